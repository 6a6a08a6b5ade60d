// tpl_dyes.cu — generic engine kernels instantiated for the `dyes` template (doc/examples/dyes.jl).
// Resident blocks per SM the register allocation aims for (profiles/r1_generic_kernel_occupancy.md): small state record.
#define MCU_GENERIC_MINB 8
#include "launch.hpp"

namespace mcu { MCU_DEFINE_TPL(DyesModel) }
