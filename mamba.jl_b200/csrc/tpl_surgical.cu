// tpl_surgical.cu — generic engine kernels instantiated for the `surgical` template (doc/examples/surgical.jl).
// Resident blocks per SM the register allocation aims for (measured on B200, 128-thread blocks, profiles/r1_generic_kernel_occupancy.md):
// small state records want full occupancy (latency hiding beats spills), large ones (rats) want registers.
#define MCU_GENERIC_MINB 8
#define MCU_DENSITY_MATH_NOINLINE
#include "launch.hpp"

namespace mcu { MCU_DEFINE_TPL(SurgicalModel) }
