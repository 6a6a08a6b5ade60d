// tpl_surgical.cu — generic engine kernels instantiated for the `surgical` template (doc/examples/surgical.jl).
#include "launch.hpp"

namespace mcu { MCU_DEFINE_TPL(SurgicalModel) }
