// tpl_epil.cu — instantiates the generic engine kernels for the `epil` model template (doc/examples/epil.jl).
#define MCU_DENSITY_MATH_NOINLINE
#include "launch.hpp"
namespace mcu {
MCU_DEFINE_TPL(EpilModel)
}
