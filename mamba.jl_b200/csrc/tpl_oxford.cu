// tpl_oxford.cu — instantiates the generic engine kernels for the `oxford` model template (doc/examples/oxford.jl).
#define MCU_DENSITY_MATH_NOINLINE
#include "launch.hpp"
namespace mcu {
MCU_DEFINE_TPL(OxfordModel)
}
