// tpl_blocker.cu — instantiates the generic engine kernels for the `blocker` model template (doc/examples/blocker.jl).
#define MCU_GENERIC_MINB 6
#define MCU_DENSITY_MATH_NOINLINE
#include "launch.hpp"
namespace mcu {
MCU_DEFINE_TPL(BlockerModel)
}
