// tpl_equiv.cu — instantiates the generic engine kernels for the `equiv` model template (doc/examples/equiv.jl).
#define MCU_GENERIC_MINB 8
#define MCU_DENSITY_MATH_NOINLINE
#include "launch.hpp"
namespace mcu {
MCU_DEFINE_TPL(EquivModel)
}
