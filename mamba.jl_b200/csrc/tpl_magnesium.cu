// tpl_magnesium.cu — instantiates the generic engine kernels for the `magnesium` model template (doc/examples/magnesium.jl).
#define MCU_DENSITY_MATH_NOINLINE
#include "launch.hpp"
namespace mcu {
MCU_DEFINE_TPL(MagnesiumModel)
}
