// tpl_stacks.cu — instantiates the generic engine kernels for the `stacks` model template (doc/examples/stacks.jl).
#define MCU_GENERIC_MINB 12
#include "launch.hpp"
namespace mcu {
MCU_DEFINE_TPL(StacksModel)
}
