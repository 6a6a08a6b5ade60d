// tpl_salm.cu — instantiates the generic engine kernels for the `salm` model template (doc/examples/salm.jl).
#define MCU_GENERIC_MINB 8
#include "launch.hpp"
namespace mcu {
MCU_DEFINE_TPL(SalmModel)
}
