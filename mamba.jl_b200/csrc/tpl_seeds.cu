// tpl_seeds.cu — instantiates the generic engine kernels for the `seeds` model template.
#include "launch.hpp"
namespace mcu {
MCU_DEFINE_TPL(SeedsModel)
}
