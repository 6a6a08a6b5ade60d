// tpl_rats.cu — instantiates the generic engine kernels for the `rats` model template.
#include "launch.hpp"
namespace mcu {
MCU_DEFINE_TPL(RatsModel)
}
