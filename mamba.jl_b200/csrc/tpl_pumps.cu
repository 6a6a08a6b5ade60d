// tpl_pumps.cu — instantiates the generic engine kernels for the `pumps` model template.
#include "launch.hpp"
namespace mcu {
MCU_DEFINE_TPL(PumpsModel)
}
