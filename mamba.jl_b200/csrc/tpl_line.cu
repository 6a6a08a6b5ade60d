// tpl_line.cu — instantiates the generic engine kernels for the `line` model template.
#include "launch.hpp"
namespace mcu {
MCU_DEFINE_TPL(LineModel)
}
