// tpl_glm.cu — instantiates the generic engine kernels for the `glm` model template.
#include "launch.hpp"
namespace mcu {
MCU_DEFINE_TPL(GlmM)
}
