// seeds_fast.cu — placeholder; replaced by the fused seeds/AMWG kernel.
#include "launch.hpp"
namespace mcu {
int seeds_fast_launch(const SeedsModel::Data&, const RunArgs&, const DevBlock*, cudaStream_t) { return -1; }
}
