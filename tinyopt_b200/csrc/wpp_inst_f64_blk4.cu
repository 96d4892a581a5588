// warp-per-problem kernels, double, 4x4 register blocks: n = 9..27 (see wpp.cuh)
#include "wpp_inst.cuh"
namespace tob200 {
TOB200_WPP_ENTRY_DEFINE3(wpp_entry_f64_blk4, double, 4)
}
