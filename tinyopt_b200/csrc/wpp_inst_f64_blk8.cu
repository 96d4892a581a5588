// warp-per-problem kernels, double, 8x8 register blocks: n = 28..55 (see wpp.cuh)
#include "wpp_inst.cuh"
namespace tob200 {
TOB200_WPP_ENTRY_DEFINE(wpp_entry_f64_blk8, double, 8)
}
