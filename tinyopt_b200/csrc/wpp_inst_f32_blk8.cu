// warp-per-problem kernels, float, 8x8 register blocks: n = 28..55 (see wpp.cuh)
#include "wpp_inst.cuh"
namespace tob200 {
TOB200_WPP_ENTRY_DEFINE(wpp_entry_f32_blk8, float, 8)
}
