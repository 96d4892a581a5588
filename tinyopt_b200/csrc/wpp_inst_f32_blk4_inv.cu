// warp-per-problem kernels, float, 4x4 register blocks, n = 13..27: the `hessian.use_ldlt = false` variants
#define TOB200_WPP_INV_TU 1
#include "wpp_inst.cuh"
namespace tob200 {
TOB200_WPP_ENTRY_DEFINE(wpp_entry_f32_blk4_inv, float, 4)
}
