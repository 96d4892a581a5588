// warp-per-problem kernels, double, 8x8 register blocks, n = 28..55: the `hessian.use_ldlt = false` variants
#define TOB200_WPP_INV_TU 1
#include "wpp_inst.cuh"
namespace tob200 {
TOB200_WPP_ENTRY_DEFINE(wpp_entry_f64_blk8_inv, double, 8)
}
