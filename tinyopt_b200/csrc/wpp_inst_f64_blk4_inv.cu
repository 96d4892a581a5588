// warp-per-problem kernels, double, 4x4 register blocks, n = 9..27: the `hessian.use_ldlt = false` variants
#define TOB200_WPP_INV_TU 1
#include "wpp_inst.cuh"
namespace tob200 {
TOB200_WPP_ENTRY_DEFINE3(wpp_entry_f64_blk4_inv, double, 4)
}
