// warp-per-problem kernels, float, 4x4 register blocks: n = 13..27 (see wpp.cuh)
#include "wpp_inst.cuh"
namespace tob200 {
TOB200_WPP_ENTRY_DEFINE(wpp_entry_f32_blk4, float, 4)
}
