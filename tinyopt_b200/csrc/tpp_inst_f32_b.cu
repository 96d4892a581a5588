// thread-per-problem kernels, float, n = 5..8 (see tpp.cuh)
#include "tpp_inst.cuh"
namespace tob200 {
TOB200_TPP_ENTRY_DEFINE(tpp_entry_f32_b, float, 5, 8)
}
