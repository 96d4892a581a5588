// thread-per-problem kernels, float, n = 1..4 (see tpp.cuh)
#include "tpp_inst.cuh"
namespace tob200 {
TOB200_TPP_ENTRY_DEFINE(tpp_entry_f32_a, float, 1, 4)
}
