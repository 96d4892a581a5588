// thread-per-problem kernels, float, n = 9..10 (see tpp.cuh)
#include "tpp_inst.cuh"
namespace tob200 {
TOB200_TPP_ENTRY_DEFINE(tpp_entry_f32_c, float, 9, 10)
}
