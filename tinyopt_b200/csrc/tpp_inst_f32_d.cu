// thread-per-problem kernels, float, n = 11..12 (see tpp.cuh)
#include "tpp_inst.cuh"
namespace tob200 {
TOB200_TPP_ENTRY_DEFINE(tpp_entry_f32_d, float, 11, 12)
}
