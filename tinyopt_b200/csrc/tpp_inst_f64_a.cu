// thread-per-problem kernels, double, n = 1..4 (see tpp.cuh)
#include "tpp_inst.cuh"
namespace tob200 {
TOB200_TPP_ENTRY_DEFINE(tpp_entry_f64_a, double, 1, 4)
}
