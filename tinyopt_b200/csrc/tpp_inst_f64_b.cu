// thread-per-problem kernels, double, n = 5..6 (see tpp.cuh)
#include "tpp_inst.cuh"
namespace tob200 {
TOB200_TPP_ENTRY_DEFINE(tpp_entry_f64_b, double, 5, 6)
}
