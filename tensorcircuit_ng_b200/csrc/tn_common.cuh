// tn_common.cuh — operand description shared by the pairwise-contraction kernels.
#pragma once
#include <stdint.h>

namespace tcb {

// logical index convention: bit i of m <-> m_a[i] / m_c[i], etc. (host lists modes in any order)
struct ContractParams {
  int nb, nm, nn, nk;
  int8_t batch_a[32], batch_b[32], batch_c[32];
  int8_t m_a[32], m_c[32];
  int8_t n_b[32], n_c[32];
  int8_t k_a[32], k_b[32];
  int conj_a, conj_b, accumulate;
};

}  // namespace tcb
