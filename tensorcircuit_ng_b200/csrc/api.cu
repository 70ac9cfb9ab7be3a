// api.cu — the extern "C" surface declared in include/tcb200.h.
#include <stdarg.h>
#include <string.h>

#include "../../include/tcb200.h"
#include "common.cuh"

namespace tcb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      cached = 148;
  }
  return cached;
}

int launch_init_zero(void*, int, int64_t, cudaStream_t);
int launch_init_product(void*, int, const void*, int, uint64_t, cudaStream_t);
int launch_dense(void*, int, int64_t, const int*, int, const void*, int64_t, cudaStream_t);
int launch_diag(void*, int, int64_t, const int*, int, const void*, int64_t, int64_t, uint64_t,
                cudaStream_t);
int launch_pass_generate(void*, int, const int32_t*, int32_t, int, int, int, const void*, uint64_t, const void*, int,
                         cudaStream_t);
int launch_pass(const void*, void*, int, int64_t, const int32_t*, int32_t, int, int, int, const void*,
                int64_t, uint64_t, cudaStream_t);
int launch_expect_z(const void*, int, int64_t, const uint64_t*, int, uint64_t, double*, cudaStream_t);
int launch_expect_pauli(const void*, int, int64_t, uint64_t, uint64_t, int, uint64_t, double*,
                        cudaStream_t);
int launch_inner(const void*, const void*, int, int64_t, double*, cudaStream_t);
int launch_pauli_sum(const void*, int, int64_t, const uint64_t*, const uint64_t*, const void*, int, uint64_t,
                     void*, int, double*, cudaStream_t);
int launch_gate_grad(const void*, const void*, int, int64_t, const int*, int, void*, int64_t,
                     cudaStream_t);
int launch_adjoint_step(void*, void*, int, int64_t, const int*, int, const void*, int64_t, void*, int64_t,
                        cudaStream_t);
int launch_cross_marginals(const void*, const void*, int, int64_t, int, const int*, double*, int64_t, cudaStream_t);
int launch_cross_rdm(const void*, const void*, int, int64_t, int, const int*, int, double*, int64_t, cudaStream_t);
int launch_sample_prepare(const void*, int, int, double*, cudaStream_t);
int launch_sample(const void*, int, int, const double*, const double*, int64_t, int, long long*, double*,
                  cudaStream_t);
int launch_pack_half(const void*, void*, int, int, int, int, cudaStream_t);
int launch_pack_bits(void*, void*, int, int, const int*, uint64_t, uint64_t, uint64_t, int, cudaStream_t);
int launch_contract(const void*, int64_t, const void*, int64_t, void*, const tcb_contract_desc*, int,
                    cudaStream_t);

}  // namespace tcb

using namespace tcb;
#define S(x) reinterpret_cast<cudaStream_t>(x)
#define NOTNULL(p, fn) TCB_REQUIRE((p) != nullptr, fn ": null pointer argument `" #p "`")

extern "C" {

int tcb_abi_version(void) { return TCB_ABI_VERSION; }
const char* tcb_last_error(void) { return g_err; }
int tcb_release_scratch(void) {
  int dev = 0;
  cudaMemPool_t pool;
  TCB_CHECK_CUDA(cudaGetDevice(&dev));
  TCB_CHECK_CUDA(cudaDeviceSynchronize());
  TCB_CHECK_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
  TCB_CHECK_CUDA(cudaMemPoolTrimTo(pool, 0));
  return 0;
}

int tcb_device_info(int* sm, int* cc_major, int* cc_minor, uint64_t* total_mem) {
  int dev = 0;
  TCB_CHECK_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  TCB_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (sm) *sm = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  if (total_mem) *total_mem = (uint64_t)prop.totalGlobalMem;
  return 0;
}

int tcb_sv_init_zero(void* state, int nbits, int64_t batch, void* stream) {
  NOTNULL(state, "tcb_sv_init_zero");
  return launch_init_zero(state, nbits, batch, S(stream));
}

int tcb_sv_apply_dense(void* state, int nbits, int64_t batch, const int* bitpos, int k,
                       const void* mat, int64_t mat_batch_stride, void* stream) {
  NOTNULL(state, "tcb_sv_apply_dense");
  NOTNULL(bitpos, "tcb_sv_apply_dense");
  NOTNULL(mat, "tcb_sv_apply_dense");
  return launch_dense(state, nbits, batch, bitpos, k, mat, mat_batch_stride, S(stream));
}

int tcb_sv_apply_diag(void* state, int nbits, int64_t batch, const int* bitpos, int k,
                      const void* diag, int64_t diag_stride, int64_t mat_batch_stride,
                      uint64_t index_base, void* stream) {
  NOTNULL(state, "tcb_sv_apply_diag");
  NOTNULL(bitpos, "tcb_sv_apply_diag");
  NOTNULL(diag, "tcb_sv_apply_diag");
  return launch_diag(state, nbits, batch, bitpos, k, diag, diag_stride, mat_batch_stride, index_base,
                     S(stream));
}

int tcb_sv_init_product(void* state, int nbits, const void* vecs, int total_bits, uint64_t index_base,
                        void* stream) {
  NOTNULL(state, "tcb_sv_init_product");
  NOTNULL(vecs, "tcb_sv_init_product");
  return launch_init_product(state, nbits, vecs, total_bits, index_base, S(stream));
}

int tcb_sv_run_pass(void* state, int nbits, int64_t batch, const int32_t* program,
                    int32_t program_words, int tile_bits, int low_bits, int pool_elems,
                    const void* gatebuf, int64_t gate_batch_stride, uint64_t index_base, void* stream) {
  NOTNULL(state, "tcb_sv_run_pass");
  NOTNULL(program, "tcb_sv_run_pass");
  NOTNULL(gatebuf, "tcb_sv_run_pass");
  return launch_pass(state, state, nbits, batch, program, program_words, tile_bits, low_bits, pool_elems,
                     gatebuf, gate_batch_stride, index_base, S(stream));
}

int tcb_sv_run_pass_oop(const void* src, void* dst, int nbits, int64_t batch, const int32_t* program,
                        int32_t program_words, int tile_bits, int low_bits, int pool_elems,
                        const void* gatebuf, int64_t gate_batch_stride, uint64_t index_base,
                        void* stream) {
  NOTNULL(src, "tcb_sv_run_pass_oop");
  NOTNULL(dst, "tcb_sv_run_pass_oop");
  NOTNULL(program, "tcb_sv_run_pass_oop");
  NOTNULL(gatebuf, "tcb_sv_run_pass_oop");
  return launch_pass(src, dst, nbits, batch, program, program_words, tile_bits, low_bits, pool_elems,
                     gatebuf, gate_batch_stride, index_base, S(stream));
}

int tcb_sv_run_pass_generate(void* dst, int nbits, const int32_t* program, int32_t program_words, int tile_bits,
                             int low_bits, int pool_elems, const void* gatebuf, uint64_t index_base, const void* vecs,
                             int total_bits, void* stream) {
  NOTNULL(dst, "tcb_sv_run_pass_generate");
  NOTNULL(program, "tcb_sv_run_pass_generate");
  NOTNULL(gatebuf, "tcb_sv_run_pass_generate");
  NOTNULL(vecs, "tcb_sv_run_pass_generate");
  return launch_pass_generate(dst, nbits, program, program_words, tile_bits, low_bits, pool_elems, gatebuf, index_base,
                              vecs, total_bits, S(stream));
}

int tcb_sv_expect_z(const void* state, int nbits, int64_t batch, const uint64_t* zmasks, int nterms,
                    uint64_t index_base, double* out, void* stream) {
  NOTNULL(state, "tcb_sv_expect_z");
  NOTNULL(zmasks, "tcb_sv_expect_z");
  NOTNULL(out, "tcb_sv_expect_z");
  return launch_expect_z(state, nbits, batch, zmasks, nterms, index_base, out, S(stream));
}

int tcb_sv_expect_pauli(const void* state, int nbits, int64_t batch, uint64_t xmask, uint64_t zmask,
                        int ny, uint64_t index_base, double* out, void* stream) {
  NOTNULL(state, "tcb_sv_expect_pauli");
  NOTNULL(out, "tcb_sv_expect_pauli");
  return launch_expect_pauli(state, nbits, batch, xmask, zmask, ny, index_base, out, S(stream));
}

int tcb_sv_pauli_sum(const void* state, int nbits, int64_t batch, const uint64_t* xmask,
                     const uint64_t* zmask, const void* coef, int nterms, uint64_t index_base,
                     void* out_state, int accumulate, double* out_value, void* stream) {
  NOTNULL(state, "tcb_sv_pauli_sum");
  if (nterms > 0) {
    NOTNULL(xmask, "tcb_sv_pauli_sum");
    NOTNULL(zmask, "tcb_sv_pauli_sum");
    NOTNULL(coef, "tcb_sv_pauli_sum");
  }
  return launch_pauli_sum(state, nbits, batch, xmask, zmask, coef, nterms, index_base, out_state,
                          accumulate, out_value, S(stream));
}

int tcb_sv_inner(const void* a, const void* b, int nbits, int64_t batch, double* out, void* stream) {
  NOTNULL(a, "tcb_sv_inner");
  NOTNULL(b, "tcb_sv_inner");
  NOTNULL(out, "tcb_sv_inner");
  return launch_inner(a, b, nbits, batch, out, S(stream));
}

int tcb_sv_gate_grad(const void* lam, const void* psi_in, int nbits, int64_t batch, const int* bitpos,
                     int k, double* grad, int64_t grad_batch_stride, void* stream) {
  NOTNULL(lam, "tcb_sv_gate_grad");
  NOTNULL(psi_in, "tcb_sv_gate_grad");
  NOTNULL(bitpos, "tcb_sv_gate_grad");
  NOTNULL(grad, "tcb_sv_gate_grad");
  return launch_gate_grad(lam, psi_in, nbits, batch, bitpos, k, grad, grad_batch_stride, S(stream));
}

int tcb_sv_adjoint_step(void* lam, void* psi, int nbits, int64_t batch, const int* bitpos_host, int k,
                        const void* udag, int64_t udag_batch_stride, double* grad,
                        int64_t grad_batch_stride, void* stream) {
  NOTNULL(lam, "tcb_sv_adjoint_step");
  NOTNULL(psi, "tcb_sv_adjoint_step");
  NOTNULL(bitpos_host, "tcb_sv_adjoint_step");
  NOTNULL(udag, "tcb_sv_adjoint_step");
  NOTNULL(grad, "tcb_sv_adjoint_step");
  return launch_adjoint_step(lam, psi, nbits, batch, bitpos_host, k, udag, udag_batch_stride, grad,
                             grad_batch_stride, S(stream));
}

int tcb_sv_cross_marginals(const void* lam, const void* psi, int nbits, int64_t batch, int ngates,
                           const int* gate_bits_host, double* out, int64_t out_batch_stride, void* stream) {
  NOTNULL(lam, "tcb_sv_cross_marginals");
  NOTNULL(psi, "tcb_sv_cross_marginals");
  NOTNULL(out, "tcb_sv_cross_marginals");
  if (ngates > 0) NOTNULL(gate_bits_host, "tcb_sv_cross_marginals");
  return launch_cross_marginals(lam, psi, nbits, batch, ngates, gate_bits_host, out, out_batch_stride, S(stream));
}

int tcb_sv_cross_rdm(const void* lam, const void* psi, int nbits, int64_t batch, int nsel, const int* sel_bits_host,
                     int skip_low, double* out, int64_t out_batch_stride, void* stream) {
  NOTNULL(lam, "tcb_sv_cross_rdm");
  NOTNULL(psi, "tcb_sv_cross_rdm");
  NOTNULL(out, "tcb_sv_cross_rdm");
  if (nsel > 0) NOTNULL(sel_bits_host, "tcb_sv_cross_rdm");
  return launch_cross_rdm(lam, psi, nbits, batch, nsel, sel_bits_host, skip_low, out, out_batch_stride, S(stream));
}

int tcb_sv_sample_prepare(const void* state, int nbits, int seg_bits, double* cdf, void* stream) {
  NOTNULL(state, "tcb_sv_sample_prepare");
  NOTNULL(cdf, "tcb_sv_sample_prepare");
  return launch_sample_prepare(state, nbits, seg_bits, cdf, S(stream));
}

int tcb_sv_sample(const void* state, int nbits, int seg_bits, const double* cdf, const double* status,
                  int64_t shots, int mode, int64_t* out_index, double* out_prob, void* stream) {
  NOTNULL(state, "tcb_sv_sample");
  NOTNULL(cdf, "tcb_sv_sample");
  NOTNULL(status, "tcb_sv_sample");
  NOTNULL(out_index, "tcb_sv_sample");
  return launch_sample(state, nbits, seg_bits, cdf, status, shots, mode,
                       reinterpret_cast<long long*>(out_index), out_prob, S(stream));
}

int tcb_sv_pack_half(const void* state, void* buf, int nbits, int local_bit, int want, void* stream) {
  NOTNULL(state, "tcb_sv_pack_half");
  NOTNULL(buf, "tcb_sv_pack_half");
  return launch_pack_half(state, buf, nbits, local_bit, want, 0, S(stream));
}

int tcb_sv_unpack_half(void* state, const void* buf, int nbits, int local_bit, int want, void* stream) {
  NOTNULL(state, "tcb_sv_unpack_half");
  NOTNULL(buf, "tcb_sv_unpack_half");
  return launch_pack_half(state, const_cast<void*>(buf), nbits, local_bit, want, 1, S(stream));
}

int tcb_sv_pack_bits(const void* state, void* buf, int nbits, int nsel, const int* sel_bits_host,
                     uint64_t pattern, uint64_t first, uint64_t count, void* stream) {
  NOTNULL(state, "tcb_sv_pack_bits");
  NOTNULL(buf, "tcb_sv_pack_bits");
  NOTNULL(sel_bits_host, "tcb_sv_pack_bits");
  return launch_pack_bits(const_cast<void*>(state), buf, nbits, nsel, sel_bits_host, pattern, first, count, 0,
                          S(stream));
}

int tcb_sv_unpack_bits(void* state, const void* buf, int nbits, int nsel, const int* sel_bits_host,
                       uint64_t pattern, uint64_t first, uint64_t count, void* stream) {
  NOTNULL(state, "tcb_sv_unpack_bits");
  NOTNULL(buf, "tcb_sv_unpack_bits");
  NOTNULL(sel_bits_host, "tcb_sv_unpack_bits");
  return launch_pack_bits(state, const_cast<void*>(buf), nbits, nsel, sel_bits_host, pattern, first, count, 1,
                          S(stream));
}

int tcb_tn_contract(const void* a, int64_t a_offset, const void* b, int64_t b_offset, void* c,
                    const tcb_contract_desc* desc, int accumulate, void* stream) {
  NOTNULL(a, "tcb_tn_contract");
  NOTNULL(b, "tcb_tn_contract");
  NOTNULL(c, "tcb_tn_contract");
  return launch_contract(a, a_offset, b, b_offset, c, desc, accumulate, S(stream));
}

}  // extern "C"
