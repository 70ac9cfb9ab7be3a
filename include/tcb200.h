/*
 * tcb200.h — C ABI of the B200-native engine for TensorCircuit-NG's two
 * contraction hot paths (statevector evolution and pairwise tensor-network
 * contraction).  Plain pointers and sizes only; no torch / C++ types.
 *
 * The reference (/root/reference, pure Python) has no FFI; the seams this
 * library replaces are the *library call sites* of SURVEY.md §2.3:
 *
 *   K1  tn.contract_between -> backend.tensordot      tensorcircuit/cons.py:948 (also :396,:413,:450)
 *   K2  final_node.reorder_edges (full-state permute)  tensorcircuit/cons.py:960
 *   K3  tn.copy(nodes, conjugate=True) (bra copy)      tensorcircuit/basecircuit.py:161,:384
 *   K4  contractor([psi, psi*, P...]) expectation      tensorcircuit/circuit.py:899-902
 *   K6  tree.contract_core(sliced_arrays)              tensorcircuit/experimental.py:1008
 *   K7  tree.slice_arrays(input_arrays, slice_idx)     tensorcircuit/experimental.py:1007
 *
 * Conventions (verified against the reference's golden values):
 *   - a state is 2^n complex64 (interleaved re,im float32), qubit 0 = MOST significant bit
 *     of the flat index (tensorcircuit/circuit.py:711-719).  This ABI speaks in *bit
 *     positions* of the flat index: qubit q of an n-qubit register is bit (n-1-q).
 *   - a k-qubit gate matrix is row-major 2^k x 2^k complex64, row = (out_0..out_{k-1})_2,
 *     col = (in_0..in_{k-1})_2, first listed qubit most significant
 *     (tensorcircuit/gates.py:497-516, tensorcircuit/basecircuit.py:288-290).
 *   - all pointers are DEVICE pointers unless the name ends in _host.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *   - every entry point returns 0 on success, nonzero on error; tcb_last_error()
 *     returns a thread-local message.  Nothing falls back to the CPU.
 */
#ifndef TCB200_H
#define TCB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TCB_ABI_VERSION 1

/* ---- library ------------------------------------------------------------ */
int tcb_abi_version(void);
const char* tcb_last_error(void);
/* The tensor-core contraction path keeps its operand-image scratch (stream-ordered allocations from the device's
 * default memory pool) cached between calls; this hands it back to the driver (e.g. before a 64 GiB state is
 * allocated on the same device). */
int tcb_release_scratch(void);
/* number of SMs / total global memory of the current device */
int tcb_device_info(int* sm_count, int* cc_major, int* cc_minor, uint64_t* total_mem);

/* ---- statevector: unfused per-gate kernels (R7/R9 literal order) ---------
 * state      : [batch][2^nbits] complex64
 * bitpos[k]  : flat-index bit position of gate qubit 0..k-1 (gate qubit 0 = matrix MSB)
 * mat        : [batch or 1][2^k * 2^k] complex64 row-major; mat_batch_stride in complex elems (0 = shared)
 * index_base : OR-ed into the flat index when a control/diagonal qubit lives above the
 *              local shard (sharded statevector); 0 on one GPU.              */
int tcb_sv_init_zero(void* state, int nbits, int64_t batch, void* stream);
/* product state: state[x] = prod_p vecs[p][bit p of (x | index_base)], p < total_bits; vecs is a device
 * array of total_bits x 2 complex64 indexed by flat bit POSITION (positions >= nbits are rank bits of a
 * sharded state and are read from index_base).  One write pass.                                       */
int tcb_sv_init_product(void* state, int nbits, const void* vecs, int total_bits, uint64_t index_base,
                        void* stream);
int tcb_sv_apply_dense(void* state, int nbits, int64_t batch, const int* bitpos_host, int k,
                       const void* mat, int64_t mat_batch_stride, void* stream);
/* diagonal gate given as 2^k complex64 entries spaced `diag_stride` apart (2^k+1 for the
 * diagonal of a dense row-major matrix, 1 for a packed diagonal). bitpos may exceed nbits-1:
 * such bits are read from index_base (global qubits of a sharded state).                  */
int tcb_sv_apply_diag(void* state, int nbits, int64_t batch, const int* bitpos_host, int k,
                      const void* diag, int64_t diag_stride, int64_t mat_batch_stride,
                      uint64_t index_base, void* stream);

/* ---- statevector: fused tile pass (the hot kernel) ------------------------
 * One launch = one HBM pass over the state applying a whole *pass program*
 * (many gates).  `program` is an int32 word stream built by the host planner
 * (tensorcircuit_ng_b200/passplan.py; layout in csrc/pass_core.cuh).
 * `gatebuf` holds every gate matrix of the circuit, concatenated, complex64;
 * program ops address it by element offset.  Replaces the tensordot+permute
 * loop of tensorcircuit/cons.py:937-953 for circuit-shaped networks.          */
int tcb_sv_run_pass(void* state, int nbits, int64_t batch, const int32_t* program,
                    int32_t program_words, int tile_bits, int low_bits, int pool_elems, const void* gatebuf,
                    int64_t gate_batch_stride, uint64_t index_base, void* stream);
/* same kernel reading `src` and writing `dst` (out-of-place; src may equal dst).
 * tile_bits / low_bits / pool_elems repeat the program header's T / L / gate-pool size (the
 * host needs them to size the launch without reading device memory).                      */
int tcb_sv_run_pass_oop(const void* src, void* dst, int nbits, int64_t batch,
                        const int32_t* program, int32_t program_words, int tile_bits,
                        int low_bits, int pool_elems, const void* gatebuf, int64_t gate_batch_stride,
                        uint64_t index_base, void* stream);

/* The FIRST pass of a circuit that starts from the product state  prod_p vecs[p][x_p]  (vecs: [total_bits][2]
 * complex64 by flat bit position, as tcb_sv_init_product takes it; |0...0> is the product of (1, 0)): every tile is
 * generated in shared memory instead of loaded, so the initial state is never written to or read from HBM —
 * tcb_sv_init_product + tcb_sv_run_pass in one launch that only WRITES `dst`.  total_bits > nbits: the rank bits of
 * a sharded state, read from index_base.  Only tile_bits == 12, batch 1 (callers fall back to init + pass).      */
int tcb_sv_run_pass_generate(void* dst, int nbits, const int32_t* program, int32_t program_words, int tile_bits,
                             int low_bits, int pool_elems, const void* gatebuf, uint64_t index_base, const void* vecs,
                             int total_bits, void* stream);

/* ---- statevector: reductions (K3/K4 without materialising the bra) --------
 * out[b][t] (float64) += sum_x sign_t(x) |psi_b[x]|^2, sign = (-1)^popc((x|index_base) & zmask[t]).
 * One read of the state serves all nterms Z-strings.                          */
int tcb_sv_expect_z(const void* state, int nbits, int64_t batch, const uint64_t* zmasks,
                    int nterms, uint64_t index_base, double* out, void* stream);
/* general Pauli string P = i^{ny} X^{xmask} Z^{zmask}:  out[b] (2 x float64: re,im) +=
 * sum_x conj(psi[x ^ xmask]) * phase(x) * psi[x]   (xmask must be local: < 2^nbits)      */
int tcb_sv_expect_pauli(const void* state, int nbits, int64_t batch, uint64_t xmask,
                        uint64_t zmask, int ny, uint64_t index_base, double* out, void* stream);
/* Matrix-free Pauli-sum operator H = sum_t c_t P_t, P_t = X^{xmask_t} Z^{zmask_t} (a Y is a bit in both,
 * its i folded into c_t = w_t i^{ny_t}); replaces the COO matvec behind
 * tensorcircuit/templates/measurements.py:156-191 (operator_expectation / sparse_expectation) and the
 * term loop of tensorcircuit/quantum.py:2222-2358 (PauliStringSum2MVP):
 *   out_state[b][i] (=, or += when accumulate) sum_t c_t (-1)^popc(((i|index_base) ^ x_t) & z_t) psi_b[i ^ x_t]
 *   out_value[b]    (2 x float64: re, im) += <psi_b| H |psi_b>
 * Either output may be null.  xmask / zmask (uint64) and coef (complex64) are DEVICE arrays of nterms
 * entries, best sorted by xmask (each run of equal xmask costs one read of the state); xmask must be
 * local (< 2^nbits).  out_state must not alias state.                                          */
int tcb_sv_pauli_sum(const void* state, int nbits, int64_t batch, const uint64_t* xmask,
                     const uint64_t* zmask, const void* coef, int nterms, uint64_t index_base,
                     void* out_state, int accumulate, double* out_value, void* stream);
/* out[b] (2 x float64) += <a_b | b_b>  */
int tcb_sv_inner(const void* a, const void* b, int nbits, int64_t batch, double* out, void* stream);

/* ---- statevector: adjoint-mode vjp helpers (R17) --------------------------
 * grad[b][r][c] (complex128 as 2 x float64, +=) = sum_rest lam[rest, r] * conj(psi_in[rest, c])
 * over every amplitude group the k-qubit gate touches (k = 1, 2); grad_batch_stride in complex
 * elements.  With lam = torch's gradient of a real loss w.r.t. psi_out this is torch's gradient
 * w.r.t. the gate matrix (torch complex convention, tests/test_backends.py:992-996).        */
int tcb_sv_gate_grad(const void* lam, const void* psi_in, int nbits, int64_t batch,
                     const int* bitpos_host, int k, double* grad, int64_t grad_batch_stride,
                     void* stream);
/* One gate of the adjoint walk fused into a single pass over both states (what autograd's backward does
 * with two tensordots and saved intermediates per gate, tensorcircuit/backends/pytorch_backend.py:775-786):
 *   psi <- U^dagger psi (= psi_in);  grad[b][r][c] += sum_rest lam[rest,r] conj(psi_in[rest,c]);
 *   lam <- U^dagger lam.     udag: DEVICE, row-major 2^k x 2^k complex64 (k = 1, 2), batch stride in
 * complex elements (0 = shared); grad as in tcb_sv_gate_grad.                                        */
int tcb_sv_adjoint_step(void* lam, void* psi, int nbits, int64_t batch, const int* bitpos_host, int k,
                        const void* udag, int64_t udag_batch_stride, double* grad,
                        int64_t grad_batch_stride, void* stream);

/* Gradients of a run of consecutive DIAGONAL gates from one read of the two states (they commute: with
 * psi, lam the states after the run, dL/dd_j[c] = d_j[c] * out[j][c]):
 *   out[j][c] (complex128 pairs, =) = sum over amplitudes i whose gate-j bits read c of lam[i] conj(psi[i]),
 * c = bit_a(i) for a one-qubit gate (gate_bits = {a, -1}), (bit_a(i) << 1) | bit_b(i) for two qubits; out has
 * 4 slots per gate.  (The bins are combinations of +-1 moments; the distinct moment masks of the whole run are
 * accumulated 24 per read of the two states, in `out` itself for runs of up to 256 gates.)
 * batch: states [batch][2^nbits]; out_batch_stride (>= 4 * ngates) in complex128 elements.          */
int tcb_sv_cross_marginals(const void* lam, const void* psi, int nbits, int64_t batch, int ngates,
                           const int* gate_bits_host, double* out, int64_t out_batch_stride, void* stream);

/* Cross reduced density matrices of single qubits between two states, up to 10 qubits per read of both:
 *   out[t][r][c] (complex128 pairs, +=) = sum_rest lam[rest, bit_t = r] conj(psi[rest, bit_t = c])
 * for the bits t = 0..min(3,nbits)-1 (the lowest address bits, always included: a thread holds them in registers)
 * followed by the nsel (<= 7) selected bits (ascending, >= 3; they map onto lane and warp index bits).  With lam, psi the states BEFORE a layer of one-qubit gates on
 * distinct qubits, dL/dU_q = U_q out[q] for every gate of the layer (same convention as tcb_sv_gate_grad).
 * skip_low != 0: leave the low-bit entries untouched (a later call of a series only adds its selected bits). */
int tcb_sv_cross_rdm(const void* lam, const void* psi, int nbits, int64_t batch, int nsel, const int* sel_bits_host,
                     int skip_low, double* out, int64_t out_batch_stride, void* stream);

/* ---- statevector: sampling ------------------------------
 * Replaces probability() + cumsum + searchsorted of tensorcircuit/basecircuit.py:1490-1512 /
 * tensorcircuit/backends/abstract_backend.py:1828-1861 (and, with mode 1, the per-qubit conditional
 * draws of perfect sampling / measure_jit, basecircuit.py:449-558) without materialising p or its CDF.
 * prepare: cdf[s] (float64, 2^(nbits-seg_bits) entries) = inclusive cumulative mass of the segments of
 *          2^seg_bits amplitudes (seg_bits <= 12) — one read of the state.
 * sample : out_index[shot] (int64), out_prob[shot] (float64, |psi[index]|^2 / total; may be null).
 *          mode 0: status[shot] uniform in [0,1): first index with cumulative mass >= total (1 - u).
 *          mode 1: status[shot][nbits]: qubit j (qubit 0 first) reads 1 iff
 *                  u_j - P(bit_j = 0 | earlier bits) + 0.31415926e-12 > 0  (measure_jit, :516-531). */
int tcb_sv_sample_prepare(const void* state, int nbits, int seg_bits, double* cdf, void* stream);
int tcb_sv_sample(const void* state, int nbits, int seg_bits, const double* cdf, const double* status,
                  int64_t shots, int mode, int64_t* out_index, double* out_prob, void* stream);

/* ---- sharded statevector: local half of a global<->local qubit swap -------
 * Packs the amplitudes whose local bit `local_bit` == `want` into a contiguous
 * send buffer (and the inverse), so the exchange itself is one NCCL send/recv.           */
int tcb_sv_pack_half(const void* state, void* buf, int nbits, int local_bit, int want, void* stream);
int tcb_sv_unpack_half(void* state, const void* buf, int nbits, int local_bit, int want, void* stream);
/* m-qubit swap: the sub-block of the local state whose `nsel` selected bits (ascending positions)
 * equal `pattern` (bit k of pattern <-> sel_bits[k]) has 2^(nbits-nsel) amplitudes; elements
 * [first, first+count) of it, in block order, are copied to / from a contiguous buffer, so a swap
 * can be streamed through a small staging buffer chunk by chunk.                               */
int tcb_sv_pack_bits(const void* state, void* buf, int nbits, int nsel, const int* sel_bits_host,
                     uint64_t pattern, uint64_t first, uint64_t count, void* stream);
int tcb_sv_unpack_bits(void* state, const void* buf, int nbits, int nsel, const int* sel_bits_host,
                       uint64_t pattern, uint64_t first, uint64_t count, void* stream);

/* ---- tensor network: pairwise contraction (K1/K6/K7) ----------------------
 * All tensor modes have extent 2.  Every mode of A, B and C is described by the flat-index
 * bit position it occupies in each operand (-1 = absent).  Mode classes:
 *   batch (in A, B and C), M (A and C), N (B and C), K (A and B, summed).
 * A mode of A or B may also be *sliced* (K7): it is absent from the arrays as stored
 * *after* slicing — slicing is folded in by passing a_offset / b_offset (element offsets)
 * computed by the host from the slice id, so no sliced copies are ever made.           */
typedef struct tcb_contract_desc {
  int32_t n_batch, n_m, n_n, n_k;
  /* bit positions, most-significant-first is NOT required; any order is fine */
  int8_t batch_a[32], batch_b[32], batch_c[32];
  int8_t m_a[32], m_c[32];
  int8_t n_b[32], n_c[32];
  int8_t k_a[32], k_b[32];
  int32_t conj_a, conj_b; /* conjugate operand on load (K3 folded in) */
} tcb_contract_desc;

int tcb_tn_contract(const void* a, int64_t a_offset, const void* b, int64_t b_offset, void* c,
                    const tcb_contract_desc* desc_host, int accumulate, void* stream);

/* ---- plan objects (plan create / destroy, execute, vjp, workspace_size) ---------------
 * The host planner (passplan.py / planner.py + tnengine.build_schedule; any producer of the same tables)
 * lowers a gate stream or a `tree_data` plan (tensorcircuit/experimental.py:947-953) to the tables below;
 * the library owns the plan from then on and replays it with one call per circuit / per slice — what the
 * reference's contractor loop does pair by pair in Python (tensorcircuit/cons.py:937-953).
 *
 * Statevector plan.  programs_host: all pass programs concatenated (int32 words, copied to the device).
 * steps: nsteps x 16 int64 = {kind (0 fused pass | 1 dense gate | 2 diagonal gate), program offset,
 *   program words, tile_bits, low_bits, pool_elems, k, matrix offset in the gate buffer (elements),
 *   diagonal stride, bitpos[0..6]}.   gates: ngates x 10 int64 = {k (1..7), offset of the gate's dense
 *   2^k x 2^k block in the udag / grad buffers (elements), bitpos[0..6], 0} in program order.
 * execute: the whole circuit in place on `state`.  vjp: the adjoint walk, last gate first
 *   (tcb_sv_adjoint_step per 1- / 2-qubit gate; wider gates are constants: U^dagger on both states, no
 *   gradient): psi (final state) is un-computed to the initial state, lam becomes the cotangent of the
 *   initial state, grad (complex128 pairs, +=) receives dL/dU of every 1- / 2-qubit gate.              */
typedef struct tcb_sv_plan tcb_sv_plan;
int tcb_sv_plan_create(int nbits, const int32_t* programs_host, int64_t program_words, const int64_t* steps,
                       int nsteps, const int64_t* gates, int ngates, tcb_sv_plan** out);
int tcb_sv_plan_destroy(tcb_sv_plan* plan);
int64_t tcb_sv_plan_workspace_size(const tcb_sv_plan* plan); /* 0: every step is in place */
int tcb_sv_plan_execute(const tcb_sv_plan* plan, void* state, int64_t batch, const void* gatebuf,
                        int64_t gate_batch_stride, uint64_t index_base, void* stream);
int tcb_sv_plan_vjp(const tcb_sv_plan* plan, void* lam, void* psi, const void* udag, double* grad,
                    void* stream);
/* the same walk over gates [first_gate, last_gate) only (last first): lets the host interleave it with
 * run-level shortcuts such as tcb_sv_cross_marginals                                                   */
int tcb_sv_plan_vjp_range(const tcb_sv_plan* plan, int first_gate, int last_gate, void* lam, void* psi,
                          const void* udag, double* grad, void* stream);
int tcb_sv_plan_launches(const tcb_sv_plan* plan, int vjp); /* kernels one execute / vjp launches */

/* Tensor-network plan: one contraction tree as an SSA list of pairwise steps.  Ids 0..nleaves-1 are the
 * input tensors, nleaves+s is the result of step s.  step_ids: nsteps x {a, b, out}; descs[s]: the modes
 * of step s as bit positions of the (sliced) operands; out_elems[s]: elements of its result.  Slicing:
 * leaf_slice_counts[l] pairs {sliced index number i, bit position of that index in leaf l} per leaf (all
 * pairs concatenated in leaf_slice_pairs); execute(slice_bits) reads leaf l at element offset
 * sum_i ((slice_bits >> i) & 1) << bitpos — no sliced copies (tensorcircuit/experimental.py:999-1009).
 * Intermediates live in the caller's workspace (liveness-packed at create time); the last step writes
 * `out` (tcb_tn_plan_output_elems complex64 values).                                                   */
typedef struct tcb_tn_plan tcb_tn_plan;
int tcb_tn_plan_create(int nleaves, const int64_t* leaf_elems, int nsteps, const int32_t* step_ids,
                       const tcb_contract_desc* descs, const int64_t* out_elems, int nsliced,
                       const int32_t* leaf_slice_counts, const int32_t* leaf_slice_pairs, tcb_tn_plan** out);
int tcb_tn_plan_destroy(tcb_tn_plan* plan);
int64_t tcb_tn_plan_workspace_size(const tcb_tn_plan* plan); /* bytes */
int64_t tcb_tn_plan_output_elems(const tcb_tn_plan* plan);
int tcb_tn_plan_execute(const tcb_tn_plan* plan, const void* const* inputs, uint64_t slice_bits, void* out,
                        void* workspace, int64_t ws_bytes, void* stream);
int tcb_tn_plan_launches(const tcb_tn_plan* plan);

#ifdef __cplusplus
}
#endif
#endif /* TCB200_H */
