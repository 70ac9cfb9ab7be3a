// plan.cu — plan objects behind the C ABI (SURVEY §8b "what a C-ABI replacement must export":
// plan create / destroy, execute(plan, inputs, out, workspace, ws_bytes, stream), vjp, workspace_size).
//
// A statevector plan owns the device copy of its pass programs and replays them (and the per-gate steps
// the planner left unfused) with one call; its vjp is the adjoint walk over the gate list.  A
// tensor-network plan owns the launch schedule of one contraction tree (the SSA list of pairwise steps the
// host planner lowered the reference's `tree_data` to), a liveness-packed workspace layout for the
// intermediates and the slice-offset tables of the leaves; execute() runs one slice.
#include <stdint.h>

#include <algorithm>
#include <map>
#include <vector>

#include "../../include/tcb200.h"
#include "common.cuh"

namespace tcb {
int launch_dense(void*, int, int64_t, const int*, int, const void*, int64_t, cudaStream_t);
int launch_diag(void*, int, int64_t, const int*, int, const void*, int64_t, int64_t, uint64_t, cudaStream_t);
int launch_pass(const void*, void*, int, int64_t, const int32_t*, int32_t, int, int, int, const void*, int64_t,
                uint64_t, cudaStream_t);
int launch_adjoint_step(void*, void*, int, int64_t, const int*, int, const void*, int64_t, void*, int64_t,
                        cudaStream_t);
int launch_contract(const void*, int64_t, const void*, int64_t, void*, const tcb_contract_desc*, int,
                    cudaStream_t);
}  // namespace tcb

using namespace tcb;
#define S(x) reinterpret_cast<cudaStream_t>(x)
#define NOTNULL(p, fn) TCB_REQUIRE((p) != nullptr, fn ": null pointer argument `" #p "`")

// ------------------------------------------------------------------------------------------------
struct tcb_sv_plan {
  int nbits = 0;
  int32_t* d_programs = nullptr;
  struct Step {
    int kind;  // 0 fused pass, 1 dense gate, 2 diagonal gate
    int64_t prog_off;
    int prog_words, tile_bits, low_bits, pool;
    int k;
    int bitpos[8];
    int64_t mat_off, diag_stride;
  };
  struct Gate {
    int k;
    int bitpos[8];
    int64_t dense_off;
  };
  std::vector<Step> steps;
  std::vector<Gate> gates;
};

struct tcb_tn_plan {
  struct Step {
    int a, b, out;  // SSA ids: < nleaves = leaf, else intermediate
    tcb_contract_desc desc;
  };
  int nleaves = 0, nsliced = 0;
  std::vector<Step> steps;
  std::vector<int64_t> elems;          // per SSA id: element count (leaves: of the stored array)
  std::vector<int64_t> ws_off;         // per SSA id: element offset in the workspace (-1: leaf / final output)
  std::vector<std::vector<std::pair<int, int>>> leaf_slices;  // per leaf: (sliced index number, bit position)
  int64_t ws_elems = 0;
  int final_id = -1;
};

extern "C" {

// ---- statevector plans ----------------------------------------------------------------------------
int tcb_sv_plan_create(int nbits, const int32_t* programs_host, int64_t program_words, const int64_t* steps,
                       int nsteps, const int64_t* gates, int ngates, tcb_sv_plan** out) {
  NOTNULL(out, "tcb_sv_plan_create");
  TCB_REQUIRE(nbits >= 1 && nbits <= 40, "tcb_sv_plan_create: nbits=%d", nbits);
  TCB_REQUIRE(nsteps >= 0 && ngates >= 0 && program_words >= 0, "tcb_sv_plan_create: negative count");
  TCB_REQUIRE(nsteps == 0 || steps != nullptr, "tcb_sv_plan_create: null step table");
  TCB_REQUIRE(ngates == 0 || gates != nullptr, "tcb_sv_plan_create: null gate table");
  tcb_sv_plan* p = new tcb_sv_plan();
  p->nbits = nbits;
  for (int i = 0; i < nsteps; ++i) {
    const int64_t* r = steps + 16 * (size_t)i;
    tcb_sv_plan::Step s;
    s.kind = (int)r[0];
    s.prog_off = r[1];
    s.prog_words = (int)r[2];
    s.tile_bits = (int)r[3];
    s.low_bits = (int)r[4];
    s.pool = (int)r[5];
    s.k = (int)r[6];
    s.mat_off = r[7];
    s.diag_stride = r[8];
    for (int j = 0; j < 7; ++j) s.bitpos[j] = (int)r[9 + j];
    s.bitpos[7] = 0;
    const bool ok = s.kind == 0 ? (s.prog_off >= 0 && s.prog_off + s.prog_words <= program_words)
                                : ((s.kind == 1 || s.kind == 2) && s.k >= 1 && s.k <= 7 && s.mat_off >= 0);
    if (!ok) {
      delete p;
      TCB_REQUIRE(false, "tcb_sv_plan_create: malformed step %d (kind %d)", i, s.kind);
    }
    p->steps.push_back(s);
  }
  for (int i = 0; i < ngates; ++i) {
    const int64_t* r = gates + 10 * (size_t)i;
    tcb_sv_plan::Gate g;
    g.k = (int)r[0];
    g.dense_off = r[1];
    for (int j = 0; j < 7; ++j) g.bitpos[j] = (int)r[2 + j];
    g.bitpos[7] = 0;
    if (g.k < 1 || g.k > 7 || g.dense_off < 0) {
      delete p;
      TCB_REQUIRE(false, "tcb_sv_plan_create: gate %d has k=%d (1..7)", i, g.k);
    }
    p->gates.push_back(g);
  }
  if (program_words > 0) {
    NOTNULL(programs_host, "tcb_sv_plan_create");
    cudaError_t e = cudaMalloc(&p->d_programs, sizeof(int32_t) * (size_t)program_words);
    if (e == cudaSuccess)
      e = cudaMemcpy(p->d_programs, programs_host, sizeof(int32_t) * (size_t)program_words, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      if (p->d_programs) cudaFree(p->d_programs);
      delete p;
      TCB_REQUIRE(false, "tcb_sv_plan_create: %s", cudaGetErrorString(e));
    }
  }
  *out = p;
  return 0;
}

int tcb_sv_plan_destroy(tcb_sv_plan* plan) {
  if (plan) {
    if (plan->d_programs) cudaFree(plan->d_programs);
    delete plan;
  }
  return 0;
}

int64_t tcb_sv_plan_workspace_size(const tcb_sv_plan* plan) {
  (void)plan;
  return 0;  // every step works in place on the state
}

int tcb_sv_plan_execute(const tcb_sv_plan* plan, void* state, int64_t batch, const void* gatebuf,
                        int64_t gate_batch_stride, uint64_t index_base, void* stream) {
  NOTNULL(plan, "tcb_sv_plan_execute");
  NOTNULL(state, "tcb_sv_plan_execute");
  NOTNULL(gatebuf, "tcb_sv_plan_execute");
  const char* gb = reinterpret_cast<const char*>(gatebuf);
  for (const auto& s : plan->steps) {
    int rc;
    if (s.kind == 0)
      rc = launch_pass(state, state, plan->nbits, batch, plan->d_programs + s.prog_off, s.prog_words, s.tile_bits,
                       s.low_bits, s.pool, gatebuf, gate_batch_stride, index_base, S(stream));
    else if (s.kind == 1)
      rc = launch_dense(state, plan->nbits, batch, s.bitpos, s.k, gb + 8 * s.mat_off, gate_batch_stride, S(stream));
    else
      rc = launch_diag(state, plan->nbits, batch, s.bitpos, s.k, gb + 8 * s.mat_off, s.diag_stride,
                       gate_batch_stride, index_base, S(stream));
    if (rc) return rc;
  }
  return 0;
}

int tcb_sv_plan_vjp_range(const tcb_sv_plan* plan, int first_gate, int last_gate, void* lam, void* psi,
                          const void* udag, double* grad, void* stream) {
  NOTNULL(plan, "tcb_sv_plan_vjp");
  NOTNULL(lam, "tcb_sv_plan_vjp");
  NOTNULL(psi, "tcb_sv_plan_vjp");
  NOTNULL(udag, "tcb_sv_plan_vjp");
  NOTNULL(grad, "tcb_sv_plan_vjp");
  TCB_REQUIRE(first_gate >= 0 && first_gate <= last_gate && last_gate <= (int)plan->gates.size(),
              "tcb_sv_plan_vjp: gate range [%d, %d) outside the plan's %d gates", first_gate, last_gate,
              (int)plan->gates.size());
  const char* ud = reinterpret_cast<const char*>(udag);
  for (int i = last_gate - 1; i >= first_gate; --i) {
    const auto* it = &plan->gates[i];
    int rc;
    if (it->k <= 2) {
      rc = launch_adjoint_step(lam, psi, plan->nbits, 1, it->bitpos, it->k, ud + 8 * it->dense_off, 0,
                               grad + 2 * it->dense_off, 0, S(stream));
    } else {  // a constant multi-qubit gate (toffoli, fredkin ...): un-apply it on both states, no gradient
      rc = launch_dense(psi, plan->nbits, 1, it->bitpos, it->k, ud + 8 * it->dense_off, 0, S(stream));
      if (!rc) rc = launch_dense(lam, plan->nbits, 1, it->bitpos, it->k, ud + 8 * it->dense_off, 0, S(stream));
    }
    if (rc) return rc;
  }
  return 0;
}

int tcb_sv_plan_vjp(const tcb_sv_plan* plan, void* lam, void* psi, const void* udag, double* grad, void* stream) {
  NOTNULL(plan, "tcb_sv_plan_vjp");
  return tcb_sv_plan_vjp_range(plan, 0, (int)plan->gates.size(), lam, psi, udag, grad, stream);
}

int tcb_sv_plan_launches(const tcb_sv_plan* plan, int vjp) {
  if (!plan) return 0;
  if (!vjp) return (int)plan->steps.size();
  int n = 0;
  for (const auto& g : plan->gates) n += g.k <= 2 ? 1 : 2;
  return n;
}

// ---- tensor-network plans ---------------------------------------------------------------------------
int tcb_tn_plan_create(int nleaves, const int64_t* leaf_elems, int nsteps, const int32_t* step_ids,
                       const tcb_contract_desc* descs, const int64_t* out_elems, int nsliced,
                       const int32_t* leaf_slice_counts, const int32_t* leaf_slice_pairs, tcb_tn_plan** out) {
  NOTNULL(out, "tcb_tn_plan_create");
  NOTNULL(leaf_elems, "tcb_tn_plan_create");
  TCB_REQUIRE(nleaves >= 1 && nsteps >= 1, "tcb_tn_plan_create: nleaves=%d nsteps=%d", nleaves, nsteps);
  NOTNULL(step_ids, "tcb_tn_plan_create");
  NOTNULL(descs, "tcb_tn_plan_create");
  NOTNULL(out_elems, "tcb_tn_plan_create");
  TCB_REQUIRE(nsliced >= 0 && nsliced <= 62, "tcb_tn_plan_create: nsliced=%d", nsliced);
  tcb_tn_plan* p = new tcb_tn_plan();
  p->nleaves = nleaves;
  p->nsliced = nsliced;
  const int nid = nleaves + nsteps;
  p->elems.assign(nid, 0);
  p->ws_off.assign(nid, -1);
  p->leaf_slices.resize(nleaves);
  size_t pair = 0;
  for (int i = 0; i < nleaves; ++i) {
    p->elems[i] = leaf_elems[i];
    const int cnt = leaf_slice_counts ? leaf_slice_counts[i] : 0;
    for (int j = 0; j < cnt; ++j, ++pair)
      p->leaf_slices[i].push_back({leaf_slice_pairs[2 * pair], leaf_slice_pairs[2 * pair + 1]});
  }
  std::vector<char> alive(nid, 0);
  for (int i = 0; i < nleaves; ++i) alive[i] = 1;
  // first-fit free list over the workspace (element units, 32-element = 256-byte granules)
  std::map<int64_t, int64_t> free_list;  // offset -> length
  int64_t top = 0;
  auto alloc = [&](int64_t n) -> int64_t {
    n = (n + 31) & ~int64_t(31);
    for (auto it = free_list.begin(); it != free_list.end(); ++it) {
      if (it->second >= n) {
        const int64_t off = it->first, len = it->second;
        free_list.erase(it);
        if (len > n) free_list[off + n] = len - n;
        return off;
      }
    }
    // grow at the top, merging with a free block that ends there
    if (!free_list.empty()) {
      auto last = std::prev(free_list.end());
      if (last->first + last->second == top) {
        const int64_t off = last->first;
        free_list.erase(last);
        top = off + n;
        return off;
      }
    }
    const int64_t off = top;
    top += n;
    return off;
  };
  auto release = [&](int64_t off, int64_t n) {
    n = (n + 31) & ~int64_t(31);
    auto it = free_list.emplace(off, n).first;
    auto nx = std::next(it);
    if (nx != free_list.end() && it->first + it->second == nx->first) {
      it->second += nx->second;
      free_list.erase(nx);
    }
    if (it != free_list.begin()) {
      auto pv = std::prev(it);
      if (pv->first + pv->second == it->first) {
        pv->second += it->second;
        free_list.erase(it);
      }
    }
  };
  for (int s = 0; s < nsteps; ++s) {
    tcb_tn_plan::Step st;
    st.a = step_ids[3 * s];
    st.b = step_ids[3 * s + 1];
    st.out = step_ids[3 * s + 2];
    st.desc = descs[s];
    const bool ok = st.a >= 0 && st.a < nid && st.b >= 0 && st.b < nid && st.a != st.b && st.out >= nleaves &&
                    st.out < nid && alive[st.a] && alive[st.b] && !alive[st.out] && out_elems[s] >= 1;
    if (!ok) {
      delete p;
      TCB_REQUIRE(false, "tcb_tn_plan_create: step %d (%d, %d -> %d) is not a valid SSA step", s, st.a, st.b, st.out);
    }
    p->elems[st.out] = out_elems[s];
    if (s + 1 < nsteps) p->ws_off[st.out] = alloc(out_elems[s]);  // (the last step writes the caller's `out`)
    // operands die with this step (linear path: every tensor is consumed exactly once)
    for (int id : {st.a, st.b}) {
      alive[id] = 0;
      if (id >= nleaves) release(p->ws_off[id], p->elems[id]);
    }
    alive[st.out] = 1;
    p->steps.push_back(st);
  }
  p->final_id = p->steps.back().out;
  p->ws_elems = top;
  *out = p;
  return 0;
}

int tcb_tn_plan_destroy(tcb_tn_plan* plan) {
  delete plan;
  return 0;
}

int64_t tcb_tn_plan_workspace_size(const tcb_tn_plan* plan) {
  return plan ? plan->ws_elems * 8 : 0;  // bytes (complex64)
}

int64_t tcb_tn_plan_output_elems(const tcb_tn_plan* plan) { return plan ? plan->elems[plan->final_id] : 0; }

int tcb_tn_plan_execute(const tcb_tn_plan* plan, const void* const* inputs, uint64_t slice_bits, void* out,
                        void* workspace, int64_t ws_bytes, void* stream) {
  NOTNULL(plan, "tcb_tn_plan_execute");
  NOTNULL(inputs, "tcb_tn_plan_execute");
  NOTNULL(out, "tcb_tn_plan_execute");
  TCB_REQUIRE(ws_bytes >= plan->ws_elems * 8, "tcb_tn_plan_execute: workspace of %lld bytes, %lld needed",
              (long long)ws_bytes, (long long)(plan->ws_elems * 8));
  TCB_REQUIRE(plan->ws_elems == 0 || workspace != nullptr, "tcb_tn_plan_execute: null workspace");
  char* ws = reinterpret_cast<char*>(workspace);
  auto locate = [&](int id, const void*& ptr, int64_t& off) {
    off = 0;
    if (id < plan->nleaves) {
      ptr = inputs[id];
      for (const auto& pr : plan->leaf_slices[id])  // slicing folded into the leaf load: an element offset
        if ((slice_bits >> pr.first) & 1ull) off += int64_t(1) << pr.second;
    } else {
      ptr = ws + 8 * plan->ws_off[id];
    }
  };
  for (size_t s = 0; s < plan->steps.size(); ++s) {
    const auto& st = plan->steps[s];
    const void *pa, *pb;
    int64_t oa, ob;
    locate(st.a, pa, oa);
    locate(st.b, pb, ob);
    TCB_REQUIRE(pa != nullptr && pb != nullptr, "tcb_tn_plan_execute: null input tensor at step %d", (int)s);
    void* pc = s + 1 == plan->steps.size() ? out : static_cast<void*>(ws + 8 * plan->ws_off[st.out]);
    const int rc = launch_contract(pa, oa, pb, ob, pc, &st.desc, 0, S(stream));
    if (rc) return rc;
  }
  return 0;
}

int tcb_tn_plan_launches(const tcb_tn_plan* plan) { return plan ? (int)plan->steps.size() : 0; }

}  // extern "C"
