// pass_emu.cpp — TEST INFRASTRUCTURE ONLY (never loaded by the tensorcircuit_ng_b200 package).
// Compiles the host/device-neutral pass interpreter (csrc/pass_core.cuh) with g++ and runs the
// kernel's phase structure sequentially, so that the planner's programs and the kernel's index
// arithmetic are checked on CPU-only CI.  It mirrors pass_kernel.cu phase by phase.
#include <cstdint>
#include <algorithm>
#include <cstring>
#include <vector>

#include "../../tensorcircuit_ng_b200/csrc/pass_core.cuh"

using namespace tcb;

extern "C" int emu_run_pass(float* state_f, int nbits, long long batch, const int32_t* prog,
                            int prog_words, int tile_bits, int low_bits, const float* gatebuf_f,
                            long long gate_bstride, unsigned long long index_base) {
  const int32_t* hdr = prog;  // re-pointed at the staged copy below
  if (hdr[H_MAGIC] != PASS_MAGIC) return 10;
  const int T = hdr[H_T], L = hdr[H_L];
  if (T != tile_bits || L != low_bits) return 11;
  if (hdr[H_WORDS] != prog_words) return 12;
  if (hdr[H_NNONTILE] != nbits - T) return 13;
  const int nthreads = 256;
  const uint64_t tiles = 1ull << (nbits - T);
  std::vector<float4> tile4(1u << (T - 1));  // 16-byte aligned
  float2* tile = reinterpret_cast<float2*>(tile4.data());
  float2* state = reinterpret_cast<float2*>(state_f);
  const float2* gatebuf = reinterpret_cast<const float2*>(gatebuf_f);
  std::vector<int32_t> sprog(prog, prog + prog_words);  // per-"CTA" staged copy of the program
  for (long long b = 0; b < batch; ++b) {
    float2* st = state + ((size_t)b << nbits);
    const float2* gates = gatebuf + (size_t)b * gate_bstride;
    hdr = sprog.data();
    for (uint64_t tile_id = 0; tile_id < tiles; ++tile_id) {
      const uint64_t base = tile_base(tile_id, hdr);
      {  // per-tile prologue (pass_kernel.cu): gate pool + fill records on a private program copy
        std::copy(prog, prog + prog_words, sprog.begin());
        const int nfill = sprog[H_NFILL], npool = sprog[H_NPOOL];
        const int32_t* filltab = sprog.data() + prog_words - nfill;
        const int32_t* pooltab = filltab - 3 * npool;
        std::vector<float2> pool(sprog[H_POOLSIZE] + 1);
        for (int e = 0; e < npool; ++e)
          for (int i = 0; i < pooltab[3 * e + 1]; ++i) pool[pooltab[3 * e + 2] + i] = gates[pooltab[3 * e] + i];
        for (int i = 0; i < nfill; ++i) run_fill_record(sprog.data(), filltab[i], pool.data(), base | index_base);
      }
      for (int t = 0; t < (1 << T); ++t) tile[swz(t)] = st[base | tile_to_flat(t, hdr)];
      const int32_t* sp = hdr + HDR_WORDS;
      for (int s = 0; s < hdr[H_NSUB]; ++s) {
        if (sp[S_KIND] == SUB_REG) {
          const int ngroups = 1 << (T - PASS_R);
          for (int g = 0; g < ngroups; ++g)
            run_reg_subpass<PASS_R>(tile, sp, group_to_tile(g, T, PASS_R, sp), base | index_base);
        } else {
          for (int tid = 0; tid < nthreads; ++tid)
            run_smem_dense(tile, hdr, sp, gates, tid, nthreads);
        }
        sp += sp[S_WORDS];
      }
      for (int t = 0; t < (1 << T); ++t) st[base | tile_to_flat(t, hdr)] = tile[swz(t)];
    }
  }
  return 0;
}
