// pass_kernel_body.cuh — the body of the fused tile-pass kernel, included twice by pass_kernel.cu:
//   PASS_KERNEL_NAME = pass_kernel                               : the hot kernel (load -> sub-passes -> store)
//   PASS_KERNEL_NAME = pass_kernel_generate, PASS_GENERATE set   : the FIRST pass of a circuit that starts from a
//       product state: the tile is generated in shared memory instead of loaded, so |psi_0> is never written to
//       or read from HBM.  A textual second copy, not a template flag: the hot kernel's code generation must not
//       move (profiles/r2_pass_kernel_experiments.md).
// R register bits, LT = log2(threads per CTA); tile bits T = LT + R are compile-time, so every
// swizzle constant of the streaming phases is a literal.
template <int R, int LT, int MINB>
__global__ void __launch_bounds__(1 << LT, MINB) PASS_KERNEL_NAME(const PassArgs A PASS_EXTRA_PARAM) {
  constexpr int T = LT + R;
  constexpr int NT = 1 << LT;
  constexpr int N_IO = (1 << (T - 1)) / NT;  // 16-byte chunks per thread (load and store)
  constexpr int NWARPS = (NT + 31) / 32;
  static_assert(NWARPS <= 8, "grp table holds 8 warp parts");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;

  // ---- once per CTA: carve shared memory, stage the program, build the static tables ----
  const int L = __ldg(A.prog + H_L);
  const int poolsize = __ldg(A.prog + H_POOLSIZE);
  const PassSmem lay = pass_smem_layout(T, L, A.prog_words, poolsize);
  float2* tile = reinterpret_cast<float2*>(smem_raw + lay.tile);
  uint32_t* hi_flat = reinterpret_cast<uint32_t*>(smem_raw + lay.hi_flat);  // tile_to_flat(h << L) >> L
  int32_t* grp = reinterpret_cast<int32_t*>(smem_raw + lay.grp);
  float2* pool = reinterpret_cast<float2*>(smem_raw + lay.pool);
  int32_t* sprog = reinterpret_cast<int32_t*>(smem_raw + lay.prog);
  for (int w = tid; w < A.prog_words; w += NT) sprog[w] = __ldg(A.prog + w);
  __syncthreads();
  const int32_t* hdr = sprog;
  for (int h = tid; h < (1 << (T - L)); h += NT) hi_flat[h] = (uint32_t)(tile_to_flat(h << L, hdr) >> L);
  // tile number -> CTA-constant flat bits through 6-bit lookup tables (the bit-deposit loop over the 18+ non-tile
  // bits ran once per tile in every thread)
  __shared__ unsigned long long tb_lut[7][64];
  const int tb_chunks = (hdr[H_NNONTILE] + 5) / 6;
  for (int e = tid; e < tb_chunks * 64; e += NT) {
    const int c = e >> 6, v = e & 63;
    unsigned long long g = 0;
    for (int i = 0; i < 6; ++i)
      if (c * 6 + i < hdr[H_NNONTILE]) g |= (unsigned long long)((v >> i) & 1) << hdr[H_NONTILEPOS + c * 6 + i];
    tb_lut[c][v] = g;
  }
#ifdef PASS_GENERATE
  // product-state source: amplitude(x) = prod_p vecs[p][x_p] factorises into a CTA-constant part (a_lut, one
  // table per 6 non-tile bits; the rank bits of a sharded state ride in g_lo) and a tile part g_lo[t & 63] g_hi[t >> 6]
  __shared__ __align__(16) float2 g_lo[64];
  __shared__ float2 g_hi[128];
  __shared__ float2 a_lut[7][64];
  {
    float2 g0 = c_one();
    for (int p = A.nbits; p < GEN.total_bits; ++p) g0 = cmul(g0, GEN.vecs[2 * p + (int)((A.index_base >> p) & 1ull)]);
    for (int e = tid; e < 64 + 128 + tb_chunks * 64; e += NT) {
      float2 acc = c_one();
      if (e < 64) {
        for (int i = 0; i < 6 && i < T; ++i) acc = cmul(acc, GEN.vecs[2 * hdr[H_TILEPOS + i] + ((e >> i) & 1)]);
        g_lo[e] = cmul(acc, g0);
      } else if (e < 192) {
        const int v = e - 64;
        for (int i = 6; i < T; ++i) acc = cmul(acc, GEN.vecs[2 * hdr[H_TILEPOS + i] + ((v >> (i - 6)) & 1)]);
        g_hi[v] = acc;
      } else {
        const int c = (e - 192) >> 6, v = (e - 192) & 63;
        for (int i = 0; i < 6; ++i)
          if (c * 6 + i < hdr[H_NNONTILE]) acc = cmul(acc, GEN.vecs[2 * hdr[H_NONTILEPOS + c * 6 + i] + ((v >> i) & 1)]);
        a_lut[c][v] = acc;
      }
    }
  }
#endif
  const int npool = hdr[H_NPOOL];
  const int nfill = hdr[H_NFILL];
  const int nstatic = hdr[H_NFILL_STATIC];
  const int nshort_end = hdr[H_NFILL_SHORT_END];
  const int nsub = hdr[H_NSUB];
  const int32_t* pooltab = sprog + A.prog_words - nfill - 3 * npool;
  {
    // tile index of a thread = lane part | warp part, per register sub-pass
    const int32_t* sp = hdr + HDR_WORDS;
    for (int s = 0; s < nsub; ++s) {
      if (sp[S_KIND] == SUB_REG) {
        for (int i = tid; i < GRP_STRIDE; i += NT) {
          const int g = i < 32 ? i : ((i - 32) << 5);
          grp[s * GRP_STRIDE + i] = group_to_tile(g & (NT - 1), T, R, sp);
        }
      }
      sp += sp[S_WORDS];
    }
  }

  // element mapping of the streaming phases: chunk c = tid + NT*u holds amplitudes t = 2c, 2c+1
  // (16 bytes per lane: a warp instruction covers 512 contiguous bytes of the tile index space).
  // tid and NT*u have disjoint bits, so swizzle and hi_flat index split into thread + literal parts.
  const int lowmask = (1 << L) - 1;
  const int io_s0 = swz(2 * tid);
  const unsigned long long tps_mask = (1ull << A.log_tiles_per_state) - 1ull;
  const uint32_t* hf = hi_flat + ((2 * tid) >> L);
  const int hstep = (2 * NT) >> L;  // hi_flat entries per u step
  __syncthreads();                  // hi_flat and grp ready

  if (A.stagger_ns > 0) {  // experiment: de-phase the CTAs that share an SM
    const int phase = blockIdx.x / A.n_sm;
    for (int i = 0; i < phase; ++i) __nanosleep(A.stagger_ns);
  }
  TRACE_DECL
  long long cur_batch = -1;
  for (unsigned long long tg = blockIdx.x; tg < A.total_tiles; tg += gridDim.x) {
    const long long b = (long long)(tg >> A.log_tiles_per_state);
    uint64_t base = 0;  // = tile_base(tg & tps_mask, hdr)
    {
      const unsigned long long tnum = tg & tps_mask;
      for (int c = 0; c < tb_chunks; ++c) base |= tb_lut[c][(tnum >> (6 * c)) & 63ull];
    }
    const uint64_t cta_bits = base | A.index_base;
    const float2* gates = A.gatebuf + (size_t)b * A.gate_bstride;
    PROF_DECL
    TRACE(0);
#ifndef PASS_GENERATE
    // ---- load: 16-byte cp.async straight into the swizzled tile (no registers held) ----
    {
      const float2* src_b = A.src + ((size_t)b << A.nbits) + (base | (uint64_t)((2 * tid) & lowmask));
      uint64_t off[N_IO];
#pragma unroll
      for (int u = 0; u < N_IO; ++u) off[u] = (uint64_t)hf[u * hstep] << L;
#pragma unroll
      for (int u = 0; u < N_IO; ++u) cp_async16(tile + (io_s0 ^ swz(2 * NT * u)), src_b + off[u]);
      cp_async_commit();
    }
#else
    // ---- no load: the tile of the initial product state is generated in place (no write + read of |psi_0>) ----
    {
      const unsigned long long tnum = tg & tps_mask;
      float2 amp = c_one();
      for (int c = 0; c < tb_chunks; ++c) amp = cmul(amp, a_lut[c][(tnum >> (6 * c)) & 63ull]);
      const float4 lo = *reinterpret_cast<const float4*>(&g_lo[(2 * tid) & 63]);  // amplitudes t = 2c and 2c + 1
#pragma unroll
      for (int u = 0; u < N_IO; ++u) {
        const float2 ah = cmul(amp, g_hi[(2 * (tid + NT * u)) >> 6]);
        const float2 v0 = cmul(ah, make_float2(lo.x, lo.y)), v1 = cmul(ah, make_float2(lo.z, lo.w));
        *reinterpret_cast<float4*>(tile + (io_s0 ^ swz(2 * NT * u))) = make_float4(v0.x, v0.y, v1.x, v1.y);
      }
    }
#endif
    PROF_MARK(0);
    // ---- while the tile streams in: resolve the fill records (shared memory only) ----
    if (b != cur_batch) {  // (re)stage the gate pool of this batch element
      for (int e = 0; e < npool; ++e) {
        const int goff = pooltab[3 * e], cnt = pooltab[3 * e + 1], poff = pooltab[3 * e + 2];
        for (int i = tid; i < cnt; i += NT) pool[poff + i] = gates[goff + i];
      }
      cur_batch = b;
      __syncthreads();
      // records that do not depend on tile bits (fused 1q products, tables, ...): once per batch element
      prologue_fill_warps(sprog, A.prog_words, pool, 0, warp, lane, NWARPS, 0, nstatic);
    }
    prologue_fill_threads(sprog, A.prog_words, pool, cta_bits, tid, NT, nstatic, nshort_end);
    prologue_fill_warps(sprog, A.prog_words, pool, cta_bits, warp, lane, NWARPS, nshort_end, nfill);
    PROF_MARK(1);
    TRACE(1);
    cp_async_wait_all();
    __syncthreads();
    PROF_MARK(2);
    TRACE(2);

    // ---- sub-passes ----
    const int32_t* sp = hdr + HDR_WORDS;
    for (int s = 0; s < nsub; ++s) {
      if (sp[S_KIND] == SUB_REG) {
        const int tbase = grp[s * GRP_STRIDE + lane] | grp[s * GRP_STRIDE + 32 + warp];
        run_reg_subpass<R>(tile, sp, tbase, cta_bits);  // 2^(T-R) groups == NT threads
      } else {
        run_smem_dense(tile, hdr, sp, gates, tid, NT);
      }
#ifdef PASS_PROFILE
      const unsigned long long _bt = clock64();
#endif
      __syncthreads();
#ifdef PASS_PROFILE
      if (threadIdx.x == 0) atomicAdd(&g_pass_prof[11], clock64() - _bt);
#endif
      sp += sp[S_WORDS];
    }
    PROF_MARK(4);
    TRACE(3);

    // ---- store ----
    {
      float2* dst_b = A.dst + ((size_t)b << A.nbits) + (base | (uint64_t)((2 * tid) & lowmask));
      uint64_t off[N_IO];
      float4 v[N_IO];
#pragma unroll
      for (int u = 0; u < N_IO; ++u) {
        off[u] = (uint64_t)hf[u * hstep] << L;
        v[u] = *reinterpret_cast<const float4*>(tile + (io_s0 ^ swz(2 * NT * u)));
      }
#pragma unroll
      for (int u = 0; u < N_IO; ++u) stg_stream(reinterpret_cast<float4*>(dst_b + off[u]), v[u]);
    }
    PROF_MARK(5);
    TRACE(4);
    __syncthreads();  // the tile and sprog may be overwritten from here on
    PROF_MARK(6);
    TRACE(5);
    TRACE_NEXT;
  }
}
