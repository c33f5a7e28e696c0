// Prime-factor search kernel (see sgx_pfa.cuh) -- its own translation unit so that it builds in seconds.
#include "sgx_pfa.cuh"

namespace sgx {
namespace pfa {

// The scratch of a transform is dead once pass B has read it: dropping its lines from the L2 (instead of letting them
// age out and be written back) keeps the live intermediate of the 444 resident CTAs inside the L2.
__device__ __forceinline__ void l2_discard_line(const void* p) {
#ifndef SGX_EMUL
  asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory");
#else
  (void)p;
#endif
}

// Pass B of one transform for one warp: DFT over (k4, k3) of its slices of 8 tauA columns, read from the scratch.
// MODE 0: maximum of |.|^2 only (hot path: 3 instructions per point; the code phase of the winning row is found by the
//         masked kernel, which recomputes that row anyway);
// MODE 1: exact arg-max, smallest index among equal values (numpy's rule);
// MODE 2: maximum over the second-peak candidates of acquisition.py:147-159 around code phase cp.
// DIRECT: the scratch columns are in the order the radix-P2 stage of pass A produces them, block tau2 * 4 + tau1 / 8,
//         position tau1 % 8 (one padding column per 32), instead of tauA = tau1 * P2 + tau2.
template <class S, int WARPS, int MODE, bool DISCARD, bool DIRECT>
__device__ __forceinline__ void pass_b(const cpx* scr, cpx* X, int w, int lane, int cp, int chip, float& best, int& bidx) {
  constexpr int P2 = S::p2, P3 = S::p3, P4 = S::p4;
  constexpr int NA = S::NA, N = S::N, CB = S::CB, KB = 32 / CB;
  constexpr int Q1 = N / S::p1, Q2 = N / P2, Q3 = N / P3, Q4 = N / P4;
  const int jb = lane & (CB - 1), kb = lane / CB;
  // Two register sets in flight: within a slice the loads of round r+2 travel while round r is computed, and the first
  // TWO rounds of the next slice travel during this slice's second stage (the scratch of 444 CTAs does not fit the L2:
  // many of these loads come from HBM).
  static_assert(P3 / KB == 4, "four rounds per slice");
  cpx ua[P4], ub[P4];
  // round r of slice sb: k3 = kb + KB*r for the lanes' kb, all k4: the P4 pass-A slices r*P4 + k4, one 256-byte block each
  auto load_round = [&](cpx* u, int sb, int r) {
    const cpx* p = scr + ((sb * S::NSA + r * P4) * S::CA * CB) + lane;
#pragma unroll
    for (int k4 = 0; k4 < P4; ++k4) u[k4] = __ldcg(p + k4 * (S::CA * CB));
  };
  auto do_round = [&](cpx* u, int r) {
    fft::Dft<P4, true>::run(u);
    cpx* rp = X + ((kb + KB * r) * P4) * CB + jb;
#pragma unroll
    for (int t4 = 0; t4 < P4; ++t4) rp[t4 * CB] = u[t4];
  };
  if (w < S::NSB) { load_round(ua, w, 0); load_round(ub, w, 1); }
#pragma unroll 1
  for (int sb = w; sb < S::NSB; sb += WARPS) {
    const int tA = sb * CB + jb;
    static_assert(!DIRECT || (S::p1 == 31 && CB == 8 && S::NSB == 4 * P2), "block order of the direct store");
    const int t1d = (sb & 3) * CB + jb, t2d = sb >> 2;
    const bool valid = DIRECT ? t1d < S::p1 : tA < NA;
    // stage 1 (radix P4 over k4): lane (k3 = kb + KB*r, column jb); tile rows k3*P4 + k4 -> k3*P4 + tau4
    do_round(ua, 0);
    load_round(ua, sb, 2);
    do_round(ub, 1);
    load_round(ub, sb, 3);
    do_round(ua, 2);
    do_round(ub, 3);
    __syncwarp();
    if (DISCARD) {   // all four rounds of this slice are in the tile: 4 x P4 blocks of 256 bytes
      constexpr int LINES = P4 * 2;
      for (int l = lane; l < 4 * LINES; l += 32)
        l2_discard_line(scr + ((sb * S::NSA + (l / LINES) * P4) * S::CA * CB) + (l % LINES) * 16);
    }
    if (sb + WARPS < S::NSB) { load_round(ua, sb + WARPS, 0); load_round(ub, sb + WARPS, 1); }
    // stage 2 (radix P3 over k3): lane (tau4 = kb + KB*r, column jb); outputs stay in registers
    const int baseA = !valid ? 0 : DIRECT ? (t1d * Q1 + t2d * Q2) % N : ((tA / P2) * Q1 + (tA % P2) * Q2) % N;
#pragma unroll 1
    for (int t4 = kb; t4 < P4; t4 += KB) {
      if (valid) {
        cpx u[P3];
        const cpx* rp = X + t4 * CB + jb;
#pragma unroll
        for (int k3 = 0; k3 < P3; ++k3) u[k3] = rp[(k3 * P4) * CB];
        fft::Dft<P3, true>::run(u);
        if (MODE == 0) {
#pragma unroll
          for (int t3 = 0; t3 < P3; ++t3) best = fmaxf(best, fmaf(u[t3].x, u[t3].x, u[t3].y * u[t3].y));
        } else {
          int tau = (baseA + t4 * Q4) % N;
#pragma unroll
          for (int t3 = 0; t3 < P3; ++t3) {
            const float mag = fmaf(u[t3].x, u[t3].x, u[t3].y * u[t3].y);
            const bool cand = MODE == 1 || second_peak_candidate(tau, cp, chip, N);
            if (cand && (mag > best || (mag == best && tau < bidx))) { best = mag; bidx = tau; }
            tau += Q3;
            if (tau >= N) tau -= N;
          }
        }
      }
    }
    __syncwarp();
  }
}

template <int WARPS>
__device__ __forceinline__ unsigned long long block_max(unsigned long long key, unsigned long long* red, int w, int lane) {
#pragma unroll
  for (int msk = 16; msk > 0; msk >>= 1) {
    const unsigned long long o = __shfl_xor_sync(0xffffffffu, key, msk);
    key = o > key ? o : key;
  }
  if (lane == 0) red[w] = key;
  __syncthreads();
  unsigned long long k = red[0];
#pragma unroll
  for (int i = 1; i < WARPS; ++i) k = red[i] > k ? red[i] : k;
  return k;
}

// Hot kernel (MASKED = false): one key (maximum of |.|^2, index field unused) per transform.
// MASKED: one transform per (rec, prn), the winning (bin, block) of select_kernel: the code phase (exact arg-max) is
// written to sel[].codePhase, the key of the second peak (acquisition.py:147-162) to partial[].
//
// Warp-independent slices: in pass A a warp owns 4 columns (kB values) x all NA rows, in pass B 8 columns (tauA
// values) x all NB rows; every butterfly stage of a slice reads only what the same warp wrote, so the only
// block-wide barriers are the two per transform (pass A -> pass B; key reduction / scratch reuse).
// Per warp two shared-memory buffers of NA x 4 values: X (work tile) and Y (code-spectrum slice, filled by a bulk copy
// one slice ahead); the spectrum slice of the next slice travels in registers while the current one goes through
// its second stage and the scratch store.  Pass B uses X and Y together as one NB x 8 tile.
// BULK: the code slice is fetched by one TMA 1-D bulk copy per slice (cp.async.bulk, completion on a per-warp mbarrier)
// instead of 14 cp.async per lane.
template <int P1, int P2, int P3, int P4, int WARPS, int MINB, bool MASKED, bool UNROLL31 = false, bool BULK = false,
          bool DISCARD = false, bool DIRECT = false>
__global__ void __launch_bounds__(WARPS * 32, MINB) pfa_search_kernel(SearchArgs a) {
  typedef Shape<P1, P2, P3, P4> S;
  static_assert(P1 == 31, "stage 1 is the grouped radix-31 butterfly");
  constexpr int NA = S::NA, NB = S::NB, N = S::N, SROW = S::SROW;
  constexpr int CA = S::CA, CB = S::CB;            // columns per warp slice
  constexpr int KA = 32 / CA, KB = 32 / CB;        // butterflies per warp round
  static_assert(CA == 4 && P3 / KB == 4, "copy loops / register ping-pong");
  SGX_DYN_SMEM(smem);
  __shared__ unsigned long long red[2][WARPS];
  __shared__ unsigned long long cbar[WARPS];
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  unsigned cphase = 0;
  if (BULK) {
    if (tid < WARPS) mbar_init(&cbar[tid], 1);
    __syncthreads();
  }
  cpx* X = reinterpret_cast<cpx*>(smem) + (size_t)w * S::WARP_TILE;
  cpx* Y = X + NA * CA;
  cpx* scr = a.scratch + (size_t)blockIdx.x * NB * SROW;
  const int ja = lane & (CA - 1), ka = lane / CA;
  const bool act = ka < P2;
  const int off0 = (act ? ka : 0) * CA + ja;

  const cpx* src = nullptr;
  const cpx* cod = nullptr;
  long long out_index = 0;
  auto locate = [&](long long item) {
    if (!MASKED) {   // PRN fastest: the CTAs running at any moment share a handful of spectra and the 32 code spectra
      long long t = item;
      const int prn = (int)(t % a.d.nprn); t /= a.d.nprn;
      const int blk = (int)(t % a.d.blocks); t /= a.d.blocks;
      const int bin = (int)(t % a.d.nbins);
      const int rec = (int)(t / a.d.nbins);
      src = a.spec + (((long long)rec * a.d.blocks + blk) * a.d.nbins + bin) * N;
      cod = a.codeF + (long long)(a.d.prn_first + prn) * N;
      out_index = (((long long)rec * a.d.nprn + prn) * a.d.nbins + bin) * a.d.blocks + blk;
    } else {
      const PeakSel s = a.sel[item];
      const int prn = (int)(item % a.d.nprn), rec = (int)(item / a.d.nprn);
      src = a.spec + (((long long)rec * a.d.blocks + s.blk) * a.d.nbins + s.bin) * N;
      cod = a.codeF + (long long)(a.d.prn_first + prn) * N;
      out_index = item;
    }
  };
  cpx v[P1];
  // code slice `sa` -> Y (the slice is one contiguous block in the slice-major storage order: 16 bytes per lane, fully
  // coalesced), spectrum slice -> registers (a row group of P2 x CA values is 224 contiguous bytes)
  auto fetch = [&](int sa) {
    if (BULK) {
      if (lane == 0) bulk_load(Y, cod + sa * (NA * CA), (unsigned)(sizeof(cpx) * NA * CA), &cbar[w]);
    } else {
      const cpx* q = cod + sa * (NA * CA) + lane * 2;
      cpx* yd = Y + lane * 2;
#pragma unroll
      for (int it = 0; it < (NA * CA + 63) / 64; ++it)
        if (it < (NA * CA) / 64 || lane * 2 + 64 * it < NA * CA) cp_async16(yd + it * 64, q + it * 64);
      cp_async_commit();
    }
    if (act) {
      const cpx* p = src + sa * (NA * CA) + off0;
#pragma unroll
      for (int k1 = 0; k1 < P1; ++k1) v[k1] = __ldg(p + k1 * (P2 * CA));
    }
  };

#pragma unroll 1
  for (long long item = blockIdx.x; item < a.nitems; item += gridDim.x) {
    locate(item);
    if (w < S::NSA) fetch(w);
    // ---------------- pass A: DFT over (k1, k2); slice = CA values of kB = k3*P4 + k4 --------------------------------
#pragma unroll 1
    for (int sa = w; sa < S::NSA; sa += WARPS) {
      if (BULK) {
        mbar_wait(&cbar[w], cphase);
        cphase ^= 1;
      } else {
        cp_async_wait_all();
      }
      __syncwarp();
      // stage 1 (radix 31): lane (k2 = ka, column ja); rows k1*P2 + k2 -> tau1*P2 + k2
      if (act) {
        R31 bf;
        cpx* x = X + ka * CA + ja;
        x[0] = bf.prepare(v, Y + ka * CA + ja, P2 * CA);
        if (UNROLL31) {
          cpx hi, lo;
#define SGX_R31_AT(m)                                          \
  {                                                            \
    bf.pair_at<m>(hi, lo);                                     \
    x[KHI31C[m] * (P2 * CA)] = hi;                             \
    x[(P1 - KHI31C[m]) * (P2 * CA)] = lo;                      \
  }
          constexpr int KHI31C[15] = SGX_R31_KHI;
          SGX_R31_AT(0) SGX_R31_AT(1) SGX_R31_AT(2) SGX_R31_AT(3) SGX_R31_AT(4) SGX_R31_AT(5) SGX_R31_AT(6) SGX_R31_AT(7)
          SGX_R31_AT(8) SGX_R31_AT(9) SGX_R31_AT(10) SGX_R31_AT(11) SGX_R31_AT(12) SGX_R31_AT(13) SGX_R31_AT(14)
#undef SGX_R31_AT
        } else {
#pragma unroll 1
        for (int q = 0; q < 3; ++q) {
          cpx hi, lo;
#define SGX_R31_PAIR(r)                                        \
  {                                                            \
    bf.pair<r>(hi, lo);                                        \
    const int kh = KHI31[q * 5 + r];                           \
    x[kh * (P2 * CA)] = hi;                                    \
    x[(P1 - kh) * (P2 * CA)] = lo;                             \
  }
          SGX_R31_PAIR(0) SGX_R31_PAIR(1) SGX_R31_PAIR(2) SGX_R31_PAIR(3) SGX_R31_PAIR(4)
#undef SGX_R31_PAIR
          if (q < 2) bf.rotate();
        }
      }
        }
      __syncwarp();
      // the next slice's operands travel while this slice goes through stage 2 and the scratch store
      if (sa + WARPS < S::NSA) fetch(sa + WARPS);
      // stage 2 (radix P2): lane (tau1 = ka + KA*r, column ja), in place
      cpx* dblk = scr + sa * (CA * CB) + ja * CB + ka;     // DIRECT: block (tau2 * 4 + round), slice sa, column ja, position ka
#pragma unroll 1
      for (int t1 = ka; t1 < P1; t1 += KA) {
        cpx u[P2];
        cpx* rp = X + (t1 * P2) * CA + ja;
#pragma unroll
        for (int k2 = 0; k2 < P2; ++k2) u[k2] = rp[k2 * CA];
        fft::Dft<P2, true>::run(u);
        if (DIRECT) {   // one 256-byte block per store instruction: lanes (ka, ja) -> position ja * 8 + ka
#pragma unroll
          for (int t2 = 0; t2 < P2; ++t2) __stcg(dblk + t2 * (4 * S::NSA * CA * CB), u[t2]);
          dblk += S::NSA * CA * CB;
        } else {
#pragma unroll
          for (int t2 = 0; t2 < P2; ++t2) rp[t2 * CA] = u[t2];
        }
      }
      __syncwarp();
      // tile row c = tauA holds the 4 columns of this slice; a lane pair takes a row (16 bytes each); in the scratch
      // (Shape::scratch_index) the 16 rows of a request are four 64-byte runs in two lines
      if (!DIRECT) {
        const int half = (lane & 1) * 2, c0 = lane >> 1;
        cpx* dst = scr + ((c0 >> 3) * S::NSA + sa) * (CA * CB) + half * CB + (c0 & 7);
        constexpr int STEP16 = 2 * S::NSA * CA * CB;     // 16 rows further: two kA / CB groups
#pragma unroll 1
        for (int c = c0; c < NA; c += 32) {      // two rows per iteration: both loads before the four stores
          const float4 x = *reinterpret_cast<const float4*>(X + c * CA + half);
          const bool second = c + 16 < NA;
          float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
          if (second) y = *reinterpret_cast<const float4*>(X + (c + 16) * CA + half);
          __stcg(dst, make_float2(x.x, x.y));
          __stcg(dst + CB, make_float2(x.z, x.w));
          if (second) {
            __stcg(dst + STEP16, make_float2(y.x, y.y));
            __stcg(dst + STEP16 + CB, make_float2(y.z, y.w));
          }
          dst += 2 * STEP16;
        }
      }
      __syncwarp();
    }
    __syncthreads();   // the whole intermediate is in the scratch (and visible to the block)
    // ---------------- pass B ------------------------------------------------------------------------------------------
    if (!MASKED) {
      float best = 0.f;
      int unused = 0;
      pass_b<S, WARPS, 0, DISCARD, DIRECT>(scr, X, w, lane, 0, 0, best, unused);
      // (the barrier inside: every warp is done reading the scratch before the next transform overwrites it)
      const unsigned long long k = block_max<WARPS>(fft::peak_key(best, 0u), red[0], w, lane);
      if (tid == 0) a.partial[out_index] = k;
    } else {
      float best = -1.f;
      int bidx = 0x7fffffff;
      pass_b<S, WARPS, 1, false, DIRECT>(scr, X, w, lane, 0, 0, best, bidx);
      const unsigned long long k1 = block_max<WARPS>(best >= 0.f ? fft::peak_key(best, (unsigned)bidx) : 0ull, red[0], w, lane);
      const int cp = (int)fft::key_index(k1);
      best = -1.f;
      bidx = 0x7fffffff;
      pass_b<S, WARPS, 2, DISCARD, DIRECT>(scr, X, w, lane, cp, a.chip, best, bidx);
      const unsigned long long k2 = block_max<WARPS>(best >= 0.f ? fft::peak_key(best, (unsigned)bidx) : 0ull, red[1], w, lane);
      if (tid == 0) {
        a.partial[out_index] = k2;
        a.sel_out[out_index].codePhase = cp;
        a.sel_out[out_index].peak = fft::key_value(k1);
      }
    }
  }
}

template <int WARPS, int MINB, bool MASKED, bool UNROLL31 = false, bool BULK = false, bool DISCARD = false, bool DIRECT = false,
          int P2 = 7>
static int launch_cfg(SearchArgs args, DevBuf& scratch, cudaStream_t s) {
  typedef Shape<31, P2, 16, 11> S;
  int dev = 0, n_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  if (n_sm <= 0) n_sm = 148;
  auto kfn = pfa_search_kernel<31, P2, 16, 11, WARPS, MINB, MASKED, UNROLL31, BULK, DISCARD, DIRECT>;
  const size_t smem = S::smem_per_warp * WARPS;
  SGX_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, WARPS * 32, smem) != cudaSuccess || occ < 1) occ = 1;
  if (occ > MINB) occ = MINB;
  long long grid = (long long)n_sm * occ;
  if (grid > args.nitems) grid = args.nitems;
  const size_t bytes = S::scratch_per_cta * (size_t)grid;
  if (scratch.reserve(bytes)) return fail(SGX_ERR_CUDA, "cudaMalloc", "search scratch");
  args.scratch = scratch.as<cpx>();
#ifdef SGX_EMUL
  SGX_COUNTED_LAUNCH(kfn, dim3((unsigned)grid), dim3(WARPS * 32), smem, s, args);
#else
  // The scratch is rewritten and re-read by its CTA for every transform: keep it in L2 (persisting access window), so
  // that the streamed spectra do not push it out to HBM.
  static int persist_max = -1, window_max = 0;
  const char* pe = getenv("SGX_PFA_PERSIST");
  const bool want_persist = pe && pe[0] == '1';
  if (persist_max < 0) {
    persist_max = 0;
    if (want_persist) {   // the set-aside shrinks the L2 available to everything else: only when asked for
      cudaDeviceGetAttribute(&persist_max, cudaDevAttrMaxPersistingL2CacheSize, dev);
      cudaDeviceGetAttribute(&window_max, cudaDevAttrMaxAccessPolicyWindowSize, dev);
      if (const char* f = getenv("SGX_PFA_PERSIST_MB")) { if (atoi(f) > 0 && ((long long)atoi(f) << 20) < persist_max) persist_max = atoi(f) << 20; }
      if (persist_max > 0) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)persist_max);
      if (getenv("SGX_DEBUG")) fprintf(stderr, "[sgx debug] persisting L2 max %d B, window max %d B\n", persist_max, window_max);
      cudaGetLastError();
    }
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(WARPS * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  int nattr = 0;
  if (want_persist && persist_max > 0 && window_max > 0) {
    const size_t win = bytes < (size_t)window_max ? bytes : (size_t)window_max;
    attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
    attr[0].val.accessPolicyWindow.base_ptr = scratch.p;
    attr[0].val.accessPolicyWindow.num_bytes = win;
    attr[0].val.accessPolicyWindow.hitRatio = win <= (size_t)persist_max ? 1.0f : (float)((double)persist_max / (double)win);
    attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    nattr = 1;
  }
  cfg.attrs = attr;
  cfg.numAttrs = nattr;
  SGX_CUDA(cudaLaunchKernelEx(&cfg, kfn, args));
  ++g_launches;
#endif
  SGX_CUDA(cudaGetLastError());
  return SGX_OK;
}

template <bool MASKED>
static int launch_search_t(SearchArgs args, DevBuf& scratch, cudaStream_t s) {
  // [v]WC: variant, warps per CTA x CTAs per SM.  Measured on B200 (59 392 transforms; the numbers of one column are
  // from the same build): 4x3 divides the 44 + 28 slices of a transform evenly and wins every time --
  //   first slice-major build: 4x3 11.30 ms, 6x2 11.59, 12x1 12.25, 8x2 16.19 (128 registers: the prefetch spills);
  //   earlier: 7x2, 14x1, 16x1 slower still, 5x3 / 4x4 (136 / 128 registers) 25.7 ms.
  if (args.p2 == 3) return launch_cfg<4, 3, MASKED, true, true, false, true, 3>(args, scratch, s);   // N = 16 368: the default build only
  int cfg = 743;
  if (const char* e = getenv("SGX_PFA_CFG")) cfg = atoi(e);
  switch (cfg) {
    case 62: return launch_cfg<6, 2, MASKED>(args, scratch, s);
    case 121: return launch_cfg<12, 1, MASKED>(args, scratch, s);
    case 43: return launch_cfg<4, 3, MASKED>(args, scratch, s);           // rolled radix-31 groups, code slices by cp.async
    // + scratch lines discarded from L2 after pass B: DRAM traffic of the launch 35.2 -> 19.6 GB, L2 hit rate 45 -> 64 %,
    // long-scoreboard stalls 0.42 -> 0.26 per issue -- and 2 % slower (9.96 vs 9.73 ms): the kernel is not DRAM-bound
    case 643: return launch_cfg<4, 3, MASKED, true, true, true>(args, scratch, s);
    case 143: return launch_cfg<4, 3, MASKED, true>(args, scratch, s);    // radix-31 butterfly unrolled, code slices by cp.async
    case 543: return launch_cfg<4, 3, MASKED, true, true>(args, scratch, s);   // + code slices by TMA bulk copy (-1 %): 9.71 ms
    // + radix-7 outputs stored straight into the scratch (blocks in (tau2, tau1 / 8) order: every store instruction is one
    // 256-byte block), no tile write-back, no copy loop: 9.01 ms, 99.8 instructions per point, issue 69 %
    default: return launch_cfg<4, 3, MASKED, true, true, false, true>(args, scratch, s);
  }
}

int launch_search(SearchArgs args, bool masked, DevBuf& scratch, cudaStream_t s) {
  if (args.nitems <= 0) return SGX_OK;
  return masked ? launch_search_t<true>(args, scratch, s) : launch_search_t<false>(args, scratch, s);
}

}  // namespace pfa
}  // namespace sgx
