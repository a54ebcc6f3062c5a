// smg_lltc.cuh -- K1, the Hamming log-likelihood block (code/neal8.cpp:40-56), as an EXACT integer GEMM on the
// 5th-generation tensor cores (tcgen05.mma.kind::i8, accumulators in tensor memory).
//
//   LL[i][k] = -sum_j [x_ij != c_kj]/sigma_kj - sden_k
//            = -(Q_k - sum_j q_kj [x_ij == c_kj]) * 2^-s_k - sden_k ,      q_kj = round(2^s_k / sigma_kj) < 2^47,  Q_k = sum_j q_kj
//
// The match sum is a contraction over (level a, attribute j):  M[i][k] = sum_{a,j} [x_ij == a] * (q_kj [c_kj == a]).
// A = one-hot rows of X (u8 0/1, generated in shared memory from the codes, never stored), B = the six base-256 digits
// of q_kj placed at the level of the centre (u8), D = s32 per digit plane; the planes are recombined in 64-bit integers,
// so the match sum is the exact integer sum_j q_kj [x_ij == c_kj] and the only approximation is the 47-bit fixed-point
// representation of 1/sigma (relative 2^-47 of the cluster's largest weight: |dLL| < p * 2^-47 * max_j 1/sigma_kj, a few
// 1e-13 at p = 256, inside the 1e-12 gate; mismatch counts are unaffected).
//
// Tiling: an N tile is 16 clusters x 6 digit planes = 96 columns; its whole B operand (96 x KD bytes, KD = levels x pp)
// stays resident in shared memory of a persistent CTA, fetched once with bulk-async copies (TMA engine).  The CTA walks
// row tiles of 128 observations: 8 producer warps build the 128 x 32 one-hot A tiles of each k-step (one 16-byte
// compare-and-store per thread) into a 6-stage ring, one thread issues the MMAs (M=128, N=96, K=32) and commits them to
// the stage's mbarrier, 4 epilogue warps read the finished accumulator (double-buffered in TMEM) with tcgen05.ld,
// recombine the digit planes and write 16 doubles per row.  No block barrier in the main loop.
// Algorithmic bytes per launch: n*pp (X) + 8*n*K (LL); ops: 2 * n * KD * 96 * ceil(K/16) u8 MACs.
#pragma once
#include "smg_psm.cuh"

namespace smg {

#define LTC_M 128
#define LTC_NC 16
#define LTC_ND 6
#define LTC_N (LTC_NC * LTC_ND)  // 96
#define LTC_STAGES 6      // mbarrier slots (stages in use: NSTG <= 6, KG k-steps each)
#define LTC_EPI 128        // 4 epilogue warps (TMEM lane quadrants)
#define LTC_PROD 256       // 8 producer warps
#define LTC_THREADS (LTC_EPI + LTC_PROD + 32)
#define LTC_A_BYTES (LTC_M * 32)
#define LTC_QBITS 46       // q < 2^47

// bytes of the B operand of one N tile: [96/8 row groups][KDp/16 chunks][8 rows][16 bytes]
__host__ __device__ inline size_t ltc_b_bytes(int KDp) { return (size_t)LTC_N * KDp; }
__host__ __device__ inline int ltc_kdp(int pp, int mmax) { return (pp * mmax + 31) & ~31; }

// ---- per-sweep operand preparation: one CTA per cluster slot of the N tiles (clusters >= K: zero rows)
__global__ void __launch_bounds__(256) ll_tc_prep_kernel(int pp, int mmax, int KDp, const int* __restrict__ Kptr /* snapshot */,
                                                         const uint8_t* __restrict__ cen, const double* __restrict__ isg,
                                                         uint8_t* __restrict__ Bg, long long* __restrict__ Qsum,
                                                         double* __restrict__ scale, const double* __restrict__ sden,
                                                         double* __restrict__ cst) {
  const int kcl = blockIdx.x, K = *Kptr;
  const int nt = kcl / LTC_NC, kc = kcl % LTC_NC;
  if (nt * LTC_NC >= K) return;  // the whole tile is unused
  uint8_t* Bt = Bg + (size_t)nt * ltc_b_bytes(KDp);
  __shared__ double s_max[256];
  __shared__ long long s_sum[256];
  const int tid = threadIdx.x;
  const bool live = kcl < K;
  double mx = 0.0;
  if (live)
    for (int j = tid; j < pp; j += 256) mx = fmax(mx, isg[(size_t)kcl * pp + j]);
  s_max[tid] = mx;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) s_max[tid] = fmax(s_max[tid], s_max[tid + o]);
    __syncthreads();
  }
  const double wmax = s_max[0];
  const int sh = (wmax > 0.0 && isfinite(wmax)) ? LTC_QBITS - ilogb(wmax) : 0;
  long long qs = 0;
  for (int j = tid; j < pp; j += 256) {
    long long q = 0;
    int c = 0;
    if (live) {
      q = llrint(ldexp(isg[(size_t)kcl * pp + j], sh));
      c = cen[(size_t)kcl * pp + j];
    }
    qs += q;
    for (int a = 1; a <= mmax; a++) {
      const int kk = (j >> 4) * (16 * mmax) + (a - 1) * 16 + (j & 15);  // levels of a 16-attribute chunk are adjacent
#pragma unroll
      for (int d = 0; d < LTC_ND; d++) {
        const int nrow = d * LTC_NC + kc;
        const uint8_t v = (c == a) ? (uint8_t)((q >> (8 * d)) & 0xff) : (uint8_t)0;
        Bt[(size_t)(nrow >> 3) * (KDp * 8) + (size_t)(kk >> 4) * 128 + (nrow & 7) * 16 + (kk & 15)] = v;
      }
    }
  }
  // k padding beyond levels * pp
  for (int kk = mmax * pp + tid; kk < KDp; kk += 256)
#pragma unroll
    for (int d = 0; d < LTC_ND; d++) {
      const int nrow = d * LTC_NC + kc;
      Bt[(size_t)(nrow >> 3) * (KDp * 8) + (size_t)(kk >> 4) * 128 + (nrow & 7) * 16 + (kk & 15)] = 0;
    }
  s_sum[tid] = qs;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) s_sum[tid] += s_sum[tid + o];
    __syncthreads();
  }
  if (tid == 0) {
    Qsum[kcl] = s_sum[0];
    scale[kcl] = ldexp(1.0, -sh);
    // epilogue constants: Q = Qlo + Qhi 2^24 (both exact doubles), the scale and the log-normaliser sum
    cst[(size_t)kcl * 4 + 0] = (double)(s_sum[0] & 0xffffffll);
    cst[(size_t)kcl * 4 + 1] = (double)(s_sum[0] >> 24);
    cst[(size_t)kcl * 4 + 2] = ldexp(1.0, -sh);
    cst[(size_t)kcl * 4 + 3] = live ? sden[kcl] : 0.0;
  }
}

#ifdef SMG_LTC_PROFILE
#define LTC_TICK(k)                                                          \
  do {                                                                       \
    const long long _t = clock64();                                          \
    if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 4 || warp == 12) && prof) prof[k] += (unsigned long long)(_t - tk); \
    tk = _t;                                                                 \
  } while (0)
#else
#define LTC_TICK(k)
#endif

__device__ __forceinline__ void ltc_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(psm_smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(LTC_THREADS, 1)
    hamming_ll_tc_kernel(const uint8_t* __restrict__ X, int n, int pp, int mmax, int KDp, const int* __restrict__ Ksnap,
                         const uint8_t* __restrict__ Bg, const long long* __restrict__ Qsum,
                         const double* __restrict__ scale, const double* __restrict__ sden, double* __restrict__ LL, int ldl,
                         int* __restrict__ tile_ctr, int KG, int NSTG, const double* __restrict__ cst,
                         unsigned long long* __restrict__ prof) {
  extern __shared__ __align__(128) uint8_t ltc_smem[];
  __shared__ __align__(8) uint64_t s_afull[LTC_STAGES], s_afree[LTC_STAGES], s_accfull[2], s_accfree[2], s_bfull;
  __shared__ uint32_t s_tmem;
  __shared__ int s_tq[16];  // row tile of this CTA's tc-th piece of work (-1: none left), written by the producers
  __shared__ double s_cst[LTC_NC * 4];  // epilogue constants of this N tile's clusters
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = *Ksnap;  // (a copy taken before this launch: the live K may change under a concurrent proposal)
  const int NT = (K + LTC_NC - 1) / LTC_NC;
  const int CPN = (int)gridDim.x / NT;  // CTAs per N tile
  if (CPN == 0 || (int)blockIdx.x >= NT * CPN) return;
  const int nt = (int)blockIdx.x % NT;
  const int ntiles = (n + LTC_M - 1) / LTC_M;
  const int KS = KDp / 32;
  const size_t bbytes = ltc_b_bytes(KDp);
  const int pps = pp + 16;  // padded row stride of the staged codes (spreads rows over the banks)
  uint8_t* sB = ltc_smem;
  uint8_t* sX = ltc_smem + bbytes;                          // [128][pps] codes of the row tile
  uint8_t* sA = sX + (size_t)LTC_M * pps;

  if (tid == 0) {
    for (int s = 0; s < LTC_STAGES; s++) {
      psm_mbar_init(&s_afull[s], LTC_PROD / 32);  // one arrival per producer warp
      psm_mbar_init(&s_afree[s], 1);
    }
    for (int b = 0; b < 2; b++) {
      psm_mbar_init(&s_accfull[b], 1);
      psm_mbar_init(&s_accfree[b], LTC_EPI / 32);
    }
    psm_mbar_init(&s_bfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {  // 256 TMEM columns: two accumulator buffers of 96 (at columns 0 and 128)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(psm_smem_u32(&s_tmem)), "r"(256u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid < LTC_NC * 4) s_cst[tid] = cst[(size_t)nt * LTC_NC * 4 + tid];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = s_tmem;
  if (tid == 0) {
    // the resident B operand of this N tile: bulk-async copies (TMA engine) completing on an mbarrier
    const uint8_t* src = Bg + (size_t)nt * bbytes;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(psm_smem_u32(&s_bfull)), "r"((uint32_t)bbytes)
                 : "memory");
    for (size_t o = 0; o < bbytes; o += 32768) {
      const uint32_t sz = (uint32_t)((bbytes - o) < 32768 ? (bbytes - o) : 32768);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       psm_smem_u32(sB + o)),
                   "l"(src + o), "r"(sz), "r"(psm_smem_u32(&s_bfull))
                   : "memory");
    }
  }
  constexpr uint32_t IDESC = (2u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(LTC_N >> 3) << 17) | ((uint32_t)(LTC_M >> 4) << 24);
#ifdef SMG_LTC_PROFILE
  long long tk = clock64();
#endif

  // Row tiles are handed out dynamically (one counter per N tile): when the kernel shares the GPU with the split-merge
  // proposal only a few of its CTAs run at first, and they must be able to take all the work.
  if (warp >= LTC_EPI / 32 && warp < (LTC_EPI + LTC_PROD) / 32) {
    // ===== producers: thread = (row, 16-byte half of the k-step); A tile layout [row/8][2][row%8][16]
    const int pt = tid - LTC_EPI, r = pt >> 1, c = pt & 1;
    const uint32_t aoff = (uint32_t)((r >> 3) * 256 + c * 128 + (r & 7) * 16);
    int it = 0;
    for (int tc = 0;; tc++) {
      if (pt == 0) {
        const int t = atomicAdd(&tile_ctr[nt], 1);
        s_tq[tc & 15] = t < ntiles ? t : -1;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(LTC_PROD) : "memory");
      LTC_TICK(0);
      const int rt = s_tq[tc & 15];
      if (rt < 0) {
        // nothing left: one empty round of arrivals wakes the MMA thread, which finds the sentinel and stops
        const int s = it % NSTG;
        if (it >= NSTG) psm_mbar_wait(&s_afree[s], (unsigned)(((it / NSTG) - 1) & 1));
        __syncwarp();
        if (lane == 0) ltc_mbar_arrive(&s_afull[s]);
        break;
      }
      // ---- the codes of the row tile, global -> shared with coalesced 16-byte loads all in flight together (one L2
      //      latency per tile; the MMAs still queued in the ring run meanwhile)
      {
        const int cpr = pp >> 4, total = LTC_M * cpr;
        for (int base = 0; base < total; base += LTC_PROD * 8) {
          uint4 v[8];
#pragma unroll
          for (int u = 0; u < 8; u++) {
            const int id = base + u * LTC_PROD + pt;
            const int rr = id / cpr, ch = id - rr * cpr, grow = rt * LTC_M + rr;
            v[u] = (id < total && grow < n) ? *reinterpret_cast<const uint4*>(X + (size_t)grow * pp + ch * 16) : make_uint4(0, 0, 0, 0);
          }
#pragma unroll
          for (int u = 0; u < 8; u++) {
            const int id = base + u * LTC_PROD + pt;
            const int rr = id / cpr, ch = id - rr * cpr;
            if (id < total) *reinterpret_cast<uint4*>(sX + (size_t)rr * pps + ch * 16) = v[u];
          }
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(LTC_PROD) : "memory");
      LTC_TICK(1);
      // This thread's 16-byte chunks of the tile are q = c, c + 2, ...; chunk q is level (q % mmax) + 1 of the attribute
      // chunk jc = q / mmax, so one 16-byte read of codes serves mmax consecutive chunks.
      const int njc = pp >> 4;
      const uint8_t* xs = sX + (size_t)r * pps;
      auto ldx = [&](int jq) -> uint4 { return jq < njc ? *reinterpret_cast<const uint4*>(xs + jq * 16) : make_uint4(0, 0, 0, 0); };
      int jc = 0, a = c + 1;
      while (a > mmax) {
        a -= mmax;
        jc++;
      }
      uint4 x0 = ldx(jc);
      // a stage holds KG k-steps (KG MMAs per hand-over: the per-stage synchronisation cost is amortised over them)
      for (int ks0 = 0; ks0 < KS; ks0 += KG, it++) {
        const int s = it % NSTG;
        if (it >= NSTG) psm_mbar_wait(&s_afree[s], (unsigned)(((it / NSTG) - 1) & 1));
        LTC_TICK(2);
        uint8_t* st = sA + (size_t)s * KG * LTC_A_BYTES + aoff;
        const int gmax = min(KG, KS - ks0);
        for (int g = 0; g < gmax; g++) {
          const bool live = jc < njc;
          const uint32_t av = (uint32_t)a * 0x01010101u;
          uint4 o;
          o.x = live ? (__vcmpeq4(x0.x, av) & 0x01010101u) : 0u;
          o.y = live ? (__vcmpeq4(x0.y, av) & 0x01010101u) : 0u;
          o.z = live ? (__vcmpeq4(x0.z, av) & 0x01010101u) : 0u;
          o.w = live ? (__vcmpeq4(x0.w, av) & 0x01010101u) : 0u;
          *reinterpret_cast<uint4*>(st + (size_t)g * LTC_A_BYTES) = o;
          a += 2;  // next chunk of this thread
          while (a > mmax) {
            a -= mmax;
            jc++;
            x0 = ldx(jc);
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core
        __syncwarp();
        if (lane == 0) ltc_mbar_arrive(&s_afull[s]);
        LTC_TICK(3);
      }
    }
  } else if (tid == LTC_EPI + LTC_PROD) {
    // ===== MMA issuer
    psm_mbar_wait(&s_bfull, 0);
    const uint32_t b0 = psm_smem_u32(sB), a0 = psm_smem_u32(sA);
    int it = 0;
    for (int tc = 0;; tc++) {
      const int buf = tc & 1;
      // the first stage of the tile: once it is full the producers have published the tile index
      psm_mbar_wait(&s_afull[it % NSTG], (unsigned)((it / NSTG) & 1));
      LTC_TICK(4);
      const int rt = *reinterpret_cast<volatile int*>(&s_tq[tc & 15]);
      if (tc >= 2) psm_mbar_wait(&s_accfree[buf], (unsigned)(((tc >> 1) - 1) & 1));
      LTC_TICK(5);
      if (rt < 0) {
        psm_commit(&s_accfull[buf]);  // (nothing pending: arrives at once) wakes the epilogue, which finds the sentinel
        break;
      }
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int ks0 = 0; ks0 < KS; ks0 += KG, it++) {
        const int s = it % NSTG;
        if (ks0 > 0) psm_mbar_wait(&s_afull[s], (unsigned)((it / NSTG) & 1));
        LTC_TICK(4);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int gmax = min(KG, KS - ks0);
        for (int g = 0; g < gmax; g++) {
          const uint64_t da = psm_desc(a0 + (uint32_t)((s * KG + g) * LTC_A_BYTES), 128, 256);
          const uint64_t db = psm_desc(b0 + (uint32_t)((ks0 + g) * 256), 128, (uint32_t)KDp * 8);
          psm_mma_i8(tmem_d + (uint32_t)(buf * 128), da, db, IDESC, (ks0 + g) > 0 ? 1u : 0u);
        }
        psm_commit(&s_afree[s]);
        LTC_TICK(6);
      }
      psm_commit(&s_accfull[buf]);
    }
  } else if (warp < LTC_EPI / 32) {
    // ===== epilogue: warp q owns TMEM lanes [32q, 32q+32) = rows 32q + lane of the tile
    const int q = warp;
    for (int tc = 0;; tc++) {
      const int buf = tc & 1;
      psm_mbar_wait(&s_accfull[buf], (unsigned)((tc >> 1) & 1));
      LTC_TICK(7);
      const int rt = *reinterpret_cast<volatile int*>(&s_tq[tc & 15]);
      if (rt < 0) break;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      LTC_TICK(8);
      // all six digit planes in flight, one wait
      uint32_t v[LTC_ND][16];
#pragma unroll
      for (int d = 0; d < LTC_ND; d++) {
        const uint32_t taddr = tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)(buf * 128 + d * LTC_NC);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(v[d][0]), "=r"(v[d][1]), "=r"(v[d][2]), "=r"(v[d][3]), "=r"(v[d][4]), "=r"(v[d][5]), "=r"(v[d][6]),
              "=r"(v[d][7]), "=r"(v[d][8]), "=r"(v[d][9]), "=r"(v[d][10]), "=r"(v[d][11]), "=r"(v[d][12]), "=r"(v[d][13]),
              "=r"(v[d][14]), "=r"(v[d][15])
            : "r"(taddr));
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      if (lane == 0) ltc_mbar_arrive(&s_accfree[buf]);  // the accumulator buffer may be overwritten
      LTC_TICK(9);
      const int row = rt * LTC_M + 32 * q + lane;
      if (row < n) {
        double* o = LL + (size_t)row * ldl + nt * LTC_NC;
        double out[LTC_NC];
#pragma unroll
        for (int kc = 0; kc < LTC_NC; kc++) {
          // match sum M = lo + hi 2^24 with lo = planes 0..2 and hi = planes 3..5 packed in 32 bits each (a plane is at
          // most 255 * 256, so neither overflows); Q - M = (Qlo - lo) + (Qhi - hi) 2^24 with both differences exact in
          // double precision: ONE rounding, the correctly rounded value of the exact integer difference
          const uint32_t lo = v[0][kc] + (v[1][kc] << 8) + (v[2][kc] << 16), hi = v[3][kc] + (v[4][kc] << 8) + (v[5][kc] << 16);
          const double dlo = s_cst[kc * 4 + 0] - (double)lo, dhi = s_cst[kc * 4 + 1] - (double)hi;
          const double D = fma(dhi, 16777216.0, dlo);
          out[kc] = fma(-D, s_cst[kc * 4 + 2], -s_cst[kc * 4 + 3]);
        }
        if (nt * LTC_NC + LTC_NC <= K && (ldl & 3) == 0) {
#pragma unroll
          for (int e = 0; e < LTC_NC; e += 4) *reinterpret_cast<double4*>(o + e) = make_double4(out[e], out[e + 1], out[e + 2], out[e + 3]);
        } else {
#pragma unroll
          for (int kc = 0; kc < LTC_NC; kc++)
            if (nt * LTC_NC + kc < K) o[kc] = out[kc];
        }
      }
      LTC_TICK(10);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(256u) : "memory");
}

}  // namespace smg
