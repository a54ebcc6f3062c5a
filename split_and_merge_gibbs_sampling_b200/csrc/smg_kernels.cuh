// smg_kernels.cuh -- sm_100a kernels of the Neal-8 sweep and the centre/sigma update.
//
//   K1  hamming_ll_block_kernel   neal8.cpp:40-56   (existing-cluster columns, all observations)
//       aux_ll_kernel             neal8.cpp:65-69,79-92 (auxiliary components from the pool)
//   K2  neal8_scan_kernel         launcher.cpp:95-99 -> neal8.cpp:10-160 (one resident CTA)
//       scan_finalize_*           label compaction after the pass
//   K3  cluster_histogram_kernel  common_functions.cpp:480-493,572-579
//   K4  phi_draw_kernel           common_functions.cpp:495-505,560,582-589 -> hyperg.cpp:346-378
//       loglik_kernel             common_functions.cpp:379-401
//
// Data layout in HBM (all row-major, p padded to pp = multiple of 16):
//   X      uint8  [n][pp]     category codes 1..m_j (0 in the padding)
//   cen    uint8  [NS][pp]    cluster centres by slot (0 in the padding)
//   sig    fp64   [NS][pp]    sigma
//   isg    fp64   [NS][pp]    1/sigma (0 in the padding)
//   sden   fp64   [NS]        sum_j log(1+(m_j-1)/exp(1/sigma_j))
//   LL     fp64   [n][ldl]    LL[i][slot] = -sum_j [x_ij != c_j]*isg_j - sden
#pragma once
#include <cooperative_groups.h>

#include "smg_device.cuh"

namespace smg {

#define SMG_MAX_SLOTS 1024    // slot (column) capacity of the scan's shared-memory tables
#define SMG_MAX_ENTRIES 256   // K + m_aux handled per allocation draw
#define SMG_EPL (SMG_MAX_ENTRIES / 32)
#define SMG_SCAN_WARPS 32

enum StatusBits : int {
  ST_OK = 0,
  ST_SLOTS_EXHAUSTED = 1,    // more cluster births in one pass than free slots
  ST_TOO_MANY_ENTRIES = 2,   // K + m_aux > SMG_MAX_ENTRIES
  ST_WALKER = 4,             // Rcpp would have switched to Walker alias sampling (nc > 200)
  ST_LL_COLS = 8,            // K exceeds the LL matrix column capacity
  ST_BAD_PROB = 16,          // non-finite probability (Rcpp::sample would stop())
  ST_VALIDATE = 32,          // validate_state failure (common_functions.cpp:146-172)
  ST_GRID_TIMEOUT = 64       // a grid barrier of the persistent split-merge kernel timed out
};

// =============================================================================
// K1: Hamming log-likelihood block.  Tile = 128 rows x 32 slots per CTA of 256 threads;
// a warp owns 4 slots, a lane owns 4 rows (lane, lane+32, lane+64, lane+96); 16 fp64
// accumulators per thread.  X and the per-slot tables are staged through shared memory in
// j-tiles of 64 attributes: x words are read conflict-free (row stride 17 words), the centre
// and 1/sigma words are warp-wide broadcasts.
// Algorithmic bytes per row: pp (X) + 8*K (LL write); op count: K*p compare-adds.
// =============================================================================
#define LLB_ROWS 128
#define LLB_SLOTS 32
#define LLB_JT 64
#define LLB_XS (LLB_JT / 4 + 1)  // padded row stride in words

__global__ void __launch_bounds__(256) hamming_ll_block_kernel(const uint8_t* __restrict__ X, int n, int pp,
                                                               const uint8_t* __restrict__ cen,
                                                               const double* __restrict__ isg,
                                                               const double* __restrict__ sden,
                                                               const int* __restrict__ Kptr, double* __restrict__ LL,
                                                               int ldl) {
  const int K = *Kptr;
  const int slot0 = blockIdx.y * LLB_SLOTS;
  if (slot0 >= K) return;
  const int row0 = blockIdx.x * LLB_ROWS;
  __shared__ uint32_t sx[LLB_ROWS * LLB_XS];
  __shared__ uint32_t sc[LLB_SLOTS * (LLB_JT / 4)];
  __shared__ __align__(16) double sw[LLB_SLOTS * LLB_JT];
  const int tid = threadIdx.x, lane = tid & 31, wg = tid >> 5;  // wg: slot group (4 slots)
  double acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int c = 0; c < 4; c++) acc[r][c] = 0.0;

  for (int j0 = 0; j0 < pp; j0 += LLB_JT) {
    const int jt = min(LLB_JT, pp - j0);  // multiple of 16
    // stage X tile: 128 rows x jt bytes, 16-byte chunks
    for (int ch = tid; ch < LLB_ROWS * (LLB_JT / 16); ch += 256) {
      int r = ch / (LLB_JT / 16), q = ch % (LLB_JT / 16);
      uint4 v = make_uint4(0, 0, 0, 0);
      if (row0 + r < n && q * 16 < jt) v = *reinterpret_cast<const uint4*>(X + (size_t)(row0 + r) * pp + j0 + q * 16);
      uint32_t* d = &sx[r * LLB_XS + q * 4];
      d[0] = v.x;
      d[1] = v.y;
      d[2] = v.z;
      d[3] = v.w;
    }
    // stage centre tile
    for (int ch = tid; ch < LLB_SLOTS * (LLB_JT / 16); ch += 256) {
      int s = ch / (LLB_JT / 16), q = ch % (LLB_JT / 16);
      uint4 v = make_uint4(0, 0, 0, 0);
      if (slot0 + s < K && q * 16 < jt) v = *reinterpret_cast<const uint4*>(cen + (size_t)(slot0 + s) * pp + j0 + q * 16);
      *reinterpret_cast<uint4*>(&sc[s * (LLB_JT / 4) + q * 4]) = v;
    }
    // stage 1/sigma tile
    for (int ch = tid; ch < LLB_SLOTS * (LLB_JT / 2); ch += 256) {
      int s = ch / (LLB_JT / 2), q = ch % (LLB_JT / 2);
      double2 v = make_double2(0.0, 0.0);
      if (slot0 + s < K && q * 2 < jt) v = *reinterpret_cast<const double2*>(isg + (size_t)(slot0 + s) * pp + j0 + q * 2);
      *reinterpret_cast<double2*>(&sw[s * LLB_JT + q * 2]) = v;
    }
    __syncthreads();
#pragma unroll 2
    for (int jw = 0; jw < LLB_JT / 4; jw++) {
      uint32_t xw[4], cw[4];
#pragma unroll
      for (int r = 0; r < 4; r++) xw[r] = sx[(lane + 32 * r) * LLB_XS + jw];
#pragma unroll
      for (int c = 0; c < 4; c++) cw[c] = sc[(wg * 4 + c) * (LLB_JT / 4) + jw];
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const double2 w01 = *reinterpret_cast<const double2*>(&sw[(wg * 4 + c) * LLB_JT + jw * 4]);
        const double2 w23 = *reinterpret_cast<const double2*>(&sw[(wg * 4 + c) * LLB_JT + jw * 4 + 2]);
#pragma unroll
        for (int r = 0; r < 4; r++) {
          uint32_t mm = __vcmpne4(xw[r], cw[c]);  // 0xff per mismatching byte
          if (mm & 0x000000ffu) acc[r][c] += w01.x;
          if (mm & 0x0000ff00u) acc[r][c] += w01.y;
          if (mm & 0x00ff0000u) acc[r][c] += w23.x;
          if (mm & 0xff000000u) acc[r][c] += w23.y;
        }
      }
    }
    __syncthreads();
  }
  // epilogue: each thread writes 4 consecutive doubles (one 32-byte sector) per row
  double sd[4];
#pragma unroll
  for (int c = 0; c < 4; c++) sd[c] = (slot0 + wg * 4 + c < K) ? sden[slot0 + wg * 4 + c] : 0.0;
#pragma unroll
  for (int r = 0; r < 4; r++) {
    int row = row0 + lane + 32 * r;
    if (row >= n) continue;
    double* o = LL + (size_t)row * ldl + slot0 + wg * 4;
    if (slot0 + wg * 4 + 3 < K && (ldl & 3) == 0) {
      double4 v = make_double4(-acc[r][0] - sd[0], -acc[r][1] - sd[1], -acc[r][2] - sd[2], -acc[r][3] - sd[3]);
      *reinterpret_cast<double4*>(o) = v;
    } else {
#pragma unroll
      for (int c = 0; c < 4; c++)
        if (slot0 + wg * 4 + c < K) o[c] = -acc[r][c] - sd[c];
    }
  }
}

// -----------------------------------------------------------------------------
// K1, table form (attributes with at most 7 levels, i.e. every code fits 3 bits).
// Same tile as above.  Four attributes (one 32-bit word of codes) are handled by ONE shared-memory look-up:
// per (slot, word) the CTA keeps the 16 subset sums of the word's four 1/sigma values, and the 4-bit mismatch
// pattern of (x word, centre word) indexes them.  Pattern -> address is three integer ops and a DP4A:
//     t = x ^ c                      (per byte 0..7, zero iff the codes match)
//     t = (t + 0x07070707) & 0x08080808     (per byte 8 iff mismatch)
//     addr = dp4a(t, {1,2,4,8}, base)       (= base + 8 * pattern: byte offset of the fp64 subset sum)
// so a (row, slot, word) costs 6 instructions and one DADD instead of 4 x (LOP3 + DADD + 2 FSEL): the C form
// is bound by the integer ALU pipe (86% busy, fp64 pipe 22%: profiles/r01_ncu_summary.md), this one by the
// shared-memory pipe.  16 look-up entries of one (slot, word) fill one 128-byte line: conflict-free.
// -----------------------------------------------------------------------------
// byte b of entry w is 8 when bit b of w is set: XOR-swizzle of the look-up lines (see the kernel)
__constant__ uint32_t c_llt_swz[16] = {0x00000000u, 0x00000008u, 0x00000800u, 0x00000808u, 0x00080000u, 0x00080008u,
                                       0x00080800u, 0x00080808u, 0x08000000u, 0x08000008u, 0x08000800u, 0x08000808u,
                                       0x08080000u, 0x08080008u, 0x08080800u, 0x08080808u};
#define LLT_TB_BYTES (LLB_SLOTS * (LLB_JT / 4) * 16 * 8)  // 64 KB of subset sums per attribute tile
#define LLT_SMEM_BYTES (LLT_TB_BYTES + LLB_ROWS * LLB_XS * 4 + LLB_SLOTS * (LLB_JT / 4) * 4 + LLB_SLOTS * LLB_JT * 8)

__global__ void __launch_bounds__(256, 2) hamming_ll_block_t16_kernel(const uint8_t* __restrict__ X, int n, int pp,
                                                                      const uint8_t* __restrict__ cen,
                                                                      const double* __restrict__ isg,
                                                                      const double* __restrict__ sden,
                                                                      const int* __restrict__ Kptr,
                                                                      double* __restrict__ LL, int ldl) {
  const int K = *Kptr;
  const int slot0 = blockIdx.y * LLB_SLOTS;
  if (slot0 >= K) return;
  const int row0 = blockIdx.x * LLB_ROWS;
  extern __shared__ __align__(128) unsigned char smem[];
  double* tb = reinterpret_cast<double*>(smem);                                   // [32 slots][16 words][16]
  uint32_t* sx = reinterpret_cast<uint32_t*>(smem + LLT_TB_BYTES);                // [128 rows][17]
  uint32_t* sc = sx + LLB_ROWS * LLB_XS;                                          // [32 slots][16 words]
  double* sw = reinterpret_cast<double*>(sc + LLB_SLOTS * (LLB_JT / 4));          // [32 slots][64]
  const unsigned tb_base = (unsigned)__cvta_generic_to_shared(tb);
  const int tid = threadIdx.x, lane = tid & 31, wg = tid >> 5;  // wg: slot group (4 slots)
  double acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int c = 0; c < 4; c++) acc[r][c] = 0.0;

  for (int j0 = 0; j0 < pp; j0 += LLB_JT) {
    const int jt = min(LLB_JT, pp - j0);  // multiple of 16
    for (int ch = tid; ch < LLB_ROWS * (LLB_JT / 16); ch += 256) {
      int r = ch / (LLB_JT / 16), q = ch % (LLB_JT / 16);
      uint4 v = make_uint4(0, 0, 0, 0);
      if (row0 + r < n && q * 16 < jt) v = *reinterpret_cast<const uint4*>(X + (size_t)(row0 + r) * pp + j0 + q * 16);
      uint32_t* d = &sx[r * LLB_XS + q * 4];
      d[0] = v.x;
      d[1] = v.y;
      d[2] = v.z;
      d[3] = v.w;
    }
    for (int ch = tid; ch < LLB_SLOTS * (LLB_JT / 16); ch += 256) {
      int s = ch / (LLB_JT / 16), q = ch % (LLB_JT / 16);
      uint4 v = make_uint4(0, 0, 0, 0);
      if (slot0 + s < K && q * 16 < jt) v = *reinterpret_cast<const uint4*>(cen + (size_t)(slot0 + s) * pp + j0 + q * 16);
      *reinterpret_cast<uint4*>(&sc[s * (LLB_JT / 4) + q * 4]) = v;
    }
    for (int ch = tid; ch < LLB_SLOTS * (LLB_JT / 2); ch += 256) {
      int s = ch / (LLB_JT / 2), q = ch % (LLB_JT / 2);
      double2 v = make_double2(0.0, 0.0);
      if (slot0 + s < K && q * 2 < jt) v = *reinterpret_cast<const double2*>(isg + (size_t)(slot0 + s) * pp + j0 + q * 2);
      *reinterpret_cast<double2*>(&sw[s * LLB_JT + q * 2]) = v;
    }
    __syncthreads();
    // subset sums: pair = (slot, word); bit b of the pattern <-> byte b of the word <-> attribute 4*word + b.
    // Entry k of a pair lives at position k ^ word of its 128-byte line: the 32 lanes of a warp (32 consecutive
    // pairs) then store to distinct banks, and the reader folds the XOR into its mask constant for free.
    for (int pair = tid; pair < LLB_SLOTS * (LLB_JT / 4); pair += 256) {
      const double2 w01 = *reinterpret_cast<const double2*>(&sw[pair * 4]);
      const double2 w23 = *reinterpret_cast<const double2*>(&sw[pair * 4 + 2]);
      const double s01 = w01.x + w01.y, s23 = w23.x + w23.y;
      double* line = &tb[pair * 16];
      const int sz = pair & 15;
      line[0 ^ sz] = 0.0;
      line[1 ^ sz] = w01.x;
      line[2 ^ sz] = w01.y;
      line[3 ^ sz] = s01;
      line[4 ^ sz] = w23.x;
      line[5 ^ sz] = w01.x + w23.x;
      line[6 ^ sz] = w01.y + w23.x;
      line[7 ^ sz] = s01 + w23.x;
      line[8 ^ sz] = w23.y;
      line[9 ^ sz] = w01.x + w23.y;
      line[10 ^ sz] = w01.y + w23.y;
      line[11 ^ sz] = s01 + w23.y;
      line[12 ^ sz] = s23;
      line[13 ^ sz] = w01.x + s23;
      line[14 ^ sz] = w01.y + s23;
      line[15 ^ sz] = s01 + s23;
    }
    __syncthreads();
    // (the last slot tile is rarely full: a warp whose four slots all lie beyond K is skipped -- at K = 50 that is 12
    // of the 64 columns of the two tiles)
    const int ncol = K - (slot0 + wg * 4);
    if (ncol > 0)
#pragma unroll
    for (int jw = 0; jw < LLB_JT / 4; jw++) {
      // byte b of the swizzle constant is 8 when bit b of the word index is set
      const uint32_t swz = c_llt_swz[jw];  // from constant memory: stays a register operand of the LOP3 below
      uint32_t xw[4], cw[4];
#pragma unroll
      for (int r = 0; r < 4; r++) xw[r] = sx[(lane + 32 * r) * LLB_XS + jw];
#pragma unroll
      for (int c = 0; c < 4; c++) cw[c] = sc[(wg * 4 + c) * (LLB_JT / 4) + jw];
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const unsigned base = tb_base + (unsigned)((((wg * 4 + c) * (LLB_JT / 4)) + jw) * 128);
#pragma unroll
        for (int r = 0; r < 4; r++) {
          uint32_t t = (xw[r] ^ cw[c]) + 0x07070707u;
          asm("lop3.b32 %0, %1, %2, 0x08080808, 0x28;" : "=r"(t) : "r"(t), "r"(swz));  // (t ^ swz) & 0x08080808
          const unsigned addr = __dp4a(t, 0x08040201u, base);
          double v;
          asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
          acc[r][c] += v;
        }
      }
    }
    __syncthreads();
  }
  double sd[4];
#pragma unroll
  for (int c = 0; c < 4; c++) sd[c] = (slot0 + wg * 4 + c < K) ? sden[slot0 + wg * 4 + c] : 0.0;
#pragma unroll
  for (int r = 0; r < 4; r++) {
    int row = row0 + lane + 32 * r;
    if (row >= n) continue;
    double* o = LL + (size_t)row * ldl + slot0 + wg * 4;
    if (slot0 + wg * 4 + 3 < K && (ldl & 3) == 0) {
      double4 v = make_double4(-acc[r][0] - sd[0], -acc[r][1] - sd[1], -acc[r][2] - sd[2], -acc[r][3] - sd[3]);
      *reinterpret_cast<double4*>(o) = v;
    } else {
#pragma unroll
      for (int c = 0; c < 4; c++)
        if (slot0 + wg * 4 + c < K) o[c] = -acc[r][c] - sd[c];
    }
  }
}

// One warp evaluates sum_j [x_j != c_j] * isg_j for one (row, parameter vector) pair.
// Lanes take 8 attributes per 256-attribute chunk; butterfly reduction => every lane
// returns the same value, and the value for a given (row, vector) is identical no matter
// which kernel calls this (used for aux columns, columns born during a pass, member subsets).
// (no __restrict__: the scan kernel reads parameter vectors it wrote earlier in the same launch,
// so these loads must stay on the coherent path)
__device__ __forceinline__ double warp_mismatch_dot(const uint8_t* xrow, const uint8_t* crow, const double* wrow, int pp,
                                                    int lane) {
  double acc = 0.0;
  for (int j0 = lane * 8; j0 < pp; j0 += 256) {
    uint2 xv = *reinterpret_cast<const uint2*>(xrow + j0);
    uint2 cv = *reinterpret_cast<const uint2*>(crow + j0);
    uint32_t m0 = __vcmpne4(xv.x, cv.x), m1 = __vcmpne4(xv.y, cv.y);
    if (m0 | m1) {
      const double2* w = reinterpret_cast<const double2*>(wrow + j0);
      double2 w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
      if (m0 & 0x000000ffu) acc += w0.x;
      if (m0 & 0x0000ff00u) acc += w0.y;
      if (m0 & 0x00ff0000u) acc += w1.x;
      if (m0 & 0xff000000u) acc += w1.y;
      if (m1 & 0x000000ffu) acc += w2.x;
      if (m1 & 0x0000ff00u) acc += w2.y;
      if (m1 & 0x00ff0000u) acc += w3.x;
      if (m1 & 0xff000000u) acc += w3.y;
    }
  }
  return warp_sum(acc);
}

// integer mismatch count (parity checks: bit-exact against the oracle)
__global__ void mismatch_count_kernel(const uint8_t* __restrict__ X, int n, int pp, const uint8_t* __restrict__ cen,
                                      int K, int* __restrict__ out) {
  int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n * K) return;
  int i = w / K, k = w % K;
  int acc = 0;
  for (int j0 = lane * 4; j0 < pp; j0 += 128) {
    uint32_t xv = *reinterpret_cast<const uint32_t*>(X + (size_t)i * pp + j0);
    uint32_t cv = *reinterpret_cast<const uint32_t*>(cen + (size_t)k * pp + j0);
    acc += __popc(__vcmpne4(xv, cv)) >> 3;
  }
  acc = warp_sum_i(acc);
  if (lane == 0) out[w] = acc;
}

// One prior draw (launcher.cpp:67-77: centre ~ U{1..m_j}, sigma ~ HIG(v_j, w_j, m_j)) of attribute j of entry e, a pure
// function of (key, e, j): the stored pool (pool_draw_kernel) and the pool-free auxiliary components
// (aux_ll_philox_kernel, and the scan when one of them becomes a cluster) evaluate the same code.
__device__ __forceinline__ void prior_entry_attr(const RngKey& key, long long e, int j, int m, double v, double w,
                                                 int sigma_exact, int* center_out, double* sigma_out) {
  // 64-bit entry index split over the two counter words
  uint32_t o[4];
  philox4x32_10((uint32_t)e, (uint32_t)(e >> 32) | ((uint32_t)j << 8), U_POOL_CENTER | (key.sub << 8), key.sweep, key.k0,
                key.k1, o);
  const double uc = u01_from_bits(o[0], o[1]), us = u01_from_bits(o[2], o[3]);
  int center = (int)((double)m * uc + 1.0);
  if (center > m) center = m;
  double uu;
  if (sigma_exact) {
    uu = hig_inv_u_d(us, v, w, (double)m);
  } else {
    RngKey k2 = key;
    k2.sweep = key.sweep ^ ((uint32_t)(e >> 32) << 20);  // entries beyond 2^32 (never at the shapes in scope)
    SubStream rs(k2, U_POOL_SIGMA, (uint32_t)e, (uint32_t)j);
    uu = hig_draw_u_d(rs, v, w, (double)m);
  }
  *center_out = center;
  *sigma_out = -1.0 / log(uu);
}

// Pool-free auxiliary components (smg_config.aux_mode == 1): the a-th auxiliary component of observation i in this
// pass is the prior draw of the virtual entry e = i * m_aux + a under the pass's own Philox key -- fresh for every
// observation and every pass, as Neal's Algorithm 8 states it, where the reference re-uses a stored pool of
// n * m_aux draws for 1000 iterations (launcher.cpp:67-77,123-129; neal8.cpp:65-69).  Nothing is stored but the
// column value and the entry's log-normaliser sum; the scan re-derives the parameters of the rare component that
// becomes a cluster.  One warp per (i, a): a lane draws attributes lane, lane + 32, ...
__global__ void __launch_bounds__(256) aux_ll_philox_kernel(const uint8_t* __restrict__ X, int n, int pp, int p, int m_aux,
                                                            const int* __restrict__ attr, const double* __restrict__ v,
                                                            const double* __restrict__ w, RngKey key, int sigma_exact,
                                                            double* __restrict__ LLaux, int* __restrict__ aux_e,
                                                            double* __restrict__ aux_sd) {
  long long wi = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wi >= (long long)n * m_aux) return;
  const int i = (int)(wi / m_aux);
  const uint8_t* x = X + (size_t)i * pp;
  double dot = 0.0, den = 0.0;
  for (int j = lane; j < p; j += 32) {
    const int m = attr[j];
    int center;
    double sigma;
    prior_entry_attr(key, wi, j, m, v[j], w[j], sigma_exact, &center, &sigma);
    if ((int)x[j] != center) dot += 1.0 / sigma;
    den += hamming_den(sigma, m);
  }
  dot = warp_sum(dot);
  den = warp_sum(den);
  if (lane == 0) {
    LLaux[wi] = -dot - den;
    aux_e[wi] = (int)wi;
    aux_sd[wi] = den;
  }
}

// Auxiliary-component columns: for every observation i and aux slot a, draw the pool entry
// (neal8.cpp:66: sample(pool_size,1)-1 = (int)(P*u+1)-1) and evaluate its log-likelihood.
__global__ void __launch_bounds__(256) aux_ll_kernel(const uint8_t* __restrict__ X, int n, int pp, int m_aux,
                                                     const uint8_t* __restrict__ pool_cen,
                                                     const double* __restrict__ pool_isg,
                                                     const double* __restrict__ pool_sden, long long pool_size,
                                                     const double* __restrict__ tape, int tape_stride, RngKey key,
                                                     double* __restrict__ LLaux, int* __restrict__ aux_e) {
  long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= (long long)n * m_aux) return;
  const int i = (int)(w / m_aux), a = (int)(w % m_aux);
  double u = get_u(tape, (size_t)i * tape_stride + a, key, U_POOL_IDX, (uint32_t)i, (uint32_t)a);
  long long e = (long long)((double)pool_size * u + 1.0) - 1;
  if (e >= pool_size) e = pool_size - 1;
  double dot = warp_mismatch_dot(X + (size_t)i * pp, pool_cen + (size_t)e * pp, pool_isg + (size_t)e * pp, pp, lane);
  if (lane == 0) {
    LLaux[w] = -dot - pool_sden[e];
    aux_e[w] = (int)e;
  }
}

// =============================================================================
// K2: Neal-8 allocation scan.  ONE resident CTA (32 warps).  The pass is sequential in
// the observation index, but an observation whose draw re-selects its current cluster
// changes nothing, so the CTA evaluates 32 consecutive observations per round against the
// same state (one warp each), finds the first one whose draw changes the state ("event"),
// applies it and restarts right after it.  Results are identical to the one-at-a-time scan.
//
// State kept in shared memory: member counts and their logs by slot, label<->slot maps.
// Existing columns come from the precomputed LL block; a column born during the pass
// (case 3/4) is evaluated on the fly from its parameter vector.
// =============================================================================
struct ScanArgs {
  int n, pp, m_aux, ldl, K0cap;   // K0cap: number of valid LL columns (= K at pass start)
  const uint8_t* X;
  double* LL;     // [n][ldl]; columns >= K0 are filled chunk by chunk for clusters born during the pass
  const double* LLaux;
  const double* mrg;  // [n] dominance margins under the start-of-pass counts (scan_margin_kernel)
  const uint8_t* und0;  // [n padded to 4096] 1 = margin within SCAN_FAST_DRIFT of the threshold (or below)
  int* und_blk;         // [blocks] number of flagged rows per block of SCAN_BLOCK observations; zeroed again at the end
  const int* aux_e;
  const double* u_alloc;  // injected allocation uniforms (tape + m_aux, stride), or null
  int u_stride;
  RngKey key;
  int* c;        // in: labels (== slots at pass start); out: slots (finalize maps back)
  uint8_t* cen;  // slot-indexed parameter arrays
  double* sig;
  double* isg;
  double* sden;
  const uint8_t* pool_cen;
  const double* pool_sig;
  const double* pool_isg;
  const double* pool_sden;
  // pool-free auxiliary components (aux_mode 1): parameters re-derived from (aux_key, entry) when one becomes a cluster
  int aux_free, p, sigma_exact;
  RngKey aux_key;
  const int* attr;
  const double *hv, *hw;
  const double* aux_sd;  // [n][m_aux] log-normaliser sums of the auxiliary components of this pass
  int* Kptr;      // in/out
  int* counts;    // in: by label; out: by slot
  int* slot2label;  // out [NS]
  int NS;
  double log_gamma_m;  // log(gamma/m)  (neal8.cpp:78)
  int* status;
  unsigned long long* stats;  // [0] rounds, [1] events, [2] births, [3] deaths
  int* job;  // [4] mailbox of the scan cluster: {slot (-1: exit), first row, end row}
  unsigned long long* prof;  // [8] optional cycle counters of the scanner's phases (thread 0), or null
  // speculative evaluation (see the scan kernel): 0 off, 1 on, 2 on and every speculated outcome checked by an exact
  // evaluation under the state the row meets (stats[4] counts the differences -- must stay 0)
  int spec;
  int spec_rmax;      // rows evaluated per warp and speculation at most (<= SCAN_SPEC_RMAX)
  double spec_dmax;   // drift (nats) after which the rest of a speculation is dropped and made again
};

#define EVT_NONE (-1)

// Materialise LL[r0..r1)[slot] for a column born during the pass.  The rows are dealt in blocks of 128
// (4 per warp in flight) to the `nparts` CTAs of the scan cluster; `part` is this CTA's rank.  The value of
// each entry is the one warp_mismatch_dot would return (same per-lane order, same butterfly).
__device__ __forceinline__ void scan_fill_column(const ScanArgs& A, int slot, int r0, int r1, int warp, int lane, int part,
                                                 int nparts) {
  const int pp = A.pp;
  const uint8_t* crow = A.cen + (size_t)slot * pp;
  const double* wrow = A.isg + (size_t)slot * pp;
  const double sd = A.sden[slot];
  for (int row = r0 + (part * SMG_SCAN_WARPS + warp) * 4; row < r1; row += nparts * SMG_SCAN_WARPS * 4) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int j0 = lane * 8; j0 < pp; j0 += 256) {
      const uint2 cv = *reinterpret_cast<const uint2*>(crow + j0);
      uint2 xv[4];
#pragma unroll
      for (int r = 0; r < 4; r++)
        xv[r] = (row + r < r1) ? *reinterpret_cast<const uint2*>(A.X + (size_t)(row + r) * pp + j0) : cv;
      const double2* w = reinterpret_cast<const double2*>(wrow + j0);
      const double2 w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
#pragma unroll
      for (int r = 0; r < 4; r++) {
        const uint32_t m0 = __vcmpne4(xv[r].x, cv.x), m1 = __vcmpne4(xv[r].y, cv.y);
        if (m0 & 0x000000ffu) acc[r] += w0.x;
        if (m0 & 0x0000ff00u) acc[r] += w0.y;
        if (m0 & 0x00ff0000u) acc[r] += w1.x;
        if (m0 & 0xff000000u) acc[r] += w1.y;
        if (m1 & 0x000000ffu) acc[r] += w2.x;
        if (m1 & 0x0000ff00u) acc[r] += w2.y;
        if (m1 & 0x00ff0000u) acc[r] += w3.x;
        if (m1 & 0xff000000u) acc[r] += w3.y;
      }
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const double dot = warp_sum(acc[r]);
      if (lane == 0 && row + r < r1) A.LL[(size_t)(row + r) * A.ldl + slot] = -dot - sd;
    }
  }
}

// thread-block-cluster plumbing of the scan kernel (rank 0 scans, the other CTAs only fill born columns)
#define SCAN_CLUSTER 4
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_cta_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}

// ---------------------------------------------------------------------------------------------
// Dominance margins.  Under the counts at the start of the pass, observation i keeps its cluster
// `own` with certainty when
//     mrg[i] = (log(n_own - 1) + LL[i][own]) - max(others: log n_k + LL[i][k],  aux: log(gamma/m) + LLaux[i][a])
// exceeds SCAN_DOMINANCE nats: every other entry then carries less than e^-44 of the own mass, the
// normalised own probability rounds to exactly 1.0 in fp64 (256 * e^-44 < 2^-54) and Rcpp::sample returns
// it for every u (neal8.cpp:95-102).  The scan keeps the test valid while counts drift (see below).
// One warp per observation, all SMs; reads the LL block once (8*(K+m) B per row), writes 8 B per row.
// ---------------------------------------------------------------------------------------------
#define SCAN_DOMINANCE 44.0
#define SCAN_SLACK 1e-6  // covers the rounding of the drift arithmetic
#define SCAN_FAST_DRIFT 0.25  // drift allowance of the precomputed screen flags (und0)
// Drift allowance of a block's own screen: the bit map of undecided rows is computed as if the counts had already drifted
// by this much more, and is kept across the moves applied inside the block until they have used it up.  (Without it every
// move re-screened up to 4096 rows: 56% of the scan's cycles on a chain with 7700 moves per pass.)
#define SCAN_RESCREEN_SLACK 0.05
#define SCAN_BLOCK_ROWS 4096   // = SCAN_BLOCK of the scan kernel

__global__ void __launch_bounds__(256) scan_margin_kernel(int n, const int* __restrict__ Kptr, int ldl, int m_aux,
                                                          const double* __restrict__ LL,
                                                          const double* __restrict__ LLaux, const int* __restrict__ c,
                                                          const int* __restrict__ counts, double log_gamma_m,
                                                          const double* u_alloc, int u_stride, RngKey key,
                                                          double* __restrict__ mrg, uint8_t* __restrict__ und0,
                                                          int* __restrict__ und_blk) {
  __shared__ double s_lc[SMG_MAX_ENTRIES], s_lcm1[SMG_MAX_ENTRIES];
  const int K = min(*Kptr, SMG_MAX_ENTRIES);
  for (int s = threadIdx.x; s < K; s += blockDim.x) {
    const int cnt = counts[s];
    s_lc[s] = cnt > 0 ? log((double)cnt) : -CUDART_INF;
    s_lcm1[s] = cnt > 1 ? log((double)(cnt - 1)) : -CUDART_INF;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int w0 = blockIdx.x * wpb + (threadIdx.x >> 5), nw = gridDim.x * wpb;
  // Two observations per warp and step (their loads are issued together); the warp maxima go through integer REDUX
  // on order-preserving keys instead of a butterfly of double-precision shuffles.
  for (int ib = w0; ib < n; ib += 2 * nw)
  for (int half = 0; half < 2; half++) {
    const int i = ib + half * nw;
    if (i >= n) break;
    const int own = c[i];
    const double* row = LL + (size_t)i * ldl;
    double a_own, best;
    if (K <= 64 && m_aux <= 32) {
      const int e0 = lane, e1 = lane + 32;
      const double v0 = e0 < K ? row[e0] : 0.0, v1 = e1 < K ? row[e1] : 0.0;
      const double xa = lane < m_aux ? LLaux[(size_t)i * m_aux + lane] : 0.0;
      double bl = -CUDART_INF, al = -CUDART_INF;
      if (e0 < K) {
        if (e0 == own) al = s_lcm1[e0] + v0; else bl = s_lc[e0] + v0;
      }
      if (e1 < K) {
        if (e1 == own) al = s_lcm1[e1] + v1; else bl = fmax(bl, s_lc[e1] + v1);
      }
      if (lane < m_aux) bl = fmax(bl, log_gamma_m + xa);
      best = key_to_double(warp_max_key(sort_key(bl)));
      a_own = shfl_d(al, own & 31);
    } else {
      a_own = -CUDART_INF, best = -CUDART_INF;
      for (int e = lane; e < K; e += 32) {
        const double ll = row[e];
        if (e == own)
          a_own = s_lcm1[e] + ll;
        else
          best = fmax(best, s_lc[e] + ll);
      }
      for (int a = lane; a < m_aux; a += 32) best = fmax(best, log_gamma_m + LLaux[(size_t)i * m_aux + a]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a_own = fmax(a_own, shfl_xor_d(a_own, o));
        best = fmax(best, shfl_xor_d(best, o));
      }
    }
    double mg = (a_own > -CUDART_INF) ? a_own - best : -CUDART_INF;
    if (a_own > -CUDART_INF && mg <= SCAN_DOMINANCE + 2.0 * SCAN_SLACK) {
      // Not dominated -- but the observation's uniform is known (counter-based stream, or the injected tape), and
      // with R = (mass of all other entries) / (own mass) the draw returns `own` iff own is the largest entry and
      // u <= 1/(1+R).  If every other entry gains at most D nats on the own one, R grows at most by e^D: the
      // outcome stays `own` while D < log(min(1, (1-u)/u) / R).  Stored as an equivalent margin, so the scan's drift
      // test covers it unchanged; only observations whose draw really is close to a boundary (or moves) are left
      // for the exact evaluation.
      double r = 0.0;
      for (int e = lane; e < K; e += 32)
        if (e != own) r += exp(s_lc[e] + row[e] - a_own);
      for (int a = lane; a < m_aux; a += 32) r += exp(log_gamma_m + LLaux[(size_t)i * m_aux + a] - a_own);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) r += shfl_xor_d(r, o);
      const double u = get_u(u_alloc, (size_t)i * u_stride, key, U_ALLOC, (uint32_t)i, 0u);
      const double room = fmin(1.0, (1.0 - u) / u);
      if (r > 0.0 && room > 0.0) {
        const double dtol = log(room / r) - 1e-5;  // the slack covers the rounding of r and of the exact evaluation
        if (dtol > 0.0) mg = fmax(mg, dtol + SCAN_DOMINANCE + SCAN_SLACK);
      }
    }
    if (lane == 0) {
      mrg[i] = mg;
      // the scan's screen, precomputed for as long as the counts have drifted by less than SCAN_FAST_DRIFT nats
      const bool flag = !(mg > SCAN_DOMINANCE + SCAN_SLACK + SCAN_FAST_DRIFT);
      und0[i] = flag;
      if (flag) atomicAdd(&und_blk[i / SCAN_BLOCK_ROWS], 1);  // flagged rows per scan block (zeroed by the scan kernel)
    }
  }
}

#define SCAN_CHUNK (SMG_SCAN_WARPS * 32)  // observations screened per step: one per thread

// Shared-memory state of the scan CTA (counts, their logs, label <-> slot maps, drift of the counts).
struct ScanState {
  int cnt[SMG_MAX_SLOTS];
  double logc[SMG_MAX_SLOTS];
  double logcm1[SMG_MAX_SLOTS];
  int l2s[SMG_MAX_SLOTS];
  int s2l[SMG_MAX_SLOTS];
  double lc0[SMG_MAX_ENTRIES];     // log n0_k of the clusters present at the start
  double lcm1_0[SMG_MAX_ENTRIES];  // log (n0_k - 1)
  double dminus[SMG_MAX_ENTRIES];
  double Dplus;
  double maxdm;  // monotone upper bound of dminus over the clusters that had more than one member at the start
  double scr_used;  // drift accumulated by the events applied since the current block's bit map was screened
  int evt[SMG_SCAN_WARPS];
  int row[SMG_SCAN_WARPS];
  unsigned und[4 * SMG_SCAN_WARPS];  // undecided rows of the current block of 4096 observations (bit per row)
  int K, next, err;
  int serial_next;  // first row after a serial stretch of warp 0
  unsigned long long stats[8];  // [4] speculation mismatches (spec = 2), [5] exact re-evaluations, [6] speculations dropped early
  // speculative evaluation: log-counts the current speculation was evaluated with (by label), the drift since, and
  // what the walk hands back to the block
  double base_lc[64], base_lcm1[64];
  double specD;
  int spec_keep, walk_done, walk_start, walk_reason;
  int arr[64], dep[64];  // arrivals / departures by label among the moves of a group of the walk (zero between uses)
  int sp_i0, sp_nb, sp_K;       // a speculation shared with the other CTAs of the cluster: first row of the block, rows, clusters
  int wk_cmd, wk_j0;            // job of the walk for the other warps: re-examine the rows `wk_need` of group wk_j0 / 32 (cmd 1: the walk is over)
  unsigned wk_need, wk_ok;
  double gsk[64];  // by label: slack of the cluster's log-counts over the rest of the current group of the walk
  double fac[64], facm1[64];  // by label: e^(drift of log n_k) and e^(drift of log(n_k - 1)) since the base (entries >= K stay 1)
};

// likelihood of observation i under a cluster whose column is not materialised (slot >= ldl: more births in
// one pass than spare LL columns) -- whole warp, rare
__device__ __noinline__ double scan_dyn_ll(const ScanArgs& A, int i, int slot, int lane) {
  const int pp = A.pp;
  const double dot = warp_mismatch_dot(A.X + (size_t)i * pp, A.cen + (size_t)slot * pp, A.isg + (size_t)slot * pp, pp, lane);
  return -dot - A.sden[slot];
}

// Exact allocation draw of observation i by one warp (neal8.cpp:40-102 + Rcpp::sample's descending-order
// inverse CDF).  Entry e = q*32 + lane; NQ*32 >= K + m_aux.  Returns EVT_NONE when the draw leaves the
// state unchanged, the selected entry (0..K+m-1) when it changes it, -2 when the weights are not finite.
//
// TOL: also returns in *tol how far (nats) the log-counts may drift from the ones used here before the OUTCOME can change
// (see "speculative evaluation" at the scan kernel); -1 when the draw must be repeated under the state it meets.
// wrow (TOL): the NQ*32 relative weights of the draw (0 beyond K + m) are left there, tol[1] = u, tol[2] = selected entry.
template <int NQ, bool TOL = false>
__device__ __forceinline__ int scan_eval_row(const ScanArgs& A, const ScanState& S, int i, int old_slot, int K, int lane,
                                             double* tol = nullptr, float* wrow = nullptr) {
  const int m = A.m_aux, ne = K + m;
  const double* rowp = A.LL + (size_t)i * A.ldl;
  const double* auxp = A.LLaux + (size_t)i * m;
  // the allocation uniform does not depend on anything loaded below: its Philox rounds overlap the loads
  const double u = get_u(A.u_alloc, (size_t)i * A.u_stride, A.key, U_ALLOC, (uint32_t)i, 0u);
  int slot[NQ];
  double ll[NQ];
  bool anydyn = false;
#pragma unroll
  for (int q = 0; q < NQ; q++) {
    const int e = q * 32 + lane;
    slot[q] = (e < K) ? S.l2s[e] : -1;
    ll[q] = 0.0;
    if (slot[q] >= 0 && slot[q] < A.ldl)
      ll[q] = __ldcg(&rowp[slot[q]]);
    else if (e >= K && e < ne)
      ll[q] = auxp[e - K];
    anydyn |= (slot[q] >= A.ldl);
  }
  if (__any_sync(SMG_FULL, anydyn)) {
#pragma unroll
    for (int q = 0; q < NQ; q++) {
      unsigned dyn = __ballot_sync(SMG_FULL, slot[q] >= A.ldl);
      while (dyn) {
        const int src = __ffs(dyn) - 1;
        dyn &= dyn - 1;
        const double v = scan_dyn_ll(A, i, __shfl_sync(SMG_FULL, slot[q], src), lane);
        if (lane == src) ll[q] = v;
      }
    }
  }
  const bool singleton = (S.cnt[old_slot] == 1);
  // ---- log weights: existing clusters (neal8.cpp:40-56), auxiliary components (neal8.cpp:78-92)
  double ll_own = 0.0;
#pragma unroll
  for (int q = 0; q < NQ; q++)
    if (slot[q] == old_slot) ll_own = ll[q];
  if (singleton) {  // own parameters become aux slot 0 (neal8.cpp:72-75)
    // ll_own is set by exactly one lane; LL is strictly negative so != 0 identifies it
    const unsigned who = __ballot_sync(SMG_FULL, ll_own != 0.0);
    ll_own = shfl_d(ll_own, who ? (__ffs(who) - 1) : 0);
  }
  double lg[NQ];
  uint64_t mykey = 0;
  int myarg = 0x7fffffff;
#pragma unroll
  for (int q = 0; q < NQ; q++) {
    const int e = q * 32 + lane;
    lg[q] = -CUDART_INF;
    if (slot[q] >= 0) {
      const bool own = (slot[q] == old_slot);
      const int cx = S.cnt[slot[q]] - (own ? 1 : 0);
      if (cx > 0) lg[q] = (own ? S.logcm1[slot[q]] : S.logc[slot[q]]) + ll[q];
    } else if (e >= K && e < ne) {
      lg[q] = A.log_gamma_m + ((singleton && e == K) ? ll_own : ll[q]);
    }
    if (e < ne) {
      const uint64_t k = sort_key(lg[q]);
      if (k > mykey) {
        mykey = k;
        myarg = e;
      }
    }
  }
  // ---- max, exp, sum  (neal8.cpp:95-96)
  const uint64_t maxkey = warp_max_key(mykey);
  const double M = key_to_double(maxkey);
  const int argmax = (int)__reduce_min_sync(SMG_FULL, (unsigned)((mykey == maxkey) ? myarg : 0x7fffffff));  // smallest index
  double pe[NQ];
  double lsum = 0.0;
#pragma unroll
  for (int q = 0; q < NQ; q++) {
    const int e = q * 32 + lane;
    pe[q] = 0.0;
    if (e < ne) {
      const double d = lg[q] - M;
      if (d == 0.0)
        pe[q] = 1.0;
      else if (d >= -SCAN_DOMINANCE)
        pe[q] = exp(d);
    }
    lsum += pe[q];
  }
  const double Ssum = warp_sum(lsum);
  const double T = u * Ssum;  // compared against the cumulative sums of the unnormalised weights
  int new_e;
  bool fell = false;  // selected by the fall-through of the reference loop
  if (!(M > -CUDART_INF) || !(Ssum == Ssum)) {
    if (TOL) {
      tol[0] = -1.0;
      tol[1] = u;
      tol[2] = -1.0;
    }
    return -2;  // all -Inf or NaN: Rcpp::sample would stop()
  } else if (T <= 1.0) {
    new_e = argmax;  // first entry of the descending order already covers u
  } else {
    // Rcpp::sample semantics: walk the probabilities in DESCENDING order (smallest index first among
    // ties) and return the first entry whose cumulative sum reaches u; the last entry if none does.
    int mysig = 0;
#pragma unroll
    for (int q = 0; q < NQ; q++) mysig += (pe[q] > 0.0);
    const int nsig = warp_sum_i(mysig);
    new_e = -1;
    if (nsig <= 16) {
      // few significant entries: extract the maxima one by one, accumulating in that order
      unsigned rem = 0;
#pragma unroll
      for (int q = 0; q < NQ; q++) rem |= (pe[q] > 0.0) ? (1u << q) : 0u;
      double cum = 0.0;
      for (int it = 0; it < nsig; it++) {
        uint64_t ck = 0;
        int ce = 0x7fffffff;
#pragma unroll
        for (int q = 0; q < NQ; q++) {
          const uint64_t k = (uint64_t)__double_as_longlong(pe[q]);  // positive doubles order like their bits
          if (((rem >> q) & 1u) && k > ck) {
            ck = k;
            ce = q * 32 + lane;
          }
        }
        const uint64_t wk = warp_max_key(ck);
        const int we = (int)__reduce_min_sync(SMG_FULL, (unsigned)((ck == wk) ? ce : 0x7fffffff));
        cum += __longlong_as_double((long long)wk);
        if (T <= cum) {
          new_e = we;
          break;
        }
        if ((we & 31) == lane) rem &= ~(1u << (we >> 5));
      }
    } else {
      // many significant entries (burn-in): cumulative sum of everything that precedes each entry
      double G[NQ];
#pragma unroll
      for (int q = 0; q < NQ; q++) G[q] = 0.0;
#pragma unroll
      for (int qy = 0; qy < NQ; qy++) {
        unsigned sig = __ballot_sync(SMG_FULL, pe[qy] > 0.0);
        while (sig) {
          const int src = __ffs(sig) - 1;
          sig &= sig - 1;
          const double py = shfl_d(pe[qy], src);
          const int ey = qy * 32 + src;
#pragma unroll
          for (int q = 0; q < NQ; q++) {
            const int ex = q * 32 + lane;
            if (py > pe[q] || (py == pe[q] && ey < ex)) G[q] += py;
          }
        }
      }
      // candidate = entry with u*S <= G + p, largest p first (smallest index among ties)
      uint64_t bestk = 0;
      int beste = 0x7fffffff;
#pragma unroll
      for (int q = 0; q < NQ; q++) {
        const int ex = q * 32 + lane;
        if (ex < ne && pe[q] > 0.0 && T <= G[q] + pe[q]) {
          const uint64_t k = sort_key(pe[q]);
          if (k > bestk || (k == bestk && ex < beste)) {
            bestk = k;
            beste = ex;
          }
        }
      }
      const uint64_t wk = warp_max_key(bestk);
      if (wk != 0) new_e = (int)__reduce_min_sync(SMG_FULL, (unsigned)((bestk == wk) ? beste : 0x7fffffff));
    }
    if (new_e < 0) {
      // fall-through of the reference loop: last entry of the descending order
      // (smallest probability, largest index among ties)
      fell = true;
      uint64_t mink = ~0ull;
      int mine = -1;
#pragma unroll
      for (int q = 0; q < NQ; q++) {
        const int ex = q * 32 + lane;
        if (ex < ne) {
          const uint64_t k = sort_key(pe[q]);
          if (k < mink || (k == mink && ex > mine)) {
            mink = k;
            mine = ex;
          }
        }
      }
      const uint64_t wmin = ~warp_max_key(~mink);
      new_e = (int)__reduce_max_sync(SMG_FULL, (unsigned)((mink == wmin) ? mine : 0));
    }
  }
  if (TOL) {
    // With w_e the weights (count x likelihood), G the sum of the weights that precede the selected entry e* in the
    // descending order and T = u * sum(w): e* is returned iff G < T <= G + w_e*.  If every log-count moves by at most
    // D <= Dc, every weight moves by a factor within e^-D .. e^D, and only an entry whose log-weight lies within 2 Dc
    // of e*'s can change place with it in the order.  With G_lo the sum over the entries that certainly precede e* and
    // P_A the sum over the ones that may, the outcome is the same as long as
    //   2D < log(T / (G_lo + P_A))   and   2D <= log((G_lo + w_e*) / T);
    // or, with no entry allowed to change place with e* (2D below the smallest gap |log w_f - log w_e*|), as long as
    //   2D < log(T / G)   and   2D <= log((G + w_e*) / T).
    // A slack of 1e-9 nats covers the rounding of the sums (and the e^-44 cut-off of negligible entries, < 1e-17).
    double tau = -1.0;
    if (!singleton && !fell) {
      const double Dc = fmin(A.spec_dmax, 0.01);
      double pstar = 0.0, lgstar = 0.0;
#pragma unroll
      for (int q = 0; q < NQ; q++)
        if (q == (new_e >> 5)) {
          pstar = pe[q];
          lgstar = lg[q];
        }
      pstar = shfl_d(pstar, new_e & 31);
      lgstar = shfl_d(lgstar, new_e & 31);
      // (two windows for "may change place": none -- then D must also stay below half the gap to the nearest
      // log-weight -- and 2 Dc; the larger of the two tolerances holds)
      double glo = 0.0, pa = 0.0, gex = 0.0, gap = CUDART_INF;
#pragma unroll
      for (int q = 0; q < NQ; q++) {
        const int e = q * 32 + lane;
        if (e < ne && e != new_e && lg[q] > -CUDART_INF) {
          const double dl = lg[q] - lgstar;
          if (dl > 2.0 * Dc)
            glo += pe[q];
          else if (dl >= -2.0 * Dc)
            pa += pe[q];
          if (pe[q] > pstar || (pe[q] == pstar && e < new_e)) gex += pe[q];
          gap = fmin(gap, fabs(dl));
        }
      }
      glo = warp_sum(glo);
      pa = warp_sum(pa);
      gex = warp_sum(gex);
      gap = key_to_double(~warp_max_key(~sort_key(gap)));
      const double up = log((glo + pstar) / T);
      const double lo = (glo + pa) > 0.0 ? log(T / (glo + pa)) : CUDART_INF;
      const double up1 = log((gex + pstar) / T);
      const double lo1 = gex > 0.0 ? log(T / gex) : CUDART_INF;
      tau = fmax(fmin(Dc, 0.5 * fmin(up, lo)), 0.5 * fmin(fmin(up1, lo1), gap)) - 1e-9;
      if (!(tau == tau)) tau = -1.0;
    }
    tol[0] = tau;
    tol[1] = u;
    tol[2] = (double)new_e;
#pragma unroll
    for (int q = 0; q < NQ; q++) wrow[q * 32 + lane] = (float)pe[q];
  }
  // ---- does the draw change the state?
  if (new_e < K) return (S.l2s[new_e] != old_slot) ? new_e : EVT_NONE;
  return (singleton && new_e == K) ? EVT_NONE : new_e;  // singleton re-drawing its own phi: no-op
}


// Shared-memory staging of the (slot, margin) pairs of one block of SCAN_BLOCK observations.  It is filled on demand
// with 16-byte loads, only for blocks whose rows need the per-row screen or the serial walk: with the precomputed
// screen flags a quiet block never reads it.  (A cp.async ring two blocks ahead was measured at ~5500 cycles per block
// of copy issue/queueing on the one scanner SM, more than everything else in a quiet block.)
#define SCAN_PF_DEPTH 4  // = SCAN_SUPER chunks of 1024 observations
#define SCAN_DENSE_ENTER 6    // events in a row at the head of their batch before warp 0 goes serial
#define SCAN_DENSE_LEAVE 48   // rows in a row without a move before it hands back to the block
#define SCAN_SUPER 4  // chunks of 1024 observations screened and evaluated together
#define SCAN_BLOCK (SCAN_SUPER * SCAN_CHUNK)
#define SCAN_BLKCNT_MAX 1024  // per-block flag counts kept in shared memory (larger n: read from global)
#define SCAN_SPEC_RMAX 4       // rows per warp in one speculation at most
#define SCAN_SPEC_ROWS (SCAN_SPEC_RMAX * SMG_SCAN_WARPS)
#define SCAN_SPEC_WSTRIDE 65   // floats per row of the cached weights (64 entries + 1)
// The cached weights (33 KB) lie OVER the staging buffers of the block (48 KB), which a speculation does not read and which
// are filled again on demand: with them beside the buffers the kernel needed the 164 KB shared-memory configuration instead
// of the 100 KB one, and the 64 KB of L1 that went with it slowed every phase of a quiet pass (spills) by a few percent.
#define SCAN_PF_BYTES (SCAN_PF_DEPTH * SCAN_CHUNK * 12 + SCAN_BLKCNT_MAX * 4 + SCAN_SPEC_ROWS * (8 + 4 + 8))
static_assert(SCAN_SPEC_ROWS * SCAN_SPEC_WSTRIDE * 4 <= SCAN_PF_DEPTH * SCAN_CHUNK * 12, "cached weights fit the staging buffers");
__device__ __forceinline__ void scan_cp_async4(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void scan_cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void scan_cp_async8(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}

// more than 64 entries per draw: rare, kept out of line so that its registers do not weigh on the scan kernel
__device__ __noinline__ int scan_eval_row_wide(const ScanArgs& A, const ScanState& S, int i, int old_slot, int K, int lane) {
  return scan_eval_row<SMG_EPL>(A, S, i, old_slot, K, lane);
}

// State update for a draw that moves observation `ie` from `old_slot` to the existing cluster of entry `new_e`
// (cases 1 and 2 of neal8.cpp:107-137).  One thread; the caller orders it against the readers of the state.
__device__ __forceinline__ void scan_apply_move(const ScanArgs& A, ScanState& S, int K0, int ie, int old_slot, int new_e) {
  const int Kc = S.K;
  const int new_slot = S.l2s[new_e];
  const bool singleton = (S.cnt[old_slot] == 1);
  S.stats[1]++;
  A.c[ie] = new_slot;
  // drift bookkeeping of a start-of-pass cluster whose count just changed
  auto drift = [&](int s) {
    if (s >= K0) return;
    S.dminus[s] = (S.cnt[s] > 1) ? (S.lcm1_0[s] - S.logcm1[s]) : CUDART_INF;
    if (S.lcm1_0[s] > -CUDART_INF && S.dminus[s] > S.maxdm) S.maxdm = S.dminus[s];
    const double dp = S.logc[s] - S.lc0[s];
    if (dp > S.Dplus) S.Dplus = dp;  // monotone upper bound
  };
  S.cnt[new_slot]++;
  S.logcm1[new_slot] = S.logc[new_slot];
  S.logc[new_slot] = log((double)S.cnt[new_slot]);
  drift(new_slot);
  if (!singleton) {  // case 1 (neal8.cpp:107-112)
    int c0 = --S.cnt[old_slot];
    S.logc[old_slot] = S.logcm1[old_slot];
    S.logcm1[old_slot] = c0 > 1 ? log((double)(c0 - 1)) : -CUDART_INF;
    drift(old_slot);
  } else {  // case 2 (neal8.cpp:115-137): last label moves into the hole
    S.cnt[old_slot] = 0;
    S.logc[old_slot] = S.logcm1[old_slot] = -CUDART_INF;
    int lab = S.s2l[old_slot];
    int last_slot = S.l2s[Kc - 1];
    S.l2s[lab] = last_slot;
    S.s2l[last_slot] = lab;
    S.s2l[old_slot] = -1;
    if (lab == Kc - 1) S.s2l[last_slot] = -1;  // the dying cluster was the last label
    S.K = Kc - 1;
    S.stats[3]++;
  }
}

// Speculative evaluation, second line of defence (see the scan kernel): does the draw of a speculated row still return the
// entry `sel` it was speculated to, under the current counts and with the weight of every cluster free to move by another
// factor within e^-s_k .. e^s_k (s_k = S.gsk[k], the slack of the clusters the rest of the group touches; `slack` false:
// under the current counts exactly)?  The weights of the base (w[0 .. 63], relative to the largest) are rescaled by the
// drift factors of the counts (own cluster: of n - 1); with G the sum of the weights above w_sel and T = u * sum(w):
//   no weight whose interval meets w_sel's,   max G < min T,   max T <= min (G + w_sel)
// (5e-7 relative on top: the cached weights are single precision, 6e-8 each; the exact evaluation decides whatever is
// closer than that).  One warp per row, two entries per lane; the answer is uniform over the warp.
__device__ __forceinline__ bool scan_recheck_row(const ScanState& S, const float* w, int sel, int own, double u, bool slack,
                                                 int K, int m, int lane) {
  if (sel < 0 || S.cnt[own] < 2) return false;
  const int oe = S.s2l[own], ne = K + m;
  // every weight within [w (1 - s - 5e-7), w (1 + s + s^2 + 5e-7)], s the slack of its cluster (e^s <= 1 + s + s^2 for s <= 1)
  double wlo[2], whi[2];
#pragma unroll
  for (int q = 0; q < 2; q++) {
    const int e = q * 32 + lane;
    const double wf = e < ne ? (double)w[e] * (e == oe ? S.facm1[e] : (e < K ? S.fac[e] : 1.0)) : 0.0;
    const double sk = (slack && e < K) ? S.gsk[e] : 0.0;
    wlo[q] = wf * (1.0 - sk - 5e-7);
    whi[q] = (sk <= 0.5) ? wf * (1.0 + sk * (1.0 + sk) + 5e-7) : CUDART_NAN;
  }
  const double slo = shfl_d((sel >> 5) ? wlo[1] : wlo[0], sel & 31), shi = shfl_d((sel >> 5) ? whi[1] : whi[0], sel & 31);
  double Glo = 0.0, Ghi = 0.0;
  bool amb = false;
#pragma unroll
  for (int q = 0; q < 2; q++) {
    const int e = q * 32 + lane;
    if (e < ne && e != sel) {
      if (wlo[q] > shi) {  // certainly before sel in the descending order
        Glo += wlo[q];
        Ghi += whi[q];
      } else if (!(whi[q] < slo)) {  // not certainly after it (or not a number): may change place
        amb = true;
      }
    }
  }
  const double Tlo = u * warp_sum(wlo[0] + wlo[1]), Thi = u * warp_sum(whi[0] + whi[1]);
  Glo = warp_sum(Glo);
  Ghi = warp_sum(Ghi);
  if (__any_sync(SMG_FULL, amb)) return false;
  return slo > 0.0 && (Ghi == 0.0 || Tlo > Ghi) && (Thi <= Glo + slo);
}
__device__ __forceinline__ void scan_walk_barrier() { asm volatile("bar.sync 1, 1024;" ::: "memory"); }  // all warps of the scan CTA

// =============================================================================
// K2: the scan proper (one resident CTA of 32 warps).
//
// Chunks of 1024 consecutive observations.  (1) Screen: each thread tests its observation's margin
// against the drift of the counts since the start of the pass,
//     mrg[i] - dminus[own] - Dplus > SCAN_DOMINANCE,    dminus[k] = log(n0_k - 1) - log(n_k - 1),
//                                                        Dplus >= max_k (log n_k - log n0_k),
// and against every cluster born during the pass (whose columns are materialised chunk by chunk).
// Observations that pass are certain non-events.  (2) The others ("undecided") are evaluated exactly,
// 32 per round (one warp each) against the same state: Rcpp::sample's descending-order inverse CDF.
// The first one whose draw changes the state (an event) is applied, the rest of the chunk is screened
// again, and evaluation restarts right after it -- the result is the one-at-a-time scan's.
//
// The kernel runs as ONE thread-block cluster of SCAN_CLUSTER CTAs.  Rank 0 does everything above.  When a
// cluster is born (neal8.cpp:140-159) its likelihood column is needed for every later observation: rank 0
// posts {slot, rows} in a mailbox and the whole cluster evaluates the column between two cluster barriers
// (n*pp bytes of X spread over SCAN_CLUSTER SMs instead of one).  The other ranks sleep in the barrier.
// =============================================================================
__global__ void __cluster_dims__(SCAN_CLUSTER, 1, 1) __launch_bounds__(SMG_SCAN_WARPS * 32, 1)
    neal8_scan_kernel(ScanArgs A) {
  __shared__ ScanState S;
  // staging buffers of the current block (see SCAN_PF_DEPTH) and results of a speculation, by ordinal of the row among
  // the undecided rows it covers
  extern __shared__ __align__(16) unsigned char s_ring[];
  float* spec_w = reinterpret_cast<float*>(s_ring);  // [SCAN_SPEC_ROWS][SCAN_SPEC_WSTRIDE] relative weights, over the staging buffers
  double* spec_u = reinterpret_cast<double*>(s_ring + SCAN_PF_DEPTH * SCAN_CHUNK * 12 + SCAN_BLKCNT_MAX * 4);  // [SCAN_SPEC_ROWS] allocation uniforms
  float* spec_tau = reinterpret_cast<float*>(spec_u + SCAN_SPEC_ROWS);                                         // [SCAN_SPEC_ROWS]
  short* spec_row = reinterpret_cast<short*>(spec_tau + SCAN_SPEC_ROWS);                                       // [SCAN_SPEC_ROWS]
  short* spec_code = spec_row + SCAN_SPEC_ROWS;                                                                // [SCAN_SPEC_ROWS]
  short* spec_own = spec_code + SCAN_SPEC_ROWS;                                                                // [SCAN_SPEC_ROWS]
  short* spec_sel = spec_own + SCAN_SPEC_ROWS;                                                                 // [SCAN_SPEC_ROWS] selected entry (-1: none)
  // one speculated row: exact draw of observation i0 + myrow against the state St, results into the arrays of the scan CTA
  auto spec_eval = [&](const ScanState& St, int j, int myrow, int i0, int K, int lane_, float* w_, double* u_, float* tau_,
                       short* code_, short* own_, short* sel_) {
    const int i = i0 + myrow;
    const int old_slot = __ldcg(&A.c[i]);
    double tol3[3];
    const int code = scan_eval_row<2, true>(A, St, i, old_slot, K, lane_, tol3, w_ + j * SCAN_SPEC_WSTRIDE);
    if (lane_ == 0) {
      code_[j] = (short)code;
      own_[j] = (short)old_slot;
      sel_[j] = (short)(tol3[0] == -1.0 ? -1 : (int)tol3[2]);  // (-1: must be repeated under the state it meets)
      u_[j] = tol3[1];
      tau_[j] = __double2float_rd(tol3[0]);
    }
  };
  if (cluster_cta_rank() != 0) {  // ---- helpers: columns of clusters born in the pass, a share of the speculative evaluations
    const int lane_h = threadIdx.x & 31, warp_h = threadIdx.x >> 5;
    const int rank_h = (int)cluster_cta_rank();
    cooperative_groups::cluster_group cl = cooperative_groups::this_cluster();
    for (;;) {
      cluster_sync_all();  // the mailbox (and the new parameter vector / the scan CTA's state) are published
      const int slot = __ldcg(&A.job[0]), r0 = __ldcg(&A.job[1]), r1 = __ldcg(&A.job[2]);
      if (slot == -1) return;
      if (slot == -2) {
        // rows [32 rank, 32 rank + 32) of the speculation, one per warp, against a copy of the scan CTA's counts
        const ScanState* S0 = cl.map_shared_rank(&S, 0);
        const int i0 = S0->sp_i0, nb = S0->sp_nb, K = S0->sp_K;
        for (int e = threadIdx.x; e < K; e += blockDim.x) {
          const int sl = S0->l2s[e];
          S.l2s[e] = sl;
          S.cnt[sl] = S0->cnt[sl];
          S.logc[sl] = S0->logc[sl];
          S.logcm1[sl] = S0->logcm1[sl];
        }
        __syncthreads();
        const int j = rank_h * SMG_SCAN_WARPS + warp_h;
        if (j < nb)
          spec_eval(S, j, (int)cl.map_shared_rank(spec_row, 0)[j], i0, K, lane_h, cl.map_shared_rank(spec_w, 0),
                    cl.map_shared_rank(spec_u, 0), cl.map_shared_rank(spec_tau, 0), cl.map_shared_rank(spec_code, 0),
                    cl.map_shared_rank(spec_own, 0), cl.map_shared_rank(spec_sel, 0));
      } else {
        scan_fill_column(A, slot, r0, r1, warp_h, lane_h, rank_h, SCAN_CLUSTER);
      }
      cluster_sync_all();  // the column is complete / the results are in the scan CTA's shared memory
    }
  }

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = A.n, pp = A.pp, m = A.m_aux;
  const int K0 = *A.Kptr;

  for (int s = tid; s < SMG_MAX_SLOTS; s += blockDim.x) {
    const int cnt = (s < K0) ? A.counts[s] : 0;
    const double lc = cnt > 0 ? log((double)cnt) : -CUDART_INF;
    const double lcm1 = cnt > 1 ? log((double)(cnt - 1)) : -CUDART_INF;
    S.cnt[s] = cnt;
    S.logc[s] = lc;
    S.logcm1[s] = lcm1;
    S.l2s[s] = s;
    S.s2l[s] = (s < K0) ? s : -1;
    if (s < 64) {
      S.arr[s] = S.dep[s] = 0;
      S.fac[s] = S.facm1[s] = 1.0;
    }
    if (s < SMG_MAX_ENTRIES) {
      S.lc0[s] = lc;
      S.lcm1_0[s] = lcm1;
      S.dminus[s] = (cnt > 1) ? 0.0 : CUDART_INF;
    }
  }
  if (tid == 0) {
    S.K = K0;
    S.next = K0;
    S.err = 0;
    S.Dplus = 0.0;
    S.maxdm = 0.0;
    for (int q = 0; q < 8; q++) S.stats[q] = 0;
    if (K0 + m > SMG_MAX_ENTRIES) S.err |= ST_TOO_MANY_ENTRIES;
    if (K0 > A.K0cap) S.err |= ST_LL_COLS;
  }
  __syncthreads();
  bool abort_pass = false;
  int W = SMG_SCAN_WARPS;  // rows evaluated per round
  int dense_run = 0;        // consecutive events found at the head of their batch
  // phase cycle counters of thread 0 (compile with -DSMG_SCAN_PROFILE): [0] chunk prologue, [1] screen,
  // [2] batch pick, [3] evaluation, [4] event detection, [5] event application, [6] whole loop
#ifdef SMG_SCAN_PROFILE
  long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tmark = clock64();
  const long long tstart = tmark;
#define SCAN_TICK(k)                 \
  do {                               \
    const long long _t = clock64();  \
    pc[k] += _t - tmark;             \
    tmark = _t;                      \
  } while (0)
  long long wkc[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // the walk's: [0] group bounds, [1] prefix commit, [2] exact re-evaluation, [3] single move, [4..6] counts
  long long wmark = 0;
#define WALK_MARK() wmark = clock64()
#define WALK_TICK(k)                 \
  do {                               \
    const long long _t = clock64();  \
    wkc[k] += _t - wmark;            \
    wmark = _t;                      \
  } while (0)
#define WALK_COUNT(k) wkc[k]++
#else
#define SCAN_TICK(k)
#define WALK_MARK()
#define WALK_TICK(k)
#define WALK_COUNT(k)
#endif
  if (S.err) {
    if (tid == 0) atomicOr(A.status, S.err);
    abort_pass = true;
  }

  double* ring_mg = reinterpret_cast<double*>(s_ring);                                    // [SCAN_BLOCK]
  int* ring_own = reinterpret_cast<int*>(s_ring + SCAN_PF_DEPTH * SCAN_CHUNK * 8);         // [SCAN_BLOCK]
  int* s_blkcnt = reinterpret_cast<int*>(s_ring + SCAN_PF_DEPTH * SCAN_CHUNK * 12);       // [SCAN_BLKCNT_MAX]
  int R = 1;                // rows per warp of the next speculation (adapts)
  int hot = 0;              // > 0 while moves are being found: a quiet stretch (a handful of undecided rows per block, none of
                            // which moves) goes through the plain rounds, which cost less when nothing is applied
  bool one_legacy = false;  // the next round goes through the one-row-per-warp path (an event the walk does not apply)
  for (int b = tid; b < SCAN_BLKCNT_MAX && (long long)b * SCAN_BLOCK < n; b += blockDim.x) s_blkcnt[b] = A.und_blk[b];
  __syncthreads();

  // can observation i (slot `own`, margin `mg`) be anything but a certain non-event under the current state?
  // `extra`: additional drift (nats) the verdict "certain non-event" must survive -- the screen of a block is kept across
  // the events applied inside it for as long as they have used less than that (SCAN_RESCREEN_SLACK)
  auto undecided = [&](int i, int own, double mg, double extra) -> bool {
    // own slot is always a start-of-pass cluster for a row the scan has not reached yet
    if (!(mg - S.dminus[own] - S.Dplus > SCAN_DOMINANCE + SCAN_SLACK + extra)) return true;
    if (S.next > K0) {  // clusters born during this pass: compare with their materialised columns
      const double thr = S.logcm1[own] + A.LL[(size_t)i * A.ldl + own] - (SCAN_DOMINANCE + SCAN_SLACK + extra);
      const int K = S.K;
      for (int e = 0; e < K; e++) {
        const int slot = S.l2s[e];
        if (slot < K0) continue;
        if (slot >= A.ldl || !(S.logc[slot] + __ldcg(&A.LL[(size_t)i * A.ldl + slot]) < thr)) return true;
      }
    }
    return false;
  };

  // Blocks of SCAN_SUPER ring slots (SCAN_BLOCK = 4096 observations): thread t screens rows t, t+1024, ...; the
  // undecided rows of the whole block are evaluated together, so a quiet block costs one barrier and a block with a
  // handful of undecided rows one evaluation round instead of one per 1024 rows.
  static_assert(SCAN_BLOCK == SCAN_BLOCK_ROWS, "block size of the precomputed screen");
  for (int i0 = 0, blk = 0; i0 < n && !abort_pass; i0 += SCAN_BLOCK, blk++) {
    const int nrows = min(SCAN_BLOCK, n - i0);
    const int cnt_cur = blk < SCAN_BLKCNT_MAX ? s_blkcnt[blk] : A.und_blk[blk];
    // a block without a flagged row, under a state the flags are valid for: nothing to do (every thread reads the
    // same shared values; they only change in the event handling, between barriers)
    if (cnt_cur == 0 && S.next == K0 && S.maxdm + S.Dplus <= SCAN_FAST_DRIFT) {
      if (tid == 0) S.stats[0]++;
      continue;
    }
    const unsigned und_cur = *reinterpret_cast<const unsigned*>(A.und0 + i0 + 4 * tid);  // flags of rows 4*tid.. (und0 is padded)
    // The staging buffers are filled only when somebody reads them (block-uniform call sites).  mrg and c are padded,
    // and rows the scan has not reached still hold their start-of-pass slot in c.
    bool ring_ready = false;
    auto ring_wait = [&]() {
      if (ring_ready) return;
#pragma unroll
      for (int k = 0; k < SCAN_BLOCK / 2 / (SMG_SCAN_WARPS * 32); k++) {
        const int idx = 2 * (k * SMG_SCAN_WARPS * 32 + tid);
        if (i0 + idx < n) *reinterpret_cast<double2*>(&ring_mg[idx]) = __ldcg(reinterpret_cast<const double2*>(A.mrg + i0 + idx));
      }
#pragma unroll
      for (int k = 0; k < SCAN_BLOCK / 4 / (SMG_SCAN_WARPS * 32); k++) {
        const int idx = 4 * (k * SMG_SCAN_WARPS * 32 + tid);
        if (i0 + idx < n) *reinterpret_cast<int4*>(&ring_own[idx]) = __ldcg(reinterpret_cast<const int4*>(A.c + i0 + idx));
      }
      __syncthreads();
      ring_ready = true;
    };
    const int slot0 = 0;
    SCAN_TICK(0);
    int start = 0;          // rows [0, start) of the block are final
    bool screened = false;  // the undecided set below is valid for the current state
    int total_und = 0, consumed = 0;  // undecided rows from `start` on / already evaluated without an event
    int scr_kind = 0;                 // how the current bit map was made: 1 from the precomputed flags, 2 by the per-row screen
    int scr_rows = 0;                 // rows [start, scr_rows) are covered by the current screen
    bool scr_all = false;             // every row the current screen covers is marked undecided: no drift can make it wrong
    for (;;) {
      // ================= serial stretch: one warp, no block barriers =================
      // While nearly every observation moves (burn-in from a random start) speculation over a batch buys nothing
      // and every event costs several block barriers.  After SCAN_DENSE_ENTER events in a row at the head of their
      // batch, warp 0 walks the rows one at a time on its own -- evaluate, apply, next -- until a stretch of
      // SCAN_DENSE_LEAVE rows without a move, the end of the block, or a draw the block must handle together
      // (a new cluster: its column is filled by the whole cluster of CTAs; an error).
      if (!A.spec && dense_run >= SCAN_DENSE_ENTER && S.K + m <= 64) {
        ring_wait();
        __syncthreads();
        if (warp == 0) {
          int r = start, calm = 0;
          while (r < nrows && calm < SCAN_DENSE_LEAVE) {
            const int own = ring_own[(slot0 + r / SCAN_CHUNK) * SCAN_CHUNK + (r % SCAN_CHUNK)];
            const double mg = ring_mg[(slot0 + r / SCAN_CHUNK) * SCAN_CHUNK + (r % SCAN_CHUNK)];
            int code = EVT_NONE;
            if (undecided(i0 + r, own, mg, 0.0)) {
              const int K = S.K;
              if (K + m > 64) break;
              code = scan_eval_row<2>(A, S, i0 + r, own, K, lane);
              if (code != EVT_NONE && (code < 0 || code >= K)) break;  // left to the block: row r is evaluated again
            }
            if (code == EVT_NONE) {
              calm++;
            } else {
              if (lane == 0) {
                S.stats[0]++;
                scan_apply_move(A, S, K0, i0 + r, own, code);
              }
              __syncwarp();
              calm = 0;
            }
            r++;
          }
          if (lane == 0) S.serial_next = r;
        }
        __syncthreads();
        start = S.serial_next;
        dense_run = 0;
        W = SMG_SCAN_WARPS;
        screened = false;
        if (start >= nrows) {
          if (tid == 0) S.stats[0]++;
          break;
        }
      }
      // ================= screen: up to SCAN_SUPER rows per thread =================
      // In quiet stretches the whole block is screened at once; while events are dense (W small: every event
      // invalidates the screen) only the 1024-row slice that holds `start` is, the later ones when they are reached.
      if (!screened) {
        // While no cluster was born in this pass and the counts have drifted by less than SCAN_FAST_DRIFT nats, the
        // flags written by scan_margin_kernel ARE the screen: 4 rows per thread, no per-row arithmetic.
        const bool fast = (S.next == K0) && (S.maxdm + S.Dplus <= SCAN_FAST_DRIFT);  // (start-of-pass singletons are flagged anyway)
        const int q0 = start / SCAN_CHUNK;
        const int q1 = (fast || W == SMG_SCAN_WARPS) ? SCAN_SUPER : min(SCAN_SUPER, q0 + 1);
        if (fast) {
          const int r4 = tid * 4;
          unsigned nib = 0;
          if (r4 < nrows) {
            const unsigned f = und_cur;
#pragma unroll
            for (int b = 0; b < 4; b++)
              if (((f >> (8 * b)) & 0xffu) && r4 + b >= start && r4 + b < nrows) nib |= 1u << b;
          }
          unsigned wv = nib << (4 * (lane & 7));
          wv |= __shfl_xor_sync(SMG_FULL, wv, 1);
          wv |= __shfl_xor_sync(SMG_FULL, wv, 2);
          wv |= __shfl_xor_sync(SMG_FULL, wv, 4);
          if ((lane & 7) == 0) S.und[tid >> 3] = wv;  // word w covers rows [32w, 32w + 32) of the block
        } else {
        ring_wait();
#pragma unroll
        for (int q = 0; q < SCAN_SUPER; q++) {
          const int r = q * SCAN_CHUNK + tid;
          bool und = false;
          if (q >= q0 && q < q1 && r >= start && r < nrows)
            und = undecided(i0 + r, ring_own[(slot0 + q) * SCAN_CHUNK + tid], ring_mg[(slot0 + q) * SCAN_CHUNK + tid],
                            SCAN_RESCREEN_SLACK);
          const unsigned b = __ballot_sync(SMG_FULL, und);
          if (lane == 0) S.und[q * SMG_SCAN_WARPS + warp] = b;  // word w covers rows [32w, 32w + 32) of the block
        }
        }
        scr_rows = min(nrows, q1 * SCAN_CHUNK);
        screened = true;
        scr_kind = fast ? 1 : 2;
        consumed = 0;
        if (tid == 0) S.scr_used = 0.0;
        __syncthreads();
        int pc = 0;
#pragma unroll
        for (int k = 0; k < SCAN_SUPER; k++) pc += __popc(S.und[SCAN_SUPER * lane + k]);
        total_und = warp_sum_i(pc);
        scr_all = total_und == scr_rows - start;
      }
      SCAN_TICK(1);
      if (consumed >= total_und) {  // every screened row from `start` on is final
        if (scr_rows >= nrows) {
          if (tid == 0) S.stats[0]++;
          break;
        }
        start = scr_rows;  // go on with the next slice
        screened = false;
        __syncthreads();   // S.und is rewritten
        continue;
      }

      // ================= speculative evaluation =================
      // Up to R undecided rows per warp are evaluated against the SAME state (the "base"), each with the drift of the
      // log-counts its outcome tolerates (scan_eval_row<.., true>).  Warp 0 then walks the results in row order:
      // a speculated move is applied when the drift accumulated since the base is below the row's tolerance, a
      // speculated non-event is skipped under the same condition, and a row whose tolerance is used up is evaluated
      // again, exactly, under the state it meets -- so the result is the one-at-a-time scan's, and a move costs one
      // state update instead of a round of the whole block.  The walk hands back to the block for anything but a
      // plain move (birth, last member leaving, error), when the block's screen has to be redone, and when the drift
      // makes the remaining speculation not worth keeping (it is then made again from the new state).
      if (A.spec && (hot > 0 || A.spec == 2) && !one_legacy && S.K + m <= 64) {
        const int K = S.K;
        const int nb = min(R * SMG_SCAN_WARPS, total_und - consumed);
        if (tid < K) {
          const int sl = S.l2s[tid];
          S.base_lc[tid] = S.logc[sl];
          S.base_lcm1[tid] = S.logcm1[sl];
        }
        if (tid < 64) S.fac[tid] = S.facm1[tid] = 1.0;
        {
          // lane l owns words SCAN_SUPER*l .. of the undecided bit map; ordinal -> row as in the batch path below
          unsigned wb[SCAN_SUPER];
          int wpc = 0;
#pragma unroll
          for (int k = 0; k < SCAN_SUPER; k++) {
            wb[k] = S.und[SCAN_SUPER * lane + k];
            wpc += __popc(wb[k]);
          }
          int incl = wpc;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(SMG_FULL, incl, o);
            if (lane >= o) incl += y;
          }
          const int wpref = incl - wpc;
          for (int j = warp; j < nb; j += SMG_SCAN_WARPS) {
            const int t = consumed + j;
            int found = -1;
            if (t >= wpref && t < wpref + wpc) {
              int off = t - wpref;
#pragma unroll
              for (int k = 0; k < SCAN_SUPER; k++) {
                const int c = __popc(wb[k]);
                if (found < 0 && off >= 0 && off < c) {
                  unsigned bits = wb[k];
                  for (int q = 0; q < off; q++) bits &= bits - 1;  // drop the `off` lowest set bits
                  found = (SCAN_SUPER * lane + k) * 32 + __ffs(bits) - 1;
                }
                off -= c;
              }
            }
            const unsigned hit = __ballot_sync(SMG_FULL, found >= 0);
            const int myrow = __shfl_sync(SMG_FULL, found, __ffs(hit) - 1);
            if (lane == 0) spec_row[j] = (short)myrow;
          }
        }
        // more than one row per warp: the other CTAs of the cluster take rows 32 .. nb - 1 (one per warp there too)
        const bool wide = nb > SMG_SCAN_WARPS;
        if (wide) {
          if (tid == 0) {
            S.sp_i0 = i0;
            S.sp_nb = nb;
            S.sp_K = K;
            A.job[0] = -2;
          }
          cluster_sync_all();
        } else {
          __syncwarp();
        }
        if (warp < nb) spec_eval(S, warp, (int)spec_row[warp], i0, K, lane, spec_w, spec_u, spec_tau, spec_code, spec_own, spec_sel);
        if (wide) cluster_sync_all();
        SCAN_TICK(3);
        __syncthreads();
        SCAN_TICK(4);
#ifdef SMG_SCAN_PROFILE
        pc[7] += nb;
#endif
        if (warp == 0) {
          double D = 0.0;  // largest drift of a log-count since the base (running maximum)
          int nfrag = 0, done = nb, next_start = -1, reason = 0;
          auto dd = [](double x, double y) { return x == y ? 0.0 : fabs(x - y); };
          // e^x for the drift of a log-count (8th-order series: error < 1e-14 for |x| <= 0.125; NaN beyond -- and for a
          // count that fell to or rose from zero --, which fails every comparison of scan_recheck_row)
          auto drift_factor = [](double now, double base) {
            if (now == base) return 1.0;
            const double x = now - base;
            if (!(fabs(x) <= 0.125)) return CUDART_NAN;
            double r = 1.0 + x * (1.0 / 8.0);
            r = 1.0 + x * (1.0 / 7.0) * r;
            r = 1.0 + x * (1.0 / 6.0) * r;
            r = 1.0 + x * (1.0 / 5.0) * r;
            r = 1.0 + x * (1.0 / 4.0) * r;
            r = 1.0 + x * (1.0 / 3.0) * r;
            r = 1.0 + x * (1.0 / 2.0) * r;
            return 1.0 + x * r;
          };
          for (int g = 0; g * 32 < nb && !reason; g++) {
            const int jl = g * 32 + lane;
            const bool have = jl < nb;
            const int code_l = have ? (int)spec_code[jl] : EVT_NONE;
            const double tau_l = have ? (double)spec_tau[jl] : CUDART_INF;
            const int row_l = have ? (int)spec_row[jl] : -1;
            const int own_l = have ? (int)spec_own[jl] : 0;
            // a speculated move into an existing cluster ("plain"): the only kind the walk applies itself
            const bool plain_l = have && code_l >= 0 && code_l < K;
            const int ns_l = plain_l ? S.l2s[code_l] : 0;
            const int oe_l = plain_l ? S.s2l[own_l] : 0;
            unsigned live = __ballot_sync(SMG_FULL, have);
            // bounds of the group (below); they hold until a row is decided by an exact evaluation, whose outcome may
            // differ from the speculated one they were made with
            bool bounds_valid = false;
            double dmb = 0.0, dpb = 0.0, usedb = 0.0, Bg = 0.0;
            // outcome of this lane's recheck under the current bounds: the slacks cover every later state of the group, so a row
            // is re-examined once per set of bounds, not once per applied prefix
            bool rechecked_l = false, recheck_ok_l = false;
            WALK_MARK();
            while (live && !reason) {
              const bool live_l = (live >> lane) & 1u;
              WALK_COUNT(4);
              // ---- (a) as many of the next rows as possible at once.  With a_k arrivals and d_k departures among the
              // live moves of this group, the count of cluster k stays within [c_k - d_k, c_k + a_k] whatever the order,
              // so its log-counts stay within  |now - base| + max(a/c, d/(c - d))  of the base: Dg bounds the drift
              // every row of the group can meet.  The rows before the first one whose tolerance is below Dg (or that
              // is not a plain move) are final as speculated: their moves are applied together.
              // (a_k and d_k count every move of the group that is still to come when they are made; the ones applied
              // since are part of them, so the bounds hold for the rest of the group.)
              const bool pl = plain_l && live_l;
              if (!bounds_valid) {
              if (pl) {
                atomicAdd(&S.arr[code_l], 1);
                atomicAdd(&S.dep[oe_l], 1);
              }
              __syncwarp();
              double bound = 0.0;
              dmb = dpb = usedb = 0.0;
              rechecked_l = recheck_ok_l = false;
              S.gsk[lane] = S.gsk[lane + 32] = 0.0;
              __syncwarp();
              if (pl) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                  const int sl = h ? own_l : ns_l, e = h ? oe_l : code_l;
                  const int c = S.cnt[sl], a = S.arr[e], d = S.dep[e], lo = c - d;
                  if (lo < 2) {
                    bound = CUDART_INF;  // a cluster could get down to one member: one row at a time
                    S.gsk[e] = CUDART_INF;
                  } else {
                    // (upper bounds: single-precision reciprocals rounded up, times 1 + 2^-20 for the products)
                    const float fa = (float)a * 1.000001f, fd = (float)d * 1.000001f;
                    const double up0 = fa * __frcp_ru((float)c), dn0 = fd * __frcp_ru((float)lo);
                    const double up1 = fa * __frcp_ru((float)(c - 1)), dn1 = fd * __frcp_ru((float)(lo - 1));
                    bound = fmax(bound, dd(S.logc[sl], S.base_lc[e]) + fmax(up0, dn0));
                    bound = fmax(bound, dd(S.logcm1[sl], S.base_lcm1[e]) + fmax(up1, dn1));
                    // slack between ANY two states of the rest of the group (the rechecks rescale from the state they
                    // meet, not from the one the bounds were made in): at any time the count is >= lo = c - d, at most a
                    // arrivals and d departures are still to come, so |log n' - log n_t| <= max(a, d) / lo (n - 1: lo - 1)
                    const double sk = (double)(fmaxf(fa, fd) * __frcp_ru((float)(lo - 1)));
                    S.gsk[e] = sk;  // (every lane that touches label e writes the same value)
                    // what the block's screen is told (upper bounds of the running maxima the one-at-a-time updates keep)
                    if (sl < K0) {
                      if (S.lcm1_0[sl] > -CUDART_INF) dmb = fmax(dmb, S.lcm1_0[sl] - S.logcm1[sl] + dn1);
                      dpb = fmax(dpb, S.logc[sl] - S.lc0[sl] + up0);
                    }
                    if (h)
                      usedb += 1.000001f * __frcp_ru((float)(lo - 1));
                    else if (sl >= K0)
                      usedb += 1.000001f * __frcp_ru((float)c);
                  }
                }
              }
              __syncwarp();
              if (pl) {
                S.arr[code_l] = 0;
                S.dep[oe_l] = 0;
              }
              Bg = key_to_double(warp_max_key(sort_key(bound)));
              bounds_valid = true;
              }
              const double Dg = fmax(D, Bg);
              WALK_TICK(0);
              bool ok_l = tau_l > Dg || recheck_ok_l;
              unsigned need = __ballot_sync(SMG_FULL, live_l && !ok_l && !rechecked_l && (code_l == EVT_NONE || plain_l));
              if (need) {
                if (__popc(need) >= 3) {
                  // one row per warp, the whole CTA: the other warps wait for this in scan_walk_barrier()
                  if (lane == 0) {
                    S.wk_cmd = 0;
                    S.wk_j0 = g * 32;
                    S.wk_need = need;
                    S.wk_ok = 0u;
                  }
                  scan_walk_barrier();
                  {
                    const int src = __ffs(need) - 1;  // warp 0 takes the first one
                    const int j = g * 32 + src;
                    if (scan_recheck_row(S, spec_w + j * SCAN_SPEC_WSTRIDE, (int)spec_sel[j], (int)spec_own[j], spec_u[j], true, K, m, lane) && lane == 0)
                      atomicOr(&S.wk_ok, 1u << src);
                  }
                  scan_walk_barrier();
                  if ((need >> lane) & 1u) {
                    rechecked_l = true;
                    ok_l = recheck_ok_l = (S.wk_ok >> lane) & 1u;
                  }
                } else {
                  while (need) {
                    const int src = __ffs(need) - 1;
                    need &= need - 1;
                    const int j = g * 32 + src;
                    const bool ok = scan_recheck_row(S, spec_w + j * SCAN_SPEC_WSTRIDE, (int)spec_sel[j], (int)spec_own[j], spec_u[j], true, K, m, lane);
                    if (lane == src) {
                      rechecked_l = true;
                      ok_l = recheck_ok_l = ok;
                    }
                    if (!ok) break;  // the prefix ends here anyway
                  }
                }
              }
              unsigned frag = __ballot_sync(SMG_FULL, live_l && ((code_l != EVT_NONE && !plain_l) || !ok_l));
              if (A.spec == 2) frag = live;  // self-check: every row goes through the exact evaluation below
              const int f = frag ? __ffs(frag) - 1 : 32;
              const unsigned pm = live & (f >= 32 ? 0xffffffffu : ((1u << f) - 1u));
              const unsigned evp = __ballot_sync(SMG_FULL, pl) & pm;
              int L = f;  // the row handled alone below (if any)
              WALK_TICK(1);
              if (pm) {
                bool commit = true;
                if (evp) {
                  const bool mine = (evp >> lane) & 1u;
                  // the screen of the block must survive the whole prefix
                  const double maxdm1 = fmax(S.maxdm, key_to_double(warp_max_key(sort_key(mine ? dmb : 0.0))));
                  const double dplus1 = fmax(S.Dplus, key_to_double(warp_max_key(sort_key(mine ? dpb : 0.0))));
                  const double used = warp_sum(mine ? usedb : 0.0) + (dplus1 - S.Dplus);
                  if (scr_all)
                    commit = true;
                  else if (scr_kind == 2)
                    commit = S.scr_used + used <= SCAN_RESCREEN_SLACK;
                  else
                    commit = scr_kind == 1 && S.next == K0 && maxdm1 + dplus1 <= SCAN_FAST_DRIFT;
                  if (commit) {
                    if (mine) {
                      atomicAdd(&S.cnt[ns_l], 1);
                      atomicSub(&S.cnt[own_l], 1);
                      A.c[i0 + row_l] = ns_l;
                    }
                    __syncwarp();
                    double dtrue = 0.0;
                    if (mine) {
#pragma unroll
                      for (int h = 0; h < 2; h++) {  // (several lanes may write the same values for a cluster)
                        const int sl = h ? own_l : ns_l, e = h ? oe_l : code_l;
                        const int c = S.cnt[sl];
                        const double lc = log((double)c), lcm1 = c > 1 ? log((double)(c - 1)) : -CUDART_INF;
                        S.logc[sl] = lc;
                        S.logcm1[sl] = lcm1;
                        if (sl < K0) S.dminus[sl] = c > 1 ? S.lcm1_0[sl] - lcm1 : CUDART_INF;
                        S.fac[e] = drift_factor(lc, S.base_lc[e]);
                        S.facm1[e] = drift_factor(lcm1, S.base_lcm1[e]);
                        dtrue = fmax(dtrue, fmax(dd(lc, S.base_lc[e]), dd(lcm1, S.base_lcm1[e])));
                      }
                    }
                    D = fmax(D, key_to_double(warp_max_key(sort_key(dtrue))));
                    if (lane == 0) {
                      S.maxdm = maxdm1;
                      S.Dplus = dplus1;
                      S.scr_used += used;
                      S.stats[1] += __popc(evp);
                    }
                    __syncwarp();
                    next_start = __shfl_sync(SMG_FULL, row_l, 31 - __clz(evp)) + 1;
                  }
                }
                if (commit) {
                  live &= ~pm;
                  done = g * 32 + (32 - __clz(pm));
                  if (evp && done < nb && !(D <= A.spec_dmax)) reason = 2;
                  WALK_COUNT(5);
                  WALK_TICK(2);
                  continue;
                }
                L = __ffs(evp) - 1;  // the screen's allowance does not cover the prefix: its first move alone
              }
              // ---- (b) one row alone: exact evaluation when its tolerance is used up, anything but a plain move
              // handed back to the block
              int code = __shfl_sync(SMG_FULL, code_l, L);
              const double tau = shfl_d(tau_l, L);
              bool robust = tau > D;
              if (!robust) {
                const int j = g * 32 + L;
                robust = scan_recheck_row(S, spec_w + j * SCAN_SPEC_WSTRIDE, (int)spec_sel[j], (int)spec_own[j], spec_u[j], false, K, m, lane);
              }
              const int r = __shfl_sync(SMG_FULL, row_l, L);
              const int own = __shfl_sync(SMG_FULL, own_l, L);
              const int ie = i0 + r;
              if (!robust && A.spec != 2 && g * 32 + L > 0) {
                // Its tolerance is used up.  One warp evaluating one row is ~2000 dependent instructions; the whole
                // block evaluating the next rows again from here costs about the same and covers 32 R of them.
                reason = 2;
                done = g * 32 + L;
                nfrag++;
                break;
              }
              if (!robust || A.spec == 2) {
                bounds_valid = false;
                const int exact = scan_eval_row<2>(A, S, ie, own, K, lane);
                if (!robust)
                  nfrag++;
                else if (exact != code && lane == 0)
                  S.stats[4]++;  // a speculated outcome that the state did not confirm: must never happen
                code = exact;
              }
              WALK_COUNT(6);
              live &= ~((2u << L) - 1u);
              if (code == EVT_NONE) continue;
              if (code < 0 || code >= K || S.cnt[own] == 1) {  // left to the block: the row is evaluated again there
                reason = 3;
                done = g * 32 + L;
                next_start = r;
                break;
              }
              if (lane == 0) {
                const int ns = S.l2s[code];
                const double dp0 = S.Dplus, dm0 = own < K0 ? S.dminus[own] : 0.0, lcn0 = S.logc[ns];
                const int oe = S.s2l[own];
                scan_apply_move(A, S, K0, ie, own, code);
                double used = S.Dplus - dp0;
                if (own < K0) used += S.dminus[own] - dm0;
                if (ns >= K0) used += S.logc[ns] - lcn0;
                S.scr_used += used;
                bool keep;
                if (scr_all)
                  keep = true;
                else if (scr_kind == 2)
                  keep = S.scr_used <= SCAN_RESCREEN_SLACK;
                else
                  keep = scr_kind == 1 && S.next == K0 && S.maxdm + S.Dplus <= SCAN_FAST_DRIFT;
                S.spec_keep = keep;
                const double dn = fmax(dd(S.logc[ns], S.base_lc[code]), dd(S.logcm1[ns], S.base_lcm1[code]));
                const double dow = fmax(dd(S.logc[own], S.base_lc[oe]), dd(S.logcm1[own], S.base_lcm1[oe]));
                S.fac[code] = drift_factor(S.logc[ns], S.base_lc[code]);
                S.facm1[code] = drift_factor(S.logcm1[ns], S.base_lcm1[code]);
                S.fac[oe] = drift_factor(S.logc[own], S.base_lc[oe]);
                S.facm1[oe] = drift_factor(S.logcm1[own], S.base_lcm1[oe]);
                double Dn = fmax(D, fmax(dn, dow));
                if (!(Dn == Dn)) Dn = CUDART_INF;
                S.specD = Dn;
              }
              __syncwarp();
              D = S.specD;
              WALK_TICK(3);
              done = g * 32 + L + 1;
              next_start = r + 1;
              if (!S.spec_keep) {
                reason = 1;
                break;
              }
              if (done < nb && (!(D <= A.spec_dmax) || nfrag > 8 + (done >> 4))) {
                reason = 2;
                break;
              }
            }
          }
          if (lane == 0) {
            S.walk_done = done;
            S.walk_start = next_start;
            S.walk_reason = reason;
            S.stats[0]++;
            S.stats[5] += nfrag;
            if (reason == 2) S.stats[6]++;
            S.wk_cmd = 1;
          }
          scan_walk_barrier();
        } else {
          for (;;) {
            scan_walk_barrier();
            if (S.wk_cmd) break;
            const unsigned need = S.wk_need;
            if (warp < __popc(need)) {
              const int src = __fns(need, 0, warp + 1);  // the (warp+1)-th row of the job
              const int j = S.wk_j0 + src;
              if (scan_recheck_row(S, spec_w + j * SCAN_SPEC_WSTRIDE, (int)spec_sel[j], (int)spec_own[j], spec_u[j], true, K, m, lane) && lane == 0)
                atomicOr(&S.wk_ok, 1u << src);
            }
            scan_walk_barrier();
          }
        }
        __syncthreads();
        {
          const int done = S.walk_done, reason = S.walk_reason;
          hot = (S.walk_start >= 0 || reason) ? 4 : hot - 1;
          if (S.walk_start >= 0) start = S.walk_start;
          consumed += done;
          if (reason == 0) {
            R = max(1, min(2 * R, min(A.spec_rmax, SCAN_SPEC_RMAX)));
            W = SMG_SCAN_WARPS;
          } else if (reason == 1) {
            screened = false;  // the remaining rows are screened again
          } else if (reason == 2) {
            R = max(1, min(R, (done + SMG_SCAN_WARPS - 1) / SMG_SCAN_WARPS));  // about as far as this one got
          } else {
            one_legacy = true;
            W = 4;
          }
        }
        ring_ready = false;  // (the cached weights lie over the staging buffers)
        SCAN_TICK(5);
        continue;
      }
      one_legacy = false;
      // ================= batch: the next `nb` undecided rows, one warp each =================
      // (only the warps that evaluate a row run the selection; the others go straight to the barrier)
      const int nb = min(W, total_und - consumed);
      int myrow = -1;
      if (warp < nb) {
        // lane l owns words SCAN_SUPER*l .. SCAN_SUPER*l + SCAN_SUPER-1 of the undecided bit map
        unsigned wb[SCAN_SUPER];
        int wpc = 0;
#pragma unroll
        for (int k = 0; k < SCAN_SUPER; k++) {
          wb[k] = S.und[SCAN_SUPER * lane + k];
          wpc += __popc(wb[k]);
        }
        int incl = wpc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int y = __shfl_up_sync(SMG_FULL, incl, o);
          if (lane >= o) incl += y;
        }
        const int wpref = incl - wpc;
        const int t = consumed + warp;
        // the lane whose words hold the t-th undecided row locates it; everybody reads the answer from that lane
        int found = -1;
        if (t >= wpref && t < wpref + wpc) {
          int off = t - wpref;
#pragma unroll
          for (int k = 0; k < SCAN_SUPER; k++) {
            const int c = __popc(wb[k]);
            if (found < 0 && off < c) {
              unsigned bits = wb[k];
              for (int q = 0; q < off; q++) bits &= bits - 1;  // drop the `off` lowest set bits
              found = (SCAN_SUPER * lane + k) * 32 + __ffs(bits) - 1;
            }
            off -= c;
          }
        }
        const unsigned hit = __ballot_sync(SMG_FULL, found >= 0);
        myrow = __shfl_sync(SMG_FULL, found, __ffs(hit) - 1);
      }
      SCAN_TICK(2);
      const int K = S.K;
      const int ne = K + m;
      int code = EVT_NONE;
      if (myrow >= 0 && ne > SMG_MAX_ENTRIES) code = -3;
      if (myrow >= 0 && ne <= SMG_MAX_ENTRIES) {
        const int i = i0 + myrow;
        const int old_slot = __ldcg(&A.c[i]);  // rows the scan has not reached still hold their start-of-pass slot
        code = (ne <= 64) ? scan_eval_row<2>(A, S, i, old_slot, K, lane) : scan_eval_row_wide(A, S, i, old_slot, K, lane);
      }
      if (lane == 0) {
        S.evt[warp] = code;
        S.row[warp] = myrow;
      }
      SCAN_TICK(3);
      __syncthreads();
      // ---- first event of the batch (warps hold ascending rows)
      int ev = S.evt[lane];
      unsigned evm = __ballot_sync(SMG_FULL, ev != EVT_NONE && S.row[lane] >= 0);
      if (evm == 0) {
        if (tid == 0) S.stats[0]++;
        consumed += nb;
        dense_run = 0;
        W = SMG_SCAN_WARPS;  // quiet: speculate over a full batch again
        __syncthreads();     // S.evt / S.row are rewritten by the next round
        SCAN_TICK(4);
        continue;
      }
      SCAN_TICK(4);
      const int first = __ffs(evm) - 1;
      hot = 4;
      // event-dense stretches (burn-in): evaluating 32 rows per round only burns issue slots
      W = min(SMG_SCAN_WARPS, max(4, 2 * (first + 1)));
      dense_run = (first == 0) ? dense_run + 1 : 0;
      const int new_e = __shfl_sync(SMG_FULL, ev, first);
      const int erow = S.row[first];
      const int ie = i0 + erow;
      if (new_e < 0) {  // error: stop the pass
        if (tid == 0) atomicOr(A.status, new_e == -2 ? ST_BAD_PROB : ST_TOO_MANY_ENTRIES);
        abort_pass = true;
        break;
      }
      const int old_slot = A.c[ie];
      const bool singleton = (S.cnt[old_slot] == 1);
      const int Kc = S.K;
      int new_slot;
      if (new_e < Kc) {
        new_slot = S.l2s[new_e];
      } else {
        new_slot = S.next;  // birth (case 3) or parameter replacement (case 4)
        if (new_slot >= A.NS || new_slot >= SMG_MAX_SLOTS) {
          if (tid == 0) atomicOr(A.status, ST_SLOTS_EXHAUSTED);
          abort_pass = true;
          break;
        }
        // copy the auxiliary component's parameters into the new slot
        const int a = new_e - Kc;
        const long long e = A.aux_e[(size_t)ie * m + a];
        if (A.aux_free) {
          for (int j = tid; j < pp; j += blockDim.x) {
            int center = 0;
            double sigma = 1.0;
            if (j < A.p) prior_entry_attr(A.aux_key, e, j, A.attr[j], A.hv[j], A.hw[j], A.sigma_exact, &center, &sigma);
            A.cen[(size_t)new_slot * pp + j] = (uint8_t)center;
            A.sig[(size_t)new_slot * pp + j] = sigma;
            A.isg[(size_t)new_slot * pp + j] = j < A.p ? 1.0 / sigma : 0.0;
          }
          if (tid == 0) A.sden[new_slot] = A.aux_sd[(size_t)ie * m + a];
        } else {
          for (int j = tid; j < pp; j += blockDim.x) {
            A.cen[(size_t)new_slot * pp + j] = A.pool_cen[(size_t)e * pp + j];
            A.sig[(size_t)new_slot * pp + j] = A.pool_sig[(size_t)e * pp + j];
            A.isg[(size_t)new_slot * pp + j] = A.pool_isg[(size_t)e * pp + j];
          }
          if (tid == 0) A.sden[new_slot] = A.pool_sden[e];
        }
        if (new_slot < A.ldl && ie + 1 < n) {
          // materialise the new column for every later observation, on the whole cluster
          if (tid == 0) {
            A.job[0] = new_slot;
            A.job[1] = ie + 1;
            A.job[2] = n;
          }
          cluster_sync_all();  // mailbox + parameter vector visible to every CTA of the cluster
          scan_fill_column(A, new_slot, ie + 1, n, warp, lane, 0, SCAN_CLUSTER);
          cluster_sync_all();  // column complete (read below with ld.global.cg)
        }
      }
      __syncthreads();  // everyone has read the pre-event state
      if (tid == 0) {
        S.stats[0]++;
        if (new_e < Kc) {
          // drift this move adds to what the rows still to come are screened against: the global gain bound, the loss
          // term of the cluster that shrinks, the count of a cluster born in this pass that grows
          const int ns = S.l2s[new_e];
          const double dp0 = S.Dplus, dm0 = old_slot < K0 ? S.dminus[old_slot] : 0.0, lcn0 = S.logc[ns];
          scan_apply_move(A, S, K0, ie, old_slot, new_e);
          double used = S.Dplus - dp0;
          if (old_slot < K0) used += S.dminus[old_slot] - dm0;  // (inf or NaN once the cluster is down to one member: re-screen)
          if (ns >= K0) used += S.logc[ns] - lcn0;
          S.scr_used += used;
        } else {
          S.scr_used = CUDART_INF;  // a new column: every remaining row has to be compared with it
          S.stats[1]++;
          A.c[ie] = new_slot;
          auto drift = [&](int s) {
            if (s >= K0) return;
            S.dminus[s] = (S.cnt[s] > 1) ? (S.lcm1_0[s] - S.logcm1[s]) : CUDART_INF;
            if (S.lcm1_0[s] > -CUDART_INF && S.dminus[s] > S.maxdm) S.maxdm = S.dminus[s];
            const double dp = S.logc[s] - S.lc0[s];
            if (dp > S.Dplus) S.Dplus = dp;  // monotone upper bound
          };
          S.next = new_slot + 1;
          S.cnt[new_slot] = 1;
          S.logc[new_slot] = 0.0;
          S.logcm1[new_slot] = -CUDART_INF;
          if (!singleton) {  // case 3 (neal8.cpp:140-150)
            int c0 = --S.cnt[old_slot];
            S.logc[old_slot] = S.logcm1[old_slot];
            S.logcm1[old_slot] = c0 > 1 ? log((double)(c0 - 1)) : -CUDART_INF;
            drift(old_slot);
            S.l2s[Kc] = new_slot;
            S.s2l[new_slot] = Kc;
            S.K = Kc + 1;
            S.stats[2]++;
          } else {  // case 4 (neal8.cpp:153-159): same label, new parameters
            int lab = S.s2l[old_slot];
            S.cnt[old_slot] = 0;
            S.logc[old_slot] = S.logcm1[old_slot] = -CUDART_INF;
            S.s2l[old_slot] = -1;
            S.l2s[lab] = new_slot;
            S.s2l[new_slot] = lab;
          }
        }
      }
      start = erow + 1;
      __syncthreads();
      {
        // The state changed.  The bit map of the block stays valid when the event was a plain move (no cluster born or
        // closed) and the drift allowance it was made with is not used up; the rows of this batch before the event were
        // non-events and are final, the rows after it are evaluated again against the new state.
        bool keep = new_e < Kc && !singleton;
        if (keep) {
          if (scr_kind == 2)
            keep = S.scr_used <= SCAN_RESCREEN_SLACK;
          else
            keep = scr_kind == 1 && S.next == K0 && S.maxdm + S.Dplus <= SCAN_FAST_DRIFT;
        }
        if (keep)
          consumed += first + 1;
        else
          screened = false;  // the remaining rows are screened again
      }
      SCAN_TICK(5);
    }
    __syncthreads();  // everybody is done with this block's staging buffers and bit map
  }
#ifdef SMG_SCAN_PROFILE
  if (tid == 0 && A.prof) {
    pc[6] = clock64() - tstart;
    for (int q = 0; q < 8; q++) A.prof[q] += (unsigned long long)pc[q];
    for (int q = 0; q < 8; q++) A.prof[8 + q] += (unsigned long long)wkc[q];
  }
#endif
  for (int b = tid; b * SCAN_BLOCK < n; b += blockDim.x) A.und_blk[b] = 0;  // for the next pass
  // release the helpers
  if (tid == 0) A.job[0] = -1;
  cluster_sync_all();
  // publish: K, counts by slot, slot->label map
  for (int s = tid; s < A.NS && s < SMG_MAX_SLOTS; s += blockDim.x) {
    A.counts[s] = S.cnt[s];
    A.slot2label[s] = S.s2l[s];
  }
  if (tid == 0) {
    *A.Kptr = S.K;
    if (S.K > A.K0cap) atomicOr(A.status, ST_LL_COLS);
    if (A.stats)
      for (int q = 0; q < 7; q++) A.stats[q] += S.stats[q];  // ([7] belongs to the split-merge kernels)
  }
}

// After the pass: c[i] <- label of its slot; parameters gathered into label order (dst buffers).
__global__ void scan_finalize_labels_kernel(int* __restrict__ c, int n, const int* __restrict__ slot2label) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) c[i] = slot2label[c[i]];
}
__global__ void scan_finalize_params_kernel(const int* __restrict__ slot2label, int NS, int pp,
                                            const uint8_t* __restrict__ cen_s, const double* __restrict__ sig_s,
                                            const double* __restrict__ isg_s, const double* __restrict__ sden_s,
                                            const int* __restrict__ cnt_s, uint8_t* __restrict__ cen_d,
                                            double* __restrict__ sig_d, double* __restrict__ isg_d,
                                            double* __restrict__ sden_d, int* __restrict__ cnt_d) {
  int s = blockIdx.x;
  if (s >= NS) return;
  int lab = slot2label[s];
  if (lab < 0) return;
  for (int j = threadIdx.x; j < pp; j += blockDim.x) {
    cen_d[(size_t)lab * pp + j] = cen_s[(size_t)s * pp + j];
    sig_d[(size_t)lab * pp + j] = sig_s[(size_t)s * pp + j];
    isg_d[(size_t)lab * pp + j] = isg_s[(size_t)s * pp + j];
  }
  if (threadIdx.x == 0) {
    sden_d[lab] = sden_s[s];
    cnt_d[lab] = cnt_s[s];
  }
}

// =============================================================================
// K3: category histogram H[k][j][a] = #{i : c_i = k, x_ij = a+1} and member counts.
// One thread per (observation, 16-attribute chunk); integer atomics into L2.
// =============================================================================
__global__ void __launch_bounds__(256) cluster_histogram_kernel(const uint8_t* __restrict__ X, int n, int pp,
                                                                const int* __restrict__ c, int mmax,
                                                                int* __restrict__ H, int* __restrict__ counts) {
  const int chunks = pp / 16;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * chunks) return;
  const int i = (int)(t / chunks), q = (int)(t % chunks);
  const int k = c[i];
  if (q == 0) atomicAdd(&counts[k], 1);
  uint4 v = *reinterpret_cast<const uint4*>(X + (size_t)i * pp + q * 16);
  uint32_t w[4] = {v.x, v.y, v.z, v.w};
  int* Hk = H + ((size_t)k * pp + q * 16) * mmax;
#pragma unroll
  for (int b = 0; b < 16; b++) {
    int x = (w[b >> 2] >> ((b & 3) * 8)) & 0xff;
    if (x) atomicAdd(&Hk[b * mmax + (x - 1)], 1);
  }
}

// Same result with the histogram privatised in shared memory: a CTA owns one 16-attribute slice of
// the table ([K][16][mmax] ints) and a strided share of the rows, so the integer atomics stay on chip
// and each (cluster, attribute, level) cell costs one global atomic per CTA instead of one per row.
__global__ void __launch_bounds__(256) cluster_histogram_smem_kernel(const uint8_t* __restrict__ X, int n, int pp,
                                                                     const int* __restrict__ c, int mmax,
                                                                     const int* __restrict__ Kptr, int Kcap,
                                                                     int* __restrict__ H, int* __restrict__ counts) {
  extern __shared__ int s_h[];  // [K][16][mmax] then [Kcap] member counts
  const int K = min(*Kptr, Kcap);
  const int cells = K * 16 * mmax;
  int* s_cnt = s_h + (size_t)Kcap * 16 * mmax;
  const int slice = blockIdx.x;
  for (int q = threadIdx.x; q < cells; q += blockDim.x) s_h[q] = 0;
  for (int q = threadIdx.x; q < Kcap; q += blockDim.x) s_cnt[q] = 0;
  __syncthreads();
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n; i += gridDim.y * blockDim.x) {
    const int k = c[i];
    if (k >= K) continue;
    const uint4 v = *reinterpret_cast<const uint4*>(X + (size_t)i * pp + slice * 16);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    if (slice == 0) atomicAdd(&s_cnt[k], 1);
    int* Hk = s_h + (size_t)k * 16 * mmax;
#pragma unroll
    for (int b = 0; b < 16; b++) {
      const int x = (w[b >> 2] >> ((b & 3) * 8)) & 0xff;
      if (x) atomicAdd(&Hk[b * mmax + (x - 1)], 1);
    }
  }
  __syncthreads();
  const int per = 16 * mmax;
  for (int q = threadIdx.x; q < cells; q += blockDim.x) {
    const int hv = s_h[q];
    if (hv) {
      const int k = q / per, rem = q - k * per;
      atomicAdd(&H[((size_t)k * pp + slice * 16) * mmax + rem], hv);
    }
  }
  if (slice == 0)
    for (int q = threadIdx.x; q < K; q += blockDim.x)
      if (s_cnt[q]) atomicAdd(&counts[q], s_cnt[q]);
}

// Incremental form: H already holds the histogram of the labels c_hist; only the rows whose label changed since
// then are moved (subtracted from the old cluster, added to the new one), which at stationarity is a handful of
// rows per sweep instead of all n.  Integer arithmetic, so the table equals a full rebuild.  One warp per 32
// rows; a changed row is moved by the whole warp (8 attributes per lane and step).  Member counts are rebuilt
// from c (privatised in shared memory) exactly as the full kernels do.
__global__ void __launch_bounds__(256) cluster_histogram_update_kernel(const uint8_t* __restrict__ X, int n, int pp,
                                                                       const int* __restrict__ c, int* __restrict__ c_hist,
                                                                       int mmax, int Kcap, int* __restrict__ H,
                                                                       int* __restrict__ counts) {
  extern __shared__ int s_cnt[];  // [Kcap]
  for (int q = threadIdx.x; q < Kcap; q += blockDim.x) s_cnt[q] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int base = gw * 32; base < n; base += nw * 32) {
    const int i = base + lane;
    const int k = i < n ? c[i] : -1, kp = i < n ? c_hist[i] : -1;
    if (k >= 0 && k < Kcap) atomicAdd(&s_cnt[k], 1);
    unsigned chg = __ballot_sync(SMG_FULL, k != kp);
    if (k != kp) c_hist[i] = k;
    while (chg) {
      const int b = __ffs(chg) - 1;
      chg &= chg - 1;
      const int kn = __shfl_sync(SMG_FULL, k, b), ko = __shfl_sync(SMG_FULL, kp, b);
      const uint8_t* x = X + (size_t)(base + b) * pp;
      for (int off = lane * 8; off < pp; off += 256) {
        const uint2 v = *reinterpret_cast<const uint2*>(x + off);
        const uint32_t w[2] = {v.x, v.y};
#pragma unroll
        for (int q = 0; q < 8; q++) {
          const int lv = (w[q >> 2] >> ((q & 3) * 8)) & 0xff;
          if (lv) {
            if (ko >= 0 && ko < Kcap) atomicSub(&H[((size_t)ko * pp + off + q) * mmax + (lv - 1)], 1);
            if (kn >= 0 && kn < Kcap) atomicAdd(&H[((size_t)kn * pp + off + q) * mmax + (lv - 1)], 1);
          }
        }
      }
    }
  }
  __syncthreads();
  for (int q = threadIdx.x; q < Kcap; q += blockDim.x)
    if (s_cnt[q]) atomicAdd(&counts[q], s_cnt[q]);
}

// histogram of a member subset split in two groups by z (split-merge launch states):
// rows = S[0..nS) with group z[s], plus the two anchors i1 (group 0) and i2 (group 1).
__global__ void __launch_bounds__(256) subset_histogram_kernel(const uint8_t* __restrict__ X, int pp,
                                                               const int* __restrict__ S,
                                                               const int* __restrict__ nSptr,
                                                               const int* __restrict__ z, const int* __restrict__ anchors,
                                                               int mmax, int* __restrict__ H2, int* __restrict__ cnt2,
                                                               const int* enable, int enable_val) {
  if (enable && *enable != enable_val) return;
  const int chunks = pp / 16;
  const int nS = *nSptr;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)(nS + 2) * chunks) return;
  const int r = (int)(t / chunks), q = (int)(t % chunks);
  int row, g;
  if (r < nS) {
    row = S[r];
    g = z ? z[r] : 0;
  } else {
    row = anchors[r - nS];
    g = z ? (r - nS) : 0;
  }
  if (q == 0) atomicAdd(&cnt2[g], 1);
  uint4 v = *reinterpret_cast<const uint4*>(X + (size_t)row * pp + q * 16);
  uint32_t w[4] = {v.x, v.y, v.z, v.w};
  int* Hk = H2 + ((size_t)g * pp + q * 16) * mmax;
#pragma unroll
  for (int b = 0; b < 16; b++) {
    int x = (w[b >> 2] >> ((b & 3) * 8)) & 0xff;
    if (x) atomicAdd(&Hk[b * mmax + (x - 1)], 1);
  }
}

// Same, with the two histograms privatised in shared memory (2*pp*mmax ints) and a fixed grid that strides
// over the member rows: one flush of integer atomics per CTA instead of one atomic per (row, attribute).
// Counts are not touched when cnt2 == null (the restricted-scan decision kernel already knows them).
__device__ __forceinline__ void subset_hist_body(const uint8_t* __restrict__ X, int pp, const int* __restrict__ S, int nS,
                                                 const int* __restrict__ z, const int* __restrict__ anchors, int mmax,
                                                 int* __restrict__ H2, int* __restrict__ cnt2, int* s_h, int block,
                                                 int nblocks) {
  const int len = pp * mmax;
  const int ng = z ? 2 : 1;
  for (int q = threadIdx.x; q < ng * len; q += blockDim.x) s_h[q] = 0;
  __syncthreads();
  const int chunks = pp / 16;
  const long long total = (long long)(nS + 2) * chunks;
  int c0 = 0, c1 = 0;
  for (long long t = (long long)block * blockDim.x + threadIdx.x; t < total; t += (long long)nblocks * blockDim.x) {
    const int r = (int)(t / chunks), q = (int)(t % chunks);
    int row, g;
    if (r < nS) {
      row = S[r];
      g = z ? z[r] : 0;
    } else {
      row = anchors[r - nS];
      g = z ? (r - nS) : 0;
    }
    if (q == 0) {
      c0 += (g == 0);
      c1 += (g == 1);
    }
    const uint4 v = *reinterpret_cast<const uint4*>(X + (size_t)row * pp + q * 16);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    int* Hk = s_h + (size_t)g * len + (size_t)q * 16 * mmax;
#pragma unroll
    for (int b = 0; b < 16; b++) {
      const int x = (w[b >> 2] >> ((b & 3) * 8)) & 0xff;
      if (x) atomicAdd(&Hk[b * mmax + (x - 1)], 1);
    }
  }
  if (cnt2) {
    if (c0) atomicAdd(&cnt2[0], c0);
    if (c1) atomicAdd(&cnt2[1], c1);
  }
  __syncthreads();
  for (int q = threadIdx.x; q < ng * len; q += blockDim.x) {
    const int hv = s_h[q];
    if (hv) atomicAdd(&H2[q], hv);
  }
}

__global__ void __launch_bounds__(256) subset_histogram_smem_kernel(const uint8_t* __restrict__ X, int pp,
                                                                    const int* __restrict__ S,
                                                                    const int* __restrict__ nSptr,
                                                                    const int* __restrict__ z,
                                                                    const int* __restrict__ anchors, int mmax,
                                                                    int* __restrict__ H2, int* __restrict__ cnt2,
                                                                    const int* enable, int enable_val) {
  if (enable && *enable != enable_val) return;
  extern __shared__ int s_h[];  // [2][pp][mmax]
  subset_hist_body(X, pp, S, *nSptr, z, anchors, mmax, H2, cnt2, s_h, blockIdx.x, gridDim.x);
}

// =============================================================================
// K4: per-(cluster, attribute) centre and sigma draws (update_phi).
// A "job" updates one cluster: histogram row `hist`, current sigma from slot `src`,
// result written to slot `dst`.  jobs == nullptr => job k updates cluster k in place.
//   centre ~ Rcpp::sample(1..m_j, probs = softmax(-(n_k - freq)/sigma_j))   (:495-505, :195)
//   s = freq[centre];  sigma ~ HIG(v_j + s, w_j + n_k - s, m_j)             (:582-589)
// Prior draws (sample_center_1_cluster / sample_sigma_1_cluster without data,
// common_functions.cpp:199,232) use count = 0 with `prior = 1`.
// =============================================================================
struct PhiJob {
  int hist, src, dst, cnt_idx;
  uint32_t sub;      // Philox sub-phase of this job's draws
  int prior;         // 1 => draw from the prior (no data)
  int enable_mode;   // 0: always; 1: only when *enable != 0; 2: only when *enable == 0
  const double* uc;  // injected uniforms [p] of this job (centre / sigma) or null
  const double* us;
};

#define PHI_MAX_INLINE_JOBS 4

// A job may be split over several CTAs ("parts" of consecutive attributes, a multiple of 32 each): the draws are a
// long double-precision dependency chain per attribute, so spreading them over more SMs shortens the job.
__host__ __device__ inline int phi_part_chunk(int pp, int nparts) { return (((pp + nparts - 1) / nparts) + 31) & ~31; }
__host__ __device__ inline int phi_parts_for(int pp, int want) {
  const int chunk = phi_part_chunk(pp, want < 1 ? 1 : want);
  return (pp + chunk - 1) / chunk;
}

struct PhiArgs {
  int pp, p, mmax;
  int nparts;     // CTAs per job
  double* den;    // [slots][pp] per-attribute log-normalisers (scratch of the split jobs)
  int* part_cnt;  // [slots] finished parts of the job writing that slot (self-resetting)
  const int* attr;
  const double* v;
  const double* w;
  const int* H;
  const int* counts;
  // mode A (njobs > 0): the jobs listed here.  mode B (njobs == 0): job k updates cluster k in place for
  // k < *njobs_ptr, with u_center / u_sigma [job][p] (stride u_stride) and the fields `sub`, `prior` below.
  PhiJob jobs[PHI_MAX_INLINE_JOBS];
  int njobs;
  const int* njobs_ptr;
  const uint8_t* cen_src;
  const double* sig_src;
  uint8_t* cen;
  double* sig;
  double* isg;
  double* sden;
  const double* u_center;
  const double* u_sigma;
  int u_stride;
  RngKey key;
  int prior;
  int sigma_exact;  // 1 => always the one-uniform inverse-CDF sigma draw (also used whenever sigma uniforms are injected)
  const int* enable;  // device flag consulted by jobs with enable_mode != 0
  int* status;
  unsigned long long* prof;  // optional cycle counters (SMG_PHI_PROFILE): [0] centre part, [1] sigma part, [2] tail, [3] calls
};

// Centre draw of one attribute (compute_prob_centers + sample_center_1_cluster, common_functions.cpp:495-505,195):
// probs = softmax(-(n_k - freq_a)/sigma), then Rcpp::sample(1..m, 1, true, probs).  Up to 8 levels everything stays in
// registers, and when the largest probability is unique and already covers u -- the first step of Rcpp's
// descending-order walk -- the sort is skipped; ties and the other draws take the general path, which
// reproduces R's revsort permutation.  Returns the 1-based level.
#define SMG_CENTER_DOM 44.0
__device__ __forceinline__ int draw_center(const int* __restrict__ h, int nk, double sg, int m, double u) {
  {  // dominance screen (see draw_center_grp)
    int h1 = -1, h2 = -1, arg = 0, ntop = 0;
    for (int a = 0; a < m; a++) {
      const int v = h[a];
      if (v > h1) {
        h2 = h1;
        h1 = v;
        arg = a;
        ntop = 1;
      } else if (v == h1) {
        ntop++;
      } else if (v > h2) {
        h2 = v;
      }
    }
    if (ntop == 1 && (double)(h1 - max(h2, 0)) >= SMG_CENTER_DOM * sg && sg > 0.0) return arg + 1;
  }
  if (m <= 8) {
    double p[8];
    double mx = -CUDART_INF;
#pragma unroll
    for (int a = 0; a < 8; a++) {
      p[a] = (a < m) ? -((double)nk - (double)h[a]) / sg : -CUDART_INF;
      mx = p[a] > mx ? p[a] : mx;
    }
    double sum = 0.0;
#pragma unroll
    for (int a = 0; a < 8; a++)
      if (a < m) {
        p[a] = exp(p[a] - mx);
        sum += p[a];
      }
    double sum2 = 0.0;
#pragma unroll
    for (int a = 0; a < 8; a++)
      if (a < m) {
        p[a] = p[a] / sum;
        sum2 += p[a];
      }
    double qmax = -1.0;
    int arg = 0, nmax = 0;
#pragma unroll
    for (int a = 0; a < 8; a++)
      if (a < m) {
        const double q = p[a] / sum2;  // what Rcpp::sample compares against after its own normalisation
        if (q > qmax) {
          qmax = q;
          arg = a;
          nmax = 1;
        } else if (q == qmax) {
          nmax++;
        }
      }
    if (nmax == 1 && u <= qmax) return arg + 1;
    double pt[8];
#pragma unroll
    for (int a = 0; a < 8; a++) pt[a] = (a < m) ? p[a] : 0.0;
    return 1 + sample_probs_small(pt, m, u);
  }
  double pt[SMG_MAX_LEVELS];
  double mx = -CUDART_INF;
  for (int a = 0; a < m; a++) {
    pt[a] = -((double)nk - (double)h[a]) / sg;
    mx = pt[a] > mx ? pt[a] : mx;
  }
  double sum = 0.0;
  for (int a = 0; a < m; a++) {
    pt[a] = exp(pt[a] - mx);
    sum += pt[a];
  }
  for (int a = 0; a < m; a++) pt[a] = pt[a] / sum;
  return 1 + sample_probs_small(pt, m, u);
}

// The m <= 8 case by a group of PHI_G lanes, lane g holding level g+1 (hg = its count): the divisions and
// exponentials of the levels run side by side, the two normalising sums are accumulated in level order on every
// lane (same roundings as the one-thread version above).  Returns the level on every lane; *s_match = its count.
__device__ __forceinline__ int draw_center_grp(int hg, int nk, double sg, int m, double u, int g, unsigned gmask,
                                               int gbase, double* s_match) {
  const bool valid = g < m;
  {
    // Dominance screen: when the most frequent level leads every other one by >= SMG_CENTER_DOM nats of
    // (count difference)/sigma, the sums below are exactly 1.0 in double precision ((m-1) e^-44 < 2^-54), the leader's
    // normalised probability is exactly 1.0 and Rcpp::sample returns it for every u: same result, no exp / division.
    const int hv = valid ? hg : -1;
    const int h1 = __reduce_max_sync(gmask, hv);
    const unsigned top = __ballot_sync(gmask, hv == h1) & gmask;
    const int h2 = __reduce_max_sync(gmask, hv == h1 ? -1 : hv);
    if (__popc(top) == 1 && (double)(h1 - max(h2, 0)) >= SMG_CENTER_DOM * sg && sg > 0.0) {
      *s_match = (double)h1;
      return (__ffs(top) - 1 - gbase) + 1;
    }
  }
  const double lp = valid ? -((double)nk - (double)hg) / sg : -CUDART_INF;
  double mx = lp;
#pragma unroll
  for (int o = 1; o < PHI_G; o <<= 1) {
    const double y = __shfl_xor_sync(gmask, mx, o);
    mx = y > mx ? y : mx;
  }
  const double e = valid ? exp(lp - mx) : 0.0;
  double sum = 0.0;
#pragma unroll
  for (int a = 0; a < PHI_G; a++) {
    const double y = __shfl_sync(gmask, e, gbase + a);
    if (a < m) sum += y;
  }
  const double pn = valid ? e / sum : 0.0;
  double sum2 = 0.0;
#pragma unroll
  for (int a = 0; a < PHI_G; a++) {
    const double y = __shfl_sync(gmask, pn, gbase + a);
    if (a < m) sum2 += y;
  }
  const double q = valid ? pn / sum2 : -1.0;  // what Rcpp::sample compares against after its own normalisation
  double qmax = q;
#pragma unroll
  for (int o = 1; o < PHI_G; o <<= 1) {
    const double y = __shfl_xor_sync(gmask, qmax, o);
    qmax = y > qmax ? y : qmax;
  }
  const unsigned top = __ballot_sync(gmask, valid && q == qmax) & gmask;
  int center;
  if (__popc(top) == 1 && u <= qmax) {
    center = (__ffs(top) - 1 - gbase) + 1;
  } else {  // ties and the other draws: R's descending-order walk on lane 0
    double pt[PHI_G];
#pragma unroll
    for (int a = 0; a < PHI_G; a++) pt[a] = __shfl_sync(gmask, pn, gbase + a);
    center = 0;
    if (g == 0) center = 1 + sample_probs_small(pt, m, u);
    center = __shfl_sync(gmask, center, gbase);
  }
  *s_match = (double)__shfl_sync(gmask, hg, gbase + center - 1);
  return center;
}

// One CTA per job; thread j draws attribute j (and j+256, ...), then the CTA sums the per-attribute
// log-normalisers in a fixed order into sden[dst].
// One parameter-update job executed by one CTA (any block size >= 256: the first 256 threads draw, everybody
// takes part in the barriers of the final reduction).  `sh` = 256 doubles of shared memory.
__device__ __forceinline__ void phi_job_body(const PhiArgs& A, const PhiJob& J, int job, int part, int nparts, double* sh) {
  RngKey key = A.key;
  key.sub = J.sub;
  const int nk = J.prior ? 0 : A.counts[J.cnt_idx];
  if (!J.prior && nk == 0) return;  // empty cluster: untouched (common_functions.cpp:547)
  const int chunk = phi_part_chunk(A.pp, nparts);
  const int j1 = min(A.pp, (part + 1) * chunk);
#ifdef SMG_PHI_PROFILE
  long long tp0 = clock64(), tp1 = tp0, tp2 = tp0;
#endif
  // PHI_G lanes per attribute
  const int g = threadIdx.x & (PHI_G - 1), lane = threadIdx.x & 31, gbase = lane & ~(PHI_G - 1);
  const unsigned gmask = ((1u << PHI_G) - 1u) << gbase;
  for (int jb = part * chunk; jb < j1; jb += blockDim.x / PHI_G) {
    const int j = jb + (int)(threadIdx.x / PHI_G);
    if (j >= j1) continue;  // whole groups drop out together
    const size_t o = (size_t)J.dst * A.pp + j;
    if (j >= A.p) {  // padding attributes
      if (g == 0) {
        A.cen[o] = 0;
        A.sig[o] = 1.0;
        A.isg[o] = 0.0;
        A.den[o] = 0.0;
      }
      continue;
    }
    const int m = A.attr[j];
    int center;
    double s_match = 0.0;
    const double uc = get_u(J.uc, (size_t)j, key, U_CENTER, (uint32_t)job, (uint32_t)j);
    if (J.prior) {
      center = (int)((double)m * uc + 1.0);  // sample(m_j, 1): (int)(m*u + 1)
      if (center > m) center = m;
    } else {
      const double sg = A.sig_src[(size_t)J.src * A.pp + j];
      const int* h = A.H + ((size_t)J.hist * A.pp + j) * A.mmax;
      if (m <= PHI_G) {
        center = draw_center_grp(g < m ? h[g] : 0, nk, sg, m, uc, g, gmask, gbase, &s_match);
      } else {
        center = 0;
        if (g == 0) {
          center = draw_center(h, nk, sg, m, uc);
          s_match = (double)h[center - 1];
        }
        center = __shfl_sync(gmask, center, gbase);
        s_match = __shfl_sync(gmask, s_match, gbase);
      }
    }
    const double vv = A.v[j] + s_match;
    const double ww = A.w[j] + (double)nk - s_match;
    double uu = 0.5;
#ifdef SMG_PHI_PROFILE
    tp1 = clock64();
#endif
    if (J.us || A.sigma_exact) {
      if (g == 0) {
        const double us = get_u(J.us, (size_t)j, key, U_SIGMA, (uint32_t)job, (uint32_t)j);
        uu = hig_inv_u_d(us, vv, ww, (double)m);
      }
    } else {
      uu = hig_draw_u_grp(key, (uint32_t)job, (uint32_t)j, vv, ww, (double)m, g, gmask, gbase);
    }
#ifdef SMG_PHI_PROFILE
    tp2 = clock64();
#endif
    if (g == 0) {
      const double sigma = -1.0 / log(uu);
      A.cen[o] = (uint8_t)center;
      A.sig[o] = sigma;
      A.isg[o] = 1.0 / sigma;
      A.den[o] = hamming_den(sigma, m);
    }
  }
#if defined(SMG_PHI_PROFILE) && !defined(SMG_SM_HT_PROFILE)
  if (threadIdx.x == 0 && A.prof) {
    const long long tp3 = clock64();
    atomicAdd(&A.prof[0], (unsigned long long)(tp1 - tp0));
    atomicAdd(&A.prof[1], (unsigned long long)(tp2 - tp1));
    atomicAdd(&A.prof[2], (unsigned long long)(tp3 - tp2));
    atomicAdd(&A.prof[3], 1ull);
  }
#endif
  // the CTA that finishes the job last sums the normalisers of all attributes, always in the same order
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int done = nparts > 1 ? atomicAdd(&A.part_cnt[J.dst], 1) + 1 : 1;
    s_last = done == nparts;
    if (s_last && nparts > 1) A.part_cnt[J.dst] = 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x < 32) {  // thread t of the 256-thread block reduction summed j = t, t + 256, ...: same leaves, same tree
    const int lane = threadIdx.x;
    double v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      double acc = 0.0;
      for (int j = lane + 32 * k; j < A.p; j += 256) acc += __ldcg(&A.den[(size_t)J.dst * A.pp + j]);
      v[k] = acc;
    }
    const double t = tree256_warp(v);
    if (lane == 0) A.sden[J.dst] = t;
  }
}


__global__ void __launch_bounds__(256) phi_update_kernel(PhiArgs A) {
  // grid = jobs x A.nparts, parts of one job strided by the job count
  const int njg = gridDim.x / A.nparts, job = blockIdx.x % njg, part = blockIdx.x / njg;
  PhiJob J;
  if (A.njobs > 0) {
    if (job >= A.njobs) return;
    J = A.jobs[job];
    if (J.enable_mode == 1 && *A.enable == 0) return;
    if (J.enable_mode == 2 && *A.enable != 0) return;
  } else {
    if (job >= *A.njobs_ptr) return;
    J.hist = J.src = J.dst = J.cnt_idx = job;
    J.sub = A.key.sub;
    J.prior = A.prior;
    J.enable_mode = 0;
    J.uc = A.u_center ? A.u_center + (size_t)job * A.u_stride : nullptr;
    J.us = A.u_sigma ? A.u_sigma + (size_t)job * A.u_stride : nullptr;
  }
  __shared__ double sh[256];
  phi_job_body(A, J, job, part, A.nparts, sh);
}

// derive isg / den / sden from (cen, sig) for slots [0, nslots): used after host uploads
__global__ void __launch_bounds__(256) derive_terms_kernel(int nslots, int pp, int p, const int* __restrict__ attr,
                                                           const double* __restrict__ sig, double* __restrict__ isg,
                                                           double* __restrict__ den, double* __restrict__ sden) {
  const int s = blockIdx.x;
  if (s >= nslots) return;
  __shared__ double sh[256];
  double acc = 0.0;
  for (int j = threadIdx.x; j < pp; j += 256) {
    size_t o = (size_t)s * pp + j;
    double d = 0.0, w = 0.0;
    if (j < p) {
      double sg = sig[o];
      w = 1.0 / sg;
      d = hamming_den(sg, attr[j]);
    }
    isg[o] = w;
    if (den) den[o] = d;
    acc += d;
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) sden[s] = sh[0];
}

// =============================================================================
// full-data log-likelihood (common_functions.cpp:379-401): one warp per observation,
// per-CTA partials, then a fixed-order final reduction => run-to-run deterministic.
// =============================================================================
__global__ void __launch_bounds__(256) loglik_partial_kernel(const uint8_t* __restrict__ X, int n, int pp,
                                                             const int* __restrict__ c,
                                                             const uint8_t* __restrict__ cen,
                                                             const double* __restrict__ isg,
                                                             const double* __restrict__ sden,
                                                             double* __restrict__ partial) {
  __shared__ double sh[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double acc = 0.0;
  for (int i = blockIdx.x * 8 + warp; i < n; i += gridDim.x * 8) {
    int k = c[i];
    double dot = warp_mismatch_dot(X + (size_t)i * pp, cen + (size_t)k * pp, isg + (size_t)k * pp, pp, lane);
    acc += -dot - sden[k];
  }
  if (lane == 0) sh[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int q = 0; q < 8; q++) t += sh[q];
    partial[blockIdx.x] = t;
  }
}
// the same sum read off a likelihood block that was evaluated for exactly this state: sum_i LL[i][c_i]
// (the block of the next pass is computed right after the split-merge step, see sweep())
__global__ void __launch_bounds__(256) loglik_gather_kernel(const double* __restrict__ LL, int ldl, const int* __restrict__ c,
                                                            int n, double* __restrict__ partial) {
  __shared__ double sh[256];
  double acc = 0.0;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) acc += LL[(size_t)i * ldl + c[i]];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
__global__ void __launch_bounds__(256) reduce_final_kernel(const double* __restrict__ partial, int np,
                                                           double* __restrict__ out) {
  __shared__ double sh[256];
  double acc = 0.0;
  for (int q = threadIdx.x; q < np; q += 256) acc += partial[q];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sh[0];
}

// =============================================================================
// data ingest / pool
// =============================================================================
// fp64 column-major R matrix -> uint8 row-major padded; flags non-integer / out-of-range codes
__global__ void ingest_colmajor_kernel(const double* __restrict__ Xd, int n, int p, int pp,
                                       const int* __restrict__ attr, uint8_t* __restrict__ X, int* __restrict__ bad) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * pp) return;
  int i = (int)(t / pp), j = (int)(t % pp);
  uint8_t o = 0;
  if (j < p) {
    double v = Xd[(size_t)i + (size_t)n * j];
    int iv = (int)v;
    if ((double)iv != v || iv < 1 || iv > attr[j] || iv > 255)
      atomicAdd(bad, 1);
    else
      o = (uint8_t)iv;
  }
  X[t] = o;
}

// prior pool entries (launcher.cpp:67-77,123-129): centre ~ U{1..m_j}, sigma ~ HIG(v_j,w_j,m_j)
__global__ void __launch_bounds__(128) pool_draw_kernel(long long pool_size, int pp, int p, const int* __restrict__ attr,
                                                        const double* __restrict__ v, const double* __restrict__ w,
                                                        RngKey key, int sigma_exact, uint8_t* __restrict__ pcen,
                                                        double* __restrict__ psig, double* __restrict__ pisg,
                                                        double* __restrict__ pden) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= pool_size * pp) return;
  long long e = t / pp;
  int j = (int)(t % pp);
  if (j >= p) {
    pcen[t] = 0;
    psig[t] = 1.0;
    pisg[t] = 0.0;
    pden[t] = 0.0;
    return;
  }
  const int m = attr[j];
  int center;
  double sigma;
  prior_entry_attr(key, e, j, m, v[j], w[j], sigma_exact, &center, &sigma);
  pcen[t] = (uint8_t)center;
  psig[t] = sigma;
  pisg[t] = 1.0 / sigma;
  pden[t] = hamming_den(sigma, m);
}
// per-entry sum of den (one warp per entry, fixed order)
__global__ void pool_sden_kernel(long long pool_size, int pp, const double* __restrict__ pden,
                                 double* __restrict__ psden) {
  long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (w >= pool_size) return;
  double acc = 0.0;
  for (int j = lane; j < pp; j += 32) acc += pden[(size_t)w * pp + j];
  acc = warp_sum(acc);
  if (lane == 0) psden[w] = acc;
}

// validate_state (common_functions.cpp:146-172): #unique(c_i) == total_cls == number of parameter vectors, i.e.
// every label lies in [0, K), every label in [0, K) is used, and the maintained member counts are the true ones.
// Two kernels: recount (integer atomics into `tmp`, zeroed by the caller), then compare.
__global__ void validate_recount_kernel(const int* __restrict__ c, int n, const int* __restrict__ Kptr, int* tmp, int* status) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int k = c[i];
  if (k < 0 || k >= *Kptr)
    atomicOr(status, ST_VALIDATE);
  else
    atomicAdd(&tmp[k], 1);
}
__global__ void validate_compare_kernel(const int* __restrict__ Kptr, int kcap, const int* __restrict__ tmp,
                                        const int* __restrict__ counts, int* status) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= kcap) return;
  const int K = *Kptr;
  if (K < 1 || K > kcap) atomicOr(status, ST_VALIDATE);
  if (k < K ? (tmp[k] <= 0 || tmp[k] != counts[k]) : (tmp[k] != 0)) atomicOr(status, ST_VALIDATE);
}

// Synthetic Hamming-mixture data on the device (spec: code/old_code/data_generation.R:1-101, ham_mix_gen):
// x_ij = c_kj with probability 1/(1+(m_j-1)exp(-1/s)), otherwise one of the other levels uniformly; k = label of i.
__global__ void synth_generate_kernel(int n, int p, int pp, const int* __restrict__ attr, const uint8_t* __restrict__ cent,
                                      const int* __restrict__ labels, double s, RngKey key, uint8_t* __restrict__ X) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * pp) return;
  const int i = (int)(t / pp), j = (int)(t % pp);
  uint8_t o = 0;
  if (j < p) {
    const int m = attr[j];
    const int c = cent[(size_t)labels[i] * pp + j];
    uint32_t r[4];
    philox4x32_10((uint32_t)i, (uint32_t)j, 77u, key.sweep, key.k0, key.k1, r);
    const double u = u01_from_bits(r[0], r[1]);
    const double pm = 1.0 / (1.0 + ((double)m - 1.0) * exp(-1.0 / s));
    if (u < pm) {
      o = (uint8_t)c;
    } else {
      const int shift = 1 + (int)(r[2] % (uint32_t)(m - 1));  // a uniformly chosen other level
      o = (uint8_t)((c - 1 + shift) % m + 1);
    }
  }
  X[t] = o;
}

// Snapshot of one kept iteration (launcher.cpp:140-153) gathered into ONE contiguous staging block on the device, so
// that it leaves with a single copy on a side stream while the next sweep already runs:
// [0] K, status, accepted (ints) | [16] log-likelihood | [64] c_i[n] | centres [Kcap][pp] u8 | sigmas [Kcap][pp] f64
__global__ void snapshot_pack_kernel(const int* __restrict__ Kptr, const int* __restrict__ status,
                                     const int* __restrict__ accepted, const double* __restrict__ loglik,
                                     const int* __restrict__ c, int n, const uint8_t* __restrict__ cen,
                                     const double* __restrict__ sig, int rows_pp, size_t cen_off, size_t sig_off,
                                     unsigned char* __restrict__ out) {
  const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  if (t0 == 0) {
    int* h = reinterpret_cast<int*>(out);
    h[0] = *Kptr;
    h[1] = *status;
    h[2] = *accepted;
    *reinterpret_cast<double*>(out + 16) = *loglik;
  }
  int* oc = reinterpret_cast<int*>(out + 64);
  for (size_t i = t0; i < (size_t)n; i += stride) oc[i] = c[i];
  const uint32_t* cs = reinterpret_cast<const uint32_t*>(cen);
  uint32_t* od = reinterpret_cast<uint32_t*>(out + cen_off);
  for (size_t i = t0; i < (size_t)rows_pp / 4; i += stride) od[i] = cs[i];  // pp is a multiple of 16
  double* os = reinterpret_cast<double*>(out + sig_off);
  for (size_t i = t0; i < (size_t)rows_pp; i += stride) os[i] = sig[i];
}

// initial labels: sample(L, n, replace) - 1 (common_functions.cpp:174-183)
__global__ void init_assign_kernel(int n, int L, const double* __restrict__ u_inj, RngKey key, int* __restrict__ c) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double u = get_u(u_inj, i, key, U_INIT_ASSIGN, (uint32_t)i, 0u);
  int l = (int)((double)L * u + 1.0) - 1;
  c[i] = l >= L ? L - 1 : l;
}

}  // namespace smg
