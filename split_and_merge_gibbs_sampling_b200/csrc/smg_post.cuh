// smg_post.cuh -- posterior summaries on the device.
//
// The reference computes them in R after the run (realdata_analysis/zoo_simulator.R:205-215,339-344):
//   psm <- mcclust.ext::comp.psm(results$c_i + 1)          -> K7, smg_psm.cuh (tensor cores)
//   mcclust.ext::minVI(psm, cls.draw, method = "draws") / mcclust::minbinder -> point estimate among the kept draws
//   mcclust::arandi(point_estimate, ground_truth)           -> adjusted Rand index
//   LaplacesDemon::ESS / IAT(results$loglikelihood)         -> effective sample size of a trace
// Here they are read straight off the device-resident int32 PSM.  Everything that can be an integer is one, so the
// results do not depend on the order of the atomics:
//   Binder:  draws * L_B(c) = sum_{i<j} | draws [c_i = c_j] - PSM_ij |                                 (int64)
//   VI lower bound (Wade & Ghahramani 2018, eq. 13), per row i:  a_i = #{j : c_j = c_i},
//            b_i = sum_j [c_j = c_i] PSM_ij,  s_i = sum_j PSM_ij                                       (int64)
//            VI_lb(c) = (1/n) sum_i [ log2 a_i + log2 (s_i / draws) - 2 log2 (b_i / draws) ]  (fixed-order fp64 sum)
//   ARI:     contingency table by integer atomics, sums of C(n_ab, 2) in int64, Hubert-Arabie ratio in fp64.
//   IAT:     autocovariances by fixed-order block sums, Geyer's initial positive sequence on one thread.
// Rows can be a block [row0, row0 + nrows) of the matrix (the reduce-scattered PSM of smg_chains_reduce_psm): Binder and
// VI are sums over rows, the caller adds the per-rank partial sums.
#pragma once
#include "smg_device.cuh"

namespace smg {

#define POST_MAXC 8  // candidate allocations evaluated per pass over the matrix

// One warp per row i of the block; lanes stride over the columns.  cand: [ncand][n] uint8 labels.
// out_binder[q] += sum_{j > i} |draws [c_i = c_j] - PSM_ij|  (pairs counted once: by the row with the smaller index);
// rowstats [ncand][nrows][3] = (a_i, b_i, s_i)
__global__ void __launch_bounds__(256) psm_point_estimate_kernel(const int* __restrict__ psm, int n, int row0, int nrows,
                                                                 const uint8_t* __restrict__ cand, int ncand, int q0,
                                                                 long long draws, unsigned long long* __restrict__ out_binder,
                                                                 long long* __restrict__ rowstats) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= nrows) return;
  const int i = row0 + warp;
  const int nq = min(POST_MAXC, ncand - q0);
  const int* row = psm + (size_t)warp * n;  // the block's rows are stored from row0
  int ci[POST_MAXC];
  long long bind[POST_MAXC], a[POST_MAXC], b[POST_MAXC];
#pragma unroll
  for (int q = 0; q < POST_MAXC; q++) {
    ci[q] = q < nq ? cand[(size_t)(q0 + q) * n + i] : -1;
    bind[q] = a[q] = b[q] = 0;
  }
  long long s = 0;
  for (int j = lane; j < n; j += 32) {
    const long long pij = row[j];
    s += pij;
#pragma unroll
    for (int q = 0; q < POST_MAXC; q++) {
      if (q < nq) {
        const bool same = cand[(size_t)(q0 + q) * n + j] == ci[q];
        if (same) {
          a[q] += 1;
          b[q] += pij;
        }
        if (j > i) {
          const long long d = (same ? draws : 0) - pij;
          bind[q] += d < 0 ? -d : d;
        }
      }
    }
  }
  auto wsum = [](long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(SMG_FULL, v, o);
    return v;
  };
  s = wsum(s);
#pragma unroll
  for (int q = 0; q < POST_MAXC; q++) {
    if (q < nq) {
      const long long bq = wsum(bind[q]), aq = wsum(a[q]), bb = wsum(b[q]);
      if (lane == 0) {
        atomicAdd(&out_binder[q0 + q], (unsigned long long)bq);
        long long* rs = rowstats + ((size_t)(q0 + q) * nrows + warp) * 3;
        rs[0] = aq;
        rs[1] = bb;
        rs[2] = s;
      }
    }
  }
}

// VI lower bound partial sums of the block's rows, one CTA per candidate, fixed order (thread t adds rows t, t+256, ...;
// then the 256-leaf tree)
__global__ void __launch_bounds__(256) psm_vi_reduce_kernel(const long long* __restrict__ rowstats, int nrows, long long draws,
                                                            double* __restrict__ out_vi_sum) {
  const int q = blockIdx.x;
  __shared__ double sh[256];
  double acc = 0.0;
  const double ld = log2((double)draws);
  for (int r = threadIdx.x; r < nrows; r += 256) {
    const long long* rs = rowstats + ((size_t)q * nrows + r) * 3;
    acc += log2((double)rs[0]) + (log2((double)rs[2]) - ld) - 2.0 * (log2((double)rs[1]) - ld);
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out_vi_sum[q] = sh[0];
}

// contingency table of two labelings (labels 0..kmax-1 each), integer atomics
__global__ void ari_table_kernel(const int* __restrict__ a, const int* __restrict__ b, int n, int kb, int* __restrict__ tab) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(&tab[(size_t)a[i] * kb + b[i]], 1);
}
// out[0] = sum_ab C(n_ab,2), out[1] = sum_a C(n_a.,2), out[2] = sum_b C(n_.b,2)   (one CTA; int64, order-free)
__global__ void __launch_bounds__(256) ari_sums_kernel(const int* __restrict__ tab, int ka, int kb,
                                                       unsigned long long* __restrict__ out) {
  auto c2 = [](long long z) { return (unsigned long long)(z * (z - 1) / 2); };
  unsigned long long sij = 0, sa = 0, sb = 0;
  for (int e = threadIdx.x; e < ka * kb; e += 256) sij += c2(tab[e]);
  for (int r = threadIdx.x; r < ka; r += 256) {
    long long t = 0;
    for (int c = 0; c < kb; c++) t += tab[(size_t)r * kb + c];
    sa += c2(t);
  }
  for (int c = threadIdx.x; c < kb; c += 256) {
    long long t = 0;
    for (int r = 0; r < ka; r++) t += tab[(size_t)r * kb + c];
    sb += c2(t);
  }
  atomicAdd(&out[0], sij);
  atomicAdd(&out[1], sa);
  atomicAdd(&out[2], sb);
}

// autocovariance gamma_k = (1/T) sum_t (x_t - mean)(x_{t+k} - mean), k = 0..T-1, of trace blockIdx.y; one CTA per lag,
// fixed-order sums
__global__ void __launch_bounds__(256) trace_mean_kernel(const double* __restrict__ x, int T, double* __restrict__ mean) {
  const double* xr = x + (size_t)blockIdx.x * T;
  __shared__ double sh[256];
  double acc = 0.0;
  for (int t = threadIdx.x; t < T; t += 256) acc += xr[t];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) mean[blockIdx.x] = sh[0] / (double)T;
}
__global__ void __launch_bounds__(256) trace_autocov_kernel(const double* __restrict__ x, int T, const double* __restrict__ mean,
                                                            double* __restrict__ gamma) {
  const int k = blockIdx.x;
  const double* xr = x + (size_t)blockIdx.y * T;
  const double mu = mean[blockIdx.y];
  __shared__ double sh[256];
  double acc = 0.0;
  for (int t = threadIdx.x; t + k < T; t += 256) acc += (xr[t] - mu) * (xr[t + k] - mu);
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) gamma[(size_t)blockIdx.y * T + k] = sh[0] / (double)T;
}
// Geyer's initial positive sequence: tau = -1 + 2 sum_{k even} (rho_k + rho_{k+1}) while the pair sums stay positive
__global__ void trace_iat_kernel(const double* __restrict__ gamma, int T, int ntraces, double* __restrict__ iat) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= ntraces) return;
  const double* g = gamma + (size_t)r * T;
  if (!(g[0] > 0.0)) {
    iat[r] = 1.0;
    return;
  }
  double tau = -1.0;
  for (int k = 0; k + 1 < T; k += 2) {
    const double pair = g[k] / g[0] + g[k + 1] / g[0];
    if (pair <= 0.0) break;
    tau += 2.0 * pair;
  }
  iat[r] = tau > 1.0 ? tau : 1.0;
}

}  // namespace smg
