// smg_device.cuh -- device-side building blocks shared by every kernel of the
// B200 (sm_100a) Gibbs + split-merge sampler: counter-based uniforms, warp
// reductions, the log-incomplete-beta function, the HIG inverse-CDF sampler and
// the Rcpp-compatible categorical draw.
//
// Reference behaviour being reproduced (file:line in Filippo-Galli/Split_and_merge_Gibbs_sampling):
//   dhamming                  code/common_functions.cpp:355-377
//   norm_const2 / HIG density code/hyperg.cpp:11-48, code/split_merge.cpp:6-18
//   rhig (sigma sampler)      code/hyperg.cpp:346-378  (+ bisec_hyper2 :221-287, lF_conK2 :183-217)
//   Rcpp::sample with probs   call sites neal8.cpp:102, common_functions.cpp:195, split_merge.cpp:215
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace smg {

#define SMG_FULL 0xffffffffu

// ----------------------------------------------------------------------------
// Philox4x32-10 counter-based generator.  One call yields 4 x 32 bits addressed
// by (seed, sweep, site, a, b): any thread can produce the uniform of any draw
// site without a sequential stream (SURVEY Appendix A lists the sites).
// ----------------------------------------------------------------------------
enum Site : uint32_t {
  U_POOL_IDX = 1,   // neal8.cpp:66        (a = observation, b = aux slot)
  U_ALLOC = 2,      // neal8.cpp:102       (a = observation)
  U_CENTER = 3,     // common_functions.cpp:195/199 (a = job, b = attribute)
  U_SIGMA = 4,      // hyperg.cpp:373      (a = job, b = attribute)
  U_SM_PAIR = 5,    // split_merge.cpp:275 (a = 0/1)
  U_SM_LAUNCH = 6,  // split_merge.cpp:346 (a = position in S)
  U_SM_RGIBBS = 7,  // split_merge.cpp:215 (a = position in S, b = scan index)
  U_SM_ACCEPT = 8,  // split_merge.cpp:591
  U_POOL_CENTER = 9,
  U_POOL_SIGMA = 10,
  U_INIT_ASSIGN = 11,  // common_functions.cpp:180
  U_SIGMA_B = 12      // second gamma stream of the lane-pair sigma draw (a = job, b = attribute)
};

struct RngKey {
  uint32_t k0, k1;   // seed
  uint32_t sweep;    // iteration index
  uint32_t sub;      // sub-phase inside the iteration (which update_phi call, ...)
};

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                        uint32_t k1, uint32_t out[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0;
    c1 = n1;
    c2 = n2;
    c3 = n3;
    k0 += W0;
    k1 += W1;
  }
  out[0] = c0;
  out[1] = c1;
  out[2] = c2;
  out[3] = c3;
}

// uniform strictly inside (0,1) with 53 random bits, like R's unif_rand() contract
__host__ __device__ __forceinline__ double u01_from_bits(uint32_t hi, uint32_t lo) {
  uint64_t x = (((uint64_t)hi << 32) | lo) >> 11;
  return ((double)x + 0.5) * (1.0 / 9007199254740992.0);
}

__host__ __device__ __forceinline__ double philox_u01(const RngKey& k, uint32_t site, uint32_t a, uint32_t b) {
  uint32_t o[4];
  philox4x32_10(a, b, site | (k.sub << 8), k.sweep, k.k0, k.k1, o);
  return u01_from_bits(o[0], o[1]);
}

// injected uniform (parity tests / tape replay) or the Philox one
__device__ __forceinline__ double get_u(const double* inj, size_t idx, const RngKey& k, uint32_t site, uint32_t a,
                                        uint32_t b) {
  return inj ? inj[idx] : philox_u01(k, site, a, b);
}

// ----------------------------------------------------------------------------
// warp helpers
// ----------------------------------------------------------------------------
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(SMG_FULL, v, src); }
__device__ __forceinline__ double shfl_xor_d(double v, int m) { return __shfl_xor_sync(SMG_FULL, v, m); }

// deterministic butterfly sum: every lane ends with the same value
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += shfl_xor_d(v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) { return __reduce_add_sync(SMG_FULL, v); }

// The 256-leaf pairwise tree of the block reductions (sh[t] += sh[t + o], o = 128 .. 1), by one warp: lane l holds leaves l, l+32, ..., l+224 in
// v[0..7].  Same pairs in the same order as the shared-memory tree (o = 128, 64, ..., 1), so the same roundings.
// Result on lane 0.
__device__ __forceinline__ double tree256_warp(double* v) {
#pragma unroll
  for (int k = 0; k < 4; k++) v[k] += v[k + 4];
  v[0] += v[2];
  v[1] += v[3];
  v[0] += v[1];
  double x = v[0];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(SMG_FULL, x, o);
  return x;
}
// order-preserving map double -> uint64 (no NaNs expected)
__device__ __forceinline__ uint64_t sort_key(double v) {
  uint64_t b = (uint64_t)__double_as_longlong(v);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_to_double(uint64_t k) {
  uint64_t b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}
// exact warp max of doubles with two REDUX ops (hi word, then lo word among the hi-maxima)
__device__ __forceinline__ uint64_t warp_max_key(uint64_t key) {
  uint32_t hi = (uint32_t)(key >> 32), lo = (uint32_t)key;
  uint32_t mhi = __reduce_max_sync(SMG_FULL, hi);
  uint32_t mlo = __reduce_max_sync(SMG_FULL, hi == mhi ? lo : 0u);
  return ((uint64_t)mhi << 32) | mlo;
}

// ----------------------------------------------------------------------------
// per-attribute Hamming terms (common_functions.cpp:368-376):
//   dhamming = -[x != c]/s - log(1 + (m-1)/exp(1/s))
// ----------------------------------------------------------------------------
__device__ __forceinline__ double hamming_den(double s, int m) { return log(1.0 + ((double)m - 1.0) / exp(1.0 / s)); }

// ----------------------------------------------------------------------------
// log I_x(a,b): modified-Lentz continued fraction, evaluated on the smaller tail.
// Returns both log I_x(a,b) (lower) and log(1 - I_x(a,b)) (upper).
// `lb` = lbeta(a,b) supplied by the caller (computed once per draw).
// ----------------------------------------------------------------------------
// (one out-of-line copy: lgamma expands to a few hundred instructions, and the code that calls it runs once per
// proposal on cold instruction caches)
__device__ __noinline__ double lgamma_ni(double x) { return lgamma(x); }
__device__ __forceinline__ double lbeta_d(double a, double b) { return lgamma_ni(a) + lgamma_ni(b) - lgamma_ni(a + b); }

__device__ __noinline__ double betacf_d(double a, double b, double x) {
  const double EPS = 3e-16, FPMIN = 1e-300;
  double qab = a + b, qap = a + 1.0, qam = a - 1.0;
  double c = 1.0, d = 1.0 - qab * x / qap;
  if (fabs(d) < FPMIN) d = FPMIN;
  d = 1.0 / d;
  double h = d;
  for (int m = 1; m <= 100000; m++) {
    double m2 = 2.0 * m, dm = (double)m;
    double aa = dm * (b - dm) * x / ((qam + m2) * (a + m2));
    d = 1.0 + aa * d;
    if (fabs(d) < FPMIN) d = FPMIN;
    c = 1.0 + aa / c;
    if (fabs(c) < FPMIN) c = FPMIN;
    d = 1.0 / d;
    h *= d * c;
    aa = -(a + dm) * (qab + dm) * x / ((a + m2) * (qap + m2));
    d = 1.0 + aa * d;
    if (fabs(d) < FPMIN) d = FPMIN;
    c = 1.0 + aa / c;
    if (fabs(c) < FPMIN) c = FPMIN;
    d = 1.0 / d;
    double del = d * c;
    h *= del;
    if (fabs(del - 1.0) < EPS) break;
  }
  return h;
}

__device__ __noinline__ void log_ibeta_pair(double x, double a, double b, double lb, double* lower, double* upper) {
  if (x <= 0.0) {
    *lower = -CUDART_INF;
    *upper = 0.0;
    return;
  }
  if (x >= 1.0) {
    *lower = 0.0;
    *upper = -CUDART_INF;
    return;
  }
  double lbt = a * log(x) + b * log1p(-x) - lb;
  if (x < (a + 1.0) / (a + b + 2.0)) {
    double lo = lbt + log(betacf_d(a, b, x)) - log(a);
    *lower = lo;
    *upper = log1p(-exp(lo));
  } else {
    double up = lbt + log(betacf_d(b, a, 1.0 - x)) - log(b);
    *upper = up;
    *lower = log1p(-exp(up));
  }
}

// norm_const2(d=w, c=v, m) through the overflow-free identity (needs v > 1):
//   (w+1) log(m-1) - lbeta(w+1, v-1) - log I_{(m-1)/m}(w+1, v-1)
// The reference evaluates log(d+1)+(d+c)log m - log 2F1(d+c,1;d+2;(m-1)/m) with GSL, which
// overflows for clusters of ~10^3 members; both forms agree to <1e-14 rel
// where the reference is finite (tests/test_oracle_known_answers.py).
// log of the UNNORMALISED incomplete beta function, log int_0^x t^(a-1) (1-t)^(b-1) dt, for 0 < x < 1, a > 0 and ANY real
// b: Gauss's continued fraction for x^a (1-x)^b / a * 2F1(a+b, 1; a+1; x) does not need b > 0 (the complete beta
// function does).  This is what carries v_j <= 1, where the reference's Beta(w+1, v-1) proposal does not exist and
// rhig takes its inverse-CDF branch (hyperg.cpp:359-376: qbeta returns NaN, the switching test is false).
__device__ __noinline__ double log_incbeta_u(double x, double a, double b) {
  return a * log(x) + b * log1p(-x) + log(betacf_d(a, b, x)) - log(a);
}

__device__ __noinline__ double norm_const2_d(double w, double v, double m) {
  if (v <= 1.0) return (w + 1.0) * log(m - 1.0) - log_incbeta_u((m - 1.0) / m, w + 1.0, v - 1.0);
  double a = w + 1.0, b = v - 1.0, lb = lbeta_d(a, b), lo, up;
  // With data the truncation point x = (m-1)/m lies far in the upper tail of Beta(a,b).  Chernoff on the gamma
  // representation X = G_a/(G_a+G_b) gives P(X >= x) <= exp(-(a+b) KL(mu||x)), mu = a/(a+b) < x; once that is below
  // e^-80 the log of the truncated mass is 0 to far below one ulp of the result and the continued fraction (tens of
  // iterations for large shapes) is skipped.
  const double x = (m - 1.0) / m, mu = a / (a + b);
  if (x > mu && b > 0.0 && -(a * log(mu / x) + b * log((1.0 - mu) / (1.0 - x))) > 80.0) return a * log(m - 1.0) - lb;
  log_ibeta_pair(x, a, b, lb, &lo, &up);
  return a * log(m - 1.0) - lb - lo;
}

// logdensity_hig (split_merge.cpp:6-18)
__device__ __noinline__ double logdensity_hig_d(double s, double v, double w, double m) {
  double K = norm_const2_d(w, v, m);
  return K - (v + w) * log(1.0 + exp(-1.0 / s) * (m - 1.0)) - (w + 1.0) / s - 2.0 * log(s);
}

// HIG(v,w,m) draw by inversion of the CDF in u = exp(-1/sigma):
//   CDF(u) = I_x(w+1, v-1) / I_{(m-1)/m}(w+1, v-1),  x = u(m-1)/(1+u(m-1))
// which is exactly what the reference's bisection branch solves (hyperg.cpp:221-287 with
// lF_conK2 :183-217) and the law its Beta-rejection branch samples (hyperg.cpp:359-368).
// Safeguarded Newton on the log of the smaller tail; returns u (sigma = -1/log u).
// v <= 1: CDF(u) = B(x; w+1, v-1) / B(xmax; w+1, v-1) with the unnormalised incomplete beta function (any b)
__device__ __noinline__ double hig_inv_u_smallv(double Omega, double v, double w, double m) {
  const double a = w + 1.0, b = v - 1.0, xmax = (m - 1.0) / m;
  const double target = log(Omega) + log_incbeta_u(xmax, a, b);
  double lo = 0.0, hi = xmax;
  // B(x) ~ x^a / a near 0: a first guess that is exact for small Omega
  double x = xmax * pow(Omega, 1.0 / a);
  if (!(x > 0.0 && x < xmax)) x = 0.5 * xmax;
  for (int it = 0; it < 200; it++) {
    const double lB = log_incbeta_u(x, a, b);
    const double g = lB - target;
    if (g < 0.0)
      lo = x;
    else
      hi = x;
    const double dg = exp((a - 1.0) * log(x) + (b - 1.0) * log1p(-x) - lB);  // d/dx log B(x)
    double xn = x - g / dg;
    const bool newton_ok = (xn > lo && xn < hi);
    if (!newton_ok) xn = 0.5 * (lo + hi);
    const double dx = fabs(xn - x);
    x = xn;
    if (newton_ok && dx <= 1e-8 * x) break;
    if (hi - lo <= 4e-16 * hi) break;
  }
  return x / ((m - 1.0) * (1.0 - x));
}

__device__ __noinline__ double hig_inv_u_d(double Omega, double v, double w, double m) {
  if (v <= 1.0) return hig_inv_u_smallv(Omega, v, w, m);
  const double a = w + 1.0, b = v - 1.0, xmax = (m - 1.0) / m;
  const double lb = lbeta_d(a, b);
  double ltot_lo, ltot_up;
  log_ibeta_pair(xmax, a, b, lb, &ltot_lo, &ltot_up);
  // target lower-tail probability P = Omega * I_tot ; upper Q = 1 - P = tail_tot + I_tot (1 - Omega)
  const double lP = log(Omega) + ltot_lo;
  const double Q = exp(ltot_up) + exp(ltot_lo) * (1.0 - Omega);
  const bool use_lower = lP < -0.6931471805599453;  // P < 1/2
  const double target = use_lower ? lP : log(Q);
  double lo = 0.0, hi = xmax;
  // start from the normal approximation of the Beta quantile (exact enough for the large shapes of big
  // clusters; Newton on the log-tail repairs it for the small shapes of the prior)
  const double mean = a / (a + b);
  double x = mean;
  if (a > 2.0 && b > 2.0) {
    const double sd = sqrt(mean * (1.0 - mean) / (a + b + 1.0));
    const double z = use_lower ? normcdfinv(exp(lP)) : -normcdfinv(Q);
    x = mean + sd * z;
  }
  if (!(x < xmax)) x = (mean < xmax) ? 0.5 * (mean + xmax) : 0.5 * xmax;
  if (!(x > 0.0)) x = 0.5 * fmin(mean, xmax);
  for (int it = 0; it < 200; it++) {
    double l_lo, l_up;
    log_ibeta_pair(x, a, b, lb, &l_lo, &l_up);
    // g is increasing in x for the lower tail, decreasing for the upper tail
    double g = (use_lower ? l_lo : l_up) - target;
    bool below = use_lower ? (g < 0.0) : (g > 0.0);  // root lies to the right of x
    if (below)
      lo = x;
    else
      hi = x;
    double lpdf = (a - 1.0) * log(x) + (b - 1.0) * log1p(-x) - lb;
    double dg = exp(lpdf - (use_lower ? l_lo : l_up));  // d/dx log tail = +-pdf/tail
    if (!use_lower) dg = -dg;
    double xn = x - g / dg;
    const bool newton_ok = (xn > lo && xn < hi);
    if (!newton_ok) xn = 0.5 * (lo + hi);
    const double dx = fabs(xn - x);
    x = xn;
    // a Newton step below 1e-8 relative leaves an error of order 1e-16 (quadratic convergence)
    if (newton_ok && dx <= 1e-8 * x) break;
    if (hi - lo <= 4e-16 * hi) break;
  }
  return x / ((m - 1.0) * (1.0 - x));
}

// ----------------------------------------------------------------------------
// Counter-based uniform sub-stream of one draw site (variable consumption, reproducible):
// Philox counter = (a, b | ctr << 20, site | sub << 8, sweep).
// ----------------------------------------------------------------------------
struct SubStream {
  RngKey key;
  uint32_t site, a, b, ctr;
  double spare;
  bool has_spare;
  __device__ __forceinline__ SubStream(const RngKey& k, uint32_t site_, uint32_t a_, uint32_t b_)
      : key(k), site(site_), a(a_), b(b_), ctr(0), spare(0.0), has_spare(false) {}
  __device__ __forceinline__ double next() {
    if (has_spare) {
      has_spare = false;
      return spare;
    }
    uint32_t o[4];
    philox4x32_10(a, b | (ctr << 20), site | (key.sub << 8), key.sweep, key.k0, key.k1, o);
    ctr++;
    spare = u01_from_bits(o[2], o[3]);
    has_spare = true;
    return u01_from_bits(o[0], o[1]);
  }
};

// Standard normal by the ziggurat method (Marsaglia & Tsang 2000, in Doornik's 2005 form: 128 layers, layer index and
// position from separate random bits).  Double-precision dependent chains are what these draws cost on this GPU
// (a log or normcdfinv is a thousand cycles); the ziggurat's common case is one table look-up, one multiplication and one
// comparison.  Tables: g_zig_x[129] (layer edges, x[128] = 0), g_zig_r[128] = x[i+1]/x[i]; filled by the host once per
// device (smg_zig_init).  One Philox call feeds a whole attempt: o[0..1] -> position, o[2] & 127 -> layer,
// (o[2] >> 7, o[3]) -> the uniform of the caller's acceptance test.
#define SMG_ZIG_C 128
#define SMG_ZIG_R 3.442619855899
#define SMG_ZIG_V 9.91256303526217e-3
__device__ double g_zig_x[SMG_ZIG_C + 1];
__device__ double g_zig_r[SMG_ZIG_C];

struct NormU {
  double x;  // N(0,1)
  double u;  // independent U(0,1)
};
__device__ __forceinline__ void substream_raw(SubStream& rs, uint32_t o[4]) {
  philox4x32_10(rs.a, rs.b | (rs.ctr << 20), rs.site | (rs.key.sub << 8), rs.key.sweep, rs.key.k0, rs.key.k1, o);
  rs.ctr++;
}
__device__ __noinline__ NormU zig_normal_u(SubStream& rs) {
  NormU out;
  for (;;) {
    uint32_t o[4];
    substream_raw(rs, o);
    const double up = 2.0 * u01_from_bits(o[0], o[1]) - 1.0;  // position in (-1, 1)
    const int i = (int)(o[2] & (SMG_ZIG_C - 1));
    out.u = ((double)((((uint64_t)(o[2] >> 7)) << 28) | (uint64_t)(o[3] >> 4)) + 0.5) * (1.0 / 9007199254740992.0);  // 53 bits
    const double xi = __ldg(&g_zig_x[i]);
    if (fabs(up) < __ldg(&g_zig_r[i])) {  // inside the layer's rectangle: 98.8% of the attempts
      out.x = up * xi;
      return out;
    }
    if (i == 0) {  // tail beyond R
      double x, y;
      do {
        uint32_t q[4];
        substream_raw(rs, q);
        x = log(u01_from_bits(q[0], q[1])) / SMG_ZIG_R;
        y = log(u01_from_bits(q[2], q[3]));
      } while (-2.0 * y < x * x);
      out.x = up < 0.0 ? x - SMG_ZIG_R : SMG_ZIG_R - x;
      return out;
    }
    {  // wedge
      const double x = up * xi, xn = __ldg(&g_zig_x[i + 1]);
      const double f0 = exp(-0.5 * (xi * xi - x * x)), f1 = exp(-0.5 * (xn * xn - x * x));
      uint32_t q[4];
      substream_raw(rs, q);
      if (f1 + u01_from_bits(q[0], q[1]) * (f0 - f1) < 1.0) {
        out.x = x;
        return out;
      }
    }
  }
}

// Gamma(shape, 1) by Marsaglia & Tsang (2000); shape < 1 through Gamma(shape+1) * U^(1/shape)
__device__ __noinline__ double gamma_draw_d(SubStream& rs, double shape) {
  double boost = 1.0;
  if (shape < 1.0) {
    boost = pow(rs.next(), 1.0 / shape);
    shape += 1.0;
  }
  const double d = shape - 1.0 / 3.0, c = rsqrt(9.0 * d);
  for (int it = 0; it < 64; it++) {
    const NormU nu = zig_normal_u(rs);
    const double x = nu.x;
    double vv = 1.0 + c * x;
    if (vv <= 0.0) continue;
    vv = vv * vv * vv;
    const double x2 = x * x;
    if (nu.u < 1.0 - 0.0331 * x2 * x2) return boost * d * vv;
    if (log(nu.u) < 0.5 * x2 + d * (1.0 - vv + log(vv))) return boost * d * vv;
  }
  return boost * d;  // not reached in practice (acceptance > 95% per round)
}

// sigma ~ HIG(v,w,m), returned as u = exp(-1/sigma).  Same law as hyperg.cpp:346-378:
// x ~ Beta(w+1, v-1) conditioned on x <= (m-1)/m, u = x/((m-1)(1-x))  (the reference's own Beta branch,
// hyperg.cpp:359-368); after 8 rejected proposals (the truncation keeps little mass) one exact
// inverse-CDF draw (the reference's bisection branch) finishes -- the mixture is still the exact law.
__device__ inline double hig_draw_u_d(SubStream& rs, double v, double w, double m) {
  const double a = w + 1.0, b = v - 1.0;
  for (int attempt = 0; attempt < 8 && b > 0.0; attempt++) {  // b <= 0 (v <= 1): no Beta proposal, inversion only
    const double ga = gamma_draw_d(rs, a), gb = gamma_draw_d(rs, b);
    // x = ga/(ga+gb) <= (m-1)/m  <=>  u = x/((m-1)(1-x)) = ga/((m-1) gb) <= 1: one division
    const double u = ga / ((m - 1.0) * gb);
    if (u > 0.0 && u < 1.0) return u;
  }
  return hig_inv_u_d(rs.next(), v, w, m);
}

// The same draw by a group of PHI_G consecutive lanes (all of them call this; `g` = lane index inside the group,
// `gmask` its lane mask, `gbase` its first lane).  The two gamma variates of a proposal are independent, so lanes 0
// and 1 draw them side by side from two sub-streams -- the dependency chain of a proposal is one gamma draw
// instead of two.  Result valid on lane 0 of the group.
#define PHI_G 8
// reference form: the algorithm below written with the out-of-line building blocks; also the path of shapes < 1
__device__ __noinline__ double hig_draw_u_grp_ref(const RngKey& key, uint32_t sa, uint32_t sb, double v, double w, double m, int g,
                                                  unsigned gmask, int gbase) {
  SubStream rs(key, g == 1 ? U_SIGMA_B : U_SIGMA, sa, sb);
  const double a = w + 1.0, b = v - 1.0;
  double res = 0.0;
  for (int attempt = 0; attempt < 8 && b > 0.0; attempt++) {
    double gm = 0.0;
    if (g < 2) gm = gamma_draw_d(rs, g == 0 ? a : b);
    const double gb = __shfl_sync(gmask, gm, gbase + 1);
    int ok = 0;
    if (g == 0) {
      const double u = gm / ((m - 1.0) * gb);  // = x/((m-1)(1-x)) with x = gm/(gm+gb); u < 1 <=> x < (m-1)/m
      if (u > 0.0 && u < 1.0) {
        ok = 1;
        res = u;
      }
    }
    if (__shfl_sync(gmask, ok, gbase)) return res;
  }
  if (g == 0) res = hig_inv_u_d(rs.next(), v, w, m);
  return res;
}

// One Marsaglia-Tsang iteration on Philox block `blk` of the lane's sub-stream: 1 accepted (*val = the variate),
// 0 rejected, 2 the normal is not in its ziggurat rectangle (zig_normal_u consumes further blocks)
__device__ __forceinline__ int mt_iteration(const RngKey& key, uint32_t site, uint32_t sa, uint32_t sb, uint32_t blk, double d,
                                            double c, double* val, const double* zx, const double* zr) {
  uint32_t o[4];
  philox4x32_10(sa, sb | (blk << 20), site | (key.sub << 8), key.sweep, key.k0, key.k1, o);
  const double up = 2.0 * u01_from_bits(o[0], o[1]) - 1.0;
  const int zi = (int)(o[2] & (SMG_ZIG_C - 1));
  const double un = ((double)((((uint64_t)(o[2] >> 7)) << 28) | (uint64_t)(o[3] >> 4)) + 0.5) * (1.0 / 9007199254740992.0);
  const double x = up * zx[zi];
  if (!(fabs(up) < zr[zi])) return 2;
  double vv = 1.0 + c * x;
  if (vv <= 0.0) return 0;
  vv = vv * vv * vv;
  const double x2 = x * x;
  if (un < 1.0 - 0.0331 * x2 * x2 || log(un) < 0.5 * x2 + d * (1.0 - vv + log(vv))) {
    *val = d * vv;
    return 1;
  }
  return 0;
}

// The production form: the same draws from the same Philox blocks as the reference form above (every path of the library
// calls this one function), arranged for latency -- these draws sit on the critical path of every restricted scan and of
// update_phi, and a phase ends when the SLOWEST of its few hundred rejection samplers does.
//   * the common case is inline (Philox block -> ziggurat rectangle -> Marsaglia-Tsang squeeze or its exact log test);
//     ziggurat wedge / tail, shapes < 1 and the inverse-CDF fallback stay out of line;
//   * iteration k of a gamma draw reads Philox block k as long as no earlier iteration left its rectangle, so the 8 lanes
//     of the group evaluate iterations 0..3 of BOTH gamma variates at once (lane 2k + s: iteration k of variate s) and the
//     first one in order that does not reject is taken: a rejection costs no extra round trip.
// zx / zr: the ziggurat tables (g_zig_x, g_zig_r), or a copy of them in shared memory -- the look-up follows the Philox
// block in the dependency chain, and a global load there is a trip to L2 in kernels that keep little L1
__device__ __forceinline__ double hig_draw_u_grp(const RngKey& key, uint32_t sa, uint32_t sb, double v, double w, double m, int g,
                                                 unsigned gmask, int gbase, const double* zx = g_zig_x,
                                                 const double* zr = g_zig_r) {
  const double a = w + 1.0, b = v - 1.0;
  if (b < 1.0) return hig_draw_u_grp_ref(key, sa, sb, v, w, m, g, gmask, gbase);  // uniform over the group
  const int sv = g & 1;  // which variate this lane works on
  const uint32_t site = sv ? U_SIGMA_B : U_SIGMA;
  const double shape = sv ? b : a;
  const double d = shape - 1.0 / 3.0, c = rsqrt(9.0 * d);
  // speculative round: lane g evaluates iteration g >> 1
  double val = 0.0;
  const int oc = mt_iteration(key, site, sa, sb, (uint32_t)(g >> 1), d, c, &val, zx, zr);
  const unsigned nz = (__ballot_sync(gmask, oc != 0) >> gbase) & 0xffu;
  const unsigned mine = (nz >> sv) & 0x55u;                 // iterations of my variate that did not reject: bits 0,2,4,6
  const int kf = mine ? (__ffs(mine) - 1) >> 1 : 4;         // the first of them (4: none)
  const int src = gbase + 2 * (kf < 4 ? kf : 0) + sv;
  const int oc_f = __shfl_sync(gmask, oc, src);
  const double val_f = __shfl_sync(gmask, val, src);
  uint32_t ctr = (uint32_t)kf;  // blocks consumed so far by my variate
  bool have = false;
  double gm = 0.0;
  if (kf < 4 && oc_f == 1) {
    gm = val_f;
    ctr = (uint32_t)kf + 1u;
    have = true;
  }
  double res = 0.0;
  for (int attempt = 0; attempt < 8; attempt++) {
    if (g < 2 && !have) {  // sequential continuation (a wedge / tail normal, four rejections in a row, a second attempt)
      gm = d;  // gamma_draw_d's exit after 64 rejected normals (not reached in practice)
      for (int it = 0; it < 64; it++) {
        double vq = 0.0;
        int r = mt_iteration(key, site, sa, sb, ctr, d, c, &vq, zx, zr);
        ctr++;
        if (r == 2) {  // zig_normal_u from this very block
          SubStream rs(key, site, sa, sb);
          rs.ctr = ctr - 1;
          const NormU nu = zig_normal_u(rs);
          ctr = rs.ctr;
          double vv = 1.0 + c * nu.x;
          r = 0;
          if (vv > 0.0) {
            vv = vv * vv * vv;
            const double x2 = nu.x * nu.x;
            if (nu.u < 1.0 - 0.0331 * x2 * x2 || log(nu.u) < 0.5 * x2 + d * (1.0 - vv + log(vv))) {
              vq = d * vv;
              r = 1;
            }
          }
        }
        if (r == 1) {
          gm = vq;
          break;
        }
      }
    }
    have = false;
    const double gb = __shfl_sync(gmask, gm, gbase + 1);
    int ok = 0;
    if (g == 0) {
      const double u = gm / ((m - 1.0) * gb);
      if (u > 0.0 && u < 1.0) {
        ok = 1;
        res = u;
      }
    }
    if (__shfl_sync(gmask, ok, gbase)) return res;
  }
  if (g == 0) {
    uint32_t o[4];
    philox4x32_10(sa, sb | (ctr << 20), site | (key.sub << 8), key.sweep, key.k0, key.k1, o);
    res = hig_inv_u_d(u01_from_bits(o[0], o[1]), v, w, m);
  }
  return res;
}

// ----------------------------------------------------------------------------
// R's revsort (src/main/sort.c): heapsort into DESCENDING order carrying an index
// array; the tie permutation is part of the contract for centre draws.
// Small per-thread arrays (n <= SMG_MAX_LEVELS).
// ----------------------------------------------------------------------------
#define SMG_MAX_LEVELS 64

__device__ inline void revsort_d(double* a0, int* ib0, int n) {
  if (n <= 1) return;
  double* a = a0 - 1;
  int* ib = ib0 - 1;
  int l = (n >> 1) + 1, ir = n, i, j, ii;
  double ra;
  for (;;) {
    if (l > 1) {
      l = l - 1;
      ra = a[l];
      ii = ib[l];
    } else {
      ra = a[ir];
      ii = ib[ir];
      a[ir] = a[1];
      ib[ir] = ib[1];
      if (--ir == 1) {
        a[1] = ra;
        ib[1] = ii;
        return;
      }
    }
    i = l;
    j = l << 1;
    while (j <= ir) {
      if (j < ir && a[j] > a[j + 1]) ++j;
      if (ra > a[j]) {
        a[i] = a[j];
        ib[i] = ib[j];
        j += (i = j);
      } else
        j = ir + 1;
    }
    a[i] = ra;
    ib[i] = ii;
  }
}

// Rcpp::sample(x, 1, true, probs) for a short probability vector held by one thread
// (centre draws, common_functions.cpp:195): normalise, revsort descending, cumulate,
// first j < n-1 with u <= cum_j else n-1.  Returns the 0-based position. p[] is clobbered.
__device__ inline int sample_probs_small(double* p, int n, double u) {
  int perm[SMG_MAX_LEVELS];
  double sum = 0.0;
  for (int i = 0; i < n; i++) sum += p[i];
  for (int i = 0; i < n; i++) {
    p[i] /= sum;
    perm[i] = i + 1;
  }
  revsort_d(p, perm, n);
  for (int i = 1; i < n; i++) p[i] += p[i - 1];
  int j;
  for (j = 0; j < n - 1; j++)
    if (u <= p[j]) break;
  return perm[j] - 1;
}

}  // namespace smg
