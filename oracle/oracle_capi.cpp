// =============================================================================
// oracle/oracle_capi.cpp -- TEST INFRASTRUCTURE ONLY.
// extern "C" surface over smg_oracle.hpp for tests/ (ctypes), __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs.  The product library
// (split_and_merge_gibbs_sampling_b200/csrc) never links or loads this file.
// PARITY UNPINNED by the reference's own tests (it has none) - see smg_oracle.hpp.
// =============================================================================
#include <chrono>
#include <cstring>
#include <thread>

#include "smg_oracle.hpp"

using namespace smg_oracle;

extern "C" {

struct orc_opts {
  int counted, stable_hig, sigma_inverse_cdf;
  double bisect_tol;
  int bisect_max, validate;
  int det_i1;  // -1: select_observations_random; >= 0: select_observations_deterministic with this i_1
};

struct orc_data {
  int n, p;
  const double* X;  // column-major n x p
  const int* attrisize;
  double gamma;
  const double* v;
  const double* w;
};
}

namespace {
Opts mk_opts(const orc_opts* o) {
  Opts r;
  if (o) {
    r.counted = o->counted;
    r.stable_hig = o->stable_hig;
    r.sigma_inverse_cdf = o->sigma_inverse_cdf;
    r.bisect_tol = o->bisect_tol;
    r.bisect_max = o->bisect_max;
    r.validate = o->validate;
    r.det_i1 = o->det_i1;
  }
  return r;
}
AuxData mk_data(const orc_data* d) {
  AuxData a;
  a.n = d->n;
  a.p = d->p;
  a.X.assign(d->X, d->X + (size_t)d->n * d->p);
  a.attrisize.assign(d->attrisize, d->attrisize + d->p);
  a.gamma = d->gamma;
  a.v.assign(d->v, d->v + d->p);
  a.w.assign(d->w, d->w + d->p);
  return a;
}
State mk_state(int n, int p, int K, const int* c, const double* center, const double* sigma) {
  State s;
  s.K = K;
  s.c.assign(c, c + n);
  s.center.resize(K);
  s.sigma.resize(K);
  for (int k = 0; k < K; k++) {
    s.center[k].assign(center + (size_t)k * p, center + (size_t)(k + 1) * p);
    s.sigma[k].assign(sigma + (size_t)k * p, sigma + (size_t)(k + 1) * p);
  }
  return s;
}
// returns 0 ok, -1 capacity exceeded
int put_state(const State& s, int p, int Kcap, int* K, int* c, double* center, double* sigma) {
  if (s.K > Kcap) return -1;
  *K = s.K;
  if (c) std::copy(s.c.begin(), s.c.end(), c);
  for (int k = 0; k < s.K; k++) {
    if (center) std::copy(s.center[k].begin(), s.center[k].end(), center + (size_t)k * p);
    if (sigma) std::copy(s.sigma[k].begin(), s.sigma[k].end(), sigma + (size_t)k * p);
  }
  return 0;
}
void put_err(char* err, int errlen, const char* msg) {
  if (err && errlen > 0) {
    std::strncpy(err, msg, errlen - 1);
    err[errlen - 1] = 0;
  }
}
void put_diag(const Diag& d, long long* out) {
  if (!out) return;
  out[0] = (long long)d.exact_pos_ties;
  out[1] = (long long)d.near_ties;
  out[2] = (long long)d.walker_trigger;
  out[3] = (long long)d.bisect_maxiter;
}
Pool mk_pool(long pool_size, int p, const double* pc, const double* ps) {
  Pool pool;
  pool.center.resize(pool_size);
  pool.sigma.resize(pool_size);
  for (long e = 0; e < pool_size; e++) {
    pool.center[e].assign(pc + (size_t)e * p, pc + (size_t)(e + 1) * p);
    pool.sigma[e].assign(ps + (size_t)e * p, ps + (size_t)(e + 1) * p);
  }
  return pool;
}
long put_log(const Rng& r, long cap, int* lphase, int* lsite, int* la, int* lb, double* lu) {
  long n = (long)r.log.size();
  for (long i = 0; i < n && i < cap; i++) {
    if (lphase) lphase[i] = r.log[i].phase;
    if (lsite) lsite[i] = r.log[i].site;
    if (la) la[i] = r.log[i].a;
    if (lb) lb[i] = r.log[i].b;
    if (lu) lu[i] = r.log[i].u;
  }
  return n;
}
}  // namespace

extern "C" {

// ------------------------------ scalar functions ------------------------------
double orc_dhamming(int x, int c, double s, int m) { return dhamming(x, c, s, m); }
double orc_hyp2f1(double a, double b, double c, double x, int* status) { return hyp2f1_series(a, b, c, x, status); }
double orc_norm_const2(double d, double c, double m, int stable, int* err) {
  if (err) *err = 0;
  try {
    return norm_const2(d, c, m, stable != 0);
  } catch (const std::exception&) {
    if (err) *err = 1;
    return NAN;
  }
}
double orc_logdensity_hig(double s, double v, double w, double m, int stable, int* err) {
  if (err) *err = 0;
  try {
    return logdensity_hig(s, v, w, m, stable != 0);
  } catch (const std::exception&) {
    if (err) *err = 1;
    return NAN;
  }
}
double orc_lF_conK2(double u, double d, double c, double m, double lK, int stable) {
  return lF_conK2(u, d, c, m, lK, stable != 0);
}
double orc_bisec_hyper2(double d, double c, double m, double Omega, const orc_opts* o, int* err) {
  if (err) *err = 0;
  try {
    return bisec_hyper2(d, c, m, Omega, mk_opts(o));
  } catch (const std::exception&) {
    if (err) *err = 1;
    return NAN;
  }
}
double orc_pbeta(double x, double a, double b) { return pbeta_(x, a, b); }
double orc_log_ibeta(double x, double a, double b) { return log_ibeta(x, a, b); }
int orc_rhig_beta_branch(double v, double w, double m) { return rhig_beta_branch(v, w, m) ? 1 : 0; }
void orc_revsort(double* a, int* ib, int n) { revsort(a, ib, n); }

// sample_initial_assignment (common_functions.cpp:174-183) from a uniform tape of n draws
int orc_initial_assignment(int L, int n, const double* tape, int* out) {
  try {
    TapeRng r(tape, (size_t)n);
    std::vector<int> c = sample_initial_assignment(r, L, n);
    for (int i = 0; i < n; i++) out[i] = c[i];
    return 0;
  } catch (const std::exception&) {
    return -1;
  }
}

// one Rcpp-style sample(x,1,true,probs) draw; flags[0]=exact positive tie, flags[1]=near tie
int orc_sample_probs_one(const double* probs, int n, double u, int* flags) {
  Diag dg;
  try {
    int r = sample_probs_one(probs, n, u, &dg);
    if (flags) {
      flags[0] = (int)dg.exact_pos_ties;
      flags[1] = (int)dg.near_ties;
    }
    return r;
  } catch (const std::exception&) {
    return -1;
  }
}

// n draws of rhig(1,v,w,m) from a seeded stream (distribution tests)
int orc_rhig_many(double v, double w, double m, long n, unsigned long long seed, const orc_opts* o, double* out) {
  MtRng r(seed);
  Opts op = mk_opts(o);
  try {
    for (long i = 0; i < n; i++) out[i] = rhig1(r, v, w, m, op);
  } catch (const std::exception&) {
    return 1;
  }
  return 0;
}
// rhig on the inverse-CDF branch for explicit Omega values (u-space result, before -1/log)
int orc_rhig_u_from_omega(double v, double w, double m, long n, const double* omega, const orc_opts* o, double* u_out) {
  Opts op = mk_opts(o);
  try {
    for (long i = 0; i < n; i++) u_out[i] = bisec_hyper2(w, v, m, omega[i], op);
  } catch (const std::exception&) {
    return 1;
  }
  return 0;
}
void orc_rbeta_many(double a, double b, long n, unsigned long long seed, double* out) {
  MtRng r(seed);
  for (long i = 0; i < n; i++) out[i] = rbeta_cheng(r, a, b);
}

// ------------------------------ state-level -----------------------------------
double orc_loglik(const orc_data* d, int K, const int* c, const double* center, const double* sigma) {
  AuxData a = mk_data(d);
  State s = mk_state(d->n, d->p, K, c, center, sigma);
  return compute_loglikelihood(s, a);
}

// per-observation x per-cluster log-likelihood block and integer mismatch counts
void orc_ll_block(const orc_data* d, int K, const double* center, const double* sigma, double* LL /*n x K row-major*/,
                  int* mism /*n x K*/) {
  for (int i = 0; i < d->n; i++)
    for (int k = 0; k < K; k++) {
      double ll = 0.0;
      int mm = 0;
      for (int j = 0; j < d->p; j++) {
        int x = (int)d->X[(size_t)i + (size_t)d->n * j];
        int cc = (int)center[(size_t)k * d->p + j];
        ll += dhamming(x, cc, sigma[(size_t)k * d->p + j], d->attrisize[j]);
        mm += (x != cc);
      }
      if (LL) LL[(size_t)i * K + k] = ll;
      if (mism) mism[(size_t)i * K + k] = mm;
    }
}

// one Neal-8 pass over all observations (no update_phi), uniforms from `tape`
// in the reference's order: per observation m pool-index draws then 1 allocation draw.
int orc_neal8_scan(const orc_data* d, int m_aux, int Kcap, int* K, int* c, double* center, double* sigma,
                   long pool_size, const double* pool_center, const double* pool_sigma, const double* tape,
                   long tape_len, const orc_opts* o, long long* diag, char* err, int errlen) {
  try {
    AuxData a = mk_data(d);
    State s = mk_state(d->n, d->p, *K, c, center, sigma);
    Pool pool = mk_pool(pool_size, d->p, pool_center, pool_sigma);
    TapeRng r(tape, (size_t)tape_len);
    Diag dg;
    neal8_scan(a, s, m_aux, pool, r, mk_opts(o), &dg);
    put_diag(dg, diag);
    if (put_state(s, d->p, Kcap, K, c, center, sigma)) {
      put_err(err, errlen, "Kcap exceeded");
      return 2;
    }
    return 0;
  } catch (const std::exception& e) {
    put_err(err, errlen, e.what());
    return 1;
  }
}

// histogram H[k][j][a] (a < mmax) of members, and member counts
void orc_histogram(const orc_data* d, int K, const int* c, int mmax, int* H, int* counts) {
  std::fill(H, H + (size_t)K * d->p * mmax, 0);
  std::fill(counts, counts + K, 0);
  for (int i = 0; i < d->n; i++) {
    int k = c[i];
    counts[k]++;
    for (int j = 0; j < d->p; j++) {
      int x = (int)d->X[(size_t)i + (size_t)d->n * j];
      if (x >= 1 && x <= d->attrisize[j]) H[((size_t)k * d->p + j) * mmax + (x - 1)]++;
    }
  }
}

// update_phi on `clusters` (ncl==0 => all).  Uniforms consumed sequentially from `tape`;
// the tagged log of the draws is returned so a test can rebuild addressed arrays.
int orc_update_phi(const orc_data* d, int K, const int* c, double* center, double* sigma, const int* clusters, int ncl,
                   const double* tape, long tape_len, long* consumed, const orc_opts* o, long log_cap, int* lsite,
                   int* la, int* lb, double* lu, long* log_n, char* err, int errlen) {
  try {
    AuxData a = mk_data(d);
    State s = mk_state(d->n, d->p, K, c, center, sigma);
    TapeRng r(tape, (size_t)tape_len);
    r.logging = log_cap > 0;
    std::vector<int> cl(clusters, clusters + ncl);
    update_phi(s, a, r, mk_opts(o), cl);
    if (consumed) *consumed = (long)r.pos;
    int Kd;
    put_state(s, d->p, K, &Kd, nullptr, center, sigma);
    long nlog = put_log(r, log_cap, nullptr, lsite, la, lb, lu);
    if (log_n) *log_n = nlog;
    return 0;
  } catch (const std::exception& e) {
    put_err(err, errlen, e.what());
    return 1;
  }
}

// softmax centre probabilities of one cluster given sigma (compute_prob_centers)
void orc_prob_centers(const orc_data* d, int nidx, const int* indices, const double* sigma, int mmax, double* prob,
                      double* freq) {
  AuxData a = mk_data(d);
  std::vector<int> idx(indices, indices + nidx);
  std::vector<double> sg(sigma, sigma + d->p);
  std::vector<std::vector<double>> fr;
  auto pr = compute_prob_centers(a, idx, sg, &fr);
  for (int j = 0; j < d->p; j++)
    for (int q = 0; q < mmax; q++) {
      prob[(size_t)j * mmax + q] = q < (int)pr[j].size() ? pr[j][q] : 0.0;
      if (freq) freq[(size_t)j * mmax + q] = q < (int)fr[j].size() ? fr[j][q] : 0.0;
    }
}

// One split_and_merge call with every intermediate exposed.
// terms[24]: log_alpha, lg[3], pri[3], ll[3], gs_phi[3], gs_c, log_prior, log_lik, log_prop, log_ratio, u_accept
int orc_split_merge(const orc_data* d, int t, int rr, int Kcap, int* K, int* c, double* center, double* sigma,
                    const double* tape, long tape_len, long* consumed, const orc_opts* o,
                    int* info /*[8]: i1,i2,nS,is_split,accepted,KSL,KML,Kstar*/, int* S /*[n]*/,
                    int* c_SL, double* center_SL, double* sigma_SL, int* c_ML, double* center_ML, double* sigma_ML,
                    int* c_star, double* center_star, double* sigma_star, double* terms, long log_cap, int* lphase,
                    int* lsite, int* la, int* lb, double* lu, long* log_n, long long* diag, char* err, int errlen) {
  try {
    AuxData a = mk_data(d);
    State s = mk_state(d->n, d->p, *K, c, center, sigma);
    TapeRng r(tape, (size_t)tape_len);
    r.logging = log_cap > 0;
    Diag dg;
    SMTrace tr;
    int acc = split_and_merge(s, a, t, rr, r, mk_opts(o), &dg, &tr);
    if (consumed) *consumed = (long)r.pos;
    put_diag(dg, diag);
    info[0] = tr.i_1;
    info[1] = tr.i_2;
    info[2] = (int)tr.S.size();
    info[3] = tr.is_split;
    info[4] = acc;
    std::copy(tr.S.begin(), tr.S.end(), S);
    int rc = 0;
    rc |= put_state(tr.split_launch, d->p, Kcap, &info[5], c_SL, center_SL, sigma_SL);
    rc |= put_state(tr.merge_launch, d->p, Kcap, &info[6], c_ML, center_ML, sigma_ML);
    rc |= put_state(tr.star, d->p, Kcap, &info[7], c_star, center_star, sigma_star);
    rc |= put_state(s, d->p, Kcap, K, c, center, sigma);
    if (rc) {
      put_err(err, errlen, "Kcap exceeded");
      return 2;
    }
    const MHTerms& T = tr.terms;
    double tt[24] = {T.log_alpha, T.lg[0],     T.lg[1],     T.lg[2],     T.pri[0], T.pri[1],    T.pri[2],  T.ll[0],
                     T.ll[1],     T.ll[2],     T.gs_phi[0], T.gs_phi[1], T.gs_phi[2], T.gs_c,  T.log_prior, T.log_lik,
                     T.log_prop,  T.log_ratio, tr.u_accept, 0, 0, 0, 0, 0};
    std::copy(tt, tt + 24, terms);
    long nlog = put_log(r, log_cap, lphase, lsite, la, lb, lu);
    if (log_n) *log_n = nlog;
    return 0;
  } catch (const std::exception& e) {
    put_err(err, errlen, e.what());
    return 1;
  }
}

// MH building blocks on explicit states (gamma_star, gamma)
double orc_logprobgs_phi(const orc_data* d, int Ks, const int* cs, const double* centers_s, const double* sigmas_s,
                         int Kg, const int* cg, const double* centers_g, const double* sigmas_g, int chosen,
                         int stable, double* parts /*[2] centre, sigma*/, int* err) {
  if (err) *err = 0;
  try {
    AuxData a = mk_data(d);
    State gs = mk_state(d->n, d->p, Ks, cs, centers_s, sigmas_s);
    State g = mk_state(d->n, d->p, Kg, cg, centers_g, sigmas_g);
    Opts o;
    o.stable_hig = stable != 0;
    return logprobgs_phi(gs, g, a, chosen, o, parts ? &parts[0] : nullptr, parts ? &parts[1] : nullptr);
  } catch (const std::exception&) {
    if (err) *err = 1;
    return NAN;
  }
}
double orc_priors(const orc_data* d, const double* sigma_c, int stable, int* err) {
  if (err) *err = 0;
  try {
    double pr = 0;
    for (int j = 0; j < d->p; j++) {
      pr -= std::log((double)d->attrisize[j]);
      pr += logdensity_hig(sigma_c[j], d->v[j], d->w[j], d->attrisize[j], stable != 0);
    }
    return pr;
  } catch (const std::exception&) {
    if (err) *err = 1;
    return NAN;
  }
}

// prior pool (launcher.cpp:67-77) from a seeded stream; sigma on either sampler
int orc_draw_pool(const orc_data* d, long pool_size, unsigned long long seed, const orc_opts* o, double* pool_center,
                  double* pool_sigma) {
  try {
    AuxData a = mk_data(d);
    MtRng r(seed);
    Pool pool;
    Diag dg;
    draw_pool(pool, (size_t)pool_size, a, r, mk_opts(o), &dg);
    for (long e = 0; e < pool_size; e++) {
      std::copy(pool.center[e].begin(), pool.center[e].end(), pool_center + (size_t)e * d->p);
      std::copy(pool.sigma[e].begin(), pool.sigma[e].end(), pool_sigma + (size_t)e * d->p);
    }
    return 0;
  } catch (const std::exception&) {
    return 1;
  }
}

// Full chain (launcher.cpp:7-174) from a seeded Mersenne-Twister stream.
// centers/sigmas snapshots are returned only for the LAST kept iteration (Kcap x p) to keep the ABI flat.
int orc_run_chain(const orc_data* d, int m_aux, int iterations, int L, const int* c_init, int burnin, int t, int rr,
                  int neal8, int split_merge, int n8_step, int sam_step, int thinning, unsigned long long seed,
                  const orc_opts* o, long pool_size_override, int* total_cls, int* c_i /*iterations*n or null*/,
                  double* loglik, int* accepted, int* final_ass, int Kcap, double* last_center, double* last_sigma,
                  double* seconds, long long* diag, char* err, int errlen) {
  try {
    AuxData a = mk_data(d);
    ChainArgs ca;
    ca.m = m_aux;
    ca.iterations = iterations;
    ca.L = L;
    ca.c_i = c_init;
    ca.burnin = burnin;
    ca.t = t;
    ca.r = rr;
    ca.neal8 = neal8 != 0;
    ca.split_merge = split_merge != 0;
    ca.n8_step_size = n8_step;
    ca.sam_step_size = sam_step;
    ca.thinning = thinning;
    ca.pool_size_override = (size_t)pool_size_override;
    MtRng r(seed);
    auto t0 = std::chrono::steady_clock::now();
    Results res = run_markov_chain(a, ca, r, mk_opts(o));
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    for (int it = 0; it < iterations; it++) {
      if (total_cls) total_cls[it] = res.total_cls[it];
      if (loglik) loglik[it] = res.loglikelihood[it];
      if (accepted) accepted[it] = res.accepted[it];
      if (c_i && !res.c_i[it].empty()) std::copy(res.c_i[it].begin(), res.c_i[it].end(), c_i + (size_t)it * d->n);
    }
    if (final_ass) std::copy(res.final_ass.begin(), res.final_ass.end(), final_ass);
    if (last_center && iterations > 0 && (int)res.centers[iterations - 1].size() <= Kcap)
      for (size_t k = 0; k < res.centers[iterations - 1].size(); k++) {
        std::copy(res.centers[iterations - 1][k].begin(), res.centers[iterations - 1][k].end(), last_center + k * d->p);
        std::copy(res.sigmas[iterations - 1][k].begin(), res.sigmas[iterations - 1][k].end(), last_sigma + k * d->p);
      }
    put_diag(res.diag, diag);
    return 0;
  } catch (const std::exception& e) {
    put_err(err, errlen, e.what());
    return 1;
  }
}

// -----------------------------------------------------------------------------
// CPU-baseline timing helper (bench.py only).  Times the hot path of one sweep
// on a given state: the Neal-8 scan over the observations [0, n_obs) (the cost
// per observation does not depend on the index), then update_phi, one
// split-merge proposal and the log-likelihood, and returns the seconds of each
// piece.  `n_chains` independent copies run on that many threads (one chain per
// core); the returned times are the max over threads.
// out[0]=scan seconds for n_obs observations, out[1]=update_phi, out[2]=split-merge,
// out[3]=loglik, out[4]=observations actually timed
// -----------------------------------------------------------------------------
int orc_time_sweep(const orc_data* d, int m_aux, int t, int rr, int K, const int* c, const double* center,
                   const double* sigma, long pool_size, const double* pool_center, const double* pool_sigma,
                   int n_obs, int do_sm, int n_chains, unsigned long long seed, const orc_opts* o, double* out,
                   char* err, int errlen) {
  try {
    AuxData a = mk_data(d);
    Pool pool = mk_pool(pool_size, d->p, pool_center, pool_sigma);
    Opts op = mk_opts(o);
    std::vector<std::vector<double>> tms(n_chains, std::vector<double>(5, 0.0));
    std::vector<std::string> errs(n_chains);
    auto work = [&](int tid) {
      try {
        State s = mk_state(d->n, d->p, K, c, center, sigma);
        MtRng r(seed + 7919ull * tid);
        Diag dg;
        using clk = std::chrono::steady_clock;
        auto t0 = clk::now();
        int nobs = std::min(n_obs, d->n);
        if (op.counted) {
          // counted mode: same code path as neal8_scan but bounded to n_obs observations
          ScanCtx ctx;
          ctx.counts.assign(s.K, 0);
          for (int i = 0; i < d->n; i++) ctx.counts[s.c[i]]++;
          ctx.cl.num1.resize(s.K);
          ctx.cl.den.resize(s.K);
          for (int k = 0; k < s.K; k++) fill_terms(s.sigma[k], a.attrisize, ctx.cl.num1[k], ctx.cl.den[k]);
          ctx.pool.num1.resize(pool.size());
          ctx.pool.den.resize(pool.size());
          ctx.pool_ready.assign(pool.size(), 0);
          for (int i = 0; i < nobs; i++) sample_allocation(i, a, s, m_aux, pool, r, op, &dg, &ctx);
        } else {
          for (int i = 0; i < nobs; i++) sample_allocation(i, a, s, m_aux, pool, r, op, &dg, nullptr);
        }
        auto t1 = clk::now();
        update_phi(s, a, r, op, {}, &dg);
        auto t2 = clk::now();
        if (do_sm) split_and_merge(s, a, t, rr, r, op, &dg);
        auto t3 = clk::now();
        volatile double ll = compute_loglikelihood(s, a);
        (void)ll;
        auto t4 = clk::now();
        tms[tid][0] = std::chrono::duration<double>(t1 - t0).count();
        tms[tid][1] = std::chrono::duration<double>(t2 - t1).count();
        tms[tid][2] = std::chrono::duration<double>(t3 - t2).count();
        tms[tid][3] = std::chrono::duration<double>(t4 - t3).count();
        tms[tid][4] = nobs;
      } catch (const std::exception& e) {
        errs[tid] = e.what();
      }
    };
    std::vector<std::thread> th;
    for (int i = 1; i < n_chains; i++) th.emplace_back(work, i);
    work(0);
    for (auto& x : th) x.join();
    for (int q = 0; q < 5; q++) out[q] = 0;
    for (int i = 0; i < n_chains; i++) {
      if (!errs[i].empty()) {
        put_err(err, errlen, errs[i].c_str());
        return 1;
      }
      for (int q = 0; q < 5; q++) out[q] = std::max(out[q], tms[i][q]);
    }
    return 0;
  } catch (const std::exception& e) {
    put_err(err, errlen, e.what());
    return 1;
  }
}

}  // extern "C"
