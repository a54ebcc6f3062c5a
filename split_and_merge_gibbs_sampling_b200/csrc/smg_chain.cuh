// smg_chain.cuh -- host-side state of one chain (the internal_state + aux_data of
// code/common_functions.hpp:32-72, resident in HBM) and small helpers.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/smgibbs.h"
#include "smg_kernels.cuh"

namespace smg {

extern thread_local std::string g_last_error;
inline int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

#define SMG_CUDA(call)                                                                                       \
  do {                                                                                                       \
    cudaError_t _e = (call);                                                                                 \
    if (_e != cudaSuccess)                                                                                   \
      return smg::fail(SMG_ERR_CUDA, std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " + __FILE__ + \
                                         ":" + std::to_string(__LINE__));                                    \
  } while (0)

// Stream-ordered device memory from the device's default pool with an unbounded release threshold: chains
// are created and destroyed per run_markov_chain call, and cudaMalloc / cudaFree of multi-GB buffers cost tens
// of milliseconds each time; the pool hands the same memory back in microseconds.
inline cudaError_t dev_pool_init(int device) {
  static bool done[64] = {};
  if (device < 0 || device >= 64 || done[device]) return cudaSuccess;
  cudaMemPool_t pool;
  cudaError_t e = cudaDeviceGetDefaultMemPool(&pool, device);
  if (e != cudaSuccess) return e;
  unsigned long long thr = ~0ull;
  e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  done[device] = (e == cudaSuccess);
  return e;
}
// Host <-> device copies ordered on the CONSUMING stream (the chains' streams are non-blocking: a legacy-stream
// cudaMemcpy from pageable memory may return before its DMA has landed and is not ordered against them), followed by
// a wait so that the host buffer may be reused and the data is visible to every other stream.
inline cudaError_t h2d_sync(void* dst, const void* src, size_t bytes, cudaStream_t st) {
  if (bytes == 0) return cudaSuccess;
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st);
  return e != cudaSuccess ? e : cudaStreamSynchronize(st);
}
inline cudaError_t d2h_sync(void* dst, const void* src, size_t bytes, cudaStream_t st) {
  if (bytes == 0) return cudaSuccess;
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st);
  return e != cudaSuccess ? e : cudaStreamSynchronize(st);
}
template <typename T>
inline cudaError_t dev_malloc(T** ptr, size_t bytes, cudaStream_t st) {
  return cudaMallocAsync((void**)ptr, bytes ? bytes : 1, st);
}

// Philox sub-stream ids inside one iteration
enum SubPhase : uint32_t {
  SUB_SCAN = 0,
  SUB_PHI_AFTER_SCAN = 1,
  SUB_SM_SELECT = 2,
  SUB_SM_PRIOR = 3,       // prior draws of the launch states
  SUB_SM_LAUNCH = 4,      // random launch allocation
  SUB_SM_RG = 8,          // + scan index q  (restricted scans and their update_phi)
  SUB_SM_MERGE = 64,      // + update index
  SUB_SM_ACCEPT = 120,
  SUB_POOL = 121,
  SUB_INIT = 122,
  SUB_INIT_PHI = 123,
  SUB_AUX_FREE = 124      // pool-free auxiliary components of a pass
};

// extra parameter slots (beyond the SMG_MAX_SLOTS scan slots) used by the split-merge step
enum SmSlot : int { SM_SL_A = 0, SM_SL_B, SM_ML_M, SM_ST_A, SM_ST_B, SM_ST_M, SM_TMP0, SM_TMP1, SM_NSLOTS };

struct SmWork;  // split-merge workspace (smg_sm.cuh)

}  // namespace smg

struct smg_chain {
  // ---- configuration (host copies)
  int n = 0, p = 0, pp = 0, mmax = 0, m_aux = 0, L = 1, t = 10, r = 10;
  int neal8 = 0, split_merge = 1, n8_step = 1, sam_step = 1, thinning = 1, sigma_exact = 0, pair_det = 0, aux_mode = 0;
  double gamma = 1.0;
  unsigned long long seed = 0;
  int Kcap = 192, NS = SMG_MAX_SLOTS, NST = SMG_MAX_SLOTS + smg::SM_NSLOTS, ldl = 192;
  long long pool_size = 0;
  int device = 0;
  std::vector<int> h_attr;
  std::vector<double> h_v, h_w;
  // ---- device buffers
  cudaStream_t st = nullptr;
  cudaStream_t st_aux = nullptr;  // side stream: auxiliary-component columns of the NEXT pass, under this pass's scan
  cudaStream_t st_k1 = nullptr;   // side stream: likelihood block of the NEXT pass, under the split-merge proposal
  cudaEvent_t ev_phi_done = nullptr, ev_k1_done = nullptr, ev_aux_go = nullptr;
  cudaEvent_t ev_scan_done = nullptr, ev_aux_done = nullptr;
  cudaEvent_t ev_aux_t0 = nullptr, ev_aux_t1 = nullptr;  // device time of the prefetched aux pass (side stream)
  bool aux_timed = false;
  long long ll_for_iter = -1;      // the LL block holds the columns of the state at the start of this iteration (-1: stale)
  cudaEvent_t ev_k1[2] = {};      // device time of the likelihood-block kernel wherever in the sweep it was launched
  bool k1_timed = false;
  bool k1_overlap = true;        // evaluate the next pass's likelihood block beside the split-merge proposal
  bool k1_in_tail = false;       // the block of the last timed iteration ran inside its [6]-[7] interval (not overlapped)
  bool many = false;              // stepped together with other chains: keep every kernel small (no gang-scheduled grids)
  bool aux_ready = false;         // LLaux / aux_e already hold the columns of iteration aux_iter
  long long aux_iter = -1;
  uint8_t* X = nullptr;
  int* attr = nullptr;
  double *v = nullptr, *w = nullptr;
  uint8_t* cen[2] = {nullptr, nullptr};
  double *sig[2] = {nullptr, nullptr}, *isg[2] = {nullptr, nullptr}, *sden[2] = {nullptr, nullptr};
  double* den = nullptr;
  int* c_hist = nullptr;    // labels the cluster histogram H was last brought up to date with
  bool hist_valid = false;  // H matches c_hist (cleared whenever labels are uploaded from the host)
  int* phi_cnt = nullptr;  // finished parts per destination slot of a split parameter-update job
  int phi_parts = 1;       // CTAs per job in phi_update_kernel
  int cur = 0;
  int *c = nullptr, *K = nullptr, *counts = nullptr, *counts_slot = nullptr, *slot2label = nullptr;
  double *LL = nullptr, *mrg = nullptr;
  // tensor-core likelihood block (smg_lltc.cuh): digit-plane operand, per-cluster integer weight sums and scales
  bool ltc_on = false;
  int ltc_kdp = 0, ltc_sms = 0, ltc_kg = 1, ltc_nstg = 4;
  size_t ltc_smem = 0;
  uint8_t* ltc_B = nullptr;
  long long* ltc_Q = nullptr;
  double *ltc_scale = nullptr, *ltc_cst = nullptr;
  int *ltc_K = nullptr, *ltc_ctr = nullptr;
  double* LLaux[2] = {nullptr, nullptr};  // aux columns, double-buffered: [aux_buf] feeds the current pass
  int* aux_e[2] = {nullptr, nullptr};
  double* aux_sd[2] = {nullptr, nullptr};  // pool-free mode: log-normaliser sums of the auxiliary components
  int aux_buf = 0;
  int* und_blk = nullptr;   // flagged rows per scan block
  uint8_t* und0 = nullptr;  // [n padded] precomputed screen flags of the allocation scan (scan_margin_kernel)
  uint8_t* pcen = nullptr;
  double *psig = nullptr, *pisg = nullptr, *pden = nullptr, *psden = nullptr;
  bool pool_valid = false;
  int* H = nullptr;
  double *partial = nullptr, *loglik_d = nullptr;
  int* status = nullptr;
  int* accepted_d = nullptr;
  unsigned long long* stats_d = nullptr;
  int* scan_job = nullptr;       // mailbox of the scan cluster
  unsigned long long* scan_prof = nullptr;  // cycle counters of the scanner's phases
  int scan_spec = -1;                       // speculative evaluation of the scan: -1 library default, 0 / 1 / 2 (smg_debug_scan_spec)
  double* tape_d = nullptr;      // [n][m_aux+1] injected scan uniforms
  double *uc_d = nullptr, *us_d = nullptr;  // injected phi uniforms [NST][p]
  smg::SmWork* sm = nullptr;
  // ---- host-side run state
  long long iter = 0;
  int h_K = 0, h_accepted = 0;
  double h_loglik = 0.0;
  unsigned long long h_launches = 0, h_sweeps = 0, h_sm_props = 0, h_sm_acc = 0;
  cudaEvent_t ev[8] = {};
  cudaEvent_t ev_call[2] = {};
  double h_step_ms = 0.0;
  double h_timings[8] = {};
  int loglik_blocks = 0;
};

namespace smg {
PhiArgs phi_args_base(smg_chain* ch, uint32_t sub);
inline RngKey mk_key(const smg_chain* ch, uint32_t sub) {
  RngKey k;
  k.k0 = (uint32_t)ch->seed;
  k.k1 = (uint32_t)(ch->seed >> 32);
  k.sweep = (uint32_t)ch->iter;
  k.sub = sub;
  return k;
}
}  // namespace smg
