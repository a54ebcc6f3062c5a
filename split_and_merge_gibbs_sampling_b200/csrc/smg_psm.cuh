// smg_psm.cuh -- posterior similarity matrix on the 5th-generation tensor cores (tcgen05 + TMEM).
//
//   PSM[i][j] += sum_t [c_t(i) == c_t(j)]  =  (Z Z^T)[i][j],   Z = [Z_1 | Z_2 | ... | Z_T],  Z_t one-hot n x K_t
//
// The reference leaves this to R (mcclust.ext::comp.psm on results$c_i, realdata_analysis/zoo_simulator.R:339);
// here it is the one dense contraction of the path and runs as an exact integer GEMM: u8 x u8 -> s32
// (tcgen05.mma.kind::i8), accumulators in tensor memory, so the counts are bit-exact.
//
// One CTA owns a 128 x 256 tile of the matrix (D: 128 TMEM lanes x 256 columns of s32).  The operand tiles are
// never read from memory: for every buffered sweep t the CTA *generates* them in shared memory from the label
// arrays -- 128 + 256 rows of KP bytes that are zero except for a single 1 at column c_t(row) -- directly in the
// canonical K-major no-swizzle UMMA layout (8-row x 16-byte core matrices).  A stage is recycled by clearing
// just the bytes that were set.  Warp-specialised: 8 producer warps write the tiles of sweep t into stage t % NST and
// arrive on the stage's `full` mbarrier; one thread of a ninth warp waits for it, issues KP/32 MMAs (M=128, N=256,
// K=32) and tcgen05.commit's them to the stage's `free` mbarrier -- no block barrier in the main loop.
// Measured, n=2e4, KP=64: 2.6 POP/s (dense u8) at 256 buffered sweeps per flush, 3.1 at 1024; the same MMAs issued
// back to back without tile generation (SMG_PSM_MMA_ONLY=1) reach 4.2.
// Algorithmic work per flush: 2 * n^2 * KP * T integer ops; traffic: T * n label bytes in, n^2 * 4 bytes read+written once.
//
// Distributed form (smg_chains_psm_distribute): the rows of the matrix are dealt to the G ranks of a communicator in
// blocks of n / G and every rank's matrix is mapped into every other rank's address space (CUDA IPC over NVLink).  The
// epilogue then ADDS each tile straight into the memory of the rank that owns its rows -- red.global.add.s32, one
// coalesced 128-byte row segment per warp instruction -- so the accumulation and the reduce-scatter over the GPUs are
// one kernel: nothing is left to reduce at the end of the run.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace smg {

#define PSM_M 128
#define PSM_N 256
#define PSM_PRODUCERS 256                 // 8 warps write operand tiles (and run the epilogue)
#define PSM_THREADS (PSM_PRODUCERS + 32)   // + one warp whose first lane issues the MMAs

__device__ __forceinline__ uint32_t psm_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void psm_mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(psm_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void psm_mbar_wait(uint64_t* bar, unsigned parity) {
  const uint32_t a = psm_smem_u32(bar);
  unsigned ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
  }
}
// all previously issued tcgen05.mma of this thread arrive on `bar` when they have completed
__device__ __forceinline__ void psm_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(psm_smem_u32(bar))
               : "memory");
}

// Shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp):
// start address >> 4 in bits [0,14), leading byte offset >> 4 in [16,30) (stride between the two 16-byte
// K-chunks of one MMA), stride byte offset >> 4 in [32,46) (stride between 8-row groups), version 1 in [46,48).
__device__ __forceinline__ uint64_t psm_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// D[tmem] (+)= A[smem] * B[smem]^T, u8 x u8 -> s32
__device__ __forceinline__ void psm_mma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}

// byte offset of (row, k) inside an operand tile of `KP` bytes per row:
// [row / 8][k / 16][row % 8][k % 16]  =>  SBO = KP * 8 bytes, LBO = 128 bytes
template <int KP>
__device__ __forceinline__ uint32_t psm_tile_off(int row, int k) {
  return (uint32_t)((row >> 3) * (KP * 8) + (k >> 4) * 128 + (row & 7) * 16 + (k & 15));
}

template <int KP, int NST>
__global__ void __launch_bounds__(PSM_THREADS, 2)
    psm_accumulate_kernel(const uint8_t* __restrict__ labels, int n, int T, int* __restrict__ psm, int mma_only,
                          int upper_only, int* const* __restrict__ peers, int rows_per) {
  constexpr int A_BYTES = PSM_M * KP, B_BYTES = PSM_N * KP, STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t s_free[NST];
  __shared__ __align__(8) uint64_t s_full[NST];
  __shared__ __align__(8) uint64_t s_done;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * PSM_M, n0 = blockIdx.x * PSM_N;
  // The matrix is symmetric: tiles that lie entirely below the diagonal are not accumulated (half the MMAs and half the
  // read-modify-write traffic of a flush); psm_mirror_kernel copies the upper triangle down once, when the matrix is read.
  if (upper_only && n0 + PSM_N - 1 < m0) return;

  // zero every stage once; afterwards a stage is recycled by clearing the bytes that were set
  for (int q = tid; q < NST * STAGE_BYTES / 16; q += PSM_THREADS) reinterpret_cast<uint4*>(smem)[q] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int s = 0; s < NST; s++) {
      psm_mbar_init(&s_free[s], 1);
      psm_mbar_init(&s_full[s], PSM_PRODUCERS);
    }
    psm_mbar_init(&s_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {  // 256 TMEM columns: the s32 accumulator tile
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(psm_smem_u32(&s_tmem)), "r"(256u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = s_tmem;

  // instruction descriptor (cute::UMMA::InstrDescriptor): D = s32, A = B = u8, both K-major, N = 256, M = 128
  constexpr uint32_t IDESC = (2u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(PSM_N >> 3) << 17) | ((uint32_t)(PSM_M >> 4) << 24);
  const uint32_t smem_base = psm_smem_u32(smem);
  const int rowA = m0 + tid, rowB = n0 + tid;  // tid < 128 also writes a row of A
  const bool hasA = (tid < PSM_M) && (rowA < n), hasB = rowB < n;
  int oldA[NST], oldB[NST];
#pragma unroll
  for (int s = 0; s < NST; s++) oldA[s] = oldB[s] = -1;

  if (mma_only) {
    // measurement mode (SMG_PSM_MMA_ONLY=1, results are meaningless): the same MMAs on stage 0 back to back, no tile
    // generation and no barriers -- the rate the tensor pipe sustains for this instruction shape
    if (tid == 0) {
      for (int t = 0; t < T; t++) {
#pragma unroll
        for (int ks = 0; ks < KP / 32; ks++)
          psm_mma_i8(tmem_d, psm_desc(smem_base + ks * 256, 128, KP * 8), psm_desc(smem_base + A_BYTES + ks * 256, 128, KP * 8),
                     IDESC, (t > 0 || ks > 0) ? 1u : 0u);
      }
      psm_commit(&s_done);
    }
    T = (T > 0) ? T : 0;
  } else if (tid < PSM_PRODUCERS) {
    // ===== producers (8 warps): write the one-hot operand tiles of sweep t into stage t % NST =====
    // labels are fetched one group of NST sweeps ahead: their L2/HBM latency overlaps the stores of the current group
    int nxA[NST], nxB[NST];
#pragma unroll
    for (int s = 0; s < NST; s++) {
      nxA[s] = (hasA && s < T) ? (int)labels[(size_t)s * n + rowA] : -1;
      nxB[s] = (hasB && s < T) ? (int)labels[(size_t)s * n + rowB] : -1;
    }
    for (int t0 = 0; t0 < T; t0 += NST) {
      int curA[NST], curB[NST];
#pragma unroll
      for (int s = 0; s < NST; s++) {
        curA[s] = nxA[s];
        curB[s] = nxB[s];
        const int tn = t0 + NST + s;
        nxA[s] = (hasA && tn < T) ? (int)labels[(size_t)tn * n + rowA] : -1;
        nxB[s] = (hasB && tn < T) ? (int)labels[(size_t)tn * n + rowB] : -1;
      }
#pragma unroll
      for (int s = 0; s < NST; s++) {
        const int t = t0 + s;
        if (t >= T) break;
        uint8_t* stA = smem + s * STAGE_BYTES;
        uint8_t* stB = stA + A_BYTES;
        const int cA = curA[s], cB = curB[s];
        if (t >= NST) psm_mbar_wait(&s_free[s], (unsigned)(((t / NST) - 1) & 1));  // the MMAs that read this stage are done
        if (oldA[s] >= 0) stA[psm_tile_off<KP>(tid, oldA[s])] = 0;
        if (oldB[s] >= 0) stB[psm_tile_off<KP>(tid, oldB[s])] = 0;
        if (cA >= 0 && cA < KP) stA[psm_tile_off<KP>(tid, cA)] = 1;
        if (cB >= 0 && cB < KP) stB[psm_tile_off<KP>(tid, cB)] = 1;
        oldA[s] = (cA < KP) ? cA : -1;
        oldB[s] = (cB < KP) ? cB : -1;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(psm_smem_u32(&s_full[s])) : "memory");
      }
    }
  } else if (tid == PSM_PRODUCERS) {
    // ===== MMA issuer (one thread of the ninth warp): no block barrier anywhere in the main loop =====
    for (int t = 0; t < T; t++) {
      const int s = t % NST;
      psm_mbar_wait(&s_full[s], (unsigned)((t / NST) & 1));  // all 256 producers have written this stage
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a0 = smem_base + s * STAGE_BYTES, b0 = a0 + A_BYTES;
#pragma unroll
      for (int ks = 0; ks < KP / 32; ks++) {
        const uint64_t da = psm_desc(a0 + ks * 256, 128, KP * 8);
        const uint64_t db = psm_desc(b0 + ks * 256, 128, KP * 8);
        psm_mma_i8(tmem_d, da, db, IDESC, (t > 0 || ks > 0) ? 1u : 0u);
      }
      psm_commit(&s_free[s]);
    }
    psm_commit(&s_done);
  }
  // ---- epilogue: TMEM -> registers -> PSM += (each element of the matrix is owned by exactly one CTA)
  if (T > 0 && tid < PSM_PRODUCERS) {
    psm_mbar_wait(&s_done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3, h = warp >> 2;  // TMEM lanes [32q, 32q+32), columns [128h, 128h+128)
    const int row = m0 + 32 * q + lane;
#pragma unroll 1
    for (int cb = 0; cb < 4; cb++) {
      const int col0 = 128 * h + 32 * cb;
      uint32_t v[32];
      const uint32_t taddr = tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)col0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
            "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
            "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (peers) {
        // distributed matrix: through a per-warp 32 x 33 staging tile (the operand stages are free by now) so that a warp
        // instruction adds 32 consecutive columns of ONE row -- a single 128-byte segment of the owner's memory
        int* tile = reinterpret_cast<int*>(smem) + warp * (32 * 33);
#pragma unroll
        for (int e = 0; e < 32; e++) tile[lane * 33 + e] = (int)v[e];
        __syncwarp();
        const int col = n0 + col0 + lane;
        for (int r = 0; r < 32; r++) {
          const int rr = m0 + 32 * q + r;
          const int val = tile[r * 33 + lane];
          if (rr < n && col < n && val) atomicAdd(peers[rr / rows_per] + (size_t)rr * n + col, val);
        }
        __syncwarp();
      } else if (row < n) {
        int* out = psm + (size_t)row * n + n0 + col0;
        if (n0 + col0 + 32 <= n && (n & 3) == 0) {
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            int4 o = *reinterpret_cast<int4*>(out + e);
            o.x += (int)v[e];
            o.y += (int)v[e + 1];
            o.z += (int)v[e + 2];
            o.w += (int)v[e + 3];
            *reinterpret_cast<int4*>(out + e) = o;
          }
        } else {
#pragma unroll
          for (int e = 0; e < 32; e++)
            if (n0 + col0 + e < n) out[e] += (int)v[e];
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(256u) : "memory");
}

// int32 labels of a chain -> one row of the u8 label buffer
__global__ void psm_pack_labels_kernel(const int* __restrict__ c, int n, uint8_t* __restrict__ dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (uint8_t)c[i];
}

// plain CUDA-core reference of the same accumulation (parity check of the tensor-core kernel on the device)
__global__ void psm_reference_kernel(const uint8_t* __restrict__ labels, int n, int T, int* __restrict__ psm) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= n || i >= n) return;
  int acc = 0;
  for (int t = 0; t < T; t++) acc += (labels[(size_t)t * n + i] == labels[(size_t)t * n + j]);
  psm[(size_t)i * n + j] += acc;
}

// lower triangle <- upper triangle (32 x 32 tiles through shared memory, coalesced both ways); one CTA per tile pair
__global__ void __launch_bounds__(256) psm_mirror_kernel(int* __restrict__ psm, int n) {
  __shared__ int tile[32][33];
  const int bj = blockIdx.x, bi = blockIdx.y;  // source tile: rows bi*32.., cols bj*32.. with bj >= bi
  if (bj < bi) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int i = bi * 32 + r, j = bj * 32 + tx;
    tile[r][tx] = (i < n && j < n) ? psm[(size_t)i * n + j] : 0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int i = bj * 32 + r, j = bi * 32 + tx;  // destination (i, j) = source (j, i)
    if (i < n && j < n && i > j) psm[(size_t)i * n + j] = tile[tx][r];
  }
}

// distributed matrix: lower triangle of THIS rank's rows <- upper triangle, read from the rank that owns the source row
__global__ void __launch_bounds__(256) psm_mirror_dist_kernel(int* const* __restrict__ peers, int n, int rows_per, int rank) {
  __shared__ int tile[32][33];
  const int bj = blockIdx.x, bi = blockIdx.y;  // source tile: rows bi*32.. (j), cols bj*32.. (i), bj >= bi
  if (bj < bi) return;
  const int i_lo = bj * 32, i_hi = min(n, i_lo + 32) - 1;  // destination rows
  if (i_hi / rows_per < rank || i_lo / rows_per > rank) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int j = bi * 32 + r, i = bj * 32 + tx;
    tile[r][tx] = (j < n && i < n) ? __ldcg(peers[j / rows_per] + (size_t)j * n + i) : 0;
  }
  __syncthreads();
  int* mine = peers[rank];
  for (int r = ty; r < 32; r += 8) {
    const int i = bj * 32 + r, j = bi * 32 + tx;
    if (i < n && j < n && i > j && i / rows_per == rank) mine[(size_t)i * n + j] = tile[tx][r];
  }
}

}  // namespace smg
