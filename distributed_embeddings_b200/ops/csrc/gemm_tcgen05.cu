// Hand-written Blackwell GEMM with fused epilogue for the MLP layers (sm_100a only):
//
//   C[M,N] (bf16) = act( A[M,K] (bf16, K-major) x B[N,K]^T (bf16, K-major) + bias[N] )
//
// * operands are streamed by TMA (cp.async.bulk.tensor.2d, 128-byte swizzle) into a multi-stage
//   shared-memory ring guarded by full/empty mbarriers;
// * one elected thread issues tcgen05.mma (UMMA 128 x BLOCK_N x 16, kind::f16, fp32 accumulate)
//   with shared-memory matrix descriptors; the accumulator lives in TMEM (BLOCK_N columns);
// * tcgen05.commit releases smem stages / publishes the finished accumulator through mbarriers;
// * four epilogue warps read their TMEM lane quadrant with tcgen05.ld (32x32b.x32), add the bias,
//   apply ReLU, convert to bf16 and store 64-byte row segments.
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue.
//
// Replaces the cuBLAS GEMM + separate bias/activation ops that TF/XLA runs for the reference's
// Dense layers (examples/dlrm/main.py:123-145).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <mutex>
#include <tuple>

#include "de_b200.h"

namespace de {

namespace {

using bf16 = __nv_bfloat16;

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int kGemmThreads = 192;

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {  // no PTX labels: the loop lives in C++, so any number of inlined copies is fine
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
      "{%3, %4}], [%2];" ::"r"(smem_addr(smem_dst)),
      "l"(map), "r"(smem_addr(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_out) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_addr(smem_out)),
               "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS)
               : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,"
      "%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor: K-major operand tile, rows of 128 bytes, SWIZZLE_128B, 8-row
// groups 1024 bytes apart (tile base 1024-byte aligned).
__device__ __forceinline__ uint64_t make_sw128_kmajor_desc(uint32_t smem_byte_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_byte_addr & 0x3FFFF) >> 4);  // start address  [0,14)
  d |= static_cast<uint64_t>(1) << 16;                          // LBO (unused with swizzle) [16,30)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                  // SBO = 1024 B   [32,46)
  d |= static_cast<uint64_t>(1) << 46;                          // descriptor version (sm_100)
  d |= static_cast<uint64_t>(2) << 61;                          // SWIZZLE_128B   [61,64)
  return d;
}

// Instruction descriptor: D fp32, A/B bf16, both K-major, dense.
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

template <int BLOCK_N, int STAGES>
struct SmemLayout {
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBiasOffset = STAGES * kStageBytes;
  static constexpr int kColsumOffset = kBiasOffset + 2 * BLOCK_N * 4;  // bias tile x2
  static constexpr int kMaxColsum = 2048;                              // EPI 2: N <= 2048
  static constexpr int kBarOffset = kColsumOffset + kMaxColsum * 4;
  static constexpr int kNumBars = 2 * STAGES + 4;                   // full/empty + tmem full/empty x2
  static constexpr int kTotal = kBarOffset + kNumBars * 8 + 16;
};

// Persistent kernel: one CTA per SM walks the output tiles (n fastest so that concurrently
// running CTAs share A rows in L2).  Two TMEM accumulators (2 x BLOCK_N columns) let the epilogue
// of tile i drain while the MMAs of tile i+1 are already running.
// EPI 0: C = A B^T + bias            EPI 1: C = relu(A B^T + bias)
// EPI 2 (backward of a ReLU layer's input): C = (A B^T) * (act > 0), colsum[n] += sum_m C[m, n]
//        i.e. dgrad GEMM + ReLU-backward mask + bias gradient of the layer below in one kernel.
template <int BLOCK_N, int STAGES, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tn_fused_kernel(const __grid_constant__ CUtensorMap tma_a,
                     const __grid_constant__ CUtensorMap tma_b, const bf16* __restrict__ bias,
                     bf16* __restrict__ C, int64_t ldc, int M, int N, int K,
                     const bf16* __restrict__ act, int64_t ldact, float* __restrict__ colsum) {
  using L = SmemLayout<BLOCK_N, STAGES>;
  constexpr int kTmemCols = 2 * BLOCK_N;
  extern __shared__ uint8_t smem_raw[];
  // the swizzled tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  float* s_bias = reinterpret_cast<float*>(smem + L::kBiasOffset);
  float* s_colsum = reinterpret_cast<float*>(smem + L::kColsumOffset);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;
  const int tiles_n = (N + BLOCK_N - 1) / BLOCK_N;
  const int tiles_m = (M + BLOCK_M - 1) / BLOCK_M;
  const int num_tiles = tiles_n * tiles_m;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 128);  // every epilogue thread arrives
    }
    fence_barrier_init();
    fence_proxy_async();
  } else if (warp == 1) {
    tmem_alloc<kTmemCols>(tmem_ptr_smem);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / tiles_n) * BLOCK_M, n0 = (tile % tiles_n) * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
          mbar_expect_tx(&full_bar[stage], L::kStageBytes);
          tma_load_2d(sa, &tma_a, &full_bar[stage], kb * BLOCK_K, m0);
          tma_load_2d(sb, &tma_b, &full_bar[stage], kb * BLOCK_K, n0);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);  // epilogue drained this accumulator
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_addr(smem + stage * L::kStageBytes);
          const uint32_t sb = sa + L::kABytes;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t da = make_sw128_kmajor_desc(sa + k * UMMA_K * 2);
            const uint64_t db = make_sw128_kmajor_desc(sb + k * UMMA_K * 2);
            umma_f16(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // smem stage reusable once these MMAs retire
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full_bar[acc]);  // accumulator complete
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int et = (warp - 2) * 32 + lane;
    if (EPI == 2) {
      for (int i = et; i < L::kMaxColsum; i += 128) s_colsum[i] = 0.f;
    }
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m0 = (tile / tiles_n) * BLOCK_M, n0 = (tile % tiles_n) * BLOCK_N;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      float* sb_tile = s_bias + acc * BLOCK_N;
      if (EPI != 2) {
        for (int i = et; i < BLOCK_N; i += 128)
          sb_tile[i] = (bias != nullptr && n0 + i < N) ? __bfloat162float(bias[n0 + i]) : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");  // epilogue warps only
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tcgen05_fence_after();
      const int row = m0 + quad * 32 + lane;
      bf16* crow = C + static_cast<int64_t>(row) * ldc + n0;
      const bf16* arow = EPI == 2 ? act + static_cast<int64_t>(row) * ldact + n0 : nullptr;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N; c += 32) {
        uint32_t v[32];
        const uint32_t taddr =
            tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BLOCK_N + c;
        tmem_ld_32x32b_x32(taddr, v);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (EPI == 2) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 a4 = make_uint4(0, 0, 0, 0);
            if (row < M && n0 + c + j < N) a4 = *reinterpret_cast<const uint4*>(arow + c + j);
            const uint32_t aw[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float2 af = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&aw[q]));
              f[j + 2 * q] = af.x > 0.f ? f[j + 2 * q] : 0.f;
              f[j + 2 * q + 1] = af.y > 0.f ? f[j + 2 * q + 1] : 0.f;
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            f[j] += sb_tile[c + j];
            if (EPI == 1) f[j] = fmaxf(f[j], 0.f);
          }
        }
        if (row < M) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (n0 + c + j < N) {
              uint32_t packed[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                __nv_bfloat162 h = __floats2bfloat162_rn(f[j + 2 * q], f[j + 2 * q + 1]);
                packed[q] = *reinterpret_cast<uint32_t*>(&h);
              }
              *reinterpret_cast<uint4*>(crow + c + j) =
                  make_uint4(packed[0], packed[1], packed[2], packed[3]);
            }
          }
        }
        if (EPI == 2) {
          // column sums over the 32 rows of this warp: butterfly transpose-reduce, 31 shuffles;
          // afterwards lane l holds the sum of column c + l
#pragma unroll
          for (int sft = 16; sft >= 1; sft >>= 1) {
            const bool up = (lane & sft) != 0;
#pragma unroll
            for (int i = 0; i < sft; ++i) {
              const float send = up ? f[i] : f[i + sft];
              const float recv = __shfl_xor_sync(0xffffffffu, send, sft);
              f[i] = (up ? f[i + sft] : f[i]) + recv;
            }
          }
          if (n0 + c + lane < N) atomicAdd(&s_colsum[n0 + c + lane], f[0]);
        }
      }
      // hand the accumulator back to the MMA warp
      tcgen05_fence_before();
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(
                       smem_addr(&tmem_empty_bar[acc]))
                   : "memory");
    }
    if (EPI == 2) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int i = et; i < N && i < L::kMaxColsum; i += 128) {
        const float vsum = s_colsum[i];
        if (vsum != 0.f) atomicAdd(colsum + i, vsum);
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------ CTA-pair (cta_group::2) variant
// Validated on B200 (tests/test_gemm_tcgen05.py); 0.72-0.98 of cuBLASLt on the wide DLRM layers.
// Two CTAs of a cluster (one TPC) cooperate on a 256 x BLOCK_N tile: CTA r stages rows
// [r*128, r*128+128) of A and rows [r*BLOCK_N/2, ...) of B, so every operand byte is loaded once
// per pair (half the smem fill traffic per SM); the leader CTA issues tcgen05.mma.cta_group::2
// (UMMA 256 x BLOCK_N x 16) and each CTA's TMEM receives its 128 accumulator rows.
//   * both producers' TMA loads (.cta_group::2) complete on the LEADER's full barrier;
//   * tcgen05.commit ... multicast::cluster frees the smem stage / publishes the accumulator in
//     both CTAs;
//   * the epilogue threads of both CTAs hand the accumulator back on the leader's tmem_empty
//     barrier (remote mbarrier.arrive through mapa).
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP_CL:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_CL;\n"
      "bra WAIT_LOOP_CL;\n"
      "DONE_CL:\n"
      "}\n" ::"r"(smem_addr(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map,
                                                 uint32_t leader_bar_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];" ::"r"(smem_addr(smem_dst)),
      "l"(map), "r"(leader_bar_addr), "r"(c0), "r"(c1)
      : "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_out) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_addr(smem_out)),
               "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS)
               : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the same-offset mbarrier of both CTAs once the issued MMAs have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;" ::"r"(smem_addr(bar)),
      "h"(mask)
      : "memory");
}

template <int BLOCK_N, int STAGES>
struct PairSmemLayout {
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;          // this CTA's 128 rows of A
  static constexpr int kBBytes = (BLOCK_N / 2) * BLOCK_K * 2;    // this CTA's half of the B tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBiasOffset = STAGES * kStageBytes;
  static constexpr int kBarOffset = kBiasOffset + 2 * BLOCK_N * 4;
  static constexpr int kNumBars = 2 * STAGES + 4;
  static constexpr int kTotal = kBarOffset + kNumBars * 8 + 16;
};

// EPI 0: C = A B^T + bias, EPI 1: relu(...)
template <int BLOCK_N, int STAGES, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_tn_pair_kernel(const __grid_constant__ CUtensorMap tma_a,
                    const __grid_constant__ CUtensorMap tma_b, const bf16* __restrict__ bias,
                    bf16* __restrict__ C, int64_t ldc, int M, int N, int K) {
  using L = PairSmemLayout<BLOCK_N, STAGES>;
  constexpr int kTmemCols = 2 * BLOCK_N;
  static_assert(kTmemCols <= 512, "two accumulators must fit the 512 TMEM columns");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  float* s_bias = reinterpret_cast<float*>(smem + L::kBiasOffset);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;  // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;  // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;
  const int tiles_n = (N + BLOCK_N - 1) / BLOCK_N;
  const int tiles_m = (M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const int num_tiles = tiles_n * tiles_m;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);   // the leader's expect_tx arrive (the peer's copy is unused)
      mbar_init(&empty_bar[s], 1);  // multicast commit of the leader's MMA thread
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 256);  // epilogue threads of both CTAs (leader's copy is used)
    }
    fence_barrier_init();
    fence_proxy_async();
  } else if (warp == 1) {
    tmem_alloc_pair<kTmemCols>(tmem_ptr_smem);
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();  // peer barriers are initialised before anything signals them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int m0 = (tile / tiles_n) * (2 * BLOCK_M) + static_cast<int>(cta_rank) * BLOCK_M;
        const int n0 = (tile % tiles_n) * BLOCK_N + static_cast<int>(cta_rank) * (BLOCK_N / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait_cluster(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
          const uint32_t leader_full = mapa_shared(smem_addr(&full_bar[stage]), 0);
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * L::kStageBytes);
          tma_load_2d_pair(sa, &tma_a, leader_full, kb * BLOCK_K, m0);
          tma_load_2d_pair(sb, &tma_b, leader_full, kb * BLOCK_K, n0);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc(2 * BLOCK_M, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait_cluster(&tmem_empty_bar[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait_cluster(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_addr(smem + stage * L::kStageBytes);
          const uint32_t sb = sa + L::kABytes;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t da = make_sw128_kmajor_desc(sa + k * UMMA_K * 2);
            const uint64_t db = make_sw128_kmajor_desc(sb + k * UMMA_K * 2);
            umma_f16_pair(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit_pair(&empty_bar[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_pair(&tmem_full_bar[acc]);
      }
    }
  } else {
    // ===================== epilogue (warps 2..5 of both CTAs) =====================
    const int quad = warp & 3;
    const int et = (warp - 2) * 32 + lane;
    int it = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
      const int m0 = (tile / tiles_n) * (2 * BLOCK_M) + static_cast<int>(cta_rank) * BLOCK_M;
      const int n0 = (tile % tiles_n) * BLOCK_N;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      float* sb_tile = s_bias + acc * BLOCK_N;
      for (int i = et; i < BLOCK_N; i += 128)
        sb_tile[i] = (bias != nullptr && n0 + i < N) ? __bfloat162float(bias[n0 + i]) : 0.f;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait_cluster(&tmem_full_bar[acc], acc_phase);
      tcgen05_fence_after();
      const int row = m0 + quad * 32 + lane;
      bf16* crow = C + static_cast<int64_t>(row) * ldc + n0;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N; c += 32) {
        uint32_t v[32];
        const uint32_t taddr =
            tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BLOCK_N + c;
        tmem_ld_32x32b_x32(taddr, v);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          f[j] = __uint_as_float(v[j]) + sb_tile[c + j];
          if (EPI == 1) f[j] = fmaxf(f[j], 0.f);
        }
        if (row < M) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (n0 + c + j < N) {
              uint32_t packed[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                __nv_bfloat162 h = __floats2bfloat162_rn(f[j + 2 * q], f[j + 2 * q + 1]);
                packed[q] = *reinterpret_cast<uint32_t*>(&h);
              }
              *reinterpret_cast<uint4*>(crow + c + j) =
                  make_uint4(packed[0], packed[1], packed[2], packed[3]);
            }
          }
        }
      }
      // hand the accumulator back: the leader's MMA thread waits for both CTAs' epilogues
      tcgen05_fence_before();
      mbar_arrive_remote(mapa_shared(smem_addr(&tmem_empty_bar[acc]), 0));
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();  // the leader's MMAs read the peer's smem: nobody leaves early
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc_pair<kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------ host side: tensor maps
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                              const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                              const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode_fn() {
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(p);
  });
  return fn;
}

// 2-D bf16 row-major [rows, cols] tensor, box = [box_rows, 64 cols], 128-byte swizzle
bool make_tensor_map(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols,
                     int64_t row_stride_elems, int box_rows) {
  EncodeFn fn = get_encode_fn();
  if (fn == nullptr) return false;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(row_stride_elems) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(BLOCK_K), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int BLOCK_N, int STAGES, int EPI>
bool launch_one(const CUtensorMap& ta, const CUtensorMap& tb, const void* bias, void* C,
                int64_t ldc, int M, int N, int K, const void* act, int64_t ldact, float* colsum,
                int sm_count, cudaStream_t stream) {
  using L = SmemLayout<BLOCK_N, STAGES>;
  const size_t smem = L::kTotal + 1024;
  const int tiles = ((N + BLOCK_N - 1) / BLOCK_N) * ((M + BLOCK_M - 1) / BLOCK_M);
  dim3 grid(tiles < sm_count ? tiles : sm_count);
  cudaFuncSetAttribute(gemm_tn_fused_kernel<BLOCK_N, STAGES, EPI>,
                       cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  gemm_tn_fused_kernel<BLOCK_N, STAGES, EPI><<<grid, kGemmThreads, smem, stream>>>(
      ta, tb, reinterpret_cast<const bf16*>(bias), reinterpret_cast<bf16*>(C), ldc, M, N, K,
      reinterpret_cast<const bf16*>(act), ldact, colsum);
  return cudaGetLastError() == cudaSuccess;
}

template <int BLOCK_N, int STAGES>
bool launch_cfg(const CUtensorMap& ta, const CUtensorMap& tb, const void* bias, void* C,
                int64_t ldc, int M, int N, int K, int epi, const void* act, int64_t ldact,
                float* colsum, int sm_count, cudaStream_t stream) {
  if (epi == 0)
    return launch_one<BLOCK_N, STAGES, 0>(ta, tb, bias, C, ldc, M, N, K, act, ldact, colsum,
                                          sm_count, stream);
  if (epi == 1)
    return launch_one<BLOCK_N, STAGES, 1>(ta, tb, bias, C, ldc, M, N, K, act, ldact, colsum,
                                          sm_count, stream);
  return launch_one<BLOCK_N, STAGES, 2>(ta, tb, bias, C, ldc, M, N, K, act, ldact, colsum,
                                        sm_count, stream);
}

}  // namespace

// C = epilogue(A B^T). A [M,K] (lda), B [N,K] (ldb), C [M,N] (ldc): bf16, 16-byte aligned rows.
// epi 0: + bias; 1: relu(+ bias); 2: * (act > 0) and colsum[n] += column sums (N <= 2048).
bool launch_gemm_tn_fused(const void* A, int64_t lda, const void* B, int64_t ldb, const void* bias,
                          void* C, int64_t ldc, int M, int N, int K, int epi, const void* act,
                          int64_t ldact, float* colsum, int block_n, int sm_count,
                          cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return true;
  if ((lda % 8) || (ldb % 8) || (ldc % 8) || (N % 8)) return false;
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) |
       reinterpret_cast<uintptr_t>(C)) & 15)
    return false;
  if (epi == 2 && (act == nullptr || colsum == nullptr || N > 2048 || (ldact % 8) ||
                   (reinterpret_cast<uintptr_t>(act) & 15)))
    return false;
  const int bn = (block_n == 128 || block_n == 256) ? block_n : (N >= 256 ? 256 : 128);
  alignas(64) CUtensorMap ta, tb;
  if (!make_tensor_map(&ta, A, M, K, lda, BLOCK_M)) return false;
  if (!make_tensor_map(&tb, B, N, K, ldb, bn)) return false;
  if (bn == 256)
    return launch_cfg<256, 4>(ta, tb, bias, C, ldc, M, N, K, epi, act, ldact, colsum, sm_count,
                              stream);
  return launch_cfg<128, 6>(ta, tb, bias, C, ldc, M, N, K, epi, act, ldact, colsum, sm_count,
                            stream);
}

namespace {
template <int EPI>
bool launch_pair(const CUtensorMap& ta, const CUtensorMap& tb, const void* bias, void* C,
                 int64_t ldc, int M, int N, int K, int sm_count, cudaStream_t stream) {
  constexpr int BN = 256, ST = 6;
  using L = PairSmemLayout<BN, ST>;
  const size_t smem = L::kTotal + 1024;
  const int tiles = ((N + BN - 1) / BN) * ((M + 2 * BLOCK_M - 1) / (2 * BLOCK_M));
  int pairs = sm_count / 2;
  if (tiles < pairs) pairs = tiles;
  if (pairs < 1) return false;
  cudaFuncSetAttribute(gemm_tn_pair_kernel<BN, ST, EPI>,
                       cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  gemm_tn_pair_kernel<BN, ST, EPI><<<dim3(2 * pairs), kGemmThreads, smem, stream>>>(
      ta, tb, reinterpret_cast<const bf16*>(bias), reinterpret_cast<bf16*>(C), ldc, M, N, K);
  return cudaGetLastError() == cudaSuccess;
}
}  // namespace

// CTA-pair kernel (cta_group::2, 256 x 256 tile per pair); epi 0 / 1 only.
bool launch_gemm_tn_pair(const void* A, int64_t lda, const void* B, int64_t ldb, const void* bias,
                         void* C, int64_t ldc, int M, int N, int K, bool relu, int sm_count,
                         cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return true;
  if ((lda % 8) || (ldb % 8) || (ldc % 8) || (N % 8)) return false;
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) |
       reinterpret_cast<uintptr_t>(C)) & 15)
    return false;
  alignas(64) CUtensorMap ta, tb;
  if (!make_tensor_map(&ta, A, M, K, lda, BLOCK_M)) return false;
  if (!make_tensor_map(&tb, B, N, K, ldb, 128)) return false;
  return relu ? launch_pair<1>(ta, tb, bias, C, ldc, M, N, K, sm_count, stream)
              : launch_pair<0>(ta, tb, bias, C, ldc, M, N, K, sm_count, stream);
}

bool launch_gemm_tn_bias_act(const void* A, int64_t lda, const void* B, int64_t ldb,
                             const void* bias, void* C, int64_t ldc, int M, int N, int K,
                             bool relu, int block_n, int sm_count, cudaStream_t stream) {
  // block_n == 512 selects the CTA-pair (cta_group::2) kernel
  if (block_n == 512)
    return launch_gemm_tn_pair(A, lda, B, ldb, bias, C, ldc, M, N, K, relu, sm_count, stream);
  return launch_gemm_tn_fused(A, lda, B, ldb, bias, C, ldc, M, N, K, relu ? 1 : 0, nullptr, 0,
                              nullptr, block_n, sm_count, stream);
}

}  // namespace de
