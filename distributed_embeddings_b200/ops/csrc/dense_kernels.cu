// Dense-side kernels of the DLRM step for sm_100a (everything around the MLP GEMMs):
//   * dot interaction forward / backward: per sample Gram matrix F F^T of the (n_emb + 1) x D
//     feature matrix on tensor cores (mma.sync m16n8k16 bf16, one warp per sample).  The op is
//     HBM bound (7 KB in, 1 KB out per sample), the MMA only has to keep up with the loads.  The
//     backward writes the embedding gradient straight into the (symmetric) gradient buffer of
//     the embedding engine, so no extra pack / copy precedes the backward all-to-all.
//   * fused ReLU-backward + bias gradient, fused final layer + BCE loss + their backward,
//     fused SGD update of the fp32 master weights + bf16 shadow copy, input cast/pad.
//
// Capability parity: dot_interact (reference examples/dlrm/utils.py:92-113) and the TF / XLA
// elementwise + optimizer kernels the reference borrows.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>

#include "common.cuh"
#include "de_b200.h"

namespace de {

namespace {

using bf16 = __nv_bfloat16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
      "{%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int kMaxFeat = 32;     // rows of F incl. the bottom-MLP vector, padded to 32
constexpr int kWarps = 4;        // samples in flight per block

__host__ __device__ constexpr int fwd_warp_bytes(int d) {
  return (kMaxFeat * (d + 8) * 2 > kMaxFeat * 33 * 4) ? kMaxFeat * (d + 8) * 2 : kMaxFeat * 33 * 4;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// Stage F = [bottom ; emb_0 .. emb_{n-1}] (each D bf16) of one sample into smem [32][LD] with
// cp.async (LDGSTS): every 16-byte chunk of the sample is in flight before the first wait.
// Pad rows (> n_emb) are zeroed once by the caller.
template <int D>
__device__ __forceinline__ void stage_features(bf16* sF, int LD, const bf16* bottom,
                                               const bf16* emb, int n_emb, int lane) {
  constexpr int kChunks = D / 8;  // 16-byte chunks per row
  const int total = (n_emb + 1) * kChunks;
#pragma unroll 4
  for (int c = lane; c < total; c += 32) {
    const int row = c / kChunks, ch = c - row * kChunks;
    const bf16* src = row == 0 ? bottom + ch * 8 : emb + (row - 1) * D + ch * 8;
    cp_async16(sF + row * LD + ch * 8, src);
  }
  cp_async_wait_all();
}

template <int D>
__device__ __forceinline__ void zero_pad_rows(bf16* sF, int LD, int n_emb, int lane) {
  constexpr int kChunks = D / 8;
  for (int c = (n_emb + 1) * kChunks + lane; c < kMaxFeat * kChunks; c += 32) {
    const int row = c / kChunks, ch = c - row * kChunks;
    *reinterpret_cast<uint4*>(sF + row * LD + ch * 8) = make_uint4(0, 0, 0, 0);
  }
}

// z[s] = [ tril(F F^T, -1) (row major) | bottom | 0 pad ]
template <int D>
__global__ void __launch_bounds__(kWarps * 32)
interact_fwd_kernel(const bf16* __restrict__ bottom, int64_t bottom_stride,
                    const bf16* __restrict__ emb, int64_t emb_stride, int n_emb,
                    bf16* __restrict__ z, int64_t z_stride, int z_width, int64_t batch,
                    const __grid_constant__ SyncArgs sync) {
  sync_head(sync);  // every owner's pooled rows have landed in this rank's embedding output
  constexpr int LD = D + 8;
  constexpr int kWarpBytes = fwd_warp_bytes(D);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  bf16* sF = reinterpret_cast<bf16*>(smem_raw + warp * kWarpBytes);
  float* sC = reinterpret_cast<float*>(sF);  // reused after the MMAs: [32][33] fp32 (4.2 KB)
  const int nf = n_emb + 1;
  const int n_inter = nf * (nf - 1) / 2;

  for (int64_t s = static_cast<int64_t>(blockIdx.x) * kWarps + warp; s < batch;
       s += static_cast<int64_t>(gridDim.x) * kWarps) {
    const bf16* bp = bottom + s * bottom_stride;
    zero_pad_rows<D>(sF, LD, n_emb, lane);  // the C staging below may alias the pad rows
    stage_features<D>(sF, LD, bp, emb + s * emb_stride, n_emb, lane);
    __syncwarp();
    // lower triangle tiles: m-tile 0 x n-tiles {0,1}; m-tile 1 x n-tiles {0..3}
    float acc0[2][4] = {}, acc1[4][4] = {};
#pragma unroll
    for (int k = 0; k < D; k += 16) {
      uint32_t a0[4], a1[4], b01[4], b23[4];
      const int arow = lane & 15, acol = k + ((lane >> 4) << 3);
      ldmatrix_x4(a0, smem_u32(sF + arow * LD + acol));
      ldmatrix_x4(a1, smem_u32(sF + (16 + arow) * LD + acol));
      const int brow = (lane & 7) + ((lane >> 4) << 3), bcol = k + (((lane >> 3) & 1) << 3);
      ldmatrix_x4(b01, smem_u32(sF + brow * LD + bcol));         // n-tiles 0,1 (rows 0..15)
      ldmatrix_x4(b23, smem_u32(sF + (16 + brow) * LD + bcol));  // n-tiles 2,3 (rows 16..31)
      mma_bf16(acc0[0], a0, b01[0], b01[1]);
      mma_bf16(acc0[1], a0, b01[2], b01[3]);
      mma_bf16(acc1[0], a1, b01[0], b01[1]);
      mma_bf16(acc1[1], a1, b01[2], b01[3]);
      mma_bf16(acc1[2], a1, b23[0], b23[1]);
      mma_bf16(acc1[3], a1, b23[2], b23[3]);
    }
    __syncwarp();
    const int cr = lane >> 2, cc = (lane & 3) << 1;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      sC[cr * 33 + nt * 8 + cc] = acc0[nt][0];
      sC[cr * 33 + nt * 8 + cc + 1] = acc0[nt][1];
      sC[(cr + 8) * 33 + nt * 8 + cc] = acc0[nt][2];
      sC[(cr + 8) * 33 + nt * 8 + cc + 1] = acc0[nt][3];
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      sC[(16 + cr) * 33 + nt * 8 + cc] = acc1[nt][0];
      sC[(16 + cr) * 33 + nt * 8 + cc + 1] = acc1[nt][1];
      sC[(24 + cr) * 33 + nt * 8 + cc] = acc1[nt][2];
      sC[(24 + cr) * 33 + nt * 8 + cc + 1] = acc1[nt][3];
    }
    __syncwarp();
    bf16* zp = z + s * z_stride;
    // strict lower triangle in row-major order: idx = i*(i-1)/2 + j
    for (int idx = lane; idx < n_inter; idx += 32) {
      int i = static_cast<int>((1.0f + sqrtf(1.0f + 8.0f * idx)) * 0.5f);
      while (i * (i - 1) / 2 > idx) --i;
      while ((i + 1) * i / 2 <= idx) ++i;
      const int j = idx - i * (i - 1) / 2;
      zp[idx] = __float2bfloat16_rn(sC[i * 33 + j]);
    }
    for (int c = lane; c < D; c += 32) zp[n_inter + c] = bp[c];
    for (int c = n_inter + D + lane; c < z_width; c += 32) zp[c] = __float2bfloat16_rn(0.f);
    __syncwarp();
  }
  sync_tail(sync);
}

// dF = G F with G symmetric (G_ij = dz[idx(i,j)], zero diagonal); row 0 (+ the direct copy path)
// is the gradient of the bottom-MLP output, rows 1.. go to the embedding gradient buffer.
template <int D>
__global__ void __launch_bounds__(kWarps * 32)
interact_bwd_kernel(const bf16* __restrict__ bottom, int64_t bottom_stride,
                    const bf16* __restrict__ emb, int64_t emb_stride, int n_emb,
                    const bf16* __restrict__ dz, int64_t dz_stride, bf16* __restrict__ dbottom,
                    int64_t dbottom_stride, bf16* __restrict__ demb, int64_t demb_stride,
                    float emb_grad_scale, int64_t batch) {
  constexpr int LD = D + 8;
  constexpr int LDG = 40;  // G row stride (32 + 8 pad)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  bf16* sF = reinterpret_cast<bf16*>(smem_raw) + warp * (kMaxFeat * LD + kMaxFeat * LDG);
  bf16* sG = sF + kMaxFeat * LD;
  const int nf = n_emb + 1;
  const int n_inter = nf * (nf - 1) / 2;
  zero_pad_rows<D>(sF, LD, n_emb, lane);

  for (int64_t s = static_cast<int64_t>(blockIdx.x) * kWarps + warp; s < batch;
       s += static_cast<int64_t>(gridDim.x) * kWarps) {
    stage_features<D>(sF, LD, bottom + s * bottom_stride, emb + s * emb_stride, n_emb, lane);
    const bf16* dzp = dz + s * dz_stride;
    for (int c = lane; c < kMaxFeat * LDG / 8; c += 32)
      reinterpret_cast<uint4*>(sG)[c] = make_uint4(0, 0, 0, 0);
    __syncwarp();
    for (int idx = lane; idx < n_inter; idx += 32) {
      int i = static_cast<int>((1.0f + sqrtf(1.0f + 8.0f * idx)) * 0.5f);
      while (i * (i - 1) / 2 > idx) --i;
      while ((i + 1) * i / 2 <= idx) ++i;
      const int j = idx - i * (i - 1) / 2;
      const bf16 v = dzp[idx];
      sG[i * LDG + j] = v;
      sG[j * LDG + i] = v;
    }
    __syncwarp();
    // A = G (32x32): 2 m-tiles x 2 k-steps, loaded once
    uint32_t ga[2][2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
        ldmatrix_x4(ga[mt][ks],
                    smem_u32(sG + (mt * 16 + (lane & 15)) * LDG + ks * 16 + ((lane >> 4) << 3)));
    const int cr = lane >> 2, cc = (lane & 3) << 1;
    bf16* dbp = dbottom + s * dbottom_stride;
    bf16* dep = demb + s * demb_stride;
#pragma unroll 1
    for (int n0 = 0; n0 < D; n0 += 32) {  // 4 n-tiles (32 columns) per pass
      float acc[2][4][4] = {};
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        uint32_t b01[4], b23[4];
        // B = F as K x N row major -> transposed loads
        const int krow = ks * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
        const int ncol = n0 + ((lane >> 4) << 3);
        ldmatrix_x4_trans(b01, smem_u32(sF + krow * LD + ncol));
        ldmatrix_x4_trans(b23, smem_u32(sF + krow * LD + ncol + 16));
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          mma_bf16(acc[mt][0], ga[mt][ks], b01[0], b01[1]);
          mma_bf16(acc[mt][1], ga[mt][ks], b01[2], b01[3]);
          mma_bf16(acc[mt][2], ga[mt][ks], b23[0], b23[1]);
          mma_bf16(acc[mt][3], ga[mt][ks], b23[2], b23[3]);
        }
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int row = mt * 16 + cr + half * 8;
            const int col = n0 + nt * 8 + cc;
            float v0 = acc[mt][nt][half * 2], v1 = acc[mt][nt][half * 2 + 1];
            if (row == 0) {
              // + gradient of the direct concat path z[n_inter + col]
              v0 += __bfloat162float(dzp[n_inter + col]);
              v1 += __bfloat162float(dzp[n_inter + col + 1]);
              *reinterpret_cast<__nv_bfloat162*>(dbp + col) = __floats2bfloat162_rn(v0, v1);
            } else if (row <= n_emb) {
              *reinterpret_cast<__nv_bfloat162*>(dep + (row - 1) * D + col) =
                  __floats2bfloat162_rn(v0 * emb_grad_scale, v1 * emb_grad_scale);
            }
          }
        }
      }
    }
    __syncwarp();
  }
}

// ---- v2 of the interaction backward (the default; measured 287 vs 370 us for v1 at one GPU,
// profiles/r2_experiments/summary.txt).  ncu of v1: 60 % of the issue stalls are long_scoreboard, 25 %
// warps active - a warp loads a sample, waits, computes, stores, and only then touches the next
// sample.  v2 keeps the *next* sample's features and dz row in flight (cp.async into a second
// buffer) while the current one is multiplied and stored, and reads dz from shared memory
// (16-byte chunks) instead of 351 scalar global loads.  8 warps/SM x 8 KB in flight each.
__device__ __forceinline__ void cp_async_commit() {
  asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait_group() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

constexpr int kDzMax = 512;  // staged dz elements per sample (n_inter + D <= 512)

template <int D>
__device__ __forceinline__ void issue_sample(bf16* sF, int LD, bf16* sDz, const bf16* bottom,
                                             const bf16* emb, const bf16* dz, int n_emb,
                                             int dz_chunks, int lane) {
  constexpr int kChunks = D / 8;
  const int total = (n_emb + 1) * kChunks;
#pragma unroll 4
  for (int c = lane; c < total; c += 32) {
    const int row = c / kChunks, ch = c - row * kChunks;
    const bf16* src = row == 0 ? bottom + ch * 8 : emb + (row - 1) * D + ch * 8;
    cp_async16(sF + row * LD + ch * 8, src);
  }
  for (int c = lane; c < dz_chunks; c += 32) cp_async16(sDz + c * 8, dz + c * 8);
}

// Per-block table of where every 16-byte chunk of a sample's embedding-gradient row goes:
// address of the chunk for local sample 0 and the byte stride between samples.  Built once per
// block from the route pieces (or from the single local buffer).
struct ChunkDst {
  unsigned long long base;
  long long stride;
};

template <int D>
__global__ void __launch_bounds__(kWarps * 32)
interact_bwd_v2_kernel(const bf16* __restrict__ bottom, int64_t bottom_stride,
                       const bf16* __restrict__ emb, int64_t emb_stride, int n_emb,
                       const bf16* __restrict__ dz, int64_t dz_stride, bf16* __restrict__ dbottom,
                       int64_t dbottom_stride, bf16* __restrict__ demb, int64_t demb_stride,
                       float emb_grad_scale, int64_t batch,
                       const GradRoute* __restrict__ routes, int n_routes,
                       const __grid_constant__ SyncArgs sync, uint32_t* __restrict__ done_counters,
                       int chunk_rows) {
  constexpr int LD = D + 8;
  constexpr int LDG = 40;
  constexpr int kWarpElems = 2 * kMaxFeat * LD + 2 * kDzMax + kMaxFeat * LDG;
  constexpr int kRowChunks = D / 8;  // 16-byte chunks per feature row
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  bf16* base = reinterpret_cast<bf16*>(smem_raw) + warp * kWarpElems;
  bf16* sDz0 = base + 2 * kMaxFeat * LD;
  bf16* sG = sDz0 + 2 * kDzMax;
  ChunkDst* sDst = reinterpret_cast<ChunkDst*>(reinterpret_cast<bf16*>(smem_raw) +
                                               kWarps * kWarpElems);
  const int nf = n_emb + 1;
  const int n_inter = nf * (nf - 1) / 2;
  const int dz_chunks = (n_inter + D + 7) >> 3;
  const int n_chunks = n_emb * kRowChunks;
  // chunk routing table (all chunks are covered: the host checks that the pieces tile the row)
  if (routes == nullptr) {
    for (int c = threadIdx.x; c < n_chunks; c += blockDim.x) {
      sDst[c].base = reinterpret_cast<unsigned long long>(demb + c * 8);
      sDst[c].stride = demb_stride * 2;
    }
  } else {
    for (int r = 0; r < n_routes; ++r) {
      const GradRoute R = routes[r];
      const int c0 = R.src_col >> 3, nc = R.width >> 3;
      for (int c = threadIdx.x; c < nc; c += blockDim.x) {
        sDst[c0 + c].base =
            reinterpret_cast<unsigned long long>(reinterpret_cast<bf16*>(R.dst) + R.dst_col + c * 8);
        sDst[c0 + c].stride = R.dst_stride * 2;
      }
    }
  }
  zero_pad_rows<D>(base, LD, n_emb, lane);
  zero_pad_rows<D>(base + kMaxFeat * LD, LD, n_emb, lane);
  __syncthreads();

  const int64_t stride = static_cast<int64_t>(gridDim.x) * kWarps;
  int64_t s = static_cast<int64_t>(blockIdx.x) * kWarps + warp;
  int cur = 0;
  if (s < batch)
    issue_sample<D>(base, LD, sDz0, bottom + s * bottom_stride, emb + s * emb_stride,
                    dz + s * dz_stride, n_emb, dz_chunks, lane);
  cp_async_commit();
  for (; s < batch; s += stride, cur ^= 1) {
    const int64_t nxt = s + stride;
    if (nxt < batch)
      issue_sample<D>(base + (cur ^ 1) * (kMaxFeat * LD), LD, sDz0 + (cur ^ 1) * kDzMax,
                      bottom + nxt * bottom_stride, emb + nxt * emb_stride, dz + nxt * dz_stride,
                      n_emb, dz_chunks, lane);
    cp_async_commit();        // possibly empty: keeps the group arithmetic uniform
    cp_async_wait_group<1>();  // everything but the prefetch just issued has landed
    __syncwarp();
    bf16* sF = base + cur * (kMaxFeat * LD);
    const bf16* sDz = sDz0 + cur * kDzMax;
    for (int c = lane; c < kMaxFeat * LDG / 8; c += 32)
      reinterpret_cast<uint4*>(sG)[c] = make_uint4(0, 0, 0, 0);
    __syncwarp();
    for (int idx = lane; idx < n_inter; idx += 32) {
      int i = static_cast<int>((1.0f + sqrtf(1.0f + 8.0f * idx)) * 0.5f);
      while (i * (i - 1) / 2 > idx) --i;
      while ((i + 1) * i / 2 <= idx) ++i;
      const int j = idx - i * (i - 1) / 2;
      const bf16 v = sDz[idx];
      sG[i * LDG + j] = v;
      sG[j * LDG + i] = v;
    }
    __syncwarp();
    uint32_t ga[2][2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
        ldmatrix_x4(ga[mt][ks],
                    smem_u32(sG + (mt * 16 + (lane & 15)) * LDG + ks * 16 + ((lane >> 4) << 3)));
    const int cr = lane >> 2, cc = (lane & 3) << 1;
#pragma unroll 1
    for (int n0 = 0; n0 < D; n0 += 32) {
      float acc[2][4][4] = {};
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        uint32_t b01[4], b23[4];
        const int krow = ks * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
        const int ncol = n0 + ((lane >> 4) << 3);
        ldmatrix_x4_trans(b01, smem_u32(sF + krow * LD + ncol));
        ldmatrix_x4_trans(b23, smem_u32(sF + krow * LD + ncol + 16));
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          mma_bf16(acc[mt][0], ga[mt][ks], b01[0], b01[1]);
          mma_bf16(acc[mt][1], ga[mt][ks], b01[2], b01[3]);
          mma_bf16(acc[mt][2], ga[mt][ks], b23[0], b23[1]);
          mma_bf16(acc[mt][3], ga[mt][ks], b23[2], b23[3]);
        }
      }
      // dF overwrites F in place: columns [n0, n0 + 32) are read by this pass only (the ldmatrix
      // loads above are warp collective, so every lane has its operands before any lane stores)
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int row = mt * 16 + cr + half * 8;
            const int col = n0 + nt * 8 + cc;
            float v0 = acc[mt][nt][half * 2], v1 = acc[mt][nt][half * 2 + 1];
            if (row == 0) {
              v0 += __bfloat162float(sDz[n_inter + col]);
              v1 += __bfloat162float(sDz[n_inter + col + 1]);
            } else {
              v0 *= emb_grad_scale;
              v1 *= emb_grad_scale;
            }
            if (row <= n_emb)
              *reinterpret_cast<__nv_bfloat162*>(sF + row * LD + col) =
                  __floats2bfloat162_rn(v0, v1);
          }
        }
      }
    }
    __syncwarp();
    // coalesced 16-byte copy-out: row 0 -> bottom-MLP gradient, rows 1.. -> the chunk's owner
    // (local buffer, or a peer's receive buffer over NVLink: 128-256 contiguous bytes per piece)
    if (lane < kRowChunks)
      *reinterpret_cast<uint4*>(dbottom + s * dbottom_stride + lane * 8) =
          *reinterpret_cast<const uint4*>(sF + lane * 8);
    for (int c = lane; c < n_chunks; c += 32) {
      const int row = 1 + c / kRowChunks, ch = c - (row - 1) * kRowChunks;
      const ChunkDst d = sDst[c];
      *reinterpret_cast<uint4*>(d.base + static_cast<unsigned long long>(s * d.stride)) =
          *reinterpret_cast<const uint4*>(sF + row * LD + ch * 8);
    }
    __syncwarp();  // all lanes are done with buffer `cur` before the next iteration refills it
    if (done_counters != nullptr && lane == 0) {
      // streamed push: tell the copy kernel that this sample's staged rows are complete
      __threadfence();
      atomicAdd(done_counters + s / chunk_rows, 1u);
    }
  }
  cp_async_wait_group<0>();
  sync_tail(sync);  // every gradient piece of this rank is on its way to its owner
}

// dy <- dy * (y > 0) (in place) ; db[c] += sum_rows dy   (db fp32, pre-zeroed)
// Each thread owns 8 columns (one 16-byte vector) and keeps 4 rows in flight; partial column sums
// are reduced across the block in shared memory, then one atomic per column per block.
constexpr int kRbUnroll = 4;
__global__ void __launch_bounds__(256)
relu_bwd_bias_kernel(bf16* __restrict__ dy, const bf16* __restrict__ y, float* __restrict__ db,
                     int64_t rows, int cols, int rows_per_block) {
  extern __shared__ float s_part[];  // [rows_par][cols]
  const int tpr = cols >> 3;                // threads per row
  const int rows_par = blockDim.x / tpr;    // rows processed concurrently
  const int tr = threadIdx.x / tpr, tc = threadIdx.x - tr * tpr;
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_block;
  const int64_t r1 = min(rows, r0 + rows_per_block);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (tr < rows_par) {
    for (int64_t r = r0 + tr; r < r1; r += static_cast<int64_t>(rows_par) * kRbUnroll) {
      uint4 g[kRbUnroll], a[kRbUnroll];
#pragma unroll
      for (int u = 0; u < kRbUnroll; ++u) {
        const int64_t rr = r + static_cast<int64_t>(u) * rows_par;
        if (rr < r1) {
          g[u] = *reinterpret_cast<const uint4*>(dy + rr * cols + tc * 8);
          a[u] = *reinterpret_cast<const uint4*>(y + rr * cols + tc * 8);
        }
      }
#pragma unroll
      for (int u = 0; u < kRbUnroll; ++u) {
        const int64_t rr = r + static_cast<int64_t>(u) * rows_par;
        if (rr < r1) {
          uint32_t* gw = reinterpret_cast<uint32_t*>(&g[u]);
          const uint32_t* aw = reinterpret_cast<const uint32_t*>(&a[u]);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float2 gf = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&gw[i]));
            const float2 af = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&aw[i]));
            gf.x = af.x > 0.f ? gf.x : 0.f;
            gf.y = af.y > 0.f ? gf.y : 0.f;
            acc[2 * i] += gf.x;
            acc[2 * i + 1] += gf.y;
            __nv_bfloat162 h = __floats2bfloat162_rn(gf.x, gf.y);
            gw[i] = *reinterpret_cast<uint32_t*>(&h);
          }
          *reinterpret_cast<uint4*>(dy + rr * cols + tc * 8) = g[u];
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s_part[tr * cols + tc * 8 + i] = acc[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    float v = 0.f;
    for (int t = 0; t < rows_par; ++t) v += s_part[t * cols + c];
    atomicAdd(db + c, v);
  }
}

// Final layer (K -> 1) + BCE-with-logits loss + backward of both, one warp per sample:
//   logit = <x, w> + b ; loss += softplus terms ; dlogit = (sigmoid(logit) - label) * inv_batch
//   dx = dlogit * w masked by (x > 0) (x is a ReLU output) ; dw += dlogit * x ; db += dlogit ;
//   dbias_prev[c] += dx[c]  (bias gradient of the layer that produced x)
template <int PL>
__global__ void __launch_bounds__(256)
head_loss_kernel(const bf16* __restrict__ x, int K, const bf16* __restrict__ w,
                 const bf16* __restrict__ bias, const float* __restrict__ labels, int64_t batch,
                 float inv_batch, bf16* __restrict__ dx, float* __restrict__ dw,
                 float* __restrict__ db, float* __restrict__ dbias_prev,
                 float* __restrict__ loss_sum, float* __restrict__ logits_out) {
  extern __shared__ float sred[];  // [2 * K + 2] per block: dw partial, dbias_prev partial
  float* s_dw = sred;
  float* s_dbp = sred + K;
  float* s_misc = sred + 2 * K;  // [0] loss, [1] db
  for (int i = threadIdx.x; i < 2 * K + 2; i += blockDim.x) sred[i] = 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  constexpr int per_lane = PL;  // K == 32 * PL
  const float b0 = __bfloat162float(bias[0]);
  float wreg[PL];
#pragma unroll
  for (int i = 0; i < PL; ++i) wreg[i] = __bfloat162float(w[lane + 32 * i]);
  float dw_acc[PL], dbp_acc[PL];
#pragma unroll
  for (int i = 0; i < PL; ++i) dw_acc[i] = dbp_acc[i] = 0.f;
  float loss_acc = 0.f, db_acc = 0.f;
  for (int64_t s = static_cast<int64_t>(blockIdx.x) * wpb + warp; s < batch;
       s += static_cast<int64_t>(gridDim.x) * wpb) {
    float xv[PL];
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < PL; ++i) {
      if (i < per_lane) {
        xv[i] = __bfloat162float(x[s * K + lane + 32 * i]);
        dot = fmaf(xv[i], wreg[i], dot);
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, off);
    const float logit = dot + b0;
    const float label = labels[s];
    // numerically stable BCE with logits
    const float loss = fmaxf(logit, 0.f) - logit * label + log1pf(__expf(-fabsf(logit)));
    const float sig = 1.f / (1.f + __expf(-logit));
    const float dl = (sig - label) * inv_batch;
    if (lane == 0) {
      loss_acc += loss;
      db_acc += dl;
      if (logits_out) logits_out[s] = logit;
    }
#pragma unroll
    for (int i = 0; i < PL; ++i) {
      if (i < per_lane) {
        dw_acc[i] = fmaf(dl, xv[i], dw_acc[i]);
        const float g = xv[i] > 0.f ? dl * wreg[i] : 0.f;
        dbp_acc[i] += g;
        dx[s * K + lane + 32 * i] = __float2bfloat16_rn(g);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < PL; ++i) {
    if (i < per_lane) {
      atomicAdd(&s_dw[lane + 32 * i], dw_acc[i]);
      atomicAdd(&s_dbp[lane + 32 * i], dbp_acc[i]);
    }
  }
  if (lane == 0) {
    atomicAdd(&s_misc[0], loss_acc);
    atomicAdd(&s_misc[1], db_acc);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    atomicAdd(dw + i, s_dw[i]);
    atomicAdd(dbias_prev + i, s_dbp[i]);
  }
  if (threadIdx.x == 0) {
    atomicAdd(loss_sum, s_misc[0] * inv_batch);
    atomicAdd(db, s_misc[1]);
  }
}

// p32 -= lr * g32 ; p16 = bf16(p32) ; g32 = 0   (lr read from device memory: graph replay safe)
__global__ void __launch_bounds__(256)
sgd_update_kernel(float* __restrict__ p32, bf16* __restrict__ p16, float* __restrict__ g32,
                  const float* __restrict__ lr_ptr, float grad_scale, int64_t n_vec4) {
  const float step = -(*lr_ptr) * grad_scale;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_vec4;
       i += stride) {
    float4 p = reinterpret_cast<float4*>(p32)[i];
    const float4 g = reinterpret_cast<const float4*>(g32)[i];
    p.x = fmaf(step, g.x, p.x);
    p.y = fmaf(step, g.y, p.y);
    p.z = fmaf(step, g.z, p.z);
    p.w = fmaf(step, g.w, p.w);
    reinterpret_cast<float4*>(p32)[i] = p;
    reinterpret_cast<float4*>(g32)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __nv_bfloat162 a = __floats2bfloat162_rn(p.x, p.y), b = __floats2bfloat162_rn(p.z, p.w);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&a);
    o.y = *reinterpret_cast<uint32_t*>(&b);
    reinterpret_cast<uint2*>(p16)[i] = o;
  }
}

// dst[r, 0:dst_cols] = bf16(src[r, 0:src_cols]) zero padded
__global__ void cast_pad_kernel(const float* __restrict__ src, int src_cols, bf16* __restrict__ dst,
                                int dst_cols, int64_t rows) {
  const int64_t n = rows * dst_cols;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += stride) {
    const int64_t r = i / dst_cols;
    const int c = static_cast<int>(i - r * dst_cols);
    dst[i] = __float2bfloat16_rn(c < src_cols ? src[r * src_cols + c] : 0.f);
  }
}

// ---- 1-D average pooling over the concatenated embeddings ("same" padding, partial windows
// divide by the number of real elements): the memory-bound stand-in for FM / pooling
// interactions in the synthetic models (reference synthetic_models.py:150-160, Keras
// AveragePooling1D).  out[r, j] = mean(x[r, j*stride-left : (j+1)*stride-left] ∩ [0, n)).
__global__ void __launch_bounds__(256)
avgpool_fwd_kernel(const bf16* __restrict__ x, int64_t x_stride, int n, bf16* __restrict__ out,
                   int64_t out_stride, int out_len, int stride, int left, int64_t rows) {
  const int64_t total = rows * out_len;
  const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += step) {
    const int64_t r = i / out_len;
    const int j = static_cast<int>(i - r * out_len);
    const int lo = max(0, j * stride - left), hi = min(n, (j + 1) * stride - left);
    const bf16* xp = x + r * x_stride;
    float acc = 0.f;
    for (int c = lo; c < hi; ++c) acc += __bfloat162float(xp[c]);
    out[r * out_stride + j] = __float2bfloat16_rn(hi > lo ? acc / static_cast<float>(hi - lo) : 0.f);
  }
}

// dx[r, c] = dout[r, window(c)] / count(window(c))
__global__ void __launch_bounds__(256)
avgpool_bwd_kernel(const bf16* __restrict__ dout, int64_t dout_stride, int out_len,
                   bf16* __restrict__ dx, int64_t dx_stride, int n, int stride, int left,
                   int64_t rows) {
  const int64_t total = rows * n;
  const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += step) {
    const int64_t r = i / n;
    const int c = static_cast<int>(i - r * n);
    const int j = (c + left) / stride;
    const int lo = max(0, j * stride - left), hi = min(n, (j + 1) * stride - left);
    const float g = j < out_len ? __bfloat162float(dout[r * dout_stride + j]) : 0.f;
    dx[r * dx_stride + c] = __float2bfloat16_rn(g / static_cast<float>(hi - lo));
  }
}

}  // namespace

void launch_avgpool_fwd(const void* x, int64_t x_stride, int n, void* out, int64_t out_stride,
                        int out_len, int stride, int left, int64_t rows, cudaStream_t stream) {
  if (rows <= 0 || out_len <= 0) return;
  int64_t blocks = (rows * out_len + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  avgpool_fwd_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
      reinterpret_cast<const bf16*>(x), x_stride, n, reinterpret_cast<bf16*>(out), out_stride,
      out_len, stride, left, rows);
}

void launch_avgpool_bwd(const void* dout, int64_t dout_stride, int out_len, void* dx,
                        int64_t dx_stride, int n, int stride, int left, int64_t rows,
                        cudaStream_t stream) {
  if (rows <= 0 || n <= 0) return;
  int64_t blocks = (rows * n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  avgpool_bwd_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
      reinterpret_cast<const bf16*>(dout), dout_stride, out_len, reinterpret_cast<bf16*>(dx),
      dx_stride, n, stride, left, rows);
}

bool launch_interact_fwd(const void* bottom, int64_t bottom_stride, const void* emb,
                         int64_t emb_stride, int n_emb, int dim, void* z, int64_t z_stride,
                         int z_width, int64_t batch, int sm_count, cudaStream_t stream,
                         const SyncArgs& sync) {
  if (n_emb + 1 > kMaxFeat || batch <= 0) return false;
  int64_t blocks = (batch + kWarps - 1) / kWarps;
  const int64_t cap = static_cast<int64_t>(sm_count) * 8;
  if (blocks > cap) blocks = cap;
#define DE_IFWD(DD)                                                                              \
  {                                                                                              \
    const size_t smem = kWarps * fwd_warp_bytes(DD);                                             \
    cudaFuncSetAttribute(interact_fwd_kernel<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                         static_cast<int>(smem));                                                \
    interact_fwd_kernel<DD><<<static_cast<unsigned>(blocks), kWarps * 32, smem, stream>>>(       \
        reinterpret_cast<const bf16*>(bottom), bottom_stride, reinterpret_cast<const bf16*>(emb), \
        emb_stride, n_emb, reinterpret_cast<bf16*>(z), z_stride, z_width, batch, sync);          \
    return true;                                                                                 \
  }
  if (dim == 128) DE_IFWD(128)
  if (dim == 64) DE_IFWD(64)
  if (dim == 32) DE_IFWD(32)
  if (dim == 16) DE_IFWD(16)
#undef DE_IFWD
  return false;
}

bool launch_interact_bwd(const void* bottom, int64_t bottom_stride, const void* emb,
                         int64_t emb_stride, int n_emb, int dim, const void* dz,
                         int64_t dz_stride, void* dbottom, int64_t dbottom_stride, void* demb,
                         int64_t demb_stride, float emb_grad_scale, int64_t batch, int sm_count,
                         cudaStream_t stream, const GradRoute* routes, int n_routes,
                         const SyncArgs& sync, uint32_t* done_counters, int chunk_rows) {
  if (n_emb + 1 > kMaxFeat || batch <= 0 || dim % 32 != 0) return false;
  int64_t blocks = (batch + kWarps - 1) / kWarps;
  const int64_t cap = static_cast<int64_t>(sm_count) * 8;
  if (blocks > cap) blocks = cap;
  // DE_B200_INTERACT_V1=1 selects the single-buffered kernel (local gradient buffer only)
  static const bool force_v1 = [] {
    const char* v = std::getenv("DE_B200_INTERACT_V1");
    return v != nullptr && v[0] == '1';
  }();
  const int nf = n_emb + 1;
  const int dz_elems = (nf * (nf - 1) / 2 + dim + 7) / 8 * 8;
  const bool v2_ok =
      dz_elems <= kDzMax && dz_elems <= dz_stride && dz_stride % 8 == 0 &&
      bottom_stride % 8 == 0 && emb_stride % 8 == 0 && dbottom_stride % 8 == 0 &&
      ((reinterpret_cast<uintptr_t>(dz) | reinterpret_cast<uintptr_t>(bottom) |
        reinterpret_cast<uintptr_t>(emb) | reinterpret_cast<uintptr_t>(dbottom)) & 15) == 0 &&
      (dim == 128 || dim == 64) &&
      (routes != nullptr || (demb_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(demb) & 15) == 0));
  if ((routes != nullptr || done_counters != nullptr) && !v2_ok) return false;  // v2 only
  const bool has_sync = sync.state != nullptr && (sync.wait_ch >= 0 || sync.signal_ch >= 0);
  if (v2_ok && (!force_v1 || routes != nullptr || has_sync || done_counters != nullptr)) {
    // two resident blocks per SM, each warp streams its samples through a double buffer
    int64_t blocks2 = (batch + kWarps - 1) / kWarps;
    if (blocks2 > static_cast<int64_t>(sm_count) * 2) blocks2 = static_cast<int64_t>(sm_count) * 2;
#define DE_IBWD2(DD)                                                                             \
  {                                                                                              \
    const size_t smem =                                                                          \
        kWarps * (2 * kMaxFeat * (DD + 8) + 2 * kDzMax + kMaxFeat * 40) * sizeof(bf16) +         \
        static_cast<size_t>(n_emb) * (DD / 8) * sizeof(ChunkDst);                                \
    cudaFuncSetAttribute(interact_bwd_v2_kernel<DD>,                                             \
                         cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));   \
    interact_bwd_v2_kernel<DD><<<static_cast<unsigned>(blocks2), kWarps * 32, smem, stream>>>(   \
        reinterpret_cast<const bf16*>(bottom), bottom_stride, reinterpret_cast<const bf16*>(emb), \
        emb_stride, n_emb, reinterpret_cast<const bf16*>(dz), dz_stride,                         \
        reinterpret_cast<bf16*>(dbottom), dbottom_stride, reinterpret_cast<bf16*>(demb),         \
        demb_stride, emb_grad_scale, batch, routes, n_routes, sync, done_counters, chunk_rows);  \
    return true;                                                                                 \
  }
    if (dim == 128) DE_IBWD2(128)
    if (dim == 64) DE_IBWD2(64)
#undef DE_IBWD2
  }
  if (has_sync) return false;
#define DE_IBWD(DD)                                                                              \
  {                                                                                              \
    const size_t smem = kWarps * (kMaxFeat * (DD + 8) + kMaxFeat * 40) * sizeof(bf16);           \
    cudaFuncSetAttribute(interact_bwd_kernel<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                         static_cast<int>(smem));                                                \
    interact_bwd_kernel<DD><<<static_cast<unsigned>(blocks), kWarps * 32, smem, stream>>>(       \
        reinterpret_cast<const bf16*>(bottom), bottom_stride, reinterpret_cast<const bf16*>(emb), \
        emb_stride, n_emb, reinterpret_cast<const bf16*>(dz), dz_stride,                         \
        reinterpret_cast<bf16*>(dbottom), dbottom_stride, reinterpret_cast<bf16*>(demb),         \
        demb_stride, emb_grad_scale, batch);                                                     \
    return true;                                                                                 \
  }
  if (dim == 128) DE_IBWD(128)
  if (dim == 64) DE_IBWD(64)
  if (dim == 32) DE_IBWD(32)
#undef DE_IBWD
  return false;
}

void launch_relu_bwd_bias(void* dy, const void* y, float* db, int64_t rows, int cols,
                          cudaStream_t stream) {
  if (rows <= 0) return;
  const int tpr = cols / 8;
  const int threads = 256;  // cols <= 2048
  const int rows_par = threads / tpr;
  // ~16 rows per thread (4 batches of 4 in flight); >= 2048 blocks at batch 64k
  const int rows_per_block = rows_par * kRbUnroll * 4;
  const int64_t blocks = (rows + rows_per_block - 1) / rows_per_block;
  const size_t smem = static_cast<size_t>(rows_par) * cols * sizeof(float);
  relu_bwd_bias_kernel<<<static_cast<unsigned>(blocks), threads, smem, stream>>>(
      reinterpret_cast<bf16*>(dy), reinterpret_cast<const bf16*>(y), db, rows, cols,
      rows_per_block);
}

bool launch_head_loss(const void* x, int K, const void* w, const void* bias, const float* labels,
                      int64_t batch, float inv_batch, void* dx, float* dw, float* db,
                      float* dbias_prev, float* loss_sum, float* logits_out, int sm_count,
                      cudaStream_t stream) {
  if (batch <= 0) return true;
  const int threads = 256;
  int64_t blocks = (batch + 7) / 8;
  if (blocks > sm_count * 4) blocks = sm_count * 4;
  const size_t smem = (2 * K + 2) * sizeof(float);
#define DE_HEAD(PL)                                                                              \
  head_loss_kernel<PL><<<static_cast<unsigned>(blocks), threads, smem, stream>>>(                \
      reinterpret_cast<const bf16*>(x), K, reinterpret_cast<const bf16*>(w),                     \
      reinterpret_cast<const bf16*>(bias), labels, batch, inv_batch, reinterpret_cast<bf16*>(dx), \
      dw, db, dbias_prev, loss_sum, logits_out)
  switch (K) {
    case 64: DE_HEAD(2); break;
    case 128: DE_HEAD(4); break;
    case 256: DE_HEAD(8); break;
    case 512: DE_HEAD(16); break;
    case 1024: DE_HEAD(32); break;
    default: return false;
  }
#undef DE_HEAD
  return true;
}

void launch_sgd_update(float* p32, void* p16, float* g32, const float* lr_ptr, float grad_scale,
                       int64_t n, int sm_count, cudaStream_t stream) {
  const int64_t n_vec4 = n / 4;  // buffers are padded to 16 bytes
  if (n_vec4 <= 0) return;
  int64_t blocks = (n_vec4 + 255) / 256;
  if (blocks > sm_count * 8) blocks = sm_count * 8;
  sgd_update_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
      p32, reinterpret_cast<bf16*>(p16), g32, lr_ptr, grad_scale, n_vec4);
}

void launch_cast_pad(const float* src, int src_cols, void* dst, int dst_cols, int64_t rows,
                     cudaStream_t stream) {
  if (rows <= 0) return;
  int64_t blocks = (rows * dst_cols + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  cast_pad_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
      src, src_cols, reinterpret_cast<bf16*>(dst), dst_cols, rows);
}

}  // namespace de
