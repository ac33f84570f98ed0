// NVLink / NVSwitch communication kernels for sm_100a: flag barrier over peer-mapped signal pads
// and the dense-gradient all-reduce (the data-parallel half of hybrid parallelism) as a single
// kernel: reduce-scatter + all-gather over peer memory, fused with the 1/world scale and the
// bf16/fp32 handling, either with plain P2P loads/stores or with NVSwitch multicast
// (multimem.ld_reduce / multimem.st -> in-switch reduction).
//
// Replaces Horovod's allreduce / barrier use (reference dist_model_parallel.py:1260, 985).
#include "common.cuh"

namespace de {

namespace {

// ----------------------------------------------------------------------------- barrier
// grid = 1 block.  Thread t < world signals peer t and waits for peer t.
__global__ void barrier_kernel(const __grid_constant__ PeerPtrs flags, uint32_t* epoch_p, int rank,
                               int world, int channel, unsigned long long timeout_cycles,
                               int* error_flag) {
  const uint32_t epoch = *epoch_p + 1;
  const int t = threadIdx.x;
  if (t < world) {
    __threadfence_system();
    st_release_sys(flag_slot(flags.p[t], channel, rank), epoch);
    if (!wait_flag_ge(flag_slot(flags.p[rank], channel, t), epoch, timeout_cycles)) {
      if (error_flag) atomicExch(error_flag, 1 + t);
    }
  }
  __syncthreads();
  if (t == 0) *epoch_p = epoch;
}

// Multi-block entry barrier: block 0 signals, every block waits (flags are local memory).
__device__ __forceinline__ void grid_peer_barrier_enter(const PeerPtrs& flags, uint32_t epoch,
                                                        int rank, int world, int channel,
                                                        unsigned long long timeout_cycles,
                                                        int* error_flag) {
  const int t = threadIdx.x;
  if (t < world) {
    if (blockIdx.x == 0) {
      __threadfence_system();
      st_release_sys(flag_slot(flags.p[t], channel, rank), epoch);
    }
    if (!wait_flag_ge(flag_slot(flags.p[rank], channel, t), epoch, timeout_cycles)) {
      if (error_flag) atomicExch(error_flag, 1 + t);
    }
  }
  __syncthreads();
}

// Multi-block exit barrier: the last block to finish signals the peers and waits for them, so
// kernel completion implies every rank's writes into this rank's buffer have landed.
__device__ __forceinline__ void grid_peer_barrier_exit(const PeerPtrs& flags, uint32_t epoch,
                                                       uint32_t* block_counter, uint32_t* epoch_p,
                                                       int rank, int world, int channel,
                                                       unsigned long long timeout_cycles,
                                                       int* error_flag) {
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const uint32_t done = atomicAdd(block_counter, 1u);
    last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return;
  const int t = threadIdx.x;
  if (t < world) {
    __threadfence_system();
    st_release_sys(flag_slot(flags.p[t], channel, rank), epoch);
    if (!wait_flag_ge(flag_slot(flags.p[rank], channel, t), epoch, timeout_cycles)) {
      if (error_flag) atomicExch(error_flag, 1 + t);
    }
  }
  __syncthreads();
  if (t == 0) {
    *block_counter = 0;
    *epoch_p = epoch;
  }
}

// ----------------------------------------------------------------------------- all-reduce
// Rank r owns the r-th slice: it sums the slice over all peers (P2P loads), scales, and stores
// the result into every peer's buffer (P2P stores).  16-byte accesses; n_vec = elements / VEC.
__device__ __forceinline__ void acc_f32(float (&a)[4], const uint4& v) {
  a[0] += __uint_as_float(v.x);
  a[1] += __uint_as_float(v.y);
  a[2] += __uint_as_float(v.z);
  a[3] += __uint_as_float(v.w);
}
__device__ __forceinline__ void acc_bf16(float (&a)[8], const uint4& v) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
    float2 f = __bfloat1622float2(h);
    a[2 * i] += f.x;
    a[2 * i + 1] += f.y;
  }
}

template <bool BF16>
__global__ void __launch_bounds__(512)
allreduce_p2p_kernel(const __grid_constant__ PeerPtrs bufs, const __grid_constant__ PeerPtrs flags,
                     uint32_t* epoch_p, uint32_t* block_counter, int rank, int world,
                     int64_t n_vec16, float scale, int channel, unsigned long long timeout_cycles,
                     int* error_flag) {
  const uint32_t epoch = *epoch_p + 1;
  grid_peer_barrier_enter(flags, epoch, rank, world, channel, timeout_cycles, error_flag);

  const int64_t per = (n_vec16 + world - 1) / world;
  const int64_t lo = per * rank;
  const int64_t hi = min(n_vec16, lo + per);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = lo + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < hi;
       i += stride) {
    uint4 v[kMaxPeers];
#pragma unroll
    for (int p = 0; p < kMaxPeers; ++p) {
      if (p < world) {
        const int q = (rank + p) % world;  // stagger peers
        v[p] = reinterpret_cast<const uint4*>(bufs.p[q])[i];
      }
    }
    uint4 o;
    if constexpr (BF16) {
      float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int p = 0; p < kMaxPeers; ++p)
        if (p < world) acc_bf16(a, v[p]);
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __nv_bfloat162 h = __floats2bfloat162_rn(a[2 * j] * scale, a[2 * j + 1] * scale);
        w[j] = *reinterpret_cast<uint32_t*>(&h);
      }
      o = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
      float a[4] = {0, 0, 0, 0};
#pragma unroll
      for (int p = 0; p < kMaxPeers; ++p)
        if (p < world) acc_f32(a, v[p]);
      o = make_uint4(__float_as_uint(a[0] * scale), __float_as_uint(a[1] * scale),
                     __float_as_uint(a[2] * scale), __float_as_uint(a[3] * scale));
    }
#pragma unroll
    for (int p = 0; p < kMaxPeers; ++p) {
      if (p < world) {
        const int q = (rank + p) % world;
        reinterpret_cast<uint4*>(bufs.p[q])[i] = o;
      }
    }
  }
  grid_peer_barrier_exit(flags, epoch + 1, block_counter, epoch_p, rank, world, channel,
                         timeout_cycles, error_flag);
}

// NVSwitch multicast variant: one multimem.ld_reduce pulls the switch-reduced 16 bytes, one
// multimem.st broadcasts the scaled result to every GPU.
template <bool BF16>
__global__ void __launch_bounds__(512)
allreduce_multimem_kernel(void* mc_ptr, const __grid_constant__ PeerPtrs flags, uint32_t* epoch_p,
                          uint32_t* block_counter, int rank, int world, int64_t n_vec16,
                          float scale, int channel, unsigned long long timeout_cycles,
                          int* error_flag) {
  const uint32_t epoch = *epoch_p + 1;
  grid_peer_barrier_enter(flags, epoch, rank, world, channel, timeout_cycles, error_flag);
  const int64_t per = (n_vec16 + world - 1) / world;
  const int64_t lo = per * rank;
  const int64_t hi = min(n_vec16, lo + per);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  uint4* mc = reinterpret_cast<uint4*>(mc_ptr);
  for (int64_t i = lo + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < hi;
       i += stride) {
    uint4 o;
    if constexpr (BF16) {
      uint32_t x0, x1, x2, x3;
      asm volatile(
          "multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
          : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3)
          : "l"(mc + i)
          : "memory");
      uint32_t w[4] = {x0, x1, x2, x3};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&w[j]);
        float2 f = __bfloat1622float2(h);
        h = __floats2bfloat162_rn(f.x * scale, f.y * scale);
        w[j] = *reinterpret_cast<uint32_t*>(&h);
      }
      o = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
      float a, b, c, d;
      asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                   : "=f"(a), "=f"(b), "=f"(c), "=f"(d)
                   : "l"(mc + i)
                   : "memory");
      o = make_uint4(__float_as_uint(a * scale), __float_as_uint(b * scale),
                     __float_as_uint(c * scale), __float_as_uint(d * scale));
    }
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc + i),
                 "f"(__uint_as_float(o.x)), "f"(__uint_as_float(o.y)), "f"(__uint_as_float(o.z)),
                 "f"(__uint_as_float(o.w))
                 : "memory");
  }
  grid_peer_barrier_exit(flags, epoch + 1, block_counter, epoch_p, rank, world, channel,
                         timeout_cycles, error_flag);
}

// Pull-style all-to-all of index segments: segment j copies n elements from peer src_rank's
// staging buffer into the local model-parallel id buffer (the reference's 'inp_dp_to_mp'
// hvd.alltoall, dist_model_parallel.py:211, as direct NVLink reads).
template <typename T>
__global__ void gather_segments_kernel(const int64_t* __restrict__ segs,
                                       const __grid_constant__ PeerPtrs src, T* __restrict__ dst) {
  const int64_t* sg = segs + 4 * static_cast<int64_t>(blockIdx.y);
  const T* sp = reinterpret_cast<const T*>(src.p[sg[0]]) + sg[1];
  T* dp = dst + sg[2];
  const int64_t n = sg[3];
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    dp[i] = sp[i];
}

// Double-buffered input staging for graph-replayed steps: copies the active staging slot
// (chosen by a device-resident flag, so one captured graph serves both slots) into the static
// input buffers.  All sizes are multiples of 16 bytes.
struct SelectSegs {
  const uint4* src[2][4];
  uint4* dst[4];
  int64_t n16[4];
  int count;
};
__global__ void select_copy_kernel(const __grid_constant__ SelectSegs segs,
                                   const int* __restrict__ slot_flag) {
  const int slot = (*slot_flag) & 1;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int s = 0; s < segs.count; ++s) {
    const uint4* src = segs.src[slot][s];
    uint4* dst = segs.dst[s];
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < segs.n16[s];
         i += stride)
      dst[i] = src[i];
  }
}

// Ragged index exchange over peer memory (replaces the reference's two hvd.alltoall calls for
// values + row lengths and the worker-major -> feature-major transpose, dist_model_parallel.py:
// 90-166).  Every requester stages values[cap] and row_splits[b+1] of a ragged feature in
// symmetric buffers; the owner pulls them from all sources and builds the global-batch CSR:
//   vals[item_off + base_s + j] = values_s[j],  goff[s*b + i] = base_s + splits_s[i],
// with base_s = sum of the earlier sources' nnz computed on the device (no host round trip).
// segs[j] = {src_val_off, dst_item_off, splits_off, goff_off}
template <typename T>
__global__ void gather_ragged_kernel(const int64_t* __restrict__ segs,
                                     const __grid_constant__ PeerPtrs src_vals,
                                     const __grid_constant__ PeerPtrs src_splits,
                                     T* __restrict__ dst_vals, int64_t* __restrict__ goff,
                                     int64_t b, int world) {
  const int64_t* sg = segs + 4 * static_cast<int64_t>(blockIdx.y);
  const int s = blockIdx.z;
  __shared__ int64_t s_base, s_nnz;
  if (threadIdx.x == 0) {
    int64_t base = 0;
    for (int q = 0; q < s; ++q)
      base += (reinterpret_cast<const int64_t*>(src_splits.p[q]) + sg[2])[b];
    s_base = base;
    s_nnz = (reinterpret_cast<const int64_t*>(src_splits.p[s]) + sg[2])[b];
  }
  __syncthreads();
  const int64_t base = s_base, nnz = s_nnz;
  const int64_t* sp = reinterpret_cast<const int64_t*>(src_splits.p[s]) + sg[2];
  const T* vals = reinterpret_cast<const T*>(src_vals.p[s]) + sg[0];
  T* dv = dst_vals + sg[1] + base;
  int64_t* go = goff + sg[3] + static_cast<int64_t>(s) * b;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t t0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (int64_t i = t0; i < nnz; i += stride) dv[i] = vals[i];
  for (int64_t i = t0; i < b; i += stride) go[i] = base + sp[i];
  if (s == world - 1 && t0 == 0) go[b] = base + nnz;
}

template <typename S, typename D>
__global__ void copy_cast_2d_kernel(const S* __restrict__ src, int64_t src_stride,
                                    D* __restrict__ dst, int64_t dst_stride, int64_t rows,
                                    int64_t cols, float scale) {
  const int64_t n = rows * cols;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += stride) {
    const int64_t r = i / cols, c = i - r * cols;
    float v;
    if constexpr (sizeof(S) == 4) v = src[r * src_stride + c];
    else v = __bfloat162float(src[r * src_stride + c]);
    v *= scale;
    if constexpr (sizeof(D) == 4) dst[r * dst_stride + c] = v;
    else dst[r * dst_stride + c] = __float2bfloat16_rn(v);
  }
}

}  // namespace

void launch_barrier(const PeerPtrs& flags, uint32_t* epoch, int rank, int world, int channel,
                    unsigned long long timeout_cycles, int* error_flag, cudaStream_t stream) {
  barrier_kernel<<<1, 32, 0, stream>>>(flags, epoch, rank, world, channel, timeout_cycles,
                                       error_flag);
}

// epoch[0] = epoch counter (advances by 2 per all-reduce), epoch[1] = block counter
void launch_allreduce(const PeerPtrs& bufs, const PeerPtrs& flags, uint32_t* epoch, int rank,
                      int world, int64_t n_elems, float scale, bool bf16, int channel,
                      unsigned long long timeout_cycles, int* error_flag, int sm_count,
                      cudaStream_t stream) {
  const int64_t per16 = bf16 ? 8 : 4;
  const int64_t n_vec16 = (n_elems + per16 - 1) / per16;  // buffers are padded to 16 bytes
  const int threads = 512;
  int64_t blocks = ((n_vec16 + world - 1) / world + threads - 1) / threads;
  if (blocks > sm_count) blocks = sm_count;  // must be co-resident: blocks spin on flags
  if (blocks < 1) blocks = 1;
  if (bf16)
    allreduce_p2p_kernel<true><<<static_cast<unsigned>(blocks), threads, 0, stream>>>(
        bufs, flags, epoch, epoch + 1, rank, world, n_vec16, scale, channel, timeout_cycles,
        error_flag);
  else
    allreduce_p2p_kernel<false><<<static_cast<unsigned>(blocks), threads, 0, stream>>>(
        bufs, flags, epoch, epoch + 1, rank, world, n_vec16, scale, channel, timeout_cycles,
        error_flag);
}

void launch_allreduce_multimem(void* mc_ptr, const PeerPtrs& flags, uint32_t* epoch, int rank,
                               int world, int64_t n_elems, float scale, bool bf16, int channel,
                               unsigned long long timeout_cycles, int* error_flag, int sm_count,
                               cudaStream_t stream) {
  const int64_t per16 = bf16 ? 8 : 4;
  const int64_t n_vec16 = (n_elems + per16 - 1) / per16;
  const int threads = 512;
  int64_t blocks = ((n_vec16 + world - 1) / world + threads - 1) / threads;
  if (blocks > sm_count) blocks = sm_count;
  if (blocks < 1) blocks = 1;
  if (bf16)
    allreduce_multimem_kernel<true><<<static_cast<unsigned>(blocks), threads, 0, stream>>>(
        mc_ptr, flags, epoch, epoch + 1, rank, world, n_vec16, scale, channel, timeout_cycles,
        error_flag);
  else
    allreduce_multimem_kernel<false><<<static_cast<unsigned>(blocks), threads, 0, stream>>>(
        mc_ptr, flags, epoch, epoch + 1, rank, world, n_vec16, scale, channel, timeout_cycles,
        error_flag);
}

void launch_select_copy(const void* const* src0, const void* const* src1, void* const* dst,
                        const int64_t* nbytes, int count, const int* slot_flag, int sm_count,
                        cudaStream_t stream) {
  SelectSegs segs;
  segs.count = count > 4 ? 4 : count;
  int64_t total = 0;
  for (int i = 0; i < segs.count; ++i) {
    segs.src[0][i] = reinterpret_cast<const uint4*>(src0[i]);
    segs.src[1][i] = reinterpret_cast<const uint4*>(src1[i]);
    segs.dst[i] = reinterpret_cast<uint4*>(dst[i]);
    segs.n16[i] = nbytes[i] / 16;
    total += segs.n16[i];
  }
  if (total == 0) return;
  int64_t blocks = (total + 255) / 256;
  if (blocks > sm_count * 4) blocks = sm_count * 4;
  select_copy_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(segs, slot_flag);
}

void launch_gather_ragged(const int64_t* segs, int n_seg, const PeerPtrs& src_vals,
                          const PeerPtrs& src_splits, void* dst_vals, int64_t* goff, int64_t b,
                          int world, int elem_bytes, int64_t max_cap, cudaStream_t stream) {
  if (n_seg <= 0) return;
  int64_t bx = (max_cap + 1023) / 1024;
  if (bx > 32) bx = 32;
  if (bx < 1) bx = 1;
  dim3 grid(static_cast<unsigned>(bx), static_cast<unsigned>(n_seg), static_cast<unsigned>(world));
  if (elem_bytes == 8)
    gather_ragged_kernel<int64_t><<<grid, 256, 0, stream>>>(
        segs, src_vals, src_splits, reinterpret_cast<int64_t*>(dst_vals), goff, b, world);
  else
    gather_ragged_kernel<int32_t><<<grid, 256, 0, stream>>>(
        segs, src_vals, src_splits, reinterpret_cast<int32_t*>(dst_vals), goff, b, world);
}

void launch_gather_segments(const int64_t* segs, int n_seg, const PeerPtrs& src, void* dst,
                            int elem_bytes, int64_t max_seg_elems, cudaStream_t stream) {
  if (n_seg <= 0 || max_seg_elems <= 0) return;
  const int threads = 256;
  int64_t bx = (max_seg_elems + threads * 4 - 1) / (threads * 4);
  if (bx > 64) bx = 64;
  if (bx < 1) bx = 1;
  dim3 grid(static_cast<unsigned>(bx), static_cast<unsigned>(n_seg));
  if (elem_bytes == 8)
    gather_segments_kernel<int64_t><<<grid, threads, 0, stream>>>(segs, src,
                                                                  reinterpret_cast<int64_t*>(dst));
  else
    gather_segments_kernel<int32_t><<<grid, threads, 0, stream>>>(segs, src,
                                                                  reinterpret_cast<int32_t*>(dst));
}

// 16-byte fast path: same dtype, unit scale, everything 16-byte aligned
__global__ void copy_2d_vec16_kernel(const uint4* __restrict__ src, int64_t src_stride16,
                                     uint4* __restrict__ dst, int64_t dst_stride16, int64_t rows,
                                     int64_t cols16) {
  const int64_t n = rows * cols16;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += stride) {
    const int64_t r = i / cols16, c = i - r * cols16;
    dst[r * dst_stride16 + c] = src[r * src_stride16 + c];
  }
}

void launch_copy_cast_2d(const void* src, int64_t src_stride, void* dst, int64_t dst_stride,
                         int64_t rows, int64_t cols, bool src_bf16, bool dst_bf16, float scale,
                         cudaStream_t stream) {
  if (rows <= 0 || cols <= 0) return;
  {
    const int64_t per16 = src_bf16 ? 8 : 4;
    if (src_bf16 == dst_bf16 && scale == 1.0f && cols % per16 == 0 && src_stride % per16 == 0 &&
        dst_stride % per16 == 0 &&
        ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
      const int64_t n = rows * (cols / per16);
      int64_t blocks = (n + 255) / 256;
      if (blocks > 148 * 16) blocks = 148 * 16;
      copy_2d_vec16_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
          reinterpret_cast<const uint4*>(src), src_stride / per16, reinterpret_cast<uint4*>(dst),
          dst_stride / per16, rows, cols / per16);
      return;
    }
  }
  const int threads = 256;
  int64_t blocks = (rows * cols + threads - 1) / threads;
  if (blocks > 148 * 8) blocks = 148 * 8;
#define DE_CC(S, D)                                                                              \
  copy_cast_2d_kernel<S, D><<<static_cast<unsigned>(blocks), threads, 0, stream>>>(              \
      reinterpret_cast<const S*>(src), src_stride, reinterpret_cast<D*>(dst), dst_stride, rows,  \
      cols, scale)
  if (src_bf16) {
    if (dst_bf16) DE_CC(__nv_bfloat16, __nv_bfloat16);
    else DE_CC(__nv_bfloat16, float);
  } else {
    if (dst_bf16) DE_CC(float, __nv_bfloat16);
    else DE_CC(float, float);
  }
#undef DE_CC
}

}  // namespace de
