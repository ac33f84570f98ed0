// Pooled embedding lookup forward and atomic scatter-add backward for sm_100a.
//
// One persistent, descriptor-driven kernel serves every table of the rank.  A warp owns a tile
// of 32 consecutive samples of one input; inside the warp, LPR lanes cooperate on one row
// (LPR * VEC columns per pass) and 32/LPR rows are in flight side by side, with a further 4x
// unroll so that >= 4 independent 16-byte row loads per lane are outstanding (HBM3e needs ~45 KB
// in flight per SM).  Sources of ids and destinations of pooled rows are *peer-mapped* pointers:
// with world_size > 1 the kernel reads indices straight out of the requesters' staging buffers
// and stores pooled rows straight into the requesters' output tensors over NVLink (the two
// all-to-alls of the reference, dist_model_parallel.py:211 and :872, fused into the lookup).
// Tile order interleaves destination ranks (rotated by the local rank) so all NVLink egress and
// every peer's ingress stay evenly loaded instead of all GPUs bursting at peer 0.
//
// Capability parity: EmbeddingLookUpVariableHot / ...HotWide (reference
// cc/kernels/embedding_lookup_kernels.cu:175-336) + the dense tf.gather/reduce path.
#include <cstdlib>

#include "common.cuh"

namespace de {

namespace {

constexpr int kThreads = 256;
constexpr int kWarpsPerBlock = kThreads / 32;
constexpr int kTile = 32;  // samples per warp tile
constexpr int kUnroll = 4;
constexpr int kBlocksPerSM = 4;  // 64 registers / thread -> 32 resident warps per SM

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

struct TileCoord {
  int f;        // local input
  int d;        // destination (requester) rank
  int64_t g0;   // first global sample
  int nsamp;    // samples in tile
};

__device__ __forceinline__ TileCoord decode_tile(int64_t t, int n_inputs, int n_dst,
                                                 int64_t tiles_per_dst, int64_t batch,
                                                 int64_t dst_batch, int rot) {
  TileCoord c;
  int dd = static_cast<int>(t % n_dst);
  int64_t rest = t / n_dst;
  c.f = static_cast<int>(rest % n_inputs);
  int64_t chunk = rest / n_inputs;
  c.d = (dd + rot) % n_dst;
  int64_t local0 = chunk * kTile;
  c.g0 = static_cast<int64_t>(c.d) * dst_batch + local0;
  int64_t lim = min(dst_batch, batch - static_cast<int64_t>(c.d) * dst_batch);
  int64_t rem = lim - local0;
  c.nsamp = rem < kTile ? static_cast<int>(rem < 0 ? 0 : rem) : kTile;
  return c;
}

template <typename IdT>
struct IdReader {
  const IdT* direct;      // per-input direct pointer (global batch order) or nullptr
  const int64_t* offsets; // CSR
  const PeerPtrs* src;
  int64_t ids_off;
  int64_t src_batch;
  int hot;

  // number of ids of sample g and the pointer to its first id
  __device__ __forceinline__ const IdT* sample(int64_t g, int& n) const {
    if (offsets != nullptr) {
      int64_t a = offsets[g], b = offsets[g + 1];
      n = static_cast<int>(b - a);
      return direct + a;
    }
    n = hot;
    if (direct != nullptr) return direct + g * hot;
    int64_t s = g / src_batch;
    int64_t i = g - s * src_batch;
    return reinterpret_cast<const IdT*>(src->p[s]) + ids_off + i * hot;
  }
};

template <typename IdT>
__device__ __forceinline__ IdReader<IdT> make_reader(const InputDesc& D, const PeerPtrs& src,
                                                     int64_t src_batch) {
  IdReader<IdT> r;
  r.direct = reinterpret_cast<const IdT*>(D.ids);
  r.offsets = D.offsets;
  r.src = &src;
  r.ids_off = D.ids_off;
  r.src_batch = src_batch;
  r.hot = D.hotness;
  return r;
}

// =============================================================================== forward
template <typename IdT, typename OutT, int VEC>
__global__ void __launch_bounds__(kThreads, kBlocksPerSM)
lookup_fwd_kernel(const InputDesc* __restrict__ descs, int n_inputs, int64_t batch,
                  int64_t src_batch, int64_t dst_batch, int64_t dst_stride,
                  const __grid_constant__ PeerPtrs src, const __grid_constant__ PeerPtrs dst,
                  int rot) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x) >> 5;
  const int64_t n_warps = static_cast<int64_t>(gridDim.x) * kWarpsPerBlock;
  const int n_dst = static_cast<int>((batch + dst_batch - 1) / dst_batch);
  const int64_t tiles_per_dst = (dst_batch + kTile - 1) / kTile;
  const int64_t total = static_cast<int64_t>(n_inputs) * n_dst * tiles_per_dst;

  for (int64_t t = warp; t < total; t += n_warps) {
    const TileCoord tc = decode_tile(t, n_inputs, n_dst, tiles_per_dst, batch, dst_batch, rot);
    if (tc.nsamp <= 0) continue;
    const InputDesc D = descs[tc.f];
    const int W = D.width;
    const int nvec = (W + VEC - 1) / VEC;           // VEC==4 requires W % 4 == 0
    const int lpr = min(32, pow2_ceil(nvec));       // lanes per row
    const int rpw = 32 / lpr;                       // rows in flight per warp
    const int sub = lane / lpr, li = lane - sub * lpr;
    const float* table = reinterpret_cast<const float*>(D.table);
    OutT* out_base = reinterpret_cast<OutT*>(dst.p[tc.d]);
    const int64_t i0 = tc.g0 - static_cast<int64_t>(tc.d) * dst_batch;
    const IdReader<IdT> rd = make_reader<IdT>(D, src, src_batch);
    const bool onehot = (D.hotness == 1) && (D.offsets == nullptr);
    const bool skip_empty = (D.flags & 1) != 0;  // row slices: only the owner of an id writes

    for (int c0 = 0; c0 < nvec; c0 += lpr) {        // column pass (one pass when W <= 128)
      const int cv = c0 + li;
      const bool col_ok = cv < nvec;
      const int col = cv * VEC;
      if (onehot) {
        // 4 samples per lane group in flight
        for (int r0 = 0; r0 < tc.nsamp; r0 += rpw * kUnroll) {
          FVec<VEC> acc[kUnroll];
          bool ok[kUnroll];
#pragma unroll
          for (int u = 0; u < kUnroll; ++u) {
            const int r = r0 + u * rpw + sub;
            ok[u] = (r < tc.nsamp) && col_ok;
            acc[u].zero();
            if (ok[u]) {
              int n;
              const IdT* p = rd.sample(tc.g0 + r, n);
              const int64_t id = static_cast<int64_t>(*p) + D.id_shift;
              if (static_cast<uint64_t>(id) < static_cast<uint64_t>(D.sub_rows))
                acc[u] = ld_f32<VEC>(table + (D.row_base + id) * W + col);
              else if (skip_empty)
                ok[u] = false;
            }
          }
#pragma unroll
          for (int u = 0; u < kUnroll; ++u) {
            if (ok[u]) {
              const int r = r0 + u * rpw + sub;
              st_act<OutT, VEC>(out_base + (i0 + r) * dst_stride + D.dst_col + col, acc[u]);
            }
          }
        }
      } else {
        for (int r0 = 0; r0 < tc.nsamp; r0 += rpw) {
          const int r = r0 + sub;
          if (r >= tc.nsamp || !col_ok) continue;
          int n;
          const IdT* p = rd.sample(tc.g0 + r, n);
          FVec<VEC> acc;
          acc.zero();
          int h = 0, hits = 0;
          constexpr int kHotUnroll = 8;  // rows of one sample in flight
          for (; h + kHotUnroll <= n; h += kHotUnroll) {
            int64_t id[kHotUnroll];
            FVec<VEC> x[kHotUnroll];
#pragma unroll
            for (int u = 0; u < kHotUnroll; ++u)
              id[u] = static_cast<int64_t>(p[h + u]) + D.id_shift;
#pragma unroll
            for (int u = 0; u < kHotUnroll; ++u) {
              x[u].zero();
              if (static_cast<uint64_t>(id[u]) < static_cast<uint64_t>(D.sub_rows)) {
                x[u] = ld_f32<VEC>(table + (D.row_base + id[u]) * W + col);
                ++hits;
              }
            }
#pragma unroll
            for (int u = 0; u < kHotUnroll; ++u) acc.add(x[u]);
          }
          for (; h < n; ++h) {
            const int64_t id = static_cast<int64_t>(p[h]) + D.id_shift;
            if (static_cast<uint64_t>(id) < static_cast<uint64_t>(D.sub_rows)) {
              acc.add(ld_f32<VEC>(table + (D.row_base + id) * W + col));
              ++hits;
            }
          }
          if (skip_empty && hits == 0) continue;
          if (D.combiner == 1 && n > 0) acc.scale(1.0f / static_cast<float>(n));
          st_act<OutT, VEC>(out_base + (i0 + r) * dst_stride + D.dst_col + col, acc);
        }
      }
    }
  }
}

// ---- one-hot forward with TMA bulk row copies (DE_B200_LOOKUP_BULK=1; EXPERIMENTAL, written
// after the round-1 GPU budget was spent).  The LSU version above keeps at most
// kUnroll x 512 B per lane group in flight and pays registers for it (80 regs -> 3 CTAs/SM, 69 %
// of the DRAM roofline).  Here every lane asks the TMA engine for its sample's whole row
// (`cp.async.bulk` global -> shared, completion counted in bytes on an mbarrier), two tiles of 32
// rows per warp are in flight (32 KB of a 128-wide fp32 table per warp, no registers), and the
// warp then streams the rows from shared memory to the requester's output row (peer memory)
// with the same coalesced bf16 / fp32 stores as above.
constexpr int kBulkWarps = 4;      // warps per block
constexpr int kBulkStages = 2;     // tiles in flight per warp
constexpr int kBulkMaxWidth = 128; // fp32 columns per row staged in shared memory

__device__ __forceinline__ uint32_t smem_u32_addr(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void bulk_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32_addr(bar)), "r"(count));
}
__device__ __forceinline__ void bulk_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32_addr(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "BULK_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra BULK_DONE;\n"
      "bra BULK_WAIT;\n"
      "BULK_DONE:\n"
      "}\n" ::"r"(smem_u32_addr(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                              uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32_addr(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32_addr(bar))
      : "memory");
}

template <typename IdT, typename OutT>
__global__ void __launch_bounds__(kBulkWarps * 32)
lookup_fwd_bulk_kernel(const InputDesc* __restrict__ descs, int n_inputs, int64_t batch,
                       int64_t src_batch, int64_t dst_batch, int64_t dst_stride,
                       const __grid_constant__ PeerPtrs src, const __grid_constant__ PeerPtrs dst,
                       int rot) {
  extern __shared__ __align__(128) unsigned char bulk_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  // per warp: kBulkStages x [32 rows][kBulkMaxWidth] fp32, then the barriers of all warps
  float* rows = reinterpret_cast<float*>(bulk_smem) +
                static_cast<size_t>(wib) * kBulkStages * kTile * kBulkMaxWidth;
  uint64_t* bars = reinterpret_cast<uint64_t*>(
                       bulk_smem + static_cast<size_t>(kBulkWarps) * kBulkStages * kTile *
                                       kBulkMaxWidth * sizeof(float)) +
                   wib * kBulkStages;
  if (lane == 0) {
    for (int s = 0; s < kBulkStages; ++s) bulk_mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  const int64_t warp = static_cast<int64_t>(blockIdx.x) * kBulkWarps + wib;
  const int64_t n_warps = static_cast<int64_t>(gridDim.x) * kBulkWarps;
  const int n_dst = static_cast<int>((batch + dst_batch - 1) / dst_batch);
  const int64_t tiles_per_dst = (dst_batch + kTile - 1) / kTile;
  const int64_t total = static_cast<int64_t>(n_inputs) * n_dst * tiles_per_dst;

  // issue the row copies of tile t into stage s (every lane: the row of its own sample)
  auto issue = [&](int64_t t, int s) {
    const TileCoord tc = decode_tile(t, n_inputs, n_dst, tiles_per_dst, batch, dst_batch, rot);
    const InputDesc& D = descs[tc.f];
    const int W = D.width;
    const uint32_t row_bytes = static_cast<uint32_t>(W) * 4u;
    float* stage = rows + static_cast<size_t>(s) * kTile * kBulkMaxWidth;
    const float* gsrc = nullptr;
    if (lane < tc.nsamp) {
      const IdReader<IdT> rd = make_reader<IdT>(D, src, src_batch);
      int n;
      const IdT* p = rd.sample(tc.g0 + lane, n);
      const int64_t id = static_cast<int64_t>(*p) + D.id_shift;
      if (static_cast<uint64_t>(id) < static_cast<uint64_t>(D.sub_rows))
        gsrc = reinterpret_cast<const float*>(D.table) + (D.row_base + id) * W;
    }
    const unsigned valid = __ballot_sync(0xffffffffu, gsrc != nullptr);
    // the previous readers of this stage are done (__syncwarp at the end of consume); order
    // their generic-proxy reads before the async-proxy writes that follow
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (lane == 0) bulk_mbar_expect_tx(&bars[s], __popc(valid) * row_bytes);
    __syncwarp();
    if (gsrc != nullptr) {
      bulk_copy_g2s(stage + lane * kBulkMaxWidth, gsrc, row_bytes, &bars[s]);
    } else if (lane < tc.nsamp) {
      for (int c = 0; c < W; c += 4)  // out-of-range id: the sample pools to zero
        *reinterpret_cast<float4*>(stage + lane * kBulkMaxWidth + c) = make_float4(0, 0, 0, 0);
    }
  };

  uint32_t parity = 0;  // bit s: phase parity of stage s
  int64_t t = warp;
  int s = 0;
  if (t < total) issue(t, 0);
  for (; t < total; t += n_warps, s ^= 1) {
    const int64_t tn = t + n_warps;
    if (tn < total) issue(tn, s ^ 1);  // keep the next tile's rows in flight
    bulk_mbar_wait(&bars[s], (parity >> s) & 1u);
    parity ^= 1u << s;
    __syncwarp();  // zero-filled rows of other lanes are visible too
    const TileCoord tc = decode_tile(t, n_inputs, n_dst, tiles_per_dst, batch, dst_batch, rot);
    const InputDesc& D = descs[tc.f];
    const int W = D.width;
    const float* stage = rows + static_cast<size_t>(s) * kTile * kBulkMaxWidth;
    OutT* out_base = reinterpret_cast<OutT*>(dst.p[tc.d]);
    const int64_t i0 = tc.g0 - static_cast<int64_t>(tc.d) * dst_batch;
    // lane = 4 columns; one row per pass when W == 128, several rows per pass for narrow tables
    const int nvec = W >> 2;
    const int lpr = min(32, pow2_ceil(nvec));
    const int rpw = 32 / lpr;
    const int sub = lane / lpr, li = lane - sub * lpr;
    for (int r0 = 0; r0 < tc.nsamp; r0 += rpw) {
      const int r = r0 + sub;
      if (r < tc.nsamp) {
        for (int cv = li; cv < nvec; cv += lpr) {
          const float4 v = *reinterpret_cast<const float4*>(stage + r * kBulkMaxWidth + cv * 4);
          FVec<4> x;
          x.v[0] = v.x;
          x.v[1] = v.y;
          x.v[2] = v.z;
          x.v[3] = v.w;
          st_act<OutT, 4>(out_base + (i0 + r) * dst_stride + D.dst_col + cv * 4, x);
        }
      }
    }
    __syncwarp();  // all lanes are done with stage s before it is refilled
  }
}

// =============================================================================== backward
// dst_table[row] += scale * w_sample * grad_row   (vector RED, no return value)
// (the 8-column variant keeps 32 gradient floats per lane in flight: give it 128 registers,
//  at 64 it spills ~440 bytes per thread)
template <typename IdT, typename GradT, int VEC>
__global__ void __launch_bounds__(kThreads, VEC == 8 ? 2 : kBlocksPerSM)
scatter_add_bwd_kernel(const InputDesc* __restrict__ descs, int n_inputs, int64_t batch,
                       int64_t src_batch, int64_t grad_batch, int64_t grad_stride,
                       const __grid_constant__ PeerPtrs src, const __grid_constant__ PeerPtrs grad,
                       int rot, float scale, const float* __restrict__ scale_ptr) {
  if (scale_ptr != nullptr) scale *= *scale_ptr;  // device-resident lr (CUDA-graph friendly)
  const int lane = threadIdx.x & 31;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x) >> 5;
  const int64_t n_warps = static_cast<int64_t>(gridDim.x) * kWarpsPerBlock;
  const int n_dst = static_cast<int>((batch + grad_batch - 1) / grad_batch);
  const int64_t tiles_per_dst = (grad_batch + kTile - 1) / kTile;
  const int64_t total = static_cast<int64_t>(n_inputs) * n_dst * tiles_per_dst;

  for (int64_t t = warp; t < total; t += n_warps) {
    const TileCoord tc = decode_tile(t, n_inputs, n_dst, tiles_per_dst, batch, grad_batch, rot);
    if (tc.nsamp <= 0) continue;
    const InputDesc D = descs[tc.f];
    const int W = D.width;
    const int nvec = (W + VEC - 1) / VEC;
    const int lpr = min(32, pow2_ceil(nvec));
    const int rpw = 32 / lpr;
    const int sub = lane / lpr, li = lane - sub * lpr;
    float* table = reinterpret_cast<float*>(const_cast<void*>(D.table));
    const GradT* grad_base = reinterpret_cast<const GradT*>(grad.p[tc.d]);
    const int64_t i0 = tc.g0 - static_cast<int64_t>(tc.d) * grad_batch;
    const IdReader<IdT> rd = make_reader<IdT>(D, src, src_batch);

    // Pull the table rows of the *next* tile of this warp into L2 so that its reductions find
    // their lines resident (an atomic that misses L2 waits for the DRAM fill at the slice).
    {
      const int64_t tn = t + n_warps;
      if (tn < total) {
        const TileCoord nc = decode_tile(tn, n_inputs, n_dst, tiles_per_dst, batch, grad_batch,
                                         rot);
        if (lane < nc.nsamp) {
          const InputDesc& N = descs[nc.f];
          const IdReader<IdT> nrd = make_reader<IdT>(N, src, src_batch);
          int nn;
          const IdT* np = nrd.sample(nc.g0 + lane, nn);
          const int lines = (N.width * 4 + 127) >> 7;
          for (int h = 0; h < min(nn, 4); ++h) {
            const int64_t id = static_cast<int64_t>(np[h]) + N.id_shift;
            if (static_cast<uint64_t>(id) < static_cast<uint64_t>(N.sub_rows)) {
              const char* row = reinterpret_cast<const char*>(N.table) +
                                (N.row_base + id) * N.width * 4;
              for (int l = 0; l < lines; ++l) prefetch_l2(row + (l << 7));
            }
          }
        }
      }
    }

    for (int c0 = 0; c0 < nvec; c0 += lpr) {
      const int cv = c0 + li;
      const bool col_ok = cv < nvec;
      const int col = cv * VEC;
      for (int r0 = 0; r0 < tc.nsamp; r0 += rpw * kUnroll) {
        FVec<VEC> g[kUnroll];
        const IdT* p[kUnroll];
        int n[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const int r = r0 + u * rpw + sub;
          n[u] = 0;
          p[u] = nullptr;
          if (r < tc.nsamp && col_ok) {
            p[u] = rd.sample(tc.g0 + r, n[u]);
            g[u] = ld_act<GradT, VEC>(grad_base + (i0 + r) * grad_stride + D.dst_col + col);
            float w = scale;
            if (D.combiner == 1 && n[u] > 0) w /= static_cast<float>(n[u]);
            g[u].scale(w);
          }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          for (int h = 0; h < n[u]; ++h) {
            const int64_t id = static_cast<int64_t>(p[u][h]) + D.id_shift;
            if (static_cast<uint64_t>(id) < static_cast<uint64_t>(D.sub_rows))
              red_add_f32<VEC>(table + (D.row_base + id) * W + col, g[u]);
          }
        }
      }
    }
  }
}

// ---- tiny tables (DE_B200_TINY_TABLES=1; EXPERIMENTAL, written after the round-1 GPU budget
// was spent).  A table with a handful of rows receives `batch` reductions per step on the same
// few L2 lines: ncu of the kernel above shows the atomic unit of the busiest L2 slice at 56 %
// while the average slice sits at 17 %.  Here one block owns a chunk of samples of ONE tiny
// one-hot input, accumulates the gradient rows in shared memory (fp32 smem atomics), and issues
// one vector RED per touched row and 4 columns: chunk-size times fewer L2 atomics.
constexpr int kTinyChunk = 2048;  // samples per block
constexpr int kTinyUnroll = 4;

template <typename IdT, typename GradT>
__global__ void __launch_bounds__(kThreads)
tiny_scatter_add_kernel(const InputDesc* __restrict__ descs, int n_inputs, int64_t batch,
                        int64_t src_batch, int64_t grad_batch, int64_t grad_stride,
                        const __grid_constant__ PeerPtrs src,
                        const __grid_constant__ PeerPtrs grad, float scale,
                        const float* __restrict__ scale_ptr, int n_chunks) {
  extern __shared__ float s_acc[];  // [sub_rows][width]
  if (scale_ptr != nullptr) scale *= *scale_ptr;
  const int f = blockIdx.x / n_chunks;
  const int chunk = blockIdx.x - f * n_chunks;
  if (f >= n_inputs) return;
  const InputDesc D = descs[f];
  const int W = D.width;
  const int rows = static_cast<int>(D.sub_rows);
  for (int i = threadIdx.x; i < rows * W; i += kThreads) s_acc[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const IdReader<IdT> rd = make_reader<IdT>(D, src, src_batch);
  const int64_t g_lo = static_cast<int64_t>(chunk) * kTinyChunk;
  const int64_t g_hi = min(batch, g_lo + kTinyChunk);
  // warp per sample, lane per 4 columns (columns beyond 128 in further passes)
  for (int64_t g0 = g_lo + warp; g0 < g_hi; g0 += kWarpsPerBlock * kTinyUnroll) {
    for (int c0 = lane * 4; c0 < W; c0 += 128) {
      FVec<4> gv[kTinyUnroll];
      int64_t id[kTinyUnroll];
#pragma unroll
      for (int u = 0; u < kTinyUnroll; ++u) {
        const int64_t g = g0 + static_cast<int64_t>(u) * kWarpsPerBlock;
        id[u] = -1;
        if (g < g_hi) {
          int n;
          const IdT* p = rd.sample(g, n);
          id[u] = static_cast<int64_t>(p[0]) + D.id_shift;
          const int64_t d = g / grad_batch;
          const int64_t i = g - d * grad_batch;
          gv[u] = ld_act<GradT, 4>(reinterpret_cast<const GradT*>(grad.p[d]) + i * grad_stride +
                                   D.dst_col + c0);
        }
      }
#pragma unroll
      for (int u = 0; u < kTinyUnroll; ++u) {
        if (static_cast<uint64_t>(id[u]) < static_cast<uint64_t>(rows)) {
          float* a = s_acc + static_cast<int>(id[u]) * W + c0;
#pragma unroll
          for (int k = 0; k < 4; ++k) atomicAdd(a + k, gv[u].v[k] * scale);
        }
      }
    }
  }
  __syncthreads();
  float* table = reinterpret_cast<float*>(const_cast<void*>(D.table));
  const int nvec = W >> 2;
  for (int i = threadIdx.x; i < rows * nvec; i += kThreads) {
    const int r = i / nvec, c = (i - r * nvec) << 2;
    FVec<4> v;
#pragma unroll
    for (int k = 0; k < 4; ++k) v.v[k] = s_acc[r * W + c + k];
    if (v.v[0] != 0.f || v.v[1] != 0.f || v.v[2] != 0.f || v.v[3] != 0.f)
      red_add_f32<4>(table + (D.row_base + r) * W + c, v);
  }
}

// DE_B200_EMB_BLOCKS_PER_SM=1..4 caps the resident CTAs per SM of the persistent lookup / scatter
// grids (default 4 = the launch bound): fewer CTAs leave registers and shared memory for kernels
// of other streams (the MLP GEMMs overlapped with the embedding exchange).
int blocks_per_sm_cap(int compiled) {
  static const int env = [] {
    const char* v = std::getenv("DE_B200_EMB_BLOCKS_PER_SM");
    return v != nullptr ? std::atoi(v) : 0;
  }();
  return (env >= 1 && env < compiled) ? env : compiled;
}

int grid_for(int64_t total_tiles, int sm_count, int blocks_per_sm) {
  blocks_per_sm = blocks_per_sm_cap(blocks_per_sm);
  int64_t blocks = (total_tiles + kWarpsPerBlock - 1) / kWarpsPerBlock;
  int64_t cap = static_cast<int64_t>(sm_count) * blocks_per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

int64_t count_tiles(int n_inputs, int64_t batch, int64_t dst_batch) {
  int64_t n_dst = (batch + dst_batch - 1) / dst_batch;
  int64_t tiles_per_dst = (dst_batch + kTile - 1) / kTile;
  return static_cast<int64_t>(n_inputs) * n_dst * tiles_per_dst;
}

}  // namespace

#define DE_DISPATCH_FWD(IdT, OutT, VEC)                                                        \
  lookup_fwd_kernel<IdT, OutT, VEC><<<grid, kThreads, 0, stream>>>(                            \
      descs, n_inputs, batch, src_batch, dst_batch, dst_stride, src, dst, rot)

void launch_lookup_fwd(const InputDesc* descs, int n_inputs, int64_t batch, int64_t src_batch,
                       int64_t dst_batch, int64_t dst_stride, const PeerPtrs& src,
                       const PeerPtrs& dst, int rot, bool ids64, bool out_bf16, bool vec4,
                       int sm_count, cudaStream_t stream) {
  if (n_inputs <= 0 || batch <= 0) return;
  const int grid = grid_for(count_tiles(n_inputs, batch, dst_batch), sm_count, kBlocksPerSM);
  if (vec4) {
    if (ids64) {
      if (out_bf16) DE_DISPATCH_FWD(int64_t, __nv_bfloat16, 4);
      else DE_DISPATCH_FWD(int64_t, float, 4);
    } else {
      if (out_bf16) DE_DISPATCH_FWD(int32_t, __nv_bfloat16, 4);
      else DE_DISPATCH_FWD(int32_t, float, 4);
    }
  } else {
    if (ids64) {
      if (out_bf16) DE_DISPATCH_FWD(int64_t, __nv_bfloat16, 1);
      else DE_DISPATCH_FWD(int64_t, float, 1);
    } else {
      if (out_bf16) DE_DISPATCH_FWD(int32_t, __nv_bfloat16, 1);
      else DE_DISPATCH_FWD(int32_t, float, 1);
    }
  }
}

#define DE_DISPATCH_BWD(IdT, GradT, VEC)                                                       \
  scatter_add_bwd_kernel<IdT, GradT, VEC><<<grid, kThreads, 0, stream>>>(                      \
      descs, n_inputs, batch, src_batch, grad_batch, grad_stride, src, grad, rot, scale, scale_ptr)

void launch_scatter_add_bwd(const InputDesc* descs, int n_inputs, int64_t batch, int64_t src_batch,
                            int64_t grad_batch, int64_t grad_stride, const PeerPtrs& src,
                            const PeerPtrs& grad, int rot, float scale, const float* scale_ptr,
                            bool ids64, bool grad_bf16, bool vec4, int sm_count,
                            cudaStream_t stream, bool vec8) {
  if (n_inputs <= 0 || batch <= 0) return;
  const int grid = grid_for(count_tiles(n_inputs, batch, grad_batch), sm_count, kBlocksPerSM);
  if (vec8 && grad_bf16) {
    // 16-byte gradient loads: 8 columns per lane, two rows per warp instruction (peer pulls)
    if (ids64) DE_DISPATCH_BWD(int64_t, __nv_bfloat16, 8);
    else DE_DISPATCH_BWD(int32_t, __nv_bfloat16, 8);
    return;
  }
  if (vec4) {
    if (ids64) {
      if (grad_bf16) DE_DISPATCH_BWD(int64_t, __nv_bfloat16, 4);
      else DE_DISPATCH_BWD(int64_t, float, 4);
    } else {
      if (grad_bf16) DE_DISPATCH_BWD(int32_t, __nv_bfloat16, 4);
      else DE_DISPATCH_BWD(int32_t, float, 4);
    }
  } else {
    if (ids64) {
      if (grad_bf16) DE_DISPATCH_BWD(int64_t, __nv_bfloat16, 1);
      else DE_DISPATCH_BWD(int64_t, float, 1);
    } else {
      if (grad_bf16) DE_DISPATCH_BWD(int32_t, __nv_bfloat16, 1);
      else DE_DISPATCH_BWD(int32_t, float, 1);
    }
  }
}

// One-hot inputs of tables with at most max_rows rows (width % 4 == 0): shared-memory
// pre-reduction per 2048-sample chunk, then one vector RED per touched row segment.
bool launch_tiny_scatter_add(const InputDesc* descs, int n_inputs, int64_t batch,
                             int64_t src_batch, int64_t grad_batch, int64_t grad_stride,
                             const PeerPtrs& src, const PeerPtrs& grad, float scale,
                             const float* scale_ptr, bool ids64, bool grad_bf16, int max_rows,
                             int max_width, cudaStream_t stream) {
  if (n_inputs <= 0 || batch <= 0) return true;
  const size_t smem = static_cast<size_t>(max_rows) * max_width * sizeof(float);
  if (smem > 96 * 1024 || (max_width & 3)) return false;
  const int n_chunks = static_cast<int>((batch + kTinyChunk - 1) / kTinyChunk);
  const unsigned grid = static_cast<unsigned>(n_inputs) * n_chunks;
#define DE_TINY(IdT, GradT)                                                                       \
  {                                                                                               \
    cudaFuncSetAttribute(tiny_scatter_add_kernel<IdT, GradT>,                                     \
                         cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));    \
    tiny_scatter_add_kernel<IdT, GradT><<<grid, kThreads, smem, stream>>>(                        \
        descs, n_inputs, batch, src_batch, grad_batch, grad_stride, src, grad, scale, scale_ptr,  \
        n_chunks);                                                                                \
  }
  if (ids64) {
    if (grad_bf16) DE_TINY(int64_t, __nv_bfloat16) else DE_TINY(int64_t, float)
  } else {
    if (grad_bf16) DE_TINY(int32_t, __nv_bfloat16) else DE_TINY(int32_t, float)
  }
#undef DE_TINY
  return cudaGetLastError() == cudaSuccess;
}

// One-hot inputs (hotness 1, no CSR, no skip-empty flag), widths % 4 == 0 and <= 128, destination
// columns % 4 == 0: the caller checks the descriptors, this only launches.
bool launch_lookup_fwd_bulk(const InputDesc* descs, int n_inputs, int64_t batch,
                            int64_t src_batch, int64_t dst_batch, int64_t dst_stride,
                            const PeerPtrs& src, const PeerPtrs& dst, int rot, bool ids64,
                            bool out_bf16, int sm_count, cudaStream_t stream) {
  if (n_inputs <= 0 || batch <= 0) return true;
  const size_t smem = static_cast<size_t>(kBulkWarps) * kBulkStages * kTile * kBulkMaxWidth *
                          sizeof(float) +
                      kBulkWarps * kBulkStages * sizeof(uint64_t);
  const int64_t tiles = count_tiles(n_inputs, batch, dst_batch);
  int64_t blocks = (tiles + kBulkWarps - 1) / kBulkWarps;
  const int64_t cap = static_cast<int64_t>(sm_count);  // 128 KB of staging: one CTA per SM
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
#define DE_BULK(IdT, OutT)                                                                        \
  {                                                                                               \
    cudaFuncSetAttribute(lookup_fwd_bulk_kernel<IdT, OutT>,                                       \
                         cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));    \
    lookup_fwd_bulk_kernel<IdT, OutT><<<static_cast<unsigned>(blocks), kBulkWarps * 32, smem,     \
                                        stream>>>(descs, n_inputs, batch, src_batch, dst_batch,   \
                                                  dst_stride, src, dst, rot);                     \
  }
  if (ids64) {
    if (out_bf16) DE_BULK(int64_t, __nv_bfloat16) else DE_BULK(int64_t, float)
  } else {
    if (out_bf16) DE_BULK(int32_t, __nv_bfloat16) else DE_BULK(int32_t, float)
  }
#undef DE_BULK
  return cudaGetLastError() == cudaSuccess;
}

}  // namespace de
