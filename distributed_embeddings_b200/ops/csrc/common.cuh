// Device helpers shared by the sm_100a kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "de_b200.h"

namespace de {

constexpr int64_t kSentinelKey = 0x7fffffffffffffffLL;

__host__ __device__ __forceinline__ int pow2_ceil(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

// ---- small fixed-size vectors of fp32 ---------------------------------------------------
template <int VEC>
struct FVec {
  float v[VEC];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < VEC; ++i) v[i] = 0.f;
  }
  __device__ __forceinline__ void fma(float a, const FVec& x) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) v[i] = fmaf(a, x.v[i], v[i]);
  }
  __device__ __forceinline__ void add(const FVec& x) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) v[i] += x.v[i];
  }
  __device__ __forceinline__ void scale(float a) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) v[i] *= a;
  }
};

// Read-only fp32 row fragment (tables are immutable during the forward kernel).
template <int VEC>
__device__ __forceinline__ FVec<VEC> ld_f32(const float* p) {
  FVec<VEC> r;
  if constexpr (VEC == 4) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) r.v[i] = __ldg(p + i);
  }
  return r;
}

// Plain (coherent) fp32 load: used for read-modify-write of weights / optimizer state.
template <int VEC>
__device__ __forceinline__ FVec<VEC> ld_f32_rw(const float* p) {
  FVec<VEC> r;
  if constexpr (VEC == 8) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    const float4 b = *reinterpret_cast<const float4*>(p + 4);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  } else if constexpr (VEC == 4) {
    float4 t = *reinterpret_cast<const float4*>(p);
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) r.v[i] = p[i];
  }
  return r;
}

template <int VEC>
__device__ __forceinline__ void st_f32(float* p, const FVec<VEC>& x) {
  if constexpr (VEC == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(x.v[0], x.v[1], x.v[2], x.v[3]);
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) p[i] = x.v[i];
  }
}

// Typed load of an activation / gradient fragment (fp32 or bf16 in memory -> fp32 registers).
// Plain loads: the source may be peer-mapped memory written by another GPU before a barrier.
template <typename T, int VEC>
__device__ __forceinline__ FVec<VEC> ld_act(const T* p) {
  if constexpr (sizeof(T) == 4) {
    return ld_f32_rw<VEC>(reinterpret_cast<const float*>(p));
  } else if constexpr (VEC == 8) {
    FVec<VEC> r;
    const uint4 t = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
      r.v[2 * i] = f.x;
      r.v[2 * i + 1] = f.y;
    }
    return r;
  } else if constexpr (VEC == 4) {
    FVec<VEC> r;
    uint2 t = *reinterpret_cast<const uint2*>(p);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x);
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
    float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    r.v[0] = fa.x; r.v[1] = fa.y; r.v[2] = fb.x; r.v[3] = fb.y;
    return r;
  } else {
    FVec<VEC> r;
#pragma unroll
    for (int i = 0; i < VEC; ++i) r.v[i] = __bfloat162float(p[i]);
    return r;
  }
}

template <typename T, int VEC>
__device__ __forceinline__ void st_act(T* p, const FVec<VEC>& x) {
  if constexpr (sizeof(T) == 4) {
    st_f32<VEC>(reinterpret_cast<float*>(p), x);
  } else if constexpr (VEC == 4) {
    __nv_bfloat162 a = __floats2bfloat162_rn(x.v[0], x.v[1]);
    __nv_bfloat162 b = __floats2bfloat162_rn(x.v[2], x.v[3]);
    uint2 t;
    t.x = *reinterpret_cast<uint32_t*>(&a);
    t.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = t;
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) p[i] = __float2bfloat16_rn(x.v[i]);
  }
}

// Fire-and-forget vector reduction into global memory (REDG.E.ADD.F32x4 on sm_100a).
template <int VEC>
__device__ __forceinline__ void red_add_f32(float* p, const FVec<VEC>& x) {
  if constexpr (VEC == 8) {
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(x.v[0]),
                 "f"(x.v[1]), "f"(x.v[2]), "f"(x.v[3])
                 : "memory");
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p + 4),
                 "f"(x.v[4]), "f"(x.v[5]), "f"(x.v[6]), "f"(x.v[7])
                 : "memory");
  } else if constexpr (VEC == 4) {
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(x.v[0]),
                 "f"(x.v[1]), "f"(x.v[2]), "f"(x.v[3])
                 : "memory");
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i)
      asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(p + i), "f"(x.v[i]) : "memory");
  }
}

// ---- system-scope flag primitives (peer-mapped signal pads) -----------------------------
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Signal pad layout: [channel][kMaxPeers] uint32 epochs, slot j of rank r's pad is written by
// rank j only.
constexpr int kFlagChannels = 16;
__device__ __forceinline__ uint32_t* flag_slot(void* pad, int channel, int writer) {
  return reinterpret_cast<uint32_t*>(pad) + channel * kMaxPeers + writer;
}

// Bounded spin (watchdog): returns false on timeout instead of hanging the GPU.
__device__ __forceinline__ bool wait_flag_ge(const uint32_t* p, uint32_t epoch,
                                             unsigned long long timeout_cycles) {
  unsigned long long start = clock64();
  unsigned spins = 0;
  while (static_cast<int32_t>(ld_acquire_sys(p) - epoch) < 0) {
    if ((++spins & 0x3ff) == 0 && timeout_cycles && (clock64() - start) > timeout_cycles)
      return false;
    __nanosleep(20);
  }
  return true;
}

}  // namespace de
