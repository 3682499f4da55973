// Shared device/host helpers for the pseudo-speaker VAE hot path (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#ifndef PSVAE_NUM_SMS
#define PSVAE_NUM_SMS 148   // B200: 2 dies x 74 SMs; grids are sized in multiples of this
#endif

namespace psvae {

typedef __nv_bfloat16 bf16;

// error plumbing (api.cu)
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
#define PSVAE_CUDA(call)                                                     \
  do {                                                                       \
    cudaError_t _e = (call);                                                 \
    if (_e != cudaSuccess) return ::psvae::cuda_fail(_e, #call);             \
  } while (0)
#define PSVAE_LAUNCH_CHECK(what)                                             \
  do {                                                                       \
    cudaError_t _e = cudaGetLastError();                                     \
    if (_e != cudaSuccess) return ::psvae::cuda_fail(_e, what);              \
  } while (0)
#define PSVAE_TRY(call)                                                      \
  do {                                                                       \
    int _r = (call);                                                         \
    if (_r != 0) return _r;                                                  \
  } while (0)

// Programmatic dependent launch (PDL): every kernel of the step is launched with programmaticStreamSerialization, so its CTAs may be
// scheduled (and run their prologue) while the previous kernel drains; PSVAE_GRID_DEP() at the top of a kernel lets ITS successor do the
// same -- but only AFTER it has itself waited for its predecessor, so that a kernel that starts early knows: everything up to its
// predecessor's predecessor is complete and visible (the GEMM engine relies on this to preload weights).  No-ops when launched normally.
#define PSVAE_GRID_DEP()                                                   \
  do {                                                                     \
    asm volatile("griddepcontrol.wait;" ::: "memory");                     \
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");        \
  } while (0)

// launch with the PDL attribute (option "pdl", default on) -- kernels launched this way MUST start with PSVAE_GRID_DEP()
bool pdl_enabled();
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_dep(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t align_up64(int64_t a, int64_t b) { return ceil_div64(a, b) * b; }

// ---- element conversion -------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);   // .x = lo (low 16 bits)
  return *reinterpret_cast<uint32_t*>(&t);
}

// store NV consecutive values starting at ptr (16-byte aligned when NV*sizeof(T) >= 16)
template <int NV> __device__ __forceinline__ void store_vec(float* ptr, const float* v) {
  if constexpr (NV % 4 == 0) {
#pragma unroll
    for (int i = 0; i < NV; i += 4) *reinterpret_cast<float4*>(ptr + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i) ptr[i] = v[i];
  }
}
template <int NV> __device__ __forceinline__ void store_vec(bf16* ptr, const float* v) {
  if constexpr (NV % 8 == 0) {
#pragma unroll
    for (int i = 0; i < NV; i += 8) {
      uint4 u;
      u.x = pack_bf16x2(v[i], v[i + 1]);
      u.y = pack_bf16x2(v[i + 2], v[i + 3]);
      u.z = pack_bf16x2(v[i + 4], v[i + 5]);
      u.w = pack_bf16x2(v[i + 6], v[i + 7]);
      *reinterpret_cast<uint4*>(ptr + i) = u;
    }
  } else if constexpr (NV % 4 == 0) {
#pragma unroll
    for (int i = 0; i < NV; i += 4) {
      uint2 u;
      u.x = pack_bf16x2(v[i], v[i + 1]);
      u.y = pack_bf16x2(v[i + 2], v[i + 3]);
      *reinterpret_cast<uint2*>(ptr + i) = u;
    }
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i) ptr[i] = __float2bfloat16_rn(v[i]);
  }
}
template <int NV> __device__ __forceinline__ void load_vec(const float* ptr, float* v) {
  if constexpr (NV % 4 == 0) {
#pragma unroll
    for (int i = 0; i < NV; i += 4) {
      float4 t = *reinterpret_cast<const float4*>(ptr + i);
      v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = ptr[i];
  }
}
template <int NV> __device__ __forceinline__ void load_vec(const bf16* ptr, float* v) {
  if constexpr (NV % 8 == 0) {
#pragma unroll
    for (int i = 0; i < NV; i += 8) {
      uint4 u = *reinterpret_cast<const uint4*>(ptr + i);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[i + 2 * j] = __uint_as_float(w[j] << 16);
        v[i + 2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u);
      }
    }
  } else if constexpr (NV % 4 == 0) {
#pragma unroll
    for (int i = 0; i < NV; i += 4) {
      uint2 u = *reinterpret_cast<const uint2*>(ptr + i);
      v[i] = __uint_as_float(u.x << 16); v[i + 1] = __uint_as_float(u.x & 0xFFFF0000u);
      v[i + 2] = __uint_as_float(u.y << 16); v[i + 3] = __uint_as_float(u.y & 0xFFFF0000u);
    }
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = __bfloat162float(ptr[i]);
  }
}

// ---- reductions ---------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Sum over a block of up to 1024 threads; result valid in thread 0. `scratch` >= 32 floats of smem.
__device__ __forceinline__ float block_sum(float v, float* scratch) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float r = 0.f;
  if (w == 0) {
    r = lane < nw ? scratch[lane] : 0.f;
    r = warp_sum(r);
  }
  return r;
}

}  // namespace psvae
