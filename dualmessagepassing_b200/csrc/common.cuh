// common.cuh -- shared helpers for the DMPNN sparse-core kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/dmp_b200.h"

namespace dmp {

constexpr int kThreads = 256;   // 8 warps per CTA everywhere
constexpr int kNumSMs = 148;    // B200

void set_error(const char* fmt, ...);
// SMs the persistent (one-CTA-per-SM) tensor-core kernels may occupy: 148 minus dmp_set_sm_reserve()
int persistent_sms();

#define DMP_CHECK_ARG(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      ::dmp::set_error(__VA_ARGS__);        \
      return DMP_ERR_INVALID;               \
    }                                       \
  } while (0)

#define DMP_CUDA_OK(expr)                                                                  \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ::dmp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return DMP_ERR_CUDA;                                                                 \
    }                                                                                      \
  } while (0)

inline int launch_status(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
    return DMP_ERR_CUDA;
  }
  return DMP_OK;
}

// ---- vector access along the hidden dimension ---------------------------------------------------
template <int VEC> struct Vec;
template <> struct Vec<4> { using type = float4; };
template <> struct Vec<2> { using type = float2; };
template <> struct Vec<1> { using type = float; };

template <int VEC>
struct Row {
  float v[VEC];
};

// default-cached load (tables that are re-gathered: Qd/Qs/gN rows)
template <int VEC>
__device__ __forceinline__ Row<VEC> ld_row(const float* p) {
  Row<VEC> r;
  if constexpr (VEC == 4) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  } else if constexpr (VEC == 2) {
    float2 t = __ldg(reinterpret_cast<const float2*>(p));
    r.v[0] = t.x; r.v[1] = t.y;
  } else {
    r.v[0] = __ldg(p);
  }
  return r;
}

// streaming load: data touched exactly once (edge-sized operands) -- do not allocate in L1
template <int VEC>
__device__ __forceinline__ Row<VEC> ld_stream(const float* p) {
  Row<VEC> r;
  if constexpr (VEC == 4) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]) : "l"(p));
  } else if constexpr (VEC == 2) {
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];"
                 : "=f"(r.v[0]), "=f"(r.v[1]) : "l"(p));
  } else {
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r.v[0]) : "l"(p));
  }
  return r;
}

// plain (coherent) load for operands that may alias the output (in-place edge update)
template <int VEC>
__device__ __forceinline__ Row<VEC> ld_plain(const float* p) {
  Row<VEC> r;
  if constexpr (VEC == 4) {
    float4 t = *reinterpret_cast<const float4*>(p);
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  } else if constexpr (VEC == 2) {
    float2 t = *reinterpret_cast<const float2*>(p);
    r.v[0] = t.x; r.v[1] = t.y;
  } else {
    r.v[0] = *p;
  }
  return r;
}

template <int VEC>
__device__ __forceinline__ void st_row(float* p, const Row<VEC>& r) {
  if constexpr (VEC == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  } else if constexpr (VEC == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(r.v[0], r.v[1]);
  } else {
    *p = r.v[0];
  }
}

// streaming store (written once, not re-read by this kernel)
template <int VEC>
__device__ __forceinline__ void st_stream(float* p, const Row<VEC>& r) {
  if constexpr (VEC == 4) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(r.v[0], r.v[1], r.v[2], r.v[3]));
  } else if constexpr (VEC == 2) {
    __stcs(reinterpret_cast<float2*>(p), make_float2(r.v[0], r.v[1]));
  } else {
    __stcs(p, r.v[0]);
  }
}

// ---- how a row of H floats is spread over a group of G lanes ---------------------------------------
// lane l of the group owns vectors  l, l+G, ..., l+(ITER-1)*G  (each VEC floats); H <= G*ITER*VEC.
struct Shape {
  int vec, g, iter;
};

inline bool aligned_to(const void* p, int bytes) {
  return p == nullptr || (reinterpret_cast<uintptr_t>(p) % bytes) == 0;
}

// Largest vector width allowed by H, the leading dimensions and the pointers.
inline int pick_vec(int64_t H, std::initializer_list<int64_t> lds, std::initializer_list<const void*> ptrs) {
  int vec = 4;
  while (vec > 1) {
    bool ok = (H % vec) == 0;
    for (int64_t ld : lds) ok = ok && (ld % vec) == 0;
    for (const void* p : ptrs) ok = ok && aligned_to(p, vec * 4);
    if (ok) break;
    vec >>= 1;
  }
  return vec;
}

// kMaxChunk floats of a row are handled by one launch; wider rows are processed in column chunks.
constexpr int kMaxIter = 4;
inline Shape pick_shape(int64_t Hc, int vec) {
  int64_t nvec = (Hc + vec - 1) / vec;
  Shape s{vec, 32, 1};
  if (nvec <= 8) s.g = 8;
  else if (nvec <= 16) s.g = 16;
  else if (nvec <= 32) s.g = 32;
  else if (nvec <= 64) s.iter = 2;
  else s.iter = 4;
  return s;
}
inline int64_t max_chunk(int vec) { return (int64_t)32 * kMaxIter * vec; }

__device__ __forceinline__ float apply_act(float x, int act, float slope) {
  switch (act) {
    case DMP_ACT_RELU: return x > 0.f ? x : 0.f;
    case DMP_ACT_LEAKY_RELU: return x > 0.f ? x : __fmul_rn(x, slope);
    case DMP_ACT_TANH: return tanhf(x);
    case DMP_ACT_SIGMOID: return 1.f / (1.f + expf(-x));
    default: return x;
  }
}
__device__ __forceinline__ float act_grad(float x, int act, float slope) {
  switch (act) {
    case DMP_ACT_RELU: return x > 0.f ? 1.f : 0.f;
    case DMP_ACT_LEAKY_RELU: return x > 0.f ? 1.f : slope;
    case DMP_ACT_TANH: { float t = tanhf(x); return 1.f - t * t; }
    case DMP_ACT_SIGMOID: { float s = 1.f / (1.f + expf(-x)); return s * (1.f - s); }
    default: return 1.f;
  }
}

// derivative of the activation expressed through its OUTPUT y = act(x) (sign(y) == sign(x) for the
// piecewise-linear ones as long as slope > 0)
__device__ __forceinline__ float act_grad_from_output(float y, int act, float slope) {
  switch (act) {
    case DMP_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case DMP_ACT_LEAKY_RELU: return y > 0.f ? 1.f : slope;
    case DMP_ACT_TANH: return 1.f - y * y;
    case DMP_ACT_SIGMOID: return y * (1.f - y);
    default: return 1.f;
  }
}

}  // namespace dmp
