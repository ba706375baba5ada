// Forward-mode dual numbers with 6 tangents: used to differentiate the
// closed-form SE(3) exponential (rigid_body.py:59-101) w.r.t. the raw screw
// outputs (w, v) of the warp MLP inside the d(sigma)/dx sweep.
#pragma once
#include <cuda_runtime.h>

namespace nds {

struct Dual6 {
  float v;
  float d[6];
  __device__ __forceinline__ Dual6() {}
  __device__ __forceinline__ explicit Dual6(float x) : v(x) {
#pragma unroll
    for (int i = 0; i < 6; ++i) d[i] = 0.f;
  }
  __device__ __forceinline__ static Dual6 var(float x, int idx) {
    Dual6 r(x);
    r.d[idx] = 1.f;
    return r;
  }
};

__device__ __forceinline__ Dual6 operator+(const Dual6& a, const Dual6& b) {
  Dual6 r; r.v = a.v + b.v;
#pragma unroll
  for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] + b.d[i];
  return r;
}
__device__ __forceinline__ Dual6 operator-(const Dual6& a, const Dual6& b) {
  Dual6 r; r.v = a.v - b.v;
#pragma unroll
  for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] - b.d[i];
  return r;
}
__device__ __forceinline__ Dual6 operator-(const Dual6& a) {
  Dual6 r; r.v = -a.v;
#pragma unroll
  for (int i = 0; i < 6; ++i) r.d[i] = -a.d[i];
  return r;
}
__device__ __forceinline__ Dual6 operator*(const Dual6& a, const Dual6& b) {
  Dual6 r; r.v = a.v * b.v;
#pragma unroll
  for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
  return r;
}
__device__ __forceinline__ Dual6 operator*(const Dual6& a, float b) {
  Dual6 r; r.v = a.v * b;
#pragma unroll
  for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] * b;
  return r;
}
__device__ __forceinline__ Dual6 operator/(const Dual6& a, const Dual6& b) {
  Dual6 r; r.v = a.v / b.v;
  const float inv = 1.f / b.v;
#pragma unroll
  for (int i = 0; i < 6; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
  return r;
}
__device__ __forceinline__ Dual6 nsqrt(const Dual6& a) {
  Dual6 r; r.v = sqrtf(a.v);
  const float k = 0.5f / r.v;
#pragma unroll
  for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] * k;
  return r;
}
__device__ __forceinline__ Dual6 nsin(const Dual6& a) {
  Dual6 r; r.v = sinf(a.v);
  const float c = cosf(a.v);
#pragma unroll
  for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] * c;
  return r;
}
__device__ __forceinline__ Dual6 ncos(const Dual6& a) {
  Dual6 r; r.v = cosf(a.v);
  const float s = -sinf(a.v);
#pragma unroll
  for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] * s;
  return r;
}


// the same with N tangents (the tensor-core reverse sweep differentiates w.r.t. the 3 rotation inputs only; the
// translation inputs enter linearly)
template <int N>
struct DualN {
  float v;
  float d[N];
  __device__ __forceinline__ DualN() {}
  __device__ __forceinline__ explicit DualN(float x) : v(x) {
#pragma unroll
    for (int i = 0; i < N; ++i) d[i] = 0.f;
  }
  __device__ __forceinline__ static DualN var(float x, int idx) {
    DualN r(x);
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = (i == idx) ? 1.f : 0.f;
    return r;
  }
};
#define NDS_DUALN_LOOP _Pragma("unroll") for (int i = 0; i < N; ++i)
template <int N> __device__ __forceinline__ DualN<N> operator+(const DualN<N>& a, const DualN<N>& b) { DualN<N> r; r.v = a.v + b.v; NDS_DUALN_LOOP r.d[i] = a.d[i] + b.d[i]; return r; }
template <int N> __device__ __forceinline__ DualN<N> operator-(const DualN<N>& a, const DualN<N>& b) { DualN<N> r; r.v = a.v - b.v; NDS_DUALN_LOOP r.d[i] = a.d[i] - b.d[i]; return r; }
template <int N> __device__ __forceinline__ DualN<N> operator-(const DualN<N>& a) { DualN<N> r; r.v = -a.v; NDS_DUALN_LOOP r.d[i] = -a.d[i]; return r; }
template <int N> __device__ __forceinline__ DualN<N> operator*(const DualN<N>& a, const DualN<N>& b) { DualN<N> r; r.v = a.v * b.v; NDS_DUALN_LOOP r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
template <int N> __device__ __forceinline__ DualN<N> operator*(const DualN<N>& a, float b) { DualN<N> r; r.v = a.v * b; NDS_DUALN_LOOP r.d[i] = a.d[i] * b; return r; }
template <int N> __device__ __forceinline__ DualN<N> operator/(const DualN<N>& a, const DualN<N>& b) {
  DualN<N> r; r.v = a.v / b.v;
  const float inv = 1.f / b.v;
  NDS_DUALN_LOOP r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
  return r;
}
template <int N> __device__ __forceinline__ DualN<N> nsqrt(const DualN<N>& a) { DualN<N> r; r.v = sqrtf(a.v); const float k = 0.5f / r.v; NDS_DUALN_LOOP r.d[i] = a.d[i] * k; return r; }
template <int N> __device__ __forceinline__ DualN<N> nsin(const DualN<N>& a) { DualN<N> r; r.v = sinf(a.v); const float c = cosf(a.v); NDS_DUALN_LOOP r.d[i] = a.d[i] * c; return r; }
template <int N> __device__ __forceinline__ DualN<N> ncos(const DualN<N>& a) { DualN<N> r; r.v = cosf(a.v); const float sn = -sinf(a.v); NDS_DUALN_LOOP r.d[i] = a.d[i] * sn; return r; }
#undef NDS_DUALN_LOOP

}  // namespace nds
