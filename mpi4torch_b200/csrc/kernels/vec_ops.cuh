// 16-byte vector arithmetic over the shared Elem/Combine tables, guarded
// private loads/stores (unaligned tensors, ragged tails) and the fused
// epilogue (scale, accumulate) used by every reducing kernel.
#pragma once
#include "../runtime/reduce_ops.h"
#include "device_sync.cuh"

namespace m4t {

struct DevEpilogue {
  double scale_d;
  float scale_f;
  int has_scale;
  const void* acc;  // same dtype/indexing as the output, or nullptr
};

template <typename A> struct ScaleAcc {
  static __device__ __forceinline__ A apply(A v, const DevEpilogue& e) {
    return static_cast<A>(static_cast<double>(v) * e.scale_d);
  }
};
template <> struct ScaleAcc<float> {
  static __device__ __forceinline__ float apply(float v, const DevEpilogue& e) { return v * e.scale_f; }
};
template <> struct ScaleAcc<double> {
  static __device__ __forceinline__ double apply(double v, const DevEpilogue& e) { return v * e.scale_d; }
};

template <DType DT> struct VecOf {
  using E = Elem<DT>;
  using S = typename E::storage;
  using A = typename E::acc;
  static constexpr int N = 16 / static_cast<int>(sizeof(S));
  union U {
    Vec16 v;
    S s[N];
  };
  static __device__ __forceinline__ void unpack(const Vec16& v, A (&a)[N]) {
    U u;
    u.v = v;
#pragma unroll
    for (int i = 0; i < N; ++i) a[i] = E::load(u.s[i]);
  }
  static __device__ __forceinline__ Vec16 pack(const A (&a)[N]) {
    U u;
#pragma unroll
    for (int i = 0; i < N; ++i) u.s[i] = E::store(a[i]);
    return u.v;
  }
};

// Loads vector `i` (16 bytes) of a private tensor with `n` elements.  The
// fast path needs a 16-byte aligned base and a full vector.
template <DType DT>
__device__ __forceinline__ Vec16 load_private(const void* base, int64_t i, int64_t n, bool aligned) {
  using V = VecOf<DT>;
  const int64_t e0 = i * V::N;
  if (aligned && e0 + V::N <= n) return ld_vec_stream(static_cast<const char*>(base) + i * 16);
  typename V::U u;
  const typename V::S* p = static_cast<const typename V::S*>(base);
#pragma unroll
  for (int k = 0; k < V::N; ++k) u.s[k] = (e0 + k < n) ? p[e0 + k] : typename V::S(0);
  return u.v;
}

template <DType DT>
__device__ __forceinline__ void store_private(void* base, int64_t i, int64_t n, bool aligned, const Vec16& v) {
  using V = VecOf<DT>;
  const int64_t e0 = i * V::N;
  if (aligned && e0 + V::N <= n) {
    st_vec(static_cast<char*>(base) + i * 16, v);
    return;
  }
  typename V::U u;
  u.v = v;
  typename V::S* p = static_cast<typename V::S*>(base);
#pragma unroll
  for (int k = 0; k < V::N; ++k)
    if (e0 + k < n) p[e0 + k] = u.s[k];
}

// acc[] (op)= unpack(v)
template <DType DT, ReduceOp OP>
__device__ __forceinline__ void combine_into(typename VecOf<DT>::A (&acc)[VecOf<DT>::N], const Vec16& v) {
  using V = VecOf<DT>;
  using C = Combine<OP, typename V::A, V::E::is_float>;
  typename V::A b[V::N];
  V::unpack(v, b);
#pragma unroll
  for (int k = 0; k < V::N; ++k) acc[k] = C::apply(acc[k], b[k]);
}

template <DType DT, ReduceOp OP>
__device__ __forceinline__ void init_from(typename VecOf<DT>::A (&acc)[VecOf<DT>::N], const Vec16& v) {
  using V = VecOf<DT>;
  V::unpack(v, acc);
#pragma unroll
  for (int k = 0; k < V::N; ++k) acc[k] = normalise_single<OP, typename V::A>(acc[k]);
}

template <DType DT>
__device__ __forceinline__ void apply_scale(typename VecOf<DT>::A (&acc)[VecOf<DT>::N], const DevEpilogue& e) {
  using V = VecOf<DT>;
  if (e.has_scale) {
#pragma unroll
    for (int k = 0; k < V::N; ++k) acc[k] = ScaleAcc<typename V::A>::apply(acc[k], e);
  }
}

// acc[] += private accumulate tensor (vector i)
template <DType DT>
__device__ __forceinline__ void apply_accumulate(typename VecOf<DT>::A (&acc)[VecOf<DT>::N], const DevEpilogue& e,
                                                 int64_t i, int64_t n, bool aligned) {
  using V = VecOf<DT>;
  if (e.acc) {
    typename V::A b[V::N];
    V::unpack(load_private<DT>(e.acc, i, n, aligned), b);
#pragma unroll
    for (int k = 0; k < V::N; ++k) acc[k] = acc[k] + b[k];
  }
}

// acc[] += unpack(v)   (the accumulate tensor's vector, loaded by the caller ahead of use)
template <DType DT>
__device__ __forceinline__ void add_vec(typename VecOf<DT>::A (&acc)[VecOf<DT>::N], const Vec16& v) {
  using V = VecOf<DT>;
  typename V::A b[V::N];
  V::unpack(v, b);
#pragma unroll
  for (int k = 0; k < V::N; ++k) acc[k] = acc[k] + b[k];
}

__host__ inline DevEpilogue make_dev_epilogue(const Epilogue& e) {
  DevEpilogue d;
  d.scale_d = e.scale;
  d.scale_f = static_cast<float>(e.scale);
  d.has_scale = e.has_scale ? 1 : 0;
  d.acc = e.accumulate;
  return d;
}

}  // namespace m4t
