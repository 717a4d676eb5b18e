// common.cuh -- shared helpers for the sm_100a space-charge kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace scb {

template <typename T> struct cx_of;
template <> struct cx_of<float>  { using type = float2; };
template <> struct cx_of<double> { using type = double2; };
template <typename T> using cx_t = typename cx_of<T>::type;

template <typename C> struct real_of;
template <> struct real_of<float2>  { using type = float; };
template <> struct real_of<double2> { using type = double; };

template <typename C> __host__ __device__ __forceinline__ C cmake(typename real_of<C>::type a, typename real_of<C>::type b) {
    C r; r.x = a; r.y = b; return r;
}
template <typename C> __device__ __forceinline__ C cadd(C a, C b) { return cmake<C>(a.x + b.x, a.y + b.y); }
template <typename C> __device__ __forceinline__ C csub(C a, C b) { return cmake<C>(a.x - b.x, a.y - b.y); }
template <typename C> __device__ __forceinline__ C cmul(C a, C b) {
    return cmake<C>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
template <typename C> __device__ __forceinline__ C cconj(C a) { return cmake<C>(a.x, -a.y); }
template <typename C> __device__ __forceinline__ C cscale(C a, typename real_of<C>::type s) { return cmake<C>(a.x * s, a.y * s); }
// multiply by the direction's quarter-turn: DIR=-1 (forward) -> *(-i); DIR=+1 (inverse) -> *(+i)
template <int DIR, typename C> __device__ __forceinline__ C cmul_qturn(C a) {
    return DIR < 0 ? cmake<C>(a.y, -a.x) : cmake<C>(-a.y, a.x);
}
// twiddle in the direction of the transform (table holds forward roots exp(-2 pi i k/N))
template <int DIR, typename C> __device__ __forceinline__ C tw_dir(C w) { return DIR < 0 ? w : cconj(w); }

__host__ __device__ __forceinline__ int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// streaming (read-once) loads and stores: keep L1 for the gathers/twiddles
template <typename V> __device__ __forceinline__ V ld_stream(const V* p) { return __ldcs(p); }
template <typename V> __device__ __forceinline__ void st_stream(V* p, V v) { __stcs(p, v); }

}  // namespace scb
