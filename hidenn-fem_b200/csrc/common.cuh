// Shared helpers for the sm_100a kernels of the HiDeNN-FEM quadrature hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

namespace hidenn {

void set_error(const std::string& msg);

#define HIDENN_CUDA_OK(expr)                                                                   \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            hidenn::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));             \
            return 1;                                                                          \
        }                                                                                      \
    } while (0)

#define HIDENN_REQUIRE(cond, msg)                                                              \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            hidenn::set_error(std::string(msg));                                               \
            return 2;                                                                          \
        }                                                                                      \
    } while (0)

// Binds the calling thread to a plan's device for the duration of a C-ABI call and restores the previous device
// (a model on cuda:1 must work while the caller's current device is cuda:0, like the reference's torch ops).
struct DeviceScope {
    int prev = -1;
    bool switched = false;
    cudaError_t enter(int dev) {
        cudaError_t e = cudaGetDevice(&prev);
        if (e != cudaSuccess || prev == dev) return e;
        e = cudaSetDevice(dev);
        switched = (e == cudaSuccess);
        return e;
    }
    ~DeviceScope() { if (switched) cudaSetDevice(prev); }
};
constexpr int kMaxDevices = 64;
// SM count of a device, cached per device id
inline int sm_count(int dev) {
    static int cache[kMaxDevices] = {};
    if (dev < 0 || dev >= kMaxDevices) return 148;
    if (cache[dev] == 0) {
        int n = 148;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        cache[dev] = n;
    }
    return cache[dev];
}

template <typename R> struct Real2;
template <> struct Real2<double> { using type = double2; };
template <> struct Real2<float> { using type = float2; };

template <typename R> __device__ __forceinline__ typename Real2<R>::type mk2(R a, R b);
template <> __device__ __forceinline__ double2 mk2<double>(double a, double b) { return make_double2(a, b); }
template <> __device__ __forceinline__ float2 mk2<float>(float a, float b) { return make_float2(a, b); }

__device__ __forceinline__ double rcp(double x) { return __drcp_rn(x); }
__device__ __forceinline__ float rcp(float x) { return __frcp_rn(x); }

template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Fixed-order block sum (deterministic): warp shuffles, then warp 0 folds the per-warp partials.
template <typename T, int BLOCK> __device__ __forceinline__ T block_sum(T v, T* s_warp /*[BLOCK/32]*/) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) s_warp[w] = v;
    __syncthreads();
    T r = T(0);
    if (w == 0) {
        r = (lane < BLOCK / 32) ? s_warp[lane] : T(0);
        r = warp_sum(r);
    }
    return r;   // valid in warp 0
}

// slot maps: s >= 0 -> row of the free Parameter, s < 0 -> row ~s of the fixed buffer
template <typename V> __device__ __forceinline__ V load_slot(const V* __restrict__ free_v, const V* __restrict__ fixed_v, int s) {
    return s >= 0 ? __ldg(free_v + s) : __ldg(fixed_v + (~s));
}

}  // namespace hidenn
