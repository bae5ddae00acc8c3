// Platform layer: the kernel bodies in this directory are written once and compile
//  (a) with nvcc for sm_100a (the product), and
//  (b) with g++ against a host-thread CTA emulator (tests/emu: one std::thread per
//      CUDA thread, std::barrier for __syncthreads, slot exchange for shuffles),
//      which exists only so index maths / barrier placement can be checked (also
//      under -fsanitize=thread) without a GPU.  The emulator is never a product path.
#pragma once

#ifdef DP_HOST_EMU
// ------------------------------------------------------------------ host emulation
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#define DP_DEV inline
#define DP_HD inline
#define DP_GLOBAL inline
#define DP_RESTRICT __restrict__

struct float2 { float x, y; };
struct double2 { double x, y; };
struct short2 { short x, y; };
struct int2 { int x, y; };
struct uint4 { unsigned x, y, z, w; };
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
static inline double2 make_double2(double a, double b) { return double2{a, b}; }
struct dp_dim3 { unsigned x = 1, y = 1, z = 1; };

namespace dpemu {
struct Cta {
    int nthreads;
    std::unique_ptr<std::barrier<>> bar;
    std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
    std::vector<uint64_t> slots;  // [nthreads]
    std::map<std::pair<int, int>, std::unique_ptr<std::barrier<>>> named;  // bar.sync (id, count), created on first use
    std::mutex named_mu;
    explicit Cta(int nt) : nthreads(nt), slots(nt) {
        bar = std::make_unique<std::barrier<>>(nt);
        for (int w = 0; w < (nt + 31) / 32; ++w) {
            int lanes = std::min(32, nt - 32 * w);
            warp_bar.push_back(std::make_unique<std::barrier<>>(lanes));
        }
    }
};
extern thread_local Cta* cta;
extern thread_local dp_dim3 tIdx, bIdx, bDim, gDim;
}  // namespace dpemu

#define threadIdx (dpemu::tIdx)
#define blockIdx (dpemu::bIdx)
#define blockDim (dpemu::bDim)
#define gridDim (dpemu::gDim)

static inline void __syncthreads() { dpemu::cta->bar->arrive_and_wait(); }
static inline void __syncwarp() { dpemu::cta->warp_bar[dpemu::tIdx.x / 32]->arrive_and_wait(); }
// named barrier: `count` threads of the CTA meet at barrier `id` (1..15)
static inline void dp_bar_sync(int id, int count) {
    auto* c = dpemu::cta;
    std::barrier<>* b;
    {
        std::lock_guard<std::mutex> g(c->named_mu);
        auto& slot = c->named[{id, count}];
        if (!slot) slot = std::make_unique<std::barrier<>>(count);
        b = slot.get();
    }
    b->arrive_and_wait();
}

// producer side of a named barrier: counts the thread in, does not wait
static inline void dp_bar_arrive(int id, int count) {
    auto* c = dpemu::cta;
    std::barrier<>* b;
    {
        std::lock_guard<std::mutex> g(c->named_mu);
        auto& slot = c->named[{id, count}];
        if (!slot) slot = std::make_unique<std::barrier<>>(count);
        b = slot.get();
    }
    (void)b->arrive();
}
static inline int __ffs(int v) { return __builtin_ffs(v); }

template <class V>
static inline V dp_shfl_impl(V v, int src_lane) {
    static_assert(sizeof(V) <= 8, "shuffle payload");
    auto* c = dpemu::cta;
    int tid = dpemu::tIdx.x, w = tid / 32;
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(V));
    c->slots[tid] = raw;
    c->warp_bar[w]->arrive_and_wait();
    int lanes = std::min(32, c->nthreads - 32 * w);
    int s = src_lane;
    uint64_t got = (s >= 0 && s < lanes) ? c->slots[32 * w + s] : raw;
    c->warp_bar[w]->arrive_and_wait();
    V out;
    std::memcpy(&out, &got, sizeof(V));
    return out;
}
template <class V>
static inline V __shfl_xor_sync(unsigned, V v, int mask) { return dp_shfl_impl(v, (int)(dpemu::tIdx.x % 32) ^ mask); }
template <class V>
static inline V __shfl_sync(unsigned, V v, int src) { return dp_shfl_impl(v, src); }
template <class V>
static inline V __shfl_down_sync(unsigned, V v, int d) { return dp_shfl_impl(v, (int)(dpemu::tIdx.x % 32) + d); }

template <class V>
static inline V __ldg(const V* p) { return *p; }
static inline void sincospi(double a, double* s, double* c) { *s = std::sin(M_PI * a); *c = std::cos(M_PI * a); }
static inline void sincospif(float a, float* s, float* c) { *s = (float)std::sin(M_PI * (double)a); *c = (float)std::cos(M_PI * (double)a); }
static inline float rsqrtf(float a) { return 1.0f / std::sqrt(a); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }

#else
// ------------------------------------------------------------------------- CUDA
#include <cuda_runtime.h>
#include <cstdint>
#define DP_DEV __device__ __forceinline__
#define DP_HD __host__ __device__ __forceinline__
#define DP_GLOBAL __global__
#define DP_RESTRICT __restrict__
// named barrier: `count` threads (a multiple of 32) of the CTA meet at barrier `id` (1..15)
__device__ __forceinline__ void dp_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
// producer side: counts the thread in (its earlier writes are visible to the threads that bar.sync on `id`), does not wait
__device__ __forceinline__ void dp_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
#endif

// ------------------------------------------------------------- scalar type traits
template <class T> struct dp_vec2;
template <> struct dp_vec2<float> { using type = float2; };
template <> struct dp_vec2<double> { using type = double2; };

template <class T> DP_HD typename dp_vec2<T>::type dp_make2(T a, T b);
template <> DP_HD float2 dp_make2<float>(float a, float b) { return make_float2(a, b); }
template <> DP_HD double2 dp_make2<double>(double a, double b) { return make_double2(a, b); }

DP_HD float dp_fma(float a, float b, float c) { return fmaf(a, b, c); }
DP_HD double dp_fma(double a, double b, double c) { return fma(a, b, c); }
