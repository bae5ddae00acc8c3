// Packed pair of fp32 values processed by one sm_100a FFMA2 / FADD2 / FMUL2 instruction
// (PTX fma.rn.f32x2 etc., CUDA 12.9 intrinsics __ffma2_rn / __fadd2_rn / __fmul2_rn).
// Measured on B200 (tools/ubench/pipes.cu): scalar FFMA issues at 2.6 warp-inst/clk/SM
// (82 lane-FMA/clk), FFMA2 at 1.97 warp-inst/clk/SM = 126 lane-FMA/clk -- the packed form
// is the only way to reach the fp32 peak and it halves the issue-slot cost.  The v2 OF
// kernel instantiates its butterflies on cx<f2>: two complex points per register quad.
#pragma once
#include "dp_platform.cuh"

struct alignas(8) f2 {
    float x, y;
    f2() = default;
    DP_HD constexpr f2(float a) : x(a), y(a) {}
    DP_HD constexpr f2(double a) : x((float)a), y((float)a) {}
    DP_HD constexpr f2(int a) : x((float)a), y((float)a) {}
    DP_HD constexpr f2(float a, float b) : x(a), y(b) {}
};

#ifdef DP_HOST_EMU
DP_HD f2 operator+(f2 a, f2 b) { return f2(a.x + b.x, a.y + b.y); }
DP_HD f2 operator-(f2 a, f2 b) { return f2(a.x - b.x, a.y - b.y); }
DP_HD f2 operator*(f2 a, f2 b) { return f2(a.x * b.x, a.y * b.y); }
DP_HD f2 operator-(f2 a) { return f2(-a.x, -a.y); }
DP_HD f2 dp_fma(f2 a, f2 b, f2 c) { return f2(std::fma(a.x, b.x, c.x), std::fma(a.y, b.y, c.y)); }
#else
__device__ __forceinline__ float2 dp_f2_raw(f2 a) { return make_float2(a.x, a.y); }
__device__ __forceinline__ f2 dp_f2_wrap(float2 a) { return f2(a.x, a.y); }
__device__ __forceinline__ f2 operator+(f2 a, f2 b) { return dp_f2_wrap(__fadd2_rn(dp_f2_raw(a), dp_f2_raw(b))); }
// a - b as fma(b, -1, a): one FFMA2, no separate negation of the two halves
__device__ __forceinline__ f2 operator-(f2 a, f2 b) {
    return dp_f2_wrap(__ffma2_rn(dp_f2_raw(b), make_float2(-1.0f, -1.0f), dp_f2_raw(a)));
}
__device__ __forceinline__ f2 operator*(f2 a, f2 b) { return dp_f2_wrap(__fmul2_rn(dp_f2_raw(a), dp_f2_raw(b))); }
__device__ __forceinline__ f2 operator-(f2 a) { return f2(-a.x, -a.y); }
__device__ __forceinline__ f2 dp_fma(f2 a, f2 b, f2 c) {
    return dp_f2_wrap(__ffma2_rn(dp_f2_raw(a), dp_f2_raw(b), dp_f2_raw(c)));
}
#endif
