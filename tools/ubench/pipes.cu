// Micro-benchmark of the sm_100a pipes that bound the fused OF kernel: scalar vs packed
// fp32 (FFMA / FFMA2 / FADD2), fp64 (DFMA / DADD), cvt f64->f32, shared-memory LDS/STS,
// SHFL.  Prints warp-instructions per clock per SM and lane-ops per clock per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/pipes tools/ubench/pipes.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define ITERS 2048
#define NACC 8

enum { K_FFMA, K_FFMA2, K_FADD2, K_FMUL2, K_DFMA, K_DADD, K_CVT, K_LDS64, K_LDS128, K_STS64, K_SHFL, K_FFMA_FADD, K_FADD,
       K_FFMA2_LDS, K_COUNT };
const char* kname[] = {"FFMA", "FFMA2(f32x2)", "FADD2(f32x2)", "FMUL2(f32x2)", "DFMA", "DADD", "CVT.F32.F64", "LDS.64",
                       "LDS.128", "STS.64", "SHFL.BFLY", "FFMA+FADD", "FADD", "FFMA2+LDS.64(1:1)"};

__device__ __forceinline__ float2 lds64(unsigned a) { float2 v; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ float4 lds128(unsigned a) { float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ void sts64(unsigned a, float x, float y) { asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(a), "f"(x), "f"(y) : "memory"); }

template <int K> __global__ void __launch_bounds__(1024) bench(float* out, long long* clk, float a, float b) {
    extern __shared__ float4 sm[];
    const int tid = threadIdx.x;
    float x[NACC], y[NACC];
    double d[NACC];
    unsigned long long p[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
        x[i] = a * (tid + i);
        y[i] = b * (tid - i);
        d[i] = (double)x[i];
        p[i] = ((unsigned long long)__float_as_uint(x[i]) << 32) | __float_as_uint(y[i]);
    }
    const unsigned long long pa = ((unsigned long long)__float_as_uint(a) << 32) | __float_as_uint(a);
    const unsigned long long pb = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(b);
    sm[tid] = make_float4(a, b, a, b);
    __syncthreads();
    const double da = a, db = b;
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(sm);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (K == K_FFMA) x[i] = fmaf(x[i], a, b);
            if (K == K_FADD) x[i] = x[i] + a;
            if (K == K_FFMA_FADD) { x[i] = fmaf(x[i], a, b); y[i] = y[i] + a; }
            if (K == K_FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
            if (K == K_FADD2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pa));
            if (K == K_FMUL2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pa));
            if (K == K_DFMA) d[i] = fma(d[i], da, db);
            if (K == K_DADD) d[i] = d[i] + da;
            if (K == K_CVT) { x[i] += (float)d[i]; d[i] = __longlong_as_double(__double_as_longlong(d[i]) ^ (it & 1)); }
            if (K == K_LDS64) { float2 v = lds64(sbase + 8 * ((tid + i * 32 + it) & 1023)); x[i] += v.x; }
            if (K == K_LDS128) { float4 v = lds128(sbase + 16 * ((tid + i * 32 + it) & 1023)); x[i] += v.x; }
            if (K == K_STS64) { sts64(sbase + 8 * ((tid + i * 32 + it) & 1023), x[i], y[i]); }
            if (K == K_SHFL) x[i] = __shfl_xor_sync(0xffffffffu, x[i], 1 + (i & 15));
            if (K == K_FFMA2_LDS) {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
                float2 v = lds64(sbase + 8 * ((tid + i * 32 + it) & 1023));
                y[i] = v.x;
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += x[i] + y[i] + (float)d[i] + __uint_as_float((unsigned)(p[i] >> 32)) + __uint_as_float((unsigned)p[i]);
    out[blockIdx.x * blockDim.x + tid] = s;
    if (tid == 0) clk[blockIdx.x] = t1 - t0;
}

template <int K> void run(int threads, int sms) {
    float* out;
    long long* clk;
    cudaMalloc(&out, sizeof(float) * threads * sms);
    cudaMalloc(&clk, sizeof(long long) * sms);
    cudaFuncSetAttribute(bench<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    for (int rep = 0; rep < 2; ++rep) bench<K><<<sms, threads, 16384>>>(out, clk, 1.0001f, 0.5f);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<K><<<sms, threads, 16384>>>(out, clk, 1.0001f, 0.5f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[256];
    cudaMemcpy(h, clk, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double c = 0;
    for (int i = 0; i < sms; ++i) c += (double)h[i];
    c /= sms;
    const double winst = (double)ITERS * NACC * (threads / 32);
    printf("%-20s threads=%4d  cycles=%9.0f  warp-inst/clk/SM=%6.3f  lane-ops/clk/SM=%7.2f  (%.3f ms)  err=%s\n", kname[K], threads, c,
           winst / c, winst * 32 / c, ms, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
    cudaFree(clk);
}

int main() {
    cudaDeviceProp pr;
    cudaGetDeviceProperties(&pr, 0);
    printf("%s SMs=%d\n", pr.name, pr.multiProcessorCount);
    const int sms = pr.multiProcessorCount;
    for (int threads : {256, 512, 1024}) {
        run<K_FFMA>(threads, sms);
        run<K_FADD>(threads, sms);
        run<K_FFMA_FADD>(threads, sms);
        run<K_FFMA2>(threads, sms);
        run<K_FADD2>(threads, sms);
        run<K_FMUL2>(threads, sms);
        run<K_DFMA>(threads, sms);
        run<K_DADD>(threads, sms);
        run<K_CVT>(threads, sms);
        run<K_LDS64>(threads, sms);
        run<K_LDS128>(threads, sms);
        run<K_STS64>(threads, sms);
        run<K_SHFL>(threads, sms);
        run<K_FFMA2_LDS>(threads, sms);
    }
    return 0;
}
