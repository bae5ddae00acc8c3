// Micro-benchmark: time of one in-place shared-memory FFT pass over 8192 packed vectors (128 KB) per SM,
// radix 16 with 512 threads (16 vectors / thread, 128 registers) vs radix 8 with 1024 threads (8 vectors /
// thread, 64 registers), packed fp32 and fp64.  Answers: would a radix-8 / 1024-thread generation of the OF
// kernel hide the LDS -> FMA -> STS latencies that the 4-warps-per-scheduler kernel exposes?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I detprocess_b200/csrc -o tools/ubench/passes tools/ubench/passes.cu
#include <cstdio>
#include "dp_of2_kernel.cuh"

template <class T, int RAD, int NT> __global__ void __launch_bounds__(NT, 1) passk(cx<T>* g, const cx<T>* tw, int npass, long long* clk) {
    extern __shared__ __align__(16) unsigned char raw[];
    cx<T>* buf = reinterpret_cast<cx<T>*>(raw);
    const int tid = threadIdx.x;
    for (int i = tid; i < 8192 + 1024; i += NT) buf[i] = g[i & 8191];
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < npass; ++it) {
        // alternate between a large-stride and a small-stride pass (both conflict free with v + v/8 padding)
        const int stride = (it & 1) ? 8 : (8192 / RAD);
        const int col = (it & 1) ? ((tid & 7) + (tid >> 3) * 8 * RAD) : tid;
        cx<T> z[RAD];
#pragma unroll
        for (int n = 0; n < RAD; ++n) { const int v = col + n * stride; z[n] = buf[v + (v >> 3)]; }
        dp_dft<RAD, -1, T>::run(z);
        dp_twiddle<RAD, false, T>(z, dp_ldg(tw + (tid & 127)));
#pragma unroll
        for (int n = 0; n < RAD; ++n) { const int v = col + n * stride; buf[v + (v >> 3)] = z[n]; }
        __syncthreads();
    }
    const long long t1 = clock64();
    if (tid == 0) clk[blockIdx.x] = t1 - t0;
    for (int i = tid; i < 8192; i += NT) g[blockIdx.x * 8192 + i] = buf[i + (i >> 3)];
}

template <class T, int RAD, int NT> void run(const char* name) {
    cx<T>*g, *tw;
    long long* clk;
    const int sms = 148, npass = 400;
    cudaMalloc(&g, sizeof(cx<T>) * 8192 * sms);
    cudaMalloc(&tw, sizeof(cx<T>) * 128);
    cudaMalloc(&clk, sizeof(long long) * sms);
    cudaMemset(g, 0, sizeof(cx<T>) * 8192 * sms);
    cx<T> h[128];
    for (int i = 0; i < 128; ++i) { const double a = -6.283185307179586 * i / 4096; h[i] = cx<T>{(T)cos(a), (T)sin(a)}; }
    cudaMemcpy(tw, h, sizeof(h), cudaMemcpyHostToDevice);
    const size_t smem = sizeof(cx<T>) * (8192 + 1024 + 16);
    cudaFuncSetAttribute(passk<T, RAD, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    passk<T, RAD, NT><<<sms, NT, smem>>>(g, tw, npass, clk);
    passk<T, RAD, NT><<<sms, NT, smem>>>(g, tw, npass, clk);
    cudaDeviceSynchronize();
    long long hc[148];
    cudaMemcpy(hc, clk, sizeof(hc), cudaMemcpyDeviceToHost);
    double c = 0;
    for (int i = 0; i < sms; ++i) c += (double)hc[i];
    c /= sms * npass;
    const double levels = RAD == 16 ? 4 : 3;
    printf("%-28s %7.0f clk / pass of 8192 vectors   %7.0f clk per radix-2 level   err=%s\n", name, c, c / levels, cudaGetErrorString(cudaGetLastError()));
    cudaFree(g); cudaFree(tw); cudaFree(clk);
}

int main() {
    run<f2, 16, 512>("fp32x2 radix16 512 thr");
    run<f2, 8, 1024>("fp32x2 radix8 1024 thr");
    run<f2, 8, 512>("fp32x2 radix8  512 thr x2it");
    run<double, 16, 512>("fp64   radix16 512 thr");
    run<double, 8, 1024>("fp64   radix8 1024 thr");
    return 0;
}
