// One translation unit per (precision, input type): compiled several times by
// detprocess_b200/build.py with -DDP_INST_PREC={0,1} -DDP_INST_IN={0,1,2} so the
// kernel instantiations build in parallel.
//
// Geometries built: nb_samples = 1024 * R1 * P.
//   P = 1: R1 = 2..32 (fp32), 2..16 (fp64: M' = 16384 complex f64 exceeds one CTA's smem)
//   P = 2: the largest sub-FFT only (fp32: 65536 samples, fp64: 32768 samples)
#include <cuda_runtime.h>

#include "dp_of_kernel.cuh"
#include "dp_psd_kernel.cuh"
#include "dp_of_launch.hpp"

#if DP_INST_PREC == 0
using InstT = double;
constexpr int kR1Max = 16;
#else
using InstT = float;
constexpr int kR1Max = 32;
#endif
constexpr int kIN = DP_INST_IN;

namespace {

template <int R1, int P> int setup_one(int device, size_t* smem, int* grid_max, int* occ_out) {
    using K = DpOfKernel<InstT, R1, P, kIN>;
    auto kern = dp_of_kernel<InstT, R1, P, kIN>;
    *smem = K::SMEM_BYTES;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, K::NT, K::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    int sms = 0;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return (int)e;
    *grid_max = sms * occ;
    *occ_out = occ;
    return 0;
}

template <int R1, int P> int launch_one(const DpOfParams<InstT>& prm, int grid, size_t smem, cudaStream_t st) {
    dp_of_kernel<InstT, R1, P, kIN><<<grid, DpGeom<R1>::NT, smem, st>>>(prm);
    return (int)cudaGetLastError();
}

}  // namespace

#define DP_CAT2(a, b, c) a##b##_##c
#define DP_CAT(a, b, c) DP_CAT2(a, b, c)

int DP_CAT(dp_of_setup_p, DP_INST_PREC, DP_INST_IN)(int R1, int P, int device, size_t* smem, int* grid_max, int* occ) {
    if (P == 2) return R1 == kR1Max ? setup_one<kR1Max, 2>(device, smem, grid_max, occ) : -1;
    if (P != 1 || R1 > kR1Max) return -1;
    switch (R1) {
        case 2: return setup_one<2, 1>(device, smem, grid_max, occ);
        case 4: return setup_one<4, 1>(device, smem, grid_max, occ);
        case 8: return setup_one<8, 1>(device, smem, grid_max, occ);
        case 16: return setup_one<16, 1>(device, smem, grid_max, occ);
        case 32: return setup_one<(kR1Max >= 32 ? 32 : 16), 1>(device, smem, grid_max, occ);
        default: return -1;
    }
}

int DP_CAT(dp_of_launch_p, DP_INST_PREC, DP_INST_IN)(int R1, int P, const void* prm_v, int grid, size_t smem, void* st_v) {
    const DpOfParams<InstT>& prm = *reinterpret_cast<const DpOfParams<InstT>*>(prm_v);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(st_v);
    if (P == 2) return R1 == kR1Max ? launch_one<kR1Max, 2>(prm, grid, smem, st) : -1;
    if (P != 1 || R1 > kR1Max) return -1;
    switch (R1) {
        case 2: return launch_one<2, 1>(prm, grid, smem, st);
        case 4: return launch_one<4, 1>(prm, grid, smem, st);
        case 8: return launch_one<8, 1>(prm, grid, smem, st);
        case 16: return launch_one<16, 1>(prm, grid, smem, st);
        case 32: return launch_one<(kR1Max >= 32 ? 32 : 16), 1>(prm, grid, smem, st);
        default: return -1;
    }
}

// ------------------------------------------------------------------------ PSD kernels
// built for float64 traces only (IN = 0)
#if DP_INST_IN == 0
namespace {
template <int R1, int P> int psd_setup_one(int device, size_t* smem, int* grid_max) {
    using K = DpPsdKernel<InstT, R1, P, 0>;
    auto kern = dp_psd_kernel<InstT, R1, P, 0>;
    *smem = K::SMEM_BYTES;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    int occ = 0, sms = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, K::NT, K::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return (int)e;
    if (occ < 1) return -2;
    *grid_max = sms * occ;
    return 0;
}
template <int R1, int P> int psd_launch_one(const DpPsdParams<InstT>& prm, int grid, size_t smem, cudaStream_t st) {
    dp_psd_kernel<InstT, R1, P, 0><<<grid, DpGeom<R1>::NT, smem, st>>>(prm);
    return (int)cudaGetLastError();
}
}  // namespace

int DP_CAT(dp_psd_setup_p, DP_INST_PREC, 0)(int R1, int P, int device, size_t* smem, int* grid_max) {
    if (P == 2) return R1 == kR1Max ? psd_setup_one<kR1Max, 2>(device, smem, grid_max) : -1;
    if (P != 1 || R1 > kR1Max) return -1;
    switch (R1) {
        case 2: return psd_setup_one<2, 1>(device, smem, grid_max);
        case 4: return psd_setup_one<4, 1>(device, smem, grid_max);
        case 8: return psd_setup_one<8, 1>(device, smem, grid_max);
        case 16: return psd_setup_one<16, 1>(device, smem, grid_max);
        case 32: return psd_setup_one<(kR1Max >= 32 ? 32 : 16), 1>(device, smem, grid_max);
        default: return -1;
    }
}
int DP_CAT(dp_psd_launch_p, DP_INST_PREC, 0)(int R1, int P, const void* prm_v, int grid, size_t smem, void* st_v) {
    const DpPsdParams<InstT>& prm = *reinterpret_cast<const DpPsdParams<InstT>*>(prm_v);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(st_v);
    if (P == 2) return R1 == kR1Max ? psd_launch_one<kR1Max, 2>(prm, grid, smem, st) : -1;
    if (P != 1 || R1 > kR1Max) return -1;
    switch (R1) {
        case 2: return psd_launch_one<2, 1>(prm, grid, smem, st);
        case 4: return psd_launch_one<4, 1>(prm, grid, smem, st);
        case 8: return psd_launch_one<8, 1>(prm, grid, smem, st);
        case 16: return psd_launch_one<16, 1>(prm, grid, smem, st);
        case 32: return psd_launch_one<(kR1Max >= 32 ? 32 : 16), 1>(prm, grid, smem, st);
        default: return -1;
    }
}
#if DP_INST_PREC == 0
// the (precision independent) reduction kernel lives in exactly one translation unit
__global__ void dp_psd_reduce_kernel(const DpPsdReduceParams prm) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < prm.nbins) {
        const int l = prm.loc[k];
        double s = 0.0;
        for (int c = 0; c < prm.grid; ++c) s += prm.partial[(long long)c * prm.partial_per_cta + l];
        prm.sum_out[k] += s;
    }
    if (k == 0) {
        unsigned long long n = 0;
        for (int c = 0; c < prm.grid; ++c) n += prm.count[c];
        *prm.count_out += n;
    }
}
int dp_psd_reduce_launch(const void* prm_v, void* st_v) {
    const DpPsdReduceParams& prm = *reinterpret_cast<const DpPsdReduceParams*>(prm_v);
    dp_psd_reduce_kernel<<<(prm.nbins + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(st_v)>>>(prm);
    return (int)cudaGetLastError();
}
#endif
#endif
