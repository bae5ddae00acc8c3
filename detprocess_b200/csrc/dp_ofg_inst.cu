// Instantiations of the mixed-radix OF kernel (trace lengths that are not powers of two), float64 and float32.
#include <cuda_runtime.h>

#include "dp_ofg_kernel.cuh"
#include "dp_ofg_launch.hpp"

namespace {
template <class T> int setup_t(int M, int device, size_t* smem, int* grid_max) {
    *smem = DpGenKernel<T>::smem_bytes(M);
    cudaError_t e = cudaFuncSetAttribute(dp_ofg_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*smem);
    if (e != cudaSuccess) return (int)e;
    int occ = 0, sms = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dp_ofg_kernel<T>, DPG_NT, *smem);
    if (e != cudaSuccess) return (int)e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return (int)e;
    if (occ < 1) return -2;
    *grid_max = sms * occ;
    return 0;
}
}  // namespace

int dp_ofg_setup(int prec, int M, int device, size_t* smem, int* grid_max) {
    return prec == 0 ? setup_t<double>(M, device, smem, grid_max) : setup_t<float>(M, device, smem, grid_max);
}
int dp_ofg_launch(int prec, const void* prm, int grid, size_t smem, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (prec == 0)
        dp_ofg_kernel<double><<<grid, DPG_NT, smem, st>>>(*reinterpret_cast<const DpGenParams<double>*>(prm));
    else
        dp_ofg_kernel<float><<<grid, DPG_NT, smem, st>>>(*reinterpret_cast<const DpGenParams<float>*>(prm));
    return (int)cudaGetLastError();
}
