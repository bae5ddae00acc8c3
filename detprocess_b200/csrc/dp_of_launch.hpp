// Entry points of the per-(precision, input type) kernel translation units
// (dp_of_inst.cu).  Return 0 on success, a cudaError_t value, or -1 if unsupported.
#pragma once
#include <cstddef>

#define DP_DECL_INST(p, in)                                                                      \
    int dp_of_setup_p##p##_##in(int R1, int P, int device, size_t* smem, int* grid_max, int* occ);      \
    int dp_of_launch_p##p##_##in(int R1, int P, const void* prm, int grid, size_t smem, void* stream);
DP_DECL_INST(0, 0)
DP_DECL_INST(0, 1)
DP_DECL_INST(0, 2)
DP_DECL_INST(1, 0)
DP_DECL_INST(1, 1)
DP_DECL_INST(1, 2)
#undef DP_DECL_INST

typedef int (*dp_of_setup_fn)(int, int, int, size_t*, int*, int*);
typedef int (*dp_of_launch_fn)(int, int, const void*, int, size_t, void*);
static const dp_of_setup_fn dp_of_setup_table[2][3] = {{dp_of_setup_p0_0, dp_of_setup_p0_1, dp_of_setup_p0_2},
                                                       {dp_of_setup_p1_0, dp_of_setup_p1_1, dp_of_setup_p1_2}};
static const dp_of_launch_fn dp_of_launch_table[2][3] = {{dp_of_launch_p0_0, dp_of_launch_p0_1, dp_of_launch_p0_2},
                                                         {dp_of_launch_p1_0, dp_of_launch_p1_1, dp_of_launch_p1_2}};

// PSD accumulation kernels (float64 traces)
int dp_psd_setup_p0_0(int R1, int P, int device, size_t* smem, int* grid_max);
int dp_psd_setup_p1_0(int R1, int P, int device, size_t* smem, int* grid_max);
int dp_psd_launch_p0_0(int R1, int P, const void* prm, int grid, size_t smem, void* stream);
int dp_psd_launch_p1_0(int R1, int P, const void* prm, int grid, size_t smem, void* stream);
int dp_psd_reduce_launch(const void* prm, void* stream);
