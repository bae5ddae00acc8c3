// Launch entry points of the v2 OF kernel instantiation units (dp_of2_inst.cu).
// p0 = float64, p1 = packed float32; second index = input type (0 f64, 1 f32, 2 i16 with rows aligned to a sample PAIR;
// 3 f64, 4 f32, 5 i16 with rows aligned to one sample only: windows of a continuous stream).
#pragma once
#include <cstddef>

#define DP_OF2_DECL(P, I)                                                                                     \
    int dp_of2_setup_p##P##_##I(int R1, int device, size_t* smem, int* grid_max, int* occ, int* threads);       \
    int dp_of2_launch_p##P##_##I(int R1, int multi, const void* prm, int grid, size_t smem, void* stream, size_t persist_bytes);
DP_OF2_DECL(0, 0) DP_OF2_DECL(0, 1) DP_OF2_DECL(0, 2) DP_OF2_DECL(0, 3) DP_OF2_DECL(0, 4) DP_OF2_DECL(0, 5)
DP_OF2_DECL(1, 0) DP_OF2_DECL(1, 1) DP_OF2_DECL(1, 2) DP_OF2_DECL(1, 3) DP_OF2_DECL(1, 4) DP_OF2_DECL(1, 5)
#undef DP_OF2_DECL

typedef int (*dp_of2_setup_fn)(int, int, size_t*, int*, int*, int*);
typedef int (*dp_of2_launch_fn)(int, int, const void*, int, size_t, void*, size_t);
static const dp_of2_setup_fn dp_of2_setup_table[2][6] = {
    {dp_of2_setup_p0_0, dp_of2_setup_p0_1, dp_of2_setup_p0_2, dp_of2_setup_p0_3, dp_of2_setup_p0_4, dp_of2_setup_p0_5},
    {dp_of2_setup_p1_0, dp_of2_setup_p1_1, dp_of2_setup_p1_2, dp_of2_setup_p1_3, dp_of2_setup_p1_4, dp_of2_setup_p1_5}};
static const dp_of2_launch_fn dp_of2_launch_table[2][6] = {
    {dp_of2_launch_p0_0, dp_of2_launch_p0_1, dp_of2_launch_p0_2, dp_of2_launch_p0_3, dp_of2_launch_p0_4, dp_of2_launch_p0_5},
    {dp_of2_launch_p1_0, dp_of2_launch_p1_1, dp_of2_launch_p1_2, dp_of2_launch_p1_3, dp_of2_launch_p1_4, dp_of2_launch_p1_5}};

// PSD accumulation on the v2 core; second index = input type (0 f64, 1 f32, 2 i16)
#define DP_PSD2_DECL(P, I)                                                                                          \
    int dp_psd2_setup_p##P##_##I(int R1, int device, size_t* smem, int* grid_max, long long* partial_per_cta);       \
    int dp_psd2_launch_p##P##_##I(int R1, const void* prm, int grid, size_t smem, void* stream);
DP_PSD2_DECL(0, 0) DP_PSD2_DECL(0, 1) DP_PSD2_DECL(0, 2) DP_PSD2_DECL(1, 0) DP_PSD2_DECL(1, 1) DP_PSD2_DECL(1, 2)
#undef DP_PSD2_DECL
typedef int (*dp_psd2_setup_fn)(int, int, size_t*, int*, long long*);
typedef int (*dp_psd2_launch_fn)(int, const void*, int, size_t, void*);
static const dp_psd2_setup_fn dp_psd2_setup_table[2][3] = {{dp_psd2_setup_p0_0, dp_psd2_setup_p0_1, dp_psd2_setup_p0_2},
                                                           {dp_psd2_setup_p1_0, dp_psd2_setup_p1_1, dp_psd2_setup_p1_2}};
static const dp_psd2_launch_fn dp_psd2_launch_table[2][3] = {{dp_psd2_launch_p0_0, dp_psd2_launch_p0_1, dp_psd2_launch_p0_2},
                                                             {dp_psd2_launch_p1_0, dp_psd2_launch_p1_1, dp_psd2_launch_p1_2}};
