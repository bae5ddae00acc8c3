// Launch entry points of the mixed-radix OF kernel (dp_ofg_inst.cu); prec 0 = float64, 1 = float32.
#pragma once
#include <cstddef>
int dp_ofg_setup(int prec, int M, int device, size_t* smem, int* grid_max);
int dp_ofg_launch(int prec, const void* prm, int grid, size_t smem, void* stream);
