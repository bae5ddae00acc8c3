// Launch entry point of the band-amplitude kernel (dp_band_inst.cu).
#pragma once
int dp_band_launch(const void* prm, int grid, void* stream);
