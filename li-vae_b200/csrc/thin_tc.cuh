// Tensor-core (tcgen05) versions of the thin 1-channel layers (see thin_tc.cu).  Each launcher returns
// 0 = launched, 1 = shape not eligible (the caller falls back to the SIMT kernel in thin.cu), or an error.
#pragma once
#include "common.cuh"

namespace livae {
namespace tc {

int thin_tc_conv1c_fwd(int kind, const float* img, const float* w, const float* bias, int B, int H, int W,
                       void* out_bf16, uint8_t* pool_idx, cudaStream_t st);
int thin_tc_col2im(int kind, const void* x_bf16, const float* w, const float* bias, int B, int Hin, int Win, int act,
                   float* out, cudaStream_t st);
int thin_tc_wgrad(int kind, const float* src, const void* big_bf16, const uint8_t* pool_idx, int B, int H, int W,
                  float* gw, float* gb, cudaStream_t st);

}  // namespace tc
}  // namespace livae
