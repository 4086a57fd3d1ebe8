// Host interface of the tensor-core convolution kernels (conv_tc.cu), used by conv.cu.
#pragma once
#include "common.cuh"

namespace d2p {

// geometry of one 3x3 / stride-2 / TF-SAME layer over all frames (see conv.cu)
struct ConvGeo {
    int N, IH, IW, CIN, OH, OW, COUT, PT, PL, T, k;
};

// bit 0: forward, bit 1: input gradient, bit 2: weight gradient on the tensor cores
int conv_tc_mode();
bool conv_tc_supported(const ConvGeo& g);
// bytes of the per-layer scratch (weight images, statistics partials, weight-gradient partials)
size_t conv_tc_ws_bytes(const ConvGeo& g);

// a = lrelu(conv(in*scale+shift) + bias); per-tile (sum, sum of squares) partials of `a` per
// BatchNorm slice land in the scratch and are finalised by bn_forward_finalize (training only)
int conv_tc_fwd(cudaStream_t st, const ConvGeo& g, const float* in, const float* scale, const float* shift,
                const float* W, const float* bias, float* out, int training, int* nchunk, float2** partial,
                void* ws, size_t ws_bytes);
// dX = conv^T(dZ, W)
int conv_tc_dx(cudaStream_t st, const ConvGeo& g, const float* dZ, const float* W, float* dX, void* ws,
               size_t ws_bytes);
// dW += im2col(in*scale+shift)^T dZ
int conv_tc_dw(cudaStream_t st, const ConvGeo& g, const float* in, const float* scale, const float* shift,
               const float* dZ, float* dW, void* ws, size_t ws_bytes);

// RGB input layer with u8 frames (CIN = 3, COUT = 16): dW += im2col(frames)^T dZ on the tensor cores
bool conv_tc_dw3_supported(const ConvGeo& g);
size_t conv_tc_dw3_ws_bytes(const ConvGeo& g);
int conv_tc_dw3(cudaStream_t st, const ConvGeo& g, const uint8_t* in, const float* dZ, float* dW, void* ws,
                size_t ws_bytes);

}  // namespace d2p
