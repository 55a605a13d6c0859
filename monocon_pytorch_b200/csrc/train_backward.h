// Backward kernels of the training step (second half; BASELINE.json configs[2]) -- launchers on plain
// device pointers.  Compiled into the library, driven by the engine (api.cu: mc_backward_train walks the stage list through
// mc_bw_run_graph) and checked formula by formula and over the whole 50-convolution graph: on the CPU under tests/host_shim
// (sequential-thread execution of the same kernel bodies) against oracle/backward_oracle.py, and on the B200
// (tests/test_gpu_zz_train_backward.py, round 2: all green).  Correctness-first fp32 kernels; the tensor-core plan is DESIGN.md 9.
//
// All activations and gradients are fp32 NHWC.  "+=" outputs accumulate into buffers the caller zeroed at the start of the
// backward pass (a tensor with several consumers receives one contribution per consumer, the kernels of one pass run in
// stream order); "=" outputs are overwritten.
#pragma once
#ifndef MC_HOST_SHIM
#include "common.cuh"
#endif

namespace mc {

// y = conv2d(cat(src...), w, stride, pad)  (reference: model/backbone/dla.py:22-31,117-121,228-236; dla_neck.py:24-31;
// model/dense_heads/monocon_heads.py:114-131).
struct ConvBwdParams {
    const float* src[kMaxSrc];   // forward inputs (physical pitch srcWp, x offset srcXoff, like ConvParams)
    float* dsrc[kMaxSrc];        // += gradient of each source, dense NHWC [B,Hin,Win,srcC]; null = not needed (the image)
    int srcC[kMaxSrc], srcWp[kMaxSrc], srcXoff[kMaxSrc];
    int nsrc;
    int B, Hin, Win, Hout, Wout, Cin, Cout, k, stride, pad;
    const float* w;              // [k*k][Cin][Cout] (ConvLayer::w_simt)
    float* wT;                   // scratch for dgrad, k*k*Cin*Cout floats: the weights as [k*k][Cout][Cin] so that the threads of a warp
                                 // (adjacent input channels) read adjacent floats; null = read w with a stride of Cout
    const float* dy;             // gradient of the raw convolution output, NHWC [B,Hout,Wout,Cout]
    float* dw;                   // += [k*k][Cin][Cout]; null = skip
};
void launch_conv_wgrad(const ConvBwdParams& p, cudaStream_t st);
void launch_conv_dgrad(const ConvBwdParams& p, cudaStream_t st);

// z = BN_train(raw) (+ res);  y = relu ? max(z, 0) : z   (nn.BatchNorm2d in train(), dla.py:24,30,119,185,233,295; dla_neck.py:27)
struct BnBwdParams {
    const float* dy;             // gradient of y
    const float* y;              // forward output (only its sign is used, for the ReLU mask); may be null when relu == 0
    const float* raw;            // convolution output the BN normalised
    const float* mean;           // [C] batch mean of raw
    const float* inv;            // [C] rsqrt(biased batch variance + eps)
    const float* gamma;          // [C] or null (affine-free)
    long long P;                 // B*H*W
    int C, relu;
    double* sums;                // scratch, 2*C doubles
    float* draw;                 // =  gradient of raw
    float* dres;                 // += gradient of the residual input, or null
    float* dgamma;               // =  [C] or null
    float* dbeta;                // =  [C] or null
};
void launch_bn_backward(const BnBwdParams& p, cudaStream_t st);

// out[c] = sum over P rows of x[P][C]  (bias gradients: the head stems and the 1x1 output convolutions);  sums: C doubles of scratch
void launch_colsum(const float* x, long long P, int C, double* sums, float* out, cudaStream_t st);

// MaxPool2d(2, 2) (dla.py:176-177,193): dx += dy at the first maximum of each window in (ky, kx) order
void launch_maxpool2_backward(const float* x, const float* dy, float* dx, int B, int C, int Hin, int Win, cudaStream_t st);

// depthwise ConvTranspose2d k=4 s=2 p=1 (dla_neck.py:58-65): dx += ..., dw[C][4][4] += ...
void launch_upsample2_backward(const float* x, const float* w, const float* dy, float* dx, float* dw, int B, int C, int Hin, int Win,
                               cudaStream_t st);

// Heads: 1x1 output convolutions + output transforms, ReLU, AttnBatchNorm2d (monocon_heads.py:114-131,165-200;
// model/norm/attentive_norm.py:79-91,154-164), from dL/dpred down to the gradient of the pre-norm stems.
struct HeadBwdParams {
    const float* pred[kNumPred];     // forward outputs, NCHW (post sigmoid+clamp / depth transform)
    const float* dpred[kNumPred];    // dL/dpred, NCHW (mc_losses)
    const float* stems;              // [B][HW][576] pre-norm stem outputs
    const double* sums;              // [B][576][2] per-sample (sum, sum of squares) of stems (launch_attn_stats)
    const float *coefA, *coefB;      // [B][576] forward affine of this batch (launch_attn_mix_train)
    const float *att_w, *att_gamma, *att_beta;   // [9][10][64], [9][10], [9][10]
    const float *bank_w, *bank_b;    // [9][10][64]
    const float* w;                  // [65][64]
    int B, HW;
    void* scratch;                   // head_bwd_scratch_bytes(B, HW)
    float* dstems;                   // =  [B][HW][576]
    float *dw, *dbias;               // =  [65][64], [65]
    float *datt_w, *datt_gamma, *datt_beta, *dbank_w, *dbank_b;   // =
};
size_t head_bwd_scratch_bytes(int B, int HW);
void launch_head_backward(const HeadBwdParams& p, cudaStream_t st);

}  // namespace mc
