// bf16 tensor-core training step (BASELINE.json configs[2] "batch=32 training step bf16", configs[4]):
// launchers on plain device pointers.  Activations, raw convolution outputs and activation gradients are bf16 NHWC (what
// torch.autocast(bfloat16) stores); BatchNorm statistics, parameters, parameter gradients and the optimiser state are fp32.
#pragma once
#include <memory>
#include <string>

#include "engine.h"

namespace mc {

// ---- weight gradient on the tensor cores (wgrad_tc.cu) ----------------------------------------------------------------
struct WgradSrc { const void* x; int C; int Wp = 0, xoff = 0; };   // bf16 NHWC [B][H][W][C], dense; the stem's padded image: row pitch Wp pixels, x = 0 at column xoff
struct WgradDesc {
    const void* dy;                              // bf16 NHWC [B][H][W][Cout]: gradient of the raw convolution output; for a stride-2
                                                 // layer the zero-inserted gradient at INPUT resolution (dy at even rows / columns)
    WgradSrc src[kMaxSrc];                       // the forward inputs, concatenated along C in this order
    int nsrc;
    int H, W, Cout, k;                           // k = 3 (pad 1) or 1 (pad 0), stride 1; k = 7 (pad 3): the stem over the 8-channel padded image
    float* dw;                                   // += [k*k][Cin][Cout] fp32 (ConvLayer::w_simt layout)
};
struct WgradPlan;
void wgrad_tc_init();
bool wgrad_tc_supported(const WgradDesc& d);
std::shared_ptr<WgradPlan> wgrad_tc_prepare(const WgradDesc& d, int max_batch, DeviceArena& arena, const std::string& name);
void wgrad_tc_launch(const WgradPlan& plan, int B, cudaStream_t st);


// ---- bandwidth kernels of the bf16 training step (train_tc.cu); all activation pointers are bf16 NHWC ---------------------
void launch_bn_stats_bf16(const void* x, long long P, int C, double* sums /*[C][2]*/, cudaStream_t st, bool zeroed = false /* caller zeroed sums */);
// bn_finalize + bn_apply in one launch (every thread derives its channels' constants from the sums; block 0 publishes them)
void launch_bn_finalize_apply_bf16(const void* raw, void* y, const void* residual, long long P, int C, const double* sums, float eps, float momentum,
                                   const float* gamma, const float* beta, float* rmean, float* rvar, float* scale, float* shift, float* mean_out,
                                   float* inv_out, bool relu, cudaStream_t st);
void launch_bn_apply_bf16(const void* raw, void* y, const void* residual, long long P, int C, const float* scale, const float* shift, bool relu,
                          cudaStream_t st);
// mean, biased variance -> scale / shift, running statistics, batch mean / inverse std (train_forward.cu: bn_finalize_kernel)
void launch_bn_finalize(const double* sums, int C, long long P, float eps, float momentum, const float* gamma, const float* beta, float* rmean,
                        float* rvar, float* scale, float* shift, float* mean_out, float* inv_out, cudaStream_t st);
struct BnBwdTcParams {
    const void *dy, *y, *raw;     // gradient of y, y (its sign is the ReLU mask; may be null when relu == 0), raw convolution output
    const float *mean, *inv, *gamma;
    const float *fscale = nullptr, *fshift = nullptr;   // the forward's y = raw * scale + shift: with no residual the ReLU mask is taken from raw
    double* sums;                 // scratch, 2 * C doubles
    bool sums_zeroed = false;     // the caller zeroed `sums` (one memset for all layers of a pass)
    long long P;                  // B * H * W
    int C, relu;
    int up, H, W;                 // up = 1: draw is zero-inserted, [B][2H][2W][C] with this layer's pixels at even rows / columns
    void* draw;                   // =  gradient of the raw convolution output
    void* dres;                   // gradient of the residual input, or null
    int dres_acc;                 // 1: +=, 0: = (the residual's first contribution in backward order)
    float *dgamma, *dbeta;        // =
};
void launch_bn_backward_bf16(const BnBwdTcParams& p, cudaStream_t st);
void launch_maxpool2_backward_bf16(const void* x, const void* dy, void* dx, int B, int C, int Hin, int Win, bool accumulate, cudaStream_t st);
void launch_upsample2_backward_bf16(const void* x, const float* w, const void* dy, void* dx, float* dw /* += [C][16] */, int B, int C, int Hin, int Win,
                                    bool accumulate, cudaStream_t st);
// heads backward restructured for the device (train_tc_head.cu): same outputs as launch_head_backward (train_backward.h) with
// p.dstems optional (null: skipped), the pre-norm stems given separately as fp32 or bf16 (p.stems is not read), the gradient of the pre-norm stems as bf16 [B][HW][576] and the gradient of the 576 stem biases
struct HeadBwdParams;
void launch_head_backward_tc(const HeadBwdParams& p, const void* stems, bool stems_bf16, void* dstems_bf16, float* dstem_bias, cudaStream_t st);
void launch_bf16_to_f32(const void* in, float* out, long long n, cudaStream_t st);
void launch_f32_to_bf16(const float* in, void* out, long long n, cudaStream_t st);
// out[i] = bf16(idx[i] >= 0 ? master[idx[i]] : 0): the fp32 master weights into a convolution plan's bf16 layout
struct RepackJob { const float* master; const int* idx; void* out; long long start; long long n; };   // start: first BLOCK of this job; n: its elements
long long repack_blocks(long long n);                 // blocks a job of n elements takes
void launch_repack_all_bf16(const RepackJob* jobs_dev, int njobs, long long total_blocks, cudaStream_t st);
void launch_repack_bf16(const float* master, const int* idx, void* out, long long n, cudaStream_t st);

}  // namespace mc
