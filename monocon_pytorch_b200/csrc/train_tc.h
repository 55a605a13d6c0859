// bf16 tensor-core training step (BASELINE.json configs[2] "batch=32 training step bf16", configs[4]):
// launchers on plain device pointers.  Activations, raw convolution outputs and activation gradients are bf16 NHWC (what
// torch.autocast(bfloat16) stores); BatchNorm statistics, parameters, parameter gradients and the optimiser state are fp32.
#pragma once
#include <memory>
#include <string>

#include "engine.h"

namespace mc {

// ---- weight gradient on the tensor cores (wgrad_tc.cu) ----------------------------------------------------------------
struct WgradSrc { const void* x; int C; };       // bf16 NHWC [B][H][W][C], dense
struct WgradDesc {
    const void* dy;                              // bf16 NHWC [B][H][W][Cout]: gradient of the raw convolution output; for a stride-2
                                                 // layer the zero-inserted gradient at INPUT resolution (dy at even rows / columns)
    WgradSrc src[kMaxSrc];                       // the forward inputs, concatenated along C in this order
    int nsrc;
    int H, W, Cout, k;                           // k = 3 (pad 1) or 1 (pad 0), stride 1
    float* dw;                                   // += [k*k][Cin][Cout] fp32 (ConvLayer::w_simt layout)
};
struct WgradPlan;
void wgrad_tc_init();
bool wgrad_tc_supported(const WgradDesc& d);
std::shared_ptr<WgradPlan> wgrad_tc_prepare(const WgradDesc& d, int max_batch, DeviceArena& arena, const std::string& name);
void wgrad_tc_launch(const WgradPlan& plan, int B, cudaStream_t st);

}  // namespace mc
