// TEST INFRASTRUCTURE ONLY.  Lets the GPU-less build container execute the engine's HOST logic for the training step -- plan
// construction, parameter packing, the records of the backward pass (api.cu: setup_backward), mc_backward_train / mc_get_grad /
// mc_get_param / mc_train_tensor -- by linking api.cu + engine.cu (compiled by g++) against
//   * this file: a stand-in CUDA runtime on host memory (cudaMalloc = calloc, copies = memcpy, streams / events / graphs inert) and
//     plain-loop host versions of the FORWARD launchers the train-mode path calls (the CUDA ones are validated on the GPU:
//     tests/test_gpu_train_forward.py); every launcher the training path does not use aborts,
//   * the host-shim build of csrc/train_backward.cu (the real backward kernels, run sequentially).
// tests/test_host_engine.py drives the result through the C ABI.  Nothing under monocon_pytorch_b200/ ever loads it.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "engine.h"
#include "handle.h"

using namespace mc;

// ---------------------------------------------------------------------------------------------
// stand-in CUDA runtime
// ---------------------------------------------------------------------------------------------
extern "C" {
cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDeviceProperties_v2(cudaDeviceProp* p, int) {
    std::memset(p, 0, sizeof(*p));
    p->major = 10; p->minor = 0; p->multiProcessorCount = 148;
    std::strcpy(p->name, "host stand-in");
    return cudaSuccess;
}
const char* cudaGetErrorString(cudaError_t) { return "host stand-in runtime"; }
cudaError_t cudaMalloc(void** p, size_t n) { *p = std::calloc(n ? n : 1, 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
cudaError_t cudaMemset(void* p, int v, size_t n) { std::memset(p, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return cudaSuccess; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = nullptr; return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode) { return cudaErrorNotSupported; }
cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t*) { return cudaErrorNotSupported; }
cudaError_t cudaGraphInstantiate(cudaGraphExec_t*, cudaGraph_t, unsigned long long) { return cudaErrorNotSupported; }
cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t cudaGraphDestroy(cudaGraph_t) { return cudaSuccess; }
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return cudaSuccess; }
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t*, void*) { return cudaErrorNotSupported; }
cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_t, unsigned) { return cudaErrorNotSupported; }
cudaError_t cudaIpcCloseMemHandle(void*) { return cudaErrorNotSupported; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { *p = std::calloc(n ? n : 1, 1); return cudaSuccess; }
cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
}

namespace mc {

static void unused(const char* what) {
    std::fprintf(stderr, "host engine stand-in: %s is not part of the training path\n", what);
    std::abort();
}

// bf16 tensor-core training step (train_engine_tc.cu): GPU only
std::shared_ptr<TrainTc> traintc_create() { unused("traintc_create"); return nullptr; }
void traintc_before_pack(mc_handle*, int, ConvLayer&, const std::vector<float>&) { unused("traintc_before_pack"); }
void traintc_setup(mc_handle*) { unused("traintc_setup"); }
void traintc_forward(mc_handle*, const float*, int, float* const*, cudaStream_t) { unused("traintc_forward"); }
void traintc_backward(mc_handle*, int, int, int, bool, cudaStream_t) { unused("traintc_backward"); }
void traintc_debug(mc_handle*, int, int, const void**, DType*, int*, int*, int*) { unused("traintc_debug"); }

// tensor-core paths: not available here (the fp32 training engine never asks for them)
bool tc_conv_supported(const Net&, const ConvLayer&) { return false; }
void tc_conv_prepare(Net&, ConvLayer&, const std::vector<float>&) { unused("tc_conv_prepare"); }
void tc_conv_launch(const Net&, const ConvLayer&, int, cudaStream_t) { unused("tc_conv_launch"); }
void tc_kernels_init() {}
bool tc2_conv_supported(const Net&, const ConvLayer&) { return false; }
void tc2_conv_prepare(Net&, ConvLayer&, const std::vector<float>&) { unused("tc2_conv_prepare"); }
void tc2_conv_launch(const Net&, const ConvLayer&, int, cudaStream_t) { unused("tc2_conv_launch"); }
void tc2_kernels_init() {}
bool tc3_conv_supported(const Net&, const ConvLayer&) { return false; }
void tc3_conv_prepare(Net&, ConvLayer&, const std::vector<float>&) { unused("tc3_conv_prepare"); }
void tc3_conv_launch(const Net&, const ConvLayer&, int, cudaStream_t) { unused("tc3_conv_launch"); }
void tc3_kernels_init() {}
bool dcn_tc_supported(const Net&, const ConvLayer&) { return false; }
void dcn_tc_prepare(Net&, ConvLayer&, const std::vector<float>&) { unused("dcn_tc_prepare"); }
void dcn_tc_launch(const Net&, const ConvLayer&, int, cudaStream_t) { unused("dcn_tc_launch"); }
void dcn_tc_init() {}
bool head_tc_supported(DType, int) { return false; }
std::shared_ptr<HeadTcPlan> head_tc_prepare(Net&, const void*, DType, int, int, const std::vector<float>&, int) { unused("head_tc_prepare"); return nullptr; }
void launch_head_apply_tc(const HeadTcPlan&, const HeadApplyParams&, cudaStream_t) { unused("launch_head_apply_tc"); }
void head_tc_init() {}
void head_kernels_init() {}

void launch_pack_input_u8(const unsigned char*, const int*, const float*, void*, DType, int, int, int, int, int, int, int, int, cudaStream_t, const SplitInfo&) { unused("launch_pack_input_u8"); }
void launch_unpack_nchw(const void* srcv, DType dt, float* dst, int B, int C, int H, int W, cudaStream_t, const SplitInfo&) {
    if (dt != DT_F32) unused("bf16 unpack");
    const float* src = (const float*)srcv;
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c)
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) dst[(((size_t)b * C + c) * H + y) * W + x] = src[(((size_t)b * H + y) * W + x) * C + c];
}
void launch_pack_nhwc(const float*, void*, DType, int, int, int, int, cudaStream_t, const SplitInfo&) { unused("launch_pack_nhwc"); }
// eval-mode pieces: inert stand-ins (outputs left as they are), enough to walk the inference entry points' HOST logic --
// staging buffers, slots, stage tables -- under the sanitizers (tests/host_shim/asan.sh); their CUDA versions are validated on the GPU
void launch_attn_mix(const AttnMixParams&, int, cudaStream_t) {}
void launch_decode(const DecodeParams&, unsigned long long*, int*, cudaStream_t) {}
void launch_kitti_boxes(const float*, const unsigned char*, const float*, const int*, int, int, double*, float*, unsigned char*, cudaStream_t) { unused("launch_kitti_boxes"); }
void launch_gather_release(const GatherParams&, unsigned* const*, unsigned, cudaStream_t) { unused("launch_gather_release"); }
void launch_gather_wait(const unsigned*, int, unsigned, int*, cudaStream_t) { unused("launch_gather_wait"); }

// ---------------------------------------------------------------------------------------------
// forward launchers of the train-mode path, as plain loops on host memory (fp32 engine only)
// ---------------------------------------------------------------------------------------------
void launch_pack_input(const float* img, void* dstv, DType dt, int B, int C, int H, int W, int Cpad, int Wp, int xoff, cudaStream_t, const SplitInfo&) {
    if (dt != DT_F32) unused("bf16 input");
    float* dst = (float*)dstv;
    std::memset(dst, 0, sizeof(float) * (size_t)B * H * Wp * Cpad);
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c)
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) dst[(((size_t)b * H + y) * Wp + x + xoff) * Cpad + c] = img[(((size_t)b * C + c) * H + y) * W + x];
}

void launch_conv_simt(const ConvParams& p, DType dt, cudaStream_t) {
    if (dt != DT_F32) unused("bf16 convolution");
    float* dst = (float*)p.dst;
    const float* res = (const float*)p.residual;
    std::vector<float> acc(p.Cout);
    for (int n = 0; n < p.B; ++n)
        for (int oy = 0; oy < p.Hout; ++oy)
            for (int ox = 0; ox < p.Wout; ++ox) {
                std::fill(acc.begin(), acc.end(), 0.f);
                for (int ky = 0; ky < p.k; ++ky) {
                    const int iy = oy * p.stride - p.pad + ky;
                    if (iy < 0 || iy >= p.Hin) continue;
                    for (int kx = 0; kx < p.k; ++kx) {
                        const int ix = ox * p.stride - p.pad + kx;
                        if (ix < 0 || ix >= p.Win) continue;
                        int cbase = 0;
                        for (int s = 0; s < p.nsrc; ++s) {
                            const float* x = (const float*)p.src[s] + (((size_t)n * p.Hin + iy) * p.srcWp[s] + ix + p.srcXoff[s]) * p.srcC[s];
                            for (int c = 0; c < p.srcC[s]; ++c) {
                                const float xv = x[c];
                                const float* w = p.w + ((size_t)(ky * p.k + kx) * p.Cin + cbase + c) * p.Cout;
                                for (int co = 0; co < p.Cout; ++co) acc[co] += xv * w[co];
                            }
                            cbase += p.srcC[s];
                        }
                    }
                }
                const size_t o = (((size_t)n * p.Hout + oy) * p.Wout + ox) * p.Cout;
                for (int co = 0; co < p.Cout; ++co) {
                    float v = acc[co] * p.scale[co] + p.shift[co];
                    if (res) v += res[o + co];
                    if (p.relu) v = v > 0.f ? v : 0.f;
                    dst[o + co] = v;
                }
            }
}

void launch_bn_train_ex(const float* raw, float* y, const float* residual, long long P, int C, double* sums, float eps, float momentum,
                        const float* gamma, const float* beta, float* rmean, float* rvar, float* scale, float* shift, bool relu, float* mean_out,
                        float* inv_out, cudaStream_t) {
    for (int c = 0; c < C; ++c) {
        double s = 0.0, ss = 0.0;
        for (long long i = 0; i < P; ++i) { const double v = raw[i * C + c]; s += v; ss += v * v; }
        sums[2 * c] = s; sums[2 * c + 1] = ss;
        const double n = (double)P, mean = s / n;
        double var = ss / n - mean * mean;
        if (var < 0.0) var = 0.0;
        const float inv = (float)(1.0 / std::sqrt(var + (double)eps));
        const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
        scale[c] = g * inv;
        shift[c] = b - (float)mean * g * inv;
        if (mean_out) { mean_out[c] = (float)mean; inv_out[c] = inv; }
        const double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
        rmean[c] = (1.f - momentum) * rmean[c] + momentum * (float)mean;
        rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)unbiased;
    }
    for (long long i = 0; i < P * C; ++i) {
        const int c = (int)(i % C);
        float v = std::fmaf(raw[i], scale[c], shift[c]);
        if (residual) v += residual[i];
        if (relu) v = v > 0.f ? v : 0.f;
        y[i] = v;
    }
}

void launch_bn_train(float* x, const float* residual, long long P, int C, double* sums, float eps, float momentum, const float* gamma,
                     const float* beta, float* rmean, float* rvar, float* scale, float* shift, bool relu, cudaStream_t st) {
    launch_bn_train_ex(x, x, residual, P, C, sums, eps, momentum, gamma, beta, rmean, rvar, scale, shift, relu, nullptr, nullptr, st);
}

void launch_maxpool2(const void* srcv, void* dstv, DType dt, int B, int C, int Hin, int Win, cudaStream_t, const SplitInfo&, const SplitInfo&) {
    if (dt != DT_F32) unused("bf16 pool");
    const float* src = (const float*)srcv;
    float* dst = (float*)dstv;
    const int Ho = Hin / 2, Wo = Win / 2;
    for (int n = 0; n < B; ++n)
        for (int oy = 0; oy < Ho; ++oy)
            for (int ox = 0; ox < Wo; ++ox)
                for (int c = 0; c < C; ++c) {
                    float m = -INFINITY;
                    for (int ky = 0; ky < 2; ++ky)
                        for (int kx = 0; kx < 2; ++kx) m = std::fmax(m, src[(((size_t)n * Hin + 2 * oy + ky) * Win + 2 * ox + kx) * C + c]);
                    dst[(((size_t)n * Ho + oy) * Wo + ox) * C + c] = m;
                }
}

void launch_upsample2(const void* srcv, void* dstv, DType dt, const float* w, int B, int C, int Hin, int Win, cudaStream_t, const SplitInfo&, const SplitInfo&) {
    if (dt != DT_F32) unused("bf16 upsample");
    const float* src = (const float*)srcv;
    float* dst = (float*)dstv;
    const int Ho = 2 * Hin, Wo = 2 * Win;
    std::memset(dst, 0, sizeof(float) * (size_t)B * Ho * Wo * C);
    for (int n = 0; n < B; ++n)
        for (int i = 0; i < Hin; ++i)
            for (int j = 0; j < Win; ++j)
                for (int ky = 0; ky < 4; ++ky) {
                    const int oy = 2 * i - 1 + ky;
                    if (oy < 0 || oy >= Ho) continue;
                    for (int kx = 0; kx < 4; ++kx) {
                        const int ox = 2 * j - 1 + kx;
                        if (ox < 0 || ox >= Wo) continue;
                        for (int c = 0; c < C; ++c)
                            dst[(((size_t)n * Ho + oy) * Wo + ox) * C + c] += src[(((size_t)n * Hin + i) * Win + j) * C + c] * w[c * 16 + ky * 4 + kx];
                    }
                }
}

// deformable columns (csrc/dcn.cu), fp32 storage: plain loops over pixel / tap / channel with the operator's sampling rule
void launch_dcn_columns(const DcnColParams& p, DType dt, cudaStream_t) {
    if (dt != DT_F32) unused("bf16 / fp16-plane deformable columns");
    const float* off = (const float*)p.off;
    float* col = (float*)p.col;
    const int H = p.H, W = p.W;
    for (int n = 0; n < p.B; ++n)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                const size_t pix = ((size_t)n * H + y) * W + x;
                for (int k = 0; k < 9; ++k) {
                    const float dy = off[pix * p.offC + 2 * k], dx = off[pix * p.offC + 2 * k + 1], mv = off[pix * p.offC + 18 + k];
                    const float m = p.mask_logits ? 1.f / (1.f + std::exp(-mv)) : mv;
                    const float py = (float)(y - 1 + k / 3) + dy, px = (float)(x - 1 + k % 3) + dx;
                    const bool inside = py > -1.f && py < (float)H && px > -1.f && px < (float)W;
                    const int h0 = (int)std::floor(py), w0 = (int)std::floor(px);
                    const float lh = py - h0, lw = px - w0;
                    int cbase = 0;
                    for (int s = 0; s < p.nsrc; ++s) {
                        const float* src = (const float*)p.src[s];
                        const int C = p.srcC[s];
                        for (int c = 0; c < C; ++c) {
                            auto at = [&](int yy, int xx) -> float {
                                return (yy >= 0 && yy <= H - 1 && xx >= 0 && xx <= W - 1) ? src[(((size_t)n * H + yy) * W + xx) * C + c] : 0.f;
                            };
                            float v = 0.f;
                            if (inside)
                                v = (1.f - lh) * (1.f - lw) * at(h0, w0) + (1.f - lh) * lw * at(h0, w0 + 1) + lh * (1.f - lw) * at(h0 + 1, w0) + lh * lw * at(h0 + 1, w0 + 1);
                            col[pix * (size_t)(9 * p.Cin) + (size_t)k * p.Cin + cbase + c] = m * v;
                        }
                        cbase += C;
                    }
                }
            }
}

void launch_attn_stats(const void* stemsv, DType dt, double* sums, int B, int HW, cudaStream_t) {
    if (dt != DT_F32) unused("bf16 stems");
    const float* x = (const float*)stemsv;
    for (int b = 0; b < B; ++b)
        for (int ch = 0; ch < kStemTot; ++ch) {
            double s = 0.0, ss = 0.0;
            for (int p = 0; p < HW; ++p) { const double v = x[((size_t)b * HW + p) * kStemTot + ch]; s += v; ss += v * v; }
            sums[((size_t)b * kStemTot + ch) * 2] = s;
            sums[((size_t)b * kStemTot + ch) * 2 + 1] = ss;
        }
}

// sequential restatement of attn_mix_train_kernel (csrc/train_forward.cu)
void launch_attn_mix_train(const double* sums, int B, int HW, const float* att_w, const float* att_gamma, const float* att_beta, float* att_rmean,
                           float* att_rvar, const float* bank_w, const float* bank_b, float* bn_rmean, float* bn_rvar, float* coefA, float* coefB,
                           cudaStream_t) {
    const double n = (double)HW;
    std::vector<float> y((size_t)B * kStemC), a((size_t)B * kNumAff), inv(kStemC), bmeanf(kStemC);
    for (int s = 0; s < kNumStems; ++s) {
        for (int c = 0; c < kStemC; ++c) {
            const int ch = s * kStemC + c;
            double bs = 0.0, bss = 0.0;
            for (int b = 0; b < B; ++b) {
                const double sum = sums[((size_t)b * kStemTot + ch) * 2], sq = sums[((size_t)b * kStemTot + ch) * 2 + 1];
                bs += sum; bss += sq;
                const double mean = sum / n;
                double var = (sq - sum * mean) / (n - 1.0);
                if (var < 0.0) var = 0.0;
                y[(size_t)b * kStemC + c] = (float)mean * (1.0f / std::sqrt((float)var + 1e-3f));
            }
            const double N = n * B, bmean = bs / N;
            double bvar = bss / N - bmean * bmean;
            if (bvar < 0.0) bvar = 0.0;
            inv[c] = (float)(1.0 / std::sqrt(bvar + 1e-3));
            bmeanf[c] = (float)bmean;
            bn_rmean[ch] = (1.f - 0.03f) * bn_rmean[ch] + 0.03f * (float)bmean;
            bn_rvar[ch] = (1.f - 0.03f) * bn_rvar[ch] + 0.03f * (float)(bvar * N / (N - 1.0));
        }
        for (int j = 0; j < kNumAff; ++j) {
            double m = 0.0, q = 0.0;
            for (int b = 0; b < B; ++b) {
                float acc = 0.f;
                for (int k = 0; k < kStemC; ++k) acc = std::fmaf(att_w[((size_t)s * kNumAff + j) * kStemC + k], y[(size_t)b * kStemC + k], acc);
                a[(size_t)b * kNumAff + j] = acc;
                m += acc; q += (double)acc * acc;
            }
            m /= B;
            double v = q / B - m * m;
            if (v < 0.0) v = 0.0;
            const float ainv = (float)(1.0 / std::sqrt(v + 1e-5));
            const float g = att_gamma[s * kNumAff + j], be = att_beta[s * kNumAff + j];
            for (int b = 0; b < B; ++b) {
                const float t = (a[(size_t)b * kNumAff + j] - (float)m) * ainv * g + be;
                a[(size_t)b * kNumAff + j] = std::fmin(std::fmax(t + 3.f, 0.f), 6.f) / 6.f;
            }
            att_rmean[s * kNumAff + j] = 0.9f * att_rmean[s * kNumAff + j] + 0.1f * (float)m;
            att_rvar[s * kNumAff + j] = 0.9f * att_rvar[s * kNumAff + j] + 0.1f * (float)(B > 1 ? v * B / (B - 1.0) : v);
        }
        for (int b = 0; b < B; ++b)
            for (int c = 0; c < kStemC; ++c) {
                float gamma = 0.f, beta = 0.f;
                for (int j = 0; j < kNumAff; ++j) {
                    gamma = std::fmaf(a[(size_t)b * kNumAff + j], bank_w[((size_t)s * kNumAff + j) * kStemC + c], gamma);
                    beta = std::fmaf(a[(size_t)b * kNumAff + j], bank_b[((size_t)s * kNumAff + j) * kStemC + c], beta);
                }
                const float A = gamma * inv[c];
                coefA[(size_t)b * kStemTot + s * kStemC + c] = A;
                coefB[(size_t)b * kStemTot + s * kStemC + c] = beta - A * bmeanf[c];
            }
    }
}

void launch_head_apply(const HeadApplyParams& p, DType dt, cudaStream_t) {
    if (dt != DT_F32) unused("bf16 head");
    static const int o0[kNumStems] = {0, 12, 14, 18, 3, 16, 36, 39, 41}, o1[kNumStems] = {3, 14, 16, 36, 12, 18, 39, 41, 65};
    static const int pch[kNumPred] = {3, 9, 2, 2, 2, 18, 3, 2, 12, 12}, po0[kNumPred] = {0, 3, 12, 14, 16, 18, 36, 39, 41, 53};
    const float* x = (const float*)p.stems;
    float z[kStemC];
    for (int b = 0; b < p.B; ++b)
        for (int pix = 0; pix < p.HW; ++pix)
            for (int s = 0; s < kNumStems; ++s) {
                for (int c = 0; c < kStemC; ++c) {
                    const size_t ch = (size_t)s * kStemC + c;
                    const float v = std::fmaf(p.coefA[(size_t)b * kStemTot + ch], x[((size_t)b * p.HW + pix) * kStemTot + ch], p.coefB[(size_t)b * kStemTot + ch]);
                    z[c] = v > 0.f ? v : 0.f;
                }
                for (int o = o0[s]; o < o1[s]; ++o) {
                    float acc = p.bias[o];
                    for (int c = 0; c < kStemC; ++c) acc = std::fmaf(p.w[o * kStemC + c], z[c], acc);
                    int k = 0;
                    while (k + 1 < kNumPred && po0[k + 1] <= o) ++k;
                    if (o < 12) {
                        acc = 1.f / (1.f + std::exp(-acc));
                        acc = std::fmin(std::fmax(acc, 1e-4f), 1.f - 1e-4f);
                    } else if (o == 39) {
                        acc = 1.f / (1.f / (1.f + std::exp(-acc)) + 1e-12f) - 1.f;
                    }
                    p.out[k][((size_t)b * pch[k] + (o - po0[k])) * p.HW + pix] = acc;
                }
            }
}

}  // namespace mc
