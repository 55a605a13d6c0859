// Train-mode forward of the network (first half of SURVEY.md 8(f) row 1 / BASELINE.json configs[2]): every BatchNorm uses
// the statistics of the current batch and updates its running statistics (nn.BatchNorm2d in train(), momentum 0.1;
// AttnBatchNorm2d: base BN momentum 0.03, eps 1e-3, and the 10-channel BatchNorm of the attention branch over the batch,
// model/norm/attentive_norm.py:49-53,79-91,154-164).  fp32 engine (MC_PREC_FP32) only for now: the convolutions are the FFMA
// kernels writing the raw convolution output, followed by
//   bn_stats_kernel      per-channel sum / sum of squares over B*H*W (NHWC), fp64 accumulation
//   bn_finalize_kernel   mean, biased variance -> scale / shift; running_mean / running_var (unbiased) update
//   bn_apply_kernel      y = x * scale + shift (+ residual) (ReLU), in place
// and for the heads attn_stats_kernel (per-sample sums, also the batch sums) -> attn_mix_train_kernel -> head_apply_kernel.
// Raw convolution outputs are normalised in place unless the engine was finalized for the backward pass (mc_finalize_params(h, 2):
// launch_bn_train_ex keeps them and the batch mean / inverse std for csrc/train_backward.cu).
#include <algorithm>
#include <cstring>

#include "engine.h"
#include "train_tc.h"

namespace mc {

namespace {

constexpr int kBnThreads = 256;

// x: [P][C] fp32 (NHWC flattened).  sums: [C][2] doubles (zeroed by the caller).
__global__ void __launch_bounds__(kBnThreads) bn_stats_kernel(const float* __restrict__ x, long long P, int C, double* __restrict__ sums) {
    extern __shared__ double sh[];                   // [C][2]
    for (int i = threadIdx.x; i < 2 * C; i += kBnThreads) sh[i] = 0.0;
    __syncthreads();
    const int cc = C < kBnThreads ? C : kBnThreads;  // threads along the channel axis
    const int ppb = kBnThreads / cc;                 // pixels per block iteration
    const int c0 = threadIdx.x % cc, pr = threadIdx.x / cc;
    if (pr < ppb) {
        for (int c = c0; c < C; c += cc) {
            double s = 0.0, ss = 0.0;
            float fs = 0.f, fss = 0.f;
            int n = 0;
            for (long long pix = (long long)blockIdx.x * ppb + pr; pix < P; pix += (long long)gridDim.x * ppb) {
                const float v = x[pix * C + c];
                fs += v; fss = fmaf(v, v, fss);
                if (++n == 64) { s += (double)fs; ss += (double)fss; fs = 0.f; fss = 0.f; n = 0; }
            }
            s += (double)fs; ss += (double)fss;
            atomicAdd(&sh[2 * c], s);
            atomicAdd(&sh[2 * c + 1], ss);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += kBnThreads)
        if (sh[i] != 0.0) atomicAdd(&sums[i], sh[i]);
}

// one thread per channel
__global__ void bn_finalize_kernel(const double* __restrict__ sums, int C, double n, float eps, float momentum, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ rmean, float* __restrict__ rvar,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_out,
                                   float* __restrict__ inv_out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double mean = sums[2 * c] / n;
    double var = sums[2 * c + 1] / n - mean * mean;             // biased: what the normalisation uses
    if (var < 0.0) var = 0.0;
    const float inv = (float)(1.0 / sqrt(var + (double)eps));
    const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
    scale[c] = g * inv;
    shift[c] = b - (float)mean * g * inv;
    if (mean_out) { mean_out[c] = (float)mean; inv_out[c] = inv; }     // kept for the backward pass
    const double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
    rmean[c] = (1.f - momentum) * rmean[c] + momentum * (float)mean;
    rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)unbiased;
}

// in == out for the in-place forward; the backward-enabled forward keeps the raw convolution output and writes elsewhere
__global__ void __launch_bounds__(256) bn_apply_kernel(const float* in, float* x, const float* __restrict__ res, long long total4, int C,
                                                      const float* __restrict__ scale, const float* __restrict__ shift, int relu) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)((i * 4) % C);
        float4 v = reinterpret_cast<const float4*>(in)[i];
        const float4 s = *reinterpret_cast<const float4*>(scale + c), h = *reinterpret_cast<const float4*>(shift + c);
        v.x = fmaf(v.x, s.x, h.x); v.y = fmaf(v.y, s.y, h.y); v.z = fmaf(v.z, s.z, h.z); v.w = fmaf(v.w, s.w, h.w);
        if (res) {
            const float4 r = reinterpret_cast<const float4*>(res)[i];
            v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
        }
        if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        reinterpret_cast<float4*>(x)[i] = v;
    }
}

struct AttnMixTrainParams {
    const double* sums;       // [B][576][2] per-sample sums of the pre-norm stems
    int B, HW;
    const float *att_w, *att_gamma, *att_beta;       // [9][10][64], [9][10], [9][10]
    float *att_rmean, *att_rvar;                     // [9][10] running statistics of the attention BatchNorm2d(10)
    const float *bank_w, *bank_b;                    // [9][10][64]
    float *bn_rmean, *bn_rvar;                       // [576] running statistics of the affine-free base BN
    float *coefA, *coefB;                            // [B][576]
};

constexpr int kMaxTrainB = 64;

// one CTA per stem, 64 threads (one per channel); B <= 64
__global__ void __launch_bounds__(64) attn_mix_train_kernel(const AttnMixTrainParams p) {
    __shared__ float y[kMaxTrainB][kStemC];
    __shared__ float a[kMaxTrainB][kNumAff];
    const int s = blockIdx.x, c = threadIdx.x, ch = s * kStemC + c, B = p.B;
    const double n = (double)p.HW;
    // base BN (affine-free, eps 1e-3): batch statistics over B * HW, running update with momentum 0.03
    double bs = 0.0, bss = 0.0;
    for (int b = 0; b < B; ++b) {
        const double sum = p.sums[((long long)b * kStemTot + ch) * 2], sq = p.sums[((long long)b * kStemTot + ch) * 2 + 1];
        bs += sum; bss += sq;
        const double mean = sum / n;
        double var = (sq - sum * mean) / (n - 1.0);                 // unbiased instance variance (attentive_norm.py:84)
        if (var < 0.0) var = 0.0;
        y[b][c] = (float)mean * rsqrtf((float)var + 1e-3f);
    }
    const double N = n * B, bmean = bs / N;
    double bvar = bss / N - bmean * bmean;
    if (bvar < 0.0) bvar = 0.0;
    const float inv = (float)(1.0 / sqrt(bvar + 1e-3));
    p.bn_rmean[ch] = (1.f - 0.03f) * p.bn_rmean[ch] + 0.03f * (float)bmean;
    p.bn_rvar[ch] = (1.f - 0.03f) * p.bn_rvar[ch] + 0.03f * (float)(bvar * N / (N - 1.0));
    __syncthreads();
    // attention branch: conv1x1 (64 -> 10), BatchNorm2d(10) over the batch (N = B), hsigmoid
    for (int i = c; i < B * kNumAff; i += 64) {
        const int b = i / kNumAff, j = i % kNumAff;
        const float* w = p.att_w + ((long long)s * kNumAff + j) * kStemC;
        float acc = 0.f;
        for (int k = 0; k < kStemC; ++k) acc = fmaf(w[k], y[b][k], acc);
        a[b][j] = acc;
    }
    __syncthreads();
    if (c < kNumAff) {
        double m = 0.0, q = 0.0;
        for (int b = 0; b < B; ++b) { m += a[b][c]; q += (double)a[b][c] * a[b][c]; }
        m /= B;
        double v = q / B - m * m;
        if (v < 0.0) v = 0.0;
        const float ainv = (float)(1.0 / sqrt(v + 1e-5));
        const float g = p.att_gamma[s * kNumAff + c], be = p.att_beta[s * kNumAff + c];
        for (int b = 0; b < B; ++b) {
            const float t = (a[b][c] - (float)m) * ainv * g + be;
            a[b][c] = fminf(fmaxf(t + 3.f, 0.f), 6.f) / 6.f;
        }
        p.att_rmean[s * kNumAff + c] = 0.9f * p.att_rmean[s * kNumAff + c] + 0.1f * (float)m;
        p.att_rvar[s * kNumAff + c] = 0.9f * p.att_rvar[s * kNumAff + c] + 0.1f * (float)(B > 1 ? v * B / (B - 1.0) : v);
    }
    __syncthreads();
    for (int b = 0; b < B; ++b) {
        float gamma = 0.f, beta = 0.f;
        for (int j = 0; j < kNumAff; ++j) {
            gamma = fmaf(a[b][j], p.bank_w[((long long)s * kNumAff + j) * kStemC + c], gamma);
            beta = fmaf(a[b][j], p.bank_b[((long long)s * kNumAff + j) * kStemC + c], beta);
        }
        const float A = gamma * inv;
        p.coefA[(long long)b * kStemTot + ch] = A;
        p.coefB[(long long)b * kStemTot + ch] = beta - A * (float)bmean;
    }
}

}  // namespace

void launch_bn_train_ex(const float* raw, float* y, const float* residual, long long P, int C, double* sums, float eps, float momentum,
                        const float* gamma, const float* beta, float* rmean, float* rvar, float* scale, float* shift, bool relu, float* mean_out,
                        float* inv_out, cudaStream_t st) {
    MC_CHECK(C % 4 == 0 && C <= 1024, "bn_train: C must be a multiple of 4 and <= 1024");
    MC_CHECK((mean_out == nullptr) == (inv_out == nullptr), "bn_train: mean_out and inv_out come together");
    MC_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, st));
    const int cc = std::min(C, kBnThreads), ppb = kBnThreads / cc;
    const int grid = (int)std::min<long long>((P + ppb - 1) / ppb, 148 * 4);
    bn_stats_kernel<<<grid, kBnThreads, sizeof(double) * 2 * C, st>>>(raw, P, C, sums);
    MC_CUDA(cudaGetLastError());
    bn_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums, C, (double)P, eps, momentum, gamma, beta, rmean, rvar, scale, shift, mean_out, inv_out);
    MC_CUDA(cudaGetLastError());
    const long long total4 = P * C / 4;
    bn_apply_kernel<<<(int)std::min<long long>((total4 + 255) / 256, 148 * 8), 256, 0, st>>>(raw, y, residual, total4, C, scale, shift, relu ? 1 : 0);
    MC_CUDA(cudaGetLastError());
}

void launch_bn_finalize(const double* sums, int C, long long P, float eps, float momentum, const float* gamma, const float* beta, float* rmean,
                        float* rvar, float* scale, float* shift, float* mean_out, float* inv_out, cudaStream_t st) {
    MC_CHECK((mean_out == nullptr) == (inv_out == nullptr), "bn_finalize: mean_out and inv_out come together");
    bn_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums, C, (double)P, eps, momentum, gamma, beta, rmean, rvar, scale, shift, mean_out, inv_out);
    MC_CUDA(cudaGetLastError());
}

void launch_bn_train(float* x, const float* residual, long long P, int C, double* sums, float eps, float momentum, const float* gamma,
                     const float* beta, float* rmean, float* rvar, float* scale, float* shift, bool relu, cudaStream_t st) {
    launch_bn_train_ex(x, x, residual, P, C, sums, eps, momentum, gamma, beta, rmean, rvar, scale, shift, relu, nullptr, nullptr, st);
}

void launch_attn_mix_train(const double* sums, int B, int HW, const float* att_w, const float* att_gamma, const float* att_beta,
                           float* att_rmean, float* att_rvar, const float* bank_w, const float* bank_b, float* bn_rmean, float* bn_rvar,
                           float* coefA, float* coefB, cudaStream_t st) {
    MC_CHECK(B >= 2 && B <= kMaxTrainB, "train-mode AttnBN needs 2 <= B <= 64 (the 10-channel BatchNorm sees a (B,10,1,1) tensor)");
    AttnMixTrainParams p;
    p.sums = sums; p.B = B; p.HW = HW;
    p.att_w = att_w; p.att_gamma = att_gamma; p.att_beta = att_beta; p.att_rmean = att_rmean; p.att_rvar = att_rvar;
    p.bank_w = bank_w; p.bank_b = bank_b; p.bn_rmean = bn_rmean; p.bn_rvar = bn_rvar; p.coefA = coefA; p.coefB = coefB;
    attn_mix_train_kernel<<<kNumStems, 64, 0, st>>>(p);
    MC_CUDA(cudaGetLastError());
}

}  // namespace mc
