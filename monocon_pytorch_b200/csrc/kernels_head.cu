// MonoCon dense-head kernels: Attentive-Normalisation statistics / mixture, and the fused
// "normalise + ReLU + ten 1x1 convolutions + output activations" kernel.  sm_100a.
//
// Reference semantics: model/norm/attentive_norm.py:79-91,154-164 and
// model/dense_heads/monocon_heads.py:114-131,165-200.
#include "common.cuh"

namespace mc {

template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<bf16>(const bf16* p) { return __bfloat162float(*p); }

// ---------------------------------------------------------------------------------------------
// instance statistics: sums[b][c] = (sum x, sum x^2) over HW, c in [0,576).
// grid (chunks, B); each CTA reduces a slab of pixels for all 576 channels (coalesced along C),
// accumulates per-thread in fp32 over <= 64 pixels, then in fp64 through shared + global atomics.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(576) attn_stats_kernel(const T* __restrict__ x, double* __restrict__ sums, int HW,
                                                         int pix_per_cta) {
    const int b = blockIdx.y;
    const int c = threadIdx.x;                       // 576 threads: one channel each
    const int p0 = blockIdx.x * pix_per_cta;
    const int p1 = min(HW, p0 + pix_per_cta);
    const T* base = x + ((long long)b * HW) * kStemTot + c;
    double s = 0.0, ss = 0.0;
    for (int q = p0; q < p1; q += 32) {
        float fs = 0.f, fss = 0.f;
        const int qe = min(p1, q + 32);
        for (int pidx = q; pidx < qe; ++pidx) {
            float v = ldf<T>(base + (long long)pidx * kStemTot);
            fs += v;
            fss = fmaf(v, v, fss);
        }
        s += (double)fs;
        ss += (double)fss;
    }
    atomicAdd(&sums[((long long)b * kStemTot + c) * 2 + 0], s);
    atomicAdd(&sums[((long long)b * kStemTot + c) * 2 + 1], ss);
}

void launch_attn_stats(const void* stems, DType dt, double* sums, int B, int HW, cudaStream_t st) {
    MC_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * kStemTot * B, st));
    int chunks = (148 * 4 + B - 1) / B;
    int pix_per_cta = (HW + chunks - 1) / chunks;
    if (pix_per_cta < 32) pix_per_cta = 32;
    chunks = (HW + pix_per_cta - 1) / pix_per_cta;
    dim3 grid(chunks, B);
    if (dt == DT_F32) attn_stats_kernel<float><<<grid, kStemTot, 0, st>>>((const float*)stems, sums, HW, pix_per_cta);
    else attn_stats_kernel<bf16><<<grid, kStemTot, 0, st>>>((const bf16*)stems, sums, HW, pix_per_cta);
    MC_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------
// mixture: one CTA per (stem, b), 64 threads.
//   y_c   = mean_c * rsqrt(var_unbiased_c + 1e-3)                     attentive_norm.py:84-85
//   a_j   = relu6(BN10(sum_c W[j][c] y_c) + 3) / 6                    attentive_norm.py:49-53,20
//   gamma = a @ weight_, beta = a @ bias_                              attentive_norm.py:159-160
//   out   = gamma * (x - rm) * rsqrt(rv + 1e-3) + beta  = coefA * x + coefB
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) attn_mix_kernel(const AttnMixParams p) {
    const int s = blockIdx.x, b = blockIdx.y, c = threadIdx.x;
    __shared__ float y[kStemC];
    __shared__ float a[kNumAff];
    const int ch = s * kStemC + c;
    const double n = (double)p.HW;
    const double sum = p.sums[((long long)b * kStemTot + ch) * 2 + 0];
    const double sq = p.sums[((long long)b * kStemTot + ch) * 2 + 1];
    const double mean = sum / n;
    double var = (sq - sum * mean) / (n - 1.0);      // unbiased (torch.var_mean default)
    if (var < 0.0) var = 0.0;
    y[c] = (float)mean * rsqrtf((float)var + 1e-3f);
    __syncthreads();
    if (c < kNumAff) {
        const float* w = p.att_w + ((long long)s * kNumAff + c) * kStemC;
        float acc = 0.f;
        for (int i = 0; i < kStemC; ++i) acc = fmaf(w[i], y[i], acc);
        acc = fmaf(acc, p.att_scale[s * kNumAff + c], p.att_shift[s * kNumAff + c]);
        a[c] = fminf(fmaxf(acc + 3.f, 0.f), 6.f) / 6.f;
    }
    __syncthreads();
    float gamma = 0.f, beta = 0.f;
    for (int j = 0; j < kNumAff; ++j) {
        gamma = fmaf(a[j], p.bank_w[((long long)s * kNumAff + j) * kStemC + c], gamma);
        beta = fmaf(a[j], p.bank_b[((long long)s * kNumAff + j) * kStemC + c], beta);
    }
    const float inv = p.bn_inv[ch], rm = p.bn_mean[ch];
    const float A = gamma * inv;
    p.coefA[(long long)b * kStemTot + ch] = A;
    p.coefB[(long long)b * kStemTot + ch] = beta - A * rm;
}

void launch_attn_mix(const AttnMixParams& p, int B, cudaStream_t st) {
    dim3 grid(kNumStems, B);
    attn_mix_kernel<<<grid, kStemC, 0, st>>>(p);
    MC_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------
// head apply: per pixel, per stem: z = relu(coefA * x + coefB) (64 ch), then the 1x1 convs that read
// that stem, bias, and the output activation; writes the ten NCHW fp32 maps.
//   rows of w / bias (pred order):  heat 0-2 (stem 0) | kpt_heat 3-11 (stem 4) | wh 12-13 (1) | offset 14-15 (2)
//   | kpt_hm_offset 16-17 (5) | center2kpt 18-35 (3) | dim 36-38 (6) | depth 39-40 (7) | alpha_cls 41-52 (8)
//   | alpha_offset 53-64 (8)
// CTA: 32 pixels x 9 stems.  Tile staged in shared memory as fp32 (pixel-major, padded).
// ---------------------------------------------------------------------------------------------
struct OutMap { int stem, pred, ch, nch, act; };   // act: 0 none, 1 sigmoid+clamp, 2 inverse-sigmoid depth
__constant__ OutMap c_outmap[kNumOut];

constexpr int kHaPix = 32;
constexpr int kHaPitch = kStemTot + 4;    // 580 floats: float4-aligned rows, bank offset 4 per pixel

template <typename T>
__global__ void __launch_bounds__(256) head_apply_kernel(const HeadApplyParams p) {
    extern __shared__ __align__(16) float smem[];
    float* zt = smem;                                  // [32][580]
    float* ws = smem + kHaPix * kHaPitch;              // [65][64]
    const int b = blockIdx.y;
    const int p0 = blockIdx.x * kHaPix;
    const int tid = threadIdx.x;
    for (int i = tid; i < kNumOut * kStemC; i += 256) ws[i] = p.w[i];
    // stage + normalise + ReLU   (coalesced over channels)
    const T* x = reinterpret_cast<const T*>(p.stems) + ((long long)b * p.HW + p0) * kStemTot;
    const float* cA = p.coefA + (long long)b * kStemTot;
    const float* cB = p.coefB + (long long)b * kStemTot;
    for (int i = tid; i < kHaPix * kStemTot; i += 256) {
        const int pix = i / kStemTot, ch = i % kStemTot;
        float v = 0.f;
        if (p0 + pix < p.HW) v = fmaxf(fmaf(cA[ch], ldf<T>(x + (long long)pix * kStemTot + ch), cB[ch]), 0.f);
        zt[pix * kHaPitch + ch] = v;
    }
    __syncthreads();
    // 65 outputs x 32 pixels; lane = pixel, warp w handles outputs w, w+8, ...
    const int lane = tid & 31, warp = tid >> 5;
    const int pix = p0 + lane;
    for (int o = warp; o < kNumOut; o += 8) {
        const OutMap om = c_outmap[o];
        const float4* zr = reinterpret_cast<const float4*>(zt + lane * kHaPitch + om.stem * kStemC);
        const float4* wr = reinterpret_cast<const float4*>(ws + o * kStemC);
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < kStemC / 4; ++k) {
            const float4 z = zr[k], w = wr[k];
            acc = fmaf(z.x, w.x, acc);
            acc = fmaf(z.y, w.y, acc);
            acc = fmaf(z.z, w.z, acc);
            acc = fmaf(z.w, w.w, acc);
        }
        acc += p.bias[o];
        if (om.act == 1) {                              // monocon_heads.py:168-170
            acc = 1.f / (1.f + expf(-acc));
            acc = fminf(fmaxf(acc, 1e-4f), 1.f - 1e-4f);
        } else if (om.act == 2) {                       // monocon_heads.py:183
            acc = 1.f / (1.f / (1.f + expf(-acc)) + 1e-12f) - 1.f;
        }
        if (pix < p.HW) p.out[om.pred][((long long)b * om.nch + om.ch) * p.HW + pix] = acc;
    }
}

void launch_head_apply(const HeadApplyParams& p, DType dt, cudaStream_t st) {
    const size_t smem = sizeof(float) * (kHaPix * kHaPitch + kNumOut * kStemC);
    dim3 grid((p.HW + kHaPix - 1) / kHaPix, p.B);
    if (dt == DT_F32) head_apply_kernel<float><<<grid, 256, smem, st>>>(p);
    else head_apply_kernel<bf16><<<grid, 256, smem, st>>>(p);
    MC_CUDA(cudaGetLastError());
}

void head_kernels_init() {
    // pred index, stem, channels (monocon_heads.py:165-200; stems in registration order :74-88)
    const int pred_stem[kNumPred] = {0, 4, 1, 2, 5, 3, 6, 7, 8, 8};
    const int pred_nch[kNumPred] = {3, 9, 2, 2, 2, 18, 3, 2, 12, 12};
    OutMap h[kNumOut];
    int o = 0;
    for (int pi = 0; pi < kNumPred; ++pi)
        for (int c = 0; c < pred_nch[pi]; ++c) {
            int act = 0;
            if (pi == 0 || pi == 1) act = 1;
            if (pi == 7 && c == 0) act = 2;
            h[o++] = OutMap{pred_stem[pi], pi, c, pred_nch[pi], act};
        }
    MC_CUDA(cudaMemcpyToSymbol(c_outmap, h, sizeof(h)));
    const int smem = (int)(sizeof(float) * (kHaPix * kHaPitch + kNumOut * kStemC));
    MC_CUDA(cudaFuncSetAttribute(head_apply_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    MC_CUDA(cudaFuncSetAttribute(head_apply_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
}

}  // namespace mc
