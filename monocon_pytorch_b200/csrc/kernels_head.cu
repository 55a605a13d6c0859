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
// grid (chunks, B); each CTA reduces a slab of pixels for all 576 channels (16-byte loads, coalesced along C),
// accumulates per-thread in fp32 over <= 32 pixels, then in fp64 through shared memory + global atomics.
// ---------------------------------------------------------------------------------------------
// 288 threads = 72 column groups (8 channels, one 16-byte load) x 4 pixel lanes.
template <typename T> __device__ __forceinline__ void ld8f(const T* p, float (&v)[8]);
template <> __device__ __forceinline__ void ld8f<float>(const float* p, float (&v)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void ld8f<bf16>(const bf16* p, float (&v)[8]) {
    const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
        v[2 * j] = f.x; v[2 * j + 1] = f.y;
    }
}

template <typename T>
__global__ void __launch_bounds__(288, 2) attn_stats_kernel(const T* __restrict__ x, double* __restrict__ sums, int HW,
                                                         int pix_per_cta) {
    __shared__ double red[4][kStemTot][2];                   // 36.9 KB
    pdl_sync();
    const int b = blockIdx.y;
    const int cg = threadIdx.x % 72, pl = threadIdx.x / 72;  // channel group (8 ch), pixel lane (0..3)
    const int p0 = blockIdx.x * pix_per_cta;
    const int p1 = min(HW, p0 + pix_per_cta);
    const T* base = x + ((long long)b * HW) * kStemTot + cg * 8;
    double s[8], ss[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] = 0.0; ss[j] = 0.0; }
    // full blocks of 128 pixels (32 per thread): fixed trip count, four independent 16-byte loads in flight per thread
    const int nfull = (p1 - p0) / 128;
    for (int blk = 0; blk < nfull; ++blk) {
        const T* bp = base + (long long)(p0 + blk * 128 + pl) * kStemTot;
        float fs[8], fss[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { fs[j] = 0.f; fss[j] = 0.f; }
#pragma unroll 2
        for (int i = 0; i < 32; i += 4) {
            float v[4][8];
#pragma unroll
            for (int u = 0; u < 4; ++u) ld8f<T>(bp + (long long)(i + u) * 4 * kStemTot, v[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int j = 0; j < 8; ++j) { fs[j] += v[u][j]; fss[j] = fmaf(v[u][j], v[u][j], fss[j]); }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j] += (double)fs[j]; ss[j] += (double)fss[j]; }
    }
    for (int q = p0 + nfull * 128 + pl; q < p1; q += 4 * 32) {   // tail: fp32 partial sums over <= 32 pixels, then fp64
        float fs[8], fss[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { fs[j] = 0.f; fss[j] = 0.f; }
        for (int pidx = q; pidx < p1 && pidx < q + 4 * 32; pidx += 4) {
            float v[8];
            ld8f<T>(base + (long long)pidx * kStemTot, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) { fs[j] += v[j]; fss[j] = fmaf(v[j], v[j], fss[j]); }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j] += (double)fs[j]; ss[j] += (double)fss[j]; }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { red[pl][cg * 8 + j][0] = s[j]; red[pl][cg * 8 + j][1] = ss[j]; }
    __syncthreads();
    for (int c = threadIdx.x; c < kStemTot; c += 288) {
        const double a = red[0][c][0] + red[1][c][0] + red[2][c][0] + red[3][c][0];
        const double q2 = red[0][c][1] + red[1][c][1] + red[2][c][1] + red[3][c][1];
        atomicAdd(&sums[((long long)b * kStemTot + c) * 2 + 0], a);
        atomicAdd(&sums[((long long)b * kStemTot + c) * 2 + 1], q2);
    }
}

void launch_attn_stats(const void* stems, DType dt, double* sums, int B, int HW, cudaStream_t st) {
    MC_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * kStemTot * B, st));
    int chunks = (148 * 4 + B - 1) / B;
    int pix_per_cta = (HW + chunks - 1) / chunks;
    if (pix_per_cta < 64) pix_per_cta = 64;
    if (pix_per_cta > 128) pix_per_cta = (pix_per_cta + 127) / 128 * 128;      // whole 128-pixel blocks (unrolled path)
    chunks = (HW + pix_per_cta - 1) / pix_per_cta;
    dim3 grid(chunks, B);
    if (dt == DT_F32) launch_k(attn_stats_kernel<float>, grid, dim3(288), 0, st, (const float*)stems, sums, HW, pix_per_cta);
    else launch_k(attn_stats_kernel<bf16>, grid, dim3(288), 0, st, (const bf16*)stems, sums, HW, pix_per_cta);
}

// ---------------------------------------------------------------------------------------------
// mixture: one CTA per (stem, b), 64 threads.
//   y_c   = mean_c * rsqrt(var_unbiased_c + 1e-3)                     attentive_norm.py:84-85
//   a_j   = relu6(BN10(sum_c W[j][c] y_c) + 3) / 6                    attentive_norm.py:49-53,20
//   gamma = a @ weight_, beta = a @ bias_                              attentive_norm.py:159-160
//   out   = gamma * (x - rm) * rsqrt(rv + 1e-3) + beta  = coefA * x + coefB
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) attn_mix_kernel(const AttnMixParams p) {
    pdl_sync();
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
    launch_k(attn_mix_kernel, grid, dim3(kStemC), 0, st, p);
}

// ---------------------------------------------------------------------------------------------
// head apply: per pixel, per stem: z = relu(coefA * x + coefB) (64 ch), then the 1x1 convs that read
// that stem, bias, and the output activation; writes the ten NCHW fp32 maps.
//   rows of w / bias (pred order):  heat 0-2 (stem 0) | kpt_heat 3-11 (stem 4) | wh 12-13 (1) | offset 14-15 (2)
//   | kpt_hm_offset 16-17 (5) | center2kpt 18-35 (3) | dim 36-38 (6) | depth 39-40 (7) | alpha_cls 41-52 (8)
//   | alpha_offset 53-64 (8)
// CTA: 9 warps (one per stem) x 32 pixels, persistent over pixel groups.
// ---------------------------------------------------------------------------------------------
struct OutMap { int stem, pred, ch, nch, act; };   // act: 0 none, 1 sigmoid+clamp, 2 inverse-sigmoid depth
__constant__ OutMap c_outmap[kNumOut];

// v2 layout: one warp per (pixel group, stem) unit, one lane per pixel.  A thread keeps its 64 normalised inputs in registers and
// walks the (contiguous) output rows that read its stem; the 1x1 weights are broadcast from shared memory.
// Reads are 128 B (bf16) / 256 B (fp32) contiguous per thread, writes are coalesced along pixels (NCHW).
constexpr int kHaPix = 32;
constexpr int kHaThreads = kNumStems * 32;            // 288

// first / one-past-last output row (pred order) of each stem, see the table above
__constant__ int c_stem_o0[kNumStems] = {0, 12, 14, 18, 3, 16, 36, 39, 41};
__constant__ int c_stem_o1[kNumStems] = {3, 14, 16, 36, 12, 18, 39, 41, 65};

template <typename T> __device__ __forceinline__ void load_row64(const T* p, float (&z)[kStemC]);
template <> __device__ __forceinline__ void load_row64<float>(const float* p, float (&z)[kStemC]) {
#pragma unroll
    for (int i = 0; i < kStemC / 4; ++i) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(p) + i);
        z[4 * i] = v.x; z[4 * i + 1] = v.y; z[4 * i + 2] = v.z; z[4 * i + 3] = v.w;
    }
}
template <> __device__ __forceinline__ void load_row64<bf16>(const bf16* p, float (&z)[kStemC]) {
#pragma unroll
    for (int i = 0; i < kStemC / 8; ++i) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(p) + i);
        const uint32_t r[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r[j]));
            z[8 * i + 2 * j] = f.x;
            z[8 * i + 2 * j + 1] = f.y;
        }
    }
}

// work unit = (32-pixel group, stem); units are strided over all warps of the (persistent) grid so that the uneven
// number of outputs per stem (2 .. 24) balances out; no block-level synchronisation inside the loop.
// The 32 x 64 slice of a unit is fetched with coalesced 16-byte loads (8 lanes cover one pixel's 128 bytes, four
// pixels per instruction) and transposed through a per-warp shared-memory tile so that each lane ends up with its
// own pixel's 64 channels in registers.
constexpr int kHaRowBytes = kStemC * 2 + 16;          // bf16 row + 16 B pad (conflict-free 16-byte column access)

template <typename T> struct HaStage;
template <> struct HaStage<bf16> {
    static constexpr int kWarpBytes = kHaPix * kHaRowBytes;            // 4608 B
    static __device__ __forceinline__ void load(const bf16* g, int rows_valid, unsigned char* tile, int lane, float (&z)[kStemC]) {
        // g: first pixel of the group at this stem's channel offset; row pitch kStemTot elements
        const int part = lane & 7, r0 = lane >> 3;                     // 8 x 16 B per pixel row, 4 rows per instruction
#pragma unroll
        for (int i = 0; i < kHaPix / 4; ++i) {
            const int r = r0 + 4 * i;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (r < rows_valid) v = __ldg(reinterpret_cast<const uint4*>(g + (long long)r * kStemTot) + part);
            *reinterpret_cast<uint4*>(tile + r * kHaRowBytes + part * 16) = v;
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < kStemC / 8; ++i) {
            const uint4 v = *reinterpret_cast<const uint4*>(tile + lane * kHaRowBytes + i * 16);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
                z[8 * i + 2 * j] = f.x;
                z[8 * i + 2 * j + 1] = f.y;
            }
        }
        __syncwarp();
    }
};
template <> struct HaStage<float> {
    static constexpr int kWarpBytes = 16;                              // unused: fp32 mode reads rows directly
    static __device__ __forceinline__ void load(const float* g, int rows_valid, unsigned char*, int lane, float (&z)[kStemC]) {
        if (lane < rows_valid) load_row64<float>(g + (long long)lane * kStemTot, z);
    }
};

template <typename T>
__global__ void __launch_bounds__(kHaThreads, 2) head_apply_kernel(const HeadApplyParams p) {
    __shared__ __align__(16) float ws[kNumOut * kStemC];      // 16.6 KB
    __shared__ float bs[kNumOut];
    __shared__ float* outp[kNumPred];
    extern __shared__ __align__(16) unsigned char stage[];       // (kHaThreads / 32) * HaStage<T>::kWarpBytes
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < kNumOut * kStemC; i += kHaThreads) ws[i] = p.w[i];
    if (tid < kNumOut) bs[tid] = p.bias[tid];
    if (tid < kNumPred) outp[tid] = p.out[tid];
    __syncthreads();
    pdl_sync();          // the 1x1 weights above are constants; stems / coefficients are produced by the previous kernels
    unsigned char* tile = stage + warp * HaStage<T>::kWarpBytes;
    const int groups_per_img = (p.HW + kHaPix - 1) / kHaPix;
    const int units = groups_per_img * p.B * kNumStems;
    const int warps_total = gridDim.x * (kHaThreads / 32);
    for (int u = blockIdx.x * (kHaThreads / 32) + warp; u < units; u += warps_total) {
        // stem-major unit order: the warp stride (a multiple of 9) must not lock a warp onto one stem -- the stems have
        // 2 .. 24 outputs each, and a warp walking u, u + stride, ... now visits every stem in proportion
        const int total_groups = groups_per_img * p.B;
        const int stem = u / total_groups;
        const int g = u % total_groups;
        const int b = g / groups_per_img, p0 = (g % groups_per_img) * kHaPix;
        const int pix = p0 + lane;
        const int rows_valid = min(kHaPix, p.HW - p0);
        float z[kStemC];
        HaStage<T>::load(reinterpret_cast<const T*>(p.stems) + ((long long)b * p.HW + p0) * kStemTot + stem * kStemC, rows_valid,
                         tile, lane, z);
        if (pix >= p.HW) continue;
        const float4* a4 = reinterpret_cast<const float4*>(p.coefA + (long long)b * kStemTot + stem * kStemC);
        const float4* c4 = reinterpret_cast<const float4*>(p.coefB + (long long)b * kStemTot + stem * kStemC);
#pragma unroll
        for (int k = 0; k < kStemC / 4; ++k) {                 // warp-uniform addresses: one broadcast transaction each
            const float4 a = __ldg(a4 + k), c = __ldg(c4 + k);
            z[4 * k + 0] = fmaxf(fmaf(a.x, z[4 * k + 0], c.x), 0.f);
            z[4 * k + 1] = fmaxf(fmaf(a.y, z[4 * k + 1], c.y), 0.f);
            z[4 * k + 2] = fmaxf(fmaf(a.z, z[4 * k + 2], c.z), 0.f);
            z[4 * k + 3] = fmaxf(fmaf(a.w, z[4 * k + 3], c.w), 0.f);
        }
        const int o0 = c_stem_o0[stem], o1 = c_stem_o1[stem];
        auto finish = [&](int o, float acc) {
            acc += bs[o];
            const OutMap om = c_outmap[o];
            if (om.act == 1) {                              // monocon_heads.py:168-170
                acc = 1.f / (1.f + expf(-acc));
                acc = fminf(fmaxf(acc, 1e-4f), 1.f - 1e-4f);
            } else if (om.act == 2) {                       // monocon_heads.py:183
                acc = 1.f / (1.f / (1.f + expf(-acc)) + 1e-12f) - 1.f;
            }
            outp[om.pred][((long long)b * om.nch + om.ch) * p.HW + pix] = acc;
        };
        int o = o0;
        for (; o + 1 < o1; o += 2) {                        // two outputs per pass: eight independent FMA chains
            const float4* wa = reinterpret_cast<const float4*>(ws + o * kStemC);
            const float4* wb = reinterpret_cast<const float4*>(ws + (o + 1) * kStemC);
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
#pragma unroll
            for (int k = 0; k < kStemC / 4; ++k) {
                const float4 u = wa[k], v = wb[k];
                a0 = fmaf(z[4 * k], u.x, a0);     b0 = fmaf(z[4 * k], v.x, b0);
                a1 = fmaf(z[4 * k + 1], u.y, a1); b1 = fmaf(z[4 * k + 1], v.y, b1);
                a2 = fmaf(z[4 * k + 2], u.z, a2); b2 = fmaf(z[4 * k + 2], v.z, b2);
                a3 = fmaf(z[4 * k + 3], u.w, a3); b3 = fmaf(z[4 * k + 3], v.w, b3);
            }
            finish(o, (a0 + a1) + (a2 + a3));
            finish(o + 1, (b0 + b1) + (b2 + b3));
        }
        if (o < o1) {
            const float4* wr = reinterpret_cast<const float4*>(ws + o * kStemC);
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int k = 0; k < kStemC / 4; ++k) {
                const float4 w = wr[k];
                a0 = fmaf(z[4 * k], w.x, a0);
                a1 = fmaf(z[4 * k + 1], w.y, a1);
                a2 = fmaf(z[4 * k + 2], w.z, a2);
                a3 = fmaf(z[4 * k + 3], w.w, a3);
            }
            finish(o, (a0 + a1) + (a2 + a3));
        }
    }
}

void launch_head_apply(const HeadApplyParams& p, DType dt, cudaStream_t st) {
    const int units = ((p.HW + kHaPix - 1) / kHaPix) * p.B * kNumStems;
    const int blocks_needed = (units + kNumStems - 1) / kNumStems;
    const int grid = blocks_needed < 148 * 2 ? blocks_needed : 148 * 2;
    if (dt == DT_F32) launch_k(head_apply_kernel<float>, dim3(grid), dim3(kHaThreads), (kHaThreads / 32) * HaStage<float>::kWarpBytes, st, p);
    else launch_k(head_apply_kernel<bf16>, dim3(grid), dim3(kHaThreads), (kHaThreads / 32) * HaStage<bf16>::kWarpBytes, st, p);
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
    MC_CUDA(cudaFuncSetAttribute(head_apply_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (kHaThreads / 32) * HaStage<bf16>::kWarpBytes));
}

}  // namespace mc
