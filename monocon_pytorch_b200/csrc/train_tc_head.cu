// Heads backward of the bf16 tensor-core training step: the same formulas as the fp32 kernel set of train_backward.cu (1x1 output
// convolutions + output transforms, ReLU, AttnBatchNorm2d incl. its 10-channel BatchNorm over the batch: monocon_heads.py:114-131,
// 165-200; model/norm/attentive_norm.py:79-91,154-164; oracle/backward_oracle.py), restructured for the device:
//   * head_reduce_tc / head_dx_tc: thread = (4 adjacent stem channels, pixel lane); the 65 x 64 matrix of the 1x1 convolutions
//     sits in shared memory, the per-(image, channel) constants in registers, so one 16-byte load of the stems feeds 4 x (o1 - o0)
//     FMAs of the 1x1 backward and as many of its weight gradient; ~10^4 blocks instead of 18 per image.
//   * head_mix_tc: one CTA per stem, one thread per channel (the fp32 twin runs one THREAD per stem).
//   * the gradient of the pre-norm stems goes out as bf16 -- it is the dy operand of the stem convolution's dgrad / wgrad -- and the
//     gradient of the stem biases is assembled from sums the pass already has: sum_pix (K0 dout + K1 x + K2) = K0 S0 + K1 sum(x) + K2 HW.
// Checked against the fp32 twin on the same inputs (tests/test_gpu_train_tc.py).
#include <algorithm>

#include "train_backward.h"
#include "train_tc.h"

namespace mc {

namespace {

constexpr int kMaxRows = 24;             // most output rows behind one stem (dir_feat: dir_cls + dir_reg)
constexpr int kMaxB = 64;

__constant__ int t_o0[kNumStems] = {0, 12, 14, 18, 3, 16, 36, 39, 41};      // first / one-past-last row of each stem's 1x1 outputs in
__constant__ int t_o1[kNumStems] = {3, 14, 16, 36, 12, 18, 39, 41, 65};      // the 65-row pred-order matrix (train_backward.cu)
__constant__ int t_pred_ch[kNumPred] = {3, 9, 2, 2, 2, 18, 3, 2, 12, 12};
__constant__ int t_pred_o0[kNumPred] = {0, 3, 12, 14, 16, 18, 36, 39, 41, 53};

struct Scratch {
    float* draw;       // [B*HW][65]
    double* S;         // [B][576][2]
    double* colsums;   // [65]
    float* meaninv;    // [576][2]
    float* K;          // [3][B][576]
};
inline size_t a256(size_t v) { return (v + 255) / 256 * 256; }
inline Scratch carve_tc(void* base, int B, int HW) {
    char* p = (char*)base;
    Scratch s;
    s.draw = (float*)p; p += a256(sizeof(float) * (size_t)B * HW * kNumOut);
    s.S = (double*)p; p += a256(sizeof(double) * (size_t)B * kStemTot * 2);
    s.colsums = (double*)p; p += a256(sizeof(double) * kNumOut);
    s.meaninv = (float*)p; p += a256(sizeof(float) * kStemTot * 2);
    s.K = (float*)p;
    return s;
}

// dL/dpred (ten NCHW maps) -> gradient of the raw 1x1 outputs, [pixel][65] (sigmoid + clamp of the two heat-maps, depth transform).
// A block transposes 128 pixels through shared memory: the reads walk the maps along the pixel index, the writes are one contiguous
// 33 KB span (a thread per pixel scattered 65 floats at a 260-byte stride: 0.55 ms, this: see profiles).
__global__ void __launch_bounds__(256) head_draw_tc_kernel(const HeadBwdParams p, float* __restrict__ draw) {
    __shared__ float tile[kNumOut][129];
    const long long Q = (long long)p.B * p.HW, q0 = (long long)blockIdx.x * 128;
    for (int idx = threadIdx.x; idx < kNumOut * 128; idx += 256) {
        const int o = idx >> 7, px = idx & 127;
        const long long q = q0 + px;
        if (q >= Q) continue;
        const int b = (int)(q / p.HW), pix = (int)(q % p.HW);
        int k = 0;
#pragma unroll
        for (int kk = 1; kk < kNumPred; ++kk) if (o >= t_pred_o0[kk]) k = kk;
        const int ch = t_pred_ch[k], j = o - t_pred_o0[k];
        const long long i = ((long long)b * ch + j) * p.HW + pix;
        float d = p.dpred[k][i];
        if (k < 2) {
            const float v = p.pred[k][i];
            d = (v > 1e-4f && v < 1.f - 1e-4f) ? d * v * (1.f - v) : 0.f;
        } else if (k == 7 && j == 0) {
            d = -d * p.pred[k][i];
        }
        tile[o][px] = d;
    }
    __syncthreads();
    const int npx = (int)min((long long)128, Q - q0);
    for (int idx = threadIdx.x; idx < npx * kNumOut; idx += 256) {
        const int px = idx / kNumOut, o = idx - px * kNumOut;
        draw[q0 * kNumOut + idx] = tile[o][px];
    }
}

// column sums of draw [Q][65] -> dbias[65]: block partials in shared memory
__global__ void __launch_bounds__(256) head_colsum_tc_kernel(const float* __restrict__ draw, long long Q, double* __restrict__ sums) {
    __shared__ float sh[kNumOut];
    for (int i = threadIdx.x; i < kNumOut; i += 256) sh[i] = 0.f;
    __syncthreads();
    const int c = threadIdx.x % 65, lane = threadIdx.x / 65;          // 3 pixel lanes x 65 columns (195 threads busy)
    if (lane < 3) {
        float acc = 0.f;
        for (long long q = (long long)blockIdx.x * 3 + lane; q < Q; q += (long long)gridDim.x * 3) acc += draw[q * kNumOut + c];
        atomicAdd(&sh[c], acc);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kNumOut; i += 256) atomicAdd(&sums[i], (double)sh[i]);
}
__global__ void head_narrow_tc_kernel(const double* __restrict__ in, float* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)in[i];
}

__global__ void head_meaninv_tc_kernel(const double* __restrict__ sums, int B, int HW, float* __restrict__ meaninv) {
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= kStemTot) return;
    double bs = 0.0, bss = 0.0;
    for (int b = 0; b < B; ++b) { bs += sums[((long long)b * kStemTot + ch) * 2]; bss += sums[((long long)b * kStemTot + ch) * 2 + 1]; }
    const double N = (double)HW * B, m = bs / N;
    double v = bss / N - m * m;
    if (v < 0.0) v = 0.0;
    meaninv[2 * ch] = (float)m;
    meaninv[2 * ch + 1] = (float)(1.0 / sqrt(v + 1e-3));
}

template <int VW> __device__ __forceinline__ void load_stems(const float* p, float (&x)[VW]) {
    if (VW == 4) {
        const float4 v = *reinterpret_cast<const float4*>(p);
        x[0] = v.x; x[1] = v.y; x[VW - 2] = v.z; x[VW - 1] = v.w;
    } else {
        const float2 v = *reinterpret_cast<const float2*>(p);
        x[0] = v.x; x[1] = v.y;
    }
}
template <int VW> __device__ __forceinline__ void load_stems(const bf16* p, float (&x)[VW]) {
    if (VW == 4) {
        const uint2 v = *reinterpret_cast<const uint2*>(p);
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.x)), b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.y));
        x[0] = a.x; x[1] = a.y; x[VW - 2] = b.x; x[VW - 1] = b.y;
    } else {
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
        x[0] = a.x; x[1] = a.y;
    }
}

// grid (pixel chunks, B, stem): one block works on ONE stem's 64 channels -- the stems have 2 ... 24 output rows behind them, so blocks
// of different stems cost very different amounts and the block scheduler balances them (a block over all 576 channels waited for
// its dir_feat warp: 24 rows against an average of 7).  256 threads = (64 / VW) channel groups x pixel lanes.
// REDUCE: S[b][ch] += (sum dout, sum dout * xhat), dw[o][c] += sum draw[o] * relu(post).  else: dstems = K0 * dout + K1 * x + K2.
// NR = the stem's row count as a template parameter (a predicated 24-row loop issued 3.3x the instructions: ncu, 1.85e9 warp
// instructions), one launch per stem.
template <bool REDUCE, int VW, int NR, typename ST>
__global__ void __launch_bounds__(256) head_pass_tc_kernel(const HeadBwdParams p, const ST* __restrict__ stems, const float* __restrict__ draw, const float* __restrict__ meaninv,
                                                          double* __restrict__ S, const float* __restrict__ K, bf16* __restrict__ out_bf16,
                                                          float* __restrict__ out_f32, int ppb, int s) {
    constexpr int kGroups = kStemC / VW, kLanes = 256 / kGroups;
    constexpr int nrow = NR;
    __shared__ __align__(16) float w_s[NR * kStemC];
    const int b = blockIdx.y;
    const int o0 = t_o0[s];
    for (int i = threadIdx.x; i < nrow * kStemC; i += 256) w_s[i] = p.w[o0 * kStemC + i];
    __syncthreads();
    const int grp = threadIdx.x % kGroups, lane = threadIdx.x / kGroups;
    const int c = grp * VW, ch0 = s * kStemC + c;
    const long long cb = (long long)b * kStemTot + ch0;
    float A[VW], Bc[VW], mean[VW], inv[VW], K0[VW], K1[VW], K2[VW];
#pragma unroll
    for (int j = 0; j < VW; ++j) {
        A[j] = p.coefA[cb + j]; Bc[j] = p.coefB[cb + j];
        if (REDUCE) {
            mean[j] = meaninv[2 * (ch0 + j)]; inv[j] = meaninv[2 * (ch0 + j) + 1];
        } else {
            const long long plane = (long long)p.B * kStemTot;
            K0[j] = K[cb + j]; K1[j] = K[plane + cb + j]; K2[j] = K[2 * plane + cb + j];
        }
    }
    float dwk[REDUCE ? NR : 1][VW];
    float f0[VW], f1[VW];
#pragma unroll
    for (int j = 0; j < VW; ++j) { f0[j] = 0.f; f1[j] = 0.f; }
    if (REDUCE) {
#pragma unroll
        for (int k = 0; k < NR; ++k)
#pragma unroll
            for (int j = 0; j < VW; ++j) dwk[k][j] = 0.f;
    }
    const int p0 = blockIdx.x * ppb, p1 = min(p.HW, p0 + ppb);
    for (int pix = p0 + lane; pix < p1; pix += kLanes) {
        const long long q = (long long)b * p.HW + pix;
        float x[VW];
        load_stems<VW>(stems + q * kStemTot + ch0, x);
        const float* drow = draw + q * kNumOut + o0;
        float post[VW], r[VW], dout[VW];
#pragma unroll
        for (int j = 0; j < VW; ++j) { post[j] = fmaf(A[j], x[j], Bc[j]); r[j] = fmaxf(post[j], 0.f); dout[j] = 0.f; }
#pragma unroll
        for (int k = 0; k < NR; ++k) {
            const float d = __ldg(drow + k);
            const float* wr = w_s + k * kStemC + c;
#pragma unroll
            for (int j = 0; j < VW; ++j) {
                dout[j] = fmaf(d, wr[j], dout[j]);
                if (REDUCE) dwk[k][j] = fmaf(d, r[j], dwk[k][j]);
            }
        }
#pragma unroll
        for (int j = 0; j < VW; ++j) if (!(post[j] > 0.f)) dout[j] = 0.f;
        if (REDUCE) {
#pragma unroll
            for (int j = 0; j < VW; ++j) { f0[j] += dout[j]; f1[j] = fmaf(dout[j], (x[j] - mean[j]) * inv[j], f1[j]); }
        } else {
            float o[VW];
#pragma unroll
            for (int j = 0; j < VW; ++j) o[j] = fmaf(K0[j], dout[j], fmaf(K1[j], x[j], K2[j]));
            if (out_bf16) {
#pragma unroll
                for (int j = 0; j < VW; j += 2)
                    *reinterpret_cast<__nv_bfloat162*>(out_bf16 + q * kStemTot + ch0 + j) = __floats2bfloat162_rn(o[j], o[j + 1]);
            }
            if (out_f32) {
#pragma unroll
                for (int j = 0; j < VW; ++j) out_f32[q * kStemTot + ch0 + j] = o[j];
            }
        }
    }
    if (REDUCE) {
        // the pixel lanes of a block meet in shared memory first (the weight tile is dead by now)
        __syncthreads();
        float* red = w_s;                                    // [kMaxRows][64] weight-gradient partials
        for (int i = threadIdx.x; i < NR * kStemC; i += 256) red[i] = 0.f;
        __shared__ float red_s[2 * kStemC];
        for (int i = threadIdx.x; i < 2 * kStemC; i += 256) red_s[i] = 0.f;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < VW; ++j) { atomicAdd(&red_s[2 * (c + j)], f0[j]); atomicAdd(&red_s[2 * (c + j) + 1], f1[j]); }
#pragma unroll
        for (int k = 0; k < NR; ++k) {
#pragma unroll
            for (int j = 0; j < VW; ++j) atomicAdd(&red[k * kStemC + c + j], dwk[k][j]);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * kStemC; i += 256) atomicAdd(&S[((long long)b * kStemTot + s * kStemC) * 2 + i], (double)red_s[i]);
        for (int i = threadIdx.x; i < nrow * kStemC; i += 256) atomicAdd(&p.dw[o0 * kStemC + i], red[i]);
    }
}

// one CTA per stem, thread c = channel c: the K = 10 mixture algebra of AttnBatchNorm2d (forward recomputed from the per-sample sums,
// then backward), the coefficients of the element-wise pass dx = K0 dout + K1 x + K2, and the gradient of the stem bias.
__global__ void __launch_bounds__(kStemC) head_mix_tc_kernel(const HeadBwdParams p, const double* __restrict__ S, const float* __restrict__ meaninv,
                                                             float* __restrict__ K, float* __restrict__ dstem_bias) {
    extern __shared__ float sm[];
    const int B = p.B, s = blockIdx.x, c = threadIdx.x, ch = s * kStemC + c;
    float* y = sm;                              // [B][64]
    float* S0 = y + B * kStemC;                 // [B][64]
    float* S1 = S0 + B * kStemC;                // [B][64]
    float* a0h = S1 + B * kStemC;               // [B][10]
    float* a1 = a0h + B * kNumAff;
    float* a = a1 + B * kNumAff;
    float* da0 = a + B * kNumAff;
    float* ainv = da0 + B * kNumAff;            // [10]
    float* dgs = ainv + kNumAff;                // [10]
    float* dbs = dgs + kNumAff;                 // [10]
    const float* attw = p.att_w + (long long)s * kNumAff * kStemC;
    const float* bw = p.bank_w + (long long)s * kNumAff * kStemC;
    const float* bb = p.bank_b + (long long)s * kNumAff * kStemC;
    const double n = (double)p.HW;
    const float hw = (float)p.HW, cntf = (float)p.HW * (float)B;
    // ---- forward, as attn_mix_train_kernel computes it ----
    for (int b = 0; b < B; ++b) {
        const long long i = ((long long)b * kStemTot + ch) * 2;
        const double sum = p.sums[i], sq = p.sums[i + 1], mean = sum / n;
        double var = (sq - sum * mean) / (n - 1.0);
        if (var < 0.0) var = 0.0;
        y[b * kStemC + c] = (float)mean * rsqrtf((float)var + 1e-3f);
        S0[b * kStemC + c] = (float)S[i];
        S1[b * kStemC + c] = (float)S[i + 1];
    }
    __syncthreads();
    for (int i = c; i < B * kNumAff; i += kStemC) {
        const int b = i / kNumAff, j = i % kNumAff;
        float acc = 0.f;
        for (int k = 0; k < kStemC; ++k) acc = fmaf(attw[j * kStemC + k], y[b * kStemC + k], acc);
        a0h[i] = acc;
    }
    __syncthreads();
    if (c < kNumAff) {
        const int j = c;
        double m = 0.0, q = 0.0;
        for (int b = 0; b < B; ++b) { const float v = a0h[b * kNumAff + j]; m += v; q += (double)v * v; }
        m /= B;
        double v = q / B - m * m;
        if (v < 0.0) v = 0.0;
        const float iv = (float)(1.0 / sqrt(v + 1e-5));
        ainv[j] = iv;
        const float g = p.att_gamma[s * kNumAff + j], be = p.att_beta[s * kNumAff + j];
        for (int b = 0; b < B; ++b) {
            const float h = (a0h[b * kNumAff + j] - (float)m) * iv;
            a0h[b * kNumAff + j] = h;
            const float t = h * g + be;
            a1[b * kNumAff + j] = t;
            a[b * kNumAff + j] = fminf(fmaxf(t + 3.f, 0.f), 6.f) / 6.f;
        }
    }
    __syncthreads();
    // ---- backward ----
    for (int j = 0; j < kNumAff; ++j) {
        float gw = 0.f, gb = 0.f;
        for (int b = 0; b < B; ++b) {
            gw = fmaf(a[b * kNumAff + j], S1[b * kStemC + c], gw);
            gb = fmaf(a[b * kNumAff + j], S0[b * kStemC + c], gb);
        }
        p.dbank_w[((long long)s * kNumAff + j) * kStemC + c] = gw;
        p.dbank_b[((long long)s * kNumAff + j) * kStemC + c] = gb;
    }
    for (int i = c; i < B * kNumAff; i += kStemC) {
        const int b = i / kNumAff, j = i % kNumAff;
        float da = 0.f;
        for (int k = 0; k < kStemC; ++k) {
            da = fmaf(S1[b * kStemC + k], bw[j * kStemC + k], da);
            da = fmaf(S0[b * kStemC + k], bb[j * kStemC + k], da);
        }
        const float t = a1[i];
        da0[i] = (t > -3.f && t < 3.f) ? da / 6.f : 0.f;                 // hardtanh backward is strict at both ends; da1 for now
    }
    __syncthreads();
    if (c < kNumAff) {
        const int j = c;
        float dg = 0.f, db = 0.f;
        for (int b = 0; b < B; ++b) { const float d1 = da0[b * kNumAff + j]; dg = fmaf(d1, a0h[b * kNumAff + j], dg); db += d1; }
        p.datt_gamma[s * kNumAff + j] = dg;
        p.datt_beta[s * kNumAff + j] = db;
        const float g = p.att_gamma[s * kNumAff + j];
        for (int b = 0; b < B; ++b)
            da0[b * kNumAff + j] = g * ainv[j] / (float)B * ((float)B * da0[b * kNumAff + j] - db - a0h[b * kNumAff + j] * dg);
    }
    __syncthreads();
    for (int j = 0; j < kNumAff; ++j) {
        float gw = 0.f;
        for (int b = 0; b < B; ++b) gw = fmaf(da0[b * kNumAff + j], y[b * kStemC + c], gw);
        p.datt_w[((long long)s * kNumAff + j) * kStemC + c] = gw;
    }
    // coefficients of the element-wise pass
    const float mean = meaninv[2 * ch], inv = meaninv[2 * ch + 1];
    float sum1 = 0.f, sum2 = 0.f;
    for (int b = 0; b < B; ++b) {
        float wt = 0.f;
        for (int j = 0; j < kNumAff; ++j) wt = fmaf(a[b * kNumAff + j], bw[j * kStemC + c], wt);
        sum1 = fmaf(wt, S0[b * kStemC + c], sum1);
        sum2 = fmaf(wt, S1[b * kStemC + c], sum2);
    }
    double bias_acc = 0.0;
    for (int b = 0; b < B; ++b) {
        float wt = 0.f, dy = 0.f;
        for (int j = 0; j < kNumAff; ++j) {
            wt = fmaf(a[b * kNumAff + j], bw[j * kStemC + c], wt);
            dy = fmaf(da0[b * kNumAff + j], attw[j * kStemC + c], dy);
        }
        const long long i = ((long long)b * kStemTot + ch) * 2;
        const double sum = p.sums[i], sq = p.sums[i + 1], im = sum / n;
        double var = (sq - sum * im) / (n - 1.0);
        if (var < 0.0) var = 0.0;
        const float r = rsqrtf((float)var + 1e-3f);
        const float dm = dy * r, dv = dy * (float)im * (-0.5f) * r * r * r;
        const long long o = (long long)b * kStemTot + ch;
        const long long plane = (long long)B * kStemTot;
        const float k0 = inv * wt;
        const float k1 = -inv * inv * sum2 / cntf + 2.f * dv / (hw - 1.f);
        const float k2 = -inv * sum1 / cntf + inv * inv * sum2 * mean / cntf + dm / hw - 2.f * dv * (float)im / (hw - 1.f);
        K[o] = k0; K[plane + o] = k1; K[2 * plane + o] = k2;
        // sum over the image's pixels of dstems = K0 sum(dout) + K1 sum(x) + K2 HW
        bias_acc += (double)k0 * (double)S0[b * kStemC + c] + (double)k1 * sum + (double)k2 * n;
    }
    if (dstem_bias) dstem_bias[ch] = (float)bias_acc;
}

}  // namespace

#define HP_LAUNCH(RED, VW, NR, GRID, MI, SS, KK, OB, OF, PPB)                                                                              \
    do {                                                                                                                                  \
        if (stems_bf16) head_pass_tc_kernel<RED, VW, NR, bf16><<<GRID, 256, 0, st>>>(p, (const bf16*)stems, sc.draw, MI, SS, KK, OB, OF, PPB, s); \
        else head_pass_tc_kernel<RED, VW, NR, float><<<GRID, 256, 0, st>>>(p, (const float*)stems, sc.draw, MI, SS, KK, OB, OF, PPB, s);    \
    } while (0)

void launch_head_backward_tc(const HeadBwdParams& p, const void* stems, bool stems_bf16, void* dstems_bf16, float* dstem_bias, cudaStream_t st) {
    MC_CHECK(p.B >= 2 && p.B <= kMaxB && p.HW >= 2, "head_backward_tc: 2 <= B <= 64");
    const Scratch sc = carve_tc(p.scratch, p.B, p.HW);
    const long long Q = (long long)p.B * p.HW;
    head_draw_tc_kernel<<<(unsigned)((Q + 127) / 128), 256, 0, st>>>(p, sc.draw);
    MC_CUDA(cudaGetLastError());
    MC_CUDA(cudaMemsetAsync(sc.colsums, 0, sizeof(double) * kNumOut, st));
    head_colsum_tc_kernel<<<(unsigned)std::min<long long>((Q + 2) / 3, 148 * 8), 256, 0, st>>>(sc.draw, Q, sc.colsums);
    MC_CUDA(cudaGetLastError());
    head_narrow_tc_kernel<<<1, 128, 0, st>>>(sc.colsums, p.dbias, kNumOut);
    MC_CUDA(cudaGetLastError());
    head_meaninv_tc_kernel<<<(kStemTot + 63) / 64, 64, 0, st>>>(p.sums, p.B, p.HW, sc.meaninv);
    MC_CUDA(cudaGetLastError());
    MC_CUDA(cudaMemsetAsync(sc.S, 0, sizeof(double) * (size_t)p.B * kStemTot * 2, st));
    MC_CUDA(cudaMemsetAsync(p.dw, 0, sizeof(float) * kNumOut * kStemC, st));
    const int ppb = 1024;
    const dim3 grid((p.HW + ppb - 1) / ppb, p.B);
    static const int rows[kNumStems] = {3, 2, 2, 18, 9, 2, 3, 2, 24};     // t_o1 - t_o0
    for (int s = 0; s < kNumStems; ++s) {
        switch (rows[s]) {
            case 2: HP_LAUNCH(true, 2, 2, grid, sc.meaninv, sc.S, nullptr, nullptr, nullptr, ppb); break;
            case 3: HP_LAUNCH(true, 2, 3, grid, sc.meaninv, sc.S, nullptr, nullptr, nullptr, ppb); break;
            case 9: HP_LAUNCH(true, 2, 9, grid, sc.meaninv, sc.S, nullptr, nullptr, nullptr, ppb); break;
            case 18: HP_LAUNCH(true, 2, 18, grid, sc.meaninv, sc.S, nullptr, nullptr, nullptr, ppb); break;
            default: HP_LAUNCH(true, 2, 24, grid, sc.meaninv, sc.S, nullptr, nullptr, nullptr, ppb); break;
        }
        MC_CUDA(cudaGetLastError());
    }
    const size_t mix_smem = sizeof(float) * ((size_t)3 * p.B * kStemC + 4 * (size_t)p.B * kNumAff + 3 * kNumAff);
    static bool attr_set = false;
    if (!attr_set) {
        MC_CUDA(cudaFuncSetAttribute(head_mix_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(float) * (3 * kMaxB * kStemC + 4 * kMaxB * kNumAff + 3 * kNumAff))));
        attr_set = true;
    }
    head_mix_tc_kernel<<<kNumStems, kStemC, mix_smem, st>>>(p, sc.S, sc.meaninv, sc.K, dstem_bias);
    MC_CUDA(cudaGetLastError());
    const int ppb2 = 256;
    const dim3 grid2((p.HW + ppb2 - 1) / ppb2, p.B);
    for (int s = 0; s < kNumStems; ++s) {
        bf16* ob = (bf16*)dstems_bf16;
        switch (rows[s]) {
            case 2: HP_LAUNCH(false, 4, 2, grid2, nullptr, nullptr, sc.K, ob, p.dstems, ppb2); break;
            case 3: HP_LAUNCH(false, 4, 3, grid2, nullptr, nullptr, sc.K, ob, p.dstems, ppb2); break;
            case 9: HP_LAUNCH(false, 4, 9, grid2, nullptr, nullptr, sc.K, ob, p.dstems, ppb2); break;
            case 18: HP_LAUNCH(false, 4, 18, grid2, nullptr, nullptr, sc.K, ob, p.dstems, ppb2); break;
            default: HP_LAUNCH(false, 4, 24, grid2, nullptr, nullptr, sc.K, ob, p.dstems, ppb2); break;
        }
        MC_CUDA(cudaGetLastError());
    }
}

}  // namespace mc
