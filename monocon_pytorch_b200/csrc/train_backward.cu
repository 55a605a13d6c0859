// Backward kernels of the training step: one kernel family per formula of oracle/backward_oracle.py (the autograd-free
// restatement pinned to the reference's own gradients).  See train_backward.h for the status: correctness-first fp32 kernels,
// checked on the CPU under tests/host_shim and on the B200 (tests/test_gpu_zz_train_backward.py, all strict), called by the engine
// (mc_backward_train).
//
// Style rule of this file: no shared memory, no __syncthreads, no warp intrinsics -- threads are independent and meet only in
// atomicAdd.  That is what lets the same bodies run sequentially under the host shim; it costs reuse (every operand comes
// through L1/L2), which is the thing to fix once the results are pinned on the device (tensor-core dgrad/wgrad: DESIGN.md 9).
#ifdef MC_HOST_SHIM
#include "host_shim.h"
#define MC_LAUNCH(kernel, grid, block, st, ...) mc::launch_k(kernel, grid, block, 0, st, __VA_ARGS__)
#else
#include "common.cuh"
#define MC_LAUNCH(kernel, grid, block, st, ...)            \
    do {                                                   \
        kernel<<<grid, block, 0, st>>>(__VA_ARGS__);       \
        MC_CUDA(cudaGetLastError());                       \
    } while (0)
namespace mc {
static inline void zero_async(void* p, size_t bytes, cudaStream_t st) { MC_CUDA(cudaMemsetAsync(p, 0, bytes, st)); }
static inline int sm_count() { return 148; }
}  // namespace mc
#endif
#include "../../include/monocon_b200.h"
#include "train_backward.h"

#include <cstdint>
#include <cstring>
#include <string>

namespace mc {

namespace {

constexpr int kThreads = 256;
constexpr int kMaxTrainB = 64;

__device__ __forceinline__ int imin(int a, int b) { return a < b ? a : b; }

// first output row of each stem's 1x1 convolutions in the 65-row pred-order weight matrix, and one past the last
__constant__ int c_o0[kNumStems] = {0, 12, 14, 18, 3, 16, 36, 39, 41};
__constant__ int c_o1[kNumStems] = {3, 14, 16, 36, 12, 18, 39, 41, 65};
// pred index and channel count of each pred, row offset in the 65-row matrix
__constant__ int c_pred_ch[kNumPred] = {3, 9, 2, 2, 2, 18, 3, 2, 12, 12};
__constant__ int c_pred_o0[kNumPred] = {0, 3, 12, 14, 16, 18, 36, 39, 41, 53};

// ---------------------------------------------------------------------------------------------
// convolution
// ---------------------------------------------------------------------------------------------
// dw[tap][ci][co] += sum over the output rows of this slice of x[n, oy*s-p+ky, ox*s-p+kx, ci] * dy[n, oy, ox, co]
// grid (ceil(k*k*Cin*Cout / 256), row slices); one thread per weight element, co fastest (dy loads coalesce, x broadcasts).
__global__ void __launch_bounds__(kThreads) conv_wgrad_kernel(const ConvBwdParams p, int rows_per_slice) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)p.k * p.k * p.Cin * p.Cout;
    if (e >= total) return;
    const int co = (int)(e % p.Cout);
    const long long t = e / p.Cout;
    const int ci = (int)(t % p.Cin), tap = (int)(t / p.Cin);
    const int ky = tap / p.k, kx = tap % p.k;
    int s = 0, c = ci;
    while (c >= p.srcC[s]) { c -= p.srcC[s]; ++s; }
    const float* sp = p.src[s];
    const int Cs = p.srcC[s], Wp = p.srcWp[s], xo = p.srcXoff[s];
    const int rows = p.B * p.Hout;
    const int r0 = (int)blockIdx.y * rows_per_slice, r1 = imin(rows, r0 + rows_per_slice);
    float acc = 0.f;
    for (int r = r0; r < r1; ++r) {
        const int n = r / p.Hout, oy = r % p.Hout;
        const int iy = oy * p.stride - p.pad + ky;
        if (iy < 0 || iy >= p.Hin) continue;
        const float* xrow = sp + ((long long)(n * p.Hin + iy) * Wp + xo) * Cs + c;
        const float* dyrow = p.dy + (long long)r * p.Wout * p.Cout + co;
        float row = 0.f;                                  // blocked summation: one partial per output row keeps the fp32 chains short
        for (int ox = 0; ox < p.Wout; ++ox) {
            const int ix = ox * p.stride - p.pad + kx;
            if (ix < 0 || ix >= p.Win) continue;
            row = fmaf(xrow[(long long)ix * Cs], dyrow[(long long)ox * p.Cout], row);
        }
        acc += row;
    }
    atomicAdd(&p.dw[e], acc);
}

// dsrc[n, iy, ix, ci] += sum over (ky, kx, co) of dy[n, (iy+p-ky)/s, (ix+p-kx)/s, co] * w[tap][ci][co]   (where the division is exact)
// one thread per input element, ci fastest
__global__ void __launch_bounds__(kThreads) conv_dgrad_kernel(const ConvBwdParams p) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)p.B * p.Hin * p.Win * p.Cin;
    if (e >= total) return;
    const int ci = (int)(e % p.Cin);
    long long t = e / p.Cin;
    const int ix = (int)(t % p.Win);
    t /= p.Win;
    const int iy = (int)(t % p.Hin), n = (int)(t / p.Hin);
    int s = 0, c = ci;
    while (c >= p.srcC[s]) { c -= p.srcC[s]; ++s; }
    if (!p.dsrc[s]) return;
    float acc = 0.f;
    for (int ky = 0; ky < p.k; ++ky) {
        const int ty = iy + p.pad - ky;
        if (ty < 0 || ty % p.stride) continue;
        const int oy = ty / p.stride;
        if (oy >= p.Hout) continue;
        for (int kx = 0; kx < p.k; ++kx) {
            const int tx = ix + p.pad - kx;
            if (tx < 0 || tx % p.stride) continue;
            const int ox = tx / p.stride;
            if (ox >= p.Wout) continue;
            const float* dyp = p.dy + ((long long)(n * p.Hout + oy) * p.Wout + ox) * p.Cout;
            float tap = 0.f;                              // one partial per tap (chains of Cout, then k*k)
            if (p.wT) {                                   // [tap][co][ci]: coalesced across the warp's input channels
                const float* wp = p.wT + (long long)(ky * p.k + kx) * p.Cout * p.Cin + ci;
                for (int co = 0; co < p.Cout; ++co) tap = fmaf(dyp[co], wp[(long long)co * p.Cin], tap);
            } else {
                const float* wp = p.w + ((long long)(ky * p.k + kx) * p.Cin + ci) * p.Cout;
                for (int co = 0; co < p.Cout; ++co) tap = fmaf(dyp[co], wp[co], tap);
            }
            acc += tap;
        }
    }
    p.dsrc[s][((long long)(n * p.Hin + iy) * p.Win + ix) * p.srcC[s] + c] += acc;
}

// The same two kernels with four channels per thread (16-byte loads of dy / wT next to one broadcast scalar: 4 FMAs per 2 loads
// instead of 1): used when every channel count involved is a multiple of 4, which is every layer of this network.
__global__ void __launch_bounds__(kThreads) conv_wgrad4_kernel(const ConvBwdParams p, int rows_per_slice) {
    const int Co4 = p.Cout / 4;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)p.k * p.k * p.Cin * Co4;
    if (e >= total) return;
    const int co = (int)(e % Co4) * 4;
    const long long t = e / Co4;
    const int ci = (int)(t % p.Cin), tap = (int)(t / p.Cin);
    const int ky = tap / p.k, kx = tap % p.k;
    int s = 0, c = ci;
    while (c >= p.srcC[s]) { c -= p.srcC[s]; ++s; }
    const float* sp = p.src[s];
    const int Cs = p.srcC[s], Wp = p.srcWp[s], xo = p.srcXoff[s];
    const int rows = p.B * p.Hout;
    const int r0 = (int)blockIdx.y * rows_per_slice, r1 = imin(rows, r0 + rows_per_slice);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int r = r0; r < r1; ++r) {
        const int n = r / p.Hout, oy = r % p.Hout;
        const int iy = oy * p.stride - p.pad + ky;
        if (iy < 0 || iy >= p.Hin) continue;
        const float* xrow = sp + ((long long)(n * p.Hin + iy) * Wp + xo) * Cs + c;
        const float* dyrow = p.dy + (long long)r * p.Wout * p.Cout + co;
        float r0a = 0.f, r1a = 0.f, r2a = 0.f, r3a = 0.f;
        for (int ox = 0; ox < p.Wout; ++ox) {
            const int ix = ox * p.stride - p.pad + kx;
            if (ix < 0 || ix >= p.Win) continue;
            const float xv = xrow[(long long)ix * Cs];
            const float4 d = *reinterpret_cast<const float4*>(dyrow + (long long)ox * p.Cout);
            r0a = fmaf(xv, d.x, r0a); r1a = fmaf(xv, d.y, r1a); r2a = fmaf(xv, d.z, r2a); r3a = fmaf(xv, d.w, r3a);
        }
        a0 += r0a; a1 += r1a; a2 += r2a; a3 += r3a;
    }
    float* o = p.dw + ((long long)tap * p.Cin + ci) * p.Cout + co;
    atomicAdd(o, a0); atomicAdd(o + 1, a1); atomicAdd(o + 2, a2); atomicAdd(o + 3, a3);
}

__global__ void __launch_bounds__(kThreads) conv_dgrad4_kernel(const ConvBwdParams p) {
    const int Ci4 = p.Cin / 4;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)p.B * p.Hin * p.Win * Ci4;
    if (e >= total) return;
    const int ci = (int)(e % Ci4) * 4;
    long long t = e / Ci4;
    const int ix = (int)(t % p.Win);
    t /= p.Win;
    const int iy = (int)(t % p.Hin), n = (int)(t / p.Hin);
    int s = 0, c = ci;
    while (c >= p.srcC[s]) { c -= p.srcC[s]; ++s; }
    if (!p.dsrc[s]) return;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int ky = 0; ky < p.k; ++ky) {
        const int ty = iy + p.pad - ky;
        if (ty < 0 || ty % p.stride) continue;
        const int oy = ty / p.stride;
        if (oy >= p.Hout) continue;
        for (int kx = 0; kx < p.k; ++kx) {
            const int tx = ix + p.pad - kx;
            if (tx < 0 || tx % p.stride) continue;
            const int ox = tx / p.stride;
            if (ox >= p.Wout) continue;
            const float* dyp = p.dy + ((long long)(n * p.Hout + oy) * p.Wout + ox) * p.Cout;
            const float* wp = p.wT + (long long)(ky * p.k + kx) * p.Cout * p.Cin + ci;
            float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
            for (int co = 0; co < p.Cout; ++co) {
                const float d = dyp[co];
                const float4 w4 = *reinterpret_cast<const float4*>(wp + (long long)co * p.Cin);
                t0 = fmaf(d, w4.x, t0); t1 = fmaf(d, w4.y, t1); t2 = fmaf(d, w4.z, t2); t3 = fmaf(d, w4.w, t3);
            }
            a0 += t0; a1 += t1; a2 += t2; a3 += t3;
        }
    }
    float4* o = reinterpret_cast<float4*>(p.dsrc[s] + ((long long)(n * p.Hin + iy) * p.Win + ix) * p.srcC[s] + c);
    float4 v = *o;
    v.x += a0; v.y += a1; v.z += a2; v.w += a3;
    *o = v;
}

// wT[tap][co][ci] = w[tap][ci][co]
__global__ void __launch_bounds__(kThreads) conv_wT_kernel(const float* __restrict__ w, float* __restrict__ wT, int taps, int Cin, int Cout) {
    const long long total = (long long)taps * Cin * Cout;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(e % Cin);
        const long long t = e / Cin;
        const int co = (int)(t % Cout), tap = (int)(t / Cout);
        wT[e] = w[((long long)tap * Cin + ci) * Cout + co];
    }
}

// ---------------------------------------------------------------------------------------------
// BatchNorm (train mode) + ReLU + residual
// ---------------------------------------------------------------------------------------------
// thread = (channel, pixel slice): sums[c] += (sum dz, sum dz * xhat), dz = dy masked by the ReLU
__global__ void __launch_bounds__(kThreads) bn_bwd_reduce_kernel(const BnBwdParams p) {
    const long long T = (long long)gridDim.x * blockDim.x, g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nsl = T / p.C;
    const int c = (int)(g % p.C);
    const long long sl = g / p.C;
    if (sl >= nsl) return;
    const float mean = p.mean[c], inv = p.inv[c];
    double s = 0.0, ss = 0.0;
    float fs = 0.f, fss = 0.f;
    int cnt = 0;
    for (long long pix = sl; pix < p.P; pix += nsl) {
        const long long i = pix * p.C + c;
        float dz = p.dy[i];
        if (p.relu && !(p.y[i] > 0.f)) dz = 0.f;
        const float xh = (p.raw[i] - mean) * inv;
        fs += dz;
        fss = fmaf(dz, xh, fss);
        if (++cnt == 64) { s += (double)fs; ss += (double)fss; fs = 0.f; fss = 0.f; cnt = 0; }
    }
    s += (double)fs; ss += (double)fss;
    atomicAdd(&p.sums[2 * c], s);
    atomicAdd(&p.sums[2 * c + 1], ss);
}

// draw = gamma * inv * (dz - (sum dz + xhat * sum dz*xhat) / P);  dres += dz;  dgamma = sum dz*xhat;  dbeta = sum dz
__global__ void __launch_bounds__(kThreads) bn_bwd_apply_kernel(const BnBwdParams p) {
    const long long total = p.P * p.C;
    const float rn = (float)(1.0 / (double)p.P);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % p.C);
        float dz = p.dy[i];
        if (p.relu && !(p.y[i] > 0.f)) dz = 0.f;
        const float inv = p.inv[c];
        const float xh = (p.raw[i] - p.mean[c]) * inv;
        const float s0 = (float)p.sums[2 * c], s1 = (float)p.sums[2 * c + 1];
        const float g = p.gamma ? p.gamma[c] : 1.f;
        p.draw[i] = g * inv * (dz - (s0 + xh * s1) * rn);
        if (p.dres) p.dres[i] += dz;
        if (i < p.C) {
            if (p.dgamma) p.dgamma[c] = s1;
            if (p.dbeta) p.dbeta[c] = s0;
        }
    }
}

__global__ void __launch_bounds__(kThreads) colsum_kernel(const float* __restrict__ x, long long P, int C, double* __restrict__ sums) {
    const long long T = (long long)gridDim.x * blockDim.x, g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nsl = T / C;
    const int c = (int)(g % C);
    const long long sl = g / C;
    if (sl >= nsl) return;
    double s = 0.0;
    float fs = 0.f;
    int cnt = 0;
    for (long long pix = sl; pix < P; pix += nsl) {
        fs += x[pix * C + c];
        if (++cnt == 64) { s += (double)fs; fs = 0.f; cnt = 0; }
    }
    s += (double)fs;
    atomicAdd(&sums[c], s);
}

__global__ void __launch_bounds__(kThreads) narrow_kernel(const double* __restrict__ in, float* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)in[i];
}

// ---------------------------------------------------------------------------------------------
// pool / upsample
// ---------------------------------------------------------------------------------------------
// one thread per (output pixel, channel); windows are disjoint, so the += has no conflicts
__global__ void __launch_bounds__(kThreads) maxpool2_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx,
                                                                int B, int C, int Hin, int Win) {
    const int Ho = Hin / 2, Wo = Win / 2;
    const long long total = (long long)B * Ho * Wo * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        long long t = i / C;
        const int ox = (int)(t % Wo);
        t /= Wo;
        const int oy = (int)(t % Ho), n = (int)(t / Ho);
        long long best = -1;
        float bv = 0.f;
        for (int ky = 0; ky < 2; ++ky)
            for (int kx = 0; kx < 2; ++kx) {
                const long long j = ((long long)(n * Hin + 2 * oy + ky) * Win + 2 * ox + kx) * C + c;
                const float v = x[j];
                if (best < 0 || v > bv) { best = j; bv = v; }      // strictly greater: the first maximum wins (ATen)
            }
        dx[best] += dy[i];
    }
}

// thread = (channel, pixel slice): dx per input pixel, the 16 weight-gradient partials in registers, 16 atomics at the end
__global__ void __launch_bounds__(kThreads) upsample2_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ dy,
                                                                 float* __restrict__ dx, float* __restrict__ dw, int B, int C, int Hin, int Win) {
    const long long T = (long long)gridDim.x * blockDim.x, g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nsl = T / C;
    const int c = (int)(g % C);
    const long long sl = g / C;
    if (sl >= nsl) return;
    float wk[16], dwk[16];
    for (int k = 0; k < 16; ++k) { wk[k] = w[c * 16 + k]; dwk[k] = 0.f; }
    const long long P = (long long)B * Hin * Win;
    const int Ho = 2 * Hin, Wo = 2 * Win;
    for (long long pix = sl; pix < P; pix += nsl) {
        const int j = (int)(pix % Win);
        const long long t = pix / Win;
        const int i = (int)(t % Hin), n = (int)(t / Hin);
        const float xv = x[pix * C + c];
        float acc = 0.f;
        for (int ky = 0; ky < 4; ++ky) {
            const int oy = 2 * i - 1 + ky;
            if (oy < 0 || oy >= Ho) continue;
            for (int kx = 0; kx < 4; ++kx) {
                const int ox = 2 * j - 1 + kx;
                if (ox < 0 || ox >= Wo) continue;
                const float d = dy[((long long)(n * Ho + oy) * Wo + ox) * C + c];
                acc = fmaf(d, wk[ky * 4 + kx], acc);
                dwk[ky * 4 + kx] = fmaf(xv, d, dwk[ky * 4 + kx]);
            }
        }
        dx[pix * C + c] += acc;
    }
    for (int k = 0; k < 16; ++k) atomicAdd(&dw[c * 16 + k], dwk[k]);
}

// ---------------------------------------------------------------------------------------------
// heads
// ---------------------------------------------------------------------------------------------
struct HeadScratch {          // carved from HeadBwdParams::scratch
    float* draw;              // [B*HW][65]  gradient of the raw 1x1 outputs
    double* S;                // [B][576][2] per-sample (sum dout, sum dout * xhat), dout = gradient of the AttnBN output
    double* colsums;          // [65]
    float* meaninv;           // [576][2]    batch mean / rsqrt(var + 1e-3) of the base BN
    float* K;                 // [3][B][576] dx = K0 * dout + K1 * x + K2
    float* mix;               // [9][kMixFloats] work arrays of head_mix_bwd_kernel (global memory: a 26 KB stack frame per thread
                              // would make the driver reserve that much local memory for every resident thread of the device)
};
constexpr int kMixFloats = kMaxTrainB * kStemC + 4 * kMaxTrainB * kNumAff;
__host__ __device__ inline size_t align256(size_t v) { return (v + 255) / 256 * 256; }
inline HeadScratch carve(void* base, int B, int HW) {
    char* p = (char*)base;
    HeadScratch s;
    s.draw = (float*)p; p += align256(sizeof(float) * (size_t)B * HW * kNumOut);
    s.S = (double*)p; p += align256(sizeof(double) * (size_t)B * kStemTot * 2);
    s.colsums = (double*)p; p += align256(sizeof(double) * kNumOut);
    s.meaninv = (float*)p; p += align256(sizeof(float) * kStemTot * 2);
    s.K = (float*)p; p += align256(sizeof(float) * 3 * (size_t)B * kStemTot);
    s.mix = (float*)p;
    return s;
}

// one thread per pixel: dL/dpred -> gradient of the raw 1x1 outputs through sigmoid+clamp (rows 0..11) and the depth transform (row 39)
__global__ void __launch_bounds__(kThreads) head_draw_kernel(const HeadBwdParams p, float* __restrict__ draw) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (long long)p.B * p.HW) return;
    const int b = (int)(q / p.HW), pix = (int)(q % p.HW);
    for (int k = 0; k < kNumPred; ++k) {
        const int ch = c_pred_ch[k];
        for (int j = 0; j < ch; ++j) {
            const long long i = ((long long)b * ch + j) * p.HW + pix;
            float d = p.dpred[k][i];
            if (k < 2) {                                     // center / keypoint heat-maps: clamp(sigmoid(z), 1e-4, 1 - 1e-4)
                const float v = p.pred[k][i];
                d = (v > 1e-4f && v < 1.f - 1e-4f) ? d * v * (1.f - v) : 0.f;
            } else if (k == 7 && j == 0) {                   // depth = 1 / (s + 1e-12) - 1 with s = sigmoid(z):  d depth / dz =
                d = -d * p.pred[k][i];                       // -s (1 - s) / (s + 1e-12)^2 = -(1 - s) / s = -depth   (1e-12 << fp32 ulp of s)
            }
            draw[q * kNumOut + c_pred_o0[k] + j] = d;
        }
    }
}

__global__ void head_meaninv_kernel(const double* __restrict__ sums, int B, int HW, float* __restrict__ meaninv) {
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

// gradient of the AttnBN output at one element: dout = relu'(A x + B) * sum over the stem's output rows of draw * w
__device__ __forceinline__ float head_dout(const float* __restrict__ drow, const float* __restrict__ w, int o0, int o1, int c, float post) {
    if (!(post > 0.f)) return 0.f;
    float d = 0.f;
    for (int o = o0; o < o1; ++o) d = fmaf(drow[o], w[o * kStemC + c], d);
    return d;
}

// grid (slices, B); thread = (stem channel, pixel slice of image b): S[b][ch] += (dout, dout * xhat); dw[o][c] += draw[o] * relu(post)
__global__ void __launch_bounds__(kThreads) head_reduce_kernel(const HeadBwdParams p, const float* __restrict__ draw, const float* __restrict__ meaninv,
                                                               double* __restrict__ S) {
    const long long T = (long long)gridDim.x * blockDim.x, g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nsl = T / kStemTot;
    const int ch = (int)(g % kStemTot);
    const long long sl = g / kStemTot;
    if (sl >= nsl) return;
    const int b = blockIdx.y, s = ch / kStemC, c = ch % kStemC;
    const int o0 = c_o0[s], o1 = c_o1[s];
    const float A = p.coefA[(long long)b * kStemTot + ch], Bc = p.coefB[(long long)b * kStemTot + ch];
    const float mean = meaninv[2 * ch], inv = meaninv[2 * ch + 1];
    float dwk[24];
    for (int k = 0; k < 24; ++k) dwk[k] = 0.f;
    double s0 = 0.0, s1 = 0.0;
    float f0 = 0.f, f1 = 0.f;
    int cnt = 0;
    for (long long pix = sl; pix < p.HW; pix += nsl) {
        const long long q = (long long)b * p.HW + pix;
        const float x = p.stems[q * kStemTot + ch];
        const float post = fmaf(A, x, Bc);
        const float* drow = draw + q * kNumOut;
        const float dout = head_dout(drow, p.w, o0, o1, c, post);
        const float r = post > 0.f ? post : 0.f;
        for (int o = o0; o < o1; ++o) dwk[o - o0] = fmaf(drow[o], r, dwk[o - o0]);
        f0 += dout;
        f1 = fmaf(dout, (x - mean) * inv, f1);
        if (++cnt == 64) { s0 += (double)f0; s1 += (double)f1; f0 = 0.f; f1 = 0.f; cnt = 0; }
    }
    s0 += (double)f0; s1 += (double)f1;
    atomicAdd(&S[((long long)b * kStemTot + ch) * 2], s0);
    atomicAdd(&S[((long long)b * kStemTot + ch) * 2 + 1], s1);
    for (int o = o0; o < o1; ++o) atomicAdd(&p.dw[o * kStemC + c], dwk[o - o0]);
}

// one thread per stem: the K = 10 mixture algebra of AttnBatchNorm2d, forward recomputed from the per-sample sums, then backward.
// Outputs the parameter gradients of the stem's AttnBN and the per-(image, channel) coefficients of dx = K0 * dout + K1 * x + K2.
__global__ void head_mix_bwd_kernel(const HeadBwdParams p, const double* __restrict__ S, const float* __restrict__ meaninv, float* __restrict__ K,
                                    float* __restrict__ mix) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= kNumStems) return;
    const int B = p.B;
    const float hw = (float)p.HW, cntf = (float)p.HW * (float)B;
    float* base = mix + (long long)s * kMixFloats;
    float (*y)[kStemC] = reinterpret_cast<float (*)[kStemC]>(base);
    float (*a0h)[kNumAff] = reinterpret_cast<float (*)[kNumAff]>(base + kMaxTrainB * kStemC);
    float (*a1)[kNumAff] = a0h + kMaxTrainB;
    float (*a)[kNumAff] = a1 + kMaxTrainB;
    float (*da0)[kNumAff] = a + kMaxTrainB;
    float ainv[kNumAff];
    const float* attw = p.att_w + (long long)s * kNumAff * kStemC;
    const float* bw = p.bank_w + (long long)s * kNumAff * kStemC;
    const float* bb = p.bank_b + (long long)s * kNumAff * kStemC;
    const double n = (double)p.HW;
    // ---- forward, as attn_mix_train_kernel computes it -----------------------------------------------------------------
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < kStemC; ++c) {
            const long long i = ((long long)b * kStemTot + s * kStemC + c) * 2;
            const double sum = p.sums[i], sq = p.sums[i + 1], mean = sum / n;
            double var = (sq - sum * mean) / (n - 1.0);
            if (var < 0.0) var = 0.0;
            y[b][c] = (float)mean * rsqrtf((float)var + 1e-3f);
        }
    for (int j = 0; j < kNumAff; ++j) {
        double m = 0.0, q = 0.0;
        for (int b = 0; b < B; ++b) {
            float acc = 0.f;
            for (int c = 0; c < kStemC; ++c) acc = fmaf(attw[j * kStemC + c], y[b][c], acc);
            a0h[b][j] = acc;
            m += acc; q += (double)acc * acc;
        }
        m /= B;
        double v = q / B - m * m;
        if (v < 0.0) v = 0.0;
        ainv[j] = (float)(1.0 / sqrt(v + 1e-5));
        const float g = p.att_gamma[s * kNumAff + j], be = p.att_beta[s * kNumAff + j];
        for (int b = 0; b < B; ++b) {
            a0h[b][j] = (a0h[b][j] - (float)m) * ainv[j];
            a1[b][j] = a0h[b][j] * g + be;
            a[b][j] = fminf(fmaxf(a1[b][j] + 3.f, 0.f), 6.f) / 6.f;
        }
    }
    // ---- backward --------------------------------------------------------------------------------------------------------
    // mixture banks and the gradient of the mixture weights
    for (int j = 0; j < kNumAff; ++j) {
        for (int c = 0; c < kStemC; ++c) {
            float gw = 0.f, gb = 0.f;
            for (int b = 0; b < B; ++b) {
                const long long i = ((long long)b * kStemTot + s * kStemC + c) * 2;
                gw = fmaf(a[b][j], (float)S[i + 1], gw);
                gb = fmaf(a[b][j], (float)S[i], gb);
            }
            p.dbank_w[((long long)s * kNumAff + j) * kStemC + c] = gw;
            p.dbank_b[((long long)s * kNumAff + j) * kStemC + c] = gb;
        }
        float dg = 0.f, db = 0.f;
        for (int b = 0; b < B; ++b) {
            float da = 0.f;
            for (int c = 0; c < kStemC; ++c) {
                const long long i = ((long long)b * kStemTot + s * kStemC + c) * 2;
                da = fmaf((float)S[i + 1], bw[j * kStemC + c], da);
                da = fmaf((float)S[i], bb[j * kStemC + c], da);
            }
            const float d1 = (a1[b][j] > -3.f && a1[b][j] < 3.f) ? da / 6.f : 0.f;     // hardtanh backward is strict at both ends
            da0[b][j] = d1;                                                             // da1 for now
            dg = fmaf(d1, a0h[b][j], dg);
            db += d1;
        }
        p.datt_gamma[s * kNumAff + j] = dg;
        p.datt_beta[s * kNumAff + j] = db;
        const float g = p.att_gamma[s * kNumAff + j];
        for (int b = 0; b < B; ++b) da0[b][j] = g * ainv[j] / (float)B * ((float)B * da0[b][j] - db - a0h[b][j] * dg);
        for (int c = 0; c < kStemC; ++c) {
            float gw = 0.f;
            for (int b = 0; b < B; ++b) gw = fmaf(da0[b][j], y[b][c], gw);
            p.datt_w[((long long)s * kNumAff + j) * kStemC + c] = gw;
        }
    }
    // coefficients of the element-wise pass
    for (int c = 0; c < kStemC; ++c) {
        const int ch = s * kStemC + c;
        const float mean = meaninv[2 * ch], inv = meaninv[2 * ch + 1];
        float sum1 = 0.f, sum2 = 0.f;
        for (int b = 0; b < B; ++b) {
            float wt = 0.f;
            for (int j = 0; j < kNumAff; ++j) wt = fmaf(a[b][j], bw[j * kStemC + c], wt);
            const long long i = ((long long)b * kStemTot + ch) * 2;
            sum1 = fmaf(wt, (float)S[i], sum1);
            sum2 = fmaf(wt, (float)S[i + 1], sum2);
        }
        for (int b = 0; b < B; ++b) {
            float wt = 0.f, dy = 0.f;
            for (int j = 0; j < kNumAff; ++j) {
                wt = fmaf(a[b][j], bw[j * kStemC + c], wt);
                dy = fmaf(da0[b][j], attw[j * kStemC + c], dy);
            }
            const long long i = ((long long)b * kStemTot + ch) * 2;
            const double sum = p.sums[i], sq = p.sums[i + 1], im = sum / n;
            double var = (sq - sum * im) / (n - 1.0);
            if (var < 0.0) var = 0.0;
            const float r = rsqrtf((float)var + 1e-3f);
            const float dm = dy * r, dv = dy * (float)im * (-0.5f) * r * r * r;
            const long long o = (long long)b * kStemTot + ch;
            const long long plane = (long long)B * kStemTot;
            K[o] = inv * wt;
            K[plane + o] = -inv * inv * sum2 / cntf + 2.f * dv / (hw - 1.f);
            K[2 * plane + o] = -inv * sum1 / cntf + inv * inv * sum2 * mean / cntf + dm / hw - 2.f * dv * (float)im / (hw - 1.f);
        }
    }
}

// one thread per stem element: dstems = K0 * dout + K1 * x + K2
__global__ void __launch_bounds__(kThreads) head_dx_kernel(const HeadBwdParams p, const float* __restrict__ draw, const float* __restrict__ K) {
    const long long total = (long long)p.B * p.HW * kStemTot;
    const long long plane = (long long)p.B * kStemTot;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % kStemTot);
        const long long q = i / kStemTot;
        const int b = (int)(q / p.HW), s = ch / kStemC, c = ch % kStemC;
        const long long o = (long long)b * kStemTot + ch;
        const float x = p.stems[i];
        const float post = fmaf(p.coefA[o], x, p.coefB[o]);
        const float dout = head_dout(draw + q * kNumOut, p.w, c_o0[s], c_o1[s], c, post);
        p.dstems[i] = fmaf(K[o], dout, fmaf(K[plane + o], x, K[2 * plane + o]));
    }
}

inline int grid_for(long long threads_wanted, int cap_blocks) {
    long long b = (threads_wanted + kThreads - 1) / kThreads;
    if (b < 1) b = 1;
    if (b > cap_blocks) b = cap_blocks;
    return (int)b;
}

// blocks for a (channel, slice) kernel: at least ceil(C / 256) so that every channel has a thread, about four waves otherwise
inline int slice_grid(long long P, int C) {
    const long long min_blocks = (C + kThreads - 1) / kThreads;
    long long want = ((long long)C * (P < 1 ? 1 : P) + kThreads - 1) / kThreads;      // one pixel per thread at most
    const long long cap = (long long)sm_count() * 4;
    if (want > cap) want = cap;
    if (want < min_blocks) want = min_blocks;
    return (int)want;
}

}  // namespace

void launch_conv_wgrad(const ConvBwdParams& p, cudaStream_t st) {
    if (!p.dw) return;
    int csum = 0;
    for (int s = 0; s < p.nsrc; ++s) csum += p.srcC[s];
    MC_CHECK(p.nsrc >= 1 && p.nsrc <= kMaxSrc && csum == p.Cin, "conv_wgrad: sources do not add up to Cin");
    const bool quad = p.Cout % 4 == 0 && (reinterpret_cast<uintptr_t>(p.dy) & 15) == 0;
    const long long total = (long long)p.k * p.k * p.Cin * (quad ? p.Cout / 4 : p.Cout);
    const int gx = (int)((total + kThreads - 1) / kThreads);
    const int rows = p.B * p.Hout;
    int slices = sm_count() * 8 / (gx < 1 ? 1 : gx);              // fill the machine a few times over, not more
    if (slices < 1) slices = 1;
    if (slices > rows) slices = rows;
    const int rps = (rows + slices - 1) / slices;
    slices = (rows + rps - 1) / rps;
    if (quad) MC_LAUNCH(conv_wgrad4_kernel, dim3(gx, slices), dim3(kThreads), st, p, rps);
    else MC_LAUNCH(conv_wgrad_kernel, dim3(gx, slices), dim3(kThreads), st, p, rps);
}

void launch_conv_dgrad(const ConvBwdParams& p, cudaStream_t st) {
    bool any = false;
    int csum = 0;
    for (int s = 0; s < p.nsrc; ++s) { any = any || p.dsrc[s]; csum += p.srcC[s]; }
    MC_CHECK(p.nsrc >= 1 && p.nsrc <= kMaxSrc && csum == p.Cin, "conv_dgrad: sources do not add up to Cin");
    if (!any) return;
    if (p.wT) {
        const long long wn = (long long)p.k * p.k * p.Cin * p.Cout;
        MC_LAUNCH(conv_wT_kernel, dim3(grid_for(wn, sm_count() * 8)), dim3(kThreads), st, p.w, p.wT, p.k * p.k, p.Cin, p.Cout);
    }
    bool quad = p.wT != nullptr && p.Cin % 4 == 0 && (reinterpret_cast<uintptr_t>(p.wT) & 15) == 0;
    for (int s = 0; s < p.nsrc; ++s) quad = quad && p.srcC[s] % 4 == 0 && (reinterpret_cast<uintptr_t>(p.dsrc[s]) & 15) == 0;
    const long long total = (long long)p.B * p.Hin * p.Win * (quad ? p.Cin / 4 : p.Cin);
    if (quad) MC_LAUNCH(conv_dgrad4_kernel, dim3((unsigned)((total + kThreads - 1) / kThreads)), dim3(kThreads), st, p);
    else MC_LAUNCH(conv_dgrad_kernel, dim3((unsigned)((total + kThreads - 1) / kThreads)), dim3(kThreads), st, p);
}

void launch_bn_backward(const BnBwdParams& p, cudaStream_t st) {
    MC_CHECK(p.C >= 1 && p.P >= 1 && (!p.relu || p.y), "bn_backward: arguments");
    zero_async(p.sums, sizeof(double) * 2 * p.C, st);
    MC_LAUNCH(bn_bwd_reduce_kernel, dim3(slice_grid(p.P, p.C)), dim3(kThreads), st, p);
    MC_LAUNCH(bn_bwd_apply_kernel, dim3(grid_for(p.P * p.C, sm_count() * 8)), dim3(kThreads), st, p);
}

void launch_colsum(const float* x, long long P, int C, double* sums, float* out, cudaStream_t st) {
    zero_async(sums, sizeof(double) * C, st);
    MC_LAUNCH(colsum_kernel, dim3(slice_grid(P, C)), dim3(kThreads), st, x, P, C, sums);
    MC_LAUNCH(narrow_kernel, dim3((C + kThreads - 1) / kThreads), dim3(kThreads), st, (const double*)sums, out, C);
}

void launch_maxpool2_backward(const float* x, const float* dy, float* dx, int B, int C, int Hin, int Win, cudaStream_t st) {
    MC_CHECK(Hin % 2 == 0 && Win % 2 == 0, "maxpool2_backward: even input size");
    const long long total = (long long)B * (Hin / 2) * (Win / 2) * C;
    MC_LAUNCH(maxpool2_bwd_kernel, dim3(grid_for(total, sm_count() * 8)), dim3(kThreads), st, x, dy, dx, B, C, Hin, Win);
}

void launch_upsample2_backward(const float* x, const float* w, const float* dy, float* dx, float* dw, int B, int C, int Hin, int Win,
                               cudaStream_t st) {
    MC_LAUNCH(upsample2_bwd_kernel, dim3(slice_grid((long long)B * Hin * Win, C)), dim3(kThreads), st, x, w, dy, dx, dw, B, C, Hin, Win);
}

size_t head_bwd_scratch_bytes(int B, int HW) {
    return align256(sizeof(float) * (size_t)B * HW * kNumOut) + align256(sizeof(double) * (size_t)B * kStemTot * 2) +
           align256(sizeof(double) * kNumOut) + align256(sizeof(float) * kStemTot * 2) + align256(sizeof(float) * 3 * (size_t)B * kStemTot) +
           align256(sizeof(float) * kNumStems * kMixFloats);
}

void launch_head_backward(const HeadBwdParams& p, cudaStream_t st) {
    MC_CHECK(p.B >= 2 && p.B <= kMaxTrainB && p.HW >= 2, "head_backward: 2 <= B <= 64");
    const HeadScratch sc = carve(p.scratch, p.B, p.HW);
    const long long Q = (long long)p.B * p.HW;
    MC_LAUNCH(head_draw_kernel, dim3((unsigned)((Q + kThreads - 1) / kThreads)), dim3(kThreads), st, p, sc.draw);
    launch_colsum(sc.draw, Q, kNumOut, sc.colsums, p.dbias, st);
    MC_LAUNCH(head_meaninv_kernel, dim3((kStemTot + 63) / 64), dim3(64), st, p.sums, p.B, p.HW, sc.meaninv);
    zero_async(sc.S, sizeof(double) * (size_t)p.B * kStemTot * 2, st);
    zero_async(p.dw, sizeof(float) * kNumOut * kStemC, st);
    int gx = slice_grid(p.HW, kStemTot);
    const int cap = sm_count() * 4 / p.B > 3 ? sm_count() * 4 / p.B : 3;       // >= 3 blocks: 576 channels need 576 threads
    if (gx > cap) gx = cap;
    MC_LAUNCH(head_reduce_kernel, dim3(gx, p.B), dim3(kThreads), st, p, (const float*)sc.draw, (const float*)sc.meaninv, sc.S);
    MC_LAUNCH(head_mix_bwd_kernel, dim3(1), dim3(kNumStems), st, p, (const double*)sc.S, (const float*)sc.meaninv, sc.K, sc.mix);
    MC_LAUNCH(head_dx_kernel, dim3(grid_for(Q * kStemTot, sm_count() * 16)), dim3(kThreads), st, p, (const float*)sc.draw, (const float*)sc.K);
}

}  // namespace mc

// ---------------------------------------------------------------------------------------------
// per-kernel C entry points (include/monocon_b200.h, "training step, backward kernels").  Same signatures in the product
// library (device pointers, a CUDA stream) and in the host-shim test build (host pointers, stream ignored).
// ---------------------------------------------------------------------------------------------
namespace {
thread_local std::string g_bw_error;      // per thread, like mc_eval_last_error
template <class F> int bw_guard(F&& f) {
    try {
        f();
        return 0;
    } catch (const std::exception& e) {
        g_bw_error = e.what();
        return 1;
    }
}
}  // namespace

extern "C" {

const char* mc_bw_last_error() { return g_bw_error.c_str(); }

int mc_bw_conv(int nsrc, const float* const* src, float* const* dsrc, const int* srcC, const int* srcWp, const int* srcXoff, int B, int Hin,
               int Win, int Hout, int Wout, int Cout, int k, int stride, int pad, const float* w, const float* dy, float* dw, float* wT_scratch,
               void* stream) {
    return bw_guard([&]() {
        MC_CHECK(nsrc >= 1 && nsrc <= mc::kMaxSrc, "mc_bw_conv: 1..4 sources");
        mc::ConvBwdParams p;
        std::memset(&p, 0, sizeof(p));
        p.nsrc = nsrc;
        for (int s = 0; s < nsrc; ++s) {
            p.src[s] = src[s]; p.dsrc[s] = dsrc ? dsrc[s] : nullptr; p.srcC[s] = srcC[s];
            p.srcWp[s] = srcWp ? srcWp[s] : Win; p.srcXoff[s] = srcXoff ? srcXoff[s] : 0;
            p.Cin += srcC[s];
        }
        p.B = B; p.Hin = Hin; p.Win = Win; p.Hout = Hout; p.Wout = Wout; p.Cout = Cout; p.k = k; p.stride = stride; p.pad = pad;
        p.w = w; p.dy = dy; p.dw = dw; p.wT = wT_scratch;
        mc::launch_conv_wgrad(p, (cudaStream_t)stream);
        mc::launch_conv_dgrad(p, (cudaStream_t)stream);
    });
}

int mc_bw_batchnorm(const float* dy, const float* y, const float* raw, const float* mean, const float* inv, const float* gamma, long long P,
                    int C, int relu, double* sums, float* draw, float* dres, float* dgamma, float* dbeta, void* stream) {
    return bw_guard([&]() {
        mc::BnBwdParams p;
        p.dy = dy; p.y = y; p.raw = raw; p.mean = mean; p.inv = inv; p.gamma = gamma; p.P = P; p.C = C; p.relu = relu; p.sums = sums;
        p.draw = draw; p.dres = dres; p.dgamma = dgamma; p.dbeta = dbeta;
        mc::launch_bn_backward(p, (cudaStream_t)stream);
    });
}

int mc_bw_colsum(const float* x, long long P, int C, double* sums, float* out, void* stream) {
    return bw_guard([&]() { mc::launch_colsum(x, P, C, sums, out, (cudaStream_t)stream); });
}

int mc_bw_maxpool2(const float* x, const float* dy, float* dx, int B, int C, int Hin, int Win, void* stream) {
    return bw_guard([&]() { mc::launch_maxpool2_backward(x, dy, dx, B, C, Hin, Win, (cudaStream_t)stream); });
}

int mc_bw_upsample2(const float* x, const float* w, const float* dy, float* dx, float* dw, int B, int C, int Hin, int Win, void* stream) {
    return bw_guard([&]() { mc::launch_upsample2_backward(x, w, dy, dx, dw, B, C, Hin, Win, (cudaStream_t)stream); });
}

long long mc_bw_heads_scratch_bytes(int B, int HW) { return (long long)mc::head_bwd_scratch_bytes(B, HW); }

int mc_bw_heads(const float* const* pred, const float* const* dpred, const float* stems, const double* sums, const float* coefA,
                const float* coefB, const float* att_w, const float* att_gamma, const float* att_beta, const float* bank_w, const float* bank_b,
                const float* w, int B, int HW, void* scratch, float* dstems, float* dw, float* dbias, float* datt_w, float* datt_gamma,
                float* datt_beta, float* dbank_w, float* dbank_b, void* stream) {
    return bw_guard([&]() {
        mc::HeadBwdParams p;
        for (int i = 0; i < mc::kNumPred; ++i) { p.pred[i] = pred[i]; p.dpred[i] = dpred[i]; }
        p.stems = stems; p.sums = sums; p.coefA = coefA; p.coefB = coefB; p.att_w = att_w; p.att_gamma = att_gamma; p.att_beta = att_beta;
        p.bank_w = bank_w; p.bank_b = bank_b; p.w = w; p.B = B; p.HW = HW; p.scratch = scratch; p.dstems = dstems; p.dw = dw; p.dbias = dbias;
        p.datt_w = datt_w; p.datt_gamma = datt_gamma; p.datt_beta = datt_beta; p.dbank_w = dbank_w; p.dbank_b = dbank_b;
        mc::launch_head_backward(p, (cudaStream_t)stream);
    });
}

int mc_bw_run_graph(const mc_bw_tensor* T, int n_tensors, const mc_bw_op* ops, int n_ops, int B, void* stream) {
    return mc_bw_run_graph_range(T, n_tensors, ops, n_ops, B, 0, n_ops, 1, stream);
}

int mc_bw_run_graph_range(const mc_bw_tensor* T, int n_tensors, const mc_bw_op* ops, int n_ops, int B, int op_first, int op_last, int zero,
                          void* stream) {
    return bw_guard([&]() {
        cudaStream_t st = (cudaStream_t)stream;
        MC_CHECK(T && ops && n_tensors > 0 && n_ops > 0 && B >= 1, "mc_bw_run_graph: arguments");
        MC_CHECK(0 <= op_first && op_first <= op_last && op_last <= n_ops, "mc_bw_run_graph_range: 0 <= op_first <= op_last <= n_ops");
        auto tensor = [&](int i) -> const mc_bw_tensor& {
            MC_CHECK(i >= 0 && i < n_tensors, "mc_bw_run_graph: tensor index out of range");
            return T[i];
        };
        for (int i = 0; zero && i < n_tensors; ++i)
            if (T[i].g) mc::zero_async(T[i].g, sizeof(float) * (size_t)B * T[i].H * T[i].W * T[i].C, st);
        for (int i = 0; zero && i < n_ops; ++i) {
            const mc_bw_op& op = ops[i];
            if (!op.dw) continue;
            if (op.type == MC_BW_CONV) {
                int cin = 0;
                for (int s = 0; s < op.nsrc; ++s) cin += tensor(op.src[s]).C;
                mc::zero_async(op.dw, sizeof(float) * (size_t)op.k * op.k * cin * op.cout, st);
            } else if (op.type == MC_BW_UP) {
                mc::zero_async(op.dw, sizeof(float) * (size_t)tensor(op.src[0]).C * 16, st);
            }
        }
        for (int i = op_last - 1; i >= op_first; --i) {
            const mc_bw_op& op = ops[i];
            if (op.type == MC_BW_HEADS) {
                MC_CHECK(op.heads, "mc_bw_run_graph: HEADS without arguments");
                const mc_bw_tensor& stems = tensor(op.src[0]);
                MC_CHECK(stems.C == mc::kStemTot && stems.g, "mc_bw_run_graph: the stems tensor");
                const mc_bw_heads_args& a = *op.heads;
                mc::HeadBwdParams p;
                for (int k = 0; k < mc::kNumPred; ++k) { p.pred[k] = a.pred[k]; p.dpred[k] = a.dpred[k]; }
                p.stems = stems.x; p.sums = a.sums; p.coefA = a.coefA; p.coefB = a.coefB; p.att_w = a.att_w; p.att_gamma = a.att_gamma;
                p.att_beta = a.att_beta; p.bank_w = a.bank_w; p.bank_b = a.bank_b; p.w = a.w; p.B = B; p.HW = stems.H * stems.W;
                p.scratch = a.scratch; p.dstems = stems.g; p.dw = a.dw; p.dbias = a.dbias; p.datt_w = a.datt_w;
                p.datt_gamma = a.datt_gamma; p.datt_beta = a.datt_beta; p.dbank_w = a.dbank_w; p.dbank_b = a.dbank_b;
                mc::launch_head_backward(p, st);            // "=": the stems have one consumer
            } else if (op.type == MC_BW_POOL) {
                const mc_bw_tensor &s = tensor(op.src[0]), &d = tensor(op.dst);
                MC_CHECK(s.g && d.g && s.Wp == s.W && s.xoff == 0, "mc_bw_run_graph: pool tensors");
                mc::launch_maxpool2_backward(s.x, d.g, s.g, B, s.C, s.H, s.W, st);
            } else if (op.type == MC_BW_UP) {
                const mc_bw_tensor &s = tensor(op.src[0]), &d = tensor(op.dst);
                MC_CHECK(s.g && d.g && op.dw && s.Wp == s.W && s.xoff == 0, "mc_bw_run_graph: upsample tensors");
                mc::launch_upsample2_backward(s.x, op.w, d.g, s.g, op.dw, B, s.C, s.H, s.W, st);
            } else {
                MC_CHECK(op.type == MC_BW_CONV && op.nsrc >= 1 && op.nsrc <= mc::kMaxSrc, "mc_bw_run_graph: op type");
                const mc_bw_tensor& d = tensor(op.dst);
                MC_CHECK(d.g && d.C == op.cout, "mc_bw_run_graph: convolution output");
                const float* dy = d.g;
                const long long P = (long long)B * d.H * d.W;
                if (op.has_bn) {
                    mc::BnBwdParams b;
                    b.dy = d.g; b.y = d.x; b.raw = op.raw; b.mean = op.mean; b.inv = op.inv; b.gamma = op.gamma; b.P = P; b.C = op.cout;
                    b.relu = op.relu; b.sums = op.sums; b.draw = op.draw;
                    b.dres = op.residual >= 0 ? tensor(op.residual).g : nullptr;
                    b.dgamma = op.dgamma; b.dbeta = op.dbeta;
                    MC_CHECK(op.raw && op.mean && op.inv && op.draw && op.sums, "mc_bw_run_graph: BatchNorm buffers");
                    mc::launch_bn_backward(b, st);
                    dy = op.draw;
                } else {
                    MC_CHECK(!op.relu && op.residual < 0, "mc_bw_run_graph: a convolution without BatchNorm has a plain epilogue in this network");
                    if (op.dbias) mc::launch_colsum(d.g, P, op.cout, op.sums, op.dbias, st);
                }
                mc::ConvBwdParams c;
                std::memset(&c, 0, sizeof(c));
                c.nsrc = op.nsrc;
                const mc_bw_tensor& s0 = tensor(op.src[0]);
                for (int s = 0; s < op.nsrc; ++s) {
                    const mc_bw_tensor& t = tensor(op.src[s]);
                    MC_CHECK(t.H == s0.H && t.W == s0.W, "mc_bw_run_graph: concatenated sources differ in size");
                    c.src[s] = t.x; c.dsrc[s] = t.g; c.srcC[s] = t.C; c.srcWp[s] = t.Wp; c.srcXoff[s] = t.xoff;
                    c.Cin += t.C;
                }
                c.B = B; c.Hin = s0.H; c.Win = s0.W; c.Hout = d.H; c.Wout = d.W; c.Cout = op.cout; c.k = op.k; c.stride = op.stride; c.pad = op.pad;
                c.w = op.w; c.dy = dy; c.dw = op.dw; c.wT = op.wT;
                mc::launch_conv_wgrad(c, st);
                mc::launch_conv_dgrad(c, st);
            }
        }
    });
}

}  // extern "C"
