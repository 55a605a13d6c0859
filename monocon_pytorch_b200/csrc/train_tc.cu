// Bandwidth kernels of the bf16 tensor-core training step (BASELINE.json configs[2] / [4]): everything
// between the tcgen05 convolutions of forward / dgrad / wgrad.  Activations, raw convolution outputs and activation
// gradients are bf16 NHWC; statistics are accumulated in fp32 partials -> fp64; parameters and their gradients are fp32.
// One thread owns 8 adjacent channels of a pixel (one 16-byte load / store), so a warp moves 512 contiguous bytes.
//
//   forward :  bn_stats_bf16 -> bn_finalize (train_forward.cu) -> bn_apply_bf16          nn.BatchNorm2d in train(), dla.py:24,30,119,185,233
//   backward:  bn_bwd_reduce_bf16 -> bn_bwd_apply_bf16 (gradient of the raw convolution output written dense, or zero-inserted at
//              input resolution for a stride-2 convolution: its dgrad / wgrad then are stride-1 problems), maxpool2_bwd_bf16
//              (dla.py:176-177,193), upsample2_bwd_bf16 (depthwise ConvTranspose2d k4 s2 p1, dla_neck.py:58-65), f32_to_bf16
//              (gradient of the fp32 head stems), repack_bf16 (fp32 master weights -> the bf16 layouts of the convolution plans)
// Formulas: oracle/backward_oracle.py (pinned to the reference's gradients); the fp32 twins live in train_backward.cu.
#include <algorithm>
#include <cstring>

#include "train_tc.h"

namespace mc {

namespace {

constexpr int kT = 256;

struct alignas(16) V8 { __nv_bfloat162 h[4]; };

__device__ __forceinline__ void load8(const void* p, float (&f)[8]) {
    const V8 v = *reinterpret_cast<const V8*>(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(v.h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ void unpack8(const V8& v, float (&f)[8]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(v.h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ void store8(void* p, const float (&f)[8]) {
    V8 v;
#pragma unroll
    for (int i = 0; i < 4; ++i) v.h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    *reinterpret_cast<V8*>(p) = v;
}
__device__ __forceinline__ void loadf8(const float* p, float (&f)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// ---- per-channel sums over pixels: thread = (channel group, pixel lane) ----
// MODE 0: sums = (sum x, sum x^2);  MODE 1: sums = (sum dz, sum dz * xhat) with dz = dy masked by the ReLU, xhat = (raw - mean) * inv.
// Per-thread partials are fp32 over at most 64 pixels, then folded across the warp's lanes of the same channel group (shuffles) into
// fp64 accumulators in shared memory, and from there with one fp64 atomic per channel and block into `sums`.  No fp64 registers:
// the kernel is bandwidth-bound only if enough blocks are resident (ncu: 150 registers -> 12 % occupancy before).
template <int MODE>
__global__ void __launch_bounds__(kT, MODE == 0 ? 3 : 2) chan_sums_bf16_kernel(const bf16* __restrict__ x, const bf16* __restrict__ y, const bf16* __restrict__ raw,
                                                                             const float* __restrict__ mean, const float* __restrict__ inv, int relu,
                                                                             const float* __restrict__ fscale, const float* __restrict__ fshift,
                                                                             long long P, int C, double* __restrict__ sums) {
    extern __shared__ double sh[];                   // [C][2]
    for (int i = threadIdx.x; i < 2 * C; i += kT) sh[i] = 0.0;
    __syncthreads();
    pdl_sync();                                      // programmatic dependent launch: the prologue overlaps the previous kernel's tail
    const int G = C >> 3, ppb = kT / G;
    const int cg = threadIdx.x % G, pl = threadIdx.x / G;
    const bool fold = (G & (G - 1)) == 0 && G < 32;  // block-uniform; then kT % G == 0 and every thread is active
    if (pl < ppb) {
        float xa_[8], xb_[8], fsc[8], fsh[8];        // xhat = raw * xa_ + xb_
        if (MODE == 1) {
            float mu[8], iv[8];
            loadf8(mean + cg * 8, mu);
            loadf8(inv + cg * 8, iv);
#pragma unroll
            for (int j = 0; j < 8; ++j) { xa_[j] = iv[j]; xb_[j] = -mu[j] * iv[j]; }
            // relu == 2: no residual behind this BatchNorm, so the ReLU mask y > 0 is raw * scale + shift > 0 and y is not read
            if (relu == 2) { loadf8(fscale + cg * 8, fsc); loadf8(fshift + cg * 8, fsh); }
        }
        float fs[8], fq[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { fs[j] = 0.f; fq[j] = 0.f; }
        auto flush = [&]() {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float a = fs[j], b = fq[j];
                if (fold) {
                    for (int off = G; off < 32; off <<= 1) { a += __shfl_xor_sync(0xffffffffu, a, off); b += __shfl_xor_sync(0xffffffffu, b, off); }
                }
                if (!fold || (threadIdx.x & 31) < G) {
                    atomicAdd(&sh[2 * (cg * 8 + j)], (double)a);
                    atomicAdd(&sh[2 * (cg * 8 + j) + 1], (double)b);
                }
                fs[j] = 0.f; fq[j] = 0.f;
            }
        };
        int n = 0;
        constexpr int U = 2;                          // pixels per iteration: all loads are issued before the first use
        const long long stride = (long long)gridDim.x * ppb;
        // block-uniform trip count (the shuffles of flush() need whole warps): lanes beyond P add zeros
        for (long long base = (long long)blockIdx.x * ppb; base < P; base += U * stride) {
            const long long pix0 = base + pl;
            V8 xa[U], xr[U], xy[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long pix = pix0 + u * stride;
                if (pix < P) {
                    const long long o = pix * C + cg * 8;
                    xa[u] = *reinterpret_cast<const V8*>(x + o);
                    if (MODE == 1) {
                        xr[u] = *reinterpret_cast<const V8*>(raw + o);
                        if (relu == 1) xy[u] = *reinterpret_cast<const V8*>(y + o);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (pix0 + u * stride >= P) continue;
                float a[8];
                unpack8(xa[u], a);
                if (MODE == 0) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) { fs[j] += a[j]; fq[j] = fmaf(a[j], a[j], fq[j]); }
                } else {
                    float r[8];
                    unpack8(xr[u], r);
                    if (relu == 1) {
                        float yy[8];
                        unpack8(xy[u], yy);
#pragma unroll
                        for (int j = 0; j < 8; ++j) if (!(yy[j] > 0.f)) a[j] = 0.f;
                    } else if (relu == 2) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) if (!(fmaf(r[j], fsc[j], fsh[j]) > 0.f)) a[j] = 0.f;
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) { fs[j] += a[j]; fq[j] = fmaf(a[j], fmaf(r[j], xa_[j], xb_[j]), fq[j]); }
                }
            }
            if (++n == 32) { flush(); n = 0; }
        }
        flush();
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += kT) atomicAdd(&sums[i], sh[i]);
}

// y = raw * scale + shift (+ residual) (ReLU).  256 % (C / 8) == 0 (the launcher checks), so a thread keeps its channel group over the
// grid-stride loop and its constants stay in registers; U vectors per iteration with all loads first.
struct BnFin {                 // bn_finalize folded into the apply kernel: every thread derives the constants of its 8 channels from the
    const double* sums;        // batch sums (the arithmetic of bn_finalize_kernel, train_forward.cu); block 0 publishes them and updates
    double n;                  // the running statistics.  sums == null: scale / shift are read.
    float eps, momentum;
    const float *gamma, *beta;
    float *rmean, *rvar, *scale_out, *shift_out, *mean_out, *inv_out;
};
__global__ void __launch_bounds__(kT) bn_apply_bf16_kernel(const bf16* __restrict__ raw, bf16* __restrict__ y, const bf16* __restrict__ res,
                                                           unsigned total8, int C, const float* __restrict__ scale, const float* __restrict__ shift,
                                                           int relu, const BnFin fin) {
    pdl_sync();
    const int G = C >> 3;
    const int c0 = (int)(threadIdx.x % G) * 8;
    float sc[8], sf[8];
    if (fin.sums == nullptr) {
        loadf8(scale + c0, sc);
        loadf8(shift + c0, sf);
    } else {
        // one channel per thread through shared memory (the fp64 division and square root are long instruction sequences: eight of
        // them per thread cost more than the finalize launch they replace)
        __shared__ float s_sc[1024], s_sf[1024];
        for (int c = threadIdx.x; c < C; c += kT) {
            const double mean = fin.sums[2 * c] / fin.n;
            double var = fin.sums[2 * c + 1] / fin.n - mean * mean;             // biased: what the normalisation uses
            if (var < 0.0) var = 0.0;
            const float inv = (float)(1.0 / sqrt(var + (double)fin.eps));
            const float g = fin.gamma ? fin.gamma[c] : 1.f, b = fin.beta ? fin.beta[c] : 0.f;
            const float scv = g * inv, sfv = b - (float)mean * g * inv;
            s_sc[c] = scv;
            s_sf[c] = sfv;
            if (blockIdx.x == 0) {
                fin.scale_out[c] = scv;
                fin.shift_out[c] = sfv;
                if (fin.mean_out) { fin.mean_out[c] = (float)mean; fin.inv_out[c] = inv; }
                const double unbiased = fin.n > 1.0 ? var * fin.n / (fin.n - 1.0) : var;
                fin.rmean[c] = (1.f - fin.momentum) * fin.rmean[c] + fin.momentum * (float)mean;
                fin.rvar[c] = (1.f - fin.momentum) * fin.rvar[c] + fin.momentum * (float)unbiased;
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 8; ++j) { sc[j] = s_sc[c0 + j]; sf[j] = s_sf[c0 + j]; }
    }
    constexpr int U = 4;
    const unsigned stride = gridDim.x * kT;
    for (unsigned v0 = blockIdx.x * kT + threadIdx.x; v0 < total8; v0 += U * stride) {
        V8 xa[U], xr[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned v = v0 + u * stride;
            if (v < total8) {
                xa[u] = *reinterpret_cast<const V8*>(raw + (size_t)v * 8);
                if (res) xr[u] = *reinterpret_cast<const V8*>(res + (size_t)v * 8);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned v = v0 + u * stride;
            if (v >= total8) break;
            float a[8];
            unpack8(xa[u], a);
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = fmaf(a[j], sc[j], sf[j]);
            if (res) {
                float r[8];
                unpack8(xr[u], r);
#pragma unroll
                for (int j = 0; j < 8; ++j) a[j] += r[j];
            }
            if (relu) {
#pragma unroll
                for (int j = 0; j < 8; ++j) a[j] = fmaxf(a[j], 0.f);
            }
            store8(y + (size_t)v * 8, a);
        }
    }
}

// draw = gamma * inv * (dz - (sum dz + xhat * sum dz*xhat) / P);  dres (+)= dz;  dgamma = sum dz*xhat;  dbeta = sum dz
struct BnBwdTc {
    const bf16 *dy, *y, *raw;
    const float *mean, *inv, *gamma, *fscale, *fshift;
    const double* sums;
    long long P;
    int C, relu;               // relu: 0 none, 1 mask = y > 0, 2 mask = raw * fscale + fshift > 0 (no residual: y is not read)
    int up, H, W;              // up: draw is the zero-inserted tensor [B][2H][2W][C], this pixel goes to (2y, 2x)
    bf16* draw;
    bf16* dres;                // or null
    int dres_acc;              // 1: +=, 0: =
    float *dgamma, *dbeta;
};
__global__ void __launch_bounds__(kT) bn_bwd_apply_bf16_kernel(const BnBwdTc p) {
    pdl_sync();
    const int G = p.C >> 3;                                  // 256 % G == 0: the channel group of a thread is fixed
    const unsigned total8 = (unsigned)(p.P * G);
    const int c0 = (int)(threadIdx.x % G) * 8;
    // draw = g inv (dz - (s0 + xhat s1) / P), xhat = (raw - mean) inv   ==   ka * dz + kb * raw + kc
    float ka[8], kb[8], kc[8], fsc[8], fsh[8];
    if (p.relu == 2) { loadf8(p.fscale + c0, fsc); loadf8(p.fshift + c0, fsh); }
    {
        const float rn = (float)(1.0 / (double)p.P);
        float mu[8], iv[8], g[8];
        loadf8(p.mean + c0, mu);
        loadf8(p.inv + c0, iv);
        if (p.gamma) loadf8(p.gamma + c0, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float s0 = (float)p.sums[2 * (c0 + j)], s1 = (float)p.sums[2 * (c0 + j) + 1];
            const float gi = (p.gamma ? g[j] : 1.f) * iv[j];
            ka[j] = gi;
            kb[j] = -gi * iv[j] * s1 * rn;
            kc[j] = -gi * rn * (s0 - mu[j] * iv[j] * s1);
            if (blockIdx.x == 0 && threadIdx.x < G) {        // one thread per channel group publishes the parameter gradients
                if (p.dgamma) p.dgamma[c0 + j] = s1;
                if (p.dbeta) p.dbeta[c0 + j] = s0;
            }
        }
    }
    constexpr int U = 2;
    const unsigned stride = gridDim.x * kT;
    for (unsigned v0 = blockIdx.x * kT + threadIdx.x; v0 < total8; v0 += U * stride) {
        V8 xd[U], xr[U], xy[U], xs[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned v = v0 + u * stride;
            if (v < total8) {
                xd[u] = *reinterpret_cast<const V8*>(p.dy + (size_t)v * 8);
                xr[u] = *reinterpret_cast<const V8*>(p.raw + (size_t)v * 8);
                if (p.relu == 1) xy[u] = *reinterpret_cast<const V8*>(p.y + (size_t)v * 8);
                if (p.dres && p.dres_acc) xs[u] = *reinterpret_cast<const V8*>(p.dres + (size_t)v * 8);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned v = v0 + u * stride;
            if (v >= total8) break;
            float dz[8], r[8], o[8];
            unpack8(xd[u], dz);
            unpack8(xr[u], r);
            if (p.relu == 1) {
                float yy[8];
                unpack8(xy[u], yy);
#pragma unroll
                for (int j = 0; j < 8; ++j) if (!(yy[j] > 0.f)) dz[j] = 0.f;
            } else if (p.relu == 2) {
#pragma unroll
                for (int j = 0; j < 8; ++j) if (!(fmaf(r[j], fsc[j], fsh[j]) > 0.f)) dz[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = fmaf(ka[j], dz[j], fmaf(kb[j], r[j], kc[j]));
            size_t dst = (size_t)v * 8;
            if (p.up) {
                const unsigned pix = v / (unsigned)G;
                const unsigned x = pix % (unsigned)p.W, t = pix / (unsigned)p.W;
                const unsigned yy = t % (unsigned)p.H, n = t / (unsigned)p.H;
                dst = (((size_t)n * 2 * p.H + 2 * yy) * (2 * (size_t)p.W) + 2 * x) * p.C + c0;
            }
            store8(p.draw + dst, o);
            if (p.dres) {
                if (p.dres_acc) {
                    float d[8];
                    unpack8(xs[u], d);
#pragma unroll
                    for (int j = 0; j < 8; ++j) dz[j] += d[j];
                }
                store8(p.dres + (size_t)v * 8, dz);
            }
        }
    }
}

// thread = (pooled pixel, channel group): the gradient goes to the first maximum of the 2x2 window (ATen's tie rule)
__global__ void __launch_bounds__(kT) maxpool2_bwd_bf16_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, bf16* __restrict__ dx, int B, int C,
                                                               int Hin, int Win, int acc) {
    pdl_sync();
    const int G = C >> 3, Ho = Hin / 2, Wo = Win / 2;
    const int total = B * Ho * Wo * G;               // 32-bit indices (the launcher checks the input holds < 2^31 elements)
    for (int v = (int)blockIdx.x * kT + (int)threadIdx.x; v < total; v += (int)gridDim.x * kT) {
        const int c0 = (v % G) * 8;
        int t = v / G;
        const int ox = t % Wo;
        t /= Wo;
        const int oy = t % Ho;
        const int n = t / Ho;
        float g[8], w[4][8];
        load8(dy + (size_t)v * 8, g);
        int idx[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            idx[k] = (((n * Hin + 2 * oy + (k >> 1)) * Win) + 2 * ox + (k & 1)) * C + c0;
            load8(x + idx[k], w[k]);
        }
        int best[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int b = 0;
            float bv = w[0][j];
#pragma unroll
            for (int k = 1; k < 4; ++k) if (w[k][j] > bv) { bv = w[k][j]; b = k; }
            best[j] = b;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float o[8];
            if (acc) load8(dx + idx[k], o);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = (acc ? o[j] : 0.f) + (best[j] == k ? g[j] : 0.f);
            store8(dx + idx[k], o);
        }
    }
}

// depthwise ConvTranspose2d k4 s2 p1 backward.  thread = (filter row ky, 8-channel group, pixel lane): it reads the input pixel's
// 8 channels and the four output pixels (2i - 1 + ky, 2j - 1 + kx) as 16-byte vectors, keeps the 4 x 8 weight-gradient partials of
// its filter row in registers and contributes its row's part of dx; the four ky threads of a pixel are adjacent lanes (two
// shuffles).  Block partials of dw meet in shared memory, then one global atomic per (channel, tap) and block.
__global__ void __launch_bounds__(kT) upsample2_bwd_bf16_kernel(const bf16* __restrict__ x, const float* __restrict__ w, const bf16* __restrict__ dy,
                                                                bf16* __restrict__ dx, float* __restrict__ dw, int B, int C, int Hin, int Win, int acc) {
    extern __shared__ float shw[];                   // [C][16]
    for (int i = threadIdx.x; i < C * 16; i += kT) shw[i] = 0.f;
    __syncthreads();
    pdl_sync();
    const int G = C >> 3;                            // 256 % (4 G) == 0 (the launcher checks)
    const int ky = threadIdx.x & 3, cg = (threadIdx.x >> 2) % G, pl = threadIdx.x / (4 * G), ppb = kT / (4 * G);
    const int c0 = cg * 8;
    float wk[4][8], dwk[4][8];
#pragma unroll
    for (int kx = 0; kx < 4; ++kx)
#pragma unroll
        for (int j = 0; j < 8; ++j) { wk[kx][j] = w[(c0 + j) * 16 + ky * 4 + kx]; dwk[kx][j] = 0.f; }
    // 32-bit indices (the launcher checks 4 * P * C < 2^31): no 64-bit divisions in the loop (measured neutral on the step time)
    const int P = B * Hin * Win;
    const int Ho = 2 * Hin, Wo = 2 * Win;
    // block-uniform trip count (the shuffles below need whole warps); a pixel lane beyond P only skips its loads and stores
    for (int base = (int)blockIdx.x * ppb; base < P; base += (int)gridDim.x * ppb) {
        const int pix = base + pl;
        const bool valid = pix < P;
        const int j0 = pix % Win;
        const int t = pix / Win;
        const int i0 = t % Hin;
        const int n = t / Hin;
        float xv[8], a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { a[j] = 0.f; xv[j] = 0.f; }
        if (valid) load8(x + pix * C + c0, xv);
        const int oy = 2 * i0 - 1 + ky;
        if (valid && oy >= 0 && oy < Ho) {
            V8 d[4];
            bool ok[4];
#pragma unroll
            for (int kx = 0; kx < 4; ++kx) {
                const int ox = 2 * j0 - 1 + kx;
                ok[kx] = ox >= 0 && ox < Wo;
                if (ok[kx]) d[kx] = *reinterpret_cast<const V8*>(dy + (((n * Ho + oy) * Wo) + ox) * C + c0);
            }
#pragma unroll
            for (int kx = 0; kx < 4; ++kx) {
                if (!ok[kx]) continue;
                float dv[8];
                unpack8(d[kx], dv);
#pragma unroll
                for (int j = 0; j < 8; ++j) { a[j] = fmaf(dv[j], wk[kx][j], a[j]); dwk[kx][j] = fmaf(xv[j], dv[j], dwk[kx][j]); }
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) { a[j] += __shfl_xor_sync(0xffffffffu, a[j], 1); a[j] += __shfl_xor_sync(0xffffffffu, a[j], 2); }
        if (ky == 0 && valid) {
            if (acc) {
                float old[8];
                load8(dx + pix * C + c0, old);
#pragma unroll
                for (int j = 0; j < 8; ++j) a[j] += old[j];
            }
            store8(dx + pix * C + c0, a);
        }
    }
#pragma unroll
    for (int kx = 0; kx < 4; ++kx)
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(&shw[(c0 + j) * 16 + ky * 4 + kx], dwk[kx][j]);
    __syncthreads();
    for (int i = threadIdx.x; i < C * 16; i += kT) atomicAdd(&dw[i], shw[i]);
}

__global__ void __launch_bounds__(kT) f32_to_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, long long total8) {
    for (long long v = (long long)blockIdx.x * kT + threadIdx.x; v < total8; v += (long long)gridDim.x * kT) {
        float a[8];
        loadf8(in + v * 8, a);
        store8(out + v * 8, a);
    }
}

__global__ void __launch_bounds__(kT) bf16_to_f32_kernel(const bf16* __restrict__ in, float* __restrict__ out, long long total8) {
    for (long long v = (long long)blockIdx.x * kT + threadIdx.x; v < total8; v += (long long)gridDim.x * kT) {
        float a[8];
        load8(in + v * 8, a);
        *reinterpret_cast<float4*>(out + v * 8) = make_float4(a[0], a[1], a[2], a[3]);
        *reinterpret_cast<float4*>(out + v * 8 + 4) = make_float4(a[4], a[5], a[6], a[7]);
    }
}

__global__ void __launch_bounds__(kT) repack_bf16_kernel(const float* __restrict__ master, const int* __restrict__ idx, bf16* __restrict__ out, long long n) {
    for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < n; i += (long long)gridDim.x * kT) {
        const int j = idx[i];
        out[i] = __float2bfloat16(j >= 0 ? master[j] : 0.f);
    }
}

// all plans of an engine in one launch: job j owns the blocks [bstart[j], bstart[j + 1]) (kRepackPerBlock packed elements each), so the
// job is looked up once per block, not once per element
constexpr int kRepackPerBlock = kT * 16;
__global__ void __launch_bounds__(kT) repack_all_bf16_kernel(const RepackJob* __restrict__ jobs, int njobs) {
    __shared__ int s_job;
    if (threadIdx.x == 0) {
        int lo = 0, hi = njobs - 1;                    // last job whose first block <= blockIdx.x
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (jobs[mid].start <= (long long)blockIdx.x) lo = mid; else hi = mid - 1;
        }
        s_job = lo;
    }
    __syncthreads();
    const RepackJob jb = jobs[s_job];
    const long long e0 = ((long long)blockIdx.x - jb.start) * kRepackPerBlock;
    bf16* out = reinterpret_cast<bf16*>(jb.out);
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
        const long long e = e0 + i * kT + threadIdx.x;
        if (e < jb.n) {
            const int j = jb.idx[e];
            out[e] = __float2bfloat16(j >= 0 ? jb.master[j] : 0.f);
        }
    }
}

inline int grid_for(long long items, int per_block, int max_blocks) {
    return (int)std::max<long long>(1, std::min<long long>((items + per_block - 1) / per_block, max_blocks));
}

}  // namespace

void launch_bn_stats_bf16(const void* x, long long P, int C, double* sums, cudaStream_t st, bool zeroed) {
    MC_CHECK(C % 8 == 0 && C <= 1024, "bn_stats_bf16: C must be a multiple of 8 and <= 1024");
    if (!zeroed) MC_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, st));
    const int ppb = kT / (C / 8);
    launch_k(chan_sums_bf16_kernel<0>, dim3(grid_for(P, ppb * 16, 148 * 8)), dim3(kT), sizeof(double) * 2 * C, st, (const bf16*)x, (const bf16*)nullptr,
             (const bf16*)nullptr, (const float*)nullptr, (const float*)nullptr, 0, (const float*)nullptr, (const float*)nullptr, P, C, sums);
}

void launch_bn_apply_bf16(const void* raw, void* y, const void* residual, long long P, int C, const float* scale, const float* shift, bool relu,
                          cudaStream_t st) {
    const long long total8 = P * (C / 8);
    MC_CHECK(C % 8 == 0 && kT % (C / 8) == 0 && total8 < (1LL << 31), "bn_apply_bf16: C / 8 must divide 256");
    BnFin fin;
    std::memset(&fin, 0, sizeof(fin));
    launch_k(bn_apply_bf16_kernel, dim3(grid_for(total8, kT * 4, 148 * 8)), dim3(kT), 0, st, (const bf16*)raw, (bf16*)y, (const bf16*)residual, (unsigned)total8, C,
             scale, shift, relu ? 1 : 0, fin);
}

void launch_bn_finalize_apply_bf16(const void* raw, void* y, const void* residual, long long P, int C, const double* sums, float eps, float momentum,
                                   const float* gamma, const float* beta, float* rmean, float* rvar, float* scale, float* shift, float* mean_out,
                                   float* inv_out, bool relu, cudaStream_t st) {
    const long long total8 = P * (C / 8);
    MC_CHECK(C % 8 == 0 && kT % (C / 8) == 0 && total8 < (1LL << 31), "bn_finalize_apply_bf16: C / 8 must divide 256");
    MC_CHECK((mean_out == nullptr) == (inv_out == nullptr), "bn_finalize_apply_bf16: mean_out and inv_out come together");
    BnFin fin;
    fin.sums = sums; fin.n = (double)P; fin.eps = eps; fin.momentum = momentum; fin.gamma = gamma; fin.beta = beta; fin.rmean = rmean; fin.rvar = rvar;
    fin.scale_out = scale; fin.shift_out = shift; fin.mean_out = mean_out; fin.inv_out = inv_out;
    launch_k(bn_apply_bf16_kernel, dim3(grid_for(total8, kT * 4, 148 * 8)), dim3(kT), 0, st, (const bf16*)raw, (bf16*)y, (const bf16*)residual, (unsigned)total8, C,
             (const float*)nullptr, (const float*)nullptr, relu ? 1 : 0, fin);
}

void launch_bn_backward_bf16(const BnBwdTcParams& q, cudaStream_t st) {
    MC_CHECK(q.C % 8 == 0 && kT % (q.C / 8) == 0 && q.P * (q.C / 8) < (1LL << 31), "bn_backward_bf16: C / 8 must divide 256");
    if (!q.sums_zeroed) MC_CUDA(cudaMemsetAsync(q.sums, 0, sizeof(double) * 2 * q.C, st));
    const int ppb = kT / (q.C / 8);
    // the ReLU mask comes from the raw output when the forward's scale / shift are given and nothing was added before the ReLU
    const int relu = !q.relu ? 0 : ((q.fscale && q.fshift && !q.dres) ? 2 : 1);
    launch_k(chan_sums_bf16_kernel<1>, dim3(grid_for(q.P, ppb * 16, 148 * 8)), dim3(kT), sizeof(double) * 2 * q.C, st, (const bf16*)q.dy, (const bf16*)q.y,
             (const bf16*)q.raw, q.mean, q.inv, relu, q.fscale, q.fshift, q.P, q.C, q.sums);
    BnBwdTc p;
    p.dy = (const bf16*)q.dy; p.y = (const bf16*)q.y; p.raw = (const bf16*)q.raw; p.mean = q.mean; p.inv = q.inv; p.gamma = q.gamma;
    p.fscale = q.fscale; p.fshift = q.fshift;
    p.sums = q.sums; p.P = q.P; p.C = q.C; p.relu = relu; p.up = q.up; p.H = q.H; p.W = q.W; p.draw = (bf16*)q.draw;
    p.dres = (bf16*)q.dres; p.dres_acc = q.dres_acc; p.dgamma = q.dgamma; p.dbeta = q.dbeta;
    const long long total8 = q.P * (q.C / 8);
    launch_k(bn_bwd_apply_bf16_kernel, dim3(grid_for(total8, kT * 4, 148 * 8)), dim3(kT), 0, st, p);
}

void launch_maxpool2_backward_bf16(const void* x, const void* dy, void* dx, int B, int C, int Hin, int Win, bool accumulate, cudaStream_t st) {
    MC_CHECK(C % 8 == 0 && Hin % 2 == 0 && Win % 2 == 0, "maxpool2_backward_bf16: geometry");
    MC_CHECK((long long)B * Hin * Win * C < (1ll << 31), "maxpool2_backward_bf16: tensor too large for 32-bit indices");
    const long long total = (long long)B * (Hin / 2) * (Win / 2) * (C / 8);
    launch_k(maxpool2_bwd_bf16_kernel, dim3(grid_for(total, kT * 2, 148 * 16)), dim3(kT), 0, st, (const bf16*)x, (const bf16*)dy, (bf16*)dx, B, C, Hin, Win,
             accumulate ? 1 : 0);
}

void launch_upsample2_backward_bf16(const void* x, const float* w, const void* dy, void* dx, float* dw, int B, int C, int Hin, int Win, bool accumulate,
                                    cudaStream_t st) {
    MC_CHECK(C % 8 == 0 && kT % (C / 2) == 0, "upsample2_backward_bf16: C / 2 must divide 256");
    const long long P = (long long)B * Hin * Win;
    MC_CHECK(4 * P * C < (1ll << 31), "upsample2_backward_bf16: tensor too large for 32-bit indices");
    const int ppb = kT / (C / 2);
    launch_k(upsample2_bwd_bf16_kernel, dim3(grid_for(P, ppb * 8, 148 * 8)), dim3(kT), sizeof(float) * C * 16, st, (const bf16*)x, w, (const bf16*)dy, (bf16*)dx, dw,
             B, C, Hin, Win, accumulate ? 1 : 0);
}

void launch_f32_to_bf16(const float* in, void* out, long long n, cudaStream_t st) {
    MC_CHECK(n % 8 == 0, "f32_to_bf16: length must be a multiple of 8");
    f32_to_bf16_kernel<<<grid_for(n / 8, kT * 4, 148 * 16), kT, 0, st>>>(in, (bf16*)out, n / 8);
    MC_CUDA(cudaGetLastError());
}

void launch_bf16_to_f32(const void* in, float* out, long long n, cudaStream_t st) {
    MC_CHECK(n % 8 == 0, "bf16_to_f32: length must be a multiple of 8");
    bf16_to_f32_kernel<<<grid_for(n / 8, kT * 4, 148 * 16), kT, 0, st>>>((const bf16*)in, out, n / 8);
    MC_CUDA(cudaGetLastError());
}

long long repack_blocks(long long n) { return (n + kRepackPerBlock - 1) / kRepackPerBlock; }

void launch_repack_all_bf16(const RepackJob* jobs_dev, int njobs, long long total_blocks, cudaStream_t st) {
    repack_all_bf16_kernel<<<(unsigned)total_blocks, kT, 0, st>>>(jobs_dev, njobs);
    MC_CUDA(cudaGetLastError());
}

void launch_repack_bf16(const float* master, const int* idx, void* out, long long n, cudaStream_t st) {
    repack_bf16_kernel<<<grid_for(n, kT * 4, 148 * 8), kT, 0, st>>>(master, idx, (bf16*)out, n);
    MC_CUDA(cudaGetLastError());
}

}  // namespace mc
