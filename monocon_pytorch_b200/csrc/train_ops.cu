// Training-side rows of the hot path as sm_100a kernels behind the C ABI:
//   mc_generate_targets  TargetGenerator.__call__                 utils/target_generator.py:30-138, utils/tensor_ops.py:62-125
//   mc_losses            MonoConDenseHeads._get_losses + losses/*  model/dense_heads/monocon_heads.py:203-310
//   mc_optimizer_step    clip_grad_norm_(35, 2) + AdamW.step      engine/monocon_engine.py:39-53,94-100
// The reference drives these from Python (B x n_obj x 9 tiny device writes, boolean-mask gathers with host syncs, a
// foreach optimiser); here each is one or two launches with no host synchronisation.  All are HBM / latency bound:
// no tensor cores.  Float arithmetic follows the reference's float32 operation order (explicit _rn intrinsics: no FMA
// contraction) so that the integer outputs (indices, bins, masks, Gaussian radii) are bit-identical.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/monocon_b200.h"
#include "common.cuh"

namespace mc {
namespace {

constexpr int kCls = 3, kKpt = 9, kBins = 12, kMaxObj = 64;

// ---------------------------------------------------------------------------------------------
// targets
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// gaussian_radius((h, w), min_overlap = 0.3), utils/tensor_ops.py:76-98: float32 tensor arithmetic, math.sqrt in double
__device__ float gaussian_radius_dev(float h, float w) {
    const double mo = 0.3;
    const float b1 = fadd(h, w);
    const float c1 = fdiv(fmul(fmul(w, h), (float)(1 - mo)), (float)(1 + mo));
    const double sq1 = sqrt((double)fsub(fmul(b1, b1), fmul(4.f, c1)));
    const float r1 = fdiv(fsub(b1, (float)sq1), 2.f);
    const float b2 = fmul(2.f, fadd(h, w));
    const float c2 = fmul(fmul((float)(1 - mo), w), h);
    const double sq2 = sqrt((double)fsub(fmul(b2, b2), fmul(16.f, c2)));
    const float r2 = fdiv(fsub(b2, (float)sq2), 8.f);
    const double a3 = 4 * mo;
    const float b3 = fmul((float)(-2 * mo), fadd(h, w));
    const float c3 = fmul(fmul((float)(mo - 1), w), h);
    const double sq3 = sqrt((double)fsub(fmul(b3, b3), fmul((float)(4 * a3), c3)));
    const float r3 = fdiv(fadd(b3, (float)sq3), (float)(2 * a3));
    return fminf(r1, fminf(r2, r3));
}

__device__ __forceinline__ float remainder_f32(float a, float b) {      // torch.remainder on float32
    float m = fmodf(a, b);
    if (m != 0.f && ((b < 0.f) != (m < 0.f))) m = fadd(m, b);
    return m;
}

struct TargetArgs {
    mc_labels lab;
    mc_targets tgt;
    int B, M, fh, fw;
    float h_ratio, w_ratio;
};

// one CTA per image: compaction of the valid rows, per-object / per-key-point targets, then one warp per Gaussian splat
__global__ void __launch_bounds__(256) targets_kernel(const TargetArgs a) {
    __shared__ int s_rows[kMaxObj];
    __shared__ int s_n;
    __shared__ int s_cx[kMaxObj], s_cy[kMaxObj], s_rad[kMaxObj], s_cls[kMaxObj];
    __shared__ int s_kx[kMaxObj * kKpt], s_ky[kMaxObj * kKpt];     // integer key-point cell, kx = INT_MIN: no splat
    const int b = blockIdx.x, tid = threadIdx.x, M = a.M;
    if (tid == 0) {
        int n = 0;
        for (int r = 0; r < M; ++r)
            if (a.lab.mask[(size_t)b * M + r]) s_rows[n++] = r;
        s_n = n;
    }
    __syncthreads();
    const int n = s_n;
    const size_t HW = (size_t)a.fh * a.fw;
    if (tid < n) {
        const int o = tid, r = s_rows[o];
        const float* bx = a.lab.gt_bboxes + ((size_t)b * M + r) * 4;
        const float ctx = fdiv(fmul(fadd(bx[0], bx[2]), a.w_ratio), 2.f);
        const float cty = fdiv(fmul(fadd(bx[1], bx[3]), a.h_ratio), 2.f);
        const int cxi = (int)ctx, cyi = (int)cty;
        const float fbh = fmul(fsub(bx[3], bx[1]), a.h_ratio);
        const float fbw = fmul(fsub(bx[2], bx[0]), a.w_ratio);
        const int radius = max(0, (int)gaussian_radius_dev(fbh, fbw));
        const int cls = (int)a.lab.gt_labels[(size_t)b * M + r];
        s_cx[o] = cxi; s_cy[o] = cyi; s_rad[o] = radius; s_cls[o] = cls;
        const size_t bo = (size_t)b * M + o;
        a.tgt.indices[bo] = (long long)cyi * a.fw + cxi;
        a.tgt.wh[bo * 2] = fbw; a.tgt.wh[bo * 2 + 1] = fbh;
        a.tgt.offset[bo * 2] = fsub(ctx, (float)cxi); a.tgt.offset[bo * 2 + 1] = fsub(cty, (float)cyi);
        const float* b3 = a.lab.gt_bboxes_3d + ((size_t)b * M + r) * 7;
        a.tgt.dim[bo * 3] = b3[3]; a.tgt.dim[bo * 3 + 1] = b3[4]; a.tgt.dim[bo * 3 + 2] = b3[5];
        a.tgt.depth[bo] = a.lab.depths[(size_t)b * M + r];
        // _convert_angle_to_class (utils/target_generator.py:141-149)
        const double PI = 3.141592653589793;
        const float two_pi = (float)(2 * PI);
        const double apc = 2 * PI / (double)kBins;
        const float angle = remainder_f32(b3[6], two_pi);
        const float shifted = remainder_f32(fadd(angle, (float)(apc / 2)), two_pi);
        const int bin = (int)fdiv(shifted, (float)apc);
        a.tgt.alpha_cls[bo] = (float)bin;
        a.tgt.alpha_offset[bo] = fsub(shifted, (float)(bin * apc + apc / 2));
        a.tgt.mask_target[bo] = 1;
    }
    __syncthreads();
    for (int i = tid; i < n * kKpt; i += blockDim.x) {
        const int o = i / kKpt, k = i % kKpt, r = s_rows[o];
        s_kx[i] = INT_MIN; s_ky[i] = 0;
        if (a.lab.gt_kpts_valid_mask[((size_t)b * M + r) * kKpt + k] < 1) continue;
        const float* kp = a.lab.gt_kpts_2d + ((size_t)b * M + r) * 2 * kKpt + 2 * k;
        const float kx = fmul(kp[0], a.w_ratio), ky = fmul(kp[1], a.h_ratio);
        const int kxi = (int)kx, kyi = (int)ky;
        const size_t bo = (size_t)b * M + o;
        a.tgt.center2kpt_offset[bo * 2 * kKpt + 2 * k] = fsub(kx, (float)s_cx[o]);
        a.tgt.center2kpt_offset[bo * 2 * kKpt + 2 * k + 1] = fsub(ky, (float)s_cy[o]);
        a.tgt.mask_center2kpt_offset[bo * 2 * kKpt + 2 * k] = 1.f;
        a.tgt.mask_center2kpt_offset[bo * 2 * kKpt + 2 * k + 1] = 1.f;
        if (!(kxi >= 0 && kxi < a.fw && kyi >= 0 && kyi < a.fh)) continue;
        s_kx[i] = kxi; s_ky[i] = kyi;
        a.tgt.indices_kpt[bo * kKpt + k] = (long long)kyi * a.fw + kxi;
        a.tgt.kpt_heatmap_offset[bo * 2 * kKpt + 2 * k] = fsub(kx, (float)kxi);
        a.tgt.kpt_heatmap_offset[bo * 2 * kKpt + 2 * k + 1] = fsub(ky, (float)kyi);
        a.tgt.mask_kpt_heatmap_offset[bo * 2 * kKpt + 2 * k] = 1.f;
        a.tgt.mask_kpt_heatmap_offset[bo * 2 * kKpt + 2 * k + 1] = 1.f;
    }
    __syncthreads();
    // splats: job j = o * 10 + w; w = 0: centre heat-map of the object's class, w = 1..9: key-point heat-map w - 1.
    // max() is commutative and the values are >= 0, so concurrent splats use an integer atomicMax on the float bits.
    const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    for (int j = warp; j < n * (1 + kKpt); j += nwarps) {
        const int o = j / (1 + kKpt), w = j % (1 + kKpt);
        int cx, cy;
        float* canvas;
        if (w == 0) {
            cx = s_cx[o]; cy = s_cy[o];
            if (!(cx >= 0 && cx < a.fw && cy >= 0 && cy < a.fh) || s_cls[o] >= kCls) continue;   // the dataset keeps centres inside
            canvas = a.tgt.center_heatmap + ((size_t)b * kCls + s_cls[o]) * HW;
        } else {
            const int i = o * kKpt + (w - 1);
            if (s_kx[i] == INT_MIN) continue;
            cx = s_kx[i]; cy = s_ky[i];
            canvas = a.tgt.kpt_heatmap + ((size_t)b * kKpt + (w - 1)) * HW;
        }
        const int rad = s_rad[o], d = 2 * rad + 1;
        const double sigma = (double)d / 6;
        const float two_s2 = (float)(2 * sigma * sigma);
        for (int t = lane; t < d * d; t += 32) {
            const int dy = t / d - rad, dx = t % d - rad;
            const int px = cx + dx, py = cy + dy;
            if (px < 0 || px >= a.fw || py < 0 || py >= a.fh) continue;
            const float fx = (float)dx, fy = (float)dy;
            float g = expf(fdiv(-fadd(fmul(fx, fx), fmul(fy, fy)), two_s2));
            if (g < 1.1920929e-07f) g = 0.f;                      // h[h < eps * h.max()] = 0, h.max() = 1
            atomicMax(reinterpret_cast<int*>(canvas + (size_t)py * a.fw + px), __float_as_int(g));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// losses
// ---------------------------------------------------------------------------------------------
// workspace (doubles): [0..2] centre heat-map pos / neg / num_pos, [3..5] key-point heat-map, [6] error flag (N == 0)
constexpr int kWsDoubles = 8;

__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// GaussianFocalLoss partial sums (losses/focal_loss.py:21-44) over both heat-maps: elements [0, n0) = centre, [n0, n0 + n1) = key-point
__global__ void __launch_bounds__(256) focal_reduce_kernel(const float* p0, const float* t0, long long n0, const float* p1, const float* t1,
                                                          long long n1, double* ws) {
    double acc[6] = {0, 0, 0, 0, 0, 0};
    const long long total = n0 + n1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const bool second = i >= n0;
        const float p = second ? p1[i - n0] : p0[i];
        const float t = second ? t1[i - n0] : t0[i];
        const int base = second ? 3 : 0;
        if (t == 1.f) {
            const float q = 1.f - p;
            acc[base] += (double)(logf(p + 1e-12f) * (q * q));
            acc[base + 2] += 1.0;
        } else if (t < 1.f) {
            const float u = 1.f - t, u2 = u * u;
            acc[base + 1] += (double)(logf((1.f - p) + 1e-12f) * (p * p) * (u2 * u2));
        }
    }
    __shared__ double s[6][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const double v = warp_sum(acc[k]);
        if (lane == 0) s[k][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double v = 0;
        for (int w = 0; w < 8; ++w) v += s[threadIdx.x][w];
        if (v != 0) atomicAdd(&ws[threadIdx.x], v);
    }
}

// d loss / d p of the focal loss (num_pos from the reduce pass)
__global__ void __launch_bounds__(256) focal_grad_kernel(const float* p0, const float* t0, float* g0, long long n0, const float* p1,
                                                        const float* t1, float* g1, long long n1, const double* ws) {
    const long long total = n0 + n1;
    const float inv0 = ws[2] > 0 ? (float)(1.0 / ws[2]) : 1.f, inv1 = ws[5] > 0 ? (float)(1.0 / ws[5]) : 1.f;
    const bool has0 = ws[2] > 0, has1 = ws[5] > 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const bool second = i >= n0;
        const long long j = second ? i - n0 : i;
        const float p = second ? p1[j] : p0[j];
        const float t = second ? t1[j] : t0[j];
        const bool has = second ? has1 : has0;
        const float inv = second ? inv1 : inv0;
        float g = 0.f;
        if (t == 1.f) {
            // -(d/dp)[log(p + eps) (1 - p)^2]; dropped entirely when num_pos == 0 (only the negative term is kept then)
            const float q = 1.f - p;
            g = has ? -(q * q / (p + 1e-12f) - 2.f * q * logf(p + 1e-12f)) * inv : 0.f;
        } else if (t < 1.f) {
            const float u = 1.f - t, u2 = u * u, q = (1.f - p) + 1e-12f;
            g = -(-(p * p) / q + 2.f * p * logf(q)) * (u2 * u2) * inv;
        }
        (second ? g1 : g0)[j] = g;
    }
}

struct LossArgs {
    const float* pred[MC_NUM_PRED];
    float* grad[MC_NUM_PRED];
    mc_targets tgt;
    int B, M, fh, fw;
    int with_grad;
    double* ws;
    float* losses;        // [10] in the reference's dict order (monocon_heads.py:299-309)
};

__device__ __forceinline__ float sgn(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }

// block-wide sum of `cnt` doubles per thread (cnt <= 8), result valid in every thread
template <int CNT>
__device__ void block_sum(double (&v)[CNT], double* s_red /*[CNT][8]*/) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < CNT; ++k) {
        const double w = warp_sum(v[k]);
        if (lane == 0) s_red[k * 8 + warp] = w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < CNT; ++k) {
        double t = 0;
        for (int w = 0; w < 8; ++w) t += s_red[k * 8 + w];
        v[k] = t;
    }
    __syncthreads();
}

// One CTA: the eight gathered ("sparse") losses over the <= B * max_objs valid objects, their gradients scattered with
// atomicAdd (objects may share a pixel), and the final ten loss values.  Thread t walks objects t, t + 256, ...
__global__ void __launch_bounds__(256) object_losses_kernel(const LossArgs a) {
    __shared__ double s_red[8 * 8];
    const int M = a.M, total = a.B * M;
    const size_t HW = (size_t)a.fh * a.fw;
    const int K2 = 2 * kKpt;
    auto at = [&](int pi, int b, int c, long long idx, int C) -> size_t { return ((size_t)b * C + c) * HW + (size_t)idx; };
    // ---- pass A: counts and the dim-loss compensation weight ----
    double pa[5] = {0, 0, 0, 0, 0};        // N, sum mask_c2k, sum mask_kho, sum |d dim|, sum |d dim| / dim_pred
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        if (!a.tgt.mask_target[i]) continue;
        const int b = i / M;
        const long long idx = a.tgt.indices[i];
        pa[0] += 1;
        for (int c = 0; c < K2; ++c) { pa[1] += a.tgt.mask_center2kpt_offset[(size_t)i * K2 + c]; pa[2] += a.tgt.mask_kpt_heatmap_offset[(size_t)i * K2 + c]; }
        for (int c = 0; c < 3; ++c) {
            const float dp = a.pred[6][at(6, b, c, idx, 3)];
            const float ad = fabsf(dp - a.tgt.dim[(size_t)i * 3 + c]);
            pa[3] += ad;
            pa[4] += ad / dp;
        }
    }
    block_sum<5>(pa, s_red);
    const float N = (float)pa[0];
    const float den_c2k = __fadd_rn((float)pa[1], 1e-12f), den_kho = __fadd_rn((float)pa[2], 1e-12f);
    const float mean_l1 = (float)(pa[3] / (3.0 * pa[0])), mean_l = (float)(pa[4] / (3.0 * pa[0]));
    const float comp = mean_l1 / mean_l;
    // ---- pass B: loss sums + gradient scatter ----
    double pb[8] = {0, 0, 0, 0, 0, 0, 0, 0};    // wh, offset, dim, c2k, kho, alpha_cls, alpha_reg, depth
    const bool wg = a.with_grad != 0;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        if (!a.tgt.mask_target[i]) continue;
        const int b = i / M;
        const long long idx = a.tgt.indices[i];
        for (int c = 0; c < 2; ++c) {
            const float dw = a.pred[2][at(2, b, c, idx, 2)] - a.tgt.wh[(size_t)i * 2 + c];
            pb[0] += fabsf(dw);
            const float dof = a.pred[3][at(3, b, c, idx, 2)] - a.tgt.offset[(size_t)i * 2 + c];
            pb[1] += fabsf(dof);
            if (wg) {
                atomicAdd(&a.grad[2][at(2, b, c, idx, 2)], 0.1f * sgn(dw) / (2.f * N));
                atomicAdd(&a.grad[3][at(3, b, c, idx, 2)], sgn(dof) / (2.f * N));
            }
        }
        for (int c = 0; c < 3; ++c) {
            const float dp = a.pred[6][at(6, b, c, idx, 3)];
            const float dd = dp - a.tgt.dim[(size_t)i * 3 + c];
            pb[2] += (double)(fabsf(dd) / dp * comp);
            if (wg) atomicAdd(&a.grad[6][at(6, b, c, idx, 3)], comp * sgn(dd) / dp / (3.f * N));
        }
        for (int c = 0; c < K2; ++c) {
            const float m = a.tgt.mask_center2kpt_offset[(size_t)i * K2 + c];
            const float v = a.pred[5][at(5, b, c, idx, K2)] * m - a.tgt.center2kpt_offset[(size_t)i * K2 + c];
            pb[3] += fabsf(v);
            if (wg && m != 0.f) atomicAdd(&a.grad[5][at(5, b, c, idx, K2)], m * sgn(v) / den_c2k);
        }
        // kpt_heatmap_offset is gathered at indices_kpt and NOT masked (monocon_heads.py:264-275): key-points without a
        // cell read pixel 0 against a zero target, exactly as the reference does
        for (int k = 0; k < kKpt; ++k) {
            const long long ik = a.tgt.indices_kpt[(size_t)i * kKpt + k];
            for (int c = 0; c < 2; ++c) {
                const float v = a.pred[4][at(4, b, c, ik, 2)] - a.tgt.kpt_heatmap_offset[(size_t)i * K2 + 2 * k + c];
                pb[4] += fabsf(v);
                if (wg) atomicAdd(&a.grad[4][at(4, b, c, ik, 2)], sgn(v) / den_kho);
            }
        }
        const int bin = (int)a.tgt.alpha_cls[i];
        for (int c = 0; c < kBins; ++c) {
            const float x = a.pred[8][at(8, b, c, idx, kBins)];
            const float y = c == bin ? 1.f : 0.f;
            pb[5] += (double)(fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x))));
            if (wg) atomicAdd(&a.grad[8][at(8, b, c, idx, kBins)], (1.f / (1.f + expf(-x)) - y) / (12.f * N));
        }
        if (bin >= 0 && bin < kBins) {
            const float v = a.pred[9][at(9, b, bin, idx, kBins)] - a.tgt.alpha_offset[i];
            pb[6] += fabsf(v);
            if (wg) atomicAdd(&a.grad[9][at(9, b, bin, idx, kBins)], sgn(v) / N);
        } else {
            pb[6] += fabsf(a.tgt.alpha_offset[i]);
        }
        {
            const float d = a.pred[7][at(7, b, 0, idx, 2)], s = a.pred[7][at(7, b, 1, idx, 2)];
            const float dv = d - a.tgt.depth[i], e = 1.4142f * expf(-s);
            pb[7] += (double)(e * fabsf(dv) + s);
            if (wg) {
                atomicAdd(&a.grad[7][at(7, b, 0, idx, 2)], e * sgn(dv) / N);
                atomicAdd(&a.grad[7][at(7, b, 1, idx, 2)], (1.f - e * fabsf(dv)) / N);
            }
        }
    }
    block_sum<8>(pb, s_red);
    if (threadIdx.x == 0) {
        const double* ws = a.ws;
        auto focal = [&](int o) { return ws[o + 2] > 0 ? (float)(-(ws[o] + ws[o + 1]) / ws[o + 2]) : (float)(-ws[o + 1]); };
        const double n = pa[0];
        if (n <= 0) a.ws[6] = 1.0;                 // the reference asserts here (losses/l1_loss.py:15); the host wrapper raises
        a.losses[0] = focal(0);
        a.losses[1] = (float)(0.1 * pb[0] / (2.0 * n));
        a.losses[2] = (float)(pb[1] / (2.0 * n));
        a.losses[3] = (float)(pb[2] / (3.0 * n));
        a.losses[4] = (float)pb[3] / den_c2k;
        a.losses[5] = focal(3);
        a.losses[6] = (float)pb[4] / den_kho;
        a.losses[7] = (float)(pb[5] / (12.0 * n));
        a.losses[8] = (float)(pb[6] / n);
        a.losses[9] = (float)(pb[7] / n);
    }
}

// ---------------------------------------------------------------------------------------------
// clip_grad_norm_ + AdamW
// ---------------------------------------------------------------------------------------------
constexpr int kChunk = 8192;
struct OptChunk { int tensor; int offset; int len; };

__global__ void __launch_bounds__(256) grad_sumsq_kernel(const OptChunk* chunks, int nchunks, float* const* grads, double* sumsq) {
    for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const OptChunk ch = chunks[c];
        const float* g = grads[ch.tensor];
        if (!g) continue;
        double acc = 0;
        for (int i = threadIdx.x; i < ch.len; i += blockDim.x) { const float v = g[ch.offset + i]; acc += (double)v * v; }
        acc = warp_sum(acc);
        __shared__ double s[8];
        if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0;
            for (int w = 0; w < 8; ++w) t += s[w];
            atomicAdd(sumsq, t);
        }
        __syncthreads();
    }
}

struct AdamArgs {
    const OptChunk* chunks;
    int nchunks;
    float* const* params;
    float* const* grads;
    float* const* m;
    float* const* v;
    const double* sumsq;
    float* total_norm;            // optional output (device)
    float max_norm, decay, one_minus_b1, beta2, one_minus_b2, bc2_sqrt, eps, neg_step_size;
};

// torch.optim.AdamW single-tensor update (mul_(1 - lr wd), lerp_, mul_ / addcmul_, sqrt / sqrt(bc2) + eps, addcdiv_) on the
// clipped gradient; tensors without a gradient are skipped entirely, as torch does
__global__ void __launch_bounds__(256) adamw_kernel(const AdamArgs a) {
    const float total = (float)sqrt(*a.sumsq);
    const float coef = fminf(1.f, a.max_norm / (total + 1e-6f));
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.total_norm) *a.total_norm = total;
    for (int c = blockIdx.x; c < a.nchunks; c += gridDim.x) {
        const OptChunk ch = a.chunks[c];
        const float* g = a.grads[ch.tensor];
        if (!g) continue;
        float* p = a.params[ch.tensor] + ch.offset;
        float* m = a.m[ch.tensor] + ch.offset;
        float* v = a.v[ch.tensor] + ch.offset;
        g += ch.offset;
        for (int i = threadIdx.x; i < ch.len; i += blockDim.x) {
            const float gi = __fmul_rn(g[i], coef);
            float pi = __fmul_rn(p[i], a.decay);
            const float mi = __fadd_rn(m[i], __fmul_rn(a.one_minus_b1, __fsub_rn(gi, m[i])));
            const float vi = __fadd_rn(__fmul_rn(v[i], a.beta2), __fmul_rn(__fmul_rn(a.one_minus_b2, gi), gi));
            const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vi), a.bc2_sqrt), a.eps);
            pi = __fadd_rn(pi, __fmul_rn(a.neg_step_size, __fdiv_rn(mi, denom)));
            p[i] = pi; m[i] = mi; v[i] = vi;
        }
    }
}

thread_local std::string g_train_error;

template <typename F>
int guarded_train(int device, F&& f) {
    try {
        MC_CUDA(cudaSetDevice(device));
        f();
        return 0;
    } catch (const std::exception& e) {
        g_train_error = e.what();
        return 1;
    }
}

}  // namespace
}  // namespace mc

using namespace mc;

struct mc_optimizer {
    int device = 0, n = 0, nchunks = 0;
    OptChunk* d_chunks = nullptr;
    float **d_params = nullptr, **d_m = nullptr, **d_v = nullptr, **d_grads = nullptr;
    double* d_sumsq = nullptr;
    std::vector<float*> h_grads;
};

extern "C" {

const char* mc_train_last_error(void) { return g_train_error.c_str(); }

int mc_generate_targets(int device, int B, int max_objs, int feat_h, int feat_w, int pad_h, int pad_w, const mc_labels* labels,
                        const mc_targets* targets, void* stream) {
    return guarded_train(device, [&]() {
        MC_CHECK(labels && targets, "labels / targets");
        MC_CHECK(B >= 1 && max_objs >= 1 && max_objs <= kMaxObj, "B / max_objs (<= 64)");
        MC_CHECK(feat_h >= 1 && feat_w >= 1 && pad_h >= 1 && pad_w >= 1, "geometry");
        cudaStream_t st = (cudaStream_t)stream;
        const size_t HW = (size_t)feat_h * feat_w, BM = (size_t)B * max_objs;
        const mc_targets& t = *targets;
        // _create_empty_target (utils/target_generator.py:152-177): everything starts at zero
        MC_CUDA(cudaMemsetAsync(t.center_heatmap, 0, sizeof(float) * B * kCls * HW, st));
        MC_CUDA(cudaMemsetAsync(t.kpt_heatmap, 0, sizeof(float) * B * kKpt * HW, st));
        MC_CUDA(cudaMemsetAsync(t.wh, 0, sizeof(float) * BM * 2, st));
        MC_CUDA(cudaMemsetAsync(t.offset, 0, sizeof(float) * BM * 2, st));
        MC_CUDA(cudaMemsetAsync(t.dim, 0, sizeof(float) * BM * 3, st));
        MC_CUDA(cudaMemsetAsync(t.alpha_cls, 0, sizeof(float) * BM, st));
        MC_CUDA(cudaMemsetAsync(t.alpha_offset, 0, sizeof(float) * BM, st));
        MC_CUDA(cudaMemsetAsync(t.depth, 0, sizeof(float) * BM, st));
        MC_CUDA(cudaMemsetAsync(t.center2kpt_offset, 0, sizeof(float) * BM * 2 * kKpt, st));
        MC_CUDA(cudaMemsetAsync(t.kpt_heatmap_offset, 0, sizeof(float) * BM * 2 * kKpt, st));
        MC_CUDA(cudaMemsetAsync(t.indices, 0, sizeof(int64_t) * BM, st));
        MC_CUDA(cudaMemsetAsync(t.indices_kpt, 0, sizeof(int64_t) * BM * kKpt, st));
        MC_CUDA(cudaMemsetAsync(t.mask_target, 0, BM, st));
        MC_CUDA(cudaMemsetAsync(t.mask_center2kpt_offset, 0, sizeof(float) * BM * 2 * kKpt, st));
        MC_CUDA(cudaMemsetAsync(t.mask_kpt_heatmap_offset, 0, sizeof(float) * BM * 2 * kKpt, st));
        TargetArgs a;
        a.lab = *labels; a.tgt = t;
        a.B = B; a.M = max_objs; a.fh = feat_h; a.fw = feat_w;
        a.h_ratio = (float)((double)feat_h / (double)pad_h);
        a.w_ratio = (float)((double)feat_w / (double)pad_w);
        targets_kernel<<<B, 256, 0, st>>>(a);
        MC_CUDA(cudaGetLastError());
    });
}

size_t mc_losses_workspace_bytes(void) { return sizeof(double) * kWsDoubles; }

int mc_losses(int device, int B, int max_objs, int feat_h, int feat_w, const float* const pred[MC_NUM_PRED], const mc_targets* targets,
              float* losses_out, float* const grad[MC_NUM_PRED], void* workspace, void* stream) {
    return guarded_train(device, [&]() {
        MC_CHECK(pred && targets && losses_out && workspace, "arguments");
        MC_CHECK(B >= 1 && max_objs >= 1 && max_objs <= kMaxObj, "B / max_objs (<= 64)");
        static const int ch[MC_NUM_PRED] = {3, 9, 2, 2, 2, 18, 3, 2, 12, 12};
        cudaStream_t st = (cudaStream_t)stream;
        const size_t HW = (size_t)feat_h * feat_w;
        double* ws = (double*)workspace;
        MC_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * kWsDoubles, st));
        const long long n0 = (long long)B * kCls * HW, n1 = (long long)B * kKpt * HW;
        const int blocks = (int)std::min<long long>((n0 + n1 + 255) / 256, 148 * 8);
        focal_reduce_kernel<<<blocks, 256, 0, st>>>(pred[0], targets->center_heatmap, n0, pred[1], targets->kpt_heatmap, n1, ws);
        MC_CUDA(cudaGetLastError());
        LossArgs a;
        std::memset(&a, 0, sizeof(a));
        for (int i = 0; i < MC_NUM_PRED; ++i) { a.pred[i] = pred[i]; a.grad[i] = grad ? grad[i] : nullptr; }
        a.tgt = *targets;
        a.B = B; a.M = max_objs; a.fh = feat_h; a.fw = feat_w;
        a.with_grad = grad ? 1 : 0;
        a.ws = ws; a.losses = losses_out;
        if (grad) {
            for (int i = 2; i < MC_NUM_PRED; ++i) {
                MC_CHECK(grad[i] != nullptr, "grad[i]");
                MC_CUDA(cudaMemsetAsync(grad[i], 0, sizeof(float) * B * ch[i] * HW, st));
            }
        }
        object_losses_kernel<<<1, 256, 0, st>>>(a);
        MC_CUDA(cudaGetLastError());
        if (grad) {
            MC_CHECK(grad[0] && grad[1], "grad[0..1]");
            focal_grad_kernel<<<blocks, 256, 0, st>>>(pred[0], targets->center_heatmap, grad[0], n0, pred[1], targets->kpt_heatmap, grad[1], n1, ws);
            MC_CUDA(cudaGetLastError());
        }
    });
}

int mc_optimizer_create(mc_optimizer** out, int device, int n_tensors, float* const* params, float* const* exp_avg,
                        float* const* exp_avg_sq, const int64_t* numel) {
    if (!out) return 1;
    *out = nullptr;
    mc_optimizer* o = new mc_optimizer();
    int rc = guarded_train(device, [&]() {
        MC_CHECK(n_tensors >= 1 && params && exp_avg && exp_avg_sq && numel, "arguments");
        o->device = device; o->n = n_tensors;
        std::vector<OptChunk> chunks;
        for (int t = 0; t < n_tensors; ++t) {
            MC_CHECK(numel[t] >= 0 && numel[t] < (1ll << 31), "numel");
            for (int64_t off = 0; off < numel[t]; off += kChunk)
                chunks.push_back(OptChunk{t, (int)off, (int)std::min<int64_t>(kChunk, numel[t] - off)});
        }
        o->nchunks = (int)chunks.size();
        MC_CUDA(cudaMalloc(&o->d_chunks, sizeof(OptChunk) * std::max<size_t>(1, chunks.size())));
        MC_CUDA(cudaMemcpy(o->d_chunks, chunks.data(), sizeof(OptChunk) * chunks.size(), cudaMemcpyHostToDevice));
        const size_t pb = sizeof(float*) * n_tensors;
        MC_CUDA(cudaMalloc(&o->d_params, pb)); MC_CUDA(cudaMalloc(&o->d_m, pb)); MC_CUDA(cudaMalloc(&o->d_v, pb)); MC_CUDA(cudaMalloc(&o->d_grads, pb));
        MC_CUDA(cudaMemcpy(o->d_params, params, pb, cudaMemcpyHostToDevice));
        MC_CUDA(cudaMemcpy(o->d_m, exp_avg, pb, cudaMemcpyHostToDevice));
        MC_CUDA(cudaMemcpy(o->d_v, exp_avg_sq, pb, cudaMemcpyHostToDevice));
        MC_CUDA(cudaMalloc(&o->d_sumsq, sizeof(double)));
        o->h_grads.assign(n_tensors, nullptr);
    });
    if (rc) { delete o; return rc; }
    *out = o;
    return 0;
}

int mc_optimizer_step(mc_optimizer* o, float* const* grads, int step, double lr, double beta1, double beta2, double eps,
                      double weight_decay, double max_norm, float* total_norm_out, void* stream) {
    if (!o) return 1;
    return guarded_train(o->device, [&]() {
        MC_CHECK(grads != nullptr && step >= 1, "grads / step (1-based)");
        cudaStream_t st = (cudaStream_t)stream;
        if (std::memcmp(o->h_grads.data(), grads, sizeof(float*) * o->n) != 0) {          // gradient tensors moved (zero_grad(set_to_none))
            std::memcpy(o->h_grads.data(), grads, sizeof(float*) * o->n);
            MC_CUDA(cudaMemcpyAsync(o->d_grads, o->h_grads.data(), sizeof(float*) * o->n, cudaMemcpyHostToDevice, st));
        }
        MC_CUDA(cudaMemsetAsync(o->d_sumsq, 0, sizeof(double), st));
        const int blocks = std::min(o->nchunks, 148 * 8);
        grad_sumsq_kernel<<<blocks, 256, 0, st>>>(o->d_chunks, o->nchunks, o->d_grads, o->d_sumsq);
        MC_CUDA(cudaGetLastError());
        AdamArgs a;
        a.chunks = o->d_chunks; a.nchunks = o->nchunks;
        a.params = o->d_params; a.grads = o->d_grads; a.m = o->d_m; a.v = o->d_v;
        a.sumsq = o->d_sumsq; a.total_norm = total_norm_out;
        const double bc1 = 1.0 - std::pow(beta1, (double)step), bc2 = 1.0 - std::pow(beta2, (double)step);
        a.max_norm = (float)max_norm;
        a.decay = (float)(1.0 - lr * weight_decay);
        a.one_minus_b1 = (float)(1.0 - beta1);
        a.beta2 = (float)beta2;
        a.one_minus_b2 = (float)(1.0 - beta2);
        a.bc2_sqrt = (float)std::sqrt(bc2);
        a.eps = (float)eps;
        a.neg_step_size = (float)(-(lr / bc1));
        adamw_kernel<<<blocks, 256, 0, st>>>(a);
        MC_CUDA(cudaGetLastError());
    });
}

void mc_optimizer_destroy(mc_optimizer* o) {
    if (!o) return;
    cudaSetDevice(o->device);
    cudaFree(o->d_chunks); cudaFree(o->d_params); cudaFree(o->d_m); cudaFree(o->d_v); cudaFree(o->d_grads); cudaFree(o->d_sumsq);
    delete o;
}

}  // extern "C"
