// Modulated deformable convolution (DCNv2), 3x3 / stride 1 / pad 1 / dilation 1 / one offset group: the DCN variant of the
// IDAUp proj / node blocks that BASELINE.json's north_star names.  The reference repository ships the plain-convolution
// neck only (model/backbone/dla_neck.py:11-38); the operator follows the published algorithm
// of torchvision.ops.deform_conv2d (torchvision 0.26: csrc/ops/cpu/deform_conv2d_kernel.cpp, bilinear_interpolate +
// deformable_im2col), which oracle/dcn_oracle.py restates and tests/golden/dcn.npz pins.
//
// Built here as   columns  =  mask * bilinear(x, p + tap + offset)      (this file: one bandwidth kernel)
//                 y        =  columns (pixels x 9 Cin)  @  W (9 Cin x Cout)   (the tcgen05 1x1 convolution kernels)
// so the contraction runs on the tensor cores in every precision mode (bf16, fp16 hi + lo planes, FFMA twin) through the
// parity-tested kernels, and the 27-channel offset / mask field comes from an ordinary 3x3 convolution of the same plan.
//
// Column kernel: one warp per output pixel (eight pixels of one image row per block).  Lanes 0..8 turn the pixel's nine (dy, dx, mask) triples into clamped corner
// coordinates and four corner weights (zero for corners outside the image: torchvision substitutes 0 for those samples);
// the 9 x Cin / 8 work items (tap, eight channels) are then spread over the lanes, each item four 16-byte corner loads per
// stored plane and one 16-byte store per plane; the lanes of one tap read consecutive channels of the same four pixels.
// HBM roofline: algorithmic bytes per pixel = (Cin + 32 + 9 Cin) x element size (the four corner reads of a tap hit L1 / L2).
#include <cuda_fp16.h>

#include <type_traits>

#include "engine.h"

namespace mc {

namespace {

template <typename T> struct Dio;
template <> struct Dio<float> {
    static __device__ __forceinline__ float ld1(const float* p, long long) { return __ldg(p); }
    static __device__ __forceinline__ void ld8(const float* p, long long, float (&v)[8]) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    static __device__ __forceinline__ void st8(float* p, long long, const float (&v)[8]) {
        reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
};
template <> struct Dio<bf16> {
    static __device__ __forceinline__ float ld1(const bf16* p, long long) { return __bfloat162float(*p); }
    static __device__ __forceinline__ void ld8(const bf16* p, long long, float (&v)[8]) {
        const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
            v[2 * j] = f.x; v[2 * j + 1] = f.y;
        }
    }
    static __device__ __forceinline__ void st8(bf16* p, long long, const float (&v)[8]) {
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
            o[j] = *reinterpret_cast<uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(p) = make_uint4(o[0], o[1], o[2], o[3]);
    }
};
// DT_SPLIT (common.cuh): the pointer addresses the hi plane, the lo plane sits `plane` elements further; hi + lo is exact in fp32
template <> struct Dio<__half> {
    static __device__ __forceinline__ float ld1(const __half* p, long long plane) { return __half2float(p[0]) + __half2float(p[plane]); }
    static __device__ __forceinline__ void ld8(const __half* p, long long plane, float (&v)[8]) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(p)), b = __ldg(reinterpret_cast<const uint4*>(p + plane));
        const uint32_t wa[4] = {a.x, a.y, a.z, a.w}, wb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 fa = __half22float2(*reinterpret_cast<const __half2*>(&wa[j]));
            const float2 fb = __half22float2(*reinterpret_cast<const __half2*>(&wb[j]));
            v[2 * j] = fa.x + fb.x; v[2 * j + 1] = fa.y + fb.y;
        }
    }
    static __device__ __forceinline__ void st8(__half* p, long long plane, const float (&v)[8]) {
        uint32_t oh[4], ol[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float a = fminf(fmaxf(v[2 * j], -65504.f), 65504.f), b = fminf(fmaxf(v[2 * j + 1], -65504.f), 65504.f);
            const __half2 h = __floats2half2_rn(a, b);
            const float2 hf = __half22float2(h);
            const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
            oh[j] = *reinterpret_cast<const uint32_t*>(&h);
            ol[j] = *reinterpret_cast<const uint32_t*>(&l);
        }
        *reinterpret_cast<uint4*>(p) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
        *reinterpret_cast<uint4*>(p + plane) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
    }
};

// grid: x = blocks of 8 pixels along a row (one warp per pixel), y = image row (b * H + y): no index divisions; element offsets
// are 32-bit (the launcher checks every tensor holds < 2^31 elements).  ncu on the first version (64-bit pixel index split
// with divisions, 64-bit corner addresses): 80 % issue-slot utilisation at 28 % of DRAM bandwidth -- the kernel was bound by
// its own index arithmetic (1980 warp instructions per pixel for 4.5 work items of 8 channels).
template <typename T>
__global__ void __launch_bounds__(256) dcn_columns_kernel(const DcnColParams p) {
    const int lane = threadIdx.x & 31;
    const int x = (int)blockIdx.x * 8 + (int)(threadIdx.x >> 5);
    const int H = p.H, W = p.W;
    const int row = (int)blockIdx.y;                              // b * H + y
    const int b = row / H, y = row - b * H;
    pdl_sync();
    if (x >= W) return;                                           // warp-uniform
    const unsigned pix = (unsigned)row * (unsigned)W + (unsigned)x;
    constexpr bool kSplit = sizeof(T) == 2 && !std::is_same<T, bf16>::value;
    // ---- lanes 0..8: the tap's sampling position -> clamped corners + corner weights (deform_conv2d_kernel.cpp, bilinear_interpolate)
    int oa = 0, ob = 0, xa = 0, xb = 0;                           // row offsets (pixels) of the two corner rows, the two corner columns
    float c1 = 0.f, c2 = 0.f, c3 = 0.f, c4 = 0.f, m = 0.f;
    if (lane < 9) {
        const T* o = reinterpret_cast<const T*>(p.off) + (size_t)pix * p.offC;
        const float oinv = (kSplit && p.off_sc) ? p.off_sc->inv : 1.f;
        const float dy = Dio<T>::ld1(o + 2 * lane, p.off_plane) * oinv, dx = Dio<T>::ld1(o + 2 * lane + 1, p.off_plane) * oinv;
        const float mv = Dio<T>::ld1(o + 18 + lane, p.off_plane) * oinv;
        m = p.mask_logits ? 1.f / (1.f + expf(-mv)) : mv;
        const int ti = lane / 3, tj = lane - 3 * ti;
        const float py = (float)(y - 1 + ti) + dy, px = (float)(x - 1 + tj) + dx;
        int h0 = 0, w0 = 0;
        if (py > -1.f && py < (float)H && px > -1.f && px < (float)W) {
            const float fl_h = floorf(py), fl_w = floorf(px);
            h0 = (int)fl_h; w0 = (int)fl_w;
            const float lh = py - fl_h, lw = px - fl_w, hh = 1.f - lh, hw = 1.f - lw;
            const bool t_ok = h0 >= 0, b_ok = h0 + 1 <= H - 1, l_ok = w0 >= 0, r_ok = w0 + 1 <= W - 1;
            c1 = (t_ok && l_ok) ? hh * hw : 0.f;
            c2 = (t_ok && r_ok) ? hh * lw : 0.f;
            c3 = (b_ok && l_ok) ? lh * hw : 0.f;
            c4 = (b_ok && r_ok) ? lh * lw : 0.f;
        }
        // corners outside the image carry weight 0 (the operator substitutes 0 for them); their addresses are clamped into it
        oa = (b * H + min(max(h0, 0), H - 1)) * W;
        ob = (b * H + min(max(h0 + 1, 0), H - 1)) * W;
        xa = min(max(w0, 0), W - 1);
        xb = min(max(w0 + 1, 0), W - 1);
    }
    const int G = p.Cin >> 3, items = 9 * G;
    const int G0 = p.srcC[0] >> 3;
    const float rescale = kSplit ? (p.src_sc[0] ? p.src_sc[0]->inv : 1.f) * (p.col_sc ? p.col_sc->mul : 1.f) : 1.f;
    T* colp = reinterpret_cast<T*>(p.col) + (size_t)pix * (unsigned)(9 * p.Cin);
    const T* sp0 = reinterpret_cast<const T*>(p.src[0]);
    const T* sp1 = reinterpret_cast<const T*>(p.src[p.nsrc > 1 ? 1 : 0]);
    const int C0 = p.srcC[0], C1 = p.srcC[p.nsrc > 1 ? 1 : 0];
    const long long pl0 = p.src_plane[0], pl1 = p.src_plane[p.nsrc > 1 ? 1 : 0];
    float amax = 0.f;
    for (int base = 0; base < items; base += 32) {
        const int it = base + lane;
        const bool act = it < items;
        const int k = act ? (int)(((unsigned)it * p.g_magic) >> 20) : 0;       // it / G
        const int g = it - k * G;
        const int ta = __shfl_sync(0xffffffffu, oa, k), tb = __shfl_sync(0xffffffffu, ob, k);
        const int ua = __shfl_sync(0xffffffffu, xa, k), ub = __shfl_sync(0xffffffffu, xb, k);
        const float a1 = __shfl_sync(0xffffffffu, c1, k), a2 = __shfl_sync(0xffffffffu, c2, k);
        const float a3 = __shfl_sync(0xffffffffu, c3, k), a4 = __shfl_sync(0xffffffffu, c4, k);
        const float mk = __shfl_sync(0xffffffffu, m, k);
        if (!act) continue;
        const bool second = g >= G0;
        const T* sp = second ? sp1 : sp0;
        const int C = second ? C1 : C0, c = (second ? g - G0 : g) << 3;
        const long long pl = second ? pl1 : pl0;
        float v1[8], v2[8], v3[8], v4[8], o[8];
        Dio<T>::ld8(sp + (unsigned)((ta + ua) * C + c), pl, v1);
        Dio<T>::ld8(sp + (unsigned)((ta + ub) * C + c), pl, v2);
        Dio<T>::ld8(sp + (unsigned)((tb + ua) * C + c), pl, v3);
        Dio<T>::ld8(sp + (unsigned)((tb + ub) * C + c), pl, v4);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float val = a1 * v1[j] + a2 * v2[j] + a3 * v3[j] + a4 * v4[j];     // the operator's summation order
            o[j] = mk * val;
            if (kSplit) { o[j] *= rescale; amax = fmaxf(amax, fabsf(o[j])); }
        }
        Dio<T>::st8(colp + (unsigned)(k * p.Cin + (g << 3)), p.col_plane, o);
    }
    if (kSplit && p.col_amax != nullptr) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, d));
        if (lane == 0 && amax > 0.f) atomicMax(p.col_amax, __float_as_uint(amax));
    }
}

}  // namespace

void launch_dcn_columns(const DcnColParams& p_in, DType dt, cudaStream_t st) {
    DcnColParams p = p_in;
    MC_CHECK(p.nsrc >= 1 && p.nsrc <= 2, "dcn columns: one or two concatenated sources");
    int c = 0;
    for (int s = 0; s < p.nsrc; ++s) { MC_CHECK(p.srcC[s] % 8 == 0, "dcn columns: source channels must be multiples of 8"); c += p.srcC[s]; }
    MC_CHECK(c == p.Cin && p.offC >= 27, "dcn columns: channel counts");
    const long long npix = (long long)p.B * p.H * p.W;
    MC_CHECK(npix * 9 * p.Cin < (1ll << 31) && npix * p.offC < (1ll << 31) && (long long)p.B * p.H < 65536, "dcn columns: tensor too large for 32-bit element offsets");
    // it / G by multiply + shift for it < 9 * G (checked exhaustively here: G <= 128)
    const int G = p.Cin / 8;
    MC_CHECK(G >= 1 && G <= 128, "dcn columns: at most 1024 input channels");
    p.g_magic = (1u << 20) / (unsigned)G + 1u;
    for (int it = 0; it < 9 * G; ++it) MC_CHECK((int)(((unsigned)it * p.g_magic) >> 20) == it / G, "dcn columns: division constant");
    const dim3 grid((unsigned)((p.W + 7) / 8), (unsigned)(p.B * p.H)), block(256);
    if (dt == DT_F32) launch_k(dcn_columns_kernel<float>, grid, block, 0, st, p);
    else if (dt == DT_BF16) launch_k(dcn_columns_kernel<bf16>, grid, block, 0, st, p);
    else launch_k(dcn_columns_kernel<__half>, grid, block, 0, st, p);
}

}  // namespace mc
