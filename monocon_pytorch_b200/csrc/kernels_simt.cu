// FFMA (fp32 CUDA-core) kernels: the fp32-accurate convolution path, and the bandwidth-bound
// layout / pooling / up-sampling kernels shared by both precision modes.  sm_100a.
#include <cuda_fp16.h>

#include <cstdlib>

#include <algorithm>

#include "common.cuh"

namespace mc {

// ---------------------------------------------------------------------------------------------
// element helpers
// ---------------------------------------------------------------------------------------------
template <typename T> struct Elem;
template <> struct Elem<float> {
    static __device__ __forceinline__ float4 load4(const float* p) { return *reinterpret_cast<const float4*>(p); }
    static __device__ __forceinline__ void store4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
    static __device__ __forceinline__ float2 load2(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
    static __device__ __forceinline__ void store2(float* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }
    static __device__ __forceinline__ float ld(const float* p) { return *p; }
    static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
};
template <> struct Elem<bf16> {
    static __device__ __forceinline__ float4 load4(const bf16* p) {
        uint2 raw = *reinterpret_cast<const uint2*>(p);
        __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&raw.x);
        __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&raw.y);
        float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
        return make_float4(fa.x, fa.y, fb.x, fb.y);
    }
    static __device__ __forceinline__ void store4(bf16* p, float4 v) {
        __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        uint2 raw;
        raw.x = *reinterpret_cast<uint32_t*>(&a);
        raw.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(p) = raw;
    }
    static __device__ __forceinline__ float2 load2(const bf16* p) {
        const uint32_t raw = __ldg(reinterpret_cast<const uint32_t*>(p));
        return make_float2(__uint_as_float(raw << 16), __uint_as_float(raw & 0xffff0000u));
    }
    static __device__ __forceinline__ void store2(bf16* p, float2 v) {
        __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
        *reinterpret_cast<uint32_t*>(p) = *reinterpret_cast<uint32_t*>(&a);
    }
    static __device__ __forceinline__ float ld(const bf16* p) { return __bfloat162float(*p); }
    static __device__ __forceinline__ void st(bf16* p, float v) { *p = __float2bfloat16_rn(v); }
};

// DT_SPLIT storage (common.cuh): T = __half addresses the hi plane, the lo plane sits `plane` elements further.  Loads return
// hi + lo (exact in fp32: two non-overlapping 11-bit pieces) in the tensor's STORED scale; stores split a stored-scale value.
struct SplitIO {
    static __device__ __forceinline__ void split1(float v, __half& h, __half& l) {
        v = fminf(fmaxf(v, -65504.f), 65504.f);
        h = __float2half_rn(v);
        l = __float2half_rn(v - __half2float(h));
    }
    static __device__ __forceinline__ float2 load2(const __half* p, long long plane) {
        const float2 a = __half22float2(__ldg(reinterpret_cast<const __half2*>(p)));
        const float2 b = __half22float2(__ldg(reinterpret_cast<const __half2*>(p + plane)));
        return make_float2(a.x + b.x, a.y + b.y);
    }
    static __device__ __forceinline__ void store2(__half* p, long long plane, float2 v) {
        __half h0, l0, h1, l1;
        split1(v.x, h0, l0); split1(v.y, h1, l1);
        *reinterpret_cast<__half2*>(p) = __halves2half2(h0, h1);
        *reinterpret_cast<__half2*>(p + plane) = __halves2half2(l0, l1);
    }
    static __device__ __forceinline__ float4 load4(const __half* p, long long plane) {
        const uint2 ra = __ldg(reinterpret_cast<const uint2*>(p)), rb = __ldg(reinterpret_cast<const uint2*>(p + plane));
        const float2 a0 = __half22float2(*reinterpret_cast<const __half2*>(&ra.x)), a1 = __half22float2(*reinterpret_cast<const __half2*>(&ra.y));
        const float2 b0 = __half22float2(*reinterpret_cast<const __half2*>(&rb.x)), b1 = __half22float2(*reinterpret_cast<const __half2*>(&rb.y));
        return make_float4(a0.x + b0.x, a0.y + b0.y, a1.x + b1.x, a1.y + b1.y);
    }
    static __device__ __forceinline__ void store4(__half* p, long long plane, float4 v) {
        __half h[4], l[4];
        split1(v.x, h[0], l[0]); split1(v.y, h[1], l[1]); split1(v.z, h[2], l[2]); split1(v.w, h[3], l[3]);
        uint2 a, b;
        __half2 t;
        t = __halves2half2(h[0], h[1]); a.x = *reinterpret_cast<uint32_t*>(&t);
        t = __halves2half2(h[2], h[3]); a.y = *reinterpret_cast<uint32_t*>(&t);
        t = __halves2half2(l[0], l[1]); b.x = *reinterpret_cast<uint32_t*>(&t);
        t = __halves2half2(l[2], l[3]); b.y = *reinterpret_cast<uint32_t*>(&t);
        *reinterpret_cast<uint2*>(p) = a;
        *reinterpret_cast<uint2*>(p + plane) = b;
    }
};
// plane-aware accessors used by the bandwidth kernels (plane is ignored by the single-plane types)
template <typename T> __device__ __forceinline__ float2 ld2p(const T* p, long long) { return Elem<T>::load2(p); }
template <> __device__ __forceinline__ float2 ld2p<__half>(const __half* p, long long plane) { return SplitIO::load2(p, plane); }
template <typename T> __device__ __forceinline__ void st2p(T* p, long long, float2 v) { Elem<T>::store2(p, v); }
template <> __device__ __forceinline__ void st2p<__half>(__half* p, long long plane, float2 v) { SplitIO::store2(p, plane, v); }
template <typename T> __device__ __forceinline__ float4 ld4p(const T* p, long long) { return Elem<T>::load4(p); }
template <> __device__ __forceinline__ float4 ld4p<__half>(const __half* p, long long plane) { return SplitIO::load4(p, plane); }
template <typename T> __device__ __forceinline__ void st4p(T* p, long long, float4 v) { Elem<T>::store4(p, v); }
template <> __device__ __forceinline__ void st4p<__half>(__half* p, long long plane, float4 v) { SplitIO::store4(p, plane, v); }
// running maximum of |stored value| of a DT_SPLIT tensor: one atomic per warp at the end of a kernel
// (some threads of a warp may already have left the kernel: reduce over the active lanes only; the bit patterns of
// non-negative floats order like unsigned integers)
__device__ __forceinline__ void publish_amax_simt(unsigned* slot, float amax) {
    if (slot == nullptr) return;
    const unsigned m = __activemask();
    const unsigned v = __reduce_max_sync(m, __float_as_uint(amax));
    if ((int)(threadIdx.x & 31) == __ffs(m) - 1 && v != 0u) atomicMax(slot, v);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

// ---------------------------------------------------------------------------------------------
// implicit-GEMM convolution, fp32 FFMA, NHWC, multi-source K-split
//   M = B*Hout*Wout (pixels), N = Cout, K = k*k*Cin.  CTA tile BM x BN, thread tile 4 x 4, BK = 16.
// ---------------------------------------------------------------------------------------------
template <typename T, int BN>
__global__ void __launch_bounds__(256) conv_simt_kernel(const ConvParams p) {
    constexpr int NT = BN / 4;       // threads along N
    constexpr int MT = 256 / NT;     // threads along M
    constexpr int BM = MT * 4;
    constexpr int BK = 16;
    constexpr int APT = BM / 64;     // A pixels per thread
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN];

    pdl_sync();
    const int tid = threadIdx.x;
    const int tx = tid % NT, ty = tid / NT;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int M = p.B * p.Hout * p.Wout;

    const int a_cg = tid & 3, a_px = tid >> 2;
    int pn[APT], poy[APT], pox[APT];
    bool pv[APT];
#pragma unroll
    for (int j = 0; j < APT; ++j) {
        int m = m0 + a_px + j * 64;
        pv[j] = m < M;
        int mm = pv[j] ? m : 0;
        pox[j] = mm % p.Wout;
        int t = mm / p.Wout;
        poy[j] = t % p.Hout;
        pn[j] = t / p.Hout;
    }
    const int b_row = tid / NT, b_c4 = tid % NT;     // B tile: BK rows x NT float4

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int tap = 0; tap < p.k * p.k; ++tap) {
        const int ky = tap / p.k, kx = tap % p.k;
        long long off[APT];
        int offx[APT];
        bool inb[APT];
#pragma unroll
        for (int j = 0; j < APT; ++j) {
            int iy = poy[j] * p.stride - p.pad + ky, ix = pox[j] * p.stride - p.pad + kx;
            inb[j] = pv[j] && iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win;
            off[j] = ((long long)pn[j] * p.Hin + iy);      // row index; the column is added per source (pitch / x offset)
            offx[j] = ix;
        }
        int cbase = 0;
        for (int s = 0; s < p.nsrc; ++s) {
            const int Cs = p.srcC[s];
            const T* sp = reinterpret_cast<const T*>(p.src[s]);
            for (int c0 = 0; c0 < Cs; c0 += BK) {
                const bool cvalid = (c0 + a_cg * 4) < Cs;
#pragma unroll
                for (int j = 0; j < APT; ++j) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (inb[j] && cvalid)
                        v = Elem<T>::load4(sp + (off[j] * p.srcWp[s] + offx[j] + p.srcXoff[s]) * Cs + c0 + a_cg * 4);
                    As[a_cg * 4 + 0][a_px + j * 64] = v.x;
                    As[a_cg * 4 + 1][a_px + j * 64] = v.y;
                    As[a_cg * 4 + 2][a_px + j * 64] = v.z;
                    As[a_cg * 4 + 3][a_px + j * 64] = v.w;
                }
                if (b_row < BK) {
                    float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (c0 + b_row < Cs)
                        wv = *reinterpret_cast<const float4*>(
                            p.w + ((long long)tap * p.Cin + cbase + c0 + b_row) * p.Cout + n0 + b_c4 * 4);
                    *reinterpret_cast<float4*>(&Bs[b_row][b_c4 * 4]) = wv;
                }
                __syncthreads();
#pragma unroll
                for (int kk = 0; kk < BK; ++kk) {
                    float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
                    float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
                    float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
                }
                __syncthreads();
            }
            cbase += Cs;
        }
    }

    // epilogue: folded BN / bias, residual, ReLU
    const int n = n0 + tx * 4;
    const float4 sc = *reinterpret_cast<const float4*>(p.scale + n);
    const float4 sh = *reinterpret_cast<const float4*>(p.shift + n);
    T* dst = reinterpret_cast<T*>(p.dst);
    const T* res = reinterpret_cast<const T*>(p.residual);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
        float4 v = make_float4(fmaf(acc[i][0], sc.x, sh.x), fmaf(acc[i][1], sc.y, sh.y), fmaf(acc[i][2], sc.z, sh.z),
                               fmaf(acc[i][3], sc.w, sh.w));
        if (res) {
            float4 r = Elem<T>::load4(res + (long long)m * p.Cout + n);
            v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
        }
        if (p.relu) {
            v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
        }
        Elem<T>::store4(dst + (long long)m * p.Cout + n, v);
    }
}

template <typename T>
static void launch_conv_simt_t(const ConvParams& p, cudaStream_t st) {
    const int M = p.B * p.Hout * p.Wout;
    if (p.Cout % 64 == 0) {
        dim3 grid((M + 63) / 64, p.Cout / 64);
        launch_k(conv_simt_kernel<T, 64>, grid, dim3(256), 0, st, p);
    } else if (p.Cout % 32 == 0) {
        dim3 grid((M + 127) / 128, p.Cout / 32);
        launch_k(conv_simt_kernel<T, 32>, grid, dim3(256), 0, st, p);
    } else {
        MC_CHECK(p.Cout % 16 == 0, "conv_simt: Cout must be a multiple of 16");
        dim3 grid((M + 255) / 256, p.Cout / 16);
        launch_k(conv_simt_kernel<T, 16>, grid, dim3(256), 0, st, p);
    }
    MC_CUDA(cudaGetLastError());
}

void launch_conv_simt(const ConvParams& p, DType dt, cudaStream_t st) {
    for (int s = 0; s < p.nsrc; ++s) MC_CHECK(p.srcC[s] % 4 == 0, "conv_simt: source channels must be a multiple of 4");
    if (dt == DT_F32) launch_conv_simt_t<float>(p, st);
    else launch_conv_simt_t<bf16>(p, st);
}

// ---------------------------------------------------------------------------------------------
// layout kernels
// ---------------------------------------------------------------------------------------------
// one thread per physical pixel: three coalesced plane reads, one 16-byte store (C = 8 bf16 or C = 4 fp32)
template <typename T, int CPAD>
__global__ void pack_input_kernel(const float* __restrict__ img, T* __restrict__ dst, int B, int C, int H, int W, int Wp,
                                  int xoff) {
    pdl_sync();
    const int total = B * H * Wp;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int x = i % Wp - xoff;
        const int t = i / Wp;
        const int y = t % H, b = t / H;
        float v[CPAD];
#pragma unroll
        for (int c = 0; c < CPAD; ++c) v[c] = 0.f;
        if (x >= 0 && x < W) {
#pragma unroll
            for (int c = 0; c < CPAD; ++c)
                if (c < C) v[c] = __ldg(img + (((long long)b * C + c) * H + y) * W + x);
        }
        T* o = dst + (long long)i * CPAD;
#pragma unroll
        for (int c = 0; c < CPAD; c += 4) Elem<T>::store4(o + c, make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]));
    }
}

// bf16, 8-channel padded destination, even W / Wp / xoff: one thread = two adjacent pixels (one 8-byte load per colour
// plane, one 32-byte store), one CTA per row segment -- no per-thread index divisions, every warp instruction touches
// contiguous memory.  Only the interior is written: the padding columns of the destination are zero from allocation
// (DeviceArena::alloc clears) and nothing else ever writes them.
__global__ void __launch_bounds__(128) pack_input_pair_kernel(const float* __restrict__ img, bf16* __restrict__ dst, int C, int H,
                                                              int W, int Wp, int xoff, int segs) {
    pdl_sync();
    const int row = blockIdx.x / segs;                     // b * H + y
    const int xp = (blockIdx.x - row * segs) * 128 + threadIdx.x;      // pixel pair index along x
    if (2 * xp >= W) return;
    const int b = row / H, y = row - b * H;
    const long long plane = (long long)H * W;
    const float* s0 = img + ((long long)b * C * H + y) * W + 2 * xp;
    float2 v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = c < C ? __ldg(reinterpret_cast<const float2*>(s0 + c * plane)) : make_float2(0.f, 0.f);
    uint32_t o[8];
    o[0] = pack_bf16x2(v[0].x, v[1].x); o[1] = pack_bf16x2(v[2].x, 0.f); o[2] = 0u; o[3] = 0u;
    o[4] = pack_bf16x2(v[0].y, v[1].y); o[5] = pack_bf16x2(v[2].y, 0.f); o[6] = 0u; o[7] = 0u;
    bf16* d = dst + ((long long)row * Wp + xoff + 2 * xp) * 8;
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(d), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]) : "memory");
}

// DT_SPLIT destination (8-channel padded, two fp16 planes): one thread per pixel, three coalesced plane reads, one 16-byte
// store per plane.  Only the interior is written (the padding columns stay zero from allocation).
__device__ __forceinline__ void store_input_pixel_hl(__half* d, const float (&v)[3]) {
    // one 16-byte pixel [hi0 hi1 hi2 0 lo0 lo1 lo2 0]
    __half h[3], l[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) SplitIO::split1(v[c], h[c], l[c]);
    uint4 o;
    __half2 t;
    const __half z = __float2half_rn(0.f);
    t = __halves2half2(h[0], h[1]); o.x = *reinterpret_cast<uint32_t*>(&t);
    t = __halves2half2(h[2], z); o.y = *reinterpret_cast<uint32_t*>(&t);
    t = __halves2half2(l[0], l[1]); o.z = *reinterpret_cast<uint32_t*>(&t);
    t = __halves2half2(l[2], z); o.w = *reinterpret_cast<uint32_t*>(&t);
    *reinterpret_cast<uint4*>(d) = o;
}

__global__ void __launch_bounds__(256) pack_input_split_kernel(const float* __restrict__ img, __half* __restrict__ dst, int B, int C, int H,
                                                               int W, int Wp, int xoff, long long plane, const ActScale* sc, unsigned* amax_slot,
                                                               int interleaved) {
    pdl_sync();
    const float mul = sc ? sc->mul : 1.f;
    float amax = 0.f;
    const long long total = (long long)B * H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W);
        const long long t = i / W;
        const int y = (int)(t % H), b = (int)(t / H);
        float v[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 3; ++c)
            if (c < C) v[c] = __ldg(img + (((long long)b * C + c) * H + y) * W + x) * mul;
        amax = fmaxf(amax, fmaxf(fabsf(v[0]), fmaxf(fabsf(v[1]), fabsf(v[2]))));
        __half* d = dst + ((t * Wp) + xoff + x) * 8;
        if (interleaved) {
            store_input_pixel_hl(d, v);
        } else {
            SplitIO::store4(d, plane, make_float4(v[0], v[1], v[2], 0.f));
            SplitIO::store4(d + 4, plane, make_float4(0.f, 0.f, 0.f, 0.f));
        }
    }
    publish_amax_simt(amax_slot, amax);
}

void launch_pack_input(const float* img, void* dst, DType dt, int B, int C, int H, int W, int Cpad, int Wp, int xoff,
                       cudaStream_t st, const SplitInfo& so) {
    MC_CHECK((Cpad == 4 || Cpad == 8) && C <= Cpad, "pack_input: Cpad must be 4 or 8");
    if (dt == DT_SPLIT) {
        MC_CHECK(Cpad == 8 && C <= 3, "pack_input: the fp16-plane input has 8 padded channels");
        const long long total = (long long)B * H * W;
        const int grid = (int)std::min<long long>((total + 255) / 256, 148 * 16);
        launch_k(pack_input_split_kernel, dim3(grid), dim3(256), 0, st, img, (__half*)dst, B, C, H, W, Wp, xoff, so.plane, so.sc, so.amax, so.interleaved ? 1 : 0);
        return;
    }
    if (dt == DT_BF16 && Cpad == 8 && C <= 3 && W % 2 == 0 && Wp % 2 == 0 && xoff % 2 == 0 &&
        (reinterpret_cast<uintptr_t>(img) & 7) == 0 && (reinterpret_cast<uintptr_t>(dst) & 31) == 0) {
        const int segs = (W / 2 + 127) / 128;
        launch_k(pack_input_pair_kernel, dim3((unsigned)(B * H * segs)), dim3(128), 0, st, img, (bf16*)dst, C, H, W, Wp, xoff, segs);
        return;
    }
    const long long total = (long long)B * H * Wp;
    MC_CHECK(total < (1ll << 31), "pack_input: tensor too large for 32-bit indexing");
    int grid = (int)((total + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    if (dt == DT_F32) {
        if (Cpad == 4) launch_k(pack_input_kernel<float, 4>, dim3(grid), dim3(256), 0, st, img, (float*)dst, B, C, H, W, Wp, xoff);
        else launch_k(pack_input_kernel<float, 8>, dim3(grid), dim3(256), 0, st, img, (float*)dst, B, C, H, W, Wp, xoff);
    } else {
        if (Cpad == 4) launch_k(pack_input_kernel<bf16, 4>, dim3(grid), dim3(256), 0, st, img, (bf16*)dst, B, C, H, W, Wp, xoff);
        else launch_k(pack_input_kernel<bf16, 8>, dim3(grid), dim3(256), 0, st, img, (bf16*)dst, B, C, H, W, Wp, xoff);
    }
    MC_CUDA(cudaGetLastError());
}

// uint8 HWC frames -> normalised, zero-padded NHWC: Normalize + Pad + ToTensor of the reference's input pipeline
// (transforms/default_transforms.py:376-431) fused into the input packing.  lut[c * 256 + u] = float((u - mean_c) / std_c)
// is computed on the host in double exactly as numpy does; pixels outside an image's own (h0, w0) are written as zeros
// every call (Pad's canvas), since the frames of a batch may differ in size.
template <typename T, int CPAD>
__global__ void __launch_bounds__(256) pack_input_u8_kernel(const unsigned char* __restrict__ src, const int* __restrict__ hw,
                                                            const float* __restrict__ lut, T* __restrict__ dst, int B, int H0, int W0,
                                                            int H, int W, int Wp, int xoff, long long plane, const ActScale* sc,
                                                            unsigned* amax_slot, int interleaved) {
    pdl_sync();
    const float mul = sc ? sc->mul : 1.f;
    float amax = 0.f;
    const long long total = (long long)B * H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W);
        const long long t = i / W;
        const int y = (int)(t % H), b = (int)(t / H);
        float v[3] = {0.f, 0.f, 0.f};
        if (y < min(hw[2 * b], H0) && x < min(hw[2 * b + 1], W0)) {
            const unsigned char* s = src + (((long long)b * H0 + y) * W0 + x) * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) v[c] = __ldg(lut + c * 256 + s[c]) * mul;
        }
        amax = fmaxf(amax, fmaxf(fabsf(v[0]), fmaxf(fabsf(v[1]), fabsf(v[2]))));
        T* d = dst + ((t * Wp) + xoff + x) * CPAD;
        if (sizeof(T) == 2 && CPAD == 8 && interleaved) {
            store_input_pixel_hl(reinterpret_cast<__half*>(d), v);
            continue;
        }
        st4p<T>(d, plane, make_float4(v[0], v[1], v[2], 0.f));
        if (CPAD == 8) st4p<T>(d + 4, plane, make_float4(0.f, 0.f, 0.f, 0.f));
    }
    publish_amax_simt(amax_slot, amax);
}

void launch_pack_input_u8(const unsigned char* src, const int* hw, const float* lut, void* dst, DType dt, int B, int H0, int W0,
                          int H, int W, int Cpad, int Wp, int xoff, cudaStream_t st, const SplitInfo& so) {
    MC_CHECK(Cpad == 4 || Cpad == 8, "pack_input_u8: Cpad must be 4 or 8");
    const long long total = (long long)B * H * W;
    int grid = (int)std::min<long long>((total + 255) / 256, 148 * 16);
    const long long z = 0;
    const ActScale* nosc = nullptr;
    unsigned* noamax = nullptr;
    if (dt == DT_SPLIT) {
        MC_CHECK(Cpad == 8, "pack_input_u8: the fp16-plane input has 8 padded channels");
        launch_k(pack_input_u8_kernel<__half, 8>, dim3(grid), dim3(256), 0, st, src, hw, lut, (__half*)dst, B, H0, W0, H, W, Wp, xoff, so.plane, so.sc, so.amax, so.interleaved ? 1 : 0);
    } else if (dt == DT_F32) {
        if (Cpad == 4) launch_k(pack_input_u8_kernel<float, 4>, dim3(grid), dim3(256), 0, st, src, hw, lut, (float*)dst, B, H0, W0, H, W, Wp, xoff, z, nosc, noamax, 0);
        else launch_k(pack_input_u8_kernel<float, 8>, dim3(grid), dim3(256), 0, st, src, hw, lut, (float*)dst, B, H0, W0, H, W, Wp, xoff, z, nosc, noamax, 0);
    } else {
        if (Cpad == 4) launch_k(pack_input_u8_kernel<bf16, 4>, dim3(grid), dim3(256), 0, st, src, hw, lut, (bf16*)dst, B, H0, W0, H, W, Wp, xoff, z, nosc, noamax, 0);
        else launch_k(pack_input_u8_kernel<bf16, 8>, dim3(grid), dim3(256), 0, st, src, hw, lut, (bf16*)dst, B, H0, W0, H, W, Wp, xoff, z, nosc, noamax, 0);
    }
}

// NHWC <-> NCHW through a 32x32 shared-memory transpose (pixels x channels).
template <typename T>
__global__ void unpack_nchw_kernel(const T* __restrict__ src, float* __restrict__ dst, int C, int HW) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int pix = p0 + r, c = c0 + threadIdx.x;
        if (pix < HW && c < C) tile[r][threadIdx.x] = Elem<T>::ld(src + ((long long)b * HW + pix) * C + c);
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int c = c0 + r, pix = p0 + threadIdx.x;
        if (pix < HW && c < C) dst[((long long)b * C + c) * HW + pix] = tile[threadIdx.x][r];
    }
}

template <typename T>
__global__ void pack_nhwc_kernel(const float* __restrict__ src, T* __restrict__ dst, int C, int HW) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int c = c0 + r, pix = p0 + threadIdx.x;
        if (pix < HW && c < C) tile[r][threadIdx.x] = src[((long long)b * C + c) * HW + pix];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int pix = p0 + r, c = c0 + threadIdx.x;
        if (pix < HW && c < C) Elem<T>::st(dst + ((long long)b * HW + pix) * C + c, tile[threadIdx.x][r]);
    }
}

// DT_SPLIT twins of the two transposes (debug dumps and the operator tests)
__global__ void unpack_nchw_split_kernel(const __half* __restrict__ src, float* __restrict__ dst, int C, int HW, long long plane,
                                         const ActScale* sc) {
    __shared__ float tile[32][33];
    const float inv = sc ? sc->inv : 1.f;
    const int b = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int pix = p0 + r, c = c0 + threadIdx.x;
        if (pix < HW && c < C) {
            const long long o = ((long long)b * HW + pix) * C + c;
            tile[r][threadIdx.x] = (__half2float(src[o]) + __half2float(src[o + plane])) * inv;
        }
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int c = c0 + r, pix = p0 + threadIdx.x;
        if (pix < HW && c < C) dst[((long long)b * C + c) * HW + pix] = tile[threadIdx.x][r];
    }
}
__global__ void pack_nhwc_split_kernel(const float* __restrict__ src, __half* __restrict__ dst, int C, int HW, long long plane,
                                       const ActScale* sc, unsigned* amax_slot) {
    __shared__ float tile[32][33];
    const float mul = sc ? sc->mul : 1.f;
    float amax = 0.f;
    const int b = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int c = c0 + r, pix = p0 + threadIdx.x;
        if (pix < HW && c < C) tile[r][threadIdx.x] = src[((long long)b * C + c) * HW + pix];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int pix = p0 + r, c = c0 + threadIdx.x;
        if (pix < HW && c < C) {
            const float v = tile[threadIdx.x][r] * mul;
            amax = fmaxf(amax, fabsf(v));
            __half h, l;
            SplitIO::split1(v, h, l);
            const long long o = ((long long)b * HW + pix) * C + c;
            dst[o] = h;
            dst[o + plane] = l;
        }
    }
    publish_amax_simt(amax_slot, amax);
}

void launch_unpack_nchw(const void* src, DType dt, float* dst, int B, int C, int H, int W, cudaStream_t st, const SplitInfo& si) {
    dim3 grid((H * W + 31) / 32, (C + 31) / 32, B), block(32, 8);
    if (dt == DT_SPLIT) unpack_nchw_split_kernel<<<grid, block, 0, st>>>((const __half*)src, dst, C, H * W, si.plane, si.sc);
    else if (dt == DT_F32) unpack_nchw_kernel<float><<<grid, block, 0, st>>>((const float*)src, dst, C, H * W);
    else unpack_nchw_kernel<bf16><<<grid, block, 0, st>>>((const bf16*)src, dst, C, H * W);
    MC_CUDA(cudaGetLastError());
}

void launch_pack_nhwc(const float* src, void* dst, DType dt, int B, int C, int H, int W, cudaStream_t st, const SplitInfo& so) {
    dim3 grid((H * W + 31) / 32, (C + 31) / 32, B), block(32, 8);
    if (dt == DT_SPLIT) pack_nhwc_split_kernel<<<grid, block, 0, st>>>(src, (__half*)dst, C, H * W, so.plane, so.sc, so.amax);
    else if (dt == DT_F32) pack_nhwc_kernel<float><<<grid, block, 0, st>>>(src, (float*)dst, C, H * W);
    else pack_nhwc_kernel<bf16><<<grid, block, 0, st>>>(src, (bf16*)dst, C, H * W);
    MC_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------
// 2x2/s2 max-pool, NHWC, 4 channels per thread
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void maxpool2_kernel(const T* __restrict__ src, T* __restrict__ dst, int B, int C, int Hin, int Win) {
    pdl_sync();
    const int Ho = Hin / 2, Wo = Win / 2, C4 = C / 4;
    long long total = (long long)B * Ho * Wo * C4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c4 = (int)(i % C4);
        long long t = i / C4;
        int ox = (int)(t % Wo);
        t /= Wo;
        int oy = (int)(t % Ho);
        int b = (int)(t / Ho);
        const T* s = src + (((long long)b * Hin + oy * 2) * Win + ox * 2) * C + c4 * 4;
        float4 a = Elem<T>::load4(s), bb = Elem<T>::load4(s + C), c = Elem<T>::load4(s + (long long)Win * C),
               d = Elem<T>::load4(s + (long long)Win * C + C);
        float4 m = make_float4(fmaxf(fmaxf(a.x, bb.x), fmaxf(c.x, d.x)), fmaxf(fmaxf(a.y, bb.y), fmaxf(c.y, d.y)),
                               fmaxf(fmaxf(a.z, bb.z), fmaxf(c.z, d.z)), fmaxf(fmaxf(a.w, bb.w), fmaxf(c.w, d.w)));
        Elem<T>::store4(dst + i * 4, m);
    }
}

// DT_SPLIT: the maximum of hi + lo, copied as the (hi, lo) pair it came from -- source and destination share one scale, so
// nothing is re-rounded.  Ties (hi + lo equal) carry equal values whichever pair is taken.
__global__ void maxpool2_split_kernel(const __half* __restrict__ src, __half* __restrict__ dst, int B, int C, int Hin, int Win,
                                      long long plane_in, long long plane_out, unsigned* amax_slot) {
    pdl_sync();
    const int Ho = Hin / 2, Wo = Win / 2, C2 = C / 2;
    float amax = 0.f;
    long long total = (long long)B * Ho * Wo * C2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c2 = (int)(i % C2);
        long long t = i / C2;
        int ox = (int)(t % Wo);
        t /= Wo;
        int oy = (int)(t % Ho);
        int b = (int)(t / Ho);
        const __half* s = src + (((long long)b * Hin + oy * 2) * Win + ox * 2) * C + c2 * 2;
        const long long offs[4] = {0, C, (long long)Win * C, (long long)Win * C + C};
        float2 best = make_float2(0.f, 0.f);
        __half2 bh, bl;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const __half2 h = __ldg(reinterpret_cast<const __half2*>(s + offs[k])), l = __ldg(reinterpret_cast<const __half2*>(s + offs[k] + plane_in));
            const float2 hf = __half22float2(h), lf = __half22float2(l);
            const float2 v = make_float2(hf.x + lf.x, hf.y + lf.y);
            if (k == 0) { best = v; bh = h; bl = l; }
            else {
                if (v.x > best.x) { best.x = v.x; bh = __halves2half2(__low2half(h), __high2half(bh)); bl = __halves2half2(__low2half(l), __high2half(bl)); }
                if (v.y > best.y) { best.y = v.y; bh = __halves2half2(__low2half(bh), __high2half(h)); bl = __halves2half2(__low2half(bl), __high2half(l)); }
            }
        }
        amax = fmaxf(amax, fmaxf(fabsf(best.x), fabsf(best.y)));
        *reinterpret_cast<__half2*>(dst + i * 2) = bh;
        *reinterpret_cast<__half2*>(dst + i * 2 + plane_out) = bl;
    }
    publish_amax_simt(amax_slot, amax);
}

void launch_maxpool2(const void* src, void* dst, DType dt, int B, int C, int Hin, int Win, cudaStream_t st, const SplitInfo& si,
                     const SplitInfo& so) {
    MC_CHECK(C % 4 == 0 && Hin % 2 == 0 && Win % 2 == 0, "maxpool2 geometry");
    long long total = (long long)B * (Hin / 2) * (Win / 2) * (C / 4);
    int grid = (int)((total + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    if (dt == DT_SPLIT) {
        grid = (int)std::min<long long>((total * 2 + 255) / 256, 148 * 32);
        launch_k(maxpool2_split_kernel, dim3(grid), dim3(256), 0, st, (const __half*)src, (__half*)dst, B, C, Hin, Win, si.plane, so.plane, so.amax);
    } else if (dt == DT_F32) launch_k(maxpool2_kernel<float>, dim3(grid), dim3(256), 0, st, (const float*)src, (float*)dst, B, C, Hin, Win);
    else launch_k(maxpool2_kernel<bf16>, dim3(grid), dim3(256), 0, st, (const bf16*)src, (bf16*)dst, B, C, Hin, Win);
    MC_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------
// depthwise ConvTranspose2d k4 s2 p1 (IDAUp.up_i), NHWC, 8 channels per thread, weights staged in shared
// memory as [16 taps][C].
//   out[oy][ox] = sum_{ky,kx} in[(oy+1-ky)/2][(ox+1-kx)/2] * w[ky][kx]   for (oy+1-ky), (ox+1-kx) even & in range
// ---------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <> __device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void load8<bf16>(const bf16* p, float (&v)[8]) {
    const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
        v[2 * j] = f.x; v[2 * j + 1] = f.y;
    }
}
template <typename T> __device__ __forceinline__ void store8(T* p, const float (&v)[8]);
template <> __device__ __forceinline__ void store8<float>(float* p, const float (&v)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void store8<bf16>(bf16* p, const float (&v)[8]) {
    uint4 o;
    __nv_bfloat162 h;
    h = __floats2bfloat162_rn(v[0], v[1]); o.x = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2bfloat162_rn(v[2], v[3]); o.y = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2bfloat162_rn(v[4], v[5]); o.z = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2bfloat162_rn(v[6], v[7]); o.w = *reinterpret_cast<uint32_t*>(&h);
    *reinterpret_cast<uint4*>(p) = o;
}

// one thread = the 2x2 output quad of input pixel (iy, ix) x 4 channels: nine 8-byte loads, four 8-byte stores,
// 64 FMAs; the 16 x C filter taps sit in shared memory as [tap][C] (one LDS.128 per tap and thread).
//   out[2iy  ][.] <- rows (iy-1, ky=3), (iy, ky=1)        out[2iy+1][.] <- rows (iy, ky=2), (iy+1, ky=0)   (same along x)
template <typename T>
__global__ void __launch_bounds__(256, 4) upsample2_kernel(const T* __restrict__ src, T* __restrict__ dst,
                                                           const float* __restrict__ w, int B, int C, int Hin, int Win) {
    extern __shared__ __align__(16) float sw[];            // [16][C]
    for (int i = threadIdx.x; i < 16 * C; i += blockDim.x) {
        const int tap = i / C, c = i % C;
        sw[i] = w[c * 16 + tap];
    }
    __syncthreads();
    pdl_sync();          // weights are constants
    const int C4 = C / 4;
    const int total = B * Hin * Win * C4;
    const int Wo = Win * 2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c4 = i % C4;
        int t = i / C4;
        const int ix = t % Win;
        t /= Win;
        const int iy = t % Hin, b = t / Hin;
        float4 in[3][3];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int yy = iy + dy - 1, xx = ix + dx - 1;
                in[dy][dx] = (yy >= 0 && yy < Hin && xx >= 0 && xx < Win)
                                 ? Elem<T>::load4(src + ((long long)(b * Hin + yy) * Win + xx) * C + c4 * 4)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
        for (int py = 0; py < 2; ++py)
#pragma unroll
            for (int px = 0; px < 2; ++px) {
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                    for (int bq = 0; bq < 2; ++bq) {
                        // py = 0: (dy=0, ky=3), (dy=1, ky=1);  py = 1: (dy=1, ky=2), (dy=2, ky=0)
                        const int dy = py + a, ky = py == 0 ? 3 - 2 * a : 2 - 2 * a;
                        const int dx = px + bq, kx = px == 0 ? 3 - 2 * bq : 2 - 2 * bq;
                        const float4 wv = *reinterpret_cast<const float4*>(sw + (ky * 4 + kx) * C + c4 * 4);
                        const float4 v = in[dy][dx];
                        acc.x = fmaf(v.x, wv.x, acc.x);
                        acc.y = fmaf(v.y, wv.y, acc.y);
                        acc.z = fmaf(v.z, wv.z, acc.z);
                        acc.w = fmaf(v.w, wv.w, acc.w);
                    }
                Elem<T>::store4(dst + ((long long)(b * 2 * Hin + 2 * iy + py) * Wo + 2 * ix + px) * C + c4 * 4, acc);
            }
    }
}

// strip variant: one thread = two channels x a horizontal strip of input pixels.  The 16 x 2 filter taps live in
// registers for the whole strip, a 3 x 3 input window slides along x (three new 4-byte loads per step, a warp =
// 64 consecutive channels = 128 contiguous bytes per load / store), so the kernel issues no shared-memory traffic and
// ~1.75 bytes of L1 traffic per output byte (the quad kernel above: ~11, which made it L1-bound at 25 % of HBM speed).
// PF: input columns fetched as one batch (3 x PF independent loads in flight per thread) before they are consumed; the
// kernel is latency-bound otherwise (~24 resident warps per SM x 3 loads of 128 B each in flight)
template <typename T, int PF>
__global__ void __launch_bounds__(256) upsample2_strip_kernel(const T* __restrict__ src, T* __restrict__ dst,
                                                              const float* __restrict__ w, int B, int C, int Hin, int Win,
                                                              int strip, int nstrips, long long plane_in, long long plane_out,
                                                              const ActScale* sc_in, const ActScale* sc_out, unsigned* amax_slot) {
    const int C2 = C >> 1;
    // DT_SPLIT: stored-scale input -> stored-scale output (powers of two: the rescale is exact)
    const float rescale = (sc_in ? sc_in->inv : 1.f) * (sc_out ? sc_out->mul : 1.f);
    float amax = 0.f;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)B * Hin * nstrips * C2;
    if (gid >= total) return;
    const int cp = (int)(gid % C2);
    long long t = gid / C2;
    const int sx = (int)(t % nstrips);
    t /= nstrips;
    const int iy = (int)(t % Hin), b = (int)(t / Hin);
    // taps of channels 2cp, 2cp+1: w[c][ky][kx]
    float2 wr[16];
    {
        const float4* w4 = reinterpret_cast<const float4*>(w + (long long)cp * 32);
        float wa[16], wb[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 u = __ldg(w4 + i), v = __ldg(w4 + 4 + i);
            wa[4 * i] = u.x; wa[4 * i + 1] = u.y; wa[4 * i + 2] = u.z; wa[4 * i + 3] = u.w;
            wb[4 * i] = v.x; wb[4 * i + 1] = v.y; wb[4 * i + 2] = v.z; wb[4 * i + 3] = v.w;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) wr[i] = make_float2(wa[i], wb[i]);
    }
    pdl_sync();          // weights are constants
    const int x_begin = sx * strip, x_end = min(Win, x_begin + strip);
    // 32-bit element offsets (the launcher checks the tensors are < 2^31 elements), advanced by C per step
    const int rowC = Win * C;
    const int base_in = (b * Hin + iy) * rowC + cp * 2;
    const bool ok0 = iy > 0, ok2 = iy + 1 < Hin;
    const T* r0 = src + base_in - (ok0 ? rowC : 0);
    const T* r1 = src + base_in;
    const T* r2 = src + base_in + (ok2 ? rowC : 0);
    const float2 zero2 = make_float2(0.f, 0.f);
    float2 win[3][3];
    {
        const bool okl = x_begin > 0;
        const int xl = (x_begin - 1) * C, xc = x_begin * C;
        win[0][1] = (ok0 && okl) ? ld2p<T>(r0 + xl, plane_in) : zero2;
        win[1][1] = okl ? ld2p<T>(r1 + xl, plane_in) : zero2;
        win[2][1] = (ok2 && okl) ? ld2p<T>(r2 + xl, plane_in) : zero2;
        win[0][2] = ok0 ? ld2p<T>(r0 + xc, plane_in) : zero2;
        win[1][2] = ld2p<T>(r1 + xc, plane_in);
        win[2][2] = ok2 ? ld2p<T>(r2 + xc, plane_in) : zero2;
    }
    const int oRow = 2 * rowC;                               // elements per output row
    T* o = dst + ((b * 2 * Hin + 2 * iy) * 2 * Win + 2 * x_begin) * C + cp * 2;
    int xn = (x_begin + 1) * C;
    for (int ix0 = x_begin; ix0 < x_end; ix0 += PF) {
        float2 nxt[PF][3];
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            const int ix = ix0 + u;
            const bool okr = (ix + 1 < Win) && (ix < x_end);
            nxt[u][0] = (ok0 && okr) ? ld2p<T>(r0 + xn + u * C, plane_in) : zero2;
            nxt[u][1] = okr ? ld2p<T>(r1 + xn + u * C, plane_in) : zero2;
            nxt[u][2] = (ok2 && okr) ? ld2p<T>(r2 + xn + u * C, plane_in) : zero2;
        }
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            if (ix0 + u >= x_end) break;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) { win[dy][0] = win[dy][1]; win[dy][1] = win[dy][2]; win[dy][2] = nxt[u][dy]; }
#pragma unroll
            for (int py = 0; py < 2; ++py)
#pragma unroll
                for (int px = 0; px < 2; ++px) {
                    float2 acc = zero2;
#pragma unroll
                    for (int a = 0; a < 2; ++a)
#pragma unroll
                        for (int bq = 0; bq < 2; ++bq) {
                            // same tap order as the quad kernel (bit-identical fp32 accumulation)
                            const int dy = py + a, ky = py == 0 ? 3 - 2 * a : 2 - 2 * a;
                            const int dx = px + bq, kx = px == 0 ? 3 - 2 * bq : 2 - 2 * bq;
                            const float2 wv = wr[ky * 4 + kx];
                            acc.x = fmaf(win[dy][dx].x, wv.x, acc.x);
                            acc.y = fmaf(win[dy][dx].y, wv.y, acc.y);
                        }
                    if (sizeof(T) == 2 && plane_out != 0) {
                        acc.x *= rescale; acc.y *= rescale;
                        amax = fmaxf(amax, fmaxf(fabsf(acc.x), fabsf(acc.y)));
                    }
                    st2p<T>(o + py * oRow + px * C, plane_out, acc);
                }
            o += 2 * C;
        }
        xn += PF * C;
    }
    publish_amax_simt(amax_slot, amax);
}

void launch_upsample2(const void* src, void* dst, DType dt, const float* w, int B, int C, int Hin, int Win,
                      cudaStream_t st, const SplitInfo& si, const SplitInfo& so) {
    MC_CHECK(C % 4 == 0 && C <= 512, "upsample2: C must be a multiple of 4 and <= 512");
    const long long total = (long long)B * Hin * Win * (C / 4);
    MC_CHECK(total * 4 < (1ll << 31), "upsample2: tensor too large for 32-bit indexing");
    static const bool quad = [] { const char* e = std::getenv("MC_UP_QUAD"); return e && e[0] == '1'; }();
    MC_CHECK(dt != DT_SPLIT || (long long)B * Hin * Win * C * 4 < (1ll << 31), "upsample2: fp16-plane tensors use the strip kernel (32-bit offsets)");
    if ((!quad || dt == DT_SPLIT) && (long long)B * Hin * Win * C * 4 < (1ll << 31)) {       // 32-bit offsets into the 4x larger output
        // strip length: long strips amortise the window start-up, short ones keep >= ~48 warps per SM in flight
        int strip = 16;
        while (strip > 4 && (long long)B * Hin * ((Win + strip - 1) / strip) * (C / 2) < 148ll * 2048) strip >>= 1;
        if (const char* e = std::getenv("MC_UP_STRIP")) strip = std::max(2, std::atoi(e));
        const int nstrips = (Win + strip - 1) / strip;
        const long long threads = (long long)B * Hin * nstrips * (C / 2);
        const int grid = (int)((threads + 255) / 256);
        static const int pf = [] { const char* e = std::getenv("MC_UP_PF"); return (e && e[0]) ? std::atoi(e) : 2; }();   // measured: PF 1 / 2 / 4 / 8 -> 0.048 / 0.037 / 0.052 / 0.064 ms (ida_2 up-sampling): registers cost more occupancy than the deeper prefetch gains
        const long long z = 0;
        const ActScale* nosc = nullptr;
        unsigned* noamax = nullptr;
        if (dt == DT_SPLIT) {
            // fp16 planes: every window column is already two loads (hi, lo) per row, prefetch depth 1 measured best
            // (0.072 / 0.078 / 0.093 ms for PF 1 / 2 / 4 on the ida_2 up-samplings; MC_UP_PF_SPLIT overrides)
            static const int pfs = [] { const char* e = std::getenv("MC_UP_PF_SPLIT"); return (e && e[0]) ? std::atoi(e) : 1; }();
            auto kern = pfs == 1 ? upsample2_strip_kernel<__half, 1> : (pfs == 4 ? upsample2_strip_kernel<__half, 4> : upsample2_strip_kernel<__half, 2>);
            launch_k(kern, dim3(grid), dim3(256), 0, st, (const __half*)src, (__half*)dst, w, B, C, Hin, Win, strip, nstrips, si.plane, so.plane, si.sc, so.sc, so.amax);
        }
        else if (dt == DT_F32) launch_k(upsample2_strip_kernel<float, 2>, dim3(grid), dim3(256), 0, st, (const float*)src, (float*)dst, w, B, C, Hin, Win, strip, nstrips, z, z, nosc, nosc, noamax);
        else if (pf == 1) launch_k(upsample2_strip_kernel<bf16, 1>, dim3(grid), dim3(256), 0, st, (const bf16*)src, (bf16*)dst, w, B, C, Hin, Win, strip, nstrips, z, z, nosc, nosc, noamax);
        else if (pf == 2) launch_k(upsample2_strip_kernel<bf16, 2>, dim3(grid), dim3(256), 0, st, (const bf16*)src, (bf16*)dst, w, B, C, Hin, Win, strip, nstrips, z, z, nosc, nosc, noamax);
        else if (pf == 8) launch_k(upsample2_strip_kernel<bf16, 8>, dim3(grid), dim3(256), 0, st, (const bf16*)src, (bf16*)dst, w, B, C, Hin, Win, strip, nstrips, z, z, nosc, nosc, noamax);
        else launch_k(upsample2_strip_kernel<bf16, 4>, dim3(grid), dim3(256), 0, st, (const bf16*)src, (bf16*)dst, w, B, C, Hin, Win, strip, nstrips, z, z, nosc, nosc, noamax);
        return;
    }
    int grid = (int)((total + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    const size_t smem = sizeof(float) * 16 * C;
    if (dt == DT_F32) launch_k(upsample2_kernel<float>, dim3(grid), dim3(256), smem, st, (const float*)src, (float*)dst, w, B, C, Hin, Win);
    else launch_k(upsample2_kernel<bf16>, dim3(grid), dim3(256), smem, st, (const bf16*)src, (bf16*)dst, w, B, C, Hin, Win);
}

}  // namespace mc
