// Shared epilogue of the tcgen05 convolution kernels: TMEM -> registers -> folded BN / bias (+residual) (+ReLU)
// -> NHWC in one of three output modes.  One thread owns one accumulator row (= one output pixel); it drains
// NCH x 16 columns per call with all TMEM loads and all residual loads in flight before the first use (the epilogue
// is latency-bound otherwise), and moves 32 bytes per global instruction (LDG/STG.256, sm_100).
//
//   OM_BF16  : bf16 NHWC (throughput mode)
//   OM_SPLIT : the fp32-accurate mode's storage, two fp16 planes hi + lo of value * 2^e (common.cuh, DT_SPLIT); the
//              residual is read in the same format.  The kernel folds the scales into the per-channel constants it
//              stages in shared memory (sc = scale * 2^-e_in * 2^e_out, sh = shift * 2^e_out), so the only extra work
//              here is the residual's rescale, the split itself and the running maximum of |stored value|.
//   OM_F32   : fp32 NHWC (the head stems of the fp32-accurate mode, consumed by AttnBN statistics + the 1x1 heads)
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace mc {
namespace tcepi {

enum { OM_BF16 = 0, OM_SPLIT = 1, OM_F32 = 2 };

// OM_SPLIT constants of one launch
struct SplitEpi {
    long long dst_plane;   // elements between the hi and lo planes of the destination
    long long res_plane;   // ... of the residual
    float res_mul;         // 2^-e_res * 2^e_out: residual's stored value -> the destination's stored scale
};

// Fused 2x2 / stride-2 max-pool of the convolution's own output (Tree.downsample, dla.py:179,193), OM_SPLIT only.  The four
// pixels of a window are accumulator rows of ONE epilogue warp (lanes l, l ^ 1, l ^ ybit, l ^ 1 ^ ybit), so the maximum is two
// shuffles per value; the even / even lane stores the pooled pixel.  The pooled tensor shares the source's scale, and
// max commutes with the (monotonic) hi + lo split, so the result equals pooling the stored tensor.
struct PoolEpi {
    void* dst;             // this thread's pooled pixel (hi plane, first column of the row's Cout tile); null unless it is the anchor lane
    long long plane;       // elements between the planes of the pooled tensor
    int ybit;              // lane distance of the pixel one row below
};

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void ldg256(const void* p, uint32_t (&r)[8]) {
    asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint32_t (&r)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
// scale / shift live in shared memory; through a generic pointer the compiler emits generic LD.E.128 (measured: the
// epilogue's "math" phase took 2400 clocks per 64 columns), so the loads are spelled as ld.shared
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
    float4 v;
    // volatile on purpose: as a plain asm the compiler hoists all 32 loads of a 64-column block to the top and spills
    // (5925 -> 5390 img/s)
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
// two floats -> (hi pair, lo pair) of fp16; saturating so that an out-of-range value stays finite (the running maximum
// reports it, the host rescales the tensor)
__device__ __forceinline__ void split_f16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
    a = fminf(fmaxf(a, -65504.f), 65504.f);
    b = fminf(fmaxf(b, -65504.f), 65504.f);
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ float2 unpack_f16x2(uint32_t v) { return __half22float2(*reinterpret_cast<const __half2*>(&v)); }

template <int OM> struct ElemBytes { static constexpr int value = (OM == OM_F32) ? 4 : 2; };

// taddr: TMEM address of (this warp's lane quarter, first column of the block)
// sc / sh: shared-memory scale / shift of the block's first column (16-byte aligned)
// per 16-column chunk i: res[i] / dst[i] global pointers (32-byte aligned; res[i] may be null), ok[i] = store it
//   (OM_SPLIT: pointers into the hi plane; OM_F32: dst[i] points at 16 floats)
// amax: running max of |stored value| (OM_SPLIT only)
// optional phase timing (diagnostics): t[0] += clocks until the TMEM / residual loads have landed, t[1] += the rest
template <int OM, int NCH>
__device__ __forceinline__ void drain_block_ex(uint32_t taddr, const float* sc, const float* sh, const void* const (&res)[NCH],
                                               void* const (&dst)[NCH], const bool (&ok)[NCH], bool relu, const SplitEpi& se, float& amax,
                                               long long* t = nullptr, const PoolEpi* pool = nullptr, int pool_col0 = 0) {
    uint32_t v[NCH][16];
    uint32_t r[NCH][8];
    uint32_t rl[OM == OM_SPLIT ? NCH : 1][8];
    const uint32_t sca = (uint32_t)__cvta_generic_to_shared(sc), sha = (uint32_t)__cvta_generic_to_shared(sh);
    const long long c0 = t ? clock64() : 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) tmem_ld16_nowait(taddr + 16u * i, v[i]);
    if (OM != OM_F32) {
#pragma unroll
        for (int i = 0; i < NCH; ++i)
            if (res[i] != nullptr && ok[i]) {
                ldg256(res[i], r[i]);
                if (OM == OM_SPLIT) ldg256(reinterpret_cast<const uint16_t*>(res[i]) + se.res_plane, rl[OM == OM_SPLIT ? i : 0]);
            }
    }
    tmem_wait_ld();
    if (t) {          // touch the last loaded registers so that the clock below is read after the data has really arrived
        uint32_t sink;
        asm volatile("add.u32 %0, %1, %2;" : "=r"(sink) : "r"(v[NCH - 1][15]), "r"(v[0][0]));
        asm volatile("" ::"r"(sink));
    }
    const long long c1 = t ? clock64() : 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        if (!ok[i] && !(OM == OM_SPLIT && pool != nullptr)) continue;      // pooling: every lane takes part in the shuffles
        float f[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 s4 = lds128(sca + (uint32_t)(16 * i + 4 * j) * 4u);
            const float4 h4 = lds128(sha + (uint32_t)(16 * i + 4 * j) * 4u);
            f[4 * j + 0] = fmaf(__uint_as_float(v[i][4 * j + 0]), s4.x, h4.x);
            f[4 * j + 1] = fmaf(__uint_as_float(v[i][4 * j + 1]), s4.y, h4.y);
            f[4 * j + 2] = fmaf(__uint_as_float(v[i][4 * j + 2]), s4.z, h4.z);
            f[4 * j + 3] = fmaf(__uint_as_float(v[i][4 * j + 3]), s4.w, h4.w);
        }
        if (OM != OM_F32 && res[i] != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (OM == OM_BF16) {
                    const float2 hf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r[i][j]));
                    f[2 * j] += hf.x;
                    f[2 * j + 1] += hf.y;
                } else {
                    // hi + lo is exact in fp32 (two non-overlapping 11-bit pieces)
                    const float2 a = unpack_f16x2(r[i][j]), b = unpack_f16x2(rl[OM == OM_SPLIT ? i : 0][j]);
                    f[2 * j] = fmaf(a.x + b.x, se.res_mul, f[2 * j]);
                    f[2 * j + 1] = fmaf(a.y + b.y, se.res_mul, f[2 * j + 1]);
                }
            }
        }
        if (relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
        }
        if (OM == OM_SPLIT && pool != nullptr) {                               // warp-uniform
            uint32_t ph[8], pl[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float a = ok[i] ? f[2 * j] : -3.0e38f, b = ok[i] ? f[2 * j + 1] : -3.0e38f;
                a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, 1));
                b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, 1));
                a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, pool->ybit));
                b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, pool->ybit));
                split_f16x2(a, b, ph[j], pl[j]);
            }
            if (pool->dst != nullptr) {
                uint16_t* pd = reinterpret_cast<uint16_t*>(pool->dst) + pool_col0 + 16 * i;
                stg256(pd, ph);
                stg256(pd + pool->plane, pl);
            }
            if (!ok[i]) continue;
        }
        if (OM == OM_BF16) {
            uint32_t o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = pack_bf16(f[2 * j], f[2 * j + 1]);
            if (t) { const long long s0 = clock64(); stg256(dst[i], o); t[2] += clock64() - s0; }
            else stg256(dst[i], o);
        } else if (OM == OM_SPLIT) {
            uint32_t o[8], ol[8];
            float m = amax;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                m = fmaxf(m, fmaxf(fabsf(f[2 * j]), fabsf(f[2 * j + 1])));
                split_f16x2(f[2 * j], f[2 * j + 1], o[j], ol[j]);
            }
            amax = m;
            stg256(dst[i], o);
            stg256(reinterpret_cast<uint16_t*>(dst[i]) + se.dst_plane, ol);
        } else {
            uint32_t o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = __float_as_uint(f[j]);
            stg256(dst[i], o);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = __float_as_uint(f[8 + j]);
            stg256(reinterpret_cast<float*>(dst[i]) + 8, o);
        }
    }
    if (t) { const long long c2 = clock64(); t[0] += c1 - c0; t[1] += c2 - c1; }
}

// contiguous variant: chunk i lives at dst + 16 i elements (one pixel, consecutive channels)
template <int OM, int NCH>
__device__ __forceinline__ void drain_block(uint32_t taddr, const float* sc, const float* sh, const void* res, void* dst, bool valid,
                                            bool relu, const SplitEpi& se, float& amax, long long* t = nullptr, const PoolEpi* pool = nullptr,
                                            int pool_col0 = 0) {
    constexpr int EB = ElemBytes<OM>::value;
    const void* rr[NCH];
    void* dd[NCH];
    bool ok[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        rr[i] = res ? reinterpret_cast<const char*>(res) + 16 * i * EB : nullptr;
        dd[i] = reinterpret_cast<char*>(dst) + 16 * i * EB;
        ok[i] = valid;
    }
    drain_block_ex<OM, NCH>(taddr, sc, sh, rr, dd, ok, relu, se, amax, t, pool, pool_col0);
}

// drains n_cols (multiple of 16) columns of one accumulator row; BIG: 64-column blocks (more loads in flight, more registers)
template <int OM, bool BIG>
__device__ __forceinline__ void drain_row(uint32_t taddr, int n_cols, const float* sc, const float* sh, const void* res, void* dst,
                                          bool valid, bool relu, const SplitEpi& se, float& amax, long long* t = nullptr,
                                          const PoolEpi* pool = nullptr) {
    constexpr int EB = ElemBytes<OM>::value;
    const char* r = reinterpret_cast<const char*>(res);
    char* d = reinterpret_cast<char*>(dst);
    int c0 = 0;
    if (BIG && OM == OM_BF16) {
        for (; c0 + 64 <= n_cols; c0 += 64)
            drain_block<OM, 4>(taddr + c0, sc + c0, sh + c0, r ? r + c0 * EB : nullptr, d + c0 * EB, valid, relu, se, amax, t);
    }
    for (; c0 + 32 <= n_cols; c0 += 32)
        drain_block<OM, 2>(taddr + c0, sc + c0, sh + c0, r ? r + c0 * EB : nullptr, d + c0 * EB, valid, relu, se, amax, t, pool, c0);
    if (c0 + 16 <= n_cols) drain_block<OM, 1>(taddr + c0, sc + c0, sh + c0, r ? r + c0 * EB : nullptr, d + c0 * EB, valid, relu, se, amax, t, pool, c0);
}

// ---- AttnBN instance statistics in the head-stem epilogue (OM_F32) -----------------------------------------------------
// Sum over the warp's 32 lanes (= 32 pixels) of 16 per-lane values, one column per lane pair: a butterfly that halves the
// number of live values at every step (8 + 4 + 2 + 1 + 1 = 16 shuffles).  Lane l returns the total of value index
// 8 b4 + 4 b3 + 2 b2 + b1 (b_k = bit k of l); lanes l and l ^ 1 hold the same column.
__device__ __forceinline__ float warp_colsum16(const float (&v)[16], int lane) {
    float a[8], b[4], c[2];
    const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0, b1 = (lane & 2) != 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = (b4 ? v[8 + j] : v[j]) + __shfl_xor_sync(0xffffffffu, b4 ? v[j] : v[8 + j], 16);
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = (b3 ? a[4 + j] : a[j]) + __shfl_xor_sync(0xffffffffu, b3 ? a[j] : a[4 + j], 8);
#pragma unroll
    for (int j = 0; j < 2; ++j) c[j] = (b2 ? b[2 + j] : b[j]) + __shfl_xor_sync(0xffffffffu, b2 ? b[j] : b[2 + j], 4);
    float d = (b1 ? c[1] : c[0]) + __shfl_xor_sync(0xffffffffu, b1 ? c[0] : c[1], 2);
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    return d;
}
__device__ __forceinline__ int warp_colsum16_index(int lane) { return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1); }

// What one epilogue warp knows about the statistics of its 32 rows: they belong to image A, or (a tile may straddle two images)
// partly to image A and partly to image B = A + 1.  sums: [image][channel][2] doubles (sum, sum of squares), fp64 atomics.
struct StatsEpi {
    double* sums_a;        // &sums[(A * Ctot + first column of this row's Cout tile) * 2]
    long long img_stride;  // doubles between images (Ctot * 2)
    bool in_a, in_b;       // this lane's pixel is a real pixel of image A / B
    bool one;              // warp-uniform: image A exists (A < batch; the last tile's rows run past the batch)
    bool two;              // warp-uniform: some lane of the warp belongs to image B, and B exists
};

// fp32 output + statistics: drains n_cols (multiple of 16) columns of one accumulator row, 32 columns at a time
__device__ __forceinline__ void drain_row_f32_stats(uint32_t taddr, int n_cols, const float* sc, const float* sh, float* dst, bool valid,
                                                    bool relu, const StatsEpi& st, int lane) {
    const uint32_t sca = (uint32_t)__cvta_generic_to_shared(sc), sha = (uint32_t)__cvta_generic_to_shared(sh);
    const int ci = warp_colsum16_index(lane);
    for (int c0 = 0; c0 < n_cols; c0 += 32) {
        const int nch = (c0 + 32 <= n_cols) ? 2 : 1;
        uint32_t v[2][16];
        tmem_ld16_nowait(taddr + c0, v[0]);
        if (nch == 2) tmem_ld16_nowait(taddr + c0 + 16, v[1]);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            if (i >= nch) break;
            float f[16];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 s4 = lds128(sca + (uint32_t)(c0 + 16 * i + 4 * j) * 4u);
                const float4 h4 = lds128(sha + (uint32_t)(c0 + 16 * i + 4 * j) * 4u);
                f[4 * j + 0] = fmaf(__uint_as_float(v[i][4 * j + 0]), s4.x, h4.x);
                f[4 * j + 1] = fmaf(__uint_as_float(v[i][4 * j + 1]), s4.y, h4.y);
                f[4 * j + 2] = fmaf(__uint_as_float(v[i][4 * j + 2]), s4.z, h4.z);
                f[4 * j + 3] = fmaf(__uint_as_float(v[i][4 * j + 3]), s4.w, h4.w);
            }
            if (relu) {
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
            }
            if (valid) {
                uint32_t o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = __float_as_uint(f[j]);
                stg256(dst + c0 + 16 * i, o);
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = __float_as_uint(f[8 + j]);
                stg256(dst + c0 + 16 * i + 8, o);
            }
            // statistics of image A (and of image B where the warp straddles two images)
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                if (pass == 0 && !st.one) continue;
                if (pass == 1 && !st.two) break;
                const bool mine = pass == 0 ? st.in_a : st.in_b;
                float x[16], x2[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) { x[j] = mine ? f[j] : 0.f; x2[j] = x[j] * x[j]; }
                const float s1 = warp_colsum16(x, lane), s2 = warp_colsum16(x2, lane);
                if (!(lane & 1)) {
                    double* p = st.sums_a + (long long)pass * st.img_stride + (long long)(c0 + 16 * i + ci) * 2;
                    atomicAdd(p, (double)s1);
                    atomicAdd(p + 1, (double)s2);
                }
            }
        }
    }
}

// end of an epilogue role: fold the thread's running maximum into the tensor's slot (one atomic per warp)
__device__ __forceinline__ void publish_amax(unsigned* slot, float amax) {
    if (slot == nullptr) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if ((threadIdx.x & 31) == 0 && amax > 0.f) atomicMax(slot, __float_as_uint(amax));
}

}  // namespace tcepi
}  // namespace mc
