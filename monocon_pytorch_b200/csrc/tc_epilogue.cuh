// Shared epilogue of the tcgen05 convolution kernels: TMEM -> registers -> folded BN / bias (+residual) (+ReLU)
// -> bf16 NHWC.  One thread owns one accumulator row (= one output pixel); it drains NCH x 16 columns per call
// with all TMEM loads and all residual loads in flight before the first use (the epilogue is latency-bound
// otherwise), and moves 32 bytes per global instruction (LDG/STG.256, sm_100).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace mc {
namespace tcepi {

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

// taddr: TMEM address of (this warp's lane quarter, first column of the block)
// sc / sh: shared-memory scale / shift of the block's first column (16-byte aligned)
// per 16-column chunk i: res[i] / dst[i] global pointers (32-byte aligned; res[i] may be null), ok[i] = store it
// optional phase timing (diagnostics): t[0] += clocks until the TMEM / residual loads have landed, t[1] += the rest
template <int NCH>
__device__ __forceinline__ void drain_block_ex(uint32_t taddr, const float* sc, const float* sh, const __nv_bfloat16* const (&res)[NCH],
                                               __nv_bfloat16* const (&dst)[NCH], const bool (&ok)[NCH], bool relu, long long* t = nullptr) {
    uint32_t v[NCH][16];
    uint32_t r[NCH][8];
    const uint32_t sca = (uint32_t)__cvta_generic_to_shared(sc), sha = (uint32_t)__cvta_generic_to_shared(sh);
    const long long c0 = t ? clock64() : 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) tmem_ld16_nowait(taddr + 16u * i, v[i]);
#pragma unroll
    for (int i = 0; i < NCH; ++i)
        if (res[i] != nullptr && ok[i]) ldg256(res[i], r[i]);
    tmem_wait_ld();
    if (t) {          // touch the last loaded registers so that the clock below is read after the data has really arrived
        uint32_t sink;
        asm volatile("add.u32 %0, %1, %2;" : "=r"(sink) : "r"(v[NCH - 1][15]), "r"(v[0][0]));
        asm volatile("" ::"r"(sink));
    }
    const long long c1 = t ? clock64() : 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        if (!ok[i]) continue;
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
        if (res[i] != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float2 hf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r[i][j]));
                f[2 * j] += hf.x;
                f[2 * j + 1] += hf.y;
            }
        }
        if (relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
        }
        uint32_t o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = pack_bf16(f[2 * j], f[2 * j + 1]);
        if (t) { const long long s0 = clock64(); stg256(dst[i], o); t[2] += clock64() - s0; }
        else stg256(dst[i], o);
    }
    if (t) { const long long c2 = clock64(); t[0] += c1 - c0; t[1] += c2 - c1; }
}

// contiguous variant: chunk i lives at dst + 16 i (one pixel, consecutive channels)
template <int NCH>
__device__ __forceinline__ void drain_block(uint32_t taddr, const float* sc, const float* sh, const __nv_bfloat16* res,
                                            __nv_bfloat16* dst, bool valid, bool relu, long long* t = nullptr) {
    const __nv_bfloat16* rr[NCH];
    __nv_bfloat16* dd[NCH];
    bool ok[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) { rr[i] = res ? res + 16 * i : nullptr; dd[i] = dst + 16 * i; ok[i] = valid; }
    drain_block_ex<NCH>(taddr, sc, sh, rr, dd, ok, relu, t);
}

// drains n_cols (multiple of 16) columns of one accumulator row
__device__ __forceinline__ void drain_row(uint32_t taddr, int n_cols, const float* sc, const float* sh, const __nv_bfloat16* res,
                                          __nv_bfloat16* dst, bool valid, bool relu, long long* t = nullptr) {
    int c0 = 0;
    for (; c0 + 64 <= n_cols; c0 += 64) drain_block<4>(taddr + c0, sc + c0, sh + c0, res ? res + c0 : nullptr, dst + c0, valid, relu, t);
    if (c0 + 32 <= n_cols) {
        drain_block<2>(taddr + c0, sc + c0, sh + c0, res ? res + c0 : nullptr, dst + c0, valid, relu, t);
        c0 += 32;
    }
    if (c0 + 16 <= n_cols) drain_block<1>(taddr + c0, sc + c0, sh + c0, res ? res + c0 : nullptr, dst + c0, valid, relu, t);
}

}  // namespace tcepi
}  // namespace mc
