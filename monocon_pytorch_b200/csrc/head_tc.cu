// MonoCon head "apply" stage on the tensor cores (sm_100a): bf16 throughput mode, and (SPLIT) the fp32-accurate mode.
//
//   per pixel and stem s:  z = relu(coefA[b,s,:] * x + coefB[b,s,:])   (AttnBatchNorm2d + ReLU, attentive_norm.py:154-164)
//                          y = W_s z + bias                            (the 1x1 convs that read stem s, monocon_heads.py:165-200)
//                          + sigmoid/clamp (heat maps, :168-170), inverse-sigmoid depth (:183); ten NCHW fp32 maps out.
//
// The SIMT kernel (kernels_head.cu) spends its time on the 64 x 65 FMAs per pixel and on the shared-memory transpose
// that puts a pixel's 64 channels into one thread.  Here a work unit is (128-pixel tile, stem):
//   warp 0        TMA: the unit's 128 x 64 bf16 slice of the stem tensor -> 16 KB shared-memory stage, 128-byte swizzle
//                 (exactly the K-major UMMA operand layout: one pixel = one 128-byte row)
//   warps 2..9    transform the stage in place: bf16 -> fp32 FMA with the per-(image, channel) coefficients -> ReLU -> bf16
//   warp 1        four tcgen05.mma (M = 128 pixels, N = 16 / 32 padded outputs of the stem, K = 4 x 16) into the stem's
//                 column range of a 176-column fp32 accumulator in TMEM (double buffered across tiles)
//   warps 10..13  epilogue: TMEM -> bias -> activation -> coalesced NCHW stores (a warp = 32 consecutive pixels)
// The 1x1 weights (bf16, zero-padded rows) stay resident in shared memory.  Per unit the SM moves 16 KB in (TMA), 16 KB
// out and in again (transform), 16 KB to the tensor core: ~512 clk of shared-memory bandwidth, below the HBM time of the
// same 16 KB at 148 SMs, so the stage is bound by reading the stem tensor once from HBM.
//
// SPLIT (MC_PREC_FP32_TC): the stems arrive as fp32 (two 128-pixel x 32-channel TMA boxes per unit, 32 KB).  The transform
// computes z in fp32, scales it by the calibrated power of two of the pseudo-tensor "head.z" (engine.h: Net::act_scale; its
// running maximum is tracked like every other fp16-plane tensor's) and splits it into fp16 hi + lo (common.cuh, DT_SPLIT), written IN PLACE over the
// staging buffer (hi tile over box 0, lo tile over box 1; the eight lanes that touch a pixel row sit in one warp, so a
// __syncwarp between the row batch's loads and stores is the only ordering needed).  The 1x1 weights are resident as fp16
// hi / lo pieces of w * 2^ew[o] (packed on the host at prepare time), each stem is 3 x 4 MMAs (z_hi w_lo, z_lo w_hi, z_hi w_hi),
// and the epilogue multiplies the accumulator by 2^-ew[o] * 2^-e_z before the bias.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstring>

#include "engine.h"
#include "tc_epilogue.cuh"

namespace mc {

namespace {

constexpr int kHtThreads = 448;
constexpr int kHtXformWarp0 = 2, kHtXformThreads = 256;
constexpr int kHtEpiWarp0 = 10;
constexpr int kHtGroups4 = 4;               // transform groups (two warps each); stages % kHtGroups4 == 0
template <bool SPLIT> struct HtCfg {
    static constexpr int kStages = SPLIT ? 4 : 8;
    static constexpr int kTileBytes = SPLIT ? 2 * 128 * 128 : 128 * 128;   // 128 pixels x 64 channels, fp32 (two boxes) / bf16
    static constexpr int kWBytes = (SPLIT ? 2 : 1) * 176 * 128;            // resident 1x1 weights: [hi][lo] x 176 rows x 128 B
};
constexpr int kHtCols = 176;                     // padded outputs: 16,16,16,32,16,16,16,16,32
constexpr int kHtAccStride = 256;
constexpr long long kHtSpin = 4000000000LL;

// stems in registration order (monocon_heads.py:74-88); output rows in pred order, see kernels_head.cu
__constant__ int c_ht_col[kNumStems] = {0, 16, 32, 48, 80, 96, 112, 128, 144};
__constant__ int c_ht_npad[kNumStems] = {16, 16, 16, 32, 16, 16, 16, 16, 32};
__constant__ int c_ht_o0[kNumStems] = {0, 12, 14, 18, 3, 16, 36, 39, 41};
__constant__ int c_ht_o1[kNumStems] = {3, 14, 16, 36, 12, 18, 39, 41, 65};

struct HeadTcParams {
    CUtensorMap map_x;       // stems [B][HW][576]: dims (576, HW, B); bf16: box (64, 128, 1); fp32 (SPLIT): box (32, 128, 1); SWIZZLE_128B
    const float* coefA;      // [B][576]
    const float* coefB;
    const float* w;          // [65][64] fp32
    const float* bias;       // [65]
    const uint16_t* w_split; // SPLIT: [2 (hi, lo)][176 padded rows][64] fp16 bits of w * 2^ew[row]
    const float* oscale;     // SPLIT: [80] per output row: 2^-ew
    const ActScale* z_sc;    // SPLIT: scale of the post-AttnBN activations z (pseudo-tensor "head.z")
    unsigned* z_amax;        // SPLIT: running max of |z * 2^e_z|
    float* out[kNumPred];
    int B, HW, tiles_per_img;
    int* error_flag;
};

__device__ __forceinline__ uint32_t h_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void hbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(h_u32(bar)), "r"(count));
}
__device__ __forceinline__ void hbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(h_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void hbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(h_u32(bar)) : "memory");
}
__device__ __forceinline__ bool hbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(h_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void hbar_wait(uint64_t* bar, uint32_t parity, int* error_flag, int code) {
    if (hbar_try(bar, parity)) return;
    const long long t0 = clock64();
    while (!hbar_try(bar, parity)) {
        if (clock64() - t0 > kHtSpin) {
            if (error_flag) atomicExch(error_flag, code);
            __threadfence_system();
            asm volatile("trap;");
        }
    }
}
__device__ __forceinline__ void h_tma3(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(h_u32(smem)), "l"(map), "r"(h_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void h_mma(uint32_t tmem_d, uint32_t alo, uint32_t blo, uint32_t hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
        ::"r"(tmem_d), "r"(alo), "r"(blo), "r"(hi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void h_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(h_u32(bar)) : "memory");
}
__device__ __forceinline__ bool h_elect() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// explicit shared-memory accesses (through generic pointers the compiler emits generic LD.E / ST.E here)
__device__ __forceinline__ uint4 h_lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ float h_lds32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void h_sts128(uint32_t a, const uint4& v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// two fp32 -> packed bf16x2 with ReLU (lo = a, hi = b)
__device__ __forceinline__ uint32_t relu_pack(float a, float b) {
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
    return d;
}

// ---- epilogue tables: 16-column accumulator groups -> first output row (pred order) and number of real outputs ----
constexpr int kHtGroups = 11;
__host__ __device__ constexpr int grp_o0(int g) {
    constexpr int t[kHtGroups] = {0, 12, 14, 18, 34, 3, 16, 36, 39, 41, 57};
    return t[g];
}
__host__ __device__ constexpr int grp_n(int g) {
    constexpr int t[kHtGroups] = {3, 2, 2, 16, 2, 9, 2, 3, 2, 16, 8};
    return t[g];
}
// output row o -> prediction map (pred order, monocon_heads.py:190-200)
__host__ __device__ constexpr int out_pred(int o) {
    constexpr int first[kNumPred + 1] = {0, 3, 12, 14, 16, 18, 36, 39, 41, 53, 65};
    int p = 0;
    for (int i = 0; i < kNumPred; ++i)
        if (o >= first[i]) p = i;
    return p;
}
__host__ __device__ constexpr int out_first(int pred) {
    constexpr int first[kNumPred + 1] = {0, 3, 12, 14, 16, 18, 36, 39, 41, 53, 65};
    return first[pred];
}
__host__ __device__ constexpr int out_nch(int pred) {
    constexpr int first[kNumPred + 1] = {0, 3, 12, 14, 16, 18, 36, 39, 41, 53, 65};
    return first[pred + 1] - first[pred];
}

template <int G, bool SPLIT>
__device__ __forceinline__ void epi_group(const uint32_t (&v)[16], uint32_t bs, const HeadTcParams& p, int b, int pix) {
#pragma unroll
    for (int c = 0; c < grp_n(G); ++c) {
        constexpr int o0 = grp_o0(G);
        const int o = o0 + c;                               // compile-time after unrolling
        const int pred = out_pred(o0 + c);
        const int ch = o - out_first(pred);
        const int nch = out_nch(pred);
        float acc = SPLIT ? fmaf(__uint_as_float(v[c]), h_lds32(bs + 4u * (uint32_t)(80 + o)), h_lds32(bs + 4u * (uint32_t)o))
                          : __uint_as_float(v[c]) + h_lds32(bs + 4u * (uint32_t)o);
        if (pred == 0 || pred == 1) {                       // monocon_heads.py:168-170
            acc = 1.f / (1.f + expf(-acc));
            acc = fminf(fmaxf(acc, 1e-4f), 1.f - 1e-4f);
        } else if (pred == 7 && ch == 0) {                  // monocon_heads.py:183
            acc = 1.f / (1.f / (1.f + expf(-acc)) + 1e-12f) - 1.f;
        }
        p.out[pred][((long long)b * nch + ch) * p.HW + pix] = acc;
    }
}

template <bool SPLIT>
__global__ void __launch_bounds__(kHtThreads, 1) head_apply_tc_kernel(const __grid_constant__ HeadTcParams p) {
    constexpr int kHtStages = HtCfg<SPLIT>::kStages, kHtTileBytes = HtCfg<SPLIT>::kTileBytes;
    extern __shared__ __align__(1024) uint8_t ht_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ht_smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;                                          // [kHtStages][16 / 32 KB]
    uint8_t* smem_w = smem + kHtStages * kHtTileBytes;               // [hi (, lo)][176 rows][128 B], swizzled
    float* bs = reinterpret_cast<float*>(smem_w + HtCfg<SPLIT>::kWBytes);    // [80] bias (+ [80] output scale)
    uint64_t* bars = reinterpret_cast<uint64_t*>(bs + 160);
    uint64_t* full = bars;                       // TMA landed
    uint64_t* ready = bars + kHtStages;          // transformed
    uint64_t* empty = bars + 2 * kHtStages;      // MMAs retired
    uint64_t* tmem_full = bars + 3 * kHtStages;  // [2]
    uint64_t* tmem_empty = tmem_full + 2;        // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // resident 1x1 weights: bf16, K-major, 128-byte swizzle (16-byte chunk index ^= row & 7), zero rows for the padding
    for (int idx = threadIdx.x; idx < kHtCols * 8; idx += kHtThreads) {
        const int row = idx >> 3, j = idx & 7;
        int s = 0;
#pragma unroll
        for (int i = 1; i < kNumStems; ++i)
            if (row >= c_ht_col[i]) s = i;
        const int r = row - c_ht_col[s];
        const int o = c_ht_o0[s] + r;
        if (SPLIT) {          // host-packed fp16 pieces, padded rows already zero
#pragma unroll
            for (int piece = 0; piece < 2; ++piece) {
                const uint4 v = *reinterpret_cast<const uint4*>(p.w_split + ((size_t)piece * kHtCols + row) * kStemC + j * 8);
                *reinterpret_cast<uint4*>(smem_w + piece * kHtCols * 128 + row * 128 + ((j ^ (r & 7)) << 4)) = v;
            }
            continue;
        }
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (o < c_ht_o1[s]) {
            const float4 a = *reinterpret_cast<const float4*>(p.w + o * kStemC + j * 8);
            const float4 c = *reinterpret_cast<const float4*>(p.w + o * kStemC + j * 8 + 4);
            v.x = tcepi::pack_bf16(a.x, a.y); v.y = tcepi::pack_bf16(a.z, a.w);
            v.z = tcepi::pack_bf16(c.x, c.y); v.w = tcepi::pack_bf16(c.z, c.w);
        }
        *reinterpret_cast<uint4*>(smem_w + row * 128 + ((j ^ (r & 7)) << 4)) = v;
    }
    if (threadIdx.x < 80) {
        bs[threadIdx.x] = threadIdx.x < kNumOut ? p.bias[threadIdx.x] : 0.f;
        bs[80 + threadIdx.x] = (SPLIT && threadIdx.x < kNumOut) ? p.oscale[threadIdx.x] * (p.z_sc ? p.z_sc->inv : 1.f) : 1.f;
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kHtStages; ++s) { hbar_init(&full[s], 1); hbar_init(&ready[s], kHtXformThreads / kHtGroups4); hbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { hbar_init(&tmem_full[a], 1); hbar_init(&tmem_empty[a], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&p.map_x) : "memory");
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(h_u32(tmem_ptr)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy weight writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_ptr;
    pdl_sync();          // weights / bias are constants; stems and coefficients come from the previous kernels

    const int tiles = p.B * p.tiles_per_img;

    if (warp == 0) {
        // ===================== TMA producer =====================
        int stage = 0;
        uint32_t phase = 0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
            const int b = t / p.tiles_per_img, p0 = (t % p.tiles_per_img) * 128;
            for (int s = 0; s < kNumStems; ++s) {
                hbar_wait(&empty[stage], phase ^ 1u, p.error_flag, 21);
                if (h_elect()) {
                    hbar_expect_tx(&full[stage], (uint32_t)kHtTileBytes);
                    h_tma3(smem_a + stage * kHtTileBytes, &p.map_x, &full[stage], s * kStemC, p0, b);
                    if (SPLIT) h_tma3(smem_a + stage * kHtTileBytes + 128 * 128, &p.map_x, &full[stage], s * kStemC + 32, p0, b);
                }
                __syncwarp();
                if (++stage == kHtStages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);           // SBO = 8 rows x 128 B, SWIZZLE_128B
        const uint32_t a_base16 = (1u << 16) | ((h_u32(smem_a) & 0x3FFFF) >> 4);
        const uint32_t w_base16 = (1u << 16) | ((h_u32(smem_w) & 0x3FFFF) >> 4);
        const uint32_t fmt = SPLIT ? 0u : 1u;                                          // fp16 pieces / bf16
        const uint32_t idesc_base = (1u << 4) | (fmt << 7) | (fmt << 10) | ((128u >> 4) << 24);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase[2] = {0u, 0u};
        for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
            hbar_wait(&tmem_empty[acc], acc_phase[acc] ^ 1u, p.error_flag, 22);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int s = 0; s < kNumStems; ++s) {
                const uint32_t col = (uint32_t)c_ht_col[s], npad = (uint32_t)c_ht_npad[s];
                const uint32_t idesc = idesc_base | ((npad >> 3) << 17);
                const uint32_t d = tmem_base + (uint32_t)(acc * kHtAccStride) + col;
                const uint32_t alo = a_base16 + (uint32_t)stage * (kHtTileBytes >> 4);
                const uint32_t blo = w_base16 + col * (128u >> 4);
                hbar_wait(&ready[stage], phase, p.error_flag, 23);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (h_elect()) {
                    if (SPLIT) {
                        // z_hi x w_lo, z_lo x w_hi (the small terms first: the TMEM accumulator truncates), then z_hi x w_hi
                        const uint32_t a_lo_tile = alo + ((128u * 128u) >> 4), b_lo_w = blo + (((uint32_t)kHtCols * 128u) >> 4);
#pragma unroll
                        for (int k = 0; k < 4; ++k) h_mma(d, alo + 2u * k, b_lo_w + 2u * k, hi, idesc, k > 0 ? 1u : 0u);
#pragma unroll
                        for (int k = 0; k < 4; ++k) h_mma(d, a_lo_tile + 2u * k, blo + 2u * k, hi, idesc, 1u);
#pragma unroll
                        for (int k = 0; k < 4; ++k) h_mma(d, alo + 2u * k, blo + 2u * k, hi, idesc, 1u);
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) h_mma(d, alo + 2u * k, blo + 2u * k, hi, idesc, k > 0 ? 1u : 0u);
                    }
                    h_commit(&empty[stage]);
                }
                __syncwarp();
                if (++stage == kHtStages) { stage = 0; phase ^= 1u; }
            }
            if (h_elect()) h_commit(&tmem_full[acc]);
            __syncwarp();
            acc_phase[acc] ^= 1u;
            acc ^= 1;
        }
    } else if (warp < kHtEpiWarp0) {
        // ===================== transform: z = relu(A x + C) in place =====================
        // four groups of two warps; group g owns units g, g + 4, ... (a unit's latency -- barrier, LDS, FMA, STS, proxy
        // fence -- is several hundred cycles, so four units are kept in flight)
        const int tt = threadIdx.x - kHtXformWarp0 * 32;       // 0..255
        const int g = tt >> 6, tg = tt & 63;
        const int j = tg & 7, r0 = tg >> 3;                    // channel chunk (8 ch), first row; rows r0 + 8 i
        const uint32_t off = (uint32_t)(r0 * 128 + ((j ^ r0) << 4));
        const int my_tiles = (tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
        const int units = my_tiles * kNumStems;
        const float zs = (SPLIT && p.z_sc) ? p.z_sc->mul : 1.f;
        float zmax = 0.f;
        for (int n = g; n < units; n += kHtGroups4) {
            const int ti = n / kNumStems, s = n - ti * kNumStems;
            const int t = (int)blockIdx.x + ti * (int)gridDim.x;
            const int b = t / p.tiles_per_img;
            const int stage = n % kHtStages;
            const uint32_t phase = (uint32_t)(n / kHtStages) & 1u;
            const float* ca = p.coefA + (long long)b * kStemTot + s * kStemC + j * 8;
            const float* cb = p.coefB + (long long)b * kStemTot + s * kStemC + j * 8;
            const float4 a0 = __ldg(reinterpret_cast<const float4*>(ca));
            const float4 a1 = __ldg(reinterpret_cast<const float4*>(ca) + 1);
            const float4 c0 = __ldg(reinterpret_cast<const float4*>(cb));
            const float4 c1 = __ldg(reinterpret_cast<const float4*>(cb) + 1);
            hbar_wait(&full[stage], phase, p.error_flag, 24);
            if (SPLIT) {
                // fp32 in (box = j / 4, 16-byte chunks 2 (j % 4) and 2 (j % 4) + 1 of the 128-byte row), fp16 hi / lo out in place
                const uint32_t sbase = h_u32(smem_a) + (uint32_t)(stage * kHtTileBytes);
                const uint32_t in0 = sbase + (uint32_t)((j >> 2) * 128 * 128 + r0 * 128 + (((2 * (j & 3)) ^ r0) << 4));
                const uint32_t in1 = sbase + (uint32_t)((j >> 2) * 128 * 128 + r0 * 128 + (((2 * (j & 3) + 1) ^ r0) << 4));
                const uint32_t outh = sbase + off, outl = sbase + 128u * 128u + off;
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    uint4 x0[4], x1[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        x0[i] = h_lds128(in0 + (uint32_t)((h * 4 + i) * 1024));
                        x1[i] = h_lds128(in1 + (uint32_t)((h * 4 + i) * 1024));
                    }
                    __syncwarp();                 // every lane that reads these pixel rows has read them: they may be overwritten
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float z[8];
                        z[0] = fmaxf(fmaf(a0.x, __uint_as_float(x0[i].x), c0.x), 0.f) * zs;
                        z[1] = fmaxf(fmaf(a0.y, __uint_as_float(x0[i].y), c0.y), 0.f) * zs;
                        z[2] = fmaxf(fmaf(a0.z, __uint_as_float(x0[i].z), c0.z), 0.f) * zs;
                        z[3] = fmaxf(fmaf(a0.w, __uint_as_float(x0[i].w), c0.w), 0.f) * zs;
                        z[4] = fmaxf(fmaf(a1.x, __uint_as_float(x1[i].x), c1.x), 0.f) * zs;
                        z[5] = fmaxf(fmaf(a1.y, __uint_as_float(x1[i].y), c1.y), 0.f) * zs;
                        z[6] = fmaxf(fmaf(a1.z, __uint_as_float(x1[i].z), c1.z), 0.f) * zs;
                        z[7] = fmaxf(fmaf(a1.w, __uint_as_float(x1[i].w), c1.w), 0.f) * zs;
#pragma unroll
                        for (int e = 0; e < 8; ++e) zmax = fmaxf(zmax, z[e]);           // z >= 0 after the ReLU
                        uint4 oh, ol;
                        tcepi::split_f16x2(z[0], z[1], oh.x, ol.x);
                        tcepi::split_f16x2(z[2], z[3], oh.y, ol.y);
                        tcepi::split_f16x2(z[4], z[5], oh.z, ol.z);
                        tcepi::split_f16x2(z[6], z[7], oh.w, ol.w);
                        h_sts128(outh + (uint32_t)((h * 4 + i) * 1024), oh);
                        h_sts128(outl + (uint32_t)((h * 4 + i) * 1024), ol);
                    }
                    __syncwarp();
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                hbar_arrive(&ready[stage]);
                continue;
            }
            const uint32_t base = h_u32(smem_a) + (uint32_t)(stage * kHtTileBytes) + off;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint4 v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = h_lds128(base + (uint32_t)((h * 8 + i) * 1024));
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    uint4 o;
                    o.x = relu_pack(fmaf(a0.x, __uint_as_float(v[i].x << 16), c0.x), fmaf(a0.y, __uint_as_float(v[i].x & 0xffff0000u), c0.y));
                    o.y = relu_pack(fmaf(a0.z, __uint_as_float(v[i].y << 16), c0.z), fmaf(a0.w, __uint_as_float(v[i].y & 0xffff0000u), c0.w));
                    o.z = relu_pack(fmaf(a1.x, __uint_as_float(v[i].z << 16), c1.x), fmaf(a1.y, __uint_as_float(v[i].z & 0xffff0000u), c1.y));
                    o.w = relu_pack(fmaf(a1.z, __uint_as_float(v[i].w << 16), c1.z), fmaf(a1.w, __uint_as_float(v[i].w & 0xffff0000u), c1.w));
                    h_sts128(base + (uint32_t)((h * 8 + i) * 1024), o);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> async proxy (tcgen05.mma reads)
            hbar_arrive(&ready[stage]);
        }
        if (SPLIT) tcepi::publish_amax(p.z_amax, zmax);
    } else {
        // ===================== epilogue (TMEM lane quarter = warp & 3) =====================
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t bs_addr = h_u32(bs);
        int acc = 0;
        uint32_t acc_phase[2] = {0u, 0u};
        for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
            const int b = t / p.tiles_per_img, pix = (t % p.tiles_per_img) * 128 + row;
            const bool valid = pix < p.HW;
            hbar_wait(&tmem_full[acc], acc_phase[acc], p.error_flag, 25);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kHtAccStride);
            {
                uint32_t v0[16], v1[16], v2[16], v3[16];
                tcepi::tmem_ld16_nowait(t_row + 0, v0);
                tcepi::tmem_ld16_nowait(t_row + 16, v1);
                tcepi::tmem_ld16_nowait(t_row + 32, v2);
                tcepi::tmem_ld16_nowait(t_row + 48, v3);
                tcepi::tmem_wait_ld();
                if (valid) { epi_group<0, SPLIT>(v0, bs_addr, p, b, pix); epi_group<1, SPLIT>(v1, bs_addr, p, b, pix); epi_group<2, SPLIT>(v2, bs_addr, p, b, pix); epi_group<3, SPLIT>(v3, bs_addr, p, b, pix); }
            }
            {
                uint32_t v0[16], v1[16], v2[16], v3[16];
                tcepi::tmem_ld16_nowait(t_row + 64, v0);
                tcepi::tmem_ld16_nowait(t_row + 80, v1);
                tcepi::tmem_ld16_nowait(t_row + 96, v2);
                tcepi::tmem_ld16_nowait(t_row + 112, v3);
                tcepi::tmem_wait_ld();
                if (valid) { epi_group<4, SPLIT>(v0, bs_addr, p, b, pix); epi_group<5, SPLIT>(v1, bs_addr, p, b, pix); epi_group<6, SPLIT>(v2, bs_addr, p, b, pix); epi_group<7, SPLIT>(v3, bs_addr, p, b, pix); }
            }
            {
                uint32_t v0[16], v1[16], v2[16];
                tcepi::tmem_ld16_nowait(t_row + 128, v0);
                tcepi::tmem_ld16_nowait(t_row + 144, v1);
                tcepi::tmem_ld16_nowait(t_row + 160, v2);
                tcepi::tmem_wait_ld();
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                hbar_arrive(&tmem_empty[acc]);                 // accumulator is in registers: release it before the stores
                if (valid) { epi_group<8, SPLIT>(v0, bs_addr, p, b, pix); epi_group<9, SPLIT>(v1, bs_addr, p, b, pix); epi_group<10, SPLIT>(v2, bs_addr, p, b, pix); }
            }
            acc_phase[acc] ^= 1u;
            acc ^= 1;
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
    }
}

typedef CUresult (*EncodeTiledFnH)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFnH g_encode_h = nullptr;
int g_num_sms_h = 148;
template <bool SPLIT> constexpr size_t ht_smem_bytes() {
    return 1024 + (size_t)HtCfg<SPLIT>::kStages * HtCfg<SPLIT>::kTileBytes + HtCfg<SPLIT>::kWBytes + 160 * 4 + 8 * (3 * HtCfg<SPLIT>::kStages + 4) + 16;
}

}  // namespace

struct HeadTcPlan {
    HeadTcParams p;
    int* d_err = nullptr;
    bool split = false;
};

void head_tc_init() {
    int dev = 0;
    MC_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    MC_CUDA(cudaGetDeviceProperties(&prop, dev));
    g_num_sms_h = std::max(1, prop.multiProcessorCount - reserved_sms());
    if (!g_encode_h) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        MC_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        MC_CHECK(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available in this driver");
        g_encode_h = reinterpret_cast<EncodeTiledFnH>(fn);
    }
    MC_CUDA(cudaFuncSetAttribute(head_apply_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ht_smem_bytes<false>()));
    MC_CUDA(cudaFuncSetAttribute(head_apply_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ht_smem_bytes<true>()));
}

// dt: storage type of the stem tensor (bf16 in the throughput mode, fp32 in the fp32-accurate tensor-core mode)
bool head_tc_supported(DType dt, int HW) { return (dt == DT_BF16 || dt == DT_F32) && HW >= 128; }

// w_host: the ten 1x1 weights [65][64] (only read for the fp32 stems: fp16 hi / lo pieces are packed here)
std::shared_ptr<HeadTcPlan> head_tc_prepare(Net& net, const void* stems, DType stems_dt, int max_batch, int HW, const std::vector<float>& w_host,
                                            int z_tensor) {
    auto plan = std::make_shared<HeadTcPlan>();
    HeadTcParams& p = plan->p;
    std::memset(&p, 0, sizeof(p));
    MC_CHECK(g_encode_h != nullptr, "head_tc_init has not been called");
    const bool split = stems_dt == DT_F32;
    plan->split = split;
    const cuuint64_t es = split ? 4 : 2;
    cuuint64_t dims[3] = {(cuuint64_t)kStemTot, (cuuint64_t)HW, (cuuint64_t)max_batch};
    cuuint64_t str[2] = {(cuuint64_t)kStemTot * es, (cuuint64_t)HW * kStemTot * es};
    cuuint32_t box[3] = {(cuuint32_t)(split ? 32 : kStemC), 128, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode_h(&p.map_x, split ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(stems), dims, str,
                            box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ") for the head stems");
    plan->d_err = (int*)net.arena.alloc(sizeof(int));
    p.error_flag = plan->d_err;
    p.HW = HW;
    p.tiles_per_img = (HW + 127) / 128;
    if (split) {
        MC_CHECK((int)w_host.size() == kNumOut * kStemC, "head_tc: 1x1 weights");
        const std::vector<int> ew = split_weight_exponents(w_host, kNumOut);
        // padded row layout of the accumulator: stem s owns rows [col[s], col[s] + npad[s]), its outputs o0[s] .. o1[s] come first
        const int col[kNumStems] = {0, 16, 32, 48, 80, 96, 112, 128, 144};
        const int o0[kNumStems] = {0, 12, 14, 18, 3, 16, 36, 39, 41}, o1[kNumStems] = {3, 14, 16, 36, 12, 18, 39, 41, 65};
        std::vector<uint16_t> ws((size_t)2 * kHtCols * kStemC, 0);
        for (int s = 0; s < kNumStems; ++s)
            for (int o = o0[s]; o < o1[s]; ++o)
                for (int k = 0; k < kStemC; ++k)
                    for (int piece = 0; piece < 2; ++piece)
                        ws[((size_t)piece * kHtCols + col[s] + (o - o0[s])) * kStemC + k] = split_weight_piece(w_host[(size_t)o * kStemC + k], ew[o], piece == 1);
        std::vector<float> osc(80, 1.f);
        for (int o = 0; o < kNumOut; ++o) osc[o] = std::ldexp(1.f, -ew[o]);
        uint16_t* dws = (uint16_t*)net.arena.alloc(sizeof(uint16_t) * ws.size());
        float* dos = (float*)net.arena.alloc(sizeof(float) * osc.size());
        MC_CUDA(cudaMemcpy(dws, ws.data(), sizeof(uint16_t) * ws.size(), cudaMemcpyHostToDevice));
        MC_CUDA(cudaMemcpy(dos, osc.data(), sizeof(float) * osc.size(), cudaMemcpyHostToDevice));
        p.w_split = dws;
        p.oscale = dos;
        MC_CHECK(z_tensor >= 0, "head_tc: the fp32-accurate mode needs the head.z scale slot");
        p.z_sc = net.act_scale(z_tensor);
        p.z_amax = net.act_amax(z_tensor);
    }
    return plan;
}

void launch_head_apply_tc(const HeadTcPlan& plan, const HeadApplyParams& ap, cudaStream_t st) {
    HeadTcParams p = plan.p;
    p.coefA = ap.coefA; p.coefB = ap.coefB; p.w = ap.w; p.bias = ap.bias;
    for (int i = 0; i < kNumPred; ++i) p.out[i] = ap.out[i];
    p.B = ap.B;
    MC_CHECK(ap.HW == p.HW, "head_tc: feature size differs from the prepared plan");
    const int tiles = p.B * p.tiles_per_img;
    const int grid = tiles < g_num_sms_h ? tiles : g_num_sms_h;
    if (plan.split) launch_k(head_apply_tc_kernel<true>, dim3(grid), dim3(kHtThreads), ht_smem_bytes<true>(), st, p);
    else launch_k(head_apply_tc_kernel<false>, dim3(grid), dim3(kHtThreads), ht_smem_bytes<false>(), st, p);
}

}  // namespace mc
