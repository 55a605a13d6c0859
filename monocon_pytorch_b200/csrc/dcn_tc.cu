// Fused modulated deformable 3x3 convolution on tcgen05 (sm_100a): the DCN variant of the IDAUp blocks without
// the 9 * Cin column tensor in HBM.  Same arithmetic as csrc/dcn.cu + the 1x1 tensor-core layer (operator semantics of
// torchvision.ops.deform_conv2d, 3x3 / stride 1 / pad 1 / one offset group), different data path:
//
//   GEMM view:  D[pixel, cout] = sum_{tap, chunk} A_{tap,chunk}[pixel, 64] * W_{tap,chunk}[cout, 64]^T
//   A_{tap,chunk}[pixel, c] = sigmoid(logit_tap[pixel]) * bilinear(x_c, pixel + tap + offset_tap[pixel])   is PRODUCED IN THE
//   KERNEL: eight gather warps sample the NHWC activation (four 16-byte corner loads per 8 channels and stored plane, fp32 blend
//   in the operator's summation order) and write the 128 x 64 operand tile straight into the 128-byte-swizzled shared-memory
//   stage the MMA reads (generic-proxy stores -> fence.proxy.async -> mbarrier), exactly where a TMA box would have landed.
//
//   warp 0      : TMA producer of the weight tiles (one elected lane), ring of `stages` stages
//   warp 1      : tcgen05.mma issuer (one elected lane), accumulators double-buffered in TMEM (2 x 256 columns)
//   warps 2..5  : epilogue: tcgen05.ld -> folded BN (+ReLU) -> bf16 / fp16 hi + lo planes -> global (tc_epilogue.cuh)
//   warps 6..21 : gather producers: a warp owns 8 pixel rows of the tile; lane = (row of a group of four, 16-byte chunk), so
//                 every warp-wide load reads whole 128-byte lines
//   persistent CTAs, tile = 128 consecutive pixels of the flattened (image, row, column) index (the NHWC output is contiguous in
//   it, so ragged rows cost nothing), every CTA owns a contiguous range of tiles.
//
// fp32-accurate mode (DT_SPLIT): the sampled value is formed from hi + lo of the four corners in fp32 and re-split into an fp16
// hi tile and a lo tile at the SOURCE's exponent (|sample| <= max |x|: no new scale); a K-block is then the three products
// hi x w_lo, lo x w_hi, hi x w_hi of conv_tc3.cu issued back to back on the two tiles (gathering once instead of three times
// is worth more than issuing all cross terms first).
//
// Bounds per K-block (128 pixels x 64 channels of one tap): the gather moves 4 corners x 16 KB (x2 planes) through L1 -- 512
// (1024) clocks at 128 B/clk -- against 4 (12) MMAs of N = Cout; no HBM traffic beyond x, the offsets and the output.
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "engine.h"
#include "tc_epilogue.cuh"

namespace mc {

namespace {

constexpr int kTileM = 128;
constexpr int kGatherWarps = 16;
constexpr int kThreads = 64 + 128 + 32 * kGatherWarps;      // 704
constexpr int kMaxStages = 6;
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;
constexpr long long kSpinLimit = 4000000000LL;

struct DcnTcParams {
    CUtensorMap map_b;          // packed weights: rows = [K-block][piece][cout], 64 K elements per row
    const void* src[2];
    int srcC[2];
    long long src_plane[2];
    int nsrc;
    const void* off;            // NHWC offsets / mask logits, offC channels per pixel
    int offC;
    long long off_plane;
    const ActScale* off_sc;
    int H, W, B, Cin, Cout;
    int nchunks;                // Cin / 64
    int stages;
    int a_stage, b_stage;       // bytes per stage
    int b_piece;                // bytes of one weight piece tile (cout x 128)
    const float* scale;
    const float* shift;
    void* dst;
    int relu;
    int f16;
    int mask_logits;            // the modulation channels hold logits (the DCNv2 pack): apply the sigmoid here
    long long dst_plane;
    const ActScale* in_sc;
    const ActScale* out_sc;
    unsigned* amax;
    int* error_flag;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// SLEEP: nanoseconds between polls.  This kernel is bound by the instruction issue of its gather warps (ncu: a fifth of all issued
// instructions were the try_wait / clock / branch loops of the six warps that wait for them), so every other role backs off.
template <int SLEEP = 0>
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* error_flag, int code) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (SLEEP > 0) __nanosleep(SLEEP);
        if (clock64() - t0 > kSpinLimit) {           // a stuck pipeline traps instead of hanging the GPU
            if (error_flag) atomicExch(error_flag, code);
            __threadfence_system();
            asm volatile("trap;");
        }
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tma_load_2d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) { asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory"); }
// K-major operands, 128-byte rows, SWIZZLE_128B; descriptor halves as in conv_tc.cu (lo = LBO(1) | address >> 4)
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint32_t alo, uint32_t blo, uint32_t hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
        ::"r"(tmem_d), "r"(alo), "r"(blo), "r"(hi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ldg128(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// two floats -> (hi pair, lo pair) of fp16 without the saturation of tcepi::split_f16x2: a sample is a convex combination of
// stored values times a mask in (0, 1), so it cannot leave the source tensor's fp16 range
__device__ __forceinline__ void split_f16x2_nosat(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// packed fp32 pairs (sm_100: add / mul / fma .f32x2 on 64-bit register pairs): the blend of two channels per instruction -- the
// gather is bound by instruction issue, and the per-lane results are bit-identical to the scalar sequence
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
// one 32-bit word of two stored channels -> packed fp32 pair
template <bool F16> __device__ __forceinline__ f32x2 unpack_pair(uint32_t w) {
    if (F16) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
        return pk2(f.x, f.y);
    }
    return pk2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}

// eight consecutive channels of one stored plane -> fp32
template <bool F16> __device__ __forceinline__ void unpack8(const uint4& r, float (&v)[8]) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float2 f;
        if (F16) f = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
        else f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
        v[2 * j] = f.x; v[2 * j + 1] = f.y;
    }
}

template <int OM>
__global__ void __launch_bounds__(kThreads, 1) dcn_tc_kernel(const __grid_constant__ DcnTcParams p) {
    constexpr bool kSplit = OM == tcepi::OM_SPLIT;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;                                                  // [stages][1 | 2 planes][128 x 128 B]
    uint8_t* smem_b = smem + (size_t)p.stages * p.a_stage;                   // [stages][1 | 2 pieces][cout x 128 B]
    float* s_scale = reinterpret_cast<float*>(smem_b + (size_t)p.stages * p.b_stage);
    float* s_shift = s_scale + p.Cout;
    uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(s_shift + p.Cout) + 15) & ~uintptr_t(15));
    uint64_t* full_a = bars;                          // [kMaxStages] gather threads arrive
    uint64_t* full_b = bars + kMaxStages;             // [kMaxStages] TMA bytes
    uint64_t* empty_bar = bars + 2 * kMaxStages;      // [kMaxStages] MMA commit
    uint64_t* tmem_full = bars + 3 * kMaxStages;      // [2]
    uint64_t* tmem_empty = bars + 3 * kMaxStages + 2; // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 3 * kMaxStages + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long npix = (long long)p.B * p.H * p.W;
    const int total_tiles = (int)((npix + kTileM - 1) / kTileM);
    const int per_cta = total_tiles / (int)gridDim.x, rem_cta = total_tiles % (int)gridDim.x;
    const int t_begin = (int)blockIdx.x * per_cta + min((int)blockIdx.x, rem_cta);
    const int t_end = t_begin + per_cta + ((int)blockIdx.x < rem_cta ? 1 : 0);
    const int nkb = 9 * p.nchunks;

    {
        const float in_inv = (kSplit && p.in_sc) ? p.in_sc->inv : 1.f, out_mul = (kSplit && p.out_sc) ? p.out_sc->mul : 1.f;
        for (int i = threadIdx.x; i < p.Cout; i += kThreads) {
            s_scale[i] = p.scale[i] * in_inv * out_mul;
            s_shift[i] = p.shift[i] * out_mul;
        }
    }
    if (warp == 0 && lane == 0) prefetch_tmap(&p.map_b);
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full_a[s], 32 * kGatherWarps); mbar_init(&full_b[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 128); }
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_sync();

    if (warp == 0) {
        // ===================== TMA producer: weight tiles =====================
        int stage = 0;
        uint32_t phase = 0;
        for (int t = t_begin; t < t_end; ++t)
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait<200>(&empty_bar[stage], phase ^ 1u, p.error_flag, 1);
                if (elect_one()) {
                    mbar_expect_tx(&full_b[stage], (uint32_t)p.b_stage);
                    uint8_t* b_dst = smem_b + (size_t)stage * p.b_stage;
                    if (kSplit) {
                        tma_load_2d(b_dst, &p.map_b, &full_b[stage], 0, (kb * 2) * p.Cout);                 // w_lo piece
                        tma_load_2d(b_dst + p.b_piece, &p.map_b, &full_b[stage], 0, (kb * 2 + 1) * p.Cout);   // w_hi piece
                    } else {
                        tma_load_2d(b_dst, &p.map_b, &full_b[stage], 0, kb * p.Cout);
                    }
                }
                __syncwarp();
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t fmt = kSplit ? 0u : 1u;         // fp16 / bf16 operands
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.Cout >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
        const uint32_t hi = (uint32_t)((8 * 128) >> 4) | (1u << 14) | (2u << 29);
        const uint32_t a_base16 = (1u << 16) | ((smem_u32(smem_a) & 0x3FFFF) >> 4);
        const uint32_t b_base16 = (1u << 16) | ((smem_u32(smem_b) & 0x3FFFF) >> 4);
        const uint32_t a_stage16 = (uint32_t)p.a_stage >> 4, b_stage16 = (uint32_t)p.b_stage >> 4;
        const uint32_t a_plane16 = (uint32_t)(kTileM * 128) >> 4, b_piece16 = (uint32_t)p.b_piece >> 4;
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase[2] = {0u, 0u};
        for (int t = t_begin; t < t_end; ++t) {
            mbar_wait<100>(&tmem_empty[acc], acc_phase[acc] ^ 1u, p.error_flag, 2);
            tc_fence_after();
            const uint32_t d0 = tmem_base + (uint32_t)(acc * kAccStride);
            uint32_t accumulate = 0;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait<50>(&full_b[stage], phase, p.error_flag, 3);
                mbar_wait<50>(&full_a[stage], phase, p.error_flag, 5);
                tc_fence_after();
                const uint32_t alo = a_base16 + (uint32_t)stage * a_stage16;
                const uint32_t blo = b_base16 + (uint32_t)stage * b_stage16;
                if (elect_one()) {
                    if (kSplit) {
                        // hi x w_lo, lo x w_hi, hi x w_hi
#pragma unroll
                        for (int k = 0; k < 4; ++k) { umma_f16(d0, alo + 2u * k, blo + 2u * k, hi, idesc, accumulate); accumulate = 1; }
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_f16(d0, alo + a_plane16 + 2u * k, blo + b_piece16 + 2u * k, hi, idesc, 1u);
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_f16(d0, alo + 2u * k, blo + b_piece16 + 2u * k, hi, idesc, 1u);
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) { umma_f16(d0, alo + 2u * k, blo + 2u * k, hi, idesc, accumulate); accumulate = 1; }
                    }
                    umma_commit(&empty_bar[stage]);
                }
                __syncwarp();
                accumulate = 1;
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
            if (elect_one()) umma_commit(&tmem_full[acc]);
            __syncwarp();
            acc_phase[acc] ^= 1u;
            acc ^= 1;
        }
    } else if (warp < 6) {
        // ===================== epilogue (warps 2..5 = TMEM lane quarters 2, 3, 0, 1) =====================
        const int q = warp & 3;
        const int row = q * 32 + lane;
        int acc = 0;
        uint32_t acc_phase[2] = {0u, 0u};
        constexpr int EB = tcepi::ElemBytes<OM>::value;
        tcepi::SplitEpi se;
        se.dst_plane = p.dst_plane; se.res_plane = 0; se.res_mul = 1.f;
        float amax = 0.f;
        for (int t = t_begin; t < t_end; ++t) {
            const long long pix = (long long)t * kTileM + row;
            const bool valid = pix < npix;
            char* dst = reinterpret_cast<char*>(p.dst) + pix * p.Cout * EB;
            mbar_wait<500>(&tmem_full[acc], acc_phase[acc], p.error_flag, 4);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kAccStride);
            tcepi::drain_row<OM, false>(t_row, p.Cout, s_scale, s_shift, nullptr, dst, valid, p.relu != 0, se, amax);
            tc_fence_before();
            mbar_arrive(&tmem_empty[acc]);
            acc_phase[acc] ^= 1u;
            acc ^= 1;
        }
        if (kSplit) tcepi::publish_amax(p.amax, amax);
    } else {
        // ===================== gather producers (warps 6..21) =====================
        // Warp gw owns 8 rows (pixels) of the tile.  Loads: lane = (row within a group of four, 16-byte chunk of the 128-byte
        // K-block row), so one LDG.128 of the warp reads four whole 128-byte lines (a first version with lane = row read 32
        // lines per instruction and ran at a tenth of the L1 bound).  Sampling parameters: lane l computes those of row l & 7
        // once per tap; the lanes fetch the rows they load with shuffles.
        constexpr int RW = kTileM / kGatherWarps;           // rows per warp (8)
        constexpr int NIT = RW / 4;                        // groups of four rows per warp (2)
        const int gw = warp - 6, chunk8 = lane & 7, rsub = lane >> 3;
        const int prow = gw * RW + (lane & (RW - 1));      // the row whose parameters this lane computes
        const float oinv = (kSplit && p.off_sc) ? p.off_sc->inv : 1.f;
        const int H = p.H, W = p.W;
        int stage = 0;
        uint32_t phase = 0;
        // raw (dy, dx, modulation) of one tap of this lane's parameter row; the loads of tap k + 1 (or of the next tile's first tap)
        // are issued BEFORE the K-blocks of tap k are gathered, so their latency hides behind the gather (all warps reach a tap
        // boundary at about the same time: unhidden, it cost half the kernel)
        auto load_tap = [&](const char* offp, int k, float& dy, float& dx, float& mv) {
            if (kSplit) {
                const __half* o = reinterpret_cast<const __half*>(offp);
                dy = __half2float(o[2 * k]) + __half2float(o[2 * k + p.off_plane]);
                dx = __half2float(o[2 * k + 1]) + __half2float(o[2 * k + 1 + p.off_plane]);
                mv = __half2float(o[18 + k]) + __half2float(o[18 + k + p.off_plane]);
            } else {
                const bf16* o = reinterpret_cast<const bf16*>(offp);
                dy = __bfloat162float(o[2 * k]); dx = __bfloat162float(o[2 * k + 1]); mv = __bfloat162float(o[18 + k]);
            }
        };
        auto tile_offp = [&](int t) {
            const long long pix = (long long)t * kTileM + prow;
            return reinterpret_cast<const char*>(p.off) + (pix < npix ? pix : 0) * p.offC * 2;      // 2-byte elements in both tensor-core modes
        };
        float ndy = 0.f, ndx = 0.f, nmv = 0.f;
        if (t_begin < t_end) load_tap(tile_offp(t_begin), 0, ndy, ndx, nmv);
        for (int t = t_begin; t < t_end; ++t) {
            const long long pix = (long long)t * kTileM + prow;
            const bool valid = pix < npix;
            const long long pv = valid ? pix : 0;
            const int b = (int)(pv / ((long long)H * W));
            const int rem = (int)(pv - (long long)b * H * W);
            const int y = rem / W, x = rem - y * W;
            const char* offp = tile_offp(t);
            for (int k = 0; k < 9; ++k) {
                // ---- the tap's sampling position (deform_conv2d_kernel.cpp, bilinear_interpolate) ----
                const float dy = ndy * oinv, dx = ndx * oinv, mv = nmv * oinv;
                if (k + 1 < 9) load_tap(offp, k + 1, ndy, ndx, nmv);
                else if (t + 1 < t_end) load_tap(tile_offp(t + 1), 0, ndy, ndx, nmv);
                const float m = valid ? (p.mask_logits ? 1.f / (1.f + expf(-mv)) : mv) : 0.f;
                const int ti = k / 3, tj = k - 3 * ti;
                const float py = (float)(y - 1 + ti) + dy, px = (float)(x - 1 + tj) + dx;
                float a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
                int h0 = 0, w0 = 0;
                if (py > -1.f && py < (float)H && px > -1.f && px < (float)W) {
                    const float fl_h = floorf(py), fl_w = floorf(px);
                    h0 = (int)fl_h; w0 = (int)fl_w;
                    const float lh = py - fl_h, lw = px - fl_w, hh = 1.f - lh, hw = 1.f - lw;
                    const bool t_ok = h0 >= 0, b_ok = h0 + 1 <= H - 1, l_ok = w0 >= 0, r_ok = w0 + 1 <= W - 1;
                    a1 = (t_ok && l_ok) ? hh * hw : 0.f;
                    a2 = (t_ok && r_ok) ? hh * lw : 0.f;
                    a3 = (b_ok && l_ok) ? lh * hw : 0.f;
                    a4 = (b_ok && r_ok) ? lh * lw : 0.f;
                }
                const int ra = (b * H + min(max(h0, 0), H - 1)) * W, rb = (b * H + min(max(h0 + 1, 0), H - 1)) * W;
                const int xa = min(max(w0, 0), W - 1), xb = min(max(w0 + 1, 0), W - 1);
                // rows this lane loads: gw * RW + 4 * it + rsub -> parameters from lane 4 * it + rsub.  The operator multiplies the blended
                // sample by the mask; folding the mask into the four weights moves each product by at most one fp32 ulp
                a1 *= m; a2 *= m; a3 *= m; a4 *= m;
                unsigned q1[NIT], q2[NIT], q3[NIT], q4[NIT];
                float w1[NIT], w2[NIT], w3[NIT], w4[NIT];
#pragma unroll
                for (int it = 0; it < NIT; ++it) {
                    const int sl = 4 * it + rsub;
                    const int sra = __shfl_sync(0xffffffffu, ra, sl), srb = __shfl_sync(0xffffffffu, rb, sl);
                    const int sxa = __shfl_sync(0xffffffffu, xa, sl), sxb = __shfl_sync(0xffffffffu, xb, sl);
                    q1[it] = (unsigned)(sra + sxa); q2[it] = (unsigned)(sra + sxb); q3[it] = (unsigned)(srb + sxa); q4[it] = (unsigned)(srb + sxb);
                    w1[it] = __shfl_sync(0xffffffffu, a1, sl); w2[it] = __shfl_sync(0xffffffffu, a2, sl);
                    w3[it] = __shfl_sync(0xffffffffu, a3, sl); w4[it] = __shfl_sync(0xffffffffu, a4, sl);
                }
                for (int ch = 0; ch < p.nchunks; ++ch) {
                    const int c0 = ch * 64;                        // first channel of the K-block in the concatenated input
                    const bool second = c0 >= p.srcC[0];
                    const char* sp = reinterpret_cast<const char*>(second ? p.src[1] : p.src[0]);
                    const int C = second ? p.srcC[1] : p.srcC[0];
                    const unsigned c = (unsigned)((second ? c0 - p.srcC[0] : c0) + 8 * chunk8);
                    const long long plane_b = (second ? p.src_plane[1] : p.src_plane[0]) * 2;
                    mbar_wait(&empty_bar[stage], phase ^ 1u, p.error_flag, 6);
                    const uint32_t a_hi = smem_u32(smem_a + (size_t)stage * p.a_stage);
                    // all loads of a pass in flight before the first use: bf16 -- one pass over the thread's two rows (8 x 16 B);
                    // fp16 planes -- one row per pass (8 x 16 B)
                    constexpr int NR = kSplit ? 1 : NIT;           // rows per pass
#pragma unroll
                    for (int pass = 0; pass < NIT / NR; ++pass) {
                        uint4 rh[4][NR];
                        uint4 rl[kSplit ? 4 : 1][kSplit ? NR : 1];
#pragma unroll
                        for (int j = 0; j < NR; ++j) {
                            const int it = pass * NR + j;
                            const char* g1 = sp + ((size_t)q1[it] * C + c) * 2;
                            const char* g2 = sp + ((size_t)q2[it] * C + c) * 2;
                            const char* g3 = sp + ((size_t)q3[it] * C + c) * 2;
                            const char* g4 = sp + ((size_t)q4[it] * C + c) * 2;
                            rh[0][j] = ldg128(g1); rh[1][j] = ldg128(g2); rh[2][j] = ldg128(g3); rh[3][j] = ldg128(g4);
                            if (kSplit) {
                                rl[0][j] = ldg128(g1 + plane_b); rl[1][j] = ldg128(g2 + plane_b);
                                rl[kSplit ? 2 : 0][j] = ldg128(g3 + plane_b); rl[kSplit ? 3 : 0][j] = ldg128(g4 + plane_b);
                            }
                        }
#pragma unroll
                        for (int j = 0; j < NR; ++j) {
                            const int it = pass * NR + j;
                            float o[8];
                            if (!kSplit) {
                                // bf16: packed fp32 pairs (fma.rn.f32x2), two channels per instruction: 3.14 -> 2.81 ms on the 12 blocks
                                const uint32_t* c1w = reinterpret_cast<const uint32_t*>(&rh[0][j]);
                                const uint32_t* c2w = reinterpret_cast<const uint32_t*>(&rh[1][j]);
                                const uint32_t* c3w = reinterpret_cast<const uint32_t*>(&rh[2][j]);
                                const uint32_t* c4w = reinterpret_cast<const uint32_t*>(&rh[3][j]);
                                const f32x2 W1 = pk2(w1[it], w1[it]), W2 = pk2(w2[it], w2[it]), W3 = pk2(w3[it], w3[it]), W4 = pk2(w4[it], w4[it]);
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const f32x2 x1 = unpack_pair<false>(c1w[e]), x2 = unpack_pair<false>(c2w[e]);
                                    const f32x2 x3 = unpack_pair<false>(c3w[e]), x4 = unpack_pair<false>(c4w[e]);
                                    // ((w1 v1 + w2 v2) + w3 v3) + w4 v4: the operator's summation order
                                    upk2(fma2(W4, x4, fma2(W3, x3, fma2(W2, x2, mul2(W1, x1)))), o[2 * e], o[2 * e + 1]);
                                }
                            } else {
                                // fp16 planes: scalar (the packed form needs the hi + lo pairs of four corners live at once and spills
                                // under the 80-register cap of the 704-thread block: 5.8 -> 7.3 ms)
                                float v1[8], v2[8], v3[8], v4[8], l1[8], l2[8], l3[8], l4[8];
                                unpack8<true>(rh[0][j], v1); unpack8<true>(rh[1][j], v2);
                                unpack8<true>(rh[2][j], v3); unpack8<true>(rh[3][j], v4);
                                unpack8<true>(rl[0][kSplit ? j : 0], l1); unpack8<true>(rl[kSplit ? 1 : 0][kSplit ? j : 0], l2);
                                unpack8<true>(rl[kSplit ? 2 : 0][kSplit ? j : 0], l3); unpack8<true>(rl[kSplit ? 3 : 0][kSplit ? j : 0], l4);
#pragma unroll
                                for (int e = 0; e < 8; ++e)      // hi + lo is exact in fp32; then the operator's summation order
                                    o[e] = w1[it] * (v1[e] + l1[e]) + w2[it] * (v2[e] + l2[e]) + w3[it] * (v3[e] + l3[e]) + w4[it] * (v4[e] + l4[e]);
                            }
                            const int row = gw * RW + 4 * it + rsub;
                            const uint32_t dsta = a_hi + (uint32_t)row * 128u + (((uint32_t)chunk8 ^ (uint32_t)(row & 7)) << 4);   // SWIZZLE_128B
                            if (kSplit) {
                                uint32_t oh[4], ol[4];
#pragma unroll
                                for (int e = 0; e < 4; ++e) split_f16x2_nosat(o[2 * e], o[2 * e + 1], oh[e], ol[e]);
                                sts128(dsta, oh[0], oh[1], oh[2], oh[3]);
                                sts128(dsta + (uint32_t)(kTileM * 128), ol[0], ol[1], ol[2], ol[3]);
                            } else {
                                sts128(dsta, tcepi::pack_bf16(o[0], o[1]), tcepi::pack_bf16(o[2], o[3]), tcepi::pack_bf16(o[4], o[5]),
                                       tcepi::pack_bf16(o[6], o[7]));
                            }
                        }
                    }
                    fence_proxy_async();                           // generic-proxy stores -> visible to the tensor core's async proxy
                    mbar_arrive(&full_a[stage]);
                    if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
int g_num_sms = 148;
int g_max_smem = 0;

}  // namespace

struct DcnTcPlan {
    DcnTcParams p;
    void* d_w = nullptr;
    int* d_err = nullptr;
    int om = tcepi::OM_BF16;
    size_t smem_bytes = 0;
};

void dcn_tc_init() {
    int dev = 0;
    MC_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    MC_CUDA(cudaGetDeviceProperties(&prop, dev));
    g_num_sms = std::max(1, prop.multiProcessorCount - reserved_sms());
    g_max_smem = (int)prop.sharedMemPerBlockOptin;
    if (!g_encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        MC_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        MC_CHECK(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available in this driver");
        g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    MC_CUDA(cudaFuncSetAttribute(dcn_tc_kernel<tcepi::OM_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem));
    MC_CUDA(cudaFuncSetAttribute(dcn_tc_kernel<tcepi::OM_SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem));
}

// the fused kernel covers a deformable layer when its sources are tensor-core storage with whole 64-channel K-blocks
bool dcn_tc_supported(const Net& net, const ConvLayer& L) {
    if (const char* e = std::getenv("MC_DCN_FUSE")) if (e[0] == '0') return false;
    if (net.dt != DT_BF16 && net.dt != DT_SPLIT) return false;
    if (L.dcn_off < 0 || L.k != 3 || L.stride != 1 || L.pad != 1 || L.residual >= 0 || L.dst_override) return false;
    if (L.src.size() > 2 || L.cout % 16 != 0 || L.cout > 256) return false;
    for (int s : L.src)
        if (net.tensors[s].dt != net.dt || net.tensors[s].C % 64 != 0 || net.tensors[s].Wp != net.tensors[s].W) return false;
    const TensorInfo& o = net.tensors[L.dcn_off];
    if (o.dt != net.dt || o.C < 27 || o.Wp != o.W) return false;
    return net.tensors[L.dst].dt == net.dt;
}

void dcn_tc_prepare(Net& net, ConvLayer& L, const std::vector<float>& w_oihw) {
    MC_CHECK(dcn_tc_supported(net, L), "dcn_tc: layer not supported: " + L.name);
    auto plan = std::make_shared<DcnTcPlan>();
    DcnTcParams& p = plan->p;
    std::memset(&p, 0, sizeof(p));
    const bool split = net.dt == DT_SPLIT;
    const TensorInfo& s0 = net.tensors[L.src[0]];
    const TensorInfo& d = net.tensors[L.dst];
    const TensorInfo& o = net.tensors[L.dcn_off];
    MC_CHECK((long long)net.max_batch * s0.H * s0.W * L.cin_store < (1ll << 31), "dcn_tc: tensor too large for 32-bit pixel offsets");
    p.nsrc = (int)L.src.size();
    for (int s = 0; s < 2; ++s) {
        const TensorInfo& t = net.tensors[L.src[std::min(s, p.nsrc - 1)]];
        p.src[s] = t.ptr; p.srcC[s] = t.C; p.src_plane[s] = t.plane;
    }
    if (p.nsrc == 1) p.srcC[1] = 0;
    p.off = o.ptr; p.offC = o.C; p.off_plane = o.plane; p.off_sc = split ? net.act_scale(L.dcn_off) : nullptr;
    p.H = s0.H; p.W = s0.W; p.B = net.max_batch; p.Cin = L.cin_store; p.Cout = L.cout;
    p.nchunks = L.cin_store / 64;
    // weights [K-block = tap * nchunks + chunk][piece][cout][64]; fp32-accurate mode: pieces (lo, hi) of w * 2^ew[cout]
    const std::vector<int> ew = split ? split_weight_exponents(w_oihw, L.cout) : std::vector<int>();
    const int pieces = split ? 2 : 1;
    std::vector<uint16_t> w((size_t)9 * p.nchunks * pieces * L.cout * 64);
    size_t at = 0;
    for (int t = 0; t < 9; ++t)
        for (int ch = 0; ch < p.nchunks; ++ch)
            for (int pc = 0; pc < pieces; ++pc)
                for (int oc = 0; oc < L.cout; ++oc)
                    for (int kk = 0; kk < 64; ++kk) {
                        const float v = w_oihw[((size_t)oc * L.cin + ch * 64 + kk) * 9 + t];
                        w[at++] = split ? split_weight_piece(v, ew[oc], pc == 0) : bf16_bits(v);
                    }
    plan->d_w = net.arena.alloc(sizeof(uint16_t) * w.size());
    MC_CUDA(cudaMemcpy(plan->d_w, w.data(), sizeof(uint16_t) * w.size(), cudaMemcpyHostToDevice));
    L.w_packed = plan->d_w;
    plan->d_err = (int*)net.arena.alloc(sizeof(int));
    p.error_flag = plan->d_err;
    {
        cuuint64_t dims[2] = {64, (cuuint64_t)9 * p.nchunks * pieces * L.cout};
        cuuint64_t str[1] = {128};
        cuuint32_t box[2] = {64, (cuuint32_t)L.cout};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = g_encode(&p.map_b, split ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, plan->d_w, dims, str, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) throw Error("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ") for " + L.name + " (deformable weights)");
    }
    p.b_piece = (L.cout * 128 + 1023) / 1024 * 1024;
    p.a_stage = kTileM * 128 * pieces;
    p.b_stage = p.b_piece * pieces;
    const size_t fixed = 1024 + sizeof(float) * 2 * L.cout + 16 + 8 * (3 * kMaxStages + 4) + 16;
    int stages = (int)(((size_t)g_max_smem - fixed) / (size_t)(p.a_stage + p.b_stage));
    stages = std::min(stages, kMaxStages);
    if (const char* e = std::getenv("MC_DCN_STAGES")) stages = std::max(2, std::min(stages, std::atoi(e)));      // A/B knob: fewer stages leave more L1 to the gather
    MC_CHECK(stages >= 2, "dcn_tc: not enough shared memory for 2 stages: " + L.name);
    p.stages = stages;
    plan->smem_bytes = fixed + (size_t)stages * (p.a_stage + p.b_stage);
    p.scale = split ? net.upload_split_scale(L, ew) : L.scale;
    p.shift = L.shift;
    p.dst = d.ptr;
    p.relu = L.relu ? 1 : 0;
    p.f16 = split ? 1 : 0;
    p.mask_logits = L.dcn_mask_logits ? 1 : 0;
    plan->om = split ? tcepi::OM_SPLIT : tcepi::OM_BF16;
    if (split) {
        p.in_sc = net.act_scale(L.src[0]);
        p.out_sc = net.act_scale(L.dst); p.amax = net.act_amax(L.dst); p.dst_plane = d.plane;
    }
    L.dcn = plan;
}

void dcn_tc_launch(const Net& net, const ConvLayer& L, int B, cudaStream_t st) {
    MC_CHECK(L.dcn != nullptr, "dcn_tc layer not prepared: " + L.name);
    DcnTcParams p = L.dcn->p;
    p.B = B;
    // concatenated sources meet in the offset convolution's K dimension, so they share one exponent (mc_calibrate_scales)
    MC_CHECK(net.dt != DT_SPLIT || L.src.size() == 1 || net.act_exp.empty() || net.act_exp[L.src[0]] == net.act_exp[L.src[1]], "dcn_tc: sources with different scales");
    const long long npix = (long long)B * p.H * p.W;
    const int tiles = (int)((npix + kTileM - 1) / kTileM);
    const int grid = std::min(tiles, g_num_sms);
    if (L.dcn->om == tcepi::OM_SPLIT) launch_k(dcn_tc_kernel<tcepi::OM_SPLIT>, dim3(grid), dim3(kThreads), L.dcn->smem_bytes, st, p);
    else launch_k(dcn_tc_kernel<tcepi::OM_BF16>, dim3(grid), dim3(kThreads), L.dcn->smem_bytes, st, p);
}

}  // namespace mc
