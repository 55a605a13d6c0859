// tcgen05 implicit-GEMM convolution for sm_100a (B200): NHWC bf16 activations, fp32 accumulation in
// tensor memory, fused folded-BN / bias / residual / ReLU epilogue.
//
//   GEMM view:  D[pixel, cout] = sum_{kb} A_kb[pixel, BK] * W_kb[cout, BK]^T
//   a "K-block" kb = one filter tap (r,s) x one BK-channel slice of one (concat-free) source tensor.
//   A_kb is fetched by ONE 5-D TMA box per K-block straight from the NHWC activation: the box is the
//   128-pixel output tile shifted by the tap offset, out-of-bounds pixels are zero-filled by TMA (= the
//   convolution's zero padding).  Stride-2 convolutions use a space-to-depth *view* of the same memory
//   (dims: [pw*C + c, W/2, ph, H/2, N]) so that a tap is again a dense box.  The 7x7 / Cin=3 stem uses
//   an overlapping-window view of the 8-channel-padded image (x stride = one pixel, box = 8 pixels x 8 ch)
//   so that one K-block = one filter row.
//
//   warp 0      : TMA producer (one elected lane), ring of `stages` shared-memory stages
//   warp 1      : tcgen05.mma issuer (one elected lane), accumulators double-buffered in TMEM
//   warps 2..5  : epilogue: tcgen05.ld -> scale/shift (+residual) (+ReLU) -> bf16 -> global (NHWC)
//   persistent CTAs (grid = min(tiles, #SM)), every CTA owns a contiguous range of tiles (pixel tile fastest) and
//   walks it in steps of `msub` (<= 2) pixel tiles that share the weight boxes of each K-block.
//
// Reference ops replaced: nn.Conv2d + BatchNorm2d (+ReLU, +residual) in BasicBlock.forward (dla.py:34-51),
// Root.forward (:124-132), Tree.project (:181-185), Conv2dBlock.forward (dla_neck.py:34-38), the stem /
// level0 / level1 (dla.py:231-237) and the nine head stems (monocon_heads.py:114-131).
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "engine.h"
#include "tc_epilogue.cuh"

namespace mc {

namespace {

constexpr int kTileM = 128;
constexpr int kThreads = 192;
constexpr int kMaxStages = 8;
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;          // TMEM columns per accumulator stage
constexpr long long kSpinLimit = 4000000000LL;   // ~2 s at 2 GHz: a stuck pipeline traps instead of hanging

struct KBlock { int src, c, dx, p, dy, dn; };   // dn: image offset (fp32-accurate mode: max_batch selects the lo plane)

struct TcParams {
    CUtensorMap map_a[kMaxSrc];
    CUtensorMap map_b;
    const KBlock* kblocks;
    int nkb;            // K-blocks per output tile
    int G;              // K-blocks per pipeline stage
    int msub;           // pixel tiles per pipeline step (1 or 2): a pair of 128-pixel tiles shares every weight box, which
                        // cuts the L2 -> shared-memory traffic (the chip-wide L2 throughput cap, not the tensor pipe, bounds
                        // the Cout >= 128 layers when every tile fetches its own copy of the weights)
    int stages;
    int bk;             // channels per K-block (16 / 32 / 64)
    int a_bytes, b_bytes;      // bytes of one K-block's A / B tile (1024-aligned strides used in smem)
    int a_stride, b_stride;
    int n_tile;         // cout per tile
    int n_tiles;        // cout tiles
    int tw, th, tn;     // pixel tile = tw x th x tn = 128
    int tiles_x, tiles_y, tiles_n;
    int Hout, Wout, B, Cout;
    const float* scale;
    const float* shift;
    const void* residual;
    void* dst;
    int relu;
    // fp32-accurate mode (DT_SPLIT sources: fp16 hi / lo planes; see common.cuh and conv_tc3.cu)
    int f16;
    long long dst_plane, res_plane;
    const ActScale* in_sc;
    const ActScale* out_sc;
    const ActScale* res_sc;
    unsigned* amax;
    // fused 2x2 max-pool of the output (OM_SPLIT; needs tw in {2..16}, th >= 2): pooled tensor, same scale as dst
    void* pool_dst;
    long long pool_plane;
    unsigned* pool_amax;
    int* error_flag;
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
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
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* error_flag, int code) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > kSpinLimit) {
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

__device__ __forceinline__ void tma_load_5d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3,
                                            int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(smem)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// K-major shared-memory matrix descriptor (rows of `row_bytes` = the swizzle span, 8-row groups contiguous).
//   bits [0,14) start >> 4 | [16,30) LBO >> 4 (unused for swizzled K-major) | [32,46) SBO >> 4
//   [46,48) version = 1 (Blackwell) | [61,64) layout: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 6 = SWIZZLE_32B
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, int row_bytes) {
    const uint64_t layout = row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
    const uint64_t sbo = (uint64_t)(8 * row_bytes) >> 4;
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32, M = 128, N = n_tile, K = 16
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_bf16_split(uint32_t tmem_d, uint32_t alo, uint32_t blo, uint32_t hi, uint32_t idesc,
                                                uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
        ::"r"(tmem_d), "r"(alo), "r"(blo), "r"(hi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
// EG = 2: warps 6..9 are a second epilogue group; the groups take alternate steps (a step = one or two pixel tiles), so a
// step's accumulators may take two steps of MMA time to drain (the 1x1 Root / project layers issue 2-20 MMAs per tile and
// are bound by the one-thread-per-row epilogue)
template <int EG, int OM>
__global__ void __launch_bounds__(64 + 128 * EG, 1) conv_tc_kernel(const __grid_constant__ TcParams p) {
    constexpr int kThreadsK = 64 + 128 * EG;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [stages][G] A tiles, [stages][G] B tiles (1024-aligned), then scale/shift, barriers
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stage_a = p.G * p.msub * p.a_stride, stage_b = p.G * p.b_stride;
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + (size_t)p.stages * stage_a;
    float* s_scale = reinterpret_cast<float*>(smem_b + (size_t)p.stages * stage_b);
    float* s_shift = s_scale + p.Cout;
    uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(s_shift + p.Cout) + 15) & ~uintptr_t(15));
    uint64_t* full_bar = bars;                       // [kMaxStages]
    uint64_t* empty_bar = bars + kMaxStages;         // [kMaxStages]
    uint64_t* tmem_full = bars + 2 * kMaxStages;     // [2]
    uint64_t* tmem_empty = bars + 2 * kMaxStages + 2;   // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_m_tiles = p.tiles_x * p.tiles_y * p.tiles_n;
    const int total_tiles = num_m_tiles * p.n_tiles;
    // contiguous tile range of this CTA; tile = nt * num_m_tiles + mt
    const int per_cta = total_tiles / (int)gridDim.x, rem_cta = total_tiles % (int)gridDim.x;
    const int t_begin = (int)blockIdx.x * per_cta + min((int)blockIdx.x, rem_cta);
    const int t_end = t_begin + per_cta + ((int)blockIdx.x < rem_cta ? 1 : 0);
    // pairs of pixel tiles with Cout tile > 128 columns need both 256-column accumulator slots at once
    const bool big = p.msub > 1 && p.n_tile > 128;

    {
        const float in_inv = p.in_sc ? p.in_sc->inv : 1.f, out_mul = (OM == tcepi::OM_SPLIT && p.out_sc) ? p.out_sc->mul : 1.f;
        for (int i = threadIdx.x; i < p.Cout; i += kThreadsK) {
            s_scale[i] = p.scale[i] * in_inv * out_mul;
            s_shift[i] = p.shift[i] * out_mul;
        }
    }
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kMaxSrc; ++s) prefetch_tmap(&p.map_a[s]);
        prefetch_tmap(&p.map_b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 128); }
        fence_barrier_init();
    }
    if (warp == 1) {     // TMEM allocation by one full warp; the same warp frees it
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_sync();          // everything above is independent of the previous kernel's output

    // tiles of one step: `cnt` consecutive pixel tiles of the same Cout tile
    auto step_cnt = [&](int t) { return (p.msub > 1 && t + 1 < t_end && (t % num_m_tiles) + 1 < num_m_tiles) ? 2 : 1; };

    if (warp == 0) {
        // ===================== TMA producer (whole warp converged, one elected lane issues) =====================
        {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = t_begin; t < t_end;) {
                const int cnt = step_cnt(t);
                const int co0 = (t / num_m_tiles) * p.n_tile;
                int x0[2], y0[2], n0[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    int mt = t % num_m_tiles + (j < cnt ? j : 0);
                    x0[j] = (mt % p.tiles_x) * p.tw; mt /= p.tiles_x;
                    y0[j] = (mt % p.tiles_y) * p.th;
                    n0[j] = (mt / p.tiles_y) * p.tn;
                }
                for (int kb0 = 0; kb0 < p.nkb; kb0 += p.G) {
                    const int g_cnt = min(p.G, p.nkb - kb0);
                    mbar_wait(&empty_bar[stage], phase ^ 1u, p.error_flag, 1);
                    if (elect_one()) {
                        mbar_expect_tx(&full_bar[stage], (uint32_t)g_cnt * (uint32_t)(cnt * p.a_bytes + p.b_bytes));
                        for (int g = 0; g < g_cnt; ++g) {
                            const KBlock kb = p.kblocks[kb0 + g];
                            uint8_t* a_dst = smem_a + (size_t)stage * stage_a + (size_t)(g * p.msub) * p.a_stride;
                            tma_load_5d(a_dst, &p.map_a[kb.src], &full_bar[stage], kb.c, x0[0] + kb.dx, kb.p, y0[0] + kb.dy, n0[0] + kb.dn);
                            if (cnt == 2)
                                tma_load_5d(a_dst + p.a_stride, &p.map_a[kb.src], &full_bar[stage], kb.c, x0[1] + kb.dx, kb.p,
                                            y0[1] + kb.dy, n0[1] + kb.dn);
                            tma_load_2d(smem_b + (size_t)stage * stage_b + (size_t)g * p.b_stride, &p.map_b, &full_bar[stage], 0,
                                        (kb0 + g) * p.Cout + co0);
                        }
                    }
                    __syncwarp();
                    if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                }
                t += cnt;
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp converged, one elected lane issues) =====================
        {
            // instruction descriptor: fp32 accumulate, A/B bf16, both K-major, N = n_tile, M = 128
            const uint32_t fmt = p.f16 ? 0u : 1u;     // fp16 (fp32-accurate mode) or bf16 operands
            const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.n_tile >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
            const int row_bytes = p.bk * 2;
            const int ksteps = p.bk / 16;
            // descriptor halves (A and B share layout / SBO): hi = SBO | version 1 | layout; lo = LBO(=1) | address >> 4
            const uint32_t layout = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);
            const uint32_t hi = (uint32_t)((8 * row_bytes) >> 4) | (1u << 14) | (layout << 29);
            const uint32_t a_base16 = (1u << 16) | ((smem_u32(smem_a) & 0x3FFFF) >> 4);
            const uint32_t b_base16 = (1u << 16) | ((smem_u32(smem_b) & 0x3FFFF) >> 4);
            const uint32_t stage_a16 = (uint32_t)stage_a >> 4, stage_b16 = (uint32_t)stage_b >> 4;
            const uint32_t a_stride16 = (uint32_t)p.a_stride >> 4, b_stride16 = (uint32_t)p.b_stride >> 4;
            const uint32_t a_step16 = a_stride16 * (uint32_t)p.msub;
            const int nkb = p.nkb, G = p.G, stages = p.stages;
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase[2] = {0u, 0u};
            for (int t = t_begin; t < t_end;) {
                const int cnt = step_cnt(t);
                uint32_t d0, d1;
                if (!big) {
                    mbar_wait(&tmem_empty[acc], acc_phase[acc] ^ 1u, p.error_flag, 2);
                    d0 = tmem_base + (uint32_t)(acc * kAccStride);
                    d1 = d0 + (uint32_t)p.n_tile;
                } else {
                    mbar_wait(&tmem_empty[0], acc_phase[0] ^ 1u, p.error_flag, 2);
                    if (cnt == 2) mbar_wait(&tmem_empty[1], acc_phase[1] ^ 1u, p.error_flag, 2);
                    d0 = tmem_base;
                    d1 = tmem_base + (uint32_t)kAccStride;
                }
                tc_fence_after();
                uint32_t accumulate = 0;
                for (int kb0 = 0; kb0 < nkb; kb0 += G) {
                    const int g_cnt = min(G, nkb - kb0);
                    mbar_wait(&full_bar[stage], phase, p.error_flag, 3);
                    tc_fence_after();
                    uint32_t alo = a_base16 + (uint32_t)stage * stage_a16;
                    uint32_t blo = b_base16 + (uint32_t)stage * stage_b16;
                    if (elect_one()) {
                        if (cnt == 1) {
                            for (int g = 0; g < g_cnt; ++g) {
#pragma unroll 4
                                for (int k = 0; k < ksteps; ++k) {
                                    umma_bf16_split(d0, alo + 2u * k, blo + 2u * k, hi, idesc, accumulate);
                                    accumulate = 1;
                                }
                                alo += a_step16;
                                blo += b_stride16;
                            }
                        } else {
                            for (int g = 0; g < g_cnt; ++g) {
#pragma unroll 4
                                for (int k = 0; k < ksteps; ++k) {
                                    umma_bf16_split(d0, alo + 2u * k, blo + 2u * k, hi, idesc, accumulate);
                                    umma_bf16_split(d1, alo + a_stride16 + 2u * k, blo + 2u * k, hi, idesc, accumulate);
                                    accumulate = 1;
                                }
                                alo += a_step16;
                                blo += b_stride16;
                            }
                        }
                        umma_commit(&empty_bar[stage]);      // frees the smem stage when these MMAs retire
                    }
                    __syncwarp();
                    accumulate = 1;
                    if (++stage == stages) { stage = 0; phase ^= 1u; }
                }
                if (!big) {
                    if (elect_one()) umma_commit(&tmem_full[acc]);   // accumulator(s) complete -> epilogue
                    __syncwarp();
                    acc_phase[acc] ^= 1u;
                    acc ^= 1;
                } else {
                    if (elect_one()) {
                        umma_commit(&tmem_full[0]);
                        if (cnt == 2) umma_commit(&tmem_full[1]);
                    }
                    __syncwarp();
                    acc_phase[0] ^= 1u;
                    if (cnt == 2) acc_phase[1] ^= 1u;
                }
                t += cnt;
            }
        }
    } else {
        // ===================== epilogue (warps 2..5 = TMEM lane quarters 2,3,0,1) =====================
        const int q = warp & 3;
        const int row = q * 32 + lane;                       // accumulator row == pixel index inside the tile
        const int ix = row % p.tw;
        const int iy = (row / p.tw) % p.th;
        const int in = row / (p.tw * p.th);
        const int eg = (warp - 2) >> 2;
        int acc = 0;
        uint32_t acc_phase[2] = {0u, 0u};
        int ord = 0;
        constexpr int EB = tcepi::ElemBytes<OM>::value;
        tcepi::SplitEpi se;
        se.dst_plane = p.dst_plane; se.res_plane = p.res_plane;
        se.res_mul = (OM == tcepi::OM_SPLIT) ? (p.res_sc ? p.res_sc->inv : 1.f) * (p.out_sc ? p.out_sc->mul : 1.f) : 1.f;
        float amax = 0.f;
        for (int t = t_begin; t < t_end; ++ord) {
            const int cnt = step_cnt(t);
            // two groups: with one accumulator stage per step the groups take alternate steps (group g only ever touches
            // stage g and its barriers); in `big` mode a step fills both 256-column slots and group g drains slot g of
            // every step.  Either way a barrier is consumed by ONE group only: parity waits cannot skip a phase.
            if (EG == 2 && !big && (ord & 1) != eg) {
                acc ^= 1;
                t += cnt;
                continue;
            }
            const int co0 = (t / num_m_tiles) * p.n_tile;
            for (int j = 0; j < cnt; ++j) {
                if (EG == 2 && big && j != eg) continue;
                int mt = t % num_m_tiles + j;
                const int tx = mt % p.tiles_x; mt /= p.tiles_x;
                const int ty = mt % p.tiles_y;
                const int tb = mt / p.tiles_y;
                const int x = tx * p.tw + ix, y = ty * p.th + iy, n = tb * p.tn + in;
                const bool valid = (x < p.Wout) && (y < p.Hout) && (n < p.B);
                const long long pix = ((long long)n * p.Hout + y) * p.Wout + x;
                char* dst = reinterpret_cast<char*>(p.dst) + (pix * p.Cout + co0) * EB;
                const char* res = p.residual ? reinterpret_cast<const char*>(p.residual) + (pix * p.Cout + co0) * EB : nullptr;
                const int slot = big ? j : acc;              // 256-column accumulator slot holding this pixel tile
                const int col = big ? j * kAccStride : acc * kAccStride + j * p.n_tile;
                if (big || j == 0) {
                    mbar_wait(&tmem_full[slot], acc_phase[slot], p.error_flag, 4);
                    tc_fence_after();
                }
                const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)col;
                // 64-column blocks only with one epilogue group (the 320-thread variant is capped at 168 registers)
                if (OM == tcepi::OM_SPLIT && p.pool_dst != nullptr) {
                    // row = iy * tw + ix inside the tile (tn == 1, tw <= 16): the window of (y, x) is lanes l, l ^ 1, l ^ tw
                    tcepi::PoolEpi pe;
                    const bool anchor = valid && !(ix & 1) && !(iy & 1);
                    const long long ppix = ((long long)n * (p.Hout >> 1) + (y >> 1)) * (p.Wout >> 1) + (x >> 1);
                    pe.dst = anchor ? reinterpret_cast<char*>(p.pool_dst) + (ppix * p.Cout + co0) * 2 : nullptr;
                    pe.plane = p.pool_plane;
                    pe.ybit = p.tw;
                    tcepi::drain_row<OM, false>(t_row, p.n_tile, s_scale + co0, s_shift + co0, res, dst, valid, p.relu != 0, se, amax, nullptr, &pe);
                } else {
                    tcepi::drain_row<OM, EG == 1>(t_row, p.n_tile, s_scale + co0, s_shift + co0, res, dst, valid, p.relu != 0, se, amax);
                }
                if (big || j == cnt - 1) {
                    tc_fence_before();
                    mbar_arrive(&tmem_empty[slot]);          // 128 arrivals release the accumulator slot
                    acc_phase[slot] ^= 1u;
                }
            }
            if (!big) acc ^= 1;
            t += cnt;
        }
        if (OM == tcepi::OM_SPLIT) { tcepi::publish_amax(p.amax, amax); tcepi::publish_amax(p.pool_amax, amax); }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
    }
}

// ---------------------------------------------------------------------------------------------
// host side: tensor maps, K-block tables, weight packing
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
int g_num_sms = 148;
int g_max_smem = 0;

CUtensorMapSwizzle swizzle_for(int row_bytes) {
    return row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

void encode(CUtensorMap* map, void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box,
            int row_bytes, const std::string& what, bool f16 = false) {
    MC_CHECK(g_encode != nullptr, "cuTensorMapEncodeTiled entry point not resolved");
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = g_encode(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, base, dims, strides_bytes, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(row_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ") for " + what);
}

}  // namespace

struct TcConvPlan {
    TcParams p;
    KBlock* d_kblocks = nullptr;
    void* d_w = nullptr;
    int om = tcepi::OM_BF16;
    int* d_err = nullptr;
    size_t smem_bytes = 0;
    int grid = 0;
    bool stem = false;
};

void tc_kernels_init() {
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
    MC_CUDA(cudaFuncSetAttribute(conv_tc_kernel<1, tcepi::OM_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem));
    MC_CUDA(cudaFuncSetAttribute(conv_tc_kernel<2, tcepi::OM_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem));
    MC_CUDA(cudaFuncSetAttribute(conv_tc_kernel<1, tcepi::OM_SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem));
    MC_CUDA(cudaFuncSetAttribute(conv_tc_kernel<2, tcepi::OM_SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem));
}

static bool is_stem(const ConvLayer& L) { return L.k == 7 && L.cin == 3 && L.stride == 1; }

bool tc_conv_supported(const Net& net, const ConvLayer& L) {
    if (net.dt != DT_BF16 && net.dt != DT_SPLIT) return false;
    if (L.cout % 16 != 0) return false;
    for (int s : L.src)
        if (net.tensors[s].dt != net.dt) return false;
    if (net.tensors[L.dst].dt != net.dt || L.dst_override_f32) return false;
    if (L.residual >= 0 && net.tensors[L.residual].dt != net.dt) return false;
    if (is_stem(L)) return net.tensors[L.src[0]].C == 8 && L.src.size() == 1 && !net.tensors[L.src[0]].hl_interleaved;
    if (!(L.stride == 1 || L.stride == 2)) return false;
    for (int s : L.src) {
        const int C = net.tensors[s].C;
        if (!(C == 16 || C == 32 || C % 64 == 0)) return false;
        if (net.tensors[s].Wp != net.tensors[s].W) return false;
    }
    if (L.stride == 2) {
        const TensorInfo& t = net.tensors[L.src[0]];
        if (L.src.size() != 1 || (t.H & 1) || (t.W & 1) || L.k != 3 || L.pad != 1) return false;
    }
    return true;
}

// pooled: the epilogue max-pools the output with quad shuffles, so a tile must hold whole 2x2 windows inside one warp's 32 rows:
// one image per tile, 2 <= tw <= 16 columns, at least two rows
static void choose_tile(int Wout, int Hout, int B, int& tw, int& th, int& tn, bool pooled = false) {
    double best = -1;
    tw = pooled ? 16 : 128; th = pooled ? 8 : 1; tn = 1;
    for (int a = 1; a <= 128; a *= 2)
        for (int b = 1; a * b <= 128; b *= 2) {
            const int c = 128 / (a * b);
            if (pooled && (c != 1 || a < 2 || a > 16 || b < 2)) continue;
            const double tiles = (double)((Wout + a - 1) / a) * ((Hout + b - 1) / b) * ((B + c - 1) / c);
            double util = (double)Wout * Hout * B / (tiles * 128.0);
            util += 1e-6 * a - 1e-7 * c;        // tie-break: wide tiles, few images per tile
            if (util > best) { best = util; tw = a; th = b; tn = c; }
        }
}

void tc_conv_prepare(Net& net, ConvLayer& L, const std::vector<float>& w_oihw) {
    auto plan = std::make_shared<TcConvPlan>();
    TcParams& p = plan->p;
    std::memset(&p, 0, sizeof(p));
    plan->stem = is_stem(L);
    const TensorInfo& s0 = net.tensors[L.src[0]];
    const TensorInfo& d = net.tensors[L.dst];
    const int B = net.max_batch;

    // ---- K-blocks + packed weights [kb][cout][bk] ----
    int bk = 64;
    for (int s : L.src) bk = std::min(bk, net.tensors[s].C);
    if (plan->stem) bk = 64;
    // fp32-accurate mode: the K-block list is walked three times -- (hi plane, w_lo), (lo plane, w_hi), (hi plane, w_hi);
    // the cross terms first, see conv_tc3.cu
    const bool split = net.dt == DT_SPLIT;
    const std::vector<int> ew = split ? split_weight_exponents(w_oihw, L.cout) : std::vector<int>();
    std::vector<KBlock> kbs;
    std::vector<uint16_t> w;
    int pass = split ? 0 : 2;
    L.widx.clear();
    auto push_weights = [&](auto&& getter) {     // getter(o, kk) -> index into w_oihw for kk in [0, bk), -1 = zero
        for (int o = 0; o < L.cout; ++o)
            for (int kk = 0; kk < bk; ++kk) {
                const long long id = getter(o, kk);
                const float v = id >= 0 ? w_oihw[(size_t)id] : 0.f;
                if (L.keep_widx) L.widx.push_back((int)id);
                w.push_back(split ? split_weight_piece(v, ew[o], pass == 0) : bf16_bits(v));
            }
    };
    for (; pass < 3; ++pass) {
    const int dn = pass == 1 ? B : 0;
    const int kk2 = L.k * L.k;
    if (plan->stem) {
        // one K-block per filter row r: 8 pixels (7 taps + 1 zero) x 8 channels (3 + 5 zero); the activation is
        // stored with 4 zero columns left of x = 0 (pitch W + 8), so tap s of output x sits at column x + s + 1.
        for (int r = 0; r < 7; ++r) {
            kbs.push_back(KBlock{0, 0, 1, 0, r - 3, dn});
            push_weights([&](int o, int kk) -> long long {
                const int s = kk / 8, c = kk % 8;
                return (s < 7 && c < 3) ? ((long long)o * 3 + c) * 49 + r * 7 + s : -1;
            });
        }
    } else {
        int cbase = 0;
        std::vector<int> cb;
        for (int s : L.src) { cb.push_back(cbase); cbase += net.tensors[s].C; }
        for (int r = 0; r < L.k; ++r)
            for (int sx = 0; sx < L.k; ++sx)
                for (int si = 0; si < (int)L.src.size(); ++si) {
                    const int C = net.tensors[L.src[si]].C;
                    for (int c0 = 0; c0 < C; c0 += bk) {
                        KBlock kb;
                        kb.src = si; kb.dn = dn;
                        if (L.stride == 1) {
                            kb.c = c0; kb.dx = sx - L.pad; kb.p = 0; kb.dy = r - L.pad;
                        } else {          // stride 2 over the space-to-depth view
                            const int ty = r - L.pad, tx = sx - L.pad;
                            const int ph = ty & 1, pw = tx & 1;
                            kb.c = pw * C + c0; kb.dx = (tx - pw) / 2; kb.p = ph; kb.dy = (ty - ph) / 2;
                        }
                        kbs.push_back(kb);
                        const int cb0 = cb[si] + c0;
                        push_weights([&](int o, int kk) -> long long { return ((long long)o * L.cin + cb0 + kk) * kk2 + r * L.k + sx; });
                    }
                }
    }
    }
    p.nkb = (int)kbs.size();
    p.bk = bk;
    const int row_bytes = bk * 2;
    plan->d_kblocks = (KBlock*)net.arena.alloc(sizeof(KBlock) * kbs.size());
    MC_CUDA(cudaMemcpy(plan->d_kblocks, kbs.data(), sizeof(KBlock) * kbs.size(), cudaMemcpyHostToDevice));
    plan->d_w = net.arena.alloc(sizeof(uint16_t) * w.size());
    MC_CUDA(cudaMemcpy(plan->d_w, w.data(), sizeof(uint16_t) * w.size(), cudaMemcpyHostToDevice));
    L.w_packed = plan->d_w;
    plan->d_err = (int*)net.arena.alloc(sizeof(int));
    p.kblocks = plan->d_kblocks;
    p.error_flag = plan->d_err;

    // ---- tiling ----
    p.Hout = d.H; p.Wout = d.W; p.B = B; p.Cout = L.cout;
    const bool pooled = split && L.pool_dst >= 0;
    MC_CHECK(!pooled || (d.H % 2 == 0 && d.W % 2 == 0), "tc: fused max-pool needs an even output size");
    choose_tile(d.W, d.H, B, p.tw, p.th, p.tn, pooled);
    p.tiles_x = (d.W + p.tw - 1) / p.tw;
    p.tiles_y = (d.H + p.th - 1) / p.th;
    p.tiles_n = (B + p.tn - 1) / p.tn;
    int n_tile = L.cout;
    if (n_tile > 256) {
        n_tile = 256;
        while (L.cout % n_tile != 0 || n_tile % 16 != 0) n_tile -= 16;
    }
    p.n_tile = n_tile;
    p.n_tiles = L.cout / n_tile;
    p.a_bytes = kTileM * row_bytes;
    p.b_bytes = n_tile * row_bytes;
    p.a_stride = (p.a_bytes + 1023) / 1024 * 1024;
    p.b_stride = (p.b_bytes + 1023) / 1024 * 1024;
    p.G = std::max(1, std::min(64 / bk, p.nkb));
    const int total_tiles = p.tiles_x * p.tiles_y * p.tiles_n * p.n_tiles;
    const size_t fixed = 1024 + sizeof(float) * 2 * L.cout + 16 + 8 * (2 * kMaxStages + 4) + 16;
    // pair pixel tiles (shared weight boxes) when the weights are a large part of a tile's traffic, CTAs own more than
    // one tile and three pipeline stages still fit
    int msub = (total_tiles > g_num_sms && p.b_bytes >= p.a_bytes && !plan->stem) ? 2 : 1;
    if (const char* e = std::getenv("MC_V1_MSUB")) msub = std::max(1, std::min(msub, std::atoi(e)));
    if (msub == 2 && ((size_t)g_max_smem - fixed) / ((size_t)p.G * (2 * p.a_stride + p.b_stride)) < 3) msub = 1;
    p.msub = msub;
    const size_t stage_bytes = (size_t)p.G * (msub * p.a_stride + p.b_stride);
    int stages = (int)(((size_t)g_max_smem - fixed) / stage_bytes);
    stages = std::min(stages, kMaxStages);
    MC_CHECK(stages >= 2, "tc conv: not enough shared memory for 2 stages: " + L.name);
    p.stages = stages;
    plan->smem_bytes = fixed + (size_t)stages * stage_bytes;
    plan->grid = std::min(total_tiles, g_num_sms);

    // ---- tensor maps ----
    const cuuint64_t nimg = (cuuint64_t)B * (split ? 2 : 1);          // DT_SPLIT: the lo plane = images B .. 2B-1
    for (int si = 0; si < kMaxSrc; ++si) {
        const TensorInfo& t = net.tensors[L.src[std::min(si, (int)L.src.size() - 1)]];
        const cuuint64_t C = t.C, W = t.W, H = t.H;
        cuuint64_t dims[5], str[4];
        cuuint32_t box[5] = {(cuuint32_t)bk, (cuuint32_t)p.tw, 1, (cuuint32_t)p.th, (cuuint32_t)p.tn};
        if (plan->stem) {
            // overlapping windows: dim0 = 64 elements starting at a pixel, dim1 steps one pixel (8 ch = 16 B)
            const cuuint64_t Wp = t.Wp;
            dims[0] = 64; dims[1] = Wp - 7; dims[2] = 1; dims[3] = H; dims[4] = nimg;
            str[0] = 16; str[1] = Wp * 16; str[2] = Wp * 16; str[3] = H * Wp * 16;
        } else if (L.stride == 1) {
            dims[0] = C; dims[1] = W; dims[2] = 1; dims[3] = H; dims[4] = nimg;
            str[0] = C * 2; str[1] = W * C * 2; str[2] = W * C * 2; str[3] = H * W * C * 2;
        } else {
            dims[0] = 2 * C; dims[1] = W / 2; dims[2] = 2; dims[3] = H / 2; dims[4] = nimg;
            str[0] = 2 * C * 2; str[1] = W * C * 2; str[2] = 2 * W * C * 2; str[3] = H * W * C * 2;
        }
        encode(&p.map_a[si], t.ptr, 5, dims, str, box, row_bytes, L.name + " (activation)", split);
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)bk, (cuuint64_t)p.nkb * L.cout};
        cuuint64_t str[1] = {(cuuint64_t)row_bytes};
        cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)n_tile};
        encode(&p.map_b, plan->d_w, 2, dims, str, box, row_bytes, L.name + " (weights)", split);
    }
    p.scale = split ? net.upload_split_scale(L, ew) : L.scale;
    p.shift = L.shift;
    // dst_override: the RAW convolution output is wanted (a train-mode BatchNorm applies residual and ReLU afterwards)
    p.residual = (L.residual >= 0 && !L.dst_override) ? net.tensors[L.residual].ptr : nullptr;
    p.dst = L.dst_override ? L.dst_override : d.ptr;
    p.f16 = split ? 1 : 0;
    plan->om = split ? tcepi::OM_SPLIT : tcepi::OM_BF16;
    if (split) {
        p.in_sc = net.act_scale(L.src[0]);
        p.out_sc = net.act_scale(L.dst); p.amax = net.act_amax(L.dst); p.dst_plane = d.plane;
        if (L.residual >= 0) { p.res_sc = net.act_scale(L.residual); p.res_plane = net.tensors[L.residual].plane; }
        if (pooled) {
            const TensorInfo& pt = net.tensors[L.pool_dst];
            p.pool_dst = pt.ptr; p.pool_plane = pt.plane; p.pool_amax = net.act_amax(L.pool_dst);
        }
    }
    p.relu = (L.relu && !L.dst_override) ? 1 : 0;
    L.tc = plan;
}

void tc_conv_launch(const Net& net, const ConvLayer& L, int B, cudaStream_t st) {
    MC_CHECK(L.tc != nullptr, "tc conv not prepared: " + L.name);
    TcParams p = L.tc->p;
    p.B = B;
    p.tiles_n = (B + p.tn - 1) / p.tn;
    const int total_tiles = p.tiles_x * p.tiles_y * p.tiles_n * p.n_tiles;
    const int grid = std::min(total_tiles, g_num_sms);
    // two epilogue groups unless MC_V1_EG=1
    const char* e = std::getenv("MC_V1_EG");
    const int eg = (e && e[0] == '1') ? 1 : 2;
    const bool sp = L.tc->om == tcepi::OM_SPLIT;
    if (eg == 2) launch_k(sp ? conv_tc_kernel<2, tcepi::OM_SPLIT> : conv_tc_kernel<2, tcepi::OM_BF16>, dim3(grid), dim3(64 + 128 * 2), L.tc->smem_bytes, st, p);
    else launch_k(sp ? conv_tc_kernel<1, tcepi::OM_SPLIT> : conv_tc_kernel<1, tcepi::OM_BF16>, dim3(grid), dim3(kThreads), L.tc->smem_bytes, st, p);
}

}  // namespace mc
