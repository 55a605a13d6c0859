// tcgen05 convolution, "halo view" variant (v2) for sm_100a.
//
// v1 (conv_tc.cu) fetches one shifted 128-pixel box per filter tap, i.e. the activation crosses L2 -> shared
// memory nine times for a 3x3 convolution, and re-fetches the weights for every tile; measured on B200 that
// traffic (not the tensor pipe, not HBM) bounds v1 at ~20-45 % of the bf16 peak.  v2 removes both:
//
//   * ONE TMA box per (tile, 64-channel chunk) brings the (16+2) x (8*SUB+2) pixel halo tile into shared memory;
//     every filter tap is then only a different *start address* of the same tile in the UMMA shared-memory
//     descriptor (rows = pixels, 8-row groups = 8 pixels along x, SBO = halo row pitch).  tools/umma_probe.cu
//     verified on B200 that tcgen05.mma forms row addresses as start + (m/8)*SBO + (m%8)*row_pitch and applies
//     the 32/64/128-byte swizzle XOR on absolute shared-memory address bits, which is exactly how TMA wrote the
//     tile -- so arbitrary 16-byte-aligned starts and SBOs are legal views.
//   * stride-2 3x3 convolutions with C = 16 / 32 read a space-to-depth view [pw*C+c, W/2, ph, H/2, N] of the same
//     NHWC memory: one box holds all four parity planes, a tap is a start offset (row shift + pw*C*2 bytes).
//   * the 7x7 / Cin=3 stem reads 22 input rows of the 8-channel-padded image as flat 16-byte pixels, no swizzle;
//     with LBO = 16 B and SBO = row pitch the descriptor walks overlapping 8-pixel windows (a Toeplitz view), so
//     one filter row is four K=16 MMAs and the im2col matrix is never materialised.
//   * the weights of the CTA's Cout tile stay resident in shared memory for the CTA's lifetime (persistent CTAs,
//     one Cout tile per CTA, pixel tiles strided over the CTAs that share it).
//
//   warp 0: TMA producer (activation halo tiles, ring of a_slots)   warp 1: MMA issuer + TMEM allocator
//   warps 2..5: epilogue (TMEM -> scale/shift (+residual) (+ReLU) -> bf16 NHWC)
//
// Reference ops replaced: see conv_tc.cu.
#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "engine.h"
#include "tc_epilogue.cuh"

namespace mc {

namespace {

constexpr int kThreads2 = 192;          // warp 0 producer, warp 1 MMA, warps 2..5 epilogue
constexpr int kThreads2x = 320;         // + warps 6..9: second epilogue group (EG = 2 variants)
constexpr int kTileRows = 16;            // output rows per tile
constexpr int kMaxChunks = 24;           // fp32-accurate mode: three virtual chunks per real one (see conv_tc3.cu)
constexpr int kMaxPieces = 18;
constexpr int kMaxASlots = 8;
constexpr int kAccCols = 256;            // TMEM columns per accumulator stage
constexpr long long kSpinLimit2 = 4000000000LL;

struct Chunk { int src, c, p, plane, w; };   // plane: 0 = hi (or only) plane, 1 = lo plane of a DT_SPLIT source; w: first resident weight piece

struct Tc2Params {
    CUtensorMap map_a[kMaxSrc];
    CUtensorMap map_b;
    Chunk chunks[kMaxChunks];
    int nchunks;
    int nwpieces;                 // resident weight pieces (fp32-accurate mode: the two passes that use w_hi share one copy)
    int np;                       // B pieces (filter taps / filter rows) per chunk
    int piece_aoff[kMaxPieces];   // byte offset of the piece's view inside the halo tile
    int nk;                       // K=16 steps per piece (each +32 B in A and B)
    int a_layout, a_sbo, a_lbo, a_rowpitch8;   // UMMA descriptor fields of the A views; a_rowpitch8 = bytes per 8 pixels along x
    int a_tile_bytes, a_slot_stride, a_slots;
    int a_split, a_part_rows, a_part_bytes;   // the halo box is fetched as a_split TMA boxes of a_part_rows input rows each
    int b_layout, b_sbo, b_piece_stride, b_piece_bytes;
    int sub;                      // 8-pixel-wide sub-tiles per tile (tile = 16*rs x 8*sub pixels)
    int rs;                       // row stacking: one accumulator row holds `rs` vertically adjacent output pixels
                                  // (N = rs * Cout, Cout = 16): fewer, longer MMAs for the A-read-bound 16-channel layers
    int b_rows;                   // rows of one weight piece in the packed weight tensor (= rs * Cout)
    int n_tile, n_tiles, ctas_per_ntile;
    int cxmul, xmul, ax, ay;      // TMA coordinates: (c + x0*cxmul, x0*xmul + ax, p, y0 + ay, n)
    int tiles_x, tiles_y;
    int Hout, Wout, B, Cout;
    const float* scale;
    const float* shift;
    const void* residual;
    void* dst;
    int relu;
    // fp32-accurate mode (DT_SPLIT sources: fp16 hi / lo planes; see common.cuh and conv_tc3.cu)
    int f16, plane_imgs;
    long long dst_plane, res_plane;
    const ActScale* in_sc;
    const ActScale* out_sc;
    const ActScale* res_sc;
    unsigned* amax;
    // fused 2x2 max-pool of the output (OM_SPLIT, rs == 1): pooled tensor [B][Hout/2][Wout/2][Cout], same scale as dst
    void* pool_dst;
    long long pool_plane;
    unsigned* pool_amax;
    int diag;                     // timing diagnostics only (env MC_DIAG): 1 = epilogue drains TMEM but skips global
                                  // loads/stores, 2 = one MMA per chunk, 3 = both.  Results are wrong by design.
    int* error_flag;
    unsigned long long* trace;    // diagnostics (env MC_TRACE_LAYER): per-role wait cycles of CTA 0, see tc2_conv_launch
};

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count));
}
__device__ __forceinline__ void bar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ bool bar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(s_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity, int* error_flag, int code) {
    if (bar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!bar_try_wait(bar, parity)) {
        if (clock64() - t0 > kSpinLimit2) {
            if (error_flag) atomicExch(error_flag, code);
            __threadfence_system();
            asm volatile("trap;");
        }
    }
}
// bar_wait that also accumulates the cycles spent waiting (diagnostic trace)
__device__ __forceinline__ void bar_wait_t(uint64_t* bar, uint32_t parity, int* error_flag, int code, bool tr, long long& acc) {
    if (!tr) { bar_wait(bar, parity, error_flag, code); return; }
    const long long t0 = clock64();
    bar_wait(bar, parity, error_flag, code);
    acc += clock64() - t0;
}
__device__ __forceinline__ void tma5(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(s_u32(smem)), "l"(map), "r"(s_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma2(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(s_u32(smem)), "l"(map), "r"(s_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint64_t desc_of(uint32_t saddr, int layout, int sbo, int lbo) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) |
           ((uint64_t)layout << 61);
}
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// same instruction, descriptors passed as 32-bit halves (the issue loop only ever changes the low words)
__device__ __forceinline__ void mma_bf16_split(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi,
                                               uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(tmem_d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void ld_tmem16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
// one elected lane of a fully converged warp (the warp runs the role loop convergently so that descriptors, barrier
// addresses and coordinates stay in uniform registers; only the issue instructions are predicated on the elected lane)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

template <int NK, int SUB, int EG, int OM>
__global__ void __launch_bounds__(64 + 128 * EG, 1) conv_tc2_kernel(const __grid_constant__ Tc2Params p) {
    constexpr int kThreadsK = 64 + 128 * EG;
    // EG = 2 with several sub-tiles per tile: both groups work on every tile, group g drains sub-tiles g, g + 2, ...;
    // with one sub-tile per tile the groups take alternate tiles (= accumulator stages)
    constexpr bool kSplitSub = (EG == 2) && (SUB >= 2);
    extern __shared__ __align__(1024) uint8_t smem_raw2[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw2) + 1023) & ~uintptr_t(1023));
    // [B resident: nchunks*np pieces][A ring: a_slots][scale][shift][barriers]
    uint8_t* smem_b = smem;
    uint8_t* smem_a = smem + (size_t)p.nwpieces * p.b_piece_stride;
    float* s_scale = reinterpret_cast<float*>(smem_a + (size_t)p.a_slots * p.a_slot_stride);
    float* s_shift = s_scale + p.n_tile;
    uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(s_shift + p.n_tile) + 15) & ~uintptr_t(15));
    uint64_t* a_full = bars;                         // [kMaxASlots]
    uint64_t* a_empty = bars + kMaxASlots;           // [kMaxASlots]
    uint64_t* b_full = bars + 2 * kMaxASlots;        // [1]
    uint64_t* tmem_full = bars + 2 * kMaxASlots + 1; // [2]
    uint64_t* tmem_empty = bars + 2 * kMaxASlots + 3;   // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kMaxASlots + 5);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nt = blockIdx.x % p.n_tiles;           // this CTA's Cout tile (weights resident)
    const int slot = blockIdx.x / p.n_tiles;
    const int co0 = nt * p.n_tile;
    const int m_tiles = p.tiles_x * p.tiles_y * p.B;

    {
        const float in_inv = p.in_sc ? p.in_sc->inv : 1.f, out_mul = (OM == tcepi::OM_SPLIT && p.out_sc) ? p.out_sc->mul : 1.f;
        for (int i = threadIdx.x; i < p.n_tile; i += kThreadsK) {
            const int c = p.rs > 1 ? (i % p.Cout) : (co0 + i);
            s_scale[i] = p.scale[c] * in_inv * out_mul;
            s_shift[i] = p.shift[c] * out_mul;
        }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.a_slots; ++s) { bar_init(&a_full[s], 1); bar_init(&a_empty[s], 1); }
        bar_init(b_full, 1);
        for (int a = 0; a < 2; ++a) { bar_init(&tmem_full[a], 1); bar_init(&tmem_empty[a], kSplitSub ? 256 : 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s_u32(tmem_ptr)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===================== TMA producer (whole warp converged, one elected lane issues) =====================
        {
            // resident weights of this Cout tile: all pieces, one barrier
            const int pieces = p.nwpieces;
            if (elect_one()) {
                bar_expect_tx(b_full, (uint32_t)pieces * (uint32_t)p.b_piece_bytes);
                for (int i = 0; i < pieces; ++i) tma2(smem_b + (size_t)i * p.b_piece_stride, &p.map_b, b_full, 0, i * p.b_rows + co0);
            }
            __syncwarp();
            pdl_sync();      // the weights above are constants; activations below are the previous kernel's output
            int as = 0;
            uint32_t aphase = 0;
            const bool tr = (p.trace != nullptr) && blockIdx.x == 0;
            long long w_ae = 0;
            const long long t_begin = clock64();
            for (int m = slot; m < m_tiles; m += p.ctas_per_ntile) {
                const int tx = m % p.tiles_x;
                const int ty = (m / p.tiles_x) % p.tiles_y;
                const int n = m / (p.tiles_x * p.tiles_y);
                const int x0 = tx * 8 * SUB, y0 = ty * kTileRows * p.rs;
                for (int ci = 0; ci < p.nchunks; ++ci) {
                    const Chunk ch = p.chunks[ci];
                    bar_wait_t(&a_empty[as], aphase ^ 1u, p.error_flag, 11, tr, w_ae);
                    if (elect_one()) {
                        bar_expect_tx(&a_full[as], (uint32_t)p.a_tile_bytes);
                        for (int part = 0; part < p.a_split; ++part)
                            tma5(smem_a + (size_t)as * p.a_slot_stride + (size_t)part * p.a_part_bytes, &p.map_a[ch.src], &a_full[as],
                                 ch.c + x0 * p.cxmul, x0 * p.xmul + p.ax, ch.p, y0 + p.ay + part * p.a_part_rows, n + ch.plane * p.plane_imgs);
                    }
                    __syncwarp();
                    if (++as == p.a_slots) { as = 0; aphase ^= 1u; }
                }
            }
            if (tr && lane == 0) { p.trace[0] = (unsigned long long)(clock64() - t_begin); p.trace[1] = (unsigned long long)w_ae; }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp converged, one elected lane issues) =====================
        {
            const uint32_t fmt = p.f16 ? 0u : 1u;     // A / B format: fp16 (fp32-accurate mode) or bf16
            const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.n_tile >> 3) << 17) | ((128u >> 4) << 24);
            // descriptor halves: hi = SBO | version 1 | layout, lo = LBO | (address >> 4); only lo changes in the loop
            const uint32_t a_hi = (uint32_t)(p.a_sbo >> 4) | (1u << 14) | ((uint32_t)p.a_layout << 29);
            const uint32_t b_hi = (uint32_t)(p.b_sbo >> 4) | (1u << 14) | ((uint32_t)p.b_layout << 29);
            const uint32_t a_lo_c = (uint32_t)(p.a_lbo >> 4) << 16;
            const uint32_t b_lo_c = 1u << 16;
            const uint32_t rp8 = (uint32_t)p.a_rowpitch8 >> 4;
            const uint32_t n_tile = (uint32_t)p.n_tile;
            const int np = p.np, nchunks = p.nchunks, a_slots = p.a_slots;
            const uint32_t a_slot16 = (uint32_t)p.a_slot_stride >> 4, b_piece16 = (uint32_t)p.b_piece_stride >> 4;
            uint32_t aoff16[kMaxPieces];
#pragma unroll
            for (int j = 0; j < kMaxPieces; ++j) aoff16[j] = (uint32_t)p.piece_aoff[j] >> 4;
            bar_wait(b_full, 0u, p.error_flag, 12);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            int as = 0;
            uint32_t aphase = 0;
            int acc = 0;
            uint32_t acc_phase[2] = {0u, 0u};
            const uint32_t b_base16 = (s_u32(smem_b) & 0x3FFFF) >> 4;
            const uint32_t a_base16 = (s_u32(smem_a) & 0x3FFFF) >> 4;
            const bool tr = (p.trace != nullptr) && blockIdx.x == 0;
            long long w_te = 0, w_af = 0;
            const long long t_begin = clock64();
            for (int m = slot; m < m_tiles; m += p.ctas_per_ntile) {
                bar_wait_t(&tmem_empty[acc], acc_phase[acc] ^ 1u, p.error_flag, 13, tr, w_te);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d0 = tmem_base + (uint32_t)(acc * kAccCols);
                uint32_t accf = 0u;
                for (int ci = 0; ci < nchunks; ++ci) {
                    const uint32_t blo = b_lo_c | (b_base16 + (uint32_t)p.chunks[ci].w * b_piece16);
                    bar_wait_t(&a_full[as], aphase, p.error_flag, 14, tr, w_af);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t alo_slot = a_lo_c | (a_base16 + (uint32_t)as * a_slot16);
                    // ONE elected region per chunk: the tensor pipe's issue queue is shallow (tools/umma_timing.cu: a gap of
                    // ~130 clocks between groups of N <= 128 MMAs is not absorbed), so an elect + warp-sync per filter tap
                    // left the pipe idle between taps
                    if (elect_one()) {
#pragma unroll
                        for (int j = 0; j < kMaxPieces; ++j) {
                            if (j < np && !((p.diag & 2) && j > 0)) {
                                const uint32_t alo_j = alo_slot + aoff16[j];
                                const uint32_t blo_j = blo + (uint32_t)j * b_piece16;
                                const uint32_t first = (j == 0) ? accf : 1u;
#pragma unroll
                                for (int k = 0; k < NK; ++k) {
#pragma unroll
                                    for (int sj = 0; sj < SUB; ++sj)
                                        mma_bf16_split(d0 + (uint32_t)sj * n_tile, alo_j + (uint32_t)sj * rp8 + 2u * k, a_hi, blo_j + 2u * k,
                                                       b_hi, idesc, (k == 0) ? first : 1u);
                                }
                            }
                        }
                        mma_commit(&a_empty[as]);
                    }
                    __syncwarp();
                    accf = 1u;
                    if (++as == a_slots) { as = 0; aphase ^= 1u; }
                }
                if (elect_one()) mma_commit(&tmem_full[acc]);
                __syncwarp();
                acc_phase[acc] ^= 1u;
                acc ^= 1;
            }
            if (tr && lane == 0) {
                p.trace[2] = (unsigned long long)(clock64() - t_begin); p.trace[3] = (unsigned long long)w_te; p.trace[4] = (unsigned long long)w_af;
            }
        }
    } else {
        // ===================== epilogue: EG groups of four warps; with EG = 2 group g drains accumulator stage g ==========
        // (one thread per accumulator row is latency-bound -- TMEM load, residual load, ~300 dependent instructions per
        // 64 columns; on the wide head-stem tiles it took longer than the MMAs of a tile, so two groups work on alternate
        // tiles there.  Measured: head.stems 0.300 -> 0.273 ms; the full-resolution layers lose 8-10 % with two groups)
        pdl_sync();
        const int q = warp & 3;
        const int grp = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        const int iy = row >> 3, ixl = row & 7;
        int acc = grp;
        uint32_t acc_phase[2] = {0u, 0u};
        const bool tr = (p.trace != nullptr) && blockIdx.x == 0 && warp == 2;
        long long w_tf = 0;
        long long t_ph[3] = {0, 0, 0};
        const long long t_begin = clock64();
        constexpr int EB = tcepi::ElemBytes<OM>::value;
        tcepi::SplitEpi se;
        se.dst_plane = p.dst_plane; se.res_plane = p.res_plane;
        se.res_mul = (OM == tcepi::OM_SPLIT) ? (p.res_sc ? p.res_sc->inv : 1.f) * (p.out_sc ? p.out_sc->mul : 1.f) : 1.f;
        float amax = 0.f;
        int ord = 0;
        for (int m = slot; m < m_tiles; m += p.ctas_per_ntile, ++ord) {
            if (EG == 2 && !kSplitSub && (ord & 1) != grp) continue;
            if (EG == 1 || kSplitSub) acc = ord & 1;
            const int tx = m % p.tiles_x;
            const int ty = (m / p.tiles_x) % p.tiles_y;
            const int n = m / (p.tiles_x * p.tiles_y);
            const int y = (ty * kTileRows + iy) * p.rs;
            bar_wait_t(&tmem_full[acc], acc_phase[acc], p.error_flag, 15, tr, w_tf);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int sj = kSplitSub ? grp : 0; sj < SUB; sj += kSplitSub ? 2 : 1) {
                const int x = (tx * SUB + sj) * 8 + ixl;
                const long long pix = ((long long)n * p.Hout + y) * p.Wout + x;
                const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kAccCols + sj * p.n_tile);
                if (p.rs == 1) {
                    const bool valid = (x < p.Wout) && (y < p.Hout) && !(p.diag & 1);
                    char* dst = reinterpret_cast<char*>(p.dst) + (pix * p.Cout + co0) * EB;
                    const char* res = p.residual ? reinterpret_cast<const char*>(p.residual) + (pix * p.Cout + co0) * EB : nullptr;
                    // 64-column blocks only with one epilogue group (the 320-thread variants are capped at 168 registers)
                    if (OM == tcepi::OM_SPLIT && p.pool_dst != nullptr) {
                        // the 2x2 window of (y, x): lanes l, l ^ 1 (x + 1), l ^ 8 (y + 1); tiles start on even rows / columns
                        tcepi::PoolEpi pe;
                        const bool anchor = valid && !(iy & 1) && !(ixl & 1);
                        const long long ppix = ((long long)n * (p.Hout >> 1) + (y >> 1)) * (p.Wout >> 1) + (x >> 1);
                        pe.dst = anchor ? reinterpret_cast<char*>(p.pool_dst) + (ppix * p.Cout + co0) * 2 : nullptr;
                        pe.plane = p.pool_plane;
                        pe.ybit = 8;
                        tcepi::drain_row<OM, false>(t_row, p.n_tile, s_scale, s_shift, res, dst, valid, p.relu != 0, se, amax, nullptr, &pe);
                    } else {
                        tcepi::drain_row<OM, EG == 1>(t_row, p.n_tile, s_scale, s_shift, res, dst, valid, p.relu != 0, se, amax, (tr && EG == 1) ? t_ph : nullptr);
                    }
                } else {
                    // row-stacked (rs == 4, Cout == 16): 16-column chunk dy is output pixel (y + dy, x)
                    const void* rr[4] = {nullptr, nullptr, nullptr, nullptr};
                    void* dd[4];
                    bool ok[4];
#pragma unroll
                    for (int dy = 0; dy < 4; ++dy) {
                        dd[dy] = reinterpret_cast<char*>(p.dst) + (pix + (long long)dy * p.Wout) * 16 * EB;
                        ok[dy] = (x < p.Wout) && (y + dy < p.Hout);
                    }
                    tcepi::drain_block_ex<OM, 4>(t_row, s_scale, s_shift, rr, dd, ok, p.relu != 0, se, amax);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            bar_arrive(&tmem_empty[acc]);
            acc_phase[acc] ^= 1u;
        }
        if (OM == tcepi::OM_SPLIT) { tcepi::publish_amax(p.amax, amax); tcepi::publish_amax(p.pool_amax, amax); }
        if (tr && lane == 0) {
            p.trace[5] = (unsigned long long)(clock64() - t_begin); p.trace[6] = (unsigned long long)w_tf;
            p.trace[7] = (unsigned long long)t_ph[0]; p.trace[8] = (unsigned long long)t_ph[1]; p.trace[9] = (unsigned long long)t_ph[2];
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn2 g_encode2 = nullptr;
int g_num_sms2 = 148;
int g_max_smem2 = 0;

int layout_code(int row_bytes) { return row_bytes >= 128 ? 2 : (row_bytes == 64 ? 4 : (row_bytes == 32 ? 6 : 0)); }
CUtensorMapSwizzle swz(int layout) {
    return layout == 2 ? CU_TENSOR_MAP_SWIZZLE_128B : layout == 4 ? CU_TENSOR_MAP_SWIZZLE_64B
         : layout == 6 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
}
void encode2(CUtensorMap* map, void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box,
             int layout, const std::string& what, bool f16 = false) {
    MC_CHECK(g_encode2 != nullptr, "cuTensorMapEncodeTiled entry point not resolved");
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = g_encode2(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, base, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swz(layout), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ") for " + what);
}

enum Kind { K_S1 = 0, K_S2 = 1, K_STEM = 2 };

bool classify(const Net& net, const ConvLayer& L, Kind& kind) {
    if (L.k == 7 && L.cin == 3 && L.stride == 1 && L.pad == 3 && L.src.size() == 1 && net.tensors[L.src[0]].C == 8 &&
        net.tensors[L.src[0]].xoff >= 3) {
        kind = K_STEM;
        return true;
    }
    if (L.k != 3 || L.pad != 1) return false;
    for (int s : L.src)
        if (net.tensors[s].Wp != net.tensors[s].W) return false;
    if (L.stride == 1) {
        kind = K_S1;
        if (L.src.size() == 1 && (net.tensors[L.src[0]].C == 16 || net.tensors[L.src[0]].C == 32)) return true;
        for (int s : L.src)
            if (net.tensors[s].C % 64 != 0) return false;
        return true;
    }
    if (L.stride == 2) {
        kind = K_S2;
        const TensorInfo& t = net.tensors[L.src[0]];
        return L.src.size() == 1 && (t.C == 16 || t.C == 32) && (t.H % 2 == 0) && (t.W % 2 == 0);
    }
    return false;
}

}  // namespace

struct Tc2ConvPlan {
    Tc2Params p;
    void* d_w = nullptr;
    int* d_err = nullptr;
    size_t smem_bytes = 0;
    int om = tcepi::OM_BF16;
    std::vector<int> ew;          // fp32-accurate mode: weight exponents per output channel
};

typedef void (*Tc2Kernel)(const Tc2Params);
template <int OM> static Tc2Kernel kernel_om(int nk, int sub, int eg) {
    if (eg == 2) {
        if (nk == 4 && sub == 1) return conv_tc2_kernel<4, 1, 2, OM>;
        if (nk == 4 && sub == 2) return conv_tc2_kernel<4, 2, 2, OM>;
        if (nk == 2 && sub == 2) return conv_tc2_kernel<2, 2, 2, OM>;
        if (nk == 1 && sub == 4) return conv_tc2_kernel<1, 4, 2, OM>;
    }
#define MC_TC2_CASE(NK, SUB) if (nk == NK && sub == SUB) return conv_tc2_kernel<NK, SUB, 1, OM>;
    MC_TC2_CASE(1, 1) MC_TC2_CASE(1, 2) MC_TC2_CASE(1, 3) MC_TC2_CASE(1, 4)
    MC_TC2_CASE(2, 1) MC_TC2_CASE(2, 2) MC_TC2_CASE(2, 3) MC_TC2_CASE(2, 4)
    MC_TC2_CASE(4, 1) MC_TC2_CASE(4, 2) MC_TC2_CASE(4, 3) MC_TC2_CASE(4, 4)
#undef MC_TC2_CASE
    return nullptr;
}
// the fp32-output variant (head stems of the fp32-accurate mode) is never reached through this kernel: the tripled
// weights of a 64-channel layer do not fit, conv_tc3 takes those layers
static Tc2Kernel kernel_for(int nk, int sub, int eg, int om = tcepi::OM_BF16) {
    return om == tcepi::OM_SPLIT ? kernel_om<tcepi::OM_SPLIT>(nk, sub, eg) : kernel_om<tcepi::OM_BF16>(nk, sub, eg);
}
// variants that exist with two epilogue groups
static bool has_eg2(const Tc2Params& p) {
    return (p.nk == 4 && p.sub == 1) || (p.nk == 4 && p.sub == 2) || (p.nk == 2 && p.sub == 2) || (p.nk == 1 && p.sub == 4);
}
static int epi_groups_for(const Tc2Params& p) {
    const char* e = std::getenv("MC_TC2_EG");      // 1: never, 2: wherever the variant exists, 3: only the sub >= 2 variants
    if (e && e[0]) {
        const int v = std::atoi(e);
        if (v == 2) return has_eg2(p) ? 2 : 1;
        if (v == 3) return (has_eg2(p) && p.sub >= 2) || (p.nk == 4 && p.sub == 1 && p.n_tile > 64) ? 2 : 1;
        return 1;
    }
    // default (measured per layer, B = 16): two groups for the wide one-sub-tile tiles (head stems 0.267 -> 0.236 ms) and
    // for every multi-sub-tile variant that is not row-stacked (level1 0.080 -> 0.070, level2 residual layers 0.068 ->
    // 0.058 ms); the row-stacked 16-channel layers lose (level0 0.107 -> 0.117 ms)
    if (p.nk == 4 && p.sub == 1) return p.n_tile > 64 ? 2 : 1;
    return (has_eg2(p) && p.sub >= 2 && p.rs == 1) ? 2 : 1;
}
template <typename F> static void for_each_variant(F&& f) {
    const int nks[3] = {1, 2, 4};
    for (int om = 0; om < 2; ++om) {
        for (int a = 0; a < 3; ++a)
            for (int sb = 1; sb <= 4; ++sb) f(kernel_for(nks[a], sb, 1, om));
        f(kernel_for(4, 1, 2, om));
        f(kernel_for(4, 2, 2, om));
        f(kernel_for(2, 2, 2, om));
        f(kernel_for(1, 4, 2, om));
    }
}

void tc2_kernels_init() {
    int dev = 0;
    MC_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    MC_CUDA(cudaGetDeviceProperties(&prop, dev));
    g_num_sms2 = std::max(1, prop.multiProcessorCount - reserved_sms());
    g_max_smem2 = (int)prop.sharedMemPerBlockOptin;
    if (!g_encode2) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        MC_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        MC_CHECK(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available in this driver");
        g_encode2 = reinterpret_cast<EncodeTiledFn2>(fn);
    }
    for_each_variant([&](auto kern) { MC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem2)); });
}

// plan: fills `plan` and returns true when the layer fits the v2 scheme (resident weights + >= 2 halo slots)
static bool plan_tc2(const Net& net, const ConvLayer& L, Tc2ConvPlan& plan, std::vector<uint16_t>* weights, const std::vector<float>* w_oihw,
                     std::vector<int>* widx = nullptr) {
    const char* env = std::getenv("MC_TC2");
    if (env && env[0] == '0') return false;
    if ((net.dt != DT_BF16 && net.dt != DT_SPLIT) || L.cout % 16 != 0) return false;
    const bool split = net.dt == DT_SPLIT;
    for (int s : L.src)
        if (net.tensors[s].dt != net.dt) return false;
    if (net.tensors[L.dst].dt != net.dt || L.dst_override_f32) return false;              // fp32 output: conv_tc3
    if (L.residual >= 0 && net.tensors[L.residual].dt != net.dt) return false;
    if (const char* skip = std::getenv("MC_TC2_SKIP"))           // A/B knob: layers whose name contains this use v1
        if (skip[0] && L.name.find(skip) != std::string::npos) return false;
    Kind kind;
    if (!classify(net, L, kind)) return false;
    Tc2Params& p = plan.p;
    std::memset(&p, 0, sizeof(p));
    const TensorInfo& d = net.tensors[L.dst];
    p.Hout = d.H; p.Wout = d.W; p.B = net.max_batch; p.Cout = L.cout;

    // ---- geometry of the activation views ----
    int bk;              // K elements per piece (= B row elements)
    int a_row_bytes;     // bytes per halo-tile row (one pixel, or one parity pixel pair)
    std::vector<Chunk> chunks;
    if (kind == K_STEM) {
        bk = 64; p.np = 7; p.nk = 4;
        chunks.push_back(Chunk{0, (net.tensors[L.src[0]].xoff - 3) * 8, 0, 0, 0});
    } else if (kind == K_S1) {
        int cmin = 64;
        for (int s : L.src) cmin = std::min(cmin, net.tensors[s].C);
        bk = cmin;
        p.np = 9; p.nk = bk / 16;
        for (int si = 0; si < (int)L.src.size(); ++si)
            for (int c0 = 0; c0 < net.tensors[L.src[si]].C; c0 += bk) chunks.push_back(Chunk{si, c0, 0, 0, 0});
    } else {
        bk = net.tensors[L.src[0]].C;
        p.np = 9; p.nk = bk / 16;
        chunks.push_back(Chunk{0, 0, 0, 0, 0});
    }
    if ((int)chunks.size() * (split ? 3 : 1) > kMaxChunks) return false;
    const int real_chunks = (int)chunks.size();
    // fp32-accurate mode: 64-channel stride-1 layers run ~10 % faster on the streamed-weight kernel (two sub-tiles share each
    // weight box; here the doubled resident weights leave room for one sub-tile only).  MC_TC2_SPLIT64=1 keeps them here.
    if (split && kind == K_S1 && bk == 64 && !std::getenv("MC_TC2_SPLIT64")) return false;
    // fp32-accurate mode, network input: hi and lo of the three colour channels sit in ONE 16-byte pixel, so the stem needs
    // two passes -- [w_lo | 0] (hi x w_lo), then [w_hi | w_hi] (hi x w_hi + lo x w_hi in the same MMAs) -- instead of three
    const bool stem_hl = split && kind == K_STEM && net.tensors[L.src[0]].hl_interleaved;
    if (split && kind == K_STEM && !stem_hl) return false;
    // row stacking for the 16-channel full-resolution layers (stem, level0): N = 4 * 16
    const char* env_rs = std::getenv("MC_ROWSTACK");
    const int rs = (L.cout == 16 && (kind == K_STEM || kind == K_S1) && real_chunks == 1 && L.residual < 0 &&
                    !(env_rs && env_rs[0] == '0')) ? 4 : 1;
    const int Nv = rs * L.cout;                          // accumulator columns per output-pixel group
    const int krows = (kind == K_STEM ? 7 : 3) + rs - 1;  // input rows touched by one stacked output group
    if (kind == K_STEM) p.np = krows;
    else if (kind == K_S1) p.np = krows * 3;
    p.rs = rs;
    p.b_rows = Nv;
    if (split) {
        // fp32-accurate mode: hi-plane x w_lo, lo-plane x w_hi, then hi-plane x w_hi (cross terms first, conv_tc3.cu).
        // Resident weights: [w_lo of every real chunk][w_hi of every real chunk]; passes 1 and 2 share the w_hi copy.
        std::vector<Chunk> v;
        for (int pass = 0; pass < 3; ++pass) {
            if (stem_hl && pass == 1) continue;
            for (int r = 0; r < real_chunks; ++r)
                v.push_back(Chunk{chunks[r].src, chunks[r].c, chunks[r].p, (pass == 1) ? 1 : 0, ((pass == 0 ? 0 : real_chunks) + r) * p.np});
        }
        chunks.swap(v);
    } else {
        for (int r = 0; r < real_chunks; ++r) chunks[r].w = r * p.np;
    }
    const int w_chunks = split ? 2 * real_chunks : real_chunks;      // chunk-sized groups of resident weight pieces
    p.nwpieces = w_chunks * p.np;
    p.nchunks = (int)chunks.size();
    for (int i = 0; i < p.nchunks; ++i) p.chunks[i] = chunks[i];
    p.f16 = split ? 1 : 0;
    p.plane_imgs = net.max_batch;
    plan.om = split ? tcepi::OM_SPLIT : tcepi::OM_BF16;

    // ---- Cout tile: largest tile whose resident weights + 2 halo slots fit ----
    const int b_row_bytes = bk * 2;
    p.b_layout = layout_code(b_row_bytes);
    p.b_sbo = 8 * b_row_bytes;
    const size_t fixed = 1024 + 512;
    int n_tile = 0, sub = 1;
    bool fit = false;
    for (int nt_c = std::min(Nv, 256) / 16 * 16; nt_c >= 16 && !fit; nt_c -= 16) {
        if (Nv % nt_c != 0) continue;
        if (rs > 1 && nt_c != Nv) continue;
        int sub_max = std::max(1, std::min(kAccCols / nt_c, 4));     // accumulator stage = 256 TMEM columns
        if (const char* e = std::getenv("MC_TC2_SUBMAX")) sub_max = std::max(1, std::min(sub_max, std::atoi(e)));
        if (kind == K_STEM) sub_max = std::min(sub_max, 2);          // TMA inner box <= 256 elements
        for (int sb = sub_max; sb >= 1 && !fit; --sb) {
            if (sb > 1 && d.W % (8 * sb) != 0) continue;
            const int b_piece_bytes = nt_c * b_row_bytes;
            const int b_piece_stride = (b_piece_bytes + 1023) / 1024 * 1024;
            int rows;
            if (kind == K_STEM) { a_row_bytes = (8 * sb + 8) * 16; rows = kTileRows * rs + 6; }
            else if (kind == K_S1) { a_row_bytes = bk * 2; rows = (kTileRows * rs + 2) * (8 * sb + 2); }
            else { a_row_bytes = 2 * bk * 2; rows = (kTileRows + 1) * 2 * (8 * sb + 1); }
            const int a_tile_bytes = rows * a_row_bytes;
            const int a_slot_stride = (a_tile_bytes + 1023) / 1024 * 1024;
            const size_t bbytes = (size_t)p.nwpieces * b_piece_stride;
            const size_t avail = (size_t)g_max_smem2 - fixed - 8 * nt_c;
            if (bbytes + 2 * (size_t)a_slot_stride > avail) continue;
            fit = true;
            n_tile = nt_c; sub = sb;
            p.b_piece_bytes = b_piece_bytes; p.b_piece_stride = b_piece_stride;
            p.a_tile_bytes = a_tile_bytes; p.a_slot_stride = a_slot_stride;
            p.a_slots = (int)std::min<size_t>(kMaxASlots, (avail - bbytes) / a_slot_stride);
            plan.smem_bytes = fixed + 8 * nt_c + bbytes + (size_t)p.a_slots * a_slot_stride;
        }
    }
    if (!fit) return false;
    if (Nv / n_tile > 4) return false;                   // many Cout tiles re-fetch the halo too often: v1 is the better fit
    if (n_tile < 64 && Nv > n_tile) return false;       // short MMAs (N < 64) are issue / operand-read bound: v1 with wide N wins
    // Cout = 128 / 256 layers whose weights only fit as two or more resident Cout tiles (level3 3x3): v1 runs them as one
    // N = Cout tile with paired pixel tiles, measured 15-20 % faster than N = 64 here
    if (Nv > n_tile && L.cout <= 256 && !std::getenv("MC_TC2_SPLIT_OK")) return false;
    {
        // split the halo box along y into several TMA instructions (more requests in flight inside the TMA unit)
        const int box_rows = kind == K_STEM ? kTileRows * rs + 6 : (kind == K_S1 ? kTileRows * rs + 2 : kTileRows + 1);
        int want = 1;
        if (const char* e = std::getenv("MC_TC2_ASPLIT")) want = std::max(1, std::atoi(e));
        int split = 1;
        for (int k = 1; k <= want; ++k)
            if (box_rows % k == 0) split = k;
        if (kind == K_S2) split = 1;
        p.a_split = split;
        p.a_part_rows = box_rows / split;
        p.a_part_bytes = p.a_tile_bytes / split;
        if (p.a_part_bytes % 128 != 0) { p.a_split = 1; p.a_part_rows = box_rows; p.a_part_bytes = p.a_tile_bytes; }
    }
    p.n_tile = n_tile;
    p.n_tiles = Nv / n_tile;
    p.sub = sub;
    p.ctas_per_ntile = std::max(1, g_num_sms2 / p.n_tiles);
    p.tiles_x = (d.W + 8 * sub - 1) / (8 * sub);
    p.tiles_y = (d.H + kTileRows * rs - 1) / (kTileRows * rs);

    // ---- piece views (one piece = one input row [x one horizontal tap]; 8-row groups step `rs` input rows) ----
    if (kind == K_STEM) {
        p.a_layout = 0; p.a_lbo = 16; p.a_sbo = rs * a_row_bytes; p.a_rowpitch8 = 8 * 16;
        for (int r = 0; r < krows; ++r) p.piece_aoff[r] = r * a_row_bytes;
        p.cxmul = 8; p.xmul = 0; p.ax = 0; p.ay = -3;
    } else if (kind == K_S1) {
        const int PW = 8 * sub + 2;
        p.a_layout = layout_code(a_row_bytes); p.a_lbo = 16; p.a_sbo = rs * PW * a_row_bytes; p.a_rowpitch8 = 8 * a_row_bytes;
        for (int r = 0; r < krows; ++r)
            for (int s = 0; s < 3; ++s) p.piece_aoff[r * 3 + s] = (r * PW + s) * a_row_bytes;
        p.cxmul = 0; p.xmul = 1; p.ax = -1; p.ay = -1;
    } else {
        const int PW = 8 * sub + 1;
        p.a_layout = layout_code(a_row_bytes); p.a_lbo = 16; p.a_sbo = 2 * PW * a_row_bytes; p.a_rowpitch8 = 8 * a_row_bytes;
        for (int r = 0; r < 3; ++r)
            for (int s = 0; s < 3; ++s) {
                const int ty = r - 1, tx = s - 1;
                const int ph = ty & 1, pw = tx & 1, dy = (ty - ph) / 2, dx = (tx - pw) / 2;
                p.piece_aoff[r * 3 + s] = (((dy + 1) * 2 + ph) * PW + (dx + 1)) * a_row_bytes + pw * bk * 2;
            }
        p.cxmul = 0; p.xmul = 1; p.ax = -1; p.ay = -1;
    }

    // ---- weights [chunk][piece][cout][bk] ----
    if (weights && w_oihw) {
        const std::vector<float>& w = *w_oihw;
        if (split) plan.ew = split_weight_exponents(w, L.cout);
        weights->clear();
        if (widx) widx->clear();
        weights->reserve((size_t)p.nwpieces * Nv * bk);
        std::vector<int> cb;
        int cbase = 0;
        for (int s : L.src) { cb.push_back(cbase); cbase += net.tensors[s].C; }
        for (int wc = 0; wc < w_chunks; ++wc) {
            const bool want_lo = split && wc < real_chunks;
            const int ci = wc % real_chunks;                 // geometry of the real chunk (virtual chunk ci of pass 0)
            for (int j = 0; j < p.np; ++j)
                for (int nv = 0; nv < Nv; ++nv)
                    for (int kk = 0; kk < bk; ++kk) {
                        // accumulator column nv = dy * Cout + o: output row offset dy inside the stacked group
                        const int dy = nv / L.cout, o = nv % L.cout;
                        float v = 0.f;
                        long long id = -1;
                        if (kind == K_STEM) {
                            const int r = j - dy, s = kk / 8;                      // piece j = input row j of the group
                            int c = kk % 8;
                            bool lo_channel = false;                                // interleaved pixels: channels 4..6 hold the lo pieces
                            if (stem_hl && c >= 4) { c -= 4; lo_channel = true; }
                            if (r >= 0 && r < 7 && s < 7 && c < 3 && !(lo_channel && want_lo)) id = ((long long)o * 3 + c) * 49 + r * 7 + s;
                        } else if (kind == K_S1) {
                            const int r = j / 3 - dy, sx = j % 3;
                            const int cin_idx = cb[chunks[ci].src] + chunks[ci].c + kk;
                            if (r >= 0 && r < 3) id = ((long long)o * L.cin + cin_idx) * 9 + r * 3 + sx;
                        } else {
                            const int cin_idx = cb[chunks[ci].src] + chunks[ci].c + kk;
                            id = ((long long)o * L.cin + cin_idx) * 9 + j;
                        }
                        if (id >= 0) v = w[(size_t)id];
                        if (widx) widx->push_back((int)id);
                        weights->push_back(split ? split_weight_piece(v, plan.ew[o], want_lo) : bf16_bits(v));
                    }
        }
    }
    return true;
}

bool tc2_conv_supported(const Net& net, const ConvLayer& L) {
    Tc2ConvPlan tmp;
    return plan_tc2(net, L, tmp, nullptr, nullptr);
}

void tc2_conv_prepare(Net& net, ConvLayer& L, const std::vector<float>& w_oihw) {
    auto plan = std::make_shared<Tc2ConvPlan>();
    std::vector<uint16_t> w;
    MC_CHECK(plan_tc2(net, L, *plan, &w, &w_oihw, L.keep_widx ? &L.widx : nullptr), "tc2: layer not supported: " + L.name);
    const bool split = net.dt == DT_SPLIT;
    const cuuint64_t nimg = (cuuint64_t)net.max_batch * (split ? 2 : 1);       // DT_SPLIT: the lo plane = images B .. 2B-1
    Tc2Params& p = plan->p;
    Kind kind;
    classify(net, L, kind);
    const TensorInfo& d = net.tensors[L.dst];
    const int B = net.max_batch;
    plan->d_w = net.arena.alloc(sizeof(uint16_t) * w.size());
    MC_CUDA(cudaMemcpy(plan->d_w, w.data(), sizeof(uint16_t) * w.size(), cudaMemcpyHostToDevice));
    L.w_packed = plan->d_w;
    plan->d_err = (int*)net.arena.alloc(sizeof(int));
    p.error_flag = plan->d_err;
    const int bk = p.b_piece_bytes / p.n_tile / 2;

    for (int si = 0; si < kMaxSrc; ++si) {
        const TensorInfo& t = net.tensors[L.src[std::min(si, (int)L.src.size() - 1)]];
        const cuuint64_t C = t.C, W = t.W, H = t.H;
        cuuint64_t dims[5], str[4];
        cuuint32_t box[5];
        if (kind == K_STEM) {
            const cuuint64_t Wp = t.Wp;
            dims[0] = Wp * 8; dims[1] = 1; dims[2] = 1; dims[3] = H; dims[4] = nimg;
            str[0] = Wp * 16; str[1] = Wp * 16; str[2] = Wp * 16; str[3] = H * Wp * 16;
            box[0] = (cuuint32_t)((8 * p.sub + 8) * 8); box[1] = 1; box[2] = 1; box[3] = (cuuint32_t)p.a_part_rows; box[4] = 1;
        } else if (kind == K_S1) {
            dims[0] = C; dims[1] = W; dims[2] = 1; dims[3] = H; dims[4] = nimg;
            str[0] = C * 2; str[1] = W * C * 2; str[2] = W * C * 2; str[3] = H * W * C * 2;
            box[0] = (cuuint32_t)bk; box[1] = (cuuint32_t)(8 * p.sub + 2); box[2] = 1; box[3] = (cuuint32_t)p.a_part_rows; box[4] = 1;
        } else {
            dims[0] = 2 * C; dims[1] = W / 2; dims[2] = 2; dims[3] = H / 2; dims[4] = nimg;
            str[0] = 2 * C * 2; str[1] = W * C * 2; str[2] = 2 * W * C * 2; str[3] = H * W * C * 2;
            box[0] = (cuuint32_t)(2 * C); box[1] = (cuuint32_t)(8 * p.sub + 1); box[2] = 2; box[3] = kTileRows + 1; box[4] = 1;
        }
        encode2(&p.map_a[si], t.ptr, 5, dims, str, box, p.a_layout, L.name + " (activation halo)", split);
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)bk, (cuuint64_t)p.nwpieces * p.b_rows};
        cuuint64_t str[1] = {(cuuint64_t)bk * 2};
        cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)p.n_tile};
        encode2(&p.map_b, plan->d_w, 2, dims, str, box, p.b_layout, L.name + " (weights)", split);
    }
    p.scale = split ? net.upload_split_scale(L, plan->ew) : L.scale;
    p.shift = L.shift;
    // dst_override: the RAW convolution output is wanted (a train-mode BatchNorm applies residual and ReLU afterwards)
    p.residual = (L.residual >= 0 && !L.dst_override) ? net.tensors[L.residual].ptr : nullptr;
    p.dst = L.dst_override ? L.dst_override : d.ptr;
    if (split) {
        p.in_sc = net.act_scale(L.src[0]);
        p.out_sc = net.act_scale(L.dst); p.amax = net.act_amax(L.dst); p.dst_plane = d.plane;
        if (L.residual >= 0) { p.res_sc = net.act_scale(L.residual); p.res_plane = net.tensors[L.residual].plane; }
        if (L.pool_dst >= 0) {
            MC_CHECK(p.rs == 1 && d.H % 2 == 0 && d.W % 2 == 0, "tc2: fused max-pool needs an unstacked layer with even output size");
            const TensorInfo& pt = net.tensors[L.pool_dst];
            p.pool_dst = pt.ptr; p.pool_plane = pt.plane; p.pool_amax = net.act_amax(L.pool_dst);
        }
    }
    p.relu = (L.relu && !L.dst_override) ? 1 : 0;
    if (const char* e = std::getenv("MC_DIAG")) p.diag = std::atoi(e);
    L.tc2 = plan;
}

void tc2_conv_launch(const Net& net, const ConvLayer& L, int B, cudaStream_t st) {
    MC_CHECK(L.tc2 != nullptr, "tc2 conv not prepared: " + L.name);
    Tc2Params p = L.tc2->p;
    p.B = B;
    const int m_tiles = p.tiles_x * p.tiles_y * B;
    p.ctas_per_ntile = std::max(1, std::min(m_tiles, g_num_sms2 / p.n_tiles));
    const int grid = p.ctas_per_ntile * p.n_tiles;
    const int eg = epi_groups_for(p);
    Tc2Kernel kern = kernel_for(p.nk, p.sub, eg, L.tc2->om);
    MC_CHECK(kern != nullptr, "tc2: no kernel variant for nk/sub of " + L.name);
    const char* tl = std::getenv("MC_TRACE_LAYER");
    static unsigned long long* d_trace = nullptr;
    const bool trace = tl && L.name == tl;
    if (trace) {
        if (!d_trace) MC_CUDA(cudaMalloc(&d_trace, 16 * sizeof(unsigned long long)));
        MC_CUDA(cudaMemsetAsync(d_trace, 0, 16 * sizeof(unsigned long long), st));
        p.trace = d_trace;
    }
    launch_k(kern, dim3(grid), dim3(eg == 2 ? kThreads2x : kThreads2), L.tc2->smem_bytes, st, p);
    if (trace) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(st, &cs);
        if (cs == cudaStreamCaptureStatusNone) {
            unsigned long long h[16];
            MC_CUDA(cudaStreamSynchronize(st));
            MC_CUDA(cudaMemcpy(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost));
            const int tiles_cta0 = (m_tiles + p.ctas_per_ntile - 1) / p.ctas_per_ntile;
            std::fprintf(stderr, "[trace %s] tiles/CTA %d chunks %d np %d nk %d sub %d rs %d n_tile %d a_slots %d a_tile %d B | producer: loop %llu clk, wait a_empty %llu | mma: loop %llu, wait tmem_empty %llu, wait a_full %llu | epilogue: loop %llu, wait tmem_full %llu, loads-landed %llu, math+stores %llu (store issue %llu)\n",
                         L.name.c_str(), tiles_cta0, p.nchunks, p.np, p.nk, p.sub, p.rs, p.n_tile, p.a_slots, p.a_tile_bytes, h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8], h[9]);
        }
    }
}

}  // namespace mc
