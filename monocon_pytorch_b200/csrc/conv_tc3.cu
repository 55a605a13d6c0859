// tcgen05 convolution, "halo view + streamed weights" variant (v3) for sm_100a: 3x3 / stride 1 / pad 1 layers with
// Cin a multiple of 64 whose weights do not fit in shared memory (level3..5 BasicBlocks, the IDAUp proj / node
// convolutions of ida_0 / ida_1).
//
// Why: the tap-box kernel (conv_tc.cu) fetches one shifted activation box per filter tap AND one weight box per K-block
// per pixel tile; at N >= 128 that is 64-96 bytes per SM per clock at full tensor rate, above what the L2 can feed to
// 148 SMs at once (~43 B/clk/SM), so those layers sat at 35-40 % tensor pipe.  The resident-weight halo kernel
// (conv_tc2.cu) cannot hold 9 x Cin x Cout weights.  v3 combines the two ideas:
//
//   * activations: the halo tile of a (step, 64-channel chunk) is brought in ONCE and every tap is a start-address view
//     of it (conv_tc2.cu).  The M dimension walks "flattened padded rows" g = n * (H + 2) + y: accumulator row group
//     m / 8 is flattened row g0 + m / 8, m % 8 is the pixel inside an 8-pixel strip.  A halo tile is fetched row by row
//     (one small TMA box per flattened row f -> (image f / (H+2), input row f % (H+2) - 1); rows -1 and H are fully
//     out of bounds = the zero padding), so tiles run seamlessly across image boundaries and any H works; the two
//     garbage row groups per image are simply not stored.
//   * weights: streamed through their own ring, one [n_tile x 64] box per (chunk, tap); a step processes up to `sub`
//     vertically adjacent 128-pixel sub-tiles (accumulators side by side in TMEM) that all reuse each weight box.
//   * every CTA owns a contiguous, balanced (+-1) range of sub-tiles, so the last wave is not half empty.
//
//   warp 0: activation producer   warp 6: weight producer   warp 1: MMA issuer + TMEM allocator   warps 2..5: epilogue
//
// Reference ops replaced: nn.Conv2d 3x3 + BatchNorm2d (+residual) + ReLU of BasicBlock.forward (dla.py:34-51) and
// Conv2dBlock.forward (dla_neck.py:34-38, including the concat-free cat([skip, up]) of IDAUp.forward :104).
#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "engine.h"
#include "tc_epilogue.cuh"

namespace mc {

namespace {

constexpr int kThreads3 = 224;
constexpr int kMaxChunks3 = 24;       // fp32-accurate mode: three virtual chunks (hi x w_lo, lo x w_hi, hi x w_hi) per 64 channels
constexpr int kMaxASlots3 = 4;
constexpr int kMaxBSlots3 = 8;
constexpr int kMaxSub3 = 4;
constexpr long long kSpinLimit3 = 4000000000LL;

struct Chunk3 { int src, c, plane; };    // plane: 0 = hi (or the only) plane, 1 = lo plane of a DT_SPLIT source

struct Tc3Params {
    CUtensorMap map_a[kMaxSrc];   // box [64 ch][10 px][1][1 row][1 image], SWIZZLE_128B
    CUtensorMap map_b;            // box [64][n_tile], SWIZZLE_128B, rows = (chunk * 9 + tap) * Cout + cout
    Chunk3 chunks[kMaxChunks3];
    int nchunks;
    int sub;                      // sub-tiles (16 flattened rows x 8 px) per step
    int n_tile, n_tiles;
    int acc_stages;               // 2 when sub * n_tile <= half of the CTA's TMEM columns, else 1
    int tmem_cols;                // TMEM columns this CTA allocates: 512, or 256 when two CTAs share an SM (ctas_per_sm = 2)
    int acc_stride;               // TMEM columns per accumulator stage
    int ctas_per_sm;
    int H, W, B, Cout, Hp;        // Hp = H + 2
    int strips;                   // W / 8
    int tiles_g;                  // ceil(B * Hp / 16) sub-tiles per strip
    int a_row_bytes;              // shared-memory pitch of one halo row (10 px x 128 B, optionally padded to 1024)
    int a_slot_stride, a_slots;
    int b_bytes, b_slot_stride, b_slots;
    const float* scale;
    const float* shift;
    const void* residual;
    void* dst;
    int relu;
    // fp32-accurate mode (DT_SPLIT sources: fp16 hi / lo planes; see common.cuh)
    int f16;                      // operands are fp16 (else bf16)
    int plane_imgs;               // images per plane (= the engine's max_batch): image coordinate of plane 1 = n + plane_imgs
    long long dst_plane, res_plane;   // elements between the hi and lo planes of dst / residual (OM_SPLIT)
    const ActScale* in_sc;        // scale of the (shared) source exponent, null = 1
    const ActScale* out_sc;       // scale of the destination (OM_SPLIT), null = 1
    const ActScale* res_sc;
    unsigned* amax;               // running max |stored| of the destination (OM_SPLIT)
    double* stats;                // OM_F32: per-(image, channel) sum / sum of squares of the output, [B][Cout][2] (AttnBN instance
                                  // statistics of the head stems, attentive_norm.py:84-85), accumulated by the epilogue; null = off
    int diag;                     // timing diagnostics only (env MC_DIAG3; results are wrong by design): 1 = epilogue does not touch
                                  // TMEM or global memory, 2 = no activation TMA traffic after the first fill of each slot,
                                  // 4 = no weight TMA traffic after the first fill of each slot
    int* error_flag;
    unsigned long long* trace;    // diagnostics (env MC_TRACE_LAYER): per-role loop / wait cycles of CTA 0
};

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init3(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(count));
}
__device__ __forceinline__ void bar_expect_tx3(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_arrive3(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ bool bar_try_wait3(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(s32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// non-blocking probe (mbarrier.try_wait may suspend the thread for a while when the phase is not complete)
__device__ __forceinline__ bool bar_test3(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(s32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bar_wait3(uint64_t* bar, uint32_t parity, int* error_flag, int code) {
    if (bar_try_wait3(bar, parity)) return;
    const long long t0 = clock64();
    while (!bar_try_wait3(bar, parity)) {
        if (clock64() - t0 > kSpinLimit3) {
            if (error_flag) atomicExch(error_flag, code);
            __threadfence_system();
            asm volatile("trap;");
        }
    }
}
__device__ __forceinline__ void bar_wait3_t(uint64_t* bar, uint32_t parity, int* error_flag, int code, bool tr, long long& acc) {
    if (!tr) { bar_wait3(bar, parity, error_flag, code); return; }
    const long long t0 = clock64();
    bar_wait3(bar, parity, error_flag, code);
    acc += clock64() - t0;
}
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void tma5_3(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(s32(smem)), "l"(map), "r"(s32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma2_3(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(s32(smem)), "l"(map), "r"(s32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void mma_bf16_3(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(tmem_d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit3(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ bool elect3() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// Work decomposition shared by all roles.  Units are sub-tiles u = nt * S + s, s = strip * tiles_g + tg (tg fastest);
// a CTA owns [u_begin, u_end) and walks it in steps of up to `sub` sub-tiles of the same Cout tile and strip.
struct Walk {
    int u, u_end, S, tiles_g, sub;
    __device__ __forceinline__ Walk(const Tc3Params& p) {
        S = p.strips * p.tiles_g;
        tiles_g = p.tiles_g;
        sub = p.sub;
        const int total = S * p.n_tiles;
        const int per = total / (int)gridDim.x, rem = total % (int)gridDim.x;
        u = (int)blockIdx.x * per + min((int)blockIdx.x, rem);
        u_end = u + per + ((int)blockIdx.x < rem ? 1 : 0);
    }
    __device__ __forceinline__ bool done() const { return u >= u_end; }
    // sub-tiles of the step starting at u
    __device__ __forceinline__ int count() const {
        const int s = u % S;
        const int tg = s % tiles_g;
        return min(min(sub, tiles_g - tg), u_end - u);
    }
};

// EG = 2: warps 7..10 form a second epilogue group; group g drains the sub-tiles sj = g (mod 2) of every step, so a
// step's accumulators are emptied in half the time (matters where the epilogue is not hidden: single-buffered N = 256
// steps, and the tail after a CTA's last step)
template <int SUBMAX, int MINB, int EG, int OM>
__global__ void __launch_bounds__(kThreads3 + 128 * (EG - 1), MINB) conv_tc3_kernel(const __grid_constant__ Tc3Params p) {
    constexpr int kThreadsK = kThreads3 + 128 * (EG - 1);
    extern __shared__ __align__(1024) uint8_t smem_raw3[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw3) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + (size_t)p.a_slots * p.a_slot_stride;
    float* s_scale = reinterpret_cast<float*>(smem_b + (size_t)p.b_slots * p.b_slot_stride);
    float* s_shift = s_scale + p.Cout;
    uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(s_shift + p.Cout) + 15) & ~uintptr_t(15));
    uint64_t* a_full = bars;                                    // [kMaxASlots3]
    uint64_t* a_empty = a_full + kMaxASlots3;                   // [kMaxASlots3]
    uint64_t* b_full = a_empty + kMaxASlots3;                   // [kMaxBSlots3]
    uint64_t* b_empty = b_full + kMaxBSlots3;                   // [kMaxBSlots3]
    uint64_t* tmem_full = b_empty + kMaxBSlots3;                // [2]
    uint64_t* tmem_empty = tmem_full + 2;                       // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // per-CTA timeline (diagnostics): trace[16 + 4 * cta + {0: entry, 1: MMA loop begin, 2: MMA loop end, 3: exit}] in ns
    unsigned long long* tl = p.trace ? p.trace + 16 + 4 * blockIdx.x : nullptr;
    if (tl && threadIdx.x == 32) tl[0] = gtime();

    {
        // the tensor scales are constants of a forward pass (the host changes them between passes only)
        const float in_inv = p.in_sc ? p.in_sc->inv : 1.f, out_mul = (OM == tcepi::OM_SPLIT && p.out_sc) ? p.out_sc->mul : 1.f;
        for (int i = threadIdx.x; i < p.Cout; i += kThreadsK) {
            s_scale[i] = p.scale[i] * in_inv * out_mul;
            s_shift[i] = p.shift[i] * out_mul;
        }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.a_slots; ++s) { bar_init3(&a_full[s], 1); bar_init3(&a_empty[s], 1); }
        for (int s = 0; s < p.b_slots; ++s) { bar_init3(&b_full[s], 1); bar_init3(&b_empty[s], 1); }
        for (int a = 0; a < 2; ++a) { bar_init3(&tmem_full[a], 1); bar_init3(&tmem_empty[a], 128 * EG); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_ptr)), "r"((uint32_t)p.tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===================== activation producer: one TMA box per halo row, issued lane-parallel =====================
        pdl_sync();
        int as = 0;
        uint32_t aphase = 0;
        const bool tr = (p.trace != nullptr) && blockIdx.x == 0;
        long long w_ae = 0;
        const long long t_begin = clock64();
        for (Walk w(p); !w.done();) {
            const int cnt = w.count();
            const int s = w.u % w.S;
            const int strip = s / p.tiles_g, tg = s % p.tiles_g;
            const int rows = 16 * cnt + 2;
            const int f0 = 16 * tg;
            for (int ci = 0; ci < p.nchunks; ++ci) {
                const Chunk3 ch = p.chunks[ci];
                bar_wait3_t(&a_empty[as], aphase ^ 1u, p.error_flag, 31, tr, w_ae);
                if ((p.diag & 2) && aphase) {            // diagnostics: reuse the stale tile, only flip the barrier
                    if (elect3()) bar_arrive3(&a_full[as]);
                    __syncwarp();
                    if (++as == p.a_slots) { as = 0; aphase ^= 1u; }
                    continue;
                }
                if (elect3()) bar_expect_tx3(&a_full[as], (uint32_t)rows * 1280u);
                __syncwarp();
                uint8_t* slot = smem_a + (size_t)as * p.a_slot_stride;
                for (int r = lane; r < rows; r += 32) {
                    const int f = f0 + r;
                    const int n = f / p.Hp, yy = f - n * p.Hp - 1;
                    tma5_3(slot + (size_t)r * p.a_row_bytes, &p.map_a[ch.src], &a_full[as], ch.c, strip * 8 - 1, 0, yy, n + ch.plane * p.plane_imgs);
                }
                __syncwarp();
                if (++as == p.a_slots) { as = 0; aphase ^= 1u; }
            }
            w.u += cnt;
        }
        if (tr && lane == 0) { p.trace[0] = (unsigned long long)(clock64() - t_begin); p.trace[1] = (unsigned long long)w_ae; }
    } else if (warp == 6) {
        // ===================== weight producer (constants: no dependency on the previous kernel) =====================
        int bs = 0;
        uint32_t bphase = 0;
        const bool tr = (p.trace != nullptr) && blockIdx.x == 0;
        long long w_be = 0;
        const long long t_begin = clock64();
        for (Walk w(p); !w.done();) {
            const int cnt = w.count();
            const int co0 = (w.u / w.S) * p.n_tile;
            for (int ci = 0; ci < p.nchunks; ++ci) {
                for (int j = 0; j < 9; ++j) {
                    bar_wait3_t(&b_empty[bs], bphase ^ 1u, p.error_flag, 32, tr, w_be);
                    if ((p.diag & 4) && bphase) {
                        if (elect3()) bar_arrive3(&b_full[bs]);
                        __syncwarp();
                        if (++bs == p.b_slots) { bs = 0; bphase ^= 1u; }
                        continue;
                    }
                    if (elect3()) {
                        bar_expect_tx3(&b_full[bs], (uint32_t)p.b_bytes);
                        tma2_3(smem_b + (size_t)bs * p.b_slot_stride, &p.map_b, &b_full[bs], 0, (ci * 9 + j) * p.Cout + co0);
                    }
                    if (++bs == p.b_slots) { bs = 0; bphase ^= 1u; }
                }
            }
            w.u += cnt;
        }
        if (tr && lane == 0) { p.trace[2] = (unsigned long long)(clock64() - t_begin); p.trace[3] = (unsigned long long)w_be; }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp converged, one elected lane issues) =====================
        // fp32 accumulate; A / B format bf16 (1) or fp16 (0); N, M
        const uint32_t fmt = p.f16 ? 0u : 1u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.n_tile >> 3) << 17) | ((128u >> 4) << 24);
        // descriptor halves: hi = SBO | version 1 | SWIZZLE_128B, lo = LBO (1) | address >> 4
        const uint32_t a_hi = (uint32_t)(p.a_row_bytes >> 4) | (1u << 14) | (2u << 29);
        const uint32_t b_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
        const uint32_t a_base16 = (1u << 16) | ((s32(smem_a) & 0x3FFFF) >> 4);
        const uint32_t b_base16 = (1u << 16) | ((s32(smem_b) & 0x3FFFF) >> 4);
        const uint32_t a_slot16 = (uint32_t)p.a_slot_stride >> 4, b_slot16 = (uint32_t)p.b_slot_stride >> 4;
        const uint32_t row16 = (uint32_t)p.a_row_bytes >> 4;
        const uint32_t sub16 = 16u * row16;                      // next sub-tile = 16 halo rows further down
        const uint32_t n_tile = (uint32_t)p.n_tile;
        const int nchunks = p.nchunks, a_slots = p.a_slots, b_slots = p.b_slots;
        const bool dbuf = p.acc_stages == 2;
        int as = 0, bs = 0, acc = 0;
        uint32_t aphase = 0, bphase = 0;
        uint32_t acc_phase[2] = {0u, 0u};
        const bool tr = (p.trace != nullptr) && blockIdx.x == 0;
        long long w_te = 0, w_af = 0, w_bf = 0;
        const long long t_begin = clock64();
        if (tl && lane == 0) tl[1] = gtime();
        for (Walk w(p); !w.done();) {
            const int cnt = w.count();
            bar_wait3_t(&tmem_empty[acc], acc_phase[acc] ^ 1u, p.error_flag, 33, tr, w_te);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d0 = tmem_base + (uint32_t)(acc * p.acc_stride);
            uint32_t accf = 0u;
            for (int ci = 0; ci < nchunks; ++ci) {
                bar_wait3_t(&a_full[as], aphase, p.error_flag, 34, tr, w_af);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t alo_slot = a_base16 + (uint32_t)as * a_slot16;
                // ONE elected region per chunk; the elected lane itself waits for the weight boxes.  The tensor pipe's issue
                // queue is shallow (tools/umma_timing.cu), so an elect + warp-sync per tap shows up as idle tensor time.
                if (elect3()) {
                    int bsl = bs;
                    uint32_t bph = bphase;
                    // the barrier test of tap j + 1 is issued before the MMAs of tap j, so that its ~150-clock latency
                    // overlaps their issue instead of opening a gap in the (shallow) tensor-pipe queue
                    bool ready = bar_test3(&b_full[bsl], bph);
#pragma unroll
                    for (int j = 0; j < 9; ++j) {
                        if (!ready) bar_wait3_t(&b_full[bsl], bph, p.error_flag, 35, tr, w_bf);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        int bsn = bsl + 1;
                        uint32_t bpn = bph;
                        if (bsn == b_slots) { bsn = 0; bpn ^= 1u; }
                        if (j < 8) ready = bar_test3(&b_full[bsn], bpn);
                        const uint32_t alo = alo_slot + (uint32_t)(j / 3) * row16 + (uint32_t)(j % 3) * 8u;   // + 128 B per pixel
                        const uint32_t blo = b_base16 + (uint32_t)bsl * b_slot16;
                        const uint32_t first = (j == 0) ? accf : 1u;
#pragma unroll
                        for (int sj = 0; sj < SUBMAX; ++sj) {
                            if (sj < cnt) {
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    mma_bf16_3(d0 + (uint32_t)sj * n_tile, alo + (uint32_t)sj * sub16 + 2u * k, a_hi, blo + 2u * k, b_hi, idesc,
                                               (k == 0) ? first : 1u);
                            }
                        }
                        mma_commit3(&b_empty[bsl]);
                        bsl = bsn; bph = bpn;
                    }
                    mma_commit3(&a_empty[as]);
                }
                __syncwarp();
                accf = 1u;
                bs += 9;
                while (bs >= b_slots) { bs -= b_slots; bphase ^= 1u; }
                if (++as == a_slots) { as = 0; aphase ^= 1u; }
            }
            if (elect3()) mma_commit3(&tmem_full[acc]);
            __syncwarp();
            acc_phase[acc] ^= 1u;
            if (dbuf) acc ^= 1;
            w.u += cnt;
        }
        if (tl && lane == 0) tl[2] = gtime();
        if (tr && lane == 0) {
            p.trace[4] = (unsigned long long)(clock64() - t_begin); p.trace[5] = (unsigned long long)w_te;
            p.trace[6] = (unsigned long long)w_af; p.trace[7] = (unsigned long long)w_bf;
        }
    } else {
        // ===================== epilogue (warps 2..5 = TMEM lane quarters 2, 3, 0, 1) =====================
        pdl_sync();
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int grp = row >> 3, ixl = row & 7;
        const int eg = warp >= 7 ? 1 : 0;                     // epilogue group
        const bool dbuf = p.acc_stages == 2;
        int acc = 0;
        uint32_t acc_phase[2] = {0u, 0u};
        const bool tr = (p.trace != nullptr) && blockIdx.x == 0 && warp == 2;
        long long w_tf = 0;
        const long long t_begin = clock64();
        constexpr int EB = tcepi::ElemBytes<OM>::value;
        tcepi::SplitEpi se;
        se.dst_plane = p.dst_plane; se.res_plane = p.res_plane;
        se.res_mul = (OM == tcepi::OM_SPLIT) ? (p.res_sc ? p.res_sc->inv : 1.f) * (p.out_sc ? p.out_sc->mul : 1.f) : 1.f;
        float amax = 0.f;
        for (Walk w(p); !w.done();) {
            const int cnt = w.count();
            const int co0 = (w.u / w.S) * p.n_tile;
            const int s = w.u % w.S;
            const int strip = s / p.tiles_g, tg = s % p.tiles_g;
            const int x = strip * 8 + ixl;
            bar_wait3_t(&tmem_full[acc], acc_phase[acc], p.error_flag, 36, tr, w_tf);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int sj = eg; sj < ((p.diag & 1) ? 0 : cnt); sj += EG) {
                const int g = 16 * (tg + sj) + grp;
                const int n = g / p.Hp, y = g - n * p.Hp;
                const bool valid = (y < p.H) && (n < p.B) && (x < p.W);
                const long long pix = ((long long)n * p.H + y) * p.W + x;
                char* dst = reinterpret_cast<char*>(p.dst) + (pix * p.Cout + co0) * EB;
                const char* res = p.residual ? reinterpret_cast<const char*>(p.residual) + (pix * p.Cout + co0) * EB : nullptr;
                const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.acc_stride + sj * p.n_tile);
                if (OM == tcepi::OM_F32 && p.stats != nullptr) {
                    // the warp's 32 rows = flattened rows gA .. gA + 3: image A, or images A and A + 1 (pad rows belong to nobody)
                    const int gA = 16 * (tg + sj) + q * 4;
                    const int nA = gA / p.Hp, nB = (gA + 3) / p.Hp;
                    tcepi::StatsEpi st;
                    st.sums_a = p.stats + ((long long)nA * p.Cout + co0) * 2;
                    st.img_stride = (long long)p.Cout * 2;
                    st.in_a = valid && n == nA;
                    st.in_b = valid && n != nA;
                    st.one = nA < p.B;
                    st.two = (nB != nA) && (nB < p.B);
                    tcepi::drain_row_f32_stats(t_row, p.n_tile, s_scale + co0, s_shift + co0, reinterpret_cast<float*>(dst), valid, p.relu != 0, st, lane);
                    continue;
                }
                // 64-column blocks only in the 224-thread variant (the others are capped at 168 registers)
                tcepi::drain_row<OM, (MINB == 1 && EG == 1)>(t_row, p.n_tile, s_scale + co0, s_shift + co0, res, dst, valid, p.relu != 0, se, amax);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            bar_arrive3(&tmem_empty[acc]);
            acc_phase[acc] ^= 1u;
            if (dbuf) acc ^= 1;
            w.u += cnt;
        }
        if (OM == tcepi::OM_SPLIT) tcepi::publish_amax(p.amax, amax);
        if (tr && lane == 0) { p.trace[8] = (unsigned long long)(clock64() - t_begin); p.trace[9] = (unsigned long long)w_tf; }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols));
        if (tl && lane == 0) tl[3] = gtime();
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn3 g_encode3 = nullptr;
int g_num_sms3 = 148;
int g_max_smem3 = 0;

void encode3(CUtensorMap* map, void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box,
             const std::string& what, bool f16 = false) {
    MC_CHECK(g_encode3 != nullptr, "cuTensorMapEncodeTiled entry point not resolved");
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = g_encode3(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, base, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ") for " + what);
}

int env_int(const char* name, int dflt) {
    const char* e = std::getenv(name);
    return (e && e[0]) ? std::atoi(e) : dflt;
}

typedef void (*Tc3Kernel)(const Tc3Params);
template <int OM> Tc3Kernel kernel3_om(int sub, int ctas_per_sm, int eg) {
    if (ctas_per_sm == 2) return conv_tc3_kernel<1, 2, 1, OM>;
    if (eg == 2 && sub == 2) return conv_tc3_kernel<2, 1, 2, OM>;
    return sub <= 1 ? conv_tc3_kernel<1, 1, 1, OM> : (sub == 2 ? conv_tc3_kernel<2, 1, 1, OM> : conv_tc3_kernel<4, 1, 1, OM>);
}
Tc3Kernel kernel3_for(int sub, int ctas_per_sm, int eg, int om) {
    return om == tcepi::OM_SPLIT ? kernel3_om<tcepi::OM_SPLIT>(sub, ctas_per_sm, eg)
         : om == tcepi::OM_F32 ? kernel3_om<tcepi::OM_F32>(sub, ctas_per_sm, eg) : kernel3_om<tcepi::OM_BF16>(sub, ctas_per_sm, eg);
}
// two epilogue groups whenever a step has two sub-tiles (MC_TC3_EG=1 disables).  Measured on the 23 layers of this kernel
// (B = 16): 0.973 -> 0.940 ms; every layer gains 0-8 %, double-buffered ones included (shorter tail after the last step).
int epi_groups3(const Tc3Params& p) {
    if (p.sub != 2 || p.ctas_per_sm != 1) return 1;
    return env_int("MC_TC3_EG", 2) == 1 ? 1 : 2;
}

}  // namespace

struct Tc3ConvPlan {
    Tc3Params p;
    void* d_w = nullptr;
    int* d_err = nullptr;
    size_t smem_bytes = 0;
    int om = tcepi::OM_BF16;
};

void tc3_kernels_init() {
    int dev = 0;
    MC_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    MC_CUDA(cudaGetDeviceProperties(&prop, dev));
    g_num_sms3 = std::max(1, prop.multiProcessorCount - reserved_sms());
    g_max_smem3 = (int)prop.sharedMemPerBlockOptin;
    if (!g_encode3) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        MC_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        MC_CHECK(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available in this driver");
        g_encode3 = reinterpret_cast<EncodeTiledFn3>(fn);
    }
    for (int om = 0; om < 3; ++om) {
        MC_CUDA(cudaFuncSetAttribute(kernel3_for(1, 1, 1, om), cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem3));
        MC_CUDA(cudaFuncSetAttribute(kernel3_for(2, 1, 1, om), cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem3));
        MC_CUDA(cudaFuncSetAttribute(kernel3_for(4, 1, 1, om), cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem3));
        MC_CUDA(cudaFuncSetAttribute(kernel3_for(1, 2, 1, om), cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem3));
        MC_CUDA(cudaFuncSetAttribute(kernel3_for(2, 1, 2, om), cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem3));
    }
}

// fills the geometry part of the plan; false when the layer is outside this kernel's domain
static bool plan_tc3(const Net& net, const ConvLayer& L, Tc3ConvPlan& plan) {
    if (env_int("MC_TC3", 1) == 0) return false;
    if (net.dt != DT_BF16 && net.dt != DT_SPLIT) return false;
    const bool split = net.dt == DT_SPLIT;
    if (L.k != 3 || L.stride != 1 || L.pad != 1) return false;
    // the fp32-accurate mode triples the weights, so the Cout = 64 layers no longer fit the resident-weight kernel
    if (L.cout % 32 != 0 || L.cout < env_int("MC_TC3_MIN_COUT", split ? 64 : 128)) return false;
    if (const char* skip = std::getenv("MC_TC3_SKIP"))
        if (skip[0] && L.name.find(skip) != std::string::npos) return false;
    const TensorInfo& d = net.tensors[L.dst];
    if (d.W % 8 != 0) return false;
    if (L.residual >= 0 && net.tensors[L.residual].dt != d.dt) return false;
    int nch = 0;
    for (int s : L.src) {
        const TensorInfo& t = net.tensors[s];
        if (t.C % 64 != 0 || t.Wp != t.W) return false;
        if (t.dt != net.dt) return false;
        nch += t.C / 64;
    }
    if (nch * (split ? 3 : 1) > kMaxChunks3) return false;
    Tc3Params& p = plan.p;
    std::memset(&p, 0, sizeof(p));
    // fp32-accurate mode: x * w ~= hi * w_lo + lo * w_hi + hi * w_hi.  The two cross terms come FIRST: the tensor core adds
    // into the TMEM accumulator with truncation (tools/umma_accum.cu), an error relative to the accumulator's magnitude
    // per MMA, so the small terms are summed while the accumulator is still small.
    for (int pass = split ? 0 : 2; pass < 3; ++pass)
        for (int si = 0; si < (int)L.src.size(); ++si)
            for (int c0 = 0; c0 < net.tensors[L.src[si]].C; c0 += 64) p.chunks[p.nchunks++] = Chunk3{si, c0, pass == 1 ? 1 : 0};
    p.f16 = split ? 1 : 0;
    p.plane_imgs = net.max_batch;
    plan.om = L.dst_override_f32 ? tcepi::OM_F32 : (d.dt == DT_SPLIT ? tcepi::OM_SPLIT : (d.dt == DT_F32 ? tcepi::OM_F32 : tcepi::OM_BF16));
    if (!split && plan.om == tcepi::OM_SPLIT) return false;
    p.H = d.H; p.W = d.W; p.B = net.max_batch; p.Cout = L.cout; p.Hp = d.H + 2;
    p.strips = d.W / 8;
    // Cout tile: 128 columns (two sub-tiles per step, double-buffered accumulators).  256 columns (single-buffered, half the
    // weight traffic per MMA) was 5-10 % faster on level4 / ida_0.node before the second epilogue group existed; with two
    // groups 128 wins there too (level4 0.0415 -> 0.039 ms).  MC_TC3_NT=256 selects the wide tile.
    const int n_tile_dflt = 128;
    int n_tile = std::min(L.cout, env_int("MC_TC3_NT", n_tile_dflt));
    while (L.cout % n_tile != 0) n_tile -= 16;
    if (n_tile < 64) return false;
    p.n_tile = n_tile;
    p.n_tiles = L.cout / n_tile;
    // two CTAs per SM (MC_TC3_CTAS=2, experiment): 256 TMEM columns and half the shared memory each, one sub-tile per step;
    // two issuing warps feed the tensor pipe and each CTA's epilogue / prologue / tail overlaps the other's MMAs
    const int ctas = (env_int("MC_TC3_CTAS", 1) == 2 && n_tile <= 128) ? 2 : 1;
    p.ctas_per_sm = ctas;
    p.tmem_cols = ctas == 2 ? 256 : 512;
    int sub = std::max(1, std::min(p.tmem_cols / n_tile, std::min(kMaxSub3, env_int("MC_TC3_SUB", 2))));
    if (ctas == 2) sub = 1;
    if (sub == 3) sub = 2;
    p.sub = sub;
    p.acc_stages = sub * n_tile <= p.tmem_cols / 2 ? 2 : 1;
    p.acc_stride = p.tmem_cols / 2;
    p.a_row_bytes = 10 * 128;
    if (env_int("MC_TC3_ROWPAD", 0)) p.a_row_bytes = 2048;
    const int rows = 16 * sub + 2;
    p.a_slot_stride = (rows * p.a_row_bytes + 1023) / 1024 * 1024;
    p.b_bytes = n_tile * 128;
    p.b_slot_stride = p.b_bytes;                              // multiple of 1024 (n_tile >= 64, multiple of 16 -> check)
    if (p.b_slot_stride % 1024 != 0) return false;
    const size_t fixed = 1024 + sizeof(float) * 2 * L.cout + 16 + 8 * (2 * kMaxASlots3 + 2 * kMaxBSlots3 + 4) + 16;
    // two CTAs per SM: 228 KB per SM minus 1 KB reserved per block, halved
    const size_t avail = (ctas == 2 ? (size_t)(228 * 1024 - 2 * 1024) / 2 : (size_t)g_max_smem3) - fixed;
    p.a_slots = 2;
    if ((size_t)p.a_slots * p.a_slot_stride + 3 * (size_t)p.b_slot_stride > avail) return false;
    p.b_slots = (int)std::min<size_t>(kMaxBSlots3, (avail - (size_t)p.a_slots * p.a_slot_stride) / p.b_slot_stride);
    // spare room goes to a third activation slot when at least four weight slots remain
    if (p.b_slots == kMaxBSlots3) {
        const size_t left = avail - (size_t)p.a_slots * p.a_slot_stride - (size_t)p.b_slots * p.b_slot_stride;
        p.a_slots = (int)std::min<size_t>(kMaxASlots3, p.a_slots + left / p.a_slot_stride);
    }
    plan.smem_bytes = fixed + (size_t)p.a_slots * p.a_slot_stride + (size_t)p.b_slots * p.b_slot_stride;
    return true;
}

bool tc3_conv_supported(const Net& net, const ConvLayer& L) {
    Tc3ConvPlan tmp;
    return plan_tc3(net, L, tmp);
}

void tc3_conv_prepare(Net& net, ConvLayer& L, const std::vector<float>& w_oihw) {
    auto plan = std::make_shared<Tc3ConvPlan>();
    MC_CHECK(plan_tc3(net, L, *plan), "tc3: layer not supported: " + L.name);
    Tc3Params& p = plan->p;
    const TensorInfo& d = net.tensors[L.dst];
    const int B = net.max_batch;
    // weights [chunk][tap][cout][64]; fp32-accurate mode: fp16 pieces of w * 2^ew[cout], the chunks in the kernel's order
    // (hi-plane x w_lo, lo-plane x w_hi, hi-plane x w_hi)
    const bool split = net.dt == DT_SPLIT;
    const int real_chunks = split ? p.nchunks / 3 : p.nchunks;
    const std::vector<int> ew = split ? split_weight_exponents(w_oihw, L.cout) : std::vector<int>();
    std::vector<uint16_t> w;
    w.reserve((size_t)p.nchunks * 9 * L.cout * 64);
    L.widx.clear();
    std::vector<int> cb;
    int cbase = 0;
    for (int s : L.src) { cb.push_back(cbase); cbase += net.tensors[s].C; }
    for (int ci = 0; ci < p.nchunks; ++ci) {
        const bool want_lo = split && ci < real_chunks;      // pass 0 pairs the hi plane with the weights' lo piece
        for (int j = 0; j < 9; ++j)
            for (int o = 0; o < L.cout; ++o)
                for (int kk = 0; kk < 64; ++kk) {
                    const int cin_idx = cb[p.chunks[ci].src] + p.chunks[ci].c + kk;
                    const size_t id = ((size_t)o * L.cin + cin_idx) * 9 + j;
                    const float v = w_oihw[id];
                    if (L.keep_widx) L.widx.push_back((int)id);
                    w.push_back(split ? split_weight_piece(v, ew[o], want_lo) : bf16_bits(v));
                }
    }
    plan->d_w = net.arena.alloc(sizeof(uint16_t) * w.size());
    MC_CUDA(cudaMemcpy(plan->d_w, w.data(), sizeof(uint16_t) * w.size(), cudaMemcpyHostToDevice));
    L.w_packed = plan->d_w;
    plan->d_err = (int*)net.arena.alloc(sizeof(int));
    p.error_flag = plan->d_err;
    for (int si = 0; si < kMaxSrc; ++si) {
        const TensorInfo& t = net.tensors[L.src[std::min(si, (int)L.src.size() - 1)]];
        const cuuint64_t C = t.C, W = t.W, H = t.H;
        cuuint64_t dims[5] = {C, W, 1, H, (cuuint64_t)B * (split ? 2 : 1)};      // DT_SPLIT: the lo plane = images B .. 2B-1
        cuuint64_t str[4] = {C * 2, W * C * 2, W * C * 2, H * W * C * 2};
        cuuint32_t box[5] = {64, 10, 1, 1, 1};
        encode3(&p.map_a[si], t.ptr, 5, dims, str, box, L.name + " (activation halo row)", split);
    }
    {
        cuuint64_t dims[2] = {64, (cuuint64_t)p.nchunks * 9 * L.cout};
        cuuint64_t str[1] = {128};
        cuuint32_t box[2] = {64, (cuuint32_t)p.n_tile};
        encode3(&p.map_b, plan->d_w, 2, dims, str, box, L.name + " (weights)", split);
    }
    p.scale = split ? net.upload_split_scale(L, ew) : L.scale;
    p.shift = L.shift;
    // dst_override: the RAW convolution output is wanted (a train-mode BatchNorm applies residual and ReLU afterwards)
    p.residual = (L.residual >= 0 && !L.dst_override) ? net.tensors[L.residual].ptr : nullptr;
    p.dst = L.dst_override ? L.dst_override : d.ptr;
    if (split) {
        p.in_sc = net.act_scale(L.src[0]);
        if (plan->om == tcepi::OM_SPLIT) { p.out_sc = net.act_scale(L.dst); p.amax = net.act_amax(L.dst); p.dst_plane = d.plane; }
        if (L.residual >= 0) { p.res_sc = net.act_scale(L.residual); p.res_plane = net.tensors[L.residual].plane; }
    }
    p.relu = (L.relu && !L.dst_override) ? 1 : 0;
    p.diag = env_int("MC_DIAG3", 0);
    L.tc3 = plan;
}

void tc3_conv_launch(const Net& net, const ConvLayer& L, int B, cudaStream_t st) {
    MC_CHECK(L.tc3 != nullptr, "tc3 conv not prepared: " + L.name);
    Tc3Params p = L.tc3->p;
    p.B = B;
    p.tiles_g = (B * p.Hp + 15) / 16;
    if (L.stats_sums != nullptr && L.tc3->om == tcepi::OM_F32) {
        MC_CUDA(cudaMemsetAsync(L.stats_sums, 0, sizeof(double) * 2 * (size_t)p.Cout * B, st));
        p.stats = L.stats_sums;
    }
    const int total = p.strips * p.tiles_g * p.n_tiles;
    // every CTA should own whole steps where possible: with fewer sub-tiles than sub * #SM, shrink the grid so that the
    // weight boxes are still shared (MC_TC3_FILL=1 spreads over all SMs instead)
    int grid = std::min(total, g_num_sms3 * p.ctas_per_sm);
    if (!env_int("MC_TC3_FILL", 1)) grid = std::max(1, std::min(grid, (total + p.sub - 1) / p.sub));
    const char* tl = std::getenv("MC_TRACE_LAYER");
    static unsigned long long* d_trace = nullptr;
    const bool trace = tl && L.name == tl;
    if (trace) {
        if (!d_trace) MC_CUDA(cudaMalloc(&d_trace, (16 + 4 * 256) * sizeof(unsigned long long)));
        MC_CUDA(cudaMemsetAsync(d_trace, 0, (16 + 4 * 256) * sizeof(unsigned long long), st));
        p.trace = d_trace;
    }
    const int eg = epi_groups3(p);
    launch_k(kernel3_for(p.sub, p.ctas_per_sm, eg, L.tc3->om), dim3(grid), dim3(kThreads3 + 128 * (eg - 1)), L.tc3->smem_bytes, st, p);
    if (trace) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(st, &cs);
        if (cs == cudaStreamCaptureStatusNone) {
            unsigned long long h[16 + 4 * 256];
            MC_CUDA(cudaStreamSynchronize(st));
            MC_CUDA(cudaMemcpy(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost));
            {
                // per-CTA timeline relative to the earliest CTA entry (ns): min / median / max over the grid
                unsigned long long t0 = ~0ull;
                for (int c = 0; c < grid; ++c) t0 = std::min(t0, h[16 + 4 * c]);
                const char* names[4] = {"entry", "mma-begin", "mma-end", "exit"};
                std::fprintf(stderr, "[timeline3 %s]", L.name.c_str());
                for (int k = 0; k < 4; ++k) {
                    std::vector<unsigned long long> v;
                    for (int c = 0; c < grid; ++c) v.push_back(h[16 + 4 * c + k] - t0);
                    std::sort(v.begin(), v.end());
                    std::fprintf(stderr, " %s %llu/%llu/%llu ns", names[k], v.front(), v[v.size() / 2], v.back());
                }
                std::vector<unsigned long long> d;
                for (int c = 0; c < grid; ++c) d.push_back(h[16 + 4 * c + 2] - h[16 + 4 * c + 1]);
                std::sort(d.begin(), d.end());
                std::fprintf(stderr, " | mma loop duration %llu/%llu/%llu ns\n", d.front(), d[d.size() / 2], d.back());
            }
            std::fprintf(stderr, "[trace3 %s] grid %d units %d chunks %d sub %d n_tile %d acc_stages %d a_slots %d b_slots %d | A producer: loop %llu clk, wait a_empty %llu | B producer: loop %llu, wait b_empty %llu | mma: loop %llu, wait tmem_empty %llu, a_full %llu, b_full %llu | epilogue: loop %llu, wait tmem_full %llu\n",
                         L.name.c_str(), grid, total, p.nchunks, p.sub, p.n_tile, p.acc_stages, p.a_slots, p.b_slots, h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8], h[9]);
        }
    }
}

}  // namespace mc
