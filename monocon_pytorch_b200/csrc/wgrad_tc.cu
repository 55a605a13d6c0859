// tcgen05 weight-gradient kernel of the bf16 training step (BASELINE.json configs[2] / [4]) for sm_100a.
//
//   dW[tap][ci][co] += sum over pixels p of  dy[p][co] * x[p + tap][ci]          (nn.Conv2d backward w.r.t. the weight:
//   BasicBlock / Root / Tree.project dla.py:22-31,117-121,181-185; Conv2dBlock dla_neck.py:24-31; head stems monocon_heads.py:114-131)
//
// As a GEMM the reduction dimension K is the PIXEL index, so with NHWC activations both operands are "MN-major" (the
// channel index -- M for dy, N for x -- is the contiguous one).  tcgen05.mma reads MN-major bf16 operands directly
// (instruction-descriptor bits 15 / 16), and a TMA box [channels][8 px][rows] written with the 128 / 64 / 32-byte swizzle IS the
// canonical MN-major layout: one K index (pixel) per swizzle row, 8-pixel groups SBO apart.  Therefore
//   * the dy tile of a step (R rows x 8 pixels x <= 128 output channels) is one or two plain boxes,
//   * the x tile is ONE halo box ((R + 2) x 10 pixels x <= 64 input channels, out-of-bounds = the zero padding), and every
//     filter tap is the same tile seen through a different descriptor start address (+ (ky * 10 + kx) pixel rows), exactly
//     like the forward halo-view kernels (conv_tc2.cu) -- no im2col, no transposed copies of the activations,
//   * every tap owns its own accumulator D_tap[co][ci] (128 lanes x N columns) in TMEM; a CTA accumulates its whole pixel range
//     there and drains ONCE at the end with fp32 red.global.add into the master-layout gradient [tap][Cin][Cout] (split-K
//     over CTAs; a warp's 32 lanes are 32 adjacent output channels = one 128-byte reduction).
// Work items = (128-wide Cout tile) x (<= 64-wide Cin chunk of one source) x (tap group: up to 512 / N taps); each item's pixel
// tiles are split over floor(#SM / items) CTAs.  Stride-2 layers arrive here as stride-1 problems: their dy is stored
// zero-inserted at input resolution (train_tc.cu), which makes wgrad and dgrad ordinary 3x3 / stride-1 work at 4x the MMAs
// of five small layers.
//
//   warp 0: TMA producer   warp 1: MMA issuer + TMEM allocator   warps 2..5: drain (TMEM lane quarters 2, 3, 0, 1)
#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "train_tc.h"

namespace mc {

namespace {

constexpr int kWgThreads = 192;
constexpr int kWgMaxChunks = 24;
constexpr int kWgMaxStages = 6;
constexpr int kWgMaxViews = 9;
constexpr long long kWgSpinLimit = 4000000000LL;

struct WgChunk { int src, c, n; };       // source, first channel, channels (16 / 32 / 64, or 128 = two 64-channel boxes side by side)

struct WgParams {
    CUtensorMap map_dy;                  // dims {Cout, W, H, B}, box {mch, 8, R, 1}
    CUtensorMap map_x[kMaxSrc];          // dims {C, W, H, B}, box {nch, 8 + 2 pad, R + 2 pad, 1}
    WgChunk chunks[kWgMaxChunks];
    int nchunks;
    int H, W, B, Cout, Cin;              // Cin: all sources
    int k, pad;                          // 3 / 1 or 1 / 0; stem: 7 / 3
    int stem, xoff;                      // 7x7 stem over the padded 8-channel image (16-byte pixels, physical column = x + xoff)
    int xm;                              // x is the M operand, its MN blocks = horizontal taps (16 / 32 input channels; the stem when swapped)
    int m64;                             // M = 64 accumulators (rows r -> TMEM lanes (r % 16) + 32 (r / 16)): halves the A-operand reads
    int R;                               // tile rows (even)
    int tiles_x, tiles_y;
    int mch;                             // channels per dy box: min(Cout, 64)
    int co_tiles;                        // ceil(Cout / 128)
    int groups;                          // tap groups per (Cout tile, chunk)
    int taps_per_group;                  // first groups get this many, the last one the rest
    int ksplit;                          // CTAs per item
    int stages;
    int a_box_bytes, a_bytes, b_bytes_max, stage_stride;
    int cbase[kMaxSrc];                  // first input channel of each source in the concatenated Cin
    float* dw;                           // += [k*k][Cin][Cout]
    int* error_flag;
};

__device__ __forceinline__ uint32_t s32w(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32w(bar)), "r"(count));
}
__device__ __forceinline__ void wbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32w(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool wbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(s32w(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void wbar_wait(uint64_t* bar, uint32_t parity, int* error_flag, int code) {
    if (wbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!wbar_try_wait(bar, parity)) {
        if (clock64() - t0 > kWgSpinLimit) {
            if (error_flag) atomicExch(error_flag, code);
            __threadfence_system();
            asm volatile("trap;");
        }
    }
}
__device__ __forceinline__ void wtma4(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(s32w(smem)), "l"(map), "r"(s32w(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void wmma(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(tmem_d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void wcommit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32w(bar)) : "memory");
}
__device__ __forceinline__ bool welect() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void wtmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// UMMA shared-memory descriptor layout code of a swizzle row of `row_bytes`
__host__ __device__ inline uint32_t wg_layout(int row_bytes) { return row_bytes >= 128 ? 2u : (row_bytes == 64 ? 4u : 6u); }

// what a CTA works on
struct WgItem {
    int co0, m_valid, m_boxes;     // Cout tile
    int chunk;                     // input-channel chunk
    int tap0, ntaps;               // tap group
    int t0, t1;                    // pixel tiles
    __device__ __forceinline__ WgItem(const WgParams& p) {
        const int item = (int)blockIdx.x / p.ksplit, j = (int)blockIdx.x % p.ksplit;
        const int g = item % p.groups;
        const int rest = item / p.groups;
        chunk = rest % p.nchunks;
        const int ct = rest / p.nchunks;
        co0 = ct * 128;
        m_valid = min(128, p.Cout - co0);
        m_boxes = (m_valid + p.mch - 1) / p.mch;
        if (m_boxes > 2) m_boxes = 2;
        const int kk = p.stem ? 7 : (p.xm ? 3 : p.k * p.k);      // stem / xm: one view per filter row
        tap0 = g * p.taps_per_group;
        ntaps = min(p.taps_per_group, kk - tap0);
        const int tiles = p.B * p.tiles_y * p.tiles_x;
        const int per = tiles / p.ksplit, rem = tiles % p.ksplit;
        t0 = j * per + min(j, rem);
        t1 = t0 + per + (j < rem ? 1 : 0);
    }
};

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_tc_kernel(const __grid_constant__ WgParams p) {
    extern __shared__ __align__(1024) uint8_t wg_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(wg_smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * p.stage_stride + 1024);   // 1 KB slack: garbage-row reads
    uint64_t* full = bars;                       // [kWgMaxStages]
    uint64_t* empty = full + kWgMaxStages;       // [kWgMaxStages]
    uint64_t* done = empty + kWgMaxStages;       // [1]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const WgItem it(p);
    const WgChunk ch = p.chunks[it.chunk];

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) { wbar_init(&full[s], 1); wbar_init(&empty[s], 1); }
        wbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32w(tmem_ptr)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_ptr;
    pdl_sync();

    const int PW = p.stem ? 16 : 8 + 2 * p.pad;         // halo tile width in pixels (stem: 8 + 6, and one more so that N = 8 taps x 8 channels)
    const int x_boxes = ch.n > 64 ? ch.n / 64 : 1;      // N = 128: the x tile is two 64-channel halo boxes, LBO apart
    const int pbA = p.mch * 2, pbB = (ch.n > 64 ? 64 : ch.n) * 2;          // bytes per pixel row of the two tiles
    const int x_box_bytes = (p.R + 2 * p.pad) * PW * pbB;
    const int b_bytes = x_boxes * x_box_bytes;
    const int nview = p.xm ? p.Cout : (p.stem ? 64 : ch.n);   // accumulator columns per view

    if (warp == 0) {
        // ===================== producer =====================
        if (welect()) {
            int s = 0;
            uint32_t phase = 0;
            for (int t = it.t0; t < it.t1; ++t) {
                const int tx = t % p.tiles_x;
                const int r1 = t / p.tiles_x;
                const int ty = r1 % p.tiles_y, n = r1 / p.tiles_y;
                wbar_wait(&empty[s], phase ^ 1u, p.error_flag, 41);
                uint8_t* st = smem + (size_t)s * p.stage_stride;
                wbar_expect_tx(&full[s], (uint32_t)(it.m_boxes * p.a_box_bytes + b_bytes));
                for (int b = 0; b < it.m_boxes; ++b)
                    wtma4(st + (size_t)b * p.a_box_bytes, &p.map_dy, &full[s], it.co0 + b * p.mch, tx * 8, ty * p.R, n);
                for (int b = 0; b < x_boxes; ++b)
                    wtma4(st + p.a_bytes + (size_t)b * x_box_bytes, &p.map_x[ch.src], &full[s], ch.c + 64 * b, tx * 8 - p.pad + p.xoff, ty * p.R - p.pad, n);
                if (++s == p.stages) { s = 0; phase ^= 1u; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // kind::f16, fp32 accumulate, bf16 x bf16, both operands MN-major, N = chunk channels, M = 128
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(nview >> 3) << 17) | (((p.m64 ? 64u : 128u) >> 4) << 24);
        // descriptor halves: hi = SBO (distance of the two 8-pixel groups of a K = 16 step) | version 1 | swizzle;
        //                    lo = LBO (distance of the MN blocks of one swizzle row: the second 64 output channels) | address
        // dy tile (8-pixel rows, contiguous) and x halo tile (PW-pixel rows):
        const uint32_t dy_hi = (uint32_t)((8 * pbA) >> 4) | (1u << 14) | (wg_layout(pbA) << 29);
        // stem: unswizzled 16-byte pixels.  There the roles of the two offsets swap (canonical MN-major INTERLEAVE layout): SBO is
        // the distance of the 8-element MN blocks -- 16 bytes, i.e. MN block j is the pixel j further right = filter tap kx = j, so
        // ONE N = 64 MMA covers a whole filter row -- and LBO the distance of the two 8-pixel groups (the tile's row pitch)
        const uint32_t x_hi = p.stem ? ((16u >> 4) | (1u << 14)) : ((uint32_t)((PW * pbB) >> 4) | (1u << 14) | (wg_layout(pbB) << 29));
        // rows of the accumulator beyond the tile's real output channels read shifted copies of the tile (LBO = one pixel row):
        // garbage in rows nobody drains
        const uint32_t dy_lbo = (it.m_boxes == 2) ? (uint32_t)p.a_box_bytes : 128u;
        const uint32_t dy_lo0 = ((dy_lbo >> 4) << 16);
        // xm (16 / 32 input channels): x is the M operand and LBO = ONE PIXEL, so MN block j of the 128 accumulator rows is the tile
        // shifted j pixels to the right = horizontal tap kx = j (j >= 3: garbage rows): one MMA per filter ROW instead of one per tap,
        // D_ky[(kx, ci)][co], and the dy tile is the N operand
        const uint32_t x_lo0 = p.stem ? ((uint32_t)((PW * 16) >> 4) << 16)
                             : (p.xm ? ((uint32_t)(pbB >> 4) << 16) : (x_boxes == 2 ? ((uint32_t)(x_box_bytes >> 4) << 16) : (1u << 16)));
        const uint32_t a_hi = p.xm ? x_hi : dy_hi, b_hi = p.xm ? dy_hi : x_hi;
        const int ksteps = p.R / 2;
        // The issuing thread must do next to nothing between two MMAs (the tensor pipe's queue is shallow, tools/umma_timing.cu):
        // per-view descriptor offsets live in registers (fully unrolled view loop), per K-step increments are single adds.
        uint32_t voff[kWgMaxViews];
#pragma unroll
        for (int v = 0; v < kWgMaxViews; ++v) {
            const int tap = it.tap0 + min(v, it.ntaps - 1);
            const bool row_view = p.stem || p.xm;                                          // view = filter row
            const int ky = row_view ? tap : tap / p.k, kx = row_view ? 0 : tap - ky * p.k;
            voff[v] = (uint32_t)((ky * PW + kx) * pbB) >> 4;
        }
        const uint32_t dy_step = (uint32_t)(16 * pbA) >> 4, x_step = (uint32_t)(2 * PW * pbB) >> 4;
        const bool xm = p.xm != 0;
        const uint32_t nv = (uint32_t)nview;
        const int ntaps = it.ntaps;
        // ONE elected lane runs the whole tile loop, barrier waits included (an elect + warp sync per tile is idle tensor time); the
        // state of the next stage's barrier is probed before the current tile's MMAs are issued, so its latency hides behind them
        if (welect()) {
            int s = 0;
            uint32_t phase = 0;
            bool ready = wbar_try_wait(&full[0], 0u);
            for (int t = it.t0; t < it.t1; ++t) {
                if (!ready) wbar_wait(&full[s], phase, p.error_flag, 42);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                int sn = s + 1;
                uint32_t pn = phase;
                if (sn == p.stages) { sn = 0; pn ^= 1u; }
                ready = (t + 1 < it.t1) ? wbar_try_wait(&full[sn], pn) : false;
                const uint32_t a_addr = s32w(smem + (size_t)s * p.stage_stride);
                uint32_t dylo = dy_lo0 | ((a_addr & 0x3FFFF) >> 4);
                uint32_t xlo = x_lo0 | (((a_addr + (uint32_t)p.a_bytes) & 0x3FFFF) >> 4);
                uint32_t acc = (t == it.t0) ? 0u : 1u;
                for (int ks = 0; ks < ksteps; ++ks) {
                    if (!xm) {
#pragma unroll
                        for (int v = 0; v < kWgMaxViews; ++v)
                            if (v < ntaps) wmma(tmem_base + (uint32_t)v * nv, dylo, a_hi, xlo + voff[v], b_hi, idesc, acc);
                    } else {
                        // exactly 3 (a 3x3 filter's rows) or 7 (the stem's) views: no predicated slots in the issue loop (they cost
                        // level0 / level1 0.3 ms each)
#pragma unroll
                        for (int v = 0; v < 3; ++v) wmma(tmem_base + (uint32_t)v * nv, xlo + voff[v], a_hi, dylo, b_hi, idesc, acc);
                        if (ntaps == 7) {
#pragma unroll
                            for (int v = 3; v < 7; ++v) wmma(tmem_base + (uint32_t)v * nv, xlo + voff[v], a_hi, dylo, b_hi, idesc, acc);
                        }
                    }
                    acc = 1u;
                    dylo += dy_step;
                    xlo += x_step;
                }
                wcommit(&empty[s]);
                if (t == it.t1 - 1) wcommit(done);
                s = sn; phase = pn;
            }
        }
        __syncwarp();
    } else if (it.t1 > it.t0) {
        // ===================== drain: TMEM -> red.global.add.f32 =====================
        const int q = warp & 3;
        const int m = q * 32 + lane;
        const bool row_ok = m < it.m_valid;
        wbar_wait(done, 0u, p.error_flag, 43);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int ci0 = p.cbase[ch.src] + ch.c;
        if (p.xm) {
            // accumulator row r = (kx, ci), column = co, view = ky.  M = 64: row r sits in TMEM lane (r % 16) + 32 (r / 16), i.e. the
            // first 16 lanes of every lane quarter (the "half subpartition" layout of tcgen05.mma with M = 64)
            const int r = p.m64 ? (lane < 16 ? q * 16 + lane : -1) : m;
            const int kx = r >= 0 ? r / ch.n : 99, ci = r - kx * ch.n;
            const int kw = p.stem ? 7 : 3;                  // filter width
            for (int v = 0; v < it.ntaps; ++v) {
                float* out = p.dw + ((size_t)(v * kw + kx) * p.Cin + ci0 + ci) * p.Cout;
                for (int c0 = 0; c0 < p.Cout; c0 += 16) {
                    uint32_t rr[16];
                    wtmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(v * nview + c0), rr);
                    if (kx < kw) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) atomicAdd(out + c0 + j, __uint_as_float(rr[j]));
                    }
                }
            }
        } else
        for (int v = 0; v < it.ntaps; ++v) {
            const int tap = it.tap0 + v;
            // stem: view = filter row ky, column = kx * 8 + channel, i.e. dw[(ky * 7 + kx) * 8 + c] is contiguous in the column index
            float* out = p.dw + ((size_t)(p.stem ? tap * 7 : tap) * p.Cin + ci0) * p.Cout + it.co0 + m;
            const int ncol = p.stem ? 56 : ch.n;
            for (int c0 = 0; c0 < nview; c0 += 16) {
                uint32_t r[16];
                wtmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(v * nview + c0), r);
                if (row_ok) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c0 + j < ncol) atomicAdd(out + (size_t)(c0 + j) * p.Cout, __uint_as_float(r[j]));
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
    }
}

int env_wg(const char* name, int dflt) {
    const char* e = std::getenv(name);
    return (e && e[0]) ? std::atoi(e) : dflt;
}

typedef CUresult (*EncodeTiledFnW)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFnW g_encode_w = nullptr;
int g_num_sms_w = 148;
int g_max_smem_w = 0;

void encode_w(CUtensorMap* map, const void* base, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box, int row_bytes,
              const std::string& what) {
    MC_CHECK(g_encode_w != nullptr, "cuTensorMapEncodeTiled entry point not resolved");
    cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUtensorMapSwizzle sw = row_bytes >= 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : (row_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE));
    CUresult r = g_encode_w(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ") for " + what);
}

}  // namespace

struct WgradPlan {
    WgParams p;
    int* d_err = nullptr;
    size_t smem_bytes = 0;
    int items = 0;
};

void wgrad_tc_init() {
    int dev = 0;
    MC_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    MC_CUDA(cudaGetDeviceProperties(&prop, dev));
    g_num_sms_w = std::max(1, prop.multiProcessorCount - reserved_sms());
    g_max_smem_w = (int)prop.sharedMemPerBlockOptin;
    if (!g_encode_w) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        MC_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        MC_CHECK(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available in this driver");
        g_encode_w = reinterpret_cast<EncodeTiledFnW>(fn);
    }
    MC_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem_w));
}

bool wgrad_tc_supported(const WgradDesc& d) {
    if (d.k == 7)            // the stem: one padded 8-channel source (3 colour channels + 5 zeros), 4 spare columns left and right
        return d.nsrc == 1 && d.src[0].C == 8 && d.src[0].xoff >= 3 && d.src[0].Wp >= d.W + d.src[0].xoff + 4 && d.W % 8 == 0 &&
               (d.Cout == 16 || d.Cout == 32 || d.Cout % 64 == 0);
    // any H, W: tiles that stick out of the image read zeros (TMA out-of-bounds fill) on both operands
    if (!(d.k == 3 || d.k == 1) || d.nsrc < 1 || d.nsrc > kMaxSrc) return false;
    if (!(d.Cout == 16 || d.Cout == 32 || d.Cout % 64 == 0)) return false;
    int chunks = 0;
    for (int s = 0; s < d.nsrc; ++s) {
        const int C = d.src[s].C;
        if (!(C == 16 || C == 32 || C % 64 == 0)) return false;
        if (d.src[s].Wp != 0 && (d.src[s].Wp != d.W || d.src[s].xoff != 0)) return false;
        chunks += (C + 63) / 64;
    }
    return chunks <= kWgMaxChunks;
}

std::shared_ptr<WgradPlan> wgrad_tc_prepare(const WgradDesc& d, int max_batch, DeviceArena& arena, const std::string& name) {
    MC_CHECK(wgrad_tc_supported(d), "wgrad_tc: layer not supported: " + name);
    auto plan = std::make_shared<WgradPlan>();
    WgParams& p = plan->p;
    std::memset(&p, 0, sizeof(p));
    p.H = d.H; p.W = d.W; p.B = max_batch; p.Cout = d.Cout; p.k = d.k; p.pad = (d.k - 1) / 2;
    p.stem = d.k == 7 ? 1 : 0;
    p.xoff = p.stem ? d.src[0].xoff : 0;
    bool wide = d.k == 3 && env_wg("MC_WGRAD_N128", 1) != 0;
    for (int s = 0; s < d.nsrc; ++s) wide = wide && d.src[s].C % 128 == 0;
    // tile rows: the largest even divisor of H up to 16 (H = 24 -> 12, no padded rows); 16 with zero-filled rows otherwise
    p.R = std::min(16, (d.H + 1) & ~1);
    for (int r = 16; r >= 8; r -= 2)
        if (d.H % r == 0) { p.R = r; break; }
    if (wide && d.H % 8 == 0) p.R = 8;            // 42 KB per stage instead of 78: three stages within the shared-memory budget
    p.tiles_x = (d.W + 7) / 8;
    p.tiles_y = (d.H + p.R - 1) / p.R;
    p.mch = std::min(d.Cout, 64);
    p.co_tiles = (d.Cout + 127) / 128;
    int nmax = 0, cin = 0;
    for (int s = 0; s < d.nsrc; ++s) {
        p.cbase[s] = cin;
        // 128-channel chunks (N = 128 MMAs: 8 KB of operands per 64 tensor clocks instead of 6 KB per 32) where every source allows it
        const int C = d.src[s].C, n = wide ? 128 : std::min(C, 64);
        for (int c0 = 0; c0 < C; c0 += n) p.chunks[p.nchunks++] = WgChunk{s, c0, n};
        nmax = std::max(nmax, n);
        cin += C;
    }
    p.Cin = cin;
    // 3x3 over one source of 16 / 32 channels with at most 64 output channels (level0, level1, level2.tree1.conv1): see xm in the kernel
    p.xm = (!p.stem && d.k == 3 && d.nsrc == 1 && (d.src[0].C == 16 || d.src[0].C == 32) && d.Cout <= 64 && env_wg("MC_WGRAD_XM", 1)) ? 1 : 0;
    // the stem the same way round (MC_WGRAD_STEM_SWAP, default on): x (unswizzled 16-byte pixels, MN blocks = the pixels to the right =
    // kx) is the M operand, dy the N operand, one view per filter row: 7 MMAs of M = 64 x N = Cout per K step instead of 7 of 128 x 64
    if (p.stem && d.Cout <= 64 && env_wg("MC_WGRAD_STEM_SWAP", 1)) p.xm = 1;
    // M = 64 where 64 accumulator rows hold every tap: 4 blocks of 16 channels (3 taps), or the stem's 8 blocks of 8 (7 taps)
    p.m64 = (p.xm && (p.stem || d.src[0].C == 16) && env_wg("MC_WGRAD_M64", 1)) ? 1 : 0;
    const int kk = p.stem ? 7 : (p.xm ? 3 : d.k * d.k);          // stem / xm: one view per filter row
    if (p.stem) nmax = 64;
    if (p.xm) nmax = d.Cout;                      // (after the stem's 64: a swapped stem has Cout columns per view)
    p.taps_per_group = std::min(kk, std::min(kWgMaxViews, 512 / nmax));
    p.groups = (kk + p.taps_per_group - 1) / p.taps_per_group;
    // balance the groups (9 taps at N = 64: 5 + 4 rather than 8 + 1)
    p.taps_per_group = (kk + p.groups - 1) / p.groups;
    plan->items = p.co_tiles * p.nchunks * p.groups;
    const int PW = p.stem ? 16 : 8 + 2 * p.pad;
    if (p.stem) nmax = 8;
    if (p.xm) nmax = d.src[0].C;
    p.a_box_bytes = p.R * 8 * p.mch * 2;
    p.a_bytes = (std::min(2, (std::min(128, d.Cout) + p.mch - 1) / p.mch) * p.a_box_bytes + 1023) / 1024 * 1024;
    p.b_bytes_max = ((p.R + 2 * p.pad) * PW * nmax * 2 + 1023) / 1024 * 1024;      // nmax = 128: two boxes
    p.stage_stride = p.a_bytes + p.b_bytes_max;
    const size_t fixed = 1024 /*alignment*/ + 1024 /*slack*/ + 8 * (2 * kWgMaxStages + 1) + 16;
    // MC_WGRAD_STAGES (default 3: ~170 KB for the widest tiles): leaves shared memory for blocks of the bandwidth kernels that run
    // next to this one (the engine launches the weight gradients on a stream of their own)
    p.stages = (int)std::min<size_t>(std::min(kWgMaxStages, env_wg("MC_WGRAD_STAGES", 3)), ((size_t)g_max_smem_w - fixed) / p.stage_stride);
    MC_CHECK(p.stages >= 2, "wgrad_tc: tile does not fit twice into shared memory: " + name);
    plan->smem_bytes = fixed + (size_t)p.stages * p.stage_stride;
    p.dw = d.dw;
    plan->d_err = (int*)arena.alloc(sizeof(int));
    p.error_flag = plan->d_err;
    {
        const cuuint64_t C = d.Cout, W = d.W, H = d.H;
        cuuint64_t dims[4] = {C, W, H, (cuuint64_t)max_batch};
        cuuint64_t str[3] = {C * 2, W * C * 2, H * W * C * 2};
        cuuint32_t box[4] = {(cuuint32_t)p.mch, 8, (cuuint32_t)p.R, 1};
        encode_w(&p.map_dy, d.dy, dims, str, box, p.mch * 2, name + " (wgrad dy)");
    }
    for (int s = 0; s < kMaxSrc; ++s) {
        const WgradSrc& src = d.src[std::min(s, d.nsrc - 1)];
        const cuuint64_t C = src.C, W = p.stem ? src.Wp : d.W, H = d.H;
        const int n = std::min(src.C, 64);
        cuuint64_t dims[4] = {C, W, H, (cuuint64_t)max_batch};
        cuuint64_t str[3] = {C * 2, W * C * 2, H * W * C * 2};
        cuuint32_t box[4] = {(cuuint32_t)n, (cuuint32_t)PW, (cuuint32_t)(p.R + 2 * p.pad), 1};
        encode_w(&p.map_x[s], src.x, dims, str, box, n * 2, name + " (wgrad x)");
    }
    return plan;
}

void wgrad_tc_launch(const WgradPlan& plan, int B, cudaStream_t st) {
    WgParams p = plan.p;
    p.B = B;
    const int tiles = B * p.tiles_y * p.tiles_x;
    p.ksplit = std::max(1, std::min(tiles, g_num_sms_w / plan.items));
    launch_k(wgrad_tc_kernel, dim3(plan.items * p.ksplit), dim3(kWgThreads), plan.smem_bytes, st, p);
}

}  // namespace mc
