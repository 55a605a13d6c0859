// umma_timing: where does the time of a tcgen05 issue loop go?  One CTA, one issuing warp, operands resident in
// shared memory (contents irrelevant), clock64 around every group of MMAs and every tcgen05.commit.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a tools/umma_timing.cu -o gpurun_out/umma_timing ; run on a B200.
//
// For N in {64, 128, 256} and n MMAs per group in {4, 8, 36}:
//   mode 0: [n MMAs + commit(bar[g % 8])] x G, never waiting inside the loop       (the conv kernels' pattern)
//   mode 1: [n MMAs] x G, one commit at the end                                    (no per-group commits)
//   mode 2: like mode 0, but each group first waits for the commit of group g - 2  (what a 2-slot ring does)
// Printed: clocks per group (issue side), clocks until the last commit's barrier flips, and the ideal tensor time.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    const long long t0 = clock64();
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(s32(bar)), "r"(parity) : "memory");
        if (clock64() - t0 > 2000000000LL) asm volatile("trap;");
    }
}

struct Res { long long issue, total, commit_clk, copies; };

// a_sbo / a_off: stride between 8-row groups and start offset of the A view (1024 / 0 = canonical tile; 1280 / 128 =
// a filter-tap view of a 10-pixel-wide halo tile as used by conv_tc2.cu / conv_tc3.cu).  bg != 0: warp 1 streams
// 16 KB bulk copies global -> shared (L2-resident source) for the whole measurement (TMA write traffic into smem).
__global__ void __launch_bounds__(64, 1) timing_kernel(int N, int n_per_group, int G, int mode, int a_sbo, int a_off, int bg,
                                                       const uint8_t* gsrc, Res* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bars[9];
    __shared__ uint64_t bg_bar;
    __shared__ uint32_t tmem_ptr;
    __shared__ volatile int stop_flag;
    __shared__ long long bg_copies;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (49152 + 32768) / 4; i += 64) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 9; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[i])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bg_bar)));
        stop_flag = 0;
        bg_copies = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s32(&tmem_ptr)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_ptr;
    if (warp == 0) {
        // A: 128 rows x 64 bf16 (128 B rows, SWIZZLE_128B), B: N rows x 64 bf16
        const uint32_t a_addr = s32(smem) + (uint32_t)a_off, b_addr = s32(smem + 49152);
        const uint64_t hi = ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
        const uint64_t ahi = ((uint64_t)(a_sbo >> 4) << 32) | (1ull << 46) | (2ull << 61);
        const uint64_t adesc = (uint64_t)((a_addr & 0x3FFFF) >> 4) | (1ull << 16) | ahi;
        const uint64_t bdesc = (uint64_t)((b_addr & 0x3FFFF) >> 4) | (1ull << 16) | hi;
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        uint32_t phase[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        long long commit_clk = 0;
        __syncwarp();
        const long long t0 = clock64();
        for (int g = 0; g < G; ++g) {
            if (mode == 2 && g >= 2) {
                const int s = (g - 2) & 7;
                wait(&bars[s], phase[s]);
                phase[s] ^= 1u;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            if (elect_one()) {
                for (int i = 0; i < n_per_group; ++i) mma(tmem, adesc + 2u * (i & 3), bdesc + 2u * (i & 3), idesc, (g | i) ? 1u : 0u);
                if (mode != 1) {
                    const long long c0 = clock64();
                    commit(&bars[g & 7]);
                    commit_clk += clock64() - c0;
                }
            }
            __syncwarp();
        }
        const long long t1 = clock64();
        // drain: one more commit on a dedicated barrier tracks every MMA issued above
        if (elect_one()) commit(&bars[8]);
        __syncwarp();
        wait(&bars[8], 0u);
        const long long t2 = clock64();
        commit_clk = __shfl_sync(0xffffffffu, commit_clk, 0) ;
        stop_flag = 1;
        if (lane == 0) { out->issue = t1 - t0; out->total = t2 - t0; out->commit_clk = commit_clk; }
    } else if (bg) {
        // background bulk copies into a scratch region behind A and B
        uint8_t* scratch = smem + 49152 + 32768;
        uint32_t ph = 0;
        long long n = 0;
        while (!stop_flag) {
            if (lane == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bg_bar)), "r"(65536u) : "memory");
                for (int q = 0; q < 4; ++q)      // four copies in flight (same scratch region: only the traffic matters)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(s32(scratch)), "l"(gsrc + ((4 * n + q) & 63) * 16384), "r"(16384u), "r"(s32(&bg_bar)) : "memory");
            }
            __syncwarp();
            wait(&bg_bar, ph);
            ph ^= 1u;
            ++n;
        }
        if (lane == 0) bg_copies = n;
    }
    __syncthreads();
    if (threadIdx.x == 0) out->copies = bg_copies;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
    }
}

int main() {
    Res* d;
    uint8_t* gsrc;
    cudaMalloc(&d, sizeof(Res));
    cudaMalloc(&gsrc, 64 * 16384);
    cudaMemset(gsrc, 0, 64 * 16384);
    const int smem_bytes = 49152 + 32768 + 16384 + 2048;
    cudaFuncSetAttribute(timing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    const int G = 64;
    auto run = [&](int N, int n, int mode, int sbo, int off, int bg) {
        Res h{0, 0, 0, 0};
        for (int rep = 0; rep < 2; ++rep) {        // second run = warm instruction cache
            timing_kernel<<<1, 64, smem_bytes>>>(N, n, G, mode, sbo, off, bg, gsrc, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
            cudaMemcpy(&h, d, sizeof(Res), cudaMemcpyDeviceToHost);
        }
        printf("%5d %5d %4d %5d %4d %2d | %10.1f %10.1f %10.1f | %10.1f | %6.1f B/clk\n", N, n, mode, sbo, off, bg, (double)h.issue / G,
               (double)h.total / G, (double)h.commit_clk / G, (double)n * N / 2.0, (double)h.copies * 65536.0 / (double)h.total);
    };
    printf("%5s %5s %4s %5s %4s %2s | %10s %10s %10s | %10s | %s\n", "N", "n/grp", "mode", "sbo", "off", "bg", "issue/grp", "total/grp",
           "commit/grp", "ideal/grp", "bg copy rate");
    const int Ns[3] = {64, 128, 256}, ns[3] = {4, 8, 36};
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b)
            for (int mode = 0; mode < 3; ++mode) run(Ns[a], ns[b], mode, 1024, 0, 0);
    printf("-- A views of a 10-pixel halo tile (SBO 1280) and background bulk copies into shared memory\n");
    for (int a = 0; a < 3; ++a) {
        run(Ns[a], 36, 0, 1024, 0, 0);
        run(Ns[a], 36, 0, 1280, 0, 0);
        run(Ns[a], 36, 0, 1280, 128, 0);
        run(Ns[a], 36, 0, 1280, 1280 + 256, 0);
        run(Ns[a], 36, 0, 2048, 128, 0);
        run(Ns[a], 36, 0, 1024, 0, 1);
        run(Ns[a], 36, 0, 1280, 128, 1);
    }
    return 0;
}
