// umma_probe: reveals which shared-memory bytes tcgen05.mma fetches for a given K-major matrix descriptor
// (start address, SBO, LBO, base_offset, swizzle mode).  Build: nvcc -gencode arch=compute_100a,code=sm_100a
// tools/umma_probe.cu -o gpurun_out/umma_probe ; run on a B200.
//
// Method: shared memory is filled so that every 16-byte chunk c holds the value (c & 255) [pass 0] or
// (c >> 8) [pass 1] in all eight bf16 lanes (exactly representable).  B = [I_16 ; 0]: D[m][n] = A[m][k = n].
// So D[m][0] / D[m][8] decode the chunk index the hardware read for row m, K-elements 0-7 / 8-15.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

typedef __nv_bfloat16 bf16;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Probe { int start_off; int sbo; int lbo; int base_offset; int layout; };   // bytes; layout: 0 none, 2 sw128, 4 sw64, 6 sw32

constexpr int kRegion = 48 * 1024;     // probed A region (bytes)

__global__ void __launch_bounds__(128, 1) probe_kernel(Probe pr, int pass, float* out /*[128][16]*/) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    bf16* a = reinterpret_cast<bf16*>(smem);                     // kRegion bytes
    bf16* b = reinterpret_cast<bf16*>(smem + kRegion);           // 16 rows x 16 k, no swizzle: 2 core matrices (8x16B) per k-chunk
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int c = tid; c < kRegion / 16; c += 128) {
        const float v = pass == 0 ? (float)(c & 255) : (float)(c >> 8);
        for (int j = 0; j < 8; ++j) a[c * 8 + j] = __float2bfloat16(v);
    }
    // B (N = 16 rows, K = 16), K-major, no swizzle: core matrix (8 rows x 16 B) contiguous 128 B;
    // element (n, k): chunk kc = k / 8 -> offset kc * 256 (LBO) + (n / 8) * 128 (SBO) + (n % 8) * 16 + (k % 8) * 2
    for (int i = tid; i < 16 * 16; i += 128) {
        const int n = i / 16, k = i % 16;
        const int off = (k / 8) * 256 + (n / 8) * 128 + (n % 8) * 16 + (k % 8) * 2;
        b[off / 2] = __float2bfloat16(n == k ? 1.f : 0.f);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the async proxy (UMMA)
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmem_ptr)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_ptr;
    if (tid == 0) {
        const uint32_t a_addr = smem_u32(a) + pr.start_off;
        const uint64_t adesc = (uint64_t)((a_addr & 0x3FFFF) >> 4) | ((uint64_t)(pr.lbo >> 4) << 16) | ((uint64_t)(pr.sbo >> 4) << 32) |
                               (1ull << 46) | ((uint64_t)(pr.base_offset & 7) << 49) | ((uint64_t)pr.layout << 61);
        const uint32_t b_addr = smem_u32(b);
        const uint64_t bdesc = (uint64_t)((b_addr & 0x3FFFF) >> 4) | ((uint64_t)(256 >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(0u)
            : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // wait
    {
        uint32_t ok = 0;
        long long t0 = clock64();
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
            if (clock64() - t0 > 2000000000LL) { asm volatile("trap;"); }
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[16];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 16 + j] = __uint_as_float(v[j]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem));
}

static void run(const char* name, Probe pr, float* d_out) {
    std::vector<float> h0(128 * 16), h1(128 * 16);
    const size_t smem = kRegion + 1024 + 2048;
    for (int pass = 0; pass < 2; ++pass) {
        probe_kernel<<<1, 128, smem>>>(pr, pass, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); exit(1); }
        cudaMemcpy(pass == 0 ? h0.data() : h1.data(), d_out, sizeof(float) * 128 * 16, cudaMemcpyDeviceToHost);
    }
    printf("== %s: start_off=%d sbo=%d lbo=%d base_offset=%d layout=%d\n", name, pr.start_off, pr.sbo, pr.lbo, pr.base_offset, pr.layout);
    // print the byte offset (relative to the A region base) of the chunk read for k=0..7 and k=8..15, rows 0..23 and a few more
    const int rows[] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 23, 24, 31, 32, 64, 127};
    for (int r : rows) {
        const int c0 = (int)h0[r * 16 + 0] + 256 * (int)h1[r * 16 + 0];
        const int c1 = (int)h0[r * 16 + 8] + 256 * (int)h1[r * 16 + 8];
        // sanity: all 8 elements of a chunk agree
        bool ok = true;
        for (int j = 1; j < 8; ++j) ok = ok && h0[r * 16 + j] == h0[r * 16] && h0[r * 16 + 8 + j] == h0[r * 16 + 8];
        printf("  row %3d: k0-7 @ %6d  k8-15 @ %6d%s\n", r, c0 * 16, c1 * 16, ok ? "" : "  (mixed chunk!)");
    }
}

int main() {
    float* d_out;
    cudaMalloc(&d_out, sizeof(float) * 128 * 16);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRegion + 1024 + 2048);
    // canonical SW128, aligned
    run("sw128 canonical", Probe{0, 1024, 16, 0, 2}, d_out);
    // SW128, k-advance +32 B
    run("sw128 +32B", Probe{32, 1024, 16, 0, 2}, d_out);
    // SW128, start shifted by 1 / 2 / 3 rows (128 B each), SBO = 1280 (halo pitch of 10 pixels)
    run("sw128 row+1 sbo1280 bo0", Probe{128, 1280, 16, 0, 2}, d_out);
    run("sw128 row+1 sbo1280 bo1", Probe{128, 1280, 16, 1, 2}, d_out);
    run("sw128 row+2 sbo1280 bo0", Probe{256, 1280, 16, 0, 2}, d_out);
    run("sw128 row+11 sbo1280 bo0", Probe{11 * 128, 1280, 16, 0, 2}, d_out);
    run("sw128 row+11 sbo1280 bo3", Probe{11 * 128, 1280, 16, 3, 2}, d_out);
    // SW64 / SW32 canonical and shifted
    run("sw64 canonical", Probe{0, 512, 16, 0, 4}, d_out);
    run("sw64 row+1 sbo640", Probe{64, 640, 16, 0, 4}, d_out);
    run("sw32 canonical", Probe{0, 256, 16, 0, 6}, d_out);
    run("sw32 row+1 sbo320", Probe{32, 320, 16, 0, 6}, d_out);
    // no swizzle: canonical (core matrices 128 B contiguous; LBO = k-chunk stride, SBO = 8-row-group stride)
    run("none canonical lbo128 sbo256", Probe{0, 256, 128, 0, 0}, d_out);
    // no swizzle Toeplitz: rows 16 B apart (overlapping windows): LBO = 16, SBO = 128
    run("none toeplitz lbo16 sbo128", Probe{0, 128, 16, 0, 0}, d_out);
    run("none toeplitz lbo16 sbo128 +48B", Probe{48, 128, 16, 0, 0}, d_out);
    printf("done\n");
    return 0;
}
