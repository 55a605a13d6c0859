// umma_accum: how does tcgen05.mma accumulate?  (kind::f16 with fp16 operands, kind::tf32)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a tools/umma_accum.cu -o gpurun_out/umma_accum ; run on a B200.
//
// The fp32-accurate convolution mode splits every operand into two fp16 pieces (hi + lo) and issues three MMAs per
// K-block into ONE fp32 accumulator in TMEM.  Products of fp16 pieces are exact in fp32, so the result is as good as
// the accumulator's own arithmetic.  This probe measures that arithmetic: a chain of T MMAs (M = 128, N = 16, K = 16
// (f16) / 8 (tf32)) on random operands, compared with (a) the exact sum (fp64 on the host), (b) host models that add
// the exact per-MMA block sum to an fp32 accumulator with round-to-nearest / round-toward-zero.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <cmath>
#include <cfenv>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int kPool = 24;            // operand tiles kept in shared memory
constexpr int kATile = 128 * 32;     // bytes: 128 rows x 32 B (K = 16 fp16 or 8 tf32)
constexpr int kBTile = 16 * 32;
constexpr int kMaxT = 2048;

struct Seq { int T; int kind; /*0 f16, 1 tf32*/ };
__constant__ unsigned char c_ia[kMaxT];
__constant__ unsigned char c_ib[kMaxT];

__global__ void __launch_bounds__(128, 1) accum_kernel(Seq sq, const uint8_t* a_pool, const uint8_t* b_pool, float* out /*[128][16]*/) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sa = smem;
    uint8_t* sb = smem + kPool * kATile;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < kPool * kATile / 16; i += 128) reinterpret_cast<uint4*>(sa)[i] = reinterpret_cast<const uint4*>(a_pool)[i];
    for (int i = tid; i < kPool * kBTile / 16; i += 128) reinterpret_cast<uint4*>(sb)[i] = reinterpret_cast<const uint4*>(b_pool)[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmem_ptr)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_ptr;
    if (tid == 0) {
        // K-major, no swizzle: element (m, kbyte) at (kbyte / 16) * LBO + (m / 8) * SBO + (m % 8) * 16 + kbyte % 16; LBO = 128, SBO = 256
        const uint64_t hi = ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46);
        const uint32_t fmt = sq.kind == 0 ? 0u : 2u;      // f16 / tf32
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
        for (int t = 0; t < sq.T; ++t) {
            const uint64_t adesc = hi | (uint64_t)(((smem_u32(sa) + c_ia[t] * kATile) & 0x3FFFF) >> 4);
            const uint64_t bdesc = hi | (uint64_t)(((smem_u32(sb) + c_ib[t] * kBTile) & 0x3FFFF) >> 4);
            const uint32_t accf = t ? 1u : 0u;
            if (sq.kind == 0)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accf) : "memory");
            else
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accf) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    {
        uint32_t ok = 0;
        long long t0 = clock64();
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
            if (clock64() - t0 > 4000000000LL) { asm volatile("trap;"); }
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

// ---- host ----
static float to_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
static float rz_add(float acc, double blk) {
    // fp32 round-toward-zero of the exact (double) sum
    const double s = (double)acc + blk;
    float r = (float)s;                       // RN
    if (std::fabs((double)r) > std::fabs(s)) r = std::nextafterf(r, 0.f);
    return r;
}

struct Pools {
    std::vector<float> a, b;       // logical values [pool][rows][K]
    std::vector<uint8_t> ab, bb;   // device byte images
};
// kind 0: fp16, K = 16; kind 1: tf32, K = 8.  gen(which, pool, row, k) -> value
template <class G>
static Pools make_pools(int kind, G gen) {
    const int K = kind == 0 ? 16 : 8;
    Pools P;
    P.a.resize((size_t)kPool * 128 * K); P.b.resize((size_t)kPool * 16 * K);
    P.ab.assign((size_t)kPool * kATile, 0); P.bb.assign((size_t)kPool * kBTile, 0);
    for (int p = 0; p < kPool; ++p) {
        for (int which = 0; which < 2; ++which) {
            const int rows = which == 0 ? 128 : 16;
            for (int m = 0; m < rows; ++m)
                for (int k = 0; k < K; ++k) {
                    float v = gen(which, p, m, k);
                    const int esz = kind == 0 ? 2 : 4;
                    const int kb = k * esz;
                    const size_t off = (size_t)p * (which == 0 ? kATile : kBTile) + (kb / 16) * 128 + (m / 8) * 256 + (m % 8) * 16 + kb % 16;
                    uint8_t* dst = (which == 0 ? P.ab.data() : P.bb.data()) + off;
                    if (kind == 0) { __half h = __float2half_rn(v); v = __half2float(h); memcpy(dst, &h, 2); }
                    else { v = to_tf32(v); memcpy(dst, &v, 4); }
                    (which == 0 ? P.a : P.b)[((size_t)p * rows + m) * K + k] = v;
                }
        }
    }
    return P;
}

static void run_case(const char* name, int kind, int T, const Pools& P, uint8_t* d_a, uint8_t* d_b, float* d_out, uint64_t seed) {
    const int K = kind == 0 ? 16 : 8;
    std::vector<unsigned char> ia(kMaxT), ib(kMaxT);
    std::mt19937_64 rng(seed);
    for (int t = 0; t < T; ++t) { ia[t] = rng() % kPool; ib[t] = rng() % kPool; }
    cudaMemcpyToSymbol(c_ia, ia.data(), kMaxT);
    cudaMemcpyToSymbol(c_ib, ib.data(), kMaxT);
    cudaMemcpy(d_a, P.ab.data(), P.ab.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(d_b, P.bb.data(), P.bb.size(), cudaMemcpyHostToDevice);
    const size_t smem = (size_t)kPool * (kATile + kBTile) + 1024;
    accum_kernel<<<1, 128, smem>>>(Seq{T, kind}, d_a, d_b, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); exit(1); }
    std::vector<float> out(128 * 16);
    cudaMemcpy(out.data(), d_out, sizeof(float) * out.size(), cudaMemcpyDeviceToHost);
    double sum_rel = 0, sum_rel2 = 0, sum_ulp = 0, sum_ulp2 = 0, max_ulp = 0;
    double rn_ulp2 = 0, rz_ulp2 = 0, seq_ulp2 = 0;
    int match_rn = 0, match_rz = 0, match_exact = 0, n = 0;
    for (int m = 0; m < 128; ++m)
        for (int c = 0; c < 16; ++c) {
            double exact = 0;
            float acc_rn = 0.f, acc_rz = 0.f, acc_seq = 0.f;
            for (int t = 0; t < T; ++t) {
                double blk = 0;
                const float* ar = &P.a[((size_t)ia[t] * 128 + m) * K];
                const float* br = &P.b[((size_t)ib[t] * 16 + c) * K];
                for (int k = 0; k < K; ++k) { blk += (double)ar[k] * (double)br[k]; acc_seq = fmaf(ar[k], br[k], acc_seq); }
                exact += blk;
                acc_rn = (float)((double)acc_rn + blk);
                acc_rz = rz_add(acc_rz, blk);
            }
            const float got = out[m * 16 + c];
            const double ulp = std::ldexp(1.0, std::ilogb((float)exact == 0.f ? 1.f : (float)exact) - 23);
            const double err = (double)got - exact;
            const double rel = exact != 0 ? err / std::fabs(exact) : 0;
            const double sgn = exact >= 0 ? 1.0 : -1.0;     // signed toward larger magnitude
            sum_rel += rel * 1.0; sum_rel2 += rel * rel;
            sum_ulp += sgn * err / ulp; sum_ulp2 += (err / ulp) * (err / ulp);
            if (std::fabs(err / ulp) > max_ulp) max_ulp = std::fabs(err / ulp);
            rn_ulp2 += std::pow(((double)acc_rn - exact) / ulp, 2);
            rz_ulp2 += std::pow(((double)acc_rz - exact) / ulp, 2);
            seq_ulp2 += std::pow(((double)acc_seq - exact) / ulp, 2);
            match_rn += got == acc_rn; match_rz += got == acc_rz; match_exact += got == (float)exact;
            ++n;
        }
    printf("%-34s kind=%s T=%4d (K=%5d): mean signed err %+8.3f ulp (toward larger |x|)  rms %8.3f ulp  max %8.2f ulp  rel rms %.3e | "
           "host models rms: RN-per-MMA %.3f  RZ-per-MMA %.3f  fp32 FMA chain %.3f ulp | bit-equal: RN %d RZ %d RN(exact) %d of %d\n",
           name, kind == 0 ? "f16" : "tf32", T, T * K, sum_ulp / n, std::sqrt(sum_ulp2 / n), max_ulp, std::sqrt(sum_rel2 / n),
           std::sqrt(rn_ulp2 / n), std::sqrt(rz_ulp2 / n), std::sqrt(seq_ulp2 / n), match_rn, match_rz, match_exact, n);
}

int main() {
    uint8_t *d_a, *d_b; float* d_out;
    cudaMalloc(&d_a, (size_t)kPool * kATile); cudaMalloc(&d_b, (size_t)kPool * kBTile); cudaMalloc(&d_out, sizeof(float) * 128 * 16);
    const size_t smem = (size_t)kPool * (kATile + kBTile) + 1024;
    cudaFuncSetAttribute(accum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    std::mt19937_64 rng(1234);
    std::normal_distribution<float> nd(0.f, 1.f);
    std::uniform_real_distribution<float> ud(0.5f, 1.5f);
    for (int kind = 0; kind < 2; ++kind) {
        // 1. random signed operands (the convolution case: activations >= 0 after ReLU, weights signed)
        Pools Pn = make_pools(kind, [&](int which, int, int, int) { return which == 0 ? std::fabs(nd(rng)) : nd(rng) * 0.05f; });
        for (int T : {1, 2, 4, 16, 36, 108, 288, 864, 2048}) run_case("relu-act x signed-w", kind, T, Pn, d_a, d_b, d_out, 77 + T);
        // 2. all-positive products: a rounding bias shows up as a drift linear in T
        Pools Pp = make_pools(kind, [&](int, int, int, int) { return ud(rng); });
        for (int T : {1, 4, 16, 108, 864, 2048}) run_case("all-positive", kind, T, Pp, d_a, d_b, d_out, 99 + T);
        // 3. large accumulator + small addends: MMA 0 contributes ~2^10, the others ~1e-3 relative (the lo-piece MMAs of a split)
        Pools Ps = make_pools(kind, [&](int which, int p, int, int) {
            const float s = which == 0 ? ((p % 3 == 0) ? 1.f : 1.f / 2048.f) : 1.f;
            return s * (which == 0 ? std::fabs(nd(rng)) : nd(rng));
        });
        for (int T : {36, 108, 864}) run_case("mixed hi / lo magnitudes", kind, T, Ps, d_a, d_b, d_out, 5 + T);
    }
    // 4. micro tests (f16): one accumulator of 1.0, then addends below one ulp
    {
        // tile 0: A row = [1, 0...], tile 1: A row = [2^-12, 0, ...] with B col = [3 * 2^-13 ...]: product = 0.75 * 2^-23 = 0.75 ulp(1.0)
        // tile 2: A row = 16 x 2^-12, B = 2^-13 each: sixteen products of 2^-25 (1/4 ulp) -> block sum 4 ulp
        Pools Pm = make_pools(0, [&](int which, int p, int, int k) {
            if (which == 0) {
                if (p == 0) return k == 0 ? 1.f : 0.f;
                if (p == 1) return k == 0 ? std::ldexp(1.f, -12) : 0.f;
                if (p == 2) return std::ldexp(1.f, -12);
                return 0.f;
            }
            if (p == 0) return k == 0 ? 1.f : 0.f;
            if (p == 1) return k == 0 ? 3.f * std::ldexp(1.f, -13) : 0.f;   // x 2^-12 = 0.75 * 2^-23
            if (p == 2) return std::ldexp(1.f, -13);
            return 0.f;
        });
        cudaMemcpy(d_a, Pm.ab.data(), Pm.ab.size(), cudaMemcpyHostToDevice);
        cudaMemcpy(d_b, Pm.bb.data(), Pm.bb.size(), cudaMemcpyHostToDevice);
        auto go = [&](const char* what, std::vector<int> seq_a, std::vector<int> seq_b) {
            std::vector<unsigned char> ia(kMaxT, 3), ib(kMaxT, 3);
            for (size_t t = 0; t < seq_a.size(); ++t) { ia[t] = seq_a[t]; ib[t] = seq_b[t]; }
            cudaMemcpyToSymbol(c_ia, ia.data(), kMaxT); cudaMemcpyToSymbol(c_ib, ib.data(), kMaxT);
            accum_kernel<<<1, 128, smem>>>(Seq{(int)seq_a.size(), 0}, d_a, d_b, d_out);
            cudaDeviceSynchronize();
            float o; cudaMemcpy(&o, d_out, 4, cudaMemcpyDeviceToHost);
            printf("micro: %-60s -> 1 + %.4f ulp\n", what, ((double)o - 1.0) / std::ldexp(1.0, -23));
        };
        go("1.0 then 8 x (+0.75 ulp), one product per MMA [RN: +8, RZ: +0]", {0, 1, 1, 1, 1, 1, 1, 1, 1}, {0, 1, 1, 1, 1, 1, 1, 1, 1});
        go("1.0 then 1 x (16 products of 0.25 ulp in one MMA) [exact block: +4]", {0, 2}, {0, 2});
        go("1.0 then 4 x (16 products of 0.25 ulp) [exact blocks: +16]", {0, 2, 2, 2, 2}, {0, 2, 2, 2, 2});
    }
    printf("done\n");
    return 0;
}
