// TEST INFRASTRUCTURE ONLY.  Lets g++ compile monocon_pytorch_b200/csrc/train_backward.cu as host code (-DMC_HOST_SHIM) so that
// the arithmetic and indexing of its kernels can be checked in the GPU-less build container: every launch runs the kernel
// body once per (block, thread), sequentially, on host pointers.  That is only meaningful because those kernels are written
// without __syncthreads / shared memory / warp intrinsics (threads are independent and meet only in atomicAdd).
// The product library never sees this file: it is on the include path of tests/host_shim/build.sh alone, and the resulting
// tests/host_shim/_build/libtrain_backward_host.so is loaded by tests/test_backward_kernels_host.py alone.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
typedef void* cudaStream_t;
inline dim3 threadIdx, blockIdx, blockDim, gridDim;

#define __global__
#define __device__
#define __host__
#define __constant__ static const
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __ldg(p) (*(p))

struct alignas(16) float4 { float x, y, z, w; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
template <class T> inline T atomicAdd(T* p, T v) { T o = *p; *p += v; return o; }
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }

namespace mc {
struct Error : public std::runtime_error {
    explicit Error(const std::string& m) : std::runtime_error(m) {}
};
#define MC_CHECK(cond, msg)                                                                       \
    do {                                                                                          \
        if (!(cond)) throw mc::Error(std::string("check failed: ") + #cond + ": " + (msg));       \
    } while (0)
inline void pdl_sync() {}
template <class T> struct ident_t { using type = T; };
template <typename... KArgs>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t, cudaStream_t, typename ident_t<KArgs>::type... args) {
    gridDim = grid; blockDim = block;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx)
                for (unsigned tz = 0; tz < block.z; ++tz)
                    for (unsigned ty = 0; ty < block.y; ++ty)
                        for (unsigned tx = 0; tx < block.x; ++tx) {
                            blockIdx = dim3(bx, by, bz); threadIdx = dim3(tx, ty, tz);
                            kernel(args...);
                        }
}
inline void zero_async(void* p, size_t bytes, cudaStream_t) { std::memset(p, 0, bytes); }
inline int sm_count() { return 2; }            // small grids: the host run is sequential
constexpr int kNumStems = 9, kStemC = 64, kStemTot = 576, kNumAff = 10, kNumOut = 65, kNumPred = 10, kMaxSrc = 4;
}  // namespace mc
