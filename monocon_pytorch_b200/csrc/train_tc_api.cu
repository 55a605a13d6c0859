// C ABI of the bf16 tensor-core training kernels: stand-alone operator entries for kernel-level parity tests
// (include/monocon_b200.h).  The engine-driven training step lives in api.cu.
#include <cstdio>
#include <vector>

#include "../../include/monocon_b200.h"
#include "train_tc.h"

using namespace mc;

extern "C" {

int mc_conv2d_wgrad_tc(int device, const float* x, int B, int Cin, int H, int W, const float* dy, int Cout, int k, int split, float* dw,
                       void* stream, char* err, int err_len) {
    try {
        MC_CUDA(cudaSetDevice(device));
        MC_CHECK(split >= 1 && split <= kMaxSrc && Cin % split == 0, "split");
        cudaStream_t st = (cudaStream_t)stream;
        wgrad_tc_init();
        DeviceArena arena;
        const int Cs = Cin / split;
        WgradDesc d;
        d.nsrc = split; d.H = H; d.W = W; d.Cout = Cout; d.k = k; d.dw = dw;
        const bool stem = k == 7 && Cin == 3;            // dw: [49][8][Cout], the padded storage channels
        if (stem) {
            void* xs = arena.alloc((size_t)B * H * (W + 8) * 8 * 2);
            launch_pack_input(x, xs, DT_BF16, B, 3, H, W, 8, W + 8, 4, st);
            d.src[0] = WgradSrc{xs, 8, W + 8, 4};
        }
        for (int s = 0; s < split && !stem; ++s) {
            void* xs = arena.alloc((size_t)B * H * W * Cs * 2);
            for (int b = 0; b < B; ++b)
                launch_pack_nhwc(x + ((size_t)b * Cin + (size_t)s * Cs) * H * W, (char*)xs + (size_t)b * H * W * Cs * 2, DT_BF16, 1, Cs, H, W, st);
            d.src[s] = WgradSrc{xs, Cs};
        }
        void* dyb = arena.alloc((size_t)B * H * W * Cout * 2);
        launch_pack_nhwc(dy, dyb, DT_BF16, B, Cout, H, W, st);
        d.dy = dyb;
        MC_CHECK(wgrad_tc_supported(d), "geometry outside the tensor-core weight-gradient kernel");
        auto plan = wgrad_tc_prepare(d, B, arena, "mc_conv2d_wgrad_tc");
        MC_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)k * k * (stem ? 8 : Cin) * Cout, st));
        wgrad_tc_launch(*plan, B, st);
        MC_CUDA(cudaStreamSynchronize(st));
        return 0;
    } catch (const std::exception& e) {
        if (err && err_len > 0) std::snprintf(err, err_len, "%s", e.what());
        return 1;
    }
}

}  // extern "C"
