// Shared declarations of the MonoCon B200 engine (kernel parameter blocks + launchers).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdexcept>
#include <string>

namespace mc {

typedef __nv_bfloat16 bf16;

struct Error : public std::runtime_error {
    explicit Error(const std::string& m) : std::runtime_error(m) {}
};

#define MC_CUDA(expr)                                                                             \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess)                                                                    \
            throw mc::Error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " +  \
                            __FILE__ + ":" + std::to_string(__LINE__));                           \
    } while (0)

#define MC_CHECK(cond, msg)                                                                       \
    do {                                                                                          \
        if (!(cond)) throw mc::Error(std::string("check failed: ") + #cond + ": " + (msg));       \
    } while (0)

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch: every forward-path kernel is launched with the programmatic-stream-serialisation
// attribute and calls pdl_sync() after its prologue (barrier init, TMEM allocation, constant weights) and before it
// touches any activation, so that launch latency and prologues overlap the tail of the previous kernel.
// ---------------------------------------------------------------------------------------------
bool pdl_enabled();            // env MC_PDL=0 disables (engine.cu)
// SMs the persistent tensor-core kernels leave free (env MC_RESERVE_SMS, default 0).  With one CTA per SM and a static
// work split, a communication kernel (the NCCL all-gather of the multi-GPU path) that occupies an SM would make one
// CTA of every overlapping convolution wait for a free SM and double that kernel's duration.
int reserved_sms();
template <class T> struct ident_t { using type = T; };
template <typename... KArgs>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                     typename ident_t<KArgs>::type... args) {
    cudaLaunchConfig_t cfg;
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    void* argv[] = {(void*)&args...};
    MC_CUDA(cudaLaunchKernelExC(&cfg, (const void*)kernel, argv));
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_sync() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
#endif

// DT_SPLIT: the fp32-accurate tensor-core mode.  A tensor is stored as TWO fp16 images per logical image,
//   x * 2^e  =  hi + lo,   hi = fp16(x * 2^e),   lo = fp16(x * 2^e - hi)      (about 22 significant bits),
// laid out plane-major: [2 planes][max_batch][H][Wp][C] (all hi images, then all lo images), so every kernel that walks
// NHWC images through a TMA map simply sees 2 * max_batch images and selects the plane with the image coordinate.
// e is a per-tensor power-of-two exponent kept in a device table (Net::d_actmul), see engine.h.
enum DType { DT_F32 = 0, DT_BF16 = 1, DT_SPLIT = 2 };
inline size_t dtype_size(DType t) { return t == DT_BF16 ? 2 : 4; }   // bytes per logical element (DT_SPLIT: two 2-byte planes)

// Scaling of a DT_SPLIT tensor, read by kernels from the device table: mul = 2^e, inv = 2^-e.  amax: running maximum of
// |stored value| (bit pattern of a non-negative float, atomicMax), what the host calibration reads.
struct ActScale { float mul, inv; };

// ---------------------------------------------------------------------------------------------
// kernel parameter blocks (plain structs, passed by value)
// ---------------------------------------------------------------------------------------------
constexpr int kMaxSrc = 4;

// Direct / implicit-GEMM convolution on NHWC activations with up to four channel-concatenated
// sources (the reference's torch.cat([...], 1) before Root / node convs, dla.py:126, dla_neck.py:104,
// is never materialised).
struct ConvParams {
    const void* src[kMaxSrc];
    int srcC[kMaxSrc];
    int srcWp[kMaxSrc];     // physical row pitch in pixels (>= Win)
    int srcXoff[kMaxSrc];   // physical column of x = 0
    int nsrc;
    int B, Hin, Win, Hout, Wout, Cin, Cout;
    int k, stride, pad;
    const float* w;         // [k*k][Cin][Cout] fp32
    const float* scale;     // [Cout] folded BN scale (1 for bias-only convs)
    const float* shift;     // [Cout] folded BN shift / bias
    const void* residual;   // NHWC [B,Hout,Wout,Cout] or nullptr
    void* dst;              // NHWC [B,Hout,Wout,Cout]
    int relu;
};

void launch_conv_simt(const ConvParams& p, DType dt, cudaStream_t st);

// What the bandwidth kernels need to know about a DT_SPLIT tensor (ignored for the other storage types): the distance
// between its hi and lo planes in elements, its entry in the scale table and its slot in the running-maximum table.
struct SplitInfo {
    long long plane = 0;              // elements between the hi and the lo plane (= max_batch * H * Wp * C)
    bool interleaved = false;         // network input only: ONE plane of 8-channel pixels [hi0 hi1 hi2 0 lo0 lo1 lo2 0] (plane unused)
    const ActScale* sc = nullptr;     // device: this tensor's {2^e, 2^-e}
    unsigned* amax = nullptr;         // device: running max |stored| (kernels that WRITE the tensor update it), may be null
};

// NCHW fp32 image -> NHWC (C padded to Cpad with zeros); rows have a physical pitch of Wp >= W + xoff pixels,
// pixel x is stored at column x + xoff and every other column is zero (used by the tensor-core stem).
void launch_pack_input(const float* img_nchw, void* dst, DType dt, int B, int C, int H, int W, int Cpad, int Wp, int xoff,
                       cudaStream_t st, const SplitInfo& so = SplitInfo());
// uint8 HWC frames (B, H0, W0, 3) with per-image valid sizes hw[B][2] -> normalised, zero-padded NHWC (see kernels_simt.cu)
void launch_pack_input_u8(const unsigned char* src, const int* hw, const float* lut, void* dst, DType dt, int B, int H0, int W0,
                          int H, int W, int Cpad, int Wp, int xoff, cudaStream_t st, const SplitInfo& so = SplitInfo());
// NHWC (T) -> NCHW fp32 (debug / operator tests)
void launch_unpack_nchw(const void* src, DType dt, float* dst_nchw, int B, int C, int H, int W, cudaStream_t st,
                        const SplitInfo& si = SplitInfo());
// NCHW fp32 -> NHWC (T)
void launch_pack_nhwc(const float* src_nchw, void* dst, DType dt, int B, int C, int H, int W, cudaStream_t st,
                      const SplitInfo& so = SplitInfo());

// 2x2 stride-2 max-pool (Tree.downsample, dla.py:179,193) on NHWC.  DT_SPLIT: source and destination share one scale.
void launch_maxpool2(const void* src, void* dst, DType dt, int B, int C, int Hin, int Win, cudaStream_t st,
                     const SplitInfo& si = SplitInfo(), const SplitInfo& so = SplitInfo());

// Depthwise ConvTranspose2d k=4 s=2 p=1, no bias (IDAUp.up_i, dla_neck.py:58-65) on NHWC.
// w: [C][4][4] fp32 (the reference's (C,1,4,4) weight).
void launch_upsample2(const void* src, void* dst, DType dt, const float* w, int B, int C, int Hin, int Win,
                      cudaStream_t st, const SplitInfo& si = SplitInfo(), const SplitInfo& so = SplitInfo());

// Deformable-convolution columns (csrc/dcn.cu): columns[pixel][tap * Cin + c] = mask_tap * bilinear(x_c, pixel + tap + offset_tap)
// for the 3x3 / stride 1 / pad 1 modulated deformable convolution (torchvision.ops.deform_conv2d semantics).  The sources are
// NHWC tensors concatenated along C (both multiples of 8); `off` is an NHWC tensor with >= 27 channels per pixel: 18 offsets
// ((dy, dx) per tap, tap = 3 i + j) followed by the 9 modulation values (logits when mask_logits is set: the kernel applies the
// sigmoid, as the DCNv2 block does to its conv_offset output).  DT_SPLIT tensors carry their plane distance and scale entry.
struct DcnColParams {
    const void* src[2];
    int srcC[2];
    long long src_plane[2];
    const ActScale* src_sc[2];
    int nsrc;
    const void* off;
    int offC;
    long long off_plane;
    const ActScale* off_sc;
    void* col;
    long long col_plane;
    const ActScale* col_sc;
    unsigned* col_amax;
    int B, H, W, Cin;
    int mask_logits;
    unsigned g_magic;        // set by the launcher: (i * g_magic) >> 20 == i / (Cin / 8) for every work-item index of a pixel
};
void launch_dcn_columns(const DcnColParams& p, DType dt, cudaStream_t st);

// ---- heads ------------------------------------------------------------------------------------
constexpr int kNumStems = 9;
constexpr int kStemC = 64;
constexpr int kStemTot = kNumStems * kStemC;   // 576
constexpr int kNumAff = 10;                    // AttnBatchNorm2d num_affine_trans (monocon_heads.py:117)
constexpr int kNumOut = 65;                    // 3+9+2+2+2+18+3+2+12+12
constexpr int kNumPred = 10;

// Per-(b, channel) instance statistics of the nine pre-norm stem outputs (AttnWeights.forward,
// attentive_norm.py:84): partial sums accumulated in fp64 by atomics.
void launch_attn_stats(const void* stems, DType dt, double* sums /*[B][576][2]*/, int B, int HW, cudaStream_t st);

struct AttnMixParams {
    const double* sums;      // [B][576][2]
    int HW;
    const float* att_w;      // [9][10][64]  attention conv1x1 (attentive_norm.py:51)
    const float* att_scale;  // [9][10]      folded BatchNorm2d(10) (attentive_norm.py:52)
    const float* att_shift;  // [9][10]
    const float* bank_w;     // [9][10][64]  weight_ (attentive_norm.py:138)
    const float* bank_b;     // [9][10][64]  bias_
    const float* bn_mean;    // [576] running_mean of the affine-free base BN
    const float* bn_inv;     // [576] rsqrt(running_var + 1e-3)
    float* coefA;            // [B][576]  out = coefA * x + coefB  ==  gamma * BN(x) + beta
    float* coefB;            // [B][576]
};
void launch_attn_mix(const AttnMixParams& p, int B, cudaStream_t st);

struct HeadApplyParams {
    const void* stems;       // [B][HW][576]
    const float* coefA;      // [B][576]
    const float* coefB;
    const float* w;          // [65][64] the ten 1x1 convs, rows in pred order
    const float* bias;       // [65]
    float* out[kNumPred];    // NCHW fp32
    int B, HW;
};
void launch_head_apply(const HeadApplyParams& p, DType dt, cudaStream_t st);
// Per-device one-time setup (constant tables, dynamic shared-memory opt-in); call before any capture.
void head_kernels_init();

// ---- train-mode forward (train_forward.cu) ------------------------------------------------------
// BatchNorm with batch statistics on an NHWC fp32 tensor of P pixels, in place: statistics -> scale / shift + running-stat
// update -> y = x * scale + shift (+ residual) (ReLU).  gamma / beta may be null (affine-free).  sums: 2 * C doubles of scratch.
void launch_bn_train(float* x, const float* residual, long long P, int C, double* sums, float eps, float momentum, const float* gamma,
                     const float* beta, float* rmean, float* rvar, float* scale, float* shift, bool relu, cudaStream_t st);
// The same with separate input / output (the raw convolution output survives for the backward pass) and the batch mean /
// rsqrt(var + eps) written out (both null: not kept).
void launch_bn_train_ex(const float* raw, float* y, const float* residual, long long P, int C, double* sums, float eps, float momentum,
                        const float* gamma, const float* beta, float* rmean, float* rvar, float* scale, float* shift, bool relu, float* mean_out,
                        float* inv_out, cudaStream_t st);
// AttnBatchNorm2d in train mode: from the per-sample sums of attn_stats to the per-sample affine coefA / coefB
void launch_attn_mix_train(const double* sums, int B, int HW, const float* att_w, const float* att_gamma, const float* att_beta,
                           float* att_rmean, float* att_rvar, const float* bank_w, const float* bank_b, float* bn_rmean, float* bn_rvar,
                           float* coefA, float* coefB, cudaStream_t st);

// ---- decode -----------------------------------------------------------------------------------
constexpr int kMaxPeers = 8;
// Peer-memory all-gather fused into the decode kernel's tail (multi-GPU inference, one process per GPU): every
// detection row is also stored, over NVLink, into this rank's slot of each peer's gather buffer; the last CTA of the
// launch publishes the new generation number to every peer.  n == 0: off.
struct GatherParams {
    int n, rank;                          // world size, this rank
    char* peer_slot[kMaxPeers];           // peer r's buffer: start of THIS rank's slot (peer_slot[rank] = the local slot)
    unsigned* peer_data_flag[kMaxPeers];  // peer r's data_flag[buf][rank]
    const unsigned* ready;                // local ready_flag[buf][0..n): generation peer r allows us to overwrite
    unsigned* done;                       // local CTA counter of this launch
    unsigned* gen;                        // local generation counter of this buffer (device-side, so graphs replay)
    long long off_box2d, off_box3d, off_labels, off_inds, off_valid;   // byte offsets inside a slot
    int* error_flag;
};

struct DecodeParams {
    const float* pred[kNumPred];
    int B, C, H, W;          // heat-map geometry (C classes)
    const float* P2;         // [B][3][4]
    const float* invP;       // [B][4][4]
    float scale_x, scale_y;  // img_w / feat_w, img_h / feat_h
    int topk;
    float thres;
    int num_bins;            // 12
    int c2k_channels;        // 18
    float* box2d;            // [B][K][5]
    float* box3d;            // [B][K][7]
    long long* labels;       // [B][K]
    long long* inds;         // [B][K]
    unsigned char* valid;    // [B][K]
    GatherParams gather;
};
// cand: scratch of B * C*H*W 64-bit entries (NMS survivors as (score, index) composites); count: B counters.
// Two launches: batch-wide NMS + candidate append, then one CTA per image for select / gather / lift.
void launch_decode(const DecodeParams& p, unsigned long long* cand, int* count, cudaStream_t st);
// post-decode KITTI conversion of the 3D boxes (kernels_decode.cu)
void launch_kitti_boxes(const float* box3d, const unsigned char* valid, const float* P2, const int* img_hw, int B, int K, double* bbox,
                        float* alpha, unsigned char* keep, cudaStream_t st);
void launch_gather_release(const GatherParams& G, unsigned* const* peer_ready_flag_dev, unsigned gen, cudaStream_t st);
void launch_gather_wait(const unsigned* data_flag, int n, unsigned gen, int* error_flag, cudaStream_t st);

}  // namespace mc
