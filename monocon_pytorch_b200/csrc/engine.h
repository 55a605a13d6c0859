// Engine: static layer plan for DLA-34 + DLAUp + MonoCon heads, parameter store / folding,
// activation arena, launch sequence (optionally replayed as a CUDA graph).
#pragma once
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace mc {

struct TensorInfo {
    std::string name;
    int C = 0, H = 0, W = 0;   // NHWC, leading dim = max_batch
    int Wp = 0;                // physical row pitch in pixels (== W except for the padded tensor-core stem input)
    int xoff = 0;              // physical column of x = 0
    void* ptr = nullptr;
    size_t bytes = 0;
    DType dt = DT_F32;         // storage type (the net's, except the fp32 head-stem tensor of a DT_SPLIT net)
    long long plane = 0;       // DT_SPLIT: elements between the hi and the lo plane (= max_batch * H * Wp * C)
    bool hl_interleaved = false;   // DT_SPLIT network input: one plane, 8-channel pixels [hi0 hi1 hi2 0 lo0 lo1 lo2 0]: the stem then
                                   // needs two MMA passes ([w_lo | 0], then [w_hi | w_hi]) instead of three
};

struct TcConvPlan;             // tensor-core (tcgen05) launch plan, conv_tc.cu
struct Tc2ConvPlan;            // halo-view tensor-core launch plan, conv_tc2.cu
struct Tc3ConvPlan;            // halo-view + streamed-weights launch plan, conv_tc3.cu
struct DcnTcPlan;              // fused deformable convolution, dcn_tc.cu

struct ConvLayer {
    std::string name;
    std::vector<int> src;      // tensor ids, concatenated along C in this order
    int dst = -1;
    int k = 1, stride = 1, pad = 0;
    int cin = 0, cout = 0;     // cin = logical input channels (3 for the stem; storage is padded to 4)
    int cin_store = 0;         // channels of the stored sources summed
    int residual = -1;
    bool relu = false;
    double* stats_sums = nullptr;   // head stems, fp32 output on the streamed-weight kernel: the epilogue also accumulates the AttnBN
                                    // instance statistics [B][cout][2] here (set by the engine once the head buffers exist)
    int dcn_off = -1;          // >= 0: a modulated deformable 3x3 convolution sampled at this offset / mask tensor (csrc/dcn_tc.cu)
    bool dcn_mask_logits = true;
    std::shared_ptr<DcnTcPlan> dcn;
    int pool_dst = -1;         // fp32-accurate mode: the 2x2 max-pool of this layer's output is written by its own epilogue (tensor id)
    // bf16 tensor-core training (train_engine_tc.cu): the kernels write the raw convolution output here instead of the dst tensor
    // (same geometry and type), and the packers record where every packed weight element comes from
    void* dst_override = nullptr;
    bool dst_override_f32 = false;   // ... as fp32 (the head stems: AttnBN statistics and the head backward read fp32)
    bool keep_widx = false;
    std::vector<int> widx;     // per packed element: index into the OIHW weight array handed to pack_conv, -1 = structural zero
    void* w_packed = nullptr;  // the plan's device weight buffer (bf16), widx.size() elements
    // where the parameters come from (state_dict keys); several parts are concatenated along Cout
    struct Part {
        std::string wkey;      // conv weight key (OIHW)
        std::string bn;        // BatchNorm prefix ("" = none)
        std::string bias;      // conv bias key ("" = none)
        float eps = 1e-5f;
        int pad_cout = 0;      // > the weight's Cout: zero filters (scale 1, shift 0) are appended up to this many outputs
                               // (the 27-channel offset / mask convolution of a deformable block runs as a 32-channel layer)
        bool taps_to_k = false;   // the (Cout, Cin, 3, 3) weight feeds a 1x1 layer over deformable columns: K index = tap * Cin + c
    };
    std::vector<Part> parts;
    // packed device parameters
    float* w_simt = nullptr;   // [k*k][cin_store][cout]
    float* scale = nullptr;
    float* shift = nullptr;
    std::shared_ptr<TcConvPlan> tc;
    std::shared_ptr<Tc2ConvPlan> tc2;
    std::shared_ptr<Tc3ConvPlan> tc3;
    bool use_tc = false;       // any tensor-core kernel (v1 tap boxes or v2 halo views)
    bool use_tc2 = false;      // v2 halo-view kernel (conv_tc2.cu)
    bool use_tc3 = false;      // v3 halo-view kernel with streamed weights (conv_tc3.cu)
    double flops_per_image = 0;
    double bytes_per_image = 0;
};

enum OpType { OP_CONV, OP_POOL, OP_UP, OP_HEADS, OP_DCN_COL };
struct Op {
    OpType type;
    int conv = -1;             // OP_CONV: index into convs
    int src = -1, dst = -1;    // OP_POOL / OP_UP; OP_DCN_COL: dst = the column tensor
    std::vector<int> srcs;     // OP_DCN_COL: the sampled tensors (concatenated along C)
    int off = -1;              // OP_DCN_COL: the offset / mask tensor (18 offsets + 9 modulation values per pixel)
    bool mask_logits = true;   // OP_DCN_COL: the modulation values are logits (the kernel applies the sigmoid)
    std::string wkey;          // OP_UP: depthwise deconv weight key
    float* w_dev = nullptr;
};

struct HeadParams {            // device buffers of the AttnBN / 1x1 stage
    float *att_w = nullptr, *att_scale = nullptr, *att_shift = nullptr, *bank_w = nullptr, *bank_b = nullptr;
    float *bn_mean = nullptr, *bn_inv = nullptr, *w = nullptr, *bias = nullptr;
    double* sums = nullptr;
    float *coefA = nullptr, *coefB = nullptr;
};

class DeviceArena {
   public:
    ~DeviceArena();
    void* alloc(size_t bytes);
    size_t total() const { return total_; }
    // Replay: between begin_replay(first, last) and end_replay() every alloc() hands out, in order, the blocks [first, last)
    // of an earlier identical allocation sequence (sizes are checked) instead of new memory -- how mc_refresh_params repacks new
    // weights into the SAME device buffers, so that tensor maps, CUDA graphs and optimiser handles built on them stay valid.
    size_t mark() const { return blocks_.size(); }
    void begin_replay(size_t first, size_t last);
    void end_replay();

   private:
    std::vector<void*> blocks_;
    std::vector<size_t> sizes_;
    size_t total_ = 0;
    bool replay_ = false;
    size_t replay_idx_ = 0, replay_end_ = 0;
};

class Net {
   public:
    Net(int device, int max_batch, DType dt, int conv_impl);
    ~Net();

    int add_tensor(const std::string& name, int C, int H, int W, int Wp = 0, int xoff = 0);
    int add_conv(const std::string& name, const std::vector<int>& src, int cout, int k, int stride, int pad,
                 const std::vector<ConvLayer::Part>& parts, int residual, bool relu, int cin_logical = 0);
    // a convolution into an EXISTING tensor (the dgrad convolutions of the training step write / accumulate gradient tensors)
    int add_conv_to(const std::string& name, const std::vector<int>& src, int dst, int k, int stride, int pad, int residual, bool relu);
    int add_pool(int src);
    int add_up(int src, const std::string& wkey);
    // deformable-convolution columns (csrc/dcn.cu): [9 * sum of the sources' C] channels per pixel, sampled at `off`
    // a fused deformable 3x3 convolution (csrc/dcn_tc.cu): add_conv + the offset / mask tensor it samples at
    int add_dcn_conv(const std::string& name, const std::vector<int>& src, int off, int cout, const std::vector<ConvLayer::Part>& parts, bool relu,
                     bool mask_logits = true);
    int add_dcn_columns(const std::string& name, const std::vector<int>& src, int off, bool mask_logits = true);
    void alias(const std::string& name, int tensor) { aliases_[name] = tensor; }

    void allocate();           // arena for all tensors (+ the scale / running-maximum tables of a DT_SPLIT net)
    void set_tensor_dtype(int tensor, DType t);      // before allocate()
    // fp32-accurate mode (DT_SPLIT): per-tensor power-of-two scale 2^e of the stored fp16 planes, kept in a device table that
    // the kernels read at run time (so a captured CUDA graph follows a recalibration), and the running maximum of |stored value|
    // that every writer of a tensor maintains.  Tensors that meet in one convolution's K dimension (concatenated sources) and
    // a max-pool's input / output share an exponent.
    const ActScale* act_scale(int tensor) const { return d_actscale ? d_actscale + tensor : nullptr; }
    unsigned* act_amax(int tensor) const { return d_amax ? d_amax + tensor : nullptr; }
    SplitInfo split_info(int tensor) const;
    float* upload_split_scale(const ConvLayer& L, const std::vector<int>& ew);   // scale[c] * 2^-ew[c]
    void set_act_exponents(const std::vector<int>& e);                            // host -> device table
    std::vector<float> read_act_amax(bool reset);                                  // device -> host (synchronises)
    std::vector<int> act_exp;          // host mirror, one exponent per tensor
    ActScale* d_actscale = nullptr;
    unsigned* d_amax = nullptr;
    // pack one conv from host arrays (OIHW weight with the logical cin)
    void pack_conv(ConvLayer& L, const std::vector<float>& w_oihw, const std::vector<float>& scale,
                   const std::vector<float>& shift);
    void run_ops(int B, cudaStream_t st, int first = 0, int last = -1);
    void run_conv(int conv, int B, cudaStream_t st);     // one convolution of `convs` through whichever kernel family it was packed for

    int device;
    int max_batch;
    DType dt;
    int conv_impl;             // MC_CONV_*
    bool keep_master = false;  // keep the fp32 [tap][cin][cout] weights next to a tensor-core plan (training: the optimiser's master copy)
    std::vector<TensorInfo> tensors;
    std::vector<ConvLayer> convs;
    std::vector<Op> ops;
    std::map<std::string, int> aliases_;
    std::map<int, int> pooled_;       // de-duplicate Tree max-pools of the same tensor
    DeviceArena arena;
    int launches_last_run = 0;
};

// fp32-accurate mode, weights: per-output-channel exponents ew with max_k |w[o][k]| * 2^ew in [2^13, 2^14) (fp16 keeps 11
// bits down to 2^-14, so both pieces of every weight within 2^-16 of the filter's largest are exact to ~22 bits), and the
// fp16 pieces hi = fp16(w * 2^ew), lo = fp16(w * 2^ew - hi) as raw bits
std::vector<int> split_weight_exponents(const std::vector<float>& w_oihw, int cout);
uint16_t split_weight_piece(float w, int ew, bool lo);
uint16_t bf16_bits(float v);

// tensor-core path (conv_tc.cu)
bool tc_conv_supported(const Net& net, const ConvLayer& L);
void tc_conv_prepare(Net& net, ConvLayer& L, const std::vector<float>& w_oihw);
void tc_conv_launch(const Net& net, const ConvLayer& L, int B, cudaStream_t st);
void tc_kernels_init();
// halo-view tensor-core path (conv_tc2.cu)
bool tc2_conv_supported(const Net& net, const ConvLayer& L);
void tc2_conv_prepare(Net& net, ConvLayer& L, const std::vector<float>& w_oihw);
void tc2_conv_launch(const Net& net, const ConvLayer& L, int B, cudaStream_t st);
void tc2_kernels_init();
// halo-view tensor-core path with streamed weights (conv_tc3.cu)
bool tc3_conv_supported(const Net& net, const ConvLayer& L);
void tc3_conv_prepare(Net& net, ConvLayer& L, const std::vector<float>& w_oihw);
void tc3_conv_launch(const Net& net, const ConvLayer& L, int B, cudaStream_t st);
void tc3_kernels_init();
// fused deformable convolution (dcn_tc.cu)
bool dcn_tc_supported(const Net& net, const ConvLayer& L);
void dcn_tc_prepare(Net& net, ConvLayer& L, const std::vector<float>& w_oihw);
void dcn_tc_launch(const Net& net, const ConvLayer& L, int B, cudaStream_t st);
void dcn_tc_init();
// tensor-core head apply (head_tc.cu)
struct HeadTcPlan;
bool head_tc_supported(DType dt, int HW);
std::shared_ptr<HeadTcPlan> head_tc_prepare(Net& net, const void* stems, DType stems_dt, int max_batch, int HW, const std::vector<float>& w_host,
                                            int z_tensor = -1);
void launch_head_apply_tc(const HeadTcPlan& plan, const HeadApplyParams& ap, cudaStream_t st);
void head_tc_init();

}  // namespace mc
