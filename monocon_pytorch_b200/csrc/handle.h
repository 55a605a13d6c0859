// The engine handle behind the C ABI (include/monocon_b200.h): shared by api.cu (inference, fp32 training) and
// train_engine_tc.cu (bf16 tensor-core training).
#pragma once
#include <cstring>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/monocon_b200.h"
#include "engine.h"

namespace mc {
struct TrainTc;                // bf16 tensor-core training state (train_engine_tc.cu)
struct HostParam {
    std::vector<float> data;
    std::vector<int64_t> shape;
};
}  // namespace mc

using namespace mc;

struct mc_handle {
    int device = 0, max_batch = 0, H = 0, W = 0, prec = 0;
    int neck = 0;                              // MC_NECK_CONV / MC_NECK_DCN (mc_create_ex)
    DType dt = DT_BF16;
    std::unique_ptr<Net> net;
    std::unordered_map<std::string, mc::HostParam> params;
    bool finalized = false;
    size_t fin_first = 0, fin_last = 0;        // arena blocks of the first finalize (mc_refresh_params repacks into them)
    int fin_training = 0;
    std::string err;
    // plan landmarks
    int t_input = -1, t_feat = -1, t_stems = -1, t_headz = -1;
    int fh = 0, fw = 0;
    HeadParams hp;
    bool stats_fused = false;                  // the AttnBN instance statistics come out of the stem convolution's epilogue
    std::shared_ptr<HeadTcPlan> head_tc;       // tensor-core head apply (bf16 mode, conv_impl auto), else the SIMT kernel
    float* pred_own[kNumPred] = {nullptr};
    // train-mode forward (mc_finalize_params(h, 1) / mc_forward_train): per-convolution BatchNorm state
    struct BnTrain { float *gamma = nullptr, *beta = nullptr, *rmean = nullptr, *rvar = nullptr, *scale = nullptr, *shift = nullptr;
                     double* sums = nullptr; float eps = 1e-5f; int C = 0; std::string prefix; };
    bool training = false;
    std::vector<BnTrain> bn_train;             // indexed like net->convs (C == 0: no BatchNorm behind that convolution)
    // backward pass (mc_finalize_params(h, 2) / mc_backward_train; EXPERIMENTAL, see csrc/train_backward.h): what the forward
    // keeps per convolution (raw output, batch mean / inverse std) and the gradient buffers
    struct BwdConv { float *raw = nullptr, *mean = nullptr, *inv = nullptr, *dw = nullptr, *dgamma = nullptr, *dbeta = nullptr, *dbias = nullptr;
                     std::vector<int> part_cout; };
    bool backward = false, grads_valid = false;
    int last_train_B = 0;
    long long train_generation = 0;            // counts mc_forward_train calls: the saved activations belong to the LAST one
    std::vector<BwdConv> bwd_conv;             // indexed like net->convs
    std::vector<float*> bwd_g;                 // per tensor (null: the input image)
    std::vector<float*> bwd_up_dw;             // per op (OP_UP only)
    float* bwd_dw_pool = nullptr;              // all convolution weight gradients, contiguous
    size_t bwd_dw_pool_floats = 0;
    float* bwd_draw = nullptr;
    float* bwd_wT = nullptr;                   // transposed weights of the convolution being differentiated (largest layer)
    double* bwd_sums = nullptr;
    float *bwd_hdw = nullptr, *bwd_hdbias = nullptr, *bwd_datt_w = nullptr, *bwd_datt_gamma = nullptr, *bwd_datt_beta = nullptr,
          *bwd_dbank_w = nullptr, *bwd_dbank_b = nullptr;
    struct TrainTensor { std::string key; float* param = nullptr; float* grad = nullptr; int64_t numel = 0; int stage = -1; };
    std::vector<TrainTensor> train_tensors;    // every trainable buffer of the plan in the ENGINE's layout, with its gradient buffer
    std::vector<mc_bw_tensor> bwd_tensors;
    std::vector<mc_bw_op> bwd_ops;
    bool train_debug = false, head_backward_fast = true;   // mc_set_option
    std::shared_ptr<mc::TrainTc> train_tc;    // set: mc_finalize_params(h, 1 | 2) on an MC_PREC_BF16 handle (tensor-core training step)
    mc_bw_heads_args bwd_hargs;
    float *att_gamma = nullptr, *att_beta = nullptr, *att_rmean = nullptr, *att_rvar = nullptr;   // [9][10]
    float *hbn_rmean = nullptr, *hbn_rvar = nullptr;                                             // [576]
    float* d_lut = nullptr;                    // [3][256] normalisation table of the uint8 input path (mc_set_normalization)
    // decode scratch / staging
    unsigned long long* cand = nullptr;
    int* cand_count = nullptr;
    float *d_img = nullptr, *d_P2 = nullptr, *d_invP = nullptr;
    float *d_box2d = nullptr, *d_box3d = nullptr;
    long long *d_labels = nullptr, *d_inds = nullptr;
    unsigned char* d_valid = nullptr;
    int staging_topk = 0;
    // double-buffered host pipeline (mc_infer_host_submit / mc_infer_host_wait)
    struct HostSlot {
        float *d_img = nullptr, *d_P2 = nullptr, *d_invP = nullptr, *d_b2 = nullptr, *d_b3 = nullptr;
        int* d_hw = nullptr;                       // uint8 path: valid (height, width) per frame
        long long *d_lb = nullptr, *d_ix = nullptr;
        unsigned char* d_vl = nullptr;
        cudaEvent_t ev_in = nullptr, ev_done = nullptr, ev_out = nullptr;
        int topk = 0;
        bool busy = false;
    } slots[2];
    cudaStream_t st_h2d = nullptr, st_comp = nullptr, st_d2h = nullptr;
    // CUDA graph cache for mc_infer_device
    bool use_graph = false;
    struct GraphKey {
        const void *img, *hw, *P2, *invP, *b2, *b3, *lb, *ix, *vl, *gather;
        int B, topk, H0, W0;
        float thres;
        bool operator==(const GraphKey& o) const { return std::memcmp(this, &o, sizeof(GraphKey)) == 0; }
    };
    std::vector<std::pair<GraphKey, cudaGraphExec_t>> graphs;      // small cache, most recent last
    int launches = 0;
    double flops = 0, bytes = 0;
    // peer-memory all-gather of the decode outputs (mc_gather_*)
    struct Gather {
        int world = 0, rank = 0, topk = 0;
        size_t slot_bytes = 0, data_bytes = 0, block_bytes = 0;
        long long off[5] = {0, 0, 0, 0, 0};
        char* block = nullptr;                    // local: [2][world][slot] data, then the flag words
        char* peer[kMaxPeers] = {nullptr};        // peer blocks (IPC-mapped), peer[rank] = block
        bool connected = false;
        unsigned gen[2] = {0u, 0u};               // host mirror: launches issued per buffer
        unsigned** d_peer_ready[2] = {nullptr, nullptr};   // device arrays of peer ready-flag addresses
        int* d_err = nullptr;
    } gather;
};


namespace mc {
// bf16 tensor-core training step (train_engine_tc.cu), driven by api.cu:
std::shared_ptr<TrainTc> traintc_create();
// before Net::pack_conv of forward convolution `conv_index`: where its raw output goes, weight-index capture, host copy of the weights
void traintc_before_pack(mc_handle* h, int conv_index, ConvLayer& L, const std::vector<float>& w_oihw);
// after all convolutions are packed (and, for a backward-enabled engine, setup_backward has allocated the parameter gradients)
void traintc_setup(mc_handle* h);
void traintc_forward(mc_handle* h, const float* img, int B, float* const pred_out[kNumPred], cudaStream_t st);
// stages [op_first, op_last) of the engine's op list, walked downwards; zero: first segment of a pass (parameter gradients are zeroed)
void traintc_backward(mc_handle* h, int B, int op_first, int op_last, bool zero, cudaStream_t st);
// debug: kind 1 = gradient of forward tensor `index`; 2 / 3 = raw output / gradient of the raw output of forward convolution `index`
void traintc_debug(mc_handle* h, int kind, int index, const void** ptr, DType* dt, int* C, int* H, int* W);
}  // namespace mc
