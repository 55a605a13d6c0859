// C ABI (include/monocon_b200.h): builds the DLA-34 + DLAUp + MonoCon-heads plan, owns the
// parameter store, and runs forward / decode.  All file:line citations refer to the reference repo.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>

#include "../../include/monocon_b200.h"
#include "engine.h"
#include "handle.h"
#include "train_backward.h"

using namespace mc;

namespace {

const char* kStemNames[kNumStems] = {"heatmap_head", "wh_head", "offset_head", "center2kpt_offset_head",
                                     "kpt_heatmap_head", "kpt_heatmap_offset_head", "dim_head", "depth_head",
                                     "dir_feat"};   // registration order, monocon_heads.py:74-88
// pred order (monocon_heads.py:190-200) -> key of its 1x1 conv, channels
const char* kPredConv[kNumPred] = {"head.heatmap_head.3", "head.kpt_heatmap_head.3", "head.wh_head.3",
                                   "head.offset_head.3", "head.kpt_heatmap_offset_head.3",
                                   "head.center2kpt_offset_head.3", "head.dim_head.3", "head.depth_head.3",
                                   "head.dir_cls.0", "head.dir_reg.0"};
const int kPredCh[kNumPred] = {3, 9, 2, 2, 2, 18, 3, 2, 12, 12};

std::string g_create_error;
std::mutex g_mutex;

}  // namespace

namespace {
// flag words behind the data region: data_flag[2][8], ready_flag[2][8], done[2], gen[2]
inline unsigned* g_data_flag(char* block, size_t data_bytes, int buf) { return reinterpret_cast<unsigned*>(block + data_bytes) + buf * kMaxPeers; }
inline unsigned* g_ready_flag(char* block, size_t data_bytes, int buf) { return reinterpret_cast<unsigned*>(block + data_bytes) + 2 * kMaxPeers + buf * kMaxPeers; }
inline unsigned* g_done(char* block, size_t data_bytes, int buf) { return reinterpret_cast<unsigned*>(block + data_bytes) + 4 * kMaxPeers + buf; }
inline unsigned* g_gen(char* block, size_t data_bytes, int buf) { return reinterpret_cast<unsigned*>(block + data_bytes) + 4 * kMaxPeers + 2 + buf; }
}  // namespace

namespace {

// ---------------------------------------------------------------------------------------------
// plan builder: DLA-34 (model/backbone/dla.py), DLAUp (model/backbone/dla_neck.py), heads
// ---------------------------------------------------------------------------------------------
ConvLayer::Part bn_part(const std::string& wkey, const std::string& bn, float eps = 1e-5f) {
    ConvLayer::Part p;
    p.wkey = wkey; p.bn = bn; p.eps = eps;
    return p;
}

// BasicBlock (dla.py:34-51): conv3x3(s) BN ReLU conv3x3 BN (+residual) ReLU
int build_block(Net& n, const std::string& pre, int x, int cout, int stride, int residual) {
    int a = n.add_conv(pre + ".conv1", {x}, cout, 3, stride, 1, {bn_part(pre + ".conv1.weight", pre + ".bn1")}, -1, true);
    return n.add_conv(pre + ".conv2", {a}, cout, 3, 1, 1, {bn_part(pre + ".conv2.weight", pre + ".bn2")}, residual, true);
}

// Tree.forward (dla.py:187-205).  `children` carries the tensors appended for the Root concat.
int build_tree(Net& n, const std::string& pre, int levels, int cin, int cout, int stride, bool level_root, int x,
               std::vector<int> children) {
    const int bottom = stride > 1 ? n.add_pool(x) : x;                       // dla.py:193
    if (level_root) children.push_back(bottom);                              // dla.py:196-197
    if (levels == 1) {
        int residual = bottom;
        if (cin != cout)                                                     // project: conv1x1 + BN, dla.py:181-185,194
            residual = n.add_conv(pre + ".project", {bottom}, cout, 1, 1, 0,
                                  {bn_part(pre + ".project.0.weight", pre + ".project.1")}, -1, false);
        int x1 = build_block(n, pre + ".tree1", x, cout, stride, residual);  // dla.py:198
        int x2 = build_block(n, pre + ".tree2", x1, cout, 1, x1);            // dla.py:200
        std::vector<int> src = {x2, x1};                                     // Root(x2, x1, *children), dla.py:201
        for (int c : children) src.push_back(c);
        return n.add_conv(pre + ".root", src, cout, 1, 1, 0, {bn_part(pre + ".root.conv.weight", pre + ".root.bn")}, -1,
                          true);
    }
    // levels > 1: the outer project output is never consumed (tree1 is a Tree and recomputes its own
    // residual, dla.py:194 vs :198) -> its 1x1 conv is skipped; its parameters are accepted and ignored.
    int x1 = build_tree(n, pre + ".tree1", levels - 1, cin, cout, stride, false, x, {});
    children.push_back(x1);                                                  // dla.py:203
    return build_tree(n, pre + ".tree2", levels - 1, cout, cout, 1, false, x1, children);
}

// Conv2dBlock (dla_neck.py:11-38) with its 3x3 convolution replaced by a modulated deformable convolution (DCNv2 pack:
// `conv.conv_offset` = a plain 3x3 convolution with bias producing 18 offsets + 9 mask logits from the block's own input;
// `conv.weight` (Cout, Cin, 3, 3), no bias; then bn1 + ReLU as in the plain block).  Three stages: the offset convolution
// (run as a 32-channel layer, five zero filters), the deformable columns (csrc/dcn.cu), and a 1x1 convolution over the
// 9 * Cin column channels with the folded BatchNorm + ReLU epilogue.
int build_deform_block(Net& n, const std::string& pre, const std::vector<int>& src, int cout) {
    ConvLayer::Part po;
    po.wkey = pre + ".conv.conv_offset.weight";
    po.bias = pre + ".conv.conv_offset.bias";
    // 27 -> 32 output channels (zero filters); 64 in the fp32-accurate mode, where the streamed-weight halo kernel (conv_tc3.cu, Cout
    // >= 64: one halo tile serves all nine taps) beats the tap-box kernel's N = 32 MMAs.  MC_DCN_OFFC overrides.
    int offc = n.dt == DT_SPLIT ? 64 : 32;
    if (const char* e = std::getenv("MC_DCN_OFFC")) { const int v = std::atoi(e); if (v == 32 || v == 64) offc = v; }
    po.pad_cout = offc;
    const int off = n.add_conv(pre + ".conv_offset", src, offc, 3, 1, 1, {po}, -1, false);
    n.convs.back().flops_per_image *= 27.0 / offc;            // algorithmic work: the 27 real filters
    ConvLayer::Part pw = bn_part(pre + ".conv.weight", pre + ".bn1");
    // tensor-core storage: sampling and contraction in ONE kernel (csrc/dcn_tc.cu), no column tensor; MC_DCN_FUSE=0 and the
    // fp32 FFMA twin keep the two stages
    const char* e = std::getenv("MC_DCN_FUSE");
    bool c64 = true;
    for (int s : src) c64 = c64 && n.tensors[s].C % 64 == 0;
    if ((n.dt == DT_BF16 || n.dt == DT_SPLIT) && c64 && cout <= 256 && !(e && e[0] == '0')) return n.add_dcn_conv(pre, src, off, cout, {pw}, true);
    const int col = n.add_dcn_columns(pre + ".columns", src, off, true);
    pw.taps_to_k = true;
    return n.add_conv(pre, {col}, cout, 1, 1, 0, {pw}, -1, true);
}

void build_plan(mc_handle* h) {
    Net& n = *h->net;
    const int H = h->H, W = h->W;
    const int ch[6] = {16, 32, 64, 128, 256, 512};
    const int lv[6] = {1, 1, 1, 2, 2, 1};                                    // dla.py:211
    // NHWC input: fp32 mode C 3 -> 4; bf16 mode C 3 -> 8 with 4 zero columns left and right of every row,
    // the layout the tensor-core stem's overlapping-window TMA view needs (conv_tc.cu)
    if (h->dt == DT_BF16 || h->dt == DT_SPLIT) {
        h->t_input = n.add_tensor("input", 8, H, W, W + 8, 4);
        if (h->dt == DT_SPLIT) n.tensors[h->t_input].hl_interleaved = true;     // hi and lo of the 3 colour channels share one 16-byte pixel
    } else {
        h->t_input = n.add_tensor("input", 4, H, W);
    }
    int x = n.add_conv("backbone.base_layer", {h->t_input}, 16, 7, 1, 3,
                       {bn_part("backbone.base_layer.0.weight", "backbone.base_layer.1")}, -1, true, 3);   // dla.py:231-234
    x = n.add_conv("backbone.level0", {x}, 16, 3, 1, 1, {bn_part("backbone.level0.0.weight", "backbone.level0.1")}, -1, true);
    int l1 = n.add_conv("backbone.level1", {x}, 32, 3, 2, 1, {bn_part("backbone.level1.0.weight", "backbone.level1.1")}, -1, true);
    int lvl[6] = {x, l1, -1, -1, -1, -1};
    x = l1;
    for (int l = 2; l < 6; ++l) {                                            // dla.py:238-241
        x = build_tree(n, "backbone.level" + std::to_string(l), lv[l], ch[l - 1], ch[l], 2, l != 2, x, {});
        lvl[l] = x;
        n.alias("backbone.level" + std::to_string(l), x);
    }
    // DLAUp.forward (dla_neck.py:136-143) over [l2, l3, l4, l5]; IDAUp.forward (:94-106)
    std::vector<int> layers = {lvl[2], lvl[3], lvl[4], lvl[5]};
    for (int i = 0; i < 3; ++i) {
        const int start = (int)layers.size() - i - 2;
        const int cout = n.tensors[layers[start]].C;
        const std::string pre = "neck.ida_" + std::to_string(i);
        for (int j = 1; start + j < (int)layers.size(); ++j) {
            const std::string sj = std::to_string(j);
            if (h->neck == MC_NECK_DCN) {                                     // the DCNv2 variant of the two Conv2dBlocks
                int p = build_deform_block(n, pre + ".proj_" + sj, {layers[start + j]}, cout);
                int u = n.add_up(p, pre + ".up_" + sj + ".weight");
                layers[start + j] = build_deform_block(n, pre + ".node_" + sj, {layers[start + j - 1], u}, cout);
                continue;
            }
            int p = n.add_conv(pre + ".proj_" + sj, {layers[start + j]}, cout, 3, 1, 1,
                               {bn_part(pre + ".proj_" + sj + ".conv.weight", pre + ".proj_" + sj + ".bn1")}, -1, true);
            int u = n.add_up(p, pre + ".up_" + sj + ".weight");
            layers[start + j] = n.add_conv(pre + ".node_" + sj, {layers[start + j - 1], u}, cout, 3, 1, 1,
                                           {bn_part(pre + ".node_" + sj + ".conv.weight", pre + ".node_" + sj + ".bn1")}, -1, true);
        }
    }
    h->t_feat = layers.back();
    n.alias("neck.feat", h->t_feat);
    // nine 3x3 stems (conv + bias) as one convolution with Cout = 576 (monocon_heads.py:114-131)
    std::vector<ConvLayer::Part> parts;
    for (int s = 0; s < kNumStems; ++s) {
        ConvLayer::Part p;
        p.wkey = std::string("head.") + kStemNames[s] + ".0.weight";
        p.bias = std::string("head.") + kStemNames[s] + ".0.bias";
        parts.push_back(p);
    }
    h->t_stems = n.add_conv("head.stems", {h->t_feat}, kStemTot, 3, 1, 1, parts, -1, false);
    // fp32-accurate tensor-core mode: the pre-norm stems go straight to fp32 (AttnBN statistics and the 1x1 heads read fp32)
    if (h->dt == DT_SPLIT) {
        n.set_tensor_dtype(h->t_stems, DT_F32);
        // the post-AttnBN activations z never reach memory (the head kernel splits them in shared memory), but they need a
        // calibrated fp16 scale and a running maximum like every stored tensor: a one-element pseudo-tensor carries both
        h->t_headz = n.add_tensor("head.z", 1, 1, 1);
    }
    Op op;
    op.type = OP_HEADS;
    n.ops.push_back(op);
    h->fh = n.tensors[h->t_feat].H;
    h->fw = n.tensors[h->t_feat].W;
}

const HostParam& get_param(mc_handle* h, const std::string& key) {
    auto it = h->params.find(key);
    if (it == h->params.end()) throw Error("missing parameter: " + key);
    return it->second;
}

float* upload(mc_handle* h, const std::vector<float>& v) {
    float* d = (float*)h->net->arena.alloc(sizeof(float) * v.size());
    MC_CUDA(cudaMemcpy(d, v.data(), sizeof(float) * v.size(), cudaMemcpyHostToDevice));
    return d;
}

// eval-mode BatchNorm fold: y = (x - mean) * rsqrt(var + eps) * gamma + beta
void fold_bn(mc_handle* h, const std::string& bn, float eps, bool affine, std::vector<float>& scale, std::vector<float>& shift) {
    const HostParam& rm = get_param(h, bn + ".running_mean");
    const HostParam& rv = get_param(h, bn + ".running_var");
    const size_t c = rm.data.size();
    for (size_t i = 0; i < c; ++i) {
        const float inv = 1.0f / std::sqrt(rv.data[i] + eps);
        const float g = affine ? get_param(h, bn + ".weight").data[i] : 1.f;
        const float b = affine ? get_param(h, bn + ".bias").data[i] : 0.f;
        scale.push_back(g * inv);
        shift.push_back(b - rm.data[i] * g * inv);
    }
}

// Buffers and stage records of the backward pass (mc_bw_run_graph, csrc/train_backward.cu): one gradient tensor per
// activation, per convolution the raw output + batch statistics the forward keeps and the parameter gradients, and the
// engine's op list restated as mc_bw_op records (tests/test_backward_graph_host.py builds the same records on the CPU).
void setup_backward(mc_handle* h) {
    Net& n = *h->net;
    auto& a = n.arena;
    const size_t MB = (size_t)h->max_batch;
    const bool tc = h->train_tc != nullptr;      // bf16 tensor-core step: activation gradients are bf16 tensors of its own net (train_engine_tc.cu);
                                                 // only the gradient of the fp32 head stems, the parameter gradients and the stage list live here
    h->bwd_g.assign(n.tensors.size(), nullptr);
    h->bwd_tensors.resize(n.tensors.size());
    for (size_t i = 0; i < n.tensors.size(); ++i) {
        const TensorInfo& t = n.tensors[i];
        if ((int)i != h->t_input && (!tc || (int)i == h->t_stems)) h->bwd_g[i] = (float*)a.alloc(sizeof(float) * MB * t.H * t.W * t.C);
        mc_bw_tensor& b = h->bwd_tensors[i];
        b.x = (const float*)t.ptr; b.g = h->bwd_g[i]; b.C = t.C; b.H = t.H; b.W = t.W; b.Wp = t.Wp > 0 ? t.Wp : t.W; b.xoff = t.xoff;
    }
    size_t max_out = 0, max_w = 0;
    // all convolution weight gradients in ONE allocation (64-float aligned pieces): a pass zeroes them with one memset
    std::vector<size_t> dw_off(n.convs.size());
    size_t dw_total = 0;
    for (size_t i = 0; i < n.convs.size(); ++i) {
        const ConvLayer& L = n.convs[i];
        dw_off[i] = dw_total;
        dw_total += ((size_t)L.k * L.k * L.cin_store * L.cout + 63) / 64 * 64;
    }
    float* dw_pool = (float*)a.alloc(sizeof(float) * dw_total);
    h->bwd_dw_pool = dw_pool; h->bwd_dw_pool_floats = dw_total;
    for (size_t i = 0; i < n.convs.size(); ++i) {
        const ConvLayer& L = n.convs[i];
        const TensorInfo& d = n.tensors[L.dst];
        auto& bc = h->bwd_conv[i];
        const size_t elems = MB * d.H * d.W * L.cout;
        max_w = std::max(max_w, (size_t)L.k * L.k * L.cin_store * L.cout);
        MC_CHECK(L.w_simt, "backward: the convolution has no fp32 weights");
        bc.dw = dw_pool + dw_off[i];
        if (h->bn_train[i].C > 0) {
            if (!tc) bc.raw = (float*)a.alloc(sizeof(float) * elems);
            bc.mean = (float*)a.alloc(sizeof(float) * L.cout);
            bc.inv = (float*)a.alloc(sizeof(float) * L.cout);
            bc.dgamma = (float*)a.alloc(sizeof(float) * L.cout);
            bc.dbeta = (float*)a.alloc(sizeof(float) * L.cout);
            if (elems > max_out) max_out = elems;
        } else {
            bc.dbias = (float*)a.alloc(sizeof(float) * L.cout);
        }
    }
    if (!tc) {
        h->bwd_draw = (float*)a.alloc(sizeof(float) * max_out);
        h->bwd_wT = (float*)a.alloc(sizeof(float) * max_w);
    }
    h->bwd_sums = (double*)a.alloc(sizeof(double) * 2 * 1024);
    const int HW = h->fh * h->fw;
    h->bwd_hdw = (float*)a.alloc(sizeof(float) * kNumOut * kStemC);
    h->bwd_hdbias = (float*)a.alloc(sizeof(float) * kNumOut);
    h->bwd_datt_w = (float*)a.alloc(sizeof(float) * kNumStems * kNumAff * kStemC);
    h->bwd_datt_gamma = (float*)a.alloc(sizeof(float) * kNumStems * kNumAff);
    h->bwd_datt_beta = (float*)a.alloc(sizeof(float) * kNumStems * kNumAff);
    h->bwd_dbank_w = (float*)a.alloc(sizeof(float) * kNumStems * kNumAff * kStemC);
    h->bwd_dbank_b = (float*)a.alloc(sizeof(float) * kNumStems * kNumAff * kStemC);
    mc_bw_heads_args& ha = h->bwd_hargs;
    std::memset(&ha, 0, sizeof(ha));
    ha.scratch = a.alloc(head_bwd_scratch_bytes((int)MB, HW));
    ha.sums = h->hp.sums; ha.coefA = h->hp.coefA; ha.coefB = h->hp.coefB; ha.att_w = h->hp.att_w; ha.att_gamma = h->att_gamma;
    ha.att_beta = h->att_beta; ha.bank_w = h->hp.bank_w; ha.bank_b = h->hp.bank_b; ha.w = h->hp.w;
    ha.dw = h->bwd_hdw; ha.dbias = h->bwd_hdbias; ha.datt_w = h->bwd_datt_w; ha.datt_gamma = h->bwd_datt_gamma;
    ha.datt_beta = h->bwd_datt_beta; ha.dbank_w = h->bwd_dbank_w; ha.dbank_b = h->bwd_dbank_b;
    h->bwd_up_dw.assign(n.ops.size(), nullptr);
    h->bwd_ops.clear();
    for (size_t i = 0; i < n.ops.size(); ++i) {
        const Op& op = n.ops[i];
        mc_bw_op o;
        std::memset(&o, 0, sizeof(o));
        o.dst = -1; o.residual = -1;
        if (op.type == OP_CONV) {
            const ConvLayer& L = n.convs[op.conv];
            const auto& bc = h->bwd_conv[op.conv];
            const auto& bt = h->bn_train[op.conv];
            o.type = MC_BW_CONV;
            o.nsrc = (int)L.src.size();
            MC_CHECK(o.nsrc <= 4, "backward: more than four concatenated sources");
            for (int s = 0; s < o.nsrc; ++s) o.src[s] = L.src[s];
            o.dst = L.dst; o.residual = L.residual; o.relu = L.relu ? 1 : 0;
            o.k = L.k; o.stride = L.stride; o.pad = L.pad; o.cout = L.cout;
            o.w = L.w_simt; o.dw = bc.dw; o.dbias = bc.dbias; o.wT = h->bwd_wT;
            o.has_bn = bt.C > 0 ? 1 : 0;
            o.raw = bc.raw; o.mean = bc.mean; o.inv = bc.inv; o.gamma = bt.gamma; o.dgamma = bc.dgamma; o.dbeta = bc.dbeta;
            o.draw = h->bwd_draw; o.sums = h->bwd_sums;
        } else if (op.type == OP_POOL) {
            o.type = MC_BW_POOL; o.nsrc = 1; o.src[0] = op.src; o.dst = op.dst;
        } else if (op.type == OP_UP) {
            h->bwd_up_dw[i] = (float*)a.alloc(sizeof(float) * n.tensors[op.src].C * 16);
            o.type = MC_BW_UP; o.nsrc = 1; o.src[0] = op.src; o.dst = op.dst; o.w = op.w_dev; o.dw = h->bwd_up_dw[i];
        } else {
            o.type = MC_BW_HEADS; o.nsrc = 1; o.src[0] = h->t_stems; o.heads = &h->bwd_hargs;
        }
        h->bwd_ops.push_back(o);
    }
    // The trainable buffers as the engine holds them (packed convolution weights, BatchNorm weight / bias, biases, upsampling
    // taps, head matrices) next to their gradient buffers: AdamW and the gradient norm are element-wise, so the optimiser can
    // step these in place without ever unpacking to the state_dict layout (mc_optimizer_* over mc_train_tensor pointers).
    int stage = -1;                              // index in n.ops of the op whose backward finishes the gradient
    auto add = [&](const std::string& key, float* param, float* grad, size_t numel) {
        mc_handle::TrainTensor t;
        t.key = key; t.param = param; t.grad = grad; t.numel = (int64_t)numel; t.stage = stage;
        MC_CHECK(param && grad && numel > 0, "backward: incomplete trainable tensor " + key);
        h->train_tensors.push_back(t);
    };
    h->train_tensors.clear();
    std::vector<int> conv_stage(n.convs.size(), -1);
    int heads_stage = -1;
    for (size_t i = 0; i < n.ops.size(); ++i) {
        if (n.ops[i].type == OP_CONV) conv_stage[n.ops[i].conv] = (int)i;
        if (n.ops[i].type == OP_HEADS) heads_stage = (int)i;
    }
    for (size_t i = 0; i < n.convs.size(); ++i) {
        ConvLayer& L = n.convs[i];
        const auto& bc = h->bwd_conv[i];
        const auto& bt = h->bn_train[i];
        stage = conv_stage[i];
        add(L.name + ".weight[packed]", L.w_simt, bc.dw, (size_t)L.k * L.k * L.cin_store * L.cout);
        if (bt.C > 0) {
            add(bt.prefix + ".weight", bt.gamma, bc.dgamma, L.cout);
            add(bt.prefix + ".bias", bt.beta, bc.dbeta, L.cout);
        } else {
            add(L.name + ".bias[packed]", L.shift, bc.dbias, L.cout);     // scale stays 1, shift is the bias
        }
    }
    for (size_t i = 0; i < n.ops.size(); ++i)
        if (n.ops[i].type == OP_UP) { stage = (int)i; add(n.ops[i].wkey, n.ops[i].w_dev, h->bwd_up_dw[i], (size_t)n.tensors[n.ops[i].src].C * 16); }
    stage = heads_stage;
    add("head.1x1.weight[packed]", h->hp.w, h->bwd_hdw, (size_t)kNumOut * kStemC);
    add("head.1x1.bias[packed]", h->hp.bias, h->bwd_hdbias, kNumOut);
    add("head.attention.0.weight[packed]", h->hp.att_w, h->bwd_datt_w, (size_t)kNumStems * kNumAff * kStemC);
    add("head.attention.1.weight[packed]", h->att_gamma, h->bwd_datt_gamma, (size_t)kNumStems * kNumAff);
    add("head.attention.1.bias[packed]", h->att_beta, h->bwd_datt_beta, (size_t)kNumStems * kNumAff);
    add("head.weight_[packed]", h->hp.bank_w, h->bwd_dbank_w, (size_t)kNumStems * kNumAff * kStemC);
    add("head.bias_[packed]", h->hp.bank_b, h->bwd_dbank_b, (size_t)kNumStems * kNumAff * kStemC);
}

void finalize(mc_handle* h) {
    Net& n = *h->net;
    MC_CUDA(cudaSetDevice(h->device));
    const bool refresh = h->finalized;          // mc_refresh_params: same plan, same buffers, new values
    if (refresh) n.arena.begin_replay(h->fin_first, h->fin_last);
    else h->fin_first = n.arena.mark();
    struct ReplayGuard {                         // leave replay mode on every exit path
        DeviceArena& a; bool on;
        ~ReplayGuard() { if (on) { try { a.end_replay(); } catch (...) {} } }
    } guard{n.arena, refresh};
    if (h->training) h->bn_train.assign(n.convs.size(), mc_handle::BnTrain());
    if (h->backward) h->bwd_conv.assign(n.convs.size(), mc_handle::BwdConv());
    int conv_index = -1;
    for (auto& L : n.convs) {
        ++conv_index;
        std::vector<float> w, scale, shift;
        const int kk = L.k * L.k;
        for (const auto& part : L.parts) {
            const HostParam& wp = get_param(h, part.wkey);
            if (part.taps_to_k) {
                // deformable block: (Cout, Cin, 3, 3) -> the 1x1 layer's (Cout, 9 Cin) with K index = tap * Cin + c (csrc/dcn.cu)
                MC_CHECK(L.k == 1 && wp.shape.size() == 4 && wp.shape[1] * 9 == L.cin && wp.shape[2] == 3 && wp.shape[3] == 3, "shape of " + part.wkey);
                const int co = (int)wp.shape[0], ci = (int)wp.shape[1];
                const size_t base = w.size();
                w.resize(base + (size_t)co * ci * 9);
                for (int o = 0; o < co; ++o)
                    for (int c = 0; c < ci; ++c)
                        for (int t = 0; t < 9; ++t) w[base + ((size_t)o * 9 + t) * ci + c] = wp.data[((size_t)o * ci + c) * 9 + t];
            } else {
                MC_CHECK(wp.shape.size() == 4 && wp.shape[1] == L.cin && wp.shape[2] == L.k && wp.shape[3] == L.k,
                         "shape of " + part.wkey);
                w.insert(w.end(), wp.data.begin(), wp.data.end());
            }
            (void)kk;
            if (h->backward) h->bwd_conv[conv_index].part_cout.push_back((int)wp.shape[0]);
            if (!part.bn.empty() && h->training) {
                // train mode: the convolution writes its raw output (scale 1, shift 0); the BatchNorm parameters stay separate
                MC_CHECK(L.parts.size() == 1, "train mode: one BatchNorm per convolution");
                auto& bt = h->bn_train[conv_index];
                bt.C = L.cout; bt.eps = part.eps; bt.prefix = part.bn;
                bt.gamma = upload(h, get_param(h, part.bn + ".weight").data);
                bt.beta = upload(h, get_param(h, part.bn + ".bias").data);
                bt.rmean = upload(h, get_param(h, part.bn + ".running_mean").data);
                bt.rvar = upload(h, get_param(h, part.bn + ".running_var").data);
                bt.scale = (float*)n.arena.alloc(sizeof(float) * L.cout);
                bt.shift = (float*)n.arena.alloc(sizeof(float) * L.cout);
                bt.sums = (double*)n.arena.alloc(sizeof(double) * 2 * L.cout);
                for (int i = 0; i < L.cout; ++i) { scale.push_back(1.f); shift.push_back(0.f); }
            } else if (!part.bn.empty()) {
                fold_bn(h, part.bn, part.eps, true, scale, shift);
            } else {
                const HostParam& b = get_param(h, part.bias);
                MC_CHECK((int64_t)b.data.size() == wp.shape[0], "shape of " + part.bias);
                for (float v : b.data) { scale.push_back(1.f); shift.push_back(v); }
            }
            for (int64_t o = wp.shape[0]; o < part.pad_cout; ++o) {          // zero filters up to the layer's channel count
                w.insert(w.end(), (size_t)L.cin * kk, 0.f);
                scale.push_back(1.f); shift.push_back(0.f);
            }
        }
        if (h->train_tc) traintc_before_pack(h, conv_index, L, w);
        n.pack_conv(L, w, scale, shift);
    }
    for (auto& op : n.ops)
        if (op.type == OP_UP) {
            const HostParam& w = get_param(h, op.wkey);              // (C,1,4,4), dla_neck.py:58-65
            MC_CHECK(w.shape.size() == 4 && w.shape[2] == 4 && w.shape[3] == 4 && w.shape[1] == 1, "shape of " + op.wkey);
            op.w_dev = upload(h, w.data);
        }
    // AttnBatchNorm2d x 9 + the ten 1x1 convs
    std::vector<float> att_w, att_scale, att_shift, bank_w, bank_b, bn_mean, bn_inv;
    std::vector<float> tr_att_gamma, tr_att_beta, tr_att_rmean, tr_att_rvar, tr_rmean, tr_rvar;
    for (int s = 0; s < kNumStems; ++s) {
        const std::string pre = std::string("head.") + kStemNames[s] + ".1";
        const HostParam& aw = get_param(h, pre + ".attn_weights.attention.0.weight");   // (10,64,1,1)
        MC_CHECK((int)aw.data.size() == kNumAff * kStemC, "shape of attention conv");
        att_w.insert(att_w.end(), aw.data.begin(), aw.data.end());
        fold_bn(h, pre + ".attn_weights.attention.1", 1e-5f, true, att_scale, att_shift);
        if (h->training) {
            const std::string ab = pre + ".attn_weights.attention.1";
            for (int j = 0; j < kNumAff; ++j) {
                tr_att_gamma.push_back(get_param(h, ab + ".weight").data[j]); tr_att_beta.push_back(get_param(h, ab + ".bias").data[j]);
                tr_att_rmean.push_back(get_param(h, ab + ".running_mean").data[j]); tr_att_rvar.push_back(get_param(h, ab + ".running_var").data[j]);
            }
            const HostParam& rm0 = get_param(h, pre + ".running_mean");
            const HostParam& rv0 = get_param(h, pre + ".running_var");
            tr_rmean.insert(tr_rmean.end(), rm0.data.begin(), rm0.data.end());
            tr_rvar.insert(tr_rvar.end(), rv0.data.begin(), rv0.data.end());
        }
        const HostParam& bw = get_param(h, pre + ".weight_");
        const HostParam& bb = get_param(h, pre + ".bias_");
        MC_CHECK((int)bw.data.size() == kNumAff * kStemC && (int)bb.data.size() == kNumAff * kStemC, "shape of weight_/bias_");
        bank_w.insert(bank_w.end(), bw.data.begin(), bw.data.end());
        bank_b.insert(bank_b.end(), bb.data.begin(), bb.data.end());
        const HostParam& rm = get_param(h, pre + ".running_mean");
        const HostParam& rv = get_param(h, pre + ".running_var");
        for (int c = 0; c < kStemC; ++c) {
            bn_mean.push_back(rm.data[c]);
            bn_inv.push_back(1.0f / std::sqrt(rv.data[c] + 1e-3f));  // eps=0.001, monocon_heads.py:117
        }
    }
    std::vector<float> w1, b1;
    for (int p = 0; p < kNumPred; ++p) {
        const HostParam& w = get_param(h, std::string(kPredConv[p]) + ".weight");
        const HostParam& b = get_param(h, std::string(kPredConv[p]) + ".bias");
        MC_CHECK((int)w.data.size() == kPredCh[p] * kStemC && (int)b.data.size() == kPredCh[p], std::string("shape of ") + kPredConv[p]);
        w1.insert(w1.end(), w.data.begin(), w.data.end());
        b1.insert(b1.end(), b.data.begin(), b.data.end());
    }
    HeadParams& hp = h->hp;
    hp.att_w = upload(h, att_w); hp.att_scale = upload(h, att_scale); hp.att_shift = upload(h, att_shift);
    hp.bank_w = upload(h, bank_w); hp.bank_b = upload(h, bank_b);
    hp.bn_mean = upload(h, bn_mean); hp.bn_inv = upload(h, bn_inv);
    hp.w = upload(h, w1); hp.bias = upload(h, b1);
    if (h->training) {
        h->att_gamma = upload(h, tr_att_gamma); h->att_beta = upload(h, tr_att_beta);
        h->att_rmean = upload(h, tr_att_rmean); h->att_rvar = upload(h, tr_att_rvar);
        h->hbn_rmean = upload(h, tr_rmean); h->hbn_rvar = upload(h, tr_rvar);
    }
    hp.sums = (double*)n.arena.alloc(sizeof(double) * 2 * kStemTot * h->max_batch);
    {
        // fp32-accurate tensor-core mode: the stem convolution's epilogue accumulates the AttnBN instance statistics itself
        // (one read of the 1.1 GB stem tensor less); MC_STATS_FUSE=0 keeps the separate pass
        const char* e = std::getenv("MC_STATS_FUSE");
        h->stats_fused = false;
        for (auto& L : n.convs)
            if (L.dst == h->t_stems) {
                L.stats_sums = nullptr;
                if (n.dt == DT_SPLIT && !h->training && L.use_tc3 && n.tensors[h->t_stems].dt == DT_F32 && !(e && e[0] == '0')) {
                    L.stats_sums = hp.sums;
                    h->stats_fused = true;
                }
            }
    }
    hp.coefA = (float*)n.arena.alloc(sizeof(float) * kStemTot * h->max_batch);
    hp.coefB = (float*)n.arena.alloc(sizeof(float) * kStemTot * h->max_batch);
    {
        const char* e = std::getenv("MC_HEAD_TC");              // A/B knob: MC_HEAD_TC=0 keeps the SIMT head kernel
        const int HW = h->fh * h->fw;
        const DType sdt = n.tensors[h->t_stems].dt;
        if (!h->training && n.conv_impl == MC_CONV_AUTO && (n.dt == DT_BF16 || n.dt == DT_SPLIT) && head_tc_supported(sdt, HW) && !(e && e[0] == '0'))
            h->head_tc = head_tc_prepare(n, n.tensors[h->t_stems].ptr, sdt, h->max_batch, HW, w1, h->t_headz);
    }
    h->flops = 0; h->bytes = 0;
    for (auto& L : n.convs) { h->flops += L.flops_per_image; h->bytes += L.bytes_per_image; }
    if (h->backward) setup_backward(h);
    if (h->train_tc) traintc_setup(h);
    if (refresh) { guard.on = false; n.arena.end_replay(); }
    else h->fin_last = n.arena.mark();
    h->finalized = true;
    h->params.clear();
    MC_CUDA(cudaDeviceSynchronize());
}

// Optional per-stage hook (profiling): called before and after every stage with the stage index.
struct StageHook {
    virtual void before(int stage, cudaStream_t st) = 0;
    virtual void after(int stage, cudaStream_t st) = 0;
    virtual ~StageHook() {}
};

// stage list: [pack_input] + one per op of the plan (OP_HEADS expands to stats/mix/apply = one stage)
int num_stages(mc_handle* h) { return 1 + (int)h->net->ops.size(); }

std::string stage_name(mc_handle* h, int stage) {
    if (stage == 0) return "pack_input";
    const Op& op = h->net->ops[stage - 1];
    if (op.type == OP_CONV) return h->net->convs[op.conv].name;
    if (op.type == OP_POOL) return h->net->tensors[op.dst].name;
    if (op.type == OP_UP) return op.wkey;
    if (op.type == OP_DCN_COL) return h->net->tensors[op.dst].name;
    return "head.attn_norm+1x1";
}

// uint8 input of mc_*_u8: frames (B, H0, W0, 3) HWC with per-image valid sizes hw[B][2] (device)
struct U8Input { const unsigned char* img; const int* hw; int H0, W0; };

void upload_lut(mc_handle* h, const double mean[3], const double stdv[3]) {
    std::vector<float> lut(3 * 256);
    for (int c = 0; c < 3; ++c)
        for (int u = 0; u < 256; ++u) lut[c * 256 + u] = (float)(((double)u - mean[c]) / stdv[c]);   // numpy: float64, then torch.Tensor -> float32
    if (!h->d_lut) h->d_lut = (float*)h->net->arena.alloc(sizeof(float) * lut.size());
    MC_CUDA(cudaMemcpy(h->d_lut, lut.data(), sizeof(float) * lut.size(), cudaMemcpyHostToDevice));
}

void run_forward(mc_handle* h, const float* img, int B, float* const pred_out[kNumPred], cudaStream_t st,
                 StageHook* hook = nullptr, const U8Input* u8 = nullptr) {
    MC_CHECK(h->finalized, "mc_finalize_params has not been called");
    MC_CHECK(B >= 1 && B <= h->max_batch, "batch out of range");
    Net& n = *h->net;
    n.launches_last_run = 0;
    const TensorInfo& in = n.tensors[h->t_input];
    if (hook) hook->before(0, st);
    if (u8) {
        MC_CHECK(u8->H0 >= 1 && u8->W0 >= 1 && u8->H0 <= h->H && u8->W0 <= h->W, "uint8 frames larger than the engine's padded geometry");
        launch_pack_input_u8(u8->img, u8->hw, h->d_lut, in.ptr, n.dt, B, u8->H0, u8->W0, h->H, h->W, in.C, in.Wp, in.xoff, st, n.split_info(h->t_input));
    } else {
        launch_pack_input(img, in.ptr, n.dt, B, 3, h->H, h->W, in.C, in.Wp, in.xoff, st, n.split_info(h->t_input));
    }
    if (hook) hook->after(0, st);
    n.launches_last_run++;
    for (int i = 0; i < (int)n.ops.size(); ++i) {
        if (hook) hook->before(i + 1, st);
        if (n.ops[i].type != OP_HEADS) {
            n.run_ops(B, st, i, i + 1);
        } else {
            const int HW = h->fh * h->fw;
            const TensorInfo& stems = n.tensors[h->t_stems];
            if (!h->stats_fused) launch_attn_stats(stems.ptr, stems.dt, h->hp.sums, B, HW, st);
            AttnMixParams mp;
            mp.sums = h->hp.sums; mp.HW = HW;
            mp.att_w = h->hp.att_w; mp.att_scale = h->hp.att_scale; mp.att_shift = h->hp.att_shift;
            mp.bank_w = h->hp.bank_w; mp.bank_b = h->hp.bank_b; mp.bn_mean = h->hp.bn_mean; mp.bn_inv = h->hp.bn_inv;
            mp.coefA = h->hp.coefA; mp.coefB = h->hp.coefB;
            launch_attn_mix(mp, B, st);
            HeadApplyParams ap;
            ap.stems = stems.ptr; ap.coefA = h->hp.coefA; ap.coefB = h->hp.coefB; ap.w = h->hp.w; ap.bias = h->hp.bias;
            for (int p = 0; p < kNumPred; ++p) ap.out[p] = pred_out[p];
            ap.B = B; ap.HW = HW;
            if (h->head_tc) launch_head_apply_tc(*h->head_tc, ap, st);
            else launch_head_apply(ap, stems.dt, st);
            n.launches_last_run += h->stats_fused ? 2 : 3;
        }
        if (hook) hook->after(i + 1, st);
    }
}

// MonoConDetector.forward in train() mode up to the prediction maps (monocon_detector.py:53-61): batch-statistic BatchNorm
// everywhere, running statistics updated.  fp32 engine only; the raw convolution outputs are normalised in place.
void run_forward_train(mc_handle* h, const float* img, int B, float* const pred_out[kNumPred], cudaStream_t st) {
    MC_CHECK(h->finalized && h->training, "mc_finalize_params(h, 1) has not been called");
    MC_CHECK(B >= 2 && B <= h->max_batch, "train mode needs 2 <= B <= max_batch");
    Net& n = *h->net;
    n.launches_last_run = 0;
    h->grads_valid = false;
    h->last_train_B = B;
    ++h->train_generation;
    const TensorInfo& in = n.tensors[h->t_input];
    launch_pack_input(img, in.ptr, n.dt, B, 3, h->H, h->W, in.C, in.Wp, in.xoff, st);
    n.launches_last_run++;
    for (int i = 0; i < (int)n.ops.size(); ++i) {
        const Op& op = n.ops[i];
        if (op.type == OP_CONV) {
            const ConvLayer& L = n.convs[op.conv];
            const auto& bt = h->bn_train[op.conv];
            const TensorInfo& s0 = n.tensors[L.src[0]];
            const TensorInfo& d = n.tensors[L.dst];
            ConvParams p;
            std::memset(&p, 0, sizeof(p));
            p.nsrc = (int)L.src.size();
            for (int s = 0; s < p.nsrc; ++s) {
                p.src[s] = n.tensors[L.src[s]].ptr; p.srcC[s] = n.tensors[L.src[s]].C;
                p.srcWp[s] = n.tensors[L.src[s]].Wp; p.srcXoff[s] = n.tensors[L.src[s]].xoff;
            }
            p.B = B; p.Hin = s0.H; p.Win = s0.W; p.Hout = d.H; p.Wout = d.W; p.Cin = L.cin_store; p.Cout = L.cout;
            p.k = L.k; p.stride = L.stride; p.pad = L.pad;
            p.w = L.w_simt; p.scale = L.scale; p.shift = L.shift;            // 1 / 0 behind a BatchNorm, 1 / bias for the head stems
            p.dst = d.ptr;
            if (bt.C == 0) {                                                 // no BatchNorm: the plan's own epilogue
                p.residual = L.residual >= 0 ? n.tensors[L.residual].ptr : nullptr;
                p.relu = L.relu ? 1 : 0;
                launch_conv_simt(p, n.dt, st);
                n.launches_last_run++;
            } else {
                p.residual = nullptr; p.relu = 0;
                const float* res = L.residual >= 0 ? (const float*)n.tensors[L.residual].ptr : nullptr;
                if (h->backward) {                                           // the raw output survives for the backward pass
                    const auto& bc = h->bwd_conv[op.conv];
                    p.dst = bc.raw;
                    launch_conv_simt(p, n.dt, st);
                    launch_bn_train_ex(bc.raw, (float*)d.ptr, res, (long long)B * d.H * d.W, L.cout, bt.sums, bt.eps, 0.1f, bt.gamma, bt.beta,
                                       bt.rmean, bt.rvar, bt.scale, bt.shift, L.relu, bc.mean, bc.inv, st);
                } else {
                    launch_conv_simt(p, n.dt, st);
                    launch_bn_train((float*)d.ptr, res, (long long)B * d.H * d.W, L.cout, bt.sums, bt.eps, 0.1f, bt.gamma, bt.beta, bt.rmean,
                                    bt.rvar, bt.scale, bt.shift, L.relu, st);
                }
                n.launches_last_run += 4;
            }
        } else if (op.type == OP_HEADS) {
            const int HW = h->fh * h->fw;
            const TensorInfo& stems = n.tensors[h->t_stems];
            launch_attn_stats(stems.ptr, n.dt, h->hp.sums, B, HW, st);
            launch_attn_mix_train(h->hp.sums, B, HW, h->hp.att_w, h->att_gamma, h->att_beta, h->att_rmean, h->att_rvar, h->hp.bank_w,
                                  h->hp.bank_b, h->hbn_rmean, h->hbn_rvar, h->hp.coefA, h->hp.coefB, st);
            HeadApplyParams ap;
            ap.stems = stems.ptr; ap.coefA = h->hp.coefA; ap.coefB = h->hp.coefB; ap.w = h->hp.w; ap.bias = h->hp.bias;
            for (int p = 0; p < kNumPred; ++p) ap.out[p] = pred_out[p];
            ap.B = B; ap.HW = HW;
            launch_head_apply(ap, n.dt, st);
            n.launches_last_run += 3;
        } else {
            n.run_ops(B, st, i, i + 1);
        }
    }
}

void run_decode(mc_handle* h, const float* const pred[kNumPred], int B, const float* P2, const float* invP, int img_h,
                int img_w, int topk, float thres, float* box2d, float* box3d, long long* labels, long long* inds,
                unsigned char* valid, cudaStream_t st, const GatherParams* gather = nullptr) {
    MC_CHECK(B >= 1 && B <= h->max_batch, "batch out of range");
    DecodeParams p;
    std::memset(&p.gather, 0, sizeof(p.gather));
    if (gather) p.gather = *gather;
    for (int i = 0; i < kNumPred; ++i) p.pred[i] = pred[i];
    p.B = B; p.C = 3; p.H = h->fh; p.W = h->fw;
    p.P2 = P2; p.invP = invP;
    p.scale_x = (float)img_w / (float)h->fw;
    p.scale_y = (float)img_h / (float)h->fh;
    p.topk = topk; p.thres = thres; p.num_bins = 12; p.c2k_channels = 18;
    p.box2d = box2d; p.box3d = box3d; p.labels = labels; p.inds = inds; p.valid = valid;
    launch_decode(p, h->cand, h->cand_count, st);
    h->net->launches_last_run += 2;
}

void ensure_staging(mc_handle* h, int topk) {
    if (h->staging_topk >= topk) return;
    auto& a = h->net->arena;
    const size_t B = h->max_batch;
    if (!h->d_img) {
        h->d_img = (float*)a.alloc(sizeof(float) * B * 3 * h->H * h->W);
        h->d_P2 = (float*)a.alloc(sizeof(float) * B * 12);
        h->d_invP = (float*)a.alloc(sizeof(float) * B * 16);
    }
    h->d_box2d = (float*)a.alloc(sizeof(float) * B * topk * 5);
    h->d_box3d = (float*)a.alloc(sizeof(float) * B * topk * 7);
    h->d_labels = (long long*)a.alloc(sizeof(long long) * B * topk);
    h->d_inds = (long long*)a.alloc(sizeof(long long) * B * topk);
    h->d_valid = (unsigned char*)a.alloc(B * topk);
    h->staging_topk = topk;
}

void infer_device(mc_handle* h, const float* img, int B, const float* P2, const float* invP, int topk, float thres,
                  float* box2d, float* box3d, long long* labels, long long* inds, unsigned char* valid, cudaStream_t st,
                  const GatherParams* gather = nullptr, const U8Input* u8 = nullptr) {
    auto body = [&](cudaStream_t s) {
        run_forward(h, img, B, h->pred_own, s, nullptr, u8);
        run_decode(h, h->pred_own, B, P2, invP, h->H, h->W, topk, thres, box2d, box3d, labels, inds, valid, s, gather);
        h->launches = h->net->launches_last_run;
    };
    if (!h->use_graph) {
        body(st);
        return;
    }
    mc_handle::GraphKey key;
    std::memset(&key, 0, sizeof(key));
    key.img = u8 ? (const void*)u8->img : (const void*)img; key.hw = u8 ? (const void*)u8->hw : nullptr;
    key.H0 = u8 ? u8->H0 : 0; key.W0 = u8 ? u8->W0 : 0;
    key.P2 = P2; key.invP = invP; key.b2 = box2d; key.b3 = box3d; key.lb = labels; key.ix = inds; key.vl = valid;
    key.gather = gather ? (const void*)gather->gen : nullptr;
    key.B = B; key.topk = topk; key.thres = thres;
    cudaGraphExec_t exec = nullptr;
    for (auto& kv : h->graphs)
        if (kv.first == key) exec = kv.second;
    if (!exec) {
        if (h->graphs.size() >= 16) {
            cudaGraphExecDestroy(h->graphs.front().second);
            h->graphs.erase(h->graphs.begin());
        }
        cudaStream_t cs;
        MC_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        cudaGraph_t graph = nullptr;
        MC_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
        try {
            body(cs);
        } catch (...) {
            cudaStreamEndCapture(cs, &graph);
            if (graph) cudaGraphDestroy(graph);
            cudaStreamDestroy(cs);
            throw;
        }
        MC_CUDA(cudaStreamEndCapture(cs, &graph));
        MC_CUDA(cudaGraphInstantiate(&exec, graph, 0));
        cudaGraphDestroy(graph);
        cudaStreamDestroy(cs);
        h->graphs.push_back({key, exec});
    }
    MC_CUDA(cudaGraphLaunch(exec, st));
}

template <typename F>
int guarded(mc_handle* h, F&& f) {
    try {
        if (h) MC_CUDA(cudaSetDevice(h->device));
        f();
        return 0;
    } catch (const std::exception& e) {
        if (h) h->err = e.what();
        else { std::lock_guard<std::mutex> g(g_mutex); g_create_error = e.what(); }
        return 1;
    }
}

}  // namespace

// =============================================================================================
extern "C" {

int mc_create(mc_handle** out, int device, int max_batch, int H, int W, int precision_mode) {
    return mc_create_ex(out, device, max_batch, H, W, precision_mode, MC_NECK_CONV);
}

int mc_create_ex(mc_handle** out, int device, int max_batch, int H, int W, int precision_mode, int neck_variant) {
    if (!out) return 1;
    *out = nullptr;
    mc_handle* h = new mc_handle();
    int rc = guarded(nullptr, [&]() {
        MC_CHECK(max_batch >= 1 && max_batch <= 1024, "max_batch");
        MC_CHECK(H >= 32 && W >= 32 && H % 32 == 0 && W % 32 == 0, "H and W must be multiples of 32");
        MC_CHECK(precision_mode == MC_PREC_BF16 || precision_mode == MC_PREC_FP32 || precision_mode == MC_PREC_FP32_TC, "precision_mode");
        MC_CHECK(neck_variant == MC_NECK_CONV || neck_variant == MC_NECK_DCN, "neck_variant");
        h->neck = neck_variant;
        int ndev = 0;
        MC_CUDA(cudaGetDeviceCount(&ndev));
        MC_CHECK(device >= 0 && device < ndev, "no such CUDA device");
        MC_CUDA(cudaSetDevice(device));
        cudaDeviceProp prop;
        MC_CUDA(cudaGetDeviceProperties(&prop, device));
        MC_CHECK(prop.major == 10, std::string("this library is built for sm_100a (B200); found ") + prop.name);
        h->device = device; h->max_batch = max_batch; h->H = H; h->W = W; h->prec = precision_mode;
        h->dt = precision_mode == MC_PREC_FP32 ? DT_F32 : (precision_mode == MC_PREC_FP32_TC ? DT_SPLIT : DT_BF16);
        h->net.reset(new Net(device, max_batch, h->dt, MC_CONV_AUTO));
        head_kernels_init();
        tc_kernels_init();
        tc2_kernels_init();
        tc3_kernels_init();
        dcn_tc_init();
        head_tc_init();
        build_plan(h);
        h->net->allocate();
        const size_t HW = (size_t)h->fh * h->fw;
        for (int p = 0; p < kNumPred; ++p)
            h->pred_own[p] = (float*)h->net->arena.alloc(sizeof(float) * max_batch * kPredCh[p] * HW);
        h->cand = (unsigned long long*)h->net->arena.alloc(sizeof(unsigned long long) * max_batch * 3 * HW);
        h->cand_count = (int*)h->net->arena.alloc(sizeof(int) * max_batch);
        const double mean[3] = {123.675, 116.28, 103.53}, stdv[3] = {58.395, 57.12, 57.375};    // dataset/monocon_dataset.py:32,39
        upload_lut(h, mean, stdv);
    });
    if (rc) { delete h; return rc; }
    *out = h;
    return 0;
}

int mc_set_param(mc_handle* h, const char* key, const float* data, const int64_t* shape, int ndim) {
    if (!h) return 1;
    return guarded(h, [&]() {
        MC_CHECK(key && data && ndim >= 0 && ndim <= 8, "mc_set_param arguments");
        // after mc_finalize_params the staged tensors wait for mc_refresh_params
        HostParam p;
        size_t n = 1;
        for (int i = 0; i < ndim; ++i) { p.shape.push_back(shape[i]); n *= (size_t)shape[i]; }
        p.data.resize(n);
        MC_CUDA(cudaMemcpy(p.data.data(), data, sizeof(float) * n, cudaMemcpyDefault));
        h->params[key] = std::move(p);
    });
}

int mc_finalize_params(mc_handle* h, int training) {
    if (!h) return 1;
    return guarded(h, [&]() {
        MC_CHECK(training >= 0 && training <= 2, "training flag: 0 inference, 1 train-mode forward, 2 forward + backward");
        MC_CHECK(!h->finalized, "parameters are already finalized: mc_refresh_params repacks new values, a new handle changes the mode");
        MC_CHECK(!training || h->neck == MC_NECK_CONV, "the deformable (MC_NECK_DCN) neck is an inference plan: no train-mode kernels for the deformable columns");
        h->fin_training = training;
        h->backward = training == 2;
        if (training) {
            MC_CHECK(h->dt == DT_F32 || h->dt == DT_BF16, "train-mode engines: MC_PREC_FP32 (FFMA, the strict twin) or MC_PREC_BF16 (tensor cores, bf16 mixed precision)");
            if (h->dt == DT_BF16) {
                h->train_tc = traintc_create();
                h->net->keep_master = true;      // the optimiser steps fp32 master weights; the bf16 plans are repacked from them every forward
            } else {
                h->net->conv_impl = MC_CONV_SIMT;
            }
            h->training = true;
        }
        finalize(h);
    });
}

int mc_refresh_params(mc_handle* h) {
    if (!h) return 1;
    return guarded(h, [&]() {
        MC_CHECK(h->finalized, "mc_refresh_params follows mc_finalize_params");
        MC_CHECK(!h->params.empty(), "mc_refresh_params: stage the new tensors with mc_set_param first (all of them)");
        MC_CUDA(cudaDeviceSynchronize());       // nothing may still read the buffers that are about to be rewritten
        h->grads_valid = false;
        finalize(h);
    });
}

int mc_forward(mc_handle* h, const float* img, int B, float* const pred_out[MC_NUM_PRED], void* stream) {
    if (!h) return 1;
    return guarded(h, [&]() {
        run_forward(h, img, B, pred_out, (cudaStream_t)stream);
        h->launches = h->net->launches_last_run;
    });
}

int mc_forward_train(mc_handle* h, const float* img, int B, float* const pred_out[MC_NUM_PRED], void* stream) {
    if (!h) return 1;
    return guarded(h, [&]() {
        if (h->train_tc) {
            MC_CHECK(h->finalized && h->training, "mc_finalize_params(h, 1 | 2) has not been called");
            MC_CHECK(B >= 2 && B <= h->max_batch, "train mode needs 2 <= B <= max_batch");
            h->grads_valid = false;
            h->last_train_B = B;
            ++h->train_generation;
            traintc_forward(h, img, B, pred_out, (cudaStream_t)stream);
        } else {
            run_forward_train(h, img, B, pred_out, (cudaStream_t)stream);
        }
        h->launches = h->net->launches_last_run;
    });
}

long long mc_train_generation(const mc_handle* h) { return h ? h->train_generation : -1; }

int mc_get_buffer(mc_handle* h, const char* key, float* out_host, int n) {
    if (!h) return 1;
    return guarded(h, [&]() {
        MC_CHECK(h->training && key && out_host, "train-mode engine / arguments");
        const std::string k(key);
        const float* src = nullptr;
        int len = 0;
        for (const auto& bt : h->bn_train) {
            if (bt.C == 0) continue;
            if (k == bt.prefix + ".running_mean") { src = bt.rmean; len = bt.C; }
            if (k == bt.prefix + ".running_var") { src = bt.rvar; len = bt.C; }
        }
        for (int s = 0; s < kNumStems && !src; ++s) {
            const std::string pre = std::string("head.") + kStemNames[s] + ".1";
            if (k == pre + ".running_mean") { src = h->hbn_rmean + s * kStemC; len = kStemC; }
            if (k == pre + ".running_var") { src = h->hbn_rvar + s * kStemC; len = kStemC; }
            if (k == pre + ".attn_weights.attention.1.running_mean") { src = h->att_rmean + s * kNumAff; len = kNumAff; }
            if (k == pre + ".attn_weights.attention.1.running_var") { src = h->att_rvar + s * kNumAff; len = kNumAff; }
        }
        if (!src) throw Error("no such BatchNorm buffer in the plan: " + k);
        MC_CHECK(n == len, "buffer length");
        MC_CUDA(cudaDeviceSynchronize());
        MC_CUDA(cudaMemcpy(out_host, src, sizeof(float) * len, cudaMemcpyDeviceToHost));
    });
}

int mc_backward_train(mc_handle* h, const float* const pred[MC_NUM_PRED], const float* const dpred[MC_NUM_PRED], int B, void* stream) {
    if (!h) return 1;
    return guarded(h, [&]() {
        MC_CHECK(h->finalized && h->backward, "mc_finalize_params(h, 2) has not been called");
        MC_CHECK(B == h->last_train_B && B >= 2, "mc_backward_train follows mc_forward_train of the same batch");
        for (int i = 0; i < kNumPred; ++i) {
            MC_CHECK(pred[i] && dpred[i], "mc_backward_train: null map");
            h->bwd_hargs.pred[i] = pred[i]; h->bwd_hargs.dpred[i] = dpred[i];
        }
        if (h->train_tc) traintc_backward(h, B, 0, (int)h->bwd_ops.size(), true, (cudaStream_t)stream);
        else if (mc_bw_run_graph(h->bwd_tensors.data(), (int)h->bwd_tensors.size(), h->bwd_ops.data(), (int)h->bwd_ops.size(), B, stream))
            throw Error(std::string("backward: ") + mc_bw_last_error());
        h->grads_valid = true;
    });
}

// One parameter of the plan in the reference's state_dict layout: its gradient (what = 0) or its current value (what = 1; the
// engine-resident optimiser steps the packed buffers in place, this is how the module gets them back).
static int fetch_parameter(mc_handle* h, const char* key, float* out_host, int64_t n, int what) {
    if (!h) return 1;
    return guarded(h, [&]() {
        MC_CHECK(h->backward && key && out_host, "needs an engine finalized with mc_finalize_params(h, 2)");
        MC_CHECK(what == 1 || h->grads_valid, "mc_get_grad follows mc_backward_train");
        const bool val = what == 1;
        MC_CUDA(cudaDeviceSynchronize());
        const std::string k(key);
        Net& net = *h->net;
        auto copy = [&](const float* dev, size_t len) {
            MC_CHECK((size_t)n == len, "gradient length of " + k);
            MC_CUDA(cudaMemcpy(out_host, dev, sizeof(float) * len, cudaMemcpyDeviceToHost));
        };
        for (size_t i = 0; i < net.convs.size(); ++i) {
            const ConvLayer& L = net.convs[i];
            const auto& bc = h->bwd_conv[i];
            int o0 = 0;
            for (size_t j = 0; j < L.parts.size(); ++j) {
                const auto& part = L.parts[j];
                const int co = bc.part_cout[j];
                if (k == part.wkey) {                        // [k*k][cin_store][cout] -> OIHW with the logical cin
                    const int kk = L.k * L.k;
                    MC_CHECK((size_t)n == (size_t)co * L.cin * kk, "gradient length of " + k);
                    std::vector<float> packed((size_t)kk * L.cin_store * L.cout);
                    MC_CUDA(cudaMemcpy(packed.data(), val ? L.w_simt : bc.dw, sizeof(float) * packed.size(), cudaMemcpyDeviceToHost));
                    for (int o = 0; o < co; ++o)
                        for (int c = 0; c < L.cin; ++c)
                            for (int t = 0; t < kk; ++t)
                                out_host[((size_t)o * L.cin + c) * kk + t] = packed[((size_t)t * L.cin_store + c) * L.cout + o0 + o];
                    return;
                }
                if (!part.bias.empty() && k == part.bias) { copy((val ? L.shift : bc.dbias) + o0, co); return; }
                if (!part.bn.empty() && k == part.bn + ".weight") { copy(val ? h->bn_train[i].gamma : bc.dgamma, L.cout); return; }
                if (!part.bn.empty() && k == part.bn + ".bias") { copy(val ? h->bn_train[i].beta : bc.dbeta, L.cout); return; }
                o0 += co;
            }
        }
        for (size_t i = 0; i < net.ops.size(); ++i)
            if (net.ops[i].type == OP_UP && k == net.ops[i].wkey) { copy(val ? net.ops[i].w_dev : h->bwd_up_dw[i], (size_t)net.tensors[net.ops[i].src].C * 16); return; }
        int o0 = 0;
        for (int p = 0; p < kNumPred; ++p) {
            if (k == std::string(kPredConv[p]) + ".weight") { copy((val ? h->hp.w : h->bwd_hdw) + (size_t)o0 * kStemC, (size_t)kPredCh[p] * kStemC); return; }
            if (k == std::string(kPredConv[p]) + ".bias") { copy((val ? h->hp.bias : h->bwd_hdbias) + o0, kPredCh[p]); return; }
            o0 += kPredCh[p];
        }
        for (int s = 0; s < kNumStems; ++s) {
            const std::string pre = std::string("head.") + kStemNames[s] + ".1";
            const size_t bank = (size_t)s * kNumAff * kStemC, aff = (size_t)s * kNumAff;
            if (k == pre + ".attn_weights.attention.0.weight") { copy((val ? h->hp.att_w : h->bwd_datt_w) + bank, kNumAff * kStemC); return; }
            if (k == pre + ".attn_weights.attention.1.weight") { copy((val ? h->att_gamma : h->bwd_datt_gamma) + aff, kNumAff); return; }
            if (k == pre + ".attn_weights.attention.1.bias") { copy((val ? h->att_beta : h->bwd_datt_beta) + aff, kNumAff); return; }
            if (k == pre + ".weight_") { copy((val ? h->hp.bank_w : h->bwd_dbank_w) + bank, kNumAff * kStemC); return; }
            if (k == pre + ".bias_") { copy((val ? h->hp.bank_b : h->bwd_dbank_b) + bank, kNumAff * kStemC); return; }
        }
        throw Error("no gradient for this key (not a parameter of the plan; the outer `project` tensors of level3/level4 get none in the reference either): " + k);
    });
}

int mc_num_backward_stages(mc_handle* h) { return (h && h->backward && h->finalized) ? (int)h->bwd_ops.size() : -1; }

int mc_backward_train_segment(mc_handle* h, const float* const pred[MC_NUM_PRED], const float* const dpred[MC_NUM_PRED], int B, int op_first,
                              int op_last, void* stream) {
    if (!h) return 1;
    return guarded(h, [&]() {
        MC_CHECK(h->finalized && h->backward, "mc_finalize_params(h, 2) has not been called");
        MC_CHECK(B == h->last_train_B && B >= 2, "mc_backward_train_segment follows mc_forward_train of the same batch");
        const int nops = (int)h->bwd_ops.size();
        for (int i = 0; i < kNumPred; ++i) {
            MC_CHECK(pred[i] && dpred[i], "mc_backward_train_segment: null map");
            h->bwd_hargs.pred[i] = pred[i]; h->bwd_hargs.dpred[i] = dpred[i];
        }
        if (h->train_tc) {
            MC_CHECK(0 <= op_first && op_first <= op_last && op_last <= nops, "mc_backward_train_segment: 0 <= op_first <= op_last <= stages");
            traintc_backward(h, B, op_first, op_last, op_last == nops, (cudaStream_t)stream);
        } else if (mc_bw_run_graph_range(h->bwd_tensors.data(), (int)h->bwd_tensors.size(), h->bwd_ops.data(), nops, B, op_first, op_last,
                                         op_last == nops ? 1 : 0, stream))
            throw Error(std::string("backward: ") + mc_bw_last_error());
        if (op_first == 0) h->grads_valid = true;
    });
}

int mc_get_grad(mc_handle* h, const char* key, float* out_host, int64_t n) { return fetch_parameter(h, key, out_host, n, 0); }
int mc_get_param(mc_handle* h, const char* key, float* out_host, int64_t n) { return fetch_parameter(h, key, out_host, n, 1); }

int mc_debug_bw_graph(mc_handle* h, const mc_bw_tensor** tensors, int* n_tensors, const mc_bw_op** ops, int* n_ops) {
    if (!h) return 1;
    return guarded(h, [&]() {
        MC_CHECK(h->backward && h->finalized && tensors && n_tensors && ops && n_ops, "mc_debug_bw_graph: backward-enabled engine");
        *tensors = h->bwd_tensors.data(); *n_tensors = (int)h->bwd_tensors.size();
        *ops = h->bwd_ops.data(); *n_ops = (int)h->bwd_ops.size();
    });
}

int mc_debug_train_dump(mc_handle* h, int kind, int index, int B, float* out_nchw, void* stream) {
    if (!h) return 1;
    return guarded(h, [&]() {
        MC_CHECK(h->backward && h->finalized && out_nchw && B >= 1 && B <= h->max_batch, "mc_debug_train_dump: backward-enabled engine");
        Net& n = *h->net;
        const void* ptr = nullptr;
        DType dt = n.dt;
        int C = 0, H = 0, W = 0;
        if (kind == 0 || kind == 1) {
            MC_CHECK(index >= 0 && index < (int)n.tensors.size(), "mc_debug_train_dump: tensor index");
            const TensorInfo& t = n.tensors[index];
            MC_CHECK(t.Wp == t.W, "padded tensors cannot be dumped");
            C = t.C; H = t.H; W = t.W;
            if (kind == 0) { ptr = t.ptr; dt = t.dt; }
            else if (h->train_tc && index != h->t_stems) traintc_debug(h, 1, index, &ptr, &dt, &C, &H, &W);
            else { ptr = h->bwd_g[index]; dt = DT_F32; }
        } else {
            MC_CHECK(index >= 0 && index < (int)n.ops.size() && n.ops[index].type == OP_CONV, "mc_debug_train_dump: index of a convolution stage");
            const int conv = n.ops[index].conv;
            if (h->train_tc) {
                traintc_debug(h, kind, conv, &ptr, &dt, &C, &H, &W);
            } else {
                const TensorInfo& d = n.tensors[n.convs[conv].dst];
                MC_CHECK(kind == 2, "mc_debug_train_dump: the fp32 engine keeps one shared scratch for the raw-output gradients");
                ptr = h->bwd_conv[conv].raw ? (const void*)h->bwd_conv[conv].raw : d.ptr; dt = DT_F32; C = d.C; H = d.H; W = d.W;
            }
        }
        MC_CHECK(ptr != nullptr, "mc_debug_train_dump: no such buffer");
        launch_unpack_nchw(ptr, dt, out_nchw, B, C, H, W, (cudaStream_t)stream);
    });
}

int mc_num_train_tensors(mc_handle* h) { return (h && h->backward && h->finalized) ? (int)h->train_tensors.size() : -1; }

int mc_train_tensor(mc_handle* h, int i, float** param, float** grad, int64_t* numel, int* stage, char* key, int key_cap) {
    if (!h) return 1;
    return guarded(h, [&]() {
        MC_CHECK(h->backward && h->finalized && i >= 0 && i < (int)h->train_tensors.size(), "mc_train_tensor: index");
        const auto& t = h->train_tensors[i];
        if (param) *param = t.param;
        if (grad) *grad = t.grad;
        if (numel) *numel = t.numel;
        if (stage) *stage = t.stage;
        if (key && key_cap > 0) { std::strncpy(key, t.key.c_str(), key_cap - 1); key[key_cap - 1] = 0; }
    });
}

int mc_decode(mc_handle* h, const float* const pred[MC_NUM_PRED], int B, const float* P2, const float* invP, int img_h,
              int img_w, int topk, float thres, float* box2d, float* box3d, int64_t* labels, int64_t* inds,
              uint8_t* valid, void* stream) {
    if (!h) return 1;
    return guarded(h, [&]() {
        run_decode(h, pred, B, P2, invP, img_h, img_w, topk, thres, box2d, box3d, (long long*)labels, (long long*)inds,
                   valid, (cudaStream_t)stream);
    });
}

int mc_infer_device(mc_handle* h, const float* img, int B, const float* P2, const float* invP, int topk, float thres,
                    float* box2d, float* box3d, int64_t* labels, int64_t* inds, uint8_t* valid, void* stream) {
    if (!h) return 1;
    return guarded(h, [&]() {
        infer_device(h, img, B, P2, invP, topk, thres, box2d, box3d, (long long*)labels, (long long*)inds, valid,
                     (cudaStream_t)stream);
    });
}

int mc_infer_host(mc_handle* h, const float* img_host, int B, const float* P2_host, const float* invP_host, int topk,
                  float thres, float* box2d_host, float* box3d_host, int64_t* labels_host, int64_t* inds_host,
                  uint8_t* valid_host, void* stream) {
    if (!h) return 1;
    return guarded(h, [&]() {
        MC_CHECK(B >= 1 && B <= h->max_batch, "batch out of range");
        cudaStream_t st = (cudaStream_t)stream;
        ensure_staging(h, topk);
        MC_CUDA(cudaMemcpyAsync(h->d_img, img_host, sizeof(float) * (size_t)B * 3 * h->H * h->W, cudaMemcpyHostToDevice, st));
        MC_CUDA(cudaMemcpyAsync(h->d_P2, P2_host, sizeof(float) * B * 12, cudaMemcpyHostToDevice, st));
        MC_CUDA(cudaMemcpyAsync(h->d_invP, invP_host, sizeof(float) * B * 16, cudaMemcpyHostToDevice, st));
        infer_device(h, h->d_img, B, h->d_P2, h->d_invP, topk, thres, h->d_box2d, h->d_box3d, h->d_labels, h->d_inds,
                     h->d_valid, st);
        const size_t n = (size_t)B * topk;
        MC_CUDA(cudaMemcpyAsync(box2d_host, h->d_box2d, sizeof(float) * n * 5, cudaMemcpyDeviceToHost, st));
        MC_CUDA(cudaMemcpyAsync(box3d_host, h->d_box3d, sizeof(float) * n * 7, cudaMemcpyDeviceToHost, st));
        MC_CUDA(cudaMemcpyAsync(labels_host, h->d_labels, sizeof(long long) * n, cudaMemcpyDeviceToHost, st));
        MC_CUDA(cudaMemcpyAsync(inds_host, h->d_inds, sizeof(long long) * n, cudaMemcpyDeviceToHost, st));
        MC_CUDA(cudaMemcpyAsync(valid_host, h->d_valid, n, cudaMemcpyDeviceToHost, st));
        MC_CUDA(cudaStreamSynchronize(st));
    });
}

// shared body of mc_infer_host_submit / mc_infer_host_u8_submit: H2D on the copy stream -> forward + decode on the compute
// stream -> D2H on the read-back stream, all asynchronous, one pinned-host slot of two
static void host_submit(mc_handle* h, int slot, const float* img_host, const unsigned char* u8_host, const int* hw_host, int H0, int W0, int B,
                        const float* P2_host, const float* invP_host, int topk, float thres, float* box2d_host, float* box3d_host,
                        int64_t* labels_host, int64_t* inds_host, uint8_t* valid_host) {
    MC_CHECK(slot == 0 || slot == 1, "slot must be 0 or 1");
    MC_CHECK(B >= 1 && B <= h->max_batch, "batch out of range");
    MC_CHECK(topk >= 1 && topk <= 128, "topk");
    auto& S = h->slots[slot];
    MC_CHECK(!S.busy, "slot is still in flight: call mc_infer_host_wait first");
    if (!h->st_h2d) {
        MC_CUDA(cudaStreamCreateWithFlags(&h->st_h2d, cudaStreamNonBlocking));
        MC_CUDA(cudaStreamCreateWithFlags(&h->st_comp, cudaStreamNonBlocking));
        MC_CUDA(cudaStreamCreateWithFlags(&h->st_d2h, cudaStreamNonBlocking));
    }
    auto& a = h->net->arena;
    const size_t MB = h->max_batch;
    if (!S.d_img) {
        S.d_img = (float*)a.alloc(sizeof(float) * MB * 3 * h->H * h->W);       // also holds the (4x smaller) uint8 frames
        S.d_P2 = (float*)a.alloc(sizeof(float) * MB * 12);
        S.d_invP = (float*)a.alloc(sizeof(float) * MB * 16);
        S.d_hw = (int*)a.alloc(sizeof(int) * MB * 2);
        MC_CUDA(cudaEventCreateWithFlags(&S.ev_in, cudaEventDisableTiming));
        MC_CUDA(cudaEventCreateWithFlags(&S.ev_done, cudaEventDisableTiming));
        MC_CUDA(cudaEventCreateWithFlags(&S.ev_out, cudaEventDisableTiming));
    }
    if (S.topk < topk) {
        S.d_b2 = (float*)a.alloc(sizeof(float) * MB * topk * 5);
        S.d_b3 = (float*)a.alloc(sizeof(float) * MB * topk * 7);
        S.d_lb = (long long*)a.alloc(sizeof(long long) * MB * topk);
        S.d_ix = (long long*)a.alloc(sizeof(long long) * MB * topk);
        S.d_vl = (unsigned char*)a.alloc(MB * topk);
        S.topk = topk;
    }
    // copy engine: inputs of this slot (its previous compute has been waited for by the caller)
    if (u8_host) {
        MC_CHECK(hw_host && H0 >= 1 && W0 >= 1 && H0 <= h->H && W0 <= h->W, "uint8 frames larger than the engine's padded geometry");
        MC_CUDA(cudaMemcpyAsync(S.d_img, u8_host, (size_t)B * H0 * W0 * 3, cudaMemcpyHostToDevice, h->st_h2d));
        MC_CUDA(cudaMemcpyAsync(S.d_hw, hw_host, sizeof(int) * B * 2, cudaMemcpyHostToDevice, h->st_h2d));
    } else {
        MC_CUDA(cudaMemcpyAsync(S.d_img, img_host, sizeof(float) * (size_t)B * 3 * h->H * h->W, cudaMemcpyHostToDevice, h->st_h2d));
    }
    MC_CUDA(cudaMemcpyAsync(S.d_P2, P2_host, sizeof(float) * B * 12, cudaMemcpyHostToDevice, h->st_h2d));
    MC_CUDA(cudaMemcpyAsync(S.d_invP, invP_host, sizeof(float) * B * 16, cudaMemcpyHostToDevice, h->st_h2d));
    MC_CUDA(cudaEventRecord(S.ev_in, h->st_h2d));
    // compute stream: both slots share the activation arena, so their forwards serialise here
    MC_CUDA(cudaStreamWaitEvent(h->st_comp, S.ev_in, 0));
    if (u8_host) {
        U8Input u8{reinterpret_cast<const unsigned char*>(S.d_img), S.d_hw, H0, W0};
        infer_device(h, nullptr, B, S.d_P2, S.d_invP, topk, thres, S.d_b2, S.d_b3, S.d_lb, S.d_ix, S.d_vl, h->st_comp, nullptr, &u8);
    } else {
        infer_device(h, S.d_img, B, S.d_P2, S.d_invP, topk, thres, S.d_b2, S.d_b3, S.d_lb, S.d_ix, S.d_vl, h->st_comp);
    }
    MC_CUDA(cudaEventRecord(S.ev_done, h->st_comp));
    // read-back
    MC_CUDA(cudaStreamWaitEvent(h->st_d2h, S.ev_done, 0));
    const size_t n = (size_t)B * topk;
    MC_CUDA(cudaMemcpyAsync(box2d_host, S.d_b2, sizeof(float) * n * 5, cudaMemcpyDeviceToHost, h->st_d2h));
    MC_CUDA(cudaMemcpyAsync(box3d_host, S.d_b3, sizeof(float) * n * 7, cudaMemcpyDeviceToHost, h->st_d2h));
    MC_CUDA(cudaMemcpyAsync(labels_host, S.d_lb, sizeof(long long) * n, cudaMemcpyDeviceToHost, h->st_d2h));
    MC_CUDA(cudaMemcpyAsync(inds_host, S.d_ix, sizeof(long long) * n, cudaMemcpyDeviceToHost, h->st_d2h));
    MC_CUDA(cudaMemcpyAsync(valid_host, S.d_vl, n, cudaMemcpyDeviceToHost, h->st_d2h));
    MC_CUDA(cudaEventRecord(S.ev_out, h->st_d2h));
    S.busy = true;
}

int mc_infer_host_submit(mc_handle* h, int slot, const float* img_host, int B, const float* P2_host, const float* invP_host,
                         int topk, float thres, float* box2d_host, float* box3d_host, int64_t* labels_host, int64_t* inds_host,
                         uint8_t* valid_host) {
    if (!h) return 1;
    return guarded(h, [&]() {
        MC_CHECK(img_host != nullptr, "img");
        host_submit(h, slot, img_host, nullptr, nullptr, 0, 0, B, P2_host, invP_host, topk, thres, box2d_host, box3d_host, labels_host,
                    inds_host, valid_host);
    });
}

int mc_infer_host_u8_submit(mc_handle* h, int slot, const uint8_t* img_hwc_host, const int32_t* hw_host, int B, int H0, int W0,
                            const float* P2_host, const float* invP_host, int topk, float thres, float* box2d_host, float* box3d_host,
                            int64_t* labels_host, int64_t* inds_host, uint8_t* valid_host) {
    if (!h) return 1;
    return guarded(h, [&]() {
        MC_CHECK(img_hwc_host != nullptr && hw_host != nullptr, "img / hw");
        host_submit(h, slot, nullptr, img_hwc_host, hw_host, H0, W0, B, P2_host, invP_host, topk, thres, box2d_host, box3d_host, labels_host,
                    inds_host, valid_host);
    });
}

int mc_infer_host_wait(mc_handle* h, int slot) {
    if (!h) return 1;
    return guarded(h, [&]() {
        MC_CHECK(slot == 0 || slot == 1, "slot must be 0 or 1");
        auto& S = h->slots[slot];
        MC_CHECK(S.busy, "nothing was submitted on this slot");
        MC_CUDA(cudaEventSynchronize(S.ev_out));
        S.busy = false;
    });
}

int mc_kitti_boxes(int device, const float* box3d, const uint8_t* valid, const float* P2, const int32_t* img_hw, int B, int K,
                   double* bbox_out, float* alpha_out, uint8_t* keep_out, void* stream) {
    return guarded(nullptr, [&]() {
        MC_CUDA(cudaSetDevice(device));
        MC_CHECK(box3d && valid && P2 && img_hw && bbox_out && alpha_out && keep_out && B >= 1 && K >= 1, "arguments");
        launch_kitti_boxes(box3d, valid, P2, img_hw, B, K, bbox_out, alpha_out, keep_out, (cudaStream_t)stream);
    });
}

int mc_set_normalization(mc_handle* h, const double mean[3], const double stdv[3]) {
    if (!h) return 1;
    return guarded(h, [&]() {
        MC_CHECK(mean && stdv && stdv[0] != 0 && stdv[1] != 0 && stdv[2] != 0, "mean / std");
        MC_CUDA(cudaDeviceSynchronize());
        upload_lut(h, mean, stdv);
    });
}

int mc_forward_u8(mc_handle* h, const uint8_t* img_hwc, const int32_t* hw, int B, int H0, int W0, float* const pred_out[MC_NUM_PRED],
                  void* stream) {
    if (!h) return 1;
    return guarded(h, [&]() {
        MC_CHECK(img_hwc && hw, "img / hw");
        U8Input u8{img_hwc, hw, H0, W0};
        run_forward(h, nullptr, B, pred_out, (cudaStream_t)stream, nullptr, &u8);
        h->launches = h->net->launches_last_run;
    });
}

int mc_infer_device_u8(mc_handle* h, const uint8_t* img_hwc, const int32_t* hw, int B, int H0, int W0, const float* P2,
                       const float* invP, int topk, float thres, float* box2d, float* box3d, int64_t* labels, int64_t* inds,
                       uint8_t* valid, void* stream) {
    if (!h) return 1;
    return guarded(h, [&]() {
        MC_CHECK(img_hwc && hw, "img / hw");
        U8Input u8{img_hwc, hw, H0, W0};
        infer_device(h, nullptr, B, P2, invP, topk, thres, box2d, box3d, (long long*)labels, (long long*)inds, valid,
                     (cudaStream_t)stream, nullptr, &u8);
    });
}

int mc_gather_create(mc_handle* h, int world, int rank, int topk, void* ipc_handle_out) {
    if (!h) return 1;
    return guarded(h, [&]() {
        auto& G = h->gather;
        MC_CHECK(G.block == nullptr, "gather block already created");
        MC_CHECK(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "world (<= 8) / rank");
        MC_CHECK(topk >= 1 && topk <= 128 && ipc_handle_out, "topk / handle");
        G.world = world; G.rank = rank; G.topk = topk;
        // slot layout = monocon_pytorch_b200/dist.py:_field_bytes (box2d, box3d, labels, inds, valid; 16-byte aligned fields)
        const size_t n = (size_t)h->max_batch * topk;
        const size_t sizes[5] = {n * 5 * 4, n * 7 * 4, n * 8, n * 8, n};
        size_t off = 0;
        for (int i = 0; i < 5; ++i) { G.off[i] = (long long)off; off += (sizes[i] + 15) / 16 * 16; }
        G.slot_bytes = off;
        G.data_bytes = 2 * (size_t)world * G.slot_bytes;
        G.block_bytes = G.data_bytes + sizeof(unsigned) * (4 * kMaxPeers + 4);
        MC_CUDA(cudaMalloc(&G.block, G.block_bytes));        // own allocation: the IPC handle covers exactly this block
        MC_CUDA(cudaMemset(G.block, 0, G.block_bytes));
        MC_CUDA(cudaMalloc(&G.d_err, sizeof(int)));
        MC_CUDA(cudaMemset(G.d_err, 0, sizeof(int)));
        cudaIpcMemHandle_t hd;
        MC_CUDA(cudaIpcGetMemHandle(&hd, G.block));
        static_assert(sizeof(hd) == 64, "cudaIpcMemHandle_t is 64 bytes");
        std::memcpy(ipc_handle_out, &hd, sizeof(hd));
        MC_CUDA(cudaDeviceSynchronize());
    });
}

int mc_gather_connect(mc_handle* h, const void* all_handles) {
    if (!h) return 1;
    return guarded(h, [&]() {
        auto& G = h->gather;
        MC_CHECK(G.block != nullptr && !G.connected && all_handles, "mc_gather_create first / already connected");
        for (int r = 0; r < G.world; ++r) {
            if (r == G.rank) { G.peer[r] = G.block; continue; }
            cudaIpcMemHandle_t hd;
            std::memcpy(&hd, (const char*)all_handles + 64 * r, 64);
            void* ptr = nullptr;
            MC_CUDA(cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess));
            G.peer[r] = (char*)ptr;
        }
        for (int buf = 0; buf < 2; ++buf) {
            unsigned* host[kMaxPeers] = {nullptr};
            for (int r = 0; r < G.world; ++r) host[r] = g_ready_flag(G.peer[r], G.data_bytes, buf) + G.rank;
            MC_CUDA(cudaMalloc(&G.d_peer_ready[buf], sizeof(unsigned*) * kMaxPeers));
            MC_CUDA(cudaMemcpy(G.d_peer_ready[buf], host, sizeof(host), cudaMemcpyHostToDevice));
        }
        G.connected = true;
    });
}

size_t mc_gather_slot_bytes(const mc_handle* h) { return h ? h->gather.slot_bytes : 0; }

int mc_gather_buffer(mc_handle* h, int buf, void** ptr) {
    if (!h || !ptr) return 1;
    return guarded(h, [&]() {
        MC_CHECK(h->gather.block != nullptr && (buf == 0 || buf == 1), "gather block / buf");
        *ptr = h->gather.block + (size_t)buf * h->gather.world * h->gather.slot_bytes;
    });
}

int mc_infer_device_gather(mc_handle* h, const float* img, int B, const float* P2, const float* invP, float thres, int buf,
                           void* stream) {
    if (!h) return 1;
    return guarded(h, [&]() {
        auto& G = h->gather;
        MC_CHECK(G.connected && (buf == 0 || buf == 1), "mc_gather_connect first / buf");
        MC_CHECK(B == h->max_batch, "the gather slots are laid out for max_batch images per rank");
        cudaStream_t st = (cudaStream_t)stream;
        GatherParams gp;
        std::memset(&gp, 0, sizeof(gp));
        gp.n = G.world; gp.rank = G.rank;
        for (int r = 0; r < G.world; ++r) {
            gp.peer_slot[r] = G.peer[r] + ((size_t)buf * G.world + G.rank) * G.slot_bytes;
            gp.peer_data_flag[r] = g_data_flag(G.peer[r], G.data_bytes, buf) + G.rank;
        }
        gp.ready = g_ready_flag(G.block, G.data_bytes, buf);
        gp.done = g_done(G.block, G.data_bytes, buf);
        gp.gen = g_gen(G.block, G.data_bytes, buf);
        gp.off_box2d = G.off[0]; gp.off_box3d = G.off[1]; gp.off_labels = G.off[2]; gp.off_inds = G.off[3]; gp.off_valid = G.off[4];
        gp.error_flag = G.d_err;
        const unsigned gen = ++G.gen[buf];
        // this rank has consumed the previous generation of `buf` (stream order): let the peers overwrite it
        launch_gather_release(gp, G.d_peer_ready[buf], gen, st);
        char* slot = gp.peer_slot[G.rank];
        infer_device(h, img, B, P2, invP, G.topk, thres, (float*)(slot + G.off[0]), (float*)(slot + G.off[1]),
                     (long long*)(slot + G.off[2]), (long long*)(slot + G.off[3]), (unsigned char*)(slot + G.off[4]), st, &gp);
        h->launches += 1;
    });
}

int mc_gather_wait(mc_handle* h, int buf, void* stream) {
    if (!h) return 1;
    return guarded(h, [&]() {
        auto& G = h->gather;
        MC_CHECK(G.connected && (buf == 0 || buf == 1), "mc_gather_connect first / buf");
        launch_gather_wait(g_data_flag(G.block, G.data_bytes, buf), G.world, G.gen[buf], G.d_err, (cudaStream_t)stream);
    });
}

int mc_get_pred_ptrs(mc_handle* h, float* out_ptrs[MC_NUM_PRED]) {
    if (!h) return 1;
    for (int p = 0; p < kNumPred; ++p) out_ptrs[p] = h->pred_own[p];
    return 0;
}

int mc_copy_pred(mc_handle* h, int B, float* const dst[MC_NUM_PRED], void* stream) {
    if (!h) return 1;
    return guarded(h, [&]() {
        MC_CHECK(B >= 1 && B <= h->max_batch, "batch out of range");
        const size_t HW = (size_t)h->fh * h->fw;
        for (int p = 0; p < kNumPred; ++p)
            MC_CUDA(cudaMemcpyAsync(dst[p], h->pred_own[p], sizeof(float) * B * kPredCh[p] * HW, cudaMemcpyDeviceToDevice,
                                    (cudaStream_t)stream));
    });
}

int mc_num_stages(mc_handle* h) { return h ? num_stages(h) + 1 : 0; }

int mc_stage_info(mc_handle* h, int stage, char* name, int name_len, double* flops_per_image, double* bytes_per_image,
                  int* is_tensor_core) {
    if (!h) return 1;
    return guarded(h, [&]() {
        const int ns = num_stages(h);
        MC_CHECK(stage >= 0 && stage <= ns, "stage index");
        std::string nm = stage == ns ? std::string("decode") : stage_name(h, stage);
        double fl = 0, by = 0;
        int tc = 0;
        if (stage >= 1 && stage < ns && h->net->ops[stage - 1].type == OP_CONV) {
            const ConvLayer& L = h->net->convs[h->net->ops[stage - 1].conv];
            fl = L.flops_per_image; by = L.bytes_per_image; tc = L.dcn_off >= 0 ? 4 : (L.use_tc2 ? 2 : (L.use_tc3 ? 3 : (L.use_tc ? 1 : 0)));      // 4 = fused deformable convolution (dcn_tc.cu)
        }
        if (name && name_len > 0) std::snprintf(name, name_len, "%s", nm.c_str());
        if (flops_per_image) *flops_per_image = fl;
        if (bytes_per_image) *bytes_per_image = by;
        if (is_tensor_core) *is_tensor_core = tc;
    });
}

int mc_profile_stages(mc_handle* h, const float* img, int B, const float* P2, const float* invP, int iters, float* ms_out,
                      void* stream) {
    if (!h) return 1;
    return guarded(h, [&]() {
        cudaStream_t st = (cudaStream_t)stream;
        const int ns = num_stages(h) + 1;
        ensure_staging(h, 30);
        struct Hook : StageHook {
            std::vector<cudaEvent_t> e0, e1;
            void before(int s, cudaStream_t st) override { cudaEventRecord(e0[s], st); }
            void after(int s, cudaStream_t st) override { cudaEventRecord(e1[s], st); }
        } hook;
        hook.e0.resize(ns); hook.e1.resize(ns);
        for (int i = 0; i < ns; ++i) { MC_CUDA(cudaEventCreate(&hook.e0[i])); MC_CUDA(cudaEventCreate(&hook.e1[i])); }
        std::vector<double> acc(ns, 0.0);
        for (int it = 0; it < iters + 1; ++it) {
            run_forward(h, img, B, h->pred_own, st, &hook);
            hook.before(ns - 1, st);
            run_decode(h, h->pred_own, B, P2, invP, h->H, h->W, 30, 0.4f, h->d_box2d, h->d_box3d, h->d_labels, h->d_inds,
                       h->d_valid, st);
            hook.after(ns - 1, st);
            MC_CUDA(cudaStreamSynchronize(st));
            if (it == 0) continue;              // first pass = warm-up
            for (int i = 0; i < ns; ++i) {
                float ms = 0.f;
                MC_CUDA(cudaEventElapsedTime(&ms, hook.e0[i], hook.e1[i]));
                acc[i] += ms;
            }
        }
        for (int i = 0; i < ns; ++i) {
            ms_out[i] = (float)(acc[i] / iters);
            cudaEventDestroy(hook.e0[i]); cudaEventDestroy(hook.e1[i]);
        }
    });
}

// fp32-accurate tensor-core mode: fit the per-tensor power-of-two scales of the fp16 planes to a sample batch.  Every writer
// of a DT_SPLIT tensor keeps the running maximum of |stored value|; a pass whose maxima all stayed below the fp16 limit
// gives the true maxima (stored / 2^e), from which each tensor gets the exponent that puts its maximum into [2^11, 2^12)
// -- 16-32x headroom before saturation, full hi + lo precision for everything within 2^-14 of the maximum.  Tensors that
// meet in one convolution's K dimension (concatenated sources) and a max-pool's input / output share the smallest exponent
// of their group.  A saturated pass lowers the offending exponents by 2^10 and repeats.
int mc_calibrate_scales(mc_handle* h, const float* img, int B, void* stream) {
    if (!h) return 1;
    return guarded(h, [&]() {
        MC_CHECK(h->finalized, "mc_finalize_params has not been called");
        MC_CHECK(h->dt == DT_SPLIT, "mc_calibrate_scales: MC_PREC_FP32_TC engines only");
        Net& n = *h->net;
        cudaStream_t st = (cudaStream_t)stream;
        const size_t nt = n.tensors.size();
        // exponent groups
        std::vector<int> parent(nt);
        for (size_t i = 0; i < nt; ++i) parent[i] = (int)i;
        std::function<int(int)> find = [&](int a) { return parent[a] == a ? a : parent[a] = find(parent[a]); };
        auto unite = [&](int a, int b) { parent[find(a)] = find(b); };
        for (const auto& L : n.convs)
            for (size_t s = 1; s < L.src.size(); ++s) unite(L.src[0], L.src[s]);
        for (const auto& op : n.ops)
            if (op.type == OP_POOL) unite(op.src, op.dst);
        std::vector<int> e = n.act_exp;
        for (int pass = 0; pass < 12; ++pass) {
            n.read_act_amax(true);
            run_forward(h, img, B, h->pred_own, st);
            MC_CUDA(cudaStreamSynchronize(st));
            const std::vector<float> amax = n.read_act_amax(false);
            bool saturated = false;
            std::vector<int> want(nt, 1 << 20);
            for (size_t i = 0; i < nt; ++i) {
                if (n.tensors[i].dt != DT_SPLIT) continue;
                const float m = amax[i];
                if (!(m < 65000.f)) { saturated = true; want[i] = e[i] - 10; continue; }     // also catches NaN
                if (m <= 0.f) { want[i] = e[i]; continue; }                                    // never written / all zero
                int ex = 0;
                std::frexp((double)m * std::ldexp(1.0, -e[i]), &ex);                           // true max = f * 2^ex, f in [0.5, 1)
                want[i] = std::max(-100, std::min(100, 12 - ex));
            }
            std::vector<int> gmin(nt, 1 << 20);
            for (size_t i = 0; i < nt; ++i) if (n.tensors[i].dt == DT_SPLIT) gmin[find((int)i)] = std::min(gmin[find((int)i)], want[i]);
            std::vector<int> ne = e;
            for (size_t i = 0; i < nt; ++i) if (n.tensors[i].dt == DT_SPLIT) ne[i] = gmin[find((int)i)];
            const bool changed = ne != e;
            e = ne;
            n.set_act_exponents(e);
            if (!saturated && !changed) break;
            if (!saturated) { /* one more pass confirms that nothing saturates at the new scales */ }
            MC_CHECK(pass < 11, "mc_calibrate_scales: activations do not settle inside the fp16 range");
        }
        n.read_act_amax(true);
        h->launches = n.launches_last_run;
    });
}

// Range use of the fp16 planes since the last call (or calibration): the largest |stored value| of any tensor as a fraction of
// the fp16 limit, and how many tensors reached it (their outputs are clamped, i.e. WRONG: recalibrate and run again).
int mc_scale_status(mc_handle* h, float* max_fraction, int* n_saturated) {
    if (!h) return 1;
    return guarded(h, [&]() {
        MC_CHECK(h->dt == DT_SPLIT && h->finalized, "mc_scale_status: finalized MC_PREC_FP32_TC engines only");
        const std::vector<float> amax = h->net->read_act_amax(true);
        float mf = 0.f;
        int ns = 0;
        for (size_t i = 0; i < amax.size(); ++i) {
            if (h->net->tensors[i].dt != DT_SPLIT) continue;
            if (!(amax[i] < 65000.f)) ++ns;
            mf = std::max(mf, amax[i] / 65504.f);
        }
        if (max_fraction) *max_fraction = mf;
        if (n_saturated) *n_saturated = ns;
    });
}

int mc_set_option(mc_handle* h, const char* name, int value) {
    if (!h) return 1;
    return guarded(h, [&]() {
        const std::string n(name ? name : "");
        if (n == "conv_impl") {
            MC_CHECK(!h->finalized, "conv_impl must be set before mc_finalize_params");
            MC_CHECK(value == MC_CONV_AUTO || value == MC_CONV_SIMT, "conv_impl value");
            MC_CHECK(value == MC_CONV_AUTO || h->dt != DT_SPLIT, "MC_PREC_FP32_TC has tensor-core convolutions only (use MC_PREC_FP32 for the FFMA kernels)");
            h->net->conv_impl = value;
        } else if (n == "use_graph") {
            h->use_graph = value != 0;
        } else if (n == "train_debug") {       // bf16 training: also keep the fp32 gradient of the head stems (mc_debug_train_dump)
            h->train_debug = value != 0;
        } else if (n == "head_backward") {     // bf16 training: 1 = the restructured heads backward (default), 0 = the fp32 twin's kernels
            h->head_backward_fast = value != 0;
        } else {
            throw Error("unknown option: " + n);
        }
    });
}

size_t mc_workspace_bytes(const mc_handle* h) { return h ? h->net->arena.total() : 0; }
int mc_num_kernel_launches(const mc_handle* h) { return h ? h->launches : 0; }
double mc_flops_per_image(const mc_handle* h) { return h ? h->flops : 0.0; }
double mc_bytes_per_image(const mc_handle* h) { return h ? h->bytes : 0.0; }

const char* mc_last_error(const mc_handle* h) {
    if (h) return h->err.c_str();
    std::lock_guard<std::mutex> g(g_mutex);
    return g_create_error.c_str();
}

void mc_destroy(mc_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    for (auto& kv : h->graphs) cudaGraphExecDestroy(kv.second);
    for (auto& S : h->slots) {
        if (S.ev_in) { cudaEventDestroy(S.ev_in); cudaEventDestroy(S.ev_done); cudaEventDestroy(S.ev_out); }
    }
    if (h->st_h2d) { cudaStreamDestroy(h->st_h2d); cudaStreamDestroy(h->st_comp); cudaStreamDestroy(h->st_d2h); }
    if (h->gather.block) {
        cudaDeviceSynchronize();
        for (int r = 0; r < h->gather.world; ++r)
            if (r != h->gather.rank && h->gather.peer[r]) cudaIpcCloseMemHandle(h->gather.peer[r]);
        for (int b = 0; b < 2; ++b) cudaFree(h->gather.d_peer_ready[b]);
        cudaFree(h->gather.d_err);
        cudaFree(h->gather.block);
    }
    delete h;
}

int mc_debug_tensor_shape(mc_handle* h, const char* name, int* C, int* H, int* W) {
    if (!h) return 1;
    return guarded(h, [&]() {
        auto it = h->net->aliases_.find(name ? name : "");
        if (it == h->net->aliases_.end()) throw Error(std::string("no such tensor: ") + (name ? name : ""));
        const TensorInfo& t = h->net->tensors[it->second];
        *C = t.C; *H = t.H; *W = t.W;
    });
}

int mc_debug_tensor(mc_handle* h, const char* name, int B, float* out_nchw, void* stream) {
    if (!h) return 1;
    return guarded(h, [&]() {
        auto it = h->net->aliases_.find(name ? name : "");
        if (it == h->net->aliases_.end()) throw Error(std::string("no such tensor: ") + (name ? name : ""));
        const TensorInfo& t = h->net->tensors[it->second];
        MC_CHECK(t.Wp == t.W, "padded tensors cannot be dumped");
        launch_unpack_nchw(t.ptr, t.dt, out_nchw, B, t.C, t.H, t.W, (cudaStream_t)stream, h->net->split_info(it->second));
    });
}

int mc_conv2d(int device, int precision_mode, int conv_impl, const float* x, int B, int Cin, int H, int W, const float* w,
              int Cout, int k, int stride, int pad, const float* scale, const float* shift, const float* residual, int relu,
              int split, float* y, void* stream, char* err, int err_len) {
    try {
        MC_CUDA(cudaSetDevice(device));
        MC_CHECK(split >= 1 && split <= kMaxSrc && Cin % split == 0, "split");
        cudaStream_t st = (cudaStream_t)stream;
        const DType dt = precision_mode == MC_PREC_FP32 ? DT_F32 : (precision_mode == MC_PREC_FP32_TC ? DT_SPLIT : DT_BF16);
        Net net(device, B, dt, conv_impl);
        tc_kernels_init();
        tc2_kernels_init();
        tc3_kernels_init();
        const int Cs = Cin / split;
        const bool stem_like = (Cin == 3 && k == 7 && dt != DT_F32);
        const int Cst = stem_like ? 8 : ((Cs % 4 == 0) ? Cs : (Cs + 3) / 4 * 4);   // storage channels (Cin=3 -> 4 / 8)
        const int Wp = stem_like ? W + 8 : W, xoff = stem_like ? 4 : 0;
        MC_CHECK(split == 1 || Cst == Cs, "split needs channel groups that are multiples of 4");
        std::vector<int> src;
        for (int s = 0; s < split; ++s) src.push_back(net.add_tensor("x" + std::to_string(s), Cst, H, W, Wp, xoff));
        if (stem_like && dt == DT_SPLIT) net.tensors[src[0]].hl_interleaved = true;
        const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
        int res = -1;
        if (residual) res = net.add_tensor("res", Cout, Ho, Wo);
        net.add_conv("conv", src, Cout, k, stride, pad, {}, res, relu != 0, Cin);
        net.allocate();
        std::vector<float> hw((size_t)Cout * Cin * k * k), hs(Cout), hb(Cout);
        MC_CUDA(cudaMemcpy(hw.data(), w, sizeof(float) * hw.size(), cudaMemcpyDefault));
        MC_CUDA(cudaMemcpy(hs.data(), scale, sizeof(float) * Cout, cudaMemcpyDefault));
        MC_CUDA(cudaMemcpy(hb.data(), shift, sizeof(float) * Cout, cudaMemcpyDefault));
        net.pack_conv(net.convs[0], hw, hs, hb);
        for (int s = 0; s < split; ++s)
            for (int b = 0; b < B; ++b) {
                // DT_SPLIT: image b of the hi plane; the lo plane follows split_info().plane elements later
                char* dstp = (char*)net.tensors[src[s]].ptr + (size_t)b * H * Wp * Cst * (dt == DT_SPLIT ? 2 : dtype_size(dt));
                if (Cst == Cs)
                    launch_pack_nhwc(x + ((size_t)b * Cin + (size_t)s * Cs) * H * W, dstp, dt, 1, Cs, H, W, st, net.split_info(src[s]));
                else
                    launch_pack_input(x + (size_t)b * Cin * H * W, dstp, dt, 1, Cs, H, W, Cst, Wp, xoff, st, net.split_info(src[s]));
            }
        if (residual) launch_pack_nhwc(residual, net.tensors[res].ptr, dt, B, Cout, Ho, Wo, st, net.split_info(res));
        net.run_ops(B, st);
        const int dstT = net.convs[0].dst;
        launch_unpack_nchw(net.tensors[dstT].ptr, net.tensors[dstT].dt, y, B, Cout, Ho, Wo, st, net.split_info(dstT));
        MC_CUDA(cudaStreamSynchronize(st));
        return 0;
    } catch (const std::exception& e) {
        if (err && err_len > 0) std::snprintf(err, err_len, "%s", e.what());
        return 1;
    }
}

int mc_deform_conv2d(int device, int precision_mode, const float* x, int B, int Cin, int H, int W, const float* offset, const float* mask,
                     const float* w, const float* bias, int Cout, int split, float* y, void* stream, char* err, int err_len) {
    try {
        MC_CUDA(cudaSetDevice(device));
        MC_CHECK(split >= 1 && split <= 2 && Cin % (8 * split) == 0, "split: one or two equal channel groups, multiples of 8");
        MC_CHECK(x && offset && mask && w && y, "null argument");
        cudaStream_t st = (cudaStream_t)stream;
        const DType dt = precision_mode == MC_PREC_FP32 ? DT_F32 : (precision_mode == MC_PREC_FP32_TC ? DT_SPLIT : DT_BF16);
        Net net(device, B, dt, MC_CONV_AUTO);
        tc_kernels_init();
        tc2_kernels_init();
        tc3_kernels_init();
        const int Cs = Cin / split;
        std::vector<int> src;
        for (int s = 0; s < split; ++s) src.push_back(net.add_tensor("x" + std::to_string(s), Cs, H, W));
        const int off = net.add_tensor("offset_mask", 32, H, W);
        // `mask` is the modulation itself, as the operator takes it.  Tensor-core storage with whole 64-channel K-blocks: the fused
        // kernel (csrc/dcn_tc.cu) unless MC_DCN_FUSE=0; otherwise columns + 1x1 layer
        const char* fe = std::getenv("MC_DCN_FUSE");
        const bool fused = (dt == DT_BF16 || dt == DT_SPLIT) && Cs % 64 == 0 && Cout <= 256 && Cout % 16 == 0 && !(fe && fe[0] == '0');
        if (fused) {
            dcn_tc_init();
            net.add_dcn_conv("dcn", src, off, Cout, {}, false, false);
        } else {
            const int col = net.add_dcn_columns("columns", src, off, false);
            net.add_conv("dcn", {col}, Cout, 1, 1, 0, {}, -1, false);
        }
        net.allocate();
        // unfused: (Cout, Cin, 3, 3) -> (Cout, 9 Cin), K index = tap * Cin + c
        const size_t HWs = (size_t)H * W;
        std::vector<float> hw((size_t)Cout * Cin * 9), hk(hw.size()), hs(Cout, 1.f), hb(Cout, 0.f);
        MC_CUDA(cudaMemcpy(hw.data(), w, sizeof(float) * hw.size(), cudaMemcpyDefault));
        if (bias) MC_CUDA(cudaMemcpy(hb.data(), bias, sizeof(float) * Cout, cudaMemcpyDefault));
        for (int o = 0; o < Cout; ++o)
            for (int c = 0; c < Cin; ++c)
                for (int t = 0; t < 9; ++t) hk[((size_t)o * 9 + t) * Cin + c] = hw[((size_t)o * Cin + c) * 9 + t];
        net.pack_conv(net.convs[0], fused ? hw : hk, hs, hb);
        // offsets (B, 18, H, W) and mask (B, 9, H, W) -> one 32-channel NCHW image per batch entry -> NHWC
        float* om = (float*)net.arena.alloc(sizeof(float) * (size_t)B * 32 * HWs);
        for (int b = 0; b < B; ++b) {
            MC_CUDA(cudaMemcpyAsync(om + (size_t)b * 32 * HWs, offset + (size_t)b * 18 * HWs, sizeof(float) * 18 * HWs, cudaMemcpyDefault, st));
            MC_CUDA(cudaMemcpyAsync(om + ((size_t)b * 32 + 18) * HWs, mask + (size_t)b * 9 * HWs, sizeof(float) * 9 * HWs, cudaMemcpyDefault, st));
        }
        launch_pack_nhwc(om, net.tensors[off].ptr, dt, B, 32, H, W, st, net.split_info(off));
        for (int s = 0; s < split; ++s)
            for (int b = 0; b < B; ++b) {
                char* dstp = (char*)net.tensors[src[s]].ptr + (size_t)b * HWs * Cs * (dt == DT_SPLIT ? 2 : dtype_size(dt));
                launch_pack_nhwc(x + ((size_t)b * Cin + (size_t)s * Cs) * HWs, dstp, dt, 1, Cs, H, W, st, net.split_info(src[s]));
            }
        net.run_ops(B, st);
        const int dstT = net.convs[0].dst;
        launch_unpack_nchw(net.tensors[dstT].ptr, net.tensors[dstT].dt, y, B, Cout, H, W, st, net.split_info(dstT));
        MC_CUDA(cudaStreamSynchronize(st));
        return 0;
    } catch (const std::exception& e) {
        if (err && err_len > 0) std::snprintf(err, err_len, "%s", e.what());
        return 1;
    }
}

}  // extern "C"
