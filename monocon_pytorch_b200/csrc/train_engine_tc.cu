// bf16 tensor-core training step, engine side (BASELINE.json configs[2] "batch=32 training step bf16" and
// configs[4]; reference: engine/monocon_engine.py:80-102 under torch.autocast-like arithmetic): mc_finalize_params(h, 1 | 2) on
// an MC_PREC_BF16 handle.
//
//   forward   every convolution runs on the inference engine's tcgen05 kernels (conv_tc / conv_tc2 / conv_tc3), writing its RAW
//             output as bf16 into a buffer of its own (the nine head stems, bias only, into their tensor); train-mode BatchNorm = bn_stats_bf16 ->
//             bn_finalize -> bn_apply_bf16 (train_tc.cu); heads: the fp32 engine's kernels on bf16 stems.
//   backward  the engine's op list walked downwards.  Per convolution: BatchNorm backward (train_tc.cu) writes the gradient of
//             the raw output as bf16 -- dense, or ZERO-INSERTED at input resolution when the convolution has stride 2 --, then
//               wgrad  = wgrad_tc_kernel (wgrad_tc.cu): dW[tap][ci][co] += sum_p dy[p][co] x[p + tap][ci], fp32 into the master layout;
//               dgrad  = the FORWARD kernels on the spatially flipped, in/out-transposed weights, one launch per source tensor of
//                        the (concat-free) convolution, reading the gradient above and writing / accumulating the source's bf16
//                        gradient tensor through the residual input of the epilogue (dst += conv(...)).
//             A stride-2 convolution's dgrad and wgrad on the zero-inserted gradient ARE stride-1 problems (4x the MMAs of five
//             small layers), so one tensor-core path covers all 50 convolutions; the 7x7 stem has its own wgrad view and no dgrad.
//   weights   the optimiser owns fp32 master weights in the [tap][cin][cout] layout; every forward starts by re-deriving all
//             bf16 plan buffers (forward layouts and dgrad layouts) from them with one gather kernel per plan (repack_bf16).
// Which contribution to a gradient tensor comes first in backward order is known statically, so the first one overwrites and
// the others accumulate: no gradient tensor is ever zeroed.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "handle.h"
#include "train_backward.h"
#include "train_tc.h"

namespace mc {

struct TrainTc {
    struct ConvT {
        void* raw = nullptr;              // bf16 [max_batch][Ho][Wo][cout] raw convolution output (BatchNorm layers); the stems tensor itself for the head stems
        int draw = -1;                    // bnet tensor: gradient of the raw output (bf16; zero-inserted for stride 2)
        std::shared_ptr<WgradPlan> wg;
        std::vector<int> dgrad;           // bnet convolutions, one per source that needs a gradient
        bool res_acc = false;             // BatchNorm backward: the residual's gradient accumulates (else: first contribution)
        std::vector<float> host_w;        // OIHW weights as handed to pack_conv (kept until the dgrad plans are built)
    };
    struct Repack { const float* master; int* idx; void* out; long long n; };
    std::unique_ptr<Net> bnet;            // gradient tensors + dgrad convolutions (own arena)
    std::vector<int> g;                   // forward tensor -> bnet tensor of its gradient (-1: the image, the fp32 stems)
    std::vector<ConvT> conv;              // indexed like net->convs
    std::vector<char> op_acc;             // per forward op (POOL / UP): its backward accumulates into the source's gradient
    std::vector<Repack> repacks;
    RepackJob* jobs_dev = nullptr;        // all of `repacks` as one launch
    long long repack_total = 0;
    int bwd_launches = 0;
    bool built = false;
    // the weight-gradient kernels run on a stream of their own: nothing in the backward walk depends on them, they are tensor-bound
    // and leave room on the SMs (shared memory, threads) for the bandwidth-bound BatchNorm kernels of the next stage
    cudaStream_t st_w = nullptr;
    cudaEvent_t ev_ready = nullptr, ev_done = nullptr;
    bool side_stream = true;
    double* sums_all = nullptr;           // [forward | backward][conv] per-channel sums, 2 x 1024 doubles each: ONE memset per pass
    float* stems_f32 = nullptr;           // fp32 copy of the stems for the fp32 twin's heads backward (option head_backward = 0)
    ~TrainTc() {
        if (st_w) cudaStreamDestroy(st_w);
        if (ev_ready) cudaEventDestroy(ev_ready);
        if (ev_done) cudaEventDestroy(ev_done);
    }
};

std::shared_ptr<TrainTc> traintc_create() {
    wgrad_tc_init();
    auto t = std::make_shared<TrainTc>();
    if (const char* e = std::getenv("MC_WGRAD_STREAM")) t->side_stream = e[0] != '0';      // A/B knob
    return t;
}

void traintc_before_pack(mc_handle* h, int conv_index, ConvLayer& L, const std::vector<float>& w_oihw) {
    TrainTc& T = *h->train_tc;
    Net& n = *h->net;
    if ((int)T.conv.size() <= conv_index) T.conv.resize(n.convs.size());
    TrainTc::ConvT& c = T.conv[conv_index];
    const TensorInfo& d = n.tensors[L.dst];
    const size_t elems = (size_t)n.max_batch * d.H * d.W * L.cout;
    const bool has_bn = h->bn_train[conv_index].C > 0;
    L.keep_widx = true;
    if (has_bn) {
        c.raw = n.arena.alloc(elems * 2);                    // same order on every finalize (arena replay of mc_refresh_params)
        L.dst_override = c.raw;
    } else {
        // the head stems (bias only): the plan's own bf16 output tensor is what AttnBN statistics, the 1x1 heads and the heads
        // backward read -- as autocast stores a convolution's output
        c.raw = d.ptr;
    }
    if (!T.built) c.host_w = w_oihw;
}

namespace {

int* upload_idx(DeviceArena& a, const std::vector<int>& v) {
    int* d = (int*)a.alloc(sizeof(int) * v.size());
    MC_CUDA(cudaMemcpy(d, v.data(), sizeof(int) * v.size(), cudaMemcpyHostToDevice));
    return d;
}

}  // namespace

void traintc_setup(mc_handle* h) {
    TrainTc& T = *h->train_tc;
    Net& n = *h->net;
    if (T.built) return;                  // mc_refresh_params: same buffers, the next forward repacks from the new master weights
    T.bnet.reset(new Net(h->device, h->max_batch, DT_BF16, MC_CONV_AUTO));
    Net& bn = *T.bnet;
    // forward plans: packed element -> master index.  widx indexes OIHW with the LOGICAL cin; master is [tap][cin_store][cout].
    for (size_t i = 0; i < n.convs.size(); ++i) {
        ConvLayer& L = n.convs[i];
        MC_CHECK(L.use_tc && L.w_packed && !L.widx.empty() && L.w_simt, "bf16 training: no tensor-core plan for " + L.name);
        const int kk = L.k * L.k;
        std::vector<int> idx(L.widx.size());
        for (size_t e = 0; e < idx.size(); ++e) {
            const int id = L.widx[e];
            if (id < 0) { idx[e] = -1; continue; }
            const int t = id % kk, c = (id / kk) % L.cin, o = id / (kk * L.cin);
            idx[e] = (t * L.cin_store + c) * L.cout + o;
        }
        T.repacks.push_back(TrainTc::Repack{L.w_simt, upload_idx(bn.arena, idx), L.w_packed, (long long)idx.size()});
        L.widx.clear(); L.widx.shrink_to_fit();
    }
    auto finish = [&]() {
        T.sums_all = (double*)bn.arena.alloc(sizeof(double) * 2 * n.convs.size() * 2048);
        std::vector<RepackJob> jobs;
        long long start = 0;
        for (const auto& r : T.repacks) { jobs.push_back(RepackJob{r.master, r.idx, r.out, start, r.n}); start += repack_blocks(r.n); }
        T.repack_total = start;
        T.jobs_dev = (RepackJob*)bn.arena.alloc(sizeof(RepackJob) * jobs.size());
        MC_CUDA(cudaMemcpy(T.jobs_dev, jobs.data(), sizeof(RepackJob) * jobs.size(), cudaMemcpyHostToDevice));
        T.built = true;
    };
    if (!h->backward) { finish(); return; }

    // gradient tensors
    T.g.assign(n.tensors.size(), -1);
    for (size_t t = 0; t < n.tensors.size(); ++t) {
        const TensorInfo& ti = n.tensors[t];
        if ((int)t == h->t_input || (int)t == h->t_stems || (int)t == h->t_headz) continue;
        MC_CHECK(ti.Wp == ti.W && ti.xoff == 0, "bf16 training: padded activation tensor " + ti.name);
        T.g[t] = bn.add_tensor("g." + ti.name, ti.C, ti.H, ti.W);
    }
    for (size_t i = 0; i < n.convs.size(); ++i) {
        const ConvLayer& L = n.convs[i];
        const TensorInfo& d = n.tensors[L.dst];
        const TensorInfo& s0 = n.tensors[L.src[0]];
        MC_CHECK(L.stride == 1 || (L.stride == 2 && L.k == 3 && L.pad == 1 && s0.H == 2 * d.H && s0.W == 2 * d.W), "bf16 training: stride of " + L.name);
        T.conv[i].draw = L.stride == 1 ? bn.add_tensor("draw." + L.name, L.cout, d.H, d.W) : bn.add_tensor("draw." + L.name, L.cout, s0.H, s0.W);
    }
    // backward order: who writes a gradient tensor first
    std::vector<char> written(n.tensors.size(), 0);
    T.op_acc.assign(n.ops.size(), 0);
    struct Pending { int conv, src_index, src_tensor; bool acc; };
    std::vector<Pending> pend;
    for (int i = (int)n.ops.size() - 1; i >= 0; --i) {
        const Op& op = n.ops[i];
        if (op.type == OP_HEADS) {
            written[h->t_stems] = 1;
        } else if (op.type == OP_CONV) {
            const ConvLayer& L = n.convs[op.conv];
            MC_CHECK(written[L.dst], "bf16 training: the output of " + L.name + " has no consumer");
            if (L.residual >= 0) {
                MC_CHECK(h->bn_train[op.conv].C > 0, "bf16 training: residual without BatchNorm");
                T.conv[op.conv].res_acc = written[L.residual] != 0;
                written[L.residual] = 1;
            }
            for (int s = 0; s < (int)L.src.size(); ++s) {
                const int t = L.src[s];
                if (t == h->t_input) continue;
                pend.push_back(Pending{op.conv, s, t, written[t] != 0});
                written[t] = 1;
            }
        } else {
            T.op_acc[i] = written[op.src];
            written[op.src] = 1;
        }
    }
    bn.allocate();
    // dgrad convolutions: y' = conv(draw, w'), w'[ci][co][a][b] = w[co][cb + ci][k-1-a][k-1-b], into / onto the source's gradient
    for (const Pending& pd : pend) {
        const ConvLayer& L = n.convs[pd.conv];
        TrainTc::ConvT& c = T.conv[pd.conv];
        int cb = 0;
        for (int s = 0; s < pd.src_index; ++s) cb += n.tensors[L.src[s]].C;
        const int Cs = n.tensors[pd.src_tensor].C, k = L.k, kk = k * k;
        const int gi = T.g[pd.src_tensor];
        const int ci = bn.add_conv_to("dgrad." + L.name + "." + std::to_string(pd.src_index), {c.draw}, gi, k, 1, L.pad, pd.acc ? gi : -1, false);
        ConvLayer& D = bn.convs[ci];
        D.keep_widx = true;
        std::vector<float> w((size_t)Cs * L.cout * kk);
        for (int o = 0; o < Cs; ++o)
            for (int i2 = 0; i2 < L.cout; ++i2)
                for (int t = 0; t < kk; ++t)
                    w[((size_t)o * L.cout + i2) * kk + t] = c.host_w[((size_t)i2 * L.cin + cb + o) * kk + (kk - 1 - t)];
        const std::vector<float> one(Cs, 1.f), zero(Cs, 0.f);
        bn.pack_conv(D, w, one, zero);
        MC_CHECK(D.use_tc && D.w_packed && !D.widx.empty(), "bf16 training: no tensor-core kernel for the dgrad of " + L.name);
        std::vector<int> idx(D.widx.size());
        for (size_t e = 0; e < idx.size(); ++e) {
            const int id = D.widx[e];
            if (id < 0) { idx[e] = -1; continue; }
            const int t = id % kk, i2 = (id / kk) % L.cout, o = id / (kk * L.cout);       // w'[o][i2][t]
            idx[e] = ((kk - 1 - t) * L.cin_store + cb + o) * L.cout + i2;
        }
        T.repacks.push_back(TrainTc::Repack{L.w_simt, upload_idx(bn.arena, idx), D.w_packed, (long long)idx.size()});
        D.widx.clear(); D.widx.shrink_to_fit();
        c.dgrad.push_back(ci);
    }
    // wgrad plans
    for (size_t i = 0; i < n.convs.size(); ++i) {
        const ConvLayer& L = n.convs[i];
        TrainTc::ConvT& c = T.conv[i];
        const TensorInfo& dr = bn.tensors[c.draw];
        WgradDesc d;
        d.dy = dr.ptr; d.nsrc = (int)L.src.size(); d.H = dr.H; d.W = dr.W; d.Cout = L.cout; d.k = L.k; d.dw = h->bwd_conv[i].dw;
        for (int s = 0; s < d.nsrc; ++s) {
            const TensorInfo& t = n.tensors[L.src[s]];
            d.src[s] = WgradSrc{t.ptr, t.C, t.Wp == t.W ? 0 : t.Wp, t.xoff};
        }
        MC_CHECK(wgrad_tc_supported(d), "bf16 training: no tensor-core weight-gradient kernel for " + L.name);
        c.wg = wgrad_tc_prepare(d, h->max_batch, bn.arena, L.name);
        c.host_w.clear(); c.host_w.shrink_to_fit();
    }
    finish();
}

void traintc_debug(mc_handle* h, int kind, int index, const void** ptr, DType* dt, int* C, int* H, int* W) {
    TrainTc& T = *h->train_tc;
    Net& n = *h->net;
    MC_CHECK(T.built && h->backward, "bf16 training: debug dump needs mc_finalize_params(h, 2)");
    if (kind == 1) {
        MC_CHECK(index >= 0 && index < (int)T.g.size() && T.g[index] >= 0, "bf16 training: tensor without gradient");
        const TensorInfo& t = T.bnet->tensors[T.g[index]];
        *ptr = t.ptr; *dt = DT_BF16; *C = t.C; *H = t.H; *W = t.W;
        return;
    }
    MC_CHECK(index >= 0 && index < (int)T.conv.size() && (kind == 2 || kind == 3), "bf16 training: debug kind / convolution index");
    if (kind == 2) {
        const TensorInfo& d = n.tensors[n.convs[index].dst];
        *ptr = T.conv[index].raw; *dt = DT_BF16; *C = d.C; *H = d.H; *W = d.W;
    } else {
        const TensorInfo& t = T.bnet->tensors[T.conv[index].draw];
        *ptr = t.ptr; *dt = DT_BF16; *C = t.C; *H = t.H; *W = t.W;
    }
}

void traintc_forward(mc_handle* h, const float* img, int B, float* const pred_out[kNumPred], cudaStream_t st) {
    TrainTc& T = *h->train_tc;
    Net& n = *h->net;
    n.launches_last_run = 0;
    launch_repack_all_bf16(T.jobs_dev, (int)T.repacks.size(), T.repack_total, st);
    n.launches_last_run++;
    MC_CUDA(cudaMemsetAsync(T.sums_all, 0, sizeof(double) * n.convs.size() * 2048, st));      // the forward half
    const TensorInfo& in = n.tensors[h->t_input];
    launch_pack_input(img, in.ptr, n.dt, B, 3, h->H, h->W, in.C, in.Wp, in.xoff, st);
    n.launches_last_run++;
    for (int i = 0; i < (int)n.ops.size(); ++i) {
        const Op& op = n.ops[i];
        if (op.type == OP_CONV) {
            const ConvLayer& L = n.convs[op.conv];
            const auto& bt = h->bn_train[op.conv];
            const TensorInfo& d = n.tensors[L.dst];
            n.run_conv(op.conv, B, st);
            n.launches_last_run++;
            if (bt.C > 0) {
                const long long P = (long long)B * d.H * d.W;
                const void* raw = T.conv[op.conv].raw;
                double* sums = T.sums_all + (size_t)op.conv * 2048;
                launch_bn_stats_bf16(raw, P, L.cout, sums, st, true);
                float *mean = nullptr, *inv = nullptr;
                if (h->backward) { mean = h->bwd_conv[op.conv].mean; inv = h->bwd_conv[op.conv].inv; }
                launch_bn_finalize_apply_bf16(raw, d.ptr, L.residual >= 0 ? n.tensors[L.residual].ptr : nullptr, P, L.cout, sums, bt.eps, 0.1f, bt.gamma,
                                              bt.beta, bt.rmean, bt.rvar, bt.scale, bt.shift, mean, inv, L.relu, st);
                n.launches_last_run += 2;
            }
        } else if (op.type == OP_HEADS) {
            const int HW = h->fh * h->fw;
            int stems_conv = -1;
            for (size_t c = 0; c < n.convs.size(); ++c)
                if (n.convs[c].dst == h->t_stems) stems_conv = (int)c;
            const void* stems = T.conv[stems_conv].raw;          // bf16 [B][HW][576], bias included
            launch_attn_stats(stems, DT_BF16, h->hp.sums, B, HW, st);
            launch_attn_mix_train(h->hp.sums, B, HW, h->hp.att_w, h->att_gamma, h->att_beta, h->att_rmean, h->att_rvar, h->hp.bank_w,
                                  h->hp.bank_b, h->hbn_rmean, h->hbn_rvar, h->hp.coefA, h->hp.coefB, st);
            HeadApplyParams ap;
            ap.stems = stems; ap.coefA = h->hp.coefA; ap.coefB = h->hp.coefB; ap.w = h->hp.w; ap.bias = h->hp.bias;
            for (int p = 0; p < kNumPred; ++p) ap.out[p] = pred_out[p];
            ap.B = B; ap.HW = HW;
            launch_head_apply(ap, DT_BF16, st);
            n.launches_last_run += 3;
        } else {
            n.run_ops(B, st, i, i + 1);
        }
    }
}

void traintc_backward(mc_handle* h, int B, int op_first, int op_last, bool zero, cudaStream_t st) {
    TrainTc& T = *h->train_tc;
    Net& n = *h->net;
    Net& bn = *T.bnet;
    MC_CHECK(T.built && h->backward, "bf16 training: backward needs mc_finalize_params(h, 2)");
    if (zero) {
        MC_CUDA(cudaMemsetAsync(h->bwd_dw_pool, 0, sizeof(float) * h->bwd_dw_pool_floats, st));
        for (size_t i = 0; i < n.ops.size(); ++i)
            if (n.ops[i].type == OP_UP) MC_CUDA(cudaMemsetAsync(h->bwd_up_dw[i], 0, sizeof(float) * (size_t)n.tensors[n.ops[i].src].C * 16, st));
        MC_CUDA(cudaMemsetAsync(T.sums_all + n.convs.size() * 2048, 0, sizeof(double) * n.convs.size() * 2048, st));     // the backward half
    }
    auto grad = [&](int t) -> void* {
        MC_CHECK(T.g[t] >= 0, "bf16 training: tensor without gradient");
        return bn.tensors[T.g[t]].ptr;
    };
    int cnt = 0;                                  // kernels launched by this segment (mc_num_kernel_launches: forward + backward)
    if (T.side_stream && !T.st_w) {
        // lowest priority: when a dgrad convolution of the main chain and a weight-gradient kernel both wait for SMs, the chain goes first
        int prio_lo = 0, prio_hi = 0;
        MC_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        MC_CUDA(cudaStreamCreateWithPriority(&T.st_w, cudaStreamNonBlocking, prio_lo));
        MC_CUDA(cudaEventCreateWithFlags(&T.ev_ready, cudaEventDisableTiming));
        MC_CUDA(cudaEventCreateWithFlags(&T.ev_done, cudaEventDisableTiming));
    }
    bool side_used = false;
    for (int i = op_last - 1; i >= op_first; --i) {
        const Op& op = n.ops[i];
        if (op.type == OP_HEADS) {
            int stems_conv = -1;
            for (size_t c = 0; c < n.convs.size(); ++c)
                if (n.convs[c].dst == h->t_stems) stems_conv = (int)c;
            const mc_bw_heads_args& a = h->bwd_hargs;
            HeadBwdParams p;
            for (int k = 0; k < kNumPred; ++k) { p.pred[k] = a.pred[k]; p.dpred[k] = a.dpred[k]; }
            p.stems = nullptr; p.sums = a.sums; p.coefA = a.coefA; p.coefB = a.coefB; p.att_w = a.att_w;
            p.att_gamma = a.att_gamma; p.att_beta = a.att_beta; p.bank_w = a.bank_w; p.bank_b = a.bank_b; p.w = a.w; p.B = B; p.HW = h->fh * h->fw;
            p.scratch = a.scratch; p.dstems = h->bwd_g[h->t_stems]; p.dw = a.dw; p.dbias = a.dbias; p.datt_w = a.datt_w;
            p.datt_gamma = a.datt_gamma; p.datt_beta = a.datt_beta; p.dbank_w = a.dbank_w; p.dbank_b = a.dbank_b;
            if (h->head_backward_fast) {
                // gradient of the stems straight to the bf16 operand of the stem convolution's dgrad / wgrad, bias gradient from the sums
                if (!h->train_debug) p.dstems = nullptr;
                launch_head_backward_tc(p, T.conv[stems_conv].raw, true, bn.tensors[T.conv[stems_conv].draw].ptr, h->bwd_conv[stems_conv].dbias, st);
                cnt += 5 + 2 * kNumStems;
            } else {
                // the fp32 twin's kernels read fp32 stems: an exact copy of the bf16 tensor (test / A-B path only)
                const long long ne = (long long)B * h->fh * h->fw * kStemTot;
                if (!T.stems_f32) T.stems_f32 = (float*)bn.arena.alloc(sizeof(float) * (size_t)h->max_batch * h->fh * h->fw * kStemTot);
                launch_bf16_to_f32(T.conv[stems_conv].raw, T.stems_f32, ne, st);
                p.stems = T.stems_f32;
                launch_head_backward(p, st);
                cnt += 9;
            }
        } else if (op.type == OP_POOL) {
            const TensorInfo& s = n.tensors[op.src];
            launch_maxpool2_backward_bf16(s.ptr, grad(op.dst), grad(op.src), B, s.C, s.H, s.W, T.op_acc[i] != 0, st);
            ++cnt;
        } else if (op.type == OP_UP) {
            const TensorInfo& s = n.tensors[op.src];
            launch_upsample2_backward_bf16(s.ptr, op.w_dev, grad(op.dst), grad(op.src), h->bwd_up_dw[i], B, s.C, s.H, s.W, T.op_acc[i] != 0, st);
            ++cnt;
        } else {
            const ConvLayer& L = n.convs[op.conv];
            const TrainTc::ConvT& c = T.conv[op.conv];
            const auto& bc = h->bwd_conv[op.conv];
            const auto& bt = h->bn_train[op.conv];
            const TensorInfo& d = n.tensors[L.dst];
            const long long P = (long long)B * d.H * d.W;
            void* draw = bn.tensors[c.draw].ptr;
            if (bt.C > 0) {
                BnBwdTcParams q;
                q.dy = grad(L.dst); q.y = d.ptr; q.raw = c.raw; q.mean = bc.mean; q.inv = bc.inv; q.gamma = bt.gamma; q.sums = T.sums_all + (n.convs.size() + (size_t)op.conv) * 2048; q.sums_zeroed = true;
                q.P = P; q.C = L.cout; q.relu = L.relu ? 1 : 0; q.up = L.stride == 2 ? 1 : 0; q.H = d.H; q.W = d.W; q.draw = draw;
                q.fscale = bt.scale; q.fshift = bt.shift;       // of this batch (the forward's bn_finalize)
                q.dres = L.residual >= 0 ? grad(L.residual) : nullptr; q.dres_acc = c.res_acc ? 1 : 0; q.dgamma = bc.dgamma; q.dbeta = bc.dbeta;
                launch_bn_backward_bf16(q, st);
                cnt += 2;
            } else {
                // the head stems (bias only): the gradient of the raw output is the fp32 gradient the head backward wrote
                const float* dst = h->bwd_g[L.dst];
                MC_CHECK(dst != nullptr && L.stride == 1, "bf16 training: a convolution without BatchNorm is the head-stem convolution");
                if (!h->head_backward_fast) {           // the restructured heads backward has written both already
                    launch_colsum(dst, P, L.cout, h->bwd_sums, bc.dbias, st);
                    launch_f32_to_bf16(dst, draw, P * L.cout, st);
                    cnt += 3;
                }
            }
            if (T.side_stream) {
                // the gradient of the raw output is complete (and the zeroing of dw, first stage of a pass, is behind it in `st`)
                MC_CUDA(cudaEventRecord(T.ev_ready, st));
                MC_CUDA(cudaStreamWaitEvent(T.st_w, T.ev_ready, 0));
                wgrad_tc_launch(*c.wg, B, T.st_w);
                side_used = true;
            } else {
                wgrad_tc_launch(*c.wg, B, st);
            }
            for (int ci : c.dgrad) bn.run_conv(ci, B, st);
            cnt += 1 + (int)c.dgrad.size();
        }
    }
    if (side_used) {                              // the segment's parameter gradients are final once `st` has passed this point
        MC_CUDA(cudaEventRecord(T.ev_done, T.st_w));
        MC_CUDA(cudaStreamWaitEvent(st, T.ev_done, 0));
    }
    h->launches += cnt;
}

}  // namespace mc
