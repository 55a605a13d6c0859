// Net: tensors, convolution layers, op list, arena, launch sequence.
#include "engine.h"

#include <cuda_fp16.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace mc {

bool pdl_enabled() {
    static const bool on = []() { const char* e = std::getenv("MC_PDL"); return !(e && e[0] == '0'); }();
    return on;
}

int reserved_sms() {
    const char* e = std::getenv("MC_RESERVE_SMS");
    const int n = (e && e[0]) ? std::atoi(e) : 0;
    return n < 0 ? 0 : (n > 64 ? 64 : n);
}

DeviceArena::~DeviceArena() {
    for (void* p : blocks_) cudaFree(p);
}

void* DeviceArena::alloc(size_t bytes) {
    void* p = nullptr;
    bytes = (bytes + 1023) / 1024 * 1024;
    if (replay_) {
        MC_CHECK(replay_idx_ < replay_end_ && sizes_[replay_idx_] == bytes, "arena replay: the allocation sequence differs from the recorded one");
        p = blocks_[replay_idx_++];
        MC_CUDA(cudaMemset(p, 0, bytes));
        return p;
    }
    MC_CUDA(cudaMalloc(&p, bytes));
    MC_CUDA(cudaMemset(p, 0, bytes));
    blocks_.push_back(p);
    sizes_.push_back(bytes);
    total_ += bytes;
    return p;
}

void DeviceArena::begin_replay(size_t first, size_t last) {
    MC_CHECK(!replay_ && first <= last && last <= blocks_.size(), "arena replay range");
    replay_ = true;
    replay_idx_ = first;
    replay_end_ = last;
}

void DeviceArena::end_replay() {
    const bool complete = replay_idx_ == replay_end_;
    replay_ = false;
    MC_CHECK(complete, "arena replay: fewer allocations than recorded");
}

Net::Net(int device_, int max_batch_, DType dt_, int conv_impl_)
    : device(device_), max_batch(max_batch_), dt(dt_), conv_impl(conv_impl_) {}

Net::~Net() {}

int Net::add_tensor(const std::string& name, int C, int H, int W, int Wp, int xoff) {
    TensorInfo t;
    t.name = name;
    t.C = C; t.H = H; t.W = W;
    t.Wp = Wp ? Wp : W;
    t.xoff = xoff;
    t.bytes = (size_t)max_batch * H * t.Wp * C * dtype_size(dt);
    t.dt = dt;
    t.plane = dt == DT_SPLIT ? (long long)max_batch * H * t.Wp * C : 0;
    tensors.push_back(t);
    aliases_[name] = (int)tensors.size() - 1;
    return (int)tensors.size() - 1;
}

int Net::add_conv(const std::string& name, const std::vector<int>& src, int cout, int k, int stride, int pad,
                  const std::vector<ConvLayer::Part>& parts, int residual, bool relu, int cin_logical) {
    MC_CHECK(!src.empty() && (int)src.size() <= kMaxSrc, "conv sources");
    ConvLayer L;
    L.name = name;
    L.src = src;
    L.k = k; L.stride = stride; L.pad = pad;
    L.cout = cout;
    L.cin_store = 0;
    const TensorInfo s0 = tensors[src[0]];            // by value: add_tensor below may reallocate `tensors`
    for (int s : src) {
        MC_CHECK(tensors[s].H == s0.H && tensors[s].W == s0.W, "conv sources must share H, W: " + name);
        L.cin_store += tensors[s].C;
    }
    L.cin = cin_logical ? cin_logical : L.cin_store;
    const int Ho = (s0.H + 2 * pad - k) / stride + 1, Wo = (s0.W + 2 * pad - k) / stride + 1;
    L.dst = add_tensor(name, cout, Ho, Wo);
    L.residual = residual;
    if (residual >= 0) {
        const TensorInfo& r = tensors[residual];
        MC_CHECK(r.C == cout && r.H == Ho && r.W == Wo, "residual geometry: " + name);
    }
    L.relu = relu;
    L.parts = parts;
    L.flops_per_image = 2.0 * Ho * Wo * (double)cout * k * k * L.cin;
    L.bytes_per_image = 2.0 * ((double)s0.H * s0.W * L.cin_store + (double)Ho * Wo * cout);   // bf16 in + out
    convs.push_back(L);
    Op op;
    op.type = OP_CONV;
    op.conv = (int)convs.size() - 1;
    ops.push_back(op);
    return L.dst;
}

int Net::add_conv_to(const std::string& name, const std::vector<int>& src, int dst, int k, int stride, int pad, int residual, bool relu) {
    MC_CHECK(!src.empty() && (int)src.size() <= kMaxSrc && dst >= 0 && dst < (int)tensors.size(), "conv sources / destination");
    ConvLayer L;
    L.name = name;
    L.src = src;
    L.k = k; L.stride = stride; L.pad = pad;
    L.cout = tensors[dst].C;
    L.cin_store = 0;
    const TensorInfo& s0 = tensors[src[0]];
    for (int s : src) {
        MC_CHECK(tensors[s].H == s0.H && tensors[s].W == s0.W, "conv sources must share H, W: " + name);
        L.cin_store += tensors[s].C;
    }
    L.cin = L.cin_store;
    const int Ho = (s0.H + 2 * pad - k) / stride + 1, Wo = (s0.W + 2 * pad - k) / stride + 1;
    MC_CHECK(tensors[dst].H == Ho && tensors[dst].W == Wo, "destination geometry: " + name);
    L.dst = dst;
    L.residual = residual;
    if (residual >= 0) {
        const TensorInfo& r = tensors[residual];
        MC_CHECK(r.C == L.cout && r.H == Ho && r.W == Wo, "residual geometry: " + name);
    }
    L.relu = relu;
    L.flops_per_image = 2.0 * Ho * Wo * (double)L.cout * k * k * L.cin;
    L.bytes_per_image = 2.0 * ((double)s0.H * s0.W * L.cin_store + (double)Ho * Wo * L.cout);
    convs.push_back(L);
    return (int)convs.size() - 1;
}

int Net::add_pool(int src) {
    auto it = pooled_.find(src);
    if (it != pooled_.end()) return it->second;       // the reference pools the same tensor twice (dla.py:193)
    const TensorInfo s = tensors[src];
    int dst = add_tensor(s.name + ".pool", s.C, s.H / 2, s.W / 2);
    Op op;
    op.type = OP_POOL;
    op.src = src; op.dst = dst;
    ops.push_back(op);
    pooled_[src] = dst;
    return dst;
}

int Net::add_up(int src, const std::string& wkey) {
    const TensorInfo s = tensors[src];
    int dst = add_tensor(wkey, s.C, s.H * 2, s.W * 2);
    Op op;
    op.type = OP_UP;
    op.src = src; op.dst = dst;
    op.wkey = wkey;
    ops.push_back(op);
    return dst;
}

int Net::add_dcn_conv(const std::string& name, const std::vector<int>& src, int off, int cout, const std::vector<ConvLayer::Part>& parts, bool relu,
                      bool mask_logits) {
    const TensorInfo o = tensors[off];
    const TensorInfo s0 = tensors[src[0]];
    MC_CHECK(o.H == s0.H && o.W == s0.W && o.C >= 27, "deformable convolution: offset tensor of " + name);
    const int dst = add_conv(name, src, cout, 3, 1, 1, parts, -1, relu);
    convs.back().dcn_off = off;
    convs.back().dcn_mask_logits = mask_logits;
    return dst;
}

int Net::add_dcn_columns(const std::string& name, const std::vector<int>& src, int off, bool mask_logits) {
    MC_CHECK(!src.empty() && src.size() <= 2, "deformable columns: one or two sources");
    const TensorInfo s0 = tensors[src[0]];
    const TensorInfo o = tensors[off];
    int cin = 0;
    for (int s : src) {
        MC_CHECK(tensors[s].H == s0.H && tensors[s].W == s0.W && tensors[s].Wp == tensors[s].W && tensors[s].C % 8 == 0, "deformable columns: source geometry of " + name);
        cin += tensors[s].C;
    }
    MC_CHECK(o.H == s0.H && o.W == s0.W && o.Wp == o.W && o.C >= 27, "deformable columns: offset tensor of " + name);
    const int dst = add_tensor(name, 9 * cin, s0.H, s0.W);
    Op op;
    op.type = OP_DCN_COL;
    op.srcs = src; op.off = off; op.dst = dst; op.mask_logits = mask_logits;
    ops.push_back(op);
    return dst;
}

void Net::allocate() {
    for (auto& t : tensors)
        if (!t.ptr) t.ptr = arena.alloc(t.bytes);
    if (dt == DT_SPLIT && !d_actscale) {
        d_actscale = (ActScale*)arena.alloc(sizeof(ActScale) * tensors.size());
        d_amax = (unsigned*)arena.alloc(sizeof(unsigned) * tensors.size());
        set_act_exponents(std::vector<int>(tensors.size(), 0));
    }
}

void Net::set_tensor_dtype(int tensor, DType t) {
    TensorInfo& ti = tensors[tensor];
    MC_CHECK(ti.ptr == nullptr, "set_tensor_dtype: before allocate()");
    ti.dt = t;
    ti.bytes = (size_t)max_batch * ti.H * ti.Wp * ti.C * dtype_size(t);
    ti.plane = t == DT_SPLIT ? (long long)max_batch * ti.H * ti.Wp * ti.C : 0;
}

SplitInfo Net::split_info(int tensor) const {
    SplitInfo s;
    if (tensors[tensor].dt == DT_SPLIT) {
        s.plane = tensors[tensor].plane;
        s.interleaved = tensors[tensor].hl_interleaved;
        s.sc = act_scale(tensor);
        s.amax = act_amax(tensor);
    }
    return s;
}

void Net::set_act_exponents(const std::vector<int>& e) {
    MC_CHECK(d_actscale != nullptr && e.size() == tensors.size(), "set_act_exponents: DT_SPLIT net, one exponent per tensor");
    std::vector<ActScale> h(e.size());
    for (size_t i = 0; i < e.size(); ++i) { h[i].mul = std::ldexp(1.f, e[i]); h[i].inv = std::ldexp(1.f, -e[i]); }
    MC_CUDA(cudaMemcpy(d_actscale, h.data(), sizeof(ActScale) * h.size(), cudaMemcpyHostToDevice));
    act_exp = e;
}

std::vector<float> Net::read_act_amax(bool reset) {
    MC_CHECK(d_amax != nullptr, "read_act_amax: DT_SPLIT net");
    std::vector<float> h(tensors.size());
    MC_CUDA(cudaDeviceSynchronize());
    MC_CUDA(cudaMemcpy(h.data(), d_amax, sizeof(float) * h.size(), cudaMemcpyDeviceToHost));   // bit patterns of non-negative floats
    if (reset) MC_CUDA(cudaMemset(d_amax, 0, sizeof(unsigned) * h.size()));
    return h;
}

float* Net::upload_split_scale(const ConvLayer& L, const std::vector<int>& ew) {
    MC_CHECK((int)ew.size() == L.cout && L.scale != nullptr, "upload_split_scale");
    std::vector<float> sc(L.cout);
    MC_CUDA(cudaMemcpy(sc.data(), L.scale, sizeof(float) * L.cout, cudaMemcpyDeviceToHost));
    for (int c = 0; c < L.cout; ++c) sc[c] = std::ldexp(sc[c], -ew[c]);
    float* d = (float*)arena.alloc(sizeof(float) * L.cout);
    MC_CUDA(cudaMemcpy(d, sc.data(), sizeof(float) * L.cout, cudaMemcpyHostToDevice));
    return d;
}

std::vector<int> split_weight_exponents(const std::vector<float>& w_oihw, int cout) {
    std::vector<int> ew(cout, 0);
    const size_t per = w_oihw.size() / (size_t)cout;
    for (int o = 0; o < cout; ++o) {
        float m = 0.f;
        for (size_t i = 0; i < per; ++i) m = std::max(m, std::fabs(w_oihw[(size_t)o * per + i]));
        if (m > 0.f && std::isfinite(m)) {
            int ex = 0;
            std::frexp(m, &ex);                 // m = f * 2^ex, f in [0.5, 1)  ->  m * 2^(14 - ex) in [2^13, 2^14)
            ew[o] = std::max(-100, std::min(100, 14 - ex));
        }
    }
    return ew;
}

uint16_t split_weight_piece(float w, int ew, bool lo) {
    const float ws = std::ldexp(w, ew);
    const __half h = __float2half_rn(ws);
    const __half r = lo ? __float2half_rn(ws - __half2float(h)) : h;
    uint16_t bits;
    std::memcpy(&bits, &r, 2);
    return bits;
}

uint16_t bf16_bits(float v) {
    const bf16 b = __float2bfloat16(v);
    uint16_t bits;
    std::memcpy(&bits, &b, 2);
    return bits;
}

void Net::pack_conv(ConvLayer& L, const std::vector<float>& w_oihw, const std::vector<float>& scale,
                    const std::vector<float>& shift) {
    const int kk = L.k * L.k;
    MC_CHECK((int)w_oihw.size() == L.cout * L.cin * kk, "weight size of " + L.name);
    MC_CHECK((int)scale.size() == L.cout && (int)shift.size() == L.cout, "scale/shift size of " + L.name);
    // [tap][cin_store][cout], zero rows for padded storage channels (stem: 3 -> 4)
    std::vector<float> w((size_t)kk * L.cin_store * L.cout, 0.f);
    for (int o = 0; o < L.cout; ++o)
        for (int c = 0; c < L.cin; ++c)
            for (int t = 0; t < kk; ++t)
                w[((size_t)t * L.cin_store + c) * L.cout + o] = w_oihw[((size_t)o * L.cin + c) * kk + t];
    const bool tc_dt = dt == DT_BF16 || dt == DT_SPLIT;
    MC_CHECK(dt != DT_SPLIT || conv_impl == 0, "the fp16-plane storage of MC_PREC_FP32_TC has tensor-core convolutions only");
    if (L.dcn_off >= 0) {
        // fused deformable convolution: tensor-core kernel only (the FFMA twin runs the unfused plan: columns + 1x1 layer)
        MC_CHECK(conv_impl == 0 && dcn_tc_supported(*this, L), "fused deformable convolution not available for " + L.name + " (MC_DCN_FUSE=0 plans the unfused stages)");
        L.use_tc = true;
        L.scale = (float*)arena.alloc(sizeof(float) * L.cout);
        L.shift = (float*)arena.alloc(sizeof(float) * L.cout);
        MC_CUDA(cudaMemcpy(L.scale, scale.data(), sizeof(float) * L.cout, cudaMemcpyHostToDevice));
        MC_CUDA(cudaMemcpy(L.shift, shift.data(), sizeof(float) * L.cout, cudaMemcpyHostToDevice));
        dcn_tc_prepare(*this, L, w_oihw);
        return;
    }
    L.use_tc2 = tc_dt && (conv_impl == 0) && tc2_conv_supported(*this, L);
    L.use_tc3 = !L.use_tc2 && tc_dt && (conv_impl == 0) && tc3_conv_supported(*this, L);
    L.use_tc = L.use_tc2 || L.use_tc3 || (tc_dt && (conv_impl == 0) && tc_conv_supported(*this, L));
    MC_CHECK(dt != DT_SPLIT || L.use_tc, "no tensor-core kernel covers this layer in the fp32-accurate mode: " + L.name);
    // Tree.downsample of this layer's output (dla.py:193) folded into its epilogue: the resident-weight halo kernel (unstacked
    // layers) and the tap-box kernel pool with quad shuffles.  MC_POOL_FUSE=0 keeps the separate kernel.
    L.pool_dst = -1;
    {
        const char* e = std::getenv("MC_POOL_FUSE");
        auto it = pooled_.find(L.dst);
        const TensorInfo& d = tensors[L.dst];
        const bool even = d.H % 2 == 0 && d.W % 2 == 0;
        const bool kernel_ok = (L.use_tc2 && !(L.cout == 16 && L.residual < 0)) || (!L.use_tc2 && !L.use_tc3);
        if (dt == DT_SPLIT && it != pooled_.end() && even && kernel_ok && L.residual < 0 && !(e && e[0] == '0')) L.pool_dst = it->second;
    }
    L.scale = (float*)arena.alloc(sizeof(float) * L.cout);
    L.shift = (float*)arena.alloc(sizeof(float) * L.cout);
    MC_CUDA(cudaMemcpy(L.scale, scale.data(), sizeof(float) * L.cout, cudaMemcpyHostToDevice));
    MC_CUDA(cudaMemcpy(L.shift, shift.data(), sizeof(float) * L.cout, cudaMemcpyHostToDevice));
    if (L.use_tc) {
        try {
            if (L.use_tc2) tc2_conv_prepare(*this, L, w_oihw);
            else if (L.use_tc3) tc3_conv_prepare(*this, L, w_oihw);
            else tc_conv_prepare(*this, L, w_oihw);
        } catch (const std::exception& e) {
            // only the overlapping-window stem view is allowed to degrade (to the FFMA kernel, still on the GPU)
            if (!(L.k == 7 && L.cin == 3) || dt == DT_SPLIT) throw;
            std::fprintf(stderr, "[monocon_b200] tensor-core stem unavailable (%s); using the FFMA stem\n", e.what());
            L.use_tc = false;
            L.use_tc2 = false;
            L.tc.reset();
            L.tc2.reset();
        }
    }
    if (!L.use_tc || keep_master) {
        L.w_simt = (float*)arena.alloc(sizeof(float) * w.size());
        MC_CUDA(cudaMemcpy(L.w_simt, w.data(), sizeof(float) * w.size(), cudaMemcpyHostToDevice));
    }
}

void Net::run_conv(int conv, int B, cudaStream_t st) {
    const ConvLayer& L = convs[conv];
    if (L.dcn_off >= 0) {
        dcn_tc_launch(*this, L, B, st);
    } else if (L.use_tc2) {
        tc2_conv_launch(*this, L, B, st);
    } else if (L.use_tc3) {
        tc3_conv_launch(*this, L, B, st);
    } else if (L.use_tc) {
        tc_conv_launch(*this, L, B, st);
    } else {
        MC_CHECK(L.w_simt != nullptr, "conv not packed: " + L.name);
        MC_CHECK(!L.dst_override_f32, "fp32 raw output needs the streamed-weight tensor-core kernel: " + L.name);
        ConvParams p;
        std::memset(&p, 0, sizeof(p));
        p.nsrc = (int)L.src.size();
        for (int s = 0; s < p.nsrc; ++s) {
            p.src[s] = tensors[L.src[s]].ptr;
            p.srcC[s] = tensors[L.src[s]].C;
            p.srcWp[s] = tensors[L.src[s]].Wp;
            p.srcXoff[s] = tensors[L.src[s]].xoff;
        }
        const TensorInfo& s0 = tensors[L.src[0]];
        const TensorInfo& d = tensors[L.dst];
        p.B = B; p.Hin = s0.H; p.Win = s0.W; p.Hout = d.H; p.Wout = d.W;
        p.Cin = L.cin_store; p.Cout = L.cout;
        p.k = L.k; p.stride = L.stride; p.pad = L.pad;
        p.w = L.w_simt; p.scale = L.scale; p.shift = L.shift;
        p.residual = (L.residual >= 0 && !L.dst_override) ? tensors[L.residual].ptr : nullptr;     // dst_override: the raw output
        p.dst = L.dst_override ? L.dst_override : d.ptr;
        p.relu = (L.relu && !L.dst_override) ? 1 : 0;
        launch_conv_simt(p, dt, st);
    }
}

void Net::run_ops(int B, cudaStream_t st, int first, int last) {
    if (last < 0) last = (int)ops.size();
    for (int i = first; i < last; ++i) {
        const Op& op = ops[i];
        if (op.type == OP_CONV) {
            run_conv(op.conv, B, st);
            ++launches_last_run;
        } else if (op.type == OP_POOL) {
            bool fused = false;
            for (const auto& L : convs)
                if (L.dst == op.src && L.pool_dst == op.dst) fused = true;
            if (fused) continue;                       // written by the producing convolution's epilogue
            const TensorInfo& s = tensors[op.src];
            launch_maxpool2(s.ptr, tensors[op.dst].ptr, dt, B, s.C, s.H, s.W, st, split_info(op.src), split_info(op.dst));
            ++launches_last_run;
        } else if (op.type == OP_UP) {
            const TensorInfo& s = tensors[op.src];
            MC_CHECK(op.w_dev != nullptr, "upsample weight missing: " + op.wkey);
            launch_upsample2(s.ptr, tensors[op.dst].ptr, dt, op.w_dev, B, s.C, s.H, s.W, st, split_info(op.src), split_info(op.dst));
            ++launches_last_run;
        } else if (op.type == OP_DCN_COL) {
            DcnColParams p;
            std::memset(&p, 0, sizeof(p));
            p.nsrc = (int)op.srcs.size();
            for (int s = 0; s < p.nsrc; ++s) {
                const TensorInfo& t = tensors[op.srcs[s]];
                const SplitInfo si = split_info(op.srcs[s]);
                MC_CHECK(t.dt == dt, "deformable columns: source storage type");
                p.src[s] = t.ptr; p.srcC[s] = t.C; p.src_plane[s] = si.plane; p.src_sc[s] = si.sc;
                p.Cin += t.C;
            }
            // concatenated sources meet in the offset convolution's K dimension, so they share one exponent (mc_calibrate_scales)
            MC_CHECK(dt != DT_SPLIT || p.nsrc == 1 || act_exp.empty() || act_exp[op.srcs[0]] == act_exp[op.srcs[1]], "deformable columns: sources with different scales");
            const TensorInfo& o = tensors[op.off];
            const TensorInfo& c = tensors[op.dst];
            MC_CHECK(o.dt == dt && c.dt == dt, "deformable columns: storage types");
            const SplitInfo so = split_info(op.off), sc = split_info(op.dst);
            p.off = o.ptr; p.offC = o.C; p.off_plane = so.plane; p.off_sc = so.sc;
            p.col = c.ptr; p.col_plane = sc.plane; p.col_sc = sc.sc; p.col_amax = sc.amax;
            p.B = B; p.H = c.H; p.W = c.W;
            p.mask_logits = op.mask_logits ? 1 : 0;
            launch_dcn_columns(p, dt, st);
            ++launches_last_run;
        }
    }
}

}  // namespace mc
