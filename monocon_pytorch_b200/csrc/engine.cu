// Net: tensors, convolution layers, op list, arena, launch sequence.
#include "engine.h"

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
    MC_CUDA(cudaMalloc(&p, bytes));
    MC_CUDA(cudaMemset(p, 0, bytes));
    blocks_.push_back(p);
    total_ += bytes;
    return p;
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

void Net::allocate() {
    for (auto& t : tensors)
        if (!t.ptr) t.ptr = arena.alloc(t.bytes);
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
    L.use_tc2 = (dt == DT_BF16) && (conv_impl == 0) && tc2_conv_supported(*this, L);
    L.use_tc3 = !L.use_tc2 && (dt == DT_BF16) && (conv_impl == 0) && tc3_conv_supported(*this, L);
    L.use_tc = L.use_tc2 || L.use_tc3 || ((dt == DT_BF16) && (conv_impl == 0) && tc_conv_supported(*this, L));
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
            if (!(L.k == 7 && L.cin == 3)) throw;
            std::fprintf(stderr, "[monocon_b200] tensor-core stem unavailable (%s); using the FFMA stem\n", e.what());
            L.use_tc = false;
            L.use_tc2 = false;
            L.tc.reset();
            L.tc2.reset();
        }
    }
    if (!L.use_tc) {
        L.w_simt = (float*)arena.alloc(sizeof(float) * w.size());
        MC_CUDA(cudaMemcpy(L.w_simt, w.data(), sizeof(float) * w.size(), cudaMemcpyHostToDevice));
    }
}

void Net::run_ops(int B, cudaStream_t st, int first, int last) {
    if (last < 0) last = (int)ops.size();
    for (int i = first; i < last; ++i) {
        const Op& op = ops[i];
        if (op.type == OP_CONV) {
            const ConvLayer& L = convs[op.conv];
            if (L.use_tc2) {
                tc2_conv_launch(*this, L, B, st);
            } else if (L.use_tc3) {
                tc3_conv_launch(*this, L, B, st);
            } else if (L.use_tc) {
                tc_conv_launch(*this, L, B, st);
            } else {
                MC_CHECK(L.w_simt != nullptr, "conv not packed: " + L.name);
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
                p.residual = L.residual >= 0 ? tensors[L.residual].ptr : nullptr;
                p.dst = d.ptr;
                p.relu = L.relu ? 1 : 0;
                launch_conv_simt(p, dt, st);
            }
            ++launches_last_run;
        } else if (op.type == OP_POOL) {
            const TensorInfo& s = tensors[op.src];
            launch_maxpool2(s.ptr, tensors[op.dst].ptr, dt, B, s.C, s.H, s.W, st);
            ++launches_last_run;
        } else if (op.type == OP_UP) {
            const TensorInfo& s = tensors[op.src];
            MC_CHECK(op.w_dev != nullptr, "upsample weight missing: " + op.wkey);
            launch_upsample2(s.ptr, tensors[op.dst].ptr, dt, op.w_dev, B, s.C, s.H, s.W, st);
            ++launches_last_run;
        }
    }
}

}  // namespace mc
