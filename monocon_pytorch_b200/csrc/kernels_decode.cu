// Fused decode: 3x3 heat-map NMS -> exact top-k (radix select, deterministic ties) -> gather of the
// regression maps at the k peaks -> 2D box, alpha bins, rot_y, 2D->3D lifting, origin shift, threshold.
// One CTA per image.  sm_100a.
//
// Reference semantics (file:line in the reference repo):
//   get_local_maximum        utils/tensor_ops.py:17-21      keep = (maxpool3x3(h) == h); h * keep
//   get_topk_from_heatmap    utils/tensor_ops.py:24-31      topk over C*H*W, cls = i // HW, ind = i % HW
//   transpose_and_gather_feat utils/tensor_ops.py:34-59     rows of the NHWC-viewed maps at `ind`
//   decode_heatmap           model/dense_heads/monocon_heads.py:399-482
//   decode_alpha             :379-396     calculate_roty :485-515     convert_pts2D_to_pts3D :518-558
//   _get_bboxes              :313-329     (y += 0.5 * dim[1])
// Precondition: heat-map values are >= 0 (they are clamp(sigmoid) in [1e-4, 1-1e-4], :168-170), so the
// IEEE bit pattern orders like the value.  Ties are broken by the lowest flat index (torch.topk leaves
// the order of equal elements unspecified).
#include "common.cuh"

namespace mc {

constexpr int kDecThreads = 1024;
constexpr int kMaxTopk = 128;

__device__ __forceinline__ int block_ordered_offset(bool flag, int* warp_cnt, int& total) {
    // returns the ordered (by thread id) rank of this thread among flagged threads of the CTA
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned bal = __ballot_sync(0xffffffffu, flag);
    const int wpre = __popc(bal & ((1u << lane) - 1u));
    __syncthreads();                       // protect warp_cnt reuse
    if (lane == 0) warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int woff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kDecThreads / 32; ++w) {
        const int c = warp_cnt[w];
        if (w < warp) woff += c;
        tot += c;
    }
    total = tot;
    return woff + wpre;
}

__global__ void __launch_bounds__(kDecThreads) decode_kernel(const DecodeParams p, unsigned* __restrict__ cand_key_all,
                                                             int* __restrict__ cand_idx_all) {
    __shared__ int warp_cnt[kDecThreads / 32];
    __shared__ unsigned hist[256];
    __shared__ unsigned s_prefix, s_need;
    __shared__ int s_count, s_nsel;
    __shared__ unsigned sel_key[kMaxTopk];
    __shared__ int sel_idx[kMaxTopk];
    __shared__ unsigned srt_key[kMaxTopk];
    __shared__ int srt_idx[kMaxTopk];

    const int b = blockIdx.x, tid = threadIdx.x;
    const int HW = p.H * p.W, N = p.C * HW, K = p.topk;
    const float* heat = p.pred[0] + (long long)b * N;
    unsigned* cand_key = cand_key_all + (long long)b * N;
    int* cand_idx = cand_idx_all + (long long)b * N;

    // ---- phase 1: NMS + ordered compaction of the surviving non-zero peaks ---------------------
    int base_off = 0;
    for (int base = 0; base < N; base += kDecThreads) {
        const int i = base + tid;
        bool keep = false;
        float v = 0.f;
        if (i < N) {
            const int c = i / HW, r = i % HW, y = r / p.W, x = r % p.W;
            v = heat[i];
            float m = v;
            const float* hc = heat + (long long)c * HW;
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy) {
                const int yy = y + dy;
                if (yy < 0 || yy >= p.H) continue;
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) {
                    const int xx = x + dx;
                    if (xx < 0 || xx >= p.W) continue;
                    m = fmaxf(m, hc[yy * p.W + xx]);
                }
            }
            keep = (m == v) && (__float_as_uint(v) != 0u) && (v > 0.f);
        }
        int total;
        const int rank = block_ordered_offset(keep, warp_cnt, total);
        if (keep) {
            cand_key[base_off + rank] = __float_as_uint(v);
            cand_idx[base_off + rank] = i;
        }
        base_off += total;
    }
    const int Nc = base_off;     // identical in every thread
    __syncthreads();             // candidates visible CTA-wide (global writes + barrier)

    if (tid == 0) s_nsel = 0;
    __syncthreads();

    if (Nc <= K) {
        // fewer peaks than k: take them all, then pad with the lowest-index zero-valued cells
        for (int j = tid; j < Nc; j += kDecThreads) { sel_key[j] = cand_key[j]; sel_idx[j] = cand_idx[j]; }
        __syncthreads();
        if (tid == 0) {
            int n = Nc, cj = 0;
            for (int i = 0; i < N && n < K; ++i) {
                while (cj < Nc && cand_idx[cj] < i) ++cj;
                if (cj < Nc && cand_idx[cj] == i) continue;     // a kept peak, already selected
                sel_key[n] = 0u; sel_idx[n] = i; ++n;
            }
            s_nsel = n;
        }
        __syncthreads();
    } else {
        // ---- phase 2: 4 x 8-bit MSB-first radix select of the K-th largest key ------------------
        if (tid == 0) { s_prefix = 0u; s_need = (unsigned)K; }
        __syncthreads();
        for (int shift = 24; shift >= 0; shift -= 8) {
            if (tid < 256) hist[tid] = 0u;
            __syncthreads();
            const unsigned prefix = s_prefix;
            const unsigned himask = (shift == 24) ? 0u : (0xffffffffu << (shift + 8));
            for (int j = tid; j < Nc; j += kDecThreads) {
                const unsigned key = cand_key[j];
                if ((key & himask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (tid == 0) {
                unsigned need = s_need, acc = 0u;
                int d = 255;
                for (; d > 0; --d) {
                    if (acc + hist[d] >= need) break;
                    acc += hist[d];
                }
                s_need = need - acc;                 // how many are needed from bin d
                s_prefix = prefix | ((unsigned)d << shift);
            }
            __syncthreads();
        }
        const unsigned T = s_prefix;                  // K-th largest key
        const int need_ties = (int)s_need;            // >= 1 elements equal to T, lowest indices first
        // ---- phase 3: collect  key > T  and the first need_ties ties (list is index-ordered) ----
        int tie_base = 0;
        for (int base = 0; base < Nc; base += kDecThreads) {
            const int j = base + tid;
            unsigned key = 0u;
            int idx = 0;
            if (j < Nc) { key = cand_key[j]; idx = cand_idx[j]; }
            const bool gt = (j < Nc) && key > T;
            const bool tie = (j < Nc) && key == T;
            int total;
            const int trank = block_ordered_offset(tie, warp_cnt, total);
            if (gt || (tie && (tie_base + trank) < need_ties)) {
                const int pos = atomicAdd(&s_nsel, 1);
                if (pos < kMaxTopk) { sel_key[pos] = key; sel_idx[pos] = idx; }
            }
            tie_base += total;
        }
        __syncthreads();
    }

    // ---- phase 4: rank sort (key desc, index asc) ------------------------------------------------
    const int nsel = min(s_nsel, K);
    if (tid < nsel) {
        const unsigned k0 = sel_key[tid];
        const int i0 = sel_idx[tid];
        int rank = 0;
        for (int j = 0; j < nsel; ++j) {
            const unsigned kj = sel_key[j];
            const int ij = sel_idx[j];
            rank += (kj > k0) || (kj == k0 && ij < i0);
        }
        srt_key[rank] = k0;
        srt_idx[rank] = i0;
    }
    __syncthreads();

    // ---- phase 5: gather + lift, one thread per detection ----------------------------------------
    if (tid < K) {
        const int t = tid;
        const float score = (t < nsel) ? __uint_as_float(srt_key[t]) : 0.f;
        const int flat = (t < nsel) ? srt_idx[t] : 0;
        const int cls = flat / HW, ind = flat % HW;
        const int ys_i = ind / p.W, xs_i = ind % p.W;
        const float xs = (float)xs_i, ys = (float)ys_i;
        auto at = [&](int pi, int nch, int ch) -> float {
            return p.pred[pi][((long long)b * nch + ch) * HW + ind];
        };
        // 2D box, monocon_heads.py:416-428
        const float w0 = at(2, 2, 0), w1 = at(2, 2, 1);
        const float tx = xs + at(3, 2, 0), ty = ys + at(3, 2, 1);
        const float x1 = (tx - w0 / 2.f) * p.scale_x, y1 = (ty - w1 / 2.f) * p.scale_y;
        const float x2 = (tx + w0 / 2.f) * p.scale_x, y2 = (ty + w1 / 2.f) * p.scale_y;
        // alpha, monocon_heads.py:379-396
        int acls = 0;
        float best = at(8, p.num_bins, 0);
        for (int k = 1; k < p.num_bins; ++k) {
            const float v = at(8, p.num_bins, k);
            if (v > best) { best = v; acls = k; }
        }
        const float PI_F = 3.14159265358979323846f;
        const float TWO_PI_F = (float)(2.0 * 3.14159265358979323846);
        const float angle_per_class = (float)((2.0 * 3.14159265358979323846) / (double)p.num_bins);
        float alpha = (float)acls * angle_per_class + at(9, p.num_bins, acls);
        if (alpha > PI_F) alpha -= TWO_PI_F;
        if (alpha < -PI_F) alpha += TWO_PI_F;
        // uncertainty-weighted score, monocon_heads.py:439-441
        const float d0 = at(7, 2, 0), d1 = at(7, 2, 1);
        const float sc = score * expf(-d1);
        // projected centre = 9th keypoint, monocon_heads.py:443-457
        const float cu = (at(5, p.c2k_channels, p.c2k_channels - 2) + xs) * p.scale_x;
        const float cv = (at(5, p.c2k_channels, p.c2k_channels - 1) + ys) * p.scale_y;
        // rot_y, monocon_heads.py:507-513
        const float* P = p.P2 + (long long)b * 12;
        float rot = alpha + atan2f(cu - P[2], P[0]);
        while (rot > PI_F) rot -= TWO_PI_F;
        while (rot < -PI_F) rot += TWO_PI_F;
        // lift, monocon_heads.py:518-558:  [u d, v d, d, 1] @ inv(viewpad)^T
        const float* I = p.invP + (long long)b * 16;
        const float h0 = cu * d0, h1 = cv * d0, h2 = d0, h3 = 1.f;
        float X = h0 * I[0] + h1 * I[1] + h2 * I[2] + h3 * I[3];
        float Y = h0 * I[4] + h1 * I[5] + h2 * I[6] + h3 * I[7];
        float Z = h0 * I[8] + h1 * I[9] + h2 * I[10] + h3 * I[11];
        const float dm0 = at(6, 3, 0), dm1 = at(6, 3, 1), dm2 = at(6, 3, 2);
        Y += dm1 * 0.5f;                                     // _get_bboxes origin shift, :320-328
        const long long o = (long long)b * K + t;
        float* b2 = p.box2d + o * 5;
        b2[0] = x1; b2[1] = y1; b2[2] = x2; b2[3] = y2; b2[4] = sc;
        float* b3 = p.box3d + o * 7;
        b3[0] = X; b3[1] = Y; b3[2] = Z; b3[3] = dm0; b3[4] = dm1; b3[5] = dm2; b3[6] = rot;
        p.labels[o] = cls;
        p.inds[o] = ind;
        p.valid[o] = (sc > p.thres) ? 1 : 0;
    }
}

void launch_decode(const DecodeParams& p, unsigned* cand_key, int* cand_idx, cudaStream_t st) {
    MC_CHECK(p.topk >= 1 && p.topk <= kMaxTopk, "decode: topk must be in [1,128]");
    MC_CHECK(p.C * p.H * p.W >= p.topk, "decode: topk larger than the heat-map");
    decode_kernel<<<p.B, kDecThreads, 0, st>>>(p, cand_key, cand_idx);
    MC_CUDA(cudaGetLastError());
}

}  // namespace mc
