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
constexpr int kNmsThreads = 256;

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

// ---- kernel 1: 3x3 NMS over the whole batch; survivors (value > 0) are appended, unordered, as 64-bit composites
//      comp = (IEEE bits of the score << idx_bits) | (idx_mask - flat_index): larger composite == larger score, then
//      lower flat index, and composites are unique, so "top-k with lowest-index tie-break" is a plain order statistic.
__global__ void __launch_bounds__(kNmsThreads) decode_nms_kernel(const float* __restrict__ heat_all, int C, int H, int W,
                                                                 int idx_bits, unsigned long long* __restrict__ cand_all,
                                                                 int* __restrict__ count_all) {
    pdl_sync();
    const int b = blockIdx.y;
    const int HW = H * W, N = C * HW;
    const float* heat = heat_all + (long long)b * N;
    unsigned long long* cand = cand_all + (long long)b * N;
    const unsigned long long idx_mask = (1ull << idx_bits) - 1ull;
    const int lane = threadIdx.x & 31;
    for (int i = blockIdx.x * kNmsThreads + threadIdx.x; i < ((N + 31) / 32) * 32; i += gridDim.x * kNmsThreads) {
        bool keep = false;
        float v = 0.f;
        if (i < N) {
            const int c = i / HW, r = i % HW, y = r / W, x = r % W;
            v = heat[i];
            float m = v;
            const float* hc = heat + (long long)c * HW;
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy) {
                const int yy = y + dy;
                if (yy < 0 || yy >= H) continue;
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) {
                    const int xx = x + dx;
                    if (xx < 0 || xx >= W) continue;
                    m = fmaxf(m, __ldg(hc + yy * W + xx));
                }
            }
            keep = (m == v) && (v > 0.f);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (bal) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&count_all[b], __popc(bal));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (keep)
                cand[base + __popc(bal & ((1u << lane) - 1u))] =
                    ((unsigned long long)__float_as_uint(v) << idx_bits) | (idx_mask - (unsigned long long)i);
        }
    }
}

// ---- kernel 2: exact k-th largest composite by MSB-first radix select (7-bit digits), gather + lift.  One CTA / image.
__global__ void __launch_bounds__(kDecThreads) decode_kernel(const DecodeParams p, int idx_bits,
                                                             const unsigned long long* __restrict__ cand_all,
                                                             const int* __restrict__ count_all) {
    __shared__ int warp_cnt[kDecThreads / 32];
    __shared__ unsigned hist[128];
    __shared__ unsigned long long s_prefix;
    __shared__ unsigned s_need;
    __shared__ int s_nsel;
    __shared__ unsigned long long sel[kMaxTopk];
    __shared__ unsigned long long srt[kMaxTopk];

    pdl_sync();
    const int b = blockIdx.x, tid = threadIdx.x;
    const int HW = p.H * p.W, N = p.C * HW, K = p.topk;
    const unsigned long long* cand = cand_all + (long long)b * N;
    const unsigned long long idx_mask = (1ull << idx_bits) - 1ull;
    const int Nc = count_all[b];
    if (tid == 0) s_nsel = 0;
    __syncthreads();

    if (Nc <= K) {
        // fewer peaks than k: take them all, then pad with the lowest-index cells that are not peaks (value 0)
        for (int j = tid; j < Nc; j += kDecThreads) sel[j] = cand[j];
        __syncthreads();
        bool flag = false;
        if (tid < 2 * K && tid < N) {         // among the first 2K cells at least K are not peaks
            flag = true;
            for (int j = 0; j < Nc; ++j)
                if ((int)(idx_mask - (sel[j] & idx_mask)) == tid) flag = false;
        }
        int total;
        const int rank = block_ordered_offset(flag, warp_cnt, total);
        if (flag && Nc + rank < K) sel[Nc + rank] = idx_mask - (unsigned long long)tid;      // score bits 0
        __syncthreads();
        if (tid == 0) s_nsel = min(K, Nc + total);
        __syncthreads();
    } else {
        if (tid == 0) { s_prefix = 0ull; s_need = (unsigned)K; }
        __syncthreads();
        for (int shift = 49; shift >= 0; shift -= 7) {          // 8 digits cover 56 >= 32 + idx_bits bits
            if (tid < 128) hist[tid] = 0u;
            __syncthreads();
            const unsigned long long prefix = s_prefix;
            for (int j = tid; j < Nc; j += kDecThreads) {
                const unsigned long long c = cand[j];
                if ((c >> (shift + 7)) == (prefix >> (shift + 7))) atomicAdd(&hist[(unsigned)(c >> shift) & 127u], 1u);
            }
            __syncthreads();
            if (tid < 32) {                                       // warp 0: suffix scan over 128 bins, 4 per lane
                const unsigned h0 = hist[4 * tid], h1 = hist[4 * tid + 1], h2 = hist[4 * tid + 2], h3 = hist[4 * tid + 3];
                const unsigned mine = h0 + h1 + h2 + h3;
                unsigned suf = mine;                              // inclusive suffix sum over lanes >= tid
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const unsigned o = __shfl_down_sync(0xffffffffu, suf, d);
                    if (tid + d < 32) suf += o;
                }
                const unsigned need = s_need;
                const unsigned bal = __ballot_sync(0xffffffffu, suf >= need);
                const int L = 31 - __clz(bal);                    // highest lane whose suffix reaches `need`
                if (tid == L) {
                    unsigned acc = suf - mine;                    // elements in bins above this lane's four
                    int digit;
                    if (acc + h3 >= need) digit = 3;
                    else if (acc + h3 + h2 >= need) { acc += h3; digit = 2; }
                    else if (acc + h3 + h2 + h1 >= need) { acc += h3 + h2; digit = 1; }
                    else { acc += h3 + h2 + h1; digit = 0; }
                    s_need = need - acc;
                    s_prefix = prefix | ((unsigned long long)(4 * tid + digit) << shift);
                }
            }
            __syncthreads();
        }
        const unsigned long long T = s_prefix;                    // the k-th largest composite (composites are unique)
        for (int j = tid; j < Nc; j += kDecThreads) {
            const unsigned long long c = cand[j];
            if (c >= T) {
                const int pos = atomicAdd(&s_nsel, 1);
                if (pos < kMaxTopk) sel[pos] = c;
            }
        }
        __syncthreads();
    }

    // ---- rank sort by composite, descending --------------------------------------------------------
    const int nsel = min(s_nsel, K);
    if (tid < nsel) {
        const unsigned long long c0 = sel[tid];
        int rank = 0;
        for (int j = 0; j < nsel; ++j) rank += sel[j] > c0;
        srt[rank] = c0;
    }
    __syncthreads();

    // ---- multi-GPU: wait until every peer has released this buffer for the generation we are about to write (the
    //      release was sent a whole forward pass ago, so this does not spin in steady state) -------------------
    const GatherParams& G = p.gather;
    unsigned want_gen = 0;
    if (G.n > 0) {
        want_gen = *reinterpret_cast<volatile unsigned*>(G.gen) + 1u;
        if (tid < G.n && tid != G.rank) {
            const volatile unsigned* rf = G.ready + tid;
            const long long t0 = clock64();
            while ((int)(*rf - want_gen) < 0) {
                if (clock64() - t0 > 4000000000LL) {
                    if (G.error_flag) atomicExch(G.error_flag, 21);
                    __threadfence_system();
                    asm volatile("trap;");
                }
            }
        }
        __syncthreads();
    }

    // ---- phase 5: gather + lift, one thread per detection ----------------------------------------
    if (tid < K) {
        const int t = tid;
        const unsigned long long comp = (t < nsel) ? srt[t] : idx_mask;
        const float score = __uint_as_float((unsigned)(comp >> idx_bits));
        const int flat = (int)(idx_mask - (comp & idx_mask));
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
        // the same row into this rank's slot of every peer's gather buffer (plain stores over NVLink)
        for (int r = 0; r < G.n; ++r) {
            if (r == G.rank) continue;
            char* base = G.peer_slot[r];
            float* q2 = reinterpret_cast<float*>(base + G.off_box2d) + o * 5;
            q2[0] = x1; q2[1] = y1; q2[2] = x2; q2[3] = y2; q2[4] = sc;
            float* q3 = reinterpret_cast<float*>(base + G.off_box3d) + o * 7;
            q3[0] = X; q3[1] = Y; q3[2] = Z; q3[3] = dm0; q3[4] = dm1; q3[5] = dm2; q3[6] = rot;
            reinterpret_cast<long long*>(base + G.off_labels)[o] = cls;
            reinterpret_cast<long long*>(base + G.off_inds)[o] = ind;
            reinterpret_cast<unsigned char*>(base + G.off_valid)[o] = (sc > p.thres) ? 1 : 0;
        }
    }
    if (G.n > 0) {
        // one system-scope fence per CTA (a fence in each of the 1024 threads costs ~0.8 ms per launch): the CTA barrier
        // orders every thread's remote stores before thread 0's fence, which is cumulative
        __syncthreads();
        if (tid == 0) {
            __threadfence_system();
            const unsigned prev = atomicAdd(G.done, 1u);
            if (prev == (unsigned)p.B - 1u) {   // last CTA of the launch: publish the generation everywhere
                *G.done = 0u;
                __threadfence_system();
                for (int r = 0; r < G.n; ++r) *reinterpret_cast<volatile unsigned*>(G.peer_data_flag[r]) = want_gen;
                *reinterpret_cast<volatile unsigned*>(G.gen) = want_gen;
                __threadfence_system();
            }
        }
    }
}

// release: tell every peer that this rank has consumed generation gen - 1 of the buffer and that they may write `gen`
__global__ void gather_release_kernel(GatherParams G, unsigned* const* peer_ready_flag, unsigned gen) {
    const int r = threadIdx.x;
    if (r < G.n) *reinterpret_cast<volatile unsigned*>(peer_ready_flag[r]) = gen;
    __threadfence_system();
}

// wait: all ranks' slots of generation >= gen have arrived in the local buffer
__global__ void gather_wait_kernel(const unsigned* data_flag, int n, unsigned gen, int* error_flag) {
    const int r = threadIdx.x;
    if (r < n) {
        const volatile unsigned* f = data_flag + r;
        const long long t0 = clock64();
        while ((int)(*f - gen) < 0) {
            if (clock64() - t0 > 4000000000LL) {
                if (error_flag) atomicExch(error_flag, 22);
                __threadfence_system();
                asm volatile("trap;");
            }
        }
    }
    __threadfence_system();
}

// ---- post-decode KITTI conversion: get_valid_bboxes_3d / convert_to_kitti_3d of
//      utils/kitti_convert_utils.py:16-171 with extract_corners_from_bboxes_3d, rotation_3d_in_axis, points_cam2img of
//      utils/geometry_ops.py:7-163, one thread per detection.  Arithmetic as in the reference: corners and rotation in
//      float32, projection and min / max in float64, image-bounds test against float32 (h, w).
__global__ void kitti_boxes_kernel(const float* __restrict__ box3d, const unsigned char* __restrict__ valid, const float* __restrict__ P2,
                                   const int* __restrict__ img_hw, int B, int K, double* __restrict__ bbox, float* __restrict__ alpha,
                                   unsigned char* __restrict__ keep) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * K) return;
    const int b = i / K;
    const float* bx = box3d + (long long)i * 7;
    const float* P = P2 + (long long)b * 12;
    const float s = sinf(bx[6]), c = cosf(bx[6]);
    double umin = 1e300, vmin = 1e300, umax = -1e300, vmax = -1e300;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        // unit-cube corner order of geometry_ops.py:37-39 (unravel_index re-ordered by [0,1,3,2,4,5,7,6]), origin (0.5, 1, 0.5)
        const int idx = (k == 2) ? 3 : (k == 3) ? 2 : (k == 6) ? 7 : (k == 7) ? 6 : k;
        const float cx = (float)((idx >> 2) & 1) - 0.5f, cy = (float)((idx >> 1) & 1) - 1.0f, cz = (float)(idx & 1) - 0.5f;
        const float x = __fmul_rn(bx[3], cx), y = __fmul_rn(bx[4], cy), z = __fmul_rn(bx[5], cz);
        const float rx = __fadd_rn(__fadd_rn(__fmul_rn(x, c), __fmul_rn(z, s)), bx[0]);
        const float ry = __fadd_rn(y, bx[1]);
        const float rz = __fadd_rn(__fadd_rn(__fmul_rn(-x, s), __fmul_rn(z, c)), bx[2]);
        const double X = rx, Y = ry, Z = rz;
        const double u = X * (double)P[0] + Y * (double)P[1] + Z * (double)P[2] + (double)P[3];
        const double v = X * (double)P[4] + Y * (double)P[5] + Z * (double)P[6] + (double)P[7];
        const double w = X * (double)P[8] + Y * (double)P[9] + Z * (double)P[10] + (double)P[11];
        const double pu = u / w, pv = v / w;
        umin = fmin(umin, pu); umax = fmax(umax, pu); vmin = fmin(vmin, pv); vmax = fmax(vmax, pv);
    }
    const double h = (double)(float)img_hw[2 * b], w_img = (double)(float)img_hw[2 * b + 1];
    const bool ok = valid[i] && (umin < w_img) && (vmin < h) && (umax > 0.0) && (vmax > 0.0);
    keep[i] = ok ? 1 : 0;
    // clip to the image (kitti_convert_utils.py:124-128)
    bbox[(long long)i * 4 + 0] = fmax(umin, 0.0);
    bbox[(long long)i * 4 + 1] = fmax(vmin, 0.0);
    bbox[(long long)i * 4 + 2] = fmin(umax, w_img);
    bbox[(long long)i * 4 + 3] = fmin(vmax, h);
    alpha[i] = __fadd_rn(-atan2f(bx[0], bx[2]), bx[6]);
}

void launch_kitti_boxes(const float* box3d, const unsigned char* valid, const float* P2, const int* img_hw, int B, int K, double* bbox,
                        float* alpha, unsigned char* keep, cudaStream_t st) {
    const int n = B * K;
    kitti_boxes_kernel<<<(n + 127) / 128, 128, 0, st>>>(box3d, valid, P2, img_hw, B, K, bbox, alpha, keep);
    MC_CUDA(cudaGetLastError());
}

void launch_gather_release(const GatherParams& G, unsigned* const* peer_ready_flag_dev, unsigned gen, cudaStream_t st) {
    gather_release_kernel<<<1, 32, 0, st>>>(G, peer_ready_flag_dev, gen);
    MC_CUDA(cudaGetLastError());
}
void launch_gather_wait(const unsigned* data_flag, int n, unsigned gen, int* error_flag, cudaStream_t st) {
    gather_wait_kernel<<<1, 32, 0, st>>>(data_flag, n, gen, error_flag);
    MC_CUDA(cudaGetLastError());
}

void launch_decode(const DecodeParams& p, unsigned long long* cand, int* count, cudaStream_t st) {
    MC_CHECK(p.topk >= 1 && p.topk <= kMaxTopk, "decode: topk must be in [1,128]");
    const int N = p.C * p.H * p.W;
    MC_CHECK(N >= p.topk, "decode: topk larger than the heat-map");
    int idx_bits = 1;
    while ((1 << idx_bits) < N) ++idx_bits;
    MC_CHECK(idx_bits <= 24, "decode: heat-map too large");
    MC_CUDA(cudaMemsetAsync(count, 0, sizeof(int) * p.B, st));
    int chunks = (N + kNmsThreads - 1) / kNmsThreads;
    if (chunks * p.B > 148 * 8) chunks = (148 * 8 + p.B - 1) / p.B;
    dim3 grid(chunks, p.B);
    launch_k(decode_nms_kernel, grid, dim3(kNmsThreads), 0, st, p.pred[0], p.C, p.H, p.W, idx_bits, cand, count);
    launch_k(decode_kernel, dim3(p.B), dim3(kDecThreads), 0, st, p, idx_bits, (const unsigned long long*)cand, (const int*)count);
}

}  // namespace mc
