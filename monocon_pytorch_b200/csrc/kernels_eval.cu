// Rotated-IoU evaluation kernels for sm_100a, behind the C ABI:
//   mc_rotate_iou      rotate_iou_gpu_eval   engine/kitti_eval/rotate_iou.py:337-379 (numba.cuda kernel :266-334)
//   mc_box3d_overlap   d3_box_overlap        engine/kitti_eval/eval.py:128-164 (numba CPU kernel after the BEV pass)
// One thread per (box, query) pair (the reference stages 64 x 64 boxes in shared memory; N, K are a few dozen per frame, the
// boxes stay in L1).  The intersection of two rotated rectangles is the convex polygon spanned by the corners of each that
// lie inside the other plus the proper edge-edge crossings (<= 16 points, duplicates included exactly as the reference
// keeps them), ordered around their centroid and fan-triangulated.  float32 arithmetic like the reference's local arrays;
// the area accumulates in double (the reference's `0.0` / `2.0` literals promote).  Matching the reference includes its
// quirk that two identical boxes give 1/3 (duplicate vertices), which the KITTI statistics were computed with.
#include <cmath>
#include <string>

#include "../../include/monocon_b200.h"
#include "common.cuh"

namespace mc {
namespace {

struct P2f { float x, y; };

__device__ __forceinline__ void corners_of(const float* b, P2f (&c)[4]) {
    const float ang = b[4], cs = cosf(ang), sn = sinf(ang);
    const float hx = (float)((double)b[2] / 2), hy = (float)((double)b[3] / 2);
    const float xs[4] = {-hx, -hx, hx, hx}, ys[4] = {-hy, hy, hy, -hy};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        c[i].x = __fadd_rn(__fadd_rn(__fmul_rn(cs, xs[i]), __fmul_rn(sn, ys[i])), b[0]);
        c[i].y = __fadd_rn(__fadd_rn(__fmul_rn(-sn, xs[i]), __fmul_rn(cs, ys[i])), b[1]);
    }
}

__device__ __forceinline__ bool inside(float px, float py, const P2f (&q)[4]) {
    const float ab0 = __fsub_rn(q[1].x, q[0].x), ab1 = __fsub_rn(q[1].y, q[0].y);
    const float ad0 = __fsub_rn(q[3].x, q[0].x), ad1 = __fsub_rn(q[3].y, q[0].y);
    const float ap0 = __fsub_rn(px, q[0].x), ap1 = __fsub_rn(py, q[0].y);
    const float abab = __fadd_rn(__fmul_rn(ab0, ab0), __fmul_rn(ab1, ab1)), abap = __fadd_rn(__fmul_rn(ab0, ap0), __fmul_rn(ab1, ap1));
    const float adad = __fadd_rn(__fmul_rn(ad0, ad0), __fmul_rn(ad1, ad1)), adap = __fadd_rn(__fmul_rn(ad0, ap0), __fmul_rn(ad1, ap1));
    return abab >= abap && abap >= 0.f && adad >= adap && adap >= 0.f;
}

__device__ __forceinline__ bool crossing(const P2f (&p1)[4], const P2f (&p2)[4], int i, int j, P2f& out) {
    const P2f A = p1[i], B = p1[(i + 1) & 3], C = p2[j], D = p2[(j + 1) & 3];
    const float BA0 = __fsub_rn(B.x, A.x), BA1 = __fsub_rn(B.y, A.y);
    const float DA0 = __fsub_rn(D.x, A.x), CA0 = __fsub_rn(C.x, A.x), DA1 = __fsub_rn(D.y, A.y), CA1 = __fsub_rn(C.y, A.y);
    const bool acd = __fmul_rn(DA1, CA0) > __fmul_rn(CA1, DA0);
    const bool bcd = __fmul_rn(__fsub_rn(D.y, B.y), __fsub_rn(C.x, B.x)) > __fmul_rn(__fsub_rn(C.y, B.y), __fsub_rn(D.x, B.x));
    if (acd == bcd) return false;
    const bool abc = __fmul_rn(CA1, BA0) > __fmul_rn(BA1, CA0);
    const bool abd = __fmul_rn(DA1, BA0) > __fmul_rn(BA1, DA0);
    if (abc == abd) return false;
    const float DC0 = __fsub_rn(D.x, C.x), DC1 = __fsub_rn(D.y, C.y);
    const float ABBA = __fsub_rn(__fmul_rn(A.x, B.y), __fmul_rn(B.x, A.y));
    const float CDDC = __fsub_rn(__fmul_rn(C.x, D.y), __fmul_rn(D.x, C.y));
    const float DH = __fsub_rn(__fmul_rn(BA1, DC0), __fmul_rn(BA0, DC1));
    out.x = __fdiv_rn(__fsub_rn(__fmul_rn(ABBA, DC0), __fmul_rn(BA0, CDDC)), DH);
    out.y = __fdiv_rn(__fsub_rn(__fmul_rn(ABBA, DC1), __fmul_rn(BA1, CDDC)), DH);
    return true;
}

// intersection area of rbox1 and rbox2 ([cx, cy, dx, dy, angle])
__device__ double intersection_area(const float* r1, const float* r2) {
    P2f p1[4], p2[4], pts[16];
    corners_of(r1, p1);
    corners_of(r2, p2);
    int n = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (inside(p1[i].x, p1[i].y, p2)) pts[n++] = p1[i];
        if (inside(p2[i].x, p2[i].y, p1)) pts[n++] = p2[i];
    }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            P2f t;
            if (n < 16 && crossing(p1, p2, i, j, t)) pts[n++] = t;
        }
    if (n == 0) return 0.0;
    float cx = 0.f, cy = 0.f;
    for (int i = 0; i < n; ++i) { cx = __fadd_rn(cx, pts[i].x); cy = __fadd_rn(cy, pts[i].y); }
    cx = __fdiv_rn(cx, (float)n); cy = __fdiv_rn(cy, (float)n);
    float key[16];
    for (int i = 0; i < n; ++i) {
        float vx = __fsub_rn(pts[i].x, cx), vy = __fsub_rn(pts[i].y, cy);
        const float d = __fsqrt_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)));
        vx = __fdiv_rn(vx, d); vy = __fdiv_rn(vy, d);
        key[i] = vy < 0.f ? __fsub_rn(-2.f, vx) : vx;
    }
    for (int i = 1; i < n; ++i) {                    // stable insertion sort, ascending key (the reference's order)
        if (key[i - 1] > key[i]) {
            const float t = key[i];
            const P2f tp = pts[i];
            int j = i;
            while (j > 0 && key[j - 1] > t) { key[j] = key[j - 1]; pts[j] = pts[j - 1]; --j; }
            key[j] = t; pts[j] = tp;
        }
    }
    double area = 0.0;
    for (int i = 0; i + 2 < n; ++i) {
        const P2f a = pts[0], b = pts[i + 1], c = pts[i + 2];
        const float cr = __fsub_rn(__fmul_rn(__fsub_rn(a.x, c.x), __fsub_rn(b.y, c.y)), __fmul_rn(__fsub_rn(a.y, c.y), __fsub_rn(b.x, c.x)));
        area += fabs((double)cr / 2.0);
    }
    return area;
}

// boxes (N,5), qboxes (K,5) -> out (N,K).  The reference evaluates devRotateIoUEval(query, box): criterion 0 divides by the
// query's area, 1 by the box's, 2 returns the intersection area.
__global__ void rotate_iou_kernel(const float* __restrict__ boxes, const float* __restrict__ qboxes, int N, int K, int criterion,
                                  float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)N * K) return;
    const int n = (int)(i / K), k = (int)(i % K);
    const float* b = boxes + (long long)n * 5;
    const float* q = qboxes + (long long)k * 5;
    const float area1 = __fmul_rn(q[2], q[3]), area2 = __fmul_rn(b[2], b[3]);
    const double inter = intersection_area(q, b);
    double r;
    if (criterion == -1) r = inter / ((double)__fadd_rn(area1, area2) - inter);
    else if (criterion == 0) r = inter / (double)area1;
    else if (criterion == 1) r = inter / (double)area2;
    else r = inter;
    out[i] = (float)r;
}

// camera boxes (N,7) / (K,7) [x, y, z, l, h, w, ry] in float64 like the reference's numpy arrays; BEV pass on
// [x, z, l, w, ry] cast to float32, then the height overlap (eval.py:128-157)
__global__ void box3d_overlap_kernel(const double* __restrict__ boxes, const double* __restrict__ qboxes, int N, int K, int criterion,
                                     float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)N * K) return;
    const int n = (int)(i / K), k = (int)(i % K);
    const double* b = boxes + (long long)n * 7;
    const double* q = qboxes + (long long)k * 7;
    const float bb[5] = {(float)b[0], (float)b[2], (float)b[3], (float)b[5], (float)b[6]};
    const float qq[5] = {(float)q[0], (float)q[2], (float)q[3], (float)q[5], (float)q[6]};
    const float rinc = (float)intersection_area(qq, bb);
    float r = rinc;
    if (rinc > 0.f) {
        const double iw = fmin(b[1], q[1]) - fmax(b[1] - b[4], q[1] - q[4]);
        if (iw > 0) {
            const double area1 = b[3] * b[4] * b[5], area2 = q[3] * q[4] * q[5];
            const double inc = iw * (double)rinc;
            const double ua = criterion == -1 ? (area1 + area2 - inc) : criterion == 0 ? area1 : criterion == 1 ? area2 : inc;
            r = (float)(inc / ua);
        } else {
            r = 0.f;
        }
    }
    out[i] = r;
}

thread_local std::string g_eval_error;

}  // namespace
}  // namespace mc

using namespace mc;

extern "C" {

const char* mc_eval_last_error(void) { return g_eval_error.c_str(); }

int mc_rotate_iou(int device, const float* boxes, const float* qboxes, int N, int K, int criterion, float* out, void* stream) {
    try {
        MC_CUDA(cudaSetDevice(device));
        MC_CHECK(N >= 0 && K >= 0 && (N == 0 || K == 0 || (boxes && qboxes && out)), "arguments");
        if (N == 0 || K == 0) return 0;
        const long long total = (long long)N * K;
        rotate_iou_kernel<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(boxes, qboxes, N, K, criterion, out);
        MC_CUDA(cudaGetLastError());
        return 0;
    } catch (const std::exception& e) {
        g_eval_error = e.what();
        return 1;
    }
}

int mc_box3d_overlap(int device, const double* boxes, const double* qboxes, int N, int K, int criterion, float* out, void* stream) {
    try {
        MC_CUDA(cudaSetDevice(device));
        MC_CHECK(N >= 0 && K >= 0 && (N == 0 || K == 0 || (boxes && qboxes && out)), "arguments");
        if (N == 0 || K == 0) return 0;
        const long long total = (long long)N * K;
        box3d_overlap_kernel<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(boxes, qboxes, N, K, criterion, out);
        MC_CUDA(cudaGetLastError());
        return 0;
    } catch (const std::exception& e) {
        g_eval_error = e.what();
        return 1;
    }
}

}  // extern "C"
