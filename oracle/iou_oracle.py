"""CPU restatement of the reference's rotated-IoU evaluation kernels.  TEST INFRASTRUCTURE ONLY.

* ``rotate_iou``      -- ``rotate_iou_gpu_eval`` (engine/kitti_eval/rotate_iou.py:337-379; device code :19-277): BEV boxes
                         [cx, cy, dx, dy, angle] (clockwise-positive angle, camera frame), intersection of the two quadrilaterals
                         = their mutually contained corners + the edge-edge intersections, ordered around the centroid,
                         fan-triangulated.
* ``d3_box_overlap``  -- ``d3_box_overlap`` / ``d3_box_overlap_kernel`` (engine/kitti_eval/eval.py:128-164): camera boxes
                         [x, y, z, l, h, w, ry]; BEV intersection area x height overlap.

Pinned: tests/golden/iou.npz is produced by tests/golden/gen_iou_golden.py from the UNMODIFIED reference kernels run under
NUMBA_ENABLE_CUDASIM=1; tests/test_iou_oracle.py holds this file to it (<= 2e-6).  Plain Python loops: small cases only.
"""
from __future__ import annotations

import math

import numpy as np

F = np.float32


def _corners(box):
    """rbbox_to_corners (:189-211): clockwise corners rotated clockwise by `angle`."""
    cx, cy, dx, dy, ang = [F(v) for v in box]
    c, s = F(math.cos(ang)), F(math.sin(ang))
    xs = [F(-dx / 2), F(-dx / 2), F(dx / 2), F(dx / 2)]
    ys = [F(-dy / 2), F(dy / 2), F(dy / 2), F(-dy / 2)]
    return [(F(c * x + s * y + cx), F(-s * x + c * y + cy)) for x, y in zip(xs, ys)]


def _inside(px, py, q):
    """point_in_quadrilateral (:157-172): projections on the edges AB and AD."""
    ab = (q[1][0] - q[0][0], q[1][1] - q[0][1])
    ad = (q[3][0] - q[0][0], q[3][1] - q[0][1])
    ap = (px - q[0][0], py - q[0][1])
    abab, abap = ab[0] * ab[0] + ab[1] * ab[1], ab[0] * ap[0] + ab[1] * ap[1]
    adad, adap = ad[0] * ad[0] + ad[1] * ad[1], ad[0] * ap[0] + ad[1] * ap[1]
    return abab >= abap >= 0 and adad >= adap >= 0


def _segment_hit(p1, p2, i, j):
    """line_segment_intersection (:72-114): proper crossing of edge i of p1 with edge j of p2, by orientation tests."""
    A, B = p1[i], p1[(i + 1) % 4]
    C, D = p2[j], p2[(j + 1) % 4]
    BA0, BA1 = B[0] - A[0], B[1] - A[1]
    DA0, CA0, DA1, CA1 = D[0] - A[0], C[0] - A[0], D[1] - A[1], C[1] - A[1]
    acd = DA1 * CA0 > CA1 * DA0
    bcd = (D[1] - B[1]) * (C[0] - B[0]) > (C[1] - B[1]) * (D[0] - B[0])
    if acd == bcd:
        return None
    abc = CA1 * BA0 > BA1 * CA0
    abd = DA1 * BA0 > BA1 * DA0
    if abc == abd:
        return None
    DC0, DC1 = D[0] - C[0], D[1] - C[1]
    ABBA = A[0] * B[1] - B[0] * A[1]
    CDDC = C[0] * D[1] - D[0] * C[1]
    DH = BA1 * DC0 - BA0 * DC1
    return (F((ABBA * DC0 - BA0 * CDDC) / DH), F((ABBA * DC1 - BA1 * CDDC) / DH))


def _intersection_area(b1, b2) -> float:
    """inter (:214-237): quadrilateral_intersection + sort_vertex_in_convex_polygon + area."""
    p1, p2 = _corners(b1), _corners(b2)
    pts = []
    for i in range(4):
        if _inside(p1[i][0], p1[i][1], p2):
            pts.append(p1[i])
        if _inside(p2[i][0], p2[i][1], p1):
            pts.append(p2[i])
    for i in range(4):
        for j in range(4):
            h = _segment_hit(p1, p2, i, j)
            if h is not None:
                pts.append(h)
    n = len(pts)
    if n == 0:
        return 0.0
    cxm = F(sum(p[0] for p in pts) / F(n))
    cym = F(sum(p[1] for p in pts) / F(n))
    keys = []
    for p in pts:                                   # monotone proxy of the polar angle around the centroid (:41-51)
        vx, vy = F(p[0] - cxm), F(p[1] - cym)
        d = F(math.sqrt(vx * vx + vy * vy))
        vx, vy = F(vx / d), F(vy / d)
        keys.append(F(-2 - vx) if vy < 0 else vx)
    order = sorted(range(n), key=lambda k: keys[k])           # insertion sort in the reference: stable, ascending
    pts = [pts[k] for k in order]
    a = 0.0
    for i in range(n - 2):
        p, q, r = pts[0], pts[i + 1], pts[i + 2]
        a += abs(((p[0] - r[0]) * (q[1] - r[1]) - (p[1] - r[1]) * (q[0] - r[0])) / 2.0)
    return a


def rotate_iou(boxes: np.ndarray, query_boxes: np.ndarray, criterion: int = -1) -> np.ndarray:
    """(N,5), (K,5) -> (N,K) float32.  Note the reference's argument order inside the kernel: devRotateIoUEval(qbox, box),
    so criterion 0 divides by the QUERY box's area and 1 by the box's (:330-333, 240-263)."""
    boxes = np.asarray(boxes, F)
    query_boxes = np.asarray(query_boxes, F)
    out = np.zeros((len(boxes), len(query_boxes)), F)
    for n, b in enumerate(boxes):
        for k, q in enumerate(query_boxes):
            area1 = F(q[2] * q[3])
            area2 = F(b[2] * b[3])
            inter = _intersection_area(q, b)
            if criterion == -1:
                out[n, k] = inter / (area1 + area2 - inter)
            elif criterion == 0:
                out[n, k] = inter / area1
            elif criterion == 1:
                out[n, k] = inter / area2
            else:
                out[n, k] = inter
    return out


def d3_box_overlap(boxes: np.ndarray, qboxes: np.ndarray, criterion: int = -1) -> np.ndarray:
    """engine/kitti_eval/eval.py:128-164 (camera boxes [x, y, z, l, h, w, ry]; y is the bottom, h extends towards -y)."""
    boxes = np.asarray(boxes)
    qboxes = np.asarray(qboxes)
    rinc = rotate_iou(boxes[:, [0, 2, 3, 5, 6]], qboxes[:, [0, 2, 3, 5, 6]], 2)
    for i in range(len(boxes)):
        for j in range(len(qboxes)):
            if rinc[i, j] > 0:
                iw = min(boxes[i, 1], qboxes[j, 1]) - max(boxes[i, 1] - boxes[i, 4], qboxes[j, 1] - qboxes[j, 4])
                if iw > 0:
                    area1 = boxes[i, 3] * boxes[i, 4] * boxes[i, 5]
                    area2 = qboxes[j, 3] * qboxes[j, 4] * qboxes[j, 5]
                    inc = iw * rinc[i, j]
                    ua = (area1 + area2 - inc) if criterion == -1 else area1 if criterion == 0 else area2 if criterion == 1 else inc
                    rinc[i, j] = inc / ua
                else:
                    rinc[i, j] = 0.0
    return rinc
