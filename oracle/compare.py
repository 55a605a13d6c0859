"""Comparison helpers of the parity checks (test infrastructure, like the rest of oracle/: imported by tests/, by
__graft_entry__.smoke() and by bench.py's parity gate -- never by the product)."""
import numpy as np


def rel_to_max(a, b) -> float:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(1e-30, np.abs(b).max()))


def rel_l2(a, b) -> float:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(1e-30, np.linalg.norm(b)))


def topk_matches(got_inds, got_labels, ref_inds, ref_labels, ref_scores, n_cells: int, near_tie: float = 0.0) -> bool:
    """Top-k agreement of (index, class) pairs.  got_*: (B, K); ref_*: (B, >= K) in the reference's order with the reference's
    raw heat-map scores (one extra reference entry covers the K / K+1 boundary).
    near_tie = 0: identical, in order.  near_tie > 0: consecutive reference entries whose scores differ by less than near_tie
    form a group; inside a group any order is accepted (and a group that straddles the K-th place may contribute any of its
    members), everything else must be identical."""
    got = np.asarray(got_labels).astype(np.int64) * n_cells + np.asarray(got_inds).astype(np.int64)
    ref = np.asarray(ref_labels).astype(np.int64) * n_cells + np.asarray(ref_inds).astype(np.int64)
    scores = np.asarray(ref_scores, dtype=np.float64)
    K, n = got.shape[1], ref.shape[1]
    for b in range(ref.shape[0]):
        if near_tie <= 0:
            if not np.array_equal(got[b], ref[b][:K]):
                return False
            continue
        i = 0
        while i < K:
            j = i
            while j + 1 < n and scores[b][j] - scores[b][j + 1] < near_tie:
                j += 1
            ref_group = set(ref[b][i:j + 1].tolist())
            got_group = set(got[b][i:min(j + 1, K)].tolist())
            if j + 1 <= K:
                if got_group != ref_group:
                    return False
            elif not got_group <= ref_group:
                return False
            i = j + 1
    return True


def count_topk_differences(got_inds, got_labels, ref_inds, ref_labels, n_cells: int):
    """(positions that differ, size of the symmetric set difference) over all images -- for reporting."""
    got = np.asarray(got_labels).astype(np.int64) * n_cells + np.asarray(got_inds).astype(np.int64)
    ref = (np.asarray(ref_labels).astype(np.int64) * n_cells + np.asarray(ref_inds).astype(np.int64))[:, :got.shape[1]]
    pos = int((got != ref).sum())
    sym = sum(len(set(g.tolist()) ^ set(r.tolist())) for g, r in zip(got, ref))
    return pos, sym
