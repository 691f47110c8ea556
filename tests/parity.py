"""Parity rule of BASELINE.json north_star, as a checker.

  * top-k doc ids and order identical to the reference, ties broken by doc id;
  * an id difference is allowed only where the reference scores lie within 1e-5 relative of the
    k-th score (or, for a pure order swap, of each other);
  * scores agree within 1e-5 relative.

`strict=True` demands bit-equal ids and scores (used against the oracle's `bm25` mode, whose float
summation order — query term order — is the one the CUDA kernel reproduces).
"""
from __future__ import annotations

import numpy as np

RTOL = 1e-5


def check_query(ref, got, n_ref: int, n_got: int, rtol: float = RTOL, strict: bool = False) -> str | None:
    """ref/got: structured hit rows of one query.  Returns None if parity holds, else a message."""
    if n_ref != n_got:
        return f"count {n_got} != reference {n_ref}"
    if n_ref == 0:
        return None
    r, g = ref[:n_ref], got[:n_got]
    if strict:
        if not np.array_equal(r["doc_id"], g["doc_id"]):
            return f"ids differ: {g['doc_id'].tolist()} vs {r['doc_id'].tolist()}"
        if not np.array_equal(r["score"].view(np.uint32), g["score"].view(np.uint32)):
            return f"scores not bit-equal: {g['score'].tolist()} vs {r['score'].tolist()}"
        return None
    # own order must be (score desc, doc asc)
    for i in range(1, n_got):
        if g["score"][i] > g["score"][i - 1] or (g["score"][i] == g["score"][i - 1] and g["doc_id"][i] < g["doc_id"][i - 1]
                                                  and g["segment_ord"][i] == g["segment_ord"][i - 1]):
            return f"result not ordered at {i}"
    # rank-wise scores within rtol
    tol = rtol * np.abs(r["score"].astype(np.float64))
    if not np.all(np.abs(g["score"].astype(np.float64) - r["score"].astype(np.float64)) <= tol + 1e-30):
        return f"scores differ beyond {rtol}: {g['score'].tolist()} vs {r['score'].tolist()}"
    kth = float(r["score"][n_ref - 1])
    ref_pos = {(int(s), int(d)): i for i, (s, d) in enumerate(zip(r["segment_ord"], r["doc_id"]))}
    for i in range(n_ref):
        if r["doc_id"][i] == g["doc_id"][i] and r["segment_ord"][i] == g["segment_ord"][i]:
            continue
        si = float(r["score"][i])
        if abs(si - kth) <= rtol * abs(kth):
            continue  # boundary tie zone
        j = ref_pos.get((int(g["segment_ord"][i]), int(g["doc_id"][i])))
        if j is not None and abs(float(r["score"][j]) - si) <= rtol * abs(si):
            continue  # swap among near-equal scores
        return f"id mismatch at rank {i}: got {int(g['doc_id'][i])}, reference {int(r['doc_id'][i])}"
    return None


def assert_parity(ref_hits, ref_counts, got_hits, got_counts, rtol: float = RTOL, strict: bool = False):
    bad = []
    for q in range(len(ref_counts)):
        msg = check_query(ref_hits[q], got_hits[q], int(ref_counts[q]), int(got_counts[q]), rtol, strict)
        if msg:
            bad.append((q, msg))
    assert not bad, f"{len(bad)} of {len(ref_counts)} queries break parity; first: {bad[:3]}"


def parity_report(ref_hits, ref_counts, got_hits, got_counts, rtol: float = RTOL) -> dict:
    n = len(ref_counts)
    strict_ok = sum(check_query(ref_hits[q], got_hits[q], int(ref_counts[q]), int(got_counts[q]), rtol, True) is None for q in range(n))
    rule_ok = sum(check_query(ref_hits[q], got_hits[q], int(ref_counts[q]), int(got_counts[q]), rtol, False) is None for q in range(n))
    return {"queries": n, "bit_exact": strict_ok, "within_rule": rule_ok}
