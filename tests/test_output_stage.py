"""CPU: the host-side output stage of the library (darwin_gpu_cigar, darwin_gpu_sam_select; no device work) against a
restatement of printer_body (software/printer.cpp:15-47 ordering + overlap suppression, :236-310 CIGAR construction) that
walks two gapped strings the way the reference does."""
import numpy as np

import darwin_b200
from darwin_b200 import abi


def reference_cigar(ref_str, qry_str, query_start, query_end, query_len):
    """printer.cpp:236-310 on gapped strings."""
    out, prev, num = [], "Z", 0
    if query_start > 0:
        out.append("%dS" % query_start)
    for r, q in zip(ref_str, qry_str):
        op = "I" if r == "-" else ("D" if q == "-" else "M")
        if op == prev:
            num += 1
        else:
            if num > 0:
                out.append("%d%s" % (num, prev))
            num = 1
        prev = op
    if num > 0:
        out.append("%d%s" % (num, prev))
    tail = query_len - query_end - 1
    if tail > 0:
        out.append("%dS" % tail)
    return "".join(out) if out else "*"


def gapped(ops):
    ref = "".join("-" if o == abi.OP_I else "A" for o in ops)
    qry = "".join("-" if o == abi.OP_D else "C" for o in ops)
    return ref, qry


def test_cigar_matches_the_printer_rules():
    rng = np.random.default_rng(1)
    pool, rows, want = [], [], []
    cases = [[], [3], [1], [2], [3] * 1000, [1, 1, 2, 2, 3, 3, 1], [2] * 7 + [3] * 12 + [1] * 300]
    for _ in range(60):
        n = int(rng.integers(1, 400))
        cases.append(rng.choice([1, 2, 3], n, p=[0.1, 0.1, 0.8]).tolist())
    for k, ops in enumerate(cases):
        qs = 0 if k % 3 == 0 else int(rng.integers(1, 5000))
        qlen_extra = 0 if k % 4 == 0 else int(rng.integers(1, 70000))
        consumed_q = sum(1 for o in ops if o != abi.OP_D)
        qe = qs + max(consumed_q, 1) - 1
        r = np.zeros(1, abi.ALN_RES)
        r["ops_offset"], r["n_ops"], r["query_start_offset"], r["query_end_offset"] = len(pool), len(ops), qs, qe
        r["flags"] = abi.ALN_EMITTED
        pool += ops
        rows.append((r, qe + 1 + qlen_extra))
        want.append(reference_cigar(*gapped(ops), qs, qe, qe + 1 + qlen_extra))
    pool = np.array(pool + [0], np.uint8)
    for (r, qlen), w in zip(rows, want):
        assert darwin_b200.cigar(r, pool, qlen) == w
    # nothing aligned, no clips: "*"
    r = np.zeros(1, abi.ALN_RES)
    assert darwin_b200.cigar(r, pool, 1) == "*"


def reference_select(read_num, score, qs, qe):
    """printer.cpp:15-47: stable sort by (read, score desc), then drop alignments mostly covered by a better one."""
    order = sorted(range(len(read_num)), key=lambda i: (read_num[i], -score[i]))
    keep = [True] * len(order)
    for a in range(len(order)):
        if not keep[a]:
            continue
        s1, e1 = qs[order[a]], qe[order[a]]
        for b in range(a + 1, len(order)):
            if not keep[b]:
                continue
            if read_num[order[b]] != read_num[order[a]]:
                break
            s2, e2 = qs[order[b]], qe[order[b]]
            s, e = max(s1, s2), min(e1, e2)
            overlap = e - s if e > s else 0
            if (2 * overlap) & 0xFFFFFFFF > (e2 - s2) & 0xFFFFFFFF:
                keep[b] = False
    return order, keep


def test_sam_select_matches_the_printer_rules():
    rng = np.random.default_rng(2)
    n = 400
    anchors, res = np.zeros(n, abi.ANCHOR), np.zeros(n, abi.ALN_RES)
    anchors["read_num"] = rng.integers(0, 40, n)
    res["score"] = rng.integers(-50, 300, n) // 10 * 10                      # many ties: stability matters
    qs = rng.integers(0, 9000, n)
    res["query_start_offset"] = qs
    res["query_end_offset"] = qs + rng.integers(0, 4000, n)
    res["flags"] = np.where(rng.random(n) < 0.85, abi.ALN_EMITTED, 0)
    order, keep = darwin_b200.sam_select(anchors, res)
    em = np.flatnonzero(res["flags"] & 1)
    w_order, w_keep = reference_select(anchors["read_num"][em].tolist(), res["score"][em].tolist(),
                                       res["query_start_offset"][em].tolist(), res["query_end_offset"][em].tolist())
    assert order.tolist() == em[w_order].tolist()
    assert keep.tolist() == w_keep
    assert 0 < keep.sum() < len(keep)
    o0, k0 = darwin_b200.sam_select(anchors[:0], res[:0])
    assert len(o0) == 0 and len(k0) == 0
