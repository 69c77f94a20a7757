"""CPU: the oracle's restatement of the first-tile filter (oracle/gact_oracle.c: gact_filter, gact_slope_filter ==
filter.cpp:28-289) against the committed golden fixture generated from the COMPILED reference
(tests/golden/make_golden.py: gen_filter) and, when oracle/_ref is built, against the reference's filter_body live."""
import os

import numpy as np
import pytest

import oracle
from darwin_b200 import abi
from conftest import GOLDEN


def locations_from(cands, read_num, res, port, slope=0.05):
    """filter_body's tail (filter.cpp:87-124 / :190-223): keep score+overlap passes per strand, slope-filter, fw then rc."""
    out = []
    for strand in (0, 1):
        m = np.nonzero((cands["strand"] == strand) & ((res["flags"] & 3) == 3))[0]
        order = port.slope_filter(read_num[m], res["score"][m], res["reference_pos"][m], res["query_pos"][m], slope)
        out.append(m[order])
    return np.concatenate(out)


def same_locations(idx, cands, read_num, res, anchors):
    return (len(idx) == len(anchors) and np.array_equal(res["score"][idx], anchors["score"]) and
            np.array_equal(res["reference_pos"][idx], anchors["reference_pos"]) and
            np.array_equal(res["query_pos"][idx], anchors["query_pos"]) and
            np.array_equal(read_num[idx], anchors["read_num"]) and np.array_equal(cands["strand"][idx], anchors["strand"]))


@pytest.fixture(scope="module")
def golden_filter():
    return np.load(os.path.join(GOLDEN, "filter_v1.npz"))


def test_port_filter_matches_golden(golden_filter):
    g = golden_filter
    port = oracle.port(abi.Scoring.from_values(*[int(x) for x in g["scoring"]]))
    res = port.filter(g["arena"], g["cands"])
    assert np.array_equal(res, g["port_res"])
    idx = locations_from(g["cands"], g["cand_read_num"], res, port)
    assert len(g["anchors"]) > 30 and same_locations(idx, g["cands"], g["cand_read_num"], res, g["anchors"])
    # the fixture exercises both strands and the slope filter
    c = g["cands"]
    assert (c["strand"] == 0).any() and (c["strand"] == 1).any() and len(idx) < int(((res["flags"] & 3) == 3).sum())


def check_custom(g, run_filter, port):
    """Hand-made candidates (low scores, tiles clamped at chromosome / read ends, a chromosome and reads shorter than the
    tile) that the compiled reference's filter_body let through with permissive thresholds: every candidate's
    (score, reference_pos, query_pos) must match, for first_tile_size 128 and 96."""
    c, rn = g["custom_cands"], g["custom_read_num"]
    assert ((c["hit"] + 128 >= c["chr_start"] + c["chr_len"]).any() and (c["offset"] + 128 >= c["read_len"]).any() and
            (c["read_len"] < 128).any() and (c["chr_len"] <= 128).any())
    for fts, key in ((128, "custom_locations"), (96, "custom_locations96")):
        res = run_filter(c, fts, 0, 0)
        assert ((res["flags"] & 3) == 3).all()
        idx = locations_from(c, rn, res, port, slope=-1.0)                 # nothing dropped: the reference's sort order only
        assert len(idx) == len(c) and same_locations(idx, c, rn, res, g[key])
    assert (g["custom_locations"]["score"] < 60).sum() > 100


def test_port_filter_custom_candidates(golden_filter):
    g = golden_filter
    port = oracle.port(abi.Scoring.from_values(*[int(x) for x in g["scoring"]]))
    check_custom(g, lambda c, fts, thr, ovl: port.filter(g["arena"], c, fts, thr, ovl), port)


def test_slope_filter_rules():
    port = oracle.port(abi.Scoring.from_values())
    # read 0: second location on the same diagonal (slope 1) is dropped, third (off-diagonal) kept; read 1 untouched;
    # equal q (division by zero -> inf) is kept; order = read asc, score desc (filter.cpp:232-234)
    rn = np.array([0, 0, 0, 1, 1], np.int32)
    sc = np.array([100, 90, 80, 70, 75], np.int32)
    rp = np.array([1000, 2000, 3000, 500, 900], np.uint32)
    qp = np.array([100, 1100, 1500, 50, 50], np.uint32)
    assert list(port.slope_filter(rn, sc, rp, qp)) == [0, 2, 4, 3]
    assert list(port.slope_filter(rn[:0], sc[:0], rp[:0], qp[:0])) == []


@pytest.mark.skipif(not oracle.have_reference(), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed", [21, 22])
def test_port_filter_matches_reference_live(seed):
    import sys
    sys.path.insert(0, GOLDEN)
    import make_golden
    ref, n_reads = make_golden.filter_case(seed, 36)
    cands, rn = ref.seed(0, n_reads)
    anchors, _ = ref.filter_last()
    port = oracle.port(abi.Scoring.from_values())
    res = port.filter(ref.arena().copy(), cands)
    idx = locations_from(cands, rn, res, port)
    assert len(cands) > 20 and same_locations(idx, cands, rn, res, anchors)
