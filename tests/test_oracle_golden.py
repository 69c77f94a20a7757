"""CPU: the oracle restatement (oracle/gact_oracle.c) against the committed golden vectors, which were
produced by the compiled reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest

import oracle
from darwin_b200 import abi
from conftest import tiles_equal, alignments_equal

SCHEMES = ("stock", "tie", "s2", "s3")


@pytest.mark.parametrize("scheme", SCHEMES)
@pytest.mark.parametrize("rule", [oracle.Port.STRIPED, oracle.Port.STREAM])
def test_tiles_exact_rules_match_reference(golden_tiles, scheme, rule):
    g = golden_tiles
    port = oracle.port(abi.Scoring.from_values(*g[scheme + "_scoring"].tolist()))
    req = g[scheme + "_req"]
    if rule == oracle.Port.STRIPED:          # the literal emulation is slow on the two 1984x960 tiles: keep one
        keep = np.flatnonzero((req["ref_size"].astype(int) * req["query_size"]) < 500000)
        keep = np.concatenate([keep, np.flatnonzero((req["ref_size"].astype(int) * req["query_size"]) >= 500000)[:1]])
    else:
        keep = np.arange(len(req))
    res, tb, _ = port.tiles(g[scheme + "_arena"], req[keep], 1, rule, tb_words_per_req=260)
    assert tiles_equal(g[scheme + "_res"][keep], g[scheme + "_tb"][keep], res, tb) == []


@pytest.mark.parametrize("scheme", SCHEMES)
def test_tiles_clean_rule_exact_when_unflagged(golden_tiles, scheme):
    """the textbook rule is bit-exact whenever the traceback never meets a long-gap candidate."""
    g = golden_tiles
    port = oracle.port(abi.Scoring.from_values(*g[scheme + "_scoring"].tolist()))
    res, tb, flags = port.tiles(g[scheme + "_arena"], g[scheme + "_req"], 1, oracle.Port.CLEAN, tb_words_per_req=260)
    bad = tiles_equal(g[scheme + "_res"], g[scheme + "_tb"], res, tb)
    assert all(flags[k] & 1 for k in bad)
    assert (flags & 1).sum() < len(flags)          # the flag is not trivially always set


@pytest.mark.parametrize("scheme", SCHEMES)
def test_tiles_score_only(golden_tiles, scheme):
    g = golden_tiles
    port = oracle.port(abi.Scoring.from_values(*g[scheme + "_scoring"].tolist()))
    res, _, _ = port.tiles(g[scheme + "_arena"], g[scheme + "_req"], 0, oracle.Port.STREAM, tb_words_per_req=1)
    assert np.array_equal(res, g[scheme + "_res_notb"])


@pytest.mark.parametrize("tag", ["T384_O64_ovl0", "T320_O128_ovl0", "T256_O64_ovl1"])
@pytest.mark.parametrize("rule", [oracle.Port.STREAM, oracle.Port.CLEAN])
def test_extend_matches_reference(golden_extend, tag, rule):
    g = golden_extend
    port = oracle.port(abi.Scoring.from_values(*g["scoring"].tolist()))
    T, O, ovl = [int(x.lstrip("TOovl")) for x in tag.split("_")]
    res, ops = port.extend(g["arena"], abi.ExtendParams(T, O, ovl, 0), g[tag + "_anchors"], g[tag + "_hits"], rule)
    assert alignments_equal(g[tag + "_res"], g[tag + "_ops"], res, ops) == []


def test_rtl_known_answer_scores():
    """RTL testbench vectors (RTL/GACT/test_data): +1/-1 scoring, gap open/extend -1, max-cell mode; long gaps
    disabled.  The software recurrence reproduces the RTL's 10 'Total score' values."""
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "rtl_kat.npz"))
    sc = abi.Scoring.from_values(1, -1, 0, -1, -1, -1000, -1)
    port = oracle.port(sc)
    arena, req = [], np.zeros(10, abi.TILE_REQ)
    pos = 0
    for k in range(10):
        r = np.frombuffer(str(g["refs"][k]).encode(), np.uint8)
        q = np.frombuffer(str(g["queries"][k]).encode(), np.uint8)
        req[k]["ref_bases_start_addr"], req[k]["ref_size"] = pos, len(r)
        pos += len(r)
        req[k]["query_bases_start_addr"], req[k]["query_size"] = pos, len(q)
        pos += len(q)
        req[k]["max_tb_steps"] = 640
        arena += [r, q]
    arena = np.concatenate(arena)
    for rule in (oracle.Port.STRIPED, oracle.Port.STREAM, oracle.Port.CLEAN):
        res, _, _ = port.tiles(arena, req, 1, rule)
        assert res["score"].tolist() == g["scores"].tolist()


# ---- the as-is half of the parity report on the CPU: restatement vs fixtures ----------------
@pytest.mark.parametrize("scheme", SCHEMES)
def test_tiles_asis_rule_on_fixture(golden_tiles, scheme):
    """Tiles on which the two builds of the reference disagree all walk through the long-insertion state."""
    import parity_cases as pc
    g = golden_tiles
    port = oracle.port(abi.Scoring.from_values(*g[scheme + "_scoring"].tolist()))
    _, _, flags = port.tiles(g[scheme + "_arena"], g[scheme + "_req"], 1, oracle.Port.STREAM, tb_words_per_req=260)
    c = pc.check_asis_rule((flags & 2) != 0, g[scheme + "_asis_same"], scheme)
    assert c["flagged"] < c["n"]


def test_config1_fixture_matches_restatement():
    """BASELINE.json configs[0] (sample reference, stock params.cfg): the committed fixture vs the restatement, and the
    as-is rule on its anchors."""
    import os
    import parity_cases as pc
    from conftest import GOLDEN, ALN_FIELDS
    g = np.load(os.path.join(GOLDEN, "config1_v1.npz"))
    port = oracle.port(abi.Scoring.from_values(*g["scoring"].tolist()))
    T, O, ovl = [int(x) for x in g["extend"]]
    res, ops = port.extend(g["arena"], abi.ExtendParams(T, O, ovl, 0), g["anchors"], g["hits"], oracle.Port.STREAM)
    assert alignments_equal(g["res"], g["ops"], res, ops, ALN_FIELDS) == []
    c = pc.check_asis_rule((res["flags"] & abi.ALN_LONG_INS_PATH) != 0, g["asis_same"], "config1")
    assert c["n"] == len(g["anchors"]) and 0 < c["asis_differs"] <= c["flagged"] < c["n"]
