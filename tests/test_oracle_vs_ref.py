"""CPU: fresh random inputs through the compiled reference (oracle/_ref) and the restatement.  Skipped only
when oracle/_ref has not been built (no /root/reference and no prebuilt library)."""
import numpy as np
import pytest

import oracle
from darwin_b200 import abi, synth
from conftest import tiles_equal

pytestmark = pytest.mark.skipif(not oracle.have_reference(), reason="oracle/_ref not built")


def _random_tiles(seed, n, maxsize):
    rng = np.random.default_rng(seed)
    arena, req, pos = [np.full(64, ord("N"), np.uint8)], np.zeros(n, abi.TILE_REQ), 64
    for k in range(n):
        R = int(rng.integers(1, maxsize + 1))
        r = synth.random_seq(rng, R)
        q = synth.mutate(rng, r, 0.1 * rng.random(), 0.08 * rng.random(), 0.08 * rng.random(), 0.002, (1, 30))
        if len(q) == 0:
            q = synth.random_seq(rng, 3)
        for name, s in (("ref", r), ("query", q)):
            req[k][name + "_bases_start_addr"], req[k][name + "_size"] = pos, len(s)
            arena.append(s)
            pos += len(s)
        req[k]["max_tb_steps"] = int(rng.choice([2 * maxsize, 48]))
        req[k]["align_fields"] = int(rng.choice([1, 21, 7, 19, 0, 6]))
        req[k]["index"] = k % 256
    return np.concatenate(arena + [np.full(64, ord("N"), np.uint8)]), req


@pytest.mark.parametrize("vals", [(2, -6, -1, -4, -2, -25, -1), (1, -1, 0, -1, -1, -1, -1), (2, -3, -1, -3, -2, -8, -1),
                                  (5, -4, -1, -10, -1, -1000, -1)])
def test_random_tiles_port_equals_reference(vals):
    sc = abi.Scoring.from_values(*vals)
    ref = oracle.reference("patched")
    ref.set_scoring(sc)
    port = oracle.port(sc)
    arena, req = _random_tiles(hash(vals) & 0xFFFF, 150, 260)
    rres, rtb = ref.tiles(arena, req, 1, tb_words_per_req=80)
    for rule in (oracle.Port.STRIPED, oracle.Port.STREAM):
        pres, ptb, _ = port.tiles(arena, req, 1, rule, tb_words_per_req=80)
        assert tiles_equal(rres, rtb, pres, ptb) == []


def test_general_matrix_port_equals_reference():
    """Non-uniform substitution matrix (transitions cheaper than transversions)."""
    m = dict(AA=3, AC=-5, AG=-2, AT=-5, CC=3, CG=-5, CT=-2, GG=3, GT=-5, TT=4)
    sc = abi.Scoring.from_values(matrix=m, sub_n=-1, go=-5, ge=-2, lgo=-20, lge=-1)
    ref = oracle.reference("patched")
    ref.set_scoring(sc)
    port = oracle.port(sc)
    arena, req = _random_tiles(5, 120, 200)
    rres, rtb = ref.tiles(arena, req, 1, tb_words_per_req=80)
    pres, ptb, _ = port.tiles(arena, req, 1, oracle.Port.STREAM, tb_words_per_req=80)
    assert tiles_equal(rres, rtb, pres, ptb) == []


SAMPLE_REF = "/root/reference/software/data/sample_ref.fa"


@pytest.mark.skipif(not __import__("os").path.exists(SAMPLE_REF), reason="reference data not present")
def test_config1_sample_reference_plumbing():
    """BASELINE.json configs[0]: the reference's own sample reference (sacCer3 chrI) with stock params.cfg; the reads
    file is missing from the reference tree (.MISSING_LARGE_BLOBS), so reads are simulated (seeded, 15 % error).
    Reference pipeline (D-SOFT, filter, extender_body) vs the restatement on the same anchors."""
    from conftest import alignments_equal, ALN_FIELDS
    seq = b"".join(l.strip() for l in open(SAMPLE_REF, "rb") if not l.startswith(b">"))
    ref = oracle.reference("patched")
    ref.load_cfg("/root/reference/software/params.cfg", 0)
    ref.reset_arena()
    ref.add_chr("chrI", seq, True)
    ref.build_index()
    rng = np.random.default_rng(1)
    g = np.frombuffer(seq, np.uint8)
    for k in range(6):
        L = int(rng.integers(6000, 9000))
        p = int(rng.integers(0, len(g) - L))
        r = synth.mutate_fast(rng, np.char.upper(g[p:p + L].view("S1")).view(np.uint8), 0.05, 0.05, 0.05)
        ref.add_read("r%d" % k, np.ascontiguousarray(synth.revcomp(r) if k % 2 else r).tobytes())
    A, H, hb = [], [], 0
    for k in range(6):
        a, h = ref.seed_filter(k, 1)
        a = a.copy()
        a["left_hits_off"] += hb
        a["right_hits_off"] += hb
        hb += len(h)
        A.append(a)
        H.append(h)
    anchors, hits = np.concatenate(A), np.concatenate(H)
    assert len(anchors) >= 6
    want, wops = ref.extend(anchors, hits)
    port = oracle.port(abi.Scoring.from_values())
    for rule in (oracle.Port.STREAM, oracle.Port.CLEAN):
        got, gops = port.extend(ref.arena(), abi.ExtendParams(384, 64, 0, 0), anchors, hits, rule)
        assert alignments_equal(want, wops, got, gops, ALN_FIELDS) == []


@pytest.mark.parametrize("seed,scheme,T,O,ovl", [(41, (2, -6, -1, -4, -2, -25, -1), 128, 32, 0),
                                                 (42, (2, -3, -1, -3, -2, -8, -1), 512, 64, 0),
                                                 (43, (2, -6, -1, -4, -2, -25, -1), 200, 150, 1),
                                                 (44, (1, -1, 0, -2, -1, -4, 0), 384, 64, 0)])
def test_extend_live_port_equals_reference(seed, scheme, T, O, ovl):
    """extender_body of the compiled reference on fresh reads (structural insertions / deletions that force large tiles,
    both strands, reads hanging over the chromosome ends) with anchors from its own D-SOFT + filter, against the
    restatement: tile sizes and scoring schemes the committed fixtures do not hold."""
    from conftest import alignments_equal
    rng = np.random.default_rng(seed)
    genome = synth.random_seq(rng, 60000)
    sc = abi.Scoring.from_values(*scheme)
    ref = oracle.reference("patched")
    ref.set_scoring(sc)
    ref.set_dsoft_defaults()
    ref.set_extend(T, O, 2, ovl)
    ref.reset_arena()
    ref.add_chr("chrL", genome.tobytes(), True)
    ref.build_index()
    n_reads = 6
    for k in range(n_reads):
        L = int(rng.integers(2500, 4500))
        p = int(rng.integers(0, len(genome) - L)) if k < 4 else (0 if k == 4 else len(genome) - L)
        src = genome[p:p + L]
        if k % 3 == 1:
            src = np.concatenate([src[:L // 2], synth.random_seq(rng, 400), src[L // 2:]])
        elif k % 3 == 2:
            src = np.concatenate([src[:L // 3], src[L // 3 + 500:]])
        r = synth.mutate(rng, src, 0.04, 0.04, 0.04, indel_run=(3, 40) if k % 2 else None)
        if k >= 4:                                                       # overhang beyond the chromosome end
            r = np.concatenate([synth.random_seq(rng, 300), r]) if k == 4 else np.concatenate([r, synth.random_seq(rng, 300)])
        if k % 2 == 0:
            r = synth.revcomp(r)
        ref.add_read("r%d" % k, np.ascontiguousarray(r).tobytes())
    A, H, hb = [], [], 0
    for k in range(n_reads):
        a, h = ref.seed_filter(k, 1)
        a = a.copy()
        a["left_hits_off"] += hb
        a["right_hits_off"] += hb
        hb += len(h)
        A.append(a)
        H.append(h)
    anchors, hits = np.concatenate(A), np.concatenate(H)
    assert len(anchors) >= n_reads - 1
    want, want_ops = ref.extend(anchors, hits)
    got, got_ops = oracle.port(sc).extend(ref.arena().copy(), abi.ExtendParams(T, O, ovl, 0), anchors, hits, oracle.Port.STREAM)
    assert alignments_equal(want, want_ops, got, got_ops) == []
    assert int((want["flags"] & 1).sum()) >= 1
