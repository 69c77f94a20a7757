"""GPU: anchors produced by the reference's own D-SOFT + first-tile filter (oracle/_ref) on a synthetic genome with
repeats and reads with structural indels -- spurious anchors, stalls and 1984x960 / 960x1984 large tiles included --
extended on the GPU and by the reference's extender_body (patched flavour).  Every alignment must be identical."""
import numpy as np
import pytest

import oracle
from darwin_b200 import abi, synth
from conftest import alignments_equal, ALN_FIELDS

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not oracle.have_reference(), reason="oracle/_ref not built")]


def build_case(seed, n_reads, genome_len=400000):
    rng = np.random.default_rng(seed)
    genome = synth.random_seq(rng, genome_len)
    # a few 3 kbp repeats (diverged copies) so that D-SOFT proposes secondary / spurious anchors
    for _ in range(6):
        a, b = int(rng.integers(0, genome_len - 4000)), int(rng.integers(0, genome_len - 4000))
        rep = synth.mutate_fast(rng, genome[a:a + 3000], 0.03, 0.01, 0.01)
        genome[b:b + len(rep)] = rep[:min(len(rep), genome_len - b)]
    ref = oracle.reference("patched")
    ref.set_scoring(abi.Scoring.from_values())
    ref.set_dsoft_defaults()
    ref.reset_arena()
    ref.add_chr("chrS", genome.tobytes(), True)
    ref.build_index()
    for k in range(n_reads):
        L = int(rng.integers(6000, 10000))
        p = int(rng.integers(0, genome_len - L))
        src = genome[p:p + L]
        if k % 5 == 1:
            src = np.concatenate([src[:L // 2], synth.random_seq(rng, int(rng.integers(200, 700))), src[L // 2:]])
        elif k % 5 == 2:
            cut = int(rng.integers(300, 900))
            src = np.concatenate([src[:L // 3], src[L // 3 + cut:]])
        r = synth.mutate_fast(rng, src, 0.05, 0.05, 0.05)
        if k % 2:
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
    return ref, np.concatenate(A), np.concatenate(H)


@pytest.mark.parametrize("T,O,ovl", [(384, 64, 0), (320, 128, 0), (512, 64, 0), (256, 64, 1)])
def test_dsoft_anchors_extend_identically(gpu, T, O, ovl):
    ref, anchors, hits = build_case(1000 + T, 60)
    ref.set_extend(T, O, 2, ovl)
    want_res, want_ops = ref.extend(anchors, hits)
    arena = ref.arena().copy()
    p = gpu(len(arena), abi.Scoring.from_values())
    p.InitializeReferenceMemory(0, arena)
    st0 = p.stats()
    res, ops = p.extender_body(anchors, hits, T, O, ovl)
    st1 = p.stats()
    assert alignments_equal(want_res, want_ops, res, ops, ALN_FIELDS) == []
    assert len(anchors) >= 60 and int((res["flags"] & 1).sum()) >= 50
    if T == 384:
        assert int(res["n_large_tiles"].sum()) > 0            # the large-tile fallback was exercised ...
        assert st1.tiles_scoreonly - st0.tiles_scoreonly > 0  # ... and some of those tiles ended in the score-only pre-pass
    assert st1.tiles_fast - st0.tiles_fast > 0
    p.close()
