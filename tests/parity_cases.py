"""Shared builders of the parity report (tests only): cases run through the COMPILED reference (oracle/_ref, both
flavours) whose results the GPU path -- and, on the CPU, the restatement -- are compared with.

The reference reads two vectors uninitialised (software/Processor.cpp:259-260, used :405-408, :444), so two
builds of the same source can disagree wherever a traceback walks through the long-insertion state.  The parity report
therefore has two halves:
  * the answer equals the PATCHED flavour everywhere (vectors initialised the way the following column does it);
  * the answer equals the AS-IS flavour wherever the long-insertion flag is clear, and every case where as-is and patched
    differ carries the flag (DARWIN_TILE_LONG_INS_PATH in DarwinTileRes.status, DARWIN_ALN_LONG_INS_PATH in DarwinAlnRes.flags).
"""
import numpy as np

import oracle
from darwin_b200 import abi, synth
from conftest import ALN_FIELDS

STOCK = (2, -6, -1, -4, -2, -25, -1)
SCHEMES = {"stock": STOCK, "tie": (1, -1, 0, -1, -1, -1, -1), "s2": (1, -1, 0, -2, -1, -4, 0), "s3": (2, -3, -1, -3, -2, -8, -1)}
ONT = (0.04, 0.03, 0.05)            # BASELINE.json configs[4]: 12 % error, sub / ins / del
PACBIO = (0.015, 0.09, 0.045)       # configs[2]: 15 % error


def indel_rich_tiles(seed, n, max_size=400, n_large=0):
    """Tiles whose paths cross long gaps (20-80 base indels) -- the inputs on which the lazy-F tie rule, the long-gap
    states and the reference's uninitialised bits matter.  Corner traceback, the four extension flag sets."""
    rng = np.random.default_rng(seed)
    parts, pos = [np.full(128, ord("N"), np.uint8)], 128
    req = np.zeros(n + n_large, abi.TILE_REQ)
    flagsets = [1, 1 | 4 | 16, 1 | 4 | 2, 1 | 16 | 2]
    for k in range(n + n_large):
        if k >= n:
            R, Q = (1984, 960) if k % 2 else (960, 1984)
        else:
            R = max_size if k % 3 == 0 else int(rng.integers(40, max_size + 1))
            Q = None
        r = synth.random_seq(rng, R)
        q = synth.mutate(rng, r, 0.04, 0.03, 0.03, indel_run=(int(rng.integers(1, 4)), 80))
        if Q is not None:
            q = np.concatenate([q, synth.random_seq(rng, max(0, Q - len(q)))])[:Q]
        elif k % 3 == 0:
            q = np.concatenate([q, synth.random_seq(rng, max(0, max_size - len(q)))])[:max_size]
        q = q[:1984] if len(q) else synth.random_seq(rng, 1)
        req[k]["ref_bases_start_addr"], req[k]["ref_size"] = pos, len(r)
        pos += len(r)
        req[k]["query_bases_start_addr"], req[k]["query_size"] = pos, len(q)
        pos += len(q)
        parts += [r, q]
        req[k]["max_tb_steps"] = 768 if k >= n else 2 * max(len(r), len(q))
        req[k]["align_fields"] = flagsets[k % 4]
        req[k]["index"] = k % 250
    parts.append(np.full(128, ord("N"), np.uint8))
    return np.concatenate(parts), req


def reference_tiles(scheme, arena, req, words):
    """(patched results, patched TB words, per-tile bool: as-is flavour identical)."""
    out = {}
    for fl in ("patched", "as-is"):
        ref = oracle.reference(fl)
        ref.set_scoring(abi.Scoring.from_values(*scheme))
        out[fl] = ref.tiles(arena, req, 1, tb_words_per_req=words)
    (rp, tp), (ra, ta) = out["patched"], out["as-is"]
    same = np.array([rp[k] == ra[k] and np.array_equal(tp[k, :(int(rp[k]["total_TB_pointers"]) + 31) // 32],
                                                       ta[k, :(int(ra[k]["total_TB_pointers"]) + 31) // 32]) for k in range(len(req))])
    return rp, tp, same


def simulated_reads(rng, genome, n_reads, read_len, err, structural=True):
    """Reads as BASELINE.json's configs describe them: uniform loci, per-base sub/ins/del, random strand; every fifth read
    carries a structural insertion or deletion (stalls a normal tile -> the 1984x960 / 960x1984 large tiles)."""
    reads = []
    for k in range(n_reads):
        L = int(rng.integers(read_len - read_len // 5, read_len + read_len // 5))
        p = int(rng.integers(0, len(genome) - L))
        src = genome[p:p + L]
        if structural and k % 5 == 1:
            src = np.concatenate([src[:L // 2], synth.random_seq(rng, int(rng.integers(200, 700))), src[L // 2:]])
        elif structural and k % 5 == 2:
            cut = int(rng.integers(300, 900))
            src = np.concatenate([src[:L // 3], src[L // 3 + cut:]])
        r = synth.mutate_fast(rng, src, *err)
        reads.append(np.ascontiguousarray(synth.revcomp(r) if k % 2 else r))
    return reads


def repeat_genome(rng, genome_len, n_repeats=6):
    """Random genome with a few diverged 3 kbp repeats, so that D-SOFT proposes secondary / spurious anchors."""
    genome = synth.random_seq(rng, genome_len)
    for _ in range(n_repeats):
        a, b = int(rng.integers(0, genome_len - 4000)), int(rng.integers(0, genome_len - 4000))
        rep = synth.mutate_fast(rng, genome[a:a + 3000], 0.03, 0.01, 0.01)[:2900]
        genome[b:b + len(rep)] = rep
    return genome


def reference_anchors(genomes, reads, T, O, ovl, scheme=STOCK):
    """The reference's own D-SOFT + first-tile filter + extender_body (both flavours) on `reads` against `genomes`.
    Returns dict(arena, anchors, hits, res, ops, asis_same, read_addr, read_len, chroms): res / ops from the patched flavour."""
    out = {}
    for fl in ("patched", "as-is"):
        ref = oracle.reference(fl)
        ref.set_scoring(abi.Scoring.from_values(*scheme))
        ref.set_dsoft_defaults()
        ref.set_extend(T, O, 2, ovl)
        ref.reset_arena()
        for k, g in enumerate(genomes):
            ref.add_chr("chr%d" % k, g.tobytes(), True)
        ref.build_index()
        addrs = []
        for k, r in enumerate(reads):
            _, a = ref.add_read("r%d" % k, r.tobytes())
            addrs.append(a)
        A, H, hb = [], [], 0
        for k in range(len(reads)):
            a, h = ref.seed_filter(k, 1)
            a = a.copy()
            a["left_hits_off"] += hb
            a["right_hits_off"] += hb
            hb += len(h)
            A.append(a)
            H.append(h)
        anchors, hits = np.concatenate(A), np.concatenate(H)
        res, ops = ref.extend(anchors, hits)
        out[fl] = dict(arena=ref.arena().copy(), anchors=anchors, hits=hits, res=res, ops=ops,
                       read_addr=np.array(addrs, np.uint64), read_len=np.array([len(r) for r in reads], np.uint32), chroms=ref.chroms())
    p, a = out["patched"], out["as-is"]
    assert np.array_equal(p["anchors"], a["anchors"]) and np.array_equal(p["hits"], a["hits"])     # score-only tiles: no UB bits
    p["asis_same"] = alignments_same(p["res"], p["ops"], a["res"], a["ops"])
    return p


def alignments_same(res_a, ops_a, res_b, ops_b, fields=ALN_FIELDS):
    """Per anchor bool: same emit decision, same reported fields, same op string."""
    same = np.zeros(len(res_a), bool)
    for k in range(len(res_a)):
        a, b = res_a[k], res_b[k]
        ok = (int(a["flags"]) & 1) == (int(b["flags"]) & 1) and all(a[f] == b[f] for f in fields)
        if ok and int(a["flags"]) & 1:
            ok = np.array_equal(ops_a[int(a["ops_offset"]):int(a["ops_offset"]) + int(a["n_ops"])],
                                ops_b[int(b["ops_offset"]):int(b["ops_offset"]) + int(b["n_ops"])])
        same[k] = ok
    return same


def check_asis_rule(flagged, asis_same, what):
    """flag clear => the as-is build of the reference gives the same answer; as-is differs => the flag is set."""
    flagged, asis_same = np.asarray(flagged, bool), np.asarray(asis_same, bool)
    unflagged_diff = np.flatnonzero(~flagged & ~asis_same)
    assert len(unflagged_diff) == 0, "%s: %d cases differ from the as-is reference without the long-insertion flag: %s" % (
        what, len(unflagged_diff), unflagged_diff[:10].tolist())
    return {"n": int(len(flagged)), "flagged": int(flagged.sum()), "asis_differs": int((~asis_same).sum())}
