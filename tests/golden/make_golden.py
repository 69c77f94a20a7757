"""Regenerates the golden fixtures under tests/golden/ from the COMPILED REFERENCE (oracle/_ref, built by
oracle/Makefile from the unmodified sources under /root/reference).  Run here (CPU container):

    python tests/golden/make_golden.py

Fixtures (all seeded; inputs are synthetic, expected outputs come from the reference itself):
  tiles_v1.npz   tile requests of many shapes/flags/scoring schemes + BatchAlignmentSIMD results (patched flavour;
                 `asis_same` marks tiles where the as-is flavour gave the identical answer)
  extend_v1.npz  synthetic 150 kbp reference + 10 reads, anchors from the reference's own D-SOFT + filter,
                 extender_body results for T/O = 384/64 and 320/128 (+ do_overlap=1)
  filter_v1.npz  3 chromosomes (one 5 kbp) + 48 reads (some hanging over chromosome ends): the candidates of the reference's
                 seeder_body, the first-tile results of its BatchAlignmentSIMD on filter_body's requests, and the
                 ExtendLocations its filter_body (tiles + slope filter) returns
  config1_v1.npz BASELINE.json configs[0]: the reference's own sample reference (software/data/sample_ref.fa, sacCer3 chrI)
                 with 40 simulated reads (the reads file is missing from the reference tree), stock params.cfg read through
                 the reference's ConfigFile: arena, anchors of its seeder + filter, extender_body results (patched; `asis_same`)
  rtl_kat.npz    the RTL testbench's 10 known-answer pairs (RTL/GACT/test_data/{ref,query}_320.txt) and
                 their "Total score" lines (test_align.txt) -- score-level vectors only
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from darwin_b200 import abi, synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"

SCHEMES = {   # (match, mismatch, N; go, ge; lgo, lge) -- SURVEY Appendix D
    "stock": (2, -6, -1, -4, -2, -25, -1),
    "tie": (1, -1, 0, -1, -1, -1, -1),
    "s2": (1, -1, 0, -2, -1, -4, 0),
    "s3": (2, -3, -1, -3, -2, -8, -1),
}
FLAGSETS = [1, 1 | 4 | 16, 1 | 4 | 2, 1 | 16 | 2, 0, 4 | 2, 1 | 8, 1 | 8 | 16 | 4 | 2]


def make_tiles(seed, n, sizes, big=0):
    rng = np.random.default_rng(seed)
    arena = [np.full(128, ord("N"), np.uint8)]
    pos = 128
    req = np.zeros(n + big, abi.TILE_REQ)
    for k in range(n + big):
        if k >= n:
            R, Qc = (1984, 960) if k % 2 else (960, 1984)
        else:
            R = int(rng.choice(sizes)) if rng.random() < 0.3 else int(rng.integers(1, max(sizes) + 1))
            Qc = None
        r = synth.random_seq(rng, R)
        q = synth.mutate(rng, r, 0.05 + 0.1 * rng.random(), 0.02 + 0.05 * rng.random(), 0.02 + 0.05 * rng.random(),
                         0.003 if k % 3 == 0 else 0.0, (2, 40) if k % 2 else None)
        if Qc is not None:
            q = np.concatenate([q, synth.random_seq(rng, max(0, Qc - len(q)))])[:Qc]
        elif rng.random() < 0.5:
            q = q[:max(1, min(len(q), int(rng.integers(1, max(sizes) + 1))))]
        if len(q) == 0:
            q = synth.random_seq(rng, 1)
        if rng.random() < 0.1:
            q = np.char.lower(q.view("S1")).view(np.uint8)           # Nt2Int accepts either case
        req[k]["ref_bases_start_addr"] = pos
        arena.append(r)
        pos += len(r)
        req[k]["query_bases_start_addr"] = pos
        arena.append(q)
        pos += len(q)
        req[k]["ref_size"] = len(r)
        req[k]["query_size"] = len(q)
        req[k]["max_tb_steps"] = 768 if k >= n else int(rng.choice([2 * max(len(r), len(q)), 64, 768]))
        req[k]["align_fields"] = 1 if k >= n else int(rng.choice(FLAGSETS))
        req[k]["index"] = k % 250
    arena.append(np.full(128, ord("N"), np.uint8))
    return np.concatenate(arena), req


def gen_tiles():
    out = {}
    for si, (name, vals) in enumerate(SCHEMES.items()):
        sc = abi.Scoring.from_values(*vals)
        arena, req = make_tiles(100 + si, 120, [64, 128, 320, 400], big=2 if name == "stock" else 0)
        res = {}
        for fl in ("patched", "as-is"):
            ref = oracle.reference(fl)
            ref.set_scoring(sc)
            res[fl] = ref.tiles(arena, req, 1, tb_words_per_req=260)
        rp, tp = res["patched"]
        ra, ta = res["as-is"]
        same = np.array([(rp[k] == ra[k]) and np.array_equal(tp[k], ta[k]) for k in range(len(req))])
        # score-only pass (do_traceback = 0), as the first-tile filter calls it (filter.cpp:40,:77)
        ref = oracle.reference("patched")
        ref.set_scoring(sc)
        r0, _ = ref.tiles(arena, req, 0, tb_words_per_req=1)
        used = int(((rp["total_TB_pointers"].astype(int) + 31) // 32).max())
        out[name + "_scoring"] = np.array(vals, np.int32)
        out[name + "_arena"] = arena
        out[name + "_req"] = req
        out[name + "_res"] = rp
        out[name + "_tb"] = tp[:, :used + 1].copy()
        out[name + "_res_notb"] = r0
        out[name + "_asis_same"] = same
        print("tiles", name, len(req), "as-is identical:", int(same.sum()))
    np.savez_compressed(os.path.join(OUT, "tiles_v1.npz"), **out)


ALN_CMP = ("flags", "n_ops", "reference_start_offset", "reference_end_offset", "query_start_offset", "query_end_offset",
           "n_tiles", "n_large_tiles", "score", "cells")


def same_alignments(rp, opsp, ra, opsa):
    """Per anchor: do the two flavours agree on every reported field and on the op string?  (`ops_offset` is a position in
    the flavour's own pool and is not compared.)"""
    return np.array([all(rp[k][f] == ra[k][f] for f in ALN_CMP) and
                     np.array_equal(opsp[rp[k]["ops_offset"]:rp[k]["ops_offset"] + rp[k]["n_ops"]],
                                    opsa[ra[k]["ops_offset"]:ra[k]["ops_offset"] + ra[k]["n_ops"]])
                     for k in range(len(rp))])


def gen_extend():
    rng = np.random.default_rng(7)
    genome = synth.random_seq(rng, 150000)
    reads = []
    for k in range(10):
        L = int(rng.integers(5000, 8000))
        p = int(rng.integers(0, len(genome) - L))
        src = genome[p:p + L]
        if k % 4 == 1:      # structural insertion in the read: stalls a normal tile, chained hits remain -> large tile
            src = np.concatenate([src[:L // 2], synth.random_seq(rng, 450), src[L // 2:]])
        elif k % 4 == 2:    # structural deletion
            src = np.concatenate([src[:L // 3], src[L // 3 + 600:]])
        r = synth.mutate(rng, src, 0.05, 0.05, 0.05, indel_run=(3, 60) if k % 2 else None)
        if k % 3 == 0:
            r = synth.revcomp(r)
        reads.append(np.ascontiguousarray(r))
    out = {}
    for (T, O, ovl) in ((384, 64, 0), (320, 128, 0), (256, 64, 1)):
        res = {}
        for fl in ("patched", "as-is"):
            ref = oracle.reference(fl)
            ref.set_scoring(abi.Scoring.from_values(*SCHEMES["stock"]))
            ref.set_dsoft_defaults()
            ref.set_extend(T, O, 2, ovl)
            ref.reset_arena()
            ref.add_chr("chrS", genome.tobytes(), True)
            ref.build_index()
            for k, r in enumerate(reads):
                ref.add_read("r%d" % k, r.tobytes())
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
            r_, ops = ref.extend(anchors, hits)
            used = int((r_["ops_offset"] + r_["n_ops"]).max())
            res[fl] = (anchors, hits, r_, ops[:used].copy(), ref.arena().copy())
        (anchors, hits, rp, opsp, arena) = res["patched"]
        (_, _, ra, opsa, _) = res["as-is"]
        same = same_alignments(rp, opsp, ra, opsa)
        tag = "T%d_O%d_ovl%d" % (T, O, ovl)
        out[tag + "_anchors"] = anchors
        out[tag + "_hits"] = hits
        out[tag + "_res"] = rp
        out[tag + "_ops"] = opsp
        out[tag + "_asis_same"] = same
        out["arena"] = arena
        print("extend", tag, "anchors", len(anchors), "emitted", int((rp["flags"] & 1).sum()), "tiles", int(rp["n_tiles"].sum()),
              "large", int(rp["n_large_tiles"].sum()), "as-is identical", int(same.sum()))
    out["scoring"] = np.array(SCHEMES["stock"], np.int32)
    np.savez_compressed(os.path.join(OUT, "extend_v1.npz"), **out)


def filter_case(seed=11, n_reads=48):
    """Shared by the generator and the live tests: (reference driver with arena + index + reads loaded)."""
    rng = np.random.default_rng(seed)
    ref = oracle.reference("patched")
    ref.set_scoring(abi.Scoring.from_values(*SCHEMES["stock"]))
    ref.set_dsoft_defaults()
    ref.reset_arena()
    chrs = [synth.random_seq(rng, 100), synth.random_seq(rng, 180000), synth.random_seq(rng, 5000), synth.random_seq(rng, 90000)]
    chrs[0][:] = chrs[3][500:600]            # the first chromosome is SHORTER than a first tile (filter.cpp:57 falls back to 0)
    a = chrs[1]
    for _ in range(5):                       # diverged repeats -> secondary candidates, slope-filter work
        x, y = int(rng.integers(0, 170000)), int(rng.integers(0, 170000))
        rep = synth.mutate_fast(rng, a[x:x + 3000], 0.03, 0.01, 0.01)[:2900]
        a[y:y + len(rep)] = rep
    for k, c in enumerate(chrs):
        ref.add_chr("chr%d" % k, c.tobytes(), True)
    ref.build_index()
    for k in range(n_reads):
        c = chrs[1 + k % 3]
        L = int(rng.integers(600, min(9000, len(c) - 10))) if k % 8 != 7 else int(rng.integers(70, 128))   # reads shorter than a tile
        if k % 6 == 0:
            p = len(c) - L                   # ends exactly at the chromosome end: tiles clamp to chr_end - first_tile_size
        elif k % 6 == 1:
            p = 0
        else:
            p = int(rng.integers(0, len(c) - L))
        r = synth.mutate_fast(rng, c[p:p + L], 0.05, 0.05, 0.05)
        if k % 2:
            r = synth.revcomp(r)
        ref.add_read("r%d" % k, np.ascontiguousarray(r).tobytes())
    return ref, n_reads


def custom_candidates(ref, n_reads, seed=5, per_read=6):
    """Candidates D-SOFT would rarely propose: random loci (low scores), hits within a tile of a chromosome end, offsets
    within a tile of the read end, the sub-tile chromosome, both strands.  Sorted by read within each strand."""
    rng = np.random.default_rng(seed)
    arena_chr = [(int(ref.lib.dref_chr_start(k)), int(ref.lib.dref_chr_len(k))) for k in range(ref.lib.dref_num_chr())]
    hit, off, rn, st = [], [], [], []
    for strand in (0, 1):
        for r in range(n_reads):
            rl = int(ref.lib.dref_read_len(r))
            for j in range(per_read):
                cs, cl = arena_chr[int(rng.integers(0, len(arena_chr)))]
                mode = (r + j) % 4
                h = cs + (int(rng.integers(0, cl)) if mode < 2 else max(0, cl - 1 - int(rng.integers(0, 140))))
                o = int(rng.integers(0, rl)) if mode % 2 == 0 else max(0, rl - 1 - int(rng.integers(0, 140)))
                hit.append(h); off.append(o); rn.append(r); st.append(strand)
    return np.array(hit, np.uint64), np.array(off, np.uint64), np.array(rn, np.int32), np.array(st, np.uint8)


def gen_filter():
    ref, n_reads = filter_case()
    # part 2 first (it only needs the arena): hand-made candidates through the reference's filter_body with thresholds
    # that let EVERY candidate through (score >= 0, min_overlap 0, slope filter off) -> per-candidate reference values
    hit, off, rn2, st = custom_candidates(ref, n_reads)
    ref.set_first_tile(128, 0, 0, -1.0)
    cands2, crn2 = ref.seed_custom(0, n_reads, hit, off, rn2, st)
    all_loc, _ = ref.filter_last()
    ref.set_first_tile(96, 0, 0, -1.0)                       # a first_tile_size other than 128 (ragged lanes on the GPU)
    all_loc96, _ = ref.filter_last()
    ref.set_first_tile()
    cands, rn = ref.seed(0, n_reads)
    anchors, hits = ref.filter_last()
    arena = ref.arena().copy()
    # the reference's own BatchAlignmentSIMD on the requests filter_body builds (restated by the port; equality of the
    # final ExtendLocations above pins that restatement)
    port = oracle.port(abi.Scoring.from_values(*SCHEMES["stock"]))
    pres = port.filter(arena, cands)
    np.savez_compressed(os.path.join(OUT, "filter_v1.npz"), arena=arena, cands=cands, cand_read_num=rn, anchors=anchors,
                        hits=hits, port_res=pres, scoring=np.array(SCHEMES["stock"], np.int32),
                        custom_cands=cands2, custom_read_num=crn2, custom_locations=all_loc, custom_locations96=all_loc96)
    print("filter: candidates", len(cands), "rc", int(cands["strand"].sum()), "locations", len(anchors),
          "| custom", len(cands2), "locations", len(all_loc), len(all_loc96), "score<60:", int((all_loc["score"] < 60).sum()))


def gen_config1(n_reads=40):
    """configs[0]: software/data/sample_ref.fa + software/params.cfg, reference-guided.  The arena (chromosome + reads as the
    reference's reader lays them out) is committed, so the GPU test needs neither /root/reference nor the read simulator."""
    seq = b"".join(l.strip() for l in open(os.path.join(REF, "software", "data", "sample_ref.fa"), "rb") if not l.startswith(b">"))
    g = np.char.upper(np.frombuffer(seq, np.uint8).view("S1")).view(np.uint8)
    res = {}
    for fl in ("patched", "as-is"):
        rng = np.random.default_rng(1)
        ref = oracle.reference(fl)
        ref.load_cfg(os.path.join(REF, "software", "params.cfg"), 0)
        ref.reset_arena()
        ref.add_chr("sacCer3.chrI", seq, True)
        ref.build_index()
        addrs, lens = [], []
        for k in range(n_reads):
            L = int(rng.integers(8000, 12000))
            p = int(rng.integers(0, len(g) - L))
            r = synth.mutate_fast(rng, g[p:p + L], 0.05, 0.05, 0.05)
            if rng.random() < 0.5:
                r = synth.revcomp(r)
            r = np.ascontiguousarray(r)
            _, addr = ref.add_read("r%d" % k, r.tobytes())
            addrs.append(addr)
            lens.append(len(r))
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
        r_, ops = ref.extend(anchors, hits)
        used = int((r_["ops_offset"] + r_["n_ops"]).max())
        res[fl] = (anchors, hits, r_, ops[:used].copy(), ref.arena().copy(), np.array(addrs, np.uint64), np.array(lens, np.uint32), ref.chroms())
    anchors, hits, rp, opsp, arena, addrs, lens, chroms = res["patched"]
    _, _, ra, opsa, _, _, _, _ = res["as-is"]
    same = same_alignments(rp, opsp, ra, opsa)
    # 2 bits per base where possible keeps the fixture small: the arena is stored as is (deflate does the rest)
    np.savez_compressed(os.path.join(OUT, "config1_v1.npz"), arena=arena, read_addr=addrs, read_len=lens, chroms=chroms,
                        anchors=anchors, hits=hits, res=rp, ops=opsp, asis_same=same, scoring=np.array(SCHEMES["stock"], np.int32),
                        extend=np.array([384, 64, 0], np.int32))
    print("config1: reads", n_reads, "anchors", len(anchors), "emitted", int((rp["flags"] & 1).sum()), "tiles", int(rp["n_tiles"].sum()),
          "large", int(rp["n_large_tiles"].sum()), "as-is identical", int(same.sum()))


def gen_rtl():
    d = os.path.join(REF, "RTL", "GACT", "test_data")
    refs = open(os.path.join(d, "ref_320.txt")).read().split()
    qrys = open(os.path.join(d, "query_320.txt")).read().split()
    scores = [int(l.split(":")[1]) for l in open(os.path.join(d, "test_align.txt")) if l.startswith("Total score")]
    assert len(refs) == len(qrys) == len(scores) == 10
    np.savez_compressed(os.path.join(OUT, "rtl_kat.npz"), refs=np.array(refs), queries=np.array(qrys),
                        scores=np.array(scores, np.int32))
    print("rtl", scores)


if __name__ == "__main__":
    oracle.build()
    gen_tiles()
    gen_extend()
    gen_filter()
    gen_config1()
    gen_rtl()
