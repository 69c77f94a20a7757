"""GPU: the reference's own pipeline (its seeder and filter, unmodified) with libdarwin_gact.so swapped in through the
C++ host adapter (darwin_b200/host/): g_BatchAlignmentSIMD -> darwin_gpu_tiles (first-tile filter),
extender_body -> gpu_extender_body -> darwin_gpu_extend.  The ExtendAlignments (offsets, strand, AlignmentScore and the
gapped strings) must equal those of the untouched CPU pipeline."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
from darwin_b200 import abi, synth

pytestmark = pytest.mark.gpu
LIB = os.path.join(os.path.dirname(os.path.abspath(oracle.__file__)), "_ref", "libdarwin_ref_gpu.so")


@pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libdarwin_ref_gpu.so not built (needs /root/reference at build time)")
def test_reference_pipeline_with_gpu_processor():
    ref = oracle.Reference.__new__(oracle.Reference)
    ref.lib = C.CDLL(LIB)
    L = ref.lib
    L.dref_arena.restype = C.c_void_p
    L.dref_arena_position.restype = C.c_uint64
    L.dref_add_chr.restype = C.c_uint64
    ref.set_scoring(abi.Scoring.from_values())
    ref.set_dsoft_defaults()
    ref.set_extend(384, 64, 2, 0)
    L.dref_set_dsoft(14, 3, 64, 26, 1000, 40, 1000, 4, 128, 60, 64, 1000, C.c_float(0.05))
    ref.reset_arena()
    rng = np.random.default_rng(42)
    genome = synth.random_seq(rng, 120000)
    ref.add_chr("chrS", genome.tobytes(), True)
    ref.build_index()
    nreads = 8
    for k in range(nreads):
        Lr = int(rng.integers(4000, 6000))
        p = int(rng.integers(0, len(genome) - Lr))
        src = genome[p:p + Lr]
        if k % 4 == 1:
            src = np.concatenate([src[:Lr // 2], synth.random_seq(rng, 450), src[Lr // 2:]])    # forces large tiles
        r = synth.mutate(rng, src, 0.05, 0.05, 0.05)
        if k % 2:
            r = synth.revcomp(r)
        ref.add_read("r%d" % k, np.ascontiguousarray(r).tobytes())
    cap = 64 << 20
    buf_cpu, buf_gpu = C.create_string_buffer(cap), C.create_string_buffer(cap)
    n_cpu = L.dref_pipeline(0, nreads, 0, buf_cpu, C.c_uint64(cap))           # software Processor + extender_body
    assert n_cpu > 0
    assert L.dref_gpu_init(1) == 0                                            # install the GPU table, upload the arena
    n_gpu = L.dref_pipeline(0, nreads, 1, buf_gpu, C.c_uint64(cap))           # GPU filter tiles + gpu_extender_body
    L.dref_use_cpu_table()
    L.dref_gpu_shutdown()
    assert n_gpu == n_cpu
    assert buf_gpu.value == buf_cpu.value


@pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libdarwin_ref_gpu.so not built (needs /root/reference at build time)")
@pytest.mark.parametrize("do_overlap", [0, 1])
def test_output_stage_bytes_are_identical(do_overlap):
    """The reference's own printer_body (printer.cpp: SAM with CIGAR / overlap suppression, MHAP in overlap mode) fed by
    the GPU stages prints byte-for-byte what it prints after the CPU stages."""
    ref = oracle.Reference.__new__(oracle.Reference)
    ref.lib = C.CDLL(LIB)
    L = ref.lib
    L.dref_arena.restype = C.c_void_p
    L.dref_arena_position.restype = C.c_uint64
    L.dref_add_chr.restype = C.c_uint64
    ref.set_scoring(abi.Scoring.from_values())
    ref.set_dsoft_defaults()
    ref.set_extend(384, 64, 2, do_overlap)
    ref.reset_arena()
    rng = np.random.default_rng(43 + do_overlap)
    genome = synth.random_seq(rng, 100000)
    rep = synth.mutate_fast(rng, genome[1000:4000], 0.02, 0.01, 0.01)[:2900]
    genome[60000:60000 + len(rep)] = rep                                   # secondary alignments -> overlap suppression
    ref.add_chr("chrS", genome.tobytes(), True)
    ref.build_index()
    nreads = 10
    for k in range(nreads):
        Lr = int(rng.integers(3000, 5000))
        p = int(rng.integers(0, len(genome) - Lr)) if k else 500
        r = synth.mutate(rng, genome[p:p + Lr], 0.04, 0.04, 0.04)
        if k % 2:
            r = synth.revcomp(r)
        ref.add_read("read_%d" % k, np.ascontiguousarray(r).tobytes())
    cap = 64 << 20
    buf_cpu, buf_gpu = C.create_string_buffer(cap), C.create_string_buffer(cap)
    n_cpu = L.dref_pipeline(0, nreads, 8, buf_cpu, C.c_uint64(cap))            # reference stages + reference printer
    assert n_cpu > (0 if do_overlap else nreads)
    assert L.dref_gpu_init(1) == 0
    try:
        n_gpu = L.dref_pipeline(0, nreads, 8 | 2, buf_gpu, C.c_uint64(cap))    # GPU filter + GPU extender + reference printer
    finally:
        L.dref_use_cpu_table()
        L.dref_gpu_shutdown()
    assert n_gpu == n_cpu and buf_gpu.value == buf_cpu.value
    if not do_overlap:
        text = buf_cpu.value.decode()
        assert text.startswith("@HD\tVN:1.6") and "\tAS:i:" in text and "M" in text.split("\n")[2].split("\t")[5]


@pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libdarwin_ref_gpu.so not built (needs /root/reference at build time)")
def test_native_output_stage_prints_the_reference_sam():
    """gpu_sam_body (reads in, SAM out: darwin_gpu_align_reads + darwin_gpu_sam_select + darwin_gpu_cigar, no gapped strings, no
    printer.cpp) against the reference's CPU stages + its own printer_body: same header, same lines (QNAME, FLAG, RNAME, POS,
    CIGAR with soft clips, SEQ of the right strand, AS / ZS), same overlap suppression -- byte for byte."""
    ref = oracle.Reference.__new__(oracle.Reference)
    ref.lib = C.CDLL(LIB)
    L = ref.lib
    L.dref_arena.restype = C.c_void_p
    L.dref_arena_position.restype = C.c_uint64
    L.dref_add_chr.restype = C.c_uint64
    ref.set_scoring(abi.Scoring.from_values())
    ref.set_dsoft_defaults()
    ref.set_extend(384, 64, 2, 0)
    ref.reset_arena()
    rng = np.random.default_rng(77)
    g1, g2 = synth.random_seq(rng, 90000), synth.random_seq(rng, 40000)
    for a, b in ((1000, 50000), (20000, 70000)):                           # diverged repeats: secondary alignments, suppression
        rep = synth.mutate_fast(rng, g1[a:a + 3000], 0.02, 0.01, 0.01)[:2900]
        g1[b:b + len(rep)] = rep
    ref.add_chr("chrA", g1.tobytes(), True)
    ref.add_chr("chrB", g2.tobytes(), True)
    ref.build_index()
    nreads = 24
    for k in range(nreads):
        g = g2 if k % 5 == 4 else g1
        Lr = int(rng.integers(3000, 6000))
        p = int(rng.integers(0, len(g) - Lr)) if k > 2 else (500, 19500, len(g1) - Lr)[k]
        src = g[p:p + Lr]
        if k % 6 == 1:
            src = np.concatenate([src[:Lr // 2], synth.random_seq(rng, 450), src[Lr // 2:]])      # structural insertion: large tiles
        r = synth.mutate(rng, src, 0.04, 0.04, 0.04)
        if k % 7 == 3:
            r = np.concatenate([synth.random_seq(rng, 300), r, synth.random_seq(rng, 200)])       # unaligned flanks: soft clips
        if k % 2:
            r = synth.revcomp(r)
        ref.add_read("read_%d some description" % k, np.ascontiguousarray(r).tobytes())
    cap = 64 << 20
    buf_cpu, buf_gpu = C.create_string_buffer(cap), C.create_string_buffer(cap)
    n_cpu = L.dref_pipeline(0, nreads, 8, buf_cpu, C.c_uint64(cap))            # reference stages + reference printer
    assert n_cpu > nreads
    assert L.dref_gpu_init(1) == 0
    try:
        assert L.dref_gpu_seed_index() == 0
        n_gpu = L.dref_pipeline(0, nreads, 8 | 5, buf_gpu, C.c_uint64(cap))    # all stages on the GPU + native output stage
    finally:
        L.dref_use_cpu_table()
        L.dref_gpu_shutdown()
    assert n_gpu == n_cpu
    assert buf_gpu.value == buf_cpu.value
    lines = [l.split("\t") for l in buf_cpu.value.decode().split("\n") if l and not l.startswith("@")]
    assert any(l[1] == "80" for l in lines) and any(l[1] == "64" for l in lines)               # both strands
    assert any("S" in l[5] for l in lines) and any("I" in l[5] and "D" in l[5] for l in lines)  # clips and gaps in the CIGARs
