"""GPU: the whole reference-guided pipeline, multi-threaded -- the reference's own seeder_body on every host thread,
gpu_filter_body + gpu_extender_body behind the per-GPU combiner (cross-read batches) -- against the reference's CPU
pipeline on the same reads: byte-identical alignments (offsets, strand, AlignmentScore, gapped strings)."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
from darwin_b200 import abi, synth

pytestmark = pytest.mark.gpu
LIB = os.path.join(os.path.dirname(os.path.abspath(oracle.__file__)), "_ref", "libdarwin_ref_gpu.so")


def first_difference(mode, per_batch, n_gpu, n_cpu, buf_gpu, buf_cpu):
    """Failure message: which canonical alignment line differs first (read, chromosome, strand, offsets, score)."""
    g, c = buf_gpu.value.split(b"\n"), buf_cpu.value.split(b"\n")
    for k, (a, b) in enumerate(zip(g, c)):
        if a != b:
            return "mode %d, %d reads per batch: %d vs %d alignments; line %d differs: GPU %r | CPU %r" % (
                mode, per_batch, n_gpu, n_cpu, k, a[:80], b[:80])
    return "mode %d, %d reads per batch: %d vs %d alignments; common prefix identical (%d vs %d lines)" % (
        mode, per_batch, n_gpu, n_cpu, len(g), len(c))


def load_driver():
    ref = oracle.Reference.__new__(oracle.Reference)
    ref.lib = C.CDLL(LIB)
    L = ref.lib
    L.dref_arena.restype = C.c_void_p
    L.dref_arena_position.restype = C.c_uint64
    L.dref_add_chr.restype = C.c_uint64
    L.dref_anchor_hits_total.restype = C.c_uint64
    return ref, L


def load_case(ref, seed, genome_len, n_reads, read_len, err=(0.015, 0.09, 0.045), do_overlap=0, tile=(384, 64)):
    rng = np.random.default_rng(seed)
    ref.set_scoring(abi.Scoring.from_values())
    ref.set_dsoft_defaults()
    ref.set_extend(tile[0], tile[1], 2, do_overlap)
    ref.reset_arena()
    genome = synth.random_seq(rng, genome_len)
    for _ in range(max(2, genome_len // 500000)):
        a, b = int(rng.integers(0, genome_len - 4000)), int(rng.integers(0, genome_len - 4000))
        rep = synth.mutate_fast(rng, genome[a:a + 3000], 0.03, 0.01, 0.01)[:2900]
        genome[b:b + len(rep)] = rep
    ref.add_chr("chrS", genome.tobytes(), True)
    ref.build_index()
    for k in range(n_reads):
        L = int(rng.integers(read_len // 2, read_len))
        p = int(rng.integers(0, genome_len - L))
        src = genome[p:p + L]
        if k % 7 == 3:
            src = np.concatenate([src[:L // 2], synth.random_seq(rng, 400), src[L // 2:]])      # stalls -> large tiles
        r = synth.mutate_fast(rng, src, *err)
        if k % 2:
            r = synth.revcomp(r)
        ref.add_read("r%d" % k, np.ascontiguousarray(r).tobytes())


@pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libdarwin_ref_gpu.so not built (needs /root/reference at build time)")
def test_multithreaded_pipeline_matches_cpu_pipeline():
    ref, L = load_driver()
    n_reads = 96
    load_case(ref, 5, 600000, n_reads, 6000)
    cap = 256 << 20
    buf_cpu, buf_gpu = C.create_string_buffer(cap), C.create_string_buffer(cap)
    stats = (C.c_double * 8)()
    threads = min(8, os.cpu_count() or 1)
    n_cpu = L.dref_pipeline_mt(0, n_reads, threads, 4, 0, buf_cpu, C.c_uint64(cap), stats)
    assert n_cpu > n_reads // 2
    assert L.dref_gpu_init(1) == 0
    try:
        assert L.dref_gpu_seed_index() == 0                      # seed position table on the GPU (for mode 3)
        # mode 1: GPU extender; 2: + GPU first-tile filter; 3: + GPU D-SOFT (no reference stage left but slopeFilter)
        # mode 4: the resident pipeline call (darwin_gpu_align_reads) behind gpu_align_body
        for mode, per_batch in ((2, 4), (1, 7), (2, 1), (3, 5), (3, 32), (4, 6), (4, 48)):
            n_gpu = L.dref_pipeline_mt(0, n_reads, threads, per_batch, mode, buf_gpu, C.c_uint64(cap), stats)
            assert n_gpu == n_cpu and buf_gpu.value == buf_cpu.value, first_difference(mode, per_batch, n_gpu, n_cpu, buf_gpu, buf_cpu)
        cs = (C.c_uint64 * 12)()
        L.dref_combiner_stats(cs)
        # every request reached the device through the combiner; whether two threads' requests happened to ride in one device
        # call depends on timing (two threads per lane here), so merging itself is asserted in tests/test_combiner.py, where
        # the stand-in device has a fixed latency
        assert 0 < cs[2] <= cs[5] and cs[11] >= 1, list(cs)
    finally:
        L.dref_use_cpu_table()
        L.dref_gpu_shutdown()


@pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libdarwin_ref_gpu.so not built (needs /root/reference at build time)")
def test_pipeline_across_all_visible_gpus():
    """Tokens spread over every visible GPU (each with its own arena replica and seed position table, reads sharded by
    token, no data-path collective): same bytes as the CPU pipeline.  On a one-GPU box this repeats the single-GPU case."""
    import torch
    gpus = max(1, torch.cuda.device_count())
    ref, L = load_driver()
    n_reads = 64
    load_case(ref, 6, 500000, n_reads, 5000)
    cap = 256 << 20
    buf_cpu, buf_gpu = C.create_string_buffer(cap), C.create_string_buffer(cap)
    stats = (C.c_double * 8)()
    threads = min(8, os.cpu_count() or 1)
    n_cpu = L.dref_pipeline_mt(0, n_reads, threads, 4, 0, buf_cpu, C.c_uint64(cap), stats)
    assert n_cpu > n_reads // 2
    assert L.dref_gpu_init(gpus) == 0
    try:
        assert L.dref_gpu_seed_index() == 0
        for mode, per_batch in ((4, 3), (2, 5)):
            n_gpu = L.dref_pipeline_mt(0, n_reads, threads, per_batch, mode, buf_gpu, C.c_uint64(cap), stats)
            assert n_gpu == n_cpu and buf_gpu.value == buf_cpu.value, (gpus, mode)
    finally:
        L.dref_use_cpu_table()
        L.dref_gpu_shutdown()


@pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libdarwin_ref_gpu.so not built (needs /root/reference at build time)")
@pytest.mark.parametrize("do_overlap,tile", [(1, (384, 64)), (0, (320, 128)), (1, (256, 64))])
def test_pipeline_other_modes(do_overlap, tile):
    """De novo overlap mode (argv[3] = 1: D-SOFT stops after N+1 seeds, SV window of one bin, large tiles keep T) and other
    tile geometries through the staged and the resident GPU pipeline."""
    ref, L = load_driver()
    n_reads = 48
    try:
        load_case(ref, 8 + do_overlap, 400000, n_reads, 5000, do_overlap=do_overlap, tile=tile)
        cap = 256 << 20
        buf_cpu, buf_gpu = C.create_string_buffer(cap), C.create_string_buffer(cap)
        stats = (C.c_double * 8)()
        n_cpu = L.dref_pipeline_mt(0, n_reads, 4, 4, 0, buf_cpu, C.c_uint64(cap), stats)
        assert n_cpu > n_reads // 2
        assert L.dref_gpu_init(1) == 0
        try:
            assert L.dref_gpu_seed_index() == 0
            for mode, per_batch in ((3, 6), (4, 6), (4, 48)):
                n_gpu = L.dref_pipeline_mt(0, n_reads, 4, per_batch, mode, buf_gpu, C.c_uint64(cap), stats)
                assert n_gpu == n_cpu and buf_gpu.value == buf_cpu.value, (do_overlap, tile, mode)
        finally:
            L.dref_use_cpu_table()
            L.dref_gpu_shutdown()
    finally:
        ref.set_extend(384, 64, 2, 0)


@pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libdarwin_ref_gpu.so not built (needs /root/reference at build time)")
def test_de_novo_all_vs_all_overlap():
    """argv[3] = 1: the read set is its own reference -- every read is a chromosome of the seed position table
    (hundreds of short chromosomes: chromosome look-ups, tiles clamped at chromosome ends, extension running off both
    ends) and is aligned against all of them.  Staged and resident GPU pipelines vs the reference's CPU stages."""
    ref, L = load_driver()
    rng = np.random.default_rng(31)
    genome = synth.random_seq(rng, 150000)
    reads = []
    for k in range(90):                                            # ~2.5x coverage: neighbouring reads overlap
        Lr = int(rng.integers(2500, 6000))
        p = int(rng.integers(0, len(genome) - Lr))
        r = synth.mutate_fast(rng, genome[p:p + Lr], 0.03, 0.03, 0.03)
        reads.append(np.ascontiguousarray(synth.revcomp(r) if k % 3 == 0 else r))
    try:
        ref.set_scoring(abi.Scoring.from_values())
        ref.set_dsoft_defaults()
        ref.set_extend(384, 64, 2, 1)
        ref.reset_arena()
        for k, r in enumerate(reads):
            ref.add_chr("read%d" % k, r.tobytes(), True)
        ref.build_index()
        for k, r in enumerate(reads):
            ref.add_read("read%d" % k, r.tobytes())
        n_reads = len(reads)
        cap = 256 << 20
        buf_cpu, buf_gpu = C.create_string_buffer(cap), C.create_string_buffer(cap)
        stats = (C.c_double * 8)()
        n_cpu = L.dref_pipeline_mt(0, n_reads, 4, 5, 0, buf_cpu, C.c_uint64(cap), stats)
        assert n_cpu > n_reads                                     # every read finds itself and its neighbours
        assert L.dref_gpu_init(1) == 0
        try:
            assert L.dref_gpu_seed_index() == 0
            for mode, per_batch in ((2, 5), (3, 7), (4, 4), (4, 90)):
                n_gpu = L.dref_pipeline_mt(0, n_reads, 4, per_batch, mode, buf_gpu, C.c_uint64(cap), stats)
                assert n_gpu == n_cpu and buf_gpu.value == buf_cpu.value, first_difference(mode, per_batch, n_gpu, n_cpu, buf_gpu, buf_cpu)
            # the reference's MHAP printer (printer.cpp:100-180) on top of the GPU stages
            sam_cpu, sam_gpu = C.create_string_buffer(cap), C.create_string_buffer(cap)
            n2 = L.dref_pipeline(0, n_reads, 8 | 3, sam_gpu, C.c_uint64(cap))
            L.dref_use_cpu_table()
            n1 = L.dref_pipeline(0, n_reads, 8, sam_cpu, C.c_uint64(cap))
            assert n1 > 0 and n1 == n2 and sam_cpu.value == sam_gpu.value
        finally:
            L.dref_use_cpu_table()
            L.dref_gpu_shutdown()
    finally:
        ref.set_extend(384, 64, 2, 0)
