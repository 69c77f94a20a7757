"""Reference-shaped pipeline (host threads play the reference's tokens) by adapter mode, host threads, reads per batch and
lanes: one case loaded once, every combination timed twice.  mode 4 = gpu_align_body (ExtendAlignments with gapped strings
for the unchanged printer), mode 5 = gpu_sam_body (reads in, SAM text out: darwin_gpu_align_reads + darwin_gpu_sam_select +
darwin_gpu_cigar).  Usage: python tests/tools/e2e_sweep.py [n_reads] [genome_bp]"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_e2e import load_driver, load_case
n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
genome = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
cores = os.cpu_count() or 1
t0 = time.time()
ref, L = load_driver()
load_case(ref, 9, genome, n_reads, 13333)
print("case: %d reads (mean 10 kbp, PacBio-like 1.5/9/4.5) against %d bp, %d host cores; loaded in %.1f s" % (n_reads, genome, cores, time.time() - t0), flush=True)
stats = (C.c_double * 8)()
# (lanes, [(mode, threads, reads per batch), ...])
plan = [(4, [(4, 48, 256), (5, 48, 256), (5, 32, 512), (5, 16, 1024), (5, 2 * cores, 256)]),
        (8, [(5, 48, 256), (5, 32, 512), (4, 48, 256)])]
for lanes, runs in plan:
    os.environ["DARWIN_GPU_LANES"] = str(lanes)
    t0 = time.time()
    assert L.dref_gpu_init(1) == 0
    assert L.dref_gpu_seed_index() == 0
    print("lanes %d: processors + arena upload + seed position table %.2f s" % (lanes, time.time() - t0), flush=True)
    for mode, threads, per_batch in runs:
        best = None
        for rep in range(3):
            n = L.dref_pipeline_mt(0, n_reads, threads, per_batch, mode, None, C.c_uint64(0), stats)
            assert n >= 0
            if rep and (best is None or stats[0] < best[0]):
                best = (stats[0], n)
        print("lanes %d mode %d threads %2d, %4d reads per batch: %.3f s, %6.0f reads/s, %d %s" % (
            lanes, mode, threads, per_batch, best[0], n_reads / best[0], best[1], "SAM lines" if mode == 5 else "alignments"), flush=True)
        if mode == 4:
            hp = (C.c_double * 3)()
            L.dref_host_profile(hp)
            print("    gpu_align_body thread-seconds over 3 runs: requests %.2f, blocked in the combiner %.2f, gapped strings %.2f" % tuple(hp), flush=True)
    L.dref_use_cpu_table(); L.dref_gpu_shutdown()
