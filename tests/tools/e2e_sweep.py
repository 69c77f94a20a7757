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
plan = [(4, [(4, 48, 256), (5, 48, 256), (5, 16, 1024), (5, 8, 2048), (5, 4, 4096), (5, 16, 2048)]),
        (2, [(5, 16, 1024), (5, 8, 2048)]),
        (8, [(5, 48, 256), (5, 32, 512)])]
if os.environ.get("E2E_SWEEP_PLAN") == "short":
    plan = [(4, [(5, 48, 256), (5, 16, 1024), (5, 8, 2048), (5, 4, 4096)])]
ph = (C.c_double * 6)()


def phases():
    L.dref_combiner_phases(ph)
    return list(ph)


for lanes, runs in plan:
    os.environ["DARWIN_GPU_LANES"] = str(lanes)
    t0 = time.time()
    assert L.dref_gpu_init(1) == 0
    assert L.dref_gpu_seed_index() == 0
    print("lanes %d: processors + arena upload + seed position table %.2f s" % (lanes, time.time() - t0), flush=True)
    for mode, threads, per_batch in runs:
        best = None
        for rep in range(3):
            p0 = phases()
            n = L.dref_pipeline_mt(0, n_reads, threads, per_batch, mode, None, C.c_uint64(0), stats)
            assert n >= 0
            p1 = phases()
            if rep and (best is None or stats[0] < best[0]):
                best = (stats[0], n, [b - a for a, b in zip(p0, p1)])
        d = best[2]
        print("lanes %d mode %d threads %2d, %4d reads per batch: %.3f s, %6.0f reads/s, %d %s | combining threads (summed over lanes): "
              "uploads %.3f s, inside device calls %.3f s, merge + scatter %.3f s; %d device calls for %d requests" % (
                  lanes, mode, threads, per_batch, best[0], n_reads / best[0], best[1], "SAM lines" if mode == 5 else "alignments",
                  d[0], d[1], d[2], d[3], d[4]), flush=True)
        if mode == 4:
            hp = (C.c_double * 3)()
            L.dref_host_profile(hp)
            print("    gpu_align_body thread-seconds over 3 runs: requests %.2f, blocked in the combiner %.2f, gapped strings %.2f" % tuple(hp), flush=True)
    L.dref_use_cpu_table(); L.dref_gpu_shutdown()
if os.environ.get("E2E_SWEEP_TIMING"):
    # per-phase host timings of the library calls (DARWIN_GPU_TIMING, stderr) for ONE lane and a few large batches
    os.environ["DARWIN_GPU_LANES"] = "1"; os.environ["DARWIN_GPU_TIMING"] = "1"
    assert L.dref_gpu_init(1) == 0 and L.dref_gpu_seed_index() == 0
    for rep in range(2):
        sys.stderr.write("== one lane, one thread, 2048 reads per batch, 4096 reads, pass %d\n" % rep); sys.stderr.flush()
        n = L.dref_pipeline_mt(0, min(n_reads, 4096), 1, 2048, 5, None, C.c_uint64(0), stats)
        print("one lane, one thread, 2048 reads per batch: %.3f s for %d reads (%.0f reads/s)" % (stats[0], min(n_reads, 4096), min(n_reads, 4096) / stats[0]), flush=True)
    L.dref_use_cpu_table(); L.dref_gpu_shutdown()
