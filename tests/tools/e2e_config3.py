"""Analysis script (not part of bench.py) for SURVEY 8(d) config 3: synthetic reference + PacBio-like 10 kbp reads at 15 %
error (sub 1.5 / ins 9 / del 4.5), reference-guided, END TO END: the reference's own D-SOFT (seeder_body) on all host
threads + first-tile filter and GACT extension on the GPU behind the cross-read combiner, next to the reference's CPU
pipeline on a bounded sample of the same reads.  Usage: python tests/tools/e2e_config3.py [genome_bp] [n_reads] [cpu_sample]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_e2e import load_driver, load_case  # noqa: E402

genome_bp = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
cpu_sample = int(sys.argv[3]) if len(sys.argv) > 3 else 320
threads = int(os.environ.get("E2E_THREADS", os.cpu_count() or 1))
ref, L = load_driver()
t0 = time.time()
load_case(ref, 9, genome_bp, n_reads, 13333)           # read lengths uniform in [6.7, 13.3] kbp, mean 10 kbp
print("loaded %d bp + %d reads, index built in %.1f s; %d host threads" % (genome_bp, n_reads, time.time() - t0, threads))
stats = (C.c_double * 8)()
n = L.dref_pipeline_mt(0, cpu_sample, threads, 1, 0, None, C.c_uint64(0), stats)
cpu = list(stats)
print("CPU pipeline (reference stages, %d threads), %d reads: %.2f s -> %.1f reads/s | thread-seconds seed %.2f filter %.2f extend %.2f | %d alignments" % (
    threads, cpu_sample, cpu[0], cpu_sample / cpu[0], cpu[2], cpu[3], cpu[4], n))
assert L.dref_gpu_init(1) == 0
for per_batch in (64, 16):
    for rep in range(2):
        n = L.dref_pipeline_mt(0, n_reads, threads, per_batch, 2, None, C.c_uint64(0), stats)
    g = list(stats)
    cs = (C.c_uint64 * 12)()
    L.dref_combiner_stats(cs)
    print("GPU pipeline (host D-SOFT x %d threads + GPU filter/extend, %d reads per host batch), %d reads: %.2f s -> %.0f reads/s "
          "| thread-seconds seed %.2f filter %.2f extend %.2f | %d alignments | combiner calls/requests: filter %d/%d extend %d/%d" % (
              threads, per_batch, n_reads, g[0], n_reads / g[0], g[2], g[3], g[4], n, cs[1], cs[4], cs[2], cs[5]))
    print("  speed-up over the CPU pipeline: %.1fx ; host D-SOFT bound (seed thread-seconds / threads): %.0f reads/s" % (
        (n_reads / g[0]) / (cpu_sample / cpu[0]), n_reads / (g[2] / threads)))
L.dref_use_cpu_table()
L.dref_gpu_shutdown()
