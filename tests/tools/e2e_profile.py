"""Where does the host time go?  Single host thread, one lane: per-stage wall times of the GPU pipeline are not hidden behind
other threads.  Usage: python tests/tools/e2e_profile.py [n_reads] [per_batch]"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_e2e import load_driver, load_case
n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
per_batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
threads = int(sys.argv[3]) if len(sys.argv) > 3 else 1
mode = int(sys.argv[4]) if len(sys.argv) > 4 else 2          # 2: host D-SOFT + GPU filter/extend; 3: GPU D-SOFT as well
ref, L = load_driver()
load_case(ref, 9, 10_000_000, n_reads, 13333)
stats = (C.c_double * 8)()
gpus = int(os.environ.get('E2E_GPUS', '1'))
assert L.dref_gpu_init(gpus) == 0
if mode >= 3:
    import time as _t
    t0 = _t.time(); assert L.dref_gpu_seed_index() == 0; print("seed position table on the GPU: %.2f s" % (_t.time() - t0))
for rep in range(3):
    n = L.dref_pipeline_mt(0, n_reads, threads, per_batch, mode, None, C.c_uint64(0), stats)
    g = list(stats)
    nb = (n_reads + per_batch - 1) // per_batch
    print("gpus %d mode %d threads %d, %d reads, %d per batch: wall %.3f s (%.0f reads/s) | seed %.3f filter %.3f extend %.3f thread-s | per batch: seed %.2f ms filter %.2f ms extend %.2f ms" % (
        gpus, mode, threads, n_reads, per_batch, g[0], n_reads / g[0], g[2], g[3], g[4], g[2] / nb * 1e3, g[3] / nb * 1e3, g[4] / nb * 1e3))
if mode >= 4:
    hp = (C.c_double * 3)()
    L.dref_host_profile(hp)
    print("gpu_align_body thread-seconds over the 3 repetitions: building requests %.2f, blocked in the combiner %.2f, rebuilding ExtendAlignments %.2f" % tuple(hp))
L.dref_use_cpu_table(); L.dref_gpu_shutdown()
