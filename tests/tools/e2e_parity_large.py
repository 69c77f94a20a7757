"""One-off scale check (not a test): the whole pipeline on a few thousand 10 kbp reads, reference CPU stages vs the GPU
pipeline in every mode, canonical alignment text compared byte for byte.  Usage: python tests/tools/e2e_parity_large.py [n_reads] [genome_bp]"""
import ctypes as C, hashlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_e2e import load_driver, load_case
n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
genome = int(sys.argv[2]) if len(sys.argv) > 2 else 5_000_000
ref, L = load_driver()
read_len = int(os.environ.get('E2E_READ_LEN', '13333'))
err = tuple(float(x) for x in os.environ.get('E2E_ERR', '0.015,0.09,0.045').split(','))      # sub, ins, del
tile = tuple(int(x) for x in os.environ.get('E2E_TILE', '384,64').split(','))
load_case(ref, 21, genome, n_reads, read_len, err=err, tile=tile)
cap = 1 << 30
buf_cpu, buf_gpu = C.create_string_buffer(cap), C.create_string_buffer(cap)
stats = (C.c_double * 8)()
threads = int(os.environ.get("E2E_THREADS", os.cpu_count() or 1))
t0 = time.time()
n_cpu = L.dref_pipeline_mt(0, n_reads, threads, 1, 0, buf_cpu, C.c_uint64(cap), stats)
cpu_text = buf_cpu.value
print("CPU pipeline: %d alignments, %.1f MB of text, %.1f s, sha1 %s" % (n_cpu, len(cpu_text) / 1e6, time.time() - t0, hashlib.sha1(cpu_text).hexdigest()[:16]))
assert L.dref_gpu_init(1) == 0 and L.dref_gpu_seed_index() == 0
ok = True
for mode, per_batch in ((2, 64), (3, 64), (4, 64), (4, 1000)):
    t0 = time.time()
    n_gpu = L.dref_pipeline_mt(0, n_reads, threads, per_batch, mode, buf_gpu, C.c_uint64(cap), stats)
    same = n_gpu == n_cpu and buf_gpu.value == cpu_text
    ok &= same
    print("GPU mode %d, %4d reads per batch: %d alignments, %.2f s, identical: %s" % (mode, per_batch, n_gpu, time.time() - t0, same))
L.dref_use_cpu_table(); L.dref_gpu_shutdown()
print("ALL IDENTICAL" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
