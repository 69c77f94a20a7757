"""Where does the end-to-end tile leg lose time against the device-resident rate?  The bench's multi-lane e2e loop (host
buffers through Processor.BatchAlignmentSIMD, one handle per host thread) in variants: with / without the per-step ASCII
upload, different lane counts; prints GCUPS, ms per step and the mean CUDA-event kernel time per call.
Usage: python scripts/e2e_lanes.py [n_tiles] [steps_per_lane] [lanes:upload,...]"""
import os, sys, threading, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import darwin_b200
from darwin_b200 import abi
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
k_each = int(sys.argv[2]) if len(sys.argv) > 2 else 8
TILE = bench.TILE
arena, req = bench.make_workload(n, 1)
cells = float(n) * TILE * TILE
tbw = 2 * TILE // 32 + 2
sc = abi.Scoring.from_values()


def pinned(shape, dtype):
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    t = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    return t.numpy().view(dtype).reshape(shape)


h_arena = pinned(arena.shape, np.uint8); h_arena[:] = arena


def run(L, upload, stagger=True):
    procs = [darwin_b200.Processor(L * len(arena), 0)]
    procs[0].InitializeScoringParameters(sc)
    lanes = []
    for k in range(L):
        if k:
            procs.append(darwin_b200.Processor(0, 0, parent=procs[0]))
        r = pinned(req.shape, abi.TILE_REQ); r[:] = req
        r["ref_bases_start_addr"] += k * len(arena); r["query_bases_start_addr"] += k * len(arena)
        lanes.append((procs[k], k * len(arena), r, pinned((n,), abi.TILE_RES), pinned((n, tbw), np.uint64)))
        procs[k].InitializeReferenceMemory(k * len(arena), h_arena)
        procs[k].BatchAlignmentSIMD(r, 1, tbw, out=(lanes[k][3], lanes[k][4]))
    kms, ups, calls = [], [], []

    def work(p, off, hreq, hres, htb, wait_for, uploaded):
        if wait_for is not None and stagger:
            wait_for.wait()
        for _ in range(k_each):
            t0 = time.perf_counter()
            if upload:
                p.InitializeReferenceMemory(off, h_arena)
            t1 = time.perf_counter()
            uploaded.set()
            p.BatchAlignmentSIMD(hreq, 1, tbw, out=(hres, htb))
            t2 = time.perf_counter()
            kms.append(p.stats().last_kernel_ms); ups.append((t1 - t0) * 1e3); calls.append((t2 - t1) * 1e3)

    ev = [threading.Event() for _ in range(L)]
    th = [threading.Thread(target=work, args=(*lanes[k], ev[k - 1] if k else None, ev[k])) for k in range(L)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for x in th:
        x.start()
    for x in th:
        x.join()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    steps = L * k_each
    print("lanes %d upload %d stagger %d: %.0f GCUPS, %.1f ms/step | per call: kernel events %.1f ms, upload %.1f ms, tiles call %.1f ms" % (
        L, upload, stagger, cells * steps / wall / 1e9, wall * 1e3 / steps, np.mean(kms), np.mean(ups), np.mean(calls)), flush=True)
    for p in reversed(procs):
        p.close()


cases = [tuple(int(y) for y in x.split(':')) for x in sys.argv[3].split(',')] if len(sys.argv) > 3 else [(1, 0), (1, 1), (2, 0), (2, 1), (4, 0), (4, 1), (6, 1)]
for L, up in cases:
    run(L, up)
