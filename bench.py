#!/usr/bin/env python
"""bench.py -- GACT throughput on B200 (BASELINE.json: "GACT GCUPS ... vs reference TBB+AVX CPU").

Workload (config.workload = "gact_tiles_T320_O128"): BASELINE.json configs[1] / SURVEY 8(d).2 -- independent
320x320 tiles on random sequences with 15 % mutations (5/5/5 sub/ins/del), half left-extension
(start_end) and half right-extension (reverse_ref|reverse_query|start_end) requests, corner traceback,
max_tb_steps 640.  One "step" = one pass of the hot path over the whole batch (default 1M tiles per GPU).

  value : GCUPS with the packed arena, requests and outputs resident in HBM (CUDA events on the library's stream)
  e2e   : GCUPS through the public call `Processor.BatchAlignmentSIMD` with HOST buffers: ASCII upload of the
          batch's sequences + request H2D + result/TB-word D2H inside the timed region, every step; the steps are
          issued from two host threads (one handle each, one shared arena replica: the reference's token model), so
          one step's copies overlap the other lane's kernels.  e2e.single_lane_value = the same from one thread
  roofline: integer-pipe cell-update roofline of SURVEY 8(d): achieved = cells/s * 32 int-ops, peak = measured
          packed-int16 ALU issue rate of this GPU (darwin_gpu_int_peak microbenchmark, run live)
  cpu_baseline: the reference's own BatchAlignmentSIMD (oracle/_ref, compiled from the unmodified sources) on a
          bounded sample of the same tiles, all host cores

`--impl reference` times that CPU path alone (rank 0 only).  Multi-GPU: launched under torchrun, one rank per GPU,
each rank owns an independent shard of tiles (weak scaling), no data-path collective; the barrier and the max over
ranks go through torch.distributed (NCCL).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TILE, OVERLAP = 320, 128
OPS_PER_CELL = 32          # algorithmic integer ops per DP cell


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons of one GPU during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.stop_flag = index, [], set(), None, False

    def run(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def make_workload(n_tiles, seed):
    from darwin_b200 import synth
    chunks, reqs, base = [], [], 0
    step = 100000
    for c0 in range(0, n_tiles, step):
        n = min(step, n_tiles - c0)
        arena, req = synth.tile_batch_fast(seed * 1000 + c0 // step, n, TILE)
        arena = arena[:n * 2 * TILE]
        req["ref_bases_start_addr"] += base
        req["query_bases_start_addr"] += base
        base += len(arena)
        chunks.append(arena)
        reqs.append(req)
    arena = np.concatenate(chunks + [np.full(128, ord("N"), np.uint8)])
    return arena, np.concatenate(reqs)


def cpu_reference_leg(arena, req, seconds_target, threads=None):
    """The reference's own BatchAlignmentSIMD (oracle/_ref) on a bounded sample, all host cores."""
    import oracle
    from darwin_b200 import abi
    kind = "reference" if oracle.have_reference() else "port"
    cores = threads or os.cpu_count() or 1
    cells_per_tile = TILE * TILE
    if kind == "reference":
        ref = oracle.reference("as-is")
        ref.set_scoring(abi.Scoring.from_values())
        # ~0.2 GCUPS/core: size the sample for `seconds_target`
        n = int(min(len(req), max(cores * 8, seconds_target * cores * 0.2e9 / cells_per_tile)))
        sample = np.ascontiguousarray(req[:n])
        _, _, secs = ref.tiles(arena, sample, 1, tb_words_per_req=44, threads=cores)
    else:
        port = oracle.port(abi.Scoring.from_values())
        cores = 1
        n = int(min(len(req), max(8, seconds_target * 0.02e9 / cells_per_tile)))
        sample = np.ascontiguousarray(req[:n])
        t0 = time.time()
        port.tiles(arena, sample, 1, oracle.Port.STREAM, tb_words_per_req=44)
        secs = time.time() - t0
    gcups = n * cells_per_tile / secs / 1e9
    return {"value": gcups, "unit": "GCUPS", "cores": cores, "kind": kind,
            "sample": "%d of the workload's %dx%d tiles, %.1f s wall, %s" % (
                n, TILE, TILE, secs, "oracle/_ref BatchAlignmentSIMD (AVX2), std::thread x cores" if kind == "reference"
                else "oracle/gact_oracle.c scalar port"),
            "tiles_per_s": n / secs, "sample_tiles": n, "sample_ms": secs * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    arena, req = make_workload(min(args.tiles, 200000), 1)
    per_step = max(4.0, min(30.0, 150.0 / max(1, args.steps + args.warmup)))
    vals, ms = [], []
    cb = None
    for s in range(args.warmup + args.steps):
        cb = cpu_reference_leg(arena, req, per_step)
        if s >= args.warmup:
            vals.append(cb["value"])
            ms.append(cb["sample_ms"])
    v = float(np.mean(vals))
    cb["value"] = v
    line = {"impl": "reference", "metric": "gact_gcups", "value": v, "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": float(np.mean(ms)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int16", "data": "synthetic",
            "config": {"workload": "gact_tiles_T%d_O%d" % (TILE, OVERLAP), "tile_size": TILE, "tile_overlap": OVERLAP,
                       "error_rate": 0.15, "tiles_per_step": cb["sample_tiles"],
                       "note": "bounded sample of the same tile batch per step (ms_per_step is the sample's)"},
            "cpu_baseline": cb,
            "e2e": {"value": v, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "reads_per_s_equiv": cb["tiles_per_s"] * (TILE - OVERLAP) / 2.0 / 10000.0}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--tiles", type=int, default=1000000, help="tiles per GPU per step")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extend-reads", type=int, default=16000,
                    help="secondary measurement: reads of 10 kbp through the in-kernel anchor walker (0 = skip)")
    ap.add_argument("--e2e-lanes", type=int, default=4,
                    help="host threads (handles sharing one arena replica) issuing the end-to-end steps")
    ap.add_argument("--filter-tiles", type=int, default=400000,
                    help="secondary measurement: first-tile filter candidates through darwin_gpu_filter (0 = skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import darwin_b200
    from darwin_b200 import abi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback exists for the GACT path)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    warm = max(3, args.warmup)
    arena, req = make_workload(args.tiles, 1 + rank)          # each rank owns an independent shard (weak scaling)
    n = len(req)
    cells = float(n) * TILE * TILE
    tbw = 2 * TILE // 16 + 2
    sc = abi.Scoring.from_values()
    proc = darwin_b200.Processor(len(arena), local)
    proc.InitializeScoringParameters(sc)
    proc.InitializeReferenceMemory(0, arena)

    # ---- device-resident leg (value) -------------------------------------------------------------------
    dev = torch.device("cuda", local)
    d_req = torch.from_numpy(req.view(np.uint8).reshape(n, -1)).to(dev)
    d_res = torch.zeros((n, abi.TILE_RES.itemsize), dtype=torch.uint8, device=dev)
    d_tb = torch.zeros((n, tbw), dtype=torch.int64, device=dev)
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    launches0 = proc.stats().kernel_launches
    for _ in range(warm):
        proc.BatchAlignmentSIMD_device(d_req.data_ptr(), n, d_res.data_ptr(), d_tb.data_ptr(), tbw, TILE, TILE)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    st0 = proc.stats()
    launches_t0 = st0.kernel_launches
    kernel_ms = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        proc.BatchAlignmentSIMD_device(d_req.data_ptr(), n, d_res.data_ptr(), d_tb.data_ptr(), tbw, TILE, TILE)
        kernel_ms.append(proc.stats().last_kernel_ms)          # CUDA events on the library's own stream
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    st_end = proc.stats()
    launches = st_end.kernel_launches - launches_t0
    dev_ms = float(np.sum(kernel_ms))
    t = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, wall_ms_max = float(t[0]), float(t[1])
    value = world * cells * args.steps / (dev_ms_max * 1e-3) / 1e9

    # ---- end-to-end leg: host buffers through the public call ----------------------------------------------
    # Page-locked host buffers (what a production host keeps for DMA); every step uploads the ASCII sequences,
    # sends the requests and reads every result and TB word back.
    def pinned(shape, dtype):
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        t = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        return t.numpy().view(dtype).reshape(shape)
    h_arena = pinned(arena.shape, np.uint8); h_arena[:] = arena
    h_req = pinned(req.shape, abi.TILE_REQ); h_req[:] = req
    h_res = pinned((n,), abi.TILE_RES)
    h_tb = pinned((n, tbw), np.uint64)
    e2e_steps = max(1, min(args.steps, 3))
    h2d = len(arena) + req.nbytes
    d2h = n * abi.TILE_RES.itemsize + n * tbw * 8
    proc.InitializeReferenceMemory(0, h_arena)
    proc.BatchAlignmentSIMD(h_req, 1, tbw, out=(h_res, h_tb))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        proc.InitializeReferenceMemory(0, h_arena)             # H2D of the step's sequences (ASCII) + device packing
        res, tb = proc.BatchAlignmentSIMD(h_req, 1, tbw, out=(h_res, h_tb))   # H2D requests, kernels, D2H results + TB words
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_single = world * cells * e2e_steps / (float(t[0]) * 1e-3) / 1e9

    # The same calls from two host threads, one lane (handle) each over one shared arena replica -- the reference's own
    # concurrency model (TBB tokens, main.cpp:615-624): the H2D + packing of one step overlaps the kernels
    # of the other lane's step, and so do the D2H tails.  Every step still moves all of its inputs and outputs.
    proc.close()
    L = max(1, args.e2e_lanes)
    lane_procs = [darwin_b200.Processor(L * len(arena), local)]
    lane_procs[0].InitializeScoringParameters(sc)
    lanes = [(lane_procs[0], 0, h_req, h_res, h_tb)]
    for k in range(1, L):
        lane_procs.append(darwin_b200.Processor(0, local, parent=lane_procs[0]))
        r = pinned(req.shape, abi.TILE_REQ); r[:] = req
        r["ref_bases_start_addr"] += k * len(arena); r["query_bases_start_addr"] += k * len(arena)
        lanes.append((lane_procs[k], k * len(arena), r, pinned((n,), abi.TILE_RES), pinned((n, tbw), np.uint64)))
    lane_err = []

    def lane_steps(p, off, hreq, hres, htb, k, wait_for, uploaded):
        try:
            if wait_for is not None:
                wait_for.wait()                                  # out of lockstep: this lane's copies meet the other's kernels
            for _ in range(k):
                p.InitializeReferenceMemory(off, h_arena)
                uploaded.set()
                p.BatchAlignmentSIMD(hreq, 1, tbw, out=(hres, htb))
        except Exception as e:                                   # a failed lane fails the bench, loudly
            lane_err.append(e)
            uploaded.set()

    def run_lanes(k_each):
        ev = [threading.Event() for _ in range(L)]              # lane k starts once lane k-1 has uploaded its first step
        th = [threading.Thread(target=lane_steps, args=(*lanes[k], k_each, ev[k - 1] if k else None, ev[k])) for k in range(L)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        if lane_err:
            raise lane_err[0]

    ref_res, ref_tb = h_res.copy(), h_tb.copy()
    run_lanes(1)                                                 # warm both lanes; both must reproduce the serial results
    used = (ref_res["total_TB_pointers"].astype(np.int64) + 31) // 32
    mask = np.arange(tbw)[None, :] < used[:, None]
    for _, _, _, hres, htb in lanes:
        if not (np.array_equal(hres, ref_res) and np.array_equal(htb[mask], ref_tb[mask])):
            raise SystemExit("bench: a lane's results differ from the single-lane results")
    k_each = max(4, 2 * args.steps)                              # per lane; a step is ~0.08 s, the ramp and drain of the lanes ~0.05 s
    e2e_steps = L * k_each
    barrier()
    t0 = time.perf_counter()
    run_lanes(k_each)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * cells * e2e_steps / (float(t[0]) * 1e-3) / 1e9
    for lp in reversed(lane_procs):
        lp.close()
    proc = darwin_b200.Processor(len(arena), local)
    proc.InitializeScoringParameters(sc)
    sampler.stop_flag = True
    sampler.join(timeout=3)

    # ---- secondary: whole anchors through extender_body (in-kernel tile walking), stock params.cfg T=384/O=64 ----
    # Every rank extends its own shard of reads against its own replica of the packed reference (no
    # data-path collective); reads/s = all ranks' reads / slowest rank.
    extend_info = None
    seed_info = None
    align_info = None
    if args.extend_reads > 0:
        from darwin_b200 import synth
        ref_len = 4000000
        ex_arena, ex_anchors, ex_hits = synth.anchor_batch(7 + rank, args.extend_reads, 10000, ref_len)
        ex = darwin_b200.Processor(len(ex_arena), local)
        ex.InitializeScoringParameters(sc)
        ex.InitializeReferenceMemory(0, ex_arena)
        ex.extender_body(ex_anchors, ex_hits, 384, 64, 0)                          # warm-up at full size (buffers grown once)
        ex_out = (pinned((len(ex_anchors),), abi.ALN_RES), pinned((int(ex_anchors["read_len"].sum()) * 2,), np.uint8))
        barrier()
        t0 = time.perf_counter()
        ex_res, ex_ops = ex.extender_body(ex_anchors, ex_hits, 384, 64, 0, out=ex_out)   # anchors + hits H2D, ops D2H
        ex_wall = time.perf_counter() - t0
        ex_st = ex.stats()
        ex_cells = float(ex_res["cells"].sum())
        # D-SOFT on the GPU for the same reads: seed position table build + seeding of both strands
        chroms = np.zeros(1, abi.CHROM)
        chroms["start"], chroms["len_unpadded"] = 128, ref_len
        t0 = time.perf_counter()
        ex.build_seed_index(abi.SeedParams.stock(), chroms, 128 + ref_len + ((-ref_len) % 128))
        ix_s = time.perf_counter() - t0
        sreads = np.zeros(len(ex_anchors), abi.SEED_READ)
        sreads["read_addr"], sreads["read_len"] = ex_anchors["read_addr"], ex_anchors["read_len"]
        ex.seeder_body(sreads)                                                     # warm-up at full size
        barrier()
        t0 = time.perf_counter()
        sb, sa, sp = ex.seeder_body(sreads)
        seed_wall = time.perf_counter() - t0
        seed_ms = ex.stats().last_kernel_ms
        # the whole reference-guided pipeline in one resident call (darwin_gpu_align_reads): D-SOFT, first tiles, slope filter,
        # extension of every surviving location -- reads in, alignments (coordinates, score, op strings) out
        al_out = (pinned((len(ex_anchors) * 8,), abi.ANCHOR), pinned((len(ex_anchors) * 8,), abi.ALN_RES),
                  pinned((int(ex_anchors["read_len"].sum()) * 3,), np.uint8))
        ex.align_reads(sreads, out=al_out)                                         # warm-up at full size
        barrier()
        t0 = time.perf_counter()
        al_anchors, al_res, _ = ex.align_reads(sreads, out=al_out)
        al_wall = time.perf_counter() - t0
        al_ms = ex.stats().last_kernel_ms
        al_n = float(int((al_res["flags"] & 1).sum()))
        al_cells = float(al_res["cells"].sum())
        agg = torch.tensor([ex_st.last_kernel_ms, ex_wall * 1e3, seed_ms, seed_wall * 1e3, al_ms, al_wall * 1e3], dtype=torch.float64, device=dev)
        tot = torch.tensor([ex_cells, float(int((ex_res["flags"] & 1).sum())), float(int(ex_res["n_tiles"].sum())), float(len(sa)),
                            al_n, al_cells], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(agg, op=dist.ReduceOp.MAX)
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        k_ms, w_ms, sk_ms, sw_ms, ak_ms, aw_ms = [float(x) for x in agg]
        cells_all, aligned_all, tiles_all, seed_anchors_all, al_n_all, al_cells_all = [float(x) for x in tot]
        reads_all = args.extend_reads * world
        extend_info = {"workload": "extend_10kbp_T384_O64", "reads": int(reads_all), "reads_per_gpu": int(args.extend_reads),
                       "aligned": int(aligned_all), "tiles": int(tiles_all), "cells": cells_all, "kernel_ms": k_ms,
                       "gcups_kernel": cells_all / (k_ms * 1e-3) / 1e9,
                       "reads_per_s_kernel": reads_all / (k_ms * 1e-3),
                       "reads_per_s_e2e": reads_all / (w_ms * 1e-3),
                       "note": "one anchor per read at its true locus, synthetic chained hits; every rank extends its own shard "
                               "(max over ranks); D-SOFT not included"}
        seed_info = {"workload": "dsoft_10kbp_k14_w3", "reads": int(reads_all), "index_build_s": ix_s, "reference_bp": ref_len,
                     "kernel_ms": sk_ms, "reads_per_s_kernel": reads_all / (sk_ms * 1e-3), "reads_per_s_e2e": reads_all / (sw_ms * 1e-3),
                     "anchors": int(seed_anchors_all),
                     "note": "both strands of every read: minimizers, table look-ups, bin counting, candidates + chained hits "
                             "(darwin_gpu_seed); e2e includes the D2H of all chained hits"}
        align_info = {"workload": "align_reads_10kbp_stock_params", "reads": int(reads_all), "alignments": int(al_n_all),
                      "cells": al_cells_all, "kernel_ms": ak_ms, "reads_per_s_kernel": reads_all / (ak_ms * 1e-3),
                      "reads_per_s_e2e": reads_all / (aw_ms * 1e-3), "gcups_e2e": al_cells_all / (aw_ms * 1e-3) / 1e9,
                      "note": "resident reads in, alignments out through ONE call per rank: D-SOFT + first-tile filter + slope "
                              "filter + GACT extension on the GPU (darwin_gpu_align_reads); e2e includes the D2H of all op strings"}
        ex.close()

    # ---- secondary: first-tile filter (128x128 score-only, max-cell mode; filter.cpp:28-122) through darwin_gpu_filter ----
    filter_info = None
    if args.filter_tiles > 0 and rank == 0:
        from darwin_b200 import synth
        fa, freq = synth.tile_batch_fast(5, args.filter_tiles, 128, mode="filter")
        fc = np.zeros(len(freq), abi.FILTER_CAND)
        # one candidate per tile: a 256-base "chromosome" + 128-base "read" whose first tile is exactly the generated pair
        fc["chr_start"] = freq["ref_bases_start_addr"]; fc["chr_len"] = 256; fc["hit"] = freq["ref_bases_start_addr"]
        fc["read_addr"] = freq["query_bases_start_addr"]; fc["read_len"] = 128; fc["offset"] = 0
        fc["strand"] = (np.arange(len(fc)) & 1).astype(np.uint8)
        fp = darwin_b200.Processor(len(fa), local)
        fp.InitializeScoringParameters(sc)
        fp.InitializeReferenceMemory(0, fa)
        fp.filter_body(fc[:1024])
        t0 = time.perf_counter()
        fres = fp.filter_body(fc)                                     # candidates H2D, 3 kernels, results D2H
        fwall = time.perf_counter() - t0
        fms = fp.stats().last_kernel_ms
        filter_info = {"workload": "first_tile_128x128_score_only", "tiles": int(len(fc)), "kernel_ms": fms,
                       "gcups_kernel": len(fc) * 16384 / (fms * 1e-3) / 1e9, "tiles_per_s_kernel": len(fc) / (fms * 1e-3),
                       "tiles_per_s_e2e": len(fc) / fwall, "pass_rate": float((fres["flags"] & 1).mean()),
                       "packed_tiles": int(fp.stats().tiles_filter)}
        fp.close()

    # checksum of the last step (guards against "fast because wrong"): every tile must have produced a path
    assert int((res["total_TB_pointers"] > 0).sum()) > 0.99 * n, "tiles without traceback"

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        int_peak, int_detail = None, None
        try:
            int_detail = proc.int_peak()
            int_peak = 2.0 * max(int_detail[:3])      # packed s16x2 ops only (alu pipe); IADD3/LOP3/IMAD are reported
        except Exception:
            pass
        per_gpu_gcups = value / world
        roof = {"bound": "int_alu", "achieved": per_gpu_gcups * OPS_PER_CELL, "peak": int_peak,
                "unit": "Gint-op/s", "frac": (per_gpu_gcups * OPS_PER_CELL / int_peak) if int_peak else None,
                # dram__bytes_read+write of tiles_kernel<5> from the committed ncu capture (profiles/r1_s4_tiles_kernel5_*:
                # 76.6 MB read + 12.6 MB written per 200k-tile launch = 446 B per tile), scaled to this launch's tile count;
                # algorithmic bytes: ~460 B per tile (320 B packed bases + 32 B request + 16 B result + the used TB words)
                "traffic": 446.0 * n,
                "peak_detail_glaneops": dict(zip(["vimnmx_u16x2", "viaddmnmx_u16x2", "vimnmx3_u16x2", "iadd3", "lop3_3reg", "imad", "lop3_2reg", "lop3_imm", "prmt", "shfl_idx"], int_detail)) if int_detail else None,
                "note": "SURVEY 8(d) integer-pipe roofline: 32 algorithmic int-ops per cell; peak = measured packed "
                        "s16x2 DPX/ALU issue rate (2 cells per lane-op) of this GPU; HBM is not the bound "
                        "(%.3f B/cell algorithmic)" % ((2 * TILE / 2 + 32 + 16 + tbw * 8) / (TILE * TILE)),
                "hbm_frac_of_measured": (n * (2 * TILE / 2 + 32 + 16 + tbw * 8) / (np.mean(kernel_ms) * 1e-3) / 1e9) /
                peaks.get("hbm_gbs", 6650.0)}
        line = {"metric": "gact_gcups", "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": warm,
                "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "int16", "data": "synthetic",
                "config": {"workload": "gact_tiles_T%d_O%d" % (TILE, OVERLAP), "tiles_per_gpu": n, "tile_size": TILE,
                           "tile_overlap": OVERLAP, "error_rate": 0.15, "cells_per_step": cells * world,
                           "l2": "inputs+outputs per step (%d MB) exceed the 126 MB L2" % ((len(arena) // 2 + req.nbytes + d2h) >> 20)},
                "tiles_per_s": world * n * args.steps / (dev_ms_max * 1e-3),
                "reads_per_s_equiv": world * n * args.steps / (dev_ms_max * 1e-3) * (TILE - OVERLAP) / 2.0 / 10000.0,
                "wall_ms_per_step": wall_ms_max / args.steps,
                "clocks": sampler.summary(),
                "e2e": {"value": e2e_value, "unit": "GCUPS", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "steps": e2e_steps, "lanes": L, "single_lane_value": e2e_single,
                        "note": "one host thread per lane, one handle each over one arena replica (the reference's token model); "
                                "every step uploads its sequences and requests and reads all results + TB words back"},
                "gpu_launches": int(launches), "roofline": roof,
                "tiles": {"fast": int(st_end.tiles_fast - st0.tiles_fast), "exact": int(st_end.tiles_exact - st0.tiles_exact),
                          "rerun": int(st_end.tiles_rerun - st0.tiles_rerun)}}
        if extend_info:
            line["extend"] = extend_info
        if filter_info:
            line["filter"] = filter_info
        if seed_info:
            line["seed"] = seed_info
        if align_info:
            line["align"] = align_info
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_reference_leg(arena, req, args.cpu_seconds)
        print(json.dumps(line))
    proc.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
