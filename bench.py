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
# ALU-pipe utilisation of the dominant kernel in the committed `ncu --set full` capture (profiles/, sm__inst_executed_pipe_alu
# .avg.pct_of_peak_sustained_active); the honest ceiling figure next to roofline.frac, which the tagged-score formulation
# pushes above 1 (it needs fewer than the 32 nominal ops per cell)
NCU_ALU_PIPE_BUSY = {"value": 81.6, "source": "profiles/r2_final_tiles_kernel_r160_5_ncu_summary.txt"}


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons of one GPU during the timed regions (B200_PROFILING.md), in-process through
    NVML (no fork per sample: eight ranks forking nvidia-smi at 10 Hz cost the round-1 end-to-end legs host time)."""
    REASONS = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"), ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
               ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"), ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap"))

    def __init__(self, index, period=0.25):
        super().__init__(daemon=True)
        self.index, self.period, self.samples, self.reasons, self.max_mhz, self.stop_flag = index, period, [], set(), None, False
        self.nvml, self.handle, self.how = None, None, "nvml"
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            # torchrun leaves CUDA_VISIBLE_DEVICES alone on this box: CUDA ordinal == NVML index; honour a mask if there is one
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [x for x in vis.split(",") if x.strip() != ""]
                if index < len(ids) and ids[index].strip().isdigit():
                    phys = int(ids[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml, self.how = None, "nvidia-smi"

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout
        f = [x.strip() for x in out.strip().split(",")]
        self.samples.append(float(f[0]))
        self.max_mhz = float(f[1])
        for (n, _), v in zip(self.REASONS, f[2:6]):
            if v.lower().startswith("active"):
                self.reasons.add(n)

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is None:
                    self._sample_smi()
                else:
                    nv = self.nvml
                    self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                    bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                    for n, attr in self.REASONS:
                        if bits & int(getattr(nv, attr)):
                            self.reasons.add(n)
            except Exception:
                pass
            time.sleep(self.period if self.nvml is not None else 1.0)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples), "how": self.how}


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
        _, _, secs = ref.tiles(arena, sample, 1, tb_words_per_req=22, threads=cores)
    else:
        port = oracle.port(abi.Scoring.from_values())
        cores = 1
        n = int(min(len(req), max(8, seconds_target * 0.02e9 / cells_per_tile)))
        sample = np.ascontiguousarray(req[:n])
        t0 = time.time()
        port.tiles(arena, sample, 1, oracle.Port.STREAM, tb_words_per_req=22)
        secs = time.time() - t0
    gcups = n * cells_per_tile / secs / 1e9
    return {"value": gcups, "unit": "GCUPS", "cores": cores, "kind": kind,
            "sample": "%d of the workload's %dx%d tiles, %.1f s wall, %s" % (
                n, TILE, TILE, secs, "oracle/_ref BatchAlignmentSIMD (AVX2), std::thread x cores, posix_memalign in place of tbbmalloc" if kind == "reference"
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


# =====================================================================================================================
# Read-level legs: BASELINE.json configs[2] (10 Mbp + 10 k PacBio-like reads, 1 GPU), configs[3] (250 Mbp + a FIXED set of
# 200 k reads, strong-sharded over the ranks, reference replicated), configs[4] (50 kbp ONT-like reads, tile_size sweep,
# de novo overlap mode).  Every timed pass starts from page-locked HOST reads: read upload (ASCII H2D + packing), D-SOFT,
# first tiles, slope filter, extension and the D2H of locations / results / op strings are all inside the timer.
# =====================================================================================================================
PACBIO = (0.015, 0.09, 0.045)      # sub / ins / del, 15 %
ONT = (0.04, 0.03, 0.05)           # 12 %


def _pinned(shape, dtype):
    import torch
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    t = torch.empty(max(nbytes, 16), dtype=torch.uint8, pin_memory=True)
    return t.numpy()[:nbytes].view(dtype).reshape(shape)


class ReadLeg:
    """One reference replica + this rank's reads, two lanes (host threads, one handle each over the shared arena replica
    and seed position table) that take chunks of reads alternately: the upload and the host part of one chunk overlap the
    kernels of the other lane's chunk."""

    def __init__(self, local, sc, genome_len, n_total, lo, hi, read_len, err, seed, chunk_reads, lanes=2, self_reference=False):
        import torch
        import darwin_b200
        from darwin_b200 import abi, workloads
        self.abi = abi
        dev = torch.device("cuda", local)
        self.case = c = workloads.ReadSetCase(dev, genome_len, n_total, lo, hi, read_len, err, seed)
        self.chunk = max(1, min(chunk_reads, c.n))
        self.n = c.n
        self.self_reference = self_reference
        if self_reference:
            # de novo mode (software/README.md:26): the reads file is its own reference -- every read is a chromosome of the
            # table AND a read; arena = [128 N][reads as chromosomes][reads]
            self.ref_end = abi_ref_end = 128 + c.n * c.stride
            self.arena_bytes = abi_ref_end + c.n * c.stride + 128
            chroms = np.zeros(c.n, abi.CHROM)
            chroms["start"] = 128 + np.arange(c.n, dtype=np.uint64) * c.stride
            chroms["len_unpadded"] = read_len
            self.chroms = chroms
            self.seed_reads = c.seed_reads.copy()
            self.seed_reads["read_addr"] = abi_ref_end + np.arange(c.n, dtype=np.uint64) * np.uint64(c.stride)
        else:
            self.ref_end, self.arena_bytes, self.chroms, self.seed_reads = c.ref_end, c.arena_bytes, c.chroms, c.seed_reads
        self.procs = [darwin_b200.Processor(self.arena_bytes, local)]
        self.procs[0].InitializeScoringParameters(sc)
        t0 = time.perf_counter()
        if self_reference:
            self.procs[0].InitializeReferenceMemory(0, np.full(128, ord("N"), np.uint8))
            self.procs[0].InitializeReferenceMemory(128, c.reads_numpy(0, c.n))
        else:
            self.procs[0].InitializeReferenceMemory(0, c.ref_numpy())
        self.ref_upload_s = time.perf_counter() - t0
        for _ in range(1, max(1, lanes)):
            self.procs.append(darwin_b200.Processor(0, local, parent=self.procs[0]))
        self.index_s = None
        self.out = []
        per_read = 3 if not self_reference else 24                    # locations per read the output buffers are sized for
        for _ in self.procs:
            cap = self.chunk * per_read + 64
            self.out.append((_pinned((cap,), abi.ANCHOR), _pinned((cap,), abi.ALN_RES),
                             _pinned((int(self.chunk * (per_read if self_reference else 2.6) * read_len * 1.1) + 65536,), np.uint8)))

    def build_index(self, do_overlap=0):
        t0 = time.perf_counter()
        ref_size = self.ref_end
        self.procs[0].build_seed_index(self.abi.SeedParams.stock(do_overlap), self.chroms, ref_size)
        self.index_s = time.perf_counter() - t0

    def one_pass(self, params, n_reads=None, lanes=None):
        """All reads of the shard (or the first n_reads) through upload + darwin_gpu_align_reads, chunk by chunk.
        Returns dict(wall_s, kernel_ms, seed_ms, filter_ms, extend_ms, alignments, locations, cells, ops, h2d, d2h)."""
        n = self.n if n_reads is None else min(self.n, n_reads)
        chunks = [(a, min(n, a + self.chunk)) for a in range(0, n, self.chunk)]
        tot = {"kernel_ms": 0.0, "seed_ms": 0.0, "filter_ms": 0.0, "extend_ms": 0.0, "alignments": 0, "locations": 0, "cells": 0.0,
               "ops": 0, "h2d": 0, "d2h": 0, "score_sum": 0}
        lock = threading.Lock()
        errs = []
        nxt = [0]

        def lane(k):
            p, (o_an, o_res, o_ops) = self.procs[k], self.out[k]
            try:
                while True:
                    with lock:
                        i = nxt[0]
                        nxt[0] += 1
                    if i >= len(chunks):
                        return
                    a, b = chunks[i]
                    ascii_ = self.case.reads_numpy(a, b)
                    p.InitializeReadMemory(int(self.seed_reads["read_addr"][a]), ascii_)          # H2D + packing
                    an, res, _ = p.align_reads(self.seed_reads[a:b], params, out=(o_an, o_res, o_ops))
                    st = p.stats()
                    em = (res["flags"] & 1) != 0
                    nops = int(res["n_ops"][em].sum())
                    with lock:
                        tot["kernel_ms"] += st.last_kernel_ms; tot["seed_ms"] += st.last_seed_ms
                        tot["filter_ms"] += st.last_filter_ms; tot["extend_ms"] += st.last_extend_ms
                        tot["alignments"] += int(em.sum()); tot["locations"] += len(res)
                        tot["cells"] += float(res["cells"].sum()); tot["ops"] += nops
                        tot["score_sum"] += int(res["score"][em].sum())
                        tot["h2d"] += ascii_.nbytes + (b - a) * 16
                        tot["d2h"] += len(res) * (64 + 56) + nops
            except Exception as e:
                errs.append(e)

        t0 = time.perf_counter()
        th = [threading.Thread(target=lane, args=(k,)) for k in range(min(lanes or len(self.procs), len(self.procs)))]
        for x in th:
            x.start()
        for x in th:
            x.join()
        tot["wall_s"] = time.perf_counter() - t0
        if errs:
            raise errs[0]
        tot["reads"] = n
        return tot

    def close(self):
        for p in reversed(self.procs):
            p.close()
        self.procs = []


def _agg(dist, dev, world, tot, keys_max, keys_sum):
    import torch
    a = torch.tensor([tot[k] for k in keys_max], dtype=torch.float64, device=dev)
    b = torch.tensor([float(tot[k]) for k in keys_sum], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(a, op=dist.ReduceOp.MAX)
        dist.all_reduce(b, op=dist.ReduceOp.SUM)
    out = {k: float(v) for k, v in zip(keys_max, a)}
    out.update({k: float(v) for k, v in zip(keys_sum, b)})
    return out


def run_read_leg(name, note, local, rank, world, dist, barrier, sc, int_peak, genome_len, n_total, strong, read_len, err, seed,
                 chunk_reads, tiles, passes=1, lanes=2, self_reference=False, do_overlap=0, wave_sized=False):
    """tiles: list of (tile_size, tile_overlap) run on the same resident case; returns {name or name_T<T>: info}."""
    import torch
    from darwin_b200 import abi, workloads
    dev = torch.device("cuda", local)
    if strong:
        per = -(-n_total // world)
        per += (-per) % workloads.BLOCK
        lo, hi = min(n_total, rank * per), min(n_total, (rank + 1) * per)
    else:
        lo, hi = 0, n_total
        seed = seed + 1000 * rank                                       # weak scaling: every rank owns its own read set
    if strong:
        # strong scaling: keep at least four chunks per rank (two per lane) so that a rank's uploads, host work and op D2H still
        # overlap the other lane's kernels when its shard gets small; never below what fills the device a few times over
        chunk_reads = max(4000, min(chunk_reads, -(-(hi - lo) // 4)))
    leg = ReadLeg(local, sc, genome_len, n_total, lo, hi, read_len, err, seed, chunk_reads, lanes, self_reference)
    leg.build_index(do_overlap)
    out = {}
    chunk_cap, n_target = leg.chunk, leg.n
    for (T, O) in tiles:
        prm = abi.AlignParams.stock(T, O, do_overlap)
        slots, n_run = None, leg.n
        if wave_sized:
            # few long reads: one warp per anchor, ~270 tiles each -- a launch runs in whole waves of `slots` anchors, so a chunk is
            # a little under a whole number of waves (darwin_gpu_extend_slots; 2 003 anchors on 1 776 warps took two wave times)
            slots = leg.procs[0].extend_slots(T)
            waves = max(1, -(-1500 // slots))
            leg.chunk = max(1, min(chunk_cap, int(0.96 * slots * waves)))
            per_round = leg.chunk * len(leg.procs)
            n_run = min(n_target, per_round * max(1, int(round(n_target / per_round))))
        leg.one_pass(prm, n_reads=min(n_run, 2 * leg.chunk * len(leg.procs)))       # warm-up: buffers grown, pools filled
        acc = None
        barrier()
        t0 = time.perf_counter()
        for _ in range(passes):
            t = leg.one_pass(prm, n_reads=n_run)
            acc = t if acc is None else {k: (acc[k] + t[k]) for k in acc}
        barrier()
        acc["wall_s"] = time.perf_counter() - t0
        # one more pass from ONE lane over a few chunks: kernels of different lanes share the device, so the per-stage CUDA-event
        # times above overlap; the single-lane pass gives the extension kernel's own rate (the roofline fraction of this leg)
        solo = leg.one_pass(prm, n_reads=min(n_run, 2 * leg.chunk), lanes=1)
        acc["solo_extend_ms"], acc["solo_cells"] = solo["extend_ms"], solo["cells"]
        a = _agg(dist, dev, world, acc, ["wall_s", "kernel_ms", "extend_ms", "seed_ms", "filter_ms", "solo_extend_ms"],
                 ["reads", "alignments", "locations", "cells", "ops", "h2d", "d2h", "score_sum", "solo_cells"])
        lanes_n = len(leg.procs)
        # kernel times are summed over the lanes' calls, which overlap on the device: the device-time rate uses the per-lane
        # share as a lower bound of the busy time, the end-to-end rate uses the wall clock
        gcups_ext = a["cells"] / world / (a["extend_ms"] * 1e-3) / 1e9 if a["extend_ms"] else None
        gcups_solo = a["solo_cells"] / world / (a["solo_extend_ms"] * 1e-3) / 1e9 if a["solo_extend_ms"] else None
        info = {"workload": name, "reference_bp": genome_len, "reads": int(a["reads"] / passes), "read_len": read_len,
                "error_profile_sub_ins_del": list(err), "tile_size": T, "tile_overlap": O, "do_overlap": do_overlap,
                "scaling": "strong" if strong else "weak", "passes": passes, "lanes": lanes_n, "chunk_reads": leg.chunk,
                "locations": int(a["locations"] / passes), "alignments": int(a["alignments"] / passes), "cells": a["cells"] / passes,
                "reads_per_s_e2e": a["reads"] / a["wall_s"], "gcups_e2e": a["cells"] / a["wall_s"] / 1e9,
                "wall_ms": a["wall_s"] * 1e3 / passes,
                "extend_kernel_ms_slowest_rank": a["extend_ms"] / passes, "seed_kernel_ms": a["seed_ms"] / passes,
                "filter_kernel_ms": a["filter_ms"] / passes,
                "gcups_extend_kernel_lanes_overlapped": gcups_ext,
                "gcups_extend_kernel_per_gpu": gcups_solo,
                "roofline_frac_extend_kernel": (gcups_solo * OPS_PER_CELL / int_peak) if (gcups_solo and int_peak) else None,
                "extend_kernel_note": "gcups_extend_kernel_per_gpu: single-lane pass over %d reads per rank (kernel alone on the device, slowest rank); "
                                      "..._lanes_overlapped: cells / summed CUDA-event times of the timed passes, whose lanes share the device (lower bound)" % min(n_run, 2 * leg.chunk),
                "extend_slots": slots,
                "h2d_bytes": int(a["h2d"] / passes), "d2h_bytes": int(a["d2h"] / passes),
                "index_build_s": leg.index_s, "reference_upload_s": leg.ref_upload_s, "score_checksum": int(a["score_sum"] / passes),
                "note": note}
        out[name if len(tiles) == 1 else "%s_T%d" % (name, T)] = info
    leg.close()
    del leg
    torch.cuda.empty_cache()
    return out


def cpu_reads_leg(genome_len, read_len, err, seed, T, O, n_sample, threads=None):
    """The reference's own per-read chain (seeder_body -> filter_body -> extender_body, one read per batch as main.cpp feeds
    them) on a bounded sample of the same read set, all host cores (oracle/_ref, as-is flavour)."""
    import torch
    import ctypes as C
    import oracle
    from darwin_b200 import abi, workloads
    if not oracle.have_reference():
        return None
    cores = threads or os.cpu_count() or 1
    cpu = torch.device("cpu")
    genome = workloads.genome_codes(genome_len, seed, torch.device("cuda") if torch.cuda.is_available() else cpu)
    blk = workloads.simulate_block(genome, 0, seed, read_len, err, n=min(n_sample, workloads.BLOCK)).cpu().numpy()
    ga = np.frombuffer(b"ACGT", np.uint8)[genome.cpu().numpy()]
    ref = oracle.reference("as-is")
    ref.set_scoring(abi.Scoring.from_values())
    ref.set_dsoft_defaults()
    ref.set_extend(T, O, 2, 0)
    ref.reset_arena()
    t0 = time.perf_counter()
    ref.add_chr("chrS", ga.tobytes(), True)
    ref.build_index()
    index_s = time.perf_counter() - t0
    n = len(blk)
    for k in range(n):
        ref.add_read("r%d" % k, np.ascontiguousarray(blk[k, :read_len]).tobytes())
    st = (C.c_double * 8)()
    alns = ref.lib.dref_pipeline_cpu_mt(0, n, cores, st)
    return {"kind": "reference", "cores": cores, "sample": "%d reads of the same set, seeder_body -> filter_body -> extender_body per read, "
            "std::thread x %d, %.1f s wall (index build %.1f s not counted)" % (n, cores, st[0], index_s),
            "reads_per_s": n / st[0], "gcups": st[5] / st[0] / 1e9, "alignments": int(alns),
            "stage_thread_seconds_seed_filter_extend": [st[2], st[3], st[4]], "unit": "reads/s", "value": n / st[0]}


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
    ap.add_argument("--config3-reads", type=int, default=10000,
                    help="configs[2]: 10 Mbp reference, this many PacBio-like 10 kbp reads per GPU, reads in -> alignments out (0 = skip)")
    ap.add_argument("--config4-reads", type=int, default=200000,
                    help="configs[3]: 250 Mbp replicated reference, a FIXED set of this many 10 kbp reads strong-sharded over the ranks (0 = skip)")
    ap.add_argument("--config4-genome", type=int, default=250000000)
    ap.add_argument("--config5-reads", type=int, default=4000,
                    help="configs[4]: ONT-like 50 kbp reads per GPU, tile_size 256/512/1024 + de novo overlap mode (0 = skip)")
    ap.add_argument("--read-lanes", type=int, default=2, help="host threads (lanes) feeding the read-level legs")
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
    # TB words per tile: a corner traceback takes at most Q + R = 2 * TILE steps, 32 two-bit ops per 64-bit word
    # (Processor.cpp:568-582) -> 20 words at T = 320, + 2 spare; the reference returns only the used words per tile
    tbw = 2 * TILE // 32 + 2
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

    # ---- secondary: whole anchors through extender_body (in-kernel tile walking), stock params.cfg T=384/O=64 ----
    # Every rank extends its own shard of reads against its own replica of the packed reference (no
    # data-path collective); reads/s = all ranks' reads / slowest rank.
    extend_info = None
    seed_info = None
    if args.extend_reads > 0:
        from darwin_b200 import synth
        ref_len = 4000000
        ex_arena, ex_anchors, ex_hits = synth.anchor_batch(7 + rank, args.extend_reads, 10000, ref_len)
        ex = darwin_b200.Processor(len(ex_arena), local)
        ex.InitializeScoringParameters(sc)
        ex.InitializeReferenceMemory(0, ex_arena)
        ex.extender_body(ex_anchors, ex_hits, 384, 64, 0)                          # warm-up at full size (buffers grown once)
        ex_out = (pinned((len(ex_anchors),), abi.ALN_RES), pinned((int(ex_anchors["read_len"].sum()) * 2,), np.uint8))
        reads_at = int(ex_anchors["read_addr"].min())                              # the reads follow the chromosome in the arena
        h_reads = pinned((len(ex_arena) - reads_at,), np.uint8); h_reads[:] = ex_arena[reads_at:]
        barrier()
        t0 = time.perf_counter()
        ex.InitializeReadMemory(reads_at, h_reads)                                 # the reads' ASCII H2D + packing is part of the step
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
        agg = torch.tensor([ex_st.last_kernel_ms, ex_wall * 1e3, seed_ms, seed_wall * 1e3], dtype=torch.float64, device=dev)
        tot = torch.tensor([ex_cells, float(int((ex_res["flags"] & 1).sum())), float(int(ex_res["n_tiles"].sum())), float(len(sa))],
                           dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(agg, op=dist.ReduceOp.MAX)
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        k_ms, w_ms, sk_ms, sw_ms = [float(x) for x in agg]
        cells_all, aligned_all, tiles_all, seed_anchors_all = [float(x) for x in tot]
        reads_all = args.extend_reads * world
        extend_info = {"workload": "extend_10kbp_T384_O64", "reads": int(reads_all), "reads_per_gpu": int(args.extend_reads),
                       "aligned": int(aligned_all), "tiles": int(tiles_all), "cells": cells_all, "kernel_ms": k_ms,
                       "gcups_kernel": cells_all / (k_ms * 1e-3) / 1e9,
                       "reads_per_s_kernel": reads_all / (k_ms * 1e-3),
                       "reads_per_s_e2e": reads_all / (w_ms * 1e-3),
                       "note": "one anchor per read at its true locus, synthetic chained hits; every rank extends its own shard "
                               "(max over ranks); e2e = read upload (ASCII H2D + packing) + anchors/hits H2D + kernels + op D2H; D-SOFT not included"}
        seed_info = {"workload": "dsoft_10kbp_k14_w3", "reads": int(reads_all), "index_build_s": ix_s, "reference_bp": ref_len,
                     "kernel_ms": sk_ms, "reads_per_s_kernel": reads_all / (sk_ms * 1e-3), "reads_per_s_e2e": reads_all / (sw_ms * 1e-3),
                     "anchors": int(seed_anchors_all),
                     "note": "both strands of every read: minimizers, table look-ups, bin counting, candidates + chained hits "
                             "(darwin_gpu_seed); e2e includes the D2H of all chained hits"}
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
        fp.filter_body(fc)                                            # warm-up at full size (buffers grown once)
        t0 = time.perf_counter()
        fres = fp.filter_body(fc)                                     # candidates H2D, 3 kernels, results D2H
        fwall = time.perf_counter() - t0
        fms = fp.stats().last_kernel_ms
        filter_info = {"workload": "first_tile_128x128_score_only", "tiles": int(len(fc)), "kernel_ms": fms,
                       "gcups_kernel": len(fc) * 16384 / (fms * 1e-3) / 1e9, "tiles_per_s_kernel": len(fc) / (fms * 1e-3),
                       "tiles_per_s_e2e": len(fc) / fwall, "pass_rate": float((fres["flags"] & 1).mean()),
                       "packed_tiles": int(fp.stats().tiles_filter)}
        fp.close()

    # ---- read-level legs: BASELINE.json configs[2], [3], [4] (reads in -> alignments out, host buffers, upload timed) ----
    int_peak, int_detail = None, None
    try:
        int_detail = proc.int_peak()
        int_peak = 2.0 * max(int_detail[:3])      # packed s16x2 ops only (alu pipe); IADD3/LOP3/IMAD are reported
    except Exception:
        pass
    read_legs = {}
    if args.config3_reads > 0:
        read_legs.update(run_read_leg(
            "config3", "BASELINE.json configs[2]: synthetic 10 Mbp reference + PacBio-like 10 kbp reads (15 %: sub 1.5 / ins 9 / del 4.5), "
            "stock params.cfg, reference-guided; every rank runs its own read set against its own replica",
            local, rank, world, dist, barrier, sc, int_peak, 10000000, args.config3_reads, False, 10000, PACBIO, 31,
            5000, [(384, 64)], passes=2, lanes=args.read_lanes))
    if args.config4_reads > 0:
        read_legs.update(run_read_leg(
            "config4", "BASELINE.json configs[3]: synthetic chr1-scale reference replicated per GPU + ONE fixed set of 10 kbp reads at 15 % "
            "(5/5/5) sharded contiguously over the ranks (strong scaling); reads/s = set size / slowest rank",
            local, rank, world, dist, barrier, sc, int_peak, args.config4_genome, args.config4_reads, True, 10000, (0.05, 0.05, 0.05), 41,
            12500, [(384, 64)], passes=1, lanes=args.read_lanes))
    if args.config5_reads > 0:
        read_legs.update(run_read_leg(
            "config5", "BASELINE.json configs[4]: ONT-like 50 kbp reads at 12 % (sub 4 / ins 3 / del 5) against a 20 Mbp reference, tile_overlap 64",
            local, rank, world, dist, barrier, sc, int_peak, 20000000, args.config5_reads, False, 50000, ONT, 51,
            max(1000, int(args.config5_reads * 0.6)), [(256, 64), (512, 64), (1024, 64)], passes=1, lanes=args.read_lanes, wave_sized=True))
        read_legs.update(run_read_leg(
            "config5_denovo", "BASELINE.json configs[4], de novo mode (argv[3] = 1): the read set is its own reference, all-vs-all, "
            "50 kbp ONT-like reads at ~5x coverage of a 10 Mbp genome, tile_size 256",
            local, rank, world, dist, barrier, sc, int_peak, 10000000, max(200, args.config5_reads // 4), False, 50000, ONT, 52,
            500, [(256, 64)], passes=1, lanes=args.read_lanes, self_reference=True, do_overlap=1))
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # the reference's CPU chain on bounded samples of the same read sets (reported baseline, N = 1 only)
        try:
            if "config3" in read_legs:
                read_legs["config3"]["cpu_baseline"] = cpu_reads_leg(10000000, 10000, PACBIO, 31, 384, 64, 512)
            if "config5_T512" in read_legs:
                read_legs["config5_T512"]["cpu_baseline"] = cpu_reads_leg(20000000, 50000, ONT, 51, 512, 64, 64)
        except Exception as e:                                     # the baseline is a report, never a reason to lose the bench line
            read_legs["cpu_baseline_error"] = repr(e)

    sampler.stop_flag = True                                       # clocks were sampled over every timed region above
    sampler.join(timeout=3)

    # checksum of the last step (guards against "fast because wrong"): every tile must have produced a path
    assert int((res["total_TB_pointers"] > 0).sum()) > 0.99 * n, "tiles without traceback"

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        per_gpu_gcups = value / world
        roof = {"bound": "int_alu", "achieved": per_gpu_gcups * OPS_PER_CELL, "peak": int_peak,
                "unit": "Gint-op/s", "frac": (per_gpu_gcups * OPS_PER_CELL / int_peak) if int_peak else None,
                # dram__bytes_read+write of tiles_kernel_r160<5> from the committed ncu capture of the shipped build
                # (profiles/r2_final_tiles_kernel_r160_5_*: 76.7 MB read + 11.8 MB written per 200k-tile launch = 442 B per tile;
                # the writes vary between captures with what L2 still holds at the end: 16.1 MB in r2_final_narrow_*),
                # scaled to this launch's tile count; algorithmic bytes: ~460 B per tile (320 B packed bases + 32 B request +
                # 16 B result + the used TB words)
                "traffic": 442.0 * n, "traffic_source": "constant_from_ncu (profiles/: dram__bytes_read+write per tile of the committed capture x tiles; not a live counter)",
                "alu_pipe_busy_ncu_pct": NCU_ALU_PIPE_BUSY,
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
        line.update(read_legs)
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_reference_leg(arena, req, args.cpu_seconds)
        print(json.dumps(line))
    proc.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
