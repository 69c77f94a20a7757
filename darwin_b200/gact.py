"""Python binding of libdarwin_gact.so (the C-ABI in include/darwin_gpu.h).

The names mirror the reference's Processor seam (software/Processor.h:50-63) and its extender stage
(software/graph.h:219-229): `Processor.InitializeScoringParameters`, `.InitializeReferenceMemory`,
`.InitializeReadMemory`, `.BatchAlignmentSIMD` and `.extender_body`.  There is no CPU fallback: if the
CUDA library is missing or no device is usable, construction raises.
"""
import ctypes as C
import os

import numpy as np

from . import abi

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libdarwin_gact.so")
_lib = None


class DarwinGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("darwin_gpu error %d: %s" % (code, msg))
        self.code = code


def load_library():
    """Load the in-tree CUDA library; raises if it has not been built (python __graft_entry__.py)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise FileNotFoundError(_LIB_PATH + " is missing: run `python __graft_entry__.py` (nvcc, sm_100a). "
                                    "There is no CPU fallback for the GACT path.")
        L = C.CDLL(os.environ.get("DARWIN_GPU_LIB", _LIB_PATH))      # override: A/B of differently built libraries (scripts/)
        L.darwin_gpu_version.restype = C.c_char_p
        L.darwin_gpu_last_error.restype = C.c_char_p
        L.darwin_gpu_last_error.argtypes = [C.c_void_p]
        L.darwin_gpu_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_uint64]
        L.darwin_gpu_create_shared.argtypes = [C.POINTER(C.c_void_p), C.c_void_p]
        L.darwin_gpu_destroy.argtypes = [C.c_void_p]
        L.darwin_gpu_set_scoring.argtypes = [C.c_void_p, C.POINTER(abi.Scoring)]
        L.darwin_gpu_upload.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        L.darwin_gpu_upload_spans.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.darwin_gpu_tiles.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.darwin_gpu_tiles_device.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.darwin_gpu_extend.argtypes = [C.c_void_p, C.POINTER(abi.ExtendParams), C.c_void_p, C.c_int, C.c_void_p,
                                        C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64]
        L.darwin_gpu_filter.argtypes = [C.c_void_p, C.POINTER(abi.FilterParams), C.c_void_p, C.c_int, C.c_void_p]
        L.darwin_gpu_seed_index.argtypes = [C.c_void_p, C.POINTER(abi.SeedParams), C.c_void_p, C.c_int, C.c_uint64]
        L.darwin_gpu_seed_index_share.argtypes = [C.c_void_p, C.c_void_p]
        L.darwin_gpu_seed_index_read.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                                                 C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
        L.darwin_gpu_seed.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64),
                                      C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        L.darwin_gpu_align_reads.argtypes = [C.c_void_p, C.POINTER(abi.AlignParams), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                             C.c_uint64, C.POINTER(C.c_uint64), C.c_void_p, C.c_uint64]
        L.darwin_gpu_stats.argtypes = [C.c_void_p, C.POINTER(abi.GpuStats)]
        L.darwin_gpu_cigar.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        L.darwin_gpu_sam_select.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
        L.darwin_gpu_int_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.darwin_gpu_extend_slots.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        _lib = L
    return _lib


EXPORTS = ("darwin_gpu_create", "darwin_gpu_create_shared", "darwin_gpu_destroy", "darwin_gpu_set_scoring", "darwin_gpu_upload",
           "darwin_gpu_tiles", "darwin_gpu_tiles_device", "darwin_gpu_extend", "darwin_gpu_filter", "darwin_gpu_seed_index",
           "darwin_gpu_seed_index_share", "darwin_gpu_seed_index_read", "darwin_gpu_seed", "darwin_gpu_align_reads", "darwin_gpu_host_alloc", "darwin_gpu_host_free", "darwin_gpu_upload_spans", "darwin_gpu_stats", "darwin_gpu_int_peak", "darwin_gpu_extend_slots",
           "darwin_gpu_last_error", "darwin_gpu_version", "darwin_gpu_cigar", "darwin_gpu_sam_select")


class Processor:
    """One GPU-backed Processor (== one `token` of the reference, main.cpp:615-624)."""

    def __init__(self, arena_bytes, device=0, parent=None):
        """`parent`: another Processor of the same device whose arena replica this one shares (a further lane)."""
        self.lib = load_library()
        self.h = C.c_void_p()
        if parent is not None:
            rc = self.lib.darwin_gpu_create_shared(C.byref(self.h), parent.h)
            arena_bytes, device = parent.arena_bytes, parent.device
        else:
            rc = self.lib.darwin_gpu_create(C.byref(self.h), int(device), C.c_uint64(int(arena_bytes)))
        if rc:
            msg = "no usable CUDA device" if rc == abi.ERR_NO_DEVICE else self.lib.darwin_gpu_last_error(None).decode()
            self.h = None                          # a failed create leaves no handle behind
            raise DarwinGpuError(rc, msg)
        self.arena_bytes = int(arena_bytes)
        self.device = device
        self._parent = parent                  # keeps the arena owner alive

    def close(self):
        if getattr(self, "h", None):
            self.lib.darwin_gpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise DarwinGpuError(rc, self.lib.darwin_gpu_last_error(self.h).decode())

    # g_InitializeScoringParameters (Processor.cpp:48-80)
    def InitializeScoringParameters(self, scoring):
        self._check(self.lib.darwin_gpu_set_scoring(self.h, C.byref(scoring)))

    # g_InitializeReferenceMemory / g_InitializeReadMemory (Processor.cpp:82-85, sender.cpp:4-97)
    def InitializeReferenceMemory(self, arena_addr, ascii_bytes):
        a = ascii_bytes if isinstance(ascii_bytes, np.ndarray) else np.frombuffer(ascii_bytes, np.uint8)
        a = np.ascontiguousarray(a)
        self._check(self.lib.darwin_gpu_upload(self.h, C.c_uint64(int(arena_addr)), abi.ptr(a), C.c_uint64(a.size)))

    InitializeReadMemory = InitializeReferenceMemory

    def upload_spans(self, spans):
        """Several uploads in one call: `spans` = [(arena_addr, uint8 array), ...] (darwin_gpu_upload_spans)."""
        keep = [np.ascontiguousarray(a if isinstance(a, np.ndarray) else np.frombuffer(a, np.uint8)) for _, a in spans]
        rec = np.zeros(len(spans), np.dtype([("arena_addr", "<u8"), ("ascii", "<u8"), ("n", "<u8")]))
        for k, ((addr, _), a) in enumerate(zip(spans, keep)):
            rec[k] = (int(addr), a.ctypes.data, a.size)
        self._check(self.lib.darwin_gpu_upload_spans(self.h, abi.ptr(rec), len(rec)))

    # g_BatchAlignmentSIMD (Processor.cpp:718-762)
    def BatchAlignmentSIMD(self, requests, do_traceback=1, tb_words_per_req=None, out=None):
        """`out=(res, tb)` lets the caller supply (page-locked) result buffers; otherwise they are allocated."""
        req = requests if (isinstance(requests, np.ndarray) and requests.dtype == abi.TILE_REQ and
                           requests.flags["C_CONTIGUOUS"]) else np.ascontiguousarray(requests, dtype=abi.TILE_REQ)
        n = len(req)
        if tb_words_per_req is None:
            tb_words_per_req = (int(req["max_tb_steps"].max()) // 16 + 2) if n else 1
        if out is not None:
            res, tb = out
            assert res.dtype == abi.TILE_RES and len(res) >= n
            assert (not do_traceback) or (tb.dtype == np.uint64 and tb.shape[0] >= n and tb.shape[1] == tb_words_per_req)
        else:
            res = np.empty(n, abi.TILE_RES)
            tb = np.empty((n, tb_words_per_req), np.uint64) if do_traceback else None
        self._check(self.lib.darwin_gpu_tiles(self.h, int(do_traceback), abi.ptr(req), n, abi.ptr(res),
                                              abi.ptr(tb) if tb is not None else None, int(tb_words_per_req)))
        return res, tb

    def BatchAlignmentSIMD_device(self, d_req, n, d_res, d_tb, tb_words_per_req, max_ref_size, max_query_size,
                                  do_traceback=1):
        """Device-resident variant (bench `value` leg): raw device pointers (ints)."""
        self._check(self.lib.darwin_gpu_tiles_device(self.h, int(do_traceback), C.c_void_p(d_req), int(n),
                                                     C.c_void_p(d_res), C.c_void_p(d_tb), int(tb_words_per_req),
                                                     int(max_ref_size), int(max_query_size)))

    # SeedPosTable construction (seed_pos_table.cpp:41-160) from the chromosomes already in the arena
    def build_seed_index(self, params, chroms, reference_size):
        ch = np.ascontiguousarray(chroms, dtype=abi.CHROM)
        self._check(self.lib.darwin_gpu_seed_index(self.h, C.byref(params), abi.ptr(ch), len(ch), C.c_uint64(int(reference_size))))

    def seed_index_arrays(self):
        """(buckets, positions, max_occ) copied back from the device (tests)."""
        nb, npos, mo = C.c_uint64(0), C.c_uint64(0), C.c_uint32(0)
        self._check(self.lib.darwin_gpu_seed_index_read(self.h, None, 0, None, 0, C.byref(nb), C.byref(npos), C.byref(mo)))
        b = np.zeros(nb.value + 1, np.uint32)
        p = np.zeros(max(npos.value, 1), np.uint32)
        self._check(self.lib.darwin_gpu_seed_index_read(self.h, abi.ptr(b), C.c_uint64(len(b)), abi.ptr(p), C.c_uint64(len(p)), None, None, None))
        return b, p[:npos.value], mo.value

    # seeder_body::operator() (seeder.cpp:6-55) -> SeedPosTable::DSOFT (seed_pos_table.cpp:252-553) for resident reads
    def seeder_body(self, reads):
        """reads: SEED_READ array.  Returns (anchor_begin[2n+1], anchors, pool): anchors of read r, strand s are
        anchors[anchor_begin[2r+s]:anchor_begin[2r+s+1]] in the reference's order."""
        rd = np.ascontiguousarray(reads, dtype=abi.SEED_READ)
        n = len(rd)
        begin = np.zeros(2 * n + 1, np.uint32)
        acap, pcap = max(64, 8 * n), max(1 << 16, 16384 * n)
        for _ in range(3):
            anchors = np.empty(acap, abi.SEED_ANCHOR)
            pool = np.empty(pcap, np.uint64)
            na, npool = C.c_uint64(0), C.c_uint64(0)
            rc = self.lib.darwin_gpu_seed(self.h, abi.ptr(rd), n, abi.ptr(begin), abi.ptr(anchors), C.c_uint64(acap), C.byref(na),
                                          abi.ptr(pool), C.c_uint64(pcap), C.byref(npool))
            if rc == abi.ERR_CAPACITY:
                acap, pcap = max(acap, int(na.value)), max(pcap, int(npool.value))
                continue
            self._check(rc)
            return begin, anchors[:na.value], pool[:npool.value]
        raise DarwinGpuError(abi.ERR_CAPACITY, "seed output capacity")

    # seeder_body -> filter_body -> extender_body of main.cpp:590-624 in one resident call
    def align_reads(self, reads, params=None, out=None):
        """reads: SEED_READ array (resident).  Returns (anchors, DarwinAlnRes, ops): the locations handed to the extension
        (anchors[i].read_num indexes `reads`; forward-strand locations first) and their alignments."""
        rd = np.ascontiguousarray(reads, dtype=abi.SEED_READ)
        prm = params or abi.AlignParams.stock()
        n = len(rd)
        cap = max(64, 8 * n)
        ops_cap = int(rd["read_len"].astype(np.int64).sum()) * 6 + 65536
        for _ in range(4):
            if out is not None:
                anchors, res, ops = out
                cap, ops_cap = min(len(anchors), len(res)), len(ops)
            else:
                anchors, res, ops = np.empty(cap, abi.ANCHOR), np.empty(cap, abi.ALN_RES), np.empty(ops_cap, np.uint8)
            n_out = C.c_uint64(0)
            rc = self.lib.darwin_gpu_align_reads(self.h, C.byref(prm), abi.ptr(rd), n, abi.ptr(anchors), abi.ptr(res), C.c_uint64(cap),
                                                 C.byref(n_out), abi.ptr(ops), C.c_uint64(ops_cap))
            if rc == abi.ERR_CAPACITY and out is None:
                if n_out.value > cap:
                    cap = int(n_out.value)
                else:
                    ops_cap *= 2
                continue
            self._check(rc)
            return anchors[:n_out.value], res[:n_out.value], ops
        raise DarwinGpuError(abi.ERR_CAPACITY, "align output capacity")

    # the tile part of filter_body::operator() (filter.cpp:28-122, :131-223) for a batch of D-SOFT candidates
    def filter_body(self, cands, first_tile_size=128, first_tile_score_threshold=60, min_overlap=1000, out=None):
        """Returns a DarwinFilterRes array (score, reference_pos, query_pos, flags) aligned with `cands`;
        the slope filter (filter.cpp:227-289) is host work (darwin_b200.hostlogic.slope_filter)."""
        cd = cands if (isinstance(cands, np.ndarray) and cands.dtype == abi.FILTER_CAND and cands.flags["C_CONTIGUOUS"]) \
            else np.ascontiguousarray(cands, dtype=abi.FILTER_CAND)
        n = len(cd)
        res = out if out is not None else np.empty(n, abi.FILTER_RES)
        assert res.dtype == abi.FILTER_RES and len(res) >= n
        prm = abi.FilterParams(int(first_tile_size), int(first_tile_score_threshold), int(min_overlap), 0)
        self._check(self.lib.darwin_gpu_filter(self.h, C.byref(prm), abi.ptr(cd), n, abi.ptr(res)))
        return res

    # extender_body::operator() (extender.cpp:9-1065) for a batch of anchors
    def extender_body(self, anchors, hit_pool, tile_size=384, tile_overlap=64, do_overlap=0, ops_cap=None, out=None):
        """Returns (DarwinAlnRes array, op pool).  `out=(res, ops)` supplies caller-owned (ideally page-locked) buffers."""
        an = anchors if (isinstance(anchors, np.ndarray) and anchors.dtype == abi.ANCHOR and anchors.flags["C_CONTIGUOUS"]) \
            else np.ascontiguousarray(anchors, dtype=abi.ANCHOR)
        hp = np.ascontiguousarray(hit_pool, dtype=np.uint64)
        n = len(an)
        if out is not None:
            res, ops = out
            assert res.dtype == abi.ALN_RES and len(res) >= n and ops.dtype == np.uint8
            ops_cap = len(ops)
        else:
            res = np.empty(n, abi.ALN_RES)
            if ops_cap is None:
                ops_cap = int(an["read_len"].astype(np.int64).sum()) * 3 + 65536
            ops = np.empty(ops_cap, np.uint8)
        prm = abi.ExtendParams(int(tile_size), int(tile_overlap), int(do_overlap), 0)
        self._check(self.lib.darwin_gpu_extend(self.h, C.byref(prm), abi.ptr(an), n,
                                               abi.ptr(hp) if len(hp) else None, C.c_uint64(len(hp)),
                                               abi.ptr(res), abi.ptr(ops), C.c_uint64(ops_cap)))
        return res, ops

    def extend_slots(self, tile_size):
        """Anchors one extension launch keeps in flight at this tile_size (darwin_gpu_extend_slots): the wave size of long reads."""
        n = C.c_int(0)
        self._check(self.lib.darwin_gpu_extend_slots(self.h, int(tile_size), C.byref(n)))
        return n.value

    def int_peak(self):
        """Measured issue rates (G lane-ops/s): VIMNMX.U16x2, VIADDMNMX.U16x2, VIMNMX3.U16x2, IADD3, LOP3 (3 regs),
        IMAD, LOP3 (2 regs), LOP3 (immediate), PRMT, SHFL.IDX."""
        out = (C.c_double * 10)()
        self._check(self.lib.darwin_gpu_int_peak(self.h, out))
        return list(out)

    def int_peak_gops(self):
        """P_int of SURVEY 8(d) in G int16-element-ops/s: best PACKED (alu-pipe) lane-op rate x 2 cells per lane-op."""
        return 2.0 * max(self.int_peak()[:3])

    def stats(self):
        s = abi.GpuStats()
        self._check(self.lib.darwin_gpu_stats(self.h, C.byref(s)))
        return s


def cigar(res_row, ops_pool, query_length):
    """CIGAR of one alignment (printer_body::AlignmentToSam, printer.cpp:236-301) from its op string (darwin_gpu_cigar)."""
    lib = load_library()
    r = np.ascontiguousarray(np.asarray(res_row, dtype=abi.ALN_RES).reshape(1))
    ops = np.ascontiguousarray(ops_pool, dtype=np.uint8)
    cap = 16 * (int(r["n_ops"][0]) + 4)
    out = np.empty(cap, np.uint8)
    n = C.c_uint64(0)
    rc = lib.darwin_gpu_cigar(abi.ptr(r), abi.ptr(ops), int(query_length), abi.ptr(out), C.c_uint64(cap), C.byref(n))
    if rc:
        raise DarwinGpuError(rc, "darwin_gpu_cigar")
    return out[:n.value].tobytes().decode()


def sam_select(anchors, res):
    """Print order and overlap suppression of printer_body::sam_printer (printer.cpp:15-47): (order, keep)."""
    lib = load_library()
    an = np.ascontiguousarray(anchors, dtype=abi.ANCHOR)
    rs = np.ascontiguousarray(res, dtype=abi.ALN_RES)
    order = np.zeros(max(len(rs), 1), np.uint32)
    keep = np.zeros(max(len(rs), 1), np.uint8)
    n = C.c_uint64(0)
    rc = lib.darwin_gpu_sam_select(abi.ptr(an), abi.ptr(rs), C.c_uint64(len(rs)), abi.ptr(order), abi.ptr(keep), C.byref(n))
    if rc:
        raise DarwinGpuError(rc, "darwin_gpu_sam_select")
    return order[:n.value], keep[:n.value].astype(bool)


def read_params_cfg(path):
    """params.cfg semantics of the reference (software/ConfigFile.cpp, main.cpp:183-230):
    `[section]` headers, `key = value`, `//`-style comments ignored.  Returns {section: {key: value}}."""
    out, sec = {}, None
    for raw in open(path):
        line = raw.split("//")[0].split("#")[0].strip()
        if not line:
            continue
        if line.startswith("[") and line.endswith("]"):
            sec = line[1:-1].strip()
            out.setdefault(sec, {})
        elif "=" in line and sec is not None:
            k, v = line.split("=", 1)
            out[sec][k.strip()] = v.strip()
    return out


def scoring_from_cfg(cfg):
    s = cfg["GACT_scoring"]
    return abi.Scoring(*[int(float(s[k])) for k in (
        "sub_AA", "sub_AC", "sub_AG", "sub_AT", "sub_CC", "sub_CG", "sub_CT", "sub_GG", "sub_GT", "sub_TT",
        "sub_N", "gap_open", "gap_extend", "long_gap_open", "long_gap_extend")])
