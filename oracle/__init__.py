"""TEST INFRASTRUCTURE ONLY -- loaders for the parity oracles.

* ``port()``            -> oracle/libgact_oracle.so, our CPU restatement (oracle/gact_oracle.c)
* ``reference(flavour)``-> oracle/_ref/libdarwin_ref{,_patched}.so, the reference's own translation
                          units compiled unmodified (recipe: oracle/Makefile, driver: oracle/ref_driver.cpp)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this
package; the product (darwin_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from darwin_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_cache = {}


def build(target="all"):
    """(Re)build the oracle libraries; the `ref` target is a no-op when /root/reference is absent."""
    subprocess.run(["make", "-C", _HERE, target], check=True, stdout=subprocess.DEVNULL)


def _load(path):
    if path not in _cache:
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle`)")
        _cache[path] = C.CDLL(path)
    return _cache[path]


class Port:
    """oracle/gact_oracle.c through ctypes."""
    STRIPED, STREAM, CLEAN = 0, 1, 2

    def __init__(self, scoring):
        self.lib = _load(os.path.join(_HERE, "libgact_oracle.so"))
        self.sc = (C.c_int * 64)()          # GactScoring (25+4+11 ints) with slack
        self.lib.gact_scoring_init(self.sc, C.byref(scoring))
        self.lib.gact_alignment_score.restype = C.c_int

    def tiles(self, dram, req, do_traceback=1, rule=1, tb_words_per_req=None):
        n = len(req)
        res = np.zeros(n, abi.TILE_RES)
        if tb_words_per_req is None:
            tb_words_per_req = int(req["max_tb_steps"].max()) // 16 + 2 if n else 1
        tb = np.zeros((n, tb_words_per_req), np.uint64)
        flags = np.zeros(n, np.uint32)
        rc = self.lib.gact_tiles(self.sc, abi.ptr(dram), int(do_traceback), int(rule), abi.ptr(req), n,
                                 abi.ptr(res), abi.ptr(tb), tb_words_per_req, abi.ptr(flags))
        if rc:
            raise RuntimeError("gact_tiles rc=%d" % rc)
        return res, tb, flags

    def filter(self, dram, cands, first_tile_size=128, threshold=60, min_overlap=1000):
        """The tile part of filter_body (filter.cpp:28-122, :131-223) for a candidate array."""
        cd = np.ascontiguousarray(cands, dtype=abi.FILTER_CAND)
        res = np.zeros(len(cd), abi.FILTER_RES)
        prm = abi.FilterParams(int(first_tile_size), int(threshold), int(min_overlap), 0)
        rc = self.lib.gact_filter(self.sc, abi.ptr(dram), C.byref(prm), abi.ptr(cd), len(cd), abi.ptr(res))
        if rc:
            raise RuntimeError("gact_filter rc=%d" % rc)
        return res

    def slope_filter(self, read_num, score, reference_pos, query_pos, slope_threshold=0.05):
        """filter_body::slopeFilter (filter.cpp:227-289) for one strand; returns the kept indices in output order."""
        rn = np.ascontiguousarray(read_num, np.int32)
        sc = np.ascontiguousarray(score, np.int32)
        rp = np.ascontiguousarray(reference_pos, np.uint32)
        qp = np.ascontiguousarray(query_pos, np.uint32)
        order = np.zeros(max(len(rn), 1), np.int32)
        self.lib.gact_slope_filter.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p]
        n = self.lib.gact_slope_filter(abi.ptr(rn), abi.ptr(sc), abi.ptr(rp), abi.ptr(qp), len(rn), C.c_float(slope_threshold),
                                       abi.ptr(order))
        return order[:n].copy()

    def extend(self, dram, params, anchors, hit_pool, rule=1, ops_cap=None):
        n = len(anchors)
        res = np.zeros(n, abi.ALN_RES)
        if ops_cap is None:
            ops_cap = int(anchors["read_len"].astype(np.int64).sum()) * 3 + 65536
        ops = np.zeros(ops_cap, np.uint8)
        hp = hit_pool if len(hit_pool) else np.zeros(1, np.uint64)
        rc = self.lib.gact_extend(self.sc, abi.ptr(dram), C.byref(params), int(rule), abi.ptr(anchors), n,
                                  abi.ptr(hp), abi.ptr(res), abi.ptr(ops), C.c_uint64(ops_cap))
        if rc:
            raise RuntimeError("gact_extend rc=%d" % rc)
        return res, ops


class DsoftPort:
    """oracle/dsoft_oracle.c: the seed position table and SeedPosTable::DSOFT restated in C."""

    class _Index(C.Structure):
        _fields_ = [("k", C.c_int), ("w", C.c_int), ("max_stride", C.c_int), ("bin_size", C.c_uint32),
                    ("kmer_max_occurence", C.c_uint32), ("n_buckets", C.c_uint64), ("n_positions", C.c_uint64),
                    ("buckets", C.c_void_p), ("positions", C.c_void_p)]

    def __init__(self, dram, chroms, reference_size, params):
        self.lib = _load(os.path.join(_HERE, "libgact_oracle.so"))
        self.ix = self._Index()
        self.dram = np.ascontiguousarray(dram)
        self.params = params
        st = np.ascontiguousarray(chroms["start"], np.uint32)
        ln = np.ascontiguousarray(chroms["len_unpadded"], np.uint32)
        rc = self.lib.dsoft_index_build(C.byref(self.ix), abi.ptr(self.dram), abi.ptr(st), abi.ptr(ln), len(st),
                                        C.c_uint32(int(reference_size)), params.seed_size, params.minimizer_window,
                                        C.c_uint32(params.seed_occurence_multiple), C.c_uint32(params.bin_size), params.max_stride)
        if rc:
            raise MemoryError("dsoft_index_build")

    def close(self):
        if self.ix.buckets:
            self.lib.dsoft_index_free(C.byref(self.ix))

    def index_arrays(self):
        b = np.frombuffer((C.c_uint32 * (self.ix.n_buckets + 1)).from_address(self.ix.buckets), np.uint32)
        p = np.frombuffer((C.c_uint32 * max(self.ix.n_positions, 1)).from_address(self.ix.positions), np.uint32)
        return b, p[:self.ix.n_positions]

    def minimizers(self, seq):
        s = np.concatenate([np.frombuffer(seq, np.uint8), np.full(64, ord("N"), np.uint8)])
        out = np.zeros(len(s) + 32, np.uint64)
        self.lib.dsoft_minimizers.restype = C.c_uint64
        n = self.lib.dsoft_minimizers(abi.ptr(s), C.c_uint32(len(seq)), self.params.seed_size, self.params.minimizer_window, abi.ptr(out))
        return out[:n]

    def query(self, seq):
        """seq: ASCII bytes/array of one strand (padded internally with 'N' like the reader does)."""
        s = np.concatenate([np.frombuffer(seq, np.uint8) if not isinstance(seq, np.ndarray) else seq, np.full(160, ord("N"), np.uint8)])
        n_seq = len(s) - 160
        acap, pcap = 4096, 1 << 20
        while True:
            anchors = np.zeros(acap, abi.SEED_ANCHOR)
            pool = np.zeros(pcap, np.uint64)
            used = C.c_uint64(0)
            n = self.lib.dsoft_query(C.byref(self.ix), abi.ptr(s), C.c_uint32(n_seq), self.params.num_seeds, self.params.threshold,
                                     self.params.do_overlap, abi.ptr(anchors), acap, abi.ptr(pool), C.c_uint64(pcap), C.byref(used))
            if n in (-2, -3):
                acap, pcap = acap * 4, pcap * 4
                continue
            return anchors[:n].copy(), pool[:used.value].copy()


class Reference:
    """The compiled reference (oracle/_ref) through oracle/ref_driver.cpp."""

    def __init__(self, flavour="patched"):
        name = "libdarwin_ref_patched.so" if flavour == "patched" else "libdarwin_ref.so"
        self.lib = _load(os.path.join(_HERE, "_ref", name))
        L = self.lib
        L.dref_flavour.restype = C.c_char_p
        L.dref_arena.restype = C.c_void_p
        L.dref_arena_reference_size.restype = C.c_uint64
        L.dref_arena_position.restype = C.c_uint64
        L.dref_add_chr.restype = C.c_uint64
        L.dref_anchor_hits_total.restype = C.c_uint64
        L.dref_tiles_mt.restype = C.c_double
        L.dref_extend_mt.restype = C.c_double
        self.flavour = L.dref_flavour().decode()

    # -- configuration ---------------------------------------------------------------------
    def load_cfg(self, path, do_overlap=0):
        if self.lib.dref_load_cfg(path.encode(), int(do_overlap)):
            raise RuntimeError("cannot read " + path)

    def set_scoring(self, scoring):
        self.lib.dref_set_scoring(C.byref(scoring))

    def set_extend(self, tile_size, tile_overlap, batch_size=2, do_overlap=0):
        self.lib.dref_set_extend(tile_size, tile_overlap, batch_size, do_overlap)

    def set_dsoft_defaults(self):
        """software/params.cfg:18-35"""
        self.lib.dref_set_dsoft(14, 3, 64, 26, 1000, 40, 1000, 4, 128, 60, 64, 1000, C.c_float(0.05))

    # -- arena ------------------------------------------------------------------------------
    def reset_arena(self):
        if self.lib.dref_reset_arena():
            raise MemoryError("reference arena")

    def add_chr(self, name, seq, index=True):
        return self.lib.dref_add_chr(name.encode(), seq, C.c_uint64(len(seq)), int(index))

    def build_index(self):
        if self.lib.dref_build_index():
            raise RuntimeError("no minimizers collected")

    def add_read(self, name, seq):
        addr = C.c_uint64(0)
        num = self.lib.dref_add_read(name.encode(), seq, C.c_uint64(len(seq)), C.byref(addr))
        return num, addr.value

    def arena(self, nbytes=None):
        """numpy view (uint8) of the first nbytes of the reference's byte arena."""
        if nbytes is None:
            nbytes = self.lib.dref_arena_position()
        buf = (C.c_uint8 * nbytes).from_address(self.lib.dref_arena())
        return np.frombuffer(buf, np.uint8)

    # -- tiles ------------------------------------------------------------------------------
    def tiles(self, dram, req, do_traceback=1, tb_words_per_req=None, threads=0):
        n = len(req)
        res = np.zeros(n, abi.TILE_RES)
        if tb_words_per_req is None:
            tb_words_per_req = int(req["max_tb_steps"].max()) // 16 + 2 if n else 1
        tb = np.zeros((n, tb_words_per_req), np.uint64)
        d = abi.ptr(dram) if dram is not None else None
        if threads:
            secs = self.lib.dref_tiles_mt(d, int(do_traceback), abi.ptr(req), n, abi.ptr(res), abi.ptr(tb),
                                          tb_words_per_req, int(threads))
            return res, tb, secs
        rc = self.lib.dref_tiles(d, int(do_traceback), abi.ptr(req), n, abi.ptr(res), abi.ptr(tb), tb_words_per_req)
        if rc:
            raise RuntimeError("dref_tiles rc=%d" % rc)
        return res, tb

    # -- anchors / extension --------------------------------------------------------------------
    def seed_filter(self, first, count):
        n = self.lib.dref_seed_filter(int(first), int(count))
        if n < 0:
            raise RuntimeError("index not built")
        nh = self.lib.dref_anchor_hits_total()
        anchors = np.zeros(max(n, 1), abi.ANCHOR)
        hits = np.zeros(max(nh, 1), np.uint64)
        got = self.lib.dref_get_anchors(abi.ptr(anchors), n, abi.ptr(hits), C.c_uint64(nh), C.c_uint64(0))
        assert got == n
        return anchors[:n], hits[:nh]

    def _fetch_anchors(self, n):
        nh = self.lib.dref_anchor_hits_total()
        anchors = np.zeros(max(n, 1), abi.ANCHOR)
        hits = np.zeros(max(nh, 1), np.uint64)
        got = self.lib.dref_get_anchors(abi.ptr(anchors), n, abi.ptr(hits), C.c_uint64(nh), C.c_uint64(0))
        assert got == n
        return anchors[:n], hits[:nh]

    def seed(self, first, count):
        """seeder_body alone on reads [first, first+count): returns (candidates, read_num per candidate) as
        filter_body would see them (forward strand first)."""
        n = self.lib.dref_seed(int(first), int(count))
        if n < 0:
            raise RuntimeError("index not built")
        self._last_seed_count = int(count)
        cands = np.zeros(max(n, 1), abi.FILTER_CAND)
        rn = np.zeros(max(n, 1), np.int32)
        got = self.lib.dref_get_candidates(abi.ptr(cands), abi.ptr(rn), n)
        assert got == n
        return cands[:n], rn[:n]

    def _fetch_candidates(self, n):
        cands = np.zeros(max(n, 1), abi.FILTER_CAND)
        rn = np.zeros(max(n, 1), np.int32)
        got = self.lib.dref_get_candidates(abi.ptr(cands), abi.ptr(rn), n)
        assert got == n
        return cands[:n], rn[:n]

    def seed_custom(self, first, count, hit, offset, read_num, strand):
        """Hand-made seeder output (sorted by read within each strand): returns (candidates, read_num) like seed()."""
        ho = (np.asarray(hit, np.uint64) << np.uint64(32)) | np.asarray(offset, np.uint64)
        rn = np.ascontiguousarray(read_num, np.int32)
        st = np.ascontiguousarray(strand, np.uint8)
        n = self.lib.dref_seed_custom(int(first), int(count), abi.ptr(np.ascontiguousarray(ho)), abi.ptr(rn), abi.ptr(st), len(rn))
        if n != len(rn):
            raise RuntimeError("dref_seed_custom rc=%d" % n)
        return self._fetch_candidates(n)

    def set_first_tile(self, first_tile_size=128, threshold=60, min_overlap=1000, slope_threshold=0.05):
        """params.cfg [GACT_first_tile]; the D-SOFT parameters stay at the stock values (software/params.cfg:18-35)."""
        self.lib.dref_set_dsoft(14, 3, 64, 26, 1000, 40, 1000, 4, int(first_tile_size), int(threshold), 64, int(min_overlap),
                                C.c_float(slope_threshold))

    def seed_anchors(self):
        """Full output of the last seed(): (anchor_begin[2n+1], anchors, pool) in the layout of darwin_gpu_seed."""
        L = self.lib
        L.dref_get_seed_anchors.restype = C.c_int64
        n_reads = self._last_seed_count
        acap, pcap = 1 << 16, 1 << 22
        while True:
            anchors = np.zeros(acap, abi.SEED_ANCHOR)
            pool = np.zeros(pcap, np.uint64)
            begin = np.zeros(2 * n_reads + 1, np.uint32)
            used = C.c_uint64(0)
            n = L.dref_get_seed_anchors(abi.ptr(anchors), C.c_uint64(acap), abi.ptr(begin), abi.ptr(pool), C.c_uint64(pcap), C.byref(used))
            if n == abi.ERR_CAPACITY:
                acap, pcap = acap * 4, pcap * 4
                continue
            if n < 0:
                raise RuntimeError("dref_get_seed_anchors rc=%d" % n)
            return begin, anchors[:n].copy(), pool[:used.value].copy()

    def read_addr(self, k):
        self.lib.dref_read_addr.restype = C.c_uint64
        return int(self.lib.dref_read_addr(int(k)))

    def chroms(self):
        out = np.zeros(max(self.lib.dref_num_chr(), 1), abi.CHROM)
        n = self.lib.dref_get_chroms(abi.ptr(out), len(out))
        return out[:n]

    def seed_params(self):
        p = abi.SeedParams()
        self.lib.dref_get_seed_params(C.byref(p))
        return p

    def filter_last(self, gpu=False):
        """filter_body (the reference's, or the GPU host adapter's with gpu=True) on the last seed() output."""
        n = self.lib.dref_filter_last_gpu() if gpu else self.lib.dref_filter_last()
        if n < 0:
            raise RuntimeError("filter_last rc=%d" % n)
        return self._fetch_anchors(n)

    def extend(self, anchors, hit_pool, ops_cap=None):
        n = len(anchors)
        res = np.zeros(n, abi.ALN_RES)
        if ops_cap is None:
            ops_cap = int(anchors["read_len"].astype(np.int64).sum()) * 3 + 65536
        ops = np.zeros(ops_cap, np.uint8)
        hp = hit_pool if len(hit_pool) else np.zeros(1, np.uint64)
        rc = self.lib.dref_extend(abi.ptr(anchors), n, abi.ptr(hp), abi.ptr(res), abi.ptr(ops), C.c_uint64(ops_cap))
        if rc:
            raise RuntimeError("dref_extend rc=%d" % rc)
        return res, ops

    def extend_mt(self, anchors, hit_pool, threads):
        cells = C.c_uint64(0)
        alns = C.c_uint64(0)
        hp = hit_pool if len(hit_pool) else np.zeros(1, np.uint64)
        secs = self.lib.dref_extend_mt(abi.ptr(anchors), len(anchors), abi.ptr(hp), int(threads),
                                       C.byref(cells), C.byref(alns))
        return secs, cells.value, alns.value


def port(scoring):
    return Port(scoring)


def reference(flavour="patched"):
    """A driver handle in a DEFINED state.  The configuration of the compiled reference is process-global (`cfg`, like the
    reference's own main.cpp), so a test that left tile_size = 1024 / do_overlap = 1 behind would silently change how the
    next test seeds its reads: every new handle starts from the stock params.cfg values."""
    r = Reference(flavour)
    r.set_dsoft_defaults()
    r.set_extend(384, 64, 2, 0)
    return r


def have_reference():
    return os.path.exists(os.path.join(_HERE, "_ref", "libdarwin_ref_patched.so"))
