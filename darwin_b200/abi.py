"""ctypes mirror of include/darwin_gpu.h (the C-ABI of libdarwin_gact.so).

Field order and widths follow the header exactly; the header in turn follows
the reference's wire types (software/Darwin.bond:42-141, software/graph.h:83-121).
numpy structured dtypes with the same layout are provided for bulk buffers.
"""
import ctypes as C

import numpy as np

# align_fields bits (software/graph.h:22-26)
REVERSE_REF = 1 << 4
COMPLEMENT_REF = 1 << 3
REVERSE_QUERY = 1 << 2
COMPLEMENT_QUERY = 1 << 1
START_END = 1

OP_I, OP_D, OP_M = 1, 2, 3          # software/Processor.h:14 (states % 4)
MAX_TILE = 1984                      # software/extender.cpp:70-75

ALN_EMITTED = 1 << 0
ALN_OPS_OVERFLOW = 1 << 1
ALN_EXACT_RERUN = 1 << 2
ALN_LONG_INS_PATH = 1 << 3
TILE_LONG_INS_PATH = 0x10          # DarwinTileRes.status bit (include/darwin_gpu.h)

OK, ERR_NO_DEVICE, ERR_INVALID, ERR_CUDA, ERR_CAPACITY, ERR_NOT_READY, ERR_NOMEM = 0, -1, -2, -3, -4, -5, -6


class Scoring(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "sub_AA", "sub_AC", "sub_AG", "sub_AT", "sub_CC", "sub_CG", "sub_CT", "sub_GG", "sub_GT",
        "sub_TT", "sub_N", "gap_open", "gap_extend", "long_gap_open", "long_gap_extend")]

    @classmethod
    def from_values(cls, match=2, mismatch=-6, sub_n=-1, go=-4, ge=-2, lgo=-25, lge=-1, matrix=None):
        """params.cfg defaults (software/params.cfg:2-16) unless overridden."""
        m = matrix or dict(AA=match, AC=mismatch, AG=mismatch, AT=mismatch, CC=match, CG=mismatch,
                           CT=mismatch, GG=match, GT=mismatch, TT=match)
        return cls(m["AA"], m["AC"], m["AG"], m["AT"], m["CC"], m["CG"], m["CT"], m["GG"], m["GT"], m["TT"],
                   sub_n, go, ge, lgo, lge)

    def as_tuple(self):
        return tuple(getattr(self, n) for n, _ in self._fields_)


class ExtendParams(C.Structure):
    _fields_ = [("tile_size", C.c_int32), ("tile_overlap", C.c_int32), ("do_overlap", C.c_int32),
                ("reserved", C.c_int32)]


class FilterParams(C.Structure):
    """params.cfg [GACT_first_tile] (software/params.cfg:29-34) as darwin_gpu_filter takes it."""
    _fields_ = [("first_tile_size", C.c_int32), ("first_tile_score_threshold", C.c_int32), ("min_overlap", C.c_int32),
                ("reserved", C.c_int32)]


class SeedParams(C.Structure):
    """params.cfg [DSOFT_params] (software/params.cfg:18-27)."""
    _fields_ = [(n, C.c_int32) for n in ("seed_size", "minimizer_window", "bin_size", "threshold", "num_seeds",
                                         "seed_occurence_multiple", "max_stride", "do_overlap")]

    @classmethod
    def stock(cls, do_overlap=0):
        return cls(14, 3, 64, 26, 1000, 40, 4, do_overlap)


class AlignParams(C.Structure):
    """What darwin_gpu_align_reads needs beyond the seeding parameters: params.cfg [GACT_first_tile], [GACT_extend]."""
    _fields_ = [("filter", FilterParams), ("extend", ExtendParams), ("slope_threshold", C.c_float), ("reserved", C.c_int32)]

    @classmethod
    def stock(cls, tile_size=384, tile_overlap=64, do_overlap=0):
        return cls(FilterParams(128, 60, 1000, 0), ExtendParams(tile_size, tile_overlap, do_overlap, 0), 0.05, 0)


class GpuStats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("tiles_fast", C.c_uint64), ("tiles_exact", C.c_uint64),
                ("tiles_rerun", C.c_uint64), ("cells", C.c_uint64), ("cells_exact", C.c_uint64), ("tiles_xfast", C.c_uint64), ("last_kernel_ms", C.c_float), ("reserved", C.c_float),
                ("tiles_filter", C.c_uint64), ("last_seed_ms", C.c_float), ("last_filter_ms", C.c_float), ("last_extend_ms", C.c_float),
                ("reserved2", C.c_float), ("tiles_scoreonly", C.c_uint64)]


TILE_REQ = np.dtype([
    ("ref_bases_start_addr", "<u8"), ("query_bases_start_addr", "<u8"), ("score_threshold", "<u4"),
    ("index", "<u2"), ("ref_size", "<u2"), ("query_size", "<u2"), ("max_tb_steps", "<u2"),
    ("align_fields", "u1"), ("reserved", "u1", (3,))], align=False)
assert TILE_REQ.itemsize == 32

TILE_RES = np.dtype([
    ("score", "<i4"), ("ref_offset", "<u2"), ("query_offset", "<u2"), ("ref_max_pos", "<u2"),
    ("query_max_pos", "<u2"), ("total_TB_pointers", "<u2"), ("index", "u1"), ("status", "u1")], align=False)
assert TILE_RES.itemsize == 16

FILTER_CAND = np.dtype([
    ("read_addr", "<u8"), ("hit", "<u4"), ("offset", "<u4"), ("chr_start", "<u4"), ("chr_len", "<u4"),
    ("read_len", "<u4"), ("strand", "u1"), ("reserved", "u1", (3,))], align=False)
assert FILTER_CAND.itemsize == 32

FILTER_RES = np.dtype([("score", "<i4"), ("reference_pos", "<u4"), ("query_pos", "<u4"), ("flags", "<u4")], align=False)
assert FILTER_RES.itemsize == 16
FILTER_SCORE_OK, FILTER_OVERLAP_OK = 1, 2

CHROM = np.dtype([("start", "<u4"), ("len_unpadded", "<u4")], align=False)
SEED_READ = np.dtype([("read_addr", "<u8"), ("read_len", "<u4"), ("reserved", "<u4")], align=False)
SEED_ANCHOR = np.dtype([("hit_offset", "<u8"), ("left_off", "<u8"), ("right_off", "<u8"), ("left_n", "<u4"), ("right_n", "<u4")],
                       align=False)
assert CHROM.itemsize == 8 and SEED_READ.itemsize == 16 and SEED_ANCHOR.itemsize == 32

ANCHOR = np.dtype([
    ("read_addr", "<u8"), ("reference_pos", "<u4"), ("query_pos", "<u4"), ("chr_start", "<u4"),
    ("ref_len", "<u4"), ("read_len", "<u4"), ("read_num", "<i4"), ("chr_id", "<i4"), ("score", "<i4"),
    ("left_hits_off", "<u4"), ("left_hits_n", "<u4"), ("right_hits_off", "<u4"), ("right_hits_n", "<u4"),
    ("strand", "u1"), ("reserved", "u1", (7,))], align=False)
assert ANCHOR.itemsize == 64

ALN_RES = np.dtype([
    ("ops_offset", "<u8"), ("cells", "<u8"), ("n_ops", "<u4"), ("reference_start_offset", "<u4"),
    ("reference_end_offset", "<u4"), ("query_start_offset", "<u4"), ("query_end_offset", "<u4"),
    ("n_left_ops", "<u4"), ("n_tiles", "<u4"), ("n_large_tiles", "<u4"), ("score", "<i4"), ("flags", "<u4")],
    align=False)
assert ALN_RES.itemsize == 56


def ptr(a, ctype=C.c_void_p):
    """numpy array -> C pointer (None passes NULL)."""
    if a is None:
        return None
    return a.ctypes.data_as(ctype)
