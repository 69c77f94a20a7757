import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _build_native():
    """Build the oracle libraries (gcc) once per session; the CUDA library is built by __graft_entry__.build()."""
    import oracle
    import __graft_entry__ as entry
    try:
        entry.build()                      # nvcc (no-op when the library is up to date) + oracle libraries
    except Exception:
        oracle.build("port")               # no nvcc here: the prebuilt libdarwin_gact.so (if any) is used as is
        if os.path.isdir("/root/reference/software"):
            oracle.build("ref")


@pytest.fixture(scope="session")
def golden_tiles():
    return np.load(os.path.join(GOLDEN, "tiles_v1.npz"))


@pytest.fixture(scope="session")
def golden_extend():
    return np.load(os.path.join(GOLDEN, "extend_v1.npz"))


@pytest.fixture(scope="session")
def gpu():
    """A GPU-backed Processor factory; fails loudly (no skip, no fallback) when the CUDA library is missing."""
    import darwin_b200

    def make(arena_bytes, scoring):
        p = darwin_b200.Processor(arena_bytes)
        p.InitializeScoringParameters(scoring)
        return p
    return make


TILE_FIELDS = ("score", "ref_offset", "query_offset", "ref_max_pos", "query_max_pos", "total_TB_pointers", "index")


def tiles_equal(res_a, tb_a, res_b, tb_b):
    """Index list of tiles whose result struct or used TB words differ.  Of `status` only the error nibble is compared:
    bit 4 (DARWIN_TILE_LONG_INS_PATH) is information the reference does not return."""
    bad = []
    for k in range(len(res_a)):
        if any(res_a[k][f] != res_b[k][f] for f in TILE_FIELDS) or (int(res_a[k]["status"]) ^ int(res_b[k]["status"])) & 0x0F:
            bad.append(k)
            continue
        nw = (int(res_a[k]["total_TB_pointers"]) + 31) // 32
        if nw and not np.array_equal(tb_a[k, :nw], tb_b[k, :nw]):
            bad.append(k)
    return bad


# fields the reference driver reports (n_left_ops is known only to our implementations)
ALN_FIELDS = ("n_ops", "reference_start_offset", "reference_end_offset", "query_start_offset", "query_end_offset",
              "n_tiles", "n_large_tiles", "score", "cells")
ALN_FIELDS_OURS = ALN_FIELDS + ("n_left_ops",)


def alignments_equal(res_a, ops_a, res_b, ops_b, fields=ALN_FIELDS):
    bad = []
    for k in range(len(res_a)):
        a, b = res_a[k], res_b[k]
        ea, eb = int(a["flags"]) & 1, int(b["flags"]) & 1
        ok = ea == eb and all(a[f] == b[f] for f in fields)
        if ok and ea:
            ok = np.array_equal(ops_a[int(a["ops_offset"]):int(a["ops_offset"]) + int(a["n_ops"])],
                                ops_b[int(b["ops_offset"]):int(b["ops_offset"]) + int(b["n_ops"])])
        if not ok:
            bad.append(k)
    return bad
