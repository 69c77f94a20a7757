"""CPU: the cross-read batcher of the host adapter (darwin_b200/host/darwin_gpu_combiner.h) -- requests of many host
threads merged into single device calls and scattered back -- with the oracle standing in for the device
(tests/cpp/test_combiner.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from darwin_b200 import abi, synth
from conftest import ROOT


@pytest.fixture(scope="module")
def selftest_lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("combiner") / "libcombiner_test.so")
    src = os.path.join(ROOT, "tests", "cpp", "test_combiner.cpp")
    cmd = ["g++", "-std=c++11", "-O1", "-fPIC", "-shared", "-pthread", "-I", os.path.join(ROOT, "include"), src, "-o", out,
           "-L", os.path.join(ROOT, "oracle"), "-lgact_oracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle")]
    subprocess.run(cmd, check=True)
    return C.CDLL(out)


@pytest.mark.parametrize("threads", [1, 6])
def test_combiner_merges_and_scatters(selftest_lib, threads):
    arena, anchors, hits = synth.anchor_batch(3, 12, 1500, 60000)
    t_arena, req = synth.tile_batch(5, 24, 96)
    base = len(arena)
    req = req.copy()
    req["ref_bases_start_addr"] += base
    req["query_bases_start_addr"] += base
    dram = np.concatenate([arena, t_arena])
    # first-tile candidates at the anchors' own loci, both strands' request shapes
    cands = np.zeros(len(anchors), abi.FILTER_CAND)
    cands["read_addr"], cands["read_len"] = anchors["read_addr"], anchors["read_len"]
    cands["chr_start"], cands["chr_len"] = anchors["chr_start"], anchors["ref_len"]
    cands["hit"] = np.maximum(anchors["reference_pos"].astype(np.int64) - 40, anchors["chr_start"])
    cands["offset"] = np.maximum(anchors["query_pos"].astype(np.int64) - 40, 0)
    cands["strand"] = anchors["strand"]
    sc = abi.Scoring.from_values()
    stats = np.zeros(12, np.uint64)
    rc = selftest_lib.combiner_selftest(C.byref(sc), abi.ptr(dram), C.c_uint64(len(dram)), abi.ptr(req), len(req),
                                        abi.ptr(cands), len(cands), abi.ptr(anchors), len(anchors), abi.ptr(hits),
                                        C.c_uint64(len(hits)), 320, 128, threads, abi.ptr(stats))
    assert rc == 0, rc
    calls, requests, merged = stats[0:3], stats[3:6], stats[6:9]
    assert list(requests) == [threads * 4, threads * 3, threads * 3]        # tiles: 3 rounds + the upload-only request
    assert stats[10] == threads * 3 and (threads == 1 or (stats[9] < stats[10] and stats[11] >= 2))      # seeding requests
    if threads == 1:
        assert list(calls) == list(requests)
    else:
        assert (calls < requests).all() and (merged >= 2).all()             # requests really rode together


@pytest.mark.parametrize("threads", [1, 7])
def test_combiner_align_requests_merge_and_scatter(selftest_lib, threads):
    """ALIGN requests (gpu_align_body / gpu_sam_body -> darwin_gpu_align_reads): every caller gets exactly its own reads'
    locations -- forward part then reverse part, read numbers rebased to its own batch, op strings dense in that order --
    whatever it was merged with, including a caller that brings no reads."""
    stats = np.zeros(4, np.uint64)
    rounds = 4
    rc = selftest_lib.combiner_align_selftest(threads, 9, rounds, abi.ptr(stats))
    assert rc == 0, rc
    calls, requests, merged = int(stats[0]), int(stats[1]), int(stats[2])
    assert requests == threads * rounds
    if threads == 1:
        assert calls == requests and merged == 1
    else:
        assert calls < requests and merged >= 2
