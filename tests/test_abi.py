"""CPU: the C-ABI shared library loads, exports every symbol include/darwin_gpu.h declares, the ctypes/numpy
mirrors match the header's struct sizes, and -- with no GPU -- the product fails loudly instead of falling back."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import darwin_b200
from darwin_b200 import abi
from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "darwin_gpu.h")


def _declared_functions():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"\b(darwin_gpu_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = darwin_b200.load_library()
    names = _declared_functions()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(darwin_b200.gact.EXPORTS)


def test_struct_sizes_match_header(tmp_path):
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include "darwin_gpu.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n",'
                    'sizeof(DarwinScoring),sizeof(DarwinTileReq),sizeof(DarwinTileRes),sizeof(DarwinAnchor),'
                    'sizeof(DarwinAlnRes),sizeof(DarwinExtendParams),sizeof(DarwinGpuStats));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert got == [C.sizeof(abi.Scoring), abi.TILE_REQ.itemsize, abi.TILE_RES.itemsize, abi.ANCHOR.itemsize,
                   abi.ALN_RES.itemsize, C.sizeof(abi.ExtendParams), C.sizeof(abi.GpuStats)]


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(darwin_b200.DarwinGpuError) as e:
        darwin_b200.Processor(1 << 20)
    assert e.value.code == abi.ERR_NO_DEVICE


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "darwin_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) and f != "smoke.py":
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "oracle/" not in txt.replace("oracle/gact_oracle.c", "").replace("oracle/dsoft_oracle.c", ""), f   # comments may NAME the CPU twins


def test_dropin_library_resolves_all_symbols():
    """oracle/_ref/libdarwin_ref_gpu.so (reference TUs + our C++ host adapter, linked against libdarwin_gact.so) must load with
    every symbol bound -- an unresolved symbol there only shows up as a failed GPU test otherwise."""
    import ctypes
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libdarwin_ref_gpu.so")
    if not os.path.exists(path):
        import pytest
        pytest.skip("oracle/_ref/libdarwin_ref_gpu.so not built (needs /root/reference at build time)")
    lib = ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL | os.RTLD_NOW)
    for sym in ("dref_gpu_init", "dref_pipeline", "dref_pipeline_mt", "dref_pipeline_cpu_mt"):
        assert hasattr(lib, sym), sym
