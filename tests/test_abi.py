"""The C-ABI library loads on a CPU-only box and exports every symbol the
header declares; compute entry points refuse to run without a device."""
import subprocess
import sys

import pytest

from code_saturne_b200 import api


def test_library_loads_and_exports_all_header_symbols():
    L = api.load_library()
    syms = api.header_symbols()
    assert len(syms) >= 60
    missing = [s for s in syms if not hasattr(L, s)]
    assert missing == []


def test_header_cites_reference_interfaces():
    txt = open(api.HEADER_PATH).read()
    for ref in ("cs_matrix.cpp", "cs_sles_it.cpp", "cs_multigrid.cpp", "cs_halo.cpp",
                "cs_sles.h"):
        assert ref in txt


def test_no_cpu_fallback_without_device():
    """csb_init must fail loudly when there is no CUDA device."""
    code = ("import os; os.environ['CUDA_VISIBLE_DEVICES']='';"
            "from code_saturne_b200 import api\n"
            "try:\n    api.init(0)\n    print('INIT_OK')\n"
            "except api.CsbError as e:\n    print('INIT_FAILED', e)\n")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True,
                         cwd=api._ROOT).stdout
    assert "INIT_FAILED" in out and "no CPU fallback" in out


def test_objects_need_init():
    L = api.load_library()
    import ctypes as C
    import os
    if os.environ.get("CUDA_VISIBLE_DEVICES", None) != "":
        # in-process check only meaningful before init; run in a subprocess
        code = ("import os; os.environ['CUDA_VISIBLE_DEVICES']='';"
                "from code_saturne_b200 import api; L=api.load_library();"
                "print('H', L.csb_multigrid_create(0), L.csb_sles_it_create(0,0,10,0))")
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True,
                             cwd=api._ROOT).stdout
        assert "H None None" in out
