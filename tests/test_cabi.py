"""The C-ABI library loads and exports every symbol include/genlm_trie_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

from genlm_backend_b200 import _lib
from genlm_backend_b200.build import LIB_PATH

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "genlm_trie_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gt_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    names = declared_symbols()
    assert len(names) >= 18
    lib = ctypes.CDLL(LIB_PATH)
    for name in names:
        assert hasattr(lib, name), name
        assert name in _lib.SIGNATURES, f"{name} is declared in the header but not bound in _lib.py"
    assert set(_lib.SIGNATURES) == set(names)


def test_status_and_errors_without_a_gpu():
    assert _lib.lib.gt_version() >= 1
    assert _lib.lib.gt_num_nodes(None) == -1
    rc = _lib.lib.gt_build(None, None, 0, None)
    assert rc == 1 and b"bad argument" in _lib.lib.gt_last_error()
    info = _lib.PlanInfo()
    assert _lib.lib.gt_get_plan_info(None, ctypes.byref(info)) != 0
    assert _lib.lib.gt_workspace_bytes(None, 10) == 0
    # compute entry points validate arguments before touching CUDA
    assert _lib.lib.gt_weight_reduce(None, None, 0, 1, 1, None, None, 0, 1, 1, 0, None, 0, None) == 1
    assert _lib.lib.gt_lse_sample(None, 0, 1, 10, 10, None, 0, 0, 1.0, 0, 0, None, None, None) == 1


def test_library_is_built_for_sm_100a():
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        return
    out = subprocess.run([cuobjdump, "--list-elf", LIB_PATH], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
    assert "sm_100a" in out, out
