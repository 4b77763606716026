"""The C-ABI shared library loads on a CPU-only box and exports exactly what include/v3s.h declares
(no compute calls here: those need a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "v3s.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(v3s_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    from v3s import _lib
    names = declared_symbols()
    assert len(names) >= 30
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    unbound = [n for n in names if n not in _lib.SIGNATURES]
    assert not unbound, unbound
    undeclared = [n for n in _lib.SIGNATURES if n not in names]
    assert not undeclared, undeclared
    assert lib.v3s_version() == 100


def test_error_plumbing_without_gpu():
    from v3s import _lib
    # argument validation happens before any CUDA call: a bad shape returns non-zero with a message
    rc = _lib.lib.v3s_block_matvec(None, 3, 1, 10, None, 99, None, None, None)
    assert rc != 0 and "block size" in _lib.last_error()
    rc = _lib.lib.v3s_sparse_block_matvec(None, 0, 1, 10, None, 8, None, None)
    assert rc != 0 and "operator descriptor" in _lib.last_error()
    assert _lib.lib.v3s_xf_floats(130) == 2048
    assert _lib.lib.v3s_synth_workspace_bytes(64) == 64 * 64 * 16
