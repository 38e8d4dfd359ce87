"""The C-ABI library loads and exports every symbol include/gnn_b200.h declares
(no compute calls: runs without a GPU)."""
import ctypes
import os
import re

from conftest import ROOT
from graphneuralnetwork_b200 import _lib

HEADER = os.path.join(ROOT, "include", "gnn_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gnn_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_functions():
    names = declared_functions()
    assert len(names) >= 30
    assert "gnn_spmm_csr_f32" in names and "gnn_gat_fused_fwd_f32" in names and "gnn_gather_reduce_f32" in names


def test_binding_table_matches_header():
    assert sorted(_lib.SIGNATURES) == declared_functions()


def test_library_exports_every_symbol(lib):
    raw = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in declared_functions():
        assert hasattr(raw, name), f"{name} missing from libgnn_b200.so"


def test_version_and_error_strings(lib):
    assert lib.gnn_version() >= 100
    assert lib.gnn_status_string(0) == b"ok"
    assert lib.gnn_status_string(3) == b"unsupported shape"
    assert lib.gnn_launch_count() >= 0


def test_tuning_knobs(lib):
    _lib.set_tuning("sage.smem_kb", 64)
    assert _lib.get_tuning("sage.smem_kb") == 64
    _lib.set_tuning("sage.smem_kb", 48)
    try:
        _lib.set_tuning("no.such.key", 1)
    except _lib.GnnError as e:
        assert "unknown tuning key" in str(e)
    else:
        raise AssertionError("unknown key accepted")


def test_bad_arguments_return_status_not_abort(lib):
    # argument validation happens before any CUDA call, so this is safe without a GPU
    rc = lib.gnn_spmm_csr_f32(None, None, None, None, None, 4, 4, 8, 8, 8, None)
    assert rc == 1 and b"null pointer" in lib.gnn_last_error_string()
    rc = lib.gnn_gather_reduce_f32(None, 8, 4, None, 64, 4, 2, 8, 7, None, 8, None, None)
    assert rc == 1 and b"unknown reduce" in lib.gnn_last_error_string()
    rc = lib.gnn_gat_fused_fwd_f32(None, None, None, 8, None, None, 4, 0, 64, 8, 0.2, 0, 0, None, None, None, 8,
                                   None, None, None, 0, 0, 0, None)
    assert rc == 3  # more than 32 heads per call
