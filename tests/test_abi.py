"""CPU-only: the C-ABI library builds, loads and exports every symbol include/pangnn_b200.h declares
(no compute calls here — those are the -m gpu tests)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "pangnn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pangnn_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_something():
    syms = declared_symbols()
    assert "pangnn_gcn_aggregate" in syms and "pangnn_edge_score_bwd" in syms and len(syms) >= 20


def test_library_builds_and_exports_every_declared_symbol():
    from pangnn_b200 import build
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in the header but not exported"


def test_ctypes_table_matches_header():
    from pangnn_b200 import _abi
    assert sorted(_abi.SIGNATURES) == declared_symbols()
    lib = _abi.load()
    assert lib.pangnn_abi_version() == _abi.ABI_VERSION


def test_header_arity_matches_ctypes_table():
    from pangnn_b200 import _abi
    src = open(os.path.join(ROOT, "include", "pangnn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    for name, params in re.findall(r"\b(pangnn_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src):
        params = params.strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(_abi.SIGNATURES[name][1]), name


def test_ops_refuse_cpu_tensors():
    import torch
    from pangnn_b200 import _abi, ops
    with pytest.raises(_abi.PangnnError):
        ops.csr_build(torch.zeros(2, 3, dtype=torch.long), 4)
    with pytest.raises(_abi.PangnnError):
        ops.gcn_aggregate(None, None, None, torch.zeros(4, 64), 4)


def test_band_edge_count_is_host_arithmetic():
    """pangnn_neighbour_band_edges launches nothing: closed form vs the oracle's band (src/dataset.py:351-366)."""
    from oracle import preprocess as op
    from pangnn_b200 import _abi
    lib = _abi.load()
    for N in (1, 2, 3, 7, 50):
        for n in (0, 1, 3, 4, 9):
            assert lib.pangnn_neighbour_band_edges(N, n) == op.neighbour_band(N, n).shape[1], (N, n)
    assert lib.pangnn_neighbour_band_edges(0, 3) == 0
    assert lib.pangnn_neighbour_band_edges(10 ** 6, 3) == 7 * 10 ** 6 - 12


def test_oracle_component_labels_known_answer():
    from oracle import postprocess as opp
    import numpy as np
    lab = opp.component_labels(np.array([0, 1, 5, 7, 8]), np.array([1, 2, 6, 7, 9]), np.array([1, 1, 1, 1, 0]), 10)
    assert lab.tolist() == [0, 0, 0, 3, 4, 5, 5, 7, 8, 9]
    assert opp.groups(lab) == [[0, 1, 2], [5, 6]]


def test_fnv1a64_known_answers():
    """The host side of the parser's id matching: FNV-1a 64 (known vectors of the published algorithm)."""
    from pangnn_b200 import ops
    h = ops.fnv1a64(["", "a", "foobar", "FFOKMCCD_00001"])
    assert int(h[0]) == 0xcbf29ce484222325 and int(h[1]) == 0xaf63dc4c8601ec8c and int(h[2]) == 0x85944171f73967e8
    ref = 0xcbf29ce484222325
    for ch in b"FFOKMCCD_00001":
        ref = ((ref ^ ch) * 0x100000001b3) % (1 << 64)
    assert int(h[3]) == ref
