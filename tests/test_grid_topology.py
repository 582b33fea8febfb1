"""CPU: the closed-form grid topology (csrc/grid_topology.h, shared by host and device)
against the numpy oracle - edges, and the stable CSR by destination / by source."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import graph_build as ogb

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def harness():
    out = os.path.join(tempfile.gettempdir(), f"gnc_topo_harness_{os.getuid()}.so")
    src = os.path.join(HERE, "cpu_harness", "grid_topology_harness.cpp")
    subprocess.check_call(["g++", "-O1", "-shared", "-fPIC", "-o", out, src])
    lib = ctypes.CDLL(out)
    lib.harness_num_edges.restype = ctypes.c_longlong
    return lib


def _check(lib, H, W, diag):
    ref = ogb.grid_edges(H, W, bool(diag))
    E = lib.harness_num_edges(H, W, diag)
    assert E == ref.shape[1]
    src, dst = np.zeros(E, np.int64), np.zeros(E, np.int64)
    lib.harness_edges(H, W, diag, src.ctypes.data_as(ctypes.c_void_p), dst.ctypes.data_as(ctypes.c_void_p))
    assert np.array_equal(src, ref[0]) and np.array_equal(dst, ref[1])
    for which in (0, 1):
        rp, eid = np.full(H * W + 1, -1, np.int32), np.full(E, -1, np.int32)
        lib.harness_csr(H, W, diag, which, rp.ctypes.data_as(ctypes.c_void_p), eid.ctypes.data_as(ctypes.c_void_p))
        orp, oeid = ogb.csr_by_key(ref[which], H * W)
        assert np.array_equal(rp, orp) and np.array_equal(eid, oeid)


def test_small_grids_exhaustive(harness):
    for H in range(1, 7):
        for W in range(1, 7):
            for diag in (0, 1):
                _check(harness, H, W, diag)


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 70), st.integers(1, 70), st.integers(0, 1))
def test_random_grids(harness, H, W, diag):
    _check(harness, H, W, diag)


def test_csr_round_trip_property():
    # edge_index -> CSR(dst) -> edge list recovers the multiset of edges, rows ascend
    ei = ogb.grid_edges(9, 13, True)
    rp, eid = ogb.csr_by_key(ei[1], 9 * 13)
    assert rp[-1] == ei.shape[1] and np.all(np.diff(rp) >= 0)
    for v in range(9 * 13):
        seg = eid[rp[v]:rp[v + 1]]
        assert np.all(ei[1, seg] == v) and np.all(np.diff(seg) > 0)
    assert np.array_equal(np.sort(eid), np.arange(ei.shape[1]))
