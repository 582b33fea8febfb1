"""CPU: libgnc.so builds for sm_100a, loads, and exports exactly the symbols that
include/gnc.h declares (no compute calls here)."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "gnc.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(gnc_[a-z0-9_]+)\s*\(", text))


def test_header_matches_binding(libgnc):
    from graphnet_classifier_b200 import _lib
    declared = _declared()
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(libgnc, name), name


def test_exports_and_arch(libgnc):
    from graphnet_classifier_b200 import _lib
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.lib_path()], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (gnc_[a-z0-9_]+)", out))
    assert _declared() <= exported
    elf = subprocess.run(["cuobjdump", "-lelf", _lib.lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in elf, elf


def test_host_only_entry_points(libgnc):
    assert libgnc.gnc_version() >= 100
    assert libgnc.gnc_grid_num_edges(128, 128, 0) == 2 * 128 * 127
    assert libgnc.gnc_grid_num_edges(128, 128, 1) == 2 * 128 * 127 + 2 * 127 * 127
    assert libgnc.gnc_grid_num_edges(1, 1, 1) == 0
    assert libgnc.gnc_csr_workspace(1000) >= 1000
    assert libgnc.gnc_linear_wgrad_workspace(100000, 128, 384) >= 128 * 384
    libgnc.gnc_reset_launch_count()
    assert libgnc.gnc_launch_count() == 0
    # argument validation happens before any CUDA call
    assert libgnc.gnc_agg_csr_sum_f32(None, None, None, 4, 10, 8, None, 8, 0, None) == 1
    assert b"ld" in libgnc.gnc_last_error() or b"bad" in libgnc.gnc_last_error()
