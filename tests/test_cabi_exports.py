"""CPU: the C-ABI library loads and exports every symbol include/gfc.h declares
(no compute calls without a GPU)."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "gfc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gfc_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    import gnnfc
    lib = ctypes.CDLL(gnnfc._cabi.LIB_PATH)
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), "libgfc.so does not export %s" % n


def test_binding_covers_header():
    import gnnfc
    assert sorted(gnnfc._cabi.SIGNATURES) == _declared()


def test_version_and_errors_without_gpu():
    import gnnfc
    C = gnnfc._cabi
    assert C.version() == 100
    # argument validation happens before any CUDA call
    rc = C.lib.gfc_gso_build(None, 4, 3, 2.0, 0, None, None, None)
    assert rc == C.GFC_ERR_BAD_ARG and b"NULL" in C.lib.gfc_last_error()
    rc = C.lib.gfc_gso_build(None, -1, 3, 2.0, 0, None, None, None)
    assert rc == C.GFC_ERR_BAD_ARG
    rc = C.lib.gfc_filter_fwd(None, None, None, None, None, 2, 3, 8, 8, 0, 1, 0, 0.0, 0, None, 0, None)
    assert rc == C.GFC_ERR_BAD_ARG and b"bad shape" in C.lib.gfc_last_error()
    rc = C.lib.gfc_filter_fwd(None, None, None, None, None, 2, 3, 8, 8, 2, 1, 7, 0.0, 0, None, 0, None)
    assert rc == C.GFC_ERR_BAD_ARG and b"activation" in C.lib.gfc_last_error()
    with pytest.raises(C.GfcError):
        C.check(rc, "gfc_filter_fwd")
    # empty batch is a no-op, not an error
    assert C.lib.gfc_gso_build(None, 0, 3, 2.0, 0, None, None, None) == C.GFC_OK


def test_mask_handover_host_side_without_gpu():
    """gfc_use_mask / gfc_filter_mask_bytes / the option switches are host logic (include/gfc.h): sizes, alignment
    check and option keys can be checked without a device."""
    import gnnfc
    C = gnnfc._cabi
    assert C.lib.gfc_filter_mask_bytes(65536, 64, 128, 128, 4) == (65536 // 2) * 2048     # cfg3: 2 graphs per 128-row tile
    assert C.lib.gfc_filter_mask_bytes(16384, 12, 128, 128, 3) == -(-16384 // 10) * 2048  # cfg4 shape: 10 graphs per tile
    assert C.lib.gfc_filter_mask_bytes(1, 3, 64, 64, 3) == 2048
    assert C.lib.gfc_filter_mask_bytes(4096, 8, 32, 32, 3) == 0                            # narrow features: no such kernel
    assert C.lib.gfc_filter_mask_bytes(0, 8, 128, 128, 3) == 0
    assert C.lib.gfc_use_mask(ctypes.c_void_p(0x1004), 4096) == C.GFC_ERR_BAD_ARG
    assert b"aligned" in C.lib.gfc_last_error()
    assert C.lib.gfc_use_mask(None, 0) == C.GFC_OK                                         # NULL cancels
    assert C.lib.gfc_mask_filled() == 0
    for key in (C.OPT_WIDE_MASK_HANDOVER, C.OPT_WIDE_FWD_MASK, C.OPT_CSR_STAGE_IDX):
        assert C.lib.gfc_set_option(key, 0) == C.GFC_OK
        assert C.lib.gfc_set_option(key, 1) == C.GFC_OK                                    # defaults restored
    assert C.lib.gfc_set_option(12345, 1) == C.GFC_ERR_BAD_ARG


def test_option_environment_hook():
    """GFC_SET_OPTIONS="key=value,..." (an A/B aid of the tools) is applied when the binding loads"""
    code = ("import gnnfc; C = gnnfc._cabi; "
            "assert C.lib.gfc_set_option(C.OPT_CSR_STAGE_IDX, 1) == 0; print('ok')")
    env = dict(os.environ, GFC_SET_OPTIONS="10=0,9=1", PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    out = subprocess.run(["python", "-c", code], capture_output=True, text=True, env=env, cwd=ROOT)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-400:]
    bad = subprocess.run(["python", "-c", "import gnnfc"], capture_output=True, text=True,
                         env=dict(env, GFC_SET_OPTIONS="nonsense"), cwd=ROOT)
    assert bad.returncode != 0                                                             # malformed: fails loudly


def test_library_is_sm100a_only_and_uses_tensor_cores():
    import gnnfc
    out = subprocess.run(["cuobjdump", "-lelf", gnnfc._cabi.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_tile_planner_shapes():
    """host logic: which BASELINE configs take the fused path, and their tiling"""
    import gnnfc
    C = gnnfc._cabi
    cfg2 = C.tile_plan(4096, 8, 32, 32, 3, backward=True, from_positions=True)
    assert cfg2["ok"] and cfg2["graphs_per_tile"] == 16 and cfg2["rows_padded"] == 128
    assert cfg2["taps_in_smem"] == 1 and cfg2["dh_in_registers"] == 1
    assert cfg2["smem_bytes"] <= 113 * 1024            # two CTAs per SM
    cfg1 = C.tile_plan(64, 3, 128, 128, 3)
    assert cfg1["ok"] and cfg1["graphs_per_tile"] * 3 == cfg1["rows"] <= cfg1["rows_padded"]
    cfg3 = C.tile_plan(65536, 64, 128, 128, 4, backward=True)
    assert cfg3["ok"] and cfg3["grid"] == 148 and cfg3["smem_bytes"] <= 232448
    assert C.lib.gfc_filter_path(256, 1024, 32, 32, 5, 1, 0) == 2   # too large for smem -> workspace path
    assert C.lib.gfc_filter_path(8, 6, 5, 7, 4, 2, 0) == 2          # E = 2 / odd features
    assert C.lib.gfc_filter_workspace_bytes(4096, 8, 32, 32, 3, 1, 0) == 0
    assert C.lib.gfc_filter_workspace_bytes(4096, 8, 32, 32, 3, 1, 1) > 0
