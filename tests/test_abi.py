"""The C-ABI library loads on a CPU-only host and exports every symbol include/tta_b200.h declares
(no compute calls here: there is no GPU)."""
import os
import re

from multimodal_tta_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "tta_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tta_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_loads(lib):
    assert os.path.exists(_lib.LIB_PATH)
    assert lib.tta_abi_version() == 1
    assert lib.tta_last_error() is not None


def test_every_header_symbol_is_exported(lib):
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/tta_b200.h but not exported"


def test_ctypes_table_matches_header(lib):
    assert set(_lib.EXPORTED_SYMBOLS) == set(_header_symbols())


def test_geometry_queries_without_gpu(lib):
    # pure host-side queries
    assert lib.tta_conv_tc_supported(0, 3, 1, 32, 32) == 1
    assert lib.tta_conv_tc_supported(0, 5, 1, 32, 32) == 0
    assert lib.tta_conv_tc_ntile(0, 3, 1, 32, 1) == 32 and lib.tta_conv_tc_ntile(0, 3, 1, 512, 1) == 128
    assert lib.tta_conv_tc_ntile(1, 3, 2, 128, 1) == 32 and lib.tta_conv_tc_ntile(1, 3, 2, 128, 0) == 64
    assert lib.tta_conv_tc_ntile(0, 3, 1, 3, 1) == 16
    assert lib.tta_conv_tc_ngroups(1, 3, 2) == 2 and lib.tta_conv_tc_gmax(1, 3, 2) == 18
    assert lib.tta_conv_small_supported(3, 1, 3, 3) == 1 and lib.tta_conv_small_supported(3, 2, 3, 3) == 0
    assert lib.tta_norm_workspace_floats(2, 4, 64 ** 3) > 0


def test_bad_arguments_are_reported_not_crashed(lib):
    rc = lib.tta_adam_step(0, 0, 0, 0, 10, 1e-3, 0.9, 0.999, 1e-8, 1.0, 0, 0)
    assert rc == 1 and b"null pointer" in lib.tta_last_error()
    rc = lib.tta_head_entropy(0, 0, 1, 9, 8, 1, 1.0, 1.0, 1, 0, 0, 0, 0, 0, 0, 0, 0)
    assert rc == 1


def test_plan_recorder_lifecycle_without_gpu(lib):
    """tta_plan (plan_create / tta_step): while a section is being recorded the launching entry points
    store themselves instead of touching the device -- which is why this runs on a CPU-only host."""
    import ctypes
    h = ctypes.c_void_p()
    assert lib.tta_plan_create(ctypes.byref(h)) == 0 and h.value
    assert lib.tta_plan_num_launches(h, 0) == 0 and lib.tta_plan_num_launches(h, 99) == -1
    assert lib.tta_plan_end() == 1 and b"no recording" in lib.tta_last_error()
    assert lib.tta_plan_begin(h, 3) == 0
    assert lib.tta_plan_begin(h, 0) == 1 and b"already active" in lib.tta_last_error()
    # recorded, not validated or launched: argument checks happen at replay, like the launch itself
    assert lib.tta_adam_step(0, 0, 0, 0, 10, 1e-3, 0.9, 0.999, 1e-8, 1.0, 0, 0) == 0
    assert lib.tta_norm_stats(0, 0, 1, 1, 8, 0, 1e-5, 0, 0, 0, 1, 0) == 0
    assert lib.tta_plan_end() == 0
    assert lib.tta_plan_num_launches(h, 3) == 2 and lib.tta_plan_num_launches(h, 0) == 0
    # replaying surfaces the recorded call's own argument error (nothing reaches the device)
    assert lib.tta_plan_run(h, 3, 0) == 1 and b"null pointer" in lib.tta_last_error()
    assert lib.tta_plan_begin(h, 3) == 0 and lib.tta_plan_end() == 0        # re-recording drops the old content
    assert lib.tta_plan_num_launches(h, 3) == 0 and lib.tta_plan_run(h, 3, 0) == 0 and lib.tta_step(h, 0) == 0
    assert lib.tta_plan_run(h, 8, 0) == 1
    assert lib.tta_workspace_bytes(2, 4, 64 ** 3) == 4 * lib.tta_norm_workspace_floats(2, 4, 64 ** 3)
    assert lib.tta_plan_destroy(h) == 0
