"""The C-ABI library loads without a GPU and exports every symbol include/rtg_splat_b200.h declares; argument
validation that happens before any CUDA call behaves as documented. No compute calls here."""
import ctypes as C
import os

import pytest

from rtg_slam_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_present_and_loads():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    L = _lib.lib()
    assert L.rtg_version() >= 100


def test_exports_every_header_symbol():
    L = _lib.lib()
    syms = _lib.header_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/rtg_splat_b200.h but not exported"
    assert set(syms) == set(_lib.SIGNATURES), "ctypes signature table and header disagree"


def test_view_struct_layout_matches_header():
    # 14 4-byte scalars then 4 pointers
    assert C.sizeof(_lib.RtgSplatView) == 14 * 4 + 4 * 8
    assert _lib.RtgSplatView.viewmatrix.offset == 56
    assert C.sizeof(_lib.RtgAdamGroup) == 48
    assert C.sizeof(_lib.RtgHistoryMerge) == 2 * 4 + 12 * 8 + 4 * 4 and _lib.RtgHistoryMerge.features_dc_stride.offset == 104


def test_workspace_bytes_is_host_only_and_monotone():
    L = _lib.lib()
    g, i, b = C.c_size_t(), C.c_size_t(), C.c_size_t()
    assert L.rtg_splat_workspace_bytes(1000, 480, 640, 10000, C.byref(g), C.byref(i), C.byref(b)) == 0
    g2, i2, b2 = C.c_size_t(), C.c_size_t(), C.c_size_t()
    assert L.rtg_splat_workspace_bytes(2000, 480, 640, 20000, C.byref(g2), C.byref(i2), C.byref(b2)) == 0
    assert g2.value > g.value and b2.value > b.value and i2.value == i.value
    assert g.value >= 1000 * 76 and i.value >= 480 * 640 * 4 and b.value >= 10000 * 12
    assert L.rtg_splat_workspace_bytes(-1, 480, 640, 0, None, None, None) == -1
    assert b"bad sizes" in L.rtg_last_error()


def test_forward_rejects_null_view_before_touching_cuda():
    L = _lib.lib()
    rc = L.rtg_splat_forward(None, 0, 0, *([None] * 8), None, None, None, 0, *([None] * 8), None, None, None, None)
    assert rc == -1 and b"view is NULL" in L.rtg_last_error()


def test_adam_rejects_bad_groups():
    L = _lib.lib()
    assert L.rtg_adam_step(None, 0, 0.9, 0.999, 1e-15, 1, None) == -1
    arr = (_lib.RtgAdamGroup * 1)()
    assert L.rtg_adam_step(arr, 1, 0.9, 0.999, 1e-15, 0, None) == -1  # step must be >= 1


def test_side_entry_points_validate_before_touching_cuda():
    """The widened rows and the two-call backward reject bad arguments on the host."""
    L = _lib.lib()
    nul = [None]
    assert L.rtg_accumulate_gaussian_error(-1, 4, 4, *(nul * 5), 0.1, 0.1, 0.1, 1, *(nul * 6)) == -1
    assert L.rtg_accumulate_gaussian_error(4, 4, 0, *(nul * 5), 0.1, 0.1, 0.1, 1, *(nul * 6)) == 0      # P == 0: nothing to do
    assert L.rtg_accumulate_gaussian_error(4, 4, 3, *(nul * 5), 0.1, 0.1, 0.1, 1, *(nul * 6)) == -1
    assert b"NULL output" in L.rtg_last_error()
    assert L.rtg_tile_mean(16, 16, None, 0.5, None, None, None) == -1
    assert L.rtg_transmission_tile_mask(0, 16, None, 0.5, None, None, None) == -1
    assert L.rtg_color_error(16, 16, None, None, None, None) == -1
    assert L.rtg_frame_preprocess(None, 16, 16, 1, 5, 2.0, 2.0, 0.3, 5.0, 1.0, 1.0, 0.0, 0.0, 0.2, *(nul * 7)) == -1
    assert b"rtg_frame_preprocess" in L.rtg_last_error()
    assert L.rtg_loss_l1(*(nul * 6), 16, 16, 0, 0.8, 1.0, 0.1, *(nul * 5)) == -1
    assert L.rtg_normal_map(None, None, 16, 16, None, None) == -1
    assert L.rtg_map_history_merge(None, None) == -1 and b"merge is NULL" in L.rtg_last_error()
    hm = _lib.RtgHistoryMerge()
    hm.P, hm.max_weight = 0, 0.5
    assert L.rtg_map_history_merge(C.byref(hm), None) == 0          # empty map: nothing to do
    hm.P = 4
    assert L.rtg_map_history_merge(C.byref(hm), None) == -1 and b"NULL pointer" in L.rtg_last_error()
    hm.max_weight = 0.0
    assert L.rtg_map_history_merge(C.byref(hm), None) == 0          # disabled (mapper.py:213-214)
    assert L.rtg_ssim_loss(None, None, 3, 16, 16, None, None, None, None) == -1 and b"rtg_ssim_loss" in L.rtg_last_error()
    # partial sums (one double per 16x16 tile and channel, 256-byte aligned) + three derivative maps
    assert L.rtg_ssim_workspace_bytes(3, 680, 1200) == ((3 * 43 * 75 * 8 + 255) // 256) * 256 + 3 * 3 * 680 * 1200 * 4
    assert L.rtg_ssim_workspace_bytes(0, 16, 16) == 0
    for fn in (L.rtg_splat_backward, L.rtg_splat_backward_render, L.rtg_splat_backward_finish):
        assert fn(None, 1, 16, *(nul * 6), *(nul * 4), 0, None, *(nul * 4), None, *(nul * 8), None) == -1
        assert b"view is NULL" in L.rtg_last_error()


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.RtgError, match="no CPU or PyTorch fallback"):
        _lib.lib()
