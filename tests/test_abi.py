"""The C-ABI library builds, loads and exports every symbol include/d2p.h
declares (no compute calls: runs without a GPU)."""
import os
import re

from demo2program_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, 'include', 'd2p.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(d2p_[a-z0-9_]+)\s*\(', src)))


def test_header_declares_functions():
    fns = _header_functions()
    assert 'd2p_conv_encoder_fwd' in fns and 'd2p_clip_adam_step' in fns
    assert len(fns) >= 25


def test_library_exports_every_declared_symbol(lib):
    for name in _header_functions():
        assert hasattr(lib, name), 'libd2p.so does not export ' + name


def test_ctypes_signatures_cover_header(lib):
    assert sorted(_lib.SIGNATURES) == _header_functions()


def test_version_and_error_string(lib):
    assert lib.d2p_version() >= 100
    assert lib.d2p_last_error() is not None


def test_argument_errors_are_reported_not_thrown(lib):
    # null buffers -> D2P_ERR_ARG with a message; needs no GPU
    rc = lib.d2p_axpby(None, 1.0, None, 0.0, 4, None)
    assert rc == -1
    assert b'axpby' in lib.d2p_last_error()
    rc = lib.d2p_seq_weights(None, 4, 3, 1.0, 5, None, None, None)
    assert rc == -1


def test_sizes_query_without_gpu(lib):
    import ctypes as C
    d = _lib.ConvDesc()
    d.B, d.k, d.T, d.h, d.w, d.d = 32, 10, 20, 8, 8, 16
    d.frames_dtype, d.n_layers = _lib.D2P_U8, 3
    for i, c in enumerate((16, 32, 48)):
        d.layers[i].cout = c
    assert lib.d2p_conv_encoder_feature_dim(C.byref(d)) == 48
    n = 6400
    want = n * (4 * 4 * 16 + 2 * 2 * 32 + 48) + 4 * 10 * (16 + 32 + 48)
    assert lib.d2p_conv_encoder_saved_floats(C.byref(d)) == want
    assert lib.d2p_conv_encoder_ws_bytes(C.byref(d)) > 0
    # ViZDoom geometry: 80 -> 40 -> 20 -> 10 -> 5 -> 3, feature 3*3*48
    d.h = d.w = 80
    d.d, d.n_layers = 3, 5
    for i, c in enumerate((16, 32, 48, 48, 48)):
        d.layers[i].cout = c
    assert lib.d2p_conv_encoder_feature_dim(C.byref(d)) == 432


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', str(tmp_path / 'nope.so'))
    import pytest
    with pytest.raises(_lib.D2PError):
        _lib.load()


def test_integration_md_lists_every_entry_point():
    """INTEGRATION.md section 6 is the map from each C-ABI entry point to the reference site it replaces."""
    doc = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    missing = [f for f in _header_functions() if f not in doc]
    assert not missing, missing
