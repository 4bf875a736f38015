"""CPU suite, part 2: the C-ABI library and the host-side mirror, without any compute.

* libgcanet_b200.so loads and exports every symbol include/gcanet_b200.h declares;
* argument validation returns status codes + messages (never exits, never touches a device);
* the Python front end refuses CPU tensors instead of falling back.
"""
import ctypes
import os
import re

import pytest
import torch

import gcanet_b200 as gb
from gcanet_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "gcanet_b200.h")).read()
    return sorted(set(re.findall(r"GCANET_API[^;(]*?\b(gcanet_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = _declared_symbols()
    assert len(names) >= 20
    raw = ctypes.CDLL(_cabi.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} is declared in include/gcanet_b200.h but not exported"
    # and the ctypes table covers the whole header (no entry point left unbound)
    assert sorted(_cabi.SIGNATURES) == names


def test_abi_version_and_status_strings():
    L = _cabi.lib()
    assert L.gcanet_abi_version() == 2
    assert L.gcanet_status_string(0) == b"ok"
    assert b"workspace" in L.gcanet_status_string(-2)


def test_scan_probe_read_without_a_bracketed_scan_is_an_error():
    L = _cabi.lib()
    ms = ctypes.c_float(0.0)
    assert L.gcanet_knn_probe_arm(0) == 0                      # disarming creates nothing and needs no device
    assert L.gcanet_knn_probe_read(ctypes.byref(ms)) == -1
    assert b"probe" in L.gcanet_last_error()
    assert L.gcanet_knn_probe_read(None) == -1


def test_knn_columns_follow_reference_dilation():
    L = _cabi.lib()
    import numpy as np
    for k1, k2 in [(20, 20), (10, 20), (50, 80), (7, 20), (1, 1), (3, 10)]:
        assert L.gcanet_knn_graph_columns(k1, k2) == len(np.arange(0, k2, k2 // k1))   # M4:32
    assert L.gcanet_knn_graph_columns(0, 5) == 0 and L.gcanet_knn_graph_columns(6, 5) == 0


def test_argument_validation_without_device():
    L = _cabi.lib()
    null = ctypes.c_void_p(0)
    one = ctypes.c_void_p(256)     # never dereferenced: validation fails first
    assert L.gcanet_knn_graph(null, 1, 3, 10, 2, 2, 0, null, null, null, 0, null) == -1
    assert b"null pointer" in L.gcanet_last_error()
    assert L.gcanet_knn_graph(one, 1, 3, 10, 20, 20, 0, one, null, one, 1 << 20, null) == -1
    assert b"exceeds the number of points" in L.gcanet_last_error()
    assert L.gcanet_knn_graph(one, 1, 3, 10, 2, 2, 1, one, null, one, 1 << 20, null) == -1
    assert b"C = 6" in L.gcanet_last_error()
    assert L.gcanet_knn_graph(one, 1, 3, 10, 2, 2, 0, one, null, null, 0, null) == -2     # workspace
    assert L.gcanet_knn_cuda(one, 5, one, 5, 3, 6, 1, 0, one, one, null, 0, null) == -1
    d = _cabi.EdgeConvDesc(2, 100, 64, 64, 48, 20, 2, 1e-5, 0.2)                             # Cout % 32 != 0
    assert L.gcanet_edgeconv_saved_bytes(ctypes.byref(d)) == 0
    assert b"Cout" in L.gcanet_last_error()
    d = _cabi.EdgeConvDesc(16, 10000, 64, 64, 128, 50, 2, 1e-5, 0.2)
    saved = L.gcanet_edgeconv_saved_bytes(ctypes.byref(d))
    # [P|Q] + ysel + ysum (fp32) + arg (u8) per (point, channel): 4*(2+1+1)+1 = 17 bytes
    assert abs(saved - 16 * 10000 * 128 * 17) < 1 << 16
    assert L.gcanet_edgeconv_workspace_bytes(ctypes.byref(d)) > 0


def test_flag_constants_match_the_header_and_unknown_flags_are_rejected():
    from gcanet_b200 import functional as G
    text = open(os.path.join(ROOT, "include", "gcanet_b200.h")).read()
    for name, val in (("GCANET_KNN_FLAG_BRUTE_FORCE", G.KNN_FLAG_BRUTE_FORCE), ("GCANET_KNN_FLAG_UNORDERED", G.KNN_FLAG_UNORDERED),
                      ("GCANET_KNN_FLAG_NO_PRUNE", G.KNN_FLAG_NO_PRUNE)):
        m = re.search(r"#define\s+" + name + r"\s+(0x[0-9a-fA-F]+)", text)
        assert m and int(m.group(1), 16) == val, name
    L = _cabi.lib()
    null, one = ctypes.c_void_p(0), ctypes.c_void_p(256)
    # every known flag combination passes validation (and then fails on the missing workspace, never on the flags)
    for flags in (0x100, 0x200, 0x400, 0x600, 0x700):
        assert L.gcanet_knn_graph(one, 1, 64, 2048, 20, 20, flags, one, null, null, 0, null) == -2
        assert L.gcanet_knn_graph_workspace_bytes(1, 64, 2048, 20, flags) > 0
    assert L.gcanet_knn_graph(one, 1, 64, 2048, 20, 20, 0x800, one, null, one, 1 << 30, null) == -1
    assert b"unknown flag" in L.gcanet_last_error()


def test_cpu_tensors_are_refused():
    x = torch.randn(1, 3, 32)
    for fn in (lambda: gb.knn(x, 4, 4), lambda: gb.get_graph_feature(x, 4, 4),
               lambda: gb.KNN(2)(x, x), lambda: gb.grouping_operation(x, torch.zeros(1, 2, 2, dtype=torch.int32)),
               lambda: gb.DGCNNEncoderGn(mode=0, nn_nb=4, input_channels=6)(x)):
        with pytest.raises(RuntimeError):
            fn()


def test_encoder_state_dict_keys_match_reference_layout():
    enc = gb.DGCNNEncoderGn(mode=5, nn_nb=80, input_channels=6)
    sd = enc.state_dict()
    assert tuple(sd["conv1.0.weight"].shape) == (64, 12, 1, 1)
    assert tuple(sd["conv2.0.weight"].shape) == (64, 128, 1, 1)
    assert tuple(sd["conv3.0.weight"].shape) == (128, 128, 1, 1)
    for key in ("bn1.weight", "bn2.bias", "bn3.weight", "bn4.weight", "bn5.bias", "conv1.1.weight",
                "mlp1.weight", "mlp1.bias", "bnmlp1.weight"):
        assert key in sd
    from oracle import dgcnn_oracle as orc
    ref = orc.DGCNNEncoderGn(mode=5, nn_nb=80, input_channels=6)
    assert list(ref.state_dict()) == list(sd)
    enc.load_state_dict(ref.state_dict())


def test_dispatcher_ops_registered_with_shape_inference():
    """gcanet_b200.torch_ops: the operators exist under torch.ops.gcanet_b200 and their fake implementations infer the
    reference's shapes / dtypes on meta tensors (no kernel runs, no GPU needed)."""
    import torch
    import gcanet_b200.torch_ops  # noqa: F401
    ops = torch.ops.gcanet_b200
    x = torch.empty(2, 64, 1000, device="meta")
    idx = ops.knn_graph(x, 10, 40, 0, True)
    assert idx.shape == (2, 1000, 10) and idx.dtype == torch.int64
    d, i = ops.knn_cuda(torch.empty(2, 3, 500, device="meta"), torch.empty(2, 3, 70, device="meta"), 4)
    assert d.shape == (2, 4, 70) and i.dtype == torch.int64
    g = ops.group_points(torch.empty(2, 16, 500, device="meta"), torch.empty(2, 70, 4, dtype=torch.int32, device="meta"))
    assert g.shape == (2, 16, 70, 4)
    o_nc, o_cn, saved = ops.edgeconv_forward(torch.empty(2, 1000, 64, device="meta"),
                                             torch.empty(2, 1000, 20, dtype=torch.int32, device="meta"),
                                             torch.empty(128, 128, device="meta"), torch.empty(128, device="meta"),
                                             torch.empty(128, device="meta"), 64, 2, 1e-5, 0.2)
    assert o_nc.shape == (2, 1000, 128) and o_cn.shape == (2, 128, 1000) and saved.dtype == torch.uint8 and saved.numel() > 0


def test_sppnet_encoder_variant_keeps_that_files_constructor_and_keys():
    """models/sppnet.py:148-175: (mode, input_channels, nn_nb), conv1 takes 2 * input_channels, same parameter names."""
    enc = gb.SppnetDGCNNEncoderGn(0, 3, 40)
    assert enc.k == 40 and enc.mode == 0 and enc.conv1[0].weight.shape == (64, 6, 1, 1)
    enc5 = gb.SppnetDGCNNEncoderGn(mode=5, input_channels=6, nn_nb=80)
    assert enc5.conv1[0].weight.shape == (64, 12, 1, 1)
    assert set(enc.state_dict()) == set(gb.DGCNNEncoderGn(mode=0, nn_nb=40, input_channels=6).state_dict())
    with pytest.raises(RuntimeError, match="no CPU path"):
        enc(torch.zeros(1, 3, 64))
