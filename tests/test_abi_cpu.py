"""CPU: the C-ABI library builds for sm_100a, loads, exports every symbol include/cf_b200.h declares, and every
op fails loudly (no CPU / PyTorch fallback) when there is no sm_100 device."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "cf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"CF_API\s+[\w\s\*]+?\b(cf_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol(dcf):
    assert os.path.exists(dcf.SO_PATH), "build it: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(dcf.SO_PATH)
    names = _declared_symbols()
    assert len(names) >= 19 and "cf_fusion_fwd" in names and "cf_nms_sat" in names
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/cf_b200.h but not exported"
    # the Python signature table binds exactly the declared surface
    assert sorted(dcf._lib.SIGNATURES) == names


def _declared_prototypes():
    """name -> list of C parameter declarations, parsed from the header."""
    text = open(os.path.join(ROOT, "include", "cf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"CF_API\s+[\w\s\*]+?\b(cf_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        args = " ".join(m.group(2).split())
        protos[m.group(1)] = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
    return protos


def _kind(c_decl):
    """Coarse class of a C parameter: pointer / i32 / i64 / f32 / size."""
    if "*" in c_decl:
        return "ptr"
    for key, kind in (("int64_t", "i64"), ("int32_t", "i32"), ("size_t", "size"), ("float", "f32"), ("double", "f64"),
                      ("int", "i32")):
        if re.search(rf"\b{key}\b", c_decl):
            return kind
    raise AssertionError(f"unclassified parameter: {c_decl}")


def test_ctypes_signatures_match_the_header(dcf):
    """Every entry of the ctypes table has the argument COUNT and the argument KINDS (pointer / int32 / int64 / float /
    size_t, in order) of its prototype in include/cf_b200.h -- a stale table would shift every later argument."""
    protos = _declared_prototypes()
    table = dcf._lib.SIGNATURES
    assert sorted(protos) == sorted(table)

    def ctype_kind(t):
        if t is ctypes.c_void_p or t is ctypes.c_char_p or (isinstance(t, type) and issubclass(t, ctypes._Pointer)):
            return "ptr"
        return {ctypes.c_int32: "i32", ctypes.c_int: "i32", ctypes.c_int64: "i64", ctypes.c_longlong: "i64",
                ctypes.c_float: "f32", ctypes.c_double: "f64", ctypes.c_size_t: "size"}[t]

    for name, params in protos.items():
        _, argtypes = table[name]
        assert len(argtypes) == len(params), f"{name}: header has {len(params)} parameters, ctypes table {len(argtypes)}"
        for i, (decl, t) in enumerate(zip(params, argtypes)):
            assert _kind(decl) == ctype_kind(t), f"{name}: parameter {i} `{decl}` bound as {t}"


def test_abi_version_and_error_channel(dcf):
    lib = dcf.load()
    assert lib.cf_abi_version() == 2
    assert isinstance(lib.cf_last_error(), bytes)
    assert lib.cf_fusion_workspace_bytes(128, 0, 2, 96, 64) >= 2 * 2 * 128 * 128 * 2 + 2 * 96 * 64 * 4
    assert lib.cf_nms_workspace_bytes(2, 2048) >= 2 * 2048 * 32 * 8
    assert lib.cf_bucket_workspace_bytes(4, 140, 124) == 4 * 140 * 124 * 4
    assert lib.cf_gather_workspace_bytes(1, 128, 120, 160, 1) == 0   # channels_last needs no re-layout


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_fallback_without_gpu(dcf):
    lib = dcf.load()
    assert lib.cf_device_check() == -3                     # CF_ERR_ARCH
    assert b"no CPU fallback" in lib.cf_last_error()
    pts = torch.zeros(1, 8, 3)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        dcf.ops.bucket_points(pts, torch.tensor([3]), dcf.ops.BucketGrid(0, 0, 1, 4, 4))
    layer = dcf.ContinuousFusion(32, 32, geom=(0, 0, 1, 1))
    with pytest.raises(RuntimeError):
        layer(torch.zeros(1, 32, 4, 4), torch.zeros(1, 32, 4, 4), pts, torch.tensor([3]))
    with pytest.raises(RuntimeError):
        dcf.PostProcess(device="cpu").nms_sat_indices([torch.zeros(3, 7)])


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "deep_continuous_fusion_for_multi-sensor_3d_object_detection_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the CPU oracle", "").replace("fp32 oracle", "").replace(
                    "brute-force oracle", "").lower() or f == "synthetic.py", f"{f} mentions the oracle"
