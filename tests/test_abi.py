"""CPU-only checks of the drop-in boundary: libmorl_b200.so loads, exports exactly the symbols include/morl_b200.h
declares, the ctypes signature table mirrors the header, and argument errors are reported without touching a device."""

import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "morl_b200.h")


def _header_decls():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"MORL_API\s+([\w\s\*]+?)\s*\b(morl_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(3).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        decls[m.group(2)] = n
    return decls


def test_library_builds_and_loads():
    from morl_baselines_b200.csrc import build

    path = build.build()
    assert os.path.exists(path)
    from morl_baselines_b200 import _lib

    lib = _lib.load()
    assert lib.morl_version() == 100


def test_header_and_library_export_the_same_symbols():
    from morl_baselines_b200 import _lib

    decls = _header_decls()
    assert len(decls) >= 15
    lib = _lib.load()
    for name in decls:
        assert hasattr(lib, name), f"{name} declared in include/morl_b200.h but not exported by libmorl_b200.so"
    assert set(decls) == set(_lib.SIGNATURES), set(decls) ^ set(_lib.SIGNATURES)
    for name, nargs in decls.items():
        assert len(_lib.SIGNATURES[name][1]) == nargs, f"{name}: header has {nargs} parameters, ctypes table {len(_lib.SIGNATURES[name][1])}"


def test_exported_symbols_are_only_the_abi():
    import subprocess

    from morl_baselines_b200 import _lib

    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert exported == set(_header_decls()), exported ^ set(_header_decls())


def test_argument_errors_need_no_device():
    from morl_baselines_b200 import _lib

    lib = _lib.load()
    rc = lib.morl_envelope_td_f32(None, None, None, None, None, 0.99, 4, 4, 4, 3, 0, 0, None, None, None, None)
    assert rc == -1  # MORL_ERR_NULL
    assert b"NULL" in lib.morl_last_error()
    rc = lib.morl_pareto_mask_f32(16, 10, 9, 1, 16, None)  # D = 9 > MORL_MAX_D
    assert rc == -4
    with pytest.raises(_lib.MorlB200Error):
        _lib.check(rc, "morl_pareto_mask_f32")


def test_ops_refuse_cpu_tensors():
    import torch as th

    from morl_baselines_b200 import _lib, ops

    with pytest.raises(_lib.MorlB200Error):
        ops.pareto_mask(th.zeros(4, 2))
    with pytest.raises(_lib.MorlB200Error):
        ops.envelope_td(th.zeros(2, 2, 2, 3), th.zeros(2, 2, 2, 3), th.zeros(2, 3), th.zeros(2, 3), th.zeros(2), 0.99)


def test_fusion_coverage_predicates_need_no_device():
    """The `*_supported` predicates that decide between a fused kernel and the launches it replaces are pure host functions: their answers
    at the BASELINE shapes and just outside them (both sides are CUDA paths; an unsupported shape is refused by the fused entry point)."""
    from morl_baselines_b200 import _lib

    lib = _lib.load()
    F16, BF16 = _lib.FMT_F16X2, _lib.FMT_BF16X3
    # fused output layers + envelope + Bellman: north-star, configs[1] (minecart: |W| 32, |A| 6, d 3), and what falls outside
    assert lib.morl_qhead_envelope_supported(F16, 1024, 64, 8, 3, 256) == 1
    assert lib.morl_qhead_envelope_supported(F16, 256, 32, 6, 3, 256) == 1
    assert lib.morl_qhead_envelope_supported(BF16, 1024, 64, 8, 3, 256) == 0   # f16x2 planes only
    assert lib.morl_qhead_envelope_supported(F16, 1024, 48, 8, 3, 256) == 0    # |W| must divide 128
    assert lib.morl_qhead_envelope_supported(F16, 1024, 64, 11, 3, 256) == 0   # |A| d > 32 (and |W||A| not a multiple of 16)
    assert lib.morl_qhead_envelope_supported(F16, 1024, 64, 8, 5, 256) == 0    # d in 2..4
    assert lib.morl_qhead_envelope_supported(F16, 1024, 64, 8, 3, 320) == 0    # K <= 256
    assert lib.morl_qhead_gemm_supported(F16, 65536, 24, 256) == 1 and lib.morl_qhead_gemm_supported(F16, 65536, 40, 256) == 0
    assert lib.morl_qhead_gemm_supported(F16, 1000, 24, 256) == 0              # M % 128
    # chained hidden layers: 256-wide layers, at least two 128-row tiles
    assert lib.morl_gemm_chain_supported(F16, 65536, 256) == 1 and lib.morl_gemm_chain_supported(BF16, 8192, 256) == 1
    assert lib.morl_gemm_chain_supported(F16, 65536, 128) == 0 and lib.morl_gemm_chain_supported(F16, 128, 256) == 0
    # refused with an error code and a message, not a crash
    rc = lib.morl_qhead_envelope_td_f32(F16, 16, 16, 0, None, None, 16, 16, 0, None, None, None, None, 256, 16, 16, 16, 0.99, 1024, 48, 8, 3, 0, 0, 0, 16, None, None, None,
                                        None, None)
    assert rc == -4 and b"unsupported configuration" in lib.morl_last_error()
