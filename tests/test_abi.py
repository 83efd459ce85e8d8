"""The C-ABI library loads, exports every symbol include/gbnf.h declares, and refuses to run without a GPU."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest
import torch

import gbnf_b200
from gbnf_b200 import _lib
from oracle import gbnf_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "gbnf.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gbnf_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/gbnf.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header drifted apart"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.lib_path()], capture_output=True, text=True).stdout
    for n in names:
        assert re.search(rf"\bT {n}\b", out), n
    assert lib.gbnf_abi_version() == 2


def test_struct_layouts_match_header():
    assert C.sizeof(_lib.Config) == 12 * 4
    assert C.sizeof(_lib.StepParams) == (7 + 2 * 2 * _lib.GBNF_MAX_LAYERS + 3) * 8      # + invconv_w / _winv / _logdet (ABI 2)
    assert C.sizeof(_lib.ComponentParams) == 16
    assert C.sizeof(_lib.Info) == 6 * 4 + 2 * 8 + 2 * 4


def test_built_for_sm100a_with_lineinfo():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu():
    lib = _lib.load()
    h = C.c_void_p()
    cfg = _lib.Config(kind=1, D=43, h=512, K=5, C=8, depth=1)
    rc = lib.gbnf_create(C.byref(h), C.byref(cfg))
    assert rc == -2 and not h.value
    assert b"no CPU fallback" in lib.gbnf_last_error()


def test_argument_validation_without_gpu():
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.gbnf_create(C.byref(h), None) == -1
    for bad in (dict(kind=7), dict(D=1), dict(D=4000), dict(C=0), dict(depth=9), dict(act=5), dict(gemm_mode=3),
                dict(kind=1, act=2), dict(kind=1, act=3), dict(kind=0, act=3, depth=3), dict(kind=0, act=3, gemm_mode=1),
                dict(kind=0, glow_invconv=1), dict(kind=1, glow_invconv=1, gemm_mode=2), dict(kind=1, glow_invconv=2)):
        kw = dict(kind=1, D=43, h=512, K=5, C=8, depth=1)
        kw.update(bad)
        assert lib.gbnf_create(C.byref(h), C.byref(_lib.Config(**kw))) == -1, bad
        assert lib.gbnf_last_error()


def test_sample_component_host_helper_matches_oracle():
    lib = _lib.load()
    rng = np.random.default_rng(0)
    for _ in range(200):
        n = int(rng.integers(1, 12))
        rho = rng.random(n).astype(np.float32) + 0.01
        u = float(rng.random())
        exclude = int(rng.integers(-1, n)) if n > 1 else -1
        j = C.c_int32(-1)
        arr = (C.c_float * n)(*rho.tolist())
        assert lib.gbnf_sample_component(arr, n, u, exclude, C.byref(j)) == 0
        assert j.value == orc.sample_component(rho, n, u, exclude)
        assert j.value != exclude
    assert lib.gbnf_sample_component(None, 3, 0.5, -1, C.byref(j)) == -1
