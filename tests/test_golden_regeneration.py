"""Provenance of the fixtures: regenerating them from the unmodified reference (tests/golden/make_golden.py) reproduces the committed
files bit for bit.  Needs /root/reference (this container only; skipped on the GPU box)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference tree is not on this machine")
@pytest.mark.parametrize("name", ["realnvp_d5_mixed", "glow_d6_invconv", "realnvp_d6_residual"])
def test_fixture_regenerates_bit_identically(tmp_path, name):
    code = ("import sys; sys.argv=['make_golden.py']; import importlib.util as u; "
            f"s=u.spec_from_file_location('mg', r'{os.path.join(GOLDEN, 'make_golden.py')}'); m=u.module_from_spec(s); "
            f"s.loader.exec_module(m); m.build_case('{name}', out_dir=r'{tmp_path}')")
    res = subprocess.run([sys.executable, "-W", "ignore", "-c", code], capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, PYTHONDONTWRITEBYTECODE="1"))
    assert res.returncode == 0, res.stderr[-2000:]
    new = dict(np.load(os.path.join(str(tmp_path), name + ".npz")))
    old = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    assert set(old) <= set(new)          # (fixtures written before a later key was added to the script carry fewer keys)
    for k in old:
        assert new[k].dtype == old[k].dtype and new[k].shape == old[k].shape, k
        assert np.array_equal(new[k], old[k], equal_nan=True) if old[k].dtype.kind == "f" else np.array_equal(new[k], old[k]), k
