#!/usr/bin/env python
"""Stage the UNMODIFIED reference for the benchmark's reference arm.

    python baseline/vendor_reference.py          # /root/reference -> baseline/_ref/   (git-ignored, travels with gpurun)

The reference (robert-giaquinto/gradient-boosted-normalizing-flows) is a flat set of Python scripts without setup.py /
pyproject.toml, so `pip install --target baseline/_ref /root/reference` has nothing to install.  This script is the
equivalent: it copies the files the density path imports, byte for byte, into the git-ignored `baseline/_ref/` so that
`bench.py --impl reference` and the `gpu_torch_baseline` leg can run the reference's own `models.boosted_flow.BoostedFlow`
on the GPU box (where /root/reference does not exist).  Nothing under baseline/_ref is ever committed or imported by the
product; `python baseline/vendor_reference.py --check` verifies the copies against their sources.
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("GBNF_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
DIRS = ("models", "utils", "optimization")
FILES = ("density_experiment.py", "toy_experiment.py", "LICENSE", "requirements.txt")


def _pairs():
    for d in DIRS:
        for f in sorted(os.listdir(os.path.join(REF, d))):
            if f.endswith(".py"):
                yield os.path.join(REF, d, f), os.path.join(DST, d, f)
    for f in FILES:
        yield os.path.join(REF, f), os.path.join(DST, f)


def vendor(check=False):
    if not os.path.isdir(REF):
        raise SystemExit(f"{REF} not found: the staged copy under baseline/_ref (if present) is used as is")
    n = 0
    for src, dst in _pairs():
        if check:
            assert os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False), f"{dst} differs from {src}"
        else:
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
        n += 1
    print(f"{'verified' if check else 'staged'} {n} reference files in {DST}")
    return DST


if __name__ == "__main__":
    vendor(check="--check" in sys.argv[1:])
