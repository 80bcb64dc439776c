#!/usr/bin/env python3
"""Stages the UNMODIFIED reference sources of the headline path under oracle/_ref/ (git-ignored; it travels to the GPU box with
the snapshot, where /root/reference does not exist) so that `bench.py --impl reference` can time the reference's own numba
code: interpolator.py, solve_triangular.py and sn/pantheon.py (BASELINE.json config 0).  Nothing is edited: the data loader
the script imports (y2022pantheonSHOES.data, whose covariance blob is missing from the checkout) is pre-seeded
in sys.modules by the caller.  Run by __graft_entry__.build() whenever the reference checkout is present."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("COSMO_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ["interpolator.py", "solve_triangular.py", os.path.join("sn", "pantheon.py")]


def stage():
    if not os.path.isdir(REF):
        return False
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
    return True


if __name__ == "__main__":
    print("staged" if stage() else f"{REF} not present: nothing staged", file=sys.stderr)
