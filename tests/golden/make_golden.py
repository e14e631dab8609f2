#!/usr/bin/env python
"""Generates the parity fixtures.  Run in the build container (needs /root/reference and the
reference binaries of oracle/ref_build/build_all.sh):

    python tests/golden/make_golden.py [case-name ...]

  tests/golden/short/<case>.c<NNNN>.dmp   state after N cycles, written by the reference's own
                                          Kokkos-OpenMP build run on the case's ini
  tests/golden/reference/<dir>/<file>     copies of the golden dumps the reference ships for
                                          its regression suite (data files, /root/reference/reference)
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from cases import FULL, HOST_SHORT, ORACLE_FULL, ORACLE_SHORT, SHORT, ini_for  # noqa: E402
from oracle.ref_run import cleanup, run_reference  # noqa: E402

REFERENCE_GOLDENS = "/root/reference/reference"


def main():
    only = set(sys.argv[1:])
    os.makedirs(os.path.join(HERE, "short"), exist_ok=True)
    for case in SHORT + HOST_SHORT + ORACLE_SHORT:
        if only and case.name not in only:
            continue
        wanted = [0] + list(case.cycles)
        if all(os.path.exists(os.path.join(HERE, "short", f"{case.name}.c{n:04d}.dmp")) for n in wanted):
            continue
        # Axis::SymmetrizeEx1Side adds the radial EMFs around the pole with atomics (axis.cpp:39-42): only a single
        # thread gives a reproducible (ascending-k) order, and that is the order the CUDA kernel uses
        threads = 1 if case.setup in ("axis", "axis25d") else None
        res = run_reference(case.ref, ini_for(case, dump_every_cycle=True), maxcycles=max(case.cycles), keep=True, timeout=600,
                            threads=threads)
        for n in wanted:
            src = os.path.join(res["workdir"], f"dump.{n:04d}.dmp")
            shutil.copy(src, os.path.join(HERE, "short", f"{case.name}.c{n:04d}.dmp"))
        cleanup(res)
        print("generated", case.name, flush=True)
    for case in FULL + ORACLE_FULL:
        if only and case.name not in only:
            continue
        dst = os.path.join(HERE, "reference", case.golden)
        if not os.path.exists(dst):
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copy(os.path.join(REFERENCE_GOLDENS, case.golden), dst)
            os.chmod(dst, 0o644)
            print("copied", case.golden, flush=True)


if __name__ == "__main__":
    main()
