#!/usr/bin/env python
"""Generates the config-size parity fixtures.  Run in the build container (needs the reference binaries of
oracle/ref_build/build_all.sh, i.e. /root/reference):

    python tests/golden/make_config_size.py [case ...]

BASELINE.json quotes its configurations at sizes whose dumps are far too large to commit (C2: 1024^2, C3: 128^3,
C4: 256^3 per GPU, C5: 96 x 48 x 96 here).  For each of them the reference's own Kokkos-OpenMP build runs a few
cycles at that size and this script records, per dumped field, the SHA-256 of its bytes (and min / max / mean as a
human-readable cross-check) in tests/golden/config_size/<case>.json.  Bit-identical arrays <=> equal digests, so
the GPU test (tests/test_config_size_gpu.py) checks the same thing a full dump comparison would.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from cases import CONFIG_SIZE, ini_for  # noqa: E402
from idefix_b200.dump import readDump  # noqa: E402
from oracle.ref_run import cleanup, run_reference  # noqa: E402


def digest(arr):
    a = np.ascontiguousarray(np.asarray(arr, dtype=np.float64))
    return {"sha256": hashlib.sha256(a.tobytes()).hexdigest(), "shape": list(a.shape), "min": float(a.min()),
            "max": float(a.max()), "mean": float(a.mean())}


def main():
    only = set(sys.argv[1:])
    out_dir = os.path.join(HERE, "config_size")
    os.makedirs(out_dir, exist_ok=True)
    for case in CONFIG_SIZE:
        if only and case.name not in only:
            continue
        path = os.path.join(out_dir, f"{case.name}.json")
        if os.path.exists(path):
            continue
        ncyc = max(case.cycles)
        res = run_reference(case.ref, ini_for(case, dump_every_cycle=True), maxcycles=ncyc, keep=True, timeout=3000)
        rec = {"case": case.name, "reference_binary": case.ref, "cycles": {}}
        for n in case.cycles:
            d = readDump(os.path.join(res["workdir"], f"dump.{n:04d}.dmp"))
            rec["cycles"][str(n)] = {"time": float(d.data["time"][0]), "dt": float(d.data["dt"][0]),
                                     "fields": {k: digest(v) for k, v in d.data.items() if k.startswith(("Vc-", "Vs-"))}}
        cleanup(res)
        with open(path, "w") as fh:
            json.dump(rec, fh, indent=1, sort_keys=True)
        print("generated", case.name, flush=True)


if __name__ == "__main__":
    main()
