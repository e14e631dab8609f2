"""Diagnostic (GPU box): runs axis cases through the C++ host and lists, per dump and field, where they differ from
the reference build's dump.  usage: python tools/axis_diag.py <case> [<case> ...]"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cases import HOST_SHORT, ini_for  # noqa: E402
from idefix_b200.dump import readDump  # noqa: E402

HOST = os.path.join(ROOT, "idefix_b200", "host")
built = {}


def exe_for(order, vp):
    key = (order, vp)
    if key not in built:
        out = os.path.join(tempfile.mkdtemp(), "idefix")
        cmd = ["make", "-s", "-C", HOST, f"PROBLEM={os.path.join(HOST, 'problems', 'AxisMHD')}", "MHD=YES", f"ORDER={order}",
               f"OUT={out}"] + (["VP=1"] if vp else [])
        subprocess.run(cmd, check=True, capture_output=True, text=True)
        built[key] = out
    return built[key]


for name in sys.argv[1:]:
    case = [c for c in HOST_SHORT if c.name == name][0]
    wd = tempfile.mkdtemp()
    with open(os.path.join(wd, "idefix.ini"), "w") as fh:
        fh.write(ini_for(case, dump_every_cycle=True))
    res = subprocess.run([exe_for(case.order, case.ref.endswith("_vp")), "-i", "idefix.ini", "-maxcycles", str(max(case.cycles))],
                         cwd=wd, capture_output=True, text=True, timeout=300)
    print(f"=== {name}: rc={res.returncode}")
    for n in (0,) + tuple(case.cycles):
        mine = readDump(os.path.join(wd, f"dump.{n:04d}.dmp"))
        gold = readDump(os.path.join(ROOT, "tests", "golden", "short", f"{name}.c{n:04d}.dmp"))
        for f, ref in gold.data.items():
            if not f.startswith(("Vc-", "Vs-", "Ve-")) or f not in mine.data:
                continue
            a, b = np.asarray(mine.data[f]), np.asarray(ref)
            bad = np.argwhere(a != b)
            if len(bad):
                ii, jj, kk = (sorted(set(bad[:, d].tolist())) for d in range(3))
                print(f"  c{n:04d} {f}: {len(bad)}/{a.size} differ, max {np.abs(a - b).max():.2e} (scale {np.abs(b).max():.2e}); "
                      f"i {ii[:8]}{'...' if len(ii) > 8 else ''} j {jj[:8]} k {kk[:8]}{'...' if len(kk) > 8 else ''} of shape {a.shape}")
            else:
                print(f"  c{n:04d} {f}: equal")
