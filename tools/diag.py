"""GPU diagnostic: prints parity of every SHORT case (no assertion); used while bringing kernels up."""
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity  # noqa: E402
from cases import SHORT  # noqa: E402
from idefix_b200.dump import readDump  # noqa: E402

only = sys.argv[1:]
for case in SHORT:
    if only and not any(o in case.name for o in only):
        continue
    for fused in (False, True):
        try:
            s = parity.make(case)
            done = 0
            line = []
            for n in case.cycles:
                parity.advance_cycles(s, n - done, fused)
                done = n
                res = parity.compare_with_dump(s, readDump(f"{ROOT}/tests/golden/short/{case.name}.c{n:04d}.dmp"))
                worst = max(res.items(), key=lambda kv: kv[1][0])
                line.append(f"c{n}: {worst[0]} relL1={worst[1][0]:.2e} max={worst[1][1]:.2e}")
            divb = s.data.hydro.CheckDivB() if s.definitions.mhd else 0.0
            print(f"{case.name:32s} {'fused' if fused else 'gran '} | " + " | ".join(line) + f" | divB={divb:.1e}", flush=True)
            s.data.close()
        except Exception as e:
            print(f"{case.name:32s} {'fused' if fused else 'gran '} | EXC {type(e).__name__}: {str(e)[:200]}", flush=True)
            traceback.print_exc(limit=2)
