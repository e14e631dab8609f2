"""Runs a reference binary built by oracle/ref_build (the reference's OWN Kokkos-OpenMP CPU
implementation, compiled from /root/reference) on an ini text and returns its dumps.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, as the checker or the timed CPU baseline -- never by the
product path (idefix_b200/ does not import anything from oracle/).
"""
import os
import re
import shutil
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def have(name):
    return os.access(os.path.join(REF_DIR, name, "idefix"), os.X_OK)


def run_reference(name, ini_text, maxcycles=None, threads=None, keep=False, timeout=3600, nowrite=False):
    """-> dict(workdir, dumps=[paths sorted], perf=cell updates/s or None, log=str).

    With ``dump_every_cycle`` style inis ([Output] dmp 1e-7) dump.000N is the state after N cycles."""
    exe = os.path.join(REF_DIR, name, "idefix")
    if not os.access(exe, os.X_OK):
        raise FileNotFoundError(f"{exe}: build it with oracle/ref_build/build_all.sh (needs /root/reference)")
    wd = tempfile.mkdtemp(prefix=f"idfx_ref_{name}_")
    with open(os.path.join(wd, "idefix.ini"), "w") as fh:
        fh.write(ini_text)
    cmd = [exe, "-i", "idefix.ini", "-nolog"]
    if maxcycles is not None:
        cmd += ["-maxcycles", str(maxcycles)]
    if nowrite:
        cmd += ["-nowrite"]
    env = dict(os.environ)
    nthreads = threads or os.cpu_count() or 1
    env.update(OMP_NUM_THREADS=str(nthreads), OMP_PROC_BIND="spread", OMP_PLACES="threads")
    res = subprocess.run(cmd, cwd=wd, env=env, capture_output=True, text=True, timeout=timeout)
    log = res.stdout + res.stderr
    if res.returncode != 0:
        raise RuntimeError(f"reference run failed ({res.returncode}):\n{log[-3000:]}")
    m = re.search(r"Main: Perfs are ([0-9.eE+-]+) cell updates/second", log)
    dumps = sorted(os.path.join(wd, f) for f in os.listdir(wd) if f.endswith(".dmp"))
    out = {"workdir": wd, "dumps": dumps, "perf": float(m.group(1)) if m else None, "log": log,
           "threads": nthreads}
    if not keep and not dumps:
        shutil.rmtree(wd, ignore_errors=True)
    return out


def cleanup(result):
    shutil.rmtree(result["workdir"], ignore_errors=True)
