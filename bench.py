#!/usr/bin/env python
"""bench.py -- cell-updates/s of the explicit Godunov MHD update (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--block B] [--emf NAME]

Workload (config.workload): the reference's 3-D Orszag-Tang vortex (test/MHD/OrszagTang3D, the
set-up of its published benchmark, doc/source/performances.rst:5-10): HLLD + PLM + constrained
transport, RK2, periodic, FP64, B^3 interior cells PER GPU (default 512^3 = BASELINE config 4 on
one GPU).  N > 1 (launched by torchrun, one rank per GPU) block-decomposes 2x1x1 / 2x2x1 / 2x2x2
with the halo exchange that replaces Mpi::ExchangeX1/X2/X3 -- weak scaling, per-GPU block fixed.
A "step" is one full RK cycle (all stages) = one cell update of every interior cell
(src/main.cpp:237-238).

`value`  : device-resident cycles timed with CUDA events on the launching stream, max over ranks.
`e2e`    : the same cycle through the public host API with HOST buffers: pinned-host -> device
           copy of the state (Vc, Vs) a set-up hands over, one host-synchronous cycle, device -> host
           read of its result (t, dt, status; --e2e-full-state: the whole evolved state too), every
           step, all inside the timed region.
`roofline`: the dominant kernel (the slowest directional sweep), algorithmic bytes
           / CUDA-event duration measured live by idfx_timing, against MEASURED_PEAKS.json.
`cpu_baseline`: the reference's own Kokkos-OpenMP build (oracle/_ref, kind "reference") timed on
           this box's host cores on a bounded sample (a smaller block, a few cycles).
`--impl reference` runs only that CPU arm and prints it in the same JSON shape.
"""
import argparse
import ctypes as C
import json
import os
import re
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DECOMP = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}
METRIC = "cell-updates/s (FP64, 3D MHD HLLD+PLM+CT Orszag-Tang, RK2)"


def ot3d_ini(n, emf, nstages=2, log=1000000):
    emf_line = f"emf       {emf}\n" if emf else ""
    return f"""[Grid]
X1-grid    1  0.0  {n[0]}  u  1.0
X2-grid    1  0.0  {n[1]}  u  1.0
X3-grid    1  0.0  {n[2]}  u  1.0
[TimeIntegrator]
CFL         0.9
tstop       1000.0
first_dt    1.e-4
nstages     {nstages}
[Hydro]
solver    hlld
{emf_line}[Boundary]
X1-beg    periodic
X1-end    periodic
X2-beg    periodic
X2-end    periodic
X3-beg    periodic
X3-end    periodic
[Output]
log    {log}
"""


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation on the host cores (oracle/_ref)
# ------------------------------------------------------------------------------------------------
def run_cpu_reference(block, emf, steps, warmup):
    """-> (cell-updates/s over the last `steps` cycles, threads, sample description)"""
    from oracle.ref_run import cleanup, have, run_reference
    if not have("ot3d_plm"):
        raise RuntimeError("oracle/_ref/ot3d_plm/idefix is missing (build it with oracle/ref_build/build_all.sh)")
    n = (block, block, block)
    ini = ot3d_ini(n, emf, log=1)
    res = run_reference("ot3d_plm", ini, maxcycles=warmup + steps + 1, nowrite=True, timeout=3000)
    rates = []
    for line in res["log"].splitlines():
        m = re.match(r"TimeIntegrator:\s+\S+\s+\|\s+(\d+)\s+\|\s+\S+\s+\|\s+([0-9.eE+-]+)", line)
        if m and int(m.group(1)) > warmup:
            rates.append(float(m.group(2)))
    cleanup(res)
    rates = rates[:steps]
    if not rates:
        raise RuntimeError("could not parse the reference log:\n" + res["log"][-2000:])
    value = len(rates) / sum(1.0 / r for r in rates)   # cells*cycles / total time
    sample = (f"{block}^3 OT3D HLLD+PLM+{emf or 'uct_contact'} RK2, {len(rates)} cycles after {warmup} warm-up, "
              f"reference Kokkos-OpenMP build, {res['threads']} threads")
    return value, res["threads"], sample


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    block = args.cpu_block
    t0 = time.time()
    value, threads, sample = run_cpu_reference(block, args.emf, args.steps, args.warmup)
    cells = block ** 3
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "cell-updates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * cells / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, block_override=block),
        "cpu_baseline": {"value": value, "unit": "cell-updates/s", "cores": threads, "kind": "reference",
                         "sample": sample},
        "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.time() - t0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, block_override=None):
    b = block_override or args.block
    nproc = DECOMP[args.gpus]
    return {"workload": "test/MHD/OrszagTang3D (3-D Orszag-Tang), HLLD + PLM + CT(%s), RK2, periodic" %
                        (args.emf or "uct_contact"),
            "block_per_gpu": [b, b, b], "global_grid": [b * nproc[0], b * nproc[1], b * nproc[2]],
            "decomposition": list(nproc), "halo": "none" if args.gpus == 1 else args.halo,
            "l2": "inputs (>= 50 GB of state arrays at 512^3) far larger than the 126 MB L2; no flush needed"}


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self._stop, self._th = index, [], threading.Event(), None

    def _loop(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                self.samples.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._th = threading.Thread(target=self._loop, daemon=True)
        self._th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._th.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx.append(float(s[1]))
                for name, flag in zip(names, s[2:6]):
                    if flag.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def sweep_algorithmic_bytes(nvar, ndim, emf, stage0_share, direction):
    """DESIGN.md 'algorithmic bytes': per interior cell and per sweep launch (average over the RK stages).
    Every sweep reads Vc (nvar) + the normal Vs (1) and writes the face EMF outputs (2 EMFs + contact
    weight, or 2 + 4 UCT coefficients).  The x1 sweep rebuilds Uc from Vc: it WRITES the nvar-ndim non-CT
    conservative variables (and their copy Uc0 on stage 0), InvDt (stage 0) and, for uct_contact, the
    ndim cell-centred EMFs; the x2/x3 sweeps read-modify-write Uc and InvDt (stage 0)."""
    face = {"uct_contact": 3, None: 3, "uct0": 2, "arithmetic": 2, "uct_hll": 6, "uct_hlld": 6}[emf]
    nu = nvar - ndim
    if direction == 0:
        centre = ndim if emf in (None, "uct_contact") else 0   # cell-centred EMFs written for the contact average
        doubles = nvar + 1 + nu + nu * stage0_share + face + centre + 1 * stage0_share
    else:
        doubles = nvar + 1 + 2 * nu + face + 2 * stage0_share
    return 8.0 * doubles


def main_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from idefix_b200 import _abi
    from idefix_b200.setups import OrszagTang3D

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this framework has no CPU fallback "
                         "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nproc = DECOMP[world]
    xproc = (rank % nproc[0], (rank // nproc[0]) % nproc[1], rank // (nproc[0] * nproc[1]))
    b = args.block
    n = (b * nproc[0], b * nproc[1], b * nproc[2])
    lib = _abi.lib()

    setup = OrszagTang3D(order=2, ini=ot3d_ini(n, args.emf), nproc=nproc, xproc=xproc, device=local)
    data = setup.data
    h = data._handle
    stream = torch.cuda.current_stream()
    _abi.check(lib.idfx_set_stream(h, C.c_void_p(stream.cuda_stream)))
    exchanger = None
    if world > 1:
        from idefix_b200.halo import HaloExchanger
        exchanger = HaloExchanger(data, nproc, xproc, mode=args.halo)
    setup.initialise()
    l0 = data.launch_count()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up --------------------------------------------------------------------------------
    _abi.check(lib.idfx_cycles(h, args.warmup))
    barrier()

    # ---- timed region: K device-resident cycles ----------------------------------------------------
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _abi.check(lib.idfx_timing(h, 1, None))
    l1 = data.launch_count()
    with ClockSampler(local) as clocks:
        barrier()
        ev0.record(stream)
        _abi.check(lib.idfx_cycles(h, args.steps))
        ev1.record(stream)
        barrier()
    ms = ev0.elapsed_time(ev1)
    launches = data.launch_count() - l1
    tim = (C.c_double * 12)()
    _abi.check(lib.idfx_timing(h, 2, tim))
    _abi.check(lib.idfx_timing(h, 0, None))
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    cells_global = n[0] * n[1] * n[2]
    value = cells_global * args.steps / (ms * 1e-3)

    # sanity: the run must have produced finite data and a solenoidal field
    nan = data.CheckNan()
    divb = data.hydro.CheckDivB()
    mathexc = data.math_exceptions()
    if nan or not divb < 1e-10 or mathexc:
        raise SystemExit(f"bench.py: invalid state after the timed region (nan={nan}, divB={divb}, "
                         f"fast-math range-check failures={mathexc})")

    # ---- roofline of the dominant kernel -------------------------------------------------------------
    classes = ["sweep_x1", "sweep_x2", "sweep_x3", "constrained_transport", "stage_epilogue", "boundaries"]
    kern = {c: {"launch_brackets": int(tim[2 * i]), "ms_total": tim[2 * i + 1]} for i, c in enumerate(classes)}
    sweeps = [(tim[2 * i + 1] / max(tim[2 * i], 1), i) for i in range(3)]
    dur_ms, top = max(sweeps)
    nstages = 2
    bytes_cell = sweep_algorithmic_bytes(8, 3, args.emf, 1.0 / nstages, top)
    alg_bytes = bytes_cell * b ** 3
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (dur_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "sweep_dram_bytes.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as fh:
                tj = json.load(fh)
            if tj.get("block") == b and tj.get("emf") == (args.emf or "uct_contact"):
                traffic = tj["bytes_per_launch"][classes[top]]
        except Exception:
            traffic = None
    kname = ["pencilRowKernel<Phys<3,3,mhd>,HLLD,PLM,CART> (x1 sweep)",
             "pencilMarchKernel<Phys<3,3,mhd>,HLLD,PLM,JDIR,CART> (x2 sweep)",
             "pencilMarchKernel<Phys<3,3,mhd>,HLLD,PLM,KDIR,CART> (x3 sweep)"][top]
    roofline = {"bound": "hbm", "kernel": kname,
                "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "algorithmic_bytes_per_launch": alg_bytes,
                "algorithmic_bytes_per_cell": bytes_cell, "avg_launch_ms": dur_ms,
                "cycle_alg_bytes_per_cell_update": 320,
                "cycle_frac_of_peak": value / world * 320 / (peak * 1e9),
                "kernel_ms_total": kern, "timed_region_ms": ms}

    # ---- e2e: host buffers in, host buffers out, every step ----------------------------------------------
    Vc_dims, _ = data.field_dims(_abi.FIELD_VC)
    Vs_dims, _ = data.field_dims(_abi.FIELD_VS)
    hVc = torch.empty(Vc_dims, dtype=torch.float64, pin_memory=True)
    hVs = torch.empty(Vs_dims, dtype=torch.float64, pin_memory=True)
    _abi.check(lib.idfx_download(h, _abi.FIELD_VC, C.c_void_p(hVc.data_ptr())))
    _abi.check(lib.idfx_download(h, _abi.FIELD_VS, C.c_void_p(hVs.data_ptr())))
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    tt, dd, ff = C.c_double(), C.c_double(), C.c_int()
    barrier()
    t0 = time.perf_counter()
    ev0.record(stream)
    for _ in range(e2e_steps):
        # the step's inputs: the state a set-up / restart hands over, from pinned host memory
        _abi.check(lib.idfx_upload(h, _abi.FIELD_VC, C.c_void_p(hVc.data_ptr())))
        _abi.check(lib.idfx_upload(h, _abi.FIELD_VS, C.c_void_p(hVs.data_ptr())))
        # one cycle through the host-synchronous entry point; its result (t, dt, status flags -- what the
        # reference's main loop looks at after TimeIntegrator::Cycle) is read back from the device
        _abi.check(lib.idfx_cycle(h, C.byref(tt), C.byref(dd), C.byref(ff)))
        if args.e2e_full_state:
            _abi.check(lib.idfx_download(h, _abi.FIELD_VC, C.c_void_p(hVc.data_ptr())))
            _abi.check(lib.idfx_download(h, _abi.FIELD_VS, C.c_void_p(hVs.data_ptr())))
    ev1.record(stream)
    barrier()
    e2e_ms = max(ev0.elapsed_time(ev1), (time.perf_counter() - t0) * 1e3)
    if world > 1:
        t = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    state_bytes = (hVc.numel() + hVs.numel()) * 8
    e2e = {"value": cells_global * e2e_steps / (e2e_ms * 1e-3), "unit": "cell-updates/s",
           "h2d_bytes_per_step": state_bytes, "d2h_bytes_per_step": state_bytes if args.e2e_full_state else 20,
           "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
           "api": "DataBlockHost.SyncToDevice (idfx_upload Vc,Vs from pinned host memory) -> idfx_cycle -> (t, dt, status)"
                  + (" + SyncFromDevice (idfx_download Vc,Vs)" if args.e2e_full_state else "")}

    # ---- CPU baseline beside it (rank 0, N = 1 only) --------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            v, threads, sample = run_cpu_reference(args.cpu_block, args.emf, args.cpu_steps, 1)
            cpu = {"value": v, "unit": "cell-updates/s", "cores": threads, "kind": "reference", "sample": sample}
        except Exception as e:  # the CPU arm is a reported baseline; its absence must be visible, not fatal
            cpu = {"value": None, "unit": "cell-updates/s", "cores": 0, "kind": "reference",
                   "sample": f"unavailable: {e}"[:300]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "cell-updates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args), "clocks": clocks.summary(), "e2e": e2e,
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "per_gpu": value / world, "divB": divb,
            "published_h100_per_gpu": 3.08e8,
        }
        print(json.dumps(line), flush=True)
    if exchanger is not None:
        exchanger.close()
    data.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--block", type=int, default=512, help="interior cells per direction per GPU")
    ap.add_argument("--emf", default=None, help="EMF averaging ([Hydro] emf); default = the reference's default uct_contact")
    ap.add_argument("--halo", default="nccl", choices=["nccl", "p2p"])
    ap.add_argument("--e2e-steps", type=int, default=1000000, help="cap on the e2e steps (default: K)")
    ap.add_argument("--e2e-full-state", action="store_true",
                    help="e2e also reads the whole evolved state back every step (default: the cycle's result t, dt, status)")
    ap.add_argument("--cpu-block", type=int, default=128, help="block of the bounded CPU sample")
    ap.add_argument("--cpu-steps", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.gpus not in DECOMP:
        raise SystemExit("--gpus must be 1, 2, 4 or 8")
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        return main_reference(args)
    return main_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
