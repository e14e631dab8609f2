"""TimeIntegrator mirror (src/timeIntegrator.cpp:23-85 ctor, :246-414 Cycle) and the
main time loop (src/main.cpp:188-225).

Two ways to advance, giving bit-identical fields:
  * ``Cycle(data)``            the reference's own call sequence, one ABI call per reference
                               call site, dt logic on the host exactly as in Cycle();
  * ``data``-side fused path   ``idfx_cycles`` runs whole cycles on the device with the dt
                               control logic in a kernel (no host round trip per cycle).
"""
import ctypes as C
import math

from ._abi import IdefixError, check, lib


class TimeIntegrator:
    def __init__(self, inp, data):
        self.nstages = inp.Get(int, "TimeIntegrator", "nstages", 0)
        self.haveFixedDt = inp.CheckEntry("TimeIntegrator", "fixed_dt") > 0
        self.cfl, self.cflMaxVar, self.fixedDt = data.cfl, data.cfl_max_var, data.fixed_dt
        self.cyclePeriod = inp.GetOrSet(int, "Output", "log", 0, 100)
        self.checkNanPeriodicity = inp.GetOrSet(int, "TimeIntegrator", "check_nan", 0, 100)
        self.maxdivB = inp.GetOrSet(float, "TimeIntegrator", "maxdivB", 0, 1e-6)
        self.ncycles = 0
        self.wc, self.w0 = [0.0, 0.0], [0.0, 0.0]
        if self.nstages == 2:
            self.wc[0], self.w0[0] = 0.5, 0.5
        if self.nstages == 3:
            self.wc[0], self.w0[0] = 0.25, 0.75
            self.wc[1], self.w0[1] = 2.0 / 3.0, 1.0 / 3.0

    def GetNCycles(self):
        return self.ncycles

    def Cycle(self, data):
        """one integration cycle through the granular entry points, reference order"""
        t, dt = data._time()
        if self.ncycles % self.cyclePeriod == 0 and data.defs.mhd:
            divb = data.hydro.CheckDivB()
            if divb > self.maxdivB:
                raise RuntimeError(f"TimeIntegrator::Cycle divB>maxdivB check failed. divB={divb}")
        t0 = t
        data.ResetStage()
        newdt = None
        for stage in range(self.nstages):
            data.SetBoundaries()
            data.PrimToCons()
            if self.nstages > 1 and stage == 0:
                data.state_copy()
            data.hydro.EvolveStage(t, dt)
            t += dt
            if self.ncycles % self.checkNanPeriodicity == 0:
                if data.CheckNan() > 0:
                    raise RuntimeError("Nan found after integration cycle")
            if stage == 0 and not self.haveFixedDt:
                newdt = self.cfl * data.ComputeTimestep()
            if stage > 0:
                wcs, w0s = self.wc[stage - 1], self.w0[stage - 1]
                data.state_add_and_store(wcs, w0s)
                t = wcs * t + w0s * t0
            data.set_time(t, dt)
            data.ConsToPrim()
        t = t0 + dt
        if not self.haveFixedDt:
            if newdt > self.cflMaxVar * dt:
                dt = self.cflMaxVar * dt
            else:
                if self.ncycles == 0 and newdt < 0.5 * dt:
                    raise IdefixError(f"Your guessed first_dt is too large. My next dt={newdt}\n"
                                      "Try to reduce first_dt in the ini file.")
                dt = newdt
            if dt < 1e-15:
                raise RuntimeError(f"dt = {dt} is too small.")
        else:
            dt = self.fixedDt
        data.set_time(t, dt)
        self.ncycles += 1


def run(data, tint, tstop, maxcycles=-1, fused=True, check_every=64):
    """main.cpp:188-225: cycle until tstop (last step clipped) or maxcycles.

    fused=True drives idfx_cycles in batches (the device clips the last step itself);
    fused=False calls TimeIntegrator.Cycle like the reference's main loop."""
    h = data._handle
    if not fused:
        while data.t < tstop:
            t, dt = data._time()
            if tstop - t < dt:
                data.set_time(t, tstop - t)
            tint.Cycle(data)
            if maxcycles >= 0 and tint.ncycles >= maxcycles:
                break
        return tint.ncycles
    check(lib().idfx_set_tstop(h, tstop))
    t, dt, flags = C.c_double(), C.c_double(), C.c_int()
    n = C.c_longlong()
    while True:
        check(lib().idfx_get_time(h, C.byref(t), C.byref(dt)))
        check(lib().idfx_get_ncycles(h, C.byref(n)))
        if not (t.value < tstop) or (maxcycles >= 0 and n.value >= maxcycles):
            break
        # batch size that cannot overshoot tstop: dt grows at most by CFL_max_var per cycle, so n
        # cycles advance at most dt*(g^n-1)/(g-1); the device clips the very last step itself
        batch = 1
        if dt.value > 0:
            ratio = (tstop - t.value) / dt.value
            g = tint.cflMaxVar
            if tint.haveFixedDt or g <= 1.0:
                est = int(ratio)
            else:
                est = int(math.log1p((g - 1.0) * ratio) / math.log(g))
            batch = max(1, min(check_every, est))
        if maxcycles >= 0:
            batch = min(batch, maxcycles - n.value)
        if batch > 1:
            check(lib().idfx_cycles(h, batch - 1))
        check(lib().idfx_cycle(h, C.byref(t), C.byref(dt), C.byref(flags)))
        if flags.value & 1:
            raise RuntimeError(f"dt = {dt.value} is too small.")
        if flags.value & 2:
            raise IdefixError("Your guessed first_dt is too large.\nTry to reduce first_dt in the ini file.")
        if flags.value & 4:
            raise RuntimeError("the sweeps left the validated range of their branch-free FP64 divisions / square roots "
                               "(idfx_math_exceptions > 0): results no longer guaranteed bit-identical with the "
                               "reference; rebuild libidefix_b200.so with FASTDIV=0 for this regime")
    check(lib().idfx_get_ncycles(h, C.byref(n)))
    tint.ncycles = n.value
    return n.value
