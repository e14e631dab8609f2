// timeIntegrator.hpp -- TimeIntegrator (reference: src/timeIntegrator.hpp, timeIntegrator.cpp:23-85
// constructor, :246-414 Cycle, :87-243 log line).
//
// Cycle() has two bit-identical implementations:
//  * the reference's own step sequence through the granular entry points, used whenever a user
//    hook (step functions, user-defined boundaries) must run on the host between the steps;
//  * idfx_cycle / idfx_cycles: whole cycles enqueued on the device, dt control included.
#include "idefix.hpp"
#ifndef IDEFIX_B200_HOST_TIMEINTEGRATOR_HPP_
#define IDEFIX_B200_HOST_TIMEINTEGRATOR_HPP_

class TimeIntegrator {
 public:
  TimeIntegrator(Input &input, DataBlock &data) {
    nstages = input.Get<int>("TimeIntegrator", "nstages", 0);
    haveFixedDt = input.CheckEntry("TimeIntegrator", "fixed_dt") > 0;
    if (haveFixedDt) fixedDt = input.Get<real>("TimeIntegrator", "fixed_dt", 0);
    cfl = data.cfl;
    cflMaxVar = data.cflMaxVar;
    cyclePeriod = input.GetOrSet<int>("Output", "log", 0, 100);
    checkNanPeriodicity = input.GetOrSet<int>("TimeIntegrator", "check_nan", 0, 100);
    maxdivB = input.GetOrSet<real>("TimeIntegrator", "maxdivB", 0, 1e-6);
    if (nstages == 2) {
      wc[0] = 0.5;
      w0[0] = 0.5;
    }
    if (nstages == 3) {
      wc[0] = 0.25;
      w0[0] = 0.75;
      wc[1] = 2.0 / 3.0;
      w0[1] = 1.0 / 3.0;
    }
    if (nstages < 1 || nstages > 3) IDEFIX_ERROR("The time integrator nstages should be 1, 2 or 3");
    IDFX_ABI(idfx_set_tstop(data.handle, input.Get<real>("TimeIntegrator", "tstop", 0)));
    data.PushTime();
    start = std::chrono::steady_clock::now();
    lastLog = start;
  }

  long long GetNCycles() const { return ncycles; }
  long long ncycles{0};

  void ShowConfig() {
    const char *names[] = {"", "Euler", "RK2", "RK3"};
    idfx::cout << "TimeIntegrator: using " << names[nstages] << " integrator." << std::endl;
    if (haveFixedDt) idfx::cout << "TimeIntegrator: Using fixed dt time stepping. Ignoring CFL and first_dt." << std::endl;
    else idfx::cout << "TimeIntegrator: Using adaptive dt time stepping with CFL=" << cfl << "." << std::endl;
  }

  bool CheckForMaxRuntime() { return false; }

  // cycles that can be enqueued without the host looking at the state in between
  void SetDeviceBatch(int n) { deviceBatch = n < 1 ? 1 : n; }

  void Cycle(DataBlock &data) {
    if (ncycles % cyclePeriod == 0) ShowLog(data);
    const bool hooks = data.haveUserStepFirst || data.haveUserStepLast || data.hydro->userDefBoundaryFunc ||
                       data.hydro->userDefBoundaryFuncOld || data.hydro->haveInternalBoundary ||
                       data.hydro->haveEmfBoundary;
    if (hooks) {
      CycleReferenceOrder(data);
      return;
    }
    // fused path.  The device clips the last step at tstop itself (main.cpp:191); a batch never crosses
    // a log line, and the caller bounds it so that no periodic output is skipped.
    long long batch = std::min<long long>(deviceBatch, cyclePeriod - (ncycles % cyclePeriod));
    int flags = 0;
    if (batch > 1) {
      IDFX_ABI(idfx_cycles(data.handle, static_cast<int>(batch - 1)));
      ncycles += batch - 1;
    }
    IDFX_ABI(idfx_cycle(data.handle, &data.t, &data.dt, &flags));
    ncycles++;
    long long done = 0;
    IDFX_ABI(idfx_get_ncycles(data.handle, &done));
    if (flags & 2)
      IDEFIX_ERROR("Your guessed first_dt is too large.\nTry to reduce first_dt in the ini file.");
    if (flags & 1) {
      std::stringstream msg;
      msg << "dt = " << data.dt << " is too small.";
      throw std::runtime_error(msg.str());
    }
    if (flags & 4)
      throw std::runtime_error(
          "the sweeps met operands outside the validated range of their branch-free FP64 divisions / square roots "
          "(zero-crossing, subnormal, beyond 2^+-140, infinite or NaN: idfx_math_exceptions > 0); the state is no "
          "longer guaranteed to be bit-identical with the reference.  Rebuild libidefix_b200.so with FASTDIV=0 for "
          "this regime.");
    if (ncycles % checkNanPeriodicity == 0 && data.CheckNan() > 0)
      throw std::runtime_error("Nan found after integration cycle");
  }

 private:
  int nstages{1};
  bool haveFixedDt{false};
  real fixedDt{0}, cfl{0}, cflMaxVar{1.1}, maxdivB{1e-6};
  int cyclePeriod{100}, checkNanPeriodicity{100};
  int deviceBatch{1};
  real wc[2] = {0, 0}, w0[2] = {0, 0};
  std::chrono::steady_clock::time_point start, lastLog;
  long long lastLogCycle{0};

  // timeIntegrator.cpp:246-414 step by step
  void CycleReferenceOrder(DataBlock &data) {
    data.PushTime();
    const real t0 = data.t;
    real newdt = 0.0;
    data.LaunchUserStepFirst();
    data.ResetStage();
    for (int stage = 0; stage < nstages; stage++) {
      data.SetBoundaries();
      data.PrimToCons();
      if (nstages > 1 && stage == 0) IDFX_ABI(idfx_state_copy(data.handle));
      data.EvolveStage();
      data.t += data.dt;
      if (ncycles % checkNanPeriodicity == 0 && data.CheckNan() > 0)
        throw std::runtime_error("Nan found after integration cycle");
      if (stage == 0 && !haveFixedDt) newdt = cfl * data.ComputeTimestep();
      if (stage > 0) {
        const real wcs = wc[stage - 1], w0s = w0[stage - 1];
        IDFX_ABI(idfx_state_add_and_store(data.handle, wcs, w0s));
        data.t = wcs * data.t + w0s * t0;
      }
      data.PushTime();
      data.ConsToPrim();
    }
    data.LaunchUserStepLast();
    data.t = t0 + data.dt;
    if (!haveFixedDt) {
      if (newdt > cflMaxVar * data.dt) {
        data.dt = cflMaxVar * data.dt;
      } else {
        if (ncycles == 0 && newdt < 0.5 * data.dt) {
          std::stringstream msg;
          msg << "Your guessed first_dt is too large. My next dt=" << newdt << std::endl
              << "Try to reduce first_dt in the ini file.";
          IDEFIX_ERROR(msg);
        }
        data.dt = newdt;
      }
      if (data.dt < 1e-15) {
        std::stringstream msg;
        msg << "dt = " << data.dt << " is too small.";
        throw std::runtime_error(msg.str());
      }
    } else {
      data.dt = fixedDt;
    }
    data.PushTime();
    ncycles++;
  }

  // timeIntegrator.cpp:87-243: "TimeIntegrator: <t> | <cycle> | <dt> | <cell updates/s> | <div B>"
  void ShowLog(DataBlock &data) {
    const auto now = std::chrono::steady_clock::now();
    if (ncycles == 0) {
      idfx::cout << "TimeIntegrator: " << std::setw(10) << "time" << " | " << std::setw(10) << "cycle" << " | "
                 << std::setw(10) << "time step" << " | " << std::setw(10) << "cell (updates/s)";
#if MHD == YES
      idfx::cout << " | " << std::setw(10) << "div B";
#endif
      idfx::cout << std::endl;
    }
    const double secs = std::chrono::duration<double>(now - lastLog).count();
    const double cells = static_cast<double>(data.mygrid->np_int[0]) * data.mygrid->np_int[1] * data.mygrid->np_int[2];
    idfx::cout << "TimeIntegrator: " << std::scientific << std::setw(10) << data.t << " | " << std::setw(10) << ncycles
               << " | " << std::setw(10) << data.dt << " | " << std::setw(10);
    if (ncycles > lastLogCycle && secs > 0) idfx::cout << cells * (ncycles - lastLogCycle) / secs;
    else idfx::cout << "N/A";
#if MHD == YES
    const real divB = data.hydro->CheckDivB();
    idfx::cout << " | " << std::setw(10) << divB;
    if (divB > maxdivB) {
      idfx::cout << std::endl;
      std::stringstream msg;
      msg << "TimeIntegrator::Cycle divB>maxdivB check failed. divB=" << divB;
      throw std::runtime_error(msg.str());
    }
#endif
    idfx::cout << std::endl;
    lastLog = std::chrono::steady_clock::now();
    lastLogCycle = ncycles;
  }
};

#endif  // IDEFIX_B200_HOST_TIMEINTEGRATOR_HPP_
