// main.cpp -- the driver of a problem built against the B200 host layer: same flow, command line
// and log lines as the reference's src/main.cpp:60-300 (Input -> Grid -> DataBlock -> Setup ->
// Output -> InitFlow or restart -> time loop -> "Main: Perfs are ... cell updates/second").
#include "idefix.hpp"
#include "setup.hpp"

__attribute__((weak)) Setup::~Setup() {}

int main(int argc, char **argv) {
  int returnCode = 0;
  {
    int ndev = 0;
    if (idfx_device_count(&ndev) != 0 || ndev == 0)
      IDEFIX_ERROR("no CUDA device: this build has no CPU fallback");
    Input input(argc, argv);
    input.ShowConfig();
    Grid grid(input);
    DataBlock data(grid, input);
    TimeIntegrator Tint(input, data);
    Output output(input, data);
    Setup mysetup(input, grid, data, output);
    grid.ShowConfig();
    data.ShowConfig();
    Tint.ShowConfig();

    if (input.restartRequested) {
      if (input.forceInitRequested) {
        mysetup.InitFlow(data);
        data.DeriveVectorPotential();
      }
      idfx::cout << "Main: Restarting from dump file." << std::endl;
      if (!output.RestartFromDump(data, input.restartFileNumber)) {
        idfx::cout << "Main: restart aborted." << std::endl;
        input.restartRequested = false;
      } else {
        data.SetBoundaries();
      }
    }
    if (!input.restartRequested) {
      idfx::cout << "Main: Creating initial conditions." << std::endl;
      mysetup.InitFlow(data);
      data.DeriveVectorPotential();   // does something only when the vector potential is evolved (main.cpp:174)
      data.SetBoundaries();
      data.Validate();
      output.CheckForWrites(data);
    }

    idfx::cout << "Main: Cycling Time Integrator..." << std::endl;
    IDFX_CUDA(cudaDeviceSynchronize());
    const auto t0 = std::chrono::steady_clock::now();
    output.ResetTimer();
    const real tstop = input.Get<real>("TimeIntegrator", "tstop", 0);
    // without periodic outputs nothing on the host needs the state between log lines
    const bool batched = !output.HasPeriodicOutputs();

    while (data.t < tstop) {
      if (tstop - data.t < data.dt) data.dt = tstop - data.t;
      if (batched) {
        // cycles left before tstop / maxcycles: dt grows by at most CFL_max_var per cycle and the device
        // clips the last step itself, so a batch never overshoots
        long long room = 1 << 20;
        if (input.maxCycles >= 0) room = input.maxCycles - Tint.GetNCycles();
        if (data.dt > 0 && data.cflMaxVar > 1.0 && data.fixedDt <= 0) {
          const double ratio = (tstop - data.t) / data.dt, g = data.cflMaxVar;
          room = std::min<long long>(room, static_cast<long long>(std::log1p((g - 1.0) * ratio) / std::log(g)));
        } else if (data.dt > 0) {
          room = std::min<long long>(room, static_cast<long long>((tstop - data.t) / data.dt));
        }
        Tint.SetDeviceBatch(static_cast<int>(std::max<long long>(1, std::min<long long>(room, 64))));
      }
      try {
        Tint.Cycle(data);
      } catch (std::exception &e) {
        idfx::cout << "Main: WARNING! Caught an exception in TimeIntegrator." << std::endl;
        idfx::cout << e.what() << std::endl;
        idfx::cout << "Main: attempting to save the current state for inspection." << std::endl;
        output.ForceWriteVtk(data);
        idfx::cout << "Main: Aborting current calculation." << std::endl;
        returnCode = 1;
        break;
      }
      output.CheckForWrites(data);
      if (input.maxCycles >= 0 && Tint.GetNCycles() >= input.maxCycles) {
        idfx::cout << "Main: Reached maximum number of integration cycles." << std::endl;
        break;
      }
    }
    IDFX_CUDA(cudaDeviceSynchronize());
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const double perfs = secs / grid.np_int[IDIR] / grid.np_int[JDIR] / grid.np_int[KDIR] / Tint.GetNCycles() * idfx::psize;
    idfx::cout << "Main: Reached t=" << std::setprecision(17) << data.t << std::setprecision(6) << std::endl;
    idfx::cout << "Main: Completed in " << static_cast<long long>(secs) << " seconds and " << Tint.GetNCycles()
               << " cycle" << (Tint.GetNCycles() != 1 ? "s" : "") << std::endl;
    idfx::cout << "Main: Perfs are " << std::scientific << 1 / perfs << " cell updates/second" << std::endl;
    unsigned int mathExceptions = 0;
    IDFX_ABI(idfx_math_exceptions(data.handle, &mathExceptions));
    if (mathExceptions > 0)
      IDEFIX_WARNING("the branch-free FP64 division / square root of the sweeps met " + std::to_string(mathExceptions) +
                     " out-of-range operands (zero, subnormal, infinite or NaN); rebuild libidefix_b200.so with "
                     "FASTDIV=0 for strictly IEEE results in that regime");
    idfx::cout << "Outputs represent " << static_cast<int>(100.0 * output.GetTimer() / secs) << "% of total run time."
               << std::endl;
  }
  if (returnCode > 0) idfx::cout << "Main: Job's done (with errors)" << std::endl;
  else idfx::cout << "Main: Job's done" << std::endl;
  return returnCode;
}
