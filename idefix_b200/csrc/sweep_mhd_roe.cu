// Roe sweeps, adiabatic EOS (roeMHD.hpp)
#include "sweep_inst.cuh"
int idfxk::sweepMhdRoe(const PhysCfg &c, int dir, const SweepArgs &a, cudaStream_t st) {
  IDFX_SWEEP_MHD_ALL(SOLVER_ROE, false)
  return sweepMhdRoeIso(c, dir, a, st);
}
