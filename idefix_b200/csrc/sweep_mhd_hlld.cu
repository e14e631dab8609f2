// HLLD sweeps, adiabatic EOS (hlldMHD.hpp)
#include "sweep_inst.cuh"
int idfxk::sweepMhdHlld(const PhysCfg &c, int dir, const SweepArgs &a, cudaStream_t st) {
  IDFX_SWEEP_MHD_ALL(SOLVER_HLLD, false)
  return sweepMhdHlldIso(c, dir, a, st);
}
