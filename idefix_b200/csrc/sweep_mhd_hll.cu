// HLL sweeps (hllMHD.hpp)
#include "sweep_inst.cuh"
int idfxk::sweepMhdHll(const PhysCfg &c, int dir, const SweepArgs &a, cudaStream_t st) {
  IDFX_SWEEP_MHD_ALL(SOLVER_HLL, false)
  IDFX_SWEEP_MHD_ALL(SOLVER_HLL, true)
  return -1;
}
