// TVDLF (Rusanov) sweeps (tvdlfMHD.hpp)
#include "sweep_inst.cuh"
int idfxk::sweepMhdTvdlf(const PhysCfg &c, int dir, const SweepArgs &a, cudaStream_t st) {
  IDFX_SWEEP_MHD_ALL(SOLVER_TVDLF, false)
  IDFX_SWEEP_MHD_ALL(SOLVER_TVDLF, true)
  return -1;
}
