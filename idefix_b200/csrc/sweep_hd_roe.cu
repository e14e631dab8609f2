// Hydrodynamic sweeps with the Roe solver (HDsolvers/roeHD.hpp; riemann_roe_hd.cuh)
#include "sweep_inst.cuh"
int idfxk::sweepHDRoe(const PhysCfg &c, int dir, const SweepArgs &a, cudaStream_t st) {
  IDFX_SWEEP_HD_ALL(SOLVER_ROE, false)
  IDFX_SWEEP_HD_ALL(SOLVER_ROE, true)
  return -1;
}
