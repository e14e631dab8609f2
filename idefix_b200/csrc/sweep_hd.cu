// Hydrodynamic sweeps: TVDLF, HLL, HLLC (HDsolvers/*.hpp)
#include "sweep_inst.cuh"
int idfxk::sweepHD(const PhysCfg &c, int dir, const SweepArgs &a, cudaStream_t st) {
  if (c.solver == SOLVER_TVDLF) {
    IDFX_SWEEP_HD_ALL(SOLVER_TVDLF, false)
    IDFX_SWEEP_HD_ALL(SOLVER_TVDLF, true)
  } else if (c.solver == SOLVER_HLL) {
    IDFX_SWEEP_HD_ALL(SOLVER_HLL, false)
    IDFX_SWEEP_HD_ALL(SOLVER_HLL, true)
  } else if (c.solver == SOLVER_HLLC) {
    IDFX_SWEEP_HD_ALL(SOLVER_HLLC, false)
    IDFX_SWEEP_HD_ALL(SOLVER_HLLC, true)
  }
  return -1;
}
