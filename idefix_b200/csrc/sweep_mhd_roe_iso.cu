// Roe sweeps, isothermal EOS (roeMHD.hpp)
#include "sweep_inst.cuh"
int idfxk::sweepMhdRoeIso(const PhysCfg &c, int dir, const SweepArgs &a, cudaStream_t st) {
  IDFX_SWEEP_MHD_ALL(SOLVER_ROE, true)
  return -1;
}
