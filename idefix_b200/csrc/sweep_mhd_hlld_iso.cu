// HLLD sweeps, isothermal EOS (hlldMHD.hpp:459-572)
#include "sweep_inst.cuh"
int idfxk::sweepMhdHlldIso(const PhysCfg &c, int dir, const SweepArgs &a, cudaStream_t st) {
  IDFX_SWEEP_MHD_ALL(SOLVER_HLLD, true)
  return -1;
}
