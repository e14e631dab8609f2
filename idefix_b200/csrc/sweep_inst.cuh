// sweep_inst.cuh -- run-time -> compile-time dispatch of the sweep kernels.
// The list of (DIMENSIONS, COMPONENTS) pairs is the set the reference's test suite uses.
#ifndef IDFX_SWEEP_INST_CUH_
#define IDFX_SWEEP_INST_CUH_

#include "kernels.h"
#include "sweep.cuh"
#include "sweep_pencil.cuh"

template <class P, int SOLVER, int ORDER, bool CART>
inline int sweepPencil(int dir, const SweepArgs &a, cudaStream_t st) {
  if (dir == 0) return (int)launchPencilRow<P, SOLVER, ORDER, CART>(a, st);
  if constexpr (P::ndim >= 2) {
    if (dir == 1) return (int)launchPencilMarch<P, SOLVER, ORDER, 1, CART>(a, st);
  }
  if constexpr (P::ndim == 3) {
    if (dir == 2) return (int)launchPencilMarch<P, SOLVER, ORDER, 2, CART>(a, st);
  }
  return -1;
}

template <class P, int SOLVER, int ORDER>
inline int sweepDir(int dir, const SweepArgs &a, cudaStream_t st) {
  if (a.g.geometry == GEO_CARTESIAN && a.g.regular) return sweepPencil<P, SOLVER, ORDER, true>(dir, a, st);
  return sweepPencil<P, SOLVER, ORDER, false>(dir, a, st);
}

template <class P, int SOLVER>
inline int sweepOrder(int order, int dir, const SweepArgs &a, cudaStream_t st) {
  switch (order) {
#ifndef IDFX_DEV_MINIMAL
    case 1: return sweepDir<P, SOLVER, 1>(dir, a, st);
    case 3: return sweepDir<P, SOLVER, 3>(dir, a, st);
    case 4: return sweepDir<P, SOLVER, 4>(dir, a, st);
#endif
    case 2: return sweepDir<P, SOLVER, 2>(dir, a, st);
  }
  return -1;
}

#define IDFX_SWEEP_CASE(ND, NC, MHDF, ISOF, SOLVER)                              \
  if (c.ndim == ND && c.ncomp == NC && c.iso == (ISOF ? 1 : 0))                  \
    return sweepOrder<Phys<ND, NC, MHDF, ISOF>, SOLVER>(c.order, dir, a, st);

// IDFX_DEV_MINIMAL (make DEV=1): only the 3-D PLM instantiations, for quick kernel experiments
#ifdef IDFX_DEV_MINIMAL
#define IDFX_SWEEP_MHD_ALL(SOLVER, ISOF) IDFX_SWEEP_CASE(3, 3, true, ISOF, SOLVER)
#define IDFX_SWEEP_HD_ALL(SOLVER, ISOF) IDFX_SWEEP_CASE(3, 3, false, ISOF, SOLVER)
#else
#define IDFX_SWEEP_MHD_ALL(SOLVER, ISOF)   \
  IDFX_SWEEP_CASE(1, 2, true, ISOF, SOLVER) \
  IDFX_SWEEP_CASE(1, 3, true, ISOF, SOLVER) \
  IDFX_SWEEP_CASE(2, 2, true, ISOF, SOLVER) \
  IDFX_SWEEP_CASE(2, 3, true, ISOF, SOLVER) \
  IDFX_SWEEP_CASE(3, 3, true, ISOF, SOLVER)

#define IDFX_SWEEP_HD_ALL(SOLVER, ISOF)      \
  IDFX_SWEEP_CASE(1, 1, false, ISOF, SOLVER) \
  IDFX_SWEEP_CASE(1, 2, false, ISOF, SOLVER) \
  IDFX_SWEEP_CASE(1, 3, false, ISOF, SOLVER) \
  IDFX_SWEEP_CASE(2, 2, false, ISOF, SOLVER) \
  IDFX_SWEEP_CASE(2, 3, false, ISOF, SOLVER) \
  IDFX_SWEEP_CASE(3, 3, false, ISOF, SOLVER)
#endif

#endif  // IDFX_SWEEP_INST_CUH_
