// common.cuh -- device-side descriptors shared by all kernels.
//
// The reference specialises its kernels at compile time through definitions.hpp
// macros (src/idefix.hpp:47-108).  Here the same information is carried by the
// Phys<> template (phys.cuh) for what changes the variable list, and by the plain
// structs below for everything else.
#ifndef IDFX_COMMON_CUH_
#define IDFX_COMMON_CUH_

#include <cuda_runtime.h>
#include <cstdint>

#define IDFX_HD __host__ __device__ __forceinline__
#ifdef IDFX_HOST_TEST
// tests/host_cpp/riemann_host.cu compiles the per-interface physics for the HOST as well, to compare it with the
// CPU oracle without a GPU (the *Fast sequences of fastmath.cuh stay device-only and are not reached there)
#define IDFX_D __host__ __device__ __forceinline__
#else
#define IDFX_D __device__ __forceinline__
#endif

enum : int { IDIR = 0, JDIR = 1, KDIR = 2 };
enum : int { GEO_CARTESIAN = 1, GEO_CYLINDRICAL = 2, GEO_POLAR = 3, GEO_SPHERICAL = 4 };
enum : int { SOLVER_TVDLF = 0, SOLVER_HLL = 1, SOLVER_HLLC = 2, SOLVER_HLLD = 3, SOLVER_ROE = 4 };
enum : int { EMF_ARITHMETIC = 0, EMF_UCT0 = 1, EMF_UCT_CONTACT = 2, EMF_UCT_HLL = 3, EMF_UCT_HLLD = 4 };
enum : int { BC_INTERNAL = 0, BC_PERIODIC = 1, BC_REFLECTIVE = 2, BC_OUTFLOW = 3,
             BC_SHEARINGBOX = 4, BC_AXIS = 5, BC_USERDEF = 6, BC_UNDEFINED = 7 };

// Geometry of one block as seen by the kernels.  All pointers are device pointers.
struct DevGrid {
  int ndim, geometry, regular;
  int nx[3];              // np_tot
  int beg[3], end[3];     // interior range [beg,end)
  int ng[3];              // nghost
  int off[3];             // IOFFSET/JOFFSET/KOFFSET (1 for active dims)
  // cell-centred 3-D arrays: index = (k*nx[1] + j)*nx[0] + i
  long long sj, sk, sn;   // strides: j, k, variable
  // staggered array Vs(n,k,j,i): dims (ndim, nx2+off2, nx1+off1, nx0+off0)
  long long vsj, vsk, vsn;
  const double *x[3], *xl[3], *xr[3], *dx[3], *xgc[3];
  // spherical helpers (makeGeometry.cpp:93-121), valid only when geometry==SPHERICAL
  const double *rt, *sinx2m, *tanx2m, *sinx2, *tanx2, *dmu;
  // separable factors: dV(k,j,i) = (vol[0][i]*vol[1][j])*vol[2][k]   (makeGeometry.cpp:66-84)
  const double *vol[3];
  // A[d](k,j,i) = (area[d][0][i]*area[d][1][j])*area[d][2][k]      (makeGeometry.cpp:139-211)
  const double *area[3][3];
  // irregular-grid PLM weights (extrapolateToFaces.hpp:43-71), per direction
  const double *wp[3], *wm[3], *cp[3], *cm[3], *dp[3], *dm[3];
};

IDFX_HD long long cidx(const DevGrid &g, int k, int j, int i) {
  return ((long long)k * g.nx[1] + j) * g.nx[0] + i;
}
IDFX_HD long long sidx(const DevGrid &g, int k, int j, int i) {
  return (long long)k * g.vsk + (long long)j * g.vsj + i;
}
IDFX_HD double cellVolume(const DevGrid &g, int k, int j, int i) {
  return g.vol[0][i] * g.vol[1][j] * g.vol[2][k];
}
template <int D>
IDFX_HD double faceArea(const DevGrid &g, int k, int j, int i) {
  return g.area[D][0][i] * g.area[D][1][j] * g.area[D][2][k];
}
IDFX_HD double faceAreaDir(const DevGrid &g, int d, int k, int j, int i) {
  return g.area[d][0][i] * g.area[d][1][j] * g.area[d][2][k];
}

// Constrained-transport work arrays (constrainedTransport.hpp:22-144)
struct DevEmf {
  double *ex, *ey, *ez;                       // corner (edge) EMFs
  double *exj, *exk, *eyi, *eyk, *ezi, *ezj;  // face-centred EMFs
  double *Ex1, *Ex2, *Ex3;                    // cell-centred EMFs (uct0 / uct_contact)
  double *svx, *svy, *svz;                    // contact upwinding weights
  double *axL, *axR, *ayL, *ayR, *azL, *azR;  // UCT-HLL/HLLD a coefficients
  double *dxL, *dxR, *dyL, *dyR, *dzL, *dzR;  // UCT-HLL/HLLD d coefficients
};

// Everything a sweep needs
struct SweepArgs {
  DevGrid g;
  DevEmf emf;
  const double *Vc;   // (nvar,k,j,i)
  double *Uc;         // (nvar,k,j,i)
  const double *Vs;   // (ndim,k,j,i) staggered
  double *invDt;      // (k,j,i)
  const double *dt;   // device scalar
  double gamma, csiso;
  int emfMode;
  int perpExt;        // perpendicular extension of the face box (hlldMHD.hpp:32-49)
  int doInvDt;        // accumulate InvDt (needed only when the time step is read back)
  int doSource;       // add geometrical source terms after this sweep (last sweep only)
  int fusedP2C;       // Uc is rebuilt from Vc in this sweep instead of read (first sweep of the fused path)
  int invDtInit;      // InvDt accumulator starts from zero in this sweep (first sweep after ResetStage)
  int validateInputs;       // x1 sweep: check that the primitives lie in the envelope of fastmath.cuh
  unsigned int *mathFlag;   // incremented by threads whose branch-free division / sqrt failed a range check (fastmath.cuh)
  double *fluxRho;    // when tracers are on: raw Riemann mass flux of every interface of this sweep (else null)
  double *Uc0;        // fused path, first stage: also save the rebuilt Uc as the "begin" state (may be null)
  // TMA (tma.cuh): HOST pointer to the block's CUtensorMap descriptor of Vc for the x1 sweep (boxes of 32 + 2H cells
  // x all variables), null when Vc does not qualify; the launcher passes the descriptor by value and sets useTma.
  // tmapNvar = number of variables a box carries.
  const void *tmapRow;
  int tmapNvar, tmapRowCells;
  int useTma;
  const unsigned char *shockFlags;   // shock flattening: 1 where the cell is flagged (shockFlattening.hpp), null = off
  double *staleFlux;  // 1-D HD Roe only: (nvar, nx0) area-weighted flux of the previous call (roeHD.hpp:312-328), else null
};

#endif  // IDFX_COMMON_CUH_
