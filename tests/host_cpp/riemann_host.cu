// riemann_host.cu -- the product's per-interface Riemann solvers (idefix_b200/csrc/riemann_*.cuh, the very
// functions the sweep kernels inline) compiled for the HOST, behind one C entry point, so that the CPU test
// suite can compare them bit for bit with the oracle's restatement on arbitrary left / right states -- including
// the rarely taken branches (supersonic sides, HLLD degeneracies, Roe entropy fix and HLL switch) that the
// golden runs visit seldom or never.  Test harness only; built by tests/test_riemann_host.py with
//   nvcc -DIDFX_HOST_TEST -fmad=false -Xcompiler -ffp-contract=off -shared ...
#define IDFX_HOST_TEST 1
#include "riemann_hd.cuh"
#include "riemann_mhd.cuh"
#include "riemann_roe.cuh"
#include "riemann_roe_hd.cuh"
#include "emf_store.cuh"

template <class P, int SOLVER, int DIR>
static void solveAll(const Eos &eos, long n, const double *vL, const double *vR, double *F, double *aux4,
                     int emfMode, double *emf7) {
  for (long q = 0; q < n; q++) {
    RiemannAux aux{0.0, 0.0, 0.0, 0.0};
    double f[P::nvar];
    for (int v = 0; v < P::nvar; v++) f[v] = 0.0;
    const double *l = vL + q * P::nvar, *r = vR + q * P::nvar;
    if constexpr (P::mhd) {
      if constexpr (SOLVER == SOLVER_TVDLF) riemannTvdlfMHD<P, DIR>(l, r, eos, f, aux);
      if constexpr (SOLVER == SOLVER_HLL) riemannHllMHD<P, DIR>(l, r, eos, f, aux);
      if constexpr (SOLVER == SOLVER_HLLD) riemannHlldMHD<P, DIR>(l, r, eos, f, aux);
      if constexpr (SOLVER == SOLVER_ROE) riemannRoeMHD<P, DIR>(l, r, eos, f, aux);
    } else {
      if constexpr (SOLVER == SOLVER_TVDLF) riemannTvdlfHD<P, DIR>(l, r, eos, f, aux);
      if constexpr (SOLVER == SOLVER_HLL) riemannHllHD<P, DIR>(l, r, eos, f, aux);
      if constexpr (SOLVER == SOLVER_HLLC) riemannHllcHD<P, DIR>(l, r, eos, f, aux);
      if constexpr (SOLVER == SOLVER_ROE) aux.c2iso = riemannRoeHD<P, DIR>(l, r, eos, f, aux) ? 1.0 : 0.0;   // keep-flux flag
    }
    for (int v = 0; v < P::nvar; v++) F[q * P::nvar + v] = f[v];
    aux4[4 * q + 0] = aux.cmax;
    aux4[4 * q + 1] = aux.sl;
    aux4[4 * q + 2] = aux.sr;
    aux4[4 * q + 3] = aux.c2iso;
    if constexpr (P::mhd && DIR < P::ndim) {
      if (emf7 != nullptr) {   // what the interface hands to constrained transport (emf_store.cuh)
        double *e = emf7 + 7 * q;
        DevEmf dev{};
        dev.ezi = dev.ezj = dev.eyk = e + 0;   // Et of the direction
        dev.eyi = dev.exj = dev.exk = e + 1;   // Eb
        dev.svx = dev.svy = dev.svz = e + 2;
        dev.axL = dev.ayL = dev.azL = e + 3;
        dev.axR = dev.ayR = dev.azR = e + 4;
        dev.dxL = dev.dyL = dev.dzL = e + 5;
        dev.dxR = dev.dyR = dev.dzR = e + 6;
        const EmfFaceArrays A = emfFaceArrays<P::ndim, DIR>(dev);
        storeEmf<P, DIR, SOLVER>(emfMode, A, 0, f, aux, l, r);
      }
    }
  }
}

template <class P, int SOLVER>
static int byDir(int dir, const Eos &eos, long n, const double *vL, const double *vR, double *F, double *aux4,
                 int emfMode, double *emf7) {
  if (dir == 0) { solveAll<P, SOLVER, 0>(eos, n, vL, vR, F, aux4, emfMode, emf7); return 0; }
  if constexpr (P::ncomp >= 2) if (dir == 1) { solveAll<P, SOLVER, 1>(eos, n, vL, vR, F, aux4, emfMode, emf7); return 0; }
  if constexpr (P::ncomp == 3) if (dir == 2) { solveAll<P, SOLVER, 2>(eos, n, vL, vR, F, aux4, emfMode, emf7); return 0; }
  return 1;
}

template <class P>
static int bySolver(int solver, int dir, const Eos &eos, long n, const double *vL, const double *vR, double *F,
                    double *aux4, int emfMode, double *emf7) {
  if (solver == SOLVER_TVDLF) return byDir<P, SOLVER_TVDLF>(dir, eos, n, vL, vR, F, aux4, emfMode, emf7);
  if (solver == SOLVER_HLL) return byDir<P, SOLVER_HLL>(dir, eos, n, vL, vR, F, aux4, emfMode, emf7);
  if constexpr (P::mhd) {
    if (solver == SOLVER_HLLD) return byDir<P, SOLVER_HLLD>(dir, eos, n, vL, vR, F, aux4, emfMode, emf7);
    if (solver == SOLVER_ROE) return byDir<P, SOLVER_ROE>(dir, eos, n, vL, vR, F, aux4, emfMode, emf7);
  } else {
    if (solver == SOLVER_HLLC) return byDir<P, SOLVER_HLLC>(dir, eos, n, vL, vR, F, aux4, emfMode, emf7);
    if (solver == SOLVER_ROE) return byDir<P, SOLVER_ROE>(dir, eos, n, vL, vR, F, aux4, emfMode, emf7);
  }
  return 2;
}

template <int ND, int NC>
static int byPhysics(int mhd, int iso, int solver, int dir, const Eos &eos, long n, const double *vL,
                     const double *vR, double *F, double *aux4, int emfMode, double *emf7) {
  if (mhd && iso) return bySolver<Phys<ND, NC, true, true>>(solver, dir, eos, n, vL, vR, F, aux4, emfMode, emf7);
  if (mhd) return bySolver<Phys<ND, NC, true, false>>(solver, dir, eos, n, vL, vR, F, aux4, emfMode, emf7);
  if (iso) return bySolver<Phys<ND, NC, false, true>>(solver, dir, eos, n, vL, vR, F, aux4, emfMode, emf7);
  return bySolver<Phys<ND, NC, false, false>>(solver, dir, eos, n, vL, vR, F, aux4, emfMode, emf7);
}

// states are packed (n, nvar) in the reference's variable order; aux4 receives (cmax, sl, sr, c2iso) per state and,
// for MHD with dir < ndim and emf7 != NULL, emf7 receives (Et, Eb, SV, aL, aR, dL, dR) as stored for emfMode.
// returns 0, or non-zero for a combination the product does not implement
extern "C" int idfx_host_riemann(int ndim, int ncomp, int mhd, int iso, int solver, int dir, double gamma,
                                 double csiso, long n, const double *vL, const double *vR, double *F, double *aux4,
                                 int emfMode, double *emf7) {
  Eos eos;
  eos.gamma = gamma;
  eos.csiso = csiso;
#define IDFX_CASE(ND, NC) \
  if (ndim == ND && ncomp == NC) \
    return byPhysics<ND, NC>(mhd, iso, solver, dir, eos, n, vL, vR, F, aux4, emfMode, emf7);
  IDFX_CASE(1, 1)
  IDFX_CASE(2, 2)
  IDFX_CASE(3, 3)
  IDFX_CASE(2, 3)
  IDFX_CASE(1, 3)
#undef IDFX_CASE
  return 3;
}

// ---- reconstruction (recon.cuh): the per-CELL form used by the pencil sweeps, evaluated along a 1-D array and
// re-assembled per interface: vL[i] = right face value of cell i-1, vR[i] = left face value of cell i ----------
#include "recon.cuh"

template <int ORDER>
static void reconstructAll(long n, const double *q, int positive, double dx, double *vL, double *vR) {
  const double wc[6] = {0, 0, 0, 0, 0, 0};
  for (long i = 3; i < n - 2; i++) {
    double lo[2], hi[2];
    for (int side = 0; side < 2; side++) {
      const long c = i - 1 + side;
      const double st[5] = {q[c - 2], q[c - 1], q[c], q[c + 1], q[c + 2]};
      reconstructCell<ORDER, false>(st, positive != 0, true, wc, dx, lo[side], hi[side]);
    }
    vL[i] = hi[0];
    vR[i] = lo[1];
  }
}

extern "C" int idfx_host_reconstruct(int order, long n, const double *q, int positive, double dx, double *vL,
                                     double *vR) {
  if (order == 2) reconstructAll<2>(n, q, positive, dx, vL, vR);
  else if (order == 3) reconstructAll<3>(n, q, positive, dx, vL, vR);
  else if (order == 4) reconstructAll<4>(n, q, positive, dx, vL, vR);
  else return 1;
  return 0;
}

// ---- conservative <-> primitive conversion with the pressure floor (phys.cuh) -----------------------------------
template <class P>
static void convertAll(const Eos &eos, long n, double *u, double *v, double *u2) {
  for (long q = 0; q < n; q++) {
    consToPrim<P>(v + q * P::nvar, u + q * P::nvar, eos);       // may rewrite the energy
    primToCons<P>(u2 + q * P::nvar, v + q * P::nvar, eos);
  }
}

extern "C" int idfx_host_convert(int ncomp, int mhd, int iso, double gamma, double csiso, long n, double *u, double *v,
                                 double *u2) {
  Eos eos;
  eos.gamma = gamma;
  eos.csiso = csiso;
#define IDFX_CASE(NC, M, I) \
  if (ncomp == NC && mhd == M && iso == I) { convertAll<Phys<NC, NC, M != 0, I != 0>>(eos, n, u, v, u2); return 0; }
  IDFX_CASE(1, 0, 0) IDFX_CASE(2, 0, 0) IDFX_CASE(3, 0, 0) IDFX_CASE(1, 0, 1) IDFX_CASE(3, 0, 1)
  IDFX_CASE(1, 1, 0) IDFX_CASE(2, 1, 0) IDFX_CASE(3, 1, 0) IDFX_CASE(2, 1, 1) IDFX_CASE(3, 1, 1)
#undef IDFX_CASE
  return 1;
}

// ---- irregular-grid reconstruction (recon.cuh, every non-cartesian geometry): per-cell weights {wp, wm, cp, cm, dp, dm}
// of extrapolateToFaces.hpp:43-71 and the local cell width (LimO3) ------------------------------------------------
template <int ORDER>
static void reconstructIrregularAll(long n, const double *q, int positive, const double *w6, const double *dx, double *vL,
                                    double *vR) {
  for (long i = 3; i < n - 2; i++) {
    double lo[2], hi[2];
    for (int side = 0; side < 2; side++) {
      const long c = i - 1 + side;
      const double st[5] = {q[c - 2], q[c - 1], q[c], q[c + 1], q[c + 2]};
      const double wc[6] = {w6[6 * c], w6[6 * c + 1], w6[6 * c + 2], w6[6 * c + 3], w6[6 * c + 4], w6[6 * c + 5]};
      reconstructCell<ORDER, false>(st, positive != 0, false, wc, dx[c], lo[side], hi[side]);
    }
    vL[i] = hi[0];
    vR[i] = lo[1];
  }
}
extern "C" int idfx_host_reconstruct_irregular(int order, long n, const double *q, int positive, const double *w6,
                                               const double *dx, double *vL, double *vR) {
  if (order == 2) reconstructIrregularAll<2>(n, q, positive, w6, dx, vL, vR);
  else if (order == 3) reconstructIrregularAll<3>(n, q, positive, w6, dx, vL, vR);
  else if (order == 4) reconstructIrregularAll<4>(n, q, positive, w6, dx, vL, vR);
  else return 1;
  return 0;
}

// ---- geometrical source terms of polar / spherical grids (sweep.cuh::geometricSource, addSourceTerms.hpp:96-192) ---
// (sweep.cuh also holds updateCell, whose fire-and-forget accumulation is a device atomic: give the host pass of this
// test build a plain stand-in so that the header compiles; updateCell itself is not called here)
template <class T>
static __host__ __device__ inline T idfxHostAtomicAdd(T *p, T v) {
  const T old = *p;
  *p = old + v;
  return old;
}
#define atomicAdd idfxHostAtomicAdd
#include "sweep.cuh"
#undef atomicAdd

template <class P>
static void sourceAll(const DevGrid &g, const Eos &eos, double dt, long n, const double *v, const int *jj, const int *ii,
                      double *s1, double *s2) {
  for (long c = 0; c < n; c++) geometricSource<P>(g, v + c * P::nvar, eos, dt, jj[c], ii[c], s1[c], s2[c]);
}
// x1 / rt: per radial index, tanx2: per latitudinal index; the n cells are addressed by (jj, ii)
extern "C" int idfx_host_geometric_source(int geometry, int ndim, int ncomp, int mhd, int iso, double gamma, double csiso,
                                          double dt, long n, const double *v, const int *jj, const int *ii, const double *x1,
                                          const double *rt, const double *tanx2, double *s1, double *s2) {
  DevGrid g{};
  g.geometry = geometry;
  g.ndim = ndim;
  g.x[0] = x1;
  g.rt = rt;
  g.tanx2 = tanx2;
  const Eos eos{gamma, csiso};
#define SRC_CASE(ND, NC, M, I) \
  if (ndim == ND && ncomp == NC && mhd == M && iso == I) { sourceAll<Phys<ND, NC, (M != 0), (I != 0)>>(g, eos, dt, n, v, jj, ii, s1, s2); return 0; }
  SRC_CASE(3, 3, 1, 0) SRC_CASE(2, 3, 1, 0) SRC_CASE(2, 2, 0, 0) SRC_CASE(3, 3, 0, 0) SRC_CASE(2, 2, 1, 0)
  SRC_CASE(2, 3, 1, 1) SRC_CASE(2, 3, 0, 1)
#undef SRC_CASE
  return 1;
}

// ---- area weighting of the interface flux and the conservative update of a cell in curvilinear geometry
// (sweep.cuh::areaWeightFlux + updateCell: calcRightHandSide.hpp:152-195, 330-551, addSourceTerms.hpp) ----------------
// Arrays in the reference's layouts: Vc / Uc / Flux (nvar, k, j, i), InvDt / cMax / area / vol (k, j, i); area[c] = area
// of the lower face of cell c along DIR.  Every active cell is updated in place like the sweeps do it.
template <class P, int DIR>
static void rhsAll(const DevGrid &g, const Eos &eos, double dt, int doSource, const double *Vc, double *Uc, double *invDt,
                   const double *Flux, const double *cMax, const double *area, const double *vol) {
  SweepArgs a{};
  a.g = g;
  a.Vc = Vc;
  a.Uc = Uc;
  a.invDt = invDt;
  a.doInvDt = 1;
  a.doSource = doSource;
  const long long s = DIR == 0 ? 1 : (DIR == 1 ? g.sj : g.sk);
  for (int k = g.beg[2]; k < g.end[2]; k++)
    for (int j = g.beg[1]; j < g.end[1]; j++)
      for (int i = g.beg[0]; i < g.end[0]; i++) {
        const long long c = cidx(g, k, j, i);
        double Fm[P::nvar], Fp[P::nvar];
        for (int nv = 0; nv < P::nvar; nv++) {
          Fm[nv] = Flux[nv * g.sn + c];
          Fp[nv] = Flux[nv * g.sn + c + s];
        }
        areaWeightFlux<P, DIR>(g, g.geometry, area[c], k, j, i, Fm);
        areaWeightFlux<P, DIR>(g, g.geometry, area[c + s], k + (DIR == 2), j + (DIR == 1), i + (DIR == 0), Fp);
        const int idir = DIR == 0 ? i : (DIR == 1 ? j : k);
        updateCell<P, DIR>(a, g, g.geometry, eos, dt, vol[c], g.dx[DIR][idir], k, j, i, c, Fm, Fp, cMax[c], cMax[c + s],
                           [&](int nv) { return Vc[nv * g.sn + c]; });
      }
}
template <class P>
static int rhsByDir(int dir, const DevGrid &g, const Eos &eos, double dt, int doSource, const double *Vc, double *Uc,
                    double *invDt, const double *Flux, const double *cMax, const double *area, const double *vol) {
  if (dir == 0) rhsAll<P, 0>(g, eos, dt, doSource, Vc, Uc, invDt, Flux, cMax, area, vol);
  else if (dir == 1) rhsAll<P, 1>(g, eos, dt, doSource, Vc, Uc, invDt, Flux, cMax, area, vol);
  else if (dir == 2 && P::ndim == 3) rhsAll<P, (P::ndim == 3 ? 2 : 0)>(g, eos, dt, doSource, Vc, Uc, invDt, Flux, cMax, area, vol);
  else return 1;
  return 0;
}
// coords: x1, x1l, dx1, rt (n1 each), dx2, sinx2, sinx2m, dmu, tanx2 (n2 each), in this order
extern "C" int idfx_host_curvilinear_rhs(int geometry, int ndim, int ncomp, int mhd, int iso, int dir, double gamma,
                                         double csiso, double dt, int doSource, const int *nx, const int *ng,
                                         const double *const *coords, const double *Vc, double *Uc, double *invDt,
                                         const double *Flux, const double *cMax, const double *area, const double *vol) {
  DevGrid g{};
  g.geometry = geometry;
  g.ndim = ndim;
  for (int d = 0; d < 3; d++) {
    g.nx[d] = nx[d];
    g.ng[d] = ng[d];
    g.beg[d] = ng[d];
    g.end[d] = nx[d] - ng[d];
    g.off[d] = d < ndim ? 1 : 0;
  }
  g.sj = nx[0];
  g.sk = (long long)nx[0] * nx[1];
  g.sn = g.sk * nx[2];
  g.x[0] = coords[0];
  g.xl[0] = coords[1];
  g.dx[0] = coords[2];
  g.rt = coords[3];
  g.dx[1] = coords[4];
  g.sinx2 = coords[5];
  g.sinx2m = coords[6];
  g.dmu = coords[7];
  g.tanx2 = coords[8];
  g.dx[2] = coords[9];
  const Eos eos{gamma, csiso};
#define RHS_CASE(ND, NC, M, I) \
  if (ndim == ND && ncomp == NC && mhd == M && iso == I)  \
    return rhsByDir<Phys<ND, NC, (M != 0), (I != 0)>>(dir, g, eos, dt, doSource, Vc, Uc, invDt, Flux, cMax, area, vol);
  RHS_CASE(3, 3, 1, 0) RHS_CASE(2, 3, 1, 0) RHS_CASE(2, 2, 0, 0) RHS_CASE(3, 3, 0, 0) RHS_CASE(2, 3, 1, 1)
#undef RHS_CASE
  return 1;
}
