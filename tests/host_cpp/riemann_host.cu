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

template <class P, int SOLVER, int DIR>
static void solveAll(const Eos &eos, long n, const double *vL, const double *vR, double *F, double *aux4) {
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
    }
    for (int v = 0; v < P::nvar; v++) F[q * P::nvar + v] = f[v];
    aux4[4 * q + 0] = aux.cmax;
    aux4[4 * q + 1] = aux.sl;
    aux4[4 * q + 2] = aux.sr;
    aux4[4 * q + 3] = aux.c2iso;
  }
}

template <class P, int SOLVER>
static int byDir(int dir, const Eos &eos, long n, const double *vL, const double *vR, double *F, double *aux4) {
  if (dir == 0) { solveAll<P, SOLVER, 0>(eos, n, vL, vR, F, aux4); return 0; }
  if constexpr (P::ncomp >= 2) if (dir == 1) { solveAll<P, SOLVER, 1>(eos, n, vL, vR, F, aux4); return 0; }
  if constexpr (P::ncomp == 3) if (dir == 2) { solveAll<P, SOLVER, 2>(eos, n, vL, vR, F, aux4); return 0; }
  return 1;
}

template <class P>
static int bySolver(int solver, int dir, const Eos &eos, long n, const double *vL, const double *vR, double *F,
                    double *aux4) {
  if (solver == SOLVER_TVDLF) return byDir<P, SOLVER_TVDLF>(dir, eos, n, vL, vR, F, aux4);
  if (solver == SOLVER_HLL) return byDir<P, SOLVER_HLL>(dir, eos, n, vL, vR, F, aux4);
  if constexpr (P::mhd) {
    if (solver == SOLVER_HLLD) return byDir<P, SOLVER_HLLD>(dir, eos, n, vL, vR, F, aux4);
    if constexpr (!P::iso) if (solver == SOLVER_ROE) return byDir<P, SOLVER_ROE>(dir, eos, n, vL, vR, F, aux4);
  } else {
    if (solver == SOLVER_HLLC) return byDir<P, SOLVER_HLLC>(dir, eos, n, vL, vR, F, aux4);
  }
  return 2;
}

template <int NC>
static int byPhysics(int mhd, int iso, int solver, int dir, const Eos &eos, long n, const double *vL,
                     const double *vR, double *F, double *aux4) {
  if (mhd && iso) return bySolver<Phys<NC, NC, true, true>>(solver, dir, eos, n, vL, vR, F, aux4);
  if (mhd) return bySolver<Phys<NC, NC, true, false>>(solver, dir, eos, n, vL, vR, F, aux4);
  if (iso) return bySolver<Phys<NC, NC, false, true>>(solver, dir, eos, n, vL, vR, F, aux4);
  return bySolver<Phys<NC, NC, false, false>>(solver, dir, eos, n, vL, vR, F, aux4);
}

// states are packed (n, nvar) in the reference's variable order; aux4 receives (cmax, sl, sr, c2iso) per state.
// returns 0, or non-zero for a combination the product does not implement
extern "C" int idfx_host_riemann(int ncomp, int mhd, int iso, int solver, int dir, double gamma, double csiso,
                                 long n, const double *vL, const double *vR, double *F, double *aux4) {
  Eos eos;
  eos.gamma = gamma;
  eos.csiso = csiso;
  if (ncomp == 1) return byPhysics<1>(mhd, iso, solver, dir, eos, n, vL, vR, F, aux4);
  if (ncomp == 2) return byPhysics<2>(mhd, iso, solver, dir, eos, n, vL, vR, F, aux4);
  if (ncomp == 3) return byPhysics<3>(mhd, iso, solver, dir, eos, n, vL, vR, F, aux4);
  return 3;
}
