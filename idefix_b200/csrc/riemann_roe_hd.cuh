// riemann_roe_hd.cuh -- Roe solver for hydrodynamics (the solver of the reference's default test/HD/sod run).
//
// Reference: src/fluid/RiemannSolver/HDsolvers/roeHD.hpp:41-360 with ROE_AVERAGE 0 (arithmetic mean state,
// roeHD.hpp:17), entropy fix delta = 1e-7 on the two acoustic modes, HLL flux across strong shocks
// (|dp|/min(p) > 1/2 in a compression; isothermal: |drho|/min(rho) * a^2).
//
// Verified bit for bit against the oracle on the host (tests/test_riemann_host.py); dispatched by the sweep kernels
// (sweep.cuh::riemannDispatch, sweep_hd_roe.cu); GPU parity: tests/test_roe_hd_iso_gpu.py.
// Trap: with DIMENSIONS == 1 the reference compiles the HLL branch out (roeHD.hpp:314) and leaves
// Flux(nv,k,j,i) as the previous call wrote it; this function then returns true and leaves F untouched, and the
// x1 sweep substitutes the flux it kept from its previous call (SweepArgs::staleFlux, sweep_pencil.cuh).
#ifndef IDFX_RIEMANN_ROE_HD_CUH_
#define IDFX_RIEMANN_ROE_HD_CUH_

#include "riemann_hd.cuh"

template <class P, int DIR, class VL, class VR>
IDFX_D bool riemannRoeHD(const VL &vL, const VR &vR, const Eos &eos, double *F, RiemannAux &aux) {
  static_assert(!P::mhd, "hydrodynamics only");
  using D = Dirs<P, DIR>;
  constexpr int Xn = D::Xn;
  constexpr int NV = P::nvar;
  // mode numbering of roeHD.hpp:20-36
  constexpr int I0 = 0, I1 = 1, IE = 2, I2 = (P::energy ? 3 : 2), I3 = (P::energy ? 4 : 3);
  double uL[NV], uR[NV], fL[NV], fR[NV], dv[NV], um[NV];
  double Rc[NV][NV], lambda[NV], eta[NV];
#pragma unroll
  for (int nv = 0; nv < NV; nv++) dv[nv] = vR[nv] - vL[nv];
  double a2L, a2R;
  if constexpr (P::energy) {
    a2L = idfxSqrt<P::fast>(eos.gamma * idfxDiv<P::fast>(vL[P::PRS], vL[P::RHO]));
    a2R = idfxSqrt<P::fast>(eos.gamma * idfxDiv<P::fast>(vR[P::PRS], vR[P::RHO]));
  } else {
    a2L = 0.5 * (eos.csiso + eos.csiso);
    a2R = a2L;
  }
  a2L = a2L * a2L;
  a2R = a2R * a2R;
  primToCons<P>(uL, vL, eos);
  primToCons<P>(uR, vR, eos);
  physFlux<P, DIR>(fL, vL, uL, a2L);
  physFlux<P, DIR>(fR, vR, uR, a2R);
#pragma unroll
  for (int nv = 0; nv < NV; nv++) um[nv] = 0.5 * (vR[nv] + vL[nv]);
  double a, a2, h = 0.0, vel2 = 0.0;
  if constexpr (P::energy) {
    const double gamma_m1 = eos.gamma - 1;
    a2 = idfxDiv<P::fast>(eos.gamma * um[P::PRS], um[P::RHO]);
    a = idfxSqrt<P::fast>(a2);
    vel2 = um[P::MX1] * um[P::MX1];
    if constexpr (P::ncomp >= 2) vel2 = vel2 + um[P::MX2] * um[P::MX2];
    if constexpr (P::ncomp == 3) vel2 = vel2 + um[P::MX3] * um[P::MX3];
    h = 0.5 * vel2 + idfxDiv<P::fast>(a2, gamma_m1);
  } else {
    a2 = 0.5 * (a2L + a2R);
    a = idfxSqrt<P::fast>(a2);
  }
#pragma unroll
  for (int p = 0; p < NV; p++) {
    lambda[p] = 0.0;
    eta[p] = 0.0;
#pragma unroll
    for (int q = 0; q < NV; q++) Rc[p][q] = 0;
  }
  // u - c_s and u + c_s
  lambda[I0] = um[Xn] - a;
  lambda[I1] = um[Xn] + a;
  if constexpr (P::energy) {
    eta[I0] = idfxDiv<P::fast>(0.5, a2) * (dv[P::PRS] - dv[Xn] * um[P::RHO] * a);
    eta[I1] = idfxDiv<P::fast>(0.5, a2) * (dv[P::PRS] + dv[Xn] * um[P::RHO] * a);
  } else {
    eta[I0] = 0.5 * (dv[P::RHO] - idfxDiv<P::fast>(um[P::RHO] * dv[Xn], a));
    eta[I1] = 0.5 * (dv[P::RHO] + idfxDiv<P::fast>(um[P::RHO] * dv[Xn], a));
  }
  Rc[P::RHO][I0] = 1.0;
  Rc[P::RHO][I1] = 1.0;
  Rc[Xn][I0] = um[Xn] - a;
  Rc[Xn][I1] = um[Xn] + a;
  if constexpr (P::ncomp >= 2) {
    Rc[D::Xt][I0] = um[D::Xt];
    Rc[D::Xt][I1] = um[D::Xt];
  }
  if constexpr (P::ncomp == 3) {
    Rc[D::Xb][I0] = um[D::Xb];
    Rc[D::Xb][I1] = um[D::Xb];
  }
  if constexpr (P::energy) {
    Rc[P::ENG][I0] = h - um[Xn] * a;
    Rc[P::ENG][I1] = h + um[Xn] * a;
    // entropy mode
    lambda[IE] = um[Xn];
    eta[IE] = dv[P::RHO] - idfxDiv<P::fast>(dv[P::PRS], a2);
    Rc[P::RHO][IE] = 1.0;
    Rc[P::MX1][IE] = um[P::MX1];
    if constexpr (P::ncomp >= 2) Rc[P::MX2][IE] = um[P::MX2];
    if constexpr (P::ncomp == 3) Rc[P::MX3][IE] = um[P::MX3];
    Rc[P::ENG][IE] = 0.5 * vel2;
  }
  // shear modes
  if constexpr (P::ncomp >= 2) {
    lambda[I2] = um[Xn];
    eta[I2] = um[P::RHO] * dv[D::Xt];
    Rc[D::Xt][I2] = 1.0;
    if constexpr (P::energy) Rc[P::ENG][I2] = um[D::Xt];
  }
  if constexpr (P::ncomp == 3) {
    lambda[I3] = um[Xn];
    eta[I3] = um[P::RHO] * dv[D::Xb];
    Rc[D::Xb][I3] = 1.0;
    if constexpr (P::energy) Rc[P::ENG][I3] = um[D::Xb];
  }
  aux.cmax = fabs(um[Xn]) + a;
  aux.sl = lambda[I0];
  aux.sr = lambda[I1];
  aux.c2iso = 0.0;

  double scrh;
  if constexpr (P::energy) {
    scrh = fabs(vL[P::PRS] - vR[P::PRS]);
    scrh = idfxDiv<P::fast>(scrh, fmin(vL[P::PRS], vR[P::PRS]));
  } else {
    scrh = fabs(vL[P::RHO] - vR[P::RHO]);
    scrh = idfxDiv<P::fast>(scrh, fmin(vL[P::RHO], vR[P::RHO]));
    scrh *= a * a;
  }
  if (scrh > 0.5 && (vR[Xn] < vL[Xn])) {
    if constexpr (P::ndim > 1) {
      const double bmin = fmin(0.0, lambda[0]);
      const double bmax = fmax(0.0, lambda[1]);
      const double scrh1 = idfxDiv<P::fast>(1.0, (bmax - bmin));
#pragma unroll
      for (int nv = 0; nv < NV; nv++) {
        F[nv] = bmin * bmax * (uR[nv] - uL[nv]) + bmax * fL[nv] - bmin * fR[nv];
        F[nv] *= scrh1;
      }
      return false;
    } else {
      return true;   // the reference leaves the previous flux in place (roeHD.hpp:314)
    }
  }
  double alambda[NV];
#pragma unroll
  for (int nv = 0; nv < NV; nv++) alambda[nv] = fabs(lambda[nv]);
  const double delta = 1.e-7;
  if (alambda[0] <= delta) alambda[0] = idfxDiv<P::fast>(0.5 * lambda[0] * lambda[0], delta) + 0.5 * delta;
  if (alambda[1] <= delta) alambda[1] = idfxDiv<P::fast>(0.5 * lambda[1] * lambda[1], delta) + 0.5 * delta;
#pragma unroll
  for (int nv = 0; nv < NV; nv++) {
    double f = fL[nv] + fR[nv];
#pragma unroll
    for (int nv2 = 0; nv2 < NV; nv2++) f -= alambda[nv2] * eta[nv2] * Rc[nv][nv2];
    F[nv] = f * 0.5;
  }
  return false;
}

#endif  // IDFX_RIEMANN_ROE_HD_CUH_
