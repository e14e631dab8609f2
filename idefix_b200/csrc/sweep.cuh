// sweep.cuh -- one directional sweep of the Godunov update as a single kernel.
//
// It fuses what the reference runs as three kernels per direction with two
// DRAM round trips of the 8-variable flux array:
//   K8  CalcRiemannFlux  (e.g. hlldMHD.hpp:147)   reconstruct + Riemann solver + EMF storage
//   K9  Correct Flux     (calcRightHandSide.hpp:16-196)
//   K10 CalcRightHandSide(calcRightHandSide.hpp:201-551)  flux divergence, InvDt, Uc update
// (+ K11 AddSourceTerms, addSourceTerms.hpp:16-195, after the last direction).
//
// This header holds the per-interface / per-cell building blocks; the kernels that use them
// are in sweep_pencil.cuh.
#ifndef IDFX_SWEEP_CUH_
#define IDFX_SWEEP_CUH_

#include "common.cuh"
#include "phys.cuh"
#include "recon.cuh"
#include "riemann_hd.cuh"
#include "riemann_mhd.cuh"
#include "riemann_roe.cuh"
#include "emf_store.cuh"
#include "riemann_roe_hd.cuh"

// Returns true when the solver asks the caller to keep the flux of the PREVIOUS call at this interface: only the
// HD Roe solver in one dimension, across a strong shock (roeHD.hpp:312-328, the HLL branch is compiled out there and
// Flux(nv,k,j,i) is left as the last CalcRightHandSide wrote it, i.e. area-weighted).  The x1 sweep then substitutes
// SweepArgs::staleFlux.
template <class P, int SOLVER, int DIR, class VL, class VR>
IDFX_D bool riemannDispatch(const VL &vL, const VR &vR, const Eos &eos, double *F, RiemannAux &aux) {
  if constexpr (P::mhd) {
    if constexpr (SOLVER == SOLVER_TVDLF) riemannTvdlfMHD<P, DIR>(vL, vR, eos, F, aux);
    if constexpr (SOLVER == SOLVER_HLL) riemannHllMHD<P, DIR>(vL, vR, eos, F, aux);
    if constexpr (SOLVER == SOLVER_HLLD) riemannHlldMHD<P, DIR>(vL, vR, eos, F, aux);
    if constexpr (SOLVER == SOLVER_ROE) riemannRoeMHD<P, DIR>(vL, vR, eos, F, aux);
  } else {
    if constexpr (SOLVER == SOLVER_TVDLF) riemannTvdlfHD<P, DIR>(vL, vR, eos, F, aux);
    if constexpr (SOLVER == SOLVER_HLL) riemannHllHD<P, DIR>(vL, vR, eos, F, aux);
    if constexpr (SOLVER == SOLVER_HLLC) riemannHllcHD<P, DIR>(vL, vR, eos, F, aux);
    if constexpr (SOLVER == SOLVER_ROE) return riemannRoeHD<P, DIR>(vL, vR, eos, F, aux);
  }
  return false;
}

// Index of the azimuthal momentum / field for the curvilinear corrections (idefix.hpp:129-187)
template <class P>
struct CurvIdx {
  // returns -1 when the component does not exist
  static IDFX_HD int mphi(int geo) {
    if (geo == GEO_CYLINDRICAL) return (P::ncomp == 3 ? P::MX3 : -1);
    if (geo == GEO_POLAR) return (P::ncomp >= 2 ? P::MX2 : -1);
    if (geo == GEO_SPHERICAL) return (P::ncomp == 3 ? P::MX3 : -1);
    return -1;
  }
  static IDFX_HD int bphi(int geo) {
    if (!P::mhd) return -1;
    if (geo == GEO_CYLINDRICAL) return (P::ncomp == 3 ? P::BX3 : -1);
    if (geo == GEO_POLAR) return (P::ncomp >= 2 ? P::BX2 : -1);
    if (geo == GEO_SPHERICAL) return (P::ncomp == 3 ? P::BX3 : -1);
    return -1;
  }
  static IDFX_HD int bth(int geo) {
    if (!P::mhd) return -1;
    if (geo == GEO_SPHERICAL) return (P::ncomp >= 2 ? P::BX2 : -1);
    return -1;
  }
};

// Geometrical source terms of one cell (addSourceTerms.hpp:96-192), rotation/Fargo excluded.
template <class P, class V>
IDFX_D void geometricSource(const DevGrid &g, const V &v, const Eos &eos, double dt, int j, int i,
                            double &s1, double &s2) {
  s1 = 0.0;
  s2 = 0.0;
  const int geo = g.geometry;
  [[maybe_unused]] double c2iso = eos.csiso * eos.csiso;
  if (geo == GEO_CYLINDRICAL) {
    if constexpr (P::ncomp == 3) {
      double vphi = v[P::MX3];
      double Sm = v[P::RHO] * vphi * vphi;
      if constexpr (P::iso) Sm += v[P::RHO] * c2iso;
      else Sm += v[P::PRS];
      if constexpr (P::mhd) {
        Sm -= v[P::BX3] * v[P::BX3];
        Sm += 0.5 * (v[P::BX1] * v[P::BX1] + v[P::BX2] * v[P::BX2] + v[P::BX3] * v[P::BX3]);
      }
      s1 = idfxDiv<P::fast>(dt * Sm, g.x[0][i]);
    }
  } else if (geo == GEO_POLAR) {
    if constexpr (P::ncomp >= 2) {
      double vphi = v[P::MX2] + 0.0;
      double Sm = v[P::RHO] * vphi * vphi;
      if constexpr (P::iso) Sm += v[P::RHO] * c2iso;
      else Sm += v[P::PRS];
      if constexpr (P::mhd) {
        Sm -= v[P::BX2] * v[P::BX2];
        double b2 = v[P::BX1] * v[P::BX1];
        b2 = b2 + v[P::BX2] * v[P::BX2];
        if constexpr (P::ncomp == 3) b2 = b2 + v[P::BX3] * v[P::BX3];
        Sm += 0.5 * b2;
      }
      s1 = idfxDiv<P::fast>(dt * Sm, g.x[0][i]);
    }
  } else if (geo == GEO_SPHERICAL) {
    double vphi = 0.0;
    if constexpr (P::ncomp == 3) vphi = v[P::MX3];
    vphi = vphi + 0.0;
    double vv = 0.0;
    if constexpr (P::ncomp >= 2) vv = vv + v[P::MX2] * v[P::MX2];
    if constexpr (P::ncomp == 3) vv = vv + vphi * vphi;
    double Sm = v[P::RHO] * vv;
    if constexpr (P::iso) Sm += 2.0 * v[P::RHO] * c2iso;
    else Sm += 2.0 * v[P::PRS];
    if constexpr (P::mhd) {
      double hs = 0.0;
      if constexpr (P::ncomp >= 2) hs = hs + v[P::BX2] * v[P::BX2];
      if constexpr (P::ncomp == 3) hs = hs + v[P::BX3] * v[P::BX3];
      Sm -= hs;
      double b2 = v[P::BX1] * v[P::BX1];
      if constexpr (P::ncomp >= 2) b2 = b2 + v[P::BX2] * v[P::BX2];
      if constexpr (P::ncomp == 3) b2 = b2 + v[P::BX3] * v[P::BX3];
      Sm += b2;
    }
    s1 = idfxDiv<P::fast>(dt * Sm, g.rt[i]);
    if constexpr (P::ncomp >= 2) {
      const double ct = idfxDiv<P::fast>(1.0, g.tanx2[j]);
      double cc = 0.0;
      cc = cc - v[P::MX2] * v[P::MX1];
      if constexpr (P::ncomp == 3) cc = cc + ct * vphi * vphi;
      Sm = v[P::RHO] * cc;
      if constexpr (P::iso) Sm += ct * c2iso * v[P::RHO];
      else Sm += ct * v[P::PRS];
      if constexpr (P::mhd) {
        double hs = 0.0;
        hs = hs + v[P::BX2] * v[P::BX1];
        if constexpr (P::ncomp == 3) hs = hs - ct * v[P::BX3] * v[P::BX3];
        Sm += hs;
        double b2 = v[P::BX1] * v[P::BX1];
        b2 = b2 + v[P::BX2] * v[P::BX2];
        if constexpr (P::ncomp == 3) b2 = b2 + v[P::BX3] * v[P::BX3];
        Sm += 0.5 * ct * b2;
      }
      s2 = idfxDiv<P::fast>(dt * Sm, g.rt[i]);
    }
  }
}

// K9: area weighting and curvilinear factors of the interface flux (calcRightHandSide.hpp:152-195)
template <class P, int DIR>
IDFX_D void areaWeightFlux(const DevGrid &g, int geo, double A0, int k, int j, int i, double *F) {
  constexpr int NV = P::nvar;
  const int mphi = CurvIdx<P>::mphi(geo), bphi = CurvIdx<P>::bphi(geo), bth = CurvIdx<P>::bth(geo);
#pragma unroll
  for (int nv = 0; nv < NV; nv++) {
    double Ax = A0;
    if (geo != GEO_CARTESIAN) {
      if constexpr (DIR == IDIR) {
        if (geo == GEO_POLAR || geo == GEO_CYLINDRICAL) {
          if (nv == mphi) Ax *= fabs(g.xl[0][i]);
          if (nv == bphi) Ax = 1.0;
        } else if (geo == GEO_SPHERICAL) {
          if (nv == mphi) Ax *= fabs(g.xl[0][i]);
          if (nv == bth || nv == bphi) Ax = g.xl[0][i];
        }
      } else if constexpr (DIR == JDIR) {
        if (geo == GEO_SPHERICAL && P::ncomp == 3) {
          if (nv == mphi) Ax *= fabs(g.sinx2m[j]);
          if (nv == bphi) Ax = 1.0;
        }
      }
    }
    F[nv] = F[nv] * Ax;
  }
}
template <class P, int DIR>
IDFX_D void areaWeightFlux(const DevGrid &g, int geo, int k, int j, int i, double *F) {
  areaWeightFlux<P, DIR>(g, geo, faceArea<DIR>(g, k, j, i), k, j, i, F);
}

// K10 (+K11): flux divergence, time-step estimate, geometrical sources and conservative update of
// one cell from the area-weighted fluxes of its two interfaces along DIR
// (calcRightHandSide.hpp:201-551, addSourceTerms.hpp:16-195).  idir = index of the cell along DIR.
// vcAt(nv) / ucAt(nv) / invDtPrev() fetch the cell's primitives, conservatives and InvDt accumulator
// from wherever the calling kernel keeps them; they are only called when the value is needed.
// dV = volume of the cell, dxc = its width along DIR (both as the reference stores them).
template <class P, int DIR, class VcF>
IDFX_D void updateCell(const SweepArgs &a, const DevGrid &g, int geo, const Eos &eos, double dt, double dV, double dxc,
                       int k, int j, int i, long long cell, const double *Fm, const double *Fp, double cmaxm,
                       double cmaxp, VcF vcAt) {
  constexpr int NV = P::nvar;
  const double dtdV = idfxDiv<P::fast>(dt, dV);
  double rhs[NV];
#pragma unroll
  for (int nv = 0; nv < NV; nv++) rhs[nv] = -dtdV * (Fp[nv] - Fm[nv]);

  if (geo != GEO_CARTESIAN) {
    const int mphi = CurvIdx<P>::mphi(geo), bphi = CurvIdx<P>::bphi(geo), bth = CurvIdx<P>::bth(geo);
    if constexpr (DIR == IDIR) {
#pragma unroll
      for (int nv = 0; nv < NV; nv++) {
        if (nv == mphi) rhs[nv] = idfxDiv<P::fast>(rhs[nv], g.x[0][i]);
        if constexpr (P::mhd) {
          if (geo == GEO_POLAR || geo == GEO_CYLINDRICAL) {
            if (nv == bphi) rhs[nv] = idfxDiv<P::fast>(-dt, g.dx[0][i]) * (Fp[nv] - Fm[nv]);
          } else if (geo == GEO_SPHERICAL) {
            if (nv == bth || nv == bphi) {
              const double q = idfxDiv<P::fast>(dt, (g.x[0][i] * g.dx[0][i]));
              rhs[nv] = -q * ((Fp[nv] - Fm[nv]));
            }
          }
        }
      }
    } else if constexpr (DIR == JDIR) {
      if (geo == GEO_SPHERICAL && P::ncomp == 3) {
#pragma unroll
        for (int nv = 0; nv < NV; nv++) {
          if (nv == mphi) rhs[nv] = idfxDiv<P::fast>(rhs[nv], fabs(g.sinx2[j]));
          if constexpr (P::mhd) {
            if (nv == bphi) rhs[nv] = idfxDiv<P::fast>(-dt, (g.rt[i] * g.dx[1][j])) * (Fp[nv] - Fm[nv]);
          }
        }
      }
    }
  }

  // time-step control (calcRightHandSide.hpp:426-531)
  if (a.doInvDt) {
    double dl = dxc;
    if (geo == GEO_POLAR) {
      if constexpr (DIR == JDIR) dl = dl * g.x[0][i];
    } else if (geo == GEO_SPHERICAL) {
      if constexpr (DIR == JDIR) dl = dl * g.rt[i];
      if constexpr (DIR == KDIR) dl = idfxDiv<P::fast>(dl * g.rt[i] * g.dmu[j], g.dx[1][j]);
    }
    // invDtInit: first sweep after ResetStage, the accumulator is known to be zero.  Otherwise the
    // accumulation InvDt += term is a fire-and-forget RED.ADD.F64 resolved in L2 (same IEEE addition as
    // load-add-store, but no load latency and no register / shared-memory staging of the old value)
    const double term = idfxDiv<P::fast>(0.5 * (cmaxp + cmaxm), (dl));
    if (a.invDtInit) a.invDt[cell] = 0.0 + term;
    else atomicAdd(a.invDt + cell, term);
  }

  // conservative update; CT-evolved field components are skipped (calcRightHandSide.hpp:541-549)
  double s1 = 0.0, s2 = 0.0;
  double ucell[NV];
  if (a.doSource || a.fusedP2C) {
    double v[NV];
#pragma unroll
    for (int nv = 0; nv < NV; nv++) v[nv] = vcAt(nv);
    if (a.doSource) geometricSource<P>(g, v, eos, dt, j, i, s1, s2);
    // fused path: the conservative state of the cell is rebuilt from the primitives
    // (Fluid::ConvertPrimToCons) instead of being read back, and optionally saved as the
    // "begin" state of the Runge-Kutta cycle (StateContainer::CopyFrom)
    if (a.fusedP2C) primToCons<P>(ucell, v, eos);
  }
#pragma unroll
  for (int nv = 0; nv < NV; nv++) {
    if (P::mhd && nv >= P::BX1 && nv < P::BX1 + P::ndim) continue;
    double *up = a.Uc + nv * g.sn + cell;
    if (a.fusedP2C || a.doSource) {
      const double u0 = a.fusedP2C ? ucell[nv] : *up;
      if (a.fusedP2C && a.Uc0 != nullptr) a.Uc0[nv * g.sn + cell] = u0;
      double u = u0 + rhs[nv];
      if (a.doSource) {
        if (nv == P::MX1) u += s1;
        if (P::ncomp >= 2 && nv == P::MX2) u += s2;
      }
      *up = u;
    } else {
      // Uc += rhs: reduction in L2 (RED.ADD.F64), bit-identical to load-add-store since no other thread
      // touches this cell during the sweep
      atomicAdd(up, rhs[nv]);
    }
  }
}

#endif  // IDFX_SWEEP_CUH_
