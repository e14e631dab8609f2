// emf_store.cuh -- what each interface hands over to constrained transport.
//
// Reference: src/fluid/RiemannSolver/MHDsolvers/storeFlux.hpp:15-208 and the
// per-direction array/sign mapping of hlldMHD.hpp:78-145.
#ifndef IDFX_EMF_STORE_CUH_
#define IDFX_EMF_STORE_CUH_

#include "fastmath.cuh"
#include "phys.cuh"
#include "riemann_hd.cuh"
#include "riemann_mhd.cuh"

// face arrays written by a sweep along DIR
struct EmfFaceArrays {
  double *Et, *Eb, *SV, *aL, *aR, *dL, *dR;
  double st, sb;
};

template <int NDIM, int DIR>
__host__ __device__ inline EmfFaceArrays emfFaceArrays(const DevEmf &e) {
  EmfFaceArrays a;
  a.st = 1.0;
  a.sb = 1.0;
  if (DIR == IDIR) {
    a.st = -1.0;
    if (NDIM == 3) a.sb = +1.0;
    a.Et = e.ezi; a.Eb = e.eyi; a.SV = e.svx;
    a.aL = e.axL; a.aR = e.axR; a.dL = e.dxL; a.dR = e.dxR;
  } else if (DIR == JDIR) {
    a.st = +1.0;
    if (NDIM == 3) a.sb = -1.0;
    a.Et = e.ezj; a.Eb = e.exj; a.SV = e.svy;
    a.aL = e.ayL; a.aR = e.ayR; a.dL = e.dyL; a.dR = e.dyR;
  } else {
    a.st = -1.0;
    if (NDIM == 3) a.sb = +1.0;
    a.Et = e.eyk; a.Eb = e.exk; a.SV = e.svz;
    a.aL = e.azL; a.aR = e.azR; a.dL = e.dzL; a.dR = e.dzR;
  }
  return a;
}

// storeFlux.hpp:15-31 (arithmetic, uct0) and :33-44 (uct_contact)
template <class P, int DIR>
IDFX_D void storeFluxEmf(const EmfFaceArrays &A, long long idx, const double *F, bool contact) {
  using D = Dirs<P, DIR>;
  if constexpr (P::ncomp > 1) {
    A.Et[idx] = A.st * F[D::BXt];
    if constexpr (P::ndim == 3) A.Eb[idx] = A.sb * F[D::BXb];
  }
  if (contact) {
    double s = 0.5;
    if (F[P::RHO] > EPS_UCT_CONTACT) s = 1.0;
    if (F[P::RHO] < -EPS_UCT_CONTACT) s = 0.0;
    A.SV[idx] = s;
  }
}

// storeFlux.hpp:46-75
template <class P, int DIR, class VL, class VR>
IDFX_D void storeHllEmf(const EmfFaceArrays &A, long long idx, double sl, double sr, const VL &vL,
                        const VR &vR) {
  using D = Dirs<P, DIR>;
  const double ar = fmax(0.0, sr);
  const double al = fmin(0.0, sl);
  const double scrh = idfxDiv<P::fast>(1.0, (ar - al));
  if constexpr (P::ncomp > 1) {
    A.Et[idx] = -A.st * (ar * vL[D::Xt] - al * vR[D::Xt]) * scrh;
    if constexpr (P::ncomp == 3) A.Eb[idx] = -A.sb * (ar * vL[D::Xb] - al * vR[D::Xb]) * scrh;
  }
  A.aL[idx] = ar * scrh;
  A.aR[idx] = -al * scrh;
  const double d = -al * ar * scrh;
  A.dR[idx] = d;
  A.dL[idx] = d;
}

// storeFlux.hpp:77-208
template <class P, int DIR, class VL, class VR>
IDFX_D void storeHlldEmf(const EmfFaceArrays &A, long long idx, double c2iso, double sl, double sr,
                         const VL &vL, const VR &vR) {
  using D = Dirs<P, DIR>;
  constexpr int Xn = D::Xn, BXn = D::BXn, RHO = P::RHO;
  const double ptL = totalPressure<P>(vL, c2iso);
  const double ptR = totalPressure<P>(vR, c2iso);
  // conservative density / normal momentum of the two states (K_PrimToCons)
  const double uLrho = vL[RHO], uRrho = vR[RHO];
  const double uLmn = vL[Xn] * vL[RHO], uRmn = vR[Xn] * vR[RHO];

  bool revert_to_hll = false;
  const double Bn = idfxDiv<P::fast>((sr * vR[BXn] - sl * vL[BXn]), (sr - sl));
  double chiL, chiR, nuLR, nuL, nuR, SaL, SaR, Sc;
  const double duL = sl - vL[Xn];
  const double duR = sr - vR[Xn];
  double scrh;
  if constexpr (P::energy) {
    scrh = idfxDiv<P::fast>(1.0, (duR * uRrho - duL * uLrho));
    Sc = (duR * uRmn - duL * uLmn - ptR + ptL) * scrh;
    const double usLRHO = idfxDiv<P::fast>(uLrho * duL, (sl - Sc));
    const double usRRHO = idfxDiv<P::fast>(uRrho * duR, (sr - Sc));
    const double sqrL = idfxSqrt<P::fast>(usLRHO);
    const double sqrR = idfxSqrt<P::fast>(usRRHO);
    SaL = Sc - idfxDiv<P::fast>(fabs(Bn), sqrL);
    SaR = Sc + idfxDiv<P::fast>(fabs(Bn), sqrR);
  } else {
    scrh = idfxDiv<P::fast>(1.0, (sr - sl));
    const double rho_h = (uRrho * duR - uLrho * duL) * scrh;
    Sc = idfxDiv<P::fast>((sl * uRrho * duR - sr * uLrho * duL) * scrh, rho_h);
    const double sqrho_h = idfxSqrt<P::fast>(rho_h);
    SaL = Sc - idfxDiv<P::fast>(fabs(Bn), sqrho_h);
    SaR = Sc + idfxDiv<P::fast>(fabs(Bn), sqrho_h);
  }
  if ((SaL - sl) < 1e-4 * (Sc - sl)) revert_to_hll = true;
  if ((SaR - sr) > -1e-4 * (sr - Sc)) revert_to_hll = true;
  chiL = idfxDiv<P::fast>((vL[Xn] - Sc) * (sl - Sc), (SaL + sl - 2.0 * Sc));
  chiR = idfxDiv<P::fast>((vR[Xn] - Sc) * (sr - Sc), (SaR + sr - 2.0 * Sc));

  nuL = idfxDiv<P::fast>((SaL + sl), (fabs(SaL) + fabs(sl)));
  nuR = idfxDiv<P::fast>((SaR + sr), (fabs(SaR) + fabs(sr)));
  nuLR = idfxDiv<P::fast>((SaL + SaR), (fabs(SaL) + fabs(SaR)));

  double oaL, oaR, odL, odR;
  if (fabs(SaR - SaL) > 1.e-9 * fabs(sr - sl)) {
    odL = 0.5 * (chiL * nuL - chiL * nuLR) + 0.5 * (fabs(SaL) - nuLR * SaL);
    odR = 0.5 * (chiR * nuR - chiR * nuLR) + 0.5 * (fabs(SaR) - nuLR * SaR);
    oaL = 0.5 * (1.0 + nuLR);
    oaR = 0.5 * (1.0 - nuLR);
  } else {  // degenerate limit Bx -> 0
    odL = 0.5 * chiL * nuL + 0.5 * fabs(SaL);
    odR = 0.5 * chiR * nuR + 0.5 * fabs(SaR);
    oaL = 0.5;
    oaR = 0.5;
  }
  const double ar = fmax(0.0, sr);
  const double al = fmin(0.0, sl);
  scrh = idfxDiv<P::fast>(1.0, (ar - al));
  if (revert_to_hll) {
    oaL = ar * scrh;
    oaR = -al * scrh;
    odR = -al * ar * scrh;
    odL = odR;
  }
  A.aL[idx] = oaL;
  A.aR[idx] = oaR;
  A.dL[idx] = odL;
  A.dR[idx] = odR;
  if constexpr (P::ncomp > 1) {
    A.Et[idx] = -A.st * (ar * vL[D::Xt] - al * vR[D::Xt]) * scrh;
    if constexpr (P::ncomp == 3) A.Eb[idx] = -A.sb * (ar * vL[D::Xb] - al * vR[D::Xb]) * scrh;
  }
}

template <class P, int DIR, int SOLVER, class VL, class VR>
IDFX_D void storeEmf(int mode, const EmfFaceArrays &A, long long idx, const double *F, const RiemannAux &aux,
                     const VL &vL, const VR &vR) {
  if (mode == EMF_ARITHMETIC || mode == EMF_UCT0) {
    storeFluxEmf<P, DIR>(A, idx, F, false);
  } else if (mode == EMF_UCT_CONTACT) {
    storeFluxEmf<P, DIR>(A, idx, F, true);
  } else if (mode == EMF_UCT_HLL) {
    storeHllEmf<P, DIR>(A, idx, aux.sl, aux.sr, vL, vR);
  } else if (mode == EMF_UCT_HLLD) {
    // only HLLD and Roe provide what this scheme needs (constrainedTransport.hpp:254-259)
    if constexpr (SOLVER == SOLVER_HLLD || SOLVER == SOLVER_ROE) {
      storeHlldEmf<P, DIR>(A, idx, aux.c2iso, aux.sl, aux.sr, vL, vR);
    }
  }
}

#endif  // IDFX_EMF_STORE_CUH_
