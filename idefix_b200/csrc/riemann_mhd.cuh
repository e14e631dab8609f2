// riemann_mhd.cuh -- MHD approximate Riemann solvers (TVDLF, HLL, HLLD) on one interface.
//
// Reference: src/fluid/RiemannSolver/MHDsolvers/tvdlfMHD.hpp:147-225,
// hllMHD.hpp:155-445 (Hall terms are out of scope), hlldMHD.hpp:147-575.
// vL/vR already carry the face-centred normal field in BXn (hlldMHD.hpp:162-164).
#ifndef IDFX_RIEMANN_MHD_CUH_
#define IDFX_RIEMANN_MHD_CUH_

#include "fastmath.cuh"
#include "phys.cuh"
#include "riemann_hd.cuh"

// fast magnetosonic speed of one state (hlldMHD.hpp:191-207)
template <class P, int DIR, class V>
IDFX_D double fastSpeed(const V &v, double gpr) {
  using D = Dirs<P, DIR>;
  double b1 = v[D::BXn];
  double b2 = 0.0, b3 = 0.0;
  if constexpr (P::ncomp >= 2) b2 = v[D::BXt];
  if constexpr (P::ncomp == 3) b3 = v[D::BXb];
  const double Btmag2 = b2 * b2 + b3 * b3;
  const double Bmag2 = b1 * b1 + Btmag2;
  double c = gpr - Bmag2;
  c = gpr + Bmag2 + idfxSqrt<P::fast>(c * c + 4.0 * gpr * Btmag2);
  return idfxSqrt<P::fast>(idfxDiv<P::fast>(0.5 * c, v[P::RHO]));
}

template <class P, class V>
IDFX_D double totalPressure(const V &v, double c2iso) {
  double b2 = v[P::BX1] * v[P::BX1];
  if constexpr (P::ncomp >= 2) b2 = b2 + v[P::BX2] * v[P::BX2];
  if constexpr (P::ncomp == 3) b2 = b2 + v[P::BX3] * v[P::BX3];
  if constexpr (P::energy) {
    return v[P::PRS] + 0.5 * b2;
  } else {
    return c2iso * v[P::RHO] + 0.5 * b2;
  }
}

// tvdlfMHD.hpp:147-210
template <class P, int DIR, class VL, class VR>
IDFX_D void riemannTvdlfMHD(const VL &vL, const VR &vR, const Eos &eos, double *F, RiemannAux &aux) {
  using D = Dirs<P, DIR>;
  constexpr int Xn = D::Xn;
  double uL[P::nvar], uR[P::nvar], fL[P::nvar], fR[P::nvar];
  double c2iso = 0.0;
  double gpr;
  const double rho = 0.5 * (vL[P::RHO] + vR[P::RHO]);
  const double vn = 0.5 * (vL[Xn] + vR[Xn]);
  if constexpr (P::energy) {
    gpr = eos.gamma * (0.5 * (vL[P::PRS] + vR[P::PRS]));
  } else {
    c2iso = 0.5 * (eos.csiso + eos.csiso);
    c2iso *= c2iso;
    gpr = c2iso * rho;
  }
  double Bt2 = 0.0;
  if constexpr (P::ncomp >= 2) {
    const double bt = 0.5 * (vL[D::BXt] + vR[D::BXt]);
    Bt2 = Bt2 + bt * bt;
  }
  if constexpr (P::ncomp == 3) {
    const double bb = 0.5 * (vL[D::BXb] + vR[D::BXb]);
    Bt2 = Bt2 + bb * bb;
  }
  const double bn = 0.5 * (vL[D::BXn] + vR[D::BXn]);
  const double B2 = Bt2 + bn * bn;
  // note: the reference re-uses cRL here, which cancels the +B2 (tvdlfMHD.hpp:192-194)
  double cRL = gpr - B2;
  cRL = cRL + B2 + idfxSqrt<P::fast>(cRL * cRL + 4.0 * gpr * Bt2);
  cRL = idfxSqrt<P::fast>(idfxDiv<P::fast>(0.5 * cRL, rho));
  const double cmax = fmax(fabs(vn + cRL), fabs(vn - cRL));
  primToCons<P>(uL, vL, eos);
  primToCons<P>(uR, vR, eos);
  physFlux<P, DIR>(fL, vL, uL, c2iso);
  physFlux<P, DIR>(fR, vR, uR, c2iso);
#pragma unroll
  for (int nv = 0; nv < P::nvar; nv++) F[nv] = 0.5 * (fL[nv] + fR[nv] + cmax * (uL[nv] - uR[nv]));
  aux.cmax = cmax;
  aux.sl = -cmax;
  aux.sr = cmax;
  aux.c2iso = c2iso;
}

// hllMHD.hpp:155-445 without Hall
template <class P, int DIR, class VL, class VR>
IDFX_D void riemannHllMHD(const VL &vL, const VR &vR, const Eos &eos, double *F, RiemannAux &aux) {
  using D = Dirs<P, DIR>;
  constexpr int Xn = D::Xn;
  double uL[P::nvar], uR[P::nvar], fL[P::nvar], fR[P::nvar];
  double c2iso = 0.0;
  double gprL, gprR;
  if constexpr (P::energy) {
    gprL = eos.gamma * vL[P::PRS];
    gprR = eos.gamma * vR[P::PRS];
  } else {
    c2iso = 0.5 * (eos.csiso + eos.csiso);
    c2iso *= c2iso;
    gprL = c2iso * vL[P::RHO];
    gprR = c2iso * vR[P::RHO];
  }
  const double cL = fastSpeed<P, DIR>(vL, gprL);
  const double cR = fastSpeed<P, DIR>(vR, gprR);
  const double sl = fmin(vL[Xn] - cL, vR[Xn] - cR);
  const double sr = fmax(vL[Xn] + cL, vR[Xn] + cR);
  aux.cmax = fmax(fabs(sl), fabs(sr));
  aux.sl = sl;
  aux.sr = sr;
  aux.c2iso = c2iso;
  primToCons<P>(uL, vL, eos);
  primToCons<P>(uR, vR, eos);
#pragma unroll
  for (int nv = 0; nv < P::nvar; nv++) {
    fL[nv] = uL[nv];
    fR[nv] = uR[nv];
  }
  physFlux<P, DIR>(fL, vL, fL, c2iso);
  physFlux<P, DIR>(fR, vR, fR, c2iso);
  // density and momenta divide by (sr-sl); field and energy multiply by its reciprocal (hllMHD.hpp:395-433)
  if (sl > 0) {
#pragma unroll
    for (int nv = 0; nv < P::nvar; nv++) F[nv] = fL[nv];
  } else if (sr < 0) {
#pragma unroll
    for (int nv = 0; nv < P::nvar; nv++) F[nv] = fR[nv];
  } else {
    const double rcp = idfxDiv<P::fast>(1.0, (sr - sl));
#pragma unroll
    for (int nv = 0; nv < P::nvar; nv++) {
      const double f = sl * sr * uR[nv] - sl * sr * uL[nv] + sr * fL[nv] - sl * fR[nv];
      if (nv <= P::ncomp) {
        F[nv] = idfxDiv<P::fast>(f, (sr - sl));
      } else {
        F[nv] = f * rcp;
      }
    }
  }
}

// hlldMHD.hpp:147-575
template <class P, int DIR, class VL, class VR>
IDFX_D void riemannHlldMHD(const VL &vL, const VR &vR, const Eos &eos, double *F, RiemannAux &aux) {
  using D = Dirs<P, DIR>;
  constexpr int Xn = D::Xn, Xt = D::Xt, Xb = D::Xb;
  constexpr int BXn = D::BXn, BXt = D::BXt, BXb = D::BXb;
  constexpr int RHO = P::RHO;
  double uL[P::nvar], uR[P::nvar], fL[P::nvar], fR[P::nvar];
  double c2iso = 0.0;
  double gprL, gprR;
  if constexpr (P::energy) {
    gprL = eos.gamma * vL[P::PRS];
    gprR = eos.gamma * vR[P::PRS];
  } else {
    c2iso = 0.5 * (eos.csiso + eos.csiso);
    c2iso *= c2iso;
    gprL = c2iso * vL[RHO];
    gprR = c2iso * vR[RHO];
  }
  const double cL = fastSpeed<P, DIR>(vL, gprL);
  const double cR = fastSpeed<P, DIR>(vR, gprR);
  const double sl = fmin(vL[Xn] - cL, vR[Xn] - cR);
  const double sr = fmax(vL[Xn] + cL, vR[Xn] + cR);
  aux.cmax = fmax(fabs(sl), fabs(sr));
  aux.sl = sl;
  aux.sr = sr;
  aux.c2iso = c2iso;

  primToCons<P>(uL, vL, eos);
  primToCons<P>(uR, vR, eos);
#pragma unroll
  for (int nv = 0; nv < P::nvar; nv++) {
    fL[nv] = uL[nv];
    fR[nv] = uR[nv];
  }
  physFlux<P, DIR>(fL, vL, fL, c2iso);
  physFlux<P, DIR>(fR, vR, fR, c2iso);

  if (sl > 0) {
#pragma unroll
    for (int nv = 0; nv < P::nvar; nv++) F[nv] = fL[nv];
    return;
  }
  if (sr < 0) {
#pragma unroll
    for (int nv = 0; nv < P::nvar; nv++) F[nv] = fR[nv];
    return;
  }

  double usL[P::nvar], usR[P::nvar];
  double scrh, scrhL, scrhR, duL, duR, sBx, Bx, SM, S1L, S1R;

  if constexpr (P::energy) {
    constexpr int ENG = P::ENG;
    const double ptL = totalPressure<P>(vL, 0.0);
    const double ptR = totalPressure<P>(vR, 0.0);
    [[maybe_unused]] double vsL = 0.0, vsR = 0.0, wsL = 0.0, wsR = 0.0;
    bool revert_to_hllc = false;

    scrh = idfxDiv<P::fast>(1.0, (sr - sl));
    Bx = (sr * vR[BXn] - sl * vL[BXn]) * scrh;
    sBx = (Bx > 0.0 ? 1.0 : -1.0);

    duL = sl - vL[Xn];
    duR = sr - vR[Xn];

    scrh = idfxDiv<P::fast>(1.0, (duR * uR[RHO] - duL * uL[RHO]));
    SM = (duR * uR[Xn] - duL * uL[Xn] - ptR + ptL) * scrh;

    double pts = duR * uR[RHO] * ptL - duL * uL[RHO] * ptR + vL[RHO] * vR[RHO] * duR * duL * (vR[Xn] - vL[Xn]);
    pts *= scrh;

    usL[RHO] = idfxDiv<P::fast>(uL[RHO] * duL, (sl - SM));
    usR[RHO] = idfxDiv<P::fast>(uR[RHO] * duR, (sr - SM));

    const double sqrL = idfxSqrt<P::fast>(usL[RHO]);
    const double sqrR = idfxSqrt<P::fast>(usR[RHO]);

    S1L = SM - idfxDiv<P::fast>(fabs(Bx), sqrL);
    S1R = SM + idfxDiv<P::fast>(fabs(Bx), sqrR);

    // degeneracy: fall back on HLLC-type star states (hlldMHD.hpp:311-340)
    if ((S1L - sl) < 1.e-4 * (SM - sl)) revert_to_hllc = true;
    if ((S1R - sr) > -1.e-4 * (sr - SM)) revert_to_hllc = true;

    if (revert_to_hllc) {
      scrh = idfxDiv<P::fast>(1.0, (sr - sl));
      // only the field components of the HLL state are needed
      {
        double h = sr * uR[BXn] - sl * uL[BXn] + fL[BXn] - fR[BXn];
        h *= scrh;
        usL[BXn] = usR[BXn] = h;
      }
      if constexpr (P::ncomp >= 2) {
        double h = sr * uR[BXt] - sl * uL[BXt] + fL[BXt] - fR[BXt];
        h *= scrh;
        usL[BXt] = usR[BXt] = h;
      }
      if constexpr (P::ncomp == 3) {
        double h = sr * uR[BXb] - sl * uL[BXb] + fL[BXb] - fR[BXb];
        h *= scrh;
        usL[BXb] = usR[BXb] = h;
      }
      S1L = S1R = SM;
    } else {
      scrhL = idfxDiv<P::fast>((uL[RHO] * duL * duL - Bx * Bx), (uL[RHO] * duL * (sl - SM) - Bx * Bx));
      scrhR = idfxDiv<P::fast>((uR[RHO] * duR * duR - Bx * Bx), (uR[RHO] * duR * (sr - SM) - Bx * Bx));
      usL[BXn] = Bx;
      usR[BXn] = Bx;
      if constexpr (P::ncomp >= 2) {
        usL[BXt] = uL[BXt] * scrhL;
        usR[BXt] = uR[BXt] * scrhR;
      }
      if constexpr (P::ncomp == 3) {
        usL[BXb] = uL[BXb] * scrhL;
        usR[BXb] = uR[BXb] * scrhR;
      }
    }

    scrhL = idfxDiv<P::fast>(Bx, (uL[RHO] * duL));
    scrhR = idfxDiv<P::fast>(Bx, (uR[RHO] * duR));

    if constexpr (P::ncomp >= 2) {
      vsL = vL[Xt] - scrhL * (usL[BXt] - uL[BXt]);
      vsR = vR[Xt] - scrhR * (usR[BXt] - uR[BXt]);
    }
    if constexpr (P::ncomp == 3) {
      wsL = vL[Xb] - scrhL * (usL[BXb] - uL[BXb]);
      wsR = vR[Xb] - scrhR * (usR[BXb] - uR[BXb]);
    }

    usL[Xn] = usL[RHO] * SM;
    usR[Xn] = usR[RHO] * SM;
    if constexpr (P::ncomp >= 2) {
      usL[Xt] = usL[RHO] * vsL;
      usR[Xt] = usR[RHO] * vsR;
    }
    if constexpr (P::ncomp == 3) {
      usL[Xb] = usL[RHO] * wsL;
      usR[Xb] = usR[RHO] * wsR;
    }

    // energy of the star states
    scrhL = vL[Xn] * Bx;
    if constexpr (P::ncomp >= 2) scrhL = scrhL + vL[Xt] * uL[BXt];
    if constexpr (P::ncomp == 3) scrhL = scrhL + vL[Xb] * uL[BXb];
    {
      double s = SM * Bx;
      if constexpr (P::ncomp >= 2) s = s + vsL * usL[BXt];
      if constexpr (P::ncomp == 3) s = s + wsL * usL[BXb];
      scrhL -= s;
    }
    usL[ENG] = duL * uL[ENG] - ptL * vL[Xn] + pts * SM + Bx * scrhL;
    usL[ENG] = idfxDiv<P::fast>(usL[ENG], sl - SM);

    scrhR = vR[Xn] * Bx;
    if constexpr (P::ncomp >= 2) scrhR = scrhR + vR[Xt] * uR[BXt];
    if constexpr (P::ncomp == 3) scrhR = scrhR + vR[Xb] * uR[BXb];
    {
      double s = SM * Bx;
      if constexpr (P::ncomp >= 2) s = s + vsR * usR[BXt];
      if constexpr (P::ncomp == 3) s = s + wsR * usR[BXb];
      scrhR -= s;
    }
    usR[ENG] = duR * uR[ENG] - ptR * vR[Xn] + pts * SM + Bx * scrhR;
    usR[ENG] = idfxDiv<P::fast>(usR[ENG], sr - SM);

    if (S1L >= 0.0) {  // region L*
#pragma unroll
      for (int nv = 0; nv < P::nvar; nv++) F[nv] = fL[nv] + sl * (usL[nv] - uL[nv]);
    } else if (S1R <= 0.0) {  // region R*
#pragma unroll
      for (int nv = 0; nv < P::nvar; nv++) F[nv] = fR[nv] + sr * (usR[nv] - uR[nv]);
    } else {  // double-star regions, exist only when Bx != 0
      [[maybe_unused]] double vss = 0.0, wss = 0.0;
      double ussl[P::nvar], ussr[P::nvar];
      ussl[RHO] = usL[RHO];
      ussr[RHO] = usR[RHO];
      if constexpr (P::ncomp >= 2) {
        vss = sqrL * vsL + sqrR * vsR + (usR[BXt] - usL[BXt]) * sBx;
        vss = idfxDiv<P::fast>(vss, sqrL + sqrR);
      }
      if constexpr (P::ncomp == 3) {
        wss = sqrL * wsL + sqrR * wsR + (usR[BXb] - usL[BXb]) * sBx;
        wss = idfxDiv<P::fast>(wss, sqrL + sqrR);
      }
      ussl[Xn] = ussl[RHO] * SM;
      ussr[Xn] = ussr[RHO] * SM;
      if constexpr (P::ncomp >= 2) {
        ussl[Xt] = ussl[RHO] * vss;
        ussr[Xt] = ussr[RHO] * vss;
      }
      if constexpr (P::ncomp == 3) {
        ussl[Xb] = ussl[RHO] * wss;
        ussr[Xb] = ussr[RHO] * wss;
      }
      ussl[BXn] = ussr[BXn] = Bx;
      if constexpr (P::ncomp >= 2) {
        ussl[BXt] = sqrL * usR[BXt] + sqrR * usL[BXt] + sqrL * sqrR * (vsR - vsL) * sBx;
        ussl[BXt] = idfxDiv<P::fast>(ussl[BXt], sqrL + sqrR);
        ussr[BXt] = ussl[BXt];
      }
      if constexpr (P::ncomp == 3) {
        ussl[BXb] = sqrL * usR[BXb] + sqrR * usL[BXb] + sqrL * sqrR * (wsR - wsL) * sBx;
        ussl[BXb] = idfxDiv<P::fast>(ussl[BXb], sqrL + sqrR);
        ussr[BXb] = ussl[BXb];
      }
      // energy jump across the rotational waves
      scrhL = SM * Bx;
      if constexpr (P::ncomp >= 2) scrhL = scrhL + vsL * usL[BXt];
      if constexpr (P::ncomp == 3) scrhL = scrhL + wsL * usL[BXb];
      {
        double s = SM * Bx;
        if constexpr (P::ncomp >= 2) s = s + vss * ussl[BXt];
        if constexpr (P::ncomp == 3) s = s + wss * ussl[BXb];
        scrhL -= s;
      }
      scrhR = SM * Bx;
      if constexpr (P::ncomp >= 2) scrhR = scrhR + vsR * usR[BXt];
      if constexpr (P::ncomp == 3) scrhR = scrhR + wsR * usR[BXb];
      {
        double s = SM * Bx;
        if constexpr (P::ncomp >= 2) s = s + vss * ussr[BXt];
        if constexpr (P::ncomp == 3) s = s + wss * ussr[BXb];
        scrhR -= s;
      }
      ussl[ENG] = usL[ENG] - sqrL * scrhL * sBx;
      ussr[ENG] = usR[ENG] + sqrR * scrhR * sBx;

      if (SM >= 0.0) {  // region L**
#pragma unroll
        for (int nv = 0; nv < P::nvar; nv++)
          F[nv] = fL[nv] + S1L * (ussl[nv] - usL[nv]) + sl * (usL[nv] - uL[nv]);
      } else {  // region R**
#pragma unroll
        for (int nv = 0; nv < P::nvar; nv++)
          F[nv] = fR[nv] + S1R * (ussr[nv] - usR[nv]) + sr * (usR[nv] - uR[nv]);
      }
    }
  } else {
    // isothermal HLLD (hlldMHD.hpp:459-572)
    double usc[P::nvar];
    bool revert_to_hll = false;
    scrh = idfxDiv<P::fast>(1.0, (sr - sl));
    duL = sl - vL[Xn];
    duR = sr - vR[Xn];
    Bx = (sr * vR[BXn] - sl * vL[BXn]) * scrh;
    const double rho = (uR[RHO] * duR - uL[RHO] * duL) * scrh;
    F[RHO] = (sl * uR[RHO] * duR - sr * uL[RHO] * duL) * scrh;
    const double sqrho = idfxSqrt<P::fast>(rho);
    SM = idfxDiv<P::fast>(F[RHO], rho);
    S1L = SM - idfxDiv<P::fast>(fabs(Bx), sqrho);
    S1R = SM + idfxDiv<P::fast>(fabs(Bx), sqrho);
    if ((S1L - sl) < 1.e-4 * (sr - sl)) revert_to_hll = true;
    if ((S1R - sr) > -1.e-4 * (sr - sl)) revert_to_hll = true;
    if (revert_to_hll) {
      scrh = idfxDiv<P::fast>(1.0, (sr - sl));
#pragma unroll
      for (int nv = 0; nv < P::nvar; nv++) {
        F[nv] = sl * sr * (uR[nv] - uL[nv]) + sr * fL[nv] - sl * fR[nv];
        F[nv] *= scrh;
      }
    } else {
      F[Xn] = (sr * fL[Xn] - sl * fR[Xn] + sr * sl * (uR[Xn] - uL[Xn])) * scrh;
      F[BXn] = sr * sl * (uR[BXn] - uL[BXn]) * scrh;
      scrhL = idfxDiv<P::fast>(1.0, ((sl - S1L) * (sl - S1R)));
      scrhR = idfxDiv<P::fast>(1.0, ((sr - S1L) * (sr - S1R)));
      if constexpr (P::ncomp >= 2) {
        usL[Xt] = rho * vL[Xt] - Bx * uL[BXt] * (SM - vL[Xn]) * scrhL;
        usR[Xt] = rho * vR[Xt] - Bx * uR[BXt] * (SM - vR[Xn]) * scrhR;
      }
      if constexpr (P::ncomp == 3) {
        usL[Xb] = rho * vL[Xb] - Bx * uL[BXb] * (SM - vL[Xn]) * scrhL;
        usR[Xb] = rho * vR[Xb] - Bx * uR[BXb] * (SM - vR[Xn]) * scrhR;
      }
      if constexpr (P::ncomp >= 2) {
        usL[BXt] = idfxDiv<P::fast>(uL[BXt], rho) * (uL[RHO] * duL * duL - Bx * Bx) * scrhL;
        usR[BXt] = idfxDiv<P::fast>(uR[BXt], rho) * (uR[RHO] * duR * duR - Bx * Bx) * scrhR;
      }
      if constexpr (P::ncomp == 3) {
        usL[BXb] = idfxDiv<P::fast>(uL[BXb], rho) * (uL[RHO] * duL * duL - Bx * Bx) * scrhL;
        usR[BXb] = idfxDiv<P::fast>(uR[BXb], rho) * (uR[RHO] * duR * duR - Bx * Bx) * scrhR;
      }
      if (S1L >= 0.0) {
        if constexpr (P::ncomp >= 2) {
          F[Xt] = fL[Xt] + sl * (usL[Xt] - uL[Xt]);
          F[BXt] = fL[BXt] + sl * (usL[BXt] - uL[BXt]);
        }
        if constexpr (P::ncomp == 3) {
          F[Xb] = fL[Xb] + sl * (usL[Xb] - uL[Xb]);
          F[BXb] = fL[BXb] + sl * (usL[BXb] - uL[BXb]);
        }
      } else if (S1R <= 0.0) {
        if constexpr (P::ncomp >= 2) {
          F[Xt] = fR[Xt] + sr * (usR[Xt] - uR[Xt]);
          F[BXt] = fR[BXt] + sr * (usR[BXt] - uR[BXt]);
        }
        if constexpr (P::ncomp == 3) {
          F[Xb] = fR[Xb] + sr * (usR[Xb] - uR[Xb]);
          F[BXb] = fR[BXb] + sr * (usR[BXb] - uR[BXb]);
        }
      } else {
        sBx = (Bx > 0.0 ? 1.0 : -1.0);
        if constexpr (P::ncomp >= 2) {
          usc[Xt] = 0.5 * (usR[Xt] + usL[Xt] + (usR[BXt] - usL[BXt]) * sBx * sqrho);
          usc[BXt] = 0.5 * (usR[BXt] + usL[BXt] + idfxDiv<P::fast>((usR[Xt] - usL[Xt]) * sBx, sqrho));
          F[Xt] = usc[Xt] * SM - Bx * usc[BXt];
          F[BXt] = usc[BXt] * SM - idfxDiv<P::fast>(Bx * usc[Xt], rho);
        }
        if constexpr (P::ncomp == 3) {
          usc[Xb] = 0.5 * (usR[Xb] + usL[Xb] + (usR[BXb] - usL[BXb]) * sBx * sqrho);
          usc[BXb] = 0.5 * (usR[BXb] + usL[BXb] + idfxDiv<P::fast>((usR[Xb] - usL[Xb]) * sBx, sqrho));
          F[Xb] = usc[Xb] * SM - Bx * usc[BXb];
          F[BXb] = usc[BXb] * SM - idfxDiv<P::fast>(Bx * usc[Xb], rho);
        }
      }
    }
  }
}

#endif  // IDFX_RIEMANN_MHD_CUH_
