// riemann_hd.cuh -- hydrodynamic approximate Riemann solvers acting on one interface.
//
// Reference: src/fluid/RiemannSolver/HDsolvers/tvdlfHD.hpp:20-99, hllHD.hpp:20-117,
// hllcHD.hpp:20-166.  Inputs are the reconstructed left/right primitive states; the
// output is the interface flux and the fastest signal speed (cMax of the reference).
#ifndef IDFX_RIEMANN_HD_CUH_
#define IDFX_RIEMANN_HD_CUH_

#include "fastmath.cuh"
#include "phys.cuh"

struct RiemannAux {
  double cmax;   // fastest signal speed at this interface
  double sl, sr; // outermost wave speeds (used by the UCT-HLL / UCT-HLLD EMF coefficients)
  double c2iso;  // squared isothermal sound speed handed to the EMF storage (0 when adiabatic)
};

template <class P, class VL, class VR>
IDFX_D void soundSpeedsHD(const VL &vL, const VR &vR, const Eos &eos, double &cL, double &cR) {
  if constexpr (P::energy) {
    cL = idfxSqrt<P::fast>(eos.gamma * idfxDiv<P::fast>(vL[P::PRS], vL[P::RHO]));
    cR = idfxSqrt<P::fast>(eos.gamma * idfxDiv<P::fast>(vR[P::PRS], vR[P::RHO]));
  } else {
    cL = 0.5 * (eos.csiso + eos.csiso);
    cR = cL;
  }
}

// tvdlfHD.hpp:49-93
template <class P, int DIR, class VL, class VR>
IDFX_D void riemannTvdlfHD(const VL &vL, const VR &vR, const Eos &eos, double *F, RiemannAux &aux) {
  constexpr int Xn = Dirs<P, DIR>::Xn;
  double uL[P::nvar], uR[P::nvar], fL[P::nvar], fR[P::nvar];
  double cRL;
  const double vnRL = 0.5 * (vL[Xn] + vR[Xn]);
  if constexpr (P::energy) {
    const double pRL = 0.5 * (vL[P::PRS] + vR[P::PRS]);
    const double rRL = 0.5 * (vL[P::RHO] + vR[P::RHO]);
    cRL = idfxSqrt<P::fast>(eos.gamma * idfxDiv<P::fast>(pRL, rRL));
  } else {
    cRL = 0.5 * (eos.csiso + eos.csiso);
  }
  const double cmax = fmax(fabs(vnRL + cRL), fabs(vnRL - cRL));
  primToCons<P>(uL, vL, eos);
  primToCons<P>(uR, vR, eos);
  physFlux<P, DIR>(fL, vL, uL, cRL * cRL);
  physFlux<P, DIR>(fR, vR, uR, cRL * cRL);
#pragma unroll
  for (int nv = 0; nv < P::nvar; nv++) F[nv] = 0.5 * (fL[nv] + fR[nv] - cmax * (uR[nv] - uL[nv]));
  aux.cmax = cmax;
  aux.sl = -cmax;
  aux.sr = cmax;
  aux.c2iso = 0.0;
}

// hllHD.hpp:48-110
template <class P, int DIR, class VL, class VR>
IDFX_D void riemannHllHD(const VL &vL, const VR &vR, const Eos &eos, double *F, RiemannAux &aux) {
  constexpr int Xn = Dirs<P, DIR>::Xn;
  double uL[P::nvar], uR[P::nvar], fL[P::nvar], fR[P::nvar];
  double cL, cR;
  soundSpeedsHD<P>(vL, vR, eos, cL, cR);
  const double SL = fmin(vL[Xn] - cL, vR[Xn] - cR);
  const double SR = fmax(vL[Xn] + cL, vR[Xn] + cR);
  aux.cmax = fmax(fabs(SL), fabs(SR));
  aux.sl = SL;
  aux.sr = SR;
  aux.c2iso = 0.0;
  primToCons<P>(uL, vL, eos);
  primToCons<P>(uR, vR, eos);
  physFlux<P, DIR>(fL, vL, uL, cL * cL);
  physFlux<P, DIR>(fR, vR, uR, cR * cR);
  if (SL > 0) {
#pragma unroll
    for (int nv = 0; nv < P::nvar; nv++) F[nv] = fL[nv];
  } else if (SR < 0) {
#pragma unroll
    for (int nv = 0; nv < P::nvar; nv++) F[nv] = fR[nv];
  } else {
#pragma unroll
    for (int nv = 0; nv < P::nvar; nv++) {
      double f = SL * SR * uR[nv] - SL * SR * uL[nv] + SR * fL[nv] - SL * fR[nv];
      F[nv] = idfxDiv<P::fast>(f, (SR - SL));
    }
  }
}

// hllcHD.hpp:40-160
template <class P, int DIR, class VL, class VR>
IDFX_D void riemannHllcHD(const VL &vL, const VR &vR, const Eos &eos, double *F, RiemannAux &aux) {
  using D = Dirs<P, DIR>;
  constexpr int Xn = D::Xn;
  double uL[P::nvar], uR[P::nvar], fL[P::nvar], fR[P::nvar];
  double cL, cR;
  soundSpeedsHD<P>(vL, vR, eos, cL, cR);
  const double SL = fmin(vL[Xn] - cL, vR[Xn] - cR);
  const double SR = fmax(vL[Xn] + cL, vR[Xn] + cR);
  aux.cmax = fmax(fabs(SL), fabs(SR));
  aux.sl = SL;
  aux.sr = SR;
  aux.c2iso = 0.0;
  primToCons<P>(uL, vL, eos);
  primToCons<P>(uR, vR, eos);
  physFlux<P, DIR>(fL, vL, uL, cL * cL);
  physFlux<P, DIR>(fR, vR, uR, cR * cR);
  if (SL > 0) {
#pragma unroll
    for (int nv = 0; nv < P::nvar; nv++) F[nv] = fL[nv];
  } else if (SR < 0) {
#pragma unroll
    for (int nv = 0; nv < P::nvar; nv++) F[nv] = fR[nv];
  } else {
    double usL[P::nvar], usR[P::nvar];
    double vs;
    if constexpr (P::energy) {
      const double qL = vL[P::PRS] + uL[Xn] * (vL[Xn] - SL);
      const double qR = vR[P::PRS] + uR[Xn] * (vR[Xn] - SR);
      const double wL = vL[P::RHO] * (vL[Xn] - SL);
      const double wR = vR[P::RHO] * (vR[Xn] - SR);
      vs = idfxDiv<P::fast>((qR - qL), (wR - wL));
      usL[P::RHO] = idfxDiv<P::fast>(uL[P::RHO] * (SL - vL[Xn]), (SL - vs));
      usR[P::RHO] = idfxDiv<P::fast>(uR[P::RHO] * (SR - vR[Xn]), (SR - vs));
      usL[Xn] = usL[P::RHO] * vs;
      usR[Xn] = usR[P::RHO] * vs;
      if constexpr (P::ncomp >= 2) {
        usL[D::Xt] = usL[P::RHO] * vL[D::Xt];
        usR[D::Xt] = usR[P::RHO] * vR[D::Xt];
      }
      if constexpr (P::ncomp == 3) {
        usL[D::Xb] = usL[P::RHO] * vL[D::Xb];
        usR[D::Xb] = usR[P::RHO] * vR[D::Xb];
      }
      usL[P::ENG] = idfxDiv<P::fast>(uL[P::ENG], vL[P::RHO]) + (vs - vL[Xn]) * (vs + idfxDiv<P::fast>(vL[P::PRS], (vL[P::RHO] * (SL - vL[Xn]))));
      usR[P::ENG] = idfxDiv<P::fast>(uR[P::ENG], vR[P::RHO]) + (vs - vR[Xn]) * (vs + idfxDiv<P::fast>(vR[P::PRS], (vR[P::RHO] * (SR - vR[Xn]))));
      usL[P::ENG] *= usL[P::RHO];
      usR[P::ENG] *= usR[P::RHO];
    } else {
      const double scrh = idfxDiv<P::fast>(1.0, (SR - SL));
      const double rho = (SR * uR[P::RHO] - SL * uL[P::RHO] - fR[P::RHO] + fL[P::RHO]) * scrh;
      const double mx = (SR * uR[Xn] - SL * uL[Xn] - fR[Xn] + fL[Xn]) * scrh;
      usL[P::RHO] = usR[P::RHO] = rho;
      usL[Xn] = usR[Xn] = mx;
      vs = (SR * fL[P::RHO] - SL * fR[P::RHO] + SR * SL * (uR[P::RHO] - uL[P::RHO]));
      vs *= scrh;
      vs = idfxDiv<P::fast>(vs, rho);
      if constexpr (P::ncomp >= 2) {
        usL[D::Xt] = rho * vL[D::Xt];
        usR[D::Xt] = rho * vR[D::Xt];
      }
      if constexpr (P::ncomp == 3) {
        usL[D::Xb] = rho * vL[D::Xb];
        usR[D::Xb] = rho * vR[D::Xb];
      }
    }
    if (vs >= 0.0) {
#pragma unroll
      for (int nv = 0; nv < P::nvar; nv++) F[nv] = fL[nv] + SL * (usL[nv] - uL[nv]);
    } else {
#pragma unroll
      for (int nv = 0; nv < P::nvar; nv++) F[nv] = fR[nv] + SR * (usR[nv] - uR[nv]);
    }
  }
}

#endif  // IDFX_RIEMANN_HD_CUH_
