// phys.cuh -- variable indexing, EOS, conservative<->primitive conversion and the
// physical flux, as per-cell register functions.
//
// Reference: src/idefix.hpp:74-108 (variable indices), src/physics.hpp:16-38 (nvar),
// src/fluid/convertConsToPrim.hpp:16-100 (K_ConsToPrim / K_PrimToCons),
// src/fluid/RiemannSolver/flux.hpp:29-95 (K_Flux), src/fluid/eos/eos_adiabatic.hpp:39-45.
// The association order of every floating-point expression follows the reference so
// that results are bit-comparable when compiled with -fmad=false.
#ifndef IDFX_PHYS_CUH_
#define IDFX_PHYS_CUH_

#include "common.cuh"
#include "fastmath.cuh"

#define SMALL_PRESSURE_FIX (1.0e-5)   // src/fluid/fluid_defs.hpp:23
#define EPS_UCT_CONTACT (1.0e-6)      // src/fluid/fluid_defs.hpp:24

// FAST_: divisions and square roots of the interface solve use the branch-free sequences of fastmath.cuh
// (the sweep kernels try Phys<...,true> first and fall back on Phys<...,false> when a range check failed)
template <int ND, int NC, bool MHD_, bool ISO_, bool FAST_ = false>
struct Phys {
  static constexpr bool fast = FAST_;
  using Fast = Phys<ND, NC, MHD_, ISO_, true>;
  using Exact = Phys<ND, NC, MHD_, ISO_, false>;
  static constexpr int ndim = ND;
  static constexpr int ncomp = NC;
  static constexpr bool mhd = MHD_;
  static constexpr bool iso = ISO_;
  static constexpr bool energy = !ISO_;
  static constexpr int RHO = 0;
  static constexpr int MX1 = 1;
  static constexpr int MX2 = (NC >= 2 ? 2 : 255);
  static constexpr int MX3 = (NC == 3 ? 3 : 254);
  static constexpr int ENG = (ISO_ ? 250 : NC + 1);
  static constexpr int PRS = ENG;
  static constexpr int BX1 = (MHD_ ? NC + 1 + (ISO_ ? 0 : 1) : 253);
  static constexpr int BX2 = (MHD_ && NC >= 2 ? BX1 + 1 : 252);
  static constexpr int BX3 = (MHD_ && NC == 3 ? BX1 + 2 : 251);
  static constexpr int nvar = 1 + NC + (ISO_ ? 0 : 1) + (MHD_ ? NC : 0);
  // number of B components carried by constrained transport (first ndim ones)
  static constexpr int nct = (MHD_ ? ND : 0);
};

// per-direction component bookkeeping (normal / tangential / bi-tangential)
template <class P, int DIR>
struct Dirs {
  static constexpr int Xn = DIR + P::MX1;
  static constexpr int Xt = (DIR == IDIR ? P::MX2 : P::MX1);
  static constexpr int Xb = (DIR == KDIR ? P::MX2 : P::MX3);
  static constexpr int BXn = DIR + P::BX1;
  static constexpr int BXt = (DIR == IDIR ? P::BX2 : P::BX1);
  static constexpr int BXb = (DIR == KDIR ? P::BX2 : P::BX3);
};

struct Eos {
  double gamma;
  double csiso;
  IDFX_HD double internalEnergy(double prs) const { return prs / (gamma - 1.0); }
  IDFX_HD double pressure(double eint) const { return eint * (gamma - 1.0); }
};

// convertConsToPrim.hpp:70-100
template <class P, class V>
IDFX_HD void primToCons(double *u, const V &v, const Eos &eos) {
  u[P::RHO] = v[P::RHO];
  u[P::MX1] = v[P::MX1] * v[P::RHO];
  if constexpr (P::ncomp >= 2) u[P::MX2] = v[P::MX2] * v[P::RHO];
  if constexpr (P::ncomp == 3) u[P::MX3] = v[P::MX3] * v[P::RHO];
  if constexpr (P::mhd) {
    u[P::BX1] = v[P::BX1];
    if constexpr (P::ncomp >= 2) u[P::BX2] = v[P::BX2];
    if constexpr (P::ncomp == 3) u[P::BX3] = v[P::BX3];
  }
  if constexpr (P::energy) {
    double v2 = v[P::MX1] * v[P::MX1];
    if constexpr (P::ncomp >= 2) v2 = v2 + v[P::MX2] * v[P::MX2];
    if constexpr (P::ncomp == 3) v2 = v2 + v[P::MX3] * v[P::MX3];
    if constexpr (P::mhd) {
      double b2 = u[P::BX1] * u[P::BX1];
      if constexpr (P::ncomp >= 2) b2 = b2 + u[P::BX2] * u[P::BX2];
      if constexpr (P::ncomp == 3) b2 = b2 + u[P::BX3] * u[P::BX3];
      u[P::ENG] = idfxDiv<P::fast>(v[P::PRS], eos.gamma - 1.0) + 0.5 * v[P::RHO] * v2 + 0.5 * b2;
    } else {
      u[P::ENG] = idfxDiv<P::fast>(v[P::PRS], eos.gamma - 1.0) + 0.5 * v[P::RHO] * v2;
    }
  }
}

// convertConsToPrim.hpp:16-67.  Returns true when the pressure floor fired (u[ENG] rewritten).
template <class P>
IDFX_HD bool consToPrim(double *v, double *u, const Eos &eos) {
  bool fixed = false;
  v[P::RHO] = u[P::RHO];
  v[P::MX1] = u[P::MX1] / u[P::RHO];
  if constexpr (P::ncomp >= 2) v[P::MX2] = u[P::MX2] / u[P::RHO];
  if constexpr (P::ncomp == 3) v[P::MX3] = u[P::MX3] / u[P::RHO];
  if constexpr (P::mhd) {
    v[P::BX1] = u[P::BX1];
    if constexpr (P::ncomp >= 2) v[P::BX2] = u[P::BX2];
    if constexpr (P::ncomp == 3) v[P::BX3] = u[P::BX3];
  }
  if constexpr (P::energy) {
    double m2 = u[P::MX1] * u[P::MX1];
    if constexpr (P::ncomp >= 2) m2 = m2 + u[P::MX2] * u[P::MX2];
    if constexpr (P::ncomp == 3) m2 = m2 + u[P::MX3] * u[P::MX3];
    double kin = 0.5 / u[P::RHO] * m2;
    if constexpr (P::mhd) {
      double b2 = u[P::BX1] * u[P::BX1];
      if constexpr (P::ncomp >= 2) b2 = b2 + u[P::BX2] * u[P::BX2];
      if constexpr (P::ncomp == 3) b2 = b2 + u[P::BX3] * u[P::BX3];
      double mag = 0.5 * b2;
      v[P::PRS] = eos.pressure(u[P::ENG] - kin - mag);
      if (v[P::PRS] <= 0.0) {
        v[P::PRS] = SMALL_PRESSURE_FIX;
        u[P::ENG] = eos.internalEnergy(v[P::PRS]) + kin + mag;
        fixed = true;
      }
    } else {
      v[P::PRS] = eos.pressure(u[P::ENG] - kin);
      if (v[P::PRS] <= 0.0) {
        v[P::PRS] = SMALL_PRESSURE_FIX;
        u[P::ENG] = eos.internalEnergy(v[P::PRS]) + kin;
        fixed = true;
      }
    }
  }
  return fixed;
}

// flux.hpp:29-95.  f may alias u (the MHD solvers call it that way, hlldMHD.hpp:246-252):
// every u[] entry is consumed before the corresponding f[] entry is written.
template <class P, int DIR, class V>
IDFX_HD void physFlux(double *f, const V &v, const double *u, double c2iso) {
  using D = Dirs<P, DIR>;
  constexpr int Xn = D::Xn;
  const double un = u[Xn];
  [[maybe_unused]] double ueng = 0.0;
  if constexpr (P::energy) ueng = u[P::ENG];
  if constexpr (P::mhd) {
    constexpr int BXn = D::BXn;
    const double m1 = u[P::MX1];
    [[maybe_unused]] double m2 = 0.0, m3 = 0.0;
    if constexpr (P::ncomp >= 2) m2 = u[P::MX2];
    if constexpr (P::ncomp == 3) m3 = u[P::MX3];
    f[P::RHO] = un;
    f[P::MX1] = m1 * v[Xn] - v[BXn] * v[P::BX1];
    if constexpr (P::ncomp >= 2) f[P::MX2] = m2 * v[Xn] - v[BXn] * v[P::BX2];
    if constexpr (P::ncomp == 3) f[P::MX3] = m3 * v[Xn] - v[BXn] * v[P::BX3];
    f[BXn] = 0.0;
    if constexpr (P::ncomp >= 2) f[D::BXt] = v[Xn] * v[D::BXt] - v[BXn] * v[D::Xt];
    if constexpr (P::ncomp == 3) f[D::BXb] = v[Xn] * v[D::BXb] - v[BXn] * v[D::Xb];
    double bmag2 = v[P::BX1] * v[P::BX1];
    if constexpr (P::ncomp >= 2) bmag2 = bmag2 + v[P::BX2] * v[P::BX2];
    if constexpr (P::ncomp == 3) bmag2 = bmag2 + v[P::BX3] * v[P::BX3];
    double ptot;
    if constexpr (P::energy) {
      ptot = v[P::PRS] + 0.5 * bmag2;
      double vb = v[P::MX1] * v[P::BX1];
      if constexpr (P::ncomp >= 2) vb = vb + v[P::MX2] * v[P::BX2];
      if constexpr (P::ncomp == 3) vb = vb + v[P::MX3] * v[P::BX3];
      f[P::ENG] = (ueng + ptot) * v[Xn] - v[BXn] * vb;
    } else {
      ptot = c2iso * v[P::RHO] + 0.5 * bmag2;
    }
    f[Xn] += ptot;
  } else {
    const double m1 = u[P::MX1];
    [[maybe_unused]] double m2 = 0.0, m3 = 0.0;
    if constexpr (P::ncomp >= 2) m2 = u[P::MX2];
    if constexpr (P::ncomp == 3) m3 = u[P::MX3];
    f[P::RHO] = un;
    f[P::MX1] = m1 * v[Xn];
    if constexpr (P::ncomp >= 2) f[P::MX2] = m2 * v[Xn];
    if constexpr (P::ncomp == 3) f[P::MX3] = m3 * v[Xn];
    double ptot;
    if constexpr (P::energy) {
      ptot = v[P::PRS];
      f[P::ENG] = (ueng + ptot) * v[Xn];
    } else {
      ptot = c2iso * v[P::RHO];
    }
    f[Xn] += ptot;
  }
}

#endif  // IDFX_PHYS_CUH_
