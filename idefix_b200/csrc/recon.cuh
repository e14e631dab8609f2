// recon.cuh -- slope limiters and face reconstruction (one variable at a time).
//
// Reference: src/fluid/RiemannSolver/slopeLimiter.hpp:16-191 and
// src/fluid/RiemannSolver/extrapolateToFaces.hpp:75-309.
// Interface i-1/2 separates cell i-1 (left state vL) from cell i (right state vR).
// `q` points at cell i of the variable, `s` is the element stride along the sweep.
#ifndef IDFX_RECON_CUH_
#define IDFX_RECON_CUH_

#include "common.cuh"
#include "fastmath.cuh"

namespace lim {

// slopeLimiter.hpp:16-23 -- the reference's MinMod shadows its result and therefore
// always returns zero; reproduced on purpose.
IDFX_HD double minmodRef(double, double) { return 0.0; }

// slopeLimiter.hpp:54-57
IDFX_HD double vanLeer(double dvp, double dvm) {
  return (dvp * dvm > 0.0 ? 2.0 * dvp * dvm / (dvp + dvm) : 0.0);
}
IDFX_DEV double vanLeerFast(double dvp, double dvm, double cp, double cm) {
  const bool on = dvp * dvm > 0.0;
  const double q = idfxDivFast(on ? dvp * dvm * (cp * dvm + cm * dvp) : 1.0,
                               on ? dvp * dvp + dvm * dvm + (cp + cm - 2.0) * dvp * dvm : 1.0);
  return on ? q : 0.0;
}
// the same with the branch-free division (device only; fastmath.cuh).  The quotient is formed unconditionally
// and selected afterwards, so the limiter stays inside one basic block; when the product is not positive the
// operands are replaced by harmless ones so that the range check is not tripped by a value that is discarded.
IDFX_DEV double vanLeerFast(double dvp, double dvm) {
  const bool on = dvp * dvm > 0.0;
  const double q = idfxDivFast(on ? 2.0 * dvp * dvm : 1.0, on ? dvp + dvm : 1.0);
  return on ? q : 0.0;
}
// slopeLimiter.hpp:60-65 (non-uniform spacing)
IDFX_HD double vanLeer(double dvp, double dvm, double cp, double cm) {
  return (dvp * dvm > 0.0
              ? dvp * dvm * (cp * dvm + cm * dvp) / (dvp * dvp + dvm * dvm + (cp + cm - 2.0) * dvp * dvm)
              : 0.0);
}
// slopeLimiter.hpp:67-75
IDFX_HD double mc(double dvp, double dvm) {
  double dq = 0.0;
  if (dvp * dvm > 0.0) {
    double dqc = 0.5 * (dvp + dvm);
    double d2q = 2.0 * (fabs(dvp) < fabs(dvm) ? dvp : dvm);
    dq = fabs(d2q) < fabs(dqc) ? d2q : dqc;
  }
  return dq;
}
// slopeLimiter.hpp:25-52.  FAST: branch-free divisions (fastmath.cuh), device code of the sweeps only
template <bool FAST = false>
IDFX_HD double limO3(double dvp, double dvm, double dx) {
  const double r = 0.1;
  const double eps = 1.e-12;
  double th = idfxDiv<FAST>(dvm, (dvp + 1.e-16));
  double q = idfxDiv<FAST>((2.0 + th), 3.0);
  double a = fmin(1.5, 2.0 * th);
  a = fmin(q, a);
  double b = fmax(-0.5 * th, a);
  double c = fmin(q, b);
  double psi = fmax(0.0, c);
  double eta = r * dx;
  eta = idfxDiv<FAST>((dvm * dvm + dvp * dvp), (eta * eta));
  double l;
  if (eta <= 1.0 - eps) {
    l = q;
  } else if (eta >= 1.0 + eps) {
    l = psi;
  } else {
    psi = (1.0 - (eta - 1.0) / eps) * q + (1.0 + (eta - 1.0) / eps) * psi;
    l = 0.5 * psi;
  }
  return l;
}

IDFX_HD int sgn(double v) { return (0.0 < v) - (v < 0.0); }

// slopeLimiter.hpp:121-140
IDFX_HD void ppmLimitFace(double vm1, double v0, double vp1, double vp2, double &vph) {
  if ((vp1 - vph) * (vph - v0) < 0.0) {
    const double deltaL = (vm1 - 2 * v0 + vp1);
    const double deltaC = 3 * (v0 - 2 * vph + vp1);
    const double deltaR = (v0 - 2 * vp1 + vp2);
    double delta = 0.0;
    if (sgn(deltaL) == sgn(deltaC) && sgn(deltaR) == sgn(deltaC)) {
      const double C = 1.25;
      delta = C * fmin(fabs(deltaL), fabs(deltaR));
      delta = sgn(deltaC) * fmin(delta, fabs(deltaC));
    }
    vph = 0.5 * (v0 + vp1) - delta / 6.0;
  }
}

// slopeLimiter.hpp:142-190: limited parabola in cell "0"; vl / vr are its left / right face values
IDFX_HD void ppmStates(double vm2, double vm1, double v0, double vp1, double vp2, double &vl, double &vr) {
  const int n = 2;
  vr = 7.0 / 12.0 * (v0 + vp1) - 1.0 / 12.0 * (vm1 + vp2);
  vl = 7.0 / 12.0 * (vm1 + v0) - 1.0 / 12.0 * (vm2 + vp1);
  ppmLimitFace(vm2, vm1, v0, vp1, vl);
  ppmLimitFace(vm1, v0, vp1, vp2, vr);

  double d2qf = 6.0 * (vl + vr - 2.0 * v0);
  double d2qc0 = vm1 + vp1 - 2.0 * v0;
  double d2qcp1 = v0 + vp2 - 2.0 * vp1;
  double d2qcm1 = vm2 + v0 - 2.0 * vm1;
  double d2q = 0.0;
  if (sgn(d2qf) == sgn(d2qc0) && sgn(d2qf) == sgn(d2qcp1) && sgn(d2qf) == sgn(d2qcm1)) {
    const double C = 1.25;
    d2q = fmin(fabs(d2qc0), fabs(d2qcp1));
    d2q = C * fmin(fabs(d2qcm1), d2q);
    d2q = sgn(d2qf) * fmin(fabs(d2qf), d2q);
  }
  double qmax = fmax(fmax(fabs(vm1), fabs(v0)), fabs(vp1));
  double rho = 0.0;
  if (fabs(d2qf) > 1e-12 * qmax) rho = d2q / d2qf;

  if (((vr - v0) * (v0 - vl) <= 0) || (vm1 - v0) * (v0 - vp1) <= 0) {
    if (rho <= (1.0 - 1e-12)) {
      vl = v0 - rho * (v0 - vl);
      vr = v0 + rho * (vr - v0);
    }
  } else {
    if (fabs(vr - v0) >= n * fabs(v0 - vl)) vr = v0 + n * (v0 - vl);
    if (fabs(vl - v0) >= n * fabs(v0 - vr)) vl = v0 + n * (v0 - vr);
  }
}
}  // namespace lim

// Per-CELL form of the same reconstruction: st[k] holds cell (c - 2 + k), k = 0..4 (orders 1-3 read
// st[1..3] only).  vl / vr are the values of cell c on its left / right face, i.e. the right state
// vR of interface c and the left state vL of interface c+1.  A kernel that walks along the sweep
// direction evaluates this once per cell and re-uses vr at the next interface; the expressions (and
// therefore the bits) are those the reference evaluates twice (extrapolateToFaces.hpp:89-307).
template <int ORDER, bool FAST = false>
IDFX_D void reconstructCell(const double (&st)[5], bool positive, bool regular, const double (&wc)[6], double dxc,
                            double &vl, double &vr) {
  const double qm = st[1], q0 = st[2], qp = st[3];
  if constexpr (ORDER == 1) {
    vl = q0;
    vr = q0;
  } else if constexpr (ORDER == 2) {
    if (regular) {
      const double dvm = q0 - qm;
      const double dvp = qp - q0;
      double dv;
      if constexpr (FAST && IDFX_FASTDIV) dv = lim::vanLeerFast(dvp, dvm);
      else dv = lim::vanLeer(dvp, dvm);
      vr = q0 + 0.5 * dv;
      vl = q0 - 0.5 * dv;
    } else {
      // wc = {wp, wm, cp, cm, dp, dm} of cell c (extrapolateToFaces.hpp:43-71)
      double dvm = q0 - qm;
      double dvp = qp - q0;
      dvm *= wc[1];
      dvp *= wc[0];
      double dv;
      if constexpr (FAST && IDFX_FASTDIV) dv = lim::vanLeerFast(dvp, dvm, wc[2], wc[3]);
      else dv = lim::vanLeer(dvp, dvm, wc[2], wc[3]);
      vr = q0 + wc[4] * dv;
      vl = q0 - wc[5] * dv;
    }
  } else if constexpr (ORDER == 3) {
    const double dvm = q0 - qm;
    const double dvp = qp - q0;
    double dv = dvp * lim::limO3<FAST>(dvp, dvm, dxc);
    vr = q0 + 0.5 * dv;
    if (positive && vr <= 0.0) {
      dv = lim::minmodRef(dvp, dvm);
      vr = q0 + 0.5 * dv;
    }
    dv = dvm * lim::limO3<FAST>(dvm, dvp, dxc);
    vl = q0 - 0.5 * dv;
    if (positive && vl <= 0.0) {
      dv = lim::minmodRef(dvp, dvm);
      vl = q0 - 0.5 * dv;
    }
  } else {
    lim::ppmStates(st[0], qm, q0, qp, st[4], vl, vr);
    if (positive && (vr <= 0.0 || vl <= 0.0)) {
      const double dv = lim::vanLeer(qp - q0, q0 - qm);
      if (vr <= 0.0) vr = q0 + 0.5 * dv;
      if (vl <= 0.0) vl = q0 - 0.5 * dv;
    }
  }
}

#endif  // IDFX_RECON_CUH_
