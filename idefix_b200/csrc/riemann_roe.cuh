// riemann_roe.cuh -- Roe-type linearised MHD solver (Cargo & Gallice eigenvectors).
//
// Reference: src/fluid/RiemannSolver/MHDsolvers/roeMHD.hpp:180-654.
// The wave decomposition is kept mode by mode (lambda, eta, R column) but the
// dissipation sum is accumulated in mode order 0..nvar-1 exactly like the
// reference's dense Rc[nvar][nvar] product; entries that are structurally zero are
// skipped, which leaves every partial sum unchanged.
#ifndef IDFX_RIEMANN_ROE_CUH_
#define IDFX_RIEMANN_ROE_CUH_

#include "phys.cuh"
#include "riemann_hd.cuh"
#include "riemann_mhd.cuh"

template <class P>
struct RoeModes {
  static constexpr int KFASTM = 0;
  static constexpr int KFASTP = 1;
  static constexpr int KDIVB = 2;
  static constexpr int KENTRP = (P::energy ? 3 : -1);
  static constexpr int KSLOWM = (P::energy ? 4 : 3);
  static constexpr int KSLOWP = (P::energy ? 5 : 4);
  static constexpr int KALFVM = (P::energy ? 6 : 5);
  static constexpr int KALFVP = (P::energy ? 7 : 6);
};

template <class P, int DIR, class VL, class VR>
IDFX_D void riemannRoeMHD(const VL &vL, const VR &vR, const Eos &eos, double *F, RiemannAux &aux) {
  using D = Dirs<P, DIR>;
  using M = RoeModes<P>;
  constexpr int Xn = D::Xn, Xt = D::Xt, Xb = D::Xb;
  constexpr int BXn = D::BXn, BXt = D::BXt, BXb = D::BXb;
  constexpr int RHO = P::RHO;
  constexpr int NV = P::nvar;
  const double delta = 1.e-6;

  double dV[NV], uL[NV], uR[NV], dU[NV], fL[NV], fR[NV];
  double Rc[NV][NV];

#pragma unroll
  for (int nv = 0; nv < NV; nv++) dV[nv] = vR[nv] - vL[nv];

  primToCons<P>(uL, vL, eos);
  primToCons<P>(uR, vR, eos);

  double a, a2, a2L, a2R;
  [[maybe_unused]] double gamma = eos.gamma;
  if constexpr (P::energy) {
    a2L = 1.0;
    a2R = 1.0;
  } else {
    a2L = 0.5 * (eos.csiso + eos.csiso);
    a2L = a2L * a2L;
    a2R = a2L;
  }
#pragma unroll
  for (int nv = 0; nv < NV; nv++) {
    fL[nv] = uL[nv];
    fR[nv] = uR[nv];
    dU[nv] = uR[nv] - uL[nv];
  }
  physFlux<P, DIR>(fL, vL, fL, a2L);
  physFlux<P, DIR>(fR, vR, fR, a2R);

#pragma unroll
  for (int n1 = 0; n1 < NV; n1++) {
#pragma unroll
    for (int n2 = 0; n2 < NV; n2++) Rc[n1][n2] = 0;
  }

  // Roe averages (roeMHD.hpp:258-300)
  const double sqr_rho_L = idfxSqrt<P::fast>(vL[RHO]);
  const double sqr_rho_R = idfxSqrt<P::fast>(vR[RHO]);
  double sl = idfxDiv<P::fast>(sqr_rho_L, (sqr_rho_L + sqr_rho_R));
  double sr = idfxDiv<P::fast>(sqr_rho_R, (sqr_rho_L + sqr_rho_R));
  const double rho = sr * vL[RHO] + sl * vR[RHO];
  const double sqrt_rho = idfxSqrt<P::fast>(rho);

  [[maybe_unused]] double u = 0, v = 0, w = 0, Bx = 0, By = 0, Bz = 0, bx = 0, by = 0, bz = 0;
  u = sl * vL[Xn] + sr * vR[Xn];
  if constexpr (P::ncomp >= 2) v = sl * vL[Xt] + sr * vR[Xt];
  if constexpr (P::ncomp == 3) w = sl * vL[Xb] + sr * vR[Xb];
  Bx = sr * vL[BXn] + sl * vR[BXn];
  if constexpr (P::ncomp >= 2) By = sr * vL[BXt] + sl * vR[BXt];
  if constexpr (P::ncomp == 3) Bz = sr * vL[BXb] + sl * vR[BXb];
  const double sBx = (Bx >= 0.0 ? 1.0 : -1.0);
  bx = idfxDiv<P::fast>(Bx, sqrt_rho);
  if constexpr (P::ncomp >= 2) by = idfxDiv<P::fast>(By, sqrt_rho);
  if constexpr (P::ncomp == 3) bz = idfxDiv<P::fast>(Bz, sqrt_rho);
  double bt2 = 0.0;
  if constexpr (P::ncomp >= 2) bt2 = bt2 + by * by;
  if constexpr (P::ncomp == 3) bt2 = bt2 + bz * bz;
  const double b2 = bx * bx + bt2;
  const double Btmag = idfxSqrt0<P::fast>(bt2 * rho);

  double X = dV[BXn] * dV[BXn];
  if constexpr (P::ncomp >= 2) X = X + dV[BXt] * dV[BXt];
  if constexpr (P::ncomp == 3) X = X + dV[BXb] * dV[BXb];
  X = idfxDiv<P::fast>(X, (sqr_rho_L + sqr_rho_R) * (sqr_rho_L + sqr_rho_R) * 2.0);

  double Bmag2L = vL[P::BX1] * vL[P::BX1];
  double Bmag2R = vR[P::BX1] * vR[P::BX1];
  if constexpr (P::ncomp >= 2) {
    Bmag2L = Bmag2L + vL[P::BX2] * vL[P::BX2];
    Bmag2R = Bmag2R + vR[P::BX2] * vR[P::BX2];
  }
  if constexpr (P::ncomp == 3) {
    Bmag2L = Bmag2L + vL[P::BX3] * vL[P::BX3];
    Bmag2R = Bmag2R + vR[P::BX3] * vR[P::BX3];
  }
  [[maybe_unused]] double vel2 = 0, Hgas = 0;
  if constexpr (P::energy) {
    const double pL = vL[P::PRS] + 0.5 * Bmag2L;
    const double pR = vR[P::PRS] + 0.5 * Bmag2R;
    double vdm = u * dU[Xn];
    if constexpr (P::ncomp >= 2) vdm = vdm + v * dU[Xt];
    if constexpr (P::ncomp == 3) vdm = vdm + w * dU[Xb];
    double BdB = Bx * dU[BXn];
    if constexpr (P::ncomp >= 2) BdB = BdB + By * dU[BXt];
    if constexpr (P::ncomp == 3) BdB = BdB + Bz * dU[BXb];
    vel2 = u * u;
    if constexpr (P::ncomp >= 2) vel2 = vel2 + v * v;
    if constexpr (P::ncomp == 3) vel2 = vel2 + w * w;
    dV[P::PRS] = (gamma - 1.0) * ((0.5 * vel2 - X) * dV[RHO] - vdm + dU[P::ENG] - BdB);
    const double HL = idfxDiv<P::fast>((uL[P::ENG] + pL), vL[RHO]);
    const double HR = idfxDiv<P::fast>((uR[P::ENG] + pR), vR[RHO]);
    const double H = sl * HL + sr * HR;
    Hgas = H - b2;
    a2 = (2.0 - gamma) * X + (gamma - 1.0) * (Hgas - 0.5 * vel2);
  } else {
    a2 = 0.5 * (a2L + a2R) + X;
  }

  // fast / slow / Alfven speeds (roeMHD.hpp:336-371)
  double scrh = a2 - b2;
  const double ca2 = bx * bx;
  scrh = scrh * scrh + 4.0 * bt2 * a2;
  scrh = idfxSqrt0<P::fast>(scrh);
  const double cf2 = 0.5 * (a2 + b2 + scrh);
  const double cs2 = idfxDiv<P::fast>(a2 * ca2, cf2);
  const double cf = idfxSqrt<P::fast>(cf2);
  const double cs = idfxSqrt0<P::fast>(cs2);
  const double ca = idfxSqrt0<P::fast>(ca2);
  a = idfxSqrt<P::fast>(a2);

  double alpha_f, alpha_s;
  if (cf == cs) {
    alpha_f = 1.0;
    alpha_s = 0.0;
  } else if (a <= cs) {
    alpha_f = 0.0;
    alpha_s = 1.0;
  } else if (cf <= a) {
    alpha_f = 1.0;
    alpha_s = 0.0;
  } else {
    scrh = idfxDiv<P::fast>(1.0, (cf2 - cs2));
    alpha_f = (a2 - cs2) * scrh;
    alpha_s = (cf2 - a2) * scrh;
    alpha_f = fmax(0.0, alpha_f);
    alpha_s = fmax(0.0, alpha_s);
    alpha_f = idfxSqrt0<P::fast>(alpha_f);
    alpha_s = idfxSqrt0<P::fast>(alpha_s);
  }

  [[maybe_unused]] double beta_y = 0, beta_z = 0;
  if (Btmag > 1.e-9) {
    if constexpr (P::ncomp == 2) beta_y = (By >= 0.0 ? 1.0 : -1.0);
    if constexpr (P::ncomp == 3) {
      beta_y = idfxDiv<P::fast>(By, Btmag);
      beta_z = idfxDiv<P::fast>(Bz, Btmag);
    }
  } else {
    if constexpr (P::ncomp == 2) beta_y = 1.0;
    if constexpr (P::ncomp == 3) beta_z = beta_y = 1.0;
  }

  double lambda[NV], alambda[NV], eta[NV];
  [[maybe_unused]] double beta_dv = 0.0, beta_dB = 0.0, beta_v = 0.0;

  // ---- fast wave u - cf
  int kk = M::KFASTM;
  lambda[kk] = u - cf;
  scrh = alpha_s * cs * sBx;
  if constexpr (P::ncomp >= 2) {
    beta_dv = beta_dv + beta_y * dV[Xt];
    beta_dB = beta_dB + beta_y * dV[BXt];
  }
  if constexpr (P::ncomp == 3) {
    beta_dv = beta_dv + beta_z * dV[Xb];
    beta_dB = beta_dB + beta_z * dV[BXb];
  }
  Rc[RHO][kk] = alpha_f;
  Rc[Xn][kk] = alpha_f * lambda[kk];
  if constexpr (P::ncomp >= 2) Rc[Xt][kk] = alpha_f * v + scrh * beta_y;
  if constexpr (P::ncomp == 3) Rc[Xb][kk] = alpha_f * w + scrh * beta_z;
  if constexpr (P::ncomp >= 2) Rc[BXt][kk] = idfxDiv<P::fast>(alpha_s * a * beta_y, sqrt_rho);
  if constexpr (P::ncomp == 3) Rc[BXb][kk] = idfxDiv<P::fast>(alpha_s * a * beta_z, sqrt_rho);
  if constexpr (P::energy) {
    if constexpr (P::ncomp >= 2) beta_v = beta_v + beta_y * v;
    if constexpr (P::ncomp == 3) beta_v = beta_v + beta_z * w;
    Rc[P::ENG][kk] = alpha_f * (Hgas - u * cf) + scrh * beta_v + idfxDiv<P::fast>(alpha_s * a * Btmag, sqrt_rho);
    eta[kk] = alpha_f * (X * dV[RHO] + dV[P::PRS]);
  } else {
    eta[kk] = alpha_f * a2 * dV[RHO];
  }
  eta[kk] += rho * scrh * beta_dv - rho * alpha_f * cf * dV[Xn] + sqrt_rho * alpha_s * a * beta_dB;
  eta[kk] *= idfxDiv<P::fast>(0.5, a2);

  // ---- fast wave u + cf
  kk = M::KFASTP;
  lambda[kk] = u + cf;
  Rc[RHO][kk] = alpha_f;
  Rc[Xn][kk] = alpha_f * lambda[kk];
  if constexpr (P::ncomp >= 2) Rc[Xt][kk] = alpha_f * v - scrh * beta_y;
  if constexpr (P::ncomp == 3) Rc[Xb][kk] = alpha_f * w - scrh * beta_z;
  if constexpr (P::ncomp >= 2) Rc[BXt][kk] = Rc[BXt][M::KFASTM];
  if constexpr (P::ncomp == 3) Rc[BXb][kk] = Rc[BXb][M::KFASTM];
  if constexpr (P::energy) {
    Rc[P::ENG][kk] = alpha_f * (Hgas + u * cf) - scrh * beta_v + idfxDiv<P::fast>(alpha_s * a * Btmag, sqrt_rho);
    eta[kk] = alpha_f * (X * dV[RHO] + dV[P::PRS]) - rho * scrh * beta_dv + rho * alpha_f * cf * dV[Xn] +
              sqrt_rho * alpha_s * a * beta_dB;
  } else {
    eta[kk] = alpha_f * (0. * X + a2) * dV[RHO] - rho * scrh * beta_dv + rho * alpha_f * cf * dV[Xn] +
              sqrt_rho * alpha_s * a * beta_dB;
  }
  eta[kk] *= idfxDiv<P::fast>(0.5, a2);

  // ---- entropy wave
  if constexpr (P::energy) {
    kk = M::KENTRP;
    lambda[kk] = u;
    Rc[RHO][kk] = 1.0;
    Rc[Xn][kk] = u;
    if constexpr (P::ncomp >= 2) Rc[Xt][kk] = v;
    if constexpr (P::ncomp == 3) Rc[Xb][kk] = w;
    Rc[P::ENG][kk] = 0.5 * vel2 + idfxDiv<P::fast>((gamma - 2.0), (gamma - 1.0)) * X;
    eta[kk] = idfxDiv<P::fast>(((a2 - X) * dV[RHO] - dV[P::PRS]), a2);
  }

  // ---- div B wave: carries no jump with CT
  kk = M::KDIVB;
  lambda[kk] = u;
  eta[kk] = 0.0;

  if constexpr (P::ncomp > 1) {
    // ---- slow wave u - cs
    scrh = alpha_f * cf * sBx;
    kk = M::KSLOWM;
    lambda[kk] = u - cs;
    Rc[RHO][kk] = alpha_s;
    Rc[Xn][kk] = alpha_s * lambda[kk];
    if constexpr (P::ncomp >= 2) Rc[Xt][kk] = alpha_s * v - scrh * beta_y;
    if constexpr (P::ncomp == 3) Rc[Xb][kk] = alpha_s * w - scrh * beta_z;
    if constexpr (P::ncomp >= 2) Rc[BXt][kk] = idfxDiv<P::fast>(-alpha_f * a * beta_y, sqrt_rho);
    if constexpr (P::ncomp == 3) Rc[BXb][kk] = idfxDiv<P::fast>(-alpha_f * a * beta_z, sqrt_rho);
    if constexpr (P::energy) {
      Rc[P::ENG][kk] = alpha_s * (Hgas - u * cs) - scrh * beta_v - idfxDiv<P::fast>(alpha_f * a * Btmag, sqrt_rho);
      eta[kk] = alpha_s * (X * dV[RHO] + dV[P::PRS]) - rho * scrh * beta_dv - rho * alpha_s * cs * dV[Xn] -
                sqrt_rho * alpha_f * a * beta_dB;
    } else {
      eta[kk] = alpha_s * (0. * X + a2) * dV[RHO] - rho * scrh * beta_dv - rho * alpha_s * cs * dV[Xn] -
                sqrt_rho * alpha_f * a * beta_dB;
    }
    eta[kk] *= idfxDiv<P::fast>(0.5, a2);

    // ---- slow wave u + cs
    kk = M::KSLOWP;
    lambda[kk] = u + cs;
    Rc[RHO][kk] = alpha_s;
    Rc[Xn][kk] = alpha_s * lambda[kk];
    if constexpr (P::ncomp >= 2) Rc[Xt][kk] = alpha_s * v + scrh * beta_y;
    if constexpr (P::ncomp == 3) Rc[Xb][kk] = alpha_s * w + scrh * beta_z;
    if constexpr (P::ncomp >= 2) Rc[BXt][kk] = Rc[BXt][M::KSLOWM];
    if constexpr (P::ncomp == 3) Rc[BXb][kk] = Rc[BXb][M::KSLOWM];
    if constexpr (P::energy) {
      Rc[P::ENG][kk] = alpha_s * (Hgas + u * cs) + scrh * beta_v - idfxDiv<P::fast>(alpha_f * a * Btmag, sqrt_rho);
      eta[kk] = alpha_s * (X * dV[RHO] + dV[P::PRS]) + rho * scrh * beta_dv + rho * alpha_s * cs * dV[Xn] -
                sqrt_rho * alpha_f * a * beta_dB;
    } else {
      eta[kk] = alpha_s * (0. * X + a2) * dV[RHO] + rho * scrh * beta_dv + rho * alpha_s * cs * dV[Xn] -
                sqrt_rho * alpha_f * a * beta_dB;
    }
    eta[kk] *= idfxDiv<P::fast>(0.5, a2);
  }

  if constexpr (P::ncomp == 3) {
    // ---- Alfven wave u - ca
    kk = M::KALFVM;
    lambda[kk] = u - ca;
    Rc[Xt][kk] = -rho * beta_z;
    Rc[Xb][kk] = +rho * beta_y;
    Rc[BXt][kk] = -sBx * sqrt_rho * beta_z;
    Rc[BXb][kk] = sBx * sqrt_rho * beta_y;
    if constexpr (P::energy) Rc[P::ENG][kk] = -rho * (v * beta_z - w * beta_y);
    eta[kk] = +beta_y * dV[Xb] - beta_z * dV[Xt] + idfxDiv<P::fast>(sBx, sqrt_rho) * (beta_y * dV[BXb] - beta_z * dV[BXt]);
    eta[kk] *= 0.5;

    // ---- Alfven wave u + ca
    kk = M::KALFVP;
    lambda[kk] = u + ca;
    Rc[Xt][kk] = -Rc[Xt][M::KALFVM];
    Rc[Xb][kk] = -Rc[Xb][M::KALFVM];
    Rc[BXt][kk] = Rc[BXt][M::KALFVM];
    Rc[BXb][kk] = Rc[BXb][M::KALFVM];
    if constexpr (P::energy) Rc[P::ENG][kk] = -Rc[P::ENG][M::KALFVM];
    eta[kk] = -beta_y * dV[Xb] + beta_z * dV[Xt] + idfxDiv<P::fast>(sBx, sqrt_rho) * (beta_y * dV[BXb] - beta_z * dV[BXt]);
    eta[kk] *= 0.5;
  }

  aux.cmax = fabs(u) + cf;
  aux.sl = lambda[M::KFASTM];
  aux.sr = lambda[M::KFASTP];
  aux.c2iso = a2L;   // what the reference hands to K_StoreHLLD (roeMHD.hpp:645)

#pragma unroll
  for (int nv = 0; nv < NV; nv++) alambda[nv] = fabs(lambda[nv]);

  // entropy fix (roeMHD.hpp:606-623)
  if (alambda[M::KFASTM] < 0.5 * delta)
    alambda[M::KFASTM] = idfxDiv<P::fast>(lambda[M::KFASTM] * lambda[M::KFASTM], delta) + 0.25 * delta;
  if (alambda[M::KFASTP] < 0.5 * delta)
    alambda[M::KFASTP] = idfxDiv<P::fast>(lambda[M::KFASTP] * lambda[M::KFASTP], delta) + 0.25 * delta;
  if constexpr (P::ncomp > 1) {
    if (alambda[M::KSLOWM] < 0.5 * delta)
      alambda[M::KSLOWM] = idfxDiv<P::fast>(lambda[M::KSLOWM] * lambda[M::KSLOWM], delta) + 0.25 * delta;
    if (alambda[M::KSLOWP] < 0.5 * delta)
      alambda[M::KSLOWP] = idfxDiv<P::fast>(lambda[M::KSLOWP] * lambda[M::KSLOWP], delta) + 0.25 * delta;
  }

#pragma unroll
  for (int n1 = 0; n1 < NV; n1++) {
    double acc = 0.0;
#pragma unroll
    for (int n2 = 0; n2 < NV; n2++) acc += alambda[n2] * eta[n2] * Rc[n1][n2];
    F[n1] = 0.5 * (fL[n1] + fR[n1] - acc);
  }
}

#endif  // IDFX_RIEMANN_ROE_CUH_
