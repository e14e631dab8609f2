// ct.cu -- constrained transport: corner EMF averaging and staggered-field update.
//
// Reference: src/fluid/constrainedTransport/calcCornerEmf.hpp:15-282 (arithmetic, uct0,
// uct_contact), calcRiemannEmf.hpp:22-431 (uct_hll / uct_hlld), evolveMagField.hpp:15-135.
// Every kernel is a pure streaming stencil; threads run along i.
#include "kernels.h"
#include "recon.cuh"

namespace {

constexpr int TBX = 64, TBY = 4;

struct Box {
  int lo[3], hi[3];  // [lo,hi)
};

inline dim3 boxGrid(const Box &b) {
  int n0 = b.hi[0] - b.lo[0], n1 = b.hi[1] - b.lo[1], n2 = b.hi[2] - b.lo[2];
  if (n0 < 1) n0 = 1;
  if (n1 < 1) n1 = 1;
  if (n2 < 1) n2 = 1;
  return dim3((n0 + TBX - 1) / TBX, (n1 + TBY - 1) / TBY, n2);
}

#define BOX_INDEX(b)                                              \
  const int i = (b).lo[0] + blockIdx.x * TBX + threadIdx.x;       \
  const int j = (b).lo[1] + blockIdx.y * TBY + threadIdx.y;       \
  const int k = (b).lo[2] + blockIdx.z;                           \
  if (i >= (b).hi[0] || j >= (b).hi[1] || k >= (b).hi[2]) return;

// calcCornerEmf.hpp:85-126
template <int ND>
__global__ void cellCenteredEmfKernel(DevGrid g, DevEmf e, const double *__restrict__ Vc, int ncomp, int vx1,
                                      int bx1, Box b) {
  BOX_INDEX(b)
  const long long c = cidx(g, k, j, i);
  double v1 = 0, v2 = 0, v3 = 0, B1 = 0, B2 = 0, B3 = 0;
  v1 = Vc[(vx1 + 0) * g.sn + c];
  B1 = Vc[(bx1 + 0) * g.sn + c];
  if (ncomp >= 2) {
    v2 = Vc[(vx1 + 1) * g.sn + c];
    B2 = Vc[(bx1 + 1) * g.sn + c];
  }
  if (ncomp == 3) {
    v3 = Vc[(vx1 + 2) * g.sn + c];
    B3 = Vc[(bx1 + 2) * g.sn + c];
  }
  if (ND == 3) {
    e.Ex1[c] = (v3 * B2 - v2 * B3);
    e.Ex2[c] = (v1 * B3 - v3 * B1);
  }
  e.Ex3[c] = (v2 * B1 - v1 * B2);
}

// calcCornerEmf.hpp:44-82 and the corner part of uct0 (:174-191)
template <int ND>
__global__ void arithmeticAverageKernel(DevGrid g, DevEmf e, Box b) {
  BOX_INDEX(b)
  const long long c = cidx(g, k, j, i);
  const long long sj = g.sj, sk = g.sk;
  const double w = 0.25;
  if (ND == 3) {
    e.ex[c] = w * (e.exj[c] + e.exj[c - sk] + e.exk[c] + e.exk[c - sj]);
    e.ey[c] = w * (e.eyi[c] + e.eyi[c - sk] + e.eyk[c] + e.eyk[c - 1]);
  }
  e.ez[c] = w * (e.ezi[c] + e.ezi[c - sj] + e.ezj[c] + e.ezj[c - 1]);
}

// calcCornerEmf.hpp:150-172 -- modifies the face EMFs in place
template <int ND>
__global__ void uct0FaceKernel(DevGrid g, DevEmf e, Box b) {
  BOX_INDEX(b)
  const long long c = cidx(g, k, j, i);
  const long long sj = g.sj, sk = g.sk;
  if (ND == 3) {
    double exj = e.exj[c] * 2.0, exk = e.exk[c] * 2.0, eyi = e.eyi[c] * 2.0, eyk = e.eyk[c] * 2.0;
    exj -= 0.5 * (e.Ex1[c - sj] + e.Ex1[c]);
    exk -= 0.5 * (e.Ex1[c - sk] + e.Ex1[c]);
    eyi -= 0.5 * (e.Ex2[c - 1] + e.Ex2[c]);
    eyk -= 0.5 * (e.Ex2[c - sk] + e.Ex2[c]);
    e.exj[c] = exj;
    e.exk[c] = exk;
    e.eyi[c] = eyi;
    e.eyk[c] = eyk;
  }
  double ezi = e.ezi[c] * 2.0, ezj = e.ezj[c] * 2.0;
  ezi -= 0.5 * (e.Ex3[c - 1] + e.Ex3[c]);
  ezj -= 0.5 * (e.Ex3[c - sj] + e.Ex3[c]);
  e.ezi[c] = ezi;
  e.ezj[c] = ezj;
}

// calcCornerEmf.hpp:197-282
template <int ND>
__global__ void contactAverageKernel(DevGrid g, DevEmf e, Box b) {
  BOX_INDEX(b)
  const long long c = cidx(g, k, j, i);
  const long long sj = g.sj, sk = g.sk;
  const double *wsx = e.svx, *wsy = e.svy, *wsz = e.svz;
  {
    const double *ezi = e.ezi, *ezj = e.ezj, *Ex3 = e.Ex3;
    double ez_l2 = (1 - wsx[c - sj]) * (ezj[c] - Ex3[c - sj]) + (wsx[c - sj]) * (ezj[c - 1] - Ex3[c - sj - 1]);
    double ez_r2 = (1 - wsx[c]) * (ezj[c] - Ex3[c]) + (wsx[c]) * (ezj[c - 1] - Ex3[c - 1]);
    double ez_l1 = (1 - wsy[c - 1]) * (ezi[c] - Ex3[c - 1]) + (wsy[c - 1]) * (ezi[c - sj] - Ex3[c - sj - 1]);
    double ez_r1 = (1 - wsy[c]) * (ezi[c] - Ex3[c]) + (wsy[c]) * (ezi[c - sj] - Ex3[c - sj]);
    e.ez[c] = 0.25 * (ez_l2 + ez_r2 + ez_l1 + ez_r1 + ezi[c] + ezi[c - sj] + ezj[c] + ezj[c - 1]);
  }
  if (ND == 3) {
    const double *exj = e.exj, *exk = e.exk, *eyi = e.eyi, *eyk = e.eyk, *Ex1 = e.Ex1, *Ex2 = e.Ex2;
    double ex_l3 = (1 - wsy[c - sk]) * (exk[c] - Ex1[c - sk]) + (wsy[c - sk]) * (exk[c - sj] - Ex1[c - sk - sj]);
    double ex_r3 = (1 - wsy[c]) * (exk[c] - Ex1[c]) + (wsy[c]) * (exk[c - sj] - Ex1[c - sj]);
    double ex_l2 = (1 - wsz[c - sj]) * (exj[c] - Ex1[c - sj]) + (wsz[c - sj]) * (exj[c - sk] - Ex1[c - sk - sj]);
    double ex_r2 = (1 - wsz[c]) * (exj[c] - Ex1[c]) + (wsz[c]) * (exj[c - sk] - Ex1[c - sk]);
    e.ex[c] = 0.25 * (ex_l3 + ex_r3 + ex_l2 + ex_r2 + exj[c] + exj[c - sk] + exk[c] + exk[c - sj]);

    double ey_l3 = (1 - wsx[c - sk]) * (eyk[c] - Ex2[c - sk]) + (wsx[c - sk]) * (eyk[c - 1] - Ex2[c - sk - 1]);
    double ey_r3 = (1 - wsx[c]) * (eyk[c] - Ex2[c]) + (wsx[c]) * (eyk[c - 1] - Ex2[c - 1]);
    double ey_l1 = (1 - wsz[c - 1]) * (eyi[c] - Ex2[c - 1]) + (wsz[c - 1]) * (eyi[c - sk] - Ex2[c - sk - 1]);
    double ey_r1 = (1 - wsz[c]) * (eyi[c] - Ex2[c]) + (wsz[c]) * (eyi[c - sk] - Ex2[c - sk]);
    e.ey[c] = 0.25 * (ey_l3 + ey_r3 + ey_l1 + ey_r1 + eyi[c] + eyi[c - sk] + eyk[c] + eyk[c - 1]);
  }
}

// transverse reconstruction of a face quantity towards the corner (calcRiemannEmf.hpp:80-126):
// q at the corner-adjacent position `p`, pm = p - s.  qL = value extrapolated from pm, qR from p.
template <bool PPM>
__device__ __forceinline__ void cornerStates(const double *__restrict__ q, long long p, long long s, double &qL,
                                             double &qR) {
  const long long pm = p - s;
  if (PPM) {
    double dummy;
    lim::ppmStates(q[pm - 2 * s], q[pm - s], q[pm], q[pm + s], q[pm + 2 * s], dummy, qL);
    lim::ppmStates(q[p - 2 * s], q[p - s], q[p], q[p + s], q[p + 2 * s], qR, dummy);
  } else {
    double d = lim::mc(q[pm + s] - q[pm], q[pm] - q[pm - s]);
    qL = q[pm] + 0.5 * d;
    d = lim::mc(q[p + s] - q[p], q[p] - q[p - s]);
    qR = q[p] - 0.5 * d;
  }
}

// One (a,d)-weighted contribution: coefficients averaged across `sa`, states reconstructed along `sr`.
template <bool PPM>
__device__ __forceinline__ double riemannContribution(const double *aLp, const double *aRp, const double *dLp,
                                                      const double *dRp, long long c, long long sa,
                                                      const double *__restrict__ B, long long cB, long long srB,
                                                      const double *__restrict__ E, long long sr, double sign) {
  // c - sa is the "minus" neighbour in the averaging direction
  const double aL = 0.5 * (aLp[c - sa] + aLp[c]);
  const double aR = 0.5 * (aRp[c - sa] + aRp[c]);
  const double dL = 0.5 * (dLp[c - sa] + dLp[c]);
  const double dR = 0.5 * (dRp[c - sa] + dRp[c]);
  double bL, bR, vL, vR;
  cornerStates<PPM>(B, cB, srB, bL, bR);
  cornerStates<PPM>(E, c, sr, vL, vR);
  const double phi = dR * bR - dL * bL;
  return (aL * vL * bL + aR * vR * bR) + sign * phi;
}

// calcRiemannEmf.hpp:62-428
template <int ND, bool PPM>
__global__ void riemannAverageKernel(DevGrid g, DevEmf e, const double *__restrict__ Vs, Box b) {
  BOX_INDEX(b)
  const long long c = cidx(g, k, j, i);
  const long long sj = g.sj, sk = g.sk;
  const long long s = sidx(g, k, j, i);
  const double *B1 = Vs, *B2 = Vs + g.vsn, *B3 = Vs + 2 * g.vsn;
  // X-face data: a/d from the x sweep averaged in j, By reconstructed along i, vx-type face EMF ezj along i
  double ez = riemannContribution<PPM>(e.axL, e.axR, e.dxL, e.dxR, c, sj, B2, s, 1, e.ezj, 1, +1.0);
  double ex = 0.0, ey = 0.0;
  if (ND == 3) {
    ey = riemannContribution<PPM>(e.axL, e.axR, e.dxL, e.dxR, c, sk, B3, s, 1, e.eyk, 1, -1.0);
    ex = riemannContribution<PPM>(e.ayL, e.ayR, e.dyL, e.dyR, c, sk, B3, s, g.vsj, e.exk, sj, +1.0);
  }
  ez += riemannContribution<PPM>(e.ayL, e.ayR, e.dyL, e.dyR, c, 1, B1, s, g.vsj, e.ezi, sj, -1.0);
  if (ND == 3) {
    ey += riemannContribution<PPM>(e.azL, e.azR, e.dzL, e.dzR, c, 1, B1, s, g.vsk, e.eyi, sk, +1.0);
    ex += riemannContribution<PPM>(e.azL, e.azR, e.dzL, e.dzR, c, sj, B2, s, g.vsk, e.exj, sk, -1.0);
    e.ex[c] = ex;
    e.ey[c] = ey;
  }
  e.ez[c] = ez;
}

// evolveMagField.hpp:53-131
// CART: cartesian geometry fixed at compile time (drops the curvilinear branches and their registers, which
// doubles the resident warps of this latency-bound kernel).  A thread owns an (i, j) column and marches over a
// chunk of KCH planes: the corner EMFs of plane k+1 it loads for the x1 / x2 faces are those of plane k of the next
// step (kept in registers: one third fewer L2 requests), and the loads of a step do not wait for the stores of the
// previous one.
constexpr int EVOLVE_KCH = 8;
template <int ND, bool CART>
__global__ void __launch_bounds__(TBX *TBY, 4)
    evolveMagFieldKernel(DevGrid g, DevEmf e, double *__restrict__ Vs, double *__restrict__ Vs0, const double *dtp,
                         int haveAxis, int mode, double wc, double w0, Box b) {
  const int i = b.lo[0] + blockIdx.x * TBX + threadIdx.x;
  const int j = b.lo[1] + blockIdx.y * TBY + threadIdx.y;
  const int k0 = b.lo[2] + blockIdx.z * EVOLVE_KCH;
  if (i >= b.hi[0] || j >= b.hi[1] || k0 >= b.hi[2]) return;
  const int k1 = min(k0 + EVOLVE_KCH, b.hi[2]);
  const double dt = *dtp;
  const long long sj = g.sj, sk = g.sk;
  const double *__restrict__ Ex1 = e.ex, *__restrict__ Ex2 = e.ey, *__restrict__ Ex3 = e.ez;
  const double *dx1 = g.dx[0], *dx2 = g.dx[1], *dx3 = g.dx[2];
  const int geo = CART ? (int)GEO_CARTESIAN : g.geometry;
  long long c = cidx(g, k0, j, i);
  long long s = sidx(g, k0, j, i);
  // corner EMFs of the current plane that the next plane re-uses (3-D only)
  double ex1c = 0.0, ex2c = 0.0;
  if (ND == 3) {
    ex1c = Ex1[c];
    ex2c = Ex2[c];
  }
#pragma unroll 2
  for (int k = k0; k < k1; k++, c += sk, s += g.vsk) {
    double rhsx1 = 0.0, rhsx2 = 0.0, rhsx3 = 0.0;
    double ex1p = 0.0, ex2p = 0.0;   // plane k+1
    if (ND == 3) {
      ex1p = Ex1[c + sk];
      ex2p = Ex2[c + sk];
    }
    if (geo == GEO_CARTESIAN) {
      const double ex3c = Ex3[c];
      if (ND >= 2) rhsx1 = rhsx1 - dt / dx2[j] * (Ex3[c + sj] - ex3c);
      if (ND == 3) rhsx1 = rhsx1 + dt / dx3[k] * (ex2p - ex2c);
      if (ND >= 2) {
        rhsx2 = dt / dx1[i] * (Ex3[c + 1] - ex3c);
        if (ND == 3) rhsx2 = rhsx2 - dt / dx3[k] * (ex1p - ex1c);
      }
      if (ND == 3) rhsx3 = -dt / dx1[i] * (Ex2[c + 1] - ex2c) + dt / dx2[j] * (Ex1[c + sj] - ex1c);
    } else if (geo == GEO_CYLINDRICAL) {
      const double *x1 = g.x[0], *x1p = g.xr[0], *x1m = g.xl[0];
      rhsx1 = -dt / dx2[j] * (Ex3[c + sj] - Ex3[c]);
      if (ND >= 2) rhsx2 = dt * (fabs(x1p[i]) * Ex3[c + 1] - fabs(x1m[i]) * Ex3[c]) / fabs(x1[i] * dx1[i]);
    } else if (geo == GEO_POLAR) {
      const double *x1 = g.x[0], *x1m = g.xl[0];
      if (ND >= 2) rhsx1 = rhsx1 - dt / (fabs(x1m[i]) * dx2[j]) * (Ex3[c + sj] - Ex3[c]);
      if (ND == 3) rhsx1 = rhsx1 + dt / dx3[k] * (ex2p - ex2c);
      if (ND >= 2) {
        rhsx2 = dt / dx1[i] * (Ex3[c + 1] - Ex3[c]);
        if (ND == 3) rhsx2 = rhsx2 - dt / dx3[k] * (ex1p - ex1c);
      }
      if (ND == 3)
        rhsx3 = dt / (fabs(x1[i])) *
                (-(x1m[i + 1] * Ex2[c + 1] - x1m[i] * ex2c) / dx1[i] + (Ex1[c + sj] - ex1c) / dx2[j]);
    } else {  // spherical
      const double *x1 = g.x[0], *x1m = g.xl[0];
      const double dV2 = g.dmu[j];
      const double Ax2p = fabs(g.sinx2m[j + 1]);
      double Ax2m = fabs(g.sinx2m[j]);
      if (ND >= 2) rhsx1 = rhsx1 - dt / (x1m[i] * dV2) * (Ax2p * Ex3[c + sj] - Ax2m * Ex3[c]);
      if (ND == 3) rhsx1 = rhsx1 + dt * dx2[j] / (x1m[i] * dV2 * dx3[k]) * (ex2p - ex2c);
      if (ND >= 2) {
        if (haveAxis) {
          if (fabs(Ax2m) < 1e-12) Ax2m = 1.0;
        }
        rhsx2 = dt / (x1[i] * dx1[i]) * (x1m[i + 1] * Ex3[c + 1] - x1m[i] * Ex3[c]);
        if (ND == 3) rhsx2 = rhsx2 - dt / (x1[i] * Ax2m * dx3[k]) * (ex1p - ex1c);
      }
      if (ND == 3)
        rhsx3 = -dt / (x1[i] * dx1[i]) * (x1m[i + 1] * Ex2[c + 1] - x1m[i] * ex2c) +
                dt / (x1[i] * dx2[j]) * (Ex1[c + sj] - ex1c);
    }
    const double rhs[3] = {rhsx1, rhsx2, rhsx3};
#pragma unroll
    for (int d = 0; d < ND; d++) {
      const long long a = d * g.vsn + s;
      const double old = Vs[a];
      double upd = old + rhs[d];
      if (mode == 1) Vs0[a] = old;
      if (mode == 2) upd = wc * upd + w0 * Vs0[a];
      Vs[a] = upd;
    }
    ex1c = ex1p;
    ex2c = ex2p;
  }
}

// ConstrainedTransport::EvolveVectorPotential (evolveVectorPotential.hpp:17-44): Ve += -dt E on [beg, end + 1) of
// the active directions.  mode as in evolveMagFieldKernel: 1 saves the old value in Ve0 (first stage of the fused
// path), 2 folds StateContainer::AddAndStore in (Ve <- wc (Ve - dt E) + w0 Ve0).
// Ve(comp, k, j, i) shares the layout of Vs; DIMENSIONS == 3: comp = AX1e, AX2e, AX3e; DIMENSIONS == 2: AX3e only.
template <int ND>
__global__ void evolveVectorPotentialKernel(DevGrid g, DevEmf e, double *__restrict__ Ve, double *__restrict__ Ve0,
                                            const double *dtp, int mode, double wc, double w0, Box b) {
  BOX_INDEX(b)
  const double dt = *dtp;
  const long long c = cidx(g, k, j, i);
  const long long s = sidx(g, k, j, i);
  const double *E[3] = {e.ex, e.ey, e.ez};
  constexpr int NE = (ND == 3) ? 3 : 1;
#pragma unroll
  for (int n = 0; n < NE; n++) {
    const double emf = (ND == 3) ? E[n][c] : e.ez[c];
    const long long a = n * g.vsn + s;
    const double old = Ve[a];
    double upd = old + (-dt * emf);
    if (mode == 1) Ve0[a] = old;
    if (mode == 2) upd = wc * upd + w0 * Ve0[a];
    Ve[a] = upd;
  }
}

// ConstrainedTransport::ComputeMagFieldFromA (evolveVectorPotential.hpp:48-163): Vs = curl Ve, every geometry
template <int ND>
__global__ void magFieldFromAKernel(DevGrid g, const double *__restrict__ Ve, double *__restrict__ Vs, int haveAxis, Box b) {
  BOX_INDEX(b)
  const long long s = sidx(g, k, j, i);
  const long long sj = g.vsj, sk = g.vsk;
  const double *A1 = Ve, *A2 = Ve + g.vsn, *A3 = (ND == 3) ? Ve + 2 * g.vsn : Ve;   // AX1e, AX2e, AX3e
  const double *dx1 = g.dx[0], *dx2 = g.dx[1], *dx3 = g.dx[2];
  double b1 = 0.0, b2 = 0.0, b3 = 0.0;
  const int geo = g.geometry;
  if (geo == GEO_CARTESIAN) {
    b1 = 0.0;
    if (ND >= 2) b1 = b1 + 1 / dx2[j] * (A3[s + sj] - A3[s]);
    if (ND == 3) b1 = b1 - 1 / dx3[k] * (A2[s + sk] - A2[s]);
    if (ND >= 2) {
      b2 = -1 / dx1[i] * (A3[s + 1] - A3[s]);
      if (ND == 3) b2 = b2 + 1 / dx3[k] * (A1[s + sk] - A1[s]);
    }
    if (ND == 3) b3 = +1 / dx1[i] * (A2[s + 1] - A2[s]) - 1 / dx2[j] * (A1[s + sj] - A1[s]);
  } else if (geo == GEO_CYLINDRICAL) {
    const double *x1 = g.x[0], *x1p = g.xr[0], *x1m = g.xl[0];
    b1 = +1 / dx2[j] * (A3[s + sj] - A3[s]);
    if (ND >= 2) b2 = -(fabs(x1p[i]) * A3[s + 1] - fabs(x1m[i]) * A3[s]) / fabs(x1[i] * dx1[i]);
  } else if (geo == GEO_POLAR) {
    const double *x1 = g.x[0], *x1m = g.xl[0];
    b1 = 0.0;
    if (ND >= 2) b1 = b1 + 1 / (fabs(x1m[i]) * dx2[j]) * (A3[s + sj] - A3[s]);
    if (ND == 3) b1 = b1 - 1 / dx3[k] * (A2[s + sk] - A2[s]);
    if (ND >= 2) {
      b2 = -1 / dx1[i] * (A3[s + 1] - A3[s]);
      if (ND == 3) b2 = b2 + 1 / dx3[k] * (A1[s + sk] - A1[s]);
    }
    if (ND == 3)
      b3 = 1 / (fabs(x1[i])) * ((x1m[i + 1] * A2[s + 1] - x1m[i] * A2[s]) / dx1[i] - (A1[s + sj] - A1[s]) / dx2[j]);
  } else {  // spherical
    const double *x1 = g.x[0], *x1m = g.xl[0];
    const double dV2 = g.dmu[j];
    const double Ax2p = fabs(g.sinx2m[j + 1]);
    double Ax2m = fabs(g.sinx2m[j]);
    b1 = 0.0;
    if (ND >= 2) b1 = b1 + 1 / (x1m[i] * dV2) * (Ax2p * A3[s + sj] - Ax2m * A3[s]);
    if (ND == 3) b1 = b1 - dx2[j] / (x1m[i] * dV2 * dx3[k]) * (A2[s + sk] - A2[s]);
    if (ND >= 2) {
      if (haveAxis) {
        if (fabs(Ax2m) < 1e-12) Ax2m = 1.0;
      }
      b2 = -1 / (x1[i] * dx1[i]) * (x1m[i + 1] * A3[s + 1] - x1m[i] * A3[s]);
      if (ND == 3) b2 = b2 + 1 / (x1[i] * Ax2m * dx3[k]) * (A1[s + sk] - A1[s]);
    }
    if (ND == 3)
      b3 = +1 / (x1[i] * dx1[i]) * (x1m[i + 1] * A2[s + 1] - x1m[i] * A2[s]) - 1 / (x1[i] * dx2[j]) * (A1[s + sj] - A1[s]);
  }
  Vs[s] = b1;
  if (ND >= 2) Vs[g.vsn + s] = b2;
  if (ND == 3) Vs[2 * g.vsn + s] = b3;
}

Box cornerBox(const DevGrid &g, int extra) {
  Box b;
  for (int d = 0; d < 3; d++) {
    b.lo[d] = g.beg[d] - extra * g.off[d];
    b.hi[d] = g.end[d] + g.off[d];
  }
  return b;
}

}  // namespace

#define LAUNCH_ND(ndim, kern, grid, block, stream, ...)                         \
  do {                                                                          \
    if ((ndim) == 2) kern<2><<<grid, block, 0, stream>>>(__VA_ARGS__);          \
    else kern<3><<<grid, block, 0, stream>>>(__VA_ARGS__);                      \
  } while (0)

// ConstrainedTransport::CalcCornerEMF (calcCornerEmf.hpp:15-38)
int idfxk::calcCornerEmf(const DevGrid &g, const DevEmf &e, const double *Vc, const double *Vs, int emfMode,
                         int order, int ncomp, int vx1, int bx1, cudaStream_t st, long long *launches,
                         int emfFromSweep) {
  if (g.ndim < 2) return 0;
  const dim3 block(TBX, TBY, 1);
  Box corner = cornerBox(g, 0);
  if (emfMode == EMF_ARITHMETIC) {
    LAUNCH_ND(g.ndim, arithmeticAverageKernel, boxGrid(corner), block, st, g, e, corner);
    (*launches)++;
  } else if (emfMode == EMF_UCT0 || emfMode == EMF_UCT_CONTACT) {
    Box all;
    for (int d = 0; d < 3; d++) {
      all.lo[d] = 0;
      all.hi[d] = g.nx[d];
    }
    // uct_contact: the cell-centred EMFs of every cell the averaging reads ([beg-1, end] in each direction)
    // were written by the x1 sweep of this stage (sweep_pencil.cuh); the stand-alone kernel remains for uct0
    // (which also reads beg-2) and for callers that have not run a sweep (emfFromSweep == 0)
    if (!(emfFromSweep && emfMode == EMF_UCT_CONTACT)) {
      LAUNCH_ND(g.ndim, cellCenteredEmfKernel, boxGrid(all), block, st, g, e, Vc, ncomp, vx1, bx1, all);
      (*launches)++;
    }
    if (emfMode == EMF_UCT0) {
      Box fb;
      for (int d = 0; d < 3; d++) {
        fb.lo[d] = g.beg[d] - g.off[d];
        fb.hi[d] = g.end[d] + g.off[d];
      }
      LAUNCH_ND(g.ndim, uct0FaceKernel, boxGrid(fb), block, st, g, e, fb);
      LAUNCH_ND(g.ndim, arithmeticAverageKernel, boxGrid(corner), block, st, g, e, corner);
      (*launches) += 2;
    } else {
      LAUNCH_ND(g.ndim, contactAverageKernel, boxGrid(corner), block, st, g, e, corner);
      (*launches)++;
    }
  } else {
    if (order == 4) {
      if (g.ndim == 2) riemannAverageKernel<2, true><<<boxGrid(corner), block, 0, st>>>(g, e, Vs, corner);
      else riemannAverageKernel<3, true><<<boxGrid(corner), block, 0, st>>>(g, e, Vs, corner);
    } else {
      if (g.ndim == 2) riemannAverageKernel<2, false><<<boxGrid(corner), block, 0, st>>>(g, e, Vs, corner);
      else riemannAverageKernel<3, false><<<boxGrid(corner), block, 0, st>>>(g, e, Vs, corner);
    }
    (*launches)++;
  }
  return (int)cudaGetLastError();
}

// ConstrainedTransport::EvolveMagField (evolveMagField.hpp:15-135)
int idfxk::evolveMagField(const DevGrid &g, const DevEmf &e, double *Vs, double *Vs0, const double *dt, int haveAxis,
                          int mode, double wc, double w0, cudaStream_t st, long long *launches) {
  const dim3 block(TBX, TBY, 1);
  Box b = cornerBox(g, 0);
  if (g.ndim == 1) {
    return 0;  // EvolveStage only calls CT for DIMENSIONS >= 2 (evolveStage.hpp:70-87)
  }
  dim3 grid = boxGrid(b);
  grid.z = (grid.z + EVOLVE_KCH - 1) / EVOLVE_KCH;   // a thread marches over EVOLVE_KCH planes
  if (g.geometry == GEO_CARTESIAN) {
    if (g.ndim == 2) evolveMagFieldKernel<2, true><<<grid, block, 0, st>>>(g, e, Vs, Vs0, dt, haveAxis, mode, wc, w0, b);
    else evolveMagFieldKernel<3, true><<<grid, block, 0, st>>>(g, e, Vs, Vs0, dt, haveAxis, mode, wc, w0, b);
  } else {
    if (g.ndim == 2) evolveMagFieldKernel<2, false><<<grid, block, 0, st>>>(g, e, Vs, Vs0, dt, haveAxis, mode, wc, w0, b);
    else evolveMagFieldKernel<3, false><<<grid, block, 0, st>>>(g, e, Vs, Vs0, dt, haveAxis, mode, wc, w0, b);
  }
  (*launches)++;
  return (int)cudaGetLastError();
}

// EvolveVectorPotential + (optionally) the RK combination of Ve
int idfxk::evolveVectorPotential(const DevGrid &g, const DevEmf &e, double *Ve, double *Ve0, const double *dt, int mode,
                                 double wc, double w0, cudaStream_t st, long long *launches) {
  if (g.ndim == 1) return 0;
  const dim3 block(TBX, TBY, 1);
  Box b = cornerBox(g, 0);
  if (g.ndim == 2) evolveVectorPotentialKernel<2><<<boxGrid(b), block, 0, st>>>(g, e, Ve, Ve0, dt, mode, wc, w0, b);
  else evolveVectorPotentialKernel<3><<<boxGrid(b), block, 0, st>>>(g, e, Ve, Ve0, dt, mode, wc, w0, b);
  (*launches)++;
  return (int)cudaGetLastError();
}

// ComputeMagFieldFromA on [beg, end + 1) of the active directions
int idfxk::magFieldFromA(const DevGrid &g, const double *Ve, double *Vs, int haveAxis, cudaStream_t st, long long *launches) {
  if (g.ndim == 1) return 0;
  const dim3 block(TBX, TBY, 1);
  Box b = cornerBox(g, 0);
  if (g.ndim == 2) magFieldFromAKernel<2><<<boxGrid(b), block, 0, st>>>(g, Ve, Vs, haveAxis, b);
  else magFieldFromAKernel<3><<<boxGrid(b), block, 0, st>>>(g, Ve, Vs, haveAxis, b);
  (*launches)++;
  return (int)cudaGetLastError();
}
