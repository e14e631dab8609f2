// pointwise.cu -- per-cell kernels: conservative<->primitive conversion, cell-centred
// field reconstruction, Runge-Kutta state combination, and their fused stage
// prologue / epilogue forms.
//
// Reference: src/fluid/convertConsToPrim.hpp:106-189 (array loops),
// src/fluid/boundary/boundary.hpp:389-404 (ReconstructVcField),
// src/dataBlock/stateContainer.cpp:19-37,75-104 (CopyFrom / AddAndStore),
// src/fluid/enroll.hpp:116-129 (ResetStage).
#include "kernels.h"
#include "phys.cuh"

namespace {

constexpr int TB = 256;

// box [lo, hi) of cells
struct CellBox {
  int lo[3], hi[3];
};


inline int blocksFor(long long n) { return (int)((n + TB - 1) / TB); }

// run F<P>::run(args...) for the Phys matching c
#define PHYS_CASE(ND, NC, MHDF, ISOF, CALL)                                                  \
  if (c.ndim == ND && c.ncomp == NC && (c.mhd != 0) == MHDF && (c.iso != 0) == ISOF) {       \
    using P = Phys<ND, NC, MHDF, ISOF>;                                                      \
    CALL;                                                                                    \
    return (int)cudaGetLastError();                                                          \
  }
#define PHYS_DISPATCH(CALL)                \
  PHYS_CASE(1, 1, false, false, CALL)      \
  PHYS_CASE(1, 2, false, false, CALL)      \
  PHYS_CASE(1, 3, false, false, CALL)      \
  PHYS_CASE(2, 2, false, false, CALL)      \
  PHYS_CASE(2, 3, false, false, CALL)      \
  PHYS_CASE(3, 3, false, false, CALL)      \
  PHYS_CASE(1, 1, false, true, CALL)       \
  PHYS_CASE(1, 2, false, true, CALL)       \
  PHYS_CASE(1, 3, false, true, CALL)       \
  PHYS_CASE(2, 2, false, true, CALL)       \
  PHYS_CASE(2, 3, false, true, CALL)       \
  PHYS_CASE(3, 3, false, true, CALL)       \
  PHYS_CASE(1, 2, true, false, CALL)       \
  PHYS_CASE(1, 3, true, false, CALL)       \
  PHYS_CASE(2, 2, true, false, CALL)       \
  PHYS_CASE(2, 3, true, false, CALL)       \
  PHYS_CASE(3, 3, true, false, CALL)       \
  PHYS_CASE(1, 2, true, true, CALL)        \
  PHYS_CASE(1, 3, true, true, CALL)        \
  PHYS_CASE(2, 2, true, true, CALL)        \
  PHYS_CASE(2, 3, true, true, CALL)        \
  PHYS_CASE(3, 3, true, true, CALL)        \
  return -1;

template <class P>
__global__ void primToConsKernel(DevGrid g, const double *__restrict__ Vc, double *__restrict__ Uc, Eos eos,
                                 long long n) {
  const long long c = (long long)blockIdx.x * TB + threadIdx.x;
  if (c >= n) return;
  double v[P::nvar], u[P::nvar];
#pragma unroll
  for (int nv = 0; nv < P::nvar; nv++) v[nv] = Vc[nv * g.sn + c];
  primToCons<P>(u, v, eos);
#pragma unroll
  for (int nv = 0; nv < P::nvar; nv++) Uc[nv * g.sn + c] = u[nv];
}

template <class P>
__global__ void consToPrimKernel(DevGrid g, double *__restrict__ Vc, double *__restrict__ Uc, Eos eos, long long n) {
  const long long c = (long long)blockIdx.x * TB + threadIdx.x;
  if (c >= n) return;
  double v[P::nvar], u[P::nvar];
#pragma unroll
  for (int nv = 0; nv < P::nvar; nv++) u[nv] = Uc[nv * g.sn + c];
  const bool fixed = consToPrim<P>(v, u, eos);
#pragma unroll
  for (int nv = 0; nv < P::nvar; nv++) Vc[nv * g.sn + c] = v[nv];
  if constexpr (P::energy) {
    if (fixed) Uc[P::ENG * g.sn + c] = u[P::ENG];
  }
}

// V(BXn) = 0.5*(Vs(n,.) + Vs(n,.+1)) for the CT components, over the box [lo, hi) of cells
// (the whole block, or one ghost slab: the boundary conditions never touch a face of an active cell, so
// after our own stage epilogue the active cells already hold exactly these values)
__global__ void reconstructVcFieldKernel(DevGrid g, const double *__restrict__ Vs, double *__restrict__ V, int bx1,
                                         CellBox b) {
  const int i = b.lo[0] + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = b.lo[1] + blockIdx.y * blockDim.y + threadIdx.y;
  const int k = b.lo[2] + blockIdx.z;
  if (i >= b.hi[0] || j >= b.hi[1] || k >= b.hi[2]) return;
  const long long c = cidx(g, k, j, i);
  const long long s = sidx(g, k, j, i);
  V[(bx1 + 0) * g.sn + c] = 0.5 * (Vs[s] + Vs[s + 1]);
  if (g.ndim >= 2) V[(bx1 + 1) * g.sn + c] = 0.5 * (Vs[g.vsn + s] + Vs[g.vsn + s + g.vsj]);
  if (g.ndim == 3) V[(bx1 + 2) * g.sn + c] = 0.5 * (Vs[2 * g.vsn + s] + Vs[2 * g.vsn + s + g.vsk]);
}

// shockFlattening.hpp:68-150.  Sums are associated as the reference's D_EXPAND writes them (left to right).
template <int ND>
__global__ void findShockKernel(DevGrid g, const double *__restrict__ Vc, unsigned char *__restrict__ flags, int vx1,
                                int prs, int rho, double smoothing, double cs, CellBox b) {
  const int i = b.lo[0] + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = b.lo[1] + blockIdx.y * blockDim.y + threadIdx.y;
  const int k = b.lo[2] + blockIdx.z;
  if (i >= b.hi[0] || j >= b.hi[1] || k >= b.hi[2]) return;
  const long long c = cidx(g, k, j, i);
  const long long sj = g.sj, sk = g.sk, sn = g.sn;
  const double *v1 = Vc + (long long)vx1 * sn, *v2 = Vc + (long long)(vx1 + 1) * sn, *v3 = Vc + (long long)(vx1 + 2) * sn;
  double divV;
  if (g.geometry == GEO_CARTESIAN) {
    divV = (v1[c + 1] - v1[c - 1]) / (2 * g.dx[0][i]);
    if (ND >= 2) divV = divV + (v2[c + sj] - v2[c - sj]) / (2 * g.dx[1][j]);
    if (ND == 3) divV = divV + (v3[c + sk] - v3[c - sk]) / (2 * g.dx[2][k]);
  } else {
    divV = (v1[c + 1] + v1[c]) * faceAreaDir(g, 0, k, j, i + 1) - (v1[c - 1] + v1[c]) * faceAreaDir(g, 0, k, j, i);
    if (ND >= 2)
      divV = divV + (v2[c + sj] + v2[c]) * faceAreaDir(g, 1, k, j + 1, i) - (v2[c - sj] + v2[c]) * faceAreaDir(g, 1, k, j, i);
    if (ND == 3)
      divV = divV + (v3[c + sk] + v3[c]) * faceAreaDir(g, 2, k + 1, j, i) - (v3[c - sk] + v3[c]) * faceAreaDir(g, 2, k, j, i);
    divV = 0.5 * divV / cellVolume(g, k, j, i);
  }
  unsigned char flag = 0;
  if (divV < 0.0) {
    // pressure of a cell: Vc(PRS), or rho cs^2 for the isothermal equation of state
    auto P = [&](long long cc) { return prs >= 0 ? Vc[(long long)prs * sn + cc] : Vc[(long long)rho * sn + cc] * cs * cs; };
    double pmin = P(c);
    double pR = P(c + 1), pL = P(c - 1);
    if (prs >= 0) {   // shockFlattening.hpp:128-141 (min with i+1 first, then i-1)
      pmin = fmin(pmin, pR);
      pmin = fmin(pmin, pL);
    } else {          // :104-108 (min with i-1 first)
      pmin = fmin(pmin, pL);
      pmin = fmin(pmin, pR);
    }
    double gradP = fabs(pR - pL);
    if (ND >= 2) {
      pR = P(c + sj);
      pL = P(c - sj);
      if (prs >= 0) {
        pmin = fmin(pmin, pR);
        pmin = fmin(pmin, pL);
      } else {
        pmin = fmin(pmin, pL);
        pmin = fmin(pmin, pR);
      }
      gradP += fabs(pR - pL);
    }
    if (ND == 3) {
      pR = P(c + sk);
      pL = P(c - sk);
      if (prs >= 0) {
        pmin = fmin(pmin, pR);
        pmin = fmin(pmin, pL);
      } else {
        pmin = fmin(pmin, pL);
        pmin = fmin(pmin, pR);
      }
      gradP += fabs(pR - pL);
    }
    if (gradP > smoothing * pmin) flag = 1;
  }
  flags[c] = flag;
}

__global__ void fillKernel(double *p, long long n, double v) {
  const long long c = (long long)blockIdx.x * TB + threadIdx.x;
  if (c < n) p[c] = v;
}
__global__ void copyKernel(double *__restrict__ d, const double *__restrict__ s, long long n) {
  const long long c = (long long)blockIdx.x * TB + threadIdx.x;
  if (c < n) d[c] = s[c];
}
__global__ void addAndStoreKernel(double *__restrict__ d, const double *__restrict__ s, long long n, double wl,
                                  double wr) {
  const long long c = (long long)blockIdx.x * TB + threadIdx.x;
  if (c < n) d[c] = wl * d[c] + wr * s[c];
}

// ---- fused stage epilogue -----------------------------------------------------------------
// One pass over the INTERIOR cells (the fused path is only used when every ghost cell is
// refilled by the next SetBoundaries, i.e. no user-defined boundary):
//   stages >= 1: u <- wc*Uc + w0*Uc0                                  (StateContainer::AddAndStore)
//   u(B_ct) <- face average of Vs (already combined by evolveMagField)  (ReconstructVcField(Uc))
//   Vc <- ConsToPrim(u)                                                 (Fluid::ConvertConsToPrim)
// Uc is written back only when keepUc is set (it is rebuilt from Vc by the next stage anyway).
template <class P>
__global__ void stageEpilogueKernel(DevGrid g, double *__restrict__ Vc, double *__restrict__ Uc,
                                    const double *__restrict__ Vs, const double *__restrict__ Uc0, Eos eos,
                                    int combine, double wc, double w0, int keepUc, unsigned int *mathFlag, CellBox box,
                                    const double *__restrict__ invDt, unsigned long long *dtMinBits) {
  const int ni = box.hi[0] - box.lo[0], nj = box.hi[1] - box.lo[1], nk = box.hi[2] - box.lo[2];
  const long long t = (long long)blockIdx.x * TB + threadIdx.x;
  const bool valid = t < (long long)ni * nj * nk;
  double dtCell = __longlong_as_double(0x7ff0000000000000LL);   // +inf
  if (valid) {
    const int i = box.lo[0] + (int)(t % ni);
    const int j = box.lo[1] + (int)((t / ni) % nj);
    const int k = box.lo[2] + (int)(t / ((long long)ni * nj));
    const long long c = cidx(g, k, j, i);
    double u[P::nvar], v[P::nvar];
    if constexpr (P::mhd) {
      const long long s = sidx(g, k, j, i);
      u[P::BX1] = 0.5 * (Vs[s] + Vs[s + 1]);
      if constexpr (P::ndim >= 2) u[P::BX2] = 0.5 * (Vs[g.vsn + s] + Vs[g.vsn + s + g.vsj]);
      if constexpr (P::ndim == 3) u[P::BX3] = 0.5 * (Vs[2 * g.vsn + s] + Vs[2 * g.vsn + s + g.vsk]);
    }
#pragma unroll
    for (int nv = 0; nv < P::nvar; nv++) {
      if (P::mhd && nv >= P::BX1 && nv < P::BX1 + P::ndim) continue;
      double x = Uc[nv * g.sn + c];
      if (combine) x = wc * x + w0 * Uc0[nv * g.sn + c];
      u[nv] = x;
    }
    consToPrim<P>(v, u, eos);
    // the primitives the next sweeps start from must lie in the validated envelope of the branch-free
    // division / square root (fastmath.cuh); free here, this kernel is bandwidth-bound
    bool bad = false;
#pragma unroll
    for (int nv = 0; nv < P::nvar; nv++) {
      Vc[nv * g.sn + c] = v[nv];
      if (keepUc) Uc[nv * g.sn + c] = u[nv];
      bad |= !idfxStateInRange(v[nv]);
    }
    if (bad && mathFlag != nullptr) atomicAdd(mathFlag, 1u);
    // DataBlock::ComputeTimestep (dataBlock.cpp:333-364) folded in: the sweeps of this stage have finished
    // accumulating InvDt, its reciprocal is reduced here instead of in a pass of its own
    if (invDt != nullptr) dtCell = 1.0 / invDt[c];
  }
  if (invDt != nullptr) {
#ifndef IDFX_DT_CTA
    // Few cells lower the running minimum.  Every warp compares its cells with a cached copy of it -- possibly
    // stale, hence larger, because the minimum only decreases: a stale copy costs a superfluous atomic, never a
    // missed one -- and only warps that hold a lower value reduce and go to the atomic.  No barrier: a
    // __syncthreads() here keeps whole CTAs of this streaming kernel resident until their slowest load returns.
    // (One unconditional atomic per warp on a single address serialises in L2: measured +5 ms at 512^3.)
    const unsigned long long mine = (dtCell >= 0.0) ? (unsigned long long)__double_as_longlong(dtCell) : ~0ull;
    const unsigned long long seen = *dtMinBits;
    if (__any_sync(0xffffffffu, mine < seen)) {
      unsigned long long m = mine;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, m, o);
        m = other < m ? other : m;
      }
      if ((threadIdx.x & 31) == 0) atomicMin(dtMinBits, m);
    }
#else
    // warp shuffles, one shared-memory step per CTA, and an atomic only from CTAs that lower the running minimum
    __shared__ double warpMins[TB / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dtCell = fmin(dtCell, __shfl_xor_sync(0xffffffffu, dtCell, o));
    if ((threadIdx.x & 31) == 0) warpMins[threadIdx.x >> 5] = dtCell;
    __syncthreads();
    if (threadIdx.x == 0) {
      double m = warpMins[0];
#pragma unroll
      for (int w = 1; w < TB / 32; w++) m = fmin(m, warpMins[w]);
      if (m >= 0.0) {
        const unsigned long long bits = (unsigned long long)__double_as_longlong(m);
        if (bits < *reinterpret_cast<volatile unsigned long long *>(dtMinBits)) atomicMin(dtMinBits, bits);
      }
    }
#endif
  }
}

}  // namespace

int idfxk::primToCons(const PhysCfg &c, const DevGrid &g, const double *Vc, double *Uc, double gamma,
                      cudaStream_t st) {
  const long long n = g.sn;
  const Eos eos{gamma, 0.0};
  PHYS_DISPATCH((primToConsKernel<P><<<blocksFor(n), TB, 0, st>>>(g, Vc, Uc, eos, n)))
}

int idfxk::consToPrim(const PhysCfg &c, const DevGrid &g, double *Vc, double *Uc, double gamma, cudaStream_t st) {
  const long long n = g.sn;
  const Eos eos{gamma, 0.0};
  PHYS_DISPATCH((consToPrimKernel<P><<<blocksFor(n), TB, 0, st>>>(g, Vc, Uc, eos, n)))
}

int idfxk::reconstructVcField(const PhysCfg &c, const DevGrid &g, const double *Vs, double *V, cudaStream_t st,
                              int ghostsOnly) {
  auto launch = [&](const CellBox &b) {
    const int n0 = b.hi[0] - b.lo[0], n1 = b.hi[1] - b.lo[1], n2 = b.hi[2] - b.lo[2];
    if (n0 <= 0 || n1 <= 0 || n2 <= 0) return;
    const dim3 blk = (n0 <= 4) ? dim3(4, 64, 1) : dim3(64, 4, 1);
    const dim3 grd((n0 + blk.x - 1) / blk.x, (n1 + blk.y - 1) / blk.y, n2);
    reconstructVcFieldKernel<<<grd, blk, 0, st>>>(g, Vs, V, c.BX1, b);
  };
  CellBox all;
  for (int d = 0; d < 3; d++) {
    all.lo[d] = 0;
    all.hi[d] = g.nx[d];
  }
  if (!ghostsOnly) {
    launch(all);
  } else {
    // the ghost shell as six disjoint slabs: x3 slabs span everything, x2 slabs the active k range, x1 slabs
    // the active k and j ranges
    for (int d = 2; d >= 0; d--) {
      if (g.ng[d] == 0) continue;
      CellBox lo = all, hi = all;
      for (int e = d + 1; e < 3; e++) {
        lo.lo[e] = hi.lo[e] = g.beg[e];
        lo.hi[e] = hi.hi[e] = g.end[e];
      }
      lo.hi[d] = g.beg[d];
      hi.lo[d] = g.end[d];
      launch(lo);
      launch(hi);
    }
  }
  return (int)cudaGetLastError();
}

int idfxk::findShock(const PhysCfg &c, const DevGrid &g, const double *Vc, unsigned char *flags, double smoothing,
                     double csiso, cudaStream_t st) {
  CellBox b;
  for (int d = 0; d < 3; d++) {
    b.lo[d] = g.beg[d] - g.off[d];
    b.hi[d] = g.end[d] + g.off[d];
  }
  const dim3 blk(64, 4, 1);
  const dim3 grd((b.hi[0] - b.lo[0] + 63) / 64, (b.hi[1] - b.lo[1] + 3) / 4, b.hi[2] - b.lo[2]);
  const int prs = c.iso ? -1 : c.ENG;
  if (g.ndim == 1) findShockKernel<1><<<grd, blk, 0, st>>>(g, Vc, flags, c.MX1, prs, c.RHO, smoothing, csiso, b);
  else if (g.ndim == 2) findShockKernel<2><<<grd, blk, 0, st>>>(g, Vc, flags, c.MX1, prs, c.RHO, smoothing, csiso, b);
  else findShockKernel<3><<<grd, blk, 0, st>>>(g, Vc, flags, c.MX1, prs, c.RHO, smoothing, csiso, b);
  return (int)cudaGetLastError();
}

int idfxk::fill(double *p, long long n, double value, cudaStream_t st) {
  fillKernel<<<blocksFor(n), TB, 0, st>>>(p, n, value);
  return (int)cudaGetLastError();
}
int idfxk::copy(double *dst, const double *src, long long n, cudaStream_t st) {
  copyKernel<<<blocksFor(n), TB, 0, st>>>(dst, src, n);
  return (int)cudaGetLastError();
}
int idfxk::addAndStore(double *dst, const double *src, long long n, double wl, double wr, cudaStream_t st) {
  addAndStoreKernel<<<blocksFor(n), TB, 0, st>>>(dst, src, n, wl, wr);
  return (int)cudaGetLastError();
}

int idfxk::stageEpilogue(const PhysCfg &c, const DevGrid &g, double *Vc, double *Uc, const double *Vs,
                         const double *Uc0, double gamma, int combine, double wc, double w0, int keepUc,
                         cudaStream_t st, unsigned int *mathFlag, const int *lo, const int *hi, const double *invDt,
                         unsigned long long *dtMinBits) {
  CellBox box;
  for (int d = 0; d < 3; d++) {
    box.lo[d] = lo ? lo[d] : g.beg[d];
    box.hi[d] = hi ? hi[d] : g.end[d];
    if (box.hi[d] <= box.lo[d]) return 0;
  }
  const long long n = (long long)(box.hi[0] - box.lo[0]) * (box.hi[1] - box.lo[1]) * (box.hi[2] - box.lo[2]);
  const Eos eos{gamma, 0.0};
  PHYS_DISPATCH((stageEpilogueKernel<P><<<blocksFor(n), TB, 0, st>>>(g, Vc, Uc, Vs, Uc0, eos, combine, wc, w0,
                                                                      keepUc, mathFlag, box, invDt, dtMinBits)))
}
