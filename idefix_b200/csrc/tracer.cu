// tracer.cu -- passive tracers (reference: src/fluid/tracer/tracer.hpp:58-125, tracer.cpp:12-49).
//
// The reference advects tracers with separate kernels around the fluid update, and so do we: the
// directional sweep leaves the raw Riemann mass flux Flux(RHO) of every interface in a one-variable
// scratch array (SweepArgs::fluxRho, the only part of the flux array that ever exists in HBM, and
// only when tracers are on); tracerSweepKernel turns it into the upwinded tracer fluxes of the two
// interfaces of a cell and updates the conservative tracer in one pass.  Tracer variables live behind
// the fluid variables in Vc / Uc (index nvar .. nvar+ntracer-1), like in the reference.
#include "kernels.h"
#include "recon.cuh"

namespace {

constexpr int TBX = 64, TBY = 4;

// Tracer::CalcFlux: upwind PLM (van Leer, regular-grid formula) face value times the mass flux times the area
template <int DIR>
__device__ __forceinline__ double tracerFlux(const DevGrid &g, const double *__restrict__ V,
                                             const double *__restrict__ fluxRho, long long c, long long s, int k,
                                             int j, int i) {
  const double f = fluxRho[c];
  double vface;
  if (f > 0) {
    const double dvm = V[c - s] - V[c - 2 * s];
    const double dvp = V[c] - V[c - s];
    const double dv = lim::vanLeer(dvp, dvm);
    vface = V[c - s] + 0.5 * dv;
  } else {
    const double dvm = V[c] - V[c - s];
    const double dvp = V[c + s] - V[c];
    const double dv = lim::vanLeer(dvp, dvm);
    vface = V[c] - 0.5 * dv;
  }
  return f * vface * faceArea<DIR>(g, k, j, i);
}

// Tracer::CalcFlux + Tracer::CalcRightHandSide of one direction over the active cells
template <int DIR>
__global__ void tracerSweepKernel(DevGrid g, const double *__restrict__ Vtr, double *__restrict__ Utr,
                                  const double *__restrict__ fluxRho, const double *dtp, int ntr) {
  const int i = g.beg[0] + blockIdx.x * TBX + threadIdx.x;
  const int j = g.beg[1] + blockIdx.y * TBY + threadIdx.y;
  const int k = g.beg[2] + blockIdx.z;
  if (i >= g.end[0] || j >= g.end[1] || k >= g.end[2]) return;
  const double dt = *dtp;
  const long long c = cidx(g, k, j, i);
  const long long s = (DIR == IDIR) ? 1 : (DIR == JDIR ? g.sj : g.sk);
  const double dV = cellVolume(g, k, j, i);
  const int ip = i + (DIR == IDIR), jp = j + (DIR == JDIR), kp = k + (DIR == KDIR);
  for (int n = 0; n < ntr; n++) {
    const double *V = Vtr + n * g.sn;
    const double Fl = tracerFlux<DIR>(g, V, fluxRho, c, s, k, j, i);
    const double Fr = tracerFlux<DIR>(g, V, fluxRho, c + s, s, kp, jp, ip);
    Utr[n * g.sn + c] += -dt / dV * (Fr - Fl);
  }
}

// Tracer::ConvertPrimToCons over the whole block (+ optional copy into the "begin" state)
__global__ void tracerPrimToConsKernel(const double *__restrict__ Vc, double *__restrict__ Uc, double *__restrict__ Uc0,
                                       long long ncell, int rho, int first, int ntr) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  const double r = Vc[rho * ncell + c];
  for (int n = first; n < first + ntr; n++) {
    const double u = Vc[n * ncell + c] * r;
    Uc[n * ncell + c] = u;
    if (Uc0) Uc0[n * ncell + c] = u;
  }
}

// Tracer::ConvertConsToPrim over the whole block; combine != 0 first applies the Runge-Kutta combination
// u <- wc*u + w0*u0 (StateContainer::AddAndStore) of the fused path
__global__ void tracerConsToPrimKernel(double *__restrict__ Vc, double *__restrict__ Uc, const double *__restrict__ Uc0,
                                       long long ncell, int rho, int first, int ntr, int combine, double wc,
                                       double w0, DevGrid g, int interiorOnly) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  if (interiorOnly) {
    const int i = (int)(c % g.nx[0]);
    const int j = (int)((c / g.nx[0]) % g.nx[1]);
    const int k = (int)(c / ((long long)g.nx[0] * g.nx[1]));
    if (i < g.beg[0] || i >= g.end[0] || j < g.beg[1] || j >= g.end[1] || k < g.beg[2] || k >= g.end[2]) return;
  }
  const double r = Uc[rho * ncell + c];
  for (int n = first; n < first + ntr; n++) {
    double u = Uc[n * ncell + c];
    if (combine) {
      u = wc * u + w0 * Uc0[n * ncell + c];
      Uc[n * ncell + c] = u;
    }
    Vc[n * ncell + c] = u / r;
  }
}

}  // namespace

int idfxk::tracerSweep(const DevGrid &g, int dir, const double *Vtr, double *Utr, const double *fluxRho,
                       const double *dt, int ntr, cudaStream_t st) {
  const dim3 block(TBX, TBY, 1);
  const dim3 grid((g.end[0] - g.beg[0] + TBX - 1) / TBX, (g.end[1] - g.beg[1] + TBY - 1) / TBY, g.end[2] - g.beg[2]);
  if (dir == 0) tracerSweepKernel<0><<<grid, block, 0, st>>>(g, Vtr, Utr, fluxRho, dt, ntr);
  else if (dir == 1) tracerSweepKernel<1><<<grid, block, 0, st>>>(g, Vtr, Utr, fluxRho, dt, ntr);
  else tracerSweepKernel<2><<<grid, block, 0, st>>>(g, Vtr, Utr, fluxRho, dt, ntr);
  return (int)cudaGetLastError();
}

int idfxk::tracerPrimToCons(const PhysCfg &c, const DevGrid &g, const double *Vc, double *Uc, double *Uc0,
                            cudaStream_t st) {
  const long long n = g.sn;
  tracerPrimToConsKernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(Vc, Uc, Uc0, n, c.RHO, c.nvar, c.ntr);
  return (int)cudaGetLastError();
}

int idfxk::tracerConsToPrim(const PhysCfg &c, const DevGrid &g, double *Vc, double *Uc, const double *Uc0,
                            int combine, double wc, double w0, int interiorOnly, cudaStream_t st) {
  const long long n = g.sn;
  tracerConsToPrimKernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(Vc, Uc, Uc0, n, c.RHO, c.nvar, c.ntr, combine, wc,
                                                                      w0, g, interiorOnly);
  return (int)cudaGetLastError();
}
