// sweep_march.cuh -- directional sweep for the x2 / x3 directions as a marching ("pencil") kernel.
//
// Same fusion as sweep.cuh (K8 reconstruct + Riemann + EMF storage, K9 flux correction, K10
// right-hand side, K11 source terms in one kernel; reference: hlldMHD.hpp:147,
// calcRightHandSide.hpp:16-196,201-551, addSourceTerms.hpp:16-195) with a different work
// decomposition, chosen from the ncu capture of the tile kernel (profiles/r1a_sweep_ncu_full.md:
// 8 resident warps per SM, long-scoreboard stalls on the stencil loads, 12.5-25 % of the Riemann
// problems solved twice at tile overlaps):
//   * a thread owns one (i, perpendicular) column and marches along the sweep direction over a
//     chunk of interfaces; threads run along i, so every global access stays coalesced;
//   * every interface is solved exactly once (one extra per chunk); the previous interface's
//     area-weighted flux is kept on chip and differenced in the same thread, so the flux array
//     still never exists in HBM;
//   * the reconstruction stencil lives in a per-thread shared-memory ring that is fed by
//     cp.async one cell ahead of its use, together with the next cell's Uc / InvDt and the next
//     face's normal field: each primitive is read from HBM once per sweep and the load latency
//     overlaps the ~1.5k-instruction Riemann solve of the current interface without holding
//     registers; no __syncthreads is needed because a thread only reads what it copied itself;
//   * 128-thread CTAs, three resident per SM (12 warps instead of 8).
#ifndef IDFX_SWEEP_MARCH_CUH_
#define IDFX_SWEEP_MARCH_CUH_

#include "sweep.cuh"

namespace march {

IDFX_D unsigned smemAddr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
IDFX_D void cpAsync8(double *dstShared, const double *srcGlobal) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smemAddr(dstShared)), "l"(srcGlobal) : "memory");
}
IDFX_D void cpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
IDFX_D void cpAsyncWait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int ORDER>
struct Stencil {
  static constexpr int L = (ORDER == 4) ? 3 : 2;   // cells to the left of the interface
  static constexpr int R = (ORDER == 4) ? 3 : 2;   // cells to the right (cell m itself included)
  static constexpr int W = L + R + 1;              // ring slots: stencil + one cell in flight
};

}  // namespace march

// shared memory needed by one CTA of TX threads
template <class P, int ORDER>
constexpr size_t sweepMarchSmem(int TX) {
  return sizeof(double) * (size_t)TX *
         (march::Stencil<ORDER>::W * P::nvar   // stencil ring
          + 2 * (P::nvar + 1)                  // Uc + InvDt of the cell to update, double buffered
          + (P::nvar + 1)                      // previous interface: area-weighted flux + signal speed
          + 2);                                // normal field at the interface, double buffered
}

template <class P, int SOLVER, int ORDER, int DIR, int TX>
__global__ void __launch_bounds__(TX, (ORDER == 4 ? 2 : 3)) sweepMarchKernel(const SweepArgs a, const int chunk) {
  static_assert(DIR == JDIR || DIR == KDIR, "the marching kernel sweeps along x2 or x3");
  constexpr int NV = P::nvar;
  using D = Dirs<P, DIR>;
  using St = march::Stencil<ORDER>;
  constexpr int W = St::W;
  constexpr int PDIR = (DIR == JDIR) ? KDIR : JDIR;  // the other perpendicular direction

  extern __shared__ double smem[];
  const int tx = threadIdx.x;
  double *ring = smem + tx;                         // ring[(slot * NV + nv) * TX]
  double *ucb = smem + W * NV * TX + tx;            // ucb[(buf * (NV+1) + nv) * TX]
  double *fprev = ucb + 2 * (NV + 1) * TX;          // fprev[nv * TX]
  double *vsb = fprev + (NV + 1) * TX;              // vsb[buf * TX]

  const DevGrid &g = a.g;
  const int ext = a.perpExt;
  // ---- column owned by this thread ------------------------------------------------------------
  const int ilo = g.beg[0] - ext, ni = g.end[0] - g.beg[0] + 2 * ext;
  const int pext = (P::ndim > PDIR) ? ext : 0;
  const int plo = g.beg[PDIR] - pext, np = g.end[PDIR] - g.beg[PDIR] + 2 * pext;
  const long long lin = (long long)blockIdx.x * TX + tx;
  if (lin >= (long long)ni * np) return;
  const int i = ilo + (int)(lin % ni);
  const int p = plo + (int)(lin / ni);
  const int mb = g.beg[DIR] + blockIdx.y * chunk;                      // first interface of the chunk
  const int me = min(mb + chunk, g.end[DIR]);                          // last interface (cells mb..me-1)
  const long long stride = (DIR == JDIR) ? g.sj : g.sk;
  const long long vstride = (DIR == JDIR) ? g.vsj : g.vsk;
  // index of the cell / face with sweep coordinate 0
  const long long cell0 = (DIR == JDIR) ? cidx(g, p, 0, i) : cidx(g, 0, p, i);
  const long long face0 = DIR * g.vsn + ((DIR == JDIR) ? sidx(g, p, 0, i) : sidx(g, 0, p, i));
  const bool columnInterior = (i >= g.beg[0]) && (i < g.end[0]) && (p >= g.beg[PDIR]) && (p < g.end[PDIR]);

  const Eos eos{a.gamma, a.csiso};
  const double dt = *a.dt;
  const int geo = g.geometry;
  const bool regular = (g.regular != 0);
  const PlmWeights w{g.wp[DIR], g.wm[DIR], g.cp[DIR], g.cm[DIR], g.dp[DIR], g.dm[DIR]};
  const bool needUc = columnInterior && !a.fusedP2C;
  const bool needInvDt = columnInterior && a.doInvDt && !a.invDtInit;

  // ---- prologue: stencil of the first interface, its normal field -------------------------------
  // ring slot of cell m: (m - (mb - L)) mod W
#pragma unroll
  for (int c = 0; c < St::L + St::R; c++) {
    const long long cell = cell0 + (long long)(mb - St::L + c) * stride;
#pragma unroll
    for (int nv = 0; nv < NV; nv++) march::cpAsync8(ring + (c * NV + nv) * TX, a.Vc + nv * g.sn + cell);
  }
  if constexpr (P::mhd) march::cpAsync8(vsb, a.Vs + face0 + (long long)mb * vstride);
  march::cpAsyncCommit();

  int slot0 = 0;  // ring slot holding cell (m - L)
  for (int m = mb; m <= me; m++) {
    const int r = m - mb;
    // ---- prefetch what the next interface needs ------------------------------------------------
    if (m < me) {
      const int sl = (slot0 + St::L + St::R) % W;   // the free slot receives cell m + R
      const long long cell = cell0 + (long long)(m + St::R) * stride;
#pragma unroll
      for (int nv = 0; nv < NV; nv++) march::cpAsync8(ring + (sl * NV + nv) * TX, a.Vc + nv * g.sn + cell);
      if constexpr (P::mhd) march::cpAsync8(vsb + ((r + 1) & 1) * TX, a.Vs + face0 + (long long)(m + 1) * vstride);
      const long long cu = cell0 + (long long)m * stride;  // cell m is updated at the next interface
      double *ub = ucb + ((r + 1) & 1) * (NV + 1) * TX;
      if (needUc) {
#pragma unroll
        for (int nv = 0; nv < NV; nv++) {
          if (P::mhd && nv >= P::BX1 && nv < P::BX1 + P::ndim) continue;
          march::cpAsync8(ub + nv * TX, a.Uc + nv * g.sn + cu);
        }
      }
      if (needInvDt) march::cpAsync8(ub + NV * TX, a.invDt + cu);
    }
    march::cpAsyncCommit();
    march::cpAsyncWait<1>();   // everything but the group just committed has landed

    // ---- interface m: reconstruct, solve, hand the EMFs to CT ---------------------------------------
    double vL[NV], vR[NV];
#pragma unroll
    for (int nv = 0; nv < NV; nv++) {
      if (P::mhd && nv == D::BXn) continue;
      double st[6];
      st[0] = 0.0;
      st[5] = 0.0;
      // st[k] = cell (m - 3 + k); the ring holds cells m-L .. m+R-1 (+ cell m+R in flight)
#pragma unroll
      for (int k = 3 - St::L; k < 3 + St::R; k++) {
        const int sl = (slot0 + (k - (3 - St::L))) % W;
        st[k] = ring[(sl * NV + nv) * TX];
      }
      const bool positive = (nv == P::RHO) || (P::energy && nv == P::PRS);
      reconstructStencil<ORDER>(st, positive, regular, w, g.dx[DIR], m, vL[nv], vR[nv]);
    }
    if constexpr (P::mhd) {
      const double bn = vsb[(r & 1) * TX];
      vL[D::BXn] = bn;
      vR[D::BXn] = bn;
    }
    const int j = (DIR == JDIR) ? m : p;
    const int k = (DIR == JDIR) ? p : m;
    const long long cellm = cell0 + (long long)m * stride;

    double F[NV];
    RiemannAux aux;
    riemannDispatch<P, SOLVER, DIR>(vL, vR, eos, F, aux);
    if constexpr (P::mhd) {
      const EmfFaceArrays A = emfFaceArrays<P::ndim, DIR>(a.emf);
      storeEmf<P, DIR, SOLVER>(a.emfMode, A, cellm, F, aux, vL, vR);
    }
    areaWeightFlux<P, DIR>(g, geo, k, j, i, F);

    // ---- cell m-1: flux difference, time-step estimate, conservative update ---------------------------
    if (r > 0 && columnInterior) {
      const int jc = (DIR == JDIR) ? m - 1 : p;
      const int kc = (DIR == JDIR) ? p : m - 1;
      const long long cell = cellm - stride;
      double Fm[NV];
#pragma unroll
      for (int nv = 0; nv < NV; nv++) Fm[nv] = fprev[nv * TX];
      const double cmaxm = fprev[NV * TX];
      const double *ub = ucb + (r & 1) * (NV + 1) * TX;
      // primitives of cell m-1 are still in the ring (slot of cell m - 1 = slot0 + L - 1)
      const int slc = (slot0 + St::L - 1) % W;
      updateCell<P, DIR>(a, g, geo, eos, dt, kc, jc, i, m - 1, cell, Fm, F, cmaxm, aux.cmax,
                         [&](int nv) { return ring[(slc * NV + nv) * TX]; },
                         [&](int nv) { return ub[nv * TX]; }, [&]() { return ub[NV * TX]; });
    }
#pragma unroll
    for (int nv = 0; nv < NV; nv++) fprev[nv * TX] = F[nv];
    fprev[NV * TX] = aux.cmax;
    slot0 = (slot0 + 1) % W;
  }
}

template <class P, int SOLVER, int ORDER, int DIR>
inline cudaError_t launchSweepMarch(const SweepArgs &a, cudaStream_t stream) {
  constexpr int TX = 128;
  const DevGrid &g = a.g;
  constexpr int PDIR = (DIR == JDIR) ? KDIR : JDIR;
  const int ext = a.perpExt;
  const long long ni = g.end[0] - g.beg[0] + 2 * ext;
  const long long np = g.end[PDIR] - g.beg[PDIR] + 2 * ((P::ndim > PDIR) ? ext : 0);
  const int nsweep = g.end[DIR] - g.beg[DIR];
  const long long ctasPerLayer = (ni * np + TX - 1) / TX;
  // chunk length: long enough to amortise the extra interface per chunk, short enough to fill the GPU
  int chunk = 64;
  while (chunk > 8 && ctasPerLayer * ((nsweep + chunk - 1) / chunk) < 4 * 148 * 3) chunk >>= 1;
  if (chunk > nsweep) chunk = nsweep;
  const dim3 grid((unsigned)ctasPerLayer, (unsigned)((nsweep + chunk - 1) / chunk), 1);
  constexpr size_t smem = sweepMarchSmem<P, ORDER>(TX);
  static bool configured = false;
  auto kern = sweepMarchKernel<P, SOLVER, ORDER, DIR, TX>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  kern<<<grid, TX, smem, stream>>>(a, chunk);
  return cudaGetLastError();
}

#endif  // IDFX_SWEEP_MARCH_CUH_
