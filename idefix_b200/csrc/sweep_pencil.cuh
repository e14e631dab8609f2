// sweep_pencil.cuh -- the directional sweep of the Godunov update, one kernel per direction.
//
// Fuses what the reference runs as three kernels per direction with two DRAM round trips of the
// 8-variable flux array (K8 reconstruct + Riemann + EMF storage, e.g. hlldMHD.hpp:147; K9 flux
// correction, calcRightHandSide.hpp:16-196; K10 right-hand side, calcRightHandSide.hpp:201-551;
// K11 geometrical source terms, addSourceTerms.hpp:16-195).  Design points, each taken from an
// ncu capture of the previous kernels (profiles/r1a_*, profiles/r1b_*):
//   * threads always run along x1 (unit stride), so every global access is coalesced;
//   * the reconstruction is evaluated once per CELL (reconstructCell): the right-face value of a
//     cell is the left state of the next interface, so the slope limiter (one FP64 division per
//     variable) is paid once instead of twice per interface;
//   * every interface is solved once; the previous interface's flux stays on chip and the flux
//     array never exists in HBM;
//   * primitives arrive through cp.async into shared memory one step ahead of their use (x2 / x3:
//     a per-thread ring along the sweep; x1: a per-warp row buffer), each primitive is read from
//     HBM once per sweep and no __syncthreads is needed in the marching loop;
//   * the separable geometry factors a step needs are staged in shared memory / registers before
//     the loop instead of being fetched with dependent global loads at every interface;
//   * CART = true specialises cartesian uniform grids at compile time (no curvilinear branches).
#ifndef IDFX_SWEEP_PENCIL_CUH_
#define IDFX_SWEEP_PENCIL_CUH_

#include "sweep.cuh"
#include "tma.cuh"
#include "tmem.cuh"

// IDFX_TMEM = 1: the x2 / x3 marches keep their per-thread carry-over (left state of the next interface, flux and
// signal speed of the previous one: 17 doubles) in tensor memory instead of shared memory (tmem.cuh).  That takes
// 17 KB of the 45 KB of shared memory a CTA needs, which the SM turns into L1 for the register spills.
#ifndef IDFX_TMEM
#define IDFX_TMEM 1
#endif

// resident CTAs per SM the sweep kernels are compiled for (register cap = 65536 / (threads * MINB))
#ifndef IDFX_MARCH_MINB
#define IDFX_MARCH_MINB 4
#endif
#ifndef IDFX_ROW_MINB
#define IDFX_ROW_MINB 3
#endif

namespace pencil {

// Host-side per-DEVICE caches: one process may drive blocks on several GPUs (IdfxBlockDesc::device), and both the
// opt-in dynamic shared-memory size of a kernel and the SM count belong to the device that is current at the launch.
inline int currentDevice() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev;
}
inline int smCount() {
  static int sms[64] = {};
  const int dev = currentDevice() & 63;
  if (sms[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    sms[dev] = n > 0 ? n : 148;
  }
  return sms[dev];
}
template <class K>
inline cudaError_t ensureDynamicSmem(K kern, size_t bytes, unsigned long long &doneMask) {
  const int dev = currentDevice();
  if (dev < 64 && ((doneMask >> dev) & 1ull)) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess && dev < 64) doneMask |= (1ull << dev);
  return e;
}

IDFX_D unsigned smemAddr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
IDFX_D void cpAsync8(double *dstShared, const double *srcGlobal) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smemAddr(dstShared)), "l"(srcGlobal) : "memory");
}
// same with a precomputed 32-bit shared-memory address (saves the generic->shared conversion per copy)
IDFX_D void cpAsync8s(unsigned dstShared, const double *srcGlobal) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dstShared), "l"(srcGlobal) : "memory");
}
IDFX_D void cpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
IDFX_D void cpAsyncWait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int ORDER>
struct Halo {
  static constexpr int H = (ORDER == 4) ? 2 : 1;   // cells needed on each side of a cell to reconstruct it
  static constexpr int W = 2 * H + 2;              // ring slots: stencil of one cell + one cell in flight
  // LATE variant (cartesian marches): the next cell is prefetched right AFTER the reconstruction, into the
  // slot of the cell that just left the stencil; the Riemann solve that follows hides the latency
  static constexpr int WL = 2 * H + 1;
};

// rows of the per-CTA geometry table along the marching direction
enum : int { T_DX = 0, T_VOL = 1, T_AREA = 2, T_W = 3 /* wp wm cp cm dp dm */, T_ROWS_CART = 3, T_ROWS_GEN = 9 };

// Reconstruction-independent tail of an interface: tracer mass flux, face EMF storage, area weighting
template <class P, int SOLVER, int DIR, class VL, class VR>
IDFX_D void finishInterface(const SweepArgs &a, const DevGrid &g, int geo, double A0, int k, int j, int i,
                            long long cell, const VL &vL, const VR &vR, double *F, const RiemannAux &aux) {
  if (a.fluxRho != nullptr) a.fluxRho[cell] = F[P::RHO];   // for the tracer kernels (Tracer::CalcFlux)
  if constexpr (P::mhd) {
    const EmfFaceArrays A = emfFaceArrays<P::ndim, DIR>(a.emf);
    storeEmf<typename P::Fast, DIR, SOLVER>(a.emfMode, A, cell, F, aux, vL, vR);
  }
  areaWeightFlux<P, DIR>(g, geo, A0, k, j, i, F);
}

// Riemann problem of one interface with the branch-free divisions / square roots of fastmath.cuh
// (Phys<...>::Fast).  A failed range check is recorded in the per-thread flag and reported through
// SweepArgs::mathFlag at the end of the kernel (reportMathFlag below): when the flag of a run is zero --
// which the parity tests and bench.py assert -- every quotient and root was nvcc's own correctly rounded one.
template <class P, int SOLVER, int DIR>
IDFX_D bool solveInterface(const double *vL, const double *vR, const Eos &eos, double *F, RiemannAux &aux) {
  return riemannDispatch<typename P::Fast, SOLVER, DIR>(vL, vR, eos, F, aux);
}

// range validation of the branch-free arithmetic (fastmath.cuh): the primitives of a cell entering the sweeps
// and the fluxes leaving the solver
template <int NV, class V>
IDFX_D bool cellOutOfRange(V vcAt) {
  bool bad = false;
#pragma unroll
  for (int nv = 0; nv < NV; nv++) bad |= !idfxStateInRange(vcAt(nv));
  return bad;
}
// Shock flattening (extrapolateToFaces.hpp:98-104,141-147,182-189): in a flagged cell PLM and LimO3 use the
// reference's MinModLim, which always returns zero, i.e. face values q0 + 0 and q0 - 0; PPM ignores the flag
template <int ORDER>
IDFX_D void flattenCell(bool flat, double q0, double &vl, double &vr) {
  if constexpr (ORDER == 2 || ORDER == 3) {
    if (flat) {
      vr = q0 + 0.0;
      vl = q0 - 0.0;
    }
  }
}

template <int NV>
IDFX_D bool fluxNotFinite(const double *F, double cmax) {
  unsigned m = (unsigned)__double2hiint(cmax) & 0x7ff00000u;
  bool bad = (m == 0x7ff00000u);
#pragma unroll
  for (int nv = 0; nv < NV; nv++) bad |= !idfxFinite(F[nv]);
  return bad;
}
IDFX_D void reportMathFlag(const SweepArgs &a, bool bad) {
#if IDFX_FASTDIV
  if (bad) atomicAdd(a.mathFlag, 1u);
#endif
}

}  // namespace pencil

// =================================================================================================
// x2 / x3 sweeps: a thread owns one (i, perpendicular) column and marches along the sweep direction.
// =================================================================================================
template <class P, int ORDER, bool CART>
constexpr size_t pencilMarchSmem(int TX, int chunk) {
  return sizeof(double) * ((size_t)TX * ((CART ? pencil::Halo<ORDER>::WL : pencil::Halo<ORDER>::W) * P::nvar   // stencil ring
                                         + (IDFX_TMEM ? 0 : P::nvar)        // left state of the next interface
                                         + (IDFX_TMEM ? 0 : (P::nvar + 1))  // previous interface: flux + signal speed
                                         + 2)                               // normal field at the interface (x2)
                           + (size_t)(CART ? pencil::T_ROWS_CART : pencil::T_ROWS_GEN) * (chunk + 2)
                           + (IDFX_TMEM ? 2 : 0));                          // TMEM base address slot
}

// FLAT: the block has shock flattening on (a.shockFlags set).  A compile-time switch here, unlike in the x1 kernel:
// testing the flag per step costs the marches 5 % (measured, profiles/r2_ab_experiments.md) even when it is off.
template <class P, int SOLVER, int ORDER, int DIR, int TX, bool CART, bool FLAT>
__global__ void __launch_bounds__(TX, (ORDER == 4 ? 2 : IDFX_MARCH_MINB)) pencilMarchKernel(const SweepArgs a, const int chunk) {
  static_assert(DIR == JDIR || DIR == KDIR, "the marching kernel sweeps along x2 or x3");
  constexpr int NV = P::nvar;
  using D = Dirs<P, DIR>;
  constexpr bool LATE = CART;   // prefetch after the reconstruction (no cell behind the stencil is needed)
  constexpr int H = pencil::Halo<ORDER>::H, W = LATE ? pencil::Halo<ORDER>::WL : pencil::Halo<ORDER>::W;
  constexpr int PDIR = (DIR == JDIR) ? KDIR : JDIR;  // the other perpendicular direction

  extern __shared__ double smem[];
  const int tx = threadIdx.x;
  double *ring = smem + tx;                         // ring[(slot * NV + nv) * TX]
#if IDFX_TMEM
  // carry-over in tensor memory: columns [2 nv, 2 nv + 1] = left state nv, [2 NV + 2 nv, ..] = previous flux nv,
  // [4 NV, 4 NV + 1] = previous signal speed
  constexpr int TCOLS = (4 * NV + 2 <= 32) ? 32 : 64;
  constexpr unsigned T_VLS = 0, T_FPREV = 2 * NV, T_CMAX = 4 * NV;
  double *vsb = smem + W * NV * TX + tx;            // vsb[buf * TX]
  double *tab = smem + (size_t)TX * (W * NV + 2);   // tab[row * (chunk+2) + r]
#else
  double *vls = smem + W * NV * TX + tx;            // vls[nv * TX]
  double *fprev = vls + NV * TX;                    // fprev[nv * TX]
  double *vsb = fprev + (NV + 1) * TX;              // vsb[buf * TX]
  double *tab = smem + (size_t)TX * (W * NV + NV + (NV + 1) + 2);   // tab[row * (chunk+2) + r]
#endif
  const int tlen = chunk + 2;
  const unsigned ringS = pencil::smemAddr(ring), vsbS = pencil::smemAddr(vsb);

  const DevGrid &g = a.g;
  const int ext = a.perpExt;
  const int mb = g.beg[DIR] + blockIdx.y * chunk;                      // first interface of the chunk
  const int me = min(mb + chunk, g.end[DIR]);                          // last interface (cells mb..me-1)

  // ---- geometry along the march, entries for cells / faces mb-1 .. me -------------------------------
  for (int r = tx; r <= me - mb + 1; r += TX) {
    const int c = mb - 1 + r;
    tab[pencil::T_DX * tlen + r] = g.dx[DIR][c];
    tab[pencil::T_VOL * tlen + r] = g.vol[DIR][c];
    tab[pencil::T_AREA * tlen + r] = g.area[DIR][DIR][c];
    if constexpr (!CART) {
      tab[(pencil::T_W + 0) * tlen + r] = g.wp[DIR][c];
      tab[(pencil::T_W + 1) * tlen + r] = g.wm[DIR][c];
      tab[(pencil::T_W + 2) * tlen + r] = g.cp[DIR][c];
      tab[(pencil::T_W + 3) * tlen + r] = g.cm[DIR][c];
      tab[(pencil::T_W + 4) * tlen + r] = g.dp[DIR][c];
      tab[(pencil::T_W + 5) * tlen + r] = g.dm[DIR][c];
    }
  }
#if IDFX_TMEM
  unsigned *tslot = reinterpret_cast<unsigned *>(tab + (size_t)(CART ? pencil::T_ROWS_CART : pencil::T_ROWS_GEN) * tlen);
  const unsigned tbase = tmem::allocate<TCOLS>(tslot);   // contains the __syncthreads that publishes tab
#else
  __syncthreads();
#endif

  // ---- column owned by this thread ------------------------------------------------------------------
  const int ilo = g.beg[0] - ext, ni = g.end[0] - g.beg[0] + 2 * ext;
  const int pext = (P::ndim > PDIR) ? ext : 0;
  const int plo = g.beg[PDIR] - pext, np = g.end[PDIR] - g.beg[PDIR] + 2 * pext;
  long long lin = (long long)blockIdx.x * TX + tx;
#if IDFX_TMEM
  // tcgen05.ld / st are warp-collective: threads past the last column stay in the march as copies of the last
  // column (same loads, same face-EMF values to the same addresses) and only skip the cell update
  const bool columnValid = lin < (long long)ni * np;
  if (!columnValid) lin = (long long)ni * np - 1;
#else
  constexpr bool columnValid = true;
  if (lin >= (long long)ni * np) return;
#endif
  const int i = ilo + (int)(lin % ni);
  const int p = plo + (int)(lin / ni);
  const long long stride = (DIR == JDIR) ? g.sj : g.sk;
  const long long vstride = (DIR == JDIR) ? g.vsj : g.vsk;
  const long long cell0 = (DIR == JDIR) ? cidx(g, p, 0, i) : cidx(g, 0, p, i);
  const long long face0 = DIR * g.vsn + ((DIR == JDIR) ? sidx(g, p, 0, i) : sidx(g, 0, p, i));
  const bool columnInterior = columnValid && (i >= g.beg[0]) && (i < g.end[0]) && (p >= g.beg[PDIR]) && (p < g.end[PDIR]);

  const Eos eos{a.gamma, a.csiso};
  const double dt = *a.dt;
  const int geo = CART ? (int)GEO_CARTESIAN : g.geometry;
  const bool regular = CART ? true : (g.regular != 0);
  // separable geometry factors of this column, multiplied at every step in the reference's order
  // dV = (vol0[i]*vol1[j])*vol2[k], A = (area[D][0][i]*area[D][1][j])*area[D][2][k] (makeGeometry.cpp:66-84,139-211)
  double inv0, inv1 = 1.0, inv2, inv3 = 1.0;
  if constexpr (DIR == JDIR) {
    inv0 = g.vol[0][i];
    inv1 = g.vol[2][p];
    inv2 = g.area[DIR][0][i];
    inv3 = g.area[DIR][2][p];
  } else {
    inv0 = g.vol[0][i] * g.vol[1][p];
    inv2 = g.area[DIR][0][i] * g.area[DIR][1][p];
  }

  // ---- prologue: cells mb-1-H .. mb+H, the normal field of the first interface ---------------------------
  {
    const double *src = a.Vc + cell0 + (long long)(mb - 1 - H) * stride;
#pragma unroll
    for (int c = 0; c < W; c++) {
#pragma unroll
      for (int nv = 0; nv < NV; nv++) pencil::cpAsync8s(ringS + (c * NV + nv) * TX * 8, src + nv * g.sn);
      src += stride;
    }
  }
  if constexpr (P::mhd) pencil::cpAsync8s(vsbS, a.Vs + face0 + (long long)mb * vstride);
  pencil::cpAsyncCommit();
  pencil::cpAsyncWait<0>();
  // right-face values of cell mb-1 = left state of the first interface
  {
    const double dxc = tab[pencil::T_DX * tlen];
    double wc[6] = {0, 0, 0, 0, 0, 0};
    if constexpr (!CART) {
#pragma unroll
      for (int n = 0; n < 6; n++) wc[n] = tab[(pencil::T_W + n) * tlen];
    }
#pragma unroll
    for (int nv = 0; nv < NV; nv++) {
      if (P::mhd && nv == D::BXn) continue;
      double st[5];
      st[0] = st[4] = 0.0;
#pragma unroll
      for (int d = -H; d <= H; d++) st[2 + d] = ring[((H + d) * NV + nv) * TX];
      const bool positive = (nv == P::RHO) || (P::energy && nv == P::PRS);
      double vl, vr;
      reconstructCell<ORDER>(st, positive, regular, wc, dxc, vl, vr);
      if constexpr (FLAT) pencil::flattenCell<ORDER>(a.shockFlags[cell0 + (long long)(mb - 1) * stride] != 0, st[2], vl, vr);
#if IDFX_TMEM
      __syncwarp();
      tmem::store1(tbase + T_VLS + 2 * nv, vr);
#else
      vls[nv * TX] = vr;
#endif
    }
  }

  bool mathBad = false;
  int q = 0;  // ring slot of cell m-H-1: free, receives cell m+H+1 (LATE: cell m+H, issued at the end of step m-1)
  const double *pre = a.Vc + cell0 + (long long)(mb + H + (LATE ? 0 : 1)) * stride;
  long long cellm = cell0 + (long long)mb * stride;
  if constexpr (LATE) {
    // cell mb+H completes the stencil of cell mb
#pragma unroll
    for (int nv = 0; nv < NV; nv++) pencil::cpAsync8s(ringS + (q * NV + nv) * TX * 8, pre + nv * g.sn);
    pencil::cpAsyncCommit();
    pre += stride;
  }
  for (int m = mb; m <= me; m++) {
    const int r = m - mb;
    if constexpr (!LATE) {
      // ---- prefetch what the next interface needs ----------------------------------------------------
      if (m < me) {
#pragma unroll
        for (int nv = 0; nv < NV; nv++) pencil::cpAsync8s(ringS + (q * NV + nv) * TX * 8, pre + nv * g.sn);
        if constexpr (P::mhd) pencil::cpAsync8s(vsbS + ((r + 1) & 1) * TX * 8, a.Vs + face0 + (long long)(m + 1) * vstride);
      }
      pencil::cpAsyncCommit();
      pencil::cpAsyncWait<1>();   // everything but the group just committed has landed
    } else {
      pencil::cpAsyncWait<0>();   // cell m+H and the normal field of interface m (issued one step ago)
    }

    // ---- cell m: face values; interface m: left state from cell m-1, right state from cell m -----------
    int sc = q + 1 + H;         // slot of cell m
    if (sc >= W) sc -= W;
    double vL[NV], vR[NV];
    {
      const double dxc = tab[pencil::T_DX * tlen + r + 1];
      double wc[6] = {0, 0, 0, 0, 0, 0};
      if constexpr (!CART) {
#pragma unroll
        for (int n = 0; n < 6; n++) wc[n] = tab[(pencil::T_W + n) * tlen + r + 1];
      }
      int so[2 * H + 1];
#pragma unroll
      for (int d = -H; d <= H; d++) {
        int s = q + 1 + H + d;
        if (s >= W) s -= W;
        so[d + H] = s * NV * TX;
      }
      bool flat = false;
      if constexpr (FLAT) flat = a.shockFlags[cellm] != 0;
#if IDFX_TMEM
      // left state = right-face values of cell m-1, parked in tensor memory one step ago
      tmem::Pending1 vlp[NV];
      __syncwarp();
      tmem::waitStores();
#pragma unroll
      for (int nv = 0; nv < NV; nv++) {
        if (P::mhd && nv == D::BXn) continue;
        vlp[nv] = tmem::load1(tbase + T_VLS + 2 * nv);
      }
      tmem::waitLoads();
#pragma unroll
      for (int nv = 0; nv < NV; nv++) {
        if (P::mhd && nv == D::BXn) continue;
        vL[nv] = vlp[nv].value();
      }
#endif
#pragma unroll
      for (int nv = 0; nv < NV; nv++) {
        if (P::mhd && nv == D::BXn) continue;
        double st[5];
        st[0] = st[4] = 0.0;
#pragma unroll
        for (int d = -H; d <= H; d++) st[2 + d] = ring[so[d + H] + nv * TX];
        const bool positive = (nv == P::RHO) || (P::energy && nv == P::PRS);
        double vr;
        reconstructCell<ORDER, true>(st, positive, regular, wc, dxc, vR[nv], vr);
        pencil::flattenCell<ORDER>(flat, st[2], vR[nv], vr);
#if IDFX_TMEM
        __syncwarp();
        tmem::store1(tbase + T_VLS + 2 * nv, vr);
#else
        vL[nv] = vls[nv * TX];
        vls[nv * TX] = vr;
#endif
      }
    }
    if constexpr (LATE) {
      // cell m-H has left the stencil: its slot receives cell m+H+1 while the Riemann problem is solved
      if (m < me) {
        int qn = q + 1;
        if (qn >= W) qn -= W;
#pragma unroll
        for (int nv = 0; nv < NV; nv++) pencil::cpAsync8s(ringS + (qn * NV + nv) * TX * 8, pre + nv * g.sn);
        if constexpr (P::mhd) pencil::cpAsync8s(vsbS + ((r + 1) & 1) * TX * 8, a.Vs + face0 + (long long)(m + 1) * vstride);
      }
      pencil::cpAsyncCommit();
    }
    if constexpr (P::mhd) {
      const double bn = vsb[(r & 1) * TX];
      vL[D::BXn] = bn;
      vR[D::BXn] = bn;
    }
    const int j = (DIR == JDIR) ? m : p;
    const int k = (DIR == JDIR) ? p : m;
    double A0;
    if constexpr (DIR == JDIR) A0 = inv2 * tab[pencil::T_AREA * tlen + r + 1] * inv3;
    else A0 = inv2 * tab[pencil::T_AREA * tlen + r + 1];

    double F[NV];
    RiemannAux aux;
    pencil::solveInterface<P, SOLVER, DIR>(vL, vR, eos, F, aux);
    mathBad |= pencil::fluxNotFinite<NV>(F, aux.cmax);
    pencil::finishInterface<P, SOLVER, DIR>(a, g, geo, A0, k, j, i, cellm, vL, vR, F, aux);

    // ---- cell m-1: flux difference, time-step estimate, conservative update ---------------------------
#if IDFX_TMEM
    // flux and signal speed of interface m-1 come back from tensor memory (collective: outside the branch)
    double Fm[NV], cmaxm;
    {
      __syncwarp();
      tmem::Pending1 fp[NV + 1];
#pragma unroll
      for (int nv = 0; nv <= NV; nv++) fp[nv] = tmem::load1(tbase + T_FPREV + 2 * nv);
      tmem::waitLoads();
#pragma unroll
      for (int nv = 0; nv < NV; nv++) Fm[nv] = fp[nv].value();
      cmaxm = fp[NV].value();
    }
#endif
    if (r > 0 && columnInterior) {
      const int jc = (DIR == JDIR) ? m - 1 : p;
      const int kc = (DIR == JDIR) ? p : m - 1;
      double dV;
      if constexpr (DIR == JDIR) dV = inv0 * tab[pencil::T_VOL * tlen + r] * inv1;
      else dV = inv0 * tab[pencil::T_VOL * tlen + r];
#if !IDFX_TMEM
      double Fm[NV];
#pragma unroll
      for (int nv = 0; nv < NV; nv++) Fm[nv] = fprev[nv * TX];
      const double cmaxm = fprev[NV * TX];
#endif
      int sp = sc - 1;          // slot of cell m-1 (its primitives are still in the ring)
      if (sp < 0) sp += W;
      const double *vprev = ring + sp * NV * TX;
      updateCell<typename P::Fast, DIR>(a, g, geo, eos, dt, dV, tab[pencil::T_DX * tlen + r], kc, jc, i, cellm - stride, Fm, F,
                         cmaxm, aux.cmax, [&](int nv) { return vprev[nv * TX]; });
    }
#if IDFX_TMEM
    __syncwarp();
#pragma unroll
    for (int nv = 0; nv < NV; nv++) tmem::store1(tbase + T_FPREV + 2 * nv, F[nv]);
    tmem::store1(tbase + T_CMAX, aux.cmax);
#else
#pragma unroll
    for (int nv = 0; nv < NV; nv++) fprev[nv * TX] = F[nv];
    fprev[NV * TX] = aux.cmax;
#endif
    q = (q + 1 == W) ? 0 : q + 1;
    pre += stride;
    cellm += stride;
  }
  pencil::reportMathFlag(a, mathBad && columnValid);
#if IDFX_TMEM
  tmem::waitStores();
  tmem::release<TCOLS>(tslot);
#endif
}

template <class P, int SOLVER, int ORDER, int DIR, bool CART, bool FLAT = false>
inline cudaError_t launchPencilMarch(const SweepArgs &a, cudaStream_t stream) {
  if constexpr (!FLAT && (ORDER == 2 || ORDER == 3)) {   // PPM and first order ignore the flags (flattenCell)
    if (a.shockFlags != nullptr) return launchPencilMarch<P, SOLVER, ORDER, DIR, CART, true>(a, stream);
  }
  constexpr int TX = 128;
  const DevGrid &g = a.g;
  constexpr int PDIR = (DIR == JDIR) ? KDIR : JDIR;
  const int ext = a.perpExt;
  const long long ni = g.end[0] - g.beg[0] + 2 * ext;
  const long long np = g.end[PDIR] - g.beg[PDIR] + 2 * ((P::ndim > PDIR) ? ext : 0);
  const int nsweep = g.end[DIR] - g.beg[DIR];
  const long long ctasPerLayer = (ni * np + TX - 1) / TX;
  // chunk length: long enough to amortise the extra cell per chunk, short enough to fill the GPU
  int chunk = 64;
  while (chunk > 8 && ctasPerLayer * ((nsweep + chunk - 1) / chunk) < 4 * pencil::smCount() * 4) chunk >>= 1;
  if (chunk > nsweep) chunk = nsweep;
  const dim3 grid((unsigned)ctasPerLayer, (unsigned)((nsweep + chunk - 1) / chunk), 1);
  const size_t smem = pencilMarchSmem<P, ORDER, CART>(TX, 64);
  static unsigned long long configured = 0;   // bit d: attribute set on device d
  auto kern = pencilMarchKernel<P, SOLVER, ORDER, DIR, TX, CART, FLAT>;
  if (cudaError_t e = pencil::ensureDynamicSmem(kern, smem, configured); e != cudaSuccess) return e;
  kern<<<grid, TX, pencilMarchSmem<P, ORDER, CART>(TX, chunk), stream>>>(a, chunk);
  return cudaGetLastError();
}

// =================================================================================================
// x1 sweep: a warp owns a strip of 32 consecutive cells of one x1 row (30 of them updated, the two
// outer lanes provide the neighbouring face value / flux) and marches over the rows of an x2 chunk.
// Lane l holds cell c = cs - 1 + l; it reconstructs cell c, solves the interface on the LEFT face of
// c with the right-face value of lane l-1, and updates cell c with the flux of lane l+1.
// =================================================================================================
// one row buffer = one TMA box [nvar][32 + 2H cells], padded to a multiple of 128 bytes (TMA destination alignment)
template <class P, int ORDER>
constexpr int pencilRowBufferDoubles() {
  return (P::nvar * (32 + 2 * pencil::Halo<ORDER>::H) + 15) / 16 * 16;
}
template <class P, int ORDER>
constexpr int pencilRowDoublesPerWarp() {
  return 2 * pencilRowBufferDoubles<P, ORDER>()   // two row buffers of primitives (current row, next row in flight)
         + P::nvar * 32               // right-face values
         + (P::nvar + 1) * 32         // fluxes + signal speed
         + 2 * 32;                    // normal field, double buffered
}

template <class P, int ORDER, bool CART>
constexpr size_t pencilRowSmem(int warps, int chunk) {
  return sizeof(double) * ((size_t)warps * pencilRowDoublesPerWarp<P, ORDER>() + 2 * (size_t)chunk + 2 * (size_t)warps);
}

template <class P, int SOLVER, int ORDER, int WARPS, bool CART>
__global__ void __launch_bounds__(WARPS * 32, (ORDER == 4 ? 2 : IDFX_ROW_MINB))
    pencilRowKernel(const SweepArgs a, const int chunk, const int nstrips, const __grid_constant__ CUtensorMap vcBoxes) {
  constexpr int DIR = IDIR;
  constexpr int NV = P::nvar;
  using D = Dirs<P, DIR>;
  constexpr int H = pencil::Halo<ORDER>::H;
  constexpr int RW = 32 + 2 * H;                      // cells per row buffer
  constexpr int STRIP = 30;                           // cells updated per warp and row

  constexpr int BUF = pencilRowBufferDoubles<P, ORDER>();   // doubles between the two row buffers
  extern __shared__ __align__(128) double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double *wbase = smem + (size_t)warp * pencilRowDoublesPerWarp<P, ORDER>();
  double *rows = wbase;                               // rows[buf * BUF + nv * RW + x], x = cell - (cs-1-H)
  double *vrs = rows + 2 * BUF;                       // vrs[nv * 32 + lane]
  double *fls = vrs + NV * 32;                        // fls[nv * 32 + lane]
  double *vsb = fls + (NV + 1) * 32;                  // vsb[buf * 32 + lane]
  double *tab = smem + (size_t)WARPS * pencilRowDoublesPerWarp<P, ORDER>();   // tab[row * chunk + r]
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(tab + 2 * chunk) + 2 * warp;   // one per row buffer
  const unsigned barS = pencil::smemAddr(bars);
  const bool useTma = (a.useTma != 0);
  if (lane == 0) {
    tma::barrierInit(barS, 1);
    tma::barrierInit(barS + 8, 1);
    tma::fenceBarrierInit();
  }

  const DevGrid &g = a.g;
  const int ext = a.perpExt;
  const int jext = (P::ndim > 1) ? ext : 0, kext = (P::ndim > 2) ? ext : 0;
  const int jlo = g.beg[1] - jext, jhi = g.end[1] + jext;
  const int jb = jlo + blockIdx.y * chunk;            // first row of the chunk
  const int je = min(jb + chunk, jhi);                // one past the last row

  // ---- geometry along the march (x2) -----------------------------------------------------------------
  for (int r = threadIdx.x; r < je - jb; r += WARPS * 32) {
    tab[r] = g.vol[1][jb + r];
    tab[chunk + r] = g.area[DIR][1][jb + r];
  }
  __syncthreads();

  // ---- strip owned by this warp ----------------------------------------------------------------------
  const long long wlin = (long long)blockIdx.x * WARPS + warp;
  const int nk = g.end[2] - g.beg[2] + 2 * kext;
  if (wlin >= (long long)nstrips * nk) return;
  const int s = (int)(wlin % nstrips);
  const int k = g.beg[2] - kext + (int)(wlin / nstrips);
  const int cs = g.beg[0] + s * STRIP;                // first cell this warp updates
  const int c = cs - 1 + lane;                        // cell (and left-face interface) of this lane
  const int nx0 = g.nx[0];
  const bool cellLoad = (c < nx0);                    // c >= 0 always (beg[0] >= H + 1 when ORDER > 1)
  const bool faceLane = (lane >= 1) && (c <= g.end[0]);           // solves interface c
  const bool cellLane = (lane >= 1) && (lane <= STRIP) && (c < g.end[0]);   // updates cell c
  const bool kInterior = (k >= g.beg[2]) && (k < g.end[2]);

  const Eos eos{a.gamma, a.csiso};
  const double dt = *a.dt;
  const int geo = CART ? (int)GEO_CARTESIAN : g.geometry;
  const bool regular = CART ? true : (g.regular != 0);
  // lane-invariant geometry (same association as the reference: (f0[i]*f1[j])*f2[k])
  const int cg = cellLoad ? c : nx0 - 1;
  const double vol0 = g.vol[0][cg], vol2 = g.vol[2][k];
  const double ar0 = g.area[DIR][0][cg], ar2 = g.area[DIR][2][k];
  const double dxc = g.dx[0][cg];
  double wc[6] = {0, 0, 0, 0, 0, 0};
  if constexpr (!CART) {
    wc[0] = g.wp[0][cg]; wc[1] = g.wm[0][cg]; wc[2] = g.cp[0][cg];
    wc[3] = g.cm[0][cg]; wc[4] = g.dp[0][cg]; wc[5] = g.dm[0][cg];
  }

  // x offsets of the row-buffer entries this lane copies: cell cs-1-H+lane, and (lanes < 2H) cell cs+31-H+lane
  const int x0 = cs - 1 - H;                          // cell index of row-buffer entry 0
  const bool ld0 = (x0 + lane >= 0) && (x0 + lane < nx0);
  const bool ld1 = (lane < 2 * H) && (x0 + 32 + lane < nx0);

  const unsigned rowsS = pencil::smemAddr(rows), vsbS = pencil::smemAddr(vsb);
  unsigned phases = 0;   // bit b: parity of the next completion of row buffer b (TMA path)
  auto issueRow = [&](int j, int buf) {
    if (useTma) {
      // one box [nvar][RW cells] per warp and row; cells outside the array are zero-filled and never used
      if (lane == 0) {
        tma::barrierExpect(barS + 8 * buf, NV * RW * 8);
        tma::load4D(rowsS + buf * BUF * 8, &vcBoxes, barS + 8 * buf, x0, j, k, 0);
      }
    } else {
      const long long rowBase = cidx(g, k, j, 0);
      const unsigned dst = rowsS + (buf * BUF + lane) * 8;
      const double *src = a.Vc + rowBase + x0 + lane;
#pragma unroll
      for (int nv = 0; nv < NV; nv++) {   // all of them: the update rebuilds Uc from the cell's primitives
        if (ld0) pencil::cpAsync8s(dst + nv * RW * 8, src + nv * g.sn);
        if (ld1) pencil::cpAsync8s(dst + (nv * RW + 32) * 8, src + nv * g.sn + 32);
      }
    }
    if constexpr (P::mhd) {
      if (faceLane) pencil::cpAsync8s(vsbS + (buf * 32 + lane) * 8, a.Vs + DIR * g.vsn + sidx(g, k, j, c));
    }
  };
  auto rowInterior = [&](int j) { return kInterior && (j >= g.beg[1]) && (j < g.end[1]); };

  issueRow(jb, 0);
  pencil::cpAsyncCommit();
  bool mathBad = false;

  for (int j = jb; j < je; j++) {
    const int r = j - jb;
    const int buf = r & 1;
    if (j + 1 < je) issueRow(j + 1, buf ^ 1);
    pencil::cpAsyncCommit();
    pencil::cpAsyncWait<1>();
    if (useTma) {
      tma::barrierWait(barS + 8 * buf, (phases >> buf) & 1u);
      phases ^= (1u << buf);
    }
    __syncwarp();   // the row buffer was filled by all lanes of the warp

    // ---- cell c: face values --------------------------------------------------------------------------
    const double *row = rows + buf * BUF + H + lane;   // row[nv*RW + d] = cell c + d
    double vL[NV], vR[NV];
    const bool flat = (a.shockFlags != nullptr) && cellLoad && (a.shockFlags[cidx(g, k, j, c)] != 0);
#pragma unroll
    for (int nv = 0; nv < NV; nv++) {
      if (P::mhd && nv == D::BXn) continue;
      double st[5];
      st[0] = st[4] = 0.0;
#pragma unroll
      for (int d = -H; d <= H; d++) st[2 + d] = row[nv * RW + d];
      const bool positive = (nv == P::RHO) || (P::energy && nv == P::PRS);
      double vr;
      reconstructCell<ORDER, true>(st, positive, regular, wc, dxc, vR[nv], vr);
      pencil::flattenCell<ORDER>(flat, st[2], vR[nv], vr);
      vrs[nv * 32 + lane] = vr;
    }
    __syncwarp();

    // ---- every primitive of the cells this lane owns must lie in the validated envelope (fastmath.cuh) ----------
    // (only when the state was not produced by our own stage epilogue, which validates what it writes)
#if IDFX_FASTDIV
    if (a.validateInputs && ((lane >= 1 && lane <= STRIP) || (lane == 0 && s == 0) || (lane == 31 && s == nstrips - 1)) &&
        c <= g.end[0])
      mathBad |= pencil::cellOutOfRange<NV>([&](int nv) { return row[nv * RW]; });
#endif

    // ---- cell-centred EMFs of cell c (CalcCellCenteredEMF, calcCornerEmf.hpp:85-126) ------------------------
    // uct_contact averages them onto the corners from the cells [beg-1, end] of every direction; the
    // primitives are already on chip here, so the separate full-array kernel (6 reads + 3 writes per cell) is
    // folded into this sweep.  Each cell is written by exactly one lane: the strip's own 30 cells, plus the
    // cell in front of the first strip and the one behind the last.  (uct0 also reads beg-2: it keeps the
    // stand-alone kernel.)
#ifndef IDFX_ROW_NOFOLD
    if constexpr (P::mhd && P::ndim >= 2) {
      if (a.emfMode == EMF_UCT_CONTACT) {
        const bool owner = (lane >= 1 && lane <= STRIP) || (lane == 0 && s == 0) || (lane == 31 && s == nstrips - 1);
        if (owner && c <= g.end[0]) {
          const long long cc = cidx(g, k, j, c);
          const double v1 = row[P::MX1 * RW], v2 = row[P::MX2 * RW], B1 = row[P::BX1 * RW], B2 = row[P::BX2 * RW];
          if constexpr (P::ndim == 3) {
            const double v3 = row[P::MX3 * RW], B3 = row[P::BX3 * RW];
            a.emf.Ex1[cc] = (v3 * B2 - v2 * B3);
            a.emf.Ex2[cc] = (v1 * B3 - v3 * B1);
          }
          a.emf.Ex3[cc] = (v2 * B1 - v1 * B2);
        }
      }
    }
#endif

    // ---- interface c (left face of cell c) -----------------------------------------------------------------
    double F[NV];
    RiemannAux aux;
    aux.cmax = 0.0;
    const long long cell = cidx(g, k, j, c);
    if (faceLane) {
#pragma unroll
      for (int nv = 0; nv < NV; nv++) {
        if (P::mhd && nv == D::BXn) continue;
        vL[nv] = vrs[nv * 32 + lane - 1];
      }
      if constexpr (P::mhd) {
        const double bn = vsb[buf * 32 + lane];
        vL[D::BXn] = bn;
        vR[D::BXn] = bn;
      }
      const double A0 = ar0 * tab[chunk + r] * ar2;
      const bool keep = pencil::solveInterface<P, SOLVER, DIR>(vL, vR, eos, F, aux);
      // 1-D HD Roe: the flux array of the reference persists between calls and a strong shock leaves the entry of
      // the previous call (already area-weighted) in place, which CalcRightHandSide then weights again
      constexpr bool STALE = (!P::mhd && P::ndim == 1 && SOLVER == SOLVER_ROE);
      if constexpr (STALE) {
        if (keep) {
#pragma unroll
          for (int nv = 0; nv < NV; nv++) F[nv] = a.staleFlux[nv * nx0 + c];
        }
      }
      mathBad |= pencil::fluxNotFinite<NV>(F, aux.cmax);
      pencil::finishInterface<P, SOLVER, DIR>(a, g, geo, A0, k, j, c, cell, vL, vR, F, aux);
      if constexpr (STALE) {
#pragma unroll
        for (int nv = 0; nv < NV; nv++) a.staleFlux[nv * nx0 + c] = F[nv];
      }
#pragma unroll
      for (int nv = 0; nv < NV; nv++) fls[nv * 32 + lane] = F[nv];
      fls[NV * 32 + lane] = aux.cmax;
    }
    __syncwarp();

    // ---- cell c: flux difference, time-step estimate, conservative update ---------------------------------
    if (cellLane && rowInterior(j)) {
      double Fp[NV];
#pragma unroll
      for (int nv = 0; nv < NV; nv++) Fp[nv] = fls[nv * 32 + lane + 1];
      const double dV = vol0 * tab[r] * vol2;
      updateCell<typename P::Fast, DIR>(a, g, geo, eos, dt, dV, dxc, k, j, c, cell, F, Fp, aux.cmax, fls[NV * 32 + lane + 1],
                         [&](int nv) { return row[nv * RW]; });
    }
    __syncwarp();   // vrs / fls are rewritten by the next row
  }
  pencil::reportMathFlag(a, mathBad);
}

template <class P, int SOLVER, int ORDER, bool CART>
inline cudaError_t launchPencilRow(const SweepArgs &a, cudaStream_t stream) {
  constexpr int WARPS = 4;
  const DevGrid &g = a.g;
  const int ext = a.perpExt;
  const int n0 = g.end[0] - g.beg[0];
  const int nstrips = (n0 + 29) / 30;
  const int nj = g.end[1] - g.beg[1] + 2 * ((P::ndim > 1) ? ext : 0);
  const int nk = g.end[2] - g.beg[2] + 2 * ((P::ndim > 2) ? ext : 0);
  const long long ctasPerChunk = ((long long)nstrips * nk + WARPS - 1) / WARPS;
  int chunk = 64;
  while (chunk > 8 && ctasPerChunk * ((nj + chunk - 1) / chunk) < 4 * pencil::smCount() * IDFX_ROW_MINB) chunk >>= 1;
  if (chunk > nj) chunk = nj;
  const dim3 grid((unsigned)ctasPerChunk, (unsigned)((nj + chunk - 1) / chunk), 1);
  static unsigned long long configured = 0;   // bit d: attribute set on device d
  auto kern = pencilRowKernel<P, SOLVER, ORDER, WARPS, CART>;
  if (cudaError_t e = pencil::ensureDynamicSmem(kern, pencilRowSmem<P, ORDER, CART>(WARPS, 64), configured); e != cudaSuccess)
    return e;
  SweepArgs b = a;
  const CUtensorMap *map = reinterpret_cast<const CUtensorMap *>(a.tmapRow);
  b.useTma = (map != nullptr && P::nvar == a.tmapNvar && a.tmapRowCells == 32 + 2 * pencil::Halo<ORDER>::H) ? 1 : 0;
  static const CUtensorMap none = {};
  kern<<<grid, WARPS * 32, pencilRowSmem<P, ORDER, CART>(WARPS, chunk), stream>>>(b, chunk, nstrips, b.useTma ? *map : none);
  return cudaGetLastError();
}

#endif  // IDFX_SWEEP_PENCIL_CUH_
