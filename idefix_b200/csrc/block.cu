// block.cu -- the C ABI (include/idefix_b200.h): block life cycle, geometry set-up,
// the granular entry points mirroring the reference's call sites, and the fused
// TimeIntegrator::Cycle path.
//
// Host-side geometry follows src/dataBlock/makeGeometry.cpp:16-217 and
// src/fluid/RiemannSolver/extrapolateToFaces.hpp:43-71 expression by expression (same
// libm, same association) so that the separable 1-D factors reproduce the reference's
// 3-D dV / A arrays bit for bit.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/idefix_b200.h"
#include "kernels.h"
#include "tma.cuh"

namespace {

thread_local std::string g_error;

int fail(const std::string &msg) {
  g_error = msg;
  return 1;
}

#define CUDA_TRY(call)                                                                        \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(std::string(#call) + ": " + cudaGetErrorString(e_));                        \
  } while (0)

#define KTRY(call)                                                                            \
  do {                                                                                        \
    int e_ = (call);                                                                          \
    if (e_ == -1) return fail(std::string(#call) + ": configuration not instantiated");       \
    if (e_ != 0) return fail(std::string(#call) + ": " + cudaGetErrorString((cudaError_t)e_)); \
  } while (0)

}  // namespace

struct IdfxBlock {
  IdfxBlockDesc desc;
  PhysCfg cfg;
  DevGrid g;
  DevEmf emf;
  double *Vc = nullptr, *Uc = nullptr, *Vs = nullptr, *Uc0 = nullptr, *Vs0 = nullptr, *invDt = nullptr;
  long long nCells = 0, nVs = 0;
  double *state = nullptr;                  // device: t, dt, t0, tstop
  unsigned long long *redBits = nullptr;    // device: [0] dt-min bits, [1] divB-max bits
  long long *ncyclesDev = nullptr;
  int *statusDev = nullptr, *nanDev = nullptr;
  double *hostState = nullptr;              // pinned mirror (4 doubles + scratch)
  cudaStream_t stream = nullptr;
  std::vector<void *> allocs;
  long long launches = 0;
  int nproc[3] = {1, 1, 1};
  bool haveUserdef = false;
  bool fused = true;
  IdfxUserBoundaryFn userBc = nullptr;
  void *userBcData = nullptr;
  IdfxHaloFn haloFn = nullptr;
  void *haloData = nullptr;
  double stageTime = 0.0;                              // t handed to idfx_evolve_stage (granular path)
  bool stageTimeKnown = false;
  IdfxStageHookFn stageHook[2] = {nullptr, nullptr};   // internal boundary, EMF boundary
  void *stageHookData[2] = {nullptr, nullptr};
  IdfxReduceFn reduceFn = nullptr;
  void *reduceData = nullptr;
  double *haloSend[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
  double *haloRecv[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
  double *haloPeer[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
  // peer-to-peer exchange through device mailboxes (idfx_p2p_connect): no host callback, graph-capturable
  bool p2p = false;
  int p2pRank = 0, p2pWorld = 1;
  unsigned long long *mailbox = nullptr, *p2pSeq = nullptr;
  unsigned int *p2pTicket = nullptr;
  unsigned long long *peerMailbox[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
  unsigned long long *allMailbox[idfxk::P2P_MAX_RANKS] = {};
  std::vector<std::vector<double>> hostCoords;  // owned copies of the descriptor's coordinate arrays
  double wc[2] = {0, 0}, w0[2] = {0, 0};
  bool ownStream = true;
  unsigned int *mathFlag = nullptr;   // device counter of failed fast-path range checks (fastmath.cuh)
  double *fluxRho = nullptr;   // mass flux of the current sweep, for the tracer kernels (only with tracers)
  double *staleFlux = nullptr; // 1-D HD Roe: area-weighted flux of the previous call (roeHD.hpp:312-328)
  unsigned char *shockFlags = nullptr;   // [Hydro] shockFlattening: FlagShock per cell (shockFlattening.hpp)
  double *Ve = nullptr, *Ve0 = nullptr;  // EVOLVE_VECTOR_POTENTIAL: edge-centred potential (laid out like Vs) + "begin" copy
  int nVe = 0;                           // components of Ve: 3 (AX1e, AX2e, AX3e) in 3-D, 1 (AX3e) in 2-D
  // TMA descriptor of Vc for the x1 sweep (tma.cuh): boxes [nvar][32 + 2H cells]
  CUtensorMap tmRow;
  bool haveTmRow = false;
  int tmRowCells = 0;
  // steady-state cycle of the fused path captured as a CUDA graph (launch-bound small grids: ~40 launches per cycle)
  cudaGraphExec_t cycleGraph = nullptr;
  long long cycleGraphLaunches = 0;
  bool useGraph = true;
  // polar axis (axis.hpp:52-56): sides of x2 that lie on it, and whether x3 covers the full circle
  bool haveAxis = false, axisLeft = false, axisRight = false, axisTwoPi = false;
  bool vcFieldFresh = false;   // active cells of Vc(BXn) equal the face average of Vs (set by the fused stage epilogue)
  // overlap of the ghost-zone work with the update of the interior (fused path): the boundary conditions / halo
  // exchange of the NEXT stage run on a side stream as soon as the cells they read (the shell: active cells within
  // nghost of a face) have their new primitives, while the main stream converts the rest of the block
  cudaStream_t sideStream = nullptr;
  cudaEvent_t evShell = nullptr, evGhosts = nullptr;
  bool overlap = true, forceOverlap = false;
  bool ghostsDone = false;     // the ghost zones already hold the boundary conditions of the current Vc / Vs
  // per-kernel-class timing with CUDA events on the launching stream (idfx_timing)
  bool timing = false;
  std::vector<cudaEvent_t> evPool;      // pairs: [2n] start, [2n+1] stop
  std::vector<int> evCat;               // category of pair n
  size_t evUsed = 0;
};

static void dropCycleGraph(IdfxBlock *b) {
  if (b->cycleGraph) cudaGraphExecDestroy(b->cycleGraph);
  b->cycleGraph = nullptr;
}

enum { TIME_SWEEP_X1 = 0, TIME_SWEEP_X2 = 1, TIME_SWEEP_X3 = 2, TIME_CT = 3, TIME_EPILOGUE = 4, TIME_BOUNDARY = 5,
       TIME_NCAT = 6 };

// RAII bracket: records an event pair around the launches issued while it is alive
struct TimeScope {
  IdfxBlock *b;
  size_t slot = (size_t)-1;
  TimeScope(IdfxBlock *blk, int cat) : b(blk) {
    if (!b->timing) return;
    if (b->evUsed * 2 + 2 > b->evPool.size()) {
      if (b->evPool.size() >= 2 * 16384) return;  // bounded pool: stop sampling, never grow without limit
      cudaEvent_t e0, e1;
      if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return;
      b->evPool.push_back(e0);
      b->evPool.push_back(e1);
      b->evCat.push_back(cat);
    }
    slot = b->evUsed++;
    b->evCat[slot] = cat;
    cudaEventRecord(b->evPool[2 * slot], b->stream);
  }
  ~TimeScope() {
    if (slot != (size_t)-1) cudaEventRecord(b->evPool[2 * slot + 1], b->stream);
  }
};

namespace {

template <typename T>
int devAlloc(IdfxBlock *b, T **p, long long n) {
  if (n <= 0) n = 1;
  void *q = nullptr;
  cudaError_t e = cudaMalloc(&q, (size_t)n * sizeof(T));
  if (e != cudaSuccess) return fail(std::string("cudaMalloc: ") + cudaGetErrorString(e));
  e = cudaMemsetAsync(q, 0, (size_t)n * sizeof(T), b->stream);
  if (e != cudaSuccess) return fail(std::string("cudaMemset: ") + cudaGetErrorString(e));
  b->allocs.push_back(q);
  *p = (T *)q;
  return 0;
}

int upload1D(IdfxBlock *b, const double **dst, const std::vector<double> &v) {
  double *p = nullptr;
  if (devAlloc(b, &p, (long long)v.size())) return 1;
  CUDA_TRY(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice, b->stream));
  CUDA_TRY(cudaStreamSynchronize(b->stream));
  *dst = p;
  return 0;
}

// DataBlock::MakeGeometry + ExtrapolateToFaces::ComputePLMweights on the host
int buildGeometry(IdfxBlock *b) {
  const IdfxBlockDesc &d = b->desc;
  DevGrid &g = b->g;
  const int geo = d.geometry;
  const int nd = d.ndim;
  int nx[3];
  for (int dir = 0; dir < 3; dir++) nx[dir] = g.nx[dir];
  const double *x1 = d.x[0], *x2 = d.x[1], *x3 = d.x[2];
  const double *x1m = d.xl[0], *x2m = d.xl[1];
  const double *x1p = d.xr[0], *x2p = d.xr[1];
  const double *dx1 = d.dx[0], *dx2 = d.dx[1], *dx3 = d.dx[2];
  (void)x3;

  std::vector<double> xgc[3], vol[3], area[3][3];
  std::vector<double> rt(nx[0], 0.0), sinx2m(nx[1] + 1, 0.0), tanx2m(nx[1], 0.0), sinx2(nx[1], 0.0),
      tanx2(nx[1], 0.0), dmu(nx[1], 0.0);
  for (int dir = 0; dir < 3; dir++) {
    xgc[dir].assign(nx[dir], 0.0);
    vol[dir].assign(nx[dir], 1.0);
    for (int e = 0; e < 3; e++) area[dir][e].assign(nx[e] + 1, 1.0);
  }
  // volumes (makeGeometry.cpp:66-84)
  for (int i = 0; i < nx[0]; i++) {
    if (geo == IDFX_CARTESIAN) vol[0][i] = dx1[i];
    else if (geo == IDFX_CYLINDRICAL || geo == IDFX_POLAR) vol[0][i] = std::fabs(x1p[i] * x1p[i] - x1m[i] * x1m[i]) / 2.0;
    else vol[0][i] = std::fabs(x1p[i] * x1p[i] * x1p[i] - x1m[i] * x1m[i] * x1m[i]) / 3.0;
  }
  if (nd >= 2) {
    for (int j = 0; j < nx[1]; j++) {
      if (geo == IDFX_SPHERICAL) vol[1][j] = std::fabs(std::cos(x2m[j]) - std::cos(x2p[j]));
      else vol[1][j] = dx2[j];
    }
  }
  if (nd == 3) {
    for (int k = 0; k < nx[2]; k++) vol[2][k] = (geo == IDFX_CYLINDRICAL) ? 1.0 : dx3[k];
  }
  // geometrical centres and spherical helpers (makeGeometry.cpp:91-130)
  for (int i = 0; i < nx[0]; i++) {
    if (geo == IDFX_CARTESIAN) {
      xgc[0][i] = x1[i];
    } else if (geo == IDFX_CYLINDRICAL || geo == IDFX_POLAR) {
      xgc[0][i] = x1[i] + dx1[i] * dx1[i] / (12.0 * x1[i]);
    } else {
      xgc[0][i] = x1[i] + 2.0 * x1[i] * dx1[i] * dx1[i] / (12.0 * x1[i] * x1[i] + dx1[i] * dx1[i]);
      rt[i] = (x1p[i] * x1p[i] * x1p[i] - x1m[i] * x1m[i] * x1m[i]) / (x1p[i] * x1p[i] - x1m[i] * x1m[i]) / 1.5;
    }
  }
  for (int j = 0; j < nx[1]; j++) {
    if (geo != IDFX_SPHERICAL) {
      xgc[1][j] = x2[j];
    } else {
      const double xL = x2m[j], xR = x2p[j];
      xgc[1][j] = (std::sin(xR) - std::sin(xL) + xL * std::cos(xL) - xR * std::cos(xR)) / (std::cos(xL) - std::cos(xR));
      sinx2m[j] = std::sin(xL);
      tanx2m[j] = std::tan(xL);
      sinx2[j] = std::sin(x2[j]);
      tanx2[j] = std::tan(x2[j]);
      dmu[j] = std::fabs(std::cos(xL) - std::cos(xR));
    }
  }
  if (geo == IDFX_SPHERICAL) sinx2m[nx[1]] = std::sin(x2p[nx[1] - 1]);  // used by evolveMagField at j+1 only below np_tot
  for (int k = 0; k < nx[2]; k++) xgc[2][k] = d.x[2][k];

  // areas (makeGeometry.cpp:139-211); factor e of A[dir] multiplies along dimension e, D_EXPAND order
  // A1
  for (int i = 0; i <= nx[0]; i++) {
    double f;
    const bool last = (i == nx[0]);
    if (geo == IDFX_CARTESIAN) f = 1.0;
    else if (geo == IDFX_CYLINDRICAL || geo == IDFX_POLAR) f = last ? std::fabs(x1p[i - 1]) : std::fabs(x1m[i]);
    else f = last ? x1p[i - 1] * x1p[i - 1] : x1m[i] * x1m[i];
    area[0][0][i] = f;
  }
  if (nd >= 2)
    for (int j = 0; j < nx[1]; j++) area[0][1][j] = (geo == IDFX_SPHERICAL) ? std::fabs(std::cos(x2m[j]) - std::cos(x2p[j])) : dx2[j];
  if (nd == 3)
    for (int k = 0; k < nx[2]; k++) area[0][2][k] = (geo == IDFX_CYLINDRICAL) ? 1.0 : dx3[k];
  // A2
  for (int i = 0; i < nx[0]; i++) {
    double f;
    if (geo == IDFX_CARTESIAN || geo == IDFX_POLAR) f = dx1[i];
    else if (geo == IDFX_CYLINDRICAL) f = (nd >= 2) ? std::fabs(x1[i]) * dx1[i] : std::fabs(x1[i]);
    else f = x1[i] * dx1[i];
    area[1][0][i] = f;
  }
  if (nd >= 2) {
    for (int j = 0; j <= nx[1]; j++) {
      double f = 1.0;
      if (geo == IDFX_SPHERICAL) f = (j == nx[1]) ? std::fabs(std::sin(x2p[j - 1])) : std::fabs(std::sin(x2m[j]));
      area[1][1][j] = f;
    }
  }
  if (nd == 3)
    for (int k = 0; k < nx[2]; k++) area[1][2][k] = (geo == IDFX_CYLINDRICAL) ? 1.0 : dx3[k];
  // A3
  for (int i = 0; i < nx[0]; i++) {
    double f;
    if (geo == IDFX_CARTESIAN) f = dx1[i];
    else if (geo == IDFX_CYLINDRICAL) f = 1.0;
    else f = x1[i] * dx1[i];
    area[2][0][i] = f;
  }
  if (nd >= 2)
    for (int j = 0; j < nx[1]; j++) area[2][1][j] = (geo == IDFX_CYLINDRICAL) ? 1.0 : dx2[j];
  // area[2][2] stays 1

  for (int dir = 0; dir < 3; dir++) {
    std::vector<double> vx(d.x[dir], d.x[dir] + nx[dir]), vl(d.xl[dir], d.xl[dir] + nx[dir] ),
        vr(d.xr[dir], d.xr[dir] + nx[dir]), vd(d.dx[dir], d.dx[dir] + nx[dir]);
    // one extra entry so that x1m(i+1) style accesses at the last index stay in bounds
    vx.push_back(vx.back());
    vl.push_back(vr.back());
    vr.push_back(vr.back());
    vd.push_back(vd.back());
    if (upload1D(b, &g.x[dir], vx) || upload1D(b, &g.xl[dir], vl) || upload1D(b, &g.xr[dir], vr) ||
        upload1D(b, &g.dx[dir], vd) || upload1D(b, &g.xgc[dir], xgc[dir]) || upload1D(b, &g.vol[dir], vol[dir]))
      return 1;
    for (int e = 0; e < 3; e++)
      if (upload1D(b, &g.area[dir][e], area[dir][e])) return 1;
  }
  if (upload1D(b, &g.rt, rt) || upload1D(b, &g.sinx2m, sinx2m) || upload1D(b, &g.tanx2m, tanx2m) ||
      upload1D(b, &g.sinx2, sinx2) || upload1D(b, &g.tanx2, tanx2) || upload1D(b, &g.dmu, dmu))
    return 1;

  // irregular-grid PLM weights (extrapolateToFaces.hpp:43-71)
  for (int dir = 0; dir < 3; dir++) {
    const int n = nx[dir];
    std::vector<double> wp(n, 0.0), wm(n, 0.0), cp(n, 0.0), cm(n, 0.0), dp(n, 0.0), dm(n, 0.0);
    if (!d.regular_grid && dir < nd) {
      const double *dx = d.dx[dir], *xr = d.xr[dir];
      const std::vector<double> &gc = xgc[dir];
      for (int i = 1; i < n - 1; i++) {
        wp[i] = dx[i] / (gc[i + 1] - gc[i]);
        wm[i] = dx[i] / (gc[i] - gc[i - 1]);
        cp[i] = (gc[i + 1] - gc[i]) / (xr[i] - gc[i]);
        cm[i] = (gc[i] - gc[i - 1]) / (gc[i] - xr[i - 1]);
        dp[i] = (xr[i] - gc[i]) / dx[i];
        dm[i] = (gc[i] - xr[i - 1]) / dx[i];
      }
    }
    if (upload1D(b, &g.wp[dir], wp) || upload1D(b, &g.wm[dir], wm) || upload1D(b, &g.cp[dir], cp) ||
        upload1D(b, &g.cm[dir], cm) || upload1D(b, &g.dp[dir], dp) || upload1D(b, &g.dm[dir], dm))
      return 1;
  }
  return 0;
}

int setupPhys(IdfxBlock *b) {
  const IdfxBlockDesc &d = b->desc;
  PhysCfg &c = b->cfg;
  c.ndim = d.ndim;
  c.ncomp = d.ncomp;
  c.mhd = d.mhd ? 1 : 0;
  c.iso = d.isothermal ? 1 : 0;
  c.order = d.order;
  c.solver = d.solver;
  c.emf = d.emf;
  const int energy = c.iso ? 0 : 1;
  c.RHO = 0;
  c.MX1 = 1;
  c.MX2 = (c.ncomp >= 2 ? 2 : 255);
  c.MX3 = (c.ncomp == 3 ? 3 : 254);
  c.ENG = energy ? c.ncomp + 1 : 250;
  c.BX1 = c.mhd ? c.ncomp + 1 + energy : 253;
  c.BX2 = (c.mhd && c.ncomp >= 2) ? c.BX1 + 1 : 252;
  c.BX3 = (c.mhd && c.ncomp == 3) ? c.BX1 + 2 : 251;
  c.nvar = 1 + c.ncomp + energy + (c.mhd ? c.ncomp : 0);
  c.ntr = d.ntracer;
  return 0;
}

int validate(const IdfxBlockDesc &d) {
  if (d.ndim < 1 || d.ndim > 3) return fail("ndim must be 1..3");
  if (d.ncomp < d.ndim || d.ncomp > 3) return fail("ncomp must satisfy ndim <= ncomp <= 3");
  if (d.order < 1 || d.order > 4) return fail("order must be 1..4");
  if (d.shock_flattening < 0.0) return fail("shockFlattening needs a non-negative smoothing value");
  if (d.evolve_vector_potential && !(d.mhd && d.ndim >= 2))
    return fail("EVOLVE_VECTOR_POTENTIAL is not compatible with 1D MHD");   // fluid.hpp:549-551
  if (d.nstages < 1 || d.nstages > 3) return fail("nstages must be 1..3");
  if (d.ntracer < 0 || d.ntracer > 8) return fail("The number of passive tracers should be between 0 and 8");
  if (d.geometry < IDFX_CARTESIAN || d.geometry > IDFX_SPHERICAL) return fail("unknown geometry");
  if (d.geometry == IDFX_CYLINDRICAL && d.ndim == 3) return fail("CYLINDRICAL is 2-D only; use POLAR");
  if (d.mhd) {
    if (d.solver != IDFX_TVDLF && d.solver != IDFX_HLL && d.solver != IDFX_HLLD && d.solver != IDFX_ROE)
      return fail("hllc Riemann solver requires a HD fluid");
    if (d.emf < IDFX_EMF_ARITHMETIC || d.emf > IDFX_EMF_UCT_HLLD) return fail("Unknown EMF averaging scheme");
    if (d.emf == IDFX_EMF_UCT_HLLD && (d.solver == IDFX_HLL || d.solver == IDFX_TVDLF))
      return fail("HLLD EMF reconstruction is only compatible with HLLD or ROE Riemann solvers");
  } else {
    if (d.solver == IDFX_HLLD) return fail("hlld Riemann solver requires a MHD fluid");
  }
  for (int dir = 0; dir < d.ndim; dir++) {
    const int need = (d.order == 4) ? 3 : 2;
    if (d.nghost[dir] < need) return fail("nghost too small for the reconstruction order");
    if (d.np_int[dir] < 1) return fail("np_int must be >= 1");
    const int lb = d.lbound[dir], rb = d.rbound[dir];
    for (int t : {lb, rb}) {
      if (t == IDFX_BC_SHEARINGBOX) return fail("shearingbox boundaries are outside the scope of this build");
      if (t == IDFX_BC_AXIS) {
        if (dir != 1) return fail("Axis Boundaries are only applicable to X2");   // grid.cpp:98-101
        if (d.geometry != IDFX_SPHERICAL)
          return fail("Axis boundary conditions are only designed to handle spherical geometry");   // axis.hpp:112-114
        if (d.ndim == 3 && (d.lbound[2] == IDFX_BC_INTERNAL || d.rbound[2] == IDFX_BC_INTERNAL))
          return fail("axis boundaries with a domain decomposition along X3 are outside the scope of this build");
      }
      if (t < 0 || t >= IDFX_BC_UNDEFINED) return fail("undefined boundary condition");
    }
  }
  for (int dir = 0; dir < 3; dir++)
    if (!d.x[dir] || !d.xl[dir] || !d.xr[dir] || !d.dx[dir]) return fail("coordinate arrays must be provided");
  return 0;
}

SweepArgs baseSweepArgs(IdfxBlock *b) {
  SweepArgs a;
  std::memset(&a, 0, sizeof(a));
  a.g = b->g;
  a.emf = b->emf;
  a.Vc = b->Vc;
  a.Uc = b->Uc;
  a.Vs = b->Vs;
  a.invDt = b->invDt;
  a.dt = b->state + 1;
  a.gamma = b->desc.gamma;
  a.csiso = b->desc.csiso;
  a.emfMode = b->cfg.emf;
  a.perpExt = 0;
  if (b->cfg.mhd) {
    a.perpExt = 1;
    // HLL / HLLD widen the face box for the UCT-HLLx stencils; Roe and TVDLF never do
    // (hlldMHD.hpp:32-37, hllMHD.hpp:30-35 vs roeMHD.hpp:63-73, tvdlfMHD.hpp:31)
    if ((b->cfg.emf == EMF_UCT_HLL || b->cfg.emf == EMF_UCT_HLLD) &&
        (b->cfg.solver == SOLVER_HLL || b->cfg.solver == SOLVER_HLLD))
      a.perpExt = -1;  // resolved per direction: nghost[dir]
  }
  a.doInvDt = 1;
  a.fluxRho = b->fluxRho;
  a.staleFlux = b->staleFlux;
  a.shockFlags = b->shockFlags;
  a.tmapRow = b->haveTmRow ? &b->tmRow : nullptr;
  a.tmapNvar = b->cfg.nvar;
  a.tmapRowCells = b->tmRowCells;
  a.mathFlag = b->mathFlag;
  a.validateInputs = b->vcFieldFresh ? 0 : 1;   // our own stage epilogue already validated what it wrote
  return a;
}

int launchSweepDir(IdfxBlock *b, int dir, SweepArgs a) {
  if (a.perpExt == -1) a.perpExt = b->desc.nghost[dir];
  TimeScope ts(b, TIME_SWEEP_X1 + dir);
  int rc;
  const PhysCfg &c = b->cfg;
  if (!c.mhd) {
    rc = (c.solver == SOLVER_ROE) ? idfxk::sweepHDRoe(c, dir, a, b->stream) : idfxk::sweepHD(c, dir, a, b->stream);
  } else {
    switch (c.solver) {
      case SOLVER_TVDLF: rc = idfxk::sweepMhdTvdlf(c, dir, a, b->stream); break;
      case SOLVER_HLL: rc = idfxk::sweepMhdHll(c, dir, a, b->stream); break;
      case SOLVER_HLLD: rc = idfxk::sweepMhdHlld(c, dir, a, b->stream); break;
      case SOLVER_ROE: rc = idfxk::sweepMhdRoe(c, dir, a, b->stream); break;
      default: rc = -1;
    }
  }
  b->launches++;
  return rc;
}

// host copy of (t, dt); synchronises the stream
int readState(IdfxBlock *b) {
  CUDA_TRY(cudaMemcpyAsync(b->hostState, b->state, 4 * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
  CUDA_TRY(cudaStreamSynchronize(b->stream));
  return 0;
}

bool needsExchange(const IdfxBlock *b, int dir) { return b->nproc[dir] > 1; }

// a user hook of the stage (IdfxStageHook): the stream is drained, the callback runs with the current time, and the
// device is drained again (the callback may have launched kernels on other streams)
int callStageHook(IdfxBlock *b, int which, double t, bool tKnown) {
  if (!b->stageHook[which]) return 0;
  if (tKnown) {
    CUDA_TRY(cudaStreamSynchronize(b->stream));
  } else {
    if (readState(b)) return 1;
    t = b->hostState[0];
  }
  b->stageHook[which](b, t, b->stageHookData[which]);
  CUDA_TRY(cudaDeviceSynchronize());
  return 0;
}

// a side exchanges ghosts with another block when it is internal, or periodic with nproc[dir] > 1
// (mpi.cpp:262-266: MPI_PROC_NULL neighbours on physical boundaries)
bool sideFacesBlock(const IdfxBlock *b, int dir, int side) {
  const int type = side == 0 ? b->desc.lbound[dir] : b->desc.rbound[dir];
  return type == IDFX_BC_INTERNAL || (type == IDFX_BC_PERIODIC && b->nproc[dir] > 1);
}

// Mpi::ExchangeX1/X2/X3 as two kernels over peer memory (boundary.cu: haloP2PKernel)
int p2pExchange(IdfxBlock *b, int dir) {
  idfxk::P2PChannel ch;
  ch.dir = dir;
  for (int s = 0; s < 2; s++) {
    ch.active[s] = sideFacesBlock(b, dir, s) ? 1 : 0;
    ch.peerRecv[s] = b->haloPeer[dir][s];
    ch.localRecv[s] = b->haloRecv[dir][s];
    ch.peerMailbox[s] = b->peerMailbox[dir][s];
    if (ch.active[s] && (!ch.peerRecv[s] || !ch.peerMailbox[s] || !ch.localRecv[s]))
      return fail("peer-to-peer halo exchange: a neighbour of this block has not been connected (idfx_p2p_connect)");
  }
  ch.mailbox = b->mailbox;
  ch.seq = b->p2pSeq;
  ch.ticket = b->p2pTicket;
  KTRY(idfxk::haloP2P(b->cfg, b->g, b->Vc, b->Vs, true, ch, b->stream, &b->launches));
  KTRY(idfxk::haloP2P(b->cfg, b->g, b->Vc, b->Vs, false, ch, b->stream, &b->launches));
  return 0;
}

// Boundary::EnforceBoundaryDir (boundary.hpp:287-385) + ReconstructNormalField (:408-530) of one direction
int enforceBoundaryDirImpl(IdfxBlock *b, int dir, double t, bool tKnown) {
  const IdfxBlockDesc &d = b->desc;
  {
    if (needsExchange(b, dir)) {
      if (b->p2p) {
        if (p2pExchange(b, dir)) return 1;
      } else {
        if (!b->haloFn) return fail("nproc > 1 in a direction but no halo exchange has been enrolled");
        b->haloFn(b, dir, b->haloData);
      }
    }
    for (int side = 0; side < 2; side++) {
      const int type = side == 0 ? d.lbound[dir] : d.rbound[dir];
      switch (type) {
        case IDFX_BC_INTERNAL: break;
        case IDFX_BC_PERIODIC:
          if (needsExchange(b, dir)) break;  // periodicity enforced by the exchange
          KTRY(idfxk::enforcePhysicalBoundary(b->cfg, b->g, b->Vc, b->Vs, dir, side, BC_PERIODIC, b->stream, &b->launches));
          break;
        case IDFX_BC_REFLECTIVE:
          KTRY(idfxk::enforcePhysicalBoundary(b->cfg, b->g, b->Vc, b->Vs, dir, side, BC_REFLECTIVE, b->stream, &b->launches));
          break;
        case IDFX_BC_OUTFLOW:
          KTRY(idfxk::enforcePhysicalBoundary(b->cfg, b->g, b->Vc, b->Vs, dir, side, BC_OUTFLOW, b->stream, &b->launches));
          break;
        case IDFX_BC_AXIS:   // boundary.hpp:313-314,359-360
          KTRY(idfxk::enforceAxisBoundary(b->cfg, b->g, b->Vc, b->Vs, side, b->axisTwoPi ? 1 : 0, b->stream, &b->launches));
          break;
        case IDFX_BC_USERDEF: {
          if (!b->userBc)
            return fail("No function has been enrolled to define your own boundary conditions");
          double tt = t;
          if (!tKnown) {
            if (readState(b)) return 1;
            tt = b->hostState[0];
          } else {
            CUDA_TRY(cudaStreamSynchronize(b->stream));
          }
          b->userBc(b, dir, side, tt, b->userBcData);
          CUDA_TRY(cudaDeviceSynchronize());
          break;
        }
        default: return fail("Boundary condition type is not yet implemented");
      }
    }
    if (b->cfg.mhd) {
      const int left = !(d.lbound[dir] == IDFX_BC_PERIODIC || d.lbound[dir] == IDFX_BC_INTERNAL);
      const int right = !(d.rbound[dir] == IDFX_BC_PERIODIC || d.rbound[dir] == IDFX_BC_INTERNAL);
      // across an axis the phi direction is reversed (boundary.hpp:471-474, 3-D only)
      const int signLeft = (dir == 1 && d.ndim == 3 && b->axisLeft) ? -1 : 1;
      const int signRight = (dir == 1 && d.ndim == 3 && b->axisRight) ? -1 : 1;
      KTRY(idfxk::reconstructNormalField(b->cfg, b->g, b->Vs, dir, left, right, b->stream, &b->launches, signLeft, signRight));
      if (dir == 1 && b->haveAxis)   // boundary.hpp:496-499
        KTRY(idfxk::axisRegularizeBx2s(b->g, b->Vs, b->axisTwoPi ? 1 : 0, b->axisLeft ? 1 : 0, b->axisRight ? 1 : 0, b->stream,
                                       &b->launches));
    }
  }
  return 0;
}

// Boundary::SetBoundaries without the final ReconstructVcField (boundary.hpp:252-275)
int setBoundariesCore(IdfxBlock *b, double t, bool tKnown) {
  TimeScope ts(b, TIME_BOUNDARY);
  if (b->stageHook[IDFX_HOOK_INTERNAL_BOUNDARY]) {   // boundary.hpp:243-251
    b->vcFieldFresh = false;                          // the hook may have changed Vs anywhere
    if (callStageHook(b, IDFX_HOOK_INTERNAL_BOUNDARY, t, tKnown)) return 1;
  }
  for (int dir = 0; dir < b->desc.ndim; dir++)
    if (enforceBoundaryDirImpl(b, dir, t, tKnown)) return 1;
  return 0;
}

int reconstructVc(IdfxBlock *b, double *V, int ghostsOnly = 0) {
  if (!b->cfg.mhd) return 0;
  KTRY(idfxk::reconstructVcField(b->cfg, b->g, b->Vs, V, b->stream, ghostsOnly));
  b->launches++;
  return 0;
}

// Fluid::EvolveStage (evolveStage.hpp:41-90); fusedStage < 0 : reference-like behaviour
int evolveStageImpl(IdfxBlock *b, int fusedStage) {
  const IdfxBlockDesc &d = b->desc;
  const bool fused = fusedStage >= 0;
  const bool multi = d.nstages > 1;
  if (fused && b->cfg.ntr > 0) {
    // Tracer::ConvertPrimToCons (+ the "begin" copy on the first stage); the fluid part is done by the x1 sweep
    KTRY(idfxk::tracerPrimToCons(b->cfg, b->g, b->Vc, b->Uc, (fusedStage == 0 && multi) ? b->Uc0 : nullptr, b->stream));
    b->launches++;
  }
  if (b->shockFlags) {   // RiemannSolver::CalcFlux<IDIR> (calcFlux.hpp:31-34): flag the shocks once per stage
    KTRY(idfxk::findShock(b->cfg, b->g, b->Vc, b->shockFlags, d.shock_flattening, d.csiso, b->stream));
    b->launches++;
  }
  for (int dir = 0; dir < d.ndim; dir++) {
    SweepArgs a = baseSweepArgs(b);
    a.doSource = (dir == d.ndim - 1 && d.geometry != IDFX_CARTESIAN) ? 1 : 0;
    if (fused) {
      a.doInvDt = (fusedStage == 0) ? 1 : 0;
      a.invDtInit = (dir == 0) ? 1 : 0;
      a.fusedP2C = (dir == 0) ? 1 : 0;
      a.Uc0 = (dir == 0 && fusedStage == 0 && multi) ? b->Uc0 : nullptr;
    }
    KTRY(launchSweepDir(b, dir, a));
    if (b->cfg.ntr > 0) {   // Tracer::CalcFlux / CalcRightHandSide (evolveStage.hpp:22-31)
      const long long off = (long long)b->cfg.nvar * b->nCells;
      KTRY(idfxk::tracerSweep(b->g, dir, b->Vc + off, b->Uc + off, b->fluxRho, b->state + 1, b->cfg.ntr, b->stream));
      b->launches++;
    }
  }
  if (b->cfg.mhd && d.ndim >= 2) {
    TimeScope ts(b, TIME_CT);
    KTRY(idfxk::calcCornerEmf(b->g, b->emf, b->Vc, b->Vs, b->cfg.emf, b->cfg.order, b->cfg.ncomp, b->cfg.MX1,
#ifdef IDFX_ROW_NOFOLD
                              b->cfg.BX1, b->stream, &b->launches, 0));
#else
                              b->cfg.BX1, b->stream, &b->launches, 1));
#endif
    // EnforceEMFBoundary (enforceEMFBoundary.hpp:29-30); granular callers passed the stage time to idfx_evolve_stage
    if (callStageHook(b, IDFX_HOOK_EMF_BOUNDARY, b->stageTime, b->stageTimeKnown)) return 1;
    if (b->haveAxis)   // enforceEMFBoundary.hpp:32-34
      KTRY(idfxk::axisRegularizeEmf(b->g, b->emf.ex, b->emf.ez, b->axisTwoPi ? 1 : 0, b->axisLeft ? 1 : 0,
                                    b->axisRight ? 1 : 0, b->stream, &b->launches));
    int mode = 0;
    double wc = 0, w0 = 0;
    if (fused && multi) {
      mode = (fusedStage == 0) ? 1 : 2;
      if (fusedStage > 0) {
        wc = b->wc[fusedStage - 1];
        w0 = b->w0[fusedStage - 1];
      }
    }
    if (b->Ve) {
      // EVOLVE_VECTOR_POTENTIAL (evolveStage.hpp:78-81): Ve += -dt E, then Vs = curl Ve.  On the fused path the
      // Runge-Kutta combination of Ve (StateContainer::AddAndStore) is folded into the first kernel, so the Vs derived
      // here is already the one ConvertConsToPrim would derive after the combination
      KTRY(idfxk::evolveVectorPotential(b->g, b->emf, b->Ve, b->Ve0, b->state + 1, mode, wc, w0, b->stream, &b->launches));
      KTRY(idfxk::magFieldFromA(b->g, b->Ve, b->Vs, b->haveAxis ? 1 : 0, b->stream, &b->launches));
    } else {
      KTRY(idfxk::evolveMagField(b->g, b->emf, b->Vs, b->Vs0, b->state + 1, b->haveAxis ? 1 : 0, mode, wc, w0, b->stream,
                                 &b->launches));
    }
    if (!fused) {
      if (reconstructVc(b, b->Uc)) return 1;
    }
  }
  return 0;
}

int timestepReduceImpl(IdfxBlock *b) {
  KTRY(idfxk::timestepReduce(b->g, b->invDt, b->redBits, b->stream));
  b->launches++;
  return 0;
}

// MPI_Iallreduce(MIN) of timeIntegrator.cpp:318: the enrolled hook reduces the device scalar in place
// across blocks, enqueued on the block's stream (a non-negative double orders like its bit pattern)
void crossBlockReduce(IdfxBlock *b, double *devValue, int op) {
  if (b->nproc[0] * b->nproc[1] * b->nproc[2] <= 1) return;
  if (b->p2p) {
    idfxk::P2PReduce r;
    r.rank = b->p2pRank;
    r.world = b->p2pWorld;
    r.mailbox = b->mailbox;
    r.seq = b->p2pSeq;
    for (int q = 0; q < idfxk::P2P_MAX_RANKS; q++) r.allMailbox[q] = b->allMailbox[q];
    idfxk::p2pAllReduce(r, devValue, op == IDFX_REDUCE_MIN ? 0 : 1, b->stream);
    b->launches++;
  } else if (b->reduceFn) {
    b->reduceFn(b, devValue, op, b->reduceData);
  }
}

// TimeIntegrator::Cycle as a sequence of the reference's own steps (timeIntegrator.cpp:246-414)
int cycleReferenceOrder(IdfxBlock *b) {
  const IdfxBlockDesc &d = b->desc;
  KTRY(idfxk::dtControlBegin(b->state, b->redBits, b->stream));
  KTRY(idfxk::fill(b->invDt, b->nCells, 0.0, b->stream));
  b->launches += 2;
  for (int stage = 0; stage < d.nstages; stage++) {
    if (setBoundariesCore(b, 0.0, false)) return 1;
    if (reconstructVc(b, b->Vc)) return 1;
    KTRY(idfxk::primToCons(b->cfg, b->g, b->Vc, b->Uc, d.gamma, b->stream));
    b->launches++;
    if (b->cfg.ntr > 0) {
      KTRY(idfxk::tracerPrimToCons(b->cfg, b->g, b->Vc, b->Uc, nullptr, b->stream));
      b->launches++;
    }
    if (d.nstages > 1 && stage == 0) {
      KTRY(idfxk::copy(b->Uc0, b->Uc, b->nCells * (b->cfg.nvar + b->cfg.ntr), b->stream));
      b->launches++;
      if (b->cfg.mhd) {   // states["current"] holds Ve instead of Vs when the potential is evolved (fluid.hpp:548-559)
        if (b->Ve) KTRY(idfxk::copy(b->Ve0, b->Ve, b->nVs * b->nVe, b->stream));
        else KTRY(idfxk::copy(b->Vs0, b->Vs, b->nVs * d.ndim, b->stream));
        b->launches++;
      }
    }
    if (evolveStageImpl(b, -1)) return 1;
    const double wc = stage > 0 ? b->wc[stage - 1] : 0.0, w0 = stage > 0 ? b->w0[stage - 1] : 0.0;
    KTRY(idfxk::stageTimeAdvance(b->state, stage, wc, w0, b->stream));
    b->launches++;
    if (stage == 0 && !(d.fixed_dt > 0.0)) {
      if (timestepReduceImpl(b)) return 1;
      crossBlockReduce(b, (double *)b->redBits, IDFX_REDUCE_MIN);
    }
    if (stage > 0) {
      KTRY(idfxk::addAndStore(b->Uc, b->Uc0, b->nCells * (b->cfg.nvar + b->cfg.ntr), wc, w0, b->stream));
      b->launches++;
      if (b->cfg.mhd) {
        if (b->Ve) KTRY(idfxk::addAndStore(b->Ve, b->Ve0, b->nVs * b->nVe, wc, w0, b->stream));
        else KTRY(idfxk::addAndStore(b->Vs, b->Vs0, b->nVs * d.ndim, wc, w0, b->stream));
        b->launches++;
      }
    }
    if (b->Ve) KTRY(idfxk::magFieldFromA(b->g, b->Ve, b->Vs, b->haveAxis ? 1 : 0, b->stream, &b->launches));   // convertConsToPrim.hpp:117-119
    if (reconstructVc(b, b->Uc)) return 1;
    KTRY(idfxk::consToPrim(b->cfg, b->g, b->Vc, b->Uc, d.gamma, b->stream));
    b->launches++;
    if (b->cfg.ntr > 0) {
      KTRY(idfxk::tracerConsToPrim(b->cfg, b->g, b->Vc, b->Uc, nullptr, 0, 0.0, 0.0, 0, b->stream));
      b->launches++;
    }
  }
  KTRY(idfxk::dtControlEnd(b->state, b->redBits, d.cfl, d.cfl_max_var, d.fixed_dt, b->ncyclesDev, b->statusDev,
                           b->mathFlag, b->stream));
  b->launches++;
  return 0;
}

// ghost zones of the current state: Boundary::SetBoundaries + the ghost shell of ReconstructVcField
int fillGhosts(IdfxBlock *b) {
  if (setBoundariesCore(b, 0.0, false)) return 1;
  // the active cells already hold the face average when the previous step on this block was our own
  // stage epilogue (b->vcFieldFresh): only the ghost cells changed with the boundary conditions
  return reconstructVc(b, b->Vc, b->vcFieldFresh ? 1 : 0);
}

// The stage epilogue, and -- overlapped with it -- the ghost zones of the next stage.  The boundary kernels and the
// halo exchange only READ the shell (active cells within nghost of a face, plus the faces of Vs, which the epilogue
// does not write) and only WRITE ghost cells, which the epilogue does not touch: once the shell is converted they
// run on the side stream while the main stream converts the core of the block.
int epilogueAndGhosts(IdfxBlock *b, int stage, double wc, double w0, int keepUc, bool reduceDt) {
  const IdfxBlockDesc &d = b->desc;
  const DevGrid &g = b->g;
  auto epi = [&](const int *lo, const int *hi) -> int {
    KTRY(idfxk::stageEpilogue(b->cfg, g, b->Vc, b->Uc, b->Vs, b->Uc0, d.gamma, stage > 0 ? 1 : 0, wc, w0, keepUc, b->stream,
                              b->mathFlag, lo, hi, reduceDt ? b->invDt : nullptr, reduceDt ? b->redBits : nullptr));
    b->launches++;
    return 0;
  };
  int core_lo[3], core_hi[3];
  // worth it only when there is a neighbour to wait for: on a single block the ghost fill is 0.6 ms of tiny kernels,
  // and running them beside the bandwidth-bound epilogue (plus the six thin shell launches) costs slightly more than
  // it hides (measured, profiles/r2_overlap_ab.md).  IDFX_OVERLAP=1 forces it on.
  const bool multi = b->nproc[0] * b->nproc[1] * b->nproc[2] > 1;
  bool haveCore = b->overlap && (multi || b->forceOverlap) && b->cfg.ntr == 0 && !b->userBc && !b->haloFn && !b->haveAxis &&
                  !b->stageHook[0] && !b->stageHook[1];
  for (int a = 0; a < 3; a++) {
    const int w = (a < d.ndim) ? g.ng[a] : 0;
    core_lo[a] = g.beg[a] + w;
    core_hi[a] = g.end[a] - w;
    if (core_hi[a] - core_lo[a] < 8 * w) haveCore = false;   // small blocks: nothing worth overlapping
  }
  if (!haveCore) {
    TimeScope ts(b, TIME_EPILOGUE);
    return epi(nullptr, nullptr);
  }
  {
    TimeScope ts(b, TIME_EPILOGUE);
    // shell = six disjoint slabs: x3 slabs span the whole active (x1, x2) extent, x2 slabs the core in x3, x1 slabs
    // the core in x2 and x3
    for (int a = d.ndim - 1; a >= 0; a--) {
      for (int side = 0; side < 2; side++) {
        int lo[3], hi[3];
        for (int e = 0; e < 3; e++) {
          lo[e] = (e > a) ? core_lo[e] : g.beg[e];
          hi[e] = (e > a) ? core_hi[e] : g.end[e];
        }
        if (side == 0) hi[a] = core_lo[a];
        else lo[a] = core_hi[a];
        if (epi(lo, hi)) return 1;
      }
    }
  }
  b->vcFieldFresh = true;
  CUDA_TRY(cudaEventRecord(b->evShell, b->stream));
  CUDA_TRY(cudaStreamWaitEvent(b->sideStream, b->evShell, 0));
  {
    cudaStream_t main = b->stream;
    b->stream = b->sideStream;          // the boundary / exchange launchers enqueue on b->stream
    const int rc = fillGhosts(b);
    b->stream = main;
    if (rc) return 1;
  }
  CUDA_TRY(cudaEventRecord(b->evGhosts, b->sideStream));
  {
    TimeScope ts(b, TIME_EPILOGUE);
    if (epi(core_lo, core_hi)) return 1;
  }
  CUDA_TRY(cudaStreamWaitEvent(b->stream, b->evGhosts, 0));
  b->ghostsDone = true;
  return 0;
}

// The same cycle with the point-wise steps folded into the sweeps / CT / one epilogue.
int cycleFused(IdfxBlock *b, bool lastOfBatch) {
  const IdfxBlockDesc &d = b->desc;
  KTRY(idfxk::dtControlBegin(b->state, b->redBits, b->stream));
  b->launches++;
  for (int stage = 0; stage < d.nstages; stage++) {
    if (!b->ghostsDone) {
      if (fillGhosts(b)) return 1;
    }
    b->ghostsDone = false;
    if (evolveStageImpl(b, stage)) return 1;
    const double wc = stage > 0 ? b->wc[stage - 1] : 0.0, w0 = stage > 0 ? b->w0[stage - 1] : 0.0;
    KTRY(idfxk::stageTimeAdvance(b->state, stage, wc, w0, b->stream));
    b->launches++;
    // DataBlock::ComputeTimestep after the first stage (timeIntegrator.cpp:313-323): min(1 / InvDt) is reduced by the
    // epilogue kernel of this stage, which passes over every active cell anyway
    const bool reduceDt = (stage == 0 && !(d.fixed_dt > 0.0));
    // Uc is rebuilt from Vc by the x1 sweep of the next stage, so its write-back (8 of the 24-29 doubles the
    // epilogue moves per cell) is only needed where somebody can look at it: after the last stage of the last
    // cycle of this call, or when the tracer kernels need Uc(RHO)
    const bool last = lastOfBatch && stage == d.nstages - 1;
    const int keepUc = (b->cfg.ntr > 0 || last) ? 1 : 0;
    if (last) {
      // nothing follows inside this call: the caller gets the state back (and may change it before the next one)
      TimeScope ts(b, TIME_EPILOGUE);
      KTRY(idfxk::stageEpilogue(b->cfg, b->g, b->Vc, b->Uc, b->Vs, b->Uc0, d.gamma, stage > 0 ? 1 : 0, wc, w0, keepUc,
                                b->stream, b->mathFlag, nullptr, nullptr, reduceDt ? b->invDt : nullptr,
                                reduceDt ? b->redBits : nullptr));
      b->launches++;
    } else {
      if (epilogueAndGhosts(b, stage, wc, w0, keepUc, reduceDt)) return 1;
    }
    if (reduceDt) crossBlockReduce(b, (double *)b->redBits, IDFX_REDUCE_MIN);
    if (b->cfg.ntr > 0) {   // RK combination + Tracer::ConvertConsToPrim with the final Uc(RHO)
      KTRY(idfxk::tracerConsToPrim(b->cfg, b->g, b->Vc, b->Uc, b->Uc0, stage > 0 ? 1 : 0, wc, w0, 1, b->stream));
      b->launches++;
    }
    // a user boundary function may write anywhere; the axis regularisation rewrites B_theta on an ACTIVE face
    b->vcFieldFresh = !b->userBc && !b->haveAxis;
  }
  KTRY(idfxk::dtControlEnd(b->state, b->redBits, d.cfl, d.cfl_max_var, d.fixed_dt, b->ncyclesDev, b->statusDev,
                           b->mathFlag, b->stream));
  b->launches++;
  return 0;
}

int fieldPtr(IdfxBlock *b, int field, double **p, int dims[4]) {
  const DevGrid &g = b->g;
  dims[1] = g.nx[2];
  dims[2] = g.nx[1];
  dims[3] = g.nx[0];
  dims[0] = 1;
  switch (field) {
    case IDFX_FIELD_VC: *p = b->Vc; dims[0] = b->cfg.nvar + b->cfg.ntr; break;
    case IDFX_FIELD_UC: *p = b->Uc; dims[0] = b->cfg.nvar + b->cfg.ntr; break;
    case IDFX_FIELD_UC0: *p = b->Uc0; dims[0] = b->cfg.nvar + b->cfg.ntr; break;
    case IDFX_FIELD_VS:
    case IDFX_FIELD_VS0:
      *p = (field == IDFX_FIELD_VS) ? b->Vs : b->Vs0;
      dims[0] = b->cfg.mhd ? b->desc.ndim : 0;
      dims[1] = g.nx[2] + g.off[2];
      dims[2] = g.nx[1] + g.off[1];
      dims[3] = g.nx[0] + g.off[0];
      break;
    case IDFX_FIELD_VE:
    case IDFX_FIELD_VE0:
      *p = (field == IDFX_FIELD_VE) ? b->Ve : b->Ve0;
      dims[0] = b->nVe;
      dims[1] = g.nx[2] + g.off[2];
      dims[2] = g.nx[1] + g.off[1];
      dims[3] = g.nx[0] + g.off[0];
      break;
    case IDFX_FIELD_INVDT: *p = b->invDt; break;
    case IDFX_FIELD_EX: *p = b->emf.ex; break;
    case IDFX_FIELD_EY: *p = b->emf.ey; break;
    case IDFX_FIELD_EZ: *p = b->emf.ez; break;
    default: return fail("unknown field id");
  }
  return 0;
}

}  // namespace

// =============================================================================================
extern "C" {

const char *idfx_last_error(void) { return g_error.c_str(); }
const char *idfx_version(void) { return "idefix_b200 0.1 (sm_100a, FP64, fmad=false)"; }

int idfx_device_count(int *count) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    *count = 0;
    return fail(std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
  }
  *count = n;
  return 0;
}

int idfx_block_create(const IdfxBlockDesc *desc, IdfxBlock **out) {
  if (!desc || !out) return fail("null argument");
  *out = nullptr;
  if (validate(*desc)) return 1;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("no CUDA device available: idefix_b200 has no CPU fallback");
  if (desc->device < 0 || desc->device >= ndev) return fail("invalid device ordinal");
  CUDA_TRY(cudaSetDevice(desc->device));

  IdfxBlock *b = new IdfxBlock();
  b->desc = *desc;
  // own the coordinate arrays
  b->hostCoords.resize(12);
  for (int dir = 0; dir < 3; dir++) {
    const int n = desc->np_int[dir] + 2 * desc->nghost[dir];
    const double *src[4] = {desc->x[dir], desc->xl[dir], desc->xr[dir], desc->dx[dir]};
    for (int q = 0; q < 4; q++) b->hostCoords[dir * 4 + q].assign(src[q], src[q] + n);
    b->desc.x[dir] = b->hostCoords[dir * 4 + 0].data();
    b->desc.xl[dir] = b->hostCoords[dir * 4 + 1].data();
    b->desc.xr[dir] = b->hostCoords[dir * 4 + 2].data();
    b->desc.dx[dir] = b->hostCoords[dir * 4 + 3].data();
  }
  setupPhys(b);
  // Axis::Axis (axis.hpp:116-136): which sides of x2 lie on the polar axis, and whether x3 goes all the way round
  b->axisLeft = desc->ndim >= 2 && desc->lbound[1] == IDFX_BC_AXIS;
  b->axisRight = desc->ndim >= 2 && desc->rbound[1] == IDFX_BC_AXIS;
  b->haveAxis = b->axisLeft || b->axisRight;
  if (b->haveAxis) {
    const int kb = desc->nghost[2], ke = desc->nghost[2] + desc->np_int[2];
    b->axisTwoPi = std::fabs(b->desc.xr[2][ke - 1] - b->desc.xl[2][kb] - 2.0 * M_PI) < 1e-10;
  }
  CUDA_TRY(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
  if (std::getenv("IDFX_NO_GRAPH")) b->useGraph = false;   // A/B switch: always launch the cycle kernel by kernel
  if (std::getenv("IDFX_NO_OVERLAP")) b->overlap = false;  // A/B switch: boundaries strictly before the sweeps
  if (std::getenv("IDFX_OVERLAP")) b->forceOverlap = true; // ... or overlapped even on a single block
  CUDA_TRY(cudaStreamCreateWithFlags(&b->sideStream, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreateWithFlags(&b->evShell, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&b->evGhosts, cudaEventDisableTiming));

  DevGrid &g = b->g;
  std::memset(&g, 0, sizeof(g));
  g.ndim = desc->ndim;
  g.geometry = desc->geometry;
  g.regular = desc->regular_grid ? 1 : 0;
  for (int dir = 0; dir < 3; dir++) {
    g.ng[dir] = desc->nghost[dir];
    g.nx[dir] = desc->np_int[dir] + 2 * desc->nghost[dir];
    g.beg[dir] = desc->nghost[dir];
    g.end[dir] = desc->nghost[dir] + desc->np_int[dir];
    g.off[dir] = dir < desc->ndim ? 1 : 0;
  }
  g.sj = g.nx[0];
  g.sk = (long long)g.nx[0] * g.nx[1];
  g.sn = g.sk * g.nx[2];
  g.vsj = g.nx[0] + g.off[0];
  g.vsk = g.vsj * (g.nx[1] + g.off[1]);
  g.vsn = g.vsk * (g.nx[2] + g.off[2]);
  b->nCells = g.sn;
  b->nVs = g.vsn;

  int rc = 0;
  rc |= buildGeometry(b);
  const int nv = b->cfg.nvar + b->cfg.ntr;
  rc = rc || devAlloc(b, &b->mathFlag, 1);
  if (b->cfg.ntr > 0) rc = rc || devAlloc(b, &b->fluxRho, b->nCells);
  if (!b->cfg.mhd && b->cfg.ndim == 1 && b->cfg.solver == SOLVER_ROE) rc = rc || devAlloc(b, &b->staleFlux, b->nCells * b->cfg.nvar);
  if (desc->shock_flattening > 0.0) rc = rc || devAlloc(b, &b->shockFlags, b->nCells);
  rc = rc || devAlloc(b, &b->Vc, b->nCells * nv);
  rc = rc || devAlloc(b, &b->Uc, b->nCells * nv);
  rc = rc || devAlloc(b, &b->invDt, b->nCells);
  if (desc->nstages > 1) rc = rc || devAlloc(b, &b->Uc0, b->nCells * nv);
  std::memset(&b->emf, 0, sizeof(b->emf));
  if (b->cfg.mhd) {
    rc = rc || devAlloc(b, &b->Vs, b->nVs * desc->ndim);
    if (desc->nstages > 1) rc = rc || devAlloc(b, &b->Vs0, b->nVs * desc->ndim);
    if (desc->evolve_vector_potential) {
      b->nVe = (desc->ndim == 3) ? 3 : 1;
      rc = rc || devAlloc(b, &b->Ve, b->nVs * b->nVe);
      if (desc->nstages > 1) rc = rc || devAlloc(b, &b->Ve0, b->nVs * b->nVe);
    }
    DevEmf &e = b->emf;
    const long long n = b->nCells;
    const int nd = desc->ndim;
    if (nd >= 2) {
      rc = rc || devAlloc(b, &e.ez, n) || devAlloc(b, &e.ezi, n) || devAlloc(b, &e.ezj, n);
      if (nd == 3 || desc->ncomp == 3) rc = rc || devAlloc(b, &e.eyi, n) || devAlloc(b, &e.exj, n);
      if (nd == 3)
        rc = rc || devAlloc(b, &e.ex, n) || devAlloc(b, &e.ey, n) || devAlloc(b, &e.exk, n) || devAlloc(b, &e.eyk, n);
      if (desc->emf == IDFX_EMF_UCT0 || desc->emf == IDFX_EMF_UCT_CONTACT) {
        rc = rc || devAlloc(b, &e.Ex3, n);
        if (nd == 3) rc = rc || devAlloc(b, &e.Ex1, n) || devAlloc(b, &e.Ex2, n);
      }
      if (desc->emf == IDFX_EMF_UCT_CONTACT) {
        rc = rc || devAlloc(b, &e.svx, n) || devAlloc(b, &e.svy, n);
        if (nd == 3) rc = rc || devAlloc(b, &e.svz, n);
      }
      if (desc->emf == IDFX_EMF_UCT_HLL || desc->emf == IDFX_EMF_UCT_HLLD) {
        rc = rc || devAlloc(b, &e.axL, n) || devAlloc(b, &e.axR, n) || devAlloc(b, &e.dxL, n) || devAlloc(b, &e.dxR, n);
        rc = rc || devAlloc(b, &e.ayL, n) || devAlloc(b, &e.ayR, n) || devAlloc(b, &e.dyL, n) || devAlloc(b, &e.dyR, n);
        if (nd == 3)
          rc = rc || devAlloc(b, &e.azL, n) || devAlloc(b, &e.azR, n) || devAlloc(b, &e.dzL, n) || devAlloc(b, &e.dzR, n);
      }
    } else {
      // 1-D MHD: the sweep still stores the transverse EMFs; give them somewhere to go
      rc = rc || devAlloc(b, &e.ezi, n) || devAlloc(b, &e.eyi, n) || devAlloc(b, &e.svx, n);
      rc = rc || devAlloc(b, &e.axL, n) || devAlloc(b, &e.axR, n) || devAlloc(b, &e.dxL, n) || devAlloc(b, &e.dxR, n);
    }
  }
  if (!rc && !std::getenv("IDFX_NO_TMA")) {   // A/B switch: per-lane cp.async loads in the x1 sweep
    b->tmRowCells = 32 + 2 * ((desc->order == 4) ? 2 : 1);
    b->haveTmRow = tma::encodeCellBox(&b->tmRow, b->Vc, g.nx[0], g.nx[1], g.nx[2], b->cfg.nvar, b->tmRowCells);
  }
  rc = rc || devAlloc(b, &b->state, 4);
  rc = rc || devAlloc(b, &b->redBits, 2);
  rc = rc || devAlloc(b, &b->ncyclesDev, 1);
  rc = rc || devAlloc(b, &b->statusDev, 1);
  rc = rc || devAlloc(b, &b->nanDev, 1);
  if (rc) {
    idfx_block_destroy(b);
    return 1;
  }
  if (cudaMallocHost((void **)&b->hostState, 16 * sizeof(double)) != cudaSuccess) {
    idfx_block_destroy(b);
    return fail("cudaMallocHost failed");
  }
  // TimeIntegrator ctor (timeIntegrator.cpp:23-85)
  if (desc->nstages == 2) {
    b->wc[0] = 0.5;
    b->w0[0] = 0.5;
  }
  if (desc->nstages == 3) {
    b->wc[0] = 0.25;
    b->w0[0] = 0.75;
    b->wc[1] = 2.0 / 3.0;
    b->w0[1] = 1.0 / 3.0;
  }
  b->hostState[0] = 0.0;
  b->hostState[1] = desc->fixed_dt > 0.0 ? desc->fixed_dt : desc->first_dt;
  b->hostState[2] = 0.0;
  b->hostState[3] = -1.0;
  cudaMemcpyAsync(b->state, b->hostState, 4 * sizeof(double), cudaMemcpyHostToDevice, b->stream);
  for (int dir = 0; dir < desc->ndim; dir++)
    if (desc->lbound[dir] == IDFX_BC_USERDEF || desc->rbound[dir] == IDFX_BC_USERDEF) b->haveUserdef = true;
  b->fused = !b->haveUserdef;
  if (cudaStreamSynchronize(b->stream) != cudaSuccess) {
    idfx_block_destroy(b);
    return fail("device initialisation failed");
  }
  *out = b;
  return 0;
}

int idfx_block_destroy(IdfxBlock *b) {
  if (b) dropCycleGraph(b);
  if (!b) return 0;
  cudaSetDevice(b->desc.device);
  if (b->stream) cudaStreamSynchronize(b->stream);
  for (void *p : b->allocs) cudaFree(p);
  if (b->hostState) cudaFreeHost(b->hostState);
  for (cudaEvent_t e : b->evPool) cudaEventDestroy(e);
  if (b->evShell) cudaEventDestroy(b->evShell);
  if (b->evGhosts) cudaEventDestroy(b->evGhosts);
  if (b->sideStream) cudaStreamDestroy(b->sideStream);
  if (b->stream && b->ownStream) cudaStreamDestroy(b->stream);
  delete b;
  return 0;
}

int idfx_sync(IdfxBlock *b) {
  cudaSetDevice(b->desc.device);   // one process may drive blocks on several devices
  CUDA_TRY(cudaStreamSynchronize(b->stream));
  return 0;
}

int idfx_field_info(IdfxBlock *b, int field, double **dev, int dims[4]) {
  cudaSetDevice(b->desc.device);   // one process may drive blocks on several devices
  double *p = nullptr;
  if (fieldPtr(b, field, &p, dims)) return 1;
  if (dev) *dev = p;
  return 0;
}

int idfx_upload(IdfxBlock *b, int field, const double *host) {
  cudaSetDevice(b->desc.device);   // one process may drive blocks on several devices
  double *p = nullptr;
  int dims[4];
  if (fieldPtr(b, field, &p, dims)) return 1;
  if (!p) return fail("field is not allocated in this configuration");
  const size_t n = (size_t)dims[0] * dims[1] * dims[2] * dims[3];
  CUDA_TRY(cudaMemcpyAsync(p, host, n * sizeof(double), cudaMemcpyHostToDevice, b->stream));
  CUDA_TRY(cudaStreamSynchronize(b->stream));
  return 0;
}

int idfx_download(IdfxBlock *b, int field, double *host) {
  cudaSetDevice(b->desc.device);   // one process may drive blocks on several devices
  double *p = nullptr;
  int dims[4];
  if (fieldPtr(b, field, &p, dims)) return 1;
  if (!p) return fail("field is not allocated in this configuration");
  const size_t n = (size_t)dims[0] * dims[1] * dims[2] * dims[3];
  CUDA_TRY(cudaMemcpyAsync(host, p, n * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
  CUDA_TRY(cudaStreamSynchronize(b->stream));
  return 0;
}

int idfx_set_time(IdfxBlock *b, double t, double dt) {
  cudaSetDevice(b->desc.device);   // one process may drive blocks on several devices
  CUDA_TRY(cudaStreamSynchronize(b->stream));
  b->hostState[8] = t;
  b->hostState[9] = dt;
  CUDA_TRY(cudaMemcpyAsync(b->state, b->hostState + 8, 2 * sizeof(double), cudaMemcpyHostToDevice, b->stream));
  CUDA_TRY(cudaStreamSynchronize(b->stream));
  return 0;
}

int idfx_get_time(IdfxBlock *b, double *t, double *dt) {
  cudaSetDevice(b->desc.device);   // one process may drive blocks on several devices
  if (readState(b)) return 1;
  if (t) *t = b->hostState[0];
  if (dt) *dt = b->hostState[1];
  return 0;
}

int idfx_set_tstop(IdfxBlock *b, double tstop) {
  cudaSetDevice(b->desc.device);   // one process may drive blocks on several devices
  CUDA_TRY(cudaStreamSynchronize(b->stream));
  b->hostState[10] = tstop;
  CUDA_TRY(cudaMemcpyAsync(b->state + 3, b->hostState + 10, sizeof(double), cudaMemcpyHostToDevice, b->stream));
  CUDA_TRY(cudaStreamSynchronize(b->stream));
  return 0;
}

int idfx_get_ncycles(IdfxBlock *b, long long *n) {
  cudaSetDevice(b->desc.device);   // one process may drive blocks on several devices
  CUDA_TRY(cudaMemcpyAsync(b->hostState + 12, b->ncyclesDev, sizeof(long long), cudaMemcpyDeviceToHost, b->stream));
  CUDA_TRY(cudaStreamSynchronize(b->stream));
  std::memcpy(n, b->hostState + 12, sizeof(long long));
  return 0;
}

int idfx_launch_count(IdfxBlock *b, long long *count) {
  cudaSetDevice(b->desc.device);   // one process may drive blocks on several devices
  *count = b->launches;
  return 0;
}

int idfx_set_stream(IdfxBlock *b, void *stream) {
  cudaSetDevice(b->desc.device);   // one process may drive blocks on several devices
  CUDA_TRY(cudaStreamSynchronize(b->stream));
  dropCycleGraph(b);
  if (b->ownStream && b->stream) cudaStreamDestroy(b->stream);
  b->stream = (cudaStream_t)stream;
  b->ownStream = false;
  return 0;
}

// mode 1: start sampling, 0: stop, 2: collect into out[2*TIME_NCAT] = {launch brackets, total ms} per class
// (X1, X2, X3 sweeps, constrained transport, stage epilogue, boundaries) and reset the samples.
int idfx_timing(IdfxBlock *b, int mode, double *out) {
  cudaSetDevice(b->desc.device);   // one process may drive blocks on several devices
  if (mode == 0 || mode == 1) {
    b->timing = (mode == 1);
    if (mode == 1) b->evUsed = 0;
    return 0;
  }
  if (mode != 2 || !out) return fail("idfx_timing: mode must be 0, 1 or 2 (with an output array)");
  CUDA_TRY(cudaStreamSynchronize(b->stream));
  for (int c = 0; c < 2 * TIME_NCAT; c++) out[c] = 0.0;
  for (size_t n = 0; n < b->evUsed; n++) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, b->evPool[2 * n], b->evPool[2 * n + 1]) != cudaSuccess) continue;
    out[2 * b->evCat[n]] += 1.0;
    out[2 * b->evCat[n] + 1] += (double)ms;
  }
  b->evUsed = 0;
  return 0;
}

int idfx_set_nproc(IdfxBlock *b, const int nproc[3]) {
  cudaSetDevice(b->desc.device);   // one process may drive blocks on several devices
  dropCycleGraph(b);
  for (int d = 0; d < 3; d++) b->nproc[d] = nproc[d] < 1 ? 1 : nproc[d];
  return 0;
}

// ---- granular entry points -----------------------------------------------------------------
int idfx_reset_stage(IdfxBlock *b) {
  cudaSetDevice(b->desc.device);   // one process may drive blocks on several devices
  KTRY(idfxk::fill(b->invDt, b->nCells, 0.0, b->stream));
  b->launches++;
  return 0;
}

int idfx_enforce_boundary_dir(IdfxBlock *b, double t, int dir) {
  cudaSetDevice(b->desc.device);   // one process may drive blocks on several devices
  if (dir < 0 || dir >= b->desc.ndim) return fail("idfx_enforce_boundary_dir: direction out of range");
  b->vcFieldFresh = false;
  TimeScope ts(b, TIME_BOUNDARY);
  return enforceBoundaryDirImpl(b, dir, t, true);
}

int idfx_set_boundaries(IdfxBlock *b, double t) {
  cudaSetDevice(b->desc.device);   // one process may drive blocks on several devices
  if (setBoundariesCore(b, t, true)) return 1;
  return reconstructVc(b, b->Vc);
}

int idfx_reconstruct_vc_field(IdfxBlock *b, int field) {
  cudaSetDevice(b->desc.device);   // one process may drive blocks on several devices
  if (field == IDFX_FIELD_VC) return reconstructVc(b, b->Vc);
  if (field == IDFX_FIELD_UC) return reconstructVc(b, b->Uc);
  return fail("idfx_reconstruct_vc_field: field must be VC or UC");
}

int idfx_prim_to_cons(IdfxBlock *b) {
  cudaSetDevice(b->desc.device);   // one process may drive blocks on several devices
  KTRY(idfxk::primToCons(b->cfg, b->g, b->Vc, b->Uc, b->desc.gamma, b->stream));
  b->launches++;
  if (b->cfg.ntr > 0) {
    KTRY(idfxk::tracerPrimToCons(b->cfg, b->g, b->Vc, b->Uc, nullptr, b->stream));
    b->launches++;
  }
  return 0;
}

int idfx_derive_vector_potential(IdfxBlock *b) {
  cudaSetDevice(b->desc.device);
  if (!b->Ve) return 0;
  b->vcFieldFresh = false;
  KTRY(idfxk::magFieldFromA(b->g, b->Ve, b->Vs, b->haveAxis ? 1 : 0, b->stream, &b->launches));
  return 0;
}

int idfx_cons_to_prim(IdfxBlock *b) {
  cudaSetDevice(b->desc.device);   // one process may drive blocks on several devices
  if (b->Ve) KTRY(idfxk::magFieldFromA(b->g, b->Ve, b->Vs, b->haveAxis ? 1 : 0, b->stream, &b->launches));   // convertConsToPrim.hpp:117-119
  if (reconstructVc(b, b->Uc)) return 1;
  KTRY(idfxk::consToPrim(b->cfg, b->g, b->Vc, b->Uc, b->desc.gamma, b->stream));
  b->launches++;
  if (b->cfg.ntr > 0) {
    KTRY(idfxk::tracerConsToPrim(b->cfg, b->g, b->Vc, b->Uc, nullptr, 0, 0.0, 0.0, 0, b->stream));
    b->launches++;
  }
  return 0;
}

int idfx_state_copy(IdfxBlock *b) {
  cudaSetDevice(b->desc.device);   // one process may drive blocks on several devices
  if (!b->Uc0) return fail("states[\"begin\"] exists only when nstages > 1");
  KTRY(idfxk::copy(b->Uc0, b->Uc, b->nCells * (b->cfg.nvar + b->cfg.ntr), b->stream));
  b->launches++;
  if (b->cfg.mhd) {
    if (b->Ve) KTRY(idfxk::copy(b->Ve0, b->Ve, b->nVs * b->nVe, b->stream));
    else KTRY(idfxk::copy(b->Vs0, b->Vs, b->nVs * b->desc.ndim, b->stream));
    b->launches++;
  }
  return 0;
}

int idfx_state_add_and_store(IdfxBlock *b, double wc, double w0) {
  cudaSetDevice(b->desc.device);   // one process may drive blocks on several devices
  if (!b->Uc0) return fail("states[\"begin\"] exists only when nstages > 1");
  KTRY(idfxk::addAndStore(b->Uc, b->Uc0, b->nCells * (b->cfg.nvar + b->cfg.ntr), wc, w0, b->stream));
  b->launches++;
  if (b->cfg.mhd) {
    if (b->Ve) KTRY(idfxk::addAndStore(b->Ve, b->Ve0, b->nVs * b->nVe, wc, w0, b->stream));
    else KTRY(idfxk::addAndStore(b->Vs, b->Vs0, b->nVs * b->desc.ndim, wc, w0, b->stream));
    b->launches++;
  }
  return 0;
}

int idfx_evolve_stage(IdfxBlock *b, double t, double dt) {
  cudaSetDevice(b->desc.device);   // one process may drive blocks on several devices
  b->stageTime = t;
  b->stageTimeKnown = true;
  // dt lives on the device; only touch it when the caller's value differs from the mirror
  if (readState(b)) return 1;
  if (b->hostState[1] != dt) {
    b->hostState[9] = dt;
    CUDA_TRY(cudaMemcpyAsync(b->state + 1, b->hostState + 9, sizeof(double), cudaMemcpyHostToDevice, b->stream));
  }
  const int rc = evolveStageImpl(b, -1);
  b->stageTimeKnown = false;
  return rc;
}

int idfx_compute_timestep(IdfxBlock *b, double *dt_min) {
  cudaSetDevice(b->desc.device);   // one process may drive blocks on several devices
  b->hostState[11] = INFINITY;
  CUDA_TRY(cudaMemcpyAsync(b->redBits, b->hostState + 11, sizeof(double), cudaMemcpyHostToDevice, b->stream));
  if (timestepReduceImpl(b)) return 1;
  crossBlockReduce(b, (double *)b->redBits, IDFX_REDUCE_MIN);
  CUDA_TRY(cudaMemcpyAsync(b->hostState + 11, b->redBits, sizeof(double), cudaMemcpyDeviceToHost, b->stream));
  CUDA_TRY(cudaStreamSynchronize(b->stream));
  *dt_min = b->hostState[11];
  return 0;
}

int idfx_check_divb(IdfxBlock *b, double *divb_max) {
  cudaSetDevice(b->desc.device);   // one process may drive blocks on several devices
  if (!b->cfg.mhd) {
    *divb_max = 0.0;
    return 0;
  }
  CUDA_TRY(cudaMemsetAsync(b->redBits + 1, 0, sizeof(double), b->stream));
  KTRY(idfxk::checkDivB(b->cfg, b->g, b->Vs, b->redBits + 1, b->stream));
  b->launches++;
  crossBlockReduce(b, (double *)(b->redBits + 1), IDFX_REDUCE_MAX);
  CUDA_TRY(cudaMemcpyAsync(b->hostState + 13, b->redBits + 1, sizeof(double), cudaMemcpyDeviceToHost, b->stream));
  CUDA_TRY(cudaStreamSynchronize(b->stream));
  *divb_max = b->hostState[13];
  return 0;
}

int idfx_check_nan(IdfxBlock *b, int *nan_count) {
  cudaSetDevice(b->desc.device);   // one process may drive blocks on several devices
  CUDA_TRY(cudaMemsetAsync(b->nanDev, 0, sizeof(int), b->stream));
  KTRY(idfxk::checkNan(b->cfg, b->g, b->Vc, b->Vs, b->nanDev, b->stream));
  b->launches += b->cfg.mhd ? 2 : 1;
  CUDA_TRY(cudaMemcpyAsync(b->hostState + 14, b->nanDev, sizeof(int), cudaMemcpyDeviceToHost, b->stream));
  CUDA_TRY(cudaStreamSynchronize(b->stream));
  int n;
  std::memcpy(&n, b->hostState + 14, sizeof(int));
  *nan_count = n;
  return 0;
}

// ---- fused path -------------------------------------------------------------------------------
namespace {
// The steady-state cycle (every cycle of a call but the first, which re-validates the caller's state, and the
// last, which leaves Uc behind) is identical launch for launch: it is captured once into a CUDA graph and
// replayed, which removes the per-launch CPU cost that dominates small grids (2-D 1024^2: ~40 launches of
// ~15 us each per 0.7 ms cycle).  Only for a single block without host callbacks on a capturable stream.
bool graphEligible(const IdfxBlock *b, int ncycles) {
  const bool single = b->nproc[0] * b->nproc[1] * b->nproc[2] == 1;
  // a decomposed block qualifies when its exchanges and reductions are the library's own peer-to-peer kernels
  return b->useGraph && b->fused && ncycles >= 4 && !b->userBc && !b->timing && b->stream != nullptr &&
         !b->stageHook[0] && !b->stageHook[1] && ((single && !b->haloFn) || b->p2p);
}
}  // namespace

int idfx_cycles(IdfxBlock *b, int ncycles) {
  cudaSetDevice(b->desc.device);   // one process may drive blocks on several devices
  // the caller holds device pointers to the state (idfx_field_info) and may have changed it since the last call
  b->vcFieldFresh = false;
  b->ghostsDone = false;
  if (graphEligible(b, ncycles)) {
    if (cycleFused(b, false)) return 1;                       // first cycle: eager (full validation of the inputs)
    int done = 1;
    if (!b->cycleGraph) {
      const long long l0 = b->launches;
      cudaGraph_t graph = nullptr;
      if (cudaStreamBeginCapture(b->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        const int rc = cycleFused(b, false);
        const cudaError_t ce = cudaStreamEndCapture(b->stream, &graph);
        if (rc == 0 && ce == cudaSuccess && graph != nullptr &&
            cudaGraphInstantiate(&b->cycleGraph, graph, 0) == cudaSuccess) {
          b->cycleGraphLaunches = b->launches - l0;
        } else {
          b->cycleGraph = nullptr;
          b->useGraph = false;                                 // never try again on this block
        }
        if (graph) cudaGraphDestroy(graph);
        b->launches = l0;                                      // the capture itself launched nothing
      } else {
        b->useGraph = false;
      }
      cudaGetLastError();                                      // clear a capture failure, the eager path follows
    }
    if (b->cycleGraph) {
      for (; done < ncycles - 1; done++) {
        CUDA_TRY(cudaGraphLaunch(b->cycleGraph, b->stream));
        b->launches += b->cycleGraphLaunches;
      }
    }
    for (; done < ncycles; done++)
      if (cycleFused(b, done == ncycles - 1)) return 1;
    return 0;
  }
  for (int n = 0; n < ncycles; n++) {
    if (b->fused) {
      if (cycleFused(b, n == ncycles - 1)) return 1;
    } else {
      if (cycleReferenceOrder(b)) return 1;
    }
  }
  return 0;
}

int idfx_cycle(IdfxBlock *b, double *t, double *dt, int *status_flags) {
  cudaSetDevice(b->desc.device);   // one process may drive blocks on several devices
  if (idfx_cycles(b, 1)) return 1;
  CUDA_TRY(cudaMemcpyAsync(b->hostState, b->state, 4 * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
  CUDA_TRY(cudaMemcpyAsync(b->hostState + 14, b->statusDev, sizeof(int), cudaMemcpyDeviceToHost, b->stream));
  CUDA_TRY(cudaStreamSynchronize(b->stream));
  if (t) *t = b->hostState[0];
  if (dt) *dt = b->hostState[1];
  int s;
  std::memcpy(&s, b->hostState + 14, sizeof(int));
  if (status_flags) *status_flags = s;
  return 0;
}

// Number of threads whose branch-free division / square root failed its range check since the block was
// created (fastmath.cuh).  Zero = every quotient and root of the sweeps was the correctly rounded IEEE one.
int idfx_math_exceptions(IdfxBlock *b, unsigned int *count) {
  cudaSetDevice(b->desc.device);   // one process may drive blocks on several devices
  CUDA_TRY(cudaMemcpyAsync(count, b->mathFlag, sizeof(unsigned int), cudaMemcpyDeviceToHost, b->stream));
  CUDA_TRY(cudaStreamSynchronize(b->stream));
  return 0;
}

// selects the reference-ordered (un-fused) kernel sequence; used by the parity tests
int idfx_set_fused(IdfxBlock *b, int fused) {
  cudaSetDevice(b->desc.device);   // one process may drive blocks on several devices
  dropCycleGraph(b);
  if (fused && b->haveUserdef) return fail("the fused path is not available with user-defined boundaries");
  b->fused = fused != 0;
  return 0;
}

// ---- callbacks -----------------------------------------------------------------------------------
int idfx_enroll_userdef_boundary(IdfxBlock *b, IdfxUserBoundaryFn fn, void *user) {
  cudaSetDevice(b->desc.device);   // one process may drive blocks on several devices
  dropCycleGraph(b);
  b->userBc = fn;
  b->userBcData = user;
  return 0;
}
int idfx_enroll_stage_hook(IdfxBlock *b, int which, IdfxStageHookFn fn, void *user) {
  cudaSetDevice(b->desc.device);
  if (which != IDFX_HOOK_INTERNAL_BOUNDARY && which != IDFX_HOOK_EMF_BOUNDARY) return fail("unknown stage hook");
  if (which == IDFX_HOOK_EMF_BOUNDARY && !(b->cfg.mhd && b->desc.ndim >= 2))
    return fail("EMF boundary hooks need constrained transport (MHD, DIMENSIONS >= 2)");
  dropCycleGraph(b);
  b->stageHook[which] = fn;
  b->stageHookData[which] = user;
  return 0;
}
int idfx_enroll_reduce(IdfxBlock *b, IdfxReduceFn fn, void *user) {
  cudaSetDevice(b->desc.device);   // one process may drive blocks on several devices
  dropCycleGraph(b);
  b->reduceFn = fn;
  b->reduceData = user;
  return 0;
}
int idfx_enroll_halo_exchange(IdfxBlock *b, IdfxHaloFn fn, void *user) {
  cudaSetDevice(b->desc.device);   // one process may drive blocks on several devices
  dropCycleGraph(b);
  b->haloFn = fn;
  b->haloData = user;
  return 0;
}

// ---- halo exchange building blocks ------------------------------------------------------------------
int idfx_halo_buffer_size(IdfxBlock *b, int dir, size_t *ndoubles) {
  cudaSetDevice(b->desc.device);   // one process may drive blocks on several devices
  if (dir < 0 || dir >= b->desc.ndim) return fail("invalid direction");
  *ndoubles = (size_t)idfxk::haloBufferSize(b->cfg, b->g, dir);
  return 0;
}

static int ensureHaloBuffers(IdfxBlock *b, int dir) {
  if (b->haloSend[dir][0]) return 0;
  const long long n = idfxk::haloBufferSize(b->cfg, b->g, dir);
  for (int s = 0; s < 2; s++) {
    if (devAlloc(b, &b->haloSend[dir][s], n)) return 1;
    if (devAlloc(b, &b->haloRecv[dir][s], n)) return 1;
  }
  CUDA_TRY(cudaStreamSynchronize(b->stream));
  return 0;
}

int idfx_halo_buffers(IdfxBlock *b, int dir, int side, double **send, double **recv) {
  cudaSetDevice(b->desc.device);   // one process may drive blocks on several devices
  if (dir < 0 || dir >= b->desc.ndim || side < 0 || side > 1) return fail("invalid direction / side");
  if (ensureHaloBuffers(b, dir)) return 1;
  if (send) *send = b->haloSend[dir][side];
  if (recv) *recv = b->haloRecv[dir][side];
  return 0;
}

int idfx_halo_pack(IdfxBlock *b, int dir) {
  cudaSetDevice(b->desc.device);   // one process may drive blocks on several devices
  if (dir < 0 || dir >= b->desc.ndim) return fail("invalid direction");
  if (ensureHaloBuffers(b, dir)) return 1;
  for (int side = 0; side < 2; side++) {
    if (!sideFacesBlock(b, dir, side)) continue;
    // with a bound peer the slab is written straight into the neighbour's receive buffer (NVLink P2P store)
    double *dst = b->haloPeer[dir][side] ? b->haloPeer[dir][side] : b->haloSend[dir][side];
    KTRY(idfxk::haloPack(b->cfg, b->g, b->Vc, b->Vs, dir, side, dst, b->stream, &b->launches));
  }
  return 0;
}

int idfx_halo_unpack(IdfxBlock *b, int dir) {
  cudaSetDevice(b->desc.device);   // one process may drive blocks on several devices
  if (dir < 0 || dir >= b->desc.ndim) return fail("invalid direction");
  if (ensureHaloBuffers(b, dir)) return 1;
  for (int side = 0; side < 2; side++) {
    if (!sideFacesBlock(b, dir, side)) continue;
    KTRY(idfxk::haloUnpack(b->cfg, b->g, b->Vc, b->Vs, dir, side, b->haloRecv[dir][side], b->stream, &b->launches));
  }
  return 0;
}

int idfx_halo_bind_peer(IdfxBlock *b, int dir, int side, double *peer_recv) {
  cudaSetDevice(b->desc.device);   // one process may drive blocks on several devices
  if (dir < 0 || dir >= b->desc.ndim || side < 0 || side > 1) return fail("invalid direction / side");
  b->haloPeer[dir][side] = peer_recv;
  return 0;
}

int idfx_halo_ipc_handle(IdfxBlock *b, int dir, int side, void *handle64) {
  cudaSetDevice(b->desc.device);   // one process may drive blocks on several devices
  if (dir < 0 || dir >= b->desc.ndim || side < 0 || side > 1) return fail("invalid direction / side");
  if (ensureHaloBuffers(b, dir)) return 1;
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, b->haloRecv[dir][side]));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  std::memcpy(handle64, &h, 64);
  return 0;
}

// ---- peer-to-peer exchange through device mailboxes ------------------------------------------------------------------
static int ensureMailbox(IdfxBlock *b) {
  if (b->mailbox) return 0;
  if (devAlloc(b, &b->mailbox, idfxk::P2P_MB_WORDS) || devAlloc(b, &b->p2pSeq, idfxk::P2P_SEQ_WORDS) ||
      devAlloc(b, &b->p2pTicket, 2))
    return 1;
  CUDA_TRY(cudaStreamSynchronize(b->stream));
  return 0;
}

int idfx_p2p_mailbox(IdfxBlock *b, void **mailbox, void *ipc_handle64) {
  cudaSetDevice(b->desc.device);
  if (ensureMailbox(b)) return 1;
  if (mailbox) *mailbox = b->mailbox;
  if (ipc_handle64) {
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, b->mailbox));
    std::memcpy(ipc_handle64, &h, 64);
  }
  return 0;
}

int idfx_p2p_connect(IdfxBlock *b, int rank, int world, void *const peer_recv[3][2], void *const peer_mailbox[3][2],
                     void *const *all_mailbox) {
  cudaSetDevice(b->desc.device);
  if (world < 1 || world > idfxk::P2P_MAX_RANKS || rank < 0 || rank >= world)
    return fail("idfx_p2p_connect: rank / world out of range");
  if (ensureMailbox(b)) return 1;
  KTRY(idfxk::p2pPreload());
  dropCycleGraph(b);
  for (int d = 0; d < b->desc.ndim; d++) {
    if (b->nproc[d] <= 1) continue;
    if (ensureHaloBuffers(b, d)) return 1;
    for (int s = 0; s < 2; s++) {
      if (!sideFacesBlock(b, d, s)) continue;
      if (!peer_recv[d][s] || !peer_mailbox[d][s]) return fail("idfx_p2p_connect: missing neighbour pointers");
      b->haloPeer[d][s] = (double *)peer_recv[d][s];
      b->peerMailbox[d][s] = (unsigned long long *)peer_mailbox[d][s];
    }
  }
  for (int q = 0; q < world; q++) {
    if (!all_mailbox || !all_mailbox[q]) return fail("idfx_p2p_connect: missing mailbox of a block of the domain");
    b->allMailbox[q] = (unsigned long long *)all_mailbox[q];
  }
  b->p2pRank = rank;
  b->p2pWorld = world;
  b->p2p = true;
  return 0;
}

int idfx_p2p_timeouts(IdfxBlock *b, unsigned long long *count) {
  cudaSetDevice(b->desc.device);
  *count = 0;
  if (!b->p2pSeq) return 0;
  CUDA_TRY(cudaMemcpyAsync(count, b->p2pSeq + idfxk::P2P_SEQ_TIMEOUTS, sizeof(*count), cudaMemcpyDeviceToHost, b->stream));
  CUDA_TRY(cudaStreamSynchronize(b->stream));
  return 0;
}

int idfx_p2p_enable_peer_access(int device, int peer_device) {
  if (device == peer_device) return 0;
  CUDA_TRY(cudaSetDevice(device));
  int can = 0;
  CUDA_TRY(cudaDeviceCanAccessPeer(&can, device, peer_device));
  if (!can) return fail("device cannot access its peer's memory");
  const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(cudaGetErrorString(e));
  cudaGetLastError();
  return 0;
}

int idfx_halo_open_ipc(IdfxBlock *b, const void *handle64, double **ptr) {
  cudaSetDevice(b->desc.device);   // one process may drive blocks on several devices
  (void)b;
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle64, 64);
  void *p = nullptr;
  CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *ptr = (double *)p;
  return 0;
}

}  // extern "C"
