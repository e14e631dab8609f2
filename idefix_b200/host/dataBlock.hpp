// dataBlock.hpp -- DataBlock and Fluid<Phys> / Hydro on top of the C ABI
// (reference: src/dataBlock/dataBlock.hpp:54-183, src/fluid/fluid.hpp:53-247,
// src/fluid/enroll.hpp:13-129, src/fluid/boundary/boundary.hpp:28-35).
//
// The public data members user set-ups touch are kept (data.x, data.beg/end/np_tot/nghost,
// data.t/dt, data.hydro->Vc/Vs/Uc/InvDt ...).  The field arrays are non-owning views of the
// device memory owned by the IdfxBlock, so a user's idefix_for kernel, the DataBlockHost mirror
// and the library's kernels all work on the same (n,k,j,i) arrays.  Every method that launched
// Kokkos kernels in the reference forwards to one idfx_* entry point.
#include "idefix.hpp"
#ifndef IDEFIX_B200_HOST_DATABLOCK_HPP_
#define IDEFIX_B200_HOST_DATABLOCK_HPP_

class DataBlock;
struct DefaultPhysics {};
template <typename Phys> class Fluid;
using Hydro = Fluid<DefaultPhysics>;

template <typename Phys>
using UserDefBoundaryFunc = void (*)(Fluid<Phys> *, int dir, BoundarySide side, const real t);
using UserDefBoundaryFuncOld = void (*)(DataBlock &, int dir, BoundarySide side, const real t);
template <typename Phys>
using InternalBoundaryFunc = void (*)(Fluid<Phys> *, const real t);
using InternalBoundaryFuncOld = void (*)(DataBlock &, const real t);
using EmfBoundaryFunc = void (*)(DataBlock &, const real t);
using StepFunc = void (*)(DataBlock &, const real t, const real dt);

// Stand-in for Fluid::emf (ConstrainedTransport, constrainedTransport.hpp:22-144): the corner EMFs a user EMF
// boundary function rewrites
struct ConstrainedTransport {
  IdefixArray3D<real> ex, ey, ez;
};

enum Solver { TVDLF = IDFX_TVDLF, HLL = IDFX_HLL, HLLC = IDFX_HLLC, HLLD = IDFX_HLLD, ROE = IDFX_ROE };

// Stand-in for Fluid::boundary (the reference exposes SetBoundaries / EnforceBoundaryDir there)
template <typename Phys>
class Boundary {
 public:
  explicit Boundary(Fluid<Phys> *f) : fluid(f) {}
  void SetBoundaries(real t);
  void EnforceBoundaryDir(real t, int dir);
  void ReconstructVcField(IdefixArray4D<real> &);

 private:
  Fluid<Phys> *fluid;
};

template <typename Phys>
class Fluid {
 public:
  IdefixArray4D<real> Vc;      // primitive variables (nvar,k,j,i)
  IdefixArray4D<real> Vs;      // staggered field (ndim,k+1,j+1,i+1)
  IdefixArray4D<real> Ve;      // edge-centred vector potential (only with EVOLVE_VECTOR_POTENTIAL)
  std::vector<std::string> VeName;
  IdefixArray4D<real> Uc;      // conservative variables
  IdefixArray3D<real> InvDt;   // inverse time step accumulator
  std::vector<std::string> VcName, VsName;
  DataBlock *data{nullptr};
  std::unique_ptr<Boundary<Phys>> boundary;
  std::unique_ptr<ConstrainedTransport> emf;   // corner EMFs (data.hydro->emf->ex ..., as in the reference)
  Solver mySolver{HLL};
  int emfScheme{IDFX_EMF_UCT_CONTACT};         // [Hydro] emf
  real gamma{5.0 / 3.0};
  real csiso{0.0};
  bool haveTracer{false};
  int nTracer{0};
  real shockFlattening{0.0};   // [Hydro] shockFlattening <smoothing>, 0 = off
  bool haveSourceTerms{GEOMETRY != CARTESIAN};
  bool haveUserSourceTerm{false};
  std::string prefix{"Hydro"};
  UserDefBoundaryFunc<Phys> userDefBoundaryFunc{nullptr};
  UserDefBoundaryFuncOld userDefBoundaryFuncOld{nullptr};
  InternalBoundaryFunc<Phys> internalBoundaryFunc{nullptr};
  InternalBoundaryFuncOld internalBoundaryFuncOld{nullptr};
  EmfBoundaryFunc emfBoundaryFunc{nullptr};
  bool haveInternalBoundary{false}, haveEmfBoundary{false};

  Fluid(Input &input, DataBlock *d) : data(d) {
    boundary = std::make_unique<Boundary<Phys>>(this);
    const std::string s = input.Get<std::string>("Hydro", "solver", 0);
    if (s == "tvdlf") mySolver = TVDLF;
    else if (s == "hll") mySolver = HLL;
#if MHD == YES
    else if (s == "hlld") mySolver = HLLD;
#else
    else if (s == "hllc") mySolver = HLLC;
#endif
    else if (s == "roe") mySolver = ROE;
    else IDEFIX_ERROR(std::string("Unknown ") + (MHD == YES ? "MHD" : "HD") + " solver type " + s);
#if MHD == YES
    if (input.CheckEntry("Hydro", "emf") >= 0) {
      const std::string e = input.Get<std::string>("Hydro", "emf", 0);
      if (e == "arithmetic") emfScheme = IDFX_EMF_ARITHMETIC;
      else if (e == "uct0") emfScheme = IDFX_EMF_UCT0;
      else if (e == "uct_contact") emfScheme = IDFX_EMF_UCT_CONTACT;
      else if (e == "uct_hll") emfScheme = IDFX_EMF_UCT_HLL;
      else if (e == "uct_hlld") emfScheme = IDFX_EMF_UCT_HLLD;
      else IDEFIX_ERROR("Unknown EMF averaging scheme " + e);
      // constrainedTransport.hpp:254-259
      if (emfScheme == IDFX_EMF_UCT_HLLD && mySolver != HLLD)
        IDEFIX_ERROR("uct_hlld averaging is only compatible with the HLLD Riemann solver");
    }
#endif
#ifdef ISOTHERMAL
    if (input.Get<std::string>("Hydro", "csiso", 0) != "constant")
      IDEFIX_ERROR("csiso admits only constant entries in this build");
    csiso = input.Get<real>("Hydro", "csiso", 1);
#else
    gamma = input.GetOrSet<real>("Hydro", "gamma", 0, 5.0 / 3.0);
#endif
    if (input.CheckEntry("Hydro", "tracer") >= 0) {   // fluid.hpp:315-323
      haveTracer = true;
      nTracer = input.Get<int>("Hydro", "tracer", 0);
      if (nTracer < 1) IDEFIX_ERROR("The number of passive tracers should be >= 1");
    }
    if (input.CheckEntry("Hydro", "shockFlattening") >= 0)   // riemannSolver.hpp:154-159
      shockFlattening = input.Get<real>("Hydro", "shockFlattening", 0);
    const char *unsupported[] = {"viscosity", "resistivity", "ambipolar", "hall",
                                 "rotation", "gravity", "fargo", "TDiffusion", "bragViscosity", "bragTDiffusion",
                                 "drag"};
    for (const char *opt : unsupported)
      if (input.CheckEntry("Hydro", opt) >= 0)
        IDEFIX_ERROR(std::string("[Hydro] ") + opt + " is outside the scope of this build (explicit Godunov path only)");
    VcName.push_back("RHO");
    for (int c = 0; c < COMPONENTS; c++) VcName.push_back("VX" + std::to_string(c + 1));
#if HAVE_ENERGY
    VcName.push_back("PRS");
#endif
#if MHD == YES
    for (int c = 0; c < COMPONENTS; c++) VcName.push_back("BX" + std::to_string(c + 1));
    for (int c = 0; c < DIMENSIONS; c++) VsName.push_back("BX" + std::to_string(c + 1) + "s");
#ifdef EVOLVE_VECTOR_POTENTIAL
#if DIMENSIONS < 3
    VeName.push_back("AX3e");   // fluid.hpp:667-690
#else
    for (int c = 0; c < 3; c++) VeName.push_back("AX" + std::to_string(c + 1) + "e");
#endif
#endif
#endif
    for (int n = 0; n < nTracer; n++) VcName.push_back("TR" + std::to_string(n));   // fluid.hpp:634
  }

  real GetGamma() const { return gamma; }

  // ---- enrolment (enroll.hpp:13-113) ------------------------------------------------------------
  void EnrollUserDefBoundary(UserDefBoundaryFunc<Phys> f);
  void EnrollUserDefBoundary(UserDefBoundaryFuncOld f);
  void EnrollInternalBoundary(InternalBoundaryFunc<Phys> f);      // enroll.hpp:33-50
  void EnrollInternalBoundary(InternalBoundaryFuncOld f);
  void EnrollEmfBoundary(EmfBoundaryFunc f);                      // enroll.hpp:71-78
  template <typename F>
  void EnrollUserSourceTerm(F) { IDEFIX_ERROR("user source terms are outside the scope of this build"); }
  template <typename F>
  void EnrollIsoSoundSpeed(F) { IDEFIX_ERROR("user-defined sound speeds are outside the scope of this build"); }

  // ---- forwards to the C ABI ------------------------------------------------------------------------
  void ConvertConsToPrim();
  void ConvertPrimToCons();
  void ResetStage();
  void EvolveStage(const real t, const real dt);
  real CheckDivB();
  int CheckNan();
  void ShowConfig();
};

class DataBlock {
 public:
  std::array<IdefixArray1D<real>, 3> x, xr, xl, dx;   // device copies of the block coordinates
  std::array<real, 3> xbeg{}, xend{};
  std::array<int, 3> np_tot{}, np_int{}, nghost{}, beg{}, end{}, gbeg{}, gend{};
  std::array<BoundaryType, 3> lbound{}, rbound{};
  real t{0.0}, dt{0.0};
  Grid *mygrid{nullptr};
  std::unique_ptr<Hydro> hydro;
  bool haveAxis{false};          // an X2 boundary lies on the polar axis (dataBlock.cpp:43-50)
  bool haveUserStepFirst{false}, haveUserStepLast{false};
  StepFunc userStepFirst{nullptr}, userStepLast{nullptr};
  IdfxBlock *handle{nullptr};   // the C-ABI block behind this DataBlock
  cudaStream_t stream{nullptr};  // the stream the library's kernels run on
  // time-integrator options the device-side dt control needs
  int nstages{1};
  real cfl{0.0}, cflMaxVar{1.1}, firstDt{1e-10}, fixedDt{-1.0};

  DataBlock(Grid &grid, Input &input) : mygrid(&grid) {
    for (int dir = 0; dir < 3; dir++) {
      np_int[dir] = grid.np_int[dir];
      nghost[dir] = grid.nghost[dir];
      np_tot[dir] = grid.np_tot[dir];
      beg[dir] = nghost[dir];
      end[dir] = nghost[dir] + np_int[dir];
      gbeg[dir] = nghost[dir];
      gend[dir] = nghost[dir] + np_int[dir];
      if (grid.lbound[dir] == axis || grid.rbound[dir] == axis) haveAxis = true;
      lbound[dir] = grid.lbound[dir];
      rbound[dir] = grid.rbound[dir];
      xbeg[dir] = grid.xbeg[dir];
      xend[dir] = grid.xend[dir];
    }
    hydro = std::make_unique<Hydro>(input, this);
    nstages = input.Get<int>("TimeIntegrator", "nstages", 0);
    if (input.CheckEntry("TimeIntegrator", "fixed_dt") > 0) {
      fixedDt = input.Get<real>("TimeIntegrator", "fixed_dt", 0);
    } else {
      cfl = input.Get<real>("TimeIntegrator", "CFL", 0);
      cflMaxVar = input.GetOrSet<real>("TimeIntegrator", "CFL_max_var", 0, 1.1);
      firstDt = input.GetOrSet<real>("TimeIntegrator", "first_dt", 0, 1.0e-10);
    }

    IdfxBlockDesc d{};
    d.ndim = DIMENSIONS;
    d.ncomp = COMPONENTS;
    d.mhd = (MHD == YES);
    d.isothermal = !HAVE_ENERGY;
    d.geometry = GEOMETRY;
    d.order = ORDER;
    d.solver = hydro->mySolver;
    d.emf = hydro->emfScheme;
    d.nstages = nstages;
    for (int dir = 0; dir < 3; dir++) {
      d.np_int[dir] = np_int[dir];
      d.nghost[dir] = nghost[dir];
      d.lbound[dir] = lbound[dir];   // BoundaryType values == enum IdfxBoundary (idefix.hpp:220)
      d.rbound[dir] = rbound[dir];
      d.x[dir] = grid.x[dir].data();
      d.xl[dir] = grid.xl[dir].data();
      d.xr[dir] = grid.xr[dir].data();
      d.dx[dir] = grid.dx[dir].data();
    }
    d.regular_grid = grid.isRegularCartesian;
    d.device = input.device;
    d.gamma = hydro->gamma;
    d.csiso = hydro->csiso;
    d.cfl = cfl;
    d.cfl_max_var = cflMaxVar;
    d.first_dt = firstDt;
    d.fixed_dt = fixedDt;
    d.ntracer = hydro->nTracer;
    d.shock_flattening = hydro->shockFlattening;
#ifdef EVOLVE_VECTOR_POTENTIAL
#if MHD == YES && DIMENSIONS >= 2
    d.evolve_vector_potential = 1;
#else
    IDEFIX_ERROR("EVOLVE_VECTOR_POTENTIAL is not compatible with 1D MHD");   // fluid.hpp:549-551
#endif
#endif
    IDFX_CUDA(cudaSetDevice(input.device));
    IDFX_ABI(idfx_block_create(&d, &handle));
    // user kernels (idefix_for) run on the legacy default stream.  The library gets a BLOCKING stream of its
    // own: blocking streams synchronise implicitly with the legacy default stream, so everything stays ordered
    // without explicit fences (like the reference's single execution space), and -- unlike the legacy stream --
    // it can be captured, which lets idfx_cycles replay the steady-state cycle as a CUDA graph
    IDFX_CUDA(cudaStreamCreate(&stream));
    IDFX_ABI(idfx_set_stream(handle, stream));
    dt = firstDt;
    if (fixedDt > 0) dt = fixedDt;

    // views of the block's state arrays
    double *p = nullptr;
    int n[4];
    IDFX_ABI(idfx_field_info(handle, IDFX_FIELD_VC, &p, n));
    hydro->Vc = IdefixArray4D<real>(p, n[0], n[1], n[2], n[3]);
    IDFX_ABI(idfx_field_info(handle, IDFX_FIELD_UC, &p, n));
    hydro->Uc = IdefixArray4D<real>(p, n[0], n[1], n[2], n[3]);
    IDFX_ABI(idfx_field_info(handle, IDFX_FIELD_INVDT, &p, n));
    hydro->InvDt = IdefixArray3D<real>(p, n[1], n[2], n[3]);
#if MHD == YES
    IDFX_ABI(idfx_field_info(handle, IDFX_FIELD_VS, &p, n));
    hydro->Vs = IdefixArray4D<real>(p, n[0], n[1], n[2], n[3]);
#ifdef EVOLVE_VECTOR_POTENTIAL
    IDFX_ABI(idfx_field_info(handle, IDFX_FIELD_VE, &p, n));
    hydro->Ve = IdefixArray4D<real>(p, n[0], n[1], n[2], n[3]);
#endif
#if DIMENSIONS >= 2
    hydro->emf = std::make_unique<ConstrainedTransport>();
    IDFX_ABI(idfx_field_info(handle, IDFX_FIELD_EZ, &p, n));
    hydro->emf->ez = IdefixArray3D<real>(p, n[1], n[2], n[3]);
#if DIMENSIONS == 3
    IDFX_ABI(idfx_field_info(handle, IDFX_FIELD_EX, &p, n));
    hydro->emf->ex = IdefixArray3D<real>(p, n[1], n[2], n[3]);
    IDFX_ABI(idfx_field_info(handle, IDFX_FIELD_EY, &p, n));
    hydro->emf->ey = IdefixArray3D<real>(p, n[1], n[2], n[3]);
#endif
#endif
#endif
    for (int dir = 0; dir < 3; dir++) {
      x[dir] = IdefixArray1D<real>("DataBlock_x", np_tot[dir]);
      xl[dir] = IdefixArray1D<real>("DataBlock_xl", np_tot[dir]);
      xr[dir] = IdefixArray1D<real>("DataBlock_xr", np_tot[dir]);
      dx[dir] = IdefixArray1D<real>("DataBlock_dx", np_tot[dir]);
      Kokkos::deep_copy(x[dir], grid.x[dir]);
      Kokkos::deep_copy(xl[dir], grid.xl[dir]);
      Kokkos::deep_copy(xr[dir], grid.xr[dir]);
      Kokkos::deep_copy(dx[dir], grid.dx[dir]);
    }
  }
  ~DataBlock() {
    if (handle) idfx_block_destroy(handle);
    if (stream) cudaStreamDestroy(stream);
  }
  DataBlock(const DataBlock &) = delete;
  DataBlock &operator=(const DataBlock &) = delete;

  // ---- forwards (dataBlock.hpp:120-160) ----------------------------------------------------------
  void PushTime() { IDFX_ABI(idfx_set_time(handle, t, dt)); }
  void PullTime() { IDFX_ABI(idfx_get_time(handle, &t, &dt)); }
  void SetBoundaries() { IDFX_ABI(idfx_set_boundaries(handle, t)); }
  // Vs = curl Ve when the vector potential is the evolved variable, a no-op otherwise (dataBlock.cpp:366-373)
  void DeriveVectorPotential() { IDFX_ABI(idfx_derive_vector_potential(handle)); }
  void ResetStage() { hydro->ResetStage(); }
  void ConsToPrim() { hydro->ConvertConsToPrim(); }
  void PrimToCons() { hydro->ConvertPrimToCons(); }
  void EvolveStage() { hydro->EvolveStage(t, dt); }
  void Validate() {
    if (CheckNan() > 0) IDEFIX_ERROR("DataBlock::Validate: the initial state contains NaNs");
  }
  real ComputeTimestep() {
    real v;
    IDFX_ABI(idfx_compute_timestep(handle, &v));
    return v;
  }
  int CheckNan() { return hydro->CheckNan(); }
  void EnrollUserStepFirst(StepFunc f) {
    userStepFirst = f;
    haveUserStepFirst = true;
  }
  void EnrollUserStepLast(StepFunc f) {
    userStepLast = f;
    haveUserStepLast = true;
  }
  void LaunchUserStepFirst() {
    if (haveUserStepFirst) userStepFirst(*this, t, dt);
  }
  void LaunchUserStepLast() {
    if (haveUserStepLast) userStepLast(*this, t, dt);
  }
  void ShowConfig() {
    hydro->ShowConfig();
    if (haveAxis) {   // axis.cpp:12-23
      idfx::cout << "Axis: Axis regularisation ENABLED." << std::endl;
      if (std::fabs(mygrid->xend[KDIR] - mygrid->xbeg[KDIR] - 2.0 * M_PI) < 1e-10)
        idfx::cout << "Axis: Full 2pi regularisation around the axis." << std::endl;
      else
        idfx::cout << "Axis: Fractional (2pi/N) regularisation around the axis." << std::endl;
    }
  }
};

// ---- Fluid / Boundary members that need the complete DataBlock -----------------------------------------
template <typename Phys> void Fluid<Phys>::ConvertConsToPrim() { IDFX_ABI(idfx_cons_to_prim(data->handle)); }
template <typename Phys> void Fluid<Phys>::ConvertPrimToCons() { IDFX_ABI(idfx_prim_to_cons(data->handle)); }
template <typename Phys> void Fluid<Phys>::ResetStage() { IDFX_ABI(idfx_reset_stage(data->handle)); }
template <typename Phys> void Fluid<Phys>::EvolveStage(const real t, const real dt) {
  IDFX_ABI(idfx_evolve_stage(data->handle, t, dt));
}
template <typename Phys> real Fluid<Phys>::CheckDivB() {
  real v = 0.0;
  IDFX_ABI(idfx_check_divb(data->handle, &v));
  return v;
}
template <typename Phys> int Fluid<Phys>::CheckNan() {
  int n = 0;
  IDFX_ABI(idfx_check_nan(data->handle, &n));
  return n;
}
template <typename Phys> void Boundary<Phys>::SetBoundaries(real t) {
  IDFX_ABI(idfx_set_boundaries(fluid->data->handle, t));
}
template <typename Phys> void Boundary<Phys>::EnforceBoundaryDir(real t, int dir) {
  IDFX_ABI(idfx_enforce_boundary_dir(fluid->data->handle, t, dir));
}
template <typename Phys> void Boundary<Phys>::ReconstructVcField(IdefixArray4D<real> &V) {
  IDFX_ABI(idfx_reconstruct_vc_field(fluid->data->handle, V.data() == fluid->Uc.data() ? IDFX_FIELD_UC : IDFX_FIELD_VC));
}

namespace idfxhost {
template <typename Phys>
inline void userBoundaryTrampoline(IdfxBlock *, int dir, int side, double t, void *user) {
  auto *f = static_cast<Fluid<Phys> *>(user);
  if (f->userDefBoundaryFunc) f->userDefBoundaryFunc(f, dir, side == 0 ? left : right, t);
  else if (f->userDefBoundaryFuncOld) f->userDefBoundaryFuncOld(*f->data, dir, side == 0 ? left : right, t);
}
}  // namespace idfxhost

namespace idfxhost {
template <typename Phys>
inline void internalBoundaryTrampoline(IdfxBlock *, double t, void *user) {
  auto *f = static_cast<Fluid<Phys> *>(user);
  if (f->internalBoundaryFunc) f->internalBoundaryFunc(f, t);
  else if (f->internalBoundaryFuncOld) f->internalBoundaryFuncOld(*f->data, t);
}
template <typename Phys>
inline void emfBoundaryTrampoline(IdfxBlock *, double t, void *user) {
  auto *f = static_cast<Fluid<Phys> *>(user);
  if (f->emfBoundaryFunc) f->emfBoundaryFunc(*f->data, t);
}
}  // namespace idfxhost

template <typename Phys> void Fluid<Phys>::EnrollInternalBoundary(InternalBoundaryFunc<Phys> f) {
  haveInternalBoundary = true;
  internalBoundaryFunc = f;
  IDFX_ABI(idfx_enroll_stage_hook(data->handle, IDFX_HOOK_INTERNAL_BOUNDARY, &idfxhost::internalBoundaryTrampoline<Phys>, this));
}
template <typename Phys> void Fluid<Phys>::EnrollInternalBoundary(InternalBoundaryFuncOld f) {
  haveInternalBoundary = true;
  internalBoundaryFuncOld = f;
  IDFX_ABI(idfx_enroll_stage_hook(data->handle, IDFX_HOOK_INTERNAL_BOUNDARY, &idfxhost::internalBoundaryTrampoline<Phys>, this));
}
template <typename Phys> void Fluid<Phys>::EnrollEmfBoundary(EmfBoundaryFunc f) {
  haveEmfBoundary = true;
  emfBoundaryFunc = f;
  IDFX_ABI(idfx_enroll_stage_hook(data->handle, IDFX_HOOK_EMF_BOUNDARY, &idfxhost::emfBoundaryTrampoline<Phys>, this));
}

template <typename Phys> void Fluid<Phys>::EnrollUserDefBoundary(UserDefBoundaryFunc<Phys> f) {
  userDefBoundaryFunc = f;
  IDFX_ABI(idfx_enroll_userdef_boundary(data->handle, &idfxhost::userBoundaryTrampoline<Phys>, this));
}
template <typename Phys> void Fluid<Phys>::EnrollUserDefBoundary(UserDefBoundaryFuncOld f) {
  userDefBoundaryFuncOld = f;
  IDFX_ABI(idfx_enroll_userdef_boundary(data->handle, &idfxhost::userBoundaryTrampoline<Phys>, this));
}

template <typename Phys> void Fluid<Phys>::ShowConfig() {
  // src/fluid/showConfig.hpp:20-120 (the lines the reference's test harness greps are kept verbatim)
  idfx::cout << "Fluid: solving " << (MHD == YES ? "MHD" : "HD") << " equations." << std::endl;
  idfx::cout << "Fluid: Reconstruction: ";
  if (ORDER == 1) idfx::cout << "1st order (donor cell)" << std::endl;
  if (ORDER == 2) idfx::cout << "2nd order (PLM Van Leer)" << std::endl;
  if (ORDER == 3) idfx::cout << "3rd order (LimO3)" << std::endl;
  if (ORDER == 4) idfx::cout << "4th order (PPM)" << std::endl;
#if HAVE_ENERGY
  idfx::cout << "Fluid: EOS: ideal with gamma=" << gamma << std::endl;
#else
  idfx::cout << "Fluid: EOS: isothermal with cs=" << csiso << "." << std::endl;
#endif
  const char *names[] = {"tvdlf (Lax-Friedrich)", "HLL", "HLLC", "HLLD", "ROE"};
  idfx::cout << "Fluid: Riemann solver: " << names[mySolver] << "." << std::endl;
#if MHD == YES
  const char *emfs[] = {"Arithmetic", "UCT0", "UCT_CONTACT", "UCT_HLL", "UCT_HLLD"};
  idfx::cout << "Fluid: EMF averaging: " << emfs[emfScheme] << "." << std::endl;
#endif
}

#endif  // IDEFIX_B200_HOST_DATABLOCK_HPP_
