// A tour of the API a problem file can use:
//  * Input::Get / GetOrSet of a user block,
//  * a user-owned device array filled by idefix_for (KOKKOS_LAMBDA), mirrored with
//    Kokkos::create_mirror_view / deep_copy,
//  * idefix_reduce with Kokkos::Sum / Kokkos::Max on the block's own device arrays,
//  * an analysis function (Output::EnrollAnalysis), user-defined VTK variables
//    (Output::EnrollUserDefVariables) and a step hook (DataBlock::EnrollUserStepLast).
#include "idefix.hpp"
#include "setup.hpp"

namespace {
real amplitude = 0.0;
int stepCalls = 0;
IdefixArray3D<real> blob;   // user-owned device array: the initial density perturbation

void CountSteps(DataBlock &data, const real t, const real dt) { stepCalls++; }

void Analysis(DataBlock &data) {
  IdefixArray4D<real> Vc = data.hydro->Vc;
  IdefixArray1D<real> dx1 = data.dx[IDIR], dx2 = data.dx[JDIR];
  real mass = 0.0, vmax = 0.0;
  idefix_reduce("Mass", data.beg[KDIR], data.end[KDIR], data.beg[JDIR], data.end[JDIR], data.beg[IDIR], data.end[IDIR],
                KOKKOS_LAMBDA(int k, int j, int i, real &acc) { acc += Vc(RHO, k, j, i) * dx1(i) * dx2(j); },
                Kokkos::Sum<real>(mass));
  idefix_reduce("Vmax", data.beg[KDIR], data.end[KDIR], data.beg[JDIR], data.end[JDIR], data.beg[IDIR], data.end[IDIR],
                KOKKOS_LAMBDA(int k, int j, int i, real &acc) { acc = FMAX(acc, FABS(Vc(VX1, k, j, i))); },
                Kokkos::Max<real>(vmax));
  // the same mass on the host mirror
  DataBlockHost d(data);
  d.SyncFromDevice();
  real massHost = 0.0;
  for (int j = d.beg[JDIR]; j < d.end[JDIR]; j++)
    for (int i = d.beg[IDIR]; i < d.end[IDIR]; i++) massHost += d.Vc(RHO, 0, j, i) * d.dx[IDIR](i) * d.dx[JDIR](j);
  std::FILE *f = std::fopen("analysis.txt", "a");
  std::fprintf(f, "%.17g %.17g %.17g %.17g %d\n", data.t, mass, massHost, vmax, stepCalls);
  std::fclose(f);
}

void UserVars(DataBlock &data, UserDefVariablesContainer &vars) {
  DataBlockHost d(data);
  d.SyncFromDevice();
  IdefixHostArray3D<real> mach = vars["MACH"], blobOut = vars["BLOB"];
  IdefixHostArray3D<real> blobHost = Kokkos::create_mirror_view(blob);
  Kokkos::deep_copy(blobHost, blob);
  const real gamma = data.hydro->GetGamma();
  for (int k = 0; k < d.np_tot[KDIR]; k++)
    for (int j = 0; j < d.np_tot[JDIR]; j++)
      for (int i = 0; i < d.np_tot[IDIR]; i++) {
        const real cs = sqrt(gamma * d.Vc(PRS, k, j, i) / d.Vc(RHO, k, j, i));
        mach(k, j, i) = sqrt(d.Vc(VX1, k, j, i) * d.Vc(VX1, k, j, i) + d.Vc(VX2, k, j, i) * d.Vc(VX2, k, j, i)) / cs;
        blobOut(k, j, i) = blobHost(k, j, i);
      }
}
void FillBlob(DataBlock &data, const real width) {
  blob = IdefixArray3D<real>("blob", data.np_tot[KDIR], data.np_tot[JDIR], data.np_tot[IDIR]);
  IdefixArray3D<real> b = blob;
  IdefixArray1D<real> x1 = data.x[IDIR], x2 = data.x[JDIR];
  const real amp = amplitude;
  idefix_for("FillBlob", 0, data.np_tot[KDIR], 0, data.np_tot[JDIR], 0, data.np_tot[IDIR],
             KOKKOS_LAMBDA(int k, int j, int i) {
               const real r2 = (x1(i) - 0.5) * (x1(i) - 0.5) + (x2(j) - 0.5) * (x2(j) - 0.5);
               b(k, j, i) = amp * exp(-r2 / (width * width));
             });
}
}  // namespace

Setup::Setup(Input &input, Grid &grid, DataBlock &data, Output &output) {
  amplitude = input.Get<real>("Setup", "amplitude", 0);
  const real width = input.GetOrSet<real>("Setup", "width", 0, 0.1);
  output.EnrollAnalysis(&Analysis);
  output.EnrollUserDefVariables(&UserVars);
  data.EnrollUserStepLast(&CountSteps);
  FillBlob(data, width);   // (nvcc, like Kokkos-CUDA, does not allow device lambdas inside a constructor)
}

void Setup::InitFlow(DataBlock &data) {
  // initial condition written ON THE DEVICE from the user array, then pulled into the host mirror
  IdefixArray4D<real> Vc = data.hydro->Vc;
  IdefixArray3D<real> b = blob;
  idefix_for("InitFlow", 0, data.np_tot[KDIR], 0, data.np_tot[JDIR], 0, data.np_tot[IDIR],
             KOKKOS_LAMBDA(int k, int j, int i) {
               Vc(RHO, k, j, i) = 1.0 + b(k, j, i);
               Vc(VX1, k, j, i) = 1.0;
               Vc(VX2, k, j, i) = 0.5;
               Vc(PRS, k, j, i) = 1.0;
             });
  Kokkos::fence();
}
