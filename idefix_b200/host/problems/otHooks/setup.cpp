// The 2-D Orszag-Tang vortex of test/MHD/OrszagTang with two user hooks enrolled, written against the host layer
// exactly as a problem for the reference would be (src/fluid/enroll.hpp:33,71):
//   * an internal boundary function that rewrites nothing when called with mode 0 (identity: the run must then equal
//     the un-hooked one bit for bit) and floors the density when `[Setup] rhofloor` is set;
//   * an EMF boundary function that multiplies the corner EMF ez by `[Setup] emfscale` (1 = identity) through
//     data.hydro->emf->ez, between CalcCornerEMF and EvolveMagField.
#include "idefix.hpp"
#include "setup.hpp"

static real rhoFloor = 0.0;
static real emfScale = 1.0;
static int hookCalls[2] = {0, 0};

void InternalBoundary(Hydro *hydro, const real t) {
  hookCalls[0]++;
  if (rhoFloor <= 0.0) return;
  IdefixArray4D<real> Vc = hydro->Vc;
  DataBlock &data = *hydro->data;
  const real floor = rhoFloor;
  idefix_for("InternalBoundary", 0, data.np_tot[KDIR], 0, data.np_tot[JDIR], 0, data.np_tot[IDIR],
             KOKKOS_LAMBDA(int k, int j, int i) {
               if (Vc(RHO, k, j, i) < floor) Vc(RHO, k, j, i) = floor;
             });
}

void EmfBoundary(DataBlock &data, const real t) {
  hookCalls[1]++;
  if (emfScale == 1.0) return;
  IdefixArray3D<real> ez = data.hydro->emf->ez;
  const real scale = emfScale;
  idefix_for("EmfBoundary", 0, data.np_tot[KDIR], 0, data.np_tot[JDIR], 0, data.np_tot[IDIR],
             KOKKOS_LAMBDA(int k, int j, int i) { ez(k, j, i) *= scale; });
}

Setup::Setup(Input &input, Grid &grid, DataBlock &data, Output &output) {
  rhoFloor = input.GetOrSet<real>("Setup", "rhofloor", 0, 0.0);
  emfScale = input.GetOrSet<real>("Setup", "emfscale", 0, 1.0);
  data.hydro->EnrollInternalBoundary(&InternalBoundary);
  data.hydro->EnrollEmfBoundary(&EmfBoundary);
}

Setup::~Setup() {
  idfx::cout << "Setup: hooks called: internal boundary " << hookCalls[0] << ", emf boundary " << hookCalls[1] << std::endl;
}

void Setup::InitFlow(DataBlock &data) {
  DataBlockHost host(data);
  const real twoPi = 2.0 * M_PI;
  const real B0 = 1.0 / sqrt(4.0 * M_PI);
  const real rho0 = 25.0 / (36.0 * M_PI), prs0 = 5.0 / (12.0 * M_PI);
  for (int k = 0; k < host.np_tot[KDIR]; k++) {
    for (int j = 0; j < host.np_tot[JDIR]; j++) {
      const real sy = sin(twoPi * host.x[JDIR](j));
      for (int i = 0; i < host.np_tot[IDIR]; i++) {
        const real xc = host.x[IDIR](i);
        host.Vc(RHO, k, j, i) = rho0;
        host.Vc(PRS, k, j, i) = prs0;
        host.Vc(VX1, k, j, i) = -sy;
        host.Vc(VX2, k, j, i) = sin(twoPi * xc);
        host.Vs(BX1s, k, j, i) = -B0 * sy;
        host.Vs(BX2s, k, j, i) = B0 * sin(4.0 * M_PI * xc);
      }
    }
  }
  host.SyncToDevice();
}
