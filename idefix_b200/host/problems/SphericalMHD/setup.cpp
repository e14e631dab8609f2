// A magnetised, differentially rotating shell between two cones (no polar axis in the domain):
// non-uniform density and pressure, a non-axisymmetric velocity perturbation and a field that derives
// from a vector potential with all three components, so that every curvilinear term of the update
// (angular-momentum fluxes, field-flux factors, geometric sources, the spherical curl of the EMFs) is
// exercised.  Compiles unchanged against the reference and against the B200 host layer.
#include "idefix.hpp"
#include "setup.hpp"

Setup::Setup(Input &input, Grid &grid, DataBlock &data, Output &output) {}

void Setup::InitFlow(DataBlock &data) {
  DataBlockHost d(data);
  const int n1 = d.np_tot[IDIR], n2 = d.np_tot[JDIR], n3 = d.np_tot[KDIR];
  IdefixHostArray4D<real> A("A", 3, n3 + 1, n2 + 1, n1 + 1);
  const real B0 = 0.3;
  for (int k = 0; k < n3; k++) {
    for (int j = 0; j < n2; j++) {
      for (int i = 0; i < n1; i++) {
        const real r = d.x[IDIR](i), th = d.x[JDIR](j), ph = d.x[KDIR](k);
        const real rl = d.xl[IDIR](i), thl = d.xl[JDIR](j), phl = d.xl[KDIR](k);
        d.Vc(RHO, k, j, i) = 1.0 / r * (1.0 + 0.2 * sin(th) * cos(ph));
        d.Vc(PRS, k, j, i) = 0.6 / (r * r) * (1.0 + 0.1 * cos(2.0 * ph));
        d.Vc(VX1, k, j, i) = 0.05 * sin(3.0 * ph) * sin(2.0 * th);
        d.Vc(VX2, k, j, i) = 0.04 * cos(ph) * sin(th);
        d.Vc(VX3, k, j, i) = 0.3 / sqrt(r) * sin(th);
        // edge-centred vector potential: A_r on (i, j-1/2, k-1/2), A_theta on (i-1/2, j, k-1/2), A_phi on (i-1/2, j-1/2, k)
        A(IDIR, k, j, i) = 0.1 * B0 * cos(thl) * sin(phl) / r;
        A(JDIR, k, j, i) = 0.1 * B0 * sin(phl) * sin(th) / rl;
        A(KDIR, k, j, i) = 0.5 * B0 * rl * sin(thl) * (1.0 + 0.1 * cos(ph));
      }
    }
  }
  d.MakeVsFromAmag(A);
  d.SyncToDevice();
}
