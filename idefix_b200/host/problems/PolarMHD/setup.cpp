// A magnetised rotating annulus in polar geometry (R, phi, z) with a radially stretched grid: exercises
// the polar angular-momentum flux, the B_phi flux factor, the radial geometric source term, the
// irregular-grid reconstruction weights and ReconstructNormalField on reflective / outflow sides.
#include "idefix.hpp"
#include "setup.hpp"

Setup::Setup(Input &input, Grid &grid, DataBlock &data, Output &output) {}

void Setup::InitFlow(DataBlock &data) {
  DataBlockHost d(data);
  const int n1 = d.np_tot[IDIR], n2 = d.np_tot[JDIR], n3 = d.np_tot[KDIR];
  IdefixHostArray4D<real> A("A", 3, n3 + 1, n2 + 1, n1 + 1);
  const real B0 = 0.25;
  for (int k = 0; k < n3; k++) {
    for (int j = 0; j < n2; j++) {
      for (int i = 0; i < n1; i++) {
        const real R = d.x[IDIR](i), ph = d.x[JDIR](j), z = d.x[KDIR](k);
        const real Rl = d.xl[IDIR](i), phl = d.xl[JDIR](j), zl = d.xl[KDIR](k);
        d.Vc(RHO, k, j, i) = (1.0 + 0.3 * cos(ph) * exp(-4.0 * z * z)) / R;
        d.Vc(PRS, k, j, i) = 0.5 / (R * R) * (1.0 + 0.1 * sin(2.0 * ph));
        d.Vc(VX1, k, j, i) = 0.03 * sin(ph) * cos(3.0 * z);
        d.Vc(VX2, k, j, i) = 0.4 / sqrt(R);
        d.Vc(VX3, k, j, i) = 0.02 * cos(2.0 * ph);
        // A_R on (i, j-1/2, k-1/2), A_phi on (i-1/2, j, k-1/2), A_z on (i-1/2, j-1/2, k)
        A(IDIR, k, j, i) = 0.05 * B0 * sin(phl) * zl;
        A(JDIR, k, j, i) = 0.5 * B0 * Rl * (1.0 + 0.2 * zl);
        A(KDIR, k, j, i) = 0.1 * B0 * Rl * cos(phl);
      }
    }
  }
  d.MakeVsFromAmag(A);
  d.SyncToDevice();
}
