// An isothermal, magnetised shear flow in 2.5 dimensions: periodic along x, open along y, with an out-of-plane
// velocity and field so that the third vector component (evolved through the cell-centred conservative
// variables, not by constrained transport) is exercised together with the isothermal Riemann solvers and
// ReconstructNormalField on the outflow sides.
#include "idefix.hpp"
#include "setup.hpp"

Setup::Setup(Input &input, Grid &grid, DataBlock &data, Output &output) {}

void Setup::InitFlow(DataBlock &data) {
  DataBlockHost d(data);
  const real twoPi = 2.0 * M_PI;
  const real B0 = 0.4;
  IdefixHostArray4D<real> A("A", 3, d.np_tot[KDIR] + KOFFSET, d.np_tot[JDIR] + 1, d.np_tot[IDIR] + 1);
  for (int k = 0; k < d.np_tot[KDIR]; k++) {
    for (int j = 0; j < d.np_tot[JDIR]; j++) {
      for (int i = 0; i < d.np_tot[IDIR]; i++) {
        const real x = d.x[IDIR](i), y = d.x[JDIR](j);
        const real xl = d.xl[IDIR](i), yl = d.xl[JDIR](j);
        d.Vc(RHO, k, j, i) = 1.0 + 0.3 * sin(twoPi * x) * exp(-8.0 * (y - 0.5) * (y - 0.5));
        d.Vc(VX1, k, j, i) = 0.5 * tanh(10.0 * (y - 0.5));
        d.Vc(VX2, k, j, i) = 0.1 * sin(twoPi * x);
        d.Vc(VX3, k, j, i) = 0.2 * cos(twoPi * x) * sin(M_PI * y);
        d.Vc(BX3, k, j, i) = 0.3 * B0 * sin(twoPi * x + 1.0);
        // in-plane field from A_z given on the cell corners (i-1/2, j-1/2)
        A(KDIR, k, j, i) = B0 * (yl + 0.05 * cos(twoPi * xl) * sin(twoPi * yl));
      }
    }
  }
  d.MakeVsFromAmag(A);
  d.SyncToDevice();
}
