// Rotating annulus in (R, phi) with a two-armed spiral perturbation: the hydrodynamic solvers, the polar
// angular-momentum flux and the centrifugal / pressure geometric source term in two dimensions.
#include "idefix.hpp"
#include "setup.hpp"

Setup::Setup(Input &input, Grid &grid, DataBlock &data, Output &output) {}

void Setup::InitFlow(DataBlock &data) {
  DataBlockHost d(data);
  for (int k = 0; k < d.np_tot[KDIR]; k++)
    for (int j = 0; j < d.np_tot[JDIR]; j++)
      for (int i = 0; i < d.np_tot[IDIR]; i++) {
        const real R = d.x[IDIR](i), ph = d.x[JDIR](j);
        d.Vc(RHO, k, j, i) = (1.0 + 0.3 * cos(2.0 * ph - 3.0 * R)) / R;
        d.Vc(PRS, k, j, i) = 0.3 / (R * R);
        d.Vc(VX1, k, j, i) = 0.05 * sin(2.0 * ph);
        d.Vc(VX2, k, j, i) = 0.8 / sqrt(R);
      }
  d.SyncToDevice();
}
