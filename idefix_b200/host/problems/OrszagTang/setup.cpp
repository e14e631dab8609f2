// 2-D Orszag-Tang vortex on [0,1]^2: uniform density and pressure, a single velocity vortex and a
// double-periodic magnetic field given directly on the cell faces.
#include "idefix.hpp"
#include "setup.hpp"

Setup::Setup(Input &input, Grid &grid, DataBlock &data, Output &output) {}

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
#ifdef EVOLVE_VECTOR_POTENTIAL
        // A_z on the cell corners: the potential is the evolved variable, the code derives the face field from it
        const real xf = host.xl[IDIR](i), yf = host.xl[JDIR](j);
        host.Ve(AX3e, k, j, i) = B0 / (2.0 * M_PI) * (cos(2.0 * M_PI * yf) + cos(4.0 * M_PI * xf) / 2.0);
#else
        host.Vs(BX1s, k, j, i) = -B0 * sy;
        host.Vs(BX2s, k, j, i) = B0 * sin(4.0 * M_PI * xc);
#endif
      }
    }
  }
  host.SyncToDevice();
}
