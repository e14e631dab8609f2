// 3-D Orszag-Tang vortex on [0,1]^3.  The magnetic field derives from a vector potential given on
// the cell edges, so div B vanishes to round-off from the first step (DataBlockHost::MakeVsFromAmag).
#include "idefix.hpp"
#include "setup.hpp"

Setup::Setup(Input &input, Grid &grid, DataBlock &data, Output &output) {}

void Setup::InitFlow(DataBlock &data) {
  DataBlockHost host(data);
  const int n1 = host.np_tot[IDIR], n2 = host.np_tot[JDIR], n3 = host.np_tot[KDIR];
  IdefixHostArray4D<real> A("VectorPotential", 3, n3 + 1, n2 + 1, n1 + 1);
  const real twoPi = 2.0 * M_PI;
  const real B0 = 1.0 / sqrt(4.0 * M_PI);
  const real amp = B0 / twoPi;
  for (int k = 0; k < n3; k++) {
    const real cz = cos(twoPi * host.x[KDIR](k));
    for (int j = 0; j < n2; j++) {
      const real sy = sin(twoPi * host.x[JDIR](j));
      const real cyl = cos(twoPi * host.xl[JDIR](j));
      for (int i = 0; i < n1; i++) {
        const real xc = host.x[IDIR](i), xf = host.xl[IDIR](i);
        host.Vc(RHO, k, j, i) = 25.0 / (36.0 * M_PI);
        host.Vc(PRS, k, j, i) = 5.0 / (12.0 * M_PI);
        host.Vc(VX1, k, j, i) = -sy;
        host.Vc(VX2, k, j, i) = sin(twoPi * xc) + cz;
        host.Vc(VX3, k, j, i) = cos(twoPi * xc);
        if (data.hydro->haveTracer) {   // passive tracers: two step functions
          host.Vc(TRG, k, j, i) = xc > 0.5 ? 1.0 : 0.0;
          if (data.hydro->nTracer > 1) host.Vc(TRG + 1, k, j, i) = host.x[KDIR](k) > 0.5 ? 1.0 : 0.0;
        }
        A(IDIR, k, j, i) = amp * cyl;
        A(JDIR, k, j, i) = amp * sin(twoPi * xf);
        A(KDIR, k, j, i) = amp * (cyl + cos(4.0 * M_PI * xf) / 2.0);
#ifdef EVOLVE_VECTOR_POTENTIAL
        // the potential itself is the evolved variable: hand it over, the code derives the face field
        host.Ve(AX1e, k, j, i) = A(IDIR, k, j, i);
        host.Ve(AX2e, k, j, i) = A(JDIR, k, j, i);
        host.Ve(AX3e, k, j, i) = A(KDIR, k, j, i);
#endif
      }
    }
  }
#ifndef EVOLVE_VECTOR_POTENTIAL
  host.MakeVsFromAmag(A);
#endif
  host.SyncToDevice();
}
