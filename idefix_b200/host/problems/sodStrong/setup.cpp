// Strong 1-D blast tube: the pressure jump (1000 : 0.01) drives interfaces through the strong-shock switch of the
// HD Roe solver, which in one dimension keeps the flux of the previous call (roeHD.hpp:312-328).
// Written against the host layer exactly as a problem for the reference would be.
#include "idefix.hpp"
#include "setup.hpp"

Setup::Setup(Input &input, Grid &grid, DataBlock &data, Output &output) {}

void Setup::InitFlow(DataBlock &data) {
  DataBlockHost host(data);
  for (int i = 0; i < host.np_tot[IDIR]; i++) {
    const bool right = host.x[IDIR](i) > HALF_F;
    host.Vc(RHO, 0, 0, i) = ONE_F;
    host.Vc(VX1, 0, 0, i) = ZERO_F;
    host.Vc(PRS, 0, 0, i) = right ? 0.01 : 1000.0;
  }
  host.SyncToDevice();
}
