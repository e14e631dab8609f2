// Isothermal shock tube (ISOTHERMAL defined: no pressure variable, constant sound speed from the ini file):
// Written against the host layer exactly as a problem for the reference would be.
#include "idefix.hpp"
#include "setup.hpp"

namespace {
struct TubeState {
  real rho, prs;
};
constexpr TubeState kLeft{1.0, 1.0}, kRight{0.125, 0.1};
}  // namespace

Setup::Setup(Input &input, Grid &grid, DataBlock &data, Output &output) {}

void Setup::InitFlow(DataBlock &data) {
  DataBlockHost host(data);
  for (int i = 0; i < host.np_tot[IDIR]; i++) {
    const TubeState &s = (host.x[IDIR](i) > HALF_F) ? kRight : kLeft;
    for (int k = 0; k < host.np_tot[KDIR]; k++) {
      for (int j = 0; j < host.np_tot[JDIR]; j++) {
        host.Vc(RHO, k, j, i) = s.rho;
        host.Vc(VX1, k, j, i) = ZERO_F;
#if HAVE_ENERGY
        host.Vc(PRS, k, j, i) = s.prs;
#endif
      }
    }
  }
  host.SyncToDevice();
}
