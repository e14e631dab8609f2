// The same shock tube with USER-DEFINED boundaries: the enrolled function fills the ghost zones
// with a zero-gradient copy of the last active cell from a device loop (idefix_for + KOKKOS_LAMBDA
// on the block's own Vc array) -- for this problem (no inflow) identical to "outflow", which is what
// tests/test_host_cpp.py checks.
#include "idefix.hpp"
#include "setup.hpp"

static void ZeroGradient(Hydro *hydro, int dir, BoundarySide side, const real t) {
  if (dir != IDIR) return;
  DataBlock &data = *hydro->data;
  IdefixArray4D<real> Vc = hydro->Vc;
  const int nvar = Vc.extent(0);
  const int ghosts = data.nghost[IDIR];
  const int first = (side == left) ? 0 : data.end[IDIR];
  const int source = (side == left) ? data.beg[IDIR] : data.end[IDIR] - 1;
  idefix_for("ZeroGradient", 0, nvar, 0, data.np_tot[KDIR], 0, data.np_tot[JDIR], first, first + ghosts,
             KOKKOS_LAMBDA(int n, int k, int j, int i) { Vc(n, k, j, i) = Vc(n, k, j, source); });
}

Setup::Setup(Input &input, Grid &grid, DataBlock &data, Output &output) {
  data.hydro->EnrollUserDefBoundary(&ZeroGradient);
}

void Setup::InitFlow(DataBlock &data) {
  DataBlockHost host(data);
  for (int i = 0; i < host.np_tot[IDIR]; i++) {
    const bool rightHalf = host.x[IDIR](i) > HALF_F;
    host.Vc(RHO, 0, 0, i) = rightHalf ? 0.125 : 1.0;
    host.Vc(VX1, 0, 0, i) = ZERO_F;
    host.Vc(PRS, 0, 0, i) = rightHalf ? 0.1 : 1.0;
  }
  host.SyncToDevice();
}
