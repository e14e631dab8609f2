// Rotating magnetised disc wedge in (r, theta): sub-Keplerian rotation, a density stratified in latitude, a small
// poloidal velocity perturbation, a toroidal field and a poloidal field from A_phi (exactly solenoidal through
// DataBlockHost::MakeVsFromAmag).  Exercises the 2-D spherical area / volume factors, the theta geometric source
// terms (cot theta), the B_theta / B_phi flux factors and the cell-centred evolution of the phi components.
#include "idefix.hpp"
#include "setup.hpp"

Setup::Setup(Input &input, Grid &grid, DataBlock &data, Output &output) {}

void Setup::InitFlow(DataBlock &data) {
  DataBlockHost d(data);
  const real B0 = 0.2;
  IdefixHostArray4D<real> A("A", 3, d.np_tot[KDIR] + KOFFSET, d.np_tot[JDIR] + 1, d.np_tot[IDIR] + 1);
  for (int k = 0; k < d.np_tot[KDIR]; k++) {
    for (int j = 0; j < d.np_tot[JDIR]; j++) {
      for (int i = 0; i < d.np_tot[IDIR]; i++) {
        const real r = d.x[IDIR](i), th = d.x[JDIR](j);
        const real rl = d.xl[IDIR](i), thl = d.xl[JDIR](j);
        const real R = r * sin(th), zh = cos(th) / 0.2;
        d.Vc(RHO, k, j, i) = exp(-0.5 * zh * zh) / (R * sqrt(R)) + 0.01;
        d.Vc(PRS, k, j, i) = 0.02 * d.Vc(RHO, k, j, i) / R;
        d.Vc(VX1, k, j, i) = 0.02 * sin(4.0 * th) / r;
        d.Vc(VX2, k, j, i) = 0.01 * cos(2.0 * r);
        d.Vc(VX3, k, j, i) = 0.9 / sqrt(R);
        d.Vc(BX3, k, j, i) = 0.1 * B0 / R;
        A(KDIR, k, j, i) = B0 * sqrt(rl * sin(thl)) * (1.0 + 0.1 * cos(3.0 * thl));   // A_phi on the cell corners
      }
    }
  }
  d.MakeVsFromAmag(A);
  d.SyncToDevice();
}
