// Axisymmetric flow along the polar axis of a rotating, magnetised shell: uniform velocity and field along z (both
// regular on the axis), solid-body rotation, a toroidal field proportional to the cylindrical radius and a dense blob
// on the northern axis: the 2-D axis path (no average around the pole, B_theta = 0 and E_phi = 0 on the axis).
// Compiles unchanged against the reference and against the B200 host layer.
#include "idefix.hpp"
#include "setup.hpp"

Setup::Setup(Input &input, Grid &grid, DataBlock &data, Output &output) {}

void Setup::InitFlow(DataBlock &data) {
  DataBlockHost d(data);
  const int n1 = d.np_tot[IDIR], n2 = d.np_tot[JDIR], n3 = d.np_tot[KDIR];
  IdefixHostArray4D<real> A("A", 3, n3 + KOFFSET, n2 + 1, n1 + 1);
  const real B0 = 0.3, vz = 0.4, omega = 0.2, bphi = 0.15;
  for (int k = 0; k < n3; k++) {
    for (int j = 0; j < n2; j++) {
      for (int i = 0; i < n1; i++) {
        const real r = d.x[IDIR](i), th = d.x[JDIR](j);
        const real rl = d.xl[IDIR](i), thl = d.xl[JDIR](j);
        const real R = r * sin(th), z = r * cos(th);
        const real dist2 = R * R + (z - 2.0) * (z - 2.0);
        d.Vc(RHO, k, j, i) = 1.0 + 0.5 * exp(-dist2 / 0.5);
        d.Vc(PRS, k, j, i) = 1.0;
        d.Vc(VX1, k, j, i) = vz * cos(th);
        d.Vc(VX2, k, j, i) = -vz * sin(th);
        d.Vc(VX3, k, j, i) = omega * R;
        d.Vc(BX3, k, j, i) = bphi * R;
        A(IDIR, k, j, i) = 0.0;
        A(JDIR, k, j, i) = 0.0;
        A(KDIR, k, j, i) = 0.5 * B0 * rl * sin(thl);   // uniform B_z
      }
    }
  }
  d.MakeVsFromAmag(A);
  d.SyncToDevice();
}
