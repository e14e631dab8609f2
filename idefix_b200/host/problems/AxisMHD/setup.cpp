// A uniform flow and a uniform magnetic field, both inclined to the polar axis, cross the poles of a spherical
// shell; a dense blob sits on the northern axis.  Everything is regular on the axis when written in cartesian
// components (the field derives from A = B x r / 2, whose radial component vanishes identically), so the axis
// boundary condition (ghost cells taken from the opposite meridian with the signs of theta- and phi-components
// flipped), the average of the radial EMF around the pole and the regularisation of B_theta on the axis all see
// non-trivial, phi-dependent data.  Compiles unchanged against the reference and against the B200 host layer.
#include "idefix.hpp"
#include "setup.hpp"

Setup::Setup(Input &input, Grid &grid, DataBlock &data, Output &output) {}

void Setup::InitFlow(DataBlock &data) {
  DataBlockHost d(data);
  const int n1 = d.np_tot[IDIR], n2 = d.np_tot[JDIR], n3 = d.np_tot[KDIR];
  IdefixHostArray4D<real> A("A", 3, n3 + 1, n2 + 1, n1 + 1);
  const real Bx = 0.3, Bz = 0.2;          // uniform field (Bx, 0, Bz)
  const real vx = 0.4, vy = 0.3, vz = 0.2; // uniform flow
  // cartesian components of A = B x r / 2 at a point, projected on e_theta / e_phi
  auto Atheta = [&](real r, real th, real ph) {
    const real x = r * sin(th) * cos(ph), y = r * sin(th) * sin(ph), z = r * cos(th);
    const real Ax = -0.5 * Bz * y, Ay = 0.5 * (Bz * x - Bx * z), Az = 0.5 * Bx * y;
    return Ax * cos(th) * cos(ph) + Ay * cos(th) * sin(ph) - Az * sin(th);
  };
  auto Aphi = [&](real r, real th, real ph) {
    const real x = r * sin(th) * cos(ph), y = r * sin(th) * sin(ph), z = r * cos(th);
    const real Ax = -0.5 * Bz * y, Ay = 0.5 * (Bz * x - Bx * z);
    return -Ax * sin(ph) + Ay * cos(ph);
  };
  for (int k = 0; k < n3; k++) {
    for (int j = 0; j < n2; j++) {
      for (int i = 0; i < n1; i++) {
        const real r = d.x[IDIR](i), th = d.x[JDIR](j), ph = d.x[KDIR](k);
        const real rl = d.xl[IDIR](i), thl = d.xl[JDIR](j), phl = d.xl[KDIR](k);
        const real x = r * sin(th) * cos(ph), y = r * sin(th) * sin(ph), z = r * cos(th);
        const real dist2 = (x - 0.3) * (x - 0.3) + y * y + (z - 2.0) * (z - 2.0);
        d.Vc(RHO, k, j, i) = 1.0 + 0.5 * exp(-dist2 / 0.5);
        d.Vc(PRS, k, j, i) = 1.0;
        // unit vectors: e_r = (s c, s s, c), e_theta = (c c, c s, -s), e_phi = (-s, c, 0)
        d.Vc(VX1, k, j, i) = vx * sin(th) * cos(ph) + vy * sin(th) * sin(ph) + vz * cos(th);
        d.Vc(VX2, k, j, i) = vx * cos(th) * cos(ph) + vy * cos(th) * sin(ph) - vz * sin(th);
        d.Vc(VX3, k, j, i) = -vx * sin(ph) + vy * cos(ph);
        // edge-centred potential: A_r on (i, j-1/2, k-1/2), A_theta on (i-1/2, j, k-1/2), A_phi on (i-1/2, j-1/2, k)
        A(IDIR, k, j, i) = 0.0;
        A(JDIR, k, j, i) = Atheta(rl, th, phl);
        A(KDIR, k, j, i) = Aphi(rl, thl, ph);
#ifdef EVOLVE_VECTOR_POTENTIAL
        // the potential itself is the evolved variable; the code derives the face field from it
        d.Ve(AX1e, k, j, i) = A(IDIR, k, j, i);
        d.Ve(AX2e, k, j, i) = A(JDIR, k, j, i);
        d.Ve(AX3e, k, j, i) = A(KDIR, k, j, i);
#endif
      }
    }
  }
#ifndef EVOLVE_VECTOR_POTENTIAL
  d.MakeVsFromAmag(A);
#endif
  d.SyncToDevice();
}
