// dataBlockHost.hpp -- DataBlockHost: the host mirror a set-up's InitFlow fills
// (reference: src/dataBlock/dataBlockHost.hpp:25-102, dataBlockHost.cpp:108-262).
#include "idefix.hpp"
#ifndef IDEFIX_B200_HOST_DATABLOCKHOST_HPP_
#define IDEFIX_B200_HOST_DATABLOCKHOST_HPP_

class DataBlockHost {
 public:
  std::array<IdefixHostArray1D<real>, 3> x, xr, xl, dx;
  IdefixHostArray4D<real> Vc;
  IdefixHostArray4D<real> Uc;
  IdefixHostArray3D<real> InvDt;
#if MHD == YES
  IdefixHostArray4D<real> Vs;
  IdefixHostArray4D<real> Ve;      // edge-centred vector potential, allocated with EVOLVE_VECTOR_POTENTIAL only
  IdefixHostArray3D<real> Ex1, Ex2, Ex3;   // corner EMFs
#endif
  std::array<real, 3> xbeg{}, xend{};
  std::array<int, 3> np_tot{}, np_int{}, nghost{}, beg{}, end{}, gbeg{}, gend{};
  std::array<BoundaryType, 3> lbound{}, rbound{};
  real t{0.0}, dt{0.0};

  DataBlockHost() = default;
  // constructor from a device DataBlock (does not sync any data, like the reference)
  explicit DataBlockHost(DataBlock &datain) : data(&datain) {
    np_tot = datain.np_tot;
    np_int = datain.np_int;
    nghost = datain.nghost;
    beg = datain.beg;
    end = datain.end;
    gbeg = datain.gbeg;
    gend = datain.gend;
    lbound = datain.lbound;
    rbound = datain.rbound;
    xbeg = datain.xbeg;
    xend = datain.xend;
    t = datain.t;
    dt = datain.dt;
    for (int dir = 0; dir < 3; dir++) {
      x[dir] = datain.mygrid->x[dir];
      xl[dir] = datain.mygrid->xl[dir];
      xr[dir] = datain.mygrid->xr[dir];
      dx[dir] = datain.mygrid->dx[dir];
    }
    const auto &h = *datain.hydro;
    Vc = IdefixHostArray4D<real>("DataBlockHost_Vc", h.Vc.extent(0), h.Vc.extent(1), h.Vc.extent(2), h.Vc.extent(3));
    Uc = IdefixHostArray4D<real>("DataBlockHost_Uc", h.Uc.extent(0), h.Uc.extent(1), h.Uc.extent(2), h.Uc.extent(3));
    InvDt = IdefixHostArray3D<real>("DataBlockHost_InvDt", np_tot[KDIR], np_tot[JDIR], np_tot[IDIR]);
#if MHD == YES
    Vs = IdefixHostArray4D<real>("DataBlockHost_Vs", h.Vs.extent(0), h.Vs.extent(1), h.Vs.extent(2), h.Vs.extent(3));
#ifdef EVOLVE_VECTOR_POTENTIAL
    Ve = IdefixHostArray4D<real>("DataBlockHost_Ve", h.Ve.extent(0), h.Ve.extent(1), h.Ve.extent(2), h.Ve.extent(3));
#endif
    Ex1 = IdefixHostArray3D<real>("DataBlockHost_Ex1", np_tot[KDIR], np_tot[JDIR], np_tot[IDIR]);
    Ex2 = IdefixHostArray3D<real>("DataBlockHost_Ex2", np_tot[KDIR], np_tot[JDIR], np_tot[IDIR]);
    Ex3 = IdefixHostArray3D<real>("DataBlockHost_Ex3", np_tot[KDIR], np_tot[JDIR], np_tot[IDIR]);
#endif
  }

  // pushes everything the mirror holds: t, dt, Vc, InvDt, Vs, corner EMFs, Uc (dataBlockHost.cpp:108-145)
  void SyncToDevice() {
    data->t = t;
    data->dt = dt;
    data->PushTime();
    IDFX_ABI(idfx_upload(data->handle, IDFX_FIELD_VC, Vc.data()));
    IDFX_ABI(idfx_upload(data->handle, IDFX_FIELD_INVDT, InvDt.data()));
    IDFX_ABI(idfx_upload(data->handle, IDFX_FIELD_UC, Uc.data()));
#if MHD == YES
    IDFX_ABI(idfx_upload(data->handle, IDFX_FIELD_VS, Vs.data()));
#ifdef EVOLVE_VECTOR_POTENTIAL
    IDFX_ABI(idfx_upload(data->handle, IDFX_FIELD_VE, Ve.data()));
#endif
#if DIMENSIONS >= 2
    IDFX_ABI(idfx_upload(data->handle, IDFX_FIELD_EZ, Ex3.data()));
#endif
#if DIMENSIONS == 3
    IDFX_ABI(idfx_upload(data->handle, IDFX_FIELD_EX, Ex1.data()));
    IDFX_ABI(idfx_upload(data->handle, IDFX_FIELD_EY, Ex2.data()));
#endif
#endif
  }

  void SyncFromDevice() {
    data->PullTime();
    t = data->t;
    dt = data->dt;
    IDFX_ABI(idfx_download(data->handle, IDFX_FIELD_VC, Vc.data()));
    IDFX_ABI(idfx_download(data->handle, IDFX_FIELD_INVDT, InvDt.data()));
    IDFX_ABI(idfx_download(data->handle, IDFX_FIELD_UC, Uc.data()));
#if MHD == YES
    IDFX_ABI(idfx_download(data->handle, IDFX_FIELD_VS, Vs.data()));
#ifdef EVOLVE_VECTOR_POTENTIAL
    IDFX_ABI(idfx_download(data->handle, IDFX_FIELD_VE, Ve.data()));
#endif
#endif
  }

#if MHD == YES
  // Vs = curl(A) on the staggered mesh (dataBlockHost.cpp:180-262); Ain(dir,k,j,i) lives on cell edges
  void MakeVsFromAmag(IdefixHostArray4D<real> &Ain) {
    auto &dx1 = dx[IDIR], &dx2 = dx[JDIR], &dx3 = dx[KDIR];
    auto &x1m = xl[IDIR], &x1 = x[IDIR];
    [[maybe_unused]] auto &x2m = xl[JDIR];
    (void)dx2; (void)dx3; (void)x1m; (void)x1;
    for (int k = beg[KDIR]; k < end[KDIR] + KOFFSET; k++) {
      for (int j = beg[JDIR]; j < end[JDIR] + JOFFSET; j++) {
        for (int i = beg[IDIR]; i < end[IDIR] + IOFFSET; i++) {
#if GEOMETRY == CARTESIAN
          Vs(BX1s, k, j, i) = D_EXPAND(ZERO_F, +1 / dx2(j) * (Ain(KDIR, k, j + 1, i) - Ain(KDIR, k, j, i)),
                                       -1 / dx3(k) * (Ain(JDIR, k + 1, j, i) - Ain(JDIR, k, j, i)));
#if DIMENSIONS >= 2
          Vs(BX2s, k, j, i) = D_EXPAND(-1 / dx1(i) * (Ain(KDIR, k, j, i + 1) - Ain(KDIR, k, j, i)), ,
                                       +1 / dx3(k) * (Ain(IDIR, k + 1, j, i) - Ain(IDIR, k, j, i)));
#endif
#if DIMENSIONS == 3
          Vs(BX3s, k, j, i) = 1 / dx1(i) * (Ain(JDIR, k, j, i + 1) - Ain(JDIR, k, j, i)) -
                              1 / dx2(j) * (Ain(IDIR, k, j + 1, i) - Ain(IDIR, k, j, i));
#endif
#elif GEOMETRY == POLAR
          Vs(BX1s, k, j, i) = D_EXPAND(ZERO_F, +1 / (x1m(i) * dx2(j)) * (Ain(KDIR, k, j + 1, i) - Ain(KDIR, k, j, i)),
                                       -1 / dx3(k) * (Ain(JDIR, k + 1, j, i) - Ain(JDIR, k, j, i)));
#if DIMENSIONS >= 2
          Vs(BX2s, k, j, i) = D_EXPAND(-1 / dx1(i) * (Ain(KDIR, k, j, i + 1) - Ain(KDIR, k, j, i)), ,
                                       +1 / dx3(k) * (Ain(IDIR, k + 1, j, i) - Ain(IDIR, k, j, i)));
#endif
#if DIMENSIONS == 3
          Vs(BX3s, k, j, i) = 1 / (x1(i) * dx1(i)) * (x1m(i + 1) * Ain(JDIR, k, j, i + 1) - x1m(i) * Ain(JDIR, k, j, i)) -
                              1 / (x1(i) * dx2(j)) * (Ain(IDIR, k, j + 1, i) - Ain(IDIR, k, j, i));
#endif
#elif GEOMETRY == SPHERICAL
          // the reference reads x2m(j+1), sin/cos of the face colatitudes (dataBlockHost.cpp:228-262)
          const real th0 = x2m(j), th1 = (j + 1 < np_tot[JDIR]) ? x2m(j + 1) : xr[JDIR](j);
          [[maybe_unused]] const real dmu = std::cos(th0) - std::cos(th1);
          Vs(BX1s, k, j, i) = D_EXPAND(ZERO_F,
                                       +1 / (x1m(i) * dmu) * (std::sin(th1) * Ain(KDIR, k, j + 1, i) - std::sin(th0) * Ain(KDIR, k, j, i)),
                                       -dx2(j) / (x1m(i) * dmu * dx3(k)) * (Ain(JDIR, k + 1, j, i) - Ain(JDIR, k, j, i)));
#if DIMENSIONS >= 2
          real Ax2m = std::fabs(std::sin(th0));
          if (std::fabs(Ax2m) < 1e-12) Ax2m = ONE_F;
          Vs(BX2s, k, j, i) = D_EXPAND(-1 / (x1(i) * dx1(i)) * (x1m(i + 1) * Ain(KDIR, k, j, i + 1) - x1m(i) * Ain(KDIR, k, j, i)), ,
                                       +1 / (x1(i) * Ax2m * dx3(k)) * (Ain(IDIR, k + 1, j, i) - Ain(IDIR, k, j, i)));
#endif
#if DIMENSIONS == 3
          Vs(BX3s, k, j, i) = 1 / (x1(i) * dx1(i)) * (x1m(i + 1) * Ain(JDIR, k, j, i + 1) - x1m(i) * Ain(JDIR, k, j, i)) -
                              1 / (x1(i) * dx2(j)) * (Ain(IDIR, k, j + 1, i) - Ain(IDIR, k, j, i));
#endif
#else
          IDEFIX_ERROR("MakeVsFromAmag: geometry not yet defined");
#endif
        }
      }
    }
  }
#endif

 private:
  DataBlock *data{nullptr};
};

#endif  // IDEFIX_B200_HOST_DATABLOCKHOST_HPP_
