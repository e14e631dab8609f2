// grid.hpp -- Grid / GridHost: global grid sizes, boundary types and coordinates
// (reference: src/grid.hpp, src/grid.cpp:45-269, src/gridHost.cpp:44-266).
//
// Coordinates are generated expression by expression like GridHost::MakeGrid (uniform "u",
// logarithmic "l", stretched "s+" / "s-" patches with the same Newton iteration) with the same
// libm calls, so they are bit-identical with the reference's.
#include "idefix.hpp"
#ifndef IDEFIX_B200_HOST_GRID_HPP_
#define IDEFIX_B200_HOST_GRID_HPP_

class Grid {
 public:
  std::array<IdefixHostArray1D<real>, 3> x, xr, xl, dx;   // host copies (GridHost members)
  std::array<real, 3> xbeg{}, xend{};
  std::array<int, 3> np_tot{1, 1, 1}, np_int{1, 1, 1}, nghost{0, 0, 0};
  std::array<BoundaryType, 3> lbound{undefined, undefined, undefined}, rbound{undefined, undefined, undefined};
  std::array<int, 3> nproc{1, 1, 1}, xproc{0, 0, 0};
  bool haveAxis{false};
  bool isRegularCartesian{GEOMETRY == CARTESIAN};   // grid.cpp:49-51

  Grid() = default;
  explicit Grid(Input &input) {
    for (int dir = 0; dir < DIMENSIONS; dir++) {
      nghost[dir] = (ORDER < 4) ? 2 : 3;   // grid.cpp:61-65
      const std::string label = "X" + std::to_string(dir + 1) + "-grid";
      const int npatch = input.Get<int>("Grid", label, 0);
      np_int[dir] = 0;
      for (int p = 0; p < npatch; p++) np_int[dir] += input.Get<int>("Grid", label, 2 + 3 * p);
    }
    for (int dir = 0; dir < 3; dir++) np_tot[dir] = np_int[dir] + 2 * nghost[dir];
    for (int dir = 0; dir < DIMENSIONS; dir++) {
      for (int side = 0; side < 2; side++) {
        const std::string label = "X" + std::to_string(dir + 1) + (side == 0 ? "-beg" : "-end");
        const std::string name = input.Get<std::string>("Boundary", label, 0);
        BoundaryType t = undefined;
        if (name == "outflow") t = outflow;
        else if (name == "reflective") t = reflective;
        else if (name == "periodic") t = periodic;
        else if (name == "internal") t = internal;
        else if (name == "shearingbox") t = shearingbox;
        else if (name == "axis") t = axis;
        else if (name == "userdef") t = userdef;
        else IDEFIX_ERROR("Unknown boundary type " + name);
        if (t == axis) {   // grid.cpp:98-104
          if (dir != JDIR) IDEFIX_ERROR("Axis Boundaries are only applicable to X2");
          haveAxis = true;
        }
        if (t == shearingbox) IDEFIX_ERROR("shearing-box boundaries are outside the scope of this build");
        (side == 0 ? lbound : rbound)[dir] = t;
      }
    }
    MakeGrid(input);
  }

  void ShowConfig() {
    idfx::cout << "Grid: full grid size is " << std::endl;
    const char *bname[] = {"internal", "periodic", "reflective", "outflow", "shearingbox", "axis", "userdef", "undefined"};
    for (int dir = 0; dir < DIMENSIONS; dir++) {
      idfx::cout << "\t Direction X" << (dir + 1) << ": " << bname[lbound[dir]] << "\t" << xbeg[dir] << "...."
                 << np_int[dir] << "...." << xend[dir] << "\t" << bname[rbound[dir]] << std::endl;
    }
  }

 private:
  struct Patch {
    std::string type;
    real start, end;
    int size;
  };

  void MakeGrid(Input &input) {
    for (int dir = 0; dir < 3; dir++) {
      const int n = np_tot[dir];
      x[dir] = IdefixHostArray1D<real>("Grid_x", n);
      xl[dir] = IdefixHostArray1D<real>("Grid_xl", n);
      xr[dir] = IdefixHostArray1D<real>("Grid_xr", n);
      dx[dir] = IdefixHostArray1D<real>("Grid_dx", n);
      auto &X = x[dir], &XL = xl[dir], &XR = xr[dir], &DX = dx[dir];
      if (dir >= DIMENSIONS) {   // unused direction: one cell (gridHost.cpp:246-262)
        real xs = -0.5, xe = 0.5;
        if (GEOMETRY == SPHERICAL && dir == JDIR) {
          xs = 0.0;
          xe = M_PI;
        }
        xbeg[dir] = xs;
        xend[dir] = xe;
        for (int i = 0; i < n; i++) {
          DX(i) = xe - xs;
          X(i) = 0.5 * (xs + xe);
          XL(i) = xs;
          XR(i) = xe;
        }
        continue;
      }
      const std::string label = "X" + std::to_string(dir + 1) + "-grid";
      const int npatch = input.Get<int>("Grid", label, 0);
      std::vector<Patch> patch(npatch);
      for (int p = 0; p < npatch; p++) {
        patch[p].type = input.Get<std::string>("Grid", label, 3 + 3 * p);
        patch[p].start = input.Get<real>("Grid", label, 1 + 3 * p);
        patch[p].end = input.Get<real>("Grid", label, 4 + 3 * p);
        patch[p].size = input.Get<int>("Grid", label, 2 + 3 * p);
      }
      xbeg[dir] = patch[0].start;
      xend[dir] = patch[npatch - 1].end;
      const int ng = nghost[dir];
      // first pass: uniform and logarithmic patches
      int idx = ng;
      for (int p = 0; p < npatch; p++) {
        const Patch &P = patch[p];
        const int gs = (p == 0) ? ng : 0, ge = (p == npatch - 1) ? ng : 0;
        if (P.type == "u") {
          const real delta = (P.end - P.start) / P.size;
          for (int i = -gs; i < P.size + ge; i++) {
            DX(i + idx) = delta;
            X(i + idx) = P.start + (i + 0.5) * delta;
            XL(i + idx) = P.start + i * delta;
            XR(i + idx) = P.start + (i + 1) * delta;
          }
        } else if (P.type == "l") {
          isRegularCartesian = false;
          const real alpha = (P.end + std::fabs(P.start) - P.start) / std::fabs(P.start);
          for (int i = -gs; i < P.size + ge; i++) {
            XL(i + idx) = P.start * std::pow(alpha, static_cast<double>(i) / static_cast<double>(P.size));
            XR(i + idx) = P.start * std::pow(alpha, static_cast<double>(i + 1) / static_cast<double>(P.size));
            DX(i + idx) = XR(i + idx) - XL(i + idx);
            X(i + idx) = 0.5 * (XR(i + idx) + XL(i + idx));
          }
        } else if (P.type != "s+" && P.type != "s-") {
          IDEFIX_ERROR("GridHost::MakeGrid: Unknown grid type :" + P.type);
        }
        idx += P.size;
      }
      // second pass: stretched patches lean on a neighbouring uniform / log patch
      idx = ng;
      for (int p = 0; p < npatch; p++) {
        const Patch &P = patch[p];
        const int gs = (p == 0) ? ng : 0, ge = (p == npatch - 1) ? ng : 0;
        if (P.type == "s+" || P.type == "s-") {
          isRegularCartesian = false;
          const int ref = (P.type == "s+") ? p + 1 : p - 1;
          if (ref < 0 || ref >= npatch)
            IDEFIX_ERROR("You're attempting to construct a stretched patch from a non-existent patch");
          if (patch[ref].type != "u" && (patch[ref].type != "l" || ref == p + 1))
            IDEFIX_ERROR("You can only construct a stretched patch from a uniform grid or AFTER a logarithmic grid");
          const int offset = (ref == p + 1) ? P.size : -1;
          const real delta = DX(idx + offset);
          const real logdelta = std::log((P.end - P.start) / delta);
          if (std::fabs((P.end - P.start) / P.size - delta) < 1e-10)
            IDEFIX_ERROR("A Stretch grid can be defined only if the stretched domain has a mean\n"
                         "spacing different from the reference uniform grid.");
          real q = 1.05;
          for (int it = 0; it <= 50; it++) {   // Newton iterations (gridHost.cpp:165-175)
            const real f = std::log((std::pow(q, P.size + 1) - q) / (q - 1)) - logdelta;
            const real fp = ((P.size + 1) * std::pow(q, P.size) - 1) / (std::pow(q, P.size + 1) - q) - 1 / (q - 1);
            const real dq = f / fp;
            q = q - dq;
            if (std::fabs(dq) < 1e-14 * q) break;
            if (it == 50) IDEFIX_ERROR("Failed to create the stretched grid");
          }
          for (int i = -gs; i < P.size + ge; i++) {
            if (P.type == "s-") {
              XL(i + idx) = P.start + q * (std::pow(q, i) - 1) / (q - 1) * delta;
              XR(i + idx) = P.start + q * (std::pow(q, i + 1) - 1) / (q - 1) * delta;
              DX(i + idx) = std::pow(q, i + 1) * delta;
            } else {
              XL(i + idx) = P.end - q * (std::pow(q, P.size - i) - 1) / (q - 1) * delta;
              XR(i + idx) = P.end - q * (std::pow(q, P.size - i - 1) - 1) / (q - 1) * delta;
              DX(i + idx) = std::pow(q, P.size - i) * delta;
            }
            X(i + idx) = 0.5 * (XR(i + idx) + XL(i + idx));
          }
        }
        idx += P.size;
      }
    }
    // Axis treatment (gridHost.cpp:227-263): the domain must touch the axis, and the theta spacing is mirrored
    // across it (the ghost faces are rebuilt outwards from the axis face; the cell centres keep their values)
    if (haveAxis) {
#if GEOMETRY != SPHERICAL
      IDEFIX_ERROR("Axis boundaries only compatible with Spherical boundary conditions");
#endif
#if DIMENSIONS < 2
      IDEFIX_ERROR("Axis Boundaries requires at least two dimenions");
#endif
      const real smallNumber = 1e-10;
      if ((std::fabs(xbeg[JDIR]) > smallNumber) && (lbound[JDIR] == axis))
        IDEFIX_ERROR("Axis Boundaries requires your X2 domain to start at exactly x2=0.0");
      if ((std::fabs(xend[JDIR] - M_PI) > smallNumber) && (rbound[JDIR] == axis))
        IDEFIX_ERROR("Axis Boundaries requires your X2 domain to end at exactly x2=Pi");
      if (lbound[JDIR] == axis) {
        const int jref = nghost[JDIR];
        for (int j = jref - 1; j >= 0; j--) {
          dx[JDIR](j) = dx[JDIR](2 * jref - j - 1);
          xl[JDIR](j) = xl[JDIR](j + 1) - dx[JDIR](j);
          xr[JDIR](j) = xl[JDIR](j + 1);
        }
      }
      if (rbound[JDIR] == axis) {
        const int jref = nghost[JDIR] + np_int[JDIR] - 1;
        for (int j = jref + 1; j < np_tot[JDIR]; j++) {
          dx[JDIR](j) = dx[JDIR](2 * jref - j + 1);
          xr[JDIR](j) = xr[JDIR](j - 1) + dx[JDIR](j);
          xl[JDIR](j) = xr[JDIR](j - 1);
        }
      }
    }
  }
};

// GridHost in the reference is the host mirror of Grid; Grid already lives on the host here.
class GridHost : public Grid {
 public:
  GridHost() = default;
  explicit GridHost(Grid &g) : Grid(g) {}
  void SyncFromDevice() {}
  void SyncToDevice() {}
};

#endif  // IDEFIX_B200_HOST_GRID_HPP_
