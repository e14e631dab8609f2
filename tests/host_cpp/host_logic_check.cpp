// Host-only check of the C++ host layer (runs without a GPU): parses an ini file with Input, builds the Grid
// and prints what tests/test_host_cpp.py compares with the Python mirror (itself pinned against the reference):
// coordinates as hex floats, sizes, boundary types, GetOrSet defaults.
#include "idefix.hpp"

int main(int argc, char **argv) {
  Input input(argc, argv);
  Grid grid(input);
  std::printf("regular %d\n", grid.isRegularCartesian ? 1 : 0);
  for (int dir = 0; dir < 3; dir++) {
    std::printf("dir %d np_int %d np_tot %d nghost %d lbound %d rbound %d xbeg %a xend %a\n", dir, grid.np_int[dir],
                grid.np_tot[dir], grid.nghost[dir], static_cast<int>(grid.lbound[dir]), static_cast<int>(grid.rbound[dir]),
                grid.xbeg[dir], grid.xend[dir]);
    for (int i = 0; i < grid.np_tot[dir]; i++)
      std::printf("x %d %d %a %a %a %a\n", dir, i, grid.x[dir](i), grid.xl[dir](i), grid.xr[dir](i), grid.dx[dir](i));
  }
  std::printf("gamma %a\n", input.GetOrSet<real>("Hydro", "gamma", 0, 5.0 / 3.0));
  std::printf("cflmaxvar %a\n", input.GetOrSet<real>("TimeIntegrator", "CFL_max_var", 0, 1.1));
  std::printf("nstages %d\n", input.Get<int>("TimeIntegrator", "nstages", 0));
  std::printf("solver %s\n", input.Get<std::string>("Hydro", "solver", 0).c_str());
  std::printf("check %d %d %d\n", input.CheckEntry("Hydro", "emf"), input.CheckEntry("Grid", "X1-grid"),
              input.CheckBlock("Nope") ? 1 : 0);
  return 0;
}
