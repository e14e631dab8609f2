// 1-D isothermal hydrodynamic shock tube
#define COMPONENTS 1
#define DIMENSIONS 1
#define ISOTHERMAL
#define GEOMETRY CARTESIAN
