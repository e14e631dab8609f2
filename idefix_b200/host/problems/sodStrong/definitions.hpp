// 1-D hydrodynamic blast tube (adiabatic): pressure ratio 1e5 across x = 0.5
#define COMPONENTS 1
#define DIMENSIONS 1
#define GEOMETRY CARTESIAN
