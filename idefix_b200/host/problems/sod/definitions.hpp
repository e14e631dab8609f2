// 1-D hydrodynamic shock tube (BASELINE config 1)
#define COMPONENTS 1
#define DIMENSIONS 1
#define GEOMETRY CARTESIAN
