// 2-D Orszag-Tang vortex (BASELINE config 2)
#define COMPONENTS 2
#define DIMENSIONS 2
#define GEOMETRY CARTESIAN
