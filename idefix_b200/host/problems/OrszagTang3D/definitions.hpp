// 3-D Orszag-Tang vortex (BASELINE config 4; the reference's published benchmark problem)
#define COMPONENTS 3
#define DIMENSIONS 3
#define GEOMETRY CARTESIAN
