// 2-D hydrodynamics in polar geometry (R, phi)
#define COMPONENTS 2
#define DIMENSIONS 2
#define GEOMETRY POLAR
