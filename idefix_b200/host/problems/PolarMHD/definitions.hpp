// 3-D MHD in polar (cylindrical R, phi, z) geometry with a stretched radial grid
#define COMPONENTS 3
#define DIMENSIONS 3
#define GEOMETRY POLAR
