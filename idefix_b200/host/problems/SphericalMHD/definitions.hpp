// 3-D MHD in spherical geometry (r, theta, phi) on a logarithmic radial grid (BASELINE config 5)
#define COMPONENTS 3
#define DIMENSIONS 3
#define GEOMETRY SPHERICAL
