// 2.5-D MHD in spherical geometry (r, theta) with the polar axis on both sides of the theta range
#define COMPONENTS 3
#define DIMENSIONS 2
#define GEOMETRY SPHERICAL
