// 3-D MHD in spherical geometry with the polar axis inside the domain (theta from 0 and / or to pi)
#define COMPONENTS 3
#define DIMENSIONS 3
#define GEOMETRY SPHERICAL
