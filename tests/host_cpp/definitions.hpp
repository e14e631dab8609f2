// configuration of the host-logic check program (no device code is run)
#define COMPONENTS 3
#define DIMENSIONS 3
#ifndef GEOMETRY
#define GEOMETRY SPHERICAL
#endif
