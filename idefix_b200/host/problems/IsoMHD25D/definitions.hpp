// 2.5-D isothermal MHD: two dimensions, three vector components (B_z is evolved cell-centred through Uc,
// calcRightHandSide.hpp:541-549), constant sound speed
#define COMPONENTS 3
#define DIMENSIONS 2
#define ISOTHERMAL
#define GEOMETRY CARTESIAN
