// 2-D Orszag-Tang vortex with the user hooks of the stage enrolled (internal boundary + EMF boundary)
#define COMPONENTS 2
#define DIMENSIONS 2
#define GEOMETRY CARTESIAN
