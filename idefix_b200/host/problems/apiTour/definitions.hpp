// 2-D hydrodynamics: a tour of the user-facing API (device loops, reductions, analysis, user variables, step hooks)
#define COMPONENTS 2
#define DIMENSIONS 2
#define GEOMETRY CARTESIAN
