// 2.5-D MHD in spherical geometry (r, theta) with three vector components: v_phi and B_phi are carried
// cell-centred, the poloidal field by constrained transport
#define COMPONENTS 3
#define DIMENSIONS 2
#define GEOMETRY SPHERICAL
