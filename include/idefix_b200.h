/* idefix_b200.h -- C ABI of the B200-native explicit Godunov (M)HD update.
 *
 * One opaque handle (IdfxBlock) per (GPU, sub-domain).  Every entry point
 * mirrors one call site of the reference's hot loop (cited as file:line under
 * the reference tree) so that the reference's DataBlock / Fluid /
 * TimeIntegrator methods can forward to it 1:1.
 *
 * Conventions
 *  - plain C types only; all arrays are FP64; index order (n,k,j,i), i fastest
 *    (reference layout: src/arrays.hpp:12-19, src/idefix.hpp:20).
 *  - every function returns 0 on success, non-zero on failure; the message
 *    is retrievable with idfx_last_error().  No exception crosses the ABI.
 *  - kernels are enqueued on the block's stream; calls returning host scalars
 *    synchronise that stream.
 *  - there is NO CPU fallback: creating a block without a CUDA device fails.
 */
#ifndef IDEFIX_B200_H_
#define IDEFIX_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- enumerations (values of the boundary enum follow src/idefix.hpp:220) ---- */
enum IdfxGeometry { IDFX_CARTESIAN = 1, IDFX_CYLINDRICAL = 2, IDFX_POLAR = 3, IDFX_SPHERICAL = 4 };
enum IdfxSolver   { IDFX_TVDLF = 0, IDFX_HLL = 1, IDFX_HLLC = 2, IDFX_HLLD = 3, IDFX_ROE = 4 };
enum IdfxEmf      { IDFX_EMF_ARITHMETIC = 0, IDFX_EMF_UCT0 = 1, IDFX_EMF_UCT_CONTACT = 2,
                    IDFX_EMF_UCT_HLL = 3, IDFX_EMF_UCT_HLLD = 4 };
enum IdfxBoundary { IDFX_BC_INTERNAL = 0, IDFX_BC_PERIODIC = 1, IDFX_BC_REFLECTIVE = 2,
                    IDFX_BC_OUTFLOW = 3, IDFX_BC_SHEARINGBOX = 4, IDFX_BC_AXIS = 5,
                    IDFX_BC_USERDEF = 6, IDFX_BC_UNDEFINED = 7 };

/* Field identifiers for idfx_field_info / idfx_upload / idfx_download. */
enum IdfxField {
  IDFX_FIELD_VC = 0,     /* primitive variables        (nvar, n3, n2, n1)              */
  IDFX_FIELD_UC = 1,     /* conservative variables     (nvar, n3, n2, n1)              */
  IDFX_FIELD_VS = 2,     /* staggered field            (ndim, n3+o3, n2+o2, n1+1)      */
  IDFX_FIELD_INVDT = 3,  /* inverse time-step          (n3, n2, n1)                    */
  IDFX_FIELD_EX = 4,     /* corner EMFs                (n3, n2, n1)                    */
  IDFX_FIELD_EY = 5,
  IDFX_FIELD_EZ = 6,
  IDFX_FIELD_UC0 = 7,    /* states["begin"] copy of Uc                                  */
  IDFX_FIELD_VS0 = 8,    /* states["begin"] copy of Vs                                  */
  IDFX_FIELD_VE = 9,     /* EVOLVE_VECTOR_POTENTIAL: edge-centred vector potential, laid out like Vs:
                            (3, n3+o3, n2+o2, n1+1) = AX1e, AX2e, AX3e in 3-D, (1, ...) = AX3e in 2-D */
  IDFX_FIELD_VE0 = 10    /* states["begin"] copy of Ve                                  */
};

/* Compile-time configuration of the reference (definitions.hpp macros and CMake
 * options, CMakeLists.txt:12-33) expressed as a run-time descriptor. */
typedef struct IdfxBlockDesc {
  int ndim;            /* DIMENSIONS 1..3                                               */
  int ncomp;           /* COMPONENTS 1..3 (>= ndim)                                     */
  int mhd;             /* Idefix_MHD                                                    */
  int isothermal;      /* ISOTHERMAL defined (constant cs only)                         */
  int geometry;        /* enum IdfxGeometry                                             */
  int order;           /* Idefix_RECONSTRUCTION: 1 constant, 2 PLM, 3 LimO3, 4 PPM      */
  int solver;          /* [Hydro] solver   (riemannSolver.hpp:100-140)                  */
  int emf;             /* [Hydro] emf      (constrainedTransport.hpp:157-178)           */
  int nstages;         /* [TimeIntegrator] nstages 1..3                                 */
  int np_int[3];       /* interior cells of THIS block                                  */
  int nghost[3];       /* ghost depth (0 in unused dims; grid.cpp:55-66)                */
  int lbound[3];       /* enum IdfxBoundary, left side per direction                    */
  int rbound[3];       /* enum IdfxBoundary, right side per direction                   */
  int regular_grid;    /* Grid::isRegularCartesian (grid.cpp:49-51, gridHost.cpp:89)    */
  int device;          /* CUDA device ordinal                                           */
  double gamma;        /* [Hydro] gamma (eos_adiabatic.hpp:21)                          */
  double csiso;        /* [Hydro] csiso constant <value>                                */
  double cfl;          /* [TimeIntegrator] CFL                                          */
  double cfl_max_var;  /* [TimeIntegrator] CFL_max_var (default 1.1)                    */
  double first_dt;     /* [TimeIntegrator] first_dt                                     */
  double fixed_dt;     /* [TimeIntegrator] fixed_dt, <= 0 when absent                   */
  /* cell coordinates of this block incl. ghosts, HOST pointers, length
   * np_int[d]+2*nghost[d] (DataBlock::x/xl/xr/dx, dataBlock.cpp:68-84). */
  const double *x[3];
  const double *xl[3];
  const double *xr[3];
  const double *dx[3];
  int ntracer;         /* [Hydro] tracer N: passive tracers stored behind the fluid variables in
                          Vc / Uc, index nvar .. nvar+N-1 (fluid.hpp:315-323, tracer.hpp:58-125) */
  double shock_flattening; /* [Hydro] shockFlattening <smoothing> (riemannSolver.hpp:154-159, shockFlattening.hpp):
                          > 0 flags the cells of converging flow with |grad P| > smoothing * min(P) before the x1
                          sweep of every stage; PLM / LimO3 drop to the reference's MinMod there.  0 = off */
  int evolve_vector_potential; /* EVOLVE_VECTOR_POTENTIAL (Idefix_EVOLVE_VECTOR_POTENTIAL, idefix.hpp:116-125): the
                          evolved magnetic variable is the edge-centred vector potential Ve (Ve += -dt E), the face
                          field is derived from it (Vs = curl Ve, evolveVectorPotential.hpp); needs MHD, DIMENSIONS >= 2 */
} IdfxBlockDesc;

typedef struct IdfxBlock IdfxBlock;

/* ---- life cycle ------------------------------------------------------------ */
const char *idfx_last_error(void);
const char *idfx_version(void);
int idfx_device_count(int *count);
/* replaces DataBlock::DataBlock + Fluid ctor allocation (dataBlock.cpp:22-160, fluid.hpp:529-577) */
int idfx_block_create(const IdfxBlockDesc *desc, IdfxBlock **out);
int idfx_block_destroy(IdfxBlock *blk);
int idfx_sync(IdfxBlock *blk);
/* Enqueue on a caller-provided cudaStream_t instead of the block's own stream (the reference
 * runs everything on Kokkos' default execution-space instance, src/loop.hpp:56-61). */
int idfx_set_stream(IdfxBlock *blk, void *cuda_stream);

/* ---- field access ----------------------------------------------------------- */
/* dims[4] receives (n, n3, n2, n1); dev receives the device pointer (may be NULL). */
int idfx_field_info(IdfxBlock *blk, int field, double **dev, int dims[4]);
/* DataBlockHost::SyncToDevice / SyncFromDevice (dataBlockHost.cpp:108-190) */
int idfx_upload(IdfxBlock *blk, int field, const double *host);
int idfx_download(IdfxBlock *blk, int field, double *host);
int idfx_set_time(IdfxBlock *blk, double t, double dt);
int idfx_get_time(IdfxBlock *blk, double *t, double *dt);

/* ---- granular entry points: one per reference call site ---------------------- */
int idfx_reset_stage(IdfxBlock *blk);                 /* Fluid::ResetStage        enroll.hpp:116-129            */
int idfx_set_boundaries(IdfxBlock *blk, double t);    /* Boundary::SetBoundaries  boundary.hpp:240-282          */
int idfx_enforce_boundary_dir(IdfxBlock *blk, double t, int dir);  /* boundary.hpp:287-385 (+ReconstructNormalField) */
int idfx_reconstruct_vc_field(IdfxBlock *blk, int field);          /* boundary.hpp:389-404; field = VC or UC */
int idfx_prim_to_cons(IdfxBlock *blk);                /* Fluid::ConvertPrimToCons convertConsToPrim.hpp:153-189 */
int idfx_cons_to_prim(IdfxBlock *blk);                /* Fluid::ConvertConsToPrim convertConsToPrim.hpp:106-149 */
int idfx_state_copy(IdfxBlock *blk);                  /* StateContainer::CopyFrom stateContainer.cpp:19-37      */
int idfx_state_add_and_store(IdfxBlock *blk, double wc, double w0);  /* stateContainer.cpp:75-104              */
int idfx_evolve_stage(IdfxBlock *blk, double t, double dt);          /* Fluid::EvolveStage evolveStage.hpp:41-90 */
int idfx_compute_timestep(IdfxBlock *blk, double *dt_min);           /* DataBlock::ComputeTimestep dataBlock.cpp:333-364 */
/* DataBlock::DeriveVectorPotential (dataBlock.cpp:366-373): Vs = curl Ve; what main() calls after InitFlow when the
 * vector potential is the evolved variable (main.cpp:141-174).  No-op otherwise. */
int idfx_derive_vector_potential(IdfxBlock *blk);
int idfx_check_divb(IdfxBlock *blk, double *divb_max);               /* Fluid::CheckDivB checkDivB.hpp:15-68     */
int idfx_check_nan(IdfxBlock *blk, int *nan_count);                  /* Fluid::CheckNan  checkNan.hpp:15-131     */

/* ---- fused path: TimeIntegrator::Cycle (timeIntegrator.cpp:246-414) ---------- */
/* Runs ncycles full RK cycles with the dt control logic evaluated on the device
 * (no host round trip); results are identical to the granular sequence. */
int idfx_cycles(IdfxBlock *blk, int ncycles);
/* One cycle, host-synchronous: returns the new t and dt like Cycle() leaves them
 * in DataBlock::t/dt.  status_flags: bit0 dt<1e-15, bit1 first_dt too large, bit2 the sweeps left the
 * validated envelope of their branch-free FP64 divisions / square roots (idfx_math_exceptions > 0: operands that
 * are subnormal, beyond 2^+-140, infinite or NaN) -- the results are no longer guaranteed to be the reference's
 * and the host layers stop the run; a library built with `make FASTDIV=0` has no such envelope. */
int idfx_cycle(IdfxBlock *blk, double *t, double *dt, int *status_flags);
/* Limit dt so that the run ends at tstop (main.cpp:191). Negative disables. */
int idfx_set_tstop(IdfxBlock *blk, double tstop);
/* fused = 0 runs the cycle as the reference's own kernel sequence (one kernel per reference
 * kernel, whole-array point-wise passes); fused = 1 (default unless a boundary is userdef)
 * folds the point-wise passes into the sweeps.  Both give bit-identical Vc / Vs. */
int idfx_set_fused(IdfxBlock *blk, int fused);
int idfx_get_ncycles(IdfxBlock *blk, long long *ncycles);
/* The sweeps evaluate FP64 divisions and square roots with the branch-free form of nvcc's own Newton
 * sequences (idefix_b200/csrc/fastmath.cuh); a thread whose operands fail nvcc's range check (zero,
 * subnormal, infinite or NaN operands or results) is counted here instead of taking nvcc's slow path.
 * count == 0 (asserted by the parity tests and bench.py) means every quotient and root was the correctly
 * rounded IEEE value.  Build with `make FASTDIV=0` for a library that uses the ordinary operators only. */
int idfx_math_exceptions(IdfxBlock *blk, unsigned int *count);
/* Device self-test of the tensor-memory (TMEM) scratchpad the x2 / x3 sweeps keep their per-thread carry-over in
 * (csrc/tmem.cuh): four co-resident CTAs per SM write, re-read and compare a data pattern.  *mismatches == 0 on a
 * healthy sm_100a device.  No reference counterpart (the reference has no tensor-memory path). */
int idfx_selftest_tmem(int device, long long *mismatches);
/* Device self-test of the TMA loads of the sweeps (csrc/tma.cuh): rank-4 FP64 boxes [nvar][1][1][cells] per warp,
 * completion on a per-warp mbarrier, boxes crossing the array bounds (zero fill), compared element by element with
 * plain loads.  *mismatches == 0 on a healthy device; -2 = the driver refused the tensor-map descriptor. */
int idfx_selftest_tma(int device, long long *mismatches);
/* number of kernels launched by this block since creation (bench "gpu_launches") */
int idfx_launch_count(IdfxBlock *blk, long long *count);

/* Per-kernel-class device timing with CUDA events on the launching stream (the role of the
 * reference's -profile region timers, src/profiler.cpp:59-203).  mode 1 starts sampling, 0 stops,
 * 2 collects: out[2*c] = number of brackets, out[2*c+1] = total milliseconds, for the classes
 * c = 0,1,2 X1/X2/X3 sweep kernel, 3 constrained transport, 4 stage epilogue, 5 boundaries. */
#define IDFX_TIMING_CLASSES 6
int idfx_timing(IdfxBlock *blk, int mode, double *out);

/* ---- callbacks --------------------------------------------------------------- */
/* UserDefBoundaryFunc (boundary.hpp:28-35): invoked for sides declared userdef.
 * side: 0 left, 1 right.  The block's stream is synchronised around the call. */
/* Host callbacks inside a stage, the user hooks of src/fluid/enroll.hpp:33,71:
 *   IDFX_HOOK_INTERNAL_BOUNDARY  InternalBoundaryFunc: first thing in Boundary::SetBoundaries (boundary.hpp:243-251),
 *                                may rewrite Vc / Vs anywhere (e.g. cap a velocity, floor a density);
 *   IDFX_HOOK_EMF_BOUNDARY       EmfBoundaryFunc: in EnforceEMFBoundary (enforceEMFBoundary.hpp:29-30), between
 *                                CalcCornerEMF and EvolveMagField, may rewrite the corner EMFs (IDFX_FIELD_EX/EY/EZ).
 * The library synchronises its stream before the call and the device after it, so the callback may launch its own
 * kernels on any stream (or touch the arrays through idfx_field_info pointers).  A block with a hook runs its
 * cycles without CUDA-graph replay. */
enum IdfxStageHook { IDFX_HOOK_INTERNAL_BOUNDARY = 0, IDFX_HOOK_EMF_BOUNDARY = 1 };
typedef void (*IdfxStageHookFn)(IdfxBlock *blk, double t, void *user);
int idfx_enroll_stage_hook(IdfxBlock *blk, int which, IdfxStageHookFn fn, void *user);

typedef void (*IdfxUserBoundaryFn)(IdfxBlock *blk, int dir, int side, double t, void *user);
int idfx_enroll_userdef_boundary(IdfxBlock *blk, IdfxUserBoundaryFn fn, void *user);
/* Halo exchange hook, replaces Mpi::ExchangeX1/X2/X3 (mpi.cpp:247,449,646) for
 * sides declared internal or periodic with more than one block in `dir`.
 * Called from idfx_set_boundaries / idfx_cycles once per direction. */
typedef void (*IdfxHaloFn)(IdfxBlock *blk, int dir, void *user);
int idfx_enroll_halo_exchange(IdfxBlock *blk, IdfxHaloFn fn, void *user);

/* Cross-block reduction hook, replaces the MPI_(I)allreduce sites of the hot path: the CFL time
 * step (MIN, timeIntegrator.cpp:318,358) and div B (MAX, checkDivB.hpp:63).  The hook must reduce
 * the device scalar *dev_value in place over all blocks, enqueued on the block's stream. */
enum IdfxReduceOp { IDFX_REDUCE_MIN = 0, IDFX_REDUCE_MAX = 1 };
typedef void (*IdfxReduceFn)(IdfxBlock *blk, double *dev_value, int op, void *user);
int idfx_enroll_reduce(IdfxBlock *blk, IdfxReduceFn fn, void *user);

/* ---- halo exchange building blocks (mpi.cpp:79-134 buffer layout) ------------ */
/* Number of doubles in the send/recv buffer of direction dir. */
int idfx_halo_buffer_size(IdfxBlock *blk, int dir, size_t *ndoubles);
/* Device pointers of the block-owned send/recv buffers; side 0 = left, 1 = right. */
int idfx_halo_buffers(IdfxBlock *blk, int dir, int side, double **send, double **recv);
/* Pack the two send buffers of `dir` (Buffer::Pack, mpi.hpp:49-110). */
int idfx_halo_pack(IdfxBlock *blk, int dir);
/* Unpack the two recv buffers of `dir` into the ghost zones (mpi.hpp:112-172). */
int idfx_halo_unpack(IdfxBlock *blk, int dir);
/* Peer-to-peer variant: bind the neighbour's recv buffer (a device pointer valid in
 * this process, e.g. from cudaIpcOpenMemHandle) so that idfx_halo_pack writes
 * straight into it over NVLink. side = side of THIS block that faces the peer. */
int idfx_halo_bind_peer(IdfxBlock *blk, int dir, int side, double *peer_recv);
int idfx_halo_ipc_handle(IdfxBlock *blk, int dir, int side, void *handle64);
int idfx_halo_open_ipc(IdfxBlock *blk, const void *handle64, double **ptr);
/* tell the block that `dir` has more than one block (nproc[dir] > 1) */
int idfx_set_nproc(IdfxBlock *blk, const int nproc[3]);

/* ---- peer-to-peer exchange over NVLink / NVSwitch, no host in the loop ------------------------
 * Replaces Mpi::ExchangeX1/X2/X3 (src/mpi.cpp:247-846) and the MPI_Iallreduce of the time step
 * (src/timeIntegrator.cpp:318,357-359) for the blocks of ONE box: per direction one kernel stores
 * both slabs straight into the neighbours' receive buffers and raises a message number in their
 * device "mailbox", a second one waits for the incoming numbers and fills the ghost zones; the dt
 * minimum travels through the same mailboxes.  Once connected, idfx_set_boundaries / idfx_cycles
 * need no halo or reduce callback and the whole cycle (exchanges included) replays as a CUDA graph.
 *
 * idfx_p2p_mailbox: device pointer of this block's mailbox and (ipc_handle64 != NULL) its 64-byte
 *   cudaIpcMemHandle_t for blocks living in other processes (open it with idfx_halo_open_ipc).
 * idfx_p2p_connect: rank / world = position of this block in the domain (rank = xproc[0] +
 *   nproc[0]*(xproc[1] + nproc[1]*xproc[2]), the reference's Cartesian communicator, grid.cpp:222-223);
 *   peer_recv[dir][side] = receive buffer (idfx_halo_buffers) of ghost side 1-side of the neighbour
 *   on `side`; peer_mailbox[dir][side] = that neighbour's mailbox; all_mailbox[world] = every block's
 *   mailbox (own included).  Pointers must be valid on this block's device (same process with peer
 *   access enabled, or cudaIpcOpenMemHandle).  Call idfx_set_nproc first.
 * Blocks driven by ONE host thread must be given their work with the non-blocking idfx_cycles on
 * every block before any of them is synchronised (a block's receive kernel waits for its neighbour). */
int idfx_p2p_mailbox(IdfxBlock *blk, void **mailbox, void *ipc_handle64);
int idfx_p2p_connect(IdfxBlock *blk, int rank, int world, void *const peer_recv[3][2],
                     void *const peer_mailbox[3][2], void *const *all_mailbox);
/* number of waits on a neighbour that gave up (tens of seconds without an answer): non-zero = the exchange is broken */
int idfx_p2p_timeouts(IdfxBlock *blk, unsigned long long *count);
/* cudaDeviceEnablePeerAccess(device -> peer_device) for blocks of one process on several GPUs */
int idfx_p2p_enable_peer_access(int device, int peer_device);

#ifdef __cplusplus
}
#endif
#endif  /* IDEFIX_B200_H_ */
