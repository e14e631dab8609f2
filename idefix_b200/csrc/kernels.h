// kernels.h -- host-callable launchers of the CUDA kernels (internal interface
// between block.cu, which implements the C ABI, and the kernel translation units).
#ifndef IDFX_KERNELS_H_
#define IDFX_KERNELS_H_

#include "common.cuh"

// Run-time image of the compile-time configuration of the reference
struct PhysCfg {
  int ndim, ncomp, mhd, iso;
  int nvar;
  int ntr;   // passive tracers stored behind the nvar fluid variables (fluid.hpp:315-323)
  int order, solver, emf;
  // variable indices (255.. when absent)
  int RHO, MX1, MX2, MX3, ENG, BX1, BX2, BX3;
};

namespace idfxk {

// ---- sweeps (sweep_*.cu) ----------------------------------------------------
// returns cudaError_t as int; -1 when the (physics, solver, order) combination is not instantiated
int sweepHD(const PhysCfg &c, int dir, const SweepArgs &a, cudaStream_t st);
int sweepHDRoe(const PhysCfg &c, int dir, const SweepArgs &a, cudaStream_t st);
int sweepMhdTvdlf(const PhysCfg &c, int dir, const SweepArgs &a, cudaStream_t st);
int sweepMhdHll(const PhysCfg &c, int dir, const SweepArgs &a, cudaStream_t st);
int sweepMhdHlld(const PhysCfg &c, int dir, const SweepArgs &a, cudaStream_t st);
int sweepMhdHlldIso(const PhysCfg &c, int dir, const SweepArgs &a, cudaStream_t st);
int sweepMhdRoe(const PhysCfg &c, int dir, const SweepArgs &a, cudaStream_t st);
int sweepMhdRoeIso(const PhysCfg &c, int dir, const SweepArgs &a, cudaStream_t st);

// ---- constrained transport (ct.cu) ------------------------------------------
int calcCornerEmf(const DevGrid &g, const DevEmf &e, const double *Vc, const double *Vs, int emfMode, int order,
                  int ncomp, int vx1, int bx1, cudaStream_t st, long long *launches, int emfFromSweep = 0);
// mode 0: Vs += rhs (reference behaviour); 1: additionally save the old Vs into Vs0 (first stage of the
// fused path); 2: Vs <- wc*(Vs+rhs) + w0*Vs0 (later stages of the fused path)
int evolveMagField(const DevGrid &g, const DevEmf &e, double *Vs, double *Vs0, const double *dt, int haveAxis,
                   int mode, double wc, double w0, cudaStream_t st, long long *launches);

// EVOLVE_VECTOR_POTENTIAL (evolveVectorPotential.hpp): Ve += -dt E (mode as evolveMagField), Vs = curl Ve
int evolveVectorPotential(const DevGrid &g, const DevEmf &e, double *Ve, double *Ve0, const double *dt, int mode,
                          double wc, double w0, cudaStream_t st, long long *launches);
int magFieldFromA(const DevGrid &g, const double *Ve, double *Vs, int haveAxis, cudaStream_t st, long long *launches);

// ---- point-wise kernels (pointwise.cu) ---------------------------------------
struct Eos2 { double gamma, csiso; };
int primToCons(const PhysCfg &c, const DevGrid &g, const double *Vc, double *Uc, double gamma, cudaStream_t st);
int consToPrim(const PhysCfg &c, const DevGrid &g, double *Vc, double *Uc, double gamma, cudaStream_t st);
// cell-centred field from the staggered one over the whole block (boundary.hpp:389-404)
int reconstructVcField(const PhysCfg &c, const DevGrid &g, const double *Vs, double *V, cudaStream_t st,
                       int ghostsOnly = 0);
// ShockFlattening::FindShock (shockFlattening.hpp:68-178): flags[c] = 1 in cells of converging flow whose pressure
// jump exceeds smoothing * min(P) over the cell and its neighbours, on the box [beg - 1, end + 1) of every direction
int findShock(const PhysCfg &c, const DevGrid &g, const double *Vc, unsigned char *flags, double smoothing, double csiso,
              cudaStream_t st);
int fill(double *p, long long n, double value, cudaStream_t st);
int copy(double *dst, const double *src, long long n, cudaStream_t st);
// dst = wl*dst + wr*src (stateContainer.cpp:91-98)
int addAndStore(double *dst, const double *src, long long n, double wl, double wr, cudaStream_t st);
// fused stage epilogue over the interior: RK combine of Uc, Uc(B) <- Vs, Vc <- C2P(Uc)
// over the box [lo, hi) of cells (null: the whole interior)
int stageEpilogue(const PhysCfg &c, const DevGrid &g, double *Vc, double *Uc, const double *Vs, const double *Uc0,
                  double gamma, int combine, double wc, double w0, int keepUc, cudaStream_t st,
                  unsigned int *mathFlag = nullptr, const int *lo = nullptr, const int *hi = nullptr,
                  const double *invDt = nullptr, unsigned long long *dtMinBits = nullptr);
// invDt / dtMinBits non-null: also reduce min(1 / InvDt) over the box into *dtMinBits (ordered bits), i.e.
// DataBlock::ComputeTimestep folded into the epilogue of the first stage

// ---- passive tracers (tracer.cu) ------------------------------------------------
// Tracer::CalcFlux + CalcRightHandSide of one direction from the mass flux the sweep stored
int tracerSweep(const DevGrid &g, int dir, const double *Vtr, double *Utr, const double *fluxRho, const double *dt,
                int ntr, cudaStream_t st);
// Tracer::ConvertPrimToCons over the whole block; Uc0 != null also saves the result as the "begin" state
int tracerPrimToCons(const PhysCfg &c, const DevGrid &g, const double *Vc, double *Uc, double *Uc0, cudaStream_t st);
// Tracer::ConvertConsToPrim; combine: u <- wc*u + w0*u0 first (fused path)
int tracerConsToPrim(const PhysCfg &c, const DevGrid &g, double *Vc, double *Uc, const double *Uc0, int combine,
                     double wc, double w0, int interiorOnly, cudaStream_t st);

// ---- boundaries (boundary.cu) -------------------------------------------------
int enforcePhysicalBoundary(const PhysCfg &c, const DevGrid &g, double *Vc, double *Vs, int dir, int side, int type,
                            cudaStream_t st, long long *launches);
int reconstructNormalField(const PhysCfg &c, const DevGrid &g, double *Vs, int dir, int left, int right,
                           cudaStream_t st, long long *launches, int signLeft = 1, int signRight = 1);
// polar axis of spherical grids (axis.cpp): ghost cells across the axis, B_theta on the axis, EMFs on the axis
int enforceAxisBoundary(const PhysCfg &c, const DevGrid &g, double *Vc, double *Vs, int side, int twoPi, cudaStream_t st,
                        long long *launches);
int axisRegularizeBx2s(const DevGrid &g, double *Vs, int twoPi, int left, int right, cudaStream_t st, long long *launches);
int axisRegularizeEmf(const DevGrid &g, double *ex, double *ez, int twoPi, int left, int right, cudaStream_t st,
                      long long *launches);
// halo pack / unpack (mpi.hpp:49-172, mpi.cpp:79-134)
long long haloBufferSize(const PhysCfg &c, const DevGrid &g, int dir);
int haloPack(const PhysCfg &c, const DevGrid &g, const double *Vc, const double *Vs, int dir, int side, double *buf,
             cudaStream_t st, long long *launches);
int haloUnpack(const PhysCfg &c, const DevGrid &g, double *Vc, double *Vs, int dir, int side, const double *buf,
               cudaStream_t st, long long *launches);

// ---- peer-to-peer halo exchange and reduction through device mailboxes (boundary.cu) ------------------------------
constexpr int P2P_MAX_RANKS = 64;
// mailbox of a block (64-bit words, written by its neighbours / by every block of the domain):
//   DATA[2 dir + side]  number of the last message that landed in the receive buffer of ghost side `side`
//   ACK [2 dir + side]  number of the last message of OUR side-`side` slab the neighbour has unpacked
//   REDUCE[4 rank + 2 parity + {0: value bits, 1: reduction number}]
enum : int { P2P_MB_DATA = 0, P2P_MB_ACK = 6, P2P_MB_REDUCE = 12, P2P_MB_WORDS = 12 + 4 * P2P_MAX_RANKS };
// private counters of a block (device memory, advanced by the kernels: graph replays stay valid)
enum : int { P2P_SEQ_SEND = 0, P2P_SEQ_RECV = 6, P2P_SEQ_REDUCE = 12, P2P_SEQ_TIMEOUTS = 13, P2P_SEQ_WORDS = 16 };
struct P2PChannel {
  int dir;
  int active[2];                       // side exchanges with another block
  double *peerRecv[2];                 // neighbour's receive buffer our side-s slab is stored into
  const double *localRecv[2];          // our receive buffer of ghost side s
  unsigned long long *mailbox;         // ours
  unsigned long long *peerMailbox[2];  // the neighbour's on side s
  unsigned long long *seq;             // our counters
  unsigned int *ticket;                // [2] last-CTA-done tickets (send, receive)
};
struct P2PReduce {
  int rank, world;
  unsigned long long *mailbox, *seq;
  unsigned long long *allMailbox[P2P_MAX_RANKS];
};
// one kernel: both sides, Vc + every Vs component; send = true stores into the neighbours, false unpacks
int haloP2P(const PhysCfg &c, const DevGrid &g, double *Vc, double *Vs, bool send, const P2PChannel &ch, cudaStream_t st,
            long long *launches);
int p2pPreload();
// in-place min (op 0) / max (op 1) of *devValue over all blocks
int p2pAllReduce(const P2PReduce &r, double *devValue, int op, cudaStream_t st);

// ---- reductions (reduce.cu) ----------------------------------------------------
// min over the interior of 1/InvDt, written as ordered bits to *outBits (must be preset to +inf bits)
int timestepReduce(const DevGrid &g, const double *invDt, unsigned long long *outBits, cudaStream_t st);
// device-side dt control (timeIntegrator.cpp:313-323,380-408); state: [0]=t [1]=dt [2]=t0 [3]=tstop
int dtControlBegin(double *state, unsigned long long *dtMinBits, cudaStream_t st);
int dtControlEnd(double *state, unsigned long long *dtMinBits, double cfl, double cflMaxVar, double fixedDt,
                 long long *ncycles, int *status, const unsigned int *mathFlag, cudaStream_t st);
int stageTimeAdvance(double *state, int stage, double wc, double w0, cudaStream_t st);
int checkDivB(const PhysCfg &c, const DevGrid &g, const double *Vs, unsigned long long *outBits, cudaStream_t st);
int checkNan(const PhysCfg &c, const DevGrid &g, const double *Vc, const double *Vs, int *count, cudaStream_t st);

}  // namespace idfxk

#endif  // IDFX_KERNELS_H_
