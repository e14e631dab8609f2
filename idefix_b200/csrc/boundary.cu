// boundary.cu -- physical boundary conditions, normal-field reconstruction in the
// ghost zones and the halo pack / unpack kernels of the multi-GPU exchange.
//
// Reference: src/fluid/boundary/boundary.hpp:577-809 (periodic / reflective / outflow),
// :408-530 (ReconstructNormalField), :969-1100 (loop bounds of BoundaryForAll/X1s/X2s/X3s);
// src/mpi.cpp:247-846 and src/mpi.hpp:49-172 (which slabs travel between neighbours).
#include "kernels.h"

namespace {

constexpr int TBX = 64, TBY = 4;

struct Box {
  int lo[3], hi[3];
};
inline dim3 boxGrid(const Box &b, int nz = 1, dim3 block = dim3(TBX, TBY, 1)) {
  int n0 = b.hi[0] - b.lo[0], n1 = b.hi[1] - b.lo[1], n2 = b.hi[2] - b.lo[2];
  if (n0 < 1) n0 = 1;
  if (n1 < 1) n1 = 1;
  if (n2 < 1) n2 = 1;
  return dim3((n0 + block.x - 1) / block.x, (n1 + block.y - 1) / block.y, n2 * nz);
}
// thread block of the ghost-fill kernel: an x1 slab is only nghost cells wide along i, so the threads
// go along j there (a 64 x 4 block would idle 60 of its 64 columns)
inline dim3 slabBlock(const Box &b) {
  return (b.hi[0] - b.lo[0] <= 4) ? dim3(4, 64, 1) : dim3(TBX, TBY, 1);
}
inline bool emptyBox(const Box &b) {
  return b.hi[0] <= b.lo[0] || b.hi[1] <= b.lo[1] || b.hi[2] <= b.lo[2];
}

// ghost slab of direction dir / side for an array whose active extent along `stag` is
// one larger (stag = -1: cell centred).  boundary.hpp:969-1100.
Box ghostSlab(const DevGrid &g, int dir, int side, int stag) {
  Box b;
  for (int d = 0; d < 3; d++) {
    const int nint = (g.end[d] - g.beg[d]) + (d == stag ? 1 : 0);
    const int ntot = g.nx[d] + (d == stag ? 1 : 0);
    if (d == dir) {
      b.lo[d] = side * (g.ng[d] + nint);
      b.hi[d] = g.ng[d] + side * (g.ng[d] + nint);
    } else {
      b.lo[d] = 0;
      b.hi[d] = ntot;
    }
  }
  return b;
}

// One kernel serves the three standard boundary types for any array (cell-centred variable
// or staggered component): out(k,j,i) = sign * in(ref), with the reference index along dir.
//   type periodic  : ref = ng + (x + ng*(n-1)) % n                 (boundary.hpp:591-603)
//   type reflective: ref = 2*(ng + side*n) - x - 1                  (boundary.hpp:693-699)
//   type outflow   : ref = ng + side*(n-1)                          (boundary.hpp:757-769)
// n is the number of ACTIVE CELLS (np_int) along dir in all cases, as in the reference.
__global__ void ghostFillKernel(double *__restrict__ q, long long sj, long long sk, Box b, int nvars,
                                long long sn, int dir, int side, int type, int ng, int nint, int signMask,
                                int zeroInflowVar) {
  const int i = b.lo[0] + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = b.lo[1] + blockIdx.y * blockDim.y + threadIdx.y;
  const int nk = b.hi[2] - b.lo[2];
  const int k = b.lo[2] + (int)(blockIdx.z % nk);
  const int n = (int)(blockIdx.z / nk);
  if (i >= b.hi[0] || j >= b.hi[1] || n >= nvars) return;
  int x = (dir == 0) ? i : (dir == 1 ? j : k);
  int ref;
  if (type == BC_PERIODIC) {
    ref = ng + (x + ng * (nint - 1)) % nint;
  } else if (type == BC_REFLECTIVE) {
    ref = 2 * (ng + side * nint) - x - 1;
  } else {
    ref = ng + side * (nint - 1);
  }
  const int ir = (dir == 0) ? ref : i, jr = (dir == 1) ? ref : j, kr = (dir == 2) ? ref : k;
  const long long dst = n * sn + (long long)k * sk + (long long)j * sj + i;
  const long long src = n * sn + (long long)kr * sk + (long long)jr * sj + ir;
  double v = q[src];
  if (type == BC_REFLECTIVE) {
    if ((signMask >> n) & 1) v = -1.0 * v;
    else v = 1.0 * v;
  } else if (type == BC_OUTFLOW) {
    // no inflow through an outflow boundary (boundary.hpp:763-767)
    const int sign = 1 - 2 * side;
    if (n == zeroInflowVar && (sign * v >= 0.0)) v = 0.0;
  }
  q[dst] = v;
}

// boundary.hpp:434-527: march outwards from the last active face so that div B = 0 in the ghosts
__global__ void reconstructNormalKernel(DevGrid g, double *__restrict__ Vs, int dir, int left, int right,
                                        int signLeft, int signRight) {
  // thread owns a line along dir; (a,b) index the two perpendicular directions
  const int p0 = (dir == 0) ? 1 : 0;
  const int p1 = (dir == 2) ? 1 : 2;
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y * blockDim.y + threadIdx.y;
  if (a >= g.nx[p0] || b >= g.nx[p1]) return;
  const int nd = g.ndim;
  double *B1 = Vs, *B2 = Vs + g.vsn, *B3 = Vs + 2 * g.vsn;
  auto S = [&](int k, int j, int i) { return sidx(g, k, j, i); };
  auto div2 = [&](int k, int j, int i, int skip) -> double {
    // sum of the perpendicular flux differences, in the reference's order
    double s = 0.0;
    if (skip == 0) {
      if (nd >= 2) s = s + faceAreaDir(g, 1, k, j + 1, i) * B2[S(k, j + 1, i)] - faceAreaDir(g, 1, k, j, i) * B2[S(k, j, i)];
      if (nd == 3) s = s + faceAreaDir(g, 2, k + 1, j, i) * B3[S(k + 1, j, i)] - faceAreaDir(g, 2, k, j, i) * B3[S(k, j, i)];
    }
    return s;
  };
  if (dir == 0) {
    const int j = a, k = b;
    if (left) {
      for (int i = g.beg[0] - 1; i >= 0; i--) {
        B1[S(k, j, i)] = 1.0 / faceAreaDir(g, 0, k, j, i) *
                         (faceAreaDir(g, 0, k, j, i + 1) * B1[S(k, j, i + 1)] + (div2(k, j, i, 0)));
      }
    }
    if (right) {
      for (int i = g.end[0]; i < g.nx[0]; i++) {
        B1[S(k, j, i + 1)] = 1.0 / faceAreaDir(g, 0, k, j, i + 1) *
                             (faceAreaDir(g, 0, k, j, i) * B1[S(k, j, i)] - (div2(k, j, i, 0)));
      }
    }
  } else if (dir == 1) {
    const int i = a, k = b;
    if (left) {
      for (int j = g.beg[1] - 1; j >= 0; j--) {
        double s = faceAreaDir(g, 0, k, j, i + 1) * B1[S(k, j, i + 1)] - faceAreaDir(g, 0, k, j, i) * B1[S(k, j, i)];
        if (nd == 3)
          s = s + signLeft * (faceAreaDir(g, 2, k + 1, j, i) * B3[S(k + 1, j, i)] -
                              faceAreaDir(g, 2, k, j, i) * B3[S(k, j, i)]);
        B2[S(k, j, i)] = 1.0 / faceAreaDir(g, 1, k, j, i) * (faceAreaDir(g, 1, k, j + 1, i) * B2[S(k, j + 1, i)] + (s));
      }
    }
    if (right) {
      for (int j = g.end[1]; j < g.nx[1]; j++) {
        double s = faceAreaDir(g, 0, k, j, i + 1) * B1[S(k, j, i + 1)] - faceAreaDir(g, 0, k, j, i) * B1[S(k, j, i)];
        if (nd == 3)
          s = s + signRight * (faceAreaDir(g, 2, k + 1, j, i) * B3[S(k + 1, j, i)] -
                               faceAreaDir(g, 2, k, j, i) * B3[S(k, j, i)]);
        B2[S(k, j + 1, i)] =
            1.0 / faceAreaDir(g, 1, k, j + 1, i) * (faceAreaDir(g, 1, k, j, i) * B2[S(k, j, i)] - (s));
      }
    }
  } else {
    const int i = a, j = b;
    if (left) {
      for (int k = g.beg[2] - 1; k >= 0; k--) {
        const double s = faceAreaDir(g, 0, k, j, i + 1) * B1[S(k, j, i + 1)] - faceAreaDir(g, 0, k, j, i) * B1[S(k, j, i)] +
                         faceAreaDir(g, 1, k, j + 1, i) * B2[S(k, j + 1, i)] - faceAreaDir(g, 1, k, j, i) * B2[S(k, j, i)];
        B3[S(k, j, i)] = 1.0 / faceAreaDir(g, 2, k, j, i) * (faceAreaDir(g, 2, k + 1, j, i) * B3[S(k + 1, j, i)] + (s));
      }
    }
    if (right) {
      for (int k = g.end[2]; k < g.nx[2]; k++) {
        const double s = faceAreaDir(g, 0, k, j, i + 1) * B1[S(k, j, i + 1)] - faceAreaDir(g, 0, k, j, i) * B1[S(k, j, i)] +
                         faceAreaDir(g, 1, k, j + 1, i) * B2[S(k, j + 1, i)] - faceAreaDir(g, 1, k, j, i) * B2[S(k, j, i)];
        B3[S(k + 1, j, i)] =
            1.0 / faceAreaDir(g, 2, k + 1, j, i) * (faceAreaDir(g, 2, k, j, i) * B3[S(k, j, i)] - (s));
      }
    }
  }
}

// ---- polar axis of spherical grids (src/fluid/boundary/axis.cpp) ---------------------------------------------
// Axis::EnforceAxisBoundary (axis.cpp:279-357): ghost cells across the axis come from the opposite meridian (full
// 2 pi: k shifted by half the active cells) or from the same one (a fraction of 2 pi), mirrored in theta, with the
// sign of the theta / phi components flipped (symmetryVc / symmetryVs, axis.hpp:128-162).
__global__ void axisFillKernel(double *__restrict__ q, long long sj, long long sk, long long sn, Box b, int nvars,
                               int jref, int offset, int signMask, int twoPi, int ngk, int nkint) {
  const int i = b.lo[0] + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = b.lo[1] + blockIdx.y * blockDim.y + threadIdx.y;
  const int nk = b.hi[2] - b.lo[2];
  const int k = b.lo[2] + (int)(blockIdx.z % nk);
  const int n = (int)(blockIdx.z / nk);
  if (i >= b.hi[0] || j >= b.hi[1] || n >= nvars) return;
  const int kcomp = twoPi ? ngk + ((k - ngk + nkint / 2) % nkint) : k;
  const long long dst = n * sn + (long long)k * sk + (long long)j * sj + i;
  const long long src = n * sn + (long long)kcomp * sk + (long long)(2 * jref - j + offset) * sj + i;
  const double v = q[src];
  q[dst] = ((signMask >> n) & 1) ? -1.0 * v : 1.0 * v;
}

// Axis::RegularizeBX2s (axis.cpp:362-395).  Full 2 pi: the reference is built with
// AXIS_BX2S_USE_ATHENA_REGULARISATION (axis.hpp:21), i.e. FixBx2sAxisGhostAverage (axis.cpp:244-272): the value on
// the axis is the mean of its two neighbours in theta, 3-D only, active k.  A fraction of 2 pi: zero on the axis.
__global__ void axisBx2sKernel(DevGrid g, double *__restrict__ Vs, int twoPi, int left, int right) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= g.nx[0] || k >= g.nx[2]) return;
  double *B2 = Vs + g.vsn;
  for (int side = 0; side < 2; side++) {
    if (!(side == 0 ? left : right)) continue;
    const int jaxis = side == 0 ? g.beg[1] : g.end[1];
    if (twoPi) {
      if (g.ndim == 3 && k >= g.beg[2] && k < g.end[2])
        B2[sidx(g, k, jaxis, i)] = 0.5 * (B2[sidx(g, k, jaxis - 1, i)] + B2[sidx(g, k, jaxis + 1, i)]);
    } else {
      B2[sidx(g, k, jaxis, i)] = 0.0;
    }
  }
}

// Axis::RegularizeEMFs (axis.cpp:26-84,160-177) on one side: the radial EMF on the axis is its mean around the pole
// (full 2 pi, 3-D) or zero (fraction of 2 pi), the azimuthal EMF on the axis is zero.  One thread per radial index
// adds the active k in ascending order, which is the order of the reference's atomic accumulation when it runs on
// one thread (measured: descending or pairwise orders differ from it in the last bit; with several OpenMP threads
// the reference itself is not reproducible here).
__global__ void axisEmfKernel(DevGrid g, double *__restrict__ ex, double *__restrict__ ez, int jref, int twoPi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.nx[0]) return;
  if (g.ndim == 3) {
    double value = 0.0;
    if (twoPi) {
      double sum = 0.0;
      for (int k = g.beg[2]; k < g.end[2]; k++) sum += ex[cidx(g, k, jref, i)];
      value = sum / ((double)(g.end[2] - g.beg[2]));
    }
    for (int k = 0; k < g.nx[2]; k++) ex[cidx(g, k, jref, i)] = value;
  }
  for (int k = 0; k < g.nx[2]; k++) ez[cidx(g, k, jref, i)] = 0.0;
}

// ---- halo pack / unpack ----------------------------------------------------------------------
struct VarList {
  int n;
  int v[16];
};

template <bool PACK>
__global__ void haloCopyKernel(double *__restrict__ q, long long sj, long long sk, long long sn, VarList vars, Box b,
                               double *__restrict__ buf) {
  const int n0 = b.hi[0] - b.lo[0], n1 = b.hi[1] - b.lo[1], n2 = b.hi[2] - b.lo[2];
  const int ii = blockIdx.x * blockDim.x + threadIdx.x;
  const int jj = blockIdx.y * blockDim.y + threadIdx.y;
  const int kk = (int)(blockIdx.z % n2);
  const int n = (int)(blockIdx.z / n2);
  if (ii >= n0 || jj >= n1 || n >= vars.n) return;
  const long long a = vars.v[n] * sn + (long long)(b.lo[2] + kk) * sk + (long long)(b.lo[1] + jj) * sj + (b.lo[0] + ii);
  const long long p = (((long long)n * n2 + kk) * n1 + jj) * n0 + ii;
  if (PACK) buf[p] = q[a];
  else q[a] = buf[p];
}

inline long long boxCount(const Box &b) {
  if (emptyBox(b)) return 0;
  return (long long)(b.hi[0] - b.lo[0]) * (b.hi[1] - b.lo[1]) * (b.hi[2] - b.lo[2]);
}

// variables of Vc that travel: everything except the CT-evolved field components (boundary.hpp:164-186)
VarList haloVars(const PhysCfg &c) {
  VarList l;
  l.n = 0;
  for (int n = 0; n < c.nvar + c.ntr; n++) {
    if (c.mhd && n >= c.BX1 && n < c.BX1 + c.ndim) continue;
    l.v[l.n++] = n;
  }
  return l;
}

// box of cells/faces sent (pack) or received (unpack) in direction dir on `side`
// comp = -1 for cell-centred data, 0..2 for the staggered component.  mpi.cpp:269-320,469-520,666-716
Box haloBox(const DevGrid &g, int dir, int side, int comp, bool pack) {
  Box b;
  for (int d = 0; d < 3; d++) {
    if (d == dir) {
      const int ng = g.ng[d];
      const int shift = (comp == d) ? 1 : 0;  // the shared face is not exchanged
      if (pack) {
        if (side == 0) { b.lo[d] = ng + shift; b.hi[d] = 2 * ng + shift; }
        else           { b.lo[d] = g.end[d] - ng; b.hi[d] = g.end[d]; }
      } else {
        if (side == 0) { b.lo[d] = 0; b.hi[d] = ng; }
        else           { b.lo[d] = g.end[d] + shift; b.hi[d] = g.end[d] + ng + shift; }
      }
    } else if (d < dir) {
      // directions already exchanged: whole extent, ghosts included (corners ride along)
      b.lo[d] = 0;
      b.hi[d] = g.nx[d] + ((comp == d) ? 1 : 0);
    } else {
      b.lo[d] = g.beg[d];
      b.hi[d] = g.end[d] + ((comp == d && d < g.ndim) ? 1 : 0);
    }
  }
  return b;
}

template <bool PACK>
int haloTransfer(const PhysCfg &c, const DevGrid &g, double *Vc, double *Vs, int dir, int side, double *buf,
                 cudaStream_t st, long long *launches) {
  const dim3 block(TBX, TBY, 1);
  long long offset = 0;
  {
    const VarList vars = haloVars(c);
    const Box b = haloBox(g, dir, side, -1, PACK);
    if (!emptyBox(b) && vars.n > 0) {
      const dim3 blk = slabBlock(b);   // x1 slabs are only nghost cells wide along i
      haloCopyKernel<PACK><<<boxGrid(b, vars.n, blk), blk, 0, st>>>(Vc, g.sj, g.sk, g.sn, vars, b, buf + offset);
      (*launches)++;
      offset += boxCount(b) * vars.n;
    }
  }
  if (c.mhd) {
    for (int comp = 0; comp < g.ndim; comp++) {
      VarList one;
      one.n = 1;
      one.v[0] = comp;
      const Box b = haloBox(g, dir, side, comp, PACK);
      if (emptyBox(b)) continue;
      const dim3 blk = slabBlock(b);
      haloCopyKernel<PACK><<<boxGrid(b, 1, blk), blk, 0, st>>>(Vs, g.vsj, g.vsk, g.vsn, one, b, buf + offset);
      (*launches)++;
      offset += boxCount(b);
    }
  }
  return (int)cudaGetLastError();
}

}  // namespace

long long idfxk::haloBufferSize(const PhysCfg &c, const DevGrid &g, int dir) {
  long long n = boxCount(haloBox(g, dir, 0, -1, true)) * haloVars(c).n;
  if (c.mhd)
    for (int comp = 0; comp < g.ndim; comp++) n += boxCount(haloBox(g, dir, 0, comp, true));
  return n;
}

int idfxk::haloPack(const PhysCfg &c, const DevGrid &g, const double *Vc, const double *Vs, int dir, int side,
                    double *buf, cudaStream_t st, long long *launches) {
  return haloTransfer<true>(c, g, const_cast<double *>(Vc), const_cast<double *>(Vs), dir, side, buf, st, launches);
}
int idfxk::haloUnpack(const PhysCfg &c, const DevGrid &g, double *Vc, double *Vs, int dir, int side,
                      const double *buf, cudaStream_t st, long long *launches) {
  return haloTransfer<false>(c, g, Vc, Vs, dir, side, const_cast<double *>(buf), st, launches);
}

// =================================================================================================
// Peer-to-peer halo exchange (replaces Mpi::ExchangeX1/X2/X3, src/mpi.cpp:247-846, on one NVLink/NVSwitch box).
//
// One SEND kernel and one RECEIVE kernel per direction, for both sides and for Vc + every Vs component at once:
//   send:    wait until the neighbours have consumed the previous message of each channel (ack flags in OUR mailbox),
//            store the slabs straight into the neighbours' receive buffers over NVLink (the reference's buffer
//            layout, mpi.cpp:79-134), then the last CTA to finish publishes the message number in the NEIGHBOURS'
//            mailboxes (system-scope release);
//   receive: spin on the message numbers in OUR mailbox (system-scope acquire), copy the received slabs into the
//            ghost zones, then the last CTA acknowledges in the neighbours' mailboxes.
// Message numbers live in device memory and are advanced by the kernels themselves, so the pair is identical launch
// for launch and a whole RK cycle including its exchanges can be captured in a CUDA graph; nothing synchronises with
// the host.  A neighbour that lags simply keeps the spinning CTAs waiting; since every block runs the same
// send -> receive sequence, a wait is always for something the neighbour does without waiting for us (no deadlock).
// =================================================================================================
struct HaloSeg {
  double *arr;            // Vc or Vs
  long long sj, sk, sn;   // strides of arr
  VarList vars;
  Box box;                // cells / faces that travel
  long long offset;       // first element of this segment in the flat index space of the plan
  long long count;        // elements of this segment
  long long bufOffset;    // first element of this segment inside its side's message buffer
  int side;               // 0: left, 1: right
  __host__ __device__ long long sideOffset() const { return bufOffset; }
};
struct HaloPlan {
  int nseg;
  long long total;
  HaloSeg seg[8];
};

__device__ __forceinline__ unsigned long long loadAcquireSys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void storeReleaseSys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// spin until *flag >= want.  A neighbour that never answers (crashed process, blocks of one host thread driven in
// the wrong order) must not hang the device for ever: after ~2^27 polls (tens of seconds) the wait gives up and
// counts a time-out in the block's counters (idfx_p2p_timeouts), which the host turns into an error.
__device__ __forceinline__ void waitAtLeast(const unsigned long long *flag, unsigned long long want,
                                            unsigned long long *seq) {
  unsigned int polls = 0;
  while (loadAcquireSys(flag) < want) {
    __nanosleep(polls < 64 ? 32 : 256);
    if (++polls > (1u << 27)) {
      atomicAdd(seq + idfxk::P2P_SEQ_TIMEOUTS, 1ull);
      break;
    }
  }
}

template <bool SEND>
__global__ void __launch_bounds__(256) haloP2PKernel(HaloPlan plan, idfxk::P2PChannel ch) {
  using namespace idfxk;
  // ---- wait: acknowledgements of everything sent so far / arrival of the next message ----------------------------
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; s++) {
      if (!ch.active[s]) continue;
      if (SEND) {
        const unsigned long long sent = ch.seq[P2P_SEQ_SEND + 2 * ch.dir + s];
        waitAtLeast(ch.mailbox + P2P_MB_ACK + 2 * ch.dir + s, sent, ch.seq);
      } else {
        const unsigned long long want = ch.seq[P2P_SEQ_RECV + 2 * ch.dir + s] + 1;
        waitAtLeast(ch.mailbox + P2P_MB_DATA + 2 * ch.dir + s, want, ch.seq);
      }
    }
  }
  __syncthreads();
  // ---- copy ---------------------------------------------------------------------------------------------------
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < plan.total; e += (long long)gridDim.x * blockDim.x) {
    int sg = 0;
    while (sg + 1 < plan.nseg && e >= plan.seg[sg].offset + plan.seg[sg].count) sg++;
    const HaloSeg &g = plan.seg[sg];
    const long long l = e - g.offset;
    const int n0 = g.box.hi[0] - g.box.lo[0], n1 = g.box.hi[1] - g.box.lo[1], n2 = g.box.hi[2] - g.box.lo[2];
    const int ii = (int)(l % n0);
    const int jj = (int)((l / n0) % n1);
    const int kk = (int)((l / ((long long)n0 * n1)) % n2);
    const int n = (int)(l / ((long long)n0 * n1 * n2));
    const long long a = g.vars.v[n] * g.sn + (long long)(g.box.lo[2] + kk) * g.sk + (long long)(g.box.lo[1] + jj) * g.sj +
                        (g.box.lo[0] + ii);
    // position inside the side's own message buffer
    const long long pos = l + g.sideOffset();
    if (SEND) ch.peerRecv[g.side][pos] = g.arr[a];
    else g.arr[a] = ch.localRecv[g.side][pos];
  }
  // ---- publish: the last CTA to finish advances the message numbers and raises the neighbours' flags -----------------
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int *ticket = ch.ticket + (SEND ? 0 : 1);
    const unsigned int t = atomicAdd(ticket, 1u);
    if (t == gridDim.x - 1) {
      *ticket = 0u;
      __threadfence_system();
      for (int s = 0; s < 2; s++) {
        if (!ch.active[s]) continue;
        unsigned long long *counter = ch.seq + (SEND ? P2P_SEQ_SEND : P2P_SEQ_RECV) + 2 * ch.dir + s;
        const unsigned long long m = *counter + 1;
        *counter = m;
        // our side-s slab fills the neighbour's ghost side 1-s; its data / ack flags are indexed by ITS side
        storeReleaseSys(ch.peerMailbox[s] + (SEND ? P2P_MB_DATA : P2P_MB_ACK) + 2 * ch.dir + (1 - s), m);
      }
    }
  }
}

// All-reduce (min / max) of one non-negative double over the blocks of the domain through the mailboxes: every
// block stores its value into slot [rank] of every mailbox, then reads all slots of its own.  Replaces
// MPI_Iallreduce(MIN) of timeIntegrator.cpp:318 (dt) and the MPI_MAX of checkDivB.hpp.  One CTA, one thread per block
// of the domain; slots are double-buffered by the parity of the reduction number (a block cannot be two reductions
// ahead of another one, since finishing a reduction needs everybody's contribution to it).
__global__ void p2pAllReduceKernel(idfxk::P2PReduce r, double *value, int op) {
  using namespace idfxk;
  __shared__ double vals[P2P_MAX_RANKS];
  const int t = threadIdx.x;
  const unsigned long long m = r.seq[P2P_SEQ_REDUCE] + 1;
  const int par = (int)(m & 1ull);
  if (t < r.world) {
    unsigned long long *dst = r.allMailbox[t] + P2P_MB_REDUCE + 4 * r.rank + 2 * par;
    const unsigned long long bits = (unsigned long long)__double_as_longlong(*value);
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(bits) : "memory");
    __threadfence_system();
    storeReleaseSys(dst + 1, m);
    const unsigned long long *src = r.mailbox + P2P_MB_REDUCE + 4 * t + 2 * par;
    waitAtLeast(src + 1, m, r.seq);
    vals[t] = __longlong_as_double((long long)loadAcquireSys(src));
  }
  __syncthreads();
  if (t == 0) {
    double v = vals[0];
    for (int q = 1; q < r.world; q++) v = (op == 0) ? fmin(v, vals[q]) : fmax(v, vals[q]);
    *value = v;
    r.seq[P2P_SEQ_REDUCE] = m;
  }
}

int idfxk::haloP2P(const PhysCfg &c, const DevGrid &g, double *Vc, double *Vs, bool send, const P2PChannel &ch,
                   cudaStream_t st, long long *launches) {
  HaloPlan plan;
  plan.nseg = 0;
  plan.total = 0;
  for (int side = 0; side < 2; side++) {
    if (!ch.active[side]) continue;
    long long bufOffset = 0;   // same layout inside a side's buffer as haloTransfer (mpi.cpp:79-134)
    auto add = [&](double *arr, long long sj, long long sk, long long sn, const VarList &vars, const Box &b) {
      if (emptyBox(b) || vars.n == 0) return;
      HaloSeg &sg = plan.seg[plan.nseg++];
      sg.arr = arr;
      sg.sj = sj;
      sg.sk = sk;
      sg.sn = sn;
      sg.vars = vars;
      sg.box = b;
      sg.side = side;
      sg.count = boxCount(b) * vars.n;
      sg.offset = plan.total;
      sg.bufOffset = bufOffset;
      plan.total += sg.count;
      bufOffset += sg.count;
    };
    add(Vc, g.sj, g.sk, g.sn, haloVars(c), haloBox(g, ch.dir, side, -1, send));
    if (c.mhd) {
      for (int comp = 0; comp < g.ndim; comp++) {
        VarList one;
        one.n = 1;
        one.v[0] = comp;
        add(Vs, g.vsj, g.vsk, g.vsn, one, haloBox(g, ch.dir, side, comp, send));
      }
    }
  }
  if (plan.nseg == 0) return 0;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // few CTAs per SM: a receive kernel may spin while the neighbour is still busy, and must leave room for whatever
  // else shares the device (a second block of the same process, the other stream of an overlapped exchange)
  long long blocks = (plan.total + 2047) / 2048;
  if (blocks > 2LL * sms) blocks = 2LL * sms;
  if (blocks < 1) blocks = 1;
  if (send) haloP2PKernel<true><<<(unsigned)blocks, 256, 0, st>>>(plan, ch);
  else haloP2PKernel<false><<<(unsigned)blocks, 256, 0, st>>>(plan, ch);
  (*launches)++;
  return (int)cudaGetLastError();
}

// Forces the (lazily loaded) peer-to-peer kernels into the context now: loading a kernel for its first launch may
// need the device to drain, which never happens while a receive kernel is waiting for a neighbour whose work the same
// host thread has not enqueued yet.
int idfxk::p2pPreload() {
  cudaFuncAttributes at;
  cudaError_t e = cudaFuncGetAttributes(&at, haloP2PKernel<true>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&at, haloP2PKernel<false>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&at, p2pAllReduceKernel);
  return (int)e;
}

int idfxk::p2pAllReduce(const P2PReduce &r, double *devValue, int op, cudaStream_t st) {
  p2pAllReduceKernel<<<1, P2P_MAX_RANKS, 0, st>>>(r, devValue, op);
  return (int)cudaGetLastError();
}

// Boundary::EnforcePeriodic / EnforceReflective / EnforceOutflow for one side
int idfxk::enforcePhysicalBoundary(const PhysCfg &c, const DevGrid &g, double *Vc, double *Vs, int dir, int side,
                                   int type, cudaStream_t st, long long *launches) {
  const dim3 block(TBX, TBY, 1);
  const int nint = g.end[dir] - g.beg[dir];
  const int ng = g.ng[dir];
  {
    // cell-centred variables (BoundaryForAll)
    const Box b = ghostSlab(g, dir, side, -1);
    int signMask = 0;
    if (type == BC_REFLECTIVE) {
      // normal velocity and tangential field components change sign (boundary.hpp:697)
      for (int n = 0; n < c.nvar; n++) {
        const bool flip = (n == c.MX1 + dir) || (c.mhd && n >= c.BX1 && n < c.BX1 + c.ncomp && n != c.BX1 + dir);
        if (flip) signMask |= (1 << n);
      }
    }
    const dim3 blk = slabBlock(b);
    const int nall = c.nvar + c.ntr;   // tracers behave like scalars (boundary.hpp:136)
    ghostFillKernel<<<boxGrid(b, nall, blk), blk, 0, st>>>(Vc, g.sj, g.sk, b, nall, g.sn, dir, side, type, ng, nint,
                                                           signMask, c.MX1 + dir);
    (*launches)++;
  }
  if (c.mhd) {
    for (int comp = 0; comp < g.ndim; comp++) {
      // reflective / outflow leave the normal component to ReconstructNormalField (boundary.hpp:702-736)
      if (type != BC_PERIODIC && comp == dir) continue;
      const Box b = ghostSlab(g, dir, side, comp);
      const int signMask = (type == BC_REFLECTIVE) ? 1 : 0;
      const dim3 blk = slabBlock(b);
      ghostFillKernel<<<boxGrid(b, 1, blk), blk, 0, st>>>(Vs + comp * g.vsn, g.vsj, g.vsk, b, 1, 0, dir, side, type, ng,
                                                          nint, signMask, -1);
      (*launches)++;
    }
  }
  return (int)cudaGetLastError();
}

int idfxk::reconstructNormalField(const PhysCfg &c, const DevGrid &g, double *Vs, int dir, int left, int right,
                                  cudaStream_t st, long long *launches, int signLeft, int signRight) {
  if (!c.mhd || (!left && !right)) return 0;
  const int p0 = (dir == 0) ? 1 : 0;
  const int p1 = (dir == 2) ? 1 : 2;
  const dim3 block(32, 4, 1);
  const dim3 grid((g.nx[p0] + 31) / 32, (g.nx[p1] + 3) / 4, 1);
  reconstructNormalKernel<<<grid, block, 0, st>>>(g, Vs, dir, left, right, signLeft, signRight);
  (*launches)++;
  return (int)cudaGetLastError();
}

// Axis::EnforceAxisBoundary for one side of x2
int idfxk::enforceAxisBoundary(const PhysCfg &c, const DevGrid &g, double *Vc, double *Vs, int side, int twoPi,
                               cudaStream_t st, long long *launches) {
  const int jref = side == 0 ? g.beg[1] : g.end[1] - 1;
  const int offset = side == 0 ? -1 : 1;
  const int ngk = g.ng[2], nkint = g.end[2] - g.beg[2];
  {
    const Box b = ghostSlab(g, 1, side, -1);
    int signMask = 0;   // VX2, VX3, BX2, BX3 change sign across the axis (axis.hpp:133-144); tracers do not
    if (c.ncomp >= 2) signMask |= 1 << (c.MX1 + 1);
    if (c.ncomp >= 3) signMask |= 1 << (c.MX1 + 2);
    if (c.mhd && c.ncomp >= 2) signMask |= 1 << (c.BX1 + 1);
    if (c.mhd && c.ncomp >= 3) signMask |= 1 << (c.BX1 + 2);
    const int nall = c.nvar + c.ntr;
    const dim3 blk = slabBlock(b);
    axisFillKernel<<<boxGrid(b, nall, blk), blk, 0, st>>>(Vc, g.sj, g.sk, g.sn, b, nall, jref, offset, signMask, twoPi,
                                                          ngk, nkint);
    (*launches)++;
  }
  if (c.mhd) {
    for (int comp = 0; comp < g.ndim; comp++) {
      if (comp == 1) continue;   // the normal component is reconstructed from div B = 0 (axis.cpp:335)
      const Box b = ghostSlab(g, 1, side, comp);
      const dim3 blk = slabBlock(b);
      axisFillKernel<<<boxGrid(b, 1, blk), blk, 0, st>>>(Vs + comp * g.vsn, g.vsj, g.vsk, 0, b, 1, jref, offset,
                                                         comp == 0 ? 0 : 1, twoPi, ngk, nkint);
      (*launches)++;
    }
  }
  return (int)cudaGetLastError();
}

int idfxk::axisRegularizeBx2s(const DevGrid &g, double *Vs, int twoPi, int left, int right, cudaStream_t st,
                              long long *launches) {
  const dim3 block(32, 4, 1);
  const dim3 grid((g.nx[0] + 31) / 32, (g.nx[2] + 3) / 4, 1);
  axisBx2sKernel<<<grid, block, 0, st>>>(g, Vs, twoPi, left, right);
  (*launches)++;
  return (int)cudaGetLastError();
}

int idfxk::axisRegularizeEmf(const DevGrid &g, double *ex, double *ez, int twoPi, int left, int right, cudaStream_t st,
                             long long *launches) {
  for (int side = 0; side < 2; side++) {
    if (!(side == 0 ? left : right)) continue;
    const int jref = side == 0 ? g.beg[1] : g.end[1];
    axisEmfKernel<<<(g.nx[0] + 63) / 64, 64, 0, st>>>(g, ex, ez, jref, twoPi);
    (*launches)++;
  }
  return (int)cudaGetLastError();
}
