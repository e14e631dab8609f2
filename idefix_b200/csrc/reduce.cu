// reduce.cu -- global reductions (CFL time step, div B, NaN count) and the device-side
// time-step controller that lets a whole RK cycle run without a host round trip.
//
// Reference: src/dataBlock/dataBlock.cpp:333-364 + src/reduce.hpp:62-81 (ComputeTimestep),
// src/timeIntegrator.cpp:313-323,380-408 (dt logic), src/fluid/checkDivB.hpp:15-68,
// src/fluid/checkNan.hpp:15-60.
// min / max are order independent, so the results are bit-identical to the reference's.
// Reductions: warp shuffles, one shared-memory step per CTA, one 64-bit atomic per CTA on the
// order-preserving bit pattern of a non-negative double.
#include "kernels.h"

namespace {

constexpr int TB = 256;

__device__ __forceinline__ double warpMin(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warpMax(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// grid-stride loop over the interior cells; persistent grid of (SM count x resident CTAs) blocks
__global__ void timestepKernel(DevGrid g, const double *__restrict__ invDt, unsigned long long *out) {
  __shared__ double sm[TB / 32];
  const int ni = g.end[0] - g.beg[0], nj = g.end[1] - g.beg[1], nk = g.end[2] - g.beg[2];
  const long long n = (long long)ni * nj * nk;
  double m = __longlong_as_double(0x7ff0000000000000LL);  // +inf
  for (long long t = (long long)blockIdx.x * TB + threadIdx.x; t < n; t += (long long)gridDim.x * TB) {
    const int i = g.beg[0] + (int)(t % ni);
    const int j = g.beg[1] + (int)((t / ni) % nj);
    const int k = g.beg[2] + (int)(t / ((long long)ni * nj));
    m = fmin(1.0 / invDt[cidx(g, k, j, i)], m);
  }
  m = warpMin(m);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = (threadIdx.x < TB / 32) ? sm[threadIdx.x] : __longlong_as_double(0x7ff0000000000000LL);
    m = warpMin(m);
    if (threadIdx.x == 0 && m >= 0.0) atomicMin(out, (unsigned long long)__double_as_longlong(m));
  }
}

__global__ void divBKernel(DevGrid g, const double *__restrict__ Vs, unsigned long long *out) {
  __shared__ double sm[TB / 32];
  const int ni = g.end[0] - g.beg[0], nj = g.end[1] - g.beg[1], nk = g.end[2] - g.beg[2];
  const long long n = (long long)ni * nj * nk;
  const int nd = g.ndim;
  double m = 0.0;
  for (long long t = (long long)blockIdx.x * TB + threadIdx.x; t < n; t += (long long)gridDim.x * TB) {
    const int i = g.beg[0] + (int)(t % ni);
    const int j = g.beg[1] + (int)((t / ni) % nj);
    const int k = g.beg[2] + (int)(t / ((long long)ni * nj));
    const long long s = sidx(g, k, j, i);
    double dB1 = 0, dB2 = 0, dB3 = 0, d1 = 0, d2 = 0, d3 = 0, B1 = 0, B2 = 0, B3 = 0;
    {
      const double Ap = faceAreaDir(g, 0, k, j, i + 1), Am = faceAreaDir(g, 0, k, j, i);
      const double bp = Vs[s + 1], bm = Vs[s];
      dB1 = (Ap * bp - Am * bm);
      d1 = 0.5 * (Ap + Am);
      B1 = 0.5 * (bp + bm);
    }
    if (nd >= 2) {
      const double Ap = faceAreaDir(g, 1, k, j + 1, i), Am = faceAreaDir(g, 1, k, j, i);
      const double bp = Vs[g.vsn + s + g.vsj], bm = Vs[g.vsn + s];
      dB2 = (Ap * bp - Am * bm);
      d2 = 0.5 * (Ap + Am);
      B2 = 0.5 * (bp + bm);
    }
    if (nd == 3) {
      const double Ap = faceAreaDir(g, 2, k + 1, j, i), Am = faceAreaDir(g, 2, k, j, i);
      const double bp = Vs[2 * g.vsn + s + g.vsk], bm = Vs[2 * g.vsn + s];
      dB3 = (Ap * bp - Am * bm);
      d3 = 0.5 * (Ap + Am);
      B3 = 0.5 * (bp + bm);
    }
    double amplitude = 1e-40;
    double asum = fabs(B1) * d1;
    if (nd >= 2) asum = asum + fabs(B2) * d2;
    if (nd == 3) asum = asum + fabs(B3) * d3;
    amplitude += asum;
    double div = dB1;
    if (nd >= 2) div = div + dB2;
    if (nd == 3) div = div + dB3;
    m = fmax(fabs(div) / amplitude, m);
  }
  m = warpMax(m);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = (threadIdx.x < TB / 32) ? sm[threadIdx.x] : 0.0;
    m = warpMax(m);
    if (threadIdx.x == 0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
  }
}

__global__ void nanCountKernel(const double *__restrict__ q, long long sj, long long sk, long long sn, int nvars,
                               int lo0, int lo1, int lo2, int n0, int n1, int n2, int *out) {
  const long long n = (long long)n0 * n1 * n2 * nvars;
  int cnt = 0;
  for (long long t = (long long)blockIdx.x * TB + threadIdx.x; t < n; t += (long long)gridDim.x * TB) {
    const int i = lo0 + (int)(t % n0);
    const int j = lo1 + (int)((t / n0) % n1);
    const int k = lo2 + (int)((t / ((long long)n0 * n1)) % n2);
    const int v = (int)(t / ((long long)n0 * n1 * n2));
    if (isnan(q[v * sn + (long long)k * sk + (long long)j * sj + i])) cnt++;
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if ((threadIdx.x & 31) == 0 && cnt > 0) atomicAdd(out, cnt);
}

// state[0]=t, [1]=dt, [2]=t0 (time at cycle start), [3]=tstop (<0: none)
__global__ void dtBeginKernel(double *state, unsigned long long *dtMinBits) {
  // main.cpp:191: clip the step so that the run ends exactly at tstop
  const double tstop = state[3];
  if (tstop >= 0.0 && tstop - state[0] < state[1]) state[1] = tstop - state[0];
  state[2] = state[0];
  *dtMinBits = 0x7ff0000000000000ULL;
}

// timeIntegrator.cpp:302,333: t += dt after each stage, then t = wc*t + w0*t0 for stages > 0
__global__ void stageTimeKernel(double *state, int stage, double wc, double w0) {
  double t = state[0] + state[1];
  if (stage > 0) t = wc * t + w0 * state[2];
  state[0] = t;
}

// timeIntegrator.cpp:380-408
__global__ void dtEndKernel(double *state, const unsigned long long *dtMinBits, double cfl, double cflMaxVar,
                            double fixedDt, long long *ncycles, int *status, const unsigned int *mathFlag) {
  // the sweeps of this cycle left the validated envelope of their branch-free divisions / roots (fastmath.cuh):
  // the state is no longer guaranteed to be the reference's, the host must stop (status bit 2)
  if (mathFlag != nullptr && *mathFlag != 0u) *status |= 4;
  const double dt = state[1];
  state[0] = state[2] + dt;
  if (fixedDt > 0.0) {
    state[1] = fixedDt;
  } else {
    const double newdt = cfl * __longlong_as_double((long long)*dtMinBits);
    if (newdt > cflMaxVar * dt) {
      state[1] = cflMaxVar * dt;
    } else {
      if (*ncycles == 0 && newdt < 0.5 * dt) *status |= 2;  // first_dt too large
      state[1] = newdt;
    }
    if (state[1] < 1e-15) *status |= 1;  // dt too small
  }
  *ncycles = *ncycles + 1;
}

int persistentBlocks() {
  static int nb[64] = {};   // per device: a process may drive blocks on several GPUs
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (nb[dev] == 0) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    nb[dev] = sms > 0 ? sms * 8 : 1184;
  }
  return nb[dev];
}

}  // namespace

int idfxk::timestepReduce(const DevGrid &g, const double *invDt, unsigned long long *outBits, cudaStream_t st) {
  timestepKernel<<<persistentBlocks(), TB, 0, st>>>(g, invDt, outBits);
  return (int)cudaGetLastError();
}

int idfxk::dtControlBegin(double *state, unsigned long long *dtMinBits, cudaStream_t st) {
  dtBeginKernel<<<1, 1, 0, st>>>(state, dtMinBits);
  return (int)cudaGetLastError();
}
int idfxk::stageTimeAdvance(double *state, int stage, double wc, double w0, cudaStream_t st) {
  stageTimeKernel<<<1, 1, 0, st>>>(state, stage, wc, w0);
  return (int)cudaGetLastError();
}
int idfxk::dtControlEnd(double *state, unsigned long long *dtMinBits, double cfl, double cflMaxVar, double fixedDt,
                        long long *ncycles, int *status, const unsigned int *mathFlag, cudaStream_t st) {
  dtEndKernel<<<1, 1, 0, st>>>(state, dtMinBits, cfl, cflMaxVar, fixedDt, ncycles, status, mathFlag);
  return (int)cudaGetLastError();
}

int idfxk::checkDivB(const PhysCfg &c, const DevGrid &g, const double *Vs, unsigned long long *outBits,
                     cudaStream_t st) {
  divBKernel<<<persistentBlocks(), TB, 0, st>>>(g, Vs, outBits);
  return (int)cudaGetLastError();
}

int idfxk::checkNan(const PhysCfg &c, const DevGrid &g, const double *Vc, const double *Vs, int *count,
                    cudaStream_t st) {
  const int n0 = g.end[0] - g.beg[0], n1 = g.end[1] - g.beg[1], n2 = g.end[2] - g.beg[2];
  nanCountKernel<<<persistentBlocks(), TB, 0, st>>>(Vc, g.sj, g.sk, g.sn, c.nvar, g.beg[0], g.beg[1], g.beg[2], n0,
                                                    n1, n2, count);
  if (c.mhd) {
    nanCountKernel<<<persistentBlocks(), TB, 0, st>>>(Vs, g.vsj, g.vsk, g.vsn, g.ndim, g.beg[0], g.beg[1], g.beg[2],
                                                      n0 + g.off[0], n1 + g.off[1], n2 + g.off[2], count);
  }
  return (int)cudaGetLastError();
}
