// selftest.cu -- device self-test of the tensor-memory scratchpad the sweep kernels rely on (tmem.cuh).
//
// The marches park per-thread doubles in TMEM with four CTAs co-resident per SM, each owning its own column
// allocation.  This kernel reproduces that situation with a data pattern: every thread writes doubles that encode
// (CTA, thread, slot) into its lane, lets the other warps run, reads them back many times and counts mismatches.
#include "../../include/idefix_b200.h"
#include "tmem.cuh"

#include <string>

namespace {

constexpr int TX = 128, SLOTS = 17, TCOLS = 64;

__device__ __forceinline__ double pattern(unsigned cta, unsigned tid, int slot, int round) {
  return (double)cta * 1048576.0 + (double)tid * 4096.0 + (double)slot * 64.0 + (double)round + 0.5;
}

__global__ void __launch_bounds__(TX, 4) tmemSelfTestKernel(unsigned long long *mismatches, int rounds) {
  __shared__ unsigned slot;
  // shared memory of the size the marches use, so that residency (4 CTAs / SM) is the same
  extern __shared__ double pad[];
  pad[threadIdx.x] = 0.0;
  const unsigned base = tmem::allocate<TCOLS>(&slot);
  unsigned long long bad = 0;
  for (int round = 0; round < rounds; round++) {
    __syncwarp();
#pragma unroll
    for (int s = 0; s < SLOTS; s++) tmem::store1(base + 2 * s, pattern(blockIdx.x, threadIdx.x, s, round));
    tmem::waitStores();
    // 8-wide path on the first eight slots
    __syncwarp();
    const tmem::Pending8 p8 = tmem::load8(base);
    tmem::Pending1 p1[SLOTS];
#pragma unroll
    for (int s = 0; s < SLOTS; s++) p1[s] = tmem::load1(base + 2 * s);
    tmem::waitLoads();
#pragma unroll
    for (int s = 0; s < SLOTS; s++) {
      const double want = pattern(blockIdx.x, threadIdx.x, s, round);
      if (p1[s].value() != want) bad++;
      if (s < 8 && p8.value(s) != want) bad++;
    }
    pad[threadIdx.x] += 1.0;
  }
  if (bad) atomicAdd(mismatches, bad);
  tmem::release<TCOLS>(&slot);
}

}  // namespace

extern "C" int idfx_selftest_tmem(int device, long long *mismatches) {
  if (!mismatches) return 1;
  *mismatches = -1;
  if (cudaSetDevice(device) != cudaSuccess) return 1;
  unsigned long long *d = nullptr;
  if (cudaMalloc(&d, sizeof(*d)) != cudaSuccess) return 1;
  cudaMemset(d, 0, sizeof(*d));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const size_t smem = 28 * 1024;
  cudaFuncSetAttribute(tmemSelfTestKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  tmemSelfTestKernel<<<sms * 4 * 6, TX, smem>>>(d, 200);
  unsigned long long h = 0;
  const cudaError_t e = cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return 1;
  *mismatches = (long long)h;
  return 0;
}

// ---- TMA self-test -------------------------------------------------------------------------------------------------
// The sweeps load their stencil boxes with cp.async.bulk.tensor (tma.cuh): rank-4 FP64 descriptor over a
// (nvar, n3, n2, n1) array, one box [nvar][1][1][cells] per warp and step, completion on a per-warp mbarrier, boxes
// that start before the array or run past its end (zero fill).  This kernel does exactly that on a pattern array and
// compares every element of every box with a plain global load.
#include "tma.cuh"

namespace {

constexpr int ST_NV = 8, ST_CELLS = 34, ST_WARPS = 4;

__global__ void __launch_bounds__(ST_WARPS * 32) tmaSelfTestKernel(const double *arr, int n1, int n2, int n3,
                                                                   unsigned long long *mismatches,
                                                                   const __grid_constant__ CUtensorMap boxes) {
  extern __shared__ __align__(128) double sm[];
  constexpr int BUF = (ST_NV * ST_CELLS + 15) / 16 * 16;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double *rows = sm + warp * 2 * BUF;
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(sm + ST_WARPS * 2 * BUF) + 2 * warp;
  const unsigned rowsS = (unsigned)__cvta_generic_to_shared(rows), barS = (unsigned)__cvta_generic_to_shared(bars);
  if (lane == 0) {
    tma::barrierInit(barS, 1);
    tma::barrierInit(barS + 8, 1);
    tma::fenceBarrierInit();
  }
  __syncthreads();
  const int k = blockIdx.x % n3;
  // some boxes start before 0, some run past n1; always at an even index (an odd one is an illegal instruction)
  const int x0 = -2 + 30 * (int)(blockIdx.x / n3) + 8 * warp;
  unsigned phases = 0;
  unsigned long long bad = 0;
  auto issue = [&](int j, int buf) {
    if (lane == 0) {
      tma::barrierExpect(barS + 8 * buf, ST_NV * ST_CELLS * 8);
      tma::load4D(rowsS + buf * BUF * 8, &boxes, barS + 8 * buf, x0, j, k, 0);
    }
  };
  issue(0, 0);
  for (int j = 0; j < n2; j++) {
    const int buf = j & 1;
    if (j + 1 < n2) issue(j + 1, buf ^ 1);
    tma::barrierWait(barS + 8 * buf, (phases >> buf) & 1u);
    phases ^= (1u << buf);
    __syncwarp();
    for (int e = lane; e < ST_NV * ST_CELLS; e += 32) {
      const int nv = e / ST_CELLS, x = x0 + e % ST_CELLS;
      const double want = (x >= 0 && x < n1) ? arr[(((long long)nv * n3 + k) * n2 + j) * n1 + x] : 0.0;
      if (rows[buf * BUF + e] != want) bad++;
    }
    __syncwarp();
  }
  if (bad) atomicAdd(mismatches, bad);
}

__global__ void tmaPatternKernel(double *arr, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) arr[i] = 1.0 + (double)i * 0.5;
}

}  // namespace

extern "C" int idfx_selftest_tma(int device, long long *mismatches) {
  if (!mismatches) return 1;
  *mismatches = -1;
  if (cudaSetDevice(device) != cudaSuccess) return 1;
  const int n1 = 100, n2 = 9, n3 = 5;
  const long long n = (long long)ST_NV * n3 * n2 * n1;
  double *arr = nullptr;
  unsigned long long *d = nullptr;
  if (cudaMalloc(&arr, n * sizeof(double)) != cudaSuccess || cudaMalloc(&d, sizeof(*d)) != cudaSuccess) return 1;
  cudaMemset(d, 0, sizeof(*d));
  tmaPatternKernel<<<(unsigned)((n + 255) / 256), 256>>>(arr, n);
  CUtensorMap map;
  int rc = 1;
  if (tma::encodeCellBox(&map, arr, n1, n2, n3, ST_NV, ST_CELLS)) {
    constexpr int BUF = (ST_NV * ST_CELLS + 15) / 16 * 16;
    const size_t smem = sizeof(double) * (ST_WARPS * 2 * BUF + 2 * ST_WARPS);
    cudaFuncSetAttribute(tmaSelfTestKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    tmaSelfTestKernel<<<n3 * 4, ST_WARPS * 32, smem>>>(arr, n1, n2, n3, d, map);
    unsigned long long h = 0;
    if (cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost) == cudaSuccess) {
      *mismatches = (long long)h;
      rc = 0;
    }
  } else {
    *mismatches = -2;   // the driver refused the descriptor
    rc = 0;
  }
  cudaFree(arr);
  cudaFree(d);
  return rc;
}
