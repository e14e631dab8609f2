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
