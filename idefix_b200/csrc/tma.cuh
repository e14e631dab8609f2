// tma.cuh -- tensor-memory-accelerator (TMA) loads of the sweep kernels: device wrappers + host-side descriptors.
//
// A warp of the x1 sweep needs, per row, one box of primitives: [nvar][1][1][32 + 2H cells along x1].
// With cp.async that is 2 x nvar predicated LDGSTS per lane plus their 64-bit address arithmetic, and the copies
// allocate in L1 on their way -- evicting the register spills that live there (ncu: every long-scoreboard stall of
// the sweeps sits on an LDL).  One elected lane issues ONE cp.async.bulk.tensor instead (SASS: UTMALDG); the box goes
// L2 -> shared memory without touching L1 and completes on a per-warp mbarrier.
// (The x2 / x3 marches keep per-lane cp.async: their TMA variant -- one box per ring slot, issued in the middle of a
// step between tcgen05 accesses -- never completed on the device in this round's experiments and was withdrawn.)
//
// Descriptor limits (CUDA driver API, cuTensorMapEncodeTiled): base address 16-byte aligned, every stride a
// multiple of 16 bytes (=> nx1 even for FP64 rows), box inner extent a multiple of 16 bytes.  And, measured on the
// B200 (idfx_selftest_tma under compute-sanitizer): the FIRST ELEMENT of a box must itself sit on a 16-byte boundary,
// i.e. an FP64 box must start at an even x1 index -- an odd start coordinate is an illegal-instruction fault.  Boxes
// of the x1 sweep start at cell 30 * strip, which is even by construction.  An array narrower than the box is accepted
// by the encoder but the copy never completes.  Blocks that do not qualify keep the per-lane cp.async path.
#ifndef IDFX_TMA_CUH_
#define IDFX_TMA_CUH_

#include <cuda.h>   // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)

#include "common.cuh"

namespace tma {

#define IDFX_TMA_DEV __device__ __forceinline__

IDFX_TMA_DEV void barrierInit(unsigned bar, unsigned arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(arrivals) : "memory");
}
// makes freshly initialised barriers visible to the async proxy
IDFX_TMA_DEV void fenceBarrierInit() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// orders this thread's generic-proxy shared-memory accesses before subsequent async-proxy (TMA) accesses
IDFX_TMA_DEV void fenceProxyAsync() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

IDFX_TMA_DEV void barrierExpect(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// one poll: true when the phase with the given parity of the barrier has completed
IDFX_TMA_DEV bool barrierTryWait(unsigned bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// waits for the phase with the given parity.  A copy that never completes (a box the hardware rejects -- the issuing
// lane then dies with an illegal-instruction fault while the other lanes keep polling --, a byte count that does not
// match the box) must surface as an error, not as a hung device: after ~2 s of polling the kernel traps.
IDFX_TMA_DEV void barrierWait(unsigned bar, unsigned parity) {
  if (barrierTryWait(bar, parity)) return;
  const long long t0 = clock64();
  while (!barrierTryWait(bar, parity)) {
    if (clock64() - t0 > (1ll << 32)) __trap();
  }
}

// box of a rank-4 tensor (x1, x2, x3, variable) -> shared memory, completion counted in bytes on `bar`
IDFX_TMA_DEV void load4D(unsigned dstShared, const CUtensorMap *map, unsigned bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          dstShared),
      "l"(reinterpret_cast<unsigned long long>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// ---- host ------------------------------------------------------------------------------------------------
// Descriptor of a (nvar, n3, n2, n1) FP64 array for boxes of `boxCells` cells along x1 and all `nvar` variables
// of one (x2, x3) position.  Returns false when the array does not qualify (see the limits above) or the
// driver entry point is unavailable; the caller then stays on the cp.async path.
inline bool encodeCellBox(CUtensorMap *out, const double *base, int n1, int n2, int n3, int nvar, int boxCells) {
  typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                               const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                               CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  static bool looked = false;
  if (!looked) {
    looked = true;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      encode = reinterpret_cast<EncodeFn>(fn);
    cudaGetLastError();
  }
  if (!encode) return false;
  // (a box wider than the array itself -- tiny test grids -- is accepted by the encoder but never completes on the
  // device: measured, idfx_selftest_tma / tests/test_oracle_gpu.py)
  if ((reinterpret_cast<unsigned long long>(base) & 15ull) || (n1 & 1) || boxCells > 256 || nvar > 256 ||
      ((boxCells * 8) & 15) || n1 < boxCells)
    return false;
  const cuuint64_t dims[4] = {(cuuint64_t)n1, (cuuint64_t)n2, (cuuint64_t)n3, (cuuint64_t)nvar};
  const cuuint64_t strides[3] = {(cuuint64_t)n1 * 8, (cuuint64_t)n1 * n2 * 8, (cuuint64_t)n1 * n2 * n3 * 8};
  const cuuint32_t box[4] = {(cuuint32_t)boxCells, 1, 1, (cuuint32_t)nvar};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<double *>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

}  // namespace tma

#endif  // IDFX_TMA_CUH_
