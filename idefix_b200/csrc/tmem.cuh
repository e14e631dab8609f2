// tmem.cuh -- Blackwell tensor memory (TMEM) used as a per-thread FP64 scratchpad by the sweep kernels.
//
// The sweeps are bound by FP64 dependent-issue latency at 16 resident warps per SM: registers (128 per thread)
// and shared memory (the per-thread stencil ring) are both full, and what spills goes to local memory, whose
// L1 share is what shared memory leaves over.  The 256 KB of TMEM per SM are otherwise idle on this path (no
// tensor-core work), and a thread can park private values there: 512 columns x 128 lanes x 32 bit, lane l of
// the CTA's allocation belongs to thread l (warp w reaches lanes 32*(w%4) .. 32*(w%4)+31 with the 32x32b
// shape), a double takes two consecutive columns.  tcgen05.ld latency is ~12 cycles (LDS: ~29).
//
// Rules (PTX ISA, tcgen05): alloc / dealloc / ld / st are .sync.aligned -- every thread of the warp must execute
// them convergently; columns are allocated per CTA in powers of two >= 32 by ONE warp, which also frees them; a
// tcgen05.ld result may be read only after tcgen05.wait::ld, a stored value re-read only after tcgen05.wait::st.
#ifndef IDFX_TMEM_CUH_
#define IDFX_TMEM_CUH_

#include "common.cuh"

namespace tmem {

#define IDFX_TMEM_DEV __device__ __forceinline__

// One warp of the CTA allocates NCOLS columns; the base address is published through *slot (shared memory).
// Call from ALL threads of the CTA (contains a __syncthreads); returns the per-warp base address (lane field set).
template <int NCOLS>
IDFX_TMEM_DEV unsigned allocate(unsigned *slot) {
  static_assert(NCOLS >= 32 && NCOLS <= 512 && (NCOLS & (NCOLS - 1)) == 0, "TMEM columns: power of two in [32, 512]");
  if ((threadIdx.x >> 5) == 0) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(slot);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s), "n"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned base = *slot;
  return base + ((((threadIdx.x >> 5) & 3u) * 32u) << 16);
}

// Call from ALL threads of the CTA once nobody touches the allocation any more (contains a __syncthreads).
template <int NCOLS>
IDFX_TMEM_DEV void release(const unsigned *slot) {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) {
    const unsigned base = *slot;
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(NCOLS) : "memory");
  }
}

IDFX_TMEM_DEV void waitLoads() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
IDFX_TMEM_DEV void waitStores() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// One double = columns [col, col+1] of the calling thread's lane.  `taddr` = allocate()'s return value + column.
IDFX_TMEM_DEV void store1(unsigned taddr, double v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(__double2loint(v)),
               "r"(__double2hiint(v))
               : "memory");
}
// raw halves of a load in flight; combine with value() after waitLoads()
struct Pending1 {
  int lo, hi;
  IDFX_TMEM_DEV double value() const { return __hiloint2double(hi, lo); }
};
IDFX_TMEM_DEV Pending1 load1(unsigned taddr) {
  Pending1 p;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(p.lo), "=r"(p.hi) : "r"(taddr) : "memory");
  return p;
}

// Four doubles = 8 consecutive columns in one instruction.
IDFX_TMEM_DEV void store4(unsigned taddr, const double *v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(__double2loint(v[0])), "r"(__double2hiint(v[0])), "r"(__double2loint(v[1])), "r"(__double2hiint(v[1])),
               "r"(__double2loint(v[2])), "r"(__double2hiint(v[2])), "r"(__double2loint(v[3])), "r"(__double2hiint(v[3]))
               : "memory");
}
struct Pending4 {
  int w[8];
  IDFX_TMEM_DEV double value(int n) const { return __hiloint2double(w[2 * n + 1], w[2 * n]); }
};
IDFX_TMEM_DEV Pending4 load4(unsigned taddr) {
  Pending4 p;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(p.w[0]), "=r"(p.w[1]), "=r"(p.w[2]), "=r"(p.w[3]), "=r"(p.w[4]), "=r"(p.w[5]), "=r"(p.w[6]), "=r"(p.w[7])
               : "r"(taddr)
               : "memory");
  return p;
}

// Eight doubles = 16 consecutive columns in one instruction.
IDFX_TMEM_DEV void store8(unsigned taddr, const double *v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(
          taddr),
      "r"(__double2loint(v[0])), "r"(__double2hiint(v[0])), "r"(__double2loint(v[1])), "r"(__double2hiint(v[1])),
      "r"(__double2loint(v[2])), "r"(__double2hiint(v[2])), "r"(__double2loint(v[3])), "r"(__double2hiint(v[3])),
      "r"(__double2loint(v[4])), "r"(__double2hiint(v[4])), "r"(__double2loint(v[5])), "r"(__double2hiint(v[5])),
      "r"(__double2loint(v[6])), "r"(__double2hiint(v[6])), "r"(__double2loint(v[7])), "r"(__double2hiint(v[7]))
      : "memory");
}
struct Pending8 {
  int w[16];
  IDFX_TMEM_DEV double value(int n) const { return __hiloint2double(w[2 * n + 1], w[2 * n]); }
};
IDFX_TMEM_DEV Pending8 load8(unsigned taddr) {
  Pending8 p;
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(p.w[0]), "=r"(p.w[1]), "=r"(p.w[2]), "=r"(p.w[3]), "=r"(p.w[4]), "=r"(p.w[5]), "=r"(p.w[6]), "=r"(p.w[7]),
        "=r"(p.w[8]), "=r"(p.w[9]), "=r"(p.w[10]), "=r"(p.w[11]), "=r"(p.w[12]), "=r"(p.w[13]), "=r"(p.w[14]), "=r"(p.w[15])
      : "r"(taddr)
      : "memory");
  return p;
}

}  // namespace tmem

#endif  // IDFX_TMEM_CUH_
