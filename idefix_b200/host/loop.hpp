// loop.hpp -- idefix_for / idefix_reduce (reference: src/loop.hpp, src/reduce.hpp:62-81).
//
// The reference dispatches these to Kokkos policies; here each one is a plain CUDA kernel launch
// of the user's extended lambda on the default stream (the same stream the library's block uses
// unless idfx_set_stream says otherwise), threads running along the last (fastest) index.
// Ranges are [begin, end) like the reference's.
#include "idefix.hpp"
#ifndef IDEFIX_B200_HOST_LOOP_HPP_
#define IDEFIX_B200_HOST_LOOP_HPP_

#ifdef __CUDACC__
namespace idfxloop {
template <typename F>
__global__ void for1D(int ib, int ie, F f) {
  const int i = ib + blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ie) f(i);
}
template <typename F>
__global__ void for2D(int jb, int je, int ib, int ie, F f) {
  const int i = ib + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = jb + blockIdx.y * blockDim.y + threadIdx.y;
  if (i < ie && j < je) f(j, i);
}
template <typename F>
__global__ void for3D(int kb, int ke, int jb, int je, int ib, int ie, F f) {
  const int i = ib + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = jb + blockIdx.y * blockDim.y + threadIdx.y;
  const int k = kb + blockIdx.z;
  if (i < ie && j < je && k < ke) f(k, j, i);
}
template <typename F>
__global__ void for4D(int nb, int ne, int kb, int ke, int jb, int je, int ib, int ie, F f) {
  const int i = ib + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = jb + blockIdx.y * blockDim.y + threadIdx.y;
  const int nk = blockIdx.z;
  const int n = nb + nk / (ke - kb), k = kb + nk % (ke - kb);
  if (i < ie && j < je && n < ne) f(n, k, j, i);
}
inline void check(const char *name) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) IDEFIX_ERROR(std::string("idefix_for(") + name + "): " + cudaGetErrorString(e));
}

// reduction: every block folds its cells in a fixed order and writes one partial; the partials are
// folded on the host in block order, so the result does not depend on scheduling.
template <typename T, typename Op, typename F>
__global__ void reduce3D(int kb, int ke, int jb, int je, int ib, int ie, F f, T init, T *partials) {
  __shared__ T sh[256];
  const int ni = ie - ib, nj = je - jb;
  const long long n = static_cast<long long>(ni) * nj * (ke - kb);
  T acc = init;
  for (long long c = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; c < n;
       c += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int i = ib + static_cast<int>(c % ni);
    const int j = jb + static_cast<int>((c / ni) % nj);
    const int k = kb + static_cast<int>(c / (static_cast<long long>(ni) * nj));
    f(k, j, i, acc);
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] = Op::apply(sh[threadIdx.x], sh[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) partials[blockIdx.x] = sh[0];
}
}  // namespace idfxloop

template <typename F>
inline void idefix_for(const std::string &name, int ib, int ie, F f) {
  if (ie <= ib) return;
  idfxloop::for1D<<<(ie - ib + 127) / 128, 128>>>(ib, ie, f);
  idfxloop::check(name.c_str());
}
template <typename F>
inline void idefix_for(const std::string &name, int jb, int je, int ib, int ie, F f) {
  if (ie <= ib || je <= jb) return;
  dim3 b(64, 4), g((ie - ib + 63) / 64, (je - jb + 3) / 4);
  idfxloop::for2D<<<g, b>>>(jb, je, ib, ie, f);
  idfxloop::check(name.c_str());
}
template <typename F>
inline void idefix_for(const std::string &name, int kb, int ke, int jb, int je, int ib, int ie, F f) {
  if (ie <= ib || je <= jb || ke <= kb) return;
  dim3 b(64, 4), g((ie - ib + 63) / 64, (je - jb + 3) / 4, ke - kb);
  idfxloop::for3D<<<g, b>>>(kb, ke, jb, je, ib, ie, f);
  idfxloop::check(name.c_str());
}
template <typename F>
inline void idefix_for(const std::string &name, int nb, int ne, int kb, int ke, int jb, int je, int ib, int ie, F f) {
  if (ie <= ib || je <= jb || ke <= kb || ne <= nb) return;
  dim3 b(64, 4), g((ie - ib + 63) / 64, (je - jb + 3) / 4, (ke - kb) * (ne - nb));
  idfxloop::for4D<<<g, b>>>(nb, ne, kb, ke, jb, je, ib, ie, f);
  idfxloop::check(name.c_str());
}

namespace Kokkos {
template <typename T> struct Sum {
  T &ref;
  explicit Sum(T &r) : ref(r) {}
  static IDFX_HOST_DEVICE T apply(T a, T b) { return a + b; }
  static T init() { return T(0); }
};
template <typename T> struct Min {
  T &ref;
  explicit Min(T &r) : ref(r) {}
  static IDFX_HOST_DEVICE T apply(T a, T b) { return a < b ? a : b; }
  static T init() { return std::numeric_limits<T>::max(); }
};
template <typename T> struct Max {
  T &ref;
  explicit Max(T &r) : ref(r) {}
  static IDFX_HOST_DEVICE T apply(T a, T b) { return a > b ? a : b; }
  static T init() { return std::numeric_limits<T>::lowest(); }
};
}  // namespace Kokkos

template <typename F, template <typename> class R, typename T>
inline void idefix_reduce(const std::string &name, int kb, int ke, int jb, int je, int ib, int ie, F f, R<T> red) {
  red.ref = R<T>::init();
  if (ie <= ib || je <= jb || ke <= kb) return;
  const int nblocks = 1024;
  T *partials = nullptr;
  IDFX_CUDA(cudaMalloc(reinterpret_cast<void **>(&partials), nblocks * sizeof(T)));
  idfxloop::reduce3D<T, R<T>><<<nblocks, 256>>>(kb, ke, jb, je, ib, ie, f, R<T>::init(), partials);
  idfxloop::check(name.c_str());
  std::vector<T> host(nblocks);
  IDFX_CUDA(cudaMemcpy(host.data(), partials, nblocks * sizeof(T), cudaMemcpyDeviceToHost));
  IDFX_CUDA(cudaFree(partials));
  T acc = R<T>::init();
  for (int b = 0; b < nblocks; b++) acc = R<T>::apply(acc, host[b]);
  red.ref = acc;
}
#endif  // __CUDACC__

#endif  // IDEFIX_B200_HOST_LOOP_HPP_
