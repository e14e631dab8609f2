// arrays.hpp -- IdefixArray{1..4}D / IdefixHostArray{1..4}D (reference: src/arrays.hpp:12-40).
//
// Same contract as the reference's Kokkos::View aliases: a cheap, copyable, reference-counted
// handle; LayoutRight (last index fastest); operator() usable in device lambdas; extent(n),
// data(), size(); zero-initialised on allocation.  A handle can also wrap memory it does not own
// (the state arrays of the IdfxBlock, obtained through idfx_field_info), so user kernels, host
// mirrors and the library all see the same memory.
#include "idefix.hpp"
#ifndef IDEFIX_B200_HOST_ARRAYS_HPP_
#define IDEFIX_B200_HOST_ARRAYS_HPP_

struct IdfxDeviceSpace {};
struct IdfxHostSpace {};

template <typename T, int N, typename Space>
class IdfxArray {
 public:
  using value_type = T;
  IDFX_HOST_DEVICE IdfxArray() {}
  // owning allocation, zero filled (Kokkos::View initialises its allocation)
  template <typename... Ext>
  explicit IdfxArray(const std::string &label, Ext... ext) {
    static_assert(sizeof...(Ext) == N, "wrong number of extents");
    const long long e[N] = {static_cast<long long>(ext)...};
    long long n = 1;
    for (int d = 0; d < N; d++) {
      ext_[d] = static_cast<int>(e[d]);
      n *= e[d];
    }
    (void)label;
    if (n > 0) {
      if constexpr (std::is_same<Space, IdfxDeviceSpace>::value) {
        IDFX_CUDA(cudaMalloc(reinterpret_cast<void **>(&ptr_), n * sizeof(T)));
        IDFX_CUDA(cudaMemset(ptr_, 0, n * sizeof(T)));
      } else {
        ptr_ = static_cast<T *>(std::calloc(n, sizeof(T)));
        if (ptr_ == nullptr) IDEFIX_ERROR("host allocation failed");
      }
      count_ = new int(1);
    }
  }
  // non-owning view of existing memory
  template <typename... Ext>
  explicit IdfxArray(T *p, Ext... ext) : ptr_(p) {
    static_assert(sizeof...(Ext) == N, "wrong number of extents");
    const int e[N] = {static_cast<int>(ext)...};
    for (int d = 0; d < N; d++) ext_[d] = e[d];
  }
  IDFX_HOST_DEVICE IdfxArray(const IdfxArray &o) : ptr_(o.ptr_), count_(o.count_) {
    for (int d = 0; d < N; d++) ext_[d] = o.ext_[d];
#ifndef __CUDA_ARCH__
    if (count_) ++*count_;
#endif
  }
  IDFX_HOST_DEVICE IdfxArray &operator=(const IdfxArray &o) {
    if (this == &o) return *this;
#ifndef __CUDA_ARCH__
    release();
#endif
    ptr_ = o.ptr_;
    count_ = o.count_;
    for (int d = 0; d < N; d++) ext_[d] = o.ext_[d];
#ifndef __CUDA_ARCH__
    if (count_) ++*count_;
#endif
    return *this;
  }
  IDFX_HOST_DEVICE ~IdfxArray() {
#ifndef __CUDA_ARCH__
    release();
#endif
  }

  IDFX_HOST_DEVICE T &operator()(int i) const {
    static_assert(N == 1, "rank mismatch");
    return ptr_[i];
  }
  IDFX_HOST_DEVICE T &operator()(int j, int i) const {
    static_assert(N == 2, "rank mismatch");
    return ptr_[static_cast<long long>(j) * ext_[1] + i];
  }
  IDFX_HOST_DEVICE T &operator()(int k, int j, int i) const {
    static_assert(N == 3, "rank mismatch");
    return ptr_[(static_cast<long long>(k) * ext_[1] + j) * ext_[2] + i];
  }
  IDFX_HOST_DEVICE T &operator()(int n, int k, int j, int i) const {
    static_assert(N == 4, "rank mismatch");
    return ptr_[((static_cast<long long>(n) * ext_[1] + k) * ext_[2] + j) * ext_[3] + i];
  }
  IDFX_HOST_DEVICE int extent(int d) const { return ext_[d]; }
  IDFX_HOST_DEVICE T *data() const { return ptr_; }
  IDFX_HOST_DEVICE long long size() const {
    long long n = 1;
    for (int d = 0; d < N; d++) n *= ext_[d];
    return n;
  }
  IDFX_HOST_DEVICE bool is_allocated() const { return ptr_ != nullptr; }

 private:
  void release() {
    if (count_ && --*count_ == 0) {
      if constexpr (std::is_same<Space, IdfxDeviceSpace>::value) cudaFree(ptr_);
      else std::free(ptr_);
      delete count_;
    }
    count_ = nullptr;
    ptr_ = nullptr;
  }
  T *ptr_{nullptr};
  int ext_[N] = {};
  int *count_{nullptr};
};

template <typename T> using IdefixArray1D = IdfxArray<T, 1, IdfxDeviceSpace>;
template <typename T> using IdefixArray2D = IdfxArray<T, 2, IdfxDeviceSpace>;
template <typename T> using IdefixArray3D = IdfxArray<T, 3, IdfxDeviceSpace>;
template <typename T> using IdefixArray4D = IdfxArray<T, 4, IdfxDeviceSpace>;
template <typename T> using IdefixHostArray1D = IdfxArray<T, 1, IdfxHostSpace>;
template <typename T> using IdefixHostArray2D = IdfxArray<T, 2, IdfxHostSpace>;
template <typename T> using IdefixHostArray3D = IdfxArray<T, 3, IdfxHostSpace>;
template <typename T> using IdefixHostArray4D = IdfxArray<T, 4, IdfxHostSpace>;

// The few Kokkos entry points user set-ups call directly (deep_copy, create_mirror_view, fence).
namespace Kokkos {
template <typename T, int N, typename SD, typename SS>
inline void deep_copy(const IdfxArray<T, N, SD> &dst, const IdfxArray<T, N, SS> &src) {
  if (dst.size() != src.size()) IDEFIX_ERROR("deep_copy: size mismatch");
  if (dst.size() == 0) return;
  IDFX_CUDA(cudaMemcpy(dst.data(), src.data(), dst.size() * sizeof(T), cudaMemcpyDefault));
}
template <typename T, int N, typename S>
inline void deep_copy(const IdfxArray<T, N, S> &dst, const T &value) {
  std::vector<T> tmp(dst.size(), value);
  if (dst.size() == 0) return;
  IDFX_CUDA(cudaMemcpy(dst.data(), tmp.data(), dst.size() * sizeof(T), cudaMemcpyDefault));
}
template <typename T, typename S>
inline IdfxArray<T, 1, IdfxHostSpace> create_mirror_view(const IdfxArray<T, 1, S> &a) {
  return IdfxArray<T, 1, IdfxHostSpace>("mirror", a.extent(0));
}
template <typename T, typename S>
inline IdfxArray<T, 2, IdfxHostSpace> create_mirror_view(const IdfxArray<T, 2, S> &a) {
  return IdfxArray<T, 2, IdfxHostSpace>("mirror", a.extent(0), a.extent(1));
}
template <typename T, typename S>
inline IdfxArray<T, 3, IdfxHostSpace> create_mirror_view(const IdfxArray<T, 3, S> &a) {
  return IdfxArray<T, 3, IdfxHostSpace>("mirror", a.extent(0), a.extent(1), a.extent(2));
}
template <typename T, typename S>
inline IdfxArray<T, 4, IdfxHostSpace> create_mirror_view(const IdfxArray<T, 4, S> &a) {
  return IdfxArray<T, 4, IdfxHostSpace>("mirror", a.extent(0), a.extent(1), a.extent(2), a.extent(3));
}
inline void fence() { IDFX_CUDA(cudaDeviceSynchronize()); }
}  // namespace Kokkos

#endif  // IDEFIX_B200_HOST_ARRAYS_HPP_
