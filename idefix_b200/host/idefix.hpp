// idefix.hpp -- the user-facing C++ surface of Idefix, re-implemented on top of the C ABI
// (include/idefix_b200.h -> libidefix_b200.so, hand-written sm_100a CUDA).
//
// A problem directory written for the reference (definitions.hpp + setup.cpp + idefix.ini)
// compiles against this header unchanged: same macros (src/idefix.hpp:47-220), same array
// handles (src/arrays.hpp), same loop helpers (src/loop.hpp), same classes and public data
// members (Input src/input.hpp, Grid src/grid.hpp, DataBlock src/dataBlock/dataBlock.hpp,
// Fluid/Hydro src/fluid/fluid.hpp, DataBlockHost src/dataBlock/dataBlockHost.hpp, Output
// src/output/output.hpp, Setup src/setup.hpp).  Like the reference, everything is compiled per
// problem (the physics configuration is a set of macros); the arithmetic of the update itself
// lives in the shared library and is selected at run time through IdfxBlockDesc.
//
// Out of scope: Fargo, gravity, non-ideal MHD, dust, RKL, planets, MPI.
#ifndef IDEFIX_B200_HOST_IDEFIX_HPP_
#define IDEFIX_B200_HOST_IDEFIX_HPP_

#include <cuda_runtime.h>

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "idefix_b200.h"

// ---- compile-time configuration (src/idefix.hpp:22-108) ----------------------------------------
#define YES 255
#define NO 0
#define CARTESIAN 1
#define CYLINDRICAL 2
#define POLAR 3
#define SPHERICAL 4
#define IDIR 0
#define JDIR 1
#define KDIR 2

#include "definitions.hpp"

#ifndef MHD
#define MHD NO
#endif
#ifndef ORDER
#define ORDER 2
#endif
#ifndef GEOMETRY
#define GEOMETRY CARTESIAN
#endif
#ifdef ISOTHERMAL
#define HAVE_ENERGY 0
#else
#define HAVE_ENERGY 1
#endif

#if COMPONENTS < DIMENSIONS
#error "COMPONENTS must be >= DIMENSIONS"
#endif

using real = double;
#define ZERO_F 0.0
#define HALF_F 0.5
#define ONE_F 1.0
#define TWO_F 2.0
#define THREE_F 3.0
#define FOUR_F 4.0
#define ONE_FOURTH_F 0.25
#define ONE_THIRD_F (1.0 / 3.0)
#define ONE_SIXTH_F (1.0 / 6.0)
#define SMALL_NUMBER 1e-10
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#define IDEFIX_VERSION "2.2.02-b200"

// variable indices (src/idefix.hpp:74-94): RHO, MX1.., ENG/PRS, BX1..
#define RHO 0
#define MX1 1
#define MX2 2
#define MX3 3
#define VX1 MX1
#define VX2 MX2
#define VX3 MX3
#if HAVE_ENERGY
#define ENG (COMPONENTS + 1)
#define PRS ENG
#endif
#if MHD == YES
#define BX1 (COMPONENTS + 1 + HAVE_ENERGY)
#define BX2 (BX1 + 1)
#define BX3 (BX1 + 2)
#define NVAR (1 + 2 * COMPONENTS + HAVE_ENERGY)
#else
#define NVAR (1 + COMPONENTS + HAVE_ENERGY)
#endif
#define TRG NVAR
#define BX1s 0
#define BX2s 1
#define BX3s 2
// edge-centred variables (idefix.hpp:116-125)
#ifdef EVOLVE_VECTOR_POTENTIAL
#if DIMENSIONS < 3
#define AX1e 250
#define AX2e 251
#define AX3e 0
#else
#define AX1e 0
#define AX2e 1
#define AX3e 2
#endif
#endif

#if DIMENSIONS == 1
#define IOFFSET 1
#define JOFFSET 0
#define KOFFSET 0
#define D_EXPAND(a, b, c) a
#define D_SELECT(a, b, c) a
#elif DIMENSIONS == 2
#define IOFFSET 1
#define JOFFSET 1
#define KOFFSET 0
#define D_EXPAND(a, b, c) a b
#define D_SELECT(a, b, c) b
#else
#define IOFFSET 1
#define JOFFSET 1
#define KOFFSET 1
#define D_EXPAND(a, b, c) a b c
#define D_SELECT(a, b, c) c
#endif
#if COMPONENTS == 1
#define EXPAND(a, b, c) a
#define SELECT(a, b, c) a
#elif COMPONENTS == 2
#define EXPAND(a, b, c) a b
#define SELECT(a, b, c) b
#else
#define EXPAND(a, b, c) a b c
#define SELECT(a, b, c) c
#endif

#ifdef __CUDACC__
#define KOKKOS_INLINE_FUNCTION __host__ __device__ inline
#define KOKKOS_FUNCTION __host__ __device__
#define KOKKOS_LAMBDA [=] __host__ __device__
#define KOKKOS_CLASS_LAMBDA [ =, *this ] __host__ __device__
#define IDFX_HOST_DEVICE __host__ __device__
#else
#define KOKKOS_INLINE_FUNCTION inline
#define KOKKOS_FUNCTION
#define KOKKOS_LAMBDA [=]
#define KOKKOS_CLASS_LAMBDA [ =, *this ]
#define IDFX_HOST_DEVICE
#endif
#define FMAX(a, b) ((a) > (b) ? (a) : (b))
#define FMIN(a, b) ((a) < (b) ? (a) : (b))
#define FABS(a) fabs(a)
#define FSIGN(a) ((a) >= 0.0 ? 1.0 : -1.0)

enum BoundaryType { internal, periodic, reflective, outflow, shearingbox, axis, userdef, undefined };
enum BoundarySide { left, right };

// ---- logging and errors (src/global.hpp, src/error.cpp:14-56) ------------------------------------
namespace idfx {
class IdefixOutStream {
 public:
  template <typename T>
  IdefixOutStream &operator<<(const T &x) {
    std::cout << x;
    if (log_.is_open()) log_ << x;
    return *this;
  }
  IdefixOutStream &operator<<(std::ostream &(*m)(std::ostream &)) {
    m(std::cout);
    if (log_.is_open()) m(log_);
    return *this;
  }
  void enableLogFile(const std::string &name) { log_.open(name); }

 private:
  std::ofstream log_;
};
inline IdefixOutStream cout;
inline int prank = 0;
inline int psize = 1;
inline void pushRegion(const std::string &) {}
inline void popRegion() {}
[[noreturn]] inline void safeExit(int code) {
  std::cout.flush();
  std::exit(code);
}
}  // namespace idfx

inline void idefixErrorImpl(const std::string &msg, const char *file, int line) {
  idfx::cout << std::endl
             << "------------------------------------------------------------------------------" << std::endl
             << "### FATAL ERROR in function " << file << " line " << line << std::endl
             << msg << std::endl
             << "------------------------------------------------------------------------------" << std::endl;
  idfx::safeExit(1);
}
inline std::string idefixMsgString(const std::string &s) { return s; }
inline std::string idefixMsgString(const std::stringstream &s) { return s.str(); }
#define IDEFIX_ERROR(msg) idefixErrorImpl(idefixMsgString(msg), __FILE__, __LINE__)
#define IDEFIX_WARNING(msg) (idfx::cout << "### WARNING: " << idefixMsgString(msg) << std::endl)
#define IDEFIX_DEPRECATED(msg) IDEFIX_WARNING(msg)
#define IDFX_CUDA(call)                                                                    \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess) IDEFIX_ERROR(std::string("CUDA: ") + cudaGetErrorString(e_));   \
  } while (0)
#define IDFX_ABI(call)                                  \
  do {                                                  \
    if ((call) != 0) IDEFIX_ERROR(idfx_last_error());   \
  } while (0)

#include "arrays.hpp"
#include "loop.hpp"
#include "input.hpp"
#include "grid.hpp"
#include "dataBlock.hpp"
#include "dataBlockHost.hpp"
#include "output.hpp"
#include "timeIntegrator.hpp"

#endif  // IDEFIX_B200_HOST_IDEFIX_HPP_
