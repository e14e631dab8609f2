/* Hand-written Kokkos 4.5 configuration for the CPU oracle build (OpenMP + Serial
 * back-ends, C++17, no TPLs).  It stands in for the header the reference's CMake
 * would generate from src/kokkos/cmake/KokkosCore_config.h.in; see oracle/ref_build/Makefile. */
#if !defined(KOKKOS_MACROS_HPP) || defined(KOKKOS_CORE_CONFIG_H)
#error "Do not include KokkosCore_config.h directly; include Kokkos_Macros.hpp instead."
#else
#define KOKKOS_CORE_CONFIG_H
#endif
#define KOKKOS_VERSION 40500
#define KOKKOS_VERSION_MAJOR 4
#define KOKKOS_VERSION_MINOR 5
#define KOKKOS_VERSION_PATCH 0
#define KOKKOS_ENABLE_SERIAL
#define KOKKOS_ENABLE_OPENMP
#define KOKKOS_ENABLE_CXX17
#define KOKKOS_ENABLE_COMPLEX_ALIGN
#define KOKKOS_ENABLE_IMPL_MDSPAN
#define KOKKOS_ENABLE_LIBDL
