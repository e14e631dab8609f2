#ifndef KOKKOS_FWD_HPP_
#define KOKKOS_FWD_HPP_
#include <fwd/Kokkos_Fwd_OPENMP.hpp>
#include <fwd/Kokkos_Fwd_SERIAL.hpp>
#endif
