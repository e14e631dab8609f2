#ifndef KOKKOS_SETUP_HPP_
#define KOKKOS_SETUP_HPP_
#endif
