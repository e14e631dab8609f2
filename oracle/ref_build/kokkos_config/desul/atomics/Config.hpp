/* desul configuration for the CPU oracle build: host (GCC builtin) atomics only */
#ifndef DESUL_ATOMICS_CONFIG_HPP_
#define DESUL_ATOMICS_CONFIG_HPP_
#endif
