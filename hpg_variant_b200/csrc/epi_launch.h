// epi_launch.h -- the search kernels are instantiated in one translation unit per kernel family (epi_k_*.cu), compiled in
// parallel by build.py; epi_capi.cu gets the kernel it needs for a layout through these getters (nullptr: no such variant).
#pragma once
#include "epi_types.h"

namespace hpgv {

typedef void (*search_kernel_t)(const SearchArgs);

// bw: 3 (tri layout, order 2 only), 4, 7 (8-word slots, seven words used), 8
search_kernel_t kernel_search2(int bw, bool single, bool balanced);
search_kernel_t kernel_search3(int bw, bool single, bool balanced);
search_kernel_t kernel_search3v2(int bw, bool single, bool balanced);
search_kernel_t kernel_search3v3(int bw, bool single);        // balanced cohorts, up to 5 counter words per cell

}  // namespace hpgv
