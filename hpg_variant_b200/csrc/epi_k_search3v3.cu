// epi_k_search3v3.cu -- instantiations of search3v3_kernel (epi_kernels.cuh), one translation unit per kernel family.
#include "epi_kernels.cuh"
#include "epi_launch.h"

namespace hpgv {

search_kernel_t kernel_search3v3(int bw, bool single) {
#define HPGV_VARIANT(BW, SINGLE) if (bw == BW && single == SINGLE) return (search_kernel_t) search3v3_kernel<BW, SINGLE, 5>
    HPGV_VARIANT(4, true);
    HPGV_VARIANT(7, true);
    HPGV_VARIANT(7, false);
    HPGV_VARIANT(8, true);
    HPGV_VARIANT(8, false);
#undef HPGV_VARIANT
    return nullptr;
}

}  // namespace hpgv
