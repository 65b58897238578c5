// epi_k_search2.cu -- instantiations of search2_kernel (epi_kernels.cuh), one translation unit per kernel family.
#include "epi_kernels.cuh"
#include "epi_launch.h"

namespace hpgv {

search_kernel_t kernel_search2_tri(bool balanced);          // epi_k_search2_tri.cu

search_kernel_t kernel_search2(int bw, bool single, bool balanced) {
    if (bw == 3) return kernel_search2_tri(balanced);
#define HPGV_VARIANT(BW, SINGLE) if (bw == BW && single == SINGLE) return balanced ? (search_kernel_t) search2_kernel<BW, SINGLE, true> : (search_kernel_t) search2_kernel<BW, SINGLE, false>
    HPGV_VARIANT(4, true);
    HPGV_VARIANT(7, true);
    HPGV_VARIANT(7, false);
    HPGV_VARIANT(8, true);
    HPGV_VARIANT(8, false);
#undef HPGV_VARIANT
    return nullptr;
}

}  // namespace hpgv
