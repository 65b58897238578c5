// epi_k_search2_tri.cu -- the tri-layout variants of search2_kernel, compiled with the step sequencing of the producer lane
// in shared memory (HPGV_PS_IN_SMEM, see search2_kernel): at 96 registers and 20 warps the dozen registers it frees in the
// counting loop are worth 3.5 % on c2; the other variants (epi_k_search2.cu) keep it in locals, which is faster for them.
#define HPGV_PS_IN_SMEM 1
#include "epi_kernels.cuh"
#include "epi_launch.h"

namespace hpgv {

search_kernel_t kernel_search2_tri(bool balanced) {
    return balanced ? (search_kernel_t) search2_kernel<3, true, true> : (search_kernel_t) search2_kernel<3, true, false>;
}

}  // namespace hpgv
