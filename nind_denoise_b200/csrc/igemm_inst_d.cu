// Kernel instantiations, part d: the igemm_kernel variants are spread over four translation units so that
// the library builds in parallel (each variant carries 16 epilogue instantiations).
#include "igemm_host.cuh"

namespace nind {
#define X NIND_IGEMM_DEFINE
X(12812, 128, 1, 2, false, false) X(12832, 128, 3, 2, false, false) X(128129, 128, 1, 2, false, true)
#undef X
}  // namespace nind
