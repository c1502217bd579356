// Kernel instantiations, part c: the igemm_kernel variants are spread over four translation units so that
// the library builds in parallel (each variant carries 16 epilogue instantiations).
#include "igemm_host.cuh"

namespace nind {
#define X NIND_IGEMM_DEFINE
X(12811, 128, 1, 1, false, false) X(12831, 128, 3, 1, false, false) X(25612, 256, 1, 2, false, false)
#undef X
}  // namespace nind
