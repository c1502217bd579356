// Kernel instantiations, part b: the igemm_kernel variants are spread over four translation units so that
// the library builds in parallel (each variant carries 16 epilogue instantiations).
#include "igemm_host.cuh"

namespace nind {
#define X NIND_IGEMM_DEFINE
X(6412, 64, 1, 2, false, false) X(6432, 64, 3, 2, false, false) X(25611, 256, 1, 1, false, false)
#undef X
}  // namespace nind
