// Kernel instantiations, part a: the igemm_kernel variants are spread over four translation units so that
// the library builds in parallel (each variant carries 16 epilogue instantiations).
#include "igemm_host.cuh"

namespace nind {
#define X NIND_IGEMM_DEFINE
X(6411, 64, 1, 1, false, false) X(6431, 64, 3, 1, false, false) X(64118, 64, 1, 1, true, false)
#undef X
}  // namespace nind
