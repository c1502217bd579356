// Crop indexing shared by the gather kernels (aux.cuh): tiler mirror + in-network pad.
#pragma once
#include "ptx.cuh"

namespace nind {

// ------------------------------------------------------------------ crop gather + im2col
// Reference: OneImageDS.__getitem__ (src/nind_denoise/denoise_image.py:129-174) builds crop i as
//   crop[c, r, q] = img[c, sym(y0 + r, H), sym(x0 + q, W)]      (edge-INCLUSIVE mirror, np.flip)
// and the network then pads it itself:
//   UtNet: nn.ReflectionPad2d(2) (UtNet.py:27,98) — edge-EXCLUSIVE reflection of the crop's own pixels
//   UNet : Conv2d(padding=1) zero padding (ThirdPartyNets.py:66)
// gather_pad8_kernel fuses both and writes the padded crop as 8 bf16 channels per pixel (RGB hi parts, RGB lo
// parts x - bf16(x), two zeros), so that the input keeps ~16 bits of mantissa and the first 3x3 convolution
// (C_in = 3) runs on the tensor cores as an implicit GEMM.
struct GatherParams {
  const float* src;       // planar fp32
  long long src_img;      // floats between consecutive crops' source images (0 = one shared image)
  long long src_plane;    // floats between colour planes
  int src_w, src_h;       // image size
  const int2* origin;     // per crop (x0, y0) of the crop window in the image; null = (0, 0)
  int crop_h, crop_w;     // crop size fed to the network
  int pad;                // in-network padding of the first conv (UtNet: 2 reflect, UNet: 1 zero)
  int reflect;            // 1 = edge-exclusive reflection, 0 = zeros
  int out_h, out_w;       // padded crop size (crop + 2*pad)
  int n_crops;
  __nv_bfloat16* dst;     // [n_crops][out_h][out_w][8]
};

__device__ __forceinline__ int sym_index(int t, int n) {
  // edge-inclusive mirror; repeated until inside (images are far larger than the margins)
  t = t < 0 ? -t - 1 : t;
  t = t >= n ? 2 * n - 1 - t : t;
  return t < 0 ? 0 : t;
}

}  // namespace nind
