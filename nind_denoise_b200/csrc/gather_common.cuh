// Shared between the stand-alone gather kernel (aux.cuh) and the fused first layer of igemm.cuh:
// crop indexing (tiler mirror + in-network pad) and the 64-wide bf16 im2col row of one pixel.
#pragma once
#include "ptx.cuh"

namespace nind {

// ------------------------------------------------------------------ crop gather + im2col
// Reference: OneImageDS.__getitem__ (src/nind_denoise/denoise_image.py:129-174) builds crop i as
//   crop[c, r, q] = img[c, sym(y0 + r, H), sym(x0 + q, W)]      (edge-INCLUSIVE mirror, np.flip)
// and the network then pads it itself:
//   UtNet: nn.ReflectionPad2d(2) (UtNet.py:27,98) — edge-EXCLUSIVE reflection of the crop's own pixels
//   UNet : Conv2d(padding=1) zero padding (ThirdPartyNets.py:66)
// This kernel fuses both with the im2col of the first 3x3 convolution (C_in = 3): for every output
// pixel of that convolution it writes a 64-channel bf16 vector
//   k in [0,27)  : hi part of tap element e = k       (e = (ky*3+kx)*3 + c)
//   k in [27,54) : lo part (x - bf16(x)) of element k-27   -> the input keeps ~16 bits of mantissa
//   k in [54,64) : 0
// so that the first layer runs as a K=64 per-pixel GEMM on the tensor cores.
struct GatherParams {
  const float* src;       // planar fp32
  long long src_img;      // floats between consecutive crops' source images (0 = one shared image)
  long long src_plane;    // floats between colour planes
  int src_w, src_h;       // image size
  const int2* origin;     // per crop (x0, y0) of the crop window in the image; null = (0, 0)
  int crop_h, crop_w;     // crop size fed to the network
  int pad;                // in-network padding of the first conv (UtNet: 2 reflect, UNet: 1 zero)
  int reflect;            // 1 = edge-exclusive reflection, 0 = zeros
  int out_h, out_w;       // first-conv output size (crop + 2*pad - 2)
  int n_crops;
  __nv_bfloat16* dst;     // [n_crops][out_h][out_w][64]
};

__device__ __forceinline__ int sym_index(int t, int n) {
  // edge-inclusive mirror; repeated until inside (images are far larger than the margins)
  t = t < 0 ? -t - 1 : t;
  t = t >= n ? 2 * n - 1 - t : t;
  return t < 0 ? 0 : t;
}

// 64 fp32 values of the im2col row of first-conv output pixel (b, y, x): [0,27) hi, [27,54) lo, rest 0.
__device__ __forceinline__ void im2col_row(const GatherParams& p, int b, int y, int x, float (&h)[64]) {
  int x0 = 0, y0 = 0;
  if (p.origin) {
    const int2 o = p.origin[b];
    x0 = o.x;
    y0 = o.y;
  }
  const float* img = p.src + b * p.src_img;
  long long rowoff[3];
  int col[3];
  bool vy[3], vx[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    int r = y + k - p.pad, q = x + k - p.pad;  // crop coordinates
    if (p.reflect) {
      r = r < 0 ? -r : (r >= p.crop_h ? 2 * p.crop_h - 2 - r : r);
      q = q < 0 ? -q : (q >= p.crop_w ? 2 * p.crop_w - 2 - q : q);
      vy[k] = vx[k] = true;
    } else {
      vy[k] = r >= 0 && r < p.crop_h;
      vx[k] = q >= 0 && q < p.crop_w;
    }
    rowoff[k] = (long long)sym_index(y0 + r, p.src_h) * p.src_w;
    col[k] = sym_index(x0 + q, p.src_w);
  }
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int e = (ky * 3 + kx) * 3 + c;
        float f = 0.f;
        if (vy[ky] && vx[kx]) f = __ldg(img + c * p.src_plane + rowoff[ky] + col[kx]);
        const float hi = __bfloat162float(__float2bfloat16_rn(f));
        h[e] = hi;
        h[27 + e] = f - hi;
      }
#pragma unroll
  for (int k = 54; k < 64; ++k) h[k] = 0.f;
}

}  // namespace nind
