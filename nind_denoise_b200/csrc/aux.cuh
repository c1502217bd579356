// Memory-bound kernels around the convolutions: crop gather (+ in-network pad), 2x2 max-pool, the
// output-centric stitch and the file-format conversions either side of the path.
#pragma once
#include "gather_common.cuh"
#include "ptx.cuh"

namespace nind {

// ------------------------------------------------------------------ crop gather -> padded 8-channel tensor
// Writes the PADDED crop as 8 bf16 channels
// per pixel — R,G,B hi parts, R,G,B lo parts (x - bf16(x)), two zeros = 16 B — for the first layer's
// 3x3 implicit GEMM (igemm_kernel C8 mode).  out_h/out_w = crop + 2*pad.
__global__ void __launch_bounds__(256) gather_pad8_kernel(const GatherParams p) {
  const long long total = (long long)p.n_crops * p.out_h * p.out_w;
  for (long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x; gid < total;
       gid += (long long)gridDim.x * blockDim.x) {
    long long pix = gid;
    const int px = (int)(pix % p.out_w);
    pix /= p.out_w;
    const int py = (int)(pix % p.out_h);
    const int b = (int)(pix / p.out_h);
    int x0 = 0, y0 = 0;
    if (p.origin) {
      const int2 o = p.origin[b];
      x0 = o.x;
      y0 = o.y;
    }
    int r = py - p.pad, q = px - p.pad;  // crop coordinates
    bool inside = true;
    if (p.reflect) {
      r = r < 0 ? -r : (r >= p.crop_h ? 2 * p.crop_h - 2 - r : r);
      q = q < 0 ? -q : (q >= p.crop_w ? 2 * p.crop_w - 2 - q : q);
    } else {
      inside = r >= 0 && r < p.crop_h && q >= 0 && q < p.crop_w;
    }
    float hi[3] = {0.f, 0.f, 0.f}, lo[3] = {0.f, 0.f, 0.f};
    if (inside) {
      const float* img = p.src + b * p.src_img + (long long)sym_index(y0 + r, p.src_h) * p.src_w +
                         sym_index(x0 + q, p.src_w);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float f = __ldg(img + c * p.src_plane);
        hi[c] = __bfloat162float(__float2bfloat16_rn(f));
        lo[c] = f - hi[c];
      }
    }
    uint4 o;
    o.x = pack_bf16x2(hi[0], hi[1]);
    o.y = pack_bf16x2(hi[2], lo[0]);
    o.z = pack_bf16x2(lo[1], lo[2]);
    o.w = 0u;
    reinterpret_cast<uint4*>(p.dst)[gid] = o;
  }
}

// ------------------------------------------------------------------ 2x2 max-pool (NHWC bf16)
// nn.MaxPool2d(2) (UtNet.py:34, ThirdPartyNets.py:93).  `in`/`out` are already offset to the first
// interior pixel and the channel sub-range; strides in elements.
struct PoolParams {
  const __nv_bfloat16* in;
  long long i_img, i_row;
  int i_pix;
  __nv_bfloat16* out;
  long long o_img, o_row;
  int o_pix;
  int n, ho, wo, c;  // c multiple of 8
};

__global__ void __launch_bounds__(256) maxpool2_kernel(const PoolParams p) {
  const int c8 = p.c >> 3;
  const long long total = (long long)p.n * p.ho * p.wo * c8;
  for (long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x; gid < total;
       gid += (long long)gridDim.x * blockDim.x) {
    const int cc = (int)(gid % c8);
    long long r = gid / c8;
    const int xo = (int)(r % p.wo);
    r /= p.wo;
    const int yo = (int)(r % p.ho);
    const int b = (int)(r / p.ho);
    const __nv_bfloat16* ip = p.in + b * p.i_img + (long long)(2 * yo) * p.i_row + (long long)(2 * xo) * p.i_pix + cc * 8;
    const uint4 a = *reinterpret_cast<const uint4*>(ip);
    const uint4 bq = *reinterpret_cast<const uint4*>(ip + p.i_pix);
    const uint4 cq = *reinterpret_cast<const uint4*>(ip + p.i_row);
    const uint4 dq = *reinterpret_cast<const uint4*>(ip + p.i_row + p.i_pix);
    uint4 o;
    const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&bq);
    const __nv_bfloat162* pc = reinterpret_cast<const __nv_bfloat162*>(&cq);
    const __nv_bfloat162* pd = reinterpret_cast<const __nv_bfloat162*>(&dq);
    __nv_bfloat162* po = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i) po[i] = __hmax2(__hmax2(pa[i], pb[i]), __hmax2(pc[i], pd[i]));
    *reinterpret_cast<uint4*>(p.out + b * p.o_img + (long long)yo * p.o_row + (long long)xo * p.o_pix + cc * 8) = o;
  }
}

// ------------------------------------------------------------------ stitch
// Output-centric restatement of trim + make_seamless_edges + overlap-add
// (denoise_image.py:204-213, 250-267): every output pixel sums, in raster crop order, the
// contributions  w * net_out[crop][c][pad + y - ay][pad + x - ax]  of the crops whose useful area
// covers it, w in {1, 1/2, 1/4}.  No atomics; identical association order to the reference loop.
struct StitchParams {
  const float* crops;     // [crop_end - crop_begin][3][cs][cs] network outputs
  int crop_begin, crop_end;
  float* out;             // planar; row y of plane c is at out[c*out_plane + (y - out_y0)*W]
  long long out_plane;    // band output: (y_end-y_begin)*W with out_y0 = y_begin; whole image: H*W, 0
  int out_y0;
  int y_begin, y_end;     // rows written
  int W, H, cs, ucs, ol, pad, stride, nx, ny;
};

__global__ void __launch_bounds__(256) stitch_kernel(const StitchParams p) {
  const int band_h = p.y_end - p.y_begin;
  const long long total = (long long)band_h * p.W;
  const int wmax = p.cs - 2 * p.pad;  // widest useful area
  for (long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x; gid < total;
       gid += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(gid % p.W);
    const int yb = (int)(gid / p.W);
    const int y = yb + p.y_begin;
    int yi_lo = (y - wmax + p.stride) / p.stride;  // ceil((y - wmax + 1) / stride) for non-negative
    if (y - wmax + 1 <= 0) yi_lo = 0;
    int yi_hi = y / p.stride;
    if (yi_hi > p.ny - 1) yi_hi = p.ny - 1;
    int xi_lo = (x - wmax + p.stride) / p.stride;
    if (x - wmax + 1 <= 0) xi_lo = 0;
    int xi_hi = x / p.stride;
    if (xi_hi > p.nx - 1) xi_hi = p.nx - 1;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    for (int yi = yi_lo; yi <= yi_hi; ++yi) {
      const int ay = p.stride * yi;
      const int y1pad = max(0, ay - p.pad + p.cs - p.H);
      const int hcrop = p.cs - max(p.pad, y1pad) - p.pad;
      const int dy = y - ay;
      if (dy >= hcrop) continue;
      float wy = 1.f;
      if (ay != 0 && dy < p.ol) wy *= 0.5f;
      if (ay + p.ucs < p.H && p.ol && dy >= hcrop - p.ol) wy *= 0.5f;
      for (int xi = xi_lo; xi <= xi_hi; ++xi) {
        const int idx = yi * p.nx + xi;
        if (idx < p.crop_begin || idx >= p.crop_end) continue;
        const int ax = p.stride * xi;
        const int x1pad = max(0, ax - p.pad + p.cs - p.W);
        const int wcrop = p.cs - max(p.pad, x1pad) - p.pad;
        const int dx = x - ax;
        if (dx >= wcrop) continue;
        float w = wy;
        if (ax != 0 && dx < p.ol) w *= 0.5f;
        if (ax + p.ucs < p.W && p.ol && dx >= wcrop - p.ol) w *= 0.5f;
        const float* cp = p.crops + ((long long)(idx - p.crop_begin) * 3 * p.cs + (p.pad + dy)) * p.cs + p.pad + dx;
        const long long plane = (long long)p.cs * p.cs;
        s0 += w * __ldg(cp);
        s1 += w * __ldg(cp + plane);
        s2 += w * __ldg(cp + 2 * plane);
      }
    }
    const long long o = (long long)(y - p.out_y0) * p.W + x;
    p.out[o] = s0;
    p.out[p.out_plane + o] = s1;
    p.out[2 * p.out_plane + o] = s2;
  }
}

// ------------------------------------------------------------------ plain fp32 crop gather
// OneImageDS.__getitem__ (denoise_image.py:129-174) as a stand-alone op: crops[i][c][r][q] =
// img[c][sym(y0+r)][sym(x0+q)], bit-exact copies.  Used by the OneImageDS mirror and the parity tests.
struct CropGatherParams {
  const float* src;
  long long src_plane;
  int src_w, src_h;
  const int2* origin;
  int cs, n_crops;
  float* dst;  // [n_crops][3][cs][cs]
};

__global__ void __launch_bounds__(256) gather_crops_kernel(const CropGatherParams p) {
  const long long total = (long long)p.n_crops * 3 * p.cs * p.cs;
  for (long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x; gid < total;
       gid += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(gid % p.cs);
    long long r = gid / p.cs;
    const int rr = (int)(r % p.cs);
    r /= p.cs;
    const int c = (int)(r % 3);
    const int b = (int)(r / 3);
    const int2 o = p.origin[b];
    const int iy = sym_index(o.y + rr, p.src_h), ix = sym_index(o.x + q, p.src_w);
    p.dst[gid] = __ldg(p.src + c * p.src_plane + (long long)iy * p.src_w + ix);
  }
}

// ------------------------------------------------------------------ seam accumulation (multi-GPU)
// dst[i] += src[i]: a rank adds the partial sums another rank computed for image rows it owns (the grid row two
// crop ranges share, or the `ol` seam rows), in rank = raster order.
template <bool VEC>
__global__ void __launch_bounds__(256) add_rows_kernel(float* __restrict__ dst, long long dst_plane,
                                                       const float* __restrict__ src, long long src_plane, int planes,
                                                       long long count) {
  if (VEC) {  // count, the plane strides (in floats) and both pointers are multiples of 4 floats
    const long long n4 = count >> 2, total = n4 * planes;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
      const long long pl = i / n4, k = i - pl * n4;
      float4* d = reinterpret_cast<float4*>(dst + pl * dst_plane) + k;
      const float4 b = __ldg(reinterpret_cast<const float4*>(src + pl * src_plane) + k);
      float4 a = *d;
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      *d = a;
    }
  } else {
    const long long total = count * planes;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
      const long long pl = i / count, k = i - pl * count;
      dst[pl * dst_plane + k] += __ldg(src + pl * src_plane + k);
    }
  }
}

// ------------------------------------------------------------------ file formats either side of the path
// image_to_chw_kernel = img_path_to_np_flt after the decode (common/libs/np_imgops.py:19-28): interleaved HWC
// pixels as cv2 returns them (BGR when bgr = 1) of type u8 / u16 / f32 -> planar RGB fp32, x/255, x/65535 or
// unchanged (IEEE division: bit-identical to numpy's).
// chw_to_image_kernel = the quantisation of tensor_to_imgfile (common/libs/pt_helpers.py:24-32): u16
// clip(0,1)*65535 rounded half-to-even; u8 clip(0,1)*255 + 0.5 truncated (torchvision.utils.save_image);
// f32 unclamped.
struct PixConvParams {
  void* hwc;
  float* chw;
  int h, w;
  int dtype;  // 0 u8, 1 u16, 2 f32
  int bgr;
};

__global__ void __launch_bounds__(256) image_to_chw_kernel(const PixConvParams p) {
  const long long plane = (long long)p.h * p.w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < plane;
       i += (long long)gridDim.x * blockDim.x) {
    float v[3];
    if (p.dtype == 0) {
      const uint8_t* s = static_cast<const uint8_t*>(p.hwc) + 3 * i;
#pragma unroll
      for (int c = 0; c < 3; ++c) v[c] = __fdiv_rn((float)s[c], 255.f);
    } else if (p.dtype == 1) {
      const uint16_t* s = static_cast<const uint16_t*>(p.hwc) + 3 * i;
#pragma unroll
      for (int c = 0; c < 3; ++c) v[c] = __fdiv_rn((float)s[c], 65535.f);
    } else {
      const float* s = static_cast<const float*>(p.hwc) + 3 * i;
#pragma unroll
      for (int c = 0; c < 3; ++c) v[c] = s[c];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) p.chw[(p.bgr ? 2 - c : c) * plane + i] = v[c];
  }
}

__global__ void __launch_bounds__(256) chw_to_image_kernel(const PixConvParams p) {
  const long long plane = (long long)p.h * p.w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < plane;
       i += (long long)gridDim.x * blockDim.x) {
    float v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = p.chw[(p.bgr ? 2 - c : c) * plane + i];
    if (p.dtype == 0) {
      uint8_t* d = static_cast<uint8_t*>(p.hwc) + 3 * i;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float q = __fadd_rn(__fmul_rn(fminf(fmaxf(v[c], 0.f), 1.f), 255.f), 0.5f);
        d[c] = (uint8_t)fminf(fmaxf(q, 0.f), 255.f);
      }
    } else if (p.dtype == 1) {
      uint16_t* d = static_cast<uint16_t*>(p.hwc) + 3 * i;
#pragma unroll
      for (int c = 0; c < 3; ++c) d[c] = (uint16_t)rintf(__fmul_rn(fminf(fmaxf(v[c], 0.f), 1.f), 65535.f));
    } else {
      float* d = static_cast<float*>(p.hwc) + 3 * i;
#pragma unroll
      for (int c = 0; c < 3; ++c) d[c] = v[c];
    }
  }
}

}  // namespace nind
