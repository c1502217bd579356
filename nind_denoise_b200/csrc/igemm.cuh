// Implicit-GEMM convolution for sm_100a: TMA -> shared memory -> tcgen05.mma -> TMEM -> fused epilogue.
//
// One kernel serves every dense layer of the NIND denoisers (reference:
// src/nind_denoise/networks/UtNet.py:27-88, ThirdPartyNets.py:62-136):
//   * 3x3 "valid" convolution over an NHWC bf16 buffer (Conv2d k=3; ConvTranspose2d k=3 s=1 and
//     Conv2d k=3 p=1 become valid convolutions because their inputs are stored with a zero frame),
//   * 1x1 convolution / per-pixel GEMM (the first layer after im2col, ConvTranspose2d k=2 s=2 with a
//     depth-to-space scatter epilogue),
//   * the 1x1 output head fused into the epilogue of the last 3x3 layer.
//
// GEMM view: M = output pixels (tile = 16 rows x 8 pixels = 128 UMMA rows), N = output channels,
// K = taps x input channels, walked as (64-channel chunk) x (tap).  The input patch of a tile
// ((16+2) x (8+2) pixels x 64 channels) is loaded ONCE per chunk by a 3-D TMA box with 128-byte
// swizzle; the nine taps are nine UMMA descriptors into the same patch (start address moved by
// (ky*10+kx) rows, 8-row groups 10 rows apart), so shared memory — not L2 — serves the 9x reuse.
//
// Warp roles (256 threads, 1 CTA/SM, persistent over tiles):
//   warp 0: TMA producer for activation patches      warp 1: TMA producer for weight tiles
//   warp 2: single-thread tcgen05.mma issuer         warp 3: TMEM allocator
//   warps 4-7: epilogue (tcgen05.ld -> bias/activation -> bf16 -> global), double-buffered TMEM.
#pragma once
#include "ptx.cuh"

namespace nind {

enum EpiMode : int { EPI_STORE = 0, EPI_D2S = 1, EPI_HEAD = 2 };
enum ActKind : int { ACT_NONE = 0, ACT_PRELU = 1, ACT_ELU = 2, ACT_HARDSWISH = 3 };

constexpr int IG_MAX_STAGES = 32;
constexpr int IG_BAR_BYTES = 2048;   // mbarriers + TMEM base slot
constexpr int IG_EPI_BYTES = 3072;   // staged bias [2][256] fp32 + head weights [3][64]+[3]
constexpr int IG_THREADS = 256;
constexpr int IG_TILE_H = 16;
constexpr int IG_TILE_W = 8;

struct IgemmParams {
  // tile grid
  int tiles_x, tiles_y, tiles_n, total_tiles;
  // K loop
  int kchunks, taps;
  // shared-memory pipeline geometry
  uint32_t a_stage_bytes;  // distance between A stages (multiple of 1024)
  uint32_t a_tx_bytes;     // bytes TMA delivers per A stage
  uint32_t a_sbo;          // bytes per patch row = distance between 8-row groups of the A operand
  int sa, sb, ws;          // stage counts; ws = weights stay resident in shared memory
  // output geometry
  int hs_in;               // stored rows per image of the input buffer
  int rows_total;          // images * hs_in
  int h_valid, w_valid;    // valid output rows per image / columns
  int n_total;             // valid output columns of the GEMM
  int epi_mode, act;
  float slope;
  const float* bias;
  __nv_bfloat16* out;      // already offset by halo and channel offset
  long long o_img, o_row;  // element strides
  int o_pix;
  int d2s_cout;            // EPI_D2S: channels per sub-pixel
  // EPI_HEAD: 1x1 conv to 3 channels (+ optional sigmoid), fp32 planar output
  const float* head_w;     // [3][64]
  const float* head_b;     // [3]
  float* head_out;
  long long h_img, h_plane;
  int h_row;
  int h_unpad;             // output pixel (y,x) -> (y-h_unpad, x-h_unpad)
  int h_size_y, h_size_x;  // output plane size
  int head_sigmoid;
  int* err;
};

__host__ __device__ inline size_t igemm_smem_bytes(int n_tile, int sa, uint32_t a_stage_bytes, int sb) {
  return 1024 + (size_t)sa * a_stage_bytes + (size_t)sb * n_tile * 128 + IG_BAR_BYTES + IG_EPI_BYTES;
}

template <int N_TILE>
__global__ void __launch_bounds__(IG_THREADS, 1)
igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const IgemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr uint32_t B_BYTES = N_TILE * 128;
  constexpr uint32_t TMEM_COLS = 2 * N_TILE;
  constexpr uint32_t IDESC = umma_idesc_bf16(128, N_TILE);

  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = sbase;
  const uint32_t b_base = a_base + (uint32_t)p.sa * p.a_stage_bytes;
  const uint32_t bar_base = b_base + (uint32_t)p.sb * B_BYTES;
  const uint32_t a_full = bar_base;
  const uint32_t a_empty = bar_base + 8 * IG_MAX_STAGES;
  const uint32_t b_full = bar_base + 16 * IG_MAX_STAGES;
  const uint32_t b_empty = bar_base + 24 * IG_MAX_STAGES;
  const uint32_t t_full = bar_base + 32 * IG_MAX_STAGES;
  const uint32_t t_empty = t_full + 16;
  const uint32_t tmem_slot = t_full + 32;
  const uint32_t epi_base = bar_base + IG_BAR_BYTES;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.sa; ++s) {
      mbar_init(a_full + 8 * s, 1);
      mbar_init(a_empty + 8 * s, 1);
    }
    for (int s = 0; s < p.sb; ++s) {
      mbar_init(b_full + 8 * s, 1);
      mbar_init(b_empty + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(t_full + 8 * s, 1);
      mbar_init(t_empty + 8 * s, 4);
    }
    mbar_fence_init();
  }
  if (warp == 3) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int tiles_xy = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    // ------------------------------------------------ activation-patch producer
    if (lane == 0) {
      uint32_t s = 0, ph = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int r = tile % tiles_xy;
        const int yt = r / p.tiles_x, xt = r - yt * p.tiles_x;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(a_empty + 8 * s, ph ^ 1, p.err, 1);
          mbar_arrive_expect_tx(a_full + 8 * s, p.a_tx_bytes);
          tma_load_3d(a_base + s * p.a_stage_bytes, &tmA, a_full + 8 * s, kc * 64, xt * IG_TILE_W,
                      yt * IG_TILE_H);
          if (++s == (uint32_t)p.sa) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ weight-tile producer
    if (lane == 0) {
      uint32_t s = 0, ph = 0;
      int tl = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tl) {
        if (p.ws && tl > 0) break;
        const int nt = tile / tiles_xy;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          for (int t = 0; t < p.taps; ++t) {
            if (!p.ws) mbar_wait(b_empty + 8 * s, ph ^ 1, p.err, 2);
            mbar_arrive_expect_tx(b_full + 8 * s, B_BYTES);
            tma_load_2d(b_base + s * B_BYTES, &tmB, b_full + 8 * s, kc * 64,
                        t * p.n_total + nt * N_TILE);
            if (++s == (uint32_t)p.sb) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------ MMA issuer
    // The whole warp walks the loop (warp-uniform control flow and addresses, so descriptors live
    // in uniform registers); one elected lane issues tcgen05.mma / tcgen05.commit.
    constexpr uint32_t DESC_HI_B = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t desc_hi_a = (p.a_sbo >> 4) | (1u << 14) | (2u << 29);
    const uint32_t tap_w = p.taps == 9 ? 3u : 1u;
    const uint32_t pitch16 = (p.a_sbo >> 4);  // one patch row, in 16-byte units
    uint32_t sa_i = 0, pha = 0, sb_i = 0, phb = 0;
    int tl = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tl) {
      const uint32_t acc = tl & 1, aph = (tl >> 1) & 1;
      mbar_wait(t_empty + 8 * acc, aph ^ 1, p.err, 3);
      tc_fence_after();
      const uint32_t d = tmem_base + acc * N_TILE;
      uint32_t accum = 0;
      for (int kc = 0; kc < p.kchunks; ++kc) {
        mbar_wait(a_full + 8 * sa_i, pha, p.err, 4);
        const uint32_t a_lo0 = (((a_base + sa_i * p.a_stage_bytes) >> 4) & 0x3FFF) | (1u << 16);
        for (uint32_t ky = 0; ky < tap_w; ++ky) {
          for (uint32_t kx = 0; kx < tap_w; ++kx) {
            if (!(p.ws && tl > 0)) mbar_wait(b_full + 8 * sb_i, phb, p.err, 5);
            tc_fence_after();
            const uint32_t a_lo = a_lo0 + ky * pitch16 + kx * 8;  // (ky*pitch + kx) rows of 128 B
            const uint32_t b_lo = (((b_base + sb_i * B_BYTES) >> 4) & 0x3FFF) | (1u << 16);
            if (elect_one_sync()) {
#pragma unroll
              for (uint32_t k = 0; k < 4; ++k) {
                umma_bf16_lohi(d, a_lo + 2 * k, desc_hi_a, b_lo + 2 * k, DESC_HI_B, IDESC, accum);
                accum = 1;
              }
              if (!p.ws) umma_commit(b_empty + 8 * sb_i);
            }
            __syncwarp();
            accum = 1;
            if (++sb_i == (uint32_t)p.sb) { sb_i = 0; phb ^= 1; }
          }
        }
        if (elect_one_sync()) umma_commit(a_empty + 8 * sa_i);
        __syncwarp();
        if (++sa_i == (uint32_t)p.sa) { sa_i = 0; pha ^= 1; }
      }
      if (elect_one_sync()) umma_commit(t_full + 8 * acc);
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ------------------------------------------------ epilogue
    const int quarter = warp & 3;
    const int etid = threadIdx.x - 128;  // 0..127 among the epilogue warps
    float* bias_s = reinterpret_cast<float*>(smem_raw + (epi_base - smem_u32(smem_raw)));  // [2][256]
    float* head_s = bias_s + 512;                                                          // [3][64] + [3]
    if (p.epi_mode == EPI_HEAD) {
      for (int i = etid; i < 195; i += 128) head_s[i] = i < 192 ? __ldg(p.head_w + i) : __ldg(p.head_b + i - 192);
    }
    int tl = 0, prev_nt = -1, bsel = 1;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tl) {
      const int nt = tile / tiles_xy;
      const int r = tile % tiles_xy;
      const int yt = r / p.tiles_x, xt = r - yt * p.tiles_x;
      const uint32_t acc = tl & 1, aph = (tl >> 1) & 1;

      if (nt != prev_nt) {  // (re)stage this N-tile's bias; uniform over the four epilogue warps
        prev_nt = nt;
        bsel ^= 1;
        for (int i = etid; i < N_TILE; i += 128) {
          const int n = nt * N_TILE + i;
          float bv = 0.f;
          if (n < p.n_total) bv = __ldg(p.bias + (p.epi_mode == EPI_D2S ? n % p.d2s_cout : n));
          bias_s[bsel * 256 + i] = bv;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }

      const int row = quarter * 32 + lane;
      const int yflat = yt * IG_TILE_H + (row >> 3);
      const int x = xt * IG_TILE_W + (row & 7);
      const int b = yflat / p.hs_in;
      const int y = yflat - b * p.hs_in;
      const bool valid = (yflat < p.rows_total) && (y < p.h_valid) && (x < p.w_valid);

      mbar_wait(t_full + 8 * acc, aph, p.err, 6);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * N_TILE;

      float h0 = 0.f, h1 = 0.f, h2 = 0.f;
#pragma unroll 1
      for (int c = 0; c < N_TILE / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + c * 32, v);
        tmem_wait_ld();
        const int n = nt * N_TILE + c * 32;
        if (n >= p.n_total) continue;
        __nv_bfloat16* dst;
        if (p.epi_mode == EPI_D2S) {
          const int q = n / p.d2s_cout;
          const int co = n - q * p.d2s_cout;
          dst = p.out + b * p.o_img + (long long)(2 * y + (q >> 1)) * p.o_row +
                (long long)(2 * x + (q & 1)) * p.o_pix + co;
        } else {
          dst = p.out + b * p.o_img + (long long)y * p.o_row + (long long)x * p.o_pix + n;
        }
        float f[32];
        const float4* bp = reinterpret_cast<const float4*>(bias_s + bsel * 256 + c * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bb = bp[j];
          f[4 * j + 0] = __uint_as_float(v[4 * j + 0]) + bb.x;
          f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + bb.y;
          f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + bb.z;
          f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + bb.w;
        }
        if (p.act == ACT_PRELU) {  // also ReLU (slope 0); branch-free
          const float sl = p.slope;
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f) + sl * fminf(f[j], 0.f);
        } else if (p.act == ACT_ELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = f[j] > 0.f ? f[j] : (__expf(f[j]) - 1.f);
        } else if (p.act == ACT_HARDSWISH) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = f[j] * fminf(fmaxf(f[j] + 3.f, 0.f), 6.f) * (1.f / 6.f);
        }
        if (p.epi_mode == EPI_HEAD) {
          const float* w0 = head_s + n;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            h0 = fmaf(f[j], w0[j], h0);
            h1 = fmaf(f[j], w0[64 + j], h1);
            h2 = fmaf(f[j], w0[128 + j], h2);
          }
        } else if (valid) {
          uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o;
            o.x = pack_bf16x2(f[8 * j + 0], f[8 * j + 1]);
            o.y = pack_bf16x2(f[8 * j + 2], f[8 * j + 3]);
            o.z = pack_bf16x2(f[8 * j + 4], f[8 * j + 5]);
            o.w = pack_bf16x2(f[8 * j + 6], f[8 * j + 7]);
            d4[j] = o;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(t_empty + 8 * acc);

      if (p.epi_mode == EPI_HEAD) {
        const int oy = y - p.h_unpad, ox = x - p.h_unpad;
        if (valid && oy >= 0 && ox >= 0 && oy < p.h_size_y && ox < p.h_size_x) {
          float o0 = h0 + head_s[192], o1 = h1 + head_s[193], o2 = h2 + head_s[194];
          if (p.head_sigmoid) {
            o0 = 1.f / (1.f + __expf(-o0));
            o1 = 1.f / (1.f + __expf(-o1));
            o2 = 1.f / (1.f + __expf(-o2));
          }
          float* ho = p.head_out + b * p.h_img + (long long)oy * p.h_row + ox;
          ho[0] = o0;
          ho[p.h_plane] = o1;
          ho[2 * p.h_plane] = o2;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 3) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace nind
