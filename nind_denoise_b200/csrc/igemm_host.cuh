// Host side of the implicit-GEMM kernel: tensor-map encoding, tiling/pipeline choices, launch.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cstring>
#include <string>
#include "igemm.cuh"

namespace nind {

// NHWC bf16 activation buffer as stored (halo frames included in hs/ws).
struct ActBuf {
  __nv_bfloat16* ptr = nullptr;
  int b = 0, hs = 0, ws = 0, c = 0;
  size_t elems() const { return (size_t)b * hs * ws * c; }
};

struct ConvSpec {
  ActBuf in;
  int in_coff = 0, cin = 0;  // channels [in_coff, in_coff + cin) of `in`
  int taps = 9;              // 9: 3x3 valid conv over the stored buffer; 1: per-pixel GEMM
  const __nv_bfloat16* w = nullptr;  // packed [taps][n_total][cin], cin contiguous
  int n_total = 0;
  const float* bias = nullptr;
  int act = ACT_NONE;
  float slope = 0.f;
  int epi_mode = EPI_STORE;
  ActBuf out;  // EPI_STORE / EPI_D2S destination
  int out_coff = 0, out_halo = 0;
  int d2s_cout = 0;
  ActBuf pool;  // optional fused 2x2 max-pool destination (EPI_STORE), written at halo pool_halo, channel 0
  int pool_halo = 0;
  // EPI_HEAD
  const float* head_w = nullptr;
  const float* head_b = nullptr;
  float* head_out = nullptr;  // [B][3][hy][hx] fp32
  int head_unpad = 0, head_hy = 0, head_hx = 0, head_sigmoid = 0, head_clamp = 0;
  // tuning
  bool pair = false;  // EXPERIMENTAL pixel-pair mode (C_out = 64 3x3 layers as N = 128 GEMMs); `w` must then point to
                      // the weights re-packed by pack_pair_weights()
  int flat = -1;  // flat (1-D) tiles for narrow maps: -1 auto, 0 off, 1 force (error if illegal), 2 wherever legal
  int n_tile = 0;     // 0 = auto
  int cg = 0;         // CTAs per tile group: 0 = auto, 1, or 2 (cta_group::2 pair)
  bool c8 = false;    // first layer over the 8-channel padded-crop tensor (C_in = 3 as hi/lo bf16)
  int force_ws = -1;  // -1 = auto
  int dual = -1;      // two MMA issuer warps: -1 auto (on), 0 off, 1 on
  int wide = -1;      // 128-byte staging rows / one TMA store per 64 channels: -1 auto, 0 off, 1 on (where legal)
  int max_ctas = 0;   // 0 = number of SMs
};

struct IgemmLaunch {
  CUtensorMap tmA, tmB;
  CUtensorMap tmC4, tmC1;  // destination (TMA-store epilogue): 4-row and 1-row boxes of 8 pixels x 32 channels
  IgemmParams p;
  int n_tile = 0, tps = 1, cg = 1;
  bool c8 = false;  // first layer: 3x3 conv over an 8-channel (16 B/pixel) tensor, no-swizzle descriptors
  bool pdl = true;  // programmatic dependent launch: overlap this kernel's prologue with its predecessor's tail
  bool pm = false;  // pixel-pair mode
  size_t smem = 0;
  int grid = 0;
  double flops = 0;  // EXECUTED 2*MAC over the valid output pixels (see build_igemm)
};

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                        const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_tmapEncodeTiled tmap_encoder() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tmapEncodeTiled>(f);
  }
  return fn;
}

// bf16 tensor map, 128-byte swizzle, zero OOB fill.  dims/strides innermost first.
// swizzle: 128 (default), 64, or 0 (none)
inline bool encode_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims,
                             const uint64_t* strides_bytes /* rank-1 */, const uint32_t* box,
                             std::string* why, int swizzle = 128) {
  PFN_tmapEncodeTiled enc = tmap_encoder();
  if (!enc) {
    if (why) *why = "cuTensorMapEncodeTiled entry point not available";
    return false;
  }
  cuuint64_t gd[5], gs[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i + 1 < rank) gs[i] = strides_bytes[i];
  }
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs,
                   bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : (swizzle == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (why) {
      char buf[256];
      snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed: %d (rank %d dims %llu %llu %llu box %u %u %u)",
               (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
               (unsigned long long)(rank > 2 ? dims[2] : 0), box[0], rank > 1 ? box[1] : 0,
               rank > 2 ? box[2] : 0);
      *why = buf;
    }
    return false;
  }
  return true;
}

constexpr size_t IG_SMEM_LIMIT = 227 * 1024;

inline int device_sm_count() {
  static int n[64] = {0};  // per device ordinal
  int dev = 0;
  cudaGetDevice(&dev);
  dev = dev < 0 || dev >= 64 ? 0 : dev;
  if (!n[dev]) {
    cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
    if (n[dev] <= 0) n[dev] = 148;
  }
  return n[dev];
}

// Pixel-pair mode weights.  `w9` = the tap-major packing every 3x3 layer uses, [9][64][C] (host copy);
// result = [CTA rank 0..1][block g = chunk*3 + ky][96 rows][64 k]: rows 0..63 = this CTA's half of the
// N = 128 MMA's weight tile (pair tap j = 1-e; CTA rank a = output pixel a: kx = 2 - a for e = 0, 1 - a for
// e = 1), rows 64..95 = its half (c_out = 32*rank ..) of the N = 64 MMA's tile (pair tap j = e, kx = 2e).
// Chunks are ordered (input pixel e, 64-channel block).
inline void pack_pair_weights(const __nv_bfloat16* w9, int C, std::vector<__nv_bfloat16>* out) {
  const int sub_chunks = C / 64, kchunks = 2 * sub_chunks, blocks = kchunks * 3;
  out->assign((size_t)2 * blocks * 96 * 64, __float2bfloat16(0.f));
  auto W = [&](int ky, int kx, int co, int c) { return w9[((size_t)(ky * 3 + kx) * 64 + co) * C + c]; };
  for (int r = 0; r < 2; ++r)
    for (int kc = 0; kc < kchunks; ++kc) {
      const int e = kc / sub_chunks, cb = (kc % sub_chunks) * 64;
      for (int ky = 0; ky < 3; ++ky) {
        __nv_bfloat16* blk = out->data() + ((size_t)(r * blocks + kc * 3 + ky) * 96) * 64;
        const int kx128 = e == 0 ? 2 - r : 1 - r, kx64 = 2 * e;
        for (int co = 0; co < 64; ++co)
          for (int c = 0; c < 64; ++c) blk[(size_t)co * 64 + c] = W(ky, kx128, co, cb + c);
        for (int i = 0; i < 32; ++i)
          for (int c = 0; c < 64; ++c) blk[(size_t)(64 + i) * 64 + c] = W(ky, kx64, r * 32 + i, cb + c);
      }
    }
}

inline bool build_igemm(const ConvSpec& s, IgemmLaunch* L, std::string* why) {
  auto fail = [&](const char* m) {
    if (why) *why = m;
    return false;
  };
  if (s.taps != 9 && s.taps != 1) return fail("taps must be 1 or 9");
  if (!s.c8 && (s.cin % 64 != 0 || s.cin <= 0)) return fail("input channels must be a multiple of 64");
  if (s.c8 && (s.taps != 9 || s.cin != 8 || s.n_total != 64 || s.in.c != 8 || s.in_coff != 0))
    return fail("the 8-channel first-layer path needs taps=9, C=8, N=64");
  if (s.n_total % 64 != 0) return fail("output columns must be a multiple of 64");
  if (s.epi_mode == EPI_D2S && (s.d2s_cout % 64 != 0)) return fail("depth-to-space needs C_out multiple of 64");
  if (s.in.c % 8 != 0) return fail("buffer channel count must be a multiple of 8");
  // the epilogue addresses its destinations with 32-bit element offsets
  if (s.epi_mode != EPI_HEAD && s.out.elems() >= (1ull << 32)) return fail("destination buffer too large (>= 2^32 elements): use a smaller batch");
  if (s.pool.ptr && s.pool.elems() >= (1ull << 32)) return fail("pool buffer too large (>= 2^32 elements)");
  const int shrink = s.taps == 9 ? 2 : 0;
  IgemmParams& p = L->p;
  memset(&p, 0, sizeof p);
  memset(&L->tmC4, 0, sizeof L->tmC4);
  memset(&L->tmC1, 0, sizeof L->tmC1);

  if (s.pair && (s.taps != 9 || s.c8 || s.n_total != 64 || s.in_coff != 0 || s.cin != s.in.c ||
                 (s.cin != 64 && s.cin != 128) || (s.in.ws & 1) || s.epi_mode == EPI_D2S))
    return fail("pixel-pair mode needs a 3x3 layer with C_out = 64, C_in = 64|128 filling its buffer, even width");
  L->pm = s.pair;
  int n_tile = s.n_tile;
  if (s.pair) n_tile = 128;  // GEMM N = (pixel of the pair, c_out)
  if (n_tile == 0) n_tile = s.n_total >= 256 ? 256 : (s.n_total >= 128 ? 128 : 64);
  if (n_tile != 64 && n_tile != 128 && n_tile != 256) return fail("n_tile must be 64/128/256");
  if (s.epi_mode == EPI_HEAD && ((n_tile != 64 && !s.pair) || s.n_total != 64)) return fail("head needs N=64");
  L->n_tile = n_tile;

  p.w_valid = s.in.ws - shrink;
  p.h_valid = s.in.hs - shrink;
  p.hs_in = s.in.hs;
  p.rows_total = s.in.b * s.in.hs;
  // pair mode tiles 8 pixel PAIRS across
  const int tiles_x_real = s.pair ? (p.w_valid / 2 + IG_TILE_W - 1) / IG_TILE_W : (p.w_valid + IG_TILE_W - 1) / IG_TILE_W;
  int cg = s.cg;
  if (s.pair) cg = 2;
  // measured on B200 (profiles/r01_probe_cta_pair.log): the CTA pair wins on every 3x3 layer except
  // 64->128 (-3 %), and on the 1x1 / 2x2-s2 GEMMs only when K is large (C_in >= 512)
  if (s.c8) cg = 1;
  L->c8 = s.c8;
  if (cg == 0) {
    const bool pair = s.taps == 9 ? (s.cin >= 128 || s.n_total <= 64) : (s.cin >= 512);
    cg = pair ? 2 : 1;
  }
  if (cg != 1 && cg != 2) return fail("cg must be 1 or 2");
  L->cg = cg;
  // Flat tiles (128 consecutive pixels of the input's row-major index space) when the 8-pixel tile columns
  // would waste more than the (taps-1) garbage columns per row do; the 2-D box is limited to 256 rows.
  bool flat = false;
  {
    const int box_rows = s.taps == 9 ? 130 + 2 * s.in.ws : 128;
    const double eff_tile = (double)p.w_valid / (IG_TILE_W * tiles_x_real), eff_flat = (double)p.w_valid / s.in.ws;
    const bool can = !s.c8 && !s.pair && s.epi_mode != EPI_HEAD && !s.pool.ptr && box_rows <= 256;
    if (s.flat == 1 && !can) return fail("flat tiles need a narrow map (<= 63 px), no fused pool / head, not the first layer");
    flat = can && (s.flat >= 1 || (s.flat == -1 && eff_flat > eff_tile + 0.02));
  }
  // Pair the two tiles of a CTA pair along x when that wastes nothing (even tile count), else along y:
  // rows are batch-flattened (hundreds of tile rows), so a padded odd row count costs < 1 %.
  const int tiles_y_real = (p.rows_total - shrink + IG_TILE_H - 1) / IG_TILE_H;
  if (flat) {
    const long long n_flat = (long long)p.rows_total * s.in.ws;
    const int tiles = (int)((n_flat + 127) / 128);
    p.flat = 1;
    p.flat_pitch = s.in.ws;
    p.pair_y = cg == 2 ? 1 : 0;
    p.tiles_x = 1;
    p.tiles_y = (tiles + cg - 1) / cg;
  } else {
    p.pair_y = (cg == 2 && (tiles_x_real & 1)) ? 1 : 0;
    p.tiles_x = p.pair_y ? tiles_x_real : (tiles_x_real + cg - 1) / cg;
    p.tiles_y = p.pair_y ? (tiles_y_real + 1) / 2 : tiles_y_real;
  }
  p.tiles_n = s.pair ? 1 : (s.n_total + n_tile - 1) / n_tile;
  p.total_tiles = p.tiles_x * p.tiles_y * p.tiles_n;
  p.kchunks = s.c8 ? 1 : (s.pair ? 2 * s.cin / 64 : s.cin / 64);
  p.taps = s.taps;

  // A operand: one TMA box per (tile, 64-channel chunk).  3x3: the (16+2) x (8+2) pixel patch, whose
  // nine tap views are descriptor offsets (verified on B200: the UMMA 128B-swizzle XOR is applied to
  // absolute shared-memory address bits, so a start address moved by whole 128-byte rows is legal).
  uint32_t boxA[3];
  if (flat) {
    boxA[0] = 64; boxA[1] = s.taps == 9 ? 130 + 2 * s.in.ws : 128; boxA[2] = 1;
    p.a_tx_bytes = boxA[1] * 128; p.a_stage_bytes = (p.a_tx_bytes + 1023) & ~1023; p.a_sbo = 1024;
    p.tap_pitch16 = (uint32_t)s.in.ws * 8;
  } else if (s.pair) {  // 8 + 1 pixel pairs x 16 + 2 rows of the [rows, W/2, 2C] view
    boxA[0] = 64; boxA[1] = 9; boxA[2] = 18;
    p.a_tx_bytes = 9 * 18 * 128; p.a_stage_bytes = 21504; p.a_sbo = 1152;
  } else if (s.taps == 1) {
    boxA[0] = 64; boxA[1] = 8; boxA[2] = 16;
    p.a_tx_bytes = 16384; p.a_stage_bytes = 16384; p.a_sbo = 1024;
  } else if (s.c8) {
    boxA[0] = 8; boxA[1] = 10; boxA[2] = 18;
    p.a_tx_bytes = 10 * 18 * 16; p.a_stage_bytes = 3072; p.a_sbo = 160;
  } else {
    boxA[0] = 64; boxA[1] = 10; boxA[2] = 18;
    p.a_tx_bytes = 10 * 18 * 128; p.a_stage_bytes = 23552; p.a_sbo = 1280;
  }

  if (!flat) p.tap_pitch16 = p.a_sbo >> 4;

  // Store-bound layers (first layer, 2x2/s2 up-convs) hand the TMA unit 128-byte rows; the MMA-bound ones keep
  // the 64-byte halves, whose staging is half the size (the activation ring needs the shared memory more).
  const bool wide_ok = !flat && !s.pair && s.epi_mode != EPI_HEAD && !s.pool.ptr;
  const bool wide = wide_ok && (s.wide == 1 || (s.wide == -1 && (s.taps == 1 || s.c8)));
  p.wide = wide ? 1 : 0;

  // pipeline depth / weights-stationary decision
  const int tps = (s.c8 || s.pair) ? 1 : ((s.taps == 9 && n_tile <= 128) ? 3 : 1);  // taps per weight stage
  L->tps = tps;
  const int kt = s.c8 ? 2 : p.kchunks * (p.taps / tps);  // c8: 10 KB of weights in two 8 KB "stages"
  bool ws = false;
  if (s.force_ws != 0 && p.tiles_n == 1 && kt <= IG_MAX_STAGES &&
      igemm_smem_bytes(n_tile, tps, cg, 2, p.a_stage_bytes, kt, s.pair, wide) <= IG_SMEM_LIMIT)
    ws = true;
  if (s.force_ws == 1 && !ws) return fail("weights do not fit in shared memory");
  auto fits = [&](int sa, int sb) {
    return igemm_smem_bytes(n_tile, tps, cg, sa, p.a_stage_bytes, sb, s.pair, wide) <= IG_SMEM_LIMIT;
  };
  if (s.pair) {  // resident weights: 36 KB per (e, 64-channel block) per CTA, counted in 8 KB "stages"
    ws = true;
    p.sb = p.kchunks * 9 / 2;
    p.sa = 2;
    if (!fits(p.sa, p.sb)) return fail("pixel-pair weights do not fit in shared memory");
    while (p.sa < 4 && fits(p.sa + 1, p.sb)) ++p.sa;
  } else if (ws) {
    p.sb = kt;
    p.sa = 2;
    // The first layer's patches are 3 KB and its MMA phase per tile is ~200 clk, so the ring must cover the LOAD
    // LATENCY, not the consumption rate: with the HBM interface busy writing the 8x larger output, a patch arrives
    // ~5700 clk after its TMA is issued (profiles/r02_pipeline_trace_fine.log) — 6 stages capped the kernel at one
    // tile per ~950 clk.
    const int sa_max = s.c8 ? 24 : 6;
    while (p.sa < sa_max && fits(p.sa + 1, p.sb)) ++p.sa;
  } else {
    p.sb = 2;
    p.sa = 2;
    if (!fits(p.sa, p.sb)) return fail("pipeline does not fit in shared memory");
    const int sb_max = tps == 3 ? 4 : 8;
    while (p.sb < (tps == 3 ? 3 : 4) && fits(p.sa, p.sb + 1)) ++p.sb;
    while (p.sa < 3 && fits(p.sa + 1, p.sb)) ++p.sa;
    while (p.sb < sb_max && fits(p.sa, p.sb + 1)) ++p.sb;
    while (p.sa < 4 && fits(p.sa + 1, p.sb)) ++p.sa;
  }
  p.ws = ws ? 1 : 0;
  p.dual = s.dual != 0 ? 1 : 0;
  L->smem = igemm_smem_bytes(n_tile, tps, cg, p.sa, p.a_stage_bytes, p.sb, s.pair, wide);

  // tensor maps
  {
    const uint64_t dims[3] = {(uint64_t)s.cin, (uint64_t)s.in.ws, (uint64_t)p.rows_total};
    const uint64_t strides[2] = {(uint64_t)s.in.c * 2, (uint64_t)s.in.ws * s.in.c * 2};
    if (s.c8) {  // rows of (x, c) flattened: 10 px x 8 ch = 160 contiguous bytes per patch row
      const uint64_t dimsA[2] = {(uint64_t)s.in.ws * 8, (uint64_t)p.rows_total};
      const uint64_t stridesA[1] = {(uint64_t)s.in.ws * 16};
      const uint32_t boxA2[2] = {80, 18};
      if (!encode_tmap_bf16(&L->tmA, s.in.ptr, 2, dimsA, stridesA, boxA2, why, 0)) return false;
    } else if (s.pair) {  // [rows][W/2 pairs][2C channels]
      const uint64_t dimsA[3] = {(uint64_t)2 * s.cin, (uint64_t)s.in.ws / 2, (uint64_t)p.rows_total};
      const uint64_t stridesA[2] = {(uint64_t)s.in.c * 4, (uint64_t)s.in.ws * s.in.c * 2};
      if (!encode_tmap_bf16(&L->tmA, s.in.ptr, 3, dimsA, stridesA, boxA, why)) return false;
    } else if (flat) {  // [all pixels of the buffer][channels]
      const uint64_t dimsA[2] = {(uint64_t)s.cin, (uint64_t)p.rows_total * s.in.ws};
      const uint64_t stridesA[1] = {(uint64_t)s.in.c * 2};
      if (!encode_tmap_bf16(&L->tmA, s.in.ptr + s.in_coff, 2, dimsA, stridesA, boxA, why)) return false;
    } else if (!encode_tmap_bf16(&L->tmA, s.in.ptr + s.in_coff, 3, dims, strides, boxA, why)) {
      return false;
    }
    if (s.c8) {  // weights [10 half-K blocks][64 n][8 ch]
      const uint64_t dimsB[3] = {8, 64, 10};
      const uint64_t stridesB[2] = {16, 1024};
      const uint32_t boxB[3] = {8, 64, 10};
      if (!encode_tmap_bf16(&L->tmB, s.w, 3, dimsB, stridesB, boxB, why, 0)) return false;
    } else if (s.pair) {  // pack_pair_weights(): [2 ranks x blocks x 96 rows][64]
      const uint64_t dimsB[2] = {64, (uint64_t)2 * p.kchunks * 3 * 96};
      const uint64_t stridesB[1] = {128};
      const uint32_t boxB[2] = {64, 96};
      if (!encode_tmap_bf16(&L->tmB, s.w, 2, dimsB, stridesB, boxB, why)) return false;
    } else {
      const uint64_t dimsB[2] = {(uint64_t)s.cin, (uint64_t)s.taps * s.n_total};
      const uint64_t stridesB[1] = {(uint64_t)s.cin * 2};
      const uint32_t boxB[2] = {64, (uint32_t)(n_tile / cg)};
      if (!encode_tmap_bf16(&L->tmB, s.w, 2, dimsB, stridesB, boxB, why)) return false;
    }
  }

  // epilogue
  p.n_total = s.pair ? 128 : s.n_total;
  p.epi_mode = s.epi_mode;
  p.act = s.act;
  p.slope = s.slope;
  p.bias = s.bias;
  if (s.epi_mode == EPI_HEAD) {
    p.head_out = s.head_out;
    if (!s.head_w || !s.head_b) return fail("head needs weights and bias");
    if (cudaMemcpy(p.head_c, s.head_w, 192 * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(p.head_c + 192, s.head_b, 3 * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess)
      return fail("cannot read the head weights");
    p.head_c[195] = 0.f;
    p.h_unpad = s.head_unpad; p.h_size_y = s.head_hy; p.h_size_x = s.head_hx;
    p.h_row = s.head_hx; p.h_plane = (long long)s.head_hy * s.head_hx; p.h_img = 3 * p.h_plane;
    p.head_sigmoid = s.head_sigmoid;
    p.head_clamp = s.head_clamp;
  } else {
    const int up = s.epi_mode == EPI_D2S ? 2 : 1;
    if (s.out.hs < up * p.h_valid + 2 * s.out_halo || s.out.ws < up * p.w_valid + 2 * s.out_halo)
      return fail("destination buffer too small");
    const int cols = s.epi_mode == EPI_D2S ? s.d2s_cout : s.n_total;
    if (s.out_coff + cols > s.out.c) return fail("destination channel range out of bounds");
    if (s.out.b != s.in.b) return fail("batch mismatch");
    p.o_pix = s.out.c;
    p.o_row = (long long)s.out.ws * s.out.c;
    p.o_img = (long long)s.out.hs * p.o_row;
    p.out = s.out.ptr + (long long)s.out_halo * p.o_row + (long long)s.out_halo * p.o_pix + s.out_coff;
    p.d2s_cout = s.d2s_cout;
    p.c_coff = s.out_coff;
    if (!flat) {
      // Destination of the TMA-store epilogue: the interior of the output buffer as
      // [channel][sub-x][x][row][image]; "sub-x" is 1 for a plain store, the pixel of the pair in pair mode and
      // dx of the 2x2 block for depth-to-space (whose rows are 2y + dy).  The extents are the VALID ones, so the
      // tensor map clips tile overhang, the garbage rows between the images of the batch-flattened row space
      // and rows past the last image.
      const int sx = (s.pair || s.epi_mode == EPI_D2S) ? 2 : 1;
      const uint64_t wv = s.pair ? (uint64_t)p.w_valid / 2 : (uint64_t)p.w_valid;
      const uint64_t hv = s.epi_mode == EPI_D2S ? 2 * (uint64_t)p.h_valid : (uint64_t)p.h_valid;
      const uint64_t dimsC[5] = {(uint64_t)s.out.c, (uint64_t)sx, wv, hv, (uint64_t)s.out.b};
      const uint64_t stridesC[4] = {(uint64_t)p.o_pix * 2, (uint64_t)p.o_pix * 2 * sx, (uint64_t)p.o_row * 2,
                                    (uint64_t)p.o_img * 2};
      const uint32_t bc = wide ? 64 : 32;  // channels per box row
      const int sw = wide ? 128 : 64;
      const uint32_t box4[5] = {bc, 1, 8, 4, 1}, box1[5] = {bc, 1, 8, 1, 1};
      const __nv_bfloat16* baseC = s.out.ptr + (long long)s.out_halo * p.o_row + (long long)s.out_halo * p.o_pix;
      if (s.epi_mode == EPI_D2S && !(s.out.hs & 1)) {
        // Four input rows of one sub-pixel row dy are four output rows 2 apart: a single box needs
        // [channel][dx][x][dy][input row], and the image index is folded into the last dimension (stored rows per
        // image are even, so image b starts hs/2 row pairs after image b-1).  The folded dimension also covers the
        // halo rows between images, which the kernel never addresses: it only uses this map for warps whose four
        // rows are valid rows of ONE image, everything else goes row by row through tmC1.  (An odd number of stored
        // rows — UNet levels whose skip tensor has an odd size — keeps d2s_hs2 = 0: row-by-row stores only.)
        p.d2s_hs2 = s.out.hs / 2;
        const uint64_t dimsD[5] = {(uint64_t)s.out.c, 2, wv, 2, (uint64_t)s.out.b * p.d2s_hs2};
        const uint64_t stridesD[4] = {(uint64_t)p.o_pix * 2, (uint64_t)p.o_pix * 4, (uint64_t)p.o_row * 2,
                                      (uint64_t)p.o_row * 4};
        const uint32_t boxD[5] = {bc, 1, 8, 1, 4};
        if (!encode_tmap_bf16(&L->tmC4, baseC, 5, dimsD, stridesD, boxD, why, sw)) return false;
      } else if (!encode_tmap_bf16(&L->tmC4, baseC, 5, dimsC, stridesC, box4, why, sw)) {
        return false;
      }
      if (!encode_tmap_bf16(&L->tmC1, baseC, 5, dimsC, stridesC, box1, why, sw)) return false;
    }
    if (s.pool.ptr) {
      if (s.epi_mode != EPI_STORE) return fail("fused pooling needs a plain store epilogue");
      if ((p.h_valid | p.w_valid | p.hs_in) & 1) return fail("fused pooling needs even map sizes");
      if (s.pool.hs < p.h_valid / 2 + 2 * s.pool_halo || s.pool.ws < p.w_valid / 2 + 2 * s.pool_halo ||
          s.pool.c < s.n_total || s.pool.b != s.in.b)
        return fail("pooled destination buffer too small");
      p.pl_pix = s.pool.c;
      p.pl_row = (long long)s.pool.ws * s.pool.c;
      p.pl_img = (long long)s.pool.hs * p.pl_row;
      p.pool_out = s.pool.ptr + (long long)s.pool_halo * p.pl_row + (long long)s.pool_halo * p.pl_pix;
    }
  }
  {
    const int sms = s.max_ctas > 0 ? s.max_ctas : device_sm_count();
    const int groups_max = std::max(1, sms / cg);
    L->grid = cg * (p.total_tiles < groups_max ? p.total_tiles : groups_max);
  }
  // Executed work: every valid OUTPUT pixel x taps x C_in x C_out.  For the ConvTranspose2d 3x3 layers (valid
  // convolutions over a zero-framed input) this includes the taps that fall on the zero frame, so it is ~5 %
  // above SURVEY §8d's algorithmic count (ConvTranspose2d: INPUT pixels), which nind_denoise_b200/flops.py
  // restates and bench.py's roofline uses.
  L->flops = 2.0 * s.in.b * (double)p.h_valid * p.w_valid * s.n_total * (s.c8 ? 3 : s.cin) * s.taps;
  return true;
}

// One launcher per kernel variant.  The variants are spread over several translation units (igemm_inst_*.cu)
// so that nvcc compiles them in parallel; NIND_IGEMM_SINGLE_TU (tools/probe.cu) defines all of them here.
template <int N, int T, int CG, bool G = false, bool PM = false>
inline cudaError_t igemm_launch_t(const IgemmLaunch& L, const IgemmParams& p, cudaStream_t st) {
  static bool attr_done[64] = {false};  // cudaFuncSetAttribute is per device
  int dev = 0;
  cudaGetDevice(&dev);
  dev = dev < 0 || dev >= 64 ? 0 : dev;
  if (!attr_done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(igemm_kernel<N, T, CG, G, PM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)IG_SMEM_LIMIT);
    if (e != cudaSuccess) return e;
    attr_done[dev] = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(L.grid);
  cfg.blockDim = dim3(ig_threads(N, PM));
  cfg.dynamicSmemBytes = L.smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // see pdl_wait() in the kernel
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = L.pdl ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, igemm_kernel<N, T, CG, G, PM>, L.tmA, L.tmB, L.tmC4, L.tmC1, p);
}

#define NIND_IGEMM_VARIANTS(X)                                                                      \
  X(6411, 64, 1, 1, false, false) X(6431, 64, 3, 1, false, false) X(6412, 64, 1, 2, false, false)   \
  X(6432, 64, 3, 2, false, false) X(64118, 64, 1, 1, true, false)                                   \
  X(12811, 128, 1, 1, false, false) X(12831, 128, 3, 1, false, false) X(12812, 128, 1, 2, false, false) \
  X(12832, 128, 3, 2, false, false) X(128129, 128, 1, 2, false, true)                               \
  X(25611, 256, 1, 1, false, false) X(25612, 256, 1, 2, false, false)
#define NIND_IGEMM_DECLARE(key, N, T, CG, G, PM) \
  cudaError_t igemm_launch_##key(const IgemmLaunch& L, const IgemmParams& p, cudaStream_t st);
#define NIND_IGEMM_DEFINE(key, N, T, CG, G, PM)                                               \
  cudaError_t igemm_launch_##key(const IgemmLaunch& L, const IgemmParams& p, cudaStream_t st) { \
    return igemm_launch_t<N, T, CG, G, PM>(L, p, st);                                         \
  }
NIND_IGEMM_VARIANTS(NIND_IGEMM_DECLARE)
#ifdef NIND_IGEMM_SINGLE_TU
NIND_IGEMM_VARIANTS(NIND_IGEMM_DEFINE)
#endif

inline long long default_wait_cycles() {
  // bound of every mbarrier wait: NIND_WAIT_MS (default ~1 s at 2 GHz); under ncu replay, a debugger or
  // time-slicing a healthy kernel can need more
  static long long c = 0;
  if (!c) {
    const char* e = getenv("NIND_WAIT_MS");
    const double ms = e ? atof(e) : 0.0;
    c = ms > 0 ? (long long)(ms * 2.0e6) : (1ll << 31);
  }
  return c;
}

inline cudaError_t launch_igemm(const IgemmLaunch& L, int* err_flag, cudaStream_t st,
                                long long* trace = nullptr) {
  IgemmParams p = L.p;
  p.err = err_flag;
  p.trace = trace;
  p.wait_cycles = default_wait_cycles();
  if (L.c8) return igemm_launch_64118(L, p, st);
  if (L.pm) return igemm_launch_128129(L, p, st);
  const int key = L.n_tile * 100 + L.tps * 10 + L.cg;
  switch (key) {
    case 6411: return igemm_launch_6411(L, p, st);
    case 6431: return igemm_launch_6431(L, p, st);
    case 12811: return igemm_launch_12811(L, p, st);
    case 12831: return igemm_launch_12831(L, p, st);
    case 25611: return igemm_launch_25611(L, p, st);
    case 6412: return igemm_launch_6412(L, p, st);
    case 6432: return igemm_launch_6432(L, p, st);
    case 12812: return igemm_launch_12812(L, p, st);
    case 12832: return igemm_launch_12832(L, p, st);
    case 25612: return igemm_launch_25612(L, p, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace nind
