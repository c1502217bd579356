// Stand-alone GPU probe (no torch): checks the UMMA shared-memory descriptor semantics the
// implicit-GEMM kernel relies on, then checks the kernel itself against a naive CUDA-core
// convolution.  Each invocation runs ONE test so that a trap in one cannot mask the others.
//
//   probe desc <r0> <sbo_bytes> <bo_mode>
//   probe tmem                                TMEM -> register read throughput per tcgen05.ld shape
//   probe hbm [GiB]                          read / write / copy ceilings of plain kernels
//   probe conv <taps> <cin> <n_total> <B> <Hs> <Ws> <a_mode> <bo_mode> <epi> [n_tile] [ws] [ctas]
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/probe tools/probe.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../nind_denoise_b200/csrc/igemm_host.cuh"

using namespace nind;

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                     \
    }                                                                              \
  } while (0)

static uint32_t rng_state = 12345u;
static float frand() {
  rng_state = rng_state * 1664525u + 1013904223u;
  return ((rng_state >> 8) & 0xFFFF) / 32768.0f - 1.0f;
}

// ------------------------------------------------------------------ descriptor probe
__global__ void __launch_bounds__(128, 1)
desc_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  float* out, int r0, uint32_t sbo, int bo_mode, int* err) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_s = sbase;               // 256 rows x 128 B
  const uint32_t b_s = sbase + 256 * 128;   // 64 rows x 128 B
  const uint32_t bar_ld = b_s + 64 * 128;
  const uint32_t bar_mma = bar_ld + 8;
  const uint32_t slot = bar_ld + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar_ld, 1);
    mbar_init(bar_mma, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(slot, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(slot));
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar_ld, 256 * 128 + 64 * 128);
    tma_load_2d(a_s, &tmA, bar_ld, 0, 0);
    tma_load_2d(b_s, &tmB, bar_ld, 0, 0);
    mbar_wait(bar_ld, 0, err, 1);
    tc_fence_after();
    const uint32_t a_addr = a_s + r0 * 128;
    const uint32_t bo = bo_mode ? ((a_addr >> 7) & 7) : 0;
    for (int k = 0; k < 4; ++k)
      umma_bf16(tmem_base, umma_desc_sw128(a_addr + 32 * k, sbo, bo), umma_desc_sw128(b_s + 32 * k, 1024),
                umma_idesc_bf16(128, 64), k > 0);
    umma_commit(bar_mma);
  }
  __syncwarp();
  mbar_wait(bar_mma, 0, err, 2);
  tc_fence_after();
  for (int c = 0; c < 2; ++c) {
    uint32_t v[32];
    tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, v);
    tmem_wait_ld();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 64);
}

static int run_desc(int r0, int sbo, int bo_mode) {
  const int R = 256;
  std::vector<__nv_bfloat16> hA(R * 64), hB(64 * 64);
  std::vector<float> fA(R * 64), fB(64 * 64);
  for (int i = 0; i < R * 64; ++i) { hA[i] = __float2bfloat16(frand()); fA[i] = __bfloat162float(hA[i]); }
  for (int i = 0; i < 64 * 64; ++i) { hB[i] = __float2bfloat16(frand()); fB[i] = __bfloat162float(hB[i]); }
  __nv_bfloat16 *dA, *dB;
  float* dO;
  int* dErr;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dO, 128 * 64 * 4));
  CK(cudaMalloc(&dErr, 4));
  CK(cudaMemset(dErr, 0, 4));
  CK(cudaMemset(dO, 0, 128 * 64 * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap tA, tB;
  std::string why;
  {
    uint64_t d[2] = {64, (uint64_t)R}, s[1] = {128};
    uint32_t bx[2] = {64, 256};
    if (!encode_tmap_bf16(&tA, dA, 2, d, s, bx, &why)) { printf("%s\n", why.c_str()); return 2; }
    uint64_t d2[2] = {64, 64};
    uint32_t bx2[2] = {64, 64};
    if (!encode_tmap_bf16(&tB, dB, 2, d2, s, bx2, &why)) { printf("%s\n", why.c_str()); return 2; }
  }
  const size_t smem = 1024 + 256 * 128 + 64 * 128 + 64;
  CK(cudaFuncSetAttribute(desc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  desc_probe_kernel<<<1, 128, smem>>>(tA, tB, dO, r0, (uint32_t)sbo, bo_mode, dErr);
  CK(cudaGetLastError());
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("DESC r0=%d sbo=%d bo=%d : kernel failed: %s\n", r0, sbo, bo_mode, cudaGetErrorString(e)); return 1; }
  std::vector<float> hO(128 * 64);
  CK(cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost));
  // expected under "absolute address" semantics
  double maxerr = 0;
  int bad = 0;
  for (int i = 0; i < 128; ++i) {
    const int arow = r0 + (i / 8) * (sbo / 128) + (i % 8);
    for (int n = 0; n < 64; ++n) {
      float acc = 0;
      for (int k = 0; k < 64; ++k) acc += fA[arow * 64 + k] * fB[n * 64 + k];
      const double d = fabs(acc - hO[i * 64 + n]);
      if (d > maxerr) maxerr = d;
      if (d > 1e-2) ++bad;
    }
  }
  printf("DESC r0=%d sbo=%d bo=%d : maxerr=%.5f bad=%d/8192 -> %s\n", r0, sbo, bo_mode, maxerr, bad,
         bad == 0 ? "PASS" : "FAIL");
  if (bad) {
    // Which A row does each output row actually correspond to?  (diagnostic)
    for (int i = 0; i < 24; ++i) {
      int best = -1;
      for (int r = 0; r < 256 && best < 0; ++r) {
        bool ok = true;
        for (int n = 0; n < 8 && ok; ++n) {
          float acc = 0;
          for (int k = 0; k < 64; ++k) acc += fA[r * 64 + k] * fB[n * 64 + k];
          if (fabs(acc - hO[i * 64 + n]) > 1e-2) ok = false;
        }
        if (ok) best = r;
      }
      printf("  out row %d matches A row %d (expected %d)\n", i, best, r0 + (i / 8) * (sbo / 128) + (i % 8));
    }
  }
  return bad ? 1 : 0;
}


// ------------------------------------------------------------------ SWIZZLE_NONE descriptor probe
// A: P "pixels" of 8 bf16 (16 B) contiguous in smem; K = 16 = two 8-wide halves LBO apart; 8-row core
// matrices (rows 16 B apart) SBO apart.  B: [2 halves][64 n][8] (LBO 1024, SBO 128).
__global__ void __launch_bounds__(128, 1)
desc0_probe_kernel(const __nv_bfloat16* gA, const __nv_bfloat16* gB, float* out, int r0, int lbo_px, int sbo_px,
                   int swap, int* err) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t a_s = sbase;              // 1024 pixels x 16 B
  const uint32_t b_s = sbase + 16384;      // 2 x 64 x 16 B
  const uint32_t bar_mma = b_s + 2048;
  const uint32_t slot = bar_mma + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 1024; i += 128) reinterpret_cast<uint4*>(sgen)[i] = reinterpret_cast<const uint4*>(gA)[i];
  for (int i = threadIdx.x; i < 128; i += 128) reinterpret_cast<uint4*>(sgen + 16384)[i] = reinterpret_cast<const uint4*>(gB)[i];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    mbar_init(bar_mma, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(slot, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(slot));
  if (threadIdx.x == 0) {
    auto desc = [&](uint32_t addr, uint32_t lbo, uint32_t sbo) {
      uint64_t d = 0;
      d |= (uint64_t)((addr >> 4) & 0x3FFF);
      d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
      d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
      d |= (uint64_t)1 << 46;  // version; layout type 0 = SWIZZLE_NONE
      return d;
    };
    const uint32_t la = lbo_px * 16, sa = sbo_px * 16;
    const uint64_t da = swap ? desc(a_s + r0 * 16, sa, la) : desc(a_s + r0 * 16, la, sa);
    const uint64_t db = swap ? desc(b_s, 128, 1024) : desc(b_s, 1024, 128);
    umma_bf16(tmem_base, da, db, umma_idesc_bf16(128, 64), 0);
    umma_commit(bar_mma);
  }
  __syncwarp();
  mbar_wait(bar_mma, 0, err, 2);
  tc_fence_after();
  for (int c = 0; c < 2; ++c) {
    uint32_t v[32];
    tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, v);
    tmem_wait_ld();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 64);
}

static int run_desc0(int r0, int lbo_px, int sbo_px, int swap) {
  const int P = 1024;
  std::vector<__nv_bfloat16> hA(P * 8), hB(2 * 64 * 8);
  std::vector<float> fA(P * 8), fB(2 * 64 * 8);
  for (int i = 0; i < P * 8; ++i) { hA[i] = __float2bfloat16(frand()); fA[i] = __bfloat162float(hA[i]); }
  for (int i = 0; i < 2 * 64 * 8; ++i) { hB[i] = __float2bfloat16(frand()); fB[i] = __bfloat162float(hB[i]); }
  __nv_bfloat16 *dA, *dB;
  float* dO;
  int* dErr;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dO, 128 * 64 * 4));
  CK(cudaMalloc(&dErr, 4));
  CK(cudaMemset(dErr, 0, 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  const size_t smem = 1024 + 16384 + 2048 + 64;
  CK(cudaFuncSetAttribute(desc0_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  desc0_probe_kernel<<<1, 128, smem>>>(dA, dB, dO, r0, lbo_px, sbo_px, swap, dErr);
  CK(cudaGetLastError());
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("DESC0 kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> hO(128 * 64);
  CK(cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  double maxerr = 0;
  for (int i = 0; i < 128; ++i)
    for (int n = 0; n < 64; ++n) {
      float acc = 0;
      for (int h = 0; h < 2; ++h) {
        const int px = r0 + (i / 8) * sbo_px + (i % 8) + h * lbo_px;
        for (int c = 0; c < 8; ++c) acc += fA[px * 8 + c] * fB[(h * 64 + n) * 8 + c];
      }
      const double d = fabs(acc - hO[i * 64 + n]);
      if (d > maxerr) maxerr = d;
      if (d > 1e-2) ++bad;
    }
  printf("DESC0 r0=%d lbo_px=%d sbo_px=%d swap=%d : maxerr=%.5f bad=%d/8192 -> %s\n", r0, lbo_px, sbo_px, swap, maxerr,
         bad, bad == 0 ? "PASS" : "FAIL");
  return bad ? 1 : 0;
}

// ------------------------------------------------------------------ naive reference conv
// in: NHWC bf16 [B][Hs][Ws][C] (channels coff..coff+cin used); w: [taps][n_total][cin] bf16;
// out: fp32 [B][Hv][Wv][n_total] after bias + activation.
__global__ void naive_conv_kernel(const __nv_bfloat16* in, int B, int Hs, int Ws, int C, int coff, int cin,
                                  const __nv_bfloat16* w, int taps, int n_total, const float* bias, int act,
                                  float slope, float* out) {
  const int tw = taps == 9 ? 3 : 1;
  const int Hv = Hs - (tw - 1), Wv = Ws - (tw - 1);
  const long long total = (long long)B * Hv * Wv * n_total;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = i % n_total;
    long long r = i / n_total;
    const int x = r % Wv; r /= Wv;
    const int y = r % Hv;
    const int b = r / Hv;
    float acc = 0.f;
    for (int t = 0; t < taps; ++t) {
      const int ky = t / tw, kx = t % tw;
      const __nv_bfloat16* ip = in + (((long long)b * Hs + y + ky) * Ws + x + kx) * C + coff;
      const __nv_bfloat16* wp = w + ((long long)t * n_total + n) * cin;
      for (int c = 0; c < cin; ++c) acc += __bfloat162float(ip[c]) * __bfloat162float(wp[c]);
    }
    acc += bias[n];
    if (act == ACT_PRELU) acc = acc > 0.f ? acc : acc * slope;
    out[i] = acc;
  }
}

static int run_conv(int argc, char** argv) {
  if (argc < 11) { printf("conv: not enough args\n"); return 2; }
  const int taps = atoi(argv[2]), cin = atoi(argv[3]), n_total = atoi(argv[4]), B = atoi(argv[5]),
            Hs = atoi(argv[6]), Ws = atoi(argv[7]), a_mode = atoi(argv[8]), bo_mode = atoi(argv[9]),
            epi = atoi(argv[10]);
  const int n_tile = argc > 11 ? atoi(argv[11]) : 0;
  const int ws = argc > 12 ? atoi(argv[12]) : -1;
  const int ctas = argc > 13 ? atoi(argv[13]) : 0;
  const int cg = argc > 14 ? atoi(argv[14]) : 0;
  const int flat = argc > 15 ? atoi(argv[15]) : -1;
  const int pair = argc > 16 ? atoi(argv[16]) : 0;   // pixel-pair mode (C_out = 64 layers)
  const int pool = argc > 17 ? atoi(argv[17]) : 0;   // fused 2x2 max-pool into a second buffer (EPI_STORE, even sizes)
  const int dual = argc > 18 ? atoi(argv[18]) : -1;  // two MMA issuer warps (-1 default = on)
  const int wide = argc > 19 ? atoi(argv[19]) : -1;  // 128-byte staging rows (-1 auto, 0 off, 1 on)
  const int tw = taps == 9 ? 3 : 1;
  const int Hv = Hs - (tw - 1), Wv = Ws - (tw - 1);
  const bool c8 = cin == 8;              // first-layer mode: 8-channel input, no-swizzle descriptors
  // read a channel sub-range to exercise offsets (pair mode views the whole buffer as [rows, W/2, 2C])
  const int Cbuf = c8 ? 8 : (pair ? cin : cin + 64), coff = (c8 || pair) ? 0 : 64;
  const int act = ACT_PRELU;
  const float slope = 0.25f;

  std::vector<__nv_bfloat16> hin((size_t)B * Hs * Ws * Cbuf), hw((size_t)taps * n_total * cin);
  for (auto& v : hin) v = __float2bfloat16(frand());
  const float wscale = 1.0f / sqrtf((float)(taps * cin));
  for (auto& v : hw) v = __float2bfloat16(frand() * wscale * 2.f);
  const int nbias = n_total;
  std::vector<float> hbias(nbias);
  for (auto& v : hbias) v = frand() * 0.5f;
  std::vector<float> hhw(3 * 64), hhb(3);
  for (auto& v : hhw) v = frand() * 0.2f;
  for (auto& v : hhb) v = frand() * 0.1f;

  __nv_bfloat16 *din, *dw, *dout = nullptr;
  float *dbias, *dref, *dhw, *dhb, *dhead = nullptr;
  int* derr;
  CK(cudaMalloc(&din, hin.size() * 2));
  CK(cudaMalloc(&dw, hw.size() * 2));
  CK(cudaMalloc(&dbias, hbias.size() * 4));
  CK(cudaMalloc(&dref, (size_t)B * Hv * Wv * n_total * 4));
  CK(cudaMalloc(&dhw, hhw.size() * 4));
  CK(cudaMalloc(&dhb, hhb.size() * 4));
  CK(cudaMalloc(&derr, 4));
  CK(cudaMemset(derr, 0, 4));
  CK(cudaMemcpy(din, hin.data(), hin.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dbias, hbias.data(), hbias.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dhw, hhw.data(), hhw.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dhb, hhb.data(), hhb.size() * 4, cudaMemcpyHostToDevice));

  ConvSpec s;
  s.in = ActBuf{din, B, Hs, Ws, Cbuf};
  __nv_bfloat16* dw8 = nullptr;
  if (c8) {  // repack [t][n][8] -> [10 half-K blocks][n][8] with blocks = taps 0..7, zero, tap 8
    std::vector<__nv_bfloat16> w8((size_t)10 * n_total * 8, __float2bfloat16(0.f));
    const int tap_of[10] = {0, 1, 2, 3, 4, 5, 6, 7, -1, 8};
    for (int blk = 0; blk < 10; ++blk)
      if (tap_of[blk] >= 0)
        for (int n = 0; n < n_total; ++n)
          for (int c = 0; c < 8; ++c) w8[((size_t)blk * n_total + n) * 8 + c] = hw[((size_t)tap_of[blk] * n_total + n) * 8 + c];
    CK(cudaMalloc(&dw8, w8.size() * 2));
    CK(cudaMemcpy(dw8, w8.data(), w8.size() * 2, cudaMemcpyHostToDevice));
    s.c8 = true;
  }
  __nv_bfloat16* dwp = nullptr;
  if (pair) {
    std::vector<__nv_bfloat16> wp;
    pack_pair_weights(hw.data(), cin, &wp);
    CK(cudaMalloc(&dwp, wp.size() * 2));
    CK(cudaMemcpy(dwp, wp.data(), wp.size() * 2, cudaMemcpyHostToDevice));
    s.pair = true;
  }
  s.in_coff = coff; s.cin = cin; s.taps = taps; s.w = c8 ? dw8 : (pair ? dwp : dw); s.n_total = n_total; s.bias = dbias;
  s.act = act; s.slope = slope; s.epi_mode = epi; (void)a_mode; (void)bo_mode;
  s.n_tile = n_tile; s.force_ws = ws; s.max_ctas = ctas; s.cg = cg; s.flat = flat; s.dual = dual; s.wide = wide;
  const int halo = 2, ocoff = 32;
  int Ho = 0, Wo = 0, Co = 0;
  const int unpad = 1;
  int d2s_cout = n_total / 4;
  if (epi == EPI_STORE) {
    Ho = Hv + 2 * halo; Wo = Wv + 2 * halo; Co = n_total + 64;
  } else if (epi == EPI_D2S) {
    Ho = 2 * Hv + 2 * halo; Wo = 2 * Wv + 2 * halo; Co = d2s_cout + 64;
    std::vector<float> b2(d2s_cout);
    for (auto& v : b2) v = frand() * 0.5f;
    // bias per co (repeat for the 4 sub-pixels in the reference)
    for (int n = 0; n < n_total; ++n) hbias[n] = b2[n % d2s_cout];
    CK(cudaMemcpy(dbias, hbias.data(), hbias.size() * 4, cudaMemcpyHostToDevice));
  }
  std::vector<__nv_bfloat16> hout;
  if (epi != EPI_HEAD) {
    hout.assign((size_t)B * Ho * Wo * Co, __float2bfloat16(-77.f));
    CK(cudaMalloc(&dout, hout.size() * 2));
    CK(cudaMemcpy(dout, hout.data(), hout.size() * 2, cudaMemcpyHostToDevice));
    s.out = ActBuf{dout, B, Ho, Wo, Co};
    s.out_coff = ocoff; s.out_halo = halo; s.d2s_cout = d2s_cout;
  } else {
    s.head_w = dhw; s.head_b = dhb; s.head_unpad = unpad; s.head_hy = Hv - 2 * unpad; s.head_hx = Wv - 2 * unpad;
    CK(cudaMalloc(&dhead, (size_t)B * 3 * s.head_hy * s.head_hx * 4));
    CK(cudaMemset(dhead, 0xFF, (size_t)B * 3 * s.head_hy * s.head_hx * 4));
    s.head_out = dhead;
  }

  __nv_bfloat16* dpool = nullptr;
  const int phalo = 1, Hp = Hv / 2 + 2 * phalo, Wp = Wv / 2 + 2 * phalo;
  std::vector<__nv_bfloat16> hpool;
  if (pool && epi == EPI_STORE) {
    hpool.assign((size_t)B * Hp * Wp * n_total, __float2bfloat16(-77.f));
    CK(cudaMalloc(&dpool, hpool.size() * 2));
    CK(cudaMemcpy(dpool, hpool.data(), hpool.size() * 2, cudaMemcpyHostToDevice));
    s.pool = ActBuf{dpool, B, Hp, Wp, n_total};
    s.pool_halo = phalo;
  }
  IgemmLaunch L;
  std::string why;
  if (!build_igemm(s, &L, &why)) { printf("build_igemm failed: %s\n", why.c_str()); return 2; }
  printf("CONV taps=%d cin=%d N=%d B=%d %dx%d a_mode=%d bo=%d epi=%d | n_tile=%d tiles=%d (x%d y%d n%d) grid=%d sa=%d sb=%d ws=%d smem=%zu\n",
         taps, cin, n_total, B, Hs, Ws, a_mode, bo_mode, epi, L.n_tile, L.p.total_tiles, L.p.tiles_x,
         L.p.tiles_y, L.p.tiles_n, L.grid, L.p.sa, L.p.sb, L.p.ws, L.smem); printf("  tps=%d cg=%d flat=%d wide=%d\n", L.tps, L.cg, L.p.flat, L.p.wide);

  naive_conv_kernel<<<1024, 256>>>(din, B, Hs, Ws, Cbuf, coff, cin, dw, taps, n_total, dbias, act, slope, dref);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());

  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(launch_igemm(L, derr, 0));
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    int herr = -1;
    printf("  igemm kernel failed: %s\n", cudaGetErrorString(e));
    (void)herr;
    return 1;
  }
  const int reps = 5;
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) CK(launch_igemm(L, derr, 0));
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= reps;
  printf("  time %.3f ms  -> %.1f TFLOP/s (useful)\n", ms, L.flops / ms * 1e-9);
  if (getenv("NIND_TRACE")) {
    long long* dtr;
    CK(cudaMalloc(&dtr, 64 * 16 * sizeof(long long)));
    CK(cudaMemset(dtr, 0, 64 * 16 * sizeof(long long)));
    CK(launch_igemm(L, derr, 0, dtr));
    CK(cudaDeviceSynchronize());
    std::vector<long long> tr(64 * 16);
    CK(cudaMemcpy(tr.data(), dtr, tr.size() * 8, cudaMemcpyDeviceToHost));
    const long long t0 = tr[0];
    printf("  trace (CTA 0, clk rel. to first A issue): tile | A_issue MMA_tempty MMA_afull MMA_done EPI_tfull EPI_tmem EPI_done"
           " | first half: ld0 staged fenced stored | second half: buf_free ld\n");
    for (int t = 0; t < 16; ++t) {
      printf("   %2d |", t);
      for (int e = 0; e < 13; ++e) printf(" %8lld", tr[t * 16 + e] ? tr[t * 16 + e] - t0 : -1);
      printf("\n");
    }
  }

  std::vector<float> href((size_t)B * Hv * Wv * n_total);
  CK(cudaMemcpy(href.data(), dref, href.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0;
  long long bad = 0, checked = 0;
  auto report = [&](const char* what, int b, int y, int x, int n, float got, float exp) {
    if (bad < 12) printf("  mismatch %s b=%d y=%d x=%d n=%d got=%f exp=%f\n", what, b, y, x, n, got, exp);
  };
  if (epi == EPI_HEAD) {
    std::vector<float> hh((size_t)B * 3 * s.head_hy * s.head_hx);
    CK(cudaMemcpy(hh.data(), dhead, hh.size() * 4, cudaMemcpyDeviceToHost));
    for (int b = 0; b < B; ++b)
      for (int c = 0; c < 3; ++c)
        for (int oy = 0; oy < s.head_hy; ++oy)
          for (int ox = 0; ox < s.head_hx; ++ox) {
            const float* rp = &href[(((size_t)b * Hv + oy + unpad) * Wv + ox + unpad) * n_total];
            float acc = hhb[c];
            for (int n = 0; n < 64; ++n) acc += rp[n] * hhw[c * 64 + n];
            const float got = hh[(((size_t)b * 3 + c) * s.head_hy + oy) * s.head_hx + ox];
            const double d = fabs(got - acc);
            ++checked;
            if (d > maxerr) maxerr = d;
            if (!(d <= 2e-3 + 2e-3 * fabs(acc))) { report("head", b, oy, ox, c, got, acc); ++bad; }
          }
  } else {
    CK(cudaMemcpy(hout.data(), dout, hout.size() * 2, cudaMemcpyDeviceToHost));
    std::vector<uint8_t> touched((size_t)B * Ho * Wo * Co, 0);
    for (int b = 0; b < B; ++b)
      for (int y = 0; y < Hv; ++y)
        for (int x = 0; x < Wv; ++x)
          for (int n = 0; n < n_total; ++n) {
            const float exp = href[(((size_t)b * Hv + y) * Wv + x) * n_total + n];
            size_t di;
            if (epi == EPI_STORE) {
              di = (((size_t)b * Ho + y + halo) * Wo + x + halo) * Co + ocoff + n;
            } else {
              const int q = n / d2s_cout, co = n % d2s_cout;
              di = (((size_t)b * Ho + 2 * y + (q >> 1) + halo) * Wo + 2 * x + (q & 1) + halo) * Co + ocoff + co;
            }
            touched[di] = 1;
            const float got = __bfloat162float(hout[di]);
            const double d = fabs(got - exp);
            ++checked;
            if (d > maxerr) maxerr = d;
            if (!(d <= 0.02 + 0.01 * fabs(exp))) { report("out", b, y, x, n, got, exp); ++bad; }
          }
    // everything else must be untouched (sentinel)
    long long stray = 0;
    for (size_t i = 0; i < hout.size(); ++i)
      if (!touched[i] && __bfloat162float(hout[i]) != -77.f) ++stray;
    if (stray) { printf("  %lld stray writes outside the valid region\n", stray); bad += stray; }
    if (dpool) {  // the pooled tensor must be exactly the 2x2 maximum of the stored (bf16) tensor
      CK(cudaMemcpy(hpool.data(), dpool, hpool.size() * 2, cudaMemcpyDeviceToHost));
      long long pbad = 0;
      for (int b = 0; b < B; ++b)
        for (int y = 0; y < Hp; ++y)
          for (int x = 0; x < Wp; ++x)
            for (int n = 0; n < n_total; ++n) {
              const float got = __bfloat162float(hpool[(((size_t)b * Hp + y) * Wp + x) * n_total + n]);
              float exp = -77.f;
              const int py = y - phalo, px = x - phalo;
              if (py >= 0 && py < Hv / 2 && px >= 0 && px < Wv / 2) {
                exp = -1e30f;
                for (int dy = 0; dy < 2; ++dy)
                  for (int dx = 0; dx < 2; ++dx)
                    exp = fmaxf(exp, __bfloat162float(hout[(((size_t)b * Ho + 2 * py + dy + halo) * Wo + 2 * px + dx + halo) * Co + ocoff + n]));
              }
              if (got != exp) { if (pbad < 6) printf("  pool mismatch b=%d y=%d x=%d n=%d got=%f exp=%f\n", b, y, x, n, got, exp); ++pbad; }
            }
      printf("  pool: %lld mismatches\n", pbad);
      bad += pbad;
    }
  }
  int herr = 0;
  CK(cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost));
  printf("  checked=%lld maxerr=%.5f bad=%lld err_flag=%d -> %s\n", checked, maxerr, bad, herr,
         bad == 0 ? "PASS" : "FAIL");
  return bad ? 1 : 0;
}

// ------------------------------------------------------------------ HBM ceilings (probe hbm [GiB])
// What a plain CUDA-core kernel reaches on this GPU for read-only, write-only, copy and the store patterns of the
// conv epilogues (64-byte halves of a 128-byte pixel written by different instructions; 128 of every 256 bytes =
// one channel half of a concat buffer).  MODE: 0 read, 1 write, 2 copy, 3 write in 64 B halves, 4 write 128 of 256 B,
// 5 read 1 : write 2 (the 2x2/s2 up-convs' mix).
template <int MODE>
__global__ void __launch_bounds__(256) hbm_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n16,
                                                  uint32_t* sink) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint4 acc = make_uint4(0, 0, 0, 0);
  const uint4 val = make_uint4(threadIdx.x, blockIdx.x, 3, 4);
  if (MODE == 0) {
    for (; i < n16; i += stride) { const uint4 v = __ldg(src + i); acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w; }
    if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345678u) *sink = 1;
  } else if (MODE == 1) {
    for (; i < n16; i += stride) dst[i] = val;
  } else if (MODE == 2) {
    for (; i < n16; i += stride) dst[i] = __ldg(src + i);
  } else if (MODE == 3) {
    // a warp covers 8 pixels x 128 B in two instructions: lanes (pixel = lane / 4, chunk = lane % 4), halves 0 then 1
    const size_t warps = stride / 32, w0 = i / 32;
    const int lane = threadIdx.x & 31;
    for (size_t w = w0; w * 64 + 63 < n16; w += warps) {
      uint4* base = dst + w * 64 + (size_t)(lane >> 2) * 8 + (lane & 3);
      base[0] = val;
      base[4] = val;
    }
  } else if (MODE == 4) {
    // 128 contiguous bytes of every 256: 16-byte chunk c of the dense index -> chunk (c / 8) * 16 + c % 8
    for (; i < n16 / 2; i += stride) dst[(i >> 3) * 16 + (i & 7)] = val;
  } else {
    for (; i < n16 / 2; i += stride) {
      const uint4 v = __ldg(src + i);
      dst[2 * i] = v;
      dst[2 * i + 1] = val;
    }
  }
}

static int run_hbm(double gib) {
  const size_t bytes = (size_t)(gib * (1ull << 30)), n16 = bytes / 16;
  uint4 *a, *b;
  uint32_t* sink;
  CK(cudaMalloc(&a, bytes));
  CK(cudaMalloc(&b, bytes));
  CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(a, 1, bytes));
  CK(cudaMemset(b, 2, bytes));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const char* names[6] = {"read only", "write only", "copy (read + write)", "write, 64 B halves of a 128 B pixel",
                          "write 128 of every 256 B", "read 1 : write 2"};
  for (int mode = 0; mode < 6; ++mode) {
    for (int per_sm : {4, 8}) {
      const int grid = sms * per_sm;
      float best = 1e30f;
      for (int it = 0; it < 6; ++it) {
        CK(cudaEventRecord(e0));
        switch (mode) {
          case 0: hbm_kernel<0><<<grid, 256>>>(a, b, n16, sink); break;
          case 1: hbm_kernel<1><<<grid, 256>>>(a, b, n16, sink); break;
          case 2: hbm_kernel<2><<<grid, 256>>>(a, b, n16, sink); break;
          case 3: hbm_kernel<3><<<grid, 256>>>(a, b, n16, sink); break;
          case 4: hbm_kernel<4><<<grid, 256>>>(a, b, n16, sink); break;
          default: hbm_kernel<5><<<grid, 256>>>(a, b, n16, sink); break;
        }
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (it > 0 && ms < best) best = ms;
      }
      // bytes that cross the HBM interface
      const double moved = mode == 2 ? 2.0 * bytes : (mode == 4 ? 0.5 * bytes : (mode == 5 ? 1.5 * bytes : 1.0 * bytes));
      printf("hbm %-38s %d CTAs/SM  %7.3f ms  %7.1f GB/s\n", names[mode], per_sm, best, moved / best / 1e6);
    }
  }
  // cudaMemsetAsync / cudaMemcpyAsync D2D for comparison (the driver's own kernels / copy engines)
  for (int k = 0; k < 2; ++k) {
    float best = 1e30f;
    for (int it = 0; it < 4; ++it) {
      CK(cudaEventRecord(e0));
      if (k == 0) CK(cudaMemsetAsync(b, 0, bytes));
      else CK(cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice));
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (it > 0 && ms < best) best = ms;
    }
    printf("hbm %-38s            %7.3f ms  %7.1f GB/s\n", k == 0 ? "cudaMemsetAsync" : "cudaMemcpyAsync D2D", best,
           (k == 0 ? 1.0 : 2.0) * bytes / best / 1e6);
  }
  return 0;
}

// ------------------------------------------------------------------ TMEM read throughput (probe tmem)
// How fast can the epilogue warps drain accumulators?  One CTA per SM, W warps (4, 8 or 16: 1, 2 or 4 per TMEM lane
// quarter), each loops tcgen05.ld of 4 KB + tcgen05.wait::ld over its quarter.  SHAPE 0: 32x32b.x32 (32 lanes x 32
// columns, what the epilogue uses), 1: 32x32b.x16 x 2, 2: 16x256b.x8 x 2 halves (16 lanes x 64 columns each),
// 3: 16x128b.x16 x 2 halves.  DEPTH = loads in flight before each wait.
#define TM_REGS32(v) "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), \
  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), \
  "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), \
  "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
#define TM_FMT32 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}"
template <int SHAPE>
__device__ __forceinline__ void tm_load4k(uint32_t taddr, uint32_t (&v)[32]) {
  if (SHAPE == 0) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " TM_FMT32 ", [%32];" : TM_REGS32(v) : "r"(taddr) : "memory");
  } else if (SHAPE == 1) {
    uint32_t(&a)[16] = reinterpret_cast<uint32_t(&)[16]>(v[0]);
    uint32_t(&b)[16] = reinterpret_cast<uint32_t(&)[16]>(v[16]);
    tmem_ld_32x32(taddr, a);
    tmem_ld_32x32(taddr + 16, b);
  } else if (SHAPE == 2) {  // two 16-lane halves x 32 columns each (x4 = 16 regs)
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr + (16u << 16)) : "memory");
  } else {  // 16x128b.x8: 16 lanes x 32 columns, 16 regs; two halves
    asm volatile("tcgen05.ld.sync.aligned.16x128b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.ld.sync.aligned.16x128b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr + (16u << 16)) : "memory");
  }
}

template <int SHAPE, int DEPTH>
__global__ void __launch_bounds__(512, 1) tmem_read_kernel(int iters, long long* clk, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    tmem_alloc(smem_u32(&slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t v[DEPTH][32];
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) tm_load4k<SHAPE>(base + ((it * DEPTH + d) * 32 & 511 & ~31), v[d]);
    tmem_wait_ld();
#pragma unroll
    for (int d = 0; d < DEPTH; ++d)
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= v[d][j];
  }
  const long long t1 = clock64();
  if (acc == 0x12345u) *sink = acc;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}

template <int SHAPE, int DEPTH>
static void run_tmem_one(const char* name, int warps, long long* dclk, uint32_t* sink, int sms) {
  const int iters = 2000;
  tmem_read_kernel<SHAPE, DEPTH><<<sms, warps * 32>>>(iters, dclk, sink);
  CK(cudaDeviceSynchronize());
  tmem_read_kernel<SHAPE, DEPTH><<<sms, warps * 32>>>(iters, dclk, sink);
  CK(cudaDeviceSynchronize());
  std::vector<long long> h(sms);
  CK(cudaMemcpy(h.data(), dclk, sms * 8, cudaMemcpyDeviceToHost));
  double avg = 0;
  for (auto c : h) avg += (double)c;
  avg /= sms;
  printf("tmem %-22s depth %d  %2d warps: %7.1f B/clk/SM  (%.0f clk per 4 KB load per warp)\n", name, DEPTH, warps,
         (double)warps * iters * DEPTH * 4096.0 / avg, avg / (iters * DEPTH));
}

static int run_tmem() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  long long* dclk;
  uint32_t* sink;
  CK(cudaMalloc(&dclk, sms * 8));
  CK(cudaMalloc(&sink, 4));
  for (int warps : {4, 8, 16}) {
    run_tmem_one<0, 1>("32x32b.x32", warps, dclk, sink, sms);
    run_tmem_one<0, 2>("32x32b.x32", warps, dclk, sink, sms);
    run_tmem_one<1, 1>("32x32b.x16 x2", warps, dclk, sink, sms);
    run_tmem_one<2, 1>("16x256b.x4 x2", warps, dclk, sink, sms);
    run_tmem_one<2, 2>("16x256b.x4 x2", warps, dclk, sink, sms);
    run_tmem_one<3, 1>("16x128b.x8 x2", warps, dclk, sink, sms);
  }
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 2) { printf("usage: probe desc|conv ...\n"); return 2; }
  if (!strcmp(argv[1], "desc")) {
    if (argc < 5) return 2;
    return run_desc(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]));
  }
  if (!strcmp(argv[1], "desc0")) {
    if (argc < 6) return 2;
    return run_desc0(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]));
  }
  if (!strcmp(argv[1], "conv")) return run_conv(argc, argv);
  if (!strcmp(argv[1], "tmem")) return run_tmem();
  if (!strcmp(argv[1], "hbm")) return run_hbm(argc > 2 ? atof(argv[2]) : 4.0);
  return 2;
}
