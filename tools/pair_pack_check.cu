// Host-only check (no GPU needed) of the EXPERIMENTAL pixel-pair mode's weight packing and MMA schedule:
// emulates, with plain loops over pack_pair_weights()'s output, exactly the products the kernel issues
// (per chunk / ky: one N = 128 MMA at pair tap 1-e, one N = 64 MMA at pair tap e into accumulator half e) and
// compares with the direct 3x3 convolution.   nvcc -DNIND_PAIR_MODE=1 -o pair_pack_check tools/pair_pack_check.cu
#define NIND_PAIR_MODE 1
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../nind_denoise_b200/csrc/igemm_host.cuh"
using namespace nind;

int main() {
  int bad = 0;
  for (int C : {64, 128}) {
    const int H = 7, W = 12;  // input; valid output (H-2) x (W-2)
    std::vector<float> in((size_t)H * W * C), w((size_t)9 * 64 * C);
    srand(7 + C);
    for (auto& v : in) v = (rand() % 17 - 8) / 8.f;
    std::vector<__nv_bfloat16> w9(w.size());
    for (size_t i = 0; i < w.size(); ++i) { w[i] = (rand() % 15 - 7) / 16.f; w9[i] = __float2bfloat16(w[i]); }
    std::vector<__nv_bfloat16> wp;
    pack_pair_weights(w9.data(), C, &wp);
    const int sub_chunks = C / 64, kchunks = 2 * sub_chunks, blocks = kchunks * 3;
    auto Bv = [&](int r, int g, int row, int c) { return __bfloat162float(wp[(((size_t)r * blocks + g) * 96 + row) * 64 + c]); };
    auto A = [&](int y, int pair, int pc) {  // pair-view channel pc = e*C + c
      const int e = pc / C, c = pc % C, x = 2 * pair + e;
      return (y < H && x < W) ? in[((size_t)y * W + x) * C + c] : 0.f;
    };
    double maxerr = 0;
    for (int y = 0; y < H - 2; ++y)
      for (int P = 0; P < (W - 2) / 2; ++P) {
        std::vector<double> D(128, 0.0);
        for (int kc = 0; kc < kchunks; ++kc) {
          const int e = kc / sub_chunks;
          for (int ky = 0; ky < 3; ++ky) {
            const int g = kc * 3 + ky;
            for (int n = 0; n < 128; ++n)  // N = 128 MMA, pair tap j = 1-e; CTA rank n/64 supplies rows n%64
              for (int c = 0; c < 64; ++c) D[n] += A(y + ky, P + (1 - e), kc * 64 + c) * Bv(n / 64, g, n % 64, c);
            for (int n = 0; n < 64; ++n)   // N = 64 MMA, pair tap j = e, accumulator columns e*64 ..; rank n/32
              for (int c = 0; c < 64; ++c) D[e * 64 + n] += A(y + ky, P + e, kc * 64 + c) * Bv(n / 32, g, 64 + n % 32, c);
          }
        }
        for (int a = 0; a < 2; ++a)
          for (int co = 0; co < 64; ++co) {
            double ref = 0;
            for (int ky = 0; ky < 3; ++ky)
              for (int kx = 0; kx < 3; ++kx)
                for (int c = 0; c < C; ++c)
                  ref += in[((size_t)(y + ky) * W + (2 * P + a + kx)) * C + c] * w[((size_t)(ky * 3 + kx) * 64 + co) * C + c];
            const double err = std::fabs(ref - D[a * 64 + co]);
            if (err > maxerr) maxerr = err;
          }
      }
    printf("C=%d: max |pair schedule - 3x3 conv| = %.3e %s\n", C, maxerr, maxerr < 1e-9 ? "OK" : "MISMATCH");
    bad += maxerr >= 1e-9;
  }
  return bad;
}
