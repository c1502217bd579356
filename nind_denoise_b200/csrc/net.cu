// C ABI (include/nind_b200.h): weight packing, per-shape execution plans, tiled denoise driver.
//
// Reference being replaced (paths relative to /root/reference):
//   networks/UtNet.py:14-109, networks/ThirdPartyNets.py:62-169   network definitions
//   nn_common.py:116-138                                           factory + load_state_dict
//   denoise_image.py:88-174, 204-213, 240-267                      crop grid, gather, stitch
#include <atomic>
#include <cmath>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/nind_b200.h"
#include "aux.cuh"
#include "igemm_host.cuh"

using namespace nind;

namespace {

thread_local std::string g_err;
std::atomic<int64_t> g_launches{0};

int fail(int code, const std::string& m) {
  g_err = m;
  return code;
}
#define CUDA_TRY(x)                                                                          \
  do {                                                                                       \
    cudaError_t e_ = (x);                                                                    \
    if (e_ != cudaSuccess) return fail(NIND_E_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_)); \
  } while (0)

struct PackedLayer {
  int taps = 9, cin = 0, n_total = 0, cout = 0;
  __nv_bfloat16* w = nullptr;  // device [taps][n_total][cin]
  __nv_bfloat16* w_pair = nullptr;  // device, pack_pair_weights() layout (3x3 layers with C_out = C_in = 64)
  float* bias = nullptr;       // device
  int act = ACT_NONE;
  float slope = 0.f;
};

enum StepKind { STEP_GATHER = 0, STEP_IGEMM = 1, STEP_POOL = 2 };

struct Step {
  int kind = STEP_IGEMM;
  std::string name;
  IgemmLaunch ig;
  GatherParams g;
  PoolParams pl;
  double flops = 0;
  double bytes = 0;  // algorithmic bytes moved (memory-bound steps)
};

struct Plan {
  int b = 0, h = 0, w = 0;
  size_t bytes = 0;          // activation arena owned by this plan
  unsigned long long last_use = 0;
  std::vector<void*> bufs;
  std::vector<Step> steps;
  int gather_step = -1, head_step = -1;
  ~Plan() {
    for (void* p : bufs) cudaFree(p);
  }
};

struct GridGeom {
  int W, H, cs, ucs, ol, pad, stride, nx, ny;
  int size() const { return nx * ny; }
};

// denoise_image.py:100-104.  All integer; ceil of a quotient of ints done exactly.
bool make_grid(int W, int H, int cs, int ucs, int ol, GridGeom* g) {
  if (W <= 0 || H <= 0 || cs <= 0 || ucs <= 0 || ol < 0 || ucs > cs || ucs - ol <= 0) return false;
  g->W = W; g->H = H; g->cs = cs; g->ucs = ucs; g->ol = ol;
  g->stride = ucs - ol;
  g->pad = (cs - ucs) / 2;
  auto ceil_div = [](int a, int b) { return a >= 0 ? (a + b - 1) / b : -((-a) / b); };
  g->nx = ceil_div(W - ucs, g->stride) + 1;
  g->ny = ceil_div(H - ucs, g->stride) + 1;
  return g->nx >= 1 && g->ny >= 1;
}

void crop_entry(const GridGeom& g, int i, nind_crop* c) {
  const int yi = i / g.nx, xi = i - yi * g.nx;
  const int x0 = g.stride * xi - g.pad, y0 = g.stride * yi - g.pad;
  const int x1pad = std::max(0, x0 + g.cs - g.W), y1pad = std::max(0, y0 + g.cs - g.H);
  c->x0 = x0; c->y0 = y0;
  c->ud_x0 = g.pad; c->ud_y0 = g.pad;
  c->ud_x1 = g.cs - std::max(g.pad, x1pad);
  c->ud_y1 = g.cs - std::max(g.pad, y1pad);
  c->start_x = x0 + g.pad; c->start_y = y0 + g.pad;
}

void band_of(const GridGeom& g, int cb, int ce, int* y0, int* y1) {
  nind_crop a, b;
  crop_entry(g, cb, &a);
  crop_entry(g, ce - 1, &b);
  *y0 = a.start_y;
  *y1 = std::min(g.H, b.start_y + (b.ud_y1 - b.ud_y0));
}

}  // namespace

struct nind_net {
  int arch = 0, funit = 64, act_kind = ACT_PRELU, device = 0;
  std::map<std::string, PackedLayer> layers;
  float* head_w = nullptr;
  float* head_b = nullptr;
  int* err_flag = nullptr;      // device alias of err_host
  int* err_host = nullptr;      // mapped, page-locked: survives a kernel trap
  cudaEvent_t ev_last = nullptr; // last use of the shared scratch (plan arenas, crops_buf, origin_buf)
  std::map<std::vector<int>, std::unique_ptr<Plan>> plans;
  unsigned long long plan_tick = 0;
  // tiled driver scratch
  float* crops_buf = nullptr;
  size_t crops_cap = 0;
  int2* origin_buf = nullptr;
  size_t origin_cap = 0;
  std::vector<int> origin_key;  // geometry + range the device table currently holds
  // host-buffer pipeline: two (device image, device output) slots so that consecutive images overlap
  struct HostSlot {
    float* img = nullptr;
    float* out = nullptr;
    size_t img_cap = 0, out_cap = 0;
    cudaEvent_t img_free = nullptr, out_free = nullptr;  // last compute read of img / last D2H read of out
    bool used = false;
  };
  HostSlot slots[2];
  unsigned host_seq = 0;
  cudaStream_t s_in = nullptr, s_comp = nullptr, s_out = nullptr;
  std::vector<cudaEvent_t> ev_in, ev_done;
  std::vector<cudaEvent_t> ev_rows;   // host-range pipeline: recorded on s_comp after each step's stitch
  std::vector<int> rows_done;         // band rows [y0, rows_done[k]) are final after step k (last enqueue)
  cudaEvent_t ev_join = nullptr;
  int flat = -1;
  int wide = -1;   // 128-byte staging rows in the TMA-store epilogue: -1 auto (store-bound layers), 0 off, 1 wherever legal
  int pair64 = 1;  // pixel-pair mode for the C_out = 64 3x3 layers (validated on B200, profiles/r02_pair_mode_first_light.log)
  int host_first = -1, host_last = -1;  // crops in the first / last pipeline step (-1: one grid row)
  // options
  int n_tile_deep = 256, max_ctas = 0, cg = 0, fuse_pool = 1, dual = -1;
  // Programmatic dependent launch of the conv kernels: measured on B200 (profiles/r02_pdl_ab.log) +0.2 % at the
  // default batch and -6 % with 28-crop forwards, so it is off; "pdl" = 1 switches it on.
  int pdl = 0;
  // timing
  int timing = 0;
  std::vector<std::string> t_names;
  std::vector<float> t_ms;
  std::vector<double> t_flops, t_bytes;

  void free_layers() {
    for (auto& kv : layers) {
      cudaFree(kv.second.w);
      cudaFree(kv.second.w_pair);
      cudaFree(kv.second.bias);
    }
    layers.clear();
    cudaFree(head_w);
    cudaFree(head_b);
    head_w = head_b = nullptr;
  }
  ~nind_net() {
    plans.clear();
    free_layers();
    if (err_host) cudaFreeHost(err_host);
    if (ev_last) cudaEventDestroy(ev_last);
    cudaFree(crops_buf);
    cudaFree(origin_buf);
    for (auto& sl : slots) {
      cudaFree(sl.img);
      cudaFree(sl.out);
      if (sl.img_free) cudaEventDestroy(sl.img_free);
      if (sl.out_free) cudaEventDestroy(sl.out_free);
    }
    for (auto e : ev_in) cudaEventDestroy(e);
    for (auto e : ev_done) cudaEventDestroy(e);
    for (auto e : ev_rows) cudaEventDestroy(e);
    if (s_in) { cudaStreamDestroy(s_in); cudaStreamDestroy(s_comp); cudaStreamDestroy(s_out); }
  }
};

namespace {

// ------------------------------------------------------------------ state_dict access
struct HostTensor {
  std::vector<float> v;
  std::vector<int64_t> shape;
};

int fetch(const nind_tensor* ts, int n, const std::string& name, const std::vector<int64_t>& want,
          HostTensor* out) {
  for (int i = 0; i < n; ++i) {
    if (name != ts[i].name) continue;
    if (ts[i].ndim < 0 || ts[i].ndim > 4 || !ts[i].data)
      return fail(NIND_E_WEIGHTS, "tensor " + name + " has an illegal descriptor (ndim must be 0..4, data non-null)");
    size_t count = 1;
    std::vector<int64_t> shp(ts[i].shape, ts[i].shape + ts[i].ndim);
    for (auto d : shp) count *= (size_t)d;
    if (!want.empty() && shp != want) {
      std::string m = "tensor " + name + " has shape [";
      for (auto d : shp) m += std::to_string(d) + ",";
      m += "] expected [";
      for (auto d : want) m += std::to_string(d) + ",";
      return fail(NIND_E_WEIGHTS, m + "]");
    }
    out->v.resize(count);
    out->shape = shp;
    CUDA_TRY(cudaMemcpy(out->v.data(), ts[i].data, count * sizeof(float), cudaMemcpyDefault));
    return 0;
  }
  return fail(NIND_E_WEIGHTS, "state_dict has no tensor named " + name);
}

int upload_layer(nind_net* net, const std::string& name, PackedLayer& L, const std::vector<__nv_bfloat16>& w,
                 const std::vector<float>& bias) {
  CUDA_TRY(cudaMalloc(&L.w, w.size() * sizeof(__nv_bfloat16)));
  CUDA_TRY(cudaMemcpy(L.w, w.data(), w.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMalloc(&L.bias, bias.size() * sizeof(float)));
  CUDA_TRY(cudaMemcpy(L.bias, bias.data(), bias.size() * sizeof(float), cudaMemcpyHostToDevice));
  if (L.taps == 9 && L.n_total == 64 && L.cin == 64) {  // pixel-pair mode candidate (igemm.cuh "PM")
    std::vector<__nv_bfloat16> wp;
    pack_pair_weights(w.data(), L.cin, &wp);
    CUDA_TRY(cudaMalloc(&L.w_pair, wp.size() * sizeof(__nv_bfloat16)));
    CUDA_TRY(cudaMemcpy(L.w_pair, wp.data(), wp.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
  }
  net->layers[name] = L;
  return 0;
}

// Conv2d weight [co][ci][3][3] (optionally scaled per co) -> [t][co][ci]
void pack_conv3(const float* w, int co, int ci, const float* scale, std::vector<__nv_bfloat16>* out) {
  out->assign((size_t)9 * co * ci, __float2bfloat16(0.f));
  for (int o = 0; o < co; ++o)
    for (int i = 0; i < ci; ++i)
      for (int t = 0; t < 9; ++t) {
        float v = w[((size_t)o * ci + i) * 9 + t];
        if (scale) v *= scale[o];
        (*out)[((size_t)t * co + o) * ci + i] = __float2bfloat16(v);
      }
}
// ConvTranspose2d(k=3,s=1) weight [ci][co][3][3] == conv over the 2-padded input with flipped taps.
void pack_tconv3(const float* w, int ci, int co, std::vector<__nv_bfloat16>* out) {
  out->assign((size_t)9 * co * ci, __float2bfloat16(0.f));
  for (int i = 0; i < ci; ++i)
    for (int o = 0; o < co; ++o)
      for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx) {
          const float v = w[(((size_t)i * co + o) * 3 + (2 - ky)) * 3 + (2 - kx)];
          (*out)[((size_t)(ky * 3 + kx) * co + o) * ci + i] = __float2bfloat16(v);
        }
}
// ConvTranspose2d(k=2,s=2) weight [ci][co][2][2] -> per-pixel GEMM [(dy*2+dx)*co + o][ci]
void pack_up2(const float* w, int ci, int co, std::vector<__nv_bfloat16>* out) {
  out->assign((size_t)4 * co * ci, __float2bfloat16(0.f));
  for (int i = 0; i < ci; ++i)
    for (int o = 0; o < co; ++o)
      for (int q = 0; q < 4; ++q)
        (*out)[((size_t)q * co + o) * ci + i] = __float2bfloat16(w[((size_t)i * co + o) * 4 + q]);
}
// First conv [co][3][3][3] for the 8-channel path: [5 MMAs][2 K-halves][co][8], K-half (j,h) = tap 2j+h except the
// last MMA, whose halves are (zero block, tap 8); the 8 channels are R,G,B hi | R,G,B lo | 0 0, so hi and lo
// parts meet the same weight.
void pack_first_c8(const float* w, int co, const float* scale, std::vector<__nv_bfloat16>* out) {
  out->assign((size_t)10 * co * 8, __float2bfloat16(0.f));
  const int tap_of[10] = {0, 1, 2, 3, 4, 5, 6, 7, -1, 8};
  for (int blk = 0; blk < 10; ++blk) {
    const int t = tap_of[blk];
    if (t < 0) continue;
    for (int o = 0; o < co; ++o)
      for (int c = 0; c < 3; ++c) {
        float v = w[((size_t)o * 3 + c) * 9 + t];
        if (scale) v *= scale[o];
        (*out)[((size_t)blk * co + o) * 8 + c] = __float2bfloat16(v);
        (*out)[((size_t)blk * co + o) * 8 + 3 + c] = __float2bfloat16(v);
      }
  }
}

int load_head(nind_net* net, const nind_tensor* ts, int n, const std::string& name) {
  HostTensor w, b;
  int rc;
  if ((rc = fetch(ts, n, name + ".weight", {3, 64, 1, 1}, &w))) return rc;
  if ((rc = fetch(ts, n, name + ".bias", {3}, &b))) return rc;
  CUDA_TRY(cudaMalloc(&net->head_w, 3 * 64 * sizeof(float)));
  CUDA_TRY(cudaMalloc(&net->head_b, 3 * sizeof(float)));
  CUDA_TRY(cudaMemcpy(net->head_w, w.v.data(), 3 * 64 * sizeof(float), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(net->head_b, b.v.data(), 3 * sizeof(float), cudaMemcpyHostToDevice));
  return 0;
}

int load_utnet(nind_net* net, const nind_tensor* ts, int n) {
  const int f = net->funit;
  if (f != 64) return fail(NIND_E_UNSUPPORTED, "UtNet: this build supports funit=64 only");
  int rc;
  auto act_of = [&](const std::string& act_name, PackedLayer* L) -> int {
    L->act = net->act_kind;
    L->slope = 0.f;
    if (net->act_kind == ACT_PRELU) {
      HostTensor s;
      int r = fetch(ts, n, act_name + ".weight", {1}, &s);
      if (r) return r;
      L->slope = s.v[0];
    }
    return 0;
  };
  auto conv = [&](const std::string& name, const std::string& act_name, int ci, int co, bool transposed) -> int {
    HostTensor w, b;
    std::vector<int64_t> shp = transposed ? std::vector<int64_t>{ci, co, 3, 3} : std::vector<int64_t>{co, ci, 3, 3};
    if ((rc = fetch(ts, n, name + ".weight", shp, &w))) return rc;
    if ((rc = fetch(ts, n, name + ".bias", {co}, &b))) return rc;
    PackedLayer L;
    L.cin = ci; L.cout = co; L.n_total = co; L.taps = 9;
    if ((rc = act_of(act_name, &L))) return rc;
    std::vector<__nv_bfloat16> pw;
    if (transposed) pack_tconv3(w.v.data(), ci, co, &pw);
    else pack_conv3(w.v.data(), co, ci, nullptr, &pw);
    return upload_layer(net, name, L, pw, b.v);
  };
  // first layer: 3 -> f as a 3x3 implicit GEMM over the 8-channel (hi/lo split) padded crop
  {
    HostTensor w, b;
    if ((rc = fetch(ts, n, "convs1.0.weight", {f, 3, 3, 3}, &w))) return rc;
    if ((rc = fetch(ts, n, "convs1.0.bias", {f}, &b))) return rc;
    PackedLayer L;
    L.cin = 8; L.cout = f; L.n_total = f; L.taps = 9;
    if ((rc = act_of("convs1.1", &L))) return rc;
    std::vector<__nv_bfloat16> pw;
    pack_first_c8(w.v.data(), f, nullptr, &pw);
    if ((rc = upload_layer(net, "convs1.0", L, pw, b.v))) return rc;
  }
  if ((rc = conv("convs1.2", "convs1.3", f, f, false))) return rc;
  int c = f;
  for (int lvl = 2; lvl <= 4; ++lvl) {
    const std::string p = "convs" + std::to_string(lvl);
    if ((rc = conv(p + ".0", p + ".1", c, 2 * c, false))) return rc;
    if ((rc = conv(p + ".2", p + ".3", 2 * c, 2 * c, false))) return rc;
    c *= 2;
  }
  if ((rc = conv("bottom.0", "bottom.1", 8 * f, 16 * f, false))) return rc;
  if ((rc = conv("bottom.2", "bottom.3", 16 * f, 16 * f, true))) return rc;
  int width = 16 * f;
  for (int lvl = 1; lvl <= 4; ++lvl) {
    const int half = width / 2;
    const std::string up = "up" + std::to_string(lvl), tc = "tconvs" + std::to_string(lvl);
    HostTensor w, b;
    if ((rc = fetch(ts, n, up + ".weight", {width, half, 2, 2}, &w))) return rc;
    if ((rc = fetch(ts, n, up + ".bias", {half}, &b))) return rc;
    PackedLayer L;
    L.cin = width; L.cout = half; L.n_total = 4 * half; L.taps = 1; L.act = ACT_NONE;
    std::vector<__nv_bfloat16> pw;
    pack_up2(w.v.data(), width, half, &pw);
    if ((rc = upload_layer(net, up, L, pw, b.v))) return rc;
    if ((rc = conv(tc + ".0", tc + ".1", width, half, true))) return rc;
    if ((rc = conv(tc + ".2", tc + ".3", half, half, true))) return rc;
    width = half;
  }
  return load_head(net, ts, n, "tconvs4.4");
}

int load_unet(nind_net* net, const nind_tensor* ts, int n) {
  int rc;
  // conv3x3(pad 1) + BatchNorm(eval) + ReLU, BN folded: w' = w*g/sqrt(var+eps), b' = (b-mean)*g/sqrt(var+eps)+beta
  auto conv_bn = [&](const std::string& conv_name, const std::string& bn_name, int ci, int co, bool first) -> int {
    HostTensor w, b, g, beta, mean, var;
    if ((rc = fetch(ts, n, conv_name + ".weight", {co, ci, 3, 3}, &w))) return rc;
    if ((rc = fetch(ts, n, conv_name + ".bias", {co}, &b))) return rc;
    if ((rc = fetch(ts, n, bn_name + ".weight", {co}, &g))) return rc;
    if ((rc = fetch(ts, n, bn_name + ".bias", {co}, &beta))) return rc;
    if ((rc = fetch(ts, n, bn_name + ".running_mean", {co}, &mean))) return rc;
    if ((rc = fetch(ts, n, bn_name + ".running_var", {co}, &var))) return rc;
    std::vector<float> scale(co), bias(co);
    for (int o = 0; o < co; ++o) {
      scale[o] = g.v[o] / std::sqrt(var.v[o] + 1e-5f);
      bias[o] = (b.v[o] - mean.v[o]) * scale[o] + beta.v[o];
    }
    PackedLayer L;
    L.cout = co; L.n_total = co; L.act = ACT_PRELU; L.slope = 0.f;  // ReLU
    std::vector<__nv_bfloat16> pw;
    if (first) {
      L.cin = 8; L.taps = 9;
      pack_first_c8(w.v.data(), co, scale.data(), &pw);
    } else {
      L.cin = ci; L.taps = 9;
      pack_conv3(w.v.data(), co, ci, scale.data(), &pw);
    }
    return upload_layer(net, conv_name, L, pw, bias);
  };
  auto dconv = [&](const std::string& p, int ci, int co, bool first) -> int {
    if ((rc = conv_bn(p + ".0", p + ".1", ci, co, first))) return rc;
    return conv_bn(p + ".3", p + ".4", co, co, false);
  };
  if ((rc = dconv("inc.conv.conv", 3, 64, true))) return rc;
  const int dn[4][2] = {{64, 128}, {128, 256}, {256, 512}, {512, 512}};
  for (int i = 0; i < 4; ++i)
    if ((rc = dconv("down" + std::to_string(i + 1) + ".mpconv.1.conv", dn[i][0], dn[i][1], false))) return rc;
  const int upc[4][2] = {{1024, 256}, {512, 128}, {256, 64}, {128, 64}};
  for (int i = 0; i < 4; ++i) {
    const std::string p = "up" + std::to_string(i + 1);
    const int half = upc[i][0] / 2;
    HostTensor w, b;
    if ((rc = fetch(ts, n, p + ".up.weight", {half, half, 2, 2}, &w))) return rc;
    if ((rc = fetch(ts, n, p + ".up.bias", {half}, &b))) return rc;
    PackedLayer L;
    L.cin = half; L.cout = half; L.n_total = 4 * half; L.taps = 1; L.act = ACT_NONE;
    std::vector<__nv_bfloat16> pw;
    pack_up2(w.v.data(), half, half, &pw);
    if ((rc = upload_layer(net, p + ".up", L, pw, b.v))) return rc;
    if ((rc = dconv(p + ".conv.conv", upc[i][0], upc[i][1], false))) return rc;
  }
  return load_head(net, ts, n, "outc.conv");
}

int load_weights(nind_net* net, const nind_tensor* ts, int n) {
  net->plans.clear();  // plans hold pointers into the old weights
  net->free_layers();
  return net->arch == NIND_ARCH_UTNET ? load_utnet(net, ts, n) : load_unet(net, ts, n);
}

// ------------------------------------------------------------------ plans
struct PlanBuilder {
  nind_net* net;
  Plan* plan;
  int rc = 0;
  bool last_pool_fused = false;

  ActBuf alloc(int b, int hs, int ws, int c) {
    ActBuf a;
    a.b = b; a.hs = hs; a.ws = ws; a.c = c;
    if (rc) return a;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, a.elems() * sizeof(__nv_bfloat16));
    if (e == cudaSuccess) e = cudaMemset(p, 0, a.elems() * sizeof(__nv_bfloat16));
    if (e != cudaSuccess) {
      rc = fail(NIND_E_CUDA, std::string("activation arena: ") + cudaGetErrorString(e));
      if (p) cudaFree(p);
      return a;
    }
    plan->bufs.push_back(p);
    plan->bytes += a.elems() * sizeof(__nv_bfloat16);
    a.ptr = static_cast<__nv_bfloat16*>(p);
    return a;
  }

  // 3x3 / 1x1 layer writing bf16 (EPI_STORE or EPI_D2S)
  void conv(const std::string& name, const ActBuf& in, int in_coff, const ActBuf& out, int out_coff,
            int out_halo, int epi, const ActBuf* pool = nullptr, int pool_halo = 0,
            bool c8 = false) {
    if (rc) return;
    auto it = net->layers.find(name);
    if (it == net->layers.end()) { rc = fail(NIND_E_WEIGHTS, "layer not loaded: " + name); return; }
    const PackedLayer& L = it->second;
    ConvSpec s;
    s.in = in; s.in_coff = in_coff; s.cin = L.cin; s.taps = L.taps; s.w = L.w; s.n_total = L.n_total;
    s.bias = L.bias; s.act = L.act; s.slope = L.slope; s.epi_mode = epi;
    s.out = out; s.out_coff = out_coff; s.out_halo = out_halo; s.d2s_cout = L.cout;
    // the fused pool pairs tile rows/columns, which needs even map sizes (UNet at cs 440 has a 55-wide level)
    last_pool_fused = false;
    if (pool && net->fuse_pool && !((in.hs | in.ws) & 1)) {
      s.pool = *pool; s.pool_halo = pool_halo;
      last_pool_fused = true;
    }
    s.c8 = c8;
    s.max_ctas = net->max_ctas;
    s.cg = net->cg;
    s.dual = net->dual;
    s.wide = net->wide;
    s.flat = net->flat == 1 ? 2 : net->flat;  // 1 = on every layer where it is legal
    if (L.n_total >= 256) s.n_tile = net->n_tile_deep;
    maybe_pair(s, L, in, in_coff);
    if (rc) return;
    Step st;
    st.kind = STEP_IGEMM; st.name = name;
    std::string why;
    if (!build_igemm(s, &st.ig, &why)) { rc = fail(NIND_E_INVALID, name + ": " + why); return; }
    st.flops = st.ig.flops;
    {
      const IgemmParams& q = st.ig.p;
      const double in_b = (double)in.b * in.hs * in.ws * L.cin * 2;
      const double out_px = (double)in.b * q.h_valid * q.w_valid * (epi == EPI_D2S ? 4 : 1);
      st.bytes = in_b + out_px * (epi == EPI_D2S ? L.cout : L.n_total) * 2 + (s.pool.ptr ? out_px / 4 * L.n_total * 2 : 0);
    }
    plan->steps.push_back(st);
  }

  // Pixel-pair mode (option "pair64"): eligible C_out = 64 3x3 layers run as N = 128 GEMMs over pixel pairs.
  void maybe_pair(ConvSpec& s, const PackedLayer& L, const ActBuf& in, int in_coff) {
    if (rc || !net->pair64 || !L.w_pair || in_coff != 0 || L.cin != in.c || (in.ws & 1) || s.epi_mode == EPI_D2S ||
        s.c8)
      return;
    s.w = L.w_pair;
    s.pair = true;
    s.flat = 0;
  }

  void head(const std::string& name, const ActBuf& in, int unpad, int hy, int hx, int sigmoid) {
    if (rc) return;
    auto it = net->layers.find(name);
    if (it == net->layers.end()) { rc = fail(NIND_E_WEIGHTS, "layer not loaded: " + name); return; }
    const PackedLayer& L = it->second;
    ConvSpec s;
    s.in = in; s.cin = L.cin; s.taps = L.taps; s.w = L.w; s.n_total = L.n_total;
    s.bias = L.bias; s.act = L.act; s.slope = L.slope; s.epi_mode = EPI_HEAD;
    s.head_w = net->head_w; s.head_b = net->head_b; s.head_out = nullptr;
    s.head_unpad = unpad; s.head_hy = hy; s.head_hx = hx; s.head_sigmoid = sigmoid;
    s.max_ctas = net->max_ctas;
    s.cg = net->cg;
    s.dual = net->dual;
    s.wide = net->wide;
    maybe_pair(s, L, in, 0);
    if (rc) return;
    Step st;
    st.kind = STEP_IGEMM; st.name = name + "+head";
    std::string why;
    if (!build_igemm(s, &st.ig, &why)) { rc = fail(NIND_E_INVALID, name + ": " + why); return; }
    st.flops = st.ig.flops + 2.0 * in.b * hy * hx * 64 * 3;
    plan->head_step = (int)plan->steps.size();
    plan->steps.push_back(st);
  }

  // 2x2 max-pool of channels [coff, coff+c) of the interior of `in` (halo hi) into the interior of out.
  void pool(const std::string& name, const ActBuf& in, int hi, int coff, int c, const ActBuf& out, int ho) {
    if (rc) return;
    Step st;
    st.kind = STEP_POOL; st.name = name;
    PoolParams& p = st.pl;
    p.i_pix = in.c; p.i_row = (long long)in.ws * in.c; p.i_img = (long long)in.hs * p.i_row;
    p.in = in.ptr + hi * p.i_row + (long long)hi * p.i_pix + coff;
    p.o_pix = out.c; p.o_row = (long long)out.ws * out.c; p.o_img = (long long)out.hs * p.o_row;
    p.out = out.ptr + ho * p.o_row + (long long)ho * p.o_pix;
    p.n = in.b; p.ho = out.hs - 2 * ho; p.wo = out.ws - 2 * ho; p.c = c;
    st.bytes = (double)in.b * p.ho * p.wo * c * 2 * 5;
    plan->steps.push_back(st);
  }

  static GatherParams gather_geom(const ActBuf& x0, int crop_h, int crop_w, int pad, int reflect) {
    GatherParams g;
    memset(&g, 0, sizeof g);
    g.crop_h = crop_h; g.crop_w = crop_w; g.pad = pad; g.reflect = reflect;
    g.out_h = x0.hs; g.out_w = x0.ws; g.n_crops = x0.b; g.dst = x0.ptr;
    return g;
  }

  void gather(const ActBuf& x0, int crop_h, int crop_w, int pad, int reflect) {
    if (rc) return;
    Step st;
    st.kind = STEP_GATHER; st.name = "gather+pad8";
    GatherParams& g = st.g;
    g = gather_geom(x0, crop_h, crop_w, pad, reflect);
    st.bytes = (double)x0.b * (3.0 * crop_h * crop_w * 4 + (double)x0.hs * x0.ws * x0.c * 2);
    plan->gather_step = (int)plan->steps.size();
    plan->steps.push_back(st);
  }
};

bool utnet_legal(int s) { return s >= 104 && (s - 56) % 16 == 0; }

int build_utnet_plan(nind_net* net, Plan* plan, int B, int H, int W) {
  if (!utnet_legal(H) || !utnet_legal(W))
    return fail(NIND_E_INVALID, "UtNet: crop height/width must be 16a+56 with a>=3 (e.g. 120, 248, 504); got " +
                                    std::to_string(H) + "x" + std::to_string(W));
  PlanBuilder pb{net, plan};
  const int f = net->funit;
  // spatial sizes per level (h, w): e = encoder output, p = pooled
  int eh[5], ew[5], ph[5], pw[5];
  eh[1] = H; ew[1] = W;
  for (int l = 1; l <= 4; ++l) {
    ph[l] = eh[l] / 2; pw[l] = ew[l] / 2;
    if (l < 4) { eh[l + 1] = ph[l] - 4; ew[l + 1] = pw[l] - 4; }
  }
  // encoder
  // first layer input: the padded crops as 8 bf16 channels/pixel (default), or the 64-wide im2col
  ActBuf x0 = pb.alloc(B, H + 4, W + 4, 8);
  pb.gather(x0, H, W, 2, 1);
  ActBuf cat[5];
  ActBuf cur;  // input of the level's first conv
  for (int l = 1; l <= 4; ++l) {
    const int c = f << (l - 1);
    const std::string p = "convs" + std::to_string(l);
    ActBuf a = pb.alloc(B, eh[l] + 2, ew[l] + 2, c);
    if (l == 1) pb.conv(p + ".0", x0, 0, a, 0, 0, EPI_STORE, nullptr, 0, true);
    else pb.conv(p + ".0", cur, 0, a, 0, 0, EPI_STORE);
    cat[l] = pb.alloc(B, eh[l] + 4, ew[l] + 4, 2 * c);  // [up | skip], 2-px zero frame for the ConvT
    cur = pb.alloc(B, ph[l], pw[l], c);
    pb.conv(p + ".2", a, 0, cat[l], c, 2, EPI_STORE, &cur, 0);  // + fused MaxPool2d(2) into `cur`
    if (!pb.last_pool_fused) pb.pool("maxpool" + std::to_string(l), cat[l], 2, c, c, cur, 0);
  }
  // bottom
  ActBuf bt0 = pb.alloc(B, ph[4] - 2 + 4, pw[4] - 2 + 4, 16 * f);
  pb.conv("bottom.0", cur, 0, bt0, 0, 2, EPI_STORE);
  ActBuf dec = pb.alloc(B, ph[4], pw[4], 16 * f);
  pb.conv("bottom.2", bt0, 0, dec, 0, 0, EPI_STORE);
  // decoder
  for (int lvl = 1; lvl <= 4; ++lvl) {
    const int l = 5 - lvl;            // encoder level whose skip is concatenated
    const int c = f << (l - 1);       // channels after this decoder level
    const std::string tc = "tconvs" + std::to_string(lvl);
    pb.conv("up" + std::to_string(lvl), dec, 0, cat[l], 0, 2, EPI_D2S);
    ActBuf t = pb.alloc(B, eh[l] + 2 + 4, ew[l] + 2 + 4, c);
    pb.conv(tc + ".0", cat[l], 0, t, 0, 2, EPI_STORE);
    if (lvl < 4) {
      dec = pb.alloc(B, eh[l] + 4, ew[l] + 4, c);
      pb.conv(tc + ".2", t, 0, dec, 0, 0, EPI_STORE);
    } else {
      pb.head(tc + ".2", t, 2, H, W, 0);
    }
  }
  return pb.rc;
}

int build_unet_plan(nind_net* net, Plan* plan, int B, int H, int W) {
  // Any size works as in the reference: MaxPool2d floors, and the F.pad of the upsampled tensor to the
  // skip's size (ThirdPartyNets.py:114-118: zeros on the right/bottom) is the never-written, zero-filled
  // remainder of the concat buffer's up half.
  if (H < 16 || W < 16)
    return fail(NIND_E_INVALID, "UNet: crop height/width must be at least 16; got " + std::to_string(H) + "x" +
                                    std::to_string(W));
  PlanBuilder pb{net, plan};
  // level l (0..4) spatial size
  int sh[5], sw[5];
  for (int l = 0; l < 5; ++l) { sh[l] = H >> l; sw[l] = W >> l; }
  const int ch[5] = {64, 128, 256, 512, 512};
  ActBuf x0 = pb.alloc(B, H + 2, W + 2, 8);
  pb.gather(x0, H, W, 1, 0);
  // every 3x3 conv is padding=1: its input buffer carries a 1-px zero frame
  ActBuf cat[4];  // [skip | up] for decoder levels, skip written by the encoder
  ActBuf cur;
  const char* enc_name[5] = {"inc.conv.conv", "down1.mpconv.1.conv", "down2.mpconv.1.conv",
                             "down3.mpconv.1.conv", "down4.mpconv.1.conv"};
  for (int l = 0; l < 5; ++l) {
    const std::string p = enc_name[l];
    ActBuf mid = pb.alloc(B, sh[l] + 2, sw[l] + 2, ch[l]);
    if (l == 0) pb.conv(p + ".0", x0, 0, mid, 0, 1, EPI_STORE, nullptr, 0, true);
    else pb.conv(p + ".0", cur, 0, mid, 0, 1, EPI_STORE);
    if (l < 4) {
      cat[l] = pb.alloc(B, sh[l] + 2, sw[l] + 2, 2 * ch[l]);
      cur = pb.alloc(B, sh[l + 1] + 2, sw[l + 1] + 2, ch[l]);
      pb.conv(p + ".3", mid, 0, cat[l], 0, 1, EPI_STORE, &cur, 1);  // + fused MaxPool2d(2)
      if (!pb.last_pool_fused) pb.pool("maxpool" + std::to_string(l + 1), cat[l], 1, 0, ch[l], cur, 1);
    } else {
      cur = pb.alloc(B, sh[l], sw[l], ch[l]);
      pb.conv(p + ".3", mid, 0, cur, 0, 0, EPI_STORE);
    }
  }
  const int outc[4] = {256, 128, 64, 64};
  for (int i = 0; i < 4; ++i) {
    const int l = 3 - i;  // skip level
    const std::string p = "up" + std::to_string(i + 1);
    pb.conv(p + ".up", cur, 0, cat[l], ch[l], 1, EPI_D2S);
    ActBuf mid = pb.alloc(B, sh[l] + 2, sw[l] + 2, outc[i]);
    pb.conv(p + ".conv.conv.0", cat[l], 0, mid, 0, 1, EPI_STORE);
    if (i < 3) {
      cur = pb.alloc(B, sh[l], sw[l], outc[i]);
      pb.conv(p + ".conv.conv.3", mid, 0, cur, 0, 0, EPI_STORE);
    } else {
      pb.head(p + ".conv.conv.3", mid, 0, H, W, 1);
    }
  }
  return pb.rc;
}

int get_plan(nind_net* net, int B, int H, int W, Plan** out) {
  std::vector<int> key{B, H, W};
  auto it = net->plans.find(key);
  if (it != net->plans.end()) {
    it->second->last_use = ++net->plan_tick;
    *out = it->second.get();
    return 0;
  }
  std::unique_ptr<Plan> plan(new Plan);
  plan->b = B; plan->h = H; plan->w = W;
  int rc = net->arch == NIND_ARCH_UTNET ? build_utnet_plan(net, plan.get(), B, H, W)
                                         : build_unet_plan(net, plan.get(), B, H, W);
  if (rc) return rc;
  // Keep a few plans alive (each owns an activation arena): least recently used go first once there are
  // more than 8 or they hold more than 48 GB.  (cudaFree waits for the device, so work still queued on an
  // evicted plan's buffers has finished.)
  plan->last_use = ++net->plan_tick;
  for (;;) {
    size_t total = plan->bytes;
    auto lru = net->plans.end();
    for (auto i = net->plans.begin(); i != net->plans.end(); ++i) {
      total += i->second->bytes;
      if (lru == net->plans.end() || i->second->last_use < lru->second->last_use) lru = i;
    }
    if (lru == net->plans.end() || (net->plans.size() < 8 && total <= (48ull << 30))) break;
    net->plans.erase(lru);
  }
  *out = plan.get();
  net->plans[key] = std::move(plan);
  return 0;
}

int grid_for(long long work_items) {
  const long long blocks = (work_items + 255) / 256;
  const long long cap = (long long)device_sm_count() * 16;
  return (int)std::max(1LL, std::min(blocks, cap));
}

int run_plan(nind_net* net, Plan* plan, const GatherParams& gsrc, float* head_out, cudaStream_t st,
             int head_clamp = 0) {
  std::vector<cudaEvent_t> ev;
  if (net->timing) {
    ev.resize(plan->steps.size() + 1);
    for (auto& e : ev) CUDA_TRY(cudaEventCreate(&e));
    CUDA_TRY(cudaEventRecord(ev[0], st));
  }
  for (size_t i = 0; i < plan->steps.size(); ++i) {
    Step& s = plan->steps[i];
    if (s.kind == STEP_GATHER) {
      GatherParams g = s.g;
      g.src = gsrc.src; g.src_img = gsrc.src_img; g.src_plane = gsrc.src_plane;
      g.src_w = gsrc.src_w; g.src_h = gsrc.src_h; g.origin = gsrc.origin;
      gather_pad8_kernel<<<grid_for((long long)g.n_crops * g.out_h * g.out_w), 256, 0, st>>>(g);
    } else if (s.kind == STEP_POOL) {
      maxpool2_kernel<<<grid_for((long long)s.pl.n * s.pl.ho * s.pl.wo * (s.pl.c / 8)), 256, 0, st>>>(s.pl);
    } else {
      IgemmLaunch L = s.ig;
      if ((int)i == plan->head_step) { L.p.head_out = head_out; L.p.head_clamp = head_clamp; }
      L.pdl = net->pdl && !net->timing;  // per-layer timing wants the launches serialised

      cudaError_t e = launch_igemm(L, net->err_flag, st);
      if (e != cudaSuccess) return fail(NIND_E_CUDA, s.name + ": " + cudaGetErrorString(e));
    }
    ++g_launches;
    if (net->timing) CUDA_TRY(cudaEventRecord(ev[i + 1], st));
  }
  CUDA_TRY(cudaGetLastError());
  if (net->timing) {
    CUDA_TRY(cudaStreamSynchronize(st));
    net->t_names.clear(); net->t_ms.clear(); net->t_flops.clear(); net->t_bytes.clear();
    for (size_t i = 0; i < plan->steps.size(); ++i) {
      float ms = 0;
      cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
      net->t_names.push_back(plan->steps[i].name);
      net->t_ms.push_back(ms);
      net->t_flops.push_back(plan->steps[i].flops);
      net->t_bytes.push_back(plan->steps[i].bytes);
    }
    for (auto& e : ev) cudaEventDestroy(e);
  }
  return 0;
}

// The role code of a pipeline time-out lives in mapped host memory: readable without a CUDA call, also after
// the trap has poisoned the context.
int check_err_flag(nind_net* net) {
  const int h = *reinterpret_cast<volatile int*>(net->err_host);
  if (h) return fail(NIND_E_KERNEL, "kernel pipeline time-out, role code " + std::to_string(h) +
                                        " (1 A-producer, 2 B-producer, 3 MMA/TMEM, 4 MMA/A, 5 MMA/B, 6 epilogue)");
  return 0;
}

// Every entry point runs on the handle's device, whatever the caller's current device is.
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    if (prev == dev) prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};
#define ENTER(net)                                                              \
  DeviceGuard guard_((net)->device);                                            \
  if (!guard_.ok) return fail(NIND_E_CUDA, "cannot switch to the handle's device"); \
  { int rc_ = check_err_flag(net); if (rc_) return rc_; }

// The plan arenas, crops_buf and origin_buf are shared by every entry point of a handle: a call's compute
// stream first waits for the last use recorded by the previous call (which may have been on another stream).
int scratch_acquire(nind_net* net, cudaStream_t st) {
  CUDA_TRY(cudaStreamWaitEvent(st, net->ev_last, 0));
  return 0;
}
int scratch_release(nind_net* net, cudaStream_t st) {
  CUDA_TRY(cudaEventRecord(net->ev_last, st));
  return 0;
}

int ensure(void** p, size_t* cap, size_t bytes) {
  if (*cap >= bytes) return 0;
  if (*p) cudaFree(*p);
  *p = nullptr; *cap = 0;
  CUDA_TRY(cudaMalloc(p, bytes));
  *cap = bytes;
  return 0;
}

}  // namespace

// ====================================================================== C ABI
extern "C" {

const char* nind_last_error(void) { return g_err.c_str(); }

int64_t nind_kernel_launches(void) { return g_launches.load(); }

int nind_device_info(int* sm_major, int* sm_minor, int* sm_count) {
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
  if (sm_major) *sm_major = prop.major;
  if (sm_minor) *sm_minor = prop.minor;
  if (sm_count) *sm_count = prop.multiProcessorCount;
  return 0;
}

int nind_net_create(int arch, int funit, int activation, const nind_tensor* tensors, int n_tensors,
                    nind_net** out) {
  if (!out || !tensors) return fail(NIND_E_INVALID, "null argument");
  if (arch != NIND_ARCH_UTNET && arch != NIND_ARCH_UNET) return fail(NIND_E_INVALID, "unknown architecture");
  if (activation < NIND_ACT_PRELU || activation > NIND_ACT_HARDSWISH)
    return fail(NIND_E_INVALID, "unknown activation function");
  int major = 0, minor = 0;
  int rc = nind_device_info(&major, &minor, nullptr);
  if (rc) return rc;
  if (major != 10) return fail(NIND_E_UNSUPPORTED, "this library needs an sm_100 (B200) device; found sm_" +
                                                       std::to_string(major) + std::to_string(minor));
  std::unique_ptr<nind_net> net(new nind_net);
  net->arch = arch;
  net->funit = funit;
  net->act_kind = activation == NIND_ACT_PRELU ? ACT_PRELU : (activation == NIND_ACT_ELU ? ACT_ELU : ACT_HARDSWISH);
  CUDA_TRY(cudaGetDevice(&net->device));
  CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&net->err_host), sizeof(int), cudaHostAllocMapped));
  *net->err_host = 0;
  CUDA_TRY(cudaHostGetDevicePointer(reinterpret_cast<void**>(&net->err_flag), net->err_host, 0));
  CUDA_TRY(cudaEventCreateWithFlags(&net->ev_last, cudaEventDisableTiming));
  if ((rc = load_weights(net.get(), tensors, n_tensors))) return rc;
  *out = net.release();
  return 0;
}

int nind_net_load(nind_net* net, const nind_tensor* tensors, int n_tensors) {
  if (!net || !tensors) return fail(NIND_E_INVALID, "null argument");
  ENTER(net);
  CUDA_TRY(cudaDeviceSynchronize());
  return load_weights(net, tensors, n_tensors);
}

void nind_net_destroy(nind_net* net) {
  if (!net) return;
  DeviceGuard guard(net->device);
  cudaDeviceSynchronize();
  delete net;
}

int nind_net_device(nind_net* net, int* device) {
  if (!net || !device) return fail(NIND_E_INVALID, "null argument");
  *device = net->device;
  return 0;
}

int nind_set_timing(nind_net* net, int enabled) {
  if (!net) return fail(NIND_E_INVALID, "null handle");
  net->timing = enabled;
  return 0;
}

int nind_get_layer_times(nind_net* net, int max_layers, const char** names, float* ms, double* flops,
                         int* n_layers) {
  if (!net || !n_layers) return fail(NIND_E_INVALID, "null argument");
  const int n = (int)net->t_ms.size();
  *n_layers = n;
  for (int i = 0; i < n && i < max_layers; ++i) {
    if (names) names[i] = net->t_names[i].c_str();
    if (ms) ms[i] = net->t_ms[i];
    if (flops) flops[i] = net->t_flops[i];
  }
  return 0;
}

int nind_get_layer_bytes(nind_net* net, int max_layers, double* bytes, int* n_layers) {
  if (!net || !n_layers) return fail(NIND_E_INVALID, "null argument");
  const int n = (int)net->t_bytes.size();
  *n_layers = n;
  for (int i = 0; i < n && i < max_layers; ++i)
    if (bytes) bytes[i] = net->t_bytes[i];
  return 0;
}

int nind_set_option(nind_net* net, const char* key, int value) {
  if (!net || !key) return fail(NIND_E_INVALID, "null argument");
  ENTER(net);
  const std::string k = key;
  if (k == "host_first") {  // crops in the first step of the host pipeline (-1: up to the first grid-row boundary)
    net->host_first = value;
    return 0;
  }
  if (k == "host_last") {
    net->host_last = value;
    return 0;
  }
  if (k == "n_tile_deep") {
    if (value != 128 && value != 256) return fail(NIND_E_INVALID, "n_tile_deep must be 128 or 256");
    net->n_tile_deep = value;
  } else if (k == "max_ctas") {
    net->max_ctas = value;
  } else if (k == "cta_group") {
    if (value < 0 || value > 2) return fail(NIND_E_INVALID, "cta_group must be 0 (auto), 1 or 2");
    net->cg = value;
  } else if (k == "fuse_pool") {
    net->fuse_pool = value ? 1 : 0;
  } else if (k == "pair64") {  // pixel-pair mode for the C_out = 64 3x3 layers (0 | 1)
    net->pair64 = value ? 1 : 0;
  } else if (k == "pdl") {  // programmatic dependent launch of the conv kernels (0 | 1)
    net->pdl = value ? 1 : 0;
    return 0;
  } else if (k == "dual_issuer") {  // two MMA issuer warps on alternate tiles (0 | 1)
    net->dual = value ? 1 : 0;
  } else if (k == "wide_store") {  // 128-byte epilogue staging rows: -1 auto, 0 off, 1 wherever legal
    net->wide = value;
  } else if (k == "flat") {  // flat (1-D) tiles on narrow maps: -1 auto, 0 off, 1 wherever legal
    net->flat = value;
  } else {
    return fail(NIND_E_INVALID, "unknown option " + k);
  }
  CUDA_TRY(cudaDeviceSynchronize());  // plans still in flight own the arenas that are about to be freed
  net->plans.clear();
  return 0;
}

int nind_net_forward_ex(nind_net* net, const float* in_nchw, float* out_nchw, int batch, int h, int w, int flags,
                        void* stream) {
  if (!net || !in_nchw || !out_nchw) return fail(NIND_E_INVALID, "null argument");
  if (batch <= 0) return fail(NIND_E_INVALID, "batch must be positive");
  if (flags & ~NIND_FWD_CLAMP01) return fail(NIND_E_INVALID, "unknown forward flag");
  ENTER(net);
  Plan* plan = nullptr;
  int rc = get_plan(net, batch, h, w, &plan);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GatherParams g;
  memset(&g, 0, sizeof g);
  g.src = in_nchw; g.src_img = 3LL * h * w; g.src_plane = (long long)h * w; g.src_w = w; g.src_h = h;
  g.origin = nullptr;
  if ((rc = scratch_acquire(net, st))) return rc;
  rc = run_plan(net, plan, g, out_nchw, st, (flags & NIND_FWD_CLAMP01) ? 1 : 0);
  scratch_release(net, st);
  return rc;
}

int nind_net_forward(nind_net* net, const float* in_nchw, float* out_nchw, int batch, int h, int w,
                     void* stream) {
  return nind_net_forward_ex(net, in_nchw, out_nchw, batch, h, w, 0, stream);
}

int nind_crop_table(int width, int height, int cs, int ucs, int ol, nind_crop* table, int* n_crops) {
  GridGeom g;
  if (!make_grid(width, height, cs, ucs, ol, &g)) return fail(NIND_E_INVALID, "illegal crop geometry");
  if (n_crops) *n_crops = g.size();
  if (table)
    for (int i = 0; i < g.size(); ++i) crop_entry(g, i, &table[i]);
  return 0;
}

int nind_band_rows(int width, int height, int cs, int ucs, int ol, int crop_begin, int crop_end,
                   int* band_y0, int* band_y1) {
  GridGeom g;
  if (!make_grid(width, height, cs, ucs, ol, &g)) return fail(NIND_E_INVALID, "illegal crop geometry");
  if (crop_begin < 0 || crop_end > g.size() || crop_begin >= crop_end)
    return fail(NIND_E_INVALID, "illegal crop range");
  int y0, y1;
  band_of(g, crop_begin, crop_end, &y0, &y1);
  if (band_y0) *band_y0 = y0;
  if (band_y1) *band_y1 = y1;
  return 0;
}

static int upload_origins(nind_net* net, const GridGeom& g, int crop_begin, int n, cudaStream_t st) {
  int rc;
  const std::vector<int> key{g.W, g.H, g.cs, g.ucs, g.ol, crop_begin, n};
  if (net->origin_buf && key == net->origin_key) return 0;  // same table as last call: no copy, no sync
  net->origin_key.clear();
  if ((rc = ensure(reinterpret_cast<void**>(&net->origin_buf), &net->origin_cap, (size_t)n * sizeof(int2))))
    return rc;
  std::vector<int2> origins(n);
  for (int i = 0; i < n; ++i) {
    nind_crop c;
    crop_entry(g, crop_begin + i, &c);
    origins[i] = make_int2(c.x0, c.y0);
  }
  CUDA_TRY(cudaMemcpyAsync(net->origin_buf, origins.data(), (size_t)n * sizeof(int2), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaStreamSynchronize(st));  // `origins` is a stack-lifetime pageable buffer
  net->origin_key = key;
  return 0;
}

// Stitch rows [y0, y1) from `crops` (outputs of crops [crop_begin, crop_end)) into `out`.
// whole_image: `out` is the full [3][H][W] image; else a dense band [3][y1-y0][W].
static int launch_stitch(const GridGeom& g, const float* crops, int crop_begin, int crop_end, float* out,
                         int y0, int y1, bool whole_image, cudaStream_t st) {
  StitchParams sp;
  sp.crops = crops; sp.crop_begin = crop_begin; sp.crop_end = crop_end; sp.out = out;
  sp.y_begin = y0; sp.y_end = y1;
  sp.out_plane = whole_image ? (long long)g.H * g.W : (long long)(y1 - y0) * g.W;
  sp.out_y0 = whole_image ? 0 : y0;
  sp.W = g.W; sp.H = g.H; sp.cs = g.cs; sp.ucs = g.ucs; sp.ol = g.ol; sp.pad = g.pad; sp.stride = g.stride;
  sp.nx = g.nx; sp.ny = g.ny;
  stitch_kernel<<<grid_for((long long)(y1 - y0) * g.W), 256, 0, st>>>(sp);
  ++g_launches;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

static int check_range(int width, int height, int cs, int ucs, int ol, int crop_begin, int crop_end, GridGeom* g) {
  if (!make_grid(width, height, cs, ucs, ol, g)) return fail(NIND_E_INVALID, "illegal crop geometry");
  if (crop_begin < 0 || crop_end > g->size() || crop_begin >= crop_end)
    return fail(NIND_E_INVALID, "illegal crop range");
  return 0;
}

int nind_gather_crops(nind_net* net, const float* img_chw, int height, int width, int cs, int ucs, int ol,
                      int crop_begin, int crop_end, float* crops_out, void* stream) {
  if (!net || !img_chw || !crops_out) return fail(NIND_E_INVALID, "null argument");
  ENTER(net);
  GridGeom g;
  int rc;
  if ((rc = check_range(width, height, cs, ucs, ol, crop_begin, crop_end, &g))) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int n = crop_end - crop_begin;
  if ((rc = scratch_acquire(net, st))) return rc;
  if ((rc = upload_origins(net, g, crop_begin, n, st))) return rc;
  CropGatherParams p;
  p.src = img_chw; p.src_plane = (long long)height * width; p.src_w = width; p.src_h = height;
  p.origin = net->origin_buf; p.cs = cs; p.n_crops = n; p.dst = crops_out;
  gather_crops_kernel<<<grid_for((long long)n * 3 * cs * cs), 256, 0, st>>>(p);
  ++g_launches;
  CUDA_TRY(cudaGetLastError());
  return scratch_release(net, st);
}

int nind_stitch_crops(const float* crops, int height, int width, int cs, int ucs, int ol, int crop_begin,
                      int crop_end, float* out_band, int* band_y0, int* band_y1, void* stream) {
  if (!crops || !out_band) return fail(NIND_E_INVALID, "null argument");
  GridGeom g;
  int rc;
  if ((rc = check_range(width, height, cs, ucs, ol, crop_begin, crop_end, &g))) return rc;
  int y0, y1;
  band_of(g, crop_begin, crop_end, &y0, &y1);
  if (band_y0) *band_y0 = y0;
  if (band_y1) *band_y1 = y1;
  return launch_stitch(g, crops, crop_begin, crop_end, out_band, y0, y1, false, static_cast<cudaStream_t>(stream));
}

// Crops [a, b) in forwards of at most `batch` crops, balanced: ceil(n / batch) forwards of near-equal size
// (532 crops at batch 168 run as 4 x 133, not 3 x 168 + 28: the small tail plan is the least efficient one).
static void balanced_steps(int a, int b, int batch, std::vector<std::pair<int, int>>* steps) {
  const int n = b - a;
  if (n <= 0) return;
  const int k = (n + batch - 1) / batch;
  for (int i = 0; i < k; ++i) steps->push_back({a + (int)((long long)n * i / k), a + (int)((long long)n * (i + 1) / k)});
}

int nind_tiled_denoise(nind_net* net, const float* img_chw, float* out_band, int height, int width,
                       int cs, int ucs, int ol, int crop_begin, int crop_end, int batch,
                       int* band_y0, int* band_y1, void* stream) {
  if (!net || !img_chw || !out_band) return fail(NIND_E_INVALID, "null argument");
  ENTER(net);
  GridGeom g;
  int rc;
  if ((rc = check_range(width, height, cs, ucs, ol, crop_begin, crop_end, &g))) return rc;
  if (batch <= 0) return fail(NIND_E_INVALID, "batch must be positive");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int n = crop_end - crop_begin;
  if ((rc = scratch_acquire(net, st))) return rc;
  if ((rc = ensure(reinterpret_cast<void**>(&net->crops_buf), &net->crops_cap,
                   (size_t)n * 3 * cs * cs * sizeof(float))))
    return rc;
  if ((rc = upload_origins(net, g, crop_begin, n, st))) return rc;
  std::vector<std::pair<int, int>> steps;
  balanced_steps(0, n, batch, &steps);
  for (auto& sp : steps) {
    Plan* plan = nullptr;
    if ((rc = get_plan(net, sp.second - sp.first, cs, cs, &plan))) return rc;
    GatherParams gp;
    memset(&gp, 0, sizeof gp);
    gp.src = img_chw; gp.src_img = 0; gp.src_plane = (long long)height * width; gp.src_w = width; gp.src_h = height;
    gp.origin = net->origin_buf + sp.first;
    if ((rc = run_plan(net, plan, gp, net->crops_buf + (size_t)sp.first * 3 * cs * cs, st))) return rc;
  }
  int y0, y1;
  band_of(g, crop_begin, crop_end, &y0, &y1);
  if (band_y0) *band_y0 = y0;
  if (band_y1) *band_y1 = y1;
  if ((rc = launch_stitch(g, net->crops_buf, crop_begin, crop_end, out_band, y0, y1, false, st))) return rc;
  return scratch_release(net, st);
}

// Band rows that are final once every crop with index < `upto` of the range [cb, ce) has been forwarded.
static int rows_final(const GridGeom& g, int cb, int ce, int upto) {
  int y0, y1;
  band_of(g, cb, ce, &y0, &y1);
  if (upto >= ce) return y1;
  if (upto <= cb) return y0;
  return std::min(y1, std::max(y0, g.stride * (upto / g.nx)));
}

int nind_tiled_denoise_step(nind_net* net, const float* img_chw, float* out_img, int height, int width, int cs,
                            int ucs, int ol, int crop_begin, int crop_end, int step_begin, int step_end,
                            int* rows_begin, int* rows_end, void* stream) {
  if (!net || !img_chw || !out_img) return fail(NIND_E_INVALID, "null argument");
  ENTER(net);
  GridGeom g;
  int rc;
  if ((rc = check_range(width, height, cs, ucs, ol, crop_begin, crop_end, &g))) return rc;
  if (step_begin < crop_begin || step_end > crop_end || step_begin >= step_end)
    return fail(NIND_E_INVALID, "illegal step range");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int n = crop_end - crop_begin;
  Plan* plan = nullptr;
  if ((rc = get_plan(net, step_end - step_begin, cs, cs, &plan))) return rc;
  if ((rc = scratch_acquire(net, st))) return rc;
  if ((rc = ensure(reinterpret_cast<void**>(&net->crops_buf), &net->crops_cap, (size_t)n * 3 * cs * cs * sizeof(float))))
    return rc;
  if ((rc = upload_origins(net, g, crop_begin, n, st))) return rc;
  GatherParams gp;
  memset(&gp, 0, sizeof gp);
  gp.src = img_chw; gp.src_img = 0; gp.src_plane = (long long)height * width; gp.src_w = width; gp.src_h = height;
  gp.origin = net->origin_buf + (step_begin - crop_begin);
  if ((rc = run_plan(net, plan, gp, net->crops_buf + (size_t)(step_begin - crop_begin) * 3 * cs * cs, st))) return rc;
  const int r0 = rows_final(g, crop_begin, crop_end, step_begin), r1 = rows_final(g, crop_begin, crop_end, step_end);
  if (rows_begin) *rows_begin = r0;
  if (rows_end) *rows_end = r1;
  if (r1 > r0 && (rc = launch_stitch(g, net->crops_buf, crop_begin, crop_end, out_img, r0, r1, true, st))) return rc;
  return scratch_release(net, st);
}

int nind_copy_planes(float* dst, long long dst_plane, const float* src, long long src_plane, int planes,
                     long long count, void* stream) {
  if (!dst || !src || count < 0 || planes < 0) return fail(NIND_E_INVALID, "illegal argument");
  if (count == 0 || planes == 0) return 0;
  for (int p = 0; p < planes; ++p)
    CUDA_TRY(cudaMemcpyAsync(dst + (size_t)p * dst_plane, src + (size_t)p * src_plane, (size_t)count * sizeof(float),
                             cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
  return 0;
}

// Peer memory: a buffer of one process (rank) mapped into the others of the node through CUDA IPC.  The opener
// maps it on ITS current device with lazy peer access, so plain device-to-device copies and kernel stores from
// that device go straight over NVLink.
int nind_peer_alloc(size_t bytes, void** ptr, unsigned char* handle64) {
  if (!ptr || !handle64 || !bytes) return fail(NIND_E_INVALID, "illegal argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  CUDA_TRY(cudaMalloc(ptr, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, *ptr);
  if (e != cudaSuccess) {
    cudaFree(*ptr);
    *ptr = nullptr;
    return fail(NIND_E_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
  }
  memcpy(handle64, &h, 64);
  return 0;
}

int nind_peer_open(const unsigned char* handle64, void** ptr) {
  if (!ptr || !handle64) return fail(NIND_E_INVALID, "null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  CUDA_TRY(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

int nind_peer_close(void* ptr) {
  if (!ptr) return 0;
  CUDA_TRY(cudaIpcCloseMemHandle(ptr));
  return 0;
}

int nind_peer_free(void* ptr) {
  if (!ptr) return 0;
  CUDA_TRY(cudaFree(ptr));
  return 0;
}

int nind_add_rows(float* dst, long long dst_plane, const float* src, long long src_plane, int planes, long long count,
                  void* stream) {
  if (!dst || !src || count < 0 || planes < 0) return fail(NIND_E_INVALID, "illegal argument");
  if (count == 0 || planes == 0) return 0;
  const bool vec = !((count | dst_plane | src_plane) & 3) &&
                   !((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15);
  if (vec)
    add_rows_kernel<true><<<grid_for(planes * (count / 4)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        dst, dst_plane, src, src_plane, planes, count);
  else
    add_rows_kernel<false><<<grid_for(planes * count), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        dst, dst_plane, src, src_plane, planes, count);
  ++g_launches;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// Pipeline steps of the host entry for crops [cb, ce): the first and the last step end / start at a grid-row
// boundary (so that compute starts after one grid row of the image has been uploaded and only one grid row of
// output is downloaded after the last forward), unless that would make them shorter than half a grid row — a
// tiny forward costs more than the copy it hides —; the crops in between run in balanced forwards of at most
// `batch` crops.  host_first / host_last (nind_set_option) override the two end steps.
static void host_steps(const nind_net* net, const GridGeom& g, int cb, int ce, int batch,
                       std::vector<std::pair<int, int>>* steps) {
  const int n = ce - cb;
  int first = 0, last = 0;
  if (n > batch || n > 2 * g.nx) {
    const int row_end = (cb / g.nx + 1) * g.nx;        // first grid-row boundary after cb
    const int row_begin = ((ce - 1) / g.nx) * g.nx;    // start of the grid row that holds the last crop
    first = row_end - cb;
    while (first < (g.nx + 1) / 2) first += g.nx;
    last = ce - row_begin;
    while (last < (g.nx + 1) / 2) last += g.nx;
    first = std::min(first, batch);
    last = std::min(last, batch);
  }
  if (net->host_first >= 0) first = net->host_first;
  if (net->host_last >= 0) last = net->host_last;
  if (first + last >= n) {  // short range: two steps (first | rest), or one
    if (first > 0 && first < n) last = n - first;
    else first = last = 0;
  }
  if (first) steps->push_back({cb, cb + first});
  balanced_steps(cb + first, ce - last, batch, steps);
  if (last) steps->push_back({ce - last, ce});
}

int nind_plan_steps(nind_net* net, int width, int height, int cs, int ucs, int ol, int crop_begin, int crop_end,
                    int batch, int* bounds, int max_bounds, int* n_steps) {
  if (!net || !n_steps) return fail(NIND_E_INVALID, "null argument");
  GridGeom g;
  int rc;
  if ((rc = check_range(width, height, cs, ucs, ol, crop_begin, crop_end, &g))) return rc;
  if (batch <= 0) return fail(NIND_E_INVALID, "batch must be positive");
  std::vector<std::pair<int, int>> steps;
  host_steps(net, g, crop_begin, crop_end, batch, &steps);
  *n_steps = (int)steps.size();
  if (bounds) {
    if (max_bounds < (int)steps.size() + 1) return fail(NIND_E_INVALID, "bounds array too small");
    for (size_t i = 0; i < steps.size(); ++i) bounds[i] = steps[i].first;
    bounds[steps.size()] = steps.back().second;
  }
  return 0;
}

// Enqueue crops [cb, ce) of one image on the three-stream host pipeline (no synchronisation).  Rows
// [d2h_y0, d2h_y1) of the stitched band are downloaded to `out_chw_host` as they complete; `d_out`
// (optional) receives the device image (full [3][H][W] layout) the band is stitched into.
static int enqueue_host_range_impl(nind_net* net, const float* img_chw_host, float* out_chw_host, int height,
                                   int width, int cs, int ucs, int ol, int batch, int cb, int ce, int d2h_y0,
                                   int d2h_y1, float** d_out, bool* enqueued) {
  // Pipelined over steps of crops in raster order: the H2D copy of the image rows a step needs, the
  // forward of its crops, the stitch of the output rows it completes and their D2H copy run on three
  // streams, so PCIe traffic hides behind compute when the host buffers are pinned.  Two device
  // (image, output) slots let image k+1's upload overlap image k's compute and download.
  GridGeom g;
  int rc;
  if ((rc = check_range(width, height, cs, ucs, ol, cb, ce, &g))) return rc;
  if (d2h_y0 < 0 || d2h_y1 > height || d2h_y0 > d2h_y1) return fail(NIND_E_INVALID, "illegal download row range");
  const size_t plane = (size_t)height * width;
  const size_t bytes = 3 * plane * sizeof(float);
  nind_net::HostSlot& S = net->slots[net->host_seq & 1];
  if ((rc = ensure(reinterpret_cast<void**>(&S.img), &S.img_cap, bytes))) return rc;
  if ((rc = ensure(reinterpret_cast<void**>(&S.out), &S.out_cap, bytes))) return rc;
  if (d_out) *d_out = S.out;
  const int n = ce - cb;
  if (!net->s_in) {
    CUDA_TRY(cudaStreamCreateWithFlags(&net->s_in, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&net->s_comp, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&net->s_out, cudaStreamNonBlocking));
  }
  if (!S.img_free) {
    CUDA_TRY(cudaEventCreateWithFlags(&S.img_free, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&S.out_free, cudaEventDisableTiming));
  }
  std::vector<std::pair<int, int>> steps;
  host_steps(net, g, cb, ce, batch, &steps);
  while (net->ev_in.size() < steps.size()) {
    cudaEvent_t a, b;
    CUDA_TRY(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
    net->ev_in.push_back(a);
    net->ev_done.push_back(b);
  }
  while (net->ev_rows.size() < steps.size()) {
    cudaEvent_t a;
    CUDA_TRY(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
    net->ev_rows.push_back(a);
  }
  net->rows_done.clear();
  // plans first: building one allocates (and zero-fills) its arena, which must not happen between enqueues
  std::vector<Plan*> plans(steps.size(), nullptr);
  for (size_t k = 0; k < steps.size(); ++k)
    if ((rc = get_plan(net, steps[k].second - steps[k].first, cs, cs, &plans[k]))) return rc;
  if ((rc = scratch_acquire(net, net->s_comp))) return rc;
  if ((rc = ensure(reinterpret_cast<void**>(&net->crops_buf), &net->crops_cap, (size_t)n * 3 * cs * cs * sizeof(float))))
    return rc;
  if ((rc = upload_origins(net, g, cb, n, net->s_comp))) return rc;
  *enqueued = true;
  if (S.used) {  // the image that used this slot two calls ago must be done with it
    CUDA_TRY(cudaStreamWaitEvent(net->s_in, S.img_free, 0));
    CUDA_TRY(cudaStreamWaitEvent(net->s_comp, S.out_free, 0));
  }
  // H2D: the image rows each step newly needs, in increasing order (planar image -> one 2-D copy of 3
  // plane segments).  Rows the mirror padding reflects to lie inside the range: at the top they are the
  // first rows; at the bottom edge (window overshoots the image) they can lie above the window's own
  // first row, so the range starts there.
  const int y_first = g.stride * (cb / g.nx) - g.pad, y_last_end = g.stride * ((ce - 1) / g.nx) - g.pad + cs;
  int rows_lo = std::max(0, y_first), rows_hi = std::min(height, y_last_end);
  if (y_last_end > height) rows_lo = std::min(rows_lo, std::max(0, 2 * height - y_last_end));
  if (y_first < 0) rows_hi = std::max(rows_hi, std::min(height, -y_first));
  int uploaded = rows_lo;
  for (size_t k = 0; k < steps.size(); ++k) {
    const int yi = (steps[k].second - 1) / g.nx;
    const int need = std::min(rows_hi, std::max(g.stride * yi - g.pad + cs, y_first < 0 ? -y_first : 0));
    const int upto = k + 1 == steps.size() ? rows_hi : std::max(uploaded, need);
    if (upto > uploaded) {
      const size_t off = (size_t)uploaded * width;
      CUDA_TRY(cudaMemcpy2DAsync(S.img + off, plane * sizeof(float), img_chw_host + off, plane * sizeof(float),
                                 (size_t)(upto - uploaded) * width * sizeof(float), 3, cudaMemcpyHostToDevice,
                                 net->s_in));
      uploaded = upto;
    }
    CUDA_TRY(cudaEventRecord(net->ev_in[k], net->s_in));
  }
  int y0, y1;
  band_of(g, cb, ce, &y0, &y1);
  int done = y0;  // band rows [y0, done) are final
  for (size_t k = 0; k < steps.size(); ++k) {
    const int ia = steps[k].first, ib = steps[k].second;
    CUDA_TRY(cudaStreamWaitEvent(net->s_comp, net->ev_in[k], 0));
    GatherParams gp;
    memset(&gp, 0, sizeof gp);
    gp.src = S.img; gp.src_img = 0; gp.src_plane = (long long)plane; gp.src_w = width; gp.src_h = height;
    gp.origin = net->origin_buf + (ia - cb);
    if ((rc = run_plan(net, plans[k], gp, net->crops_buf + (size_t)(ia - cb) * 3 * cs * cs, net->s_comp))) return rc;
    if (ib == ce) CUDA_TRY(cudaEventRecord(S.img_free, net->s_comp));
    // output rows completed by this step: every crop of the range that touches them has index < ib
    const int r0 = done;
    const int r1 = ib == ce ? y1 : std::min(y1, std::max(done, g.stride * (ib / g.nx)));
    if (r1 > r0) {
      if ((rc = launch_stitch(g, net->crops_buf, cb, ce, S.out, r0, r1, true, net->s_comp))) return rc;
      done = r1;
      const int c0 = std::max(r0, d2h_y0), c1 = std::min(r1, d2h_y1);
      if (c1 > c0) {
        CUDA_TRY(cudaEventRecord(net->ev_done[k], net->s_comp));
        CUDA_TRY(cudaStreamWaitEvent(net->s_out, net->ev_done[k], 0));
        const size_t off = (size_t)c0 * width;
        CUDA_TRY(cudaMemcpy2DAsync(out_chw_host + off, plane * sizeof(float), S.out + off, plane * sizeof(float),
                                   (size_t)(c1 - c0) * width * sizeof(float), 3, cudaMemcpyDeviceToHost, net->s_out));
      }
    }
    CUDA_TRY(cudaEventRecord(net->ev_rows[k], net->s_comp));
    net->rows_done.push_back(done);
  }
  if ((rc = scratch_release(net, net->s_comp))) return rc;
  CUDA_TRY(cudaEventRecord(S.out_free, net->s_out));
  S.used = true;
  ++net->host_seq;
  return 0;
}

static int enqueue_host_range(nind_net* net, const float* img_chw_host, float* out_chw_host, int height, int width,
                              int cs, int ucs, int ol, int batch, int cb, int ce, int d2h_y0, int d2h_y1,
                              float** d_out) {
  if (!net || !img_chw_host || !out_chw_host) return fail(NIND_E_INVALID, "null argument");
  if (batch <= 0) return fail(NIND_E_INVALID, "batch must be positive");
  ENTER(net);
  bool enqueued = false;
  const int rc = enqueue_host_range_impl(net, img_chw_host, out_chw_host, height, width, cs, ucs, ol, batch, cb, ce,
                                         d2h_y0, d2h_y1, d_out, &enqueued);
  if (rc && enqueued) {
    // Part of the image is already in flight: wait for it, so that the caller may free its (pinned) buffers
    // as soon as it sees the error, and leave the slot bookkeeping as it was (host_seq only advances on success).
    const std::string keep = g_err;
    cudaStreamSynchronize(net->s_in);
    cudaStreamSynchronize(net->s_comp);
    cudaStreamSynchronize(net->s_out);
    g_err = keep;
  }
  return rc;
}

static int enqueue_host_image(nind_net* net, const float* img_chw_host, float* out_chw_host, int height, int width,
                              int cs, int ucs, int ol, int batch) {
  GridGeom g;
  if (!make_grid(width, height, cs, ucs, ol, &g)) return fail(NIND_E_INVALID, "illegal crop geometry");
  return enqueue_host_range(net, img_chw_host, out_chw_host, height, width, cs, ucs, ol, batch, 0, g.size(), 0,
                            height, nullptr);
}

int nind_host_sync(nind_net* net) {
  if (!net) return fail(NIND_E_INVALID, "null handle");
  DeviceGuard guard(net->device);
  if (net->s_in) {
    CUDA_TRY(cudaStreamSynchronize(net->s_out));
    CUDA_TRY(cudaStreamSynchronize(net->s_comp));
    CUDA_TRY(cudaStreamSynchronize(net->s_in));
  }
  return check_err_flag(net);
}

int nind_tiled_denoise_host_range(nind_net* net, const float* img_chw_host, float* out_chw_host, int height,
                                  int width, int cs, int ucs, int ol, int batch, int crop_begin, int crop_end,
                                  int d2h_y0, int d2h_y1, float** d_out) {
  return enqueue_host_range(net, img_chw_host, out_chw_host, height, width, cs, ucs, ol, batch, crop_begin, crop_end,
                            d2h_y0, d2h_y1, d_out);
}

int nind_host_join(nind_net* net, void* stream) {
  if (!net) return fail(NIND_E_INVALID, "null handle");
  ENTER(net);
  if (!net->s_comp) return 0;
  if (!net->ev_join) CUDA_TRY(cudaEventCreateWithFlags(&net->ev_join, cudaEventDisableTiming));
  CUDA_TRY(cudaEventRecord(net->ev_join, net->s_comp));
  CUDA_TRY(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), net->ev_join, 0));
  return 0;
}

int nind_host_join_rows(nind_net* net, int y, void* stream) {
  if (!net) return fail(NIND_E_INVALID, "null handle");
  ENTER(net);
  for (size_t k = 0; k < net->rows_done.size(); ++k)
    if (net->rows_done[k] >= y) {
      CUDA_TRY(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), net->ev_rows[k], 0));
      return 0;
    }
  return fail(NIND_E_INVALID, "nind_host_join_rows: the last host-range call does not complete that row");
}

int nind_host_rows_done(nind_net* net, int* rows, int max_rows, int* n) {
  if (!net || !n) return fail(NIND_E_INVALID, "null argument");
  *n = (int)net->rows_done.size();
  if (rows) {
    if (max_rows < *n) return fail(NIND_E_INVALID, "rows array too small");
    for (int k = 0; k < *n; ++k) rows[k] = net->rows_done[k];
  }
  return 0;
}

int nind_host_register(void* ptr, size_t bytes) {
  if (!ptr || !bytes) return fail(NIND_E_INVALID, "null argument");
  CUDA_TRY(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
  return 0;
}

int nind_host_unregister(void* ptr) {
  if (!ptr) return fail(NIND_E_INVALID, "null argument");
  CUDA_TRY(cudaHostUnregister(ptr));
  return 0;
}

int nind_tiled_denoise_host_async(nind_net* net, const float* img_chw_host, float* out_chw_host, int height,
                                  int width, int cs, int ucs, int ol, int batch) {
  return enqueue_host_image(net, img_chw_host, out_chw_host, height, width, cs, ucs, ol, batch);
}

int nind_tiled_denoise_host(nind_net* net, const float* img_chw_host, float* out_chw_host, int height,
                            int width, int cs, int ucs, int ol, int batch) {
  int rc = enqueue_host_image(net, img_chw_host, out_chw_host, height, width, cs, ucs, ol, batch);
  if (rc) return rc;
  return nind_host_sync(net);
}

// ---------------------------------------------------------------- file formats either side of the path
int nind_image_to_chw_f32(const void* src_hwc, int dtype, int height, int width, int bgr, float* dst_chw,
                          void* stream) {
  if (!src_hwc || !dst_chw) return fail(NIND_E_INVALID, "null argument");
  if (dtype < NIND_PIX_U8 || dtype > NIND_PIX_F32) return fail(NIND_E_INVALID, "unknown pixel type");
  if (height <= 0 || width <= 0) return fail(NIND_E_INVALID, "illegal image size");
  PixConvParams p;
  p.hwc = const_cast<void*>(src_hwc); p.chw = dst_chw; p.h = height; p.w = width; p.dtype = dtype; p.bgr = bgr ? 1 : 0;
  image_to_chw_kernel<<<grid_for((long long)height * width), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  ++g_launches;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int nind_chw_f32_to_image(const float* src_chw, int height, int width, int dtype, int bgr, void* dst_hwc,
                          void* stream) {
  if (!src_chw || !dst_hwc) return fail(NIND_E_INVALID, "null argument");
  if (dtype < NIND_PIX_U8 || dtype > NIND_PIX_F32) return fail(NIND_E_INVALID, "unknown pixel type");
  if (height <= 0 || width <= 0) return fail(NIND_E_INVALID, "illegal image size");
  PixConvParams p;
  p.hwc = dst_hwc; p.chw = const_cast<float*>(src_chw); p.h = height; p.w = width; p.dtype = dtype; p.bgr = bgr ? 1 : 0;
  chw_to_image_kernel<<<grid_for((long long)height * width), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  ++g_launches;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // extern "C"
