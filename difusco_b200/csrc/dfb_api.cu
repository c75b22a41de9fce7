// C-ABI of difusco_b200 (include/difusco_b200.h): context, weight packing, graph preparation and
// the orchestration of one forward / one denoise step / the whole denoise loop.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/difusco_b200.h"
#include "common.cuh"
#include "edge_layer_fp32.cuh"
#include "edge_layer_tc.cuh"
#include "edge_layer_v2.cuh"
#include "kernels_small.cuh"
#include "knn.cuh"
#include "tsp_decode.cuh"

using namespace dfb;

static std::string g_create_error;

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

struct dfb_ctx {
  int device = 0;
  std::string err;
  // ---- model ----
  bool weights_loaded = false;
  int L = 0, out_channels = 0, node_only = 0;
  int agg_mode = AGG_SUM;
  int edge_impl = DFB_EDGE_IMPL_TC;
  DevBuf wbuf, wbuf16, layers_dev;
  std::vector<LayerParams> layers;
  TimeParams tp{};
  HeadParams hp{};
  const float *Wt_node = nullptr, *b_node = nullptr, *Wt_edge = nullptr, *b_edge = nullptr;
  const float *dimt128 = nullptr, *dimt256 = nullptr;
  float* lut = nullptr;   // [2][256] categorical edge-embedding LUT (inside wbuf)
  float* cl0 = nullptr;   // [2][256] C_0 * lut (layer 0's GEMM1 by table lookup); followed by [2][256] zeros for the MIS e0 = 0 case
  // ---- graph ----
  bool graph_ready = false, points_ready = false;
  GraphDev g{};
  int gn_segments = 1;
  int max_seg = 0;   // most node segments inside one 32-edge group (pair kernel handles up to v2::MAXSEG)
  DevBuf d_row, d_col, d_perm, d_rowptr, d_grp_first, d_grp_pair, d_ei_stage;
  // ---- workspace ----
  DevBuf e, h, h0, uvab, uvab0, partials, feat, tvec, tvals, gn_part, gn_stats, d_points, d_xt, d_u;
  DevBuf opt_points, opt_tours, opt_pos, opt_dnext, opt_cand, opt_tiles, opt_state, opt_best;   // 2-opt (row f3)
  int tvec_steps_cap = 0;
  // ---- step staging (pinned) + captured loop ----
  // dfb_denoise_step / dfb_denoise never allocate, never synchronise the host with the stream and never touch the
  // heap after the first call of a shape: timesteps and posterior constants go through two pinned staging slots
  // (guarded by an event each), the per-step table lives in device memory, and the whole loop is replayed as ONE
  // CUDA graph that is re-captured only when the shape / buffers / implementation change.
  static constexpr int STAGE_SLOTS = 2, MAX_STEPS = 4096;
  float* h_tvals[STAGE_SLOTS] = {nullptr, nullptr};
  StepParams* h_steps[STAGE_SLOTS] = {nullptr, nullptr};
  cudaEvent_t stage_ev[STAGE_SLOTS] = {nullptr, nullptr};
  int stage_next = 0;
  DevBuf d_steps;
  uint64_t buf_gen = 0;          // bumped whenever a device buffer is (re)allocated or the graph / weights change
  bool capture_enabled = true;
  bool capture_broken = false;
  cudaGraphExec_t loop_exec = nullptr;
  cudaStream_t loop_stream = nullptr;   // the captured loop runs on the library's own stream (the caller's may be the
  cudaEvent_t loop_in = nullptr, loop_out = nullptr;   // legacy default stream, which cannot be captured), fenced by events
  struct LoopKey {
    uint64_t buf_gen = ~0ull;
    int steps = 0, diffusion = 0, impl = 0, agg = 0, pair = 0;
    const void* uniforms = nullptr;
    bool operator==(const LoopKey& o) const {
      return buf_gen == o.buf_gen && steps == o.steps && diffusion == o.diffusion && impl == o.impl && agg == o.agg &&
             pair == o.pair && uniforms == o.uniforms;
    }
  } loop_key;
  int64_t loop_launches = 0;     // kernel launches inside one replay of the captured loop
  // ---- accounting ----
  int64_t launches = 0;
  bool profiling = false;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  TcState tc;
  v2::State pair;   // round-2 CTA-pair kernel (middle layers of the product path)
  bool pair_enabled = true;    // DFB_PAIR_KERNEL=0 routes every layer to the single-CTA kernel (A/B timing)
  bool serpentine = true;      // DFB_SERPENTINE=0: every layer sweeps the edge stream upwards (A/B timing)
};

#define FAIL(ctx, code, ...)                         \
  do {                                               \
    char _b[512];                                    \
    snprintf(_b, sizeof(_b), __VA_ARGS__);           \
    (ctx)->err = _b;                                 \
    return (code);                                   \
  } while (0)

#define CK(ctx, call)                                                                       \
  do {                                                                                      \
    cudaError_t _e = (call);                                                                \
    if (_e != cudaSuccess)                                                                  \
      FAIL(ctx, DFB_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define CKL(ctx)                                                                           \
  do {                                                                                      \
    (ctx)->launches++;                                                                      \
    cudaError_t _e = cudaGetLastError();                                                    \
    if (_e != cudaSuccess)                                                                  \
      FAIL(ctx, DFB_E_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

static int ensure(dfb_ctx* ctx, DevBuf& b, size_t bytes) {
  if (bytes <= b.cap) return DFB_OK;
  ctx->buf_gen++;   // a captured loop that baked the old pointer is stale
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
  size_t want = bytes + (bytes >> 3);   // slack so slightly larger graphs do not reallocate
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) {
    e = cudaMalloc(&b.p, bytes);
    want = bytes;
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    FAIL(ctx, DFB_E_NOMEM, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
  }
  b.cap = want;
  return DFB_OK;
}
#define ENS(ctx, buf, bytes)                     \
  do {                                           \
    int _r = ensure(ctx, buf, bytes);            \
    if (_r) return _r;                           \
  } while (0)

static bool is_device_ptr(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// ================================================================================================
extern "C" int dfb_abi_version(void) { return DFB_ABI_VERSION; }

extern "C" const char* dfb_last_error(const dfb_ctx* ctx) {
  return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

extern "C" int dfb_create(dfb_ctx** out, int device) {
  if (!out) return DFB_E_INVALID;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) +
                     " (difusco_b200 has no CPU fallback)";
    cudaGetLastError();
    return DFB_E_CUDA;
  }
  if (device < 0 || device >= n) {
    g_create_error = "device index out of range";
    return DFB_E_INVALID;
  }
  if ((e = cudaSetDevice(device)) != cudaSuccess) {
    g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
    return DFB_E_CUDA;
  }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major != 10) {
    char b[160];
    snprintf(b, sizeof(b), "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
             prop.major, prop.minor);
    g_create_error = b;
    return DFB_E_UNSUPPORTED;
  }
  dfb_ctx* ctx = new dfb_ctx();
  ctx->device = device;
  e = cudaFuncSetAttribute(k_edge_layer_fp32, cudaFuncAttributeMaxDynamicSharedMemorySize, EF_SMEM);
  if (e != cudaSuccess) {
    g_create_error = std::string("cudaFuncSetAttribute(fp32 edge kernel): ") + cudaGetErrorString(e);
    delete ctx;
    return DFB_E_CUDA;
  }
  int r = tc_init(&ctx->tc, prop.multiProcessorCount);
  if (r != 0) {
    g_create_error = "tcgen05 edge kernel setup failed: " + ctx->tc.err;
    delete ctx;
    return DFB_E_CUDA;
  }
  for (int i = 0; i < dfb_ctx::STAGE_SLOTS; ++i) {
    if ((e = cudaHostAlloc((void**)&ctx->h_tvals[i], dfb_ctx::MAX_STEPS * sizeof(float), cudaHostAllocDefault)) != cudaSuccess ||
        (e = cudaHostAlloc((void**)&ctx->h_steps[i], dfb_ctx::MAX_STEPS * sizeof(StepParams), cudaHostAllocDefault)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&ctx->stage_ev[i], cudaEventDisableTiming)) != cudaSuccess) {
      g_create_error = std::string("pinned staging: ") + cudaGetErrorString(e);
      delete ctx;
      return DFB_E_CUDA;
    }
  }
  if ((e = cudaStreamCreateWithFlags(&ctx->loop_stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&ctx->loop_in, cudaEventDisableTiming)) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&ctx->loop_out, cudaEventDisableTiming)) != cudaSuccess) {
    g_create_error = std::string("loop stream: ") + cudaGetErrorString(e);
    delete ctx;
    return DFB_E_CUDA;
  }
  {
    const char* cg = getenv("DFB_GRAPH_CAPTURE");
    if (cg && atoi(cg) == 0) ctx->capture_enabled = false;
  }
  {
    const char* pk = getenv("DFB_PAIR_KERNEL");
    if (pk) ctx->pair_enabled = atoi(pk) != 0;
    const char* sp = getenv("DFB_SERPENTINE");
    if (sp) ctx->serpentine = atoi(sp) != 0;
  }
  r = v2::init(&ctx->pair, &ctx->tc);
  if (r != 0) {
    g_create_error = "tcgen05 pair kernel setup failed: " + ctx->tc.err;
    delete ctx;
    return DFB_E_CUDA;
  }
  *out = ctx;
  return DFB_OK;
}

extern "C" int dfb_destroy(dfb_ctx* ctx) {
  if (!ctx) return DFB_OK;
  cudaSetDevice(ctx->device);
  DevBuf* bufs[] = {&ctx->wbuf, &ctx->wbuf16, &ctx->layers_dev, &ctx->d_row, &ctx->d_col, &ctx->d_perm,
                    &ctx->d_rowptr, &ctx->d_grp_first, &ctx->d_grp_pair, &ctx->d_ei_stage, &ctx->e, &ctx->h,
                    &ctx->h0, &ctx->uvab, &ctx->uvab0, &ctx->partials, &ctx->feat, &ctx->tvec, &ctx->tvals,
                    &ctx->gn_part, &ctx->gn_stats, &ctx->d_points, &ctx->d_xt, &ctx->d_u, &ctx->opt_points, &ctx->opt_tours,
                    &ctx->opt_pos, &ctx->opt_dnext, &ctx->opt_cand, &ctx->opt_tiles, &ctx->opt_state, &ctx->opt_best};
  for (DevBuf* b : bufs)
    if (b->p) cudaFree(b->p);
  for (cudaEvent_t ev : ctx->ev_pool) cudaEventDestroy(ev);
  if (ctx->loop_exec) cudaGraphExecDestroy(ctx->loop_exec);
  if (ctx->loop_stream) cudaStreamDestroy(ctx->loop_stream);
  if (ctx->loop_in) cudaEventDestroy(ctx->loop_in);
  if (ctx->loop_out) cudaEventDestroy(ctx->loop_out);
  if (ctx->d_steps.p) cudaFree(ctx->d_steps.p);
  for (int i = 0; i < dfb_ctx::STAGE_SLOTS; ++i) {
    if (ctx->h_tvals[i]) cudaFreeHost(ctx->h_tvals[i]);
    if (ctx->h_steps[i]) cudaFreeHost(ctx->h_steps[i]);
    if (ctx->stage_ev[i]) cudaEventDestroy(ctx->stage_ev[i]);
  }
  tc_destroy(&ctx->tc);
  delete ctx;
  return DFB_OK;
}

extern "C" int dfb_set_edge_impl(dfb_ctx* ctx, int impl) {
  if (!ctx) return DFB_E_INVALID;
  if (impl != DFB_EDGE_IMPL_TC && impl != DFB_EDGE_IMPL_FP32 && impl != DFB_EDGE_IMPL_TC1) FAIL(ctx, DFB_E_INVALID, "unknown edge impl %d", impl);
  ctx->edge_impl = impl;
  return DFB_OK;
}

extern "C" int64_t dfb_launch_count(const dfb_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ================================================================================================
// weights
// ================================================================================================
static uint16_t f2bf16_rn(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  uint32_t r = u + 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(r >> 16);
}
static float bf16_to_f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

extern "C" int dfb_load_weights(dfb_ctx* ctx, int n_layers, int hidden_dim, int out_channels,
                                int node_feature_only, int n_tensors, const char* const* names,
                                const float* const* tensors, const int64_t* numels) {
  if (!ctx) return DFB_E_INVALID;
  CK(ctx, cudaSetDevice(ctx->device));
  if (hidden_dim != H) FAIL(ctx, DFB_E_UNSUPPORTED, "hidden_dim %d: kernels are specialised for 256", hidden_dim);
  if (n_layers < 1 || n_layers > 64) FAIL(ctx, DFB_E_INVALID, "n_layers %d out of range", n_layers);
  if (out_channels != 1 && out_channels != 2) FAIL(ctx, DFB_E_INVALID, "out_channels must be 1 or 2");
  std::map<std::string, std::pair<const float*, int64_t>> sd;
  for (int i = 0; i < n_tensors; ++i) {
    std::string k = names[i];
    if (k.rfind("model.", 0) == 0) k = k.substr(6);
    sd[k] = {tensors[i], numels[i]};
  }
  auto get = [&](const std::string& k, int64_t n, const float** out) -> int {
    auto it = sd.find(k);
    if (it == sd.end()) FAIL(ctx, DFB_E_INVALID, "state_dict key '%s' missing", k.c_str());
    if (it->second.second != n)
      FAIL(ctx, DFB_E_INVALID, "state_dict key '%s' has %lld elements, expected %lld", k.c_str(),
           (long long)it->second.second, (long long)n);
    *out = it->second.first;
    return DFB_OK;
  };
#define GET(k, n, out)                 \
  do {                                 \
    int _r = get(k, n, out);           \
    if (_r) return _r;                 \
  } while (0)

  const int L = n_layers;
  // ---- fp32 arena layout ----
  std::vector<float> arena;
  auto put = [&](size_t n) {
    size_t off = arena.size();
    arena.resize(off + ((n + 63) / 64) * 64, 0.0f);   // 256-byte aligned slots
    return off;
  };
  auto put_T = [&](const float* W, int out_f, int in_f) {   // [out][in] -> in-major [in][out]
    size_t off = put((size_t)out_f * in_f);
    for (int o = 0; o < out_f; ++o)
      for (int i = 0; i < in_f; ++i) arena[off + (size_t)i * out_f + o] = W[(size_t)o * in_f + i];
    return off;
  };
  auto put_v = [&](const float* v, int n) {
    size_t off = put(n);
    memcpy(&arena[off], v, n * sizeof(float));
    return off;
  };
  struct LOff {
    size_t Wt_uvab, b_uvab, Wt_C, Wt_O, b_O, hg, hb, eg, eb, og, ob, Wt_tau, b_tau;
  };
  std::vector<LOff> lo(L);
  std::vector<uint16_t> arena16((size_t)(L * 12 + 4) * H * H);
  const float* p;
  for (int l = 0; l < L; ++l) {
    std::string pre = "layers." + std::to_string(l) + ".";
    const float *W[4], *b[4], *WC, *bC;
    const char* nm[4] = {"U", "V", "A", "B"};
    for (int q = 0; q < 4; ++q) {
      GET(pre + nm[q] + ".weight", H * H, &W[q]);
      GET(pre + nm[q] + ".bias", H, &b[q]);
    }
    GET(pre + "C.weight", H * H, &WC);
    GET(pre + "C.bias", H, &bC);
    lo[l].Wt_uvab = put((size_t)H * 4 * H);
    lo[l].b_uvab = put(4 * H);
    for (int q = 0; q < 4; ++q) {
      for (int o = 0; o < H; ++o) {
        for (int i = 0; i < H; ++i) arena[lo[l].Wt_uvab + (size_t)i * 4 * H + q * H + o] = W[q][(size_t)o * H + i];
        arena[lo[l].b_uvab + q * H + o] = b[q][o] + (q == 3 ? bC[o] : 0.0f);
      }
    }
    lo[l].Wt_C = put_T(WC, H, H);
    GET(pre + "norm_h.weight", H, &p); lo[l].hg = put_v(p, H);
    GET(pre + "norm_h.bias", H, &p);   lo[l].hb = put_v(p, H);
    GET(pre + "norm_e.weight", H, &p); lo[l].eg = put_v(p, H);
    GET(pre + "norm_e.bias", H, &p);   lo[l].eb = put_v(p, H);
    std::string po = "per_layer_out." + std::to_string(l) + ".";
    const float* WO;
    GET(po + "0.weight", H, &p); lo[l].og = put_v(p, H);
    GET(po + "0.bias", H, &p);   lo[l].ob = put_v(p, H);
    GET(po + "2.weight", H * H, &WO); lo[l].Wt_O = put_T(WO, H, H);
    GET(po + "2.bias", H, &p);   lo[l].b_O = put_v(p, H);
    std::string pt = "time_embed_layers." + std::to_string(l) + ".1.";
    GET(pt + "weight", H * TE, &p); lo[l].Wt_tau = put_T(p, H, TE);
    GET(pt + "bias", H, &p);        lo[l].b_tau = put_v(p, H);
    // bf16 hi/lo split of C and O, [out][in] K-major (the tensor-core B operand)
    uint16_t* a16 = &arena16[(size_t)l * 12 * H * H];
    for (int i = 0; i < H * H; ++i) {
      uint16_t hi = f2bf16_rn(WC[i]);
      a16[i] = hi;
      a16[H * H + i] = f2bf16_rn(WC[i] - bf16_to_f(hi));
      hi = f2bf16_rn(WO[i]);
      a16[2 * H * H + i] = hi;
      a16[3 * H * H + i] = f2bf16_rn(WO[i] - bf16_to_f(hi));
      for (int qq = 0; qq < 4; ++qq) {   // U, V, A, B: [out][in] K-major, hi then lo
        hi = f2bf16_rn(W[qq][i]);
        a16[(size_t)(4 + 2 * qq) * H * H + i] = hi;
        a16[(size_t)(5 + 2 * qq) * H * H + i] = f2bf16_rn(W[qq][i] - bf16_to_f(hi));
      }
    }
  }
  size_t o_node_W, o_node_b, o_edge_W, o_edge_b, o_t0W, o_t0b, o_t2W, o_t2b, o_gng, o_gnb, o_outW, o_outb;
  GET("node_embed.weight", H * H, &p); o_node_W = put_T(p, H, H);
  for (int i = 0; i < H * H; ++i) {
    uint16_t hi = f2bf16_rn(p[i]);
    arena16[(size_t)(L * 12 + 2) * H * H + i] = hi;
    arena16[(size_t)(L * 12 + 3) * H * H + i] = f2bf16_rn(p[i] - bf16_to_f(hi));
  }
  GET("node_embed.bias", H, &p);       o_node_b = put_v(p, H);
  GET("edge_embed.weight", H * H, &p); o_edge_W = put_T(p, H, H);
  for (int i = 0; i < H * H; ++i) {
    uint16_t hi = f2bf16_rn(p[i]);
    arena16[(size_t)(L * 12 + 0) * H * H + i] = hi;
    arena16[(size_t)(L * 12 + 1) * H * H + i] = f2bf16_rn(p[i] - bf16_to_f(hi));
  }
  GET("edge_embed.bias", H, &p);       o_edge_b = put_v(p, H);
  GET("time_embed.0.weight", TE * H, &p); o_t0W = put_T(p, TE, H);
  GET("time_embed.0.bias", TE, &p);       o_t0b = put_v(p, TE);
  GET("time_embed.2.weight", TE * TE, &p); o_t2W = put_T(p, TE, TE);
  GET("time_embed.2.bias", TE, &p);        o_t2b = put_v(p, TE);
  GET("out.0.weight", H, &p); o_gng = put_v(p, H);
  GET("out.0.bias", H, &p);   o_gnb = put_v(p, H);
  GET("out.2.weight", out_channels * H, &p); o_outW = put_v(p, out_channels * H);
  GET("out.2.bias", out_channels, &p);       o_outb = put_v(p, out_channels);

  // frequency tables: computed by the Python host with the reference's own torch expressions and
  // passed as pseudo-tensors when available (bit-identical tables); otherwise computed here.
  size_t o_freqs = put(TE), o_d128 = put(TE), o_d256 = put(H), o_lut = put(2 * H), o_cl0 = put(4 * H);
  auto it = sd.find("__const.time_freqs");
  for (int m = 0; m < TE; ++m)
    arena[o_freqs + m] = (it != sd.end() && it->second.second == TE)
                             ? it->second.first[m]
                             : expf((-9.210340371976184f * (float)m) / (float)TE);
  it = sd.find("__const.dimt_pos");
  for (int m = 0; m < TE; ++m)
    arena[o_d128 + m] = (it != sd.end() && it->second.second == TE)
                            ? it->second.first[m]
                            : powf(10000.0f, (2.0f * (float)(m / 2)) / (float)TE);
  it = sd.find("__const.dimt_scalar");
  for (int m = 0; m < H; ++m)
    arena[o_d256 + m] = (it != sd.end() && it->second.second == H)
                            ? it->second.first[m]
                            : powf(10000.0f, (2.0f * (float)(m / 2)) / (float)H);

  ENS(ctx, ctx->wbuf, arena.size() * sizeof(float));
  ENS(ctx, ctx->wbuf16, arena16.size() * sizeof(uint16_t));
  ENS(ctx, ctx->layers_dev, L * sizeof(LayerParams));
  CK(ctx, cudaMemcpy(ctx->wbuf.p, arena.data(), arena.size() * sizeof(float), cudaMemcpyHostToDevice));
  CK(ctx, cudaMemcpy(ctx->wbuf16.p, arena16.data(), arena16.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
  const float* base = (const float*)ctx->wbuf.p;
  const uint16_t* base16 = (const uint16_t*)ctx->wbuf16.p;
  ctx->layers.resize(L);
  for (int l = 0; l < L; ++l) {
    LayerParams& lp = ctx->layers[l];
    lp.Wt_uvab = base + lo[l].Wt_uvab; lp.b_uvab = base + lo[l].b_uvab;
    lp.Wt_C = base + lo[l].Wt_C; lp.Wt_O = base + lo[l].Wt_O; lp.b_O = base + lo[l].b_O;
    lp.ln_h_g = base + lo[l].hg; lp.ln_h_b = base + lo[l].hb;
    lp.ln_e_g = base + lo[l].eg; lp.ln_e_b = base + lo[l].eb;
    lp.ln_o_g = base + lo[l].og; lp.ln_o_b = base + lo[l].ob;
    lp.Wt_tau = base + lo[l].Wt_tau; lp.b_tau = base + lo[l].b_tau;
    lp.C_hi = base16 + (size_t)l * 12 * H * H; lp.C_lo = lp.C_hi + H * H;
    lp.O_hi = lp.C_hi + 2 * H * H;            lp.O_lo = lp.C_hi + 3 * H * H;
  }
  CK(ctx, cudaMemcpy(ctx->layers_dev.p, ctx->layers.data(), L * sizeof(LayerParams), cudaMemcpyHostToDevice));
  ctx->Wt_node = base + o_node_W; ctx->b_node = base + o_node_b;
  ctx->Wt_edge = base + o_edge_W; ctx->b_edge = base + o_edge_b;
  ctx->tp.freqs = base + o_freqs; ctx->tp.Wt0 = base + o_t0W; ctx->tp.b0 = base + o_t0b;
  ctx->tp.Wt2 = base + o_t2W;     ctx->tp.b2 = base + o_t2b;
  ctx->hp.gn_g = base + o_gng; ctx->hp.gn_b = base + o_gnb; ctx->hp.W = base + o_outW; ctx->hp.b = base + o_outb;
  ctx->hp.out_channels = out_channels;
  ctx->dimt128 = base + o_d128; ctx->dimt256 = base + o_d256;
  ctx->lut = (float*)ctx->wbuf.p + o_lut;
  ctx->cl0 = (float*)ctx->wbuf.p + o_cl0;   // second half stays zero (arena slots are zero-initialised)
  ctx->L = L; ctx->out_channels = out_channels; ctx->node_only = node_feature_only;

  int r = tc_bind_weights(&ctx->tc, ctx->layers.data(), L);
  if (r) FAIL(ctx, DFB_E_CUDA, "tensor-map setup failed: %s", ctx->tc.err.c_str());
  r = v2::bind_weights(&ctx->pair, &ctx->tc, ctx->wbuf16.p, L);
  if (r) FAIL(ctx, DFB_E_CUDA, "tensor-map setup failed: %s", ctx->tc.err.c_str());

  // categorical edge-embedding LUT: edge_embed(edge_pos_embed(x)) for x in {0, 1}
  if (!node_feature_only) {
    ENS(ctx, ctx->feat, (size_t)LIN_ROWS * H * sizeof(float));
    float x01[2] = {0.0f, 1.0f};
    ENS(ctx, ctx->tvals, 4096 * sizeof(float));
    CK(ctx, cudaMemcpy(ctx->tvals.p, x01, sizeof(x01), cudaMemcpyHostToDevice));
    k_scalar_features<<<2, H>>>((const float*)ctx->tvals.p, nullptr, ctx->dimt256, (float*)ctx->feat.p, 2);
    CKL(ctx);
    k_linear<<<dim3(1, 1), 256>>>((const float*)ctx->feat.p, ctx->Wt_edge, ctx->b_edge, ctx->lut, 2, H);
    CKL(ctx);
    // layer 0 never needs its GEMM1: C_0 applied to the two possible input rows (fp32 FFMA; b_C rides in B h's bias)
    k_linear<<<dim3(1, 1), 256>>>(ctx->lut, ctx->layers[0].Wt_C, nullptr, ctx->cl0, 2, H);
    CKL(ctx);
    CK(ctx, cudaDeviceSynchronize());
  }
  ctx->weights_loaded = true;
  ctx->buf_gen++;
  ctx->points_ready = false;
  return DFB_OK;
}

extern "C" int dfb_set_aggregation(dfb_ctx* ctx, int mode) {
  if (!ctx) return DFB_E_INVALID;
  if (mode < AGG_SUM || mode > AGG_MAX) FAIL(ctx, DFB_E_INVALID, "unknown aggregation %d", mode);
  ctx->agg_mode = mode;
  return DFB_OK;
}

// ================================================================================================
// graph
// ================================================================================================
extern "C" int dfb_prepare_graph(dfb_ctx* ctx, const int64_t* edge_index, int64_t V64, int64_t E64,
                                 int gn_segments, void* stream_) {
  if (!ctx) return DFB_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream_;
  CK(ctx, cudaSetDevice(ctx->device));
  if (!ctx->weights_loaded) FAIL(ctx, DFB_E_INVALID, "dfb_load_weights must be called first");
  if (V64 <= 0 || E64 <= 0 || V64 > 0x7fffffff / 4 || E64 > 0x7ffffff0)
    FAIL(ctx, DFB_E_INVALID, "bad graph size V=%lld E=%lld", (long long)V64, (long long)E64);
  const int V = (int)V64, E = (int)E64;
  if (gn_segments < 1) FAIL(ctx, DFB_E_INVALID, "gn_segments must be >= 1");
  {
    int R = ctx->node_only ? V : E;
    if (R % gn_segments) FAIL(ctx, DFB_E_INVALID, "gn_segments %d does not divide %d rows", gn_segments, R);
  }
  std::vector<int64_t> stage;
  const int64_t* ei = edge_index;
  if (is_device_ptr(edge_index)) {
    stage.resize((size_t)2 * E);
    CK(ctx, cudaMemcpyAsync(stage.data(), edge_index, (size_t)2 * E * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CK(ctx, cudaStreamSynchronize(st));
    ei = stage.data();
  }
  const int64_t* row64 = ei;
  const int64_t* col64 = ei + E;
  std::vector<int> rowptr((size_t)V + 1, 0);
  bool sorted = true;
  for (int s = 0; s < E; ++s) {
    int64_t r = row64[s], c = col64[s];
    if (r < 0 || r >= V || c < 0 || c >= V)
      FAIL(ctx, DFB_E_INVALID, "edge %d = (%lld,%lld) out of range for %d nodes", s, (long long)r, (long long)c, V);
    rowptr[(size_t)r + 1]++;
    if (s && r < row64[s - 1]) sorted = false;
  }
  for (int i = 0; i < V; ++i) rowptr[i + 1] += rowptr[i];
  std::vector<int> row(E), col(E), perm;
  if (sorted) {
    for (int s = 0; s < E; ++s) {
      row[s] = (int)row64[s];
      col[s] = (int)col64[s];
    }
  } else {   // stable counting sort by row
    perm.resize(E);
    std::vector<int> cur(rowptr.begin(), rowptr.end() - 1);
    for (int s = 0; s < E; ++s) {
      int pos = cur[row64[s]]++;
      perm[pos] = s;
      row[pos] = (int)row64[s];
      col[pos] = (int)col64[s];
    }
  }
  const int nG = (E + GROUP - 1) / GROUP;
  std::vector<int> gfirst(nG), gpair((size_t)nG + 1);
  int np = 0, max_seg = 0;
  for (int gI = 0; gI < nG; ++gI) {
    int s0 = gI * GROUP, s1 = std::min(E, s0 + GROUP) - 1;
    gfirst[gI] = row[s0];
    gpair[gI] = np;
    np += row[s1] - row[s0] + 1;
    int segs = 1;
    for (int q = s0 + 1; q <= s1; ++q) segs += row[q] != row[q - 1];
    max_seg = std::max(max_seg, segs);
  }
  ctx->max_seg = max_seg;
  gpair[nG] = np;

  ENS(ctx, ctx->d_row, (size_t)E * 4);
  ENS(ctx, ctx->d_col, (size_t)E * 4);
  ENS(ctx, ctx->d_rowptr, ((size_t)V + 1) * 4);
  ENS(ctx, ctx->d_grp_first, (size_t)nG * 4);
  ENS(ctx, ctx->d_grp_pair, ((size_t)nG + 1) * 4);
  CK(ctx, cudaMemcpyAsync(ctx->d_row.p, row.data(), (size_t)E * 4, cudaMemcpyHostToDevice, st));
  CK(ctx, cudaMemcpyAsync(ctx->d_col.p, col.data(), (size_t)E * 4, cudaMemcpyHostToDevice, st));
  CK(ctx, cudaMemcpyAsync(ctx->d_rowptr.p, rowptr.data(), ((size_t)V + 1) * 4, cudaMemcpyHostToDevice, st));
  CK(ctx, cudaMemcpyAsync(ctx->d_grp_first.p, gfirst.data(), (size_t)nG * 4, cudaMemcpyHostToDevice, st));
  CK(ctx, cudaMemcpyAsync(ctx->d_grp_pair.p, gpair.data(), ((size_t)nG + 1) * 4, cudaMemcpyHostToDevice, st));
  if (!sorted) {
    ENS(ctx, ctx->d_perm, (size_t)E * 4);
    CK(ctx, cudaMemcpyAsync(ctx->d_perm.p, perm.data(), (size_t)E * 4, cudaMemcpyHostToDevice, st));
  }
  CK(ctx, cudaStreamSynchronize(st));   // host vectors go out of scope
  GraphDev& g = ctx->g;
  g.V = V; g.E = E;
  g.row = (const int*)ctx->d_row.p; g.col = (const int*)ctx->d_col.p;
  g.perm = sorted ? nullptr : (const int*)ctx->d_perm.p;
  g.rowptr = (const int*)ctx->d_rowptr.p;
  g.n_groups = nG; g.grp_first = (const int*)ctx->d_grp_first.p; g.grp_pair = (const int*)ctx->d_grp_pair.p;
  g.n_pairs = np;
  ctx->gn_segments = gn_segments;

  // workspace
  const size_t Epad = (size_t)((E + 127) / 128) * 128;   // whole 128-row tiles for the tensor-core kernel
  ENS(ctx, ctx->e, Epad * H * sizeof(float));
  // the padding rows of the last tile are carried through every layer like real rows (and discarded): they must start
  // finite, or a stale NaN would travel with them
  if (Epad > (size_t)E) CK(ctx, cudaMemsetAsync((float*)ctx->e.p + (size_t)E * H, 0, (Epad - E) * H * sizeof(float), st));
  ENS(ctx, ctx->h, (size_t)V * H * sizeof(float));
  ENS(ctx, ctx->h0, (size_t)V * H * sizeof(float));
  ENS(ctx, ctx->uvab, (size_t)V * 4 * H * sizeof(float));
  ENS(ctx, ctx->uvab0, (size_t)V * 4 * H * sizeof(float));
  ENS(ctx, ctx->partials, (size_t)np * H * sizeof(float));
  const size_t feat_rows = 65536;
  ENS(ctx, ctx->feat, feat_rows * H * sizeof(float));
  {
    int R = ctx->node_only ? V : E;
    int rps = R / gn_segments;
    int bps = (rps + GN_ROWS_PER_BLOCK - 1) / GN_ROWS_PER_BLOCK;
    ENS(ctx, ctx->gn_part, std::max((size_t)gn_segments * bps, (size_t)1024) * 32 * 2 * sizeof(double));
    ENS(ctx, ctx->gn_stats, (size_t)gn_segments * 32 * 2 * sizeof(float));
  }
  ENS(ctx, ctx->d_xt, (size_t)std::max(V, E) * sizeof(float));
  ctx->graph_ready = true;
  ctx->buf_gen++;
  ctx->points_ready = false;
  return DFB_OK;
}

// node-side linears h -> [U|V|A|B] h of layer l: tensor-core path (product) or fp32 FFMA (validation impl)
static int node_linears(dfb_ctx* ctx, int l, const float* h, float* uvab, int V, cudaStream_t st) {
  if (ctx->edge_impl == DFB_EDGE_IMPL_FP32) {
    dim3 grid((V + LIN_ROWS - 1) / LIN_ROWS, 4);
    k_linear<<<grid, 256, 0, st>>>(h, ctx->layers[l].Wt_uvab, ctx->layers[l].b_uvab, uvab, V, 4 * H);
    CKL(ctx);
    return DFB_OK;
  }
  int r = tc_launch_linear(&ctx->tc, l * 12 * H + 4 * H, 4, h, uvab, ctx->layers[l].b_uvab, V, ctx->g, ctx->layers[l], st);
  if (r) FAIL(ctx, DFB_E_CUDA, "tcgen05 node linear: %s", ctx->tc.err.c_str());
  ctx->launches += 1;
  return DFB_OK;
}

// rows X[R][256] -> Y = X Wt + b through the generic linear, chunked features
static int linear_rows(dfb_ctx* ctx, const float* X, const float* Wt, const float* b, float* Y, int R, int N,
                       cudaStream_t st) {
  dim3 grid((R + LIN_ROWS - 1) / LIN_ROWS, N / 256);
  k_linear<<<grid, 256, 0, st>>>(X, Wt, b, Y, R, N);
  CKL(ctx);
  return DFB_OK;
}

// embedding linears (256 -> 256): which = 0 edge_embed, 1 node_embed.  Tensor-core path unless the fp32 validation
// implementation is selected.
static int embed_rows(dfb_ctx* ctx, int which, const float* X, float* Y, int R, cudaStream_t st) {
  if (ctx->edge_impl == DFB_EDGE_IMPL_FP32 || !ctx->graph_ready)
    return linear_rows(ctx, X, which ? ctx->Wt_node : ctx->Wt_edge, which ? ctx->b_node : ctx->b_edge, Y, R, H, st);
  int r = tc_launch_linear(&ctx->tc, (ctx->L * 12 + 2 * which) * H, 1, X, Y, which ? ctx->b_node : ctx->b_edge, R, ctx->g,
                           ctx->layers[0], st);
  if (r) FAIL(ctx, DFB_E_CUDA, "tcgen05 embedding linear: %s", ctx->tc.err.c_str());
  ctx->launches += 1;
  return DFB_OK;
}

extern "C" int dfb_set_points(dfb_ctx* ctx, const float* points, void* stream_) {
  if (!ctx) return DFB_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream_;
  CK(ctx, cudaSetDevice(ctx->device));
  if (!ctx->graph_ready) FAIL(ctx, DFB_E_INVALID, "dfb_prepare_graph must be called first");
  if (ctx->node_only) FAIL(ctx, DFB_E_INVALID, "dfb_set_points is for the TSP encoder (node_feature_only=0)");
  const int V = ctx->g.V;
  const float* dp = points;
  if (!is_device_ptr(points)) {
    ENS(ctx, ctx->d_points, (size_t)V * 2 * sizeof(float));
    CK(ctx, cudaMemcpyAsync(ctx->d_points.p, points, (size_t)V * 2 * sizeof(float), cudaMemcpyHostToDevice, st));
    dp = (const float*)ctx->d_points.p;
  }
  const int CH = 65536;
  for (int v0 = 0; v0 < V; v0 += CH) {
    int n = std::min(CH, V - v0);
    k_pos_features<<<n, H, 0, st>>>(dp + (size_t)v0 * 2, ctx->dimt128, (float*)ctx->feat.p, n);
    CKL(ctx);
    int r = embed_rows(ctx, 1, (const float*)ctx->feat.p, (float*)ctx->h0.p + (size_t)v0 * H, n, st);
    if (r) return r;
  }
  // layer 0's node linears are step-invariant too
  int r = node_linears(ctx, 0, (const float*)ctx->h0.p, (float*)ctx->uvab0.p, V, st);
  if (r) return r;
  ctx->points_ready = true;
  return DFB_OK;
}

// ================================================================================================
// one forward (+ optional fused posterior)
// ================================================================================================
// gn_blocks: when non-null and the pair kernel runs this layer, it also produces the head's GroupNorm partial sums
// (ctx->gn_part) and *gn_blocks receives the number of partial blocks; otherwise *gn_blocks stays 0.
static int launch_edge_layer(dfb_ctx* ctx, int l, const float* uvab, const float* tvec_edge, int write_e,
                             int e_zero, const float* xt_for_lut, cudaStream_t st, int* gn_blocks = nullptr) {
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (ctx->profiling) {
    if (ctx->ev_used + 2 > ctx->ev_pool.size()) {
      for (int i = 0; i < 256; ++i) {
        cudaEvent_t ev;
        CK(ctx, cudaEventCreate(&ev));
        ctx->ev_pool.push_back(ev);
      }
    }
    ev0 = ctx->ev_pool[ctx->ev_used++];
    ev1 = ctx->ev_pool[ctx->ev_used++];
    CK(ctx, cudaEventRecord(ev0, st));
  }
  if (ctx->edge_impl == DFB_EDGE_IMPL_FP32) {
    if (e_zero) CK(ctx, cudaMemsetAsync(ctx->e.p, 0, (size_t)ctx->g.E * H * sizeof(float), st));
    if (xt_for_lut) {
      k_lut_expand<<<(ctx->g.E + 3) / 4, 256, 0, st>>>(xt_for_lut, ctx->g.perm, ctx->lut, (float*)ctx->e.p, ctx->g.E);
      CKL(ctx);
    }
    k_edge_layer_fp32<<<ctx->g.n_groups, 256, EF_SMEM, st>>>((float*)ctx->e.p, uvab, (float*)ctx->partials.p,
                                                            ctx->g, ctx->layers[l], tvec_edge, write_e,
                                                            ctx->agg_mode);
    CKL(ctx);
  } else if (ctx->edge_impl == DFB_EDGE_IMPL_TC && ctx->pair_enabled && ctx->max_seg <= v2::MAXSEG && write_e &&
             (!(e_zero || xt_for_lut) || ctx->L > 1)) {
    // the CTA-pair kernel: middle layers read and write the edge stream; layer 0 (table rows in) runs in LUT mode
    const bool lut_mode = e_zero || xt_for_lut;
    const float* cl = lut_mode ? (e_zero ? ctx->cl0 + 2 * H : ctx->cl0) : nullptr;       // e0 = 0: zero tables
    const float* lut = lut_mode ? (e_zero ? ctx->cl0 + 2 * H : ctx->lut) : nullptr;
    // the last layer that writes e (TSP: L-1, read next by the head from row 0 up; MIS: L-2, read by the last layer's
    // kernel from tile 0 up) sweeps the edge stream downwards, the one before it upwards, and so on
    const int last_writer = ctx->node_only ? ctx->L - 2 : ctx->L - 1;
    const int sweep_down = (((last_writer - l) & 1) == 0 && l <= last_writer && ctx->serpentine) ? 1 : 0;
    int r = v2::launch(&ctx->pair, &ctx->tc, l, (float*)ctx->e.p, uvab, (float*)ctx->partials.p, ctx->g, ctx->layers[l],
                       tvec_edge, ctx->agg_mode, st, gn_blocks ? (double*)ctx->gn_part.p : nullptr, gn_blocks,
                       xt_for_lut, cl, lut, sweep_down);
    if (r) FAIL(ctx, DFB_E_CUDA, "tcgen05 pair edge layer: %s", ctx->tc.err.c_str());
    ctx->launches += ctx->tc.last_launches;
  } else {
    int r = tc_launch_edge_layer(&ctx->tc, l, (float*)ctx->e.p, uvab, (float*)ctx->partials.p, ctx->g,
                                 ctx->layers[l], tvec_edge, write_e, e_zero, xt_for_lut, ctx->lut,
                                 ctx->agg_mode, st);
    if (r) FAIL(ctx, r == -3 ? DFB_E_UNSUPPORTED : DFB_E_CUDA, "tcgen05 edge layer: %s", ctx->tc.err.c_str());
    ctx->launches += ctx->tc.last_launches;
  }
  if (ctx->profiling) CK(ctx, cudaEventRecord(ev1, st));
  return DFB_OK;
}

// tvec: [L][256] for this step.  binary_xt: xt in {0,1} guaranteed (categorical denoise state).
static int run_forward(dfb_ctx* ctx, const float* xt, const float* tvec, bool binary_xt, PosteriorArgs pa,
                       cudaStream_t st) {
  const GraphDev& g = ctx->g;
  const int V = g.V, E = g.E, L = ctx->L;
  float* h = (float*)ctx->h.p;
  float* e = (float*)ctx->e.p;
  float* uvab = (float*)ctx->uvab.p;
  const float* xt_lut = nullptr;
  int e_zero = 0;
  if (!ctx->node_only) {
    if (!ctx->points_ready) FAIL(ctx, DFB_E_INVALID, "dfb_set_points must be called before a TSP forward");
    CK(ctx, cudaMemcpyAsync(h, ctx->h0.p, (size_t)V * H * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (binary_xt) {
      xt_lut = xt;   // layer 0 reads the 2-row LUT instead of a materialised e0
    } else {         // general values (Gaussian diffusion): e0 = edge_embed(edge_pos_embed(xt))
      const int CH = 65536;
      for (int s0 = 0; s0 < E; s0 += CH) {
        int n = std::min(CH, E - s0);
        // with a permutation the index array addresses the whole xt; without, offset the input
        k_scalar_features<<<n, H, 0, st>>>(g.perm ? xt : xt + s0, g.perm ? g.perm + s0 : nullptr, ctx->dimt256,
                                           (float*)ctx->feat.p, n);
        CKL(ctx);
        int r = embed_rows(ctx, 0, (const float*)ctx->feat.p, e + (size_t)s0 * H, n, st);
        if (r) return r;
      }
    }
  } else {
    const int CH = 65536;
    for (int v0 = 0; v0 < V; v0 += CH) {
      int n = std::min(CH, V - v0);
      k_scalar_features<<<n, H, 0, st>>>(xt + v0, nullptr, ctx->dimt256, (float*)ctx->feat.p, n);
      CKL(ctx);
      int r = embed_rows(ctx, 1, (const float*)ctx->feat.p, h + (size_t)v0 * H, n, st);
      if (r) return r;
    }
    e_zero = 1;   // gnn_encoder.py:407: e0 = zeros
  }
  int gn_fused_blocks = 0;
  for (int l = 0; l < L; ++l) {
    const float* uv = uvab;
    if (l == 0 && !ctx->node_only) {
      uv = (const float*)ctx->uvab0.p;
    } else {
      int r = node_linears(ctx, l, h, uvab, V, st);
      if (r) return r;
    }
    const float* tv = tvec + (size_t)l * H;
    int write_e = !(ctx->node_only && l == L - 1);
    // the last layer of the sparse TSP encoder also accumulates the head's GroupNorm statistics (one segment only)
    const bool want_gn = !ctx->node_only && l == L - 1 && ctx->gn_segments == 1;
    int r = launch_edge_layer(ctx, l, uv, ctx->node_only ? nullptr : tv, write_e, (l == 0) ? e_zero : 0,
                              (l == 0) ? xt_lut : nullptr, st, want_gn ? &gn_fused_blocks : nullptr);
    if (r) return r;
    if (ctx->node_only || l < L - 1) {   // TSP never reads h after the last layer (gnn_encoder.py:400)
      k_node_update<<<(V + 7) / 8, 256, 0, st>>>(h, uv, (const float*)ctx->partials.p, g, ctx->layers[l].ln_h_g,
                                                 ctx->layers[l].ln_h_b, ctx->node_only ? tv : nullptr, ctx->agg_mode);
      CKL(ctx);
    }
  }
  // head
  const float* Z = ctx->node_only ? h : e;
  const int R = ctx->node_only ? V : E;
  const int rps = R / ctx->gn_segments;
  const int bps = (rps + GN_ROWS_PER_BLOCK - 1) / GN_ROWS_PER_BLOCK;
  if (gn_fused_blocks > 0) {   // partial sums came out of the last edge layer's epilogue: no extra read of e
    k_gn_final<<<dim3(1, 32), 256, 0, st>>>((const double*)ctx->gn_part.p, gn_fused_blocks, rps, (float*)ctx->gn_stats.p);
    CKL(ctx);
  } else {
    k_gn_partial<<<dim3(bps, ctx->gn_segments), 256, 0, st>>>(Z, rps, (double*)ctx->gn_part.p);
    CKL(ctx);
    k_gn_final<<<dim3(ctx->gn_segments, 32), 256, 0, st>>>((const double*)ctx->gn_part.p, bps, rps, (float*)ctx->gn_stats.p);
    CKL(ctx);
  }
  k_head<<<(R + 255) / 256, 256, 0, st>>>(Z, R, rps, (const float*)ctx->gn_stats.p, ctx->node_only ? nullptr : g.perm,
                                      ctx->hp, pa);
  CKL(ctx);
  return DFB_OK;
}

// Pinned staging slot for this call: waits (host side) only if the copies of the call that used the slot two calls
// ago have not executed yet, i.e. the host never runs more than one call ahead of the device.
static int stage_acquire(dfb_ctx* ctx, int* slot) {
  *slot = ctx->stage_next;
  ctx->stage_next = (ctx->stage_next + 1) % dfb_ctx::STAGE_SLOTS;
  CK(ctx, cudaEventSynchronize(ctx->stage_ev[*slot]));
  return DFB_OK;
}

// all time-MLP outputs of the S timesteps staged in h_tvals[slot] -> ctx->tvec [S][L][256]
static int compute_tvecs(dfb_ctx* ctx, int slot, int S, cudaStream_t st) {
  ENS(ctx, ctx->tvals, std::max<size_t>(4096, (size_t)S) * sizeof(float));
  ENS(ctx, ctx->tvec, (size_t)S * ctx->L * H * sizeof(float));
  CK(ctx, cudaMemcpyAsync(ctx->tvals.p, ctx->h_tvals[slot], (size_t)S * sizeof(float), cudaMemcpyHostToDevice, st));
  k_time_vectors<<<S, 256, 0, st>>>((const float*)ctx->tvals.p, ctx->tp, (const LayerParams*)ctx->layers_dev.p,
                                    ctx->L, (float*)ctx->tvec.p);
  CKL(ctx);
  return DFB_OK;
}

extern "C" int dfb_encoder_forward(dfb_ctx* ctx, const float* xt, float t, float* out, void* stream_) {
  if (!ctx) return DFB_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream_;
  CK(ctx, cudaSetDevice(ctx->device));
  if (!ctx->graph_ready) FAIL(ctx, DFB_E_INVALID, "dfb_prepare_graph must be called first");
  int slot;
  int r = stage_acquire(ctx, &slot);
  if (r) return r;
  ctx->h_tvals[slot][0] = t;
  r = compute_tvecs(ctx, slot, 1, st);
  if (r) return r;
  CK(ctx, cudaEventRecord(ctx->stage_ev[slot], st));
  PosteriorArgs pa{};
  pa.mode = HEAD_FORWARD;
  pa.net_out = out;
  return run_forward(ctx, xt, (const float*)ctx->tvec.p, false, pa, st);
}

static int step_args(dfb_ctx* ctx, int diffusion_type, const float* consts, int last, PosteriorArgs* pa) {
  if (diffusion_type == DFB_DIFFUSION_CATEGORICAL) {
    if (ctx->out_channels != 2) FAIL(ctx, DFB_E_INVALID, "categorical diffusion needs out_channels == 2");
    pa->mode = HEAD_CATEGORICAL;
  } else if (diffusion_type == DFB_DIFFUSION_GAUSSIAN) {
    if (ctx->out_channels != 1) FAIL(ctx, DFB_E_INVALID, "gaussian diffusion needs out_channels == 1");
    pa->mode = HEAD_GAUSSIAN;
  } else {
    FAIL(ctx, DFB_E_INVALID, "Unknown diffusion type %d", diffusion_type);
  }
  if (consts)
    for (int i = 0; i < 4; ++i) pa->c[i] = consts[i];
  pa->last = last;
  return DFB_OK;
}

extern "C" int dfb_denoise_step(dfb_ctx* ctx, int diffusion_type, const float* xt_in, float t,
                                const float* consts, int last, const float* uniforms, uint64_t seed,
                                int step_index, float* xt_out, float* p_out, float* net_out, void* stream_) {
  if (!ctx) return DFB_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream_;
  CK(ctx, cudaSetDevice(ctx->device));
  if (!ctx->graph_ready) FAIL(ctx, DFB_E_INVALID, "dfb_prepare_graph must be called first");
  PosteriorArgs pa{};
  int r = step_args(ctx, diffusion_type, consts, last, &pa);
  if (r) return r;
  int slot;
  r = stage_acquire(ctx, &slot);
  if (r) return r;
  ctx->h_tvals[slot][0] = t;
  r = compute_tvecs(ctx, slot, 1, st);
  if (r) return r;
  CK(ctx, cudaEventRecord(ctx->stage_ev[slot], st));
  pa.xt_in = xt_in; pa.uniforms = uniforms; pa.seed = seed; pa.step = (unsigned)step_index;
  pa.xt_out = xt_out; pa.p_out = p_out; pa.net_out = net_out;
  return run_forward(ctx, xt_in, (const float*)ctx->tvec.p, diffusion_type == DFB_DIFFUSION_CATEGORICAL, pa, st);
}

// the `steps` forwards + posteriors of the loop, every per-step quantity read from device tables
static int enqueue_loop(dfb_ctx* ctx, int diffusion_type, float* xt, int steps, const float* uniforms, cudaStream_t st) {
  const size_t N = ctx->node_only ? ctx->g.V : ctx->g.E;
  for (int i = 0; i < steps; ++i) {
    PosteriorArgs pa{};
    int r = step_args(ctx, diffusion_type, nullptr, 0, &pa);
    if (r) return r;
    pa.sp = (const StepParams*)ctx->d_steps.p + i;
    pa.xt_in = xt; pa.xt_out = xt;
    pa.uniforms = uniforms ? uniforms + (size_t)i * N : nullptr;
    r = run_forward(ctx, xt, (const float*)ctx->tvec.p + (size_t)i * ctx->L * H,
                    diffusion_type == DFB_DIFFUSION_CATEGORICAL, pa, st);
    if (r) return r;
  }
  return DFB_OK;
}

extern "C" int dfb_denoise(dfb_ctx* ctx, int diffusion_type, float* xt, int steps, const int32_t* t1,
                           const float* consts, const int32_t* last_flags, const float* uniforms,
                           uint64_t seed, void* stream_) {
  if (!ctx) return DFB_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream_;
  CK(ctx, cudaSetDevice(ctx->device));
  if (!ctx->graph_ready) FAIL(ctx, DFB_E_INVALID, "dfb_prepare_graph must be called first");
  if (steps < 1 || steps > dfb_ctx::MAX_STEPS) FAIL(ctx, DFB_E_INVALID, "steps %d out of range", steps);
  if (!ctx->node_only && !ctx->points_ready) FAIL(ctx, DFB_E_INVALID, "dfb_set_points must be called before a TSP forward");
  {
    PosteriorArgs chk{};
    int r = step_args(ctx, diffusion_type, nullptr, 0, &chk);
    if (r) return r;
  }
  int slot;
  int r = stage_acquire(ctx, &slot);
  if (r) return r;
  for (int i = 0; i < steps; ++i) {
    ctx->h_tvals[slot][i] = (float)t1[i];
    StepParams& sp = ctx->h_steps[slot][i];
    for (int k = 0; k < 4; ++k) sp.c[k] = consts[4 * i + k];
    sp.last = last_flags[i];
    sp.step = (unsigned)i;
    sp.seed = seed;
  }
  ENS(ctx, ctx->d_steps, (size_t)dfb_ctx::MAX_STEPS * sizeof(StepParams));
  r = compute_tvecs(ctx, slot, steps, st);
  if (r) return r;
  CK(ctx, cudaMemcpyAsync(ctx->d_steps.p, ctx->h_steps[slot], (size_t)steps * sizeof(StepParams), cudaMemcpyHostToDevice, st));
  CK(ctx, cudaEventRecord(ctx->stage_ev[slot], st));
  // the loop state lives in the context's own buffer, so the captured graph does not depend on the caller's pointer
  const size_t N = ctx->node_only ? ctx->g.V : ctx->g.E;
  float* x = (float*)ctx->d_xt.p;
  if (xt != x) CK(ctx, cudaMemcpyAsync(x, xt, N * sizeof(float), cudaMemcpyDeviceToDevice, st));

  const bool want_graph = ctx->capture_enabled && !ctx->capture_broken && !ctx->profiling;
  if (want_graph) {
    dfb_ctx::LoopKey key;
    key.buf_gen = ctx->buf_gen; key.steps = steps; key.diffusion = diffusion_type; key.impl = ctx->edge_impl;
    key.agg = ctx->agg_mode; key.pair = ctx->pair_enabled; key.uniforms = uniforms;
    if (!ctx->loop_exec || !(ctx->loop_key == key)) {
      if (ctx->loop_exec) {
        cudaGraphExecDestroy(ctx->loop_exec);
        ctx->loop_exec = nullptr;
      }
      const int64_t l0 = ctx->launches;
      cudaGraph_t graph = nullptr;
      cudaError_t ce = cudaStreamBeginCapture(ctx->loop_stream, cudaStreamCaptureModeThreadLocal);
      if (ce == cudaSuccess) {
        r = enqueue_loop(ctx, diffusion_type, x, steps, uniforms, ctx->loop_stream);
        ce = cudaStreamEndCapture(ctx->loop_stream, &graph);
        if (r == DFB_OK && ce == cudaSuccess) ce = cudaGraphInstantiate(&ctx->loop_exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
      }
      ctx->loop_launches = ctx->launches - l0;
      ctx->launches = l0;
      if (r != DFB_OK || ce != cudaSuccess || !ctx->loop_exec) {
        // capture is an optimisation: fall back to plain launches (and stop trying) rather than fail the call
        cudaGetLastError();
        ctx->loop_exec = nullptr;
        ctx->capture_broken = true;
        if (r != DFB_OK) return r;
      } else {
        ctx->loop_key = key;
      }
    }
  }
  if (want_graph && ctx->loop_exec) {
    CK(ctx, cudaEventRecord(ctx->loop_in, st));
    CK(ctx, cudaStreamWaitEvent(ctx->loop_stream, ctx->loop_in, 0));
    CK(ctx, cudaGraphLaunch(ctx->loop_exec, ctx->loop_stream));
    CK(ctx, cudaEventRecord(ctx->loop_out, ctx->loop_stream));
    CK(ctx, cudaStreamWaitEvent(st, ctx->loop_out, 0));
    ctx->launches += ctx->loop_launches;
  } else {
    r = enqueue_loop(ctx, diffusion_type, x, steps, uniforms, st);
    if (r) return r;
  }
  if (xt != x) CK(ctx, cudaMemcpyAsync(xt, x, N * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return DFB_OK;
}

extern "C" int dfb_set_graph_capture(dfb_ctx* ctx, int enabled) {
  if (!ctx) return DFB_E_INVALID;
  ctx->capture_enabled = enabled != 0;
  ctx->capture_broken = false;
  return DFB_OK;
}

extern "C" int dfb_denoise_host(dfb_ctx* ctx, int diffusion_type, const float* points,
                                const int64_t* edge_index, int64_t V, int64_t E, int gn_segments,
                                const float* xt0, int steps, const int32_t* t1, const float* consts,
                                const int32_t* last_flags, uint64_t seed, float* heatmap_out, void* stream_) {
  if (!ctx) return DFB_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream_;
  CK(ctx, cudaSetDevice(ctx->device));
  int r = dfb_prepare_graph(ctx, edge_index, V, E, gn_segments, st);
  if (r) return r;
  if (!ctx->node_only) {
    if (!points) FAIL(ctx, DFB_E_INVALID, "points required for TSP");
    r = dfb_set_points(ctx, points, st);
    if (r) return r;
  }
  const size_t N = ctx->node_only ? (size_t)V : (size_t)E;
  CK(ctx, cudaMemcpyAsync(ctx->d_xt.p, xt0, N * sizeof(float), cudaMemcpyHostToDevice, st));
  r = dfb_denoise(ctx, diffusion_type, (float*)ctx->d_xt.p, steps, t1, consts, last_flags, nullptr, seed, st);
  if (r) return r;
  CK(ctx, cudaMemcpyAsync(heatmap_out, ctx->d_xt.p, N * sizeof(float), cudaMemcpyDeviceToHost, st));
  CK(ctx, cudaStreamSynchronize(st));
  return DFB_OK;
}

// ================================================================================================
extern "C" int dfb_profile_begin(dfb_ctx* ctx) {
  if (!ctx) return DFB_E_INVALID;
  ctx->profiling = true;
  ctx->ev_used = 0;
  return DFB_OK;
}
extern "C" int dfb_profile_end(dfb_ctx* ctx, double* ms, int64_t* n) {
  if (!ctx) return DFB_E_INVALID;
  CK(ctx, cudaSetDevice(ctx->device));
  CK(ctx, cudaDeviceSynchronize());
  double tot = 0.0;
  for (size_t i = 0; i + 1 < ctx->ev_used; i += 2) {
    float t = 0.f;
    CK(ctx, cudaEventElapsedTime(&t, ctx->ev_pool[i], ctx->ev_pool[i + 1]));
    tot += t;
  }
  if (ms) *ms = tot;
  if (n) *n = (int64_t)(ctx->ev_used / 2);
  ctx->profiling = false;
  ctx->ev_used = 0;
  return DFB_OK;
}

// ================================================================================================
// Test hook: run only GEMM1 of layer `layer` (acc = e_in * C^T, split-bf16 on tensor cores) on the
// prepared graph's tiling and dump the TMEM accumulator.  Isolates descriptors / TMA / TMEM
// plumbing from the epilogue math in tests/test_tc_gemm.py.
extern "C" int dfb_debug_edge_gemm(dfb_ctx* ctx, int layer, const float* e_in, float* acc_out, void* stream_) {
  if (!ctx) return DFB_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream_;
  CK(ctx, cudaSetDevice(ctx->device));
  if (!ctx->graph_ready) FAIL(ctx, DFB_E_INVALID, "dfb_prepare_graph must be called first");
  if (layer < 0 || layer >= ctx->L) FAIL(ctx, DFB_E_INVALID, "layer out of range");
  ctx->tc.debug_acc = acc_out;
  int r;
  if (ctx->edge_impl == DFB_EDGE_IMPL_TC && ctx->pair_enabled && ctx->max_seg <= v2::MAXSEG)
    r = v2::launch(&ctx->pair, &ctx->tc, layer, const_cast<float*>(e_in), (const float*)ctx->uvab.p,
                   (float*)ctx->partials.p, ctx->g, ctx->layers[layer], nullptr, AGG_SUM, st);
  else
    r = tc_launch_edge_layer(&ctx->tc, layer, const_cast<float*>(e_in), (const float*)ctx->uvab.p,
                             (float*)ctx->partials.p, ctx->g, ctx->layers[layer], nullptr, 0, 0, nullptr,
                             ctx->lut, AGG_SUM, st);
  ctx->tc.debug_acc = nullptr;
  if (r) FAIL(ctx, DFB_E_CUDA, "tcgen05 edge layer: %s", ctx->tc.err.c_str());
  ctx->launches += 1;
  return DFB_OK;
}

// Test/tuning hook: read and reset the per-phase cycle counters of the tcgen05 edge kernel
// (filled only when DFB_TC_PROBE has bit 7 set, --prof build).  out[32] host.
extern "C" int dfb_debug_phase_cycles(dfb_ctx* ctx, unsigned long long* out) {
  if (!ctx || !out) return DFB_E_INVALID;
  CK(ctx, cudaSetDevice(ctx->device));
  CK(ctx, cudaDeviceSynchronize());
  CK(ctx, cudaMemcpy(out, ctx->tc.phase_cycles, 32 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  CK(ctx, cudaMemset(ctx->tc.phase_cycles, 0, 32 * sizeof(unsigned long long)));
  return DFB_OK;
}

#ifdef DFB_PHASE_PROF
// Tuning build only (not part of the ABI): time line of cluster 0's leader CTA, see g_pair_trace in edge_layer_v2.cuh.  out[512].
extern "C" int dfb_debug_pair_trace(long long* out) {
  if (!out) return DFB_E_INVALID;
  if (cudaDeviceSynchronize() != cudaSuccess) return DFB_E_CUDA;
  if (cudaMemcpyFromSymbol(out, dfb::v2::g_pair_trace, sizeof(long long) * 512) != cudaSuccess) return DFB_E_CUDA;
  return DFB_OK;
}
#endif

// Diagnostic: watchdog record of the tcgen05 kernel (host-mapped, readable even after a launch failure):
// out[0] = wait-site code (0 = none), out[1] = blockIdx.x, out[2] = parity waited for, out[3] = threadIdx.x.
extern "C" int dfb_debug_watchdog(dfb_ctx* ctx, int* out) {
  if (!ctx || !out || !ctx->tc.error_host) return DFB_E_INVALID;
  for (int i = 0; i < 4; ++i) out[i] = ((volatile int*)ctx->tc.error_host)[i];
  return DFB_OK;
}

// ================================================================================================
// Row f1 (the step before the path): sparse k-NN graph of one TSP instance.
// Replaces TSPGraphDataset.__getitem__'s KDTree query + edge_index assembly (co_datasets/tsp_graph_dataset.py:52-62).
//   points      (N,2) float64, HOST or DEVICE (the reference parses coordinates to float64 and queries in float64)
//   edge_index  (2, N*K) int64, DEVICE: row = arange(N).repeat_interleave(K) (+ node_offset), col = neighbours in
//               ascending distance (self first) (+ node_offset: block-diagonal batching, pl_meta_model.py:177-184)
extern "C" int dfb_knn_graph(dfb_ctx* ctx, const double* points, int64_t num_nodes, int k, int64_t node_offset,
                             int64_t* edge_index, void* stream_) {
  if (!ctx) return DFB_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream_;
  CK(ctx, cudaSetDevice(ctx->device));
  if (num_nodes < 1 || k < 1 || k > num_nodes) FAIL(ctx, DFB_E_INVALID, "bad kNN size N=%lld K=%d", (long long)num_nodes, k);
  const size_t smem = (size_t)num_nodes * sizeof(double);
  if (smem > 200 * 1024) FAIL(ctx, DFB_E_UNSUPPORTED, "kNN graph: N=%lld exceeds the shared-memory brute-force limit (25600)", (long long)num_nodes);
  if (!is_device_ptr(edge_index)) FAIL(ctx, DFB_E_INVALID, "edge_index must be a device pointer");
  const double* dp = points;
  if (!is_device_ptr(points)) {
    ENS(ctx, ctx->d_points, (size_t)num_nodes * 2 * sizeof(double));
    CK(ctx, cudaMemcpyAsync(ctx->d_points.p, points, (size_t)num_nodes * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
    dp = (const double*)ctx->d_points.p;
  }
  CK(ctx, cudaFuncSetAttribute(k_knn_bruteforce, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_knn_bruteforce<<<(int)num_nodes, 256, smem, st>>>(dp, (int)num_nodes, k, (long long*)edge_index, (long long)node_offset);
  CKL(ctx);
  return DFB_OK;
}

// ================================================================================================
// Rows f2 / f3 (the steps after the path): tour merge on the host, 2-opt on the GPU.  See tsp_decode.cuh.
extern "C" int dfb_tsp_merge_sparse(const double* points, int64_t n, const float* heat, const int64_t* edge_index, int64_t E,
                                    int mode, int64_t* tour, int64_t* merge_iterations) {
  if (!points || !heat || !edge_index || !tour || !merge_iterations || n < 3 || n > 0x7fffffff / 2 || E < 0 || (mode != 0 && mode != 1))
    return DFB_E_INVALID;
  int r = tspmerge::merge_sparse(points, (int)n, heat, edge_index, E, mode, tour, merge_iterations);
  return r < 0 ? DFB_E_INVALID : r;
}

extern "C" int dfb_tsp_merge_order(int64_t n, const int64_t* order, int64_t count, int64_t* tour, int64_t* merge_iterations) {
  if (!order || !tour || !merge_iterations || n < 3 || n > 0x7fffffff / 2 || count < 0) return DFB_E_INVALID;
  int r = tspmerge::merge_order((int)n, order, count, tour, merge_iterations);
  return r < 0 ? DFB_E_INVALID : r;
}

extern "C" int dfb_two_opt(dfb_ctx* ctx, const double* points, int64_t n, int64_t* tours, int64_t batch, int64_t max_iterations,
                           int64_t* iterations_out, void* stream_) {
  if (!ctx) return DFB_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream_;
  CK(ctx, cudaSetDevice(ctx->device));
  if (!points || !tours || !iterations_out) FAIL(ctx, DFB_E_INVALID, "two_opt: null argument");
  if (n < 3 || n > 46340 || batch < 1 || batch > 65535) FAIL(ctx, DFB_E_INVALID, "two_opt: bad size n=%lld batch=%lld (n in [3, 46340], batch in [1, 65535])", (long long)n, (long long)batch);
  const int N = (int)n, B = (int)batch;
  for (int64_t k = 0; k < batch * (n + 1); ++k)
    if (tours[k] < 0 || tours[k] >= n) FAIL(ctx, DFB_E_INVALID, "two_opt: tour entry %lld out of range", (long long)tours[k]);
  const int T = (N + TWOOPT_TILE - 1) / TWOOPT_TILE;
  std::vector<int2> tiles;
  for (int a = 0; a < T; ++a)
    for (int b = a; b < T; ++b) tiles.push_back(make_int2(a, b));
  const int ntiles = (int)tiles.size();
  ENS(ctx, ctx->opt_points, (size_t)N * 2 * sizeof(double));
  ENS(ctx, ctx->opt_tours, (size_t)B * (N + 1) * sizeof(long long));
  ENS(ctx, ctx->opt_pos, (size_t)B * (N + 1) * 2 * sizeof(double));
  ENS(ctx, ctx->opt_dnext, (size_t)B * N * sizeof(double));
  ENS(ctx, ctx->opt_cand, (size_t)B * ntiles * sizeof(TwoOptCand));
  ENS(ctx, ctx->opt_tiles, (size_t)ntiles * sizeof(int2));
  ENS(ctx, ctx->opt_state, sizeof(TwoOptState));
  ENS(ctx, ctx->opt_best, (size_t)B * sizeof(TwoOptCand));
  double* d_points = (double*)ctx->opt_points.p;
  long long* d_tours = (long long*)ctx->opt_tours.p;
  double* d_pos = (double*)ctx->opt_pos.p;
  double* d_dnext = (double*)ctx->opt_dnext.p;
  TwoOptCand* d_cand = (TwoOptCand*)ctx->opt_cand.p;
  int2* d_tiles = (int2*)ctx->opt_tiles.p;
  TwoOptState* d_state = (TwoOptState*)ctx->opt_state.p;
  CK(ctx, cudaMemcpyAsync(d_points, points, (size_t)N * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
  CK(ctx, cudaMemcpyAsync(d_tours, tours, (size_t)B * (N + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
  CK(ctx, cudaMemcpyAsync(d_tiles, tiles.data(), (size_t)ntiles * sizeof(int2), cudaMemcpyHostToDevice, st));
  CK(ctx, cudaMemsetAsync(d_state, 0, sizeof(TwoOptState), st));
  k_twoopt_init<<<dim3((N + 256) / 256, B), 256, 0, st>>>(d_points, d_tours, d_pos, d_dnext, N);
  CKL(ctx);
  TwoOptState hs{};
  int chunk = 8;
  while (true) {
    for (int c = 0; c < chunk; ++c) {
      k_twoopt_eval<<<dim3(ntiles, B), 256, 0, st>>>(d_pos, d_dnext, d_tiles, d_cand, d_state, N, ntiles);
      CKL(ctx);
      k_twoopt_apply<<<1, 1024, 0, st>>>(d_tours, d_pos, d_dnext, d_cand, d_state, (TwoOptCand*)ctx->opt_best.p, N, B, ntiles,
                                         (long long)max_iterations);
      CKL(ctx);
    }
    CK(ctx, cudaMemcpyAsync(&hs, d_state, sizeof(hs), cudaMemcpyDeviceToHost, st));
    CK(ctx, cudaStreamSynchronize(st));
    if (hs.done) break;
    if (chunk < 64) chunk *= 2;
  }
  CK(ctx, cudaMemcpyAsync(tours, d_tours, (size_t)B * (N + 1) * sizeof(long long), cudaMemcpyDeviceToHost, st));
  CK(ctx, cudaStreamSynchronize(st));
  *iterations_out = hs.iterations;
  return DFB_OK;
}

// Row f4: text heat map for tsp_mcts (convert_numpy_to_txt.py:57-73).  Most entries are exactly zero after the
// sparsification, so those are copied as a literal; the rest go through printf's correctly rounded "%.6f" (what
// Python's f"{x:.6f}" produces as well).
extern "C" int dfb_write_heatmap_txt(const char* path, int64_t n, const double* matrix) {
  if (!path || !matrix || n < 1) return DFB_E_INVALID;
  FILE* f = fopen(path, "wb");
  if (!f) return DFB_E_INVALID;
  std::vector<char> line((size_t)n * 28 + 2);
  bool ok = fprintf(f, "%lld\n", (long long)n) > 0;
  for (int64_t r = 0; r < n && ok; ++r) {
    char* w = line.data();
    const double* row = matrix + r * n;
    for (int64_t c = 0; c < n; ++c) {
      if (c) *w++ = ' ';
      double x = row[c];
      if (x == 0.0 && !std::signbit(x)) {
        memcpy(w, "0.000000", 8);
        w += 8;
      } else {
        const int k = snprintf(w, 27, "%.6f", x);   // snprintf returns the UNtruncated length
        if (k < 0 || k > 26) {                      // |x| >= ~1e19 or non-finite garbage: not a heat map
          fclose(f);
          return DFB_E_INVALID;
        }
        w += k;
      }
    }
    *w++ = '\n';
    ok = fwrite(line.data(), 1, (size_t)(w - line.data()), f) == (size_t)(w - line.data());
  }
  ok = (fclose(f) == 0) && ok;
  return ok ? DFB_OK : DFB_E_INVALID;
}
