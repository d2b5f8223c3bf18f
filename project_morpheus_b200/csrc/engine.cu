// libsnacb engine: weight packing, the plan-driven layer pipeline and the C-ABI entry points
// declared in include/snacb.h.  Replaces the reference's convert_to_audio / model.decode path
// (Morpheus_Client/tts_engine/speechpipe.py:64-137) for whole decode ticks.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "kernels.h"
#include "snacb.h"

using namespace snacb;

namespace {

thread_local std::string g_create_error;

struct RuDev {
  float *a1, *i1, *dw_w, *dw_b, *a2, *i2, *pw_w, *pw_b;
  __half* pw16;  // SNACB_PREC_FP16: [C][C]; SNACB_PREC_FP16X3: [C][3C] = [hi | lo | hi]
};
struct BlockDev {
  float *alpha, *inv, *ct_w /*[s*Cout][2*Cin]*/, *ct_b, *noise_w;
  float* ct_b2;   // W_n b: the NoiseBlock GEMM's share of the conv bias (composed ConvT + NoiseBlock kernel)
  __half *ct16, *noise16;
  __half* ctn16;  // blocks 0 / 1, fp16 recipe: stacked [conv rows | W_n conv rows] weight, [2 s Cout][2 Cin]; else null
  RuDev ru[3];
};
struct DevWeights {
  QuantW q;
  float *head_dw_w /*[7][768]*/, *head_dw_b, *head_pw_w, *head_pw_b;
  __half* head_pw16;
  BlockDev blk[4];
  float *tail_alpha, *tail_inv, *tail_w /*[7][64]*/, *tail_b;
};

struct EncBlockDev {
  RuDev ru[3];
  float *alpha, *inv, *down_w /*[2C][2s*C], k-major*/, *down_b;
};
struct EncDev {
  float *in_w, *in_b;
  EncBlockDev blk[4];
  float *out_dw_w /*[7][768]*/, *out_dw_b;
  float *inproj_w[3], *inproj_b[3], *cb_norm[3];
};
constexpr int kEncRates[4] = {2, 4, 8, 8};
constexpr int kEncDim = 48;
constexpr int kVqStrides[3] = {4, 2, 1};

enum KClass { KC_DEINT = 0, KC_CODES, KC_DW, KC_SNAKE, KC_GEMM1, KC_CONVT, KC_TAIL, KC_RU, KC_RUW, KC_BLK, KC_COUNT };
const char* const kClassName[KC_COUNT] = {"deinterleave", "from_codes", "dwconv_snake", "snake", "gemm_1x1",
                                          "gemm_convt", "tail_pack", "ru_fused",  // one kernel per ResidualUnit (k_ru_tc / k_ru_x / k_ru_p)
                                          "ru_fused_tmem",                      // k_ru_w: residual stream initialised in tensor memory
                                          "block_fused_tail"};                  // k_blk_tail: 3 ResidualUnits + decoder tail in one kernel

struct Prof {
  bool on = false;
  struct Rec { int cls; cudaEvent_t a, b; };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  double ms[KC_COUNT] = {}, flops[KC_COUNT] = {}, bytes[KC_COUNT] = {};
  int64_t launches[KC_COUNT] = {};
  cudaEvent_t get() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }
};

}  // namespace

struct snacb_engine {
  snacb_config cfg{};
  int device = 0;
  std::string err;
  bool loaded = false;
  DevWeights w{};
  void* warena = nullptr;
  size_t warena_bytes = 0;
  __half* harena = nullptr;  // fp16 copies of the GEMM-shaped weights (tensor-core recipe)
  EncDev enc{};
  void* earena = nullptr;    // encoder weights (snacb_load_encoder_weights)
  bool enc_loaded = false;
  char* ws = nullptr;
  size_t ws_bytes = 0;
  // pinned + device staging for the host-buffer API and the item tables
  char* pin = nullptr;
  size_t pin_bytes = 0;
  char* dstage = nullptr;
  size_t dstage_bytes = 0;
  // pipelined host ticks (snacb_decode_windows_host_submit / _wait): two slots, the device -> host copy of tick t runs on
  // copy_stream under the kernels of tick t+1
  struct PipeSlot {
    char* pin = nullptr; size_t pin_bytes = 0;
    char* dst = nullptr; size_t dst_bytes = 0;
    cudaEvent_t tail_done = nullptr, copied = nullptr;
    bool busy = false;
    int n_win = 0;
    int16_t* h_pcm = nullptr; int32_t* h_status = nullptr;  // caller's destinations
    bool direct = false;                                      // D2H went straight into the caller's pinned buffers
    size_t pcm_off = 0, st_off = 0;
  } pipe[2];
  int pipe_next = 0;
  cudaStream_t copy_stream = nullptr;
  Item* pin_items = nullptr;
  size_t pin_items_cap = 0;
  cudaEvent_t items_ev = nullptr;
  int64_t launches = 0;
  int prefetch_ahead = 0;  // SM count when L2 prefetch-ahead is on (SNACB_PREFETCH env, default on)
  bool ru256 = false;      // decoder block 1 (C = 256) ResidualUnits through the persistent fused kernel (SNACB_RU256 env)
  bool ruw = true;         // decoder block 1 ResidualUnits through k_ru_w (SNACB_RUW=0 falls back to k_dw_tc + k_gemm_ws)
  bool ruw128 = true;      // decoder block 2 ResidualUnits through k_ru_w<128> (SNACB_RUW128=0 falls back to k_ru_tc<128>)
  bool convt_n2 = true;    // blocks 1 / 3: ConvTranspose1d + NoiseBlock as ONE GEMM over a composed weight (SNACB_CONVT_N2=0: round-1 kernels)
  // CUDA graphs of small host-API ticks (latency mode): key -> instantiated graph
  struct GraphEntry { cudaGraphExec_t exec = nullptr; uint64_t gen = 0; int calls = 0; bool disabled = false; };
  std::map<std::vector<long long>, GraphEntry> graphs;
  cudaStream_t cap_stream = nullptr;
  uint64_t gen = 1;       // bumped whenever a buffer a captured graph points into is re-allocated
  int graph_max_win = 256;
  int64_t graph_launches = 0;
  // chunk lanes: chunks of one call run concurrently on side streams (L2-resident working sets, no wave tails)
  std::vector<cudaStream_t> lane_streams;
  std::vector<cudaEvent_t> lane_done;
  cudaEvent_t lane_fork = nullptr;
  Prof prof;
  // tap
  int tap_stage = -1;
  float* tap_buf = nullptr;
  size_t tap_cap = 0;
  int tap_rows = 0, tap_ch = 0, tap_lo = 0, tap_items = 0;
};

namespace {

int fail(snacb_engine* e, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (e) e->err = buf; else g_create_error = buf;
  return code;
}

#define CU(e, call)                                                                                   \
  do {                                                                                                \
    cudaError_t _c = (call);                                                                          \
    if (_c != cudaSuccess)                                                                            \
      return fail(e, SNACB_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_c), __FILE__, __LINE__); \
  } while (0)

int check_launch(snacb_engine* e, const char* what) {
  cudaError_t c = cudaGetLastError();
  if (c != cudaSuccess) return fail(e, SNACB_ECUDA, "%s: launch failed: %s", what, cudaGetErrorString(c));
  return SNACB_OK;
}

// Times one launch with a pair of events on the launching stream when profiling is on.
struct ProfScope {
  Prof& p; cudaStream_t st; cudaEvent_t a = nullptr; int cls;
  ProfScope(snacb_engine* e, int cls_, double flops, double bytes, cudaStream_t st_) : p(e->prof), st(st_), cls(cls_) {
    if (!p.on) return;
    p.flops[cls] += flops; p.bytes[cls] += bytes; p.launches[cls] += 1;
    a = p.get();
    cudaEventRecord(a, st);
  }
  ~ProfScope() {
    if (!a) return;
    cudaEvent_t b = p.get();
    cudaEventRecord(b, st);
    p.recs.push_back({cls, a, b});
  }
};

int ensure_ws(snacb_engine* e, size_t bytes, cudaStream_t st) {
  if (bytes <= e->ws_bytes) return SNACB_OK;
  ++e->gen;
  if (e->ws) {
    CU(e, cudaStreamSynchronize(st));
    CU(e, cudaDeviceSynchronize());
    CU(e, cudaFree(e->ws));
    e->ws = nullptr; e->ws_bytes = 0;
  }
  bytes = (bytes + (size_t(1) << 20)) & ~((size_t(1) << 20) - 1);
  cudaError_t c = cudaMalloc((void**)&e->ws, bytes);
  if (c != cudaSuccess) return fail(e, SNACB_ENOMEM, "workspace cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(c));
  e->ws_bytes = bytes;
  return SNACB_OK;
}

struct Bump {
  char* p; size_t off = 0;
  explicit Bump(char* base) : p(base) {}
  template <class T> T* take(size_t n) {
    off = (off + 255) & ~size_t(255);
    T* r = reinterpret_cast<T*>(p + off);
    off += n * sizeof(T);
    return r;
  }
};
inline size_t pad256(size_t b) { return (b + 255) & ~size_t(255); }

// ---------------------------------------------------------------------------------- weights
struct HostPack {
  std::vector<float> data;
  size_t add(const float* src, size_t n) {
    size_t off = (data.size() + 63) & ~size_t(63);
    data.resize(off + n);
    if (src) memcpy(data.data() + off, src, n * sizeof(float));
    return off;
  }
};

size_t add_inv(HostPack& hp, const float* alpha, int n) {
  size_t off = hp.add(nullptr, n);
  for (int i = 0; i < n; ++i) hp.data[off + i] = 1.0f / (alpha[i] + 1e-9f);  // (alpha + 1e-9).reciprocal()
  return off;
}
size_t add_transposed(HostPack& hp, const float* w, int C, int k) {  // [C][k] -> [k][C]
  size_t off = hp.add(nullptr, (size_t)C * k);
  for (int c = 0; c < C; ++c)
    for (int j = 0; j < k; ++j) hp.data[off + (size_t)j * C + c] = w[(size_t)c * k + j];
  return off;
}
// ConvTranspose1d weight [Cin][Cout][2s] -> polyphase GEMM operand [s*Cout][2*Cin] (see k_gemm_f32).
size_t add_convt(HostPack& hp, const float* w, int Cin, int Cout, int s) {
  const int p = (s + 1) / 2, k = 2 * s;
  size_t off = hp.add(nullptr, (size_t)s * Cout * 2 * Cin);
  for (int r = 0; r < s; ++r) {
    const int k_main = r + p;
    const int k_side = (r < s - p) ? r + p + s : r + p - s;
    for (int co = 0; co < Cout; ++co) {
      float* row = hp.data.data() + off + ((size_t)r * Cout + co) * 2 * Cin;
      for (int ci = 0; ci < Cin; ++ci) {
        row[ci] = w[((size_t)ci * Cout + co) * k + k_main];
        row[Cin + ci] = w[((size_t)ci * Cout + co) * k + k_side];
      }
    }
  }
  return off;
}

// ---------------------------------------------------------------------------------- pipeline
struct NoiseCfg {
  int mode; const float* tensor; long long stride; uint64_t seed; const unsigned long long* d_keys;
  const unsigned long long* d_seed;  // device-resident seed or nullptr
};

void tap(snacb_engine* e, int stage, const float* p, Rng r, int C, int n_items, bool first_chunk, cudaStream_t st) {
  if (e->tap_stage != stage || !first_chunk || !e->tap_buf) return;
  size_t n = (size_t)n_items * r.n() * C;
  if (n > e->tap_cap) n = e->tap_cap;
  cudaMemcpyAsync(e->tap_buf, p, n * sizeof(float), cudaMemcpyDeviceToDevice, st);
  e->tap_rows = r.n(); e->tap_ch = C; e->tap_lo = r.lo; e->tap_items = n_items;
}

// Runs one uniform group of items through the layer stack (fp32 recipe).
int run_group_f32(snacb_engine* e, const Plan& P, const Item* d_items, int n_total, int out_len, Rng tail_out,
                  const int32_t* c0, const int32_t* c1, const int32_t* c2, int pitch0, const NoiseCfg& nz,
                  int32_t* d_status, float* wav, int16_t* pcm, char* ws_free, size_t ws_avail,
                  cudaStream_t st) {
  const DevWeights& W = e->w;
  const size_t Zf = (size_t)P.z.n() * kLatent, Hf = (size_t)P.h.n() * kLatent, S = P.max_stage_floats;
  const size_t per_item = (pad256(Zf * 4) + pad256(Hf * 4) + 3 * pad256(S * 4)) + 1024;
  int chunk = e->cfg.chunk_items > 0 ? e->cfg.chunk_items : 32;
  if ((size_t)chunk * per_item > ws_avail) chunk = (int)(ws_avail / per_item);
  if (chunk < 1) return fail(e, SNACB_ENOMEM, "workspace too small for one item");
  const int F = P.T0 / 4;
  const int noise_off[4] = {0, 32 * F, 288 * F, 1312 * F};

  for (int start = 0; start < n_total; start += chunk) {
    const int n = std::min(chunk, n_total - start);
    const bool first = start == 0;
    Bump bp(ws_free);
    float* Z = bp.take<float>(Zf * n);
    float* H0 = bp.take<float>(Hf * n);
    float* X = bp.take<float>(S * n);
    float* Y = bp.take<float>(S * n);
    float* A = bp.take<float>(S * n);
    GroupCtx g{d_items ? d_items + start : nullptr, start, n, out_len, P.T0, st, &e->launches, e->cfg.flags};

    auto gemm = [&](const GemmArgs& a) {
      const double M = (double)n * a.m_r.n(), segs = a.epi == EPI_CONVT ? 2.0 : 1.0;
      const double flops = 2.0 * M * a.N * a.K * segs;
      const double by = 4.0 * ((double)n * a.a_r.n() * a.K + (double)a.N * a.K * segs + M * a.N * (a.R ? 2.0 : 1.0));
      ProfScope ps(e, a.epi == EPI_CONVT ? KC_CONVT : KC_GEMM1, flops, by, st);
      launch_gemm_f32(g, a);
    };
    auto dwconv = [&](const DwArgs& d) {
      const double el = (double)n * d.out_r.n() * d.C;
      ProfScope ps(e, KC_DW, el * (14.0 + (d.a1 ? 28.0 : 0.0) + (d.a2 ? 4.0 : 0.0)), 4.0 * ((double)n * d.in_r.n() * d.C + el), st);
      launch_dwconv(g, d);
    };
    {
      ProfScope ps(e, KC_CODES, 2.0 * 24 * kLatent * n * P.z.n(), 4.0 * kLatent * n * P.z.n(), st);
      launch_from_codes(g, W.q, c0, c1, c2, pitch0, P.z, Z);
    }
    tap(e, 0, Z, P.z, kLatent, n, first, st);
    {
      DwArgs d{Z, P.z, H0, P.h, kLatent, 1, 1, W.head_dw_w, W.head_dw_b, nullptr, nullptr, nullptr, nullptr};
      dwconv(d);
      tap(e, 1, H0, P.h, kLatent, n, first, st);
    }
    {
      GemmArgs a{};
      a.epi = EPI_BIAS; a.A = H0; a.lda = kLatent; a.a_r = P.h; a.W = W.head_pw_w; a.ldw = kLatent;
      a.bias = W.head_pw_b; a.K = kLatent; a.N = kDecDim; a.m_r = P.h; a.out = X; a.o_r = P.h; a.ldo = kDecDim;
      a.up = 1;
      gemm(a);
      tap(e, 2, X, P.h, kDecDim, n, first, st);
    }
    for (int b = 0; b < 4; ++b) {
      const BlockPlan& B = P.b[b];
      const BlockDev& Wb = W.blk[b];
      const int sid = 3 + 9 * b;
      {
        const double el = (double)n * B.in.n() * B.Cin;
        ProfScope ps(e, KC_SNAKE, 4.0 * el, 8.0 * el, st);
        launch_snake(g, X, A, B.in, B.Cin, Wb.alpha, Wb.inv);
      }
      tap(e, sid + 0, A, B.in, B.Cin, n, first, st);
      {
        GemmArgs a{};
        a.epi = EPI_CONVT; a.A = A; a.lda = B.Cin; a.a_r = B.in; a.W = Wb.ct_w; a.ldw = 2 * B.Cin;
        a.bias = Wb.ct_b; a.K = B.Cin; a.N = B.s * B.Cout; a.m_r = B.q; a.s = B.s; a.p = B.p; a.Cout = B.Cout;
        a.out = Y; a.o_r = B.ct; a.ldo = B.Cout; a.up = B.up_out;
        gemm(a);
        tap(e, sid + 1, Y, B.ct, B.Cout, n, first, st);
      }
      if (nz.mode != SNACB_NOISE_OFF) {
        GemmArgs a{};
        a.epi = EPI_NOISE; a.A = Y; a.lda = B.Cout; a.a_r = B.ct; a.W = Wb.noise_w; a.ldw = B.Cout;
        a.K = B.Cout; a.N = B.Cout; a.m_r = B.ct; a.out = X; a.o_r = B.ct; a.ldo = B.Cout;
        a.R = Y; a.r_r = B.ct; a.ldr = B.Cout; a.up = B.up_out;
        a.noise = NoiseSrc{nz.mode, nz.tensor, nz.stride, noise_off[b], (unsigned long long)nz.seed, nz.d_keys, b, nz.d_seed};
        gemm(a);
      } else {
        std::swap(X, Y);  // x + 0 * h == x
      }
      tap(e, sid + 2, X, B.ct, B.Cout, n, first, st);
      Rng cur = B.ct;
      for (int r = 0; r < 3; ++r) {
        const RuDev& R = Wb.ru[r];
        DwArgs d{X, cur, A, B.r[r], B.Cout, kDil[r], B.up_out, R.dw_w, R.dw_b, R.a1, R.i1, R.a2, R.i2};
        dwconv(d);
        tap(e, sid + 3 + 2 * r, A, B.r[r], B.Cout, n, first, st);
        GemmArgs a{};
        a.epi = EPI_RESID; a.A = A; a.lda = B.Cout; a.a_r = B.r[r]; a.W = R.pw_w; a.ldw = B.Cout; a.bias = R.pw_b;
        a.K = B.Cout; a.N = B.Cout; a.m_r = B.r[r]; a.out = Y; a.o_r = B.r[r]; a.ldo = B.Cout;
        a.R = X; a.r_r = cur; a.ldr = B.Cout; a.up = B.up_out;
        gemm(a);
        tap(e, sid + 4 + 2 * r, Y, B.r[r], B.Cout, n, first, st);
        std::swap(X, Y);
        cur = B.r[r];
      }
    }
    TailArgs t{X, P.b[3].r[2], tail_out, W.tail_alpha, W.tail_inv, W.tail_w, W.tail_b, d_status, wav, pcm, false};
    {
      const double smp = (double)n * tail_out.n();
      ProfScope ps(e, KC_TAIL, smp * (2.0 * 448 + 4.0 * 64), 4.0 * (double)n * P.b[3].r[2].n() * 64 + smp * 2.0, st);
      launch_tail(g, t);
    }
    int rc = check_launch(e, "layer pipeline");
    if (rc) return rc;
  }
  return SNACB_OK;
}

// Tensor-core recipe: fp16 operands through TMA + tcgen05, fp32 residual stream, fused epilogues.
// Buffers per chunk: Z (latent, fp32), X/Y (fp32 residual stream ping-pong), P/Q (fp16 GEMM operands).
int run_group_tc(snacb_engine* e, const Plan& P, const Item* d_items, int n_total, int out_len, Rng tail_out,
                 const int32_t* c0, const int32_t* c1, const int32_t* c2, int pitch0, const NoiseCfg& nz,
                 int32_t* d_status, float* wav, int16_t* pcm, char* ws_free, size_t ws_avail, int chunk,
                 int lanes, cudaStream_t st_caller) {
  const DevWeights& W = e->w;
  const size_t Zf = (size_t)P.z.n() * kLatent, S = P.max_stage_floats;
  const int F = P.T0 / 4;
  const int noise_off[4] = {0, 32 * F, 288 * F, 1312 * F};
  cudaError_t ce = cudaSuccess;
  const int n_chunks = (n_total + chunk - 1) / chunk;
  lanes = std::max(1, std::min(lanes, n_chunks));
  const size_t lane_bytes = lanes > 1 ? (ws_avail / lanes) & ~size_t(255) : ws_avail;
  if (lanes > 1) {
    while ((int)e->lane_streams.size() < lanes) {
      cudaStream_t s2; cudaEvent_t ev;
      CU(e, cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
      CU(e, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      e->lane_streams.push_back(s2); e->lane_done.push_back(ev);
    }
    if (!e->lane_fork) CU(e, cudaEventCreateWithFlags(&e->lane_fork, cudaEventDisableTiming));
    CU(e, cudaEventRecord(e->lane_fork, st_caller));
    for (int l = 0; l < lanes; ++l) CU(e, cudaStreamWaitEvent(e->lane_streams[l], e->lane_fork, 0));
  }

  for (int start = 0, ci = 0; start < n_total; start += chunk, ++ci) {
    const int n = std::min(chunk, n_total - start);
    const bool first = start == 0;
    const int lane = ci % lanes;
    cudaStream_t st = lanes > 1 ? e->lane_streams[lane] : st_caller;
    Bump bp(ws_free + (size_t)lane * (lanes > 1 ? lane_bytes : 0));
    float* Z = bp.take<float>(Zf * n);
    float* X = bp.take<float>(S * n);
    float* Y = bp.take<float>(S * n);
    __half* P16 = bp.take<__half>(S * n);
    __half* Q16 = bp.take<__half>(S * n);
    if (bp.off > lane_bytes) return fail(e, SNACB_ENOMEM, "workspace overflow in tensor-core pipeline");
    GroupCtx g{d_items ? d_items + start : nullptr, start, n, out_len, P.T0, st, &e->launches, e->cfg.flags};
    bool tail_done = false;

    auto gemm = [&](const TcGemmArgs& a) {
      if (ce != cudaSuccess) return;
      const double M = (double)n * a.a_rows, segs = a.epi == EPI_CONVT ? 2.0 : 1.0;
      const double flops = 2.0 * M * a.N * a.K * segs;
      const double by = 2.0 * (M * a.K * segs + (double)a.N * a.K * segs) +
                        (double)n * a.o_r.n() * a.ldo * ((a.out32 ? 4.0 : 0.0) + (a.out16 ? 2.0 : 0.0) + (a.R ? 4.0 : 0.0));
      ProfScope ps(e, a.epi == EPI_CONVT ? KC_CONVT : KC_GEMM1, flops, by, st);
      ce = launch_gemm_tc(g, a);
    };
    if (codes_head_supported(P.z) && e->tap_stage != 0) {  // from_codes + decoder.model.0 fused, z stays on chip
      const double el = (double)n * P.h.n() * kLatent;
      ProfScope ps(e, KC_CODES, 2.0 * 24 * kLatent * n * P.z.n() + el * 14.0, 2.0 * el + 28.0 * n * P.z.n(), st);
      launch_codes_head(g, W.q, c0, c1, c2, pitch0, P.z, P.h, W.head_dw_w, W.head_dw_b, P16);
    } else {
      {
        ProfScope ps(e, KC_CODES, 2.0 * 24 * kLatent * n * P.z.n(), 4.0 * kLatent * n * P.z.n(), st);
        launch_from_codes(g, W.q, c0, c1, c2, pitch0, P.z, Z);
      }
      tap(e, 0, Z, P.z, kLatent, n, first, st);
      // decoder.model.0: depthwise k7 -> fp16 operand of the 1x1
      DwArgs d{Z, P.z, nullptr, P.h, kLatent, 1, 1, W.head_dw_w, W.head_dw_b, nullptr, nullptr, nullptr, nullptr};
      const double el = (double)n * P.h.n() * kLatent;
      ProfScope ps(e, KC_DW, el * 14.0, 4.0 * (double)n * P.z.n() * kLatent + 2.0 * el, st);
      launch_dwconv_half(g, d, P16);
    }
    __half* Ain = Q16;    // block input operand (Snake'd, fp16)
    __half* Aother = P16;
    {  // decoder.model.1: 1x1 768->1024 + bias, then block 0's Snake, written as fp16
      TcGemmArgs a{};
      a.epi = EPI_BIAS; a.A = P16; a.K = kLatent; a.a_rows = P.h.n(); a.a_lo = P.h.lo; a.W = W.head_pw16; a.N = kDecDim;
      a.bias = W.head_pw_b; a.out16 = Ain; a.o_r = P.h; a.ldo = kDecDim; a.sn_alpha = W.blk[0].alpha; a.sn_inv = W.blk[0].inv;
      a.up = 1;
      if (e->tap_stage == 2) a.out32 = X;
      gemm(a);
      tap(e, 2, X, P.h, kDecDim, n, first, st);
    }
    for (int b = 0; b < 4 && ce == cudaSuccess; ++b) {
      const BlockPlan& B = P.b[b];
      const BlockDev& Wb = W.blk[b];
      const int sid = 3 + 9 * b;
      const bool noisy = nz.mode != SNACB_NOISE_OFF;
      __half* Y16 = Aother;  // transposed-conv output as the noise GEMM operand
      const bool fuse_cn = noisy && convt_noise_supported(B.Cin, B.Cout) && !(e->cfg.flags & SNACB_FLAG_NO_CONVT_NOISE_FUSION) &&
                           e->tap_stage != sid + 1;
      // blocks 1 / 3: ConvT + NoiseBlock as one GEMM over the composed stacked weight (built at load time when enabled)
      const bool compose = noisy && Wb.ctn16 != nullptr && e->tap_stage != sid + 1 &&
                           !(e->cfg.flags & (SNACB_FLAG_NO_CONVT_NOISE_FUSION | SNACB_FLAG_NO_CONVT_NOISE_COMPOSE));
      {
        TcGemmArgs a{};
        a.epi = EPI_CONVT; a.A = Ain; a.K = B.Cin; a.a_rows = B.in.n(); a.a_lo = B.in.lo; a.W = Wb.ct16; a.N = B.s * B.Cout;
        a.bias = Wb.ct_b; a.s = B.s; a.p = B.p; a.Cout = B.Cout; a.o_r = B.ct; a.ldo = B.Cout; a.up = B.up_out;
        if (compose) {
          // blocks 1 / 3: y + n (W_n y) from ONE GEMM over the stacked weight [W_c | W_n W_c] (used for every tick size)
          a.out32 = X;
          a.noise = NoiseSrc{nz.mode, nz.tensor, nz.stride, noise_off[b], (unsigned long long)nz.seed, nz.d_keys, b, nz.d_seed};
          if (ce == cudaSuccess) {
            const double M = (double)n * a.a_rows;
            ProfScope ps(e, KC_CONVT, 2.0 * M * a.N * a.K * 2.0 * 2.0,
                         2.0 * (M * a.K * 2.0 + 2.0 * (double)a.N * a.K * 2.0) + 4.0 * (double)n * B.ct.n() * B.Cout, st);
            ce = launch_convt_noise2_tc(g, a, Wb.ctn16, Wb.ct_b2);
          }
        } else if (fuse_cn) {
          a.out32 = X;
          a.noise = NoiseSrc{nz.mode, nz.tensor, nz.stride, noise_off[b], (unsigned long long)nz.seed, nz.d_keys, b, nz.d_seed};
          if (ce == cudaSuccess) {
            const double M = (double)n * a.a_rows;
            ProfScope ps(e, KC_CONVT, 2.0 * M * a.N * a.K * 2.0 + 2.0 * M * B.s * B.Cout * B.Cout,
                         2.0 * (M * a.K * 2.0 + (double)a.N * a.K * 2.0) + 4.0 * (double)n * B.ct.n() * B.Cout, st);
            ce = launch_convt_noise_tc(g, a, Wb.noise16);
          }
        } else {
          a.out32 = noisy ? Y : X; a.out16 = noisy ? Y16 : nullptr;
          gemm(a);
          tap(e, sid + 1, noisy ? Y : X, B.ct, B.Cout, n, first, st);
        }
      }
      if (noisy && !fuse_cn && !compose) {
        TcGemmArgs a{};
        a.epi = EPI_NOISE; a.A = Y16; a.K = B.Cout; a.a_rows = B.ct.n(); a.a_lo = B.ct.lo; a.W = Wb.noise16; a.N = B.Cout;
        a.out32 = X; a.o_r = B.ct; a.ldo = B.Cout; a.R = Y; a.r_r = B.ct; a.ldr = B.Cout; a.up = B.up_out;
        a.noise = NoiseSrc{nz.mode, nz.tensor, nz.stride, noise_off[b], (unsigned long long)nz.seed, nz.d_keys, b, nz.d_seed};
        gemm(a);
      }
      tap(e, sid + 2, X, B.ct, B.Cout, n, first, st);
      // block 3: the three ResidualUnits + decoder tail + PCM pack as ONE kernel (residual stream in tensor memory)
      if (b == 3 && blk_tc_supported(B.Cout, true) && !(e->cfg.flags & (SNACB_FLAG_NO_BLOCK_FUSION | SNACB_FLAG_NO_RU_FUSION |
                                                                      SNACB_FLAG_PERSISTENT_RU | SNACB_FLAG_TAIL_FUSION)) &&
          !(e->tap_stage > sid + 2 && e->tap_stage <= sid + 8) && ce == cudaSuccess) {
        BlkTcArgs u{};
        u.x = X; u.in_r = B.ct; u.C = B.Cout; u.up = B.up_out;
        double el = 0.0;
        for (int r = 0; r < 3; ++r) {
          const RuDev& R = Wb.ru[r];
          u.ru[r] = BlkTcArgs::Ru{R.dw_w, R.dw_b, R.a1, R.i1, R.a2, R.i2, R.pw_b, R.pw_w, R.pw16};
          el += (double)n * B.r[r].n() * B.Cout;
        }
        u.sn_alpha = W.tail_alpha; u.sn_inv = W.tail_inv; u.tail_w7 = W.tail_w; u.tail_b = W.tail_b; u.tail_out = tail_out;
        u.status = d_status; u.wav = wav; u.pcm = pcm;
        const double smp = (double)n * tail_out.n();
        ProfScope ps(e, KC_BLK, 2.0 * el * B.Cout + el * 24.0 + smp * (2.0 * 448 + 4.0 * 64),
                     (double)n * B.ct.n() * B.Cout * 4.0 + smp * 2.0, st);
        ce = launch_blk_tc(g, u);
        tail_done = true;
        continue;
      }
      // X holds the residual stream; Ain (dead after the transposed conv) becomes the dw output operand
      __half* D16 = Ain;
      __half* Anext = Aother;
      Rng cur = B.ct;
      for (int r = 0; r < 3 && ce == cudaSuccess; ++r) {
        const RuDev& R = Wb.ru[r];
        // block 1 (C = 256): fused persistent ResidualUnit kernel with the residual stream initialised in tensor memory
        if (ruw_tc_supported(B.Cout) && e->ruw && (B.Cout != 128 || e->ruw128) && !(e->cfg.flags & (SNACB_FLAG_NO_RU_FUSION | SNACB_FLAG_PERSISTENT_RU | SNACB_FLAG_FUSE_RU256)) &&
            e->tap_stage != sid + 4 + 2 * r) {  // every tick size: a window's samples must not depend on its tick
          const bool last = (r == 2) && (b < 3);
          const bool want32 = !last;
          RuTcArgs u{X, cur, B.r[r], B.Cout, kDil[r], B.up_out, R.dw_w, R.dw_b, R.a1, R.i1, R.a2, R.i2, R.pw16, R.pw_b,
                     want32 ? Y : nullptr, last ? Anext : nullptr, last ? W.blk[b + 1].alpha : nullptr,
                     last ? W.blk[b + 1].inv : nullptr, 0, false};
          const double el = (double)n * B.r[r].n() * B.Cout;
          ProfScope ps(e, KC_RUW, 2.0 * el * B.Cout + el * 24.0,
                       (double)n * cur.n() * B.Cout * 4.0 + el * ((want32 ? 4.0 : 0.0) + (last ? 2.0 : 0.0)), st);
          ce = launch_ruw_tc(g, u);
          tap(e, sid + 4 + 2 * r, Y, B.r[r], B.Cout, n, first, st);
          std::swap(X, Y);
          cur = B.r[r];
          continue;
        }
        const bool ru_persist = (e->cfg.flags & SNACB_FLAG_PERSISTENT_RU) != 0 || (B.Cout == 256 && (e->ru256 || (e->cfg.flags & SNACB_FLAG_FUSE_RU256)));
        if (ru_tc_supported(B.Cout, ru_persist) && !(e->cfg.flags & SNACB_FLAG_NO_RU_FUSION)) {  // fused dw + 1x1 + residual (blocks 2, 3)
          const bool last = (r == 2) && (b < 3);
          const bool want32 = !last || e->tap_stage == sid + 4 + 2 * r;
          RuTcArgs u{X, cur, B.r[r], B.Cout, kDil[r], B.up_out, R.dw_w, R.dw_b, R.a1, R.i1, R.a2, R.i2, R.pw16, R.pw_b,
                     want32 ? Y : nullptr, last ? Anext : nullptr, last ? W.blk[b + 1].alpha : nullptr,
                     last ? W.blk[b + 1].inv : nullptr, e->prefetch_ahead * (B.Cout == 64 ? 3 : 2), ru_persist};
          // last ResidualUnit of the decoder: fuse Snake(64) -> conv k7 64->1 -> tanh -> slice -> int16 pack
          const bool fuse_tail = (b == 3) && (r == 2) && !ru_persist && (e->cfg.flags & SNACB_FLAG_TAIL_FUSION) &&
                                 e->tap_stage != sid + 4 + 2 * r;
          if (fuse_tail) {
            u.out32 = nullptr; u.sn_alpha = W.tail_alpha; u.sn_inv = W.tail_inv;
            u.tail_w7 = W.tail_w; u.tail_b = W.tail_b; u.tail_out = tail_out; u.status = d_status; u.wav = wav; u.pcm = pcm;
            tail_done = true;
          }
          if (ce == cudaSuccess) {
            const double el = (double)n * B.r[r].n() * B.Cout;
            ProfScope ps(e, KC_RU, 2.0 * el * B.Cout + el * 24.0,
                         (double)n * cur.n() * B.Cout * 4.0 + el * ((want32 ? 4.0 : 0.0) + (last ? 2.0 : 0.0)), st);
            ce = launch_ru_tc(g, u);
          }
          tap(e, sid + 4 + 2 * r, Y, B.r[r], B.Cout, n, first, st);
          std::swap(X, Y);
          cur = B.r[r];
          continue;
        }
        {
          DwTcArgs d{X, cur, D16, B.r[r], B.Cout, kDil[r], B.up_out, R.dw_w, R.dw_b, R.a1, R.i1, R.a2, R.i2};
          const double el = (double)n * B.r[r].n() * B.Cout;
          ProfScope ps(e, KC_DW, el * 24.0, (double)n * cur.n() * B.Cout * 4.0 + el * 2.0, st);
          launch_dw_tc(g, d);
        }
        const bool last = (r == 2) && (b < 3);  // the block output only feeds the next block's Snake + ConvT
        TcGemmArgs a{};
        a.epi = EPI_RESID; a.A = D16; a.K = B.Cout; a.a_rows = B.r[r].n(); a.a_lo = B.r[r].lo; a.W = R.pw16; a.N = B.Cout;
        a.bias = R.pw_b; a.o_r = B.r[r]; a.ldo = B.Cout; a.R = X; a.r_r = cur; a.ldr = B.Cout; a.up = B.up_out;
        const bool want32 = !last || e->tap_stage == sid + 4 + 2 * r;
        a.out32 = want32 ? Y : nullptr;
        if (last) { a.out16 = Anext; a.sn_alpha = W.blk[b + 1].alpha; a.sn_inv = W.blk[b + 1].inv; }
        gemm(a);
        tap(e, sid + 4 + 2 * r, Y, B.r[r], B.Cout, n, first, st);
        std::swap(X, Y);
        cur = B.r[r];
      }
      Ain = Anext;
      Aother = D16;
    }
    if (ce != cudaSuccess) return fail(e, SNACB_ECUDA, "tensor-core GEMM launch failed: %s", cudaGetErrorString(ce));
    if (!tail_done) {
      const double smp = (double)n * tail_out.n();
      ProfScope ps(e, KC_TAIL, smp * (2.0 * 448 + 4.0 * 64), 4.0 * (double)n * P.b[3].r[2].n() * 64 + smp * 2.0, st);
      TailArgs t{X, P.b[3].r[2], tail_out, W.tail_alpha, W.tail_inv, W.tail_w, W.tail_b, d_status, wav, pcm, true};
      launch_tail(g, t);
    }
    int rc = check_launch(e, "tensor-core layer pipeline");
    if (rc) return rc;
  }
  if (lanes > 1) {
    for (int l = 0; l < lanes; ++l) {
      CU(e, cudaEventRecord(e->lane_done[l], e->lane_streams[l]));
      CU(e, cudaStreamWaitEvent(st_caller, e->lane_done[l], 0));
    }
  }
  return SNACB_OK;
}

size_t per_item_bytes_x3(const Plan& P) {
  const size_t Sb = pad256(P.max_stage_floats * 4);
  return pad256((size_t)P.z.n() * kLatent * 4) + pad256((size_t)P.h.n() * kLatent * 4) + 4 * Sb + 2048;
}

// Split-operand recipe (SNACB_PREC_FP16X3): the precision safety net for checkpoints whose activations need more than
// fp16's 11 bits / 65504 range at the GEMM operands.  Every GEMM-shaped layer still runs on tcgen05 (kind::f16), but on
// two-term fp16 splits of BOTH operands, three products per k-block (hi*hi + hi*lo + lo*hi, SURVEY App. E: 117 dB,
// <= 1 LSB); everything else is the fp32 CUDA-core recipe's kernels (exact sinf Snake, fp32 depthwise).  One split
// kernel per GEMM turns the fp32 operand into [hi | lo]; the GEMM's K loop runs over the concatenated segments.
int run_group_x3(snacb_engine* e, const Plan& P, const Item* d_items, int n_total, int out_len, Rng tail_out,
                 const int32_t* c0, const int32_t* c1, const int32_t* c2, int pitch0, const NoiseCfg& nz,
                 int32_t* d_status, float* wav, int16_t* pcm, char* ws_free, size_t ws_avail, cudaStream_t st) {
  const DevWeights& W = e->w;
  const size_t Zf = (size_t)P.z.n() * kLatent, Hf = (size_t)P.h.n() * kLatent, S = P.max_stage_floats;
  const size_t per_item = per_item_bytes_x3(P);
  int chunk = e->cfg.chunk_items > 0 ? e->cfg.chunk_items : 256;
  if ((size_t)chunk * per_item > ws_avail) chunk = (int)(ws_avail / per_item);
  if (chunk < 1) return fail(e, SNACB_ENOMEM, "workspace too small for one item");
  const int F = P.T0 / 4;
  const int noise_off[4] = {0, 32 * F, 288 * F, 1312 * F};
  cudaError_t ce = cudaSuccess;
  for (int start = 0; start < n_total && ce == cudaSuccess; start += chunk) {
    const int n = std::min(chunk, n_total - start);
    Bump bp(ws_free);
    float* Z = bp.take<float>(Zf * n);
    float* H0 = bp.take<float>(Hf * n);
    float* X = bp.take<float>(S * n);
    float* Y = bp.take<float>(S * n);
    float* A = bp.take<float>(S * n);
    __half* A16 = bp.take<__half>(2 * S * n);
    GroupCtx g{d_items ? d_items + start : nullptr, start, n, out_len, P.T0, st, &e->launches, e->cfg.flags};
    auto gemm = [&](TcGemmArgs a, const float* a32, Rng a_r, int K, const float* sn_a, const float* sn_i) {
      if (ce != cudaSuccess) return;
      {
        ProfScope ps(e, KC_SNAKE, 4.0 * n * a_r.n() * K, 8.0 * n * a_r.n() * K, st);
        launch_split16(a32, A16, (size_t)n * a_r.n(), K, sn_a, sn_i, st, &e->launches);
      }
      a.A = A16; a.K = K; a.a_rows = a_r.n(); a.a_lo = a_r.lo; a.split = 1;
      const double M = (double)n * a.a_rows, segs = (a.epi == EPI_CONVT ? 2.0 : 1.0) * 3.0;
      ProfScope ps(e, a.epi == EPI_CONVT ? KC_CONVT : KC_GEMM1, 2.0 * M * a.N * a.K * segs,
                   2.0 * (M * a.K * 2.0 + (double)a.N * a.K * segs) + (double)n * a.o_r.n() * a.ldo * (4.0 + (a.R ? 4.0 : 0.0)), st);
      ce = launch_gemm_tc(g, a);
    };
    auto dwconv = [&](const DwArgs& d) {
      const double el = (double)n * d.out_r.n() * d.C;
      ProfScope ps(e, KC_DW, el * (14.0 + (d.a1 ? 28.0 : 0.0) + (d.a2 ? 4.0 : 0.0)), 4.0 * ((double)n * d.in_r.n() * d.C + el), st);
      launch_dwconv(g, d);
    };
    {
      ProfScope ps(e, KC_CODES, 2.0 * 24 * kLatent * n * P.z.n(), 4.0 * kLatent * n * P.z.n(), st);
      launch_from_codes(g, W.q, c0, c1, c2, pitch0, P.z, Z);
    }
    dwconv(DwArgs{Z, P.z, H0, P.h, kLatent, 1, 1, W.head_dw_w, W.head_dw_b, nullptr, nullptr, nullptr, nullptr});
    {
      TcGemmArgs a{};
      a.epi = EPI_BIAS; a.W = W.head_pw16; a.N = kDecDim; a.bias = W.head_pw_b; a.out32 = X; a.o_r = P.h; a.ldo = kDecDim; a.up = 1;
      gemm(a, H0, P.h, kLatent, nullptr, nullptr);
    }
    for (int b = 0; b < 4 && ce == cudaSuccess; ++b) {
      const BlockPlan& B = P.b[b];
      const BlockDev& Wb = W.blk[b];
      {
        TcGemmArgs a{};
        a.epi = EPI_CONVT; a.W = Wb.ct16; a.N = B.s * B.Cout; a.bias = Wb.ct_b; a.s = B.s; a.p = B.p; a.Cout = B.Cout;
        a.out32 = Y; a.o_r = B.ct; a.ldo = B.Cout; a.up = B.up_out;
        gemm(a, X, B.in, B.Cin, Wb.alpha, Wb.inv);  // the block's Snake is applied by the split kernel
      }
      if (nz.mode != SNACB_NOISE_OFF) {
        TcGemmArgs a{};
        a.epi = EPI_NOISE; a.W = Wb.noise16; a.N = B.Cout; a.out32 = X; a.o_r = B.ct; a.ldo = B.Cout;
        a.R = Y; a.r_r = B.ct; a.ldr = B.Cout; a.up = B.up_out;
        a.noise = NoiseSrc{nz.mode, nz.tensor, nz.stride, noise_off[b], (unsigned long long)nz.seed, nz.d_keys, b, nz.d_seed};
        gemm(a, Y, B.ct, B.Cout, nullptr, nullptr);
      } else {
        std::swap(X, Y);
      }
      Rng cur = B.ct;
      for (int r = 0; r < 3 && ce == cudaSuccess; ++r) {
        const RuDev& R = Wb.ru[r];
        dwconv(DwArgs{X, cur, A, B.r[r], B.Cout, kDil[r], B.up_out, R.dw_w, R.dw_b, R.a1, R.i1, R.a2, R.i2});
        TcGemmArgs a{};
        a.epi = EPI_RESID; a.W = R.pw16; a.N = B.Cout; a.bias = R.pw_b; a.out32 = Y; a.o_r = B.r[r]; a.ldo = B.Cout;
        a.R = X; a.r_r = cur; a.ldr = B.Cout; a.up = B.up_out;
        gemm(a, A, B.r[r], B.Cout, nullptr, nullptr);
        std::swap(X, Y);
        cur = B.r[r];
      }
    }
    if (ce != cudaSuccess) return fail(e, SNACB_ECUDA, "split-operand GEMM launch failed: %s", cudaGetErrorString(ce));
    {
      const double smp = (double)n * tail_out.n();
      ProfScope ps(e, KC_TAIL, smp * (2.0 * 448 + 4.0 * 64), 4.0 * (double)n * P.b[3].r[2].n() * 64 + smp * 2.0, st);
      TailArgs t{X, P.b[3].r[2], tail_out, W.tail_alpha, W.tail_inv, W.tail_w, W.tail_b, d_status, wav, pcm, false};
      launch_tail(g, t);
    }
    int rc = check_launch(e, "split-operand layer pipeline");
    if (rc) return rc;
  }
  return SNACB_OK;
}

// ---- sizing shared by every entry point
inline bool is_tc(const snacb_engine* e) { return e->cfg.precision == SNACB_PREC_FP16 || e->cfg.precision == SNACB_PREC_FP16X3; }
Plan plan_for(const snacb_engine* e, int T0, Rng out, bool clip) {
  return make_plan(T0, out, clip, is_tc(e));
}
size_t per_item_bytes(const snacb_engine* e, const Plan& P) {
  const size_t Zb = pad256((size_t)P.z.n() * kLatent * 4), Sb = pad256(P.max_stage_floats * 4);
  if (e->cfg.precision == SNACB_PREC_FP16) return Zb + 2 * Sb + 2 * pad256(P.max_stage_floats * 2) + 2048;
  if (e->cfg.precision == SNACB_PREC_FP16X3) return per_item_bytes_x3(P);
  return Zb + pad256((size_t)P.h.n() * kLatent * 4) + 3 * Sb + 1024;
}
constexpr size_t kActBudget = size_t(6) << 30;  // activation workspace cap per engine
int default_chunk(const snacb_engine* e) { return e->cfg.chunk_items > 0 ? e->cfg.chunk_items : (e->cfg.precision == SNACB_PREC_FP16X3 ? 256 : 1024); }
int default_lanes(const snacb_engine* e) { return e->cfg.precision == SNACB_PREC_FP16 ? std::max(1, std::min(e->cfg.lanes, 8)) : 1; }
size_t act_bytes(const snacb_engine* e, const Plan& P, int count) {
  const size_t per = per_item_bytes(e, P);
  const int chunk = std::max(1, std::min(count, default_chunk(e)));
  const int lanes = std::max(1, std::min(default_lanes(e), (count + chunk - 1) / chunk));
  return std::max(per, std::min(kActBudget, per * chunk * lanes)) + 4096 + 256 * (size_t)lanes;
}

int run_group(snacb_engine* e, const Plan& P, const Item* d_items, int n_total, int out_len, Rng tail_out,
              const int32_t* c0, const int32_t* c1, const int32_t* c2, int pitch0, const NoiseCfg& nz,
              int32_t* d_status, float* wav, int16_t* pcm, char* ws_free, size_t ws_avail, cudaStream_t st) {
  if (e->cfg.precision == SNACB_PREC_FP16) {
    const size_t per = per_item_bytes(e, P);
    int chunk = std::min(default_chunk(e), std::max(1, n_total));
    int lanes = std::max(1, std::min(default_lanes(e), (n_total + chunk - 1) / chunk));
    if ((size_t)chunk * per * lanes + 256 * (size_t)lanes > ws_avail) { lanes = 1; }
    if ((size_t)chunk * per > ws_avail) chunk = (int)(ws_avail / per);
    if (chunk < 1) return fail(e, SNACB_ENOMEM, "workspace too small for one item");
    return run_group_tc(e, P, d_items, n_total, out_len, tail_out, c0, c1, c2, pitch0, nz, d_status, wav, pcm, ws_free,
                        ws_avail, chunk, lanes, st);
  }
  if (e->cfg.precision == SNACB_PREC_FP16X3)
    return run_group_x3(e, P, d_items, n_total, out_len, tail_out, c0, c1, c2, pitch0, nz, d_status, wav, pcm, ws_free, ws_avail, st);
  return run_group_f32(e, P, d_items, n_total, out_len, tail_out, c0, c1, c2, pitch0, nz, d_status, wav, pcm, ws_free,
                       ws_avail, st);
}

int upload_items(snacb_engine* e, const std::vector<Item>& items, Item* d_dst, cudaStream_t st) {
  if (!e->items_ev) CU(e, cudaEventCreateWithFlags(&e->items_ev, cudaEventDisableTiming));
  else CU(e, cudaEventSynchronize(e->items_ev));
  if (items.size() > e->pin_items_cap) {
    if (e->pin_items) CU(e, cudaFreeHost(e->pin_items));
    e->pin_items_cap = items.size() * 2 + 64;
    CU(e, cudaMallocHost((void**)&e->pin_items, e->pin_items_cap * sizeof(Item)));
  }
  memcpy(e->pin_items, items.data(), items.size() * sizeof(Item));
  CU(e, cudaMemcpyAsync(d_dst, e->pin_items, items.size() * sizeof(Item), cudaMemcpyHostToDevice, st));
  CU(e, cudaEventRecord(e->items_ev, st));
  return SNACB_OK;
}

}  // namespace

// ====================================================================================== C ABI
extern "C" {

int snacb_create(snacb_engine** out, const snacb_config* cfg) {
  if (!out || !cfg) return fail(nullptr, SNACB_EINVAL, "snacb_create: null argument");
  if (cfg->abi_version != SNACB_ABI_VERSION) return fail(nullptr, SNACB_EINVAL, "snacb_create: ABI version mismatch");
  int ndev = 0;
  cudaError_t c = cudaGetDeviceCount(&ndev);
  if (c != cudaSuccess || ndev == 0)
    return fail(nullptr, SNACB_ECUDA, "snacb_create: no CUDA device (%s); there is no CPU fallback",
                c == cudaSuccess ? "device count 0" : cudaGetErrorString(c));
  if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, SNACB_EINVAL, "snacb_create: bad device ordinal %d", cfg->device);
  cudaDeviceProp prop{};
  cudaGetDeviceProperties(&prop, cfg->device);
  if (prop.major != 10)
    return fail(nullptr, SNACB_ECUDA, "snacb_create: device %d is sm_%d%d; this library is built for sm_100a only",
                cfg->device, prop.major, prop.minor);
  if (cfg->precision != SNACB_PREC_FP32 && cfg->precision != SNACB_PREC_FP16 && cfg->precision != SNACB_PREC_FP16X3)
    return fail(nullptr, SNACB_EINVAL, "snacb_create: unknown precision %d", cfg->precision);
  c = cudaSetDevice(cfg->device);
  if (c != cudaSuccess) return fail(nullptr, SNACB_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(c));
  snacb_engine* e = new snacb_engine();
  e->cfg = *cfg;
  e->device = cfg->device;
  {
    const char* pf = getenv("SNACB_PREFETCH");
    e->prefetch_ahead = (pf && pf[0] == '0') ? 0 : prop.multiProcessorCount;
    const char* r2 = getenv("SNACB_RU256");
    e->ru256 = r2 && r2[0] == '1';
    const char* rw = getenv("SNACB_RUW");
    e->ruw = !(rw && rw[0] == '0');
    const char* rw2 = getenv("SNACB_RUW128");
    e->ruw128 = !(rw2 && rw2[0] == '0');
    const char* cn2 = getenv("SNACB_CONVT_N2");  // bit mask over the blocks (kernels_tc.cu), "0" = none
    e->convt_n2 = !(cn2 && cn2[0] == '0' && cn2[1] == 0) && !(cfg->flags & SNACB_FLAG_NO_CONVT_NOISE_COMPOSE);
    const char* gr = getenv("SNACB_GRAPHS");
    if (gr) e->graph_max_win = atoi(gr);  // 0 disables the CUDA-graph path
  }
  *out = e;
  return SNACB_OK;
}

void snacb_destroy(snacb_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  if (e->warena) cudaFree(e->warena);
  if (e->harena) cudaFree(e->harena);
  if (e->earena) cudaFree(e->earena);
  if (e->ws) cudaFree(e->ws);
  if (e->dstage) cudaFree(e->dstage);
  for (auto& sl : e->pipe) {
    if (sl.dst) cudaFree(sl.dst);
    if (sl.pin) cudaFreeHost(sl.pin);
    if (sl.tail_done) cudaEventDestroy(sl.tail_done);
    if (sl.copied) cudaEventDestroy(sl.copied);
  }
  if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
  if (e->pin) cudaFreeHost(e->pin);
  if (e->pin_items) cudaFreeHost(e->pin_items);
  if (e->items_ev) cudaEventDestroy(e->items_ev);
  for (auto& kv : e->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  if (e->cap_stream) cudaStreamDestroy(e->cap_stream);
  for (auto sidestream : e->lane_streams) cudaStreamDestroy(sidestream);
  for (auto ev : e->lane_done) cudaEventDestroy(ev);
  if (e->lane_fork) cudaEventDestroy(e->lane_fork);
  for (auto& r : e->prof.recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  for (auto ev : e->prof.pool) cudaEventDestroy(ev);
  delete e;
}

const char* snacb_last_error(const snacb_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

size_t snacb_workspace_bytes(const snacb_engine* e) { return e ? e->ws_bytes + e->dstage_bytes : 0; }

int64_t snacb_launch_count(const snacb_engine* e) { return e ? e->launches : 0; }
int64_t snacb_graph_launch_count(const snacb_engine* e) { return e ? e->graph_launches : 0; }

int snacb_load_weights(snacb_engine* e, const snacb_weights* w) {
  if (!e || !w) return fail(e, SNACB_EINVAL, "snacb_load_weights: null argument");
  CU(e, cudaSetDevice(e->device));
  HostPack hp;
  struct Fix { float** dst; size_t off; };
  std::vector<Fix> fix;
  DevWeights& D = e->w;
  auto put = [&](float** dst, const float* src, size_t n) { fix.push_back({dst, hp.add(src, n)}); };
  const float* must[] = {w->head_dw_w, w->head_dw_b, w->head_pw_w, w->head_pw_b, w->tail_alpha, w->tail_w, w->tail_b};
  for (const float* p : must) if (!p) return fail(e, SNACB_EINVAL, "snacb_load_weights: null tensor");
  for (int l = 0; l < 3; ++l) {
    if (!w->codebook[l] || !w->outproj_w[l] || !w->outproj_b[l]) return fail(e, SNACB_EINVAL, "snacb_load_weights: null quantizer tensor");
    put(const_cast<float**>(&D.q.codebook[l]), w->codebook[l], (size_t)SNACB_CODEBOOK_SIZE * 8);
    put(const_cast<float**>(&D.q.w[l]), w->outproj_w[l], (size_t)kLatent * 8);
    put(const_cast<float**>(&D.q.b[l]), w->outproj_b[l], kLatent);
  }
  fix.push_back({&D.head_dw_w, add_transposed(hp, w->head_dw_w, kLatent, 7)});
  put(&D.head_dw_b, w->head_dw_b, kLatent);
  put(&D.head_pw_w, w->head_pw_w, (size_t)kDecDim * kLatent);
  put(&D.head_pw_b, w->head_pw_b, kDecDim);
  int cin = kDecDim;
  for (int b = 0; b < 4; ++b) {
    const snacb_block_weights& s = w->block[b];
    const int cout = cin / 2, st = kRates[b];
    if (!s.alpha || !s.convt_w || !s.convt_b || !s.noise_w) return fail(e, SNACB_EINVAL, "snacb_load_weights: null block tensor");
    put(&D.blk[b].alpha, s.alpha, cin);
    fix.push_back({&D.blk[b].inv, add_inv(hp, s.alpha, cin)});
    fix.push_back({&D.blk[b].ct_w, add_convt(hp, s.convt_w, cin, cout, st)});
    put(&D.blk[b].ct_b, s.convt_b, cout);
    put(&D.blk[b].noise_w, s.noise_w, (size_t)cout * cout);
    {
      const size_t off = hp.add(nullptr, cout);
      for (int o = 0; o < cout; ++o) {
        double acc = 0.0;
        for (int c = 0; c < cout; ++c) acc += (double)s.noise_w[(size_t)o * cout + c] * s.convt_b[c];
        hp.data[off + o] = (float)acc;
      }
      fix.push_back({&D.blk[b].ct_b2, off});
    }
    for (int r = 0; r < 3; ++r) {
      const snacb_ru_weights& u = s.ru[r];
      RuDev& R = D.blk[b].ru[r];
      if (!u.alpha1 || !u.dw_w || !u.dw_b || !u.alpha2 || !u.pw_w || !u.pw_b) return fail(e, SNACB_EINVAL, "snacb_load_weights: null RU tensor");
      put(&R.a1, u.alpha1, cout);
      fix.push_back({&R.i1, add_inv(hp, u.alpha1, cout)});
      fix.push_back({&R.dw_w, add_transposed(hp, u.dw_w, cout, 7)});
      put(&R.dw_b, u.dw_b, cout);
      put(&R.a2, u.alpha2, cout);
      fix.push_back({&R.i2, add_inv(hp, u.alpha2, cout)});
      put(&R.pw_w, u.pw_w, (size_t)cout * cout);
      put(&R.pw_b, u.pw_b, cout);
    }
    cin = cout;
  }
  put(&D.tail_alpha, w->tail_alpha, 64);
  fix.push_back({&D.tail_inv, add_inv(hp, w->tail_alpha, 64)});
  fix.push_back({&D.tail_w, add_transposed(hp, w->tail_w, 64, 7)});
  put(&D.tail_b, w->tail_b, 1);

  const size_t bytes = hp.data.size() * sizeof(float);
  CU(e, cudaDeviceSynchronize());
  if (e->warena) { CU(e, cudaFree(e->warena)); e->warena = nullptr; }
  CU(e, cudaMalloc(&e->warena, bytes));
  e->warena_bytes = bytes;
  CU(e, cudaMemcpy(e->warena, hp.data.data(), bytes, cudaMemcpyHostToDevice));
  for (const Fix& f : fix) *f.dst = reinterpret_cast<float*>(e->warena) + f.off;

  // fp16 operand copies for the tensor-core recipe
  if (e->harena) { CU(e, cudaFree(e->harena)); e->harena = nullptr; }
  if (is_tc(e)) {
    const bool x3 = e->cfg.precision == SNACB_PREC_FP16X3;  // operands [N][nseg * 3K] = per tap [hi | lo | hi]
    struct H { __half** dst; const float* src; int N, nseg, K; size_t off; };
    std::vector<H> hs;
    size_t hoff = 0;
    auto hput = [&](__half** dst, const float* src, int N, int nseg, int K) {
      hs.push_back({dst, src, N, nseg, K, hoff});
      hoff += ((size_t)N * nseg * K * (x3 ? 3 : 1) + 127) & ~size_t(127);
    };
    hput(&D.head_pw16, D.head_pw_w, kDecDim, 1, kLatent);
    int ci = kDecDim;
    for (int b = 0; b < 4; ++b) {
      const int co = ci / 2;
      hput(&D.blk[b].ct16, D.blk[b].ct_w, kRates[b] * co, 2, ci);
      hput(&D.blk[b].noise16, D.blk[b].noise_w, co, 1, co);
      for (int r = 0; r < 3; ++r) hput(&D.blk[b].ru[r].pw16, D.blk[b].ru[r].pw_w, co, 1, co);
      ci = co;
    }
    // composed ConvT + NoiseBlock weights of the wide blocks (single-pass fp16 recipe only)
    size_t ctn_off[4] = {0, 0, 0, 0};
    {
      int c_in = kDecDim;
      for (int b = 0; b < 4; ++b) {
        const int co = c_in / 2;
        D.blk[b].ctn16 = nullptr;
        if (!x3 && e->convt_n2 && convt_noise2_supported(c_in, co)) {
          ctn_off[b] = hoff + 1;  // +1: marks "present"
          hoff += ((size_t)2 * kRates[b] * co * 2 * c_in + 127) & ~size_t(127);
        }
        c_in = co;
      }
    }
    CU(e, cudaMalloc((void**)&e->harena, hoff * sizeof(__half)));
    for (const H& h : hs) {
      *h.dst = e->harena + h.off;
      if (x3) launch_split_w(h.src, *h.dst, h.N, h.nseg, h.K, 0);
      else launch_to_half(h.src, *h.dst, (size_t)h.N * h.nseg * h.K, 0);
    }
    {
      int c_in = kDecDim;
      for (int b = 0; b < 4; ++b) {
        const int co = c_in / 2;
        if (ctn_off[b]) {
          D.blk[b].ctn16 = e->harena + (ctn_off[b] - 1);
          launch_compose_ctn(D.blk[b].ct_w, D.blk[b].noise_w, co, kRates[b] * co, 2 * c_in, D.blk[b].ctn16, 0);
        }
        c_in = co;
      }
    }
    CU(e, cudaGetLastError());
    CU(e, cudaDeviceSynchronize());
  }
  e->loaded = true;
  ++e->gen;
  return SNACB_OK;
}

// ---------------------------------------------------------------------------------- encoder (N4)
int snacb_load_encoder_weights(snacb_engine* e, const snacb_encoder_weights* w) {
  if (!e || !w) return fail(e, SNACB_EINVAL, "snacb_load_encoder_weights: null argument");
  if (!e->loaded) return fail(e, SNACB_ESTATE, "snacb_load_encoder_weights: load the decode-path weights first");
  CU(e, cudaSetDevice(e->device));
  if (!w->in_w || !w->in_b || !w->out_dw_w || !w->out_dw_b) return fail(e, SNACB_EINVAL, "snacb_load_encoder_weights: null tensor");
  HostPack hp;
  struct Fix { float** dst; size_t off; };
  std::vector<Fix> fix;
  EncDev& D = e->enc;
  auto put = [&](float** dst, const float* src, size_t n) { fix.push_back({dst, hp.add(src, n)}); };
  put(&D.in_w, w->in_w, (size_t)kEncDim * 7);
  put(&D.in_b, w->in_b, kEncDim);
  int C = kEncDim;
  for (int b = 0; b < 4; ++b) {
    const snacb_enc_block_weights& sb = w->block[b];
    const int st = kEncRates[b], k = 2 * st;
    if (!sb.alpha || !sb.down_w || !sb.down_b) return fail(e, SNACB_EINVAL, "snacb_load_encoder_weights: null block tensor");
    for (int r = 0; r < 3; ++r) {
      const snacb_ru_weights& u = sb.ru[r];
      RuDev& R = D.blk[b].ru[r];
      if (!u.alpha1 || !u.dw_w || !u.dw_b || !u.alpha2 || !u.pw_w || !u.pw_b) return fail(e, SNACB_EINVAL, "snacb_load_encoder_weights: null RU tensor");
      put(&R.a1, u.alpha1, C);
      fix.push_back({&R.i1, add_inv(hp, u.alpha1, C)});
      fix.push_back({&R.dw_w, add_transposed(hp, u.dw_w, C, 7)});
      put(&R.dw_b, u.dw_b, C);
      put(&R.a2, u.alpha2, C);
      fix.push_back({&R.i2, add_inv(hp, u.alpha2, C)});
      put(&R.pw_w, u.pw_w, (size_t)C * C);
      put(&R.pw_b, u.pw_b, C);
    }
    put(&D.blk[b].alpha, sb.alpha, C);
    fix.push_back({&D.blk[b].inv, add_inv(hp, sb.alpha, C)});
    {  // Conv1d weight [2C][C][k] -> GEMM operand [2C][k*C]: column kk*C + ci
      const size_t off = hp.add(nullptr, (size_t)2 * C * k * C);
      for (int co = 0; co < 2 * C; ++co)
        for (int ci = 0; ci < C; ++ci)
          for (int kk = 0; kk < k; ++kk)
            hp.data[off + ((size_t)co * k + kk) * C + ci] = sb.down_w[((size_t)co * C + ci) * k + kk];
      fix.push_back({&D.blk[b].down_w, off});
    }
    put(&D.blk[b].down_b, sb.down_b, 2 * C);
    C *= 2;
  }
  fix.push_back({&D.out_dw_w, add_transposed(hp, w->out_dw_w, kLatent, 7)});
  put(&D.out_dw_b, w->out_dw_b, kLatent);
  std::vector<float> cb((size_t)SNACB_CODEBOOK_SIZE * 8);
  for (int l = 0; l < 3; ++l) {
    if (!w->inproj_w[l] || !w->inproj_b[l]) return fail(e, SNACB_EINVAL, "snacb_load_encoder_weights: null in_proj tensor");
    put(&D.inproj_w[l], w->inproj_w[l], (size_t)8 * kLatent);
    put(&D.inproj_b[l], w->inproj_b[l], 8);
    CU(e, cudaMemcpy(cb.data(), e->w.q.codebook[l], cb.size() * sizeof(float), cudaMemcpyDeviceToHost));
    const size_t off = hp.add(nullptr, cb.size());
    for (int k = 0; k < SNACB_CODEBOOK_SIZE; ++k) {  // F.normalize(codebook): x / max(||x||_2, 1e-12)
      float n2 = 0.0f;
      for (int d = 0; d < 8; ++d) n2 += cb[(size_t)k * 8 + d] * cb[(size_t)k * 8 + d];
      const float rn = 1.0f / std::max(std::sqrt(n2), 1e-12f);
      for (int d = 0; d < 8; ++d) hp.data[off + (size_t)k * 8 + d] = cb[(size_t)k * 8 + d] * rn;
    }
    fix.push_back({&D.cb_norm[l], off});
  }
  const size_t bytes = hp.data.size() * sizeof(float);
  CU(e, cudaDeviceSynchronize());
  if (e->earena) { CU(e, cudaFree(e->earena)); e->earena = nullptr; }
  CU(e, cudaMalloc(&e->earena, bytes));
  CU(e, cudaMemcpy(e->earena, hp.data.data(), bytes, cudaMemcpyHostToDevice));
  for (const Fix& f : fix) *f.dst = reinterpret_cast<float*>(e->earena) + f.off;
  e->enc_loaded = true;
  return SNACB_OK;
}

int snacb_encode(snacb_engine* e, const float* d_audio, int32_t batch, int32_t n_samples, int32_t* d_c0, int32_t* d_c1,
                 int32_t* d_c2, float* d_latent, void* stream) {
  if (!e) return SNACB_EINVAL;
  if (!e->enc_loaded) return fail(e, SNACB_ESTATE, "snacb_encode: encoder weights not loaded");
  if (!d_audio || !d_c0 || !d_c1 || !d_c2 || batch < 0 || n_samples <= 0 || n_samples % 2048)
    return fail(e, SNACB_EINVAL, "snacb_encode: bad argument (n_samples must be a positive multiple of 2048)");
  if (batch == 0) return SNACB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  CU(e, cudaSetDevice(e->device));
  const EncDev& D = e->enc;
  // largest stage: [T + 2][48] (the padded copy before the first strided conv); three ping-pong buffers
  const size_t S = (size_t)(n_samples + 8) * kEncDim;
  const size_t per_item = 3 * pad256(S * 4) + 1024;
  int rc = ensure_ws(e, std::min<size_t>(per_item * batch, std::max<size_t>(per_item, kActBudget)), st);
  if (rc) return rc;
  const int chunk = (int)std::max<size_t>(1, std::min<size_t>(batch, e->ws_bytes / per_item));
  int32_t* codes[3] = {d_c0, d_c1, d_c2};
  for (int start = 0; start < batch; start += chunk) {
    const int n = std::min(chunk, batch - start);
    Bump bp(e->ws);
    float* X = bp.take<float>(S * n);
    float* Y = bp.take<float>(S * n);
    float* A = bp.take<float>(S * n);
    int T = n_samples, C = kEncDim;
    {
      ProfScope ps(e, KC_DW, 14.0 * n * T * C, 4.0 * n * T * (1.0 + C), st);
      launch_enc_in(d_audio + (size_t)start * n_samples, n, T, D.in_w, D.in_b, X, st, &e->launches);
    }
    for (int b = 0; b < 4; ++b) {
      const EncBlockDev& Wb = D.blk[b];
      const int s = kEncRates[b], p = (s + 1) / 2;
      GroupCtx g{nullptr, 0, n, 0, T, st, &e->launches, e->cfg.flags};  // rows [0, T) are the whole sequence at this rate
      const Rng all{0, T};
      for (int r = 0; r < 3; ++r) {
        const RuDev& R = Wb.ru[r];
        {
          ProfScope ps(e, KC_DW, 46.0 * n * T * C, 8.0 * n * T * C, st);
          launch_dwconv(g, DwArgs{X, all, A, all, C, kDil[r], 1, R.dw_w, R.dw_b, R.a1, R.i1, R.a2, R.i2});
        }
        GemmArgs a{};
        a.epi = EPI_RESID; a.A = A; a.lda = C; a.a_r = all; a.W = R.pw_w; a.ldw = C; a.bias = R.pw_b; a.K = C; a.N = C;
        a.m_r = all; a.out = Y; a.o_r = all; a.ldo = C; a.R = X; a.r_r = all; a.ldr = C; a.up = 1;
        ProfScope ps(e, KC_GEMM1, 2.0 * n * T * C * C, 12.0 * n * T * C, st);
        launch_gemm_f32(g, a);
        std::swap(X, Y);
      }
      {
        ProfScope ps(e, KC_SNAKE, 4.0 * n * T * C, 8.0 * n * T * C, st);
        launch_snake_pad(X, A, n, T, C, s, p, Wb.alpha, Wb.inv, st, &e->launches);
      }
      const int To = T / s;
      GroupCtx go{nullptr, 0, n, 0, To, st, &e->launches, e->cfg.flags};
      GemmArgs a{};
      a.epi = EPI_BIAS; a.A = A; a.lda = s * C; a.a_r = Rng{0, (T + s) / s}; a.W = Wb.down_w; a.ldw = 2 * s * C; a.bias = Wb.down_b;
      a.K = 2 * s * C; a.N = 2 * C; a.m_r = Rng{0, To}; a.out = Y; a.o_r = Rng{0, To}; a.ldo = 2 * C; a.up = 1;
      {
        ProfScope ps(e, KC_GEMM1, 2.0 * n * To * a.K * a.N, 4.0 * n * (T * C + To * 2.0 * C), st);
        launch_gemm_f32(go, a);
      }
      std::swap(X, Y);
      T = To; C *= 2;
    }
    {
      GroupCtx g{nullptr, 0, n, 0, T, st, &e->launches, e->cfg.flags};
      ProfScope ps(e, KC_DW, 14.0 * n * T * C, 8.0 * n * T * C, st);
      launch_dwconv(g, DwArgs{X, Rng{0, T}, Y, Rng{0, T}, C, 1, 1, D.out_dw_w, D.out_dw_b, nullptr, nullptr, nullptr, nullptr});
    }
    if (d_latent)
      CU(e, cudaMemcpyAsync(d_latent + (size_t)start * T * kLatent, Y, (size_t)n * T * kLatent * 4, cudaMemcpyDeviceToDevice, st));
    for (int l = 0; l < 3; ++l) {  // residual VQ: Y is the running residual
      ProfScope ps(e, KC_CODES, 2.0 * n * (T / kVqStrides[l]) * (8.0 * kLatent * 2 + 8.0 * SNACB_CODEBOOK_SIZE), 8.0 * n * T * kLatent, st);
      launch_vq_level(Y, n, T, kVqStrides[l], D.inproj_w[l], D.inproj_b[l], D.cb_norm[l], e->w.q.codebook[l], e->w.q.w[l], e->w.q.b[l],
                      codes[l] + (size_t)start * (T / kVqStrides[l]), st, &e->launches);
    }
    rc = check_launch(e, "encoder pipeline");
    if (rc) return rc;
  }
  return SNACB_OK;
}

static int deinterleave_common(snacb_engine* e, const int32_t* d_tokens, int32_t tokens_stride, const int32_t* h_ntok,
                               int32_t ntok_uniform, int32_t n_win, int32_t max_frames, bool raw, int32_t* d_c0,
                               int32_t* d_c1, int32_t* d_c2, int32_t* d_status, void* stream) {
  if (!e) return SNACB_EINVAL;
  if (n_win < 0 || max_frames < 1 || !d_tokens || !d_c0 || !d_c1 || !d_c2 || !d_status)
    return fail(e, SNACB_EINVAL, "snacb_deinterleave: bad argument");
  if (n_win == 0) return SNACB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  CU(e, cudaSetDevice(e->device));
  const int32_t* d_ntok = nullptr;
  if (h_ntok) {
    for (int i = 0; i < n_win; ++i)
      if (h_ntok[i] < 0 || h_ntok[i] > tokens_stride) return fail(e, SNACB_EINVAL, "snacb_deinterleave: ntok[%d] out of range", i);
    const size_t need = pad256((size_t)n_win * 4);
    if (need > e->dstage_bytes) {
      CU(e, cudaDeviceSynchronize());
      if (e->dstage) CU(e, cudaFree(e->dstage));
      e->dstage_bytes = need * 2;
      CU(e, cudaMalloc((void**)&e->dstage, e->dstage_bytes));
    }
    CU(e, cudaMemcpyAsync(e->dstage, h_ntok, (size_t)n_win * 4, cudaMemcpyHostToDevice, st));
    d_ntok = reinterpret_cast<const int32_t*>(e->dstage);
  } else if (ntok_uniform < 0 || ntok_uniform > tokens_stride) {
    return fail(e, SNACB_EINVAL, "snacb_deinterleave: ntok_uniform out of range");
  }
  {
    ProfScope ps(e, KC_DEINT, 0.0, 8.0 * (double)n_win * 7 * max_frames, st);
    launch_deinterleave(d_tokens, tokens_stride, d_ntok, ntok_uniform, n_win, max_frames, raw, d_c0, d_c1, d_c2,
                        d_status, st, &e->launches);
  }
  return check_launch(e, "deinterleave");
}

int snacb_deinterleave(snacb_engine* e, const int32_t* d_tokens, int32_t tokens_stride, const int32_t* h_ntok,
                       int32_t ntok_uniform, int32_t n_win, int32_t max_frames, int32_t* d_c0, int32_t* d_c1,
                       int32_t* d_c2, int32_t* d_status, void* stream) {
  return deinterleave_common(e, d_tokens, tokens_stride, h_ntok, ntok_uniform, n_win, max_frames, false, d_c0, d_c1,
                             d_c2, d_status, stream);
}

int snacb_deinterleave_raw(snacb_engine* e, const int32_t* d_raw, int32_t tokens_stride, const int32_t* h_ntok,
                           int32_t ntok_uniform, int32_t n_win, int32_t max_frames, int32_t* d_c0, int32_t* d_c1,
                           int32_t* d_c2, int32_t* d_status, void* stream) {
  return deinterleave_common(e, d_raw, tokens_stride, h_ntok, ntok_uniform, n_win, max_frames, true, d_c0, d_c1, d_c2,
                             d_status, stream);
}

// d_keys_in / d_seed_in: keys and seed already on the device (CUDA-graph path); else h_keys / seed are used.
static int decode_windows_core(snacb_engine* e, const int32_t* d_tokens, int32_t tokens_stride, const int32_t* h_ntok,
                               int32_t ntok_uniform, int32_t n_win, int32_t noise_mode, const float* d_noise,
                               int64_t noise_stride, uint64_t seed, const uint64_t* h_keys,
                               const unsigned long long* d_keys_in, const unsigned long long* d_seed_in, int16_t* d_pcm,
                               int32_t* d_status, void* stream) {
  if (!e) return SNACB_EINVAL;
  if (!e->loaded) return fail(e, SNACB_ESTATE, "snacb_decode_windows: weights not loaded");
  if (n_win < 0 || !d_tokens || !d_pcm || !d_status) return fail(e, SNACB_EINVAL, "snacb_decode_windows: bad argument");
  if (noise_mode < 0 || noise_mode > 2 || (noise_mode == SNACB_NOISE_TENSOR && !d_noise))
    return fail(e, SNACB_EINVAL, "snacb_decode_windows: bad noise arguments");
  if (n_win == 0) return SNACB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  CU(e, cudaSetDevice(e->device));

  // frame counts per window (host metadata)
  int maxF = 1;
  std::map<int, std::vector<int>> groups;  // F -> windows
  if (h_ntok) {
    for (int i = 0; i < n_win; ++i) {
      if (h_ntok[i] < 0 || h_ntok[i] > tokens_stride) return fail(e, SNACB_EINVAL, "ntok[%d] out of range", i);
      const int F = h_ntok[i] / 7;
      maxF = std::max(maxF, F);
      if (F >= 2) groups[F].push_back(i);
    }
  } else {
    if (ntok_uniform < 0 || ntok_uniform > tokens_stride) return fail(e, SNACB_EINVAL, "ntok_uniform out of range");
    maxF = std::max(1, ntok_uniform / 7);
  }
  if (noise_mode == SNACB_NOISE_TENSOR && noise_stride < (int64_t)kNoisePerFrame * maxF)
    return fail(e, SNACB_EINVAL, "noise_stride %lld < 3360*%d", (long long)noise_stride, maxF);

  // fixed part of the workspace: codes, ntok, keys, item tables
  const size_t codes_b = pad256((size_t)n_win * maxF * 4) + pad256((size_t)n_win * maxF * 8) + pad256((size_t)n_win * maxF * 16);
  const size_t fixed = codes_b + 3 * pad256((size_t)n_win * 16) + 4096;
  // activation part: sized for the largest group plan
  size_t act = 0;
  std::vector<int> Fs;
  if (h_ntok) for (auto& kv : groups) Fs.push_back(kv.first); else if (maxF >= 2) Fs.push_back(maxF);
  const Rng slice{2048, 4096};
  for (int F : Fs) {
    Plan P = plan_for(e, 4 * F, e->cfg.trim ? slice : Rng{0, 2048 * F}, true);
    const int cnt = h_ntok ? (int)groups[F].size() : n_win;
    act = std::max(act, act_bytes(e, P, cnt));
  }
  int rc = ensure_ws(e, fixed + act, st);
  if (rc) return rc;
  Bump bp(e->ws);
  int32_t* c0 = bp.take<int32_t>((size_t)n_win * maxF);
  int32_t* c1 = bp.take<int32_t>((size_t)n_win * maxF * 2);
  int32_t* c2 = bp.take<int32_t>((size_t)n_win * maxF * 4);
  int32_t* d_ntok = bp.take<int32_t>(n_win);
  unsigned long long* d_keys = bp.take<unsigned long long>(n_win);
  Item* d_items = bp.take<Item>(n_win);
  char* ws_free = e->ws + pad256(bp.off);
  const size_t ws_avail = e->ws_bytes - pad256(bp.off);

  if (h_ntok) CU(e, cudaMemcpyAsync(d_ntok, h_ntok, (size_t)n_win * 4, cudaMemcpyHostToDevice, st));
  const unsigned long long* keys = d_keys_in;
  if (noise_mode == SNACB_NOISE_PHILOX && h_keys && !d_keys_in) {
    CU(e, cudaMemcpyAsync(d_keys, h_keys, (size_t)n_win * 8, cudaMemcpyHostToDevice, st));
    keys = d_keys;
  }
  {
    ProfScope ps(e, KC_DEINT, 0.0, 8.0 * (double)n_win * 7 * maxF, st);
    launch_deinterleave(d_tokens, tokens_stride, h_ntok ? d_ntok : nullptr, ntok_uniform, n_win, maxF, false, c0, c1, c2,
                        d_status, st, &e->launches);
  }
  CU(e, cudaMemsetAsync(d_pcm, 0, (size_t)n_win * 2048 * sizeof(int16_t), st));
  NoiseCfg nz{noise_mode, d_noise, (long long)noise_stride, seed, keys, d_seed_in};

  if (!h_ntok) {
    if (maxF >= 2 && ntok_uniform >= 14) {
      Plan P = plan_for(e, 4 * maxF, e->cfg.trim ? slice : Rng{0, 2048 * maxF}, true);
      rc = run_group(e, P, nullptr, n_win, 2048, slice, c0, c1, c2, maxF, nz, d_status, nullptr, d_pcm, ws_free, ws_avail, st);
      if (rc) return rc;
    }
  } else {
    std::vector<Item> all;
    std::vector<std::pair<int, std::pair<size_t, int>>> runs;  // F, (offset, count)
    for (auto& kv : groups) {
      runs.push_back({kv.first, {all.size(), (int)kv.second.size()}});
      for (int i : kv.second) all.push_back(Item{i, 0, (int64_t)i * 2048});
    }
    if (!all.empty()) {
      rc = upload_items(e, all, d_items, st);
      if (rc) return rc;
      for (auto& r : runs) {
        Plan P = plan_for(e, 4 * r.first, e->cfg.trim ? slice : Rng{0, 2048 * r.first}, true);
        rc = run_group(e, P, d_items + r.second.first, r.second.second, 2048, slice, c0, c1, c2, maxF, nz, d_status,
                           nullptr, d_pcm, ws_free, ws_avail, st);
        if (rc) return rc;
      }
    }
  }
  return check_launch(e, "decode_windows");
}

int snacb_decode_windows(snacb_engine* e, const int32_t* d_tokens, int32_t tokens_stride, const int32_t* h_ntok,
                         int32_t ntok_uniform, int32_t n_win, int32_t noise_mode, const float* d_noise,
                         int64_t noise_stride, uint64_t seed, const uint64_t* h_keys, int16_t* d_pcm,
                         int32_t* d_status, void* stream) {
  return decode_windows_core(e, d_tokens, tokens_stride, h_ntok, ntok_uniform, n_win, noise_mode, d_noise, noise_stride, seed,
                             h_keys, nullptr, nullptr, d_pcm, d_status, stream);
}

static bool is_pinned(const void* p) {
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

int snacb_decode_windows_host(snacb_engine* e, const int32_t* h_tokens, int32_t tokens_stride, const int32_t* h_ntok,
                              int32_t ntok_uniform, int32_t n_win, int32_t noise_mode, const float* h_noise,
                              int64_t noise_stride, uint64_t seed, const uint64_t* h_keys, int16_t* h_pcm,
                              int32_t* h_status, void* stream) {
  if (!e) return SNACB_EINVAL;
  if (n_win < 0 || !h_tokens || !h_pcm || !h_status) return fail(e, SNACB_EINVAL, "snacb_decode_windows_host: bad argument");
  if (n_win == 0) return SNACB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  CU(e, cudaSetDevice(e->device));
  const size_t tok_n = (size_t)n_win * tokens_stride * 4;
  const size_t tok_b = pad256(tok_n);
  const size_t pcm_b = pad256((size_t)n_win * 4096);
  const size_t st_b = pad256((size_t)n_win * 4);
  const size_t nz_b = (noise_mode == SNACB_NOISE_TENSOR && h_noise) ? pad256((size_t)n_win * noise_stride * 4) : 0;
  const size_t key_b = pad256((size_t)n_win * 8), seed_b = 256;
  const size_t total = tok_b + pcm_b + st_b + nz_b + key_b + seed_b;
  if (total > e->pin_bytes) {
    CU(e, cudaStreamSynchronize(st));
    if (e->pin) CU(e, cudaFreeHost(e->pin));
    e->pin_bytes = total + total / 2;
    CU(e, cudaMallocHost((void**)&e->pin, e->pin_bytes));
    ++e->gen;
  }
  // device staging lives in its own allocation (the workspace may be re-allocated by the decode)
  if (total > e->dstage_bytes) {
    CU(e, cudaDeviceSynchronize());
    if (e->dstage) CU(e, cudaFree(e->dstage));
    e->dstage_bytes = total + total / 2;
    CU(e, cudaMalloc((void**)&e->dstage, e->dstage_bytes));
    ++e->gen;
  }
  char* hp = e->pin; char* dp = e->dstage;
  int32_t* d_tok = reinterpret_cast<int32_t*>(dp);
  int16_t* d_pcm = reinterpret_cast<int16_t*>(dp + tok_b);
  int32_t* d_st = reinterpret_cast<int32_t*>(dp + tok_b + pcm_b);
  float* d_nz = nz_b ? reinterpret_cast<float*>(dp + tok_b + pcm_b + st_b) : nullptr;
  const size_t key_off = tok_b + pcm_b + st_b + nz_b, seed_off = key_off + key_b;

  // ---- latency mode: small uniform ticks replay a captured CUDA graph (H2D + every kernel + D2H in one launch)
  const bool graph_ok = e->graph_max_win > 0 && n_win <= e->graph_max_win && !h_ntok && noise_mode != SNACB_NOISE_TENSOR &&
                        !e->prof.on && e->tap_stage < 0 && e->loaded && ntok_uniform >= 0 && ntok_uniform <= tokens_stride;
  if (graph_ok) {
    const std::vector<long long> key{n_win, tokens_stride, ntok_uniform, noise_mode, h_keys ? 1 : 0};
    snacb_engine::GraphEntry& ge = e->graphs[key];
    if (ge.exec && ge.gen != e->gen) { cudaGraphExecDestroy(ge.exec); ge.exec = nullptr; ge.calls = 0; }
    if (!ge.disabled && ge.calls >= 1) {
      memcpy(hp, h_tokens, tok_n);
      if (h_keys) memcpy(hp + key_off, h_keys, (size_t)n_win * 8);
      memcpy(hp + seed_off, &seed, 8);
      if (!ge.exec) {
        if (!e->cap_stream) CU(e, cudaStreamCreateWithFlags(&e->cap_stream, cudaStreamNonBlocking));
        cudaStream_t cs = e->cap_stream;
        const uint64_t gen0 = e->gen;
        cudaGraph_t graph = nullptr;
        int rc = SNACB_OK;
        if (cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); rc = SNACB_ECUDA; }
        if (rc == SNACB_OK) {
          cudaMemcpyAsync(d_tok, hp, tok_n, cudaMemcpyHostToDevice, cs);
          if (h_keys) cudaMemcpyAsync(dp + key_off, hp + key_off, (size_t)n_win * 8, cudaMemcpyHostToDevice, cs);
          cudaMemcpyAsync(dp + seed_off, hp + seed_off, 8, cudaMemcpyHostToDevice, cs);
          rc = decode_windows_core(e, d_tok, tokens_stride, nullptr, ntok_uniform, n_win, noise_mode, nullptr, 0, seed, nullptr,
                                   h_keys ? reinterpret_cast<const unsigned long long*>(dp + key_off) : nullptr,
                                   reinterpret_cast<const unsigned long long*>(dp + seed_off), d_pcm, d_st, cs);
          cudaMemcpyAsync(hp + tok_b, d_pcm, (size_t)n_win * 4096, cudaMemcpyDeviceToHost, cs);
          cudaMemcpyAsync(hp + tok_b + pcm_b, d_st, (size_t)n_win * 4, cudaMemcpyDeviceToHost, cs);
          const cudaError_t ce = cudaStreamEndCapture(cs, &graph);
          if (ce != cudaSuccess || !graph) { cudaGetLastError(); rc = rc ? rc : SNACB_ECUDA; }
        }
        if (rc == SNACB_OK && e->gen == gen0 && cudaGraphInstantiate(&ge.exec, graph, 0) == cudaSuccess) {
          ge.gen = e->gen;
        } else {
          cudaGetLastError();
          ge.exec = nullptr; ge.disabled = true;  // this shape falls back to the plain path for good
        }
        if (graph) cudaGraphDestroy(graph);
      }
      if (ge.exec) {
        ++ge.calls;
        CU(e, cudaGraphLaunch(ge.exec, st));
        ++e->graph_launches;
        CU(e, cudaStreamSynchronize(st));
        memcpy(h_pcm, hp + tok_b, (size_t)n_win * 4096);
        memcpy(h_status, hp + tok_b + pcm_b, (size_t)n_win * 4);
        return SNACB_OK;
      }
    }
    ++ge.calls;  // first call of a shape runs the plain path (sizes the workspace, sets kernel attributes)
  }

  const bool tok_pinned = is_pinned(h_tokens), pcm_pinned = is_pinned(h_pcm);
  const void* src_tok = h_tokens;
  if (!tok_pinned) { memcpy(hp, h_tokens, tok_n); src_tok = hp; }
  CU(e, cudaMemcpyAsync(d_tok, src_tok, tok_n, cudaMemcpyHostToDevice, st));
  if (nz_b) {
    memcpy(hp + tok_b + pcm_b + st_b, h_noise, (size_t)n_win * noise_stride * 4);
    CU(e, cudaMemcpyAsync(d_nz, hp + tok_b + pcm_b + st_b, (size_t)n_win * noise_stride * 4, cudaMemcpyHostToDevice, st));
  }
  int rc = snacb_decode_windows(e, d_tok, tokens_stride, h_ntok, ntok_uniform, n_win, noise_mode, d_nz, noise_stride, seed,
                                h_keys, d_pcm, d_st, stream);
  if (rc) return rc;
  void* dst_pcm = pcm_pinned ? (void*)h_pcm : (void*)(hp + tok_b);
  CU(e, cudaMemcpyAsync(dst_pcm, d_pcm, (size_t)n_win * 4096, cudaMemcpyDeviceToHost, st));
  CU(e, cudaMemcpyAsync(hp + tok_b + pcm_b, d_st, (size_t)n_win * 4, cudaMemcpyDeviceToHost, st));
  CU(e, cudaStreamSynchronize(st));
  if (!pcm_pinned) memcpy(h_pcm, hp + tok_b, (size_t)n_win * 4096);
  memcpy(h_status, hp + tok_b + pcm_b, (size_t)n_win * 4);
  return SNACB_OK;
}

// N3 on the GPU: the tick's PCM goes from the decoder tail straight into the pinned per-stream rings of `g` (the ring
// kernel crossfades consecutive chunks of a stream when an overlap is configured); only the statuses come back as a copy.
int snacb_decode_windows_to_ring(snacb_engine* e, snacb_egress* g, const int32_t* h_tokens, int32_t tokens_stride,
                                 const int32_t* h_ntok, int32_t ntok_uniform, int32_t n_win, int32_t noise_mode, uint64_t seed,
                                 const uint64_t* h_keys, const int32_t* h_slots, const int32_t* h_eos, int32_t* h_status,
                                 int32_t* h_emitted, void* stream) {
  if (!e) return SNACB_EINVAL;
  if (!g || n_win < 0 || !h_tokens || !h_slots || !h_status || noise_mode == SNACB_NOISE_TENSOR)
    return fail(e, SNACB_EINVAL, "snacb_decode_windows_to_ring: bad argument (noise off / philox only)");
  if (n_win == 0) return SNACB_OK;
  std::vector<int32_t> slots(h_slots, h_slots + n_win);
  std::vector<int64_t> before((size_t)n_win, 0);
  for (int i = 0; i < n_win; ++i) {
    if (slots[i] < 0) continue;
    const int64_t room = snacb_egress_room(g, slots[i]);
    if (room < 0) return fail(e, SNACB_EINVAL, "snacb_decode_windows_to_ring: slot %d out of range", slots[i]);
    before[(size_t)i] = snacb_egress_written(g, slots[i]);
    if (room < 2048) { slots[i] = -1; before[(size_t)i] = -1; }  // a stalled consumer: this window goes nowhere (h_emitted = -1)
  }
  cudaStream_t st = (cudaStream_t)stream;
  CU(e, cudaSetDevice(e->device));
  const size_t tok_n = (size_t)n_win * tokens_stride * 4;
  const size_t tok_b = pad256(tok_n), pcm_b = pad256((size_t)n_win * 4096), st_b = pad256((size_t)n_win * 4);
  const size_t total = tok_b + pcm_b + st_b + pad256((size_t)n_win * 8) + 256;  // same block layout as the host call
  if (total > e->pin_bytes) {
    CU(e, cudaStreamSynchronize(st));
    if (e->pin) CU(e, cudaFreeHost(e->pin));
    e->pin_bytes = total + total / 2;
    CU(e, cudaMallocHost((void**)&e->pin, e->pin_bytes));
    ++e->gen;
  }
  if (total > e->dstage_bytes) {
    CU(e, cudaDeviceSynchronize());
    if (e->dstage) CU(e, cudaFree(e->dstage));
    e->dstage_bytes = total + total / 2;
    CU(e, cudaMalloc((void**)&e->dstage, e->dstage_bytes));
    ++e->gen;
  }
  char* hp = e->pin; char* dp = e->dstage;
  int32_t* d_tok = reinterpret_cast<int32_t*>(dp);
  int16_t* d_pcm = reinterpret_cast<int16_t*>(dp + tok_b);
  int32_t* d_st = reinterpret_cast<int32_t*>(dp + tok_b + pcm_b);
  const void* src_tok = h_tokens;
  if (!is_pinned(h_tokens)) { memcpy(hp, h_tokens, tok_n); src_tok = hp; }
  CU(e, cudaMemcpyAsync(d_tok, src_tok, tok_n, cudaMemcpyHostToDevice, st));
  int rc = snacb_decode_windows(e, d_tok, tokens_stride, h_ntok, ntok_uniform, n_win, noise_mode, nullptr, 0, seed, h_keys, d_pcm,
                                d_st, stream);
  if (rc) return rc;
  rc = snacb_egress_push_device(g, n_win, slots.data(), d_pcm, 2048, 2048, d_st, h_eos, stream);
  if (rc) return fail(e, rc, "snacb_decode_windows_to_ring: %s", snacb_egress_last_error(g));
  ++e->launches;
  CU(e, cudaMemcpyAsync(hp + tok_b + pcm_b, d_st, (size_t)n_win * 4, cudaMemcpyDeviceToHost, st));
  rc = snacb_egress_sync(g, stream);
  if (rc) return fail(e, rc, "snacb_decode_windows_to_ring: %s", snacb_egress_last_error(g));
  memcpy(h_status, hp + tok_b + pcm_b, (size_t)n_win * 4);
  if (h_emitted)
    for (int i = 0; i < n_win; ++i)
      h_emitted[i] = before[(size_t)i] < 0 ? -1 : (slots[i] < 0 ? 0 : (int32_t)(snacb_egress_written(g, slots[i]) - before[(size_t)i]));
  return SNACB_OK;
}

int snacb_decode_windows_host_submit(snacb_engine* e, const int32_t* h_tokens, int32_t tokens_stride, const int32_t* h_ntok,
                                     int32_t ntok_uniform, int32_t n_win, int32_t noise_mode, uint64_t seed,
                                     const uint64_t* h_keys, int16_t* h_pcm, int32_t* h_status, void* stream, int32_t* ticket) {
  if (!e) return SNACB_EINVAL;
  if (n_win <= 0 || !h_tokens || !h_pcm || !h_status || !ticket || noise_mode == SNACB_NOISE_TENSOR)
    return fail(e, SNACB_EINVAL, "snacb_decode_windows_host_submit: bad argument (n_win > 0, noise off / philox only)");
  cudaStream_t st = (cudaStream_t)stream;
  CU(e, cudaSetDevice(e->device));
  const int si = e->pipe_next;
  snacb_engine::PipeSlot& sl = e->pipe[si];
  if (sl.busy) return fail(e, SNACB_ESTATE, "snacb_decode_windows_host_submit: two ticks already in flight, wait for the older one first");
  if (!e->copy_stream) CU(e, cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
  if (!sl.tail_done) {
    CU(e, cudaEventCreateWithFlags(&sl.tail_done, cudaEventDisableTiming));
    CU(e, cudaEventCreateWithFlags(&sl.copied, cudaEventDisableTiming));
  }
  const size_t tok_n = (size_t)n_win * tokens_stride * 4;
  const size_t tok_b = pad256(tok_n), pcm_b = pad256((size_t)n_win * 4096), st_b = pad256((size_t)n_win * 4);
  const size_t key_b = pad256((size_t)n_win * 8), ntok_b = pad256((size_t)n_win * 4);
  const size_t total = tok_b + pcm_b + st_b + key_b + ntok_b;
  if (total > sl.pin_bytes) {
    if (sl.pin) CU(e, cudaFreeHost(sl.pin));
    sl.pin_bytes = total + total / 2;
    CU(e, cudaMallocHost((void**)&sl.pin, sl.pin_bytes));
  }
  if (total > sl.dst_bytes) {
    CU(e, cudaDeviceSynchronize());
    if (sl.dst) CU(e, cudaFree(sl.dst));
    sl.dst_bytes = total + total / 2;
    CU(e, cudaMalloc((void**)&sl.dst, sl.dst_bytes));
  }
  char* hp = sl.pin; char* dp = sl.dst;
  int32_t* d_tok = reinterpret_cast<int32_t*>(dp);
  int16_t* d_pcm = reinterpret_cast<int16_t*>(dp + tok_b);
  int32_t* d_st = reinterpret_cast<int32_t*>(dp + tok_b + pcm_b);
  const size_t key_off = tok_b + pcm_b + st_b, ntok_off = key_off + key_b;
  // inputs are copied into the slot's pinned block: the caller's arrays may be re-used as soon as this returns
  memcpy(hp, h_tokens, tok_n);
  const uint64_t* keys_p = nullptr;
  if (h_keys) { memcpy(hp + key_off, h_keys, (size_t)n_win * 8); keys_p = reinterpret_cast<const uint64_t*>(hp + key_off); }
  const int32_t* ntok_p = nullptr;
  if (h_ntok) { memcpy(hp + ntok_off, h_ntok, (size_t)n_win * 4); ntok_p = reinterpret_cast<const int32_t*>(hp + ntok_off); }
  CU(e, cudaMemcpyAsync(d_tok, hp, tok_n, cudaMemcpyHostToDevice, st));
  int rc = snacb_decode_windows(e, d_tok, tokens_stride, ntok_p, ntok_uniform, n_win, noise_mode, nullptr, 0, seed, keys_p, d_pcm,
                                d_st, stream);
  if (rc) return rc;
  CU(e, cudaEventRecord(sl.tail_done, st));
  CU(e, cudaStreamWaitEvent(e->copy_stream, sl.tail_done, 0));
  sl.direct = is_pinned(h_pcm) && is_pinned(h_status);
  sl.pcm_off = tok_b; sl.st_off = tok_b + pcm_b;
  CU(e, cudaMemcpyAsync(sl.direct ? (void*)h_pcm : (void*)(hp + sl.pcm_off), d_pcm, (size_t)n_win * 4096, cudaMemcpyDeviceToHost, e->copy_stream));
  CU(e, cudaMemcpyAsync(sl.direct ? (void*)h_status : (void*)(hp + sl.st_off), d_st, (size_t)n_win * 4, cudaMemcpyDeviceToHost, e->copy_stream));
  CU(e, cudaEventRecord(sl.copied, e->copy_stream));
  sl.busy = true; sl.n_win = n_win; sl.h_pcm = h_pcm; sl.h_status = h_status;
  *ticket = si;
  e->pipe_next = si ^ 1;
  return SNACB_OK;
}

int snacb_decode_windows_host_wait(snacb_engine* e, int32_t ticket) {
  if (!e) return SNACB_EINVAL;
  if (ticket < 0 || ticket > 1 || !e->pipe[ticket].busy) return fail(e, SNACB_EINVAL, "snacb_decode_windows_host_wait: no tick in flight under this ticket");
  snacb_engine::PipeSlot& sl = e->pipe[ticket];
  CU(e, cudaSetDevice(e->device));
  CU(e, cudaEventSynchronize(sl.copied));
  if (!sl.direct) {
    memcpy(sl.h_pcm, sl.pin + sl.pcm_off, (size_t)sl.n_win * 4096);
    memcpy(sl.h_status, sl.pin + sl.st_off, (size_t)sl.n_win * 4);
  }
  sl.busy = false;
  return SNACB_OK;
}

int snacb_decode_codes(snacb_engine* e, const int32_t* d_c0, const int32_t* d_c1, const int32_t* d_c2, int32_t B,
                       int32_t F, int32_t noise_mode, const float* d_noise, uint64_t seed, float* d_wav, int16_t* d_pcm,
                       void* stream) {
  if (!e) return SNACB_EINVAL;
  if (!e->loaded) return fail(e, SNACB_ESTATE, "snacb_decode_codes: weights not loaded");
  if (B < 0 || F < 1 || !d_c0 || !d_c1 || !d_c2 || (!d_wav && !d_pcm)) return fail(e, SNACB_EINVAL, "snacb_decode_codes: bad argument");
  if (noise_mode < 0 || noise_mode > 2 || (noise_mode == SNACB_NOISE_TENSOR && !d_noise))
    return fail(e, SNACB_EINVAL, "snacb_decode_codes: bad noise arguments");
  if (B == 0) return SNACB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  CU(e, cudaSetDevice(e->device));
  NoiseCfg nz{noise_mode, d_noise, (long long)kNoisePerFrame * F, seed, nullptr, nullptr};
  const int kTileFrames = 8;
  if (F <= 2 * kTileFrames) {
    Plan P = plan_for(e, 4 * F, Rng{0, 2048 * F}, true);
    int rc = ensure_ws(e, act_bytes(e, P, B), st);
    if (rc) return rc;
    return run_group(e, P, nullptr, B, 2048 * F, P.out, d_c0, d_c1, d_c2, F, nz, nullptr, d_wav, d_pcm, e->ws, e->ws_bytes, st);
  }
  // long sequence: uniform time tiles with halo recompute; rows outside [0, T) are explicit zeros.
  const int tiles = (F + kTileFrames - 1) / kTileFrames;
  std::vector<Item> items;
  items.reserve((size_t)B * tiles);
  for (int b = 0; b < B; ++b)
    for (int k = 0; k < tiles; ++k)
      items.push_back(Item{b, k * kTileFrames * 4, (int64_t)b * 2048 * F + (int64_t)k * kTileFrames * 2048});
  Plan P = plan_for(e, 4 * F, Rng{0, 2048 * kTileFrames}, false);
  const size_t tab = pad256(items.size() * sizeof(Item)) + 256;
  int rc = ensure_ws(e, tab + act_bytes(e, P, (int)items.size()), st);
  if (rc) return rc;
  Item* d_items = reinterpret_cast<Item*>(e->ws);
  rc = upload_items(e, items, d_items, st);
  if (rc) return rc;
  return run_group(e, P, d_items, (int)items.size(), 2048 * kTileFrames, P.out, d_c0, d_c1, d_c2, F, nz, nullptr, d_wav,
                       d_pcm, e->ws + tab, e->ws_bytes - tab, st);
}

int snacb_fill_noise(snacb_engine* e, uint64_t seed, const uint64_t* h_keys, int32_t n_win, int32_t F, float* d_noise,
                     int64_t noise_stride, void* stream) {
  if (!e) return SNACB_EINVAL;
  if (n_win < 0 || F < 1 || !d_noise || noise_stride < (int64_t)kNoisePerFrame * F) return fail(e, SNACB_EINVAL, "snacb_fill_noise: bad argument");
  if (n_win == 0) return SNACB_OK;
  if (n_win > 65535) return fail(e, SNACB_EINVAL, "snacb_fill_noise: at most 65535 windows per call");
  cudaStream_t st = (cudaStream_t)stream;
  CU(e, cudaSetDevice(e->device));
  const unsigned long long* keys = nullptr;
  if (h_keys) {
    int rc = ensure_ws(e, pad256((size_t)n_win * 8), st);
    if (rc) return rc;
    CU(e, cudaMemcpyAsync(e->ws, h_keys, (size_t)n_win * 8, cudaMemcpyHostToDevice, st));
    keys = reinterpret_cast<const unsigned long long*>(e->ws);
  }
  launch_fill_noise(seed, keys, n_win, F, d_noise, (long long)noise_stride, st, &e->launches);
  return check_launch(e, "fill_noise");
}

int snacb_profile_enable(snacb_engine* e, int32_t on) {
  if (!e) return SNACB_EINVAL;
  Prof& p = e->prof;
  cudaSetDevice(e->device);
  for (auto& r : p.recs) { cudaEventSynchronize(r.b); p.pool.push_back(r.a); p.pool.push_back(r.b); }
  p.recs.clear();
  for (int c = 0; c < KC_COUNT; ++c) { p.ms[c] = p.flops[c] = p.bytes[c] = 0.0; p.launches[c] = 0; }
  p.on = on != 0;
  return SNACB_OK;
}

int snacb_profile_read(snacb_engine* e, snacb_kernel_stat* out, int32_t cap) {
  if (!e || (cap > 0 && !out)) return SNACB_EINVAL;
  Prof& p = e->prof;
  CU(e, cudaSetDevice(e->device));
  for (auto& r : p.recs) {
    CU(e, cudaEventSynchronize(r.b));
    float ms = 0.f;
    CU(e, cudaEventElapsedTime(&ms, r.a, r.b));
    p.ms[r.cls] += ms;
    p.pool.push_back(r.a); p.pool.push_back(r.b);
  }
  p.recs.clear();
  for (int c = 0; c < KC_COUNT && c < cap; ++c) {
    memset(&out[c], 0, sizeof out[c]);
    snprintf(out[c].name, sizeof out[c].name, "%s", kClassName[c]);
    out[c].launches = p.launches[c]; out[c].ms = p.ms[c]; out[c].flops = p.flops[c]; out[c].bytes = p.bytes[c];
  }
  return KC_COUNT;
}

int snacb_set_tap(snacb_engine* e, int32_t stage, float* d_buf, size_t capacity_floats) {
  if (!e) return SNACB_EINVAL;
  e->tap_stage = stage; e->tap_buf = d_buf; e->tap_cap = capacity_floats;
  e->tap_rows = e->tap_ch = e->tap_lo = e->tap_items = 0;
  return SNACB_OK;
}

int snacb_get_tap_shape(const snacb_engine* e, int32_t* rows, int32_t* channels, int32_t* t_lo, int32_t* items) {
  if (!e) return SNACB_EINVAL;
  if (rows) *rows = e->tap_rows;
  if (channels) *channels = e->tap_ch;
  if (t_lo) *t_lo = e->tap_lo;
  if (items) *items = e->tap_items;
  return SNACB_OK;
}

int snacb_plan(int32_t frames, int32_t out_lo, int32_t out_hi, int32_t clip, int32_t* ranges) {
  if (frames < 1 || out_hi <= out_lo || !ranges) return SNACB_EINVAL;
  const Plan P = make_plan(4 * frames, Rng{out_lo, out_hi}, clip != 0);
  int k = 0;
  auto put = [&](Rng r) { ranges[k++] = r.lo; ranges[k++] = r.hi; };
  put(P.z); put(P.h);
  for (int b = 0; b < 4; ++b) {
    put(P.b[b].in); put(P.b[b].q); put(P.b[b].ct);
    put(P.b[b].r[0]); put(P.b[b].r[1]); put(P.b[b].r[2]);
  }
  return SNACB_OK;
}

}  // extern "C"
