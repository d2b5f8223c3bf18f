// Shared declarations of the SNAC-24k decode engine (host plan + device helpers).
//
// Data layout in HBM: every activation is CHANNELS-LAST fp32, [item][row][C], where an item is one
// window (streaming) or one time tile of a long sequence, and row r holds relative time
// t_rel = lo + r of that stage.  Absolute time is t_abs = t_rel + shift0 * up(stage); rows whose
// t_abs falls outside [0, T_stage) are stored as ZERO by the kernel that produces them, so they are
// exactly the zero padding the reference's convolutions see and consumers never bounds-check time.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <string>

namespace snacb {

constexpr int kLatent = 768;
constexpr int kDecDim = 1024;
constexpr int kRates[4] = {8, 8, 4, 2};
constexpr int kDil[3] = {1, 3, 9};
constexpr int kNoisePerFrame = 3360;

// One unit of work of a uniform group.
struct Item {
  int32_t code_row;  // row of the code / noise / status arrays this item reads
  int32_t shift0;    // origin of the item's relative time axis, in latent steps
  int64_t dst;       // first output sample's index in the wav / pcm destination
};

struct Rng {
  int lo, hi;
  __host__ __device__ int n() const { return hi - lo; }
};

struct BlockPlan {
  int Cin, Cout, s, p;
  int up_in, up_out;  // cumulative upsampling of the block's input / output time axis
  Rng in;             // rows of the block input (Snake'd operand of the transposed conv)
  Rng q;              // input positions q iterated by the polyphase GEMM
  Rng ct;             // transposed-conv / noise output rows
  Rng r[3];           // output rows of the three residual units; r[2] is the block output
};

struct Plan {
  int T0;        // latent steps of the whole sequence (4 per frame): time bounds are T0*up
  Rng z, h;      // latent rows; head (dw7 + 1x1) output rows == block 0 input rows
  BlockPlan b[4];
  Rng out;       // waveform samples produced per item (relative)
  size_t max_stage_floats;  // largest [rows][C] activation of one item
};

inline int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

// Back-propagate the needed row ranges from the requested output samples.
// clip: intersect every range with [0, T_stage) - valid when every item of the group has shift0 = 0.
// halo_in: keep each block's input range unclipped, so the rows q-1 / q+1 the transposed conv reads
// at the sequence edges exist as explicit zero rows (needed by the TMA-fed tensor-core path, whose
// shifted operand tile is a plain row offset into the flattened [items x rows] operand).
inline Plan make_plan(int T0, Rng out, bool clip, bool halo_in = false) {
  Plan P{};
  P.T0 = T0;
  P.out = out;
  auto clipr = [&](Rng r, int up) {
    if (clip) { r.lo = std::max(r.lo, 0); r.hi = std::min(r.hi, T0 * up); }
    return r;
  };
  int cin = kDecDim, up = 1;
  for (int b = 0; b < 4; ++b) {
    BlockPlan& B = P.b[b];
    B.Cin = cin; B.Cout = cin / 2; B.s = kRates[b]; B.p = (B.s + 1) / 2;
    B.up_in = up; B.up_out = up * B.s;
    cin /= 2; up *= B.s;
  }
  Rng need = clipr(Rng{out.lo - 3, out.hi + 3}, 512);  // tail conv k7 pad 3
  for (int b = 3; b >= 0; --b) {
    BlockPlan& B = P.b[b];
    B.r[2] = need;
    B.r[1] = clipr(Rng{B.r[2].lo - 27, B.r[2].hi + 27}, B.up_out);
    B.r[0] = clipr(Rng{B.r[1].lo - 9, B.r[1].hi + 9}, B.up_out);
    B.ct = clipr(Rng{B.r[0].lo - 3, B.r[0].hi + 3}, B.up_out);
    int q0 = floordiv(B.ct.lo, B.s), r0 = B.ct.lo - q0 * B.s;
    int q1 = floordiv(B.ct.hi - 1, B.s), r1 = (B.ct.hi - 1) - q1 * B.s;
    B.q = Rng{q0, q1 + 1};
    B.in = Rng{(r0 < B.s - B.p) ? q0 - 1 : q0, ((r1 >= B.s - B.p) ? q1 + 1 : q1) + 1};
    if (!halo_in) B.in = clipr(B.in, B.up_in);
    B.q = clipr(B.q, B.up_in);
    need = B.in;
  }
  P.h = need;
  P.z = clipr(Rng{P.h.lo - 3, P.h.hi + 3}, 1);
  size_t mx = std::max((size_t)P.z.n() * kLatent, (size_t)P.h.n() * kDecDim);
  for (int b = 0; b < 4; ++b) {
    mx = std::max(mx, (size_t)P.b[b].in.n() * P.b[b].Cin);
    mx = std::max(mx, (size_t)P.b[b].ct.n() * P.b[b].Cout);
  }
  P.max_stage_floats = mx;
  return P;
}

// ------------------------------------------------------------------------------------ device side
struct ItemRef {
  int code_row;
  int shift0;
  long long dst;
};

// items == nullptr: implicit uniform items (row = base + i, no shift, dense destination).
__device__ __forceinline__ ItemRef get_item(const Item* items, int base, int i, int out_len) {
  if (items) {
    Item it = items[i];
    return ItemRef{it.code_row, it.shift0, it.dst};
  }
  return ItemRef{base + i, 0, (long long)(base + i) * out_len};
}

__device__ __forceinline__ float snake_exact(float x, float alpha, float inv) {
  float s = sinf(alpha * x);
  return x + inv * (s * s);
}

// Philox4x32-10, one call per noise value: counter = (t_abs, block, key.lo, key.hi), key = seed.
__device__ __forceinline__ float philox_normal(uint64_t seed, uint64_t key, uint32_t block, uint32_t t) {
  uint32_t c0 = t, c1 = block, c2 = (uint32_t)key, c3 = (uint32_t)(key >> 32);
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  // Box-Muller on two 32-bit uniforms in (0,1]
  float u1 = ((float)c0 + 1.0f) * 2.3283064365386963e-10f;
  float u2 = (float)c1 * 2.3283064365386963e-10f;
  u1 = fminf(fmaxf(u1, 1e-10f), 1.0f);
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

// Noise descriptor shared by the fp32 and tensor-core noise epilogues.
struct NoiseSrc {
  int mode;              // SNACB_NOISE_*
  const float* tensor;   // [rows][stride], value of (row, block b, t) at row*stride + off + t
  long long stride;
  int off;               // offset of this block's noise inside a row
  unsigned long long seed;
  const unsigned long long* keys;  // per code_row stream key (nullptr -> code_row)
  int block;
  const unsigned long long* seed_ptr;  // device-resident seed (CUDA-graph replays); nullptr -> `seed`
};

__device__ __forceinline__ float noise_at(const NoiseSrc& ns, int code_row, int t_abs) {
  if (ns.mode == 1) return ns.tensor[(long long)code_row * ns.stride + ns.off + t_abs];
  if (ns.mode == 2) {
    unsigned long long key = ns.keys ? ns.keys[code_row] : (unsigned long long)code_row;
    return philox_normal(ns.seed_ptr ? *ns.seed_ptr : ns.seed, key, (uint32_t)ns.block, (uint32_t)t_abs);
  }
  return 0.0f;
}

}  // namespace snacb
