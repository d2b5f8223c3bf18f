// Whole-DecoderBlock ResidualUnit chain in ONE persistent kernel (tensor-core recipe, sm_100a):
//
//   x1 = x0 + W1 Snake(dw7_d1(Snake(x0))) + b1;  x2 = ... d3 ...;  x3 = ... d9 ...
//   block 3 (C = 64) additionally: Snake -> conv k7 64->1 -> tanh -> slice -> trunc(y * 32767) int16
//
// Replaces, for one decoder block, the three ResidualUnits of the third-party `snac` DecoderBlock (oracle
// `snac_ref.py` DecoderBlock / ResidualUnit; call site Morpheus_Client/tts_engine/speechpipe.py:118) and, for the last
// block, the decoder tail + the window slice / PCM pack of speechpipe.py:122-129.  Versus one kernel per ResidualUnit
// (k_ru_tc / k_ru_x + k_tail) the fp32 residual stream never leaves the SM:
//
//   * the residual stream x lives in TENSOR MEMORY: a tile of 256 time rows x C channels is two 128-lane accumulators;
//     it is written once (tcgen05.st of the block input) and every ResidualUnit's 1x1 conv ACCUMULATES onto it
//     (tcgen05.mma with the accumulate flag) - the residual add costs nothing and no epilogue ever re-reads x;
//   * shared memory holds s = Snake(x) (fp32, the depthwise conv's input) for the tile, produced exactly once per
//     element by the pass that reads the accumulator (row per lane, 16 columns per step), so the depthwise units do
//     no Snake of their inputs at all (the per-unit kernels Snake every input 22/16 times);
//   * 39 halo rows per side (3 * (1 + 3 + 9)) are recomputed per tile: rows whose inputs are missing hold garbage
//     that no valid row ever reads (a GEMM row depends only on its own operand row);
//   * biases are never written back to the accumulator: the pass after unit r adds the accumulated bias
//     b1 + .. + br on the fly.
//
// One CTA = 512 threads, all of them walk the same phases (load -> [depthwise -> MMA -> Snake pass] x 3 -> tail);
// two CTAs per SM cover each other's barrier / MMA-completion waits.  Everything a tile needs from HBM is its
// 256 x C fp32 input rows (cp.async, L2-prefetched one tile ahead); it writes 2 bytes per emitted sample.
#include <cstdlib>

#include "snacb.h"
#include "tc_ptx.cuh"

namespace snacb {
namespace {

constexpr int kBlkThreads = 512;
constexpr int kBlkWarps = kBlkThreads / 32;
constexpr int kBlkRows = 256;  // tile rows = 2 MMA sub-tiles of 128
constexpr int kBlkHalo = 39;   // 3 * (1 + 3 + 9) rows of context per side

struct BlkDev {
  const Item* items; int base, out_len, T0;
  const float* x; int in_lo, in_rows;  // block input (ConvTranspose1d + NoiseBlock output), fp32 [item][in_rows][C]
  int up;                               // time scale of this block's rows
  int o_lo, o_n;                        // relative time of output element 0 and outputs per item
  int tile_stride, row_off;             // tile t: row i <-> relative time o_lo + t * tile_stride + row_off + i
  int tiles_per_item, total_tiles;
  const float* w7[3]; const float* dw_b[3]; const float* a1[3]; const float* i1[3]; const float* a2[3]; const float* i2[3];
  const float* pw_b[3];
  const float* sn_alpha; const float* sn_inv;  // Snake after the block (decoder tail / next block)
  const float* tail_w7; const float* tail_b; int32_t* status; float* wav; int16_t* pcm;
  int dbg;  // SNACB_BLK_DBG bits (timing experiments, results wrong): 1 no depthwise units, 2 no Snake passes, 4 no tail, 8 no MMA
};

template <int C> struct BlkSmem {
  static constexpr int kPitch = C + 4;  // floats per activation row: 16-byte aligned, conflict-free row-per-lane float4
  // fp16 operand of the 1x1 GEMM in the canonical NO-SWIZZLE K-major UMMA layout, 16-byte K chunk outermost:
  // byte(row, col) = (col / 8) * kLbo + row * 16 + (col % 8) * 2  (core matrix = 8 rows x 16 B contiguous, SBO = 128 B,
  // LBO = 256 rows * 16 B + 16 B of padding so the eight chunks of one row fall into different banks).  A depthwise
  // unit's outputs are then base + immediate: no per-output address arithmetic (SWIZZLE_128B XORs the row into it).
  static constexpr int kLbo = kBlkRows * 16 + 16;
  static constexpr int kABytes = (C / 8) * kLbo;
  static constexpr int kWBytes = C * C * 2;  // 1x1 weight of the current unit [k-block][C rows][128 B], SWIZZLE_128B (TMA)
  static constexpr int kSBytes = kBlkRows * kPitch * 4;
  static constexpr int kConstBytes = 4 * 3 * C * 4;  // per pass: accumulated bias, 2 * alpha, -1 / (2 alpha)
  static constexpr int kTailBytes = 7 * C * 4;
  static constexpr int kOffA = kWBytes;  // the weight sits first: its swizzled tiles need 1024-byte alignment
  static constexpr int kOffS = kOffA + kABytes;
  static constexpr int kOffConst = kOffS + kSBytes;
  static constexpr int kOffTail = kOffConst + kConstBytes;
  static constexpr int kOffBar = kOffTail + kTailBytes;
  static constexpr int kBytes = kOffBar + 64;
  static_assert(kOffA % 16 == 0 && kOffS % 16 == 0 && kOffBar % 8 == 0, "alignment");
};

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, bool valid) {
  const int n = valid ? 16 : 0;  // src-size 0: the 16 destination bytes are zero-filled (rows outside the item)
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
// K-major operand without swizzle: 8-row x 16-byte core matrices, `lbo` bytes between the two K chunks of one MMA,
// `sbo` bytes between 8-row groups.
__device__ __forceinline__ uint64_t umma_desc_k_noswz(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell); layout type bits [61,64) = 0: SWIZZLE_NONE
  return d;
}

// Snake in its cosine form with the constant term left out:  Snake(x) = x + sin^2(a x) / a = [x - h cos(2 a x)] + h,
// h = 1 / (2 (a + 1e-9)).  This returns the bracket (one FMUL2, two MUFU.COS with their range-reduction multiply, one
// FFMA2 per channel pair - the sin^2 form needs a second FMUL2); the "+ h" is a per-channel constant that the consumer
// of the value has folded into its bias (depthwise conv: b + h sum_k w_k; tail conv likewise; fp32 consumers only).
// Zero padding stays exact: x = 0 gives -h, i.e. Snake = 0.
__device__ __forceinline__ float2 snake_c(float2 x, float2 al2, float2 nh) {
  const float2 t = __fmul2_rn(al2, x);
  return __ffma2_rn(nh, make_float2(__cosf(t.x), __cosf(t.y)), x);
}

// Depthwise k7 (dilation DIL) of one channel pair over L outputs of one residue class, inputs already Snake'd in
// shared memory: out[j] = Snake2(b + sum_k w[k] * s[j + k]) with s[m] at p0 + m * DIL * PITCH.  Transposed form:
// every input is scattered into the (up to) seven accumulators it feeds, seven independent FFMA2 per input.
struct FirW {
  float2 al2, iv2, bias, w[7];
  // h1: the constant the producer of s left out (see snake_c); folded into the conv's bias here
  __device__ __forceinline__ void load(const float* w7, const float* dw_b, const float* i1, const float* a2, const float* i2,
                                       int C, int c) {
    // the unit's second Snake feeds the fp16 GEMM operand: it keeps the sin^2 form, whose values are small where x is
    // small (dropping the constant h would shift them by up to 1 and cost fp16 resolution: measured -2 dB SNR)
    al2 = *reinterpret_cast<const float2*>(a2 + c); iv2 = *reinterpret_cast<const float2*>(i2 + c);
    float2 sw = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      w[k] = *reinterpret_cast<const float2*>(w7 + k * C + c);
      sw.x += w[k].x; sw.y += w[k].y;
    }
    const float2 b = *reinterpret_cast<const float2*>(dw_b + c), h1 = *reinterpret_cast<const float2*>(i1 + c);
    bias = make_float2(fmaf(0.5f * h1.x, sw.x, b.x), fmaf(0.5f * h1.y, sw.y, b.y));
  }
};
template <int PITCH, int DIL, int L, typename Sink>
__device__ __forceinline__ void fir_unit(const float* p0, const FirW& W, Sink&& sink) {
  float2 acc[7];
#pragma unroll
  for (int m = 0; m < L + 6; ++m) {
    const float2 v = *reinterpret_cast<const float2*>(p0 + m * (DIL * PITCH));
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      const int j = m - k;
      if (j >= 0 && j < L) acc[j % 7] = __ffma2_rn(W.w[k], v, (k == 0) ? W.bias : acc[j % 7]);
    }
    if (m >= 6) sink(m - 6, snake2(acc[(m - 6) % 7], W.al2, W.iv2));
  }
}

// Depthwise phase of one ResidualUnit over the whole tile.  Valid output rows are [LO, 256 - LO) (LO = rows of context
// consumed so far); they are cut into NU = DIL * S units (residue class rho, segment) of LMAX or LMAX - 1 outputs, a
// warp takes one unit for 32 channel pairs at a time (a warp reads / writes one whole row segment: conflict-free).
template <int C, int DIL>
__device__ __forceinline__ void dw_phase(const BlkDev& a, int r, const float* sX, uint8_t* sA, int warp, int lane) {
  using S_ = BlkSmem<C>;
  constexpr int PITCH = S_::kPitch;
  constexpr int LO = (DIL == 1) ? 3 : (DIL == 3) ? 12 : 39;
  constexpr int N = kBlkRows - 2 * LO;
  constexpr int S = (DIL == 1) ? 16 : (DIL == 3) ? 5 : 3;  // segments per residue class
  constexpr int NU = DIL * S;
  constexpr int NMAX = (N + DIL - 1) / DIL;
  constexpr int LMAX = (NMAX + S - 1) / S;
  constexpr int PG = C / 64;  // groups of 32 channel pairs
#pragma unroll 1
  for (int w = warp; w < NU * PG; w += kBlkWarps) {
    const int u = w / PG, pg = w - u * PG;
    const int c = pg * 64 + 2 * lane;
    FirW W;
    W.load(a.w7[r], a.dw_b[r], a.i1[r], a.a2[r], a.i2[r], C, c);
    const int rho = u % DIL, seg = u / DIL;
    const int n = (N - rho + DIL - 1) / DIL;  // outputs of this residue class
    const int k0 = seg * n / S, L = (seg + 1) * n / S - k0;
    const int first = LO + rho + k0 * DIL;  // tile row of the unit's first output
    const float* p0 = sX + (first - 3 * DIL) * PITCH + c;
    uint8_t* ap = sA + (c >> 3) * S_::kLbo + (c & 7) * 2 + first * 16;
    auto sink = [&](int j, float2 v) { *reinterpret_cast<__half2*>(ap + j * (DIL * 16)) = f2h2_sat(v.x, v.y); };
    if (L == LMAX) fir_unit<PITCH, DIL, LMAX>(p0, W, sink);
    else fir_unit<PITCH, DIL, LMAX - 1>(p0, W, sink);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
}

// Pass over the accumulator, row per lane and 16 columns per step (warp w: TMEM lane quarter w % 4, column groups
// w / 4, w / 4 + 4, ..).  STAGE 0: the block input x0 goes from the staged tile into tensor memory; STAGE r > 0:
// x_r = D + (accumulated bias of units 1..r).  Either way s = Snake_next(x) - h replaces the tile row in shared memory
// (x is taken as zero for rows outside the sequence: they are the next conv's zero padding).  Rows [R0, R1) only.
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
// NC columns of one row: x (+ accumulated bias) -> Snake_next(x) - h -> the tile row in shared memory
template <int C, int STAGE, bool CHECK, int NC>
__device__ __forceinline__ void snake_cols(const uint32_t (&r)[NC], const float* cst, float* sp, bool live) {
#pragma unroll
  for (int j = 0; j < NC / 4; ++j) {
    float2 lo = make_float2(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]));
    float2 hi = make_float2(__uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
    if (STAGE > 0) {
      const float4 b4 = *reinterpret_cast<const float4*>(cst + 4 * j);
      lo = __fadd2_rn(lo, make_float2(b4.x, b4.y));
      hi = __fadd2_rn(hi, make_float2(b4.z, b4.w));
    }
    if (CHECK && !live) { lo = make_float2(0.f, 0.f); hi = lo; }
    const float4 al = *reinterpret_cast<const float4*>(cst + C + 4 * j);
    const float4 nh = *reinterpret_cast<const float4*>(cst + 2 * C + 4 * j);
    lo = snake_c(lo, make_float2(al.x, al.y), make_float2(nh.x, nh.y));
    hi = snake_c(hi, make_float2(al.z, al.w), make_float2(nh.z, nh.w));
    *reinterpret_cast<float4*>(sp + 4 * j) = make_float4(lo.x, lo.y, hi.x, hi.y);
  }
}

// STAGE 0: the staged block input goes into tensor memory (tcgen05.st) and is Snake'd in place.
template <int C, bool CHECK>
__device__ __forceinline__ void snake_pass0(float* sX, const float* sConst, uint32_t tmem_base, int warp, int lane, int t_abs0,
                                            int t_hi) {
  constexpr int PITCH = BlkSmem<C>::kPitch;
  const int q = warp & 3;
#pragma unroll 1
  for (int g = warp >> 2; g < C / 16; g += kBlkWarps / 4) {
    const float* cst = sConst + g * 16;
#pragma unroll
    for (int sub = 0; sub < 2; ++sub) {
      const int i = sub * 128 + q * 32 + lane;
      float* sp = sX + i * PITCH + g * 16;
      uint32_t r[16];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 v = *reinterpret_cast<const float4*>(sp + 4 * j);
        r[4 * j] = __float_as_uint(v.x); r[4 * j + 1] = __float_as_uint(v.y);
        r[4 * j + 2] = __float_as_uint(v.z); r[4 * j + 3] = __float_as_uint(v.w);
      }
      tmem_st16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(sub * C + g * 16), r);
      snake_cols<C, 0, CHECK, 16>(r, cst, sp, (unsigned)(t_abs0 + i) < (unsigned)t_hi);
    }
  }
  tmem_st_wait();
}

// STAGE r > 0: x_r = D + (accumulated bias of units 1..r) read out of tensor memory, 8 columns per tcgen05.ld with the
// next load in flight while the current columns are Snake'd (the load latency is otherwise exposed once per step).
// Only the sub-tiles whose 32 rows of this warp intersect [R0, R1) are touched.
template <int C, int STAGE, int R0, int R1, bool CHECK>
__device__ __forceinline__ void snake_pass(float* sX, const float* sConst, uint32_t tmem_base, int warp, int lane, int t_abs0,
                                           int t_hi) {
  constexpr int PITCH = BlkSmem<C>::kPitch;
  const int q = warp & 3;
  const int sub_lo = (q * 32 + 32 > R0) ? 0 : 1, sub_hi = (128 + q * 32 < R1) ? 2 : 1;
#pragma unroll 1
  for (int g = warp >> 2; g < C / 16; g += kBlkWarps / 4) {
    const float* cst = sConst + STAGE * 3 * C + g * 16;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * 16);
    uint32_t ra[8], rb[8];
    if (sub_lo < sub_hi) tmem_ld8_nowait(taddr + sub_lo * C, ra);
#pragma unroll 1
    for (int sub = sub_lo; sub < sub_hi; ++sub) {
      const int i = sub * 128 + q * 32 + lane;
      float* sp = sX + i * PITCH + g * 16;
      const bool live = (unsigned)(t_abs0 + i) < (unsigned)t_hi;
      tmem_ld_wait();
      tmem_ld8_nowait(taddr + sub * C + 8, rb);
      snake_cols<C, STAGE, CHECK, 8>(ra, cst, sp, live);
      tmem_ld_wait();
      if (sub + 1 < sub_hi) tmem_ld8_nowait(taddr + (sub + 1) * C, ra);
      snake_cols<C, STAGE, CHECK, 8>(rb, cst + 8, sp + 8, live);
    }
  }
}

template <int C>
__global__ void __launch_bounds__(kBlkThreads, 2) k_blk_tail(const __grid_constant__ CUtensorMap tmW0,
                                                             const __grid_constant__ CUtensorMap tmW1,
                                                             const __grid_constant__ CUtensorMap tmW2, const BlkDev a) {
  using S = BlkSmem<C>;
  constexpr int PITCH = S::kPitch;
  constexpr int KB = C / BK;
  constexpr int kTO = kBlkRows - 2 * kBlkHalo - 6;  // samples emitted per tile
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sW = smem;
  uint8_t* sA = smem + S::kOffA;
  float* sX = reinterpret_cast<float*>(smem + S::kOffS);
  float* sConst = reinterpret_cast<float*>(smem + S::kOffConst);  // [4 passes][bias, 2 alpha, -h][C]
  float* sTw = reinterpret_cast<float*>(smem + S::kOffTail);      // [7][C] tail conv weight
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kOffBar);  // [0] weight landed, [1] MMAs complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  float* tail_bias = reinterpret_cast<float*>(tmem_slot + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar_w = smem_u32(&bars[0]), bar_mma = smem_u32(&bars[1]);

  if (tid == 0) {
    if (smem_u32(smem) & 1023u) { printf("snacb: k_blk dynamic shared memory is not 1024-byte aligned\n"); __trap(); }
    mbar_init(bar_w, 1);
    mbar_init(bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 2 * C);
  // per-pass constants: pass p runs after p ResidualUnits.  Accumulated bias of the 1x1 convs; Snake of the NEXT
  // consumer as (2 alpha, -h).
  for (int e = tid; e < 4 * C; e += kBlkThreads) {
    const int p = e / C, c = e - p * C;
    float b = 0.0f;
    for (int i = 0; i < p; ++i) b += a.pw_b[i][c];
    sConst[(p * 3 + 0) * C + c] = b;
    sConst[(p * 3 + 1) * C + c] = 2.0f * ((p < 3) ? a.a1[p][c] : a.sn_alpha[c]);
    sConst[(p * 3 + 2) * C + c] = -0.5f * ((p < 3) ? a.i1[p][c] : a.sn_inv[c]);
  }
  for (int e = tid; e < 7 * C; e += kBlkThreads) sTw[e] = a.tail_w7[e];
  if (warp == 2) {  // tail bias + the tail Snake's left-out constant through the tail conv
    float v = 0.0f;
    for (int e = lane; e < 7 * C; e += 32) v = fmaf(a.tail_w7[e], 0.5f * a.sn_inv[e % C], v);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) *tail_bias = v + a.tail_b[0];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const float tail_b = *tail_bias;
  uint32_t nstep = 0;  // ResidualUnit steps done by this CTA: phase parity of both barriers
  auto load_w = [&](const CUtensorMap* tm) {
    mbar_arrive_expect_tx(bar_w, S::kWBytes);
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) tma_load_2d(smem_u32(sW + kb * C * 128), tm, bar_w, kb * BK, 0);
  };
  if (tid == 0 && (int)blockIdx.x < a.total_tiles) load_w(&tmW0);

#pragma unroll 1
  for (int lin = blockIdx.x; lin < a.total_tiles; lin += gridDim.x) {
    const int item = lin / a.tiles_per_item, t = lin - item * a.tiles_per_item;
    const ItemRef it = get_item(a.items, a.base, item, a.out_len);
    const int tau0 = a.o_lo + t * a.tile_stride + a.row_off;  // relative time of tile row 0
    const int t_abs0 = tau0 + it.shift0 * a.up, t_hi = a.T0 * a.up;
    const bool more = lin + (int)gridDim.x < a.total_tiles;
    if (tid == 32 && more) {  // this CTA's next tile -> L2 while this one is computed
      const int nl = lin + gridDim.x;
      const int nit = nl / a.tiles_per_item, nt = nl - nit * a.tiles_per_item;
      const int r0 = a.o_lo + nt * a.tile_stride + a.row_off - a.in_lo;
      const int r_lo = max(r0, 0), r_hi = min(r0 + kBlkRows, a.in_rows);
      if (r_hi > r_lo)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.x + ((size_t)nit * a.in_rows + r_lo) * C),
                     "r"((uint32_t)(r_hi - r_lo) * C * 4)
                     : "memory");
    }
    // ---- load: the tile's 256 input rows -> shared memory (rows outside the item's buffer are zeros)
    {
      constexpr int CH = C / 4;                       // 16-byte chunks per row
      constexpr int RS = kBlkThreads / CH;            // rows covered by one sweep of the CTA
      const int xr0 = tau0 - a.in_lo, row = tid / CH, ch = tid - row * CH;
      const float* src = a.x + ((size_t)item * a.in_rows + xr0 + row) * C + ch * 4;
      const uint32_t dst = smem_u32(sX + row * PITCH + ch * 4);
      if (xr0 >= 0 && xr0 + kBlkRows <= a.in_rows) {  // interior tile: no row checks
#pragma unroll
        for (int k = 0; k < kBlkRows / RS; ++k) cp_async_16(dst + k * (RS * PITCH * 4), src + (size_t)k * RS * C, true);
      } else {
#pragma unroll
        for (int k = 0; k < kBlkRows / RS; ++k) {
          const int xr = xr0 + row + k * RS;
          const bool ok = xr >= 0 && xr < a.in_rows;
          cp_async_16(dst + k * (RS * PITCH * 4), ok ? src + (size_t)k * RS * C : a.x, ok);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
    }
    // rows of the tile outside the sequence (warp-uniform test: interior tiles take the select-free code)
    const bool interior = t_abs0 >= 0 && t_abs0 + kBlkRows <= t_hi;
    if (interior) snake_pass0<C, false>(sX, sConst, tmem_base, warp, lane, t_abs0, t_hi);
    else snake_pass0<C, true>(sX, sConst, tmem_base, warp, lane, t_abs0, t_hi);
    tc_fence_before();
    __syncthreads();

    auto ru_step = [&](auto dil_c, int r, const CUtensorMap* next_w) {
      constexpr int DIL = decltype(dil_c)::value;
      if (!(a.dbg & 1)) dw_phase<C, DIL>(a, r, sX, sA, warp, lane);
      __syncthreads();
      if (tid == 0) {
        mbar_wait(bar_w, nstep & 1u);
        tc_fence_after();
        constexpr uint32_t idesc = umma_idesc_f16(C);
#pragma unroll
        for (int sub = 0; sub < 2; ++sub)
#pragma unroll
          for (int k = 0; k < C / 16; ++k) {  // one MMA = 16 channels = two 16-byte K chunks of the operand
            const uint64_t da = umma_desc_k_noswz(smem_u32(sA + 2 * k * S::kLbo + sub * (BM * 16)), S::kLbo, 128);
            const uint64_t db = umma_desc_k_sw128(smem_u32(sW + (k / 4) * (C * 128))) + 2 * (k % 4);
            if (!(a.dbg & 8)) umma_f16(tmem_base + sub * C, da, db, idesc, 1u);  // D += A W^T on top of the residual stream
          }
        umma_commit(bar_mma);
      }
      mbar_wait(bar_mma, nstep & 1u);
      tc_fence_after();
      if (tid == 0 && next_w) load_w(next_w);  // the MMAs have finished reading this unit's weight
      ++nstep;
    };
    ru_step(IntC<1>{}, 0, &tmW1);
    if (a.dbg & 2) {}
    else if (interior) snake_pass<C, 1, 3, kBlkRows - 3, false>(sX, sConst, tmem_base, warp, lane, t_abs0, t_hi);
    else snake_pass<C, 1, 3, kBlkRows - 3, true>(sX, sConst, tmem_base, warp, lane, t_abs0, t_hi);
    tc_fence_before();
    __syncthreads();
    ru_step(IntC<3>{}, 1, &tmW2);
    if (a.dbg & 2) {}
    else if (interior) snake_pass<C, 2, 12, kBlkRows - 12, false>(sX, sConst, tmem_base, warp, lane, t_abs0, t_hi);
    else snake_pass<C, 2, 12, kBlkRows - 12, true>(sX, sConst, tmem_base, warp, lane, t_abs0, t_hi);
    tc_fence_before();
    __syncthreads();
    ru_step(IntC<9>{}, 2, more ? &tmW0 : nullptr);
    if (a.dbg & 2) {}
    else if (interior) snake_pass<C, 3, kBlkHalo, kBlkRows - kBlkHalo, false>(sX, sConst, tmem_base, warp, lane, t_abs0, t_hi);
    else snake_pass<C, 3, kBlkHalo, kBlkRows - kBlkHalo, true>(sX, sConst, tmem_base, warp, lane, t_abs0, t_hi);
    tc_fence_before();
    __syncthreads();

    // ---- decoder tail on the tile: y[o] = tanh(b + sum_{k,c} w[k][c] * s[39 + o + k][c]), o = 0..kTO-1.
    // Warp w owns outputs [11 w, 11 w + 11), lane = channel pair: every tile row is read ONCE (LDS.64, transposed-form
    // FIR over the 7 taps as in the depthwise units), then the 11 per-lane partial sums are reduced across the warp.
    // (One thread per output re-reads each row seven times: that version spent half of the kernel's shared-memory
    // wavefronts here and kept only 11 of 16 warps busy.)
    if (!(a.dbg & 4)) {
      constexpr int LT = (kTO + kBlkWarps - 1) / kBlkWarps;
      static_assert(C == 64, "tail: one lane per channel pair");
      const int o0 = warp * LT;
      float2 w[7];
#pragma unroll
      for (int k = 0; k < 7; ++k) w[k] = *reinterpret_cast<const float2*>(sTw + k * C + 2 * lane);
      const float* p0 = sX + (kBlkHalo + o0) * PITCH + 2 * lane;
      float vals[LT];
      float2 acc[7];
#pragma unroll
      for (int m = 0; m < LT + 6; ++m) {
        const float2 v = *reinterpret_cast<const float2*>(p0 + m * PITCH);
#pragma unroll
        for (int k = 0; k < 7; ++k) {
          const int j = m - k;
          if (j >= 0 && j < LT) acc[j % 7] = (k == 0) ? __fmul2_rn(w[0], v) : __ffma2_rn(w[k], v, acc[j % 7]);
        }
        if (m >= 6) vals[m - 6] = acc[(m - 6) % 7].x + acc[(m - 6) % 7].y;
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1)
#pragma unroll
        for (int j = 0; j < LT; ++j) vals[j] += __shfl_xor_sync(0xffffffffu, vals[j], off);
      float mine = 0.0f;
#pragma unroll
      for (int j = 0; j < LT; ++j)
        if (lane == j) mine = vals[j];
      const int o = o0 + lane;
      const int oi = t * kTO + o;  // emitted sample index inside [0, o_n)
      const int t_abs = a.o_lo + oi + it.shift0 * a.up;
      if (lane < LT && o < kTO && oi < a.o_n && t_abs >= 0 && t_abs < t_hi &&
          !(a.status && a.status[it.code_row] != SNACB_WIN_OK)) {
        float y = tanhf(mine + tail_b);
        const long long d = it.dst + oi;
        if (!(fabsf(y) <= 1.0f)) {  // NaN: never emitted, the window is reported (SNACB_WIN_NONFINITE) instead
          y = 0.0f;
          if (a.status) a.status[it.code_row] = SNACB_WIN_NONFINITE;
        }
        if (a.wav) a.wav[d] = y;
        if (a.pcm) a.pcm[d] = (int16_t)(y * 32767.0f);
      }
    }
    __syncthreads();  // the tile is free for the next load
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * C);
  }
}

}  // namespace

bool blk_tc_supported(int C, bool tail) { return C == 64 && tail; }

cudaError_t launch_blk_tc(const GroupCtx& g, const BlkTcArgs& a) {
  if (!blk_tc_supported(a.C, a.tail_w7 != nullptr) || g.n_items <= 0 || a.tail_out.n() <= 0) return cudaErrorInvalidValue;
  constexpr int C = 64;
  CUtensorMap mw[3];
  for (int r = 0; r < 3; ++r)
    if (!get_tmap(a.ru[r].pw16, C, C, C, &mw[r])) return cudaErrorNotSupported;
  BlkDev d{};
  d.items = g.items; d.base = g.base; d.out_len = g.out_len; d.T0 = g.T0;
  d.x = a.x; d.in_lo = a.in_r.lo; d.in_rows = a.in_r.n(); d.up = a.up;
  constexpr int kTO = kBlkRows - 2 * kBlkHalo - 6;
  d.o_lo = a.tail_out.lo; d.o_n = a.tail_out.n();
  static const int dbg = [] { const char* v = getenv("SNACB_BLK_DBG"); return v ? atoi(v) : 0; }();
  d.dbg = dbg;
  d.tile_stride = kTO; d.row_off = -(kBlkHalo + 3);
  d.tiles_per_item = (d.o_n + kTO - 1) / kTO;
  const long long total = (long long)d.tiles_per_item * g.n_items;
  if (total >= (1LL << 31)) return cudaErrorInvalidValue;
  d.total_tiles = (int)total;
  for (int r = 0; r < 3; ++r) {
    d.w7[r] = a.ru[r].w7; d.dw_b[r] = a.ru[r].dw_b; d.a1[r] = a.ru[r].a1; d.i1[r] = a.ru[r].i1;
    d.a2[r] = a.ru[r].a2; d.i2[r] = a.ru[r].i2; d.pw_b[r] = a.ru[r].pw_b;
  }
  d.sn_alpha = a.sn_alpha; d.sn_inv = a.sn_inv;
  d.tail_w7 = a.tail_w7; d.tail_b = a.tail_b; d.status = a.status; d.wav = a.wav; d.pcm = a.pcm;
  static bool attr_dev[kMaxDev] = {};
  bool& attr_set = attr_dev[cur_dev()];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_blk_tail<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, BlkSmem<C>::kBytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_blk_tail<C>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    if (getenv("SNACB_DEBUG")) {
      int n = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_blk_tail<C>, kBlkThreads, BlkSmem<C>::kBytes);
      fprintf(stderr, "snacb: k_blk_tail<%d> occupancy query %d CTAs/SM, %d bytes smem\n", C, n, BlkSmem<C>::kBytes);
      for (int b = 0; b <= BlkSmem<C>::kBytes; b += 8192) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_blk_tail<C>, kBlkThreads, b);
        fprintf(stderr, " %d:%d", b, n);
      }
      for (int th = 128; th <= 1024; th += 128) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_blk_tail<C>, th, 32768);
        fprintf(stderr, " th%d:%d", th, n);
      }
      cudaFuncAttributes fa;
      cudaFuncGetAttributes(&fa, k_blk_tail<C>);
      fprintf(stderr, "\n regs %d static smem %zu local %zu maxdyn %d\n", fa.numRegs, fa.sharedSizeBytes, fa.localSizeBytes, fa.maxDynamicSharedSizeBytes);
    }
    attr_set = true;
  }
  // two CTAs per SM by construction (64 registers x 512 threads, <= 113 KB of shared memory each)
  const int grid = (int)std::min<long long>(total, 2LL * sm_count());
  k_blk_tail<C><<<grid, kBlkThreads, BlkSmem<C>::kBytes, g.stream>>>(mw[0], mw[1], mw[2], d);
  ++*g.launches;
  return cudaGetLastError();
}

}  // namespace snacb
