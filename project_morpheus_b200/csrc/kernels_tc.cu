// Tensor-core recipe (SNACB_PREC_FP16) of the SNAC-24k decoder's GEMM-shaped layers for sm_100a:
// tcgen05.mma (kind::f16, fp16 operands, fp32 accumulators in TMEM) fed by TMA, warp-specialised
// (TMA producer / MMA issuer / 4 epilogue warps), with the layer's bias / residual / noise / Snake
// work fused into the epilogue, plus the register-sliding-window depthwise k=7 kernel that produces
// the fp16 GEMM operand of every ResidualUnit.
//
// Replaces (third-party `snac` decoder, called at Morpheus_Client/tts_engine/speechpipe.py:118):
//   k_gemm_tc<.., EPI_BIAS>   decoder.model.1 (1x1 768->1024)  + the Snake of block 0
//   k_gemm_tc<.., EPI_CONVT>  DecoderBlock ConvTranspose1d (k=2s, stride s) as a polyphase GEMM
//   k_gemm_tc<.., EPI_NOISE>  NoiseBlock  x + n[t] * (W_n x)
//   k_gemm_tc<.., EPI_RESID>  ResidualUnit 1x1 + residual add (+ the next block's Snake)
//   k_dw_tc                   ResidualUnit Snake -> depthwise k7 (dil 1/3/9) -> Snake
//
// GEMM view: D[m][n] = sum_k A[m][k] * W[n][k]; rows m = time steps (channels-last activations, so K
// is contiguous = "K-major" for both operands), n = output channels.  One CTA computes a 128 x BN
// tile: 128 TMEM lanes x BN fp32 columns.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "kernels.h"
#include "snacb.h"
#include "tc_ptx.cuh"

namespace snacb {
namespace {

constexpr int kTcThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
const bool g_no_ws = [] { const char* v = getenv("SNACB_NO_WS"); return v && v[0] == '1'; }();
const bool g_cn_p = [] { const char* v = getenv("SNACB_CN_PERSISTENT"); return !(v && v[0] == '0'); }();
const bool g_convt_p = [] { const char* v = getenv("SNACB_CONVT_PERSISTENT"); return !(v && v[0] == '0'); }();
// cluster + TMA multicast of the A stream (opt-in: measured 1.47 vs 1.42 ms per tick - the layers are HBM-bound on the
// fp32 residual stream, not on the L2 -> SM operand traffic)
const bool g_ws_cluster = [] { const char* v = getenv("SNACB_WS_CLUSTER"); return v && v[0] == '1'; }();


// One depthwise unit: L outputs at rows first, first+DIL, ... of one channel pair from L+6 inputs at
// p0 + m*DIL*C (m = 0..L+5; bit m of `mask` says the row exists, absent rows are the conv's zero pad).
// Output j = Snake2(b + sum_k w[k] * Snake1(in[j + k])) is handed to sink(j, value).
template <int C, int DIL, int L, typename Sink>
__device__ __forceinline__ void dw_unit(const float* p0, uint32_t mask_lo, uint32_t mask_hi, const DwPairW& W, Sink&& sink) {
  // Transposed-form FIR: input m, once Snake'd, is scattered into the (up to) seven outputs j = m-6..m it
  // feeds (tap k = m - j).  The seven FFMA2 of one input are independent of each other and consecutive
  // updates of one accumulator are seven instructions apart, so the FMA latency is covered inside a single
  // warp (a sliding-window gather would chain seven dependent FFMA2 per output).
  float2 acc[7];
#pragma unroll
  for (int m = 0; m < L + 6; ++m) {
    float2 v = make_float2(0.f, 0.f);
    const bool ok = (m < 32) ? ((mask_lo >> m) & 1u) : ((mask_hi >> (m - 32)) & 1u);
    if (ok) v = __ldg(reinterpret_cast<const float2*>(p0 + (long long)m * (DIL * C)));
    v = snake2(v, W.al1, W.iv1);
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      const int j = m - k;  // output fed through tap k
      if (j >= 0 && j < L) acc[j % 7] = __ffma2_rn(W.w[k], v, (k == 0) ? W.bias : acc[j % 7]);
    }
    if (m >= 6) sink(m - 6, snake2(acc[(m - 6) % 7], W.al2, W.iv2));
  }
}
// Same unit with its inputs already in shared memory (a TMA-loaded fp32 tile, row pitch C floats): plain LDS.64
// with immediate offsets, no predicates (rows outside the sequence are zeros in the tile), no global-load latency.
template <int C, int DIL, int L, typename Sink>
__device__ __forceinline__ void dw_unit_smem(const float* p0, const DwPairW& W, Sink&& sink) {
  float2 acc[7];
#pragma unroll
  for (int m = 0; m < L + 6; ++m) {
    const float2 v = snake2(*reinterpret_cast<const float2*>(p0 + m * (DIL * C)), W.al1, W.iv1);
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      const int j = m - k;
      if (j >= 0 && j < L) acc[j % 7] = __ffma2_rn(W.w[k], v, (k == 0) ? W.bias : acc[j % 7]);
    }
    if (m >= 6) sink(m - 6, snake2(acc[(m - 6) % 7], W.al2, W.iv2));
  }
}
// Unit lengths of the 128-row tile decomposition used by the fused ResidualUnit kernels (unit u of a channel pair:
// DIL 1 -> rows 16u..16u+15; DIL 3 -> residue u%3 of the 48-row segment u/3; DIL 9 -> residue u): the number of
// outputs that fall inside the tile is a compile-time constant per group, so no unit computes (or loads inputs for)
// rows beyond row 127 and the per-output bound check disappears.
template <int DIL, typename F>
__device__ __forceinline__ void unit_len_dispatch(int u, F&& f) {
  if (DIL == 1) f(IntC<16>{});
  else if (DIL == 9) { if (u < 2) f(IntC<15>{}); else f(IntC<14>{}); }
  else { if (u < 6) f(IntC<16>{}); else if (u % 3 != 2) f(IntC<11>{}); else f(IntC<10>{}); }
}
// Same unit with its L+6 inputs staged through shared memory by cp.async: all loads of the unit are in
// flight at once without holding a register each (the register-resident version lets the compiler
// interleave loads and uses, which exposes the global latency at almost every input - ncu: long
// scoreboard = 74 % of the stalls of k_dw_tc), then the thread waits ONCE and reads its own slots back
// (stage[m * stride]: the lanes of a warp are contiguous, conflict-free).  No block barrier is needed: a
// thread only reads what it copied itself.
__device__ __forceinline__ void cp_async_8(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = smem_u32(smem_dst);
  const int n = valid ? 8 : 0;  // src-size 0 -> the 8 destination bytes are zero-filled (conv zero padding)
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
// NREG: the last NREG inputs stay in registers (plain loads issued before the wait) when shared memory is short.
template <int C, int DIL, int L, int NREG, typename Sink>
__device__ __forceinline__ void dw_unit_staged(const float* p0, uint32_t mask, const DwPairW& W, float2* stage, int stride,
                                               Sink&& sink) {
  static_assert(L + 6 <= 32, "mask is 32 bits");
  constexpr int NS = L + 6 - NREG;  // staged inputs
#pragma unroll
  for (int m = 0; m < NS; ++m) {
    const bool ok = (mask >> m) & 1u;
    cp_async_8(stage + m * stride, ok ? (const void*)(p0 + (long long)m * (DIL * C)) : (const void*)p0, ok);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  float2 tail[NREG > 0 ? NREG : 1];
#pragma unroll
  for (int m = 0; m < NREG; ++m) {
    tail[m] = make_float2(0.f, 0.f);
    if ((mask >> (NS + m)) & 1u) tail[m] = __ldg(reinterpret_cast<const float2*>(p0 + (long long)(NS + m) * (DIL * C)));
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  float2 acc[7];
#pragma unroll
  for (int m = 0; m < L + 6; ++m) {
    const float2 v = snake2(m < NS ? stage[m * stride] : tail[m < NS ? 0 : m - NS], W.al1, W.iv1);
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      const int j = m - k;
      if (j >= 0 && j < L) acc[j % 7] = __ffma2_rn(W.w[k], v, (k == 0) ? W.bias : acc[j % 7]);
    }
    if (m >= 6) sink(m - 6, snake2(acc[(m - 6) % 7], W.al2, W.iv2));
  }
}
// bit m set <=> 0 <= r0 + m*DIL < rows, for m in [0, n)
template <int DIL>
__device__ __forceinline__ void row_mask(int r0, int rows, int n, uint32_t& lo, uint32_t& hi) {
  // first valid m: ceil(-r0 / DIL) if r0 < 0; one past last valid m: ceil((rows - r0) / DIL)
  int m0 = (r0 < 0) ? (-r0 + DIL - 1) / DIL : 0;
  int m1 = (rows - r0 + DIL - 1) / DIL;
  m0 = min(max(m0, 0), n); m1 = min(max(m1, m0), n);
  const unsigned long long bits = (m1 >= 64 ? ~0ull : ((1ull << m1) - 1ull)) & ~((1ull << m0) - 1ull);
  lo = (uint32_t)bits; hi = (uint32_t)(bits >> 32);
}

struct TcDev {
  const Item* items; int base, out_len, T0, n_items;
  int K, nseg;              // K per segment; ConvT has 2 segments (taps of q and q +- 1)
  int stages;               // smem pipeline depth of this launch
  int a_rows, a_lo;         // operand rows per item, relative time of operand row 0
  int s, p, Cout;           // ConvT only
  const float* bias;
  float* out32; __half* out16; int o_lo, o_rows, ldo;
  const float* sn_alpha; const float* sn_inv;
  const float* R; int r_lo, r_rows, ldr;
  NoiseSrc noise;
  int up;
  int n_tiles;              // k_gemm_tc: N tiles per M tile (linear grid, N tile fastest)
  int split;                // split-operand recipe: segment = tap * 3 + term, A columns [hi | lo]
  const float* bias2;       // composed ConvT + NoiseBlock kernel: W_n b (the noise GEMM's share of the conv bias), [Cout]
};

template <int BN> struct TcSmem {
  static constexpr int kMaxStages = (BN == 64) ? 4 : 3;
  static constexpr int kStageBytes = BM * BK * 2 + BN * BK * 2;
  static constexpr int kMetaBytes = 3 * BM * 4 + 128;
  // stages = min(k-blocks, kMaxStages): short-K layers keep little shared memory so more CTAs share an SM
  static constexpr int bytes(int stages) { return stages * kStageBytes + kMetaBytes + 1024; }  // + alignment slack
  static_assert(kStageBytes >= 8 * 32 * 16 * 4, "epilogue staging aliases pipeline stage 0");
};

template <int BN, int EPI>
__global__ void __launch_bounds__(kTcThreads, 3) k_gemm_tc(const __grid_constant__ CUtensorMap tmA,
                                                         const __grid_constant__ CUtensorMap tmW, const TcDev a) {
  using S = TcSmem<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  const int kStages = a.stages;
  uint8_t* meta = smem + kStages * S::kStageBytes;
  int* meta_out = reinterpret_cast<int*>(meta);
  int* meta_res = meta_out + BM;
  float* meta_nz = reinterpret_cast<float*>(meta_res + BM);
  uint64_t* bars = reinterpret_cast<uint64_t*>(meta + 3 * BM * 4);  // full[kStages], empty[kStages], accum
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S::kMaxStages + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // N tile fastest: the CTAs sharing an operand tile run back to back (one HBM read, the rest L2 hits)
  const int m0 = (blockIdx.x / a.n_tiles) * BM, n0 = (blockIdx.x % a.n_tiles) * BN;
  const long long Mtot = (long long)a.n_items * a.a_rows;
  int phase = 0, delta = 0;
  if (EPI == EPI_CONVT) { phase = n0 / a.Cout; delta = (phase < a.s - a.p) ? -1 : 1; }
  const int kb_per_seg = a.K / BK;
  const int num_kb = kb_per_seg * a.nseg;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(smem_u32(&bars[i]), 1);
      mbar_init(smem_u32(&bars[kStages + i]), 1);
    }
    mbar_init(smem_u32(&bars[2 * kStages]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int st = kb % kStages;
        mbar_wait(smem_u32(&bars[kStages + st]), ((kb / kStages) & 1) ^ 1);
        const uint32_t full = smem_u32(&bars[st]);
        const uint32_t sa = smem_u32(smem + st * S::kStageBytes), sb = sa + BM * BK * 2;
        const int seg = kb / kb_per_seg, kk = (kb - seg * kb_per_seg) * BK;
        mbar_arrive_expect_tx(full, S::kStageBytes);
        if (a.split) {  // terms hi*hi, hi*lo, lo*hi of tap seg / 3: the operand's lo half sits K columns to the right
          const int term = seg % 3, tap = seg / 3;
          tma_load_2d(sa, &tmA, full, kk + (term == 2 ? a.K : 0), m0 + (tap ? delta : 0));
        } else {
          tma_load_2d(sa, &tmA, full, kk, m0 + (seg ? delta : 0));
        }
        tma_load_2d(sb, &tmW, full, seg * a.K + kk, n0);
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (one thread)
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int st = kb % kStages;
        mbar_wait(smem_u32(&bars[st]), (kb / kStages) & 1);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + st * S::kStageBytes), sb = sa + BM * BK * 2;
        const uint64_t da = umma_desc_k_sw128(sa), db = umma_desc_k_sw128(sb);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)  // UMMA_K = 16 fp16 = 32 bytes: advance the start address by 2 (x16 B)
          umma_f16(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
        umma_commit(smem_u32(&bars[kStages + st]));  // frees the stage once these MMAs have read it
      }
      umma_commit(smem_u32(&bars[2 * kStages]));     // accumulator complete
    }
  } else {
    // ===================================================================== epilogue (warps 2..9)
    const int q = warp & 3;            // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;  // the two warps of a quarter split the tile's columns
    if (half == 0) {                   // row metadata, one tile row per thread of the first four warps
      const int trow = q * 32 + lane;
      const long long gm = (long long)m0 + trow;
      int oi = -1, ri = -1;
      float nz = 0.0f;
      if (gm < Mtot) {
        const int item = (int)(gm / a.a_rows), j = (int)(gm - (long long)item * a.a_rows);
        const ItemRef it = get_item(a.items, a.base, item, a.out_len);
        int t_rel = a.a_lo + j;
        if (EPI == EPI_CONVT) t_rel = t_rel * a.s + phase;
        const int orow = t_rel - a.o_lo;
        if (orow >= 0 && orow < a.o_rows) {
          const int t_abs = t_rel + it.shift0 * a.up;
          const bool live = (t_abs >= 0) && (t_abs < a.T0 * a.up);
          oi = item * a.o_rows + orow;
          if (EPI == EPI_RESID || EPI == EPI_NOISE) ri = item * a.r_rows + (t_rel - a.r_lo);
          if (EPI == EPI_NOISE && live) nz = noise_at(a.noise, it.code_row, t_abs);
          if (!live) oi |= (int)kLiveFlag;
        }
      }
      meta_out[trow] = oi; meta_res[trow] = ri; meta_nz[trow] = nz;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");  // epilogue warps only
    const int c4 = lane & 3, r8 = lane >> 2;
    constexpr int NH = BN / 32;  // 16-column half-chunks per warp
    const int ocol0 = n0 + half * (BN / 2) + c4 * 4 - ((EPI == EPI_CONVT) ? phase * a.Cout : 0);
    int oi4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) oi4[i] = meta_out[q * 32 + r8 + 8 * i];
    float4 res[4];
    auto load_res = [&](int h) {  // residual / carrier values of this lane for half-chunk h (independent of the MMA)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        res[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (oi4[i] >= 0)
          res[i] = __ldg(reinterpret_cast<const float4*>(a.R + (size_t)meta_res[q * 32 + r8 + 8 * i] * a.ldr + ocol0 + h * 16));
      }
    };
    if (EPI == EPI_RESID || EPI == EPI_NOISE) load_res(0);
    mbar_wait(smem_u32(&bars[2 * kStages]), 0);
    tc_fence_after();
    float* stg = reinterpret_cast<float*>(smem) + (warp - 2) * (32 * 16);  // aliases stage 0: all MMAs are done
#pragma unroll 1
    for (int h = 0; h < NH; ++h) {
      uint32_t r[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * (BN / 2) + h * 16), r);
      float4 v[4];
      epi_transpose16(stg, lane, r, v);
      const int ocol = ocol0 + h * 16;
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), al = b4, iv = b4;
      if (a.bias) b4 = *reinterpret_cast<const float4*>(a.bias + ocol);
      if (a.sn_alpha) { al = *reinterpret_cast<const float4*>(a.sn_alpha + ocol); iv = *reinterpret_cast<const float4*>(a.sn_inv + ocol); }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int oi = oi4[i];
        if (oi < 0) continue;
        const bool live = !(oi & (int)kLiveFlag);
        oi &= (int)(kLiveFlag - 1);
        float4 x = add4(v[i], b4);
        if (EPI == EPI_NOISE) {
          const float nz = meta_nz[q * 32 + r8 + 8 * i];
          const float2 n2 = make_float2(nz, nz);
          const float2 lo = __ffma2_rn(n2, make_float2(x.x, x.y), make_float2(res[i].x, res[i].y));
          const float2 hi = __ffma2_rn(n2, make_float2(x.z, x.w), make_float2(res[i].z, res[i].w));
          x = make_float4(lo.x, lo.y, hi.x, hi.y);
        } else if (EPI == EPI_RESID) {
          x = add4(x, res[i]);
        }
        if (!live) x = make_float4(0.f, 0.f, 0.f, 0.f);
        const size_t o = (size_t)oi * a.ldo + ocol;
        if (a.out32) *reinterpret_cast<float4*>(a.out32 + o) = x;
        if (a.out16) {
          if (a.sn_alpha) {
            const float2 lo = snake2(make_float2(x.x, x.y), make_float2(al.x, al.y), make_float2(iv.x, iv.y));
            const float2 hi = snake2(make_float2(x.z, x.w), make_float2(al.z, al.w), make_float2(iv.z, iv.w));
            x = make_float4(lo.x, lo.y, hi.x, hi.y);
          }
          store_half4(a.out16 + o, x);
        }
      }
      if ((EPI == EPI_RESID || EPI == EPI_NOISE) && h + 1 < NH) load_res(h + 1);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

// ============================================================================ weight-stationary 1x1 GEMM
// Persistent variant of k_gemm_tc for the wide 1x1 layers of decoder blocks 0 and 1 (C = 512 / 256: the
// ResidualUnit and NoiseBlock GEMMs).  With 128 x 128 tiles and K = C those GEMMs are bound by the L2 ->
// shared-memory traffic of re-loading the weight tile for every M tile (64 FLOP/B).  Here a CTA owns one
// 128-wide slice of output channels, keeps that slice of W ([128][K] fp16, K <= 512: <= 128 KB) in shared
// memory for its whole life and streams M tiles through a 4-stage TMA ring; two TMEM accumulators let the
// epilogue of tile i overlap the MMAs of tile i+1.  One CTA per SM, gridDim.x = (#N slices) * P.
constexpr int kWsStages = 4;
struct WsSmem {
  static constexpr int kABytes = kWsStages * BM * BK * 2;       // 64 KB ring
  static constexpr int kStgBytes = 8 * 32 * 16 * 4;             // epilogue transposes
  static constexpr int kMetaBytes = 2 * 3 * BM * 4 + 256;       // two row-metadata buffers + barriers
  static constexpr int bytes(int K) { return K * 128 * 2 + kABytes + kStgBytes + kMetaBytes + 1024; }
};

// CL > 1: the CL CTAs that own the CL output-channel slices of the same M tiles form a thread-block cluster and
// share the operand stream - every k-block tile of A is fetched from L2 once (by CTA kb % CL) and multicast into
// all CL shared memories, instead of CL times (the C = 512 layers are otherwise bound by that L2 -> SM traffic).
template <int EPI, int CL>
__global__ void __launch_bounds__(kTcThreads, 1) k_gemm_ws(const __grid_constant__ CUtensorMap tmA,
                                                         const __grid_constant__ CUtensorMap tmW, const TcDev a) {
  constexpr int BN = 128;
  constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1u);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  const int KB = a.K / BK;
  uint8_t* sW = smem;                                 // [KB][128 rows][128 B]
  uint8_t* sA = sW + KB * (BN * 128);                 // [kWsStages][128 rows][128 B]
  float* sStg = reinterpret_cast<float*>(sA + WsSmem::kABytes);
  int* meta = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(sStg) + WsSmem::kStgBytes);  // [2][3][BM]
  uint64_t* bars = reinterpret_cast<uint64_t*>(meta + 2 * 3 * BM);
  // bars: [0] w_full, [1..4] a_full, [5..8] a_empty, [9..10] t_full, [11..12] t_empty
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_slices = a.ldo / BN;  // output width == N for these layers
  const int slice = blockIdx.x % n_slices, p = blockIdx.x / n_slices, P = gridDim.x / n_slices;
  const int n0 = slice * BN;
  const long long Mtot = (long long)a.n_items * a.a_rows;
  const int m_tiles = (int)((Mtot + BM - 1) / BM);

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    for (int i = 0; i < kWsStages; ++i) { mbar_init(smem_u32(&bars[1 + i]), 1); mbar_init(smem_u32(&bars[5 + i]), CL); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&bars[9 + i]), 1); mbar_init(smem_u32(&bars[11 + i]), 256); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 2 * BN);
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // peers' barriers are initialised before anyone multicasts into them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t crank = (CL > 1) ? cluster_ctarank() : 0u;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(smem_u32(&bars[0]), (uint32_t)(KB * BN * 128));
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(smem_u32(sW + kb * (BN * 128)), &tmW, smem_u32(&bars[0]), kb * BK, n0);
      int it = 0;  // running k-block counter across tiles
      for (int mt = p; mt < m_tiles; mt += P) {
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int st = it % kWsStages;
          mbar_wait(smem_u32(&bars[5 + st]), ((it / kWsStages) & 1) ^ 1);
          mbar_arrive_expect_tx(smem_u32(&bars[1 + st]), BM * BK * 2);
          if (CL == 1) tma_load_2d(smem_u32(sA + st * (BM * 128)), &tmA, smem_u32(&bars[1 + st]), kb * BK, mt * BM);
          else if ((uint32_t)(it % CL) == crank)
            tma_load_2d_mc(smem_u32(sA + st * (BM * 128)), &tmA, smem_u32(&bars[1 + st]), kb * BK, mt * BM, kMask);
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(BN);
      mbar_wait(smem_u32(&bars[0]), 0);
      int it = 0, ti = 0;
      for (int mt = p; mt < m_tiles; mt += P, ++ti) {
        const int buf = ti & 1;
        mbar_wait(smem_u32(&bars[11 + buf]), ((ti >> 1) & 1) ^ 1);  // epilogue drained this accumulator
        tc_fence_after();
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int st = it % kWsStages;
          mbar_wait(smem_u32(&bars[1 + st]), (it / kWsStages) & 1);
          tc_fence_after();
          const uint64_t da = umma_desc_k_sw128(smem_u32(sA + st * (BM * 128)));
          const uint64_t db = umma_desc_k_sw128(smem_u32(sW + kb * (BN * 128)));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_f16(tmem_base + buf * BN, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
          if (CL == 1) umma_commit(smem_u32(&bars[5 + st]));
          else umma_commit_mc(smem_u32(&bars[5 + st]), kMask);  // the stage is free once ALL CTAs of the cluster consumed it
        }
        umma_commit(smem_u32(&bars[9 + buf]));
      }
    }
  } else {
    // ===================================================================== epilogue (warps 2..9)
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int c4 = lane & 3, r8 = lane >> 2;
    float* stg = sStg + (warp - 2) * (32 * 16);
    float* st_p = stg + lane * 16;
    const int st_x = (lane >> 1) & 3;
    const float* ld_p = stg + r8 * 16 + ((c4 ^ ((r8 >> 1) & 3)) << 2);
    constexpr int NH = BN / 32;
    const int ocol0 = n0 + half * (BN / 2) + c4 * 4;
    int ti = 0;
    for (int mt = p; mt < m_tiles; mt += P, ++ti) {
      const int buf = ti & 1;
      int* m_out = meta + buf * 3 * BM;
      int* m_res = m_out + BM;
      float* m_nz = reinterpret_cast<float*>(m_res + BM);
      if (half == 0) {  // row metadata of this tile, one row per thread of the first four epilogue warps
        const int trow = q * 32 + lane;
        const long long gm = (long long)mt * BM + trow;
        int oi = -1, ri = -1;
        float nz = 0.0f;
        if (gm < Mtot) {
          const int item = (int)(gm / a.a_rows), j = (int)(gm - (long long)item * a.a_rows);
          const ItemRef itr = get_item(a.items, a.base, item, a.out_len);
          const int t_rel = a.a_lo + j;
          const int orow = t_rel - a.o_lo;
          if (orow >= 0 && orow < a.o_rows) {
            const int t_abs = t_rel + itr.shift0 * a.up;
            const bool live = (t_abs >= 0) && (t_abs < a.T0 * a.up);
            oi = item * a.o_rows + orow;
            ri = item * a.r_rows + (t_rel - a.r_lo);
            if (EPI == EPI_NOISE && live) nz = noise_at(a.noise, itr.code_row, t_abs);
            if (!live) oi |= (int)kLiveFlag;
          }
        }
        m_out[trow] = oi; m_res[trow] = ri; m_nz[trow] = nz;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      int oi4[4], ri4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { oi4[i] = m_out[q * 32 + r8 + 8 * i]; ri4[i] = m_res[q * 32 + r8 + 8 * i]; }
      // the whole residual / carrier block of this warp (4 column steps x 4 rows per lane) is requested before the
      // accumulator is waited for: one DRAM latency per tile instead of one per column step
      float4 res[NH][4];
#pragma unroll
      for (int h = 0; h < NH; ++h)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          res[h][i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (oi4[i] >= 0) res[h][i] = __ldg(reinterpret_cast<const float4*>(a.R + (size_t)ri4[i] * a.ldr + ocol0 + h * 16));
        }
      mbar_wait(smem_u32(&bars[9 + buf]), (ti >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        uint32_t r[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + half * (BN / 2) + h * 16), r);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<float4*>(st_p + ((j ^ st_x) << 2)) =
              make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                          __uint_as_float(r[4 * j + 3]));
        __syncwarp();
        float4 v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = *reinterpret_cast<const float4*>(ld_p + i * 128);
        __syncwarp();
        const int ocol = ocol0 + h * 16;
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), al = b4, iv = b4;
        if (a.bias) b4 = __ldg(reinterpret_cast<const float4*>(a.bias + ocol));
        if (a.sn_alpha) { al = __ldg(reinterpret_cast<const float4*>(a.sn_alpha + ocol)); iv = __ldg(reinterpret_cast<const float4*>(a.sn_inv + ocol)); }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          int oi = oi4[i];
          if (oi < 0) continue;
          const bool live = !(oi & (int)kLiveFlag);
          oi &= (int)(kLiveFlag - 1);
          float4 x = add4(v[i], b4);
          if (EPI == EPI_NOISE) {
            const float nz = m_nz[q * 32 + r8 + 8 * i];
            const float2 n2 = make_float2(nz, nz);
            const float2 lo = __ffma2_rn(n2, make_float2(x.x, x.y), make_float2(res[h][i].x, res[h][i].y));
            const float2 hi = __ffma2_rn(n2, make_float2(x.z, x.w), make_float2(res[h][i].z, res[h][i].w));
            x = make_float4(lo.x, lo.y, hi.x, hi.y);
          } else {
            x = add4(x, res[h][i]);
          }
          if (!live) x = make_float4(0.f, 0.f, 0.f, 0.f);
          const size_t o = (size_t)oi * a.ldo + ocol;
          if (a.out32) *reinterpret_cast<float4*>(a.out32 + o) = x;
          if (a.out16) {
            if (a.sn_alpha) {
              const float2 lo = snake2(make_float2(x.x, x.y), make_float2(al.x, al.y), make_float2(iv.x, iv.y));
              const float2 hi = snake2(make_float2(x.z, x.w), make_float2(al.z, al.w), make_float2(iv.z, iv.w));
              x = make_float4(lo.x, lo.y, hi.x, hi.y);
            }
            store_half4(a.out16 + o, x);
          }
        }
      }
      tc_fence_before();
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[11 + buf])) : "memory");  // accumulator free
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // nobody leaves while a peer may still write into its shared memory / barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

// ============================================================================ ConvT + NoiseBlock fused
// x = y + n[t] * (W_n y),  y = ConvTranspose1d(Snake(x_in)) + b, for decoder blocks whose output width
// Cout fits one tile (blocks 2 and 3: Cout = 128 / 64).  The polyphase transposed conv accumulates a
// 128 x Cout tile in TMEM (D1); the epilogue warps add the bias, round to fp16 and write the tile
// straight into the swizzled K-major operand layout in the (now idle) pipeline stage 0 while the
// producer TMA-loads W_n into stage 1; a second tcgen05.mma chain builds D2 = fp16(y) W_n^T beside D1;
// the final epilogue combines D1 + b + n*D2 per row and stores fp32 x coalesced.  Saves the fp32 +
// fp16 round trip of y through HBM (12 B per element) and one launch per block.
template <int BN>
__global__ void __launch_bounds__(kTcThreads, (BN == 64) ? 3 : 2) k_convt_noise_tc(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmW,
                                                                const __grid_constant__ CUtensorMap tmN, const TcDev a) {
  using S = TcSmem<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  const int kStages = a.stages;  // >= 3 (host-checked): stage 0 = fp16 y tile, stage 1 = W_n, stage 2 = transposes
  uint8_t* meta = smem + kStages * S::kStageBytes;
  int* meta_out = reinterpret_cast<int*>(meta);
  uint64_t* bars = reinterpret_cast<uint64_t*>(meta + 3 * BM * 4);  // full[], empty[], accum1, wn_full, y_ready, accum2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S::kMaxStages + 4);
  const uint32_t bar_acc1 = smem_u32(&bars[2 * S::kMaxStages]), bar_wn = smem_u32(&bars[2 * S::kMaxStages + 1]);
  const uint32_t bar_y = smem_u32(&bars[2 * S::kMaxStages + 2]), bar_acc2 = smem_u32(&bars[2 * S::kMaxStages + 3]);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // phase fastest: the s CTAs that read the same operand rows are launched back to back, so the tile comes
  // from HBM once and from L2 s - 1 times (m-tile-fastest order re-read it from HBM for every phase)
  const int phase = blockIdx.x % a.s;  // one N tile per polyphase component (BN == Cout)
  const int m0 = (blockIdx.x / a.s) * BM;
  const int delta = (phase < a.s - a.p) ? -1 : 1;
  const long long Mtot = (long long)a.n_items * a.a_rows;
  const int kb_per_seg = a.K / BK;
  const int num_kb = kb_per_seg * 2;
  constexpr int KB2 = BN / BK;  // k-blocks of the noise GEMM (K = Cout)

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(smem_u32(&bars[i]), 1);
      mbar_init(smem_u32(&bars[S::kMaxStages + i]), 1);
    }
    mbar_init(bar_acc1, 1); mbar_init(bar_wn, 1); mbar_init(bar_y, 1); mbar_init(bar_acc2, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 2 * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  uint8_t* sY = smem;                       // [KB2][128 rows][128 B]
  uint8_t* sWn = smem + S::kStageBytes;     // [KB2][BN rows][128 B]

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int st = kb % kStages;
        mbar_wait(smem_u32(&bars[S::kMaxStages + st]), ((kb / kStages) & 1) ^ 1);
        const uint32_t full = smem_u32(&bars[st]);
        const uint32_t sa = smem_u32(smem + st * S::kStageBytes), sb = sa + BM * BK * 2;
        const int seg = kb / kb_per_seg, kk = (kb - seg * kb_per_seg) * BK;
        mbar_arrive_expect_tx(full, S::kStageBytes);
        tma_load_2d(sa, &tmA, full, kk, m0 + (seg ? delta : 0));
        tma_load_2d(sb, &tmW, full, seg * a.K + kk, phase * BN);
      }
      mbar_wait(bar_acc1, 0);  // every conv MMA has finished reading the stages
      mbar_arrive_expect_tx(bar_wn, BN * BN * 2);
#pragma unroll
      for (int kb = 0; kb < KB2; ++kb) tma_load_2d(smem_u32(sWn + kb * BN * 128), &tmN, bar_wn, kb * BK, 0);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int st = kb % kStages;
        mbar_wait(smem_u32(&bars[st]), (kb / kStages) & 1);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + st * S::kStageBytes), sb = sa + BM * BK * 2;
        const uint64_t da = umma_desc_k_sw128(sa), db = umma_desc_k_sw128(sb);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) umma_f16(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
        umma_commit(smem_u32(&bars[S::kMaxStages + st]));
      }
      umma_commit(bar_acc1);
      mbar_wait(bar_wn, 0);
      mbar_wait(bar_y, 0);
      tc_fence_after();
#pragma unroll
      for (int kb = 0; kb < KB2; ++kb) {
        const uint64_t da = umma_desc_k_sw128(smem_u32(sY + kb * BM * 128));
        const uint64_t db = umma_desc_k_sw128(smem_u32(sWn + kb * BN * 128));
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) umma_f16(tmem_base + BN, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
      }
      umma_commit(bar_acc2);
    }
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int trow = q * 32 + lane;  // TMEM lane == tile row owned by this thread in the row-per-lane phases
    int my_oi = -1;
    float my_nz = 0.0f;
    {
      const long long gm = (long long)m0 + trow;
      if (gm < Mtot) {
        const int item = (int)(gm / a.a_rows), j = (int)(gm - (long long)item * a.a_rows);
        const ItemRef it = get_item(a.items, a.base, item, a.out_len);
        const int t_rel = (a.a_lo + j) * a.s + phase;
        const int orow = t_rel - a.o_lo;
        if (orow >= 0 && orow < a.o_rows) {
          const int t_abs = t_rel + it.shift0 * a.up;
          const bool live = (t_abs >= 0) && (t_abs < a.T0 * a.up);
          my_oi = item * a.o_rows + orow;
          if (live) my_nz = noise_at(a.noise, it.code_row, t_abs);
          else my_oi |= (int)kLiveFlag;
        }
      }
      if (half == 0) meta_out[trow] = my_oi;
    }
    constexpr int NH = BN / 32;  // 16-column half-chunks per warp
    const int colbase = half * (BN / 2);
    mbar_wait(bar_acc1, 0);
    tc_fence_after();
    // ---- phase 1: y = D1 + b -> fp16 -> operand tile (row per lane is exactly the K-major layout)
#pragma unroll 1
    for (int h = 0; h < NH; ++h) {
      const int col = colbase + h * 16;
      uint32_t r[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)col, r);
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + col + 4 * j));
        const __half2 h0 = f2h2_sat(__uint_as_float(r[4 * j]) + b4.x, __uint_as_float(r[4 * j + 1]) + b4.y);
        const __half2 h1 = f2h2_sat(__uint_as_float(r[4 * j + 2]) + b4.z, __uint_as_float(r[4 * j + 3]) + b4.w);
        pk[2 * j] = *reinterpret_cast<const uint32_t*>(&h0);
        pk[2 * j + 1] = *reinterpret_cast<const uint32_t*>(&h1);
      }
      uint8_t* rowp = sY + (col / BK) * (BM * 128) + trow * 128;
      const int ch = ((col % BK) * 2) >> 4;  // first of the two 16-byte chunks
      *reinterpret_cast<uint4*>(rowp + (((ch) ^ (trow & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *reinterpret_cast<uint4*>(rowp + (((ch + 1) ^ (trow & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (warp == 2 && lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_y) : "memory");
    // ---- phase 2: x = (D1 + b) + n * D2, transposed through stage 2, stored coalesced
    const int c4 = lane & 3, r8 = lane >> 2;
    int oi4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) oi4[i] = meta_out[q * 32 + r8 + 8 * i];
    float* stg = reinterpret_cast<float*>(smem + 2 * S::kStageBytes) + (warp - 2) * (32 * 16);
    const bool live = my_oi >= 0 && !(my_oi & (int)kLiveFlag);
    mbar_wait(bar_acc2, 0);
    tc_fence_after();
#pragma unroll 1
    for (int h = 0; h < NH; ++h) {
      const int col = colbase + h * 16;
      uint32_t d1[16], d2[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)col, d1);
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(BN + col), d2);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + col + 4 * j));
        const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float y = __uint_as_float(d1[4 * j + e]) + bb[e];
          const float x = fmaf(my_nz, __uint_as_float(d2[4 * j + e]), y);
          d1[4 * j + e] = __float_as_uint(live ? x : 0.0f);
        }
      }
      float4 v[4];
      epi_transpose16(stg, lane, d1, v);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (oi4[i] < 0) continue;
        const size_t o = (size_t)(oi4[i] & (int)(kLiveFlag - 1)) * a.ldo + col + c4 * 4;
        *reinterpret_cast<float4*>(a.out32 + o) = v[i];
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

// ------------------------------------------------------------------------------------ tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

struct MapKey {
  const void* p; long long rows; int cols, box_rows;
  bool operator==(const MapKey& o) const { return p == o.p && rows == o.rows && cols == o.cols && box_rows == o.box_rows; }
};
struct MapHash {
  size_t operator()(const MapKey& k) const {
    return std::hash<const void*>()(k.p) ^ (std::hash<long long>()(k.rows) * 1000003u) ^ ((size_t)k.cols << 20) ^ (size_t)k.box_rows;
  }
};

// [rows][cols] fp16 row-major, box = 64 columns x box_rows rows, 128-byte swizzle, zero fill out of bounds.
}  // namespace
bool get_tmap(const void* ptr, long long rows, int cols, int box_rows, CUtensorMap* out) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapHash> cache;
  std::lock_guard<std::mutex> lk(mu);
  const MapKey key{ptr, rows, cols, box_rows};
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return true; }
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, m);
  *out = m;
  return true;
}

namespace {
// fp32 activation tensor [items][rows][C] (channels-last), box = [1][box_rows][C], no swizzle: rows outside
// [0, rows) of the addressed item arrive as zeros - exactly the conv's zero padding.
bool get_tmap_x3(const float* ptr, int C, int rows, int n_items, int box_rows, CUtensorMap* out, int box_c = 0) {
  if (box_c <= 0) box_c = C;
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapHash> cache;
  std::lock_guard<std::mutex> lk(mu);
  const MapKey key{ptr, (long long)rows * 65536 + n_items, C * 1024 + box_c, box_rows};
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return true; }
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)rows, (cuuint64_t)n_items};
  cuuint64_t strides[2] = {(cuuint64_t)C * 4, (cuuint64_t)rows * C * 4};
  cuuint32_t box[3] = {(cuuint32_t)box_c, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMap m;
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, m);
  *out = m;
  return true;
}

template <int BN, int EPI>
cudaError_t launch_tc_t(const CUtensorMap& ma, const CUtensorMap& mw, const TcDev& d, dim3 grid, cudaStream_t st) {
  static bool attr_set_dev[kMaxDev] = {};
  bool& attr_set = attr_set_dev[cur_dev()];  // kernel attributes are per device
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_gemm_tc<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         TcSmem<BN>::bytes(TcSmem<BN>::kMaxStages));
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  k_gemm_tc<BN, EPI><<<grid, kTcThreads, TcSmem<BN>::bytes(d.stages), st>>>(ma, mw, d);
  return cudaGetLastError();
}

template <int BN>
cudaError_t launch_tc_bn(int epi, const CUtensorMap& ma, const CUtensorMap& mw, const TcDev& d, dim3 grid, cudaStream_t st) {
  switch (epi) {
    case EPI_BIAS: return launch_tc_t<BN, EPI_BIAS>(ma, mw, d, grid, st);
    case EPI_RESID: return launch_tc_t<BN, EPI_RESID>(ma, mw, d, grid, st);
    case EPI_NOISE: return launch_tc_t<BN, EPI_NOISE>(ma, mw, d, grid, st);
    default: return launch_tc_t<BN, EPI_CONVT>(ma, mw, d, grid, st);
  }
}

// ============================================================================ persistent ConvTranspose1d GEMM
// The wide transposed convs (decoder blocks 0 / 1: K = 2*Cin = 2048 / 1024, N = s*Cout = 4096 / 2048) are the
// tensor-bound layers.  Persistent variant of k_gemm_tc<128, EPI_CONVT>: one CTA per SM walks (m tile, phase/n
// tile) pairs round-robin, A and W stream through a 4-stage TMA ring that never drains between tiles, and two
// TMEM accumulators let the 8 epilogue warps (bias, fp32 + fp16 stores) work on tile i while the MMAs of tile
// i+1 run.
template <int BN> struct CtpSmem {
  static constexpr int kStages = 4;
  static constexpr int kStageBytes = BM * BK * 2 + BN * BK * 2;
  static constexpr int kStgBytes = 8 * 32 * 16 * 4;
  static constexpr int kBytes = kStages * kStageBytes + kStgBytes + 4 * BM * 4 + 256 + 1024;  // meta: [2][BM] rows + [2][BM] noise
};

// N2 (composed ConvT + NoiseBlock, blocks 0 / 1): the NoiseBlock x = y + n[t] * (W_n y) of the conv output y = A W_c^T + b
// needs no second pass over y: W_n y = A (W_n W_c)^T + W_n b is a GEMM over the SAME operand tiles with a composed weight.
// The weight matrix handed in is stacked per 128-channel slice: 128 rows of W_c, then the 128 matching rows of W_n W_c -
// one 256-row operand box, one N = 256 MMA: accumulator columns [0,128) hold the conv, [128,256) the noise GEMM, and the
// epilogue writes y + b + n[t] * (noise + W_n b) as fp32.  Twice the MMAs of the plain conv (the tensor pipe has room:
// 54-70 % busy), but the separate noise GEMM kernel, its fp16 operand copy and one fp32 round trip of y are gone.
template <int BN, bool N2>
__global__ void __launch_bounds__(kTcThreads, 1) k_convt_p(const __grid_constant__ CUtensorMap tmA,
                                                         const __grid_constant__ CUtensorMap tmW, const TcDev a,
                                                         const int n_tiles, const int total_tiles) {
  using S = CtpSmem<BN>;
  constexpr int CN = N2 ? BN / 2 : BN;  // output channels per tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  float* sStg = reinterpret_cast<float*>(smem + S::kStages * S::kStageBytes);
  int* meta = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(sStg) + S::kStgBytes);  // [2][BM] output row of each tile row
  float* meta_nz = reinterpret_cast<float*>(meta + 2 * BM);                              // [2][BM] noise sample of each tile row (N2)
  uint64_t* bars = reinterpret_cast<uint64_t*>(meta + 4 * BM);
  // bars: [0..3] full, [4..7] empty, [8..9] t_full, [10..11] t_empty
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long Mtot = (long long)a.n_items * a.a_rows;
  const int kb_per_seg = a.K / BK, num_kb = 2 * kb_per_seg;

  if (threadIdx.x == 0) {
    for (int i = 0; i < S::kStages; ++i) { mbar_init(smem_u32(&bars[i]), 1); mbar_init(smem_u32(&bars[4 + i]), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&bars[8 + i]), 1); mbar_init(smem_u32(&bars[10 + i]), 256); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 2 * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int nt = t % n_tiles, mt = t / n_tiles;
        const int n0 = nt * BN, m0 = mt * BM;  // n0: first row of the (stacked) weight matrix
        const int phase = (nt * CN) / a.Cout, delta = (phase < a.s - a.p) ? -1 : 1;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int st = it % S::kStages;
          mbar_wait(smem_u32(&bars[4 + st]), ((it / S::kStages) & 1) ^ 1);
          const uint32_t full = smem_u32(&bars[st]);
          const uint32_t sa = smem_u32(smem + st * S::kStageBytes), sb = sa + BM * BK * 2;
          const int seg = kb / kb_per_seg, kk = (kb - seg * kb_per_seg) * BK;
          mbar_arrive_expect_tx(full, S::kStageBytes);
          tma_load_2d(sa, &tmA, full, kk, m0 + (seg ? delta : 0));
          tma_load_2d(sb, &tmW, full, seg * a.K + kk, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(BN);
      int it = 0, ti = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++ti) {
        const int buf = ti & 1;
        mbar_wait(smem_u32(&bars[10 + buf]), ((ti >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int st = it % S::kStages;
          mbar_wait(smem_u32(&bars[st]), (it / S::kStages) & 1);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + st * S::kStageBytes), sb = sa + BM * BK * 2;
          const uint64_t da = umma_desc_k_sw128(sa), db = umma_desc_k_sw128(sb);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_f16(tmem_base + buf * BN, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
          umma_commit(smem_u32(&bars[4 + st]));
        }
        umma_commit(smem_u32(&bars[8 + buf]));
      }
    }
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int c4 = lane & 3, r8 = lane >> 2;
    float* stg = sStg + (warp - 2) * (32 * 16);
    float* st_p = stg + lane * 16;
    const int st_x = (lane >> 1) & 3;
    const float* ld_p = stg + r8 * 16 + ((c4 ^ ((r8 >> 1) & 3)) << 2);
    constexpr int NH = BN / 32;
    int ti = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++ti) {
      const int nt = t % n_tiles, mt = t / n_tiles;
      const int c0 = nt * CN, phase = c0 / a.Cout;  // first output channel (phase-major) of the tile
      const int buf = ti & 1;
      int* m_out = meta + buf * BM;
      float* m_nz = meta_nz + buf * BM;
      if (half == 0) {
        const int trow = q * 32 + lane;
        const long long gm = (long long)mt * BM + trow;
        int oi = -1;
        float nz = 0.0f;
        if (gm < Mtot) {
          const int item = (int)(gm / a.a_rows), j = (int)(gm - (long long)item * a.a_rows);
          const ItemRef itr = get_item(a.items, a.base, item, a.out_len);
          const int t_rel = (a.a_lo + j) * a.s + phase;
          const int orow = t_rel - a.o_lo;
          if (orow >= 0 && orow < a.o_rows) {
            const int t_abs = t_rel + itr.shift0 * a.up;
            oi = item * a.o_rows + orow;
            const bool live = (t_abs >= 0) && (t_abs < a.T0 * a.up);
            if (N2 && live) nz = noise_at(a.noise, itr.code_row, t_abs);
            if (!live) oi |= (int)kLiveFlag;
          }
        }
        m_out[trow] = oi;
        if (N2) m_nz[trow] = nz;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      int oi4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) oi4[i] = m_out[q * 32 + r8 + 8 * i];
      const int ocol0 = c0 - phase * a.Cout + half * (CN / 2) + c4 * 4;
      mbar_wait(smem_u32(&bars[8 + buf]), (ti >> 1) & 1);
      tc_fence_after();
      if (N2) {
        // row per lane: x = (D1 + b) + n[row] * (D2 + W_n b), then the usual transpose to 64-byte row segments
        const float2 nz2 = make_float2(m_nz[q * 32 + lane], m_nz[q * 32 + lane]);
        const uint32_t d1 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + half * (CN / 2));
#pragma unroll
        for (int h = 0; h < CN / 32; ++h) {
          uint32_t r1[16], r2[16];
          tmem_ld16_nowait(d1 + (uint32_t)(h * 16), r1);
          tmem_ld16_nowait(d1 + (uint32_t)(CN + h * 16), r2);
          tmem_ld_wait();
          const int colr = c0 - phase * a.Cout + half * (CN / 2) + h * 16;  // this step's first channel (all lanes)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(a.bias + colr + 4 * j));
            const float4 b2 = __ldg(reinterpret_cast<const float4*>(a.bias2 + colr + 4 * j));
            float2 lo = __fadd2_rn(make_float2(__uint_as_float(r1[4 * j]), __uint_as_float(r1[4 * j + 1])), make_float2(b1.x, b1.y));
            float2 hi = __fadd2_rn(make_float2(__uint_as_float(r1[4 * j + 2]), __uint_as_float(r1[4 * j + 3])), make_float2(b1.z, b1.w));
            const float2 nlo = __fadd2_rn(make_float2(__uint_as_float(r2[4 * j]), __uint_as_float(r2[4 * j + 1])), make_float2(b2.x, b2.y));
            const float2 nhi = __fadd2_rn(make_float2(__uint_as_float(r2[4 * j + 2]), __uint_as_float(r2[4 * j + 3])), make_float2(b2.z, b2.w));
            lo = __ffma2_rn(nz2, nlo, lo);
            hi = __ffma2_rn(nz2, nhi, hi);
            *reinterpret_cast<float4*>(st_p + ((j ^ st_x) << 2)) = make_float4(lo.x, lo.y, hi.x, hi.y);
          }
          __syncwarp();
          float4 v[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) v[i] = *reinterpret_cast<const float4*>(ld_p + i * 128);
          __syncwarp();
          const int ocol = ocol0 + h * 16;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            int oi = oi4[i];
            if (oi < 0) continue;
            const bool live = !(oi & (int)kLiveFlag);
            oi &= (int)(kLiveFlag - 1);
            *reinterpret_cast<float4*>(a.out32 + (size_t)oi * a.ldo + ocol) = live ? v[i] : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        tc_fence_before();
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[10 + buf])) : "memory");
        continue;
      }
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        uint32_t r[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + half * (BN / 2) + h * 16), r);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<float4*>(st_p + ((j ^ st_x) << 2)) =
              make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                          __uint_as_float(r[4 * j + 3]));
        __syncwarp();
        float4 v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = *reinterpret_cast<const float4*>(ld_p + i * 128);
        __syncwarp();
        const int ocol = ocol0 + h * 16;
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + ocol));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          int oi = oi4[i];
          if (oi < 0) continue;
          const bool live = !(oi & (int)kLiveFlag);
          oi &= (int)(kLiveFlag - 1);
          float4 x = add4(v[i], b4);
          if (!live) x = make_float4(0.f, 0.f, 0.f, 0.f);
          const size_t o = (size_t)oi * a.ldo + ocol;
          if (a.out32) *reinterpret_cast<float4*>(a.out32 + o) = x;
          if (a.out16) store_half4(a.out16 + o, x);
        }
      }
      tc_fence_before();
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[10 + buf])) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

// ============================================================================ persistent ConvT + NoiseBlock
// Same math as k_convt_noise_tc, one CTA per SM walking (m tile, phase) pairs (phase fastest: the s CTAs
// working on one m tile at a time share its operand rows through L2).  The per-tile latencies of the
// one-shot kernel - launch, barrier / TMEM setup, pipeline fill, the W_n load between the two chains -
// are paid once per SM, and the accumulator pair (D1 conv, D2 noise GEMM) is double-buffered in TMEM so
// the conv MMAs of tile i+1 run under the two epilogue phases of tile i.
// CMP (composed weight, see k_convt_p<BN, true>): the operand box of W is [W_c rows | (W_n W_c) rows] of the phase, ONE
// MMA chain of N = 2 BN fills the accumulator pair, and the kernel loses its first epilogue phase, the y tile, the W_n
// tile and the second MMA chain: what remains of the epilogue is the final x = (D1 + b) + n (D2 + W_n b).
template <int BN, bool CMP = false> struct CnpSmem {
  static constexpr int kStages = (BN == 64) ? (CMP ? 5 : 6) : 3;
  static constexpr int kStageBytes = BM * BK * 2 + (CMP ? 2 : 1) * BN * BK * 2;
  static constexpr int kWnBytes = CMP ? 0 : BN * BN * 2;
  static constexpr int kStgBytes = 8 * 32 * 16 * 4;
  static constexpr int kYBytes = CMP ? kStgBytes : BM * BN * 2;  // fp16 y tile, one per epilogue set; the set's phase-2 transposes alias it
  static constexpr int kMetaBytes = 4 * BM * 4;         // [2][BM] output row + [2][BM] noise sample of each tile row
  static constexpr int kBytes = kStages * kStageBytes + kWnBytes + 2 * kYBytes + kMetaBytes + 2 * BN * 4 + 256 + 1024;
  static_assert(kYBytes >= kStgBytes, "staging aliases the y tile");
  static_assert(kBytes <= 227 * 1024, "shared memory");
};
// warp 0 TMA, warp 1 MMA, warps 2..9 / 10..17 two epilogue sets (even / odd tiles of this CTA, one TMEM accumulator
// pair each), warps 18..21 output row + Philox noise of upcoming tiles, one row per thread
constexpr int kCnpThreads = 64 + 2 * 256 + 128;

template <int BN, bool CMP>
__global__ void __launch_bounds__(kCnpThreads, 1) k_convt_noise_p(const __grid_constant__ CUtensorMap tmA,
                                                               const __grid_constant__ CUtensorMap tmW,
                                                               const __grid_constant__ CUtensorMap tmN, const TcDev a,
                                                               const int total_tiles) {
  using S = CnpSmem<BN, CMP>;
  constexpr int NS = S::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* sWn = smem + NS * S::kStageBytes;          // [KB2][BN rows][128 B]
  uint8_t* sY = sWn + S::kWnBytes;                    // [2 sets][KB2][128 rows][128 B]
  int* meta = reinterpret_cast<int*>(sY + 2 * S::kYBytes);  // [2][BM] output row of each tile row
  float* meta_nz = reinterpret_cast<float*>(meta + 2 * BM);                              // [2][BM] noise sample of each tile row
  float* sBias = meta_nz + 2 * BM;                                                       // [BN] conv bias (same for every phase), [BN] W_n b (CMP)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + 2 * BN);
  // bars: [0..NS) full, [NS..2NS) empty, then c1_full[2], c2_full[2], acc_empty[2], y_ready[2], wn_full, meta_full[2], meta_empty[2]
  uint64_t* bx = bars + 2 * NS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bx + 13);
  const int n_my = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // tiles of this CTA
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long Mtot = (long long)a.n_items * a.a_rows;
  const int kb_per_seg = a.K / BK, num_kb = 2 * kb_per_seg;
  constexpr int KB2 = BN / BK;  // k-blocks of the noise GEMM (K = Cout)

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(smem_u32(&bars[i]), 1); mbar_init(smem_u32(&bars[NS + i]), 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bx[i]), 1); mbar_init(smem_u32(&bx[2 + i]), 1); mbar_init(smem_u32(&bx[4 + i]), 256);
    }
    mbar_init(smem_u32(&bx[6]), 1); mbar_init(smem_u32(&bx[7]), 1); mbar_init(smem_u32(&bx[8]), 1);
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&bx[9 + i]), 128); mbar_init(smem_u32(&bx[11 + i]), 256); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 4 * BN);
  if (threadIdx.x >= 64 && threadIdx.x < 64 + BN) sBias[threadIdx.x - 64] = a.bias[threadIdx.x - 64];
  if (CMP && threadIdx.x >= 64 + BN && threadIdx.x < 64 + 2 * BN) sBias[threadIdx.x - 64] = a.bias2[threadIdx.x - 64 - BN];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      if (!CMP) {
        mbar_arrive_expect_tx(smem_u32(&bx[8]), S::kWnBytes);
#pragma unroll
        for (int kb = 0; kb < KB2; ++kb) tma_load_2d(smem_u32(sWn + kb * BN * 128), &tmN, smem_u32(&bx[8]), kb * BK, 0);
      }
      int it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int phase = t % a.s, m0 = (t / a.s) * BM;
        const int delta = (phase < a.s - a.p) ? -1 : 1;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int st = it % NS;
          mbar_wait(smem_u32(&bars[NS + st]), ((it / NS) & 1) ^ 1);
          const uint32_t full = smem_u32(&bars[st]);
          const uint32_t sa = smem_u32(smem + st * S::kStageBytes), sb = sa + BM * BK * 2;
          const int seg = kb / kb_per_seg, kk = (kb - seg * kb_per_seg) * BK;
          mbar_arrive_expect_tx(full, S::kStageBytes);
          tma_load_2d(sa, &tmA, full, kk, m0 + (seg ? delta : 0));
          tma_load_2d(sb, &tmW, full, seg * a.K + kk, phase * (CMP ? 2 * BN : BN));
        }
      }
    }
  } else if (warp == 1) {
    if (CMP && lane == 0) {
      // composed weight: one chain of N = 2 BN fills [D1 | D2]
      constexpr uint32_t idesc2 = umma_idesc_f16(2 * BN);
      int it = 0;
      for (int ti = 0; ti < n_my; ++ti) {
        const int buf = ti & 1;
        mbar_wait(smem_u32(&bx[4 + buf]), ((ti >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int st = it % NS;
          mbar_wait(smem_u32(&bars[st]), (it / NS) & 1);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + st * S::kStageBytes), sb = sa + BM * BK * 2;
          const uint64_t da = umma_desc_k_sw128(sa), db = umma_desc_k_sw128(sb);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_f16(tmem_base + buf * 2 * BN, da + 2 * k, db + 2 * k, idesc2, (kb | k) ? 1u : 0u);
          umma_commit(smem_u32(&bars[NS + st]));
        }
        umma_commit(smem_u32(&bx[2 + buf]));
      }
    } else if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(BN);
      int it = 0;
      // The noise GEMM of tile i is two to eight MMAs that the epilogue warps are waiting for, the conv chain of
      // tile i+1 is long and paced by TMA: the short chain is issued the moment its operand tile is ready, between
      // two k-blocks of the long one, instead of queueing behind it.
      int c2_next = 0;  // next tile whose noise GEMM has not been issued
      auto noise_chain = [&](bool block) -> bool {
        const int buf = c2_next & 1;
        const uint32_t bar_y = smem_u32(&bx[6 + buf]), par = (uint32_t)(c2_next >> 1) & 1u;
        if (block) mbar_wait(bar_y, par);
        else if (!mbar_test(bar_y, par)) return false;
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < KB2; ++kb) {
          const uint64_t da = umma_desc_k_sw128(smem_u32(sY + buf * S::kYBytes + kb * BM * 128));
          const uint64_t db = umma_desc_k_sw128(smem_u32(sWn + kb * BN * 128));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_f16(tmem_base + buf * 2 * BN + BN, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
        }
        umma_commit(smem_u32(&bx[2 + buf]));
        ++c2_next;
        return true;
      };
      mbar_wait(smem_u32(&bx[8]), 0);
      for (int ti = 0; ti < n_my; ++ti) {
        const int buf = ti & 1;
        while (c2_next + 2 <= ti) noise_chain(true);  // the accumulators of this buffer are drained only after tile ti-2 finished
        mbar_wait(smem_u32(&bx[4 + buf]), ((ti >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int st = it % NS;
          uint32_t spins = 0;
          while (!mbar_test(smem_u32(&bars[st]), (it / NS) & 1)) {
            if (c2_next < ti) noise_chain(false);
            if (++spins > (1u << 26)) { printf("snacb: k_convt_noise_p ring timeout\n"); __trap(); }
          }
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + st * S::kStageBytes), sb = sa + BM * BK * 2;
          const uint64_t da = umma_desc_k_sw128(sa), db = umma_desc_k_sw128(sb);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_f16(tmem_base + buf * 2 * BN, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
          umma_commit(smem_u32(&bars[NS + st]));
          if (c2_next < ti) noise_chain(false);
        }
        umma_commit(smem_u32(&bx[buf]));
      }
      while (c2_next < n_my) noise_chain(true);
    }
  } else if (warp >= 18) {
    // ---- row bookkeeping ahead of the epilogue: output row, live flag and the Philox noise sample
    const int trow = (warp - 18) * 32 + lane;
    int ti = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++ti) {
      const int phase = t % a.s, m0 = (t / a.s) * BM;
      const int buf = ti & 1;
      int my_oi = -1;
      float my_nz = 0.0f;
      const long long gm = (long long)m0 + trow;
      if (gm < Mtot) {
        const int item = (int)(gm / a.a_rows), j = (int)(gm - (long long)item * a.a_rows);
        const ItemRef itr = get_item(a.items, a.base, item, a.out_len);
        const int t_rel = (a.a_lo + j) * a.s + phase;
        const int orow = t_rel - a.o_lo;
        if (orow >= 0 && orow < a.o_rows) {
          const int t_abs = t_rel + itr.shift0 * a.up;
          const bool live = (t_abs >= 0) && (t_abs < a.T0 * a.up);
          my_oi = item * a.o_rows + orow;
          if (live) my_nz = noise_at(a.noise, itr.code_row, t_abs);
          else my_oi |= (int)kLiveFlag;
        }
      }
      mbar_wait(smem_u32(&bx[11 + buf]), ((ti >> 1) & 1) ^ 1);  // the epilogue has read this buffer's previous tile
      meta[buf * BM + trow] = my_oi;
      meta_nz[buf * BM + trow] = my_nz;
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bx[9 + buf])) : "memory");
    }
  } else {
    // set 0 (warps 2..9) takes this CTA's even tiles, set 1 (warps 10..17) the odd ones: tile i+1 goes through its
    // two epilogue phases while tile i is still in its own, each set on its own accumulator pair and y tile
    const int set = (warp - 2) >> 3, ew = (warp - 2) & 7;
    const int q = warp & 3, half = ew >> 2;
    const int trow = q * 32 + lane;  // TMEM lane == tile row owned by this thread in the row-per-lane phases
    const int c4 = lane & 3, r8 = lane >> 2;
    constexpr int NH = BN / 32;  // 16-column steps per warp
    const int colbase = half * (BN / 2);
    const int buf = set;
    uint8_t* sYs = sY + set * S::kYBytes;
    float* stg = reinterpret_cast<float*>(sYs) + ew * (32 * 16);
    const int* m_out = meta + buf * BM;
    for (int ti = set; ti < n_my; ti += 2) {
      mbar_wait(smem_u32(&bx[9 + buf]), (ti >> 1) & 1);
      const int my_oi = m_out[trow];
      const float my_nz = meta_nz[buf * BM + trow];
      const uint32_t d1_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 2 * BN + colbase);
      if (!CMP) {
      mbar_wait(smem_u32(&bx[buf]), (ti >> 1) & 1);
      tc_fence_after();
      asm volatile("bar.sync %0, 256;" ::"r"(1 + set) : "memory");  // the set is done with its transposes (they alias the y tile)
      // ---- phase 1: y = D1 + b -> fp16 -> operand tile of the noise GEMM (row per lane is exactly the K-major layout)
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        const int col = colbase + h * 16;
        uint32_t r[16];
        tmem_ld16(d1_addr + (uint32_t)(h * 16), r);
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 b4 = *reinterpret_cast<const float4*>(sBias + col + 4 * j);
          const float2 lo = __fadd2_rn(make_float2(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1])), make_float2(b4.x, b4.y));
          const float2 hi = __fadd2_rn(make_float2(__uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])), make_float2(b4.z, b4.w));
          const __half2 h0 = f2h2_sat(lo.x, lo.y), h1 = f2h2_sat(hi.x, hi.y);
          pk[2 * j] = *reinterpret_cast<const uint32_t*>(&h0);
          pk[2 * j + 1] = *reinterpret_cast<const uint32_t*>(&h1);
        }
        uint8_t* rowp = sYs + (col / BK) * (BM * 128) + trow * 128;
        const int ch = ((col % BK) * 2) >> 4;  // first of the two 16-byte chunks
        *reinterpret_cast<uint4*>(rowp + (((ch) ^ (trow & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(rowp + (((ch + 1) ^ (trow & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tc_fence_before();
      asm volatile("bar.sync %0, 256;" ::"r"(1 + set) : "memory");
      if (ew == 0 && lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bx[6 + buf])) : "memory");
      }  // !CMP
      // ---- phase 2: x = (D1 + b) + n * D2, transposed through the staging tile, stored coalesced
      int oi4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) oi4[i] = m_out[q * 32 + r8 + 8 * i];
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bx[11 + buf])) : "memory");
      const bool dead = (my_oi >= 0) && (my_oi & (int)kLiveFlag);  // row exists but lies outside the stream: stored as zeros
      const bool any_dead = __any_sync(0xffffffffu, dead);
      const bool all_rows = __all_sync(0xffffffffu, (oi4[0] | oi4[1] | oi4[2] | oi4[3]) >= 0);
      const float2 nz2 = make_float2(my_nz, my_nz);
      mbar_wait(smem_u32(&bx[2 + buf]), (ti >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        const int col = colbase + h * 16;
        uint32_t d1[16], d2[16];
        tmem_ld16_nowait(d1_addr + (uint32_t)(h * 16), d1);
        tmem_ld16_nowait(d1_addr + (uint32_t)(BN + h * 16), d2);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 b4 = *reinterpret_cast<const float4*>(sBias + col + 4 * j);
          float2 lo = __fadd2_rn(make_float2(__uint_as_float(d1[4 * j]), __uint_as_float(d1[4 * j + 1])), make_float2(b4.x, b4.y));
          float2 hi = __fadd2_rn(make_float2(__uint_as_float(d1[4 * j + 2]), __uint_as_float(d1[4 * j + 3])), make_float2(b4.z, b4.w));
          float2 n_lo = make_float2(__uint_as_float(d2[4 * j]), __uint_as_float(d2[4 * j + 1]));
          float2 n_hi = make_float2(__uint_as_float(d2[4 * j + 2]), __uint_as_float(d2[4 * j + 3]));
          if (CMP) {  // the composed GEMM lacks the conv bias's way through W_n
            const float4 c4b = *reinterpret_cast<const float4*>(sBias + BN + col + 4 * j);
            n_lo = __fadd2_rn(n_lo, make_float2(c4b.x, c4b.y));
            n_hi = __fadd2_rn(n_hi, make_float2(c4b.z, c4b.w));
          }
          lo = __ffma2_rn(nz2, n_lo, lo);
          hi = __ffma2_rn(nz2, n_hi, hi);
          d1[4 * j] = __float_as_uint(lo.x); d1[4 * j + 1] = __float_as_uint(lo.y);
          d1[4 * j + 2] = __float_as_uint(hi.x); d1[4 * j + 3] = __float_as_uint(hi.y);
        }
        if (any_dead) {
          if (dead) {
#pragma unroll
            for (int e = 0; e < 16; ++e) d1[e] = 0u;
          }
        }
        float4 v[4];
        epi_transpose16(stg, lane, d1, v);
        float* op = a.out32 + col + c4 * 4;
        if (all_rows) {
#pragma unroll
          for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(op + (size_t)(oi4[i] & (int)(kLiveFlag - 1)) * a.ldo) = v[i];
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (oi4[i] < 0) continue;
            *reinterpret_cast<float4*>(op + (size_t)(oi4[i] & (int)(kLiveFlag - 1)) * a.ldo) = v[i];
          }
        }
      }
      tc_fence_before();
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bx[4 + buf])) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 4 * BN);
  }
}

}  // namespace
int sm_count() {  // of the current device (an engine may live on any ordinal)
  static int n_dev[kMaxDev] = {};
  int& n = n_dev[cur_dev()];
  if (!n) {
    int dev = 0, v = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n = v;
  }
  return n;
}
namespace {

template <int EPI, int CL>
cudaError_t launch_ws_t(const CUtensorMap& ma, const CUtensorMap& mw, const TcDev& d, int n_slices, cudaStream_t st) {
  static int max_p_dev[kMaxDev] = {};  // co-resident clusters (CL > 1) / CTAs per slice (CL == 1), per device
  int& max_p = max_p_dev[cur_dev()];
  if (!max_p) {
    cudaError_t e = cudaFuncSetAttribute(k_gemm_ws<EPI, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, WsSmem::bytes(512));
    if (e != cudaSuccess) return e;
    if (CL > 1) {
      cudaLaunchConfig_t q{};
      q.gridDim = dim3(CL * 64); q.blockDim = dim3(kTcThreads); q.dynamicSmemBytes = WsSmem::bytes(512);
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      q.attrs = at; q.numAttrs = 1;
      int n = 0;
      e = cudaOccupancyMaxActiveClusters(&n, k_gemm_ws<EPI, CL>, &q);
      if (e != cudaSuccess || n < 1) { cudaGetLastError(); return cudaErrorNotSupported; }
      max_p = n;
    } else {
      max_p = std::max(1, sm_count() / std::max(1, n_slices));
    }
  }
  const int P = (CL > 1) ? max_p : std::max(1, sm_count() / n_slices);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(n_slices * P)); cfg.blockDim = dim3(kTcThreads); cfg.dynamicSmemBytes = WsSmem::bytes(d.K); cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = (CL > 1) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, k_gemm_ws<EPI, CL>, ma, mw, d);
}
template <int EPI>
cudaError_t launch_ws_e(const CUtensorMap& ma, const CUtensorMap& mw, const TcDev& d, int n_slices, cudaStream_t st) {
  if (g_ws_cluster && n_slices == 4) {
    cudaError_t e = launch_ws_t<EPI, 4>(ma, mw, d, n_slices, st);
    if (e != cudaErrorNotSupported) return e;
  }
  if (g_ws_cluster && n_slices == 2) {
    cudaError_t e = launch_ws_t<EPI, 2>(ma, mw, d, n_slices, st);
    if (e != cudaErrorNotSupported) return e;
  }
  return launch_ws_t<EPI, 1>(ma, mw, d, n_slices, st);
}

template <int BN>
cudaError_t launch_cnp_t(const CUtensorMap& ma, const CUtensorMap& mw, const CUtensorMap& mn, const TcDev& d, int total,
                         cudaStream_t st) {
  static bool attr_set_dev[kMaxDev] = {};
  bool& attr_set = attr_set_dev[cur_dev()];  // kernel attributes are per device
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_convt_noise_p<BN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CnpSmem<BN>::kBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  k_convt_noise_p<BN, false><<<std::min(total, sm_count()), kCnpThreads, CnpSmem<BN>::kBytes, st>>>(ma, mw, mn, d, total);
  return cudaGetLastError();
}
template <int BN>
cudaError_t launch_cnp2_t(const CUtensorMap& ma, const CUtensorMap& mw, const TcDev& d, int total, cudaStream_t st) {
  static bool attr_set_dev[kMaxDev] = {};
  bool& attr_set = attr_set_dev[cur_dev()];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_convt_noise_p<BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CnpSmem<BN, true>::kBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  k_convt_noise_p<BN, true><<<std::min(total, sm_count()), kCnpThreads, CnpSmem<BN, true>::kBytes, st>>>(ma, mw, mw, d, total);
  return cudaGetLastError();
}

template <int BN>
cudaError_t launch_cn_t(const CUtensorMap& ma, const CUtensorMap& mw, const CUtensorMap& mn, const TcDev& d, dim3 grid,
                        cudaStream_t st) {
  static bool attr_set_dev[kMaxDev] = {};
  bool& attr_set = attr_set_dev[cur_dev()];  // kernel attributes are per device
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_convt_noise_tc<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         TcSmem<BN>::bytes(TcSmem<BN>::kMaxStages));
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  k_convt_noise_tc<BN><<<grid, kTcThreads, TcSmem<BN>::bytes(d.stages), st>>>(ma, mw, mn, d);
  return cudaGetLastError();
}

}  // namespace

bool convt_noise_supported(int Cin, int Cout) { return (Cout == 64 || Cout == 128) && (2 * Cin / BK) >= 3; }

// a: the EPI_CONVT arguments (out32 = x, out16 unused); noise_w16 [Cout][Cout] fp16; a.noise set.
cudaError_t launch_convt_noise_tc(const GroupCtx& g, const TcGemmArgs& a, const __half* noise_w16) {
  const long long Mtot = (long long)g.n_items * a.a_rows;
  if (Mtot <= 0) return cudaSuccess;
  const int bn = a.Cout;
  if (!convt_noise_supported(a.K, a.Cout) || a.N != a.s * a.Cout || !a.out32) return cudaErrorInvalidValue;
  CUtensorMap ma, mw, mn;
  if (!get_tmap(a.A, Mtot, a.K, BM, &ma) || !get_tmap(a.W, a.N, 2 * a.K, bn, &mw) || !get_tmap(noise_w16, bn, bn, bn, &mn))
    return cudaErrorNotSupported;
  TcDev d{};
  d.items = g.items; d.base = g.base; d.out_len = g.out_len; d.T0 = g.T0; d.n_items = g.n_items;
  d.stages = 3;  // stage 0 / 1 / 2 are re-used as y tile / W_n / transposes; 3 stages keep 2-3 CTAs per SM
  d.K = a.K; d.nseg = 2; d.a_rows = a.a_rows; d.a_lo = a.a_lo; d.s = a.s; d.p = a.p; d.Cout = a.Cout;
  d.bias = a.bias; d.out32 = a.out32; d.o_lo = a.o_r.lo; d.o_rows = a.o_r.n(); d.ldo = a.ldo; d.noise = a.noise; d.up = a.up;
  const long long total = ((Mtot + BM - 1) / BM) * a.s;
  cudaError_t e;
  if (g_cn_p && !(g.flags & SNACB_FLAG_NO_PERSISTENT_CONVT) && total >= 2LL * sm_count() && total < (1LL << 30)) {
    e = (bn == 128) ? launch_cnp_t<128>(ma, mw, mn, d, (int)total, g.stream) : launch_cnp_t<64>(ma, mw, mn, d, (int)total, g.stream);
  } else {
    dim3 grid((unsigned)total);
    e = (bn == 128) ? launch_cn_t<128>(ma, mw, mn, d, grid, g.stream) : launch_cn_t<64>(ma, mw, mn, d, grid, g.stream);
  }
  ++*g.launches;
  return e;
}

int tc_tile_n(const TcGemmArgs& a) {
  const int per = (a.epi == EPI_CONVT) ? a.Cout : a.N;
  return per >= 128 ? 128 : 64;
}

cudaError_t launch_gemm_tc(const GroupCtx& g, const TcGemmArgs& a) {
  const long long Mtot = (long long)g.n_items * a.a_rows;
  if (Mtot <= 0) return cudaSuccess;
  const int nseg = ((a.epi == EPI_CONVT) ? 2 : 1) * (a.split ? 3 : 1);
  const int bn = tc_tile_n(a);
  if (a.K % BK || a.N % bn || (a.epi == EPI_CONVT && a.Cout % bn) || (a.split && a.out16)) return cudaErrorInvalidValue;
  CUtensorMap ma, mw;
  if (!get_tmap(a.A, Mtot, a.split ? 2 * a.K : a.K, BM, &ma) || !get_tmap(a.W, a.N, nseg * a.K, bn, &mw)) return cudaErrorNotSupported;
  // wide square 1x1 layers (blocks 0 / 1): persistent weight-stationary kernel
  const bool ws = !a.split && !g_no_ws && (a.epi == EPI_RESID || a.epi == EPI_NOISE) && a.N == a.ldo && a.N == a.K && (a.K == 256 || a.K == 512) &&
                  Mtot >= 4 * BM * (long long)(sm_count() / (a.N / 128));
  TcDev d{};
  d.items = g.items; d.base = g.base; d.out_len = g.out_len; d.T0 = g.T0; d.n_items = g.n_items;
  d.stages = std::min((a.K / BK) * nseg, bn == 128 ? TcSmem<128>::kMaxStages : TcSmem<64>::kMaxStages);
  d.K = a.K; d.nseg = nseg; d.a_rows = a.a_rows; d.a_lo = a.a_lo; d.s = a.s; d.p = a.p; d.Cout = a.Cout;
  d.bias = a.bias; d.out32 = a.out32; d.out16 = a.out16; d.o_lo = a.o_r.lo; d.o_rows = a.o_r.n(); d.ldo = a.ldo;
  d.sn_alpha = a.sn_alpha; d.sn_inv = a.sn_inv; d.R = a.R; d.r_lo = a.r_r.lo; d.r_rows = a.r_r.n(); d.ldr = a.ldr;
  d.noise = a.noise; d.up = a.up; d.split = a.split;
  if (ws) {
    const int n_slices = a.N / 128;
    cudaError_t e = (a.epi == EPI_RESID) ? launch_ws_e<EPI_RESID>(ma, mw, d, n_slices, g.stream)
                                         : launch_ws_e<EPI_NOISE>(ma, mw, d, n_slices, g.stream);
    ++*g.launches;
    return e;
  }
  if (!a.split && g_convt_p && !(g.flags & SNACB_FLAG_NO_PERSISTENT_CONVT) && a.epi == EPI_CONVT && a.Cout % 256 == 0 && !a.sn_alpha) {
    // wide transposed convs: persistent kernel with 128 x 256 tiles (87 FLOP per byte of L2 -> SM operand traffic
    // instead of 65 at 128 x 128, which caps those layers near 770 TFLOP/s)
    const int n_tiles = a.N / 256, m_tiles = (int)((Mtot + BM - 1) / BM);
    const long long total = (long long)n_tiles * m_tiles;
    CUtensorMap mw256;
    if (total >= 2LL * sm_count() && get_tmap(a.W, a.N, nseg * a.K, 256, &mw256)) {
      static bool attr_set_dev[kMaxDev] = {};
  bool& attr_set = attr_set_dev[cur_dev()];  // kernel attributes are per device
      if (!attr_set) {
        cudaError_t e0 = cudaFuncSetAttribute(k_convt_p<256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CtpSmem<256>::kBytes);
        if (e0 != cudaSuccess) return e0;
        attr_set = true;
      }
      k_convt_p<256, false><<<sm_count(), kTcThreads, CtpSmem<256>::kBytes, g.stream>>>(ma, mw256, d, n_tiles, (int)total);
      ++*g.launches;
      return cudaGetLastError();
    }
  }
  d.n_tiles = a.N / bn;
  dim3 grid((unsigned)(((Mtot + BM - 1) / BM) * d.n_tiles));
  cudaError_t e = (bn == 128) ? launch_tc_bn<128>(a.epi, ma, mw, d, grid, g.stream) : launch_tc_bn<64>(a.epi, ma, mw, d, grid, g.stream);
  ++*g.launches;
  return e;
}

// ---- composed ConvT + NoiseBlock (blocks 0 / 1), see k_convt_p<BN, true>
namespace {
// Stacked fp16 weight [2N][K2] from the packed fp32 conv weight ct [N = s * Cout][K2 = 2 Cin] and the NoiseBlock weight
// wn [Cout][Cout]: per slice i of `slice` = min(Cout, 128) rows, rows [2 slice i, 2 slice i + slice) = ct rows [slice i, ..),
// the next `slice` rows = the matching (W_n ct_phase) rows, (W_n ct_phase)[o][k] = sum_c wn[o][c] * ct[phase * Cout + c][k].  Runs once per weight load.
__global__ void __launch_bounds__(256) k_compose_ctn(const float* __restrict__ ct, const float* __restrict__ wn, int Cout, int N, int K2,
                                                     int slice, __half* __restrict__ out) {
  __shared__ float s_wn[8][64];  // 8 output rows x 64 contraction steps
  const int k = blockIdx.x * 32 + (threadIdx.x & 31), ty = threadIdx.x >> 5;  // 32 columns x 8 rows per block
  const int n = blockIdx.y * 8 + ty;
  const int phase = n / Cout, o = n - phase * Cout;  // a block's 8 rows share the phase (Cout % 8 == 0)
  float acc = 0.0f;
  for (int c0 = 0; c0 < Cout; c0 += 64) {
    for (int e = threadIdx.x; e < 8 * 64; e += 256) {
      const int rr = e >> 6, cc = e & 63;
      s_wn[rr][cc] = wn[(size_t)(blockIdx.y * 8 + rr - phase * Cout) * Cout + c0 + cc];
    }
    __syncthreads();
    if (k < K2) {
#pragma unroll 8
      for (int c = 0; c < 64; ++c) acc = fmaf(s_wn[ty][c], ct[(size_t)(phase * Cout + c0 + c) * K2 + k], acc);
    }
    __syncthreads();
  }
  if (k < K2) {
    const size_t row = (size_t)(n / slice) * 2 * slice + (n % slice);
    out[row * K2 + k] = __float2half_rn(ct[(size_t)n * K2 + k]);
    out[(row + slice) * K2 + k] = __float2half_rn(acc);
  }
  (void)o;
}

}  // namespace
void launch_compose_ctn(const float* ct, const float* wn, int Cout, int N, int K2, __half* out, cudaStream_t st) {
  dim3 grid((unsigned)((K2 + 31) / 32), (unsigned)(N / 8));
  k_compose_ctn<<<grid, 256, 0, st>>>(ct, wn, Cout, N, K2, std::min(Cout, 128), out);
}

// Which blocks take the composed kernel: SNACB_CONVT_N2 = bit mask over the decoder blocks (default 10 = blocks 1 and 3).
// Block 1 (Cin 512 -> Cout 256): its conv leaves the tensor pipe 54 % busy, so the doubled MMA work is cheaper than the
// separate noise GEMM (216 + 185 -> 316 us per 1024-window tick).  Block 3 (128 -> 64, k_convt_noise_p<64, composed>):
// 258 -> 205 us, the kernel loses its first epilogue phase and the second MMA chain.  Block 0's conv is already 70 %
// tensor-busy (380 us composed against 221 + 140) and block 2's composed operand stream (48 KB per k-block, three stages)
// is slower than the two-chain kernel (291 against 267 us): both keep their round-1 kernels.
bool convt_noise2_supported(int Cin, int Cout) {
  static const int mask = [] { const char* v = getenv("SNACB_CONVT_N2"); return v ? atoi(v) : 10; }();
  const int b = (Cout == 512) ? 0 : (Cout == 256) ? 1 : (Cout == 128) ? 2 : (Cout == 64) ? 3 : -1;
  return b >= 0 && ((mask >> b) & 1) && Cin == 2 * Cout && Cin % BK == 0;
}

namespace {
template <int BN>
cudaError_t launch_ctn2_t(const CUtensorMap& ma, const CUtensorMap& mw, const TcDev& d, int n_tiles, long long total, cudaStream_t st) {
  static bool attr_set_dev[kMaxDev] = {};
  bool& attr_set = attr_set_dev[cur_dev()];
  if (!attr_set) {
    cudaError_t e0 = cudaFuncSetAttribute(k_convt_p<BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CtpSmem<BN>::kBytes);
    if (e0 != cudaSuccess) return e0;
    attr_set = true;
  }
  const int grid = (int)std::min<long long>(total, sm_count());
  k_convt_p<BN, true><<<grid, kTcThreads, CtpSmem<BN>::kBytes, st>>>(ma, mw, d, n_tiles, (int)total);
  return cudaGetLastError();
}
}  // namespace

// a: the EPI_CONVT arguments of the layer (fp32 output only); W2: the stacked weight [2 N][2 K]; bias2 = W_n b.
cudaError_t launch_convt_noise2_tc(const GroupCtx& g, const TcGemmArgs& a, const __half* W2, const float* bias2) {
  if (a.epi != EPI_CONVT || !a.out32 || a.out16 || a.sn_alpha || a.split || !W2 || !bias2 || a.K != 2 * a.Cout || a.K % BK ||
      a.N != a.s * a.Cout || g.n_items <= 0 || a.a_rows <= 0 || (a.Cout != 512 && a.Cout != 256 && a.Cout != 128 && a.Cout != 64))
    return cudaErrorInvalidValue;
  const long long Mtot = (long long)g.n_items * a.a_rows;
  const int CN = std::min(a.Cout, 128), BN2 = 2 * CN;  // output channels / accumulator columns per tile
  CUtensorMap ma, mw;
  if (!get_tmap(a.A, Mtot, a.K, BM, &ma) || !get_tmap(W2, 2LL * a.N, 2 * a.K, BN2, &mw)) return cudaErrorNotSupported;
  TcDev d{};
  d.items = g.items; d.base = g.base; d.out_len = g.out_len; d.T0 = g.T0; d.n_items = g.n_items;
  d.K = a.K; d.nseg = 2; d.a_rows = a.a_rows; d.a_lo = a.a_lo; d.s = a.s; d.p = a.p; d.Cout = a.Cout;
  d.bias = a.bias; d.bias2 = bias2; d.out32 = a.out32; d.o_lo = a.o_r.lo; d.o_rows = a.o_r.n(); d.ldo = a.ldo;
  d.noise = a.noise; d.up = a.up;
  const int n_tiles = a.N / CN, m_tiles = (int)((Mtot + BM - 1) / BM);
  const long long total = (long long)n_tiles * m_tiles;
  if (total >= (1LL << 31)) return cudaErrorInvalidValue;
  cudaError_t e;
  if (a.Cout <= 128) {
    // narrow blocks: the two-epilogue-set kernel (one tile = one phase of one m tile, all Cout channels)
    const long long tot = (long long)m_tiles * a.s;
    if (tot >= (1LL << 31)) return cudaErrorInvalidValue;
    e = (a.Cout == 128) ? launch_cnp2_t<128>(ma, mw, d, (int)tot, g.stream) : launch_cnp2_t<64>(ma, mw, d, (int)tot, g.stream);
  } else {
    e = (BN2 == 256) ? launch_ctn2_t<256>(ma, mw, d, n_tiles, total, g.stream) : launch_ctn2_t<128>(ma, mw, d, n_tiles, total, g.stream);
  }
  ++*g.launches;
  return e;
}

// ============================================================================ depthwise k=7 -> fp16
// Each thread owns two adjacent channels and walks one residue class (mod dil) of a 32*dil-row
// segment of 16 outputs (22 inputs, staged through shared memory by cp.async), every input Snake'd once.  Output = Snake(b + sum_k w[k] * Snake(x[t + (k-3)*dil])) as fp16, the
// K-major operand of the ResidualUnit's 1x1 GEMM.
template <int C, int DIL>
__global__ void __launch_bounds__(256) k_dw_tc(const Item* items, int base, int out_len, int T0, DwTcArgs a) {
  constexpr int CP = C / 2;
  const int item = blockIdx.y;
  const int idx = blockIdx.x * 256 + threadIdx.x;
  const int cp = idx % CP, rest = idx / CP;
  const int rho = rest % DIL, seg = rest / DIL;
  const int out_rows = a.out_r.n(), in_rows = a.in_r.n();
  __shared__ float2 stage[22 * 256];
  const int row0 = seg * 16 * DIL + rho;
  if (row0 >= out_rows) return;
  const int c = cp * 2;
  const ItemRef it = get_item(items, base, item, out_len);
  DwPairW W;
  W.load(a.w7, a.bias, a.a1, a.i1, a.a2, a.i2, C, c);
  const float* x = a.in + (size_t)item * in_rows * C + c;
  __half* o = a.out + (size_t)item * out_rows * C + c;
  const int t_first = a.out_r.lo + row0;                 // relative time of this thread's first output
  const int in_first = t_first - 3 * DIL - a.in_r.lo;    // operand row of window element 0
  const int t_hi = T0 * a.up;
  const int t_abs0 = t_first + it.shift0 * a.up;
  uint32_t mlo, mhi;
  row_mask<DIL>(in_first, in_rows, 22, mlo, mhi);
  dw_unit_staged<C, DIL, 16, 0>(x + (long long)in_first * C, mlo, W, stage + threadIdx.x, 256, [&](int j, float2 v) {
    const int orow = row0 + j * DIL;
    if (orow < out_rows) {
      const int t_abs = t_abs0 + j * DIL;
      if (t_abs < 0 || t_abs >= t_hi) v = make_float2(0.f, 0.f);
      *reinterpret_cast<__half2*>(o + (size_t)orow * C) = f2h2_sat(v.x, v.y);
    }
  });
}

template <int C>
void launch_dw_tc_c(const GroupCtx& g, const DwTcArgs& a) {
  const int nseg = (a.out_r.n() + 16 * a.dil - 1) / (16 * a.dil);
  dim3 grid((unsigned)(((long long)nseg * a.dil * (C / 2) + 255) / 256), g.n_items);
  if (a.dil == 1) k_dw_tc<C, 1><<<grid, 256, 0, g.stream>>>(g.items, g.base, g.out_len, g.T0, a);
  else if (a.dil == 3) k_dw_tc<C, 3><<<grid, 256, 0, g.stream>>>(g.items, g.base, g.out_len, g.T0, a);
  else k_dw_tc<C, 9><<<grid, 256, 0, g.stream>>>(g.items, g.base, g.out_len, g.T0, a);
}

bool launch_dw_x(const GroupCtx& g, const DwTcArgs& a);  // TMA-staged variant (defined below), false = not applicable

void launch_dw_tc(const GroupCtx& g, const DwTcArgs& a) {
  if (launch_dw_x(g, a)) return;
  switch (a.C) {
    case 64: launch_dw_tc_c<64>(g, a); break;
    case 128: launch_dw_tc_c<128>(g, a); break;
    case 256: launch_dw_tc_c<256>(g, a); break;
    default: launch_dw_tc_c<512>(g, a); break;
  }
  ++*g.launches;
}

// ============================================================================ fused ResidualUnit
// x' = x + W_pw * Snake(dw7_dil(Snake(x))) + b for one 128-row time tile of one item, all C channels
// (C = 64 or 128: decoder blocks 3 and 2, where the 1x1 weight fits in shared memory whole and the
// layer is bound by activation traffic, not by the GEMM).  Nine worker warps build the fp16 GEMM
// operand straight into the 128-byte-swizzled K-major shared-memory layout tcgen05 reads (register
// sliding windows over global loads: every input Snake'd once per residue class), one control warp
// TMA-loads the weight and issues the MMAs, then eight warps run the epilogue out of TMEM (bias +
// residual + optional next-block Snake) with coalesced stores.  Versus k_dw_tc + k_gemm_tc this
// removes the fp16 operand round trip through HBM and one launch per unit.
namespace {
constexpr int kRuThreads = 320;
constexpr int kRuWorkers = 288;

struct RuDev {
  const Item* items; int base, out_len, T0;
  const float* x; int in_lo, in_rows;   // residual stream in: rows per item, relative time of row 0
  int out_lo, out_rows, up;
  const float* w7; const float* dw_b; const float* a1; const float* i1; const float* a2; const float* i2;
  const float* pw_b;
  float* out32; __half* out16; const float* sn_alpha; const float* sn_inv;
  int prefetch_ahead;   // CTAs resident on the whole GPU: the tile this far ahead in launch order is prefetched into L2
  // fused decoder tail (TAIL variant, last ResidualUnit of block 3): Snake(64) -> conv k7 64->1 -> tanh -> pack
  int tile_stride, row_off;            // tile t covers out rows [t*tile_stride + row_off, +128)
  int tail_lo, tail_n;                 // emitted samples: relative times [tail_lo, tail_lo + tail_n)
  const float* tail_w7; const float* tail_b; int32_t* status; float* wav; int16_t* pcm;
};

constexpr int kTailPitch = 80;  // floats per row of the Snake'd output tile (conflict-free float4 rows)
template <int C> struct RuSmem {
  static constexpr int kABytes = BM * C * 2;   // operand tile; the epilogue staging (16 KB) aliases it after the MMAs
  static constexpr int kWBytes = C * C * 2;
  static constexpr int kInBytes = 0;  // (cp.async input staging measured slower here than register loads: 2.22 vs 2.01 ms)
  static constexpr int kMetaBytes = 64;
  static constexpr int kBytes = kABytes + kWBytes + kInBytes + kMetaBytes + 1024;
  static constexpr int kTailBytes = BM * kTailPitch * 4 + 4 * 64 * 4;  // TAIL variant: output tile + partial sums
  static_assert(kABytes >= 8 * 32 * 16 * 4, "staging aliases the operand tile");
};

template <int C, int DIL, bool TAIL = false>
__global__ void __launch_bounds__(kRuThreads, (C == 64) ? 3 : 2) k_ru_tc(const __grid_constant__ CUtensorMap tmW, const RuDev a) {
  using S = RuSmem<C>;
  constexpr int KB = C / BK;  // k-blocks of 64 channels
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* sA = smem;
  uint8_t* sW = smem + S::kABytes;
  float* sStg = reinterpret_cast<float*>(smem);  // valid once the accumulator barrier has fired
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kABytes + S::kWBytes + S::kInBytes);  // [0] weight landed, [1] accumulator complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int item = blockIdx.y;
  // first output row (index into the item's out rows) of this tile; TAIL tiles overlap by 6 rows and may start
  // before row 0 / run past the last row (those rows are the final conv's zero padding)
  const int row0 = TAIL ? (int)blockIdx.x * a.tile_stride + a.row_off : (int)blockIdx.x * BM;
  const ItemRef it = get_item(a.items, a.base, item, a.out_len);
  float* sTail = reinterpret_cast<float*>(smem + S::kABytes + S::kWBytes + S::kInBytes + S::kMetaBytes);  // TAIL only

  if (tid == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) tmem_alloc(smem_u32(tmem_slot), C);
  if (a.prefetch_ahead > 0 && tid == 64) {
    // Pull the input rows of the tile that will take this CTA's place into L2 now: its loads then cost an
    // L2 hit instead of a DRAM round trip (the kernel streams, nothing is resident at chunk = 1024 items).
    const long long lin = (long long)blockIdx.y * gridDim.x + blockIdx.x + a.prefetch_ahead;
    const int pit = (int)(lin / gridDim.x), ptile = (int)(lin - (long long)pit * gridDim.x);
    if (pit < (int)gridDim.y) {
      const int prow0 = TAIL ? ptile * a.tile_stride + a.row_off : ptile * BM;
      const int r_lo = max(a.out_lo + prow0 - 3 * DIL - a.in_lo, 0);
      const int r_hi = min(a.out_lo + prow0 + BM + 3 * DIL - a.in_lo, a.in_rows);
      const char* p = reinterpret_cast<const char*>(a.x + ((size_t)pit * a.in_rows + r_lo) * C);
      const uint32_t bytes = (uint32_t)(r_hi - r_lo) * C * 4;  // rows are contiguous: one bulk prefetch covers the tile
      if (r_hi > r_lo) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 9) {
    if (lane == 0) {  // 1x1 weight [C][C] fp16 -> KB swizzled k-block tiles of [C rows][128 B]
      mbar_arrive_expect_tx(smem_u32(&bars[0]), S::kWBytes);
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(smem_u32(sW + kb * C * 128), &tmW, smem_u32(&bars[0]), kb * BK, 0);
    }
  } else {
    // ---- depthwise stage: unit = (channel pair, residue class / segment), 16 outputs, 22 loads
    constexpr int CP = C / 2;
    constexpr int U = (DIL == 1) ? 8 : 9;
    const float* xin = a.x + (size_t)item * a.in_rows * C;
#pragma unroll 1
    for (int idx = tid; idx < CP * U; idx += kRuWorkers) {
      const int cp = idx % CP, u = idx / CP;
      int first;  // tile-relative row of the unit's first output
      if (DIL == 1) first = u * 16;
      else if (DIL == 3) first = (u % 3) + (u / 3) * 48;
      else first = u;
      const int c = cp * 2;
      DwPairW W;
      W.load(a.w7, a.dw_b, a.a1, a.i1, a.a2, a.i2, C, c);
      const int in_first = a.out_lo + row0 + first - 3 * DIL - a.in_lo;  // operand row of window element 0
      // K-major SWIZZLE_128B operand tile: byte (row, col) -> row*128 + ((col/16 ^ row%8) * 16) + col%16
      uint8_t* a_kb = sA + (c / BK) * (BM * 128) + ((c % BK) * 2 & 15);
      const int chunk = ((c % BK) * 2) >> 4;
      unit_len_dispatch<DIL>(u, [&](auto len) {
        constexpr int L = decltype(len)::value;
        uint32_t mlo, mhi;
        row_mask<DIL>(in_first, a.in_rows, L + 6, mlo, mhi);
        dw_unit<C, DIL, L>(xin + (long long)in_first * C + c, mlo, mhi, W, [&](int j, float2 v) {
          const int trow = first + j * DIL;  // < 128 by construction of the unit length
          *reinterpret_cast<__half2*>(a_kb + trow * 128 + (((chunk ^ trow) & 7) << 4)) = f2h2_sat(v.x, v.y);
        });
      });
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
  }
  __syncthreads();

  if (warp == 9) {
    if (lane == 0) {
      mbar_wait(smem_u32(&bars[0]), 0);
      tc_fence_after();
      constexpr uint32_t idesc = umma_idesc_f16(C);
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        const uint64_t da = umma_desc_k_sw128(smem_u32(sA + kb * (BM * 128)));
        const uint64_t db = umma_desc_k_sw128(smem_u32(sW + kb * (C * 128)));
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) umma_f16(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
      }
      umma_commit(smem_u32(&bars[1]));
    }
  } else if (warp < 8) {
    // ---- epilogue: warp w reads TMEM lanes 32*(w%4).., the two warps of a quarter split the columns.
    // After the 16-column transpose lane l owns rows (l/4 + 8i), i = 0..3, columns 4*(l%4)..+3: its four
    // rows are an arithmetic progression, so one base pointer + immediates address everything.
    const int q = warp & 3, half = warp >> 2;
    constexpr int NH = C / 32;  // 16-column half-chunks per warp
    const int c4 = lane & 3, r8 = lane >> 2;
    float* stg = sStg + warp * (32 * 16);
    float* st_p = stg + lane * 16;
    const int st_x = (lane >> 1) & 3;
    const float* ld_p = stg + r8 * 16 + ((c4 ^ ((r8 >> 1) & 3)) << 2);
    const int orow_b = row0 + q * 32 + r8;  // first of this lane's four rows (stride 8)
    const int tabs_b = a.out_lo + orow_b + it.shift0 * a.up;
    uint32_t vmask = 0, lmask = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (orow_b + 8 * i >= 0 && orow_b + 8 * i < a.out_rows) {
        vmask |= 1u << i;
        const int t = tabs_b + 8 * i;
        if (t >= 0 && t < a.T0 * a.up) lmask |= 1u << i;
      }
    }
    const int colb = half * (C / 2) + c4 * 4;
    const size_t ob = ((size_t)item * a.out_rows + orow_b) * C + colb;
    const float* rp = a.x + ((size_t)item * a.in_rows + (a.out_lo + orow_b - a.in_lo)) * C + colb;
    // residual rows of this lane, requested one column step ahead (two steps ahead measured slower: register spills)
    constexpr int PD = 1;
    float4 res[PD][4];
    auto load_res = [&](int h) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        res[h % PD][i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((vmask >> i) & 1u) res[h % PD][i] = __ldg(reinterpret_cast<const float4*>(rp + i * 8 * C + h * 16));
      }
    };
#pragma unroll
    for (int h = 0; h < PD && h < NH; ++h) load_res(h);
    mbar_wait(smem_u32(&bars[1]), 0);
    tc_fence_after();
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      uint32_t r[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * (C / 2) + h * 16), r);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<float4*>(st_p + ((j ^ st_x) << 2)) =
            make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                        __uint_as_float(r[4 * j + 3]));
      __syncwarp();
      float4 v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = *reinterpret_cast<const float4*>(ld_p + i * 128);
      __syncwarp();
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.pw_b + colb + h * 16));
      float4 al = make_float4(0.f, 0.f, 0.f, 0.f), iv = al;
      if (a.sn_alpha) {
        al = __ldg(reinterpret_cast<const float4*>(a.sn_alpha + colb + h * 16));
        iv = __ldg(reinterpret_cast<const float4*>(a.sn_inv + colb + h * 16));
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (!((vmask >> i) & 1u)) {
          if (TAIL) *reinterpret_cast<float4*>(sTail + (q * 32 + r8 + 8 * i) * kTailPitch + colb + h * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
          continue;
        }
        float4 x = add4(add4(v[i], b4), res[h % PD][i]);
        if (!((lmask >> i) & 1u)) x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (TAIL) {  // Snake of the decoder tail, kept on chip for the final conv
          const float2 lo = snake2(make_float2(x.x, x.y), make_float2(al.x, al.y), make_float2(iv.x, iv.y));
          const float2 hi = snake2(make_float2(x.z, x.w), make_float2(al.z, al.w), make_float2(iv.z, iv.w));
          *reinterpret_cast<float4*>(sTail + (q * 32 + r8 + 8 * i) * kTailPitch + colb + h * 16) = make_float4(lo.x, lo.y, hi.x, hi.y);
          continue;
        }
        if (a.out32) *reinterpret_cast<float4*>(a.out32 + ob + i * 8 * C + h * 16) = x;
        if (a.out16) {
          if (a.sn_alpha) {
            const float2 lo = snake2(make_float2(x.x, x.y), make_float2(al.x, al.y), make_float2(iv.x, iv.y));
            const float2 hi = snake2(make_float2(x.z, x.w), make_float2(al.z, al.w), make_float2(iv.z, iv.w));
            x = make_float4(lo.x, lo.y, hi.x, hi.y);
          }
          store_half4(a.out16 + ob + i * 8 * C + h * 16, x);
        }
      }
      if (h + PD < NH) load_res(h + PD);
    }
    tc_fence_before();
    if (TAIL) {
      // ---- decoder tail on the tile: y[o] = tanh(b + sum_{k,c} w[k][c] * s[o + k][c]), o = 0..121 (rows o..o+6)
      asm volatile("bar.sync 2, 256;" ::: "memory");  // the eight epilogue warps: tile complete in shared memory
      float* part = sTail + BM * kTailPitch;           // [4][64]
      const int etid = tid;                            // 0..255
      const int sx = etid & 63, qc = etid >> 6;
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {
        const int o = pass * 64 + sx;
        float acc = 0.0f;
        if (o < a.tile_stride) {
#pragma unroll
          for (int k = 0; k < 7; ++k)
#pragma unroll
            for (int c = 0; c < 16; c += 4) {
              const float4 w4 = __ldg(reinterpret_cast<const float4*>(a.tail_w7 + k * 64 + qc * 16 + c));
              const float4 x4 = *reinterpret_cast<const float4*>(sTail + (o + k) * kTailPitch + qc * 16 + c);
              acc = fmaf(w4.x, x4.x, acc); acc = fmaf(w4.y, x4.y, acc); acc = fmaf(w4.z, x4.z, acc); acc = fmaf(w4.w, x4.w, acc);
            }
        }
        part[qc * 64 + sx] = acc;
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (qc == 0 && o < a.tile_stride) {
          const int oi = (int)blockIdx.x * a.tile_stride + o;  // emitted sample index inside [0, tail_n)
          const int t_abs = a.tail_lo + oi + it.shift0 * 512;
          if (oi < a.tail_n && t_abs >= 0 && t_abs < a.T0 * 512 && !(a.status && a.status[it.code_row] != SNACB_WIN_OK)) {
            float y = tanhf(((part[sx] + part[64 + sx]) + (part[128 + sx] + part[192 + sx])) + a.tail_b[0]);
            const long long d = it.dst + oi;
            if (!(fabsf(y) <= 1.0f)) {
              y = 0.0f;
              if (a.status) a.status[it.code_row] = SNACB_WIN_NONFINITE;
            }
            if (a.wav) a.wav[d] = y;
            if (a.pcm) a.pcm[d] = (int16_t)(y * 32767.0f);
          }
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");
      }
    }
  }
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C);
  }
}

// ============================================================================ fused ResidualUnit, TMA-staged input
// k_ru_tc for C = 64 (decoder block 3) with the fp32 input rows of the tile (128 + 6*DIL rows, halo included) brought
// into shared memory by ONE TMA load of a 3-D tensor map [item][row][C] (rows outside the item come back as zeros =
// the conv's zero padding): the depthwise units read them with LDS.64 at immediate offsets - no per-input LDG,
// predicate or address arithmetic, nothing in flight in registers - and the epilogue takes the residual from the
// same tile instead of re-reading global memory.  70-76 KB per CTA, still three CTAs per SM.
template <int DIL> struct RuxSmem {
  static constexpr int C = 64;
  static constexpr int kABytes = BM * C * 2;
  static constexpr int kWBytes = C * C * 2;
  static constexpr int kBoxRows = BM + 6 * DIL;                        // rows the tile's outputs depend on
  static constexpr int kXBytes = kBoxRows * C * 4;  // unit lengths are exact (unit_len_dispatch): no unit reads past the box
  static constexpr int kBytes = kABytes + kWBytes + kXBytes + 64 + 1024;
};

template <int DIL>
__global__ void __launch_bounds__(kRuThreads, 3) k_ru_x(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX,
                                                        const RuDev a) {
  constexpr int C = 64;
  using S = RuxSmem<DIL>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* sA = smem;
  uint8_t* sW = smem + S::kABytes;
  float* sX = reinterpret_cast<float*>(smem + S::kABytes + S::kWBytes);
  float* sStg = reinterpret_cast<float*>(smem);  // valid once the accumulator barrier has fired
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kABytes + S::kWBytes + S::kXBytes);  // [0] weight, [1] accumulator, [2] input tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int item = blockIdx.y;
  const int row0 = (int)blockIdx.x * BM;
  const ItemRef it = get_item(a.items, a.base, item, a.out_len);

  if (tid == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    mbar_init(smem_u32(&bars[2]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // the loads start before anything else of the CTA's setup (the issuing thread initialised the barriers itself)
    mbar_arrive_expect_tx(smem_u32(&bars[2]), (uint32_t)(S::kBoxRows * C * 4));
    tma_load_3d(smem_u32(sX), &tmX, smem_u32(&bars[2]), 0, a.out_lo + row0 - 3 * DIL - a.in_lo, item);
    mbar_arrive_expect_tx(smem_u32(&bars[0]), S::kWBytes);
    tma_load_2d(smem_u32(sW), &tmW, smem_u32(&bars[0]), 0, 0);
  }
  if (warp == 9) tmem_alloc(smem_u32(tmem_slot), C);
  if (a.prefetch_ahead > 0 && tid == 64) {
    const long long lin = (long long)blockIdx.y * gridDim.x + blockIdx.x + a.prefetch_ahead;
    const int pit = (int)(lin / gridDim.x), ptile = (int)(lin - (long long)pit * gridDim.x);
    if (pit < (int)gridDim.y) {
      const int r_lo = max(a.out_lo + ptile * BM - 3 * DIL - a.in_lo, 0);
      const int r_hi = min(a.out_lo + ptile * BM + BM + 3 * DIL - a.in_lo, a.in_rows);
      const char* p = reinterpret_cast<const char*>(a.x + ((size_t)pit * a.in_rows + r_lo) * C);
      if (r_hi > r_lo) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"((uint32_t)(r_hi - r_lo) * C * 4) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp != 9) {
    constexpr int CP = C / 2;
    constexpr int U = (DIL == 1) ? 8 : 9;
    const int idx = tid;  // CP * U <= 288 workers: one unit per thread
    if (idx < CP * U) {
      const int cp = idx % CP, u = idx / CP;
      int first;
      if (DIL == 1) first = u * 16;
      else if (DIL == 3) first = (u % 3) + (u / 3) * 48;
      else first = u;
      const int c = cp * 2;
      DwPairW W;
      W.load(a.w7, a.dw_b, a.a1, a.i1, a.a2, a.i2, C, c);
      uint8_t* a_kb = sA + ((c * 2) & 15);
      const int chunk = (c * 2) >> 4;
      mbar_wait(smem_u32(&bars[2]), 0);
      unit_len_dispatch<DIL>(u, [&](auto len) {
        dw_unit_smem<C, DIL, decltype(len)::value>(sX + first * C + c, W, [&](int j, float2 v) {
          const int trow = first + j * DIL;  // < 128 by construction of the unit length
          *reinterpret_cast<__half2*>(a_kb + trow * 128 + (((chunk ^ trow) & 7) << 4)) = f2h2_sat(v.x, v.y);
        });
      });
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  if (warp == 9) {
    if (lane == 0) {
      mbar_wait(smem_u32(&bars[0]), 0);
      tc_fence_after();
      constexpr uint32_t idesc = umma_idesc_f16(C);
      const uint64_t da = umma_desc_k_sw128(smem_u32(sA)), db = umma_desc_k_sw128(smem_u32(sW));
#pragma unroll
      for (int k = 0; k < BK / 16; ++k) umma_f16(tmem_base, da + 2 * k, db + 2 * k, idesc, k ? 1u : 0u);
      umma_commit(smem_u32(&bars[1]));
    }
  } else if (warp < 8) {
    const int q = warp & 3, half = warp >> 2;
    constexpr int NH = C / 32;
    const int c4 = lane & 3, r8 = lane >> 2;
    float* stg = sStg + warp * (32 * 16);
    float* st_p = stg + lane * 16;
    const int st_x = (lane >> 1) & 3;
    const float* ld_p = stg + r8 * 16 + ((c4 ^ ((r8 >> 1) & 3)) << 2);
    const int orow_b = row0 + q * 32 + r8;
    const int tabs_b = a.out_lo + orow_b + it.shift0 * a.up;
    uint32_t vmask = 0, lmask = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (orow_b + 8 * i < a.out_rows) {
        vmask |= 1u << i;
        const int t = tabs_b + 8 * i;
        if (t >= 0 && t < a.T0 * a.up) lmask |= 1u << i;
      }
    }
    const int colb = half * (C / 2) + c4 * 4;
    const size_t ob = ((size_t)item * a.out_rows + orow_b) * C + colb;
    const float* rs = sX + (q * 32 + r8 + 3 * DIL) * C + colb;  // residual = centre-tap input rows of the tile
    mbar_wait(smem_u32(&bars[1]), 0);
    tc_fence_after();
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      uint32_t r[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * (C / 2) + h * 16), r);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<float4*>(st_p + ((j ^ st_x) << 2)) =
            make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                        __uint_as_float(r[4 * j + 3]));
      __syncwarp();
      float4 v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = *reinterpret_cast<const float4*>(ld_p + i * 128);
      __syncwarp();
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.pw_b + colb + h * 16));
      float4 al = make_float4(0.f, 0.f, 0.f, 0.f), iv = al;
      if (a.sn_alpha) {
        al = __ldg(reinterpret_cast<const float4*>(a.sn_alpha + colb + h * 16));
        iv = __ldg(reinterpret_cast<const float4*>(a.sn_inv + colb + h * 16));
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (!((vmask >> i) & 1u)) continue;
        const float4 res = *reinterpret_cast<const float4*>(rs + i * 8 * C + h * 16);
        float4 x = add4(add4(v[i], b4), res);
        if (!((lmask >> i) & 1u)) x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.out32) *reinterpret_cast<float4*>(a.out32 + ob + i * 8 * C + h * 16) = x;
        if (a.out16) {
          if (a.sn_alpha) {
            const float2 lo = snake2(make_float2(x.x, x.y), make_float2(al.x, al.y), make_float2(iv.x, iv.y));
            const float2 hi = snake2(make_float2(x.z, x.w), make_float2(al.z, al.w), make_float2(iv.z, iv.w));
            x = make_float4(lo.x, lo.y, hi.x, hi.y);
          }
          store_half4(a.out16 + ob + i * 8 * C + h * 16, x);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C);
  }
}

// ============================================================================ depthwise k=7 -> fp16, TMA-staged input
// k_dw_tc for the wide blocks (C = 256 / 512) with the same input staging as k_ru_x: one CTA = 128 output rows x 64
// channels of one item, its (128 + 6*DIL) x 64 fp32 input rows arrive by one 3-D TMA load (zero-filled outside the
// item), every thread runs one 16-output unit out of shared memory.  Replaces 22 cp.async + address / predicate
// arithmetic per unit (the standalone kernel is issue-bound: ncu 0.65 instructions per cycle and scheduler).
template <int DIL>
__global__ void __launch_bounds__((DIL == 1) ? 256 : 288, (DIL == 1) ? 4 : 3)
k_dw_x(const __grid_constant__ CUtensorMap tmX, const Item* items, int base, int out_len, int T0, DwTcArgs a, int n_cb, int pf) {
  constexpr int CB = 64;
  using S = RuxSmem<DIL>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  float* sX = reinterpret_cast<float*>(smem);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + S::kXBytes);
  const int tid = threadIdx.x;
  const int item = blockIdx.y;
  const int tile = (int)blockIdx.x / n_cb, cb = (int)blockIdx.x % n_cb;  // channel block fastest
  const int row0 = tile * BM;
  const int out_rows = a.out_r.n();
  if (tid == 0) {
    mbar_init(smem_u32(bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    mbar_arrive_expect_tx(smem_u32(bar), (uint32_t)(S::kBoxRows * CB * 4));
    tma_load_3d(smem_u32(sX), &tmX, smem_u32(bar), cb * CB, a.out_r.lo + row0 - 3 * DIL - a.in_r.lo, item);
  }
  if (pf > 0 && cb == 0 && tid == 32) {
    // the row tile `pf` CTAs ahead in launch order (all its channel blocks: whole rows, contiguous) -> L2
    const long long lin = (long long)blockIdx.y * gridDim.x + blockIdx.x + pf;
    const int pit = (int)(lin / gridDim.x), ptile = (int)(lin - (long long)pit * gridDim.x) / n_cb;
    if (pit < (int)gridDim.y) {
      const int in_rows = a.in_r.n();
      const int r_lo = max(a.out_r.lo + ptile * BM - 3 * DIL - a.in_r.lo, 0);
      const int r_hi = min(a.out_r.lo + ptile * BM + BM + 3 * DIL - a.in_r.lo, in_rows);
      const char* p = reinterpret_cast<const char*>(a.in + ((size_t)pit * in_rows + r_lo) * a.C);
      if (r_hi > r_lo) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"((uint32_t)(r_hi - r_lo) * a.C * 4) : "memory");
    }
  }
  constexpr int CP = CB / 2;
  const int cp = tid % CP, u = tid / CP;
  int first;
  if (DIL == 1) first = u * 16;
  else if (DIL == 3) first = (u % 3) + (u / 3) * 48;
  else first = u;
  const int cl = cp * 2, c = cb * CB + cl;
  const ItemRef it = get_item(items, base, item, out_len);
  DwPairW W;
  W.load(a.w7, a.bias, a.a1, a.i1, a.a2, a.i2, a.C, c);
  __half* o = a.out + ((size_t)item * out_rows + row0) * a.C + c;
  const int t_abs0 = a.out_r.lo + row0 + it.shift0 * a.up, t_hi = T0 * a.up;
  mbar_wait(smem_u32(bar), 0);
  unit_len_dispatch<DIL>(u, [&](auto len) {
    dw_unit_smem<CB, DIL, decltype(len)::value>(sX + first * CB + cl, W, [&](int j, float2 v) {
      const int trow = first + j * DIL;  // < 128 by construction of the unit length
      if (row0 + trow < out_rows) {
        const int t_abs = t_abs0 + trow;
        if (t_abs < 0 || t_abs >= t_hi) v = make_float2(0.f, 0.f);
        *reinterpret_cast<__half2*>(o + (size_t)trow * a.C) = f2h2_sat(v.x, v.y);
      }
    });
  });
}

// ============================================================================ persistent fused ResidualUnit
// Same math as k_ru_tc, restructured so that nothing waits on anything it does not need:
// one persistent CTA per SM walks tiles (item, 128 rows) round-robin with three kinds of warps running
// concurrently on different tiles -
//   warps 0..8   workers : depthwise + Snake -> fp16 operand tile A[ab] in the swizzled K-major layout
//   warp  9      control : TMA-loads W once; per tile waits for A[ab], issues the tcgen05 MMAs into
//                          accumulator D[tb] (TMEM double-buffered), commits a_free / t_full
//   warps 10..17 epilogue: prefetch the residual rows of their tile while the MMAs run, drain D[tb]
//                          (bias + residual + optional next-block Snake), coalesced stores, release D[tb]
// so the workers start tile i+1 while tile i is in the tensor core / epilogue, the weight, the TMEM
// allocation and the barrier setup are paid once per SM instead of once per tile, and C = 256
// (decoder block 1: W = 128 KB) fits because only one CTA lives on an SM.
// NW worker warps, one control warp, NE (4 or 8) epilogue warps
template <int C> struct RupCfg {
  static constexpr int NW = 11, NE = 8;  // 20 warps: registers are allocated per 4 warps, 96 each
  static constexpr int PD = (C >= 128) ? 4 : 2;  // residual prefetch depth of the epilogue (16-column steps)
  static constexpr int kThreads = (NW + 1 + NE) * 32;
};
template <int C> struct RupSmem {
  static constexpr int kNA = (C <= 128) ? 2 : 1;            // operand tile buffers
  static constexpr int kABytes = BM * C * 2;
  static constexpr int kWBytes = C * C * 2;
  static constexpr int kStgBytes = 8 * 32 * 16 * 4;
  static constexpr int kBytes = kWBytes + kNA * kABytes + kStgBytes + 256 + 1024;
};

template <int C, int DIL>
__global__ void __launch_bounds__(RupCfg<C>::kThreads, 1) k_ru_p(const __grid_constant__ CUtensorMap tmW, const RuDev a,
                                                         const int tiles_per_item, const int total_tiles) {
  using S = RupSmem<C>;
  constexpr int KB = C / BK;
  constexpr int NA = S::kNA;
  constexpr int NW = RupCfg<C>::NW, NE = RupCfg<C>::NE;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* sW = smem;
  uint8_t* sA = smem + S::kWBytes;
  float* sStg = reinterpret_cast<float*>(sA + NA * S::kABytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sStg) + S::kStgBytes);
  // bars: [0] w_full, [1..2] a_full, [3..4] a_free, [5..6] t_full, [7..8] t_empty
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bars[1 + i]), NW * 32);
      mbar_init(smem_u32(&bars[3 + i]), 1);
      mbar_init(smem_u32(&bars[5 + i]), 1);
      mbar_init(smem_u32(&bars[7 + i]), NE * 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == NW) tmem_alloc(smem_u32(tmem_slot), 2 * C);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < NW) {
    // ===================================================================== workers
    constexpr int CP = C / 2;
    constexpr int U = (DIL == 1) ? 8 : 9;
    int i = 0;
    for (int lin = blockIdx.x; lin < total_tiles; lin += gridDim.x, ++i) {
      const int item = lin / tiles_per_item, row0 = (lin - item * tiles_per_item) * BM;
      const int ab = i % NA;
      if (tid == 0) {  // pull the rows of this CTA's next tile into L2 while this one is computed
        const int nl = lin + gridDim.x;
        if (nl < total_tiles) {
          const int nit = nl / tiles_per_item, nrow0 = (nl - nit * tiles_per_item) * BM;
          const int r_lo = max(a.out_lo + nrow0 - 3 * DIL - a.in_lo, 0);
          const int r_hi = min(a.out_lo + nrow0 + BM + 3 * DIL - a.in_lo, a.in_rows);
          if (r_hi > r_lo)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.x + ((size_t)nit * a.in_rows + r_lo) * C),
                         "r"((uint32_t)(r_hi - r_lo) * C * 4)
                         : "memory");
        }
      }
      mbar_wait(smem_u32(&bars[3 + ab]), ((i / NA) & 1) ^ 1);  // the MMAs that read A[ab] last time are done
      const float* xin = a.x + (size_t)item * a.in_rows * C;
      uint8_t* sAb = sA + ab * S::kABytes;
#pragma unroll 1
      for (int idx = tid; idx < CP * U; idx += NW * 32) {
        const int cp = idx % CP, u = idx / CP;
        int first;
        if (DIL == 1) first = u * 16;
        else if (DIL == 3) first = (u % 3) + (u / 3) * 48;
        else first = u;
        const int c = cp * 2;
        DwPairW W;
        W.load(a.w7, a.dw_b, a.a1, a.i1, a.a2, a.i2, C, c);
        const int in_first = a.out_lo + row0 + first - 3 * DIL - a.in_lo;
        uint32_t mlo, mhi;
        row_mask<DIL>(in_first, a.in_rows, 22, mlo, mhi);
        uint8_t* a_kb = sAb + (c / BK) * (BM * 128) + ((c % BK) * 2 & 15);
        const int chunk = ((c % BK) * 2) >> 4;
        dw_unit<C, DIL, 16>(xin + (long long)in_first * C + c, mlo, mhi, W, [&](int j, float2 v) {
          const int trow = first + j * DIL;
          if (trow < BM)
            *reinterpret_cast<__half2*>(a_kb + trow * 128 + (((chunk ^ trow) & 7) << 4)) = f2h2_sat(v.x, v.y);
        });
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[1 + ab])) : "memory");
    }
  } else if (warp == NW) {
    // ===================================================================== control: W load + MMA issue
    if (lane == 0) {
      mbar_arrive_expect_tx(smem_u32(&bars[0]), S::kWBytes);
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(smem_u32(sW + kb * C * 128), &tmW, smem_u32(&bars[0]), kb * BK, 0);
      mbar_wait(smem_u32(&bars[0]), 0);
      constexpr uint32_t idesc = umma_idesc_f16(C);
      int i = 0;
      for (int lin = blockIdx.x; lin < total_tiles; lin += gridDim.x, ++i) {
        const int ab = i % NA, tb = i & 1;
        mbar_wait(smem_u32(&bars[1 + ab]), (i / NA) & 1);          // operand tile written
        mbar_wait(smem_u32(&bars[7 + tb]), ((i >> 1) & 1) ^ 1);    // accumulator drained
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          const uint64_t da = umma_desc_k_sw128(smem_u32(sA + ab * S::kABytes + kb * (BM * 128)));
          const uint64_t db = umma_desc_k_sw128(smem_u32(sW + kb * (C * 128)));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_f16(tmem_base + tb * C, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
        }
        umma_commit(smem_u32(&bars[3 + ab]));
        umma_commit(smem_u32(&bars[5 + tb]));
      }
    }
  } else {
    // ===================================================================== epilogue (warps 10..17)
    const int ew = warp - NW - 1;
    const int q = warp & 3;
    constexpr int NH = C / 32;                    // 16-column steps per warp and column half
    constexpr int PD = RupCfg<C>::PD;            // residual prefetch depth (steps)
    const int c4 = lane & 3, r8 = lane >> 2;
    float* stg = sStg + ew * (32 * 16);
    float* st_p = stg + lane * 16;
    const int st_x = (lane >> 1) & 3;
    const float* ld_p = stg + r8 * 16 + ((c4 ^ ((r8 >> 1) & 3)) << 2);
    int i = 0;
    for (int lin = blockIdx.x; lin < total_tiles; lin += gridDim.x, ++i) {
      const int item = lin / tiles_per_item, row0 = (lin - item * tiles_per_item) * BM;
      const int tb = i & 1;
      const ItemRef it = get_item(a.items, a.base, item, a.out_len);
      const int orow_b = row0 + q * 32 + r8;  // first of this lane's four rows (stride 8)
      const int tabs_b = a.out_lo + orow_b + it.shift0 * a.up;
      uint32_t vmask = 0, lmask = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (orow_b + 8 * k < a.out_rows) {
          vmask |= 1u << k;
          const int t = tabs_b + 8 * k;
          if (t >= 0 && t < a.T0 * a.up) lmask |= 1u << k;
        }
      }
      bool waited = false;
#pragma unroll 1
      for (int half = (NE == 8) ? (ew >> 2) : 0; half < ((NE == 8) ? (ew >> 2) + 1 : 2); ++half) {
      const int colb = half * (C / 2) + c4 * 4;
      const size_t ob = ((size_t)item * a.out_rows + orow_b) * C + colb;
      const float* rp = a.x + ((size_t)item * a.in_rows + (a.out_lo + orow_b - a.in_lo)) * C + colb;
      float4 res[PD][4];
#pragma unroll
      for (int h = 0; h < PD; ++h)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          res[h][k] = make_float4(0.f, 0.f, 0.f, 0.f);
          if ((vmask >> k) & 1u) res[h][k] = __ldg(reinterpret_cast<const float4*>(rp + k * 8 * C + h * 16));
        }
      if (!waited) {
        mbar_wait(smem_u32(&bars[5 + tb]), (i >> 1) & 1);
        tc_fence_after();
        waited = true;
      }
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        uint32_t r[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tb * C + half * (C / 2) + h * 16), r);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<float4*>(st_p + ((j ^ st_x) << 2)) =
              make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                          __uint_as_float(r[4 * j + 3]));
        __syncwarp();
        float4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = *reinterpret_cast<const float4*>(ld_p + k * 128);
        __syncwarp();
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.pw_b + colb + h * 16));
        float4 al = make_float4(0.f, 0.f, 0.f, 0.f), iv = al;
        if (a.sn_alpha) {
          al = __ldg(reinterpret_cast<const float4*>(a.sn_alpha + colb + h * 16));
          iv = __ldg(reinterpret_cast<const float4*>(a.sn_inv + colb + h * 16));
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (!((vmask >> k) & 1u)) continue;
          float4 x = add4(add4(v[k], b4), res[h % PD][k]);
          if (!((lmask >> k) & 1u)) x = make_float4(0.f, 0.f, 0.f, 0.f);
          if (a.out32) *reinterpret_cast<float4*>(a.out32 + ob + k * 8 * C + h * 16) = x;
          if (a.out16) {
            if (a.sn_alpha) {
              const float2 lo = snake2(make_float2(x.x, x.y), make_float2(al.x, al.y), make_float2(iv.x, iv.y));
              const float2 hi = snake2(make_float2(x.z, x.w), make_float2(al.z, al.w), make_float2(iv.z, iv.w));
              x = make_float4(lo.x, lo.y, hi.x, hi.y);
            }
            store_half4(a.out16 + ob + k * 8 * C + h * 16, x);
          }
        }
        if (h + PD < NH) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            res[h % PD][k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if ((vmask >> k) & 1u) res[h % PD][k] = __ldg(reinterpret_cast<const float4*>(rp + k * 8 * C + (h + PD) * 16));
          }
        }
      }
      }
      tc_fence_before();
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[7 + tb])) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NW) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * C);
  }
}

template <int C, int DIL>
cudaError_t launch_rup_t(const CUtensorMap& mw, const RuDev& d, int tiles_per_item, int total_tiles, cudaStream_t st) {
  static bool attr_set_dev[kMaxDev] = {};
  bool& attr_set = attr_set_dev[cur_dev()];  // kernel attributes are per device
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_ru_p<C, DIL>, cudaFuncAttributeMaxDynamicSharedMemorySize, RupSmem<C>::kBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int grid = std::min(total_tiles, sm_count());
  k_ru_p<C, DIL><<<grid, RupCfg<C>::kThreads, RupSmem<C>::kBytes, st>>>(mw, d, tiles_per_item, total_tiles);
  return cudaGetLastError();
}
template <int C>
cudaError_t launch_rup_c(int dil, const CUtensorMap& mw, const RuDev& d, int tpi, int total, cudaStream_t st) {
  if (dil == 1) return launch_rup_t<C, 1>(mw, d, tpi, total, st);
  if (dil == 3) return launch_rup_t<C, 3>(mw, d, tpi, total, st);
  return launch_rup_t<C, 9>(mw, d, tpi, total, st);
}

template <int C, int DIL>
cudaError_t launch_ru_t(const CUtensorMap& mw, const RuDev& d, dim3 grid, cudaStream_t st) {
  static bool attr_set_dev[kMaxDev] = {};
  bool& attr_set = attr_set_dev[cur_dev()];  // kernel attributes are per device
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_ru_tc<C, DIL>, cudaFuncAttributeMaxDynamicSharedMemorySize, RuSmem<C>::kBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  k_ru_tc<C, DIL><<<grid, kRuThreads, RuSmem<C>::kBytes, st>>>(mw, d);
  return cudaGetLastError();
}
cudaError_t launch_ru_tail(const CUtensorMap& mw, const RuDev& d, dim3 grid, cudaStream_t st) {
  constexpr int bytes = RuSmem<64>::kBytes + RuSmem<64>::kTailBytes;
  static bool attr_set_dev[kMaxDev] = {};
  bool& attr_set = attr_set_dev[cur_dev()];  // kernel attributes are per device
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_ru_tc<64, 9, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  k_ru_tc<64, 9, true><<<grid, kRuThreads, bytes, st>>>(mw, d);
  return cudaGetLastError();
}
const bool g_ru_tma = [] { const char* v = getenv("SNACB_RU_TMA"); return !(v && v[0] == '0'); }();
template <int DIL>
cudaError_t launch_ru_x(const CUtensorMap& mw, const RuDev& d, dim3 grid, cudaStream_t st) {
  CUtensorMap mx;
  if (!get_tmap_x3(d.x, 64, d.in_rows, (int)grid.y, RuxSmem<DIL>::kBoxRows, &mx)) return cudaErrorNotSupported;
  static bool attr_set_dev[kMaxDev] = {};
  bool& attr_set = attr_set_dev[cur_dev()];  // kernel attributes are per device
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_ru_x<DIL>, cudaFuncAttributeMaxDynamicSharedMemorySize, RuxSmem<DIL>::kBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  k_ru_x<DIL><<<grid, kRuThreads, RuxSmem<DIL>::kBytes, st>>>(mw, mx, d);
  return cudaGetLastError();
}
template <int C>
cudaError_t launch_ru_c(int dil, const CUtensorMap& mw, const RuDev& d, dim3 grid, cudaStream_t st) {
  if (C == 64 && g_ru_tma && d.in_rows < 65536 && grid.y < 65536) {
    if (dil == 1) return launch_ru_x<1>(mw, d, grid, st);
    if (dil == 3) return launch_ru_x<3>(mw, d, grid, st);
    return launch_ru_x<9>(mw, d, grid, st);
  }
  if (dil == 1) return launch_ru_t<C, 1>(mw, d, grid, st);
  if (dil == 3) return launch_ru_t<C, 3>(mw, d, grid, st);
  return launch_ru_t<C, 9>(mw, d, grid, st);
}

}  // namespace

// measured 0.61 vs 0.59 ms per tick for the standalone depthwise class (HBM-bound either way): opt-in
const bool g_dw_tma = [] { const char* v = getenv("SNACB_DW_TMA"); return v && v[0] == '1'; }();
template <int DIL>
bool launch_dw_x_t(const GroupCtx& g, const DwTcArgs& a, const CUtensorMap& mx) {
  constexpr int bytes = RuxSmem<DIL>::kXBytes + 16 + 128;
  static bool attr_set_dev[kMaxDev] = {};
  bool& attr_set = attr_set_dev[cur_dev()];  // kernel attributes are per device
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_dw_x<DIL>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) { cudaGetLastError(); return false; }
    attr_set = true;
  }
  const int n_cb = a.C / 64, tiles = (a.out_r.n() + BM - 1) / BM;
  dim3 grid((unsigned)(tiles * n_cb), (unsigned)g.n_items);
  static const int pf_on = [] { const char* v = getenv("SNACB_PREFETCH"); return (v && v[0] == '0') ? 0 : 1; }();
  const int pf = pf_on ? sm_count() * ((DIL == 1) ? 4 : 3) : 0;
  k_dw_x<DIL><<<grid, (DIL == 1) ? 256 : 288, bytes, g.stream>>>(mx, g.items, g.base, g.out_len, g.T0, a, n_cb, pf);
  return true;
}
bool launch_dw_x(const GroupCtx& g, const DwTcArgs& a) {
  if (!g_dw_tma || a.C % 64 || a.C < 128 || g.n_items >= 65536 || a.in_r.n() >= 65536 || a.out_r.n() <= 0) return false;
  CUtensorMap mx;
  const int box = BM + 6 * a.dil;
  if (!get_tmap_x3(a.in, a.C, a.in_r.n(), g.n_items, box, &mx, 64)) return false;
  return a.dil == 1 ? launch_dw_x_t<1>(g, a, mx) : a.dil == 3 ? launch_dw_x_t<3>(g, a, mx) : launch_dw_x_t<9>(g, a, mx);
}

bool ru_tc_supported(int C, bool persistent) { return C == 64 || C == 128 || (C == 256 && persistent); }

cudaError_t launch_ru_tc(const GroupCtx& g, const RuTcArgs& a) {
  if (!ru_tc_supported(a.C, a.persistent) || g.n_items <= 0 || a.out_r.n() <= 0) return cudaErrorInvalidValue;
  CUtensorMap mw;
  if (!get_tmap(a.pw16, a.C, a.C, a.C, &mw)) return cudaErrorNotSupported;
  RuDev d{};
  d.items = g.items; d.base = g.base; d.out_len = g.out_len; d.T0 = g.T0;
  d.x = a.x; d.in_lo = a.in_r.lo; d.in_rows = a.in_r.n(); d.out_lo = a.out_r.lo; d.out_rows = a.out_r.n(); d.up = a.up;
  d.w7 = a.w7; d.dw_b = a.dw_b; d.a1 = a.a1; d.i1 = a.i1; d.a2 = a.a2; d.i2 = a.i2; d.pw_b = a.pw_b;
  d.out32 = a.out32; d.out16 = a.out16; d.sn_alpha = a.sn_alpha; d.sn_inv = a.sn_inv;
  d.prefetch_ahead = a.prefetch_ahead;
  const int tpi = (a.out_r.n() + BM - 1) / BM;
  cudaError_t e;
  if (a.tail_w7) {  // last ResidualUnit of block 3 + decoder tail + PCM pack in one kernel
    if (a.C != 64 || a.dil != 9) return cudaErrorInvalidValue;
    d.tile_stride = BM - 6; d.row_off = a.tail_out.lo - 3 - a.out_r.lo; d.tail_lo = a.tail_out.lo; d.tail_n = a.tail_out.n();
    d.tail_w7 = a.tail_w7; d.tail_b = a.tail_b; d.status = a.status; d.wav = a.wav; d.pcm = a.pcm;
    dim3 grid((unsigned)((a.tail_out.n() + d.tile_stride - 1) / d.tile_stride), (unsigned)g.n_items);
    e = launch_ru_tail(mw, d, grid, g.stream);
    ++*g.launches;
    return e;
  }
  if (a.persistent) {
    const int total = tpi * g.n_items;
    e = (a.C == 64) ? launch_rup_c<64>(a.dil, mw, d, tpi, total, g.stream)
        : (a.C == 128) ? launch_rup_c<128>(a.dil, mw, d, tpi, total, g.stream)
                       : launch_rup_c<256>(a.dil, mw, d, tpi, total, g.stream);
  } else {
    dim3 grid((unsigned)tpi, (unsigned)g.n_items);
    e = (a.C == 64) ? launch_ru_c<64>(a.dil, mw, d, grid, g.stream) : launch_ru_c<128>(a.dil, mw, d, grid, g.stream);
  }
  ++*g.launches;
  return e;
}

namespace {
// ============================================================================ from_codes + decoder head depthwise
// z = sum_l outproj_l(codebook_l[code_l]) (strides 4/2/1) for the latent rows of one item, kept in shared memory,
// then decoder.model.0 (depthwise k7) straight to the fp16 operand of the 768->1024 GEMM.  One CTA per item, a
// thread owns 3 channels and keeps their 3 x 8 out-projection weights in registers across all rows (the standalone
// k_from_codes re-reads them for every row and writes z to HBM only for k_dwconv to read it back).
__global__ void __launch_bounds__(256) k_codes_head(const Item* items, int base, int out_len, int T0, QuantW q,
                                                    const int32_t* __restrict__ c0, const int32_t* __restrict__ c1,
                                                    const int32_t* __restrict__ c2, int pitch0, Rng z, Rng h,
                                                    const float* __restrict__ w7, const float* __restrict__ dw_b,
                                                    __half* __restrict__ out) {
  extern __shared__ float zs[];          // [z rows][768] then [z rows][24] embeddings
  const int zr = z.n(), hr = h.n();
  float* es = zs + (size_t)zr * kLatent;
  const int i = blockIdx.x, tid = threadIdx.x;
  const ItemRef it = get_item(items, base, i, out_len);
  for (int e = tid; e < zr * 24; e += 256) {
    const int j = e / 24, r = e - j * 24, l = r >> 3, d = r & 7;
    const int u = z.lo + j + it.shift0;
    float v = 0.0f;
    if (u >= 0 && u < T0) {
      int k = (l == 0) ? c0[(size_t)it.code_row * pitch0 + (u >> 2)]
              : (l == 1) ? c1[(size_t)it.code_row * 2 * pitch0 + (u >> 1)] : c2[(size_t)it.code_row * 4 * pitch0 + u];
      k = min(max(k, 0), SNACB_CODEBOOK_SIZE - 1);  // rejected windows stay memory-safe
      v = q.codebook[l][(size_t)k * 8 + d];
    }
    es[e] = v;
  }
  __syncthreads();
#pragma unroll
  for (int cc = 0; cc < 3; ++cc) {
    const int c = tid + cc * 256;
    float w[24], bsum = 0.0f;
#pragma unroll
    for (int l = 0; l < 3; ++l) {
      const float4 a = *reinterpret_cast<const float4*>(q.w[l] + (size_t)c * 8), b = *reinterpret_cast<const float4*>(q.w[l] + (size_t)c * 8 + 4);
      w[8 * l] = a.x; w[8 * l + 1] = a.y; w[8 * l + 2] = a.z; w[8 * l + 3] = a.w;
      w[8 * l + 4] = b.x; w[8 * l + 5] = b.y; w[8 * l + 6] = b.z; w[8 * l + 7] = b.w;
    }
    for (int j = 0; j < zr; ++j) {
      const int u = z.lo + j + it.shift0;
      float acc = 0.0f;
      if (u >= 0 && u < T0) {
#pragma unroll
        for (int l = 0; l < 3; ++l) {  // same association as k_from_codes: per level dot product + bias, summed over levels
          float d = w[8 * l] * es[j * 24 + 8 * l];
#pragma unroll
          for (int t = 1; t < 8; ++t) d = fmaf(w[8 * l + t], es[j * 24 + 8 * l + t], d);
          acc += d + q.b[l][c];
        }
      }
      zs[(size_t)j * kLatent + c] = acc;
    }
    (void)bsum;
  }
  __syncthreads();
  for (int e = tid; e < hr * kLatent; e += 256) {
    const int j = e / kLatent, c = e - j * kLatent;
    const int t_rel = h.lo + j, t_abs = t_rel + it.shift0;
    float acc = 0.0f;
    if (t_abs >= 0 && t_abs < T0) {
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        const int r = t_rel + k - 3 - z.lo;
        const float v = (r >= 0 && r < zr) ? zs[(size_t)r * kLatent + c] : 0.0f;
        acc = fmaf(w7[k * kLatent + c], v, acc);
      }
      acc += dw_b[c];
    }
    out[((size_t)i * hr + j) * kLatent + c] = __low2half(f2h2_sat(acc, 0.0f));
  }
}

}  // namespace

bool codes_head_supported(Rng z) { return (size_t)z.n() * (kLatent + 24) * sizeof(float) <= 200 * 1024; }

void launch_codes_head(const GroupCtx& g, const QuantW& q, const int32_t* c0, const int32_t* c1, const int32_t* c2, int pitch0,
                       Rng z, Rng h, const float* w7, const float* dw_b, __half* out) {
  const size_t smem = (size_t)z.n() * (kLatent + 24) * sizeof(float);
  static size_t max_set_dev[kMaxDev] = {};
  size_t& max_set = max_set_dev[cur_dev()];
  if (max_set < 48 * 1024) max_set = 48 * 1024;
  if (smem > max_set) {
    cudaFuncSetAttribute(k_codes_head, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    max_set = smem;
  }
  k_codes_head<<<g.n_items, 256, smem, g.stream>>>(g.items, g.base, g.out_len, g.T0, q, c0, c1, c2, pitch0, z, h, w7, dw_b, out);
  ++*g.launches;
}

namespace {
// fp32 -> fp16 weight conversion (load time)
__global__ void k_to_half(const float* __restrict__ in, __half* __restrict__ out, size_t n) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n) out[i] = __low2half(f2h2_sat(in[i], 0.0f));  // weights beyond fp16 range clamp, never inf
}
}  // namespace
void launch_to_half(const float* in, __half* out, size_t n, cudaStream_t st) {
  if (n) k_to_half<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, out, n);
}

namespace {
// x -> (hi, lo) with hi = fp16(x) (saturating), lo = fp16(x - hi): hi + lo carries ~22 bits of x
__device__ __forceinline__ void split2(float a, float b, __half2& hi, __half2& lo) {
  hi = f2h2_sat(a, b);
  const float2 h = __half22float2(hi);
  lo = f2h2_sat(a - h.x, b - h.y);
}
__global__ void __launch_bounds__(256) k_split16(const float* __restrict__ in, __half* __restrict__ out, size_t rows, int C,
                                                 const float* __restrict__ alpha, const float* __restrict__ inv) {
  const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;  // one float4 of one row
  const int c4 = C / 4;
  if (idx >= rows * c4) return;
  const size_t row = idx / c4;
  const int c = (int)(idx - row * c4) * 4;
  float4 v = *reinterpret_cast<const float4*>(in + row * C + c);
  if (alpha) {  // exact Snake (sinf), like the fp32 recipe
    const float4 al = *reinterpret_cast<const float4*>(alpha + c), iv = *reinterpret_cast<const float4*>(inv + c);
    v.x = snake_exact(v.x, al.x, iv.x); v.y = snake_exact(v.y, al.y, iv.y);
    v.z = snake_exact(v.z, al.z, iv.z); v.w = snake_exact(v.w, al.w, iv.w);
  }
  __half2 h0, l0, h1, l1;
  split2(v.x, v.y, h0, l0);
  split2(v.z, v.w, h1, l1);
  __half* o = out + row * 2 * C + c;
  uint2 ph, pl;
  ph.x = *reinterpret_cast<uint32_t*>(&h0); ph.y = *reinterpret_cast<uint32_t*>(&h1);
  pl.x = *reinterpret_cast<uint32_t*>(&l0); pl.y = *reinterpret_cast<uint32_t*>(&l1);
  *reinterpret_cast<uint2*>(o) = ph;
  *reinterpret_cast<uint2*>(o + C) = pl;
}
__global__ void __launch_bounds__(256) k_split_w(const float* __restrict__ in, __half* __restrict__ out, int N, int nseg, int K) {
  const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
  const size_t tot = (size_t)N * nseg * K;
  if (idx >= tot) return;
  const size_t n = idx / ((size_t)nseg * K);
  const int rem = (int)(idx - n * nseg * K), seg = rem / K, k = rem - seg * K;
  const float x = in[idx];
  const __half hi = __low2half(f2h2_sat(x, 0.0f));
  const __half lo = __low2half(f2h2_sat(x - __half2float(hi), 0.0f));
  __half* o = out + (n * nseg + seg) * 3 * (size_t)K;
  o[k] = hi; o[K + k] = lo; o[2 * K + k] = hi;
}
}  // namespace
void launch_split16(const float* in, __half* out, size_t rows, int C, const float* alpha, const float* inv, cudaStream_t st,
                    int64_t* launches) {
  const size_t n = rows * (C / 4);
  if (!n) return;
  k_split16<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, out, rows, C, alpha, inv);
  ++*launches;
}
void launch_split_w(const float* in, __half* out, int N, int nseg, int K, cudaStream_t st) {
  const size_t n = (size_t)N * nseg * K;
  if (n) k_split_w<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, out, N, nseg, K);
}

}  // namespace snacb
