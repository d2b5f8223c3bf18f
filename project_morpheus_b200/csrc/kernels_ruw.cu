// Fused ResidualUnit for the WIDE decoder block 1 (C = 256):  x' = x + W Snake(dw7_dil(Snake(x))) + b  in one
// persistent, warp-specialised kernel (oracle `snac_ref.py` ResidualUnit; third-party `snac`, called at
// Morpheus_Client/tts_engine/speechpipe.py:118).  Replaces the k_dw_tc + k_gemm_ws pair, which moves 16 bytes per
// element through HBM (fp32 x in, fp16 operand out; fp16 operand + fp32 residual in, fp32 out): this kernel reads x
// once and writes x' once (8 B per element, plus the depthwise halo rows, which come from L2).
//
// One CTA per SM walks (item, 128-row tile) pairs.  Per tile the 256 channels are processed as four 64-channel blocks:
//   producer (1 thread)   TMA-loads the fp32 rows of a block (128 + 6*dil rows, halo included, zero-filled outside the
//                         item = the conv's zero padding) as two [rows][32 ch] boxes with the 128-byte swizzle, two
//                         blocks in flight; streams the 1x1 weight as [256 n][64 k] k-block tiles (L2-resident) through
//                         a two-stage ring
//   workers (16 warps)    depthwise units out of shared memory (Snake -> k7 -> Snake), fp16 operand k-block written in
//                         the tcgen05 SWIZZLE_128B layout
//   init (4 warps)        the tile's OWN rows of the block (+ the 1x1 conv's bias) go from shared memory straight into
//                         TENSOR MEMORY (row per lane, conflict-free thanks to the swizzle, tcgen05.st): the accumulator
//                         starts as x + b
//   MMA (1 thread)        once the four blocks are in: D += A W^T, k-block by k-block (N = 256) - the residual add is
//                         the accumulate flag
//   epilogue (4 + 4 warps) D -> fp32 x' (or, for the block's last unit, the next block's Snake -> fp16 operand): row per
//                         lane from tensor memory into swizzled staging tiles, out through TMA tensor stores.  Two teams:
//                         the dedicated warps take the even 128-byte column slabs, the init warps the odd ones.  512 / C
//                         accumulators in tensor memory (2 at C = 256, 4 at C = 128) decouple it from the tiles being built
#include <cstdlib>

#include "snacb.h"
#include "tc_ptx.cuh"

namespace snacb {
namespace {

// C = 256 (decoder block 1) and C = 128 (block 2): C / 64 k-blocks of the 1x1 GEMM, C / 32 input blocks per tile (one TMA
// box of 32 channels each)
// Two worker groups of nine warps; group g builds the blocks b with b % 2 == g; the input ring (block counter % stages)
// keeps the next block of each group in flight while it computes.  One depthwise unit of <= 9 outputs per thread and block:
// 16 channel pairs x up to 18 units.
constexpr int kRwGroupWarps = 9, kRwWorkers = 2 * kRwGroupWarps, kRwInit = 4, kRwEpi = 4;
constexpr int kRwThreads = (2 + kRwInit + kRwWorkers + kRwEpi) * 32;  // warp 0 producer, 1 MMA, 2..5 init, 6..23 workers, 24..27 epilogue

struct RwDev {
  const Item* items; int base, out_len, T0;
  const float* x; int in_lo, in_rows, out_lo, out_rows, up;
  const float* w7; const float* dw_b; const float* a1; const float* i1; const float* a2; const float* i2; const float* pw_b;
  float* out32; __half* out16; const float* sn_alpha; const float* sn_inv;
  int tiles_per_item, total_tiles;
  int dbg;  // SNACB_RUW_DBG bits (profiling experiments): 1 workers skip the units, 2 init skips the copy, 4 epilogue skips the stores, 8 no Snake in the fp16 epilogue, 16 staging written but no TMA store issued, 256 one epilogue team
};

template <int C, int DIL> struct RwSmem {
  static constexpr int kBoxRows = BM + 6 * DIL;
  // [rows][32 ch] fp32, 128-byte rows (SWIZZLE_128B: the pattern is a function of the ADDRESS, so every box starts on
  // a 1024-byte boundary and the 16-byte chunk of (row, c) is (c/4) ^ (row % 8))
  static constexpr int kXBytes = (kBoxRows * 128 + 1023) / 1024 * 1024;
  static constexpr int kABytes = BM * C * 2;                   // whole-K operand tile: C / 64 k-blocks of [128][128 B]
  static constexpr int kWStage = C * BK * 2;                   // [C n][64 k] fp16
  // Input ring: FOUR stages, so stage s always belongs to worker group s % 2.  (An odd depth hands a stage to the
  // groups alternately; a group then skips every other use of its barrier and its parity wait can pass two uses early -
  // found the hard way.)  At C = 256 there is room for only one weight stage (64-channel slice of W, 32 KB).
  static constexpr int kStages = 4;
  static constexpr int kWStages = (C == 256) ? 1 : 2;
  // Epilogue staging tiles: the TMA store of slab i reads its tile while slab i+1 is written into the next one.  (With a
  // single tile every slab waited out the previous store: 8 slabs x ~1 us per tile made the epilogue, not the depthwise
  // workers, the bound of the kernel.)
  static constexpr int kOutBufs = (C == 256) ? 2 : 4;        // split evenly between the two epilogue teams
  static constexpr int kOutBytes = BM * 128;                   // epilogue staging: [128 rows][128 B], SWIZZLE_128B, TMA store
  static constexpr int kOffX = 0;
  static constexpr int kOffA = kStages * kXBytes;
  static constexpr int kOffW = kOffA + kABytes;
  static constexpr int kOffOut = kOffW + kWStages * kWStage;
  static constexpr int kOffSn = kOffOut + kOutBufs * kOutBytes;   // next block's Snake constants [alpha C][1/alpha C] (fp16-emitting launches)
  static constexpr int kOffBar = kOffSn + 2 * C * 4;
  static constexpr int kAcc = 512 / C;                         // accumulators in tensor memory: 2 (C = 256) or 4 (C = 128)
  static constexpr int kBytes = kOffBar + 320;
  static_assert(kBytes <= 227 * 1024, "shared memory");
};

// Depthwise unit over the swizzled block.  Input m of the unit sits at box row r0 + m * DIL; its 16-byte chunk is XORed
// with (row % 8) by the swizzle.  (row % 8) = ((r0 % 8) + (m * DIL) % 8) % 8, so the caller hands in the eight
// addresses base[k] = block + r0 * 128 + swizzled offset for row phase (r0 + k) % 8: every load is base[(m*DIL)%8] plus
// an immediate - no per-input address arithmetic.
template <int DIL, int L, typename Sink>
__device__ __forceinline__ void dw_unit_swz(const uint8_t* const (&base)[8], const DwPairW& W, Sink&& sink) {
  float2 acc[7];
#pragma unroll
  for (int m = 0; m < L + 6; ++m) {
    const float2 x = *reinterpret_cast<const float2*>(base[(m * DIL) & 7] + m * DIL * 128);
    // first Snake in its cosine form without the constant: x - h cos(2 a x) (al1 = 2a, iv1 = -h; the "+ h" is folded into
    // the conv's bias by the caller; zero-padded rows give -h, i.e. Snake = 0) - one packed instruction less per input
    const float2 t = __fmul2_rn(W.al1, x);
    const float2 v = __ffma2_rn(W.iv1, make_float2(__cosf(t.x), __cosf(t.y)), x);
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      const int j = m - k;
      if (j >= 0 && j < L) acc[j % 7] = __ffma2_rn(W.w[k], v, (k == 0) ? W.bias : acc[j % 7]);
    }
    if (m >= 6) sink(m - 6, snake2(acc[(m - 6) % 7], W.al2, W.iv2));
  }
}
// Units of the 128-row tile: unit u -> (first output row, outputs).  DIL 1: 16 runs of 8 rows.  DIL 3: residue classes of
// four 27-row segments (9 outputs each) + the last 20 rows (7, 7, 6).  DIL 9: each residue class (15 or 14 rows) in two
// halves.  Returns false for unused slots.
template <int DIL>
__device__ __forceinline__ bool rw_unit(int u, int& first, int& len) {
  if (DIL == 1) { first = 8 * u; len = 8; return u < 16; }
  if (DIL == 3) {
    if (u < 12) { first = 27 * (u / 3) + (u % 3); len = 9; return true; }
    first = 108 + (u - 12); len = (u == 14) ? 6 : 7;
    return u < 15;
  }
  const int rho = u % 9, part = u / 9;
  const int l0 = (rho < 2) ? 8 : 7;
  first = rho + (part ? 9 * l0 : 0); len = part ? 7 : l0;
  return u < 18;
}

template <int CC, int DIL>
__global__ void __launch_bounds__(kRwThreads, 1) k_ru_w(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                                                        const __grid_constant__ CUtensorMap tmO32, const __grid_constant__ CUtensorMap tmO16,
                                                        const RwDev a) {
  using S = RwSmem<CC, DIL>;
  constexpr int C = CC, kRwKB = C / 64, kRwNB = C / 32;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sX = smem + S::kOffX;   // [4 stages][rows][128 B]
  uint8_t* sA = smem + S::kOffA;   // [4 k-blocks][128 rows][128 B]
  uint8_t* sW = smem + S::kOffW;   // [2 stages][256 rows][128 B]
  uint8_t* sOut = smem + S::kOffOut;
  constexpr int NST = S::kStages, NWS = S::kWStages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kOffBar);
  // x_full[<=4] 0..3, x_free[<=4] 4..7, w_full[2] 8..9, w_free[2] 10..11, a_full[4] 12..15, a_free[4] 16..19, d_init[2] 20..21,
  // t_full[2] 22..23, t_empty[2] 24..25
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 32);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  auto bar = [&](int i) { return smem_u32(&bars[i]); };
  constexpr int X_FULL = 0, X_FREE = 4, W_FULL = 8, W_FREE = 10, A_FULL = 12, A_FREE = 16, D_INIT = 20, T_FULL = 24, T_EMPTY = 28;
  constexpr int NACC = S::kAcc;

  if (tid == 0) {
    if (smem_u32(smem) & 1023u) { printf("snacb: k_ru_w dynamic shared memory is not 1024-byte aligned\n"); __trap(); }
    for (int i = 0; i < NST; ++i) { mbar_init(bar(X_FULL + i), 1); mbar_init(bar(X_FREE + i), kRwGroupWarps + kRwInit); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(W_FULL + i), 1);
      mbar_init(bar(W_FREE + i), 1);
    }
    for (int i = 0; i < NACC; ++i) {
      mbar_init(bar(D_INIT + i), kRwInit);
      mbar_init(bar(T_FULL + i), 1);
      mbar_init(bar(T_EMPTY + i), kRwEpi + kRwInit);  // both epilogue teams
    }
    for (int i = 0; i < kRwKB; ++i) { mbar_init(bar(A_FULL + i), kRwWorkers); mbar_init(bar(A_FREE + i), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), NACC * C);
  float* sSn = reinterpret_cast<float*>(smem + S::kOffSn);
  if (a.out16 != nullptr && a.sn_alpha != nullptr)
    for (int c = tid; c < C; c += kRwThreads) { sSn[c] = a.sn_alpha[c]; sSn[C + c] = a.sn_inv[c]; }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_my = ((int)a.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  // ---- epilogue of one tile, shared by two TEAMS of four warps (one warp per TMEM lane quarter each): the dedicated
  // epilogue warps take the even 128-byte column slabs, the init warps - idle most of the time - the odd ones.  (With one
  // team the epilogue was the slowest stage of the pipeline: skipping it made the fp16-emitting launches 70-100 us
  // faster.)  D (-> Snake of the next block) goes row per lane into a swizzled staging tile and leaves through ONE TMA
  // tensor store per slab (fp32: 32 channels, fp16: 64): coalesced, asynchronous, rows beyond the item clipped by the
  // tensor map.  Each team has its own staging tiles, named barrier and store-issuing thread.
  auto epilogue_tile = [&](int ti, int team, int et, int& slab_ctr) {
    constexpr int kTeamBufs = S::kOutBufs / 2;
    const int hq = warp & 3;  // TMEM lane quarter
    const int trow = hq * 32 + lane;
    const bool half_out = a.out16 != nullptr;               // the block's last unit emits the fp16 operand only
    const int cols_per_store = half_out ? 64 : 32;
    const int tb = ti % NACC;
    const int lin = blockIdx.x + ti * gridDim.x;
    const int item = lin / a.tiles_per_item, row0 = (lin - item * a.tiles_per_item) * BM;
    const ItemRef it = get_item(a.items, a.base, item, a.out_len);
    const int t_abs = a.out_lo + row0 + trow + it.shift0 * a.up;
    const bool live = t_abs >= 0 && t_abs < a.T0 * a.up;
    const bool all_live = __all_sync(0xffffffffu, live);  // the usual case: no masking instructions at all
    mbar_wait(bar(T_FULL + tb), (ti / NACC) & 1, 800 + team);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(hq * 32) << 16) + (uint32_t)(tb * C);
    const int nteams = (a.dbg & 256) ? 1 : 2;
    for (int c0 = team * cols_per_store; c0 < C; c0 += nteams * cols_per_store, ++slab_ctr) {
      if ((a.dbg & 4) || team >= nteams) break;
      uint8_t* stage = sOut + (team * kTeamBufs + slab_ctr % kTeamBufs) * S::kOutBytes;
      uint8_t* rowp = stage + trow * 128;
      // the store that last used this staging tile (kTeamBufs slabs of this team ago) has finished READING it
      if (et == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kTeamBufs - 1) : "memory");
      asm volatile("bar.sync %0, 128;" ::"r"(1 + team) : "memory");
      if (!half_out) {
        // fp32: the accumulator already holds x + b + W s: 32 columns = 8 chunks straight from tensor memory to the
        // staging row (the bias went in with the init, rows outside the sequence are zeroed only where there are any)
        uint32_t r0[16], r1[16];
        tmem_ld16_nowait(taddr + c0, r0);
        tmem_ld16_nowait(taddr + c0 + 16, r1);
        tmem_ld_wait();
        if (!all_live && !live) {
#pragma unroll
          for (int j = 0; j < 16; ++j) { r0[j] = 0u; r1[j] = 0u; }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          *reinterpret_cast<uint4*>(rowp + ((((j) ^ trow) & 7) << 4)) = make_uint4(r0[4 * j], r0[4 * j + 1], r0[4 * j + 2], r0[4 * j + 3]);
          *reinterpret_cast<uint4*>(rowp + ((((4 + j) ^ trow) & 7) << 4)) = make_uint4(r1[4 * j], r1[4 * j + 1], r1[4 * j + 2], r1[4 * j + 3]);
        }
      } else {
        // fp16 (the next block's Snake applied): 64 columns = 8 chunks of 8 halves
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int col = c0 + g * 16;
          uint32_t r[16];
          tmem_ld16(taddr + col, r);
          float4 x[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            x[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                               __uint_as_float(r[4 * j + 3]));
            if (a.sn_alpha && !(a.dbg & 8)) {
              const float4 al = *reinterpret_cast<const float4*>(sSn + col + 4 * j);      // broadcast reads
              const float4 iv = *reinterpret_cast<const float4*>(sSn + C + col + 4 * j);
              const float2 lo = snake2(make_float2(x[j].x, x[j].y), make_float2(al.x, al.y), make_float2(iv.x, iv.y));
              const float2 hi = snake2(make_float2(x[j].z, x[j].w), make_float2(al.z, al.w), make_float2(iv.z, iv.w));
              x[j] = make_float4(lo.x, lo.y, hi.x, hi.y);
            }
            if (!all_live && !live) x[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const __half2 h0 = f2h2_sat(x[2 * j].x, x[2 * j].y), h1 = f2h2_sat(x[2 * j].z, x[2 * j].w);
            const __half2 h2 = f2h2_sat(x[2 * j + 1].x, x[2 * j + 1].y), h3 = f2h2_sat(x[2 * j + 1].z, x[2 * j + 1].w);
            uint4 pk;
            pk.x = *reinterpret_cast<const uint32_t*>(&h0); pk.y = *reinterpret_cast<const uint32_t*>(&h1);
            pk.z = *reinterpret_cast<const uint32_t*>(&h2); pk.w = *reinterpret_cast<const uint32_t*>(&h3);
            *reinterpret_cast<uint4*>(rowp + ((((g * 2 + j) ^ trow) & 7) << 4)) = pk;
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync %0, 128;" ::"r"(1 + team) : "memory");
      if (et == 0 && !(a.dbg & 16)) {
        const CUtensorMap* tm = half_out ? &tmO16 : &tmO32;
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                         reinterpret_cast<uint64_t>(tm)),
                     "r"(smem_u32(stage)), "r"(c0), "r"(row0), "r"(item)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar(T_EMPTY + tb)) : "memory");
  };

  if (warp == 0) {
    // ===================================================================== producer
    // one thread feeds both rings by polling: the input blocks of tile i+1 must not queue behind weight stages that
    // only free up when tile i's MMAs run
    if (lane == 0) {
      const int total_x = n_my * kRwNB, total_w = n_my * kRwKB;
      int nb = 0, wk = 0;
      uint32_t spins = 0;
      while (nb < total_x || wk < total_w) {
        bool progress = false;
        if (nb < total_x && mbar_test(bar(X_FREE + (nb % NST)), ((nb / NST) & 1) ^ 1)) {
          const int ti = nb / kRwNB, b = nb - ti * kRwNB, st = nb % NST;
          const int lin = blockIdx.x + ti * gridDim.x;
          const int item = lin / a.tiles_per_item, row0 = (lin - item * a.tiles_per_item) * BM;
          const int brow = a.out_lo + row0 - 3 * DIL - a.in_lo;  // first box row in the item's input rows
          if (b == 0 && ti + 1 < n_my) {
            // The boxes are 128-byte pieces of 1 KB rows: fetched piecewise from DRAM they waste its bursts (measured:
            // workers wait on x_full for 31 % of the time).  The rows of this CTA's NEXT tile are one contiguous region:
            // one bulk L2 prefetch brings them in with full bursts, the boxes then come out of L2.
            const int nl = lin + gridDim.x;
            const int nit = nl / a.tiles_per_item, nrow0 = (nl - nit * a.tiles_per_item) * BM;
            const int nb0 = a.out_lo + nrow0 - 3 * DIL - a.in_lo;
            const int r_lo = max(nb0, 0), r_hi = min(nb0 + S::kBoxRows, a.in_rows);
            if (r_hi > r_lo)
              asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.x + ((size_t)nit * a.in_rows + r_lo) * C),
                           "r"((uint32_t)(r_hi - r_lo) * C * 4)
                           : "memory");
          }
          mbar_arrive_expect_tx(bar(X_FULL + st), S::kBoxRows * 128);
          tma_load_3d(smem_u32(sX + st * S::kXBytes), &tmX, bar(X_FULL + st), b * 32, brow, item);
          ++nb;
          progress = true;
        }
        if (wk < total_w && mbar_test(bar(W_FREE + (wk % NWS)), ((wk / NWS) & 1) ^ 1)) {
          const int st = wk % NWS, kb = wk % kRwKB;
          mbar_arrive_expect_tx(bar(W_FULL + st), S::kWStage);
          tma_load_2d(smem_u32(sW + st * S::kWStage), &tmW, bar(W_FULL + st), kb * BK, 0);
          ++wk;
          progress = true;
        }
        if (progress) spins = 0;
        else if (++spins > (1u << 27)) { printf("snacb: k_ru_w producer timeout\n"); __trap(); }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(C);
      int wk = 0;
      for (int ti = 0; ti < n_my; ++ti) {
        const int tb = ti % NACC;
        mbar_wait(bar(D_INIT + tb), (ti / NACC) & 1, 100);  // the accumulator holds x for the whole tile
        tc_fence_after();
        for (int kb = 0; kb < kRwKB; ++kb, ++wk) {
          const int st = wk % NWS;
          mbar_wait(bar(A_FULL + kb), ti & 1, 200 + kb);
          mbar_wait(bar(W_FULL + st), (wk / NWS) & 1, 300 + st);
          tc_fence_after();
          const uint64_t da = umma_desc_k_sw128(smem_u32(sA + kb * (BM * 128)));
          const uint64_t db = umma_desc_k_sw128(smem_u32(sW + st * S::kWStage));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_f16(tmem_base + tb * C, da + 2 * k, db + 2 * k, idesc, 1u);
          umma_commit(bar(W_FREE + st));
          umma_commit(bar(A_FREE + kb));  // this k-block of the operand tile may be rebuilt for the next tile
        }
        umma_commit(bar(T_FULL + tb));    // accumulator complete
      }
    }
  } else if (warp < 2 + kRwInit) {
    // ===================================================================== init: x -> tensor memory
    const int hq = warp & 3;  // a warp reaches TMEM lanes 32 * (warp % 4) .. + 31
    const int row = hq * 32 + lane + 3 * DIL;  // box row of this thread's tile row
    const int it_ = tid - 2 * 32;              // thread of the team (0..127)
    int nb = 0, slab_ctr = 0;
    for (int ti = 0; ti < n_my; ++ti) {
      const int tb = ti % NACC;
      mbar_wait(bar(T_EMPTY + tb), ((ti / NACC) & 1) ^ 1, 400);  // both epilogue teams drained this accumulator
      tc_fence_after();
      for (int b = 0; b < kRwNB; ++b, ++nb) {
        const int st = nb % NST;
        mbar_wait(bar(X_FULL + st), (nb / NST) & 1, 500 + b);
        const uint8_t* rp = sX + st * S::kXBytes + row * 128;
#pragma unroll
        for (int g = 0; g < 2 && !(a.dbg & 2); ++g) {
          uint32_t r[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 v = *reinterpret_cast<const float4*>(rp + ((((g * 4 + j) ^ row) & 7) << 4));
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.pw_b + b * 32 + g * 16 + 4 * j));
            r[4 * j] = __float_as_uint(v.x + b4.x); r[4 * j + 1] = __float_as_uint(v.y + b4.y);
            r[4 * j + 2] = __float_as_uint(v.z + b4.z); r[4 * j + 3] = __float_as_uint(v.w + b4.w);
          }
          asm volatile(
              "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
              "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(
                  tmem_base + ((uint32_t)(hq * 32) << 16) + (uint32_t)(tb * C + b * 32 + g * 16)),
              "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
              "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
              : "memory");
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar(X_FREE + st)) : "memory");
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar(D_INIT + tb)) : "memory");
      // second duty: the odd column slabs of the PREVIOUS tile's epilogue (its accumulator completes while this tile's
      // blocks are still being built, and the next init needs that epilogue finished anyway)
      if (ti >= 1) epilogue_tile(ti - 1, 1, it_, slab_ctr);
    }
    if (n_my >= 1) epilogue_tile(n_my - 1, 1, it_, slab_ctr);
    if (it_ == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else if (warp < 2 + kRwInit + kRwWorkers) {
    // ===================================================================== workers: depthwise units
    const int wt = tid - (2 + kRwInit) * 32;                      // 0..575
    const int grp = wt / (kRwGroupWarps * 32), gt = wt - grp * (kRwGroupWarps * 32);  // group, thread of the group
    const int p = gt & 15, u = gt >> 4;  // channel pair of the 32-channel block, unit slot: one unit per thread and block
    int first, len;
    const bool has_unit = rw_unit<DIL>(u, first, len);
    // byte offset of this thread's channel pair inside a box row whose phase is (first + k) % 8, plus the unit's first row
    int xoff[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) xoff[k] = first * 128 + ((((p >> 1) ^ (first + k)) & 7) << 4) + (p & 1) * 8;
    for (int ti = 0; ti < n_my; ++ti) {
      for (int b = grp; b < kRwNB; b += 2) {
        const int nb = ti * kRwNB + b, st = nb % NST, kb = b >> 1;
        DwPairW W;  // requested before the waits: the latency of these loads hides behind them
        W.load(a.w7, a.dw_b, a.a1, a.i1, a.a2, a.i2, C, b * 32 + p * 2);
        {  // cosine-form first Snake: scale, sign and the folded constant h * sum_k w[k]
          float2 sw = make_float2(0.f, 0.f);
#pragma unroll
          for (int k = 0; k < 7; ++k) { sw.x += W.w[k].x; sw.y += W.w[k].y; }
          W.bias = make_float2(fmaf(0.5f * W.iv1.x, sw.x, W.bias.x), fmaf(0.5f * W.iv1.y, sw.y, W.bias.y));
          W.al1 = make_float2(2.0f * W.al1.x, 2.0f * W.al1.y);
          W.iv1 = make_float2(-0.5f * W.iv1.x, -0.5f * W.iv1.y);
        }
        if (ti > 0) mbar_wait(bar(A_FREE + kb), (ti - 1) & 1, 600 + kb);  // the previous tile's MMAs have read this k-block
        mbar_wait(bar(X_FULL + st), (nb / NST) & 1, 700 + b);
        if (has_unit && !(a.dbg & 1)) {
          const uint8_t* xs = sX + st * S::kXBytes;
          const uint8_t* base[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) base[k] = xs + xoff[k];
          // operand byte (row, col): col = (b % 2) * 32 + 2p inside the k-block -> chunk (col * 2) / 16, SWIZZLE_128B
          uint8_t* a_kb = sA + kb * (BM * 128) + ((p * 4) & 15);
          const int chunk = (b & 1) * 4 + (p >> 2);
          auto sink = [&](int j, float2 v) {
            const int trow = first + j * DIL;
            *reinterpret_cast<__half2*>(a_kb + trow * 128 + (((chunk ^ trow) & 7) << 4)) = f2h2_sat(v.x, v.y);
          };
          if (len == 9) dw_unit_swz<DIL, 9>(base, W, sink);
          else if (len == 8) dw_unit_swz<DIL, 8>(base, W, sink);
          else if (len == 7) dw_unit_swz<DIL, 7>(base, W, sink);
          else dw_unit_swz<DIL, 6>(base, W, sink);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar(A_FULL + kb)) : "memory");
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar(X_FREE + st)) : "memory");
        }
      }
    }
  } else {
    // ===================================================================== epilogue, team 0 (even slabs)
    const int et = tid - (2 + kRwInit + kRwWorkers) * 32;  // 0..127
    int slab_ctr = 0;
    for (int ti = 0; ti < n_my; ++ti) epilogue_tile(ti, 0, et, slab_ctr);
    if (et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // every store has landed before the CTA exits
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, NACC * C);
  }
}

// fp32 activation tensor [items][rows][C], box = [1][box_rows][32 channels], 128-byte swizzle, zero fill out of bounds
bool get_tmap_x3_swz(const float* ptr, int C, int rows, int n_items, int box_rows, CUtensorMap* out) {
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  if (!fn) return false;
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)rows, (cuuint64_t)n_items};
  cuuint64_t strides[2] = {(cuuint64_t)C * 4, (cuuint64_t)rows * C * 4};
  cuuint32_t box[3] = {32u, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// output tensor [items][rows][C] (fp32 or fp16), box = 128 bytes of channels x 128 rows, 128-byte swizzle: the TMA store
// clips rows beyond `rows`
bool get_tmap_out(const void* ptr, bool half, int C, int rows, int n_items, CUtensorMap* out) {
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  if (!fn) return false;
  const int es = half ? 2 : 4;
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)rows, (cuuint64_t)n_items};
  cuuint64_t strides[2] = {(cuuint64_t)C * es, (cuuint64_t)rows * C * es};
  cuuint32_t box[3] = {(cuuint32_t)(128 / es), (cuuint32_t)BM, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return fn(out, half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(ptr), dims, strides, box,
            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int C, int DIL>
cudaError_t launch_ruw_t(const CUtensorMap& mw, const RwDev& d, const float* x, int n_items, cudaStream_t st) {
  CUtensorMap mx, mo32, mo16;
  if (!get_tmap_x3_swz(x, C, d.in_rows, n_items, RwSmem<C, DIL>::kBoxRows, &mx)) return cudaErrorNotSupported;
  // exactly one of the two outputs is written by a launch; the other map aliases it (never used)
  const void* o32 = d.out32 ? (const void*)d.out32 : (const void*)d.out16;
  const void* o16 = d.out16 ? (const void*)d.out16 : (const void*)d.out32;
  if (!get_tmap_out(o32, false, C, d.out_rows, n_items, &mo32) || !get_tmap_out(o16, true, C, d.out_rows, n_items, &mo16))
    return cudaErrorNotSupported;
  static bool attr_dev[kMaxDev] = {};
  bool& attr_set = attr_dev[cur_dev()];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_ru_w<C, DIL>, cudaFuncAttributeMaxDynamicSharedMemorySize, RwSmem<C, DIL>::kBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int grid = std::min(d.total_tiles, sm_count());
  k_ru_w<C, DIL><<<grid, kRwThreads, RwSmem<C, DIL>::kBytes, st>>>(mx, mw, mo32, mo16, d);
  return cudaGetLastError();
}

}  // namespace

bool ruw_tc_supported(int C) { return C == 256 || C == 128; }

cudaError_t launch_ruw_tc(const GroupCtx& g, const RuTcArgs& a) {
  if (!ruw_tc_supported(a.C) || g.n_items <= 0 || a.out_r.n() <= 0 || g.n_items >= 65536 || a.in_r.n() >= 65536) return cudaErrorInvalidValue;
  if ((a.out32 != nullptr) == (a.out16 != nullptr)) return cudaErrorInvalidValue;  // one output per launch (fp32 stream or fp16 operand)
  CUtensorMap mw;
  if (!get_tmap(a.pw16, a.C, a.C, a.C, &mw)) return cudaErrorNotSupported;
  RwDev d{};
  d.items = g.items; d.base = g.base; d.out_len = g.out_len; d.T0 = g.T0;
  d.x = a.x; d.in_lo = a.in_r.lo; d.in_rows = a.in_r.n(); d.out_lo = a.out_r.lo; d.out_rows = a.out_r.n(); d.up = a.up;
  d.w7 = a.w7; d.dw_b = a.dw_b; d.a1 = a.a1; d.i1 = a.i1; d.a2 = a.a2; d.i2 = a.i2; d.pw_b = a.pw_b;
  d.out32 = a.out32; d.out16 = a.out16; d.sn_alpha = a.sn_alpha; d.sn_inv = a.sn_inv;
  static const int dbg = [] { const char* v = getenv("SNACB_RUW_DBG"); return v ? atoi(v) : 0; }();
  d.dbg = dbg;
  d.tiles_per_item = (a.out_r.n() + BM - 1) / BM;
  const long long total = (long long)d.tiles_per_item * g.n_items;
  if (total >= (1LL << 31)) return cudaErrorInvalidValue;
  d.total_tiles = (int)total;
  cudaError_t e;
  if (a.C == 256)
    e = (a.dil == 1) ? launch_ruw_t<256, 1>(mw, d, a.x, g.n_items, g.stream)
        : (a.dil == 3) ? launch_ruw_t<256, 3>(mw, d, a.x, g.n_items, g.stream)
                       : launch_ruw_t<256, 9>(mw, d, a.x, g.n_items, g.stream);
  else
    e = (a.dil == 1) ? launch_ruw_t<128, 1>(mw, d, a.x, g.n_items, g.stream)
        : (a.dil == 3) ? launch_ruw_t<128, 3>(mw, d, a.x, g.n_items, g.stream)
                       : launch_ruw_t<128, 9>(mw, d, a.x, g.n_items, g.stream);
  ++*g.launches;
  return e;
}

}  // namespace snacb
