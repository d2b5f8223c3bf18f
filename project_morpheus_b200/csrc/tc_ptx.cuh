// PTX wrappers (mbarrier, TMA, tcgen05 / TMEM), UMMA descriptors and the packed-fp32x2 helpers shared by the
// tensor-core translation units (kernels_tc.cu, kernels_blk.cu).  Everything is in an anonymous namespace:
// each translation unit gets its own inlined copy.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdio>

#include "kernels.h"

namespace snacb {
namespace {

constexpr int BM = 128;  // rows per tile = TMEM lanes
constexpr int BK = 64;   // fp16 per k-block = one 128-byte swizzle span
constexpr uint32_t kLiveFlag = 0x40000000u;
constexpr int kMaxDev = 64;
// ordinal of the current device: per-kernel attributes (dynamic shared memory size) and the SM count are per device
inline int cur_dev() {
  int d = 0;
  cudaGetDevice(&d);
  return d & (kMaxDev - 1);
}

template <int N> struct IntC { static constexpr int value = N; };  // compile-time int handed to generic lambdas

// ------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Non-blocking probe (try_wait may suspend the thread for a while before it reports failure).
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug becomes a trap (CUDA error) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int tag = 0) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > ((tag >= 100 && tag < 400) ? (1u << 24) : (1u << 22))) {  // dependants time out after what they wait for
      printf("snacb: mbarrier timeout, block (%d,%d) thread %d tag %d parity %u\n", (int)blockIdx.x, (int)blockIdx.y, (int)threadIdx.x, tag,
             parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// Multicast variants (thread-block clusters): the tile lands at the same shared-memory offset of every CTA in
// `mask` and completes the transaction on each destination CTA's own mbarrier at that offset.
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once every previously issued tcgen05.mma of this thread has completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// Epilogue transpose: a warp holds a 32-row x 16-column fp32 block one row per lane (tcgen05.ld
// 32x32b.x16); after the trip through its private 2 KB of swizzled shared memory lane l holds, for
// i = 0..3, the float4 of row (l/4 + 8i), columns 4*(l%4)..+3 - so global accesses are 64-byte row
// segments.  16-byte chunk index is XORed with (row/2)%4: conflict-free on both sides.
__device__ __forceinline__ void epi_transpose16(float* stg, int lane, const uint32_t (&r)[16], float4 (&v)[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
    *reinterpret_cast<float4*>(stg + lane * 16 + ((j ^ ((lane >> 1) & 3)) << 2)) =
        make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                    __uint_as_float(r[4 * j + 3]));
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = (lane >> 2) + 8 * i;
    v[i] = *reinterpret_cast<const float4*>(stg + row * 16 + (((lane & 3) ^ ((row >> 1) & 3)) << 2));
  }
  __syncwarp();
}
// fp32 pair -> fp16 pair, round to nearest, SATURATING to +-65504 (F2FP.SATFINITE, one instruction like the plain
// conversion): an activation that outgrows fp16 on a trained checkpoint clamps instead of turning into inf -> NaN.
__device__ __forceinline__ __half2 f2h2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return *reinterpret_cast<__half2*>(&r);
}
__device__ __forceinline__ void store_half4(__half* p, float4 v) {
  const __half2 h0 = f2h2_sat(v.x, v.y), h1 = f2h2_sat(v.z, v.w);
  uint2 pk;
  pk.x = *reinterpret_cast<const uint32_t*>(&h0);
  pk.y = *reinterpret_cast<const uint32_t*>(&h1);
  *reinterpret_cast<uint2*>(p) = pk;
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) {
  const float2 lo = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
  const float2 hi = __fadd2_rn(make_float2(a.z, a.w), make_float2(b.z, b.w));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}
// 1024-byte aligned view of the dynamic shared memory that keeps the shared address space (no generic ld/st)
__device__ __forceinline__ uint8_t* smem_align1024(uint8_t* raw) { return raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u); }

// Shared-memory matrix descriptor of a K-major operand tile written by TMA with SWIZZLE_128B:
// rows of 128 bytes (64 fp16 of K), 8-row groups of 1024 bytes (SBO), tile base 1024-byte aligned.
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);  // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                    // leading byte offset (unused for swizzled K-major), bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset = 8 rows * 128 B, bits [32,46)
  d |= (uint64_t)1 << 46;                    // descriptor version (Blackwell), bits [46,48)
  d |= (uint64_t)2 << 61;                    // layout type SWIZZLE_128B, bits [61,64)
  return d;
}
// kind::f16 instruction descriptor: fp16 A/B (K-major), fp32 D, M = 128, N = BN.
__host__ __device__ constexpr uint32_t umma_idesc_f16(int n) {
  return (1u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// Packed fp32x2 arithmetic (sm_100 FFMA2/FMUL2/FADD2): the depthwise + Snake work is issue-bound, and
// every value here comes as a channel pair.
__device__ __forceinline__ float2 snake2(float2 x, float2 al, float2 iv) {
  const float2 t = __fmul2_rn(al, x);
  float2 s = make_float2(__sinf(t.x), __sinf(t.y));
  s = __fmul2_rn(s, s);
  return __ffma2_rn(iv, s, x);
}

// Per-channel-pair constants of one ResidualUnit's depthwise stage (Snake -> depthwise k7 -> Snake)
struct DwPairW {
  float2 al1, iv1, al2, iv2, bias, w[7];
  __device__ __forceinline__ void load(const float* w7, const float* dw_b, const float* a1, const float* i1, const float* a2,
                                       const float* i2, int C, int c) {
    al1 = *reinterpret_cast<const float2*>(a1 + c); iv1 = *reinterpret_cast<const float2*>(i1 + c);
    al2 = *reinterpret_cast<const float2*>(a2 + c); iv2 = *reinterpret_cast<const float2*>(i2 + c);
    bias = *reinterpret_cast<const float2*>(dw_b + c);
#pragma unroll
    for (int k = 0; k < 7; ++k) w[k] = *reinterpret_cast<const float2*>(w7 + k * C + c);
  }
};

}  // namespace

// [rows][cols] fp16 row-major, box = 64 columns x box_rows rows, 128-byte swizzle (cached per process).
bool get_tmap(const void* ptr, long long rows, int cols, int box_rows, CUtensorMap* out);
int sm_count();

}  // namespace snacb
