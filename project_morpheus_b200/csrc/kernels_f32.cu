// fp32 CUDA-core kernels of the SNAC-24k decode path: the exact (SNACB_PREC_FP32) recipe, the
// bring-up reference for the tensor-core kernels, and the integer / gather / pack kernels that are
// shared by both recipes.  All activations are channels-last fp32 (see snacb_common.cuh).
//
// Reference call sites replaced (Morpheus_Client/tts_engine/speechpipe.py):
//   k_deinterleave  <- :72-111  (frame truncation, de-interleave loop, validator)   [bit-exact]
//   k_from_codes    <- :118     quantizer.from_codes (embedding + out_proj + repeat_interleave + sum)
//   k_dwconv/k_gemm_f32/k_snake <- :118  decoder (depthwise k7, 1x1, ConvTranspose1d, NoiseBlock, ResidualUnit)
//   k_tail          <- :118 tail (Snake, conv k7 64->1, tanh) + :122-129 (slice, *32767, int16 trunc)
#include "kernels.h"
#include "snacb.h"

namespace snacb {

// ============================================================================ NS-1 integer kernel
// One warp per window.  Slot s of a frame goes to: 0->L0[f]; 1->L1[2f]; 2->L2[4f]; 3->L2[4f+1];
// 4->L1[2f+1]; 5->L2[4f+2]; 6->L2[4f+3]   (speechpipe.py:84-98).
__global__ void k_deinterleave(const int32_t* __restrict__ tok, int stride, const int32_t* __restrict__ ntok,
                               int ntok_u, int n_win, int maxF, int raw, int32_t* __restrict__ c0,
                               int32_t* __restrict__ c1, int32_t* __restrict__ c2, int32_t* __restrict__ status) {
  const int w = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= n_win) return;
  const int n = ntok ? ntok[w] : ntok_u;
  int F = n / SNACB_TOKENS_PER_FRAME;  // speechpipe.py:72 num_frames = len // 7
  if (F > maxF) F = maxF;
  const int32_t* t = tok + (size_t)w * stride;
  bool bad = false, top = false;
  for (int f = lane; f < maxF; f += 32) {
    int32_t v[7];
#pragma unroll
    for (int s = 0; s < 7; ++s) {
      int32_t x = 0;
      if (f < F) {
        x = t[7 * f + s];
        if (raw) x = x - 10 - 4096 * s;  // speechpipe.py:181 with index%7 == s for aligned windows
        bad |= (x < 0) | (x > 4096);      // speechpipe.py:108-110 ('>' not '>=': 4096 passes, Q1)
        top |= (x == 4096);
      }
      v[s] = x;
    }
    c0[(size_t)w * maxF + f] = v[0];
    int32_t* p1 = c1 + (size_t)w * 2 * maxF + 2 * f;
    p1[0] = v[1]; p1[1] = v[4];
    int32_t* p2 = c2 + (size_t)w * 4 * maxF + 4 * f;
    p2[0] = v[2]; p2[1] = v[3]; p2[2] = v[5]; p2[3] = v[6];
  }
  bad = __any_sync(0xffffffffu, bad);
  top = __any_sync(0xffffffffu, top);
  if (lane == 0) {
    int st = SNACB_WIN_OK;
    if (n < SNACB_TOKENS_PER_FRAME || bad) st = SNACB_WIN_REJECTED;  // :69-70, :108-111
    else if (top) st = SNACB_WIN_CODE4096;                           // embedding would raise
    else if (F == 1) st = SNACB_WIN_EMPTY;                           // :122 empty slice
    status[w] = st;
  }
}

void launch_deinterleave(const int32_t* d_tokens, int tokens_stride, const int32_t* d_ntok, int ntok_uniform,
                         int n_win, int max_frames, bool raw, int32_t* c0, int32_t* c1, int32_t* c2,
                         int32_t* status, cudaStream_t st, int64_t* launches) {
  if (n_win <= 0) return;
  const int threads = 256, wpb = threads / 32;
  k_deinterleave<<<(n_win + wpb - 1) / wpb, threads, 0, st>>>(d_tokens, tokens_stride, d_ntok, ntok_uniform,
                                                               n_win, max_frames, raw ? 1 : 0, c0, c1, c2, status);
  ++*launches;
}

// ============================================================================ NS-2 from_codes
// z[t][c] = sum_l ( b_l[c] + sum_d W_l[c][d] * codebook_l[code_l[t / stride_l]][d] ), strides 4/2/1.
__global__ void __launch_bounds__(256) k_from_codes(const Item* items, int base, int out_len, QuantW q,
                                                    const int32_t* __restrict__ c0, const int32_t* __restrict__ c1,
                                                    const int32_t* __restrict__ c2, int pitch0, Rng z, int T0,
                                                    float* __restrict__ out) {
  const int j = blockIdx.x, i = blockIdx.y;
  const ItemRef it = get_item(items, base, i, out_len);
  const int u = z.lo + j + it.shift0;  // absolute latent step
  float* o = out + ((size_t)i * z.n() + j) * kLatent;
  if (u < 0 || u >= T0) {
    for (int c = threadIdx.x; c < kLatent; c += 256) o[c] = 0.0f;
    return;
  }
  int k[3];
  k[0] = c0[(size_t)it.code_row * pitch0 + (u >> 2)];
  k[1] = c1[(size_t)it.code_row * 2 * pitch0 + (u >> 1)];
  k[2] = c2[(size_t)it.code_row * 4 * pitch0 + u];
  float e[3][8];
#pragma unroll
  for (int l = 0; l < 3; ++l) {
    int kk = min(max(k[l], 0), SNACB_CODEBOOK_SIZE - 1);  // rejected windows stay memory-safe
    const float4* cb = reinterpret_cast<const float4*>(q.codebook[l] + (size_t)kk * 8);
    float4 a = cb[0], b = cb[1];
    e[l][0] = a.x; e[l][1] = a.y; e[l][2] = a.z; e[l][3] = a.w;
    e[l][4] = b.x; e[l][5] = b.y; e[l][6] = b.z; e[l][7] = b.w;
  }
  for (int c = threadIdx.x; c < kLatent; c += 256) {
    float zsum = 0.0f;
#pragma unroll
    for (int l = 0; l < 3; ++l) {
      const float4* w = reinterpret_cast<const float4*>(q.w[l] + (size_t)c * 8);
      float4 a = w[0], b = w[1];
      float d = a.x * e[l][0];
      d = fmaf(a.y, e[l][1], d); d = fmaf(a.z, e[l][2], d); d = fmaf(a.w, e[l][3], d);
      d = fmaf(b.x, e[l][4], d); d = fmaf(b.y, e[l][5], d); d = fmaf(b.z, e[l][6], d); d = fmaf(b.w, e[l][7], d);
      zsum += d + q.b[l][c];
    }
    o[c] = zsum;
  }
}

void launch_from_codes(const GroupCtx& g, const QuantW& q, const int32_t* c0, const int32_t* c1, const int32_t* c2,
                       int pitch0, Rng z, float* out) {
  dim3 grid(z.n(), g.n_items);
  k_from_codes<<<grid, 256, 0, g.stream>>>(g.items, g.base, g.out_len, q, c0, c1, c2, pitch0, z, g.T0, out);
  ++*g.launches;
}

// ============================================================================ depthwise k=7
// out[t][c] = post( b[c] + sum_k w[k][c] * pre(in[t + (k-3)*dil][c]) ); pre/post = Snake or identity.
template <typename T> __device__ __forceinline__ void store_act(T* p, float v);
template <> __device__ __forceinline__ void store_act<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void store_act<__half>(__half* p, float v) { *p = __float2half_rn(fminf(fmaxf(v, -65504.0f), 65504.0f)); }

template <bool PRE, bool POST, typename OutT = float>
__global__ void __launch_bounds__(256) k_dwconv(const Item* items, int base, int out_len, int T0, DwArgs a, OutT* outp) {
  const int i = blockIdx.y;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  const int rows = a.out_r.n();
  if (idx >= (long long)rows * a.C) return;
  const int j = (int)(idx / a.C), c = (int)(idx - (long long)j * a.C);
  const ItemRef it = get_item(items, base, i, out_len);
  const int t_rel = a.out_r.lo + j;
  const int t_abs = t_rel + it.shift0 * a.up;
  OutT* o = outp + ((size_t)i * rows + j) * a.C + c;
  if (t_abs < 0 || t_abs >= T0 * a.up) { store_act<OutT>(o, 0.0f); return; }
  const int in_rows = a.in_r.n();
  const float* x = a.in + (size_t)i * in_rows * a.C + c;
  float al = 0.f, iv = 0.f;
  if (PRE) { al = a.a1[c]; iv = a.i1[c]; }
  float acc = 0.0f;
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    const int r = t_rel + (k - 3) * a.dil - a.in_r.lo;
    float v = (r >= 0 && r < in_rows) ? x[(size_t)r * a.C] : 0.0f;
    if (PRE) v = snake_exact(v, al, iv);
    acc = fmaf(a.w7[k * a.C + c], v, acc);
  }
  acc += a.bias[c];
  if (POST) acc = snake_exact(acc, a.a2[c], a.i2[c]);
  store_act<OutT>(o, acc);
}

void launch_dwconv(const GroupCtx& g, const DwArgs& a) {
  const long long n = (long long)a.out_r.n() * a.C;
  dim3 grid((unsigned)((n + 255) / 256), g.n_items);
  const bool pre = a.a1 != nullptr, post = a.a2 != nullptr;
  if (pre && post) k_dwconv<true, true><<<grid, 256, 0, g.stream>>>(g.items, g.base, g.out_len, g.T0, a, a.out);
  else if (!pre && !post) k_dwconv<false, false><<<grid, 256, 0, g.stream>>>(g.items, g.base, g.out_len, g.T0, a, a.out);
  else if (pre) k_dwconv<true, false><<<grid, 256, 0, g.stream>>>(g.items, g.base, g.out_len, g.T0, a, a.out);
  else k_dwconv<false, true><<<grid, 256, 0, g.stream>>>(g.items, g.base, g.out_len, g.T0, a, a.out);
  ++*g.launches;
}

void launch_dwconv_half(const GroupCtx& g, const DwArgs& a, __half* out16) {
  const long long n = (long long)a.out_r.n() * a.C;
  dim3 grid((unsigned)((n + 255) / 256), g.n_items);
  k_dwconv<false, false, __half><<<grid, 256, 0, g.stream>>>(g.items, g.base, g.out_len, g.T0, a, out16);
  ++*g.launches;
}

// ============================================================================ Snake (block head)
__global__ void __launch_bounds__(256) k_snake(const float* __restrict__ in, float* __restrict__ out, long long n,
                                               int C, const float* __restrict__ alpha, const float* __restrict__ inv) {
  const long long idx = ((long long)blockIdx.x * 256 + threadIdx.x) * 4;
  if (idx >= n) return;
  const int c = (int)(idx % C);
  float4 x = *reinterpret_cast<const float4*>(in + idx);
  float4 al = *reinterpret_cast<const float4*>(alpha + c);
  float4 iv = *reinterpret_cast<const float4*>(inv + c);
  float4 y;
  y.x = snake_exact(x.x, al.x, iv.x); y.y = snake_exact(x.y, al.y, iv.y);
  y.z = snake_exact(x.z, al.z, iv.z); y.w = snake_exact(x.w, al.w, iv.w);
  *reinterpret_cast<float4*>(out + idx) = y;
}

void launch_snake(const GroupCtx& g, const float* in, float* out, Rng r, int C, const float* alpha, const float* inv) {
  const long long n = (long long)g.n_items * r.n() * C;  // Snake(0) == 0, so padded rows stay zero
  if (n <= 0) return;
  k_snake<<<(unsigned)((n / 4 + 255) / 256), 256, 0, g.stream>>>(in, out, n, C, alpha, inv);
  ++*g.launches;
}

// ============================================================================ fp32 GEMM (1x1 / ConvT)
// out[m][n] = epilogue( sum_k A[m][k] * W[n][k] ).  64x64x16 tiles, 256 threads, 4x4 per thread.
// EPI_CONVT runs the transposed conv as a polyphase GEMM: row m is an input position q, column
// n = r*Cout + co is (phase r, channel co), K = 2*Cin: segment 0 reads Snake(x)[q] against tap r+p,
// segment 1 reads Snake(x)[q-1] (tap r+p+s) for r < s-p, else Snake(x)[q+1] (tap r+p-s); the result
// lands at output time q*s + r.  (k = 2s, stride s, pad s/2 => exactly two taps per output sample.)
template <int EPI>
__global__ void __launch_bounds__(256) k_gemm_f32(const Item* items, int base, int out_len, int T0, int n_items,
                                                  GemmArgs a) {
  __shared__ __align__(16) float As[16][68];
  __shared__ __align__(16) float Bs[16][68];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  const int m_rows = a.m_r.n();
  const long long Mtot = (long long)n_items * m_rows;
  const int a_rows = a.a_r.n();

  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  int delta = 0, phase = 0;
  if (EPI == EPI_CONVT) { phase = n0 / a.Cout; delta = (phase < a.s - a.p) ? -1 : 1; }
  const float* arow[2] = {nullptr, nullptr};
  {
    const long long gm = (long long)m0 + lrow;
    if (gm < Mtot) {
      const int item = (int)(gm / m_rows), j = (int)(gm - (long long)item * m_rows);
      const int nseg = (EPI == EPI_CONVT) ? 2 : 1;
      for (int sgm = 0; sgm < nseg; ++sgm) {
        const int row = a.m_r.lo + j + (sgm ? delta : 0) - a.a_r.lo;
        if (row >= 0 && row < a_rows) arow[sgm] = a.A + ((size_t)item * a_rows + row) * a.lda;
      }
    }
  }
  const float* wrow = (n0 + lrow < a.N) ? a.W + (size_t)(n0 + lrow) * a.ldw : nullptr;  // N need not be a multiple of 64

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  const int nseg = (EPI == EPI_CONVT) ? 2 : 1;
  for (int sgm = 0; sgm < nseg; ++sgm) {
    const float* ar = arow[sgm];
    for (int k0 = 0; k0 < a.K; k0 += 16) {
      float4 av = ar ? *reinterpret_cast<const float4*>(ar + k0 + lk) : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 wv = wrow ? *reinterpret_cast<const float4*>(wrow + (size_t)sgm * a.K + k0 + lk) : make_float4(0.f, 0.f, 0.f, 0.f);
      As[lk + 0][lrow] = av.x; As[lk + 1][lrow] = av.y; As[lk + 2][lrow] = av.z; As[lk + 3][lrow] = av.w;
      Bs[lk + 0][lrow] = wv.x; Bs[lk + 1][lrow] = wv.y; Bs[lk + 2][lrow] = wv.z; Bs[lk + 3][lrow] = wv.w;
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) {
        const float4 av4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        const float4 bv4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        const float am[4] = {av4.x, av4.y, av4.z, av4.w};
        const float bn[4] = {bv4.x, bv4.y, bv4.z, bv4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(am[i], bn[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  // ---- epilogue
  const int o_rows = a.o_r.n();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long gm = (long long)m0 + ty * 4 + i;
    if (gm >= Mtot) continue;
    const int item = (int)(gm / m_rows), j = (int)(gm - (long long)item * m_rows);
    const ItemRef it = get_item(items, base, item, out_len);
    int t_rel = a.m_r.lo + j;
    if (EPI == EPI_CONVT) t_rel = t_rel * a.s + phase;
    const int orow = t_rel - a.o_r.lo;
    if (orow < 0 || orow >= o_rows) continue;
    const int t_abs = t_rel + it.shift0 * a.up;
    const bool live = (t_abs >= 0) && (t_abs < T0 * a.up);
    const int ncol0 = n0 + tx * 4 - ((EPI == EPI_CONVT) ? phase * a.Cout : 0);
    if (n0 + tx * 4 >= a.N) continue;
    float v[4];
#pragma unroll
    for (int jn = 0; jn < 4; ++jn) v[jn] = acc[i][jn];
    if (EPI == EPI_BIAS || EPI == EPI_CONVT || EPI == EPI_RESID) {
      const float4 b4 = *reinterpret_cast<const float4*>(a.bias + ncol0);
      v[0] += b4.x; v[1] += b4.y; v[2] += b4.z; v[3] += b4.w;
    }
    if (EPI == EPI_RESID || EPI == EPI_NOISE) {
      const int rrow = t_rel - a.r_r.lo;
      const float4 r4 = *reinterpret_cast<const float4*>(a.R + ((size_t)item * a.r_r.n() + rrow) * a.ldr + ncol0);
      if (EPI == EPI_NOISE) {
        const float nz = live ? noise_at(a.noise, it.code_row, t_abs) : 0.0f;
        v[0] = r4.x + nz * v[0]; v[1] = r4.y + nz * v[1]; v[2] = r4.z + nz * v[2]; v[3] = r4.w + nz * v[3];
      } else {
        v[0] += r4.x; v[1] += r4.y; v[2] += r4.z; v[3] += r4.w;
      }
    }
    if (!live) { v[0] = v[1] = v[2] = v[3] = 0.0f; }
    *reinterpret_cast<float4*>(a.out + ((size_t)item * o_rows + orow) * a.ldo + ncol0) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

void launch_gemm_f32(const GroupCtx& g, const GemmArgs& a) {
  const long long Mtot = (long long)g.n_items * a.m_r.n();
  if (Mtot <= 0) return;
  dim3 grid((unsigned)((Mtot + 63) / 64), (a.N + 63) / 64);
  switch (a.epi) {
    case EPI_BIAS: k_gemm_f32<EPI_BIAS><<<grid, 256, 0, g.stream>>>(g.items, g.base, g.out_len, g.T0, g.n_items, a); break;
    case EPI_RESID: k_gemm_f32<EPI_RESID><<<grid, 256, 0, g.stream>>>(g.items, g.base, g.out_len, g.T0, g.n_items, a); break;
    case EPI_NOISE: k_gemm_f32<EPI_NOISE><<<grid, 256, 0, g.stream>>>(g.items, g.base, g.out_len, g.T0, g.n_items, a); break;
    default: k_gemm_f32<EPI_CONVT><<<grid, 256, 0, g.stream>>>(g.items, g.base, g.out_len, g.T0, g.n_items, a); break;
  }
  ++*g.launches;
}

// ============================================================================ tail + NS-4 pack
// y[t] = tanh( b + sum_{c,k} w[k][c] * Snake(x[t+k-3][c]) );  pcm = (int16) trunc(y * 32767)
// (speechpipe.py:127: no rounding, no clip).  64 samples per CTA, 4 channel quarters per sample.
template <bool FAST>
__global__ void __launch_bounds__(256) k_tail(const Item* items, int base, int out_len, int T0, TailArgs a) {
  __shared__ __align__(16) float xs[70][68];  // row pitch 68 floats: 16-byte aligned rows, conflict-free float4 reads
  __shared__ __align__(16) float ws[7][64];
  __shared__ float part[4][64];
  const int i = blockIdx.y, tid = threadIdx.x;
  const ItemRef it = get_item(items, base, i, out_len);
  const int t0 = a.out_r.lo + blockIdx.x * 64;  // first output sample (relative) of this CTA
  const int x_rows = a.x_r.n();
  const float* x = a.x + (size_t)i * x_rows * 64;
  {
    const int c = (tid & 15) * 4;  // 16 float4 per row: a thread keeps its four channels for every row it stages
    const float4 al = *reinterpret_cast<const float4*>(a.alpha + c), iv = *reinterpret_cast<const float4*>(a.inv + c);
    const int row_b = t0 - 3 - a.x_r.lo;
#pragma unroll
    for (int r = tid >> 4; r < 70; r += 16) {
      const int row = row_b + r;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row >= 0 && row < x_rows) v = *reinterpret_cast<const float4*>(x + (size_t)row * 64 + c);
      float4 o;
      if (FAST) {  // packed fp32x2 Snake (MUFU sin), same arithmetic as the tensor-core recipe's other Snakes
        float2 t = __fmul2_rn(make_float2(al.x, al.y), make_float2(v.x, v.y));
        float2 sn = make_float2(__sinf(t.x), __sinf(t.y));
        const float2 lo = __ffma2_rn(make_float2(iv.x, iv.y), __fmul2_rn(sn, sn), make_float2(v.x, v.y));
        t = __fmul2_rn(make_float2(al.z, al.w), make_float2(v.z, v.w));
        sn = make_float2(__sinf(t.x), __sinf(t.y));
        const float2 hi = __ffma2_rn(make_float2(iv.z, iv.w), __fmul2_rn(sn, sn), make_float2(v.z, v.w));
        o = make_float4(lo.x, lo.y, hi.x, hi.y);
      } else {
        o.x = snake_exact(v.x, al.x, iv.x); o.y = snake_exact(v.y, al.y, iv.y);
        o.z = snake_exact(v.z, al.z, iv.z); o.w = snake_exact(v.w, al.w, iv.w);
      }
      *reinterpret_cast<float4*>(&xs[r][c]) = o;
    }
  }
  for (int e = tid; e < 7 * 64; e += 256) ws[e >> 6][e & 63] = a.w7[e];
  __syncthreads();
  const int sx = tid & 63, qc = tid >> 6;  // output sample, channel quarter
  float acc = 0.0f;
  if (FAST) {
    float2 a2 = make_float2(0.f, 0.f), b2 = a2;  // two independent packed chains
#pragma unroll
    for (int k = 0; k < 7; ++k)
#pragma unroll
      for (int c = 0; c < 16; c += 4) {
        const float4 w4 = *reinterpret_cast<const float4*>(&ws[k][qc * 16 + c]);   // warp-uniform: broadcast
        const float4 x4 = *reinterpret_cast<const float4*>(&xs[sx + k][qc * 16 + c]);
        a2 = __ffma2_rn(make_float2(w4.x, w4.y), make_float2(x4.x, x4.y), a2);
        b2 = __ffma2_rn(make_float2(w4.z, w4.w), make_float2(x4.z, x4.w), b2);
      }
    acc = (a2.x + a2.y) + (b2.x + b2.y);
  } else {
#pragma unroll
    for (int k = 0; k < 7; ++k)
#pragma unroll
      for (int c = 0; c < 16; c += 4) {
        const float4 w4 = *reinterpret_cast<const float4*>(&ws[k][qc * 16 + c]);   // warp-uniform: broadcast
        const float4 x4 = *reinterpret_cast<const float4*>(&xs[sx + k][qc * 16 + c]);
        acc = fmaf(w4.x, x4.x, acc); acc = fmaf(w4.y, x4.y, acc); acc = fmaf(w4.z, x4.z, acc); acc = fmaf(w4.w, x4.w, acc);
      }
  }
  part[qc][sx] = acc;
  __syncthreads();
  if (qc != 0) return;
  const int t_rel = t0 + sx;
  if (t_rel >= a.out_r.hi) return;
  const int t_abs = t_rel + it.shift0 * 512;
  if (t_abs < 0 || t_abs >= T0 * 512) return;
  if (a.status && a.status[it.code_row] != SNACB_WIN_OK) return;
  float y = tanhf(((part[0][sx] + part[1][sx]) + (part[2][sx] + part[3][sx])) + a.bias[0]);
  const long long d = it.dst + (t_rel - a.out_r.lo);
  if (!(fabsf(y) <= 1.0f)) {  // NaN (an overflowed activation upstream): never emitted, the window is reported instead
    y = 0.0f;
    if (a.status) a.status[it.code_row] = SNACB_WIN_NONFINITE;
  }
  if (a.wav) a.wav[d] = y;
  if (a.pcm) a.pcm[d] = (int16_t)(y * 32767.0f);
}

void launch_tail(const GroupCtx& g, const TailArgs& a) {
  dim3 grid((a.out_r.n() + 63) / 64, g.n_items);
  if (a.fast) k_tail<true><<<grid, 256, 0, g.stream>>>(g.items, g.base, g.out_len, g.T0, a);
  else k_tail<false><<<grid, 256, 0, g.stream>>>(g.items, g.base, g.out_len, g.T0, a);
  ++*g.launches;
}

// ============================================================================ Philox noise dump
__global__ void k_fill_noise(unsigned long long seed, const unsigned long long* keys, int F, float* out,
                             long long stride) {
  const int w = blockIdx.y;
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= kNoisePerFrame * F) return;
  const int len[4] = {32 * F, 256 * F, 1024 * F, 2048 * F};
  int b = 0, t = idx;
  while (t >= len[b]) { t -= len[b]; ++b; }
  const unsigned long long key = keys ? keys[w] : (unsigned long long)w;
  out[(long long)w * stride + idx] = philox_normal(seed, key, (uint32_t)b, (uint32_t)t);
}

void launch_fill_noise(uint64_t seed, const unsigned long long* d_keys, int n_win, int F, float* d_noise,
                       long long stride, cudaStream_t st, int64_t* launches) {
  if (n_win <= 0 || F <= 0) return;
  dim3 grid((kNoisePerFrame * F + 255) / 256, n_win);
  k_fill_noise<<<grid, 256, 0, st>>>((unsigned long long)seed, d_keys, F, d_noise, stride);
  ++*launches;
}

}  // namespace snacb
