// SNAC-24k ENCODER: audio -> latent -> residual vector quantisation -> the 3 code levels the decode
// path consumes.  Restates the third-party `snac` package's `SNAC.encode` (oracle: oracle/snac_ref.py Encoder /
// ResidualVectorQuantize.encode; the reference points at it from the finetune / data-prep flow,
// Orpheus-TTS/README.md:111-120): Conv1d 1->48 k7, four EncoderBlocks (ResidualUnit d = 1, 3, 9 -> Snake -> strided
// Conv1d k = 2s, stride s = 2/4/8/8), depthwise Conv1d k7 on 768 channels, then per level: avg_pool (stride 4/2/1) ->
// in_proj 768->8 -> nearest codebook entry on L2-normalised vectors -> out_proj -> residual update.
//
// This is a voice-prompt / data-prep path, not the decode hot path: everything runs as exact fp32 CUDA-core kernels,
// re-using the fp32 recipe's depthwise and GEMM kernels (ResidualUnits, the strided conv as a GEMM over overlapping
// K-windows of a zero-padded Snake'd copy).  The kernels below are the pieces the decoder does not have.
#include "kernels.h"
#include "snacb.h"

namespace snacb {
namespace {

// audio [B][T] -> [B][T][48] : Conv1d(1 -> 48, k7, pad 3) + bias
__global__ void __launch_bounds__(256) k_enc_in(const float* __restrict__ audio, int T, const float* __restrict__ w /*[48][7]*/,
                                                const float* __restrict__ bias, float* __restrict__ out) {
  const int b = blockIdx.y;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)T * 48) return;
  const int t = (int)(idx / 48), c = (int)(idx - (long long)t * 48);
  const float* x = audio + (size_t)b * T;
  float acc = 0.0f;
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    const int u = t + k - 3;
    const float v = (u >= 0 && u < T) ? x[u] : 0.0f;
    acc = fmaf(w[c * 7 + k], v, acc);
  }
  out[((size_t)b * T + t) * 48 + c] = acc + bias[c];
}

// Snake then zero padding: in [B][T][C] -> out [B][T + s][C], row j holds Snake(x[j - p]) (zero outside [0, T)):
// the strided conv's K-window of output t is then the contiguous run of 2s rows starting at row s * t.
__global__ void __launch_bounds__(256) k_snake_pad(const float* __restrict__ in, float* __restrict__ out, int T, int C, int s, int p,
                                                   const float* __restrict__ alpha, const float* __restrict__ inv) {
  const int b = blockIdx.y;
  const long long idx = ((long long)blockIdx.x * 256 + threadIdx.x) * 4;
  if (idx >= (long long)(T + s) * C) return;
  const int j = (int)(idx / C), c = (int)(idx - (long long)j * C);
  const int t = j - p;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (t >= 0 && t < T) {
    v = *reinterpret_cast<const float4*>(in + ((size_t)b * T + t) * C + c);
    const float4 al = *reinterpret_cast<const float4*>(alpha + c), iv = *reinterpret_cast<const float4*>(inv + c);
    v.x = snake_exact(v.x, al.x, iv.x); v.y = snake_exact(v.y, al.y, iv.y);
    v.z = snake_exact(v.z, al.z, iv.z); v.w = snake_exact(v.w, al.w, iv.w);
  }
  *reinterpret_cast<float4*>(out + ((size_t)b * (T + s) + j) * C + c) = v;
}

// One residual-VQ level for one pooled time step: avg_pool -> in_proj -> normalise -> argmax over the normalised
// codebook of -(|e|^2 - 2 e.c + |c|^2) (first index wins ties, like torch.max) -> residual -= out_proj(codebook[idx]).
__global__ void __launch_bounds__(256) k_vq_level(float* __restrict__ residual /*[B][T][768]*/, int T, int stride,
                                                  const float* __restrict__ w_in /*[8][768]*/, const float* __restrict__ b_in,
                                                  const float* __restrict__ cb_norm /*[4096][8]*/, const float* __restrict__ cb /*[4096][8]*/,
                                                  const float* __restrict__ w_out /*[768][8]*/, const float* __restrict__ b_out,
                                                  int32_t* __restrict__ codes /*[B][T/stride]*/) {
  __shared__ float pooled[kLatent];
  __shared__ float ze[8];
  __shared__ float best_v[8];
  __shared__ int best_i[8];
  __shared__ float emb[8];
  const int tp = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Tp = T / stride;
  float* r0 = residual + ((size_t)b * T + (size_t)tp * stride) * kLatent;
  const float inv_s = 1.0f / (float)stride;
  for (int c = tid; c < kLatent; c += 256) {
    float a = 0.0f;
    for (int j = 0; j < stride; ++j) a += r0[(size_t)j * kLatent + c];
    pooled[c] = a * inv_s;
  }
  __syncthreads();
  {  // 8 warps, one projected dimension each
    float a = 0.0f;
    for (int c = lane; c < kLatent; c += 32) a = fmaf(w_in[warp * kLatent + c], pooled[c], a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) ze[warp] = a + b_in[warp];
  }
  __syncthreads();
  float e[8], n2 = 0.0f;
#pragma unroll
  for (int d = 0; d < 8; ++d) { e[d] = ze[d]; n2 = fmaf(e[d], e[d], n2); }
  const float rn = 1.0f / fmaxf(sqrtf(n2), 1e-12f);  // F.normalize: x / max(||x||, eps)
  float e2 = 0.0f;
#pragma unroll
  for (int d = 0; d < 8; ++d) { e[d] *= rn; e2 = fmaf(e[d], e[d], e2); }
  float bv = -3.0e38f;
  int bi = 0;
  for (int k = tid; k < SNACB_CODEBOOK_SIZE; k += 256) {
    const float4 c0 = *reinterpret_cast<const float4*>(cb_norm + (size_t)k * 8), c1 = *reinterpret_cast<const float4*>(cb_norm + (size_t)k * 8 + 4);
    const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
    float dot = 0.0f, c2 = 0.0f;
#pragma unroll
    for (int d = 0; d < 8; ++d) { dot = fmaf(e[d], cc[d], dot); c2 = fmaf(cc[d], cc[d], c2); }
    const float score = -((e2 - 2.0f * dot) + c2);
    if (score > bv) { bv = score; bi = k; }  // k ascending per thread: first maximum kept
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  if (lane == 0) { best_v[warp] = bv; best_i[warp] = bi; }
  __syncthreads();
  if (tid == 0) {
    float v = best_v[0];
    int i = best_i[0];
    for (int w = 1; w < 8; ++w)
      if (best_v[w] > v || (best_v[w] == v && best_i[w] < i)) { v = best_v[w]; i = best_i[w]; }
    best_i[0] = i;
    codes[(size_t)b * Tp + tp] = i;
  }
  __syncthreads();
  const int idx = best_i[0];
  if (tid < 8) emb[tid] = cb[(size_t)idx * 8 + tid];
  __syncthreads();
  for (int c = tid; c < kLatent; c += 256) {
    float q = b_out[c];
#pragma unroll
    for (int d = 0; d < 8; ++d) q = fmaf(w_out[c * 8 + d], emb[d], q);
    for (int j = 0; j < stride; ++j) r0[(size_t)j * kLatent + c] -= q;
  }
}

}  // namespace

void launch_enc_in(const float* audio, int B, int T, const float* w, const float* bias, float* out, cudaStream_t st, int64_t* launches) {
  dim3 grid((unsigned)(((long long)T * 48 + 255) / 256), B);
  k_enc_in<<<grid, 256, 0, st>>>(audio, T, w, bias, out);
  ++*launches;
}
void launch_snake_pad(const float* in, float* out, int B, int T, int C, int s, int p, const float* alpha, const float* inv,
                      cudaStream_t st, int64_t* launches) {
  dim3 grid((unsigned)(((long long)(T + s) * C / 4 + 255) / 256), B);
  k_snake_pad<<<grid, 256, 0, st>>>(in, out, T, C, s, p, alpha, inv);
  ++*launches;
}
void launch_vq_level(float* residual, int B, int T, int stride, const float* w_in, const float* b_in, const float* cb_norm,
                     const float* cb, const float* w_out, const float* b_out, int32_t* codes, cudaStream_t st, int64_t* launches) {
  dim3 grid(T / stride, B);
  k_vq_level<<<grid, 256, 0, st>>>(residual, T, stride, w_in, b_in, cb_norm, cb, w_out, b_out, codes);
  ++*launches;
}

}  // namespace snacb
