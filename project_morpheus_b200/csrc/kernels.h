// Launcher declarations shared by engine.cu and the kernel translation units.
#pragma once
#include <cuda_fp16.h>

#include "snacb_common.cuh"

namespace snacb {

// Group-wide context handed to every launcher.
struct GroupCtx {
  const Item* items;  // device pointer already offset to the chunk's first item, or nullptr
  int base;           // implicit-item base row (chunk start) when items == nullptr
  int n_items;        // items in this chunk
  int out_len;        // samples per item in the destination (implicit dst = row * out_len)
  int T0;             // latent steps of the whole sequence
  cudaStream_t stream;
  int64_t* launches;  // launch counter
  int flags = 0;      // snacb_config.flags (kernel selection switches)
};

// ---- integer kernel ------------------------------------------------------------------------
void launch_deinterleave(const int32_t* d_tokens, int tokens_stride, const int32_t* d_ntok,
                         int ntok_uniform, int n_win, int max_frames, bool raw, int32_t* c0,
                         int32_t* c1, int32_t* c2, int32_t* status, cudaStream_t st, int64_t* launches);

// ---- fp32 CUDA-core layer kernels ------------------------------------------------------------
struct QuantW {
  const float* codebook[3];
  const float* w[3];  // [768][8]
  const float* b[3];  // [768]
};
void launch_from_codes(const GroupCtx& g, const QuantW& q, const int32_t* c0, const int32_t* c1,
                       const int32_t* c2, int pitch0, Rng z, float* out);

struct DwArgs {
  const float* in; Rng in_r;
  float* out; Rng out_r;
  int C, dil, up;
  const float* w7;  // [7][C]
  const float* bias;
  const float* a1; const float* i1;  // pre-Snake alpha, 1/(alpha+1e-9); nullptr = none
  const float* a2; const float* i2;  // post-Snake
};
void launch_dwconv(const GroupCtx& g, const DwArgs& a);

void launch_snake(const GroupCtx& g, const float* in, float* out, Rng r, int C, const float* alpha,
                  const float* inv);

enum { EPI_BIAS = 0, EPI_RESID = 1, EPI_NOISE = 2, EPI_CONVT = 3 };
struct GemmArgs {
  int epi;
  const float* A; int lda; Rng a_r;   // operand rows (per item) and its K-major row pitch
  const float* W; int ldw;            // [N][ldw]
  const float* bias;                  // [N] (EPI_CONVT: [Cout])
  int K, N;
  Rng m_r;                            // rows iterated per item (EPI_CONVT: positions q)
  int s, p, Cout;                     // EPI_CONVT only
  float* out; Rng o_r; int ldo;
  const float* R; Rng r_r; int ldr;   // EPI_RESID residual / EPI_NOISE carrier x
  NoiseSrc noise;
  int up;                             // time scale of the OUTPUT rows
};
void launch_gemm_f32(const GroupCtx& g, const GemmArgs& a);

// ---- tensor-core recipe (kernels_tc.cu) ------------------------------------------------------
struct TcGemmArgs {
  int epi;                         // EPI_*
  const __half* A; int K;          // operand [n_items * a_rows][K] fp16, K contiguous; K per segment
  int a_rows, a_lo;                // rows per item (all are iterated) and relative time of row 0
  const __half* W; int N;          // [N][nseg*K] fp16 (EPI_CONVT: nseg = 2, N = s*Cout)
  const float* bias;               // [N] (EPI_CONVT: [Cout]); nullable
  int s, p, Cout;                  // EPI_CONVT only
  float* out32; __half* out16;     // either may be null; same row pitch ldo
  Rng o_r; int ldo;
  const float* sn_alpha; const float* sn_inv;  // Snake applied to the fp16 output (nullable)
  const float* R; Rng r_r; int ldr;            // EPI_RESID residual / EPI_NOISE carrier (fp32)
  NoiseSrc noise;
  int up;
  // split-operand recipe (SNACB_PREC_FP16X3): A is [rows][2K] = [hi | lo] fp16 halves of the fp32 operand, W is
  // [N][nseg * 3K] = per tap [W_hi | W_lo | W_hi]; the GEMM runs over the concatenated K and yields
  // A_hi W_hi + A_hi W_lo + A_lo W_hi (the fp32 product to ~2^-22).  Forces the generic one-tile kernel.
  int split = 0;
};
cudaError_t launch_gemm_tc(const GroupCtx& g, const TcGemmArgs& a);
// composed ConvT + NoiseBlock for the wide blocks (Cout % 256 == 0): one GEMM over a stacked weight [conv rows | W_n conv rows]
bool convt_noise2_supported(int Cin, int Cout);
void launch_compose_ctn(const float* ct, const float* wn, int Cout, int N, int K2, __half* out, cudaStream_t st);
cudaError_t launch_convt_noise2_tc(const GroupCtx& g, const TcGemmArgs& a, const __half* W2, const float* bias2);
// fp32 [rows][C] (-> optional Snake) -> fp16 [rows][2C] = [hi | lo]: the split operand of the recipe above
void launch_split16(const float* in, __half* out, size_t rows, int C, const float* alpha, const float* inv, cudaStream_t st,
                    int64_t* launches);
// weight [N][nseg * K] fp32 -> [N][nseg * 3K] fp16 = per segment [hi | lo | hi]
void launch_split_w(const float* in, __half* out, int N, int nseg, int K, cudaStream_t st);
int tc_tile_n(const TcGemmArgs& a);
// ConvTranspose1d + NoiseBlock in one kernel (blocks whose Cout is 64 / 128)
bool convt_noise_supported(int Cin, int Cout);
cudaError_t launch_convt_noise_tc(const GroupCtx& g, const TcGemmArgs& convt, const __half* noise_w16);

struct DwTcArgs {
  const float* in; Rng in_r;
  __half* out; Rng out_r;
  int C, dil, up;
  const float* w7; const float* bias;
  const float* a1; const float* i1; const float* a2; const float* i2;
};
void launch_dw_tc(const GroupCtx& g, const DwTcArgs& a);
// Fused ResidualUnit (C = 64 / 128): x' = x + W * Snake(dw7(Snake(x))) + b in one kernel.
struct RuTcArgs {
  const float* x; Rng in_r;        // residual stream in (fp32)
  Rng out_r; int C, dil, up;
  const float* w7; const float* dw_b; const float* a1; const float* i1; const float* a2; const float* i2;
  const __half* pw16; const float* pw_b;
  float* out32; __half* out16; const float* sn_alpha; const float* sn_inv;
  int prefetch_ahead;
  bool persistent;   // k_ru_p (persistent, warp-specialised, also C = 256) instead of k_ru_tc
  // non-null tail_w7: fuse the decoder tail (Snake = sn_alpha/sn_inv, conv k7 64->1, tanh, slice, int16 pack)
  const float* tail_w7; const float* tail_b; Rng tail_out; int32_t* status; float* wav; int16_t* pcm;
};
bool ru_tc_supported(int C, bool persistent);
// wide fused ResidualUnit (C = 256, kernels_ruw.cu): TMA-staged input blocks, residual stream initialised in tensor memory
bool ruw_tc_supported(int C);
cudaError_t launch_ruw_tc(const GroupCtx& g, const RuTcArgs& a);
cudaError_t launch_ru_tc(const GroupCtx& g, const RuTcArgs& a);
// Whole DecoderBlock ResidualUnit chain (+ decoder tail and PCM pack) in one persistent kernel (kernels_blk.cu): the
// residual stream stays in tensor memory, Snake'd activations in shared memory.  x = ConvTranspose1d + NoiseBlock output.
struct BlkTcArgs {
  const float* x; Rng in_r; int C, up;
  struct Ru { const float *w7, *dw_b, *a1, *i1, *a2, *i2, *pw_b, *pw_w; const __half* pw16; } ru[3];
  const float* sn_alpha; const float* sn_inv;  // Snake after the block (decoder tail)
  const float* tail_w7; const float* tail_b; Rng tail_out; int32_t* status; float* wav; int16_t* pcm;
};
bool blk_tc_supported(int C, bool tail);
cudaError_t launch_blk_tc(const GroupCtx& g, const BlkTcArgs& a);
// from_codes + decoder.model.0 (depthwise k7) -> fp16 operand of the 768->1024 GEMM, z never leaves the SM
bool codes_head_supported(Rng z);
void launch_codes_head(const GroupCtx& g, const QuantW& q, const int32_t* c0, const int32_t* c1, const int32_t* c2, int pitch0,
                       Rng z, Rng h, const float* w7, const float* dw_b, __half* out);
void launch_dwconv_half(const GroupCtx& g, const DwArgs& a, __half* out16);  // plain dw k7 -> fp16 (decoder head)
void launch_to_half(const float* in, __half* out, size_t n, cudaStream_t st);

struct TailArgs {
  const float* x; Rng x_r;  // [item][rows][64]
  Rng out_r;
  const float* alpha; const float* inv;
  const float* w7;  // [7][64]
  const float* bias;  // [1]
  int32_t* status;  // per code_row; nullptr or skip rows whose status != 0; set to SNACB_WIN_NONFINITE by the kernel
  float* wav;   // nullable
  int16_t* pcm; // nullable
  bool fast;    // MUFU sin for the Snake (tensor-core recipe)
};
void launch_tail(const GroupCtx& g, const TailArgs& a);

// ---- encoder (kernels_enc.cu)
void launch_enc_in(const float* audio, int B, int T, const float* w, const float* bias, float* out, cudaStream_t st, int64_t* launches);
void launch_snake_pad(const float* in, float* out, int B, int T, int C, int s, int p, const float* alpha, const float* inv,
                      cudaStream_t st, int64_t* launches);
void launch_vq_level(float* residual, int B, int T, int stride, const float* w_in, const float* b_in, const float* cb_norm,
                     const float* cb, const float* w_out, const float* b_out, int32_t* codes, cudaStream_t st, int64_t* launches);

void launch_fill_noise(uint64_t seed, const unsigned long long* d_keys, int n_win, int F, float* d_noise,
                       long long stride, cudaStream_t st, int64_t* launches);

}  // namespace snacb
