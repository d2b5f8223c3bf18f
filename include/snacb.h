/*
 * snacb.h - C ABI of the B200-native SNAC-24k token->waveform engine (libsnacb.so).
 *
 * The reference (DocWobble/Project_Morpheus) is pure Python and has NO FFI: its hot path is
 *   Morpheus_Client/tts_engine/speechpipe.py:64-137   convert_to_audio(multiframe, count)
 *   Morpheus_Client/tts_engine/speechpipe.py:118       model.decode(codes)   (third-party `snac`)
 * so this header is the boundary a maintainer would bind with ctypes; every
 * entry point cites the reference lines it replaces.  Plain C types only: pointers and sizes, no
 * C++/torch types, no exceptions across the boundary.
 *
 * Conventions
 *   - "d_" pointers are device memory on the engine's GPU, "h_" pointers are host memory.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  All work is
 *     enqueued asynchronously on it unless stated; a handle is bound to one device and is not
 *     re-entrant (one handle per GPU worker).
 *   - Return value: 0 (SNACB_OK) or a negative SNACB_E* code; snacb_last_error() gives the text.
 *   - The caller owns every buffer it passes; the engine owns packed weights and its workspace.
 *   - There is no CPU fallback: without a CUDA device snacb_create() fails.
 */
#ifndef SNACB_H
#define SNACB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNACB_ABI_VERSION 1

/* return codes */
#define SNACB_OK 0
#define SNACB_EINVAL (-1)   /* bad argument */
#define SNACB_ECUDA (-2)    /* CUDA runtime / driver error */
#define SNACB_ESTATE (-3)   /* weights not loaded, engine misuse */
#define SNACB_ENOMEM (-4)

/* per-window status written by the integer kernel (speechpipe.py:69-70,108-111,122) */
#define SNACB_WIN_OK 0        /* decoded, 2048 int16 samples written                               */
#define SNACB_WIN_REJECTED 1  /* reference returns None: < 7 tokens, or a code < 0 or > 4096        */
#define SNACB_WIN_CODE4096 2  /* code == 4096 passes the reference validator, then F.embedding raises */
#define SNACB_WIN_EMPTY 3     /* one whole frame only: slice [2048:4096) is empty, reference returns b'' */
#define SNACB_WIN_NONFINITE 4 /* the decode produced a non-finite sample (fp16 operand range exceeded): PCM withheld, see SNACB_PREC_* */

/* noise modes for NoiseBlock (x + randn[B,1,T] * W_n x inside the third-party decoder) */
#define SNACB_NOISE_OFF 0     /* zeros                                                  */
#define SNACB_NOISE_TENSOR 1  /* caller-provided noise, layout below (parity mode A)    */
#define SNACB_NOISE_PHILOX 2  /* counter-based Philox4x32-10 keyed by (seed,key,block,t) */

/* arithmetic recipes for the GEMM-shaped layers (1x1 convs, ConvTranspose1d) */
#define SNACB_PREC_FP32 0     /* CUDA-core fp32 FMA: the exact/bring-up path              */
#define SNACB_PREC_FP16 1     /* tcgen05 kind::f16, fp16 operands, fp32 TMEM accumulators */
#define SNACB_PREC_FP16X3 2   /* tcgen05 kind::f16 on two-term fp16 splits of both operands (hi*hi + hi*lo + lo*hi): fp32-grade
                                 products for checkpoints that outgrow single-pass fp16; exact sinf Snake, fp32 depthwise */

/* snacb_config.flags */
#define SNACB_FLAG_NO_RU_FUSION 1 /* tensor-core recipe: run every ResidualUnit as dw kernel + GEMM kernel */
#define SNACB_FLAG_PERSISTENT_RU 4 /* persistent warp-specialised ResidualUnit kernel (also fuses C = 256)   */
#define SNACB_FLAG_TAIL_FUSION 8 /* fuse the decoder tail into the last ResidualUnit kernel (measured slower: off) */
#define SNACB_FLAG_NO_CONVT_NOISE_FUSION 2 /* run ConvTranspose1d and NoiseBlock as two GEMM kernels   */
#define SNACB_FLAG_NO_PERSISTENT_CONVT 16 /* transposed convs through the one-shot (one tile per CTA) kernels  */
#define SNACB_FLAG_NO_BLOCK_FUSION 64 /* decoder block 3: one kernel per ResidualUnit + tail kernel instead of the whole-block kernel */
#define SNACB_FLAG_NO_CONVT_NOISE_COMPOSE 128 /* decoder blocks 0 / 1: ConvTranspose1d and NoiseBlock as two kernels instead of one GEMM over the composed weight */
#define SNACB_FLAG_FUSE_RU256 32 /* decoder block 1 ResidualUnits through the persistent fused kernel (measured on par: off) */

/* fixed geometry of hubertsiuzdak/snac_24khz (the only model the reference loads, speechpipe.py:42) */
#define SNACB_LATENT 768
#define SNACB_DECODER_DIM 1024
#define SNACB_CODEBOOK_SIZE 4096
#define SNACB_CODEBOOK_DIM 8
#define SNACB_TOKENS_PER_FRAME 7
#define SNACB_SAMPLES_PER_FRAME 2048
#define SNACB_NOISE_PER_FRAME 3360 /* 32+256+1024+2048 noise values per frame per window */

typedef struct snacb_engine snacb_engine;

typedef struct snacb_config {
  int32_t abi_version;   /* SNACB_ABI_VERSION */
  int32_t device;        /* CUDA ordinal */
  int32_t precision;     /* SNACB_PREC_* */
  int32_t chunk_items;   /* windows processed per pass through the layer stack (0 = default);
                            sized so one pass's activations stay L2-resident */
  int32_t trim;          /* 1 = compute only the dependency cone of the emitted slice (exact) */
  int32_t flags;         /* SNACB_FLAG_* (bring-up switches; 0 = production)                  */
  int32_t lanes;         /* chunks of one call that run concurrently on internal streams (0/1 = none) */
  int32_t reserved[9];
} snacb_config;

/* One residual unit: x + W_pw * Snake(dw7_dil(Snake(x))) */
typedef struct snacb_ru_weights {
  const float* alpha1; /* [C]      Snake alpha before the depthwise conv           */
  const float* dw_w;   /* [C][7]   depthwise k=7 weight, weight-norm folded        */
  const float* dw_b;   /* [C]                                                       */
  const float* alpha2; /* [C]                                                       */
  const float* pw_w;   /* [C][C]   1x1 conv weight [Cout][Cin], folded             */
  const float* pw_b;   /* [C]                                                       */
} snacb_ru_weights;

typedef struct snacb_block_weights {
  const float* alpha;    /* [Cin]                block-head Snake                            */
  const float* convt_w;  /* [Cin][Cout][2*s]     ConvTranspose1d weight, PyTorch layout,
                                                 folded per INPUT channel (weight_norm dim 0) */
  const float* convt_b;  /* [Cout]                                                           */
  const float* noise_w;  /* [Cout][Cout]         NoiseBlock 1x1, no bias                     */
  snacb_ru_weights ru[3]; /* dilations 1, 3, 9 */
} snacb_block_weights;

/* Host pointers to fp32 arrays, weight-norm already folded (w = g*v/||v||).  Copied and
 * re-packed by snacb_load_weights(); may be freed afterwards. */
typedef struct snacb_weights {
  const float* codebook[3];  /* [4096][8]  RVQ level l codebook                      */
  const float* outproj_w[3]; /* [768][8]   out_proj 1x1 (8 -> latent)                */
  const float* outproj_b[3]; /* [768]                                                */
  const float* head_dw_w;    /* [768][7]   decoder.model.0 (depthwise k7)            */
  const float* head_dw_b;    /* [768]                                                */
  const float* head_pw_w;    /* [1024][768] decoder.model.1                          */
  const float* head_pw_b;    /* [1024]                                               */
  snacb_block_weights block[4]; /* strides 8,8,4,2; channels 1024->512->256->128->64 */
  const float* tail_alpha;   /* [64]       decoder.model.6                           */
  const float* tail_w;       /* [64][7]    decoder.model.7 (64 -> 1, k7)             */
  const float* tail_b;       /* [1]                                                  */
} snacb_weights;

/* ---- lifetime ------------------------------------------------------------------------------ */

/* Replaces the module-level model load + device placement, speechpipe.py:38-61. */
int snacb_create(snacb_engine** out, const snacb_config* cfg);
void snacb_destroy(snacb_engine* e);
/* Text of the last error on this engine (or of the last failed snacb_create when e == NULL). */
const char* snacb_last_error(const snacb_engine* e);
/* Replaces SNAC.from_pretrained(...).to(device) weight upload, speechpipe.py:43,49 (weight-norm
 * folded once instead of on every forward). Synchronous. */
int snacb_load_weights(snacb_engine* e, const snacb_weights* w);
/* Bytes of device workspace currently held by the engine. */
size_t snacb_workspace_bytes(const snacb_engine* e);
/* Number of kernels this engine has launched since creation (bench.py's gpu_launches claim). */
int64_t snacb_launch_count(const snacb_engine* e);
/* Number of CUDA-graph replays: small uniform ticks of snacb_decode_windows_host (<= 256 windows, env
 * SNACB_GRAPHS=<max windows>, 0 = off) are captured once per shape and replayed as ONE launch. */
int64_t snacb_graph_launch_count(const snacb_engine* e);

/* ---- NS-1: integer de-interleave + validate ------------------------------------------------ */

/* Replaces speechpipe.py:72-111 (frame truncation, de-interleave loop, range validator), batched.
 *   d_tokens  [n_win][tokens_stride] int32 token ids (offsets already removed, i.e. what
 *             turn_token_into_id returns); window i uses its first h_ntok[i] entries
 *             (h_ntok == NULL: every window has ntok_uniform tokens).
 *   F_i = ntok_i / 7 whole frames; codes are written densely at row pitch max_frames:
 *   d_c0 [n_win][max_frames], d_c1 [n_win][2*max_frames], d_c2 [n_win][4*max_frames] (int32,
 *   entries beyond F_i zeroed), d_status [n_win] = SNACB_WIN_*.  Bit-exact vs the reference. */
int snacb_deinterleave(snacb_engine* e, const int32_t* d_tokens, int32_t tokens_stride,
                       const int32_t* h_ntok, int32_t ntok_uniform, int32_t n_win,
                       int32_t max_frames, int32_t* d_c0, int32_t* d_c1, int32_t* d_c2,
                       int32_t* d_status, void* stream);

/* Raw-token variant (north_star item 1): d_raw holds N of "<custom_token_N>" for ALIGNED streams,
 * position p of window i has slot (p % 7); id = N - 10 - 4096*(p%7) (speechpipe.py:181). */
int snacb_deinterleave_raw(snacb_engine* e, const int32_t* d_raw, int32_t tokens_stride,
                           const int32_t* h_ntok, int32_t ntok_uniform, int32_t n_win,
                           int32_t max_frames, int32_t* d_c0, int32_t* d_c1, int32_t* d_c2,
                           int32_t* d_status, void* stream);

/* ---- the streaming path: tokens -> PCM ------------------------------------------------------ */

/* Replaces convert_to_audio (speechpipe.py:64-137) for a whole decode tick: de-interleave +
 * validate + SNAC decode + slice [2048:4096) + trunc(x*32767) int16 pack, one call for all windows.
 *   d_tokens/h_ntok/ntok_uniform as above (any mix of window lengths; windows are grouped by
 *   frame count internally).
 *   d_noise   SNACB_NOISE_TENSOR: [n_win][noise_stride] float32, window i holds the four noise
 *             rows of an F_i-frame decode back to back (lengths 32F,256F,1024F,2048F), i.e.
 *             torch.cat([n_b[i,0,:] for b in 0..3]); noise_stride >= 3360*max F.  Else NULL.
 *   seed      SNACB_NOISE_PHILOX key; h_keys (optional, [n_win] uint64) distinguishes streams.
 *   d_pcm     [n_win][2048] int16; rows of windows whose status != SNACB_WIN_OK are zeroed.
 *   d_status  [n_win] int32 SNACB_WIN_*.  A bad window never poisons the batch. */
int snacb_decode_windows(snacb_engine* e, const int32_t* d_tokens, int32_t tokens_stride,
                         const int32_t* h_ntok, int32_t ntok_uniform, int32_t n_win,
                         int32_t noise_mode, const float* d_noise, int64_t noise_stride,
                         uint64_t seed, const uint64_t* h_keys, int16_t* d_pcm, int32_t* d_status,
                         void* stream);

/* Same contract with HOST buffers (pageable or pinned): stages through the engine's pinned
 * buffers, H2D + kernels + D2H on `stream`, returns after the PCM and status are in h_pcm/h_status.
 * This is the call the Python convert_to_audio / convert_to_audio_batch make. */
int snacb_decode_windows_host(snacb_engine* e, const int32_t* h_tokens, int32_t tokens_stride,
                              const int32_t* h_ntok, int32_t ntok_uniform, int32_t n_win,
                              int32_t noise_mode, const float* h_noise, int64_t noise_stride,
                              uint64_t seed, const uint64_t* h_keys, int16_t* h_pcm,
                              int32_t* h_status, void* stream);

/* ---- the one-shot path: codes -> waveform --------------------------------------------------- */

/* Replaces model.decode(codes) (speechpipe.py:118; quantizer.from_codes + decoder) for B sequences
 * of F frames: d_c0 [B][F], d_c1 [B][2F], d_c2 [B][4F] int32 in [0,4095] (caller validated).
 * Writes d_wav [B][2048F] float32 and/or d_pcm [B][2048F] int16 (either may be NULL).  Long
 * sequences are time-tiled with halo recompute internally.  d_noise: [B][3360F] as above. */
int snacb_decode_codes(snacb_engine* e, const int32_t* d_c0, const int32_t* d_c1,
                       const int32_t* d_c2, int32_t B, int32_t F, int32_t noise_mode,
                       const float* d_noise, uint64_t seed, float* d_wav, int16_t* d_pcm,
                       void* stream);

/* Writes the noise SNACB_NOISE_PHILOX would use for (seed, keys) in the SNACB_NOISE_TENSOR layout,
 * so the oracle can replay a production decode (byte-identical replay under a fixed seed). */
int snacb_fill_noise(snacb_engine* e, uint64_t seed, const uint64_t* h_keys, int32_t n_win,
                     int32_t F, float* d_noise, int64_t noise_stride, void* stream);

/* ---- measurement ---------------------------------------------------------------------------- */

/* Per-kernel-class device timing (CUDA events on the launching stream around every launch of the
 * class) for bench.py's roofline object.  Off by default; when on, every decode call records two
 * events per launch.  snacb_profile_read() synchronises the recorded events, ACCUMULATES them into
 * per-class totals, writes up to `cap` rows and returns the number of classes (negative on error);
 * snacb_profile_enable(e, 1) clears the totals. */
typedef struct snacb_kernel_stat {
  char name[32];      /* kernel class, e.g. "gemm_1x1", "gemm_convt", "dwconv", "tail" */
  int64_t launches;
  double ms;          /* summed device time of the launches                                  */
  double flops;       /* executed floating-point operations (2*MAC for GEMM-shaped kernels)  */
  double bytes;       /* algorithmic global-memory bytes read + written by the launches      */
} snacb_kernel_stat;
int snacb_profile_enable(snacb_engine* e, int32_t on);
int snacb_profile_read(snacb_engine* e, snacb_kernel_stat* out, int32_t cap);

/* ---- bring-up / parity taps ----------------------------------------------------------------- */

/* Layer-wise parity: after the next decode, d_buf receives the fp32 activation of `stage`
 * (channels-last [items][rows][channels]) for the first chunk. stage < 0 disables the tap.
 * Stage ids: 0 z, 1 head dw, 2 head 1x1, 3+9b+{0 snake,1 convT,2 noise,3..8 (dw,out) of RU 0..2}. */
int snacb_set_tap(snacb_engine* e, int32_t stage, float* d_buf, size_t capacity_floats);
/* Geometry of the last captured tap: rows per item, channels, first row's time index, items. */
int snacb_get_tap_shape(const snacb_engine* e, int32_t* rows, int32_t* channels, int32_t* t_lo,
                        int32_t* items);

/* Host-only: the row ranges the engine computes for an F-frame sequence when samples
 * [out_lo, out_hi) are requested (exact dependency cone).  ranges receives 26
 * (lo, hi) pairs: z, head, then per decoder block: in, q, convT, ru0, ru1, ru2. Needs no GPU. */
int snacb_plan(int32_t frames, int32_t out_lo, int32_t out_hi, int32_t clip, int32_t* ranges);

/* Pipelined form of snacb_decode_windows_host for throughput: _submit copies the inputs, enqueues H2D + kernels on
 * `stream` and the PCM / status D2H on an internal copy stream, and returns a ticket without waiting; _wait blocks until
 * that tick's h_pcm / h_status (the pointers given to _submit; pinned ones are written directly) are complete.  At most
 * two ticks in flight (submit t+1, then wait t): the D2H and the host work of tick t overlap the kernels of tick t+1.
 * noise_mode: SNACB_NOISE_OFF or SNACB_NOISE_PHILOX. */
int snacb_decode_windows_host_submit(snacb_engine* e, const int32_t* h_tokens, int32_t tokens_stride, const int32_t* h_ntok,
                                     int32_t ntok_uniform, int32_t n_win, int32_t noise_mode, uint64_t seed,
                                     const uint64_t* h_keys, int16_t* h_pcm, int32_t* h_status, void* stream, int32_t* ticket);
int snacb_decode_windows_host_wait(snacb_engine* e, int32_t ticket);

/* ---- N2: token ingress for many streams (host only, no GPU) ---------------------------------- */

/* ---- SNAC encoder: audio -> the 3 code levels `snacb_decode_codes` consumes.  Replaces the third-party
 * `snac` package's SNAC.encode (referenced from Orpheus-TTS/README.md:111-120 data preparation); fp32 CUDA-core kernels.
 * All tensors folded fp32 host arrays in PyTorch layouts: in_w [48][7] (Conv1d 1->48), block b (C = 48 * 2^b, stride
 * 2/4/8/8): ru[3] as in the decoder, alpha [C], down_w [2C][C][2s], down_b [2C]; out_dw_w [768][7]; inproj_w [8][768]. */
typedef struct snacb_enc_block_weights {
  snacb_ru_weights ru[3];
  const float* alpha;
  const float* down_w;
  const float* down_b;
} snacb_enc_block_weights;

typedef struct snacb_encoder_weights {
  const float* in_w;
  const float* in_b;
  snacb_enc_block_weights block[4];
  const float* out_dw_w;
  const float* out_dw_b;
  const float* inproj_w[3];
  const float* inproj_b[3];
} snacb_encoder_weights;

/* Needs snacb_load_weights first (codebooks and out-projections are shared with the decode path). */
int snacb_load_encoder_weights(snacb_engine* e, const snacb_encoder_weights* w);
/* d_audio [batch][n_samples] float32 on the device, n_samples a multiple of 2048 (caller pads with zeros like
 * SNAC.preprocess); d_c0 [batch][n/2048], d_c1 [batch][n/1024], d_c2 [batch][n/512] int32; d_latent (nullable)
 * [batch][n/512][768] receives the encoder output z before quantisation. */
int snacb_encode(snacb_engine* e, const float* d_audio, int32_t batch, int32_t n_samples, int32_t* d_c0, int32_t* d_c1,
                 int32_t* d_c2, float* d_latent, void* stream);

/* Replaces, batched over streams, the per-token Python of speechpipe.py:146-189 (turn_token_into_id: the last
 * "<custom_token_N>" of the stripped string -> N - 10 - 4096 * (accepted_count % 7), None on any parse failure)
 * and the window control flow of tokens_decoder, speechpipe.py:191-293 (ids <= 0 dropped without advancing the
 * slot, first chunk after 7 accepted tokens until one decode returned non-None, then the last 28 / last 49 ids
 * every 7 accepted tokens, end-of-stream flush padded to 28 with the last id).  Streams are slots 0..n-1. */
typedef struct snacb_ingest snacb_ingest;
/* turn_token_into_id(text, index), speechpipe.py:146-189: returns 1 and *id, or 0 where the reference returns None. */
int snacb_parse_token(const char* text, int32_t len, int32_t index, int64_t* id);
int snacb_ingest_create(snacb_ingest** out, int32_t n_streams);
void snacb_ingest_destroy(snacb_ingest* g);
/* A new request in this slot (barge-in, swap): drops everything queued for it. */
int snacb_ingest_reset(snacb_ingest* g, int32_t stream);
/* n token strings (one generator item each): string i is blob[offsets[i] .. offsets[i+1]) for streams[i]. */
int snacb_ingest_push(snacb_ingest* g, int32_t n, const int32_t* streams, const char* blob, const int64_t* offsets);
/* The producer of this stream is exhausted (the flush rule applies once its queue is consumed). */
int snacb_ingest_finish(snacb_ingest* g, int32_t stream);
/* The next ready window of every stream (at most one per stream, slot order, at most max_win): row w of
 * tokens[max_win][tokens_stride >= 49] gets ntok[w] ids (7, 28 or 49; zero-filled beyond), stream_of[w] its slot.
 * Returns the number of windows (>= 0) or a negative status.  The rows are what snacb_decode_windows_host takes. */
int32_t snacb_ingest_tick(snacb_ingest* g, int32_t max_win, int32_t* tokens, int32_t tokens_stride, int32_t* ntok,
                          int32_t* stream_of);
/* Per-window statuses (SNACB_WIN_*) of the windows the last tick returned: latches the first-chunk rule. */
int snacb_ingest_result(snacb_ingest* g, int32_t n, const int32_t* stream_of, const int32_t* status);
/* 1 when the stream is finished, flushed and has nothing queued. */
int snacb_ingest_done(const snacb_ingest* g, int32_t stream);
/* which: 0 accepted tokens, 1 rejected token strings, 2 windows emitted. */
int64_t snacb_ingest_stat(const snacb_ingest* g, int32_t which);

/* ---- N3: PCM egress (host only, no GPU) ----------------------------------------------------- */

/* server.py:50-70 riff_header(): the 44-byte RIFF/WAVE header with unknown (0xFFFFFFFF) lengths, mono PCM16. Returns 44. */
int snacb_riff_header(int32_t sample_rate, uint8_t* out44);
/* orchestrator/stitcher.py:10-79 stitch_chunks(): overlap-add crossfade of consecutive int16 chunks, byte-identical to
 * the numpy code (float64 tail carried between chunks, linspace fades, truncation on emission). */
typedef struct snacb_stitcher snacb_stitcher;
int snacb_stitch_create(snacb_stitcher** out, int32_t sample_rate, double overlap_ms);
void snacb_stitch_destroy(snacb_stitcher* s);
/* Feed one chunk; writes what the reference yields for it.  Returns samples written, or -(needed) if cap is too small
 * (nothing consumed).  *emitted = 1 when a chunk is yielded (an eos chunk is yielded even if empty), *out_eos its flag. */
int64_t snacb_stitch_push(snacb_stitcher* s, const int16_t* pcm, int64_t n, int32_t eos, int16_t* out, int64_t cap,
                          int32_t* emitted, int32_t* out_eos);
/* The chunk source ended without an eos chunk: the kept tail (yielded with eos = True when non-empty). */
int64_t snacb_stitch_flush(snacb_stitcher* s, int16_t* out, int64_t cap);
int64_t snacb_stitch_overlap_samples(const snacb_stitcher* s);
/* One stitcher per stream slot, one call per decode tick: chunk i = pcm[i * pcm_stride .. + len) (a row of the PCM matrix
 * the decode returned) for stream slots[i]; row i of out[n][out_stride >= len + overlap] gets out_len[i] samples
 * (-1 = the reference yields nothing for this chunk), out_eos[i] its eos flag.  eos_in may be NULL (no chunk ends). */
typedef struct snacb_stitch_bank snacb_stitch_bank;
int snacb_stitch_bank_create(snacb_stitch_bank** out, int32_t n_streams, int32_t sample_rate, double overlap_ms);
void snacb_stitch_bank_destroy(snacb_stitch_bank* b);
int snacb_stitch_bank_reset(snacb_stitch_bank* b, int32_t stream);
int snacb_stitch_bank_push(snacb_stitch_bank* b, int32_t n, const int32_t* slots, const int16_t* pcm, int64_t pcm_stride,
                           int64_t len, const int32_t* eos_in, int16_t* out, int64_t out_stride, int64_t* out_len,
                           int32_t* out_eos);

/* ---- N3 on the GPU: decode ticks written straight into pinned per-stream PCM rings ------------
 * Replaces, for every stream of a tick at once, orchestrator/stitcher.py:10-79 (crossfade of consecutive chunks, same
 * float64 arithmetic, done by a kernel right after the decoder tail), orchestrator/ring_buffer.py:27-83 (the byte ring)
 * and the pull(chunk_size) re-chunking of tts_engine/llama_local.py:120-150 (snacb_egress_read).  A ring is
 * `ring_samples` int16 samples (multiple of 8, >= 4096) of pinned host memory per slot that the GPU writes through its
 * device mapping; the PCM matrix of the tick never crosses PCIe as a separate copy.
 * Threading: pushes / sync / flush belong to ONE producer thread and stream at a time (the tick loop); reads of a slot may
 * come from another thread (the consumer of that stream) once the tick that wrote the samples has been synchronised; a
 * slot is reset or re-used only while no push that names it is in flight. */
typedef struct snacb_egress snacb_egress;
int snacb_egress_create(snacb_egress** out, int32_t device, int32_t n_slots, int32_t ring_samples, int32_t sample_rate,
                        double overlap_ms);
void snacb_egress_destroy(snacb_egress* g);
const char* snacb_egress_last_error(const snacb_egress* g);
int64_t snacb_egress_overlap_samples(const snacb_egress* g);
/* Host address of a slot's ring (for zero-copy consumers). */
const int16_t* snacb_egress_ring_base(const snacb_egress* g, int32_t slot);
/* Asynchronous on `stream`: chunk i = d_pcm[i * pcm_stride .. + len) (device memory) joins the ring of slot h_slots[i]
 * (-1 = not for a ring); windows whose d_status[i] != SNACB_WIN_OK contribute nothing (d_status may be NULL); h_eos[i] != 0
 * closes the stream (tail emitted, no fade kept), may be NULL.  A slot may appear once per call.  SNACB_ESTATE if a ring
 * could overflow: nothing is launched, read first. */
int snacb_egress_push_device(snacb_egress* g, int32_t n_win, const int32_t* h_slots, const int16_t* d_pcm, int64_t pcm_stride,
                             int32_t len, const int32_t* d_status, const int32_t* h_eos, void* stream);
/* Synchronises `stream`; afterwards available / read see everything pushed on it. */
int snacb_egress_sync(snacb_egress* g, void* stream);
int64_t snacb_egress_available(const snacb_egress* g, int32_t slot);
/* Samples a push may still add to the slot before SNACB_ESTATE (ring full). */
int64_t snacb_egress_room(const snacb_egress* g, int32_t slot);
/* Up to max_samples unread samples of the slot into dst; returns the count (0 = nothing yet). */
int64_t snacb_egress_read(snacb_egress* g, int32_t slot, int16_t* dst, int64_t max_samples);
/* The stream ended without an eos chunk: the kept tail is emitted (synchronous). */
int snacb_egress_flush(snacb_egress* g, int32_t slot, void* stream);
/* Barge-in / slot reuse: drop unread samples and the kept tail (host-only and immediate, the device state follows with
 * the slot's next push; no push of this slot may be in flight). */
int snacb_egress_reset(snacb_egress* g, int32_t slot, void* stream);
/* Zero-call consumers: host addresses of the write cursors (pinned int64 [n_slots], published by the kernels) and the read
 * cursors (int64 [n_slots]); a consumer copies samples [rpos, min(wpos, rpos + n)) out of snacb_egress_ring_base(slot)
 * (modulo ring_samples) and advances rpos[slot] itself - what snacb_egress_read does. */
int snacb_egress_cursors(snacb_egress* g, const int64_t** wpos, int64_t** rpos);
/* Samples written to the slot since its last reset (as of the last snacb_egress_sync). */
int64_t snacb_egress_written(const snacb_egress* g, int32_t slot);
/* One tick from host token buffers into the rings: snacb_decode_windows_host with the PCM going to slot h_slots[i] of
 * `g` instead of a host matrix (H2D tokens -> kernels -> ring kernel -> D2H statuses, synchronous).  h_emitted (may be
 * NULL) receives the samples window i added to its ring, or -1 when that ring had no room for a window: such a window
 * is decoded to nowhere so that one stalled consumer cannot fail the tick of everybody else. */
int snacb_decode_windows_to_ring(snacb_engine* e, snacb_egress* g, const int32_t* h_tokens, int32_t tokens_stride,
                                 const int32_t* h_ntok, int32_t ntok_uniform, int32_t n_win, int32_t noise_mode, uint64_t seed,
                                 const uint64_t* h_keys, const int32_t* h_slots, const int32_t* h_eos, int32_t* h_status,
                                 int32_t* h_emitted, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SNACB_H */
