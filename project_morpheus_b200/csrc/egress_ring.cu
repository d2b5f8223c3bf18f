// PCM egress, GPU side: the decode tick writes every stream's samples straight into a pinned,
// host-visible per-stream ring, crossfading consecutive chunks on the way when an overlap is configured.
//
//   k_stitch_ring    Morpheus_Client/orchestrator/stitcher.py:10-79   overlap-add of consecutive chunks of one stream
//                    Morpheus_Client/orchestrator/ring_buffer.py:27-83 the byte ring the consumer reads from
//                    Morpheus_Client/tts_engine/llama_local.py:120-150 pull(chunk_size) re-chunking (served by _read)
//
// One CTA per window of the tick.  The chunk (2048 int16 samples in HBM, written by the decoder tail) is joined to the
// slot's kept tail with the reference's float64 arithmetic restated operation for operation (same as csrc/egress.cpp:
// fades = numpy.linspace(.., endpoint=False) = i*step + start with both roundings, two rounded products and one rounded
// sum, truncation on emission; __dmul_rn / __dadd_rn / __ddiv_rn keep the compiler from contracting them), and the
// emitted samples go to the ring through its device mapping in 16-byte stores; the write cursor of the slot is published
// to pinned memory after a system-scope fence.  Windows whose status is not SNACB_WIN_OK emit nothing (the reference
// yields nothing for them).  The host side only moves a read cursor and memcpy's out of pinned memory.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "snacb.h"

namespace {

constexpr int kMaxOverlap = 4096;  // samples of kept tail (170 ms at 24 kHz): 32 KB of shared memory per CTA

struct SlotState {          // device side, one per slot
  long long wpos;           // samples ever written to the slot's ring
  int tail_len;             // kept samples (not yet truncated), <= overlap
  int done;                 // an eos chunk was emitted
};

__device__ __forceinline__ double linspace_at(double start, double delta, int num, int i) {
  const double step = __ddiv_rn(delta, (double)num);
  return __dadd_rn(__dmul_rn((double)i, step), start);
}
__device__ __forceinline__ short trunc_i16(double v) { return (short)(int)v; }

// slots[i] < 0: window i does not go to a ring.  pcm rows of `len` samples at `pcm_stride`.
__global__ void __launch_bounds__(256) k_stitch_ring(const int* __restrict__ slots, const int* __restrict__ status,
                                                     const int* __restrict__ eos_in, const short* __restrict__ pcm,
                                                     long long pcm_stride, int len, int overlap, int ring_samples,
                                                     SlotState* __restrict__ state, double* __restrict__ tails,
                                                     short* __restrict__ ring, long long* __restrict__ wpos_pub) {
  extern __shared__ double s_tail[];
  const int w = blockIdx.x, s = slots[w];
  if (s < 0) return;
  SlotState stt = state[s];
  const int flags = eos_in ? eos_in[w] : 0;  // bit 0: this chunk ends the stream, bit 1: the slot was reset since its last push
  if (flags & 2) stt = SlotState{0, 0, 0};
  const int eos = flags & 1;
  const int n = (status == nullptr || status[w] == SNACB_WIN_OK) ? len : 0;
  if (stt.done || (n == 0 && !eos)) {
    if ((flags & 2) && threadIdx.x == 0) { state[s] = stt; wpos_pub[s] = 0; }
    return;
  }
  const int tn = stt.tail_len;
  double* g_tail = tails + (size_t)s * overlap;
  for (int i = threadIdx.x; i < tn; i += blockDim.x) s_tail[i] = g_tail[i];
  __syncthreads();
  const int ov = (tn && overlap > 0) ? min(min(overlap, tn), n) : 0;
  const int total = tn + n - ov;
  const int keep = eos ? 0 : (overlap > 0 ? min(total, overlap) : 0);
  const int n_out = total - keep;
  const int a = tn - ov;
  const short* src = pcm + (size_t)w * pcm_stride;
  auto val = [&](int i) -> double {   // joined array [tail[0:a] | crossfade[0:ov] | pcm[ov:n]]
    if (i < a) return s_tail[i];
    if (i < tn) {
      const int j = i - a;
      const double fo = __dmul_rn(s_tail[i], linspace_at(1.0, -1.0, ov, j));
      const double fi = __dmul_rn((double)src[j], linspace_at(0.0, 1.0, ov, j));
      return __dadd_rn(fo, fi);
    }
    return (double)src[i - tn + ov];
  };
  auto sample = [&](int i) -> short { return i < tn ? trunc_i16(val(i)) : src[i - tn + ov]; };
  // emission: destination groups of 8 samples (16 bytes); ring_samples is a multiple of 8, so a group never wraps
  short* rbase = ring + (size_t)s * ring_samples;
  const long long w0 = stt.wpos;
  const long long g0 = w0 >> 3, g1 = (w0 + n_out + 7) >> 3;
  for (long long g = g0 + threadIdx.x; g < g1; g += blockDim.x) {
    const long long p0 = g << 3;
    const long long lo = max(p0, w0), hi = min(p0 + 8, w0 + (long long)n_out);
    short* dst = rbase + (size_t)(p0 % ring_samples);
    if (hi - lo == 8) {
      union { uint4 v; short h[8]; } u;
#pragma unroll
      for (int k = 0; k < 8; ++k) u.h[k] = sample((int)(p0 - w0) + k);
      *reinterpret_cast<uint4*>(dst) = u.v;
    } else {
      for (long long p = lo; p < hi; ++p) dst[p - p0] = sample((int)(p - w0));
    }
  }
  // new tail (reads the old one from shared memory only)
  for (int k = threadIdx.x; k < keep; k += blockDim.x) g_tail[k] = val(n_out + k);
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    stt.wpos = w0 + n_out;
    stt.tail_len = keep;
    stt.done = eos ? 1 : 0;
    state[s] = stt;
    wpos_pub[s] = stt.wpos;
    __threadfence_system();
  }
}

// End of a stream without an eos chunk: the kept tail is emitted (truncated) and the slot is closed.
__global__ void k_ring_flush(int s, int overlap, int ring_samples, SlotState* state, const double* tails, short* ring,
                             long long* wpos_pub) {
  SlotState stt = state[s];
  if (stt.done) return;
  const double* g_tail = tails + (size_t)s * overlap;
  short* rbase = ring + (size_t)s * ring_samples;
  for (int i = threadIdx.x; i < stt.tail_len; i += blockDim.x) rbase[(stt.wpos + i) % ring_samples] = trunc_i16(g_tail[i]);
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    stt.wpos += stt.tail_len;
    stt.tail_len = 0;
    stt.done = 1;
    state[s] = stt;
    wpos_pub[s] = stt.wpos;
    __threadfence_system();
  }
}

}  // namespace

struct snacb_egress {
  int device = 0;
  int n_slots = 0, ring_samples = 0, overlap = 0, sample_rate = 24000;
  short* h_ring = nullptr;        // pinned + mapped [n_slots][ring_samples]
  short* d_ring = nullptr;        // its device alias
  long long* h_wpos = nullptr;    // pinned + mapped [n_slots], published by the kernels
  long long* d_wpos = nullptr;
  SlotState* d_state = nullptr;
  double* d_tails = nullptr;
  int* d_args = nullptr;          // [3][cap] slots | eos, staged through h_args
  int* h_args = nullptr;          // pinned
  int args_cap = 0;
  cudaEvent_t args_free = nullptr;  // the previous push has consumed h_args
  std::vector<long long> rpos;    // host read cursors (snacb_egress_cursors hands their address to zero-call consumers)
  std::vector<long long> wbound;  // upper bound of the device write cursor (pushes in flight included)
  std::vector<int> seen;          // duplicate-slot check, tick stamp
  std::vector<char> lazy_reset;   // the slot was reset on the host; the device state follows with its next push
  int stamp = 0;
  std::string err;
};

namespace {
int efail(snacb_egress* g, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (g) g->err = buf;
  return code;
}
#define CUE(g, x)                                                                                        \
  do {                                                                                                   \
    cudaError_t _e = (x);                                                                                \
    if (_e != cudaSuccess) return efail(g, SNACB_ECUDA, "%s: %s", #x, cudaGetErrorString(_e));          \
  } while (0)
}  // namespace

extern "C" {

int snacb_egress_create(snacb_egress** out, int32_t device, int32_t n_slots, int32_t ring_samples, int32_t sample_rate,
                        double overlap_ms) {
  if (!out || n_slots <= 0 || ring_samples < 4096 || (ring_samples & 7) || sample_rate <= 0 || overlap_ms < 0) return SNACB_EINVAL;
  const long long ov = (long long)(overlap_ms * (double)sample_rate / 1000.0);  // Python int(): toward zero
  if (ov > kMaxOverlap || 2 * ov + 2048 > ring_samples) return SNACB_EINVAL;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); return SNACB_ECUDA; }
  snacb_egress* g = new (std::nothrow) snacb_egress();
  if (!g) return SNACB_ENOMEM;
  g->device = device; g->n_slots = n_slots; g->ring_samples = ring_samples; g->overlap = (int)ov; g->sample_rate = sample_rate;
  auto bail = [&](int code) { snacb_egress_destroy(g); return code; };
  if (cudaSetDevice(device) != cudaSuccess) return bail(SNACB_ECUDA);
  const size_t ring_b = (size_t)n_slots * ring_samples * sizeof(short);
  if (cudaHostAlloc((void**)&g->h_ring, ring_b, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) return bail(SNACB_ENOMEM);
  if (cudaHostGetDevicePointer((void**)&g->d_ring, g->h_ring, 0) != cudaSuccess) return bail(SNACB_ECUDA);
  if (cudaHostAlloc((void**)&g->h_wpos, (size_t)n_slots * 8, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) return bail(SNACB_ENOMEM);
  if (cudaHostGetDevicePointer((void**)&g->d_wpos, g->h_wpos, 0) != cudaSuccess) return bail(SNACB_ECUDA);
  memset(g->h_wpos, 0, (size_t)n_slots * 8);
  if (cudaMalloc((void**)&g->d_state, (size_t)n_slots * sizeof(SlotState)) != cudaSuccess) return bail(SNACB_ENOMEM);
  if (cudaMemset(g->d_state, 0, (size_t)n_slots * sizeof(SlotState)) != cudaSuccess) return bail(SNACB_ECUDA);
  if (cudaMalloc((void**)&g->d_tails, (size_t)n_slots * (ov > 0 ? ov : 1) * sizeof(double)) != cudaSuccess) return bail(SNACB_ENOMEM);
  if (cudaEventCreateWithFlags(&g->args_free, cudaEventDisableTiming) != cudaSuccess) return bail(SNACB_ECUDA);
  if (cudaDeviceSynchronize() != cudaSuccess) return bail(SNACB_ECUDA);
  g->rpos.assign((size_t)n_slots, 0);
  g->wbound.assign((size_t)n_slots, 0);
  g->seen.assign((size_t)n_slots, 0);
  g->lazy_reset.assign((size_t)n_slots, 0);
  *out = g;
  return SNACB_OK;
}

void snacb_egress_destroy(snacb_egress* g) {
  if (!g) return;
  cudaSetDevice(g->device);
  cudaDeviceSynchronize();
  if (g->h_ring) cudaFreeHost(g->h_ring);
  if (g->h_wpos) cudaFreeHost(g->h_wpos);
  if (g->h_args) cudaFreeHost(g->h_args);
  if (g->d_state) cudaFree(g->d_state);
  if (g->d_tails) cudaFree(g->d_tails);
  if (g->d_args) cudaFree(g->d_args);
  if (g->args_free) cudaEventDestroy(g->args_free);
  cudaGetLastError();
  delete g;
}

const char* snacb_egress_last_error(const snacb_egress* g) { return g ? g->err.c_str() : "null egress"; }
int64_t snacb_egress_overlap_samples(const snacb_egress* g) { return g ? g->overlap : -1; }
const int16_t* snacb_egress_ring_base(const snacb_egress* g, int32_t slot) {
  return (g && slot >= 0 && slot < g->n_slots) ? g->h_ring + (size_t)slot * g->ring_samples : nullptr;
}

// Asynchronous on `stream`: window i of the tick (device PCM row i, `len` samples, device status i or NULL = all OK) is
// appended to the ring of slot h_slots[i] (-1 = skip).  A slot may appear once per call.  SNACB_ESTATE when a ring could
// overflow (the consumer must read first); nothing is launched in that case.
int snacb_egress_push_device(snacb_egress* g, int32_t n_win, const int32_t* h_slots, const int16_t* d_pcm, int64_t pcm_stride,
                             int32_t len, const int32_t* d_status, const int32_t* h_eos, void* stream) {
  if (!g) return SNACB_EINVAL;
  if (n_win < 0 || len < 0 || (n_win > 0 && (!h_slots || !d_pcm)) || pcm_stride < len)
    return efail(g, SNACB_EINVAL, "snacb_egress_push_device: bad argument");
  if (n_win == 0) return SNACB_OK;
  if (len + 2 * g->overlap > g->ring_samples) return efail(g, SNACB_EINVAL, "snacb_egress_push_device: chunk longer than the ring");
  cudaStream_t st = (cudaStream_t)stream;
  CUE(g, cudaSetDevice(g->device));
  ++g->stamp;
  for (int i = 0; i < n_win; ++i) {
    const int s = h_slots[i];
    if (s < 0) continue;
    if (s >= g->n_slots) return efail(g, SNACB_EINVAL, "snacb_egress_push_device: slot %d out of range", s);
    if (g->seen[(size_t)s] == g->stamp) return efail(g, SNACB_EINVAL, "snacb_egress_push_device: slot %d twice in one tick", s);
    g->seen[(size_t)s] = g->stamp;
    if (g->wbound[(size_t)s] + len + g->overlap - g->rpos[(size_t)s] > g->ring_samples)
      return efail(g, SNACB_ESTATE, "snacb_egress_push_device: ring of slot %d is full (read it first)", s);
  }
  if (n_win > g->args_cap) {
    CUE(g, cudaStreamSynchronize(st));
    CUE(g, cudaEventSynchronize(g->args_free));
    if (g->h_args) CUE(g, cudaFreeHost(g->h_args));
    if (g->d_args) CUE(g, cudaFree(g->d_args));
    g->h_args = nullptr; g->d_args = nullptr;
    g->args_cap = n_win + n_win / 2 + 64;
    CUE(g, cudaMallocHost((void**)&g->h_args, (size_t)g->args_cap * 2 * sizeof(int)));
    CUE(g, cudaMalloc((void**)&g->d_args, (size_t)g->args_cap * 2 * sizeof(int)));
  }
  CUE(g, cudaEventSynchronize(g->args_free));  // the previous push's H2D copy has read h_args
  memcpy(g->h_args, h_slots, (size_t)n_win * sizeof(int));
  int* flags = g->h_args + g->args_cap;
  for (int i = 0; i < n_win; ++i) {
    const int s = h_slots[i];
    flags[i] = (h_eos && h_eos[i]) ? 1 : 0;
    if (s >= 0 && g->lazy_reset[(size_t)s]) { flags[i] |= 2; g->lazy_reset[(size_t)s] = 0; }
  }
  CUE(g, cudaMemcpyAsync(g->d_args, g->h_args, (size_t)n_win * sizeof(int), cudaMemcpyHostToDevice, st));
  CUE(g, cudaMemcpyAsync(g->d_args + g->args_cap, flags, (size_t)n_win * sizeof(int), cudaMemcpyHostToDevice, st));
  CUE(g, cudaEventRecord(g->args_free, st));
  const size_t smem = (size_t)(g->overlap > 0 ? g->overlap : 1) * sizeof(double);
  k_stitch_ring<<<n_win, 256, smem, st>>>(g->d_args, d_status, g->d_args + g->args_cap,
                                          reinterpret_cast<const short*>(d_pcm), (long long)pcm_stride, len, g->overlap,
                                          g->ring_samples, g->d_state, g->d_tails, g->d_ring, g->d_wpos);
  CUE(g, cudaGetLastError());
  for (int i = 0; i < n_win; ++i)
    if (h_slots[i] >= 0) g->wbound[(size_t)h_slots[i]] += len + g->overlap;
  return SNACB_OK;
}

// After the stream that carried the pushes has been synchronised: the published cursors are exact again.
int snacb_egress_sync(snacb_egress* g, void* stream) {
  if (!g) return SNACB_EINVAL;
  CUE(g, cudaSetDevice(g->device));
  CUE(g, cudaStreamSynchronize((cudaStream_t)stream));
  for (int s = 0; s < g->n_slots; ++s) g->wbound[(size_t)s] = g->h_wpos[s];
  return SNACB_OK;
}

// Samples that can still be pushed to the slot before its ring would overflow (pushes in flight counted in full).
int64_t snacb_egress_room(const snacb_egress* g, int32_t slot) {
  if (!g || slot < 0 || slot >= g->n_slots) return SNACB_EINVAL;
  const long long room = (long long)g->ring_samples - (g->wbound[(size_t)slot] - g->rpos[(size_t)slot]) - g->overlap;
  return room > 0 ? room : 0;
}

int64_t snacb_egress_available(const snacb_egress* g, int32_t slot) {
  if (!g || slot < 0 || slot >= g->n_slots) return SNACB_EINVAL;
  return *(volatile long long*)&g->h_wpos[slot] - g->rpos[(size_t)slot];
}

// Up to max_samples of the slot's unread samples into dst (plain host memory); returns the count.
int64_t snacb_egress_read(snacb_egress* g, int32_t slot, int16_t* dst, int64_t max_samples) {
  if (!g || slot < 0 || slot >= g->n_slots || max_samples < 0 || (max_samples > 0 && !dst)) return SNACB_EINVAL;
  const long long w = *(volatile long long*)&g->h_wpos[slot];
  long long r = g->rpos[(size_t)slot];
  const long long n = (w - r) < max_samples ? (w - r) : max_samples;
  const short* base = g->h_ring + (size_t)slot * g->ring_samples;
  long long left = n;
  int16_t* o = dst;
  while (left > 0) {
    const long long off = r % g->ring_samples;
    const long long run = (g->ring_samples - off) < left ? (g->ring_samples - off) : left;
    memcpy(o, base + off, (size_t)run * sizeof(short));
    o += run; r += run; left -= run;
  }
  g->rpos[(size_t)slot] = r;
  return n;
}

int snacb_egress_flush(snacb_egress* g, int32_t slot, void* stream) {
  if (!g || slot < 0 || slot >= g->n_slots) return SNACB_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  CUE(g, cudaSetDevice(g->device));
  if (g->lazy_reset[(size_t)slot]) return SNACB_OK;  // reset and never pushed since: nothing is kept
  if (g->wbound[(size_t)slot] + g->overlap - g->rpos[(size_t)slot] > g->ring_samples)
    return efail(g, SNACB_ESTATE, "snacb_egress_flush: ring of slot %d is full (read it first)", slot);
  k_ring_flush<<<1, 256, 0, st>>>(slot, g->overlap, g->ring_samples, g->d_state, g->d_tails, g->d_ring, g->d_wpos);
  CUE(g, cudaGetLastError());
  CUE(g, cudaStreamSynchronize(st));
  g->wbound[(size_t)slot] = g->h_wpos[slot];
  return SNACB_OK;
}

// Barge-in / slot reuse: unread samples and the kept tail are dropped, cursors restart at zero.  Host-only and
// immediate: the device-side state of the slot is cleared by the kernel of its next push (no push of this slot may be in
// flight, which the one-window-per-stream-per-tick discipline guarantees).
int snacb_egress_reset(snacb_egress* g, int32_t slot, void* stream) {
  (void)stream;
  if (!g || slot < 0 || slot >= g->n_slots) return SNACB_EINVAL;
  g->lazy_reset[(size_t)slot] = 1;
  g->h_wpos[slot] = 0;
  g->rpos[(size_t)slot] = 0;
  g->wbound[(size_t)slot] = 0;
  return SNACB_OK;
}

// For consumers that read the rings without a call per read (the Python adapter): host addresses of the published write
// cursors (pinned, int64 [n_slots], samples written since the slot's reset) and of the read cursors (int64 [n_slots]: the
// consumer advances rpos[slot] itself after copying samples [rpos, rpos + n) out of snacb_egress_ring_base(slot); the
// room check of the pushes reads them).  Valid for the lifetime of the object.
int snacb_egress_cursors(snacb_egress* g, const int64_t** wpos, int64_t** rpos) {
  if (!g || !wpos || !rpos) return SNACB_EINVAL;
  static_assert(sizeof(long long) == sizeof(int64_t), "cursor width");
  *wpos = reinterpret_cast<const int64_t*>(g->h_wpos);
  *rpos = reinterpret_cast<int64_t*>(g->rpos.data());
  return SNACB_OK;
}

// Samples ever written to the slot since its last reset (the published device cursor).
int64_t snacb_egress_written(const snacb_egress* g, int32_t slot) {
  if (!g || slot < 0 || slot >= g->n_slots) return SNACB_EINVAL;
  return *(volatile long long*)&g->h_wpos[slot];
}

}  // extern "C"
