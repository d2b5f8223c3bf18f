// PCM egress: what happens to the int16 chunks right after the decode path, host side.
//
//   snacb_riff_header   Morpheus_Client/server.py:50-70        the 44-byte streaming RIFF/WAVE header (unknown length)
//   snacb_stitch_*      Morpheus_Client/orchestrator/stitcher.py:10-79   overlap-add crossfade of consecutive chunks
//
// The stitcher is restated operation for operation so the bytes are identical to numpy's: the kept tail stays in
// float64 between chunks (the reference only truncates what it emits), the fades are numpy.linspace(.., endpoint=False)
// = i * step + start with both roundings, fade_out + fade_in is two rounded products and one rounded sum, and emission
// is astype('<i2') = truncation toward zero.  No GPU code here.
#include <cstdint>
#include <cstring>
#include <new>
#include <vector>

#include "snacb.h"

struct snacb_stitcher {
  int32_t sample_rate = 24000;
  int64_t overlap = 0;          // int(overlap_ms * sample_rate / 1000.0)
  std::vector<double> tail;     // kept samples, not yet truncated
  bool done = false;            // an eos chunk was emitted
  std::vector<double> work;
};

struct snacb_stitch_bank {
  std::vector<snacb_stitcher> s;
};

namespace {
inline void put_le16(uint8_t* p, uint16_t v) { p[0] = (uint8_t)(v & 0xff); p[1] = (uint8_t)(v >> 8); }
inline void put_le32(uint8_t* p, uint32_t v) { for (int i = 0; i < 4; ++i) p[i] = (uint8_t)((v >> (8 * i)) & 0xff); }
// numpy.linspace(start, start + delta, num, endpoint=False)[i]
inline double linspace_at(double start, double delta, int64_t num, int64_t i) {
  volatile double step = delta / (double)num;  // volatile: each operation rounds on its own (no contraction)
  volatile double y = (double)i * step;
  return y + start;
}
inline int16_t trunc_i16(double v) { return (int16_t)(int32_t)v; }
}  // namespace

extern "C" {

int snacb_riff_header(int32_t sample_rate, uint8_t* out44) {
  if (!out44 || sample_rate <= 0) return SNACB_EINVAL;
  memcpy(out44, "RIFF", 4);
  put_le32(out44 + 4, 0xFFFFFFFFu);
  memcpy(out44 + 8, "WAVE", 4);
  memcpy(out44 + 12, "fmt ", 4);
  put_le32(out44 + 16, 16);
  put_le16(out44 + 20, 1);                                  // PCM
  put_le16(out44 + 22, 1);                                  // mono
  put_le32(out44 + 24, (uint32_t)sample_rate);
  put_le32(out44 + 28, (uint32_t)sample_rate * 2u);         // byte rate
  put_le16(out44 + 32, 2);                                  // block align
  put_le16(out44 + 34, 16);                                 // bits per sample
  memcpy(out44 + 36, "data", 4);
  put_le32(out44 + 40, 0xFFFFFFFFu);
  return 44;
}

int snacb_stitch_create(snacb_stitcher** out, int32_t sample_rate, double overlap_ms) {
  if (!out || sample_rate <= 0) return SNACB_EINVAL;
  snacb_stitcher* s = new (std::nothrow) snacb_stitcher();
  if (!s) return SNACB_ENOMEM;
  s->sample_rate = sample_rate;
  const double ov = overlap_ms * (double)sample_rate / 1000.0;
  s->overlap = (int64_t)ov;  // Python int(): toward zero
  *out = s;
  return SNACB_OK;
}

void snacb_stitch_destroy(snacb_stitcher* s) { delete s; }

// One chunk in, the samples the reference yields for it out.  Returns the number of samples written (0 = the
// reference yields nothing for this chunk), or -(needed) when cap is too small (state untouched).  *emitted is 1
// when a chunk is yielded (an eos chunk is yielded even when empty), *out_eos its eos flag.
int64_t snacb_stitch_push(snacb_stitcher* s, const int16_t* pcm, int64_t n, int32_t eos, int16_t* out, int64_t cap,
                          int32_t* emitted, int32_t* out_eos) {
  if (!s || n < 0 || (n > 0 && !pcm) || !emitted || !out_eos || cap < 0 || (cap > 0 && !out)) return SNACB_EINVAL;
  *emitted = 0;
  *out_eos = 0;
  if (s->done) return 0;  // the reference's loop has exited after the eos chunk
  const int64_t tn = (int64_t)s->tail.size();
  const int64_t ov = (tn && s->overlap > 0) ? (s->overlap < tn ? (s->overlap < n ? s->overlap : n) : (tn < n ? tn : n)) : 0;
  const int64_t total = tn + n - ov;  // len(pcm) after the join
  const int64_t keep = eos ? 0 : (s->overlap > 0 ? (total <= s->overlap ? total : s->overlap) : 0);
  const int64_t n_out = total - keep;
  if (n_out > cap) return -n_out;
  // conceptual joined array w = [tail[0 : tn-ov] | crossfade[0 : ov] | pcm[ov : n]]; out = w[0 : n_out], new tail = w[n_out :]
  const int64_t a = tn - ov;
  auto val = [&](int64_t i) -> double {
    if (i < a) return s->tail[(size_t)i];
    if (i < tn) {
      const int64_t j = i - a;
      volatile double fo = s->tail[(size_t)i] * linspace_at(1.0, -1.0, ov, j);
      volatile double fi = (double)pcm[j] * linspace_at(0.0, 1.0, ov, j);
      return fo + fi;
    }
    return (double)pcm[i - tn + ov];
  };
  const int64_t head = n_out < tn ? n_out : tn;
  for (int64_t i = 0; i < head; ++i) out[i] = trunc_i16(val(i));
  if (n_out > tn) memcpy(out + tn, pcm + ov, (size_t)(n_out - tn) * sizeof(int16_t));  // untouched samples: straight copy
  std::vector<double>& w = s->work;
  w.resize((size_t)keep);
  for (int64_t k = 0; k < keep; ++k) w[(size_t)k] = val(n_out + k);
  s->tail.swap(w);
  if (eos) {
    s->done = true;
    *emitted = 1;
    *out_eos = 1;
  } else if (s->overlap > 0 && total <= s->overlap) {
    *emitted = 0;  // not enough to emit: everything went into the tail
  } else {
    *emitted = 1;
  }
  return n_out;
}

// The chunk iterator ended without an eos chunk: the remaining tail, yielded with eos = True.
int64_t snacb_stitch_flush(snacb_stitcher* s, int16_t* out, int64_t cap) {
  if (!s || cap < 0 || (cap > 0 && !out)) return SNACB_EINVAL;
  if (s->done) return 0;
  const int64_t tn = (int64_t)s->tail.size();
  if (tn > cap) return -tn;
  for (int64_t i = 0; i < tn; ++i) out[i] = trunc_i16(s->tail[(size_t)i]);
  s->tail.clear();
  s->done = true;
  return tn;
}

int64_t snacb_stitch_overlap_samples(const snacb_stitcher* s) { return s ? s->overlap : -1; }

// ---- a bank of stitchers: one call per decode tick for all streams (rows of the PCM matrix the decode returned)
int snacb_stitch_bank_create(snacb_stitch_bank** out, int32_t n_streams, int32_t sample_rate, double overlap_ms) {
  if (!out || n_streams < 0 || sample_rate <= 0) return SNACB_EINVAL;
  snacb_stitch_bank* b = new (std::nothrow) snacb_stitch_bank();
  if (!b) return SNACB_ENOMEM;
  b->s.resize((size_t)n_streams);
  for (auto& st : b->s) {
    st.sample_rate = sample_rate;
    st.overlap = (int64_t)(overlap_ms * (double)sample_rate / 1000.0);
  }
  *out = b;
  return SNACB_OK;
}

void snacb_stitch_bank_destroy(snacb_stitch_bank* b) { delete b; }

int snacb_stitch_bank_reset(snacb_stitch_bank* b, int32_t stream) {
  if (!b || stream < 0 || (size_t)stream >= b->s.size()) return SNACB_EINVAL;
  snacb_stitcher& st = b->s[(size_t)stream];
  st.tail.clear();
  st.done = false;
  return SNACB_OK;
}

int snacb_stitch_bank_push(snacb_stitch_bank* b, int32_t n, const int32_t* slots, const int16_t* pcm, int64_t pcm_stride,
                           int64_t len, const int32_t* eos_in, int16_t* out, int64_t out_stride, int64_t* out_len,
                           int32_t* out_eos) {
  if (!b || n < 0 || len < 0 || (n > 0 && (!slots || !pcm || !out || !out_len || !out_eos))) return SNACB_EINVAL;
  for (int32_t i = 0; i < n; ++i) {
    if (slots[i] < 0 || (size_t)slots[i] >= b->s.size()) return SNACB_EINVAL;
    if (out_stride < len + b->s[(size_t)slots[i]].overlap) return SNACB_EINVAL;
  }
  for (int32_t i = 0; i < n; ++i) {
    int32_t emitted = 0, eos = 0;
    const int64_t w = snacb_stitch_push(&b->s[(size_t)slots[i]], pcm + (size_t)i * pcm_stride, len, eos_in ? eos_in[i] : 0,
                                        out + (size_t)i * out_stride, out_stride, &emitted, &eos);
    if (w < 0) return SNACB_EINVAL;
    out_len[i] = emitted ? w : -1;
    out_eos[i] = eos;
  }
  return SNACB_OK;
}

}  // extern "C"
