// Token ingress for many streams: the host-side control flow in front of the decode path.
//
// Restates, per stream and in C++, what the reference does one Python call per token:
//   turn_token_into_id   Morpheus_Client/tts_engine/speechpipe.py:146-189  ("<custom_token_N>" -> N - 10 - 4096*(count%7),
//                        last occurrence in the stripped string, must end with '>', int() failures -> None)
//   tokens_decoder       Morpheus_Client/tts_engine/speechpipe.py:191-293  (accept ids > 0, count them, first chunk after
//                        7 tokens until one decode succeeded, then every 7 accepted tokens the last 28 / last 49,
//                        end-of-stream flush with padding)
// and batches it: token strings of any number of streams go in as one blob, the next ready window of every stream
// comes out as rows of an int32 matrix that snacb_decode_windows_host takes as is.  No GPU code here.
//
// Same state machine as project_morpheus_b200/tokens.py (WindowPlanner) + scheduler.py (TickScheduler); the parity
// tests drive both with the same strings.  Divergences from Python's int(): digits, '+', '-', '_' between digits and
// ASCII whitespace only (Python also accepts non-ASCII digits / spaces), numbers are clamped to +-2^62.
#include <cstdint>
#include <cstring>
#include <deque>
#include <new>
#include <string>
#include <vector>

#include "snacb.h"

namespace {

constexpr int kTokensPerFrame = 7;
constexpr int kFirstWindow = 7, kShortWindow = 28, kLongWindow = 49;
constexpr char kPrefix[] = "<custom_token_";
constexpr int kPrefixLen = 14;

// str.strip() whitespace (ASCII subset): includes the separators 0x1c-0x1f
inline bool py_space(unsigned char c) { return c == ' ' || (c >= 0x09 && c <= 0x0d) || (c >= 0x1c && c <= 0x1f); }
// whitespace int() skips around the number: ' ', \t \n \v \f \r only - int("\x1c7") raises ValueError
// (speechpipe.py:179-181 turns that into None)
inline bool int_space(unsigned char c) { return c == ' ' || (c >= 0x09 && c <= 0x0d); }

// Python int(str) for the ASCII subset.  Returns false where int() raises ValueError.
bool py_int(const char* p, const char* e, long long* out) {
  while (p < e && int_space((unsigned char)*p)) ++p;
  while (e > p && int_space((unsigned char)e[-1])) --e;
  if (p >= e) return false;
  bool neg = false;
  if (*p == '+' || *p == '-') { neg = (*p == '-'); ++p; }
  if (p >= e || *p < '0' || *p > '9') return false;  // a digit must follow the sign ("_1", "" are invalid)
  long long v = 0;
  bool prev_us = false;
  constexpr long long kClamp = 1LL << 62;
  for (; p < e; ++p) {
    const char c = *p;
    if (c == '_') {
      if (prev_us) return false;
      prev_us = true;
      continue;
    }
    if (c < '0' || c > '9') return false;
    prev_us = false;
    v = (v <= (kClamp - 9) / 10) ? v * 10 + (c - '0') : kClamp;
  }
  if (prev_us) return false;  // trailing underscore
  *out = neg ? -v : v;
  return true;
}

// turn_token_into_id: false = None.
bool token_id(const char* s, int len, int count, long long* id) {
  const char* b = s;
  const char* e = s + len;
  // `"<custom_token_" not in token_string` is tested on the unstripped string; the prefix has no whitespace, so
  // searching the stripped text is equivalent.
  while (b < e && py_space((unsigned char)*b)) ++b;
  while (e > b && py_space((unsigned char)e[-1])) --e;
  if (e - b < kPrefixLen + 1 || e[-1] != '>') return false;
  const char* start = nullptr;
  for (const char* q = e - kPrefixLen; q >= b; --q)
    if (memcmp(q, kPrefix, kPrefixLen) == 0) { start = q; break; }
  if (!start) return false;
  long long n;
  if (!py_int(start + kPrefixLen, e - 1, &n)) return false;
  *id = n - 10 - (long long)(count % kTokensPerFrame) * 4096;
  return true;
}

struct Stream {
  std::string blob;             // pending token strings, back to back
  std::deque<int32_t> lens;     // their lengths
  size_t head = 0;              // offset of the first pending string in blob
  std::vector<int32_t> ids;     // accepted ids (the reference's `buffer`)
  long long count = 0;
  bool first_done = false, first_pending = false, finished = false, flushed = false;
  void clear() { *this = Stream(); }
};

inline int32_t wrap32(long long v) { return (int32_t)(uint32_t)(unsigned long long)v; }  // np.int64 -> np.int32 cast

}  // namespace

struct snacb_ingest {
  std::vector<Stream> s;
  int64_t accepted = 0, rejected = 0, windows = 0;
};

extern "C" {

int snacb_parse_token(const char* text, int32_t len, int32_t index, int64_t* id) {
  if (!text || len < 0 || !id) return SNACB_EINVAL;
  long long v;
  if (!token_id(text, len, (int)(((index % kTokensPerFrame) + kTokensPerFrame) % kTokensPerFrame), &v)) return 0;
  *id = v;
  return 1;
}

int snacb_ingest_create(snacb_ingest** out, int32_t n_streams) {
  if (!out || n_streams < 0) return SNACB_EINVAL;
  snacb_ingest* g = new (std::nothrow) snacb_ingest();
  if (!g) return SNACB_ENOMEM;
  g->s.resize((size_t)n_streams);
  *out = g;
  return SNACB_OK;
}

void snacb_ingest_destroy(snacb_ingest* g) { delete g; }

int snacb_ingest_reset(snacb_ingest* g, int32_t stream) {
  if (!g || stream < 0 || (size_t)stream >= g->s.size()) return SNACB_EINVAL;
  g->s[(size_t)stream].clear();
  return SNACB_OK;
}

int snacb_ingest_push(snacb_ingest* g, int32_t n, const int32_t* streams, const char* blob, const int64_t* offsets) {
  if (!g || n < 0 || (n > 0 && (!streams || !blob || !offsets))) return SNACB_EINVAL;
  for (int32_t i = 0; i < n; ++i) {
    if (streams[i] < 0 || (size_t)streams[i] >= g->s.size() || offsets[i + 1] < offsets[i]) return SNACB_EINVAL;
    if (g->s[(size_t)streams[i]].finished) return SNACB_ESTATE;
  }
  for (int32_t i = 0; i < n; ++i) {
    Stream& st = g->s[(size_t)streams[i]];
    const int64_t len = offsets[i + 1] - offsets[i];
    if (st.lens.empty() && st.head) { st.blob.clear(); st.head = 0; }
    st.blob.append(blob + offsets[i], (size_t)len);
    st.lens.push_back((int32_t)len);
  }
  return SNACB_OK;
}

int snacb_ingest_finish(snacb_ingest* g, int32_t stream) {
  if (!g || stream < 0 || (size_t)stream >= g->s.size()) return SNACB_EINVAL;
  g->s[(size_t)stream].finished = true;
  return SNACB_OK;
}

// One window per stream per tick (a stream's next window may depend on the outcome of its previous one), slot order.
int32_t snacb_ingest_tick(snacb_ingest* g, int32_t max_win, int32_t* tokens, int32_t tokens_stride, int32_t* ntok,
                          int32_t* stream_of) {
  if (!g || max_win < 0 || (max_win > 0 && (!tokens || !ntok || !stream_of)) || tokens_stride < kLongWindow) return SNACB_EINVAL;
  int32_t n_out = 0;
  for (size_t si = 0; si < g->s.size() && n_out < max_win; ++si) {
    Stream& st = g->s[si];
    if (st.first_pending) continue;  // its first-chunk probe is still being decoded (pipelined ticks): the next window depends on it
    int emit = 0;      // window length in tokens, taken from the end of ids
    bool pad = false;  // end-of-stream window shorter than 28: padded with the last id
    while (!st.lens.empty() && !emit) {
      const int32_t len = st.lens.front();
      st.lens.pop_front();
      long long id;
      const bool ok = token_id(st.blob.data() + st.head, len, (int)(st.count % kTokensPerFrame), &id);
      st.head += (size_t)len;
      if (!ok || id <= 0) { ++g->rejected; continue; }
      st.ids.push_back(wrap32(id));
      ++st.count;
      ++g->accepted;
      if (!st.first_done) {
        if (st.count >= kFirstWindow) { st.first_pending = true; emit = kFirstWindow; }
      } else if (st.count % kTokensPerFrame == 0) {
        if ((int)st.ids.size() >= kLongWindow) emit = kLongWindow;
        else if ((int)st.ids.size() >= kShortWindow) emit = kShortWindow;
      }
    }
    if (!emit && st.lens.empty() && st.finished && !st.flushed) {
      st.flushed = true;
      if ((int)st.ids.size() >= kLongWindow) emit = kLongWindow;
      else if ((int)st.ids.size() >= kShortWindow) emit = kShortWindow;
      else if ((int)st.ids.size() >= kTokensPerFrame) { emit = (int)st.ids.size(); pad = true; }
    }
    if (!emit) continue;
    int32_t* row = tokens + (size_t)n_out * tokens_stride;
    const int32_t* src = st.ids.data() + (st.ids.size() - (size_t)emit);
    memcpy(row, src, (size_t)emit * 4);
    int len_out = emit;
    if (pad) {
      for (int k = emit; k < kShortWindow; ++k) row[k] = src[emit - 1];
      len_out = kShortWindow;
    }
    for (int k = len_out; k < tokens_stride; ++k) row[k] = 0;
    ntok[n_out] = len_out;
    stream_of[n_out] = (int32_t)si;
    ++n_out;
    ++g->windows;
    // the reference keeps the whole buffer; only the last 49 ids can ever be used again
    if (st.ids.size() > 4096) st.ids.erase(st.ids.begin(), st.ids.end() - kLongWindow);
  }
  return n_out;
}

int snacb_ingest_result(snacb_ingest* g, int32_t n, const int32_t* stream_of, const int32_t* status) {
  if (!g || n < 0 || (n > 0 && (!stream_of || !status))) return SNACB_EINVAL;
  for (int32_t i = 0; i < n; ++i) {
    if (stream_of[i] < 0 || (size_t)stream_of[i] >= g->s.size()) return SNACB_EINVAL;
    Stream& st = g->s[(size_t)stream_of[i]];
    if (st.first_pending) {
      st.first_pending = false;
      if (status[i] == SNACB_WIN_OK || status[i] == SNACB_WIN_EMPTY) st.first_done = true;  // convert returned non-None
    }
  }
  return SNACB_OK;
}

int snacb_ingest_done(const snacb_ingest* g, int32_t stream) {
  if (!g || stream < 0 || (size_t)stream >= g->s.size()) return SNACB_EINVAL;
  const Stream& st = g->s[(size_t)stream];
  return (st.finished && st.flushed && st.lens.empty()) ? 1 : 0;
}

int64_t snacb_ingest_stat(const snacb_ingest* g, int32_t which) {
  if (!g) return -1;
  return which == 0 ? g->accepted : which == 1 ? g->rejected : g->windows;
}

}  // extern "C"
