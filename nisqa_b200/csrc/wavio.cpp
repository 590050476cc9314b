// wavio.cpp - native RIFF/WAVE (and FLAC, flac.cpp) ingest (the caller side of the hot path).
// Replaces what lb.load(path, sr=None[, mono=False]) + the ms_channel pick do at reference
// nisqa/NISQA_lib.py:2298-2306 (libsndfile conversion rules, float32 mean over channels).
//
// Two-step C ABI so that the CALLER owns the memory: nisqa_wav_probe() parses the header,
// nisqa_wav_decode() reads the samples straight into the caller's (pinned) batch buffer - mono
// PCM16 is a single fread into place, no intermediate copy.  Both are thread-safe and are driven
// from a Python thread pool (ctypes releases the GIL), one file per task.
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <functional>
#include <thread>
#include <vector>

#include "../../include/nisqa_b200.h"
#include "flac.h"

namespace {

constexpr int kMaxChannels = 256;     // libsndfile's own limit is 1024; no speech corpus comes close

struct WavInfo {
  int tag = 0, channels = 0, bits = 0;
  int32_t sample_rate = 0;
  int64_t data_off = 0, n_frames = 0;
};

bool parse_header(FILE* f, WavInfo* w) {
  unsigned char h[12];
  if (fread(h, 1, 12, f) != 12 || memcmp(h, "RIFF", 4) != 0 || memcmp(h + 8, "WAVE", 4) != 0) return false;
  bool have_fmt = false;
  int64_t pos = 12;
  for (;;) {
    unsigned char ck[8];
    if (fseek(f, pos, SEEK_SET) != 0 || fread(ck, 1, 8, f) != 8) return false;
    const uint32_t sz = (uint32_t)ck[4] | ((uint32_t)ck[5] << 8) | ((uint32_t)ck[6] << 16) | ((uint32_t)ck[7] << 24);
    if (memcmp(ck, "fmt ", 4) == 0) {
      unsigned char b[40] = {0};
      const size_t want = sz < 40 ? sz : 40;
      if (sz < 16 || fread(b, 1, want, f) != want) return false;
      w->tag = b[0] | (b[1] << 8);
      w->channels = b[2] | (b[3] << 8);
      w->sample_rate = (int32_t)((uint32_t)b[4] | ((uint32_t)b[5] << 8) | ((uint32_t)b[6] << 16) | ((uint32_t)b[7] << 24));
      w->bits = b[14] | (b[15] << 8);
      const int block_align = b[12] | (b[13] << 8);
      if (w->tag == 0xFFFE && sz >= 26) w->tag = b[24] | (b[25] << 8);   // WAVE_FORMAT_EXTENSIBLE
      // implausible / inconsistent headers are "Could not load file", not a multi-GB decode buffer
      if (w->channels < 1 || w->channels > kMaxChannels || w->sample_rate < 1) return false;
      if (w->bits < 8 || w->bits > 64 || (w->bits & 7) || block_align != w->channels * (w->bits / 8)) return false;
      have_fmt = true;
    } else if (memcmp(ck, "data", 4) == 0) {
      if (!have_fmt || w->channels < 1) return false;
      const int width = w->bits / 8;
      if (width < 1) return false;
      // clamp to the real file size (streamed files may carry a bogus length)
      fseek(f, 0, SEEK_END);
      const int64_t fsize = ftell(f);
      int64_t bytes = sz;
      if (pos + 8 + bytes > fsize) bytes = fsize - (pos + 8);
      if (bytes < 0) return false;
      w->data_off = pos + 8;
      w->n_frames = bytes / ((int64_t)width * w->channels);
      return true;
    }
    pos += 8 + (int64_t)sz + (sz & 1);
  }
}

bool supported(const WavInfo& w) {
  if (w.tag == 1) return w.bits == 8 || w.bits == 16 || w.bits == 24 || w.bits == 32;
  if (w.tag == 3) return w.bits == 32 || w.bits == 64;
  if (w.tag == 6 || w.tag == 7) return w.bits == 8;            // G.711 A-law / mu-law (telephony corpora)
  return false;
}

// ITU-T G.711 expansion of one code to a 16-bit linear sample (the tables of libsndfile's alaw.c / ulaw.c)
inline int g711_expand(unsigned char code, bool alaw) {
  if (alaw) {
    const int c = code ^ 0x55, seg = (c & 0x70) >> 4;
    int t = (c & 0x0F) << 4;
    t = (seg == 0) ? t + 8 : (t + 0x108) << (seg - 1);
    return (c & 0x80) ? t : -t;
  }
  const int c = ~code & 0xFF;
  const int t = (((c & 0x0F) << 3) + 0x84) << ((c & 0x70) >> 4);
  return (c & 0x80) ? 0x84 - t : t - 0x84;
}

// libsndfile's conversion of one stored sample to float32
inline float sample_f32(const unsigned char* p, const WavInfo& w) {
  if (w.tag == 1) {
    switch (w.bits) {
      case 8: return ((float)p[0] - 128.0f) * (1.0f / 128.0f);
      case 16: return (float)(int16_t)(p[0] | (p[1] << 8)) * (1.0f / 32768.0f);
      case 24: {
        int32_t v = p[0] | (p[1] << 8) | (p[2] << 16);
        v = (v ^ 0x800000) - 0x800000;
        return (float)v * (1.0f / 8388608.0f);
      }
      default: {
        const int32_t v = (int32_t)((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24));
        return (float)((double)v * (1.0 / 2147483648.0));
      }
    }
  }
  if (w.tag == 6 || w.tag == 7) return (float)g711_expand(p[0], w.tag == 6) * (1.0f / 32768.0f);
  if (w.bits == 32) { float v; memcpy(&v, p, 4); return v; }
  double d; memcpy(&d, p, 8); return (float)d;
}

}  // namespace

extern "C" {

// Header only.  kind_out: NISQA_FMT_S16 if the clip can be delivered as int16 without loss (PCM16 and
// either mono or a channel pick), else NISQA_FMT_F32.  Returns 0, or NISQA_ERR_INVALID for anything
// the reference would answer with "Could not load file".
int nisqa_wav_probe(const char* path, int32_t ms_channel, int32_t* sample_rate, int64_t* n_frames,
                    int32_t* channels, int32_t* kind_out) {
  if (!path) return NISQA_ERR_INVALID;
  if (nisqa::flac_is(path)) {       // FLAC: integers of 8..32 bits, delivered like PCM of that width
    nisqa::FlacInfo fi;
    if (!nisqa::flac_probe(path, &fi) || fi.channels < 1 || fi.channels > 8) return NISQA_ERR_INVALID;
    if (ms_channel >= 0 && fi.channels > 1 && ms_channel >= fi.channels) return NISQA_ERR_INVALID;
    if (sample_rate) *sample_rate = fi.sample_rate;
    if (n_frames) *n_frames = fi.n_frames;
    if (channels) *channels = fi.channels;
    if (kind_out) *kind_out = (fi.bits == 16 && (fi.channels == 1 || ms_channel >= 0)) ? NISQA_FMT_S16 : NISQA_FMT_F32;
    return 0;
  }
  FILE* f = fopen(path, "rb");
  if (!f) return NISQA_ERR_INVALID;
  WavInfo w;
  const bool ok = parse_header(f, &w) && supported(w);
  fclose(f);
  if (!ok) return NISQA_ERR_INVALID;
  if (ms_channel >= 0 && w.channels > 1 && ms_channel >= w.channels) return NISQA_ERR_INVALID;
  if (sample_rate) *sample_rate = w.sample_rate;
  if (n_frames) *n_frames = w.n_frames;
  if (channels) *channels = w.channels;
  if (kind_out) *kind_out = (w.tag == 1 && w.bits == 16 && (w.channels == 1 || ms_channel >= 0)) ? NISQA_FMT_S16 : NISQA_FMT_F32;
  return 0;
}

// Decode into dst (capacity cap_frames samples of out_fmt).  out_fmt NISQA_FMT_S16 is only legal when
// probe reported S16; NISQA_FMT_F32 is always legal (PCM16 is then scaled by 1/32768).
// ms_channel < 0: mono mix = float32 mean over the channels (librosa.to_mono); else that channel
// (ignored for mono files, lib:2301).  Returns the number of frames written, or a negative status.
// FLAC: decode to integers (flac.cpp), then libsndfile's rules - int16 as is, float = v * 2^-(bits-1), mono mix = float32
// mean over the channels
static int64_t flac_decode_into(const char* path, int32_t ms_channel, int32_t out_fmt, void* dst, int64_t cap_frames) {
  nisqa::FlacInfo fi;
  std::vector<int32_t> pcm;
  if (!nisqa::flac_decode_all(path, &fi, &pcm) || fi.n_frames > cap_frames) return NISQA_ERR_INVALID;
  const int ch = fi.channels;
  if (ms_channel >= 0 && ch > 1 && ms_channel >= ch) return NISQA_ERR_INVALID;
  const int pick = (ch > 1 && ms_channel >= 0) ? ms_channel : -1;
  if (out_fmt == NISQA_FMT_S16) {
    if (!(fi.bits == 16 && (ch == 1 || pick >= 0))) return NISQA_ERR_INVALID;
    int16_t* o = static_cast<int16_t*>(dst);
    for (int64_t i = 0; i < fi.n_frames; ++i) o[i] = (int16_t)pcm[(size_t)i * ch + (pick >= 0 ? pick : 0)];
  } else if (out_fmt == NISQA_FMT_F32) {
    float* o = static_cast<float*>(dst);
    const float scale = 1.0f / (float)(1u << (fi.bits - 1));
    for (int64_t i = 0; i < fi.n_frames; ++i) {
      const int32_t* p = pcm.data() + (size_t)i * ch;
      if (ch == 1) o[i] = (float)p[0] * scale;
      else if (pick >= 0) o[i] = (float)p[pick] * scale;
      else {
        float s2 = 0.f;
        for (int c = 0; c < ch; ++c) s2 += (float)p[c] * scale;
        o[i] = s2 / (float)ch;
      }
    }
  } else {
    return NISQA_ERR_INVALID;
  }
  return fi.n_frames;
}

static int64_t wav_decode_impl(const char* path, int32_t ms_channel, int32_t out_fmt, void* dst, int64_t cap_frames) {
  if (!path || !dst) return NISQA_ERR_INVALID;
  if (nisqa::flac_is(path)) return flac_decode_into(path, ms_channel, out_fmt, dst, cap_frames);
  FILE* f = fopen(path, "rb");
  if (!f) return NISQA_ERR_INVALID;
  WavInfo w;
  if (!parse_header(f, &w) || !supported(w) || w.n_frames > cap_frames ||
      (ms_channel >= 0 && w.channels > 1 && ms_channel >= w.channels)) { fclose(f); return NISQA_ERR_INVALID; }
  const int width = w.bits / 8, ch = w.channels;
  const int pick = (ch > 1 && ms_channel >= 0) ? ms_channel : -1;
  if (fseek(f, w.data_off, SEEK_SET) != 0) { fclose(f); return NISQA_ERR_INVALID; }
  int64_t done = 0;
  if (out_fmt == NISQA_FMT_S16) {
    if (!(w.tag == 1 && w.bits == 16 && (ch == 1 || pick >= 0))) { fclose(f); return NISQA_ERR_INVALID; }
    int16_t* o = static_cast<int16_t*>(dst);
    if (ch == 1) {                                   // straight into place (little-endian host), no stdio copy
      const int fd = fileno(f);
      size_t want = (size_t)w.n_frames * 2, off = 0;
      while (off < want) {
        const ssize_t got = pread(fd, reinterpret_cast<char*>(o) + off, want - off, (off_t)(w.data_off + (int64_t)off));
        if (got <= 0) break;
        off += (size_t)got;
      }
      done = (int64_t)(off / 2);
    } else {
      std::vector<int16_t> buf((size_t)8192 * ch);
      while (done < w.n_frames) {
        const int64_t want = std::min<int64_t>(8192, w.n_frames - done);
        const int64_t got = (int64_t)fread(buf.data(), (size_t)2 * ch, (size_t)want, f);
        if (got <= 0) break;
        for (int64_t i = 0; i < got; ++i) o[done + i] = buf[(size_t)i * ch + pick];
        done += got;
      }
    }
  } else if (out_fmt == NISQA_FMT_F32) {
    float* o = static_cast<float*>(dst);
    std::vector<unsigned char> buf((size_t)8192 * ch * width);
    while (done < w.n_frames) {
      const int64_t want = std::min<int64_t>(8192, w.n_frames - done);
      const int64_t got = (int64_t)fread(buf.data(), (size_t)width * ch, (size_t)want, f);
      if (got <= 0) break;
      for (int64_t i = 0; i < got; ++i) {
        const unsigned char* p = buf.data() + (size_t)i * ch * width;
        if (ch == 1) o[done + i] = sample_f32(p, w);
        else if (pick >= 0) o[done + i] = sample_f32(p + (size_t)pick * width, w);
        else {
          // numpy.mean(axis=0) over float32 channels: sequential float32 sum, then / n
          float s = 0.f;
          for (int c = 0; c < ch; ++c) s += sample_f32(p + (size_t)c * width, w);
          o[done + i] = s / (float)ch;
        }
      }
      done += got;
    }
  } else { fclose(f); return NISQA_ERR_INVALID; }
  fclose(f);
  return done == w.n_frames ? done : (int64_t)NISQA_ERR_INVALID;
}

// no exception crosses the ABI (and none may escape a std::thread of run_parallel: that is std::terminate)
int64_t nisqa_wav_decode(const char* path, int32_t ms_channel, int32_t out_fmt, void* dst, int64_t cap_frames) {
  try {
    return wav_decode_impl(path, ms_channel, out_fmt, dst, cap_frames);
  } catch (...) {
    return NISQA_ERR_INVALID;
  }
}


// ---- whole-batch forms: one call per batch, the files are spread over n_threads native threads
// (per-file calls from Python are bound by interpreter overhead at a few thousand files per second).
static void run_parallel(int n, int n_threads, const std::function<void(int)>& fn) {
  n_threads = std::max(1, std::min(n_threads, n));
  if (n_threads == 1) { for (int i = 0; i < n; ++i) fn(i); return; }
  std::atomic<int> next(0);
  std::vector<std::thread> pool;
  for (int t = 0; t < n_threads; ++t)
    pool.emplace_back([&]() { for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) fn(i); });
  for (auto& th : pool) th.join();
}

// status[i] = 0 or NISQA_ERR_INVALID per file; returns the number of files that failed.
int nisqa_wav_probe_batch(int n, const char* const* paths, int32_t ms_channel, int n_threads,
                          int32_t* sample_rate, int64_t* n_frames, int32_t* kind, int32_t* status) {
  if (n < 0 || (n > 0 && (!paths || !sample_rate || !n_frames || !kind || !status))) return NISQA_ERR_INVALID;
  std::atomic<int> bad(0);
  run_parallel(n, n_threads, [&](int i) {
    int32_t ch = 0;
    status[i] = nisqa_wav_probe(paths[i], ms_channel, &sample_rate[i], &n_frames[i], &ch, &kind[i]);
    if (status[i] != 0) { sample_rate[i] = 0; n_frames[i] = 0; kind[i] = NISQA_FMT_F32; bad.fetch_add(1); }
  });
  return bad.load();
}

// clip i is decoded to base + elem_offsets[i] * sizeof(out_fmt element), capacity cap_frames[i]
int nisqa_wav_decode_batch(int n, const char* const* paths, int32_t ms_channel, int32_t out_fmt, void* base,
                           const int64_t* elem_offsets, const int64_t* cap_frames, int n_threads,
                           int32_t* status) {
  if (n < 0 || (n > 0 && (!paths || !base || !elem_offsets || !cap_frames || !status))) return NISQA_ERR_INVALID;
  const size_t esz = out_fmt == NISQA_FMT_S16 ? 2 : 4;
  std::atomic<int> bad(0);
  run_parallel(n, n_threads, [&](int i) {
    const int64_t got = nisqa_wav_decode(paths[i], ms_channel, out_fmt, static_cast<char*>(base) + (size_t)elem_offsets[i] * esz, cap_frames[i]);
    status[i] = got < 0 ? (int32_t)got : 0;
    if (got < 0) bad.fetch_add(1);
  });
  return bad.load();
}

}  // extern "C"
