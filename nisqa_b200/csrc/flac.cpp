// flac.cpp - native FLAC reader of the ingest ("non-WAV formats that libsndfile handled").
// The reference reads every file through lb.load -> soundfile -> libsndfile (nisqa/NISQA_lib.py:2298-2306), which
// decodes FLAC to integers and scales them by 2^-(bits-1) when float is asked for; this file does the same for the
// FLAC subset every encoder produces (the format's "subset" and beyond: all four subframe types, fixed predictors 0-4,
// LPC up to order 32, Rice and Rice2 partitions with escape codes, wasted bits, the three stereo decorrelations, 8-32 bit
// samples, variable block sizes), with the frame header's CRC-8 and the frame's CRC-16 verified - a damaged file is
// "Could not load file", never silently wrong samples.  Written from the format specification (xiph.org FLAC format /
// RFC 9639); PARITY UNPINNED like the rest of the ingest's third-party formats: there is no FLAC encoder or decoder in
// this environment, the tests round-trip through an independent Python encoder (tests/flac_enc.py).
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "flac.h"

namespace nisqa {
namespace {

// MSB-first bit reader over the file image: a 64-bit window refilled byte-wise, unary runs counted with clz
struct BitReader {
  const uint8_t* p;
  size_t n;
  size_t byte = 0;          // next byte to load into the window
  uint64_t win = 0;         // the unread bits, left-aligned
  int have = 0;             // valid bits in `win`
  bool bad = false;
  BitReader(const uint8_t* p_, size_t n_) : p(p_), n(n_) {}
  void refill() {
    while (have <= 56 && byte < n) { win |= (uint64_t)p[byte++] << (56 - have); have += 8; }
  }
  uint32_t bits(int k) {    // k <= 32
    if (k == 0) return 0;
    if (have < k) { refill(); if (have < k) { bad = true; have = 0; win = 0; return 0; } }
    const uint32_t v = (uint32_t)(win >> (64 - k));
    win <<= k; have -= k;
    return v;
  }
  int64_t sbits(int k) {    // two's complement, k <= 33 (side channel of 32-bit streams)
    if (k == 0) return 0;
    uint64_t v;
    if (k > 32) { v = (uint64_t)bits(k - 32) << 32; v |= bits(32); }
    else v = bits(k);
    const uint64_t sign = 1ull << (k - 1);
    return (int64_t)((v ^ sign) - sign);
  }
  uint32_t unary() {        // number of 0 bits before the next 1
    uint32_t q = 0;
    for (;;) {
      if (have == 0) { refill(); if (have == 0) { bad = true; return 0; } }
      if (win == 0) { q += (uint32_t)have; have = 0; if (q > (1u << 24)) { bad = true; return 0; } continue; }
      const int z = __builtin_clzll(win);          // < have: a set bit lies inside the valid part
      q += (uint32_t)z;
      win <<= (z + 1); have -= z + 1;
      return q;
    }
  }
  size_t bit_pos() const { return byte * 8 - (size_t)have; }
  void align() { const int drop = have & 7; win <<= drop; have -= drop; }
};

struct CrcTables {
  uint8_t t8[256];
  uint16_t t16[256];
  CrcTables() {
    for (int i = 0; i < 256; ++i) {
      uint8_t c = (uint8_t)i;                       // polynomial x^8 + x^2 + x + 1
      for (int b = 0; b < 8; ++b) c = (uint8_t)((c & 0x80) ? (c << 1) ^ 0x07 : (c << 1));
      t8[i] = c;
      uint16_t w = (uint16_t)(i << 8);              // polynomial x^16 + x^15 + x^2 + 1
      for (int b = 0; b < 8; ++b) w = (uint16_t)((w & 0x8000) ? (w << 1) ^ 0x8005 : (w << 1));
      t16[i] = w;
    }
  }
};
const CrcTables kCrc;

uint8_t crc8(const uint8_t* p, size_t n) {
  uint8_t c = 0;
  for (size_t i = 0; i < n; ++i) c = kCrc.t8[c ^ p[i]];
  return c;
}
uint16_t crc16(const uint8_t* p, size_t n) {
  uint16_t c = 0;
  for (size_t i = 0; i < n; ++i) c = (uint16_t)((c << 8) ^ kCrc.t16[(c >> 8) ^ p[i]]);
  return c;
}

bool read_file(const char* path, std::vector<uint8_t>* out) {
  FILE* f = fopen(path, "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END);
  const long sz = ftell(f);
  if (sz < 0 || sz > (1l << 31)) { fclose(f); return false; }
  out->resize((size_t)sz);
  fseek(f, 0, SEEK_SET);
  const bool ok = sz == 0 || fread(out->data(), 1, (size_t)sz, f) == (size_t)sz;
  fclose(f);
  return ok;
}

// metadata: returns the byte offset of the first frame, 0 on error
size_t parse_metadata(const std::vector<uint8_t>& d, FlacInfo* info) {
  if (d.size() < 42 || memcmp(d.data(), "fLaC", 4) != 0) return 0;
  size_t pos = 4;
  bool have_si = false;
  for (;;) {
    if (pos + 4 > d.size()) return 0;
    const bool last = (d[pos] & 0x80) != 0;
    const int type = d[pos] & 0x7F;
    const size_t len = ((size_t)d[pos + 1] << 16) | ((size_t)d[pos + 2] << 8) | d[pos + 3];
    pos += 4;
    if (pos + len > d.size()) return 0;
    if (type == 0) {
      if (len < 34) return 0;
      const uint8_t* s = d.data() + pos;
      info->min_block = (s[0] << 8) | s[1];
      info->max_block = (s[2] << 8) | s[3];
      info->sample_rate = (int32_t)(((uint32_t)s[10] << 12) | ((uint32_t)s[11] << 4) | (s[12] >> 4));
      info->channels = ((s[12] >> 1) & 7) + 1;
      info->bits = (((s[12] & 1) << 4) | (s[13] >> 4)) + 1;
      info->n_frames = ((int64_t)(s[13] & 0x0F) << 32) | ((int64_t)s[14] << 24) | ((int64_t)s[15] << 16) | ((int64_t)s[16] << 8) | s[17];
      have_si = true;
    }
    pos += len;
    if (last) break;
  }
  if (!have_si || info->sample_rate < 1 || info->bits < 4 || info->bits > 32) return 0;
  return pos;
}

bool read_residual(BitReader& br, int order, int blocksize, int64_t* out /*[blocksize], residual at i >= order*/) {
  const int method = (int)br.bits(2);
  if (method > 1) return false;
  const int pbits = method == 0 ? 4 : 5;
  const int escape = method == 0 ? 15 : 31;
  const int porder = (int)br.bits(4);
  const int nparts = 1 << porder;
  if ((blocksize >> porder) << porder != blocksize && porder > 0) return false;
  int i = order;
  for (int part = 0; part < nparts; ++part) {
    int count = blocksize >> porder;
    if (part == 0) count -= order;
    if (count < 0) return false;
    const int k = (int)br.bits(pbits);
    if (k == escape) {
      const int raw = (int)br.bits(5);
      for (int j = 0; j < count; ++j) out[i++] = raw ? br.sbits(raw) : 0;
    } else {
      for (int j = 0; j < count; ++j) {
        const uint64_t q = br.unary();
        const uint64_t v = (q << k) | (k ? br.bits(k) : 0u);
        out[i++] = (int64_t)(v >> 1) ^ -(int64_t)(v & 1);
      }
    }
    if (br.bad) return false;
  }
  return i == blocksize;
}

bool read_subframe(BitReader& br, int bps, int blocksize, int64_t* s) {
  if (br.bits(1) != 0) return false;
  const int type = (int)br.bits(6);
  int wasted = 0;
  if (br.bits(1)) wasted = (int)br.unary() + 1;
  const int eb = bps - wasted;
  if (eb < 1 || br.bad) return false;
  if (type == 0) {                                   // CONSTANT
    const int64_t v = br.sbits(eb);
    for (int i = 0; i < blocksize; ++i) s[i] = v;
  } else if (type == 1) {                            // VERBATIM
    for (int i = 0; i < blocksize; ++i) s[i] = br.sbits(eb);
  } else if (type >= 8 && type <= 12) {              // FIXED, order type - 8
    const int o = type - 8;
    if (o > blocksize) return false;
    for (int i = 0; i < o; ++i) s[i] = br.sbits(eb);
    if (!read_residual(br, o, blocksize, s)) return false;
    for (int i = o; i < blocksize; ++i) {
      switch (o) {
        case 0: break;
        case 1: s[i] += s[i - 1]; break;
        case 2: s[i] += 2 * s[i - 1] - s[i - 2]; break;
        case 3: s[i] += 3 * s[i - 1] - 3 * s[i - 2] + s[i - 3]; break;
        default: s[i] += 4 * s[i - 1] - 6 * s[i - 2] + 4 * s[i - 3] - s[i - 4]; break;
      }
    }
  } else if (type >= 32) {                           // LPC, order type - 31
    const int o = type - 31;
    if (o > blocksize) return false;
    for (int i = 0; i < o; ++i) s[i] = br.sbits(eb);
    const int prec = (int)br.bits(4) + 1;
    if (prec == 16) return false;
    const int shift = (int)br.sbits(5);
    if (shift < 0) return false;
    int64_t coef[32];
    for (int j = 0; j < o; ++j) coef[j] = br.sbits(prec);
    if (!read_residual(br, o, blocksize, s)) return false;
    for (int i = o; i < blocksize; ++i) {
      int64_t pred = 0;
      for (int j = 0; j < o; ++j) pred += coef[j] * s[i - 1 - j];
      s[i] += pred >> shift;
    }
  } else {
    return false;                                    // reserved subframe type
  }
  if (wasted)
    for (int i = 0; i < blocksize; ++i) s[i] = (int64_t)((uint64_t)s[i] << wasted);
  return !br.bad;
}

}  // namespace

bool flac_is(const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) return false;
  char m[4];
  const bool ok = fread(m, 1, 4, f) == 4 && memcmp(m, "fLaC", 4) == 0;
  fclose(f);
  return ok;
}

bool flac_probe(const char* path, FlacInfo* info) {
  std::vector<uint8_t> d;
  if (!read_file(path, &d)) return false;
  FlacInfo fi;
  if (!parse_metadata(d, &fi)) return false;
  if (fi.n_frames == 0) {          // unknown length in STREAMINFO (streamed encoders): count by decoding
    std::vector<int32_t> all;
    if (!flac_decode_all(path, &fi, &all)) return false;
  }
  *info = fi;
  return true;
}

// Decodes every frame: samples interleaved [frame][channel] as int32.  info->n_frames is set to the decoded count when
// STREAMINFO does not carry it; a mismatch with a declared count is an error.
bool flac_decode_all(const char* path, FlacInfo* info, std::vector<int32_t>* out) {
  std::vector<uint8_t> d;
  if (!read_file(path, &d)) return false;
  FlacInfo fi;
  size_t pos = parse_metadata(d, &fi);
  if (!pos) return false;
  const int ch = fi.channels;
  out->clear();
  // (a declared length is only a hint until the frames confirm it: a damaged header must not reserve gigabytes)
  if (fi.n_frames > 0) out->reserve((size_t)std::min<int64_t>(fi.n_frames, (int64_t)d.size() * 64) * ch);
  std::vector<int64_t> sub[8];
  int64_t total = 0;
  while (pos + 6 <= d.size()) {
    const uint8_t* h = d.data() + pos;
    if (h[0] != 0xFF || (h[1] & 0xFE) != 0xF8) return false;                   // sync code 11111111 111110 + reserved 0
    BitReader br(h, d.size() - pos);
    br.bits(16);
    const int bs_code = (int)br.bits(4), sr_code = (int)br.bits(4), ch_code = (int)br.bits(4), ss_code = (int)br.bits(3);
    if (br.bits(1) != 0) return false;
    {                                                                           // UTF-8 style coded frame / sample number
      const uint32_t b0 = br.bits(8);
      int extra = 0;
      if (b0 & 0x80) { uint32_t m = 0x40; while (b0 & m) { ++extra; m >>= 1; } if (extra == 0 || extra > 6) return false; }
      for (int i = 0; i < extra; ++i) if ((br.bits(8) & 0xC0) != 0x80) return false;
    }
    int blocksize;
    if (bs_code == 0) return false;
    else if (bs_code == 1) blocksize = 192;
    else if (bs_code <= 5) blocksize = 576 << (bs_code - 2);
    else if (bs_code == 6) blocksize = (int)br.bits(8) + 1;
    else if (bs_code == 7) blocksize = (int)br.bits(16) + 1;
    else blocksize = 256 << (bs_code - 8);
    if (sr_code == 12) br.bits(8);
    else if (sr_code == 13 || sr_code == 14) br.bits(16);
    else if (sr_code == 15) return false;
    if (br.bad) return false;
    const size_t hdr_bytes = br.bit_pos() >> 3;
    if (crc8(h, hdr_bytes) != (uint8_t)br.bits(8)) return false;
    static const int ss_table[8] = {0, 8, 12, -1, 16, 20, 24, 32};
    const int bps = ss_code == 0 ? fi.bits : ss_table[ss_code];
    if (bps < 4 || bps != fi.bits) return false;                                // (one sample size per stream)
    int n_ch;
    if (ch_code < 8) n_ch = ch_code + 1;
    else if (ch_code <= 10) n_ch = 2;
    else return false;
    if (n_ch != ch) return false;
    for (int c = 0; c < ch; ++c) {
      sub[c].resize((size_t)blocksize);
      const bool side = (ch_code == 8 && c == 1) || (ch_code == 9 && c == 0) || (ch_code == 10 && c == 1);
      if (!read_subframe(br, bps + (side ? 1 : 0), blocksize, sub[c].data())) return false;
    }
    br.align();
    const size_t body = br.bit_pos() >> 3;
    if (pos + body + 2 > d.size()) return false;
    if (crc16(h, body) != (uint16_t)((h[body] << 8) | h[body + 1])) return false;
    for (int i = 0; i < blocksize; ++i) {
      int64_t a = sub[0][i], b = ch > 1 ? sub[1][i] : 0;
      if (ch_code == 8) b = a - b;                              // left, side  -> right = left - side
      else if (ch_code == 9) a = a + b;                         // side, right -> left = side + right
      else if (ch_code == 10) {                                 // mid, side
        const int64_t mid = (a << 1) | (b & 1), sd = b;
        a = (mid + sd) >> 1; b = (mid - sd) >> 1;
      }
      out->push_back((int32_t)a);
      if (ch > 1) out->push_back((int32_t)b);
      for (int c = 2; c < ch; ++c) out->push_back((int32_t)sub[c][i]);
    }
    total += blocksize;
    pos += body + 2;
    if (fi.n_frames > 0 && total >= fi.n_frames) break;
  }
  if (fi.n_frames > 0) {
    if (total < fi.n_frames) return false;
    out->resize((size_t)fi.n_frames * ch);                       // (the last block of a stream may be padded by some encoders)
  } else {
    fi.n_frames = total;
  }
  *info = fi;
  return true;
}

}  // namespace nisqa
