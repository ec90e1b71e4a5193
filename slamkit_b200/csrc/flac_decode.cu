// Host-side FLAC decoder: the image has no audio decoder (torchaudio 2.11 needs torchcodec, no
// soundfile / ffmpeg), and the reference's example data and LibriSpeech-style corpora are FLAC.  This is a small
// from-the-spec decoder for the subset that matters here (<= 8 channels, <= 24 bits/sample, constant / verbatim / fixed /
// LPC subframes, Rice and Rice2 residuals with escape partitions, all four stereo decorrelation modes, CRC-8 / CRC-16
// checked).  Pure host code -- it only lives in this library so that cli/extract_features.py can call it via ctypes.
// Replaces `torchaudio.info(...).num_frames` and `torchaudio.load(...)` + channel mean (cli/extract_features.py:45-57).
#include "kernels.h"
#include "../../include/slamkit_b200.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#define SK_TRY_RC2(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)

namespace {

struct BitReader {
  const uint8_t* p;
  size_t n, pos = 0;   // byte position
  uint64_t acc = 0;
  int nbits = 0;
  bool eof = false;
  BitReader(const uint8_t* d, size_t len) : p(d), n(len) {}
  uint32_t read(int bits) {
    while (nbits < bits) {
      if (pos >= n) { eof = true; acc <<= 8; }
      else acc = (acc << 8) | p[pos++];
      nbits += 8;
    }
    nbits -= bits;
    const uint32_t v = (uint32_t)((acc >> nbits) & ((bits == 32) ? 0xffffffffull : ((1ull << bits) - 1)));
    acc &= (1ull << nbits) - 1;
    return v;
  }
  int32_t read_signed(int bits) {
    if (bits == 0) return 0;
    const uint32_t v = read(bits);
    return (int32_t)(v << (32 - bits)) >> (32 - bits);
  }
  void align() { nbits -= nbits % 8; acc &= (nbits ? ((1ull << nbits) - 1) : 0); }
  size_t byte_pos() const { return pos - nbits / 8; }
  int64_t read_rice(int k) {
    uint32_t q = 0;
    while (read(1) == 0) {
      if (eof) return 0;
      ++q;
    }
    const uint32_t v = (q << k) | (k ? read(k) : 0);
    return (int64_t)(v >> 1) ^ -(int64_t)(v & 1);
  }
};

uint8_t crc8(const uint8_t* d, size_t n) {
  uint8_t c = 0;
  for (size_t i = 0; i < n; ++i) {
    c ^= d[i];
    for (int b = 0; b < 8; ++b) c = (c & 0x80) ? (uint8_t)((c << 1) ^ 0x07) : (uint8_t)(c << 1);
  }
  return c;
}
uint16_t crc16(const uint8_t* d, size_t n) {
  uint16_t c = 0;
  for (size_t i = 0; i < n; ++i) {
    c ^= (uint16_t)d[i] << 8;
    for (int b = 0; b < 8; ++b) c = (c & 0x8000) ? (uint16_t)((c << 1) ^ 0x8005) : (uint16_t)(c << 1);
  }
  return c;
}

struct StreamInfo {
  int sample_rate = 0, channels = 0, bps = 0;
  int64_t total = 0;
  uint8_t md5[16];
  size_t audio_start = 0;
};

int read_file(const char* path, std::vector<uint8_t>& out) {
  FILE* f = fopen(path, "rb");
  SK_REQUIRE(f, "flac: cannot open %s", path);
  fseek(f, 0, SEEK_END);
  const long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  out.resize(n > 0 ? n : 0);
  const size_t got = n > 0 ? fread(out.data(), 1, n, f) : 0;
  fclose(f);
  SK_REQUIRE((long)got == n, "flac: short read on %s", path);
  return 0;
}

int parse_header(const std::vector<uint8_t>& d, StreamInfo& si) {
  SK_REQUIRE(d.size() > 42 && memcmp(d.data(), "fLaC", 4) == 0, "flac: missing fLaC marker");
  size_t pos = 4;
  bool last = false, have = false;
  while (!last) {
    SK_REQUIRE(pos + 4 <= d.size(), "flac: truncated metadata");
    last = d[pos] & 0x80;
    const int type = d[pos] & 0x7f;
    const size_t len = ((size_t)d[pos + 1] << 16) | ((size_t)d[pos + 2] << 8) | d[pos + 3];
    pos += 4;
    SK_REQUIRE(pos + len <= d.size(), "flac: truncated metadata block");
    if (type == 0) {
      SK_REQUIRE(len >= 34, "flac: short STREAMINFO");
      const uint8_t* s = d.data() + pos;
      si.sample_rate = (s[10] << 12) | (s[11] << 4) | (s[12] >> 4);
      si.channels = ((s[12] >> 1) & 7) + 1;
      si.bps = (((s[12] & 1) << 4) | (s[13] >> 4)) + 1;
      si.total = ((int64_t)(s[13] & 0xf) << 32) | ((int64_t)s[14] << 24) | (s[15] << 16) | (s[16] << 8) | s[17];
      memcpy(si.md5, s + 18, 16);
      have = true;
    }
    pos += len;
  }
  SK_REQUIRE(have, "flac: no STREAMINFO block");
  si.audio_start = pos;
  return 0;
}

int decode_residual(BitReader& br, int order, int blocksize, int32_t* out /* starts at sample `order` */) {
  const int method = br.read(2);
  SK_REQUIRE(method < 2, "flac: reserved residual coding method");
  const int pbits = method == 0 ? 4 : 5;
  const int escape = method == 0 ? 15 : 31;
  const int porder = br.read(4);
  const int parts = 1 << porder;
  SK_REQUIRE((blocksize % parts) == 0 || porder == 0, "flac: bad partition order");
  int idx = 0;
  for (int p = 0; p < parts; ++p) {
    int count = (blocksize >> porder) - (p == 0 ? order : 0);
    SK_REQUIRE(count >= 0, "flac: negative partition size");
    const int k = br.read(pbits);
    if (k == escape) {
      const int nb = br.read(5);
      for (int i = 0; i < count; ++i) out[idx++] = br.read_signed(nb);
    } else {
      for (int i = 0; i < count; ++i) out[idx++] = (int32_t)br.read_rice(k);
    }
    SK_REQUIRE(!br.eof, "flac: unexpected end of stream in residual");
  }
  return 0;
}

int decode_subframe(BitReader& br, int bps, int blocksize, int32_t* s) {
  SK_REQUIRE(br.read(1) == 0, "flac: subframe padding bit set");
  const int type = br.read(6);
  int wasted = 0;
  if (br.read(1)) {
    wasted = 1;
    while (br.read(1) == 0) {
      ++wasted;
      SK_REQUIRE(!br.eof, "flac: eof in wasted bits");
    }
  }
  bps -= wasted;
  if (type == 0) {
    const int32_t v = br.read_signed(bps);
    for (int i = 0; i < blocksize; ++i) s[i] = v;
  } else if (type == 1) {
    for (int i = 0; i < blocksize; ++i) s[i] = br.read_signed(bps);
  } else if (type >= 8 && type <= 12) {
    const int order = type - 8;
    SK_REQUIRE(order <= blocksize, "flac: fixed order > block size");
    for (int i = 0; i < order; ++i) s[i] = br.read_signed(bps);
    SK_TRY_RC2(decode_residual(br, order, blocksize, s + order));
    for (int i = order; i < blocksize; ++i) {
      int64_t pred = 0;
      switch (order) {
        case 1: pred = s[i - 1]; break;
        case 2: pred = 2 * (int64_t)s[i - 1] - s[i - 2]; break;
        case 3: pred = 3 * (int64_t)s[i - 1] - 3 * (int64_t)s[i - 2] + s[i - 3]; break;
        case 4: pred = 4 * (int64_t)s[i - 1] - 6 * (int64_t)s[i - 2] + 4 * (int64_t)s[i - 3] - s[i - 4]; break;
        default: break;
      }
      s[i] = (int32_t)(s[i] + pred);
    }
  } else if (type >= 32) {
    const int order = (type & 31) + 1;
    SK_REQUIRE(order <= blocksize, "flac: LPC order > block size");
    for (int i = 0; i < order; ++i) s[i] = br.read_signed(bps);
    const int prec = br.read(4) + 1;
    SK_REQUIRE(prec != 16, "flac: invalid LPC precision");
    const int shift = br.read_signed(5);
    SK_REQUIRE(shift >= 0, "flac: negative LPC shift");
    int32_t coef[32];
    for (int i = 0; i < order; ++i) coef[i] = br.read_signed(prec);
    SK_TRY_RC2(decode_residual(br, order, blocksize, s + order));
    for (int i = order; i < blocksize; ++i) {
      int64_t sum = 0;
      for (int j = 0; j < order; ++j) sum += (int64_t)coef[j] * s[i - 1 - j];
      s[i] = (int32_t)(s[i] + (sum >> shift));
    }
  } else {
    SK_REQUIRE(false, "flac: reserved subframe type %d", type);
  }
  if (wasted)
    for (int i = 0; i < blocksize; ++i) s[i] = (int32_t)((uint32_t)s[i] << wasted);
  return 0;
}

// Decodes the whole stream into interleaved int32 PCM (channels-major within a sample). Returns samples per channel.
int decode_all(const std::vector<uint8_t>& d, const StreamInfo& si, std::vector<int32_t>& pcm, int64_t* n_out) {
  size_t pos = si.audio_start;
  std::vector<int32_t> ch[8];
  int64_t done = 0;
  pcm.clear();
  while (pos + 2 <= d.size()) {
    if (!(d[pos] == 0xff && (d[pos + 1] & 0xfe) == 0xf8)) break;   // no sync code: trailing data
    BitReader br(d.data() + pos, d.size() - pos);
    br.read(14);
    SK_REQUIRE(br.read(1) == 0, "flac: reserved bit set in frame header");
    br.read(1);  // blocking strategy
    const int bs_code = br.read(4), sr_code = br.read(4), ch_assign = br.read(4), ss_code = br.read(3);
    SK_REQUIRE(br.read(1) == 0, "flac: reserved bit set in frame header");
    int first = br.read(8);   // UTF-8 style coded frame / sample number
    if (first >= 0xc0) {
      int extra = 0;
      while (first & (0x40 >> extra)) ++extra;
      for (int i = 0; i <= extra; ++i) br.read(8);
    }
    int blocksize;
    if (bs_code == 1) blocksize = 192;
    else if (bs_code >= 2 && bs_code <= 5) blocksize = 576 << (bs_code - 2);
    else if (bs_code == 6) blocksize = br.read(8) + 1;
    else if (bs_code == 7) blocksize = br.read(16) + 1;
    else if (bs_code >= 8) blocksize = 256 << (bs_code - 8);
    else { SK_REQUIRE(false, "flac: reserved block size code"); }
    if (sr_code == 12) br.read(8);
    else if (sr_code == 13 || sr_code == 14) br.read(16);
    const size_t hdr_len = br.byte_pos();
    const uint8_t want8 = (uint8_t)br.read(8);
    SK_REQUIRE(crc8(d.data() + pos, hdr_len) == want8, "flac: frame header CRC-8 mismatch at byte %zu", pos);
    static const int ss_table[8] = {0, 8, 12, 0, 16, 20, 24, 0};
    const int bps = ss_code == 0 ? si.bps : ss_table[ss_code];
    SK_REQUIRE(bps > 0, "flac: reserved sample size code");
    const int nch = ch_assign < 8 ? ch_assign + 1 : 2;
    SK_REQUIRE(ch_assign <= 10 && nch == si.channels, "flac: channel layout changes mid-stream");
    for (int c = 0; c < nch; ++c) {
      ch[c].resize(blocksize);
      const int side = (ch_assign == 8 && c == 1) || (ch_assign == 9 && c == 0) || (ch_assign == 10 && c == 1);
      SK_TRY_RC2(decode_subframe(br, bps + side, blocksize, ch[c].data()));
    }
    br.align();
    const size_t body_len = br.byte_pos();
    const uint16_t want16 = (uint16_t)br.read(16);
    SK_REQUIRE(!br.eof, "flac: truncated frame at byte %zu", pos);
    SK_REQUIRE(crc16(d.data() + pos, body_len) == want16, "flac: frame CRC-16 mismatch at byte %zu", pos);
    if (ch_assign == 8) for (int i = 0; i < blocksize; ++i) ch[1][i] = ch[0][i] - ch[1][i];
    else if (ch_assign == 9) for (int i = 0; i < blocksize; ++i) ch[0][i] = ch[0][i] + ch[1][i];
    else if (ch_assign == 10)
      for (int i = 0; i < blocksize; ++i) {
        const int32_t side = ch[1][i];
        const int32_t mid = (int32_t)(((uint32_t)ch[0][i] << 1) | (side & 1));
        ch[0][i] = (mid + side) >> 1;
        ch[1][i] = (mid - side) >> 1;
      }
    const size_t base = pcm.size();
    pcm.resize(base + (size_t)blocksize * nch);
    for (int i = 0; i < blocksize; ++i)
      for (int c = 0; c < nch; ++c) pcm[base + (size_t)i * nch + c] = ch[c][i];
    done += blocksize;
    pos += br.byte_pos();
  }
  *n_out = done;
  return 0;
}

}  // namespace

extern "C" {

int sk_flac_info(const char* path, int32_t* sample_rate, int32_t* channels, int32_t* bits_per_sample,
                 int64_t* n_samples, uint8_t* md5_16) {
  SK_REQUIRE(path, "sk_flac_info: null path");
  std::vector<uint8_t> d;
  SK_TRY_RC2(read_file(path, d));
  StreamInfo si;
  SK_TRY_RC2(parse_header(d, si));
  if (sample_rate) *sample_rate = si.sample_rate;
  if (channels) *channels = si.channels;
  if (bits_per_sample) *bits_per_sample = si.bps;
  if (n_samples) *n_samples = si.total;
  if (md5_16) memcpy(md5_16, si.md5, 16);
  return 0;
}

// Interleaved int32 PCM, capacity in samples-per-channel; returns the decoded count through n_decoded.
int sk_flac_decode_i32(const char* path, int32_t* pcm_host, int64_t capacity, int64_t* n_decoded) {
  SK_REQUIRE(path && pcm_host && n_decoded, "sk_flac_decode_i32: null argument");
  std::vector<uint8_t> d;
  SK_TRY_RC2(read_file(path, d));
  StreamInfo si;
  SK_TRY_RC2(parse_header(d, si));
  std::vector<int32_t> pcm;
  int64_t n = 0;
  SK_TRY_RC2(decode_all(d, si, pcm, &n));
  SK_REQUIRE(n <= capacity, "sk_flac_decode_i32: buffer too small (%lld > %lld samples)", (long long)n, (long long)capacity);
  memcpy(pcm_host, pcm.data(), pcm.size() * sizeof(int32_t));
  *n_decoded = n;
  return 0;
}

}  // extern "C"
