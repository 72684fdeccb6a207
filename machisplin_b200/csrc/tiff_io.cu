// GeoTIFF in / out for the rasters either side of the hot path: what `terra::rast(path)` hands to
// machisplin.mltps() (README Example 1 / V73:68-70: the bundled alt / slope / TWI rasters are INT16, 128 x 128 tiles, uncompressed
// or LZW, GDAL_NODATA, ModelPixelScale + ModelTiepoint) and what `terra::writeRaster()` leaves behind (V73:1011, 1020: FLT4S).
// Host-only code (no device work, no mb_ctx): tiles / strips are decoded by all host threads straight into the caller's float32
// plane - pinned memory when it comes from the Python / R side of mb_mltps_predict - so that the H2D copy of DESIGN.md section 6
// can start on the rows that are ready.
//
// Supported: classic little- or big-endian TIFF, striped or tiled, chunky or planar bands, 8 / 16 / 32-bit integers and 32 / 64-bit
// floats, compression none (1), LZW (5: TIFF 6.0 code stream, MSB first, early change) and PackBits (32773), predictor 1 and 2.  Refused with a message: BigTIFF, Deflate / JPEG / ZSTD,
// floating-point predictor 3, palette / bilevel images.
#include "common.cuh"

#include <sys/mman.h>
#include <sys/stat.h>
#include <fcntl.h>
#include <unistd.h>

#include <atomic>
#include <cmath>
#include <cstdlib>
#include <thread>

namespace mb {
namespace {

struct Mapped {
  const uint8_t* p = nullptr;
  size_t n = 0;
  int fd = -1;
  explicit Mapped(const char* path) {
    fd = ::open(path, O_RDONLY);
    if (fd < 0) throw Error(MB_E_ARG, std::string("cannot open '") + path + "'");
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < 8) { ::close(fd); throw Error(MB_E_ARG, std::string("'") + path + "' is not a TIFF file"); }
    n = (size_t)st.st_size;
    void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m == MAP_FAILED) { ::close(fd); throw Error(MB_E_NOMEM, "mmap failed"); }
    p = static_cast<const uint8_t*>(m);
  }
  ~Mapped() { if (p) munmap(const_cast<uint8_t*>(p), n); if (fd >= 0) ::close(fd); }
  Mapped(const Mapped&) = delete;
  Mapped& operator=(const Mapped&) = delete;
};

struct Meta {
  bool le = true;
  int width = 0, height = 0, spp = 1, bits = 0, fmt = 1, compression = 1, predictor = 1, planar = 1, photometric = 1;
  bool tiled = false;
  int cw = 0, ch = 0;                       // chunk (tile or strip) size
  std::vector<uint64_t> off, cnt;
  bool has_scale = false, has_tie = false, has_nodata = false;
  double scale[3] = {1, 1, 0}, tie[6] = {0, 0, 0, 0, 0, 0}, nodata = 0;
  int epsg = 0;
};

struct Rd {
  const uint8_t* p; size_t n; bool le;
  void need(size_t o, size_t k) const { if (o > n || k > n - o) throw Error(MB_E_ARG, "TIFF: offset outside the file"); }
  uint16_t u16(size_t o) const { need(o, 2); return le ? (uint16_t)(p[o] | p[o + 1] << 8) : (uint16_t)(p[o] << 8 | p[o + 1]); }
  uint32_t u32(size_t o) const {
    need(o, 4);
    return le ? ((uint32_t)p[o] | (uint32_t)p[o + 1] << 8 | (uint32_t)p[o + 2] << 16 | (uint32_t)p[o + 3] << 24)
              : ((uint32_t)p[o] << 24 | (uint32_t)p[o + 1] << 16 | (uint32_t)p[o + 2] << 8 | (uint32_t)p[o + 3]);
  }
  double f64(size_t o) const {
    need(o, 8);
    uint64_t v = 0;
    for (int i = 0; i < 8; ++i) v |= (uint64_t)p[o + i] << (le ? 8 * i : 8 * (7 - i));
    double d;
    std::memcpy(&d, &v, 8);
    return d;
  }
};

constexpr int kTypeSize[13] = {0, 1, 1, 2, 4, 8, 1, 1, 2, 4, 8, 4, 8};

Meta parse(const Mapped& f) {
  Meta m;
  if (f.p[0] == 'I' && f.p[1] == 'I') m.le = true;
  else if (f.p[0] == 'M' && f.p[1] == 'M') m.le = false;
  else throw Error(MB_E_ARG, "not a TIFF file (byte-order mark)");
  Rd r{f.p, f.n, m.le};
  const uint16_t magic = r.u16(2);
  if (magic == 43) throw Error(MB_E_UNSUPPORTED, "BigTIFF is not supported");
  if (magic != 42) throw Error(MB_E_ARG, "not a TIFF file (magic number)");
  const size_t ifd = r.u32(4);
  const int nent = r.u16(ifd);
  int rows_per_strip = 0, tw = 0, th = 0;
  std::vector<uint64_t> soff, scnt, toff, tcnt;
  auto values = [&](size_t e, std::vector<double>& out) {        // entry -> numbers
    const int type = r.u16(e + 2);
    const uint32_t count = r.u32(e + 4);
    if (type < 1 || type > 12) return;
    const size_t bytes = (size_t)kTypeSize[type] * count;
    const size_t at = bytes <= 4 ? e + 8 : r.u32(e + 8);
    r.need(at, bytes);
    out.resize(count);
    for (uint32_t i = 0; i < count; ++i) {
      const size_t o = at + (size_t)i * kTypeSize[type];
      switch (type) {
        case 1: case 7: out[i] = f.p[o]; break;
        case 2: out[i] = f.p[o]; break;
        case 6: out[i] = (int8_t)f.p[o]; break;
        case 3: out[i] = r.u16(o); break;
        case 8: out[i] = (int16_t)r.u16(o); break;
        case 4: out[i] = r.u32(o); break;
        case 9: out[i] = (int32_t)r.u32(o); break;
        case 12: out[i] = r.f64(o); break;
        case 11: { uint32_t v = r.u32(o); float fl; std::memcpy(&fl, &v, 4); out[i] = fl; } break;
        case 5: out[i] = r.u32(o + 4) ? (double)r.u32(o) / r.u32(o + 4) : 0.0; break;
        case 10: out[i] = r.u32(o + 4) ? (double)(int32_t)r.u32(o) / (int32_t)r.u32(o + 4) : 0.0; break;
      }
    }
  };
  std::vector<double> v;
  std::vector<double> geokeys;
  for (int i = 0; i < nent; ++i) {
    const size_t e = ifd + 2 + (size_t)12 * i;
    const int tag = r.u16(e);
    v.clear();
    values(e, v);
    if (v.empty()) continue;
    auto u64s = [&](std::vector<uint64_t>& dst) { dst.resize(v.size()); for (size_t k = 0; k < v.size(); ++k) dst[k] = (uint64_t)v[k]; };
    switch (tag) {
      case 256: m.width = (int)v[0]; break;
      case 257: m.height = (int)v[0]; break;
      case 258:
        m.bits = (int)v[0];
        for (double b : v) if ((int)b != m.bits) throw Error(MB_E_UNSUPPORTED, "TIFF: bands of different bit depth");
        break;
      case 259: m.compression = (int)v[0]; break;
      case 262: m.photometric = (int)v[0]; break;
      case 273: u64s(soff); break;
      case 277: m.spp = (int)v[0]; break;
      case 278: rows_per_strip = v[0] > 2147483647.0 ? 0 : (int)v[0]; break;
      case 279: u64s(scnt); break;
      case 284: m.planar = (int)v[0]; break;
      case 317: m.predictor = (int)v[0]; break;
      case 322: tw = (int)v[0]; break;
      case 323: th = (int)v[0]; break;
      case 324: u64s(toff); break;
      case 325: u64s(tcnt); break;
      case 339: m.fmt = (int)v[0]; break;
      case 33550: if (v.size() >= 2) { m.has_scale = true; m.scale[0] = v[0]; m.scale[1] = v[1]; } break;
      case 33922: if (v.size() >= 6) { m.has_tie = true; for (int k = 0; k < 6; ++k) m.tie[k] = v[k]; } break;
      case 34735: geokeys = v; break;
      case 42113: {                          // GDAL_NODATA: ASCII number
        std::string s;
        for (double c : v) if (c > 0) s.push_back((char)c);
        char* end = nullptr;
        const double nd = std::strtod(s.c_str(), &end);
        if (end != s.c_str()) { m.has_nodata = true; m.nodata = nd; }
      } break;
      default: break;
    }
  }
  for (size_t k = 4; k + 3 < geokeys.size(); k += 4)               // GeoKeyDirectory: (key, location, count, value)
    if (((int)geokeys[k] == 2048 || (int)geokeys[k] == 3072) && (int)geokeys[k + 1] == 0) m.epsg = (int)geokeys[k + 3];
  MB_REQUIRE(m.width > 0 && m.height > 0, "TIFF: image size missing");
  MB_REQUIRE(m.spp >= 1 && m.spp <= 64, "TIFF: unsupported number of bands");
  if (m.fmt == 4 || m.fmt == 0) m.fmt = 1;
  if (!((m.fmt == 1 || m.fmt == 2) && (m.bits == 8 || m.bits == 16 || m.bits == 32)) && !(m.fmt == 3 && (m.bits == 32 || m.bits == 64)))
    throw Error(MB_E_UNSUPPORTED, "TIFF: unsupported sample type (" + std::to_string(m.bits) + " bits, format " + std::to_string(m.fmt) + ")");
  if (m.photometric == 3) throw Error(MB_E_UNSUPPORTED, "TIFF: palette images are not rasters of values");
  if (m.compression != 1 && m.compression != 5 && m.compression != 32773)
    throw Error(MB_E_UNSUPPORTED, "TIFF: compression scheme " + std::to_string(m.compression) + " is not supported (none, LZW, PackBits are)");
  if (m.predictor != 1 && m.predictor != 2)
    throw Error(MB_E_UNSUPPORTED, "TIFF: predictor " + std::to_string(m.predictor) + " is not supported (1 and 2 are)");
  MB_REQUIRE(m.planar == 1 || m.planar == 2, "TIFF: bad PlanarConfiguration");
  if (!toff.empty()) {
    MB_REQUIRE(tw > 0 && th > 0 && toff.size() == tcnt.size(), "TIFF: inconsistent tile tags");
    m.tiled = true; m.cw = tw; m.ch = th; m.off = toff; m.cnt = tcnt;
  } else {
    MB_REQUIRE(!soff.empty() && soff.size() == scnt.size(), "TIFF: no strips and no tiles");
    m.tiled = false; m.cw = m.width;
    m.ch = rows_per_strip > 0 ? std::min(rows_per_strip, m.height) : m.height;
    m.off = soff; m.cnt = scnt;
  }
  const size_t across = ((size_t)m.width + m.cw - 1) / m.cw, down = ((size_t)m.height + m.ch - 1) / m.ch;
  const size_t want = across * down * (m.planar == 2 ? m.spp : 1);
  MB_REQUIRE(m.off.size() >= want, "TIFF: fewer chunks than the image needs");
  for (size_t k = 0; k < want; ++k)
    if (m.off[k] > f.n || m.cnt[k] > f.n - m.off[k]) throw Error(MB_E_ARG, "TIFF: chunk outside the file");
  return m;
}

// ---- decompression -------------------------------------------------------------------------------------------------
// TIFF LZW: MSB-first codes of 9..12 bits, Clear = 256, EOI = 257, the width grows one code early.  Returns bytes written.
size_t lzw_decode(const uint8_t* src, size_t n, uint8_t* dst, size_t cap) {
  static thread_local std::vector<uint16_t> prefix(4096);
  static thread_local std::vector<uint8_t> suffix(4096);
  static thread_local std::vector<uint16_t> length(4096);
  static thread_local std::vector<uint8_t> spill;
  for (int i = 0; i < 256; ++i) { prefix[i] = 0; suffix[i] = (uint8_t)i; length[i] = 1; }
  size_t out = 0, ip = 0;
  uint32_t acc = 0;
  int nbits = 0, width = 9, next = 258, prev = -1;
  while (out < cap) {
    while (nbits < width && ip < n) { acc = (acc << 8) | src[ip++]; nbits += 8; }
    if (nbits < width) break;                                   // ran out of data without EOI: tolerated
    const int code = (int)((acc >> (nbits - width)) & ((1u << width) - 1));
    nbits -= width;
    if (code == 257) break;
    if (code == 256) { width = 9; next = 258; prev = -1; continue; }
    const bool kwkwk = prev >= 0 && code == next;               // string = str(prev) + first byte of str(prev)
    if ((prev < 0 && code > 255) || (!kwkwk && code >= next)) throw Error(MB_E_ARG, "TIFF: corrupt LZW stream");
    const int from = kwkwk ? prev : code;
    const size_t len = (size_t)length[from] + (kwkwk ? 1 : 0);
    uint8_t* w = dst + out;
    const bool fits = len <= cap - out;
    if (!fits) { spill.resize(len); w = spill.data(); }         // the last string may run past the chunk: keep what fits
    {
      size_t q = length[from];
      int c = from;
      while (q > 0) { w[--q] = suffix[c]; c = prefix[c]; }
      if (kwkwk) w[len - 1] = w[0];
    }
    const uint8_t first = w[0];
    if (fits) out += len;
    else { std::memcpy(dst + out, w, cap - out); out = cap; }
    if (prev >= 0 && next < 4096) {
      prefix[next] = (uint16_t)prev;
      suffix[next] = first;
      length[next] = (uint16_t)(length[prev] + 1);
      ++next;
      if (next + 1 >= (1 << width) && width < 12) ++width;      // "early change": one code before the table needs it
    }
    prev = code;
  }
  return out;
}

size_t packbits_decode(const uint8_t* src, size_t n, uint8_t* dst, size_t cap) {
  size_t ip = 0, out = 0;
  while (ip < n && out < cap) {
    const int8_t h = (int8_t)src[ip++];
    if (h >= 0) {
      const size_t k = std::min<size_t>((size_t)h + 1, std::min(n - ip, cap - out));
      std::memcpy(dst + out, src + ip, k);
      ip += (size_t)h + 1; out += k;
    } else if (h != -128) {
      if (ip >= n) break;
      const size_t k = std::min<size_t>((size_t)(1 - h), cap - out);
      std::memset(dst + out, src[ip++], k);
      out += k;
    }
  }
  return out;
}

// TIFF LZW encoder (writer side): same code stream as above, hash table keyed by (prefix << 8 | byte).
void lzw_encode(const uint8_t* src, size_t n, std::vector<uint8_t>& out) {
  constexpr int kHash = 1 << 14;
  static thread_local std::vector<int32_t> hkey(kHash), hval(kHash);
  uint32_t acc = 0;
  int nbits = 0, width = 9, next = 258;
  auto put = [&](int code) {
    acc = (acc << width) | (uint32_t)code;
    nbits += width;
    while (nbits >= 8) { out.push_back((uint8_t)(acc >> (nbits - 8))); nbits -= 8; }
  };
  auto reset = [&] { std::fill(hkey.begin(), hkey.end(), -1); next = 258; width = 9; };
  reset();
  put(256);
  if (n == 0) { put(257); if (nbits) out.push_back((uint8_t)(acc << (8 - nbits))); return; }
  int cur = src[0];
  for (size_t i = 1; i < n; ++i) {
    const int c = src[i];
    const int32_t key = (cur << 8) | c;
    uint32_t h = ((uint32_t)key * 2654435761u) >> 18;
    int found = -1;
    while (hkey[h] != -1) {
      if (hkey[h] == key) { found = hval[h]; break; }
      h = (h + 1) & (kHash - 1);
    }
    if (found >= 0) { cur = found; continue; }
    put(cur);
    hkey[h] = key; hval[h] = next++;
    if (next >= 4094) { put(256); reset(); }                    // table full (libtiff resets at CODE_MAX - 1)
    else if (next >= (1 << width) && width < 12) ++width;       // the decoder's table is one entry behind: its early change
    cur = c;
  }
  put(cur);
  // the decoder adds an entry for the last string as well: keep the width in step before EOI
  ++next;
  if (next >= (1 << width) && width < 12) ++width;
  put(257);
  if (nbits) out.push_back((uint8_t)(acc << (8 - nbits)));
}

template <class F>
void parallel_chunks(size_t n, int nthreads, F&& fn) {
  int nt = nthreads > 0 ? nthreads : (int)std::thread::hardware_concurrency();
  nt = (int)std::max<size_t>(1, std::min<size_t>((size_t)std::max(nt, 1), n));
  std::atomic<size_t> nexti{0};
  std::vector<std::exception_ptr> errs(nt);
  auto work = [&](int w) {
    try {
      for (;;) {
        const size_t i = nexti.fetch_add(1);
        if (i >= n) break;
        fn(i);
      }
    } catch (...) { errs[w] = std::current_exception(); }
  };
  std::vector<std::thread> thr;
  for (int w = 1; w < nt; ++w) thr.emplace_back(work, w);
  work(0);
  for (auto& t : thr) t.join();
  for (auto& e : errs) if (e) std::rethrow_exception(e);
}

inline uint64_t bswap(uint64_t v, int bytes) {
  uint64_t r = 0;
  for (int i = 0; i < bytes; ++i) r |= ((v >> (8 * i)) & 0xff) << (8 * (bytes - 1 - i));
  return r;
}

void read_band(const Mapped& f, const Meta& m, int band, float* out, int nthreads) {
  MB_REQUIRE(band >= 0 && band < m.spp, "TIFF: no such band");
  const int bytes = m.bits / 8;
  const int cspp = m.planar == 1 ? m.spp : 1;                   // samples per pixel inside one chunk
  const size_t across = ((size_t)m.width + m.cw - 1) / m.cw, down = ((size_t)m.height + m.ch - 1) / m.ch;
  const size_t plane0 = m.planar == 2 ? (size_t)band * across * down : 0;
  const int sel = m.planar == 1 ? band : 0;
  const bool host_le = true;
  parallel_chunks(across * down, nthreads, [&](size_t ci) {
    static thread_local std::vector<uint8_t> buf;
    const size_t cx = ci % across, cy = ci / across;
    const int rows = m.tiled ? m.ch : std::min(m.ch, m.height - (int)cy * m.ch);
    const size_t raw = (size_t)rows * m.cw * cspp * bytes;
    const uint8_t* src = f.p + m.off[plane0 + ci];
    const size_t nsrc = (size_t)m.cnt[plane0 + ci];
    const uint8_t* data = src;
    if (m.compression != 1) {
      buf.resize(raw);
      const size_t got = m.compression == 5 ? lzw_decode(src, nsrc, buf.data(), raw) : packbits_decode(src, nsrc, buf.data(), raw);
      if (got < raw) std::memset(buf.data() + got, 0, raw - got);
      data = buf.data();
    } else if (nsrc < raw) {
      buf.assign(raw, 0);
      std::memcpy(buf.data(), src, nsrc);
      data = buf.data();
    }
    const int r0 = (int)cy * m.ch, c0 = (int)cx * m.cw;
    const int nr = std::min(rows, m.height - r0), nc = std::min(m.cw, m.width - c0);
    for (int rr = 0; rr < nr; ++rr) {
      const uint8_t* line = data + (size_t)rr * m.cw * cspp * bytes;
      float* dst = out + (size_t)(r0 + rr) * m.width + c0;
      uint64_t carry[64];                                       // predictor 2: running sum per sample of the pixel
      for (int s = 0; s < cspp; ++s) carry[s] = 0;
      for (int cc = 0; cc < nc; ++cc) {
        uint64_t raw_v = 0;
        for (int s = 0; s < cspp; ++s) {
          if (m.predictor == 1 && s != sel) continue;
          const uint8_t* q = line + ((size_t)cc * cspp + s) * bytes;
          uint64_t v = 0;
          std::memcpy(&v, q, bytes);                            // host is little-endian
          if (m.le != host_le) v = bswap(v, bytes);
          if (m.predictor == 2) {
            carry[s] = (carry[s] + v) & (bytes == 8 ? ~0ull : ((1ull << (8 * bytes)) - 1));
            v = carry[s];
          }
          if (s == sel) raw_v = v;
        }
        double val;
        if (m.fmt == 3) {
          if (bytes == 4) { uint32_t u = (uint32_t)raw_v; float fl; std::memcpy(&fl, &u, 4); val = fl; }
          else { std::memcpy(&val, &raw_v, 8); }
        } else if (m.fmt == 2) {
          val = bytes == 1 ? (double)(int8_t)raw_v : bytes == 2 ? (double)(int16_t)raw_v : (double)(int32_t)raw_v;
        } else {
          val = (double)raw_v;
        }
        dst[cc] = (m.has_nodata && val == m.nodata) ? __builtin_nanf("") : (float)val;
      }
    }
  });
}

mb_grid grid_of(const Meta& m) {
  mb_grid g;
  g.nrow = m.height; g.ncol = m.width;
  if (m.has_scale && m.has_tie) {
    g.xmin = m.tie[3] - m.tie[0] * m.scale[0];
    g.ymax = m.tie[4] + m.tie[1] * m.scale[1];
    g.xmax = g.xmin + m.width * m.scale[0];
    g.ymin = g.ymax - m.height * m.scale[1];
  } else {
    g.xmin = 0; g.xmax = m.width; g.ymin = 0; g.ymax = m.height;
  }
  return g;
}

// ---- writer ----------------------------------------------------------------------------------------------------------
struct Out {
  std::vector<uint8_t> b;
  void u16(uint16_t v) { b.push_back((uint8_t)v); b.push_back((uint8_t)(v >> 8)); }
  void u32(uint32_t v) { for (int i = 0; i < 4; ++i) b.push_back((uint8_t)(v >> (8 * i))); }
  void f64(double d) { uint64_t v; std::memcpy(&v, &d, 8); for (int i = 0; i < 8; ++i) b.push_back((uint8_t)(v >> (8 * i))); }
};

template <class T>
void write_f32(const char* path, const mb_grid& g, const T* data, int compression, int epsg, int nthreads) {
  MB_REQUIRE(path && data, "NULL argument");
  check_grid(&g);
  MB_REQUIRE(compression == 1 || compression == 5, "compression must be 1 (none) or 5 (LZW)");
  constexpr int kTile = 256;
  const size_t across = ((size_t)g.ncol + kTile - 1) / kTile, down = ((size_t)g.nrow + kTile - 1) / kTile;
  const size_t ntile = across * down;
  std::vector<std::vector<uint8_t>> enc(ntile);
  parallel_chunks(ntile, nthreads, [&](size_t ti) {
    static thread_local std::vector<float> raw;
    raw.assign((size_t)kTile * kTile, __builtin_nanf(""));
    const int r0 = (int)(ti / across) * kTile, c0 = (int)(ti % across) * kTile;
    const int nr = std::min(kTile, g.nrow - r0), nc = std::min(kTile, g.ncol - c0);
    for (int rr = 0; rr < nr; ++rr)
      for (int cc = 0; cc < nc; ++cc) raw[(size_t)rr * kTile + cc] = (float)data[(size_t)(r0 + rr) * g.ncol + c0 + cc];
    const uint8_t* bytes = reinterpret_cast<const uint8_t*>(raw.data());
    const size_t nb = raw.size() * 4;
    if (compression == 1) enc[ti].assign(bytes, bytes + nb);
    else { enc[ti].reserve(nb / 2); lzw_encode(bytes, nb, enc[ti]); }
  });
  // layout: header | tile data | offsets | counts | doubles | geokeys | nodata | IFD
  uint64_t pos = 8;
  std::vector<uint32_t> toff(ntile), tcnt(ntile);
  for (size_t i = 0; i < ntile; ++i) {
    toff[i] = (uint32_t)pos; tcnt[i] = (uint32_t)enc[i].size();
    pos += enc[i].size();
    if (pos & 1) ++pos;
    if (pos > 0xfffffff0ull) throw Error(MB_E_UNSUPPORTED, "raster too large for a classic TIFF file (4 GB)");
  }
  Out tail;
  const uint32_t off_toff = (uint32_t)pos;
  for (uint32_t v : toff) tail.u32(v);
  const uint32_t off_tcnt = off_toff + 4 * (uint32_t)ntile;
  for (uint32_t v : tcnt) tail.u32(v);
  const uint32_t off_scale = off_tcnt + 4 * (uint32_t)ntile;
  const double rx = (g.xmax - g.xmin) / g.ncol, ry = (g.ymax - g.ymin) / g.nrow;
  tail.f64(rx); tail.f64(ry); tail.f64(0.0);
  const uint32_t off_tie = off_scale + 24;
  tail.f64(0); tail.f64(0); tail.f64(0); tail.f64(g.xmin); tail.f64(g.ymax); tail.f64(0);
  const uint32_t off_keys = off_tie + 48;
  std::vector<uint16_t> keys = {1, 1, 0, 0, 1025, 0, 1, 1};            // GTRasterTypeGeoKey = RasterPixelIsArea
  if (epsg > 0) {
    const bool geographic = (epsg >= 4000 && epsg < 5000);
    keys.insert(keys.begin() + 4, {1024, 0, 1, (uint16_t)(geographic ? 2 : 1)});
    keys.insert(keys.end(), {(uint16_t)(geographic ? 2048 : 3072), 0, 1, (uint16_t)epsg});
  }
  keys[3] = (uint16_t)(keys.size() / 4 - 1);
  for (uint16_t k : keys) tail.u16(k);
  const uint32_t off_nodata = off_keys + 2 * (uint32_t)keys.size();
  const char nd[4] = {'n', 'a', 'n', 0};
  for (char c : nd) tail.b.push_back((uint8_t)c);
  const uint32_t off_ifd = off_nodata + 4;
  struct Ent { uint16_t tag, type; uint32_t count, value; };
  const uint32_t one_tile = ntile == 1;
  std::vector<Ent> ents = {
      {256, 4, 1, (uint32_t)g.ncol}, {257, 4, 1, (uint32_t)g.nrow}, {258, 3, 1, 32}, {259, 3, 1, (uint32_t)compression},
      {262, 3, 1, 1}, {277, 3, 1, 1}, {284, 3, 1, 1}, {322, 3, 1, kTile}, {323, 3, 1, kTile},
      {324, 4, (uint32_t)ntile, one_tile ? toff[0] : off_toff}, {325, 4, (uint32_t)ntile, one_tile ? tcnt[0] : off_tcnt},
      {339, 3, 1, 3}, {33550, 12, 3, off_scale}, {33922, 12, 6, off_tie}, {34735, 3, (uint32_t)keys.size(), off_keys},
      {42113, 2, 4, off_nodata}};
  Out ifd;
  ifd.u16((uint16_t)ents.size());
  for (const Ent& e : ents) { ifd.u16(e.tag); ifd.u16(e.type); ifd.u32(e.count); ifd.u32(e.value); }
  ifd.u32(0);
  FILE* fp = std::fopen(path, "wb");
  if (!fp) throw Error(MB_E_ARG, std::string("cannot create '") + path + "'");
  bool ok = true;
  Out head;
  head.b = {'I', 'I', 42, 0};
  head.u32(off_ifd);
  ok &= std::fwrite(head.b.data(), 1, 8, fp) == 8;
  uint64_t at = 8;
  for (size_t i = 0; i < ntile && ok; ++i) {
    ok &= std::fwrite(enc[i].data(), 1, enc[i].size(), fp) == enc[i].size();
    at += enc[i].size();
    if (at & 1) { ok &= std::fputc(0, fp) != EOF; ++at; }
  }
  ok &= std::fwrite(tail.b.data(), 1, tail.b.size(), fp) == tail.b.size();
  ok &= std::fwrite(ifd.b.data(), 1, ifd.b.size(), fp) == ifd.b.size();
  ok &= std::fclose(fp) == 0;
  if (!ok) throw Error(MB_E_ARG, std::string("write to '") + path + "' failed");
}


// =========================================================================================================================
// Device path: mb_tiff_read_f32_dev.  The file's COMPRESSED bytes cross PCIe (through a pair of pinned bounce buffers), the
// GPU undoes LZW, the horizontal predictor, the sample type and the NoData value, and the float32 plane is born in HBM - where
// mb_mltps_predict_dev wants it.  One warp per tile / strip:
//   k_tiff_lzw     lane 0 walks the code stream; the dictionary holds, per code, WHERE in the already decoded output its string
//                  starts and how long it is (every LZW string is a copy of earlier output + 1 byte), 24 KB of shared memory; strings
//                  of 16 bytes or more are copied by the whole warp.  Same early-change rule and tolerance of a missing
//                  EOI as the host decoder.
//   k_tiff_unpack  rows of the decoded chunk -> output plane, coalesced; predictor 2 as a warp-wide inclusive scan per 32 pixels
//                  with a running carry (modulo 2^bits, like libtiff).
// Big-endian files, PackBits and predictor 2 on pixel-interleaved multi-band chunks take the host decoder + one upload.
// =========================================================================================================================
struct ChunkJob {
  unsigned long long src;      // offset of the chunk's bytes in the compressed device buffer
  unsigned int nsrc;           // their number
  int r0, c0, nr, nc;          // where the chunk's pixels go: rows [r0, r0 + nr), columns [c0, c0 + nc)
  int rows;                    // rows held by the chunk buffer (tile height, or the strip's real height)
};
struct UnpackArgs {
  const unsigned char* raw;    // decoded (or uncompressed) chunk bytes: chunk j at raw_base[j]
  const unsigned long long* raw_off;
  const ChunkJob* jobs;
  int cw, cspp, sel, bytes, fmt, predictor, width;
  int has_nodata; double nodata;
  float* out;
};

__global__ void __launch_bounds__(32) k_tiff_lzw(const unsigned char* __restrict__ comp, const ChunkJob* __restrict__ jobs,
                                                   const unsigned long long* __restrict__ raw_off, unsigned char* __restrict__ raw,
                                                   const unsigned long long* __restrict__ raw_cap, int* __restrict__ err) {
  __shared__ unsigned int s_pos[4096];
  __shared__ unsigned short s_len[4096];
  const ChunkJob jb = jobs[blockIdx.x];
  const unsigned char* src = comp + jb.src;
  unsigned char* dst = raw + raw_off[blockIdx.x];
  const size_t cap = (size_t)raw_cap[blockIdx.x];
  const size_t n = jb.nsrc;
  const int lane = threadIdx.x;
  // lane 0 decodes; the other lanes wait for copy orders (len >= 16) or the end
  size_t out = 0, ip = 0;
  unsigned int acc = 0;
  int nbits = 0, width = 9, next = 258, prev = -1;
  size_t prev_pos = 0;
  int prev_len = 0;
  for (;;) {
    // ---- lane 0: next code -> (from, len, kwkwk) ----
    unsigned int cmd_from = 0; int cmd_len = -1;     // -1 = stop, 0 = nothing to copy (clear code)
    int kw = 0;
    if (lane == 0) {
      if (out < cap) {
        while (nbits < width && ip < n) { acc = (acc << 8) | src[ip++]; nbits += 8; }
        if (nbits >= width) {
          const int code = (int)((acc >> (nbits - width)) & ((1u << width) - 1));
          nbits -= width;
          if (code == 257) cmd_len = -1;
          else if (code == 256) { width = 9; next = 258; prev = -1; cmd_len = 0; }
          else {
            const bool kwkwk = prev >= 0 && code == next;
            if ((prev < 0 && code > 255) || (!kwkwk && code >= next)) { *err = 1; cmd_len = -1; }
            else {
              size_t pos; int len;
              if (kwkwk) { pos = prev_pos; len = prev_len + 1; kw = 1; }
              else if (code < 256) { pos = (size_t)-1; len = 1; }
              else { pos = s_pos[code]; len = s_len[code]; }
              // new dictionary entry: str(prev) + first byte of this string = the prev_len + 1 bytes starting at prev_pos
              if (prev >= 0 && next < 4096) {
                s_pos[next] = (unsigned int)prev_pos;
                s_len[next] = (unsigned short)(prev_len + 1);
                ++next;
                if (next + 1 >= (1 << width) && width < 12) ++width;
              }
              const size_t room = cap - out;
              const int k = (size_t)len <= room ? len : (int)room;
              if (code < 256 && !kwkwk) { dst[out] = (unsigned char)code; cmd_len = 0; }
              else if (k < 16) { for (int i = 0; i < k; ++i) dst[out + i] = dst[pos + i]; cmd_len = 0; }   // forward copy: overlap is the point (kwkwk)
              else { cmd_from = (unsigned int)pos; cmd_len = k; }
              prev = code; prev_pos = out; prev_len = len;
              if (cmd_len == 0) out += k;
            }
          }
        }
      }
    }
    cmd_len = __shfl_sync(0xffffffffu, cmd_len, 0);
    if (cmd_len < 0) break;
    if (cmd_len > 0) {
      cmd_from = __shfl_sync(0xffffffffu, cmd_from, 0);
      const unsigned long long o = __shfl_sync(0xffffffffu, (unsigned long long)out, 0);
      kw = __shfl_sync(0xffffffffu, kw, 0);
      // source [from, from + len) and destination [o, o + len) overlap only in the kwkwk case (from + len - 1 == o): the last byte
      // equals the first, which is old output
      for (int i = lane; i < cmd_len; i += 32) {
        const size_t sidx = (kw && i == cmd_len - 1 && (size_t)cmd_from + i == o) ? (size_t)cmd_from : (size_t)cmd_from + i;
        dst[o + i] = dst[sidx];
      }
      __syncwarp();
      if (lane == 0) out += cmd_len;
    }
  }
  // zero what the stream did not cover (truncated chunk), like the host path
  out = __shfl_sync(0xffffffffu, (unsigned long long)out, 0);
  for (size_t i = out + lane; i < cap; i += 32) dst[i] = 0;
}

__device__ __forceinline__ unsigned long long tiff_load(const unsigned char* q, int bytes) {
  unsigned long long v = 0;
  for (int i = 0; i < bytes; ++i) v |= (unsigned long long)q[i] << (8 * i);      // little-endian samples
  return v;
}
__device__ __forceinline__ float tiff_value(unsigned long long raw_v, int bytes, int fmt, int has_nodata, double nodata) {
  double val;
  if (fmt == 3) val = bytes == 4 ? (double)__uint_as_float((unsigned int)raw_v) : __longlong_as_double((long long)raw_v);
  else if (fmt == 2) val = bytes == 1 ? (double)(signed char)raw_v : bytes == 2 ? (double)(short)raw_v : (double)(int)raw_v;
  else val = (double)raw_v;
  return (has_nodata && val == nodata) ? __int_as_float(0x7fc00000) : (float)val;
}

__global__ void __launch_bounds__(256) k_tiff_unpack(UnpackArgs a) {
  const ChunkJob jb = a.jobs[blockIdx.x];
  const unsigned char* data = a.raw + a.raw_off[blockIdx.x];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t pix = (size_t)a.cspp * a.bytes;
  const unsigned long long mask = a.bytes == 8 ? ~0ull : ((1ull << (8 * a.bytes)) - 1ull);
  for (int rr = warp; rr < jb.nr; rr += 8) {
    const unsigned char* line = data + (size_t)rr * a.cw * pix;
    float* dst = a.out + (size_t)(jb.r0 + rr) * a.width + jb.c0;
    unsigned long long carry = 0;
    for (int c0 = 0; c0 < jb.nc; c0 += 32) {
      const int cc = c0 + lane;
      unsigned long long v = cc < jb.nc ? tiff_load(line + (size_t)cc * pix + (size_t)a.sel * a.bytes, a.bytes) : 0ull;
      if (a.predictor == 2) {                       // cspp == 1 here (host side checks): inclusive scan over the row
        for (int d = 1; d < 32; d <<= 1) {
          const unsigned long long u = __shfl_up_sync(0xffffffffu, v, d);
          if (lane >= d) v += u;
        }
        v = (v + carry) & mask;
        carry = __shfl_sync(0xffffffffu, v, 31);
      }
      if (cc < jb.nc) dst[cc] = tiff_value(v, a.bytes, a.fmt, a.has_nodata, a.nodata);
    }
  }
}

// pinned bounce buffers, one pair per process (the reader is synchronous on the host side)
struct Bounce {
  unsigned char* p[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  size_t cap = 0;
  void ensure(size_t want) {
    if (cap >= want) return;
    for (int i = 0; i < 2; ++i) {
      if (p[i]) cudaFreeHost(p[i]);
      p[i] = nullptr;
      if (cudaMallocHost(reinterpret_cast<void**>(&p[i]), want) != cudaSuccess) { cap = 0; throw Error(MB_E_NOMEM, "cudaMallocHost (GeoTIFF bounce buffer)"); }
      if (!ev[i] && cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess) throw Error(MB_E_CUDA, "cudaEventCreate");
    }
    cap = want;
  }
};
Bounce& bounce() { static Bounce b; return b; }

void read_band_dev(mb_ctx* ctx, const Mapped& f, const Meta& m, int band, float* out_dev, int nthreads, cudaStream_t st,
                   mb_tiff_dev_stats* stats) {
  MB_REQUIRE(band >= 0 && band < m.spp, "TIFF: no such band");
  const int bytes = m.bits / 8;
  const int cspp = m.planar == 1 ? m.spp : 1;
  const size_t across = ((size_t)m.width + m.cw - 1) / m.cw, down = ((size_t)m.height + m.ch - 1) / m.ch;
  const size_t plane0 = m.planar == 2 ? (size_t)band * across * down : 0;
  const int sel = m.planar == 1 ? band : 0;
  const size_t nchunk = across * down;
  bool gpu_ok = m.le && (m.compression == 1 || m.compression == 5) && !(m.predictor == 2 && cspp > 1) && nchunk > 0;
  if (gpu_ok && m.compression == 1)               // a truncated uncompressed chunk is zero-padded by the host decoder only
    for (size_t ci = 0; ci < nchunk; ++ci) {
      const int rows = m.tiled ? m.ch : std::min(m.ch, m.height - (int)(ci / across) * m.ch);
      if (m.cnt[plane0 + ci] < (uint64_t)rows * m.cw * cspp * bytes) gpu_ok = false;
    }
  if (stats) std::memset(stats, 0, sizeof *stats);
  Arena& ar = ctx->arena;
  if (!gpu_ok) {
    // host decoder into a pinned plane, one upload
    const size_t ncell = (size_t)m.width * m.height;
    Bounce& b = bounce();
    b.ensure(std::max<size_t>(ncell * sizeof(float), size_t(32) << 20));
    read_band(f, m, band, reinterpret_cast<float*>(b.p[0]), nthreads);
    MB_CUDA(cudaMemcpyAsync(out_dev, b.p[0], ncell * sizeof(float), cudaMemcpyHostToDevice, st));
    MB_CUDA(cudaStreamSynchronize(st));
    if (stats) { stats->decoded_on_gpu = 0; stats->h2d_bytes = (int64_t)(ncell * sizeof(float)); stats->chunks = (int32_t)nchunk; }
    return;
  }
  // ---- jobs -----------------------------------------------------------------------------------------------------------
  std::vector<ChunkJob> jobs(nchunk);
  std::vector<unsigned long long> raw_off(nchunk), raw_cap(nchunk);
  size_t comp_total = 0, raw_total = 0;
  for (size_t ci = 0; ci < nchunk; ++ci) {
    const size_t cx = ci % across, cy = ci / across;
    const int rows = m.tiled ? m.ch : std::min(m.ch, m.height - (int)cy * m.ch);
    const size_t raw = (size_t)rows * m.cw * cspp * bytes;
    ChunkJob& j = jobs[ci];
    j.src = comp_total;
    j.nsrc = (unsigned int)m.cnt[plane0 + ci];
    j.r0 = (int)cy * m.ch; j.c0 = (int)cx * m.cw;
    j.nr = std::min(rows, m.height - j.r0); j.nc = std::min(m.cw, m.width - j.c0);
    j.rows = rows;
    comp_total += (j.nsrc + 15u) & ~size_t(15);
    raw_cap[ci] = raw;
    if (m.compression == 1) {
      // the chunk's own bytes are the raw bytes; a short chunk is padded with zeros on the device
      raw_off[ci] = raw_total;
      raw_total += (raw + 15) & ~size_t(15);
    } else {
      raw_off[ci] = raw_total;
      raw_total += (raw + 15) & ~size_t(15);
    }
  }
  unsigned char* d_comp = ar.take_n<unsigned char>(std::max<size_t>(comp_total, 16));
  unsigned char* d_raw = m.compression == 1 ? nullptr : ar.take_n<unsigned char>(std::max<size_t>(raw_total, 16));
  // ---- compressed bytes -> device through the pinned pair ---------------------------------------------------------------------
  Bounce& b = bounce();
  b.ensure(size_t(32) << 20);
  {
    size_t ci = 0;
    int which = 0;
    bool used[2] = {false, false};
    while (ci < nchunk) {
      size_t first = ci, fill = 0;
      while (ci < nchunk && fill + ((jobs[ci].nsrc + 15u) & ~size_t(15)) <= b.cap) { fill += (jobs[ci].nsrc + 15u) & ~size_t(15); ++ci; }
      MB_REQUIRE(ci > first, "TIFF: a chunk is larger than the bounce buffer (32 MB)");
      if (used[which]) MB_CUDA(cudaEventSynchronize(b.ev[which]));
      unsigned char* hp = b.p[which];
      const size_t base = (size_t)jobs[first].src;
      parallel_chunks(ci - first, nthreads, [&](size_t k) {
        const ChunkJob& j = jobs[first + k];
        std::memcpy(hp + (j.src - base), f.p + m.off[plane0 + first + k], j.nsrc);
      });
      MB_CUDA(cudaMemcpyAsync(d_comp + base, hp, fill, cudaMemcpyHostToDevice, st));
      MB_CUDA(cudaEventRecord(b.ev[which], st));
      used[which] = true;
      which ^= 1;
    }
  }
  ChunkJob* d_jobs = ar.upload(jobs.data(), jobs.size(), st);
  unsigned long long* d_cap = ar.upload(raw_cap.data(), raw_cap.size(), st);
  int* d_err = ar.take_n<int>(1);
  MB_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), st));
  const unsigned char* raw_src = d_comp;
  std::vector<unsigned long long> src_off(nchunk);
  unsigned long long* d_roff;
  if (m.compression == 5) {
    d_roff = ar.upload(raw_off.data(), raw_off.size(), st);
    MB_LAUNCH(ctx, "k_tiff_lzw", st) k_tiff_lzw<<<(unsigned)nchunk, 32, 0, st>>>(d_comp, d_jobs, d_roff, d_raw, d_cap, d_err);
    raw_src = d_raw;
  } else {
    // uncompressed: unpack straight from the uploaded bytes
    for (size_t ci = 0; ci < nchunk; ++ci) src_off[ci] = jobs[ci].src;
    d_roff = ar.upload(src_off.data(), src_off.size(), st);
  }
  UnpackArgs ua{raw_src, d_roff, d_jobs, m.cw, cspp, sel, bytes, m.fmt, m.predictor, m.width, m.has_nodata ? 1 : 0, m.nodata, out_dev};
  MB_LAUNCH(ctx, "k_tiff_unpack", st) k_tiff_unpack<<<(unsigned)nchunk, 256, 0, st>>>(ua);
  MB_CUDA(cudaGetLastError());
  int herr = 0;
  MB_CUDA(cudaMemcpyAsync(&herr, d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
  MB_CUDA(cudaStreamSynchronize(st));      // the job tables are host vectors; the bounce buffers are reused by the next call
  if (herr) throw Error(MB_E_ARG, "TIFF: corrupt LZW stream");
  if (stats) { stats->decoded_on_gpu = 1; stats->h2d_bytes = (int64_t)comp_total; stats->chunks = (int32_t)nchunk; }
}

}  // namespace
}  // namespace mb

using namespace mb;

extern "C" {

int mb_tiff_info(const char* path, mb_tiff_meta* out) {
  return guarded([&] {
    MB_REQUIRE(path && out, "NULL argument");
    Mapped f(path);
    const Meta m = parse(f);
    std::memset(out, 0, sizeof(*out));
    out->grid = grid_of(m);
    out->nbands = m.spp;
    out->bits = m.bits;
    out->sample_format = m.fmt;
    out->compression = m.compression;
    out->predictor = m.predictor;
    out->tiled = m.tiled ? 1 : 0;
    out->chunk_w = m.cw; out->chunk_h = m.ch;
    out->has_georef = (m.has_scale && m.has_tie) ? 1 : 0;
    out->has_nodata = m.has_nodata ? 1 : 0;
    out->nodata = m.nodata;
    out->epsg = m.epsg;
  });
}

int mb_tiff_read_f32(const char* path, int band, float* out, int nthreads) {
  return guarded([&] {
    MB_REQUIRE(path && out, "NULL argument");
    Mapped f(path);
    const Meta m = parse(f);
    read_band(f, m, band, out, nthreads);
  });
}

int mb_tiff_read_f32_dev(mb_ctx* ctx, const char* path, int band, float* out_dev, int nthreads, void* stream, mb_tiff_dev_stats* stats) {
  return guarded([&] {
    MB_REQUIRE(ctx && path && out_dev, "NULL argument");
    MB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    ctx->arena.begin(st);
    Mapped f(path);
    const Meta m = parse(f);
    read_band_dev(ctx, f, m, band, out_dev, nthreads, st, stats);
  });
}

int mb_tiff_write_f32(const char* path, const mb_grid* g, const float* data, int compression, int epsg, int nthreads) {
  return guarded([&] {
    MB_REQUIRE(g, "grid is NULL");
    write_f32(path, *g, data, compression, epsg, nthreads);
  });
}

int mb_tiff_write_f64(const char* path, const mb_grid* g, const double* data, int compression, int epsg, int nthreads) {
  return guarded([&] {
    MB_REQUIRE(g, "grid is NULL");
    write_f32(path, *g, data, compression, epsg, nthreads);
  });
}

}  // extern "C"
