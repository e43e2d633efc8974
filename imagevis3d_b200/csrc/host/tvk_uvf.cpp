// UVF container reader / writer behind include/tvk_uvf.h (host code; see that header for the
// reference functions each part replaces).  File layout (little endian), as the reference writes it:
//   "UVF-DATA" u8 bigEndian u64 version(5) u64 checksumSemantics(2 = MD5) u64 checksumLen(16)
//   u8[16] md5(file[49..]) u64 offsetToFirstBlock(0)                       -- 57 bytes
//   blocks: u64 idLen, id, u64 semantics, u64 compression(0), u64 offsetToNextBlock (0 = last), payload
//   TOC payload = ExtendedOctree: u32 type, u64 components, u8 normals, u64 size[3], f64 aspect[3],
//     u64 brick[3], u32 overlap, u32 version(2), u64 totalSize, u32 compressionLevel,
//     ToC (36 B per brick: u64 offset-from-octree-start, u64 length, u32 compression, u64 validLength,
//     u32 atlasW, u32 atlasH), brick payloads in ToC order
#include "tvk_uvf.h"
#include "tvk_cuda.h"

#include <zlib.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

namespace {

// ---------------------------------------------------------------------------- MD5 (RFC 1321)
struct Md5 {
  uint32_t a = 0x67452301u, b = 0xefcdab89u, c = 0x98badcfeu, d = 0x10325476u;
  uint64_t total = 0;
  uint8_t buf[64];
  size_t fill = 0;

  static uint32_t rol(uint32_t x, int s) { return (x << s) | (x >> (32 - s)); }
  void block(const uint8_t* p) {
    static const uint32_t K[64] = {
        0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501, 0x698098d8, 0x8b44f7af,
        0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821, 0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa,
        0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8, 0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8,
        0x676f02d9, 0x8d2a4c8a, 0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70,
        0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665, 0xf4292244, 0x432aff97,
        0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1, 0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1,
        0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
    static const int S[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9, 14, 20, 5, 9, 14, 20, 5, 9, 14, 20, 5, 9, 14, 20,
                              4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};
    uint32_t m[16];
    memcpy(m, p, 64);
    uint32_t A = a, B = b, C = c, D = d;
    for (int i = 0; i < 64; ++i) {
      uint32_t f;
      int g;
      if (i < 16) { f = (B & C) | (~B & D); g = i; }
      else if (i < 32) { f = (D & B) | (~D & C); g = (5 * i + 1) & 15; }
      else if (i < 48) { f = B ^ C ^ D; g = (3 * i + 5) & 15; }
      else { f = C ^ (B | ~D); g = (7 * i) & 15; }
      const uint32_t t = D;
      D = C; C = B;
      B = B + rol(A + f + K[i] + m[g], S[i]);
      A = t;
    }
    a += A; b += B; c += C; d += D;
  }
  void update(const void* data, size_t n) {
    const uint8_t* p = static_cast<const uint8_t*>(data);
    total += n;
    if (fill) {
      const size_t k = std::min(n, 64 - fill);
      memcpy(buf + fill, p, k);
      fill += k; p += k; n -= k;
      if (fill == 64) { block(buf); fill = 0; }
    }
    for (; n >= 64; p += 64, n -= 64) block(p);
    if (n) { memcpy(buf, p, n); fill = n; }
  }
  void final(uint8_t out[16]) {
    const uint64_t bits = total * 8;
    const uint8_t pad = 0x80;
    update(&pad, 1);
    const uint8_t zero = 0;
    while (fill != 56) update(&zero, 1);
    update(&bits, 8);
    memcpy(out, &a, 4); memcpy(out + 4, &b, 4); memcpy(out + 8, &c, 4); memcpy(out + 12, &d, 4);
  }
};

void set_err(char* err, size_t n, const std::string& m) {
  if (err && n) { strncpy(err, m.c_str(), n - 1); err[n - 1] = 0; }
}

struct Buf {                       // little-endian serialiser
  std::vector<uint8_t> v;
  template <typename T> void put(T x) { const size_t o = v.size(); v.resize(o + sizeof(T)); memcpy(&v[o], &x, sizeof(T)); }
  void str(const std::string& s) { v.insert(v.end(), s.begin(), s.end()); }
  void raw(const void* p, size_t n) { const uint8_t* b = static_cast<const uint8_t*>(p); v.insert(v.end(), b, b + n); }
};

void block_header(Buf& b, const std::string& id, uint64_t semantics, uint64_t payload_bytes, bool last) {
  b.put<uint64_t>(id.size());
  b.str(id);
  b.put<uint64_t>(semantics);
  b.put<uint64_t>(0);                                                // COS_NONE
  b.put<uint64_t>(last ? 0 : id.size() + 32 + payload_bytes);       // DataBlock::GetOffsetToNextBlock
}

enum { BS_REG_NDIM_GRID = 1, BS_KEY_VALUE = 4, BS_HIST1D = 5, BS_HIST2D = 6, BS_MAXMIN = 7, BS_GEOMETRY = 8, BS_TOC = 9 };

bool md5_of_file_tail(FILE* f, uint64_t from, uint8_t digest[16]) {
  if (fseeko(f, (off_t)from, SEEK_SET)) return false;
  Md5 m;
  std::vector<uint8_t> chunk(1u << 24);
  for (;;) {
    const size_t n = fread(chunk.data(), 1, chunk.size(), f);
    if (n) m.update(chunk.data(), n);
    if (n < chunk.size()) break;
  }
  m.final(digest);
  return !ferror(f);
}

// zCompress (ZlibCompression.cpp): one deflate(Z_FINISH) into a buffer of the uncompressed size;
// returns the uncompressed size when the stream does not fit (the caller then stores the brick raw)
size_t z_compress(const uint8_t* src, size_t n, uint8_t* dst, uint32_t level) {
  z_stream s;
  memset(&s, 0, sizeof(s));
  if (deflateInit(&s, (int)std::min<uint32_t>(level, 9)) != Z_OK) return n;
  s.avail_in = (uInt)n; s.next_in = const_cast<Bytef*>(src);
  s.avail_out = (uInt)n; s.next_out = dst;
  const int ret = deflate(&s, Z_FINISH);
  const size_t out = n - s.avail_out;
  deflateEnd(&s);
  return ret == Z_STREAM_END ? out : n;
}

bool z_decompress(const uint8_t* src, size_t srclen, uint8_t* dst, size_t n) {
  if (srclen > 0xFFFFFFFFull || n > 0xFFFFFFFFull) return false;       // zlib's one-shot interface counts in 32 bits
  z_stream s;
  memset(&s, 0, sizeof(s));
  if (inflateInit(&s) != Z_OK) return false;
  s.avail_in = (uInt)srclen; s.next_in = const_cast<Bytef*>(src);
  s.avail_out = (uInt)n; s.next_out = dst;
  const int ret = inflate(&s, Z_FINISH);
  const bool ok = ret == Z_STREAM_END && s.avail_out == 0;
  inflateEnd(&s);
  return ok;
}

// bzip2 and LZMA bricks (ExtendedOctree.cpp:344-353 -> BzlibCompression.cpp:60-77, LzmaCompression.cpp:130-150): the
// reference links its vendored bzip2 / LZMA SDK; here the system's libbz2 / liblzma are opened on first use (no
// headers in the image, so the few public declarations needed are restated below).
bool bz_decompress(const uint8_t* src, size_t srclen, uint8_t* dst, size_t n) {
  if (srclen > 0xFFFFFFFFull || n > 0xFFFFFFFFull) return false;       // unsigned int lengths in the bzip2 API
  typedef int (*fn_t)(char*, unsigned int*, char*, unsigned int, int, int);      // BZ2_bzBuffToBuffDecompress
  static fn_t fn = [] {
    void* h = dlopen("libbz2.so.1.0", RTLD_NOW);
    if (!h) h = dlopen("libbz2.so.1", RTLD_NOW);
    return h ? reinterpret_cast<fn_t>(dlsym(h, "BZ2_bzBuffToBuffDecompress")) : nullptr;
  }();
  if (!fn) return false;
  unsigned int out = (unsigned int)n;
  const int rc = fn(reinterpret_cast<char*>(dst), &out, reinterpret_cast<char*>(const_cast<uint8_t*>(src)), (unsigned int)srclen, 0, 0);
  return rc == 0 && out == n;
}

// liblzma public ABI (lzma/base.h, lzma/filter.h, lzma/lzma12.h)
struct xz_stream {
  const uint8_t* next_in; size_t avail_in; uint64_t total_in;
  uint8_t* next_out; size_t avail_out; uint64_t total_out;
  const void* allocator; void* internal;
  void* reserved_ptr[4]; uint64_t reserved_int12[2]; size_t reserved_int34[2]; int reserved_enum[2];
};
struct xz_filter { uint64_t id; void* options; };
struct xz_options_lzma {
  uint32_t dict_size; const uint8_t* preset_dict; uint32_t preset_dict_size;
  uint32_t lc, lp, pb; int mode; uint32_t nice_len; int mf; uint32_t depth;
  uint8_t reserved[128];                                  // ext_* and reserved members of newer versions, zero
};
// A raw LZMA1 stream without end marker, as LzmaEncode(writeEndMark = 0) writes it.  The reader takes the
// properties from the compression level in the header (ExtendedOctree.cpp:54): lc 3, lp 0, pb 2 for every level
// (LzmaEncProps_Normalize); the dictionary only has to cover the brick.
bool lzma_decompress(const uint8_t* src, size_t srclen, uint8_t* dst, size_t n) {
  typedef int (*dec_t)(xz_stream*, const xz_filter*);
  typedef int (*code_t)(xz_stream*, int);
  typedef void (*end_t)(xz_stream*);
  static dec_t dec; static code_t code; static end_t end;
  static bool ok = [] {
    void* h = dlopen("liblzma.so.5", RTLD_NOW);
    if (!h) return false;
    dec = reinterpret_cast<dec_t>(dlsym(h, "lzma_raw_decoder"));
    code = reinterpret_cast<code_t>(dlsym(h, "lzma_code"));
    end = reinterpret_cast<end_t>(dlsym(h, "lzma_end"));
    return dec && code && end;
  }();
  if (!ok) return false;
  xz_options_lzma opt;
  memset(&opt, 0, sizeof(opt));
  uint32_t dict = 4096;
  while (dict < n && dict < (1u << 30)) dict <<= 1;
  opt.dict_size = dict; opt.lc = 3; opt.lp = 0; opt.pb = 2;
  xz_filter filters[2] = {{0x4000000000000001ull /* LZMA_FILTER_LZMA1 */, &opt}, {~0ull /* LZMA_VLI_UNKNOWN */, nullptr}};
  xz_stream st;
  memset(&st, 0, sizeof(st));
  if (dec(&st, filters) != 0) return false;
  st.next_in = src; st.avail_in = srclen; st.next_out = dst; st.avail_out = n;
  bool good = false;
  for (;;) {
    const size_t before_in = st.avail_in, before_out = st.avail_out;
    const int rc = code(&st, 0 /* LZMA_RUN */);
    if (st.avail_out == 0 && (rc == 0 || rc == 1)) { good = true; break; }
    if (rc != 0 || (st.avail_in == before_in && st.avail_out == before_out)) break;
  }
  end(&st);
  return good;
}

// LZ4 block format with a known decompressed size (what LZ4_decompress_fast consumes), bounds checked
bool lz4_decompress(const uint8_t* src, size_t srclen, uint8_t* dst, size_t n) {
  const uint8_t* ip = src; const uint8_t* const iend = src + srclen;
  uint8_t* op = dst; uint8_t* const oend = dst + n;
  for (;;) {
    if (ip >= iend) return false;
    const unsigned token = *ip++;
    size_t lit = token >> 4;
    if (lit == 15) { unsigned b; do { if (ip >= iend) return false; b = *ip++; lit += b; } while (b == 255); }
    if (lit > (size_t)(iend - ip) || lit > (size_t)(oend - op)) return false;
    memcpy(op, ip, lit); op += lit; ip += lit;
    if (op == oend) return true;                     // the last sequence has literals only
    if (iend - ip < 2) return false;
    const size_t off = ip[0] | ((size_t)ip[1] << 8); ip += 2;
    if (off == 0 || off > (size_t)(op - dst)) return false;
    size_t ml = token & 15;
    if (ml == 15) { unsigned b; do { if (ip >= iend) return false; b = *ip++; ml += b; } while (b == 255); }
    ml += 4;
    if (ml > (size_t)(oend - op)) return false;
    const uint8_t* m = op - off;
    for (size_t i = 0; i < ml; ++i) op[i] = m[i];    // byte-wise: matches may overlap their own output
    op += ml;
  }
}

uint32_t bytes_of(uint32_t component_type) {
  static const uint32_t w[10] = {1, 2, 4, 8, 1, 2, 4, 8, 4, 8};   // U8 U16 U32 U64 I8 I16 I32 I64 F32 F64
  return component_type < 10 ? w[component_type] : 0;
}

// ExtendedOctree::ComputeMetadata + ComputeBrickSize through the planner the builder uses
bool plan_sizes(const tvk_uvf_volume& v, tvk_build_info& info, std::vector<uint32_t>& sizes) {
  if (v.brick_size[0] != v.brick_size[1] || v.brick_size[0] != v.brick_size[2]) return false;
  tvk_build_desc d;
  memset(&d, 0, sizeof(d));
  for (int a = 0; a < 3; ++a) { if (v.volume_size[a] == 0 || v.volume_size[a] > 0xFFFFFFFFull) return false; d.volume_size[a] = (uint32_t)v.volume_size[a]; }
  d.brick_size = (uint32_t)v.brick_size[0]; d.overlap = v.overlap;
  d.bytes_per_voxel = 1;
  if (tvk_build_plan(&d, &info) != TVK_OK) return false;
  const uint32_t core = d.brick_size - 2 * d.overlap;
  sizes.resize(3 * info.total_bricks);
  uint64_t i = 0;
  for (uint32_t l = 0; l < info.lod_count; ++l)
    for (uint32_t z = 0; z < info.lod_bricks[l][2]; ++z)
      for (uint32_t y = 0; y < info.lod_bricks[l][1]; ++y)
        for (uint32_t x = 0; x < info.lod_bricks[l][0]; ++x, ++i) {
          const uint32_t c[3] = {x, y, z};
          for (int a = 0; a < 3; ++a) {
            const bool last = c[a] >= info.lod_bricks[l][a] - 1;
            const uint32_t rem = info.lod_size[l][a] % core;
            sizes[3 * i + a] = (last && rem) ? 2 * d.overlap + rem : d.brick_size;
          }
        }
  return true;
}

}  // namespace

extern "C" void tvk_md5(const void* data, uint64_t n, uint8_t digest[16]) {
  Md5 m;
  m.update(data, (size_t)n);
  m.final(digest);
}

// =========================================================================== writer
extern "C" int tvk_uvf_write(const char* path, const tvk_uvf_write_desc* d, char* err, size_t errlen) {
  if (!path || !d || !d->store || !d->brick_offsets || !d->brick_sizes || !d->minmax || d->total_bricks == 0) {
    set_err(err, errlen, "tvk_uvf_write: missing argument");
    return TVK_ERR_INVALID;
  }
  const uint32_t bpv = bytes_of(d->vol.component_type) * (uint32_t)d->vol.component_count;
  if (!bpv) { set_err(err, errlen, "tvk_uvf_write: bad component type"); return TVK_ERR_INVALID; }
  if (d->vol.compression != TVK_UVF_COMP_NONE && d->vol.compression != TVK_UVF_COMP_ZLIB) {
    set_err(err, errlen, "tvk_uvf_write: only uncompressed and zlib bricks can be written");
    return TVK_ERR_UNSUPPORTED;
  }
  {   // the ToC the caller hands over must be the ExtendedOctree's
    tvk_build_info info;
    std::vector<uint32_t> sizes;
    if (!plan_sizes(d->vol, info, sizes) || info.total_bricks != d->total_bricks ||
        memcmp(sizes.data(), d->brick_sizes, sizeof(uint32_t) * sizes.size()) != 0) {
      set_err(err, errlen, "tvk_uvf_write: brick table does not match the octree of this volume");
      return TVK_ERR_INVALID;
    }
  }
  const uint64_t nb = d->total_bricks;
  const std::string toc_id = d->toc_block_id ? d->toc_block_id : "Volume converted by ImageVis3D";
  FILE* f = fopen(path, "wb+");
  if (!f) { set_err(err, errlen, std::string("tvk_uvf_write: cannot create ") + path); return TVK_ERR_INVALID; }

  const uint64_t kGlobal = 57;
  const uint64_t toc_hdr = toc_id.size() + 32;
  const uint64_t oct_hdr = 4 + 8 + 1 + 24 + 24 + 24 + 4 + 4 + 8 + 4 + nb * 36;   // ExtendedOctree::ComputeHeaderSize
  const uint64_t oct_pos = kGlobal + toc_hdr;

  // ---- brick payloads (compressed in parallel, written in ToC order)
  std::vector<uint64_t> t_off(nb), t_len(nb), t_valid(nb);
  std::vector<uint32_t> t_comp(nb);
  bool ok = fseeko(f, (off_t)(oct_pos + oct_hdr), SEEK_SET) == 0;
  uint64_t cur = oct_hdr;
  const uint8_t* store = static_cast<const uint8_t*>(d->store);
  const unsigned workers = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  const uint64_t batch = d->vol.compression == TVK_UVF_COMP_ZLIB ? workers * 4 : 1;
  std::vector<std::vector<uint8_t>> zbuf(batch);
  for (uint64_t b0 = 0; b0 < nb && ok; b0 += batch) {
    const uint64_t bn = std::min(batch, nb - b0);
    if (d->vol.compression == TVK_UVF_COMP_ZLIB) {
      std::atomic<uint64_t> next(0);
      auto work = [&]() {
        for (uint64_t k; (k = next.fetch_add(1)) < bn;) {
          const uint64_t i = b0 + k;
          const size_t n = (size_t)d->brick_sizes[3 * i] * d->brick_sizes[3 * i + 1] * d->brick_sizes[3 * i + 2] * bpv;
          zbuf[k].resize(n);
          const size_t z = z_compress(store + d->brick_offsets[i], n, zbuf[k].data(), d->vol.compression_level);
          t_len[i] = z < n ? z : n;
          t_comp[i] = z < n ? TVK_UVF_COMP_ZLIB : TVK_UVF_COMP_NONE;
        }
      };
      std::vector<std::thread> th;
      for (unsigned w = 1; w < std::min<uint64_t>(workers, bn); ++w) th.emplace_back(work);
      work();
      for (auto& t : th) t.join();
    }
    for (uint64_t k = 0; k < bn && ok; ++k) {
      const uint64_t i = b0 + k;
      const size_t n = (size_t)d->brick_sizes[3 * i] * d->brick_sizes[3 * i + 1] * d->brick_sizes[3 * i + 2] * bpv;
      t_valid[i] = n;
      if (d->vol.compression != TVK_UVF_COMP_ZLIB) { t_len[i] = n; t_comp[i] = TVK_UVF_COMP_NONE; }
      t_off[i] = cur;
      const uint8_t* src = t_comp[i] == TVK_UVF_COMP_ZLIB ? zbuf[k].data() : store + d->brick_offsets[i];
      ok = fwrite(src, 1, (size_t)t_len[i], f) == t_len[i];
      cur += t_len[i];
    }
  }
  const uint64_t oct_size = cur;                                   // m_iSize: header + ToC + payloads

  // ---- the remaining blocks, serialised in memory
  Buf tail;
  const bool has_h1 = d->hist1d && d->hist1d_size, has_h2 = d->hist2d && d->hist2d_x && d->hist2d_y;
  if (has_h1) {
    block_header(tail, "1D Histogram for datablock " + toc_id, BS_HIST1D, 8 + 8 * d->hist1d_size, false);
    tail.put<uint64_t>(d->hist1d_size);
    tail.raw(d->hist1d, 8 * d->hist1d_size);
  }
  if (has_h2) {
    block_header(tail, "2D Histogram for datablock " + toc_id, BS_HIST2D, 4 + 16 + 8 * d->hist2d_x * d->hist2d_y, false);
    tail.put<float>(d->max_grad_magnitude);
    tail.put<uint64_t>(d->hist2d_x);
    tail.put<uint64_t>(d->hist2d_y);
    tail.raw(d->hist2d, 8 * d->hist2d_x * d->hist2d_y);
  }
  block_header(tail, "Brick Max/Min Values", BS_MAXMIN, 16 + 32 * nb, false);
  tail.put<uint64_t>(nb);
  tail.put<uint64_t>(1);
  for (uint64_t i = 0; i < nb; ++i) {
    tail.put<double>(d->minmax[2 * i]); tail.put<double>(d->minmax[2 * i + 1]);
    tail.put<double>(-DBL_MAX); tail.put<double>(DBL_MAX);         // MaxMinDataBlock::SetDataFromFlatVector
  }
  {
    uint64_t chars = 0;
    for (uint32_t i = 0; i < d->kv_count; ++i) chars += strlen(d->kv_keys[i]) + strlen(d->kv_values[i]);
    block_header(tail, "KeyValue Pair Block", BS_KEY_VALUE, 8 + chars + 16ull * d->kv_count, true);
    tail.put<uint64_t>(d->kv_count);
    for (uint32_t i = 0; i < d->kv_count; ++i) {
      tail.put<uint64_t>(strlen(d->kv_keys[i])); tail.str(d->kv_keys[i]);
      tail.put<uint64_t>(strlen(d->kv_values[i])); tail.str(d->kv_values[i]);
    }
  }
  ok = ok && fwrite(tail.v.data(), 1, tail.v.size(), f) == tail.v.size();

  // ---- global header, TOC block header, octree header + ToC
  Buf head;
  head.str("UVF-DATA");
  head.put<uint8_t>(0);            // little endian
  head.put<uint64_t>(5);           // UVFVERSION
  head.put<uint64_t>(2);           // CS_MD5
  head.put<uint64_t>(16);
  for (int i = 0; i < 16; ++i) head.put<uint8_t>(0);
  head.put<uint64_t>(0);           // ulOffsetToFirstDataBlock
  block_header(head, toc_id, BS_TOC, oct_size, false);
  head.put<uint32_t>(d->vol.component_type);
  head.put<uint64_t>(d->vol.component_count);
  head.put<uint8_t>(0);            // m_bPrecomputedNormals
  for (int a = 0; a < 3; ++a) head.put<uint64_t>(d->vol.volume_size[a]);
  for (int a = 0; a < 3; ++a) head.put<double>(d->vol.aspect[a]);
  for (int a = 0; a < 3; ++a) head.put<uint64_t>(d->vol.brick_size[a]);
  head.put<uint32_t>(d->vol.overlap);
  head.put<uint32_t>(2);           // ExtendedOctree m_iVersion
  head.put<uint64_t>(oct_size);
  head.put<uint32_t>(d->vol.compression_level);
  for (uint64_t i = 0; i < nb; ++i) {
    head.put<uint64_t>(t_off[i]); head.put<uint64_t>(t_len[i]); head.put<uint32_t>(t_comp[i]);
    head.put<uint64_t>(t_valid[i]); head.put<uint32_t>(0); head.put<uint32_t>(0);
  }
  ok = ok && head.v.size() == oct_pos + oct_hdr && fseeko(f, 0, SEEK_SET) == 0 && fwrite(head.v.data(), 1, head.v.size(), f) == head.v.size();
  // ---- checksum over everything behind it (UVF::ComputeChecksum), patched in at offset 33
  uint8_t digest[16];
  ok = ok && fflush(f) == 0 && md5_of_file_tail(f, 49, digest) && fseeko(f, 33, SEEK_SET) == 0 && fwrite(digest, 1, 16, f) == 16;
  ok = (fclose(f) == 0) && ok;
  if (!ok) { set_err(err, errlen, std::string("tvk_uvf_write: I/O error writing ") + path); return TVK_ERR_INVALID; }
  return TVK_OK;
}

// =========================================================================== reader
struct tvk_uvf {
  FILE* f = nullptr;
  std::string err;
  tvk_uvf_info info{};
  uint64_t oct_pos = 0;                       // file offset of the octree header
  std::vector<uint64_t> t_off, t_len;         // absolute file offset, stored length
  std::vector<uint32_t> t_comp, sizes;
  std::vector<double> minmax;                 // 4 per brick
  std::vector<uint64_t> hist1d, hist2d;
  std::vector<std::pair<std::string, std::string>> kv;
  uint32_t bpv = 0;
  uint64_t rdb_data_pos = 0;                  // file offset of the finest level's first brick
  uint64_t file_size = 0;
  uint32_t want_ts = 0;                       // which time step's blocks this handle reads (UVFDataset::FindSuitableDataBlocks)
};

namespace {
struct Rd {
  FILE* f; bool ok = true;
  template <typename T> T get() { T x{}; if (fread(&x, sizeof(T), 1, f) != 1) ok = false; return x; }
  std::string str(uint64_t n) {
    std::string s;
    if (n > (1u << 20)) { ok = false; return s; }
    s.resize((size_t)n);
    if (n && fread(&s[0], 1, (size_t)n, f) != n) ok = false;
    return s;
  }
};
}  // namespace

// No C++ exception may cross the C ABI: every entry point that allocates from sizes found in the FILE runs inside this
// guard (std::bad_alloc -> TVK_ERR_NOMEM, anything else -> TVK_ERR_INVALID).  On-disk sizes are checked against the
// file size before they are trusted (a UVF's MD5 only guards against accidental corruption, and verifying it is optional).
#define TVK_UVF_GUARD(expr_on_error, ...)                                                         \
  try { __VA_ARGS__ }                                                                             \
  catch (const std::bad_alloc&) { expr_on_error; return TVK_ERR_NOMEM; }                          \
  catch (...) { expr_on_error; return TVK_ERR_INVALID; }

static int uvf_open_impl(const char* path, int verify, uint32_t timestep, tvk_uvf** out, char* err, size_t errlen, tvk_uvf*& u);

extern "C" int tvk_uvf_open_timestep(const char* path, int verify, uint32_t timestep, tvk_uvf** out, char* err, size_t errlen) {
  if (!path || !out) return TVK_ERR_INVALID;
  *out = nullptr;
  tvk_uvf* u = nullptr;
  TVK_UVF_GUARD({ if (u) tvk_uvf_close(u); *out = nullptr; set_err(err, errlen, "out of memory or corrupt file while parsing the UVF headers"); },
                return uvf_open_impl(path, verify, timestep, out, err, errlen, u);)
}
extern "C" int tvk_uvf_open(const char* path, int verify, tvk_uvf** out, char* err, size_t errlen) {
  return tvk_uvf_open_timestep(path, verify, 0, out, err, errlen);
}

static int uvf_open_impl(const char* path, int verify, uint32_t timestep, tvk_uvf** out, char* err, size_t errlen, tvk_uvf*& u) {
  FILE* f = fopen(path, "rb");
  if (!f) { set_err(err, errlen, "file not found or access denied"); return TVK_ERR_INVALID; }
  u = new tvk_uvf();
  u->f = f;
  u->want_ts = timestep;
  // UVFDataset::FindSuitableDataBlocks (uvfDataset.cpp:635-745) hands the k-th volume / 1-D histogram / 2-D histogram /
  // max-min block of the file to time step k, in file order
  uint32_t toc_seen = 0, h1_seen = 0, h2_seen = 0, mm_seen = 0;
  if (fseeko(f, 0, SEEK_END) == 0) u->file_size = (uint64_t)ftello(f);
  fseeko(f, 0, SEEK_SET);
  u->info.toc_block = -1;
  u->info.rdb_block = -1;
  auto fail = [&](const std::string& m, int code) { set_err(err, errlen, m); tvk_uvf_close(u); u = nullptr; return code; };
  Rd r{f};
  char magic[8];
  if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "UVF-DATA", 8) != 0) return fail("file magic not found", TVK_ERR_INVALID);
  const uint8_t big = r.get<uint8_t>();
  if (big) return fail("big-endian UVF files are not supported", TVK_ERR_UNSUPPORTED);
  u->info.uvf_version = r.get<uint64_t>();
  const uint64_t cs_sem = r.get<uint64_t>(), cs_len = r.get<uint64_t>();
  if (!r.ok || cs_len > 64) return fail("corrupt global header", TVK_ERR_INVALID);
  uint8_t cs[64];
  if (cs_len && fread(cs, 1, (size_t)cs_len, f) != cs_len) return fail("corrupt global header", TVK_ERR_INVALID);
  const uint64_t first = r.get<uint64_t>();
  if (!r.ok) return fail("file too small to be a UVF file", TVK_ERR_INVALID);
  if (verify && cs_sem != 0) {
    if (cs_sem != 2 || cs_len != 16) return fail("only MD5 checksums can be verified", TVK_ERR_UNSUPPORTED);
    uint8_t digest[16];
    if (!md5_of_file_tail(f, 33 + 16, digest) || memcmp(digest, cs, 16) != 0) return fail("UVF::VerifyChecksum: checksum mismatch", TVK_ERR_INVALID);
    u->info.checksum_ok = 1;
  }
  uint64_t pos = 8 + 1 + 24 + cs_len + 8 + first;          // GlobalHeader::GetDataPos
  for (;;) {
    if (fseeko(f, (off_t)pos, SEEK_SET)) return fail("corrupt block chain", TVK_ERR_INVALID);
    const uint64_t idlen = r.get<uint64_t>();
    const std::string id = r.str(idlen);
    const uint64_t sem = r.get<uint64_t>();
    (void)r.get<uint64_t>();                                // block compression scheme (unused by the reference too)
    const uint64_t next = r.get<uint64_t>();
    if (!r.ok) return fail("corrupt block header", TVK_ERR_INVALID);
    const uint32_t bi = u->info.block_count++;
    if (bi < 16) u->info.block_semantics[bi] = (uint32_t)sem;
    if (sem == BS_TOC && toc_seen++ == u->want_ts) {
      u->info.toc_block = (int32_t)bi;
      u->oct_pos = (uint64_t)ftello(f);
      tvk_uvf_volume& v = u->info.vol;
      v.component_type = r.get<uint32_t>();
      v.component_count = r.get<uint64_t>();
      const uint8_t normals = r.get<uint8_t>();
      for (int a = 0; a < 3; ++a) v.volume_size[a] = r.get<uint64_t>();
      for (int a = 0; a < 3; ++a) v.aspect[a] = r.get<double>();
      for (int a = 0; a < 3; ++a) v.brick_size[a] = r.get<uint64_t>();
      v.overlap = r.get<uint32_t>();
      if (u->info.uvf_version > 4) {
        u->info.octree_version = r.get<uint32_t>();
        if (u->info.octree_version == 0) return fail("corrupt octree header", TVK_ERR_INVALID);
        (void)r.get<uint64_t>();                            // m_iSize
        if (u->info.octree_version > 1) v.compression_level = r.get<uint32_t>();
      } else {
        return fail("UVF versions before 5 store the octree without a version field; not supported", TVK_ERR_UNSUPPORTED);
      }
      if (!r.ok || normals || v.component_count != 1 || bytes_of(v.component_type) == 0)
        return fail("unsupported octree (multi-component or precomputed normals)", TVK_ERR_UNSUPPORTED);
      u->bpv = bytes_of(v.component_type);
      tvk_build_info bi2;
      if (!plan_sizes(v, bi2, u->sizes)) return fail("unsupported octree geometry (non-cubic bricks?)", TVK_ERR_UNSUPPORTED);
      u->info.lod_count = bi2.lod_count;
      u->info.total_bricks = bi2.total_bricks;
      const uint64_t nb = bi2.total_bricks;
      if (nb == 0 || nb > u->file_size / 36) return fail("brick table larger than the file", TVK_ERR_INVALID);   // 36 bytes per ToC entry
      u->t_off.resize(nb); u->t_len.resize(nb); u->t_comp.resize(nb);
      for (uint64_t i = 0; i < nb; ++i) {
        const uint64_t rel = r.get<uint64_t>();
        u->t_off[i] = u->oct_pos + rel;
        u->t_len[i] = r.get<uint64_t>();
        u->t_comp[i] = r.get<uint32_t>();
        (void)r.get<uint64_t>(); (void)r.get<uint32_t>(); (void)r.get<uint32_t>();
        const uint64_t raw = (uint64_t)u->sizes[3 * i] * u->sizes[3 * i + 1] * u->sizes[3 * i + 2] * u->bpv;
        u->info.store_bytes += raw;
        // a brick lies inside the file, and a stored (possibly compressed) brick is never much larger than its voxels
        if (!r.ok || rel > u->file_size || u->t_len[i] > u->file_size || u->t_off[i] + u->t_len[i] > u->file_size ||
            u->t_len[i] > raw + raw / 2 + 65536 || raw > 0xFFFFFFFFull)
          return fail("corrupt brick table (entry outside the file or larger than its brick)", TVK_ERR_INVALID);
      }
      if (!r.ok) return fail("truncated brick table", TVK_ERR_INVALID);
      v.compression = u->t_comp[0];
    } else if (sem == BS_REG_NDIM_GRID && u->info.rdb_block < 0) {
      // RasterDataBlock::GetHeaderFromFile (RasterDataBlock.cpp:154-208); only 3-D scalar 8/16-bit grids are decoded
      const uint64_t dims = r.get<uint64_t>();
      bool usable = r.ok && dims == 3;
      uint64_t dom[3] = {0, 0, 0}, brick[3] = {0, 0, 0}, ovl[3] = {0, 0, 0}, groups[3] = {0, 0, 0};
      if (usable) {
        for (int a = 0; a < 3; ++a) (void)r.get<uint64_t>();                 // domain semantics
        for (int a = 0; a < 16; ++a) (void)r.get<double>();                  // (n+1)^2 domain transformation
        for (int a = 0; a < 3; ++a) dom[a] = r.get<uint64_t>();
        for (int a = 0; a < 3; ++a) brick[a] = r.get<uint64_t>();
        for (int a = 0; a < 3; ++a) ovl[a] = r.get<uint64_t>();
        for (int a = 0; a < 3; ++a) (void)r.get<uint64_t>();                 // LOD decimation factors
        for (int a = 0; a < 3; ++a) groups[a] = r.get<uint64_t>();
        uint64_t lodIdx = 1;
        for (int a = 0; a < 3; ++a) if (groups[a] >= lodIdx) lodIdx = groups[a] + 1;
        uint64_t levels = 0;
        usable = lodIdx <= 64;                                                // a LOD group index beyond that is a corrupt header
        for (uint64_t i = 0; usable && i < lodIdx && r.ok; ++i) levels = std::max<uint64_t>(levels, r.get<uint64_t>());
        usable = usable && r.ok;
        if (!usable) { r.ok = true; if (next == 0) break; pos += next; continue; }
        const uint64_t edim = r.get<uint64_t>();
        usable = r.ok && edim == 1;
        if (usable) {
          const uint64_t esize = r.get<uint64_t>();
          usable = r.ok && esize == 1;
          if (usable) {
            (void)r.get<uint64_t>();                                          // element semantic
            const uint64_t bits = r.get<uint64_t>();
            (void)r.get<uint64_t>();                                          // mantissa
            const uint8_t sgn = r.get<uint8_t>();
            const uint64_t toData = r.get<uint64_t>();
            usable = r.ok && (bits == 8 || bits == 16);
            for (int a = 0; a < 3; ++a) usable = usable && dom[a] > 0 && dom[a] <= (1ull << 20) && brick[a] > ovl[a];
            usable = usable && dom[0] * dom[1] * dom[2] * (bits / 8) <= u->file_size;      // the finest level is stored in full
            if (usable) {
              u->info.rdb_block = (int32_t)bi;
              u->info.rdb_bits = (uint32_t)bits; u->info.rdb_signed = sgn ? 1u : 0u; u->info.rdb_lod_levels = levels;
              for (int a = 0; a < 3; ++a) { u->info.rdb_domain[a] = dom[a]; u->info.rdb_brick[a] = brick[a]; u->info.rdb_overlap[a] = ovl[a]; }
              u->rdb_data_pos = (uint64_t)ftello(f) + toData;
            }
          }
        }
      }
      r.ok = true;                                                            // an undecodable grid block is skipped, not fatal
    } else if (sem == BS_HIST1D && h1_seen++ == u->want_ts) {
      const uint64_t n = r.get<uint64_t>();
      if (!r.ok || n > (1u << 24)) return fail("corrupt 1D histogram", TVK_ERR_INVALID);
      u->hist1d.resize((size_t)n);
      if (n && fread(u->hist1d.data(), 8, (size_t)n, f) != n) return fail("corrupt 1D histogram", TVK_ERR_INVALID);
      u->info.hist1d_size = n;
    } else if (sem == BS_HIST2D && h2_seen++ == u->want_ts) {
      u->info.max_grad_magnitude = r.get<float>();
      const uint64_t nx = r.get<uint64_t>(), ny = r.get<uint64_t>();
      if (!r.ok || nx * ny > (1u << 26)) return fail("corrupt 2D histogram", TVK_ERR_INVALID);
      u->hist2d.resize((size_t)(nx * ny));
      if (nx * ny && fread(u->hist2d.data(), 8, (size_t)(nx * ny), f) != nx * ny) return fail("corrupt 2D histogram", TVK_ERR_INVALID);
      u->info.hist2d_x = nx; u->info.hist2d_y = ny;
    } else if (sem == BS_MAXMIN && mm_seen++ == u->want_ts) {
      const uint64_t n = r.get<uint64_t>(), comps = r.get<uint64_t>();
      if (!r.ok || comps == 0 || comps > 64 || n > u->file_size / (32 * comps)) return fail("corrupt max/min block", TVK_ERR_INVALID);
      u->minmax.resize((size_t)(4 * n));
      std::vector<double> row(4 * comps);
      for (uint64_t i = 0; i < n; ++i) {
        if (fread(row.data(), 8, (size_t)(4 * comps), f) != 4 * comps) return fail("corrupt max/min block", TVK_ERR_INVALID);
        memcpy(&u->minmax[4 * i], row.data(), 32);          // component 0
      }
      u->info.maxmin_bricks = n;
    } else if (sem == BS_KEY_VALUE && u->kv.empty()) {
      const uint64_t n = r.get<uint64_t>();
      if (n > u->file_size / 16) return fail("corrupt key/value block", TVK_ERR_INVALID);
      for (uint64_t i = 0; i < n && r.ok; ++i) {
        const std::string k = r.str(r.get<uint64_t>());
        const std::string v = r.str(r.get<uint64_t>());
        u->kv.emplace_back(k, v);
      }
      if (!r.ok) return fail("corrupt key/value block", TVK_ERR_INVALID);
      u->info.kv_count = (uint32_t)u->kv.size();
    }
    if (next == 0) break;
    pos += next;
  }
  u->info.timesteps = toc_seen;
  if (toc_seen && u->want_ts >= toc_seen) return fail("time step " + std::to_string(u->want_ts) + " requested, the file has " + std::to_string(toc_seen), TVK_ERR_INVALID);
  *out = u;
  return TVK_OK;
}

extern "C" void tvk_uvf_close(tvk_uvf* u) {
  if (!u) return;
  if (u->f) fclose(u->f);
  delete u;
}

extern "C" const char* tvk_uvf_last_error(const tvk_uvf* u) { return u ? u->err.c_str() : "null handle"; }

extern "C" int tvk_uvf_get_info(const tvk_uvf* u, tvk_uvf_info* info) {
  if (!u || !info) return TVK_ERR_INVALID;
  *info = u->info;
  return TVK_OK;
}

extern "C" int tvk_uvf_get_toc(const tvk_uvf* u, uint64_t* off, uint64_t* len, uint32_t* comp, uint32_t* sizes3, uint64_t n) {
  if (!u || u->info.toc_block < 0 || n != u->info.total_bricks) return TVK_ERR_INVALID;
  if (off) memcpy(off, u->t_off.data(), 8 * n);
  if (len) memcpy(len, u->t_len.data(), 8 * n);
  if (comp) memcpy(comp, u->t_comp.data(), 4 * n);
  if (sizes3) memcpy(sizes3, u->sizes.data(), 12 * n);
  return TVK_OK;
}

static int uvf_read_bricks_impl(tvk_uvf* u, uint64_t first, uint64_t count, void* dst, const uint64_t* dst_offsets);
extern "C" int tvk_uvf_read_bricks(tvk_uvf* u, uint64_t first, uint64_t count, void* dst, const uint64_t* dst_offsets) {
  if (!u || u->info.toc_block < 0 || !dst || first > u->info.total_bricks || count > u->info.total_bricks - first) return TVK_ERR_INVALID;
  TVK_UVF_GUARD({ u->err = "out of memory or corrupt data while reading bricks"; }, return uvf_read_bricks_impl(u, first, count, dst, dst_offsets);)
}
static int uvf_read_bricks_impl(tvk_uvf* u, uint64_t first, uint64_t count, void* dst, const uint64_t* dst_offsets) {
  std::vector<uint8_t> tmp;
  uint64_t packed = 0;
  for (uint64_t k = 0; k < count; ++k) {
    const uint64_t i = first + k;
    const size_t n = (size_t)u->sizes[3 * i] * u->sizes[3 * i + 1] * u->sizes[3 * i + 2] * u->bpv;
    uint8_t* o = static_cast<uint8_t*>(dst) + (dst_offsets ? dst_offsets[k] : packed);
    packed += n;
    if (fseeko(u->f, (off_t)u->t_off[i], SEEK_SET)) { u->err = "seek failed"; return TVK_ERR_INVALID; }
    if (u->t_comp[i] == TVK_UVF_COMP_NONE) {
      if (u->t_len[i] != n || fread(o, 1, n, u->f) != n) { u->err = "short read"; return TVK_ERR_INVALID; }
      continue;
    }
    tmp.resize((size_t)u->t_len[i]);
    if (fread(tmp.data(), 1, tmp.size(), u->f) != tmp.size()) { u->err = "short read"; return TVK_ERR_INVALID; }
    bool ok;
    if (u->t_comp[i] == TVK_UVF_COMP_ZLIB) ok = z_decompress(tmp.data(), tmp.size(), o, n);
    else if (u->t_comp[i] == TVK_UVF_COMP_LZ4) ok = lz4_decompress(tmp.data(), tmp.size(), o, n);
    else if (u->t_comp[i] == TVK_UVF_COMP_LZMA) ok = lzma_decompress(tmp.data(), tmp.size(), o, n);
    else if (u->t_comp[i] == TVK_UVF_COMP_BZIP) ok = bz_decompress(tmp.data(), tmp.size(), o, n);
    else { u->err = "unknown brick compression"; return TVK_ERR_UNSUPPORTED; }
    if (!ok) { u->err = "brick decompression failed"; return TVK_ERR_INVALID; }
  }
  return TVK_OK;
}

// the bricks as the file stores them (compressed or not): what a GPU-side inflate pages from
extern "C" int tvk_uvf_read_bricks_raw(tvk_uvf* u, uint64_t first, uint64_t count, void* dst, const uint64_t* dst_offsets) {
  if (!u || u->info.toc_block < 0 || !dst || first > u->info.total_bricks || count > u->info.total_bricks - first) return TVK_ERR_INVALID;
  uint64_t packed = 0;
  for (uint64_t k = 0; k < count; ++k) {
    const uint64_t i = first + k;
    uint8_t* o = static_cast<uint8_t*>(dst) + (dst_offsets ? dst_offsets[k] : packed);
    packed += u->t_len[i];
    if (fseeko(u->f, (off_t)u->t_off[i], SEEK_SET)) { u->err = "seek failed"; return TVK_ERR_INVALID; }
    if (fread(o, 1, (size_t)u->t_len[i], u->f) != (size_t)u->t_len[i]) { u->err = "short read"; return TVK_ERR_INVALID; }
  }
  return TVK_OK;
}

// RasterDataBlock::ComputeBricks (RasterDataBlock.cpp:321-349): brick sizes along one axis
static std::vector<uint64_t> rdb_axis_bricks(uint64_t size, uint64_t brick, uint64_t overlap) {
  std::vector<uint64_t> v;
  if (size <= brick) { v.push_back(size); return v; }
  do {
    if (size + overlap <= brick) { v.push_back(size); break; }
    v.push_back(brick);
    size = size + overlap - brick;
  } while (size > overlap);
  return v;
}

static int uvf_read_rdb_impl(tvk_uvf* u, void* dst, uint64_t dst_bytes);
extern "C" int tvk_uvf_read_rdb_volume(tvk_uvf* u, void* dst, uint64_t dst_bytes) {
  if (!u || !dst || u->info.rdb_block < 0) return TVK_ERR_INVALID;
  TVK_UVF_GUARD({ u->err = "out of memory or corrupt data while reading the RasterDataBlock"; }, return uvf_read_rdb_impl(u, dst, dst_bytes);)
}
static int uvf_read_rdb_impl(tvk_uvf* u, void* dst, uint64_t dst_bytes) {
  const tvk_uvf_info& i = u->info;
  const uint64_t bpe = i.rdb_bits / 8;
  if (dst_bytes != i.rdb_domain[0] * i.rdb_domain[1] * i.rdb_domain[2] * bpe) { u->err = "destination size does not match the domain"; return TVK_ERR_INVALID; }
  std::vector<uint64_t> bs[3];
  for (int a = 0; a < 3; ++a) bs[a] = rdb_axis_bricks(i.rdb_domain[a], i.rdb_brick[a], i.rdb_overlap[a]);
  // finest level first in the block; bricks in Cartesian-product order, x fastest (:289-312, :365-391)
  if (fseeko(u->f, (off_t)u->rdb_data_pos, SEEK_SET)) { u->err = "seek failed"; return TVK_ERR_INVALID; }
  uint8_t* out = static_cast<uint8_t*>(dst);
  std::vector<uint8_t> buf;
  for (size_t kz = 0; kz < bs[2].size(); ++kz)
    for (size_t ky = 0; ky < bs[1].size(); ++ky)
      for (size_t kx = 0; kx < bs[0].size(); ++kx) {
        const uint64_t sx = bs[0][kx], sy = bs[1][ky], sz = bs[2][kz];
        buf.resize((size_t)(sx * sy * sz * bpe));
        if (fread(buf.data(), 1, buf.size(), u->f) != buf.size()) { u->err = "short read in RasterDataBlock"; return TVK_ERR_INVALID; }
        // brick k starts at k * (brickSize - overlap); neighbouring bricks share `overlap` voxels (identical data)
        const uint64_t ox = kx * (i.rdb_brick[0] - i.rdb_overlap[0]), oy = ky * (i.rdb_brick[1] - i.rdb_overlap[1]),
                       oz = kz * (i.rdb_brick[2] - i.rdb_overlap[2]);
        for (uint64_t z = 0; z < sz; ++z)
          for (uint64_t y = 0; y < sy; ++y)
            memcpy(out + ((ox) + (oy + y) * i.rdb_domain[0] + (oz + z) * i.rdb_domain[0] * i.rdb_domain[1]) * bpe,
                   buf.data() + (y + z * sy) * sx * bpe, (size_t)(sx * bpe));
      }
  return TVK_OK;
}

extern "C" int tvk_uvf_get_minmax(const tvk_uvf* u, double* mm, uint64_t n) {
  if (!u || !mm || n != u->info.maxmin_bricks) return TVK_ERR_INVALID;
  memcpy(mm, u->minmax.data(), 32 * n);
  return TVK_OK;
}
extern "C" int tvk_uvf_get_hist1d(const tvk_uvf* u, uint64_t* bins, uint64_t n) {
  if (!u || !bins || n != u->info.hist1d_size) return TVK_ERR_INVALID;
  memcpy(bins, u->hist1d.data(), 8 * n);
  return TVK_OK;
}
extern "C" int tvk_uvf_get_hist2d(const tvk_uvf* u, uint64_t* bins, uint64_t n) {
  if (!u || !bins || n != u->info.hist2d_x * u->info.hist2d_y) return TVK_ERR_INVALID;
  memcpy(bins, u->hist2d.data(), 8 * n);
  return TVK_OK;
}
extern "C" int tvk_uvf_get_kv(const tvk_uvf* u, uint32_t i, const char** key, const char** value) {
  if (!u || i >= u->kv.size() || !key || !value) return TVK_ERR_INVALID;
  *key = u->kv[i].first.c_str();
  *value = u->kv[i].second.c_str();
  return TVK_OK;
}
