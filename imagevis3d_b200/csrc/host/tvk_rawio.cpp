// Raw-volume front ends: what ImageVis3D reads BEFORE the ExtendedOctree builder sees a flat volume.
//   QVIS .dat / .raw pairs   IO/QVISConverter.cpp:60-150  (ConvertToRAW: header parse)
//   NRRD .nrrd / .nhdr       IO/NRRDConverter.cpp:60-440  (raw and gzip encodings; text / hex / bzip2 are refused)
//   the "KEY: value" files   IO/KeyValueFileParser.cpp:230-300
// The parsers fill a tvk_raw_desc (the out-parameters of AbstrConverter::ConvertToRAW); tvk_raw_read then delivers the
// voxels in host byte order, and the renderer host (tvkr_load_volume_file) quantises them on the device (tvk_quantize, row B5)
// and bricks them (tvk_build_octree).  Host code only: no CUDA in here.
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>
#include <zlib.h>
#include "tvk_renderer.h"

namespace {

std::string trim_left(const std::string& s) { size_t a = s.find_first_not_of(" \t\r\n"); return a == std::string::npos ? std::string() : s.substr(a); }
std::string trim_right(const std::string& s) { size_t b = s.find_last_not_of(" \t\r\n"); return b == std::string::npos ? std::string() : s.substr(0, b + 1); }
std::string trim(const std::string& s) { return trim_right(trim_left(s)); }
std::string upper(std::string s) { std::transform(s.begin(), s.end(), s.begin(), ::toupper); return s; }

struct KeyVal {                                            // KeyValPair (KeyValueFileParser.cpp:47-138)
  std::string key, value;
  std::vector<std::string> tokens;
  std::vector<int> ints; std::vector<uint32_t> uints; std::vector<float> floats;
};

template <class T> bool from_string(T& out, const std::string& s) {    // SysTools::FromString
  std::istringstream iss(s);
  return !(iss >> std::dec >> out).fail();
}

struct KeyValueFile {                                      // KeyValueFileParser::ParseFile / ParseKeyValueLine (:230-300)
  std::vector<KeyVal> kv;
  uint64_t stop_pos = 0;                                   // byte position after the empty line that ends an attached header
  bool readable = false;
  KeyValueFile(const std::string& path, bool stop_on_empty) {
    std::ifstream f(path.c_str(), std::ios::binary);
    if (!f.is_open()) return;
    readable = true;
    std::string line;
    while (!f.eof()) {
      std::getline(f, line);
      line = trim_left(line);
      if (!line.empty() && line[line.size() - 1] == 13) line = line.substr(0, line.size() - 1);
      if (stop_on_empty && line.empty()) { if (!f.eof()) stop_pos = (uint64_t)f.tellg(); break; }
      const size_t c = line.find_first_of('#');
      if (c != std::string::npos) line = line.substr(0, c);
      line = trim(line);
      if (line.empty()) continue;
      const size_t t = line.find_first_of(':');
      if (t == std::string::npos) continue;
      KeyVal k;
      k.key = upper(trim_right(line.substr(0, t)));
      k.value = trim(line.substr(t + 1));
      if (k.key.empty() || k.value.empty()) continue;
      std::istringstream ss(k.value);
      std::string tok;
      while (ss >> tok) {
        k.tokens.push_back(tok);
        int i = 0; uint32_t u = 0; float fl = 0;
        k.ints.push_back(from_string(i, tok) ? i : 0);
        k.uints.push_back(from_string(u, tok) ? u : 0);
        k.floats.push_back(from_string(fl, tok) ? fl : 0.0f);
      }
      kv.push_back(k);
    }
  }
  const KeyVal* get(const char* key) const {
    for (const KeyVal& k : kv) if (k.key == key) return &k;
    return nullptr;
  }
};

std::string dir_of(const std::string& path) {              // SysTools::GetPath: up to and including the last separator
  const size_t s = path.find_last_of("/\\");
  return s == std::string::npos ? std::string() : path.substr(0, s + 1);
}

bool host_is_big_endian() { const uint16_t v = 1; return *reinterpret_cast<const uint8_t*>(&v) == 0; }

int fail(char* err, size_t n, const std::string& msg) { if (err && n) snprintf(err, n, "%s", msg.c_str()); return 1; }

int parse_qvis(const std::string& path, tvk_raw_desc& d, char* err, size_t errlen) {          // QVISConverter.cpp:60-150
  KeyValueFile p(path, false);
  if (!p.readable) return fail(err, errlen, "cannot read " + path);
  snprintf(d.title, sizeof(d.title), "Qvis data");
  d.header_skip = 0; d.component_bits = 8; d.component_count = 1; d.is_signed = 0; d.is_float = 0;
  d.convert_endianness = host_is_big_endian() ? 1 : 0;
  const KeyVal* fmt = p.get("FORMAT");
  if (!fmt) return fail(err, errlen, "not a QVIS .dat file (no FORMAT)");
  const std::string F = upper(fmt->value);
  if (F == "CHAR" || F == "UCHAR" || F == "BYTE") { d.component_bits = 8; }                   // QVis cannot store signed 8 bit
  else if (F == "SHORT") { d.is_signed = 1; d.component_bits = 16; }
  else if (F == "USHORT") { d.component_bits = 16; }
  else if (F == "FLOAT") { d.is_signed = 1; d.component_bits = 32; d.is_float = 1; }
  else if (F == "UCHAR4") { d.component_bits = 8; d.component_count = 4; }
  else if (F == "USHORT3") { d.component_bits = 16; d.component_count = 3; }
  else if (F == "USHORT4") { d.component_bits = 16; d.component_count = 4; }
  const KeyVal* obj = p.get("OBJECTFILENAME");
  if (!obj) return fail(err, errlen, "not a valid QVIS .dat file (no ObjectFileName)");
  const KeyVal* res = p.get("RESOLUTION");
  if (!res || res->uints.size() != 3) return fail(err, errlen, "not a valid QVIS .dat file (Resolution)");
  for (int a = 0; a < 3; ++a) d.size[a] = res->uints[a];
  const KeyVal* en = p.get("ENDIANESS");
  if (en && upper(en->value) == "BIG") d.convert_endianness = !d.convert_endianness;
  const KeyVal* st = p.get("SLICETHICKNESS");
  if (!st || st->uints.size() != 3) { d.aspect[0] = d.aspect[1] = d.aspect[2] = 1.0f; }
  else {
    const float m = std::max(st->floats[0], std::max(st->floats[1], st->floats[2]));
    for (int a = 0; a < 3; ++a) d.aspect[a] = st->floats[a] / m;
  }
  snprintf(d.data_file, sizeof(d.data_file), "%s%s", dir_of(path).c_str(), obj->value.c_str());
  d.encoding = TVK_RAW_ENC_RAW;
  return 0;
}

int parse_nrrd(const std::string& path, tvk_raw_desc& d, char* err, size_t errlen) {          // NRRDConverter.cpp:60-440
  {
    std::ifstream f(path.c_str());
    std::string first;
    if (!f.is_open()) return fail(err, errlen, "cannot open " + path);
    std::getline(f, first);
    if (first.substr(0, 7) != "NRRD000") return fail(err, errlen, "not a NRRD file (missing magic)");
  }
  snprintf(d.title, sizeof(d.title), "NRRD data");
  d.component_bits = 8; d.component_count = 1; d.is_signed = 0; d.is_float = 0;
  d.size[0] = d.size[1] = d.size[2] = 1;
  d.aspect[0] = d.aspect[1] = d.aspect[2] = 1.0f;
  KeyValueFile p(path, true);
  if (!p.readable) return fail(err, errlen, "cannot read " + path);
  const KeyVal* ty = p.get("TYPE");
  if (!ty) return fail(err, errlen, "NRRD: no \"type\"");
  const std::string T = upper(ty->value);
  auto any = [&](std::initializer_list<const char*> l) { for (const char* s : l) if (T == s) return true; return false; };
  if (any({"SIGNED CHAR", "INT8", "INT8_T"})) { d.is_signed = 1; d.component_bits = 8; }
  else if (any({"UCHAR", "UNSIGNED CHAR", "UINT8", "UINT8_T"})) { d.component_bits = 8; }
  else if (any({"SHORT", "SHORT INT", "SIGNED SHORT", "SIGNED SHORT INT", "INT16", "INT16_T"})) { d.is_signed = 1; d.component_bits = 16; }
  else if (any({"USHORT", "UNSIGNED SHORT", "UNSIGNED SHORT INT", "UINT16", "UINT16_T"})) { d.component_bits = 16; }
  else if (any({"INT", "SIGNED INT", "INT32", "INT32_T"})) { d.is_signed = 1; d.component_bits = 32; }
  else if (any({"UINT", "UNSIGNED INT", "UINT32", "UINT32_T"})) { d.component_bits = 32; }
  else if (any({"LONGLONG", "LONG LONG", "LONG LONG INT", "SIGNED LONG LONG", "SIGNED LONG LONG INT", "INT64", "INT64_T"})) { d.is_signed = 1; d.component_bits = 64; }
  // (the reference flags the unsigned 64-bit types as signed too, NRRDConverter.cpp:172-179: kept)
  else if (any({"ULONGLONG", "UNSIGNED LONG LONG", "UNSIGNED LONG LONG INT", "UINT64", "UINT64_T"})) { d.is_signed = 1; d.component_bits = 64; }
  else if (any({"FLOAT", "FLOAT32"})) { d.is_signed = 1; d.is_float = 1; d.component_bits = 32; }
  else if (any({"DOUBLE", "FLOAT64"})) { d.is_signed = 1; d.is_float = 1; d.component_bits = 64; }
  else return fail(err, errlen, "NRRD: unsupported \"type\" " + ty->value);
  const KeyVal* sz = p.get("SIZES");
  if (!sz) return fail(err, errlen, "NRRD: no \"sizes\"");
  {
    size_t j = 0;
    for (size_t i = 0; i < sz->uints.size(); ++i)
      if (sz->uints[i] > 1) { if (j > 2) return fail(err, errlen, "only 3D NRRDs are supported"); d.size[j++] = sz->uints[i]; }
  }
  if (!p.get("DIMENSION")) return fail(err, errlen, "NRRD: no \"dimension\"");
  const KeyVal* df1 = p.get("DATA FILE");
  const KeyVal* df2 = p.get("DATAFILE");
  bool detached;
  std::string raw;
  if (!df1 && !df2) {
    d.header_skip = p.stop_pos; raw = path; detached = false;
    if (d.header_skip == 0) return fail(err, errlen, "NRRD: neither a detached data file nor data after the header");
  } else {
    if (df1) raw = dir_of(path) + df1->value;
    if (df2) raw = dir_of(path) + df2->value;
    d.header_skip = 0; detached = true;
  }
  if (const KeyVal* sp = p.get("SPACINGS")) {
    size_t j = 0;
    for (size_t i = 0; i < sz->uints.size(); ++i) {
      if (sp->floats.size() <= i) break;
      if (sz->uints[i] > 1 && j < 3) d.aspect[j++] = sp->floats[i];
    }
  }
  // "space directions": each axis scales the aspect by its length (NRRDConverter.cpp:262-300).  The scan is the
  // reference's, including its substr length 1 + iStart - iEnd: for a number of two or more characters that wraps to
  // "the rest of the token" (the stream extraction then reads the number), for a ONE-character number it is 0 -- an
  // empty string, which a failed extraction turns into 0 (C++11).  A malformed token stops the scan where it is.
  if (const KeyVal* sd = p.get("SPACE DIRECTIONS")) {
    if (sd->tokens.size() == 3) {
      auto numch = [](char c) { return isdigit((unsigned char)c) || c == '.' || c == '-' || c == 'e'; };
      bool broken = false;
      for (size_t dim = 0; dim < 3 && !broken; ++dim) {
        const std::string& t = sd->tokens[dim];
        float axis[3] = {0, 0, 0};
        size_t iStart = 0;
        for (size_t v = 0; v < 3 && !broken; ++v) {
          while (!numch(t[iStart])) { if (++iStart >= t.length()) { broken = true; break; } }
          if (broken) break;
          size_t iEnd = iStart + 1;
          while (numch(t[iEnd])) { if (++iEnd >= t.length()) { broken = true; break; } }
          if (broken) break;
          float val = 0.0f;
          std::istringstream iss(t.substr(iStart, 1 + iStart - iEnd));
          iss >> val;
          axis[v] = val;
          iStart = iEnd;
        }
        if (!broken) d.aspect[dim] *= std::sqrt(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2]);
      }
    }
  }
  for (int a = 0; a < 3; ++a) if (d.aspect[a] == 0.0f) d.aspect[a] = 1.0f;
  int line_skip = 0, byte_skip = 0;
  if (const KeyVal* k = p.get("LINE SKIP")) line_skip = k->ints.empty() ? 0 : k->ints[0];
  if (const KeyVal* k = p.get("LINESKIP")) line_skip = k->ints.empty() ? 0 : k->ints[0];
  if (const KeyVal* k = p.get("BYTE SKIP")) byte_skip = k->ints.empty() ? 0 : k->ints[0];
  if (const KeyVal* k = p.get("BYTESKIP")) byte_skip = k->ints.empty() ? 0 : k->ints[0];
  if (line_skip < 0) line_skip = 0;
  if (byte_skip == -1 && line_skip != 0) line_skip = 0;
  uint64_t line_skip_bytes = 0;
  if (line_skip != 0) {
    std::ifstream f(raw.c_str(), std::ios::binary);
    std::string line;
    int i = 0;
    while (f.is_open() && !f.eof()) { std::getline(f, line); if (++i == line_skip) { line_skip_bytes = (uint64_t)f.tellg(); break; } }
  }
  bool big = false;
  if (const KeyVal* k = p.get("ENDIAN")) big = upper(k->value) == "BIG";
  d.convert_endianness = big != host_is_big_endian() ? 1 : 0;
  const KeyVal* enc = p.get("ENCODING");
  if (!enc) return fail(err, errlen, "NRRD: no \"encoding\"");
  const std::string E = upper(enc->value);
  snprintf(d.data_file, sizeof(d.data_file), "%s", raw.c_str());
  const uint64_t payload = (uint64_t)d.component_bits / 8 * d.component_count * d.size[0] * d.size[1] * d.size[2];
  if (E == "RAW") {
    d.encoding = TVK_RAW_ENC_RAW;
    if (byte_skip == -1) {
      std::ifstream f(raw.c_str(), std::ios::binary | std::ios::ate);
      if (!f.is_open()) return fail(err, errlen, "cannot open " + raw);
      d.header_skip = (uint64_t)f.tellg() - payload;
    } else if (byte_skip != 0) {
      if (detached) d.header_skip = (uint64_t)byte_skip; else d.header_skip += (uint64_t)byte_skip;
    }
    d.header_skip += line_skip_bytes;
    return 0;
  }
  if (E == "GZIP" || E == "GZ") {
    // the reference inflates into a temporary file (NRRDConverter.cpp:407-425: header skip + line skip, then byte skip
    // INSIDE the inflated stream); tvk_raw_read inflates in memory with the same skips
    d.encoding = TVK_RAW_ENC_GZIP;
    d.header_skip += line_skip_bytes;
    d.inflated_skip = byte_skip > 0 ? (uint64_t)byte_skip : 0;
    return 0;
  }
  return fail(err, errlen, "NRRD encoding \"" + enc->value + "\" is not supported (raw and gzip are)");
}

}  // namespace

extern "C" int tvk_raw_probe(const char* path, tvk_raw_desc* out, char* err, size_t errlen) {
  if (!path || !out) return 1;
  memset(out, 0, sizeof(*out));
  const std::string p(path);
  const size_t dot = p.find_last_of('.');
  const std::string ext = dot == std::string::npos ? std::string() : upper(p.substr(dot + 1));
  if (ext == "DAT") return parse_qvis(p, *out, err, errlen);
  if (ext == "NRRD" || ext == "NHDR") return parse_nrrd(p, *out, err, errlen);
  return fail(err, errlen, "unknown volume file type ." + ext + " (QVIS .dat and NRRD .nrrd/.nhdr are read)");
}

extern "C" int tvk_raw_read(const tvk_raw_desc* d, void* dst, uint64_t dst_bytes, char* err, size_t errlen) {
  if (!d || !dst) return 1;
  const uint64_t n = (uint64_t)d->component_bits / 8 * d->component_count * d->size[0] * d->size[1] * d->size[2];
  if (dst_bytes < n) return fail(err, errlen, "destination too small");
  FILE* f = fopen(d->data_file, "rb");
  if (!f) return fail(err, errlen, std::string("cannot open ") + d->data_file);
  bool ok = fseeko(f, (off_t)d->header_skip, SEEK_SET) == 0;
  if (ok && d->encoding == TVK_RAW_ENC_RAW) {
    ok = fread(dst, 1, (size_t)n, f) == (size_t)n;
  } else if (ok) {
    std::vector<uint8_t> in;
    uint8_t buf[1 << 16];
    size_t got;
    while ((got = fread(buf, 1, sizeof(buf), f)) > 0) in.insert(in.end(), buf, buf + got);
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    ok = inflateInit2(&zs, 15 + 32) == Z_OK;               // gzip or zlib wrapper, auto-detected
    if (ok) {
      std::vector<uint8_t> skipbuf((size_t)d->inflated_skip);
      zs.next_in = in.data(); zs.avail_in = (uInt)in.size();
      uint64_t done = 0;
      int rc = Z_OK;
      // first the bytes to skip inside the stream, then the payload
      for (int pass = 0; pass < 2 && ok; ++pass) {
        uint8_t* o = pass == 0 ? skipbuf.data() : static_cast<uint8_t*>(dst);
        const uint64_t want = pass == 0 ? d->inflated_skip : n;
        done = 0;
        while (done < want && rc != Z_STREAM_END) {
          const uInt chunk = (uInt)std::min<uint64_t>(want - done, 1u << 30);
          zs.next_out = o + done; zs.avail_out = chunk;
          rc = inflate(&zs, Z_NO_FLUSH);
          if (rc != Z_OK && rc != Z_STREAM_END) { ok = false; break; }
          done += chunk - zs.avail_out;
          if (zs.avail_in == 0 && zs.avail_out != 0 && rc != Z_STREAM_END) { ok = false; break; }
        }
        if (done != want) ok = false;
      }
      inflateEnd(&zs);
    }
  }
  fclose(f);
  if (!ok) return fail(err, errlen, std::string("short or corrupt data in ") + d->data_file);
  if (d->convert_endianness && d->component_bits > 8) {
    const size_t w = d->component_bits / 8;
    uint8_t* b = static_cast<uint8_t*>(dst);
    for (uint64_t i = 0; i + w <= n; i += w) std::reverse(b + i, b + i + w);
  }
  return 0;
}
