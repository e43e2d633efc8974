// TEST INFRASTRUCTURE ONLY -- never linked into the product.
//
// Small command line driver around the UNMODIFIED reference IO/builder sources
// (compiled in place from /root/reference by oracle/ref_build/Makefile into
// oracle/_ref/).  It is used (a) to pin the C restatement in oracle/ against the
// real reference at arbitrary sizes and (b) to generate the golden fixtures in
// tests/golden/ (see tests/golden/make_golden.py).
//
//   ref_driver convert <in.raw> <sx> <sy> <sz> <u8|u16> <brick> <overlap>
//                      <median 0|1> <clamp 0|1> <out.bin> [hist 0|1]
//     runs TOCBlock::FlatDataToBrickedLOD (-> ExtendedOctreeConverter::Convert,
//     Tuvok/IO/UVF/TOCBlock.cpp:135, ExtendedOctreeConverter.cpp:128), then
//     Histogram1DDataBlock::Compute / Histogram2DDataBlock::Compute and dumps
//     everything in a flat little-endian format (see dump layout below).
//   ref_driver uvfinfo <file.uvf> <out.bin>
//     opens a UVF (MD5 verified) and dumps the first RasterDataBlock brick, the
//     1D histogram and the global min/max (used for DemoData/c60.uvf).
//   ref_driver tf1d <file.1dt> <out.bin>   dumps size + GetByteArray (RGBA8)
//   ref_driver quantize <in.raw> <i8|u8|i16|u16|i32|u32|f32|i64|u64|f64> <elements> <to8bit 0|1> <out.bin>
//     calls the reference's Quantize<T,U> / BinningQuantize<T,U> / AbstrConverter::Process8Bits /
//     QuantizeTo8Bit templates (IO/Quantize.h, IO/AbstrConverter.cpp) in the order of the file-static
//     dispatcher quantize() of IO/RAWConverter.cpp:196-352 (one component, one time step) and dumps
//     u32 outBits; u32 generated; u64 nbytes; data; u64 nh; u64 hist[nh]
//
// Dump layout of `convert` (all little endian):
//   char[8] "TVKREF1\0"; u32 lodCount;
//   per LOD: u64 lodSize[3]; u64 brickCount[3]; f64 aspect[3];
//   u64 totalBricks; per brick (ToC order): u64 size[3]; f64 min; f64 max;
//   then per brick: raw voxels (size.volume()*bytesPerVoxel);
//   u64 n1; u64 hist1d[n1]; f32 maxGrad; u64 nx; u64 ny; u64 hist2d[nx*ny]
#include <cstdio>
#include <list>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <memory>
#include <chrono>

#include "Controller/Controller.h"
#include "IO/UVF/UVF.h"
#include "IO/UVF/TOCBlock.h"
#include "IO/UVF/MaxMinDataBlock.h"
#include "IO/UVF/Histogram1DDataBlock.h"
#include "IO/UVF/Histogram2DDataBlock.h"
#include "IO/UVF/RasterDataBlock.h"
#include "IO/TransferFunction1D.h"
#include "IO/TransferFunction2D.h"
#include "IO/QVISConverter.h"
#include "IO/NRRDConverter.h"
#include "Basics/SysTools.h"
#include "Basics/LargeRAWFile.h"
#include "IO/AbstrConverter.h"
#include "IO/Quantize.h"
#include "IO/UVF/KeyValuePairDataBlock.h"
#include "Basics/nonstd.h"
#include "IO/uvfDataset.h"
#include "IO/DynamicBrickingDS.h"

using namespace tuvok;

template <typename T> static void put(FILE* f, const T& v) { fwrite(&v, sizeof(T), 1, f); }

// header + per-brick sizes / min / max + brick payloads of a TOC block (TVKREF1 layout, see above)
static void dump_octree(FILE* f, const TOCBlock& toc, const MaxMinDataBlock& mmb, size_t bpv, uint64_t brick) {
  const MaxMinDataBlock* mm = &mmb;
  fwrite("TVKREF1\0", 1, 8, f);
  const uint32_t lods = (uint32_t)toc.GetLoDCount();
  put(f, lods);
  uint64_t total = 0;
  for (uint32_t l = 0; l < lods; ++l) {
    UINT64VECTOR3 s = toc.GetLODDomainSize(l), b = toc.GetBrickCount(l);
    DOUBLEVECTOR3 a = toc.GetBrickAspect(UINT64VECTOR4(0, 0, 0, l));
    put(f, s.x); put(f, s.y); put(f, s.z);
    put(f, b.x); put(f, b.y); put(f, b.z);
    put(f, a.x); put(f, a.y); put(f, a.z);
    total += b.volume();
  }
  put(f, total);
  size_t idx = 0;
  for (uint32_t l = 0; l < lods; ++l) {
    UINT64VECTOR3 b = toc.GetBrickCount(l);
    for (uint64_t z = 0; z < b.z; ++z) for (uint64_t y = 0; y < b.y; ++y) for (uint64_t x = 0; x < b.x; ++x) {
      UINT64VECTOR3 s = toc.GetBrickSize(UINT64VECTOR4(x, y, z, l));
      put(f, s.x); put(f, s.y); put(f, s.z);
      const MinMaxBlock& v = mm->GetValue(idx++, 0);
      put(f, v.minScalar); put(f, v.maxScalar);
    }
  }
  std::vector<uint8_t> buf(brick * brick * brick * bpv);
  for (uint32_t l = 0; l < lods; ++l) {
    UINT64VECTOR3 b = toc.GetBrickCount(l);
    for (uint64_t z = 0; z < b.z; ++z) for (uint64_t y = 0; y < b.y; ++y) for (uint64_t x = 0; x < b.x; ++x) {
      UINT64VECTOR4 c(x, y, z, l);
      UINT64VECTOR3 s = toc.GetBrickSize(c);
      toc.GetData(buf.data(), c);
      fwrite(buf.data(), 1, size_t(s.volume()) * bpv, f);
    }
  }
}

static int do_convert(int argc, char** argv) {
  if (argc < 12) { fprintf(stderr, "convert: not enough arguments\n"); return 2; }
  const std::string in = argv[2];
  const uint64_t sx = strtoull(argv[3], 0, 10), sy = strtoull(argv[4], 0, 10), sz = strtoull(argv[5], 0, 10);
  const bool u16 = !strcmp(argv[6], "u16");
  const uint64_t brick = strtoull(argv[7], 0, 10);
  const uint32_t overlap = (uint32_t)strtoul(argv[8], 0, 10);
  const bool median = atoi(argv[9]) != 0, clamp = atoi(argv[10]) != 0;
  const std::string out = argv[11];
  const bool hist = (argc > 12) ? atoi(argv[12]) != 0 : true;
  const std::string tmp = out + ".octree.tmp";

  LargeRAWFile_ptr src(new LargeRAWFile(in));
  if (!src->Open(false)) { fprintf(stderr, "cannot open %s\n", in.c_str()); return 1; }

  std::shared_ptr<MaxMinDataBlock> mm(new MaxMinDataBlock(1));
  TOCBlock toc(UVF::ms_ulReaderVersion);
  auto t0 = std::chrono::steady_clock::now();
  bool ok = toc.FlatDataToBrickedLOD(src, tmp,
      u16 ? ExtendedOctree::CT_UINT16 : ExtendedOctree::CT_UINT8, 1,
      UINT64VECTOR3(sx, sy, sz), DOUBLEVECTOR3(1, 1, 1),
      UINT64VECTOR3(brick, brick, brick), overlap, median, clamp,
      size_t(1) << 31, mm, &Controller::Debug::Out(), CT_NONE, 0, LT_SCANLINE);
  auto t1 = std::chrono::steady_clock::now();
  if (!ok) { fprintf(stderr, "conversion failed\n"); return 1; }
  fprintf(stderr, "REF_CONVERT_SECONDS %.6f\n", std::chrono::duration<double>(t1 - t0).count());

  FILE* f = fopen(out.c_str(), "wb");
  if (!f) return 1;
  dump_octree(f, toc, *mm, u16 ? 2 : 1, brick);
  if (hist) {
    Histogram1DDataBlock h1;
    auto t2 = std::chrono::steady_clock::now();
    h1.Compute(&toc, 0);
    auto t3 = std::chrono::steady_clock::now();
    const std::vector<uint64_t>& bins = h1.GetHistogram();
    uint64_t n1 = bins.size();
    put(f, n1);
    fwrite(bins.data(), 8, bins.size(), f);
    Histogram2DDataBlock h2;
    h2.Compute(&toc, 0, bins.size(), mm->GetGlobalValue().maxScalar);
    auto t4 = std::chrono::steady_clock::now();
    fprintf(stderr, "REF_HIST1D_SECONDS %.6f\nREF_HIST2D_SECONDS %.6f\n",
            std::chrono::duration<double>(t3 - t2).count(), std::chrono::duration<double>(t4 - t3).count());
    float mg = h2.GetMaxGradMagnitude();
    put(f, mg);
    const std::vector<std::vector<uint64_t>>& b2 = h2.GetHistogram();
    uint64_t nx = b2.size(), ny = nx ? b2[0].size() : 0;
    put(f, nx); put(f, ny);
    for (uint64_t i = 0; i < nx; ++i) fwrite(b2[i].data(), 8, ny, f);
  } else {
    uint64_t z = 0; float zf = 0; put(f, z); put(f, zf); put(f, z); put(f, z);
  }
  fclose(f);
  remove(tmp.c_str());
  return 0;
}

static int do_uvfinfo(int argc, char** argv) {
  if (argc < 4) return 2;
  std::string fn(argv[2]);
  UVF uvf(std::wstring(fn.begin(), fn.end()));
  std::string err;
  if (!uvf.Open(false, true, false, &err)) { fprintf(stderr, "open failed: %s\n", err.c_str()); return 1; }
  FILE* f = fopen(argv[3], "wb");
  fwrite("TVKUVF1\0", 1, 8, f);
  const RasterDataBlock* rdb = NULL; const Histogram1DDataBlock* h1 = NULL; const MaxMinDataBlock* mm = NULL;
  const Histogram2DDataBlock* h2 = NULL;
  for (uint64_t i = 0; i < uvf.GetDataBlockCount(); ++i) {
    const DataBlock* b = uvf.GetDataBlock(i).get();
    if (!rdb) rdb = dynamic_cast<const RasterDataBlock*>(b);
    if (!h1) h1 = dynamic_cast<const Histogram1DDataBlock*>(b);
    if (!h2) h2 = dynamic_cast<const Histogram2DDataBlock*>(b);
    if (!mm) mm = dynamic_cast<const MaxMinDataBlock*>(b);
  }
  if (!rdb || !h1 || !mm) { fprintf(stderr, "missing blocks\n"); return 1; }
  std::vector<uint64_t> lod(1, 0), brick(3, 0);
  std::vector<uint64_t> bs = rdb->GetBrickSize(lod, brick);
  uint64_t dom[3] = {rdb->ulDomainSize[0], rdb->ulDomainSize[1], rdb->ulDomainSize[2]};
  fwrite(dom, 8, 3, f);
  uint64_t bsz[3] = {bs[0], bs[1], bs[2]};
  fwrite(bsz, 8, 3, f);
  uint64_t bits = rdb->ulElementBitSize[0][0];
  put(f, bits);
  std::vector<unsigned char> data;
  rdb->GetData(data, lod, brick);
  uint64_t n = data.size(); put(f, n);
  fwrite(data.data(), 1, data.size(), f);
  const std::vector<uint64_t>& bins = h1->GetHistogram();
  uint64_t n1 = bins.size(); put(f, n1); fwrite(bins.data(), 8, n1, f);
  double mn = mm->GetGlobalValue().minScalar, mx = mm->GetGlobalValue().maxScalar;
  put(f, mn); put(f, mx);
  float mg = h2 ? h2->GetMaxGradMagnitude() : 0.0f; put(f, mg);
  fclose(f);
  uvf.Close();
  return 0;
}

static int do_tf1d(int argc, char** argv) {
  if (argc < 4) return 2;
  const std::string tfn(argv[2]);
  TransferFunction1D tf(tfn);
  std::vector<unsigned char> bytes;
  tf.GetByteArray(bytes);
  FILE* f = fopen(argv[3], "wb");
  uint64_t n = tf.GetSize(); put(f, n);
  fwrite(bytes.data(), 1, bytes.size(), f);
  uint64_t lo = tf.GetNonZeroLimits().x; uint64_t hi = tf.GetNonZeroLimits().y;
  put(f, lo); put(f, hi);
  fclose(f);
  return 0;
}

// ref_driver rawinfo <file.dat | file.nrrd | file.nhdr>: what the reference's converter reports for the file
// (QVISConverter / NRRDConverter::ConvertToRAW), one "key value" per line
static int do_rawinfo(int argc, char** argv) {
  if (argc < 3) return 2;
  const std::string fn(argv[2]);
  const std::string ext = SysTools::ToUpperCase(SysTools::GetExt(fn));
  uint64_t skip = 0; unsigned bits = 0; uint64_t comps = 0; bool conv = false, sgn = false, flt = false, del = false;
  UINT64VECTOR3 size; FLOATVECTOR3 aspect; std::string title, inter;
  bool ok;
  if (ext == "DAT") { QVISConverter c; ok = c.ConvertToRAW(fn, "/tmp/", true, skip, bits, comps, conv, sgn, flt, size, aspect, title, inter, del); }
  else { NRRDConverter c; ok = c.ConvertToRAW(fn, "/tmp/", true, skip, bits, comps, conv, sgn, flt, size, aspect, title, inter, del); }
  if (!ok) { printf("ok 0\n"); return 0; }
  printf("ok 1\nskip %llu\nbits %u\ncomps %llu\nconv %d\nsigned %d\nfloat %d\nsize %llu %llu %llu\naspect %.9g %.9g %.9g\nfile %s\ndelete %d\n",
         (unsigned long long)skip, bits, (unsigned long long)comps, (int)conv, (int)sgn, (int)flt, (unsigned long long)size.x,
         (unsigned long long)size.y, (unsigned long long)size.z, aspect.x, aspect.y, aspect.z, inter.c_str(), (int)del);
  return 0;
}

// ref_driver tfresample <in.f32 (n*4 floats)> <n> <target> <out.f32>: TransferFunction1D::Resample on caller data
static int do_tfresample(int argc, char** argv) {
  if (argc < 6) return 2;
  const size_t n = strtoull(argv[3], 0, 10), target = strtoull(argv[4], 0, 10);
  TransferFunction1D tf(n);
  std::vector<FLOATVECTOR4>* data = tf.GetColorData().get();
  FILE* f = fopen(argv[2], "rb");
  if (!f) return 1;
  for (size_t i = 0; i < n; ++i) { float v[4]; if (fread(v, 4, 4, f) != 4) return 1; (*data)[i] = FLOATVECTOR4(v[0], v[1], v[2], v[3]); }
  fclose(f);
  tf.Resample(target);
  data = tf.GetColorData().get();
  f = fopen(argv[5], "wb");
  for (size_t i = 0; i < data->size(); ++i) { float v[4] = {(*data)[i].x, (*data)[i].y, (*data)[i].z, (*data)[i].w}; fwrite(v, 4, 4, f); }
  fclose(f);
  return 0;
}

// ref_driver writeuvf <in.raw> <sx> <sy> <sz> <u8|u16> <brick> <overlap> <median> <compression> <level> <out.uvf>
//   the file assembly of RAWConverter::ConvertRAWDataset (IO/RAWConverter.cpp:527-681) for an input that
//   needs no quantisation: TOC block, 1-D / 2-D histogram, max/min, key/value block, MD5.
static int do_writeuvf(int argc, char** argv) {
  if (argc < 13) return 2;
  const std::string in = argv[2], out = argv[12];
  const uint64_t sx = strtoull(argv[3], 0, 10), sy = strtoull(argv[4], 0, 10), sz = strtoull(argv[5], 0, 10);
  const bool u16 = !strcmp(argv[6], "u16");
  const uint64_t brick = strtoull(argv[7], 0, 10);
  const uint32_t overlap = (uint32_t)strtoul(argv[8], 0, 10);
  const bool median = atoi(argv[9]) != 0;
  const uint32_t comp = (uint32_t)atoi(argv[10]), level = (uint32_t)atoi(argv[11]);
  const std::string tmp = out + ".octree.tmp";
  LargeRAWFile_ptr src(new LargeRAWFile(in));
  if (!src->Open(false)) return 1;
  UVF uvfFile(std::wstring(out.begin(), out.end()));
  GlobalHeader gh;
  gh.bIsBigEndian = EndianConvert::IsBigEndian();
  gh.ulChecksumSemanticsEntry = UVFTables::CS_MD5;
  uvfFile.SetGlobalHeader(gh);
  std::shared_ptr<MaxMinDataBlock> mm(new MaxMinDataBlock(1));
  std::shared_ptr<Histogram2DDataBlock> h2(new Histogram2DDataBlock());
  std::shared_ptr<TOCBlock> toc(new TOCBlock(UVF::ms_ulReaderVersion));
  toc->strBlockID = "Volume converted by ImageVis3D";
  if (!toc->FlatDataToBrickedLOD(src, tmp, u16 ? ExtendedOctree::CT_UINT16 : ExtendedOctree::CT_UINT8, 1,
                                 UINT64VECTOR3(sx, sy, sz), DOUBLEVECTOR3(1, 1, 1), UINT64VECTOR3(brick, brick, brick), overlap,
                                 median, false, size_t(1) << 31, mm, &Controller::Debug::Out(), COMPRESSION_TYPE(comp), level,
                                 LT_SCANLINE)) return 1;
  if (!uvfFile.AddDataBlock(toc)) return 1;
  Histogram1DDataBlock h1;
  if (!h1.Compute(toc.get(), 0)) return 1;
  if (!h2->Compute(toc.get(), 0, h1.GetHistogram().size(), mm->GetGlobalValue().maxScalar)) return 1;
  uvfFile.AddDataBlock(std::shared_ptr<Histogram1DDataBlock>(&h1, nonstd::null_deleter()));
  uvfFile.AddDataBlock(h2);
  uvfFile.AddDataBlock(mm);
  std::shared_ptr<KeyValuePairDataBlock> meta(new KeyValuePairDataBlock());
  meta->AddPair("Source Endianness", "little");
  meta->AddPair("Source Type", "integer");
  meta->AddPair("Source Bitwidth", u16 ? "16" : "8");
  uvfFile.AddDataBlock(meta);
  uvfFile.Create();
  uvfFile.Close();
  src->Close();
  return 0;
}

// ref_driver readuvf <file.uvf> <out.bin>: UVF::Open with checksum verification, then the TVKREF1 dump of
// the first TOC block with the file's own histogram / max-min blocks, followed by u64 nkv and, per pair,
// u64 len, key, u64 len, value.

// ref_driver writeuvfts <out.uvf> <sx> <sy> <sz> <u8|u16> <brick> <overlap> <raw0> <raw1> ...
//   a TIME-DEPENDENT file, block order as the time-step loop of RAWConverter::ConvertRAWDataset writes it
//   (IO/RAWConverter.cpp:548-661): per step a TOC block, the 1-D histogram block (the loop computes it for the first step
//   only and stores the same object again for every later step, :630-637), a 2-D histogram and a max/min block; the
//   key/value block once at the end.
static int do_writeuvfts(int argc, char** argv) {
  if (argc < 10) return 2;
  const std::string out = argv[2];
  const uint64_t sx = strtoull(argv[3], 0, 10), sy = strtoull(argv[4], 0, 10), sz = strtoull(argv[5], 0, 10);
  const bool u16 = !strcmp(argv[6], "u16");
  const uint64_t brick = strtoull(argv[7], 0, 10);
  const uint32_t overlap = (uint32_t)strtoul(argv[8], 0, 10);
  UVF uvfFile(std::wstring(out.begin(), out.end()));
  GlobalHeader gh;
  gh.bIsBigEndian = EndianConvert::IsBigEndian();
  gh.ulChecksumSemanticsEntry = UVFTables::CS_MD5;
  uvfFile.SetGlobalHeader(gh);
  Histogram1DDataBlock h1;
  std::vector<std::shared_ptr<TOCBlock>> keep;
  for (int t = 9; t < argc; ++t) {
    LargeRAWFile_ptr src(new LargeRAWFile(std::string(argv[t])));
    if (!src->Open(false)) return 1;
    std::shared_ptr<MaxMinDataBlock> mm(new MaxMinDataBlock(1));
    std::shared_ptr<Histogram2DDataBlock> h2(new Histogram2DDataBlock());
    std::shared_ptr<TOCBlock> toc(new TOCBlock(UVF::ms_ulReaderVersion));
    toc->strBlockID = "Volume converted by ImageVis3D";
    const std::string tmp = out + "." + std::to_string(t) + ".octree.tmp";
    if (!toc->FlatDataToBrickedLOD(src, tmp, u16 ? ExtendedOctree::CT_UINT16 : ExtendedOctree::CT_UINT8, 1,
                                   UINT64VECTOR3(sx, sy, sz), DOUBLEVECTOR3(1, 1, 1), UINT64VECTOR3(brick, brick, brick), overlap,
                                   false, false, size_t(1) << 31, mm, &Controller::Debug::Out(), CT_NONE, 1, LT_SCANLINE)) return 1;
    if (!uvfFile.AddDataBlock(toc)) return 1;
    keep.push_back(toc);
    if (h1.GetHistogram().empty() && !h1.Compute(toc.get(), 0)) return 1;
    if (!h2->Compute(toc.get(), 0, h1.GetHistogram().size(), mm->GetGlobalValue().maxScalar)) return 1;
    uvfFile.AddDataBlock(std::shared_ptr<Histogram1DDataBlock>(&h1, nonstd::null_deleter()));
    uvfFile.AddDataBlock(h2);
    uvfFile.AddDataBlock(mm);
    src->Close();
  }
  std::shared_ptr<KeyValuePairDataBlock> meta(new KeyValuePairDataBlock());
  meta->AddPair("Source Type", "integer");
  uvfFile.AddDataBlock(meta);
  uvfFile.Create();
  uvfFile.Close();
  return 0;
}

// ref_driver dsdump <file.uvf> <timestep> <out.bin>: the file through the reference's UVFDataset (uvfDataset.cpp)
//     char[8] "TVKDSD1\0"; u64 timesteps; u32 lods; u64 total; per brick (LOD-major, index order): u32 size[3];
//     f64 min; f64 max (MaxMinForKey); then the payloads (GetBrick); u64 n1; u64 hist1d[n1] (Get1DHistogram of step 0 --
//     the dataset only ever exposes m_timesteps[0]'s histograms, uvfDataset.cpp:747-750); f32 maxGradMagnitude(timestep)
static int do_dsdump(int argc, char** argv) {
  if (argc < 5) return 2;
  UVFDataset ds(argv[2], 1024, false);
  const size_t ts = (size_t)atoi(argv[3]);
  FILE* f = fopen(argv[4], "wb");
  if (!f) return 1;
  fwrite("TVKDSD1\0", 1, 8, f);
  put(f, (uint64_t)ds.GetNumberOfTimesteps());
  const uint32_t lods = (uint32_t)ds.GetLODLevelCount();
  put(f, lods);
  std::vector<uint64_t> per_lod(lods);
  uint64_t total = 0;
  for (uint32_t l = 0; l < lods; ++l) { const UINTVECTOR3 lay = ds.GetBrickLayout(l, ts); per_lod[l] = (uint64_t)lay[0] * lay[1] * lay[2]; total += per_lod[l]; }
  put(f, total);
  for (uint32_t l = 0; l < lods; ++l)
    for (uint64_t i = 0; i < per_lod[l]; ++i) {
      const BrickKey k(ts, l, i);
      const UINTVECTOR3 n = ds.GetBrickVoxelCounts(k);
      put(f, (uint32_t)n[0]); put(f, (uint32_t)n[1]); put(f, (uint32_t)n[2]);
      const MinMaxBlock mm = ds.MaxMinForKey(k);
      put(f, mm.minScalar); put(f, mm.maxScalar);
    }
  std::vector<uint8_t> d8; std::vector<uint16_t> d16;
  const bool wide = ds.GetBitWidth() == 16;
  for (uint32_t l = 0; l < lods; ++l)
    for (uint64_t i = 0; i < per_lod[l]; ++i) {
      const BrickKey k(ts, l, i);
      if (!(wide ? ds.GetBrick(k, d16) : ds.GetBrick(k, d8))) { fclose(f); return 1; }
      if (wide) fwrite(d16.data(), 2, d16.size(), f); else fwrite(d8.data(), 1, d8.size(), f);
    }
  fclose(f);
  return 0;
}

// ref_driver tf2d <in.2dt> <out.2dt>: TransferFunction2D::Load(file) + Save(file) of the reference (the text round trip;
//   the swatch rasterisation needs Qt and is compiled out)
static int do_tf2d(int argc, char** argv) {
  if (argc < 4) return 2;
  TransferFunction2D tf;
  if (!tf.Load(std::string(argv[2]))) return 1;
  return tf.Save(std::string(argv[3])) ? 0 : 1;
}

// ref_driver rebrick <file.uvf> <bx> <by> <bz> <mm 0|1|2> <out.bin>
//   opens the file with the reference's UVFDataset, wraps it in the reference's DynamicBrickingDS (target brick size
//   INCLUDING ghost voxels, min/max mode MM_SOURCE / MM_PRECOMPUTE / MM_DYNAMIC -- IO/DynamicBrickingDS.h) and dumps
//     char[8] "TVKRBK1\0"; u32 lods; per LOD: u32 layout[3], u64 domain[3]; u64 total;
//     per brick (LOD-major, index order): u32 size[3]; f64 min; f64 max   (MaxMinForKey)
//     then per brick: payload (GetBrick)
static int do_rebrick(int argc, char** argv) {
  if (argc < 8) return 2;
  std::shared_ptr<UVFDataset> ds(new UVFDataset(argv[2], 1024, false));
  const unsigned mode = (unsigned)atoi(argv[6]);
  std::array<size_t, 3> bs = {{(size_t)atoi(argv[3]), (size_t)atoi(argv[4]), (size_t)atoi(argv[5])}};
  std::unique_ptr<DynamicBrickingDS> dyn;
  try { dyn.reset(new DynamicBrickingDS(ds, bs, 1u << 30, (DynamicBrickingDS::MinMaxMode)mode)); }
  catch (const std::exception& e) { fprintf(stderr, "rebrick refused: %s\n", e.what()); return 3; }
  FILE* f = fopen(argv[7], "wb");
  if (!f) return 1;
  fwrite("TVKRBK1\0", 1, 8, f);
  const uint32_t lods = (uint32_t)dyn->GetLODLevelCount();
  put(f, lods);
  uint64_t total = 0;
  std::vector<uint64_t> per_lod(lods);
  for (uint32_t l = 0; l < lods; ++l) {
    const UINTVECTOR3 lay = dyn->GetBrickLayout(l, 0);
    const UINT64VECTOR3 dom = dyn->GetDomainSize(l, 0);
    put(f, (uint32_t)lay[0]); put(f, (uint32_t)lay[1]); put(f, (uint32_t)lay[2]);
    put(f, (uint64_t)dom[0]); put(f, (uint64_t)dom[1]); put(f, (uint64_t)dom[2]);
    per_lod[l] = (uint64_t)lay[0] * lay[1] * lay[2];
    total += per_lod[l];
  }
  put(f, total);
  for (uint32_t l = 0; l < lods; ++l)
    for (uint64_t i = 0; i < per_lod[l]; ++i) {
      const BrickKey k(0, l, i);
      const UINTVECTOR3 n = dyn->GetBrickMetadata(k).n_voxels;
      put(f, (uint32_t)n[0]); put(f, (uint32_t)n[1]); put(f, (uint32_t)n[2]);
      const MinMaxBlock mm = dyn->MaxMinForKey(k);
      put(f, mm.minScalar); put(f, mm.maxScalar);
    }
  std::vector<uint8_t> data;
  std::vector<uint16_t> data16;                              // (GetBrick is typed: the 8-bit overload copies 8-bit elements)
  const bool wide = dyn->GetBitWidth() == 16;
  for (uint32_t l = 0; l < lods; ++l)
    for (uint64_t i = 0; i < per_lod[l]; ++i) {
      const BrickKey k(0, l, i);
      if (!(wide ? dyn->GetBrick(k, data16) : dyn->GetBrick(k, data))) { fprintf(stderr, "GetBrick failed\n"); fclose(f); return 1; }
      if (wide) fwrite(data16.data(), 2, data16.size(), f);
      else fwrite(data.data(), 1, data.size(), f);
    }
  fclose(f);
  return 0;
}

static int do_readuvf(int argc, char** argv) {
  if (argc < 4) return 2;
  std::string fn(argv[2]);
  UVF uvf(std::wstring(fn.begin(), fn.end()));
  std::string err;
  if (!uvf.Open(true, true, false, &err)) { fprintf(stderr, "open failed: %s\n", err.c_str()); return 1; }
  const TOCBlock* toc = NULL; const Histogram1DDataBlock* h1 = NULL; const MaxMinDataBlock* mm = NULL;
  const Histogram2DDataBlock* h2 = NULL; const KeyValuePairDataBlock* kv = NULL;
  for (uint64_t i = 0; i < uvf.GetDataBlockCount(); ++i) {
    const DataBlock* b = uvf.GetDataBlock(i).get();
    if (!toc) toc = dynamic_cast<const TOCBlock*>(b);
    if (!h1) h1 = dynamic_cast<const Histogram1DDataBlock*>(b);
    if (!h2) h2 = dynamic_cast<const Histogram2DDataBlock*>(b);
    if (!mm) mm = dynamic_cast<const MaxMinDataBlock*>(b);
    if (!kv) kv = dynamic_cast<const KeyValuePairDataBlock*>(b);
  }
  if (!toc || !h1 || !h2 || !mm || !kv) { fprintf(stderr, "missing blocks\n"); return 1; }
  FILE* f = fopen(argv[3], "wb");
  if (!f) return 1;
  dump_octree(f, *toc, *mm, toc->GetComponentTypeSize(), toc->GetMaxBrickSize().x);
  const std::vector<uint64_t>& bins = h1->GetHistogram();
  uint64_t n1 = bins.size();
  put(f, n1);
  fwrite(bins.data(), 8, bins.size(), f);
  float mg = h2->GetMaxGradMagnitude();
  put(f, mg);
  const std::vector<std::vector<uint64_t>>& b2 = h2->GetHistogram();
  uint64_t nx = b2.size(), ny = nx ? b2[0].size() : 0;
  put(f, nx); put(f, ny);
  for (uint64_t i = 0; i < nx; ++i) fwrite(b2[i].data(), 8, ny, f);
  uint64_t nkv = kv->GetKeyCount();
  put(f, nkv);
  for (uint64_t i = 0; i < nkv; ++i) {
    const std::string k = kv->GetKeyByIndex(i), v = kv->GetValueByIndex(i);
    uint64_t l = k.size(); put(f, l); fwrite(k.data(), 1, k.size(), f);
    l = v.size(); put(f, l); fwrite(v.data(), 1, v.size(), f);
  }
  fclose(f);
  uvf.Close();
  return 0;
}

static int do_quantize(int argc, char** argv) {
  if (argc < 7) { fprintf(stderr, "quantize: not enough arguments\n"); return 2; }
  const std::string in = argv[2], type = argv[3], out = argv[6];
  const uint64_t n = strtoull(argv[4], 0, 10);
  const bool to8 = atoi(argv[5]) != 0;
  const std::string tmp = out + ".quantized.tmp";
  LargeRAWFile src(in);
  src.Open(false);
  if (!src.IsOpen()) return 3;
  const bool fp = type[0] == 'f', sgn = type[0] == 'i' || fp;
  unsigned bits = (unsigned)atoi(type.c_str() + 1);
  Histogram1DDataBlock h;
  BStreamDescriptor bsd;
  bsd.elements = n; bsd.components = 1; bsd.width = bits / 8; bsd.is_signed = sgn; bsd.fp = fp;
  bsd.big_endian = EndianConvert::IsBigEndian(); bsd.timesteps = 1;
  bool target = false;
  if (to8 && bits > 8) {
    target = AbstrConverter::QuantizeTo8Bit(src, tmp, bits, n, sgn, fp, &h);
    bits = 8;
  } else {
    size_t bc = 0;
    switch (bits) {
      case 8: target = AbstrConverter::Process8Bits(src, tmp, n, sgn, &h); break;
      case 16:
        if (sgn) target = Quantize<short, unsigned short>(src, bsd, tmp, &h);
        else {
          target = Quantize<unsigned short, unsigned short>(src, bsd, tmp, &h, &bc);
          if (bc > 0 && bc <= 256) target = BinningQuantize<unsigned short, unsigned char>(src, tmp, n, bits, &h);
        }
        break;
      case 32:
        if (fp) target = BinningQuantize<float, unsigned short>(src, tmp, n, bits, &h);
        else {
          bool binned = false;
          if (sgn) target = Quantize<int32_t, unsigned short>(src, bsd, tmp, &h);
          else {
            target = Quantize<uint32_t, unsigned short>(src, bsd, tmp, &h, &bc);
            if (bc > 0 && bc <= 256) { target = BinningQuantize<uint32_t, unsigned char>(src, tmp, n, bits, &h); binned = true; }
          }
          if (!binned) bits = 16;
        }
        break;
      case 64:
        if (fp) target = BinningQuantize<double, unsigned short>(src, tmp, n, bits, &h);
        else {
          bool binned = false;
          if (sgn) target = Quantize<int64_t, unsigned short>(src, bsd, tmp, &h);
          else {
            target = Quantize<uint64_t, unsigned short>(src, bsd, tmp, &h, &bc);
            if (bc > 0 && bc <= 256) { target = BinningQuantize<uint64_t, unsigned char>(src, tmp, n, bits, &h); binned = true; }
          }
          if (!binned) bits = 16;
        }
        break;
      default: return 2;
    }
  }
  src.Close();
  FILE* f = fopen(out.c_str(), "wb");
  if (!f) return 3;
  put<uint32_t>(f, bits);
  put<uint32_t>(f, target ? 1u : 0u);
  {
    FILE* d = fopen((target ? tmp : in).c_str(), "rb");
    std::vector<unsigned char> buf;
    if (d) {
      fseek(d, 0, SEEK_END); const long sz = ftell(d); fseek(d, 0, SEEK_SET);
      buf.resize((size_t)sz);
      if (sz && fread(buf.data(), 1, (size_t)sz, d) != (size_t)sz) return 3;
      fclose(d);
    }
    put<uint64_t>(f, (uint64_t)buf.size());
    if (!buf.empty()) fwrite(buf.data(), 1, buf.size(), f);
  }
  const std::vector<uint64_t>& hist = h.GetHistogram();
  put<uint64_t>(f, (uint64_t)hist.size());
  if (!hist.empty()) fwrite(hist.data(), sizeof(uint64_t), hist.size(), f);
  fclose(f);
  remove(tmp.c_str());
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: ref_driver convert|uvfinfo|tf1d ...\n"); return 2; }
  if (!strcmp(argv[1], "convert")) return do_convert(argc, argv);
  if (!strcmp(argv[1], "uvfinfo")) return do_uvfinfo(argc, argv);
  if (!strcmp(argv[1], "tf1d")) return do_tf1d(argc, argv);
  if (!strcmp(argv[1], "tfresample")) return do_tfresample(argc, argv);
  if (!strcmp(argv[1], "rawinfo")) return do_rawinfo(argc, argv);
  if (!strcmp(argv[1], "quantize")) return do_quantize(argc, argv);
  if (!strcmp(argv[1], "writeuvf")) return do_writeuvf(argc, argv);
  if (!strcmp(argv[1], "readuvf")) return do_readuvf(argc, argv);
  if (!strcmp(argv[1], "rebrick")) return do_rebrick(argc, argv);
  if (!strcmp(argv[1], "tf2d")) return do_tf2d(argc, argv);
  if (!strcmp(argv[1], "writeuvfts")) return do_writeuvfts(argc, argv);
  if (!strcmp(argv[1], "dsdump")) return do_dsdump(argc, argv);
  return 2;
}
