/*
 * tvk_uvf.h -- C ABI for the UVF container at the two ends of the hot path: writing the K2 builder's
 * output as a UVF v5 file that stock ImageVis3D opens, and reading a stock UVF's bricks into (pinned)
 * host memory for the pager.  Host code only; no CUDA call is made here.
 *
 * Replaces, under /root/reference/Tuvok/IO/:
 *   writer  RAWConverter::ConvertRAWDataset's file assembly (RAWConverter.cpp:527-681):
 *           UVF::SetGlobalHeader / AddDataBlock / Create / Close (UVF/UVF.cpp:306-395,103-136),
 *           GlobalHeader::CopyHeaderToFile (UVF/GlobalHeader.cpp:88-95), DataBlock::CopyHeaderToFile
 *           (UVF/DataBlock.cpp:89-103), ExtendedOctree::WriteHeader (ExtendedOctree.cpp:499-549),
 *           ExtendedOctreeConverter::ComputeStatsAndCompressAll's zlib branch (.cpp:466-564),
 *           Histogram1D/2DDataBlock::CopyToFile, MaxMinDataBlock::CopyToFile (MaxMinDataBlock.cpp:97-120),
 *           KeyValuePairDataBlock::CopyToFile, UVF::ComputeChecksum (MD5, UVF.cpp:169-236)
 *   reader  UVF::Open / ParseGlobalHeader / ParseDataBlocks / VerifyChecksum (UVF.cpp:82-101,138-167,238-304),
 *           ExtendedOctree::Open / GetBrickData (ExtendedOctree.cpp:87-165,313-360) with
 *           zDecompress / lz4Decompress, the *DataBlock::GetHeaderFromFile readers, and what
 *           UVFDataset::Open pulls out of them (uvfDataset.cpp:83-216)
 * Little-endian files only (the reference writes the byte order of the machine it runs on).
 * Supported brick compression: none, zlib, lz4, lzma, bzip2 (reading; the last two through the system's liblzma /
 * libbz2, opened on first use); none, zlib (writing).  Blocks of other kinds
 * (legacy RasterDataBlock, geometry) are listed by tvk_uvf_info but not decoded.
 * All functions return 0 on success; on failure `err` (if given) receives a message.
 */
#ifndef TVK_UVF_H
#define TVK_UVF_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ExtendedOctree::COMPONENT_TYPE (ExtendedOctree.h) */
#define TVK_UVF_CT_UINT8  0
#define TVK_UVF_CT_UINT16 1
/* COMPRESSION_TYPE (ExtendedOctree.h) */
#define TVK_UVF_COMP_NONE 0
#define TVK_UVF_COMP_ZLIB 1
#define TVK_UVF_COMP_LZMA 2
#define TVK_UVF_COMP_LZ4  3
#define TVK_UVF_COMP_BZIP 4

typedef struct tvk_uvf_volume {     /* the global header of the TOC block's ExtendedOctree */
  uint32_t component_type;          /* TVK_UVF_CT_*                                        */
  uint64_t component_count;         /* 1                                                   */
  uint64_t volume_size[3];
  double   aspect[3];               /* vVolumeAspect / scale                               */
  uint64_t brick_size[3];           /* max brick size incl. overlap (cubic)                */
  uint32_t overlap;
  uint32_t compression;             /* writer: requested brick compression                 */
  uint32_t compression_level;       /* zlib 0..9 (reference default 1)                     */
} tvk_uvf_volume;

typedef struct tvk_uvf_write_desc {
  tvk_uvf_volume vol;
  const char* toc_block_id;         /* NULL -> "Volume converted by ImageVis3D"            */
  uint64_t total_bricks;            /* all LODs down to 1x1x1, ToC order                   */
  const uint64_t* brick_offsets;    /* byte offset of each brick inside `store`            */
  const uint32_t* brick_sizes;      /* 3 per brick, voxels                                 */
  const void* store;                /* host memory (e.g. the builder's store copied back)  */
  const double* minmax;             /* 2 per brick {min, max}                              */
  const uint64_t* hist1d; uint64_t hist1d_size;          /* may be NULL: block omitted     */
  const uint64_t* hist2d; uint64_t hist2d_x, hist2d_y;   /* [x][y]; may be NULL            */
  float max_grad_magnitude;
  const char* const* kv_keys; const char* const* kv_values; uint32_t kv_count;
} tvk_uvf_write_desc;

int tvk_uvf_write(const char* path, const tvk_uvf_write_desc* d, char* err, size_t errlen);

typedef struct tvk_uvf tvk_uvf;

typedef struct tvk_uvf_info {
  uint64_t uvf_version;
  uint32_t block_count;
  uint32_t block_semantics[16];     /* UVFTables::BlockSemanticTable of the first 16 blocks */
  int32_t  toc_block;               /* index of the first TOC block, -1 if none             */
  tvk_uvf_volume vol;               /* valid if toc_block >= 0 (compression: of brick 0)    */
  uint32_t octree_version;
  uint32_t lod_count;
  uint64_t total_bricks;
  uint64_t store_bytes;             /* sum of the uncompressed brick sizes                  */
  uint64_t hist1d_size, hist2d_x, hist2d_y;
  float    max_grad_magnitude;
  uint64_t maxmin_bricks;
  uint32_t kv_count;
  uint32_t checksum_ok;             /* 1 verified, 0 not verified (verify_checksum = 0)     */
  /* legacy RasterDataBlock (UVF versions < 5, e.g. DemoData/c60.uvf): finest level only    */
  int32_t  rdb_block;               /* index of the first 3-D scalar RasterDataBlock, -1 if none */
  uint32_t rdb_bits, rdb_signed;    /* element width (8 / 16) and signedness                */
  uint64_t rdb_domain[3], rdb_brick[3], rdb_overlap[3], rdb_lod_levels;
  uint32_t timesteps;               /* TOC blocks in the file = time steps (UVFDataset::GetNumberOfTimesteps, uvfDataset.cpp:621-623) */
} tvk_uvf_info;

int  tvk_uvf_open(const char* path, int verify_checksum, tvk_uvf** out, char* err, size_t errlen);
/* the same for time step `timestep` of a time-dependent file: every getter below then answers for the k-th volume /
 * histogram / max-min block of the file (UVFDataset::FindSuitableDataBlocks, uvfDataset.cpp:635-745); tvk_uvf_open = step 0 */
int  tvk_uvf_open_timestep(const char* path, int verify_checksum, uint32_t timestep, tvk_uvf** out, char* err, size_t errlen);
void tvk_uvf_close(tvk_uvf* u);
int  tvk_uvf_get_info(const tvk_uvf* u, tvk_uvf_info* info);
/* per brick: position / stored length in the file, compression, size in voxels */
int  tvk_uvf_get_toc(const tvk_uvf* u, uint64_t* file_offsets, uint64_t* lengths, uint32_t* compression,
                     uint32_t* sizes3, uint64_t n);
/* ExtendedOctree::GetBrickData for bricks [first, first+count): decompressed, each written at
 * dst + dst_offsets[i] (i relative to `first`); dst_offsets == NULL packs them tightly. */
int  tvk_uvf_read_bricks(tvk_uvf* u, uint64_t first, uint64_t count, void* dst, const uint64_t* dst_offsets);
/* RasterDataBlock::GetData over all bricks of the finest level, stitched back into the flat x-fastest volume
 * (rdb_domain voxels of rdb_bits/8 bytes): what DynamicBrickingDS (IO/DynamicBrickingDS.cpp) starts from when it
 * re-bricks a legacy / large-brick file for the ray-guided renderer. */
/* the same bricks AS STORED (still compressed): lengths from tvk_uvf_get_toc; dst_offsets NULL = packed */
int  tvk_uvf_read_bricks_raw(tvk_uvf* u, uint64_t first, uint64_t count, void* dst, const uint64_t* dst_offsets);
int  tvk_uvf_read_rdb_volume(tvk_uvf* u, void* dst, uint64_t dst_bytes);
int  tvk_uvf_get_minmax(const tvk_uvf* u, double* minmax4, uint64_t n);   /* min, max, minGrad, maxGrad per brick */
int  tvk_uvf_get_hist1d(const tvk_uvf* u, uint64_t* bins, uint64_t n);
int  tvk_uvf_get_hist2d(const tvk_uvf* u, uint64_t* bins, uint64_t n);
/* key/value i: returned pointers stay valid until tvk_uvf_close */
int  tvk_uvf_get_kv(const tvk_uvf* u, uint32_t i, const char** key, const char** value);
const char* tvk_uvf_last_error(const tvk_uvf* u);

/* MD5 of a memory range (the checksum UVF uses); digest = 16 bytes */
void tvk_md5(const void* data, uint64_t n, uint8_t digest[16]);

#ifdef __cplusplus
}
#endif
#endif /* TVK_UVF_H */
