// octree_file.h -- reader of the ExtendedOctree on-disk layout (the payload of a UVF TOC block) for the
// streaming path: header + table of contents are parsed once, bricks are pread() straight into the
// library's pinned staging memory (no std::vector hop of Dataset::GetBrick) and decoded there when the
// TOC marks them compressed.  Replaces (reference file:line):
//   ExtendedOctree::Open            IO/UVF/ExtendedOctree/ExtendedOctree.cpp:87-165
//   ExtendedOctree::ComputeMetadata IO/UVF/ExtendedOctree/ExtendedOctree.cpp:188-243
//   ExtendedOctree::GetBrickData    IO/UVF/ExtendedOctree/ExtendedOctree.cpp:313-360
//   UVFDataset::GetBrick (TOC path) IO/uvfDataset.cpp:1690-1712
//   zDecompress / lz4Decompress     IO/UVF/ExtendedOctree/ZlibCompression.cpp, Lz4Compression.cpp
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace tvk {

// ExtendedOctree.h:65-73
enum OctreeCodec : uint32_t { OC_NONE = 0, OC_ZLIB = 1, OC_LZMA = 2, OC_LZ4 = 3, OC_BZLIB = 4, OC_LZHAM = 5 };

struct OctreeToc {
  uint64_t offset, length, valid_length;
  uint32_t codec, atlas_w, atlas_h;
};

struct OctreeFile {
  int fd = -1;
  uint64_t base = 0;                // offset of the octree header inside the file (the TOC block payload)
  uint32_t component_type = 0;      // ExtendedOctree::COMPONENT_TYPE
  uint64_t component_count = 0;
  bool precomputed_normals = false;
  uint64_t vol[3] = {0, 0, 0};
  double aspect[3] = {1, 1, 1};
  uint64_t brick[3] = {0, 0, 0};    // max brick size incl. ghost
  uint32_t overlap = 0, version = 0, compression_level = 0;
  uint64_t total_size = 0;
  std::vector<uint64_t> lod_size;   // 3 per LoD
  std::vector<uint64_t> lod_layout; // 3 per LoD
  std::vector<uint64_t> lod_first;  // first TOC index of each LoD
  std::vector<OctreeToc> toc;
  std::string error;

  ~OctreeFile();
  bool open(const char* path, uint64_t offset, uint64_t uvf_file_version);
  void close();
  uint32_t lod_count() const { return (uint32_t)lod_first.size(); }
  uint64_t lod0_brick_count() const { return lod_layout.size() >= 3 ? lod_layout[0] * lod_layout[1] * lod_layout[2] : 0; }
  size_t element_bytes() const;
  uint64_t brick_index(uint32_t x, uint32_t y, uint32_t z, uint32_t lod) const;
  void brick_size(uint32_t x, uint32_t y, uint32_t z, uint32_t lod, uint32_t out[3]) const;
  // voxels of one brick (x fastest, own size incl. ghost) into dst; thread-safe (pread + private scratch).
  // Returns false and sets `err` on IO / codec errors.
  bool read_brick(uint64_t index, size_t uncompressed_bytes, void* dst, size_t cap, std::string* err) const;
};

// UVF container walk (IO/UVF/UVF.cpp:140-290, GlobalHeader.cpp:39-47, DataBlock.cpp:60-72): magic "UVF-DATA", global
// header, linked list of data blocks.  Finds the `timestep`-th TOC block (UVFTables::BS_TOC_BLOCK) and the
// `timestep`-th MaxMin block (BS_MAXMIN_VALUES; UVFDataset pairs them by order, IO/uvfDataset.cpp:640-700).
struct UvfScan {
  uint64_t file_version = 0;
  uint64_t toc_payload_offset = 0;   // byte offset of the ExtendedOctree header
  uint64_t n_blocks = 0, n_toc = 0;
  bool have_maxmin = false;
  uint64_t maxmin_components = 0;
  std::vector<double> maxmin;        // 4 doubles per brick (component 0; component 3 of RGBA data), TOC order
  // the `timestep`-th 1D / 2D histogram blocks (Histogram1DDataBlock.cpp:48-58, Histogram2DDataBlock.cpp:51-66)
  bool have_hist1d = false, have_hist2d = false;
  uint64_t hist1d_size = 0, hist1d_filled = 0;   // bins; index of the last non-zero bin + 1 (Grid1D::GetFilledSize)
  float max_grad_magnitude = 0.0f;               // Histogram2DDataBlock::GetMaxGradMagnitude
  uint64_t hist2d_size[2] = {0, 0};
  std::string error;
};
bool uvf_scan(const char* path, uint64_t timestep, UvfScan* out);
// UVFDataset::ComputeRange for a TOC-based file (IO/uvfDataset.cpp:1140-1155): min / max scalar over the LoD-0 bricks
bool uvf_range(const UvfScan& sc, uint64_t lod0_bricks, double* lo, double* hi);

// LZ4 block format (the reference calls LZ4_decompress_fast: the decoder knows only the output size)
bool lz4_block_decode(const uint8_t* src, size_t src_len, uint8_t* dst, size_t dst_len);

}  // namespace tvk
