// octree_file.cpp -- see octree_file.h.  Host-side IO of the streaming path (plain C++, no CUDA).
#include "octree_file.h"

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <cmath>
#include <cstring>
#include <memory>

namespace tvk {

namespace {

// sequential little-endian reader over pread (the writer stores native endianness; a big-endian host is
// not a B200 host)
struct Cursor {
  int fd;
  uint64_t pos;
  bool ok = true;
  template <typename V> V get() {
    V v{};
    if (ok && pread(fd, &v, sizeof(V), (off_t)pos) != (ssize_t)sizeof(V)) ok = false;
    pos += sizeof(V);
    return v;
  }
};

bool pread_all(int fd, void* dst, size_t n, uint64_t off) {
  uint8_t* p = static_cast<uint8_t*>(dst);
  while (n) {
    const ssize_t r = pread(fd, p, n, (off_t)off);
    if (r <= 0) return false;
    p += r; off += (uint64_t)r; n -= (size_t)r;
  }
  return true;
}

const size_t kTypeBytes[10] = {1, 2, 4, 8, 1, 2, 4, 8, 4, 8};   // ExtendedOctree::COMPONENT_TYPE order

}  // namespace

OctreeFile::~OctreeFile() { close(); }

void OctreeFile::close() {
  if (fd >= 0) ::close(fd);
  fd = -1;
}

size_t OctreeFile::element_bytes() const {
  return component_type < 10 ? kTypeBytes[component_type] * (size_t)component_count : 0;
}

// ExtendedOctree::Open (ExtendedOctree.cpp:87-165) + ComputeMetadata (:188-243)
bool OctreeFile::open(const char* path, uint64_t offset, uint64_t uvf_file_version) {
  close();
  fd = ::open(path, O_RDONLY);
  if (fd < 0) { error = std::string("cannot open ") + path; return false; }
  base = offset;
  Cursor c{fd, offset};
  component_type = c.get<uint32_t>();
  component_count = c.get<uint64_t>();
  precomputed_normals = c.get<uint8_t>() != 0;     // `bool` is one byte in the writer
  for (int i = 0; i < 3; i++) vol[i] = c.get<uint64_t>();
  for (int i = 0; i < 3; i++) aspect[i] = c.get<double>();
  for (int i = 0; i < 3; i++) brick[i] = c.get<uint64_t>();
  overlap = c.get<uint32_t>();
  version = 0;
  if (uvf_file_version > 4) {
    version = c.get<uint32_t>();
    if (version == 0) { error = "octree version 0 in a UVF >= 5 file (corrupt)"; return false; }
  }
  total_size = version > 0 ? c.get<uint64_t>() : 0;
  compression_level = version > 1 ? c.get<uint32_t>() : 0;
  if (!c.ok) { error = "short read in the octree header"; return false; }
  if (component_type >= 10 || component_count == 0 || vol[0] * vol[1] * vol[2] == 0 ||
      aspect[0] * aspect[1] * aspect[2] == 0.0 || brick[0] * brick[1] * brick[2] == 0) {
    error = "zero field in the octree header";
    return false;
  }
  if (precomputed_normals && component_count != 4) { error = "precomputed normals need 4 components"; return false; }
  for (int i = 0; i < 3; i++)
    if (brick[i] <= 2ull * overlap) { error = "brick size does not exceed twice the overlap"; return false; }

  // LoD table: halve (ceil) until 1^3; bricks = ceil(size / (brick - 2*overlap))
  lod_size.clear(); lod_layout.clear(); lod_first.clear();
  uint64_t s[3] = {vol[0], vol[1], vol[2]}, first = 0;
  bool start = true;
  do {
    if (!start)
      for (int i = 0; i < 3; i++)
        if (s[i] > 1) s[i] = (uint64_t)std::ceil(s[i] / 2.0);
    start = false;
    uint64_t n = 1;
    for (int i = 0; i < 3; i++) {
      lod_size.push_back(s[i]);
      const uint64_t l = (uint64_t)std::ceil(s[i] / double(brick[i] - 2ull * overlap));
      lod_layout.push_back(l);
      n *= l;
    }
    lod_first.push_back(first);
    first += n;
  } while (s[0] > 1 || s[1] > 1 || s[2] > 1);

  // table of contents
  toc.resize((size_t)first);
  if (version > 0) {
    for (OctreeToc& t : toc) {
      t.offset = c.get<uint64_t>();
      t.length = c.get<uint64_t>();
      t.codec = c.get<uint32_t>();
      t.valid_length = c.get<uint64_t>();
      t.atlas_w = c.get<uint32_t>();
      t.atlas_h = c.get<uint32_t>();
    }
  } else {
    // version 0 (UVF file version <= 4) stores {length, codec} only and the bricks follow back to back from
    // ExtendedOctree::ComputeHeaderSize() -- restated as the reference computes it for version 0 (:477-490:
    // no normals flag, 32-bit brick sizes, and TOCEntry::SizeInFile(0) = 28 bytes per entry)
    const uint64_t header = 4 + 8 + 3 * 8 + 3 * 8 + 3 * 4 + 4 + (uint64_t)toc.size() * (8 + 4 + 8 + 8);
    uint64_t off = header;
    for (OctreeToc& t : toc) {
      t.offset = off;
      t.length = c.get<uint64_t>();
      t.codec = c.get<uint32_t>();
      t.valid_length = t.length;
      t.atlas_w = t.atlas_h = 0;
      off += t.length;
    }
  }
  if (!c.ok) { error = "short read in the octree table of contents"; return false; }
  struct stat st;
  if (fstat(fd, &st) == 0)
    for (const OctreeToc& t : toc)
      if (base + t.offset + t.length > (uint64_t)st.st_size) { error = "a brick lies beyond the end of the file"; return false; }
  error.clear();
  return true;
}

// ExtendedOctree::BrickCoordsToIndex (ExtendedOctree.cpp:394-401)
uint64_t OctreeFile::brick_index(uint32_t x, uint32_t y, uint32_t z, uint32_t lod) const {
  const uint64_t* l = &lod_layout[3 * (size_t)lod];
  return lod_first[lod] + x + y * l[0] + z * l[0] * l[1];
}

// ExtendedOctree::ComputeBrickSize (ExtendedOctree.cpp:276-285)
void OctreeFile::brick_size(uint32_t x, uint32_t y, uint32_t z, uint32_t lod, uint32_t out[3]) const {
  const uint32_t co[3] = {x, y, z};
  for (int i = 0; i < 3; i++) {
    const uint64_t core = brick[i] - 2ull * overlap, px = lod_size[3 * (size_t)lod + i];
    const bool last = co[i] + 1 >= lod_layout[3 * (size_t)lod + i];
    out[i] = (uint32_t)((last && (px % core)) ? 2ull * overlap + px % core : brick[i]);
  }
}

bool uvf_scan(const char* path, uint64_t timestep, UvfScan* out) {
  *out = UvfScan();
  const int fd = ::open(path, O_RDONLY);
  if (fd < 0) { out->error = std::string("cannot open ") + path; return false; }
  struct stat st;
  fstat(fd, &st);
  const uint64_t fsize = (uint64_t)st.st_size;
  char magic[8] = {0};
  bool ok = pread_all(fd, magic, 8, 0) && std::memcmp(magic, "UVF-DATA", 8) == 0;
  if (!ok) { out->error = "not a UVF file (magic)"; ::close(fd); return false; }
  Cursor c{fd, 8};
  const uint8_t big_endian = c.get<uint8_t>();
  out->file_version = c.get<uint64_t>();
  c.get<uint64_t>();                                   // checksum semantics
  const uint64_t cs_len = c.get<uint64_t>();
  if (!c.ok || big_endian || cs_len > 1024) { out->error = big_endian ? "big-endian UVF files are not supported" : "corrupt global header"; ::close(fd); return false; }
  c.pos += cs_len;
  const uint64_t to_first = c.get<uint64_t>();
  uint64_t off = c.pos + to_first;                     // GlobalHeader::GetDataPos
  uint64_t toc_seen = 0, mm_seen = 0;
  for (;;) {
    if (off + 32 > fsize) { out->error = "data block list runs past the end of the file"; ::close(fd); return false; }
    Cursor b{fd, off};
    const uint64_t id_len = b.get<uint64_t>();
    if (id_len > 65536) { out->error = "corrupt data block header"; ::close(fd); return false; }
    b.pos += id_len;
    const uint64_t semantics = b.get<uint64_t>();
    b.get<uint64_t>();                                 // compression scheme of the block (none is ever written)
    const uint64_t to_next = b.get<uint64_t>();
    if (!b.ok) { out->error = "short read in a data block header"; ::close(fd); return false; }
    out->n_blocks++;
    if (semantics == 9) {                              // UVFTables::BS_TOC_BLOCK
      if (toc_seen == timestep) out->toc_payload_offset = b.pos;
      toc_seen++;
    } else if (semantics == 7) {                       // UVFTables::BS_MAXMIN_VALUES (MaxMinDataBlock.cpp:67-95)
      if (mm_seen == timestep) {
        const uint64_t n = b.get<uint64_t>(), comps = b.get<uint64_t>();
        if (!b.ok || comps == 0 || comps > 16 || b.pos + n * comps * 32 > fsize) { out->error = "corrupt MaxMin block"; ::close(fd); return false; }
        out->maxmin_components = comps;
        std::vector<double> all((size_t)(n * comps * 4));
        if (!pread_all(fd, all.data(), all.size() * 8, b.pos)) { out->error = "short read in the MaxMin block"; ::close(fd); return false; }
        const uint64_t pick = comps == 4 ? 3 : 0;      // UVFDataset::MaxMinForKey, IO/uvfDataset.cpp:1183-1188
        out->maxmin.resize((size_t)n * 4);
        for (uint64_t i = 0; i < n; i++)
          std::memcpy(&out->maxmin[(size_t)i * 4], &all[(size_t)((i * comps + pick) * 4)], 32);
        out->have_maxmin = true;
      }
      mm_seen++;
    }
    if (to_next == 0) break;
    off += to_next;
  }
  ::close(fd);
  out->n_toc = toc_seen;
  if (toc_seen <= timestep) { out->error = toc_seen ? "timestep out of range" : "no TOC block (legacy raster-data UVF: use the generic brick source)"; return false; }
  return true;
}

// LZ4 block format: token (literal length | match length), literals, 16-bit offset, 255-continued lengths.
// The block ends with literals only.
bool lz4_block_decode(const uint8_t* src, size_t src_len, uint8_t* dst, size_t dst_len) {
  const uint8_t* ip = src;
  const uint8_t* const iend = src + src_len;
  uint8_t* op = dst;
  uint8_t* const oend = dst + dst_len;
  while (ip < iend) {
    const unsigned token = *ip++;
    size_t lit = token >> 4;
    if (lit == 15) {
      unsigned b;
      do { if (ip >= iend) return false; b = *ip++; lit += b; } while (b == 255);
    }
    if (lit > (size_t)(iend - ip) || lit > (size_t)(oend - op)) return false;
    std::memcpy(op, ip, lit);
    ip += lit; op += lit;
    if (op == oend) return true;             // last sequence: literals only
    if (iend - ip < 2) return false;
    const size_t off = (size_t)ip[0] | ((size_t)ip[1] << 8);
    ip += 2;
    if (off == 0 || off > (size_t)(op - dst)) return false;
    size_t ml = token & 15u;
    if (ml == 15) {
      unsigned b;
      do { if (ip >= iend) return false; b = *ip++; ml += b; } while (b == 255);
    }
    ml += 4;
    if (ml > (size_t)(oend - op)) return false;
    const uint8_t* m = op - off;
    for (size_t i = 0; i < ml; i++) op[i] = m[i];    // byte-wise: the match may overlap its own output
    op += ml;
  }
  return op == oend;
}

// ExtendedOctree::GetBrickData (ExtendedOctree.cpp:313-360)
bool OctreeFile::read_brick(uint64_t index, size_t uncompressed_bytes, void* dst, size_t cap, std::string* err) const {
  if (index >= toc.size()) { if (err) *err = "brick index out of range"; return false; }
  const OctreeToc& t = toc[(size_t)index];
  if (uncompressed_bytes > cap) { if (err) *err = "destination too small"; return false; }
  if (t.atlas_w != 0 && t.atlas_h != 0) { if (err) *err = "2D-atlas packed bricks are not supported"; return false; }
  if (t.codec == OC_NONE) {
    if (t.length != uncompressed_bytes) { if (err) *err = "stored brick length does not match its geometry"; return false; }
    if (!pread_all(fd, dst, (size_t)t.length, base + t.offset)) { if (err) *err = "short read"; return false; }
    return true;
  }
  std::unique_ptr<uint8_t[]> tmp(new uint8_t[(size_t)t.length]);
  if (!pread_all(fd, tmp.get(), (size_t)t.length, base + t.offset)) { if (err) *err = "short read"; return false; }
  switch (t.codec) {
    case OC_ZLIB: {
      uLongf n = (uLongf)uncompressed_bytes;
      const int rc = uncompress(static_cast<Bytef*>(dst), &n, tmp.get(), (uLong)t.length);
      if (rc != Z_OK || n != uncompressed_bytes) { if (err) *err = "zlib stream is corrupt"; return false; }
      return true;
    }
    case OC_LZ4:
      if (!lz4_block_decode(tmp.get(), (size_t)t.length, static_cast<uint8_t*>(dst), uncompressed_bytes)) {
        if (err) *err = "lz4 block is corrupt";
        return false;
      }
      return true;
    case OC_LZMA: if (err) *err = "LZMA-compressed bricks are not supported by this reader"; return false;
    case OC_BZLIB: if (err) *err = "bzip2-compressed bricks are not supported by this reader"; return false;
    default: if (err) *err = "unknown brick compression"; return false;
  }
}

}  // namespace tvk
