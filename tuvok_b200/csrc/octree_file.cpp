// octree_file.cpp -- see octree_file.h.  Host-side IO of the streaming path (plain C++, no CUDA).
#include "octree_file.h"

#include <dlfcn.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>
#include <vector>

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
  struct stat st;
  if (fstat(fd, &st) != 0 || st.st_size < 0) { error = "cannot stat the file"; return false; }
  const uint64_t fsize = (uint64_t)st.st_size;
  if (offset > fsize) { error = "octree header offset beyond the end of the file"; return false; }
  if (component_type >= 10 || component_count == 0 || vol[0] == 0 || vol[1] == 0 || vol[2] == 0 ||
      aspect[0] * aspect[1] * aspect[2] == 0.0 || brick[0] == 0 || brick[1] == 0 || brick[2] == 0) {
    error = "zero field in the octree header";
    return false;
  }
  // a hostile header must not drive allocations or wrap the arithmetic below: every field is bounded by what a
  // 32-bit renderer geometry can address (tvk_open_octree_file refuses larger ones anyway)
  for (int i = 0; i < 3; i++)
    if (vol[i] > 0xffffffffull || brick[i] > 0xffffffffull) { error = "octree header: size field exceeds 32 bits"; return false; }
  if (component_count > 16) { error = "octree header: more than 16 components"; return false; }
  if (precomputed_normals && component_count != 4) { error = "precomputed normals need 4 components"; return false; }
  for (int i = 0; i < 3; i++)
    if (brick[i] <= 2ull * overlap) { error = "brick size does not exceed twice the overlap"; return false; }

  // LoD table: halve (ceil) until 1^3; bricks = ceil(size / (brick - 2*overlap))
  lod_size.clear(); lod_layout.clear(); lod_first.clear();
  uint64_t s[3] = {vol[0], vol[1], vol[2]}, first = 0;
  bool start = true;
  // every brick needs at least one table entry in the file (12 bytes in version 0, 40 bytes later)
  const uint64_t toc_entry_bytes = version > 0 ? 40 : 12;
  const uint64_t max_bricks = (fsize - offset) / toc_entry_bytes;
  do {
    if (lod_first.size() >= 64) { error = "octree header: more than 64 levels of detail"; return false; }
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
    // n < 2^96 cannot be formed: each factor is < 2^32 and the product is checked stepwise against the file size
    if (lod_layout[lod_layout.size() - 3] > max_bricks || lod_layout[lod_layout.size() - 3] * lod_layout[lod_layout.size() - 2] > max_bricks ||
        n > max_bricks || first + n > max_bricks) {
      error = "octree header implies more bricks than the file can hold a table of contents for";
      return false;
    }
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
      if (t.length > fsize || off > fsize) { error = "a brick lies beyond the end of the file"; return false; }
      off += t.length;
    }
  }
  if (!c.ok) { error = "short read in the octree table of contents"; return false; }
  const uint64_t room = fsize - base;      // base <= fsize was checked above
  for (const OctreeToc& t : toc)           // overflow-safe: no sum of file-supplied values is formed
    if (t.offset > room || t.length > room - t.offset) { error = "a brick lies beyond the end of the file"; return false; }
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
  if (!c.ok || c.pos > fsize || to_first > fsize - c.pos) { out->error = "corrupt global header"; ::close(fd); return false; }
  uint64_t off = c.pos + to_first;                     // GlobalHeader::GetDataPos
  uint64_t toc_seen = 0, mm_seen = 0, h1_seen = 0, h2_seen = 0;
  for (;;) {
    if (off > fsize || fsize - off < 32) { out->error = "data block list runs past the end of the file"; ::close(fd); return false; }
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
        if (!b.ok || comps == 0 || comps > 16 || b.pos > fsize || n > (fsize - b.pos) / (comps * 32)) { out->error = "corrupt MaxMin block"; ::close(fd); return false; }
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
    } else if (semantics == 5) {                       // UVFTables::BS_1D_HISTOGRAM
      if (h1_seen == timestep) {
        const uint64_t n = b.get<uint64_t>();
        if (!b.ok || b.pos > fsize || n > (fsize - b.pos) / 8) { out->error = "corrupt 1D histogram block"; ::close(fd); return false; }
        std::vector<uint64_t> bins((size_t)n);
        if (n && !pread_all(fd, bins.data(), (size_t)n * 8, b.pos)) { out->error = "short read in the 1D histogram block"; ::close(fd); return false; }
        out->hist1d_size = n;
        for (uint64_t i = 0; i < n; i++) if (bins[(size_t)i] != 0) out->hist1d_filled = i + 1;
        out->have_hist1d = true;
      }
      h1_seen++;
    } else if (semantics == 6) {                       // UVFTables::BS_2D_HISTOGRAM
      if (h2_seen == timestep) {
        out->max_grad_magnitude = b.get<float>();
        out->hist2d_size[0] = b.get<uint64_t>(); out->hist2d_size[1] = b.get<uint64_t>();
        if (!b.ok) { out->error = "corrupt 2D histogram block"; ::close(fd); return false; }
        out->have_hist2d = true;
      }
      h2_seen++;
    }
    if (to_next == 0) break;
    if (to_next > fsize - off) { out->error = "data block list runs past the end of the file"; ::close(fd); return false; }
    off += to_next;
    if (out->n_blocks > (1u << 20)) { out->error = "data block list does not end"; ::close(fd); return false; }
  }
  ::close(fd);
  out->n_toc = toc_seen;
  if (toc_seen <= timestep) { out->error = toc_seen ? "timestep out of range" : "no TOC block (legacy raster-data UVF: use the generic brick source)"; return false; }
  return true;
}

bool uvf_range(const UvfScan& sc, uint64_t lod0_bricks, double* lo, double* hi) {
  const uint64_t n = std::min<uint64_t>(lod0_bricks, sc.maxmin.size() / 4);
  if (!sc.have_maxmin || n == 0) return false;
  double a = sc.maxmin[0], b = sc.maxmin[1];
  for (uint64_t i = 1; i < n; i++) {
    a = std::min(a, sc.maxmin[(size_t)i * 4]);
    b = std::max(b, sc.maxmin[(size_t)i * 4 + 1]);
  }
  *lo = a; *hi = b;
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

// ---- LZMA (raw LZMA1 stream, no end marker, known output size) ------------------------------------------------------
// The reference compresses bricks with the LZMA SDK's LzmaEncode (IO/UVF/ExtendedOctree/LzmaCompression.cpp:105-127,
// writeEndMark = 0) and decodes them with LzmaDecode into a buffer of the brick's size (:129-149); the 5 property bytes
// are not stored with the brick but re-derived from the compression level in the octree header
// (ExtendedOctree::InitLzmaCompression, ExtendedOctree.cpp:51-55).  LzmaEncProps_Normalize leaves lc = 3, lp = 0,
// pb = 2 at every level, and the dictionary size does not matter when the whole output buffer is the dictionary, so the
// level is not needed here.  This is the LZMA decoder of the published specification, written for one-shot decoding.
namespace {
struct LzmaRange {
  const uint8_t* p; const uint8_t* end;
  uint32_t range = 0xFFFFFFFFu, code = 0;
  bool bad = false;
  uint8_t next() { if (p < end) return *p++; bad = true; return 0; }
  void init() {
    if (next() != 0) bad = true;
    for (int i = 0; i < 4; i++) code = (code << 8) | next();
  }
  void normalize() { if (range < (1u << 24)) { range <<= 8; code = (code << 8) | next(); } }
  unsigned bit(uint16_t* prob) {
    const uint32_t bound = (range >> 11) * *prob;
    unsigned b;
    if (code < bound) { *prob = (uint16_t)(*prob + (((1u << 11) - *prob) >> 5)); range = bound; b = 0; }
    else { *prob = (uint16_t)(*prob - (*prob >> 5)); code -= bound; range -= bound; b = 1; }
    normalize();
    return b;
  }
  uint32_t direct(unsigned n) {
    uint32_t r = 0;
    do {
      range >>= 1; code -= range;
      const uint32_t t = 0u - (code >> 31);
      code += range & t;
      if (code == range) bad = true;
      normalize();
      r = (r << 1) + t + 1;
    } while (--n);
    return r;
  }
  unsigned tree(uint16_t* probs, unsigned bits) {
    unsigned m = 1;
    for (unsigned i = 0; i < bits; i++) m = (m << 1) + bit(&probs[m]);
    return m - (1u << bits);
  }
  unsigned rtree(uint16_t* probs, unsigned bits) {
    unsigned m = 1, sym = 0;
    for (unsigned i = 0; i < bits; i++) { const unsigned b = bit(&probs[m]); m = (m << 1) + b; sym |= b << i; }
    return sym;
  }
};
struct LzmaLen {
  uint16_t choice, choice2, low[16][8], mid[16][8], high[256];
  void init() {
    choice = choice2 = 1024;
    for (auto& r : low) for (auto& v : r) v = 1024;
    for (auto& r : mid) for (auto& v : r) v = 1024;
    for (auto& v : high) v = 1024;
  }
  unsigned decode(LzmaRange& rc, unsigned pos_state) {
    if (rc.bit(&choice) == 0) return rc.tree(low[pos_state], 3);
    if (rc.bit(&choice2) == 0) return 8 + rc.tree(mid[pos_state], 3);
    return 16 + rc.tree(high, 8);
  }
};
}  // namespace

static bool lzma_decode(const uint8_t* src, size_t n_src, uint8_t* dst, size_t n_dst) {
  const unsigned lc = 3, lp = 0, pb = 2;
  LzmaRange rc; rc.p = src; rc.end = src + n_src;
  rc.init();
  std::vector<uint16_t> lit((size_t)0x300 << (lc + lp), 1024);
  uint16_t is_match[12 << 4], is_rep[12], is_rep_g0[12], is_rep_g1[12], is_rep_g2[12], is_rep0_long[12 << 4];
  uint16_t pos_slot[4][64], pos_dec[115], align[16];
  for (auto& v : is_match) v = 1024; for (auto& v : is_rep0_long) v = 1024;
  for (int i = 0; i < 12; i++) is_rep[i] = is_rep_g0[i] = is_rep_g1[i] = is_rep_g2[i] = 1024;
  for (auto& r : pos_slot) for (auto& v : r) v = 1024;
  for (auto& v : pos_dec) v = 1024; for (auto& v : align) v = 1024;
  LzmaLen len_dec, rep_len_dec; len_dec.init(); rep_len_dec.init();
  uint32_t rep0 = 0, rep1 = 0, rep2 = 0, rep3 = 0;
  unsigned state = 0;
  size_t pos = 0;
  while (pos < n_dst && !rc.bad) {
    const unsigned ps = (unsigned)pos & ((1u << pb) - 1);
    if (rc.bit(&is_match[(state << 4) + ps]) == 0) {               // literal
      const unsigned prev = pos ? dst[pos - 1] : 0;
      uint16_t* probs = &lit[(size_t)0x300 * ((((unsigned)pos & ((1u << lp) - 1)) << lc) + (prev >> (8 - lc)))];
      unsigned sym = 1;
      if (state >= 7) {
        unsigned mb = dst[pos - rep0 - 1];
        do {
          const unsigned mbit = (mb >> 7) & 1; mb <<= 1;
          const unsigned b = rc.bit(&probs[((1 + mbit) << 8) + sym]);
          sym = (sym << 1) | b;
          if (mbit != b) break;
        } while (sym < 0x100);
      }
      while (sym < 0x100) sym = (sym << 1) | rc.bit(&probs[sym]);
      dst[pos++] = (uint8_t)sym;
      state = state < 4 ? 0 : state < 10 ? state - 3 : state - 6;
      continue;
    }
    unsigned len;
    if (rc.bit(&is_rep[state]) != 0) {
      if (pos == 0) return false;
      if (rc.bit(&is_rep_g0[state]) == 0) {
        if (rc.bit(&is_rep0_long[(state << 4) + ps]) == 0) {         // short rep
          state = state < 7 ? 9 : 11;
          if ((size_t)rep0 + 1 > pos) return false;
          dst[pos] = dst[pos - rep0 - 1]; pos++;
          continue;
        }
      } else {
        uint32_t dist;
        if (rc.bit(&is_rep_g1[state]) == 0) dist = rep1;
        else {
          if (rc.bit(&is_rep_g2[state]) == 0) dist = rep2;
          else { dist = rep3; rep3 = rep2; }
          rep2 = rep1;
        }
        rep1 = rep0; rep0 = dist;
      }
      len = rep_len_dec.decode(rc, ps);
      state = state < 7 ? 8 : 11;
    } else {
      rep3 = rep2; rep2 = rep1; rep1 = rep0;
      len = len_dec.decode(rc, ps);
      state = state < 7 ? 7 : 10;
      const unsigned len_state = len < 3 ? len : 3;
      const unsigned slot = rc.tree(pos_slot[len_state], 6);
      if (slot < 4) rep0 = slot;
      else {
        const unsigned nb = (slot >> 1) - 1;
        uint32_t dist = (2u | (slot & 1u)) << nb;
        if (slot < 14) dist += rc.rtree(pos_dec + dist - slot, nb);
        else { dist += rc.direct(nb - 4) << 4; dist += rc.rtree(align, 4); }
        rep0 = dist;
      }
      if (rep0 == 0xFFFFFFFFu) break;                                // end marker (not written by the reference)
    }
    if ((size_t)rep0 + 1 > pos) return false;
    len += 2;
    if (len > n_dst - pos) return false;                             // a match must not run past the brick
    const uint8_t* m = dst + pos - rep0 - 1;
    for (unsigned i = 0; i < len; i++) dst[pos + i] = m[i];          // byte-wise: the match may overlap its own output
    pos += len;
  }
  return !rc.bad && pos == n_dst;
}

// ---- bzip2: the runtime library of the system (no header in the image: the one entry point is declared here) --------
// BzlibCompression.cpp:38-52 calls BZ2_bzBuffToBuffDecompress(dst, &dstLen, src, srcLen, 0, 0).
static bool bz2_decode(const uint8_t* src, size_t n_src, uint8_t* dst, size_t n_dst, std::string* err) {
  typedef int (*fn_t)(char*, unsigned int*, char*, unsigned int, int, int);
  static fn_t fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    for (const char* name : {"libbz2.so.1.0", "libbz2.so.1", "libbz2.so"}) {
      if (void* h = dlopen(name, RTLD_NOW | RTLD_LOCAL)) { fn = (fn_t)dlsym(h, "BZ2_bzBuffToBuffDecompress"); if (fn) break; }
    }
  }
  if (!fn) { if (err) *err = "bzip2-compressed bricks need the system's libbz2 runtime library, which was not found"; return false; }
  unsigned int n = (unsigned int)n_dst;
  const int rc = fn((char*)dst, &n, (char*)src, (unsigned int)n_src, 0, 0);
  if (rc != 0 || n != n_dst) { if (err) *err = "bzip2 stream is corrupt"; return false; }
  return true;
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
    case OC_LZMA:
      if (!lzma_decode(tmp.get(), (size_t)t.length, static_cast<uint8_t*>(dst), uncompressed_bytes)) {
        if (err) *err = "LZMA stream is corrupt";
        return false;
      }
      return true;
    case OC_BZLIB:
      return bz2_decode(tmp.get(), (size_t)t.length, static_cast<uint8_t*>(dst), uncompressed_bytes, err);
    default: if (err) *err = "unknown brick compression"; return false;
  }
}

}  // namespace tvk
