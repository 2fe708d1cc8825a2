// ref_octree -- drives the UNMODIFIED reference ExtendedOctreeConverter
// (compiled in place from /root/reference, see oracle/Makefile) so that the
// oracle's bricking / LOD pyramid / min-max restatement can be checked
// bit-exactly against the reference itself.  Test infrastructure only; built
// into oracle/_ref/ (git-ignored).  No reference source is copied here.
//
// usage: ref_octree <in.raw> <out.bin> <dtype u8|u16|f32|rgba8> X Y Z brick overlap clamp median
//        (rgba8: four interleaved 8-bit components per voxel; min / max are those of component 3, what a renderer
//         sees for colour data, uvfDataset.cpp:1188)
//                   [octree_out [compression 0 none|1 zlib|3 lz4 [layout 0 scanline|1 morton|2 hilbert|3 random]]]
// octree_out: the ExtendedOctree file the converter wrote (= payload of a UVF TOC block) is kept there, so the
// product's file reader (tvk_open_octree_file) can be tested on reference-written data.
// out.bin: u64 lodCount, u64 brickCount, then per brick (TOC order):
//          u64 sx,sy,sz, f64 min,max, raw voxels
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "IO/UVF/ExtendedOctree/ExtendedOctreeConverter.h"
#include "IO/UVF/ExtendedOctree/ExtendedOctree.h"
#include "DebugOut/AbstrDebugOut.h"

class NullOut : public AbstrDebugOut {
public:
  virtual void printf(enum DebugChannel, const char*, const char*) {}
  virtual void printf(const char*) const {}
};

int main(int argc, char** argv) {
  if (argc < 11 || argc > 14) { fprintf(stderr, "bad args\n"); return 2; }
  std::string in = argv[1], out = argv[2], dt = argv[3];
  UINT64VECTOR3 vol(strtoull(argv[4], 0, 10), strtoull(argv[5], 0, 10), strtoull(argv[6], 0, 10));
  uint64_t brick = strtoull(argv[7], 0, 10);
  uint32_t overlap = (uint32_t)strtoul(argv[8], 0, 10);
  bool clamp = atoi(argv[9]) != 0, median = atoi(argv[10]) != 0;
  ExtendedOctree::COMPONENT_TYPE ct =
      (dt == "u8" || dt == "rgba8") ? ExtendedOctree::CT_UINT8 : dt == "u16" ? ExtendedOctree::CT_UINT16 : ExtendedOctree::CT_FLOAT32;
  const uint64_t comps = dt == "rgba8" ? 4 : 1;

  NullOut dbg;
  const bool keep = argc > 11;
  std::string tmp = keep ? std::string(argv[11]) : out + ".octree";
  const COMPRESSION_TYPE comp = argc > 12 ? (COMPRESSION_TYPE)atoi(argv[12]) : CT_NONE;
  const LAYOUT_TYPE layout = argc > 13 ? (LAYOUT_TYPE)atoi(argv[13]) : LT_SCANLINE;
  BrickStatVec stats;
  {
    ExtendedOctreeConverter c(UINT64VECTOR3(brick, brick, brick), overlap, 1ull << 30, dbg);
    if (!c.Convert(in, 0, ct, comps, vol, DOUBLEVECTOR3(1, 1, 1), tmp, 0, &stats, comp, comp == CT_LZ4 ? 1 : 6,
                   median, clamp, layout)) {
      fprintf(stderr, "convert failed\n");
      return 1;
    }
  }
  ExtendedOctree e;
  if (!e.Open(tmp, 0, 5)) { fprintf(stderr, "open failed\n"); return 1; }
  FILE* f = fopen(out.c_str(), "wb");
  uint64_t lods = e.GetLODCount(), total = 0;
  for (uint64_t l = 0; l < lods; l++) total += e.GetBrickCount(l).volume();
  fwrite(&lods, 8, 1, f);
  fwrite(&total, 8, 1, f);
  std::vector<uint8_t> buf;
  uint64_t idx = 0;
  for (uint64_t l = 0; l < lods; l++) {
    UINT64VECTOR3 bc = e.GetBrickCount(l);
    for (uint64_t z = 0; z < bc.z; z++)
      for (uint64_t y = 0; y < bc.y; y++)
        for (uint64_t x = 0; x < bc.x; x++, idx++) {
          UINT64VECTOR4 co(x, y, z, l);
          UINT64VECTOR3 bs = e.ComputeBrickSize(co);
          size_t bytes = size_t(bs.volume() * e.GetComponentTypeSize() * comps);
          buf.resize(bytes);
          e.GetBrickData(&buf[0], co);
          uint64_t s[3] = {bs.x, bs.y, bs.z};
          const size_t si = size_t(idx * comps + (comps - 1));
          double mm[2] = {stats[si].minScalar, stats[si].maxScalar};
          fwrite(s, 8, 3, f);
          fwrite(mm, 8, 2, f);
          fwrite(&buf[0], 1, bytes, f);
        }
  }
  fclose(f);
  e.Close();
  if (!keep) remove(tmp.c_str());
  return 0;
}
