// ref_dataset -- opens a .uvf with the UNMODIFIED reference UVFDataset (IO/uvfDataset.cpp and what it needs, compiled
// in place from /root/reference) and dumps the data interface the renderers consume: LoD count, domain
// sizes, brick layouts, overlap, bit width, scale, range, and per brick its centre / extents / voxel counts / texture
// coordinates / min-max / voxels.  tests/test_dataset_ref.py compares the oracle's classic-path brick metadata and the
// product's file source with it.  Test infrastructure only.
//
// usage: ref_dataset <in.uvf> <out.txt> [max_brick_size]
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "StdTuvokDefines.h"
#include "IO/uvfDataset.h"

using namespace tuvok;

static unsigned long long fnv1a(const unsigned char* p, size_t n) {
  unsigned long long h = 1469598103934665603ull;
  for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ull; }
  return h;
}

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: ref_dataset in.uvf out.txt [max_brick_size]\n"); return 2; }
  const unsigned max_brick = argc > 3 ? (unsigned)atoi(argv[3]) : 256;
  UVFDataset ds(argv[1], max_brick, false);
  if (ds.GetLODLevelCount() == 0) { fprintf(stderr, "open failed\n"); return 1; }
  FILE* o = fopen(argv[2], "w");
  const unsigned lods = ds.GetLODLevelCount();
  const DOUBLEVECTOR3 sc = ds.GetScale();
  const UINTVECTOR3 ov = ds.GetBrickOverlapSize();
  const UINTVECTOR3 mu = ds.GetMaxUsedBrickSizes();
  const std::pair<double, double> rg = ds.GetRange();
  fprintf(o, "lods %u largest_single %zu bits %u comps %llu signed %d float %d overlap %u %u %u maxused %u %u %u\n", lods,
          ds.GetLargestSingleBrickLOD(0), ds.GetBitWidth(), (unsigned long long)ds.GetComponentCount(), int(ds.GetIsSigned()),
          int(ds.GetIsFloat()), ov.x, ov.y, ov.z, mu.x, mu.y, mu.z);
  fprintf(o, "scale %a %a %a range %a %a maxgrad %a total %llu\n", sc.x, sc.y, sc.z, rg.first, rg.second,
          (double)ds.MaxGradientMagnitude(), (unsigned long long)ds.GetTotalBrickCount());
  for (unsigned l = 0; l < lods; l++) {
    const UINT64VECTOR3 d = ds.GetDomainSize(l, 0);
    const UINTVECTOR3 lay = ds.GetBrickLayout(l, 0);
    fprintf(o, "lod %u domain %llu %llu %llu layout %u %u %u\n", l, (unsigned long long)d.x, (unsigned long long)d.y,
            (unsigned long long)d.z, lay.x, lay.y, lay.z);
    const size_t n = size_t(lay.x) * lay.y * lay.z;
    for (size_t i = 0; i < n; i++) {
      const BrickKey k(0, l, i);
      const BrickMD& md = ds.GetBrickMetadata(k);
      BrickTable::const_iterator it = ds.BricksBegin();
      for (; it != ds.BricksEnd(); ++it) if (it->first == k) break;
      const std::pair<FLOATVECTOR3, FLOATVECTOR3> t = ds.GetTextCoords(it, false);
      const MinMaxBlock mm = ds.MaxMinForKey(k);
      const UINTVECTOR3 vc = ds.GetBrickVoxelCounts(k);
      unsigned long long h = 0;
      const size_t nv = size_t(vc.x) * vc.y * vc.z;
      if (ds.GetBitWidth() == 8) { std::vector<uint8_t> v; ds.GetBrick(k, v); h = fnv1a((const unsigned char*)v.data(), nv); }
      else if (ds.GetBitWidth() == 16) { std::vector<uint16_t> v; ds.GetBrick(k, v); h = fnv1a((const unsigned char*)v.data(), nv * 2); }
      else { std::vector<float> v; ds.GetBrick(k, v); h = fnv1a((const unsigned char*)v.data(), nv * 4); }
      fprintf(o, "brick %u %zu center %a %a %a ext %a %a %a vox %u %u %u tmin %a %a %a tmax %a %a %a mm %a %a first %d %d %d last %d %d %d fnv %016llx\n",
              l, i, (double)md.center.x, (double)md.center.y, (double)md.center.z, (double)md.extents.x, (double)md.extents.y,
              (double)md.extents.z, vc.x, vc.y, vc.z, (double)t.first.x, (double)t.first.y, (double)t.first.z,
              (double)t.second.x, (double)t.second.y, (double)t.second.z, mm.minScalar, mm.maxScalar,
              int(ds.BrickIsFirstInDimension(0, k)), int(ds.BrickIsFirstInDimension(1, k)), int(ds.BrickIsFirstInDimension(2, k)),
              int(ds.BrickIsLastInDimension(0, k)), int(ds.BrickIsLastInDimension(1, k)), int(ds.BrickIsLastInDimension(2, k)), h);
    }
  }
  fclose(o);
  return 0;
}
