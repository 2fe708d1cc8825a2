// ref_dynbrick -- wraps a .uvf opened by the UNMODIFIED reference UVFDataset in the UNMODIFIED reference
// DynamicBrickingDS (IO/DynamicBrickingDS.cpp, IO/BrickCache.cpp compiled in place from /root/reference) and dumps what
// the renderers then see: LoD count, per level the domain size and brick layout, per target brick its voxel counts, the
// MM_PRECOMPUTE min / max and an FNV-1a hash of the voxels GetBrick returns.  tests/test_rebrick.py compares the
// oracle's restatement (oracle/orc.py rebrick_*) with it.  A constructor that throws (target size not a divisor of the
// source's) prints "throws <what>" -- the reference's own tuneven KATs.  Test infrastructure only.
//
// usage: ref_dynbrick <in.uvf> <out.txt> <bx> <by> <bz> [source_max_brick]
#include <array>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <stdexcept>
#include <vector>
#include "StdTuvokDefines.h"
#include "IO/DynamicBrickingDS.h"
#include "IO/uvfDataset.h"

using namespace tuvok;

static unsigned long long fnv1a(const unsigned char* p, size_t n) {
  unsigned long long h = 1469598103934665603ull;
  for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ull; }
  return h;
}

int main(int argc, char** argv) {
  if (argc < 6) { fprintf(stderr, "usage: ref_dynbrick in.uvf out.txt bx by bz [source_max_brick]\n"); return 2; }
  const unsigned max_brick = argc > 6 ? (unsigned)atoi(argv[6]) : 256;
  std::shared_ptr<UVFDataset> ds(new UVFDataset(argv[1], max_brick, false, false));
  if (ds->GetLODLevelCount() == 0) { fprintf(stderr, "open failed\n"); return 1; }
  FILE* o = fopen(argv[2], "w");
  const std::array<size_t, 3> bs = {{(size_t)atoi(argv[3]), (size_t)atoi(argv[4]), (size_t)atoi(argv[5])}};
  try {
    DynamicBrickingDS dyn(ds, bs, size_t(64) << 20, DynamicBrickingDS::MM_PRECOMPUTE);
    const unsigned lods = dyn.GetLODLevelCount();
    const UINTVECTOR3 mb = dyn.GetMaxBrickSize(), mu = dyn.GetMaxUsedBrickSizes(), ov = dyn.GetBrickOverlapSize();
    fprintf(o, "lods %u total %llu maxbrick %u %u %u maxused %u %u %u overlap %u %u %u bits %u\n", lods,
            (unsigned long long)dyn.GetTotalBrickCount(), mb.x, mb.y, mb.z, mu.x, mu.y, mu.z, ov.x, ov.y, ov.z, dyn.GetBitWidth());
    for (unsigned l = 0; l < lods; l++) {
      const UINT64VECTOR3 d = dyn.GetDomainSize(l, 0);
      const UINTVECTOR3 lay = dyn.GetBrickLayout(l, 0);
      fprintf(o, "lod %u domain %llu %llu %llu layout %u %u %u\n", l, (unsigned long long)d.x, (unsigned long long)d.y,
              (unsigned long long)d.z, lay.x, lay.y, lay.z);
      const size_t n = size_t(lay.x) * lay.y * lay.z;
      for (size_t i = 0; i < n; i++) {
        const BrickKey k(0, l, i);
        const UINTVECTOR3 vc = dyn.GetBrickVoxelCounts(k);
        const MinMaxBlock mm = dyn.MaxMinForKey(k);
        const size_t nv = size_t(vc.x) * vc.y * vc.z;
        unsigned long long h = 0;
        if (dyn.GetBitWidth() == 8) { std::vector<uint8_t> v; dyn.GetBrick(k, v); h = fnv1a((const unsigned char*)v.data(), nv); }
        else if (dyn.GetBitWidth() == 16) { std::vector<uint16_t> v; dyn.GetBrick(k, v); h = fnv1a((const unsigned char*)v.data(), nv * 2); }
        else { std::vector<float> v; dyn.GetBrick(k, v); h = fnv1a((const unsigned char*)v.data(), nv * 4); }
        fprintf(o, "brick %u %zu vox %u %u %u mm %a %a fnv %016llx\n", l, i, vc.x, vc.y, vc.z, mm.minScalar, mm.maxScalar, h);
      }
    }
  } catch (const std::runtime_error& e) {
    fprintf(o, "throws %s\n", e.what());
  }
  fclose(o);
  return 0;
}
