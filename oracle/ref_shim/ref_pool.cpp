// ref_pool -- drives the UNMODIFIED reference GLVolumePool (Renderer/GL/GLVolumePool.cpp, compiled in
// place from /root/reference together with GLTexture3D/GLTexture/VisibilityState/Threads/BrickedDataset/
// LinearIndexDataset/Dataset) over the recording GL stand-in of gl_null.cpp, and dumps what the shader
// would see: the R32UI metadata texture (page table), the brick-pool atlas, plus the CPU-side slot table
// and the visibility counts.  tests/test_pool_ref.py replays the same scenario through the oracle
// restatement (oracle/orc_pool.cpp) and -- on the GPU box -- through libtvkcuda.so and compares bit for
// bit.  Test infrastructure only; nothing here ships.
//
//   ref_pool <scenario.txt> <result.txt>
//
// scenario (text, one directive per line):
//   vol X Y Z | brick B | overlap O | bits 8|16|32 | float 0|1 | pool PX PY PZ | max3d N | lods N
//   layout <lod> LX LY LZ                     (brick layout per LOD, N lines)
//   sizes <file>   u32[3] per brick, TOC order (LOD-major, z, y, x): voxel counts incl. ghost
//   minmax <file>  f64[4] per brick, TOC order
//   bricks <file>  optional: tightly packed voxels of every brick, TOC order
//   create                                     (constructs the pool, DM_SYNC)
//   autopool <bytes>                           (instead of create: the pool size GPUMemMan::GetVolumePool picks for that budget)
//   first | vis1d a b | vis2d a b c d | visiso v | upload n (x y z lod)*n | dump | glsl <strategy 0..3> <out file>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "StdTuvokDefines.h"
#define private public        // test shim only: MasterController::m_pSystemInfo is set by hand for the `autopool` directive
#include "Controller/MasterController.h"
#undef private
#include "Controller/Controller.h"
#include "Basics/SystemInfo.h"
#include "Renderer/GPUMemMan/GPUMemMan.h"
#include "IO/LinearIndexDataset.h"
#include "Renderer/AbstrRenderer.h"
#include "Renderer/VisibilityState.h"
#include "Renderer/GL/GLVolumePool.h"
#include "gl_null.h"

using namespace tuvok;

namespace {

struct Scenario {
  uint32_t vol[3] = {0, 0, 0}, brick = 0, overlap = 0, bits = 8, pool[3] = {0, 0, 0}, lods = 0;
  bool is_float = false;
  std::vector<UINTVECTOR3> layout;
  std::vector<uint32_t> sizes;      // 3 per brick
  std::vector<double> minmax;       // 4 per brick
  std::vector<uint8_t> bricks;      // optional voxel payload
  std::vector<uint64_t> brick_off;  // byte offset of each brick in `bricks`
  std::vector<uint64_t> lod_first;  // first TOC index of each LOD
};

std::vector<uint8_t> slurp(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) { fprintf(stderr, "ref_pool: cannot open %s\n", path.c_str()); exit(2); }
  return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

// The dataset the pool talks to: geometry + min/max + voxels come from the scenario files; every query
// the pool makes goes through the reference's own BrickedDataset / LinearIndexDataset code.
class ScenarioDataset : public LinearIndexDataset {
public:
  explicit ScenarioDataset(const Scenario& s) : S(s) {
    NBricksHint(S.sizes.size() / 3);
    for (uint32_t lod = 0; lod < S.lods; lod++) {
      const UINTVECTOR3 L = S.layout[lod];
      for (uint32_t i = 0; i < L.volume(); i++) {
        const uint64_t toc = S.lod_first[lod] + i;
        BrickMD md;
        md.center = FLOATVECTOR3(0, 0, 0);
        md.extents = FLOATVECTOR3(1, 1, 1);
        md.n_voxels = UINTVECTOR3(S.sizes[toc * 3], S.sizes[toc * 3 + 1], S.sizes[toc * 3 + 2]);
        AddBrick(BrickKey(0, lod, i), md);
      }
    }
  }
  uint64_t Toc(const BrickKey& k) const { return S.lod_first[std::get<1>(k)] + std::get<2>(k); }

  virtual UINTVECTOR3 GetBrickLayout(size_t lod, size_t) const { return S.layout[lod]; }
  virtual float MaxGradientMagnitude() const { return 1.0f; }
  virtual MinMaxBlock MaxMinForKey(const BrickKey& k) const {
    const uint64_t t = Toc(k);
    return MinMaxBlock(S.minmax[t * 4], S.minmax[t * 4 + 1], S.minmax[t * 4 + 2], S.minmax[t * 4 + 3]);
  }
  template <typename T> bool Fetch(const BrickKey& k, std::vector<T>& v) const {
    const uint64_t t = Toc(k);
    const size_t n = size_t(S.sizes[t * 3]) * S.sizes[t * 3 + 1] * S.sizes[t * 3 + 2];
    v.resize(n);
    if (sizeof(T) * 8 != S.bits) return false;
    if (S.bricks.empty()) { std::fill(v.begin(), v.end(), T(0)); return true; }
    memcpy(v.data(), &S.bricks[S.brick_off[t]], n * sizeof(T));
    return true;
  }
  virtual bool GetBrick(const BrickKey& k, std::vector<uint8_t>& v) const { return Fetch(k, v); }
  virtual bool GetBrick(const BrickKey& k, std::vector<int8_t>& v) const { return Fetch(k, v); }
  virtual bool GetBrick(const BrickKey& k, std::vector<uint16_t>& v) const { return Fetch(k, v); }
  virtual bool GetBrick(const BrickKey& k, std::vector<int16_t>& v) const { return Fetch(k, v); }
  virtual bool GetBrick(const BrickKey& k, std::vector<uint32_t>& v) const { return Fetch(k, v); }
  virtual bool GetBrick(const BrickKey& k, std::vector<int32_t>& v) const { return Fetch(k, v); }
  virtual bool GetBrick(const BrickKey& k, std::vector<float>& v) const { return Fetch(k, v); }
  virtual bool GetBrick(const BrickKey& k, std::vector<double>& v) const { return Fetch(k, v); }
  virtual unsigned GetLODLevelCount() const { return S.lods; }
  virtual UINT64VECTOR3 GetDomainSize(const size_t lod = 0, const size_t = 0) const {
    UINT64VECTOR3 d(S.vol[0], S.vol[1], S.vol[2]);
    for (size_t i = 0; i < lod; i++) d = UINT64VECTOR3((d.x + 1) / 2, (d.y + 1) / 2, (d.z + 1) / 2);
    return d;
  }
  virtual UINTVECTOR3 GetBrickOverlapSize() const { return UINTVECTOR3(S.overlap, S.overlap, S.overlap); }
  virtual UINT64VECTOR3 GetEffectiveBrickSize(const BrickKey& k) const {
    const uint64_t t = Toc(k);
    return UINT64VECTOR3(S.sizes[t * 3] - 2 * S.overlap, S.sizes[t * 3 + 1] - 2 * S.overlap,
                         S.sizes[t * 3 + 2] - 2 * S.overlap);
  }
  virtual unsigned GetBitWidth() const { return S.bits; }
  virtual uint64_t GetComponentCount() const { return 1; }
  virtual bool GetIsSigned() const { return S.is_float; }
  virtual bool GetIsFloat() const { return S.is_float; }
  virtual bool IsSameEndianness() const { return true; }
  virtual std::pair<double, double> GetRange() const { return std::make_pair(0.0, 1.0); }
  virtual UINTVECTOR3 GetMaxBrickSize() const { return UINTVECTOR3(S.brick, S.brick, S.brick); }
  virtual bool Export(uint64_t, const std::string&, bool) const { return false; }
  virtual bool ApplyFunction(uint64_t, bfqn*, void*, uint64_t) const { return false; }
  virtual Dataset* Create(const std::string&, uint64_t, bool) const { return NULL; }
  virtual const char* Name() const { return "scenario"; }

private:
  const Scenario& S;
};

// exposes the protected CPU-side state of the unmodified class
class PoolProbe : public GLVolumePool {
public:
  PoolProbe(const UINTVECTOR3& ps, LinearIndexDataset* ds)
      : GLVolumePool(ps, ds, GL_LINEAR, true, GLVolumePool::DM_SYNC) {}
  const std::vector<uint32_t>& Meta() const { return m_vBrickMetadata; }
  const std::vector<PoolSlotData>& Slots() const { return m_vPoolSlotData; }
  const std::vector<uint32_t>& LodOffsets() const { return m_vLoDOffsetTable; }
  uint32_t Total() const { return m_iTotalBrickCount; }
  unsigned MetaTex() const { return m_pPoolMetadataTexture->GetGLID(); }
  unsigned DataTex() const { return m_pPoolDataTexture->GetGLID(); }
};

uint64_t fnv1a(const uint8_t* p, size_t n) {
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ull; }
  return h;
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: ref_pool scenario.txt result.txt [pool_dump.bin]\n"); return 2; }
  std::ifstream in(argv[1]);
  FILE* out = fopen(argv[2], "w");
  if (!in || !out) { fprintf(stderr, "ref_pool: cannot open files\n"); return 2; }

  Scenario S;
  ScenarioDataset* ds = NULL;
  PoolProbe* pool = NULL;
  std::string line;
  while (std::getline(in, line)) {
    std::istringstream ls(line);
    std::string op;
    if (!(ls >> op) || op[0] == '#') continue;
    if (op == "vol") ls >> S.vol[0] >> S.vol[1] >> S.vol[2];
    else if (op == "brick") ls >> S.brick;
    else if (op == "overlap") ls >> S.overlap;
    else if (op == "bits") ls >> S.bits;
    else if (op == "float") { int f; ls >> f; S.is_float = f != 0; }
    else if (op == "pool") ls >> S.pool[0] >> S.pool[1] >> S.pool[2];
    else if (op == "max3d") { int m; ls >> m; glnull_set_max_3d(m); }
    else if (op == "lods") { ls >> S.lods; S.layout.resize(S.lods); }
    else if (op == "layout") { uint32_t l; ls >> l; ls >> S.layout[l].x >> S.layout[l].y >> S.layout[l].z; }
    else if (op == "sizes") { std::string p; ls >> p; auto b = slurp(p); S.sizes.resize(b.size() / 4); memcpy(S.sizes.data(), b.data(), b.size()); }
    else if (op == "minmax") { std::string p; ls >> p; auto b = slurp(p); S.minmax.resize(b.size() / 8); memcpy(S.minmax.data(), b.data(), b.size()); }
    else if (op == "bricks") { std::string p; ls >> p; S.bricks = slurp(p); }
    else if (op == "autopool") {
      // the pool size the reference picks itself: the UNMODIFIED GPUMemMan::GetVolumePool (GPUMemMan.cpp:766-844, compiled in
      // place, the rest of the class dropped by --gc-sections) with GetMaxUsableGPUMem = the given budget.  Neither
      // GPUMemMan nor SystemInfo is constructed (their constructors pull in the whole renderer): the member function only
      // reads m_iAllocatedGPUMemory (0) and the SystemInfo's usable-GPU-memory field, so both live in zeroed storage.
      unsigned long long budget;
      ls >> budget;
      uint64_t first = 0, off = 0;
      for (uint32_t l = 0; l < S.lods; l++) { S.lod_first.push_back(first); first += S.layout[l].volume(); }
      for (size_t b = 0; b < S.sizes.size() / 3; b++) {
        S.brick_off.push_back(off);
        off += uint64_t(S.sizes[b * 3]) * S.sizes[b * 3 + 1] * S.sizes[b * 3 + 2] * (S.bits / 8);
      }
      ds = new ScenarioDataset(S);
      static uint64_t si_store[(sizeof(SystemInfo) + 7) / 8], mm_store[(sizeof(GPUMemMan) + 7) / 8];
      SystemInfo* si = reinterpret_cast<SystemInfo*>(si_store);
      si->SetMaxUsableGPUMem(budget);
      Controller::Instance().m_pSystemInfo = si;
      GPUMemMan* mm = reinterpret_cast<GPUMemMan*>(mm_store);
      GLVolumePool* ap = mm->GetVolumePool(ds, GL_LINEAR, 0);
      if (!ap) { fprintf(out, "autopool failed\n"); }
      else {
        const UINTVECTOR3 cap = ap->GetPoolCapacity();
        const UINTVECTOR3 mb = ds->GetMaxUsedBrickSizes();
        fprintf(out, "autopool %u %u %u capacity %u %u %u\n", cap.x * mb.x, cap.y * mb.y, cap.z * mb.z, cap.x, cap.y, cap.z);
      }
      Controller::Instance().m_pSystemInfo = NULL;
    }
    else if (op == "create") {
      uint64_t first = 0, off = 0;
      for (uint32_t l = 0; l < S.lods; l++) { S.lod_first.push_back(first); first += S.layout[l].volume(); }
      for (size_t b = 0; b < S.sizes.size() / 3; b++) {
        S.brick_off.push_back(off);
        off += uint64_t(S.sizes[b * 3]) * S.sizes[b * 3 + 1] * S.sizes[b * 3 + 2] * (S.bits / 8);
      }
      ds = new ScenarioDataset(S);
      pool = new PoolProbe(UINTVECTOR3(S.pool[0], S.pool[1], S.pool[2]), ds);
      const UINTVECTOR3 cap = pool->GetPoolCapacity();
      fprintf(out, "create total %u lods %u capacity %u %u %u offsets", pool->Total(), pool->GetLoDCount(), cap.x, cap.y, cap.z);
      for (uint32_t o : pool->LodOffsets()) fprintf(out, " %u", o);
      uint32_t dim[3], bpt;
      glnull_texture(pool->MetaTex(), dim, &bpt);
      fprintf(out, " metadim %u %u %u\n", dim[0], dim[1], dim[2]);
    } else if (op == "first") {
      // GLGridLeaper::RegisterDataset / Initialize: the single lowest-resolution brick goes to the last slot
      const UINTVECTOR4 last(0, 0, 0, pool->GetLoDCount() - 1);
      pool->UploadFirstBrick(ds->IndexFrom4D(last, 0));
      fprintf(out, "first\n");
    } else if (op == "vis1d" || op == "vis2d" || op == "visiso") {
      VisibilityState vs;
      if (op == "vis1d") { double a, b; ls >> a >> b; vs.NeedsUpdate(a, b); }
      else if (op == "vis2d") { double a, b, c, d; ls >> a >> b >> c >> d; vs.NeedsUpdate(a, b, c, d); }
      else { double v; ls >> v; vs.NeedsUpdate(v); }
      const UINTVECTOR4 c = pool->RecomputeVisibility(vs, 0, true);
      fprintf(out, "counts %u %u %u %u\n", c.x, c.y, c.z, c.w);
    } else if (op == "upload") {
      uint32_t n; ls >> n;
      std::vector<UINTVECTOR4> ids(n);
      for (uint32_t i = 0; i < n; i++) ls >> ids[i].x >> ids[i].y >> ids[i].z >> ids[i].w;
      const uint32_t paged = pool->UploadBricks(ids, false);
      fprintf(out, "paged %u\n", paged);
    } else if (op == "glsl") {
      // the GLSL the pool generates for the shader side (GetBrick, ComputeLOD, TransformToPoolSpace, samplePool, ...)
      int strategy; std::string path; ls >> strategy >> path;
      const std::string g = pool->GetShaderFragment(0, 1, (GLVolumePool::MissingBrickStrategy)strategy, "");
      std::ofstream f(path); f << g; f.close();
      fprintf(out, "glsl %zu\n", g.size());
    } else if (op == "dump") {
      uint32_t dim[3], bpt;
      const uint32_t* tex = reinterpret_cast<const uint32_t*>(glnull_texture(pool->MetaTex(), dim, &bpt));
      const size_t n = size_t(dim[0]) * dim[1] * dim[2];
      // what the shader sees (texture) must equal the CPU copy; report both so the test can tell
      bool same = n == pool->Meta().size() && memcmp(tex, pool->Meta().data(), n * 4) == 0;
      fprintf(out, "meta %zu texture_equals_cpu %d", n, int(same));
      for (size_t i = 0; i < n; i++) fprintf(out, " %u", tex[i]);
      fprintf(out, "\nslots %zu", pool->Slots().size());
      for (const PoolSlotData& s : pool->Slots())
        fprintf(out, " %d %llu %u %u %u", s.m_iBrickID, (unsigned long long)s.m_iTimeOfCreation, s.PositionInPool().x,
                s.PositionInPool().y, s.PositionInPool().z);
      const uint8_t* atlas = glnull_texture(pool->DataTex(), dim, &bpt);
      const size_t nb = size_t(dim[0]) * dim[1] * dim[2] * bpt;
      fprintf(out, "\natlas %u %u %u %u %016llx\n", dim[0], dim[1], dim[2], bpt, (unsigned long long)fnv1a(atlas, nb));
      if (argc > 3) { FILE* f = fopen(argv[3], "wb"); fwrite(atlas, 1, nb, f); fclose(f); }
    } else {
      fprintf(stderr, "ref_pool: unknown directive '%s'\n", op.c_str());
      return 2;
    }
  }
  fclose(out);
  delete pool;
  delete ds;
  return 0;
}
