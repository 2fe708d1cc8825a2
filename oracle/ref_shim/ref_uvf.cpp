// ref_uvf -- writes a complete UVF container with the UNMODIFIED reference code (UVF, GlobalHeader, TOCBlock::
// FlatDataToBrickedLOD = ExtendedOctreeConverter, Histogram1DDataBlock, MaxMinDataBlock, KeyValuePairDataBlock;
// compiled in place from /root/reference by oracle/Makefile) the way RAWConverter::ConvertRAWDataset assembles one
// (IO/RAWConverter.cpp:553-690): per timestep a TOC block, a 1D histogram block and a MaxMin block, then a
// key/value block.  A 1D and a 2D histogram block accompany every TOC block: UVFDataset only accepts files whose block
// counts match (IO/uvfDataset.cpp:484-497).  The 2D histogram is computed with 16 value bins to keep the file small.  The product's container walk (tvk_open_uvf) is tested on these files.  Test infrastructure only.
//
// usage: ref_uvf <in.raw> <out.uvf> <dtype u8|u16|f32|rgba8> X Y Z brick overlap compression(0|1|3) layout(0..3) [timesteps]
//        (rgba8: four interleaved 8-bit components; TOC + MaxMin block with four components, no histogram blocks -- the
//         histogram classes take scalar data only)
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include "StdTuvokDefines.h"
#include "Basics/LargeRAWFile.h"
#include "IO/UVF/UVF.h"
#include "IO/UVF/GlobalHeader.h"
#include "IO/UVF/TOCBlock.h"
#include "IO/UVF/MaxMinDataBlock.h"
#include "IO/UVF/Histogram1DDataBlock.h"
#include "IO/UVF/Histogram2DDataBlock.h"
#include "IO/UVF/KeyValuePairDataBlock.h"
#include "DebugOut/AbstrDebugOut.h"

class NullOut : public AbstrDebugOut {
public:
  virtual void printf(enum DebugChannel, const char*, const char*) {}
  virtual void printf(const char*) const {}
};

int main(int argc, char** argv) {
  if (argc < 11) { fprintf(stderr, "bad args\n"); return 2; }
  const std::string in = argv[1], out = argv[2], dt = argv[3];
  const UINT64VECTOR3 vol(strtoull(argv[4], 0, 10), strtoull(argv[5], 0, 10), strtoull(argv[6], 0, 10));
  const uint64_t brick = strtoull(argv[7], 0, 10);
  const uint32_t overlap = (uint32_t)strtoul(argv[8], 0, 10);
  const COMPRESSION_TYPE comp = (COMPRESSION_TYPE)atoi(argv[9]);
  const LAYOUT_TYPE layout = (LAYOUT_TYPE)atoi(argv[10]);
  const int timesteps = argc > 11 ? atoi(argv[11]) : 1;
  const ExtendedOctree::COMPONENT_TYPE ct =
      (dt == "u8" || dt == "rgba8") ? ExtendedOctree::CT_UINT8 : dt == "u16" ? ExtendedOctree::CT_UINT16 : ExtendedOctree::CT_FLOAT32;
  const uint64_t comps = dt == "rgba8" ? 4 : 1;
  NullOut dbg;
  remove(out.c_str());
  std::wstring wout(out.begin(), out.end());
  UVF uvf(wout);
  GlobalHeader gh;
  gh.bIsBigEndian = EndianConvert::IsBigEndian();
  gh.ulChecksumSemanticsEntry = UVFTables::CS_MD5;
  uvf.SetGlobalHeader(gh);
  std::vector<std::shared_ptr<TOCBlock>> tocs;
  std::vector<std::shared_ptr<MaxMinDataBlock>> mms;
  std::vector<std::shared_ptr<Histogram1DDataBlock>> hists;
  std::vector<std::shared_ptr<Histogram2DDataBlock>> hists2;
  for (int ts = 0; ts < timesteps; ts++) {
    std::shared_ptr<MaxMinDataBlock> mm(new MaxMinDataBlock(size_t(comps)));
    std::shared_ptr<TOCBlock> toc(new TOCBlock(UVF::ms_ulReaderVersion));
    toc->strBlockID = "Volume converted by ref_uvf";
    const std::string tmp = out + "." + std::to_string(ts) + ".tmp";    // (no fixed-size buffer: test paths are long)
    if (!toc->FlatDataToBrickedLOD(in, tmp, ct, comps, vol, DOUBLEVECTOR3(1, 1, 1), UINT64VECTOR3(brick, brick, brick), overlap,
                                   false, false, size_t(1) << 30, mm, &dbg, comp, comp == CT_LZ4 ? 1 : 6, layout)) {
      fprintf(stderr, "brick generation failed\n");
      return 1;
    }
    uvf.AddDataBlock(toc);
    if (dt != "f32" && comps == 1) {
      std::shared_ptr<Histogram1DDataBlock> h(new Histogram1DDataBlock());
      if (h->Compute(toc.get(), 0)) {
        uvf.AddDataBlock(h);
        hists.push_back(h);
        std::shared_ptr<Histogram2DDataBlock> h2(new Histogram2DDataBlock());
        if (h2->Compute(toc.get(), 0, 16, mm->GetGlobalValue().maxScalar)) { uvf.AddDataBlock(h2); hists2.push_back(h2); }
      }
    }
    uvf.AddDataBlock(mm);
    tocs.push_back(toc); mms.push_back(mm);
  }
  std::shared_ptr<KeyValuePairDataBlock> kv(new KeyValuePairDataBlock());
  kv->AddPair("Data Source", "ref_uvf");
  uvf.AddDataBlock(kv);
  if (!uvf.Create()) { fprintf(stderr, "UVF::Create failed\n"); return 1; }
  uvf.Close();
  for (int ts = 0; ts < timesteps; ts++) {
    const std::string tmp = out + "." + std::to_string(ts) + ".tmp";
    remove(tmp.c_str());
  }
  return 0;
}
