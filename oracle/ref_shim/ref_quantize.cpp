// ref_quantize -- runs the UNMODIFIED reference quantiser (IO/Quantize.h: Quantize<T, U>; AbstrConverter::Process8Bits)
// on a raw file, the way RAWConverter's quantize() calls it (IO/RAWConverter.cpp:205-300).  Compiled in place from
// /root/reference by oracle/Makefile; tests/test_quantize.py compares the oracle restatement with it.  Test infrastructure.
//
//   ref_quantize <in.raw> <type i8|u8|i16|u16|i32|u32|f32|f64> <n values> <out bits 8|16> <out.raw> <hist.txt>
//     hist.txt: line 1 = "changed <0|1>", then one bin count per line
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>

#include "StdTuvokDefines.h"
#include "Basics/BStream.h"
#include "Basics/EndianConvert.h"
#include "Basics/LargeRAWFile.h"
#include "IO/AbstrConverter.h"
#include "IO/UVF/Histogram1DDataBlock.h"
#include "IO/Quantize.h"

template <typename T, typename U>
static bool run(LargeRAWFile& in, uint64_t n, const std::string& out, Histogram1DDataBlock* h) {
  BStreamDescriptor bsd;
  bsd.elements = n; bsd.components = 1; bsd.width = sizeof(T);
  bsd.is_signed = ctti<T>::is_signed; bsd.fp = std::is_floating_point<T>::value;
  bsd.big_endian = EndianConvert::IsBigEndian(); bsd.timesteps = 1;
  return Quantize<T, U>(in, bsd, out, h);
}

int main(int argc, char** argv) {
  if (argc < 7) { fprintf(stderr, "usage: ref_quantize in.raw type n bits out.raw hist.txt\n"); return 2; }
  const std::string type = argv[2], out = argv[5];
  const uint64_t n = strtoull(argv[3], nullptr, 10);
  const int bits = atoi(argv[4]);
  LargeRAWFile in(argv[1]);
  in.Open(false);
  if (!in.IsOpen()) return 2;
  Histogram1DDataBlock hist;
  bool changed = false;
  if (type == "i8") changed = AbstrConverter::Process8Bits(in, out, n, true, &hist);
  else if (type == "u8") changed = AbstrConverter::Process8Bits(in, out, n, false, &hist);
#define CASE(NAME, T) \
  else if (type == NAME) changed = bits == 8 ? run<T, unsigned char>(in, n, out, &hist) : run<T, unsigned short>(in, n, out, &hist);
  CASE("i16", short) CASE("u16", unsigned short) CASE("i32", int32_t) CASE("u32", uint32_t) CASE("f32", float) CASE("f64", double)
  else return 2;
  in.Close();
  FILE* f = fopen(argv[6], "w");
  if (!f) return 2;
  fprintf(f, "changed %d\n", changed ? 1 : 0);
  const std::vector<uint64_t>& v = hist.GetHistogram();
  for (size_t i = 0; i < v.size(); i++) fprintf(f, "%llu\n", (unsigned long long)v[i]);
  fclose(f);
  return 0;
}
