// ref_hash -- the UNMODIFIED reference GLHashTable (Renderer/GL/GLHashTable.cpp, compiled in place over the recording
// null-GL): prints the GLSL it generates for the shader side of the miss-report table (Serialize / HashValue /
// AccessHashTable / Hash -- tests/test_hash_ref.py compiles that text as C++ and EXECUTES it) and decodes a table
// with its own GetData() / Int2Vector.  Test infrastructure only.
//
//   ref_hash glsl   LX LY LZ table_size rehash max_tex            -> GLSL fragment on stdout (+ "//texsize W H")
//   ref_hash decode LX LY LZ table_size rehash max_tex table.bin  -> "n" then n lines "x y z lod"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>
#include "StdTuvokDefines.h"
#include <GL/glew.h>
#include "Renderer/GL/GLHashTable.h"
#include "Renderer/GL/GLTexture.h"
#include "gl_null.h"

using namespace tuvok;

// GLHashTable keeps its texture private; the recording GL hands out ids 1, 2, ... in creation order
int main(int argc, char** argv) {
  if (argc < 8) { fprintf(stderr, "bad args\n"); return 2; }
  const std::string mode = argv[1];
  const UINTVECTOR3 lay(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]));
  const uint32_t size = (uint32_t)atoi(argv[5]), rehash = (uint32_t)atoi(argv[6]);
  glnull_set_max_2d(atoi(argv[7]));
  GLHashTable ht(lay, size, rehash, true, "hash");
  ht.InitGL();
  uint32_t dim[3], bpt;
  glnull_texture(1, dim, &bpt);
  if (mode == "glsl") {
    const std::string s = ht.GetShaderFragment(5);
    fwrite(s.data(), 1, s.size(), stdout);
    printf("//texsize %u %u\n", dim[0], dim[1]);
    return 0;
  }
  if (argc < 9) return 2;
  std::ifstream f(argv[8], std::ios::binary);
  std::vector<char> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  std::vector<uint32_t> table(size_t(dim[0]) * dim[1], 0u);
  memcpy(table.data(), buf.data(), std::min(buf.size(), table.size() * 4));
  glnull_write_texture(1, table.data(), table.size() * 4);     // what the shader's imageAtomicCompSwap calls left behind
  const std::vector<UINTVECTOR4> req = ht.GetData();
  printf("%zu\n", req.size());
  for (const UINTVECTOR4& r : req) printf("%u %u %u %u\n", r.x, r.y, r.z, r.w);
  return 0;
}
