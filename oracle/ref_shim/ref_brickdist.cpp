// ref_brickdist -- the reference's OWN brick_distance (a file-static function of Renderer/AbstrRenderer.cpp:808-841: distance
// from the eye to the closest of a brick's eight corners, each pulled towards the centre by 0.4999) and Brick ordering
// (operator<, AbstrRenderer.h:104-106).  The only way to call a file-static function without copying it is to compile its
// translation unit: AbstrRenderer.cpp is #included here (in place, from /root/reference), every function in its own section,
// and the linker drops everything that main() does not reach (--gc-sections).  tests/test_host_ref.py compares the oracle's
// classic-path brick distances (orc_classic.cpp) with it.  Test infrastructure only.
//
// usage: ref_brickdist <in.txt> <out.txt>     in: 16 floats model view, then per line: cx cy cz ex ey ez
#include <cstdio>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>
#include "Renderer/AbstrRenderer.cpp"

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  std::ifstream in(argv[1]);
  FILE* out = fopen(argv[2], "w");
  FLOATMATRIX4 mv;
  for (int i = 0; i < 16; i++) in >> mv.array[i];
  std::vector<tuvok::Brick> bricks;
  float c[6];
  while (in >> c[0] >> c[1] >> c[2] >> c[3] >> c[4] >> c[5]) {
    tuvok::Brick b;
    b.vCenter = FLOATVECTOR3(c[0], c[1], c[2]);
    b.vExtension = FLOATVECTOR3(c[3], c[4], c[5]);
    b.vCoords = UINTVECTOR3(uint32_t(bricks.size()), 0, 0);     // remembers the input position
    b.fDistance = brick_distance(b, mv);
    bricks.push_back(b);
    fprintf(out, "dist %a\n", (double)b.fDistance);
  }
  fclose(out);
  return 0;
}
