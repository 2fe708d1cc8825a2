"""The miss-report hash table against the UNMODIFIED reference GLHashTable (oracle/_ref/ref_hash, compiled
in place over the recording null-GL):
  * shader side: the GLSL that GLHashTable::GetShaderFragment GENERATES (Serialize / HashValue / AccessHashTable / Hash)
    is compiled here as C++ -- uvec4 / ivec2 / imageAtomicCompSwap supplied by a 20-line prelude, the generated text
    itself untouched -- and EXECUTED on a sequence of brick requests; the resulting table and the per-request rehash
    counts must equal the oracle's (orc_hash_insert = the arithmetic of report_missing in orc_render.c, which the CUDA
    kernel's miss lists are compared with on the GPU);
  * CPU side: the reference's GetData() / Int2Vector decode of that table == orc_hash_decode, in table order;
  * 1D and 2D (folded, small GL_MAX_TEXTURE_SIZE) table textures, full tables (probe chain exhausted)."""
import os
import subprocess

import numpy as np
import pytest

from oracle import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "ref_hash")


def _have():
    if not os.path.exists(BIN) and os.path.isdir("/root/reference/IO"):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", "ref"], stdout=subprocess.DEVNULL)
    return os.path.exists(BIN)


pytestmark = pytest.mark.skipif(not _have(), reason="oracle/_ref/ref_hash not built (reference tree absent)")

PRELUDE = r"""
#include <cstdio>
#include <vector>
typedef unsigned int uint;
struct uvec4 { uint x, y, z, w; };
struct ivec2 { int x, y; ivec2(int a, int b) : x(a), y(b) {} };
struct Image { std::vector<uint> t; int w; };
static Image hashhashTable;
static uint imageAtomicCompSwap(Image& img, int pos, uint cmp, uint val) {
  uint old = img.t[pos]; if (old == cmp) img.t[pos] = val; return old;
}
static uint imageAtomicCompSwap(Image& img, ivec2 pos, uint cmp, uint val) {
  return imageAtomicCompSwap(img, pos.y * img.w + pos.x, cmp, val);
}
"""

MAIN = r"""
int main(int argc, char** argv) {
  int w = 0, h = 0, n = 0;
  if (scanf("%d %d %d", &w, &h, &n) != 3) return 2;
  hashhashTable.t.assign((size_t)w * h, 0u); hashhashTable.w = w;
  for (int i = 0; i < n; i++) {
    uvec4 b;
    if (scanf("%u %u %u %u", &b.x, &b.y, &b.z, &b.w) != 4) return 2;
    printf("%u\n", hashHash(b));
  }
  FILE* f = fopen(argv[1], "wb");
  fwrite(hashhashTable.t.data(), 4, hashhashTable.t.size(), f);
  fclose(f);
  return 0;
}
"""


def run_reference_glsl(tmp_path, layout, size, rehash, max_tex, requests):
    """Compile the generated GLSL as C++ and run it.  Returns (table u32[w*h], rehash counts, decoded requests)."""
    args = [str(v) for v in layout] + [str(size), str(rehash), str(max_tex)]
    glsl = subprocess.run([BIN, "glsl"] + args, capture_output=True, text=True, check=True).stdout
    lines = glsl.splitlines()
    w, h = (int(v) for v in [l for l in lines if l.startswith("//texsize")][0].split()[1:3])
    body = "\n".join(l for l in lines if not l.startswith("#version") and not l.startswith("layout("))
    assert "imageAtomicCompSwap" in body and "hashSerialize" in body
    src = tmp_path / "glsl_as_cpp.cpp"
    src.write_text(PRELUDE + body + MAIN)
    exe = tmp_path / "glsl_as_cpp"
    subprocess.check_call(["g++", "-O1", "-o", str(exe), str(src)])
    table_bin = tmp_path / "table.bin"
    inp = "%d %d %d\n" % (w, h, len(requests)) + "\n".join("%d %d %d %d" % tuple(r) for r in requests) + "\n"
    out = subprocess.run([str(exe), str(table_bin)], input=inp, capture_output=True, text=True, check=True).stdout
    counts = [int(v) for v in out.split()]
    table = np.fromfile(table_bin, np.uint32)
    dec = subprocess.run([BIN, "decode"] + args + [str(table_bin)], capture_output=True, text=True, check=True).stdout.split()
    n = int(dec[0])
    decoded = np.array(dec[1:1 + 4 * n], np.uint32).reshape(n, 4)
    return table, counts, decoded, (w, h)


@pytest.mark.parametrize("layout,lods,size,rehash,max_tex,n_req", [
    ((4, 3, 3), 3, 31, 10, 16384, 20),        # small table: collisions and rehashing
    ((16, 16, 16), 5, 509, 10, 16384, 300),   # the reference defaults (RState.HashTableSize 509, RehashCount 10)
    ((8, 8, 8), 4, 13, 10, 16384, 40),        # more requests than entries: probe chains run out
    ((16, 16, 16), 5, 509, 10, 64, 300),      # GL_MAX_TEXTURE_SIZE 64: the table folds into a 2D texture
    ((64, 64, 64), 7, 4099, 3, 16384, 2000),
])
def test_generated_glsl_and_getdata_match_oracle(tmp_path, layout, lods, size, rehash, max_tex, n_req):
    rng = np.random.default_rng(size + n_req)
    req = np.stack([rng.integers(0, max(1, layout[0] >> 0), n_req), rng.integers(0, layout[1], n_req),
                    rng.integers(0, layout[2], n_req), rng.integers(0, lods, n_req)], axis=1).astype(np.uint32)
    req[n_req // 2:n_req // 2 + 5] = req[:5]          # repeated requests hit their own entry
    table, counts, decoded, (w, h) = run_reference_glsl(tmp_path, layout, size, rehash, max_tex, req)
    if max_tex < size:
        assert h > 1                                    # 2D-folded texture; entries stay linear (y * w + x)
    mine = np.zeros(size, np.uint32)
    my_counts = [orc.hash_insert(mine, rehash, layout, *[int(v) for v in r]) for r in req]
    assert my_counts == counts
    assert np.array_equal(table[:size], mine)
    assert not table[size:].any()
    assert np.array_equal(orc.hash_decode(mine, layout), decoded)
    if size == 13:
        assert max(counts) == rehash                    # some reports were dropped: the chain ran out
