"""Runs the reference's OWN GLSL on the CPU (test infrastructure).

The GridLeaper fragment shader (Shaders/GLGridLeaper-blend.glsl + the Method / GradientTools / lighting / Compositing
files it links with), the GLSL that the unmodified GLVolumePool generates for the page-table walk (through
oracle/_ref/ref_pool) and the GLSL that GLHashTable generates for the miss reports (oracle/_ref/ref_hash) are read at test
time, put through a purely SYNTACTIC rewrite (storage / parameter qualifiers, array constructors, float literal suffixes,
`main`), compiled with g++ against oracle/glsl/glsl_emu.h and executed one fragment after the other.  No shader text is
copied into the repository.  What the emulation fixes where GL leaves things implementation-defined is the arithmetic
contract of DESIGN.md section 4 (see glsl_emu.h)."""
import os
import re
import struct
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHADERS = "/root/reference/Shaders"
EMU = os.path.join(ROOT, "oracle", "glsl")
REF_POOL = os.path.join(ROOT, "oracle", "_ref", "ref_pool")
REF_HASH = os.path.join(ROOT, "oracle", "_ref", "ref_hash")

TYPES = r"(?:vec[234]|ivec[234]|uvec[234]|mat4x4|mat4|float|uint|int|bool)"


def available():
    return os.path.isdir(SHADERS) and os.path.exists(REF_POOL) and os.path.exists(REF_HASH)


def _array_ctor(text):
    """`TYPE name[N] = TYPE[]( ... );`  ->  `TYPE name[N] = { ... };`"""
    out, pos = [], 0
    for m in re.finditer(r"=\s*" + TYPES + r"\[\]\(", text):
        depth, i = 1, m.end()
        while depth:
            depth += {"(": 1, ")": -1}.get(text[i], 0)
            i += 1
        out.append(text[pos:m.start()] + "= {" + text[m.end():i - 1] + "}")
        pos = i
    out.append(text[pos:])
    return "".join(out)


def rewrite(text, main_name="shader_main"):
    """GLSL 4.20 -> C++ (syntax only; every statement and expression is kept as written)."""
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)                                   # licence blocks
    text = re.sub(r"^\s*#version.*$", "", text, flags=re.M)
    text = re.sub(r"^\s*layout\s*\(pixel_center_integer\).*$", "extern vec4 gl_FragCoord;", text, flags=re.M)
    text = re.sub(r"^\s*layout\s*\([^)]*\)\s*(?:coherent\s+)?uniform\s+(\w+)\s+(\w+)\s*;", r"extern \1 \2;", text, flags=re.M)
    text = re.sub(r"^\s*layout\s*\(location\s*=\s*\d+\)\s*out\s+(\w+)\s+(\w+)\s*;", r"extern \1 \2;", text, flags=re.M)
    text = re.sub(r"^in\s+(\w+)\s+(\w+)\s*;", r"extern \1 \2;", text, flags=re.M)      # fragment shader inputs
    text = _array_ctor(text)
    text = re.sub(r"^\s*uniform\s+(\w+)\s+(\w+(?:\[\d+\])?)\s*=", r"\1 \2 =", text, flags=re.M)   # initialised uniforms
    text = re.sub(r"^\s*uniform\s+(\w+)\s+(\w+)\s*;", r"extern \1 \2;", text, flags=re.M)
    # parameter qualifiers: arrays decay to pointers (reference semantics), everything else needs a reference
    text = re.sub(r"\b(?:out|inout)\s+(" + TYPES + r")\s+(\w+)\s*(\[\d+\])", r"\1 \2\3", text)
    text = re.sub(r"\b(?:out|inout)\s+(" + TYPES + r")\s+(\w+)", r"\1& \2", text)
    text = re.sub(r"\bin\s+(" + TYPES + r")\s+", r"\1 ", text)
    text = re.sub(r"\bvoid\s+main\s*\(\s*(?:void)?\s*\)", "void %s()" % main_name, text)
    # GLSL floating literals are single precision
    text = re.sub(r"(?<![\w.])(\d+\.\d*|\.\d+)(?![\w.])", r"\1f", text)
    return text


def read_shader(name):
    with open(os.path.join(SHADERS, name)) as f:
        return f.read()


def generated_glsl(tmp, octree, vol_size, brick, overlap, dtype, pool_size, strategy, finest, hash_size, rehash):
    """The pool fragment from the reference GLVolumePool and the hash fragment from GLHashTable for this scene."""
    import pool_ref
    tmp = str(tmp)
    pool_ref.run(tmp, octree, vol_size, brick, overlap, dtype, pool_size, [("first",)], with_voxels=False)
    scen = os.path.join(tmp, "scenario.txt")
    out = os.path.join(tmp, "pool.glsl")
    with open(scen, "a") as f:
        f.write("glsl %d %s\n" % (strategy, out))
    subprocess.check_call([REF_POOL, scen, os.path.join(tmp, "result.txt")], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    pool = open(out).read()
    h = subprocess.run([REF_HASH, "glsl", str(finest[0]), str(finest[1]), str(finest[2]), str(hash_size), str(rehash), "16384"],
                       capture_output=True, text=True, check=True).stdout
    for a, b in (("hashHashValue", "HashValue"), ("hashAccessHashTable", "AccessHashTable"), ("hashSerialize", "Serialize"),
                 ("hashHash", "Hash"), ("hashhashTable", "hashTable")):
        h = h.replace(a, b)
    return pool, h


METHOD = {(0, False): "GLGridLeaper-Method-1D.glsl", (0, True): "GLGridLeaper-Method-1D-L.glsl",
          (1, False): "GLGridLeaper-Method-2D.glsl", (1, True): "GLGridLeaper-Method-2D-L.glsl"}

DRIVER = r"""
// ---- definitions of what the shader text declares `extern` ----
float sampleRateModifier; mat4x4 mEyeToModel;
float fTransScale, fGradientScale, fLoDFactor, fLevelZeroWorldSpaceError, fIsoval;
vec3 vLightAmbient, vLightDiffuse, vLightSpecular, vModelSpaceLightDir, vModelSpaceEyePos, vDomainScale, volumeAspect;
sampler2D rayStartPoint, rayStartColor; sampler1D dummy0, dummy1; usampler3D metaData; sampler3D volumePool;
@TF_TYPE@ transferFunction; uimage1D hashTable;
vec4 gl_FragCoord, accRayColor, rayResumeColor, rayResumePos; vec3 vPosInViewCoords;
unsigned long long g_samples = 0;

#include <cstdio>
#include <cstdlib>
static std::vector<char> slurp(const char* p) {
  FILE* f = fopen(p, "rb"); fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
  std::vector<char> b(n); if (fread(b.data(), 1, n, f) != (size_t)n) abort(); fclose(f); return b;
}
int main(int argc, char** argv) {
  std::vector<char> in = slurp(argv[1]);
  const char* p = in.data();
  auto u32 = [&]() { uint32_t v; memcpy(&v, p, 4); p += 4; return v; };
  auto f32 = [&]() { float v; memcpy(&v, p, 4); p += 4; return v; };
  auto v3 = [&]() { vec3 v; v.x = f32(); v.y = f32(); v.z = f32(); return v; };
  const uint32_t W = u32(), H = u32();
  for (int i = 0; i < 16; i++) mEyeToModel.a[i] = f32();
  sampleRateModifier = f32(); fTransScale = f32(); fGradientScale = f32(); fLoDFactor = f32(); fLevelZeroWorldSpaceError = f32();
  vLightAmbient = v3(); vLightDiffuse = v3(); vLightSpecular = v3(); vModelSpaceLightDir = v3(); vModelSpaceEyePos = v3();
  vDomainScale = v3();
  const uint32_t md[3] = {u32(), u32(), u32()}, ps[3] = {u32(), u32(), u32()};
  const uint32_t dtype = u32(), nearest = u32(), tfw = u32(), tfh = u32(), hash_size = u32();
  const float norm = f32();
  const size_t npx = (size_t)W * H;
  const float* entry = (const float*)p; p += npx * 16;
  const float* start = (const float*)p; p += npx * 16;
  const float* exit_eye = (const float*)p; p += npx * 12;
  const uint8_t* covered = (const uint8_t*)p; p += npx;
  metaData.d = (const uint32_t*)p; metaData.w = md[0]; metaData.h = md[1]; metaData.z = md[2]; p += (size_t)md[0] * md[1] * md[2] * 4;
  const size_t es = dtype == 0 ? 1 : dtype == 1 ? 2 : 4;
  volumePool.d = p; volumePool.w = ps[0]; volumePool.h = ps[1]; volumePool.z = ps[2]; volumePool.dtype = dtype;
  volumePool.norm = norm; volumePool.nearest = nearest != 0; p += (size_t)ps[0] * ps[1] * ps[2] * es;
  transferFunction.rgba8 = (const uint8_t*)p; transferFunction.w = tfw; set_tf_height(transferFunction, tfh); p += (size_t)tfw * tfh * 4;
  std::vector<uint32_t> hash(hash_size, 0u);
  hashTable.d = hash.data();
  rayStartPoint.f32 = entry; rayStartPoint.w = W; rayStartPoint.h = H;
  rayStartColor.f32 = start; rayStartColor.w = W; rayStartColor.h = H;
  std::vector<float> out(npx * 12, 0.0f);
  for (uint32_t y = 0; y < H; y++)
    for (uint32_t x = 0; x < W; x++) {
      const size_t i = (size_t)y * W + x;
      if (!covered[i]) continue;                       // no back face rasterised: the cleared render targets stay 0
      gl_FragCoord = vec4((float)x, (float)y, 0.0f, 1.0f);          // pixel_center_integer
      vPosInViewCoords = vec3(exit_eye[3 * i], exit_eye[3 * i + 1], exit_eye[3 * i + 2]);
      accRayColor = rayResumeColor = rayResumePos = vec4();
      shader_main();
      memcpy(&out[i * 4], &accRayColor.x, 16);
      memcpy(&out[npx * 4 + i * 4], &rayResumeColor.x, 16);
      memcpy(&out[npx * 8 + i * 4], &rayResumePos.x, 16);
    }
  FILE* f = fopen(argv[2], "wb");
  fwrite(out.data(), 4, out.size(), f);
  fwrite(hash.data(), 4, hash.size(), f);
  fwrite(&g_samples, 8, 1, f);
  fclose(f);
  return 0;
}
"""

PRELUDE = r"""
#include "glsl_emu.h"
struct uimage1D { uint32_t* d = nullptr; };
static uint imageAtomicCompSwap(uimage1D& img, int pos, uint cmp, uint val) { uint old = img.d[pos]; if (old == cmp) img.d[pos] = val; return old; }
static void set_tf_height(sampler1D&, uint32_t) {}
static void set_tf_height(sampler2D& s, uint32_t h) { s.h = (int)h; }
"""


def build(tmp, mode, lighting, pool_glsl, hash_glsl):
    """Translation unit = emulation header + rewritten reference shader text + driver; returns the executable."""
    parts = [PRELUDE, rewrite(hash_glsl), rewrite(pool_glsl)]
    names = ["Compositing.glsl", "lighting.glsl", "GLGridLeaper-GradientTools.glsl", METHOD[(mode, bool(lighting))],
             "GLGridLeaper-blend.glsl"]
    for n in names:
        parts.append("// ---- %s (read from the reference tree, syntactic rewrite only)\n" % n + rewrite(read_shader(n)))
    tf_type = "sampler2D" if mode == 1 else "sampler1D"
    src = os.path.join(str(tmp), "shader_as_cpp.cpp")
    with open(src, "w") as f:
        f.write("\n".join(parts) + DRIVER.replace("@TF_TYPE@", tf_type))
    exe = os.path.join(str(tmp), "shader_as_cpp")
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-mfma", "-ffp-contract=off", "-fno-fast-math", "-w", "-I", EMU, "-o", exe, src])
    return exe


def run(exe, tmp, params, emm, exit_eye, entry, start_color, covered, meta, meta_dim, atlas, tf, norm, light, domain_scale):
    """One GridLeaper raycast pass through the executed reference GLSL.  Returns (out0, out1, out2, hash)."""
    w, h = params.width, params.height
    buf = [struct.pack("<II", w, h), np.asarray(emm, np.float32).tobytes(),
           struct.pack("<5f", params.sample_rate_modifier, params.trans_scale, params.gradient_scale, params.lod_factor,
                       light["lzwse"])]
    for k in ("ambient", "diffuse", "specular", "light_dir_m", "eye_m"):
        buf.append(np.asarray(light[k], np.float32).tobytes())
    buf.append(np.asarray(domain_scale, np.float32).tobytes())
    buf.append(struct.pack("<3I", *meta_dim))
    buf.append(struct.pack("<3I", atlas.shape[2], atlas.shape[1], atlas.shape[0]))
    buf.append(struct.pack("<5I", params.dtype, params.nearest, params.tf_w, params.tf_h, params.hash_size))
    buf.append(struct.pack("<f", norm))
    npx = w * h
    buf += [np.ascontiguousarray(entry, np.float32).tobytes(), np.ascontiguousarray(start_color, np.float32).tobytes(),
            np.ascontiguousarray(exit_eye, np.float32).tobytes(), np.ascontiguousarray(covered, np.uint8).tobytes()]
    m = np.zeros(meta_dim[0] * meta_dim[1] * meta_dim[2], np.uint32)
    m[:len(meta)] = meta
    buf += [m.tobytes(), np.ascontiguousarray(atlas).tobytes(), np.ascontiguousarray(tf, np.uint8).tobytes()]
    fin, fout = os.path.join(str(tmp), "scene.bin"), os.path.join(str(tmp), "out.bin")
    with open(fin, "wb") as f:
        f.write(b"".join(buf))
    subprocess.check_call([exe, fin, fout])
    raw = np.fromfile(fout, np.uint8)
    img = np.frombuffer(raw[:npx * 48].tobytes(), np.float32).reshape(3, h * w, 4)
    hsh = np.frombuffer(raw[npx * 48:npx * 48 + params.hash_size * 4].tobytes(), np.uint32)
    return img[0].copy(), img[1].copy(), img[2].copy(), hsh.copy()


ISO_DRIVER = r"""
float sampleRateModifier, fLoDFactor, fLevelZeroWorldSpaceError, fIsoval; mat4x4 mEyeToModel, mModelToEye, mModelViewIT;
vec3 vDomainScale, volumeAspect;
sampler2D rayStartPoint, rayStartNormal; usampler3D metaData; sampler3D volumePool; uimage1D hashTable;
vec4 gl_FragCoord, rayHitPos, rayHitNormal, rayResumePos, rayResumeNormal; vec3 vPosInViewCoords;
#include <cstdio>
#include <cstdlib>
static std::vector<char> slurp(const char* p) {
  FILE* f = fopen(p, "rb"); fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
  std::vector<char> b(n); if (fread(b.data(), 1, n, f) != (size_t)n) abort(); fclose(f); return b;
}
int main(int argc, char** argv) {
  std::vector<char> in = slurp(argv[1]);
  const char* p = in.data();
  auto u32 = [&]() { uint32_t v; memcpy(&v, p, 4); p += 4; return v; };
  auto f32 = [&]() { float v; memcpy(&v, p, 4); p += 4; return v; };
  const uint32_t W = u32(), H = u32();
  for (int i = 0; i < 16; i++) mEyeToModel.a[i] = f32();
  for (int i = 0; i < 16; i++) mModelToEye.a[i] = f32();
  for (int i = 0; i < 16; i++) mModelViewIT.a[i] = f32();
  sampleRateModifier = f32(); fIsoval = f32(); fLoDFactor = f32(); fLevelZeroWorldSpaceError = f32();
  vDomainScale.x = f32(); vDomainScale.y = f32(); vDomainScale.z = f32();
  const uint32_t md[3] = {u32(), u32(), u32()}, ps[3] = {u32(), u32(), u32()};
  const uint32_t dtype = u32(), nearest = u32(), hash_size = u32();
  const float norm = f32();
  const size_t npx = (size_t)W * H;
  const float* entry = (const float*)p; p += npx * 16;
  const float* start = (const float*)p; p += npx * 16;
  const float* exit_eye = (const float*)p; p += npx * 12;
  const uint8_t* covered = (const uint8_t*)p; p += npx;
  metaData.d = (const uint32_t*)p; metaData.w = md[0]; metaData.h = md[1]; metaData.z = md[2]; p += (size_t)md[0] * md[1] * md[2] * 4;
  volumePool.d = p; volumePool.w = ps[0]; volumePool.h = ps[1]; volumePool.z = ps[2]; volumePool.dtype = dtype;
  volumePool.norm = norm; volumePool.nearest = nearest != 0;
  std::vector<uint32_t> hash(hash_size, 0u);
  hashTable.d = hash.data();
  rayStartPoint.f32 = entry; rayStartPoint.w = W; rayStartPoint.h = H;
  rayStartNormal.f32 = start; rayStartNormal.w = W; rayStartNormal.h = H;
  std::vector<float> out(npx * 16, 0.0f);
  for (uint32_t y = 0; y < H; y++)
    for (uint32_t x = 0; x < W; x++) {
      const size_t i = (size_t)y * W + x;
      if (!covered[i]) continue;
      gl_FragCoord = vec4((float)x, (float)y, 0.0f, 1.0f);
      vPosInViewCoords = vec3(exit_eye[3 * i], exit_eye[3 * i + 1], exit_eye[3 * i + 2]);
      rayHitPos = rayHitNormal = rayResumePos = rayResumeNormal = vec4();
      shader_main();
      memcpy(&out[i * 4], &rayHitPos.x, 16);
      memcpy(&out[npx * 4 + i * 4], &rayHitNormal.x, 16);
      memcpy(&out[npx * 8 + i * 4], &rayResumePos.x, 16);
      memcpy(&out[npx * 12 + i * 4], &rayResumeNormal.x, 16);
    }
  FILE* f = fopen(argv[2], "wb");
  fwrite(out.data(), 4, out.size(), f);
  fwrite(hash.data(), 4, hash.size(), f);
  fclose(f);
  return 0;
}
"""

COMPOSE_DRIVER = r"""
sampler2D texRayHitPos, texRayHitNormal; vec3 vLightAmbient, vLightDiffuse, vLightSpecular, vLightDir; vec2 vScreensize, vProjParam;
vec4 gl_FragCoord, gl_FragColor; float gl_FragDepth; bool g_discarded;
#include <cstdio>
#include <cstdlib>
int main(int argc, char** argv) {
  FILE* f = fopen(argv[1], "rb");
  uint32_t W, H; float u[14];
  if (fread(&W, 4, 1, f) != 1 || fread(&H, 4, 1, f) != 1 || fread(u, 4, 14, f) != 14) abort();
  const size_t npx = (size_t)W * H;
  std::vector<float> pos(npx * 4), nrm(npx * 4), out(npx * 4, 0.0f);
  if (fread(pos.data(), 4, npx * 4, f) != npx * 4 || fread(nrm.data(), 4, npx * 4, f) != npx * 4) abort();
  fclose(f);
  vLightAmbient = vec3(u[0], u[1], u[2]); vLightDiffuse = vec3(u[3], u[4], u[5]); vLightSpecular = vec3(u[6], u[7], u[8]);
  vLightDir = vec3(u[9], u[10], u[11]); vProjParam = vec2(u[12], u[13]); vScreensize = vec2((float)W, (float)H);
  texRayHitPos.f32 = pos.data(); texRayHitPos.w = W; texRayHitPos.h = H;
  texRayHitNormal.f32 = nrm.data(); texRayHitNormal.w = W; texRayHitNormal.h = H;
  for (uint32_t y = 0; y < H; y++)
    for (uint32_t x = 0; x < W; x++) {
      gl_FragCoord = vec4((float)x + 0.5f, (float)y + 0.5f, 0.0f, 1.0f);     // default pixel centres
      gl_FragColor = vec4(); g_discarded = false;
      compose_main();
      if (!g_discarded) memcpy(&out[((size_t)y * W + x) * 4], &gl_FragColor.x, 16);   // cleared target stays 0 otherwise
    }
  f = fopen(argv[2], "wb");
  fwrite(out.data(), 4, out.size(), f);
  fclose(f);
  return 0;
}
"""


def _compile(tmp, name, source):
    src = os.path.join(str(tmp), name + ".cpp")
    with open(src, "w") as f:
        f.write(source)
    exe = os.path.join(str(tmp), name)
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-mfma", "-ffp-contract=off", "-fno-fast-math", "-w", "-I", EMU, "-o", exe, src])
    return exe


def build_iso(tmp, pool_glsl, hash_glsl):
    """GLGridLeaper-iso.glsl + Method-iso + GradientTools (+ generated pool / hash fragments) as one executable."""
    parts = [PRELUDE, rewrite(hash_glsl), rewrite(pool_glsl)]
    for n in ("Compositing.glsl", "GLGridLeaper-GradientTools.glsl", "GLGridLeaper-Method-iso.glsl", "GLGridLeaper-iso.glsl"):
        parts.append("// ---- %s\n" % n + rewrite(read_shader(n)))
    return _compile(tmp, "iso_as_cpp", "\n".join(parts) + ISO_DRIVER)


def run_iso(exe, tmp, params, u, exit_eye, ray_start, start_normal, covered, meta, meta_dim, atlas):
    w, h = params.width, params.height
    mvit = np.asarray(u["mv_inv"], np.float32).reshape(4, 4).T.copy()     # uploaded array of transpose(inverse(MV))
    buf = [struct.pack("<II", w, h), np.asarray(u["emm"], np.float32).tobytes(), np.asarray(u["model_to_eye"], np.float32).tobytes(),
           mvit.tobytes(), struct.pack("<4f", params.sample_rate_modifier, params.isoval, params.lod_factor, u["lzwse"]),
           np.asarray(u["domain_scale"], np.float32).tobytes(), struct.pack("<3I", *meta_dim),
           struct.pack("<3I", atlas.shape[2], atlas.shape[1], atlas.shape[0]),
           struct.pack("<3I", params.dtype, params.nearest, params.hash_size), struct.pack("<f", u["norm"]),
           np.ascontiguousarray(ray_start, np.float32).tobytes(), np.ascontiguousarray(start_normal, np.float32).tobytes(),
           np.ascontiguousarray(exit_eye, np.float32).tobytes(), np.ascontiguousarray(covered, np.uint8).tobytes()]
    m = np.zeros(meta_dim[0] * meta_dim[1] * meta_dim[2], np.uint32)
    m[:len(meta)] = meta
    buf += [m.tobytes(), np.ascontiguousarray(atlas).tobytes()]
    fin, fout = os.path.join(str(tmp), "iso_scene.bin"), os.path.join(str(tmp), "iso_out.bin")
    with open(fin, "wb") as f:
        f.write(b"".join(buf))
    subprocess.check_call([exe, fin, fout])
    raw = np.fromfile(fout, np.uint8)
    npx = w * h
    img = np.frombuffer(raw[:npx * 64].tobytes(), np.float32).reshape(4, npx, 4)
    hsh = np.frombuffer(raw[npx * 64:npx * 64 + params.hash_size * 4].tobytes(), np.uint32)
    return [img[i].copy() for i in range(4)], hsh.copy()


def build_compose(tmp):
    """Compose-FS.glsl (deferred isosurface lighting, compatibility profile: texture2D, gl_FragColor, discard)."""
    pre = PRELUDE + "#define discard { g_discarded = true; return; }\nextern vec4 gl_FragCoord, gl_FragColor; extern float gl_FragDepth; extern bool g_discarded;\n"
    return _compile(tmp, "compose_as_cpp", pre + rewrite(read_shader("Compose-FS.glsl"), "compose_main") + COMPOSE_DRIVER)


def run_compose(exe, tmp, w, h, ambient, diffuse, specular, light_dir, hit_pos, hit_normal):
    fin, fout = os.path.join(str(tmp), "compose.bin"), os.path.join(str(tmp), "compose_out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("<II", w, h))
        f.write(np.asarray(list(ambient) + list(diffuse) + list(specular) + list(light_dir) + [0.0, 0.0], np.float32).tobytes())
        f.write(np.ascontiguousarray(hit_pos, np.float32).tobytes())
        f.write(np.ascontiguousarray(hit_normal, np.float32).tobytes())
    subprocess.check_call([exe, fin, fout])
    return np.fromfile(fout, np.float32).reshape(h * w, 4)
