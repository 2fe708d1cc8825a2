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


def rewrite(text, main_name="shader_main", tls=False):
    """GLSL 4.20 -> C++ (syntax only; every statement and expression is kept as written)."""
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)                                   # licence blocks
    text = re.sub(r"^\s*#version.*$", "", text, flags=re.M)
    frag = "extern thread_local" if tls else "extern"        # per-fragment state (inputs / outputs of one invocation)
    text = re.sub(r"^\s*layout\s*\(pixel_center_integer\).*$", frag + " vec4 gl_FragCoord;", text, flags=re.M)
    text = re.sub(r"^\s*layout\s*\([^)]*\)\s*(?:coherent\s+)?uniform\s+(\w+)\s+(\w+)\s*;", r"extern \1 \2;", text, flags=re.M)
    text = re.sub(r"^\s*layout\s*\(location\s*=\s*\d+\)\s*out\s+(\w+)\s+(\w+)\s*;", frag + r" \1 \2;", text, flags=re.M)
    text = re.sub(r"^in\s+(\w+)\s+(\w+)\s*;", frag + r" \1 \2;", text, flags=re.M)      # fragment shader inputs
    text = re.sub(r"^\s*varying\s+(\w+)\s+(\w+)\s*;", r"extern \1 \2;", text, flags=re.M)
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
    with open(os.path.join(SHADERS, name), encoding="latin-1") as f:     # author names in the comment headers
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


def build(tmp, mode, lighting, pool_glsl, hash_glsl, color=False):
    """Translation unit = emulation header + rewritten reference shader text + driver; returns the executable.
    color: the GLGridLeaper-Method-*-color.glsl variants GLGridLeaper picks for 4-component data (GLGridLeaper.cpp:766-795)."""
    parts = [PRELUDE, rewrite(hash_glsl), rewrite(pool_glsl)]
    method = METHOD[(mode, bool(lighting))]
    if color:
        method = method.replace(".glsl", "-color.glsl")
    names = ["Compositing.glsl", "lighting.glsl", "GLGridLeaper-GradientTools.glsl", method, "GLGridLeaper-blend.glsl"]
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


def build_iso(tmp, pool_glsl, hash_glsl, color=False):
    """GLGridLeaper-iso.glsl + Method-iso(-color) + GradientTools (+ generated pool / hash fragments) as one executable."""
    parts = [PRELUDE, rewrite(hash_glsl), rewrite(pool_glsl)]
    for n in ("Compositing.glsl", "GLGridLeaper-GradientTools.glsl",
              "GLGridLeaper-Method-iso-color.glsl" if color else "GLGridLeaper-Method-iso.glsl", "GLGridLeaper-iso.glsl"):
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


def build_compose(tmp, color=False):
    """Compose-FS.glsl / Compose-Color-FS.glsl (deferred isosurface lighting, compatibility profile: texture2D, gl_FragColor,
    discard)."""
    pre = PRELUDE + "#define discard { g_discarded = true; return; }\nextern vec4 gl_FragCoord, gl_FragColor; extern float gl_FragDepth; extern bool g_discarded;\n"
    name = "Compose-Color-FS.glsl" if color else "Compose-FS.glsl"
    return _compile(tmp, "compose_as_cpp", pre + rewrite(read_shader(name), "compose_main") + COMPOSE_DRIVER)


def run_compose(exe, tmp, w, h, ambient, diffuse, specular, light_dir, hit_pos, hit_normal):
    fin, fout = os.path.join(str(tmp), "compose.bin"), os.path.join(str(tmp), "compose_out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("<II", w, h))
        f.write(np.asarray(list(ambient) + list(diffuse) + list(specular) + list(light_dir) + [0.0, 0.0], np.float32).tobytes())
        f.write(np.ascontiguousarray(hit_pos, np.float32).tobytes())
        f.write(np.ascontiguousarray(hit_normal, np.float32).tobytes())
    subprocess.check_call([exe, fin, fout])
    return np.fromfile(fout, np.float32).reshape(h * w, 4)


# ---------------------------------------------------------------------------------------------------------------------
# classic per-brick raycaster (GLRaycaster): the fragment shaders are the reference's; the per-brick pass setup around
# them (near-plane / front-face entry FBO in RGBA16F, back-face fragments, GL under-blending) is restated in the driver
# ---------------------------------------------------------------------------------------------------------------------
CLASSIC_FILES = {
    (0, False): ["Compositing.glsl", "Volume3D.glsl", "VRender1D.glsl", "VRender1D-BScale.glsl", "VRender1DProxy.glsl",
                 "GLRaycaster-1D-FS.glsl"],
    (0, True): ["Compositing.glsl", "Volume3D.glsl", "lighting.glsl", "VRender1DLit.glsl", "GLRaycaster-1D-light-FS.glsl"],
    (1, False): ["Compositing.glsl", "Volume3D.glsl", "GLRaycaster-2D-FS.glsl"],
    (1, True): ["Compositing.glsl", "Volume3D.glsl", "lighting.glsl", "GLRaycaster-2D-light-FS.glsl"],
}

CLASSIC_DRIVER = r"""
sampler3D texVolume; @TF_TYPE@ texTrans; sampler2D texRayExitPos, texRayExit;
float fTransScale, fGradientScale, fStepScale, fRayStepsize, TFuncBias; int ScaleMethod;
vec2 vScreensize; vec3 vVoxelStepsize, vDomainScale, vLightAmbient, vLightDiffuse, vLightSpecular, vLightDir, vEyePos;
vec4 gl_FragCoord, gl_FragColor; mat4x4 gl_TextureMatrix[1]; mat3 gl_NormalMatrix;
#include <cstdio>
#include <cstdlib>
static std::vector<char> slurp(const char* p) {
  FILE* f = fopen(p, "rb"); fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
  std::vector<char> b(n); if (fread(b.data(), 1, n, f) != (size_t)n) abort(); fclose(f); return b;
}
static void mul4(const float* a, const float* b, float* o) {     // row-vector convention: o = a * b
  float t[16];
  for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) { float s = 0.0f; for (int k = 0; k < 4; k++) s = s + a[r * 4 + k] * b[k * 4 + c]; t[r * 4 + c] = s; }
  memcpy(o, t, 64);
}
static float half_round(float f) {                                // what a GL_RGBA16F render target stores
  uint32_t x; memcpy(&x, &f, 4);
  const uint32_t sign = x & 0x80000000u; uint32_t a = x & 0x7fffffffu;
  if (a >= 0x7f800000u) return f;
  if (a >= 0x477ff000u) { uint32_t r = sign | 0x7f800000u; float o; memcpy(&o, &r, 4); return o; }
  if (a < 0x33000001u) { uint32_t r = sign; float o; memcpy(&o, &r, 4); return o; }
  float af; memcpy(&af, &a, 4);
  int e; frexpf(af, &e);
  const int ulp = (e - 1 < -14 ? -14 : e - 1) - 10;
  float q = ldexpf(nearbyintf(ldexpf(af, -ulp)), ulp);
  uint32_t r; memcpy(&r, &q, 4); r |= sign; float o; memcpy(&o, &r, 4); return o;
}
int main(int argc, char** argv) {
  std::vector<char> in = slurp(argv[1]);
  const char* p = in.data();
  auto u32 = [&]() { uint32_t v; memcpy(&v, p, 4); p += 4; return v; };
  auto f32 = [&]() { float v; memcpy(&v, p, 4); p += 4; return v; };
  auto v3 = [&]() { vec3 v; v.x = f32(); v.y = f32(); v.z = f32(); return v; };
  const uint32_t W = u32(), H = u32();
  float inv_proj[16], imv[16];
  for (int i = 0; i < 16; i++) inv_proj[i] = f32();
  for (int i = 0; i < 16; i++) imv[i] = f32();
  fTransScale = f32(); fGradientScale = f32(); fStepScale = f32();
  const float sample_rate = f32(), norm = f32();
  vDomainScale = v3(); vLightAmbient = v3(); vLightDiffuse = v3(); vLightSpecular = v3(); vLightDir = v3();
  const uint32_t dtype = u32(), nearest = u32(), tfw = u32(), tfh = u32(), n_bricks = u32();
  ScaleMethod = 0; TFuncBias = 0.0f;
  vScreensize = vec2((float)W, (float)H);
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) gl_NormalMatrix.a[r * 3 + c] = imv[r * 4 + c];   // transpose(inverse(MV)) rows
  texTrans.rgba8 = (const uint8_t*)p; texTrans.w = tfw; set_tf_height(texTrans, tfh); p += (size_t)tfw * tfh * 4;
  const size_t npx = (size_t)W * H;
  std::vector<float> out(npx * 4, 0.0f), fbo(npx * 4, 0.0f), near_pt(npx * 3), org_pt(npx * 3, 0.0f), dir_pt(npx * 3);
  const bool ortho = inv_proj[11] == 0.0f;           // parallel projection (m_bOrthoView): w' does not depend on z
  for (uint32_t y = 0; y < H; y++)
    for (uint32_t x = 0; x < W; x++) {               // Render3DPreLoop: the near plane fills the entry FBO first
      const float nx = ((float)x + 0.5f) / (float)W * 2.0f - 1.0f, ny = ((float)y + 0.5f) / (float)H * 2.0f - 1.0f;
      const float* m = inv_proj;
      const float rx = nx * m[0] + ny * m[4] + -1.0f * m[8] + 1.0f * m[12], ry = nx * m[1] + ny * m[5] + -1.0f * m[9] + 1.0f * m[13];
      const float rz = nx * m[2] + ny * m[6] + -1.0f * m[10] + 1.0f * m[14], rw = nx * m[3] + ny * m[7] + -1.0f * m[11] + 1.0f * m[15];
      const size_t i = (size_t)y * W + x;
      near_pt[3 * i] = rx / rw; near_pt[3 * i + 1] = ry / rw; near_pt[3 * i + 2] = rz / rw;
      for (int k = 0; k < 3; k++) fbo[4 * i + k] = half_round(near_pt[3 * i + k]);
      for (int k = 0; k < 3; k++) dir_pt[3 * i + k] = near_pt[3 * i + k];
      if (ortho) {                                    // eye-space ray a + s * b through the near (s = 1) and far plane points
        const float fx = nx * m[0] + ny * m[4] + 1.0f * m[8] + 1.0f * m[12], fy = nx * m[1] + ny * m[5] + 1.0f * m[9] + 1.0f * m[13];
        const float fz = nx * m[2] + ny * m[6] + 1.0f * m[10] + 1.0f * m[14], fw = nx * m[3] + ny * m[7] + 1.0f * m[11] + 1.0f * m[15];
        const float far_pt[3] = {fx / fw, fy / fw, fz / fw};
        for (int k = 0; k < 3; k++) { dir_pt[3 * i + k] = far_pt[k] - near_pt[3 * i + k]; org_pt[3 * i + k] = near_pt[3 * i + k] - dir_pt[3 * i + k]; }
      }
    }
  texRayExitPos.f32 = fbo.data(); texRayExitPos.w = W; texRayExitPos.h = H;
  auto xf = [&](const float* m, float x, float y, float z) { return vec3(x * m[0] + y * m[4] + z * m[8] + 1.0f * m[12], x * m[1] + y * m[5] + z * m[9] + 1.0f * m[13], x * m[2] + y * m[6] + z * m[10] + 1.0f * m[14]); };
  const vec3 o4 = xf(imv, 0.0f, 0.0f, 0.0f);
  unsigned long long frags = 0;
  for (uint32_t bi = 0; bi < n_bricks; bi++) {
    const vec3 c = v3(), e = v3(), tmin = v3(), tmax = v3();
    const uint32_t nv[3] = {u32(), u32(), u32()}, empty = u32();
    const size_t es = dtype == 0 ? 1 : dtype == 1 ? 2 : 4;
    const char* data = p;
    if (!empty) p += (size_t)nv[0] * nv[1] * nv[2] * es;
    if (empty) continue;
    texVolume.d = data; texVolume.w = nv[0]; texVolume.h = nv[1]; texVolume.z = nv[2]; texVolume.dtype = dtype;
    texVolume.norm = norm; texVolume.nearest = nearest != 0;
    const vec3 pmin = c - vec3(e.x / 2.0f, e.y / 2.0f, e.z / 2.0f), pmax = c + vec3(e.x / 2.0f, e.y / 2.0f, e.z / 2.0f);
    const vec3 tsc = (tmin - tmax) / (pmin - pmax);
    vVoxelStepsize = vec3(1.0f / (float)nv[0], 1.0f / (float)nv[1], 1.0f / (float)nv[2]);
    const vec3 rs = (e * vVoxelStepsize) * (0.5f * 1.0f / sample_rate);
    fRayStepsize = fminf(rs.x, fminf(rs.y, rs.z));
    // GLRaycaster::ComputeEyeToTextureMatrix: eye -> world -> texture, as ONE matrix in gl_TextureMatrix[0]
    float t1[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, -pmax.x, -pmax.y, -pmax.z, 1};
    float sc[16] = {tsc.x, 0, 0, 0, 0, tsc.y, 0, 0, 0, 0, tsc.z, 0, 0, 0, 0, 1};
    float t2[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, tmax.x, tmax.y, tmax.z, 1};
    float m[16]; mul4(imv, t1, m); mul4(m, sc, m); mul4(m, t2, m);
    memcpy(gl_TextureMatrix[0].a, m, 64);
    const float lo[3] = {pmin.x, pmin.y, pmin.z}, hi[3] = {pmax.x, pmax.y, pmax.z};
    for (uint32_t y = 0; y < H; y++)
      for (uint32_t x = 0; x < W; x++) {
        const size_t i = (size_t)y * W + x;
        const vec3 pn(near_pt[3 * i], near_pt[3 * i + 1], near_pt[3 * i + 2]);
        const vec3 pa(org_pt[3 * i], org_pt[3 * i + 1], org_pt[3 * i + 2]), pb(dir_pt[3 * i], dir_pt[3 * i + 1], dir_pt[3 * i + 2]);
        const vec3 n4 = xf(imv, pn.x, pn.y, pn.z);
        const vec3 o4 = xf(imv, pa.x, pa.y, pa.z);
        const float o[3] = {o4.x, o4.y, o4.z}, d[3] = {n4.x - o4.x, n4.y - o4.y, n4.z - o4.z};
        float s_in = -INFINITY, s_out = INFINITY; bool miss = false;
        for (int k = 0; k < 3; k++) {
          if (d[k] == 0.0f) { if (o[k] < lo[k] || o[k] > hi[k]) miss = true; continue; }
          const float t0 = (lo[k] - o[k]) / d[k], t1_ = (hi[k] - o[k]) / d[k];
          s_in = fmaxf(s_in, fminf(t0, t1_)); s_out = fminf(s_out, fmaxf(t0, t1_));
        }
        if (miss || !(s_out > fmaxf(s_in, 1.0f))) continue;
        if (s_in > 1.0f) { const vec3 fe = ortho ? pa + pb * s_in : pn * s_in; fbo[4 * i] = half_round(fe.x); fbo[4 * i + 1] = half_round(fe.y); fbo[4 * i + 2] = half_round(fe.z); }
        vEyePos = ortho ? pa + pb * s_out : pn * s_out;          // interpolated back-face position (eye space)
        gl_FragCoord = vec4((float)x + 0.5f, (float)y + 0.5f, 0.0f, 1.0f);
        gl_FragColor = vec4();
        classic_main();
        frags++;
        float* dst = &out[4 * i];                               // GL blending ONE_MINUS_DST_ALPHA, ONE
        const float k = 1.0f - dst[3];
        dst[0] = fmaf(k, gl_FragColor.x, dst[0]); dst[1] = fmaf(k, gl_FragColor.y, dst[1]);
        dst[2] = fmaf(k, gl_FragColor.z, dst[2]); dst[3] = fmaf(k, gl_FragColor.w, dst[3]);
      }
  }
  FILE* f = fopen(argv[2], "wb");
  fwrite(out.data(), 4, out.size(), f);
  fclose(f);
  return 0;
}
"""


def build_classic(tmp, mode, lighting):
    pre = PRELUDE + "extern vec4 gl_FragCoord, gl_FragColor; extern mat4x4 gl_TextureMatrix[1]; extern mat3 gl_NormalMatrix;\n"
    parts = [pre]
    for n in CLASSIC_FILES[(mode, bool(lighting))]:
        parts.append("// ---- %s\n" % n + rewrite(read_shader(n), "classic_main"))
    tf_type = "sampler2D" if mode == 1 else "sampler1D"
    return _compile(tmp, "classic_as_cpp", "\n".join(parts) + CLASSIC_DRIVER.replace("@TF_TYPE@", tf_type))


def run_classic(exe, tmp, params, inv_proj, imv, step_scale, norm, domain_scale, light, bricks, n_bricks, brick_arrays, tf):
    w, h = params.width, params.height
    buf = [struct.pack("<II", w, h), np.asarray(inv_proj, np.float32).tobytes(), np.asarray(imv, np.float32).tobytes(),
           struct.pack("<5f", params.trans_scale, params.gradient_scale, step_scale, params.sample_rate_modifier, norm),
           np.asarray(domain_scale, np.float32).tobytes()]
    for k in ("ambient", "diffuse", "specular", "dir"):
        buf.append(np.asarray(light[k], np.float32).tobytes())
    buf.append(struct.pack("<5I", params.dtype, params.nearest, params.tf_w, params.tf_h, n_bricks))
    buf.append(np.ascontiguousarray(tf, np.uint8).tobytes())
    for i in range(n_bricks):
        b = bricks[i]
        buf.append(np.asarray(list(b.center) + list(b.ext) + list(b.tex_min) + list(b.tex_max), np.float32).tobytes())
        buf.append(struct.pack("<4I", b.n_vox[0], b.n_vox[1], b.n_vox[2], int(b.empty)))
        if not b.empty:
            buf.append(np.ascontiguousarray(brick_arrays[i]).tobytes())
    fin, fout = os.path.join(str(tmp), "classic.bin"), os.path.join(str(tmp), "classic_out.bin")
    with open(fin, "wb") as f:
        f.write(b"".join(buf))
    subprocess.check_call([exe, fin, fout])
    return np.fromfile(fout, np.float32).reshape(h * w, 4)


# ---------------------------------------------------------------------------------------------------------------------
# CPU baseline: the same executed reference GLSL, all host cores (what bench.py times as cpu_baseline kind "reference").
# The per-fragment globals become thread_local and the fragments of a frame are distributed with OpenMP; the binary
# takes the scene file of run(), renders `reps` frames and prints the seconds of the fastest one.
# ---------------------------------------------------------------------------------------------------------------------
def build_baseline(tmp, out_exe, mode, lighting, pool_glsl, hash_glsl):
    parts = [PRELUDE, rewrite(hash_glsl, tls=True), rewrite(pool_glsl, tls=True)]
    for n in ["Compositing.glsl", "lighting.glsl", "GLGridLeaper-GradientTools.glsl", METHOD[(mode, bool(lighting))],
              "GLGridLeaper-blend.glsl"]:
        parts.append("// ---- %s\n" % n + rewrite(read_shader(n), tls=True))
    drv = DRIVER.replace("@TF_TYPE@", "sampler2D" if mode == 1 else "sampler1D")
    drv = drv.replace("vec4 gl_FragCoord, accRayColor, rayResumeColor, rayResumePos; vec3 vPosInViewCoords;",
                      "thread_local vec4 gl_FragCoord, accRayColor, rayResumeColor, rayResumePos; thread_local vec3 vPosInViewCoords;")
    drv = drv.replace("#include <cstdio>", "#include <cstdio>\n#include <omp.h>")
    drv = drv.replace("  for (uint32_t y = 0; y < H; y++)\n    for (uint32_t x = 0; x < W; x++) {\n      const size_t i = (size_t)y * W + x;\n      if (!covered[i]) continue;",
                      "  const int reps = argc > 3 ? atoi(argv[3]) : 1;\n  double best = 1e30;\n  for (int rep = 0; rep < reps; rep++) {\n  const double t0 = omp_get_wtime();\n#pragma omp parallel for schedule(dynamic, 1)\n  for (int64_t y = 0; y < (int64_t)H; y++)\n    for (uint32_t x = 0; x < W; x++) {\n      const size_t i = (size_t)y * W + x;\n      if (!covered[i]) continue;")
    drv = drv.replace("      memcpy(&out[npx * 8 + i * 4], &rayResumePos.x, 16);\n    }\n",
                      "      memcpy(&out[npx * 8 + i * 4], &rayResumePos.x, 16);\n    }\n  const double dt = omp_get_wtime() - t0; if (dt < best) best = dt;\n  }\n  printf(\"%.9f %d\\n\", best, omp_get_max_threads());\n", 1)
    assert "omp_get_wtime() - t0" in drv and "#pragma omp parallel for" in drv
    src = os.path.join(str(tmp), "baseline_as_cpp.cpp")
    with open(src, "w") as f:
        f.write("\n".join(parts) + drv)
    subprocess.check_call(["g++", "-std=c++14", "-O3", "-fopenmp", "-mfma", "-ffp-contract=off", "-fno-fast-math", "-w", "-I", EMU,
                           "-o", out_exe, src])
    return out_exe


def scene_file(path, params, u, exit_eye, entry, start_color, covered, meta, meta_dim, atlas, tf):
    """Writes the input file of the DRIVER (same layout as run())."""
    w, h = params.width, params.height
    buf = [struct.pack("<II", w, h), np.asarray(u["emm"], np.float32).tobytes(),
           struct.pack("<5f", params.sample_rate_modifier, params.trans_scale, params.gradient_scale, params.lod_factor, u["lzwse"])]
    for k in ("ambient", "diffuse", "specular", "light_dir_m", "eye_m", "domain_scale"):
        buf.append(np.asarray(u[k], np.float32).tobytes())
    buf.append(struct.pack("<3I", *meta_dim))
    buf.append(struct.pack("<3I", atlas.shape[2], atlas.shape[1], atlas.shape[0]))
    buf.append(struct.pack("<5I", params.dtype, params.nearest, params.tf_w, params.tf_h, params.hash_size))
    buf.append(struct.pack("<f", u["norm"]))
    buf += [np.ascontiguousarray(entry, np.float32).tobytes(), np.ascontiguousarray(start_color, np.float32).tobytes(),
            np.ascontiguousarray(exit_eye, np.float32).tobytes(), np.ascontiguousarray(covered, np.uint8).tobytes()]
    m = np.zeros(meta_dim[0] * meta_dim[1] * meta_dim[2], np.uint32)
    m[:len(meta)] = meta
    buf += [m.tobytes(), np.ascontiguousarray(atlas).tobytes(), np.ascontiguousarray(tf, np.uint8).tobytes()]
    with open(path, "wb") as f:
        f.write(b"".join(buf))


# ---------------------------------------------------------------------------------------------------------------------
# HQ MIP frames: GLRaycaster-MIP-Rot-FS.glsl + Volume3D.glsl executed per brick with the pass setup of
# GLRaycaster::RenderHQMIPInLoop (front faces into the RGBA16F entry FBO, back-face fragments, BE_MAX blending), then
# Transfer-MIP-FS.glsl over the blended maximum image.  The driver is the classic one with the blend equation and
# the final pass exchanged.
# ---------------------------------------------------------------------------------------------------------------------
def _mip_driver():
    d = CLASSIC_DRIVER.replace("@TF_TYPE@", "sampler1D")
    subs = [
        ("sampler2D texRayExitPos, texRayExit;", "sampler2D texRayExitPos, texRayExit, texLast; vec4 gl_TexCoord[1];"),
        ("        classic_main();\n", "        mip_main();\n"),
        ("        float* dst = &out[4 * i];                               // GL blending ONE_MINUS_DST_ALPHA, ONE\n"
         "        const float k = 1.0f - dst[3];\n"
         "        dst[0] = fmaf(k, gl_FragColor.x, dst[0]); dst[1] = fmaf(k, gl_FragColor.y, dst[1]);\n"
         "        dst[2] = fmaf(k, gl_FragColor.z, dst[2]); dst[3] = fmaf(k, gl_FragColor.w, dst[3]);\n",
         "        float* dst = &out[4 * i];                               // GL blending BF_ONE, BE_MAX\n"
         "        dst[0] = fmaxf(dst[0], gl_FragColor.x); dst[1] = fmaxf(dst[1], gl_FragColor.y);\n"
         "        dst[2] = fmaxf(dst[2], gl_FragColor.z); dst[3] = fmaxf(dst[3], gl_FragColor.w);\n"),
        ("  FILE* f = fopen(argv[2], \"wb\");\n  fwrite(out.data(), 4, out.size(), f);\n",
         "  std::vector<float> fin(npx * 4, 0.0f);                       // Transfer-MIP-FS over a full-screen quad\n"
         "  texLast.f32 = out.data(); texLast.w = W; texLast.h = H;\n"
         "  for (uint32_t y = 0; y < H; y++)\n"
         "    for (uint32_t x = 0; x < W; x++) {\n"
         "      gl_TexCoord[0] = vec4(((float)x + 0.5f) / (float)W, ((float)y + 0.5f) / (float)H, 0.0f, 1.0f);\n"
         "      gl_FragColor = vec4();\n"
         "      transfer_main();\n"
         "      memcpy(&fin[4 * ((size_t)y * W + x)], &gl_FragColor.x, 16);\n"
         "    }\n"
         "  FILE* f = fopen(argv[2], \"wb\");\n  fwrite(fin.data(), 4, fin.size(), f);\n  fwrite(out.data(), 4, out.size(), f);\n"),
    ]
    for a, b in subs:
        assert a in d, a
        d = d.replace(a, b, 1)
    return d


def build_mip(tmp):
    pre = PRELUDE + ("extern vec4 gl_FragCoord, gl_FragColor, gl_TexCoord[1]; extern mat4x4 gl_TextureMatrix[1]; "
                     "extern mat3 gl_NormalMatrix;\n")
    parts = [pre]
    for n, main in (("Volume3D.glsl", "unused_main"), ("GLRaycaster-MIP-Rot-FS.glsl", "mip_main"),
                    ("Transfer-MIP-FS.glsl", "transfer_main")):
        parts.append("// ---- %s\n" % n + rewrite(read_shader(n), main))
    return _compile(tmp, "mip_as_cpp", "\n".join(parts) + _mip_driver())


def run_mip(exe, tmp, params, inv_proj, imv, norm, bricks, n_bricks, brick_arrays, tf1d):
    """Returns (rgba [h*w, 4], blended maximum image [h*w, 4] = (max, max, max, coverage))."""
    w, h = params.width, params.height
    zero3 = [0.0, 0.0, 0.0]
    tf = np.ascontiguousarray(tf1d, np.uint8).reshape(-1, 4)
    buf = [struct.pack("<II", w, h), np.asarray(inv_proj, np.float32).tobytes(), np.asarray(imv, np.float32).tobytes(),
           struct.pack("<5f", params.trans_scale, 1.0, 1.0, params.sample_rate_modifier, norm)]
    for _ in range(5):                                             # domain scale + light terms: unused by the MIP shaders
        buf.append(np.asarray(zero3, np.float32).tobytes())
    buf.append(struct.pack("<5I", params.dtype, params.nearest, len(tf), 1, n_bricks))
    buf.append(tf.tobytes())
    for i in range(n_bricks):
        b = bricks[i]
        buf.append(np.asarray(list(b.center) + list(b.ext) + list(b.tex_min) + list(b.tex_max), np.float32).tobytes())
        buf.append(struct.pack("<4I", b.n_vox[0], b.n_vox[1], b.n_vox[2], int(b.empty)))
        if not b.empty:
            buf.append(np.ascontiguousarray(brick_arrays[i]).tobytes())
    fin, fout = os.path.join(str(tmp), "mip.bin"), os.path.join(str(tmp), "mip_out.bin")
    with open(fin, "wb") as f:
        f.write(b"".join(buf))
    subprocess.check_call([exe, fin, fout])
    raw = np.fromfile(fout, np.float32).reshape(2, h * w, 4)
    return raw[0], raw[1]


# ---------------------------------------------------------------------------------------------------------------------
# Stereo eye composition (GLRenderer::EndFrame, GLRenderer.cpp:758-812): Compose-{Anaglyphs,Scanline,SBS,AF}-FS.glsl
# executed over a full-screen quad on two RGBA32F eye images (GL_NEAREST FBOs).
# ---------------------------------------------------------------------------------------------------------------------
STEREO_DRIVER = r"""
sampler2D texRightEye, texLeftEye; vec2 vScreensize; float fSplitCoord; int iAlternatingFrameID;
vec4 gl_FragColor, gl_TexCoord[1];
#include <cstdio>
#include <cstdlib>
int main(int argc, char** argv) {
  FILE* f = fopen(argv[1], "rb");
  uint32_t hd[5]; float split;
  if (fread(hd, 4, 5, f) != 5 || fread(&split, 4, 1, f) != 1) abort();
  const uint32_t W = hd[0], H = hd[1], mode = hd[2], swap = hd[3];
  const size_t n = (size_t)W * H * 4;
  std::vector<float> l(n), r(n), out(n);
  if (fread(l.data(), 4, n, f) != n || fread(r.data(), 4, n, f) != n) abort();
  fclose(f);
  // m_bStereoEyeSwap exchanges the texture units the two FBOs are bound to (GLRenderer.cpp:773-779)
  texLeftEye.f32 = swap ? r.data() : l.data(); texRightEye.f32 = swap ? l.data() : r.data();
  texLeftEye.w = texRightEye.w = W; texLeftEye.h = texRightEye.h = H;
  vScreensize = vec2((float)W, (float)H); fSplitCoord = split; iAlternatingFrameID = (int)hd[4];
  for (uint32_t y = 0; y < H; y++)
    for (uint32_t x = 0; x < W; x++) {
      gl_TexCoord[0] = vec4(((float)x + 0.5f) / (float)W, ((float)y + 0.5f) / (float)H, 0.0f, 1.0f);
      gl_FragColor = vec4();
      if (mode == 0) rb_main(); else if (mode == 1) scan_main(); else if (mode == 2) sbs_main(); else af_main();
      memcpy(&out[4 * ((size_t)y * W + x)], &gl_FragColor.x, 16);
    }
  f = fopen(argv[2], "wb"); fwrite(out.data(), 4, n, f); fclose(f);
  return 0;
}
"""


def build_stereo(tmp):
    parts = [PRELUDE + "extern vec4 gl_FragColor, gl_TexCoord[1];\n"]
    for n, main in (("Compose-Anaglyphs-FS.glsl", "rb_main"), ("Compose-Scanline-FS.glsl", "scan_main"),
                    ("Compose-SBS-FS.glsl", "sbs_main"), ("Compose-AF-FS.glsl", "af_main")):
        parts.append("// ---- %s\n" % n + rewrite(read_shader(n), main))
    return _compile(tmp, "stereo_as_cpp", "\n".join(parts) + STEREO_DRIVER)


def run_stereo(exe, tmp, mode, left, right, eye_swap=False, alternating_frame_id=0, split_coord=0.5):
    h, w = left.shape[:2]
    fin, fout = os.path.join(str(tmp), "stereo.bin"), os.path.join(str(tmp), "stereo_out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("<5If", w, h, mode, int(bool(eye_swap)), alternating_frame_id, split_coord))
        f.write(np.ascontiguousarray(left, np.float32).tobytes())
        f.write(np.ascontiguousarray(right, np.float32).tobytes())
    subprocess.check_call([exe, fin, fout])
    return np.fromfile(fout, np.float32).reshape(h, w, 4)


# ---------------------------------------------------------------------------------------------------------------------
# Classic isosurface frames: GLRaycaster-ISO-FS.glsl + RefineIsosurface.glsl + Volume3D.glsl executed per
# brick with the pass setup of GLRaycaster::Render3DInLoop's RM_ISOSURFACE branch: two float targets, gl_FragDepth
# under the depth test DF_LESS (cleared to 1), no blending.  The classic driver with the output stage exchanged.
# ---------------------------------------------------------------------------------------------------------------------
def _classic_iso_driver():
    d = CLASSIC_DRIVER.replace("@TF_TYPE@", "sampler1D")
    subs = [
        ("vec4 gl_FragCoord, gl_FragColor; mat4x4 gl_TextureMatrix[1]; mat3 gl_NormalMatrix;",
         "vec4 gl_FragCoord, gl_FragColor, gl_FragData[2]; mat4x4 gl_TextureMatrix[1]; mat3 gl_NormalMatrix;\n"
         "float fIsoval, gl_FragDepth; vec2 vProjParam; int iTileID; bool g_discarded;"),
        ("  ScaleMethod = 0; TFuncBias = 0.0f;\n",
         "  ScaleMethod = 0; TFuncBias = 0.0f;\n"
         "  fIsoval = fTransScale; vProjParam = vec2(fGradientScale, fStepScale);   // carried in the unused header slots\n"),
        ("  std::vector<float> out(npx * 4, 0.0f), fbo(npx * 4, 0.0f), near_pt(npx * 3), org_pt(npx * 3, 0.0f), dir_pt(npx * 3);",
         "  std::vector<float> out(npx * 4, 0.0f), out2(npx * 4, 0.0f), depthb(npx, 1.0f), fbo(npx * 4, 0.0f), near_pt(npx * 3), org_pt(npx * 3, 0.0f), dir_pt(npx * 3);"),
        ("        gl_FragColor = vec4();\n        classic_main();\n",
         "        gl_FragData[0] = vec4(); gl_FragData[1] = vec4(); g_discarded = false; iTileID = (int)bi;\n        iso_main();\n"),
        ("        float* dst = &out[4 * i];                               // GL blending ONE_MINUS_DST_ALPHA, ONE\n"
         "        const float k = 1.0f - dst[3];\n"
         "        dst[0] = fmaf(k, gl_FragColor.x, dst[0]); dst[1] = fmaf(k, gl_FragColor.y, dst[1]);\n"
         "        dst[2] = fmaf(k, gl_FragColor.z, dst[2]); dst[3] = fmaf(k, gl_FragColor.w, dst[3]);\n",
         "        if (g_discarded) continue;\n"
         "        const float dz = fminf(fmaxf(gl_FragDepth, 0.0f), 1.0f);   // depth range clamp, then DF_LESS\n"
         "        if (!(dz < depthb[i])) continue;\n"
         "        depthb[i] = dz;\n"
         "        memcpy(&out[4 * i], &gl_FragData[0].x, 16); memcpy(&out2[4 * i], &gl_FragData[1].x, 16);\n"),
        ("  fwrite(out.data(), 4, out.size(), f);\n", "  fwrite(out.data(), 4, out.size(), f);\n  fwrite(out2.data(), 4, out2.size(), f);\n"),
    ]
    for a, b in subs:
        assert a in d, a
        d = d.replace(a, b, 1)
    return d


def build_classic_iso(tmp):
    pre = PRELUDE + ("#define discard { g_discarded = true; return; }\n"
                     "extern vec4 gl_FragCoord, gl_FragColor, gl_FragData[2]; extern mat4x4 gl_TextureMatrix[1]; extern mat3 gl_NormalMatrix;\n"
                     "extern float gl_FragDepth; extern bool g_discarded;\n")
    parts = [pre]
    for n, main in (("Volume3D.glsl", "unused_main"), ("RefineIsosurface.glsl", "unused_main2"), ("GLRaycaster-ISO-FS.glsl", "iso_main")):
        parts.append("// ---- %s\n" % n + rewrite(read_shader(n), main))
    src = "\n".join(parts)
    # a swizzle passed as an `inout` argument: GLSL copies in and out, a C++ reference cannot bind to the proxy
    call = "RefineIsosurface(vRayIncTex, vHitPosTex.xyz, fIsoval)"
    assert src.count(call) == 1
    src = src.replace(call, "[&] { vec3 io_ = vHitPosTex.xyz; vec3 r_ = RefineIsosurface(vRayIncTex, io_, fIsoval); "
                            "vHitPosTex.xyz = io_; return r_; }()")
    return _compile(tmp, "classic_iso_as_cpp", src + _classic_iso_driver())


def run_classic_iso(exe, tmp, params, inv_proj, imv, norm, domain_scale, proj_param, bricks, n_bricks, brick_arrays):
    """Returns (hit_pos [h*w, 4], hit_normal [h*w, 4]) of the executed GLRaycaster-ISO-FS brick loop."""
    w, h = params.width, params.height
    zero3 = [0.0, 0.0, 0.0]
    buf = [struct.pack("<II", w, h), np.asarray(inv_proj, np.float32).tobytes(), np.asarray(imv, np.float32).tobytes(),
           struct.pack("<5f", params.isoval, proj_param[0], proj_param[1], params.sample_rate_modifier, norm),
           np.asarray(domain_scale, np.float32).tobytes()]
    for _ in range(4):
        buf.append(np.asarray(zero3, np.float32).tobytes())
    buf.append(struct.pack("<5I", params.dtype, params.nearest, 1, 1, n_bricks))
    buf.append(np.zeros(4, np.uint8).tobytes())                    # a 1-texel transfer function nobody reads
    for i in range(n_bricks):
        b = bricks[i]
        buf.append(np.asarray(list(b.center) + list(b.ext) + list(b.tex_min) + list(b.tex_max), np.float32).tobytes())
        buf.append(struct.pack("<4I", b.n_vox[0], b.n_vox[1], b.n_vox[2], int(b.empty)))
        if not b.empty:
            buf.append(np.ascontiguousarray(brick_arrays[i]).tobytes())
    fin, fout = os.path.join(str(tmp), "ciso.bin"), os.path.join(str(tmp), "ciso_out.bin")
    with open(fin, "wb") as f:
        f.write(b"".join(buf))
    subprocess.check_call([exe, fin, fout])
    raw = np.fromfile(fout, np.float32).reshape(2, h * w, 4)
    return raw[0], raw[1]


# ---------------------------------------------------------------------------------------------------------------------
# ClearView: after a brick's GLRaycaster-ISO-FS pass the same back faces run GLRaycaster-ISO-CV-FS.glsl
# into m_pFBOCVHit (reading the first pass's targets as texLastHit / texLastHitPos), and Compose-CV-FS.glsl shades the
# four targets.  The classic isosurface driver with the second pass and the composition added; the ClearView
# parameters travel in the (otherwise unused) transfer-function block of the scene file.
# ---------------------------------------------------------------------------------------------------------------------
def _classic_cv_driver():
    d = _classic_iso_driver()
    subs = [
        ("float fIsoval, gl_FragDepth; vec2 vProjParam; int iTileID; bool g_discarded;",
         "float fIsoval, gl_FragDepth; vec2 vProjParam; int iTileID; bool g_discarded;\n"
         "sampler2D texLastHit, texLastHitPos, texRayHitPos, texRayHitNormal, texRayHitPos2, texRayHitNormal2;\n"
         "vec3 vLightDiffuse2, vCVParam, vCVPickPos;"),
        ("  std::vector<float> out(npx * 4, 0.0f), out2(npx * 4, 0.0f), depthb(npx, 1.0f),",
         "  const float* cvp = (const float*)texTrans.rgba8;              // cv isovalue, diffuse2, (size, context, border), pick\n"
         "  const float iso1 = fIsoval, iso2 = cvp[0];\n"
         "  vLightDiffuse2 = vec3(cvp[1], cvp[2], cvp[3]); vCVParam = vec3(cvp[4], cvp[5], cvp[6]); vCVPickPos = vec3(cvp[7], cvp[8], cvp[9]);\n"
         "  std::vector<float> out3(npx * 4, 0.0f), out4(npx * 4, 0.0f), depthb2(npx, 1.0f), fin(npx * 4, 0.0f);\n"
         "  std::vector<float> out(npx * 4, 0.0f), out2(npx * 4, 0.0f), depthb(npx, 1.0f),"),
        ("        gl_FragData[0] = vec4(); gl_FragData[1] = vec4(); g_discarded = false; iTileID = (int)bi;\n        iso_main();\n"
         "        frags++;\n",
         "        gl_FragData[0] = vec4(); gl_FragData[1] = vec4(); g_discarded = false; iTileID = (int)bi; fIsoval = iso1;\n        iso_main();\n"
         "        frags++;\n"),
        ("        if (g_discarded) continue;\n"
         "        const float dz = fminf(fmaxf(gl_FragDepth, 0.0f), 1.0f);   // depth range clamp, then DF_LESS\n"
         "        if (!(dz < depthb[i])) continue;\n"
         "        depthb[i] = dz;\n"
         "        memcpy(&out[4 * i], &gl_FragData[0].x, 16); memcpy(&out2[4 * i], &gl_FragData[1].x, 16);\n",
         "        if (!g_discarded) {\n"
         "          const float dz = fminf(fmaxf(gl_FragDepth, 0.0f), 1.0f);   // depth range clamp, then DF_LESS\n"
         "          if (dz < depthb[i]) { depthb[i] = dz; memcpy(&out[4 * i], &gl_FragData[0].x, 16); memcpy(&out2[4 * i], &gl_FragData[1].x, 16); }\n"
         "        }\n"
         "        // second pass of the same brick (GLRaycaster.cpp:429-444): m_pFBOIsoHit bound as texLastHit / texLastHitPos\n"
         "        texLastHit.f32 = out.data(); texLastHit.w = W; texLastHit.h = H;\n"
         "        texLastHitPos.f32 = out2.data(); texLastHitPos.w = W; texLastHitPos.h = H;\n"
         "        gl_FragData[0] = vec4(); gl_FragData[1] = vec4(); g_discarded = false; fIsoval = iso2;\n"
         "        cv_main();\n"
         "        if (!g_discarded) {\n"
         "          const float dz = fminf(fmaxf(gl_FragDepth, 0.0f), 1.0f);\n"
         "          if (dz < depthb2[i]) { depthb2[i] = dz; memcpy(&out3[4 * i], &gl_FragData[0].x, 16); memcpy(&out4[4 * i], &gl_FragData[1].x, 16); }\n"
         "        }\n"),
        ("  fwrite(out.data(), 4, out.size(), f);\n  fwrite(out2.data(), 4, out2.size(), f);\n",
         "  texRayHitPos.f32 = out.data(); texRayHitNormal.f32 = out2.data(); texRayHitPos2.f32 = out3.data(); texRayHitNormal2.f32 = out4.data();\n"
         "  texRayHitPos.w = texRayHitNormal.w = texRayHitPos2.w = texRayHitNormal2.w = W;\n"
         "  texRayHitPos.h = texRayHitNormal.h = texRayHitPos2.h = texRayHitNormal2.h = H;\n"
         "  for (uint32_t y = 0; y < H; y++)\n"
         "    for (uint32_t x = 0; x < W; x++) {                          // GLRenderer::ComposeSurfaceImage, ClearView branch\n"
         "      gl_FragCoord = vec4((float)x + 0.5f, (float)y + 0.5f, 0.0f, 1.0f);\n"
         "      gl_FragColor = vec4(); g_discarded = false;\n"
         "      compose_cv_main();\n"
         "      if (!g_discarded) memcpy(&fin[4 * ((size_t)y * W + x)], &gl_FragColor.x, 16);\n"
         "    }\n"
         "  fwrite(out.data(), 4, out.size(), f);\n  fwrite(out2.data(), 4, out2.size(), f);\n"
         "  fwrite(out3.data(), 4, out3.size(), f);\n  fwrite(out4.data(), 4, out4.size(), f);\n  fwrite(fin.data(), 4, fin.size(), f);\n"),
    ]
    for a, b in subs:
        assert a in d, a
        d = d.replace(a, b, 1)
    # the FILE* f of the output stage is opened before the composition loop writes: move the declaration up
    return d


def build_classic_cv(tmp):
    pre = PRELUDE + ("#define discard { g_discarded = true; return; }\n"
                     "extern vec4 gl_FragCoord, gl_FragColor, gl_FragData[2]; extern mat4x4 gl_TextureMatrix[1]; extern mat3 gl_NormalMatrix;\n"
                     "extern float gl_FragDepth; extern bool g_discarded;\n")
    parts = [pre]
    for n, main in (("Volume3D.glsl", "unused_main"), ("RefineIsosurface.glsl", "unused_main2"), ("GLRaycaster-ISO-FS.glsl", "iso_main"),
                    ("GLRaycaster-ISO-CV-FS.glsl", "cv_main"), ("Compose-CV-FS.glsl", "compose_cv_main")):
        parts.append("// ---- %s\n" % n + rewrite(read_shader(n), main))
    src = "\n".join(parts)
    call = "RefineIsosurface(vRayIncTex, vHitPosTex.xyz, fIsoval)"      # inout swizzle argument: copy in, copy out
    assert src.count(call) == 2
    src = src.replace(call, "[&] { vec3 io_ = vHitPosTex.xyz; vec3 r_ = RefineIsosurface(vRayIncTex, io_, fIsoval); "
                            "vHitPosTex.xyz = io_; return r_; }()")
    return _compile(tmp, "classic_cv_as_cpp", src + _classic_cv_driver())


def run_classic_cv(exe, tmp, params, inv_proj, imv, norm, domain_scale, proj_param, light, cv_isoval, diffuse2, cv_param, pick,
                   bricks, n_bricks, brick_arrays):
    """Returns (hit_pos, hit_normal, cv_pos, cv_normal, rgba), each [h*w, 4]."""
    w, h = params.width, params.height
    buf = [struct.pack("<II", w, h), np.asarray(inv_proj, np.float32).tobytes(), np.asarray(imv, np.float32).tobytes(),
           struct.pack("<5f", params.isoval, proj_param[0], proj_param[1], params.sample_rate_modifier, norm),
           np.asarray(domain_scale, np.float32).tobytes()]
    for k in ("ambient", "diffuse", "specular", "dir"):
        buf.append(np.asarray(light[k], np.float32).tobytes())
    cv = np.zeros(12, np.float32)
    cv[0] = cv_isoval; cv[1:4] = diffuse2; cv[4:7] = cv_param; cv[7:10] = pick
    buf.append(struct.pack("<5I", params.dtype, params.nearest, 12, 1, n_bricks))
    buf.append(cv.tobytes())                                       # 12 "texels" of the transfer-function block
    for i in range(n_bricks):
        b = bricks[i]
        buf.append(np.asarray(list(b.center) + list(b.ext) + list(b.tex_min) + list(b.tex_max), np.float32).tobytes())
        buf.append(struct.pack("<4I", b.n_vox[0], b.n_vox[1], b.n_vox[2], int(b.empty)))
        if not b.empty:
            buf.append(np.ascontiguousarray(brick_arrays[i]).tobytes())
    fin, fout = os.path.join(str(tmp), "ccv.bin"), os.path.join(str(tmp), "ccv_out.bin")
    with open(fin, "wb") as f:
        f.write(b"".join(buf))
    subprocess.check_call([exe, fin, fout])
    raw = np.fromfile(fout, np.float32).reshape(5, h * w, 4)
    return raw[0], raw[1], raw[2], raw[3], raw[4]
