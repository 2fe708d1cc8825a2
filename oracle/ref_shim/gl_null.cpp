// gl_null -- a RECORDING stand-in for the handful of OpenGL entry points the reference's
// GLVolumePool / GLTexture3D reach (there is no GL in this image).  Texture storage is kept on the
// host so that oracle/_ref/ref_pool can dump what the shader WOULD see: the R32UI metadata texture
// (the page table) and the brick-pool atlas, exactly as the unmodified reference code uploaded them
// with glTexImage3D / glTexSubImage3D.  Test infrastructure only.
#include <GL/glew.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>
#include "gl_null.h"

namespace {
struct Tex {
  uint32_t w = 0, h = 0, d = 0, bpt = 0;   // bytes per texel
  std::vector<uint8_t> data;
};
std::map<GLuint, Tex> g_tex;
GLuint g_next = 1, g_bound = 0;
int g_max3d = 2048, g_max2d = 16384;

uint32_t texel_bytes(GLenum format, GLenum type) {
  uint32_t comps = 1;
  switch (format) {
    case GL_RGB: comps = 3; break;
    case GL_RGBA: comps = 4; break;
    default: comps = 1; break;   // GL_LUMINANCE, GL_RED_INTEGER
  }
  uint32_t b = 1;
  switch (type) {
    case GL_UNSIGNED_SHORT: case GL_SHORT: b = 2; break;
    case GL_UNSIGNED_INT: case GL_INT: case GL_FLOAT: b = 4; break;
    default: b = 1; break;
  }
  return comps * b;
}

void GLAPIENTRY nullActiveTexture(GLenum) {}
void GLAPIENTRY nullBindImageTexture(GLuint, GLuint, GLint, GLboolean, GLint, GLenum, GLenum) {}
void GLAPIENTRY nullMemoryBarrier(GLbitfield) {}

void GLAPIENTRY recTexImage3D(GLenum, GLint, GLint, GLsizei w, GLsizei h, GLsizei d, GLint, GLenum format,
                              GLenum type, const void* pixels) {
  Tex& t = g_tex[g_bound];
  t.w = w; t.h = h; t.d = d; t.bpt = texel_bytes(format, type);
  t.data.assign(size_t(w) * h * d * t.bpt, 0);
  if (pixels) memcpy(t.data.data(), pixels, t.data.size());
}

void GLAPIENTRY recTexSubImage3D(GLenum, GLint, GLint x, GLint y, GLint z, GLsizei w, GLsizei h, GLsizei d,
                                 GLenum format, GLenum type, const void* pixels) {
  Tex& t = g_tex[g_bound];
  const uint32_t bpt = texel_bytes(format, type);
  if (bpt != t.bpt || uint32_t(x + w) > t.w || uint32_t(y + h) > t.h || uint32_t(z + d) > t.d) {
    fprintf(stderr, "gl_null: glTexSubImage3D out of range / format mismatch\n");
    abort();
  }
  const uint8_t* src = static_cast<const uint8_t*>(pixels);
  for (GLsizei k = 0; k < d; k++)
    for (GLsizei j = 0; j < h; j++)
      memcpy(&t.data[((size_t(z + k) * t.h + (y + j)) * t.w + x) * bpt], src + (size_t(k) * h + j) * w * bpt,
             size_t(w) * bpt);
}
}  // namespace

extern "C" {
PFNGLACTIVETEXTUREPROC __glewActiveTexture = nullActiveTexture;
PFNGLTEXIMAGE3DPROC __glewTexImage3D = recTexImage3D;
PFNGLBINDIMAGETEXTUREPROC __glewBindImageTexture = nullBindImageTexture;
PFNGLMEMORYBARRIERPROC __glewMemoryBarrier = nullMemoryBarrier;
PFNGLTEXSUBIMAGE3DPROC __glewTexSubImage3D = recTexSubImage3D;

void GLAPIENTRY glBindTexture(GLenum, GLuint id) { g_bound = id; }
void GLAPIENTRY glDeleteTextures(GLsizei n, const GLuint* ids) { for (GLsizei i = 0; i < n; i++) g_tex.erase(ids[i]); }
void GLAPIENTRY glGenTextures(GLsizei n, GLuint* ids) { for (GLsizei i = 0; i < n; i++) { ids[i] = g_next++; g_tex[ids[i]]; } }
GLenum GLAPIENTRY glGetError(void) { return GL_NO_ERROR; }
void GLAPIENTRY glGetIntegerv(GLenum pname, GLint* v) {
  *v = (pname == GL_MAX_3D_TEXTURE_SIZE_EXT) ? g_max3d : (pname == GL_MAX_TEXTURE_SIZE) ? g_max2d
     : (pname == GL_TEXTURE_BINDING_3D || pname == GL_TEXTURE_BINDING_2D || pname == GL_TEXTURE_BINDING_1D) ? GLint(g_bound) : 0;
}
// 1D / 2D textures (the miss-report hash table) live in the same store with d = 1 (and h = 1)
void GLAPIENTRY glTexImage2D(GLenum, GLint, GLint, GLsizei w, GLsizei h, GLint, GLenum format, GLenum type, const void* px) {
  recTexImage3D(0, 0, 0, w, h, 1, 0, format, type, px);
}
void GLAPIENTRY glTexImage1D(GLenum, GLint, GLint, GLsizei w, GLint, GLenum format, GLenum type, const void* px) {
  recTexImage3D(0, 0, 0, w, 1, 1, 0, format, type, px);
}
void GLAPIENTRY glTexSubImage2D(GLenum, GLint, GLint x, GLint y, GLsizei w, GLsizei h, GLenum format, GLenum type, const void* px) {
  recTexSubImage3D(0, 0, x, y, 0, w, h, 1, format, type, px);
}
void GLAPIENTRY glTexSubImage1D(GLenum, GLint, GLint x, GLsizei w, GLenum format, GLenum type, const void* px) {
  recTexSubImage3D(0, 0, x, 0, 0, w, 1, 1, format, type, px);
}
void GLAPIENTRY glGetTexImage(GLenum, GLint, GLenum, GLenum, void* dst) {
  const Tex& t = g_tex[g_bound];
  memcpy(dst, t.data.data(), t.data.size());
}
void GLAPIENTRY glPixelStorei(GLenum, GLint) {}
void GLAPIENTRY glTexParameteri(GLenum, GLenum, GLint) {}
}

void glnull_set_max_3d(int v) { g_max3d = v; }
void glnull_set_max_2d(int v) { g_max2d = v; }
void glnull_write_texture(unsigned id, const void* src, size_t bytes) {
  auto it = g_tex.find(id);
  if (it != g_tex.end()) memcpy(it->second.data.data(), src, bytes < it->second.data.size() ? bytes : it->second.data.size());
}
const uint8_t* glnull_texture(unsigned id, uint32_t dim[3], uint32_t* bytes_per_texel) {
  auto it = g_tex.find(id);
  if (it == g_tex.end()) return nullptr;
  dim[0] = it->second.w; dim[1] = it->second.h; dim[2] = it->second.d;
  *bytes_per_texel = it->second.bpt;
  return it->second.data.data();
}
