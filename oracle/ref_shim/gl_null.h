// gl_null.h -- test-side accessors of the recording GL stand-in (see gl_null.cpp).
#pragma once
#include <cstddef>
#include <cstdint>
void glnull_set_max_3d(int v);   // what glGetIntegerv(GL_MAX_3D_TEXTURE_SIZE) answers
const uint8_t* glnull_texture(unsigned gl_id, uint32_t dim[3], uint32_t* bytes_per_texel);
void glnull_set_max_2d(int v);   // GL_MAX_TEXTURE_SIZE
void glnull_write_texture(unsigned gl_id, const void* src, size_t bytes);   // what a shader's image stores left behind
