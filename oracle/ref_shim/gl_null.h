// gl_null.h -- test-side accessors of the recording GL stand-in (see gl_null.cpp).
#pragma once
#include <cstdint>
void glnull_set_max_3d(int v);   // what glGetIntegerv(GL_MAX_3D_TEXTURE_SIZE) answers
const uint8_t* glnull_texture(unsigned gl_id, uint32_t dim[3], uint32_t* bytes_per_texel);
