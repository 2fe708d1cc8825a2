// glsl_stub -- link-time stand-ins for the GLSLProgram members GLVolumePool::Enable / GLTexture::Bind reference.
// ref_pool never calls Enable (no shader exists without GL); these only satisfy the linker.
#include "Renderer/GL/GLSLProgram.h"
namespace tuvok {
void GLSLProgram::Enable() {}
void GLSLProgram::Set(const char*, float) const {}
void GLSLProgram::SetTexture(const std::string&, const GLTexture&) {}
}
