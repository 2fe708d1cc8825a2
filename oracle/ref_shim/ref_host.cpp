// ref_host -- runs the UNMODIFIED reference host-side arithmetic that feeds the ray caster: the matrix
// conventions of Basics/Vectors.h (BuildLookAt, Perspective, operator*, inverse), the frustum culling /
// LOD selection of Renderer/CullingLOD.cpp and the 1D transfer function of IO/TransferFunction1D.cpp
// (SetStdFunction, GetByteArray, non-zero limits).  Compiled in place from /root/reference by
// oracle/Makefile; tests/test_host_ref.py compares the oracle restatement and tvk_compute_view with it.
// Floats are printed as C99 hex floats (%a) so the comparison is bit-exact.  Test infrastructure only.
//
//   ref_host <commands.txt> <result.txt>
//     view  ex ey ez  ax ay az  ux uy uz  fov aspect near far pixels_y
//     mul   a[16] b[16]                       (Tuvok row-vector product a*b, FLOATMATRIX4::operator*)
//     inverse m[16]
//     cull  fov aspect near far pixels_y  model[16] view[16] proj[16]  n  (cx cy cz ex ey ez vx vy vz)*n
//     tf1d  n center inv_gradient
//     uniforms mv[16] vol[3] scale[3] light_dir[3] eye[3]   (SetupRaycastShader / ComputeEyeToModelMatrix)
//     stereo ex ey ez  ax ay az  ux uy uz  fov aspect near far focal_length eye_dist
//           (FLOATMATRIX4::BuildStereoLookAtAndProjection as GLRenderer::ComputeViewAndProjection calls it)
//     clipbox plane_world[4] rotation[16] translation[16] extent[3]
//           (GLGridLeaper::FillBBoxVBO, GLGridLeaper.cpp:506-532: Plane() * inverse(rotation*translation), normal normalised,
//            then Clipper::BoxPlane on the 12 triangles of the box [-extent/2, extent/2]; prints the plane and the triangles)
//     miprot window(0 sagittal,1 axial,2 coronal) flipx flipy angle_deg region_rotation[16] view[16]
//     mipo width height                     (the parallel projection of an HQ MIP frame, m_bOrthoView)
//           (the statements of GLRenderer::RenderHQMIPPreLoop + GLRaycaster::RenderHQMIPPreLoop on FLOATMATRIX4)
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "StdTuvokDefines.h"
#include <GL/glew.h>   // defines __GL_H__: Vectors.h only declares BuildLookAt / Perspective next to its GL helpers
#include "Basics/Vectors.h"
#include "IO/TransferFunction1D.h"
#include "Renderer/CullingLOD.h"
#include "Basics/Clipper.h"

using namespace tuvok;

static void read16(std::istream& s, FLOATMATRIX4& m) { for (int i = 0; i < 16; i++) s >> m.array[i]; }
static void put16(FILE* o, const char* tag, const FLOATMATRIX4& m) {
  fprintf(o, "%s", tag);
  for (int i = 0; i < 16; i++) fprintf(o, " %a", (double)m.array[i]);
  fprintf(o, "\n");
}

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: ref_host commands.txt result.txt\n"); return 2; }
  std::ifstream in(argv[1]);
  FILE* out = fopen(argv[2], "w");
  if (!in || !out) return 2;
  std::string line;
  while (std::getline(in, line)) {
    std::istringstream ls(line);
    std::string op;
    if (!(ls >> op)) continue;
    if (op == "view") {
      FLOATVECTOR3 e, a, u; float fov, aspect, n, f; uint32_t py;
      ls >> e.x >> e.y >> e.z >> a.x >> a.y >> a.z >> u.x >> u.y >> u.z >> fov >> aspect >> n >> f >> py;
      FLOATMATRIX4 view, proj;
      view.BuildLookAt(e, a, u);            // GLRenderer::ComputeViewAndProjection, GLRenderer.cpp:908-912
      proj.Perspective(fov, aspect, n, f);
      CullingLOD c;
      c.SetScreenParams(fov, aspect, n, f, py);   // GLRenderer::SetViewPort, GLRenderer.cpp:886-888
      put16(out, "view", view);
      put16(out, "proj", proj);
      fprintf(out, "lodfactor %a\n", (double)c.GetLoDFactor());
    } else if (op == "mul") {
      FLOATMATRIX4 a, b; read16(ls, a); read16(ls, b);
      put16(out, "mul", a * b);
    } else if (op == "stereo") {
      FLOATVECTOR3 e, a, u; float fov, aspect, n, f, focal, dist;
      ls >> e.x >> e.y >> e.z >> a.x >> a.y >> a.z >> u.x >> u.y >> u.z >> fov >> aspect >> n >> f >> focal >> dist;
      FLOATMATRIX4 vl, vr, pl, pr;
      FLOATMATRIX4::BuildStereoLookAtAndProjection(e, a, u, fov, aspect, n, f, focal, dist, vl, vr, pl, pr);   // GLRenderer.cpp:905-910
      put16(out, "view_left", vl); put16(out, "view_right", vr);
      put16(out, "proj_left", pl); put16(out, "proj_right", pr);
    } else if (op == "miprot") {
      int wm, fx, fy; float angle;
      ls >> wm >> fx >> fy >> angle;
      FLOATMATRIX4 region_rotation, view; read16(ls, region_rotation); read16(ls, view);
      // GLRenderer::RenderHQMIPPreLoop, GLRenderer.cpp:1256-1285
      double dPI = 3.141592653589793238462643383;
      FLOATMATRIX4 matRotDir, matFlipX, matFlipY, maMIPRotation;
      if (wm == 0) { FLOATMATRIX4 matTemp; matRotDir.RotationX(-dPI/2.0); matTemp.RotationY(-dPI/2.0); matRotDir = matRotDir * matTemp; }
      else if (wm == 1) matRotDir.RotationX(-dPI/2.0);
      if (fx) matFlipY.Scaling(-1,1,1);
      if (fy) matFlipX.Scaling(1,-1,1);
      maMIPRotation.RotationY(dPI*double(angle)/180.0);
      maMIPRotation = matRotDir * region_rotation * matFlipX * matFlipY * maMIPRotation;
      put16(out, "miprot", maMIPRotation);
      put16(out, "mipmv", maMIPRotation * view);   // GLRaycaster::RenderHQMIPPreLoop, GLRaycaster.cpp:489 (perspective)
    } else if (op == "mipo") {
      // the parallel projection of an HQ MIP frame under m_bOrthoView: the statement sequence of GLRenderer.cpp:1183-1197 on
      // the reference's own vector / matrix classes
      unsigned w, h;
      ls >> w >> h;
      UINTVECTOR2 m_vWinSize(w, h);
      FLOATMATRIX4 maOrtho;
      DOUBLEVECTOR2 vWinAspectRatio = 1.0 / DOUBLEVECTOR2(m_vWinSize);
      vWinAspectRatio = vWinAspectRatio / vWinAspectRatio.maxVal();
      float fRoot2Scale = (vWinAspectRatio.x < vWinAspectRatio.y) ?
                          std::max(1.0f, 1.414213f * float(vWinAspectRatio.x/vWinAspectRatio.y)) :
                          1.414213f;
      maOrtho.Ortho(-0.5f*fRoot2Scale/float(vWinAspectRatio.x),
                    +0.5f*fRoot2Scale/float(vWinAspectRatio.x),
                    -0.5f*fRoot2Scale/float(vWinAspectRatio.y),
                    +0.5f*fRoot2Scale/float(vWinAspectRatio.y),
                    -100.0f, 100.0f);
      put16(out, "mipo", maOrtho);
    } else if (op == "clipbox") {
      float pw[4]; FLOATMATRIX4 rot, tra; FLOATVECTOR3 ext;
      ls >> pw[0] >> pw[1] >> pw[2] >> pw[3]; read16(ls, rot); read16(ls, tra); ls >> ext.x >> ext.y >> ext.z;
      const PLANE<float> plane(pw[0], pw[1], pw[2], pw[3]);
      // GLGridLeaper.cpp:518-524
      FLOATMATRIX4 inv = (rot * tra).inverse();
      PLANE<float> transformed = plane * inv;
      const FLOATVECTOR3 normal(transformed.xyz().normalized());
      const float d = transformed.d();
      fprintf(out, "plane %a %a %a %a\n", (double)normal.x, (double)normal.y, (double)normal.z, (double)d);
      // the box as 12 triangles (what MaxMinBoxToVector emits: two per face; winding plays no role for BoxPlane)
      const FLOATVECTOR3 lo = FLOATVECTOR3(0, 0, 0) - ext / 2.0f, hi = FLOATVECTOR3(0, 0, 0) + ext / 2.0f;
      std::vector<FLOATVECTOR3> pos;
      for (int axis = 0; axis < 3; axis++)
        for (int side = 0; side < 2; side++) {
          FLOATVECTOR3 c[4];
          for (int k = 0; k < 4; k++) {
            float v[3];
            v[axis] = side ? hi[axis] : lo[axis];
            const int a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
            v[a1] = (k == 1 || k == 2) ? hi[a1] : lo[a1];
            v[a2] = (k >= 2) ? hi[a2] : lo[a2];
            c[k] = FLOATVECTOR3(v[0], v[1], v[2]);
          }
          pos.push_back(c[0]); pos.push_back(c[1]); pos.push_back(c[2]);
          pos.push_back(c[2]); pos.push_back(c[3]); pos.push_back(c[0]);
        }
      Clipper::BoxPlane(pos, normal, d);
      fprintf(out, "tris %zu", pos.size() / 3);
      for (size_t i = 0; i < pos.size(); i++) fprintf(out, " %a %a %a", (double)pos[i].x, (double)pos[i].y, (double)pos[i].z);
      fprintf(out, "\n");
    } else if (op == "inverse") {
      FLOATMATRIX4 a; read16(ls, a);
      put16(out, "inverse", a.inverse());
    } else if (op == "cull") {
      float fov, aspect, n, f; uint32_t py;
      ls >> fov >> aspect >> n >> f >> py;
      FLOATMATRIX4 model, view, proj; read16(ls, model); read16(ls, view); read16(ls, proj);
      CullingLOD c;
      c.SetScreenParams(fov, aspect, n, f, py);
      c.SetProjectionMatrix(proj);
      c.SetViewMatrix(view);
      c.SetModelMatrix(model);
      c.Update();
      size_t cnt; ls >> cnt;
      fprintf(out, "cull %zu", cnt);
      for (size_t i = 0; i < cnt; i++) {
        FLOATVECTOR3 ctr, ext; UINTVECTOR3 vox;
        ls >> ctr.x >> ctr.y >> ctr.z >> ext.x >> ext.y >> ext.z >> vox.x >> vox.y >> vox.z;
        fprintf(out, " %d %d", int(c.IsVisible(ctr, ext)), c.GetLODLevel(ctr, ext, vox));
      }
      fprintf(out, "\n");
    } else if (op == "uniforms") {
      // the statement sequence of GLGridLeaper::SetupRaycastShader / ComputeEyeToModelMatrix (GLGridLeaper.cpp:560-573,
      // 690-752) and AbstrRenderer::GetVolumeAABB (AbstrRenderer.cpp:1102-1109) on the reference's own vector classes
      FLOATMATRIX4 mv; read16(ls, mv);
      UINTVECTOR3 vDomainSize; FLOATVECTOR3 vScale, vLightDir, vEye;
      ls >> vDomainSize.x >> vDomainSize.y >> vDomainSize.z >> vScale.x >> vScale.y >> vScale.z;
      ls >> vLightDir.x >> vLightDir.y >> vLightDir.z >> vEye.x >> vEye.y >> vEye.z;
      FLOATVECTOR3 vExtend = FLOATVECTOR3(vDomainSize) * vScale;
      vExtend /= vExtend.maxVal();
      vScale /= vScale.minVal();
      FLOATVECTOR3 vCenter(0, 0, 0);
      FLOATMATRIX4 mTrans, mScale, mNormalize;
      mTrans.Translation(-vCenter);
      mScale.Scaling(1.0f / vExtend);
      mNormalize.Translation(0.5f, 0.5f, 0.5f);
      const FLOATMATRIX4 emm = mv.inverse() * mTrans * mScale * mNormalize;
      const FLOATVECTOR3 scale = 1.0f / vScale;
      const FLOATVECTOR3 l = (FLOATVECTOR4(vLightDir, 0.0f) * emm).xyz().normalized();
      const FLOATVECTOR3 e = (FLOATVECTOR4(vEye, 1.0f) * emm).xyz();
      put16(out, "emm", emm);
      put16(out, "m2e", emm.inverse());
      put16(out, "mvinv", mv.inverse());
      fprintf(out, "vecs %a %a %a %a %a %a %a %a %a %a %a %a\n", (double)scale.x, (double)scale.y, (double)scale.z, (double)l.x,
              (double)l.y, (double)l.z, (double)e.x, (double)e.y, (double)e.z, (double)vExtend.x, (double)vExtend.y, (double)vExtend.z);
    } else if (op == "tf1d") {
      size_t n; float center, inv;
      ls >> n >> center >> inv;
      TransferFunction1D tf(n);
      tf.SetStdFunction(center, inv);
      std::vector<unsigned char> bytes;
      tf.GetByteArray(bytes);
      fprintf(out, "tf1d %zu limits %llu %llu bytes", n, (unsigned long long)tf.GetNonZeroLimits().x,
              (unsigned long long)tf.GetNonZeroLimits().y);
      for (unsigned char b : bytes) fprintf(out, " %u", unsigned(b));
      fprintf(out, "\n");
    } else {
      fprintf(stderr, "ref_host: unknown command %s\n", op.c_str());
      return 2;
    }
  }
  fclose(out);
  return 0;
}
