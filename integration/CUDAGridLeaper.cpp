// CUDAGridLeaper.cpp -- see CUDAGridLeaper.h.  Every method names the reference member it mirrors (file:line in the
// reference tree) and the ABI call that replaces the GL work.  Threading and ownership are the reference's: all calls
// come from the one thread that used to own the GL context; the tvk_ctx is not thread-safe; the brick callback is
// invoked synchronously from tvk_render on that thread and writes into pinned staging memory owned by the library.
// Errors follow the reference's convention -- bool + T_ERROR -- fed from the ABI's status code and tvk_last_error.
#include "CUDAGridLeaper.h"

#include <algorithm>
#include <cstring>
#include <stdexcept>

#include "Basics/SystemInfo.h"
#include "Controller/Controller.h"
#include "IO/LinearIndexDataset.h"
#include "IO/TransferFunction1D.h"
#include "IO/TransferFunction2D.h"
#include "Renderer/RenderRegion.h"

using namespace tuvok;

CUDAGridLeaper::CUDAGridLeaper(MasterController* pMasterController, bool bUseOnlyPowerOfTwo, bool bDownSampleTo8Bits,
                               bool bDisableBorder)
  : AbstrRenderer(pMasterController, bUseOnlyPowerOfTwo, bDownSampleTo8Bits, bDisableBorder),
    m_ctx(NULL), m_pToc(NULL), m_bConverged(false), m_bSortLast(false) {
  m_bSupportsMeshes = false;                                               // as GLGridLeaper.cpp:74
  std::memset(&m_params, 0, sizeof(m_params));
}

CUDAGridLeaper::~CUDAGridLeaper() { Cleanup(); }

void CUDAGridLeaper::Cleanup() {
  if (m_ctx) { tvk_destroy(m_ctx); m_ctx = NULL; }
}

bool CUDAGridLeaper::Fail(int rc) const {
  if (rc != TVK_OK) T_ERROR("%s", tvk_last_error(m_ctx));
  return rc != TVK_OK;
}

void CUDAGridLeaper::Log(void* user, int channel, const char* source, const char* msg) {   // Controller/Controller.h:68-84
  AbstrDebugOut* o = static_cast<CUDAGridLeaper*>(user)->m_pMasterController->DebugOut();
  if (channel == 2) o->Error(source, "%s", msg);
  else if (channel == 1) o->Warning(source, "%s", msg);
  else o->Message(source, "%s", msg);
}

// GLGridLeaper::RegisterDataset (GLGridLeaper.cpp:105-132): must be a LinearIndexDataset
bool CUDAGridLeaper::RegisterDataset(Dataset* ds) {
  if (!AbstrRenderer::RegisterDataset(ds)) return false;                   // AbstrRenderer.cpp:237-264
  m_pToc = dynamic_cast<LinearIndexDataset*>(ds);
  if (!m_pToc) { T_ERROR("Currently, this renderer works only with a LinearIndexDataset."); return false; }
  return true;
}

// GLGridLeaper::Initialize (GLGridLeaper.cpp:247-264): context -> tvk_ctx, dataset, transfer functions, pool
bool CUDAGridLeaper::Initialize(std::shared_ptr<Context>) {
  if (!m_pToc) return false;
  const RendererState& rs = m_pMasterController->RState;                   // MasterController.h:61-74
  tvk_device_cfg cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.device = 0;
  cfg.max_gpu_mem = m_pMasterController->SysInfo()->GetMaxUsableGPUMem();
  cfg.hash_table_size = rs.HashTableSize;
  cfg.rehash_count = rs.RehashCount;
  cfg.brick_strategy = int32_t(rs.BStrategy);
  if (tvk_create(&cfg, &m_ctx) != TVK_OK) { T_ERROR("%s", tvk_last_error(NULL)); return false; }
  tvk_set_log_callback(m_ctx, &CUDAGridLeaper::Log, this);

  const size_t ts = m_iTimestep;
  tvk_volume_desc d;
  std::memset(&d, 0, sizeof(d));
  const UINT64VECTOR3 dom = m_pToc->GetDomainSize(0, ts);                  // IO/Dataset.h:131
  const DOUBLEVECTOR3 sc = m_pToc->GetScale();                             // IO/Dataset.h:133
  const UINTVECTOR3 mb = m_pToc->GetMaxUsedBrickSizes();                   // IO/BrickedDataset.h:71
  for (int i = 0; i < 3; i++) {
    d.domain_size[i] = uint32_t(dom[i]); d.scale[i] = float(sc[i]); d.max_brick_size[i] = mb[i];
  }
  d.overlap = m_pToc->GetBrickOverlapSize()[0];                            // IO/Dataset.h:134
  d.dtype = m_pToc->GetIsFloat() ? TVK_F32 : m_pToc->GetBitWidth() == 8 ? TVK_U8 : TVK_U16;
  // AbstrRenderer::ColorData (AbstrRenderer.cpp:1417-1425): four 8-bit components -> the "-color" methods (k_color.cu);
  // MaxMinForKey already answers with the alpha component for such data (uvfDataset.cpp:1188)
  if (m_pToc->GetComponentCount() == 4 && m_pToc->GetBitWidth() == 8) d.dtype = TVK_RGBA8;
  else if (m_pToc->GetComponentCount() != 1) { T_ERROR("CUDAGridLeaper: %u-component data is not supported", unsigned(m_pToc->GetComponentCount())); return false; }
  d.range_max = MaxValue();                                                // AbstrRenderer.cpp:860-866
  d.max_gradient_magnitude = m_pToc->MaxGradientMagnitude();               // IO/Dataset.h:82
  // MaxMinForKey for all bricks of the pool LoDs in TOC order (what GLVolumePool.cpp:225-235 copies)
  m_minmax.clear();
  const size_t lods = size_t(m_pToc->GetLargestSingleBrickLOD(ts)) + 1;
  for (size_t l = 0; l < lods; l++) {
    const UINTVECTOR3 lay = m_pToc->GetBrickLayout(l, ts);                 // IO/LinearIndexDataset.h:21
    for (uint32_t z = 0; z < lay.z; z++)
      for (uint32_t y = 0; y < lay.y; y++)
        for (uint32_t x = 0; x < lay.x; x++) {
          const MinMaxBlock m = m_pToc->MaxMinForKey(m_pToc->IndexFrom4D(UINTVECTOR4(x, y, z, uint32_t(l)), ts));
          m_minmax.push_back(m.minScalar); m_minmax.push_back(m.maxScalar);
          m_minmax.push_back(m.minGradient); m_minmax.push_back(m.maxGradient);
        }
  }
  d.brick_count = m_minmax.size() / 4;
  d.minmax = m_minmax.data();
  if (Fail(tvk_set_volume(m_ctx, &d, &CUDAGridLeaper::FetchBrick, this))) return false;

  Push1DTrans();                                                           // GLRenderer.cpp:163-262 loads the TFs
  Push2DTrans();
  PushParams();
  PushClipPlane();
  return !Fail(tvk_create_pool(m_ctx, NULL));                              // GLGridLeaper::CreateVolumePool :83-103
}

// Dataset::GetBrick(key, vector<T>&) (IO/Dataset.h:93-100) -> library-owned pinned staging memory
int CUDAGridLeaper::FetchBrick(void* user, uint32_t x, uint32_t y, uint32_t z, uint32_t lod, void* dst, size_t cap) {
  CUDAGridLeaper* self = static_cast<CUDAGridLeaper*>(user);
  const BrickKey k = self->m_pToc->IndexFrom4D(UINTVECTOR4(x, y, z, lod), self->m_iTimestep);
  bool ok;
  size_t bytes;
  switch (self->m_pToc->GetBitWidth()) {
    case 8:
      ok = self->m_pToc->GetBrick(k, self->m_b8);  bytes = self->m_b8.size();
      std::memcpy(dst, self->m_b8.data(), std::min(bytes, cap)); break;
    case 16:
      ok = self->m_pToc->GetBrick(k, self->m_b16); bytes = self->m_b16.size() * 2;
      std::memcpy(dst, self->m_b16.data(), std::min(bytes, cap)); break;
    default:
      ok = self->m_pToc->GetBrick(k, self->m_b32); bytes = self->m_b32.size() * 4;
      std::memcpy(dst, self->m_b32.data(), std::min(bytes, cap)); break;
  }
  return ok && bytes <= cap ? 0 : 1;
}

// TransferFunction1D::GetByteArray / GetNonZeroLimits (TransferFunction1D.cpp:311-371)
void CUDAGridLeaper::Set1DTrans(const std::vector<unsigned char>& rgba) { m_p1DTrans->Set(rgba); Push1DTrans(); }
void CUDAGridLeaper::Changed1DTrans() { AbstrRenderer::Changed1DTrans(); Push1DTrans(); }    // GLGridLeaper.cpp:622-626
void CUDAGridLeaper::Changed2DTrans() { AbstrRenderer::Changed2DTrans(); Push2DTrans(); }    // GLGridLeaper.cpp:628-632

void CUDAGridLeaper::Push1DTrans() {
  if (!m_ctx || !m_p1DTrans) return;
  std::vector<unsigned char> b;
  m_p1DTrans->GetByteArray(b);
  const UINT64VECTOR2 nz = m_p1DTrans->GetNonZeroLimits();
  Fail(tvk_set_tf1d(m_ctx, b.data(), uint32_t(m_p1DTrans->GetSize()), nz.x, nz.y));
}

void CUDAGridLeaper::Push2DTrans() {
  if (!m_ctx || !m_p2DTrans) return;
  unsigned char* b = NULL;
  m_p2DTrans->GetByteArray(&b);                                            // TransferFunction2D.cpp:222-238 (Qt rasteriser)
  const UINT64VECTOR4 nz = m_p2DTrans->GetNonZeroLimits();                 // :378-397
  const uint64_t lim[4] = {nz.x, nz.y, nz.z, nz.w};
  Fail(tvk_set_tf2d(m_ctx, b, uint32_t(m_p2DTrans->GetSize().x), uint32_t(m_p2DTrans->GetSize().y), lim));
  delete[] b;
}

// view / mode state -> tvk_render_params (what SetupRaycastShader reads, GLGridLeaper.cpp:690-752)
void CUDAGridLeaper::PushParams() {
  if (!m_ctx) return;
  tvk_render_params& p = m_params;
  tvk_default_params(&p, m_vWinSize.x, m_vWinSize.y);
  const std::shared_ptr<RenderRegion3D> rr = GetFirst3DRegion();
  if (!rr) return;
  if (m_bUserMatrices) {                                                   // GLRenderer.cpp:893-898
    const FLOATMATRIX4 mv = rr->rotation * rr->translation * m_UserView;
    tvk_compute_view(&p, m_vWinSize.x, m_vWinSize.y, rr->rotation.array, rr->translation.array, &m_vEye.x, &m_vAt.x,
                     &m_vUp.x, m_fFOV, m_fZNear, m_fZFar, 1.0f);
    std::memcpy(p.model_view, mv.array, 64);
    std::memcpy(p.projection, m_UserProjection.array, 64);
  } else {
    tvk_compute_view(&p, m_vWinSize.x, m_vWinSize.y, rr->rotation.array, rr->translation.array,     // GLRenderer.cpp:627,892-919
                     &m_vEye.x, &m_vAt.x, &m_vUp.x, m_fFOV, m_fZNear, m_fZFar, 1.0f);
  }
  p.mode = int32_t(m_eRenderMode);
  p.lighting = m_bUseLighting;
  p.sample_rate_modifier = m_fSampleRateModifier;
  p.isovalue = GetIsoValue();
  p.nearest = m_eInterpolant == NearestNeighbor;
  std::memcpy(p.ambient, &m_cAmbient.x, 16);
  std::memcpy(p.diffuse, &m_cDiffuse.x, 16);
  std::memcpy(p.specular, &m_cSpecular.x, 16);
  std::memcpy(p.light_dir, &m_vLightDir.x, 12);
  std::memcpy(p.eye, &m_vEye.x, 12);
  std::memcpy(p.iso_color, &m_vIsoColor.x, 12);
  Fail(tvk_set_params(m_ctx, &p));
}

void CUDAGridLeaper::SetRendermode(ERenderMode m) { AbstrRenderer::SetRendermode(m); PushParams(); }
void CUDAGridLeaper::SetIsoValue(float v) { AbstrRenderer::SetIsoValue(v); PushParams(); }
void CUDAGridLeaper::SetSampleRateModifier(float f) { AbstrRenderer::SetSampleRateModifier(f); PushParams(); }
void CUDAGridLeaper::SetInterpolant(Interpolant i) { AbstrRenderer::SetInterpolant(i); PushParams(); }
void CUDAGridLeaper::Resize(const UINTVECTOR2& s) { AbstrRenderer::Resize(s); PushParams(); }
void CUDAGridLeaper::SetViewPort(UINTVECTOR2, UINTVECTOR2, bool) {}
void CUDAGridLeaper::UpdateLightParamsInShaders() { PushParams(); }

// GLGridLeaper::FillBBoxVBO (GLGridLeaper.cpp:506-532): the world-space plane goes back to model space with the
// reference's own PLANE / FLOATMATRIX4 arithmetic; the library cuts every ray's entry / exit with it
void CUDAGridLeaper::PushClipPlane() {
  if (!m_ctx) return;
  const std::shared_ptr<RenderRegion3D> rr = GetFirst3DRegion();
  if (!m_bClipPlaneOn || !rr) { Fail(tvk_set_clip_plane(m_ctx, 0, NULL)); return; }
  FLOATMATRIX4 inv = (rr->rotation * rr->translation).inverse();
  PLANE<float> transformed = m_ClipPlane.Plane() * inv;
  const FLOATVECTOR3 normal(transformed.xyz().normalized());
  const float plane[4] = {normal.x, normal.y, normal.z, transformed.d()};
  Fail(tvk_set_clip_plane(m_ctx, 1, plane));
}
void CUDAGridLeaper::SetClipPlane(RenderRegion* rr, const ExtendedPlane& plane) { AbstrRenderer::SetClipPlane(rr, plane); PushClipPlane(); }
void CUDAGridLeaper::EnableClipPlane(RenderRegion* rr) { AbstrRenderer::EnableClipPlane(rr); PushClipPlane(); }
void CUDAGridLeaper::DisableClipPlane(RenderRegion* rr) { AbstrRenderer::DisableClipPlane(rr); PushClipPlane(); }

// ClearView (AbstrRenderer.cpp:1247-1360): every setter forwards the whole state; classic isosurface frames use it
void CUDAGridLeaper::PushCV() {
  if (!m_ctx) return;
  const float col[3] = {m_vCVColor.x, m_vCVColor.y, m_vCVColor.z};
  const float pos[4] = {m_vCVPos.x, m_vCVPos.y, m_vCVPos.z, m_vCVPos.w};
  Fail(tvk_set_clearview(m_ctx, m_bDoClearView, GetCVIsoValue(), col, m_fCVSize, m_fCVContextScale, m_fCVBorderScale, pos));
}
void CUDAGridLeaper::SetCV(bool b) { AbstrRenderer::SetCV(b); PushCV(); }
void CUDAGridLeaper::SetCVIsoValue(float v) { AbstrRenderer::SetCVIsoValue(v); PushCV(); }
void CUDAGridLeaper::SetCVColor(const FLOATVECTOR3& c) { AbstrRenderer::SetCVColor(c); PushCV(); }
void CUDAGridLeaper::SetCVSize(float v) { AbstrRenderer::SetCVSize(v); PushCV(); }
void CUDAGridLeaper::SetCVContextScale(float v) { AbstrRenderer::SetCVContextScale(v); PushCV(); }
void CUDAGridLeaper::SetCVBorderScale(float v) { AbstrRenderer::SetCVBorderScale(v); PushCV(); }
void CUDAGridLeaper::SetCVFocusPosFVec(const FLOATVECTOR4& p) { AbstrRenderer::SetCVFocusPosFVec(p); PushCV(); }

// GLRenderer::Paint -> GLGridLeaper::Render3DRegion (GLRenderer.cpp:571-667, GLGridLeaper.cpp:914-1154)
bool CUDAGridLeaper::Paint() {
  if (!AbstrRenderer::Paint()) return false;
  const std::shared_ptr<RenderRegion3D> rr = GetFirst3DRegion();
  if (!rr) return false;
  if (rr->isBlank) { PushParams(); PushClipPlane(); }                      // new view: new ray-entry buffer (:925-945)
  if (m_bDoStereoRendering) {
    if (!PaintStereo()) return false;
    m_bConverged = true;
  } else {
    tvk_frame_stats st;
    if (Fail(tvk_render(m_ctx, &st))) return false;                        // one subframe; pages missing bricks in
    m_bConverged = st.converged != 0;                                      // GLGridLeaper.cpp:1097-1099
    m_pMasterController->IncrementPerfCounter(PERF_RAYCAST, st.ms_raycast);              // Basics/PerfCounter.h:7-44
    m_pMasterController->IncrementPerfCounter(PERF_READ_HTABLE, st.ms_read_htable);
    m_pMasterController->IncrementPerfCounter(PERF_UPLOAD_BRICKS, st.ms_upload_bricks);
    m_pMasterController->IncrementPerfCounter(PERF_RENDER, st.ms_total);
  }
  rr->isBlank = false;
  return true;
}

bool CUDAGridLeaper::CheckForRedraw() { return !m_bConverged || AbstrRenderer::CheckForRedraw(); }   // GLGridLeaper.cpp:872-890

// GLRenderer::Pick (GLRenderer.cpp:2856-2872), same exceptions
FLOATVECTOR3 CUDAGridLeaper::Pick(const UINTVECTOR2& mousePos) const {
  if (m_eRenderMode != RM_ISOSURFACE)
    throw std::runtime_error("Can only determine pick locations in isosurface rendering mode.");
  float v[3];
  if (tvk_pick(m_ctx, mousePos.x, mousePos.y, v) != TVK_OK) throw std::range_error("No intersection.");
  return FLOATVECTOR3(v[0], v[1], v[2]);
}

bool CUDAGridLeaper::CaptureRGBA8(std::vector<uint8_t>& out) {
  out.resize(size_t(m_vWinSize.area()) * 4);
  return !Fail(tvk_read_rgba8(m_ctx, out.data(), 0));
}

// stereo: GLRenderer::ComputeViewAndProjection (GLRenderer.cpp:892-919) + EndFrame (GLRenderer.cpp:758-812); the eyes are
// two converged frames through the same pool
bool CUDAGridLeaper::PaintStereo() {
  const std::shared_ptr<RenderRegion3D> rr = GetFirst3DRegion();
  tvk_render_params eye[2] = {m_params, m_params};
  if (m_bUserMatrices) {                                                   // GLRenderer.cpp:893-898
    const FLOATMATRIX4 rt = rr->rotation * rr->translation;
    std::memcpy(eye[0].model_view, (rt * m_UserViewLeft).array, 64);
    std::memcpy(eye[0].projection, m_UserProjectionLeft.array, 64);
    std::memcpy(eye[1].model_view, (rt * m_UserViewRight).array, 64);
    std::memcpy(eye[1].projection, m_UserProjectionRight.array, 64);
  } else if (Fail(tvk_compute_stereo_view(&eye[0], &eye[1], m_vWinSize.x, m_vWinSize.y, rr->rotation.array,
                                          rr->translation.array, &m_vEye.x, &m_vAt.x, &m_vUp.x, m_fFOV, m_fZNear,
                                          m_fZFar, 1.0f, m_fStereoFocalLength, m_fStereoEyeDist))) {
    return false;
  }
  for (int e = 0; e < 2; e++) {                                            // EStereoID SI_LEFT_OR_MONO, SI_RIGHT
    tvk_frame_stats st;
    if (Fail(tvk_set_params(m_ctx, &eye[e])) || Fail(tvk_paint(m_ctx, 0, &st)) || Fail(tvk_stereo_keep_eye(m_ctx, e)))
      return false;
  }
  return !Fail(tvk_stereo_compose(m_ctx, int(m_eStereoMode), m_bStereoEyeSwap, m_iAlternatingFrameID, 0.5f));
}

// GLRenderer::RenderHQMIPPreLoop (GLRenderer.cpp:1256-1285) + GLRaycaster::RenderHQMIPPreLoop (GLRaycaster.cpp:481-492)
bool CUDAGridLeaper::PaintHQMIP(const FLOATMATRIX4& regionRotation, int windowMode, bool flipX, bool flipY) {
  const double dPI = 3.141592653589793238462643383;
  FLOATMATRIX4 matRotDir, matFlipX, matFlipY;
  if (windowMode == 0) {                                                   // RenderRegion::WM_SAGITTAL
    FLOATMATRIX4 matTemp;
    matRotDir.RotationX(-dPI / 2.0); matTemp.RotationY(-dPI / 2.0);
    matRotDir = matRotDir * matTemp;
  } else if (windowMode == 1) {                                            // WM_AXIAL
    matRotDir.RotationX(-dPI / 2.0);
  }
  if (flipX) matFlipY.Scaling(-1, 1, 1);
  if (flipY) matFlipX.Scaling(1, -1, 1);
  m_maMIPRotation.RotationY(dPI * double(m_fMIPRotationAngle) / 180.0);
  m_maMIPRotation = matRotDir * regionRotation * matFlipX * matFlipY * m_maMIPRotation;
  tvk_render_params p = m_params;
  if (m_bOrthoView) {
    // GLRenderer.cpp:1183-1197: the parallel projection of the MIP frame; GLRaycaster.cpp:486-487: model view = the rotation
    FLOATMATRIX4 maOrtho;
    DOUBLEVECTOR2 vWinAspectRatio = 1.0 / DOUBLEVECTOR2(m_vWinSize);
    vWinAspectRatio = vWinAspectRatio / vWinAspectRatio.maxVal();
    const float fRoot2Scale = (vWinAspectRatio.x < vWinAspectRatio.y)
                                  ? std::max(1.0f, 1.414213f * float(vWinAspectRatio.x / vWinAspectRatio.y))
                                  : 1.414213f;
    // FLOATMATRIX4::Ortho (Vectors.h:1279-1284) is only declared under USEGL; this library-side renderer is GL-free, so the
    // six entries are written out
    const float l = -0.5f * fRoot2Scale / float(vWinAspectRatio.x), r = +0.5f * fRoot2Scale / float(vWinAspectRatio.x);
    const float b = -0.5f * fRoot2Scale / float(vWinAspectRatio.y), t = +0.5f * fRoot2Scale / float(vWinAspectRatio.y);
    const float zn = -100.0f, zf = 100.0f;
    maOrtho.array[0] = 2.0f / (r - l); maOrtho.array[12] = -(r + l) / (r - l);
    maOrtho.array[5] = 2.0f / (t - b); maOrtho.array[13] = -(t + b) / (t - b);
    maOrtho.array[10] = -2.0f / (zf - zn); maOrtho.array[14] = -(zf + zn) / (zf - zn);
    maOrtho.array[15] = 1.0f;
    std::memcpy(p.projection, maOrtho.array, 64);
    std::memcpy(p.model_view, m_maMIPRotation.array, 64);
  } else {
    std::memcpy(p.model_view, (m_maMIPRotation * m_mView[0]).array, 64);   // GLRaycaster.cpp:489 (perspective rays)
  }
  tvk_frame_stats st;
  return !Fail(tvk_set_params(m_ctx, &p)) && !Fail(tvk_render_mip(m_ctx, m_bMIPLOD, &st));
}

// sort-last across the GPUs of the box: partition, slice exchange (NCCL), n-way over and the RGBA8 gather all run inside
// the library on one stream (tvk_sortlast_frame); the host application only distributes the communicator id
bool CUDAGridLeaper::InitSortLast(const uint8_t commId[TVK_COMM_ID_BYTES], int rank, int nRanks, int policy) {
  m_bSortLast = !Fail(tvk_sortlast_init(m_ctx, commId, rank, nRanks, policy));
  return m_bSortLast;
}

bool CUDAGridLeaper::PaintSortLast(std::vector<uint8_t>* gatheredOnRank0) {
  if (!m_bSortLast) return false;
  const std::shared_ptr<RenderRegion3D> rr = GetFirst3DRegion();
  if (rr && rr->isBlank) { PushParams(); PushClipPlane(); rr->isBlank = false; }
  tvk_sortlast_stats st;
  if (Fail(tvk_sortlast_frame(m_ctx, &st))) return false;
  m_bConverged = st.frame.converged != 0;
  if (gatheredOnRank0) {
    gatheredOnRank0->resize(size_t(m_vWinSize.area()) * 4);
    return !Fail(tvk_sortlast_read_rgba8(m_ctx, gatheredOnRank0->data(), 0));
  }
  return true;
}

bool CUDAGridLeaper::IsVolumeResident(const BrickKey&) const { return false; }     // bricks live in the pool, not one by one
bool CUDAGridLeaper::CropDataset(const std::string&, bool) { return false; }       // as GLGridLeaper: not supported

// the `case CUDA_GRIDLEAPER:` of MasterController::RequestNewVolumeRenderer (MasterController.cpp:144-212); instantiating
// the class here also makes the compiler prove that no pure virtual of AbstrRenderer is left open
namespace tuvok {
AbstrRenderer* NewCUDAGridLeaper(MasterController* mc, bool bUseOnlyPowerOfTwo, bool bDownSampleTo8Bits, bool bDisableBorder) {
  return new CUDAGridLeaper(mc, bUseOnlyPowerOfTwo, bDownSampleTo8Bits, bDisableBorder);
}
}
