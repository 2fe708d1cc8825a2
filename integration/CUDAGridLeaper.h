// CUDAGridLeaper -- the reference-side half of the drop-in boundary: a tuvok::AbstrRenderer subclass
// that forwards the renderer interface to the C ABI of include/tvk.h (libtvkcuda.so), the way GLGridLeaper forwards it
// to OpenGL.  It lives in the reference tree as Renderer/CUDA/CUDAGridLeaper.{h,cpp}; it is kept here so that it can be
// compiled against the reference's own headers (tests/test_integration_shim.py runs g++ on it with -I/root/reference:
// every pure virtual of Renderer/AbstrRenderer.h:112-881 is overridden, every member it touches exists with that type).
// Tuvok itself cannot be LINKED in this image (Qt, GL, bison absent), so the shim is syntax- and type-checked only.
#pragma once
#ifndef TUVOK_CUDAGRIDLEAPER_H
#define TUVOK_CUDAGRIDLEAPER_H

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "Renderer/AbstrRenderer.h"
#include "tvk.h"

namespace tuvok {

class LinearIndexDataset;

class CUDAGridLeaper : public AbstrRenderer {
public:
  CUDAGridLeaper(MasterController* pMasterController, bool bUseOnlyPowerOfTwo, bool bDownSampleTo8Bits,
                 bool bDisableBorder);
  virtual ~CUDAGridLeaper();

  virtual ERendererType GetRendererType() const { return RT_RC; }                  // AbstrRenderer.h:322
  virtual bool SupportsClearView() { return true; }                               // classic isosurface frames only

  // GLGridLeaper::RegisterDataset (GLGridLeaper.cpp:105-132) / Initialize (:247-264) / Cleanup (:134-160)
  virtual bool RegisterDataset(Dataset* ds);
  virtual bool Initialize(std::shared_ptr<Context> ctx);
  virtual void Cleanup();

  // transfer functions (AbstrRenderer.h:189-195; GLGridLeaper.cpp:622-632)
  virtual void Set1DTrans(const std::vector<unsigned char>& rgba);
  virtual void Changed1DTrans();
  virtual void Changed2DTrans();

  // state that SetupRaycastShader reads (GLGridLeaper.cpp:690-752)
  virtual void SetRendermode(ERenderMode eRenderMode);                             // GLGridLeaper.cpp:639-645
  virtual void SetIsoValue(float fIsovalue);                                       // :634-637
  virtual void SetSampleRateModifier(float fSampleRateModifier);
  virtual void SetInterpolant(Interpolant eInterpolant);
  virtual void Resize(const UINTVECTOR2& vWinSize);                                // AbstrRenderer.cpp:474-478
  virtual void SetViewPort(UINTVECTOR2 lower_left, UINTVECTOR2 upper_right, bool decrease_screen_res);
  virtual void UpdateLightParamsInShaders();

  // clip plane (AbstrRenderer.h:215-219; GLGridLeaper.cpp:506-532,1337-1356)
  virtual void SetClipPlane(RenderRegion* renderRegion, const ExtendedPlane& plane);
  virtual void EnableClipPlane(RenderRegion* renderRegion = NULL);
  virtual void DisableClipPlane(RenderRegion* renderRegion = NULL);

  // ClearView (AbstrRenderer.cpp:1247-1360)
  virtual void SetCV(bool bEnable);
  virtual void SetCVIsoValue(float fIsovalue);
  virtual void SetCVColor(const FLOATVECTOR3& vColor);
  virtual void SetCVSize(float fSize);
  virtual void SetCVContextScale(float fScale);
  virtual void SetCVBorderScale(float fScale);
  virtual void SetCVFocusPosFVec(const FLOATVECTOR4& vPos);

  // GLRenderer::Paint -> GLGridLeaper::Render3DRegion (GLRenderer.cpp:571-667, GLGridLeaper.cpp:914-1154)
  virtual bool Paint();
  virtual bool CheckForRedraw();                                                   // GLGridLeaper.cpp:872-890
  virtual FLOATVECTOR3 Pick(const UINTVECTOR2& mousePos) const;                    // GLRenderer.cpp:2856-2872

  // GLFrameCapture read-back of GetLastFBO() (GLFrameCapture.cpp:72-85)
  bool CaptureRGBA8(std::vector<uint8_t>& out);
  // 2D windows in MIP mode: GLRenderer::Render2DView's HQ branch (GLRenderer.cpp:1183-1253)
  bool PaintHQMIP(const FLOATMATRIX4& regionRotation, int windowMode, bool flipX, bool flipY);
  // sort-last over the GPUs of one box: every process owns one CUDAGridLeaper; `commId` is tvk_sortlast_unique_id's result
  // broadcast by the host application (MPI, a socket, a file -- the ABI does not care)
  bool InitSortLast(const uint8_t commId[TVK_COMM_ID_BYTES], int rank, int nRanks, int policy = TVK_SL_OCTANT);
  bool PaintSortLast(std::vector<uint8_t>* gatheredOnRank0);

  // nothing to do without GL state (AbstrRenderer.h:267,480,650,651,679,851)
  virtual void NewFrameClear(const RenderRegion&) {}
  virtual void FixedFunctionality() const {}
  virtual void SyncStateManager() {}
  virtual void ClearColorBuffer() const {}
  virtual bool IsVolumeResident(const BrickKey& key) const;
  virtual bool CropDataset(const std::string& strTempDir, bool bKeepOldData);

private:
  static int  FetchBrick(void* user, uint32_t x, uint32_t y, uint32_t z, uint32_t lod, void* dst, size_t cap);
  static void Log(void* user, int channel, const char* source, const char* msg);
  bool Fail(int rc) const;
  void Push1DTrans();
  void Push2DTrans();
  void PushParams();
  void PushClipPlane();
  void PushCV();
  bool PaintStereo();

  tvk_ctx*            m_ctx;
  LinearIndexDataset* m_pToc;
  bool                m_bConverged;
  bool                m_bSortLast;
  tvk_render_params   m_params;
  std::vector<double>   m_minmax;
  std::vector<uint8_t>  m_b8;
  std::vector<uint16_t> m_b16;
  std::vector<float>    m_b32;
};

}  // namespace tuvok
#endif
