// See CudaGridLeaperRenderer.h.  Added to a Tuvok build next to Renderer/GL/GLGridLeaper.cpp; registered in
// MasterController::RequestNewVolumeRenderer (Controller/MasterController.cpp:144-215) as CUDA_GRIDLEAPER.
#include "CudaGridLeaperRenderer.h"

#include <cstring>
#include <stdexcept>

#include "Controller/Controller.h"
#include "IO/LinearIndexDataset.h"
#include "IO/TransferFunction1D.h"
#include "IO/TransferFunction2D.h"
#include "Renderer/RenderRegion.h"

namespace tuvok {

CudaGridLeaperRenderer::CudaGridLeaperRenderer(MasterController* pMasterController, bool bUseOnlyPowerOfTwo,
                                               bool bDownSampleTo8Bits, bool bDisableBorder)
    : AbstrRenderer(pMasterController, bUseOnlyPowerOfTwo, bDownSampleTo8Bits, bDisableBorder),
      m_pToCDataset(NULL), m_bUploaded(false), m_iUploadedTimestep(0) {}

CudaGridLeaperRenderer::~CudaGridLeaperRenderer() { m_impl.Cleanup(); }

// ---------------------------------------------------------------------------------------------- life cycle
bool CudaGridLeaperRenderer::Initialize(std::shared_ptr<Context> ctx) {
  if (!AbstrRenderer::Initialize(ctx)) return false;                      // needs a registered dataset
  if (!m_impl.Initialize(ctx ? ctx->GetShareGroupID() : 0)) {
    T_ERROR("CUDA renderer: %s", m_impl.LastError());
    return false;
  }
  return UploadDataset();
}

void CudaGridLeaperRenderer::Cleanup() { m_impl.Cleanup(); m_bUploaded = false; }

bool CudaGridLeaperRenderer::RegisterDataset(Dataset* ds) {              // GLGridLeaper.cpp:96-110
  LinearIndexDataset* lin = dynamic_cast<LinearIndexDataset*>(ds);
  if (!lin) {
    T_ERROR("the CUDA ray-caster needs a bricked LinearIndexDataset (a UVF with a TOC block)");
    return false;
  }
  if (!AbstrRenderer::RegisterDataset(ds)) return false;
  m_pToCDataset = lin;
  m_bUploaded = false;
  return true;
}

bool CudaGridLeaperRenderer::UploadDataset() {
  if (!m_pToCDataset) return false;
  LinearIndexDataset& ds = *m_pToCDataset;
  if (ds.GetComponentCount() != 1 || ds.GetIsSigned() || ds.GetIsFloat() || (ds.GetBitWidth() != 8 && ds.GetBitWidth() != 16)) {
    T_ERROR("the CUDA ray-caster renders 8/16-bit unsigned single-component volumes");
    return false;
  }
  const size_t ts = m_iTimestep;
  tvk::BrickedVolume* v = new tvk::BrickedVolume();
  const UINT64VECTOR3 dom = ds.GetDomainSize(0, ts);
  const DOUBLEVECTOR3 scale = ds.GetScale();
  const UINTVECTOR3 maxBrick = ds.GetMaxBrickSize(), overlap = ds.GetBrickOverlapSize();
  v->domain[0] = uint32_t(dom.x); v->domain[1] = uint32_t(dom.y); v->domain[2] = uint32_t(dom.z);
  v->scale[0] = float(scale.x); v->scale[1] = float(scale.y); v->scale[2] = float(scale.z);
  v->bpv = ds.GetBitWidth() / 8;
  v->brick_size = maxBrick.x;
  v->overlap = overlap.x;
  // per brick, ExtendedOctree ToC order = LOD 0 first, x fastest (LinearIndexDataset::IndexFrom4D): sizes + min/max
  const unsigned lods = ds.GetLODLevelCount();
  std::vector<BrickKey> keys;
  for (unsigned lod = 0; lod < lods; ++lod) {
    const UINTVECTOR3 layout = ds.GetBrickLayout(lod, ts);
    for (uint32_t z = 0; z < layout.z; ++z)
      for (uint32_t y = 0; y < layout.y; ++y)
        for (uint32_t x = 0; x < layout.x; ++x) {
          const BrickKey key = ds.IndexFrom4D(UINTVECTOR4(x, y, z, lod), ts);
          const UINTVECTOR3 n = ds.GetBrickVoxelCounts(key);
          const MinMaxBlock mm = ds.MaxMinForKey(key);
          keys.push_back(key);
          v->sizes.push_back(n.x); v->sizes.push_back(n.y); v->sizes.push_back(n.z);
          v->minmax.push_back(mm.minScalar); v->minmax.push_back(mm.maxScalar);
        }
  }
  std::shared_ptr<const Histogram1D> h1 = ds.Get1DHistogram();
  if (h1) {
    v->hist1d.resize(h1->GetFilledSize());
    for (size_t i = 0; i < v->hist1d.size(); ++i) v->hist1d[i] = h1->Get(i);
  }
  v->max_grad_magnitude = ds.MaxGradientMagnitude();
  if (!m_impl.PrepareHostVolume(v)) {
    T_ERROR("CUDA renderer: %s", m_impl.LastError());
    delete v;
    return false;
  }
  // brick payloads -> the pinned host store the pager streams from (replaces GLVolumePool::UploadBrick's
  // per-request Dataset::GetBrick + glTexSubImage3D, GLVolumePool.cpp:673-717,1362-1393)
  std::vector<uint8_t> b8;
  std::vector<uint16_t> b16;
  for (size_t i = 0; i < keys.size(); ++i) {
    const size_t voxels = size_t(v->sizes[3 * i]) * v->sizes[3 * i + 1] * v->sizes[3 * i + 2];
    bool ok;
    if (v->bpv == 1) { ok = ds.GetBrick(keys[i], b8) && b8.size() >= voxels; if (ok) memcpy(v->host_store + v->offsets[i], b8.data(), voxels); }
    else { ok = ds.GetBrick(keys[i], b16) && b16.size() >= voxels; if (ok) memcpy(v->host_store + v->offsets[i], b16.data(), voxels * 2); }
    if (!ok) { T_ERROR("could not read brick %u of the dataset", unsigned(i)); return false; }
  }
  if (!m_impl.AdoptVolume(v, 0)) {
    T_ERROR("CUDA renderer: %s", m_impl.LastError());
    return false;
  }
  m_bUploaded = true;
  m_iUploadedTimestep = ts;
  m_impl.Resize(m_vWinSize.x, m_vWinSize.y);
  Changed1DTrans();
  return true;
}

void CudaGridLeaperRenderer::Resize(const UINTVECTOR2& vWinSize) {
  AbstrRenderer::Resize(vWinSize);
  m_impl.Resize(vWinSize.x, vWinSize.y);
}

void CudaGridLeaperRenderer::SetViewPort(UINTVECTOR2 lower_left, UINTVECTOR2 upper_right, bool) {
  m_impl.Resize(upper_right.x - lower_left.x, upper_right.y - lower_left.y);
}

// ---------------------------------------------------------------------------------------------- state
void CudaGridLeaperRenderer::Set1DTrans(const std::vector<unsigned char>& rgba) {
  AbstrRenderer::Free1DTrans();
  m_impl.Set1DTrans(rgba);
  ScheduleCompleteRedraw();
}

void CudaGridLeaperRenderer::Changed1DTrans() {                           // GLGridLeaper.cpp:632-635
  if (m_p1DTrans && m_bUploaded) {
    std::vector<unsigned char> bytes;
    m_p1DTrans->GetByteArray(bytes);
    m_impl.Set1DTrans(bytes);
  }
  AbstrRenderer::Changed1DTrans();
}

void CudaGridLeaperRenderer::Changed2DTrans() {                           // GLGridLeaper.cpp:637-645
  if (m_p2DTrans && m_bUploaded) {
    unsigned char* bytes = NULL;                                           // allocated by GetByteArray (TransferFunction2D.cpp:222-238)
    m_p2DTrans->GetByteArray(&bytes);
    const UINT64VECTOR4& l = m_p2DTrans->GetNonZeroLimits();
    double nz[4] = {double(l.x), double(l.y), double(l.z), double(l.w)};
    const VECTOR2<size_t> size = m_p2DTrans->GetSize();
    m_impl.Set2DTrans(bytes, uint32_t(size.x), uint32_t(size.y), nz);
    delete[] bytes;
  }
  AbstrRenderer::Changed2DTrans();
}

void CudaGridLeaperRenderer::SetRendermode(ERenderMode eRenderMode) {
  AbstrRenderer::SetRendermode(eRenderMode);
  if (m_bUploaded && !m_impl.SetRendermode(int(eRenderMode))) T_ERROR("CUDA renderer: %s", m_impl.LastError());
}

void CudaGridLeaperRenderer::SetIsoValue(float fIsovalue) {
  AbstrRenderer::SetIsoValue(fIsovalue);
  if (m_bUploaded) m_impl.SetIsoValue(fIsovalue);
}

void CudaGridLeaperRenderer::SetBlendPrecision(EBlendPrecision eBlendPrecision) {
  AbstrRenderer::SetBlendPrecision(eBlendPrecision);                       // rays accumulate in fp32 registers: nothing to switch
}

void CudaGridLeaperRenderer::SetInterpolant(Interpolant eInterpolant) {
  AbstrRenderer::SetInterpolant(eInterpolant);                             // the pool is sampled GL_LINEAR like GLGridLeaper's
}

void CudaGridLeaperRenderer::UpdateLightParamsInShaders() {
  m_impl.SetLightColors(&m_cAmbient.x, &m_cDiffuse.x, &m_cSpecular.x, tvk::Vec3(m_vLightDir.x, m_vLightDir.y, m_vLightDir.z));
}

void CudaGridLeaperRenderer::PushState() {
  m_impl.SetViewParameters(m_fFOV, m_fZNear, m_fZFar, tvk::Vec3(m_vEye.x, m_vEye.y, m_vEye.z),
                           tvk::Vec3(m_vAt.x, m_vAt.y, m_vAt.z), tvk::Vec3(m_vUp.x, m_vUp.y, m_vUp.z));
  std::shared_ptr<RenderRegion3D> rr = GetFirst3DRegion();
  if (rr) {
    tvk::Matrix4 r, t;                                                     // FLOATMATRIX4 is 16 floats, m11 first
    memcpy(r.m, &rr->rotation.m11, 64);
    memcpy(t.m, &rr->translation.m11, 64);
    m_impl.SetRotation4x4(r);
    m_impl.SetTranslation4x4(t);
  }
  m_impl.SetUseLighting(m_bUseLighting);
  m_impl.SetSampleRateModifier(m_fSampleRateModifier);
  m_impl.SetIsosurfaceColor(tvk::Vec3(m_vIsoColor.x, m_vIsoColor.y, m_vIsoColor.z));
  m_impl.SetLightColors(&m_cAmbient.x, &m_cDiffuse.x, &m_cSpecular.x, tvk::Vec3(m_vLightDir.x, m_vLightDir.y, m_vLightDir.z));
  const PLANE<float>& pl = m_ClipPlane.Plane();
  const float plane[4] = {pl.x, pl.y, pl.z, pl.w};
  m_impl.SetClipPlane(plane);
  m_impl.EnableClipPlane(m_bClipPlaneOn);
  m_impl.SetRendererTarget(int(m_eRendererTarget));
  // stereo: AbstrRenderer's setters only record the state (AbstrRenderer.cpp:1387-1410); m_bDoStereoRendering is the
  // request restricted to a single 3-D region (AbstrRenderer.h:444)
  m_impl.SetStereo(m_bDoStereoRendering);
  m_impl.SetStereoEyeDist(m_fStereoEyeDist);
  m_impl.SetStereoFocalLength(m_fStereoFocalLength);
  m_impl.SetStereoMode(int(m_eStereoMode));
  m_impl.SetStereoEyeSwap(m_bStereoEyeSwap);
  // ClearView (classic renderer type only): state AbstrRenderer's setters recorded (AbstrRenderer.cpp:1247-1256,1304-1361)
  m_impl.SetCV(m_bDoClearView);
  m_impl.SetCVIsoValue(GetCVIsoValue());
  m_impl.SetCVColor(tvk::Vec3(m_vCVColor.x, m_vCVColor.y, m_vCVColor.z));
  m_impl.SetCVSize(m_fCVSize);
  m_impl.SetCVContextScale(m_fCVContextScale);
  m_impl.SetCVBorderScale(m_fCVBorderScale);
  m_impl.SetCVFocusPosFVec(&m_vCVPos.x);
}

// ---------------------------------------------------------------------------------------------- frame loop
bool CudaGridLeaperRenderer::Paint() {                                     // GLRenderer::Paint + GLGridLeaper::Render3DRegion
  if (!AbstrRenderer::Paint()) return false;
  if (m_bDatasetIsInvalid || !m_bUploaded) return true;
  // AbstrRenderer::SetTimestep (AbstrRenderer.h:573) is not virtual: it only records m_iTimestep and schedules a redraw.
  // The brick store here holds ONE time step, so a change is picked up where the redraw arrives.
  if (m_iTimestep != m_iUploadedTimestep && !UploadDataset()) return false;
  // A window that shows ONE 2-D region (slice view / MIP): GLRenderer::Render2DView's job.  The region covers the whole
  // target here; ImageVis3D's 2x2 layout of several regions per window is not mapped.
  if (renderRegions.size() == 1 && renderRegions[0]->is2D()) {
    RenderRegion* r2 = renderRegions[0].get();
    if (r2->isBlank) {
      PushState();
      bool fx = false, fy = false;
      Get2DFlipMode(r2, fx, fy);
      const bool mip = GetUseMIP(r2);
      const bool ok = (mip && m_eRendererTarget == RT_CAPTURE)
                          ? m_impl.RenderHQMIP(int(r2->windowMode), fx, fy, m_fMIPRotationAngle, m_bOrthoView, m_bMIPLOD)
                          : m_impl.Render2DView(int(r2->windowMode), GetSliceDepth(r2), mip, fx, fy);
      if (!ok) { T_ERROR("CUDA renderer: %s", m_impl.LastError()); return false; }
      r2->isBlank = false; r2->isTargetBlank = false; r2->redrawMask = false;
      ++m_iFrameCounter;
    }
    return true;
  }
  std::shared_ptr<RenderRegion3D> rr = GetFirst3DRegion();
  if (rr && rr->isBlank) {                                                 // a new view: everything may have changed
    PushState();
    m_impl.ScheduleCompleteRedraw();
  }
  if (!m_impl.Paint()) {
    T_ERROR("CUDA renderer: %s", m_impl.LastError());
    return false;
  }
  if (rr) { rr->isBlank = false; rr->isTargetBlank = false; rr->redrawMask = false; }
  ++m_iFrameCounter;
  return true;
}

bool CudaGridLeaperRenderer::CheckForRedraw() {                            // GLGridLeaper.cpp:872-890
  if (m_vWinSize.area() == 0) return false;
  std::shared_ptr<RenderRegion3D> rr = GetFirst3DRegion();
  return (rr && rr->isBlank) || m_impl.CheckForRedraw();
}

bool CudaGridLeaperRenderer::CaptureSingleFrame(const std::string& strFilename, bool bPreserveTransparency) const {
  return m_impl.CaptureSingleFrame(strFilename, bPreserveTransparency);
}

FLOATVECTOR3 CudaGridLeaperRenderer::Pick(const UINTVECTOR2& mousePos) const {      // GLRenderer.cpp:2856-2872
  float p[3];
  if (!const_cast<tvk::CudaGridLeaper&>(m_impl).Pick(int(mousePos.x), int(mousePos.y), p))
    throw std::runtime_error(m_impl.LastError());          // "Can only determine pick locations in isosurface ..." / "No intersection."
  return FLOATVECTOR3(p[0], p[1], p[2]);
}

void CudaGridLeaperRenderer::NewFrameClear(const RenderRegion&) {}         // the march kernel clears what it rewrites
bool CudaGridLeaperRenderer::CropDataset(const std::string&, bool) { return false; }
void CudaGridLeaperRenderer::FixedFunctionality() const {}                 // no GL state
void CudaGridLeaperRenderer::SyncStateManager() {}
void CudaGridLeaperRenderer::ClearColorBuffer() const {}

void CudaGridLeaperRenderer::PH_ClearWorkingSet() { m_impl.ClearWorkingSet(); }
UINTVECTOR4 CudaGridLeaperRenderer::PH_RecalculateVisibility() {
  m_impl.RecalcVisibility();
  return UINTVECTOR4(0, 0, 0, 0);
}
bool CudaGridLeaperRenderer::PH_Converged() const { return m_impl.Converged(); }

bool CudaGridLeaperRenderer::IsVolumeResident(const BrickKey& key) const {
  if (!m_pToCDataset || !m_impl.Pool()) return false;
  const UINTVECTOR4 b = m_pToCDataset->IndexTo4D(key);
  return m_impl.Pool()->IsBrickResident(b.x, b.y, b.z, b.w);
}

}  // namespace tuvok
