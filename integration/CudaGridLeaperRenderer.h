// Reference-side adapter: the B200 ray-caster as a tuvok::AbstrRenderer.
//
// This file is compiled INSIDE a Tuvok build (it includes the reference's headers); it is not part of
// libtvk_cuda.so.  tests/test_integration_adapter.py compiles it with -fsyntax-only against the headers
// under /root/reference (where that tree is mounted) so that every override below is checked against the
// reference's real signatures:
//   pure virtuals   AbstrRenderer.h:189 Set1DTrans, :201 SetViewPort, :245 Pick, :267 NewFrameClear, :409 Cleanup,
//                   :480 CropDataset, :650 FixedFunctionality, :651 SyncStateManager, :679 ClearColorBuffer,
//                   :680 UpdateLightParamsInShaders, :851 IsVolumeResident
//   overridden      :192,:195 Changed1D/2DTrans, :294-296 PH_*, :322 GetRendererType, :323 SetRendermode,
//                   :324 SetBlendPrecision, :399 CaptureSingleFrame, :406 CheckForRedraw, :418 SetIsoValue,
//                   :431 Resize, :433 Paint, :460 RegisterDataset, :466 Initialize, :468 SetInterpolant
// Every method below carries `override`, so a signature that drifts from the reference's is a compile error.
// View state (SetViewPos / SetViewDir / SetUpDir / SetViewParameters, AbstrRenderer.h:250-261,505-513) is NOT virtual
// in the reference: the adapter reads m_vEye / m_vAt / m_vUp / m_fFOV / m_fZNear / m_fZFar whenever the 3-D region is
// blank, exactly where GLGridLeaper recomputes its ray entry (GLGridLeaper.cpp:914-1154, rr.isBlank).
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "Renderer/AbstrRenderer.h"
#include "Renderer/Context.h"
#include "tvk_host.h"

namespace tuvok {

class LinearIndexDataset;

/// tuvok.createContext for the CUDA renderer: the "share group" is the CUDA device ordinal
/// (BatchRenderer/BatchContext.h:47-66 is the GL counterpart).
class CudaContext : public Context {
 public:
  explicit CudaContext(int device) : Context(device) {}
};

class CudaGridLeaperRenderer : public AbstrRenderer {
 public:
  CudaGridLeaperRenderer(MasterController* pMasterController, bool bUseOnlyPowerOfTwo, bool bDownSampleTo8Bits,
                         bool bDisableBorder);
  virtual ~CudaGridLeaperRenderer();

  // --- life cycle
  bool Initialize(std::shared_ptr<Context> ctx) override;
  void Cleanup() override;
  bool RegisterDataset(Dataset* ds) override;
  void Resize(const UINTVECTOR2& vWinSize) override;
  void SetViewPort(UINTVECTOR2 lower_left, UINTVECTOR2 upper_right, bool decrease_screen_res) override;

  // --- state
  void Set1DTrans(const std::vector<unsigned char>& rgba) override;
  void Changed1DTrans() override;
  void Changed2DTrans() override;
  void SetRendermode(ERenderMode eRenderMode) override;
  void SetIsoValue(float fIsovalue) override;
  void SetBlendPrecision(EBlendPrecision eBlendPrecision) override;
  void SetInterpolant(Interpolant eInterpolant) override;
  ERendererType GetRendererType() const override { return RT_RC; }

  // --- frame loop
  bool Paint() override;
  bool CheckForRedraw() override;
  bool CaptureSingleFrame(const std::string& strFilename, bool bPreserveTransparency) const override;
  FLOATVECTOR3 Pick(const UINTVECTOR2&) const override;
  void NewFrameClear(const RenderRegion&) override;
  bool CropDataset(const std::string& strTempDir, bool bKeepOldData) override;
  void FixedFunctionality() const override;
  void SyncStateManager() override;

  // --- "paper hacks" the scripts use (BatchRenderer: repeat paint() until converged())
  void PH_ClearWorkingSet() override;
  UINTVECTOR4 PH_RecalculateVisibility() override;
  bool PH_Converged() const override;

 protected:
  void ClearColorBuffer() const override;
  void UpdateLightParamsInShaders() override;
  bool IsVolumeResident(const BrickKey& key) const override;

 private:
  /// walks the LinearIndexDataset once into a tvk::BrickedVolume (pinned host brick store) and hands it to the
  /// CUDA host: replaces GLGridLeaper::LoadDataset -> GPUMemMan::GetVolumePool + GLVolumePool::UploadFirstBrick.
  bool UploadDataset();
  /// copies view / light / clip / sampling state out of the AbstrRenderer members (called when the region is blank)
  void PushState();

  mutable tvk::CudaGridLeaper m_impl;
  LinearIndexDataset* m_pToCDataset;
  bool m_bUploaded;
  size_t m_iUploadedTimestep;   ///< the time step whose bricks the CUDA renderer holds
};

}  // namespace tuvok
