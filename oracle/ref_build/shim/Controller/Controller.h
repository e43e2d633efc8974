// TEST INFRASTRUCTURE ONLY (oracle/_ref build).
// Minimal stand-in for Tuvok/Controller/Controller.h so that the reference's
// IO / UVF / ExtendedOctree sources compile *unmodified* from /root/reference
// without dragging in MasterController (Lua, GL, Qt).  Include guards match the
// reference header so that later relative includes of the real file are no-ops.
#ifndef TUVOK_CONTROLLER_H
#define TUVOK_CONTROLLER_H
#ifndef TUVOK_STACK_TIMER_H
#define TUVOK_STACK_TIMER_H
#endif

#include <list>
#include <string>
#include "Basics/PerfCounter.h"
#include "DebugOut/AbstrDebugOut.h"
#include "DebugOut/ConsoleOut.h"

namespace tuvok {

class ShimSysInfo {
public:
  unsigned long long GetMaxUsableCPUMem() const { return 4ull << 30; }
  unsigned long long GetMaxUsableGPUMem() const { return 4ull << 30; }
  unsigned long long GetCPUMemSize() const { return 8ull << 30; }
  unsigned GetNumberOfCPUs() const { return 1; }
};

class ShimIOMan {
public:
  unsigned long long GetIncoresize() const { return 1ull << 20; }
  unsigned long long GetMaxBrickSize() const { return 256; }     // UVFDataset::Crop only
  unsigned long long GetBrickOverlap() const { return 2; }
};

class ShimMasterController {
public:
  ShimIOMan* IOMan() { return nullptr; }                    // AbstrConverter::GetIncoreSize -> DEFAULT_INCORESIZE
  const ShimIOMan& IOMan() const { static ShimIOMan m; return m; }
  void IncrementPerfCounter(enum PerfCounter, double) {}
  AbstrDebugOut* DebugOut() { static ConsoleOut out; static bool once = (out.SetOutput(true, false, false, false), true); (void)once; return &out; }
  const AbstrDebugOut* DebugOut() const { return const_cast<ShimMasterController*>(this)->DebugOut(); }
  const ShimSysInfo& SysInfo() const { static ShimSysInfo s; return s; }
};

namespace Controller {
  inline ShimMasterController& Instance() { static ShimMasterController m; return m; }
  inline const ShimMasterController& ConstInstance() { return Instance(); }
  inline const ShimMasterController& Const() { return Instance(); }
  namespace Debug {
    inline AbstrDebugOut& Out() { return *(Controller::Instance().DebugOut()); }
    inline const AbstrDebugOut& ConstOut() { return *(Controller::Instance().DebugOut()); }
  }
}

struct StackTimer { StackTimer(enum PerfCounter) {} };

} // namespace tuvok

#define TimedStatement(pc, block) do { block; } while (0)
#define T_ERROR(...)  do { tuvok::Controller::Debug::Out().Error(_func_, __VA_ARGS__); } while (0)
#define WARNING(...)  do { } while (0)
#define MESSAGE(...)  do { } while (0)
#define OTHER(...)    do { } while (0)

#endif
