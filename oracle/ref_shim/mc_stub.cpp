// mc_stub -- minimal stand-in for the parts of tuvok::MasterController that the
// reference IO / pool sources reach through Controller::Instance() (perf
// counters and debug output only).  Lets oracle/_ref link the reference's data
// and pool code without Lua, GL or the IO manager.  Test infrastructure only.
#include "Controller/MasterController.h"
#include "IO/IOManager.h"
#include "IO/AbstrConverter.h"
#include "Basics/Mesh.h"
#include "LuaScripting/LuaScripting.h"
#include "LuaScripting/LuaMemberReg.h"
#include "LuaScripting/TuvokSpecific/LuaIOManagerProxy.h"

namespace tuvok {

MasterController::MasterController()
    : m_pSystemInfo(NULL), m_pGPUMemMan(NULL), m_pIOManager(NULL),
      m_bDeleteDebugOutOnExit(false), m_bExperimentalFeatures(false),
      m_pActiveRenderer(NULL) {
  for (size_t i = 0; i < PERF_END; i++) m_Perf[i] = 0.0;
  RState.BStrategy = RendererState::BS_SkipTwoLevels;
}

MasterController::~MasterController() {}

double MasterController::PerfQuery(enum PerfCounter pc) {
  double v = m_Perf[pc];
  m_Perf[pc] = 0.0;
  return v;
}

void MasterController::IncrementPerfCounter(enum PerfCounter pc, double amount) {
  m_Perf[pc] += amount;
}

AbstrDebugOut* MasterController::DebugOut() { return &m_DefaultOut; }
const AbstrDebugOut* MasterController::DebugOut() const { return &m_DefaultOut; }

}  // namespace tuvok
