#include "common.cuh"

namespace dd {
thread_local char g_err[512] = "";
long long g_launches = 0;
int g_sm_budget = 0;
}  // namespace dd

extern "C" const char* dd_last_error(void) { return dd::g_err; }
extern "C" int dd_abi_version(void) { return 1; }
extern "C" long long dd_launch_count(void) { return dd::g_launches; }
extern "C" int dd_set_sm_budget(int sms) {
  const int prev = dd::sm_budget();
  dd::g_sm_budget = sms;
  return prev;
}
