// core.cu -- error string, version, launch counter of libb2rl.so.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "common.cuh"

namespace b2rl {
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static int g_pdl = -1;          // -1: read B2RL_PDL on first use (default on)
bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("B2RL_PDL");
    g_pdl = (e && strcmp(e, "0") == 0) ? 0 : 1;
  }
  return g_pdl != 0;
}
void set_pdl(int on) { g_pdl = on ? 1 : 0; }
}  // namespace b2rl

extern "C" int b2rl_version(void) { return 100; }
extern "C" const char* b2rl_last_error(void) { return b2rl::g_err; }
extern "C" int64_t b2rl_launch_count(void) { return b2rl::g_launches.load(); }
extern "C" void b2rl_reset_launch_count(void) { b2rl::g_launches.store(0); }
extern "C" void b2rl_set_pdl(int32_t on) { b2rl::set_pdl(on); }
