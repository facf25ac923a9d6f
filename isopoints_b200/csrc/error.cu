// Error channel of the C ABI: every entry point returns an int status; the message of the
// last failure on the calling thread is kept here (the Python layer turns it into the
// RuntimeError the reference's TORCH_CHECK / AT_CUDA_CHECK would have raised).
#include "common.cuh"
#include <stdarg.h>
#include <atomic>

namespace isob200 {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace isob200

extern "C" {
long long isob200_launch_count(void) { return isob200::g_launches.load(); }
const char* isob200_last_error(void) { return isob200::g_err; }
int isob200_abi_version(void) { return 2; }   // 2: round 2 (fused splat epilogue, fps_ws, occ-backward workspace signature)
int isob200_compiled_arch(void) {
#ifdef ISOB200_ARCH
  return ISOB200_ARCH;
#else
  return 0;
#endif
}
}
