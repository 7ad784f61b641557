// C-ABI plumbing shared by all kernel files: error string, SM count, launch counter.
#include <stdarg.h>
#include <string.h>

#include "mh_b200.h"
#include "mh_common.cuh"

namespace mh {
static thread_local char g_err[512] = "";
long long g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}
}  // namespace mh

extern "C" const char* mh_last_error(void) { return mh::g_err; }
extern "C" int mh_version(void) { return 100; }
extern "C" long long mh_launch_count(void) { return mh::g_launches; }
