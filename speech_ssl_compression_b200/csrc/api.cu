// C-ABI plumbing shared by all kernel files: error string, SM count, launch counter.
#include <stdarg.h>
#include <stdlib.h>
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

static const unsigned long long* g_drop_offset = nullptr;
const unsigned long long* dropout_offset_ptr() { return g_drop_offset; }

__global__ void counter_add_kernel(unsigned long long* c, unsigned long long v) { *c += v; }

bool pdl_enabled(int family) {  // MH_PDL: bit mask of kernel families (1 GEMM, 2 attention, 4 norm / column sums)
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("MH_PDL");
    on = e == nullptr ? 0 : atoi(e);  // default off: measured neutral for GEMM / attention (21.40 vs 21.34 ms per step), -0.56 ms for norm
  }
  return (on & family) != 0;
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

/* Device-side 64-bit counter mixed into every dropout seed.  A CUDA graph that replays the same
 * captured seeds still draws fresh masks each step if it bumps the counter (mh_counter_add). */
extern "C" int mh_set_dropout_offset_ptr(const unsigned long long* device_counter) {
  mh::g_drop_offset = device_counter;
  return 0;
}
extern "C" int mh_counter_add(unsigned long long* device_counter, unsigned long long v, void* stream) {
  mh::counter_add_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>(device_counter, v);
  MH_LAUNCH_CHECK();
  ++mh::g_launches;
  return 0;
}
