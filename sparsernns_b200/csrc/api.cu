// api.cu - error plumbing, version and launch accounting of libs5b200.
#include "common.cuh"

namespace s5 {

static thread_local char g_err[512] = "";
static thread_local long long g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

long long& launch_counter() { return g_launches; }

int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 148;
    cached = prop.multiProcessorCount;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace s5

extern "C" int s5b200_version(void) { return S5B200_VERSION; }
extern "C" const char* s5b200_last_error(void) { return s5::g_err; }
extern "C" long long s5b200_launch_count(void) { return s5::g_launches; }
extern "C" void s5b200_reset_launch_count(void) { s5::g_launches = 0; }
