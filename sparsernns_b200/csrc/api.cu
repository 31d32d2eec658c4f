// api.cu - error plumbing, version and launch accounting of libs5b200.
#include <vector>

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

// SM count of the current device.  If the query fails the error is recorded (s5b200_last_error) and 1 is returned: every
// grid sized from it stays valid - in particular a cooperative launch can never be over-subscribed by a guess - and the
// caller's next CUDA call reports the underlying failure.
int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { set_error("sm_count: cudaGetDevice: %s", cudaGetErrorString(e)); return 1; }
  if (dev != cached_dev) {
    int n = 0;
    e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess || n < 1) { set_error("sm_count: cudaDeviceGetAttribute: %s", cudaGetErrorString(e)); return 1; }
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

struct ProfRec { cudaEvent_t a, b; int cls; };
static thread_local bool g_prof_on = false;
static thread_local std::vector<ProfRec>* g_prof = nullptr;

ProfScope::ProfScope(int kernel_class, cudaStream_t stream) : idx(-1), st(stream) {
  if (!g_prof_on) return;
  ProfRec r;
  r.cls = kernel_class;
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
  cudaEventRecord(r.a, st);
  g_prof->push_back(r);
  idx = (int)g_prof->size() - 1;
}
ProfScope::~ProfScope() {
  if (idx >= 0) cudaEventRecord((*g_prof)[idx].b, st);
}

}  // namespace s5

extern "C" int s5b200_profile_begin(void) {
  using namespace s5;
  if (!g_prof) g_prof = new std::vector<ProfRec>();
  for (auto& r : *g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  g_prof->clear();
  g_prof_on = true;
  return S5_OK;
}
extern "C" int s5b200_profile_end(double* ms_per_class, long long* launches_per_class) {
  using namespace s5;
  g_prof_on = false;
  if (!ms_per_class || !launches_per_class) S5_FAIL(S5_E_BADARG, "profile_end: null pointer");
  for (int i = 0; i < S5_K_COUNT; ++i) { ms_per_class[i] = 0.0; launches_per_class[i] = 0; }
  if (!g_prof) return S5_OK;
  for (auto& r : *g_prof) {
    float ms = 0.f;
    S5_CUDA(cudaEventSynchronize(r.b));
    S5_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
    if (r.cls >= 0 && r.cls < S5_K_COUNT) { ms_per_class[r.cls] += ms; launches_per_class[r.cls]++; }
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  g_prof->clear();
  return S5_OK;
}

extern "C" int s5b200_version(void) { return S5B200_VERSION; }
extern "C" const char* s5b200_last_error(void) { return s5::g_err; }
extern "C" long long s5b200_launch_count(void) { return s5::g_launches; }
extern "C" void s5b200_reset_launch_count(void) { s5::g_launches = 0; }
