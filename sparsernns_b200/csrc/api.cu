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
