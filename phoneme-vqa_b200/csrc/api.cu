// api.cu — ABI version, error string, launch counter, device queries.
#include "common.cuh"
#include <mutex>

namespace pvqa {

static thread_local char tl_err[512] = "";
char* last_error_buf() { return tl_err; }

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tl_err, sizeof(tl_err), fmt, ap);
  va_end(ap);
  return code;
}

std::atomic<long long> g_launch_count{0};
const unsigned long long* g_rng_base = nullptr;
std::atomic<int> g_attn_bwd_waves{1};

int num_sms() {
  // cached per device id; the library is used with one device per process (one rank per GPU)
  static int cached_dev = -1, cached_sms = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return cached_sms;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) {
      cached_sms = n;
      cached_dev = dev;
    }
  }
  return cached_sms;
}

}  // namespace pvqa

extern "C" int pvqa_abi_version(void) { return PVQA_ABI_VERSION; }
extern "C" const char* pvqa_last_error(void) { return pvqa::last_error_buf(); }
extern "C" int64_t pvqa_launch_count(void) { return (int64_t)pvqa::g_launch_count.load(); }
extern "C" int pvqa_set_attn_bwd_waves(int k) {
  if (k < 1 || k > 16) return pvqa::fail(PVQA_ERR_SHAPE, "pvqa_set_attn_bwd_waves: %d is outside 1..16", k);
  pvqa::g_attn_bwd_waves.store(k);
  return PVQA_OK;
}
extern "C" int pvqa_set_rng_step_counter(const uint64_t* device_counter) {
  pvqa::g_rng_base = reinterpret_cast<const unsigned long long*>(device_counter);
  return PVQA_OK;
}
