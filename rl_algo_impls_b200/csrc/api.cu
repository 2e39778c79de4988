// libb200rl: version, error reporting, cached device attributes.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace b200rl {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return B200RL_ECUDA;
  }
  return B200RL_OK;
}

const DeviceInfo& device_info() {
  // read-only after first use; one entry per device ordinal
  static DeviceInfo info[64];
  static bool ready[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  if (!ready[dev]) {
    DeviceInfo d{148, 227 * 1024};
    cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&d.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    info[dev] = d;
    ready[dev] = true;
  }
  return info[dev];
}

}  // namespace b200rl

extern "C" int b200rl_version(void) { return B200RL_VERSION; }
extern "C" const char* b200rl_last_error(void) { return b200rl::g_error; }

// Page-lock / release a host buffer the caller already owns (a vec env's output array), so that H2D copies out of
// it are plain DMAs.  A refusal (range already registered, not registrable, ...) is reported as a code and the
// runtime's last-error state is cleared: the caller simply keeps using a staging buffer.
extern "C" int b200rl_host_register(void* ptr, size_t bytes) {
  using namespace b200rl;
  B200RL_REQUIRE(ptr != nullptr && bytes > 0, "host_register: null pointer or empty range");
  cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterDefault);
  if (e != cudaSuccess) {
    set_error("host_register: %s", cudaGetErrorString(e));
    (void)cudaGetLastError();
    return B200RL_ECUDA;
  }
  return B200RL_OK;
}

extern "C" int b200rl_host_unregister(void* ptr) {
  using namespace b200rl;
  B200RL_REQUIRE(ptr != nullptr, "host_unregister: null pointer");
  cudaError_t e = cudaHostUnregister(ptr);
  if (e != cudaSuccess) {
    set_error("host_unregister: %s", cudaGetErrorString(e));
    (void)cudaGetLastError();
    return B200RL_ECUDA;
  }
  return B200RL_OK;
}
