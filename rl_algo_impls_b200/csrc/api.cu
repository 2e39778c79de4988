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

// One env step's host -> device copies in one call: dst[i] <- src[i] (bytes[i] each), cudaMemcpyAsync on `stream`, in
// order.  The sources are page-locked (b200rl_host_register, or pinned allocations): plain DMAs, nothing staged.
// Replaces the per-field obs / mask uploads of sync_step_rollout.py:181-216 (there: torch.as_tensor(...).to(device)
// inside policy.step, actor_critic.py:306-318).
extern "C" int b200rl_h2d_batch(int n, void* const* dst_host, const void* const* src_host, const int64_t* bytes_host,
                                b200rl_stream_t stream) {
  using namespace b200rl;
  B200RL_REQUIRE(n >= 0 && (n == 0 || (dst_host && src_host && bytes_host)), "h2d_batch: null pointer");
  for (int i = 0; i < n; ++i) {
    B200RL_REQUIRE(bytes_host[i] >= 0 && (bytes_host[i] == 0 || (dst_host[i] && src_host[i])), "h2d_batch: copy %d is null", i);
    if (bytes_host[i] == 0) continue;
    cudaError_t e = cudaMemcpyAsync(dst_host[i], src_host[i], (size_t)bytes_host[i], cudaMemcpyHostToDevice, (cudaStream_t)stream);
    if (e != cudaSuccess) {
      set_error("h2d_batch: copy %d (%lld bytes): %s", i, (long long)bytes_host[i], cudaGetErrorString(e));
      (void)cudaGetLastError();
      return B200RL_ECUDA;
    }
  }
  return B200RL_OK;
}
