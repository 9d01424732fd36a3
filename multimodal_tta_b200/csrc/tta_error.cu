// Error plumbing of the C-ABI: every entry point returns an int status; the message of the last
// failure on the calling thread is read back with tta_last_error().
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "tta_common.cuh"

static thread_local char g_err[512] = "";

void tta_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool tta_pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* v = getenv("TTA_PDL");
    on = (v != nullptr && strcmp(v, "1") == 0) ? 1 : 0;  // opt-in: measured slower inside the CUDA graph
  }
  return on != 0;
}

bool tta_pdl_family(int family_bit) {
  static int off = -1;
  if (off < 0) {
    const char* v = getenv("TTA_PDL_OFF");
    off = v ? atoi(v) : 0;
  }
  return (off & family_bit) == 0;
}

int tta_check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    tta_set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
    return TTA_ERR_CUDA;
  }
  return TTA_OK;
}

extern "C" {
const char* tta_last_error(void) { return g_err; }

// ABI version of include/tta_b200.h this library was built from.
int tta_abi_version(void) { return 1; }

// Reports the compute capability of the current device (major*10+minor), or <0 on error.
int tta_device_sm(void) {
  int dev = 0, maj = 0, min = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) return -1;
  return maj * 10 + min;
}
}
