// placeholder until the tcgen05 kernel lands (same exported symbols)
#include "tta_common.cuh"
extern "C" {
int tta_conv_tc_supported(int mode, int K, int stride, int cin, int cout) { return 0; }
int tta_conv_tc_ntile(int cout) { return 16; }
long long tta_conv_tc_packed_bytes(int mode, int K, int stride, int cin, int cout) { return 0; }
int tta_conv_tc(const uint16_t*, const uint16_t*, long long, int, int, int, int, int, int, const void*,
                const float*, float*, long long, int, int, int, int, int, int, int, int, int, cudaStream_t) {
  tta_set_error("tta_conv_tc: not built");
  return TTA_ERR_UNSUPPORTED;
}
}
