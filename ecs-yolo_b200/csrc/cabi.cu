// Library-level pieces of the C ABI: error string, device query, version.
#include <stdarg.h>
#include <string.h>

#include "ecsy_common.cuh"
#include "../../include/ecsy.h"

static thread_local char g_err[512] = "";

void ecsy_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int ecsy_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      sms = n;
    else
      return 148;
  }
  return sms;
}

extern "C" const char* ecsy_last_error(void) { return g_err; }
extern "C" int ecsy_abi_version(void) { return ECSY_ABI_VERSION; }
extern "C" int ecsy_sm_count(void) { return ecsy_num_sms(); }
