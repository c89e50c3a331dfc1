// Library-wide C-ABI plumbing: version + thread-local error string.
#include "common.cuh"

namespace dvis {
char *last_error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}
}  // namespace dvis

extern "C" int dvis_abi_version(void) { return DVIS_B200_ABI_VERSION; }
extern "C" const char *dvis_last_error(void) { return dvis::last_error_buffer(); }
