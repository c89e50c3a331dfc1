// Library-wide C-ABI plumbing: version + thread-local error string.
#include "common.cuh"

namespace dvis {
char *last_error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}
}  // namespace dvis

namespace dvis {
bool g_pdl = false;   // programmatic dependent launch for the temporal-stage kernels (dvis_set_pdl)
}
extern "C" int dvis_set_pdl(int enabled) {
  dvis::g_pdl = enabled != 0;
  return DVIS_OK;
}
extern "C" int dvis_abi_version(void) { return DVIS_B200_ABI_VERSION; }
extern "C" const char *dvis_last_error(void) { return dvis::last_error_buffer(); }

// ---- driver entry point for tensor-map encoding (resolved once; no link-time dependency on libcuda) ----
#include "tc05.cuh"
namespace dvis {
PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<PFN_encodeTiled>(p);
  }();
  return fn;
}
}  // namespace dvis
