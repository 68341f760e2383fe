// Library-wide state: last-error string, launch counter, version.
#include "common.cuh"

namespace effdet {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
}  // namespace effdet

extern "C" const char *effdet_last_error(void) { return effdet::g_err; }
extern "C" int effdet_version(void) { return 100; }
extern "C" long long effdet_launch_count(void) { return effdet::g_launches.load(); }

extern "C" int effdet_zero(void *ptr, size_t bytes, void *stream) {
    if (bytes == 0) return EFFDET_OK;
    EFFDET_REQUIRE(ptr, "null pointer");
    EFFDET_CUDA(cudaMemsetAsync(ptr, 0, bytes, effdet::as_stream(stream)));
    return EFFDET_OK;
}
