// Library-level entry points of the C ABI (include/retisar_b200.h) and the error channel.
#include "rtb_common.cuh"

#include <mutex>
#include <set>
#include <utility>

namespace rtb {

bool first_use_on_device(const void *func) {
    static std::mutex mu;
    static std::set<std::pair<const void *, int>> seen;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    return seen.insert(std::make_pair(func, dev)).second;
}

char *last_error_buffer() {
    static thread_local char buf[512] = "";
    return buf;
}

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error_buffer(), 512, fmt, ap);
    va_end(ap);
    return code;
}

}  // namespace rtb

extern "C" {

int rtb_version(void) { return 100; }

const char *rtb_last_error(void) { return rtb::last_error_buffer(); }

int rtb_device_count(void) {
    int n = 0;
    cudaError_t err = cudaGetDeviceCount(&n);
    if (err != cudaSuccess) {
        rtb::fail(RTB_ERR_CUDA, "cudaGetDeviceCount failed: %s", cudaGetErrorString(err));
        return -RTB_ERR_CUDA;
    }
    return n;
}

// Page-locked host buffers for the host-buffer entry points (rtb_*_process_block_host): blocks that
// the caller fills on the CPU and the library copies to the device.  write_combined != 0 allocates
// write-combined memory (not cached by the CPU: fast to fill sequentially and to DMA, slow to read
// back on the host -- use it for input blocks only).
int rtb_host_alloc(void **ptr, unsigned long long bytes, int write_combined) {
    if (!ptr || bytes == 0) return rtb::fail(RTB_ERR_INVALID, "bad host allocation request");
    cudaError_t err = cudaHostAlloc(ptr, (size_t)bytes, write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault);
    if (err != cudaSuccess) {
        *ptr = nullptr;
        return rtb::fail(RTB_ERR_NOMEM, "cudaHostAlloc(%llu) failed: %s", bytes, cudaGetErrorString(err));
    }
    return RTB_OK;
}

int rtb_host_free(void *ptr) {
    if (ptr && cudaFreeHost(ptr) != cudaSuccess) return rtb::fail(RTB_ERR_CUDA, "cudaFreeHost failed");
    return RTB_OK;
}

}  // extern "C"
