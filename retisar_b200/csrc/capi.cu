// Library-level entry points of the C ABI (include/retisar_b200.h) and the error channel.
#include "rtb_common.cuh"

namespace rtb {

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

}  // extern "C"
