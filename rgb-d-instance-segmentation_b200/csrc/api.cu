// C-ABI plumbing: version and thread-local error text (include/rgbd_b200.h).
#include <stdarg.h>
#include "common.cuh"
#include "rgbd_b200.h"

static thread_local char g_err[512] = "";

void rgbd_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int rgbd_abi_version(void) { return RGBD_ABI_VERSION; }
extern "C" const char* rgbd_last_error(void) { return g_err; }

// Per-device launch state (common.cuh): looked up for the CURRENT device on every launch, cached per device.
int rgbd_device_info(RgbdDeviceInfo* info) {
    static std::mutex m;
    static RgbdDeviceInfo cache[RGBD_MAX_DEVICES] = {};
    int dev = 0;
    RGBD_CHECK_CUDA(cudaGetDevice(&dev));
    RGBD_CHECK_ARG(dev >= 0 && dev < RGBD_MAX_DEVICES, "device ordinal %d out of range", dev);
    std::lock_guard<std::mutex> g(m);
    if (!cache[dev].num_sms) {
        int cc_major = 0;
        RGBD_CHECK_CUDA(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, dev));
        if (cc_major != 10) {
            rgbd_set_error("device %d has compute capability %d.x; librgbd_b200 is built for sm_100a only", dev, cc_major);
            return RGBD_ERR_UNSUPPORTED;
        }
        cache[dev].device = dev;
        RGBD_CHECK_CUDA(cudaDeviceGetAttribute(&cache[dev].max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        RGBD_CHECK_CUDA(cudaDeviceGetAttribute(&cache[dev].num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    *info = cache[dev];
    return RGBD_OK;
}
