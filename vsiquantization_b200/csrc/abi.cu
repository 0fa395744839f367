// abi.cu -- library-level entry points and the host-side helpers shared by the launchers.
#include <stdlib.h>

#include <mutex>

#include "common.cuh"

namespace vsiq {

static std::mutex g_props_mutex;
static DeviceProps g_props[64];
static bool g_props_valid[64];

int get_device_props(DeviceProps* out) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (dev < 0 || dev >= 64) return (int)cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> lk(g_props_mutex);
    if (!g_props_valid[dev]) {
        DeviceProps p;
        if ((e = cudaDeviceGetAttribute(&p.sm_count, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return (int)e;
        if ((e = cudaDeviceGetAttribute(&p.cc_major, cudaDevAttrComputeCapabilityMajor, dev)) != cudaSuccess) return (int)e;
        if ((e = cudaDeviceGetAttribute(&p.cc_minor, cudaDevAttrComputeCapabilityMinor, dev)) != cudaSuccess) return (int)e;
        g_props[dev] = p;
        g_props_valid[dev] = true;
    }
    *out = g_props[dev];
    return 0;
}

// Grid policy.  Default: a persistent grid of (SM count x resident CTAs per SM) CTAs walking the tiles
// with a grid stride.  VSIQ_GRID_WAVES=k (k >= 1) allows k x that many CTAs; VSIQ_GRID_WAVES=0 launches
// one CTA per tile (tuning knobs, read once).
static int grid_waves() {
    static int waves = -1;
    if (waves < 0) {
        const char* s = getenv("VSIQ_GRID_WAVES");
        waves = s ? atoi(s) : 1;
        if (waves < 0) waves = 1;
    }
    return waves;
}

int grid_for(uint32_t n_ctas_wanted, int ctas_per_sm) {
    DeviceProps dp;
    if (int e = get_device_props(&dp)) return -e;
    if (n_ctas_wanted < 1) n_ctas_wanted = 1;
    const int waves = grid_waves();
    if (waves == 0) return (int)(n_ctas_wanted > 0x7fffffffu ? 0x7fffffffu : n_ctas_wanted);
    const uint64_t cap = (uint64_t)dp.sm_count * (uint64_t)ctas_per_sm * (uint64_t)waves;
    return (int)(n_ctas_wanted < cap ? n_ctas_wanted : cap);
}

bool aligned32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31u) == 0; }

int check_layout(const vsiq_layout* l) {
    if (!l) return VSIQ_ERR_INVALID_ARG;
    if (l->outer < 0 || l->channels < 1 || l->inner < 0) return VSIQ_ERR_INVALID_ARG;
    if (l->outer > 0 && l->inner > 0) {
        // total element count must fit comfortably in int64 and the tile count in 31 bits
        const long double n = (long double)l->outer * (long double)l->channels * (long double)l->inner;
        if (n >= 9.0e18L) return VSIQ_ERR_INVALID_ARG;
    }
    return VSIQ_OK;
}

int fill_qp(const vsiq_qparams* in, QPDev* out) {
    if (!in) return VSIQ_ERR_INVALID_ARG;
    if (in->qmin >= in->qmax) return VSIQ_ERR_INVALID_ARG;
    if ((in->scale_dtype != VSIQ_F32 && in->scale_dtype != VSIQ_F64) ||
        (in->zp_dtype != VSIQ_F32 && in->zp_dtype != VSIQ_F64))
        return VSIQ_ERR_INVALID_ARG;
    out->scale = in->scale;
    out->zp = in->zero_point;
    out->scale_f64 = in->scale_dtype == VSIQ_F64;
    out->zp_f64 = in->zp_dtype == VSIQ_F64;
    out->scale_host = in->scale_host;
    out->zp_host = in->zp_host;
    out->zp_learned = in->zp_learned ? 1 : 0;
    out->lo = (float)in->qmin;
    out->hi = (float)in->qmax;
    return VSIQ_OK;
}

}  // namespace vsiq

extern "C" int vsiq_version(void) { return VSIQ_VERSION; }

extern "C" const char* vsiq_error_string(int code) {
    switch (code) {
        case VSIQ_OK: return "ok";
        case VSIQ_ERR_INVALID_ARG: return "vsiq: invalid argument";
        case VSIQ_ERR_WORKSPACE: return "vsiq: workspace missing or too small";
        case VSIQ_ERR_UNSUPPORTED: return "vsiq: unsupported combination";
        case VSIQ_ERR_NO_DEVICE: return "vsiq: no usable CUDA device";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "vsiq: unknown error";
}

extern "C" int vsiq_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    vsiq::DeviceProps dp;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n < 1) {
        (void)cudaGetLastError();
        return VSIQ_ERR_NO_DEVICE;
    }
    if (int e = vsiq::get_device_props(&dp)) return e;
    if (sm_count) *sm_count = dp.sm_count;
    if (cc_major) *cc_major = dp.cc_major;
    if (cc_minor) *cc_minor = dp.cc_minor;
    return VSIQ_OK;
}
