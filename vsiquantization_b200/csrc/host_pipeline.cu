// host_pipeline.cu -- the end-to-end entry: forward + STE backward over HOST buffers.
//
// The range is cut into chunks; chunk i uses slot i % n_slots (its own stream and device staging
// buffers): H2D x,g -> fused fwd+bwd kernel (16 B/element on the device) -> D2H y,dx.  Stream order
// serialises the reuse of a slot, different slots overlap, so both PCIe directions and the kernel run
// concurrently.  This is the path bench.py's `e2e` key times (reference call it replaces:
// UniformQuantizer.quantize + autograd backward on host tensors, quantizers/uniform.py:34-56).
#include <new>

#include "common.cuh"

struct vsiq_host_pipeline {
    int64_t chunk;
    int n_slots;
    int device;
    float* dev[4];  // x, g, y, dx staging: n_slots * chunk floats each
    cudaStream_t* streams;
    int64_t last_launches;
};

extern "C" int vsiq_host_pipeline_create(vsiq_host_pipeline** out, int64_t chunk_elems, int n_slots) {
    if (!out || chunk_elems < 1024 || n_slots < 1 || n_slots > 16) return VSIQ_ERR_INVALID_ARG;
    vsiq_host_pipeline* p = new (std::nothrow) vsiq_host_pipeline();
    if (!p) return (int)cudaErrorMemoryAllocation;
    p->chunk = (chunk_elems + 7) / 8 * 8;  // keep every slot 32-byte aligned
    p->n_slots = n_slots;
    p->last_launches = 0;
    p->streams = nullptr;
    for (int k = 0; k < 4; ++k) p->dev[k] = nullptr;
    cudaError_t e = cudaGetDevice(&p->device);
    const size_t bytes = (size_t)p->chunk * (size_t)n_slots * sizeof(float);
    for (int k = 0; k < 4 && e == cudaSuccess; ++k) e = cudaMalloc((void**)&p->dev[k], bytes);
    if (e == cudaSuccess) {
        p->streams = new (std::nothrow) cudaStream_t[n_slots]();
        if (!p->streams) e = cudaErrorMemoryAllocation;
    }
    for (int s = 0; s < n_slots && e == cudaSuccess; ++s)
        e = cudaStreamCreateWithFlags(&p->streams[s], cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        vsiq_host_pipeline_destroy(p);
        return (int)e;
    }
    *out = p;
    return VSIQ_OK;
}

extern "C" int vsiq_host_pipeline_destroy(vsiq_host_pipeline* p) {
    if (!p) return VSIQ_OK;
    if (p->streams) {
        for (int s = 0; s < p->n_slots; ++s)
            if (p->streams[s]) {
                cudaStreamSynchronize(p->streams[s]);
                cudaStreamDestroy(p->streams[s]);
            }
        delete[] p->streams;
    }
    for (int k = 0; k < 4; ++k)
        if (p->dev[k]) cudaFree(p->dev[k]);
    delete p;
    return VSIQ_OK;
}

extern "C" int64_t vsiq_host_pipeline_last_launches(const vsiq_host_pipeline* p) { return p ? p->last_launches : 0; }

extern "C" int vsiq_host_pipeline_fwd_bwd(vsiq_host_pipeline* p, const float* x_host, const float* g_host,
                                          float* y_host, float* dx_host, int64_t n, float scale, float zero_point,
                                          int qmin, int qmax) {
    if (!p || !x_host || !g_host || !y_host || !dx_host || n < 0 || qmin >= qmax) return VSIQ_ERR_INVALID_ARG;
    vsiq_qparams qp = {};
    qp.scale = nullptr;
    qp.zero_point = nullptr;
    qp.scale_host = scale;
    qp.zp_host = zero_point;
    qp.qmin = qmin;
    qp.qmax = qmax;
    p->last_launches = 0;
    int rc = VSIQ_OK;
    int64_t i = 0;
    for (int64_t off = 0; off < n && rc == VSIQ_OK; off += p->chunk, ++i) {
        const int slot = (int)(i % p->n_slots);
        const int64_t len = (n - off) < p->chunk ? (n - off) : p->chunk;
        const size_t bytes = (size_t)len * sizeof(float);
        cudaStream_t st = p->streams[slot];
        float* dx_ = p->dev[0] + (size_t)slot * p->chunk;
        float* dg_ = p->dev[1] + (size_t)slot * p->chunk;
        float* dy_ = p->dev[2] + (size_t)slot * p->chunk;
        float* dd_ = p->dev[3] + (size_t)slot * p->chunk;
        cudaError_t e = cudaMemcpyAsync(dx_, x_host + off, bytes, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(dg_, g_host + off, bytes, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) { rc = (int)e; break; }
        vsiq_layout lay = {1, 1, len};
        rc = vsiq_fake_quant_fwd_bwd(dx_, dg_, dy_, dd_, &lay, &qp, (vsiq_stream_t)st);
        if (rc != VSIQ_OK) break;
        p->last_launches += 1;
        e = cudaMemcpyAsync(y_host + off, dy_, bytes, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(dx_host + off, dd_, bytes, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) rc = (int)e;
    }
    for (int s = 0; s < p->n_slots; ++s) {
        cudaError_t e = cudaStreamSynchronize(p->streams[s]);
        if (e != cudaSuccess && rc == VSIQ_OK) rc = (int)e;
    }
    return rc;
}
