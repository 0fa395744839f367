// host_pipeline.cu -- the end-to-end entry: forward + STE backward over HOST buffers.
//
// The range is cut into chunks; chunk i uses slot i % n_slots (device staging buffers for x, g, y, dx).  Three streams,
// one per engine -- H2D copies, the fused fwd+bwd kernel (16 B/element on the device), D2H copies -- linked by events:
//   H2D(i)    waits for kernel(i - n_slots)                 the slot's inputs have been consumed
//   kernel(i) waits for H2D(i) and D2H(i - n_slots)         inputs landed, the slot's outputs have been drained
//   D2H(i)    waits for kernel(i)
// so the inbound copy engine runs ahead of the outbound one by up to n_slots chunks instead of stalling on the D2H of
// its own slot (one stream per slot, the first form of this file, did: 47.9-52.7 ms per 2^28-element call depending on
// chunk size and slot count, this form 45.8-47.7 ms on the same box; profiles/r02_e2e_pipeline_sweep.log).  Both PCIe
// directions and the kernel run concurrently; the host spends 0.2-1.4 ms submitting a call.  This is the path bench.py's
// `e2e` key times (reference call it replaces: UniformQuantizer.quantize + autograd backward on host tensors,
// quantizers/uniform.py:34-56).
#include <chrono>
#include <new>

#include "common.cuh"

struct vsiq_host_pipeline {
    int64_t chunk;
    int n_slots;
    int device;
    float* dev[4];  // x, g, y, dx staging: n_slots * chunk floats each
    int64_t last_launches;
    int64_t last_enqueue_ns;  // host time spent submitting the last call's copies and launches (before the final wait)
    cudaStream_t s3[3];       // H2D, compute, D2H
    cudaEvent_t* ev;          // [3 * n_slots]: h2d done, kernel done, d2h done
};

extern "C" int vsiq_host_pipeline_create(vsiq_host_pipeline** out, int64_t chunk_elems, int n_slots) {
    if (!out || chunk_elems < 1024 || n_slots < 1 || n_slots > 16) return VSIQ_ERR_INVALID_ARG;
    vsiq_host_pipeline* p = new (std::nothrow) vsiq_host_pipeline();
    if (!p) return (int)cudaErrorMemoryAllocation;
    p->chunk = (chunk_elems + 7) / 8 * 8;  // keep every slot 32-byte aligned
    p->n_slots = n_slots;
    p->last_launches = 0;
    p->last_enqueue_ns = 0;
    p->ev = nullptr;
    for (int k = 0; k < 3; ++k) p->s3[k] = nullptr;
    for (int k = 0; k < 4; ++k) p->dev[k] = nullptr;
    cudaError_t e = cudaGetDevice(&p->device);
    const size_t bytes = (size_t)p->chunk * (size_t)n_slots * sizeof(float);
    for (int k = 0; k < 4 && e == cudaSuccess; ++k) e = cudaMalloc((void**)&p->dev[k], bytes);
    if (e == cudaSuccess) {
        p->ev = new (std::nothrow) cudaEvent_t[3 * n_slots]();
        if (!p->ev) e = cudaErrorMemoryAllocation;
        for (int k = 0; k < 3 && e == cudaSuccess; ++k) e = cudaStreamCreateWithFlags(&p->s3[k], cudaStreamNonBlocking);
        for (int k = 0; k < 3 * n_slots && e == cudaSuccess; ++k) e = cudaEventCreateWithFlags(&p->ev[k], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
        vsiq_host_pipeline_destroy(p);
        return (int)e;
    }
    *out = p;
    return VSIQ_OK;
}

extern "C" int vsiq_host_pipeline_destroy(vsiq_host_pipeline* p) {
    if (!p) return VSIQ_OK;
    for (int k = 0; k < 3; ++k)
        if (p->s3[k]) {
            cudaStreamSynchronize(p->s3[k]);
            cudaStreamDestroy(p->s3[k]);
        }
    if (p->ev) {
        for (int k = 0; k < 3 * p->n_slots; ++k)
            if (p->ev[k]) cudaEventDestroy(p->ev[k]);
        delete[] p->ev;
    }
    for (int k = 0; k < 4; ++k)
        if (p->dev[k]) cudaFree(p->dev[k]);
    delete p;
    return VSIQ_OK;
}

extern "C" int64_t vsiq_host_pipeline_last_launches(const vsiq_host_pipeline* p) { return p ? p->last_launches : 0; }
extern "C" int64_t vsiq_host_pipeline_last_enqueue_ns(const vsiq_host_pipeline* p) { return p ? p->last_enqueue_ns : 0; }

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
    const auto t_begin = std::chrono::steady_clock::now();
    for (int64_t off = 0; off < n && rc == VSIQ_OK; off += p->chunk, ++i) {
        const int slot = (int)(i % p->n_slots);
        const int64_t len = (n - off) < p->chunk ? (n - off) : p->chunk;
        const size_t bytes = (size_t)len * sizeof(float);
        float* dx_ = p->dev[0] + (size_t)slot * p->chunk;
        float* dg_ = p->dev[1] + (size_t)slot * p->chunk;
        float* dy_ = p->dev[2] + (size_t)slot * p->chunk;
        float* dd_ = p->dev[3] + (size_t)slot * p->chunk;
        vsiq_layout lay = {1, 1, len};
        cudaEvent_t e_in = p->ev[3 * slot], e_k = p->ev[3 * slot + 1], e_out = p->ev[3 * slot + 2];
        cudaError_t e = cudaSuccess;
        if (i >= p->n_slots) e = cudaStreamWaitEvent(p->s3[0], e_k, 0);  // the slot's inputs have been consumed
        if (e == cudaSuccess) e = cudaMemcpyAsync(dx_, x_host + off, bytes, cudaMemcpyHostToDevice, p->s3[0]);
        if (e == cudaSuccess) e = cudaMemcpyAsync(dg_, g_host + off, bytes, cudaMemcpyHostToDevice, p->s3[0]);
        if (e == cudaSuccess) e = cudaEventRecord(e_in, p->s3[0]);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(p->s3[1], e_in, 0);
        if (e == cudaSuccess && i >= p->n_slots) e = cudaStreamWaitEvent(p->s3[1], e_out, 0);  // outputs drained
        if (e != cudaSuccess) { rc = (int)e; break; }
        rc = vsiq_fake_quant_fwd_bwd(dx_, dg_, dy_, dd_, &lay, &qp, (vsiq_stream_t)p->s3[1]);
        if (rc != VSIQ_OK) break;
        p->last_launches += 1;
        e = cudaEventRecord(e_k, p->s3[1]);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(p->s3[2], e_k, 0);
        if (e == cudaSuccess) e = cudaMemcpyAsync(y_host + off, dy_, bytes, cudaMemcpyDeviceToHost, p->s3[2]);
        if (e == cudaSuccess) e = cudaMemcpyAsync(dx_host + off, dd_, bytes, cudaMemcpyDeviceToHost, p->s3[2]);
        if (e == cudaSuccess) e = cudaEventRecord(e_out, p->s3[2]);
        if (e != cudaSuccess) rc = (int)e;
    }
    p->last_enqueue_ns = (int64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(
                             std::chrono::steady_clock::now() - t_begin).count();
    for (int k = 0; k < 3; ++k) {
        cudaError_t e = cudaStreamSynchronize(p->s3[k]);
        if (e != cudaSuccess && rc == VSIQ_OK) rc = (int)e;
    }
    return rc;
}
