// fake_quant.cu -- kernels (3) and (4): uniform fake-quant forward, STE backward, LSQ backward with
// the per-channel step-size / zero-point gradients in the same pass, and the fused fwd+bwd sweep.
//
// Reference semantics reproduced here (bit-exact for y, codes and dx; fp64-accumulated sums):
//   forward   quantizers/uniform.py:54-55, :95          per channel: quantizers/lsq_module.py:254-274
//   backward  autograd of the above through RoundStraightThrough (uniform.py:258-271), torch.clamp and
//             ScaleGradient (uniform.py:242-255); grad-scale uniform.py:69-71 / lsq_module.py:327-340
//   FunLSQ    quantizers/uniform.py:144-150 (mask_mode 1)
//
// Roofline: HBM.  Algorithmic bytes per element: fwd 8 (read x, write y), bwd 12 (read g, read x,
// write dx; the mask is recomputed from x, per-channel outputs are O(C)), fused fwd+bwd 16.
#include "common.cuh"

namespace vsiq {

// ------------------------------------------------------------------------------------------ ops
struct FwdOp : OpBase {
    QP p;
    __device__ __forceinline__ void apply(const float (&a)[1], float (&o)[1]) {
        float r = fq_round(a[0], p);
        o[0] = fq_dequant(clamp_torch(r, p.lo, p.hi), p);
    }
};

struct SteBwdOp : OpBase {
    QP p;
    __device__ __forceinline__ void apply(const float (&a)[2], float (&o)[1]) {
        float r = fq_round(a[0], p);
        o[0] = ste_dx(a[1], fq_inrange(r, p), p);
    }
};

struct FwdBwdOp : OpBase {
    QP p;
    __device__ __forceinline__ void apply(const float (&a)[2], float (&o)[2]) {
        float r = fq_round(a[0], p);
        o[0] = fq_dequant(clamp_torch(r, p.lo, p.hi), p);
        o[1] = ste_dx(a[1], fq_inrange(r, p), p);
    }
};

template <int MASK_MODE, bool WANT_DZ>
struct LsqBwdOp : OpBase {
    QP p;
    float e_acc;  // sum g * ((q - z) - m * x/s)   over this thread's elements of the current tile
    float b_acc;  // sum g over clamped-out elements
    __device__ __forceinline__ void apply(const float (&a)[2], float (&o)[1]) {
        const float x = a[0], g = a[1];
        const float v = div_exact(x, p.s, p.r, p.fast);
        if (MASK_MODE == VSIQ_MASK_ROUNDED) {
            const float r = rintf(__fadd_rn(v, p.z));
            const bool m = fq_inrange(r, p);
            const float d = __fsub_rn(clamp_torch(r, p.lo, p.hi), p.z);
            o[0] = ste_dx(g, m, p);
            // reference: g*(q-z) from mul-backward minus where(m, g*s, 0) * ((x/s)/s) from div-backward.
            // v * 0 keeps the reference's NaN for infinite inputs.
            const float mv = __fmul_rn(v, m ? 1.0f : 0.0f);
            e_acc = fmaf(g, d - mv, e_acc);
            if (WANT_DZ) b_acc += m ? 0.0f : g;
        } else {
            const float small = v < p.lo ? 1.0f : 0.0f;
            const float big = v > p.hi ? 1.0f : 0.0f;
            const float mid = 1.0f - small - big;
            const float term = small * p.lo + big * p.hi + mid * (rintf(v) - v);
            e_acc = fmaf(term, g, e_acc);
            o[0] = __fmul_rn(mid, g);
        }
    }
};

// -------------------------------------------------------------------------------------- kernels
template <int GROUP, int V, class Op, int NIN, int NOUT>
__device__ __forceinline__ void run_elementwise(const float* const (&in)[NIN], float* const (&out)[NOUT],
                                                const Tiles& tiles, const QPDev& qpd) {
    Op op;
    int64_t cur_channel = -1;
    for (uint32_t t = group_index<GROUP>(); t < tiles.n_tiles; t += group_count<GROUP>()) {
        const TileCursor<GROUP> c = tile_at<GROUP>(tiles, t);
        if (c.channel != cur_channel) {
            op.p = load_qp(qpd, c.channel);
            cur_channel = c.channel;
        }
        span_apply<GROUP, V, NIN, NOUT>(in, out, c.offset, c.len, op);
    }
}

template <int GROUP, int V>
__global__ void __launch_bounds__(kThreads) fq_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                          Tiles tiles, QPDev qpd) {
    const float* const in[1] = {x};
    float* const out[1] = {y};
    run_elementwise<GROUP, V, FwdOp, 1, 1>(in, out, tiles, qpd);
}

template <int GROUP, int V>
__global__ void __launch_bounds__(kThreads) fq_bwd_ste_kernel(const float* __restrict__ x,
                                                              const float* __restrict__ g,
                                                              float* __restrict__ dx, Tiles tiles, QPDev qpd) {
    const float* const in[2] = {x, g};
    float* const out[1] = {dx};
    run_elementwise<GROUP, V, SteBwdOp, 2, 1>(in, out, tiles, qpd);
}

template <int GROUP, int V>
__global__ void __launch_bounds__(kThreads) fq_fwd_bwd_kernel(const float* __restrict__ x,
                                                              const float* __restrict__ g, float* __restrict__ y,
                                                              float* __restrict__ dx, Tiles tiles, QPDev qpd) {
    const float* const in[2] = {x, g};
    float* const out[2] = {y, dx};
    run_elementwise<GROUP, V, FwdBwdOp, 2, 2>(in, out, tiles, qpd);
}

// integer-code export (deployment path, not the training hot loop): 4-byte loads, 1-byte stores
__global__ void __launch_bounds__(kThreads) fq_codes_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                            int8_t* __restrict__ codes, int64_t n, int64_t inner,
                                                            int64_t channels, QPDev qpd) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t c = channels == 1 ? 0 : (i / inner) % channels;
        const QP p = load_qp(qpd, c);
        const float q = clamp_torch(fq_round(ld_stream1(x + i), p), p.lo, p.hi);
        if (y) y[i] = fq_dequant(q, p);
        const int qi = (q == q) ? (int)q : 0;  // NaN has no code; emit 0
        codes[i] = (int8_t)(qi & 0xff);
    }
}

// LSQ backward.  Partials: single-row tensors (per tensor) accumulate in registers across all tiles a
// group owns and emit ONE partial per group; multi-row tensors (per channel) emit one partial per
// tile.  The last CTA combines them per channel in a fixed order in fp64 and applies the grad-scale.
template <int GROUP, int V, int MASK_MODE, bool WANT_DZ>
__global__ void __launch_bounds__(kThreads)
    lsq_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g, float* __restrict__ dx, Tiles tiles,
                   QPDev qpd, void* ws, void* dscale, int ds_f64, void* dzp, int dz_f64, double gs_host,
                   const float* __restrict__ gs_dev, int64_t outer) {
    __shared__ double s_red[kWarps][2];
    const float* const in[2] = {x, g};
    float* const out[1] = {dx};
    double* partials = ws_partials(ws);
    const bool single_row = tiles.rows == 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    LsqBwdOp<MASK_MODE, WANT_DZ> op;
    double e_run = 0.0, b_run = 0.0;
    int64_t cur_channel = -1;
    for (uint32_t t = group_index<GROUP>(); t < tiles.n_tiles; t += group_count<GROUP>()) {
        const TileCursor<GROUP> c = tile_at<GROUP>(tiles, t);
        if (c.channel != cur_channel) {
            op.p = load_qp(qpd, c.channel);
            cur_channel = c.channel;
        }
        op.e_acc = 0.0f;
        op.b_acc = 0.0f;
        span_apply<GROUP, V, 2, 1>(in, out, c.offset, c.len, op);
        if (single_row) {
            e_run += (double)op.e_acc;
            if (WANT_DZ) b_run += (double)op.b_acc;
        } else {
            // per-tile flush: fp32 inside the warp, fp64 across warps
            float e = warp_sum(op.e_acc);
            float b = WANT_DZ ? warp_sum(op.b_acc) : 0.0f;
            if (GROUP == 32) {
                if (lane == 0) {
                    partials[2 * (size_t)t] = (double)e;
                    partials[2 * (size_t)t + 1] = (double)b;
                }
            } else {
                __syncthreads();  // s_red free again
                if (lane == 0) {
                    s_red[warp][0] = (double)e;
                    s_red[warp][1] = (double)b;
                }
                __syncthreads();
                if (threadIdx.x == 0) {
                    double es = 0.0, bs = 0.0;
#pragma unroll
                    for (int w = 0; w < kWarps; ++w) {
                        es += s_red[w][0];
                        bs += s_red[w][1];
                    }
                    partials[2 * (size_t)t] = es;
                    partials[2 * (size_t)t + 1] = bs;
                }
            }
        }
    }
    if (single_row) {
        double e = warp_sum(e_run);
        double b = WANT_DZ ? warp_sum(b_run) : 0.0;
        if (GROUP == 32) {
            const size_t slot = group_index<32>();
            if (lane == 0 && slot < tiles.n_tiles) {  // idle warps own no slot
                partials[2 * slot] = e;
                partials[2 * slot + 1] = b;
            }
        } else {
            if (lane == 0) {
                s_red[warp][0] = e;
                s_red[warp][1] = b;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                double es = 0.0, bs = 0.0;
#pragma unroll
                for (int w = 0; w < kWarps; ++w) {
                    es += s_red[w][0];
                    bs += s_red[w][1];
                }
                partials[2 * (size_t)blockIdx.x] = es;
                partials[2 * (size_t)blockIdx.x + 1] = bs;
            }
        }
    }

    if (!last_cta_ticket((unsigned int*)ws)) return;

    // ---- finalize (one CTA): fixed-order fp64 combination, grad-scale, dtype conversion ----
    const double gs = gs_host * (gs_dev ? (double)__ldg(gs_dev) : 1.0);
    const int64_t C = tiles.channels;
    for (int64_t c = warp; c < C; c += kWarps) {
        double e = 0.0, b = 0.0;
        if (single_row) {
            const uint32_t n_slots = group_count<GROUP>() < tiles.n_tiles ? group_count<GROUP>() : tiles.n_tiles;
            for (uint32_t i = lane; i < n_slots; i += 32) {
                e += __ldcg(partials + 2 * (size_t)i);
                b += __ldcg(partials + 2 * (size_t)i + 1);
            }
        } else {
            const int64_t items = outer * (int64_t)tiles.chunks;
            for (int64_t i = lane; i < items; i += 32) {
                const int64_t o = i / tiles.chunks, k = i - o * tiles.chunks;
                const size_t slot = (size_t)((o * C + c) * tiles.chunks + k);
                e += __ldcg(partials + 2 * slot);
                b += __ldcg(partials + 2 * slot + 1);
            }
        }
        e = warp_sum(e);
        b = warp_sum(b);
        if (lane == 0) {
            const QP p = load_qp(qpd, c);
            const double ds = gs * e;
            if (ds_f64)
                ((double*)dscale)[c] = ds;
            else
                ((float*)dscale)[c] = (float)ds;
            if (WANT_DZ) {
                const float zr = rintf(p.zf);
                const bool cz = qpd.zp_learned ? ((zr >= p.lo) && (zr <= p.hi)) : true;
                const double dz = cz ? -gs * (double)p.s * b : 0.0;
                if (dz_f64)
                    ((double*)dzp)[c] = dz;
                else
                    ((float*)dzp)[c] = (float)dz;
            }
        }
    }
}

// ------------------------------------------------------------------------------------ launchers
template <class K>
static int occupancy_grid(K kernel, uint32_t n_ctas_wanted) {
    DeviceProps dp;
    if (int e = get_device_props(&dp)) return -e;
    int per_sm = 0;
    cudaError_t ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0);
    if (ce != cudaSuccess || per_sm < 1) per_sm = 1;
    return grid_for(n_ctas_wanted, per_sm);
}

template <int GROUP>
static uint32_t ctas_for_tiles(uint32_t n_tiles) {
    return GROUP == kThreads ? n_tiles : (n_tiles + kWarps - 1) / kWarps;
}

#define VSIQ_DISPATCH_GROUP_VEC(warp_group, vec8, CALL) \
    do {                                                \
        if (warp_group) {                               \
            if (vec8) {                                 \
                CALL(32, 8);                            \
            } else {                                    \
                CALL(32, 1);                            \
            }                                           \
        } else {                                        \
            if (vec8) {                                 \
                CALL(kThreads, 8);                      \
            } else {                                    \
                CALL(kThreads, 1);                      \
            }                                           \
        }                                               \
    } while (0)

}  // namespace vsiq

using namespace vsiq;

extern "C" int vsiq_fake_quant_fwd(const float* x, float* y, void* codes, const vsiq_layout* layout,
                                   const vsiq_qparams* qp, vsiq_stream_t stream) {
    if (!x || (!y && !codes)) return VSIQ_ERR_INVALID_ARG;
    if (int e = check_layout(layout)) return e;
    QPDev qpd;
    if (int e = fill_qp(qp, &qpd)) return e;
    const int64_t n = layout->outer * layout->channels * layout->inner;
    if (n == 0) return VSIQ_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (codes) {
        if (qp->qmin < -128 || qp->qmax > 255 || (qp->qmin < 0 && qp->qmax > 127)) return VSIQ_ERR_UNSUPPORTED;
        int64_t blocks = (n + kThreads - 1) / kThreads;
        int grid = grid_for((uint32_t)(blocks > (1 << 30) ? (1 << 30) : blocks), 8);
        if (grid < 0) return -grid;
        fq_codes_kernel<<<grid, kThreads, 0, st>>>(x, y, (int8_t*)codes, n, layout->inner, layout->channels, qpd);
        return (int)cudaGetLastError();
    }
    const bool warp_group = layout->inner < kWarpGroupMaxInner;
    const bool vec8 = aligned32(x) && aligned32(y);
    Tiles tiles;
#define CALL(G, V)                                                                          \
    {                                                                                       \
        if (!make_tiles<G>(layout->outer, layout->channels, layout->inner, &tiles))         \
            return VSIQ_ERR_INVALID_ARG;                                                    \
        int grid = occupancy_grid(fq_fwd_kernel<G, V>, ctas_for_tiles<G>(tiles.n_tiles));   \
        if (grid < 0) return -grid;                                                         \
        fq_fwd_kernel<G, V><<<grid, kThreads, 0, st>>>(x, y, tiles, qpd);                   \
    }
    VSIQ_DISPATCH_GROUP_VEC(warp_group, vec8, CALL);
#undef CALL
    return (int)cudaGetLastError();
}

extern "C" int vsiq_fake_quant_bwd_ste(const float* x, const float* g, float* dx, const vsiq_layout* layout,
                                       const vsiq_qparams* qp, vsiq_stream_t stream) {
    if (!x || !g || !dx) return VSIQ_ERR_INVALID_ARG;
    if (int e = check_layout(layout)) return e;
    QPDev qpd;
    if (int e = fill_qp(qp, &qpd)) return e;
    if (layout->outer * layout->channels * layout->inner == 0) return VSIQ_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool warp_group = layout->inner < kWarpGroupMaxInner;
    const bool vec8 = aligned32(x) && aligned32(g) && aligned32(dx);
    Tiles tiles;
#define CALL(G, V)                                                                            \
    {                                                                                         \
        if (!make_tiles<G>(layout->outer, layout->channels, layout->inner, &tiles))           \
            return VSIQ_ERR_INVALID_ARG;                                                      \
        int grid = occupancy_grid(fq_bwd_ste_kernel<G, V>, ctas_for_tiles<G>(tiles.n_tiles)); \
        if (grid < 0) return -grid;                                                           \
        fq_bwd_ste_kernel<G, V><<<grid, kThreads, 0, st>>>(x, g, dx, tiles, qpd);             \
    }
    VSIQ_DISPATCH_GROUP_VEC(warp_group, vec8, CALL);
#undef CALL
    return (int)cudaGetLastError();
}

extern "C" int vsiq_fake_quant_fwd_bwd(const float* x, const float* g, float* y, float* dx,
                                       const vsiq_layout* layout, const vsiq_qparams* qp, vsiq_stream_t stream) {
    if (!x || !g || !y || !dx) return VSIQ_ERR_INVALID_ARG;
    if (int e = check_layout(layout)) return e;
    QPDev qpd;
    if (int e = fill_qp(qp, &qpd)) return e;
    if (layout->outer * layout->channels * layout->inner == 0) return VSIQ_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool warp_group = layout->inner < kWarpGroupMaxInner;
    const bool vec8 = aligned32(x) && aligned32(g) && aligned32(y) && aligned32(dx);
    Tiles tiles;
#define CALL(G, V)                                                                            \
    {                                                                                         \
        if (!make_tiles<G>(layout->outer, layout->channels, layout->inner, &tiles))           \
            return VSIQ_ERR_INVALID_ARG;                                                      \
        int grid = occupancy_grid(fq_fwd_bwd_kernel<G, V>, ctas_for_tiles<G>(tiles.n_tiles)); \
        if (grid < 0) return -grid;                                                           \
        fq_fwd_bwd_kernel<G, V><<<grid, kThreads, 0, st>>>(x, g, y, dx, tiles, qpd);          \
    }
    VSIQ_DISPATCH_GROUP_VEC(warp_group, vec8, CALL);
#undef CALL
    return (int)cudaGetLastError();
}

// Partials: 2 doubles per slot; slots = tiles (multi-row) or launched groups (single row) -- sized for
// the larger of the two so the caller need not know the launch geometry.
extern "C" size_t vsiq_lsq_bwd_workspace_bytes(const vsiq_layout* layout) {
    if (check_layout(layout)) return 0;
    Tiles tc, tw;
    size_t slots = 1;
    if (make_tiles<kThreads>(layout->outer, layout->channels, layout->inner, &tc)) slots = tc.n_tiles;
    if (layout->inner < kWarpGroupMaxInner && make_tiles<32>(layout->outer, layout->channels, layout->inner, &tw))
        slots = tw.n_tiles > slots ? tw.n_tiles : slots;
    return kWsHeader + slots * 2 * sizeof(double);
}

extern "C" int vsiq_lsq_bwd(const float* x, const float* g, float* dx, void* dscale, int dscale_dtype, void* dzp,
                            int dzp_dtype, const vsiq_layout* layout, const vsiq_qparams* qp, double grad_scale_host,
                            const float* grad_scale_dev, int mask_mode, void* workspace, size_t workspace_bytes,
                            vsiq_stream_t stream) {
    if (!x || !g || !dx || !dscale) return VSIQ_ERR_INVALID_ARG;
    if (int e = check_layout(layout)) return e;
    if (mask_mode != VSIQ_MASK_ROUNDED && mask_mode != VSIQ_MASK_FUNLSQ) return VSIQ_ERR_INVALID_ARG;
    if (mask_mode == VSIQ_MASK_FUNLSQ && dzp) return VSIQ_ERR_UNSUPPORTED;
    if ((dscale_dtype != VSIQ_F32 && dscale_dtype != VSIQ_F64) || (dzp && dzp_dtype != VSIQ_F32 && dzp_dtype != VSIQ_F64))
        return VSIQ_ERR_INVALID_ARG;
    QPDev qpd;
    if (int e = fill_qp(qp, &qpd)) return e;
    if (!workspace || workspace_bytes < vsiq_lsq_bwd_workspace_bytes(layout)) return VSIQ_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    if (layout->outer * layout->channels * layout->inner == 0) {
        // empty tensor: gradients are zero
        cudaError_t ce = cudaMemsetAsync(dscale, 0, (size_t)layout->channels * (dscale_dtype ? 8 : 4), st);
        if (ce == cudaSuccess && dzp) ce = cudaMemsetAsync(dzp, 0, (size_t)layout->channels * (dzp_dtype ? 8 : 4), st);
        return (int)ce;
    }
    const bool warp_group = layout->inner < kWarpGroupMaxInner;
    const bool vec8 = aligned32(x) && aligned32(g) && aligned32(dx);
    Tiles tiles;
#define LAUNCH(G, V, M, Z)                                                                                   \
    {                                                                                                        \
        if (!make_tiles<G>(layout->outer, layout->channels, layout->inner, &tiles))                          \
            return VSIQ_ERR_INVALID_ARG;                                                                     \
        int grid = occupancy_grid(lsq_bwd_kernel<G, V, M, Z>, ctas_for_tiles<G>(tiles.n_tiles));             \
        if (grid < 0) return -grid;                                                                          \
        lsq_bwd_kernel<G, V, M, Z><<<grid, kThreads, 0, st>>>(x, g, dx, tiles, qpd, workspace, dscale,       \
                                                              dscale_dtype, dzp, dzp_dtype, grad_scale_host, \
                                                              grad_scale_dev, layout->outer);                \
    }
#define CALL(G, V)                                   \
    {                                                \
        if (mask_mode == VSIQ_MASK_FUNLSQ) {         \
            LAUNCH(G, V, VSIQ_MASK_FUNLSQ, false);   \
        } else if (dzp) {                            \
            LAUNCH(G, V, VSIQ_MASK_ROUNDED, true);   \
        } else {                                     \
            LAUNCH(G, V, VSIQ_MASK_ROUNDED, false);  \
        }                                            \
    }
    VSIQ_DISPATCH_GROUP_VEC(warp_group, vec8, CALL);
#undef CALL
#undef LAUNCH
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------ self-test
// Counts the x bit patterns (all 2^32 of them) for which the reciprocal-based division differs from
// the IEEE division, for one divisor s.  NaN results compare equal to NaN.
namespace vsiq {
__global__ void __launch_bounds__(kThreads) division_selftest_kernel(float s, unsigned long long* mismatches) {
    const float r = __frcp_rn(s);
    const float as = fabsf(s);
    const bool fast = (as >= 9.094947017729282e-13f) && (as <= 1.099511627776e12f);
    unsigned int bad = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (1ull << 32); i += stride) {
        const float x = __uint_as_float((uint32_t)i);
        const float a = div_exact(x, s, r, fast);
        const float b = __fdiv_rn(x, s);
        const bool same = (__float_as_uint(a) == __float_as_uint(b)) || ((a != a) && (b != b));
        bad += same ? 0u : 1u;
    }
    bad = __reduce_add_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(mismatches, (unsigned long long)bad);
}
}  // namespace vsiq

extern "C" int vsiq_selftest_division(float s, unsigned long long* mismatches_dev, vsiq_stream_t stream) {
    if (!mismatches_dev) return VSIQ_ERR_INVALID_ARG;
    DeviceProps dp;
    if (int e = get_device_props(&dp)) return e;
    division_selftest_kernel<<<dp.sm_count * 8, kThreads, 0, (cudaStream_t)stream>>>(s, mismatches_dev);
    return (int)cudaGetLastError();
}
