// common.cuh -- device-side building blocks shared by every kernel of libvsiq.so (sm_100a).
//
// All kernels on this path are HBM-bound streaming passes over fp32 tensors, so the whole design is
// about keeping enough 256-bit requests in flight per SM and never touching a byte twice:
//   * a tensor is walked as [rows = outer*channels][inner]; a "tile" is a contiguous chunk of ONE row,
//     so the quantisation parameters are uniform over a tile and the inner loop has no index math;
//   * a tile is owned by a thread GROUP: a whole CTA (256 threads, 8192-element tiles) for long rows,
//     or one warp (1024-element tiles) for short rows (20x20 feature maps, small filters), so short
//     rows still keep every SM's load queue full;
//   * inside a tile every thread issues UNROLL 256-bit loads per input (LDG.E.256, sm_100+) before any
//     arithmetic, computes, then issues 256-bit stores; ragged heads/tails and unaligned rows fall to
//     a scalar peel, fully unaligned tensors to the V=1 instantiation;
//   * grids are capped at (SM count x resident CTAs) and walk tiles with a grid stride.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vsiq.h"

namespace vsiq {

constexpr int kThreads = 256;  // threads per CTA
constexpr int kWarps = kThreads / 32;
constexpr int kVec = 8;        // fp32 lanes per 256-bit access
constexpr int kUnroll = 2;     // vectors in flight per thread per input array
constexpr int kBatchesPerTile = 2;

template <int GROUP>
struct TileGeom {
    static constexpr int kBatch = GROUP * kVec * kUnroll;     // elements one group moves per batch
    static constexpr int kTile = kBatch * kBatchesPerTile;    // elements per tile
};
constexpr int kCtaTile = TileGeom<kThreads>::kTile;   // 8192
constexpr int kWarpTile = TileGeom<32>::kTile;        // 1024
constexpr int64_t kWarpGroupMaxInner = 2048;          // rows shorter than this are walked warp-per-tile

// ---------------------------------------------------------------------------------------------
// 256-bit / 32-bit streaming global accesses.  Inputs are read exactly once: bypass L1 allocation.
// ---------------------------------------------------------------------------------------------
template <int V>
struct Vec {
    float v[V];
};

__device__ __forceinline__ Vec<8> ld_stream(const float* p, Vec<8>*) {
    Vec<8> r;
    asm("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]),
          "=f"(r.v[7])
        : "l"(p));
    return r;
}
__device__ __forceinline__ Vec<1> ld_stream(const float* p, Vec<1>*) {
    Vec<1> r;
    asm("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r.v[0]) : "l"(p));
    return r;
}
__device__ __forceinline__ float ld_stream1(const float* p) {
    float r;
    asm("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream(float* p, const Vec<8>& r) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(r.v[0]), "f"(r.v[1]),
                 "f"(r.v[2]), "f"(r.v[3]), "f"(r.v[4]), "f"(r.v[5]), "f"(r.v[6]), "f"(r.v[7])
                 : "memory");
}
__device__ __forceinline__ void st_stream(float* p, const Vec<1>& r) { *p = r.v[0]; }

// ---------------------------------------------------------------------------------------------
// Quantisation parameters
// ---------------------------------------------------------------------------------------------
struct QPDev {  // kernel-argument mirror of vsiq_qparams
    const void* scale;
    const void* zp;
    int scale_f64;
    int zp_f64;
    float scale_host;
    float zp_host;
    int zp_learned;
    float lo;
    float hi;
};

struct QP {  // per-tile (uniform) values
    float s;   // scale rounded to fp32 (ATen rounds the fp64 0-dim Parameter the same way)
    float r;   // RN(1/s), hoisted out of the element loop
    float z;   // effective zero-point used by the forward
    float zf;  // raw (float) zero-point parameter
    float lo, hi;
    bool fast; // |s| in [2^-40, 2^40]: the reciprocal-based exact division below is valid
};

// IEEE-754 round-to-nearest x / s without a per-element MUFU.RCP + FCHK (the XU pipe issues only 16
// lanes/clk/SM, which would bound these kernels before HBM does).  With r = RN(1/s):
//     q0 = RN(x*r); q1 = RN(q0 + (x - q0*s)*r)   -> faithful (error < 2^-46 relative before rounding)
//     q2 = RN(q1 + (x - q1*s)*r)                 -> correctly rounded (Markstein's theorem)
// The residuals are exact FMAs as long as nothing under/overflows: guaranteed for
// 2^-60 <= |x| < 2^61 and 2^-40 <= |s| <= 2^40.  x == 0 returns x*r (signed zero of the right sign);
// everything else (denormal, huge, inf, NaN) takes the IEEE division instruction sequence.
// tests/test_gpu_division.py checks it against __fdiv_rn over ALL 2^32 values of x for many scales.
static __device__ __noinline__ float div_ieee(float x, float s) { return __fdiv_rn(x, s); }

__device__ __forceinline__ float div_exact(float x, float s, float r, bool fast) {
    const float q0 = __fmul_rn(x, r);
    const float e0 = __fmaf_rn(-q0, s, x);
    const float q1 = __fmaf_rn(e0, r, q0);
    const float e1 = __fmaf_rn(-q1, s, x);
    const float q2 = __fmaf_rn(e1, r, q1);
    const uint32_t ex = (__float_as_uint(x) >> 23) & 0xffu;
    const bool in_range = fast && ((ex - 67u) < 121u);  // biased exponent in [67, 187]
    float q = in_range ? q2 : q0;
    if (!in_range && (x != 0.0f || !fast)) q = div_ieee(x, s);  // rare: denormal / huge / inf / NaN / odd scale
    return q;
}

// torch.clamp semantics: NaN propagates, -0.0 survives a 0 lower bound (fminf/fmaxf would lose both).
__device__ __forceinline__ float clamp_torch(float r, float lo, float hi) {
    r = (r < lo) ? lo : r;
    r = (r > hi) ? hi : r;
    return r;
}

__device__ __forceinline__ QP load_qp(const QPDev& d, int64_t c) {
    QP q;
    q.lo = d.lo;
    q.hi = d.hi;
    if (d.scale)
        q.s = d.scale_f64 ? (float)__ldg((const double*)d.scale + c) : __ldg((const float*)d.scale + c);
    else
        q.s = d.scale_host;
    if (d.zp)
        q.zf = d.zp_f64 ? (float)__ldg((const double*)d.zp + c) : __ldg((const float*)d.zp + c);
    else
        q.zf = d.zp_host;
    q.z = d.zp_learned ? clamp_torch(rintf(q.zf), q.lo, q.hi) : q.zf;
    q.r = __frcp_rn(q.s);
    const float as = fabsf(q.s);
    q.fast = (as >= 9.094947017729282e-13f) && (as <= 1.099511627776e12f);  // 2^-40 .. 2^40
    return q;
}

// The reference's forward, one element (quantizers/uniform.py:54-55,95): every op individually rounded.
//   r = rint(x / s + z);  q = clamp(r);  y = (q - z) * s
__device__ __forceinline__ float fq_round(float x, const QP& p) {
    return rintf(__fadd_rn(div_exact(x, p.s, p.r, p.fast), p.z));
}
__device__ __forceinline__ float fq_dequant(float q, const QP& p) { return __fmul_rn(__fsub_rn(q, p.z), p.s); }
__device__ __forceinline__ bool fq_inrange(float r, const QP& p) { return (r >= p.lo) && (r <= p.hi); }
// autograd of the forward: dx = where(m, g*s, 0) / s   (mul-, clamp-, STE-, add-, div-backward)
__device__ __forceinline__ float ste_dx(float g, bool m, const QP& p) {
    float gd = __fmul_rn(g, p.s);
    return div_exact(m ? gd : 0.0f, p.s, p.r, p.fast);
}

struct OpBase {
    __device__ __forceinline__ void vec_done() {}
};

// ---------------------------------------------------------------------------------------------
// Tile schedule
// ---------------------------------------------------------------------------------------------
struct Tiles {
    int64_t rows;       // outer * channels
    int64_t channels;
    int64_t inner;
    uint32_t chunks;    // tiles per row
    uint32_t n_tiles;   // rows * chunks  (< 2^31)
};

template <int GROUP>
__host__ __device__ inline bool make_tiles(int64_t outer, int64_t channels, int64_t inner, Tiles* t) {
    t->rows = outer * channels;
    t->channels = channels;
    t->inner = inner;
    const int64_t tile = TileGeom<GROUP>::kTile;
    int64_t chunks = (inner + tile - 1) / tile;
    int64_t n = t->rows * chunks;
    if (chunks <= 0 || n <= 0 || n >= (int64_t(1) << 31)) return false;
    t->chunks = (uint32_t)chunks;
    t->n_tiles = (uint32_t)n;
    return true;
}

template <int GROUP>
struct TileCursor {  // which tile this thread's group owns right now
    int64_t offset;  // element offset of the tile in the tensor
    int len;         // elements in the tile
    int64_t row;
    int64_t channel;
};

template <int GROUP>
__device__ __forceinline__ uint32_t group_index() {
    return GROUP == kThreads ? blockIdx.x : blockIdx.x * kWarps + (threadIdx.x >> 5);
}
template <int GROUP>
__device__ __forceinline__ uint32_t group_count() {
    return GROUP == kThreads ? gridDim.x : gridDim.x * kWarps;
}
template <int GROUP>
__device__ __forceinline__ int group_tid() {
    return GROUP == kThreads ? threadIdx.x : (threadIdx.x & 31);
}

template <int GROUP>
__device__ __forceinline__ TileCursor<GROUP> tile_at(const Tiles& t, uint32_t idx) {
    TileCursor<GROUP> c;
    uint32_t row = idx / t.chunks;
    uint32_t chunk = idx - row * t.chunks;
    int64_t start = (int64_t)chunk * TileGeom<GROUP>::kTile;
    int64_t rem = t.inner - start;
    c.len = rem < TileGeom<GROUP>::kTile ? (int)rem : TileGeom<GROUP>::kTile;
    c.row = row;
    c.channel = t.channels == 1 ? 0 : (int64_t)(row % (uint32_t)t.channels);
    c.offset = (int64_t)row * t.inner + start;
    return c;
}

// ---------------------------------------------------------------------------------------------
// span_apply: run Op over `len` consecutive elements starting at element `off`, cooperatively by a
// GROUP of threads.  Op::apply(const float (&in)[NIN], float (&out)[NOUT]) handles one element;
// Op::vec_done() is called after every vector (<= 8 elements) so reducing ops can spill their short
// fp32 partials into fp64 accumulators.
// Base pointers are 32-byte aligned when V == 8 (checked on the host); the row offset need not be.
// ---------------------------------------------------------------------------------------------
template <int GROUP, int V, int NIN, int NOUT, class Op>
__device__ __forceinline__ void span_apply(const float* const (&in)[NIN],
                                           float* const (&out)[NOUT > 0 ? NOUT : 1], int64_t off,
                                           int len, Op& op) {
    const int tid = group_tid<GROUP>();
    int head = 0;
    if (V > 1) {
        head = (int)((V - (off & (V - 1))) & (V - 1));
        head = head < len ? head : len;
        if (tid < head) {
            float a[NIN], o[NOUT > 0 ? NOUT : 1];
#pragma unroll
            for (int k = 0; k < NIN; ++k) a[k] = ld_stream1(in[k] + off + tid);
            op.apply(a, o);
            op.vec_done();
#pragma unroll
            for (int k = 0; k < NOUT; ++k) out[k][off + tid] = o[k];
        }
    }
    const int nvec = (len - head) / V;
    const int64_t vbase = off + head;
    for (int b = 0; b < nvec; b += GROUP * kUnroll) {
        Vec<V> vin[NIN][kUnroll];
        bool ok[kUnroll];
#pragma unroll
        for (int j = 0; j < kUnroll; ++j) {
            const int vi = b + j * GROUP + tid;
            ok[j] = vi < nvec;
            if (ok[j]) {
#pragma unroll
                for (int k = 0; k < NIN; ++k)
                    vin[k][j] = ld_stream(in[k] + vbase + (int64_t)vi * V, (Vec<V>*)nullptr);
            }
        }
#pragma unroll
        for (int j = 0; j < kUnroll; ++j) {
            if (ok[j]) {
                const int vi = b + j * GROUP + tid;
                Vec<V> vout[NOUT > 0 ? NOUT : 1];
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    float a[NIN], o[NOUT > 0 ? NOUT : 1];
#pragma unroll
                    for (int k = 0; k < NIN; ++k) a[k] = vin[k][j].v[e];
                    op.apply(a, o);
#pragma unroll
                    for (int k = 0; k < NOUT; ++k) vout[k].v[e] = o[k];
                }
                op.vec_done();
#pragma unroll
                for (int k = 0; k < NOUT; ++k) st_stream(out[k] + vbase + (int64_t)vi * V, vout[k]);
            }
        }
    }
    if (V > 1) {
        const int done = head + nvec * V;
        const int i = done + tid;
        if (i < len) {
            float a[NIN], o[NOUT > 0 ? NOUT : 1];
#pragma unroll
            for (int k = 0; k < NIN; ++k) a[k] = ld_stream1(in[k] + off + i);
            op.apply(a, o);
            op.vec_done();
#pragma unroll
            for (int k = 0; k < NOUT; ++k) out[k][off + i] = o[k];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Reductions.  Within a warp: shuffles.  Across the warps of a CTA: shared memory, fixed order.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double shfl_down_d(double v, int d) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_down_sync(0xffffffffu, lo, d);
    hi = __shfl_down_sync(0xffffffffu, hi, d);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += shfl_down_d(v, d);
    return v;  // valid in lane 0
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    return v;
}
// NaN-propagating min / max (torch.min / torch.max semantics)
__device__ __forceinline__ float nanmin(float a, float b) { return (a != a) ? a : ((b != b) ? b : (b < a ? b : a)); }
__device__ __forceinline__ float nanmax(float a, float b) { return (a != a) ? a : ((b != b) ? b : (b > a ? b : a)); }
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = nanmin(v, __shfl_down_sync(0xffffffffu, v, d));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = nanmax(v, __shfl_down_sync(0xffffffffu, v, d));
    return v;
}

// "last CTA finishes" ticket: returns true in every thread of the CTA that arrives last.  All global
// writes issued by this CTA before the call are visible to the last CTA after it.
__device__ __forceinline__ bool last_cta_ticket(unsigned int* counter) {
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(counter, 1u);
        s_last = (t == gridDim.x - 1);
        if (s_last) *counter = 0;  // leave the workspace ready for the next launch
    }
    __syncthreads();
    const bool last = s_last != 0;
    if (last) __threadfence();
    return last;
}

// Workspace layout shared by the reducing kernels: a 256-byte header (ticket) then fp64 partials.
constexpr size_t kWsHeader = 256;
__device__ __forceinline__ double* ws_partials(void* ws) { return (double*)((char*)ws + kWsHeader); }

// ---------------------------------------------------------------------------------------------
// Host-side helpers (abi)
// ---------------------------------------------------------------------------------------------
struct DeviceProps {
    int sm_count;
    int cc_major, cc_minor;
};
int get_device_props(DeviceProps* out);         // cached per device
int grid_for(uint32_t n_groups_wanted, int ctas_per_sm);  // capped persistent grid (>= 1)
bool aligned32(const void* p);
int check_layout(const vsiq_layout* l);
int fill_qp(const vsiq_qparams* in, QPDev* out);

}  // namespace vsiq
