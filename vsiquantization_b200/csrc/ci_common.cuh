// ci_common.cuh -- geometry, scheduling and 128-bit accessors shared by the channel-innermost (NHWC) kernels
// (channels_inner.cu: fake-quant epilogue forward / backward; observer.cu: per-channel statistics).
//
// A tensor is [rows = N*H*W][C] with the channel fastest; a 128-bit vector holds four consecutive channels and with
// G = C/4 vector groups per row only T = (256 / G) * G threads of a CTA are active, so thread t ALWAYS owns channel group
// t % G.  One "step" is one vector per active thread (T vectors = T/G whole rows); CTAs are persistent and every CTA owns
// a range of steps:
//   * static schedule (default): CTA b takes the contiguous range
//     [b*S/grid, (b+1)*S/grid) -- no atomics, no per-tile barrier, nothing to wait for before the first load is issued,
//     and all CTAs finish within one step of each other.  This is what the YOLO-sized feature maps (13-105 M elements,
//     20-150 us per launch) want: round 1's 16-step tiles from an atomic queue left a 444-CTA grid with 1-2 tiles each.
//   * dynamic schedule: tiles of 4..16 steps from the look-ahead atomic queue (common.cuh TileQueue), hardware-like
//     balancing;
//   * interleaved schedule: CTA b takes tiles b, b + grid, b + 2 grid, ... (no atomics, the grid moves through memory as
//     one window).  VSIQ_CI_SCHED=static|dynamic|interleaved and VSIQ_CI_TILE=<steps> force a choice (A/B knobs, read once).
#pragma once

#include "common.cuh"

namespace vsiq {

constexpr int kCiVec = 4;
constexpr int kCiUnroll = 4;       // 128-bit loads in flight per thread per input
constexpr int kCiBatches = 4;      // batches per dynamic tile
constexpr int kCiTileSteps = kCiUnroll * kCiBatches;  // steps per dynamic tile
constexpr int kCiMaxChannels = 1024;

struct CiGeom {
    int64_t n_vec;       // total vectors = rows * C / 4
    int64_t steps;       // ceil(n_vec / T)
    int channels;
    int groups;          // G = C / 4
    int threads;         // T: active threads per CTA (multiple of G)
    int sched;           // kCiStatic / kCiDynamic / kCiInterleaved
    int tile_steps;      // steps per tile (tiled schedules); a multiple of kCiUnroll
    uint32_t n_tiles;    // ceil(steps / tile_steps)
};
enum { kCiStatic = 0, kCiDynamic = 1, kCiInterleaved = 2 };



inline bool make_ci_geom(int64_t rows, int64_t channels, CiGeom* g) {
    if (channels < kCiVec || channels % kCiVec != 0 || channels > kCiMaxChannels || rows <= 0) return false;
    g->channels = (int)channels;
    g->groups = (int)(channels / kCiVec);
    g->threads = (kThreads / g->groups) * g->groups;
    g->n_vec = rows * (int64_t)g->groups;
    g->steps = (g->n_vec + g->threads - 1) / g->threads;
    if (g->steps <= 0 || g->steps >= (int64_t(1) << 31) - kCiTileSteps) return false;
    g->tile_steps = kCiTileSteps;
    g->n_tiles = (uint32_t)((g->steps + kCiTileSteps - 1) / kCiTileSteps);
    g->sched = kCiDynamic;
    return true;
}

// Persistent grid: every CTA resident at once (ctas_per_sm matches the kernel's __launch_bounds__), never more CTAs than
// units of work.  Also picks the schedule and the tile size: tiles shrink (down to one batch of `min_tile` steps) until
// every CTA can expect >= 8 of them, so a 13 M-element map is not dealt out as one or two big tiles per CTA.
inline int ci_pick_grid(CiGeom* g, int sm_count, int ctas_per_sm, int min_tile) {
    int64_t grid = (int64_t)sm_count * ctas_per_sm;
    int tile = kCiTileSteps;
    while (tile > min_tile && (g->steps + tile - 1) / tile < grid * 8) tile >>= 1;
    const int tov = ci_tile_override();
    if (tov > 0) tile = (tov + min_tile - 1) / min_tile * min_tile;
    g->tile_steps = tile;
    g->n_tiles = (uint32_t)((g->steps + tile - 1) / tile);
    const int ov = ci_sched_override();
    // dynamic balancing is worth ~15 % on B200 (SMs do not all see the same bandwidth; measured against both static
    // forms on 26-210 M-element maps) unless there is at most a tile or two per CTA to balance
    g->sched = ov >= 0 ? ov : (g->steps <= grid * (int64_t)min_tile * 2 ? kCiStatic : kCiDynamic);
    if (g->sched == kCiStatic) {
        if (grid > g->steps) grid = g->steps;
    } else if (grid > (int64_t)g->n_tiles) {
        grid = g->n_tiles;
    }
    return (int)grid;
}

// Step ranges are 32-bit (steps < 2^31).  A thread's range end is clamped so that the ragged last step of the tensor
// (only the first n_vec - (steps-1)*T threads own a vector there) needs no second bounds test in the hot loop.
struct CiRange {
    uint32_t s0, s1;
};
__device__ __forceinline__ uint32_t ci_clamp_end(const CiGeom& geo, uint32_t s1, int t) {
    const uint32_t last = (uint32_t)geo.steps - 1u;
    const int tail = (int)(geo.n_vec - (int64_t)last * geo.threads);  // vectors in the last step
    return (s1 > last && t >= tail) ? last : s1;
}
// The step range of this CTA under the static schedule.
__device__ __forceinline__ CiRange ci_static_range(const CiGeom& geo, int t) {
    const uint64_t S = (uint64_t)geo.steps, b = blockIdx.x, n = gridDim.x;
    CiRange r;
    r.s0 = (uint32_t)(S * b / n);
    r.s1 = ci_clamp_end(geo, (uint32_t)(S * (b + 1) / n), t);
    return r;
}
__device__ __forceinline__ CiRange ci_tile_range(const CiGeom& geo, uint32_t tile, int t) {
    CiRange r;
    r.s0 = tile * (uint32_t)geo.tile_steps;
    const uint32_t e = r.s0 + (uint32_t)geo.tile_steps;
    r.s1 = tile < geo.n_tiles ? ci_clamp_end(geo, e < (uint32_t)geo.steps ? e : (uint32_t)geo.steps, t) : r.s0;
    return r;
}

// Work distribution of the direct-load kernels: first range, then next ranges until ci_sched_next returns false.
struct CiSched {
    TileQueue tq;
    uint32_t tile;
};
__device__ __forceinline__ CiRange ci_sched_first(const CiGeom& geo, CiSched& sc, void* ws, uint32_t* s_tile, int t) {
    if (geo.sched == kCiStatic) return ci_static_range(geo, t);
    if (geo.sched == kCiDynamic) {
        tq_init(sc.tq, (unsigned int*)ws + 1, geo.n_tiles, s_tile);
        sc.tile = tq_current(sc.tq, s_tile);
    } else {
        sc.tile = blockIdx.x;
    }
    return ci_tile_range(geo, sc.tile, t);
}
__device__ __forceinline__ bool ci_sched_next(const CiGeom& geo, CiSched& sc, uint32_t* s_tile, int t, CiRange& r) {
    if (geo.sched == kCiStatic) return false;
    if (geo.sched == kCiDynamic) {
        tq_advance(sc.tq, s_tile);
        sc.tile = tq_current(sc.tq, s_tile);
    } else {
        sc.tile += gridDim.x;
    }
    if (sc.tile >= geo.n_tiles) return false;
    r = ci_tile_range(geo, sc.tile, t);
    return true;
}

// Activation fused in front of the quantiser (the fused layer's F.relu / F.silu, modules/fused.py:133).  SiLU follows
// ATen's CUDA kernels operation by operation (ActivationSiluKernel.cu: x / (1 + exp(-x)); backward
// dy * s * (1 + x * (1 - s)) with s = 1 / (1 + exp(-x)), the inner multiply-add contracted as nvcc does), so the fused
// epilogue reproduces F.silu followed by the quantiser ON THE SAME GPU; torch-CPU's vectorised exp differs in the last
// bits, so against the reference's CPU results SiLU layers agree to rounding of exp, not bit for bit.
enum { kActNone = 0, kActRelu = 1, kActSilu = 2 };
template <int ACT>
__device__ __forceinline__ float act_fwd(float x) {
    if (ACT == kActRelu) return max_nan(x, 0.0f);
    if (ACT == kActSilu) return __fdiv_rn(x, __fadd_rn(1.0f, expf(-x)));
    return x;
}
template <int ACT>
__device__ __forceinline__ float act_bwd(float x, float d) {
    if (ACT == kActRelu) return x > 0.0f ? d : 0.0f;
    if (ACT == kActSilu) {
        const float s = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));
        return __fmul_rn(__fmul_rn(d, s), __fmaf_rn(x, __fsub_rn(1.0f, s), 1.0f));
    }
    return d;
}

struct Vec4 {
    float v[4];
};
__device__ __forceinline__ Vec4 ld4(const float* p) {
    Vec4 r;
    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]) : "l"(p));
    return r;
}
__device__ __forceinline__ void st4(float* p, const Vec4& r) {
    asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]), "f"(r.v[3]) : "memory");
}

// Fixed-order combine of per-CTA records [n_rec][width] (fp64): one CTA owns 8 consecutive entries; inside a warp the
// low three lane bits select the entry and the high two a record phase, so one load instruction of a warp reads four
// 64-byte segments and the 32 (warp, phase) pairs of the CTA walk the records 32 apart.  A thread issues ALL its loads
// (up to 16 per pass: 512 records) before the first add, so the combine costs one memory round trip instead of one per
// record (ncu: 5.5 us -> the latency of a single L2 read for a 296-record launch).  The phases are then added by two
// xor-shuffles and the eight warps through shared memory, always in the same order.  Returns the sum for entry
// (base + (lane & 7)) in every thread with warp == 0 && lane < 8.
constexpr int kCombineEntries = 8;
constexpr int kCombineDepth = 16;
__device__ __forceinline__ double ci_combine8_sum(const double* __restrict__ records, size_t width, uint32_t n_rec,
                                                  size_t idx, bool valid, double (*s_part)[kCombineEntries]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t phase = (uint32_t)warp * 4u + (uint32_t)(lane >> 3);
    double v = 0.0;
    for (uint32_t r0 = phase; r0 < n_rec; r0 += 32u * kCombineDepth) {
        double a[kCombineDepth];
#pragma unroll
        for (int k = 0; k < kCombineDepth; ++k) {
            const uint32_t r = r0 + 32u * (uint32_t)k;
            a[k] = (valid && r < n_rec) ? __ldcg(records + (size_t)r * width + idx) : 0.0;
        }
#pragma unroll
        for (int st = 1; st < kCombineDepth; st <<= 1)
#pragma unroll
            for (int k = 0; k + st < kCombineDepth; k += 2 * st) a[k] += a[k + st];
        v += a[0];
    }
    v += __longlong_as_double(__shfl_xor_sync(0xffffffffu, __double_as_longlong(v), 8));
    v += __longlong_as_double(__shfl_xor_sync(0xffffffffu, __double_as_longlong(v), 16));
    if (lane < kCombineEntries) s_part[warp][lane] = v;
    __syncthreads();
    double out = 0.0;
    if (warp == 0 && lane < kCombineEntries) {
#pragma unroll
        for (int w = 0; w < kWarps; ++w) out += s_part[w][lane];
    }
    return out;
}

}  // namespace vsiq
