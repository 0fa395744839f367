// ci_common.cuh -- geometry and 128-bit accessors shared by the channel-innermost (NHWC) kernels
// (channels_inner.cu: fake-quant epilogue forward / backward; observer.cu: per-channel statistics).
#pragma once

#include "common.cuh"

namespace vsiq {

constexpr int kCiVec = 4;
constexpr int kCiUnroll = 4;       // 128-bit loads in flight per thread per input
constexpr int kCiBatches = 4;      // batches per tile
constexpr int kCiMaxChannels = 1024;

struct CiGeom {
    int64_t n_vec;       // total vectors = rows * C / 4
    int channels;
    int groups;          // G = C / 4
    int threads;         // T: active threads per CTA (multiple of G)
    int tile_vecs;       // T * unroll * batches
    uint32_t n_tiles;
};

inline bool make_ci_geom(int64_t rows, int64_t channels, CiGeom* g) {
    if (channels < kCiVec || channels % kCiVec != 0 || channels > kCiMaxChannels || rows <= 0) return false;
    g->channels = (int)channels;
    g->groups = (int)(channels / kCiVec);
    g->threads = (kThreads / g->groups) * g->groups;
    g->tile_vecs = g->threads * kCiUnroll * kCiBatches;
    g->n_vec = rows * (int64_t)g->groups;
    const int64_t nt = (g->n_vec + g->tile_vecs - 1) / g->tile_vecs;
    if (nt <= 0 || nt >= (int64_t(1) << 31)) return false;
    g->n_tiles = (uint32_t)nt;
    return true;
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

}  // namespace vsiq
