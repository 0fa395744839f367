// peer_exchange.cu -- the one exchange of BN re-estimation under data parallelism, as ONE kernel over NVLink peer memory.
//
// reestimate_BN_stats (reference: utils/estimate_bn.py:56-99) needs, per layer and batch, the per-channel batch moments of
// the conv output BEFORE the layer's output can be normalised; with the batch sharded over R GPUs the per-channel sums
// (sum x, sum x^2) of all ranks have to be added first (SyncBN-style).  As library calls that is
//     combine kernel -> ncclAllReduce(SUM) over [C,5] doubles -> moments kernel            (3 launches, 77 times per batch
// for YOLOv8m; the collective moves 10-40 KB and costs its fixed latency).  Here every rank runs ONE small kernel:
//   1. PUSH: thread c writes w * (sum x, sum x^2) of its channel into EVERY peer's receive area (plain stores through the
//      peer pointers, fire and forget over NVLink).  Each 8-byte word carries 4 bytes of payload and the 4-byte sequence
//      number of this exchange, so a word validates itself: no fence, no separate flag, no second round trip (the
//      "low-latency" layout NCCL uses for small messages);
//   2. POLL: it then reads the words its peers pushed for channel c from its OWN memory until each shows this sequence
//      number, adds the shards IN RANK ORDER -- every rank computes bit-identical sums -- and finishes the moments: mean,
//      biased / unbiased variance, running sums (:86-87).
// Buffers are plain cudaMalloc memory exported with cudaIpcGetMemHandle (one process per GPU, one node, <= 8 ranks).
// Two receive areas (sequence parity) suffice: a rank reaches exchange k+2 only after every peer has pushed k+1, i.e. has
// finished its kernel of exchange k.  A peer that never arrives trips the timeout: the kernel raises the buffer's error
// word instead of hanging the GPU.
#include <string.h>

#include "common.cuh"

namespace vsiq {

constexpr int kPeerMaxWorld = 8;
constexpr int kPeerMaxChannels = 1024;  // the channel-innermost kernels' own limit; wider layers keep the NCCL collective
constexpr int kPeerThreads = 1024;      // one channel per thread
constexpr size_t kPeerHeader = 256;     // [8] sequence counter, [9] error word (64-bit words)
constexpr size_t kPeerWordsPerChannel = 4;  // lo / hi halves of two doubles, each with the sequence number
constexpr size_t kPeerSrcBytes = (size_t)kPeerMaxChannels * kPeerWordsPerChannel * 8;
constexpr size_t kPeerAreaBytes = kPeerSrcBytes * kPeerMaxWorld;  // one receive area: a stripe per source rank
constexpr size_t kPeerBytes = kPeerHeader + 2 * kPeerAreaBytes;

struct PeerPtrs {
    void* p[kPeerMaxWorld];
};

__device__ __forceinline__ void st_words_sys(void* p, unsigned long long a, unsigned long long b) {
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void ld_words_sys(const void* p, unsigned long long& a, unsigned long long& b) {
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(kPeerThreads)
    bn_moments_exchange_kernel(const double* __restrict__ stats, double weight, double count, int C, PeerPtrs peers,
                               int rank, int world, unsigned long long timeout_ns, float* batch_mean,
                               float* batch_var_biased, float* batch_var_unbiased, float* mean_sum, float* var_sum) {
    __shared__ int s_failed;
    char* mine = (char*)peers.p[rank];
    unsigned long long* my_words = (unsigned long long*)mine;
    const unsigned long long seq = my_words[8] + 1ull;  // written by thread 0 at the very end of the previous exchange
    const unsigned long long tag = (seq & 0xffffffffull) << 32;
    const size_t area = kPeerHeader + (size_t)(seq & 1ull) * kPeerAreaBytes;
    if (my_words[9] != 0ull) return;  // an earlier exchange timed out: fail fast (the peers' kernels time out once, then do the same)
    if (threadIdx.x == 0) s_failed = 0;
    __syncthreads();
    const int c = threadIdx.x;
    double v1 = 0.0, v2 = 0.0;
    if (c < C) {
        v1 = weight * stats[(size_t)c * VSIQ_STATS_WIDTH + 3];
        v2 = weight * stats[(size_t)c * VSIQ_STATS_WIDTH + 4];
        const unsigned long long b1 = (unsigned long long)__double_as_longlong(v1);
        const unsigned long long b2 = (unsigned long long)__double_as_longlong(v2);
        const unsigned long long w0 = tag | (b1 & 0xffffffffull), w1 = tag | (b1 >> 32);
        const unsigned long long w2 = tag | (b2 & 0xffffffffull), w3 = tag | (b2 >> 32);
        const size_t off = area + (size_t)rank * kPeerSrcBytes + (size_t)c * kPeerWordsPerChannel * 8;
#pragma unroll
        for (int r = 0; r < kPeerMaxWorld; ++r) {
            if (r < world && r != rank) {
                char* dst = (char*)peers.p[r] + off;
                st_words_sys(dst, w0, w1);
                st_words_sys(dst + 16, w2, w3);
            }
        }
    }
    double s1 = 0.0, s2 = 0.0;
    bool failed = false;
    if (c < C) {
        const unsigned long long t0 = global_timer_ns();
        for (int r = 0; r < world; ++r) {  // fixed order: every rank gets the same bits
            double a = v1, b = v2;
            if (r != rank) {
                const char* src = mine + area + (size_t)r * kPeerSrcBytes + (size_t)c * kPeerWordsPerChannel * 8;
                unsigned long long w0, w1, w2, w3;
                unsigned int spins = 0;
                for (;;) {
                    ld_words_sys(src, w0, w1);
                    ld_words_sys(src + 16, w2, w3);
                    if ((w0 & w1 & w2 & w3 & 0xffffffff00000000ull) == tag &&
                        ((w0 | w1 | w2 | w3) & 0xffffffff00000000ull) == tag)
                        break;
                    if ((++spins & 63u) == 0 && global_timer_ns() - t0 > timeout_ns) {
                        failed = true;
                        break;
                    }
                }
                if (failed) break;
                a = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
                b = __longlong_as_double((long long)((w2 & 0xffffffffull) | (w3 << 32)));
            }
            s1 += a;
            s2 += b;
        }
        if (failed) s_failed = 1;
    }
    __syncthreads();
    if (s_failed) {
        if (threadIdx.x == 0) my_words[9] = seq;  // error word: the sequence number that timed out
        return;                                  // the sequence counter does not advance; outputs are left untouched
    }
    if (c < C) {
        const double mean = s1 / count;  // as bn_moments_finalize_kernel (observer.cu)
        double var_b = s2 / count - mean * mean;
        var_b = var_b > 0.0 ? var_b : 0.0;
        const double var_u = var_b * (count / (count - 1.0));
        const float m32 = (float)mean, vu32 = (float)var_u;
        if (batch_mean) batch_mean[c] = m32;
        if (batch_var_biased) batch_var_biased[c] = (float)var_b;
        if (batch_var_unbiased) batch_var_unbiased[c] = vu32;
        if (mean_sum) mean_sum[c] = __fadd_rn(mean_sum[c], m32);  // estimate_bn.py:86
        if (var_sum) var_sum[c] = __fadd_rn(var_sum[c], vu32);     // estimate_bn.py:87
    }
    if (threadIdx.x == 0) my_words[8] = seq;
}

}  // namespace vsiq

using namespace vsiq;

extern "C" size_t vsiq_peer_buffer_bytes(void) { return kPeerBytes; }
extern "C" int vsiq_peer_max_world(void) { return kPeerMaxWorld; }
extern "C" int vsiq_peer_max_channels(void) { return kPeerMaxChannels; }

extern "C" int vsiq_peer_alloc(void** buffer, unsigned char* handle64) {
    if (!buffer || !handle64) return VSIQ_ERR_INVALID_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, kPeerBytes);
    if (e != cudaSuccess) return (int)e;
    if ((e = cudaMemset(p, 0, kPeerBytes)) != cudaSuccess || (e = cudaDeviceSynchronize()) != cudaSuccess) {
        cudaFree(p);
        return (int)e;
    }
    cudaIpcMemHandle_t h;
    if ((e = cudaIpcGetMemHandle(&h, p)) != cudaSuccess) {
        cudaFree(p);
        return (int)e;
    }
    memcpy(handle64, &h, 64);
    *buffer = p;
    return VSIQ_OK;
}

extern "C" int vsiq_peer_open(const unsigned char* handle64, void** buffer) {
    if (!handle64 || !buffer) return VSIQ_ERR_INVALID_ARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    return (int)cudaIpcOpenMemHandle(buffer, h, cudaIpcMemLazyEnablePeerAccess);
}

extern "C" int vsiq_peer_close(void* buffer) { return buffer ? (int)cudaIpcCloseMemHandle(buffer) : VSIQ_OK; }
extern "C" int vsiq_peer_free(void* buffer) { return buffer ? (int)cudaFree(buffer) : VSIQ_OK; }

extern "C" int vsiq_peer_status(const void* own_buffer, uint64_t* sequence, uint64_t* failed_sequence) {
    if (!own_buffer) return VSIQ_ERR_INVALID_ARG;
    unsigned long long w[2] = {0, 0};
    cudaError_t e = cudaMemcpy(w, (const unsigned long long*)own_buffer + 8, sizeof(w), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return (int)e;
    if (sequence) *sequence = w[0];
    if (failed_sequence) *failed_sequence = w[1];
    return VSIQ_OK;
}

extern "C" int vsiq_bn_moments_exchange(const double* stats, double weight, double global_count, int64_t channels,
                                        void* const* peer_buffers, int rank, int world, double timeout_s,
                                        float* batch_mean, float* batch_var_biased, float* batch_var_unbiased,
                                        float* mean_sum, float* var_sum, vsiq_stream_t stream) {
    if (!stats || !peer_buffers || channels <= 0 || !(global_count > 1.0)) return VSIQ_ERR_INVALID_ARG;
    if (world < 1 || world > kPeerMaxWorld || rank < 0 || rank >= world) return VSIQ_ERR_INVALID_ARG;
    if (channels > kPeerMaxChannels) return VSIQ_ERR_UNSUPPORTED;
    if (!(timeout_s > 0.0)) timeout_s = 10.0;
    PeerPtrs pp;
    for (int r = 0; r < kPeerMaxWorld; ++r) pp.p[r] = r < world ? peer_buffers[r] : nullptr;
    for (int r = 0; r < world; ++r)
        if (!pp.p[r]) return VSIQ_ERR_INVALID_ARG;
    bn_moments_exchange_kernel<<<1, kPeerThreads, 0, (cudaStream_t)stream>>>(
        stats, weight, global_count, (int)channels, pp, rank, world, (unsigned long long)(timeout_s * 1e9), batch_mean,
        batch_var_biased, batch_var_unbiased, mean_sum, var_sum);
    return (int)cudaGetLastError();
}
