// dp_exchange.cu — the data-parallel exchange step as ONE kernel over NVLink / NVSwitch peer memory.
//
// Reference: DistributedDataParallel's gradient all-reduce followed by optimizer.step()
// (main_mimic.py:44-48, utils/utils.py:179-185, run_epochs.py:130-131).  Here every rank owns 1/world of the flat
// buffers and one kernel does, for its slice,
//   reduce-scatter : g = sum_r grad_r[i]   read straight out of every peer's gradient buffer (P2P loads, fixed rank
//                    order -> the same bits on every run)
//   Adam           : m, v, p update for the slice (moments of a slice live on its owner only)
//   all-gather     : the new parameters are stored into EVERY rank's parameter buffer (P2P stores)
// so the transfers overlap the math element by element and nothing is staged.  The per-rank traffic equals a ring
// all-reduce's ((world-1)/world of the buffer in, the same out) while the optimizer's own HBM traffic drops by 1/world.
//
// Synchronisation: two flag barriers in a small symmetric flag array, stamped with a monotonically increasing
// epoch kept in device memory (the kernel is replayed from a CUDA graph: no host-side argument changes).
//   flags[r]         = epoch : rank r's gradients are final (its backward finished — stream order)
//   flags[world + r] = epoch : rank r has finished READING my gradients and WRITING my parameters
// The kernel does not complete before all peers have signalled the second flag, so whatever follows it in the stream
// (the next step's zero_grad / forward) is safe.  Waits are bounded by a CONFIGURABLE window (MOPOE_DP_TIMEOUT_S, default
// 600 s, 0 = wait forever): ranks may legitimately arrive far apart (rank-0-only evaluation / checkpointing, a loader
// stall), so the default is in NCCL-watchdog territory.  A wait that does expire does NOT trap (that would poison the
// CUDA context of every waiting rank): it records the missing flag in state[2], stops waiting, and the host raises from
// PeerExchange.check() — the step's numbers are garbage but the process can still report and shut down cleanly.
#include "common.cuh"

constexpr int DPX_MAX_WORLD = 16;
constexpr int DPX_THREADS = 256;
constexpr int DPX_UNROLL = 4;

struct DpxPeers {
    const float* grad[DPX_MAX_WORLD];
    float* param[DPX_MAX_WORLD];
    unsigned int* flags[DPX_MAX_WORLD];
};
static_assert(sizeof(DpxPeers) == sizeof(mopoe_dp_peers_t), "mopoe_dp_peers_t layout");

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// peer data: read around L1 (the lines may have been cached by an earlier step of this persistent grid-stride loop)
__device__ __forceinline__ float4 ld_peer(const float4* p) {
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// NVSwitch multicast (NVLS): one load returns the sum over every rank's copy (reduced inside the switch), one store
// lands in every rank's copy — the per-rank NVLink traffic drops from (world-1)/world of the buffer to 1/world.
__device__ __forceinline__ float4 mc_ld_reduce(const float4* mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st(float4* mc, const float4& v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

// threads [0, world) of the block poll one flag each until it reaches `epoch`
__device__ __forceinline__ void wait_flags(const unsigned int* flags, int world, unsigned int epoch,
                                           unsigned long long timeout_ns, unsigned int* state, int which) {
    if ((int)threadIdx.x < world) {
        const unsigned long long t0 = globaltimer_ns();
        while ((int)(ld_acquire_sys(flags + threadIdx.x) - epoch) < 0) {
            __nanosleep(64);
            if (timeout_ns && globaltimer_ns() - t0 > timeout_ns) {
                // error flag for the host (PeerExchange.check): 1 + barrier * 16 + the rank that never arrived
                atomicCAS(state + 2, 0u, 1u + (unsigned)which * 16u + threadIdx.x);
                break;
            }
        }
    }
    __syncthreads();
}

template <bool MC>
__global__ void __launch_bounds__(DPX_THREADS)
dp_adam_exchange_kernel(const DpxPeers peers, const float* __restrict__ mc_grad, float* __restrict__ mc_param,
                        float* __restrict__ m, float* __restrict__ v, long long n4, int rank,
                        int world, unsigned int* state, const float* __restrict__ coef, float b1, float b2, float eps,
                        float gscale, unsigned long long timeout_ns) {
    __shared__ int s_last;
    const unsigned int epoch = *reinterpret_cast<volatile unsigned int*>(state);
    unsigned int* const myflags = peers.flags[rank];
    // ---- barrier 1: everybody's gradients are final ---------------------------------------------------------------
    if (blockIdx.x == 0 && (int)threadIdx.x < world) st_release_sys(peers.flags[threadIdx.x] + rank, epoch);
    wait_flags(myflags, world, epoch, timeout_ns, state, 0);

    const float lr_c = coef[0], inv_sqrt_bc2 = coef[1];
    const long long per = (n4 + world - 1) / world;
    const long long begin = (long long)rank * per, end = begin + per < n4 ? begin + per : n4;
    float4* const p_own = reinterpret_cast<float4*>(peers.param[rank]);
    // DPX_UNROLL independent float4 per thread and iteration, every remote load issued before the first use: NVLink
    // round trips are microseconds, so the bytes in flight per SM — not the instruction count — set the throughput
    const long long stride = (long long)gridDim.x * DPX_THREADS * DPX_UNROLL;
    for (long long i0 = begin + (long long)blockIdx.x * DPX_THREADS * DPX_UNROLL + threadIdx.x; i0 < end; i0 += stride) {
        float4 g[DPX_UNROLL], pp[DPX_UNROLL], mm[DPX_UNROLL], vv[DPX_UNROLL];
#pragma unroll
        for (int u = 0; u < DPX_UNROLL; ++u) {
            const long long i = i0 + (long long)u * DPX_THREADS;
            if (i < end) g[u] = MC ? mc_ld_reduce(reinterpret_cast<const float4*>(mc_grad) + i)
                                   : ld_peer(reinterpret_cast<const float4*>(peers.grad[0]) + i);
        }
        if (!MC) {
            for (int r = 1; r < world; ++r) {
                float4 t[DPX_UNROLL];
#pragma unroll
                for (int u = 0; u < DPX_UNROLL; ++u) {
                    const long long i = i0 + (long long)u * DPX_THREADS;
                    if (i < end) t[u] = ld_peer(reinterpret_cast<const float4*>(peers.grad[r]) + i);
                }
#pragma unroll
                for (int u = 0; u < DPX_UNROLL; ++u) {          // rank order: fixed summation order
                    g[u].x += t[u].x; g[u].y += t[u].y; g[u].z += t[u].z; g[u].w += t[u].w;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < DPX_UNROLL; ++u) {
            const long long i = i0 + (long long)u * DPX_THREADS;
            if (i < end) {
                pp[u] = p_own[i];
                mm[u] = reinterpret_cast<float4*>(m)[i];
                vv[u] = reinterpret_cast<float4*>(v)[i];
            }
        }
#pragma unroll
        for (int u = 0; u < DPX_UNROLL; ++u) {
            const long long i = i0 + (long long)u * DPX_THREADS;
            if (i >= end) continue;
            float* P = &pp[u].x; float* G = &g[u].x; float* M = &mm[u].x; float* V = &vv[u].x;
#pragma unroll
            for (int k = 0; k < 4; ++k) {          // same arithmetic as adam_kernel (elementwise.cu)
                const float gk = G[k] * gscale;
                M[k] = b1 * M[k] + (1.f - b1) * gk;
                V[k] = b2 * V[k] + (1.f - b2) * gk * gk;
                P[k] -= lr_c * M[k] / (sqrtf(V[k]) * inv_sqrt_bc2 + eps);
            }
            reinterpret_cast<float4*>(m)[i] = mm[u];
            reinterpret_cast<float4*>(v)[i] = vv[u];
            if (MC) {
                mc_st(reinterpret_cast<float4*>(mc_param) + i, pp[u]);
            } else {
                for (int r = 0; r < world; ++r) reinterpret_cast<float4*>(peers.param[r])[i] = pp[u];
            }
        }
    }
    // ---- barrier 2: my stores have landed everywhere; wait until every peer is done with my buffers ---------------
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(state + 1, 1u);
        __threadfence();
        s_last = done == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last) {
        if ((int)threadIdx.x < world) st_release_sys(peers.flags[threadIdx.x] + world + rank, epoch);
        wait_flags(myflags + world, world, epoch, timeout_ns, state, 1);
        if (threadIdx.x == 0) {
            state[1] = 0;
            *reinterpret_cast<volatile unsigned int*>(state) = epoch + 1;
        }
    }
}

extern "C" int mopoe_dp_adam_exchange_ex(const mopoe_dp_peers_t* peers, const float* mc_grad, float* mc_param, float* m, float* v,
                                         int64_t n, int rank, int world, uint32_t* state, const float* coef, float beta1,
                                         float beta2, float eps, float grad_scale, int max_blocks, void* stream) {
    MOPOE_REQUIRE(peers && m && v && state && coef, "dp_adam_exchange: null argument");
    MOPOE_REQUIRE(world >= 1 && world <= DPX_MAX_WORLD && rank >= 0 && rank < world, "dp_adam_exchange: rank %d / world %d", rank,
                  world);
    MOPOE_REQUIRE(n > 0 && n % 4 == 0, "dp_adam_exchange: n=%lld must be a positive multiple of 4", (long long)n);
    uintptr_t bits = (uintptr_t)m | (uintptr_t)v;
    for (int r = 0; r < world; ++r) {
        MOPOE_REQUIRE(peers->grad[r] && peers->param[r] && peers->flags[r], "dp_adam_exchange: null peer pointer (rank %d)", r);
        bits |= (uintptr_t)peers->grad[r] | (uintptr_t)peers->param[r];
    }
    MOPOE_REQUIRE((bits & 15) == 0, "dp_adam_exchange: unaligned buffers");
    DpxPeers pk;
    memcpy(&pk, peers, sizeof(pk));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    const long long n4 = n / 4, per = (n4 + world - 1) / world;
    long long blocks = (per + DPX_THREADS * DPX_UNROLL - 1) / (DPX_THREADS * DPX_UNROLL);
    if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
    // a bucket exchanged UNDER the rest of the backward pass runs on a small grid: it has milliseconds to hide in and
    // must not take the SMs away from the kernels it overlaps with
    if (max_blocks > 0 && blocks > max_blocks) blocks = max_blocks;
    if (blocks < 1) blocks = 1;
    MOPOE_REQUIRE((mc_grad == nullptr) == (mc_param == nullptr), "dp_adam_exchange: give both multicast addresses or neither");
    MOPOE_REQUIRE((((uintptr_t)mc_grad | (uintptr_t)mc_param) & 15) == 0, "dp_adam_exchange: unaligned multicast address");
    static long long timeout_s = -1;
    if (timeout_s < 0) {
        const char* e = getenv("MOPOE_DP_TIMEOUT_S");
        timeout_s = e ? atoll(e) : 600;
        if (timeout_s < 0) timeout_s = 0;
    }
    const unsigned long long timeout_ns = (unsigned long long)timeout_s * 1000000000ull;
    if (mc_grad)
        dp_adam_exchange_kernel<true><<<(unsigned)blocks, DPX_THREADS, 0, (cudaStream_t)stream>>>(
            pk, mc_grad, mc_param, m, v, n4, rank, world, state, coef, beta1, beta2, eps, grad_scale, timeout_ns);
    else
        dp_adam_exchange_kernel<false><<<(unsigned)blocks, DPX_THREADS, 0, (cudaStream_t)stream>>>(
            pk, nullptr, nullptr, m, v, n4, rank, world, state, coef, beta1, beta2, eps, grad_scale, timeout_ns);
    MOPOE_CHECK_LAUNCH("dp_adam_exchange");
    return 0;
}

extern "C" int mopoe_dp_adam_exchange(const mopoe_dp_peers_t* peers, const float* mc_grad, float* mc_param, float* m, float* v,
                                      int64_t n, int rank, int world, uint32_t* state, const float* coef, float beta1,
                                      float beta2, float eps, float grad_scale, void* stream) {
    return mopoe_dp_adam_exchange_ex(peers, mc_grad, mc_param, m, v, n, rank, world, state, coef, beta1, beta2, eps, grad_scale, 0,
                                     stream);
}
