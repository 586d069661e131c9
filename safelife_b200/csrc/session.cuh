// session.cuh -- host-driven stepping without a launch per step.
//
// A host policy steps the batch once per host round trip (actions in, reward / done
// out: SafeLifeEnv.step as training/base_algo.py:231-236 drives it from the CPU).
// With one launch per step most of that round trip is launch and completion
// latency around a few microseconds of work.  A session keeps ONE
// resident kernel for many steps:
//   * every worker CTA loads its envs (1..8, one warp each, as k_env_step_warp) into
//     shared memory once and keeps them there until the session ends;
//   * the host posts a step by writing the actions into pinned memory, every byte
//     tagged with the step's sequence number in its high nibble (so that data and
//     "go" arrive in the same PCIe reads and no ordering between them is needed);
//   * one extra CTA, the courier, polls that buffer, and when every byte carries the
//     expected tag copies the actions into device memory and publishes the sequence
//     number there; the workers spin on that word in the L2, never on the host;
//   * the workers run the step (phase A per warp, the scalar tails on warp 0), write
//     the packed result records into pinned host memory (SlEnvBatch.host_out) and count
//     their arrivals in device memory; the last CTA to arrive issues the step's one
//     system-scope fence and writes the sequence number into the host's done word,
//     which the host spins on.
//     (Measured, tools/probes/pcie_probe.cu and tools/trace_session.py, 4096 envs: a
//     system fence costs 2 us alone and 5 us with 586 worker CTAs at it, hence one per
//     step; letting the host scan tags in the records instead of a done word costs it
//     27 ns per freshly written line, 14 us; carrying the records to the host through
//     the courier's SM alone takes 7 us; dependent PCIe reads cost 1.7 us each, so the
//     courier issues all the reads of a poll before it looks at any.)
// The courier ends the session when the host asks for it, or by itself when no step
// has been posted for idle_timeout_ms (a host that died must not leave a spinning
// kernel behind): the workers then write their envs back and the kernel exits; the
// host side notices (alive word) and starts a new session on the next step.
// All CTAs of the grid must be resident at once (the host checks occupancy).
#pragma once
#include "warp_kernels.cuh"

namespace sl {

constexpr uint32_t SESSION_EXIT = 0xffffffffu;
constexpr int SESSION_LINES = 256;             // 128-byte lines of published sequence numbers, one per SM
constexpr int SESSION_CTRL_WORDS = 32 + 32 * SESSION_LINES;

__device__ __forceinline__ unsigned int sm_id() {
    unsigned int v;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(v));
    return v;
}

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t *p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_gpu(uint32_t *p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_sys(volatile uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_volatile_u32(const volatile uint32_t *p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint4 ld_volatile_v4(const uint4 *p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Timeline of the last step (globaltimer ns) in dev_ctrl words 8..23, for
// tools/trace_session.py: 0 courier starts the poll that sees the step, 1 every tag
// seen, 2 sequence number published; worker CTA 0: 3 sees it, 4 phase A done (all
// warps), 5 tails done and records written; the last CTA to arrive: 6 arrived, 7 fenced and
// done word sent.
__device__ __forceinline__ void session_stamp(uint32_t *dev_ctrl, int slot) {
    ((unsigned long long *)(dev_ctrl + 8))[slot] = globaltimer_ns();
}

// Kernel-side copy of SlHostSession.
struct SessionParams {
    const uint8_t *host_actions;       // pinned [n_pad]: action | (seq & 15) << 4
    uint8_t *dev_actions;              // device [2][n_pad]
    uint32_t *dev_ctrl;                // device: [8..23] timeline; from word 32 on one 128-byte line per SM
                                       // with the published sequence number in its first word
    volatile uint32_t *host_ctrl;      // pinned: [0] done seq, [1] alive, [2] command (1 = end)
    int n_pad;                         // n_envs rounded up to 16
    unsigned long long idle_ns;
};

constexpr int SESSION_COURIER_CHUNKS = 4;      // 16-byte chunks a courier thread keeps in registers

// The courier CTA: host mailbox -> device mailbox.
__device__ __forceinline__ void session_courier(const SessionParams &S) {
    const int n16 = S.n_pad >> 4;
    const uint4 *src = (const uint4 *)S.host_actions;
    __shared__ int s_flag;
    uint32_t seq = 1;
    bool live = true;
    while (live) {
        const uint32_t tag = ((seq & 15u) << 4) * 0x01010101u;
        const unsigned long long t0 = globaltimer_ns();
        uint4 v[SESSION_COURIER_CHUNKS];
        for (;;) {
            bool ok = true;
            const unsigned long long t_poll = globaltimer_ns();
            // every PCIe read of the round is issued before the first result is looked at
            // (a read costs 1.7 us: three dependent ones were most of a 6.4 us poll)
            uint32_t cmd = 0;
            if (threadIdx.x == 0) cmd = ld_volatile_u32(S.host_ctrl + 2);
#pragma unroll
            for (int k = 0; k < SESSION_COURIER_CHUNKS; k++) {
                const int i = threadIdx.x + k * blockDim.x;
                v[k] = i < n16 ? ld_volatile_v4(src + i) : make_uint4(tag, tag, tag, tag);
            }
#pragma unroll
            for (int k = 0; k < SESSION_COURIER_CHUNKS; k++)
                ok &= ((v[k].x & 0xf0f0f0f0u) == tag) & ((v[k].y & 0xf0f0f0f0u) == tag) &
                      ((v[k].z & 0xf0f0f0f0u) == tag) & ((v[k].w & 0xf0f0f0f0u) == tag);
            for (int i = threadIdx.x + SESSION_COURIER_CHUNKS * blockDim.x; i < n16; i += blockDim.x) {
                const uint4 w = ld_volatile_v4(src + i);
                ok &= ((w.x & 0xf0f0f0f0u) == tag) & ((w.y & 0xf0f0f0f0u) == tag) &
                      ((w.z & 0xf0f0f0f0u) == tag) & ((w.w & 0xf0f0f0f0u) == tag);
            }
            int stop = 0;
            if (threadIdx.x == 0) stop = cmd != 0 || globaltimer_ns() - t0 > S.idle_ns;
            const int all_ok = __syncthreads_and(ok ? 1 : 0);
            if (threadIdx.x == 0) s_flag = stop;
            __syncthreads();
            // a posted step wins over a simultaneous request to end: it is run first
            if (all_ok) {
                if (threadIdx.x == 0) {
                    ((unsigned long long *)(S.dev_ctrl + 8))[0] = t_poll;
                    session_stamp(S.dev_ctrl, 1);
                }
                break;
            }
            if (s_flag) {
                live = false;
                break;
            }
        }
        if (!live) break;
        uint4 *dst = (uint4 *)(S.dev_actions + (size_t)(seq & 1u) * S.n_pad);
#pragma unroll
        for (int k = 0; k < SESSION_COURIER_CHUNKS; k++) {
            const int i = threadIdx.x + k * blockDim.x;
            if (i < n16)
                dst[i] = make_uint4(v[k].x & 0x0f0f0f0fu, v[k].y & 0x0f0f0f0fu, v[k].z & 0x0f0f0f0fu,
                                    v[k].w & 0x0f0f0f0fu);
        }
        for (int i = threadIdx.x + SESSION_COURIER_CHUNKS * blockDim.x; i < n16; i += blockDim.x) {
            const uint4 w = ld_volatile_v4(src + i);      // (tags already verified)
            dst[i] = make_uint4(w.x & 0x0f0f0f0fu, w.y & 0x0f0f0f0fu, w.z & 0x0f0f0f0fu, w.w & 0x0f0f0f0fu);
        }
        // one fence for the CTA's copies (thread 0, after the barrier has made them its
        // business), then one copy of the sequence number per SM: the workers of an SM spin
        // on their own line (586 CTAs on one line slow every access that shares its L2 slice)
        __syncthreads();
        if (threadIdx.x == 0) __threadfence();
        __syncthreads();
        for (int i = threadIdx.x; i < SESSION_LINES; i += blockDim.x)
            st_relaxed_gpu(S.dev_ctrl + 32 + 32 * i, seq);
        if (threadIdx.x == 0) session_stamp(S.dev_ctrl, 2);
        seq++;
    }
    for (int i = threadIdx.x; i < SESSION_LINES; i += blockDim.x)
        st_release_gpu(S.dev_ctrl + 32 + 32 * i, SESSION_EXIT);
    if (threadIdx.x == 0) st_release_sys(S.host_ctrl + 1, 0u);   // not alive: the host must begin a new session
}

// The session kernel: n_workers worker CTAs (epc envs each, like k_env_step_warp
// without observation / wrapper epilogue) and one courier CTA.
template <bool GR>
__global__ void __launch_bounds__(WARP_ENVS_PER_CTA * 32, 4)
k_env_session(const __grid_constant__ SlEnvBatch B, const __grid_constant__ SessionParams S,
              int flags, int epc, int n_workers) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if ((int)blockIdx.x == n_workers) {
        session_courier(S);
        return;
    }
    const int cells = B.H * B.W;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int e0 = blockIdx.x * epc;
    const int nenv = min(epc, B.n_envs - e0);
    const int group_bytes = nenv * cells * 2, state_bytes = nenv * 96;
    // the same placement as k_env_step_warp: slabs keep their offset inside a 16-byte unit
    const size_t byte0 = (size_t)e0 * cells * 2;
    const int mis = (int)(byte0 & 15);
    const int load_bytes = (mis + group_bytes + 15) & ~15;
    const bool bulk = group_bytes >= 64 &&
                      byte0 - mis + load_bytes <= (((size_t)B.n_envs * cells * 2) & ~(size_t)15);
    uint64_t *bar = (uint64_t *)smem_raw;
    int32_t *stable = (int32_t *)(smem_raw + WARP_SMEM_HEAD);
    SlEnvState *sstate = (SlEnvState *)(smem_raw + WARP_SMEM_HEAD + WARP_SMEM_TABLE);
    unsigned char *rboard = smem_raw + WARP_SMEM_HEAD + WARP_SMEM_TABLE + WARP_SMEM_STATE;
    unsigned char *rgoals = rboard + warp_slab_stride_any(cells);
    uint16_t *sboard = (uint16_t *)(rboard + (bulk ? mis : 0));
    uint16_t *sgoals = (uint16_t *)(rgoals + (bulk && GR ? mis : 0));
    uint16_t *gboard = B.board + (size_t)e0 * cells;
    uint16_t *ggoals = B.goals + (size_t)e0 * cells;
    SlEnvState *gstate = B.state + e0;
    volatile uint32_t *s_cmd = (volatile uint32_t *)(smem_raw + 32);

    CtaShared cs;
    cs.scratch = (uint16_t *)rgoals;
    cs.lock = (int *)(smem_raw + 16);
    if (bulk) {
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            *cs.lock = 0;
            mbar_expect_tx(bar, (uint32_t)((GR ? 2 : 1) * load_bytes + state_bytes));
            bulk_load(sstate, gstate, state_bytes, bar);
            bulk_load(rboard, (const unsigned char *)gboard - mis, load_bytes, bar);
            if (GR) bulk_load(rgoals, (const unsigned char *)ggoals - mis, load_bytes, bar);
        }
    } else {
        if (threadIdx.x == 0) *cs.lock = 0;
        copy_slabs(sstate, gstate, state_bytes);
        copy_slabs(sboard, gboard, group_bytes);
        if (GR) copy_slabs(sgoals, ggoals, group_bytes);
    }
    cs.table = stage_table(B, stable);
    __syncthreads();
    if (bulk) mbar_wait(bar, 0);

    const Narrow t(B.H, B.W, lane);
    const uint32_t *my_line = S.dev_ctrl + 32 + 32 * (sm_id() % SESSION_LINES);
    bool goals_dirty = false;       // warp 0: some step rewrote a shared-memory goals slab
    for (uint32_t seq = 1;; seq++) {
        if (threadIdx.x == 0) {
            uint32_t v;
            while ((v = ld_acquire_gpu(my_line)) != seq && v != SESSION_EXIT) __nanosleep(40);
            *s_cmd = v;
            if (blockIdx.x == 0) session_stamp(S.dev_ctrl, 3);
        }
        __syncthreads();
        if (*s_cmd == SESSION_EXIT) break;
        bool dirty = false;
        if (warp < nenv) {
            const int action = __ldcg(S.dev_actions + (size_t)(seq & 1u) * S.n_pad + e0 + warp);
            dirty = step_phase_a<Narrow, GR>(B, cs, t, sstate + warp, sboard + warp * cells,
                                             GR ? sgoals + warp * cells : ggoals + (size_t)warp * cells,
                                             nullptr, action);
        }
        const int any_dirty = __syncthreads_or(dirty);
        if (warp == 0) {
            const bool tr = blockIdx.x == 0 && lane == 0;
            if (tr) session_stamp(S.dev_ctrl, 4);
            goals_dirty |= any_dirty != 0;
            // every record carries the step's tag: valid for the host the moment it lands
            goals_dirty |= step_phase_b(B, cs, e0, nenv, sstate, sboard, GR ? sgoals : ggoals, flags, lane,
                                        seq);
            if (tr) session_stamp(S.dev_ctrl, 5);
            // Arrival.  The records of this CTA are ordered before its arrival at GPU scope;
            // the last CTA to arrive has thereby observed every other CTA's arrival, and ITS
            // system-scope fence orders all of their records before the done word (fences
            // are cumulative): one system fence per step, and the host spins on one word.
            __syncwarp();
            if (lane == 0) {
                __threadfence();
                if (atomicAdd(S.dev_ctrl + 1, 1u) == (uint32_t)n_workers - 1u) {
                    S.dev_ctrl[1] = 0;                      // everyone has arrived: ready for the next step
                    session_stamp(S.dev_ctrl, 6);
                    __threadfence_system();
                    st_release_sys(S.host_ctrl, seq);
                    session_stamp(S.dev_ctrl, 7);
                }
            }
        }
        // (the next round's barrier keeps the other warps behind warp 0's tails)
    }
    // the session is over: the envs go back to HBM
    fence_async_smem();
    __syncthreads();
    if (warp != 0) return;
    if (bulk) {
        const int head = (16 - mis) & 15;
        const int body = (group_bytes - head) & ~15, tail = group_bytes - head - body;
        if (lane == 0) {
            bulk_store((unsigned char *)gboard + head, (unsigned char *)sboard + head, body);
            if (GR && goals_dirty)
                bulk_store((unsigned char *)ggoals + head, (unsigned char *)sgoals + head, body);
            bulk_store(gstate, sstate, state_bytes);
        }
        const int edge = lane < 8 ? lane : (head + body) / 2 + lane - 8;
        if (lane < 16 && (lane < 8 ? 2 * lane < head : 2 * (lane - 8) < tail)) {
            gboard[edge] = sboard[edge];
            if (GR && goals_dirty) ggoals[edge] = sgoals[edge];
        }
        if (lane == 0) bulk_commit_and_wait_read();
    } else {
        for (int i = lane; i < (group_bytes >> 1); i += 32) gboard[i] = sboard[i];
        if (GR && goals_dirty)
            for (int i = lane; i < (group_bytes >> 1); i += 32) ggoals[i] = sgoals[i];
        for (int i = lane; i < (state_bytes >> 1); i += 32)
            ((uint16_t *)gstate)[i] = ((const uint16_t *)sstate)[i];
    }
}

}  // namespace sl
