// abi.cu -- extern "C" entry points of libsafelife_b200.so (include/safelife_b200.h).
//
// Argument checking mirrors the reference binding's error behaviour
// (safelife/speedups_src/module.c: ValueError for shape problems -> SL_ERR_ARG);
// nothing here throws or keeps simulation state between calls.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <time.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <emmintrin.h>

#include <mutex>
#include <vector>

#include "../../include/safelife_b200.h"
#include "cells.cuh"
#include "emd_host.h"
#include "generic_kernels.cuh"
#include "obs.cuh"
#include "wrap.cuh"
#include "warp_kernels.cuh"
#include "tile_kernels.cuh"
#include "session.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define SL_CUDA(call)                                                               \
    do {                                                                            \
        cudaError_t err__ = (call);                                                 \
        if (err__ != cudaSuccess)                                                   \
            return fail(err__ == cudaErrorNoDevice || err__ == cudaErrorInsufficientDriver \
                            ? SL_ERR_NO_DEVICE : SL_ERR_CUDA,                       \
                        "%s: %s", #call, cudaGetErrorString(err__));                \
    } while (0)

int launch_status(const char *what) {
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess)
        return fail(err == cudaErrorNoDevice || err == cudaErrorInsufficientDriver
                        ? SL_ERR_NO_DEVICE : SL_ERR_CUDA,
                    "%s launch: %s", what, cudaGetErrorString(err));
    return SL_OK;
}

// draw < (double)spawn_prob  <=>  (u64 >> 11) < ceil((double)spawn_prob * 2^53)
uint64_t spawn_threshold(float spawn_prob) {
    double t = (double)spawn_prob * 9007199254740992.0;
    if (!(t > 0.0)) return 0;
    if (t >= 9007199254740992.0) return 1ull << 53;
    return (uint64_t)ceil(t);
}

struct DeviceProps {
    int sm_count = 0, l2_bytes = 0, major = 0, minor = 0, max_smem_optin = 0, smem_per_sm = 0;
    bool ok = false;
};

int device_props(DeviceProps *out) {
    static thread_local int cached_dev = -1;
    static thread_local DeviceProps cached;
    int dev;
    SL_CUDA(cudaGetDevice(&dev));
    if (dev != cached_dev) {
        DeviceProps p;
        SL_CUDA(cudaDeviceGetAttribute(&p.sm_count, cudaDevAttrMultiProcessorCount, dev));
        SL_CUDA(cudaDeviceGetAttribute(&p.l2_bytes, cudaDevAttrL2CacheSize, dev));
        SL_CUDA(cudaDeviceGetAttribute(&p.major, cudaDevAttrComputeCapabilityMajor, dev));
        SL_CUDA(cudaDeviceGetAttribute(&p.minor, cudaDevAttrComputeCapabilityMinor, dev));
        SL_CUDA(cudaDeviceGetAttribute(&p.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        SL_CUDA(cudaDeviceGetAttribute(&p.smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
        p.ok = true;
        cached = p;
        cached_dev = dev;
    }
    *out = cached;
    return SL_OK;
}

// Opt a kernel in to `bytes` of dynamic shared memory.  The attribute is sticky
// per device, so it is set once per (kernel, device) and raised only when a
// later call needs more: a step in a host-driven loop must not pay a dozen
// driver calls.  (The only process-global state of the library; no simulation
// state.)
struct SmemGrant {
    const void *kernel;
    int dev, bytes;
};
std::mutex g_grant_mutex;
std::vector<SmemGrant> g_grants;

int prepare_smem_ptr(const void *kernel, size_t bytes);

template <typename K>
int prepare_smem(K kernel, size_t bytes) {
    return prepare_smem_ptr((const void *)kernel, bytes);
}

int prepare_smem_ptr(const void *kernel, size_t bytes) {
    DeviceProps p;
    int rc = device_props(&p);
    if (rc) return rc;
    if (bytes > (size_t)p.max_smem_optin)
        return fail(SL_ERR_ARG, "board too large: needs %zu bytes of shared memory (max %d)",
                    bytes, p.max_smem_optin);
    if (bytes <= 48 * 1024) return SL_OK;
    int dev = 0;
    SL_CUDA(cudaGetDevice(&dev));
    const void *key = kernel;
    std::lock_guard<std::mutex> lock(g_grant_mutex);
    for (SmemGrant &g : g_grants) {
        if (g.kernel == key && g.dev == dev) {
            if ((int)bytes <= g.bytes) return SL_OK;
            SL_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
            g.bytes = (int)bytes;
            return SL_OK;
        }
    }
    SL_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    g_grants.push_back({key, dev, (int)bytes});
    return SL_OK;
}

int check_board_args(const void *p, int n, int H, int W) {
    if (n < 0) return fail(SL_ERR_ARG, "n must be >= 0");
    if (H < 3 || W < 3) return fail(SL_ERR_ARG, "Board must be at least 3x3.");
    if ((long long)H * W > (1 << 24)) return fail(SL_ERR_ARG, "board too large");
    if (n > 0 && p == nullptr) return fail(SL_ERR_ARG, "null board pointer");
    return SL_OK;
}

int grid_for(int n) { return n < 1 ? 1 : (n > (1 << 20) ? (1 << 20) : n); }

// Which fast step path handles H x W boards: 1 = one warp per env (narrow, <= 32 x 32),
// 2 = several warps per env on 30 x 30 tiles, 3 = one warp per env on 64-bit rows
// (<= 64 x 64; superseded by the tiles, kept behind SAFELIFE_B200_WIDE=1 for A/B runs),
// 0 = none (generic CTA-per-env kernel).
int fast_step_path(int H, int W) {
    if (sl::warp_kernel_supports(H, W) && sl::warp_kernel_is_narrow(H, W)) return 1;
    static const bool wide = sl::env_knob("SAFELIFE_B200_WIDE", 0) != 0;
    if (wide && sl::warp_kernel_supports(H, W)) return 3;
    if (sl::tile_kernel_supports(H, W)) return 2;
    return 0;
}

// The multi-step kernels on tiles put TW warps on a board.  Measured on 50x50
// navigation boards against one warp per board on 64-bit rows (board-steps/s): 256
// boards 5.2e7 vs 2.0e7, 2048 boards 1.38e8 vs 1.02e8 -- tiles at every batch size
// (SAFELIFE_B200_WIDE=1 brings the one-warp kernel back for A/B runs).
bool multistep_on_tiles(int n, int H, int W) {
    (void)n;
    return fast_step_path(H, W) == 2;
}

int launch_multistep_tile(int mode, const uint16_t *in, const uint16_t *in2, uint16_t *out,
                          int32_t *counts, int n, int H, int W, uint64_t thr_all,
                          const uint64_t *thr_each, int n_steps, const int32_t *lead_steps,
                          uint64_t *rng, int sm_count, cudaStream_t st) {
    sl::TileParams sh = sl::tile_shape(H, W);
    // a small batch: one board per CTA, so that its boards reach as many SMs as possible
    if ((n + sh.epc - 1) / sh.epc < 2 * sm_count) sh.epc = 1;
    const size_t smem = sl::tile_multistep_smem_bytes(sh.epc, sh.TW, H * W);
    int rc = prepare_smem(sl::k_multistep_tile, smem);
    if (rc) return rc;
    int grid = (n + sh.epc - 1) / sh.epc;
    const int cap = sm_count * 4;
    if (grid > cap) grid = cap;
    sl::k_multistep_tile<<<grid, sh.epc * sh.TW * 32, smem, st>>>(
        sh, mode, in, in2, out, counts, n, H, W, thr_all, thr_each, n_steps, lead_steps, rng);
    return launch_status("k_multistep_tile");
}

int advance_impl(const uint16_t *in, uint16_t *out, int32_t *counts, int n, int H, int W,
                 float spawn_prob, int n_steps, uint64_t *rng, cudaStream_t st) {
    int rc = check_board_args(in, n, H, W);
    if (rc) return rc;
    if (n_steps < 1) return fail(SL_ERR_ARG, "n_steps must be >= 1");
    if (n == 0) return SL_OK;
    if (counts == nullptr && out == nullptr) return fail(SL_ERR_ARG, "null output pointer");
    if (sl::warp_advance_supports(H, W)) {
        DeviceProps props;
        rc = device_props(&props);
        if (rc) return rc;
        sl::launch_advance_warp(in, out, counts, n, H, W, spawn_threshold(spawn_prob), n_steps,
                                rng, props.sm_count, st);
        return launch_status("k_advance_warp");
    }
    if (multistep_on_tiles(n, H, W)) {        // beyond 32x32: several warps per board, planes in registers
        DeviceProps props;
        rc = device_props(&props);
        if (rc) return rc;
        return launch_multistep_tile(counts ? 1 : 0, in, nullptr, out, counts, n, H, W,
                                     spawn_threshold(spawn_prob), nullptr, n_steps, nullptr, rng,
                                     props.sm_count, st);
    }
    if (sl::warp_kernel_supports(H, W)) {      // up to 64x64 on one warp (SAFELIFE_B200_WIDE=1): two rows per lane
        DeviceProps props;
        rc = device_props(&props);
        if (rc) return rc;
        const bool counting = counts != nullptr;
        size_t smem = sl::multistep_wide_smem(H, W, counting);
        rc = prepare_smem(sl::k_multistep_wide, smem);
        if (rc) return rc;
        sl::k_multistep_wide<<<sl::multistep_wide_grid(n, props.sm_count), sl::WIDE_ADV_WARPS * 32, smem, st>>>(
            counting ? 1 : 0, in, nullptr, out, counts, n, H, W, spawn_threshold(spawn_prob), nullptr,
            n_steps, nullptr, rng);
        return launch_status("k_multistep_wide");
    }
    size_t smem = sl::generic_smem_bytes(H * W);
    rc = prepare_smem(sl::k_advance_generic, smem);
    if (rc) return rc;
    sl::k_advance_generic<<<grid_for(n), sl::GENERIC_THREADS, smem, st>>>(
        in, out, counts, n, H, W, spawn_threshold(spawn_prob), n_steps, rng);
    return launch_status("k_advance_generic");
}

// Scratch device buffer for the *_host entry points (per thread, grown on demand).
struct HostStage {
    void *ptr = nullptr;
    size_t cap = 0;
    int dev = -1;
    int reserve(size_t bytes) {
        int cur;
        SL_CUDA(cudaGetDevice(&cur));
        if (bytes <= cap && cur == dev) return SL_OK;
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
        SL_CUDA(cudaMalloc(&ptr, bytes));
        cap = bytes;
        dev = cur;
        return SL_OK;
    }
};
thread_local HostStage g_stage;

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace

extern "C" {

int sl_abi_version(void) { return SL_ABI_VERSION; }

const char *sl_last_error(void) { return g_err; }

int sl_device_info(int *sm_count, int *l2_bytes, int *cc_major, int *cc_minor) {
    DeviceProps p;
    int rc = device_props(&p);
    if (rc) return rc;
    if (sm_count) *sm_count = p.sm_count;
    if (l2_bytes) *l2_bytes = p.l2_bytes;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    return SL_OK;
}

#ifdef SL_TRACE
/* diagnostic builds only (tools/trace_step.py): device buffer uint64[grid][8] or NULL */
int sl_debug_trace(void *buf) {
    SL_CUDA(cudaMemcpyToSymbol(sl::sl_trace_buf, &buf, sizeof buf));
    return SL_OK;
}
#endif

int sl_warp_kernel_supported(int H, int W) { return fast_step_path(H, W) != 0 ? 1 : 0; }

int sl_advance_board(const uint16_t *in, uint16_t *out, int n, int H, int W, float spawn_prob,
                     int n_steps, uint64_t *rng, void *stream) {
    return advance_impl(in, out, nullptr, n, H, W, spawn_prob, n_steps, rng, (cudaStream_t)stream);
}

int sl_life_occupancy(const uint16_t *board, int32_t *counts, int n, int H, int W,
                      float spawn_prob, int n_steps, uint64_t *rng, void *stream) {
    if (n > 0 && counts == nullptr) return fail(SL_ERR_ARG, "null counts pointer");
    return advance_impl(board, nullptr, counts, n, H, W, spawn_prob, n_steps, rng,
                        (cudaStream_t)stream);
}

int sl_side_effect_occupancy(const uint16_t *init_board, const uint16_t *final_board,
                             int32_t *counts, int n, int H, int W, const uint64_t *spawn_thr,
                             const int32_t *num_steps, int num_samples, uint64_t *rng,
                             void *stream) {
    int rc = check_board_args(init_board, n, H, W);
    if (rc) return rc;
    if (num_samples < 1) return fail(SL_ERR_ARG, "num_samples must be >= 1");
    if (n == 0) return SL_OK;
    if (!final_board || !counts || !spawn_thr || !num_steps) return fail(SL_ERR_ARG, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (sl::warp_advance_supports(H, W)) {
        DeviceProps props;
        rc = device_props(&props);
        if (rc) return rc;
        sl::launch_side_effect_warp(init_board, final_board, counts, n, H, W, spawn_thr, num_steps,
                                    num_samples, rng, props.sm_count, st);
        return launch_status("k_side_effect_warp");
    }
    if (multistep_on_tiles(n, H, W)) {
        DeviceProps props;
        rc = device_props(&props);
        if (rc) return rc;
        return launch_multistep_tile(2, init_board, final_board, nullptr, counts, n, H, W, 0, spawn_thr,
                                     num_samples, num_steps, rng, props.sm_count, st);
    }
    if (sl::warp_kernel_supports(H, W)) {
        DeviceProps props;
        rc = device_props(&props);
        if (rc) return rc;
        size_t wsmem = sl::multistep_wide_smem(H, W, true);
        rc = prepare_smem(sl::k_multistep_wide, wsmem);
        if (rc) return rc;
        sl::k_multistep_wide<<<sl::multistep_wide_grid(n, props.sm_count), sl::WIDE_ADV_WARPS * 32, wsmem, st>>>(
            2, init_board, final_board, nullptr, counts, n, H, W, 0, spawn_thr, num_samples, num_steps, rng);
        return launch_status("k_multistep_wide");
    }
    size_t smem = sl::generic_smem_bytes(H * W);
    rc = prepare_smem(sl::k_side_effect_generic, smem);
    if (rc) return rc;
    sl::k_side_effect_generic<<<grid_for(n), sl::GENERIC_THREADS, smem, st>>>(
        init_board, final_board, counts, n, H, W, spawn_thr, num_steps, num_samples, rng);
    return launch_status("k_side_effect_generic");
}

int sl_alive_counts(const uint16_t *board, const uint16_t *goals, int64_t *out, int n, int cells,
                    void *stream) {
    if (n < 0 || cells < 0) return fail(SL_ERR_ARG, "negative size");
    if (n == 0) return SL_OK;
    if (!board || !goals || !out) return fail(SL_ERR_ARG, "null pointer");
    sl::k_alive_counts<<<grid_for(n), sl::GENERIC_THREADS, 0, (cudaStream_t)stream>>>(
        board, goals, out, n, cells);
    return launch_status("k_alive_counts");
}

int sl_execute_actions(uint16_t *board, int64_t *locations, const int64_t *actions, int n, int H,
                       int W, int n_agents, int action_stride, void *stream) {
    int rc = check_board_args(board, n, H, W);
    if (rc) return rc;
    if (n_agents < 0) return fail(SL_ERR_ARG, "Locations should be shape (n_agent, 2).");
    if (action_stride != 0 && action_stride != 1)
        return fail(SL_ERR_ARG, "action_stride must be 0 (broadcast) or 1");
    if (n == 0 || n_agents == 0) return SL_OK;
    if (!locations || !actions) return fail(SL_ERR_ARG, "null pointer");
    sl::k_execute_actions<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        board, locations, actions, n, H, W, n_agents, action_stride);
    return launch_status("k_execute_actions");
}

int sl_wrapped_label(int32_t *data, int32_t *num_labels, int n, int H, int W, void *stream) {
    if (n < 0) return fail(SL_ERR_ARG, "n must be >= 0");
    if (H < 1 || W < 1) return fail(SL_ERR_ARG, "grid must be non-empty");
    if (n == 0) return SL_OK;
    if (!data || !num_labels) return fail(SL_ERR_ARG, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    int32_t *scratch = nullptr;
    SL_CUDA(cudaMallocAsync((void **)&scratch, (size_t)n * H * W * 3 * sizeof(int32_t), st));
    sl::k_wrapped_label<<<(n + 31) / 32, 32, 0, st>>>(data, num_labels, scratch, n, H, W);
    int rc = launch_status("k_wrapped_label");
    cudaFreeAsync(scratch, st);
    return rc;
}

// ---- host-buffer variants ------------------------------------------------

int sl_advance_board_host(const uint16_t *in, uint16_t *out, int n, int H, int W, float spawn_prob,
                          int n_steps, uint64_t *rng) {
    int rc = check_board_args(in, n, H, W);
    if (rc) return rc;
    if (n_steps < 1) return fail(SL_ERR_ARG, "n_steps must be >= 1");
    if (n == 0) return SL_OK;
    if (!out) return fail(SL_ERR_ARG, "null output pointer");
    size_t bbytes = (size_t)n * H * W * 2, rbytes = rng ? (size_t)n * 32 : 0;
    rc = g_stage.reserve(align256(bbytes) + align256(rbytes) + 256);
    if (rc) return rc;
    uint16_t *d_board = (uint16_t *)g_stage.ptr;
    uint64_t *d_rng = rng ? (uint64_t *)((char *)g_stage.ptr + align256(bbytes)) : nullptr;
    SL_CUDA(cudaMemcpyAsync(d_board, in, bbytes, cudaMemcpyHostToDevice, 0));
    if (rng) SL_CUDA(cudaMemcpyAsync(d_rng, rng, rbytes, cudaMemcpyHostToDevice, 0));
    rc = advance_impl(d_board, d_board, nullptr, n, H, W, spawn_prob, n_steps, d_rng, 0);
    if (rc) return rc;
    SL_CUDA(cudaMemcpyAsync(out, d_board, bbytes, cudaMemcpyDeviceToHost, 0));
    if (rng) SL_CUDA(cudaMemcpyAsync(rng, d_rng, rbytes, cudaMemcpyDeviceToHost, 0));
    SL_CUDA(cudaStreamSynchronize(0));
    return SL_OK;
}

int sl_life_occupancy_host(const uint16_t *board, int32_t *counts, int n, int H, int W,
                           float spawn_prob, int n_steps, uint64_t *rng) {
    int rc = check_board_args(board, n, H, W);
    if (rc) return rc;
    if (n_steps < 1) return fail(SL_ERR_ARG, "n_steps must be >= 1");
    if (n == 0) return SL_OK;
    if (!counts) return fail(SL_ERR_ARG, "null counts pointer");
    size_t bbytes = (size_t)n * H * W * 2, cbytes = (size_t)n * H * W * 8 * 4,
           rbytes = rng ? (size_t)n * 32 : 0;
    rc = g_stage.reserve(align256(bbytes) + align256(cbytes) + align256(rbytes) + 256);
    if (rc) return rc;
    char *base = (char *)g_stage.ptr;
    uint16_t *d_board = (uint16_t *)base;
    int32_t *d_counts = (int32_t *)(base + align256(bbytes));
    uint64_t *d_rng = rng ? (uint64_t *)(base + align256(bbytes) + align256(cbytes)) : nullptr;
    SL_CUDA(cudaMemcpyAsync(d_board, board, bbytes, cudaMemcpyHostToDevice, 0));
    SL_CUDA(cudaMemcpyAsync(d_counts, counts, cbytes, cudaMemcpyHostToDevice, 0));
    if (rng) SL_CUDA(cudaMemcpyAsync(d_rng, rng, rbytes, cudaMemcpyHostToDevice, 0));
    rc = advance_impl(d_board, nullptr, d_counts, n, H, W, spawn_prob, n_steps, d_rng, 0);
    if (rc) return rc;
    SL_CUDA(cudaMemcpyAsync(counts, d_counts, cbytes, cudaMemcpyDeviceToHost, 0));
    if (rng) SL_CUDA(cudaMemcpyAsync(rng, d_rng, rbytes, cudaMemcpyDeviceToHost, 0));
    SL_CUDA(cudaStreamSynchronize(0));
    return SL_OK;
}

int sl_alive_counts_host(const uint16_t *board, const uint16_t *goals, int64_t *out, int n,
                         int cells) {
    if (n < 0 || cells < 0) return fail(SL_ERR_ARG, "negative size");
    if (n == 0) return SL_OK;
    if (!board || !goals || !out) return fail(SL_ERR_ARG, "null pointer");
    size_t bbytes = (size_t)n * cells * 2, obytes = (size_t)n * 72 * 8;
    int rc = g_stage.reserve(2 * align256(bbytes) + align256(obytes) + 256);
    if (rc) return rc;
    char *base = (char *)g_stage.ptr;
    uint16_t *d_board = (uint16_t *)base, *d_goals = (uint16_t *)(base + align256(bbytes));
    int64_t *d_out = (int64_t *)(base + 2 * align256(bbytes));
    SL_CUDA(cudaMemcpyAsync(d_board, board, bbytes, cudaMemcpyHostToDevice, 0));
    SL_CUDA(cudaMemcpyAsync(d_goals, goals, bbytes, cudaMemcpyHostToDevice, 0));
    rc = sl_alive_counts(d_board, d_goals, d_out, n, cells, nullptr);
    if (rc) return rc;
    SL_CUDA(cudaMemcpyAsync(out, d_out, obytes, cudaMemcpyDeviceToHost, 0));
    SL_CUDA(cudaStreamSynchronize(0));
    return SL_OK;
}

int sl_execute_actions_host(uint16_t *board, int64_t *locations, const int64_t *actions, int n,
                            int H, int W, int n_agents, int action_stride) {
    int rc = check_board_args(board, n, H, W);
    if (rc) return rc;
    if (n_agents < 0) return fail(SL_ERR_ARG, "Locations should be shape (n_agent, 2).");
    if (action_stride != 0 && action_stride != 1)
        return fail(SL_ERR_ARG, "action_stride must be 0 (broadcast) or 1");
    if (n == 0 || n_agents == 0) return SL_OK;
    if (!locations || !actions) return fail(SL_ERR_ARG, "null pointer");
    size_t bbytes = (size_t)n * H * W * 2, lbytes = (size_t)n * n_agents * 16,
           abytes = (size_t)n * (action_stride ? n_agents : 1) * 8;
    rc = g_stage.reserve(align256(bbytes) + align256(lbytes) + align256(abytes) + 256);
    if (rc) return rc;
    char *base = (char *)g_stage.ptr;
    uint16_t *d_board = (uint16_t *)base;
    int64_t *d_loc = (int64_t *)(base + align256(bbytes));
    int64_t *d_act = (int64_t *)(base + align256(bbytes) + align256(lbytes));
    SL_CUDA(cudaMemcpyAsync(d_board, board, bbytes, cudaMemcpyHostToDevice, 0));
    SL_CUDA(cudaMemcpyAsync(d_loc, locations, lbytes, cudaMemcpyHostToDevice, 0));
    SL_CUDA(cudaMemcpyAsync(d_act, actions, abytes, cudaMemcpyHostToDevice, 0));
    rc = sl_execute_actions(d_board, d_loc, d_act, n, H, W, n_agents, action_stride, nullptr);
    if (rc) return rc;
    SL_CUDA(cudaMemcpyAsync(board, d_board, bbytes, cudaMemcpyDeviceToHost, 0));
    SL_CUDA(cudaMemcpyAsync(locations, d_loc, lbytes, cudaMemcpyDeviceToHost, 0));
    SL_CUDA(cudaStreamSynchronize(0));
    return SL_OK;
}

int sl_wrapped_label_host(int32_t *data, int32_t *num_labels, int n, int H, int W) {
    if (n < 0) return fail(SL_ERR_ARG, "n must be >= 0");
    if (H < 1 || W < 1) return fail(SL_ERR_ARG, "grid must be non-empty");
    if (n == 0) return SL_OK;
    if (!data || !num_labels) return fail(SL_ERR_ARG, "null pointer");
    size_t dbytes = (size_t)n * H * W * 4, nbytes = (size_t)n * 4;
    int rc = g_stage.reserve(align256(dbytes) + align256(nbytes) + 256);
    if (rc) return rc;
    char *base = (char *)g_stage.ptr;
    int32_t *d_data = (int32_t *)base, *d_num = (int32_t *)(base + align256(dbytes));
    SL_CUDA(cudaMemcpyAsync(d_data, data, dbytes, cudaMemcpyHostToDevice, 0));
    rc = sl_wrapped_label(d_data, d_num, n, H, W, nullptr);
    if (rc) return rc;
    SL_CUDA(cudaMemcpyAsync(data, d_data, dbytes, cudaMemcpyDeviceToHost, 0));
    SL_CUDA(cudaMemcpyAsync(num_labels, d_num, nbytes, cudaMemcpyDeviceToHost, 0));
    SL_CUDA(cudaStreamSynchronize(0));
    return SL_OK;
}

int sl_emd_host(const double *a, const double *b, const double *dist, int k,
                double extra_mass_penalty, double *out) {
    if (k < 0) return fail(SL_ERR_ARG, "k must be >= 0");
    if (!out || (k > 0 && (!a || !b || !dist))) return fail(SL_ERR_ARG, "null pointer");
    for (int i = 0; i < k; i++)
        if (!(a[i] >= 0.0) || !(b[i] >= 0.0)) return fail(SL_ERR_ARG, "histograms must be >= 0");
    *out = k == 0 ? 0.0 : sl::emd_hat(a, b, dist, k, extra_mass_penalty);
    return SL_OK;
}

int sl_emd_host_batch(const double *const *a, const double *const *b, const double *const *dist,
                      const int *k, int n, double extra_mass_penalty, double *out, int n_threads) {
    if (n < 0) return fail(SL_ERR_ARG, "n must be >= 0");
    if (n == 0) return SL_OK;
    if (!a || !b || !dist || !k || !out) return fail(SL_ERR_ARG, "null pointer");
    for (int i = 0; i < n; i++) {
        if (k[i] < 0) return fail(SL_ERR_ARG, "k must be >= 0");
        if (k[i] > 0 && (!a[i] || !b[i] || !dist[i])) return fail(SL_ERR_ARG, "null pointer");
        for (int j = 0; j < k[i]; j++)
            if (!(a[i][j] >= 0.0) || !(b[i][j] >= 0.0)) return fail(SL_ERR_ARG, "histograms must be >= 0");
    }
    sl::emd_hat_batch(a, b, dist, k, n, extra_mass_penalty, out, n_threads);
    return SL_OK;
}

// ---- batched environment ---------------------------------------------------

static int check_batch(const SlEnvBatch *b) {
    if (!b) return fail(SL_ERR_ARG, "null batch");
    int rc = check_board_args(b->board, b->n_envs, b->H, b->W);
    if (rc) return rc;
    if (b->n_levels < 1 || b->max_exits < 0) return fail(SL_ERR_ARG, "bad level pool sizes");
    static_assert(sizeof(SlEnvState) == 96, "SlEnvState must stay 96 bytes");
    if (((uintptr_t)b->state & 15) != 0) return fail(SL_ERR_ARG, "state must be 16-byte aligned");
    if (b->n_envs > 0 &&
        (!b->goals || !b->state || !b->reward ||
         !b->done || !b->success || !b->times_up || !b->pool_board || !b->pool_goals ||
         !b->pool_loc || !b->pool_table || !b->pool_init_points || !b->pool_required ||
         !b->pool_spawn_thr || (b->max_exits > 0 && !b->pool_exits)))
        return fail(SL_ERR_ARG, "null pointer in SlEnvBatch");
    return SL_OK;
}

static int check_obs_spec(const SlObsSpec *o) {
    if (o->vh < 1 || o->vw < 1 || o->n_channels < 0) return fail(SL_ERR_ARG, "bad view shape");
    if (o->n_channels > 32) return fail(SL_ERR_ARG, "at most 32 output channels (bits of the uint32 view)");
    if ((long long)o->vh * o->vw > (1 << 20)) return fail(SL_ERR_ARG, "view too large");
    if (!o->obs) return fail(SL_ERR_ARG, "null observation pointer");
    for (int k = 0; k < o->n_channels; k++)
        if (o->channels[k] < 0 || o->channels[k] > 31)
            return fail(SL_ERR_ARG, "output channels are bit positions 0..31");
    return SL_OK;
}

static int launch_get_obs(const SlEnvBatch *batch, const sl::ObsParams &O, cudaStream_t st) {
    DeviceProps props;
    int rc = device_props(&props);
    if (rc) return rc;
    const size_t smem = (size_t)sl::OBS_WARPS * sl::obs_stage_bytes(O.vh, O.vw, O.n_channels);
    rc = prepare_smem(sl::k_get_obs_warp, smem);
    if (rc) return rc;
    int grid = (batch->n_envs + sl::OBS_WARPS - 1) / sl::OBS_WARPS;
    const int cap = props.sm_count * 32;
    if (grid > cap) grid = cap;
    sl::k_get_obs_warp<<<grid, sl::OBS_WARPS * 32, smem, st>>>(*batch, O);
    return launch_status("k_get_obs_warp");
}

static int check_wrap_spec(const SlEnvBatch *b, const SlWrapSpec *w) {
    const int known = SL_WRAP_MOVEMENT | SL_WRAP_MOVEMENT_PENALTY | SL_WRAP_EXIT_BONUS |
                      SL_WRAP_SIDE_EFFECT | SL_WRAP_IGNORE_REWARD_CELLS;
    if (w->flags & ~known) return fail(SL_ERR_ARG, "unknown wrapper flags");
    if (!w->wstate || !w->reward) return fail(SL_ERR_ARG, "null pointer in SlWrapSpec");
    if (((uintptr_t)w->wstate & 15) != 0) return fail(SL_ERR_ARG, "wstate must be 16-byte aligned");
    static_assert(sizeof(SlWrapState) == 48, "SlWrapState must stay 48 bytes");
    if (w->flags & SL_WRAP_MOVEMENT) {
        if (w->movement_period < 1 || w->movement_period > 8)
            return fail(SL_ERR_ARG, "movement_period must be 1..8");
        if (!w->movement_table) return fail(SL_ERR_ARG, "null movement_table");
    }
    if ((w->flags & SL_WRAP_SIDE_EFFECT) && !w->baseline) return fail(SL_ERR_ARG, "null baseline");
    (void)b;
    return SL_OK;
}

static int wrap_grid(const SlEnvBatch *batch) {
    int grid = (batch->n_envs + 7) / 8;
    return grid > 4096 ? 4096 : grid;
}

static int env_step_impl(const SlEnvBatch *batch, const uint8_t *actions, const SlObsSpec *obs,
                         const SlWrapSpec *wrap, int kernel, int flags, void *stream) {
    int rc = check_batch(batch);
    if (rc) return rc;
    if (obs) {
        rc = check_obs_spec(obs);
        if (rc) return rc;
    }
    if (wrap) {
        rc = check_wrap_spec(batch, wrap);
        if (rc) return rc;
    }
    if (batch->n_envs == 0) return SL_OK;
    if (!actions) return fail(SL_ERR_ARG, "null actions");
    if (kernel < 0 || kernel > 3) return fail(SL_ERR_ARG, "kernel must be 0, 1, 2 or 3");
    cudaStream_t st = (cudaStream_t)stream;
    const bool prefer_pipe = kernel == 3;
    if (kernel == 3) kernel = 2;
    const int path = fast_step_path(batch->H, batch->W);
    bool warp_ok = path != 0;
    if (kernel == 2 && !warp_ok)
        return fail(SL_ERR_ARG, "no warp kernel for %d x %d boards", batch->H, batch->W);
    if (flags & ~SL_STEP_AUTO_RESET) return fail(SL_ERR_ARG, "unknown step flags");
    sl::ObsParams O;
    if (obs) O = sl::make_obs_params(*obs);
    if ((kernel == 2 || kernel == 0) && path == 2) {
        static const sl::ObsParams no_obs = {};
        static const SlWrapSpec no_wrap = {};
        const sl::TileShape sh = sl::tile_shape(batch->H, batch->W);
        const int stage = obs ? sl::obs_stage_bytes(O.vh, O.vw, O.n_channels) : 0;
        const size_t smem = sl::tile_step_smem_bytes(sh.epc, batch->H * batch->W, stage);
        rc = prepare_smem(sl::k_env_step_tile, smem);
        if (rc) return rc;
        const int bulk_ok = (((uintptr_t)batch->board | (uintptr_t)batch->goals | (uintptr_t)batch->state) & 15) == 0;
        sl::launch_pdl(sl::k_env_step_tile, (batch->n_envs + sh.epc - 1) / sh.epc, sh.epc * sh.TW * 32,
                       smem, st, *batch, obs ? O : no_obs, wrap ? *wrap : no_wrap, sh, actions, bulk_ok, flags);
        return launch_status("k_env_step_tile");
    }
    if (kernel == 2 || (kernel == 0 && warp_ok)) {
        DeviceProps props;
        rc = device_props(&props);
        if (rc) return rc;
        const sl::StepPlan p = sl::plan_env_step_warp(*batch, obs ? &O : nullptr, wrap,
                                                      props.sm_count, props.smem_per_sm, prefer_pipe);
        rc = sl::launch_env_step_warp(*batch, obs ? &O : nullptr, wrap, actions, flags, p, st,
                                      [](const void *fn, size_t smem) { return prepare_smem_ptr(fn, smem); });
        if (rc) return rc;
        return launch_status("k_env_step_warp");
    }
    // generic path: the same stages as separate launches
    size_t smem = sl::generic_smem_bytes(batch->H * batch->W);
    rc = prepare_smem(sl::k_env_step_generic, smem);
    if (rc) return rc;
    sl::k_env_step_generic<<<grid_for(batch->n_envs), sl::GENERIC_THREADS, smem, st>>>(*batch, actions);
    rc = launch_status("k_env_step_generic");
    if (rc) return rc;
    if (wrap) {
        sl::k_wrap_step<<<wrap_grid(batch), 256, 0, st>>>(*batch, *wrap);
        rc = launch_status("k_wrap_step");
        if (rc) return rc;
    }
    if (flags & SL_STEP_AUTO_RESET) {
        rc = sl_env_reset(batch, batch->done, nullptr, SL_RESET_NEXT_LEVEL | SL_RESET_KEEP_STATS, stream);
        if (rc) return rc;
        if (wrap) {
            rc = sl_wrap_reset(batch, wrap, batch->done, stream);
            if (rc) return rc;
        }
    }
    return obs ? launch_get_obs(batch, O, st) : SL_OK;
}

int sl_env_step(const SlEnvBatch *batch, const uint8_t *actions, int kernel, int flags,
                void *stream) {
    return env_step_impl(batch, actions, nullptr, nullptr, kernel, flags, stream);
}

int sl_env_step_obs(const SlEnvBatch *batch, const uint8_t *actions, const SlObsSpec *obs,
                    int kernel, int flags, void *stream) {
    return env_step_impl(batch, actions, obs, nullptr, kernel, flags, stream);
}

int sl_env_step_wrapped(const SlEnvBatch *batch, const uint8_t *actions, const SlObsSpec *obs,
                        const SlWrapSpec *wrap, int kernel, int flags, void *stream) {
    if (!wrap) return fail(SL_ERR_ARG, "null wrapper spec");
    return env_step_impl(batch, actions, obs, wrap, kernel, flags, stream);
}

int sl_wrap_reset(const SlEnvBatch *batch, const SlWrapSpec *wrap, const uint8_t *mask,
                  void *stream) {
    int rc = check_batch(batch);
    if (rc) return rc;
    if (!wrap) return fail(SL_ERR_ARG, "null wrapper spec");
    rc = check_wrap_spec(batch, wrap);
    if (rc) return rc;
    if (batch->n_envs == 0) return SL_OK;
    sl::k_wrap_reset<<<wrap_grid(batch), 256, 0, (cudaStream_t)stream>>>(*batch, *wrap, mask);
    return launch_status("k_wrap_reset");
}

static int env_step_plan_impl(const SlEnvBatch *batch, const SlObsSpec *obs, int kernel, char *text,
                              int text_len, int *n_launches) {
    int rc = check_batch(batch);
    if (rc) return rc;
    if (obs) {
        rc = check_obs_spec(obs);
        if (rc) return rc;
    }
    if (kernel < 0 || kernel > 3) return fail(SL_ERR_ARG, "kernel must be 0, 1, 2 or 3");
    char buf[320];
    int launches = 0;
    const bool prefer_pipe = kernel == 3;
    if (kernel == 3) kernel = 2;
    const int path = fast_step_path(batch->H, batch->W);
    const bool warp_ok = path != 0;
    if (batch->n_envs == 0) {
        snprintf(buf, sizeof buf, "nothing to launch");
    } else if ((kernel == 2 || kernel == 0) && path == 2) {
        const sl::TileShape sh = sl::tile_shape(batch->H, batch->W);
        snprintf(buf, sizeof buf, "k_env_step_tile<%d x %d tiles, %d envs/CTA%s> grid %d", sh.TR, sh.TC,
                 sh.epc, obs ? ", obs fused" : "", (batch->n_envs + sh.epc - 1) / sh.epc);
        launches = 1;
    } else if (kernel == 2 || (kernel == 0 && warp_ok)) {
        if (!warp_ok) return fail(SL_ERR_ARG, "no warp kernel for %d x %d boards", batch->H, batch->W);
        DeviceProps props;
        rc = device_props(&props);
        if (rc) return rc;
        sl::ObsParams O;
        if (obs) O = sl::make_obs_params(*obs);
        const sl::StepPlan p = sl::plan_env_step_warp(*batch, obs ? &O : nullptr, nullptr,
                                                      props.sm_count, props.smem_per_sm, prefer_pipe);
        const char *goals = p.goals_resident ? "goals resident" : "goals in place";
        const char *ob = p.obs ? ", obs fused" : "";
        int n = 0;
        if (p.piped_groups > 0) {
            n += snprintf(buf + n, sizeof buf - n, "k_env_step_pipe<%d stages, %d CTAs/SM, %s%s> grid %d",
                          p.pipe_stages, p.pipe_ctas, goals, ob, p.pipe_grid);
            launches++;
        }
        if (p.single_groups > 0) {
            n += snprintf(buf + n, sizeof buf - n, "%sk_env_step_warp<%s, %s%s> grid %d x %d envs",
                          n ? " + " : "", p.narrow ? "Narrow" : "Wide", goals, ob, p.single_groups,
                          p.single_epc);
            launches++;
        }
    } else {
        snprintf(buf, sizeof buf, "k_env_step_generic grid %d (+ k_env_reset with SL_STEP_AUTO_RESET)%s",
                 grid_for(batch->n_envs), obs ? " + k_get_obs_warp" : "");
        launches = obs ? 2 : 1;
    }
    if (text && text_len > 0) snprintf(text, (size_t)text_len, "%s", buf);
    if (n_launches) *n_launches = launches;
    return SL_OK;
}

int sl_env_step_plan(const SlEnvBatch *batch, int kernel, char *text, int text_len,
                     int *n_launches) {
    return env_step_plan_impl(batch, nullptr, kernel, text, text_len, n_launches);
}

int sl_env_step_obs_plan(const SlEnvBatch *batch, const SlObsSpec *obs, int kernel, char *text,
                         int text_len, int *n_launches) {
    return env_step_plan_impl(batch, obs, kernel, text, text_len, n_launches);
}

namespace {
__global__ void k_launch_floor(const __grid_constant__ SlEnvBatch B, int *sink) {
    if (sink != nullptr && B.n_envs < 0) *sink = B.H;     // never true: keeps the parameter alive
}
}  // namespace

int sl_launch_floor(const SlEnvBatch *batch, void *stream) {
    int rc = check_batch(batch);
    if (rc) return rc;
    int groups = batch->n_envs > 0 ? (batch->n_envs + 7) / 8 : 1, threads = 256;
    if (batch->n_envs > 0 && fast_step_path(batch->H, batch->W) == 1) {     // the single-pass kernel's shape
        DeviceProps props;
        rc = device_props(&props);
        if (rc) return rc;
        const sl::StepPlan p = sl::plan_env_step_warp(*batch, nullptr, nullptr, props.sm_count, props.smem_per_sm);
        if (p.single_groups > 0 && p.piped_groups == 0) {
            groups = p.single_groups;
            threads = p.single_epc * 32;
        }
    }
    k_launch_floor<<<groups, threads, 0, (cudaStream_t)stream>>>(*batch, nullptr);
    return launch_status("k_launch_floor");
}

namespace {
// Completion flag for host-driven loops: after the step's launch a stream memory
// operation writes a sequence number into pinned host memory
// (cuStreamWriteValue32, resolved at run time: the library does not link libcuda),
// and the host spins on that word instead of going through cudaStreamSynchronize.
typedef CUresult (*WriteValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
struct HostFlag {
    volatile uint32_t *word = nullptr;
    uint32_t seq = 0;
    int dev = -1;
    WriteValue32Fn write = nullptr;
    bool tried = false;

    bool ready() {
        int cur = -1;
        if (cudaGetDevice(&cur) != cudaSuccess) return false;
        if (tried && cur == dev) return write != nullptr;
        tried = true;
        dev = cur;
        write = nullptr;

        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &fn, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess || fn == nullptr) {
            cudaGetLastError();
            return false;
        }
        if (word == nullptr) {
            void *p = nullptr;
            if (cudaHostAlloc(&p, 64, cudaHostAllocPortable) != cudaSuccess) {
                cudaGetLastError();
                return false;
            }
            word = (volatile uint32_t *)p;
            *word = 0;
        }
        write = (WriteValue32Fn)fn;
        return true;
    }
};
thread_local HostFlag g_host_flag;

double now_us() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}

int wait_for_stream(cudaStream_t st) {
    // Measured (tools/host_latency.py, 4096 envs): cudaStreamSynchronize returns 1.2 us
    // sooner than the completion word shows up (16.1 vs 17.3 us per step), so the
    // synchronous call uses it; SAFELIFE_B200_HOST_WAIT=1 selects the word, which the
    // submit / wait pair needs anyway.
    static const bool use_word = sl::env_knob("SAFELIFE_B200_HOST_WAIT", 0) != 0;
    HostFlag &f = g_host_flag;
    if (use_word && f.ready()) {
        const uint32_t seq = ++f.seq;
        if (f.write((CUstream)st, (CUdeviceptr)(uintptr_t)f.word, seq, 0) == CUDA_SUCCESS) {
            // the kernel's stores into pinned memory are ordered before the flag
            // (stream memory operations fence by default)
            unsigned spins = 0;
            while (*f.word != seq) {
                __builtin_ia32_pause();
                if ((++spins & 0xfffffu) == 0 && cudaStreamQuery(st) != cudaErrorNotReady) break;
            }
            if (*f.word == seq) return SL_OK;
        }
    }
    SL_CUDA(cudaStreamSynchronize(st));
    return SL_OK;
}
}  // namespace

int sl_env_step_submit(const SlEnvBatch *batch, const uint8_t *actions, int kernel, int flags,
                       void *stream, uint32_t *flag, uint32_t seq) {
    if (!flag) return fail(SL_ERR_ARG, "null completion word");
    int rc = sl_env_step(batch, actions, kernel, flags, stream);
    if (rc) return rc;
    HostFlag &f = g_host_flag;
    if (f.ready() && f.write((CUstream)stream, (CUdeviceptr)(uintptr_t)flag, seq, 0) == CUDA_SUCCESS)
        return SL_OK;
    // no stream memory operations: complete synchronously and publish the word ourselves
    SL_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    *(volatile uint32_t *)flag = seq;
    return SL_OK;
}

int sl_host_wait(const uint32_t *flag, uint32_t seq, void *stream) {
    if (!flag) return fail(SL_ERR_ARG, "null completion word");
    const volatile uint32_t *word = flag;
    unsigned spins = 0;
    while (*word != seq) {
        __builtin_ia32_pause();
        if ((++spins & 0xfffffu) == 0) {
            cudaError_t q = cudaStreamQuery((cudaStream_t)stream);
            if (q != cudaErrorNotReady && q != cudaSuccess)
                return fail(SL_ERR_CUDA, "stream failed while waiting: %s", cudaGetErrorString(q));
            if (q == cudaSuccess && *word != seq)
                return fail(SL_ERR_ARG, "stream is idle but the completion word is %u, not %u", *word, seq);
        }
    }
    return SL_OK;
}

int sl_env_step_host(const SlEnvBatch *batch, const uint8_t *actions, int kernel, int flags,
                     void *stream, double *timing_us) {
    const double t0 = timing_us ? now_us() : 0.0;
    int rc = sl_env_step(batch, actions, kernel, flags, stream);
    if (rc) return rc;
    const double t1 = timing_us ? now_us() : 0.0;
    rc = wait_for_stream((cudaStream_t)stream);
    if (rc) return rc;
    if (timing_us) {
        timing_us[0] = t1 - t0;
        timing_us[1] = now_us() - t1;
    }
    return SL_OK;
}

int sl_env_step_sync(const SlEnvBatch *batch, const uint8_t *actions, int kernel, int flags,
                     void *stream) {
    return sl_env_step_host(batch, actions, kernel, flags, stream, nullptr);
}

// ---- host sessions (session.cuh) -----------------------------------------------
int sl_host_session_alloc(SlHostSession *s, int n_envs, int write_combined) {
    if (!s || n_envs < 1) return fail(SL_ERR_ARG, "bad arguments");
    memset(s, 0, sizeof *s);
    const size_t n_pad = ((size_t)n_envs + 15) & ~(size_t)15;
    void *p = nullptr;
    SL_CUDA(cudaHostAlloc(&p, n_pad, cudaHostAllocMapped | (write_combined ? cudaHostAllocWriteCombined : 0)));
    s->host_actions = (uint8_t *)p;
    SL_CUDA(cudaHostAlloc(&p, 64, cudaHostAllocMapped));
    s->host_ctrl = (uint32_t *)p;
    memset(s->host_ctrl, 0, 64);
    SL_CUDA(cudaMalloc(&p, 2 * n_pad));
    s->dev_actions = (uint8_t *)p;
    SL_CUDA(cudaMalloc(&p, SL_SESSION_CTRL_WORDS * sizeof(uint32_t)));
    s->dev_ctrl = (uint32_t *)p;

    s->n_envs = n_envs;
    return SL_OK;
}

int sl_host_session_timeline(const SlHostSession *s, uint64_t *stamps8) {
    if (!s || !s->dev_ctrl || !stamps8) return fail(SL_ERR_ARG, "null pointer");
    SL_CUDA(cudaMemcpy(stamps8, s->dev_ctrl + 8, 64, cudaMemcpyDeviceToHost));
    return SL_OK;
}

int sl_host_session_free(SlHostSession *s) {
    if (!s) return SL_OK;
    if (s->host_actions) cudaFreeHost(s->host_actions);
    if (s->host_ctrl) cudaFreeHost(s->host_ctrl);
    if (s->dev_actions) cudaFree(s->dev_actions);
    if (s->dev_ctrl) cudaFree(s->dev_ctrl);
    memset(s, 0, sizeof *s);
    return SL_OK;
}

int sl_host_session_begin(const SlEnvBatch *batch, SlHostSession *s, int flags, void *stream) {
    int rc = check_batch(batch);
    if (rc) return rc;
    if (!s || !s->host_actions || !s->dev_actions || !s->dev_ctrl || !s->host_ctrl)
        return fail(SL_ERR_ARG, "null pointer in SlHostSession");
    if (flags & ~SL_STEP_AUTO_RESET) return fail(SL_ERR_ARG, "unknown step flags");
    if (batch->n_envs < 1) return fail(SL_ERR_ARG, "empty batch");
    if (!batch->host_out) return fail(SL_ERR_ARG, "a host session needs SlEnvBatch.host_out");
    if (fast_step_path(batch->H, batch->W) != 1)
        return fail(SL_ERR_ARG, "host sessions take boards up to 32 x 32 (got %d x %d)", batch->H, batch->W);
    if ((((uintptr_t)batch->board | (uintptr_t)batch->goals | (uintptr_t)batch->state |
          (uintptr_t)s->host_actions | (uintptr_t)s->dev_actions) & 15) != 0)
        return fail(SL_ERR_ARG, "host sessions need 16-byte aligned arrays");
    DeviceProps props;
    rc = device_props(&props);
    if (rc) return rc;
    const int cells = batch->H * batch->W;
    const bool resident = !(batch->pool_flags & SL_POOL_STATIC_GOALS);
    const void *fn = resident ? (const void *)sl::k_env_session<true> : (const void *)sl::k_env_session<false>;
    const size_t smem = sl::warp_step_smem_bytes<sl::Narrow>(cells, resident, 0);
    rc = prepare_smem_ptr(fn, smem);
    if (rc) return rc;
    // every CTA must be resident at once: as many envs per CTA (= warps) as that takes
    int epc = 0, workers = 0;
    for (int e = 1; e <= sl::WARP_ENVS_PER_CTA && !epc; e++) {
        int per_sm = 0;
        SL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, e * 32, smem));
        const int w = (batch->n_envs + e - 1) / e;
        if (w + 1 <= per_sm * props.sm_count) {
            epc = e;
            workers = w;
        }
    }
    if (!epc)
        return fail(SL_ERR_ARG, "a host session keeps every env resident: %d envs of %d x %d are too many",
                    batch->n_envs, batch->H, batch->W);
    volatile uint32_t *hc = s->host_ctrl;
    hc[0] = 0;
    hc[1] = 1;
    hc[2] = 0;
    s->seq = 0;
    s->n_envs = batch->n_envs;
    s->host_out = batch->host_out;
    for (int i = 0; i < batch->n_envs; i++) ((volatile uint8_t *)s->host_out)[8 * (size_t)i + 7] = 0;   // no stale tags
    cudaStream_t st = (cudaStream_t)stream;
    static_assert(SL_SESSION_CTRL_WORDS == sl::SESSION_CTRL_WORDS, "mailbox size");
    SL_CUDA(cudaMemsetAsync(s->dev_ctrl, 0, SL_SESSION_CTRL_WORDS * sizeof(uint32_t), st));
    const int n_pad = (batch->n_envs + 15) & ~15;
    memset(s->host_actions, 0, (size_t)n_pad);
    sl::SessionParams P;
    P.host_actions = s->host_actions;
    P.dev_actions = s->dev_actions;
    P.dev_ctrl = s->dev_ctrl;
    P.host_ctrl = s->host_ctrl;
    P.n_pad = n_pad;

    P.idle_ns = (unsigned long long)(s->idle_timeout_ms ? s->idle_timeout_ms : 200) * 1000000ull;
    if (resident)
        sl::k_env_session<true><<<workers + 1, epc * 32, smem, st>>>(*batch, P, flags, epc, workers);
    else
        sl::k_env_session<false><<<workers + 1, epc * 32, smem, st>>>(*batch, P, flags, epc, workers);
    return launch_status("k_env_session");
}

int sl_host_session_step(SlHostSession *s, const uint8_t *actions, double *timing_us) {
    if (!s || !actions) return fail(SL_ERR_ARG, "null pointer");
    volatile uint32_t *hc = s->host_ctrl;
    if (hc[1] == 0) return SL_SESSION_ENDED;
    const double t0 = timing_us ? now_us() : 0.0;
    const uint32_t seq = s->seq + 1;
    const uint8_t tag = (uint8_t)((seq & 15u) << 4);
    const int n = s->n_envs, n_pad = (n + 15) & ~15;
    // (host_actions is write-combined memory when it comes from sl_host_session_alloc: the
    // stores drain to memory and the device's PCIe reads need no snoop of this core's cache --
    // with ordinary pinned memory the courier's 4 KB read of freshly written lines takes
    // 6.4 us instead of 1.7 us, tools/trace_session.py.  Only ever written here, in order.)
    uint8_t *dst = s->host_actions;
    const __m128i lo = _mm_set1_epi8(15), tg = _mm_set1_epi8((char)tag);
    int i = 0;
    for (; i + 16 <= n; i += 16) {
        const __m128i a = _mm_loadu_si128((const __m128i *)(actions + i));
        _mm_store_si128((__m128i *)(dst + i), _mm_or_si128(_mm_and_si128(a, lo), tg));
    }
    if (i < n_pad) {
        uint8_t last[16];
        for (int k = 0; k < 16; k++) last[k] = (uint8_t)((i + k < n ? (actions[i + k] & 15u) : 0u) | tag);
        _mm_store_si128((__m128i *)(dst + i), _mm_loadu_si128((const __m128i *)last));
    }
    _mm_sfence();
    const double t1 = timing_us ? now_us() : 0.0;
    // the last CTA to finish the step writes the sequence number into the done word
    bool ended = false;
    while (hc[0] != seq) {
        if (ended) return SL_SESSION_ENDED;        // idle timeout raced the post: nothing was stepped
        if (hc[1] == 0) ended = true;              // one more look at the done word, then give up
        __builtin_ia32_pause();
    }
    s->seq = seq;
    if (timing_us) {
        timing_us[0] = t1 - t0;
        timing_us[1] = now_us() - t1;
    }
    return SL_OK;
}

int sl_host_session_end(SlHostSession *s, void *stream) {
    if (!s || !s->host_ctrl) return fail(SL_ERR_ARG, "null pointer");
    volatile uint32_t *hc = s->host_ctrl;
    hc[2] = 1;
    __sync_synchronize();
    SL_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    hc[1] = 0;
    return SL_OK;
}

int sl_env_reset(const SlEnvBatch *batch, const uint8_t *mask, const int32_t *new_level,
                 int flags, void *stream) {
    int rc = check_batch(batch);
    if (rc) return rc;
    if (batch->n_envs == 0) return SL_OK;
    size_t smem = sl::generic_smem_bytes(batch->H * batch->W);
    rc = prepare_smem(sl::k_env_reset, smem);
    if (rc) return rc;
    // grid-stride over envs: a masked reset mostly skips, so keep the grid small
    int grid = batch->n_envs < 2048 ? batch->n_envs : 2048;
    sl::k_env_reset<<<grid, sl::GENERIC_THREADS, smem, (cudaStream_t)stream>>>(
        *batch, mask, new_level, flags);
    return launch_status("k_env_reset");
}

int sl_get_obs(const SlEnvBatch *batch, void *obs, int vh, int vw, const int32_t *channels,
               int n_channels, int remove_white_goals, void *stream) {
    int rc = check_batch(batch);
    if (rc) return rc;
    if (vh < 1 || vw < 1 || n_channels < 0) return fail(SL_ERR_ARG, "bad view shape");
    if (n_channels > 32) return fail(SL_ERR_ARG, "at most 32 output channels (bits of the uint32 view)");
    if (batch->n_envs == 0) return SL_OK;
    if (!obs || (n_channels > 0 && !channels)) return fail(SL_ERR_ARG, "null pointer");
    // the channel list is a device array here (ABI v1); it travels to the kernel by value
    SlObsSpec spec = {};
    spec.obs = obs;
    spec.vh = vh;
    spec.vw = vw;
    spec.n_channels = n_channels;
    spec.remove_white_goals = remove_white_goals;
    if (n_channels > 0) {
        cudaPointerAttributes attr;
        cudaError_t err = cudaPointerGetAttributes(&attr, channels);
        if (err != cudaSuccess) {
            cudaGetLastError();
            attr.type = cudaMemoryTypeUnregistered;
        }
        if (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged) {
            SL_CUDA(cudaMemcpyAsync(spec.channels, channels, (size_t)n_channels * 4,
                                    cudaMemcpyDeviceToHost, (cudaStream_t)stream));
            SL_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
        } else {
            memcpy(spec.channels, channels, (size_t)n_channels * 4);
        }
    }
    return sl_get_obs_spec(batch, &spec, stream);
}

int sl_get_obs_spec(const SlEnvBatch *batch, const SlObsSpec *obs, void *stream) {
    int rc = check_batch(batch);
    if (rc) return rc;
    if (!obs) return fail(SL_ERR_ARG, "null observation spec");
    rc = check_obs_spec(obs);
    if (rc) return rc;
    if (batch->n_envs == 0) return SL_OK;
    return launch_get_obs(batch, sl::make_obs_params(*obs), (cudaStream_t)stream);
}

int sl_side_effect_cells(const SlEnvBatch *batch, const uint16_t *baseline, int32_t *out,
                         void *stream) {
    int rc = check_batch(batch);
    if (rc) return rc;
    if (batch->n_envs == 0) return SL_OK;
    if (!baseline || !out) return fail(SL_ERR_ARG, "null pointer");
    sl::k_side_effect_cells<<<grid_for(batch->n_envs), sl::GENERIC_THREADS, 0, (cudaStream_t)stream>>>(
        *batch, baseline, out);
    return launch_status("k_side_effect_cells");
}

}  // extern "C"
