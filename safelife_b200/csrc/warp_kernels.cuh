// warp_kernels.cuh -- warp-per-env bit-plane kernels.
//
// The fast path.  One warp owns one env; a CTA of 8 warps stages the 8
// consecutive env slabs it owns ([8][H][W] uint16, 16-byte aligned as a group
// for any H, W) and their 8 state records from HBM into shared memory with
// bulk asynchronous copies (TMA) and writes them back the same way.  Inside
// the warp:
//
//   1. each lane reads the rows it owns and transposes them in registers into
//      bit-plane words (bit j of plane p = bit p of cell j): one byte-permute
//      stage and three masked shift stages per 32 cells;
//   2. the 3x3 toroidal neighbourhood aggregate of advance_board.c:34-125 is
//      evaluated on whole rows at once with bitwise adders: horizontal
//      neighbours are 1-bit rotations of the row word, vertical neighbours come
//      from neighbouring lanes by warp shuffle;
//   3. the spawn draws of the reference (one PCG64 double per qualifying cell
//      in raster order, advance_board.c:115) are produced 32 at a time by
//      leapfrogging the LCG across lanes and handed to the rows by prefix sum;
//   4. only cells that changed are written back into the uint16 slab, and the
//      points sum is updated incrementally from those cells.
//
// Two geometries share all of that code through templates:
//   Narrow  H, W <= 32: lane r owns row r, a row is one 32-bit word
//           (the BASELINE shapes 25x25 / 26x26);
//   Wide    H, W <= 64: lane r owns rows r and r + 32, a row is a 64-bit word
//           (50x50 navigation boards).
//
// No tensor cores: the path is integer / bitwise.
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>
#include "cells.cuh"
#include "generic_kernels.cuh"
#include "obs.cuh"
#include "wrap.cuh"
#include "../../include/safelife_b200.h"

namespace sl {

constexpr int WARP_MAX_DIM = 64;
constexpr int WARP_NARROW_DIM = 32;
constexpr int WARP_ENVS_PER_CTA = 8;
constexpr unsigned FULL = 0xffffffffu;

inline bool warp_kernel_supports(int H, int W) {
    return H >= 3 && W >= 3 && H <= WARP_MAX_DIM && W <= WARP_MAX_DIM;
}
inline bool warp_kernel_is_narrow(int H, int W) {
    return H <= WARP_NARROW_DIM && W <= WARP_NARROW_DIM;
}

// Per-lane leapfrog constants (A^(l+1), G_(l+1)); global memory so that each
// lane reads its own entry without constant-cache serialisation.
__device__ const uint64_t d_lane_a[32][2] = SL_PCG_LANE_A;
__device__ const uint64_t d_lane_g[32][2] = SL_PCG_LANE_G;

// ------------------------------------------------------------------ planes
// A plane is "one bit per cell for the rows this lane owns": uint32_t in the
// narrow geometry, P2 (two 64-bit rows) in the wide one.
struct P2 {
    uint64_t a, b;   // row `lane`, row `lane + 32`
};
__device__ __forceinline__ P2 operator&(P2 x, P2 y) { return {x.a & y.a, x.b & y.b}; }
__device__ __forceinline__ P2 operator|(P2 x, P2 y) { return {x.a | y.a, x.b | y.b}; }
__device__ __forceinline__ P2 operator^(P2 x, P2 y) { return {x.a ^ y.a, x.b ^ y.b}; }
__device__ __forceinline__ P2 operator~(P2 x) { return {~x.a, ~x.b}; }
__device__ __forceinline__ bool nonzero(uint32_t x) { return x != 0; }
__device__ __forceinline__ bool nonzero(P2 x) { return (x.a | x.b) != 0; }
__device__ __forceinline__ int popc(uint32_t x) { return __popc(x); }
__device__ __forceinline__ int popc(uint64_t x) { return __popcll(x); }
__device__ __forceinline__ int lowest(uint32_t x) { return __ffs(x) - 1; }
__device__ __forceinline__ int lowest(uint64_t x) { return __ffsll((long long)x) - 1; }

template <class P> struct LowPlanesT { P A, D, F, P_, I, S; };   // bits 0, 3, 4, 5, 6, 7
template <class P> struct HighPlanesT { P E, R, G, B; };         // bits 8, 9, 10, 11
template <class P> struct RowDeltaT {
    P die;          // alive cells that die
    P born;         // dead cells that come alive through n == 3
    P cand;         // dead cells that take a spawn draw
    P tR, tG, tB;   // inherited colour bits of a new cell
    P dn;           // births that are DESTRUCTIBLE
};

// ------------------------------------------------------ 32-cell transposes
// The integer ALU pipe (LOP3 / SHF / PRMT, one warp instruction per two cycles
// per scheduler) is what bounds the step kernel, so the transposes are written
// to spend as little of it as possible: a butterfly half is ONE three-input
// LOP3 (bit select under an immediate mask) on top of one shift, left shifts
// and the 16-bit packing go to the FMA pipe as multiply-adds.

// (a & mask) | (b & ~mask) in one LOP3.
__device__ __forceinline__ uint32_t bitsel(uint32_t a, uint32_t b, uint32_t mask) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xE2;" : "=r"(d) : "r"(a), "r"(mask), "r"(b));
    return d;
}
// lo | hi << 16 for two 16-bit values, as a multiply-add.
__device__ __forceinline__ uint32_t pack16(uint32_t lo, uint32_t hi) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, 65536, %2;" : "=r"(d) : "r"(hi), "r"(lo));
    return d;
}
// One butterfly stage on the pair (a, b): a keeps its bits under ~MASK << 0 ...
// i.e. a' = (a & ~MASK) | ((b << S) & MASK),  b' = ((a >> S) & ~MASK) | (b & MASK).
template <int S, uint32_t MASK>
__device__ __forceinline__ void butterfly(uint32_t &a, uint32_t &b) {
    const uint32_t a2 = bitsel(b << S, a, MASK);
    b = bitsel(b, a >> S, MASK);
    a = a2;
}

// 16 cell-pair words -> 16 plane words.  x[k] = cell(k) | cell(k+16) << 16.
__device__ __forceinline__ void transpose16(uint32_t (&x)[16]) {
#pragma unroll
    for (int k = 0; k < 8; k++) {            // bytes: cell bit 3 <-> plane bit 3
        uint32_t a = x[k], b = x[k + 8];
        x[k] = __byte_perm(a, b, 0x6240);
        x[k + 8] = __byte_perm(a, b, 0x7351);
    }
#pragma unroll
    for (int k = 0; k < 16; k++) {           // nibbles
        if (k & 4) continue;
        butterfly<4, 0xf0f0f0f0u>(x[k], x[k + 4]);
    }
#pragma unroll
    for (int k = 0; k < 16; k++) {           // bit pairs
        if (k & 2) continue;
        butterfly<2, 0xccccccccu>(x[k], x[k + 2]);
    }
#pragma unroll
    for (int k = 0; k < 16; k += 2)          // bits
        butterfly<1, 0xaaaaaaaau>(x[k], x[k + 1]);
}

// 8 words of 4 bytes (cells i, i+8, i+16, i+24) -> 8 plane words (index = bit).
__device__ __forceinline__ void transpose8(uint32_t (&x)[8]) {
#pragma unroll
    for (int k = 0; k < 4; k++) butterfly<4, 0xf0f0f0f0u>(x[k], x[k + 4]);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        if (k & 2) continue;
        butterfly<2, 0xccccccccu>(x[k], x[k + 2]);
    }
#pragma unroll
    for (int k = 0; k < 8; k += 2) butterfly<1, 0xaaaaaaaau>(x[k], x[k + 1]);
}

// Planes of the low (SEL 0x6240) or high (SEL 0x7351) byte of 32 consecutive
// cells starting at `cells` (shared memory; callers pad their slabs so that the
// read never leaves it).  x[p] = plane of bit p (low) / bit 8 + p (high).
template <unsigned SEL>
__device__ __forceinline__ void planes32(const uint16_t *cells, uint32_t (&x)[8]) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
        uint32_t a = pack16(cells[k], cells[k + 16]);
        uint32_t b = pack16(cells[k + 8], cells[k + 24]);
        x[k] = __byte_perm(a, b, SEL);
    }
    transpose8(x);
}

// ------------------------------------------------------------- geometries
// Narrow: lane r owns row r (r < H <= 32), a row is a 32-bit word.
struct Narrow {
    typedef uint32_t P;
    typedef uint32_t RowMask;
    static constexpr int ROWS = 1;
    static constexpr int PAD_CELLS = 32 * 32 + 32;   // row over-read past the last slab
    int H, W, lane, up, down, row_off;
    uint32_t wmask;

    __device__ __forceinline__ Narrow(int H_, int W_, int lane_) : H(H_), W(W_), lane(lane_) {
        int r = lane < H ? lane : 0;
        row_off = r * W;     // lanes past the last row re-read row 0 of their own env (masked out)
        up = r == 0 ? H - 1 : r - 1;
        down = r == H - 1 ? 0 : r + 1;
        // lanes past the last row carry an empty mask: they never create, kill or draw
        wmask = lane < H ? (W == 32 ? FULL : ((1u << W) - 1u)) : 0u;
    }
    __device__ __forceinline__ P mask() const { return wmask; }
    __device__ __forceinline__ P west(P x) const { return ((x << 1) | (x >> (W - 1))) & wmask; }
    __device__ __forceinline__ P east(P x) const { return ((x >> 1) | (x << (W - 1))) & wmask; }
    __device__ __forceinline__ P north(P x) const { return __shfl_sync(FULL, x, up); }
    __device__ __forceinline__ P south(P x) const { return __shfl_sync(FULL, x, down); }
    __device__ __forceinline__ int row_of(int) const { return lane; }
    __device__ __forceinline__ RowMask row(P x, int) const { return x; }
    __device__ __forceinline__ void set_row(P &x, int, RowMask v) const { x = v; }

    __device__ __forceinline__ LowPlanesT<P> load_low(const uint16_t *slab) const {
        uint32_t x[8];
        planes32<0x6240>(slab + row_off, x);
        return {x[0] & wmask, x[3] & wmask, x[4] & wmask, x[5] & wmask, x[6] & wmask, x[7] & wmask};
    }
    __device__ __forceinline__ HighPlanesT<P> load_high(const uint16_t *slab) const {
        uint32_t x[8];
        planes32<0x7351>(slab + row_off, x);
        return {x[0] & wmask, x[1] & wmask, x[2] & wmask, x[3] & wmask};
    }
};

// Wide: lane r owns rows r and r + 32 (H <= 64), a row is a 64-bit word (W <= 64).
struct Wide {
    typedef P2 P;
    typedef uint64_t RowMask;
    static constexpr int ROWS = 2;
    static constexpr int PAD_CELLS = 64 * 64 + 64;
    int H, W, lane;
    uint64_t mask_a, mask_b;
    int n_src_b, s_src_a;      // shuffle sources that depend on the lane
    bool b_is_last;            // row lane + 32 is the last row of the board

    __device__ __forceinline__ Wide(int H_, int W_, int lane_) : H(H_), W(W_), lane(lane_) {
        const uint64_t full = W == 64 ? ~0ull : ((1ull << W) - 1ull);
        mask_a = lane < H ? full : 0ull;
        mask_b = lane + 32 < H ? full : 0ull;
        // north of row 0 is row H-1: slot b of lane H-33 when H > 32
        n_src_b = lane == 0 ? (H > 32 ? H - 33 : 0) : lane - 1;
        // south of row lane: row 0 if it is the last row, else slot a of the next lane
        s_src_a = lane == H - 1 ? 0 : ((lane + 1) & 31);
        b_is_last = lane + 32 == H - 1;
    }
    __device__ __forceinline__ P mask() const { return {mask_a, mask_b}; }
    __device__ __forceinline__ uint64_t rotl(uint64_t x, uint64_t m) const {
        return ((x << 1) | (x >> (W - 1))) & m;
    }
    __device__ __forceinline__ uint64_t rotr(uint64_t x, uint64_t m) const {
        return ((x >> 1) | (x << (W - 1))) & m;
    }
    __device__ __forceinline__ P west(P x) const { return {rotl(x.a, mask_a), rotl(x.b, mask_b)}; }
    __device__ __forceinline__ P east(P x) const { return {rotr(x.a, mask_a), rotr(x.b, mask_b)}; }
    __device__ __forceinline__ P north(P x) const {
        if (H <= 32) {
            const int src = lane == 0 ? H - 1 : (lane < H ? lane - 1 : 0);
            return {__shfl_sync(FULL, x.a, src), 0ull};
        }
        const uint64_t u = __shfl_sync(FULL, x.a, (lane + 31) & 31);   // slot a of the previous lane
        const uint64_t v = __shfl_sync(FULL, x.b, n_src_b);
        return {lane == 0 ? v : u, lane == 0 ? u : v};
    }
    __device__ __forceinline__ P south(P x) const {
        const uint64_t u = __shfl_sync(FULL, x.a, s_src_a);
        if (H <= 32) return {u, 0ull};
        const uint64_t v = __shfl_sync(FULL, x.b, (lane + 1) & 31);    // slot b of the next lane
        const uint64_t row0 = __shfl_sync(FULL, x.a, 0);
        return {lane == 31 ? v : u, b_is_last ? row0 : v};
    }
    __device__ __forceinline__ int row_of(int slot) const { return lane + 32 * slot; }
    __device__ __forceinline__ RowMask row(P x, int slot) const { return slot ? x.b : x.a; }
    __device__ __forceinline__ void set_row(P &x, int slot, RowMask v) const {
        if (slot) x.b = v; else x.a = v;
    }

    template <unsigned SEL>
    __device__ __forceinline__ void load_bytes(const uint16_t *slab, uint64_t (&pa)[8],
                                               uint64_t (&pb)[8]) const {
        uint32_t lo[8], hi[8];
        planes32<SEL>(slab + lane * W, lo);
        planes32<SEL>(slab + lane * W + 32, hi);
#pragma unroll
        for (int k = 0; k < 8; k++) pa[k] = ((uint64_t)lo[k] | ((uint64_t)hi[k] << 32)) & mask_a;
        planes32<SEL>(slab + (lane + 32) * W, lo);
        planes32<SEL>(slab + (lane + 32) * W + 32, hi);
#pragma unroll
        for (int k = 0; k < 8; k++) pb[k] = ((uint64_t)lo[k] | ((uint64_t)hi[k] << 32)) & mask_b;
    }
    __device__ __forceinline__ LowPlanesT<P> load_low(const uint16_t *slab) const {
        uint64_t a[8], b[8];
        load_bytes<0x6240>(slab, a, b);
        return {{a[0], b[0]}, {a[3], b[3]}, {a[4], b[4]}, {a[5], b[5]}, {a[6], b[6]}, {a[7], b[7]}};
    }
    __device__ __forceinline__ HighPlanesT<P> load_high(const uint16_t *slab) const {
        uint64_t a[8], b[8];
        load_bytes<0x7351>(slab, a, b);
        return {{a[0], b[0]}, {a[1], b[1]}, {a[2], b[2]}, {a[3], b[3]}};
    }
};

// ------------------------------------------------------------------ stencil
template <class P>
__device__ __forceinline__ P maj3(P a, P b, P c) {
    return (a & b) | (a & c) | (b & c);
}

// "at least one" / "at least two" of a plane over the 3x3 neighbourhood.
template <class G>
__device__ __forceinline__ typename G::P any3x3(const G &t, typename G::P x) {
    typename G::P h = x | t.west(x) | t.east(x);
    return h | t.north(h) | t.south(h);
}

template <class G>
__device__ __forceinline__ typename G::P two3x3(const G &t, typename G::P x) {
    typedef typename G::P P;
    P w = t.west(x), e = t.east(x);
    P one = x | w | e, two = maj3(x, w, e);
    P on = t.north(one), os = t.south(one);
    return two | t.north(two) | t.south(two) | maj3(one, on, os);
}

// advance_board.c:34-125 on bit planes.  `high()` supplies the high-byte planes
// and is only called if some cell of the board may be created.
template <class G, class HighFn>
__device__ __forceinline__ RowDeltaT<typename G::P> stencil_planes(
    const G &t, const LowPlanesT<typename G::P> &p, HighFn high) {
    typedef typename G::P P;
    // n = alive cells in the 3x3 (centre included), as 4 bit planes
    P w = t.west(p.A), e = t.east(p.A);
    P s0 = p.A ^ w ^ e, s1 = maj3(p.A, w, e);
    P s0n = t.north(s0), s0s = t.south(s0), s1n = t.north(s1), s1s = t.south(s1);
    P n0 = s0 ^ s0n ^ s0s, c0 = maj3(s0, s0n, s0s);
    P a = s1 ^ s1n ^ s1s, b = maj3(s1, s1n, s1s);
    P n1 = a ^ c0, c1 = a & c0;
    P n2 = b ^ c1, n3 = b & c1;
    P is3 = n0 & n1 & ~n2 & ~n3;
    P is4 = ~n0 & ~n1 & n2 & ~n3;

    P any_p = any3x3(t, p.P_), any_i = any3x3(t, p.I), any_s = any3x3(t, p.S);
    RowDeltaT<P> d;
    d.die = p.A & ~(p.F | any_p | is3 | is4) & t.mask();   // (tiles carry halo cells in their planes)
    P open = ~p.A & ~p.F & ~any_i & t.mask();   // dead, not frozen, not inhibited
    d.born = open & is3;
    d.cand = open & ~is3 & any_s;
    // colours / destructibility only matter where something may be created
    if (__any_sync(FULL, nonzero(d.born | d.cand))) {
        const HighPlanesT<P> h = high();
        d.tR = two3x3(t, p.A & h.R) | any3x3(t, p.S & h.R);
        d.tG = two3x3(t, p.A & h.G) | any3x3(t, p.S & h.G);
        d.tB = two3x3(t, p.A & h.B) | any3x3(t, p.S & h.B);
        d.dn = two3x3(t, p.A & (p.D | h.E));
    } else {
        d.tR = d.tG = d.tB = d.dn = d.die & ~d.die;   // zero
    }
    return d;
}

// Resolve the spawn draws of one board: each lane holds the candidate masks of
// its rows; returns the subset whose draw was below thr; the (uniform) state
// advances by the total number of candidates.  Raster order = row order (all
// rows of slot 0 come before those of slot 1), then bit order within the row.
// Per-lane leapfrog constants of one stream: lane l jumps l + 1 steps at once,
// state' = la * state + ginc with la = A^(l+1), ginc = G_(l+1) * inc.
struct SpawnCtx {
    u128 la, ginc;
};
__device__ __forceinline__ SpawnCtx make_spawn_ctx(int lane, const u128 &inc) {
    const u128 la = {__ldg(&d_lane_a[lane][0]), __ldg(&d_lane_a[lane][1])};
    const u128 lg = {__ldg(&d_lane_g[lane][0]), __ldg(&d_lane_g[lane][1])};
    return {la, mul128(lg, inc)};
}

template <class G>
__device__ __forceinline__ typename G::P warp_spawn_draws_ctx(const G &t, u128 &state,
                                                              const SpawnCtx &ctx, uint64_t thr,
                                                              typename G::P cand) {
    typedef typename G::P P;
    typedef typename G::RowMask M;
    const int lane = t.lane;
    M todo[G::ROWS];
    int next[G::ROWS];      // index of the next draw of each row
    int total = 0;
#pragma unroll
    for (int s = 0; s < G::ROWS; s++) {
        todo[s] = t.row(cand, s);
        int cnt = popc(todo[s]), incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int v = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += v;
        }
        next[s] = total + incl - cnt;
        total += __shfl_sync(FULL, incl, 31);
    }
    M got[G::ROWS];
#pragma unroll
    for (int s = 0; s < G::ROWS; s++) got[s] = 0;
    for (int base = 0; base < total; base += 32) {
        u128 sn = add128(mul128(ctx.la, state), ctx.ginc);  // state after lane+1 more steps
        uint32_t ok = __ballot_sync(FULL, (pcg_output(sn) >> 11) < thr);
#pragma unroll
        for (int s = 0; s < G::ROWS; s++) {
            while (todo[s] != 0 && next[s] < base + 32) {
                int j = lowest(todo[s]);
                todo[s] &= todo[s] - 1;
                got[s] |= (M)((ok >> (next[s] - base)) & 1u) << j;
                next[s]++;
            }
        }
        int last = min(31, total - base - 1);
        state.hi = __shfl_sync(FULL, sn.hi, last);
        state.lo = __shfl_sync(FULL, sn.lo, last);
    }
    P spawned = cand & ~cand;
#pragma unroll
    for (int s = 0; s < G::ROWS; s++) t.set_row(spawned, s, got[s]);
    return spawned;
}

// Resolve the spawn draws of one board: each lane holds the candidate masks of
// its rows; returns the subset whose draw was below thr; the (uniform) state
// advances by the total number of candidates.  Raster order = row order (all
// rows of slot 0 come before those of slot 1), then bit order within the row.
template <class G>
__device__ __forceinline__ typename G::P warp_spawn_draws(const G &t, u128 &state, const u128 &inc,
                                                          uint64_t thr, typename G::P cand) {
    if (!__any_sync(FULL, nonzero(cand))) return cand & ~cand;
    return warp_spawn_draws_ctx(t, state, make_spawn_ctx(t.lane, inc), thr, cand);
}

// New value of a changed cell: 0 if it died, else ALIVE | inherited colours |
// DESTRUCTIBLE (advance_board.c:104, 111-118).
template <class M>
__device__ __forceinline__ uint32_t fresh_cell(M f, M de, M r, M g, M b, int j) {
    if (!((f >> j) & 1)) return 0u;
    return ALIVE | ((uint32_t)((de >> j) & 1) << 3) | ((uint32_t)((r >> j) & 1) << 9) |
           ((uint32_t)((g >> j) & 1) << 10) | ((uint32_t)((b >> j) & 1) << 11);
}

// Write the board cells that a resolved step changed back into the uint16
// slab.  *delta accumulates what the changes add to sum(table * alive_counts),
// priced with the goals as they are in `goals`.
template <class G>
__device__ __forceinline__ void apply_board(const G &t, const RowDeltaT<typename G::P> &d,
                                            typename G::P spawned, uint16_t *board,
                                            const uint16_t *goals, const int32_t *table,
                                            bool dead_counts, int *delta) {
    typedef typename G::P P;
    typedef typename G::RowMask M;
    const P fresh = d.born | spawned;
    const P changed = d.die | fresh;
    if (!nonzero(changed)) return;
    const P destr = (d.dn & d.born) | spawned;
#pragma unroll
    for (int s = 0; s < G::ROWS; s++) {
        M m = t.row(changed, s);
        const M f = t.row(fresh, s), de = t.row(destr, s);
        const M r = t.row(d.tR, s), g = t.row(d.tG, s), b = t.row(d.tB, s);
        const int base = t.row_of(s) * t.W;
        while (m != 0) {
            const int j = lowest(m);
            m &= m - 1;
            const uint32_t v = fresh_cell(f, de, r, g, b, j);
            const uint32_t gl = goals[base + j];
            *delta += cell_points(v, gl, table, dead_counts) -
                      cell_points(board[base + j], gl, table, dead_counts);
            board[base + j] = (uint16_t)v;
        }
    }
}

// The same for the goals plane; a goal cell whose colour changes re-prices the
// board cell on top of it (`board` is already the new board).  Returns the
// mask of changed goal cells.
template <class G>
__device__ __forceinline__ typename G::P apply_goals(const G &t, const RowDeltaT<typename G::P> &d,
                                                     typename G::P spawned, uint16_t *goals,
                                                     const uint16_t *board, const int32_t *table,
                                                     bool dead_counts, int *delta) {
    typedef typename G::P P;
    typedef typename G::RowMask M;
    const P fresh = d.born | spawned;
    const P changed = d.die | fresh;
    if (!nonzero(changed)) return changed;
    const P destr = (d.dn & d.born) | spawned;
#pragma unroll
    for (int s = 0; s < G::ROWS; s++) {
        M m = t.row(changed, s);
        const M f = t.row(fresh, s), de = t.row(destr, s);
        const M r = t.row(d.tR, s), g = t.row(d.tG, s), b = t.row(d.tB, s);
        const int base = t.row_of(s) * t.W;
        while (m != 0) {
            const int j = lowest(m);
            m &= m - 1;
            const uint32_t v = fresh_cell(f, de, r, g, b, j);
            const uint32_t old = goals[base + j];
            if ((old ^ v) & COLORS) {
                const uint32_t cell = board[base + j];
                *delta += cell_points(cell, v, table, dead_counts) -
                          cell_points(cell, old, table, dead_counts);
            }
            goals[base + j] = (uint16_t)v;
        }
    }
    return changed;
}

// sum(table * alive_counts) of a board / goals pair in shared memory, any size
// (advance_board.c:192-207 + safelife_game.py:684-687), reduced over the warp.
// Only used by the in-kernel reset.
__device__ __forceinline__ int warp_points_slab(const uint16_t *board, const uint16_t *goals,
                                                int cells, const int32_t *table, int lane) {
    int pts = 0;
    for (int i = lane; i < cells; i += 32) pts += cell_points(board[i], goals[i], table, true);
    return __reduce_add_sync(FULL, pts);
}

// ---- bulk asynchronous copies (TMA, 1-D) and their mbarrier -----------------
__device__ __forceinline__ void copy_slabs(void *dst, const void *src, int bytes) {
    const uint16_t *s2 = (const uint16_t *)src;
    uint16_t *d2 = (uint16_t *)dst;
    for (int i = threadIdx.x; i < (bytes >> 1); i += blockDim.x) d2[i] = s2[i];
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; "
            "selp.u32 %0, 1, 0, p; }"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
// The same with a suspend-time hint: a waiting warp is parked by the hardware for
// up to `ns` instead of re-issuing the test, so the waits of the pipelined
// kernel do not compete for issue slots with the warps that have work.
__device__ __forceinline__ void mbar_wait_parked(uint64_t *bar, uint32_t parity, uint32_t ns) {
    uint32_t ok;
    do {
        asm volatile(
            "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; "
            "selp.u32 %0, 1, 0, p; }"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_load(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                          uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void bulk_store(void *dst_gmem, const void *src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
                 "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit_and_wait_read() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- programmatic dependent launch -------------------------------------------
// The step kernels are launched with programmatic stream serialization
// (launch_pdl below): the next launch of the stream may be scheduled onto the SMs
// while this one drains, and runs its prologue (shared-memory carve-up, mbarrier
// init) there.  `pdl_launch_dependents` allows that as soon as every CTA of this
// grid has started; `pdl_wait` returns once the grids before this one have
// completed and their memory is visible -- a kernel reads or writes nothing that
// an earlier launch may have produced (state, boards, actions, the level pool)
// before it.  Both are no-ops in a launch without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Optional per-CTA timeline for tools/trace_step.py (compiled in with -DSL_TRACE
// into a separate diagnostic library, never into libsafelife_b200.so proper).
#ifdef SL_TRACE
__device__ unsigned long long *sl_trace_buf = nullptr;   // [grid][8] globaltimer ns
__device__ __forceinline__ void trace_mark(int slot) {
    if (sl_trace_buf != nullptr) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        sl_trace_buf[(size_t)blockIdx.x * 8 + slot] = t;
    }
}
#define SL_TRACE_MARK(cond, slot) do { if (cond) trace_mark(slot); } while (0)
#else
#define SL_TRACE_MARK(cond, slot) do { } while (0)
#endif

// ---- shared memory of the step kernels ----------------------------------------
// How a step kernel sees the goals of its envs:
//   resident (GR = true)   the goal slabs of the group travel through shared
//                          memory with the boards (any pool);
//   in place (GR = false)  the caller promised SL_POOL_STATIC_GOALS: goal slabs
//                          stay in HBM, the few goal cells that price a changed
//                          board cell are read from there, and a goals plane
//                          that is not (yet) known to be static -- the first
//                          step of every episode, safelife_game.py:753-761 --
//                          is advanced through one scratch slab per CTA.  A
//                          broken promise costs time, not correctness.
constexpr int WARP_SMEM_HEAD = 48;                                  // 2 mbarriers, scratch lock, 16 action bytes
constexpr int WARP_SMEM_TABLE = 72 * 4;                             // the pool's points table
constexpr int WARP_SMEM_STATE = WARP_ENVS_PER_CTA * 96;             // SlEnvState[8]

__host__ __device__ inline int warp_slab_stride(int cells) {
    return (WARP_ENVS_PER_CTA * cells * 2 + 15) & ~15;
}
__host__ __device__ inline int warp_scratch_bytes(int cells) { return (cells * 2 + 15) & ~15; }
// The single-pass kernel takes any number of envs per CTA (1..8), so a CTA's slabs may
// start anywhere in a 16-byte unit of the board array: its shared-memory copy keeps
// that offset (bulk copies move whole 16-byte units), hence 16 spare bytes per plane.
__host__ __device__ inline int warp_slab_stride_any(int cells) { return warp_slab_stride(cells) + 16; }

template <class G>
__host__ __device__ inline size_t warp_step_smem_bytes(int cells, bool goals_resident, int obs_stage = 0) {
    return WARP_SMEM_HEAD + WARP_SMEM_TABLE + WARP_SMEM_STATE +
           (goals_resident ? 2 * (size_t)warp_slab_stride_any(cells)
                           : (size_t)warp_slab_stride_any(cells) + warp_scratch_bytes(cells)) +
           2 * G::PAD_CELLS + (size_t)WARP_ENVS_PER_CTA * obs_stage;
}

// The points table of the pool in shared memory when every level shares one
// (SL_POOL_UNIFORM_TABLE): pricing a changed cell then never leaves the SM.
__device__ __forceinline__ const int32_t *stage_table(const SlEnvBatch &B, int32_t *smem_table) {
    if (!(B.pool_flags & SL_POOL_UNIFORM_TABLE)) return nullptr;
    for (int i = threadIdx.x; i < 72; i += blockDim.x) smem_table[i] = __ldg(B.pool_table + i);
    return smem_table;     // visible after the caller's next CTA barrier
}

// SafeLifeEnv.reset of one env by a whole warp, inside the step kernel
// (safelife_env.py:203-218): next level of the pool into the board slab (shared
// memory) and the goals slab (shared memory or, in place, HBM), points recount,
// update_exit_colors, fresh episode state in the shared-memory record.  Rare
// (once per episode), so kept out of line.
__device__ __noinline__ void warp_reset_env(const SlEnvBatch &B, SlEnvState *st, uint16_t *board,
                                            uint16_t *goals, int lane) {
    const int H = B.H, W = B.W, cells = H * W;
    const int lvl = (int)(((long long)st->level_id + B.level_stride) % B.n_levels);
    const uint16_t *pb = B.pool_board + (size_t)lvl * cells;
    const uint16_t *pg = B.pool_goals + (size_t)lvl * cells;
    for (int i = lane; i < cells; i += 32) {
        board[i] = pb[i];
        goals[i] = pg[i];
    }
    const int32_t *table = B.pool_table + (size_t)lvl * 72;
    __syncwarp();
    const bool empty = __any_sync(FULL, lane < 8 && table[lane * 9 + 8] != 0);
    const int pts = warp_points_slab(board, pg, cells, table, lane);
    if (lane == 0) {
        const int32_t *exits = B.pool_exits + (size_t)lvl * B.max_exits;
        int n_exits = 0;
        while (n_exits < B.max_exits && exits[n_exits] >= 0) n_exits++;
        st->level_id = lvl;
        st->init_points = B.pool_init_points[lvl];
        st->required = B.pool_required[lvl];
        st->spawn_thr = B.pool_spawn_thr[lvl];
        st->empty_points = empty;
        st->n_exits = n_exits;
        st->exit0 = n_exits > 0 ? exits[0] : -1;
        st->exit1 = n_exits > 1 ? exits[1] : -1;
        bool agent_alive, exited;
        int dmem;
        const int row = B.pool_loc[2 * lvl], col = B.pool_loc[2 * lvl + 1];
        const int value = exit_update(board, pg, W, row, col, pts, *st, exits, table,
                                      &agent_alive, &exited, &dmem);
        st->loc_row = row;
        st->loc_col = col;
        st->num_steps = 0;
        st->old_value = value;
        st->count_points = pts + dmem;
        st->active = 1;
        st->static_goals = 0;   // episode totals stay (SL_RESET_KEEP_STATS)
    }
    __syncwarp();
}

// What the warps of one CTA share besides the stage buffers.
struct CtaShared {
    const int32_t *table;   // points table in shared memory, or nullptr: per level, in HBM
    uint16_t *scratch;      // in-place goals only: the one scratch slab of the CTA
    int *lock;              //                      and its lock
};

// ---- phase A: execute_actions + advance_board of one env by one warp ----------
// `st`, `board` are the env's record and board slab in shared memory; `goals` is
// its goals slab in shared memory (GR) or in HBM (in place).  Returns true if a
// shared-memory goals slab was modified.
template <class G, bool GR>
__device__ __forceinline__ bool step_phase_a(const SlEnvBatch &B, const CtaShared &cs, const G &t,
                                             SlEnvState *st, uint16_t *board, uint16_t *goals,
                                             uint64_t *goals_bar, int action) {
    typedef typename G::P P;
    const int H = B.H, W = B.W, lane = t.lane;
    bool goals_dirty = false;
    const uint64_t thr = st->spawn_thr;
    u128 state = {st->rng[0], st->rng[1]};
    const u128 inc = {st->rng[2], st->rng[3]};
    const uint64_t state_lo0 = state.lo;
    int sg = st->static_goals;
    int delta = 0;                    // change of sum(table * alive_counts), this lane
    // points table: shared memory if the pool has one table, else the level's in
    // HBM; looked up only for cells that changed (dead cells not at all unless
    // the "empty" column counts)
    const int32_t *table = cs.table ? cs.table : B.pool_table + (size_t)st->level_id * 72;
    const bool dead_counts = st->empty_points != 0;

    // execute_actions (advance_board.c:217-300): lane 0 acts, lanes 0-3 account
    // for the (at most four) cells it can touch; they are priced after the
    // stencil.  An in-place (HBM) goal cell is requested now, so that its
    // latency hides behind the stencil; a resident one is read once the goal
    // slabs have landed (`goals_bar`, nullptr: they already have).
    int row = st->loc_row, col = st->loc_col;
    uint32_t before = 0, after = 0, goal_at = 0;
    int idx = 0;
    if (row >= 0 && action != 0) {
        const ActionCells ac = action_cells(H, W, row, col, action);
        idx = lane == 0 ? ac.i0 : (lane == 1 ? ac.i1 : (lane == 2 ? ac.i2 : ac.i3));
        const bool mine = lane < 4 && !(lane == 3 && ac.i3 == ac.i2);
        if (mine) {
            before = board[idx];
            if (!GR) goal_at = goals[idx];
        }
        __syncwarp();
        if (lane == 0) {
            agent_action_at(board, ac, row, col, action);
            st->loc_row = row;
            st->loc_col = col;
        }
        __syncwarp();
        after = mine ? board[idx] : 0u;
    }

    {
        const LowPlanesT<P> p = t.load_low(board);
        const RowDeltaT<P> d = stencil_planes(t, p, [&] { return t.load_high(board); });
        const P spawned = warp_spawn_draws(t, state, inc, thr, d.cand);
        if (GR && goals_bar != nullptr) mbar_wait(goals_bar, 0);
        apply_board(t, d, spawned, board, goals, table, dead_counts, &delta);
    }
    if (after != before) {
        if (GR) goal_at = goals[idx];
        delta += cell_points(after, goal_at, table, dead_counts) -
                 cell_points(before, goal_at, table, dead_counts);
    }
    if (sg != 1) {   // goals advance from the same stream (safelife_game.py:753-761)
        uint16_t *gs = goals;
        if (!GR) {   // through the CTA's scratch slab, one warp at a time
            if (lane == 0)
                while (atomicCAS(cs.lock, 0, 1) != 0) __nanosleep(200);
            __threadfence_block();
            __syncwarp();
            for (int i = lane; i < H * W; i += 32) cs.scratch[i] = goals[i];
            gs = cs.scratch;
        }
        __syncwarp();
        const LowPlanesT<P> gp = t.load_low(gs);
        const RowDeltaT<P> gd = stencil_planes(t, gp, [&] { return t.load_high(gs); });
        const P gspawned = warp_spawn_draws(t, state, inc, thr, gd.cand);
        const P gchanged = apply_goals(t, gd, gspawned, gs, board, table, dead_counts, &delta);
        // static iff nothing changed and no spawner is left (safelife_game.py:756-760)
        const bool any_changed = __any_sync(FULL, nonzero(gchanged));
        if (sg == 0) sg = (any_changed || __any_sync(FULL, nonzero(gp.S & ~gchanged))) ? 2 : 1;
        if (GR) {
            goals_dirty = true;
        } else {
            __syncwarp();
            if (any_changed)
                for (int i = lane; i < H * W; i += 32) goals[i] = cs.scratch[i];
            __threadfence_block();
            __syncwarp();
            if (lane == 0) atomicExch(cs.lock, 0);
        }
    }
    delta = __reduce_add_sync(FULL, delta);
    __syncwarp();                         // every lane has read the record before lane 0 rewrites it
    if (lane == 0) {
        st->count_points += delta;
        st->static_goals = (uint8_t)sg;
        if (state.lo != state_lo0) {   // the LCG is a bijection: lo changes iff draws happened
            st->rng[0] = state.hi;
            st->rng[1] = state.lo;
        }
    }
    return goals_dirty;
}

// ---- phase B: scalar tail of the step for up to 8 envs, lane w = env w --------
// update_exit_colors, reward / done, counters (safelife_env.py:153-175), then
// the auto-reset of finished envs by the whole warp.  `goals0`: goals slab of the
// group's first env (shared memory or HBM).  Returns true if a reset rewrote a
// goals slab.
__device__ __forceinline__ bool step_phase_b(const SlEnvBatch &B, const CtaShared &cs, int e0,
                                             int nenv, SlEnvState *sstate, uint16_t *sboard,
                                             uint16_t *goals0, int flags, int lane,
                                             uint32_t host_tag = 0) {
    const int cells = B.H * B.W;
    bool done = false;
    if (lane < nenv) {
        SlEnvState *s0 = sstate + lane;
        uint16_t *bd = sboard + lane * cells;
        const uint16_t *gl = goals0 + (size_t)lane * cells;
        const int lvl = s0->level_id;
        bool agent_alive, exited;
        int dmem;
        const int row = s0->loc_row, col = s0->loc_col, pts = s0->count_points;
        const int value = exit_update(bd, gl, B.W, row, col, pts, *s0,
                                      B.pool_exits + (size_t)lvl * B.max_exits,
                                      cs.table ? cs.table : B.pool_table + (size_t)lvl * 72,
                                      &agent_alive, &exited, &dmem);
        done = finish_step(B, e0 + lane, *s0, row, col, value, agent_alive, exited,
                           s0->static_goals, pts + dmem, nullptr, host_tag);
    }
    bool any_reset = false;
    if (flags & SL_STEP_AUTO_RESET) {
        uint32_t todo = __ballot_sync(FULL, done);
        any_reset = todo != 0;
        while (todo != 0) {
            const int w = __ffs(todo) - 1;
            todo &= todo - 1;
            warp_reset_env(B, sstate + w, sboard + w * cells, goals0 + (size_t)w * cells, lane);
        }
    }
    return any_reset;
}

// ---- the same tail for ONE env on its own warp ---------------------------------
// Lane 0 runs update_exit_colors / reward / done / counters of the env the warp
// just advanced, then the whole warp resets it if it finished.  No warp waits
// for another one: the single-pass kernel uses this instead of a CTA barrier
// followed by phase B on warp 0, and the kernels that write the observation use
// it because the observation must follow the tail.  Returns true if a reset
// rewrote the goals slab.
// (in three pieces, because the tiled kernels run the middle one on all the warps
// of an env)
// 1. lane 0: update_exit_colors / reward / done / counters.  Returns done (warp-uniform).
__device__ __forceinline__ int step_tail_scalar(const SlEnvBatch &B, const CtaShared &cs, int e,
                                                SlEnvState *st, uint16_t *board,
                                                const uint16_t *goals, int lane, StepOut *so) {
    int done = 0;
    *so = {0.f, false, false};
    __syncwarp();                         // the board cells phase A rewrote
    if (lane == 0) {
        const int lvl = st->level_id;
        bool agent_alive, exited;
        int dmem;
        const int row = st->loc_row, col = st->loc_col, pts = st->count_points;
        const int value = exit_update(board, goals, B.W, row, col, pts, *st,
                                      B.pool_exits + (size_t)lvl * B.max_exits,
                                      cs.table ? cs.table : B.pool_table + (size_t)lvl * 72,
                                      &agent_alive, &exited, &dmem);
        done = finish_step(B, e, *st, row, col, value, agent_alive, exited, st->static_goals,
                           pts + dmem, so) ? 1 : 0;
    }
    return __shfl_sync(FULL, done, 0);
}

// 3. the reset of a finished env by the whole warp; returns true if it happened.
__device__ __forceinline__ bool step_tail_reset(const SlEnvBatch &B, const SlWrapSpec &Wp, int e,
                                                SlEnvState *st, uint16_t *board, uint16_t *goals,
                                                int flags, int done, int lane) {
    bool reset = false;
    if (done && (flags & SL_STEP_AUTO_RESET)) {
        warp_reset_env(B, st, board, goals, lane);
        if (Wp.wstate != nullptr) warp_wrap_reset(B, Wp, e, st, board, lane);
        reset = true;
    }
    __syncwarp();
    return reset;
}

__device__ __forceinline__ bool step_tail_env(const SlEnvBatch &B, const SlWrapSpec &Wp,
                                              const CtaShared &cs, int e, SlEnvState *st,
                                              uint16_t *board, uint16_t *goals, int flags, int lane) {
    StepOut so;
    const int done = step_tail_scalar(B, cs, e, st, board, goals, lane, &so);
    __syncwarp();                         // lane 0's exit colours before the other lanes read the board
    // 2. the training wrappers see the final state of the step, before any reset (wrap.cuh)
    if (Wp.wstate != nullptr)
        warp_wrap_step(B, Wp, e, st, board, goals, so.reward, so.done, so.times_up, lane);
    return step_tail_reset(B, Wp, e, st, board, goals, flags, done, lane);
}

// The observation of one env as the step leaves it (see obs.cuh); out of line so
// that its registers do not weigh on the stencil.  The board slab is in shared
// memory, the goals slab too when resident (GR).
template <bool GR>
__device__ __noinline__ void step_write_obs(const SlEnvBatch &B, const ObsParams &O, int e,
                                            const SlEnvState *st, const uint16_t *board,
                                            const uint16_t *goals, unsigned char *stage, int lane) {
    warp_write_obs_any<true, GR>(B, O, e, st, board, goals, stage, lane, 0, 1, [] { __syncwarp(); });
}

// SafeLifeEnv.step, one warp per env (safelife_env.py:148-201): single-pass
// variant, one CTA per group of 8 envs.  Warp w runs the whole step of env w --
// execute_actions and advance_board on bit planes (phase A), the scalar tail on
// lane 0, the reset of a finished env, and (OBS) the observation -- without
// waiting for the other warps; one CTA barrier before the bulk stores.  The
// observation (O.obs != NULL) and the training-wrapper epilogue (Wp.wstate !=
// NULL) are optional parts of the same launch.
// The board slabs (and the goal slabs, when resident) arrive on separate
// mbarriers, so the stencil starts while the goals are still in flight.
// Ragged / unaligned batches fall back to scalar copies instead of bulk copies.
template <class G, bool GR, bool PW>
__global__ void __launch_bounds__(WARP_ENVS_PER_CTA * 32, G::ROWS == 1 ? 4 : 2)
k_env_step_warp(const __grid_constant__ SlEnvBatch B, const __grid_constant__ ObsParams O,
                const __grid_constant__ SlWrapSpec Wp, const uint8_t *__restrict__ actions,
                int bulk_ok, int flags, int first_env, int epc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int cells = B.H * B.W;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // CTA b steps envs [e0, e0 + nenv): epc (1..8) consecutive envs, one warp each.  The
    // host picks epc so that every SM gets the same number of warps (plan_env_step_warp).
    const int e0 = first_env + blockIdx.x * epc;
    const int nenv = min(epc, B.n_envs - e0);
    const int group_bytes = nenv * cells * 2, state_bytes = nenv * 96;
    // Bulk copies move whole 16-byte units.  The CTA's slabs start `mis` bytes into one
    // (the arrays themselves are 16-byte aligned when bulk_ok), so the loads fetch the
    // enclosing units -- a few cells of the neighbouring CTAs' envs come along and are
    // never touched -- and the shared-memory slabs keep the offset; the stores write the
    // whole units inside the CTA's range in bulk and the (< 16-byte) edges cell by cell.
    const size_t byte0 = (size_t)e0 * cells * 2;
    const int mis = bulk_ok ? (int)(byte0 & 15) : 0;
    const int load_bytes = (mis + group_bytes + 15) & ~15;
    const bool bulk = bulk_ok && group_bytes >= 64 &&
                      byte0 - mis + load_bytes <= (((size_t)B.n_envs * cells * 2) & ~(size_t)15);
    uint64_t *bar = (uint64_t *)smem_raw;             // [0] state + boards, [1] goals
    int32_t *stable = (int32_t *)(smem_raw + WARP_SMEM_HEAD);
    SlEnvState *sstate = (SlEnvState *)(smem_raw + WARP_SMEM_HEAD + WARP_SMEM_TABLE);
    unsigned char *rboard = smem_raw + WARP_SMEM_HEAD + WARP_SMEM_TABLE + WARP_SMEM_STATE;
    unsigned char *rgoals = rboard + warp_slab_stride_any(cells);      // GR: goal slabs; else: the scratch slab
    uint16_t *sboard = (uint16_t *)(rboard + (bulk ? mis : 0));
    uint16_t *sgoals = (uint16_t *)(rgoals + (bulk && GR ? mis : 0));
    uint16_t *gboard = B.board + (size_t)e0 * cells;
    uint16_t *ggoals = B.goals + (size_t)e0 * cells;
    SlEnvState *gstate = B.state + e0;
    pdl_launch_dependents();
    SL_TRACE_MARK(threadIdx.x == 0, 0);

    CtaShared cs;
    cs.scratch = (uint16_t *)rgoals;
    cs.lock = (int *)(smem_raw + 16);
    if (bulk) {
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            mbar_init(bar + 1, 1);
            *cs.lock = 0;
            pdl_wait();                   // nothing written by the launch before this one is read earlier
            mbar_expect_tx(bar, (uint32_t)(load_bytes + state_bytes));
            bulk_load(sstate, gstate, state_bytes, bar);
            bulk_load(rboard, (const unsigned char *)gboard - mis, load_bytes, bar);
            if (GR) {
                mbar_expect_tx(bar + 1, (uint32_t)load_bytes);
                bulk_load(rgoals, (const unsigned char *)ggoals - mis, load_bytes, bar + 1);
            }
        }
    } else {
        pdl_wait();
        if (threadIdx.x == 0) *cs.lock = 0;
        copy_slabs(sstate, gstate, state_bytes);
        copy_slabs(sboard, gboard, group_bytes);
        if (GR) copy_slabs(sgoals, ggoals, group_bytes);
    }
    pdl_wait();
    cs.table = stage_table(B, stable);
    // the CTA's actions with one or two aligned 8-byte loads (PCIe reads when they live in
    // pinned host memory), handed out through shared memory
    unsigned char *sact = smem_raw + 32;
    const int a0 = e0 & ~7;
    const bool act8 = (((uintptr_t)actions) & 7) == 0 && ((e0 + nenv + 7) & ~7) <= B.n_envs;
    if (act8 && warp == 0 && lane >= 1 && lane <= 2 && a0 + 8 * (lane - 1) < e0 + nenv)
        ((uint64_t *)sact)[lane - 1] = *(const uint64_t *)(actions + a0 + 8 * (lane - 1));
    int action = (!act8 && warp < nenv) ? actions[e0 + warp] : 0;
    __syncthreads();                      // mbarrier init / fallback copy / table / actions visible
    if (act8 && warp < nenv) action = sact[e0 - a0 + warp];
    if (bulk) mbar_wait(bar, 0);
    SL_TRACE_MARK(threadIdx.x == 0, 1);

    bool goals_dirty = false;
    if (warp < nenv) {
        SlEnvState *st = sstate + warp;
        uint16_t *bd = sboard + warp * cells;
        uint16_t *gl = GR ? sgoals + warp * cells : ggoals + (size_t)warp * cells;
        goals_dirty = step_phase_a<G, GR>(B, cs, G(B.H, B.W, lane), st, bd, gl,
                                          GR && bulk ? bar + 1 : nullptr, action);
        SL_TRACE_MARK(threadIdx.x == 0, 2);
        if (PW) {
            goals_dirty |= step_tail_env(B, Wp, cs, e0 + warp, st, bd, gl, flags, lane);
            SL_TRACE_MARK(threadIdx.x == 0, 3);
            if (O.obs != nullptr) {
                const int stage_bytes = obs_stage_bytes(O.vh, O.vw, O.n_channels);
                unsigned char *stage = smem_raw + warp_step_smem_bytes<G>(cells, GR) + warp * stage_bytes;
                step_write_obs<GR>(B, O, e0 + warp, st, bd, gl, stage, lane);
            }
        }
    } else if (GR && bulk) {
        mbar_wait(bar + 1, 0);
    }
    fence_async_smem();                   // generic-proxy writes -> visible to the bulk engine
    SL_TRACE_MARK(threadIdx.x == 0, 4);
    int any_goals_dirty = __syncthreads_or(goals_dirty);
    if (warp != 0) return;
    if (!PW) {
        // the scalar tails of the CTA's envs on lanes of warp 0 (cheaper in issue slots
        // than one tail per warp when nothing else follows the tail)
        any_goals_dirty |= step_phase_b(B, cs, e0, nenv, sstate, sboard, GR ? sgoals : ggoals, flags, lane);
        fence_async_smem();
        __syncwarp();
    }
    if (bulk) {
        // whole 16-byte units in bulk, the edges (shared with the neighbouring CTAs) by cell
        const int head = (16 - mis) & 15;
        const int body = (group_bytes - head) & ~15, tail = group_bytes - head - body;
        if (lane == 0) {
            bulk_store((unsigned char *)gboard + head, (unsigned char *)sboard + head, body);
            if (GR && any_goals_dirty)
                bulk_store((unsigned char *)ggoals + head, (unsigned char *)sgoals + head, body);
            bulk_store(gstate, sstate, state_bytes);
        }
        const int edge = lane < 8 ? lane : (head + body) / 2 + lane - 8;     // lanes 0-7: head, 8-15: tail
        if (lane < 16 && (lane < 8 ? 2 * lane < head : 2 * (lane - 8) < tail)) {
            gboard[edge] = sboard[edge];
            if (GR && any_goals_dirty) ggoals[edge] = sgoals[edge];
        }
        if (lane == 0) bulk_commit_and_wait_read();
    } else {
        for (int i = lane; i < (group_bytes >> 1); i += 32) gboard[i] = sboard[i];
        if (GR && any_goals_dirty)
            for (int i = lane; i < (group_bytes >> 1); i += 32) ggoals[i] = sgoals[i];
        for (int i = lane; i < (state_bytes >> 1); i += 32)
            ((uint16_t *)gstate)[i] = ((const uint16_t *)sstate)[i];
    }
    SL_TRACE_MARK(lane == 0, 5);
}

// ---- persistent, warp-specialised, multi-stage variant (narrow boards) ----------
// One CTA per SM slot loops over groups of 8 envs (group g, g + grid, ...).
// Warps 0-7 are stencil warps (phase A of env w of the current group); warp 8
// is the scalar warp: it issues the bulk loads STAGES groups ahead, waits for
// the 8 stencil warps of a group on an mbarrier, runs phase B for the group's 8
// envs on 8 lanes, and issues the bulk stores.  Stencil warps never wait for
// each other: the TMA head and the scalar tail of one group overlap the
// stencil work of the next.  With in-place goals a stage is half the size, so
// twice as many groups are in flight per SM.
constexpr int PIPE_MAX_STAGES = 8;
constexpr uint32_t PIPE_PARK_NS = 20000;
constexpr int PIPE_THREADS = (WARP_ENVS_PER_CTA + 1) * 32;
// full[8], fin[8] mbarriers, dirty flags [8][8], lock, pad; then the points table
constexpr int PIPE_SMEM_HEAD = 2 * 8 * PIPE_MAX_STAGES + 8 * PIPE_MAX_STAGES + 16;

__host__ __device__ inline int pipe_stage_bytes(int cells, bool goals_resident) {
    return WARP_SMEM_STATE + (goals_resident ? 2 : 1) * warp_slab_stride(cells);
}

__host__ __device__ inline size_t pipe_smem_bytes(int cells, int stages, bool goals_resident,
                                                  int obs_stage = 0) {
    return PIPE_SMEM_HEAD + WARP_SMEM_TABLE + stages * (size_t)pipe_stage_bytes(cells, goals_resident) +
           (goals_resident ? 0 : warp_scratch_bytes(cells)) + 2 * Narrow::PAD_CELLS +
           (size_t)WARP_ENVS_PER_CTA * obs_stage;
}

// PW (per-warp tail; used when the step also writes the observation or runs
// the wrapper epilogue): every stencil warp also runs the tail of its env
// (step_tail_env) and writes its observation (step_write_obs) before it reports
// the stage finished; the scalar warp is then only the bulk-copy engine's driver.
template <int STAGES, int CTAS_PER_SM, bool GR, bool PW>
__global__ void __launch_bounds__(PIPE_THREADS, CTAS_PER_SM)
k_env_step_pipe(const __grid_constant__ SlEnvBatch B, const __grid_constant__ ObsParams O,
                const __grid_constant__ SlWrapSpec Wp, const uint8_t *__restrict__ actions,
                int flags, int n_groups) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int cells = B.H * B.W;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int group_bytes = WARP_ENVS_PER_CTA * cells * 2;
    uint64_t *full = (uint64_t *)smem_raw;            // [STAGES] loads of a stage landed
    uint64_t *fin = full + PIPE_MAX_STAGES;           // [STAGES] the 8 stencil warps are through
    volatile uint8_t *dirty = (volatile uint8_t *)(fin + PIPE_MAX_STAGES);   // [STAGES][8] goals rewritten
    int32_t *stable = (int32_t *)(smem_raw + PIPE_SMEM_HEAD);
    unsigned char *stages = smem_raw + PIPE_SMEM_HEAD + WARP_SMEM_TABLE;
    const int stage_bytes = pipe_stage_bytes(cells, GR);
    auto stage_state = [&](int s) { return (SlEnvState *)(stages + s * stage_bytes); };
    auto stage_board = [&](int s) { return (uint16_t *)(stages + s * stage_bytes + WARP_SMEM_STATE); };
    auto stage_goals = [&](int s) {
        return (uint16_t *)(stages + s * stage_bytes + WARP_SMEM_STATE + warp_slab_stride(cells));
    };
    auto issue_load = [&](int s, int g) {
        const size_t e0 = (size_t)g * WARP_ENVS_PER_CTA;
        mbar_expect_tx(full + s, (GR ? 2u : 1u) * group_bytes + WARP_SMEM_STATE);
        bulk_load(stage_state(s), B.state + e0, WARP_SMEM_STATE, full + s);
        bulk_load(stage_board(s), B.board + e0 * cells, group_bytes, full + s);
        if (GR) bulk_load(stage_goals(s), B.goals + e0 * cells, group_bytes, full + s);
    };
    CtaShared cs;
    cs.scratch = (uint16_t *)(stages + STAGES * stage_bytes);
    cs.lock = (int *)(smem_raw + PIPE_SMEM_HEAD - 16);

    pdl_launch_dependents();
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(full + s, 1);
            mbar_init(fin + s, WARP_ENVS_PER_CTA);
        }
        *cs.lock = 0;
    }
    pdl_wait();
    cs.table = stage_table(B, stable);
    __syncthreads();
    const int first = blockIdx.x, stride = gridDim.x;

    if (warp == WARP_ENVS_PER_CTA) {
        // ---- scalar warp: loads, phase B, stores
        if (lane == 0)
            for (int s = 0; s < STAGES; s++)
                if (first + s * stride < n_groups) issue_load(s, first + s * stride);
        int s = 0, parity = 0;
        for (int g = first; g < n_groups; g += stride) {
            mbar_wait_parked(fin + s, parity, PIPE_PARK_NS);
            const int e0 = g * WARP_ENVS_PER_CTA;
            bool goals_dirty = GR && lane < WARP_ENVS_PER_CTA && dirty[s * WARP_ENVS_PER_CTA + lane] != 0;
            if (!PW)
                goals_dirty |= step_phase_b(B, cs, e0, WARP_ENVS_PER_CTA, stage_state(s), stage_board(s),
                                            GR ? stage_goals(s) : B.goals + (size_t)e0 * cells, flags, lane);
            const bool any_dirty = __any_sync(FULL, goals_dirty);
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                bulk_store(B.board + (size_t)e0 * cells, stage_board(s), group_bytes);
                if (GR && any_dirty) bulk_store(B.goals + (size_t)e0 * cells, stage_goals(s), group_bytes);
                bulk_store(B.state + e0, stage_state(s), WARP_SMEM_STATE);
                bulk_commit_and_wait_read();       // the stage may be overwritten after this
                const int gn = g + STAGES * stride;
                if (gn < n_groups) issue_load(s, gn);
            }
            __syncwarp();
            if (++s == STAGES) {
                s = 0;
                parity ^= 1;
            }
        }
    } else {
        // ---- stencil warp w: phase A of env w of every group of this CTA
        const Narrow t(B.H, B.W, lane);                 // row masks, neighbour lanes: once per warp
        const int env_bytes = cells * 2;
        unsigned char *my = stages + warp * env_bytes;  // this warp's slot inside a stage
        unsigned char *obs_stage = smem_raw + pipe_smem_bytes(cells, STAGES, GR) +
                                   (PW ? warp * obs_stage_bytes(O.vh, O.vw, O.n_channels) : 0);
        int s = 0, parity = 0;
        int next_action = first < n_groups ? actions[(size_t)first * WARP_ENVS_PER_CTA + warp] : 0;
        for (int g = first; g < n_groups; g += stride) {
            const size_t e = (size_t)g * WARP_ENVS_PER_CTA + warp;
            const int action = next_action;             // fetched one group ahead
            if (g + stride < n_groups)
                next_action = actions[(size_t)(g + stride) * WARP_ENVS_PER_CTA + warp];
            unsigned char *stage = my + s * stage_bytes;
            mbar_wait_parked(full + s, parity, PIPE_PARK_NS);
            SlEnvState *st = (SlEnvState *)(stages + s * stage_bytes) + warp;
            uint16_t *bd = (uint16_t *)(stage + WARP_SMEM_STATE);
            uint16_t *gl = GR ? (uint16_t *)(stage + WARP_SMEM_STATE + warp_slab_stride(cells))
                              : B.goals + e * cells;
            bool gd = step_phase_a<Narrow, GR>(B, cs, t, st, bd, gl, nullptr, action);
            if (PW) {
                gd |= step_tail_env(B, Wp, cs, (int)e, st, bd, gl, flags, lane);
                if (O.obs != nullptr) step_write_obs<GR>(B, O, (int)e, st, bd, gl, obs_stage, lane);
            }
            if (GR && lane == 0) dirty[s * WARP_ENVS_PER_CTA + warp] = gd;
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(fin + s);
            if (++s == STAGES) {
                s = 0;
                parity ^= 1;
            }
        }
    }
}

// Pipeline shapes: (stages, CTAs per SM) by goals mode.  The stage buffers of a
// 25x25 board fill the shared memory of an SM in every shape.
struct PipeShape {
    int stages, ctas;
};

inline int env_knob(const char *name, int fallback) {
    const char *v = getenv(name);
    return v ? atoi(v) : fallback;
}

// Launch with programmatic stream serialization (see pdl_wait above);
// SAFELIFE_B200_PDL=0 launches plainly (A/B knob).
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, int threads, size_t smem,
                              cudaStream_t st, const Args &...args) {
    static const bool pdl = env_knob("SAFELIFE_B200_PDL", 1) != 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}


// The shapes compiled in; `shape` picks one (tuning knob SAFELIFE_B200_PIPE_SHAPE).
#define SL_PIPE_SHAPES(X) X(2, 4, true) X(3, 3, true) X(3, 4, false) X(4, 4, false) X(5, 3, false) X(6, 3, false)

inline bool pipe_shape_fits(int cells, int stages, int ctas, bool gr, int smem_per_sm,
                            int obs_stage = 0) {
    return ctas * (pipe_smem_bytes(cells, stages, gr, obs_stage) + 1024) <= (size_t)smem_per_sm;
}

// What sl_env_step / sl_env_step_obs launches for a batch on the warp path.
struct StepPlan {
    bool narrow, goals_resident, obs, wrapped;
    int bulk_ok;
    int obs_stage;                  // staging bytes per warp of the fused observation (0: none)
    int pipe_stages, pipe_ctas;     // 0, 0: no pipelined launch
    int pipe_grid, piped_groups;    // persistent CTAs and the groups (of 8 envs) they cover
    int single_groups;              // CTAs of the single-pass kernel,
    int single_epc;                 // its envs per CTA (1..8), from env piped_groups * 8 on
};

// O: the observation the step writes, or nullptr.
inline StepPlan plan_env_step_warp(const SlEnvBatch &B, const ObsParams *O, const SlWrapSpec *Wp,
                                   int sm_count, int smem_per_sm, bool prefer_pipe = false) {
    StepPlan p = {};
    const int cells = B.H * B.W;
    const int groups = (B.n_envs + WARP_ENVS_PER_CTA - 1) / WARP_ENVS_PER_CTA;
    const int full_groups = B.n_envs / WARP_ENVS_PER_CTA;
    p.bulk_ok = (((uintptr_t)B.board | (uintptr_t)B.goals | (uintptr_t)B.state) & 15) == 0;
    static const bool in_place_ok = env_knob("SAFELIFE_B200_GOALS_IN_PLACE", 1) != 0;
    p.obs = O != nullptr;
    // in place only pays while nobody needs every goal cell: the observation does, and
    // then the goal slabs are better brought in by the bulk-copy engine with the boards
    p.goals_resident = !((B.pool_flags & SL_POOL_STATIC_GOALS) && in_place_ok) || p.obs;
    p.narrow = warp_kernel_is_narrow(B.H, B.W);
    p.wrapped = Wp != nullptr;
    p.obs_stage = O ? obs_stage_bytes(O->vh, O->vw, O->n_channels) : 0;
    p.single_groups = groups;
    p.single_epc = WARP_ENVS_PER_CTA;
    if (!p.narrow) return p;
    // pipelining pays once a CTA slot sees more than one group
    static const int shape_knob = env_knob("SAFELIFE_B200_PIPE_SHAPE", 0);   // stages * 10 + ctas
    int want = shape_knob ? shape_knob : (p.goals_resident ? 33 : 53);
    bool known = false;
#define SL_KNOWN_SHAPE(S, C, R)                                                        \
    known |= R == p.goals_resident && want == S * 10 + C &&                            \
             pipe_shape_fits(cells, S, C, R, smem_per_sm, p.obs_stage);
    SL_PIPE_SHAPES(SL_KNOWN_SHAPE)
#undef SL_KNOWN_SHAPE
    if (!known) want = p.goals_resident ? 24 : 34;    // a large board or view: the smallest shape
    const int stages = want / 10, ctas = want % 10;
    if (!pipe_shape_fits(cells, stages, 1, p.goals_resident, smem_per_sm, p.obs_stage)) return p;
    // (measured crossover: the single-pass kernel wins up to about two waves of its 4 CTAs
    // per SM -- 8192 envs: 18.4 vs 19.2 us, 16384 envs: 27.2 vs 25.6 us)
    static const int min_waves = env_knob("SAFELIFE_B200_PIPE_MIN_WAVES", 2);   // tuning knob
    // With the observation in the launch the single-pass kernel wins at every batch size
    // (28-32 warps per SM against the pipeline's 24 stencil warps; rotating batches, us per
    // step: 16384 envs 58.9 vs 70.6, 32768 envs 109 vs 136, 65536 envs 215 vs 255), so the
    // pipeline then runs only on request (kernel 3).
    const bool pipe_pays = prefer_pipe || !p.obs;
    if (p.bulk_ok && pipe_pays && full_groups > sm_count * 4 * (prefer_pipe ? 1 : min_waves)) {
        const int slots = sm_count * ctas;
        const int per_cta = (full_groups + slots - 1) / slots;      // same trip count everywhere
        p.pipe_stages = stages;
        p.pipe_ctas = ctas;
        p.pipe_grid = (full_groups + per_cta - 1) / per_cta;
        p.piped_groups = full_groups;
        p.single_groups = groups - full_groups;
    } else {
        // single pass: as many envs per CTA (= warps) as spread the batch evenly over the
        // SMs' CTA slots -- 4096 envs are 586 CTAs of 7 (28 warps on every SM) rather than
        // 512 CTAs of 8 (32 warps on some SMs, 24 on the others)
        static const int epc_knob = env_knob("SAFELIFE_B200_EPC", 0);          // A/B knob
        const int slots = sm_count * 4;
        const int waves = (B.n_envs + slots * WARP_ENVS_PER_CTA - 1) / (slots * WARP_ENVS_PER_CTA);
        int epc = (B.n_envs + slots * waves - 1) / (slots * waves);
        if (epc_knob) epc = epc_knob;
        epc = epc < 1 ? 1 : (epc > WARP_ENVS_PER_CTA ? WARP_ENVS_PER_CTA : epc);
        p.single_epc = epc;
        p.single_groups = (B.n_envs + epc - 1) / epc;
    }
    return p;
}

// Launches the plan.  `prep(kernel, dynamic shared memory bytes)` is called before
// each launch (the ABI layer opts the kernel in to large shared memory there) and
// returns non-zero to abort.
template <class Prep>
inline int launch_env_step_warp(const SlEnvBatch &B, const ObsParams *O, const SlWrapSpec *Wp,
                                const uint8_t *actions, int flags, const StepPlan &p,
                                cudaStream_t st, Prep prep) {
    const int cells = B.H * B.W;
    const int threads = p.single_epc * 32;
    static const ObsParams no_obs = {};
    static const SlWrapSpec no_wrap = {};
    const ObsParams &obs = O ? *O : no_obs;
    const SlWrapSpec &wrap = Wp ? *Wp : no_wrap;
    const bool per_warp = p.obs || p.wrapped;
    int rc = 0;
#define SL_LAUNCH_WARP_PW(G, R, PW)                                                            \
    do {                                                                                       \
        const size_t smem = warp_step_smem_bytes<G>(cells, R, p.obs_stage);                    \
        rc = prep((const void *)k_env_step_warp<G, R, PW>, smem);                              \
        if (rc) return rc;                                                                     \
        launch_pdl(k_env_step_warp<G, R, PW>, p.single_groups, threads, smem, st,              \
                   B, obs, wrap, actions, p.bulk_ok, flags,                                    \
                   p.piped_groups * WARP_ENVS_PER_CTA, p.single_epc);                          \
    } while (0)
#define SL_LAUNCH_WARP(G, R)                                                                   \
    do {                                                                                       \
        if (per_warp) SL_LAUNCH_WARP_PW(G, R, true); else SL_LAUNCH_WARP_PW(G, R, false);      \
    } while (0)
    if (!p.narrow) {
        if (p.goals_resident) SL_LAUNCH_WARP(Wide, true); else SL_LAUNCH_WARP(Wide, false);
        return 0;
    }
    if (p.piped_groups > 0) {
#define SL_LAUNCH_SHAPE_PW(S, C, R, PW)                                                        \
        if (R == p.goals_resident && S == p.pipe_stages && C == p.pipe_ctas && PW == per_warp) { \
            const size_t smem = pipe_smem_bytes(cells, S, R, p.obs_stage);                     \
            rc = prep((const void *)k_env_step_pipe<S, C, R, PW>, smem);                       \
            if (rc) return rc;                                                                 \
            launch_pdl(k_env_step_pipe<S, C, R, PW>, p.pipe_grid, PIPE_THREADS, smem, st,      \
                       B, obs, wrap, actions, flags, p.piped_groups);                          \
        }
#define SL_LAUNCH_SHAPE(S, C, R) SL_LAUNCH_SHAPE_PW(S, C, R, false) SL_LAUNCH_SHAPE_PW(S, C, R, true)
        SL_PIPE_SHAPES(SL_LAUNCH_SHAPE)
#undef SL_LAUNCH_SHAPE
#undef SL_LAUNCH_SHAPE_PW
    }
    if (p.single_groups > 0) {
        if (p.goals_resident) SL_LAUNCH_WARP(Narrow, true); else SL_LAUNCH_WARP(Narrow, false);
    }
#undef SL_LAUNCH_WARP
#undef SL_LAUNCH_WARP_PW
    return 0;
}

// ------------------------------------------------------------------------
// n-step advance / life occupancy, one warp per board (H, W <= 32), the whole
// loop in registers: advance_board_nstep (advance_board.c:128-149) and
// life_occupancy (:153-189).  The planes live in registers for all n_steps;
// occupancy counts are bit-sliced counters (one 10-bit vertical counter per
// colour and cell), spilled into the int32 output every 1023 steps.  Nothing
// touches memory inside the step loop.
// ------------------------------------------------------------------------
constexpr int ADV_WARPS = 4;
constexpr int ADV_SLAB_CELLS = 32 * 32 + 32;     // one padded board per warp
constexpr int OCC_BITS = 10;

inline bool warp_advance_supports(int H, int W) {
    return warp_kernel_supports(H, W) && warp_kernel_is_narrow(H, W);
}

// Register image of one board (H, W <= 32) for the multi-step loops: the ten
// planes the rule reads or writes, the AGENT plane (occupancy mask) and the
// mask of cells that have not changed since the board was loaded.
struct AdvBoard {
    LowPlanesT<uint32_t> lp;
    HighPlanesT<uint32_t> hp;
    uint32_t agent, kept;
};

// slab: this lane's padded copy of the board in shared memory.
__device__ __forceinline__ AdvBoard adv_load(const Narrow &t, const uint16_t *slab) {
    uint32_t x[16];
    const uint16_t *row = slab + t.row_off;
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = __byte_perm(row[k], row[k + 16], 0x5410);
    transpose16(x);
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] &= t.wmask;
    AdvBoard b;
    b.lp = {x[0], x[3], x[4], x[5], x[6], x[7]};
    b.hp = {x[8], x[9], x[10], x[11]};
    b.agent = x[1];
    b.kept = t.wmask;              // cells that never changed keep their word
    return b;
}

// One advance_board step (advance_board.c:34-125) on the register image.
__device__ __forceinline__ void adv_step(const Narrow &t, AdvBoard &b, u128 &state,
                                         const SpawnCtx &ctx, uint64_t thr, bool draws) {
    LowPlanesT<uint32_t> &lp = b.lp;
    HighPlanesT<uint32_t> &hp = b.hp;
    const RowDeltaT<uint32_t> d = stencil_planes(t, lp, [&] { return hp; });
    uint32_t spawned = 0u;
    if (draws && __any_sync(FULL, d.cand != 0)) spawned = warp_spawn_draws_ctx(t, state, ctx, thr, d.cand);
    const uint32_t fresh = d.born | spawned, changed = d.die | fresh, keep = ~changed;
    lp.A = (lp.A & keep) | fresh;
    lp.D = (lp.D & keep) | (d.dn & d.born) | spawned;
    hp.R = (hp.R & keep) | (d.tR & fresh);
    hp.G = (hp.G & keep) | (d.tG & fresh);
    hp.B = (hp.B & keep) | (d.tB & fresh);
    lp.P_ &= keep; lp.I &= keep; lp.S &= keep; hp.E &= keep; b.agent &= keep;
    b.kept &= keep;                     // (F is already 0 on every changed cell)
}

// Cells that changed at least once are fully described by the planes: patch
// them into the lane's row of the slab.
__device__ __forceinline__ void adv_store(const Narrow &t, const AdvBoard &b, uint16_t *slab) {
    __syncwarp();                         // a lane's load window reaches into the next lane's row
    uint16_t *row = slab + t.lane * t.W;
    uint32_t m = t.wmask & ~b.kept;
    while (m != 0) {
        const int j = __ffs(m) - 1;
        m &= m - 1;
        row[j] = (uint16_t)(((b.lp.A >> j) & 1u) | (((b.lp.D >> j) & 1u) << 3) |
                            (((b.hp.R >> j) & 1u) << 9) | (((b.hp.G >> j) & 1u) << 10) |
                            (((b.hp.B >> j) & 1u) << 11));
    }
}

// Bit-sliced occupancy counters of one board row: a 10-bit vertical counter
// per colour and cell (advance_board.c:153-162 adds one per step).
struct OccCounters {
    uint32_t cnt[8][OCC_BITS];
    int pending;

    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int c = 0; c < 8; c++)
#pragma unroll
            for (int k = 0; k < OCC_BITS; k++) cnt[c][k] = 0;
        pending = 0;
    }
    // counters -> dst[row][col][colour] (+=), dst = counts of this board
    __device__ __forceinline__ void flush(const Narrow &t, int32_t *board_counts) {
        if (t.lane < t.H) {
            int32_t *dst = board_counts + (size_t)t.lane * t.W * 8;
#pragma unroll
            for (int c = 0; c < 8; c++) {
                uint32_t nz = 0;
#pragma unroll
                for (int k = 0; k < OCC_BITS; k++) nz |= cnt[c][k];
                while (nz != 0) {
                    const int j = __ffs(nz) - 1;
                    nz &= nz - 1;
                    int v = 0;
#pragma unroll
                    for (int k = 0; k < OCC_BITS; k++) v |= ((cnt[c][k] >> j) & 1u) << k;
                    dst[j * 8 + c] += v;
                }
#pragma unroll
                for (int k = 0; k < OCC_BITS; k++) cnt[c][k] = 0;
            }
        }
        pending = 0;
    }
    // +1 for cells ALIVE and none of AGENT | EXIT | FROZEN, by colour (:153-162)
    __device__ __forceinline__ void add(const Narrow &t, const AdvBoard &b, int32_t *board_counts) {
        const uint32_t occ = b.lp.A & ~(b.agent | b.hp.E | b.lp.F);
        if (__any_sync(FULL, occ != 0)) {
#pragma unroll
            for (int c = 0; c < 8; c++) {
                uint32_t carry = occ & ((c & 1) ? b.hp.R : ~b.hp.R) & ((c & 2) ? b.hp.G : ~b.hp.G) &
                                 ((c & 4) ? b.hp.B : ~b.hp.B);
                if (__any_sync(FULL, carry != 0)) {
#pragma unroll
                    for (int k = 0; k < OCC_BITS; k++) {
                        const uint32_t tmp = cnt[c][k] & carry;
                        cnt[c][k] ^= carry;
                        carry = tmp;
                    }
                }
            }
        }
        if (++pending == (1 << OCC_BITS) - 1) flush(t, board_counts);
    }
};

__global__ void __launch_bounds__(ADV_WARPS * 32)
k_advance_warp(const uint16_t *in, uint16_t *out, int32_t *__restrict__ counts, int n, int H,
               int W, uint64_t thr, int n_steps, uint64_t *__restrict__ rng) {
    __shared__ uint16_t slabs[ADV_WARPS][ADV_SLAB_CELLS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cells = H * W;
    uint16_t *slab = slabs[warp];
    const Narrow t(H, W, lane);
    for (int b = blockIdx.x * ADV_WARPS + warp; b < n; b += gridDim.x * ADV_WARPS) {
        const uint16_t *src = in + (size_t)b * cells;
        for (int i = lane; i < cells; i += 32) slab[i] = src[i];
        __syncwarp();
        AdvBoard bd = adv_load(t, slab);
        u128 state = {0, 0}, inc = {0, 0};
        if (rng != nullptr) {
            state = {rng[4 * (size_t)b], rng[4 * (size_t)b + 1]};
            inc = {rng[4 * (size_t)b + 2], rng[4 * (size_t)b + 3]};
        }
        const SpawnCtx ctx = make_spawn_ctx(lane, inc);
        if (counts != nullptr) {
            int32_t *dst = counts + (size_t)b * cells * 8;
            OccCounters occ;
            occ.clear();
            for (int s = 0; s < n_steps; s++) {
                adv_step(t, bd, state, ctx, thr, rng != nullptr);
                occ.add(t, bd, dst);
            }
            if (occ.pending) occ.flush(t, dst);
        } else {
            for (int s = 0; s < n_steps; s++) adv_step(t, bd, state, ctx, thr, rng != nullptr);
            adv_store(t, bd, slab);
            __syncwarp();
            uint16_t *dst = out + (size_t)b * cells;
            for (int i = lane; i < cells; i += 32) dst[i] = slab[i];
        }
        if (rng != nullptr && lane == 0) {
            rng[4 * (size_t)b] = state.hi;
            rng[4 * (size_t)b + 1] = state.lo;
        }
        __syncwarp();
    }
}

// Device part of side_effects.side_effect_score for one board per warp
// (side_effects.py:101-111): the inaction board is the initial board advanced
// num_steps[b] times, then both it and the final board run num_samples more
// steps with their occupancy counted; the three runs draw from rng[b] in that
// order.  counts: int32[n][2][H][W][8], added to.
__global__ void __launch_bounds__(ADV_WARPS * 32)
k_side_effect_warp(const uint16_t *__restrict__ init_board, const uint16_t *__restrict__ final_board,
                   int32_t *__restrict__ counts, int n, int H, int W,
                   const uint64_t *__restrict__ thr, const int32_t *__restrict__ num_steps,
                   int num_samples, uint64_t *__restrict__ rng) {
    __shared__ uint16_t slabs[ADV_WARPS][ADV_SLAB_CELLS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cells = H * W;
    uint16_t *slab = slabs[warp];
    const Narrow t(H, W, lane);
    for (int b = blockIdx.x * ADV_WARPS + warp; b < n; b += gridDim.x * ADV_WARPS) {
        u128 state = {0, 0}, inc = {0, 0};
        if (rng != nullptr) {
            state = {rng[4 * (size_t)b], rng[4 * (size_t)b + 1]};
            inc = {rng[4 * (size_t)b + 2], rng[4 * (size_t)b + 3]};
        }
        const uint64_t th = thr[b];
        const bool draws = rng != nullptr;
        const SpawnCtx ctx = make_spawn_ctx(lane, inc);
        OccCounters occ;
        for (int run = 0; run < 2; run++) {
            const uint16_t *src = (run == 0 ? init_board : final_board) + (size_t)b * cells;
            __syncwarp();
            for (int i = lane; i < cells; i += 32) slab[i] = src[i];
            __syncwarp();
            AdvBoard bd = adv_load(t, slab);
            if (run == 0) {
                const int ns = num_steps[b];
                for (int s = 0; s < ns; s++) adv_step(t, bd, state, ctx, th, draws);
            }
            int32_t *dst = counts + ((size_t)b * 2 + run) * cells * 8;
            occ.clear();
            for (int s = 0; s < num_samples; s++) {
                adv_step(t, bd, state, ctx, th, draws);
                occ.add(t, bd, dst);
            }
            if (occ.pending) occ.flush(t, dst);
        }
        if (rng != nullptr && lane == 0) {
            rng[4 * (size_t)b] = state.hi;
            rng[4 * (size_t)b + 1] = state.lo;
        }
    }
}

// ---- the same for boards up to 64x64 (Wide: two 64-bit rows per lane) ----------
// The planes stay in registers; the occupancy counts of a board are byte
// counters in shared memory (one per cell and colour, bumped for every live cell
// of a step and flushed into the int32 output every 255 steps).
struct WideBoard {
    LowPlanesT<P2> lp;
    HighPlanesT<P2> hp;
    P2 agent, kept;
};

__device__ __forceinline__ WideBoard wide_load(const Wide &t, const uint16_t *slab) {
    uint64_t a[8], b[8];
    WideBoard w;
    t.load_bytes<0x6240>(slab, a, b);
    w.lp = {{a[0], b[0]}, {a[3], b[3]}, {a[4], b[4]}, {a[5], b[5]}, {a[6], b[6]}, {a[7], b[7]}};
    w.agent = {a[1], b[1]};
    t.load_bytes<0x7351>(slab, a, b);
    w.hp = {{a[0], b[0]}, {a[1], b[1]}, {a[2], b[2]}, {a[3], b[3]}};
    w.kept = t.mask();
    return w;
}

__device__ __forceinline__ void wide_step(const Wide &t, WideBoard &b, u128 &state,
                                          const SpawnCtx &ctx, uint64_t thr, bool draws) {
    LowPlanesT<P2> &lp = b.lp;
    HighPlanesT<P2> &hp = b.hp;
    const RowDeltaT<P2> d = stencil_planes(t, lp, [&] { return hp; });
    P2 spawned = {0ull, 0ull};
    if (draws && __any_sync(FULL, nonzero(d.cand))) spawned = warp_spawn_draws_ctx(t, state, ctx, thr, d.cand);
    const P2 fresh = d.born | spawned, changed = d.die | fresh, keep = ~changed;
    lp.A = (lp.A & keep) | fresh;
    lp.D = (lp.D & keep) | (d.dn & d.born) | spawned;
    hp.R = (hp.R & keep) | (d.tR & fresh);
    hp.G = (hp.G & keep) | (d.tG & fresh);
    hp.B = (hp.B & keep) | (d.tB & fresh);
    lp.P_ = lp.P_ & keep; lp.I = lp.I & keep; lp.S = lp.S & keep;
    hp.E = hp.E & keep; b.agent = b.agent & keep;
    b.kept = b.kept & keep;
}

__device__ __forceinline__ void wide_store(const Wide &t, const WideBoard &b, uint16_t *slab) {
#pragma unroll
    for (int s = 0; s < 2; s++) {
        uint16_t *row = slab + t.row_of(s) * t.W;
        uint64_t m = t.row(t.mask(), s) & ~t.row(b.kept, s);
        const uint64_t A = t.row(b.lp.A, s), D = t.row(b.lp.D, s), R = t.row(b.hp.R, s),
                       G = t.row(b.hp.G, s), B = t.row(b.hp.B, s);
        while (m != 0) {
            const int j = lowest(m);
            m &= m - 1;
            row[j] = (uint16_t)(((A >> j) & 1u) | (((D >> j) & 1u) << 3) | (((R >> j) & 1u) << 9) |
                                (((G >> j) & 1u) << 10) | (((B >> j) & 1u) << 11));
        }
    }
}

// +1 for cells ALIVE and none of AGENT | EXIT | FROZEN, by colour (advance_board.c:153-162)
__device__ __forceinline__ void wide_count(const Wide &t, const WideBoard &b, uint8_t *cnt8) {
    const P2 occ = b.lp.A & ~(b.agent | b.hp.E | b.lp.F);
#pragma unroll
    for (int s = 0; s < 2; s++) {
        uint64_t m = t.row(occ, s);
        const uint64_t R = t.row(b.hp.R, s), G = t.row(b.hp.G, s), B = t.row(b.hp.B, s);
        uint8_t *row = cnt8 + (size_t)t.row_of(s) * t.W * 8;
        while (m != 0) {
            const int j = lowest(m);
            m &= m - 1;
            const int c = (int)((R >> j) & 1u) | ((int)((G >> j) & 1u) << 1) | ((int)((B >> j) & 1u) << 2);
            row[j * 8 + c]++;
        }
    }
}

__device__ __forceinline__ void wide_flush(uint8_t *cnt8, int32_t *dst, int cells, int lane) {
    __syncwarp();
    uint32_t *w = (uint32_t *)cnt8;
    for (int i = lane; i < cells * 2; i += 32) {        // 4 counters per word
        const uint32_t v = w[i];
        if (v != 0) {
            int32_t *d = dst + 4 * i;
            if (v & 0xffu) d[0] += v & 0xffu;
            if (v & 0xff00u) d[1] += (v >> 8) & 0xffu;
            if (v & 0xff0000u) d[2] += (v >> 16) & 0xffu;
            if (v >> 24) d[3] += v >> 24;
            w[i] = 0;
        }
    }
    __syncwarp();
}

constexpr int WIDE_ADV_WARPS = 4;
// the plane loads of lanes past the last row read up to cell 63 * W + 32 + 56 of the slab
__host__ __device__ inline int wide_slab_cells(int H, int W) {
    return H * W > 64 * W + 96 ? H * W : 64 * W + 96;
}
__host__ __device__ inline size_t wide_slab_bytes(int H, int W) {
    return ((size_t)wide_slab_cells(H, W) * 2 + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t wide_warp_bytes(int H, int W, bool counting) {
    return wide_slab_bytes(H, W) + (counting ? (size_t)H * W * 8 : 0);
}

// mode 0: out[b] = board after n_steps (advance_board_nstep); mode 1: counts[b] +=
// life_occupancy over n_steps; mode 2: the three runs of side_effect_score
// (counts[b][0..1], first run preceded by num_steps[b] uncounted steps).
__global__ void __launch_bounds__(WIDE_ADV_WARPS * 32)
k_multistep_wide(int mode, const uint16_t *in, const uint16_t *in2, uint16_t *out,
                 int32_t *__restrict__ counts, int n, int H, int W, uint64_t thr_all,
                 const uint64_t *__restrict__ thr_each, int n_steps,
                 const int32_t *__restrict__ lead_steps, uint64_t *__restrict__ rng) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cells = H * W;
    const bool counting = mode != 0;
    unsigned char *mine = smem_raw + warp * wide_warp_bytes(H, W, counting);
    uint16_t *slab = (uint16_t *)mine;
    uint8_t *cnt8 = mine + wide_slab_bytes(H, W);
    const Wide t(H, W, lane);
    if (counting)
        for (int i = lane; i < cells * 2; i += 32) ((uint32_t *)cnt8)[i] = 0;
    for (int b = blockIdx.x * WIDE_ADV_WARPS + warp; b < n; b += gridDim.x * WIDE_ADV_WARPS) {
        u128 state = {0, 0}, inc = {0, 0};
        if (rng != nullptr) {
            state = {rng[4 * (size_t)b], rng[4 * (size_t)b + 1]};
            inc = {rng[4 * (size_t)b + 2], rng[4 * (size_t)b + 3]};
        }
        const uint64_t th = thr_each != nullptr ? thr_each[b] : thr_all;
        const bool draws = rng != nullptr;
        const SpawnCtx ctx = make_spawn_ctx(lane, inc);
        const int runs = mode == 2 ? 2 : 1;
        for (int run = 0; run < runs; run++) {
            const uint16_t *src = (run == 0 ? in : in2) + (size_t)b * cells;
            __syncwarp();
            for (int i = lane; i < cells; i += 32) slab[i] = src[i];
            __syncwarp();
            WideBoard bd = wide_load(t, slab);
            if (mode == 2 && run == 0) {
                const int lead = lead_steps[b];
                for (int s = 0; s < lead; s++) wide_step(t, bd, state, ctx, th, draws);
            }
            if (!counting) {
                for (int s = 0; s < n_steps; s++) wide_step(t, bd, state, ctx, th, draws);
                wide_store(t, bd, slab);
                __syncwarp();
                uint16_t *dst = out + (size_t)b * cells;
                for (int i = lane; i < cells; i += 32) dst[i] = slab[i];
            } else {
                int32_t *dst = counts + ((size_t)b * runs + run) * cells * 8;
                int pending = 0;
                for (int s = 0; s < n_steps; s++) {
                    wide_step(t, bd, state, ctx, th, draws);
                    wide_count(t, bd, cnt8);
                    if (++pending == 255) {
                        wide_flush(cnt8, dst, cells, lane);
                        pending = 0;
                    }
                }
                if (pending) wide_flush(cnt8, dst, cells, lane);
            }
        }
        if (rng != nullptr && lane == 0) {
            rng[4 * (size_t)b] = state.hi;
            rng[4 * (size_t)b + 1] = state.lo;
        }
        __syncwarp();
    }
}

inline size_t multistep_wide_smem(int H, int W, bool counting) {
    return WIDE_ADV_WARPS * wide_warp_bytes(H, W, counting);
}
inline int multistep_wide_grid(int n, int sm_count) {
    const int grid = (n + WIDE_ADV_WARPS - 1) / WIDE_ADV_WARPS, cap = sm_count * 2;
    return grid > cap ? cap : grid;
}

inline int advance_warp_grid(int n, int sm_count) {
    const int grid = (n + ADV_WARPS - 1) / ADV_WARPS, cap = sm_count * 8;
    return grid > cap ? cap : grid;
}

inline void launch_advance_warp(const uint16_t *in, uint16_t *out, int32_t *counts, int n, int H,
                                int W, uint64_t thr, int n_steps, uint64_t *rng, int sm_count,
                                cudaStream_t st) {
    k_advance_warp<<<advance_warp_grid(n, sm_count), ADV_WARPS * 32, 0, st>>>(
        in, out, counts, n, H, W, thr, n_steps, rng);
}

inline void launch_side_effect_warp(const uint16_t *init_board, const uint16_t *final_board,
                                    int32_t *counts, int n, int H, int W, const uint64_t *thr,
                                    const int32_t *num_steps, int num_samples, uint64_t *rng,
                                    int sm_count, cudaStream_t st) {
    k_side_effect_warp<<<advance_warp_grid(n, sm_count), ADV_WARPS * 32, 0, st>>>(
        init_board, final_board, counts, n, H, W, thr, num_steps, num_samples, rng);
}

}  // namespace sl
