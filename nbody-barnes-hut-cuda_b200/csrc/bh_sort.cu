// bh_sort.cu — hand-written onesweep LSD radix sort of (u32 key, u32 value) pairs.
//
// Replaces thrust::sort_by_key(d_keys, d_keys + N, d_vals)  nbody_v5_bench.cu:262-264
// (stable ascending; ties keep ascending input position).
//
// Structure (RADIX_BITS-bit digits.  8 bits = four passes over the reference's 30-bit key.  Three passes of 10
// bits — 52 instead of 68 B/pair, SURVEY §8d — were built and measured in round 2: 0.125 vs 0.098 ms at 1M and
// 0.91 vs 0.71 ms at 16M random keys; with 1,024 bins a 4,096-key tile leaves 4-key (16-byte) runs to write, the
// look-back touches four times as many status words and the kernel spills at 80 registers.  The code stays
// generic in RADIX_BITS — `make EXTRA_NVFLAGS=-DBH_RADIX_BITS=10` rebuilds that variant):
//   1. one histogram kernel counts every pass's digits in a single read of the keys
//      (shared-memory histograms, one global atomic per bin per CTA);
//   2. a RADIX-thread kernel turns each pass's histogram into exclusive bucket bases;
//   3. one "onesweep" kernel per digit: a CTA takes a tile ticket, ranks its 6,912 keys (384 threads x 18) in
//      warp-level digit groups (stable; the groups come from one ballot per digit bit, not from match.any — see
//      digit_peers), publishes its per-digit counts,
//      resolves the cross-tile prefix by decoupled look-back on packed status words (RADIX / 256 digits
//      per thread, their look-back chains interleaved), stages the tile digit-ordered in shared memory and
//      writes runs out coalesced.
//   4. the LAST pass of the step's sort also moves the bodies: it knows every pair's final slot, so it
//      gathers posm / vel / ids by the carried index and writes them in sorted order itself — no separate
//      reorder launch, the permutation is not re-read from HBM.
// Algorithmic traffic: 4 B (histogram) + 16 B per pass per pair.
#include "bh_common.cuh"

namespace {

#ifndef BH_RADIX_BITS
#define BH_RADIX_BITS 8
#endif
constexpr int RADIX_BITS = BH_RADIX_BITS;
constexpr int RADIX = 1 << RADIX_BITS;
#ifndef BH_SORT_THREADS
#define BH_SORT_THREADS 384
#endif
constexpr int SORT_THREADS = BH_SORT_THREADS;
constexpr int SORT_WARPS = SORT_THREADS / 32;
#ifndef BH_SORT_ITEMS
#define BH_SORT_ITEMS 18
#endif
constexpr int ITEMS = BH_SORT_ITEMS;
constexpr int TILE = SORT_THREADS * ITEMS;  // 6,912 pairs per CTA (round 1: 256 x 16 = 4,096; DESIGN.md 5.3 has the sweeps)
constexpr int MAX_PASSES = 4;                // 32 key bits = 4 x 8 (or 10 + 10 + 10 + 2)
// digits per thread in the per-digit part of a pass; with more threads than digits the first RADIX threads own one each
constexpr int DPT = RADIX >= SORT_THREADS ? RADIX / SORT_THREADS : 1;
static_assert((RADIX >= SORT_THREADS ? DPT * SORT_THREADS == RADIX : true) && 32 * ITEMS < 65536 && SORT_THREADS % 32 == 0,
              "digit ownership / 16-bit warp counters");

constexpr uint32_t ST_MASK = 0x3FFFFFFFu;
constexpr uint32_t ST_LOCAL = 0x40000000u;  // tile's own count is published
constexpr uint32_t ST_INCL = 0x80000000u;   // inclusive prefix over all earlier tiles is published
constexpr unsigned SPIN_LIMIT = 1u << 24;

struct PassDesc {
    int shift[MAX_PASSES];
    uint32_t mask[MAX_PASSES];
    int passes;
};

__global__ void __launch_bounds__(SORT_THREADS) histogram_kernel(const uint32_t* __restrict__ keys, int64_t n,
                                                                PassDesc pd, uint32_t* __restrict__ hist) {
    __shared__ uint32_t s_hist[MAX_PASSES][RADIX];
    for (int i = threadIdx.x; i < MAX_PASSES * RADIX; i += SORT_THREADS) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    // warp-uniform trip count; each lane takes 4 consecutive keys (one 128-bit load).  Morton-coherent
    // input puts a whole warp in one bin for the upper digits: that case is one vote + one atomic,
    // everything else falls back to per-lane shared atomics (spread addresses, no serialisation).
    const int64_t n4 = n >> 2;
    const uint4* keys4 = reinterpret_cast<const uint4*>(keys);
    for (int64_t base = (int64_t)blockIdx.x * SORT_THREADS + (threadIdx.x & ~31); base < n4;
         base += (int64_t)gridDim.x * SORT_THREADS) {
        const int64_t i = base + bh_lane();
        const bool valid = i < n4;
        const uint4 k4 = valid ? __ldg(keys4 + i) : make_uint4(0, 0, 0, 0);
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
        const int lead = __ffs(vmask) - 1;
#pragma unroll
        for (int p = 0; p < MAX_PASSES; ++p) {
            if (p < pd.passes) {
                const uint32_t d0 = (k4.x >> pd.shift[p]) & pd.mask[p], d1 = (k4.y >> pd.shift[p]) & pd.mask[p];
                const uint32_t d2 = (k4.z >> pd.shift[p]) & pd.mask[p], d3 = (k4.w >> pd.shift[p]) & pd.mask[p];
                const uint32_t ref = __shfl_sync(0xffffffffu, d0, lead);
                const bool mine_same = (d0 == ref) & (d1 == ref) & (d2 == ref) & (d3 == ref);
                const unsigned same = __ballot_sync(0xffffffffu, valid && mine_same);
                if (same == vmask) {
                    if ((int)bh_lane() == lead) atomicAdd(&s_hist[p][ref], 4u * (uint32_t)__popc(vmask));
                } else if (valid) {
                    atomicAdd(&s_hist[p][d0], 1u); atomicAdd(&s_hist[p][d1], 1u);
                    atomicAdd(&s_hist[p][d2], 1u); atomicAdd(&s_hist[p][d3], 1u);
                }
            }
        }
    }
    // the (< 4) keys past the last full quad
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const uint32_t k = __ldg(keys + (n4 << 2) + threadIdx.x);
        for (int p = 0; p < pd.passes; ++p) atomicAdd(&s_hist[p][(k >> pd.shift[p]) & pd.mask[p]], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < pd.passes * RADIX; i += SORT_THREADS) {
        uint32_t c = (&s_hist[0][0])[i];
        if (c) atomicAdd(hist + i, c);
    }
}

// exclusive scan of each pass's 256 counters, in place (one CTA, one warp-scan per pass row)
__global__ void __launch_bounds__(RADIX) scan_hist_kernel(uint32_t* hist, int passes) {
    __shared__ uint32_t s_warp[RADIX / 32];
    for (int p = 0; p < passes; ++p) {
        uint32_t v = hist[p * RADIX + threadIdx.x];
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if ((int)bh_lane() >= o) inc += t;
        }
        if (bh_lane() == 31) s_warp[threadIdx.x >> 5] = inc;
        __syncthreads();
        uint32_t base = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) base += s_warp[w];
        hist[p * RADIX + threadIdx.x] = base + inc - v;
        __syncthreads();
    }
}

// Lanes of the warp that hold the same digit.  One ballot per digit bit instead of match.any: the ballots of a
// key and of the ITEMS keys of a thread are independent, while match.any takes longer the more distinct digits a
// warp holds (random keys: 32).
__device__ __forceinline__ unsigned digit_peers(uint32_t d) {
#ifdef BH_SORT_MATCH_ANY
    return __match_any_sync(0xffffffffu, d);
#else
    unsigned peers = 0xffffffffu;
#ifdef BH_SORT_UNIFORM_FAST
    if (__all_sync(0xffffffffu, d == __shfl_sync(0xffffffffu, d, 0))) return peers;
#endif
#pragma unroll
    for (int b = 0; b < RADIX_BITS; ++b) {
        const bool bit = (d >> b) & 1u;
        const unsigned m = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? m : ~m;
    }
    return peers;
#endif
}

// what the last pass of the step's sort moves besides the pairs
struct BodyMove {
    const float4* posm_in; const float4* vel_in; const int32_t* ids_in;
    float4* posm_out; float4* vel_out; int32_t* ids_out;
};

#ifndef BH_SORT_MIN_CTAS
#define BH_SORT_MIN_CTAS 2
#endif
template <bool IOTA, bool MOVE>
__global__ void __launch_bounds__(SORT_THREADS, BH_SORT_MIN_CTAS) onesweep_kernel(const uint32_t* __restrict__ keys_in,
                                                               const uint32_t* __restrict__ vals_in,
                                                               uint32_t* __restrict__ keys_out,
                                                               uint32_t* __restrict__ vals_out, int64_t n, int shift,
                                                               uint32_t mask, const uint32_t* __restrict__ bucket_base,
                                                               volatile uint32_t* lookback, unsigned int* ticket,
                                                               unsigned int* err_flag, BodyMove mv) {
    __shared__ uint32_t s_buf[TILE];
    __shared__ uint16_t s_whist[SORT_WARPS][RADIX];   // a warp holds 32 * ITEMS = 512 keys: counts fit 16 bits
    __shared__ uint32_t s_tile_excl[RADIX];
    __shared__ uint32_t s_global_off[RADIX];   // n < 2^30: 32-bit wrap-around arithmetic is exact
    __shared__ uint32_t s_scan[SORT_WARPS];
    __shared__ unsigned int s_tile;

    const int lane = bh_lane(), warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = threadIdx.x; i < SORT_WARPS * RADIX / 2; i += SORT_THREADS) reinterpret_cast<uint32_t*>(&s_whist[0][0])[i] = 0;
    __syncthreads();
    const unsigned tile = s_tile;
    const int64_t tile_base = (int64_t)tile * TILE;
    const int64_t warp_base = tile_base + (int64_t)warp * (32 * ITEMS);
    const int tile_valid = (int)((n - tile_base) < TILE ? (n - tile_base) : TILE);

    uint32_t key[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        int64_t i = warp_base + k * 32 + lane;
        key[k] = i < n ? __ldg(keys_in + i) : 0xFFFFFFFFu;
    }

    // values: requested now so that their latency hides behind the ranking (a 1M-pair sort has fewer
    // than two tiles per SM: nothing else would cover it)
    uint32_t val[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        int64_t i = warp_base + k * 32 + lane;
        if (IOTA) val[k] = (uint32_t)i;
        else val[k] = i < n ? __ldg(vals_in + i) : 0u;
    }

    // ---- stable in-warp ranking by digit groups ---------------------------------------------
    uint32_t rank[ITEMS];
    uint16_t* my_hist = s_whist[warp];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const uint32_t d = (key[k] >> shift) & mask;
        const unsigned peers = digit_peers(d);
        const int leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if (lane == leader) { base = my_hist[d]; my_hist[d] = (uint16_t)(base + (uint32_t)__popc(peers)); }
        base = __shfl_sync(0xffffffffu, base, leader);
        rank[k] = base + (uint32_t)__popc(peers & ((1u << lane) - 1u));
        __syncwarp();
    }
    __syncthreads();

    // ---- per-digit: warp offsets, tile count, publish, look-back --------------------------
    {
        const bool owns = (int)threadIdx.x * DPT < RADIX;            // threads beyond the digits only take part in the scans
        const int d0 = owns ? threadIdx.x * DPT : 0;                  // this thread owns digits d0 .. d0 + DPT - 1
        uint32_t running[DPT];
#pragma unroll
        for (int q = 0; q < DPT; ++q) {
            uint32_t r = 0;
            if (owns) {
#ifdef BH_SORT_FOLD
#pragma unroll
                for (int w = 0; w < SORT_WARPS; ++w) r += s_whist[w][d0 + q];
#else
#pragma unroll
                for (int w = 0; w < SORT_WARPS; ++w) {
                    const uint32_t c = s_whist[w][d0 + q];
                    s_whist[w][d0 + q] = (uint16_t)r;
                    r += c;
                }
#endif
                lookback[(size_t)tile * RADIX + d0 + q] = (tile == 0 ? ST_INCL : ST_LOCAL) | r;
            }
            running[q] = r;
        }

        // exclusive scan of the tile's digit counts across the threads (DPT consecutive digits each)
        uint32_t mine = 0;
#pragma unroll
        for (int q = 0; q < DPT; ++q) mine += running[q];
        uint32_t inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_scan[warp] = inc;
        __syncthreads();
        uint32_t excl = inc - mine;
        for (int w = 0; w < warp; ++w) excl += s_scan[w];
        uint32_t excl_in_tile[DPT];
#pragma unroll
        for (int q = 0; q < DPT; ++q) {
            excl_in_tile[q] = excl;
#ifdef BH_SORT_FOLD
            if (owns) {   // a key's slot in the staged tile = digit start in the tile + keys of earlier warps + rank in its warp
                uint32_t r = excl;   // < TILE <= 65,535: fits the 16-bit counters
#pragma unroll
                for (int w = 0; w < SORT_WARPS; ++w) {
                    const uint32_t c = s_whist[w][d0 + q];
                    s_whist[w][d0 + q] = (uint16_t)r;
                    r += c;
                }
            }
#else
            if (owns) s_tile_excl[d0 + q] = excl;
#endif
            excl += running[q];
        }

        uint32_t prior[DPT];
#pragma unroll
        for (int q = 0; q < DPT; ++q) prior[q] = 0;
        if (tile > 0 && owns) {
            // Decoupled look-back, LB_WIDE predecessors per round trip and the DPT digit chains of this thread
            // interleaved: the statuses of tiles p, p-1, ... are fetched together (independent loads) and
            // consumed in order.  The chain of dependent L2 round trips, which bounds the first wave when
            // hundreds of tiles start together, shrinks by the same factor.
#ifndef BH_SORT_LB_WIDE
#define BH_SORT_LB_WIDE 8
#endif
            constexpr int LB_WIDE = DPT == 1 ? BH_SORT_LB_WIDE : 4;
            int p[DPT];
            bool done[DPT];
#pragma unroll
            for (int q = 0; q < DPT; ++q) { p[q] = (int)tile - 1; done[q] = false; }
            unsigned spins = 0;
            bool all_done = false;
            while (!all_done) {
                uint32_t st[DPT][LB_WIDE];
#pragma unroll
                for (int q = 0; q < DPT; ++q)
#pragma unroll
                    for (int j = 0; j < LB_WIDE; ++j)
                        st[q][j] = (!done[q] && p[q] - j >= 0) ? lookback[(size_t)(p[q] - j) * RADIX + d0 + q] : 0x80000000u;
                all_done = true;
                bool progress = false;
#pragma unroll
                for (int q = 0; q < DPT; ++q) {
                    if (done[q]) continue;
#pragma unroll
                    for (int j = 0; j < LB_WIDE; ++j) {
                        if (done[q]) break;
                        const uint32_t sv = st[q][j];
                        if (sv & ST_INCL) { prior[q] += sv & ST_MASK; done[q] = true; progress = true; }
                        else if (sv & ST_LOCAL) { prior[q] += sv & ST_MASK; --p[q]; progress = true; }
                        else break;   // not published yet: poll again from this tile
                    }
                    all_done = all_done && done[q];
                }
                spins = progress ? 0 : spins + 1;
                if (!all_done && spins > SPIN_LIMIT) { atomicOr(err_flag, BH_DERR_SORT_SPIN); break; }
            }
#pragma unroll
            for (int q = 0; q < DPT; ++q) lookback[(size_t)tile * RADIX + d0 + q] = ST_INCL | (prior[q] + running[q]);
        }
        if (owns) {
#pragma unroll
            for (int q = 0; q < DPT; ++q) s_global_off[d0 + q] = bucket_base[d0 + q] + prior[q] - excl_in_tile[q];
        }
    }
    __syncthreads();

    // ---- stage digit-ordered in shared memory, write coalesced runs ------------------------
    uint32_t pos[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const uint32_t d = (key[k] >> shift) & mask;
#ifdef BH_SORT_FOLD
        pos[k] = s_whist[warp][d] + rank[k];
#else
        pos[k] = s_tile_excl[d] + s_whist[warp][d] + rank[k];
#endif
        s_buf[pos[k]] = key[k];
    }
    __syncthreads();
    uint32_t dst[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const int p = threadIdx.x + k * SORT_THREADS;
        const uint32_t kk = s_buf[p];
        const uint32_t d = (kk >> shift) & mask;
        dst[k] = s_global_off[d] + (uint32_t)p;
        if (p < tile_valid) keys_out[dst[k]] = kk;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) s_buf[pos[k]] = val[k];
    __syncthreads();
    if (!MOVE) {
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const int p = threadIdx.x + k * SORT_THREADS;
            if (p < tile_valid) vals_out[dst[k]] = s_buf[p];
        }
    } else {
        // the carried value is the body's slot in the unsorted state: gather it, write it at its final slot
        // (consecutive threads write consecutive slots within a digit run)
#pragma unroll
        for (int k0 = 0; k0 < ITEMS; k0 += 4) {
            uint32_t j[4];
            float4 pp[4], vv[4];
            int32_t id[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (k0 + k >= ITEMS) break;   // ITEMS need not be a multiple of 4 (resolved at compile time)
                const int p = threadIdx.x + (k0 + k) * SORT_THREADS;
                j[k] = s_buf[p];
                if (p < tile_valid) { pp[k] = __ldg(mv.posm_in + j[k]); vv[k] = __ldg(mv.vel_in + j[k]); id[k] = __ldg(mv.ids_in + j[k]); }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (k0 + k >= ITEMS) break;
                const int p = threadIdx.x + (k0 + k) * SORT_THREADS;
                if (p < tile_valid) {
                    const uint32_t o = dst[k0 + k];
                    vals_out[o] = j[k];
                    mv.posm_out[o] = pp[k]; mv.vel_out[o] = vv[k]; mv.ids_out[o] = id[k];
                }
            }
        }
    }
}

PassDesc make_passes(int begin_bit, int end_bit) {
    PassDesc pd{};
    int b = begin_bit;
    while (b < end_bit && pd.passes < MAX_PASSES) {
        int w = end_bit - b < RADIX_BITS ? end_bit - b : RADIX_BITS;
        pd.shift[pd.passes] = b;
        pd.mask[pd.passes] = (1u << w) - 1u;
        ++pd.passes;
        b += w;
    }
    return pd;
}

}  // namespace

int bh_sort_passes(int bits) { return (bits + RADIX_BITS - 1) / RADIX_BITS; }

BhSortPlan bh_sort_plan(int64_t n) {
    BhSortPlan p{};
    p.n = n;
    p.num_tiles = (int)((n + TILE - 1) / TILE);
    if (p.num_tiles < 1) p.num_tiles = 1;
    p.hist_bytes = sizeof(uint32_t) * MAX_PASSES * RADIX;
    p.lookback_bytes = sizeof(uint32_t) * (size_t)MAX_PASSES * p.num_tiles * RADIX;
    p.ticket_bytes = 256;  // MAX_PASSES tickets, padded
    p.total_bytes = p.hist_bytes + p.lookback_bytes + p.ticket_bytes;
    return p;
}

// Pass 0 reads (keys_src, vals_src) and writes (keys_p, vals_p); later passes ping-pong p -> q -> p ...
// The sorted pairs end in p when the pass count is odd and in q when it is even (*result_in_q).
// keys_q may alias keys_src (the source is dead after pass 0).
// With posm_in != nullptr the last pass also reorders the bodies: out[i] = in[sorted value i].
int bh_sort_pairs_move_launch(const uint32_t* keys_src, const uint32_t* vals_src, uint32_t* keys_p, uint32_t* vals_p,
                              uint32_t* keys_q, uint32_t* vals_q, int64_t n, int begin_bit, int end_bit, void* tmp,
                              bool vals_in_is_iota, unsigned int* err_flag, int* result_in_q, const float4* posm_in,
                              const float4* vel_in, const int32_t* ids_in, float4* posm_out, float4* vel_out,
                              int32_t* ids_out, cudaStream_t st) {
    if (n < 0 || begin_bit < 0 || end_bit > 32 || begin_bit >= end_bit) return BH_E_INVAL;
    if (end_bit - begin_bit > MAX_PASSES * RADIX_BITS) return BH_E_INVAL;
    if (n >= (int64_t)ST_MASK) return BH_E_UNSUPPORTED;
    PassDesc pd = make_passes(begin_bit, end_bit);
    if (result_in_q) *result_in_q = (pd.passes & 1) ? 0 : 1;
    if (n == 0) return 0;
    BhSortPlan plan = bh_sort_plan(n);
    uint32_t* hist = (uint32_t*)tmp;
    uint32_t* lookback = (uint32_t*)((char*)tmp + plan.hist_bytes);
    unsigned int* tickets = (unsigned int*)((char*)tmp + plan.hist_bytes + plan.lookback_bytes);
    BH_CUDA_TRY(cudaMemsetAsync(tmp, 0, plan.total_bytes, st));
    int hblocks = (int)((n / 4 + (int64_t)SORT_THREADS * 4 - 1) / ((int64_t)SORT_THREADS * 4));
    if (hblocks < 1) hblocks = 1;
    if (hblocks > BH_NUM_SMS_FALLBACK * 8) hblocks = BH_NUM_SMS_FALLBACK * 8;
    histogram_kernel<<<hblocks, SORT_THREADS, 0, st>>>(keys_src, n, pd, hist);
    scan_hist_kernel<<<1, RADIX, 0, st>>>(hist, pd.passes);
    const uint32_t *kin = keys_src, *vin = vals_src;
    const BodyMove none{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    const BodyMove move{posm_in, vel_in, ids_in, posm_out, vel_out, ids_out};
    for (int p = 0; p < pd.passes; ++p) {
        uint32_t* kout = (p & 1) ? keys_q : keys_p;
        uint32_t* vout = (p & 1) ? vals_q : vals_p;
        volatile uint32_t* lb = lookback + (size_t)p * plan.num_tiles * RADIX;
        const bool iota = p == 0 && vals_in_is_iota, last = posm_in != nullptr && p == pd.passes - 1;
#define BH_SWEEP(I, M, mvarg) onesweep_kernel<I, M><<<plan.num_tiles, SORT_THREADS, 0, st>>>(kin, vin, kout, vout, n, pd.shift[p], pd.mask[p], hist + p * RADIX, lb, tickets + p, err_flag, mvarg)
        if (iota && last) BH_SWEEP(true, true, move);
        else if (iota) BH_SWEEP(true, false, none);
        else if (last) BH_SWEEP(false, true, move);
        else BH_SWEEP(false, false, none);
#undef BH_SWEEP
        kin = kout;
        vin = vout;
    }
    return (int)cudaGetLastError();
}

int bh_sort_pairs_launch(const uint32_t* keys_src, const uint32_t* vals_src, uint32_t* keys_p, uint32_t* vals_p,
                         uint32_t* keys_q, uint32_t* vals_q, int64_t n, int begin_bit, int end_bit, void* tmp,
                         bool vals_in_is_iota, unsigned int* err_flag, int* result_in_q, cudaStream_t st) {
    return bh_sort_pairs_move_launch(keys_src, vals_src, keys_p, vals_p, keys_q, vals_q, n, begin_bit, end_bit, tmp,
                                     vals_in_is_iota, err_flag, result_in_q, nullptr, nullptr, nullptr, nullptr, nullptr,
                                     nullptr, st);
}
