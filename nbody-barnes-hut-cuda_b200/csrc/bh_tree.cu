// bh_tree.cu — parallel emission of the compressed octree from the sorted 30-bit Morton keys,
// and the bottom-up centre-of-mass pass.
//
// Replaces, for the whole tree at once and without a zeroed 2N node pool:
//   cudaMemset + initRootKernel      nbody_v5_bench.cu:266-267, 65-81
//   insertParticlesKernel x N/1024   nbody_v5_bench.cu:83-132, 269-275
//   cudaMemcpy(&hCount ...)          nbody_v5_bench.cu:277-278   (no host round trip here)
//   computeCOMKernel/finalizeCOM     nbody_v5_bench.cu:158-189   (no atomics at all here: prefix sums)
//
// Canonical tree (DESIGN.md §"Tree"): a CELL is a maximal run of >= 2 sorted bodies whose keys
// share exactly L leading 3-bit digits (L = level, 0..10).  L == 10 means identical keys: a
// BUCKET leaf holding its bodies as a contiguous range.  A cell with L < 10 has up to 8
// children, one per value of digit L+1: a single body, or another cell.  Bodies sit directly in
// their parent's child table (one body per leaf, as the reference's insertion tree).
//
// Karras-style construction, one thread per adjacent key pair j = (j, j+1):
//   lv(j)      = digits shared by K[j], K[j+1]   — pair j is a split point of the level-lv(j)
//                cell that contains it;
//   leader(j)  = the split point that ends the FIRST child of that cell (found from bit masks of the pair
//                levels in shared memory; galloping searches on the keys only where a cell leaves
//                the tile) — exactly one leader per cell;
//   cell id    = exclusive prefix sum of the leader flags (so ids ascend with leader position);
//   parent     = the cell owning the tighter of the two pairs that bound the range, lv(l-1) vs
//                lv(r) — its id is scan[leader(that pair)].
// Every cell and every loose body then writes itself into its parent's child table.
//
// Traversal view of the tree (written by the centre-of-mass pass, read by bh_force.cu / bh_let.cu): per cell one
// DENSE line of its children in digit order — kid_src[8c + r] = what child r contributes as a source (a loose
// body's {x,y,z,m}, a child cell's {centre of mass, mass}), kid_info[8c + r] = {stack word of the child cell
// (id << 3 | its own child count - 1 | bucket flag), its squared width as float bits (-1 for a body)}, r < number of children.  The octree of
// the reference disk holds 3.2 children per cell: dense lines let the traversal spend its lanes on children that
// exist.  cell_child stays the canonical digit-indexed table (parity tests, this pass).
#include "bh_common.cuh"

namespace {

constexpr int TB = 256;           // threads per CTA
constexpr int SCAN_ITEMS = 8;     // pairs per thread in the flag scan
constexpr int SCAN_TILE = TB * SCAN_ITEMS;

template <int LEVELS>
__device__ __forceinline__ int lv_pair(const typename BhKey<LEVELS>::type* __restrict__ K, int j, int n) {
    if (j < 0 || j >= n - 1) return -1;
    return bh_shared_digits_t<LEVELS>(__ldg(K + j), __ldg(K + j + 1));
}

// smallest index t <= j whose key shares L digits with K[j] (keys ascending)
template <int LEVELS>
__device__ int find_left(const typename BhKey<LEVELS>::type* __restrict__ K, int j, int L) {
    typedef typename BhKey<LEVELS>::type KeyT;
    if (L <= 0) return 0;
    const int s = 3 * LEVELS - 3 * L;
    const KeyT p = __ldg(K + j) >> s;
    int good = j, bad = -1, step = 1;
    for (;;) {
        int t = good - step;
        if (t < 0) { bad = -1; break; }
        if ((__ldg(K + t) >> s) == p) { good = t; step <<= 1; }
        else { bad = t; break; }
    }
    while (good - bad > 1) {
        int mid = (good + bad) >> 1;
        if ((__ldg(K + mid) >> s) == p) good = mid; else bad = mid;
    }
    return good;
}

// largest index t >= j whose key shares L digits with K[j]
template <int LEVELS>
__device__ int find_right(const typename BhKey<LEVELS>::type* __restrict__ K, int n, int j, int L) {
    typedef typename BhKey<LEVELS>::type KeyT;
    if (L <= 0) return n - 1;
    const int s = 3 * LEVELS - 3 * L;
    const KeyT p = __ldg(K + j) >> s;
    int good = j, bad = n, step = 1;
    for (;;) {
        int t = good + step;
        if (t >= n) { bad = n; break; }
        if ((__ldg(K + t) >> s) == p) { good = t; step <<= 1; }
        else { bad = t; break; }
    }
    while (bad - good > 1) {
        int mid = (good + bad) >> 1;
        if ((__ldg(K + mid) >> s) == p) good = mid; else bad = mid;
    }
    return good;
}

// ---- pass A: leader of every pair + per-tile leader counts ---------------------------------
// In terms of the pair levels lv(t) alone (keys ascending): the level-L cell around pair j, lv(j) = L, starts right
// after the nearest pair to the LEFT of j with a level below L and ends at the nearest such pair to the RIGHT; its
// leader is the first pair from its start on whose level is <= L.  A tile of 2,048 pairs answers these "nearest pair
// with level < l" questions from bit masks in shared memory — per row of 32 pairs and threshold l one ballot, per
// threshold one 64-bit word saying which rows have any — with a handful of bit operations instead of galloping
// searches over the keys.  Only a search that leaves the tile (its cell began before / ends after it: at most seven
// pairs per level and side) falls back to find_left / find_right on the keys.  Leaders also record where their cell
// ends and which of the two bounding pairs is the tighter one (-> the parent): link_kernel does no searching at all.
constexpr int TILE_ROWS = SCAN_TILE / 32;   // 64 rows of 32 pairs; pair j of a tile sits in row (j - tile0) / 32, bit (j - tile0) % 32
static_assert(TILE_ROWS == 64, "one 64-bit word of row flags per threshold");

// nearest set position strictly below (row, bit) in the masks of one threshold; -1 if none in the tile
__device__ __forceinline__ int mask_prev(const uint32_t* rows, unsigned long long any, int row, int bit) {
    const uint32_t m = rows[row] & ((1u << bit) - 1u);
    if (m) return row * 32 + 31 - __clz((int)m);
    const unsigned long long a = any & ((1ull << row) - 1ull);
    if (!a) return -1;
    const int r = 63 - __clzll((long long)a);
    return r * 32 + 31 - __clz((int)rows[r]);
}
// nearest set position at or above (row, bit); -1 if none in the tile
__device__ __forceinline__ int mask_next(const uint32_t* rows, unsigned long long any, int row, int bit) {
    const uint32_t m = rows[row] & ~((1u << bit) - 1u);
    if (m) return row * 32 + __ffs((int)m) - 1;
    const unsigned long long a = row >= 63 ? 0ull : any & ~((2ull << row) - 1ull);
    if (!a) return -1;
    const int r = __ffsll((long long)a) - 1;
    return r * 32 + __ffs((int)rows[r]) - 1;
}

// One pair: answers from the tile's masks; a search that leaves the tile continues on the keys.  *leads: the pair
// leads a cell.
template <int LEVELS>
__device__ __forceinline__ void pair_one(const typename BhKey<LEVELS>::type* __restrict__ K, int n, int tile0, int rel,
                                         const uint32_t (*s_mask)[TILE_ROWS], const unsigned long long* s_any,
                                         const int8_t* s_lv, int2* __restrict__ pair_info, int2* __restrict__ pair_aux,
                                         bool* leads) {
    const int npairs = n - 1;
    const int j = tile0 + rel, row = rel >> 5, bit = rel & 31;
    const int L = s_lv[rel];
    int lead, first;
    *leads = false;
    if (L == LEVELS) {   // inside a run of identical keys: the run's first pair leads
        first = j;
        const bool starts = j == 0 || (rel > 0 ? s_lv[rel - 1] < LEVELS : __ldg(K + j - 1) != __ldg(K + j));
        lead = starts ? j : -1;
    } else {
        bool first_in_tile = true;
        if (L == 0) {
            first = 0;
            first_in_tile = tile0 == 0;
        } else {
            const int t = mask_prev(s_mask[L], s_any[L], row, bit);
            if (t >= 0) first = tile0 + t + 1;
            else if (tile0 == 0) first = 0;
            else { first = find_left<LEVELS>(K, tile0, L); first_in_tile = first == tile0; }   // bodies tile0 .. j share L digits
        }
        if (first_in_tile) {   // first pair from the cell's start on with level <= L (pair j itself at the latest)
            const int f = first - tile0;
            lead = tile0 + mask_next(s_mask[L + 1], s_any[L + 1], f >> 5, f & 31);
        } else {
            // j leads iff it closes the first child, i.e. K[j] still shares L+1 digits with K[first]
            lead = (bh_shared_digits_t<LEVELS>(__ldg(K + first), __ldg(K + j)) >= L + 1) ? j : find_right<LEVELS>(K, n, first, L + 1);
        }
    }
    if (lead == j) {
        // the cell's last body, and the tighter of the two pairs bounding [first, last] (its level is the parent's)
        int last;
        if (L == 0) last = n - 1;
        else {
            const int t = rel == SCAN_TILE - 1 ? -1 : mask_next(s_mask[L], s_any[L], (rel + 1) >> 5, (rel + 1) & 31);
            if (t >= 0) last = tile0 + t;
            else last = find_right<LEVELS>(K, n, min(tile0 + SCAN_TILE, n - 1), L);   // bodies j .. tile end share L digits
        }
        const int la = first - 1, lb = last;   // bounding pairs (la = -1 / lb = n - 1: none on that side)
        int a = -1, b = -1;
        if (la >= 0) {
            if (la >= tile0) a = s_lv[la - tile0];
            else a = lv_pair<LEVELS>(K, la, n);
        }
        if (lb < npairs) {
            if (lb < tile0 + SCAN_TILE) b = s_lv[lb - tile0];
            else b = lv_pair<LEVELS>(K, lb, n);
        }
        const int sp = (a < 0 && b < 0) ? -1 : ((a >= b) ? la : lb);
        pair_aux[j] = make_int2(last, sp);
        *leads = true;
    }
    pair_info[j] = make_int2(lead, first);
}

template <int LEVELS>
__global__ void __launch_bounds__(TB) pair_kernel(const typename BhKey<LEVELS>::type* __restrict__ K, int n, int2* __restrict__ pair_info,
                                                 int2* __restrict__ pair_aux, int32_t* __restrict__ tile_sums) {
    __shared__ uint32_t s_mask[LEVELS + 1][TILE_ROWS];     // [l][row]: pairs of the row with level < l   (l = 1 .. LEVELS)
    __shared__ unsigned long long s_any[LEVELS + 1];        // [l]: rows with a non-empty mask
    __shared__ int8_t s_lv[SCAN_TILE];
    __shared__ int s_w[TB / 32];
    const int npairs = n - 1;
    const int tile0 = blockIdx.x * SCAN_TILE;
    const int lane = bh_lane(), warp = threadIdx.x >> 5;

#pragma unroll 2
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const int j = tile0 + k * TB + threadIdx.x;
        // pairs past the end count as level -1: they bound every cell that reaches the end of the bodies
        const int lvk = j < npairs ? bh_shared_digits_t<LEVELS>(__ldg(K + j), __ldg(K + j + 1)) : -1;
        s_lv[k * TB + threadIdx.x] = (int8_t)lvk;
        const int row = k * (TB / 32) + warp;
#pragma unroll
        for (int l = 1; l <= LEVELS; ++l) {
            const uint32_t m = __ballot_sync(0xffffffffu, lvk < l);
            if (lane == l) s_mask[l][row] = m;   // LEVELS < 32: a different lane stores each threshold
        }
    }
    __syncthreads();
    for (int l = 1 + warp; l <= LEVELS; l += TB / 32) {
        const uint32_t lo = __ballot_sync(0xffffffffu, s_mask[l][lane] != 0u);
        const uint32_t hi = __ballot_sync(0xffffffffu, s_mask[l][lane + 32] != 0u);
        if (lane == 0) s_any[l] = ((unsigned long long)hi << 32) | lo;
    }
    __syncthreads();

    // (queueing the few pairs that need the keys and handling them together afterwards was measured: slower — the
    // common path, not the occasional galloping lane, is what the kernel's instructions go to)
    int leaders = 0;
#pragma unroll 1
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const int rel = k * TB + threadIdx.x;
        if (tile0 + rel >= npairs) continue;
        bool leads;
        pair_one<LEVELS>(K, n, tile0, rel, s_mask, s_any, s_lv, pair_info, pair_aux, &leads);
        leaders += leads;
    }
    // CTA reduction of the leader count
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) leaders += __shfl_xor_sync(0xffffffffu, leaders, o);
    if (lane == 0) s_w[warp] = leaders;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < TB / 32; ++w) t += s_w[w];
        tile_sums[blockIdx.x] = t;
    }
}

// ---- pass B: exclusive scan of the tile sums (single CTA), total -> num_cells -----------------
__global__ void __launch_bounds__(1024) scan_tiles_kernel(int32_t* tile_sums, int ntiles, BhDevScalars* sc) {
    __shared__ int s_w[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < ntiles; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < ntiles ? tile_sums[i] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, inc, o);
            if ((int)bh_lane() >= o) inc += t;
        }
        if (bh_lane() == 31) s_w[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = s_w[threadIdx.x], winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, winc, o);
                if ((int)bh_lane() >= o) winc += t;
            }
            s_w[threadIdx.x] = winc - w;  // exclusive warp bases
        }
        __syncthreads();
        const int carry = s_carry;
        const int excl = carry + s_w[threadIdx.x >> 5] + inc - v;
        if (i < ntiles) tile_sums[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) { sc->num_cells = s_carry; sc->root = -1; }
}

// ---- pass C: per-pair exclusive scan value = cell id of a leading pair -----------------------
__global__ void __launch_bounds__(TB) scan_pairs_kernel(const int2* __restrict__ pair_info, int n,
                                                       const int32_t* __restrict__ tile_sums,
                                                       int32_t* __restrict__ pair_scan) {
    const int npairs = n - 1;
    const int tile0 = blockIdx.x * SCAN_TILE;
    __shared__ int s_w[TB / 32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = tile_sums[blockIdx.x];
    __syncthreads();
#pragma unroll 1
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const int j = tile0 + k * TB + threadIdx.x;
        const int v = (j < npairs && pair_info[j].x == j) ? 1 : 0;
        const unsigned bal = __ballot_sync(0xffffffffu, v);
        const int in_warp = __popc(bal & ((1u << bh_lane()) - 1u));
        if (bh_lane() == 0) s_w[threadIdx.x >> 5] = __popc(bal);
        __syncthreads();
        int wbase = 0, total = 0;
#pragma unroll
        for (int w = 0; w < TB / 32; ++w) {
            int c = s_w[w];
            if (w < (int)(threadIdx.x >> 5)) wbase += c;
            total += c;
        }
        const int carry = s_carry;
        if (j < npairs) pair_scan[j] = carry + wbase + in_warp;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + total;
        __syncthreads();
    }
}

// ---- pass D: clear the child tables / arrival counters of the cells that exist --------------
__global__ void __launch_bounds__(TB) init_cells_kernel(int4* __restrict__ child4, const BhDevScalars* __restrict__ sc) {
    const int M = sc->num_cells;
    const int4 empty = make_int4(BH_CHILD_EMPTY, BH_CHILD_EMPTY, BH_CHILD_EMPTY, BH_CHILD_EMPTY);
    for (int i = blockIdx.x * TB + threadIdx.x; i < 2 * M; i += gridDim.x * TB) child4[i] = empty;   // consecutive threads, consecutive 16 bytes
}

// ---- pass E: every cell and every loose body links itself under its parent -------------------
template <int LEVELS>
__global__ void __launch_bounds__(TB) link_kernel(const typename BhKey<LEVELS>::type* __restrict__ K, int n,
                                                 const int2* __restrict__ pair_info, const int2* __restrict__ pair_aux,
                                                 const int32_t* __restrict__ pair_scan, int4* __restrict__ cell_meta,
                                                 int32_t* __restrict__ cell_child,
                                                 uint8_t* __restrict__ kid_lv, BhDevScalars* sc) {
    for (int i = blockIdx.x * TB + threadIdx.x; i < n; i += gridDim.x * TB) {
        const typename BhKey<LEVELS>::type ki = __ldg(K + i);
        const int lv_left = lv_pair<LEVELS>(K, i - 1, n);   // pair (i-1, i)
        const int lv_here = lv_pair<LEVELS>(K, i, n);       // pair (i, i+1)

        // (1) body i as a loose child (skipped when it belongs to a bucket)
        const int Lb = max(lv_left, lv_here);
        if (Lb >= 0 && Lb < LEVELS) {
            const int sp = (lv_left >= lv_here) ? i - 1 : i;
            const int parent = pair_scan[pair_info[sp].x];
            const int slot = (int)(ki >> (3 * LEVELS - 3 * (Lb + 1))) & 7;
            cell_child[(size_t)parent * 8 + slot] = (int)(0x80000000u | (uint32_t)i);
        }

        // (2) pair i, if it leads a cell: pair_kernel left its range and the pair that names its parent
        if (i < n - 1) {
            const int2 info = pair_info[i];
            if (info.x == i) {
                const int c = pair_scan[i];
                const int L = lv_here;
                const int l = info.y;
                const int2 aux = pair_aux[i];
                const int r = aux.x, sp = aux.y;
                int parent = -1, slot = 0;
                if (sp >= 0) {
                    const int Lp = lv_pair<LEVELS>(K, sp, n);
                    if (Lp >= L) atomicOr(&sc->err, BH_DERR_TREE);
                    parent = pair_scan[pair_info[sp].x];
                    slot = (int)(ki >> (3 * LEVELS - 3 * (Lp + 1))) & 7;
                    cell_child[(size_t)parent * 8 + slot] = c;
                    kid_lv[(size_t)parent * 8 + slot] = (uint8_t)(L | ((L == LEVELS) << 7));
                } else {
                    sc->root = c;
                }
                cell_meta[c] = make_int4(l, r - l + 1, L | ((L == LEVELS) << 8) | (slot << 12), parent);
            }
        }
    }
}

// ---- centre of mass: sums over contiguous ranges of the sorted bodies -----------------------------
// A cell is a contiguous range [first, first + count) of the Morton-sorted bodies (that is what the construction
// above emits), so its moments {m, m x, m y, m z} need no walk up the tree, no arrival counters, no float atomics
// (bench:158-189 adds 4 N floats per level with atomicAdd, which also makes the reference's sums depend on the
// schedule).  A cell of at most DIRECT_MAX bodies — most cells — adds its bodies up directly, in order.  A larger
// cell takes a DIFFERENCE OF PREFIX SUMS.  Three flat kernels:
//   com_scan_kernel   per block of 4,096 bodies: the exclusive LOCAL prefix (restarts at 0 in every block) at the
//                     start of every run of 16 bodies + the block total — 2 B/body written, not 32;
//   com_base_kernel   exclusive prefix of the block totals (one CTA);
//   com_cells_kernel  one thread per cell: direct sum, or (base[b2] - base[b1]) + (P(end) - P(first)) with
//                     P(i) = run prefix + the (< 16) bodies of the run before i, added in order; then the centre of
//                     mass as bench:181-186 and the cell's entries in the dense traversal lines.
// Sums are double: m x is exact in double (24 x 24 bits); a direct sum of <= 16 such terms and a difference of
// block-local prefixes of <= 4,096 terms are both far more accurate than any float32 summation order.  Every
// addition happens in a fixed order (16 bodies per thread in sequence, Hillis-Steele across a warp, warps / chunks
// in sequence), which oracle/bh_oracle.cpp:orc_tree_com reproduces operation for operation: equal bit for bit.
constexpr int CPT = 16;           // consecutive bodies per thread (summed in sequence)
constexpr int CT = 256;           // threads per scan block
constexpr int CB = CT * CPT;      // 4,096 bodies per scan block
constexpr int DIRECT_MAX = 16;    // cells up to this many bodies are summed directly

struct __align__(32) D4 { double m, x, y, z; };

__device__ __forceinline__ D4 d4_zero() { return D4{0.0, 0.0, 0.0, 0.0}; }
__device__ __forceinline__ D4 d4_add(const D4 a, const D4 b) {
    return D4{__dadd_rn(a.m, b.m), __dadd_rn(a.x, b.x), __dadd_rn(a.y, b.y), __dadd_rn(a.z, b.z)};
}
__device__ __forceinline__ D4 d4_sub(const D4 a, const D4 b) {
    return D4{__dsub_rn(a.m, b.m), __dsub_rn(a.x, b.x), __dsub_rn(a.y, b.y), __dsub_rn(a.z, b.z)};
}
__device__ __forceinline__ D4 d4_term(const float4 p) {   // exact products
    const double m = (double)p.w;
    return D4{m, __dmul_rn(m, (double)p.x), __dmul_rn(m, (double)p.y), __dmul_rn(m, (double)p.z)};
}
__device__ __forceinline__ D4 d4_shfl_up(const D4 v, int o) {
    return D4{__shfl_up_sync(0xffffffffu, v.m, o), __shfl_up_sync(0xffffffffu, v.x, o), __shfl_up_sync(0xffffffffu, v.y, o),
              __shfl_up_sync(0xffffffffu, v.z, o)};
}
// inclusive Hillis-Steele scan over the lanes of a warp: at distance o, lane i adds the value lane i-o held
// BEFORE this step (the oracle mirrors exactly this)
__device__ __forceinline__ D4 d4_warp_inclusive(D4 v) {
    const int lane = bh_lane();
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const D4 u = d4_shfl_up(v, o);
        if (lane >= o) v = d4_add(v, u);
    }
    return v;
}

// exclusive prefix over one CTA of NW warps given each thread's value; returns the thread's exclusive prefix and,
// through `total`, the CTA total (valid in every thread)
template <int NW>
__device__ __forceinline__ D4 d4_block_exclusive(const D4 v, D4* s_w /*NW*/, D4& total) {
    const int lane = bh_lane(), warp = threadIdx.x >> 5;
    const D4 incl = d4_warp_inclusive(v);
    D4 prev = d4_shfl_up(incl, 1);
    if (lane == 0) prev = d4_zero();
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    D4 before = d4_zero(), all = d4_zero();   // warps in sequence: ((W0 + W1) + W2) ...
#pragma unroll 1
    for (int w = 0; w < NW; ++w) {
        if (w == warp) before = all;
        all = d4_add(all, s_w[w]);
    }
    total = all;
    __syncthreads();
    return d4_add(before, prev);
}

__global__ void __launch_bounds__(CT) com_scan_kernel(const float4* __restrict__ posm, int n, D4* __restrict__ runpre,
                                                     D4* __restrict__ totals) {
    __shared__ D4 s_w[CT / 32];
    const int i0 = blockIdx.x * CB + threadIdx.x * CPT;
    D4 sum = d4_zero();   // this thread's run of bodies, in sequence
#pragma unroll 4
    for (int k = 0; k < CPT; ++k)
        if (i0 + k < n) sum = d4_add(sum, d4_term(__ldg(posm + i0 + k)));
    D4 total;
    const D4 pre = d4_block_exclusive<CT / 32>(sum, s_w, total);
    // runpre[r] = sum of the bodies of this block before run r; the run that starts at n (one past the last body) too
    if (i0 <= n) runpre[i0 / CPT] = pre;
    if (threadIdx.x == 0) totals[blockIdx.x] = total;
    if (threadIdx.x == 0 && blockIdx.x == gridDim.x - 1 && (blockIdx.x + 1) * CB == n) runpre[n / CPT] = d4_zero();   // n on a block boundary
}

// P(i): local prefix at body i = its run's prefix + the bodies of the run before i, in order
__device__ __forceinline__ D4 d4_prefix_at(const float4* __restrict__ posm, const D4* __restrict__ runpre, int i) {
    const int r0 = (i / CPT) * CPT;
    D4 v = runpre[i / CPT];
    for (int k = r0; k < i; ++k) v = d4_add(v, d4_term(__ldg(posm + k)));
    return v;
}
// moments of the bodies [first, end)
__device__ __forceinline__ D4 d4_range_sum(const float4* __restrict__ posm, const D4* __restrict__ runpre,
                                           const D4* __restrict__ base, int first, int end) {
    if (end - first <= DIRECT_MAX) {
        D4 v = d4_zero();
        for (int k = first; k < end; ++k) v = d4_add(v, d4_term(__ldg(posm + k)));
        return v;
    }
    return d4_add(d4_sub(base[end / CB], base[first / CB]), d4_sub(d4_prefix_at(posm, runpre, end), d4_prefix_at(posm, runpre, first)));
}

__global__ void __launch_bounds__(1024) com_base_kernel(const D4* __restrict__ totals, int nblocks, D4* __restrict__ base) {
    __shared__ D4 s_w[32];
    __shared__ D4 s_carry;
    if (threadIdx.x == 0) s_carry = d4_zero();
    __syncthreads();
    for (int c0 = 0; c0 < nblocks; c0 += 1024) {
        const int b = c0 + threadIdx.x;
        const D4 v = b < nblocks ? totals[b] : d4_zero();
        D4 total;
        const D4 pre = d4_block_exclusive<32>(v, s_w, total);
        const D4 carry = s_carry;
        if (b < nblocks) base[b] = d4_add(carry, pre);
        __syncthreads();
        if (threadIdx.x == 0) s_carry = d4_add(carry, total);
        __syncthreads();
    }
    if (threadIdx.x == 0) base[nblocks] = s_carry;
}

// ---- second moments for the quadrupole option (bh_params.flags & BH_FLAG_QUADRUPOLE) -------------------------
// The same prefix-sum construction with six running sums per body: m x'x', m x'y', m x'z', m y'y', m y'z', m z'z',
// x' = x - o with o the centre of the key cube (keeps the sums small; the shift is undone exactly in com_cells_kernel).
// The reference has monopoles only (bench:205-213); this is the accuracy knob of SURVEY §8f N4.
struct __align__(16) D6 { double v[6]; };
__device__ __forceinline__ D6 d6_zero() { D6 r; for (int k = 0; k < 6; ++k) r.v[k] = 0.0; return r; }
__device__ __forceinline__ D6 d6_add(const D6& a, const D6& b) { D6 r; for (int k = 0; k < 6; ++k) r.v[k] = __dadd_rn(a.v[k], b.v[k]); return r; }
__device__ __forceinline__ D6 d6_sub(const D6& a, const D6& b) { D6 r; for (int k = 0; k < 6; ++k) r.v[k] = __dsub_rn(a.v[k], b.v[k]); return r; }
__device__ __forceinline__ D6 d6_shfl_up(const D6& a, int o) { D6 r; for (int k = 0; k < 6; ++k) r.v[k] = __shfl_up_sync(0xffffffffu, a.v[k], o); return r; }
__device__ __forceinline__ D6 d6_term(const float4 p, double ox, double oy, double oz) {
    const double m = (double)p.w, x = (double)p.x - ox, y = (double)p.y - oy, z = (double)p.z - oz;
    D6 r;
    r.v[0] = m * x * x; r.v[1] = m * x * y; r.v[2] = m * x * z; r.v[3] = m * y * y; r.v[4] = m * y * z; r.v[5] = m * z * z;
    return r;
}
template <int NW>
__device__ __forceinline__ D6 d6_block_exclusive(const D6 v, D6* s_w, D6& total) {
    const int lane = bh_lane(), warp = threadIdx.x >> 5;
    D6 incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const D6 u = d6_shfl_up(incl, o);
        if (lane >= o) incl = d6_add(incl, u);
    }
    D6 prev = d6_shfl_up(incl, 1);
    if (lane == 0) prev = d6_zero();
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    D6 before = d6_zero(), all = d6_zero();
#pragma unroll 1
    for (int w = 0; w < NW; ++w) {
        if (w == warp) before = all;
        all = d6_add(all, s_w[w]);
    }
    total = all;
    __syncthreads();
    return d6_add(before, prev);
}

__device__ __forceinline__ void cube_centre(const BhDevScalars* sc, double& ox, double& oy, double& oz) {
    const float half = __fmul_rn(0.5f, fmaxf(__fsub_rn(sc->bounds[3], sc->bounds[0]), 1.0f));
    ox = (double)__fadd_rn(sc->bounds[0], half); oy = (double)__fadd_rn(sc->bounds[1], half); oz = (double)__fadd_rn(sc->bounds[2], half);
}

__global__ void __launch_bounds__(CT) quad_scan_kernel(const float4* __restrict__ posm, int n, const BhDevScalars* __restrict__ sc,
                                                      D6* __restrict__ runpre, D6* __restrict__ totals) {
    __shared__ D6 s_w[CT / 32];
    double ox, oy, oz;
    cube_centre(sc, ox, oy, oz);
    const int i0 = blockIdx.x * CB + threadIdx.x * CPT;
    D6 sum = d6_zero();
#pragma unroll 2
    for (int k = 0; k < CPT; ++k)
        if (i0 + k < n) sum = d6_add(sum, d6_term(__ldg(posm + i0 + k), ox, oy, oz));
    D6 total;
    const D6 pre = d6_block_exclusive<CT / 32>(sum, s_w, total);
    if (i0 <= n) runpre[i0 / CPT] = pre;
    if (threadIdx.x == 0) totals[blockIdx.x] = total;
    if (threadIdx.x == 0 && blockIdx.x == gridDim.x - 1 && (blockIdx.x + 1) * CB == n) runpre[n / CPT] = d6_zero();
}

__device__ __forceinline__ D6 d6_prefix_at(const float4* __restrict__ posm, const D6* __restrict__ runpre, int i, double ox,
                                           double oy, double oz) {
    const int r0 = (i / CPT) * CPT;
    D6 v = runpre[i / CPT];
    for (int k = r0; k < i; ++k) v = d6_add(v, d6_term(__ldg(posm + k), ox, oy, oz));
    return v;
}
__device__ __forceinline__ D6 d6_range_sum(const float4* __restrict__ posm, const D6* __restrict__ runpre,
                                           const D6* __restrict__ base, int first, int end, double ox, double oy, double oz) {
    if (end - first <= DIRECT_MAX) {
        D6 v = d6_zero();
        for (int k = first; k < end; ++k) v = d6_add(v, d6_term(__ldg(posm + k), ox, oy, oz));
        return v;
    }
    return d6_add(d6_sub(base[end / CB], base[first / CB]),
                  d6_sub(d6_prefix_at(posm, runpre, end, ox, oy, oz), d6_prefix_at(posm, runpre, first, ox, oy, oz)));
}

__global__ void __launch_bounds__(1024) quad_base_kernel(const D6* __restrict__ totals, int nblocks, D6* __restrict__ base) {
    __shared__ D6 s_w[32];
    __shared__ D6 s_carry;
    if (threadIdx.x == 0) s_carry = d6_zero();
    __syncthreads();
    for (int c0 = 0; c0 < nblocks; c0 += 1024) {
        const int b = c0 + threadIdx.x;
        const D6 v = b < nblocks ? totals[b] : d6_zero();
        D6 total;
        const D6 pre = d6_block_exclusive<32>(v, s_w, total);
        const D6 carry = s_carry;
        if (b < nblocks) base[b] = d6_add(carry, pre);
        __syncthreads();
        if (threadIdx.x == 0) s_carry = d6_add(carry, total);
        __syncthreads();
    }
    if (threadIdx.x == 0) base[nblocks] = s_carry;
}

// what the quadrupole option adds to com_cells_kernel's arguments
struct QuadArgs {
    const D6* local2; const D6* base2;
    float4* cell_quad;   // 2 float4 per cell: {xx, xy, xz, yy}, {yz, zz, 0, 0}
    float4* kid_quad;    // 2 float4 per dense child entry (same index as kid_src)
};

template <bool QUAD>
__global__ void __launch_bounds__(TB) com_cells_kernel(const float4* __restrict__ posm, const int4* __restrict__ cell_meta,
                                                      const int32_t* __restrict__ cell_child, const D4* __restrict__ local,
                                                      const D4* __restrict__ base, float4* __restrict__ com,
                                                      float4* __restrict__ kid_src, uint2* __restrict__ kid_info, BhDevScalars* sc,
                                                      QuadArgs qa) {
    const int M = sc->num_cells;
    const int4* child4 = reinterpret_cast<const int4*>(cell_child);
    // squared width of a level-L cell: root_w^2 with 2L taken off the exponent (exact); root_w is the key
    // grid's size, clamped like bench:52
    const float root_w = fmaxf(__fsub_rn(sc->bounds[3], sc->bounds[0]), 1.0f);
    const int root_w2_bits = __float_as_int(__fmul_rn(root_w, root_w));
    for (int c = blockIdx.x * TB + threadIdx.x; c < M; c += gridDim.x * TB) {
        const int4 mt = __ldg(cell_meta + c);
        const int first = mt.x, end = mt.x + mt.y;
        const D4 s = d4_range_sum(posm, local, base, first, end);
        const float m = (float)s.m, sx = (float)s.x, sy = (float)s.y, sz = (float)s.z;
        const float inv = (m > 1e-6f) ? __fdiv_rn(1.0f, m) : 0.0f;   // bench:181-183
        const float4 cm = make_float4(__fmul_rn(sx, inv), __fmul_rn(sy, inv), __fmul_rn(sz, inv), m);
        com[c] = cm;
        float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), q1 = q0;
        if (QUAD) {
            // second moments about the cube centre o -> about the cell's centre of mass (in double: the shift is exact
            // to 1e-16, the quantities differ by ~1e-4 at worst), then the traceless form Q = 3 P - tr(P) I
            double ox, oy, oz;
            cube_centre(sc, ox, oy, oz);
            const D6 s2 = d6_range_sum(posm, qa.local2, qa.base2, first, end, ox, oy, oz);
            const double M = s.m, im = M > 0.0 ? 1.0 / M : 0.0;
            const double cx = s.x * im - ox, cy = s.y * im - oy, cz = s.z * im - oz;
            const double pxx = s2.v[0] - M * cx * cx, pxy = s2.v[1] - M * cx * cy, pxz = s2.v[2] - M * cx * cz;
            const double pyy = s2.v[3] - M * cy * cy, pyz = s2.v[4] - M * cy * cz, pzz = s2.v[5] - M * cz * cz;
            const double tr = pxx + pyy + pzz;
            q0 = make_float4((float)(3.0 * pxx - tr), (float)(3.0 * pxy), (float)(3.0 * pxz), (float)(3.0 * pyy - tr));
            q1 = make_float4((float)(3.0 * pyz), (float)(3.0 * pzz - tr), 0.f, 0.f);
            qa.cell_quad[2 * (size_t)c] = q0;
            qa.cell_quad[2 * (size_t)c + 1] = q1;
        }
        const bool bucket = (mt.z >> 8) & 1;
        int nkids = 1;   // buckets are never opened as cells: the field is unused for them
        if (!bucket) {   // the loose bodies of this cell: their entries in its dense line
            const int4 lo = __ldg(child4 + 2 * c), hi = __ldg(child4 + 2 * c + 1);
            const int e[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
            int r = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (e[q] == BH_CHILD_EMPTY) continue;
                if (e[q] < 0) {
                    kid_src[(size_t)c * 8 + r] = __ldg(posm + (e[q] & 0x7FFFFFFF));
                    kid_info[(size_t)c * 8 + r] = make_uint2(0x7FFFFFFFu, BH_KID_BODY_W2);
                }
                ++r;
            }
            nkids = r;
        }
        const unsigned word = ((unsigned)c << 3) | (unsigned)(nkids - 1);
        const int p = mt.w;
        if (p < 0) { sc->root_word = word; continue; }
        // the parent's view of this cell: source + the word the traversal pushes to open it later
        const int4 lo = __ldg(child4 + 2 * p), hi = __ldg(child4 + 2 * p + 1);
        const int e[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        const int myslot = (mt.z >> 12) & 7;
        int rank = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) rank += (q < myslot && e[q] != BH_CHILD_EMPTY);
        kid_src[(size_t)p * 8 + rank] = cm;
        kid_info[(size_t)p * 8 + rank] = make_uint2(word | (bucket ? BH_KID_BUCKET : 0u), (unsigned)(root_w2_bits - ((mt.z & 0xFF) << 24)));
        if (QUAD) {
            qa.kid_quad[2 * ((size_t)p * 8 + rank)] = q0;
            qa.kid_quad[2 * ((size_t)p * 8 + rank) + 1] = q1;
        }
    }
}

inline int capped_grid(int64_t work_items, int per_block) {
    int64_t b = (work_items + per_block - 1) / per_block;
    if (b > BH_NUM_SMS_FALLBACK * 8) b = BH_NUM_SMS_FALLBACK * 8;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace

// keys: ascending u32 keys of 10 digits (levels == 10, the reference's 30-bit key) or u64 keys of 20 digits
// (levels == 20, bh_params.key_bits = 60)
int bh_tree_launch(const void* keys, int levels, int64_t n64, int2* pair_info, int2* pair_aux, int32_t* pair_scan,
                   int32_t* scan_block_sums, int4* cell_meta, int32_t* cell_child,
                   uint8_t* kid_lv, BhDevScalars* sc, cudaStream_t st) {
    const int n = (int)n64;
    if (n < 2) return 0;
    const int ntiles = (n - 1 + SCAN_TILE - 1) / SCAN_TILE;
    if (levels == 20) pair_kernel<20><<<ntiles, TB, 0, st>>>((const uint64_t*)keys, n, pair_info, pair_aux, scan_block_sums);
    else pair_kernel<10><<<ntiles, TB, 0, st>>>((const uint32_t*)keys, n, pair_info, pair_aux, scan_block_sums);
    scan_tiles_kernel<<<1, 1024, 0, st>>>(scan_block_sums, ntiles, sc);
    scan_pairs_kernel<<<ntiles, TB, 0, st>>>(pair_info, n, scan_block_sums, pair_scan);
    init_cells_kernel<<<capped_grid(n, TB), TB, 0, st>>>(reinterpret_cast<int4*>(cell_child), sc);
    if (levels == 20)
        link_kernel<20><<<capped_grid(n, TB), TB, 0, st>>>((const uint64_t*)keys, n, pair_info, pair_aux, pair_scan, cell_meta, cell_child, kid_lv, sc);
    else
        link_kernel<10><<<capped_grid(n, TB), TB, 0, st>>>((const uint32_t*)keys, n, pair_info, pair_aux, pair_scan, cell_meta, cell_child, kid_lv, sc);
    return (int)cudaGetLastError();
}

// com_scratch: 32-byte aligned, bh_com_scratch_bytes(n) bytes — run prefixes (n / 16 + 2), block totals, block bases
static inline size_t com_runs(int64_t n) { return (size_t)(n / CPT) + 2; }
size_t bh_com_scratch_bytes(int64_t n) {
    const size_t nblocks = (size_t)((n + CB - 1) / CB) + 1;
    return sizeof(D4) * (com_runs(n) + 2 * nblocks + 4);
}
// the same for the six second moments of the quadrupole option
size_t bh_quad_scratch_bytes(int64_t n) {
    const size_t nblocks = (size_t)((n + CB - 1) / CB) + 1;
    return sizeof(D6) * (com_runs(n) + 2 * nblocks + 4);
}

// The prefix sums need the sorted bodies only: they can run beside the tree construction (bh_com_prefix_launch on
// another stream); bh_com_cells_launch needs both the tree and the sums.  quad_scratch != nullptr: also the second moments.
int bh_com_prefix_launch(const float4* posm, int64_t n64, void* com_scratch, void* quad_scratch, const BhDevScalars* sc,
                         cudaStream_t st) {
    const int n = (int)n64;
    if (n < 2) return 0;
    const int nblocks = (n + CB - 1) / CB;
    D4* local = reinterpret_cast<D4*>(com_scratch);
    D4* totals = local + com_runs(n);
    D4* base = totals + nblocks + 1;
    com_scan_kernel<<<nblocks, CT, 0, st>>>(posm, n, local, totals);
    com_base_kernel<<<1, 1024, 0, st>>>(totals, nblocks, base);
    if (quad_scratch) {
        D6* local2 = reinterpret_cast<D6*>(quad_scratch);
        D6* totals2 = local2 + com_runs(n);
        D6* base2 = totals2 + nblocks + 1;
        quad_scan_kernel<<<nblocks, CT, 0, st>>>(posm, n, sc, local2, totals2);
        quad_base_kernel<<<1, 1024, 0, st>>>(totals2, nblocks, base2);
    }
    return (int)cudaGetLastError();
}

int bh_com_cells_launch(const float4* posm, int64_t n64, const int4* cell_meta, const int32_t* cell_child, void* com_scratch,
                        float4* cell_com, float4* kid_src, uint2* kid_info, BhDevScalars* sc, void* quad_scratch,
                        float4* cell_quad, float4* kid_quad, cudaStream_t st) {
    const int n = (int)n64;
    if (n < 2) return 0;
    const int nblocks = (n + CB - 1) / CB;
    const D4* local = reinterpret_cast<const D4*>(com_scratch);
    const D4* base = local + com_runs(n) + nblocks + 1;
    QuadArgs qa{nullptr, nullptr, cell_quad, kid_quad};
    if (quad_scratch) {
        qa.local2 = reinterpret_cast<const D6*>(quad_scratch);
        qa.base2 = qa.local2 + com_runs(n) + nblocks + 1;
        com_cells_kernel<true><<<capped_grid(n, TB), TB, 0, st>>>(posm, cell_meta, cell_child, local, base, cell_com, kid_src, kid_info, sc, qa);
    } else {
        com_cells_kernel<false><<<capped_grid(n, TB), TB, 0, st>>>(posm, cell_meta, cell_child, local, base, cell_com, kid_src, kid_info, sc, qa);
    }
    return (int)cudaGetLastError();
}

int bh_com_launch(const float4* posm, int64_t n64, const int4* cell_meta, const int32_t* cell_child, void* com_scratch,
                  float4* cell_com, float4* kid_src, uint2* kid_info, BhDevScalars* sc, void* quad_scratch, float4* cell_quad,
                  float4* kid_quad, cudaStream_t st) {
    int e = bh_com_prefix_launch(posm, n64, com_scratch, quad_scratch, sc, st);
    if (e) return e;
    return bh_com_cells_launch(posm, n64, cell_meta, cell_child, com_scratch, cell_com, kid_src, kid_info, sc, quad_scratch, cell_quad,
                               kid_quad, st);
}
