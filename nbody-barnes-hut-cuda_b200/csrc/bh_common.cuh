// bh_common.cuh — shared device helpers and the context layout of the Barnes-Hut engine.
// Target: sm_100a (B200).  FP32 SIMT + integer work; no tensor cores on this path
// (BASELINE.json north_star: traversal is not a dense contraction).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include "../../include/bh.h"

#define BH_KEY_BITS 30
#define BH_MAX_LEVEL 10
#define BH_GROUP 32          // bodies per traversal chunk (one warp, one body per lane)
#define BH_NUM_SMS_FALLBACK 148   // B200: 2 dies x 74 SMs; grids are sized in multiples of it

// device error flag bits (BH_STAT_DEVICE_ERROR)
#define BH_DERR_SORT_SPIN   1   // onesweep look-back exceeded its spin budget
#define BH_DERR_STACK       2   // traversal stack guard tripped
#define BH_DERR_LOOP        4   // traversal iteration guard tripped
#define BH_DERR_TREE        8   // builder found an inconsistent range
#define BH_DERR_LET_OVERFLOW 16  // locally-essential-tree export ran out of output / queue space

// kid_info[8c + r] = {x: child cell id << 3 | its child count - 1 | BH_KID_BUCKET, y: float bits of the child cell's
// squared width root_w^2 * 4^-level}; a loose body has x = 0xFFFFFFFF... and y = bits of -1.0f, which passes
// every acceptance test `w^2 < theta^2 (d^2 + softening)` — bodies are always sources
#define BH_KID_BUCKET     0x80000000u   // child cell is an identical-key bucket (contiguous body range, never opened)
#define BH_KID_BODY_W2    0xBF800000u   // -1.0f

struct BhDevScalars {        // one small device struct, zeroed/filled by kernels
    float bounds[6];         // as d_bounds (nbody_v5_bench.cu:149-154)
    int   num_cells;         // total of the leader-flag scan
    int   root;              // id of the parentless cell
    unsigned int err;        // BH_DERR_* bits (sticky)
    unsigned int group_ticket;   // dynamic group scheduler of the force kernel
    unsigned long long inter_cell;
    unsigned long long inter_body;
    unsigned int max_stack;
    unsigned int root_word;      // stack word of the root: id << 3 | children - 1 (written by the centre-of-mass pass)
    unsigned int bbox_enc[6];    // order-preserving uint encoding of min/max during reduction
    // heavy-first scheduling of traversal chunks (bh_force.cu): chunks whose interaction list was long
    // in the previous step are handed out first in this one (costs are temporally coherent)
    unsigned int epoch;          // force launches since import; parity selects the list being read
    unsigned int heavy_n[2];
    unsigned int heavy_thresh;   // list entries above which a chunk is recorded as heavy
    unsigned long long entries_total;   // list entries of the last force launch (all sub-groups)
};

// ---- small device helpers ---------------------------------------------------------------
__device__ __forceinline__ unsigned int bh_lane() { return threadIdx.x & 31u; }

// order-preserving float <-> uint map so atomicMin/atomicMax work on floats
__device__ __forceinline__ unsigned int bh_f2ord(float f) {
    unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float bh_ord2f(unsigned int u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

// number of leading 3-bit digits two 30-bit keys share (0..10); 10 <=> equal keys
__device__ __forceinline__ int bh_shared_digits(uint32_t a, uint32_t b) {
    uint32_t x = a ^ b;
    return x == 0 ? BH_MAX_LEVEL : (__clz((int)x) - (32 - BH_KEY_BITS)) / 3;
}

// keys of LEVELS 3-bit digits: 10 = the reference's 30-bit key in a u32, 20 = that key extended by 10 more
// bits per axis (bh_params.key_bits = 60) in a u64, reference key on top
template <int LEVELS> struct BhKey;
template <> struct BhKey<10> { typedef uint32_t type; };
template <> struct BhKey<20> { typedef uint64_t type; };
template <int LEVELS>
__device__ __forceinline__ int bh_shared_digits_t(typename BhKey<LEVELS>::type a, typename BhKey<LEVELS>::type b);
template <>
__device__ __forceinline__ int bh_shared_digits_t<10>(uint32_t a, uint32_t b) { return bh_shared_digits(a, b); }
template <>
__device__ __forceinline__ int bh_shared_digits_t<20>(uint64_t a, uint64_t b) {
    const uint64_t x = a ^ b;
    return x == 0 ? 20 : (__clzll((long long)x) - 4) / 3;
}

#define BH_CUDA_TRY(expr)                                  \
    do {                                                   \
        cudaError_t _e = (expr);                           \
        if (_e != cudaSuccess) return (int)_e;             \
    } while (0)

// ---- kernel launchers (one translation unit each) ---------------------------------------
struct BhSortPlan {
    int64_t n;
    int num_tiles;
    size_t hist_bytes;      // passes * RADIX counters
    size_t lookback_bytes;  // passes * tiles * RADIX status words
    size_t ticket_bytes;    // passes tickets
    size_t total_bytes;
};
BhSortPlan bh_sort_plan(int64_t n);
int bh_sort_passes(int bits);   // radix passes a `bits`-wide key takes (the result lands in q iff even)
// Pass 0 reads (keys_src, vals_src) -> (p); then p -> q -> p ...; result in q iff the pass count
// is even (*result_in_q).  keys_q may alias keys_src.  tmp holds histogram + look-back + tickets
// (bh_sort_plan(n).total_bytes).  If vals_in_is_iota pass 0 generates value = index.
int bh_sort_pairs_launch(const uint32_t* keys_src, const uint32_t* vals_src, uint32_t* keys_p, uint32_t* vals_p,
                         uint32_t* keys_q, uint32_t* vals_q, int64_t n, int begin_bit, int end_bit, void* tmp,
                         bool vals_in_is_iota, unsigned int* err_flag, int* result_in_q, cudaStream_t st);

// same, and the last pass also reorders the bodies (out[i] = in[sorted value i]) — the step's Morton reorder
int bh_sort_pairs_move_launch(const uint32_t* keys_src, const uint32_t* vals_src, uint32_t* keys_p, uint32_t* vals_p,
                              uint32_t* keys_q, uint32_t* vals_q, int64_t n, int begin_bit, int end_bit, void* tmp,
                              bool vals_in_is_iota, unsigned int* err_flag, int* result_in_q, const float4* posm_in,
                              const float4* vel_in, const int32_t* ids_in, float4* posm_out, float4* vel_out,
                              int32_t* ids_out, cudaStream_t st);

int bh_keys_launch(const float4* posm, int64_t n, BhDevScalars* sc, uint32_t* keys, cudaStream_t st);
// 60-bit keys: hi = the reference key, lo = ten more bits per axis from the fractional part of the same float
int bh_keys60_launch(const float4* posm, int64_t n, BhDevScalars* sc, uint32_t* hi, uint32_t* lo, cudaStream_t st);
int bh_gather_u32_launch(const uint32_t* src, const uint32_t* perm, uint32_t* dst, int64_t n, cudaStream_t st);
// keys64[i] = hi_sorted[i] << 30 | lo_unsorted[perm[i]]
int bh_combine_keys_launch(const uint32_t* hi_sorted, const uint32_t* lo_unsorted, const uint32_t* perm, uint64_t* keys64,
                           int64_t n, cudaStream_t st);
int bh_bounds_enc_launch(const float4* posm, int64_t n, BhDevScalars* sc, cudaStream_t st);   // positions -> min/max images
int bh_bounds_finish_launch(BhDevScalars* sc, cudaStream_t st);                                 // images -> cube (consumes them)
int bh_reorder_launch(const float4* posm_in, const float4* vel_in, const int32_t* ids_in,
                      const uint32_t* perm, float4* posm_out, float4* vel_out, int32_t* ids_out,
                      int64_t n, cudaStream_t st);
// kid_lv: 8 bytes per cell, digit-indexed like cell_child — level | bucket<<7 of child cells.
int bh_tree_launch(const void* keys, int levels, int64_t n, int2* pair_info, int2* pair_aux, int32_t* pair_scan,
                   int32_t* scan_block_sums, int4* cell_meta, int32_t* cell_child,
                   uint8_t* kid_lv, BhDevScalars* sc, cudaStream_t st);
// kid_src / kid_info: 8 entries per cell, DENSE (the r-th existing child in digit order sits at 8c + r) — the SOURCE
// each child contributes when its parent is opened (a loose body's {x,y,z,m}, a child cell's {com,mass}) and
// {stack word of a child cell = id << 3 | its child count - 1 | BH_KID_BUCKET, squared width as float bits}.
size_t bh_com_scratch_bytes(int64_t n);
size_t bh_quad_scratch_bytes(int64_t n);   // second moments (quadrupole option)
// quad_scratch / cell_quad / kid_quad: nullptr without BH_FLAG_QUADRUPOLE.  cell_quad: 2 float4 per cell {xx,xy,xz,yy},
// {yz,zz,0,0} (traceless, about the centre of mass); kid_quad: the same per dense child entry (index as kid_src).
int bh_com_launch(const float4* posm, int64_t n, const int4* cell_meta, const int32_t* cell_child, void* com_scratch,
                  float4* cell_com, float4* kid_src, uint2* kid_info, BhDevScalars* sc, void* quad_scratch, float4* cell_quad,
                  float4* kid_quad, cudaStream_t st);
// the two halves of bh_com_launch: prefix sums over the sorted bodies (independent of the tree), then one thread per cell
int bh_com_prefix_launch(const float4* posm, int64_t n, void* com_scratch, void* quad_scratch, const BhDevScalars* sc, cudaStream_t st);
int bh_com_cells_launch(const float4* posm, int64_t n, const int4* cell_meta, const int32_t* cell_child, void* com_scratch,
                        float4* cell_com, float4* kid_src, uint2* kid_info, BhDevScalars* sc, void* quad_scratch,
                        float4* cell_quad, float4* kid_quad, cudaStream_t st);
// Optional tail of the traversal: the kick-drift-clamp of bench:227-249 for a chunk as soon as its accelerations are
// final (same arithmetic as integrate_kernel, bit for bit), and the min/max of the new positions.  posm_out == nullptr:
// the traversal only writes acc (the update is a launch of its own).
struct BhFusedUpdate {
    const float4* vel_s; const int32_t* ids_s;
    float4* posm_out; float4* vel_out; int32_t* ids_out;
    float dt, max_speed;
};
// heavy_list, heavy_flag: 2 * max_chunks u32 each (see BhDevScalars::epoch)
// ids: nullptr, or per-body ids where id < 0 marks a ghost (a source whose own acceleration is not wanted)
int bh_force_launch(const float4* posm, const void* keys, int levels, const int32_t* ids, int64_t n, int64_t first_body,
                    int64_t body_count, const int4* cell_meta, const float4* cell_com,
                    const float4* kid_src, const uint2* kid_info,
                    float4* acc, BhDevScalars* sc, uint32_t* heavy_list, uint32_t* heavy_flag, int64_t max_chunks,
                    float theta, float softening, float G, float split_alpha, int num_sms,
                    // cross-tree pass: the tree (scalars, cells, the bodies its buckets index) of ANOTHER body set;
                    // nullptr/nullptr/0 = the ordinary pass over the bodies' own tree
                    const float4* src_posm, const BhDevScalars* tree_sc, int accumulate,
                    // quadrupole option: per-cell and per-child-entry moments (bh_com_launch); nullptr = monopoles only
                    const float4* cell_quad, const float4* kid_quad, cudaStream_t st, const BhFusedUpdate* fused = nullptr);
int bh_force_prepare();
int bh_integrate_launch(const float4* posm_s, const float4* vel_s, const int32_t* ids_s, const float4* acc,
                        float4* posm, float4* vel, int32_t* ids, int64_t first_body, int64_t body_count,
                        float dt, float max_speed, BhDevScalars* sc /* min/max of the new positions accumulate here */,
                        cudaStream_t st);
int bh_import_launch(const float* px, const float* py, const float* pz, const float* vx,
                     const float* vy, const float* vz, const float* m, int64_t n, float4* posm,
                     float4* vel, int32_t* ids, cudaStream_t st);
int bh_import_pos_launch(const float* px, const float* py, const float* pz, int64_t n, float4* posm, cudaStream_t st);
int bh_import_mass_launch(const float* m, int64_t n, float4* posm, cudaStream_t st);
int bh_import_vel_launch(const float* vx, const float* vy, const float* vz, int64_t n, float4* vel, int32_t* ids, cudaStream_t st);
int bh_reorder_posm_launch(const float4* posm_in, const uint32_t* perm, float4* posm_out, int64_t n, cudaStream_t st);
int bh_reorder_rest_launch(const float4* vel_in, const int32_t* ids_in, const uint32_t* perm, float4* vel_out, int32_t* ids_out,
                           int64_t n, cudaStream_t st);
int bh_where_launch(const int32_t* ids, int64_t n, int32_t* where, cudaStream_t st);
int bh_gather3_launch(const float4* src, const int32_t* where, int64_t n, float* ox, float* oy, float* oz, cudaStream_t st);
int bh_export_launch(const float4* posm, const float4* vel, const float4* acc, const int32_t* ids,
                     int64_t n, float* px, float* py, float* pz, float* vx, float* vy, float* vz,
                     float* ax, float* ay, float* az, cudaStream_t st);
int bh_direct_launch(const float4* posm, int64_t n, const int32_t* sample_slots, int k,
                     float softening, float G, double* acc_out, cudaStream_t st);
int bh_energy_launch(const float4* posm, const float4* vel, int64_t n, float softening, float G,
                     double* ke_pe /*2 doubles, zeroed by the launcher*/, cudaStream_t st);
int bh_visuals_launch(const float4* posm, const float4* vel, const int32_t* ids, int64_t n, float* vbo_p, float* vbo_c,
                      cudaStream_t st);
int bh_momentum_launch(const float4* posm, const float4* vel, int64_t n, double* out7, cudaStream_t st);
int bh_let_export_launch(const int4* cell_meta, const int32_t* cell_child, const float4* cell_com, const float4* kid_src,
                         const uint8_t* kid_lv, const float4* posm, BhDevScalars* sc, const float* boxes_dev,
                         const float* hull_dev, int npeers, int K, float4* out, unsigned int* out_count, long long cap, int2* queue, unsigned int* qcounts,
                         long long qcap, float theta, float softening, float root_w, int levels, cudaStream_t st);
int bh_domain_boxes_launch(const uint32_t* keys, const float4* posm, long long n, const uint32_t* cuts_dev, int K,
                           float* out_dev, int* counts_dev, unsigned int* enc_dev, cudaStream_t st);
int bh_compact_real_launch(const float4* posm, const float4* vel, const int32_t* ids, const float4* acc, int64_t n,
                           int32_t* tile_scratch, float4* posm_out, float4* vel_out, int32_t* ids_out, cudaStream_t st);
