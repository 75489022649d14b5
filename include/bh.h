/*
 * bh.h — C ABI of the B200-native Barnes-Hut step engine.
 *
 * Drop-in boundary for the hot path of bgcarmin/NBody-Barnes-Hut-CUDA:
 *   void simulationStep()            nbody_v5_bench.cu:255-283 (nbody_v5.cu:298-325)
 * The reference has no plugin/FFI interface; its de-facto boundary is that
 * argument-less function working in place on 16 file-scope device pointers
 * (nbody_v5_bench.cu:31-40).  Every entry point below names the reference
 * lines it replaces.  Plain pointers and sizes only; no C++/torch types.
 *
 * Conventions
 *   - every function returns 0 on success, a positive cudaError_t value for
 *     CUDA failures, or a negative BH_E_* code; nothing throws or exits.
 *   - `stream` arguments are a cudaStream_t passed as void* (NULL = legacy
 *     default stream, which is what the reference uses everywhere).
 *   - one host thread per context; a context is bound to one device.
 *   - body slot i is body i for the whole run at this boundary
 *     (nbody_v5_bench.cu:31-40: the reference never permutes bodies); the
 *     Morton-ordered internal layout is private to the context.
 */
#ifndef BH_H_
#define BH_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BH_ABI_VERSION 2   /* 2: bh_mg_*, bh_step_part, BH_FLAG_QUADRUPOLE, BH_DBG_CELL_QUAD; checkpoint header v2 */

/* negative error codes (positive values are cudaError_t) */
#define BH_E_INVAL      (-1)  /* bad argument                                  */
#define BH_E_NOMEM      (-2)  /* host allocation failed                        */
#define BH_E_STATE      (-3)  /* call order violated (e.g. step before import) */
#define BH_E_UNSUPPORTED (-4) /* parameter combination not implemented         */
#define BH_E_DEVICE     (-5)  /* a kernel raised its device-side error flag    */
#define BH_E_IO         (-6)  /* file could not be opened / is not a checkpoint */

/* Simulation parameters — the reference's compile-time constants
 * (nbody_v5_bench.cu:13-18) plus the tree-shape knobs it hard-codes.       */
typedef struct bh_params {
    float theta;      /* THETA      0.5f   nbody_v5_bench.cu:15 */
    float G;          /* G_CONST    0.5f   nbody_v5_bench.cu:14 */
    float dt;         /* DT         0.02f  nbody_v5_bench.cu:16 */
    float softening;  /* SOFTENING  50.0f  nbody_v5_bench.cu:17 (added to r^2) */
    float max_speed;  /* MAX_SPEED  500.0f nbody_v5_bench.cu:18 */
    int   key_bits;   /* 30: 10 bits/axis Morton key, nbody_v5_bench.cu:58-61 (default).
                         60: the same key extended by 10 more bits per axis taken from the
                         fractional part of the reference's own float (p-min)/size*1023 — the top
                         30 bits stay the reference key, the tree gains 10 more levels.  For body
                         counts where a 1024^3 grid holds many bodies per cell.   */
    int   leaf_cap;   /* 1: one body per leaf, as nbody_v5_bench.cu:100-104   */
    int   flags;      /* BH_FLAG_* */
    float group_split;/* 0.5: a 32-body traversal group is cut at its coarsest key
                         boundary while ext(A)+ext(B) < group_split*ext(AuB); 0 = never */
} bh_params;

#define BH_FLAG_NO_GRAPH    1  /* launch kernels directly instead of a CUDA graph  */
#define BH_FLAG_PHASE_TIMER 2  /* record per-phase cudaEvents (implies NO_GRAPH)   */
#define BH_FLAG_QUADRUPOLE  4  /* accepted cells act with their traceless quadrupole as well (set before bh_create).
                                * The reference has monopoles only (nbody_v5_bench.cu:205-213) and that stays the
                                * default; this is an accuracy / throughput knob: the same acceptance test, a 3-6x
                                * smaller error on disc-like systems (see DESIGN.md 5.6 for where it does not pay).  Not available in the locally-essential-tree calls.            */

typedef struct bh_ctx bh_ctx;

/* Fill *p with the reference defaults listed above. */
void bh_default_params(bh_params* p);
int  bh_abi_version(void);
/* bodies per traversal chunk (Morton-consecutive bodies handled by one warp); multi-GPU slices are
 * whole chunks.                                                              */
int  bh_group_size(void);
const char* bh_error_string(int code);

/* Replaces the 16 cudaMalloc calls + H2D copies of main()
 * (nbody_v5_bench.cu:311-335): allocates every buffer once for n_max bodies
 * (~370 B/body, +300 B/body with BH_FLAG_QUADRUPOLE; 0 < n_max < 2^28: the
 * traversal's stack words are cell id << 3 plus a flag bit).               */
int  bh_create(bh_ctx** out, int64_t n_max, const bh_params* params, int device);
/* Replaces nbody_v5_bench.cu:372-387. */
void bh_destroy(bh_ctx* ctx);

/* Load the reference's SoA state (nbody_v5_bench.cu:32-35, 329-335).
 * DEVICE pointers, caller-owned, not retained.  Body i keeps id i.        */
int  bh_import_soa(bh_ctx* ctx,
                   const float* px, const float* py, const float* pz,
                   const float* vx, const float* vy, const float* vz,
                   const float* mass, int64_t n, void* stream);
/* Same, HOST pointers (pinned memory gives full PCIe rate, pageable works);
 * copies into the context's device staging buffer, then imports.  Returns
 * after the copy has completed.                                           */
int  bh_import_soa_host(bh_ctx* ctx,
                        const float* px, const float* py, const float* pz,
                        const float* vx, const float* vy, const float* vz,
                        const float* mass, int64_t n);

/* ≙ simulationStep() called nsteps times (nbody_v5_bench.cu:255-283, 357).
 * Asynchronous on `stream`; no host synchronisation inside.               */
int  bh_step(bh_ctx* ctx, int nsteps, void* stream);

/* Read the state back in the reference layout and ORIGINAL body order
 * (what display() reads after the step, nbody_v5.cu:335).  Any pointer may
 * be NULL to skip that array.  ax/ay/az are the accelerations of the most
 * recent step (d_accX.., nbody_v5_bench.cu:222-224).  DEVICE pointers.     */
int  bh_export_soa(bh_ctx* ctx,
                   float* px, float* py, float* pz,
                   float* vx, float* vy, float* vz,
                   float* ax, float* ay, float* az, void* stream);
/* Same, HOST pointers; synchronises before returning.                     */
int  bh_export_soa_host(bh_ctx* ctx,
                        float* px, float* py, float* pz,
                        float* vx, float* vy, float* vz,
                        float* ax, float* ay, float* az);

/* End-to-end convenience used by bench.py's `e2e` leg: host SoA in, nsteps
 * steps, host SoA out (positions+velocities), all copies inside the call
 * (uploads in order of need, the step in three parts behind them, the
 * export as a gather whose position half downloads first).  Calling it
 * step after step with the arrays it returned is the intended use: the
 * context then keeps its record of which traversal chunks were expensive
 * (a scheduling hint only; dropped when n changes — results never depend
 * on it).  Equivalent to bh_import_soa_host + bh_step + bh_export_soa_host. */
int  bh_step_host(bh_ctx* ctx,
                  float* px, float* py, float* pz,
                  float* vx, float* vy, float* vz,
                  const float* mass, int64_t n, int nsteps);

/* Per-phase wall time of the LAST bh_step call when BH_FLAG_PHASE_TIMER is
 * set — the phases README.md:56-60 promises and the bench never prints.   */
enum {
    BH_PHASE_KEYS = 0,   /* bounds + Morton keys   (bench:259-260)          */
    BH_PHASE_SORT,       /* radix sort + reorder   (bench:262-264)          */
    BH_PHASE_BUILD,      /* octree emission        (bench:266-275)          */
    BH_PHASE_COM,        /* centre of mass         (bench:279-280)          */
    BH_PHASE_FORCE,      /* traversal              (bench:281)              */
    BH_PHASE_UPDATE,     /* kick-drift-clamp       (bench:282)              */
    BH_PHASE_TOTAL,
    BH_PHASE_COUNT
};
int  bh_phase_ms(bh_ctx* ctx, float out[BH_PHASE_COUNT]);
/* Change BH_FLAG_* after creation (e.g. switch the phase timer on for a few steps). */
int  bh_set_flags(bh_ctx* ctx, int flags);

/* Run ONE phase of the step on the context's current state (parity tests). */
int  bh_run_phase(bh_ctx* ctx, int phase, void* stream);

/* Debug getters/setters for the parity tests: copy an internal array to /
 * from HOST memory (synchronous).  `bytes` must match the array size.     */
enum {
    BH_DBG_BOUNDS = 0,   /* float[6]   as d_bounds  (bench:149-154)         */
    BH_DBG_KEYS,         /* u32[n]     Morton keys in CURRENT internal order;
                            after BH_PHASE_SORT they are ascending          */
    BH_DBG_PERM,         /* i32[n]     sort permutation: sorted slot -> slot
                            before the sort (d_indices, bench:62,264)       */
    BH_DBG_IDS,          /* i32[n]     internal slot -> original body id    */
    BH_DBG_POSM,         /* float4[n]  x,y,z,mass  current state            */
    BH_DBG_VEL,          /* float4[n]  vx,vy,vz,0  current state            */
    BH_DBG_ACC,          /* float4[n]  ax,ay,az,work  sorted order of the last
                            step (slot i pairs with *_SORTED slot i); work =
                            interaction-list entries of the body's 32-chunk  */
    BH_DBG_CELL_META,    /* int4[cells]   first,count,level|bucket<<8|slot<<12,parent
                            (slot = which child of its parent the cell is)   */
    BH_DBG_CELL_COM,     /* float4[cells] comx,comy,comz,mass               */
    BH_DBG_CELL_CHILD,   /* i32[cells*8]  child table                       */
    BH_DBG_POSM_SORTED,  /* float4[n]  positions the last sort/force used
                            (Morton order of THIS step, before the drift)   */
    BH_DBG_VEL_SORTED,   /* float4[n]  matching velocities                  */
    BH_DBG_IDS_SORTED,   /* i32[n]     matching original ids                */
    BH_DBG_KEYS64,       /* u64[n]     key_bits = 60 only: sorted 60-bit keys
                            (reference key << 30 | low word)                */
    BH_DBG_CELL_QUAD,    /* float[cells*8] BH_FLAG_QUADRUPOLE only: xx,xy,xz,yy,yz,zz,0,0 per cell (traceless, about
                            the centre of mass)                                                   */
    BH_DBG_COUNT
};
int  bh_debug_get(bh_ctx* ctx, int what, void* dst, size_t bytes);
int  bh_debug_set(bh_ctx* ctx, int what, const void* src, size_t bytes);

/* Scalar statistics (synchronises). */
enum {
    BH_STAT_N = 0,
    BH_STAT_CELLS,            /* cells emitted by the last build            */
    BH_STAT_ROOT,             /* id of the root cell                        */
    BH_STAT_INTERACTIONS_CELL,/* accepted (body,cell) pairs, last force     */
    BH_STAT_INTERACTIONS_BODY,/* direct   (body,body) pairs, last force     */
    BH_STAT_DEVICE_ERROR,     /* sticky device error flag (0 = none)        */
    BH_STAT_STEPS,            /* steps taken since import                   */
    BH_STAT_MAX_STACK,        /* deepest traversal stack seen, last force   */
    BH_STAT_COUNT
};
int64_t bh_stat(bh_ctx* ctx, int which);

/* child-table encoding (BH_DBG_CELL_CHILD) */
#define BH_CHILD_EMPTY 0x7F7F7F7F
#define BH_CHILD_IS_BODY(c) ((c) < 0)
#define BH_CHILD_BODY(c)    ((c) & 0x7FFFFFFF)

/* ---- multi-GPU: Morton-range slices (north_star) ------------
 * Every rank holds the full state; a rank traverses and integrates only
 * groups [first_body, first_body+count) of the freshly sorted order, then
 * the ranks all-gather their updated slices (posm, vel, ids) in place.     */
int  bh_set_slice(bh_ctx* ctx, int rank, int world);
/* One step in two halves so the all-gather can overlap compute: half 0 = bounds, keys, radix sort
 * (reads positions only); half 1 = reorder, tree, centre of mass, traversal, update (needs velocities
 * and ids as well).  bh_step_half(0) + bh_step_half(1) == bh_step(ctx, 1).                          */
int  bh_step_half(bh_ctx* ctx, int half, void* stream);
/* The same step in THREE parts, for hosts that overlap transfers with compute: part 0 = cube, keys, radix sort (reads
 * positions only); part 1 = reorder of positions + masses, tree, centre of mass, traversal (still no velocity, and no
 * id unless the state came through bh_import_state); part 2 = reorder of velocities + ids, update.
 * bh_step_part(0) + (1) + (2) == bh_step(ctx, 1) bit for bit.  bh_mg_step and bh_step_host are built on it.      */
int  bh_step_part(bh_ctx* ctx, int part, void* stream);
/* Device pointers to the CURRENT Morton-ordered state (float4 posm, float4
 * vel, int32 ids) and this rank's slice [first, first+count).              */
int  bh_state_ptrs(bh_ctx* ctx, void** posm, void** vel, void** ids,
                   int64_t* n, int64_t* slice_first, int64_t* slice_count);

/* The host driver of this mode (csrc/bh_mg.cu): one PROCESS per GPU, every process creates its context with
 * bh_create, loads the FULL state (bh_import_soa*), then
 *   rank 0:  bh_mg_unique_id(id)  ->  the 128 bytes reach the other processes by whatever the host has
 *            (a file, MPI, torch.distributed — plumbing)
 *   all:     bh_mg_create(&mg, ctx, id, rank, world, device)     NCCL communicator + bh_set_slice
 *            bh_mg_step(mg, nsteps, stream)                        ≙ the frame loop nbody_v5_bench.cu:353-367
 *            bh_mg_finish(mg, stream)                              `stream` waits for the gathers still in flight
 * bh_mg_step never blocks the host: the in-place NCCL all-gathers of the updated slices run on an internal
 * stream and the next step's keys + radix sort (positions only) start while velocities and ids are still
 * crossing NVLink.  Results are bit-identical to bh_step on one GPU.  NCCL is dlopen'ed on first use.     */
#define BH_MG_ID_BYTES 128
typedef struct bh_mg bh_mg;
int  bh_mg_unique_id(void* id128);
int  bh_mg_create(bh_mg** out, bh_ctx* ctx, const void* id128, int rank, int world, int device);
int  bh_mg_step(bh_mg* mg, int nsteps, void* stream);
int  bh_mg_finish(bh_mg* mg, void* stream);
int  bh_mg_info(bh_mg* mg, int* rank, int* world, int64_t* slice_first, int64_t* slice_count, int64_t* padded_per_rank);
void bh_mg_destroy(bh_mg* mg);

/* ---- multi-GPU: locally-essential-tree exchange (north_star) -----
 * For body counts that are not replicated on every GPU.  A rank owns the bodies of one Morton-key range;
 *   every few steps it
 *   1. agrees on the global bounding cube with its peers and fixes it (bh_set_fixed_bounds) so that every
 *      rank's Morton keys live on the same grid,
 *   2. sorts its bodies by the 30-bit key (bh_sort_coarse), takes part in the election of new key ranges
 *      (equal measured work) and hands the end runs of its sorted order (bh_sorted_ptrs) to their new owners;
 *   every step it
 *   3. builds the tree of its OWN bodies (bh_import_state, bh_run_phase KEYS..COM),
 *   4. describes its domain by the tight boxes of octree-aligned key intervals (bh_let_domain_boxes),
 *   5. walks its tree once per peer against the peer's boxes (bh_let_export): the coarsest point masses —
 *      accepted cells' {centre of mass, mass}, loose bodies, bodies of rejected buckets — that may stand in
 *      for this rank's bodies anywhere inside those boxes,
 *   6. exchanges those lists, gives what it received a small tree in a second context on the same cube, and
 *      traverses both trees with its body groups (BH_PHASE_FORCE, then bh_force_from), then BH_PHASE_UPDATE
 *      and bh_export_real.
 * nbody-barnes-hut-cuda_b200/let.py is the torch.distributed driver; DESIGN.md §6 has the measurements.   */
/* Use this cube instead of computing one from the state (NULL switches back).  b = d_bounds layout
 * (min xyz, min+size xyz; nbody_v5_bench.cu:149-154).                                                  */
int  bh_set_fixed_bounds(bh_ctx* ctx, const float b[6]);
/* min xyz / max xyz of the current state (host floats; synchronises).                                   */
int  bh_local_bounds(bh_ctx* ctx, float lohi[6]);
/* Load the internal layout directly: DEVICE float4 posm {x,y,z,m}, float4 vel, int32 ids (the step carries
 * them along unchanged).  A body with id < 0 is a GHOST: it attracts like any other body but its own
 * acceleration is not wanted — the traversal leaves it out of the group boxes and skips groups made of
 * ghosts only (imported far-field points would otherwise form huge, expensive groups).  vel.w is carried
 * along untouched by the step (let.py keeps the per-body work estimate there).                          */
int  bh_import_state(bh_ctx* ctx, const void* posm, const void* vel, const int32_t* ids, int64_t n, void* stream);
/* A domain (a Morton-key range) is not convex, so it is described by K boxes: the tight AABB (lo xyz, hi
 * xyz; lo > hi when empty) of the bodies whose key lies in [cuts[k], cuts[k+1]) for k < K.  Cutting at
 * octree-cell boundaries of the key space makes every whole-cell interval convex and disjoint from the other
 * ranks' ranges.  cuts: HOST K+1 ascending keys (the last may be 2^30); lohi: HOST K x 6; body_counts: HOST
 * K or NULL.  Needs phases KEYS+SORT of the current state; synchronises.                                  */
#define BH_LET_MAX_PEERS 64
#define BH_LET_MAX_BOXES 256
int  bh_let_domain_boxes(bh_ctx* ctx, const uint32_t* cuts, int K, float* lohi, int32_t* body_counts);
/* Host-side decisions of this mode (pure functions; every rank evaluates them on all-gathered data).
 * bh_let_domain_cuts: max_boxes+1 ascending keys cutting [k_lo, k_hi) at octree-cell boundaries — the 8..64
 *   whole cells of size 8^j inside the range, its two ragged ends cut again at 8^j/64 — padded with k_hi
 *   (empty intervals); BH_E_INVAL if max_boxes is too small (190 always suffice).
 * bh_let_elect_splitters: samples = world x nsample keys (row r: keys of rank r at equal increments of its
 *   cumulative work), work[r] = total work of rank r; edges[world+1] = key ranges of equal pooled work
 *   (edges[0] = 0, edges[world] = 2^30).                                                                    */
int  bh_let_domain_cuts(uint32_t k_lo, uint32_t k_hi, int max_boxes, uint32_t* cuts);
int  bh_let_elect_splitters(const int64_t* samples, int nsample, const double* work, int world, int64_t* edges);
/* Device pointers to THIS step's Morton-sorted arrays (u32 keys ascending, float4 posm, float4 vel, int32
 * ids) after phases KEYS+SORT — a key-range owner migrates bodies to their new owners straight from these
 * (runs of the sorted order) — and to the accelerations of the last force phase (float4 ax,ay,az,work;
 * same order).  Any pointer may be NULL.  Valid until the next import/step.                               */
int  bh_sorted_ptrs(bh_ctx* ctx, void** keys, void** posm, void** vel, void** ids, void** acc, int64_t* n);
/* boxes_lohi: HOST npeers x K x 6 (lo xyz, hi xyz; lo.x > hi.x = unused box; a peer with no used box is
 * skipped); out: DEVICE float4 [npeers * cap_per_peer]; counts: HOST npeers.  A cell is emitted as one
 * point iff it passes the acceptance test at the NEAREST of the peer's boxes.  Needs the tree of the
 * current state (phases KEYS..COM).  Synchronises.  Returns BH_E_DEVICE if an output list overflowed.  */
int  bh_let_export(bh_ctx* ctx, const float* boxes_lohi, int npeers, int K, void* out, int64_t cap_per_peer,
                   int32_t* counts, void* stream);

/* Keys + sort + reorder on the reference's 30-bit key only, whatever key_bits is: enough to elect key-range
 * splitters and to find the runs of the sorted order that migrate (bh_sorted_ptrs), at half the sorting cost
 * of a 60-bit context.  The tree phases need a full BH_PHASE_KEYS + BH_PHASE_SORT afterwards.              */
int  bh_sort_coarse(bh_ctx* ctx, void* stream);
/* Cross-tree traversal: ADD to the accelerations of ctx's last force phase (BH_PHASE_FORCE) the attraction of
 * every body of `sources`, traversing sources' tree with ctx's body groups.  Both contexts: same device, same
 * key_bits, the same fixed cube (bh_set_fixed_bounds), phases KEYS..COM current; sources needs >= 2 bodies.
 * A rank of the LET mode keeps the points imported from its peers in a small context of their own and never
 * re-sorts or rebuilds its own tree for them.  Run BH_PHASE_UPDATE on ctx afterwards.                      */
int  bh_force_from(bh_ctx* ctx, bh_ctx* sources, void* stream);
/* After a step on own bodies + ghosts: copy the bodies with id >= 0 of the current state, in their (Morton)
 * order, to the DEVICE arrays posm_out / vel_out (float4) / ids_out, with vel.w = the work of the body's
 * traversal chunk in that step (see BH_DBG_ACC).  *n_real = bodies written.  Synchronises `stream`.      */
int  bh_export_real(bh_ctx* ctx, void* posm_out, void* vel_out, int32_t* ids_out, int64_t* n_real, void* stream);

/* ---- standalone pieces -------------------------------------------------- */
/* Stable LSD radix sort of (u32 key, u32 value) pairs over bits
 * [begin_bit,end_bit) — replaces thrust::sort_by_key (bench:262-264).
 * DEVICE pointers.  tmp==NULL: writes the scratch size to *tmp_bytes.      */
int  bh_sort_pairs_u32(const uint32_t* keys_in, const uint32_t* vals_in,
                       uint32_t* keys_out, uint32_t* vals_out, int64_t n,
                       int begin_bit, int end_bit,
                       void* tmp, size_t* tmp_bytes, void* stream);

/* Double-precision O(N*k) direct sum on the GPU for a sample of k bodies of
 * the context's state (accuracy reference; north_star "direct sum").
 * sample = ORIGINAL body ids (host), acc_out = double[3*k] (host).         */
int  bh_direct_sample(bh_ctx* ctx, const int32_t* sample, int k, double* acc_out);
/* Total kinetic and (softened, pairwise) potential energy in double.       */
int  bh_energy(bh_ctx* ctx, double* kinetic, double* potential);

/* ---- callers either side of the step --------------------------- */
/* ≙ updateVisualsKernel (nbody_v5.cu:278-292): interleaved xyz positions and the
 * speed colour ramp (t = min(|v|/150, 1); rgb = 0.4+0.6t, 0.3+0.4t, 1-0.7t) for
 * body i at vbo_p[3i..], vbo_c[3i..] — ORIGINAL body order.  DEVICE pointers
 * (e.g. mapped GL buffers, nbody_v5.cu:332-334); either may be NULL.           */
int  bh_export_visuals(bh_ctx* ctx, float* vbo_p, float* vbo_c, void* stream);
/* Totals in double: out[0]=mass, out[1..3]=linear momentum, out[4..6]=angular
 * momentum about the origin.                                                 */
int  bh_momentum(bh_ctx* ctx, double out[7]);
/* Text dump in the format of the older tool's output_bh.txt:1-5
 * ("# Bodies: N, Theta: t, dt: d" header, then "x y z vx vy vz" per body in
 * ORIGINAL order, %.6f).                                                      */
int  bh_dump_text(bh_ctx* ctx, const char* path);
/* Binary checkpoint of the exact internal state (Morton-ordered posm/vel/ids,
 * step count, parameters): loading it and stepping reproduces an uninterrupted
 * run bit for bit.                                                            */
int  bh_save_checkpoint(bh_ctx* ctx, const char* path);
int  bh_load_checkpoint(bh_ctx* ctx, const char* path);

/* Initial conditions (host arrays, n floats each).
 * refdisk: main()'s generator, nbody_v5_bench.cu:294-308, glibc rand().    */
int  bh_ic_refdisk(int64_t n, unsigned seed,
                   float* px, float* py, float* pz,
                   float* vx, float* vy, float* vz, float* mass);
int  bh_ic_uniform_cube(int64_t n, uint64_t seed, float half_edge,
                        float* px, float* py, float* pz,
                        float* vx, float* vy, float* vz, float* mass);
/* two refdisk-style discs (bench:297-307 per disc) at centres -/+ (sep/2,0,0) approaching with
 * -/+ (vx,vy,0): BASELINE.json configs[4].  Counter-based RNG, multi-threaded. */
int  bh_ic_two_disks(int64_t n, uint64_t seed, float sep, float vx, float vy,
                     float* px, float* py, float* pz,
                     float* vx_out, float* vy_out, float* vz_out, float* mass);
/* bodies [first, first+n) of the same system (slot i - first): a rank generates only its own share */
int  bh_ic_two_disks_range(int64_t first, int64_t n, uint64_t seed, float sep, float vx, float vy,
                           float* px, float* py, float* pz,
                           float* vx_out, float* vy_out, float* vz_out, float* mass);
int  bh_ic_plummer(int64_t n, uint64_t seed, float scale_a, float rcut_in_a,
                   float body_mass, float G,
                   float* px, float* py, float* pz,
                   float* vx, float* vy, float* vz, float* mass);

/* Box probes used as roofline denominators by bench.py. */
int  bh_probe_fp32_tflops(int device, float* tflops);
/* same with packed fma.rn.f32x2 (two FMAs per issued instruction) */
int  bh_probe_fp32x2_tflops(int device, float* tflops);
int  bh_probe_hbm_gbs(int device, float* gbs);

#ifdef __cplusplus
}
#endif
#endif /* BH_H_ */
