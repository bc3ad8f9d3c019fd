// kernels.cuh -- device data layouts and launch wrappers of the coancestry-curve path.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/fgt_companion.h"

namespace fgt {

// One chunk of a painting after run-length encoding (reference C:514-527), 32 bytes = one
// L2 sector so that a random gather of a chunk costs exactly one sector.
struct __align__(32) Chunk {
    double pos;     // (start+end)/2 in cM
    double size;    // end-start (grid_size)
    double w;       // min(size, gridweightMAX_cutoff)
    int32_t pop;    // donor population label (0 = target population)
    int32_t hapid;  // donor haplotype id (kept for debugging / NULL path grouping)
};

// error bits raised by kernels (mapped to FGT_ERR_* on the host)
enum : int { kErrMissingDonor = 1, kErrLabelRange = 2, kErrMode3Target = 4, kErrInternal = 8, kErrBadRuns = 16 };

struct SlabDesc {   // one accumulation task column: a contiguous range of fine bins
    int32_t b0, b1; // owned fine bins [b0, b1] on the FULL grid (before subtracting grid_start)
    int32_t dlo, dhi; // coarse-bin separations (k-j) that can produce those bins (superset)
};

struct ChromPrep {  // per chromosome, per call: everything the accumulation kernel reads
    int32_t nb;             // coarse (1 cM) bins = (int)chr_length + 1   (reference C:67-68)
    int32_t W;              // window_size (reference C:117 / C:1125)
    int32_t n_phys;         // packed individuals
    const Chunk *const *chunk_ptr; // [n_phys] coarse-bin-sorted chunks of each packed individual
    const double2 *const *ps_ptr;  // [n_phys] the same chunks as (pos, size) pairs -- 16-byte gathers for phase 1
    const int32_t *const *pop_ptr; // [n_phys] ... and their population labels
    const int32_t *t;       // [n_phys][nb] chunks per coarse bin          (numchunckinbin)
    const int32_t *first;   // [n_phys][nb] exclusive prefix of t          (start_bin)
    const int32_t *yoff;    // [n_phys][nb][W]: yoff[..][d] = sum_{delta=1..d} Y(j, j+delta)
    const long long *rowoff; // [n_phys][nb+1] exclusive prefix over j of the row totals
};

struct AccumArgs {
    ChromPrep cp;
    const uint2 *draws;         // draws[p] = the two rand() outputs of sampled pair p (buffer-relative)
    const long long *pair_base; // [N] first pair (buffer-relative) of pseudo-individual n on this chromosome
    const int32_t *phys_of;     // [N] packed individual whose chunks pseudo-individual n uses
    double *tensor;             // [N][K(K+1)/2][G] upper-triangle rows, fp64
    int32_t N, K, G, grid_start, mode;
    double bw, half_bw;
    const SlabDesc *slabs;
    int32_t n_slabs;
    unsigned long long *accepted;
    int *err;
};

// ---- two-phase strict accumulation (v2) -------------------------------------------------------
// Phase 1 evaluates every sampled pair once (no ordering constraint, full occupancy) and files the
// accepted updates, in sequence order, into per-(segment, slab) record regions; phase 2 lets one warp
// per (individual, slab) fold the records of its slab in the reference's order.
struct __align__(16) UpdateRec { int32_t cell; int32_t pad; double val; };   // cell: index in the individual's tensor; pad: index in the slab tile

struct __align__(16) RegionEnt { long long base; int32_t cnt; int32_t pad; };   // one record region: first record, count

constexpr int kDrawSlack = 256;   // uint32 of slack behind the draws of a batch: phase 1 prefetches up to 127 pairs past the end
constexpr int kMaxSlots = 4;   // slabs a coarse-bin separation can reach (regions per segment)

struct EvalArgs {
    ChromPrep cp;
    const uint2 *draws;           // bulk stream (rand_impl = 0): draws[p] = the two rand() outputs of pair p, buffer-relative
    const uint32_t *hist0;        // fused generator (rand_impl = 1): the 31-word glibc history at pair 0 of `pair_base` ...
    const uint32_t *jump_tab;     // ... and the jump table [level][digit][32] (NULL selects the bulk stream)
    const long long *pair_base;   // [N] stream position (in pairs) at which pseudo-individual n starts drawing
    const long long *rec_base;    // [N] first record slot (in pairs) of n in this launch's record buffer
    const int32_t *phys_of;       // [N]
    double *tensor;               // relaxed mode only: [N][K(K+1)/2][G], updated with fp64 atomics
    unsigned long long *accepted; // relaxed mode only (the commit kernel counts otherwise)
    int32_t relaxed;
    int32_t N, K, G, grid_start, mode, n_slots;
    double bw, half_bw, cutoff;
    double inv_bw;                // RN(1/bw), or 0 when the two-FMA division below is not provably exact for this bw
    const int32_t *first_slab;    // [W] first slab reachable from separation d
    UpdateRec *recs;              // [(pairs) * n_slots]: segment (n,j,d) owns n_slots regions of Y records each
    RegionEnt *regions;           // [N*n_slabs][nb-1][nd_max] region directory, slab-major (zero = empty)
    const SlabDesc *slabs;
    int32_t n_slabs, nd_max;
    int32_t slab_width;           // bins per slab (all slabs but the last); slab_magic = ceil(2^32 / slab_width)
    uint32_t slab_magic;
    int32_t d_parts;              // warps per (individual, coarse bin): the separations are split into parts of equal pair counts
    int *err;
};

struct CommitArgs {
    ChromPrep cp;
    const long long *pair_base;
    const int32_t *phys_of;
    int32_t N, K, G, grid_start, n_slots, n_slabs;
    const SlabDesc *slabs;        // dlo/dhi = separations whose reach includes the slab
    const UpdateRec *recs;
    const RegionEnt *regions;
    int32_t nd_max, tensor_is_zero;
    int32_t buckets;              // 1: lane-queue fold (no match.any) for shared-memory tiles; 0: match.any fold
    int32_t row_blocks;           // > 1: tasks are (individual, slab, block of rows_per_block tensor rows) over ROUTED records:
    int32_t rows_per_block;       //      `regions` then holds one entry per task and `recs` the routed buffer (see k_route)
    double *tensor;
    unsigned long long *accepted;
    int *task_counter;            // zeroed before the launch: persistent CTAs claim (individual, slab) tasks from it
    const struct CommitTile *task_tab;   // optional explicit task table (NULL path): tile = rows [row0, row0+rows) x bins [b0, b0+bins) of
    int32_t n_task_tab;                  // `tensor` (individual 0), one RegionEnt per task, heavy tasks first
    int32_t max_tile_cells;              // largest tile of the table (sizes the shared memory)
    int32_t max_ctas_per_sm;             // > 0: cap of the persistent grid (leaves room for a kernel of another stream)
    int32_t consumers;                   // 1: one consumer warp per CTA (shared-memory tiles, match.any fold); 4: one warp per step of a
                                         // stage, folds serialised by named barriers (STEPWISE); otherwise two
    int32_t tags;                        // consumers == 4: tag arrays, match.any only for steps in which three or more lanes share a slot
    int32_t tag_off;                     // (set by the launcher) byte offset of the tag arrays behind the tile, 0 = none
};
struct CommitTile { int32_t row0, rows, b0, bins; };
struct RouteArgs {               // large K: split every (individual, slab) record stream by block of tensor rows (k_route)
    const UpdateRec *recs;        // phase-1 records ...
    const RegionEnt *regions;     // ... and their directory [N*n_slabs][nb-1][nd_max]
    const SlabDesc *slabs;
    int32_t N, n_slabs, nb, nd_max, row_blocks, rows_per_block;
    int32_t j_chunks;             // ceil((nb-1) / 32): routing tasks per (individual, slab)
    int32_t *counts;              // [N*n_slabs*row_blocks*j_chunks] records per (task, row block) (counting pass)
    const int32_t *offs;          // [... + 1] exclusive scan of counts (scatter pass)
    UpdateRec *routed;            // records in (individual, slab, row block) order, tile index relative to the row block
    RegionEnt *routed_dir;        // [N*n_slabs*row_blocks] one region per task
};
void launch_route(const RouteArgs &a, bool scatter, cudaStream_t st);
void launch_exclusive_scan_i32_large(const int32_t *in, int32_t *out, long long n, int32_t *scratch /*[2 * (n/1024 + 2)]*/, cudaStream_t st);
void launch_eval_pairs(const EvalArgs &a, cudaStream_t st);
void launch_commit_ordered(const CommitArgs &a, int max_slab_width, cudaStream_t st);
size_t commit_tile_bytes(int K, int slab_width);

// ---- launch wrappers (all asynchronous on `st`) ---------------------------------------------
int chunk_tiles_for(int nsites);   // 2048-label tiles per painting row (sizes the per-tile count buffer)
void launch_chunk_count(const void *labels, const int32_t *rowmap, int label_bytes, int nsites, int rows, double weight_min,
                        int32_t *counts /*[rows*chunk_tiles_for(nsites)]*/, int32_t *rowsum /*[rows]*/, cudaStream_t st);
void launch_chunk_scan(const int32_t *rowsum, int32_t *row_off /*[rows+1]*/, int rows, cudaStream_t st);
void launch_exclusive_scan_i32(const int32_t *in, int32_t *out /*[n+1]*/, int n, cudaStream_t st);
void launch_chunk_emit(const void *labels, const int32_t *rowmap, int label_bytes, int nsites, int rows, double weight_min,
                       const int32_t *counts, const int32_t *row_off, const double *cum, const double *inc, double *cstart,
                       double *cend, int32_t *chap, cudaStream_t st);
// chunks straight from run-length-encoded paintings: row r of the output takes the runs that start at runs[src_begin[r]],
// row_off[rows+1] are the output chunk offsets (= run counts prefix when weight_min < 1)
void launch_chunks_from_runs(const fgt_run_t *runs, const long long *src_begin, const int32_t *row_off, int rows, int total,
                             int nsites, const double *cum, const double *inc, double *cstart, double *cend, int32_t *chap,
                             int *err, cudaStream_t st);
void launch_chunk_seq_runs(bool emit, const fgt_run_t *runs, const long long *src_begin, const int32_t *n_runs, int rows,
                           int nsites, double weight_min, int32_t *counts, const int32_t *row_off, const double *cum,
                           const double *inc, double *cstart, double *cend, int32_t *chap, int *err, cudaStream_t st);
void launch_bin_sort(int n_phys, int rows_per_ind, const int32_t *row_off, const double *cstart, const double *cend,
                     const int32_t *chap, const int32_t *donor_label, int n_donor_labels, double cutoff, int K, int mode,
                     int nb, int W, const double *Etab, double divisor, Chunk *sorted, long long n_sorted, const Chunk **chunk_ptr,
                     const double2 **ps_ptr, const int32_t **pop_ptr, int32_t *t,
                     int32_t *first, int32_t *yoff, long long *rowoff, int32_t *cntmat, int32_t *runfirst, int *err,
                     long long *totals, cudaStream_t st);
void launch_rand_fill(const uint32_t *base_hist, const uint32_t *level_tab, uint32_t *out, long long n_threads,
                      cudaStream_t st, long long first_thread = 0);
void launch_accum_strict(const AccumArgs &a, cudaStream_t st);
void launch_project(const double *tensor, int N, int K, int G, int S, const double *tw, const int32_t *pair_kh,
                    double *results, cudaStream_t st);
void launch_project_tensor(const double *tensor, int N, int K, int G, int S, const double *tw, const int32_t *pair_kh,
                           double *results, cudaStream_t st);   // FP64 tensor cores; results to ~1e-14, not bit-exact
void launch_normalise(double *results, int N, int S, int G, int N_total, double *rsbuf, double *ratio, cudaStream_t st);
void launch_mean_accum(const double *ratio, int N, int S, int G, double *means, cudaStream_t st);
void launch_outer_weights(const double *w, int S, int K, double *out, cudaStream_t st);
void launch_expand_raw(const double *tensor, int N, int K, int G, double *out_full, cudaStream_t st);

// callers of the path, opt-in entry points
void launch_null_normalise(const double *tw, const double *nul, int S, int K, int G, double *scratch /*[S*S*G]*/, double *nullcurves,
                           double *means, cudaStream_t st);
void launch_getlhood(const double *means, int R, int T, const double *times, const double *cand, int n_cand, int nparam,
                     const double *popfracs, double likpenalty, void *fit_scratch, double *lhood, cudaStream_t st);
size_t getlhood_scratch_bytes(int R, int n_cand);
void launch_bootstrap_combine(const double *part, int Cn, int Np, size_t cells, int cumulative, const int32_t *ids, int N, double *results,
                              cudaStream_t st);

// NULL-individual path
struct NullTask { int32_t lA, lB, b_lo, b_hi; };   // one ordered fold: label pair x range of fine bins
struct NullArgs {
    const Chunk *chunks;        // per haplotype, grouped by population label (stable), see k_null_group
    const int32_t *hap_base;    // [n_haps+1]
    const int32_t *lab_off;     // [n_haps][K+1] offsets (relative to hap_base) of each label group (labels 1..K)
    int32_t n_inds, ploidy, K, G, grid_start;
    const NullTask *tasks;      // costliest first
    int32_t n_tasks, max_slab_bins;
    double bw, half_bw;
    double *out;                // [K][K][G]
    unsigned long long *accepted;
};
void launch_null_group(int n_haps, const int32_t *row_off /*[n_haps+1]*/, const double *cstart, const double *cend,
                       const int32_t *chap, const int32_t *donor_label, int n_donor_labels, double cutoff, int K,
                       Chunk *grouped, int32_t *lab_off, int *err, cudaStream_t st);
void launch_null_strict(const NullArgs &a, cudaStream_t st);

// two-phase NULL path (null_impl = 2): evaluate every chunk pair once into records, fold them with k_commit_ordered
struct NullPairPlan { int32_t tb0, w; uint32_t magic; int32_t pad; };   // per (label of A, label of B): first counter of the pair's bin
                                                                        // ranges within the row of A, bins per range, ceil(2^32 / w) (0: w = 1)
struct NullRowPlan { int32_t task0, ntb; };                             // per label of A: its first task (natural numbering), its tasks
struct NullEvalArgs {
    const Chunk *chunks;          // per haplotype, grouped by population label (k_null_group)
    const int32_t *hap_base;      // [n_haps+1]
    const int32_t *lab_off;       // [n_haps][K+2]
    const int2 *tiles;            // [n_tiles] haplotype pairs (A, B) of this batch in the reference's (n, p, g, m) order
    int32_t n_tiles;
    const int32_t *own;           // [n_own] labels of A (0-based) this engine folds (all of them unless the call is sharded)
    int32_t n_own;
    int32_t K, G, grid_start;
    double bw, half_bw, inv_bw;
    const NullPairPlan *plan;     // [K][K]
    const NullRowPlan *rows;      // [K]
    const int32_t *blk_start;     // blocks of B labels (small label groups): row la takes blk[blk_start[la] .. blk_start[la+1])
    const int2 *blk;              // (first label, label behind the last) of a block: at most 32 bin ranges; NULL = one label at a time
    int32_t n_tasks;              // tasks of this engine (natural numbering: rows of A in label order)
    int32_t *counts;              // counting pass: [n_tiles][n_tasks]
    const int32_t *offs;          // scatter pass: [n_tiles][n_tasks] first record of (tile, task): the exclusive scan of the counts in
                                  // [task][tile] order, transposed back
    UpdateRec *recs;
    unsigned long long *unit_counter;   // zeroed before every launch: persistent warps claim (haplotype pair, label of A) units from it
    int32_t window;               // windowed enumeration (one label of B at a time)
    double dwin;                  // largest distance that can land on the grid (a safe superset)
    int32_t max_ctas_per_sm;      // > 0: cap of the persistent grid
};
size_t null_eval_smem(int K, bool blocks);
void launch_sum_i32(const int32_t *in, long long n, unsigned long long *total /*zeroed*/, cudaStream_t st);
void launch_transpose_i32(const int32_t *in, long long rows, long long cols, int32_t *out, cudaStream_t st);
void launch_null_eval(const NullEvalArgs &a, bool scatter, cudaStream_t st);
void launch_null_dir(const int32_t *offs, const int32_t *perm, int n_tasks, long long T, RegionEnt *dir, cudaStream_t st);

}  // namespace fgt
