/*
 * fgt_companion.h -- C ABI of the B200-native fastGLOBETROTTER companion library
 * (fastglobetrotter_b200/fastGLOBETROTTERCompanion.so).
 *
 * Part 1 is the drop-in boundary: the four symbols fastGLOBETROTTER.R binds with
 * .C() after dyn.load("fastGLOBETROTTERCompanion.so") (reference fastGLOBETROTTER.R:154,
 * R:199-232).  Names (including the typo), argument order, pointer-only calling convention,
 * "add into the caller-zeroed output" and "printf + exit(1) on error" are those of the
 * reference C file.  No R headers are involved on either side (reference C:1-7).
 *
 * Part 2 is the packed boundary used by benchmarks, multi-GPU sharding and the parity
 * tests: the same computation on already-decoded paintings (host or device resident).
 *
 * There is no CPU implementation behind any of these entry points: a machine without a
 * usable CUDA device makes every compute entry point fail loudly (exit(1) for the .C
 * symbols, a negative return code for the fgt_* ones).
 */
#ifndef FGT_COMPANION_H
#define FGT_COMPANION_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------
 * Part 1 -- R-callable entry points (replace the same-named functions of the reference)
 * ------------------------------------------------------------------------------------ */

/* replaces fastGLOBETROTTERCompanion.c:15-23; bound at fastGLOBETROTTER.R:217.
 * tempweights[(m*S+n)*K*K + k*K + l] = weights_mat[k*S+m] * weights_mat[l*S+n]. */
void mutiply_temp_weight(double *tempweights, double *weights_mat, int *num_surrogates, int *num_donors);

/* replaces fastGLOBETROTTERCompanion.c:183-683 (modes 1 and 2); bound at fastGLOBETROTTER.R:210.
 * means (S*S*num_bins doubles, row-major [(k*S+h)*num_bins + i]) is added into. */
void data_read(int *ploidy, int *ind_id_vec, char **filenameSAMPread, char **filenameRECOMread, double *binwidth,
               int *num_bins, int *grid_start, int *num_donors, int *donor_label_vec, int *num_inds_rec,
               char **recipient_label_vec, double *weight_min, double *gridweightMAX_cutoff, int *num_surrogates,
               double *temp_weights, double *means);

/* replaces fastGLOBETROTTERCompanion.c:1211-1621 (mode 3); bound at fastGLOBETROTTER.R:202. */
void data_read_mode3(int *ploidy, int *ind_id_vec, char **filenameSAMPread, char **filenameRECOMread,
                     double *binwidth, int *num_bins, int *grid_start, int *num_donors, int *donor_label_vec,
                     int *num_inds_rec, char **recipient_label_vec, double *weight_min,
                     double *gridweightMAX_cutoff, int *num_surrogates, double *temp_weights, double *means);

/* replaces fastGLOBETROTTERCompanion.c:685-1056 (NULL-individual curves); bound at fastGLOBETROTTER.R:224.
 * chrom_ind_res (K*K*num_bins doubles, [(lA*K + lB)*num_bins + b]) is added into. */
void data_read_null(int *ploidy, char **filenameSAMPread, char **filenameRECOMread, double *binwidth, int *num_bins,
                    int *grid_start, int *num_donors, int *donor_label_vec, int *num_inds_rec,
                    char **recipient_label_vec, double *weight_min, double *gridweightMAX_cutoff,
                    double *chrom_ind_res);

/* ------------------------------------------------------------------------------------
 * Part 2 -- packed boundary
 * ------------------------------------------------------------------------------------ */

#define FGT_OK 0
#define FGT_ERR_NO_DEVICE (-1)   /* no CUDA device / driver: there is no CPU fallback            */
#define FGT_ERR_CUDA (-2)        /* a CUDA call failed (message on stderr)                        */
#define FGT_ERR_MISSING_DONOR (-3) /* a chunk's donor haplotype has label -9 (reference C:517-522) */
#define FGT_ERR_LABEL_RANGE (-4) /* donor label > num_donors (reference C:102-107)                */
#define FGT_ERR_ARG (-5)         /* malformed problem description                                 */
#define FGT_ERR_RAND (-6)        /* libc rand() state is not the TYPE_3 generator                 */
#define FGT_ERR_MODE3_TARGET (-7) /* mode 3 met a target-population chunk (undefined behaviour in
                                     the reference, C:1159-1166 index with label-1 = -1)           */

/* one run of a run-length-encoded painting: SNPs [start, next run's start) copy donor haplotype `hap`
 * (1-based id, as in the samples file).  Runs of a row are maximal (neighbours differ in `hap`), the first
 * starts at SNP 0.  This is what the reference's chunk builder derives first (C:446-455, C:487-510). */
typedef struct fgt_run_t {
    uint32_t start;
    uint32_t hap;
} fgt_run_t;

/* one chromosome of decoded paintings: EITHER run-length encoded (preferred: `runs` != NULL) OR a dense
 * label matrix (`labels`). */
typedef struct fgt_chrom_t {
    int32_t nsites;          /* SNPs = rows of the recombination map (reference C:357-364)               */
    const double *pos;       /* host, [nsites] basepair positions                                         */
    const double *rate;      /* host, [nsites] recombination rate per bp (Morgans)                        */
    const void *labels;      /* dense: [n_packed_inds*ploidy*nsamples][nsites] donor haplotype ids        */
                             /* (1-based), row (ind*ploidy + p)*nsamples + s; uint16 or int32 (problem's  */
                             /* label_bytes), host or device memory                                       */
    int32_t labels_on_device;
    int32_t runs_on_device;  /* `runs` is device memory (run_off is always host memory)                   */
    const int64_t *run_off;  /* RLE: host, [rows+1]; row r owns runs[run_off[r] .. run_off[r+1])          */
    const fgt_run_t *runs;   /* RLE: [run_off[rows]] runs, host (ideally pinned) or device memory         */
} fgt_chrom_t;

typedef struct fgt_problem_t {
    int32_t mode;            /* 1 = data_read constants, 3 = data_read_mode3 constants                    */
    int32_t ploidy, nsamples, num_chrom;
    int32_t n_packed_inds;   /* individuals present in every labels matrix                                */
    int32_t label_bytes;     /* dense label matrices: 2 or 4 (ignored for RLE chromosomes)                */
    const fgt_chrom_t *chrom;
    int32_t num_inds;        /* pseudo-individuals handled by THIS call (this rank's shard)               */
    const int32_t *ind_rows; /* [i*num_chrom + h] packed individual used for pseudo-individual i          */
    int32_t num_inds_total;  /* divisor of the mean over individuals (= num_inds unless sharded)          */
    const int64_t *pair_offset; /* optional [h*num_inds + i]: position (in sampled pairs, from the rand()
                                state at call entry) at which (h,i) starts drawing; NULL = this call owns
                                the whole stream and offsets are the running sum in (h,i) order          */
    int64_t pairs_total;     /* with pair_offset: pairs drawn by ALL shards (state advance); else ignored */
    double binwidth;
    int32_t num_bins, grid_start, num_donors;
    const int32_t *donor_label_vec;
    int32_t n_donor_labels;
    double weight_min, gridweight_max_cutoff;
    int32_t num_surrogates;
    const double *temp_weights;   /* host, laid out as data_read expects it (reference C:590)            */
} fgt_problem_t;

typedef struct fgt_stats_t {
    int64_t pairs;           /* sampled chunk pairs evaluated by this call (sum of Y) / NULL path: (i,j) pairs */
    int64_t accepted;        /* pairs that updated a cell                                                 */
    int64_t chunks;          /* chunks built                                                              */
    int32_t kernel_launches; /* kernels launched by this call                                             */
    int64_t h2d_bytes, d2h_bytes;
    float ms_chunk, ms_rand, ms_accum, ms_project, ms_total; /* CUDA-event times on the library stream  */
    int32_t accum_launches;  /* launches of the accumulation kernel included in ms_accum                  */
} fgt_stats_t;

/* sampled-pair curves (data_read / data_read_mode3 semantics) from packed paintings.
 * means: host, S*S*num_bins doubles, added into. */
int fgt_data_read_packed(const fgt_problem_t *prob, double *means, fgt_stats_t *stats);

/* Multi-GPU (individuals shard, two small collectives).  Inside ONE process: set option "devices" = n and
 * data_read / data_read_mode3 / fgt_data_read_packed spread every call over n GPUs (one host thread and engine per
 * device, NCCL bound at run time).  With ONE PROCESS PER GPU (torchrun, MPI): rank 0 obtains an id, the launcher
 * broadcasts its 128 bytes, every rank calls fgt_dist_init after choosing its device (option "device"); afterwards a
 * call of fgt_data_read_packed whose num_inds_total exceeds num_inds is this rank's shard of a collective call: ranks own
 * consecutive blocks of individuals in rank order, pair counts are all-gathered and turned into stream offsets on the
 * device, the S^2 x G partial curves are all-reduced, every rank's `means` receives the sum and every rank's generator
 * state advances by the pairs of the whole call.  Raw tensors stay bit-exact; `means` agree with the single-GPU result
 * to <= 1e-12 relative (another summation order over individuals). */
int fgt_dist_unique_id(char *out128);
int fgt_dist_init(int rank, int world, const char *id128);
void fgt_dist_finalize(void);

/* pairs each (h,i) of the problem will draw: out[h*num_inds + i].  Lets shards exchange counts
 * (one all-gather) and derive pair_offset. */
int fgt_count_pairs_packed(const fgt_problem_t *prob, int64_t *out);

/* NULL-individual curves (data_read_null semantics): uses the first min(num_inds,100) entries of
 * rec_rows (packed individual of each recipient) and the first painting sample only. */
int fgt_data_read_null_packed(const fgt_problem_t *prob, const int32_t *rec_rows, double *chrom_ind_res,
                              fgt_stats_t *stats);

/* raw per-individual co-copying tensor of the last fgt_data_read_packed / data_read call
 * (needs option "keep_raw"=1 before the call).  mode 1: [num_inds][K*K][num_bins] as
 * total_counts_mat_all stands before reference C:578; mode 3: [num_chrom*num_inds][K*K][num_bins]
 * in tabulate call order.  Returns the number of doubles (or a negative error). */
int64_t fgt_get_raw_counts(double *out, int64_t capacity);

/* options: "device" (ordinal), "strict" (1 = order-preserving fp64 fold, bit-exact; 0 = fp64
 * atomics), "keep_raw", "rand_libc" (1 = draw from / advance the process' libc rand() state like
 * the reference does; 0 = private stream seeded by fgt_srand), "slab_bins", "cache" (parsed-file cache),
 * "threads" (parser threads), "verbose", "devices" (GPUs per call, single process), "rand_impl" (1 = glibc stream generated
 * inside the pair-evaluation kernel, 0 = bulk stream in HBM), "commit_impl", "cache_mb".  Tuning / test hooks (defaults are the measured optimum): "accum_impl"
 * (2 = two-phase, 1 = single-phase), "batch_pairs", "blocks", "ranges", "split_ready", "eval_parts", "spec_rand",
 * "copy_streams", "null_slab_bins", "commit_ctas", "reuse_prep" (one-shot); "commit_consumers" (consumer warps of a commit CTA: 4 = step-wise,
 * folds handed from warp to warp through named barriers, default; 1 / 2 = one / two warps that fold whole stages) with "commit_tags"
 * (1 = tag arrays resolve steps of distinct cells and of couples without match.any, default), "null_consumers" / "null_tags" for the commit of
 * data_read_null -- every setting gives the same bits.  "null_impl": data_read_null as 2 = two phases (every
 * chunk pair evaluated once into records, ordered commit), 1 = windowed kernel (one warp per label pair and slab), 0 = two phases
 * whenever its tables allow (K <= 4096, default); "null_tasks", "null_batch", "null_overlap", "null_block", "null_window" are its
 * tuning hooks.
 * "project_impl": 0 = exact projection (bit-identical curves,
 * default), 1 = FP64 tensor cores (curves to ~1e-14).  "null_world" / "null_rank": sharded data_read_null -- this
 * process folds only the tensor rows (two phases: rows of A labels; windowed kernel: label pairs) it owns and returns zeros elsewhere (every rank is given the caller's initial
 * buffer and keeps its own cells of it); the outputs have disjoint supports, so their sum (one all-reduce) is
 * bit-identical to the single-process result. */
int fgt_set_option(const char *key, double value);
void fgt_set_option_R(char **key, double *value);     /* the same for R: .C("fgt_set_option_R", "devices", 8) */
double fgt_get_option(const char *key);

/* seed of the private glibc-compatible stream (rand_libc = 0); same sequence as srand(seed). */
void fgt_srand(unsigned seed);

/* stats of the most recent .C-style call (data_read, data_read_mode3, data_read_null) */
void fgt_last_stats(fgt_stats_t *out);

/* glibc rand() emulator self-test helpers (host): fill out[n] with the next n draws of the
 * stream (private or libc-coupled) generated ON THE DEVICE, advancing it. */
int fgt_rand_fill(int32_t *out, int64_t n);

/* chunk builder alone (device): chunks of one painting row; arrays sized >= nsites.
 * Returns the number of chunks or a negative error. */
int fgt_chunk_one(const void *labels, int label_bytes, int nsites, const double *pos, const double *rate,
                  double weight_min, double cutoff, const int32_t *donor_label_vec, int n_donor_labels,
                  double *cpos, double *csize, double *cweight, double *cend, int32_t *cpop);

/* host-only: run the file layer of the .C entry points (list files, recombination maps, multithreaded
 * gz decode of the recipients' paintings; reference C:201-275, C:341-445) and report what it decoded.
 * Fatal input errors exit(1) exactly like data_read.  Output pointers may be NULL. */
int fgt_parse_probe(const char *samples_list, const char *recom_list, int ploidy, int n_rec, char **rec_labels,
                    int32_t *num_chrom, int32_t *nsamples, int32_t *nsites, int32_t *rec_index, int32_t *label_bytes,
                    uint64_t *label_sum, void **labels_out, double *pos_out, double *rate_out);

/* host-only: dense label matrix ([rows][nsites], uint16 or int32) -> run-length-encoded rows (fgt_run_t), the
 * painting format of fgt_chrom_t.  run_off: [rows+1] (may be NULL).  Returns the number of runs; call with
 * runs == NULL first to size the buffer. */
int fgt_rle_encode(const void *labels, int label_bytes, int64_t rows, int nsites, int64_t *run_off, fgt_run_t *runs,
                   int64_t capacity);

/* host-only: convert a whole *.samples.out[.gz] text file into the self-contained binary, chunk-level (run-length
 * encoded) interchange file "*.fgtrle" (SURVEY 8f.4).  The samples list handed to data_read / data_read_mode3 /
 * data_read_null may name .fgtrle files instead of text files.  Returns the number of runs written (< 0: error). */
int64_t fgt_pack_samples(const char *samples_path, int nsites, const char *out_path);

/* host-only self-test of the jump-ahead arithmetic: the 31-word history (oldest first) of the
 * glibc stream n_draws after srand(seed). */
int fgt_rand_host_jump(unsigned seed, uint64_t n_draws, uint32_t *hist_out);

/* ------------------------------------------------------------------------------------
 * Part 3 -- callers of the path, opt-in.  fastGLOBETROTTER.R runs unchanged without them; a user who wants
 * the steps around the curves on the GPU as well switches the corresponding R lines to these.  They are
 * tolerance-level by nature: R's %*% is BLAS and lm() is Householder QR.
 * ------------------------------------------------------------------------------------ */

/* replaces fastGLOBETROTTER.R:1216-1249 (also R:1427-1457, 1773-1803, 2078-2109, 2222-2253): results = tempweights %*% null,
 * symmetrise, expectation from the margins, null curves = results / expectation, means = means / null curves.
 * temp_weights: as handed to data_read ([(k*K+l)*S*S + m*S+n]); null_mat: [K*K][G] row-major -- the chrom_ind_res buffer of
 * data_read_null (after R's removal of the target rows/columns, if any, with K reduced accordingly); means: [S*S][G]
 * row-major (the `means` buffer of data_read), divided in place (may be NULL); null_curves_out: [S*S][G] or NULL. */
int fgt_null_normalise(const double *temp_weights, const double *null_mat, int num_surrogates, int num_donors, int num_bins,
                       double *means, double *null_curves_out);
/* .C flavour: .C("null_normalise", as.double(tempweights), as.double(t(null)), S, K, G, means = as.double(t(means))) */
void null_normalise(double *temp_weights, double *null_mat, int *num_surrogates, int *num_donors, int *num_bins, double *means);

/* replaces getlhood (fastGLOBETROTTER.R:441-489) evaluated for a whole list of candidate dates at once: means [n_curves][n_times]
 * row-major (n_curves = S*S), admixtimes [n_cand][nparam] (nparam 1 or 2), popfracs [S]; lhood_out [n_cand]. */
int fgt_getlhood_grid(const double *means, int n_curves, int n_times, const double *times, const double *admixtimes, int n_cand,
                      int nparam, const double *popfracs, double likpenalty, double *lhood_out);
/* replaces getmax(..., returnall=F) (R:491-520): grid 10^seq(0,3,.25) then Nelder-Mead in sqrt(date-1); par_out [nparam] dates,
 * value_out the minimised -lhood, likpenalty_out the penalty getmax derives from the first grid point. */
int fgt_getmax(const double *means, int n_curves, int n_times, const double *times, int nparam, const double *popfracs, double *par_out,
               double *value_out, double *likpenalty_out);

/* Bootstrap replicates formed on the device from per-(individual, chromosome) partial curves (BASELINE north_star; the
 * relaxed shortcut of SURVEY 8e): after ONE plain data_read / data_read_mode3 call made with option "keep_partials" = 1
 * (every individual once, in order), each replicate's curves are recombined from the kept projections instead of re-drawing
 * the sampled pairs: ind_id_vec [n_replicates][num_inds*num_chrom] as R:2058-2059 builds it, means_out [n_replicates][S*S*G].
 * NOT bit-comparable with the reference, which draws fresh pairs for every replicate; statistically the replicates resample
 * individuals per chromosome exactly like R's. */
int fgt_bootstrap_means(const int32_t *ind_id_vec, int n_replicates, double *means_out);

const char *fgt_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FGT_COMPANION_H */
