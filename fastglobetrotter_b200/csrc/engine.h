// engine.h -- host orchestration of the device pipeline (internal C++ API behind the C ABI).
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <chrono>
#include <vector>

#include "../../include/fgt_companion.h"
#include "dist.h"
#include "glibc_rand.h"
#include "kernels.cuh"

namespace fgt {

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    void *get(size_t bytes);   // grow-only; contents undefined after growth
    void release();
    template <typename T> T *as(size_t n) { return static_cast<T *>(get(n * sizeof(T))); }
};

struct ChromState {            // per chromosome device state that must outlive the prep pass
    DevBuf sorted, chunk_base, t, first, yoff, rowoff, cum, inc, etab, chunk_ptr, ps_ptr, pop_ptr;
    DevBuf run_src, run_rowoff, run_cnt;   // RLE paintings: per row first run / per block chunk offsets / runs per row
    // host copies of what the device tables above were built from: a repeat call with the same map / the same rows
    // (the R driver calls data_read ~100 times on one data set) skips the rebuild and the upload
    std::vector<double> h_pos, h_rate;
    std::vector<int64_t> h_run_off;
    int tab_W = -1, runs_blocks = -1, runs_pm = -1, runs_lo = -1, runs_hi = -1;
    std::vector<DevBuf> sorted_blocks;   // data_read path: one sorted-chunk buffer per block of individuals
    int nb = 0, W = 0;
    std::vector<long long> totals;   // [n_phys] pairs each packed individual samples on this chromosome
    long long n_chunks = 0;
};

struct Options {
    int device = 0;
    int strict = 1;
    int keep_raw = 0;
    int rand_libc = 1;
    int slab_bins = 0;      // 0 = automatic
    int cache = 1;
    int threads = 0;        // 0 = hardware concurrency
    int verbose = 0;
    int reuse_prep = 0;
    double batch_pairs = 0; // sampled pairs per batch of individuals (0 = from free device memory)
    int split_ready = 1;    // ranges launched at once when all paintings are already prepared (two lanes overlap)
    int null_slab_bins = 0; // NULL path: bins per slab (0 = G/16)
    int ranges = 8;         // accumulation launches per chromosome when the paintings stream in from the host
    int blocks = 0;         // blocks of individuals per chromosome in the copy/chunk/accumulate pipeline (0 = automatic)
    int eval_parts = 0;     // warps per (individual, coarse bin) in phase 1; 0 = automatic
    int copy_streams = 1;   // 2 = alternate the label block copies between two streams (two DMA engines)
    int project_impl = 0;   // 0 = exact (separately rounded multiply and add, bit-identical curves); 1 = FP64 tensor cores (~1e-14)
    int null_target = 0;    // NULL task table: accepted pairs per (haplotype pair, slab) a task aims at; 0 = default
    int commit_consumers = 4; // data_read: consumer warps per commit CTA: 4 = step-wise (two sets of two warps, folds handed over through named
                            // barriers; accumulation 3.30 -> 2.64 ms with commit_tags), 1 / 2 = one / two warps that fold whole stages
    int commit_tags = 1;    // commit_consumers = 4: tag arrays resolve steps of distinct cells and of couples without match.any
    int null_tags = 0;      // the same for the NULL path's commit (null_consumers = 4); its small tiles make most steps fall back
    int commit_ctas = 0;    // > 0: cap of the commit kernel's persistent CTAs per SM (room for the next range's evaluation kernel)
    int null_impl = 0;      // NULL path: 2 = two phases (every chunk pair evaluated once into records, ordered commit), 1 = k_null_strict,
                            // 0 = two phases whenever its tables allow (K <= 4096, G < 65536)
    double null_batch = 0;  // null_impl 2: chunk-pair evaluations per batch of haplotype pairs (0 = from free device memory)
    int null_tasks = 0;     // null_impl 2: commit tasks aimed at (0 = default)
    int null_block = -1;    // null_impl 2: take the labels of B in blocks (small label groups); -1 = when a label group averages < 24 chunks
    int null_window = -1;   // null_impl 2: windowed enumeration of phase 1; -1 = when under 90 % of the chunk pairs lie inside the distance window
    int null_min_batches = 4; // null_impl 2: batches per chromosome aimed at when the work allows (so that commit and evaluation overlap)
    int null_eval_ctas = 8, null_commit_ctas = 16;   // null_impl 2: persistent CTAs per SM of the two sides while they overlap
    int null_consumers = 1; // null_impl 2: consumer warps per commit CTA (1, 2 or 4; bench chromosome 93 / 86 / 110 ms with 2 / 1 / 4)
    int null_taper = 1;     // null_impl 2: shrinking last batches (the last commit has nothing to overlap with)
    int null_overlap = 1;   // null_impl 2: fold batch b on the commit stream while batch b+1 is evaluated (two record buffers)
    int null_world = 1, null_rank = 0;   // sharded NULL run: this rank folds the label pairs it owns (sum of the ranks = exact)
    int spec_rand = 1;      // generate the first batch's draws while the chunk/bin kernels run
    int accum_impl = 2;     // 2 = two-phase (evaluate once, ordered commit); 1 = single-phase slab warps
    int rand_impl = 1;      // 1 = glibc stream generated inside the pair-evaluation kernel; 0 = bulk stream in HBM (k_rand_fill)
    int route_impl = 1;     // large K: 1 = route records to row blocks when the slab tile does not fit shared memory, 0 = fold into the
                            // global tensor instead, 2 = also route tiles that fit but leave fewer than three CTAs per SM
    int commit_impl = 0;    // commit kernel, shared-memory tiles: 0 = match.any fold (measured faster: 1.64 vs 3.4 ms on the bench
                            // workload), 1 = lane-queue fold without match.any (kept for A/B runs; bit-exact as well)
    double cache_mb = 8192; // parsed-file cache budget (host bytes), least recently used entries are dropped
    int keep_partials = 0;  // keep the projected curves of every (individual, chromosome) of a data_read call on the device, so that
                            // bootstrap replicates can be formed from them without re-sampling (fgt_bootstrap_means; identity ids only)
    int devices = 1;        // GPUs one call of data_read / data_read_mode3 / fgt_data_read_packed is spread over (single process:
                            // one host thread and one engine per device, individuals in contiguous blocks, exchange over NCCL)
};

constexpr int kMaxEngines = 16;

class Engine {
  public:
    static Engine &get(int idx = 0);   // engine 0 serves single-device calls; engines 1.. exist for the "devices" option
    Options &opt;                      // one set of options for the whole library
    int set_option(const char *key, double v);
    double get_option(const char *key);
    void srand_private(unsigned seed) { private_hist_ = seeded_history(seed); }

    int data_read_packed(const fgt_problem_t &pb, double *means, fgt_stats_t *stats, int64_t *count_only, const DistCtx *dist = nullptr);
    // the same call spread over opt.devices GPUs of this process (individuals in contiguous blocks, see dist.h)
    static int data_read_multi(const fgt_problem_t &pb, double *means, fgt_stats_t *stats);
    int data_read_null_packed(const fgt_problem_t &pb, const int32_t *rec_rows, double *out, fgt_stats_t *stats, int shard_world = 0,
                              int shard_rank = 0, NcclComm reduce_comm = nullptr);   // reduce_comm: all-reduce the partial tensors on the device
    static int data_read_null_multi(const fgt_problem_t &pb, const int32_t *rec_rows, double *out, fgt_stats_t *stats);
    // callers of the path: opt-in entry points, see include/fgt_companion.h part 3
    int null_normalise(const double *tw, const double *null_mat, int S, int K, int G, double *means, double *null_curves);
    int getlhood_grid(const double *means, int R, int T, const double *times, const double *cand, int n_cand, int nparam,
                      const double *popfracs, double likpenalty, double *lhood);
    int getmax(const double *means, int R, int T, const double *times, int nparam, const double *popfracs, double *par, double *value,
               double *likpenalty_out);
    int bootstrap_means(const int32_t *ids, int n_rep, double *means_out);
    int64_t get_raw(double *out, int64_t cap);
    static int64_t get_raw_all(double *out, int64_t cap);
    int rand_fill_host(int32_t *out, int64_t n);
    int outer_weights(double *tempweights, const double *weights_mat, int S, int K);
    int chunk_one(const void *labels, int label_bytes, int nsites, const double *pos, const double *rate,
                  double weight_min, double cutoff, const int32_t *donor_label_vec, int n_donor_labels, double *cpos,
                  double *csize, double *cweight, double *cend, int32_t *cpop);
    fgt_stats_t last_stats{};
    int prepare_device() {            // CUDA context of the configured device (or the "no device" error): for host code that pins memory
        int rc = ensure_init();
        if (rc) return rc;
        return cudaSetDevice(device_) == cudaSuccess ? FGT_OK : FGT_ERR_CUDA;
    }

    int device() const { return device_; }
    long long last_total_pairs_ = 0;   // pairs drawn by the whole call (all shards) -- what the generator state advanced by
    bool snapshot_state(History &h) { return fetch_history(h); }
    void advance_state(const History &h0, long long pairs) { store_history(apply_jump(poly_pow_E(2ULL * (uint64_t)pairs), h0)); }

  private:
    Engine();
    int idx_ = 0, device_ = 0;
    int data_read_packed_impl(const fgt_problem_t &pb, double *means, fgt_stats_t *stats, int64_t *count_only, const DistCtx *dist);
    DevBuf d_rcnt_[2], d_roff_[2], d_routed_[2], d_rdir_[2];   // large K: routed records per batch lane (k_route)
    DevBuf d_partials_;                           // keep_partials: [Cn][N][S*S][G] projected curves of the last data_read call
    int part_Cn_ = 0, part_N_ = 0, part_S_ = 0, part_G_ = 0, part_mode_ = 0;
    bool part_valid_ = false;
    DevBuf d_msg_, d_msg_all_, d_streampos_;      // sharded calls: pair counts (local / gathered), running stream position
    double *pin_part_ = nullptr;                  // pinned: the all-reduced curves of a sharded call
    size_t pin_part_cap_ = 0;
    long long *pin_msg_ = nullptr;                // pinned: [n_max + 1] counts message, then the stream position read-back
    size_t pin_msg_cap_ = 0;
    int ensure_init();
    bool fetch_history(History &h);          // current stream state (libc-coupled or private)
    void store_history(const History &h);
    int setup_chromosome(const fgt_problem_t &pb, int h, ChromState &cs, fgt_stats_t &stx);
    void vtick(const char *what) const;
    std::chrono::steady_clock::time_point t_call_start_;
    int prep_count(const fgt_problem_t &pb, int h, int p0, int p1, const void *d_lab, cudaEvent_t copy_done);
    int setup_runs(const fgt_problem_t &pb, int h, int n_blocks, int p_lo, int p_hi, ChromState &cs);
    long long prep_sig(const fgt_problem_t &pb) const;
    int prep_block(const fgt_problem_t &pb, int h, int p0, int p1, int blk, const void *d_lab, cudaEvent_t copy_done,
                   ChromState &cs, fgt_stats_t &stx);
    int prep_chromosome(const fgt_problem_t &pb, int h, const int32_t *rowmap_dev, const int32_t *rowmap_host, int rows,
                        int rows_per_ind, int n_groups, bool for_null, ChromState &cs, fgt_stats_t &stx);
    std::vector<SlabDesc> make_slabs(int G, int grid_start, double bw, int W, int N) const;
    struct SlabPlan { std::vector<SlabDesc> slabs; std::vector<int32_t> slab_of, first_slab; int n_slots = 0, slab_width = 1; };
    SlabPlan make_slab_plan(int G, int grid_start, double bw, int W, int N) const;

    bool inited_ = false;
    int init_rc_ = 0;
    cudaStream_t st_ = nullptr, st_prep_ = nullptr, st_copy_ = nullptr, st_commit_ = nullptr;
    cudaStream_t st_copy2_ = nullptr;
    cudaStream_t st_aux_ = nullptr;   // small table uploads that must not queue behind the prep stream
    History private_hist_ = seeded_history(1);
    std::vector<uint32_t> level_tab_, jump_tab_;
    DevBuf d_level_tab_, d_jump_tab_, d_err_, d_accepted_, d_recbase_;
    long long prep_sig_ = 0;          // signature of the problem the kept chunk/bin tables were built from (reuse_prep)
    void *pin_rows_ = nullptr;        // pinned staging of the per-row run tables of a chromosome
    size_t pin_rows_cap_ = 0;
    cudaEvent_t rows_ready_ = nullptr;
    // scratch shared by all chromosomes (stream ordered)
    DevBuf d_labels_, d_cum_, d_inc_, d_counts_, d_rowoff_, d_segoff_, d_cstart_, d_cend_, d_chap_, d_cntmat_, d_runfirst_,
        d_donor_, d_etab_, d_rowmap_;
    DevBuf d_draws_, d_hist_, d_tensor_, d_results_, d_tw_, d_pairkh_, d_slabs_, d_pairbase_, d_physof_, d_rs_,
        d_ratio_, d_means_, d_raw_, d_totals_, d_recs_, d_cnt_, d_draws2_, d_recs2_, d_cnt2_, d_taskctr_, d_firstslab_;
    std::vector<ChromState> chroms_;
    ChromState null_cs_;              // NULL path: chunk buffers of the chromosome being folded
    DevBuf d_ntiles_, d_nown_, d_nplan_, d_nrows_, d_nperm_, d_ntab_, d_ncnt_, d_noff_, d_ndir_, d_ncnt2_, d_noff2_, d_ndir2_, d_nblk_, d_nblks_;   // two-phase NULL path (two lanes of batches)
    std::vector<int> null_row_owner_;             // sharded two-phase NULL run: rank that folds each row of A labels
    unsigned long long *null_pin_ = nullptr;      // pinned: the record total of a batch
    cudaEvent_t null_ev_[3] = {nullptr, nullptr, nullptr};   // commit of lane 0 / lane 1 done, evaluation side ready
    int null_two_phase(const fgt_problem_t &pb, ChromState &cs, const std::vector<int32_t> &laboff, int Nc, int null_world, int null_rank,
                       double *d_out, unsigned long long *d_acc, fgt_stats_t &stx, double in_window);
    std::vector<double> raw_host_;
    int64_t raw_count_ = 0;
    int launches_ = 0;
    fgt_stats_t carry_{};
    double mem_budget_bytes_ = 0;
    long long last_shape_sig_ = -1;
    long long last_first_batch_threads_ = 0;   // generator runs the previous call's first batch used (speculation size)
    double *pin_tw_ = nullptr;        // pinned staging of a large surrogate-weight table
    size_t pin_tw_cap_ = 0;
    double *pin_tab_ = nullptr;       // pinned staging of a chromosome's cM prefix / increments / exp table
    size_t pin_tab_cap_ = 0;
    cudaEvent_t tab_ready_ = nullptr;
    int32_t *pin_small_ = nullptr;   // pinned word for the chunk-count read-back
    bool count_pending_ = false;
    int count_key_[3] = {0, 0, 0};
    void *pinned_ = nullptr;
    size_t pinned_cap_ = 0;
};

}  // namespace fgt
