// engine.cu -- host orchestration: packed paintings -> chunks -> sampling schedule -> rand() stream
// -> ordered accumulation -> projection -> normalised curves.  One CUDA stream, grow-only device
// buffers, no CPU compute path.
#include "engine.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <chrono>
#include <cstring>
#include <cstdlib>
#include <future>
#include <mutex>
#include <thread>

namespace fgt {

#define CK(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            fprintf(stderr, "fastGLOBETROTTERCompanion(B200): CUDA error %s at %s:%d: %s\n",     \
                    cudaGetErrorName(e__), __FILE__, __LINE__, cudaGetErrorString(e__));          \
            return FGT_ERR_CUDA;                                                                  \
        }                                                                                         \
    } while (0)

void *DevBuf::get(size_t bytes) {
    if (bytes == 0) bytes = 16;
    if (bytes <= cap) return p;
    if (p) cudaFree(p);
    size_t want = bytes + bytes / 8 + 256;
    if (cudaMalloc(&p, want) != cudaSuccess) {
        p = nullptr; cap = 0;
        fprintf(stderr, "fastGLOBETROTTERCompanion(B200): cudaMalloc of %zu bytes failed\n", want);
        return nullptr;
    }
    cap = want;
    return p;
}

void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
}

static Options &global_options() {
    static Options o = []() {
        Options v;
        // an unchanged R script cannot call fgt_set_option: the number of GPUs per call can come from the environment
        if (const char *e = getenv("FGT_DEVICES")) { const int n = atoi(e); if (n >= 1) v.devices = std::min(n, kMaxEngines); }
        return v;
    }();
    return o;
}

Engine::Engine() : opt(global_options()) {}

Engine &Engine::get(int idx) {
    static Engine *pool = []() {
        Engine *p = new Engine[kMaxEngines];      // never destroyed: CUDA may already be gone when static destructors run
        for (int i = 0; i < kMaxEngines; i++) p[i].idx_ = i;
        return p;
    }();
    return pool[(idx >= 0 && idx < kMaxEngines) ? idx : 0];
}

int Engine::set_option(const char *key, double v) {
    std::string k(key);
    int iv = (int)v;
    if (k == "device") {
        if (inited_ && iv != opt.device) return FGT_ERR_ARG;   // buffers live on the device chosen before the first call
        opt.device = iv;
    }
    else if (k == "strict") opt.strict = iv;
    else if (k == "keep_raw") opt.keep_raw = iv;
    else if (k == "rand_libc") opt.rand_libc = iv;
    else if (k == "slab_bins") opt.slab_bins = iv;
    else if (k == "cache") opt.cache = iv;
    else if (k == "threads") opt.threads = iv;
    else if (k == "verbose") opt.verbose = iv;
    else if (k == "reuse_prep") opt.reuse_prep = iv;
    else if (k == "accum_impl") opt.accum_impl = iv;
    else if (k == "spec_rand") opt.spec_rand = iv;
    else if (k == "null_target") opt.null_target = iv;
    else if (k == "commit_ctas") opt.commit_ctas = iv;
    else if (k == "commit_consumers") opt.commit_consumers = iv;
    else if (k == "commit_tags") opt.commit_tags = iv;
    else if (k == "null_tags") opt.null_tags = iv;
    else if (k == "null_impl") opt.null_impl = iv;
    else if (k == "null_batch") opt.null_batch = v;
    else if (k == "null_tasks") opt.null_tasks = iv;
    else if (k == "null_overlap") opt.null_overlap = iv;
    else if (k == "null_block") opt.null_block = iv;
    else if (k == "null_min_batches") opt.null_min_batches = std::max(1, iv);
    else if (k == "null_taper") opt.null_taper = iv;
    else if (k == "null_consumers") opt.null_consumers = iv;
    else if (k == "null_eval_ctas") opt.null_eval_ctas = iv;
    else if (k == "null_commit_ctas") opt.null_commit_ctas = iv;
    else if (k == "null_window") opt.null_window = iv;
    else if (k == "null_world") opt.null_world = std::max(1, iv);
    else if (k == "null_rank") opt.null_rank = std::max(0, iv);
    else if (k == "project_impl") opt.project_impl = iv;
    else if (k == "copy_streams") opt.copy_streams = iv;
    else if (k == "eval_parts") opt.eval_parts = iv;
    else if (k == "batch_pairs") opt.batch_pairs = v;
    else if (k == "blocks") opt.blocks = iv;
    else if (k == "ranges") opt.ranges = iv;
    else if (k == "null_slab_bins") opt.null_slab_bins = iv;
    else if (k == "split_ready") opt.split_ready = iv;
    else if (k == "rand_impl") opt.rand_impl = iv;
    else if (k == "commit_impl") opt.commit_impl = iv;
    else if (k == "route_impl") opt.route_impl = iv;
    else if (k == "cache_mb") opt.cache_mb = v;
    else if (k == "keep_partials") opt.keep_partials = iv;
    else if (k == "devices") opt.devices = std::max(1, std::min(iv, kMaxEngines));
    else return FGT_ERR_ARG;
    return FGT_OK;
}

double Engine::get_option(const char *key) {
    std::string k(key);
    if (k == "device") return opt.device;
    if (k == "strict") return opt.strict;
    if (k == "keep_raw") return opt.keep_raw;
    if (k == "rand_libc") return opt.rand_libc;
    if (k == "slab_bins") return opt.slab_bins;
    if (k == "cache") return opt.cache;
    if (k == "threads") return opt.threads;
    if (k == "verbose") return opt.verbose;
    if (k == "reuse_prep") return opt.reuse_prep;
    if (k == "accum_impl") return opt.accum_impl;
    if (k == "spec_rand") return opt.spec_rand;
    if (k == "commit_ctas") return opt.commit_ctas;
    if (k == "commit_consumers") return opt.commit_consumers;
    if (k == "commit_tags") return opt.commit_tags;
    if (k == "null_tags") return opt.null_tags;
    if (k == "null_impl") return opt.null_impl;
    if (k == "null_batch") return opt.null_batch;
    if (k == "null_tasks") return opt.null_tasks;
    if (k == "null_overlap") return opt.null_overlap;
    if (k == "null_block") return opt.null_block;
    if (k == "null_min_batches") return opt.null_min_batches;
    if (k == "null_taper") return opt.null_taper;
    if (k == "null_consumers") return opt.null_consumers;
    if (k == "null_eval_ctas") return opt.null_eval_ctas;
    if (k == "null_commit_ctas") return opt.null_commit_ctas;
    if (k == "null_window") return opt.null_window;
    if (k == "null_world") return opt.null_world;
    if (k == "null_rank") return opt.null_rank;
    if (k == "project_impl") return opt.project_impl;
    if (k == "copy_streams") return opt.copy_streams;
    if (k == "eval_parts") return opt.eval_parts;
    if (k == "batch_pairs") return opt.batch_pairs;
    if (k == "blocks") return opt.blocks;
    if (k == "ranges") return opt.ranges;
    if (k == "null_slab_bins") return opt.null_slab_bins;
    if (k == "split_ready") return opt.split_ready;
    if (k == "rand_impl") return opt.rand_impl;
    if (k == "commit_impl") return opt.commit_impl;
    if (k == "route_impl") return opt.route_impl;
    if (k == "cache_mb") return opt.cache_mb;
    if (k == "keep_partials") return opt.keep_partials;
    if (k == "devices") return opt.devices;
    return NAN;
}

int Engine::ensure_init() {
    if (inited_) return init_rc_;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        fprintf(stderr, "fastGLOBETROTTERCompanion(B200): no usable CUDA device (%s). This library has no CPU path.\n",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        cudaGetLastError();
        return FGT_ERR_NO_DEVICE;
    }
    if (opt.device >= n) return FGT_ERR_ARG;
    device_ = (opt.device + idx_) % n;            // engine i of a multi-device call sits on the i-th device after the configured one
    CK(cudaSetDevice(device_));
    if (!st_) CK(cudaStreamCreateWithFlags(&st_, cudaStreamNonBlocking));
    if (!st_prep_) CK(cudaStreamCreateWithFlags(&st_prep_, cudaStreamNonBlocking));
    if (!st_copy_) CK(cudaStreamCreateWithFlags(&st_copy_, cudaStreamNonBlocking));
    if (!st_copy2_) CK(cudaStreamCreateWithFlags(&st_copy2_, cudaStreamNonBlocking));
    if (!st_aux_) CK(cudaStreamCreateWithFlags(&st_aux_, cudaStreamNonBlocking));
    if (!st_commit_) {
        int lo_p = 0, hi_p = 0;
        cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p);
        CK(cudaStreamCreateWithPriority(&st_commit_, cudaStreamNonBlocking, hi_p));
    }
    if (level_tab_.empty()) level_tab_ = build_level_tables();
    uint32_t *dt = d_level_tab_.as<uint32_t>(level_tab_.size());
    if (!dt) return FGT_ERR_CUDA;
    CK(cudaMemcpy(dt, level_tab_.data(), level_tab_.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    if (jump_tab_.empty()) jump_tab_ = build_jump_table();
    uint32_t *dj = d_jump_tab_.as<uint32_t>(jump_tab_.size());
    if (!dj) return FGT_ERR_CUDA;
    CK(cudaMemcpy(dj, jump_tab_.data(), jump_tab_.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    if (!d_err_.as<int>(4) || !d_accepted_.as<unsigned long long>(2)) return FGT_ERR_CUDA;
    inited_ = true;
    init_rc_ = FGT_OK;
    return FGT_OK;
}

bool Engine::fetch_history(History &h) {
    if (opt.rand_libc) return libc_snapshot(h);
    h = private_hist_;
    return true;
}

void Engine::store_history(const History &h) {
    if (opt.rand_libc) libc_store(h);
    else private_hist_ = h;
}

static int err_to_code(int bits) {
    if (bits & kErrMissingDonor) return FGT_ERR_MISSING_DONOR;
    if (bits & kErrLabelRange) return FGT_ERR_LABEL_RANGE;
    if (bits & kErrMode3Target) return FGT_ERR_MODE3_TARGET;
    if (bits & kErrBadRuns) return FGT_ERR_ARG;
    if (bits) return FGT_ERR_CUDA;
    return FGT_OK;
}

// Slab table for the strict accumulation kernel.  A slab owns fine bins [b0,b1] of the full grid;
// an update lands in bin B only if its distance d is within half a bin of B*bw, and a pair of
// coarse bins (j,k) only produces distances in (k-j-1, k-j+1), so separations outside
// [floor((b0-.5)bw), floor((b1+.5)bw)+1] cannot reach the slab (a superset is harmless).
std::vector<SlabDesc> Engine::make_slabs(int G, int grid_start, double bw, int W, int N) const {
    int sb = opt.slab_bins;
    if (sb <= 0) {
        // trade redundancy (each slab re-evaluates ~2 cM of separations beyond its own width)
        // against the number of independent warps
        const double target = 148.0 * 40.0;
        double best = -1;
        sb = G;
        for (int cand = 4; cand <= std::max(G, 4); cand = cand + std::max(1, cand / 4)) {
            int c = std::min(cand, G);
            double slabs = std::ceil((double)G / c);
            double tasks = slabs * N;
            double redundancy = (c * bw + 2.0) / (c * bw);
            double score = std::min(tasks, target) / redundancy;
            if (score > best * 1.02) { best = score; sb = c; }
            if (c == G) break;
        }
    }
    std::vector<SlabDesc> v;
    for (int b = 0; b < G; b += sb) {
        SlabDesc s;
        s.b0 = grid_start + b;
        s.b1 = grid_start + std::min(G, b + sb) - 1;
        double dmin = (s.b0 - 0.5) * bw - 1e-6, dmax = (s.b1 + 0.5) * bw + 1e-6;
        s.dlo = std::max(1, (int)std::floor(dmin));
        s.dhi = std::min(W - 1, (int)std::floor(dmax) + 1);
        if (s.dlo <= s.dhi) v.push_back(s);
    }
    return v;
}

// fine bin of a distance, literally the reference's rule (C:144-147) -- used on the host only to
// bound which bins a coarse-bin separation can reach (the rule is monotone in the distance)
static int host_fine_bin(double d, double bw) {
    double q = d / bw;
    double fl = std::floor(q);
    double frac = (q - fl) * bw;
    int b = (int)fl;
    if (frac > bw / 2.0) b += 1;
    return b;
}

// Slab plan of the two-phase accumulation: fixed-width slabs of fine bins; for every separation d
// the contiguous range of slabs it can reach (distances in [d-1, d+1]); n_slots = widest reach.
Engine::SlabPlan Engine::make_slab_plan(int G, int grid_start, double bw, int W, int N) const {
    SlabPlan p;
    int sb = opt.slab_bins;
    if (sb <= 0) {
        sb = std::max(1, (int)std::lround(1.0 / bw));          // 1 cM of bins: a separation reaches 3 slabs
        while ((long long)((G + sb - 1) / sb) * N < 2000 && sb > 4) sb = (sb + 1) / 2;
    }
    auto plan_for = [&](int width) {
        SlabPlan q;
        q.slab_width = width;
        q.slab_of.resize(G);
        for (int b = 0; b < G; b++) q.slab_of[b] = b / width;
        int n_slabs = (G + width - 1) / width;
        q.slabs.resize(n_slabs);
        for (int s = 0; s < n_slabs; s++) {
            q.slabs[s].b0 = grid_start + s * width;
            q.slabs[s].b1 = grid_start + std::min(G, (s + 1) * width) - 1;
            q.slabs[s].dlo = W; q.slabs[s].dhi = 0;
        }
        q.first_slab.assign(W, 0);
        q.n_slots = 1;
        for (int d = 1; d < W; d++) {
            int blo = host_fine_bin((double)(d - 1), bw) - grid_start, bhi = host_fine_bin((double)(d + 1), bw) - grid_start;
            if (bhi < 0 || blo >= G) { q.first_slab[d] = 0; continue; }
            blo = std::max(blo, 0); bhi = std::min(bhi, G - 1);
            int s0 = q.slab_of[blo], s1 = q.slab_of[bhi];
            q.first_slab[d] = s0;
            q.n_slots = std::max(q.n_slots, s1 - s0 + 1);
            for (int s = s0; s <= s1; s++) { q.slabs[s].dlo = std::min(q.slabs[s].dlo, d); q.slabs[s].dhi = std::max(q.slabs[s].dhi, d); }
        }
        return q;
    };
    p = plan_for(sb);
    while (p.n_slots > kMaxSlots && sb < G) { sb *= 2; p = plan_for(sb); }   // wider slabs -> fewer regions per segment
    return p;
}

// ---------------------------------------------------------------------------------------------
// chunk + bin one chromosome.  rows = painting rows to encode (through rowmap_dev if given),
// grouped rows_per_ind at a time into n_groups "individuals" (data_read) -- or, for the NULL path,
// one row per haplotype and no binning.
// ---------------------------------------------------------------------------------------------
int Engine::prep_chromosome(const fgt_problem_t &pb, int h, const int32_t *rowmap_dev, const int32_t *rowmap_host, int rows,
                            int rows_per_ind, int n_groups, bool for_null, ChromState &cs, fgt_stats_t &stx) {
    const fgt_chrom_t &c = pb.chrom[h];
    const int L = c.nsites;
    if (L <= 0) return FGT_ERR_ARG;
    // cM prefix: sequential fp64 sum exactly as reference C:494 (and the increment re-used at C:501)
    std::vector<double> cum(L), inc(L);
    {
        double total = 0.0;
        cum[0] = 0.0; inc[0] = 0.0;
        for (int j = 1; j < L; j++) {
            double step = c.rate[j - 1] * (c.pos[j] - c.pos[j - 1]) * 100;
            total = total + step;
            cum[j] = total; inc[j] = step;
        }
    }
    double *d_cum = d_cum_.as<double>(L), *d_inc = d_inc_.as<double>(L);
    if (!d_cum || !d_inc) return FGT_ERR_CUDA;
    CK(cudaMemcpyAsync(d_cum, cum.data(), L * sizeof(double), cudaMemcpyHostToDevice, st_));
    CK(cudaMemcpyAsync(d_inc, inc.data(), L * sizeof(double), cudaMemcpyHostToDevice, st_));
    stx.h2d_bytes += 2LL * L * sizeof(double);
    int32_t *d_rowoff = d_rowoff_.as<int32_t>(rows + 1);
    int32_t *d_rowsum = d_segoff_.as<int32_t>(rows);
    if (!d_rowoff || !d_rowsum) return FGT_ERR_CUDA;
    int32_t total_chunks = 0;
    double *d_cstart = nullptr, *d_cend = nullptr;
    int32_t *d_chap = nullptr;
    if (c.runs) {
        // run-length-encoded paintings: gather the selected rows' runs (host rows are compacted into one upload)
        if (!c.run_off) return FGT_ERR_ARG;
        const long long rows_all = (long long)pb.n_packed_inds * pb.ploidy * pb.nsamples;
        std::vector<long long> src(rows);
        std::vector<int32_t> cnt(rows), roff(rows + 1);
        std::vector<fgt_run_t> stage;
        long long acc = 0;
        roff[0] = 0;
        for (int r = 0; r < rows; r++) {
            const long long rr = rowmap_host ? rowmap_host[r] : r;
            if (rr < 0 || rr >= rows_all) return FGT_ERR_ARG;
            const long long b0 = c.run_off[rr], e0 = c.run_off[rr + 1];
            if (e0 <= b0 || b0 < 0) return FGT_ERR_ARG;
            cnt[r] = (int32_t)(e0 - b0);
            if (c.runs_on_device) src[r] = b0;
            else { src[r] = (long long)stage.size(); stage.insert(stage.end(), c.runs + b0, c.runs + e0); }
            acc += e0 - b0;
            if (acc > 0x7fffffffLL) return FGT_ERR_ARG;
            roff[r + 1] = (int32_t)acc;
        }
        long long *d_src = cs.run_src.as<long long>(rows);
        int32_t *d_cnt = cs.run_cnt.as<int32_t>(rows);
        if (!d_src || !d_cnt) return FGT_ERR_CUDA;
        const fgt_run_t *d_runs = c.runs;
        if (!c.runs_on_device) {
            void *buf = d_labels_.get(stage.size() * sizeof(fgt_run_t));
            if (!buf) return FGT_ERR_CUDA;
            CK(cudaMemcpyAsync(buf, stage.data(), stage.size() * sizeof(fgt_run_t), cudaMemcpyHostToDevice, st_));
            stx.h2d_bytes += (int64_t)(stage.size() * sizeof(fgt_run_t));
            d_runs = (const fgt_run_t *)buf;
        }
        CK(cudaMemcpyAsync(d_src, src.data(), rows * sizeof(long long), cudaMemcpyHostToDevice, st_));
        CK(cudaMemcpyAsync(d_cnt, cnt.data(), rows * sizeof(int32_t), cudaMemcpyHostToDevice, st_));
        stx.h2d_bytes += (int64_t)rows * 12;
        const bool simple = pb.weight_min < 1.0;
        if (simple) {
            CK(cudaMemcpyAsync(d_rowoff, roff.data(), (rows + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st_));
            total_chunks = (int32_t)acc;
        } else {
            launch_chunk_seq_runs(false, d_runs, d_src, d_cnt, rows, L, pb.weight_min, d_rowsum, nullptr, nullptr, nullptr, nullptr,
                                  nullptr, nullptr, d_err_.as<int>(4), st_);
            launch_chunk_scan(d_rowsum, d_rowoff, rows, st_);
            launches_ += 2;
            CK(cudaMemcpyAsync(&total_chunks, d_rowoff + rows, sizeof(int32_t), cudaMemcpyDeviceToHost, st_));
            stx.d2h_bytes += 4;
        }
        CK(cudaStreamSynchronize(st_));                 // the staging vectors die at scope exit
        if (total_chunks <= 0) return FGT_ERR_ARG;
        d_cstart = d_cstart_.as<double>(total_chunks); d_cend = d_cend_.as<double>(total_chunks);
        d_chap = d_chap_.as<int32_t>(total_chunks);
        if (!d_cstart || !d_cend || !d_chap) return FGT_ERR_CUDA;
        if (simple)
            launch_chunks_from_runs(d_runs, d_src, d_rowoff, rows, total_chunks, L, d_cum, d_inc, d_cstart, d_cend, d_chap,
                                    d_err_.as<int>(4), st_);
        else
            launch_chunk_seq_runs(true, d_runs, d_src, d_cnt, rows, L, pb.weight_min, nullptr, d_rowoff, d_cum, d_inc, d_cstart,
                                  d_cend, d_chap, d_err_.as<int>(4), st_);
        launches_ += 1;
    } else {
        const void *d_lab = c.labels;
        if (!d_lab) return FGT_ERR_ARG;
        if (!c.labels_on_device) {
            size_t bytes = (size_t)pb.n_packed_inds * pb.ploidy * pb.nsamples * L * pb.label_bytes;
            void *buf = d_labels_.get(bytes);
            if (!buf) return FGT_ERR_CUDA;
            CK(cudaMemcpyAsync(buf, c.labels, bytes, cudaMemcpyHostToDevice, st_));
            stx.h2d_bytes += (int64_t)bytes;
            d_lab = buf;
        }
        int32_t *d_counts = d_counts_.as<int32_t>((size_t)rows * chunk_tiles_for(L));
        if (!d_counts) return FGT_ERR_CUDA;
        launch_chunk_count(d_lab, rowmap_dev, pb.label_bytes, L, rows, pb.weight_min, d_counts, d_rowsum, st_);
        launch_chunk_scan(d_rowsum, d_rowoff, rows, st_);
        launches_ += 2;
        CK(cudaMemcpyAsync(&total_chunks, d_rowoff + rows, sizeof(int32_t), cudaMemcpyDeviceToHost, st_));
        // the copies above read host vectors that die at scope exit: synchronise here (also yields total_chunks)
        CK(cudaStreamSynchronize(st_));
        stx.d2h_bytes += 4;
        if (total_chunks <= 0) return FGT_ERR_ARG;
        d_cstart = d_cstart_.as<double>(total_chunks); d_cend = d_cend_.as<double>(total_chunks);
        d_chap = d_chap_.as<int32_t>(total_chunks);
        if (!d_cstart || !d_cend || !d_chap) return FGT_ERR_CUDA;
        launch_chunk_emit(d_lab, rowmap_dev, pb.label_bytes, L, rows, pb.weight_min, d_counts, d_rowoff, d_cum, d_inc, d_cstart, d_cend,
                          d_chap, st_);
        launches_ += 1;
    }
    cs.n_chunks = total_chunks;
    stx.chunks += total_chunks;
    Chunk *d_sorted = cs.sorted.as<Chunk>(total_chunks);
    if (!d_sorted) return FGT_ERR_CUDA;
    const int K = pb.num_donors;
    if (for_null) {
        int32_t *d_laboff = cs.yoff.as<int32_t>((size_t)rows * (K + 2));
        int32_t *d_base = cs.chunk_base.as<int32_t>(rows + 1);
        if (!d_laboff || !d_base) return FGT_ERR_CUDA;
        CK(cudaMemcpyAsync(d_base, d_rowoff, (rows + 1) * sizeof(int32_t), cudaMemcpyDeviceToDevice, st_));
        launch_null_group(rows, d_rowoff, d_cstart, d_cend, d_chap, d_donor_.as<int32_t>(pb.n_donor_labels),
                          pb.n_donor_labels, pb.gridweight_max_cutoff, K, d_sorted, d_laboff, d_err_.as<int>(4), st_);
        launches_ += 1;
        return FGT_OK;
    }
    // coarse binning + sampling schedule
    const double chr_length = cum[L - 1];                    // every painting's last chunk ends here (C:57-65)
    const int nb = (int)chr_length / 1 + 1;                  // reference C:67-68
    const int W = (int)std::ceil(pb.binwidth * pb.num_bins) + (pb.mode == 3 ? 1 : 3);   // C:117 / C:1125
    cs.nb = nb; cs.W = W;
    std::vector<double> etab(W);
    for (int d = 0; d < W; d++) etab[d] = std::exp(-0.05 * d);   // exp(-0.05*(k-j)) with the host libm, as C:129
    double *d_etab = d_etab_.as<double>(W);
    if (!d_etab) return FGT_ERR_CUDA;
    CK(cudaMemcpyAsync(d_etab, etab.data(), W * sizeof(double), cudaMemcpyHostToDevice, st_));
    int32_t *d_t = cs.t.as<int32_t>((size_t)n_groups * nb), *d_first = cs.first.as<int32_t>((size_t)n_groups * nb);
    int32_t *d_yoff = cs.yoff.as<int32_t>((size_t)n_groups * nb * W);
    long long *d_ro = cs.rowoff.as<long long>((size_t)n_groups * (nb + 1));
    const Chunk **d_cbase = cs.chunk_ptr.as<const Chunk *>(n_groups + 1);
    size_t cm = (size_t)n_groups * rows_per_ind * nb;
    int32_t *d_cnt = d_cntmat_.as<int32_t>(cm), *d_rf = d_runfirst_.as<int32_t>(cm);
    if (!d_t || !d_first || !d_yoff || !d_ro || !d_cbase || !d_cnt || !d_rf) return FGT_ERR_CUDA;
    CK(cudaMemsetAsync(d_cnt, 0, cm * sizeof(int32_t), st_));
    launch_bin_sort(n_groups, rows_per_ind, d_rowoff, d_cstart, d_cend, d_chap, d_donor_.as<int32_t>(pb.n_donor_labels),
                    pb.n_donor_labels, pb.gridweight_max_cutoff, K, pb.mode, nb, W, d_etab, pb.mode == 3 ? 8.0 : 10.0,
                    d_sorted, total_chunks, d_cbase, nullptr, nullptr, d_t, d_first, d_yoff, d_ro, d_cnt, d_rf, d_err_.as<int>(4), nullptr, st_);
    launches_ += 1;
    cs.totals.resize(n_groups);
    CK(cudaMemcpy2DAsync(cs.totals.data(), sizeof(long long), d_ro + nb, (nb + 1) * sizeof(long long), sizeof(long long),
                         n_groups, cudaMemcpyDeviceToHost, st_));
    CK(cudaStreamSynchronize(st_));   // etab (host vector) and totals
    stx.d2h_bytes += (int64_t)n_groups * 8;
    return FGT_OK;
}

// The reference walks the samples file forwards (C:418-425, C:553-560): a non-positive step in the
// id sequence re-tabulates the individual that was read last.  Returns the packed individual used
// by each pseudo-individual of chromosome h, or false if the walk leaves the packed range.
static bool resolve_walk(const int32_t *ids, int stride, int N, int n_packed, std::vector<int32_t> &used) {
    used.resize(N);
    int next_in_file = 0;
    for (int n = 0; n < N; n++) {
        int id = ids[(size_t)n * stride];
        if (n == 0) { used[0] = id; next_in_file = id + 1; }
        else {
            int skip = id - ids[(size_t)(n - 1) * stride] - 1;
            if (skip < 0) used[n] = used[n - 1];
            else { used[n] = next_in_file + skip; next_in_file = used[n] + 1; }
        }
        if (used[n] < 0 || used[n] >= n_packed) return false;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// Chunk + bin the packed individuals [p0, p1) of chromosome h (data_read path).  Runs on the prep
// stream; waits for the block's label copy if the paintings come from the host.  Leaves the block's
// sorted chunks in its own buffer (per-individual pointers in cs.chunk_ptr) and the per-individual
// pair totals in cs.totals.
// ---------------------------------------------------------------------------------------------
int Engine::prep_block(const fgt_problem_t &pb, int h, int p0, int p1, int blk, const void *d_lab, cudaEvent_t copy_done,
                       ChromState &cs, fgt_stats_t &stx) {
    const fgt_chrom_t &c = pb.chrom[h];
    const int L = c.nsites;
    const int PM = pb.ploidy * pb.nsamples;
    const int np = p1 - p0, rows = np * PM;
    const int K = pb.num_donors;
    cudaStream_t sp = st_prep_;
    const bool use_runs = c.runs != nullptr;
    const bool simple_runs = use_runs && pb.weight_min < 1.0;     // chunk == run: offsets and totals are known on the host
    const size_t r0 = (size_t)p0 * PM;
    const int32_t *row_off_dev = nullptr;
    int32_t total_chunks = 0;
    if (!pin_small_ && cudaHostAlloc((void **)&pin_small_, 64, cudaHostAllocDefault) != cudaSuccess) return FGT_ERR_CUDA;
    if (simple_runs) {
        if (copy_done) CK(cudaStreamWaitEvent(sp, copy_done, 0));
        total_chunks = (int32_t)(c.run_off[r0 + rows] - c.run_off[r0]);
        row_off_dev = (const int32_t *)cs.run_rowoff.p + r0 + blk;
    } else {
        if (!(count_pending_ && count_key_[0] == h && count_key_[1] == p0 && count_key_[2] == p1)) {
            int rc = prep_count(pb, h, p0, p1, d_lab, copy_done);
            if (rc) return rc;
        }
        count_pending_ = false;
        vtick("  prep: waiting for the chunk count");
        CK(cudaStreamSynchronize(sp));
        vtick("  prep: chunk count known");
        total_chunks = *pin_small_;
        stx.d2h_bytes += 4;
        row_off_dev = (const int32_t *)d_rowoff_.p;
    }
    if (total_chunks <= 0) return FGT_ERR_ARG;
    cs.n_chunks += total_chunks;
    stx.chunks += total_chunks;
    double *d_cstart = d_cstart_.as<double>(total_chunks), *d_cend = d_cend_.as<double>(total_chunks);
    int32_t *d_chap = d_chap_.as<int32_t>(total_chunks);
    if ((int)cs.sorted_blocks.size() <= blk) cs.sorted_blocks.resize(blk + 1);
    // 32-byte records, then the compact (pos, size) / label copies phase 1 gathers from (k_bin_sort)
    Chunk *d_sorted = (Chunk *)cs.sorted_blocks[blk].get((size_t)total_chunks * (sizeof(Chunk) + sizeof(double2) + sizeof(int32_t)));
    if (!d_cstart || !d_cend || !d_chap || !d_sorted) return FGT_ERR_CUDA;
    if (use_runs) {
        const fgt_run_t *d_runs = (const fgt_run_t *)d_lab;
        const long long *d_src = (const long long *)cs.run_src.p + r0;
        if (simple_runs)
            launch_chunks_from_runs(d_runs, d_src, row_off_dev, rows, total_chunks, L, (const double *)cs.cum.p, (const double *)cs.inc.p,
                                    d_cstart, d_cend, d_chap, (int *)d_err_.p, sp);
        else
            launch_chunk_seq_runs(true, d_runs, d_src, (const int32_t *)cs.run_cnt.p + r0, rows, L, pb.weight_min, nullptr, row_off_dev,
                                  (const double *)cs.cum.p, (const double *)cs.inc.p, d_cstart, d_cend, d_chap, (int *)d_err_.p, sp);
    } else {
        const char *lab_blk = (const char *)d_lab + (size_t)p0 * PM * L * pb.label_bytes;
        launch_chunk_emit(lab_blk, nullptr, pb.label_bytes, L, rows, pb.weight_min, (const int32_t *)d_counts_.p, row_off_dev,
                          (const double *)cs.cum.p, (const double *)cs.inc.p, d_cstart, d_cend, d_chap, sp);
    }
    const int nb = cs.nb, W = cs.W;
    size_t cm = (size_t)np * PM * nb;
    int32_t *d_cnt = d_cntmat_.as<int32_t>(cm), *d_rf = d_runfirst_.as<int32_t>(cm);
    if (!d_cnt || !d_rf) return FGT_ERR_CUDA;
    long long *d_tot = d_totals_.as<long long>(np);
    if (!d_tot) return FGT_ERR_CUDA;
    CK(cudaMemsetAsync(d_cnt, 0, cm * sizeof(int32_t), sp));
    launch_bin_sort(np, PM, row_off_dev, d_cstart, d_cend, d_chap, (const int32_t *)d_donor_.p, pb.n_donor_labels,
                    pb.gridweight_max_cutoff, K, pb.mode, nb, W, (const double *)cs.etab.p, pb.mode == 3 ? 8.0 : 10.0, d_sorted,
                    total_chunks, (const Chunk **)cs.chunk_ptr.p + p0, (const double2 **)cs.ps_ptr.p + p0,
                    (const int32_t **)cs.pop_ptr.p + p0, (int32_t *)cs.t.p + (size_t)p0 * nb, (int32_t *)cs.first.p + (size_t)p0 * nb,
                    (int32_t *)cs.yoff.p + (size_t)p0 * nb * W, (long long *)cs.rowoff.p + (size_t)p0 * (nb + 1), d_cnt, d_rf,
                    (int *)d_err_.p, d_tot, sp);
    launches_ += 2;
    CK(cudaMemcpyAsync(cs.totals.data() + p0, d_tot, np * sizeof(long long), cudaMemcpyDeviceToHost, sp));
    if (use_runs) CK(cudaMemcpyAsync(pin_small_ + 1, d_err_.p, sizeof(int), cudaMemcpyDeviceToHost, sp));
    vtick("  prep: emit + bin sort enqueued");
    CK(cudaStreamSynchronize(sp));
    stx.d2h_bytes += (int64_t)np * 8;
    if (use_runs && (pin_small_[1] & kErrBadRuns)) return FGT_ERR_ARG;     // malformed runs: nothing is accumulated
    return FGT_OK;
}

// Per-row tables of a run-length-encoded chromosome, built in pinned memory and uploaded once: where each row's runs
// start, how many it has, and -- per block of individuals -- the chunk offsets of its rows (for weight_min < 1 a chunk
// is a run, so nothing has to be counted on the device).
int Engine::setup_runs(const fgt_problem_t &pb, int h, int n_blocks, int p_lo, int p_hi, ChromState &cs) {
    const fgt_chrom_t &c = pb.chrom[h];
    const int PM = pb.ploidy * pb.nsamples, n_phys = pb.n_packed_inds;
    const size_t rows = (size_t)n_phys * PM;
    if (!c.run_off || n_blocks <= 0 || p_lo < 0 || p_hi > n_phys || p_lo >= p_hi) return FGT_ERR_ARG;
    if (cs.runs_blocks == n_blocks && cs.runs_pm == PM && cs.runs_lo == p_lo && cs.runs_hi == p_hi && cs.h_run_off.size() == rows + 1 && cs.run_src.p &&
        std::memcmp(cs.h_run_off.data(), c.run_off, (rows + 1) * sizeof(int64_t)) == 0)
        return FGT_OK;                                       // same rows as the previous call: the device tables are still valid
    cs.runs_blocks = -1;
    const size_t need = rows * sizeof(long long) + (rows + n_blocks) * sizeof(int32_t) + rows * sizeof(int32_t);
    CK(cudaStreamSynchronize(st_aux_));                      // the previous chromosome's upload from this staging area
    if (pin_rows_cap_ < need) {
        if (pin_rows_) cudaFreeHost(pin_rows_);
        pin_rows_ = nullptr; pin_rows_cap_ = 0;
        if (cudaHostAlloc(&pin_rows_, need + need / 4 + 64, cudaHostAllocDefault) != cudaSuccess) return FGT_ERR_CUDA;
        pin_rows_cap_ = need + need / 4;
    }
    long long *src = (long long *)pin_rows_;
    int32_t *roff = (int32_t *)(src + rows);
    int32_t *cnt = roff + rows + n_blocks;
    for (size_t r = 0; r < rows; r++) {
        const long long b0 = c.run_off[r], e0 = c.run_off[r + 1];
        if (b0 < 0 || e0 <= b0 || e0 - b0 > (long long)c.nsites) return FGT_ERR_ARG;
        src[r] = b0;
        cnt[r] = (int32_t)(e0 - b0);
    }
    for (int b = 0; b < n_blocks; b++) {
        const long long pc = p_hi - p_lo;
        const size_t r0 = (size_t)(p_lo + pc * b / n_blocks) * PM, r1 = (size_t)(p_lo + pc * (b + 1) / n_blocks) * PM;
        int32_t *base = roff + r0 + b;
        long long acc = 0;
        base[0] = 0;
        for (size_t r = r0; r < r1; r++) {
            acc += cnt[r];
            if (acc > 0x7fffffffLL) return FGT_ERR_ARG;
            base[r - r0 + 1] = (int32_t)acc;
        }
    }
    long long *d_src = cs.run_src.as<long long>(rows);
    int32_t *d_roff = cs.run_rowoff.as<int32_t>(rows + n_blocks), *d_cnt = cs.run_cnt.as<int32_t>(rows);
    if (!d_src || !d_roff || !d_cnt) return FGT_ERR_CUDA;
    CK(cudaMemcpyAsync(d_src, src, rows * sizeof(long long), cudaMemcpyHostToDevice, st_aux_));
    CK(cudaMemcpyAsync(d_roff, roff, (rows + n_blocks) * sizeof(int32_t), cudaMemcpyHostToDevice, st_aux_));
    CK(cudaMemcpyAsync(d_cnt, cnt, rows * sizeof(int32_t), cudaMemcpyHostToDevice, st_aux_));
    if (!rows_ready_) CK(cudaEventCreateWithFlags(&rows_ready_, cudaEventDisableTiming));
    CK(cudaEventRecord(rows_ready_, st_aux_));
    CK(cudaStreamWaitEvent(st_prep_, rows_ready_, 0));
    cs.h_run_off.assign(c.run_off, c.run_off + rows + 1);
    cs.runs_blocks = n_blocks; cs.runs_pm = PM; cs.runs_lo = p_lo; cs.runs_hi = p_hi;
    return FGT_OK;
}

void Engine::vtick(const char *what) const {
    if (opt.verbose)
        fprintf(stderr, "[fgt] %8.3f ms  %s\n",
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call_start_).count(), what);
}

// First step of prep_block, separately launchable: count the chunks of the block's rows, scan, and start the
// read-back of the total (pinned), so that host-side work can run while the labels stream through the count kernel.
int Engine::prep_count(const fgt_problem_t &pb, int h, int p0, int p1, const void *d_lab, cudaEvent_t copy_done) {
    const fgt_chrom_t &c = pb.chrom[h];
    const int L = c.nsites;
    const int PM = pb.ploidy * pb.nsamples;
    const int rows = (p1 - p0) * PM;
    cudaStream_t sp = st_prep_;
    if (!pin_small_ && cudaHostAlloc((void **)&pin_small_, 64, cudaHostAllocDefault) != cudaSuccess) return FGT_ERR_CUDA;
    if (copy_done) CK(cudaStreamWaitEvent(sp, copy_done, 0));
    int32_t *d_rowoff = d_rowoff_.as<int32_t>(rows + 1);
    int32_t *d_rowsum = d_segoff_.as<int32_t>(rows);
    if (!d_rowoff || !d_rowsum) return FGT_ERR_CUDA;
    if (c.runs) {
        // general weight_min rule on run-length-encoded rows (needs the per-row tables of setup_runs)
        const size_t r0 = (size_t)p0 * PM;
        launch_chunk_seq_runs(false, (const fgt_run_t *)d_lab, (const long long *)chroms_[h].run_src.p + r0,
                              (const int32_t *)chroms_[h].run_cnt.p + r0, rows, L, pb.weight_min, d_rowsum, nullptr, nullptr, nullptr,
                              nullptr, nullptr, nullptr, (int *)d_err_.p, sp);
    } else {
        const char *lab_blk = (const char *)d_lab + (size_t)p0 * PM * L * pb.label_bytes;
        int32_t *d_counts = d_counts_.as<int32_t>((size_t)rows * chunk_tiles_for(L));
        if (!d_counts) return FGT_ERR_CUDA;
        launch_chunk_count(lab_blk, nullptr, pb.label_bytes, L, rows, pb.weight_min, d_counts, d_rowsum, sp);
    }
    launch_chunk_scan(d_rowsum, d_rowoff, rows, sp);
    launches_ += 2;
    CK(cudaMemcpyAsync(pin_small_, d_rowoff + rows, sizeof(int32_t), cudaMemcpyDeviceToHost, sp));
    count_pending_ = true;
    count_key_[0] = h; count_key_[1] = p0; count_key_[2] = p1;
    return FGT_OK;
}

// per-chromosome tables that do not depend on the paintings: cM prefix, exp table, schedule buffers
int Engine::setup_chromosome(const fgt_problem_t &pb, int h, ChromState &cs, fgt_stats_t &stx) {
    const fgt_chrom_t &c = pb.chrom[h];
    const int L = c.nsites, n_phys = pb.n_packed_inds;
    if (L <= 0) return FGT_ERR_ARG;
    const int W = (int)std::ceil(pb.binwidth * pb.num_bins) + (pb.mode == 3 ? 1 : 3);   // C:117 / C:1125
    if (cs.tab_W == W && cs.nb > 0 && cs.cum.p && (int)cs.h_pos.size() == L && cs.t.cap >= (size_t)n_phys * cs.nb * sizeof(int32_t) &&
        std::memcmp(cs.h_pos.data(), c.pos, L * sizeof(double)) == 0 && std::memcmp(cs.h_rate.data(), c.rate, L * sizeof(double)) == 0) {
        // same map and window as the previous call: the device tables are still valid (only the per-call state is reset)
        cs.n_chunks = 0;
        cs.totals.assign(n_phys, 0);
        if (!cs.first.as<int32_t>((size_t)n_phys * cs.nb) || !cs.yoff.as<int32_t>((size_t)n_phys * cs.nb * W) ||
            !cs.rowoff.as<long long>((size_t)n_phys * (cs.nb + 1)) || !cs.chunk_ptr.as<const Chunk *>(n_phys) ||
            !cs.ps_ptr.as<const double2 *>(n_phys) || !cs.pop_ptr.as<const int32_t *>(n_phys))
            return FGT_ERR_CUDA;
        return FGT_OK;
    }
    cs.tab_W = -1;
    // the tables are built straight into pinned memory and uploaded without a host wait (the previous chromosome's
    // upload from the same staging area has to be over first)
    const size_t need = ((size_t)2 * L + W) * sizeof(double);
    CK(cudaStreamSynchronize(st_aux_));
    if (pin_tab_cap_ < need) {
        if (pin_tab_) cudaFreeHost(pin_tab_);
        pin_tab_ = nullptr; pin_tab_cap_ = 0;
        if (cudaHostAlloc((void **)&pin_tab_, need + need / 4, cudaHostAllocDefault) != cudaSuccess) return FGT_ERR_CUDA;
        pin_tab_cap_ = need + need / 4;
    }
    double *cum = pin_tab_, *inc = pin_tab_ + L, *etab = pin_tab_ + 2 * (size_t)L;
    double total = 0.0;                       // sequential fp64 sum exactly as reference C:494 (increment re-used at C:501)
    cum[0] = 0.0; inc[0] = 0.0;
    for (int j = 1; j < L; j++) {
        double step = c.rate[j - 1] * (c.pos[j] - c.pos[j - 1]) * 100;
        total = total + step;
        cum[j] = total; inc[j] = step;
    }
    const int nb = (int)cum[L - 1] / 1 + 1;                  // every painting's last chunk ends at cum[L-1] (C:57-68)
    cs.nb = nb; cs.W = W; cs.n_chunks = 0;
    for (int d = 0; d < W; d++) etab[d] = std::exp(-0.05 * d);   // exp(-0.05*(k-j)) with the host libm, as C:129
    double *d_cum = cs.cum.as<double>(L), *d_inc = cs.inc.as<double>(L), *d_etab = cs.etab.as<double>(W);
    if (!d_cum || !d_inc || !d_etab) return FGT_ERR_CUDA;
    if (!cs.t.as<int32_t>((size_t)n_phys * nb) || !cs.first.as<int32_t>((size_t)n_phys * nb) ||
        !cs.yoff.as<int32_t>((size_t)n_phys * nb * W) || !cs.rowoff.as<long long>((size_t)n_phys * (nb + 1)) ||
        !cs.chunk_ptr.as<const Chunk *>(n_phys) || !cs.ps_ptr.as<const double2 *>(n_phys) || !cs.pop_ptr.as<const int32_t *>(n_phys))
        return FGT_ERR_CUDA;
    cs.totals.assign(n_phys, 0);
    // on their own stream: the prep stream may already be streaming labels through the chunk count
    CK(cudaMemcpyAsync(d_cum, cum, L * sizeof(double), cudaMemcpyHostToDevice, st_aux_));
    CK(cudaMemcpyAsync(d_inc, inc, L * sizeof(double), cudaMemcpyHostToDevice, st_aux_));
    CK(cudaMemcpyAsync(d_etab, etab, W * sizeof(double), cudaMemcpyHostToDevice, st_aux_));
    if (!tab_ready_) CK(cudaEventCreateWithFlags(&tab_ready_, cudaEventDisableTiming));
    CK(cudaEventRecord(tab_ready_, st_aux_));
    CK(cudaStreamWaitEvent(st_prep_, tab_ready_, 0));        // chunk emit / bin sort of this chromosome come after this
    stx.h2d_bytes += 2LL * L * sizeof(double);
    cs.h_pos.assign(c.pos, c.pos + L);
    cs.h_rate.assign(c.rate, c.rate + L);
    cs.tab_W = W;
    return FGT_OK;
}

// everything the kept chunk/bin tables depend on (reuse_prep): paintings, map, mode constants, chunking options
long long Engine::prep_sig(const fgt_problem_t &pb) const {
    unsigned long long hsh = 1469598103934665603ULL;
    auto mix = [&](unsigned long long v) { hsh ^= v; hsh *= 1099511628211ULL; };
    auto mixd = [&](double d) { unsigned long long b; std::memcpy(&b, &d, sizeof b); mix(b); };
    mix((unsigned long long)pb.mode); mix((unsigned long long)pb.ploidy); mix((unsigned long long)pb.nsamples);
    mix((unsigned long long)pb.num_chrom); mix((unsigned long long)pb.n_packed_inds); mix((unsigned long long)pb.label_bytes);
    mix((unsigned long long)pb.num_bins); mix((unsigned long long)pb.num_donors); mix((unsigned long long)pb.n_donor_labels);
    mixd(pb.binwidth); mixd(pb.weight_min); mixd(pb.gridweight_max_cutoff);
    for (int i = 0; i < pb.n_donor_labels; i++) mix((unsigned long long)(unsigned)pb.donor_label_vec[i]);
    for (int h = 0; h < pb.num_chrom; h++) {
        const fgt_chrom_t &c = pb.chrom[h];
        mix((unsigned long long)c.nsites); mix((unsigned long long)(uintptr_t)c.labels); mix((unsigned long long)(uintptr_t)c.runs);
        mix((unsigned long long)(uintptr_t)c.run_off); mix((unsigned long long)(uintptr_t)c.pos); mix((unsigned long long)(uintptr_t)c.rate);
        if (c.nsites > 0 && c.pos && c.rate) { mixd(c.pos[0]); mixd(c.pos[c.nsites - 1]); mixd(c.rate[0]); mixd(c.rate[c.nsites / 2]); }
    }
    return (long long)(hsh | 1ULL);
}

int Engine::data_read_packed(const fgt_problem_t &pb, double *means, fgt_stats_t *stats, int64_t *count_only, const DistCtx *dist) {
    const int rc = data_read_packed_impl(pb, means, stats, count_only, dist);
    if (rc != FGT_OK && rc != FGT_ERR_NO_DEVICE) {
        // an error return may leave work of earlier blocks in flight on the library's streams: drain it before the
        // caller (or the next call) touches the buffers again
        cudaDeviceSynchronize();
        cudaGetLastError();
    }
    return rc;
}

int Engine::data_read_packed_impl(const fgt_problem_t &pb, double *means, fgt_stats_t *stats, int64_t *count_only, const DistCtx *dist) {
    int rc = ensure_init();
    if (rc) return rc;
    if ((pb.mode != 1 && pb.mode != 3) || pb.ploidy <= 0 || pb.nsamples <= 0 || pb.num_chrom <= 0 || pb.num_inds <= 0 ||
        pb.n_packed_inds <= 0 || pb.num_bins <= 0 || pb.num_donors <= 0 || pb.num_surrogates <= 0 || !(pb.binwidth > 0) || !pb.chrom)
        return FGT_ERR_ARG;
    for (int h = 0; h < pb.num_chrom; h++) {
        const fgt_chrom_t &c = pb.chrom[h];
        if (c.runs ? !c.run_off : (!c.labels || (pb.label_bytes != 2 && pb.label_bytes != 4))) return FGT_ERR_ARG;
    }
    CK(cudaSetDevice(device_));
    fgt_stats_t stx{};
    launches_ = 0;
    cudaEvent_t ev[4];
    for (auto &e : ev) CK(cudaEventCreate(&e));
    CK(cudaEventRecord(ev[0], st_));

    const auto t_start = std::chrono::steady_clock::now();
    t_call_start_ = t_start;
    auto tick = [&](const char *what) { vtick(what); };
    const int N = pb.num_inds, Cn = pb.num_chrom, K = pb.num_donors, G = pb.num_bins, S = pb.num_surrogates;
    const int PM = pb.ploidy * pb.nsamples;
    const int n_phys = pb.n_packed_inds;
    // sharded call: this engine handles `N` of the `num_inds_total` individuals; where its draws start in the shared
    // stream follows from the other shards' pair counts, exchanged and prefix-summed on the device (dist.h)
    const bool dist_on = dist && dist->world > 1 && !count_only;
    const int n_msg = dist_on ? std::max(pb.num_inds_total, N) : 0;      // individuals per shard in the count message (padded)
    if (dist_on && (pb.pair_offset || !dist->comm || !nccl().loaded)) return FGT_ERR_ARG;
    int *d_err = d_err_.as<int>(4);
    unsigned long long *d_acc = d_accepted_.as<unsigned long long>(2);
    CK(cudaMemsetAsync(d_err, 0, 4 * sizeof(int), st_));
    CK(cudaMemsetAsync(d_acc, 0, 2 * sizeof(unsigned long long), st_));
    int32_t *d_donor = d_donor_.as<int32_t>(pb.n_donor_labels);
    if (!d_donor) return FGT_ERR_CUDA;
    CK(cudaMemcpyAsync(d_donor, pb.donor_label_vec, pb.n_donor_labels * sizeof(int32_t), cudaMemcpyHostToDevice, st_));
    CK(cudaStreamSynchronize(st_));       // donor labels / cleared flags are visible to the other streams
    tick("donor labels up");

    // reuse_prep (one-shot): keep the chunk/bin tables of the previous call -- only if this call describes the very same
    // paintings and sampling schedule (otherwise fall back to a full preparation)
    const long long sig = prep_sig(pb);
    const bool reuse = opt.reuse_prep && (int)chroms_.size() == Cn && !chroms_.empty() && chroms_[0].nb > 0 &&
                       (int)chroms_[0].totals.size() == n_phys && sig == prep_sig_;
    opt.reuse_prep = 0;
    prep_sig_ = sig;
    // device-resident paintings: the first chromosome's chunk count needs nothing but the labels -- start it before
    // the per-call tables are built and uploaded
    count_pending_ = false;
    if (!reuse && !pb.chrom[0].runs && pb.chrom[0].labels && pb.chrom[0].labels_on_device && pb.chrom[0].nsites > 0) {
        const int nb0 = opt.blocks > 0 ? std::min(n_phys, opt.blocks) : 1;
        rc = prep_count(pb, 0, 0, (int)((long long)n_phys * 1 / nb0), pb.chrom[0].labels, nullptr);
        if (rc) return rc;
    }
    if (reuse && !count_only) {
        stx.h2d_bytes += carry_.h2d_bytes; stx.d2h_bytes += carry_.d2h_bytes; stx.chunks += carry_.chunks;
        launches_ += carry_.kernel_launches;
    }
    carry_ = fgt_stats_t{};
    if (!reuse) chroms_.resize(Cn);

    // the walk of every chromosome (which packed individual each pseudo-individual uses)
    std::vector<std::vector<int32_t>> used(Cn);
    std::vector<char> monotone(Cn, 1);
    for (int h = 0; h < Cn; h++) {
        if (!resolve_walk(pb.ind_rows + h, Cn, N, n_phys, used[h])) {
            fprintf(stderr, "fastGLOBETROTTERCompanion(B200): ind_id_vec walks outside the recipients of the samples file\n");
            return FGT_ERR_ARG;
        }
        for (int n = 1; n < N; n++) if (used[h][n] < used[h][n - 1]) monotone[h] = 0;
    }

    History h0;
    if (dist && dist->h0) std::memcpy(h0.x, dist->h0, sizeof h0.x);      // all shards start from one generator state
    else if (!count_only && !fetch_history(h0)) {
        fprintf(stderr, "fastGLOBETROTTERCompanion(B200): libc rand() is not in its default TYPE_3 state\n");
        return FGT_ERR_RAND;
    }

    tick("walk + rand state");
    // ---- tables of the accumulation that do not depend on the paintings ---------------------------
    const size_t P = (size_t)K * (K + 1) / 2;
    const size_t tens_per_ind = P * (size_t)G;
    const size_t tens_elems = (size_t)N * tens_per_ind;
    const int W0 = (int)std::ceil(pb.binwidth * pb.num_bins) + (pb.mode == 3 ? 1 : 3);
    SlabPlan plan;
    std::vector<SlabDesc> slabs;
    int impl = opt.accum_impl;
    double *d_tensor = nullptr, *d_results = nullptr, *d_tw = nullptr, *d_raw = nullptr;
    size_t tw_bytes = 0;
    bool tw_pending = false;
    std::future<void> tw_stage;
    int32_t *d_pairkh = nullptr, *d_firstslab = nullptr;
    SlabDesc *d_slabs = nullptr;
    long long *d_pairbase = nullptr;
    int32_t *d_physof = nullptr;
    uint32_t *d_hist = nullptr;
    bool tile_in_smem = false, memset_tensor = true, relaxed = false;
    int plan_max_w = 1, nd_max = 1, row_blocks = 1, rows_per_block = 0;
    const bool fused_rng = opt.rand_impl != 0;       // draws generated inside the evaluation kernel (two-phase / relaxed only)
    long long *d_recbase = nullptr;
    uint32_t *d_h0 = nullptr;
    const bool want_raw = opt.keep_raw != 0;
    raw_count_ = 0;
    if (!count_only) {
        d_tensor = d_tensor_.as<double>(tens_elems);
        d_results = d_results_.as<double>((size_t)N * S * S * G);
        d_tw = d_tw_.as<double>((size_t)K * K * S * S);
        d_pairkh = d_pairkh_.as<int32_t>(P);
        if (!d_tensor || !d_results || !d_tw || !d_pairkh) return FGT_ERR_CUDA;
        CK(cudaMemsetAsync(d_results, 0, (size_t)N * S * S * G * sizeof(double), st_));
        // The surrogate weights are only needed by the projection.  A large table (72 MB at K=150, S=20) is staged into
        // pinned memory by a helper thread and uploaded just before the first projection, instead of holding the host
        // (and with it the whole pipeline) for a pageable copy at the start of the call.
        tw_bytes = (size_t)K * K * S * S * sizeof(double);
        if (tw_bytes >= (4u << 20)) {
            if (pin_tw_cap_ < tw_bytes) {
                if (pin_tw_) cudaFreeHost(pin_tw_);
                pin_tw_ = nullptr; pin_tw_cap_ = 0;
                if (cudaHostAlloc((void **)&pin_tw_, tw_bytes, cudaHostAllocDefault) != cudaSuccess) return FGT_ERR_CUDA;
                pin_tw_cap_ = tw_bytes;
            }
            double *dst = pin_tw_;
            const double *src = pb.temp_weights;
            const size_t nbytes = tw_bytes;
            tw_stage = std::async(std::launch::async, [dst, src, nbytes]() { std::memcpy(dst, src, nbytes); });
            tw_pending = true;
        } else if (pb.chrom[0].runs ? pb.chrom[0].runs_on_device : pb.chrom[0].labels_on_device) {
            tw_pending = true;                 // small table: copied straight from the caller's buffer, also just before the projection
        } else {
            // host-resident paintings: now, before the label copies occupy the host-to-device copy engine
            CK(cudaMemcpyAsync(d_tw, pb.temp_weights, tw_bytes, cudaMemcpyHostToDevice, st_));
        }
        stx.h2d_bytes += (int64_t)K * K * S * S * 8;
        std::vector<int32_t> pairkh(P);
        size_t q = 0;
        for (int k = 0; k < K; k++)
            for (int hh = k; hh < K; hh++) pairkh[q++] = K * k + hh;
        CK(cudaMemcpyAsync(d_pairkh, pairkh.data(), P * sizeof(int32_t), cudaMemcpyHostToDevice, st_));
        if (impl != 1 && G >= 65536) impl = 1;       // phase 1 divides bin indices with a 16-bit reciprocal
        if (impl != 1) {
            plan = make_slab_plan(G, pb.grid_start, pb.binwidth, W0, N);
            slabs = plan.slabs;
            for (const SlabDesc &s : slabs)
                if (s.dhi - s.dlo + 1 > 32) impl = 1;      // the commit kernel keeps one separation per lane
            if (plan.n_slots > kMaxSlots) impl = 1;
        }
        // relaxed mode (opt-in, "strict" = 0): the evaluation kernel adds straight into the tensor with fp64 atomics
        relaxed = !opt.strict && impl != 1;
        if (impl == 1) slabs = make_slabs(G, pb.grid_start, pb.binwidth, W0, N);
        for (const SlabDesc &s : slabs) {
            plan_max_w = std::max(plan_max_w, s.b1 - s.b0 + 1);
            nd_max = std::max(nd_max, s.dhi - s.dlo + 1);
        }
        if (impl != 1 && !relaxed) tile_in_smem = commit_tile_bytes(K, plan_max_w) + 7168 <= 200 * 1024;
        // large K: the slab tile does not fit shared memory -> route every slab's records to blocks of tensor rows whose
        // tiles do (k_route), instead of folding into the global tensor.  route_impl: 0 = never (global-tensor commit),
        // 1 = when the whole tile does not fit, 2 = whenever the tile exceeds the size that keeps three CTAs on an SM.
        if (impl != 1 && !relaxed && opt.route_impl > 0) {
            const size_t want = 56 * 1024;
            const bool need = !tile_in_smem || (opt.route_impl >= 2 && commit_tile_bytes(K, plan_max_w) > want);
            if (need) {
                long long rpb = std::max<long long>(1, (long long)(want / ((size_t)plan_max_w * sizeof(double))));
                long long nblk = ((long long)P + rpb - 1) / rpb;
                if (nblk > 32) { nblk = 32; rpb = ((long long)P + 31) / 32; }
                nblk = ((long long)P + rpb - 1) / rpb;
                if (nblk > 1 && (size_t)rpb * plan_max_w * sizeof(double) + 7168 <= 200 * 1024 && (long long)P * plan_max_w < (1LL << 28)) {
                    row_blocks = (int)nblk; rows_per_block = (int)rpb; tile_in_smem = false;
                }
            }
        }
        const bool smem_commit = tile_in_smem || row_blocks > 1;
        memset_tensor = !smem_commit;             // smem tiles start from zero and overwrite every cell
        if (memset_tensor) CK(cudaMemsetAsync(d_tensor, 0, tens_elems * sizeof(double), st_));
        d_slabs = d_slabs_.as<SlabDesc>(slabs.size());
        d_pairbase = d_pairbase_.as<long long>((size_t)Cn * N);
        d_physof = d_physof_.as<int32_t>((size_t)Cn * N);
        d_hist = d_hist_.as<uint32_t>(((size_t)Cn * N + 2) * kLag + 1);
        d_recbase = d_recbase_.as<long long>((size_t)Cn * N);
        if (!d_slabs || !d_pairbase || !d_physof || !d_hist || !d_recbase) return FGT_ERR_CUDA;
        d_h0 = d_hist + ((size_t)Cn * N + 1) * kLag;       // the generator state at call entry (fused generator)
        if (dist_on) {
            if (impl == 1 || !fused_rng) {
                fprintf(stderr, "fastGLOBETROTTERCompanion(B200): sharded calls need the in-kernel generator (rand_impl=1, accum_impl=2)\n");
                return FGT_ERR_ARG;
            }
            const size_t msg = (size_t)n_msg + 1;
            if (!d_msg_.as<long long>(msg) || !d_msg_all_.as<long long>(msg * dist->world) || !d_streampos_.as<long long>(2)) return FGT_ERR_CUDA;
            if (pin_msg_cap_ < msg + 2) {
                if (pin_msg_) cudaFreeHost(pin_msg_);
                pin_msg_ = nullptr; pin_msg_cap_ = 0;
                if (cudaHostAlloc((void **)&pin_msg_, (msg + 2 + 64) * sizeof(long long), cudaHostAllocDefault) != cudaSuccess) return FGT_ERR_CUDA;
                pin_msg_cap_ = msg + 2 + 64;
            }
            CK(cudaMemsetAsync(d_streampos_.p, 0, 2 * sizeof(long long), st_));
        }
        CK(cudaMemcpyAsync(d_slabs, slabs.data(), slabs.size() * sizeof(SlabDesc), cudaMemcpyHostToDevice, st_));
        if (impl != 1) {
            d_firstslab = d_firstslab_.as<int32_t>(W0);
            if (!d_firstslab) return FGT_ERR_CUDA;
            CK(cudaMemcpyAsync(d_firstslab, plan.first_slab.data(), W0 * sizeof(int32_t), cudaMemcpyHostToDevice, st_));
        }
        if (want_raw) {
            size_t slots = (pb.mode == 1) ? (size_t)N : (size_t)N * Cn;
            d_raw = d_raw_.as<double>(slots * K * K * G);
            if (!d_raw) return FGT_ERR_CUDA;
            raw_count_ = (int64_t)slots * K * K * G;
        }
    }
    // keep_partials: only a plain run (every individual once, in order) leaves curves that bootstrap replicates can recombine
    bool keep_part = false;
    part_valid_ = false;
    if (!count_only && opt.keep_partials && !dist_on && N == n_phys) {
        keep_part = true;
        for (int h = 0; h < Cn && keep_part; h++)
            for (int n = 0; n < N; n++) if (used[h][n] != n) { keep_part = false; break; }
        if (keep_part && !d_partials_.as<double>((size_t)Cn * N * S * S * G)) return FGT_ERR_CUDA;
    }
    cudaEvent_t tables_ready = nullptr;
    if (!count_only) { cudaEventCreateWithFlags(&tables_ready, cudaEventDisableTiming); CK(cudaEventRecord(tables_ready, st_)); }
    if (tables_ready) CK(cudaStreamWaitEvent(st_commit_, tables_ready, 0));
    tick("tables uploaded");
    const int ns_rec = (impl == 1) ? 0 : plan.n_slots;
    long long budget_pairs = (long long)opt.batch_pairs;
    if (budget_pairs <= 0 && !count_only) {
        if (mem_budget_bytes_ <= 0) {                            // cudaMemGetInfo costs milliseconds: ask once
            size_t free_b = 0, total_b = 0;
            cudaMemGetInfo(&free_b, &total_b);
            free_b += d_draws_.cap + d_recs_.cap;                // our own grow-only buffers are reusable
            mem_budget_bytes_ = std::min((double)free_b * 0.6, 64e9);
        }
        const bool use_fused = fused_rng && impl != 1;
        budget_pairs = (long long)(mem_budget_bytes_ / 2 / ((use_fused ? 0.0 : 8.0) + (relaxed ? 1.0 : 16.0 * ns_rec) + (row_blocks > 1 ? 16.0 : 0.0)));   // two lanes hold a batch each
        budget_pairs = std::min<long long>(budget_pairs, 1900000000LL);      // (routed offsets are 32-bit)
    }
    tick("tables ready");

    // timing spans on the accumulation stream
    float ms_rand = 0.f, ms_accum = 0.f, ms_proj = 0.f;
    struct Span { cudaEvent_t a, b; int kind; };
    std::vector<Span> spans;
    // ---- launch the accumulation of pseudo-individuals [n0, n1) of chromosome h -------------------
    long long run = 0, my_pairs = 0;             // running stream position / pairs of this call
    size_t hist_used = 0;
    // per-range tables go through pinned staging (a pageable async copy would make the host wait for the stream)
    const size_t stage_need = (size_t)Cn * N * (2 * sizeof(long long) + sizeof(int32_t)) + ((size_t)Cn * N + 2) * kLag * sizeof(uint32_t);
    if (!count_only && pinned_cap_ < stage_need) {
        if (pinned_) cudaFreeHost(pinned_);
        pinned_ = nullptr; pinned_cap_ = 0;
        if (cudaHostAlloc(&pinned_, stage_need + 256, cudaHostAllocDefault) != cudaSuccess) return FGT_ERR_CUDA;
        pinned_cap_ = stage_need;
    }
    long long *pin_pairbase = (long long *)pinned_;
    long long *pin_recbase = pin_pairbase + (size_t)Cn * N;
    int32_t *pin_physof = (int32_t *)(pin_recbase + (size_t)Cn * N);
    uint32_t *pin_hist = (uint32_t *)(pin_physof + (size_t)Cn * N);
    std::vector<cudaEvent_t> evs;
    auto new_event = [&]() { cudaEvent_t e; cudaEventCreateWithFlags(&e, cudaEventDisableTiming); evs.push_back(e); return e; };
    auto new_timed_event = [&]() { cudaEvent_t e; cudaEventCreate(&e); evs.push_back(e); return e; };
    const bool project_per_range = !memset_tensor;   // (a global-memory tensor in mode 3 is cleared per chromosome instead)
    // Two streams: the evaluation stream (rand stream + phase 1) and the high-priority commit stream (phase 2,
    // projection -- everything that touches the tensors, so chromosomes stay ordered).  Ranges alternate between
    // two sets of record buffers, so the evaluation of range r+1 overlaps the latency-bound ordered commit of
    // range r; the commit kernel is persistent (no waiting blocks), which lets the two kernels share the SMs.
    cudaStream_t sc = st_commit_;
    auto span_begin = [&](int kind, cudaStream_t s) { Span sp; cudaEventCreate(&sp.a); cudaEventCreate(&sp.b); sp.kind = kind; cudaEventRecord(sp.a, s); spans.push_back(sp); };
    auto span_end = [&](cudaStream_t s) { cudaEventRecord(spans.back().b, s); };
    cudaEvent_t commit_done[2] = {nullptr, nullptr};
    cudaEvent_t first_eval = nullptr, last_commit = nullptr;
    int batch_seq = 0;
    int *d_taskctr = d_taskctr_.as<int>((size_t)Cn * N + 2);
    if (!count_only) {
        if (!d_taskctr) return FGT_ERR_CUDA;
        CK(cudaMemsetAsync(d_taskctr, 0, ((size_t)Cn * N + 2) * sizeof(int), st_));
    }
    // ---- speculative draws: the first batch's share of the rand() stream starts at a known position, only its
    // length depends on the chunking.  Generate as many runs as the previous call's first batch needed while the
    // chunk/bin kernels run; the batch then tops up (or ignores the surplus).
    long long spec_threads = 0, spec_A4 = -1;
    uint32_t *spec_ptr = nullptr;
    const long long shape_sig = (((((long long)N * 1000003 + pb.n_packed_inds) * 1000003 + pb.chrom[0].nsites) * 31 + pb.nsamples) * 31 +
                                 pb.ploidy) * 1009 + pb.num_bins * 4 + pb.mode;
    if (shape_sig != last_shape_sig_) last_first_batch_threads_ = 0;     // only repeat calls of one problem shape speculate
    last_shape_sig_ = shape_sig;
    if (!count_only && fused_rng && impl != 1) {
        uint32_t *pin_h0 = pin_hist + ((size_t)Cn * N + 1) * kLag;
        std::memcpy(pin_h0, h0.x, sizeof h0.x);
        CK(cudaMemcpyAsync(d_h0, pin_h0, sizeof h0.x, cudaMemcpyHostToDevice, st_));
    }
    if (!count_only && opt.spec_rand && last_first_batch_threads_ >= 256 && !(fused_rng && impl != 1)) {
        const long long lo0 = pb.pair_offset ? pb.pair_offset[0] : 0;
        spec_A4 = (2 * lo0) & ~3LL;
        spec_threads = last_first_batch_threads_ & ~255LL;
        spec_ptr = d_draws_.as<uint32_t>((size_t)spec_threads * kRun + 4 + kDrawSlack);
        if (!spec_ptr) return FGT_ERR_CUDA;
        History hs = apply_jump(poly_pow_E((uint64_t)spec_A4), h0);
        uint32_t *pin_spec = pin_hist + (size_t)Cn * N * kLag;        // the spare last slot
        std::memcpy(pin_spec, hs.x, sizeof hs.x);
        CK(cudaMemcpyAsync(d_hist + (size_t)Cn * N * kLag, pin_spec, sizeof hs.x, cudaMemcpyHostToDevice, st_));
        span_begin(4, st_);
        launch_rand_fill(d_hist + (size_t)Cn * N * kLag, (const uint32_t *)d_level_tab_.p, spec_ptr, spec_threads, st_);
        launches_ += 1;
        span_end(st_);
    }
    auto launch_range = [&](int h, int n0, int n1, cudaEvent_t ready) -> int {
        ChromState &cs = chroms_[h];
        const bool fused = fused_rng && impl != 1;
        int n = n0;
        while (n < n1) {
            const int b0 = n;
            long long lo = -1, hi = 0, acc_pairs = 0;
            std::vector<long long> abs_off, rel_off;
            while (n < n1) {
                const long long c0 = cs.totals[used[h][n]];
                if (n > b0 && acc_pairs + c0 > budget_pairs) break;
                const long long a0 = pb.pair_offset ? pb.pair_offset[(size_t)h * N + n] : run;
                run += c0; my_pairs += c0;
                abs_off.push_back(a0);
                rel_off.push_back(acc_pairs);
                lo = (lo < 0) ? a0 : std::min(lo, a0);
                hi = std::max(hi, a0 + c0);
                acc_pairs += c0;
                n++;
            }
            const int Nb = n - b0;
            const int bi = batch_seq & 1;
            DevBuf &draws_buf = bi ? d_draws2_ : d_draws_;
            DevBuf &recs_buf = bi ? d_recs2_ : d_recs_;
            DevBuf &reg_buf = bi ? d_cnt2_ : d_cnt_;
            // bulk stream: the batch's draws [A4, 2*hi) are generated into a buffer and indexed relative to it (records too);
            // fused generator: stream positions stay absolute (the kernel jumps there), records are packed per batch
            const long long A4 = fused ? 0 : ((2 * lo) & ~3LL);
            const long long nthreads = fused ? 0 : (2 * hi - A4 + kRun - 1) / kRun;
            const long long rec_pairs = fused ? acc_pairs : nthreads * (long long)kRun / 2;
            long long *hp = pin_pairbase + (size_t)h * N + b0;
            long long *hr = pin_recbase + (size_t)h * N + b0;
            int32_t *hq = pin_physof + (size_t)h * N + b0;
            for (int m = 0; m < Nb; m++) {
                hp[m] = abs_off[m] - A4 / 2;
                hr[m] = fused ? rel_off[m] : hp[m];
                hq[m] = used[h][b0 + m];
            }
            // ---- evaluation stream ----
            if (ready) CK(cudaStreamWaitEvent(st_, ready, 0));
            if (commit_done[bi]) CK(cudaStreamWaitEvent(st_, commit_done[bi], 0));     // this buffer set is free again
            uint32_t *dh = d_hist + hist_used * kLag;
            if (!fused) {
                History hh = apply_jump(poly_pow_E((uint64_t)A4), h0);
                std::memcpy(pin_hist + hist_used * kLag, hh.x, sizeof hh.x);
                CK(cudaMemcpyAsync(dh, pin_hist + hist_used * kLag, sizeof hh.x, cudaMemcpyHostToDevice, st_));
                hist_used++;
            }
            if (!dist_on)        // (sharded calls: k_dist_offsets has already written the stream positions of chromosome h)
                CK(cudaMemcpyAsync(d_pairbase + (size_t)h * N + b0, hp, Nb * sizeof(long long), cudaMemcpyHostToDevice, st_));
            CK(cudaMemcpyAsync(d_recbase + (size_t)h * N + b0, hr, Nb * sizeof(long long), cudaMemcpyHostToDevice, st_));
            CK(cudaMemcpyAsync(d_physof + (size_t)h * N + b0, hq, Nb * sizeof(int32_t), cudaMemcpyHostToDevice, st_));
            uint32_t *d_draws = nullptr;
            if (!first_eval) { first_eval = new_timed_event(); CK(cudaEventRecord(first_eval, st_)); }
            if (!fused) {
                d_draws = draws_buf.as<uint32_t>((size_t)nthreads * kRun + 4 + kDrawSlack);
                if (!d_draws) return FGT_ERR_CUDA;
                long long have = 0;                       // runs already generated speculatively for this very range
                if (batch_seq == 0 && spec_threads > 0 && A4 == spec_A4 && d_draws == spec_ptr) have = std::min(spec_threads, nthreads & ~255LL);
                if (batch_seq == 0) last_first_batch_threads_ = nthreads;
                if (have < nthreads) {
                    span_begin(0, st_);
                    launch_rand_fill(dh, (const uint32_t *)d_level_tab_.p, d_draws, nthreads, st_, have);
                    launches_ += 1;
                    span_end(st_);
                }
            }
            ChromPrep cpd{};
            cpd.nb = cs.nb; cpd.W = cs.W; cpd.n_phys = n_phys;
            cpd.chunk_ptr = (const Chunk *const *)cs.chunk_ptr.p;
            cpd.ps_ptr = (const double2 *const *)cs.ps_ptr.p; cpd.pop_ptr = (const int32_t *const *)cs.pop_ptr.p;
            cpd.t = (const int32_t *)cs.t.p; cpd.first = (const int32_t *)cs.first.p;
            cpd.yoff = (const int32_t *)cs.yoff.p; cpd.rowoff = (const long long *)cs.rowoff.p;
            const long long *pbase = d_pairbase + (size_t)h * N + b0;
            const long long *rbase = d_recbase + (size_t)h * N + b0;
            const int32_t *pphys = d_physof + (size_t)h * N + b0;
            double *tens_b = d_tensor + (size_t)b0 * tens_per_ind;
            cudaEvent_t eval_done = new_event();
            if (impl == 1) {
                CK(cudaEventRecord(eval_done, st_));
                CK(cudaStreamWaitEvent(sc, eval_done, 0));
                span_begin(1, sc);
                AccumArgs a{};
                a.cp = cpd;
                a.draws = (const uint2 *)d_draws;
                a.pair_base = pbase; a.phys_of = pphys;
                a.tensor = tens_b;
                a.N = Nb; a.K = K; a.G = G; a.grid_start = pb.grid_start; a.mode = pb.mode;
                a.bw = pb.binwidth; a.half_bw = pb.binwidth / 2.0;
                a.slabs = d_slabs; a.n_slabs = (int)slabs.size();
                a.accepted = d_acc; a.err = d_err;
                launch_accum_strict(a, sc);
                launches_ += 1;
                span_end(sc);
            } else {
                const size_t n_ent = (size_t)Nb * slabs.size() * (cs.nb - 1) * nd_max;
                RegionEnt *d_regions = nullptr;
                UpdateRec *d_recs = nullptr;
                if (!relaxed) {
                    d_regions = reg_buf.as<RegionEnt>(n_ent);
                    d_recs = recs_buf.as<UpdateRec>((size_t)(rec_pairs + 2) * plan.n_slots);
                    if (!d_regions || !d_recs) return FGT_ERR_CUDA;
                }
                EvalArgs ea{};
                ea.cp = cpd; ea.draws = (const uint2 *)d_draws;
                ea.hist0 = d_h0; ea.jump_tab = fused ? (const uint32_t *)d_jump_tab_.p : nullptr;
                ea.pair_base = pbase; ea.rec_base = rbase; ea.phys_of = pphys;
                ea.tensor = tens_b; ea.relaxed = relaxed ? 1 : 0; ea.accepted = d_acc;
                ea.N = Nb; ea.K = K; ea.G = G; ea.grid_start = pb.grid_start; ea.mode = pb.mode; ea.n_slots = plan.n_slots;
                ea.bw = pb.binwidth; ea.half_bw = pb.binwidth / 2.0; ea.cutoff = pb.gridweight_max_cutoff;
                {
                    uint64_t bits;
                    std::memcpy(&bits, &pb.binwidth, sizeof bits);
                    const bool all_ones = (bits & 0xFFFFFFFFFFFFFULL) == 0xFFFFFFFFFFFFFULL;
                    const bool normal = std::isnormal(pb.binwidth) && pb.binwidth > 1e-100 && pb.binwidth < 1e100;
                    ea.inv_bw = (!all_ones && normal) ? 1.0 / pb.binwidth : 0.0;
                }
                ea.first_slab = d_firstslab;
                ea.slab_width = plan.slab_width;
                ea.slab_magic = plan.slab_width > 1 ? (uint32_t)(((1ULL << 32) + plan.slab_width - 1) / plan.slab_width) : 0u;
                ea.recs = d_recs; ea.regions = d_regions; ea.slabs = d_slabs; ea.n_slabs = (int)slabs.size(); ea.nd_max = nd_max;
                ea.err = d_err;
                {   // enough warps to fill the GPU twice over even for a small range of individuals
                    const long long rows_jn = (long long)Nb * std::max(1, cs.nb - 1);
                    long long ds = opt.eval_parts > 0 ? opt.eval_parts : (2 * 148 * 32 + rows_jn - 1) / rows_jn;
                    ea.d_parts = (int)std::min<long long>(8, std::max<long long>(opt.eval_parts > 0 ? 1 : 4, ds));   // 4 also balances the tail of a full-size launch
                }
                if (relaxed) {
                    // the atomics land in the tensor: this launch belongs on the stream that owns the tensors
                    CK(cudaEventRecord(eval_done, st_));
                    CK(cudaStreamWaitEvent(sc, eval_done, 0));
                    span_begin(3, sc);
                    launch_eval_pairs(ea, sc);
                    launches_ += 1;
                    span_end(sc);
                } else {
                    span_begin(3, st_);
                    CK(cudaMemsetAsync(d_regions, 0, n_ent * sizeof(RegionEnt), st_));
                    launch_eval_pairs(ea, st_);
                    CommitArgs ca{};
                    ca.cp = cpd; ca.pair_base = pbase; ca.phys_of = pphys;
                    ca.N = Nb; ca.K = K; ca.G = G; ca.grid_start = pb.grid_start; ca.n_slots = plan.n_slots;
                    ca.n_slabs = (int)slabs.size();
                    ca.slabs = d_slabs; ca.recs = d_recs; ca.regions = d_regions; ca.nd_max = nd_max;
                    ca.tensor_is_zero = ((tile_in_smem || row_blocks > 1) && (pb.mode == 3 || h == 0)) ? 1 : 0;
                    if (row_blocks > 1) {
                        // large K: count, scan, stable scatter of every slab's records by block of tensor rows (evaluation stream)
                        const size_t n_rt = (size_t)Nb * slabs.size() * row_blocks;
                        const int jch = std::max(1, (cs.nb - 1 + 31) / 32);
                        const size_t n_cnt = n_rt * jch;
                        int32_t *d_rcnt = d_rcnt_[bi].as<int32_t>(n_cnt + 1 + 2 * (n_cnt / 1024 + 2)), *d_roff = d_roff_[bi].as<int32_t>(n_cnt + 1);
                        UpdateRec *d_routed = d_routed_[bi].as<UpdateRec>((size_t)rec_pairs + 2);
                        RegionEnt *d_rdir = d_rdir_[bi].as<RegionEnt>(n_rt);
                        if (!d_rcnt || !d_roff || !d_routed || !d_rdir) return FGT_ERR_CUDA;
                        RouteArgs ra{};
                        ra.recs = d_recs; ra.regions = d_regions; ra.slabs = d_slabs;
                        ra.N = Nb; ra.n_slabs = (int)slabs.size(); ra.nb = cs.nb; ra.nd_max = nd_max;
                        ra.row_blocks = row_blocks; ra.rows_per_block = rows_per_block; ra.j_chunks = jch;
                        ra.counts = d_rcnt; ra.offs = d_roff; ra.routed = d_routed; ra.routed_dir = d_rdir;
                        launch_route(ra, false, st_);
                        launch_exclusive_scan_i32_large(d_rcnt, d_roff, (long long)n_cnt, d_rcnt + n_cnt + 1, st_);
                        launch_route(ra, true, st_);
                        launches_ += 7;
                        ca.recs = d_routed; ca.regions = d_rdir;
                        ca.row_blocks = row_blocks; ca.rows_per_block = rows_per_block;
                    }
                    span_end(st_);
                    CK(cudaEventRecord(eval_done, st_));
                    CK(cudaStreamWaitEvent(sc, eval_done, 0));
                    span_begin(1, sc);
                    ca.buckets = opt.commit_impl == 1 ? 1 : 0;
                    ca.max_ctas_per_sm = opt.commit_ctas;
                    ca.consumers = opt.commit_consumers; ca.tags = opt.commit_tags;
                    ca.tensor = tens_b; ca.accepted = d_acc;
                    ca.task_counter = d_taskctr + batch_seq;
                    launch_commit_ordered(ca, plan_max_w, sc);
                    launches_ += 2;
                    span_end(sc);
                }
            }
            stx.accum_launches += 1;
            if (project_per_range && (pb.mode == 3 || h == Cn - 1) && !want_raw) {
                // these individuals' tensors are final: project them now, off the tail of the call
                span_begin(2, sc);
                if (tw_pending) {
                    const double *tw_src = pb.temp_weights;
                    if (tw_stage.valid()) { tw_stage.get(); tw_src = pin_tw_; }
                    CK(cudaMemcpyAsync(d_tw, tw_src, tw_bytes, cudaMemcpyHostToDevice, sc));
                    tw_pending = false;
                }
                (opt.project_impl == 1 ? launch_project_tensor : launch_project)(tens_b, Nb, K, G, S, d_tw, d_pairkh, d_results + (size_t)b0 * S * S * G, sc);
                launches_ += 1;
                span_end(sc);
            }
            commit_done[bi] = new_event();
            CK(cudaEventRecord(commit_done[bi], sc));
            batch_seq++;
        }
        return FGT_OK;
    };

    // ---- chromosomes: blocks of packed individuals flow  copy -> chunk/bin -> accumulate -------------
    cudaEvent_t labels_free = nullptr;            // last chunk kernel that read the shared label staging buffer
    float ms_prep_host = 0.f;
    for (int h = 0; h < Cn; h++) {
        ChromState &cs = chroms_[h];
        const fgt_chrom_t &c = pb.chrom[h];
        const bool use_runs = c.runs != nullptr;
        const bool in_hbm = use_runs ? c.runs_on_device != 0 : c.labels_on_device != 0;
        const size_t rows_all = (size_t)n_phys * PM;
        const size_t bytes_per_ind = (size_t)PM * c.nsites * pb.label_bytes;
        // only the packed individuals this call's pseudo-individuals use are copied, chunked and binned (a shard of a
        // multi-GPU call, or a bootstrap id matrix that leaves some individuals out)
        int p_lo = n_phys, p_hi = 0;
        for (int n = 0; n < N; n++) { p_lo = std::min(p_lo, used[h][n]); p_hi = std::max(p_hi, used[h][n] + 1); }
        if (reuse) { p_lo = 0; p_hi = n_phys; }
        const int p_cnt = p_hi - p_lo;
        const size_t input_bytes = use_runs ? (size_t)(c.run_off[(size_t)p_hi * PM] - c.run_off[(size_t)p_lo * PM]) * sizeof(fgt_run_t)
                                            : bytes_per_ind * p_cnt;
        int n_blocks = 1;
        if (!reuse) {
            if (!in_hbm) n_blocks = (int)std::min<size_t>(p_cnt, std::max<size_t>(1, input_bytes / (48u << 20)));
            else n_blocks = 1;
            if (opt.blocks > 0) n_blocks = std::min(p_cnt, opt.blocks);
        }
        // host-resident paintings: upload the chromosome's tables before the painting copies fill the copy engine
        if (!reuse && !in_hbm) {
            rc = setup_chromosome(pb, h, cs, stx);
            if (rc) return rc;
            if (use_runs) { rc = setup_runs(pb, h, n_blocks, p_lo, p_hi, cs); if (rc) return rc; }
        }
        const void *d_lab = use_runs ? (const void *)c.runs : c.labels;
        std::vector<cudaEvent_t> copy_done(n_blocks, nullptr);
        auto blk_lo = [&](int b) { return p_lo + (int)((long long)p_cnt * b / n_blocks); };
        // byte range of block b inside the chromosome's painting buffer (dense rows, or the rows' runs)
        auto blk_off = [&](int p) { return use_runs ? (size_t)c.run_off[(size_t)p * PM] * sizeof(fgt_run_t) : bytes_per_ind * p; };
        if (!reuse && !in_hbm) {
            void *buf = d_labels_.get(input_bytes);
            if (!buf) return FGT_ERR_CUDA;
            d_lab = (const char *)buf - blk_off(p_lo);       // addressed like the host copy: the buffer holds [p_lo, p_hi) only
            if (labels_free) { CK(cudaStreamWaitEvent(st_copy_, labels_free, 0)); CK(cudaStreamWaitEvent(st_copy2_, labels_free, 0)); }
            const char *src = use_runs ? (const char *)c.runs : (const char *)c.labels;
            for (int b = 0; b < n_blocks; b++) {
                const int p0 = blk_lo(b), p1 = blk_lo(b + 1);
                cudaStream_t cstream = (opt.copy_streams > 1 && (b & 1)) ? st_copy2_ : st_copy_;   // two DMA engines
                CK(cudaMemcpyAsync((char *)buf + (blk_off(p0) - blk_off(p_lo)), src + blk_off(p0), blk_off(p1) - blk_off(p0), cudaMemcpyHostToDevice, cstream));
                copy_done[b] = new_event();
                CK(cudaEventRecord(copy_done[b], cstream));
            }
            stx.h2d_bytes += (int64_t)input_bytes + (use_runs ? (int64_t)rows_all * 16 : 0);
        }
        if (!reuse && in_hbm) {
            // device-resident paintings: the chunk count streams the labels while the host builds the cM prefix
            if (!use_runs && !(count_pending_ && count_key_[0] == h && count_key_[1] == blk_lo(0) && count_key_[2] == blk_lo(1))) {
                rc = prep_count(pb, h, blk_lo(0), blk_lo(1), d_lab, copy_done[0]);
                if (rc) return rc;
            }
            rc = setup_chromosome(pb, h, cs, stx);
            if (rc) return rc;
            if (use_runs) { rc = setup_runs(pb, h, n_blocks, p_lo, p_hi, cs); if (rc) return rc; stx.h2d_bytes += (int64_t)rows_all * 16; }
        }
        int next_n = 0;
        for (int b = 0; b < n_blocks; b++) {
            const int p0 = blk_lo(b), p1 = blk_lo(b + 1);
            cudaEvent_t ready = nullptr;
            if (!reuse) {
                rc = prep_block(pb, h, p0, p1, b, d_lab, copy_done[b], cs, stx);
                if (rc) return rc;
                tick("block prepared");
                ready = new_event();
                CK(cudaEventRecord(ready, st_prep_));
                if (b == n_blocks - 1) labels_free = ready;
            }
            if (count_only) continue;
            if (dist_on) {
                if (b < n_blocks - 1) continue;           // a shard's stream positions need the whole chromosome's counts
                // one all-gather of the pair counts, offsets computed on the device: nothing waits on the host
                long long *msg = pin_msg_;
                CK(cudaStreamSynchronize(st_aux_));       // (the previous chromosome's message has left the staging buffer)
                msg[0] = N;
                for (int n = 0; n < N; n++) msg[1 + n] = cs.totals[used[h][n]];
                for (int n = N; n < n_msg; n++) msg[1 + n] = 0;
                const size_t mlen = (size_t)n_msg + 1;
                CK(cudaMemcpyAsync(d_msg_.p, msg, mlen * sizeof(long long), cudaMemcpyHostToDevice, st_aux_));
                cudaEvent_t msg_up = new_event();
                CK(cudaEventRecord(msg_up, st_aux_));
                CK(cudaStreamWaitEvent(st_, msg_up, 0));
                const int nrc = nccl().AllGather(d_msg_.p, d_msg_all_.p, mlen, kNcclInt64, dist->comm, st_);
                if (nrc != 0) { fprintf(stderr, "fastGLOBETROTTERCompanion(B200): ncclAllGather failed (%d)\n", nrc); return FGT_ERR_CUDA; }
                launch_dist_offsets((const long long *)d_msg_all_.p, dist->world, dist->rank, n_msg, (long long *)d_streampos_.p,
                                    d_pairbase + (size_t)h * N, st_);
                launches_ += 1;
                stx.h2d_bytes += (int64_t)(mlen * sizeof(long long));
            }
            // pseudo-individuals whose packed individual (and all earlier ones in stream order) are ready
            int n_end = next_n;
            if (b == n_blocks - 1) n_end = N;
            else if (monotone[h]) while (n_end < N && used[h][n_end] < p1) n_end++;
            const int min_range = std::max(1, (N + opt.ranges - 1) / std::max(1, opt.ranges));
            if (n_end > next_n && (b == n_blocks - 1 || n_end - next_n >= min_range)) {
                // with everything prepared at once (device-resident paintings) still launch several ranges:
                // the evaluation of one overlaps the latency-bound ordered commit of the previous one
                const int parts = (n_blocks == 1 && opt.split_ready > 1) ? std::min(opt.split_ready, n_end - next_n) : 1;
                for (int q = 0; q < parts; q++) {
                    const int a0 = next_n + (int)((long long)(n_end - next_n) * q / parts);
                    const int a1 = next_n + (int)((long long)(n_end - next_n) * (q + 1) / parts);
                    rc = launch_range(h, a0, a1, ready);
                    if (rc) return rc;
                }
                tick("range launched");
                next_n = n_end;
            }
        }
        if (count_only) {
            for (int n = 0; n < N; n++) count_only[(size_t)h * N + n] = cs.totals[used[h][n]];
            continue;
        }
        if (keep_part) {
            // the projected curves of every individual after this chromosome (mode 1: cumulative tensor; mode 3: this chromosome's)
            double *dst = (double *)d_partials_.p + (size_t)h * N * S * S * G;
            CK(cudaMemsetAsync(dst, 0, (size_t)N * S * S * G * sizeof(double), sc));
            if (tw_pending) {
                const double *tw_src = pb.temp_weights;
                if (tw_stage.valid()) { tw_stage.get(); tw_src = pin_tw_; }
                CK(cudaMemcpyAsync(d_tw, tw_src, tw_bytes, cudaMemcpyHostToDevice, sc));
                tw_pending = false;
            }
            launch_project(d_tensor, N, K, G, S, d_tw, d_pairkh, dst, sc);
            launches_ += 1;
        }
        if (pb.mode == 3 && !(project_per_range && !want_raw)) {
            span_begin(2, sc);
            if (want_raw) { launch_expand_raw(d_tensor, N, K, G, d_raw + (size_t)h * N * K * K * G, sc); launches_++; }
            if (tw_pending) {
                    const double *tw_src = pb.temp_weights;
                    if (tw_stage.valid()) { tw_stage.get(); tw_src = pin_tw_; }
                    CK(cudaMemcpyAsync(d_tw, tw_src, tw_bytes, cudaMemcpyHostToDevice, sc));
                    tw_pending = false;
                }
            (opt.project_impl == 1 ? launch_project_tensor : launch_project)(d_tensor, N, K, G, S, d_tw, d_pairkh, d_results, sc);
            launches_ += 1;
            if (memset_tensor) CK(cudaMemsetAsync(d_tensor, 0, tens_elems * sizeof(double), sc));
            span_end(sc);
        }
    }
    {
        int herr[4];
        CK(cudaStreamSynchronize(st_prep_));
        CK(cudaMemcpy(herr, d_err, sizeof herr, cudaMemcpyDeviceToHost));
        if (count_only || (herr[0] & (kErrMissingDonor | kErrLabelRange))) {
            if (count_only) { carry_ = stx; carry_.kernel_launches = launches_; }
            CK(cudaStreamSynchronize(st_));
            CK(cudaStreamSynchronize(st_commit_));
            for (auto e2 : evs) cudaEventDestroy(e2);
            for (auto &e2 : ev) cudaEventDestroy(e2);
            for (Span &s : spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
            if (herr[0]) return err_to_code(herr[0]);
            return FGT_OK;
        }
    }
    (void)ms_prep_host;
    const long long all_pairs = pb.pair_offset ? pb.pairs_total : run;
    stx.pairs = my_pairs;

    if (!count_only) { last_commit = new_timed_event(); CK(cudaEventRecord(last_commit, sc)); }
    CK(cudaEventRecord(ev[1], sc));
    if (pb.mode == 1 && !(project_per_range && !want_raw)) {
        if (want_raw) { launch_expand_raw(d_tensor, N, K, G, d_raw, sc); launches_++; }
        if (tw_pending) {
                    const double *tw_src = pb.temp_weights;
                    if (tw_stage.valid()) { tw_stage.get(); tw_src = pin_tw_; }
                    CK(cudaMemcpyAsync(d_tw, tw_src, tw_bytes, cudaMemcpyHostToDevice, sc));
                    tw_pending = false;
                }
        (opt.project_impl == 1 ? launch_project_tensor : launch_project)(d_tensor, N, K, G, S, d_tw, d_pairkh, d_results, sc);
        launches_ += 1;
    }
    double *d_rs = d_rs_.as<double>((size_t)N * S * G), *d_ratio = d_ratio_.as<double>((size_t)N * S * S * G);
    double *d_means = d_means_.as<double>((size_t)S * S * G);
    if (!d_rs || !d_ratio || !d_means) return FGT_ERR_CUDA;
    if (dist_on) CK(cudaMemsetAsync(d_means, 0, (size_t)S * S * G * sizeof(double), sc));
    else CK(cudaMemcpyAsync(d_means, means, (size_t)S * S * G * sizeof(double), cudaMemcpyHostToDevice, sc));
    launch_normalise(d_results, N, S, G, pb.num_inds_total > 0 ? pb.num_inds_total : N, d_rs, d_ratio, sc);
    launch_mean_accum(d_ratio, N, S, G, d_means, sc);
    launches_ += 2;
    if (dist_on) {
        // the one all-reduce of the S^2 x G partial curves (every shard ends up with the sum)
        const int nrc = nccl().AllReduce(d_means, d_means, (size_t)S * S * G, kNcclFloat64, kNcclSum, dist->comm, sc);
        if (nrc != 0) { fprintf(stderr, "fastGLOBETROTTERCompanion(B200): ncclAllReduce failed (%d)\n", nrc); return FGT_ERR_CUDA; }
        const size_t cells = (size_t)S * S * G;
        if (pin_part_cap_ < cells) {                 // pinned: a pageable read-back would stall the stream for ~0.1 ms
            if (pin_part_) cudaFreeHost(pin_part_);
            pin_part_ = nullptr; pin_part_cap_ = 0;
            if (cudaHostAlloc((void **)&pin_part_, cells * sizeof(double), cudaHostAllocDefault) != cudaSuccess) return FGT_ERR_CUDA;
            pin_part_cap_ = cells;
        }
        CK(cudaMemcpyAsync(pin_part_, d_means, cells * sizeof(double), cudaMemcpyDeviceToHost, sc));
        CK(cudaMemcpyAsync(pin_msg_ + n_msg + 1, d_streampos_.p, sizeof(long long), cudaMemcpyDeviceToHost, sc));   // pairs drawn by ALL shards
    } else {
        CK(cudaMemcpyAsync(means, d_means, (size_t)S * S * G * sizeof(double), cudaMemcpyDeviceToHost, sc));
        stx.h2d_bytes += (int64_t)S * S * G * 8;
    }
    stx.d2h_bytes += (int64_t)S * S * G * 8;
    CK(cudaEventRecord(ev[2], sc));
    int herr[4];
    unsigned long long hacc[2];
    CK(cudaMemcpyAsync(herr, d_err, sizeof herr, cudaMemcpyDeviceToHost, sc));
    CK(cudaMemcpyAsync(hacc, d_acc, sizeof hacc, cudaMemcpyDeviceToHost, sc));
    if (want_raw) {
        raw_host_.resize(raw_count_);
        CK(cudaMemcpyAsync(raw_host_.data(), d_raw, raw_count_ * sizeof(double), cudaMemcpyDeviceToHost, sc));
    }
    CK(cudaStreamSynchronize(sc));
    CK(cudaStreamSynchronize(st_));
    tick("all done");
    CK(cudaGetLastError());
    if (tables_ready) cudaEventDestroy(tables_ready);
    if (herr[0]) { for (auto e2 : evs) cudaEventDestroy(e2); return err_to_code(herr[0]); }

    if (keep_part) { part_valid_ = true; part_Cn_ = Cn; part_N_ = N; part_S_ = S; part_G_ = G; part_mode_ = pb.mode; }
    if (dist_on) {
        for (size_t i = 0; i < (size_t)S * S * G; i++) means[i] += pin_part_[i];       // added into the caller's buffer (C:644)
        last_total_pairs_ = pin_msg_[n_msg + 1];
    } else last_total_pairs_ = all_pairs;
    // the stream moves on by two draws per sampled pair of the WHOLE call (all shards), C:136-137
    if (!dist || dist->store_state) store_history(apply_jump(poly_pow_E(2ULL * (uint64_t)last_total_pairs_), h0));

    float ms_eval = 0.f, ms_commit = 0.f, ms_rand_spec = 0.f;
    for (Span &s : spans) {
        float t = 0.f;
        cudaEventElapsedTime(&t, s.a, s.b);
        if (s.kind == 0) ms_rand += t; else if (s.kind == 4) ms_rand_spec += t; else if (s.kind == 1) ms_commit += t; else if (s.kind == 3) ms_eval += t; else ms_proj += t;
        if (opt.verbose) {                       // device timeline of the accumulation stages, relative to the start of the call
            static const char *names[] = {"rand", "commit", "project", "eval", "rand(spec)"};
            float t0 = 0.f;
            cudaEventElapsedTime(&t0, ev[0], s.a);
            fprintf(stderr, "[fgt dev] %8.3f .. %8.3f ms  %s\n", t0, t0 + t, names[s.kind]);
        }
        cudaEventDestroy(s.a); cudaEventDestroy(s.b);
    }
    // accumulation = from the first evaluation launch to the end of the last commit (the two phases of
    // consecutive ranges overlap, so their individual spans add up to more than this)
    if (first_eval && last_commit) cudaEventElapsedTime(&ms_accum, first_eval, last_commit);
    ms_accum = std::max(0.f, ms_accum - ms_rand - ms_proj);
    (void)ms_eval; (void)ms_commit;
    float tail = 0.f;
    cudaEventElapsedTime(&tail, ev[1], ev[2]);
    cudaEventElapsedTime(&stx.ms_total, ev[0], ev[2]);
    stx.ms_rand = ms_rand + ms_rand_spec; stx.ms_accum = ms_accum; stx.ms_project = ms_proj + tail;
    stx.ms_chunk = std::max(0.f, stx.ms_total - ms_rand - ms_accum - stx.ms_project);   // chunking/binning not hidden behind the accumulation
    stx.accepted = (int64_t)hacc[0];
    stx.kernel_launches = launches_;
    for (auto e2 : evs) cudaEventDestroy(e2);
    for (auto &e2 : ev) cudaEventDestroy(e2);
    last_stats = stx;
    if (stats) *stats = stx;
    return FGT_OK;
}

// ---------------------------------------------------------------------------------------------
// One call spread over several GPUs of this process ("devices" option): one host thread and one engine per device, the
// pseudo-individuals in contiguous blocks (all chromosomes of an individual on one device: its cells must be folded in
// chromosome order).  The shards exchange their pair counts with one all-gather per chromosome and place their draws on
// the device; one all-reduce sums the partial curves (dist.h).  Raw tensors stay bit-exact, `means` are summed in
// another order than the reference's (<= 1e-12 relative).
// ---------------------------------------------------------------------------------------------
namespace {
struct MultiComms {
    std::mutex mu;
    int n = 0, first_dev = -1;
    NcclComm comm[kMaxEngines] = {};
} g_multi;
int g_last_multi = 1;       // engines that took part in the most recent data_read call (for fgt_get_raw_counts)
int g_last_multi_mode = 1;
}  // namespace

int Engine::data_read_multi(const fgt_problem_t &pb, double *means, fgt_stats_t *stats) {
    Engine &e0 = Engine::get(0);
    g_last_multi = 1;
    int rc = e0.ensure_init();
    if (rc) return rc;
    int ndev = 0;
    cudaGetDeviceCount(&ndev);
    const int n = std::min(std::min(e0.opt.devices, ndev), std::min(pb.num_inds, kMaxEngines));
    if (n <= 1 || pb.pair_offset) return e0.data_read_packed(pb, means, stats, nullptr);
    if (!nccl().loaded) {
        fprintf(stderr, "fastGLOBETROTTERCompanion(B200): \"devices\" > 1 needs NCCL (%s)\n", nccl().why);
        return FGT_ERR_ARG;
    }
    {
        std::lock_guard<std::mutex> lk(g_multi.mu);
        if (g_multi.n != n || g_multi.first_dev != e0.opt.device) {
            for (int i = 0; i < g_multi.n; i++) if (g_multi.comm[i]) nccl().CommDestroy(g_multi.comm[i]);
            int devs[kMaxEngines];
            for (int i = 0; i < n; i++) devs[i] = (e0.opt.device + i) % ndev;
            const int nrc = nccl().CommInitAll(g_multi.comm, n, devs);
            if (nrc != 0) {
                fprintf(stderr, "fastGLOBETROTTERCompanion(B200): ncclCommInitAll failed (%s)\n", nccl().GetErrorString ? nccl().GetErrorString(nrc) : "?");
                g_multi.n = 0;
                return FGT_ERR_CUDA;
            }
            g_multi.n = n; g_multi.first_dev = e0.opt.device;
        }
    }
    History h0;
    if (!e0.snapshot_state(h0)) {
        fprintf(stderr, "fastGLOBETROTTERCompanion(B200): libc rand() is not in its default TYPE_3 state\n");
        return FGT_ERR_RAND;
    }
    const int N = pb.num_inds, Cn = pb.num_chrom;
    const size_t cells = (size_t)pb.num_surrogates * pb.num_surrogates * pb.num_bins;
    std::vector<std::vector<double>> part(n, std::vector<double>(cells, 0.0));
    std::vector<fgt_stats_t> st(n);
    std::vector<int> rcs(n, 0);
    std::vector<std::thread> th;
    const auto t0 = std::chrono::steady_clock::now();
    for (int d = 0; d < n; d++) {
        th.emplace_back([&, d]() {
            const int base = N / n, rem = N % n;
            const int lo = d * base + std::min(d, rem), hi = lo + base + (d < rem ? 1 : 0);
            fgt_problem_t sub = pb;
            sub.num_inds = hi - lo;
            sub.ind_rows = pb.ind_rows + (size_t)lo * Cn;
            sub.num_inds_total = pb.num_inds_total > 0 ? pb.num_inds_total : N;
            DistCtx dc;
            dc.rank = d; dc.world = n; dc.comm = g_multi.comm[d]; dc.h0 = h0.x; dc.store_state = false;
            rcs[d] = Engine::get(d).data_read_packed(sub, part[d].data(), &st[d], nullptr, &dc);
        });
    }
    for (auto &t : th) t.join();
    for (int d = 0; d < n; d++) if (rcs[d]) return rcs[d];
    for (size_t i = 0; i < cells; i++) means[i] += part[0][i];         // every shard holds the all-reduced sum
    e0.advance_state(h0, Engine::get(0).last_total_pairs_);
    fgt_stats_t agg{};
    for (int d = 0; d < n; d++) {
        agg.pairs += st[d].pairs; agg.accepted += st[d].accepted; agg.chunks += st[d].chunks;
        agg.kernel_launches += st[d].kernel_launches; agg.h2d_bytes += st[d].h2d_bytes; agg.d2h_bytes += st[d].d2h_bytes;
        agg.accum_launches += st[d].accum_launches;
        agg.ms_chunk = std::max(agg.ms_chunk, st[d].ms_chunk); agg.ms_rand = std::max(agg.ms_rand, st[d].ms_rand);
        agg.ms_accum = std::max(agg.ms_accum, st[d].ms_accum); agg.ms_project = std::max(agg.ms_project, st[d].ms_project);
    }
    agg.ms_total = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    e0.last_stats = agg;
    if (stats) *stats = agg;
    g_last_multi = n;
    g_last_multi_mode = pb.mode;
    return FGT_OK;
}

// ---------------------------------------------------------------------------------------------
// Callers of the path, opt-in: bootstrap replicates from kept partial curves, R's null normalisation,
// the date likelihood over a grid and the date fit.  Small problems: plain allocate / launch / copy back.
// ---------------------------------------------------------------------------------------------
int Engine::bootstrap_means(const int32_t *ids, int n_rep, double *means_out) {
    int rc = ensure_init();
    if (rc) return rc;
    if (!part_valid_ || !ids || !means_out || n_rep <= 0) return FGT_ERR_ARG;
    CK(cudaSetDevice(device_));
    const int N = part_N_, Cn = part_Cn_, S = part_S_, G = part_G_;
    const size_t cells = (size_t)S * S * G;
    for (long long i = 0; i < (long long)n_rep * N * Cn; i++) if (ids[i] < 0 || ids[i] >= N) return FGT_ERR_ARG;
    int32_t *d_ids = d_physof_.as<int32_t>((size_t)N * Cn);
    double *d_res = d_results_.as<double>((size_t)N * cells), *d_rs = d_rs_.as<double>((size_t)N * S * G);
    double *d_ratio = d_ratio_.as<double>((size_t)N * cells), *d_means = d_means_.as<double>(cells);
    if (!d_ids || !d_res || !d_rs || !d_ratio || !d_means) return FGT_ERR_CUDA;
    for (int r = 0; r < n_rep; r++) {
        CK(cudaMemcpyAsync(d_ids, ids + (size_t)r * N * Cn, (size_t)N * Cn * sizeof(int32_t), cudaMemcpyHostToDevice, st_));
        launch_bootstrap_combine((const double *)d_partials_.p, Cn, N, cells, part_mode_ == 1 ? 1 : 0, d_ids, N, d_res, st_);
        launch_normalise(d_res, N, S, G, N, d_rs, d_ratio, st_);
        CK(cudaMemsetAsync(d_means, 0, cells * sizeof(double), st_));
        launch_mean_accum(d_ratio, N, S, G, d_means, st_);
        CK(cudaMemcpyAsync(means_out + (size_t)r * cells, d_means, cells * sizeof(double), cudaMemcpyDeviceToHost, st_));
        CK(cudaStreamSynchronize(st_));
    }
    CK(cudaGetLastError());
    return FGT_OK;
}

int Engine::null_normalise(const double *tw, const double *null_mat, int S, int K, int G, double *means, double *null_curves) {
    int rc = ensure_init();
    if (rc) return rc;
    if (!tw || !null_mat || S <= 0 || S > 64 || K <= 0 || G <= 0 || (!means && !null_curves)) return FGT_ERR_ARG;
    CK(cudaSetDevice(device_));
    const size_t cells = (size_t)S * S * G, ntw = (size_t)S * S * K * K, nn = (size_t)K * K * G;
    double *d_tw = d_tw_.as<double>(ntw), *d_null = d_tensor_.as<double>(nn), *d_scr = d_results_.as<double>(cells);
    double *d_nc = d_ratio_.as<double>(cells), *d_m = d_means_.as<double>(cells);
    if (!d_tw || !d_null || !d_scr || !d_nc || !d_m) return FGT_ERR_CUDA;
    CK(cudaMemcpyAsync(d_tw, tw, ntw * sizeof(double), cudaMemcpyHostToDevice, st_));
    CK(cudaMemcpyAsync(d_null, null_mat, nn * sizeof(double), cudaMemcpyHostToDevice, st_));
    if (means) CK(cudaMemcpyAsync(d_m, means, cells * sizeof(double), cudaMemcpyHostToDevice, st_));
    launch_null_normalise(d_tw, d_null, S, K, G, d_scr, d_nc, means ? d_m : nullptr, st_);
    if (means) CK(cudaMemcpyAsync(means, d_m, cells * sizeof(double), cudaMemcpyDeviceToHost, st_));
    if (null_curves) CK(cudaMemcpyAsync(null_curves, d_nc, cells * sizeof(double), cudaMemcpyDeviceToHost, st_));
    CK(cudaStreamSynchronize(st_));
    CK(cudaGetLastError());
    return FGT_OK;
}

int Engine::getlhood_grid(const double *means, int R, int T, const double *times, const double *cand, int n_cand, int nparam,
                          const double *popfracs, double likpenalty, double *lhood) {
    int rc = ensure_init();
    if (rc) return rc;
    const int S = (int)(std::sqrt((double)R) + 0.5);
    if (!means || !times || !cand || !popfracs || !lhood || R <= 0 || S * S != R || T < 3 || n_cand <= 0 || (nparam != 1 && nparam != 2))
        return FGT_ERR_ARG;
    CK(cudaSetDevice(device_));
    double *d_m = d_results_.as<double>((size_t)R * T), *d_t = d_rs_.as<double>((size_t)T + S);
    double *d_c = d_ratio_.as<double>((size_t)n_cand * nparam + n_cand);
    void *d_fit = d_tensor_.get(getlhood_scratch_bytes(R, n_cand));
    if (!d_m || !d_t || !d_c || !d_fit) return FGT_ERR_CUDA;
    double *d_pf = d_t + T, *d_lh = d_c + (size_t)n_cand * nparam;
    CK(cudaMemcpyAsync(d_m, means, (size_t)R * T * sizeof(double), cudaMemcpyHostToDevice, st_));
    CK(cudaMemcpyAsync(d_t, times, (size_t)T * sizeof(double), cudaMemcpyHostToDevice, st_));
    CK(cudaMemcpyAsync(d_pf, popfracs, (size_t)S * sizeof(double), cudaMemcpyHostToDevice, st_));
    CK(cudaMemcpyAsync(d_c, cand, (size_t)n_cand * nparam * sizeof(double), cudaMemcpyHostToDevice, st_));
    launch_getlhood(d_m, R, T, d_t, d_c, n_cand, nparam, d_pf, likpenalty, d_fit, d_lh, st_);
    CK(cudaMemcpyAsync(lhood, d_lh, (size_t)n_cand * sizeof(double), cudaMemcpyDeviceToHost, st_));
    CK(cudaStreamSynchronize(st_));
    CK(cudaGetLastError());
    return FGT_OK;
}

// getmax, fastGLOBETROTTER.R:491-520: likelihood on the grid 10^seq(0,3,.25) (13 dates, or the 91 ordered pairs), then a
// Nelder-Mead search in sqrt(date - 1) started at the best grid point (R's optim defaults: reflection 1, contraction 0.5,
// expansion 2, relative tolerance sqrt(eps), 500 iterations).  Every likelihood evaluation runs on the device.
int Engine::getmax(const double *means, int R, int T, const double *times, int nparam, const double *popfracs, double *par, double *value,
                   double *likpenalty_out) {
    if (!par || (nparam != 1 && nparam != 2)) return FGT_ERR_ARG;
    std::vector<double> x(13);
    for (int i = 0; i < 13; i++) x[i] = std::pow(10.0, 0.25 * i);
    std::vector<double> grid;
    if (nparam == 1) grid = x;
    else
        for (int i = 0; i < 13; i++)
            for (int j = i; j < 13; j++) { grid.push_back(x[i]); grid.push_back(x[j]); }       // cbind(rep(x,13:1), rep(x,13)[-to.remove])
    const int n_cand = (int)grid.size() / nparam;
    double pen = 1.0;
    int rc = getlhood_grid(means, R, T, times, grid.data(), 1, nparam, popfracs, 1.0, &pen);
    if (rc) return rc;
    std::vector<double> lh(n_cand);
    rc = getlhood_grid(means, R, T, times, grid.data(), n_cand, nparam, popfracs, pen, lh.data());
    if (rc) return rc;
    double maxlik = 0.0;
    std::vector<double> init(nparam);
    for (int j = 0; j < nparam; j++) init[j] = std::sqrt(grid[j]);
    for (int i = 0; i < n_cand; i++)
        if (lh[i] > maxlik) { maxlik = lh[i]; for (int j = 0; j < nparam; j++) init[j] = std::sqrt(grid[(size_t)i * nparam + j]); }
    int evals = 0, bad = 0;
    auto f = [&](const std::vector<double> &p) {
        double a[2], v = 0.0;
        for (int j = 0; j < nparam; j++) a[j] = 1.0 + p[j] * p[j];
        if (getlhood_grid(means, R, T, times, a, 1, nparam, popfracs, pen, &v)) bad = 1;
        evals++;
        return -v;
    };
    // Nelder-Mead (the structure of R's nmmin)
    const int n = nparam;
    const double alpha = 1.0, beta = 0.5, gamma = 2.0, reltol = 1.4901161193847656e-08;
    std::vector<std::vector<double>> P(n + 1, init);
    std::vector<double> F(n + 1);
    double size = 0.0;
    for (int i = 0; i < n; i++) size = std::max(size, 0.1 * std::fabs(init[i]));
    if (size == 0.0) size = 0.1;
    for (int j = 1; j <= n; j++) P[j][j - 1] += size;
    for (int j = 0; j <= n; j++) F[j] = f(P[j]);
    for (int iter = 0; iter < 500 && !bad; iter++) {
        int L = 0, H = 0;
        for (int j = 1; j <= n; j++) { if (F[j] < F[L]) L = j; if (F[j] > F[H]) H = j; }
        if (F[H] <= F[L] + reltol * (std::fabs(F[L]) + reltol)) break;
        std::vector<double> C(n, 0.0);
        for (int j = 0; j <= n; j++) if (j != H) for (int i = 0; i < n; i++) C[i] += P[j][i] / n;
        std::vector<double> Rp(n), Ep(n);
        for (int i = 0; i < n; i++) Rp[i] = (1.0 + alpha) * C[i] - alpha * P[H][i];
        const double fr = f(Rp);
        if (fr < F[L]) {
            for (int i = 0; i < n; i++) Ep[i] = gamma * Rp[i] + (1.0 - gamma) * C[i];
            const double fe = f(Ep);
            if (fe < fr) { P[H] = Ep; F[H] = fe; } else { P[H] = Rp; F[H] = fr; }
        } else {
            if (fr < F[H]) { P[H] = Rp; F[H] = fr; }
            // second highest
            double second = -1e300;
            for (int j = 0; j <= n; j++) if (j != H && F[j] > second) second = F[j];
            if (fr >= second) {
                for (int i = 0; i < n; i++) Ep[i] = (1.0 - beta) * P[H][i] + beta * C[i];
                const double fc = f(Ep);
                if (fc < F[H]) { P[H] = Ep; F[H] = fc; }
                else if (fr >= F[H]) {
                    for (int j = 0; j <= n; j++)
                        if (j != L) { for (int i = 0; i < n; i++) P[j][i] = beta * (P[j][i] - P[L][i]) + P[L][i]; F[j] = f(P[j]); }
                }
            }
        }
    }
    if (bad) return FGT_ERR_CUDA;
    int L = 0;
    for (int j = 1; j <= n; j++) if (F[j] < F[L]) L = j;
    for (int j = 0; j < nparam; j++) par[j] = 1.0 + P[L][j] * P[L][j];
    if (value) *value = F[L];
    if (likpenalty_out) *likpenalty_out = pen;
    return FGT_OK;
}

int64_t Engine::get_raw_all(double *out, int64_t cap) {
    // raw tensors of the last call: one engine, or the shards of a multi-device call in individual order (mode 1 only:
    // mode 3 stores its tensors in tabulate-call order, chromosome-major, which interleaves the shards)
    if (g_last_multi <= 1) return Engine::get(0).get_raw(out, cap);
    if (g_last_multi_mode != 1) return 0;
    int64_t total = 0;
    for (int d = 0; d < g_last_multi; d++) {
        const int64_t c = Engine::get(d).get_raw(nullptr, 0);
        if (c <= 0) return 0;
        if (out) {
            if (total + c > cap) return FGT_ERR_ARG;
            Engine::get(d).get_raw(out + total, c);
        }
        total += c;
    }
    return total;
}

int64_t Engine::get_raw(double *out, int64_t cap) {
    if (raw_count_ <= 0 || (int64_t)raw_host_.size() < raw_count_) return 0;
    if (out) {
        if (cap < raw_count_) return FGT_ERR_ARG;
        std::memcpy(out, raw_host_.data(), raw_count_ * sizeof(double));
    }
    return raw_count_;
}

// ---------------------------------------------------------------------------------------------
// NULL-individual curves
// ---------------------------------------------------------------------------------------------
// The NULL tensor over several GPUs of this process ("devices" option): every engine folds the label pairs it owns (the
// same owner on every chromosome, so no cell's additions are re-associated) into a zero-initialised copy, the copies have
// disjoint supports and are added to the caller's buffer on the host: bit-identical to one GPU.
int Engine::data_read_null_multi(const fgt_problem_t &pb, const int32_t *rec_rows, double *out, fgt_stats_t *stats) {
    Engine &e0 = Engine::get(0);
    int rc = e0.ensure_init();
    if (rc) return rc;
    int ndev = 0;
    cudaGetDeviceCount(&ndev);
    const int n = std::min(std::min(e0.opt.devices, ndev), kMaxEngines);
    if (n <= 1 || e0.opt.null_world > 1) return e0.data_read_null_packed(pb, rec_rows, out, stats);
    const size_t cells = (size_t)pb.num_donors * pb.num_donors * pb.num_bins;
    std::vector<std::vector<double>> part(n);
    std::vector<fgt_stats_t> st(n);
    std::vector<int> rcs(n, 0);
    std::vector<std::thread> th;
    for (int d = 0; d < n; d++) {
        part[d].assign(out, out + cells);                       // every shard continues from the caller's values on its own cells
        th.emplace_back([&, d]() { rcs[d] = Engine::get(d).data_read_null_packed(pb, rec_rows, part[d].data(), &st[d], n, d); });
    }
    for (auto &t : th) t.join();
    for (int d = 0; d < n; d++) if (rcs[d]) return rcs[d];
    std::fill(out, out + cells, 0.0);
    for (int d = 0; d < n; d++)
        for (size_t i = 0; i < cells; i++) out[i] += part[d][i];      // disjoint supports: exact
    fgt_stats_t agg = st[0];
    for (int d = 1; d < n; d++) {
        agg.accepted += st[d].accepted; agg.kernel_launches += st[d].kernel_launches;
        agg.h2d_bytes += st[d].h2d_bytes; agg.d2h_bytes += st[d].d2h_bytes;
        agg.ms_accum = std::max(agg.ms_accum, st[d].ms_accum); agg.ms_total = std::max(agg.ms_total, st[d].ms_total);
    }
    e0.last_stats = agg;
    if (stats) *stats = agg;
    return FGT_OK;
}

// Who folds label pair (la, lb) in a sharded NULL run.  The owner of a cell must not change between chromosomes (its additions
// would be re-associated).  Two-phase path: whole rows of A labels (the evaluation work is split as well); k_null_strict: label pairs.
static inline bool null_owner(bool two_phase, int la, int lb, int world, int rank) {
    if (world <= 1) return true;
    (void)two_phase;                        // (the two-phase path deals whole rows of A labels by expected work: null_row_owner_)
    return (la * 31 + lb * 17) % world == rank;
}

// Two-phase NULL accumulation of one chromosome (kernels.cu, "NULL path, two phases"): plan the commit tasks from the per-haplotype
// label counts, then per batch of haplotype pairs: count, scan, scatter, ordered commit.
int Engine::null_two_phase(const fgt_problem_t &pb, ChromState &cs, const std::vector<int32_t> &laboff, int Nc, int null_world, int null_rank,
                           double *d_out, unsigned long long *d_acc, fgt_stats_t &stx, double in_window) {
    const int K = pb.num_donors, G = pb.num_bins, P = pb.ploidy, NH = Nc * P;
    auto cl = [&](int hp, int l) { return (long long)(laboff[(size_t)hp * (K + 2) + l + 1] - laboff[(size_t)hp * (K + 2) + l]); };
    std::vector<int32_t> own;
    auto mine_row = [&](int la) { return null_world <= 1 || null_row_owner_[la] == null_rank; };
    for (int la = 0; la < K; la++) if (mine_row(la)) own.push_back(la);
    if (own.empty()) return FGT_OK;
    // evaluated chunk pairs per label pair: sum over haplotypes A of c[A][la] * (chunks labelled lb in the haplotypes of later individuals)
    std::vector<long long> suffix((size_t)(Nc + 1) * K, 0);
    for (int n = Nc - 1; n >= 0; n--)
        for (int l = 0; l < K; l++) {
            long long c = suffix[(size_t)(n + 1) * K + l];
            for (int p = 0; p < P; p++) c += cl(n * P + p, l);
            suffix[(size_t)n * K + l] = c;
        }
    std::vector<double> est((size_t)K * K, 0.0);
    double total_est = 0;
    for (int la : own)
        for (int hA = 0; hA < (Nc - 1) * P; hA++) {
            const long long ca = cl(hA, la);
            if (!ca) continue;
            const long long *sf = &suffix[(size_t)(hA / P + 1) * K];
            for (int lb = 0; lb < K; lb++) est[(size_t)la * K + lb] += (double)ca * (double)sf[lb];
        }
    for (double e : est) total_est += e;
    if (total_est <= 0) return FGT_OK;
    {   // the warps of a CTA claim the A labels of a haplotype pair heaviest first
        std::vector<double> row_est(K, 0.0);
        for (int la : own) for (int lb = 0; lb < K; lb++) row_est[la] += est[(size_t)la * K + lb];
        std::stable_sort(own.begin(), own.end(), [&](int x, int y) { return row_est[x] > row_est[y]; });
    }
    // bin ranges per label pair: hot pairs are cut into up to 32 ranges so that no commit task is much heavier than the rest
    const int want_tasks = opt.null_tasks > 0 ? opt.null_tasks : 4096;     // per shard: enough tasks to fill the GPU's commit CTAs
    int rmax = std::min(32, G);
    std::vector<NullPairPlan> plan((size_t)K * K);
    std::vector<NullRowPlan> rows(K);
    std::vector<int32_t> perm;
    std::vector<CommitTile> tab;
    int n_tasks = 0, max_tb = 1, max_w = 1;
    for (int attempt = 0; attempt < 8; attempt++) {
        const double target = std::max(total_est / want_tasks, 1.0);
        n_tasks = 0; max_tb = 1; max_w = 1;
        struct Cost { double c; int32_t nat; CommitTile t; };
        std::vector<Cost> tl;
        for (int la = 0; la < K; la++) {
            rows[la].task0 = n_tasks; rows[la].ntb = 0;
            const bool mine = mine_row(la);
            for (int lb = 0; lb < K; lb++) {
                NullPairPlan &pp = plan[(size_t)la * K + lb];
                pp.tb0 = rows[la].ntb; pp.w = G; pp.magic = G > 1 ? (uint32_t)((0x100000000ULL + G - 1) / G) : 0u; pp.pad = 0;
                if (!mine) continue;
                const double e = est[(size_t)la * K + lb];
                int R = (int)std::ceil(e / target);
                R = std::max(1, std::min(R, rmax));
                const int w = (G + R - 1) / R;
                const int nr = (G + w - 1) / w;
                pp.w = w; pp.magic = w > 1 ? (uint32_t)((0x100000000ULL + w - 1) / w) : 0u;
                max_w = std::max(max_w, w);
                for (int r = 0; r < nr; r++) {
                    CommitTile t{la * K + lb, 1, r * w, std::min(w, G - r * w)};
                    tl.push_back({e / nr, n_tasks + rows[la].ntb + r, t});
                }
                rows[la].ntb += nr;
            }
            n_tasks += rows[la].ntb;
            max_tb = std::max(max_tb, rows[la].ntb);
        }
        if (max_tb <= 4096) {
            std::stable_sort(tl.begin(), tl.end(), [](const Cost &x, const Cost &y) { return x.c > y.c; });
            perm.resize(tl.size()); tab.resize(tl.size());
            for (size_t q = 0; q < tl.size(); q++) { perm[q] = tl[q].nat; tab[q] = tl[q].t; }
            break;
        }
        rmax = std::max(1, rmax / 2);
    }
    if (max_tb > 4096 || n_tasks <= 0) return FGT_ERR_ARG;
    // haplotype pairs in the reference's order, with an upper bound of the records each can produce
    std::vector<int2> tiles;
    std::vector<long long> bound;
    std::vector<long long> ownA(NH, 0), nlab(NH, 0);
    for (int hp = 0; hp < NH; hp++) {
        nlab[hp] = laboff[(size_t)hp * (K + 2) + K];
        for (int la : own) ownA[hp] += cl(hp, la);
    }
    for (int hA = 0; hA < (Nc - 1) * P; hA++)
        for (int hB = (hA / P + 1) * P; hB < NH; hB++)
            if (ownA[hA] > 0 && nlab[hB] > 0) { tiles.push_back(make_int2(hA, hB)); bound.push_back(ownA[hA] * nlab[hB]); }
    if (tiles.empty()) return FGT_OK;
    // batch size: records of a batch are addressed with 32 bits; the buffer takes a share of the free memory
    vtick("null: plan made");
    if (mem_budget_bytes_ <= 0) {                                // cudaMemGetInfo costs milliseconds: ask once
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        free_b += d_draws_.cap + d_recs_.cap;
        mem_budget_bytes_ = std::min((double)free_b * 0.6, 64e9);
    }
    // Batches are sized by the records they are EXPECTED to produce (the chunk pairs times the share inside the distance window, with
    // a margin); the counting pass then tells the exact number before the record buffer is sized, and a batch that would not fit
    // 32-bit offsets is halved and counted again.
    const bool overlap_batches = opt.null_overlap != 0;
    const double est_frac = std::max(0.05, std::min(1.0, in_window * 1.3 + 0.02));
    for (long long &b : bound) b = std::max<long long>(1, (long long)std::ceil((double)b * est_frac));
    long long batch_evals = opt.null_batch > 0 ? (long long)opt.null_batch : (long long)std::min(1.2e9, mem_budget_bytes_ * 0.8 / sizeof(UpdateRec));
    long long max_bound = 0, total_bound = 0;
    for (long long b : bound) { max_bound = std::max(max_bound, b); total_bound += b; }
    if (!null_pin_ && cudaHostAlloc((void **)&null_pin_, 64, cudaHostAllocDefault) != cudaSuccess) return FGT_ERR_CUDA;
    {   // equal batches instead of full ones and a remainder (every batch pays for a pass over all commit tasks)
        long long nbat = (total_bound + batch_evals - 1) / batch_evals;
        if (opt.null_batch <= 0 && overlap_batches) nbat = std::max(nbat, std::min<long long>(opt.null_min_batches, total_bound / 100000000LL));   // batches to overlap
        batch_evals = std::min(batch_evals, (total_bound + nbat - 1) / nbat + max_bound);
    }
    batch_evals = std::max(batch_evals, max_bound);
    if (batch_evals >= (1LL << 31) - 64) return FGT_ERR_ARG;
    const long long max_tiles = std::max<long long>(1, 100000000LL / n_tasks);     // count / offset tables of at most 400 MB each
    // device tables
    int2 *d_tiles = d_ntiles_.as<int2>(tiles.size());
    int32_t *d_own = d_nown_.as<int32_t>(own.size());
    NullPairPlan *d_plan = d_nplan_.as<NullPairPlan>(plan.size());
    NullRowPlan *d_rows = d_nrows_.as<NullRowPlan>(rows.size());
    int32_t *d_perm = d_nperm_.as<int32_t>(perm.size());
    CommitTile *d_tab = d_ntab_.as<CommitTile>(tab.size());
    int *d_ctr = d_taskctr_.as<int>(1024);
    if (!d_tiles || !d_own || !d_plan || !d_rows || !d_perm || !d_tab || !d_ctr) return FGT_ERR_CUDA;
    CK(cudaMemcpyAsync(d_tiles, tiles.data(), tiles.size() * sizeof(int2), cudaMemcpyHostToDevice, st_));
    CK(cudaMemcpyAsync(d_own, own.data(), own.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st_));
    CK(cudaMemcpyAsync(d_plan, plan.data(), plan.size() * sizeof(NullPairPlan), cudaMemcpyHostToDevice, st_));
    CK(cudaMemcpyAsync(d_rows, rows.data(), rows.size() * sizeof(NullRowPlan), cudaMemcpyHostToDevice, st_));
    CK(cudaMemcpyAsync(d_perm, perm.data(), perm.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st_));
    CK(cudaMemcpyAsync(d_tab, tab.data(), tab.size() * sizeof(CommitTile), cudaMemcpyHostToDevice, st_));
    NullEvalArgs ea{};
    ea.chunks = (const Chunk *)cs.sorted.p; ea.hap_base = (const int32_t *)cs.chunk_base.p; ea.lab_off = (const int32_t *)cs.yoff.p;
    ea.own = d_own; ea.n_own = (int)own.size();
    ea.K = K; ea.G = G; ea.grid_start = pb.grid_start;
    ea.bw = pb.binwidth; ea.half_bw = pb.binwidth / 2.0;
    {
        uint64_t bits;
        std::memcpy(&bits, &pb.binwidth, sizeof bits);
        const bool all_ones = (bits & 0xFFFFFFFFFFFFFULL) == 0xFFFFFFFFFFFFFULL;
        const bool normal = std::isnormal(pb.binwidth) && pb.binwidth > 1e-100 && pb.binwidth < 1e100;
        ea.inv_bw = (!all_ones && normal) ? 1.0 / pb.binwidth : 0.0;
    }
    ea.plan = d_plan; ea.rows = d_rows; ea.n_tasks = n_tasks;
    long long all_lab = 0;
    for (long long c : nlab) all_lab += c;
    // small label groups (large K): the labels of B in blocks of consecutive labels holding at most 32 bin ranges
    const bool window_on = opt.null_window >= 0 ? (opt.null_window != 0) : (in_window < 0.9);
    const double group = (double)all_lab / NH / K;           // chunks of one label on a haplotype, on average
    // (the windowed enumeration works one label at a time: with half the pairs outside the window it beats the blocks down to ~12
    // chunks per group -- K = 100, 300k SNPs: 435 ms against 994 ms in blocks and 888 ms for k_null_strict)
    const bool blocks = opt.null_block >= 0 ? opt.null_block != 0 : (group < 24.0 && !(window_on && group >= 12.0));
    if (blocks) {
        std::vector<int32_t> bstart(K + 1, 0);
        std::vector<int2> blk;
        for (int la = 0; la < K; la++) {
            bstart[la] = (int32_t)blk.size();
            if (!mine_row(la)) continue;
            int lb0 = 0, acc = 0;
            for (int lb = 0; lb < K; lb++) {
                const int nr = (G + plan[(size_t)la * K + lb].w - 1) / plan[(size_t)la * K + lb].w;
                if (lb > lb0 && acc + nr > 32) { blk.push_back(make_int2(lb0, lb)); lb0 = lb; acc = 0; }
                acc += nr;
            }
            blk.push_back(make_int2(lb0, K));
        }
        bstart[K] = (int32_t)blk.size();
        int32_t *d_bs = d_nblks_.as<int32_t>(bstart.size());
        int2 *d_blk = d_nblk_.as<int2>(blk.size() + 1);
        if (!d_bs || !d_blk) return FGT_ERR_CUDA;
        CK(cudaMemcpyAsync(d_bs, bstart.data(), bstart.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st_));
        CK(cudaMemcpyAsync(d_blk, blk.data(), blk.size() * sizeof(int2), cudaMemcpyHostToDevice, st_));
        CK(cudaStreamSynchronize(st_));            // the two vectors are locals
        ea.blk_start = d_bs; ea.blk = d_blk;
    }
    if (null_eval_smem(K, blocks) > 200 * 1024) return FGT_ERR_ARG;
    ea.unit_counter = reinterpret_cast<unsigned long long *>(d_ctr + 8);
    ea.window = window_on ? 1 : 0;
    ea.dwin = (pb.grid_start + G + 1) * pb.binwidth;
    // Two lanes of batches: batch b is folded on the commit stream while batch b+1 is counted and scattered on the evaluation
    // stream (the two kinds of kernel stall on different things); both persistent grids are capped so that they fit an SM together.
    const bool overlap = opt.null_overlap != 0;
    cudaStream_t sc = overlap ? st_commit_ : st_;
    for (int e = 0; e < 3; e++)
        if (!null_ev_[e]) CK(cudaEventCreateWithFlags(&null_ev_[e], cudaEventDisableTiming));
    if (overlap) {                                   // the caller's values and this chromosome's tables are in place
        CK(cudaEventRecord(null_ev_[2], st_));
        CK(cudaStreamWaitEvent(sc, null_ev_[2], 0));
    }
    size_t t0 = 0;
    int nb = 0;
    bool lane_used[2] = {false, false};
    std::vector<cudaEvent_t> vb_events;              // verbose: per batch, events around the count pass, the scatter pass and the commit
    std::vector<long long> vb_recs;
    size_t forced_t1 = 0;                            // a batch that overflowed is retried up to here
    long long left_bound = total_bound;              // expected records of the haplotype pairs not yet in a batch
    while (t0 < tiles.size()) {
        size_t t1 = t0;
        long long sum = 0;
        // the last commit runs with nothing beside it: the work left is dealt out in shrinking batches (a half, a third of a batch)
        long long cap_now = batch_evals;
        if (overlap && opt.null_batch <= 0 && opt.null_taper) {
            if (left_bound <= batch_evals * 3 / 2) cap_now = std::max<long long>(left_bound * 3 / 5, 1);
            if (left_bound <= batch_evals / 2) cap_now = batch_evals;
        }
        while (t1 < tiles.size() && (t1 == t0 || (sum + bound[t1] <= cap_now && (long long)(t1 - t0) < max_tiles)) && (!forced_t1 || t1 < forced_t1))
            sum += bound[t1++];
        forced_t1 = 0;
        const int ln = overlap ? (nb & 1) : 0;
        const long long T = (long long)(t1 - t0);
        const long long n_cnt = (long long)n_tasks * T;
        int32_t *d_cnt = (ln ? d_ncnt2_ : d_ncnt_).as<int32_t>((size_t)n_cnt + 1 + 2 * ((size_t)n_cnt / 1024 + 2));
        int32_t *d_off = (ln ? d_noff2_ : d_noff_).as<int32_t>((size_t)n_cnt + 1);
        RegionEnt *d_dir = (ln ? d_ndir2_ : d_ndir_).as<RegionEnt>(tab.size());
        if (!d_cnt || !d_off || !d_dir) return FGT_ERR_CUDA;
        ea.tiles = d_tiles + t0; ea.n_tiles = (int)T;
        ea.max_ctas_per_sm = (overlap && nb > 0) ? opt.null_eval_ctas : 0;            // a commit of the previous batch shares the SMs
        // counts per (tile, task) -> [task][tile] -> exclusive scan = first record of every (task, tile) -> back to [tile][task]
        CK(cudaMemsetAsync(d_ctr + 8, 0, 2 * sizeof(unsigned long long), st_));
        ea.counts = d_cnt; ea.offs = nullptr; ea.recs = nullptr;
        cudaEvent_t ev_t[6] = {};
        if (opt.verbose) { for (auto &e : ev_t) cudaEventCreate(&e); cudaEventRecord(ev_t[0], st_); }
        launch_null_eval(ea, false, st_);
        if (opt.verbose) cudaEventRecord(ev_t[1], st_);
        launch_sum_i32(d_cnt, n_cnt, reinterpret_cast<unsigned long long *>(d_ctr + 10), st_);
        CK(cudaMemcpyAsync(null_pin_, d_ctr + 10, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st_));
        CK(cudaStreamSynchronize(st_));
        const unsigned long long n_recs = *null_pin_;
        if (n_recs >= (1ULL << 31) - 64) {           // the estimate was too low for this batch: half of it, counted again
            if (T <= 1) return FGT_ERR_ARG;
            forced_t1 = t0 + (size_t)(T / 2);
            continue;
        }
        UpdateRec *d_recs = (ln ? d_recs2_ : d_recs_).as<UpdateRec>((size_t)std::max<unsigned long long>(n_recs, (unsigned long long)std::min<long long>(sum, batch_evals)) + 2);
        if (!d_recs) return FGT_ERR_CUDA;
        ea.recs = d_recs;
        launch_transpose_i32(d_cnt, T, n_tasks, d_off, st_);
        launch_exclusive_scan_i32_large(d_off, d_cnt, n_cnt, d_cnt + n_cnt + 1, st_);
        launch_transpose_i32(d_cnt, n_tasks, T, d_off, st_);
        CK(cudaMemsetAsync(d_ctr + 8, 0, sizeof(unsigned long long), st_));
        ea.counts = nullptr; ea.offs = d_off;
        if (overlap && lane_used[ln]) CK(cudaStreamWaitEvent(st_, null_ev_[ln], 0));     // the lane's records and directory are free again
        if (opt.verbose) cudaEventRecord(ev_t[2], st_);
        launch_null_eval(ea, true, st_);
        if (opt.verbose) cudaEventRecord(ev_t[3], st_);
        launch_null_dir(d_cnt, d_perm, n_tasks, T, d_dir, st_);
        if (overlap) {
            CK(cudaEventRecord(null_ev_[2], st_));
            CK(cudaStreamWaitEvent(sc, null_ev_[2], 0));
        }
        int *ctr = d_ctr + 2 + ln;
        CK(cudaMemsetAsync(ctr, 0, sizeof(int), sc));
        CommitArgs ca{};
        ca.N = 1; ca.K = K; ca.G = G; ca.grid_start = pb.grid_start;
        ca.recs = d_recs; ca.regions = d_dir; ca.tensor_is_zero = 0; ca.buckets = 0;
        ca.tensor = d_out; ca.accepted = d_acc; ca.task_counter = ctr;
        ca.task_tab = d_tab; ca.n_task_tab = n_tasks; ca.max_tile_cells = max_w;
        ca.max_ctas_per_sm = (overlap && t1 < tiles.size()) ? opt.null_commit_ctas : 0;
        ca.consumers = opt.null_consumers; ca.tags = opt.null_tags;   // the next batch's evaluation shares the SMs
        if (opt.verbose) cudaEventRecord(ev_t[4], sc);
        launch_commit_ordered(ca, max_w, sc);
        if (opt.verbose) { cudaEventRecord(ev_t[5], sc); for (auto &e : ev_t) vb_events.push_back(e); vb_recs.push_back((long long)n_recs); }
        if (overlap) { CK(cudaEventRecord(null_ev_[ln], sc)); lane_used[ln] = true; }
        launches_ += 10;
        stx.accum_launches += 1;
        if (opt.verbose) { char msg[96]; snprintf(msg, sizeof msg, "null: batch of %lld haplotype pairs launched (bound %lld records)", T, sum); vtick(msg); }
        left_bound -= sum;
        t0 = t1;
        nb++;
    }
    if (overlap)                                     // the tensor and this chromosome's tables are the evaluation stream's again
        for (int ln = 0; ln < 2; ln++) if (lane_used[ln]) CK(cudaStreamWaitEvent(st_, null_ev_[ln], 0));
    if (opt.verbose && !vb_events.empty()) {
        cudaStreamSynchronize(st_);
        for (size_t b = 0; b * 6 < vb_events.size(); b++) {
            float c = 0, sca = 0, com = 0, c0 = 0, c1 = 0;
            cudaEventElapsedTime(&c, vb_events[b * 6], vb_events[b * 6 + 1]);
            cudaEventElapsedTime(&sca, vb_events[b * 6 + 2], vb_events[b * 6 + 3]);
            cudaEventElapsedTime(&com, vb_events[b * 6 + 4], vb_events[b * 6 + 5]);
            cudaEventElapsedTime(&c0, vb_events[0], vb_events[b * 6 + 4]);
            cudaEventElapsedTime(&c1, vb_events[0], vb_events[b * 6 + 5]);
            fprintf(stderr, "[fgt] dev %d null batch %zu: %lld records, count %.2f ms, scatter %.2f ms, commit %.2f ms (from %.2f to %.2f ms)\n",
                    device_, b, vb_recs[b], c, sca, com, c0, c1);
        }
        for (cudaEvent_t e : vb_events) cudaEventDestroy(e);
    }
    return FGT_OK;
}

int Engine::data_read_null_packed(const fgt_problem_t &pb, const int32_t *rec_rows, double *out, fgt_stats_t *stats, int shard_world,
                                  int shard_rank, NcclComm reduce_comm) {
    const int null_world = shard_world > 0 ? shard_world : opt.null_world, null_rank = shard_world > 0 ? shard_rank : opt.null_rank;
    int rc = ensure_init();
    if (rc) return rc;
    if (pb.ploidy <= 0 || pb.nsamples <= 0 || pb.num_chrom <= 0 || pb.num_inds <= 0 || pb.n_packed_inds <= 0 ||
        pb.num_bins <= 0 || pb.num_donors <= 0 || !(pb.binwidth > 0) || !pb.chrom)
        return FGT_ERR_ARG;
    for (int h = 0; h < pb.num_chrom; h++)
        if (pb.chrom[h].runs ? !pb.chrom[h].run_off : (!pb.chrom[h].labels || (pb.label_bytes != 2 && pb.label_bytes != 4))) return FGT_ERR_ARG;
    CK(cudaSetDevice(device_));
    t_call_start_ = std::chrono::steady_clock::now();
    fgt_stats_t stx{};
    launches_ = 0;
    cudaEvent_t e0, e1, e2;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
    CK(cudaEventRecord(e0, st_));
    const int Nc = std::min(pb.num_inds, 100);                 // reference C:687, C:704-705
    const int K = pb.num_donors, G = pb.num_bins, P = pb.ploidy;
    const int n_haps = Nc * P;
    int *d_err = d_err_.as<int>(4);
    unsigned long long *d_acc = d_accepted_.as<unsigned long long>(2);
    CK(cudaMemsetAsync(d_err, 0, 4 * sizeof(int), st_));
    CK(cudaMemsetAsync(d_acc, 0, 2 * sizeof(unsigned long long), st_));
    int32_t *d_donor = d_donor_.as<int32_t>(pb.n_donor_labels);
    if (!d_donor) return FGT_ERR_CUDA;
    CK(cudaMemcpyAsync(d_donor, pb.donor_label_vec, pb.n_donor_labels * sizeof(int32_t), cudaMemcpyHostToDevice, st_));
    std::vector<int32_t> rowmap(n_haps);
    for (int n = 0; n < Nc; n++) {
        if (rec_rows[n] < 0 || rec_rows[n] >= pb.n_packed_inds) return FGT_ERR_ARG;
        for (int p = 0; p < P; p++) rowmap[n * P + p] = (rec_rows[n] * P + p) * pb.nsamples + 0;   // first sample only, C:748
    }
    int32_t *d_rowmap = d_rowmap_.as<int32_t>(n_haps);
    if (!d_rowmap) return FGT_ERR_CUDA;
    CK(cudaMemcpyAsync(d_rowmap, rowmap.data(), n_haps * sizeof(int32_t), cudaMemcpyHostToDevice, st_));
    double *d_out = d_tensor_.as<double>((size_t)K * K * G);
    if (!d_out) return FGT_ERR_CUDA;
    // null_impl: 2 = two phases (every chunk pair evaluated once), 1 = k_null_strict (round 1: one warp per label pair and slab walks all
    // haplotype pairs), 0 = two phases whenever its tables allow (K <= 4096, G < 65536).  Measured on B200 (profiles/r2_null.md), two
    // phases against k_null_strict: tutorial 3.9x, K = 30 2.0x (100k SNPs) / 26x (400k), K = 100 1.8x / 2.2x (300k), K = 150 2.5x /
    // 3.2x (1M SNPs): with the windowed enumeration and batches sized by expected records it wins on every shape tried.
    const bool tables_ok = K <= 4096 && G < 65536 && (long long)K * K * G < (1LL << 31);     // 16-bit packed plans, 32-bit cell numbers
    bool two_phase = opt.null_impl == 2 && tables_ok;
    // the share of chunk pairs inside the distance window, from the chromosomes' genetic lengths
    std::vector<double> in_window(pb.num_chrom, 1.0);
    double mean_in_window = 0;
    for (int h = 0; h < pb.num_chrom; h++) {
        const fgt_chrom_t &c = pb.chrom[h];
        double cm = 0;
        if (c.pos && c.rate) for (int i = 0; i + 1 < c.nsites; i++) cm += (c.pos[i + 1] - c.pos[i]) * c.rate[i] * 100.0;
        const double window = (pb.grid_start + G) * pb.binwidth;
        in_window[h] = cm > 0 ? std::min(1.0, 2.0 * window / cm) : 1.0;
        mean_in_window += in_window[h] / pb.num_chrom;
    }
    if (opt.null_impl == 0 && tables_ok) two_phase = true;
    (void)mean_in_window;
    std::vector<double> own_init;
    const double *init_src = out;
    if (null_world > 1 && !two_phase) {
        // sharded run: this rank returns only the label pairs it owns (continuing from the caller's values there), zeros
        // elsewhere, so that the ranks' outputs have disjoint supports and their sum is exact
        own_init.assign(out, out + (size_t)K * K * G);
        for (int la = 0; la < K; la++)
            for (int lb = 0; lb < K; lb++)
                if (!null_owner(two_phase, la, lb, null_world, null_rank))
                    std::fill(own_init.begin() + ((size_t)la * K + lb) * G, own_init.begin() + ((size_t)la * K + lb + 1) * G, 0.0);
        init_src = own_init.data();
    }
    CK(cudaMemcpyAsync(d_out, init_src, (size_t)K * K * G * sizeof(double), cudaMemcpyHostToDevice, st_));
    if (null_world > 1 && !two_phase) CK(cudaStreamSynchronize(st_));      // own_init is a local
    CK(cudaEventRecord(e1, st_));
    ChromState &cs = null_cs_;                   // grow-only buffers kept between calls (no cudaMalloc/cudaFree per call)
    long long evals = 0;
    for (int h = 0; h < pb.num_chrom; h++) {
        rc = prep_chromosome(pb, h, d_rowmap, rowmap.data(), n_haps, 1, n_haps, true, cs, stx);
        if (rc) return rc;
        if (Nc < 2 && two_phase && null_world > 1 && h == 0 && null_rank != 0)       // nothing to fold: rank 0 returns the caller's values
            CK(cudaMemsetAsync(d_out, 0, (size_t)K * K * G * sizeof(double), st_));
        if (Nc >= 2) {
            NullArgs a{};
            a.chunks = (const Chunk *)cs.sorted.p; a.hap_base = (const int32_t *)cs.chunk_base.p;
            a.lab_off = (const int32_t *)cs.yoff.p;
            a.n_inds = Nc; a.ploidy = P; a.K = K; a.G = G; a.grid_start = pb.grid_start;
            a.bw = pb.binwidth; a.half_bw = pb.binwidth / 2.0; a.out = d_out; a.accepted = d_acc;
            // task table: per label pair, estimate from the per-haplotype label counts how many chunk pairs it
            // scans and accepts, and split it over as many slabs as keeps a task's fold work near a target
            std::vector<int32_t> laboff((size_t)n_haps * (K + 2));
            CK(cudaMemcpyAsync(laboff.data(), cs.yoff.p, laboff.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, st_));
            CK(cudaStreamSynchronize(st_));
            vtick("null: chunks grouped, label counts on the host");
            if (two_phase && null_world > 1 && h == 0) {
                // Sharded two-phase run: rows of A labels are dealt to the ranks by their expected work on the first chromosome
                // (longest processing time first), the same on every rank (same counts, same rule) and for every chromosome of
                // the call -- a cell's owner must not change, or its additions would be re-associated.  The rows of the other ranks
                // are zeroed: outputs have disjoint supports, their sum is exact.
                std::vector<double> row_est(K, 0.0);
                std::vector<long long> later(Nc + 1, 0);
                for (int n = Nc - 1; n >= 0; n--) {
                    later[n] = later[n + 1];
                    for (int p = 0; p < P; p++) later[n] += laboff[(size_t)(n * P + p) * (K + 2) + K];
                }
                for (int hA = 0; hA < (Nc - 1) * P; hA++)
                    for (int la = 0; la < K; la++)
                        row_est[la] += (double)(laboff[(size_t)hA * (K + 2) + la + 1] - laboff[(size_t)hA * (K + 2) + la]) * (double)later[hA / P + 1];
                std::vector<int> order(K);
                for (int la = 0; la < K; la++) order[la] = la;
                std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return row_est[x] > row_est[y]; });
                std::vector<double> load(null_world, 0.0);
                null_row_owner_.assign(K, 0);
                for (int la : order) {
                    int best = 0;
                    for (int r = 1; r < null_world; r++) if (load[r] < load[best]) best = r;
                    null_row_owner_[la] = best; load[best] += row_est[la];
                }
                for (int la = 0; la < K; la++)
                    if (null_row_owner_[la] != null_rank) CK(cudaMemsetAsync(d_out + (size_t)la * K * G, 0, (size_t)K * G * sizeof(double), st_));
            }
            if (two_phase) {
                rc = null_two_phase(pb, cs, laboff, Nc, null_world, null_rank, d_out, d_acc, stx, in_window[h]);
                if (rc) return rc;
            } else {
            std::vector<double> cnt_mean(K, 0.0);
            for (int hp = 0; hp < n_haps; hp++)
                for (int l = 0; l < K; l++) cnt_mean[l] += (double)(laboff[(size_t)hp * (K + 2) + l + 1] - laboff[(size_t)hp * (K + 2) + l]) / n_haps;
            struct Cost { double c; NullTask t; };
            std::vector<Cost> tl;
            int max_w = 1;
            for (int la = 0; la < K; la++)
                for (int lb = 0; lb < K; lb++) {
                    const double ca = cnt_mean[la], cb = cnt_mean[lb];
                    if (ca <= 0 || cb <= 0) continue;
                    // sharded NULL run: a rank folds the label pairs it owns, on every chromosome (the owner of a cell must
                    // not change between chromosomes, or its additions would be re-associated); the ranks' outputs have
                    // disjoint supports, so their sum is exact
                    if (null_world > 1 && !null_owner(false, la, lb, null_world, null_rank)) continue;
                    const double scan = ca * std::ceil(std::max(cb, 1.0) / 32.0);          // steps per haplotype pair
                    const double recs = ca * cb * 0.7;                                      // accepted pairs per haplotype pair
                    int ns = opt.null_slab_bins > 0 ? (G + opt.null_slab_bins - 1) / opt.null_slab_bins
                                                    : (int)std::lround(recs / (opt.null_target > 0 ? (double)opt.null_target : 96.0));
                    ns = std::max(1, std::min(ns, std::min(G, 64)));
                    const int w = (G + ns - 1) / ns;
                    for (int b = 0; b < G; b += w) {
                        NullTask t{la, lb, b, std::min(G, b + w)};
                        tl.push_back({scan * 40.0 + recs / ns / 32.0 * 300.0, t});
                        max_w = std::max(max_w, t.b_hi - t.b_lo);
                    }
                }
            std::stable_sort(tl.begin(), tl.end(), [](const Cost &x, const Cost &y) { return x.c > y.c; });
            std::vector<NullTask> tasks(tl.size());
            for (size_t q = 0; q < tl.size(); q++) tasks[q] = tl[q].t;
            NullTask *d_tasks = d_slabs_.as<NullTask>(tasks.size() + 1);
            if (!d_tasks) return FGT_ERR_CUDA;
            CK(cudaMemcpyAsync(d_tasks, tasks.data(), tasks.size() * sizeof(NullTask), cudaMemcpyHostToDevice, st_));
            a.tasks = d_tasks; a.n_tasks = (int)tasks.size(); a.max_slab_bins = max_w;
            launch_null_strict(a, st_);
            launches_ += 1;
            stx.accum_launches += 1;
            }
            // evaluated pairs (reference loop count): sum over hap pairs of |A|*|B|
            std::vector<int32_t> ro(n_haps + 1);
            CK(cudaMemcpyAsync(ro.data(), cs.chunk_base.p, (n_haps + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, st_));
            CK(cudaStreamSynchronize(st_));
            std::vector<long long> suffix(Nc + 1, 0);
            for (int n = Nc - 1; n >= 0; n--) {
                long long c = 0;
                for (int p = 0; p < P; p++) c += ro[n * P + p + 1] - ro[n * P + p];
                suffix[n] = suffix[n + 1] + c;
            }
            for (int n = 0; n < Nc - 1; n++) {
                long long c = 0;
                for (int p = 0; p < P; p++) c += ro[n * P + p + 1] - ro[n * P + p];
                evals += c * suffix[n + 1];
            }
        }
    }
    if (reduce_comm && null_world > 1) {
        // one process per GPU: the shards' tensors have disjoint supports -- one all-reduce(sum) leaves the whole tensor on every rank, exact
        if (!nccl().loaded) return FGT_ERR_ARG;
        if (nccl().AllReduce(d_out, d_out, (size_t)K * K * G, kNcclFloat64, kNcclSum, reduce_comm, st_) != 0) return FGT_ERR_CUDA;
    }
    CK(cudaEventRecord(e2, st_));
    vtick("null: all chromosomes launched");
    CK(cudaMemcpyAsync(out, d_out, (size_t)K * K * G * sizeof(double), cudaMemcpyDeviceToHost, st_));
    int herr[4];
    unsigned long long hacc[2];
    CK(cudaMemcpyAsync(herr, d_err, sizeof herr, cudaMemcpyDeviceToHost, st_));
    CK(cudaMemcpyAsync(hacc, d_acc, sizeof hacc, cudaMemcpyDeviceToHost, st_));
    CK(cudaStreamSynchronize(st_));
    vtick("null: done");
    CK(cudaGetLastError());
    if ((herr[0] & kErrMissingDonor) && Nc >= 2) return FGT_ERR_MISSING_DONOR;
    stx.pairs = evals;
    stx.accepted = (int64_t)hacc[0];
    stx.kernel_launches = launches_;
    stx.h2d_bytes += (int64_t)K * K * G * 8;
    stx.d2h_bytes += (int64_t)K * K * G * 8;
    cudaEventElapsedTime(&stx.ms_chunk, e0, e1);
    cudaEventElapsedTime(&stx.ms_accum, e1, e2);
    cudaEventElapsedTime(&stx.ms_total, e0, e2);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
    last_stats = stx;
    if (stats) *stats = stx;
    return FGT_OK;
}

// ---------------------------------------------------------------------------------------------
// small self-test entry points
// ---------------------------------------------------------------------------------------------
int Engine::rand_fill_host(int32_t *out, int64_t n) {
    int rc = ensure_init();
    if (rc) return rc;
    if (n <= 0) return FGT_OK;
    History h0;
    if (!fetch_history(h0)) return FGT_ERR_RAND;
    long long nthreads = (n + kRun - 1) / kRun;
    uint32_t *d = d_draws_.as<uint32_t>((size_t)nthreads * kRun);
    uint32_t *dh = d_hist_.as<uint32_t>(kLag);
    if (!d || !dh) return FGT_ERR_CUDA;
    CK(cudaMemcpyAsync(dh, h0.x, sizeof h0.x, cudaMemcpyHostToDevice, st_));
    launch_rand_fill(dh, d_level_tab_.as<uint32_t>(level_tab_.size()), d, nthreads, st_);
    CK(cudaMemcpyAsync(out, d, n * sizeof(int32_t), cudaMemcpyDeviceToHost, st_));
    CK(cudaStreamSynchronize(st_));
    CK(cudaGetLastError());
    store_history(apply_jump(poly_pow_E((uint64_t)n), h0));
    return FGT_OK;
}

int Engine::outer_weights(double *tempweights, const double *weights_mat, int S, int K) {
    int rc = ensure_init();
    if (rc) return rc;
    if (S <= 0 || K <= 0) return FGT_ERR_ARG;
    CK(cudaSetDevice(device_));
    double *dw = d_rs_.as<double>((size_t)S * K), *dout = d_tw_.as<double>((size_t)S * S * K * K);
    if (!dw || !dout) return FGT_ERR_CUDA;
    CK(cudaMemcpyAsync(dw, weights_mat, (size_t)S * K * sizeof(double), cudaMemcpyHostToDevice, st_));
    launch_outer_weights(dw, S, K, dout, st_);
    CK(cudaMemcpyAsync(tempweights, dout, (size_t)S * S * K * K * sizeof(double), cudaMemcpyDeviceToHost, st_));
    CK(cudaStreamSynchronize(st_));
    CK(cudaGetLastError());
    return FGT_OK;
}

int Engine::chunk_one(const void *labels, int label_bytes, int nsites, const double *pos, const double *rate,
                      double weight_min, double cutoff, const int32_t *donor_label_vec, int n_donor_labels, double *cpos,
                      double *csize, double *cweight, double *cend, int32_t *cpop) {
    int rc = ensure_init();
    if (rc) return rc;
    fgt_problem_t pb{};
    fgt_chrom_t c{};
    c.nsites = nsites; c.pos = pos; c.rate = rate; c.labels = labels; c.labels_on_device = 0;
    pb.mode = 1; pb.ploidy = 1; pb.nsamples = 1; pb.num_chrom = 1; pb.n_packed_inds = 1; pb.label_bytes = label_bytes;
    pb.chrom = &c; pb.num_inds = 1; pb.binwidth = 0.1; pb.num_bins = 10; pb.grid_start = 10; pb.num_donors = 1 << 30;
    pb.donor_label_vec = donor_label_vec; pb.n_donor_labels = n_donor_labels; pb.weight_min = weight_min;
    pb.gridweight_max_cutoff = cutoff; pb.num_surrogates = 1;
    int32_t *d_donor = d_donor_.as<int32_t>(n_donor_labels);
    int *d_err = d_err_.as<int>(4);
    CK(cudaMemsetAsync(d_err, 0, 4 * sizeof(int), st_));
    CK(cudaMemcpyAsync(d_donor, donor_label_vec, n_donor_labels * sizeof(int32_t), cudaMemcpyHostToDevice, st_));
    ChromState cs;
    fgt_stats_t stx{};
    rc = prep_chromosome(pb, 0, nullptr, nullptr, 1, 1, 1, false, cs, stx);
    if (rc) return rc;
    int n = (int)cs.n_chunks;
    std::vector<Chunk> hc(n);
    int herr[4];
    CK(cudaMemcpy(hc.data(), cs.sorted.p, n * sizeof(Chunk), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(herr, d_err, sizeof herr, cudaMemcpyDeviceToHost));
    // one painting: its chunks are position sorted, so the coarse-bin stable sort is the identity
    std::vector<double> hend(n);
    CK(cudaMemcpy(hend.data(), d_cend_.p, n * sizeof(double), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; i++) {
        cpos[i] = hc[i].pos; csize[i] = hc[i].size; cweight[i] = hc[i].w; cpop[i] = hc[i].pop; cend[i] = hend[i];
    }
    cs.sorted.release(); cs.chunk_base.release(); cs.t.release(); cs.first.release(); cs.yoff.release(); cs.rowoff.release();
    if (herr[0] & kErrMissingDonor) return FGT_ERR_MISSING_DONOR;
    return n;
}

}  // namespace fgt
