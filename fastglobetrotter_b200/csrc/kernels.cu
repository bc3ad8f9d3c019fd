// kernels.cu -- sm_100a kernels of the coancestry-curve path.
//
// Every floating-point operation whose result feeds a value that must match the reference
// bit-for-bit is written with the round-to-nearest intrinsics (__dadd_rn, __dmul_rn, __ddiv_rn):
// they are never contracted into FMAs (the reference x86-64 binary has none) and the file is
// additionally compiled with --fmad=false.
#include "kernels.cuh"
#include "dist.h"
#include "glibc_rand.h"

#include <algorithm>

namespace fgt {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// block-wide exclusive prefix of a per-thread flag (0/1); returns the exclusive prefix and the
// block total through `total`.  blockDim.x <= 1024, smem: 32 ints.
__device__ __forceinline__ int block_flag_scan(bool flag, int *warp_tot, int &total) {
    unsigned m = __ballot_sync(kFull, flag);
    int lane = lane_id(), w = warp_id(), nw = (blockDim.x + 31) >> 5;
    int excl = __popc(m & ((1u << lane) - 1));
    if (lane == 0) warp_tot[w] = __popc(m);
    __syncthreads();
    int base = 0, tot = 0;
    for (int i = 0; i < nw; i++) {
        int c = warp_tot[i];
        if (i < w) base += c;
        tot += c;
    }
    __syncthreads();
    total = tot;
    return base + excl;
}

// =============================================================================================
// 1. Chunk builder: run-length encoding of a painting against the cM prefix of the chromosome.
//    reference C:446-455 (count), C:487-510 (locate), same code in mode 3 and the NULL path.
//    Fast path (weight_min < 1): every label change opens a chunk.
// =============================================================================================
// Eight consecutive labels per thread.  For 16-bit labels the eight come from one 16-byte load of
// the flat label matrix (groups are aligned in the flat index space, rows may start anywhere).
template <typename T> struct Lab8 { T v[8]; };

template <typename T>
__device__ __forceinline__ void load8(const T *__restrict__ flat, long long g0, long long lo, long long hi, Lab8<T> &out) {
    // elements [g0, g0+8) of the flat matrix, clipped to the row [lo, hi)
    if (sizeof(T) == 2 && g0 >= lo && g0 + 8 <= hi) {
        const uint4 w = *reinterpret_cast<const uint4 *>(flat + g0);
        const unsigned ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int e = 0; e < 8; e++) out.v[e] = (T)((ww[e >> 1] >> ((e & 1) * 16)) & 0xffffu);
    } else {
#pragma unroll
        for (int e = 0; e < 8; e++) out.v[e] = (g0 + e >= lo && g0 + e < hi) ? flat[g0 + e] : (T)0;
    }
}

// bit e of the result: element g0+e is inside the row, is not its first element, and differs from its predecessor
template <typename T>
__device__ __forceinline__ unsigned change_mask(const T *__restrict__ flat, long long g0, long long lo, long long hi,
                                                const Lab8<T> &x) {
    unsigned m = 0;
    T prev = (g0 - 1 >= lo && g0 - 1 < hi) ? flat[g0 - 1] : (T)0;
#pragma unroll
    for (int e = 0; e < 8; e++) {
        const long long g = g0 + e;
        if (g > lo && g < hi && x.v[e] != prev) m |= 1u << e;
        prev = x.v[e];
    }
    return m;
}

// Streaming layout: one block per 2048-label tile of a row (256 threads x one 16-byte load), every block
// independent -- the count kernel leaves one count per (row, tile) and adds it to the row's total; after a scan of
// the row totals the emit kernel finds its output offset from the row offset plus the counts of the tiles before it.
// The kernels are instruction-bound unless the per-label work is tiny, so groups that lie wholly inside the row
// compare the eight 16-bit labels with their predecessors as four 32-bit words (funnel shift + xor).
constexpr int kTileLabels = 2048;

// first label index (<= lo, > lo - 8) of the row's first 8-label group: groups are 16-byte aligned in memory for
// 16-bit labels whatever the row length and wherever the label block starts
template <typename T>
__device__ __forceinline__ long long tile_first_group(const T *labels, long long lo) {
    const long long off = (long long)((reinterpret_cast<uintptr_t>(labels) / sizeof(T)) & 7);
    return ((lo + off) & ~7LL) - off;
}

template <typename T> struct Group8 {
    Lab8<T> x;
    __device__ __forceinline__ int label(int e) const { return (int)x.v[e]; }
};
template <> struct Group8<uint16_t> {
    uint4 w;                                            // labels 0..7 = low/high halves of w.x, w.y, w.z, w.w
    __device__ __forceinline__ int label(int e) const {
        const unsigned ww = (e >> 1) == 0 ? w.x : (e >> 1) == 1 ? w.y : (e >> 1) == 2 ? w.z : w.w;
        return (int)((ww >> ((e & 1) * 16)) & 0xffffu);
    }
};

// bit e of the result: label g0+e is inside the row, is not its first label, and differs from its predecessor
__device__ __forceinline__ unsigned tile_change_mask(const uint16_t *__restrict__ labels, long long g0, long long lo, long long hi,
                                                     Group8<uint16_t> &g, const uint4 *pre = nullptr) {
    const bool interior = g0 > lo && g0 + 8 <= hi;
    if (__all_sync(kFull, interior)) {
        g.w = pre ? *pre : *reinterpret_cast<const uint4 *>(labels + g0);
        unsigned prev = __shfl_up_sync(kFull, g.w.w, 1) >> 16;
        if (lane_id() == 0) prev = labels[g0 - 1];
        const unsigned dx = g.w.x ^ ((g.w.x << 16) | prev);
        const unsigned dy = g.w.y ^ __funnelshift_l(g.w.x, g.w.y, 16);
        const unsigned dz = g.w.z ^ __funnelshift_l(g.w.y, g.w.z, 16);
        const unsigned dw = g.w.w ^ __funnelshift_l(g.w.z, g.w.w, 16);
        unsigned m = 0;
        m |= (dx & 0xffffu) ? 1u : 0u;  m |= (dx >> 16) ? 2u : 0u;
        m |= (dy & 0xffffu) ? 4u : 0u;  m |= (dy >> 16) ? 8u : 0u;
        m |= (dz & 0xffffu) ? 16u : 0u; m |= (dz >> 16) ? 32u : 0u;
        m |= (dw & 0xffffu) ? 64u : 0u; m |= (dw >> 16) ? 128u : 0u;
        return m;
    }
    // a warp that touches a row end: element-wise with bounds (labels outside the row read as zeros and are masked)
    Lab8<uint16_t> x;
    load8(labels, g0, lo, hi, x);
    g.w.x = x.v[0] | ((unsigned)x.v[1] << 16); g.w.y = x.v[2] | ((unsigned)x.v[3] << 16);
    g.w.z = x.v[4] | ((unsigned)x.v[5] << 16); g.w.w = x.v[6] | ((unsigned)x.v[7] << 16);
    uint16_t prev = (uint16_t)__shfl_up_sync(kFull, (unsigned)x.v[7], 1);
    if (lane_id() == 0) prev = (g0 - 1 >= lo && g0 - 1 < hi) ? labels[g0 - 1] : (uint16_t)0;
    unsigned m = 0;
#pragma unroll
    for (int e = 0; e < 8; e++) {
        const long long q = g0 + e;
        if (q > lo && q < hi && x.v[e] != prev) m |= 1u << e;
        prev = x.v[e];
    }
    return m;
}

__device__ __forceinline__ unsigned tile_change_mask(const int32_t *__restrict__ labels, long long g0, long long lo, long long hi,
                                                     Group8<int32_t> &g, const uint4 * = nullptr) {
    load8(labels, g0, lo, hi, g.x);
    int32_t prev = (int32_t)__shfl_up_sync(kFull, (unsigned)g.x.v[7], 1);
    if (lane_id() == 0) prev = (g0 - 1 >= lo && g0 - 1 < hi) ? labels[g0 - 1] : 0;
    unsigned m = 0;
#pragma unroll
    for (int e = 0; e < 8; e++) {
        const long long q = g0 + e;
        if (q > lo && q < hi && g.x.v[e] != prev) m |= 1u << e;
        prev = g.x.v[e];
    }
    return m;
}

template <typename T>
__global__ void __launch_bounds__(256, 8) k_chunk_count(const T *__restrict__ labels, const int32_t *__restrict__ rowmap, int L,
                                                        int rows, int tiles, int32_t *__restrict__ counts,
                                                        int32_t *__restrict__ rowsum) {
    __shared__ int wtot[8];
    const int row = blockIdx.x / tiles, t = blockIdx.x - row * tiles;       // one tile per block (32-bit index arithmetic)
    const long long lo = (long long)(rowmap ? rowmap[row] : row) * L, hi = lo + L;
    const long long g0 = tile_first_group(labels, lo) + ((long long)t * 256 + threadIdx.x) * 8;
    Group8<T> g;
    int c = __popc(tile_change_mask(labels, g0, lo, hi, g));
    c = warp_sum(c);
    if (lane_id() == 0) wtot[warp_id()] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = (t == 0) ? 1 : 0;                       // the row's first chunk
#pragma unroll
        for (int i = 0; i < 8; i++) s += wtot[i];
        counts[blockIdx.x] = s;
        if (s) atomicAdd(&rowsum[row], s);
    }
}

template <typename T>
__global__ void __launch_bounds__(256, 8) k_chunk_emit(const T *__restrict__ labels, const int32_t *__restrict__ rowmap, int L,
                                                       int rows, int tiles, const int32_t *__restrict__ counts,
                                                       const int32_t *__restrict__ row_off, const double *__restrict__ cum,
                                                       const double *__restrict__ inc, double *__restrict__ cstart,
                                                       double *__restrict__ cend, int32_t *__restrict__ chap) {
    __shared__ int wtot[8];
    __shared__ int open_s;
    __shared__ int2 blist[kTileLabels];                // (SNP index in the row, label) of every boundary of the tile
    const int lane = lane_id(), w = warp_id();
    {
        const int row = blockIdx.x / tiles, t = blockIdx.x - row * tiles;   // one tile per block (32-bit index arithmetic)
        const long long lo = (long long)(rowmap ? rowmap[row] : row) * L, hi = lo + L;
        const long long g0 = tile_first_group(labels, lo) + ((long long)t * 256 + threadIdx.x) * 8;
        // warp 0 also sums the counts of the row's earlier tiles
        int before = 0;
        if (w == 0)
            for (int i = lane; i < t; i += 32) before += counts[(size_t)row * tiles + i];
        const int base = (threadIdx.x == 0) ? row_off[row] : 0;
        Group8<T> g;
        const unsigned m = tile_change_mask(labels, g0, lo, hi, g);
        const int mine = __popc(m);
        int inc_scan = mine;                           // inclusive warp scan of per-thread boundary counts
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(kFull, inc_scan, o);
            if (lane >= o) inc_scan += y;
        }
        if (lane == 31) wtot[w] = inc_scan;
        if (w == 0) {
            // index of the chunk that is open when this tile starts: the row's first chunk plus the boundaries before it
            before = warp_sum(before);
            if (lane == 0) {
                open_s = base + (t == 0 ? 0 : before - 1);
                if (t == 0) {
                    cstart[base] = 0.0;
                    chap[base] = (int32_t)labels[lo];
                    cend[row_off[row + 1] - 1] = cum[L - 1];
                }
            }
        }
        __syncthreads();
        // The tile's boundaries are first listed in shared memory in label order; consecutive threads then write
        // consecutive chunks (a thread emitting its own boundaries directly costs eight sparsely populated rounds
        // of scattered loads and stores per warp).
        int r = inc_scan - mine, tot = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int wt = wtot[i];
            if (i < w) r += wt;
            tot += wt;
        }
        const int j0 = (int)(g0 - lo);
#pragma unroll
        for (int e = 0; e < 8; e++) {                  // static indices keep the labels in registers
            if ((m >> e) & 1u) blist[r++] = make_int2(j0 + e, g.label(e));
        }
        __syncthreads();
        const int c0 = open_s;
        for (int k = threadIdx.x; k < tot; k += 256) {
            const int2 b = blist[k];
            const double cj = cum[b.x];
            cstart[c0 + k + 1] = cj;
            chap[c0 + k + 1] = b.y;
            cend[c0 + k] = __dsub_rn(cj, inc[b.x]);    // reference C:501: cursor minus the same increment
        }
    }
}

// General path (weight_min >= 1, never used by the R driver which passes 0, R:52): a label change
// opens a chunk only if the run that just ended is longer than weight_min SNPs.  One thread per row.
template <typename T, bool EMIT>
__global__ void k_chunk_seq(const T *__restrict__ labels, const int32_t *__restrict__ rowmap, int L, int rows, double weight_min,
                            int32_t *__restrict__ counts, const int32_t *__restrict__ row_off,
                            const double *__restrict__ cum, const double *__restrict__ inc, double *__restrict__ cstart,
                            double *__restrict__ cend, int32_t *__restrict__ chap) {
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    const T *lab = labels + (size_t)(rowmap ? rowmap[row] : row) * L;
    int c = 0, base = EMIT ? row_off[row] : 0;
    if (EMIT) { cstart[base] = 0.0; chap[base] = (int32_t)lab[0]; }
    double run = 1;
    T prev = lab[0];
    for (int j = 1; j < L; j++) {
        T cur = lab[j];
        if (cur == prev) run = run + 1;
        else {
            if (run > weight_min) {
                c++;
                if (EMIT) {
                    double cj = cum[j];
                    cstart[base + c] = cj;
                    chap[base + c] = (int32_t)cur;
                    cend[base + c - 1] = __dsub_rn(cj, inc[j]);
                }
            }
            run = 1;
        }
        prev = cur;
    }
    if (EMIT) cend[base + c] = cum[L - 1];
    else counts[row] = c + 1;
}

int chunk_tiles_for(int nsites) { return (nsites + 7 + kTileLabels - 1) / kTileLabels; }   // +7: rows start anywhere in an aligned group

void launch_chunk_count(const void *labels, const int32_t *rowmap, int label_bytes, int nsites, int rows, double weight_min,
                        int32_t *counts, int32_t *rowsum, cudaStream_t st) {
    if (weight_min < 1.0) {
        const int tiles = chunk_tiles_for(nsites);
        cudaMemsetAsync(rowsum, 0, (size_t)rows * sizeof(int32_t), st);
        // one tile per block: 8 resident blocks per SM at different stages hide the latency better than a persistent
        // loop with a one-tile-ahead prefetch (measured: 131 vs 235 us for 400 MB of labels); rows*tiles < 2^31
        const unsigned grid = (unsigned)((long long)rows * tiles);
        if (label_bytes == 2) k_chunk_count<uint16_t><<<grid, 256, 0, st>>>((const uint16_t *)labels, rowmap, nsites, rows, tiles, counts, rowsum);
        else k_chunk_count<int32_t><<<grid, 256, 0, st>>>((const int32_t *)labels, rowmap, nsites, rows, tiles, counts, rowsum);
    } else {
        int grid = (rows + 127) / 128;
        if (label_bytes == 2)
            k_chunk_seq<uint16_t, false><<<grid, 128, 0, st>>>((const uint16_t *)labels, rowmap, nsites, rows, weight_min, rowsum,
                                                              nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
        else
            k_chunk_seq<int32_t, false><<<grid, 128, 0, st>>>((const int32_t *)labels, rowmap, nsites, rows, weight_min, rowsum,
                                                             nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    }
}

void launch_chunk_emit(const void *labels, const int32_t *rowmap, int label_bytes, int nsites, int rows, double weight_min,
                       const int32_t *counts, const int32_t *row_off, const double *cum, const double *inc, double *cstart,
                       double *cend, int32_t *chap, cudaStream_t st) {
    if (weight_min < 1.0) {
        const int tiles = chunk_tiles_for(nsites);
        const unsigned grid = (unsigned)((long long)rows * tiles);
        if (label_bytes == 2)
            k_chunk_emit<uint16_t><<<grid, 256, 0, st>>>((const uint16_t *)labels, rowmap, nsites, rows, tiles, counts, row_off, cum, inc, cstart, cend, chap);
        else
            k_chunk_emit<int32_t><<<grid, 256, 0, st>>>((const int32_t *)labels, rowmap, nsites, rows, tiles, counts, row_off, cum, inc, cstart, cend, chap);
    } else {
        int grid = (rows + 127) / 128;
        if (label_bytes == 2)
            k_chunk_seq<uint16_t, true><<<grid, 128, 0, st>>>((const uint16_t *)labels, rowmap, nsites, rows, weight_min, nullptr,
                                                             row_off, cum, inc, cstart, cend, chap);
        else
            k_chunk_seq<int32_t, true><<<grid, 128, 0, st>>>((const int32_t *)labels, rowmap, nsites, rows, weight_min, nullptr,
                                                            row_off, cum, inc, cstart, cend, chap);
    }
}

// ---------------------------------------------------------------------------------------------
// Chunk builder on run-length-encoded paintings (fgt_run_t): the parse layer already found the maximal
// runs of equal donor haplotype, so a chunk is a gather of the cM prefix at its run's start (reference
// C:487-510: start = cursor at the first SNP of the run, end = cursor at the next run's first SNP minus that
// SNP's own increment, C:501).  One thread per chunk; the row is found by bisecting the row offsets.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_chunks_from_runs(const uint2 *__restrict__ runs, const long long *__restrict__ src_begin, const int32_t *__restrict__ row_off,
                   int rows, int L, const double *__restrict__ cum, const double *__restrict__ inc,
                   double *__restrict__ cstart, double *__restrict__ cend, int32_t *__restrict__ chap, int *__restrict__ err) {
    const int total = row_off[rows];
    int bad = 0;
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < total; g += gridDim.x * blockDim.x) {
        int lo = 0, hi = rows - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (row_off[mid] <= g) lo = mid; else hi = mid - 1;
        }
        const int k = g - row_off[lo], cnt = row_off[lo + 1] - row_off[lo];
        const uint2 *rr = runs + src_begin[lo];
        const uint2 me = rr[k];
        if (me.x >= (unsigned)L || (k == 0 && me.x != 0u)) { bad |= kErrBadRuns; continue; }
        cstart[g] = (k == 0) ? 0.0 : cum[me.x];
        chap[g] = (int32_t)me.y;
        if (k + 1 < cnt) {
            const uint2 nx = rr[k + 1];
            if (nx.x <= me.x || nx.x >= (unsigned)L || nx.y == me.y) { bad |= kErrBadRuns; continue; }   // runs must be maximal and ascending
            cend[g] = __dsub_rn(cum[nx.x], inc[nx.x]);
        } else cend[g] = cum[L - 1];
    }
    if (bad) atomicOr(err, bad);
}

// weight_min >= 1 on runs (never used by the R driver, R:52): a run boundary opens a chunk only if the maximal run
// that just ended is longer than weight_min SNPs (reference C:446-455); otherwise the open chunk absorbs the next run
// and keeps its own donor haplotype.  One thread per row, same results as k_chunk_seq on the dense matrix.
template <bool EMIT>
__global__ void k_chunk_seq_runs(const uint2 *__restrict__ runs, const long long *__restrict__ src_begin,
                                 const int32_t *__restrict__ n_runs, int rows, int L, double weight_min,
                                 int32_t *__restrict__ counts, const int32_t *__restrict__ row_off,
                                 const double *__restrict__ cum, const double *__restrict__ inc, double *__restrict__ cstart,
                                 double *__restrict__ cend, int32_t *__restrict__ chap, int *__restrict__ err) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    const uint2 *rr = runs + src_begin[row];
    const int cnt = n_runs[row];
    int c = 0;
    const int base = EMIT ? row_off[row] : 0;
    if (cnt <= 0 || rr[0].x != 0u) { atomicOr(err, kErrBadRuns); if (!EMIT) counts[row] = 1; return; }
    if (EMIT) { cstart[base] = 0.0; chap[base] = (int32_t)rr[0].y; }
    unsigned prev_start = 0;
    for (int k = 1; k < cnt; k++) {
        const uint2 r = rr[k];
        if (r.x <= prev_start || r.x >= (unsigned)L) { atomicOr(err, kErrBadRuns); break; }
        const double run = (double)(r.x - prev_start);             // SNPs of the maximal run that ends here
        if (run > weight_min) {
            c++;
            if (EMIT) {
                const double cj = cum[r.x];
                cstart[base + c] = cj;
                chap[base + c] = (int32_t)r.y;
                cend[base + c - 1] = __dsub_rn(cj, inc[r.x]);
            }
        }
        prev_start = r.x;
    }
    if (EMIT) cend[base + c] = cum[L - 1];
    else counts[row] = c + 1;
}

void launch_chunks_from_runs(const fgt_run_t *runs, const long long *src_begin, const int32_t *row_off, int rows, int total,
                             int nsites, const double *cum, const double *inc, double *cstart, double *cend, int32_t *chap,
                             int *err, cudaStream_t st) {
    if (total <= 0) return;
    const unsigned grid = (unsigned)std::min<long long>(((long long)total + 255) / 256, 148LL * 8);
    k_chunks_from_runs<<<grid, 256, 0, st>>>(reinterpret_cast<const uint2 *>(runs), src_begin, row_off, rows, nsites, cum, inc,
                                             cstart, cend, chap, err);
}

void launch_chunk_seq_runs(bool emit, const fgt_run_t *runs, const long long *src_begin, const int32_t *n_runs, int rows,
                           int nsites, double weight_min, int32_t *counts, const int32_t *row_off, const double *cum,
                           const double *inc, double *cstart, double *cend, int32_t *chap, int *err, cudaStream_t st) {
    const int grid = (rows + 127) / 128;
    const uint2 *r2 = reinterpret_cast<const uint2 *>(runs);
    if (emit) k_chunk_seq_runs<true><<<grid, 128, 0, st>>>(r2, src_begin, n_runs, rows, nsites, weight_min, nullptr, row_off, cum, inc, cstart, cend, chap, err);
    else k_chunk_seq_runs<false><<<grid, 128, 0, st>>>(r2, src_begin, n_runs, rows, nsites, weight_min, counts, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, err);
}

// exclusive scan of int32 counts -> out[0..n] (single block; n is "rows", at most a few 10^5)
__global__ void __launch_bounds__(1024) k_exclusive_scan_i32(const int32_t *__restrict__ in, int32_t *__restrict__ out, int n) {
    __shared__ int wtot[32];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int t0 = 0; t0 < n; t0 += blockDim.x) {
        int i = t0 + threadIdx.x;
        int v = i < n ? in[i] : 0;
        int x = v;   // inclusive warp scan
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(kFull, x, o);
            if (lane_id() >= o) x += y;
        }
        if (lane_id() == 31) wtot[warp_id()] = x;
        __syncthreads();
        int base = carry_s;
        for (int w = 0; w < warp_id(); w++) base += wtot[w];
        if (i < n) out[i] = base + x - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry_s = base + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry_s;
}

void launch_exclusive_scan_i32(const int32_t *in, int32_t *out, int n, cudaStream_t st) {
    k_exclusive_scan_i32<<<1, 1024, 0, st>>>(in, out, n);
}

void launch_chunk_scan(const int32_t *rowsum, int32_t *row_off, int rows, cudaStream_t st) {
    k_exclusive_scan_i32<<<1, 1024, 0, st>>>(rowsum, row_off, rows);
}

// =============================================================================================
// 2. Per-individual coarse binning (reference C:57-111): finish the chunk records, stable-sort
//    them by 1-cM bin, and tabulate how many pairs every (j,k) bin pair will sample
//    (reference C:121-131 / C:1128-1140) so that rand() stream offsets are a prefix sum.
//    One CTA per packed individual.
// =============================================================================================
__device__ __forceinline__ int pairs_to_sample(int t1, int t2, double e, double divisor) {
    // Y_i = (int) t_1*t_2*exp(-0.05*(k-j))/divisor ; if (t_1>0 && t_2>0 && Y_i==0) Y_i=1
    int prod = (int)((unsigned)t1 * (unsigned)t2);
    double y = __ddiv_rn(__dmul_rn((double)prod, e), divisor);
    int Y = __double2int_rz(y);
    if (t1 > 0 && t2 > 0 && Y == 0) Y = 1;
    return Y < 0 ? 0 : Y;
}

__global__ void __launch_bounds__(1024)
k_bin_sort(int rows_per_ind, const int32_t *__restrict__ row_off, const double *__restrict__ cstart,
           const double *__restrict__ cend, const int32_t *__restrict__ chap, const int32_t *__restrict__ donor_label,
           int n_donor_labels, double cutoff, int K, int mode, int nb, int W, const double *__restrict__ Etab,
           double divisor, Chunk *__restrict__ sorted, long long n_sorted, const Chunk **__restrict__ chunk_ptr,
           const double2 **__restrict__ ps_ptr, const int32_t **__restrict__ pop_ptr, int32_t *__restrict__ t_out,
           int32_t *__restrict__ first_out, int32_t *__restrict__ yoff, long long *__restrict__ rowoff,
           int32_t *__restrict__ cntmat, int32_t *__restrict__ runfirst, int *__restrict__ err, long long *__restrict__ totals) {
    const int p = blockIdx.x;
    const int r0 = p * rows_per_ind;
    const int c0 = row_off[r0];
    int32_t *cnt = cntmat + (size_t)p * rows_per_ind * nb;      // zero on entry
    int32_t *rf = runfirst + (size_t)p * rows_per_ind * nb;
    int32_t *t = t_out + (size_t)p * nb;
    int32_t *first = first_out + (size_t)p * nb;
    // compact copies for phase 1 live behind the 32-byte records of the same buffer: (pos, size) pairs, then labels
    double2 *ps = reinterpret_cast<double2 *>(sorted + n_sorted);
    int32_t *pops = reinterpret_cast<int32_t *>(ps + n_sorted);
    if (threadIdx.x == 0) {
        chunk_ptr[p] = sorted + c0;
        if (ps_ptr) { ps_ptr[p] = ps + c0; pop_ptr[p] = pops + c0; }
    }
    // phase 1: occupancy of (painting, coarse bin); first chunk of each run.  The individual's chunks are one
    // contiguous range; a thread finds the painting of its chunk by bisecting the row offsets.
    const int c_end = row_off[r0 + rows_per_ind];
    auto row_of = [&](int i) {
        int lo = 0, hi = rows_per_ind - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (row_off[r0 + mid] <= i) lo = mid; else hi = mid - 1;
        }
        return lo;
    };
    for (int i = c0 + threadIdx.x; i < c_end; i += blockDim.x) {
        const int r = row_of(i);
        const int a = row_off[r0 + r];
        double pos = __dmul_rn(__dadd_rn(cstart[i], cend[i]), 0.5);
        int bin = __double2int_rz(pos);
        bin = bin < 0 ? 0 : (bin >= nb ? nb - 1 : bin);
        atomicAdd(&cnt[r * nb + bin], 1);
        bool is_first = (i == a);
        if (!is_first) {
            double pp = __dmul_rn(__dadd_rn(cstart[i - 1], cend[i - 1]), 0.5);
            int pb = __double2int_rz(pp);
            pb = pb < 0 ? 0 : (pb >= nb ? nb - 1 : pb);
            is_first = pb != bin;
        }
        if (is_first) rf[r * nb + bin] = i;
    }
    __syncthreads();
    // phase 2: per bin, turn the per-painting counts into exclusive prefixes; t[bin] = total
    for (int bin = threadIdx.x; bin < nb; bin += blockDim.x) {
        int s = 0;
        for (int r = 0; r < rows_per_ind; r++) {
            int c = cnt[r * nb + bin];
            cnt[r * nb + bin] = s;
            s += c;
        }
        t[bin] = s;
    }
    __syncthreads();
    // phase 3: start_bin = exclusive prefix of t (one warp, 32 bins per step)
    if (threadIdx.x < 32) {
        int carry = 0;
        for (int b0 = 0; b0 < nb; b0 += 32) {
            const int bin = b0 + (int)threadIdx.x;
            const int v = bin < nb ? t[bin] : 0;
            int x = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(kFull, x, o);
                if ((int)threadIdx.x >= o) x += y;
            }
            if (bin < nb) first[bin] = carry + x - v;
            carry += __shfl_sync(kFull, x, 31);
        }
    }
    __syncthreads();
    // phase 4: stable placement + finished records
    int bad = 0;
    for (int i = c0 + threadIdx.x; i < c_end; i += blockDim.x) {
        const int r = row_of(i);
        double s = cstart[i], e = cend[i];
        double pos = __dmul_rn(__dadd_rn(s, e), 0.5);
        int bin = __double2int_rz(pos);
        bin = bin < 0 ? 0 : (bin >= nb ? nb - 1 : bin);
        int dest = c0 + first[bin] + cnt[r * nb + bin] + (i - rf[r * nb + bin]);
        double len = __dsub_rn(e, s);
        int hap = chap[i];
        int pop = (hap >= 1 && hap <= n_donor_labels) ? donor_label[hap - 1] : -9;
        if (pop == -9) bad |= kErrMissingDonor;                 // reference C:517-522
        if (mode == 1 && pop > K) bad |= kErrLabelRange;        // reference C:102-107
        if (mode == 3 && (pop > K)) bad |= kErrLabelRange;      // (would index out of bounds in the reference)
        Chunk c;
        c.pos = pos; c.size = len; c.w = (len > cutoff) ? cutoff : len; c.pop = pop; c.hapid = hap;
        sorted[dest] = c;
        if (ps_ptr) { ps[dest] = make_double2(pos, len); pops[dest] = pop; }
    }
    if (bad) atomicOr(err, bad);
    // phase 5: pairs per row of the (j,k) sampling schedule
    for (int j = threadIdx.x; j < nb; j += blockDim.x) {
        int32_t *yo = yoff + ((size_t)p * nb + j) * W;
        int t1 = t[j];
        int acc = 0;
        yo[0] = 0;
        for (int d = 1; d < W; d++) {
            int k = j + d;
            if (j < nb - 1 && k < nb) acc += pairs_to_sample(t1, t[k], Etab[d], divisor);
            yo[d] = acc;
        }
    }
    __syncthreads();
    if (threadIdx.x < 32) {                             // exclusive prefix over j of the row totals (one warp)
        long long carry = 0;
        long long *ro = rowoff + (size_t)p * (nb + 1);
        for (int j0 = 0; j0 < nb; j0 += 32) {
            const int j = j0 + (int)threadIdx.x;
            const long long v = j < nb ? (long long)yoff[((size_t)p * nb + j) * W + (W - 1)] : 0;
            long long x = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long y = __shfl_up_sync(kFull, x, o);
                if ((int)threadIdx.x >= o) x += y;
            }
            if (j < nb) ro[j] = carry + x - v;
            carry += __shfl_sync(kFull, x, 31);
        }
        if (threadIdx.x == 0) {
            ro[nb] = carry;
            if (totals) totals[p] = carry;
        }
    }
}

void launch_bin_sort(int n_phys, int rows_per_ind, const int32_t *row_off, const double *cstart, const double *cend,
                     const int32_t *chap, const int32_t *donor_label, int n_donor_labels, double cutoff, int K, int mode,
                     int nb, int W, const double *Etab, double divisor, Chunk *sorted, long long n_sorted, const Chunk **chunk_ptr,
                     const double2 **ps_ptr, const int32_t **pop_ptr, int32_t *t,
                     int32_t *first, int32_t *yoff, long long *rowoff, int32_t *cntmat, int32_t *runfirst, int *err,
                     long long *totals, cudaStream_t st) {
    k_bin_sort<<<n_phys, 1024, 0, st>>>(rows_per_ind, row_off, cstart, cend, chap, donor_label, n_donor_labels, cutoff, K,
                                      mode, nb, W, Etab, divisor, sorted, n_sorted, chunk_ptr, ps_ptr, pop_ptr, t, first, yoff, rowoff, cntmat,
                                      runfirst, err, totals);
}

// =============================================================================================
// 3. glibc rand() stream, generated in bulk.  Thread t produces draws [t*kRun, (t+1)*kRun) of the
//    buffer; its start state is the buffer's base history advanced by the base-256 digits of t
//    (level tables from glibc_rand.h).  x[n] = x[n-31] + x[n-3]; output x[n] >> 1.
// =============================================================================================
__device__ __forceinline__ void block_apply_level(uint32_t *y /*smem[61]*/, const uint32_t *__restrict__ tab, int level,
                                                  int digit) {
    // y[0..30] holds a history; replace it by the history `digit * kRun * 256^level` draws later
    if (threadIdx.x == 0)
        for (int m = 0; m < kLag - 1; m++) y[kLag + m] = y[m] + y[m + kTap];
    __syncthreads();
    uint32_t s = 0;
    if (threadIdx.x < kLag) {
        for (int i = 0; i < kLag; i++) s += tab[((size_t)level * kLag + i) * 256 + digit] * y[i + threadIdx.x];
    }
    __syncthreads();
    if (threadIdx.x < kLag) y[threadIdx.x] = s;
    __syncthreads();
}

__global__ void __launch_bounds__(256)
k_rand_fill(const uint32_t *__restrict__ base_hist, const uint32_t *__restrict__ tab, uint32_t *__restrict__ out,
            long long n_threads, unsigned blk0) {
    __shared__ uint32_t y[2 * kLag - 1];
    if (threadIdx.x < kLag) y[threadIdx.x] = base_hist[threadIdx.x];
    __syncthreads();
    unsigned long long blk = (unsigned long long)blockIdx.x + blk0;
#pragma unroll 1
    for (int lv = 1; lv < kLevels; lv++) {
        int digit = (int)((blk >> (8 * (lv - 1))) & 255);
        if (digit) block_apply_level(y, tab, lv, digit);
    }
    if (threadIdx.x == 0)
        for (int m = 0; m < kLag - 1; m++) y[kLag + m] = y[m] + y[m + kTap];
    __syncthreads();
    uint32_t yy[2 * kLag - 1];
#pragma unroll
    for (int i = 0; i < 2 * kLag - 1; i++) yy[i] = y[i];
    uint32_t s[kLag];
#pragma unroll
    for (int j = 0; j < kLag; j++) s[j] = 0;
#pragma unroll
    for (int i = 0; i < kLag; i++) {
        uint32_t c = tab[(size_t)i * 256 + threadIdx.x];     // level 0, digit = threadIdx.x
#pragma unroll
        for (int j = 0; j < kLag; j++) s[j] += c * yy[i + j];
    }
    // Each thread owns a contiguous run of the stream, so its stores would be 32 scattered sectors per
    // instruction.  Four octets (128 B) per thread are staged in shared memory (XOR-swizzled 16-byte
    // blocks, conflict free) and flushed as whole 128-byte lines: 8 lanes per line, 4 lines per store.
    extern __shared__ __align__(16) uint4 stage_all[];            // [warps][32 rows][8 blocks]
    uint4 *stg = stage_all + (size_t)warp_id() * 256;
    const int lane = lane_id();
    uint4 *grun = reinterpret_cast<uint4 *>(out) + ((long long)blk * 256 + (threadIdx.x & ~31)) * (kRun / 4);
    // grun: first uint4 of lane 0's run; lane r's run starts r * (kRun/4) uint4 later
    int flushed = 0;                                               // 128-byte groups already written per thread
#pragma unroll 1
    for (int it = 0; it < kRun / (8 * kLag); it++) {
#pragma unroll
        for (int v = 0; v < kLag; v++) {                           // 31 octets: the rotating state index returns to 0
            uint32_t o[8];
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const int i = (8 * v + r) % kLag;
                s[i] += s[(i + kTap) % kLag];
                o[r] = s[i] >> 1;
            }
            const int oct = (it * kLag + v) & 3;                   // position of this octet inside the 128-byte group
            stg[lane * 8 + ((2 * oct) ^ (lane & 7))] = make_uint4(o[0], o[1], o[2], o[3]);
            stg[lane * 8 + ((2 * oct + 1) ^ (lane & 7))] = make_uint4(o[4], o[5], o[6], o[7]);
            if (oct == 3) {
                __syncwarp();
#pragma unroll
                for (int i4 = 0; i4 < 8; i4++) {
                    const int row = 4 * i4 + (lane >> 3), cb = lane & 7;
                    const uint4 w = stg[row * 8 + (cb ^ (row & 7))];
                    if ((long long)blk * 256 + (threadIdx.x & ~31) + row < n_threads)
                        grun[(long long)row * (kRun / 4) + flushed * 8 + cb] = w;
                }
                flushed++;
                __syncwarp();
            }
        }
    }
}

void launch_rand_fill(const uint32_t *base_hist, const uint32_t *level_tab, uint32_t *out, long long n_threads,
                      cudaStream_t st, long long first_thread) {
    // threads [first_thread, n_threads) of the stream that starts at base_hist; first_thread is a multiple of 256
    if (n_threads <= first_thread) return;
    const long long blk0 = first_thread / 256;
    long long blocks = (n_threads + 255) / 256 - blk0;
    k_rand_fill<<<(unsigned)blocks, 256, 8 * 256 * sizeof(uint4), st>>>(base_hist, level_tab, out, n_threads, (unsigned)blk0);
}

// =============================================================================================
// 4. Sampled pair accumulation, strict (order-preserving) mode.
//    reference tabulate_chunks C:121-168 / tabulate_chunks_mode3 C:1128-1177.
//
//    One warp owns one (pseudo-individual, slab of fine bins) and is the ONLY writer of those
//    cells, so it may fold updates in the reference's order without atomics: it walks the
//    sampling schedule (j ascending, k ascending, draw ascending) restricted to the coarse
//    separations that can land in its slab, evaluates 32 consecutive sampled pairs per step,
//    and commits them in lane order; lanes that hit the same cell are chained through
//    shuffles so the additions happen in sequence order.
// =============================================================================================
__device__ __forceinline__ int tri_row(int lo, int hi, int K) { return lo * K - ((lo * (lo - 1)) >> 1) + (hi - lo); }

__global__ void __launch_bounds__(128) k_accum_strict(const AccumArgs a) {
    const int lane = lane_id();
    const int task = blockIdx.x * (blockDim.x >> 5) + warp_id();
    if (task >= a.N * a.n_slabs) return;
    const int slab_i = task / a.N;      // slab-major: the heavy (small separation) slabs start first
    const int n = task - slab_i * a.N;
    const SlabDesc sl = a.slabs[slab_i];
    const int phys = a.phys_of[n];
    const ChromPrep &cp = a.cp;
    const int nb = cp.nb, W = cp.W;
    const int32_t *T = cp.t + (size_t)phys * nb;
    const int32_t *F = cp.first + (size_t)phys * nb;
    const long long *RO = cp.rowoff + (size_t)phys * (nb + 1);
    const Chunk *ch = cp.chunk_ptr[phys];
    const uint2 *dr = a.draws + a.pair_base[n];
    const int G = a.G, K = a.K;
    double *tens = a.tensor + (size_t)n * ((size_t)K * (K + 1) / 2) * G;
    const double bw = a.bw, half_bw = a.half_bw;
    unsigned long long n_acc = 0;
    int bad = 0;

    for (int j = 0; j < nb - 1; j++) {
        const int t1 = T[j];
        if (t1 <= 0) continue;
        int dmax = W - 1;
        if (nb - 1 - j < dmax) dmax = nb - 1 - j;
        const int dlo = sl.dlo, dhi = sl.dhi < dmax ? sl.dhi : dmax;
        if (dlo > dhi) continue;
        const int32_t *yo = cp.yoff + ((size_t)phys * nb + j) * W;
        const int o_lo = yo[dlo - 1], o_hi = yo[dhi];
        const int R = o_hi - o_lo;
        if (R <= 0) continue;
        const uint2 *drow = dr + RO[j] + o_lo;
        const int f1 = F[j];
        for (int base = 0; base < R; base += 32) {
            const int q = base + lane;
            bool valid = q < R;
            int cell = -1;
            double val = 0.0;
            if (valid) {
                // which separation does draw q belong to: smallest d with q + o_lo < yo[d]
                int lo = dlo, hi = dhi;
                const int target = q + o_lo;
                while (lo < hi) {
                    int mid = (lo + hi) >> 1;
                    if (target < yo[mid]) hi = mid; else lo = mid + 1;
                }
                const int k = j + lo;
                const int t2 = T[k];
                const uint2 r = drow[q];
                const int c1 = f1 + (int)(r.x % (unsigned)t1);
                const int c2 = F[k] + (int)(r.y % (unsigned)t2);
                const Chunk A = ch[c1], B = ch[c2];
                bool live = true;
                if (a.mode == 1) live = (A.pop != 0) && (B.pop != 0);          // reference C:138-142
                else if (A.pop == 0 || B.pop == 0) { bad |= kErrMode3Target; live = false; }
                if (live) {
                    const double d = fabs(__dsub_rn(A.pos, B.pos));             // pow(pow(x,2),0.5), reference C:143
                    const double qd = __ddiv_rn(d, bw);
                    const double fl = floor(qd);
                    const double frac = __dmul_rn(__dsub_rn(qd, fl), bw);       // reference C:144-145
                    int bfull = __double2int_rz(fl);
                    bool ok = true;
                    if (frac <= half_bw) { /* floor */ }
                    else if (frac > half_bw) bfull += 1;
                    else ok = false;                                             // NaN: the reference keeps a stale bin
                    const double mind = __dadd_rn(__dmul_rn(A.size, 0.5), __dmul_rn(B.size, 0.5));
                    if (ok && bfull >= sl.b0 && bfull <= sl.b1 && d >= mind) {
                        const int l1 = A.pop - 1, l2 = B.pop - 1;
                        const int lo_l = l1 < l2 ? l1 : l2, hi_l = l1 < l2 ? l2 : l1;
                        cell = tri_row(lo_l, hi_l, K) * G + (bfull - a.grid_start);
                        val = __dmul_rn(A.w, B.w);
                    }
                }
            }
            // ---- ordered commit -------------------------------------------------------------
            const unsigned key = cell >= 0 ? (unsigned)cell : (0x80000000u | (unsigned)lane);
            const unsigned grp = __match_any_sync(kFull, key);
            const bool leader = (cell >= 0) && ((grp & ((1u << lane) - 1)) == 0);
            double acc = 0.0;
            if (leader) acc = tens[cell];
            unsigned rest = leader ? (grp & ~(1u << lane)) : 0u;
            acc = __dadd_rn(acc, val);
            while (__any_sync(kFull, rest != 0)) {
                const int src = rest ? (__ffs(rest) - 1) : lane;
                const double v = __shfl_sync(kFull, val, src);
                if (rest) { acc = __dadd_rn(acc, v); rest &= rest - 1; }
            }
            if (leader) tens[cell] = acc;
            n_acc += (cell >= 0);
            __syncwarp();
        }
    }
    n_acc = warp_sum_u64(n_acc);
    if (lane == 0 && n_acc) atomicAdd(a.accepted, n_acc);
    if (bad) atomicOr(a.err, bad);
}

void launch_accum_strict(const AccumArgs &a, cudaStream_t st) {
    long long tasks = (long long)a.N * a.n_slabs;
    if (tasks <= 0) return;
    const int wpb = 4;
    unsigned grid = (unsigned)((tasks + wpb - 1) / wpb);
    k_accum_strict<<<grid, wpb * 32, 0, st>>>(a);
}

// =============================================================================================
// 4b. Two-phase strict accumulation.
// =============================================================================================
// a % d for a < 2^31 (a rand() draw) and small d: with M = floor((2^32 - 1) / d) + 1 >= 2^32 / d the estimate
// q = floor(a * M / 2^32) is floor(a / d) or one more (the excess a * (M * d - 2^32) / (d * 2^32) is below 1/2), so one
// correction of a negative remainder makes it exact.  M == 0 encodes d == 1.
__device__ __forceinline__ unsigned mod_magic(unsigned d) { return d > 1 ? 0xFFFFFFFFu / d + 1u : 0u; }
__device__ __forceinline__ unsigned fastmod_u32(unsigned a, unsigned d, unsigned magic) {
    const unsigned q = __umulhi(a, magic);
    const int r = (int)(a - q * d);
    return magic ? (unsigned)(r < 0 ? r + (int)d : r) : 0u;
}

// One 256-bit load per chunk record (LDG.E.256): a random gather of a 32-byte record then costs one L1
// wavefront per lane instead of two (the evaluation kernel is L1-wavefront bound on these gathers).
__device__ __forceinline__ Chunk ldg_chunk(const Chunk *p) {
    unsigned long long a, b, c, d;
    asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];\n" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    Chunk o;
    o.pos = __longlong_as_double((long long)a);
    o.size = __longlong_as_double((long long)b);
    o.w = __longlong_as_double((long long)c);
    o.pop = (int32_t)(d & 0xffffffffull);
    o.hapid = (int32_t)(d >> 32);
    return o;
}

// Phase 1: one warp per (pseudo-individual, coarse bin j); walks k = j+1.. in order, 64 sampled pairs
// (two per lane, for memory-level parallelism) per iteration.  Everything of reference C:136-162
// except the "+=" happens here.
struct PairEval { int cell, local, slot; double val; };

// Phase 1 gathers the compact copy of a chunk: one 16-byte (pos, size) pair -- 8 per cache line instead of 4 -- and its
// label (32 per line): the kernel is bound by the L1 wavefronts of these scattered gathers.
struct PairChunk { double pos, size; int pop; };
struct BinRef { const double2 *ps; const int32_t *pop; };
__device__ __forceinline__ PairChunk ldg_pair_chunk(const double2 *ps, const int32_t *pop) {
    PairChunk c;
    const double2 v = __ldg(ps);
    c.pos = v.x; c.size = v.y; c.pop = __ldg(pop);
    return c;
}

__device__ __forceinline__ PairEval eval_pair(bool in_range, uint2 r, int t1, unsigned inv1, int t2, unsigned inv2,
                                              const BinRef binA, const BinRef binB, const EvalArgs &a, int slab0, int &bad) {
    PairEval o;
    o.cell = -1; o.local = 0; o.slot = -1; o.val = 0.0;
    if (!in_range) return o;
    // reference C:136-137: chunk1 = start_bin[j] + rand() % t_1 ; chunk2 = start_bin[k] + rand() % t_2
    const unsigned ia = fastmod_u32(r.x, (unsigned)t1, inv1), ib = fastmod_u32(r.y, (unsigned)t2, inv2);
    const PairChunk A = ldg_pair_chunk(binA.ps + ia, binA.pop + ia), B = ldg_pair_chunk(binB.ps + ib, binB.pop + ib);
    bool live = true;
    if (a.mode == 1) live = (A.pop != 0) && (B.pop != 0);                  // reference C:138-142
    else if (A.pop == 0 || B.pop == 0) { bad |= kErrMode3Target; live = false; }
    if (!live) return o;
    const double dist = fabs(__dsub_rn(A.pos, B.pos));                      // reference C:143
    // dist / bw, correctly rounded.  With y = RN(1/bw): q = RN(dist*y), r = dist - bw*q (exact, one FMA),
    // RN(q + r*y) = RN(dist/bw) whenever the significand of bw is not all ones (Markstein); the host checks that and
    // passes inv_bw = 0 otherwise.  Four fp64 instructions instead of the ~25 of the generic division.
    double qd;
    if (a.inv_bw != 0.0) {
        const double q0 = __dmul_rn(dist, a.inv_bw);
        qd = __fma_rn(__fma_rn(-q0, a.bw, dist), a.inv_bw, q0);
    } else qd = __ddiv_rn(dist, a.bw);
    const double fl = floor(qd);
    const double frac = __dmul_rn(__dsub_rn(qd, fl), a.bw);                 // reference C:144-145
    int bfull = __double2int_rz(fl);
    bool ok = true;
    if (frac <= a.half_bw) {}
    else if (frac > a.half_bw) bfull += 1;
    else ok = false;
    const int bb = bfull - a.grid_start;
    const double mind = __dadd_rn(__dmul_rn(A.size, 0.5), __dmul_rn(B.size, 0.5));
    if (ok && (unsigned)bb < (unsigned)a.G && dist >= mind) {
        const int l1 = A.pop - 1, l2 = B.pop - 1;
        const int lo_l = l1 < l2 ? l1 : l2, hi_l = l1 < l2 ? l2 : l1;
        const int row = tri_row(lo_l, hi_l, a.K);
        const int slab = a.slab_magic ? (int)__umulhi((unsigned)bb, a.slab_magic) : bb;   // bb / slab_width (exact: both < 2^16; 0 = width 1)
        const int slot = slab - slab0;
        if (!a.relaxed && (unsigned)slot >= (unsigned)a.n_slots) { bad |= kErrInternal; return o; }
        o.cell = row * a.G + bb;
        const int sb0 = slab * a.slab_width;
        const int sw = min(a.slab_width, a.G - sb0);                          // the last slab may be narrower
        o.local = row * sw + (bb - sb0);
        o.slot = slot;
        // gridweight = min(size, cutoff), as stored by the chunk builder (reference C:527)
        o.val = __dmul_rn(A.size > a.cutoff ? a.cutoff : A.size, B.size > a.cutoff ? a.cutoff : B.size);   // reference C:527 (sizes are never NaN)
    }
    return o;
}

// ---- glibc rand() inside phase 1 (rand_impl = 1) ---------------------------------------------------------------
// A warp consumes one contiguous piece of the stream (its separations of row (n, j) follow each other), so it carries
// its own generator: lanes 0..30 hold the last 31 raw words x[m-31..m-1]; the next 31 follow from
//     x[m+i] = x[m+i-31] + x[m+i-3]  =  hist[i] + (i < 3 ? hist[i+28] : x[m+i-3])
// i.e. three interleaved running sums (i mod 3): one shuffle for the three wrapped terms and a stride-3 scan
// (shuffle-up by 3, 6, 12, 24).  Lane 31 idles.  The warp reaches its start by jump-ahead on the device: the state
// D raws after the call's entry state is  sum_i c_i * Y[i+j]  with c = E^D mod (E^31 - E^28 - 1), taken one base-256
// digit of D at a time from a table of E^(digit * 256^level) (built on the host by glibc_rand.h).
constexpr int kRingRaws = 256;     // per-warp ring of generated draws (uint32, already >> 1)
constexpr int kJumpLevels = 6;     // base-256 digits of the raw offset: 2^48 raws = 1.4e14 sampled pairs per call

__device__ __forceinline__ uint32_t gen_block31(uint32_t hist, int lane) {
    const uint32_t t = __shfl_sync(kFull, hist, (lane + 28) & 31);
    uint32_t a = hist + (lane < 3 ? t : 0u);
    // shfl.up hands back its own in-range predicate: one shuffle and one predicated add per level
#pragma unroll
    for (int s = 3; s <= 24; s <<= 1)
        asm volatile("{\n.reg .pred p;\n.reg .b32 u;\nshfl.sync.up.b32 u|p, %0, %1, 0, 0xffffffff;\n@p add.u32 %0, %0, u;\n}\n" : "+r"(a) : "r"(s));
    return a;
}

// hist <- history D raws later.  scratch: >= 96 words of the warp's shared memory.
__device__ __forceinline__ void warp_jump(uint32_t &hist, unsigned long long D, const uint32_t *__restrict__ tab,
                                          uint32_t *scratch, int lane) {
    for (int lv = 0; D != 0ull && lv < kJumpLevels; lv++, D >>= 8) {
        const int digit = (int)(D & 255ull);
        if (!digit) continue;
        const uint32_t ext = gen_block31(hist, lane);              // Y[31 + lane]
        __syncwarp();
        if (lane < 31) { scratch[lane] = hist; scratch[31 + lane] = ext; }
        scratch[64 + lane] = tab[((size_t)lv * 256 + digit) * 32 + lane];
        __syncwarp();
        const int j = lane < 31 ? lane : 0;
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < 31; i++) acc += scratch[64 + i] * scratch[i + j];
        hist = acc;
        __syncwarp();
    }
}

template <bool FUSED_RNG, bool RELAXED>
__global__ void __launch_bounds__(128, 8) k_eval_pairs(const EvalArgs a) {
    __shared__ __align__(16) uint32_t ring_all[FUSED_RNG ? 4 * kRingRaws : 4];
    const int lane = lane_id();
    const ChromPrep &cp = a.cp;
    const int nb = cp.nb, W = cp.W;
    const int DS = a.d_parts;
    const long long wtask = (long long)blockIdx.x * (blockDim.x >> 5) + warp_id();
    if (wtask >= (long long)a.N * (nb - 1) * DS) return;
    const long long task = wtask / DS;
    const int part = (int)(wtask - task * DS);
    const int n = (int)(task / (nb - 1));
    const int j = (int)(task - (long long)n * (nb - 1));
    const int phys = a.phys_of[n];
    const int32_t *T = cp.t + (size_t)phys * nb;
    const int t1 = T[j];
    if (t1 <= 0) return;
    const int32_t *F = cp.first + (size_t)phys * nb;
    const int32_t *yo = cp.yoff + ((size_t)phys * nb + j) * W;
    const double2 *chps = cp.ps_ptr[phys];
    const int32_t *chpop = cp.pop_ptr[phys];
    const long long rowoff_j = cp.rowoff[(size_t)phys * (nb + 1) + j];
    const long long row_pair0 = a.pair_base[n] + rowoff_j;        // stream position (pairs) of the row's first draw
    UpdateRec *const row_recs = RELAXED ? nullptr : a.recs + (a.rec_base[n] + rowoff_j) * a.n_slots;   // the row's record regions
    const int f1 = F[j];
    const unsigned inv1 = mod_magic((unsigned)t1);
    const int ns = a.n_slots;
    const BinRef binA{chps + f1, chpop + f1};
    int dmax = W - 1;
    if (nb - 1 - j < dmax) dmax = nb - 1 - j;
    int bad = 0;
    const unsigned lt = (1u << lane) - 1;
    // this warp's share of the separations: d in (d_from, d_to], cut where the row's cumulative pair count crosses
    // part/DS of its total (small ranges of individuals would otherwise leave most of the GPU idle)
    int d_from = 0, d_to = dmax;
    if (DS > 1) {
        const int y00 = yo[0];
        const long long yrow = (long long)yo[dmax] - y00;
        const long long thr_lo = yrow * part / DS, thr_hi = yrow * (part + 1) / DS;
        int n_lo = 0, n_hi = 0;                   // separations d in [0, dmax] whose cumulative count is <= the threshold
        for (int d0 = 0; d0 <= dmax; d0 += 32) {
            const long long c = (d0 + lane <= dmax) ? (long long)yo[d0 + lane] - y00 : (1LL << 62);
            n_lo += __popc(__ballot_sync(kFull, c <= thr_lo));
            n_hi += __popc(__ballot_sync(kFull, c <= thr_hi));
        }
        d_from = n_lo - 1;
        d_to = (part == DS - 1) ? dmax : n_hi - 1;
        if (d_from >= d_to) return;
    }
    // The row's draws are consumed strictly front to back (segment d+1 starts where segment d ends).
    //   bulk stream (k_rand_fill): the 64 draws of the next iteration are fetched while the current one waits for
    //   its chunk gathers;  fused: the warp generates them into its ring as it goes.
    uint2 nr0 = make_uint2(0u, 0u), nr1 = make_uint2(0u, 0u);
    auto prefetch = [&](long long p) {
        nr0 = a.draws[p + lane];                         // may run up to 127 pairs past the batch: the buffer has kDrawSlack
        nr1 = a.draws[p + 32 + lane];
    };
    uint32_t *ring = ring_all + (FUSED_RNG ? warp_id() * kRingRaws : 0);
    uint32_t hist = 0;
    unsigned ring_w = 0, ring_r = 0;                     // raws generated / consumed so far (uniform)
    if (FUSED_RNG) {
        if (lane < 31) hist = a.hist0[lane];
        warp_jump(hist, 2ull * (unsigned long long)(row_pair0 + yo[d_from]), a.jump_tab, ring, lane);
    } else prefetch(row_pair0 + yo[d_from]);
    double *tens = RELAXED ? a.tensor + (size_t)n * ((size_t)a.K * (a.K + 1) / 2) * a.G : nullptr;
    unsigned long long n_relaxed = 0;
    for (int d = d_from + 1; d <= d_to; d++) {
        const int y0 = yo[d - 1];
        const int Y = yo[d] - y0;
        if (Y <= 0) continue;
        const int k = j + d;
        const int t2 = T[k], f2 = F[k];
        const unsigned inv2 = mod_magic((unsigned)t2);
        const long long seg = row_pair0 + y0;
        UpdateRec *reg = RELAXED ? nullptr : row_recs + (long long)y0 * ns;
        const int slab0 = a.first_slab[d];
        const BinRef binB{chps + f2, chpop + f2};
        int c[kMaxSlots] = {};
        for (int base = 0; base < Y; base += 64) {
            const int q0 = base + lane, q1 = base + 32 + lane;
            uint2 r0, r1;
            if (FUSED_RNG) {
                const unsigned need = 2u * (unsigned)min(64, Y - base);
                while (ring_w - ring_r < need) {         // four blocks of 31 at a time: at most 127 + 124 draws are buffered
#pragma unroll
                    for (int blk = 0; blk < 4; blk++) {
                        hist = gen_block31(hist, lane);
                        if (lane < 31) ring[(ring_w + 31 * blk + lane) & (kRingRaws - 1)] = hist >> 1;      // rand() = x >> 1
                    }
                    ring_w += 124u;
                }
                __syncwarp();
                const uint2 *r64 = reinterpret_cast<const uint2 *>(ring);
                r0 = r64[((ring_r >> 1) + lane) & (kRingRaws / 2 - 1)];
                r1 = r64[((ring_r >> 1) + 32 + lane) & (kRingRaws / 2 - 1)];
                ring_r += need;
                __syncwarp();
            } else {
                r0 = nr0; r1 = nr1;
                prefetch(seg + min(base + 64, Y));
            }
            PairEval e[2];
            e[0] = eval_pair(q0 < Y, r0, t1, inv1, t2, inv2, binA, binB, a, slab0, bad);
            e[1] = eval_pair(q1 < Y, r1, t1, inv1, t2, inv2, binA, binB, a, slab0, bad);
            if (RELAXED) {
                // relaxed mode (opt-in, "strict" = 0): fp64 atomics straight into the tensor -- additions land in
                // arbitrary order, so cells agree with the reference to rounding (~1e-16 relative), not bit for bit
#pragma unroll
                for (int h = 0; h < 2; h++)
                    if (e[h].cell >= 0) { atomicAdd(tens + e[h].cell, e[h].val); n_relaxed++; }
            } else {
                // ordered partition of the two steps' updates into the segment's slab regions
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    // three ballots (accepted, slot bit 0, slot bit 1) give the lane mask of every slot: no branch on the
                    // number of slots in use (slots nobody holds have empty masks)
                    static_assert(kMaxSlots == 4, "two slot bits");
                    const int sl = e[h].slot;
                    const unsigned bv = __ballot_sync(kFull, sl >= 0);
                    const unsigned b0 = __ballot_sync(kFull, (sl & 1) != 0);
                    const unsigned b1 = __ballot_sync(kFull, (sl & 2) != 0);
                    const unsigned lo = bv & ~b1, hi = bv & b1;
                    const unsigned m0 = lo & ~b0, m1 = lo & b0, m2 = hi & ~b0, m3 = hi & b0;
                    const bool s0 = (sl & 1) != 0, s1 = (sl & 2) != 0;
                    const unsigned mine = s1 ? (s0 ? m3 : m2) : (s0 ? m1 : m0);
                    const int cs = s1 ? (s0 ? c[3] : c[2]) : (s0 ? c[1] : c[0]);
                    const int pos = cs + __popc(mine & lt);
                    c[0] += __popc(m0); c[1] += __popc(m1); c[2] += __popc(m2); c[3] += __popc(m3);
                    if (sl >= 0) {
                        // one 16-byte store per record
                        const long long vb = __double_as_longlong(e[h].val);
                        *reinterpret_cast<int4 *>(&reg[(long long)sl * Y + pos]) =
                            make_int4(e[h].cell, e[h].local, (int)(vb & 0xffffffffLL), (int)(vb >> 32));
                    }
                }
            }
        }
        if (!RELAXED && lane < ns) {
            int v = c[0];
#pragma unroll
            for (int s = 1; s < kMaxSlots; s++) if (lane == s) v = c[s];
            const int slab = slab0 + lane;
            if (v > 0 && slab < a.n_slabs) {
                // region directory of task (n, slab): entry (j, d - dlo(slab)); empty entries stay zero
                RegionEnt ent;
                ent.base = (long long)(reg - a.recs) + (long long)lane * Y;
                ent.cnt = v; ent.pad = 0;
                a.regions[(((size_t)n * a.n_slabs + slab) * (nb - 1) + j) * a.nd_max + (d - a.slabs[slab].dlo)] = ent;
            }
        }
    }
    if (RELAXED) {
        n_relaxed = warp_sum_u64(n_relaxed);
        if (lane == 0 && n_relaxed) atomicAdd(a.accepted, n_relaxed);
    }
    if (bad) atomicOr(a.err, bad);
}

void launch_eval_pairs(const EvalArgs &a, cudaStream_t st) {
    long long tasks = (long long)a.N * (a.cp.nb - 1) * a.d_parts;
    if (tasks <= 0) return;
    const int wpb = 4;
    const unsigned grid = (unsigned)((tasks + wpb - 1) / wpb);
    const bool fused = a.jump_tab != nullptr;
    if (a.relaxed) {
        if (fused) k_eval_pairs<true, true><<<grid, wpb * 32, 0, st>>>(a);
        else k_eval_pairs<false, true><<<grid, wpb * 32, 0, st>>>(a);
    } else {
        if (fused) k_eval_pairs<true, false><<<grid, wpb * 32, 0, st>>>(a);
        else k_eval_pairs<false, false><<<grid, wpb * 32, 0, st>>>(a);
    }
}

// Phase 2: one CTA per (pseudo-individual, slab) folds the slab's records in sequence order
// (j ascending, k ascending, draw ascending): the "+=" of reference C:155 / C:161.
//
// Warp-specialised, Blackwell style: warp 1 is the producer -- it walks the record regions of the
// slab in order and streams them into a shared-memory ring with TMA bulk copies
// (cp.async.bulk + mbarrier complete_tx); warp 0 is the consumer -- it folds 32 records per step into
// the slab tile, which lives in shared memory when it fits (TILE_SMEM), so the ordered
// read-modify-write chain runs at shared-memory latency and global latency is off the critical path.
constexpr int kStages = 3;        // ring depth
constexpr int kStagesStepwise = 4; // ring depth of the step-wise variant (even: each set of consumers walks every other slot; 6 measured equal)
constexpr int kStageRecs = 128;   // records per ring stage (four consumer steps); stages are packed across regions
constexpr int kStageSteps = kStageRecs / 32;
constexpr int kTagSlots = 512;    // STEPWISE: slots of a tag array (two arrays per consumer warp, one byte per slot)
constexpr int kConsumers = 2;     // consumer warps per CTA; each owns the tile cells with local % kConsumers == warp

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// BUCKETS (shared-memory tiles only): the ordered fold without match.any.  A consumer warp files its records, in stream
// order, into 32 per-lane queues keyed by the low bits of the cell index (five ballots give every record its rank among
// the records of the same key in its 32-record step); lane l then folds queue l front to back.  Records of one cell always
// share a queue, so every cell still receives its addends in stream order, and different lanes never touch the same cell:
// no duplicate detection, no shuffle chains.  Queues are only folded when one is about to overflow (or the task ends), so
// a fold pass covers several hundred records.
constexpr int kQCap = 32;         // slots per lane queue (a single 32-record step always fits)
constexpr int kQueueBytes = kConsumers * (kQCap * 32 * (int)(sizeof(double) + sizeof(int32_t)) + 32 * (int)sizeof(int32_t));

// NC consumer warps per CTA (not STEPWISE).  With two, both warps read and match every record and fold only the cells they own
// (records of the other warp share one sentinel key); with one (the NULL path's default) no record is matched twice: the match.any
// unit is the binding resource of these forms, and the second warp's sentinel keys cost more than sharing the fold chains saves.
//
// STEPWISE (shared-memory tiles, NC = 4): two sets of two consumer warps; set q takes the stages k = q (mod 2) of the record
// stream, its warp h the steps 2h and 2h+1 (32 records each) of the stage.  Waiting for the stage, reading the records,
// releasing the stage and finding the lanes that share a cell run in four warps at once, a stage ahead of the folds; only
// the folds are serial -- the turn travels w0 -> w1 -> w2 -> w3 -> w0 through named barriers (one hand-over per 64 records), so
// the steps are folded in stream order.  With one consumer warp the whole per-stage chain (wait, loads, fence, four
// matches, four folds: ~2 000 cycles) is one warp's latency and the SM's match.any unit is the binding resource.
template <bool TILE_SMEM, bool BUCKETS, int NC = kConsumers, bool STEPWISE = false>
__global__ void __launch_bounds__(32 * (NC + 1)) k_commit_ordered(const CommitArgs a) {
    static_assert(!STEPWISE || (TILE_SMEM && !BUCKETS && NC == kStageSteps), "step-wise consumers: two sets of two warps");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int RS = STEPWISE ? kStagesStepwise : kStages;    // ring depth
    static_assert(kStagesStepwise % 2 == 0, "each set of step-wise consumers walks every other ring slot");
    constexpr int kRing = RS * kStageRecs * (int)sizeof(UpdateRec);
    UpdateRec *ring = reinterpret_cast<UpdateRec *>(smem_raw);                         // [RS][kStageRecs]
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + kRing);
    uint64_t *empty = full + RS;
    int *stage_cnt = reinterpret_cast<int *>(empty + RS);                              // [RS]
    int *task_slot = stage_cnt + RS;                                                   // [1] task broadcast
    unsigned char *qbase = smem_raw + kRing + 1024;                                    // ring + 1 KB control (barriers, counts), then the queues
    static_assert(2 * RS * 8 + RS * 4 + 4 <= 256, "control block layout");
    double *tile = reinterpret_cast<double *>(qbase + (BUCKETS ? kQueueBytes : 0));
    volatile unsigned char *tags = (STEPWISE && a.tag_off > 0) ? reinterpret_cast<unsigned char *>(tile) + a.tag_off : nullptr;
    const int lane = lane_id();
    const int wid = warp_id();
    const int G = a.G;
    const int P = a.K * (a.K + 1) / 2;
    const int RB = a.row_blocks > 0 ? a.row_blocks : 1;      // row blocks per slab tile (large K: records routed per block first)
    const bool tabbed = a.task_tab != nullptr;               // explicit task table (NULL path): one tile and one record region per entry
    const int total_tasks = tabbed ? a.n_task_tab : a.N * a.n_slabs * RB;
    if (threadIdx.x == 0) {
        for (int s = 0; s < RS; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], STEPWISE ? 2 : NC); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    // ring position: carried across tasks (the end marker of a task occupies a stage like any other)
    int stage = (STEPWISE && warp_id() < NC) ? warp_id() >> 1 : 0;   // STEPWISE consumers start at their set's first stage
    unsigned phase = 0;
    unsigned long long n_acc = 0;
    const unsigned lt = (1u << lane) - 1;
    bool first_turn = warp_id() == 0;                        // STEPWISE: the very first turn (warp 0's) waits for nobody

    // Persistent CTA: tasks (heavy slabs first) are claimed from a global counter, so the grid never has
    // blocks waiting and kernels of other streams (the next range's evaluation) can share the SMs.
    while (true) {
        __syncthreads();                                   // previous task fully retired (tile stored, ring drained)
        if (threadIdx.x == 0) *task_slot = atomicAdd(a.task_counter, 1);
        __syncthreads();
        const int task = *task_slot;
        if (task >= total_tasks) break;
        int n = 0, slab_i = 0, rb = 0, row0, rows, sbw, boff;
        if (tabbed) {
            const CommitTile td = a.task_tab[task];         // heavy tasks first (host order)
            row0 = td.row0; rows = td.rows; sbw = td.bins; boff = td.b0;
        } else {
            const int sr = task / a.N;                      // (slab, row block), heavy slabs first
            n = task - sr * a.N;
            slab_i = sr / RB; rb = sr - slab_i * RB;
            const SlabDesc sl = a.slabs[slab_i];
            row0 = RB > 1 ? rb * a.rows_per_block : 0;      // the tile holds tensor rows [row0, row0 + rows)
            rows = RB > 1 ? min(a.rows_per_block, P - row0) : P;
            sbw = sl.b1 - sl.b0 + 1;
            boff = sl.b0 - a.grid_start;
        }
        double *tens = a.tensor + (size_t)n * (size_t)P * G + (size_t)row0 * G;
        if (TILE_SMEM) {
            if (a.tensor_is_zero) {
                for (int idx = threadIdx.x; idx < rows * sbw; idx += blockDim.x) tile[idx] = 0.0;
            } else {
                for (int idx = threadIdx.x; idx < rows * sbw; idx += blockDim.x) {
                    const int row = idx / sbw, off = idx - row * sbw;
                    tile[idx] = tens[(size_t)row * G + boff + off];
                }
            }
        }
        __syncthreads();

        if (wid == NC) {
            // ------------------------------ producer ------------------------------
            // routed records (row blocks): one contiguous region per task; otherwise the slab's (j, d) region directory
            const long long n_ent = (RB > 1 || tabbed) ? 1 : (long long)(a.cp.nb - 1) * a.nd_max;
            const RegionEnt *ents = tabbed ? a.regions + task : a.regions + (((size_t)n * a.n_slabs + slab_i) * RB + rb) * n_ent;
            // All 32 lanes work: lane l holds entry e0 + l; a prefix sum over the batch's counts gives every region its place in
            // the task's record stream, and for every stage the batch touches each lane issues its own piece (at most one per
            // stage and lane).  Lane 0 alone opens a stage (waits for it to be empty) and closes it (count, arrive).
            // [one lane walking the regions one by one -- ~130 dependent instructions per stage -- was what bounded a CTA
            //  once the consumers had become fast]
            RegionEnt cur, nxt;
            cur.base = 0; cur.cnt = 0; cur.pad = 0;
            if (lane < n_ent) cur = ents[lane];
            int fill = 0;                                  // records already placed in the open stage
            bool open = false;
            for (long long e0 = 0; e0 < n_ent; e0 += 32) {
                nxt.base = 0; nxt.cnt = 0; nxt.pad = 0;
                if (e0 + 32 + lane < n_ent) nxt = ents[e0 + 32 + lane];     // in flight while this batch is issued
                const long long c = cur.cnt > 0 ? (long long)cur.cnt : 0;
                long long incl = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const long long t = __shfl_up_sync(kFull, incl, o);
                    if (lane >= o) incl += t;
                }
                const long long total = __shfl_sync(kFull, incl, 31);
                const long long start = incl - c;          // this lane's region inside the batch: [start, start + c)
                const UpdateRec *src = a.recs + cur.base;
                long long done = 0;                        // records of the batch already placed
                while (done < total) {
                    if (!open) {
                        if (lane == 0) { mbar_wait(&empty[stage], phase ^ 1); fence_proxy_async_smem(); }
                        __syncwarp();
                        open = true; fill = 0;
                    }
                    const long long take = min((long long)(kStageRecs - fill), total - done);   // records this stage takes now
                    const long long lo = max(start, done), hi = min(start + c, done + take);
                    if (hi > lo) {
                        const unsigned bytes = (unsigned)(hi - lo) * (unsigned)sizeof(UpdateRec);
                        mbar_expect_tx(&full[stage], bytes);
                        bulk_g2s(&ring[stage * kStageRecs + fill + (int)(lo - done)], src + (lo - start), bytes, &full[stage]);
                    }
                    fill += (int)take; done += take;
                    __syncwarp();                          // every piece of the stage is announced before lane 0 arrives
                    if (fill == kStageRecs) {
                        if (lane == 0) { stage_cnt[stage] = fill; mbar_arrive(&full[stage]); }
                        open = false;
                        if (++stage == RS) { stage = 0; phase ^= 1; }
                    }
                }
                cur = nxt;
            }
            if (open) {                                    // last, partially filled stage
                if (lane == 0) { stage_cnt[stage] = fill; mbar_arrive(&full[stage]); }
                if (++stage == RS) { stage = 0; phase ^= 1; }
            }
            for (int mk = 0; mk < (STEPWISE ? 2 : 1); mk++) {   // end marker: occupies one stage (one per set of consumer warps)
                if (lane == 0) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    stage_cnt[stage] = -1;
                    mbar_arrive(&full[stage]);
                }
                if (++stage == RS) { stage = 0; phase ^= 1; }
            }
            __syncwarp();
            stage = __shfl_sync(kFull, stage, 0);
            phase = __shfl_sync(kFull, phase, 0);
        } else {
            // ------------------------------ consumer ------------------------------
            // software pipeline: the records of stage s+1 are fetched (and its ring slot released) before stage s is
            // folded, so barrier / shared-memory latency overlaps the ordered read-modify-write work.
            // Several consumer warps read every stage; each handles only the cells it owns, so a cell's additions
            // stay in stream order.
            struct Staged { int c; int cell[kStageSteps]; double val[kStageSteps]; unsigned g[kStageSteps]; };
            // ownership: BUCKETS -- alternate blocks of 32 tile cells (the low five bits pick the lane queue, so the 32
            // lanes of a fold pass hit 32 consecutive doubles: conflict free); otherwise odd / even cells
            auto owned = [&](int local) -> bool {
                if (NC == 1) return true;
                return BUCKETS ? (((unsigned)local >> 5) % NC == (unsigned)wid) : ((unsigned)local % NC == (unsigned)wid);
            };
            auto fetch = [&]() -> Staged {
                Staged s;
                mbar_wait(&full[stage], phase);
                s.c = stage_cnt[stage];
#pragma unroll
                for (int h = 0; h < kStageSteps; h++) { s.cell[h] = -1; s.val[h] = 0.0; s.g[h] = 0; }
                if (s.c >= 0) {
                    const UpdateRec *rg = ring + stage * kStageRecs;
#pragma unroll
                    for (int h = 0; h < kStageSteps; h++)
                        if (32 * h + lane < s.c) {
                            const UpdateRec r = rg[32 * h + lane];
                            if (owned(r.pad)) { s.cell[h] = TILE_SMEM ? r.pad : r.cell; s.val[h] = r.val; }
                        }
                }
                // The refill of this stage is a TMA write (async proxy); the reads above are generic-proxy reads.
                // Without the cross-proxy fence the bulk copy can land before these loads have sampled the
                // ring (seen on B200 as a few wrong cells whenever another kernel shares the GPU).
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[stage]);     // records are in registers: the stage can be refilled
                if (++stage == RS) { stage = 0; phase ^= 1; }
                if (!BUCKETS && s.c >= 0) {
                    // Lanes that share a cell must add in lane order, so every step needs, per lane, the set of lanes
                    // with the same cell: match.any (two cycles per distinct key on the SM's single unit; records of
                    // other warps share one sentinel key)
#pragma unroll
                    for (int h = 0; h < kStageSteps; h++)
                        if (32 * h < s.c) s.g[h] = __match_any_sync(kFull, s.cell[h] >= 0 ? (unsigned)s.cell[h] : 0x80000000u);
                }
                return s;
            };
            if (STEPWISE) {
                // Two sets of two warps: set q takes the stages k = q (mod 2) of the stream, its warp h the steps 2h and 2h+1 of
                // the stage.  The turn travels w0 -> w1 (stage k) -> w2 -> w3 (stage k+1) -> w0 ...: one hand-over per 64 records.
                // kind 0: 32 distinct cells; 1: cells shared by at most two lanes, g = the partner lane (own lane: none);
                // 2: g = match.any mask
                struct Step { int cell; double val; unsigned g; int kind; };
                struct Two { int c; Step s[2]; };
                const int half = wid & 1;
                volatile unsigned char *tA = tags ? tags + wid * 2 * kTagSlots : nullptr, *tB = tA + kTagSlots;
                auto classify = [&](Step &s) {
                    // Lanes that share a cell must add in lane order.  match.any would give every lane the lanes with its
                    // cell, at ~30 + 2 cycles per distinct key on the SM's single unit -- the binding resource of the earlier
                    // forms of this kernel.  Measured on the bench workload: 41 % of the steps have 32 distinct cells, 56 % have
                    // nothing worse than pairs, 3 % a cell shared by three or more lanes.  Two private tag arrays per warp (a
                    // byte per slot, slot = cell mod kTagSlots) sort the first two classes out without match.any:
                    //   round 1: every lane writes its number at its slot and re-reads: the last writer "wins" the slot;
                    //   round 2: the lanes that lost write their number into the second array and re-read: a loser that does
                    //            not find itself is one of several losers of its slot -> the step falls back to match.any.
                    // Otherwise every slot holds one lane or one (winner, loser) couple; the two find each other through the
                    // tags (the winner checks with a shuffle that its candidate really lost THIS step's slot -- the second
                    // array may hold an older step's number), and are a pair iff their cells (not only their slots) agree.
                    const bool valid = s.cell >= 0;
                    const int slot = s.cell & (kTagSlots - 1);
                    s.kind = 2;
                    if (tA) {
                        if (valid) tA[slot] = (unsigned char)lane;
                        __syncwarp();
                        const int t = valid ? tA[slot] : lane;
                        const bool lost = valid && t != lane;
                        if (!__any_sync(kFull, lost)) s.kind = 0;
                        else {
                            if (lost) tB[slot] = (unsigned char)lane;
                            __syncwarp();
                            const int u = valid ? tB[slot] : lane;
                            const bool bad = lost && u != lane;
                            const int cand = (lost ? t : u) & 31;
                            const int c_cell = __shfl_sync(kFull, s.cell, cand);
                            const int c_lost = __shfl_sync(kFull, (int)lost, cand);
                            bool paired;
                            if (lost) paired = c_cell == s.cell;              // the winner of my slot is this step's by construction
                            else paired = valid && c_lost && c_cell == s.cell;  // my candidate lost this step, to my very cell
                            if (!__any_sync(kFull, bad)) { s.kind = 1; s.g = (unsigned)(paired ? cand : lane); }
                        }
                        __syncwarp();                              // the arrays are free for the next step
                    }
                    if (s.kind == 2) s.g = __match_any_sync(kFull, (unsigned)s.cell);
                };
                auto fetch2 = [&]() -> Two {
                    Two t;
                    mbar_wait(&full[stage], phase);
                    t.c = stage_cnt[stage];
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        t.s[h].cell = -1; t.s[h].val = 0.0; t.s[h].g = 0; t.s[h].kind = 0;
                        const int idx = 32 * (2 * half + h) + lane;
                        if (idx < t.c) {
                            const UpdateRec r = ring[stage * kStageRecs + idx];
                            t.s[h].cell = r.pad; t.s[h].val = r.val;
                        }
                    }
                    fence_proxy_async_smem();                  // generic-proxy reads before the async-proxy refill (see fetch() of the other consumer forms)
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[stage]);
                    stage += 2;                                // the next stage of this set
                    if (stage >= RS) { stage -= RS; phase ^= 1; }
#pragma unroll
                    for (int h = 0; h < 2; h++)
                        if (32 * (2 * half + h) < t.c) classify(t.s[h]);
                    return t;
                };
                auto fold1 = [&](const Step &s, double pv) {
                    if (s.kind == 0) {
                        if (s.cell >= 0) tile[s.cell] = __dadd_rn(tile[s.cell], s.val);
                    } else if (s.kind == 1) {
                        const int partner = (int)s.g;
                        if (s.cell >= 0 && partner >= lane) {              // the lower lane of a couple adds both, in lane order
                            double acc = __dadd_rn(tile[s.cell], s.val);
                            if (partner != lane) acc = __dadd_rn(acc, pv);
                            tile[s.cell] = acc;
                        }
                    } else {
                        const bool leader = (s.cell >= 0) && ((s.g & lt) == 0);
                        double acc = leader ? tile[s.cell] : 0.0;
                        unsigned rest = leader ? (s.g & ~(1u << lane)) : 0u;
                        acc = __dadd_rn(acc, s.val);
                        while (__any_sync(kFull, rest != 0)) {
                            const int src = rest ? (__ffs(rest) - 1) : lane;
                            const double v = __shfl_sync(kFull, s.val, src);
                            if (rest) { acc = __dadd_rn(acc, v); rest &= rest - 1; }
                        }
                        if (leader) tile[s.cell] = acc;
                    }
                };
                const int bar_wait = 1 + wid, bar_next = 1 + ((wid + 1) & (kStageSteps - 1));
                Two cur = fetch2();
                while (cur.c >= 0) {
                    const Two nxt = fetch2();
                    // a couple's other addend, fetched before the turn
                    const double pv0 = __shfl_sync(kFull, cur.s[0].val, cur.s[0].kind == 1 ? (int)cur.s[0].g : lane);
                    const double pv1 = __shfl_sync(kFull, cur.s[1].val, cur.s[1].kind == 1 ? (int)cur.s[1].g : lane);
                    // the turn: named barrier 1 + w counts this warp's 32 threads and the 32 of its predecessor's bar.arrive
                    // (a hardware barrier hand-over; an mbarrier token costs about twice the latency, polling a flag far more)
                    if (!first_turn) asm volatile("bar.sync %0, 64;\n" ::"r"(bar_wait) : "memory");
                    first_turn = false;
                    if (32 * (2 * half) < cur.c) fold1(cur.s[0], pv0);
                    if (32 * (2 * half + 1) < cur.c) { __syncwarp(); fold1(cur.s[1], pv1); }
                    asm volatile("bar.arrive %0, 64;\n" ::"r"(bar_next) : "memory");
                    if (half == 0) n_acc += (unsigned long long)cur.c;
                    cur = nxt;
                }
            } else if (BUCKETS) {
                double *qval = reinterpret_cast<double *>(qbase) + (size_t)wid * kQCap * 32;                         // [slot][lane]
                int32_t *qloc = reinterpret_cast<int32_t *>(qbase + kConsumers * kQCap * 32 * sizeof(double)) + (size_t)wid * kQCap * 32;
                int32_t *qcnt = reinterpret_cast<int32_t *>(qbase + kConsumers * kQCap * 32 * (sizeof(double) + sizeof(int32_t))) + wid * 32;
                qcnt[lane] = 0;
                __syncwarp();
                auto flush = [&]() {                      // every lane folds its own queue, front to back
                    __syncwarp();
                    const int nq = qcnt[lane];
                    const int maxq = __reduce_max_sync(kFull, nq);
                    for (int i = 0; i < maxq; i++) {
                        if (i < nq) {
                            const int loc = qloc[i * 32 + lane];
                            tile[loc] = __dadd_rn(tile[loc], qval[i * 32 + lane]);
                        }
                    }
                    qcnt[lane] = 0;
                    __syncwarp();
                };
                auto insert = [&](int loc, double val) {  // one 32-record step: append every owned record to the queue of its key
                    const bool valid = loc >= 0;
                    const unsigned key = (unsigned)loc & 31u;
                    unsigned M = __ballot_sync(kFull, valid);
                    if (M == 0u) return;
#pragma unroll
                    for (int t = 0; t < 5; t++) {
                        const bool bit = (key >> t) & 1u;
                        const unsigned bm = __ballot_sync(kFull, valid && bit);
                        M &= bit ? bm : ~bm;
                    }
                    const int rank = __popc(M & lt);     // records of the same key ahead of this one in the step
                    int base = valid ? qcnt[key] : 0;
                    if (__any_sync(kFull, valid && base + rank >= kQCap)) { flush(); base = 0; }
                    if (valid) {
                        const int slot = base + rank;
                        qloc[slot * 32 + key] = loc;
                        qval[slot * 32 + key] = val;
                        if ((M >> lane) == 1u) qcnt[key] = slot + 1;        // the last record of its key in this step
                    }
                    __syncwarp();
                };
                Staged cur = fetch();
                while (cur.c >= 0) {
                    const Staged nxt = fetch();
#pragma unroll
                    for (int h = 0; h < kStageSteps; h++)
                        if (32 * h < cur.c) insert(cur.cell[h], cur.val[h]);
                    if (wid == 0) n_acc += (unsigned long long)cur.c;
                    cur = nxt;
                }
                flush();
            } else {
                auto fold = [&](int cell, double val, unsigned grp) {
                    const bool leader = (cell >= 0) && ((grp & lt) == 0);
                    double acc = 0.0;
                    if (leader) acc = TILE_SMEM ? tile[cell] : tens[cell];
                    unsigned rest = leader ? (grp & ~(1u << lane)) : 0u;
                    acc = __dadd_rn(acc, val);
                    while (__any_sync(kFull, rest != 0)) {
                        const int src = rest ? (__ffs(rest) - 1) : lane;
                        const double v = __shfl_sync(kFull, val, src);
                        if (rest) { acc = __dadd_rn(acc, v); rest &= rest - 1; }
                    }
                    if (leader) { if (TILE_SMEM) tile[cell] = acc; else tens[cell] = acc; }
                    __syncwarp();
                };
                Staged cur = fetch();
                while (cur.c >= 0) {
                    const Staged nxt = fetch();
#pragma unroll
                    for (int h = 0; h < kStageSteps; h++)
                        if (32 * h < cur.c) fold(cur.cell[h], cur.val[h], cur.g[h]);
                    if (wid == 0) n_acc += (unsigned long long)cur.c;
                    cur = nxt;
                }
            }
        }
        __syncthreads();
        if (TILE_SMEM) {
            for (int idx = threadIdx.x; idx < rows * sbw; idx += blockDim.x) {
                const int row = idx / sbw, o2 = idx - row * sbw;
                tens[(size_t)row * G + boff + o2] = tile[idx];
            }
        }
    }
    if ((STEPWISE ? (wid < NC && (wid & 1) == 0) : wid == 0) && lane == 0 && n_acc) atomicAdd(a.accepted, n_acc);
}

size_t commit_tile_bytes(int K, int slab_width) { return (size_t)K * (K + 1) / 2 * slab_width * sizeof(double); }

void launch_commit_ordered(const CommitArgs &a, int max_slab_width, cudaStream_t st) {
    const int RB = a.row_blocks > 0 ? a.row_blocks : 1;
    long long tasks = a.task_tab ? (long long)a.n_task_tab : (long long)a.N * a.n_slabs * RB;
    if (tasks <= 0) return;
    const size_t ctrl_bytes = 6144 + 1024;
    static_assert(kStages * kStageRecs * sizeof(UpdateRec) == 6144, "ring size");
    const size_t tile_bytes = a.task_tab ? (size_t)a.max_tile_cells * sizeof(double)
                              : RB > 1 ? (size_t)a.rows_per_block * max_slab_width * sizeof(double) : commit_tile_bytes(a.K, max_slab_width);
    static bool attr_set[64] = {};                    // the attribute is per device (one engine per device)
    int dev = 0;
    cudaGetDevice(&dev);
    if (tile_bytes + ctrl_bytes + 6144 <= 200 * 1024) {
        if (!attr_set[dev & 63]) {
            cudaFuncSetAttribute(k_commit_ordered<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            cudaFuncSetAttribute(k_commit_ordered<true, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            cudaFuncSetAttribute(k_commit_ordered<true, false, kStageSteps, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            cudaFuncSetAttribute(k_commit_ordered<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + kQueueBytes);
            attr_set[dev & 63] = true;
        }
        const bool buckets = a.buckets != 0;
        const size_t smem = ctrl_bytes + tile_bytes + (buckets ? kQueueBytes : 0);
        long long per_sm = std::max<long long>(1, (227 * 1024) / (long long)(smem + 1024));
        if (a.max_ctas_per_sm > 0) per_sm = std::min<long long>(per_sm, a.max_ctas_per_sm);
        const long long grid = std::min<long long>(tasks, 148 * std::min<long long>(per_sm, 32));
        if (buckets) k_commit_ordered<true, true><<<(unsigned)grid, 32 * (kConsumers + 1), smem, st>>>(a);
        else if (a.consumers == kStageSteps) {
            CommitArgs b = a;
            const size_t ctrl4 = (size_t)kStagesStepwise * kStageRecs * sizeof(UpdateRec) + 1024;
            size_t smem4 = ctrl4 + tile_bytes;
            b.tag_off = 0;
            if (a.tags) {                                   // the consumer warps' tag arrays behind the tile
                b.tag_off = (int32_t)((tile_bytes + 15) & ~(size_t)15);
                smem4 = ctrl4 + (size_t)b.tag_off + (size_t)kStageSteps * 2 * kTagSlots;
                if (smem4 > 200 * 1024) { b.tag_off = 0; smem4 = ctrl4 + tile_bytes; }
            }
            long long per_sm4 = std::min<long long>(64 / (kStageSteps + 1), std::max<long long>(1, (227 * 1024) / (long long)(smem4 + 1024)));
            if (a.max_ctas_per_sm > 0) per_sm4 = std::min<long long>(per_sm4, a.max_ctas_per_sm);
            const long long grid4 = std::min<long long>(tasks, 148 * per_sm4);
            k_commit_ordered<true, false, kStageSteps, true><<<(unsigned)grid4, 32 * (kStageSteps + 1), smem4, st>>>(b);
        } else if (a.consumers == 1) {
            const long long per_sm1 = a.max_ctas_per_sm > 0 ? std::min<long long>(a.max_ctas_per_sm, 32) : std::min<long long>(32, std::max<long long>(1, (227 * 1024) / (long long)(smem + 1024)));
            const long long grid1 = std::min<long long>(tasks, 148 * per_sm1);
            k_commit_ordered<true, false, 1><<<(unsigned)grid1, 64, smem, st>>>(a);
        } else k_commit_ordered<true, false><<<(unsigned)grid, 32 * (kConsumers + 1), smem, st>>>(a);
    } else {
        const long long grid = std::min<long long>(tasks, 148 * 16);
        k_commit_ordered<false, false><<<(unsigned)grid, 32 * (kConsumers + 1), ctrl_bytes, st>>>(a);
    }
}


// ---- large K: route the records of every (individual, slab) to row blocks ---------------------------------------------
// A slab tile of K(K+1)/2 rows does not fit shared memory for K >~ 70.  Instead of folding into the global tensor (one DRAM
// read-modify-write per record on a latency-bound chain), the slab's record stream is split, in stream order, by BLOCK OF
// TENSOR ROWS: a counting pass, a scan, and a stable scatter that rewrites every record's tile index relative to its row
// block.  Phase 2 then runs its shared-memory path on (individual, slab, row block) tasks whose tiles fit.
// One warp per (individual, slab, chunk of 32 coarse bins j) walks its part of the slab's region directory in order, 32
// records per step; counts are kept per (individual, slab, row block, j chunk) in exactly the order the routed buffer is
// laid out, so one exclusive scan gives every warp of the scatter pass its own write positions and the passes are fully
// parallel while each row block's records stay in stream order.
constexpr int kRouteJ = 32;       // coarse bins per routing task

template <bool SCATTER>
__global__ void __launch_bounds__(128)
k_route(const RouteArgs a) {
    __shared__ int32_t cnt_all[4][32];
    const int lane = lane_id(), wid = warp_id();
    const int JC = a.j_chunks;
    const long long task = (long long)blockIdx.x * 4 + wid;
    if (task >= (long long)a.N * a.n_slabs * JC) return;
    const int jc = (int)(task % JC);
    const long long ns_i = task / JC;
    const int n = (int)(ns_i / a.n_slabs), slab_i = (int)(ns_i - (long long)n * a.n_slabs);
    const int RB = a.row_blocks;
    int32_t *cnt = cnt_all[wid];
    cnt[lane] = 0;
    __syncwarp();
    const long long n_ent = (long long)(a.nb - 1) * a.nd_max;
    const RegionEnt *ents = a.regions + ((size_t)n * a.n_slabs + slab_i) * n_ent;
    const long long e_lo = (long long)jc * kRouteJ * a.nd_max, e_hi = min(n_ent, (long long)(jc + 1) * kRouteJ * a.nd_max);
    const SlabDesc sl = a.slabs[slab_i];
    const int sw = sl.b1 - sl.b0 + 1;
    const unsigned sw_magic = sw > 1 ? (unsigned)(((1ULL << 32) + sw - 1) / sw) : 0u;      // local / sw (local * sw < 2^32, checked on the host)
    const unsigned rpb_magic = a.rows_per_block > 1 ? (unsigned)(((1ULL << 32) + a.rows_per_block - 1) / a.rows_per_block) : 0u;
    const size_t tbase = ((size_t)n * a.n_slabs + slab_i) * RB;                              // counter (tbase + rb) * JC + jc
    const unsigned lt = (1u << lane) - 1;
    int32_t my_off = 0;
    if (SCATTER && lane < RB) my_off = a.offs[(tbase + lane) * JC + jc];
    for (long long e0 = e_lo; e0 < e_hi; e0 += 32) {
        RegionEnt ent;
        ent.base = 0; ent.cnt = 0; ent.pad = 0;
        if (e0 + lane < e_hi) ent = ents[e0 + lane];
        unsigned live = __ballot_sync(kFull, ent.cnt > 0);
        while (live) {
            const int i = __ffs(live) - 1;
            live &= live - 1;
            const long long base = __shfl_sync(kFull, ent.base, i);
            const int c = __shfl_sync(kFull, ent.cnt, i);
            for (int off = 0; off < c; off += 32) {
                const bool v = off + lane < c;
                UpdateRec r;
                r.cell = 0; r.pad = 0; r.val = 0.0;
                if (v) r = a.recs[base + off + lane];
                const unsigned row = sw_magic ? __umulhi((unsigned)r.pad, sw_magic) : (unsigned)r.pad;
                const unsigned rb = rpb_magic ? __umulhi(row, rpb_magic) : row;
                if (!SCATTER) {
                    if (v) atomicAdd(&cnt[rb], 1);
                } else {
                    // stable: rank among the step's records of the same row block, behind those this warp filed before
                    const unsigned m = __match_any_sync(kFull, v ? rb : 0xffffffffu);
                    const int rank = __popc(m & lt);
                    const int fill = v ? cnt[rb] : 0;
                    const int dst0 = __shfl_sync(kFull, my_off, v ? (int)rb : 0);
                    __syncwarp();
                    if (v) {
                        UpdateRec o;
                        o.cell = r.cell;
                        o.pad = (int)(((unsigned)row - rb * (unsigned)a.rows_per_block) * (unsigned)sw + ((unsigned)r.pad - row * (unsigned)sw));
                        o.val = r.val;
                        a.routed[(long long)dst0 + fill + rank] = o;
                        if ((m >> lane) == 1u) cnt[rb] = fill + rank + 1;
                    }
                    __syncwarp();
                }
            }
        }
    }
    __syncwarp();
    if (!SCATTER && lane < RB) a.counts[(tbase + lane) * JC + jc] = cnt[lane];
}

// one region per commit task (individual, slab, row block): the j chunks of a row block are adjacent in the routed buffer
__global__ void k_route_dir(const int32_t *__restrict__ offs, long long n_tasks, int JC, RegionEnt *__restrict__ dir) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tasks) return;
    RegionEnt o;
    o.base = offs[t * JC];
    o.cnt = offs[(t + 1) * JC] - offs[t * JC];
    o.pad = 0;
    dir[t] = o;
}

// exclusive scan of n int32 (n up to a few 10^6): per-block sums, scan of the sums, per-block scan
__global__ void __launch_bounds__(1024) k_scan_block_sums(const int32_t *__restrict__ in, long long n, int32_t *__restrict__ sums) {
    __shared__ int wtot[32];
    const long long i = (long long)blockIdx.x * 1024 + threadIdx.x;
    int v = i < n ? in[i] : 0;
    v = warp_sum(v);
    if (lane_id() == 0) wtot[warp_id()] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        int s = wtot[threadIdx.x];
        s = warp_sum(s);
        if (threadIdx.x == 0) sums[blockIdx.x] = s;
    }
}
__global__ void __launch_bounds__(1024) k_scan_apply(const int32_t *__restrict__ in, long long n, const int32_t *__restrict__ block_off,
                                                      int32_t *__restrict__ out) {
    __shared__ int wtot[32];
    const long long i = (long long)blockIdx.x * 1024 + threadIdx.x;
    const int v = i < n ? in[i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(kFull, x, o);
        if (lane_id() >= o) x += y;
    }
    if (lane_id() == 31) wtot[warp_id()] = x;
    __syncthreads();
    int base = block_off[blockIdx.x];
    for (int w = 0; w < warp_id(); w++) base += wtot[w];
    if (i < n) out[i] = base + x - v;
    if (i == n - 1) out[n] = base + x;
}

void launch_exclusive_scan_i32_large(const int32_t *in, int32_t *out, long long n, int32_t *scratch /*[2 * (n/1024 + 2)]*/, cudaStream_t st) {
    if (n <= 0) return;
    const long long nblk = (n + 1023) / 1024;
    int32_t *sums = scratch, *sums_off = scratch + nblk + 1;
    k_scan_block_sums<<<(unsigned)nblk, 1024, 0, st>>>(in, n, sums);
    k_exclusive_scan_i32<<<1, 1024, 0, st>>>(sums, sums_off, (int)nblk);
    k_scan_apply<<<(unsigned)nblk, 1024, 0, st>>>(in, n, sums_off, out);
}

void launch_route(const RouteArgs &a, bool scatter, cudaStream_t st) {
    const long long tasks = (long long)a.N * a.n_slabs * a.j_chunks;
    if (tasks <= 0) return;
    const unsigned grid = (unsigned)((tasks + 3) / 4);
    if (scatter) {
        k_route<true><<<grid, 128, 0, st>>>(a);
        const long long nt = (long long)a.N * a.n_slabs * a.row_blocks;
        k_route_dir<<<(unsigned)((nt + 255) / 256), 256, 0, st>>>(a.offs, nt, a.j_chunks, a.routed_dir);
    } else k_route<false><<<grid, 128, 0, st>>>(a);
}

// ---- sharded calls: stream offsets on the device -------------------------------------------------------------------
// all[r][0] = individuals of shard r, all[r][1 + n] = sampled pairs of its n-th individual on this chromosome (gathered
// by one all-gather).  The reference draws chromosome-major, individual-minor (C:338, C:419) and the shards own
// consecutive blocks of individuals, so shard `rank`'s n-th individual starts at
//     *stream_pos + (pairs of the shards before it) + (pairs of its own earlier individuals);
// *stream_pos then moves on by the chromosome's total.  One block.
__global__ void __launch_bounds__(256)
k_dist_offsets(const long long *__restrict__ all, int world, int rank, int n_max, long long *stream_pos, long long *__restrict__ pair_base) {
    __shared__ long long sh[2][8];
    const int stride = n_max + 1;
    long long before = 0, total = 0;
    for (long long i = threadIdx.x; i < (long long)world * stride; i += blockDim.x) {
        const int r = (int)(i / stride), k = (int)(i - (long long)r * stride);
        if (k == 0) continue;
        if (k - 1 >= (int)all[(size_t)r * stride]) continue;             // padding behind the shard's individuals
        const long long v = all[i];
        total += v;
        if (r < rank) before += v;
    }
    before = (long long)warp_sum_u64((unsigned long long)before);
    total = (long long)warp_sum_u64((unsigned long long)total);
    if (lane_id() == 0) { sh[0][warp_id()] = before; sh[1][warp_id()] = total; }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long b = 0, t = 0;
        for (int w = 0; w < 8; w++) { b += sh[0][w]; t += sh[1][w]; }
        const long long base = *stream_pos + b;
        const long long *mine = all + (size_t)rank * stride;
        const int n_local = (int)mine[0];
        long long run = 0;
        for (int n = 0; n < n_local; n++) { pair_base[n] = base + run; run += mine[1 + n]; }
        *stream_pos += t;
    }
}

void launch_dist_offsets(const long long *all, int world, int rank, int n_max, long long *stream_pos, long long *pair_base,
                         cudaStream_t st) {
    k_dist_offsets<<<1, 256, 0, st>>>(all, world, rank, n_max, stream_pos, pair_base);
}

// =============================================================================================
// 5. Projection onto surrogate pairs (reference C:578-595, mode 3: C:1190-1203):
//    results[n][mn][i] += sum over upper-triangle rows (k<=h, in that order) of
//    temp_weights[(K*k+h)*S*S + mn] * counts[n][(k,h)][i]; multiply and add rounded separately,
//    summed in the reference's order, so the result is bit-identical.
// =============================================================================================
constexpr int kMaxGridZ = 32768;                      // individuals per launch where they index grid.y / grid.z
constexpr int PT = 64, PKK = 16, PTM_MAX = 100;      // bins per tile, pairs per slice, most surrogate pairs per tile
__global__ void __launch_bounds__(16 * (PTM_MAX / 4))
k_project(const double *__restrict__ tensor, int K, int G, int S, const double *__restrict__ tw,
          const int32_t *__restrict__ pair_kh, double *__restrict__ results) {
    const int n = blockIdx.z;
    const int SS = S * S, P = K * (K + 1) / 2;
    const int TY = blockDim.x >> 4, PTM = 4 * TY;       // the tile is PTM surrogate pairs x PT bins; 4 x 4 outputs per thread
    const int i0 = blockIdx.x * PT, m0 = blockIdx.y * PTM;
    const double *B = tensor + (size_t)n * P * G;
    double *out = results + (size_t)n * SS * G;
    __shared__ double sA[PKK][PTM_MAX + 1];   // [pair][mn]
    __shared__ double sB[PKK][PT + 1];        // [pair][i]
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    double acc[4][4];
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
        for (int v = 0; v < 4; v++) {
            int mn = m0 + ty + TY * u, i = i0 + tx + 16 * v;
            acc[u][v] = (mn < SS && i < G) ? out[(size_t)mn * G + i] : 0.0;
        }
    // A slice is PKK x PTM = 4 * blockDim.x weights: thread t loads rows prA, prA + 4, prA + 8, prA + 12 of column cA (one
    // division per thread, outside the loop over slices)
    const int prA = threadIdx.x / PTM, cA = threadIdx.x - prA * PTM;
    const bool okA = m0 + cA < SS;
    const double *twA = tw + m0 + cA;
    for (int p0 = 0; p0 < P; p0 += PKK) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int p = p0 + prA + 4 * q;
            sA[prA + 4 * q][cA] = (p < P && okA) ? twA[(size_t)pair_kh[p] * SS] : 0.0;
        }
        for (int e = threadIdx.x; e < PKK * PT; e += blockDim.x) {
            const int pr = e / PT, c = e % PT;
            const int p = p0 + pr;
            sB[pr][c] = (p < P && i0 + c < G) ? B[(size_t)p * G + i0 + c] : 0.0;
        }
        __syncthreads();
        const int lim = (P - p0) < PKK ? (P - p0) : PKK;
        for (int pr = 0; pr < lim; pr++) {
            double av[4], bv[4];
#pragma unroll
            for (int u = 0; u < 4; u++) av[u] = sA[pr][ty + TY * u];
#pragma unroll
            for (int v = 0; v < 4; v++) bv[v] = sB[pr][tx + 16 * v];
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
                for (int v = 0; v < 4; v++) acc[u][v] = __dadd_rn(acc[u][v], __dmul_rn(av[u], bv[v]));
        }
        __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
        for (int v = 0; v < 4; v++) {
            int mn = m0 + ty + TY * u, i = i0 + tx + 16 * v;
            if (mn < SS && i < G) out[(size_t)mn * G + i] = acc[u][v];
        }
}

// Tensor-core variant (opt-in, `project_impl` = 1): the same contraction on the FP64 tensor pipe
// (mma.sync m8n8k4, SASS DMMA).  Each product is fused into the running sum, so `results` -- and only `results`,
// the raw tensors are untouched -- agree with the reference to ~1e-14 relative instead of bit for bit; north_star
// allows 1e-12 on the curves.  Worth it when K and S are large: the contraction grows like S^2 K^2 G per individual
// and chromosome (4.5 GFLOP at K=150, S=20) and then dwarfs the accumulation.
constexpr int DM = 64, DN = 64, DK = 16, DPAD = 8;     // block tile (surrogate pairs x bins), k-slice; row stride 72 doubles
__device__ __forceinline__ void dmma_m8n8k4(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(128)
k_project_dmma(const double *__restrict__ tensor, int K, int G, int S, const double *__restrict__ tw,
               const int32_t *__restrict__ pair_kh, double *__restrict__ results) {
    const int n = blockIdx.z;
    const int SS = S * S, P = K * (K + 1) / 2;
    const int g0 = blockIdx.x * DN, m0 = blockIdx.y * DM;
    const double *B = tensor + (size_t)n * P * G;
    double *out = results + (size_t)n * SS * G;
    __shared__ double sA[DK][DM + DPAD];      // [pair][surrogate pair]
    __shared__ double sB[DK][DN + DPAD];      // [pair][bin]
    const int lane = lane_id(), w = warp_id();
    const int wm = (w >> 1) * 32, wn = (w & 1) * 32;            // this warp's 32 x 32 corner of the block tile
    const int grp = lane >> 2, tig = lane & 3;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int mn = m0 + wm + 8 * i + grp, g = g0 + wn + 8 * j + 2 * tig;
            acc[i][j][0] = (mn < SS && g < G) ? out[(size_t)mn * G + g] : 0.0;
            acc[i][j][1] = (mn < SS && g + 1 < G) ? out[(size_t)mn * G + g + 1] : 0.0;
        }
    for (int p0 = 0; p0 < P; p0 += DK) {
        for (int e = threadIdx.x; e < DK * DM; e += 128) {
            const int pr = e / DM, c = e % DM;
            const int p = p0 + pr;
            sA[pr][c] = (p < P && m0 + c < SS) ? tw[(size_t)pair_kh[p] * SS + m0 + c] : 0.0;
            sB[pr][c] = (p < P && g0 + c < G) ? B[(size_t)p * G + g0 + c] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < DK; kk += 4) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = sA[kk + tig][wm + 8 * i + grp];     // A fragment: row grp, column tig
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = sB[kk + tig][wn + 8 * j + grp];     // B fragment: row tig, column grp
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int mn = m0 + wm + 8 * i + grp, g = g0 + wn + 8 * j + 2 * tig;
            if (mn < SS && g < G) out[(size_t)mn * G + g] = acc[i][j][0];
            if (mn < SS && g + 1 < G) out[(size_t)mn * G + g + 1] = acc[i][j][1];
        }
}

void launch_project_tensor(const double *tensor, int N, int K, int G, int S, const double *tw, const int32_t *pair_kh,
                           double *results, cudaStream_t st) {
    if (N <= 0) return;
    const size_t P = (size_t)K * (K + 1) / 2;
    for (int n0 = 0; n0 < N; n0 += kMaxGridZ) {
        const int nn = std::min(kMaxGridZ, N - n0);
        dim3 grid((G + DN - 1) / DN, (S * S + DM - 1) / DM, nn);
        k_project_dmma<<<grid, 128, 0, st>>>(tensor + (size_t)n0 * P * G, K, G, S, tw, pair_kh, results + (size_t)n0 * S * S * G);
    }
}

void launch_project(const double *tensor, int N, int K, int G, int S, const double *tw, const int32_t *pair_kh,
                    double *results, cudaStream_t st) {
    if (N <= 0) return;
    // rows (surrogate pairs) per tile: the fewest tiles of at most PTM_MAX rows, split evenly (S=10: one tile of 100 rows)
    const int SS = S * S;
    const int tiles_m = (SS + PTM_MAX - 1) / PTM_MAX;
    const int TY = std::max(1, ((SS + tiles_m - 1) / tiles_m + 3) / 4);
    const size_t P = (size_t)K * (K + 1) / 2;
    for (int n0 = 0; n0 < N; n0 += kMaxGridZ) {          // grid.z is limited to 65535 individuals per launch
        const int nn = std::min(kMaxGridZ, N - n0);
        dim3 grid((G + PT - 1) / PT, (SS + 4 * TY - 1) / (4 * TY), nn);
        k_project<<<grid, 16 * TY, 0, st>>>(tensor + (size_t)n0 * P * G, K, G, S, tw, pair_kh, results + (size_t)n0 * SS * G);
    }
}

// =============================================================================================
// 6. Symmetrise, divide by expectation (reference C:603-648), then mean over individuals in
//    individual order.
// =============================================================================================
__global__ void __launch_bounds__(128)
k_normalise(const double *__restrict__ results, int N, int S, int G, int N_total, double *__restrict__ rsbuf,
            double *__restrict__ ratio) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = blockIdx.y;
    if (i >= G) return;
    const double *R = results + (size_t)n * S * S * G + i;
    double *rs = rsbuf + (size_t)n * S * G + i;
    double *out = ratio + (size_t)n * S * S * G + i;
    double tot = 0.0;
    for (int k = 0; k < S; k++) {
        double s = 0.0;
        for (int h = 0; h < S; h++) {
            // symmetrised entry: for k<=h the reference forms (R[k,h]+R[h,k])/2 and mirrors it (C:610-615)
            int a = k <= h ? k : h, b = k <= h ? h : k;
            double sym = __dmul_rn(__dadd_rn(R[(size_t)(a * S + b) * G], R[(size_t)(b * S + a) * G]), 0.5);
            s = __dadd_rn(s, sym);
        }
        rs[(size_t)k * G] = s;
        tot = __dadd_rn(tot, s);
    }
    const double dN = (double)N_total;
    for (int k = 0; k < S; k++)
        for (int h = 0; h < S; h++) {
            int a = k <= h ? k : h, b = k <= h ? h : k;
            double sym = __dmul_rn(__dadd_rn(R[(size_t)(a * S + b) * G], R[(size_t)(b * S + a) * G]), 0.5);
            double e_ab = __ddiv_rn(__dmul_rn(rs[(size_t)a * G], rs[(size_t)b * G]), tot);
            double e_ba = __ddiv_rn(__dmul_rn(rs[(size_t)b * G], rs[(size_t)a * G]), tot);
            double ex = __dmul_rn(__dadd_rn(e_ab, e_ba), 0.5);                       // C:628-638
            out[(size_t)(k * S + h) * G] = __ddiv_rn(__ddiv_rn(sym, ex), dN);        // C:644
        }
}

__global__ void __launch_bounds__(128)
k_mean_accum(const double *__restrict__ ratio, int N, int SS, int G, double *__restrict__ means) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int kh = blockIdx.y;
    if (i >= G) return;
    double m = means[(size_t)kh * G + i];
    const double *r = ratio + (size_t)kh * G + i;
    const size_t stride = (size_t)SS * G;
    int n = 0;
    for (; n + 8 <= N; n += 8) {                         // eight loads in flight, added in individual order
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = r[(size_t)(n + u) * stride];
#pragma unroll
        for (int u = 0; u < 8; u++) m = __dadd_rn(m, v[u]);
    }
    for (; n < N; n++) m = __dadd_rn(m, r[(size_t)n * stride]);
    means[(size_t)kh * G + i] = m;
}

// the same arithmetic with one thread per (individual, surrogate row k, bin): 32 bins x S rows per block, the row sums
// exchanged through shared memory, every thread adding them up in row order for the total (reference order)
__global__ void __launch_bounds__(1024)
k_normalise_rows(const double *__restrict__ results, int S, int G, int N_total, double *__restrict__ rsbuf,
                 double *__restrict__ ratio) {
    __shared__ double srs[32][33];
    const int lane = threadIdx.x & 31, k = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + lane;
    const int n = blockIdx.y;
    const bool live = i < G;
    const double *R = results + (size_t)n * S * S * G + (live ? i : 0);
    double s = 0.0;
    for (int h = 0; h < S; h++) {
        const int a = k <= h ? k : h, b = k <= h ? h : k;
        const double sym = __dmul_rn(__dadd_rn(R[(size_t)(a * S + b) * G], R[(size_t)(b * S + a) * G]), 0.5);   // C:610-615
        s = __dadd_rn(s, sym);
    }
    srs[k][lane] = s;
    if (live) rsbuf[((size_t)n * S + k) * G + i] = s;
    __syncthreads();
    double tot = 0.0;
    for (int q = 0; q < S; q++) tot = __dadd_rn(tot, srs[q][lane]);
    if (!live) return;
    const double dN = (double)N_total;
    double *out = ratio + (size_t)n * S * S * G + i;
    for (int h = 0; h < S; h++) {
        const int a = k <= h ? k : h, b = k <= h ? h : k;
        const double sym = __dmul_rn(__dadd_rn(R[(size_t)(a * S + b) * G], R[(size_t)(b * S + a) * G]), 0.5);
        const double e_ab = __ddiv_rn(__dmul_rn(srs[a][lane], srs[b][lane]), tot);
        const double e_ba = __ddiv_rn(__dmul_rn(srs[b][lane], srs[a][lane]), tot);
        const double ex = __dmul_rn(__dadd_rn(e_ab, e_ba), 0.5);                     // C:628-638
        out[(size_t)(k * S + h) * G] = __ddiv_rn(__ddiv_rn(sym, ex), dN);            // C:644
    }
}

void launch_normalise(double *results, int N, int S, int G, int N_total, double *rsbuf, double *ratio, cudaStream_t st) {
    if (N <= 0) return;
    if (S <= 32) {
        for (int n0 = 0; n0 < N; n0 += kMaxGridZ) {      // grid.y is limited to 65535 individuals per launch
            const int nn = std::min(kMaxGridZ, N - n0);
            dim3 grid2((G + 31) / 32, nn);
            k_normalise_rows<<<grid2, 32 * S, 0, st>>>(results + (size_t)n0 * S * S * G, S, G, N_total, rsbuf + (size_t)n0 * S * G,
                                                        ratio + (size_t)n0 * S * S * G);
        }
        return;
    }
    for (int n0 = 0; n0 < N; n0 += kMaxGridZ) {
        const int nn = std::min(kMaxGridZ, N - n0);
        dim3 grid((G + 127) / 128, nn);
        k_normalise<<<grid, 128, 0, st>>>(results + (size_t)n0 * S * S * G, nn, S, G, N_total, rsbuf + (size_t)n0 * S * G,
                                          ratio + (size_t)n0 * S * S * G);
    }
}

void launch_mean_accum(const double *ratio, int N, int S, int G, double *means, cudaStream_t st) {
    dim3 grid((G + 127) / 128, S * S);
    k_mean_accum<<<grid, 128, 0, st>>>(ratio, N, S * S, G, means);
}

// reference mutiply_temp_weight C:15-23: out[(m*S+n)*K*K + k*K + l] = w[k*S+m] * w[l*S+n]
__global__ void __launch_bounds__(256) k_outer_weights(const double *__restrict__ w, int S, int K, double *__restrict__ out) {
    const size_t KK = (size_t)K * K;
    const size_t total = (size_t)S * S * KK;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int mn = (int)(idx / KK), kl = (int)(idx % KK);
        const int m = mn / S, n = mn % S, k = kl / K, l = kl % K;
        out[idx] = __dmul_rn(w[(size_t)k * S + m], w[(size_t)l * S + n]);
    }
}

void launch_outer_weights(const double *w, int S, int K, double *out, cudaStream_t st) {
    size_t total = (size_t)S * S * K * K;
    unsigned grid = (unsigned)std::min<size_t>((total + 255) / 256, 148 * 16);
    k_outer_weights<<<grid ? grid : 1, 256, 0, st>>>(w, S, K, out);
}

// upper-triangle tensor -> the reference's full [K*K][G] layout (lower triangle stays zero)
__global__ void k_expand_raw(const double *__restrict__ tensor, int K, int G, double *__restrict__ out) {
    const int n = blockIdx.z, lo = blockIdx.y / K, hi = blockIdx.y % K;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G) return;
    const int P = K * (K + 1) / 2;
    double v = 0.0;
    if (lo <= hi) v = tensor[((size_t)n * P + tri_row(lo, hi, K)) * G + i];
    out[((size_t)n * K * K + (size_t)lo * K + hi) * G + i] = v;
}

void launch_expand_raw(const double *tensor, int N, int K, int G, double *out_full, cudaStream_t st) {
    if (N <= 0) return;
    const size_t P = (size_t)K * (K + 1) / 2;
    for (int n0 = 0; n0 < N; n0 += kMaxGridZ) {
        const int nn = std::min(kMaxGridZ, N - n0);
        dim3 grid((G + 127) / 128, K * K, nn);
        k_expand_raw<<<grid, 128, 0, st>>>(tensor + (size_t)n0 * P * G, K, G, out_full + (size_t)n0 * K * K * G);
    }
}

// =============================================================================================
// 7. NULL-individual curves (reference C:916-1042): exhaustive chunk x chunk between haplotypes
//    of different individuals, full K x K x G tensor, one global sequential fold.
//    Strict mode: one warp owns one (label of A, label of B) pair = G cells held in shared
//    memory; it enumerates, in the reference's order, only the chunk pairs with those labels
//    (haplotype chunks are pre-grouped by label, stably), 32 candidates per step.
// =============================================================================================
__global__ void __launch_bounds__(256)
k_null_group(const int32_t *__restrict__ row_off, const double *__restrict__ cstart, const double *__restrict__ cend,
             const int32_t *__restrict__ chap, const int32_t *__restrict__ donor_label, int n_donor_labels,
             double cutoff, int K, Chunk *__restrict__ grouped, int32_t *__restrict__ lab_off, int *__restrict__ err) {
    // one CTA per haplotype; stable counting sort of its chunks by label (labels 1..K; label 0 and
    // out-of-range labels are dropped: they never contribute, reference C:1020)
    extern __shared__ int32_t sm[];       // [K+1] counts -> offsets
    const int hap = blockIdx.x;
    const int a = row_off[hap], b = row_off[hap + 1];
    int32_t *off = lab_off + (size_t)hap * (K + 2);
    for (int l = threadIdx.x; l <= K + 1; l += blockDim.x) sm[l] = 0;
    __syncthreads();
    int bad = 0;
    for (int i = a + threadIdx.x; i < b; i += blockDim.x) {
        int h = chap[i];
        int pop = (h >= 1 && h <= n_donor_labels) ? donor_label[h - 1] : -9;
        if (pop == -9) bad |= kErrMissingDonor;
        if (pop >= 1 && pop <= K) atomicAdd(&sm[pop], 1);
    }
    if (bad) atomicOr(err, bad);
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int l = 1; l <= K; l++) { int c = sm[l]; sm[l] = s; off[l - 1] = s; s += c; }
        off[K] = s;
        off[K + 1] = b - a;
    }
    __syncthreads();
    // stable placement: chunk i goes after all earlier chunks with the same label.  Chunks per
    // haplotype are few hundred..tens of thousands: rank by a block-wide pass per tile.
    for (int t0 = a; t0 < b; t0 += blockDim.x) {
        int i = t0 + threadIdx.x;
        int pop = 0;
        if (i < b) {
            int h = chap[i];
            pop = (h >= 1 && h <= n_donor_labels) ? donor_label[h - 1] : -9;
            if (pop < 1 || pop > K) pop = 0;
        }
        // rank among same-label lanes in this tile, in thread order
        unsigned peers = __match_any_sync(kFull, pop);
        int rank_in_warp = __popc(peers & ((1u << lane_id()) - 1));
        int cnt_in_warp = __popc(peers);
        // serialise warps of the tile: warp w adds after warps < w
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) {
            if (warp_id() == w && pop > 0) {
                int basep = sm[pop];
                if (i < b) {
                    double s = cstart[i], e = cend[i];
                    double len = __dsub_rn(e, s);
                    Chunk c;
                    c.pos = __dmul_rn(__dadd_rn(s, e), 0.5); c.size = len; c.w = (len > cutoff) ? cutoff : len;
                    c.pop = pop; c.hapid = chap[i];
                    grouped[a + basep + rank_in_warp] = c;
                }
                __syncwarp(peers);
                if (rank_in_warp == 0) sm[pop] = basep + cnt_in_warp;
            }
            __syncthreads();
        }
    }
}

void launch_null_group(int n_haps, const int32_t *row_off, const double *cstart, const double *cend, const int32_t *chap,
                       const int32_t *donor_label, int n_donor_labels, double cutoff, int K, Chunk *grouped,
                       int32_t *lab_off, int *err, cudaStream_t st) {
    k_null_group<<<n_haps, 256, (K + 2) * sizeof(int32_t), st>>>(row_off, cstart, cend, chap, donor_label, n_donor_labels,
                                                                cutoff, K, grouped, lab_off, err);
}

// One warp per (label of A, label of B, slab of fine bins).  It walks, in the reference's loop order
// (haplotype pairs, then chunk i of A, then chunk j of B), the chunk pairs that carry its two labels:
// the B group is staged in shared memory once per haplotype pair, a step compares one A chunk with 32
// consecutive B chunks (one subtraction and three compares per pair: distance window of the slab and the
// minimum-distance rule), survivors are queued FIFO in shared memory, and every 32 queued pairs get their
// exact bin (the reference's rounding rule) and are folded, in order, into the slab's cells.  Splitting each
// label pair over narrow slabs gives K*K*n_slabs independent ordered folds; since a slab has few cells the
// fold lets each cell's first lane add its group's values straight from the queue, in lane order.
constexpr int kNullQ = 64;
constexpr int kNullStage = 256;   // B chunks staged per haplotype pair (larger groups are read from global memory)
constexpr int kNullGroup = 3 * 32 + (5 * 32 + 2) / 2 + 1;   // doubles per warp for the staged group of A chunks

__global__ void __launch_bounds__(128) k_null_strict(const NullArgs a) {
    extern __shared__ __align__(16) unsigned char nsm[];
    const int lane = lane_id(), wid = warp_id();
    const int wpb = blockDim.x >> 5;
    const int K = a.K, G = a.G;
    const long long task = (long long)blockIdx.x * wpb + wid;
    if (task >= a.n_tasks) return;
    // task table (built on the host from the label counts): costliest first; hot label pairs are split
    // over many narrow slabs, cold ones over few wide ones (every slab of a pair re-walks the pair's chunk pairs)
    const NullTask tk = a.tasks[task];
    const int lA = tk.lA, lB = tk.lB;                      // 0-based labels
    const int b_lo = tk.b_lo, b_hi = tk.b_hi;              // owned bins (relative to grid_start)
    const int sbw = b_hi - b_lo;
    const int per_warp = 3 * kNullStage + 2 * kNullQ + kNullGroup + a.max_slab_bins;         // doubles
    double *__restrict__ sposB = reinterpret_cast<double *>(nsm) + (size_t)wid * per_warp;
    double *__restrict__ shalfB = sposB + kNullStage;
    double *__restrict__ swB = shalfB + kNullStage;
    double *__restrict__ qdist = swB + kNullStage;
    double *__restrict__ qval = qdist + kNullQ;
    double *__restrict__ sApos = qval + kNullQ;          // [32] the current group of A chunks: position, half size, weight
    double *__restrict__ sAhalf = sApos + 32;
    double *__restrict__ sAw = sAhalf + 32;
    int *__restrict__ sAi = reinterpret_cast<int *>(sAw + 32);        // [4][32] window starts / sizes, [33] candidate offsets
    double *__restrict__ tile = qval + kNullQ + kNullGroup;
    double *gout = a.out + ((size_t)lA * K + lB) * G + b_lo;
    for (int i = lane; i < sbw; i += 32) tile[i] = gout[i];
    __syncwarp();
    // distances that can land in [b_lo, b_hi): a safe superset window for the prefilter
    const double dwin_lo = ((double)(a.grid_start + b_lo) - 0.5) * a.bw - 1e-9 - a.bw * 1e-6;
    const double dwin_hi = ((double)(a.grid_start + b_hi - 1) + 0.5) * a.bw + 1e-9 + a.bw * 1e-6;
    const int P = a.ploidy, NH = a.n_inds * P;
    const double bw = a.bw, half_bw = a.half_bw;
    const unsigned lt = (1u << lane) - 1;
    unsigned long long n_acc = 0;
    int qn = 0;                                             // queued pairs (uniform)
    auto fold32 = [&](int cnt) {                           // exact bin + ordered fold of the first cnt (<= 32) queued pairs
        int cell = -1;
        double val = 0.0;
        if (lane < cnt) {
            const double dist = qdist[lane];
            const double qd = __ddiv_rn(dist, bw);
            const double fl = floor(qd);
            const double frac = __dmul_rn(__dsub_rn(qd, fl), bw);    // reference C:1016-1017
            int bfull = __double2int_rz(fl);
            bool ok = true;
            if (frac <= half_bw) {}
            else if (frac > half_bw) bfull += 1;
            else ok = false;
            const int bb = bfull - a.grid_start;
            if (ok && bb >= b_lo && bb < b_hi) { cell = bb - b_lo; val = qval[lane]; }
        }
        const unsigned grp = __match_any_sync(kFull, cell >= 0 ? (unsigned)cell : (0x80000000u | (unsigned)lane));
        if (cell >= 0 && (grp & lt) == 0) {                // first lane of the cell: add the group's values in lane order
            double acc = __dadd_rn(tile[cell], val);
            unsigned rest = grp & ~(1u << lane);
            while (rest) {
                const int src = __ffs(rest) - 1;
                rest &= rest - 1;
                acc = __dadd_rn(acc, qval[src]);
            }
            tile[cell] = acc;
        }
        n_acc += __popc(__ballot_sync(kFull, cell >= 0));
        __syncwarp();
        int have = qn - 32;                                // shift the remainder of the queue to the front
        double md = 0.0, mv = 0.0;
        if (lane < have) { md = qdist[lane + 32]; mv = qval[lane + 32]; }
        __syncwarp();
        if (lane < have) { qdist[lane] = md; qval[lane] = mv; }
        __syncwarp();
    };
    for (int hA = 0; hA < (a.n_inds - 1) * P; hA++) {
        const int nA = hA / P;
        const int32_t *oA = a.lab_off + (size_t)hA * (K + 2);
        const int a0 = oA[lA], cntA = oA[lA + 1] - a0;
        if (cntA == 0) continue;
        const Chunk *CA = a.chunks + a.hap_base[hA] + a0;
        for (int hB = (nA + 1) * P; hB < NH; hB++) {
            const int32_t *oB = a.lab_off + (size_t)hB * (K + 2);
            const int b0 = oB[lB], cntB = oB[lB + 1] - b0;
            if (cntB == 0) continue;
            const Chunk *CB = a.chunks + a.hap_base[hB] + b0;
            const bool staged = cntB <= kNullStage;
            if (staged) {
                __syncwarp();
                for (int j = lane; j < cntB; j += 32) {
                    const Chunk B = CB[j];
                    sposB[j] = B.pos; shalfB[j] = __dmul_rn(B.size, 0.5); swB[j] = B.w;
                }
                __syncwarp();
            }
            auto push = [&](bool surv, double dist, double val) {      // ordered append of one step's survivors
                const unsigned m = __ballot_sync(kFull, surv);
                if (m) {
                    if (surv) { const int pos = qn + __popc(m & lt); qdist[pos] = dist; qval[pos] = val; }
                    qn += __popc(m);
                    __syncwarp();
                    if (qn >= 32) { fold32(32); qn -= 32; }
                }
            };
            auto test = [&](const Chunk &A, double halfA, int jb, bool &surv, double &dist, double &val) {
                surv = false; dist = 0.0; val = 0.0;
                if (jb < cntB) {
                    double pbv, hb, wb;
                    if (staged) { pbv = sposB[jb]; hb = shalfB[jb]; wb = swB[jb]; }
                    else { const Chunk B = CB[jb]; pbv = B.pos; hb = __dmul_rn(B.size, 0.5); wb = B.w; }
                    dist = fabs(__dsub_rn(A.pos, pbv));            // reference C:1015
                    // C:1019-1020: dist >= size_A/2 + size_B/2 ; and the slab's distance window
                    surv = dist >= dwin_lo && dist <= dwin_hi && dist >= __dadd_rn(halfA, hb);
                    val = __dmul_rn(A.w, wb);
                }
            };
            int ia = 0;
            if (staged) {
                // Windowed enumeration.  The B group is sorted by position, and only B chunks at a distance inside the slab's
                // window [dwin_lo, dwin_hi] from an A chunk can land in the owned bins: 32 A chunks at a time, every lane
                // bisects the (at most two) index ranges of its A chunk, a warp scan numbers the candidates in (A, B) order,
                // and the steps below evaluate 32 candidates each -- instead of stepping through the whole B group per A chunk.
                int *sL0 = sAi, *snl = sAi + 32, *sR0 = sAi + 64, *snr = sAi + 96, *soff = sAi + 128;
                for (; ia < cntA; ia += 32) {
                    const int my = ia + lane;
                    int L0 = 0, nl = 0, R0 = 0, nr = 0;
                    __syncwarp();
                    if (my < cntA) {
                        const Chunk A = CA[my];
                        sApos[lane] = A.pos; sAhalf[lane] = __dmul_rn(A.size, 0.5); sAw[lane] = A.w;
                        auto lower = [&](double x) {          // first index with pos >= x
                            int lo = 0, hi = cntB;
                            while (lo < hi) { const int mid = (lo + hi) >> 1; if (sposB[mid] < x) lo = mid + 1; else hi = mid; }
                            return lo;
                        };
                        auto upper = [&](double x) {          // first index with pos > x
                            int lo = 0, hi = cntB;
                            while (lo < hi) { const int mid = (lo + hi) >> 1; if (sposB[mid] <= x) lo = mid + 1; else hi = mid; }
                            return lo;
                        };
                        const double slack = 1e-9;            // rounding of pos +- window; the exact tests follow
                        L0 = lower(A.pos - dwin_hi - slack);
                        int L1 = upper(A.pos - dwin_lo + slack);
                        R0 = lower(A.pos + dwin_lo - slack);
                        const int R1 = upper(A.pos + dwin_hi + slack);
                        if (L1 > R0) { L1 = R1; R0 = R1; }   // the windows touch (the slab holds distance ~0): one range
                        nl = L1 - L0; nr = R1 - R0;
                        if (nl < 0) nl = 0;
                        if (nr < 0) nr = 0;
                    }
                    const int mine = nl + nr;
                    int incl = mine;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int y = __shfl_up_sync(kFull, incl, o);
                        if (lane >= o) incl += y;
                    }
                    const int T = __shfl_sync(kFull, incl, 31);
                    sL0[lane] = L0; snl[lane] = nl; sR0[lane] = R0; snr[lane] = nr; soff[lane] = incl - mine;
                    if (lane == 0) soff[32] = T;
                    __syncwarp();
                    for (int t0 = 0; t0 < T; t0 += 32) {
                        const int t = t0 + lane;
                        bool surv = false;
                        double dist = 0.0, val = 0.0;
                        if (t < T) {
                            int lo = 0, hi = 31;                 // owner: the last A chunk whose first candidate is <= t
                            while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (soff[mid] <= t) lo = mid; else hi = mid - 1; }
                            const int u = lo;
                            const int k = t - soff[u];
                            const int jb = k < snl[u] ? sL0[u] + k : sR0[u] + (k - snl[u]);
                            dist = fabs(__dsub_rn(sApos[u], sposB[jb]));                       // reference C:1015
                            surv = dist >= dwin_lo && dist <= dwin_hi && dist >= __dadd_rn(sAhalf[u], shalfB[jb]);   // C:1019-1020
                            val = __dmul_rn(sAw[u], swB[jb]);
                        }
                        push(surv, dist, val);
                    }
                }
            }
            for (; ia < cntA; ia++) {
                const Chunk A = CA[ia];                    // same address in every lane: one broadcast load
                const double halfA = __dmul_rn(A.size, 0.5);
                int jb0 = 0;
                for (; jb0 + 64 <= cntB + 31 && jb0 + 32 < cntB; jb0 += 64) {      // two steps at a time
                    bool s0, s1; double d0, d1, v0, v1;
                    test(A, halfA, jb0 + lane, s0, d0, v0);
                    test(A, halfA, jb0 + 32 + lane, s1, d1, v1);
                    push(s0, d0, v0);
                    push(s1, d1, v1);
                }
                for (; jb0 < cntB; jb0 += 32) {
                    bool s0; double d0, v0;
                    test(A, halfA, jb0 + lane, s0, d0, v0);
                    push(s0, d0, v0);
                }
            }
        }
    }
    if (qn > 0) { fold32(qn); qn = 0; }
    for (int i = lane; i < sbw; i += 32) gout[i] = tile[i];
    if (lane == 0 && n_acc) atomicAdd(a.accepted, n_acc);
}

void launch_null_strict(const NullArgs &a, cudaStream_t st) {
    const int wpb = 4;
    long long tasks = a.n_tasks;
    if (tasks <= 0) return;
    size_t smem = (size_t)wpb * (3 * kNullStage + 2 * kNullQ + kNullGroup + a.max_slab_bins) * sizeof(double);
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) { cudaFuncSetAttribute(k_null_strict, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr_set[dev & 63] = true; }
    k_null_strict<<<(unsigned)((tasks + wpb - 1) / wpb), wpb * 32, smem, st>>>(a);
}

// ---- NULL path, two phases (null_impl = 2) ------------------------------------------------------------------------
// Phase 1 evaluates every chunk pair of every haplotype pair exactly ONCE (k_null_strict lets every (label pair, slab)
// task re-walk all haplotype pairs) and files the accepted updates as 16-byte records; phase 2 is k_commit_ordered over
// an explicit task table: task = (label of A, label of B, range of fine bins) = a few hundred cells in shared memory, one
// contiguous record region per task.
//
// Order.  The reference adds in (n, p, g, m, i, j) order; only the relative order of the additions into ONE cell matters.
// A work item is (haplotype pair "tile", label of A): one warp walks the A chunks with that label in order and, per A
// chunk, the B chunks in label-grouped order (k_null_group is stable, so within a label of B that is the reference's j
// order).  Records of a task are laid out tile-major (tiles in (n, p, g, m) order), inside a tile in the warp's (i, j)
// order: every cell sees its addends in the reference's order.
//
// Positions.  A counting pass (same evaluation, shared-memory counters) fills counts[task][tile]; one exclusive scan
// turns that into every (task, tile) region's first record, and a task's tiles are adjacent, so phase 2 needs one
// RegionEnt per task.  The scatter pass ranks the accepted pairs of a 32-pair step without match.any (two cycles per distinct
// key on the SM's single unit): B is grouped by label, so the task numbers met in one step lie close together and the
// lanes of one task are found with one ballot per bit of (task - smallest task of the step).
struct NullStep { unsigned key; int cell, local; double val; };   // key = bin range of the lane's pair inside its label pair, ~0u = rejected

constexpr int kNullWarps = 4;          // warps per CTA of k_null_eval2
constexpr int kNullMaxRanges = 64;     // most bin ranges of one label pair (shared-memory counters per warp)

__device__ __forceinline__ NullStep null_eval_pair(const NullEvalArgs &a, const Chunk &A, double posB, double halfB, double wB, int cell0,
                                                    unsigned magic, int w, int koff, bool valid) {
    NullStep o;
    o.key = ~0u; o.cell = -1; o.local = 0; o.val = 0.0;
    if (!valid) return o;
    const double dist = fabs(__dsub_rn(A.pos, posB));                       // reference C:1015
    double qd;
    if (a.inv_bw != 0.0) {                                                  // dist / bw, correctly rounded (see eval_pair)
        const double q0 = __dmul_rn(dist, a.inv_bw);
        qd = __fma_rn(__fma_rn(-q0, a.bw, dist), a.inv_bw, q0);
    } else qd = __ddiv_rn(dist, a.bw);
    const double fl = floor(qd);
    const double frac = __dmul_rn(__dsub_rn(qd, fl), a.bw);                 // reference C:1016-1017
    int bfull = __double2int_rz(fl);
    bool ok = true;
    if (frac <= a.half_bw) {}
    else if (frac > a.half_bw) bfull += 1;
    else ok = false;
    const int bb = bfull - a.grid_start;
    const double mind = __dadd_rn(__dmul_rn(A.size, 0.5), halfB);           // reference C:1019-1020
    if (ok && (unsigned)bb < (unsigned)a.G && dist >= mind) {
        const int r = magic ? (int)__umulhi((unsigned)bb, magic) : bb;      // bb / w (exact below 2^16; magic 0: w = 1)
        o.key = (unsigned)(koff + r);
        o.local = bb - r * w;
        o.cell = cell0 + bb;
        o.val = __dmul_rn(A.w, wB);
    }
    return o;
}

// ---- phase 1: persistent warps, every warp on its own ---------------------------------------------------------------------
// A warp claims (haplotype pair, label of A) units from one global counter, heaviest labels first (no CTA barrier, no idle warps at
// the end of a haplotype pair) and stages what it needs in ITS OWN slice of shared memory: the chunks of A with that label once
// per unit, the chunks of the current label (block of labels) of B once per label pair -- a 1/|A group| overhead, and no limit on
// the haplotype length (an earlier form staged the whole B haplotype per CTA and fell back to global loads beyond 2048 chunks:
// every long chromosome of real data).  The chunk pairs of one (label of A, label of B) are dealt to the lanes 32 at a time in
// (i, j) order, so that at any time a warp writes to the few record regions of ONE label pair -- their sectors fill up and leave
// L2 whole.  BLOCKS (small label groups, large K): the labels of B in blocks of consecutive labels with at most 32 bin ranges
// between them, so that a step is full even when a label group holds a handful of chunks; the lanes then look up their own
// label's plan.
// Optional windowed enumeration (one label of B at a time): every lane bisects, for its A chunk, the index range of the position-sorted
// B group within the largest distance the grid holds; a warp scan numbers the candidates of 32 A chunks in (i, j) order and the
// steps take 32 candidates each -- on chromosomes much longer than the window most chunk pairs are never touched.
constexpr int kNullGrp = 128;                               // chunks of a group staged per warp (larger groups: global loads)
constexpr int kNullWarpBytes = kNullGrp * (16 + 8) * 2 + kNullMaxRanges * 4 + kNullGrp * 2;   // B, A, counters, labels of B

template <bool SCATTER>
__device__ __forceinline__ void null_file(const NullEvalArgs &a, int32_t *cnt, const NullStep &e, int nr, unsigned lt, int lane) {
    const bool acc = e.key != ~0u;
    if (!SCATTER) {
        if (acc) atomicAdd(&cnt[e.key], 1);
    } else {
        // stable rank inside the step: the lanes of the same bin range, in lane order (= (i, j) order)
        if (!__any_sync(kFull, acc)) return;
        const unsigned M = nr > 1 ? __match_any_sync(kFull, e.key) : __ballot_sync(kFull, acc);
        const int first = __ffs(M) - 1;
        int base = 0;
        if (acc && lane == first) base = atomicAdd(&cnt[e.key], __popc(M));   // one lane per range of the step
        base = __shfl_sync(kFull, base, first);
        if (acc) {
            UpdateRec rec;
            rec.cell = e.cell; rec.pad = e.local; rec.val = e.val;
            a.recs[base + __popc(M & lt)] = rec;
        }
    }
}

template <bool SCATTER, bool BLOCKS>
__global__ void __launch_bounds__(32 * kNullWarps) k_null_eval2(const NullEvalArgs a) {
    extern __shared__ __align__(16) unsigned char nev2_sm[];
    unsigned char *wb = nev2_sm + (size_t)warp_id() * kNullWarpBytes;
    double2 *sBph = reinterpret_cast<double2 *>(wb);                                      // [kNullGrp] (position, half size)
    double2 *sAph = sBph + kNullGrp;
    double *sBw = reinterpret_cast<double *>(sAph + kNullGrp);
    double *sAw = sBw + kNullGrp;
    int32_t *cnt = reinterpret_cast<int32_t *>(sAw + kNullGrp);
    uint16_t *sBlab = reinterpret_cast<uint16_t *>(cnt + kNullMaxRanges);
    const int K = a.K;
    uint2 *prow = reinterpret_cast<uint2 *>(nev2_sm + (size_t)kNullWarps * kNullWarpBytes) + (size_t)warp_id() * K;   // BLOCKS
    const int lane = lane_id();
    const unsigned lt = (1u << lane) - 1;
    const long long NT = a.n_tasks;
    const long long n_units = (long long)a.n_tiles * a.n_own;
    while (true) {
        long long unit = 0;
        if (lane == 0) unit = atomicAdd(a.unit_counter, 1ULL);
        unit = __shfl_sync(kFull, unit, 0);
        if (unit >= n_units) break;
        const int tile = (int)(unit / a.n_own);
        const int la = a.own[unit - (long long)tile * a.n_own];
        const int2 hh = a.tiles[tile];
        const int32_t *oA = a.lab_off + (size_t)hh.x * (K + 2);
        const int32_t *oB = a.lab_off + (size_t)hh.y * (K + 2);
        const Chunk *CB = a.chunks + a.hap_base[hh.y];
        const NullRowPlan rpl = a.rows[la];
        int32_t *grow = (SCATTER ? const_cast<int32_t *>(a.offs) : a.counts) + (size_t)tile * NT + rpl.task0;   // [tile][task]
        const int a0 = oA[la], cntA = oA[la + 1] - a0;
        if (cntA <= 0 || oB[K] <= 0) {
            if (!SCATTER) for (int t = lane; t < rpl.ntb; t += 32) grow[t] = 0;
            continue;
        }
        const Chunk *CA = a.chunks + a.hap_base[hh.x] + a0;
        const NullPairPlan *row = a.plan + (size_t)la * K;
        const bool stagedA = cntA <= kNullGrp;
        __syncwarp();
        if (stagedA)
            for (int i = lane; i < cntA; i += 32) {
                const Chunk A = ldg_chunk(CA + i);
                sAph[i] = make_double2(A.pos, __dmul_rn(A.size, 0.5)); sAw[i] = A.w;
            }
        int blk0 = 0, nblk = K;
        if (BLOCKS) {
            blk0 = a.blk_start[la]; nblk = a.blk_start[la + 1] - blk0;
            for (int lb = lane; lb < K; lb += 32) {
                const NullPairPlan pp = row[lb];
                prow[lb] = make_uint2(pp.magic, (unsigned)pp.tb0 | ((unsigned)pp.w << 16));
            }
        }
        for (int bi = 0; bi < nblk; bi++) {
            int lb = bi, lb1 = bi + 1;
            if (BLOCKS) { const int2 bk = a.blk[blk0 + bi]; lb = bk.x; lb1 = bk.y; }
            const NullPairPlan pp = row[lb];
            const int b0 = oB[lb], cntB = oB[lb1] - b0;              // chunks of B in this label (block of labels)
            const int nr = BLOCKS ? ((lb1 < K ? row[lb1].tb0 : rpl.ntb) - pp.tb0) : (a.G + pp.w - 1) / pp.w;
            int32_t *gcnt = grow + pp.tb0;
            const bool stagedB = cntB <= kNullGrp;
            __syncwarp();                                            // the previous label's steps are done with the staged group
            for (int t = lane; t < nr; t += 32) cnt[t] = (SCATTER && cntB > 0) ? gcnt[t] : 0;
            if (stagedB)
                for (int j = lane; j < cntB; j += 32) {
                    const Chunk B = ldg_chunk(CB + b0 + j);
                    sBph[j] = make_double2(B.pos, __dmul_rn(B.size, 0.5)); sBw[j] = B.w;
                    if (BLOCKS) sBlab[j] = (uint16_t)(B.pop - 1);
                }
            __syncwarp();
            if (cntB > 0) {
                const int cell0u = (la * K + lb) * a.G;
                auto eval = [&](int i, int j, bool valid) -> NullStep {
                    Chunk A;                                          // (pos, size, w) of A as eval needs them: size = 2 * half
                    double halfA;
                    if (stagedA) { const double2 ph = sAph[i]; A.pos = ph.x; halfA = ph.y; A.w = sAw[i]; }
                    else { const Chunk g = ldg_chunk(CA + i); A.pos = g.pos; halfA = __dmul_rn(g.size, 0.5); A.w = g.w; }
                    A.size = __dadd_rn(halfA, halfA);                 // exact: null_eval_pair halves it again
                    double posB, halfB, wB;
                    int labB = lb;
                    if (stagedB) {
                        const double2 ph = sBph[j]; posB = ph.x; halfB = ph.y; wB = sBw[j];
                        if (BLOCKS) labB = sBlab[j];
                    } else {
                        const Chunk B = ldg_chunk(CB + b0 + j); posB = B.pos; halfB = __dmul_rn(B.size, 0.5); wB = B.w;
                        if (BLOCKS) labB = B.pop - 1;
                    }
                    if (BLOCKS) {
                        const uint2 pl = prow[labB];
                        return null_eval_pair(a, A, posB, halfB, wB, (la * K + labB) * a.G, pl.x, (int)(pl.y >> 16), (int)(pl.y & 0xffffu) - pp.tb0, valid);
                    }
                    return null_eval_pair(a, A, posB, halfB, wB, cell0u, pp.magic, pp.w, 0, valid);
                };
                if (!BLOCKS && a.window && stagedA && stagedB) {
                    // windowed enumeration: 32 A chunks at a time, every lane the index range of B within the largest distance of the grid
                    for (int i0 = 0; i0 < cntA; i0 += 32) {
                        const int i = i0 + lane;
                        int lo = 0, n = 0;
                        if (i < cntA) {
                            const double xa = sAph[i].x;
                            const double x0 = xa - a.dwin, x1 = xa + a.dwin;
                            int l = 0, h = cntB;
                            while (l < h) { const int m = (l + h) >> 1; if (sBph[m].x < x0) l = m + 1; else h = m; }
                            lo = l; h = cntB;
                            while (l < h) { const int m = (l + h) >> 1; if (sBph[m].x <= x1) l = m + 1; else h = m; }
                            n = l - lo;
                        }
                        int incl = n;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const int y = __shfl_up_sync(kFull, incl, o);
                            if (lane >= o) incl += y;
                        }
                        const int T = __shfl_sync(kFull, incl, 31);
                        const int off = incl - n;
                        for (int t0 = 0; t0 < T; t0 += 32) {
                            const int t = t0 + lane;
                            int u = 0;                                 // owner: the last lane whose first candidate is <= t
#pragma unroll
                            for (int sft = 16; sft >= 1; sft >>= 1) {
                                const int c = u + sft;
                                const int v = __shfl_sync(kFull, off, c & 31);
                                if (c < 32 && v <= t) u = c;
                            }
                            const int lo_u = __shfl_sync(kFull, lo, u), off_u = __shfl_sync(kFull, off, u);
                            const bool valid = t < T;
                            const NullStep e = eval(valid ? i0 + u : 0, valid ? lo_u + (t - off_u) : 0, valid);
                            null_file<SCATTER>(a, cnt, e, nr, lt, lane);
                        }
                    }
                } else {
                    const long long npairs = (long long)cntA * cntB;   // (two haplotypes of 50k chunks with one label: beyond 32 bits)
                    const unsigned bmagic = cntB > 1 ? (unsigned)((0x100000000ULL + cntB - 1) / cntB) : 0u;   // q / cntB for q < 2^16 ...
                    const bool small = npairs < 65536 && cntB < 65536;                                          // ... else a real division
                    for (long long q0 = 0; q0 < npairs; q0 += 64) {
                        NullStep e[2];
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const long long q = q0 + 32 * h + lane;
                            const bool valid = q < npairs;
                            int i = 0, j = 0;
                            if (valid) {
                                i = cntB == 1 ? (int)q : (small ? (int)__umulhi((unsigned)q, bmagic) : (int)(q / cntB));
                                j = (int)(q - (long long)i * cntB);
                            }
                            e[h] = eval(i, j, valid);
                        }
                        null_file<SCATTER>(a, cnt, e[0], nr, lt, lane);
                        if (q0 + 32 < npairs) null_file<SCATTER>(a, cnt, e[1], nr, lt, lane);
                    }
                }
            }
            if (!SCATTER) {
                __syncwarp();
                for (int t = lane; t < nr; t += 32) gcnt[t] = cnt[t];
            }
        }
    }
}

// sum of n int32 counts as a 64-bit total (the records of a batch, before its buffer is sized)
__global__ void __launch_bounds__(256) k_sum_i32(const int32_t *__restrict__ in, long long n, unsigned long long *__restrict__ total) {
    unsigned long long s = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) s += (unsigned long long)in[i];
    s = warp_sum_u64(s);
    if (lane_id() == 0 && s) atomicAdd(total, s);
}
void launch_sum_i32(const int32_t *in, long long n, unsigned long long *total, cudaStream_t st) {
    if (n <= 0) return;
    k_sum_i32<<<(unsigned)std::min<long long>((n + 255) / 256, 148 * 8), 256, 0, st>>>(in, n, total);
}

// [rows][cols] int32 -> [cols][rows]
__global__ void __launch_bounds__(256) k_transpose_i32(const int32_t *__restrict__ in, long long rows, long long cols, int32_t *__restrict__ out) {
    __shared__ int32_t t[32][33];
    const long long c0 = (long long)blockIdx.x * 32, r0 = (long long)blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int k = ty; k < 32; k += 8)
        if (r0 + k < rows && c0 + tx < cols) t[k][tx] = in[(r0 + k) * cols + c0 + tx];
    __syncthreads();
    for (int k = ty; k < 32; k += 8)
        if (c0 + k < cols && r0 + tx < rows) out[(c0 + k) * rows + r0 + tx] = t[tx][k];
}
void launch_transpose_i32(const int32_t *in, long long rows, long long cols, int32_t *out, cudaStream_t st) {
    if (rows <= 0 || cols <= 0) return;
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
    k_transpose_i32<<<grid, 256, 0, st>>>(in, rows, cols, out);
}

// one record region per commit task: commit task c folds natural task perm[c], whose tiles are adjacent in the record buffer
__global__ void k_null_dir(const int32_t *__restrict__ offs, const int32_t *__restrict__ perm, int n_tasks, long long T,
                           RegionEnt *__restrict__ dir) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_tasks) return;
    const long long t = perm[c];
    RegionEnt o;
    o.base = offs[t * T];
    o.cnt = offs[(t + 1) * T] - offs[t * T];
    o.pad = 0;
    dir[c] = o;
}

void launch_null_eval(const NullEvalArgs &a, bool scatter, cudaStream_t st) {
    if (a.n_tiles <= 0 || a.n_own <= 0) return;
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
        cudaFuncSetAttribute(k_null_eval2<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(k_null_eval2<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(k_null_eval2<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(k_null_eval2<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set[dev & 63] = true;
    }
    const size_t smem = (size_t)kNullWarps * kNullWarpBytes + (a.blk ? (size_t)kNullWarps * a.K * sizeof(uint2) : 0);
    long long per_sm = std::max<long long>(1, std::min<long long>(2048 / (32 * kNullWarps), (227 * 1024) / (long long)(smem + 1024)));
    if (a.max_ctas_per_sm > 0) per_sm = std::min<long long>(per_sm, a.max_ctas_per_sm);
    const unsigned grid = (unsigned)std::min<long long>(((long long)a.n_tiles * a.n_own + kNullWarps - 1) / kNullWarps, 148 * per_sm);
    if (a.blk) {
        if (scatter) k_null_eval2<true, true><<<grid, 32 * kNullWarps, smem, st>>>(a);
        else k_null_eval2<false, true><<<grid, 32 * kNullWarps, smem, st>>>(a);
    } else {
        if (scatter) k_null_eval2<true, false><<<grid, 32 * kNullWarps, smem, st>>>(a);
        else k_null_eval2<false, false><<<grid, 32 * kNullWarps, smem, st>>>(a);
    }
}
size_t null_eval_smem(int K, bool blocks) { return (size_t)kNullWarps * kNullWarpBytes + (blocks ? (size_t)kNullWarps * K * sizeof(uint2) : 0); }

void launch_null_dir(const int32_t *offs, const int32_t *perm, int n_tasks, long long T, RegionEnt *dir, cudaStream_t st) {
    if (n_tasks <= 0) return;
    k_null_dir<<<(n_tasks + 255) / 256, 256, 0, st>>>(offs, perm, n_tasks, T, dir);
}

}  // namespace fgt

// =============================================================================================
// 8. Callers of the path, as opt-in extra entry points: the R script stays unchanged unless a user
//    switches to them.  Tolerance-level by nature (R's %*% is BLAS, lm() is Householder QR).
// =============================================================================================
namespace fgt {

// ---- 8a. null normalisation, fastGLOBETROTTER.R:1216-1249 -------------------------------------
// results = tempweights %*% null  ([S^2 x K^2] . [K^2 x G]); tw is in data_read's layout [(k*K+l)*S^2 + mn].
__global__ void __launch_bounds__(256)
k_weights_times_null(const double *__restrict__ tw, const double *__restrict__ nul, int SS, int KK, int G, double *__restrict__ out) {
    const int b = blockIdx.x * 32 + (threadIdx.x & 31);
    const int mn = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (b >= G || mn >= SS) return;
    double acc = 0.0;
    for (int p = 0; p < KK; p++) acc = __dadd_rn(acc, __dmul_rn(tw[(size_t)p * SS + mn], nul[(size_t)p * G + b]));
    out[(size_t)mn * G + b] = acc;
}

// per bin: symmetrise, expectation from the margins, divide (R:1223-1248); one thread per bin, S <= 64
__global__ void __launch_bounds__(128)
k_null_normalise(const double *__restrict__ res, int S, int G, double *__restrict__ nullcurves, double *__restrict__ means) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= G) return;
    double s2[64], s3[64];
    for (int k = 0; k < S; k++) { s2[k] = 0.0; s3[k] = 0.0; }
    double tot = 0.0;
    auto sym = [&](int k, int l) { return __dmul_rn(__dadd_rn(res[(size_t)(k * S + l) * G + b], res[(size_t)(l * S + k) * G + b]), 0.5); };
    for (int k = 0; k < S; k++)
        for (int l = 0; l < S; l++) {
            const double r = sym(k, l);
            tot = __dadd_rn(tot, r); s2[k] = __dadd_rn(s2[k], r); s3[l] = __dadd_rn(s3[l], r);
        }
    for (int k = 0; k < S; k++)
        for (int l = 0; l < S; l++) {
            const double e_kl = __ddiv_rn(__dmul_rn(s2[k], s3[l]), tot), e_lk = __ddiv_rn(__dmul_rn(s2[l], s3[k]), tot);
            const double ex = __dmul_rn(__dadd_rn(e_kl, e_lk), 0.5);
            const double nc = __ddiv_rn(sym(k, l), ex);
            const size_t o = (size_t)(k * S + l) * G + b;
            if (nullcurves) nullcurves[o] = nc;
            if (means) means[o] = __ddiv_rn(means[o], nc);
        }
}

void launch_null_normalise(const double *tw, const double *nul, int S, int K, int G, double *scratch, double *nullcurves, double *means,
                           cudaStream_t st) {
    dim3 grid((G + 31) / 32, (S * S + 7) / 8);
    k_weights_times_null<<<grid, 256, 0, st>>>(tw, nul, S * S, K * K, G, scratch);
    k_null_normalise<<<(G + 127) / 128, 128, 0, st>>>(scratch, S, G, nullcurves, means);
}

// ---- 8b. date likelihood, fastGLOBETROTTER.R:441-489 (getlhood) over a grid of candidate dates --------------------
// one thread per (candidate, curve): ordinary least squares of the curve on 1 + exp(-t*a/100) (+ a second date), the way
// lm() resolves it (a column whose part orthogonal to the earlier ones is below 1e-7 of its norm gets an NA coefficient)
struct FitOut { double lh, c1, c2, sd3; };

__global__ void __launch_bounds__(128)
k_getlhood_fit(const double *__restrict__ means, int R, int T, const double *__restrict__ times, const double *__restrict__ cand,
               int n_cand, int nparam, FitOut *__restrict__ out) {
    const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (long long)n_cand * R) return;
    const int c = (int)(id / R), i = (int)(id - (long long)c * R);
    const double a1 = cand[(size_t)c * nparam], a2 = nparam == 2 ? cand[(size_t)c * nparam + 1] : 0.0;
    const double *y = means + (size_t)i * T;
    double sy = 0, s1 = 0, s2 = 0, n1 = 0, n2 = 0;
    for (int t = 0; t < T; t++) {
        const double p1 = exp(-times[t] * a1 / 100.0), p2 = nparam == 2 ? exp(-times[t] * a2 / 100.0) : 0.0;
        sy += y[t]; s1 += p1; s2 += p2; n1 += p1 * p1; n2 += p2 * p2;
    }
    const double yb = sy / T, b1m = s1 / T, b2m = s2 / T;
    double c11 = 0, c12 = 0, c22 = 0, c1y = 0, c2y = 0;
    for (int t = 0; t < T; t++) {
        const double p1 = exp(-times[t] * a1 / 100.0) - b1m, p2 = (nparam == 2 ? exp(-times[t] * a2 / 100.0) : 0.0) - b2m, yc = y[t] - yb;
        c11 += p1 * p1; c12 += p1 * p2; c22 += p2 * p2; c1y += p1 * yc; c2y += p2 * yc;
    }
    const double tol2 = 1e-14;                 // (1e-7)^2, lm()'s rank tolerance on column norms
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    double coef1 = nan, coef2 = nan;
    const bool ok1 = c11 >= tol2 * n1 && n1 > 0.0;
    bool ok2 = false;
    if (ok1) {
        if (nparam == 2) {
            const double q22 = c22 - c12 * c12 / c11;             // squared norm of the second column's orthogonal part
            ok2 = q22 >= tol2 * n2 && n2 > 0.0;
            if (ok2) {
                coef2 = (c2y - c12 * c1y / c11) / q22;
                coef1 = (c1y - coef2 * c12) / c11;
            } else coef1 = c1y / c11;
        } else coef1 = c1y / c11;
    } else if (nparam == 2 && c22 >= tol2 * n2 && n2 > 0.0) { coef2 = c2y / c22; ok2 = true; }
    const double u1 = ok1 ? coef1 : 0.0, u2 = ok2 ? coef2 : 0.0;
    double ss = 0.0;
    const int h0 = T / 2;                      // R: res[floor(T/2):T] (1-based, inclusive) -> 0-based from floor(T/2)-1
    const int from = h0 - 1 < 0 ? 0 : h0 - 1;
    double sa = 0.0, saa = 0.0;
    int na = 0;
    for (int t = 0; t < T; t++) {
        const double p1 = exp(-times[t] * a1 / 100.0) - b1m, p2 = (nparam == 2 ? exp(-times[t] * a2 / 100.0) : 0.0) - b2m;
        const double r = (y[t] - yb) - u1 * p1 - u2 * p2;
        ss += r * r;
        if (t >= from) { const double ar = fabs(r); sa += ar; saa += ar * ar; na++; }
    }
    FitOut o;
    o.lh = -0.5 * T * log(ss / T);
    o.c1 = coef1; o.c2 = coef2;
    const double var = na > 1 ? (saa - sa * sa / na) / (na - 1) : nan;
    o.sd3 = 3.0 * sqrt(var > 0.0 ? var : 0.0);
    out[id] = o;
}

// one thread per candidate: the sum over curves and the penalty rules of R:455-488 (diagonal curves carry popfracs)
__global__ void k_getlhood_reduce(const FitOut *__restrict__ fit, int R, int n_cand, int nparam, const double *__restrict__ cand,
                                  const double *__restrict__ popfracs, double likpenalty, double *__restrict__ lhood) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cand) return;
    const int S = (int)(sqrt((double)R) + 0.5);
    double lh = 0.0, ts0 = 0.0, ts1 = 0.0;
    bool neg = false;
    for (int i = 0; i < R; i++) {
        const FitOut f = fit[(size_t)c * R + i];
        if (i % (S + 1) == 0 && i / (S + 1) < S) {             // seq.to.capture: the (k,k) curves
            const double pf = popfracs[i / S];                 // popfracs[((i-1)/sqrt(R)+1)] with R's 1-based i: k
            if (nparam == 2) {
                const double a1 = cand[(size_t)c * 2], a2 = cand[(size_t)c * 2 + 1];
                const bool na3 = isnan(f.c2);
                if (a1 != a2 && !na3) {
                    const bool s1 = f.c1 > 0, s2 = f.c2 > 0, z1 = f.c1 == 0, z2 = f.c2 == 0;
                    const int sg1 = z1 ? 0 : (s1 ? 1 : -1), sg2 = z2 ? 0 : (s2 ? 1 : -1);
                    if (sg1 != sg2 && fmin(f.c1, f.c2) < -f.sd3) neg = true;
                    ts0 += pf * fabs(f.c1); ts1 += pf * fabs(f.c2);
                }
                if (na3) neg = true;
            } else ts0 += pf * fabs(f.c1);
        }
        lh += f.lh;
    }
    double mx = nparam == 2 ? fmax(ts0, ts1) : ts0;
    if (isnan(ts0) || isnan(ts1)) { mx = 0.0; neg = true; }
    lhood[c] = (neg || mx > 2.0) ? lh / likpenalty : lh;
}

void launch_getlhood(const double *means, int R, int T, const double *times, const double *cand, int n_cand, int nparam,
                     const double *popfracs, double likpenalty, void *fit_scratch, double *lhood, cudaStream_t st) {
    const long long n = (long long)n_cand * R;
    if (n <= 0) return;
    k_getlhood_fit<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(means, R, T, times, cand, n_cand, nparam, (FitOut *)fit_scratch);
    k_getlhood_reduce<<<(n_cand + 63) / 64, 64, 0, st>>>((const FitOut *)fit_scratch, R, n_cand, nparam, cand, popfracs, likpenalty, lhood);
}
size_t getlhood_scratch_bytes(int R, int n_cand) { return (size_t)R * n_cand * sizeof(FitOut); }

// ---- 8c. bootstrap replicates from per-(individual, chromosome) projected curves ---------------------
// part[h][p][mn][b]: projection of packed individual p's tensor after chromosome h (mode 1: cumulative over chromosomes,
// so chromosome h's own share is part[h] - part[h-1]; mode 3: chromosome h alone).  A replicate's pseudo-individual i uses
// individual ids[i*Cn + h] on chromosome h (R:2058-2059).
__global__ void __launch_bounds__(256)
k_bootstrap_combine(const double *__restrict__ part, int Cn, int Np, size_t cells, int cumulative, const int32_t *__restrict__ ids,
                    int N, double *__restrict__ results) {
    const int i = blockIdx.y;                                   // pseudo-individual of this replicate
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < cells; e += (size_t)gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (int h = 0; h < Cn; h++) {
            const int p = ids[(size_t)i * Cn + h];
            double v = part[((size_t)h * Np + p) * cells + e];
            if (cumulative && h > 0) v = __dsub_rn(v, part[((size_t)(h - 1) * Np + p) * cells + e]);
            acc = __dadd_rn(acc, v);
        }
        results[(size_t)i * cells + e] = acc;
    }
}

void launch_bootstrap_combine(const double *part, int Cn, int Np, size_t cells, int cumulative, const int32_t *ids, int N, double *results,
                              cudaStream_t st) {
    if (N <= 0) return;
    dim3 grid((unsigned)std::min<size_t>((cells + 255) / 256, 64), N);
    k_bootstrap_combine<<<grid, 256, 0, st>>>(part, Cn, Np, cells, cumulative, ids, N, results);
}

}  // namespace fgt
