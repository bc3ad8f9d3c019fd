/*
 * companion_oracle.c -- CPU restatement of the coancestry-curve path of
 * fastGLOBETROTTERCompanion.c.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle: only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may build, load or call it.
 * Nothing under fastglobetrotter_b200/ links against it and the product has no
 * CPU fallback.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py runs this restatement
 * and the compiled reference (oracle/_ref, built by oracle/build_ref.py from
 * /root/reference) on the tutorial data and on synthetic paintings with the same
 * libc srand() seed and requires bit-identical raw tensors, `means` and NULL tensors.
 * The reference ships no golden vectors for this path, so the
 * compiled reference itself is the anchor; tests/golden/ holds outputs generated from it.
 *
 * It works on *packed* inputs (decoded paintings + recombination maps), not on files:
 * the text/gz layer is exercised against the compiled reference directly.
 *
 * Third-party arithmetic restated here: glibc 2.39 rand()
 * (TYPE_3 additive feedback generator, random_r.c) -- see orc_srand/orc_rand; exp() and
 * pow() are taken from the system libm exactly as the reference does.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared -o liboracle.so companion_oracle.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* glibc rand(): r[i] = r[i-31] + r[i-3] (mod 2^32), output r[i] >> 1.        */
/* Seeding per glibc srandom_r: r[0]=seed (0 -> 1), r[i] = 16807*r[i-1] mod   */
/* (2^31-1) via Schrage, then 310 outputs discarded.  (reference C:136-137     */
/* calls rand(); SURVEY section 8c verified the recurrence for 10^6 draws.)    */
/* ------------------------------------------------------------------------- */
typedef struct {
    uint32_t h[31]; /* circular history; h[pos] is the oldest word x_{n-31} */
    int pos;
    int64_t drawn;
} orc_rand_t;

static inline uint32_t orc_step(orc_rand_t *g)
{
    int p = g->pos;
    int q = p + 28;
    if (q >= 31) q -= 31;
    uint32_t x = g->h[p] + g->h[q]; /* x_{n-31} + x_{n-3} */
    g->h[p] = x;
    g->pos = (p + 1 == 31) ? 0 : p + 1;
    g->drawn++;
    return x;
}

int orc_rand(orc_rand_t *g) { return (int)(orc_step(g) >> 1); }

void orc_srand(orc_rand_t *g, unsigned seed)
{
    int32_t r[34];
    if (seed == 0) seed = 1;
    r[0] = (int32_t)seed;
    for (int i = 1; i < 31; i++) {
        long hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
        long w = 16807 * lo - 2836 * hi;
        if (w < 0) w += 2147483647;
        r[i] = (int32_t)w;
    }
    /* glibc table layout: fptr = &r[3], rptr = &r[0]; logical "oldest" word is r[3] */
    for (int i = 0; i < 31; i++) g->h[i] = (uint32_t)r[(3 + i) % 31];
    g->pos = 0;
    g->drawn = 0;
    for (int i = 0; i < 310; i++) orc_step(g);
    g->drawn = 0;
}

/* set the state from 31 history words, oldest first (used to mirror libc's live state) */
void orc_rand_set_history(orc_rand_t *g, const uint32_t *hist)
{
    memcpy(g->h, hist, sizeof g->h);
    g->pos = 0;
    g->drawn = 0;
}

void orc_rand_get_history(const orc_rand_t *g, uint32_t *hist)
{
    for (int i = 0; i < 31; i++) hist[i] = g->h[(g->pos + i) % 31];
}

size_t orc_rand_sizeof(void) { return sizeof(orc_rand_t); }

/* ------------------------------------------------------------------------- */
/* cM prefix of a recombination map (reference C:494: sequential fp64 sum;    */
/* the increment is re-used at C:501 for the chunk end).                      */
/* ------------------------------------------------------------------------- */
void orc_cm_prefix(int nsites, const double *pos, const double *rate, double *cum, double *inc)
{
    double total = 0.0;
    cum[0] = 0.0;
    inc[0] = 0.0;
    for (int j = 1; j < nsites; j++) {
        double step = rate[j - 1] * (pos[j] - pos[j - 1]) * 100;
        total = total + step;
        cum[j] = total;
        inc[j] = step;
    }
}

static inline int label_at(const void *labels, int bytes, size_t idx)
{
    return bytes == 2 ? (int)((const uint16_t *)labels)[idx] : ((const int32_t *)labels)[idx];
}

typedef struct {
    double *pos, *size, *weight, *end;
    int *pop;
    int n, cap;
} chunk_list;

static void cl_reserve(chunk_list *c, int need)
{
    if (need <= c->cap) return;
    int cap = c->cap ? c->cap : 1024;
    while (cap < need) cap *= 2;
    c->pos = realloc(c->pos, cap * sizeof(double));
    c->size = realloc(c->size, cap * sizeof(double));
    c->weight = realloc(c->weight, cap * sizeof(double));
    c->end = realloc(c->end, cap * sizeof(double));
    c->pop = realloc(c->pop, cap * sizeof(int));
    c->cap = cap;
}

static void cl_free(chunk_list *c)
{
    free(c->pos); free(c->size); free(c->weight); free(c->end); free(c->pop);
    memset(c, 0, sizeof *c);
}

/*
 * Run-length encode one painting into chunks and append them to `out`
 * (reference C:446-455 counts, C:487-531 builds; same in mode 3 C:1447-1519 and in the
 * NULL path C:897-962).  A new chunk opens at a label change only when the run that just
 * ended was longer than weight_min SNPs; the run counter restarts at every label change.
 * Returns 0, or -1 if a chunk's donor haplotype has population label -9 (reference: exit(1)).
 */
static int chunk_painting(const void *labels, int bytes, size_t row_off, int nsites, const double *cum,
                          const double *inc, double weight_min, double cutoff, const int *donor_label_vec,
                          chunk_list *out)
{
    double start = 0.0;
    int hap = label_at(labels, bytes, row_off);
    double run = 1;
    for (int j = 1; j <= nsites; j++) {
        int boundary = 0;
        if (j < nsites) {
            int a = label_at(labels, bytes, row_off + j), b = label_at(labels, bytes, row_off + j - 1);
            if (a == b) run = run + 1;
            else {
                if (run > weight_min) boundary = 1;
                run = 1;
            }
        }
        if (boundary || j == nsites) {
            double end = (j == nsites) ? cum[nsites - 1] : cum[j] - inc[j];
            int pop = donor_label_vec[hap - 1];
            if (pop == -9) return -1;
            cl_reserve(out, out->n + 1);
            int c = out->n++;
            out->pos[c] = (start + end) / 2.0;
            out->end[c] = end;
            double len = end - start;
            out->size[c] = len;
            out->weight[c] = (len > cutoff) ? cutoff : len;
            out->pop[c] = pop;
            if (j < nsites) {
                start = cum[j];
                hap = label_at(labels, bytes, row_off + j);
            }
        }
    }
    return 0;
}

/* fine-bin rounding rule of reference C:144-147 (two-sided test, not round()) */
static inline int fine_bin(double d, double bw, int grid_start, int *ok)
{
    double q = d / bw;
    double fl = floor(q);
    double frac = (q - fl) * bw;
    int b = 0;
    *ok = 0;
    if (frac <= bw / 2.0) { b = (int)fl; *ok = 1; }
    if (frac > bw / 2.0) { b = (int)fl + 1; *ok = 1; }
    return b - grid_start;
}

/*
 * Sampled pair accumulation for one (individual, chromosome): reference
 * tabulate_chunks C:47-181 (mode 1: window +3, Y/10, label-0 pairs skipped) and
 * tabulate_chunks_mode3 C:1059-1188 (window +1, Y/8, no label-0 skip).
 * counts is K*K*G, row (lo*K+hi), only lo<=hi written.  Returns the number of sampled
 * pairs (sum of Y), or -1 on a label > K (reference C:102-107 exits).
 */
static long tabulate(const chunk_list *cl, int mode, double bw, int G, int grid_start, int K, orc_rand_t *rng,
                     double *counts)
{
    int n = cl->n;
    double chr_length = 0.0;
    for (int i = 0; i < n; i++)
        if (cl->end[i] >= chr_length) chr_length = cl->end[i];
    int nb = (int)chr_length / 1 + 1;
    int *t = calloc(nb, sizeof(int)), *first = malloc(nb * sizeof(int)), *fill = malloc(nb * sizeof(int));
    for (int i = 0; i < n; i++) t[(int)cl->pos[i]]++;
    first[0] = 0;
    for (int b = 1; b < nb; b++) first[b] = first[b - 1] + t[b - 1];
    memcpy(fill, first, nb * sizeof(int));
    /* stable counting sort by coarse 1-cM bin == the reference's O(nb*n) double loop C:94-111 */
    double *spos = malloc(n * sizeof(double)), *ssize = malloc(n * sizeof(double)), *sw = malloc(n * sizeof(double));
    int *spop = malloc(n * sizeof(int));
    long total = 0;
    int bad = 0;
    for (int i = 0; i < n; i++) {
        int d = fill[(int)cl->pos[i]]++;
        spos[d] = cl->pos[i]; ssize[d] = cl->size[i]; sw[d] = cl->weight[i]; spop[d] = cl->pop[i];
        if (mode == 1 && cl->pop[i] > K) bad = 1;
    }
    if (bad) { total = -1; goto done; }
    int window = (int)ceil(bw * G) + (mode == 3 ? 1 : 3);
    double divisor = (mode == 3) ? 8 : 10;
    for (int j = 0; j < nb - 1; j++) {
        int klim = j + window > nb ? nb : j + window;
        for (int k = j + 1; k < klim; k++) {
            int t1 = t[j], t2 = t[k];
            int Y = (int)t1 * t2 * exp(-0.05 * (k - j)) / divisor;
            if (t1 > 0 && t2 > 0 && Y == 0) Y = 1;
            total += Y;
            for (int it = 0; it < Y; it++) {
                int c1 = first[j] + orc_rand(rng) % t1;
                int c2 = first[k] + orc_rand(rng) % t2;
                if (mode != 3 && (spop[c2] == 0 || spop[c1] == 0)) continue;
                double d = pow(pow(spos[c1] - spos[c2], 2.0), 0.5);
                int ok;
                int b = fine_bin(d, bw, grid_start, &ok);
                double mind = ssize[c1] / 2.0 + ssize[c2] / 2.0;
                if (ok && b >= 0 && b < G && d >= mind) {
                    int l1 = spop[c1] - 1, l2 = spop[c2] - 1;
                    if (l1 < 0 || l2 < 0) continue; /* mode 3 with a target-label chunk: UB in the reference */
                    size_t cell = (l1 > l2) ? ((size_t)l2 * K + l1) : ((size_t)l1 * K + l2);
                    double prod = (l1 > l2) ? sw[c2] * sw[c1] : sw[c1] * sw[c2];
                    counts[cell * G + b] = counts[cell * G + b] + prod;
                }
            }
        }
    }
done:
    free(t); free(first); free(fill); free(spos); free(ssize); free(sw); free(spop);
    return total;
}

/* projection onto surrogate pairs: reference C:578-595 (mode 1) / C:1190-1203 (mode 3).
 * Both walk the K(K+1)/2 upper-triangle rows in (k,h) order for every output cell, so per
 * output cell the fp64 sum order is identical. */
static void project(const double *counts, int K, int S, int G, const double *tw, double *results)
{
    for (int k = 0; k < K; k++)
        for (int h = k; h < K; h++) {
            const double *row = counts + ((size_t)k * K + h) * G;
            for (int m = 0; m < S; m++)
                for (int n = 0; n < S; n++) {
                    double w = tw[((size_t)K * k + h) * S * S + (size_t)S * m + n];
                    double *o = results + ((size_t)m * S + n) * G;
                    for (int i = 0; i < G; i++) o[i] = o[i] + w * row[i];
                }
        }
}

/* symmetrise, divide by expectation, average over individuals: reference C:599-648 */
static void normalise_add(double *res /* S*S*G, modified */, int S, int G, int N, double *means)
{
    double *rs = malloc(S * sizeof(double)), *ex = malloc((size_t)S * S * sizeof(double));
    for (int i = 0; i < G; i++) {
        for (int k = 0; k < S; k++)
            for (int h = k; h < S; h++)
                res[((size_t)k * S + h) * G + i] = (res[((size_t)k * S + h) * G + i] + res[((size_t)h * S + k) * G + i]) / 2.0;
        for (int k = 0; k < S - 1; k++)
            for (int h = k + 1; h < S; h++) res[((size_t)h * S + k) * G + i] = res[((size_t)k * S + h) * G + i];
        double tot = 0.0;
        for (int k = 0; k < S; k++) {
            rs[k] = 0.0;
            for (int h = 0; h < S; h++) rs[k] = rs[k] + res[((size_t)k * S + h) * G + i];
            tot = tot + rs[k];
        }
        for (int k = 0; k < S; k++)
            for (int h = 0; h < S; h++) ex[k * S + h] = rs[k] * rs[h] / tot;
        for (int k = 0; k < S; k++)
            for (int h = k; h < S; h++) ex[k * S + h] = (ex[k * S + h] + ex[h * S + k]) / 2.0;
        for (int k = 0; k < S - 1; k++)
            for (int h = k + 1; h < S; h++) ex[h * S + k] = ex[k * S + h];
        for (int k = 0; k < S; k++)
            for (int h = 0; h < S; h++) {
                size_t o = ((size_t)k * S + h) * G + i;
                means[o] = means[o] + (res[o] / ex[k * S + h]) / N;
            }
    }
    free(rs); free(ex);
}

/*
 * The reference walks the samples file forwards (C:418-425, C:553-560): the first
 * pseudo-individual skips to file individual id[0]; afterwards it skips id[n]-id[n-1]-1
 * individuals, and a negative skip means "tabulate the previously read individual again".
 * This returns, for every pseudo-individual, the file individual actually used.
 */
static void resolve_file_walk(const int *ids /* stride */, int stride, int N, int *used)
{
    int next_in_file = 0;
    for (int n = 0; n < N; n++) {
        int id = ids[(size_t)n * stride];
        if (n == 0) { used[0] = id; next_in_file = id + 1; continue; }
        int skip = id - ids[(size_t)(n - 1) * stride] - 1;
        if (skip < 0) used[n] = used[n - 1];
        else { used[n] = next_in_file + skip; next_in_file = used[n] + 1; }
    }
}

/*
 * data_read (mode 1) / data_read_mode3 (mode 3) on packed inputs.
 *   ind_rows[i*num_chrom + h]  index (into the packed label matrix of chromosome h, in units of
 *                              individuals) used for pseudo-individual i -- i.e. the reference's
 *                              ind_id_vec_new (C:276-283) restricted to packed rows.
 *   labels[h]                  [n_packed_inds * ploidy * nsamples][nsites[h]] donor haplotype ids
 *   raw (optional)             mode 1: [N][K*K][G] total_counts_mat_all as it stands before C:578
 *                              mode 3: [num_chrom*N][K*K][G] per tabulate call, in call order
 * Returns total sampled pairs, or a negative error code where the reference would exit(1).
 */
long orc_data_read_packed(int mode, int ploidy, int nsamples, int num_chrom, int N, const int *ind_rows,
                          const int *nsites, const double *const *pos, const double *const *rate,
                          const void *const *labels, int label_bytes, double bw, int G, int grid_start, int K,
                          const int *donor_label_vec, double weight_min, double cutoff, int S, const double *tw,
                          orc_rand_t *rng, double *means, double *raw)
{
    size_t tens = (size_t)K * K * G, rsz = (size_t)S * S * G;
    double *counts = NULL, *scratch = NULL;
    double *results = calloc((size_t)N * rsz, sizeof(double));
    if (mode == 1) counts = calloc((size_t)N * tens, sizeof(double));
    else scratch = malloc(tens * sizeof(double));
    int *used = malloc(N * sizeof(int));
    long pairs = 0, call = 0;
    chunk_list cl = {0};
    for (int h = 0; h < num_chrom; h++) {
        int L = nsites[h];
        double *cum = malloc(L * sizeof(double)), *inc = malloc(L * sizeof(double));
        orc_cm_prefix(L, pos[h], rate[h], cum, inc);
        resolve_file_walk(ind_rows + h, num_chrom, N, used);
        for (int n = 0; n < N; n++) {
            if (n == 0 || used[n] != used[n - 1] ) {
                cl.n = 0;
                for (int p = 0; p < ploidy; p++)
                    for (int s = 0; s < nsamples; s++) {
                        size_t row = ((size_t)used[n] * ploidy + p) * nsamples + s;
                        if (chunk_painting(labels[h], label_bytes, row * L, L, cum, inc, weight_min, cutoff,
                                           donor_label_vec, &cl) < 0) { pairs = -2; goto out; }
                    }
            }
            double *dst = (mode == 1) ? counts + (size_t)n * tens : scratch;
            if (mode == 3) memset(scratch, 0, tens * sizeof(double));
            long y = tabulate(&cl, mode, bw, G, grid_start, K, rng, dst);
            if (y < 0) { pairs = -3; goto out; }
            pairs += y;
            if (mode == 3) {
                if (raw) memcpy(raw + (size_t)call * tens, scratch, tens * sizeof(double));
                project(scratch, K, S, G, tw, results + (size_t)n * rsz);
            }
            call++;
        }
        free(cum); free(inc);
    }
    if (mode == 1) {
        if (raw) memcpy(raw, counts, (size_t)N * tens * sizeof(double));
        for (int n = 0; n < N; n++) project(counts + (size_t)n * tens, K, S, G, tw, results + (size_t)n * rsz);
    }
    for (int n = 0; n < N; n++) normalise_add(results + (size_t)n * rsz, S, G, N, means);
out:
    cl_free(&cl);
    free(counts); free(scratch); free(results); free(used);
    return pairs;
}

/*
 * data_read_null on packed inputs: reference C:865-1049.  Only the first painting sample of each
 * haplotype is used (nsamplestoconsiderNULL = 1, C:748) and only the first min(N,100) recipients
 * (C:687, C:704-705).  Exhaustive chunk x chunk between haplotypes of different individuals, full
 * (unsymmetrised) K x K x G tensor, one global sequential fp64 fold.
 *   rec_rows[r]  packed individual index of the r-th recipient.
 * Returns the number of evaluated chunk pairs, negative on error.
 */
long orc_data_read_null_packed(int ploidy, int nsamples, int num_chrom, int N, const int *rec_rows, const int *nsites,
                               const double *const *pos, const double *const *rate, const void *const *labels,
                               int label_bytes, double bw, int G, int grid_start, int K, const int *donor_label_vec,
                               double weight_min, double cutoff, double *chrom_ind_res)
{
    int Nc = N > 100 ? 100 : N;
    long evals = 0;
    chunk_list *haps = calloc((size_t)Nc * ploidy, sizeof(chunk_list));
    for (int h = 0; h < num_chrom; h++) {
        int L = nsites[h];
        double *cum = malloc(L * sizeof(double)), *inc = malloc(L * sizeof(double));
        orc_cm_prefix(L, pos[h], rate[h], cum, inc);
        for (int n = 0; n < Nc; n++)
            for (int p = 0; p < ploidy; p++) {
                chunk_list *c = &haps[n * ploidy + p];
                c->n = 0;
                size_t row = ((size_t)rec_rows[n] * ploidy + p) * nsamples + 0;
                /* the reference only validates label -9 for chunks it actually visits (A: n < Nc-1; B: g > n);
                 * with Nc >= 2 every haplotype is visited, with Nc == 1 none is */
                if (chunk_painting(labels[h], label_bytes, row * L, L, cum, inc, weight_min, cutoff, donor_label_vec, c) < 0
                    && Nc >= 2) { evals = -2; goto out; }
            }
        for (int n = 0; n < Nc - 1; n++)
            for (int p = 0; p < ploidy; p++) {
                const chunk_list *A = &haps[n * ploidy + p];
                for (int g = n + 1; g < Nc; g++)
                    for (int m = 0; m < ploidy; m++) {
                        const chunk_list *B = &haps[g * ploidy + m];
                        for (int i = 0; i < A->n; i++)
                            for (int j = 0; j < B->n; j++) {
                                double d = pow(pow(A->pos[i] - B->pos[j], 2.0), 0.5);
                                int ok;
                                int b = fine_bin(d, bw, grid_start, &ok);
                                double mind = A->size[i] / 2.0 + B->size[j] / 2.0;
                                if (ok && b >= 0 && b < G && d >= mind && A->pop[i] > 0 && B->pop[j] > 0) {
                                    size_t o = ((size_t)(A->pop[i] - 1) * K + (B->pop[j] - 1)) * G + b;
                                    chrom_ind_res[o] = chrom_ind_res[o] + A->weight[i] * B->weight[j];
                                }
                            }
                        evals += (long)A->n * B->n;
                    }
            }
        free(cum); free(inc);
    }
out:
    for (int i = 0; i < Nc * ploidy; i++) cl_free(&haps[i]);
    free(haps);
    return evals;
}

/* reference C:15-23 */
void orc_mutiply_temp_weight(double *tempweights, const double *weights_mat, int S, int K)
{
    for (int k = 0; k < K; k++)
        for (int l = 0; l < K; l++)
            for (int m = 0; m < S; m++)
                for (int n = 0; n < S; n++)
                    tempweights[((size_t)m * S + n) * K * K + (size_t)k * K + l] = weights_mat[k * S + m] * weights_mat[l * S + n];
}

/* chunk one painting into caller arrays (used by tests of the device chunk builder) */
int orc_chunk_one(const void *labels, int label_bytes, int nsites, const double *pos, const double *rate,
                  double weight_min, double cutoff, const int *donor_label_vec, int cap, double *cpos, double *csize,
                  double *cweight, double *cend, int *cpop)
{
    double *cum = malloc(nsites * sizeof(double)), *inc = malloc(nsites * sizeof(double));
    orc_cm_prefix(nsites, pos, rate, cum, inc);
    chunk_list cl = {0};
    int rc = chunk_painting(labels, label_bytes, 0, nsites, cum, inc, weight_min, cutoff, donor_label_vec, &cl);
    int n = rc < 0 ? rc : cl.n;
    if (rc == 0 && cl.n <= cap) {
        memcpy(cpos, cl.pos, cl.n * sizeof(double)); memcpy(csize, cl.size, cl.n * sizeof(double));
        memcpy(cweight, cl.weight, cl.n * sizeof(double)); memcpy(cend, cl.end, cl.n * sizeof(double));
        memcpy(cpop, cl.pop, cl.n * sizeof(int));
    }
    cl_free(&cl); free(cum); free(inc);
    return n;
}
