/* oracle/poa_cluster.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * abPOA's de-novo read clustering for max_n_cons = 2 (abpoa_multip_read_clu_kmedoids, abPOA/src/abpoa_output.c:1135-1180) and the
 * per-cluster most-frequent consensus (abpoa_most_frequent :549-586 with n_clu = 2), restated on the row-column MSA: every quantity
 * the reference takes from the graph here (a node's out-edge read sets restricted to a cluster, :345-352) equals a count over the MSA
 * rows of the cluster's reads, because a read's MSA row holds a node's base exactly when the read is in one of the node's out-edge sets
 * (abpoa_set_msa_seq) and the nodes of one column have distinct bases.  longcallD reaches it through abpoa_aln_msa_cons
 * (src/align.c:872-953: wb = -1, sub_aln = 0, max_n_cons = 2, min_freq = opt->min_af) for regions without a usable phase set.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <limits.h>
#include "lcd_oracle.h"

typedef struct {
    int pos, depth, var_type, count, n_uniq;
    int map[6];                 /* base code (0..4, 5 = gap) -> allele index in order of first appearance, or -1 */
} het_t;

/* read r's allele index at candidate k (cand_het_pos_t.read_id_to_allele_idx) */
static inline int alle(const uint8_t *msa, int ml, const het_t *h, int r) { return h->map[msa[(size_t)r * ml + h->pos]]; }

/* abpoa_collect_cand_het_pos (:676-800): candidate columns, identical read partitions merged with a count, priority order */
static int collect_het(const uint8_t *msa, int n_seq, int ml, int min_w, het_t *het, int *prio) {
    const int m = 5;
    int min_het = min_w / 2 > 2 ? min_w / 2 : 2;
    const int min_hom = n_seq - min_het;
    int n_het = 0;
    for (int i = 0; i < ml; ++i) {
        int depth[6] = {0, 0, 0, 0, 0, 0}, first[6] = {0, 0, 0, 0, 0, 0};
        for (int j = 0; j < n_seq; ++j) { const int b = msa[(size_t)j * ml + i]; if (++depth[b] == 1) first[b] = j; }
        int alleles[6], nu = 0, var_type = 0, total = 0;
        for (int j = 0; j < m + 1; ++j) if (depth[j] >= min_het && depth[j] <= min_hom) { alleles[nu++] = j; total += depth[j]; if (j == m) var_type = 1; }
        if (nu < 2) continue;
        for (int j = 0; j < nu - 1; ++j) for (int k = j + 1; k < nu; ++k) if (first[alleles[j]] > first[alleles[k]]) { const int t = alleles[j]; alleles[j] = alleles[k]; alleles[k] = t; }
        het_t h; h.pos = i; h.depth = total; h.var_type = var_type; h.count = 1; h.n_uniq = nu;
        for (int j = 0; j < 6; ++j) h.map[j] = -1;
        for (int j = 0; j < nu; ++j) h.map[alleles[j]] = j;
        /* allele_clu_exist (:647-674): an earlier candidate with the same ordered read partition */
        int same = -1;
        for (int k = n_het - 1; k >= 0 && same < 0; --k) {
            if (het[k].n_uniq != nu) continue;
            int eq = 1;
            for (int r = 0; r < n_seq && eq; ++r) if (alle(msa, ml, &het[k], r) != alle(msa, ml, &h, r)) eq = 0;
            if (eq) same = k;
        }
        if (same >= 0) { het[same].count++; if (var_type == 0) het[same].var_type = 0; continue; }
        het[n_het++] = h;
    }
    /* bubble sort by (count desc, depth desc, var_type asc) (:772-791): a stable sort */
    for (int i = 0; i < n_het; ++i) prio[i] = i;
    int swapped;
    do {
        swapped = 0;
        for (int j = 0; j < n_het - 1; ++j) {
            const het_t *a = &het[prio[j]], *b = &het[prio[j + 1]];
            if (a->count < b->count || (a->count == b->count && a->depth < b->depth) || (a->count == b->count && a->depth == b->depth && a->var_type > b->var_type)) {
                const int t = prio[j]; prio[j] = prio[j + 1]; prio[j + 1] = t; swapped = 1;
            }
        }
    } while (swapped);
    return n_het;
}

/* Returns the number of clusters (1 or 2); read_clu[r] = cluster of read r (all 0 for one cluster).  min_w = MAX(2, ceil(n_seq * min_freq)). */
int lcd_oracle_poa_cluster(const uint8_t *msa, int n_seq, int ml, int min_w, uint8_t *read_clu) {
    memset(read_clu, 0, (size_t)n_seq);
    if (ml <= 0 || n_seq <= 0) return 1;
    het_t *het = (het_t*)malloc((size_t)ml * sizeof(het_t)); int *prio = (int*)malloc((size_t)ml * sizeof(int));
    const int n_het = collect_het(msa, n_seq, ml, min_w, het, prio);
    int n_clu = 1;
    if (n_het >= 1) {
        /* abpoa_collect_msa_dis_matrix (:802-862): SNP columns weigh 2, gap columns 1, times the partition's count */
        int *dis = (int*)calloc((size_t)n_seq * n_seq, sizeof(int));
        for (int i = 0; i < n_seq; ++i) for (int j = i + 1; j < n_seq; ++j) {
            int d = 0;
            for (int k = 0; k < n_het; ++k) {
                const int a = alle(msa, ml, &het[k], i), b = alle(msa, ml, &het[k], j);
                if (a >= 0 && b >= 0 && a != b) d += (het[k].var_type == 0 ? 2 : 1) * het[k].count;
            }
            dis[i * n_seq + j] = dis[j * n_seq + i] = d;
        }
        /* abpoa_init_kmedoids / abpoa_collect_multi_medoids / abpoa_collect_2medoids (:865-1003) with max_n_cons = 2: the first candidate in
         * priority order that has two reads of different alleles at a positive distance gives the farthest such pair */
        int med[2] = {-1, -1}, have = 0;
        for (int t = 0; t < n_het && !have; ++t) {
            const het_t *h = &het[prio[t]];
            int max_dis = 0;
            for (int a = 0; a < h->n_uniq - 1; ++a) for (int b = a + 1; b < h->n_uniq; ++b)
                for (int r1 = 0; r1 < n_seq; ++r1) { if (alle(msa, ml, h, r1) != a) continue;
                    for (int r2 = 0; r2 < n_seq; ++r2) { if (alle(msa, ml, h, r2) != b) continue;
                        if (dis[r1 * n_seq + r2] > max_dis) { max_dis = dis[r1 * n_seq + r2]; med[0] = r1; med[1] = r2; } } }
            if (max_dis > 0) have = 1;
        }
        if (have) {
            /* abpoa_update_kmedoids (:1030-1093), at most 10 rounds (:1105-1109) */
            int *clu = (int*)malloc((size_t)2 * n_seq * sizeof(int)), ncs[2] = {0, 0};
            for (int iter = 0; iter < 10; ++iter) {
                ncs[0] = ncs[1] = 0;
                for (int i = 0; i < n_seq; ++i) {
                    int min_dis = INT_MAX, min_clu = -1, tied = 0;
                    for (int j = 0; j < 2; ++j) {
                        const int d = dis[i * n_seq + med[j]];
                        if (d < min_dis) { min_dis = d; min_clu = j; tied = 0; } else if (d == min_dis) tied = 1;
                    }
                    if (tied) min_clu = ncs[0] < ncs[1] ? 0 : 1;
                    clu[min_clu * n_seq + ncs[min_clu]++] = i;
                }
                int nm[2] = {-1, -1};
                for (int c = 0; c < 2; ++c) {
                    int best = INT_MAX;
                    for (int j = 0; j < ncs[c]; ++j) {
                        int s = 0; const int ri = clu[c * n_seq + j];
                        for (int k = 0; k < ncs[c]; ++k) if (k != j) s += dis[ri * n_seq + clu[c * n_seq + k]];
                        if (s < best) { best = s; nm[c] = ri; }
                    }
                }
                if (nm[0] > nm[1]) { const int t = nm[0]; nm[0] = nm[1]; nm[1] = t; }
                int changed = 0;
                for (int c = 0; c < 2; ++c) { if (nm[c] == -1) { changed = 0; break; } if (nm[c] != med[c]) changed = 1; }
                med[0] = nm[0]; med[1] = nm[1];
                if (!changed) break;
            }
            /* :1110-1121: both clusters need min_w reads (every read is in one of them, so the 80 % rule holds) */
            if (ncs[0] >= min_w && ncs[1] >= min_w) {
                n_clu = 2;
                for (int j = 0; j < ncs[1]; ++j) read_clu[clu[n_seq + j]] = 1;
            }
            free(clu);
        }
        free(dis);
    }
    free(het); free(prio);
    return n_clu;
}

/* abpoa_most_frequent with clusters, sub_aln = 0 (:393-424,549-586): cluster c's consensus from the MSA rows of its reads; the consensus
 * rows n_seq + c of the MSA are written too (abpoa_generate_rc_msa :177-191).  cons gets the consensus sequences back to back. */
void lcd_oracle_poa_cluster_cons(uint8_t *msa, int n_seq, int ml, int n_clu, const uint8_t *read_clu, uint8_t *cons, int32_t *cons_len) {
    int csz[2] = {0, 0};
    for (int r = 0; r < n_seq; ++r) csz[n_clu == 1 ? 0 : read_clu[r]]++;
    int at = 0;
    for (int c = 0; c < n_clu; ++c) {
        uint8_t *row = msa + (size_t)(n_seq + c) * ml;
        int cl = 0;
        for (int i = 0; i < ml; ++i) {
            int cnt[4] = {0, 0, 0, 0};
            for (int r = 0; r < n_seq; ++r) if (n_clu == 1 || read_clu[r] == c) { const int b = msa[(size_t)r * ml + i]; if (b < 4) cnt[b]++; }
            int max_c = 0, total = 0, max_base = 5;
            for (int j = 0; j < 4; ++j) { if (cnt[j] > max_c) { max_c = cnt[j]; max_base = j; } total += cnt[j]; }
            row[i] = 5;
            if (max_c >= csz[c] - total) { cons[at + cl++] = (uint8_t)max_base; row[i] = (uint8_t)max_base; }
        }
        cons_len[c] = cl; at += cl;
    }
}
