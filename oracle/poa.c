/* oracle/poa.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Scalar restatement of abPOA (vendored @6ee6279) exactly as longcallD drives it for one
 * (region, haplotype) consensus (src/align.c:762-870 abpoa_partial_aln_msa_cons and :872-953
 * abpoa_aln_msa_cons, full-cover reads): progressive sequence-to-graph alignment with the
 * convex-gap adaptive-banded DP, graph fusion, topological re-sort after every read, row-column
 * MSA and most-frequent consensus.
 *
 * Reference code restated here:
 *   simd_abpoa_align_sequence_to_subgraph   abPOA/src/abpoa_align_simd.c:1250-1332
 *   simd_abpoa_cg_first_dp / simd_abpoa_cg_dp / SIMD_SET_F          :669-688 / :935-1074 / :691-725
 *   simd_abpoa_max_in_row / _ada_max_i / _global_get_max            :1107-1130 / :1092-1105
 *   simd_abpoa_cg_backtrack                                          :309-458
 *   abpoa_get_incre_path_score                                       abpoa_graph.c:429-437
 *   abpoa_add_subgraph_alignment / abpoa_add_graph_edge / _sequence  :689-774 / :480-556 / :573-593
 *   abpoa_topological_sort (BFS index, edge sort, BFS remain)        :192-357
 *   abpoa_DFS_set_msa_rank                                           :359-410
 *   abpoa_generate_rc_msa / abpoa_set_msa_seq                        abpoa_output.c:149-192 / :105-123
 *   abpoa_most_frequent / _set_row_column_weight / _set_major_voting_cons   :549-586 / :451-475 / :393-424
 *
 * The reference DP is SIMD code whose results depend on vector granularity (band snapping to
 * vector indices, the 2-then-1 lane propagation of F beyond the predecessors' last vector, wrapping
 * int16 adds).  We restate the AVX-512BW build (32 int16 lanes, log_n = 5): a "vector" below is 32
 * consecutive int16 cells and every vector operation of the reference is spelled out per lane.
 * Problems that would need the int32 cell path (abpoa_align_simd.c:1293-1302) are rejected.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "lcd_oracle.h"

#define PN 32
#define LOGN 5
typedef int16_t sc_t;
static inline sc_t wadd(sc_t a, sc_t b) { return (sc_t)(uint16_t)((uint16_t)a + (uint16_t)b); }
static inline sc_t wsub(sc_t a, sc_t b) { return (sc_t)(uint16_t)((uint16_t)a - (uint16_t)b); }
static inline sc_t smax(sc_t a, sc_t b) { return a > b ? a : b; }
#define MAX2(a,b) ((a)>(b)?(a):(b))
#define MIN2(a,b) ((a)<(b)?(a):(b))

typedef struct {
    int base;
    int in_n, in_m, *in_id, *in_w;
    int out_n, out_m, *out_id, *out_w; uint64_t *rid;   /* rid[edge * rid_n + word] */
    int n_read, n_span;
    int aln_n, aln_m, *aln;
} pnode_t;

typedef struct {
    pnode_t *node; int n, m;
    int rid_n;
    int *idx2id, *id2idx, *maxl, *maxr, *remain, *msa_rank; int aux_m;
    int msa_ranked;
} pgraph_t;

/* ------------------------------------------------------------------ graph */
static int g_add_node(pgraph_t *g, int base) {
    if (g->n == g->m) {
        int m = g->m ? g->m * 2 : 64;
        g->node = (pnode_t*)realloc(g->node, (size_t)m * sizeof(pnode_t));
        memset(g->node + g->m, 0, (size_t)(m - g->m) * sizeof(pnode_t));
        g->m = m;
    }
    g->node[g->n].base = base;
    return g->n++;
}
static void g_free(pgraph_t *g) {
    for (int i = 0; i < g->m; ++i) {
        free(g->node[i].in_id); free(g->node[i].in_w); free(g->node[i].out_id); free(g->node[i].out_w);
        free(g->node[i].rid); free(g->node[i].aln);
    }
    free(g->node); free(g->idx2id); free(g->id2idx); free(g->maxl); free(g->maxr); free(g->remain); free(g->msa_rank);
}
/* abpoa_add_graph_edge, abpoa_graph.c:480-556 (weights are always 1: use_qv == 0) */
static void g_add_edge(pgraph_t *g, int from, int to, int check, int w, int add_rid, int read_id) {
    pnode_t *f = &g->node[from], *t = &g->node[to];
    int exist = 0, oi = -1;
    if (check) {
        for (int i = 0; i < t->in_n; ++i) if (t->in_id[i] == from) { t->in_w[i] += w; break; }
        for (int i = 0; i < f->out_n; ++i) if (f->out_id[i] == to) { f->out_w[i] += w; exist = 1; oi = i; break; }
    }
    if (!exist) {
        if (t->in_n == t->in_m) {
            t->in_m = t->in_m ? t->in_m * 2 : 4;
            t->in_id = (int*)realloc(t->in_id, (size_t)t->in_m * sizeof(int));
            t->in_w = (int*)realloc(t->in_w, (size_t)t->in_m * sizeof(int));
        }
        t->in_id[t->in_n] = from; t->in_w[t->in_n] = w; t->in_n++;
        if (f->out_n == f->out_m) {
            int m = f->out_m ? f->out_m * 2 : 4;
            f->out_id = (int*)realloc(f->out_id, (size_t)m * sizeof(int));
            f->out_w = (int*)realloc(f->out_w, (size_t)m * sizeof(int));
            f->rid = (uint64_t*)realloc(f->rid, (size_t)m * g->rid_n * sizeof(uint64_t));
            memset(f->rid + (size_t)f->out_m * g->rid_n, 0, (size_t)(m - f->out_m) * g->rid_n * sizeof(uint64_t));
            f->out_m = m;
        }
        oi = f->out_n;
        f->out_id[oi] = to; f->out_w[oi] = w; f->out_n++;
    }
    if (add_rid) f->rid[(size_t)oi * g->rid_n + read_id / 64] |= 1ULL << (read_id & 63);
    f->n_read += 1;
}
static void g_add_aligned1(pnode_t *n, int id) {
    if (n->aln_n == n->aln_m) { n->aln_m = n->aln_m ? n->aln_m * 2 : 4; n->aln = (int*)realloc(n->aln, (size_t)n->aln_m * sizeof(int)); }
    n->aln[n->aln_n++] = id;
}
/* abpoa_add_graph_aligned_node, abpoa_graph.c:456-464 */
static void g_add_aligned(pgraph_t *g, int node_id, int aligned_id) {
    for (int i = 0; i < g->node[node_id].aln_n; ++i) {
        int other = g->node[node_id].aln[i];
        g_add_aligned1(&g->node[other], aligned_id);
        g_add_aligned1(&g->node[aligned_id], other);
    }
    g_add_aligned1(&g->node[node_id], aligned_id);
    g_add_aligned1(&g->node[aligned_id], node_id);
}

/* abpoa_topological_sort, abpoa_graph.c:322-357 */
static void g_topo_sort(pgraph_t *g, int wb) {
    const int n = g->n;
    if (n > g->aux_m) {
        int m = n * 2;
        g->idx2id = (int*)realloc(g->idx2id, m * sizeof(int)); g->id2idx = (int*)realloc(g->id2idx, m * sizeof(int));
        g->maxl = (int*)realloc(g->maxl, m * sizeof(int)); g->maxr = (int*)realloc(g->maxr, m * sizeof(int));
        g->remain = (int*)realloc(g->remain, m * sizeof(int)); g->msa_rank = (int*)realloc(g->msa_rank, m * sizeof(int));
        g->aux_m = m;
    }
    int *deg = (int*)malloc((size_t)n * sizeof(int)), *q = (int*)malloc((size_t)(n + 1) * 2 * sizeof(int));
    int qh = 0, qt = 0, index = 0;
    /* abpoa_BFS_set_node_index :221-266 */
    for (int i = 0; i < n; ++i) deg[i] = g->node[i].in_n;
    q[qt++] = 0;
    while (qh < qt) {
        const int cur = q[qh++];
        g->idx2id[index] = cur; g->id2idx[cur] = index++;
        if (cur == 1) break;
        for (int i = 0; i < g->node[cur].out_n; ++i) {
            const int out = g->node[cur].out_id[i];
            if (--deg[out] == 0) {
                int ok = 1;
                for (int j = 0; j < g->node[out].aln_n; ++j) if (deg[g->node[out].aln[j]] != 0) { ok = 0; break; }
                if (!ok) continue;
                q[qt++] = out;
                for (int j = 0; j < g->node[out].aln_n; ++j) q[qt++] = g->node[out].aln[j];
            }
        }
    }
    /* abpoa_sort_in_out_ids :192-219 (exchange sort, descending weight) */
    for (int i = 0; i < n; ++i) {
        pnode_t *nd = &g->node[i];
        for (int j = 0; j < nd->in_n - 1; ++j) for (int k = j + 1; k < nd->in_n; ++k) if (nd->in_w[j] < nd->in_w[k]) {
            int t = nd->in_id[j]; nd->in_id[j] = nd->in_id[k]; nd->in_id[k] = t;
            t = nd->in_w[j]; nd->in_w[j] = nd->in_w[k]; nd->in_w[k] = t;
        }
        for (int j = 0; j < nd->out_n - 1; ++j) for (int k = j + 1; k < nd->out_n; ++k) if (nd->out_w[j] < nd->out_w[k]) {
            int t = nd->out_id[j]; nd->out_id[j] = nd->out_id[k]; nd->out_id[k] = t;
            t = nd->out_w[j]; nd->out_w[j] = nd->out_w[k]; nd->out_w[k] = t;
            for (int x = 0; x < g->rid_n; ++x) {
                uint64_t r = nd->rid[(size_t)j * g->rid_n + x]; nd->rid[(size_t)j * g->rid_n + x] = nd->rid[(size_t)k * g->rid_n + x]; nd->rid[(size_t)k * g->rid_n + x] = r;
            }
        }
    }
    if (wb >= 0) {
        for (int i = 0; i < n; ++i) { g->maxr[i] = 0; g->maxl[i] = n; }
        /* abpoa_BFS_set_node_remain :268-309 */
        for (int i = 0; i < n; ++i) { deg[i] = g->node[i].out_n; g->remain[i] = 0; }
        qh = qt = 0; q[qt++] = 1; g->remain[1] = -1;
        while (qh < qt) {
            const int cur = q[qh++];
            if (cur != 1) {
                int max_w = -1, max_id = -1;
                for (int i = 0; i < g->node[cur].out_n; ++i) if (g->node[cur].out_w[i] > max_w) { max_w = g->node[cur].out_w[i]; max_id = g->node[cur].out_id[i]; }
                g->remain[cur] = g->remain[max_id] + 1;
            }
            if (cur == 0) break;
            for (int i = 0; i < g->node[cur].in_n; ++i) { const int in = g->node[cur].in_id[i]; if (--deg[in] == 0) q[qt++] = in; }
        }
    }
    free(deg); free(q);
    g->msa_ranked = 0;
}

/* abpoa_get_incre_path_score, abpoa_graph.c:429-437 (k indexes in_id[] directly, as the caller does) */
static int path_score(const pgraph_t *g, int node_id, int k) {
    const int pre = g->node[node_id].in_id[k];
    int node_w = 0;
    for (int i = 0; i < g->node[pre].out_n; ++i) node_w += g->node[pre].out_w[i];
    const int edge_w = g->node[node_id].in_w[k];
    if (node_w == 0 || edge_w == 0) return 0;
    const int s = (int)round(log((double)edge_w / (double)node_w));
    return MAX2(s, -20);
}

/* ------------------------------------------------------------------ DP + backtrack */
typedef struct { int op, len, node_id; } gcig_t;   /* op: 0 M, 1 I, 2 D (ABPOA_C*) */
typedef struct { gcig_t *a; int n, m; } gcigar_t;
static void cig_push(gcigar_t *c, int op, int len, int node_id) {   /* abpoa_push_cigar, abpoa_align.h:54-73 */
    if (c->n == 0 || op != 1 || c->a[c->n - 1].op != 1) {
        if (c->n == c->m) { c->m = c->m ? c->m * 2 : 16; c->a = (gcig_t*)realloc(c->a, (size_t)c->m * sizeof(gcig_t)); }
        c->a[c->n].op = op; c->a[c->n].len = len; c->a[c->n].node_id = node_id; c->n++;
    } else c->a[c->n - 1].len += len;
}

/* F = max{F, (F-e)<<1, (F-2e)<<2, ...} : SIMD_SET_F, abpoa_align_simd.c:691-725 */
static void set_f(sc_t *F, int set_num, sc_t inf_min, int e) {
    sc_t t[PN];
    int cov = set_num;
    for (int s = 0; s < LOGN; ++s) {
        const int sh = 1 << s;
        const sc_t ge = (sc_t)(e << s);          /* GAP_ES[s] = GAP_ES[s-1] + GAP_ES[s-1] */
        if (set_num != PN && s > 0) cov += sh;
        for (int l = 0; l < PN; ++l) {
            sc_t v;
            if (l < sh) v = inf_min;                                       /* PRE_MIN[sh] */
            else if (set_num != PN && l > MIN2(cov, PN - 1)) v = inf_min;  /* PRE_MASK / SUF_MIN */
            else v = wsub(F[l - sh], ge);
            t[l] = v;
        }
        for (int l = 0; l < PN; ++l) F[l] = smax(F[l], t[l]);
    }
}

typedef struct {
    int wb; float wf; int match, mismatch, o1, e1, o2, e2; int inc_both_ends; int sub_aln;
} ppar_t;

/* returns 0 ok, -1 needs int32 cells */
static int poa_align(pgraph_t *g, const ppar_t *par, int beg_id, int end_id, const uint8_t *query, int qlen, gcigar_t *cig) {
    const int beg_index = g->id2idx[beg_id], end_index = g->id2idx[end_id];
    const int gn = end_index - beg_index + 1;
    const int oe1 = par->o1 + par->e1, oe2 = par->o2 + par->e2;
    const int len = qlen > gn ? qlen : gn;
    const int max_score = MAX2(qlen * par->match, len * par->e1 + par->o1);
    if (!(max_score <= INT16_MAX - par->mismatch - oe1 - oe2)) return -1;
    const sc_t inf_min = (sc_t)(MAX2(MAX2(INT16_MIN + par->mismatch, INT16_MIN + oe1), INT16_MIN + oe2) + 512 * MAX2(par->e1, par->e2));
    int mat[25];
    for (int i = 0; i < 4; ++i) { for (int j = 0; j < 4; ++j) mat[i * 5 + j] = i == j ? par->match : -par->mismatch; mat[i * 5 + 4] = 0; }
    for (int j = 0; j < 5; ++j) mat[20 + j] = 0;
    /* index_map :1259-1269 */
    uint8_t *index_map = (uint8_t*)calloc((size_t)g->n, 1);
    index_map[beg_index] = index_map[end_index] = 1;
    for (int i = beg_index; i < end_index - 1; ++i) {
        if (!index_map[i]) continue;
        const pnode_t *nd = &g->node[g->idx2id[i]];
        for (int j = 0; j < nd->out_n; ++j) index_map[g->id2idx[nd->out_id[j]]] = 1;
    }
    const int dp_sn = (qlen + 1 + PN - 1) / PN;
    const size_t row = (size_t)dp_sn * PN;
    sc_t *DP = (sc_t*)malloc((size_t)gn * 5 * row * sizeof(sc_t) + 64);
    memset(DP, 0x55, (size_t)gn * 5 * row * sizeof(sc_t));     /* the reference leaves unwritten cells undefined; never read */
    sc_t *qp = (sc_t*)malloc(5 * row * sizeof(sc_t));
    for (int k = 0; k < 5; ++k) { sc_t *p = qp + k * row; p[0] = 0; for (int j = 0; j < qlen; ++j) p[j + 1] = (sc_t)mat[k * 5 + query[j]]; for (size_t j = qlen + 1; j < row; ++j) p[j] = 0; }
    int *dp_beg = (int*)calloc((size_t)gn * 4, sizeof(int)), *dp_end = dp_beg + gn, *dp_beg_sn = dp_end + gn, *dp_end_sn = dp_beg_sn + gn;
    int **pre_index = (int**)calloc((size_t)gn, sizeof(int*)), *pre_n = (int*)calloc((size_t)gn, sizeof(int));
    for (int index_i = beg_index + 1, dp_i = 1; index_i <= end_index; ++index_i, ++dp_i) {
        const pnode_t *nd = &g->node[g->idx2id[index_i]];
        pre_index[dp_i] = (int*)malloc((size_t)MAX2(nd->in_n, 1) * sizeof(int));
        int pn_ = 0;
        for (int j = 0; j < nd->in_n; ++j) { const int pi = g->id2idx[nd->in_id[j]]; if (index_map[pi]) pre_index[dp_i][pn_++] = pi - beg_index; }
        pre_n[dp_i] = pn_;
    }
    const int w = par->wb < 0 ? qlen : par->wb + (int)(par->wf * qlen);
#define H_(r)  (DP + (size_t)(r) * 5 * row)
#define E1_(r) (H_(r) + row)
#define E2_(r) (H_(r) + 2 * row)
#define F1_(r) (H_(r) + 3 * row)
#define F2_(r) (H_(r) + 4 * row)
#define AD_BEG(id) MAX2(0, MIN2(g->maxl[id], qlen - (g->remain[id] - g->remain[end_id] - 1)) - w)
#define AD_END(id) MIN2(qlen, MAX2(g->maxr[id], qlen - (g->remain[id] - g->remain[end_id] - 1)) + w)
    /* first row :627-688 */
    if (par->wb >= 0) {
        g->maxl[beg_id] = g->maxr[beg_id] = 0;
        for (int i = 0; i < g->node[beg_id].out_n; ++i) { const int o = g->node[beg_id].out_id[i]; if (index_map[g->id2idx[o]]) g->maxl[o] = g->maxr[o] = 1; }
        dp_beg[0] = 0; dp_end[0] = AD_END(beg_id);
    } else { dp_beg[0] = 0; dp_end[0] = qlen; }
    dp_beg_sn[0] = dp_beg[0] / PN; dp_end_sn[0] = dp_end[0] / PN;
    {
        sc_t *h = H_(0), *e1 = E1_(0), *e2 = E2_(0), *f1 = F1_(0), *f2 = F2_(0);
        const int esn = MIN2(dp_end_sn[0] + 1, dp_sn - 1);
        for (int i = 0; i < (esn + 1) * PN; ++i) h[i] = e1[i] = e2[i] = inf_min;
        h[0] = 0; e1[0] = (sc_t)-oe1; e2[0] = (sc_t)-oe2; f1[0] = f2[0] = inf_min;
        for (int i = 1; i <= dp_end[0]; ++i) { f1[i] = (sc_t)(-par->o1 - par->e1 * i); f2[i] = (sc_t)(-par->o2 - par->e2 * i); h[i] = smax(f1[i], f2[i]); }
    }
    /* rows :1202-1216 */
    for (int index_i = beg_index + 1, dp_i = 1; index_i < end_index; ++index_i, ++dp_i) {
        if (!index_map[index_i]) continue;
        const int node_id = g->idx2id[index_i];
        const sc_t *q = qp + (size_t)g->node[node_id].base * row;
        sc_t *h = H_(dp_i), *e1 = E1_(dp_i), *e2 = E2_(dp_i), *f1 = F1_(dp_i), *f2 = F2_(dp_i);
        int beg, end, beg_sn, end_sn, min_pre_beg_sn, max_pre_end_sn;
        if (par->wb < 0) {
            beg = 0; end = qlen; beg_sn = 0; end_sn = end / PN; min_pre_beg_sn = 0; max_pre_end_sn = end_sn;
        } else {
            beg = AD_BEG(node_id); end = AD_END(node_id);
            beg_sn = beg / PN;
            int min_pre_beg = INT32_MAX; min_pre_beg_sn = INT32_MAX; max_pre_end_sn = -1;
            for (int i = 0; i < pre_n[dp_i]; ++i) {
                const int p = pre_index[dp_i][i];
                if (min_pre_beg > dp_beg[p]) { min_pre_beg = dp_beg[p]; min_pre_beg_sn = dp_beg_sn[p]; }
                if (max_pre_end_sn < dp_end_sn[p]) max_pre_end_sn = dp_end_sn[p];
            }
            if (beg_sn < min_pre_beg_sn) { beg = min_pre_beg; beg_sn = min_pre_beg_sn; }
            end_sn = end / PN;
        }
        dp_beg[dp_i] = beg; dp_end[dp_i] = end; dp_beg_sn[dp_i] = beg_sn; dp_end_sn[dp_i] = end_sn;
        for (int k = 0; k < pre_n[dp_i]; ++k) {
            const int p = pre_index[dp_i][k];
            const sc_t ps = (sc_t)path_score(g, node_id, k);
            const sc_t *ph = H_(p), *pe1 = E1_(p), *pe2 = E2_(p);
            const int pre_end = dp_end[p], pre_beg_sn = dp_beg_sn[p], pre_end_sn = dp_end_sn[p];
            int bsn; sc_t first;
            if (pre_beg_sn < beg_sn) { bsn = beg_sn; first = ph[beg_sn * PN - 1]; } else { bsn = pre_beg_sn; first = inf_min; }
            int esn = MIN2(MIN2((pre_end + 1) / PN, end_sn), dp_sn - 1);
            if (k == 0) {
                for (int i = beg_sn * PN; i < bsn * PN; ++i) h[i] = inf_min;
                for (int i = (esn + 1) * PN; i < (MIN2(end_sn + 1, dp_sn - 1) + 1) * PN; ++i) h[i] = inf_min;
            }
            for (int j = bsn * PN; j < (esn + 1) * PN; ++j) {
                const sc_t v = wadd(j == bsn * PN ? first : ph[j - 1], ps);
                h[j] = k == 0 ? v : smax(v, h[j]);
            }
            esn = MIN2(pre_end_sn, end_sn);
            if (k == 0) {
                for (int i = beg_sn * PN; i < bsn * PN; ++i) e1[i] = e2[i] = inf_min;
                for (int i = (esn + 1) * PN; i < (end_sn + 1) * PN; ++i) e1[i] = e2[i] = inf_min;
            }
            for (int j = bsn * PN; j < (esn + 1) * PN; ++j) {
                const sc_t v1 = wadd(pe1[j], ps), v2 = wadd(pe2[j], ps);
                e1[j] = k == 0 ? v1 : smax(v1, e1[j]);
                e2[j] = k == 0 ? v2 : smax(v2, e2[j]);
            }
        }
        for (int j = beg_sn * PN; j < (end_sn + 1) * PN; ++j) h[j] = wadd(h[j], q[j]);
        for (int i = beg_sn * PN; i < beg; ++i) h[i] = e1[i] = e2[i] = inf_min;
        for (int i = end + 1; i < (end_sn + 1) * PN; ++i) h[i] = e1[i] = e2[i] = inf_min;
        sc_t first1 = h[beg_sn * PN], first2 = first1;
        for (int sn = beg_sn; sn <= end_sn; ++sn) {
            int set_num;
            if (sn < min_pre_beg_sn) { free(index_map); return -2; }
            else if (sn > max_pre_end_sn) set_num = sn == max_pre_end_sn + 1 ? 2 : 1;
            else set_num = PN;
            sc_t *hv = h + sn * PN, *e1v = e1 + sn * PN, *e2v = e2 + sn * PN, *f1v = f1 + sn * PN, *f2v = f2 + sn * PN;
            for (int l = 0; l < PN; ++l) hv[l] = smax(smax(hv[l], e1v[l]), e2v[l]);
            for (int l = 0; l < PN; ++l) {
                f1v[l] = wsub(l == 0 ? first1 : hv[l - 1], (sc_t)oe1);
                f2v[l] = wsub(l == 0 ? first2 : hv[l - 1], (sc_t)oe2);
            }
            set_f(f1v, set_num, inf_min, par->e1);
            set_f(f2v, set_num, inf_min, par->e2);
            first1 = smax(hv[PN - 1], wadd(f1v[PN - 1], (sc_t)par->o1));
            first2 = smax(hv[PN - 1], wadd(f2v[PN - 1], (sc_t)par->o2));
            for (int l = 0; l < PN; ++l) hv[l] = smax(hv[l], smax(f1v[l], f2v[l]));
            if (sn == end_sn) for (int i = end + 1; i < (end_sn + 1) * PN; ++i) h[i] = e1[i] = e2[i] = inf_min;
            for (int l = 0; l < PN; ++l) {
                e1v[l] = smax(wsub(e1v[l], (sc_t)par->e1), wsub(hv[l], (sc_t)oe1));
                e2v[l] = smax(wsub(e2v[l], (sc_t)par->e2), wsub(hv[l], (sc_t)oe2));
            }
        }
        if (par->wb >= 0) {   /* max_in_row + ada_max_i :1107-1130 */
            int mx = inf_min, left = -1, right = -1;
            for (int i = beg; i <= end; ++i) { if (h[i] > mx) { mx = h[i]; left = right = i; } else if (h[i] == mx) right = i; }
            const pnode_t *nd = &g->node[node_id];
            for (int i = 0; i < nd->out_n; ++i) {
                const int o = nd->out_id[i];
                if (right + 1 > g->maxr[o]) g->maxr[o] = right + 1;
                if (left + 1 < g->maxl[o]) g->maxl[o] = left + 1;
            }
        }
    }
    /* global_get_max :1092-1105 */
    int best_score = inf_min, best_i = 0, best_j = 0;
    for (int i = 0; i < g->node[end_id].in_n; ++i) {
        const int in_index = g->id2idx[g->node[end_id].in_id[i]];
        if (!index_map[in_index]) continue;
        const int r = in_index - beg_index;
        const int e = qlen > dp_end[r] ? dp_end[r] : qlen;
        const int s = H_(r)[e];
        if (s > best_score) { best_score = s; best_i = r; best_j = e; }
    }
    /* backtrack :309-458 (put_gap_on_right == 0, put_gap_at_end == 0) */
    enum { M_OP = 1, E1_OP = 2, E2_OP = 4, E_OP = 6, F1_OP = 8, F2_OP = 16, F_OP = 24, ALL_OP = 31 };
    int i = best_i, j = best_j, cur_op = ALL_OP, id = g->idx2id[i + beg_index], rc = 0;
    if (best_j < qlen) cig_push(cig, 1, qlen - best_j, -1);
    while (i > 0 && j > 0) {
        const int s = mat[5 * g->node[id].base + query[j - 1]];
        int hit = 0;
        const sc_t *h = H_(i);
        for (int pass = 0; pass < 2 && !hit; ++pass) {
            if (pass == 1) {
                if (cur_op & E_OP) {          /* deletion */
                    const sc_t *e1 = E1_(i), *e2 = E2_(i);
                    for (int k = 0; k < pre_n[i] && !hit; ++k) {
                        const int p = pre_index[i][k];
                        const int ps = path_score(g, id, k);
                        if (j < dp_beg[p] || j > dp_end[p]) continue;
                        const sc_t *ph = H_(p), *pe1 = E1_(p), *pe2 = E2_(p);
                        if (cur_op & E1_OP) {
                            const int ok = (cur_op & M_OP) ? (h[j] == pe1[j] + ps) : (e1[j] == pe1[j] - par->e1 + ps);
                            if (ok) { cur_op = (ph[j] - oe1 == pe1[j]) ? (M_OP | F_OP) : E1_OP; hit = 1; }
                        }
                        if (!hit && (cur_op & E2_OP)) {
                            const int ok = (cur_op & M_OP) ? (h[j] == pe2[j] + ps) : (e2[j] == pe2[j] - par->e2 + ps);
                            if (ok) { cur_op = (ph[j] - oe2 == pe2[j]) ? (M_OP | F_OP) : E2_OP; hit = 1; }
                        }
                        if (hit) { cig_push(cig, 2, 1, id); i = p; id = g->idx2id[i + beg_index]; }
                    }
                }
                if (!hit && (cur_op & F_OP)) { /* insertion */
                    const sc_t *f1 = F1_(i), *f2 = F2_(i);
                    if (cur_op & F1_OP) {
                        if (!(cur_op & M_OP) || h[j] == f1[j]) {
                            if (h[j - 1] - oe1 == f1[j]) { cur_op = M_OP | E_OP; hit = 1; }
                            else if (f1[j - 1] - par->e1 == f1[j]) { cur_op = F1_OP; hit = 1; }
                        }
                    }
                    if (!hit && (cur_op & F2_OP)) {
                        if (!(cur_op & M_OP) || h[j] == f2[j]) {
                            if (h[j - 1] - oe2 == f2[j]) { cur_op = M_OP | E_OP; hit = 1; }
                            else if (f2[j - 1] - par->e2 == f2[j]) { cur_op = F2_OP; hit = 1; }
                        }
                    }
                    if (hit) { cig_push(cig, 1, 1, id); --j; }
                }
                if (hit) break;
            }
            /* match / mismatch: tried first, and again after E and F failed */
            if (cur_op & M_OP) {
                for (int k = 0; k < pre_n[i]; ++k) {
                    const int p = pre_index[i][k];
                    const int ps = path_score(g, id, k);
                    if (j - 1 < dp_beg[p] || j - 1 > dp_end[p]) continue;
                    if (H_(p)[j - 1] + s + ps == h[j]) {
                        cig_push(cig, 0, 1, id);
                        i = p; --j; id = g->idx2id[i + beg_index]; hit = 1; cur_op = ALL_OP;
                        break;
                    }
                }
            }
        }
        if (!hit) { rc = -3; break; }
    }
    if (rc == 0 && j > 0) cig_push(cig, 1, j, -1);
    for (int a = 0, b = cig->n - 1; a < b; ++a, --b) { gcig_t t = cig->a[a]; cig->a[a] = cig->a[b]; cig->a[b] = t; }
    for (int r = 0; r < gn; ++r) free(pre_index[r]);
    free(pre_index); free(pre_n); free(dp_beg); free(qp); free(DP); free(index_map);
    return rc;
}

/* abpoa_add_subgraph_alignment, abpoa_graph.c:689-774 */
static void poa_add_alignment(pgraph_t *g, const ppar_t *par, int beg_id, int end_id, const uint8_t *seq, int seq_l,
                              const gcigar_t *cig, int read_id) {
    const int inc = par->inc_both_ends;
    if (g->n == 2) {       /* abpoa_add_graph_sequence :573-593 */
        if (seq_l <= 0) return;
        int last = 0;
        for (int i = 0; i < seq_l; ++i) {
            const int cur = g_add_node(g, seq[i]);
            g_add_edge(g, last, cur, 0, 1, 1, read_id);
            g->node[cur].n_span = g->node[last].n_span;
            last = cur;
        }
        g_add_edge(g, last, 1, 0, 1, 1, read_id);
        g_topo_sort(g, par->wb);
        for (int i = g->id2idx[0] + 1; i < g->id2idx[1]; ++i) g->node[g->idx2id[i]].n_span += 1;
        g->node[0].n_span += 1; g->node[1].n_span += 1;
        return;
    }
    if (cig->n == 0) return;
    int query_id = -1, last_new = 0, last_id = beg_id;
    for (int c = 0; c < cig->n; ++c) {
        const int op = cig->a[c].op;
        if (op == 0) {
            const int node_id = cig->a[c].node_id;
            query_id++;
            const int add = (last_id != beg_id || inc) ? 1 : 0;
            if (g->node[node_id].base != seq[query_id]) {
                int aligned = -1;
                for (int i = 0; i < g->node[node_id].aln_n; ++i) if (g->node[g->node[node_id].aln[i]].base == seq[query_id]) { aligned = g->node[node_id].aln[i]; break; }
                if (aligned != -1) {
                    g_add_edge(g, last_id, aligned, 1 - last_new, 1, add, read_id);
                    if (!add) g->node[last_id].n_read--;
                    last_id = aligned; last_new = 0;
                } else {
                    const int nw = g_add_node(g, seq[query_id]);
                    g_add_edge(g, last_id, nw, 0, 1, add, read_id);
                    g->node[nw].n_span = g->node[last_id].n_span;
                    if (!add) g->node[last_id].n_read--;
                    last_id = nw; last_new = 1;
                    g_add_aligned(g, node_id, nw);
                }
            } else {
                g_add_edge(g, last_id, node_id, 1 - last_new, 1, add, read_id);
                if (!add) g->node[last_id].n_read--;
                last_id = node_id; last_new = 0;
            }
        } else if (op == 1) {
            const int len = cig->a[c].len;
            query_id += len;
            for (int j = len - 1; j >= 0; --j) {
                const int nw = g_add_node(g, seq[query_id - j]);
                const int add = (last_id != beg_id || inc) ? 1 : 0;
                g_add_edge(g, last_id, nw, 0, 1, add, read_id);
                g->node[nw].n_span = g->node[last_id].n_span;
                if (!add) g->node[last_id].n_read--;
                last_id = nw; last_new = 1;
            }
        }
    }
    g_add_edge(g, last_id, end_id, 1 - last_new, 1, 1, read_id);
    g_topo_sort(g, par->wb);
    for (int i = g->id2idx[beg_id] + 1; i < g->id2idx[end_id]; ++i) g->node[g->idx2id[i]].n_span += 1;
    if (inc) { g->node[beg_id].n_span += 1; g->node[end_id].n_span += 1; }
}

/* abpoa_DFS_set_msa_rank, abpoa_graph.c:359-410 */
static void poa_msa_rank(pgraph_t *g) {
    if (g->msa_ranked) return;
    const int n = g->n;
    int *deg = (int*)malloc((size_t)n * sizeof(int)), *st = (int*)malloc((size_t)(n + 1) * 2 * sizeof(int));
    for (int i = 0; i < n; ++i) deg[i] = g->node[i].in_n;
    int sp = 0, rank = 0;
    st[sp++] = 0; g->msa_rank[0] = -1;
    while (sp > 0) {
        const int cur = st[--sp];
        if (g->msa_rank[cur] < 0) {
            g->msa_rank[cur] = rank;
            for (int i = 0; i < g->node[cur].aln_n; ++i) g->msa_rank[g->node[cur].aln[i]] = rank;
            rank++;
        }
        if (cur == 1) break;
        for (int i = 0; i < g->node[cur].out_n; ++i) {
            const int out = g->node[cur].out_id[i];
            if (--deg[out] == 0) {
                int ok = 1;
                for (int j = 0; j < g->node[out].aln_n; ++j) if (deg[g->node[out].aln[j]] != 0) { ok = 0; break; }
                if (!ok) continue;
                st[sp++] = out; g->msa_rank[out] = -1;
                for (int j = 0; j < g->node[out].aln_n; ++j) { st[sp++] = g->node[out].aln[j]; g->msa_rank[g->node[out].aln[j]] = -1; }
            }
        }
    }
    free(deg); free(st);
    g->msa_ranked = 1;
}
static int node_col(const pgraph_t *g, int id) {     /* 1-based MSA column of a node (max over its aligned group) */
    int rank = g->msa_rank[id];
    for (int j = 0; j < g->node[id].aln_n; ++j) rank = MAX2(rank, g->msa_rank[g->node[id].aln[j]]);
    return rank;
}

/* is_full_upstream_subgraph / abpoa_upstream_index / abpoa_downstream_index / abpoa_subgraph_nodes, abpoa_graph.c:595-680: the sub-graph a
 * partially covering read is aligned against -- the nodes whose BFS index lies between the outermost predecessors / successors of the
 * index range of the two anchor nodes (bases of the first read), closed under in-edges on the left and out-edges on the right */
static int full_upstream(const pgraph_t *g, int up_index, int down_index, int beg_index, int end_index) {
    const int min_index = MIN2(up_index, beg_index), max_index = MAX2(down_index, end_index);
    for (int i = up_index + 1; i <= down_index; ++i) {
        const pnode_t *nd = &g->node[g->idx2id[i]];
        for (int j = 0; j < nd->in_n; ++j) { const int x = g->id2idx[nd->in_id[j]]; if (x < min_index || x > max_index) return 0; }
    }
    return 1;
}
static int upstream_index(const pgraph_t *g, int beg_index, int end_index) {
    for (;;) {
        int min_index = beg_index;
        for (int i = beg_index; i <= end_index; ++i) {
            const pnode_t *nd = &g->node[g->idx2id[i]];
            for (int j = 0; j < nd->in_n; ++j) min_index = MIN2(min_index, g->id2idx[nd->in_id[j]]);
        }
        if (full_upstream(g, min_index, beg_index, beg_index, end_index)) return min_index;
        end_index = beg_index; beg_index = min_index;
    }
}
static int downstream_index(const pgraph_t *g, int beg_index, int end_index) {
    for (;;) {
        int max_index = end_index;
        for (int i = beg_index; i <= end_index; ++i) {
            const pnode_t *nd = &g->node[g->idx2id[i]];
            for (int j = 0; j < nd->out_n; ++j) max_index = MAX2(max_index, g->id2idx[nd->out_id[j]]);
        }
        if (full_upstream(g, end_index, max_index, beg_index, end_index)) return max_index;       /* (the reference tests the upstream property here too: :653) */
        beg_index = end_index; end_index = max_index;
    }
}
static void subgraph_nodes(const pgraph_t *g, int inc_beg_id, int inc_end_id, int *exc_beg, int *exc_end) {
    const int b = g->id2idx[inc_beg_id], e = g->id2idx[inc_end_id];
    *exc_beg = g->idx2id[upstream_index(g, b, e)]; *exc_end = g->idx2id[downstream_index(g, b, e)];
}

static int poa_run(int n_seq, const uint8_t *seqs, const int64_t *seq_off, const int32_t *seq_len, const int32_t *sub_beg, const int32_t *sub_end,
                   const lcd_poa_params_t *p, uint8_t *cons, int32_t *cons_len, uint8_t *msa, int32_t *msa_len, int32_t msa_cap);
int lcd_oracle_poa(int n_seq, const uint8_t *seqs, const int64_t *seq_off, const int32_t *seq_len,
                   const lcd_poa_params_t *p, uint8_t *cons, int32_t *cons_len,
                   uint8_t *msa, int32_t *msa_len, int32_t msa_cap) {
    return poa_run(n_seq, seqs, seq_off, seq_len, NULL, NULL, p, cons, cons_len, msa, msa_len, msa_cap);
}
/* abpoa_partial_aln_msa_cons with partially covering reads (src/align.c:790-812): read r > 0 is aligned against the sub-graph between the nodes
 * sub_beg[r] and sub_end[r] (abpoa_subgraph_nodes; both 0: the whole graph) or left out altogether (sub_beg[r] < 0: :800 `continue`) */
int lcd_oracle_poa_sub(int n_seq, const uint8_t *seqs, const int64_t *seq_off, const int32_t *seq_len, const int32_t *sub_beg, const int32_t *sub_end,
                       const lcd_poa_params_t *p, uint8_t *cons, int32_t *cons_len, uint8_t *msa, int32_t *msa_len, int32_t msa_cap) {
    return poa_run(n_seq, seqs, seq_off, seq_len, sub_beg, sub_end, p, cons, cons_len, msa, msa_len, msa_cap);
}
static int poa_run(int n_seq, const uint8_t *seqs, const int64_t *seq_off, const int32_t *seq_len, const int32_t *sub_beg, const int32_t *sub_end,
                   const lcd_poa_params_t *p, uint8_t *cons, int32_t *cons_len, uint8_t *msa, int32_t *msa_len, int32_t msa_cap) {
    ppar_t par; par.wb = p->wb; par.wf = p->wf; par.match = p->match; par.mismatch = p->mismatch;
    par.o1 = p->gap_open1; par.e1 = p->gap_ext1; par.o2 = p->gap_open2; par.e2 = p->gap_ext2;
    par.inc_both_ends = p->sub_aln ? 0 : 1; par.sub_aln = p->sub_aln;
    pgraph_t g; memset(&g, 0, sizeof(g));
    g.rid_n = 1 + ((n_seq - 1) >> 6);
    g_add_node(&g, 0); g_add_node(&g, 0);
    int rc = 0;
    *cons_len = 0; *msa_len = 0;
    for (int r = 0; r < n_seq && rc == 0; ++r) {
        const uint8_t *q = seqs + seq_off[r]; const int ql = seq_len[r];
        gcigar_t cig; memset(&cig, 0, sizeof(cig));
        int exc_beg = 0, exc_end = 1;
        if (r > 0 && sub_beg && sub_beg[r] < 0) continue;
        if (r > 0 && sub_beg && sub_beg[r] > 0 && g.n > 2) {
            if (sub_beg[r] >= g.n || sub_end[r] >= g.n || sub_end[r] < 2 || sub_beg[r] < 2) { rc = -6; break; }
            subgraph_nodes(&g, sub_beg[r], sub_end[r], &exc_beg, &exc_end);
        }
        if (g.n > 2) rc = poa_align(&g, &par, exc_beg, exc_end, q, ql, &cig);      /* abpoa_align_sequence_to_subgraph: -1 when node_n <= 2 */
        if (rc == 0) poa_add_alignment(&g, &par, exc_beg, exc_end, q, ql, &cig, r);
        free(cig.a);
    }
    if (rc == 0 && g.n > 2) {
        /* most frequent consensus with one cluster (abpoa_output.c:393-475,549-586) */
        poa_msa_rank(&g);
        const int ml = g.msa_rank[1] - 1;
        int *cnt = (int*)calloc((size_t)ml * 5, sizeof(int)), *nid = (int*)calloc((size_t)ml * 5, sizeof(int));
        for (int i = 2; i < g.n; ++i) { const int col = node_col(&g, i) - 1; nid[col * 5 + g.node[i].base] = i; cnt[col * 5 + g.node[i].base] = g.node[i].n_read; }
        int *cons_ids = (int*)malloc((size_t)MAX2(ml, 1) * sizeof(int)); int cl = 0;
        for (int i = 0; i < ml; ++i) {
            int max_c = 0, total = 0, max_base = 5;
            for (int j = 0; j < 4; ++j) { const int c = cnt[i * 5 + j]; if (c > max_c) { max_c = c; max_base = j; } total += c; }
            if (max_base == 5) { if (par.sub_aln) { rc = -4; break; } continue; }       /* sub_aln: the reference would read out of bounds; else 0 >= n_seq - 0 fails: no base */
            const int gap_c = (par.sub_aln ? g.node[nid[i * 5 + max_base]].n_span : n_seq) - total;
            if (max_c >= gap_c) { cons_ids[cl] = nid[i * 5 + max_base]; cons[cl] = (uint8_t)max_base; cl++; }
        }
        *cons_len = cl;
        /* row-column MSA: n_seq rows + 1 consensus row (abpoa_generate_rc_msa) */
        if (rc == 0 && msa) {
            if ((int64_t)(n_seq + 1) * ml > msa_cap) rc = -5;
            else {
                memset(msa, 5, (size_t)(n_seq + 1) * ml);
                for (int i = 2; i < g.n; ++i) {
                    const int col = node_col(&g, i) - 1; const pnode_t *nd = &g.node[i];
                    for (int e = 0; e < nd->out_n; ++e) for (int r = 0; r < n_seq; ++r)
                        if (nd->rid[(size_t)e * g.rid_n + r / 64] >> (r & 63) & 1) msa[(size_t)r * ml + col] = (uint8_t)nd->base;
                }
                for (int i = 0; i < cl; ++i) msa[(size_t)n_seq * ml + node_col(&g, cons_ids[i]) - 1] = cons[i];
                *msa_len = ml;
            }
        }
        free(cnt); free(nid); free(cons_ids);
    }
    g_free(&g);
    return rc;
}

/* abpoa_aln_msa_cons with max_n_cons = 2 (src/align.c:872-953; wb = -1, sub_aln = 0): the progressive POA of all reads, abPOA's read
 * clustering on the row-column MSA (poa_cluster.c) and one most-frequent consensus per cluster.  cons: the consensus sequences back to
 * back; msa: (n_seq + n_cons) rows.  min_freq = opt->min_af (abpt->min_freq, a double: min_w = MAX(2, ceil(n_seq * min_freq))). */
int lcd_oracle_poa_cluster(const uint8_t *msa, int n_seq, int ml, int min_w, uint8_t *read_clu);
void lcd_oracle_poa_cluster_cons(uint8_t *msa, int n_seq, int ml, int n_clu, const uint8_t *read_clu, uint8_t *cons, int32_t *cons_len);
int lcd_oracle_poa_ncons(int n_seq, const uint8_t *seqs, const int64_t *seq_off, const int32_t *seq_len, const lcd_poa_params_t *p, double min_freq,
                         uint8_t *cons, int32_t *cons_len, int32_t *n_cons, uint8_t *read_clu, uint8_t *msa, int32_t *msa_len, int32_t msa_cap) {
    if (p->sub_aln) return -7;
    int64_t tot = 0; for (int r = 0; r < n_seq; ++r) tot += seq_len[r];
    const int64_t cap1 = (int64_t)(n_seq + 1) * (tot + 8);
    uint8_t *m1 = (uint8_t*)malloc((size_t)cap1), *c1 = (uint8_t*)malloc((size_t)tot + 8);
    int32_t cl1 = 0, ml = 0;
    cons_len[0] = cons_len[1] = 0; *n_cons = 0; *msa_len = 0;
    int rc = poa_run(n_seq, seqs, seq_off, seq_len, NULL, NULL, p, c1, &cl1, m1, &ml, (int32_t)(cap1 > INT32_MAX ? INT32_MAX : cap1));
    if (rc == 0 && ml > 0) {
        int n_clu = 1;
        memset(read_clu, 0, (size_t)n_seq);
        if (p->max_n_cons > 1) { const int cw = (int)ceil(n_seq * min_freq); n_clu = lcd_oracle_poa_cluster(m1, n_seq, ml, cw > 2 ? cw : 2, read_clu); }
        if ((int64_t)(n_seq + n_clu) * ml > msa_cap) rc = -5;
        else {
            memcpy(msa, m1, (size_t)n_seq * ml);
            lcd_oracle_poa_cluster_cons(msa, n_seq, ml, n_clu, read_clu, cons, cons_len);
            *n_cons = n_clu; *msa_len = ml;
        }
    }
    free(m1); free(c1);
    return rc;
}
