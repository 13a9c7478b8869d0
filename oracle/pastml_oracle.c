/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the pastml ACR hot path.
 *
 * A plain-C restatement, on flat level-ordered arrays, of the reference's per-node numpy/Python code
 * (/root/reference/pastml/ml.py, parsimony.py, models/*).  Every function cites the reference lines it follows
 * and keeps the reference's formulation (log10-domain accumulation, per-child rescale_log, explicit P(t)
 * matrices, numpy's summation order for ndarray.sum) rather than anything clever: it is the checker the CUDA
 * path is compared with, never the product.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.
 *
 * Parity status: PINNED -- tests/test_oracle_golden.py checks every function below against outputs of the
 * unmodified reference run in the build container (fixtures under tests/golden/, made by oracle/make_golden.py),
 * including the reference's own known answers (log-lh -110.178 / sf 3.841 for Albania MPPA+F81, 32 parsimony
 * steps, ...).
 *
 * Conventions: nodes are numbered in level order over the forest (roots first); children of a node are
 * contiguous (first_child, n_children); state sets are bit masks in a uint64 (K <= 64); all reals are doubles.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_OK 0
#define ORC_ERR_ZERO_LIKELIHOOD 1   /* ml.py:139-145 */
#define ORC_ERR_EMPTY_RESCALE 2     /* np.min of an empty array inside rescale_log, ml.py:161-163 */
#define ORC_ERR_ALLOC 3

enum { ORC_F81 = 0, ORC_HKY = 1, ORC_EIG = 2 };

typedef struct {
    int32_t kind;        /* ORC_F81 (F81, JC, EFT), ORC_HKY, ORC_EIG (CUSTOM_RATES, JTT) */
    int32_t K;
    double sf;           /* scaling factor                      models/__init__.py:269-270 */
    double tau;          /* smoothing factor                                               */
    double tau_factor;   /* forest_length / (forest_length + tau (N-1)), or 1   :39-42     */
    double mu;           /* F81: 1 / (1 - pi.pi), computed by the caller        F81Model.py:18-26 */
    double kappa;        /* HKY                                                             */
    const double *pi;    /* K frequencies                                                  */
    const double *d;     /* EIG: eigenvalues, K                                 generator.py:16-30 */
    const double *A;     /* EIG: K x K row-major                                           */
    const double *Ainv;  /* EIG: K x K row-major                                           */
} orc_model;

typedef struct {
    int32_t n_nodes;
    const int32_t *parent;
    const int32_t *first_child;
    const int32_t *n_children;
    const double *dist;
} orc_tree;

static const double ORC_LOG10E = 0.43429448190325182765;

static double min_value(void) { return log10(2.220446049250313e-16); }      /* ml.py:26 */
static double max_value(void) { return log10(1.7976931348623157e308); }    /* ml.py:27 */

/* numpy's float64 add-reduction over a contiguous run (what ndarray.sum / sum(axis=1) on a fresh C-contiguous
 * array does): sequential for n < 8, eight partial sums + tail for n <= 128, pairwise recursion above
 * (numpy pairwise_sum, block size 128). */
static double np_sum(const double *a, int n) {
    if (n < 8) {
        double s = 0.;
        if (n == 0) return 0.;
        s = a[0];
        for (int i = 1; i < n; ++i) s += a[i];
        return s;
    }
    if (n <= 128) {
        double r[8];
        int i;
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return np_sum(a, n2) + np_sum(a + n2, n - n2);
}

/* Model.transform_t, models/__init__.py:269-270 */
static double transform_t(const orc_model *m, double t) { return (t + m->tau) * m->tau_factor * m->sf; }

/* get_Pij_t for the three model families; P is K x K row-major, P[i][j] = P(i -> j). */
void orc_pij(const orc_model *m, double t, double *P) {
    const int K = m->K;
    t = transform_t(m, t);
    if (m->kind == ORC_F81) {
        /* F81Model.py:41-46: (1 - e) * pi_j + eye * e, e = 0 when mu == inf */
        double e = isinf(m->mu) ? 0. : exp(-m->mu * t);
        for (int i = 0; i < K; ++i)
            for (int j = 0; j < K; ++j)
                P[i * K + j] = (1 - e) * m->pi[j] + (i == j ? 1. : 0.) * e;
    } else if (m->kind == ORC_HKY) {
        /* HKYModel.py:53-82, states A, C, G, T = 0..3 */
        const double pa = m->pi[0], pc = m->pi[1], pg = m->pi[2], pt = m->pi[3], k = m->kappa;
        const double pag = pa + pg, pct = pc + pt;
        const double beta = .5 / (pag * pct + k * (pa * pg + pc * pt));
        const double e1 = exp(-beta * t);
        const double ect = exp(-beta * t * (1. + pct * (k - 1.))) / pct;
        const double eag = exp(-beta * t * (1. + pag * (k - 1.))) / pag;
        const double ct_sum = (pct + pag * e1) / pct;
        const double ag_sum = (pag + pct * e1) / pag;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j)
                P[i * 4 + j] = (1. * (1 - e1)) * m->pi[j];
        P[3 * 4 + 3] = pt * ct_sum + pc * ect;
        P[3 * 4 + 1] = pc * ct_sum - pc * ect;
        P[1 * 4 + 3] = pt * ct_sum - pt * ect;
        P[1 * 4 + 1] = pc * ct_sum + pt * ect;
        P[0 * 4 + 0] = pa * ag_sum + pg * eag;
        P[0 * 4 + 2] = pg * ag_sum - pg * eag;
        P[2 * 4 + 0] = pa * ag_sum - pa * eag;
        P[2 * 4 + 2] = pg * ag_sum + pa * eag;
    } else {
        /* generator.py:54-65: A . diag(exp(d t)) . A^-1 */
        double *tmp = (double *) malloc(sizeof(double) * K * K);
        for (int i = 0; i < K; ++i)
            for (int mm = 0; mm < K; ++mm)
                tmp[i * K + mm] = m->A[i * K + mm] * exp(m->d[mm] * t);
        for (int i = 0; i < K; ++i)
            for (int j = 0; j < K; ++j) {
                double s = 0.;
                for (int mm = 0; mm < K; ++mm) s += tmp[i * K + mm] * m->Ainv[mm * K + j];
                P[i * K + j] = s;
            }
        free(tmp);
    }
}

/* rescale_log, ml.py:151-171.  Returns the shift through *factor; ORC_ERR_EMPTY_RESCALE if no finite entry. */
static int rescale_log(double *a, int K, double *factor) {
    double mn = INFINITY, mx = -INFINITY;
    int any = 0;
    for (int i = 0; i < K; ++i)
        if (a[i] > -INFINITY) {
            any = 1;
            if (a[i] < mn) mn = a[i];
            if (a[i] > mx) mx = a[i];
        }
    if (!any) return ORC_ERR_EMPTY_RESCALE;
    double f = 0.;
    if (mx > max_value()) f = max_value() - mx - 1;
    else if (mn < min_value()) f = fmin(min_value() - mn + 1, max_value() - mx - 1);
    for (int i = 0; i < K; ++i) a[i] += f;
    *factor = f;
    return ORC_OK;
}

/* get_zero_clusters_with_states + alter_zero_node_allowed_states, ml.py:321-387.
 * has_state[n] != 0 iff the node carries the annotation feature (even an empty set).  For every cluster of >= 2
 * annotated nodes joined by zero-length branches that share no common allowed state, every member receives the
 * union of all members' masks (the reference assigns one array object to all of them, ml.py:383-384) and the old
 * mask is saved in initial[] / has_initial[] (the '.initial' feature, never cleared afterwards).
 * altered[] (optional) gets 1 for the nodes touched by this call.  Returns the number of altered nodes. */
int orc_alter_zero_clusters(const orc_tree *tr, const uint8_t *has_state, int K, uint64_t *allowed,
                            uint64_t *initial, uint8_t *has_initial, uint8_t *altered) {
    const int n = tr->n_nodes;
    int32_t *cluster_of = (int32_t *) malloc(sizeof(int32_t) * n);
    int32_t *members = (int32_t *) malloc(sizeof(int32_t) * n);
    int n_altered = 0;
    (void) K;
    if (altered) memset(altered, 0, n);
    /* level order: the top of each zero-connected component comes first */
    for (int v = 0; v < n; ++v) {
        int p = tr->parent[v];
        cluster_of[v] = (p >= 0 && tr->dist[v] == 0) ? cluster_of[p] : v;
    }
    /* bucket annotated nodes by component top; components are small, a simple two-pass scan per top is enough */
    int32_t *count = (int32_t *) calloc(n, sizeof(int32_t));
    int32_t *start = (int32_t *) malloc(sizeof(int32_t) * (n + 1));
    for (int v = 0; v < n; ++v) if (has_state[v]) count[cluster_of[v]]++;
    start[0] = 0;
    for (int v = 0; v < n; ++v) start[v + 1] = start[v] + count[v];
    memset(count, 0, sizeof(int32_t) * n);
    for (int v = 0; v < n; ++v) if (has_state[v]) members[start[cluster_of[v]] + count[cluster_of[v]]++] = v;
    for (int top = 0; top < n; ++top) {
        int sz = start[top + 1] - start[top];
        if (sz < 2) continue;
        uint64_t common = ~0ULL, uni = 0ULL;
        for (int q = start[top]; q < start[top + 1]; ++q) {
            common &= allowed[members[q]];
            uni |= allowed[members[q]];
        }
        if (common) continue;  /* counts.max() == len(cluster), ml.py:374 */
        for (int q = start[top]; q < start[top + 1]; ++q) {
            int v = members[q];
            initial[v] = allowed[v];
            has_initial[v] = 1;
            allowed[v] = uni;
            if (altered) altered[v] = 1;
            ++n_altered;
        }
    }
    free(cluster_of); free(members); free(count); free(start);
    return n_altered;
}

/* unalter_zero_node_allowed_states, ml.py:390-405 (for the nodes flagged in altered[]) */
void orc_unalter_allowed(int n, const uint8_t *altered, uint64_t *allowed, const uint64_t *initial) {
    for (int v = 0; v < n; ++v)
        if (altered[v]) {
            uint64_t both = allowed[v] & initial[v];
            allowed[v] = both ? both : initial[v];
        }
}

/* get_bottom_up_loglikelihood + calc_node_bu_likelihood, ml.py:82-148, without the alter/unalter bracket
 * (the caller applies orc_alter_zero_clusters first when the reference would).
 * bu[n*K], bu_sf[n]: BOTTOM_UP_LIKELIHOOD(_SF) of every node; joint_ptr[n*K] (may be NULL when marginal):
 * BOTTOM_UP_LIKELIHOOD_JOINT_STATES of every non-root node.  *loglh = sum over the forest's trees.
 * On ORC_ERR_ZERO_LIKELIHOOD err_nodes = {parent, child}. */
int orc_bottom_up(const orc_tree *tr, const orc_model *m, const uint64_t *allowed, int is_marginal,
                  double *bu, double *bu_sf, int32_t *joint_ptr, double *loglh, int32_t *err_nodes) {
    const int n = tr->n_nodes, K = m->K;
    double *P = (double *) malloc(sizeof(double) * K * K);
    double *row = (double *) malloc(sizeof(double) * K);
    double *a = (double *) malloc(sizeof(double) * K);
    if (!P || !row || !a) return ORC_ERR_ALLOC;
    int rc = ORC_OK;
    for (int v = n - 1; v >= 0 && rc == ORC_OK; --v) {
        /* log10(ones * allowed): 0 or -inf, ml.py:127 */
        for (int i = 0; i < K; ++i) a[i] = ((allowed[v] >> i) & 1ULL) ? 0. : -INFINITY;
        double factors = 0.;
        for (int c = tr->first_child[v], ce = c + tr->n_children[v]; c < ce; ++c) {
            orc_pij(m, tr->dist[c], P);
            const double *Lc = bu + (size_t) c * K;
            int all_ninf = 1;
            for (int i = 0; i < K; ++i) {
                double val;
                if (is_marginal) {
                    for (int j = 0; j < K; ++j) row[j] = P[i * K + j] * Lc[j];
                    val = np_sum(row, K);                     /* ml.py:132 */
                } else {
                    int best = 0;
                    double bestv = P[i * K] * Lc[0];
                    for (int j = 1; j < K; ++j) {
                        double x = P[i * K + j] * Lc[j];
                        if (x > bestv) { bestv = x; best = j; }   /* first maximum, ml.py:134 */
                    }
                    joint_ptr[(size_t) c * K + i] = best;
                    val = bestv;                              /* ml.py:136 */
                }
                val = val > 0. ? val : 0.;                    /* np.maximum(.., 0), ml.py:137 */
                a[i] += log10(val);                           /* ml.py:138 */
                if (a[i] != -INFINITY) all_ninf = 0;
            }
            if (all_ninf) {                                   /* ml.py:139-145 */
                if (err_nodes) { err_nodes[0] = v; err_nodes[1] = c; }
                rc = ORC_ERR_ZERO_LIKELIHOOD;
                break;
            }
            double f;
            rc = rescale_log(a, K, &f);                       /* ml.py:146 */
            if (rc != ORC_OK) break;
            factors += f;
        }
        if (rc != ORC_OK) break;
        for (int i = 0; i < K; ++i) bu[(size_t) v * K + i] = pow(10., a[i]);     /* ml.py:147 */
        double sfsum = 0.;
        for (int c = tr->first_child[v], ce = c + tr->n_children[v]; c < ce; ++c) sfsum += bu_sf[c];
        bu_sf[v] = factors + sfsum;                           /* ml.py:148 */
    }
    if (rc == ORC_OK) {
        double total = 0.;
        for (int v = 0; v < n; ++v)
            if (tr->parent[v] < 0) {
                for (int i = 0; i < K; ++i) row[i] = bu[(size_t) v * K + i] * m->pi[i];      /* ml.py:112 */
                double r;
                if (is_marginal) r = np_sum(row, K);
                else { r = row[0]; for (int i = 1; i < K; ++i) if (row[i] > r) r = row[i]; }
                total += log(r) - bu_sf[v] / ORC_LOG10E;      /* ml.py:121 */
            }
        *loglh = total;
    }
    free(P); free(row); free(a);
    return rc;
}

/* calculate_top_down_likelihood + calc_node_td_likelihood, ml.py:240-290 */
int orc_top_down(const orc_tree *tr, const orc_model *m, const double *bu, const double *bu_sf,
                 double *td, double *td_sf) {
    const int n = tr->n_nodes, K = m->K;
    double *P = (double *) malloc(sizeof(double) * K * K);
    double *b = (double *) malloc(sizeof(double) * K);
    double *q = (double *) malloc(sizeof(double) * K);
    int rc = ORC_OK;
    for (int v = 0; v < n && rc == ORC_OK; ++v) {
        int p = tr->parent[v];
        if (p < 0) {
            for (int i = 0; i < K; ++i) td[(size_t) v * K + i] = 1.;
            td_sf[v] = 0.;
            continue;
        }
        orc_pij(m, tr->dist[v], P);
        for (int i = 0; i < K; ++i) {
            double c = 0.;                                    /* L_node . P^T, ml.py:280 */
            for (int j = 0; j < K; ++j) c += bu[(size_t) v * K + j] * P[i * K + j];
            if (c <= 0.) c = 1.;                              /* ml.py:281 */
            b[i] = log10(td[(size_t) p * K + i]) + log10(bu[(size_t) p * K + i]) - log10(c);   /* ml.py:282-283 */
        }
        double factors = td_sf[p] + bu_sf[p] - bu_sf[v];      /* ml.py:284-285 */
        double f;
        rc = rescale_log(b, K, &f);
        if (rc != ORC_OK) break;
        factors += f;
        for (int i = 0; i < K; ++i) q[i] = pow(10., b[i]);
        for (int i = 0; i < K; ++i) {
            double s = 0.;                                    /* parent_likelihood . P^T, ml.py:288 */
            for (int j = 0; j < K; ++j) s += q[j] * P[i * K + j];
            td[(size_t) v * K + i] = s > 0. ? s : 0.;
        }
        td_sf[v] = factors;
    }
    free(P); free(b); free(q);
    return rc;
}

/* calculate_marginal_likelihoods + convert_likelihoods_to_probabilities, ml.py:431-465, 486-502.
 * lh[n*K], lh_sf[n]: the LIKELIHOOD / LIKELIHOOD_SF features; probs[n*K] = lh / lh.sum(). */
int orc_marginal(const orc_tree *tr, const orc_model *m, const uint64_t *allowed, const double *bu,
                 const double *bu_sf, const double *td, const double *td_sf,
                 double *lh, double *lh_sf, double *probs) {
    const int n = tr->n_nodes, K = m->K;
    double *a = (double *) malloc(sizeof(double) * K);
    int rc = ORC_OK;
    for (int v = 0; v < n; ++v) {
        for (int i = 0; i < K; ++i) {
            double fa = m->pi[i] * (double) ((allowed[v] >> i) & 1ULL);
            a[i] = log10(bu[(size_t) v * K + i]) + log10(td[(size_t) v * K + i]) + log10(fa);   /* ml.py:456-457 */
        }
        double f;
        rc = rescale_log(a, K, &f);
        if (rc != ORC_OK) break;
        for (int i = 0; i < K; ++i) lh[(size_t) v * K + i] = pow(10., a[i]);
        lh_sf[v] = f + td_sf[v] + bu_sf[v];
        if (probs) {
            double s = np_sum(lh + (size_t) v * K, K);        /* ml.py:500 */
            for (int i = 0; i < K; ++i) probs[(size_t) v * K + i] = lh[(size_t) v * K + i] / s;
        }
    }
    free(a);
    return rc;
}

/* choose_ancestral_states_map, ml.py:577-595.  lh is modified in place like the reference does (lh *= initial). */
void orc_choose_map(int n, int K, double *lh, const uint64_t *initial, const uint8_t *has_initial, uint64_t *allowed) {
    for (int v = 0; v < n; ++v) {
        double *l = lh + (size_t) v * K;
        if (has_initial[v])
            for (int i = 0; i < K; ++i) l[i] *= (double) ((initial[v] >> i) & 1ULL);
        int best = 0;
        for (int i = 1; i < K; ++i) if (l[i] > l[best]) best = i;
        allowed[v] = 1ULL << best;
    }
}

static int cmp_double(const void *x, const void *y) {
    double a = *(const double *) x, b = *(const double *) y;
    return (a > b) - (a < b);
}

/* choose_ancestral_states_mppa, ml.py:505-574.  joint_state may be NULL when force_joint == 0.
 * k_hist[K+1] accumulates the histogram of best_k (num_scenarios = prod k^hist[k]). */
void orc_choose_mppa(int n, int K, double *lh, const uint64_t *initial, const uint8_t *has_initial,
                     const int32_t *joint_state, int force_joint, uint64_t *allowed, int64_t *k_hist,
                     int32_t *best_k_out) {
    double *p = (double *) malloc(sizeof(double) * K);
    double *sorted = (double *) malloc(sizeof(double) * K);
    int *order = (int *) malloc(sizeof(int) * K);
    for (int v = 0; v < n; ++v) {
        double *l = lh + (size_t) v * K;
        if (has_initial[v])
            for (int i = 0; i < K; ++i) l[i] *= (double) ((initial[v] >> i) & 1ULL);   /* ml.py:541-542 */
        double s = np_sum(l, K);
        for (int i = 0; i < K; ++i) p[i] = l[i] / s;                                 /* ml.py:543 */
        if (force_joint) {                                                           /* ml.py:544-547 */
            int ji = joint_state[v], q = 0;
            for (int i = 0; i < K; ++i) if (i != ji) sorted[q++] = p[i];
            qsort(sorted, K - 1, sizeof(double), cmp_double);
            sorted[K - 1] = p[ji];
        } else {
            memcpy(sorted, p, sizeof(double) * K);
            qsort(sorted, K, sizeof(double), cmp_double);
        }
        int best_k = K;
        double best_corr = INFINITY;
        for (int k = 1; k <= K; ++k) {                                               /* ml.py:552-557 */
            double corr = 0.;
            double target = 1. / (double) k;
            for (int i = 0; i < K; ++i) {
                double dlt = (i < K - k ? 0. : target) - sorted[i];
                corr += dlt * dlt;
            }
            if (corr < best_corr) { best_corr = corr; best_k = k; }
        }
        /* stable sort by decreasing likelihood, lower index first; the force_joint key is a no-op (ml.py:562-565) */
        for (int i = 0; i < K; ++i) order[i] = i;
        for (int i = 1; i < K; ++i) {
            int o = order[i], j = i - 1;
            while (j >= 0 && l[order[j]] < l[o]) { order[j + 1] = order[j]; --j; }
            order[j + 1] = o;
        }
        uint64_t mask = 0;
        for (int i = 0; i < best_k; ++i) mask |= 1ULL << order[i];
        allowed[v] = mask;
        if (k_hist) k_hist[best_k]++;
        if (best_k_out) best_k_out[v] = best_k;
    }
    free(p); free(sorted); free(order);
}

/* choose_ancestral_states_joint, ml.py:598-622, preceded by unalter_zero_node_joint_states, ml.py:408-428,
 * for the nodes flagged in altered[] (may be NULL).  joint_ptr is modified like the reference modifies the
 * BOTTOM_UP_LIKELIHOOD_JOINT_STATES features. */
void orc_choose_joint(const orc_tree *tr, const orc_model *m, const double *bu, int32_t *joint_ptr,
                      const uint8_t *altered, const uint64_t *initial, int32_t *joint_state, uint64_t *allowed) {
    const int n = tr->n_nodes, K = m->K;
    if (altered)
        for (int v = 0; v < n; ++v)
            if (altered[v] && tr->parent[v] >= 0) {
                int first = 0;
                while (first < K && !((initial[v] >> first) & 1ULL)) ++first;
                if (first == K) first = 0;           /* np.argmax of all zeros */
                int cnt = 0;
                for (int i = 0; i < K; ++i) cnt += (int) ((initial[v] >> i) & 1ULL);
                int32_t *ptr = joint_ptr + (size_t) v * K;
                if (cnt == 1) for (int i = 0; i < K; ++i) ptr[i] = first;
                else for (int i = 0; i < K; ++i) if (!((initial[v] >> ptr[i]) & 1ULL)) ptr[i] = first;
            }
    for (int v = 0; v < n; ++v) {
        int p = tr->parent[v];
        int s;
        if (p < 0) {
            s = 0;
            double best = bu[(size_t) v * K] * m->pi[0];
            for (int i = 1; i < K; ++i) {
                double x = bu[(size_t) v * K + i] * m->pi[i];
                if (x > best) { best = x; s = i; }
            }
        } else s = joint_ptr[(size_t) v * K + joint_state[p]];
        joint_state[v] = s;
        allowed[v] = 1ULL << s;
    }
}

/* ------------------------------------------------------------------ parsimony (parsimony.py) */

/* get_most_common_states, parsimony.py:76-89: states with the maximal count over the given sets */
static uint64_t most_common(const uint64_t *sets, int n_sets, int K) {
    int best = 0;
    uint64_t out = 0;
    for (int s = 0; s < K; ++s) {
        int cnt = 0;
        for (int i = 0; i < n_sets; ++i) cnt += (int) ((sets[i] >> s) & 1ULL);
        if (cnt > best) { best = cnt; out = 1ULL << s; }
        else if (cnt == best && cnt > 0) out |= 1ULL << s;
    }
    return out;
}

/* initialise_parsimonious_states + uppass, parsimony.py:51-73, 92-122.
 * annot[v] = annotated state set (0 = none -> all states).  down[v] = BOTTOM_UP_PARSIMONY after the uppass,
 * preset[v] = the PARSIMONY feature as initialised (the pre-uppass sets). */
void orc_pars_uppass(const orc_tree *tr, int K, const uint64_t *annot, uint64_t *down, uint64_t *preset) {
    const int n = tr->n_nodes;
    const uint64_t all = K >= 64 ? ~0ULL : ((1ULL << K) - 1ULL);
    for (int v = 0; v < n; ++v) down[v] = preset[v] = annot[v] ? annot[v] : all;
    for (int v = n - 1; v >= 0; --v)
        if (tr->n_children[v]) {
            uint64_t mc = most_common(down + tr->first_child[v], tr->n_children[v], K);
            uint64_t inter = down[v] & mc;
            if (inter) down[v] = inter;
        }
}

/* acctran, parsimony.py:125-158 */
void orc_pars_acctran(const orc_tree *tr, const uint64_t *down, uint64_t *out) {
    const int n = tr->n_nodes;
    for (int v = 0; v < n; ++v) {
        if (tr->parent[v] < 0) out[v] = down[v];
        for (int c = tr->first_child[v], ce = c + tr->n_children[v]; c < ce; ++c) {
            uint64_t inter = out[v] & down[c];
            out[c] = inter ? inter : down[c];
        }
    }
}

/* downpass, parsimony.py:161-210.  preset[] is updated in place into the DOWNPASS result. */
int orc_pars_downpass(const orc_tree *tr, int K, const uint64_t *down, uint64_t *preset) {
    const int n = tr->n_nodes;
    const uint64_t all = K >= 64 ? ~0ULL : ((1ULL << K) - 1ULL);
    uint64_t *up = (uint64_t *) malloc(sizeof(uint64_t) * n);
    int maxc = 1;
    for (int v = 0; v < n; ++v) if (tr->n_children[v] > maxc) maxc = tr->n_children[v];
    uint64_t *buf = (uint64_t *) malloc(sizeof(uint64_t) * (maxc + 1));
    if (!up || !buf) return ORC_ERR_ALLOC;
    for (int v = 0; v < n; ++v) {
        int p = tr->parent[v];
        if (p < 0) up[v] = all;
        else {
            int q = 0;
            buf[q++] = up[p];
            for (int s = tr->first_child[p], se = s + tr->n_children[p]; s < se; ++s) if (s != v) buf[q++] = down[s];
            up[v] = most_common(buf, q, K);
        }
        uint64_t du;
        if (tr->n_children[v]) {
            int q = 0;
            buf[q++] = up[v];
            for (int c = tr->first_child[v], ce = c + tr->n_children[v]; c < ce; ++c) buf[q++] = down[c];
            du = most_common(buf, q, K);
        } else du = up[v];
        uint64_t inter = du & preset[v];
        if (inter) preset[v] = inter;
    }
    free(up); free(buf);
    return ORC_OK;
}

/* deltran, parsimony.py:213-242 (in place on the DOWNPASS result) */
void orc_pars_deltran(const orc_tree *tr, uint64_t *states) {
    for (int v = 0; v < tr->n_nodes; ++v) {
        int p = tr->parent[v];
        if (p >= 0) {
            uint64_t inter = states[v] & states[p];
            if (inter) states[v] = inter;
        }
    }
}

/* get_num_parsimonious_steps, parsimony.py:360-380, summed over the trees of the forest */
int64_t orc_pars_steps(const orc_tree *tr, int K, const uint64_t *states) {
    const int n = tr->n_nodes;
    const int64_t INF = (int64_t) 1 << 60;
    int64_t *num = (int64_t *) malloc(sizeof(int64_t) * (size_t) n * K);
    int64_t total = 0;
    for (int v = n - 1; v >= 0; --v) {
        int64_t *nv = num + (size_t) v * K;
        for (int s = 0; s < K; ++s) {
            if (!((states[v] >> s) & 1ULL)) { nv[s] = INF; continue; }
            int64_t acc = 0;
            for (int c = tr->first_child[v], ce = c + tr->n_children[v]; c < ce; ++c) {
                const int64_t *nc = num + (size_t) c * K;
                int64_t best = INF;
                for (int cs = 0; cs < K; ++cs)
                    if (nc[cs] < INF) {
                        int64_t x = (cs == s ? 0 : 1) + nc[cs];
                        if (x < best) best = x;
                    }
                acc += best;
            }
            nv[s] = acc;
        }
        if (tr->parent[v] < 0) {
            int64_t best = INF;
            for (int s = 0; s < K; ++s) if (nv[s] < best) best = nv[s];
            total += best;
        }
    }
    free(num);
    return total;
}
