/*
 * oracle/recon_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's per-column task, in the reference's own order of operations:
 *   recon()            /root/reference/src/subrecon/recon/JointBranchReconstruction.java:82-140
 *   downTreeMarginal() /root/reference/src/subrecon/recon/JointBranchReconstruction.java:191-247
 *   getLnValues()      /root/reference/src/subrecon/utils/Utils.java:33-39
 *   getLnSumComponents /root/reference/src/subrecon/utils/Utils.java:47-80
 *   TINY_QUANTITY      /root/reference/src/subrecon/Constants.java:32
 *   column loop + pool /root/reference/src/subrecon/SubRecon.java:109-142
 * and of the model calls it makes into PAL (pal_numerics.c). Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library; the CUDA product
 * never does.
 *
 * Three work modes produce bit-identical results (PAL's per-call rebuild is deterministic,
 * SURVEY B.2) and differ only in how much of the reference's redundant work they repeat:
 *   2 "jar-faithful": every setDistance = rate-matrix rebuild + elmhes/eltran/hqr2/luinverse + P
 *   1 "P-rebuild":    every setDistance = P from one eigensystem
 *   0 "hoisted":      P per (edge, rate) computed once per run
 * Build: gcc -O2 -ffp-contract=off -pthread.
 */
#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

#define NS 20
static const double TINY_QUANTITY = 1e-10;   /* Constants.java:32 */

typedef struct {
    const sro_problem *p;
    int mode;
    /* per-thread model state (SubRecon.java:113 gives every task its own model) */
    double freq[NS];            /* frequencies after the per-column model's own checkFrequencies pass */
    double Eval[NS], Evec[NS * NS], Ievc[NS * NS];
    double P[NS * NS];          /* "transProb" of the model */
    /* mode 0: hoisted matrices, [node][K][400] and root [K][400] */
    const double *hoistP;
    const double *hoistRoot;
} model_t;

/* AbstractRateMatrix::setDistance -> handleRebuild (mode 2 only) -> MatrixExponential::setDistance */
static void model_set_distance(model_t *m, double t)
{
    if (m->mode == 2) {
        double R[NS * NS];
        pal_rate_matrix(m->p->q_upper, m->freq, R);
        pal_eigensystem(R, m->Eval, m->Evec, m->Ievc);
    }
    pal_transition_probs(m->Eval, m->Evec, m->Ievc, t, m->P);
}

/* Utils.getLnSumComponents, Utils.java:47-80. Returns NAN where Java would index [-1]
 * (no element > -Double.MAX_VALUE). */
static double ln_sum_components(const double *v, int n)
{
    if (n == 1) return v[0];
    int imax = -1;
    double mx = -1.7976931348623157e308;
    for (int i = 0; i < n; ++i)
        if (v[i] > mx) { mx = v[i]; imax = i; }
    if (imax < 0) return NAN;
    double r = v[imax];
    double sum = 1.0;
    for (int i = 0; i < n; ++i) {
        if (i == imax) continue;
        sum += exp(v[i] - v[imax]);
    }
    r += log(sum);
    return r;
}

/* downTreeMarginal, JointBranchReconstruction.java:191-247 (recursive, like the reference) */
static void down_tree_marginal(model_t *m, int node, int64_t site, double *count, int k, double rate, double *out)
{
    const sro_problem *p = m->p;
    double cond[NS];
    int c0 = p->child_off[node], c1 = p->child_off[node + 1];
    if (c0 == c1) {                                       /* leaf, :194-205 */
        int state = p->states[(int64_t)p->leaf_row[node] * p->n_sites + site];
        for (int i = 0; i < NS; ++i) cond[i] = 0.0;
        if (state < NS) cond[state] = 1.0;
        else for (int i = 0; i < NS; ++i) cond[i] = 1.0;
    } else {
        for (int i = 0; i < NS; ++i) cond[i] = 1.0;       /* :208-210 */
        for (int ic = c0; ic < c1; ++ic) {                /* :212 */
            int child = p->child_idx[ic];
            double P[NS * NS];
            if (m->mode == 0) {
                memcpy(P, m->hoistP + ((int64_t)child * p->K + k) * NS * NS, sizeof(P));
            } else {
                model_set_distance(m, p->blen[child] * rate);   /* :216 */
                memcpy(P, m->P, sizeof(P));                     /* :217 */
            }
            double cc[NS];
            down_tree_marginal(m, child, site, count, k, rate, cc);   /* :219 */
            for (int i = 0; i < NS; ++i) {                /* :221-228 */
                double sum = 0.0;
                for (int j = 0; j < NS; ++j) sum += P[i * NS + j] * cc[j];
                cond[i] *= sum;
            }
        }
    }
    double big = 0.0;                                     /* :236-239 */
    for (int i = 0; i < NS; ++i) big = (cond[i] > big) ? cond[i] : big;
    for (int i = 0; i < NS; ++i) out[i] = cond[i] / (big + TINY_QUANTITY);   /* :241-243 */
    *count += log(big + TINY_QUANTITY);                   /* :244 */
}

/* recon(), JointBranchReconstruction.java:82-140 */
static void recon_column(model_t *m, int64_t site, double *mix /*[400*K] scratch*/, double *lnl, double *joint)
{
    const sro_problem *p = m->p;
    const int K = p->K;
    const int nodeA = p->child_idx[p->child_off[p->root]];
    const int nodeB = p->child_idx[p->child_off[p->root] + 1];
    for (int k = 0; k < K; ++k) {
        double rate = p->rates[k];
        double cA = 0.0, cB = 0.0, a[NS], b[NS], la[NS], lb[NS];
        down_tree_marginal(m, nodeA, site, &cA, k, rate, a);          /* :94 */
        down_tree_marginal(m, nodeB, site, &cB, k, rate, b);          /* :95 */
        for (int i = 0; i < NS; ++i) { la[i] = log(a[i]); lb[i] = log(b[i]); }
        double corr = cA + cB;                                         /* :97 */
        const double *P;
        if (m->mode == 0) P = m->hoistRoot + (int64_t)k * NS * NS;
        else { model_set_distance(m, (p->blen[nodeA] + p->blen[nodeB]) * rate); P = m->P; }   /* :99 */
        for (int ia = 0; ia < NS; ++ia) {
            double lat = p->log_pi[ia] + la[ia];                       /* :103 */
            for (int ib = 0; ib < NS; ++ib) {
                double scaled = lat + log(P[ia * NS + ib]) + lb[ib];   /* :107 */
                mix[ia * NS * K + ib * K + k] = scaled + corr;         /* :108-109, flatIndex :169-172 */
            }
        }
    }
    double log_sum = ln_sum_components(mix, NS * NS * K);              /* :115 */
    if (joint) {
        for (int ia = 0; ia < NS; ++ia)
            for (int ib = 0; ib < NS; ++ib)                            /* :121-131 */
                joint[ia * NS + ib] = exp(ln_sum_components(mix + ia * NS * K + ib * K, K) - log_sum);
    }
    *lnl = log_sum - log((double)K);                                   /* :71, :138 */
}

typedef struct {
    const sro_problem *p;
    int mode;
    int64_t site0, n;
    atomic_llong *next;
    const double *hoistP, *hoistRoot;
    const double *Eval, *Evec, *Ievc, *freq;
    double *lnl, *joint;
} job_t;

static void *worker(void *arg)
{
    job_t *j = (job_t *)arg;
    model_t m;
    memset(&m, 0, sizeof(m));
    m.p = j->p;
    m.mode = j->mode;
    m.hoistP = j->hoistP;
    m.hoistRoot = j->hoistRoot;
    memcpy(m.freq, j->freq, sizeof(m.freq));
    memcpy(m.Eval, j->Eval, sizeof(m.Eval));
    memcpy(m.Evec, j->Evec, sizeof(m.Evec));
    memcpy(m.Ievc, j->Ievc, sizeof(m.Ievc));
    double *mix = (double *)malloc(sizeof(double) * NS * NS * j->p->K);
    for (;;) {
        long long i = atomic_fetch_add(j->next, 1);     /* the pool's FIFO queue, SubRecon.java:109-116 */
        if (i >= j->n) break;
        recon_column(&m, j->site0 + i, mix, j->lnl + i, j->joint ? j->joint + i * NS * NS : NULL);
    }
    free(mix);
    return NULL;
}

int sro_run(const sro_problem *p, int64_t site0, int64_t n, int mode, int n_threads, double *lnl, double *joint)
{
    if (!p || n < 0 || site0 < 0 || site0 + n > p->n_sites || p->K < 1 || mode < 0 || mode > 2) return -1;
    if (p->child_off[p->root + 1] - p->child_off[p->root] != 2) return -2;   /* SubRecon.java:212, :59-60 */
    if (n_threads < 1) n_threads = 1;
    /* the per-column model: new <Model>(pi) -> setFrequencies -> checkFrequencies (second pass) */
    double freq[NS], R[NS * NS], Eval[NS], Evec[NS * NS], Ievc[NS * NS];
    memcpy(freq, p->pi_model, sizeof(freq));
    pal_check_frequencies(freq);
    pal_rate_matrix(p->q_upper, freq, R);
    int rc = pal_eigensystem(R, Eval, Evec, Ievc);
    if (rc) return rc;
    double *hoistP = NULL, *hoistRoot = NULL;
    if (mode == 0) {
        hoistP = (double *)malloc(sizeof(double) * (size_t)p->n_nodes * p->K * NS * NS);
        hoistRoot = (double *)malloc(sizeof(double) * (size_t)p->K * NS * NS);
        if (!hoistP || !hoistRoot) { free(hoistP); free(hoistRoot); return -4; }
        int nodeA = p->child_idx[p->child_off[p->root]], nodeB = p->child_idx[p->child_off[p->root] + 1];
        for (int v = 0; v < p->n_nodes; ++v)
            for (int k = 0; k < p->K; ++k)
                pal_transition_probs(Eval, Evec, Ievc, p->blen[v] * p->rates[k], hoistP + ((size_t)v * p->K + k) * NS * NS);
        for (int k = 0; k < p->K; ++k)
            pal_transition_probs(Eval, Evec, Ievc, (p->blen[nodeA] + p->blen[nodeB]) * p->rates[k], hoistRoot + (size_t)k * NS * NS);
    }
    atomic_llong next = 0;
    job_t job = {p, mode, site0, n, &next, hoistP, hoistRoot, Eval, Evec, Ievc, freq, lnl, joint};
    {   /* always on worker threads: they get a stack deep enough for the reference's recursion */
        pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
        pthread_attr_t attr;
        pthread_attr_init(&attr);
        pthread_attr_setstacksize(&attr, (size_t)1 << 30);   /* deep caterpillars recurse N deep */
        for (int t = 0; t < n_threads; ++t) pthread_create(&th[t], &attr, worker, &job);
        for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
        pthread_attr_destroy(&attr);
        free(th);
    }
    free(hoistP);
    free(hoistRoot);
    return 0;
}
