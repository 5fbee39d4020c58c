/* oracle/oracle.h — TEST INFRASTRUCTURE. CPU restatement of the reference hot path (see *.c). */
#ifndef SUBRECON_ORACLE_H
#define SUBRECON_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- PAL 1.51 numerics (pal_numerics.c) */
void pal_check_frequencies(double *f20);
void pal_rate_matrix(const double *q_upper190, const double *freq20, double *R400);
int  pal_eigensystem(const double *R400, double *Eval20, double *Evec400, double *Ievc400);
void pal_transition_probs(const double *Eval, const double *Evec, const double *Ievc, double t, double *P400);
int  pal_gamma_rates(int K, double alpha, double *rates);

/* ---- the per-column task (recon_oracle.c) */
typedef struct {
    int            n_nodes;      /* tree in CSR form, children in Newick order          */
    const int32_t *child_off;    /* [n_nodes+1]                                           */
    const int32_t *child_idx;    /* [child_off[n_nodes]]                                  */
    const double  *blen;         /* [n_nodes] branch length to the parent                 */
    const int32_t *leaf_row;     /* [n_nodes] alignment row for tips, -1 for internal     */
    int            root;
    int            n_taxa;
    int64_t        n_sites;
    const uint8_t *states;       /* [n_taxa][n_sites]; 0..19 amino acid, anything else = missing */
    int            K;
    const double  *rates;        /* [K]                                                   */
    const double  *q_upper;      /* [190] model table (for the jar-faithful modes)        */
    const double  *pi_model;     /* [20] frequencies handed to the per-column model ctor (SubRecon.java:113) */
    const double  *log_pi;       /* [20] ln of SubRecon's pi (SubRecon.java:223)          */
} sro_problem;

/* mode: 0 = P hoisted out of the column loop (same results, fast; "honest CPU", BASELINE.md variant C)
 *       1 = P rebuilt per (edge, rate, column) from one eigensystem       (variant B)
 *       2 = jar-faithful: rate-matrix rebuild + eigendecomposition + P per setDistance (variant A) */
int sro_run(const sro_problem *p, int64_t site0, int64_t n, int mode, int n_threads,
            double *lnl /*[n]*/, double *joint /*[n][400] or NULL*/);

#ifdef __cplusplus
}
#endif
#endif
