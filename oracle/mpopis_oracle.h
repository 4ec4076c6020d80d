/*
 * mpopis_oracle.h — CPU oracle for the MPPI/MPOPI hot path of sisl/MPOPIS.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under mpopis_b200/ may include, link or call this.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * PARITY UNPINNED: the reference is pure Julia, Julia is not installed in this image and the
 * reference ships no tests, golden vectors or fixtures. This file restates
 * the reference line by line (citations below) and is cross-checked only against an independent
 * numpy restatement (oracle/np_mirror.py) and textbook identities (tests/test_oracle_*.py).
 *
 * The entry points mirror include/mpopis_b200.h one-to-one with the prefix orc_ so the parity
 * tests issue the same call sequence to both.
 */
#ifndef MPOPIS_ORACLE_H
#define MPOPIS_ORACLE_H

#include <stdint.h>
#include "../include/mpopis_b200.h" /* cfg / enum type definitions only */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_handle orc_t;

const char *orc_last_error(void);
int orc_create(const mpopis_cfg_t *cfg, orc_t **out);
int orc_destroy(orc_t *h);
int orc_set_threads(orc_t *h, int nthreads); /* Threads.@threads over k, POL:269 */
int orc_set_car_env(orc_t *h, int32_t n_cars, const double *params18_per_car, double dt, double ddt,
                    const double *trk_x, const double *trk_y, const double *trk_w, int64_t n_trk);
int orc_set_mountaincar_env(orc_t *h, const double *params7, int64_t max_steps);
int orc_set_sigma(orc_t *h, const double *Sigma, int64_t n);
int orc_set_cma(orc_t *h, const mpopis_cma_t *cma, const double *ws, int64_t n_ws);
int orc_seed(orc_t *h, uint64_t seed);
int orc_plan(orc_t *h, const double *state, int64_t env_t, double *U_inout, double *control_out,
             int32_t *its_run_out);
int orc_plan_with_noise(orc_t *h, const double *state, int64_t env_t, double *U_inout,
                        const double *Z, const double *resample_u, double *control_out,
                        int32_t *its_run_out);
int orc_set_external_env(orc_t *h, const double *action_lo, const double *action_hi);
int orc_plan_external(orc_t *h, double *U_inout, mpopis_rollout_fn rollout, void *user, const double *Z,
                      const double *resample_u, double *control_out, int32_t *its_run_out);
int orc_fetch(orc_t *h, double *costs, double *weights, double *E, double *traj);
int orc_fetch_proposal(orc_t *h, double *Sigma_last, double *U_last);
int orc_rollout_costs(orc_t *h, const double *state, int64_t env_t, const double *U,
                      const double *U_orig, const double *E, const double *Sigma_inv,
                      double *costs_out);
int orc_weights(orc_t *h, const double *costs, int64_t K, double lambda, double *w_out);
int orc_track_query(orc_t *h, const double *pos, int64_t n, int32_t *idx_out, int32_t *idx2_out,
                    double *dist_out, uint8_t *within_out);
int orc_env_step(orc_t *h, double *state_inout, const double *action, int64_t *env_t_inout,
                 double *reward_out, uint8_t *done_out);
int orc_env_reward(orc_t *h, const double *state, uint8_t done, double *reward_out);
int orc_sample_normals(orc_t *h, int64_t step, int64_t iteration, double *Z_out);
int orc_cov_estimate(orc_t *h, int32_t sigma_est, const double *X, int64_t p, int64_t n,
                     const double *w, int32_t corrected, double *mean_out, double *cov_out);
int orc_cholesky(orc_t *h, const double *A, int64_t n, double *L_out);
int orc_inv_sqrt(orc_t *h, const double *A, int64_t n, double *C_out);
/* extra introspection for tests: per-iteration trace of the last plan */
int orc_last_shrinkage(orc_t *h, double *lambda_out);
int orc_sortperm(const double *x, int64_t n, int64_t *perm_out); /* stable, 0-based; POL:455 */
/* Philox4x32-10 block (Random123 definition) — known-answer tests pin it. */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif
