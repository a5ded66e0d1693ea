/* pgs_synth.h — deterministic synthetic pose-graph generator of the bench harness.
 *
 * Not part of the reference's interface: the reference is fed by ROS topics
 * (src/NodeDataManager.cpp:23-215).  This produces what those callbacks would have stored —
 * keyframe poses w_T_c with timestamps, loop edges (a, b, b_T_a, weight), kidnap intervals — for
 * BASELINE.json's five configurations, plus ground truth for property checks.  splitmix64-seeded,
 * host only (no CUDA). */
#ifndef PGS_SYNTH_H_
#define PGS_SYNTH_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct pgs_synth_s* pgs_synth_handle;

typedef struct pgs_synth_spec {
  int32_t n_nodes;          /* nodes per world */
  int32_t n_loop;           /* intra-world loop edges (all worlds together) */
  int32_t n_worlds;         /* 1, or 4 for config 4 */
  int32_t n_interworld;     /* inter-world loop edges (config 4: 200) */
  int32_t deadzone_nodes;   /* kidnapped nodes between worlds (config 4: 5) */
  int32_t loop_gap_min;     /* 50 */
  int32_t loop_gap_max;     /* min(N/8, 2000); 0 = that default */
  double outlier_fraction;  /* config 3: 0.10 */
  double odom_sigma_t, odom_sigma_r;   /* 0.02 m, 0.002 rad */
  double loop_sigma_t, loop_sigma_r;   /* 0.01 m, 0.001 rad */
  uint64_t seed;
} pgs_synth_spec;

/* Fills `spec` for BASELINE.json config 1..5 (seed 0xC0FFEE00 + config). Returns 0, or -1 for a bad id. */
int pgs_synth_config(int32_t config, pgs_synth_spec* spec);
int pgs_synth_create(const pgs_synth_spec* spec, pgs_synth_handle* out);
void pgs_synth_destroy(pgs_synth_handle h);
/* total nodes (all worlds + dead zones), loop edges (intra + inter), kidnap intervals */
void pgs_synth_sizes(pgs_synth_handle h, int32_t* n_nodes, int32_t* n_loop, int32_t* n_kidnaps);
/* Any pointer may be NULL.  stamps_ns[N]; q[N][4], t[N][3] = odometry ("manager") poses in their own
 * world frames; gt_q/gt_t = ground truth in one global frame; loop edges a[E], b[E], q_bTa[E][4],
 * t_bTa[E][3], weight[E], is_outlier[E]; kidnap intervals start_ns[K], end_ns[K]. */
void pgs_synth_copy(pgs_synth_handle h, int64_t* stamps_ns, double* q, double* t, double* gt_q, double* gt_t,
                    int32_t* a, int32_t* b, double* q_bTa, double* t_bTa, double* weight, uint8_t* is_outlier,
                    int64_t* kidnap_start_ns, int64_t* kidnap_end_ns);
#ifdef __cplusplus
}
#endif
#endif
