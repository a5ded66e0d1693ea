/* pgs_compose.h — C-ABI of the Composer pose assembly.
 *
 * Replaces the body of the reference's 30 Hz Composer::pose_assember_thread loop (src/Composer.cpp:24-209):
 * for every keyframe i the pose the rest of the system sees is
 *   i <= solvedUntil            the optimised pose if the solver has one, else the odometry pose (:69-88);
 *                               dead-zone (kidnapped) keyframes hang off the last pose of the world they left
 *                               through odometry (:89-101);
 *   i >  solvedUntil            the last optimised pose of the same world carried forward through odometry,
 *                               w_T_last * (M_last^-1 * M_i) (:133-136,155-165); other worlds keep their odometry
 *                               pose (:137-139); dead-zone keyframes as above (:140-146);
 *                               nothing solved yet: the odometry pose, moved into the frame of its world's set
 *                               root when that transform is known (:128-132,172-190).
 * One CUDA thread per keyframe, no dependency between keyframes (the one two-level case — a dead-zone keyframe
 * needs the assembled pose of the last keyframe of the previous world — is evaluated inline).
 *
 * All pointers are host memory; 4x4 matrices are row-major double[16]. */
#ifndef PGS_COMPOSE_H_
#define PGS_COMPOSE_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct pgs_compose_s* pgs_compose_handle;

int pgs_compose_create(int32_t device, pgs_compose_handle* out);
void pgs_compose_destroy(pgs_compose_handle h);
const char* pgs_compose_last_error(pgs_compose_handle h);

typedef struct pgs_compose_input {
  int32_t n_nodes;               /* manager->getNodeLen()                                   (Composer.cpp:55) */
  const double* mgr_T;           /* [n_nodes][16] manager->getNodePose(i), w_M_i              (:80,98,128)     */
  const int32_t* world_id;       /* [n_nodes] manager->which_world_is_this(stamp_i); < 0 = dead zone -(k+1) (:57) */
  int32_t n_slam;                /* slam->nNodes(): slam->nodePoseExists(i) <=> i < n_slam    (:73,157)        */
  const double* slam_q;          /* [n_slam][4] x,y,z,w                                       (slam->getNodePose) */
  const double* slam_t;          /* [n_slam][3]                                                                 */
  int32_t solved_until;          /* slam->solvedUntil()                                       (:35)            */
  int32_t solved_until_world;    /* which_world_is_this(stamp of node solved_until)           (:36)            */
  int32_t n_worlds;              /* manager->n_worlds()                                                        */
  const int32_t* world_end;      /* [n_worlds] manager->nodeidx_of_world_i_ended(w)           (:94,141)        */
  const int32_t* world_setid;    /* [n_worlds] worlds->find_setID_of_world_i(w)               (:58)            */
  const uint8_t* ws_exists;      /* [n_worlds] worlds->is_exist(setid(w), w)                  (:177)           */
  const double* ws_T_w;          /* [n_worlds][16] worlds->getPoseBetweenWorlds(setid(w), w)  (:181)           */
} pgs_compose_input;

/* out_T [n_nodes][16]: the assembled pose of every keyframe (the reference's lbm_fullpose / global_lmb; grouping
 * by world_id gives jmb / global_jmb).  Returns 0 or a PGS_ERR_* code of pgs.h. */
int pgs_compose_run(pgs_compose_handle h, const pgs_compose_input* in, double* out_T);
/* device time of the last run's kernel in milliseconds (CUDA events), and of the whole call incl. copies */
int pgs_compose_last_timing(pgs_compose_handle h, double* ms_kernel, double* ms_total);

#ifdef __cplusplus
}
#endif
#endif
