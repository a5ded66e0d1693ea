/* pgs_fourdof.h — C-ABI of the reference's ALTERNATIVE edge functors.
 *
 * The reference carries a second family of cost functors that its build keeps switched off: the calls are commented
 * out at src/PoseGraphSLAM.cpp:1551,1630 and the yaw-only representation sits behind `#ifdef __USE_YPR_REP`
 * (:1346-1354,1534-1548,1608-1626), a macro nothing defines.  A maintainer who turns them back on needs the same
 * thing Ceres gives the live functors — residuals and local-parameterisation Jacobians of every edge — so this entry
 * point evaluates them for a whole edge list in one launch:
 *
 *   PGS_FOURDOF_ERROR   FourDOFError                         src/CeresResidues.h:252-335
 *       AutoDiffCostFunction<.,6,4,3,4,3>; blocks (q1,t1,q2,t2); r = w * [dt ; 4 yaw, 10 pitch, 10 roll] of
 *       delta_q = (q1* q2)* q_obs, angles in DEGREES through R2ypr (:226-243).
 *   PGS_FOURDOF_SWITCH  FourDOFErrorWithSwitchingConstraints src/CeresResidues.h:338-425
 *       <.,7,4,3,4,3,1>; r = s * [dt ; 4 yaw, 10 pitch, 10 roll ; (1 - s)]; the weight is stored, not applied (:393).
 *   PGS_FOURDOF_QIN     QinFourDOFWeightError                src/CeresResidues.h:500-546
 *       <.,4,1,3,1,3>; blocks (yaw_i, t_i, yaw_j, t_j), yaw in degrees with AngleLocalParameterization (:440-456);
 *       r = [R(yaw_i, pitch_i, roll_i)^T (t_j - t_i) - t_obs ; NormalizeAngle(yaw_j - yaw_i - relative_yaw) / 10].
 *
 * Jacobians are in the tangent space Ceres' minimiser sees: per quaternion block the 3 columns of
 * EigenQuaternionParameterization (half-angle, left-multiplied — the same convention as pgs.h), per yaw block the one
 * column of AngleLocalParameterization.  Row-major per edge:
 *   PGS_FOURDOF_ERROR   r[6],  J[6][12]  columns [dtheta1(3), dt1(3), dtheta2(3), dt2(3)]
 *   PGS_FOURDOF_SWITCH  r[7],  J[7][13]  columns [dtheta1, dt1, dtheta2, dt2, ds]
 *   PGS_FOURDOF_QIN     r[4],  J[4][8]   columns [dyaw_i, dt_i(3), dyaw_j, dt_j(3)]
 *
 * All pointers are host memory, fp64; the caller allocates the outputs.  Runs on the CUDA device of the handle; there
 * is no CPU fallback (create fails with PGS_ERR_CUDA without a device).  Status codes are pgs.h's. */
#ifndef PGS_FOURDOF_H_
#define PGS_FOURDOF_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct pgs_fourdof_s* pgs_fourdof_handle;

typedef enum pgs_fourdof_kind { PGS_FOURDOF_ERROR = 0, PGS_FOURDOF_SWITCH = 1, PGS_FOURDOF_QIN = 2 } pgs_fourdof_kind;

int pgs_fourdof_create(int32_t device, pgs_fourdof_handle* out);
void pgs_fourdof_destroy(pgs_fourdof_handle h);
const char* pgs_fourdof_last_error(pgs_fourdof_handle h);

typedef struct pgs_fourdof_input {
  int32_t kind;            /* pgs_fourdof_kind */
  int32_t n_nodes;
  const double* rot;       /* ERROR / SWITCH: q[4 n_nodes] x,y,z,w  (_opt_quat_, PoseGraphSLAM.h:153)
                              QIN: ypr[3 n_nodes] degrees (_opt_ypr_, PoseGraphSLAM.cpp:16); only the yaw is a parameter */
  const double* t;         /* [3 n_nodes]  (_opt_t_) */
  int32_t n_edges;
  const int32_t* c1;       /* [n_edges] first  pose of the residual block (q1,t1 / yaw_i,t_i) */
  const int32_t* c2;       /* [n_edges] second pose of the residual block (q2,t2 / yaw_j,t_j) */
  const double* obs_rot;   /* ERROR / SWITCH: q_obs[4 n_edges] = Quaterniond(c1_T_c2 rotation) (:260,346)
                              QIN: [3 n_edges] = (relative_yaw, pitch_i, roll_i) degrees      (:501-502) */
  const double* obs_t;     /* [3 n_edges] c1_T_c2 translation / (t_x,t_y,t_z) */
  const double* weight;    /* ERROR: [n_edges]; ignored otherwise (SWITCH stores it unused, QIN hard-wires 1, :504) */
  const double* sw;        /* SWITCH: [n_edges] switch variable of each block; ignored otherwise */
} pgs_fourdof_input;

/* r: [n_edges][6|7|4], J: [n_edges][72|91|32] (may be NULL: residuals only), cost: 1/2 sum r^2 (may be NULL). */
int pgs_fourdof_evaluate(pgs_fourdof_handle h, const pgs_fourdof_input* in, double* r, double* J, double* cost);
/* device time of the last evaluation's kernel, milliseconds (CUDA events on the handle's stream) */
int pgs_fourdof_last_timing(pgs_fourdof_handle h, double* ms_kernel);

#ifdef __cplusplus
}
#endif
#endif
