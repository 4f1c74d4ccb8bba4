/* oracle/mjinternal.h -- TEST INFRASTRUCTURE (private structs + small math for the CPU oracle). */
#ifndef MJINTERNAL_H
#define MJINTERNAL_H
#include <math.h>
#include <stddef.h>
#include <string.h>
#include "mjphys.h"

#define OM_MAXNV 24
#define OM_MAXMESH 32
#define MINVAL 1e-15

enum { JNT_FREE = 0, JNT_BALL, JNT_SLIDE, JNT_HINGE };
enum { G_PLANE = 0, G_HFIELD, G_SPHERE, G_CAPSULE, G_ELLIPSOID, G_CYLINDER, G_BOX, G_MESH };
enum { CNSTR_EQUALITY = 0, CNSTR_LIMIT, CNSTR_CONTACT, CNSTR_CONTACT_FRICTION };

#define DECL_D(n) double* n; int n_##n
#define DECL_I(n) int* n; int n_##n

struct OModel {
  int nq, nv, nbody, njnt, ngeom, nsite, nu, neq, nmesh, nmocap, iterations;
  double timestep, tolerance, impratio, meaninertia, gravity[3];
  DECL_D(opt);
  DECL_I(body_parentid); DECL_D(body_pos); DECL_D(body_quat); DECL_I(body_mocapid); DECL_I(body_weldid);
  DECL_I(body_jntnum); DECL_I(body_jntadr); DECL_I(body_dofnum); DECL_I(body_dofadr);
  DECL_D(body_mass); DECL_D(body_ipos); DECL_D(body_iquat); DECL_D(body_inertia); DECL_D(body_invweight0);
  DECL_I(jnt_type); DECL_I(jnt_bodyid); DECL_I(jnt_qposadr); DECL_I(jnt_dofadr); DECL_D(jnt_pos); DECL_D(jnt_axis);
  DECL_D(jnt_range); DECL_I(jnt_limited); DECL_D(jnt_stiffness); DECL_D(jnt_margin); DECL_D(jnt_solref); DECL_D(jnt_solimp);
  DECL_D(qpos0); DECL_D(qpos_spring); DECL_I(dof_jntid); DECL_I(dof_bodyid); DECL_I(dof_parentid); DECL_D(dof_damping);
  DECL_D(dof_armature); DECL_D(dof_invweight0);
  DECL_I(geom_bodyid); DECL_I(geom_type); DECL_D(geom_size); DECL_D(geom_pos); DECL_D(geom_quat); DECL_I(geom_contype);
  DECL_I(geom_conaffinity); DECL_I(geom_condim); DECL_I(geom_priority); DECL_D(geom_friction); DECL_D(geom_solmix);
  DECL_D(geom_solref); DECL_D(geom_solimp); DECL_D(geom_margin); DECL_D(geom_gap); DECL_I(geom_dataid); DECL_D(geom_rbound);
  DECL_I(site_bodyid); DECL_D(site_pos); DECL_D(site_quat);
  DECL_I(actuator_jntid); DECL_D(actuator_kp); DECL_D(actuator_ctrlrange);
  DECL_I(eq_obj1id); DECL_I(eq_obj2id); DECL_D(eq_data); DECL_D(eq_solref); DECL_D(eq_solimp);
  double* mesh_vert[OM_MAXMESH]; int mesh_nvert[OM_MAXMESH];
  int npair; int* pair_g1; int* pair_g2;
  int* body_lastdof;
};

struct OData {
  DECL_D(qpos); DECL_D(qvel); DECL_D(ctrl); DECL_D(mocap_pos); DECL_D(mocap_quat);
  DECL_D(qacc); DECL_D(qacc_warmstart); DECL_D(qacc_smooth);
  DECL_D(xpos); DECL_D(xquat); DECL_D(xmat); DECL_D(xipos); DECL_D(ximat);
  DECL_D(geom_xpos); DECL_D(geom_xmat); DECL_D(site_xpos); DECL_D(site_xmat);
  DECL_D(dof_axis); DECL_D(dof_anchor);
  DECL_D(qM); DECL_D(qL); DECL_D(qfrc_bias); DECL_D(qfrc_passive); DECL_D(qfrc_actuator); DECL_D(qfrc_smooth);
  DECL_D(qfrc_constraint);
  DECL_D(efc_J); DECL_D(efc_pos); DECL_D(efc_margin); DECL_D(efc_D); DECL_D(efc_R); DECL_D(efc_aref); DECL_D(efc_vel);
  DECL_D(efc_force); DECL_D(efc_diagApprox); DECL_D(efc_KBIP);
  int* efc_type; int* efc_id;
  int nefc, ne, ncon, solver_iter;
  long flops;
  OContact contact[OM_MAXCON];
};

/* algorithmic flop counter of the physics passes (mul, add, div, sqrt = 1 each; counted per loop nest with its trip counts).
 * Thread-local accumulator, folded into OData.flops by om_forward / om_step; read through om_data_flops(). */
extern __thread long om_flops_acc;
#define FL(n) (om_flops_acc += (long)(n))
void om_collide(const OModel* m, OData* d);
void om_jac(const OModel* m, const OData* d, double* jacp, double* jacr, const double* point, int body);
int om_chol(double* L, const double* A, int n);
void om_chol_solve(const double* L, double* x, const double* b, int n);

/* ---- small math (row-major 3x3, quaternions w,x,y,z) */
static inline void v3zero(double* a) { a[0] = a[1] = a[2] = 0; }
static inline void v3copy(double* a, const double* b) { a[0] = b[0]; a[1] = b[1]; a[2] = b[2]; }
static inline void v3add(double* r, const double* a, const double* b) { r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; }
static inline void v3sub(double* r, const double* a, const double* b) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
static inline void v3scl(double* r, const double* a, double s) { r[0] = a[0] * s; r[1] = a[1] * s; r[2] = a[2] * s; }
static inline void v3addscl(double* r, const double* a, const double* b, double s) { r[0] = a[0] + b[0] * s; r[1] = a[1] + b[1] * s; r[2] = a[2] + b[2] * s; }
static inline double v3dot(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline double v3norm(const double* a) { return sqrt(v3dot(a, a)); }
static inline void v3cross(double* r, const double* a, const double* b) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline double v3normalize(double* a) {
  double n = v3norm(a);
  if (n < MINVAL) { a[0] = 1; a[1] = a[2] = 0; return 0; }
  a[0] /= n; a[1] /= n; a[2] /= n; return n;
}
static inline void mat_mulvec(double* r, const double* M, const double* v) {
  double x = M[0] * v[0] + M[1] * v[1] + M[2] * v[2], y = M[3] * v[0] + M[4] * v[1] + M[5] * v[2], z = M[6] * v[0] + M[7] * v[1] + M[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void mat_tmulvec(double* r, const double* M, const double* v) {
  double x = M[0] * v[0] + M[3] * v[1] + M[6] * v[2], y = M[1] * v[0] + M[4] * v[1] + M[7] * v[2], z = M[2] * v[0] + M[5] * v[1] + M[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void quat_copy(double* a, const double* b) { a[0] = b[0]; a[1] = b[1]; a[2] = b[2]; a[3] = b[3]; }
static inline void quat_mul(double* r, const double* a, const double* b) {
  double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  double z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
static inline void quat_conj(double* r, const double* a) { r[0] = a[0]; r[1] = -a[1]; r[2] = -a[2]; r[3] = -a[3]; }
static inline void quat_normalize(double* q) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}
static inline void quat2mat(double* R, const double* q) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = w * w + x * x - y * y - z * z; R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z); R[4] = w * w - x * x + y * y - z * z; R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = w * w - x * x - y * y + z * z;
}
static inline void quat_axisangle(double* q, const double* axis, double ang) {
  double s = sin(0.5 * ang);
  q[0] = cos(0.5 * ang); q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}
/* q <- q * exp(h*w/2), w in the body frame  [3P mju_quatIntegrate] */
static inline void quat_integrate(double* q, const double* w, double h) {
  double ax[3] = {w[0], w[1], w[2]};
  double n = v3norm(ax);
  if (n < MINVAL) return;
  ax[0] /= n; ax[1] /= n; ax[2] /= n;
  double dq[4], r[4];
  quat_axisangle(dq, ax, n * h);
  quat_mul(r, q, dq);
  quat_copy(q, r);
  quat_normalize(q);
}
#endif
