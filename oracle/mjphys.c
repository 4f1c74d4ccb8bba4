/* oracle/mjphys.c -- TEST INFRASTRUCTURE (see mjphys.h header: PARITY UNPINNED).
 *
 * Scalar float64 restatement of the MuJoCo 3.3.0 forward-dynamics pipeline for
 * the Meta-World MJCF feature set.  Each stage names the MuJoCo routine it
 * restates ([3P] = third-party, source not under /root/reference) and the
 * reference call site that reaches it.
 */
#include "mjphys.h"
#include "mjinternal.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ registry */
typedef struct { const char* name; int is_int; size_t off; size_t noff; } Field;
#define FD(n) {#n, 0, offsetof(OModel, n), offsetof(OModel, n_##n)}
#define FI(n) {#n, 1, offsetof(OModel, n), offsetof(OModel, n_##n)}
static const Field model_fields[] = {
  FD(opt), FI(body_parentid), FD(body_pos), FD(body_quat), FI(body_mocapid), FI(body_weldid),
  FI(body_jntnum), FI(body_jntadr), FI(body_dofnum), FI(body_dofadr),
  FD(body_mass), FD(body_ipos), FD(body_iquat), FD(body_inertia), FD(body_invweight0),
  FI(jnt_type), FI(jnt_bodyid), FI(jnt_qposadr), FI(jnt_dofadr), FD(jnt_pos), FD(jnt_axis), FD(jnt_range),
  FI(jnt_limited), FD(jnt_stiffness), FD(jnt_margin), FD(jnt_solref), FD(jnt_solimp),
  FD(qpos0), FD(qpos_spring), FI(dof_jntid), FI(dof_bodyid), FI(dof_parentid), FD(dof_damping), FD(dof_armature),
  FD(dof_invweight0),
  FI(geom_bodyid), FI(geom_type), FD(geom_size), FD(geom_pos), FD(geom_quat), FI(geom_contype), FI(geom_conaffinity),
  FI(geom_condim), FI(geom_priority), FD(geom_friction), FD(geom_solmix), FD(geom_solref), FD(geom_solimp),
  FD(geom_margin), FD(geom_gap), FI(geom_dataid), FD(geom_rbound),
  FI(site_bodyid), FD(site_pos), FD(site_quat),
  FI(actuator_jntid), FD(actuator_kp), FD(actuator_ctrlrange),
  FI(eq_obj1id), FI(eq_obj2id), FD(eq_data), FD(eq_solref), FD(eq_solimp),
};
#define NFIELD ((int)(sizeof(model_fields) / sizeof(Field)))

OModel* om_model_new(void) { return (OModel*)calloc(1, sizeof(OModel)); }

void om_model_free(OModel* m) {
  if (!m) return;
  for (int i = 0; i < NFIELD; i++) free(*(void**)((char*)m + model_fields[i].off));
  for (int i = 0; i < m->nmesh; i++) free(m->mesh_vert[i]);
  free(m->pair_g1); free(m->pair_g2); free(m->body_lastdof);
  free(m);
}

static int set_field(OModel* m, const char* name, const void* v, int n, int is_int) {
  for (int i = 0; i < NFIELD; i++) {
    const Field* f = &model_fields[i];
    if (strcmp(f->name, name) == 0 && f->is_int == is_int) {
      void** p = (void**)((char*)m + f->off);
      size_t sz = is_int ? sizeof(int) : sizeof(double);
      free(*p);
      *p = malloc(sz * (size_t)(n > 0 ? n : 1));
      memcpy(*p, v, sz * (size_t)n);
      *(int*)((char*)m + f->noff) = n;
      return 0;
    }
  }
  return -1;
}
int om_model_set_f64(OModel* m, const char* name, const double* v, int n) { return set_field(m, name, v, n, 0); }
int om_model_set_i32(OModel* m, const char* name, const int* v, int n) { return set_field(m, name, v, n, 1); }

static void* get_field(OModel* m, const char* name, int* n, int is_int) {
  for (int i = 0; i < NFIELD; i++) {
    const Field* f = &model_fields[i];
    if (strcmp(f->name, name) == 0 && f->is_int == is_int) {
      if (n) *n = *(int*)((char*)m + f->noff);
      return *(void**)((char*)m + f->off);
    }
  }
  return NULL;
}
double* om_model_f64(OModel* m, const char* name, int* n) { return (double*)get_field(m, name, n, 0); }
int* om_model_i32(OModel* m, const char* name, int* n) { return (int*)get_field(m, name, n, 1); }

int om_model_add_mesh(OModel* m, const double* vert, int nvert) {
  if (m->nmesh >= OM_MAXMESH) return -1;
  int id = m->nmesh++;
  m->mesh_vert[id] = NULL;
  m->mesh_nvert[id] = nvert;
  if (nvert > 0) {
    m->mesh_vert[id] = (double*)malloc(sizeof(double) * 3 * (size_t)nvert);
    memcpy(m->mesh_vert[id], vert, sizeof(double) * 3 * (size_t)nvert);
  }
  return id;
}

int om_model_finalize(OModel* m) {
  m->nbody = m->n_body_parentid;
  m->njnt = m->n_jnt_type;
  m->nq = m->n_qpos0;
  m->nv = m->n_dof_jntid;
  m->ngeom = m->n_geom_type;
  m->nsite = m->n_site_bodyid;
  m->nu = m->n_actuator_jntid;
  m->neq = m->n_eq_obj1id;
  if (m->nv > OM_MAXNV || m->n_opt < 8) return -1;
  m->timestep = m->opt[0]; m->tolerance = m->opt[1]; m->impratio = m->opt[2]; m->meaninertia = m->opt[3];
  m->gravity[0] = m->opt[4]; m->gravity[1] = m->opt[5]; m->gravity[2] = m->opt[6];
  m->iterations = (int)m->opt[7];
  m->nmocap = 0;
  for (int b = 0; b < m->nbody; b++) if (m->body_mocapid[b] >= 0) m->nmocap++;
  /* last dof on the path from the world to each body */
  m->body_lastdof = (int*)malloc(sizeof(int) * (size_t)m->nbody);
  m->body_lastdof[0] = -1;
  for (int b = 1; b < m->nbody; b++)
    m->body_lastdof[b] = m->body_dofnum[b] > 0 ? m->body_dofadr[b] + m->body_dofnum[b] - 1
                                                : m->body_lastdof[m->body_parentid[b]];
  /* static candidate pairs: what survives MuJoCo's body/geom filters [3P mj_collision]:
     same weld body, parent-child weld bodies (neither static), contype/conaffinity masks */
  int cap = m->ngeom * m->ngeom / 2 + 1;
  m->pair_g1 = (int*)malloc(sizeof(int) * (size_t)cap);
  m->pair_g2 = (int*)malloc(sizeof(int) * (size_t)cap);
  m->npair = 0;
  for (int g1 = 0; g1 < m->ngeom; g1++)
    for (int g2 = g1 + 1; g2 < m->ngeom; g2++) {
      int b1 = m->geom_bodyid[g1], b2 = m->geom_bodyid[g2];
      int w1 = m->body_weldid[b1], w2 = m->body_weldid[b2];
      if (w1 == w2) continue;
      int p1 = m->body_weldid[m->body_parentid[w1]], p2 = m->body_weldid[m->body_parentid[w2]];
      if (w1 != 0 && w2 != 0 && (w1 == p2 || w2 == p1)) continue;
      if (!((m->geom_contype[g1] & m->geom_conaffinity[g2]) || (m->geom_contype[g2] & m->geom_conaffinity[g1])))
        continue;
      /* order so that type(g1) <= type(g2), as MuJoCo's collision table expects */
      if (m->geom_type[g1] <= m->geom_type[g2]) { m->pair_g1[m->npair] = g1; m->pair_g2[m->npair] = g2; }
      else { m->pair_g1[m->npair] = g2; m->pair_g2[m->npair] = g1; }
      m->npair++;
    }
  return 0;
}

/* ------------------------------------------------------------------ data */
OData* om_data_new(const OModel* m) {
  OData* d = (OData*)calloc(1, sizeof(OData));
  int nb = m->nbody, nv = m->nv;
#define AL(f, n) d->f = (double*)calloc((size_t)((n) > 0 ? (n) : 1), sizeof(double)); d->n_##f = (n)
  AL(qpos, m->nq); AL(qvel, nv); AL(ctrl, m->nu); AL(mocap_pos, 3 * m->nmocap); AL(mocap_quat, 4 * m->nmocap);
  AL(qacc, nv); AL(qacc_warmstart, nv); AL(qacc_smooth, nv);
  AL(xpos, 3 * nb); AL(xquat, 4 * nb); AL(xmat, 9 * nb); AL(xipos, 3 * nb); AL(ximat, 9 * nb);
  AL(geom_xpos, 3 * m->ngeom); AL(geom_xmat, 9 * m->ngeom); AL(site_xpos, 3 * m->nsite); AL(site_xmat, 9 * m->nsite);
  AL(dof_axis, 3 * nv); AL(dof_anchor, 3 * nv);
  AL(qM, nv * nv); AL(qL, nv * nv); AL(qfrc_bias, nv); AL(qfrc_passive, nv); AL(qfrc_actuator, nv);
  AL(qfrc_smooth, nv); AL(qfrc_constraint, nv);
  AL(efc_J, OM_MAXEFC * nv); AL(efc_pos, OM_MAXEFC); AL(efc_margin, OM_MAXEFC); AL(efc_D, OM_MAXEFC);
  AL(efc_R, OM_MAXEFC); AL(efc_aref, OM_MAXEFC); AL(efc_vel, OM_MAXEFC); AL(efc_force, OM_MAXEFC);
  AL(efc_diagApprox, OM_MAXEFC); AL(efc_KBIP, 4 * OM_MAXEFC);
#undef AL
  d->efc_type = (int*)calloc(OM_MAXEFC, sizeof(int));
  d->efc_id = (int*)calloc(OM_MAXEFC, sizeof(int));
  om_reset_data(m, d);
  return d;
}

static const struct { const char* name; size_t off, noff; } data_fields[] = {
#define DF(n) {#n, offsetof(OData, n), offsetof(OData, n_##n)}
  DF(qpos), DF(qvel), DF(ctrl), DF(mocap_pos), DF(mocap_quat), DF(qacc), DF(qacc_warmstart), DF(qacc_smooth),
  DF(xpos), DF(xquat), DF(xmat), DF(xipos), DF(ximat), DF(geom_xpos), DF(geom_xmat), DF(site_xpos), DF(site_xmat),
  DF(qM), DF(qfrc_bias), DF(qfrc_passive), DF(qfrc_actuator), DF(qfrc_smooth), DF(qfrc_constraint),
  DF(efc_J), DF(efc_pos), DF(efc_margin), DF(efc_D), DF(efc_R), DF(efc_aref), DF(efc_vel), DF(efc_force),
  DF(efc_diagApprox), DF(dof_axis), DF(dof_anchor), DF(qL), DF(efc_KBIP),
#undef DF
};

void om_data_free(OData* d) {
  if (!d) return;
  for (size_t i = 0; i < sizeof(data_fields) / sizeof(data_fields[0]); i++) free(*(void**)((char*)d + data_fields[i].off));
  free(d->efc_type); free(d->efc_id);
  free(d);
}

double* om_data_f64(OData* d, const char* name, int* n) {
  for (size_t i = 0; i < sizeof(data_fields) / sizeof(data_fields[0]); i++)
    if (strcmp(data_fields[i].name, name) == 0) {
      if (n) *n = *(int*)((char*)d + data_fields[i].noff);
      return *(double**)((char*)d + data_fields[i].off);
    }
  return NULL;
}
int om_data_ncon(const OData* d) { return d->ncon; }
int om_data_nefc(const OData* d) { return d->nefc; }
const OContact* om_data_contacts(const OData* d) { return d->contact; }
int om_data_solver_iter(const OData* d) { return d->solver_iter; }
long om_data_flops(const OData* d) { return d->flops; }
__thread long om_flops_acc = 0;

/* mj_resetData [3P]: qpos <- qpos0, mocap <- model body pose, everything else zero.
   Run-time edits of model.body_pos / site_pos / eq_data are NOT undone. */
void om_reset_data(const OModel* m, OData* d) {
  for (size_t i = 0; i < sizeof(data_fields) / sizeof(data_fields[0]); i++) {
    double* p = *(double**)((char*)d + data_fields[i].off);
    int n = *(int*)((char*)d + data_fields[i].noff);
    memset(p, 0, sizeof(double) * (size_t)n);
  }
  memcpy(d->qpos, m->qpos0, sizeof(double) * (size_t)m->nq);
  for (int b = 0; b < m->nbody; b++) {
    int id = m->body_mocapid[b];
    if (id >= 0) {
      memcpy(d->mocap_pos + 3 * id, m->body_pos + 3 * b, 3 * sizeof(double));
      memcpy(d->mocap_quat + 4 * id, m->body_quat + 4 * b, 4 * sizeof(double));
    }
  }
  d->ncon = d->nefc = 0;
  d->solver_iter = 0;
}

/* ------------------------------------------------------------------ kinematics  [3P mj_kinematics] */
static void kinematics(const OModel* m, OData* d) {
  /* normalise free-joint quaternions held in qpos */
  for (int j = 0; j < m->njnt; j++)
    if (m->jnt_type[j] == JNT_FREE) quat_normalize(d->qpos + m->jnt_qposadr[j] + 3);
  double* xp = d->xpos; double* xq = d->xquat;
  v3zero(xp); xq[0] = 1; xq[1] = xq[2] = xq[3] = 0;
  quat2mat(d->xmat, xq);
  for (int b = 1; b < m->nbody; b++) {
    int p = m->body_parentid[b];
    double* pos = xp + 3 * b; double* quat = xq + 4 * b;
    int mid = m->body_mocapid[b];
    if (mid >= 0) {
      v3copy(pos, d->mocap_pos + 3 * mid);
      quat_copy(quat, d->mocap_quat + 4 * mid);
      quat_normalize(quat);
    } else {
      double t[3];
      mat_mulvec(t, d->xmat + 9 * p, m->body_pos + 3 * b);
      v3add(pos, xp + 3 * p, t);
      quat_mul(quat, xq + 4 * p, m->body_quat + 4 * b);
    }
    for (int k = 0; k < m->body_jntnum[b]; k++) {
      int j = m->body_jntadr[b] + k, qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
      double R[9];
      if (m->jnt_type[j] == JNT_FREE) {
        v3copy(pos, d->qpos + qa);
        quat_copy(quat, d->qpos + qa + 3);
        quat2mat(R, quat);
        for (int i = 0; i < 3; i++) {
          double* ax = d->dof_axis + 3 * (da + i); v3zero(ax); ax[i] = 1;
          v3copy(d->dof_anchor + 3 * (da + i), pos);
          double* ar = d->dof_axis + 3 * (da + 3 + i);
          ar[0] = R[i]; ar[1] = R[3 + i]; ar[2] = R[6 + i];
          v3copy(d->dof_anchor + 3 * (da + 3 + i), pos);
        }
      } else {
        double axis[3], anchor[3], t[3];
        quat2mat(R, quat);
        mat_mulvec(axis, R, m->jnt_axis + 3 * j);
        mat_mulvec(t, R, m->jnt_pos + 3 * j);
        v3add(anchor, pos, t);
        double q = d->qpos[qa] - m->qpos0[qa];
        if (m->jnt_type[j] == JNT_SLIDE) {
          v3addscl(pos, pos, axis, q);
        } else { /* hinge: rotate about the joint axis, keep the anchor fixed */
          double qr[4], qn[4];
          quat_axisangle(qr, m->jnt_axis + 3 * j, q);
          quat_mul(qn, quat, qr);
          quat_copy(quat, qn);
          quat_normalize(quat);
          quat2mat(R, quat);
          mat_mulvec(t, R, m->jnt_pos + 3 * j);
          v3sub(pos, anchor, t);
        }
        v3copy(d->dof_axis + 3 * da, axis);
        v3copy(d->dof_anchor + 3 * da, anchor);
      }
    }
    quat_normalize(quat);
    quat2mat(d->xmat + 9 * b, quat);
    /* inertial frame */
    double t[3], qi[4];
    mat_mulvec(t, d->xmat + 9 * b, m->body_ipos + 3 * b);
    v3add(d->xipos + 3 * b, pos, t);
    quat_mul(qi, quat, m->body_iquat + 4 * b);
    quat2mat(d->ximat + 9 * b, qi);
  }
  FL(m->nbody * 190 + (m->ngeom + m->nsite) * 75);   /* per body: 2 quat_mul (28) + 3 quat2mat (27) + 3 mat_mulvec (15) + normalise; per geom/site: 15 + 3 + 28 + 27 */
  for (int g = 0; g < m->ngeom; g++) {
    int b = m->geom_bodyid[g]; double t[3], q[4];
    mat_mulvec(t, d->xmat + 9 * b, m->geom_pos + 3 * g);
    v3add(d->geom_xpos + 3 * g, xp + 3 * b, t);
    quat_mul(q, xq + 4 * b, m->geom_quat + 4 * g);
    quat_normalize(q);
    quat2mat(d->geom_xmat + 9 * g, q);
  }
  for (int s = 0; s < m->nsite; s++) {
    int b = m->site_bodyid[s]; double t[3], q[4];
    mat_mulvec(t, d->xmat + 9 * b, m->site_pos + 3 * s);
    v3add(d->site_xpos + 3 * s, xp + 3 * b, t);
    quat_mul(q, xq + 4 * b, m->site_quat + 4 * s);
    quat_normalize(q);
    quat2mat(d->site_xmat + 9 * s, q);
  }
}

static int dof_is_rot(const OModel* m, int dof) {
  int j = m->dof_jntid[dof];
  if (m->jnt_type[j] == JNT_HINGE) return 1;
  if (m->jnt_type[j] == JNT_FREE) return (dof - m->jnt_dofadr[j]) >= 3;
  return 0;
}

/* Jacobian of a world point rigidly attached to `body` [3P mj_jac]: jacp/jacr are 3 x nv, row-major */
void om_jac(const OModel* m, const OData* d, double* jacp, double* jacr, const double* point, int body) {
  int nv = m->nv;
  if (jacp) memset(jacp, 0, sizeof(double) * 3 * (size_t)nv);
  if (jacr) memset(jacr, 0, sizeof(double) * 3 * (size_t)nv);
  for (int dof = m->body_lastdof[body]; dof >= 0; dof = m->dof_parentid[dof]) {
    const double* ax = d->dof_axis + 3 * dof;
    if (dof_is_rot(m, dof)) {
      double r[3], c[3];
      v3sub(r, point, d->dof_anchor + 3 * dof);
      v3cross(c, ax, r);
      if (jacp) for (int i = 0; i < 3; i++) jacp[i * nv + dof] = c[i];
      if (jacr) for (int i = 0; i < 3; i++) jacr[i * nv + dof] = ax[i];
    } else if (jacp) {
      for (int i = 0; i < 3; i++) jacp[i * nv + dof] = ax[i];
    }
  }
}

/* ------------------------------------------------------------------ inertia  [3P mj_crb + mj_factorM] */
static void mass_matrix(const OModel* m, OData* d) {
  int nv = m->nv;
  memset(d->qM, 0, sizeof(double) * (size_t)(nv * nv));
  double jp[3 * OM_MAXNV], jr[3 * OM_MAXNV];
  int chain[OM_MAXNV];
  for (int b = 1; b < m->nbody; b++) {
    double mass = m->body_mass[b];
    const double* in = m->body_inertia + 3 * b;
    if (mass <= 0 && in[0] <= 0 && in[1] <= 0 && in[2] <= 0) continue;
    if (m->body_lastdof[b] < 0) continue;
    om_jac(m, d, jp, jr, d->xipos + 3 * b, b);
    int nc = 0;
    for (int dof = m->body_lastdof[b]; dof >= 0; dof = m->dof_parentid[dof]) chain[nc++] = dof;
    /* world inertia tensor Iw = R diag(in) R^T */
    const double* R = d->ximat + 9 * b;
    double Iw[9];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
        Iw[3 * i + j] = R[3 * i] * in[0] * R[3 * j] + R[3 * i + 1] * in[1] * R[3 * j + 1] + R[3 * i + 2] * in[2] * R[3 * j + 2];
    FL(nc * 12 + 81 + nc * (15 + nc * 13));   /* jacobian columns, R diag R^T, nc x nc accumulation */
    for (int a = 0; a < nc; a++) {
      int da = chain[a];
      double Ija[3];
      for (int i = 0; i < 3; i++) Ija[i] = Iw[3 * i] * jr[da] + Iw[3 * i + 1] * jr[nv + da] + Iw[3 * i + 2] * jr[2 * nv + da];
      for (int c = 0; c < nc; c++) {
        int dc = chain[c];
        double v = mass * (jp[da] * jp[dc] + jp[nv + da] * jp[nv + dc] + jp[2 * nv + da] * jp[2 * nv + dc]) +
                   jr[dc] * Ija[0] + jr[nv + dc] * Ija[1] + jr[2 * nv + dc] * Ija[2];
        d->qM[da * nv + dc] += v;
      }
    }
  }
  for (int i = 0; i < nv; i++) d->qM[i * nv + i] += m->dof_armature[i];
}

/* dense Cholesky A = L L^T (lower), returns 0 on success */
int om_chol(double* L, const double* A, int n) {
  FL((long)n * n * n / 3 + 2 * n * n);
  memcpy(L, A, sizeof(double) * (size_t)(n * n));
  for (int j = 0; j < n; j++) {
    double s = L[j * n + j];
    for (int k = 0; k < j; k++) s -= L[j * n + k] * L[j * n + k];
    if (s <= 1e-300) return -1;
    s = sqrt(s);
    L[j * n + j] = s;
    for (int i = j + 1; i < n; i++) {
      double t = L[i * n + j];
      for (int k = 0; k < j; k++) t -= L[i * n + k] * L[j * n + k];
      L[i * n + j] = t / s;
    }
  }
  return 0;
}
void om_chol_solve(const double* L, double* x, const double* b, int n) {
  FL(2 * n * n + 2 * n);
  for (int i = 0; i < n; i++) {
    double s = b[i];
    for (int k = 0; k < i; k++) s -= L[i * n + k] * x[k];
    x[i] = s / L[i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = x[i];
    for (int k = i + 1; k < n; k++) s -= L[k * n + i] * x[k];
    x[i] = s / L[i * n + i];
  }
}

/* ------------------------------------------------------------------ bias forces  [3P mj_comVel + mj_rne]
   Spatial vectors [angular; linear] taken about the world origin. */
static void cross_motion(double* r, const double* v, const double* s) {
  double a[3], b[3], c[3];
  v3cross(a, v, s);          /* w x s_w */
  v3cross(b, v, s + 3);      /* w x s_v */
  v3cross(c, v + 3, s);      /* v x s_w */
  r[0] = a[0]; r[1] = a[1]; r[2] = a[2];
  r[3] = b[0] + c[0]; r[4] = b[1] + c[1]; r[5] = b[2] + c[2];
}
static void dof_spatial(const OModel* m, const OData* d, int dof, double* S) {
  const double* ax = d->dof_axis + 3 * dof;
  if (dof_is_rot(m, dof)) {
    v3copy(S, ax);
    v3cross(S + 3, d->dof_anchor + 3 * dof, ax); /* velocity of the origin: ax x (0 - anchor) */
  } else {
    v3zero(S); v3copy(S + 3, ax);
  }
}
static void rne_bias(const OModel* m, OData* d) {
  int nb = m->nbody, nv = m->nv;
  double* cvel = (double*)calloc((size_t)(6 * nb), sizeof(double));
  double* cacc = (double*)calloc((size_t)(6 * nb), sizeof(double));
  double* cfrc = (double*)calloc((size_t)(6 * nb), sizeof(double));
  double* Sdot = (double*)calloc((size_t)(6 * (nv > 0 ? nv : 1)), sizeof(double));
  cacc[3] = -m->gravity[0]; cacc[4] = -m->gravity[1]; cacc[5] = -m->gravity[2];
  FL((long)(nb - 1) * 260 + (long)nv * 70);       /* per body: velocity/acceleration propagation, I a + v x* I v, force accumulation; per dof: S, Sdot, projection */
  for (int b = 1; b < nb; b++) {
    int p = m->body_parentid[b];
    double v[6], a[6];
    memcpy(v, cvel + 6 * p, sizeof(v));
    memcpy(a, cacc + 6 * p, sizeof(a));
    for (int k = 0; k < m->body_jntnum[b]; k++) {
      int j = m->body_jntadr[b] + k, da = m->jnt_dofadr[j];
      if (m->jnt_type[j] == JNT_FREE) {
        double S[6];
        for (int i = 0; i < 3; i++) { /* world-fixed translation axes: S' = v x S */
          dof_spatial(m, d, da + i, S);
          cross_motion(Sdot + 6 * (da + i), v, S);
          for (int c = 0; c < 6; c++) v[c] += S[c] * d->qvel[da + i];
        }
        double vb[6]; memcpy(vb, v, sizeof(vb));
        for (int i = 3; i < 6; i++) { /* body-fixed rotation axes: all three see the same velocity */
          dof_spatial(m, d, da + i, S);
          cross_motion(Sdot + 6 * (da + i), vb, S);
          for (int c = 0; c < 6; c++) v[c] += S[c] * d->qvel[da + i];
        }
        for (int i = 0; i < 6; i++) for (int c = 0; c < 6; c++) a[c] += Sdot[6 * (da + i) + c] * d->qvel[da + i];
      } else {
        double S[6];
        dof_spatial(m, d, da, S);
        cross_motion(Sdot + 6 * da, v, S);
        for (int c = 0; c < 6; c++) { v[c] += S[c] * d->qvel[da]; a[c] += Sdot[6 * da + c] * d->qvel[da]; }
      }
    }
    memcpy(cvel + 6 * b, v, sizeof(v));
    memcpy(cacc + 6 * b, a, sizeof(a));
    /* body force f = I a + v x* (I v) */
    double mass = m->body_mass[b];
    const double* in = m->body_inertia + 3 * b; const double* R = d->ximat + 9 * b; const double* c = d->xipos + 3 * b;
    double Iw[9];
    for (int i = 0; i < 3; i++)
      for (int j2 = 0; j2 < 3; j2++)
        Iw[3 * i + j2] = R[3 * i] * in[0] * R[3 * j2] + R[3 * i + 1] * in[1] * R[3 * j2 + 1] + R[3 * i + 2] * in[2] * R[3 * j2 + 2];
    double t[3], pl[3], L[3], pa[3], La[3], u[3];
    /* momentum */
    v3cross(t, v, c); for (int i = 0; i < 3; i++) pl[i] = mass * (v[3 + i] + t[i]);
    mat_mulvec(L, Iw, v); v3cross(t, c, pl); v3add(L, L, t);
    /* I a */
    v3cross(t, a, c); for (int i = 0; i < 3; i++) pa[i] = mass * (a[3 + i] + t[i]);
    mat_mulvec(La, Iw, a); v3cross(t, c, pa); v3add(La, La, t);
    double* f = cfrc + 6 * b;
    v3cross(t, v, L); v3cross(u, v + 3, pl);
    for (int i = 0; i < 3; i++) f[i] = La[i] + t[i] + u[i];
    v3cross(t, v, pl);
    for (int i = 0; i < 3; i++) f[3 + i] = pa[i] + t[i];
  }
  for (int b = nb - 1; b > 0; b--) {
    int p = m->body_parentid[b];
    for (int c = 0; c < 6; c++) cfrc[6 * p + c] += cfrc[6 * b + c];
  }
  for (int dof = 0; dof < nv; dof++) {
    double S[6]; dof_spatial(m, d, dof, S);
    const double* f = cfrc + 6 * m->dof_bodyid[dof];
    double s = 0; for (int c = 0; c < 6; c++) s += S[c] * f[c];
    d->qfrc_bias[dof] = s;
  }
  free(cvel); free(cacc); free(cfrc); free(Sdot);
}

/* ------------------------------------------------------------------ passive + actuation  [3P mj_passive, mj_fwdActuation] */
static void passive_actuation(const OModel* m, OData* d) {
  int nv = m->nv;
  for (int i = 0; i < nv; i++) { d->qfrc_passive[i] = -m->dof_damping[i] * d->qvel[i]; d->qfrc_actuator[i] = 0; }
  for (int j = 0; j < m->njnt; j++) {
    if (m->jnt_type[j] == JNT_FREE || m->jnt_stiffness[j] == 0) continue;
    int qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
    d->qfrc_passive[da] -= m->jnt_stiffness[j] * (d->qpos[qa] - m->qpos_spring[qa]);
  }
  for (int u = 0; u < m->nu; u++) {
    int j = m->actuator_jntid[u];
    double c = d->ctrl[u], lo = m->actuator_ctrlrange[2 * u], hi = m->actuator_ctrlrange[2 * u + 1];
    if (c < lo) c = lo; if (c > hi) c = hi;
    d->qfrc_actuator[m->jnt_dofadr[j]] += m->actuator_kp[u] * (c - d->qpos[m->jnt_qposadr[j]]);
  }
}

/* ------------------------------------------------------------------ constraint rows  [3P mj_makeConstraint] */
static int add_row(const OModel* m, OData* d, int type, int id, double pos, double margin, double diag,
                   const double* solref, const double* solimp) {
  if (d->nefc >= OM_MAXEFC) return -1;
  int i = d->nefc++;
  memset(d->efc_J + (size_t)i * m->nv, 0, sizeof(double) * (size_t)m->nv);
  d->efc_type[i] = type; d->efc_id[i] = id; d->efc_pos[i] = pos; d->efc_margin[i] = margin;
  d->efc_diagApprox[i] = diag;
  /* impedance + stiffness/damping  [3P mj_makeImpedance: getimpedance / solref->K,B] */
  double imp_lo = solimp[0], imp_hi = solimp[1], width = solimp[2], mid = solimp[3], power = solimp[4];
  if (imp_lo < 0.0001) imp_lo = 0.0001; if (imp_lo > 0.9999) imp_lo = 0.9999;
  if (imp_hi < 0.0001) imp_hi = 0.0001; if (imp_hi > 0.9999) imp_hi = 0.9999;
  if (width < 0) width = 0;
  if (mid < 0.0001) mid = 0.0001; if (mid > 0.9999) mid = 0.9999;
  if (power < 1) power = 1;
  double imp;
  if (imp_lo == imp_hi || width <= MINVAL) imp = 0.5 * (imp_lo + imp_hi);
  else {
    double x = fabs(pos - margin) / width;
    if (x >= 1) imp = imp_hi;
    else if (x <= 0) imp = imp_lo;
    else {
      double y;
      if (power == 1) y = x;
      else if (x <= mid) y = pow(x, power) / pow(mid, power - 1);
      else y = 1 - pow(1 - x, power) / pow(1 - mid, power - 1);
      imp = imp_lo + y * (imp_hi - imp_lo);
    }
  }
  double K, B;
  if (solref[0] > 0) {
    double tc = solref[0], dr = solref[1];
    if (tc < 2 * m->timestep) tc = 2 * m->timestep; /* refsafe */
    K = 1.0 / fmax(MINVAL, imp_hi * imp_hi * tc * tc * dr * dr);
    B = 2.0 / fmax(MINVAL, imp_hi * tc);
  } else {
    K = -solref[0] / fmax(MINVAL, imp_hi * imp_hi);
    B = -solref[1] / fmax(MINVAL, imp_hi);
  }
  d->efc_KBIP[4 * i] = K; d->efc_KBIP[4 * i + 1] = B; d->efc_KBIP[4 * i + 2] = imp; d->efc_KBIP[4 * i + 3] = 0;
  d->efc_R[i] = fmax(MINVAL, (1 - imp) * diag / imp);
  return i;
}

static void make_constraints(const OModel* m, OData* d) {
  int nv = m->nv;
  d->nefc = 0; d->ne = 0;
  double jp1[3 * OM_MAXNV], jr1[3 * OM_MAXNV], jp2[3 * OM_MAXNV], jr2[3 * OM_MAXNV];
  /* ---- weld equalities  [3P mj_instantiateEquality, mjEQ_WELD]; body1 = mocap, body2 = hand */
  for (int e = 0; e < m->neq; e++) {
    int b1 = m->eq_obj1id[e], b2 = m->eq_obj2id[e];
    const double* data = m->eq_data + 11 * e;
    double p1[3], p2[3], t[3], cpos[6];
    mat_mulvec(t, d->xmat + 9 * b1, data + 3); v3add(p1, d->xpos + 3 * b1, t);
    mat_mulvec(t, d->xmat + 9 * b2, data + 0); v3add(p2, d->xpos + 3 * b2, t);
    v3sub(cpos, p1, p2);
    om_jac(m, d, jp1, jr1, p1, b1);
    om_jac(m, d, jp2, jr2, p2, b2);
    double torquescale = data[10];
    double q[4], q1n[4], q2[4];
    quat_mul(q, d->xquat + 4 * b1, data + 6);   /* q = quat(body1) * relpose */
    quat_conj(q1n, d->xquat + 4 * b2);          /* conj(quat(body2)) */
    quat_mul(q2, q1n, q);
    for (int i = 0; i < 3; i++) cpos[3 + i] = torquescale * q2[1 + i];
    double diag_t = m->body_invweight0[2 * b1] + m->body_invweight0[2 * b2];
    double diag_r = m->body_invweight0[2 * b1 + 1] + m->body_invweight0[2 * b2 + 1];
    int r0 = d->nefc;
    for (int k = 0; k < 6; k++)
      add_row(m, d, CNSTR_EQUALITY, e, cpos[k], 0, k < 3 ? diag_t : diag_r, m->eq_solref + 2 * e, m->eq_solimp + 5 * e);
    for (int dof = 0; dof < nv; dof++) {
      for (int k = 0; k < 3; k++) d->efc_J[(size_t)(r0 + k) * nv + dof] = jp1[k * nv + dof] - jp2[k * nv + dof];
      /* rotation rows: 0.5 * conj(q_body2) * (jacr1 - jacr2) * q_body1 * relpose, scaled */
      double ax[4] = {0, jr1[dof] - jr2[dof], jr1[nv + dof] - jr2[nv + dof], jr1[2 * nv + dof] - jr2[2 * nv + dof]};
      double a[4], b[4];
      quat_mul(a, q1n, ax);
      quat_mul(b, a, q);
      for (int k = 0; k < 3; k++) d->efc_J[(size_t)(r0 + 3 + k) * nv + dof] = 0.5 * torquescale * b[1 + k];
    }
    d->ne += 6;
  }
  /* ---- joint limits  [3P mj_instantiateLimit] */
  for (int j = 0; j < m->njnt; j++) {
    if (!m->jnt_limited[j] || m->jnt_type[j] == JNT_FREE) continue;
    int qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
    for (int side = -1; side <= 1; side += 2) {
      double dist = side * (m->jnt_range[2 * j + (side + 1) / 2] - d->qpos[qa]);
      if (dist < m->jnt_margin[j]) {
        int r = add_row(m, d, CNSTR_LIMIT, j, dist, m->jnt_margin[j], m->dof_invweight0[da], m->jnt_solref + 2 * j, m->jnt_solimp + 5 * j);
        if (r >= 0) d->efc_J[(size_t)r * nv + da] = -side;
      }
    }
  }
  /* ---- contacts, elliptic cones  [3P mj_instantiateContact] */
  for (int c = 0; c < d->ncon; c++) {
    OContact* con = d->contact + c;
    con->efc_address = -1;
    if (con->dist >= con->includemargin) continue; /* excluded by gap */
    int b1 = m->geom_bodyid[con->geom1], b2 = m->geom_bodyid[con->geom2];
    om_jac(m, d, jp1, jr1, con->pos, b1);
    om_jac(m, d, jp2, jr2, con->pos, b2);
    FL(2 * nv * 12 + nv * (6 + con->dim * 6) + con->dim * (2 * nv + 40));   /* two point Jacobians, frame projection, row velocity + impedance */
    double tran = m->body_invweight0[2 * b1] + m->body_invweight0[2 * b2];
    double rot = m->body_invweight0[2 * b1 + 1] + m->body_invweight0[2 * b2 + 1];
    int r0 = d->nefc;
    if (r0 + con->dim > OM_MAXEFC) break;
    con->efc_address = r0;
    for (int k = 0; k < con->dim; k++)
      add_row(m, d, k == 0 ? CNSTR_CONTACT : CNSTR_CONTACT_FRICTION, c, k == 0 ? con->dist : 0, k == 0 ? con->includemargin : 0,
              k < 3 ? tran : rot, con->solref, con->solimp);
    for (int dof = 0; dof < nv; dof++) {
      double dp[3] = {jp2[dof] - jp1[dof], jp2[nv + dof] - jp1[nv + dof], jp2[2 * nv + dof] - jp1[2 * nv + dof]};
      double dr[3] = {jr2[dof] - jr1[dof], jr2[nv + dof] - jr1[nv + dof], jr2[2 * nv + dof] - jr1[2 * nv + dof]};
      for (int k = 0; k < con->dim && k < 3; k++) d->efc_J[(size_t)(r0 + k) * nv + dof] = v3dot(con->frame + 3 * k, dp);
      for (int k = 3; k < con->dim; k++) d->efc_J[(size_t)(r0 + k) * nv + dof] = v3dot(con->frame + 3 * (k - 3), dr);
    }
    /* friction rows share the normal row's regulariser (impratio, anisotropy)  [3P mj_makeImpedance] */
    if (con->dim > 1) {
      d->efc_R[r0 + 1] = d->efc_R[r0] / m->impratio;
      con->mu = con->friction[0] * sqrt(d->efc_R[r0 + 1] / d->efc_R[r0]);
      for (int k = 2; k < con->dim; k++)
        d->efc_R[r0 + k] = d->efc_R[r0 + 1] * con->friction[0] * con->friction[0] / (con->friction[k - 1] * con->friction[k - 1]);
    }
  }
  /* ---- reference acceleration  [3P mj_referenceConstraint] */
  for (int i = 0; i < d->nefc; i++) {
    double v = 0;
    for (int dof = 0; dof < nv; dof++) v += d->efc_J[(size_t)i * nv + dof] * d->qvel[dof];
    d->efc_vel[i] = v;
    d->efc_D[i] = 1.0 / d->efc_R[i];
    d->efc_aref[i] = -d->efc_KBIP[4 * i + 1] * v - d->efc_KBIP[4 * i] * d->efc_KBIP[4 * i + 2] * (d->efc_pos[i] - d->efc_margin[i]);
  }
}

/* ------------------------------------------------------------------ solver  [3P mj_solNewton: primal, exact Hessian, exact line search]
   minimise  1/2 (a - a0)^T M (a - a0) + sum_i s_i(J a - aref)  over accelerations a. */
typedef struct {
  double cost;       /* constraint cost s */
  double gauss;
} Cost;

/* evaluates the constraint cost at jar; optionally forces (f, may be NULL) and the Hessian contribution
   accumulated into H (nv x nv, may be NULL) */
static double constraint_eval(const OModel* m, const OData* d, const double* jar, double* f, double* H) {
  int nv = m->nv;
  double cost = 0;
  FL(d->nefc * 12 + (H ? (long)d->nefc * nv * nv * 2 : 0));     /* cone evaluation per row; J^T Hc J accumulation when the Hessian is requested */
  for (int i = 0; i < d->nefc; i++) {
    int type = d->efc_type[i];
    double D = d->efc_D[i];
    if (type == CNSTR_EQUALITY || type == CNSTR_LIMIT) {
      int active = (type == CNSTR_EQUALITY) || jar[i] < 0;
      if (f) f[i] = active ? -D * jar[i] : 0;
      if (active) {
        cost += 0.5 * D * jar[i] * jar[i];
        if (H) {
          const double* J = d->efc_J + (size_t)i * nv;
          for (int a = 0; a < nv; a++) if (J[a] != 0) for (int b = 0; b < nv; b++) H[a * nv + b] += D * J[a] * J[b];
        }
      }
      continue;
    }
    /* elliptic contact block starting at the normal row */
    const OContact* con = d->contact + d->efc_id[i];
    int dim = con->dim;
    double mu = con->mu;
    if (dim == 1) {
      int active = jar[i] < 0;
      if (f) f[i] = active ? -D * jar[i] : 0;
      if (active) {
        cost += 0.5 * D * jar[i] * jar[i];
        if (H) { const double* J = d->efc_J + (size_t)i * nv; for (int a = 0; a < nv; a++) for (int b = 0; b < nv; b++) H[a * nv + b] += D * J[a] * J[b]; }
      }
      continue;
    }
    double fr[6], u[6];
    double N = jar[i] * mu, T2 = 0;
    for (int k = 1; k < dim; k++) { fr[k] = con->friction[k - 1]; u[k] = jar[i + k] * fr[k]; T2 += u[k] * u[k]; }
    double T = sqrt(T2);
    double Hc[36]; int haveH = 0;
    memset(Hc, 0, sizeof(Hc));
    if (N >= mu * T || (T <= 0 && N >= 0)) {
      /* top zone: inside the dual cone, no force */
      if (f) for (int k = 0; k < dim; k++) f[i + k] = 0;
    } else if (mu * N + T <= 0 || (T <= 0 && N < 0)) {
      /* bottom zone: fully quadratic */
      for (int k = 0; k < dim; k++) {
        double Dk = d->efc_D[i + k];
        cost += 0.5 * Dk * jar[i + k] * jar[i + k];
        if (f) f[i + k] = -Dk * jar[i + k];
        Hc[k * 6 + k] = Dk;
      }
      haveH = 1;
    } else {
      /* middle zone: distance to the cone surface */
      double Dm = D / fmax(MINVAL, mu * mu * (1 + mu * mu));
      double NmT = N - mu * T;
      cost += 0.5 * Dm * NmT * NmT;
      double g[6];
      g[0] = mu;
      for (int k = 1; k < dim; k++) g[k] = -mu * fr[k] * u[k] / T;
      if (f) for (int k = 0; k < dim; k++) f[i + k] = -Dm * NmT * g[k];
      if (H) {
        for (int a = 0; a < dim; a++) for (int b = 0; b < dim; b++) Hc[a * 6 + b] = Dm * g[a] * g[b];
        double s = -Dm * NmT * mu; /* >= 0 */
        for (int a = 1; a < dim; a++)
          for (int b = 1; b < dim; b++) {
            double t2 = -(fr[a] * u[a]) * (fr[b] * u[b]) / (T * T * T);
            if (a == b) t2 += fr[a] * fr[a] / T;
            Hc[a * 6 + b] += s * t2;
          }
      }
      haveH = 1;
    }
    if (H && haveH) {
      for (int a = 0; a < dim; a++)
        for (int b = 0; b < dim; b++) {
          double h = Hc[a * 6 + b];
          if (h == 0) continue;
          const double* Ja = d->efc_J + (size_t)(i + a) * nv; const double* Jb = d->efc_J + (size_t)(i + b) * nv;
          for (int x = 0; x < nv; x++) if (Ja[x] != 0) for (int y = 0; y < nv; y++) H[x * nv + y] += h * Ja[x] * Jb[y];
        }
    }
    i += dim - 1;
  }
  return cost;
}

/* derivatives of the line-search objective along jar + alpha*jv (constraint part only) */
static void linesearch_eval(const OModel* m, const OData* d, const double* jar, const double* jv, double alpha,
                            double* cost, double* d1, double* d2) {
  double c = 0, g = 0, h = 0;
  FL(d->nefc * 14);
  (void)m;
  for (int i = 0; i < d->nefc; i++) {
    int type = d->efc_type[i];
    double D = d->efc_D[i];
    double x = jar[i] + alpha * jv[i];
    if (type == CNSTR_EQUALITY || type == CNSTR_LIMIT || (type == CNSTR_CONTACT && d->contact[d->efc_id[i]].dim == 1)) {
      if (type == CNSTR_EQUALITY || x < 0) { c += 0.5 * D * x * x; g += D * x * jv[i]; h += D * jv[i] * jv[i]; }
      continue;
    }
    const OContact* con = d->contact + d->efc_id[i];
    int dim = con->dim; double mu = con->mu;
    double N = x * mu, Np = jv[i] * mu, T2 = 0, xv = 0, vv = 0;
    for (int k = 1; k < dim; k++) {
      double fk = con->friction[k - 1], xk = jar[i + k] + alpha * jv[i + k];
      T2 += fk * fk * xk * xk; xv += fk * fk * xk * jv[i + k]; vv += fk * fk * jv[i + k] * jv[i + k];
    }
    double T = sqrt(T2);
    if (N >= mu * T || (T <= 0 && N >= 0)) {
    } else if (mu * N + T <= 0 || (T <= 0 && N < 0)) {
      for (int k = 0; k < dim; k++) {
        double Dk = d->efc_D[i + k], xk = jar[i + k] + alpha * jv[i + k];
        c += 0.5 * Dk * xk * xk; g += Dk * xk * jv[i + k]; h += Dk * jv[i + k] * jv[i + k];
      }
    } else {
      double Dm = D / fmax(MINVAL, mu * mu * (1 + mu * mu));
      double NmT = N - mu * T, Tp = xv / T, Tpp = vv / T - xv * xv / (T * T * T);
      double r = Np - mu * Tp;
      c += 0.5 * Dm * NmT * NmT; g += Dm * NmT * r; h += Dm * (r * r - NmT * mu * Tpp);
    }
    i += dim - 1;
  }
  *cost = c; *d1 = g; *d2 = h;
}

static void mat_vec(double* y, const double* A, const double* x, int n) {
  FL(2 * n * n);
  for (int i = 0; i < n; i++) { double s = 0; for (int j = 0; j < n; j++) s += A[i * n + j] * x[j]; y[i] = s; }
}

static void solve_constraints(const OModel* m, OData* d) {
  int nv = m->nv, ne = d->nefc;
  double scale = 1.0 / (m->meaninertia * (nv > 1 ? nv : 1));
  double qacc[OM_MAXNV], Ma[OM_MAXNV], grad[OM_MAXNV], search[OM_MAXNV], Ms[OM_MAXNV], tmp[OM_MAXNV];
  double H[OM_MAXNV * OM_MAXNV], L[OM_MAXNV * OM_MAXNV];
  double* jar = (double*)malloc(sizeof(double) * (size_t)(ne + 1));
  double* jv = (double*)malloc(sizeof(double) * (size_t)(ne + 1));
  double* f = d->efc_force;
  /* warm start: keep the previous acceleration only if it beats the unconstrained one  [3P mj_warmstart] */
  double best = 0;
  for (int pass = 0; pass < 2; pass++) {
    const double* a = pass == 0 ? d->qacc_warmstart : d->qacc_smooth;
    for (int i = 0; i < ne; i++) {
      double s = 0; for (int k = 0; k < nv; k++) s += d->efc_J[(size_t)i * nv + k] * a[k];
      jar[i] = s - d->efc_aref[i];
    }
    mat_vec(Ma, d->qM, a, nv);
    double gauss = 0;
    for (int k = 0; k < nv; k++) gauss += 0.5 * (a[k] - d->qacc_smooth[k]) * (Ma[k] - d->qfrc_smooth[k]);
    double c = gauss + constraint_eval(m, d, jar, NULL, NULL);
    if (pass == 0 || c < best) { best = c; memcpy(qacc, a, sizeof(double) * (size_t)nv); }
  }
  mat_vec(Ma, d->qM, qacc, nv);
  for (int i = 0; i < ne; i++) {
    double s = 0; for (int k = 0; k < nv; k++) s += d->efc_J[(size_t)i * nv + k] * qacc[k];
    jar[i] = s - d->efc_aref[i];
  }
  double cost = best;
  int iter = 0;
  FL(3L * 2 * ne * nv + 6 * nv);            /* J a for both warm-start candidates and the selected one */
  for (; iter < m->iterations; iter++) {
    FL(2L * 2 * ne * nv + 12 * nv + 2 * ne);   /* gradient J^T f, J search, updates */
    memcpy(H, d->qM, sizeof(double) * (size_t)(nv * nv));
    constraint_eval(m, d, jar, f, H);
    double gn = 0;
    for (int k = 0; k < nv; k++) {
      double s = Ma[k] - d->qfrc_smooth[k];
      for (int i = 0; i < ne; i++) s -= d->efc_J[(size_t)i * nv + k] * f[i];
      grad[k] = s; gn += s * s;
    }
    if (scale * sqrt(gn) < m->tolerance) break;
    if (om_chol(L, H, nv) != 0) break;
    om_chol_solve(L, search, grad, nv);
    for (int k = 0; k < nv; k++) search[k] = -search[k];
    mat_vec(Ms, d->qM, search, nv);
    for (int i = 0; i < ne; i++) {
      double s = 0; for (int k = 0; k < nv; k++) s += d->efc_J[(size_t)i * nv + k] * search[k];
      jv[i] = s;
    }
    /* exact line search on phi(alpha) (convex, C1) by safeguarded Newton */
    double c1 = 0, c2 = 0;
    for (int k = 0; k < nv; k++) { c1 += search[k] * (Ma[k] - d->qfrc_smooth[k]); c2 += search[k] * Ms[k]; }
    double lo = 0, hi = -1, alpha = 0, cc, g1, g2;
    linesearch_eval(m, d, jar, jv, 0, &cc, &g1, &g2);
    double p1 = c1 + g1, p2 = c2 + g2;
    if (p1 >= 0) break; /* not a descent direction: converged to round-off */
    double p1_0 = p1;
    alpha = -p1 / p2;
    for (int ls = 0; ls < 100; ls++) {
      linesearch_eval(m, d, jar, jv, alpha, &cc, &g1, &g2);
      p1 = c1 + alpha * c2 + g1; p2 = c2 + g2;
      if (fabs(p1) < 1e-14 * fabs(p1_0) + 1e-300) break;
      if (p1 < 0) lo = alpha; else hi = alpha;
      double an = alpha - p1 / p2;
      if (hi > 0 && (an <= lo || an >= hi)) an = 0.5 * (lo + hi);
      else if (hi < 0 && an <= lo) an = 2 * alpha + 1e-12;
      if (an == alpha) break;
      alpha = an;
    }
    for (int k = 0; k < nv; k++) { qacc[k] += alpha * search[k]; Ma[k] += alpha * Ms[k]; }
    for (int i = 0; i < ne; i++) jar[i] += alpha * jv[i];
    double gauss = 0;
    for (int k = 0; k < nv; k++) gauss += 0.5 * (qacc[k] - d->qacc_smooth[k]) * (Ma[k] - d->qfrc_smooth[k]);
    double newcost = gauss + constraint_eval(m, d, jar, NULL, NULL);
    double improvement = scale * (cost - newcost);
    cost = newcost;
    if (improvement < m->tolerance) { iter++; break; }
  }
  constraint_eval(m, d, jar, f, NULL);
  d->solver_iter = iter;
  FL(2L * ne * nv);
  memcpy(d->qacc, qacc, sizeof(double) * (size_t)nv);
  for (int k = 0; k < nv; k++) {
    double s = 0; for (int i = 0; i < ne; i++) s += d->efc_J[(size_t)i * nv + k] * f[i];
    d->qfrc_constraint[k] = s;
  }
  (void)tmp;
  free(jar); free(jv);
}

/* ------------------------------------------------------------------ pipeline */
void om_forward(const OModel* m, OData* d) {
  int nv = m->nv;
  om_flops_acc = 0;
  kinematics(m, d);
  mass_matrix(m, d);
  om_chol(d->qL, d->qM, nv);
  om_collide(m, d);
  make_constraints(m, d);
  rne_bias(m, d);
  passive_actuation(m, d);
  for (int i = 0; i < nv; i++) d->qfrc_smooth[i] = d->qfrc_passive[i] - d->qfrc_bias[i] + d->qfrc_actuator[i];
  om_chol_solve(d->qL, d->qacc_smooth, d->qfrc_smooth, nv);
  if (d->nefc == 0) {
    memcpy(d->qacc, d->qacc_smooth, sizeof(double) * (size_t)nv);
    memset(d->qfrc_constraint, 0, sizeof(double) * (size_t)nv);
    d->solver_iter = 0;
  } else {
    solve_constraints(m, d);
  }
  FL(6 * 2 * nv + 6 * 60 + 8 * nv);          /* weld rows (6 x nv Jacobian, impedance), passive + actuation + qfrc_smooth */
  d->flops += om_flops_acc; om_flops_acc = 0;
}

/* semi-implicit Euler with joint damping treated implicitly  [3P mj_Euler] */
static void euler(const OModel* m, OData* d) {
  int nv = m->nv; double h = m->timestep;
  double A[OM_MAXNV * OM_MAXNV], L[OM_MAXNV * OM_MAXNV], rhs[OM_MAXNV], acc[OM_MAXNV];
  memcpy(A, d->qM, sizeof(double) * (size_t)(nv * nv));
  for (int i = 0; i < nv; i++) { A[i * nv + i] += h * m->dof_damping[i]; rhs[i] = d->qfrc_smooth[i] + d->qfrc_constraint[i]; }
  om_chol(L, A, nv);
  om_chol_solve(L, acc, rhs, nv);
  for (int i = 0; i < nv; i++) d->qvel[i] += h * acc[i];
  for (int j = 0; j < m->njnt; j++) {
    int qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
    if (m->jnt_type[j] == JNT_FREE) {
      for (int i = 0; i < 3; i++) d->qpos[qa + i] += h * d->qvel[da + i];
      quat_integrate(d->qpos + qa + 3, d->qvel + da + 3, h);
    } else {
      d->qpos[qa] += h * d->qvel[da];
    }
  }
  memcpy(d->qacc_warmstart, d->qacc, sizeof(double) * (size_t)nv);
  FL(6 * nv + m->njnt * 4);
  d->flops += om_flops_acc; om_flops_acc = 0;
}

void om_step(const OModel* m, OData* d, int nstep) {
  for (int s = 0; s < nstep; s++) {
    om_forward(m, d);
    euler(m, d);
  }
}
