/* oracle/mjphys.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU (float64, scalar, single-environment) restatement of the physics the
 * reference reaches through `mujoco==3.3.0` (pyproject.toml:28) on its hot
 * path: `mj_step` x5 per env step (metaworld/sawyer_xyz_env.py:595 via
 * MujocoEnv.do_simulation), `mj_forward` (sawyer_xyz_env.py:620), and
 * `mj_resetData` (MujocoEnv.reset reached from sawyer_xyz_env.py:678).
 *
 * PARITY UNPINNED: MuJoCo is a third-party dependency that is not vendored
 * under /root/reference and is not installed in this image, and the reference's
 * tests hold no numeric golden vectors for this path.
 * The pipeline below restates MuJoCo's published computation model
 * (kinematics -> inertia -> collision -> constraint rows with solref/solimp
 * impedance -> convex primal solve with elliptic cones -> semi-implicit Euler
 * with implicit joint damping) for exactly the MJCF feature set Meta-World uses.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 */
#ifndef MJPHYS_H
#define MJPHYS_H
#ifdef __cplusplus
extern "C" {
#endif

#define OM_MAXCON 128
#define OM_MAXEFC 512

typedef struct OModel OModel;
typedef struct OData OData;

typedef struct {
  double dist, pos[3], frame[9]; /* frame rows: normal, tangent1, tangent2 */
  double includemargin, friction[5], solref[2], solimp[5], mu;
  int dim, geom1, geom2, efc_address;
} OContact;

OModel* om_model_new(void);
void om_model_free(OModel*);
/* copies n values into the named model field (allocating it) */
int om_model_set_f64(OModel*, const char* name, const double* v, int n);
int om_model_set_i32(OModel*, const char* name, const int* v, int n);
/* add a convex mesh (hull vertices in the mesh frame) ; returns mesh id */
int om_model_add_mesh(OModel*, const double* vert, int nvert);
/* derive sizes + static collision candidate pairs; call once after all set_* */
int om_model_finalize(OModel*);
double* om_model_f64(OModel*, const char* name, int* n); /* mutable view (body_pos, site_pos, eq_data ...) */
int* om_model_i32(OModel*, const char* name, int* n);

OData* om_data_new(const OModel*);
void om_data_free(OData*);
void om_reset_data(const OModel*, OData*);               /* mj_resetData */
void om_forward(const OModel*, OData*);                  /* mj_forward */
void om_step(const OModel*, OData*, int nstep);          /* mj_step x nstep */
double* om_data_f64(OData*, const char* name, int* n);   /* qpos qvel ctrl mocap_pos mocap_quat xpos xquat xmat geom_xpos geom_xmat site_xpos site_xmat qacc efc_force ... */
int om_data_ncon(const OData*);
int om_data_nefc(const OData*);
const OContact* om_data_contacts(const OData*);
int om_data_solver_iter(const OData*);
long om_data_flops(const OData*);

#ifdef __cplusplus
}
#endif
#endif
