// Per-task obs / reward / reset bodies (hand written; included by mw_tasks.cuh).
// Task ids must match metaworld_b200/tasks.py:TASK_IDS.
#pragma once

enum {
  T_REACH = 0, T_PUSH, T_PICK_PLACE, T_DOOR_OPEN, T_DRAWER_OPEN, T_DRAWER_CLOSE, T_BUTTON_PRESS_TOPDOWN,
  T_PEG_INSERT_SIDE, T_WINDOW_OPEN, T_WINDOW_CLOSE, T_NTASK
};

// physics callbacks used by reset code (defined in mw_engine.cu)
__device__ void eng_forward(const TaskCtx& c, int lane);
__device__ void eng_sim(const TaskCtx& c, int nstep, int lane);

// ---------------------------------------------------------------- helpers shared by task families
// `_set_obj_xyz` (sawyer_xyz_env.py:351-361): qpos[9:12] = pos, qvel[9:15] = 0, set_state -> mj_forward
DEV void set_obj_xyz(const TaskCtx& c, const real* pos, int lane) {
  if (lane == 0) { for (int i = 0; i < 3; i++) c.w->qpos[9 + i] = pos[i]; for (int i = 9; i < 15; i++) c.w->qvel[i] = 0; }
  SYNCW();
  eng_forward(c, lane);
}
// free object observed as body pos + scipy quat of a geom frame (reach/push/pick-place family)
DEV void obs_body_geomquat(const TaskCtx& c, int fbody, int fgeom, real* o) {
  real R[9];
  mw_frame_pos(c.m, c.w, fbody, o);
  frame_mat(c.m, c.w, fgeom, R);
  mat2quat_scipy(R, o + 3);
}

// ---------------------------------------------------------------- observation: object slots (14 floats, zero padded)
DEV void task_obs_objects(const TaskCtx& c, real* o) {
  for (int i = 0; i < 14; i++) o[i] = 0;
  switch (c.tc->task_id) {
    case T_REACH:
      obs_body_geomquat(c, F_TASK0, F_TASK0 + 1, o);   // body "obj", geom "objGeom"  (sawyer_reach_v3.py:99-104)
      break;
    default: break;
  }
}

// ---------------------------------------------------------------- reward + info (v2)
DEV void task_reward(const TaskCtx& c, const real* obs, real* reward, real* info) {
  real tcp[3]; tcp_center(c, tcp);
  real target[3] = {c.s->target[0], c.s->target[1], c.s->target[2]};
  for (int i = 0; i < INFO_N; i++) info[i] = 0;
  switch (c.tc->task_id) {
    case T_REACH: {   // sawyer_reach_v3.py:82-98,140-162
      real hand0[3] = {c.tc->hand_init[0], c.tc->hand_init[1], c.tc->hand_init[2]};
      real d = dist3(tcp, target);
      real in_place = tol_long_tail(d, 0, (real)0.05, dist3(hand0, target));
      *reward = 10 * in_place;
      info[INFO_SUCCESS] = d <= (real)0.05; info[INFO_NEAR_OBJECT] = d; info[INFO_GRASP_SUCCESS] = 1;
      info[INFO_GRASP_REWARD] = d; info[INFO_IN_PLACE] = in_place; info[INFO_OBJ_TO_TARGET] = d; info[INFO_UNSCALED] = *reward;
    } break;
    default: *reward = 0; break;
  }
  (void)obs;
}

// ---------------------------------------------------------------- reset_model (after _reset_hand); all lanes call
// rv = the task's frozen rand_vec (Task.data['rand_vec']); writes env constants into c.s (lane 0)
DEV void task_reset_model(const TaskCtx& c, const float* rv, int lane) {
  switch (c.tc->task_id) {
    case T_REACH: {   // sawyer_reach_v3.py:119-138
      real p[3] = {rv[0], rv[1], rv[2]};
      if (lane == 0) for (int i = 0; i < 3; i++) { c.s->target[i] = rv[3 + i]; c.s->obj_init[i] = rv[i]; }
      SYNCW();
      set_obj_xyz(c, p, lane);
    } break;
    default: break;
  }
}
