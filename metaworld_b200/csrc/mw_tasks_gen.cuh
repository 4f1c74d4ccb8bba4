// Per-task obs / reward / reset bodies (hand written; included by mw_tasks.cuh).
// Task ids must match metaworld_b200/tasks.py:TASK_IDS.
#pragma once

#include "mw_task_ids.h"

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
    case T_REACH: case T_PUSH: case T_PICK_PLACE: case T_REACH_WALL: case T_SWEEP_INTO:
      obs_body_geomquat(c, F_TASK0, F_TASK0 + 1, o);   // body "obj", geom "objGeom"  (sawyer_reach_v3.py:99-104)
      break;
    case T_PUSH_WALL: case T_PICK_PLACE_WALL: case T_PUSH_BACK:   // geom objGeom xpos + scipy quat (sawyer_push_wall_v3.py:120-126)
      obs_body_geomquat(c, F_TASK0 + 1, F_TASK0 + 1, o);
      break;
    case T_SWEEP: case T_HAND_INSERT: case T_PICK_OUT_OF_HOLE:    // body obj xpos + xquat (sawyer_sweep_v3.py:92-97)
      mw_frame_pos(c.m, c.w, F_TASK0, o); mw_frame_quat(c.m, c.w, F_TASK0, o + 3);
      break;
    case T_DOOR_OPEN: {                                // geom "handle" xpos + scipy quat (sawyer_door_v3.py:97-103)
      real R[9]; mw_frame_pos(c.m, c.w, F_TASK0, o); frame_mat(c.m, c.w, F_TASK0, R); mat2quat_scipy(R, o + 3);
    } break;
    case T_DRAWER_OPEN:                                // body drawer_link + (0,-0.16,0), xquat (sawyer_drawer_open_v3.py:93-97)
      mw_frame_pos(c.m, c.w, F_TASK0, o); o[1] -= (real)0.16; mw_frame_quat(c.m, c.w, F_TASK0, o + 3);
      break;
    case T_DRAWER_CLOSE:                               // + (0,-0.16,0.05), zeros (sawyer_drawer_close_v3.py:92-96)
      mw_frame_pos(c.m, c.w, F_TASK0, o); o[1] -= (real)0.16; o[2] += (real)0.05;
      break;
    case T_BUTTON_PRESS_TOPDOWN:                       // body button + (0,0,0.193), xquat (sawyer_button_press_topdown_v3.py:92-96)
      mw_frame_pos(c.m, c.w, F_TASK0, o); o[2] += (real)0.193; mw_frame_quat(c.m, c.w, F_TASK0, o + 3);
      break;
    case T_PEG_INSERT_SIDE: {                          // site pegGrasp pos + scipy quat of the site frame (sawyer_peg_insertion_side_v3.py:130-135)
      real R[9]; mw_frame_pos(c.m, c.w, F_TASK0, o); frame_mat(c.m, c.w, F_TASK0, R); mat2quat_scipy(R, o + 3);
    } break;
    case T_WINDOW_OPEN: case T_WINDOW_CLOSE:           // handle site, zeros quat (sawyer_window_open_v3.py:102-106)
      mw_frame_pos(c.m, c.w, F_TASK0, o);
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
    case T_REACH_WALL: {   // sawyer_reach_wall_v3.py:84-101,145-167
      real hand0[3] = {c.tc->hand_init[0], c.tc->hand_init[1], c.tc->hand_init[2]};
      real d = dist3(tcp, target);
      real in_place = tol_long_tail(d, 0, (real)0.05, dist3(hand0, target));
      *reward = 10 * in_place;
      info[INFO_SUCCESS] = d <= (real)0.05; info[INFO_IN_PLACE] = in_place; info[INFO_OBJ_TO_TARGET] = d; info[INFO_UNSCALED] = *reward;
    } break;
    case T_PUSH_WALL: case T_PICK_PLACE_WALL: {   // sawyer_push_wall_v3.py:89-118,173-236 ; sawyer_pick_place_wall_v3.py:85-115,177-252
      bool pick = c.tc->task_id == T_PICK_PLACE_WALL;
      const real* obj = obs + 4; real opened = obs[3];
      real oi[3] = {c.s->obj_init[0], c.s->obj_init[1], c.s->obj_init[2]};
      real mid[3] = {pick ? target[0] : (real)-0.05, (real)0.77, pick ? (real)0.25 : obj[2]};
      real sc[3] = {pick ? (real)1 : (real)3, 1, pick ? (real)3 : (real)1};
      real a[3], b[3];
      for (int i = 0; i < 3; i++) { a[i] = (obj[i] - mid[i]) * sc[i]; b[i] = (oi[i] - mid[i]) * sc[i]; }
      real tcp_to_obj = dist3(obj, tcp), o2t = dist3(obj, target);
      real p1 = tol_long_tail(v3norm(a), 0, (real)0.05, v3norm(b));
      real p2 = tol_long_tail(o2t, 0, (real)0.05, dist3(oi, target));
      real g = gripper_caging_reward(c, obj, (real)0.015, (real)0.05, (real)0.01, (real)0.005, 1, pick ? 0 : 1);
      real r;
      if (!pick) {
        r = 2 * g;
        if (tcp_to_obj < (real)0.02 && opened > 0) { r = 2 * g + 1 + 4 * p1; if (obj[1] > (real)0.75) r = 2 * g + 1 + 4 + 3 * p2; }
      } else {
        real ipg = hamacher(g, p1);
        r = ipg;
        if (tcp_to_obj < (real)0.02 && opened > 0 && (obj[2] - (real)0.015 > oi[2])) { r = ipg + 1 + 4 * p1; if (obj[1] > (real)0.75) r = ipg + 1 + 4 + 3 * p2; }
      }
      if (o2t < (real)0.05) r = 10;
      *reward = r;
      bool touch = touching_object(c, c.tc->main_geom, (int)c.tc->p[14], (int)c.tc->p[15]);
      info[INFO_SUCCESS] = o2t <= (real)0.07; info[INFO_NEAR_OBJECT] = tcp_to_obj <= (real)0.03;
      info[INFO_GRASP_SUCCESS] = touch && opened > 0 && (obj[2] - (real)0.02 > oi[2]);
      info[INFO_GRASP_REWARD] = g; info[INFO_IN_PLACE] = p2; info[INFO_OBJ_TO_TARGET] = o2t; info[INFO_UNSCALED] = r;
    } break;
    case T_PUSH_BACK: {   // sawyer_push_back_v3.py:69-98,256-294
      const real* obj = obs + 4; real opened = obs[3];
      real oi[3] = {c.s->obj_init[0], c.s->obj_init[1], c.s->obj_init[2]};
      real tcp_to_obj = dist3(obj, tcp), t2o = dist3(obj, target), t2oi = dist3(oi, target);
      real in_place = tol_long_tail(t2o, 0, (real)0.05, t2oi);
      real g = grip_caging(c, obj, (real)0.007, (real)0.003, (real)0.01);
      real r = hamacher(g, in_place);
      if (tcp_to_obj < (real)0.01 && opened > 0 && opened < (real)0.55 && (t2oi - t2o > (real)0.01)) r += 1 + 5 * in_place;
      if (t2o < (real)0.05) r = 10;
      *reward = r;
      bool touch = touching_object(c, c.tc->main_geom, (int)c.tc->p[14], (int)c.tc->p[15]);
      info[INFO_SUCCESS] = t2o <= (real)0.07; info[INFO_NEAR_OBJECT] = tcp_to_obj <= (real)0.03;
      info[INFO_GRASP_SUCCESS] = touch && opened > 0 && (obj[2] - (real)0.02 > oi[2]);
      info[INFO_GRASP_REWARD] = g; info[INFO_IN_PLACE] = in_place; info[INFO_OBJ_TO_TARGET] = t2o; info[INFO_UNSCALED] = r;
    } break;
    case T_SWEEP: case T_SWEEP_INTO: {   // sawyer_sweep_v3.py:68-90,228-266 ; sawyer_sweep_into_goal_v3.py:68-90,219-257
      bool into = c.tc->task_id == T_SWEEP_INTO;
      const real* obj = obs + 4; real opened = obs[3];
      real oi[3] = {c.s->obj_init[0], c.s->obj_init[1], c.s->obj_init[2]};
      real tg[3] = {target[0], target[1], into ? obj[2] : target[2]};
      real o2t = dist3(obj, tg), tcp_to_obj = dist3(obj, tcp);
      real in_place = tol_long_tail(o2t, 0, (real)0.05, dist3(oi, tg));
      real g = into ? grip_caging(c, obj, (real)0.02, (real)0.005, (real)0.01) : grip_caging(c, obj, (real)0.02, (real)0.01, (real)0.005);
      real r = 2 * g + 6 * hamacher(g, in_place);
      if (o2t < (real)0.05) r = 10;
      *reward = r;
      bool touch = touching_object(c, c.tc->main_geom, (int)c.tc->p[14], (int)c.tc->p[15]);
      info[INFO_SUCCESS] = o2t <= (real)0.05; info[INFO_NEAR_OBJECT] = tcp_to_obj <= (real)0.03; info[INFO_GRASP_SUCCESS] = touch && opened > 0;
      info[INFO_GRASP_REWARD] = g; info[INFO_IN_PLACE] = in_place; info[INFO_OBJ_TO_TARGET] = o2t; info[INFO_UNSCALED] = r;
    } break;
    case T_HAND_INSERT: {   // sawyer_hand_insert_v3.py:67-95,129-175
      const real* obj = obs + 4; real opened = obs[3];
      real oi[3] = {c.s->obj_init[0], c.s->obj_init[1], c.s->obj_init[2]};
      real t2o = dist3(obj, target), tcp_to_obj = dist3(obj, tcp);
      real in_place = tol_long_tail(t2o, 0, (real)0.05, dist3(oi, target));
      real g = gripper_caging_reward(c, obj, (real)0.015, (real)0.05, (real)0.01, (real)0.005, 1, 1);
      real r = hamacher(g, in_place);
      if (tcp_to_obj < (real)0.02 && opened > 0) r += 1 + 7 * in_place;
      if (t2o < (real)0.05) r = 10;
      *reward = r;
      bool touch = touching_object(c, c.tc->main_geom, (int)c.tc->p[14], (int)c.tc->p[15]);
      info[INFO_SUCCESS] = t2o <= (real)0.05; info[INFO_NEAR_OBJECT] = tcp_to_obj <= (real)0.03;
      info[INFO_GRASP_SUCCESS] = touch && opened > 0 && (obj[2] - (real)0.02 > oi[2]);
      info[INFO_GRASP_REWARD] = g; info[INFO_IN_PLACE] = in_place; info[INFO_OBJ_TO_TARGET] = t2o; info[INFO_UNSCALED] = r;
    } break;
    case T_PICK_OUT_OF_HOLE: {   // sawyer_pick_out_of_hole_v3.py:67-98,138-208
      const real* obj = obs + 4;
      real oi[3] = {c.s->obj_init[0], c.s->obj_init[1], c.s->obj_init[2]};
      real o2t = dist3(obj, target), tcp_to_obj = dist3(obj, tcp);
      real dx = tcp[0] - oi[0], dy = tcp[1] - oi[1], radius = sqrt(dx * dx + dy * dy);
      real floorh = radius <= (real)0.03 ? (real)0 : (real)0.015 * log(radius - (real)0.03) + (real)0.15;
      real above = tcp[2] >= floorh ? (real)1 : tol_long_tail(fmax(floorh - tcp[2], (real)0), 0, (real)0.01, (real)0.02);
      real g = gripper_caging_reward(c, obj, (real)0.015, (real)0.02, (real)0.01, (real)0.03, (real)0.1, 1);
      real in_place = tol_long_tail(o2t, 0, (real)0.02, dist3(oi, target));
      real r = hamacher(g, in_place);
      bool gs = tcp_to_obj < (real)0.04 && (obj[2] - (real)0.02 > oi[2]) && !(obs[3] < (real)0.33);
      if (gs) r += 1 + 5 * hamacher(in_place, above);
      if (o2t < (real)0.05) r = 10;
      *reward = r;
      info[INFO_SUCCESS] = o2t <= (real)0.07; info[INFO_NEAR_OBJECT] = tcp_to_obj <= (real)0.03; info[INFO_GRASP_SUCCESS] = gs;
      info[INFO_GRASP_REWARD] = g; info[INFO_IN_PLACE] = in_place; info[INFO_OBJ_TO_TARGET] = o2t; info[INFO_UNSCALED] = r;
    } break;
    case T_PUSH: {   // sawyer_push_v3.py:85-113,171-213
      const real* obj = obs + 4; real opened = obs[3];
      real oi[3] = {c.s->obj_init[0], c.s->obj_init[1], c.s->obj_init[2]};
      real tcp_to_obj = dist3(obj, tcp), t2o = dist3(obj, target);
      real in_place = tol_long_tail(t2o, 0, (real)0.05, dist3(oi, target));
      real grasped = gripper_caging_reward(c, obj, (real)0.015, (real)0.05, (real)0.01, (real)0.005, 1, 1);
      real r = 2 * grasped;
      if (tcp_to_obj < (real)0.02 && opened > 0) r += 1 + r + 5 * in_place;
      if (t2o < (real)0.05) r = 10;
      *reward = r;
      bool touch = touching_object(c, c.tc->main_geom, (int)c.tc->p[14], (int)c.tc->p[15]);
      info[INFO_SUCCESS] = t2o <= (real)0.05; info[INFO_NEAR_OBJECT] = tcp_to_obj <= (real)0.03;
      info[INFO_GRASP_SUCCESS] = touch && opened > 0 && (obj[2] - (real)0.02 > oi[2]);
      info[INFO_GRASP_REWARD] = grasped; info[INFO_IN_PLACE] = in_place; info[INFO_OBJ_TO_TARGET] = t2o; info[INFO_UNSCALED] = r;
    } break;
    case T_PICK_PLACE: {   // sawyer_pick_place_v3.py:85-120,180-293 (own caging; init pads alias the live pad positions)
      const real* obj = obs + 4; real opened = obs[3];
      real oi[3] = {c.s->obj_init[0], c.s->obj_init[1], c.s->obj_init[2]}, it[3] = {c.s->init_tcp[0], c.s->init_tcp[1], c.s->init_tcp[2]};
      real t2o = dist3(obj, target), tcp_to_obj = dist3(obj, tcp);
      real in_place = tol_long_tail(t2o, 0, (real)0.05, dist3(oi, target));
      real lp[3], rp[3]; mw_frame_pos(c.m, c.w, F_LPAD, lp); mw_frame_pos(c.m, c.w, F_RPAD, rp);
      real dl = lp[1] - obj[1], dr = obj[1] - rp[1];
      real rm = fabs(fabs(obj[1] - rp[1]) - (real)0.05), lm = fabs(fabs(obj[1] - lp[1]) - (real)0.05);
      real rc = tol_long_tail(dr, (real)0.015, (real)0.05, rm), lc = tol_long_tail(dl, (real)0.015, (real)0.05, lm);
      real ycag = hamacher(lc, rc);
      real ex = tcp[0] - obj[0], ez = tcp[2] - obj[2], ix = oi[0] - it[0], iz = oi[2] - it[2];
      real xz = tol_long_tail(sqrt(ex * ex + ez * ez), 0, (real)0.005, sqrt(ix * ix + iz * iz) - (real)0.005);
      real closed = fmin(fmax((real)0, c.action[3]), (real)1);
      real caging = hamacher(ycag, xz);
      real gripping = caging > (real)0.97 ? closed : (real)0;
      real grasped = (hamacher(caging, gripping) + caging) / 2;
      real r = hamacher(grasped, in_place);
      if (tcp_to_obj < (real)0.02 && opened > 0 && (obj[2] - (real)0.01 > oi[2])) r += 1 + 5 * in_place;
      if (t2o < (real)0.05) r = 10;
      *reward = r;
      bool touch = touching_object(c, c.tc->main_geom, (int)c.tc->p[14], (int)c.tc->p[15]);
      info[INFO_SUCCESS] = t2o <= (real)0.07; info[INFO_NEAR_OBJECT] = tcp_to_obj <= (real)0.03;
      info[INFO_GRASP_SUCCESS] = touch && opened > 0 && (obj[2] - (real)0.02 > oi[2]);
      info[INFO_GRASP_REWARD] = grasped; info[INFO_IN_PLACE] = in_place; info[INFO_OBJ_TO_TARGET] = t2o; info[INFO_UNSCALED] = r;
    } break;
    case T_DOOR_OPEN: {   // sawyer_door_v3.py:68-94,131-205
      const real* hand = obs; real door[3] = {obs[4] - (real)0.05, obs[5], obs[6]};
      real theta = c.w->qpos[9];
      real grab = (fmin(fmax(c.action[3], (real)-1), (real)1) + 1) / 2;
      real dx = hand[0] - door[0], dy = hand[1] - door[1];
      real radius = sqrt(dx * dx + dy * dy), floorh = radius <= (real)0.12 ? (real)0 : (real)0.04 * log(radius - (real)0.12) + (real)0.4;
      real above = hand[2] >= floorh ? (real)1 : tol_long_tail(floorh - hand[2], 0, (real)0.01, floorh / 2);
      real e[3] = {hand[0] - door[0] - (real)0.05, hand[1] - door[1] - (real)0.03, hand[2] - door[2] + (real)0.01};
      real in_place = tol_long_tail(v3norm(e), 0, (real)0.06, (real)0.5);
      real ready = hamacher(above, in_place);
      const real PI = (real)3.141592653589793;
      real opened = (real)0.2 * (theta < -PI / 90) + (real)0.8 * tol_long_tail(PI / 2 + PI / 6 + theta, 0, (real)0.5, PI / 3);
      real r = 2 * hamacher(ready, grab) + 8 * opened;
      bool succ = fabs(obs[4] - target[0]) <= (real)0.08;
      if (succ) r = 10;
      *reward = r;
      info[INFO_SUCCESS] = succ; info[INFO_NEAR_OBJECT] = ready; info[INFO_GRASP_SUCCESS] = grab >= (real)0.5;
      info[INFO_GRASP_REWARD] = grab; info[INFO_IN_PLACE] = opened; info[INFO_OBJ_TO_TARGET] = 0; info[INFO_UNSCALED] = r;
    } break;
    case T_DRAWER_OPEN: {   // sawyer_drawer_open_v3.py:66-91,115-158
      const real* grip = obs; const real* handle = obs + 4;
      real it[3] = {c.s->init_tcp[0], c.s->init_tcp[1], c.s->init_tcp[2]};
      real herr = dist3(handle, target);
      real opening = tol_long_tail(herr, 0, (real)0.02, (real)0.2);
      real hinit[3] = {target[0], target[1] + (real)0.2, target[2]};
      real ge[3] = {(handle[0] - grip[0]) * 3, (handle[1] - grip[1]) * 3, handle[2] - grip[2]};
      real gi[3] = {(hinit[0] - it[0]) * 3, (hinit[1] - it[1]) * 3, hinit[2] - it[2]};
      real caging = tol_long_tail(v3norm(ge), 0, (real)0.01, v3norm(gi));
      real r = 5 * (caging + opening);
      *reward = r;
      info[INFO_SUCCESS] = herr <= (real)0.03; info[INFO_NEAR_OBJECT] = dist3(handle, grip) <= (real)0.03; info[INFO_GRASP_SUCCESS] = obs[3] > 0;
      info[INFO_GRASP_REWARD] = caging; info[INFO_IN_PLACE] = opening; info[INFO_OBJ_TO_TARGET] = herr; info[INFO_UNSCALED] = r;
    } break;
    case T_DRAWER_CLOSE: {   // sawyer_drawer_close_v3.py:68-90,120-172  (TARGET_RADIUS = base class 0.05)
      const real* obj = obs + 4;
      real oi[3] = {c.s->obj_init[0], c.s->obj_init[1], c.s->obj_init[2]}, it[3] = {c.s->init_tcp[0], c.s->init_tcp[1], c.s->init_tcp[2]};
      real t2o = dist3(obj, target), t2oi = dist3(oi, target);
      real in_place = tol_long_tail(t2o, 0, (real)0.05, fabs(t2oi - (real)0.05));
      real tcp_to_obj = dist3(obj, tcp), tcp_to_obj_init = dist3(oi, it);
      real reach = tol_gaussian(tcp_to_obj, 0, (real)0.005, fabs(tcp_to_obj_init - (real)0.005));
      reach = hamacher(reach, fmin(fmax((real)0, c.action[3]), (real)1));
      real r = hamacher(reach, in_place);
      if (t2o <= (real)0.065) r = 1;
      r *= 10;
      *reward = r;
      info[INFO_SUCCESS] = t2o <= (real)0.065; info[INFO_NEAR_OBJECT] = tcp_to_obj <= (real)0.01; info[INFO_GRASP_SUCCESS] = 1;
      info[INFO_GRASP_REWARD] = reach; info[INFO_IN_PLACE] = in_place; info[INFO_OBJ_TO_TARGET] = t2o; info[INFO_UNSCALED] = r;
    } break;
    case T_BUTTON_PRESS_TOPDOWN: {   // sawyer_button_press_topdown_v3.py:62-89,122-162
      const real* obj = obs + 4;
      real it[3] = {c.s->init_tcp[0], c.s->init_tcp[1], c.s->init_tcp[2]};
      real tcp_to_obj = dist3(obj, tcp), tcp_to_obj_init = dist3(obj, it);
      real o2t = fabs(target[2] - obj[2]);
      real near_b = tol_long_tail(tcp_to_obj, 0, (real)0.01, tcp_to_obj_init);
      real pressed = tol_long_tail(o2t, 0, (real)0.005, (real)c.s->scal[0]);
      real r = 5 * hamacher(1 - obs[3], near_b);
      if (tcp_to_obj <= (real)0.03) r += 5 * pressed;
      *reward = r;
      info[INFO_SUCCESS] = o2t <= (real)0.024; info[INFO_NEAR_OBJECT] = tcp_to_obj <= (real)0.05; info[INFO_GRASP_SUCCESS] = obs[3] > 0;
      info[INFO_GRASP_REWARD] = near_b; info[INFO_IN_PLACE] = pressed; info[INFO_OBJ_TO_TARGET] = o2t; info[INFO_UNSCALED] = r;
    } break;
    case T_PEG_INSERT_SIDE: {   // sawyer_peg_insertion_side_v3.py:94-128,164-249
      const real* obj = obs + 4; real opened = obs[3];
      real head[3]; mw_frame_pos(c.m, c.w, F_TASK0 + 1, head);
      real oi[3] = {c.s->obj_init[0], c.s->obj_init[1], c.s->obj_init[2]};
      real tcp_to_obj = dist3(obj, tcp);
      real e[3] = {head[0] - target[0], 2 * (head[1] - target[1]), 2 * (head[2] - target[2])};
      real e0[3] = {(real)c.s->scal[0] - target[0], 2 * ((real)c.s->scal[1] - target[1]), 2 * ((real)c.s->scal[2] - target[2])};
      real o2t = v3norm(e);
      real in_place = tol_long_tail(o2t, 0, (real)0.07, v3norm(e0));
      real cb[2];
      for (int b = 0; b < 2; b++) {   // rect_prism_tolerance(curr=head, one=tlc, zero=brc)
        real zero[3], one[3]; mw_frame_pos(c.m, c.w, F_TASK0 + 3 + 2 * b, zero); mw_frame_pos(c.m, c.w, F_TASK0 + 4 + 2 * b, one);
        bool in = true; real prod = 1;
        for (int i = 0; i < 3; i++) {
          bool ok = one[i] >= zero[i] ? (zero[i] <= head[i] && head[i] <= one[i]) : (one[i] <= head[i] && head[i] <= zero[i]);
          in = in && ok; prod *= (head[i] - zero[i]) / (one[i] - zero[i]);
        }
        cb[b] = in ? prod : (real)1;
      }
      in_place = hamacher(in_place, hamacher(cb[1], cb[0]));
      real grasped = gripper_caging_reward(c, obj, (real)0.0075, (real)0.03, (real)0.01, (real)0.005, 1, 1);
      bool lifted = tcp_to_obj < (real)0.08 && opened > 0 && (obj[2] - (real)0.01 > oi[2]);
      if (lifted) grasped = 1;
      real r = hamacher(grasped, in_place);
      if (lifted) r += 1 + 5 * in_place;
      if (o2t <= (real)0.07) r = 10;
      *reward = r;
      info[INFO_SUCCESS] = o2t <= (real)0.07; info[INFO_NEAR_OBJECT] = tcp_to_obj <= (real)0.03;
      info[INFO_GRASP_SUCCESS] = tcp_to_obj < (real)0.02 && opened > 0 && (obj[2] - (real)0.01 > oi[2]);
      info[INFO_GRASP_REWARD] = grasped; info[INFO_IN_PLACE] = in_place; info[INFO_OBJ_TO_TARGET] = o2t; info[INFO_UNSCALED] = r;
    } break;
    case T_WINDOW_OPEN: case T_WINDOW_CLOSE: {   // sawyer_window_open_v3.py:79-100,126-170 ; sawyer_window_close_v3.py:83-104,138-179
      bool closing = c.tc->task_id == T_WINDOW_CLOSE;
      real obj[3]; mw_frame_pos(c.m, c.w, F_TASK0, obj);
      real h0[3] = {c.s->scal[0], c.s->scal[1], c.s->scal[2]}, it[3] = {c.s->init_tcp[0], c.s->init_tcp[1], c.s->init_tcp[2]};
      real t2o = fabs(obj[0] - target[0]);
      real t2oi = fabs((closing ? h0[0] : (real)c.s->obj_init[0]) - target[0]);
      real in_place = tol_long_tail(t2o, 0, (real)0.05, fabs(t2oi - (real)0.05));
      real tcp_to_obj = dist3(obj, tcp), m = fabs(dist3(h0, it) - (real)0.02);
      real reach = closing ? tol_gaussian(tcp_to_obj, 0, (real)0.02, m) : tol_long_tail(tcp_to_obj, 0, (real)0.02, m);
      real r = 10 * hamacher(reach, in_place);
      *reward = r;
      info[INFO_SUCCESS] = t2o <= (real)0.05; info[INFO_NEAR_OBJECT] = tcp_to_obj <= (real)0.05; info[INFO_GRASP_SUCCESS] = 1;
      info[INFO_GRASP_REWARD] = reach; info[INFO_IN_PLACE] = in_place; info[INFO_OBJ_TO_TARGET] = t2o; info[INFO_UNSCALED] = r;
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
    case T_PUSH: {   // sawyer_push_v3.py:135-169: z of obj/goal = the object body's current height (fix_extreme_obj_pos)
      real ob[3]; mw_frame_pos(c.m, c.w, F_TASK0, ob);
      real p[3] = {rv[0], rv[1], ob[2]};
      if (lane == 0) { c.s->target[0] = rv[3]; c.s->target[1] = rv[4]; c.s->target[2] = (float)ob[2]; for (int i = 0; i < 3; i++) c.s->obj_init[i] = (float)p[i]; }
      SYNCW();
      set_obj_xyz(c, p, lane);
    } break;
    case T_PICK_PLACE: case T_REACH_WALL: case T_PICK_PLACE_WALL: case T_PICK_OUT_OF_HOLE: {   // obj = rv[:3], goal = rv[3:6]; _set_obj_xyz
      real p[3] = {rv[0], rv[1], rv[2]};
      if (lane == 0) for (int i = 0; i < 3; i++) { c.s->target[i] = rv[3 + i]; c.s->obj_init[i] = rv[i]; }
      SYNCW();
      set_obj_xyz(c, p, lane);
    } break;
    case T_PUSH_WALL: case T_PUSH_BACK: {   // sawyer_push_wall_v3.py:135-171 ; sawyer_push_back_v3.py:120-140: z = geom objGeom's current height
      real og[3]; mw_frame_pos(c.m, c.w, F_TASK0 + 1, og);
      real p[3] = {rv[0], rv[1], og[2]};
      if (lane == 0) { c.s->target[0] = rv[3]; c.s->target[1] = rv[4]; c.s->target[2] = (float)og[2]; for (int i = 0; i < 3; i++) c.s->obj_init[i] = (float)p[i]; }
      SYNCW();
      set_obj_xyz(c, p, lane);
    } break;
    case T_SWEEP: {   // sawyer_sweep_v3.py:99-115: goal = (0.5, obj y, 0.01), obj z = 0.02
      real p[3] = {rv[0], rv[1], (real)0.02};
      if (lane == 0) { c.s->target[0] = 0.5f; c.s->target[1] = rv[1]; c.s->target[2] = 0.01f; for (int i = 0; i < 3; i++) c.s->obj_init[i] = (float)p[i]; }
      SYNCW();
      set_obj_xyz(c, p, lane);
    } break;
    case T_SWEEP_INTO: {   // sawyer_sweep_into_goal_v3.py:101-121: fixed goal (0, 0.84, 0.02); obj z = body obj's current height
      real ob[3]; mw_frame_pos(c.m, c.w, F_TASK0, ob);
      real p[3] = {rv[0], rv[1], ob[2]};
      if (lane == 0) { c.s->target[0] = 0.f; c.s->target[1] = 0.84f; c.s->target[2] = 0.02f; for (int i = 0; i < 3; i++) c.s->obj_init[i] = (float)p[i]; }
      SYNCW();
      set_obj_xyz(c, p, lane);
    } break;
    case T_HAND_INSERT: {   // sawyer_hand_insert_v3.py:108-127: obj z stays 0.05
      real p[3] = {rv[0], rv[1], (real)0.05};
      if (lane == 0) for (int i = 0; i < 3; i++) { c.s->target[i] = rv[3 + i]; c.s->obj_init[i] = (float)p[i]; }
      SYNCW();
      set_obj_xyz(c, p, lane);
    } break;
    case T_DOOR_OPEN: {   // sawyer_door_v3.py:112-129
      if (lane == 0) {
        for (int i = 0; i < 3; i++) { c.s->obj_init[i] = rv[i]; c.w->shift[i] = rv[i] - c.tc->movable_pos0[i]; }
        c.s->target[0] = rv[0] - 0.3f; c.s->target[1] = rv[1] - 0.45f; c.s->target[2] = rv[2];
        c.w->qpos[9] = 0; c.w->qvel[9] = 0;
      }
      SYNCW();
      eng_forward(c, lane);
    } break;
    case T_DRAWER_OPEN: {   // sawyer_drawer_open_v3.py:99-113 (no mj_forward: observation reads the previous kinematics)
      if (lane == 0) {
        for (int i = 0; i < 3; i++) { c.s->obj_init[i] = rv[i]; c.w->shift[i] = rv[i] - c.tc->movable_pos0[i]; }
        c.s->target[0] = rv[0]; c.s->target[1] = rv[1] - 0.36f; c.s->target[2] = rv[2] + 0.09f;
      }
      SYNCW();
    } break;
    case T_DRAWER_CLOSE: {   // sawyer_drawer_close_v3.py:104-118
      if (lane == 0) {
        for (int i = 0; i < 3; i++) c.w->shift[i] = rv[i] - c.tc->movable_pos0[i];
        c.s->target[0] = rv[0]; c.s->target[1] = rv[1] - 0.16f; c.s->target[2] = rv[2] + 0.09f;
        c.w->qpos[9] = (real)-0.15;
      }
      SYNCW();
      eng_forward(c, lane);
      if (lane == 0) { real o[3]; mw_frame_pos(c.m, c.w, F_TASK0, o); c.s->obj_init[0] = (float)o[0]; c.s->obj_init[1] = (float)(o[1] - (real)0.16); c.s->obj_init[2] = (float)(o[2] + (real)0.05); }
      SYNCW();
    } break;
    case T_BUTTON_PRESS_TOPDOWN: {   // sawyer_button_press_topdown_v3.py:105-120
      if (lane == 0) for (int i = 0; i < 3; i++) { c.s->obj_init[i] = rv[i]; c.w->shift[i] = rv[i] - c.tc->movable_pos0[i]; }
      SYNCW();
      eng_forward(c, lane);
      if (lane == 0) {
        real hole[3], bs[3]; mw_frame_pos(c.m, c.w, F_TASK0 + 1, hole); mw_frame_pos(c.m, c.w, F_TASK0 + 2, bs);
        for (int i = 0; i < 3; i++) c.s->target[i] = (float)hole[i];
        c.s->scal[0] = (float)fabs(hole[2] - bs[2]);
      }
      SYNCW();
    } break;
    case T_PEG_INSERT_SIDE: {   // sawyer_peg_insertion_side_v3.py:137-162
      real p[3] = {rv[0], rv[1], rv[2]};
      if (lane == 0) {
        real head[3]; mw_frame_pos(c.m, c.w, F_TASK0 + 1, head);
        for (int i = 0; i < 3; i++) { c.s->obj_init[i] = rv[i]; c.s->scal[i] = (float)head[i]; }
      }
      SYNCW();
      set_obj_xyz(c, p, lane);
      if (lane == 0) {
        for (int i = 0; i < 3; i++) c.w->shift[i] = rv[3 + i] - c.tc->movable_pos0[i];
        c.s->target[0] = rv[3] + 0.03f; c.s->target[1] = rv[4]; c.s->target[2] = rv[5] + 0.13f;
      }
      SYNCW();
    } break;
    case T_WINDOW_OPEN: case T_WINDOW_CLOSE: {   // sawyer_window_open_v3.py:109-124 ; sawyer_window_close_v3.py:113-131
      bool closing = c.tc->task_id == T_WINDOW_CLOSE;
      if (lane == 0) {
        for (int i = 0; i < 3; i++) { c.s->obj_init[i] = rv[i]; c.s->target[i] = rv[i]; c.w->shift[i] = rv[i] - c.tc->movable_pos0[i]; }
        if (!closing) c.s->target[0] = rv[0] + 0.2f;
        real h[3]; mw_frame_pos(c.m, c.w, F_TASK0, h);          // stale kinematics, as in the reference
        c.s->scal[0] = (float)(h[0] + (closing ? (real)0.2 : (real)0)); c.s->scal[1] = (float)h[1]; c.s->scal[2] = (float)h[2];
        c.w->qpos[9] = closing ? (real)0.2 : (real)0;            // data.joint("window_slide").qpos = ..., no mj_forward
      }
      SYNCW();
    } break;
    default: break;
  }
}
