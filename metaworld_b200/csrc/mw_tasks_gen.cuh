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
  if (lane == 0) { for (int i = 0; i < 3; i++) QSET(c.w, 9 + i, pos[i]); for (int i = 9; i < 15; i++) c.w->qvel[i] = 0; }
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
    case T_REACH: case T_PUSH: case T_PICK_PLACE: case T_REACH_WALL: case T_SWEEP_INTO: case T_SHELF_PLACE:
      obs_body_geomquat(c, F_TASK0, F_TASK0 + 1, o);   // body "obj", geom "objGeom"  (sawyer_reach_v3.py:99-104)
      break;
    case T_PUSH_WALL: case T_PICK_PLACE_WALL: case T_PUSH_BACK:   // geom objGeom xpos + scipy quat (sawyer_push_wall_v3.py:120-126)
      obs_body_geomquat(c, F_TASK0 + 1, F_TASK0 + 1, o);
      break;
    case T_ASSEMBLY: case T_DISASSEMBLE:               // site RoundNut-8 pos + body RoundNut xquat (sawyer_assembly_peg_v3.py:103-108)
      mw_frame_pos(c.m, c.w, F_TASK0, o); mw_frame_quat(c.m, c.w, F_TASK0 + 1, o + 3);
      break;
    case T_HAMMER:                                     // hammer + nail_link, pos and xquat each (sawyer_hammer_v3.py:90-99)
      mw_frame_pos(c.m, c.w, F_TASK0, o); mw_frame_quat(c.m, c.w, F_TASK0, o + 3);
      mw_frame_pos(c.m, c.w, F_TASK0 + 1, o + 7); mw_frame_quat(c.m, c.w, F_TASK0 + 1, o + 10);
      break;
    case T_FAUCET_OPEN: case T_FAUCET_CLOSE:           // handle site + (0,0,-0.01), body faucetBase xquat (sawyer_faucet_open_v3.py:96-101)
      mw_frame_pos(c.m, c.w, F_TASK0, o); o[2] -= (real)0.01; mw_frame_quat(c.m, c.w, F_TASK0 + 1, o + 3);
      break;
    case T_SWEEP: case T_HAND_INSERT: case T_PICK_OUT_OF_HOLE: case T_BASKETBALL: case T_BIN_PICKING: case T_BOX_CLOSE:
      // body xpos + xquat (sawyer_sweep_v3.py:92-97 ; body "bsktball" / "obj" / "top_link")
      mw_frame_pos(c.m, c.w, F_TASK0, o); mw_frame_quat(c.m, c.w, F_TASK0, o + 3);
      break;
    case T_DOOR_OPEN: case T_DOOR_CLOSE: {             // geom "handle" xpos + scipy quat (sawyer_door_v3.py:97-103)
      real R[9]; mw_frame_pos(c.m, c.w, F_TASK0, o); frame_mat(c.m, c.w, F_TASK0, R); mat2quat_scipy(R, o + 3);
    } break;
    case T_DRAWER_OPEN:                                // body drawer_link + (0,-0.16,0), xquat (sawyer_drawer_open_v3.py:93-97)
      mw_frame_pos(c.m, c.w, F_TASK0, o); o[1] -= (real)0.16; mw_frame_quat(c.m, c.w, F_TASK0, o + 3);
      break;
    case T_DRAWER_CLOSE:                               // + (0,-0.16,0.05), zeros (sawyer_drawer_close_v3.py:92-96)
      mw_frame_pos(c.m, c.w, F_TASK0, o); o[1] -= (real)0.16; o[2] += (real)0.05;
      break;
    case T_BUTTON_PRESS_TOPDOWN: case T_BUTTON_PRESS_TOPDOWN_WALL:   // body button + (0,0,0.193), xquat (sawyer_button_press_topdown_v3.py:92-96)
      mw_frame_pos(c.m, c.w, F_TASK0, o); o[2] += (real)0.193; mw_frame_quat(c.m, c.w, F_TASK0, o + 3);
      break;
    case T_BUTTON_PRESS: case T_BUTTON_PRESS_WALL:     // body button + (0,-0.193,0), xquat (sawyer_button_press_v3.py:91-95)
      mw_frame_pos(c.m, c.w, F_TASK0, o); o[1] -= (real)0.193; mw_frame_quat(c.m, c.w, F_TASK0, o + 3);
      break;
    case T_COFFEE_BUTTON:                              // site buttonStart, constant quat (sawyer_coffee_button_v3.py:100-104)
      mw_frame_pos(c.m, c.w, F_TASK0, o); o[3] = 1;
      break;
    case T_COFFEE_PULL: case T_COFFEE_PUSH:            // body obj + scipy quat of mesh geom "mug" (sawyer_coffee_pull_v3.py:102-108)
      obs_body_geomquat(c, F_TASK0, F_TASK0 + 1, o);
      break;
    case T_DIAL_TURN: {                                // body dial + 0.05 (sin th, -cos th, 0), xquat (sawyer_dial_turn_v3.py:87-100)
      real th = c.w->qpos[9], sn, cs; sincos(th, &sn, &cs);
      mw_frame_pos(c.m, c.w, F_TASK0, o); o[0] += (real)0.05 * sn; o[1] -= (real)0.05 * cs; mw_frame_quat(c.m, c.w, F_TASK0, o + 3);
    } break;
    case T_DOOR_LOCK: case T_DOOR_UNLOCK:              // lock site pos + body door_link xquat (sawyer_door_lock_v3.py:101-105)
      mw_frame_pos(c.m, c.w, F_TASK0, o); mw_frame_quat(c.m, c.w, F_TASK0 + 1, o + 3);
      break;
    case T_PEG_INSERT_SIDE: {                          // site pegGrasp pos + scipy quat of the site frame (sawyer_peg_insertion_side_v3.py:130-135)
      real R[9]; mw_frame_pos(c.m, c.w, F_TASK0, o); frame_mat(c.m, c.w, F_TASK0, R); mat2quat_scipy(R, o + 3);
    } break;
    case T_WINDOW_OPEN: case T_WINDOW_CLOSE:           // handle site, zeros quat (sawyer_window_open_v3.py:102-106)
      mw_frame_pos(c.m, c.w, F_TASK0, o);
      break;
    case T_HANDLE_PRESS: case T_HANDLE_PRESS_SIDE: case T_HANDLE_PULL: case T_HANDLE_PULL_SIDE:   // handle site, zeros quat (sawyer_handle_press_v3.py:101-105)
      mw_frame_pos(c.m, c.w, F_TASK0, o);
      break;
    case T_LEVER_PULL:                                 // site leverStart + scipy quat of geom objGeom (sawyer_lever_pull_v3.py:106-112)
      obs_body_geomquat(c, F_TASK0, F_TASK0 + 1, o);
      break;
    case T_PEG_UNPLUG_SIDE:                            // site pegEnd + body plug1 xquat (sawyer_peg_unplug_side_v3.py:90-94)
      mw_frame_pos(c.m, c.w, F_TASK0, o); mw_frame_quat(c.m, c.w, F_TASK0 + 1, o + 3);
      break;
    case T_PLATE_SLIDE: case T_PLATE_SLIDE_SIDE: case T_PLATE_SLIDE_BACK: case T_PLATE_SLIDE_BACK_SIDE: case T_SOCCER:
      // geom puck xpos + scipy quat (sawyer_plate_slide_v3.py:97-102) ; body soccer_ball pos + scipy quat of its xmat (sawyer_soccer_v3.py:99-104)
      obs_body_geomquat(c, F_TASK0, F_TASK0, o);
      break;
    case T_STICK_PUSH: case T_STICK_PULL: {            // stick pos + scipy quat, insertion site (+0.09 y when pushing), zeros (sawyer_stick_push_v3.py:100-118)
      obs_body_geomquat(c, F_TASK0, F_TASK0, o);
      mw_frame_pos(c.m, c.w, F_TASK0 + 1, o + 7);
      if (c.tc->task_id == T_STICK_PUSH) o[8] += (real)0.09;
    } break;
    default: break;
  }
}

// state that the reference keeps "live" through numpy views and that must be refreshed before obs / reward
DEV void task_live_update(const TaskCtx& c) {
  if (c.tc->task_id == T_BASKETBALL) {   // _target_pos aliases data.site("goal").xpos = hoop frame + (overwritten) local offset
    real b[3]; mw_frame_pos(c.m, c.w, F_TASK0 + 1, b);
    for (int i = 0; i < 3; i++) c.s->target[i] = (float)(b[i] + (real)c.s->scal[3 + i]);
  }
}

// Does this task's evaluate_state read contact forces (touching_object, sawyer_xyz_env.py:392-440)?  Only then does the step
// need the constraint solve of the final mj_forward (sawyer_xyz_env.py:620); every other task reads poses only, and
// mj_forward does not feed qacc back into the state (qacc_warmstart is written by mj_step's integrator alone [3P]), so for
// them the 6th pass reduces to the kinematics with bit-identical results.  Keep in sync with the cases of task_reward below.
DEV bool task_needs_contact_forces(int task_id) {
  switch (task_id) {
    case T_PUSH_WALL: case T_PICK_PLACE_WALL: case T_PUSH_BACK: case T_SWEEP: case T_SWEEP_INTO: case T_HAND_INSERT: case T_PUSH:
    case T_PICK_PLACE: case T_COFFEE_PULL: case T_COFFEE_PUSH: case T_SHELF_PLACE: case T_SOCCER: case T_STICK_PULL: case T_STICK_PUSH:
      return true;
    default: return false;
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
    case T_BUTTON_PRESS: case T_BUTTON_PRESS_WALL: case T_BUTTON_PRESS_TOPDOWN_WALL: case T_COFFEE_BUTTON: {
      // sawyer_button_press_v3.py:126-166 ; _wall_v3.py:130-173 ; _topdown_wall_v3.py:127-167 ; sawyer_coffee_button_v3.py:133-173
      const int id = c.tc->task_id;
      const real* obj = obs + 4;
      real it[3] = {c.s->init_tcp[0], c.s->init_tcp[1], c.s->init_tcp[2]};
      real tcp_to_obj = dist3(obj, tcp), tcp_to_obj_init = dist3(obj, it);
      const int ax = id == T_BUTTON_PRESS_TOPDOWN_WALL ? 2 : 1;
      real o2t = fabs(target[ax] - obj[ax]);
      real nb = (id == T_BUTTON_PRESS || id == T_COFFEE_BUTTON) ? (real)0.05 : (real)0.01;
      real near_b = tol_long_tail(tcp_to_obj, 0, nb, tcp_to_obj_init);
      real pressed = tol_long_tail(o2t, 0, (real)0.005, id == T_COFFEE_BUTTON ? (real)0.03 : (real)c.s->scal[0]);
      real closed = fmax(obs[3], (real)0), r;
      if (id == T_BUTTON_PRESS_WALL) {
        if (tcp_to_obj > (real)0.07) r = 2 * hamacher((1 - obs[3]) / 2, near_b);
        else r = 2 + 2 * (1 + obs[3]) + 4 * pressed * pressed;
      } else if (id == T_BUTTON_PRESS_TOPDOWN_WALL) {
        r = 5 * hamacher(closed, near_b);
        if (tcp_to_obj <= (real)0.03) r += 5 * pressed;
      } else {
        r = 2 * hamacher(closed, near_b);
        if (tcp_to_obj <= (real)0.05) r += 8 * pressed;
      }
      *reward = r;
      real thr = id == T_BUTTON_PRESS_WALL ? (real)0.03 : (id == T_BUTTON_PRESS_TOPDOWN_WALL ? (real)0.024 : (real)0.02);
      info[INFO_SUCCESS] = o2t <= thr; info[INFO_NEAR_OBJECT] = tcp_to_obj <= (real)0.05; info[INFO_GRASP_SUCCESS] = obs[3] > 0;
      info[INFO_GRASP_REWARD] = near_b; info[INFO_IN_PLACE] = pressed; info[INFO_OBJ_TO_TARGET] = o2t; info[INFO_UNSCALED] = r;
    } break;
    case T_COFFEE_PULL: case T_COFFEE_PUSH: {   // sawyer_coffee_pull_v3.py:66-97,139-189
      const real* obj = obs + 4; real opened = obs[3];
      real oi[3] = {c.s->obj_init[0], c.s->obj_init[1], c.s->obj_init[2]};
      real a[3] = {2 * (obj[0] - target[0]), 2 * (obj[1] - target[1]), obj[2] - target[2]};
      real b[3] = {2 * (oi[0] - target[0]), 2 * (oi[1] - target[1]), oi[2] - target[2]};
      real t2o = v3norm(a);
      real in_place = tol_long_tail(t2o, 0, (real)0.05, v3norm(b));
      real tcp_to_obj = dist3(obj, tcp);
      real g = gripper_caging_reward(c, obj, (real)0.02, (real)0.05, (real)0.04, (real)0.05, (real)0.7, 2);
      real r = hamacher(g, in_place);
      if (tcp_to_obj < (real)0.04 && opened > 0) r += 1 + 5 * in_place;
      if (t2o < (real)0.05) r = 10;
      *reward = r;
      real plain = dist3(obj, target);
      bool touch = touching_object(c, c.tc->main_geom, (int)c.tc->p[14], (int)c.tc->p[15]);
      info[INFO_SUCCESS] = plain <= (real)0.07; info[INFO_NEAR_OBJECT] = tcp_to_obj <= (real)0.03; info[INFO_GRASP_SUCCESS] = touch && opened > 0;
      info[INFO_GRASP_REWARD] = g; info[INFO_IN_PLACE] = in_place; info[INFO_OBJ_TO_TARGET] = plain; info[INFO_UNSCALED] = r;
    } break;
    case T_DIAL_TURN: {   // sawyer_dial_turn_v3.py:63-85,122-171
      real ob[14]; task_obs_objects(c, ob);
      real push[3] = {ob[0] + (real)0.05, ob[1] + (real)0.02, ob[2] + (real)0.09};
      real p0[3] = {c.s->scal[0], c.s->scal[1], c.s->scal[2]}, it[3] = {c.s->init_tcp[0], c.s->init_tcp[1], c.s->init_tcp[2]};
      real t2o = dist3(ob, target);
      real in_place = tol_long_tail(t2o, 0, (real)0.07, fabs(dist3(p0, target) - (real)0.07));
      real tcp_to_obj = dist3(push, tcp);
      real reach = tol_gaussian(tcp_to_obj, 0, (real)0.005, fabs(dist3(p0, it) - (real)0.005));
      reach = hamacher(reach, fmin(fmax((real)0, c.action[3]), (real)1));
      real r = 10 * hamacher(reach, in_place);
      *reward = r;
      info[INFO_SUCCESS] = t2o <= (real)0.07; info[INFO_NEAR_OBJECT] = tcp_to_obj <= (real)0.01; info[INFO_GRASP_SUCCESS] = 1;
      info[INFO_GRASP_REWARD] = reach; info[INFO_IN_PLACE] = in_place; info[INFO_OBJ_TO_TARGET] = t2o; info[INFO_UNSCALED] = r;
    } break;
    case T_DOOR_CLOSE: {   // sawyer_door_close_v3.py:105-157
      const real* obj = obs + 4;
      real oi[3] = {c.s->obj_init[0], c.s->obj_init[1], c.s->obj_init[2]}, h0[3] = {c.tc->hand_init[0], c.tc->hand_init[1], c.tc->hand_init[2]};
      real tcp_to_target = dist3(tcp, target), o2t = dist3(obj, target);
      real in_place = tol_gaussian(o2t, 0, (real)0.05, dist3(oi, target));
      real hand_in_place = tol_gaussian(tcp_to_target, 0, (real)0.0125, dist3(h0, obj) + (real)0.1);
      real r = 3 * hand_in_place + 6 * in_place;
      if (o2t < (real)0.05) r = 10;
      *reward = r;
      info[INFO_SUCCESS] = o2t <= (real)0.08; info[INFO_NEAR_OBJECT] = 0; info[INFO_GRASP_SUCCESS] = 1;
      info[INFO_GRASP_REWARD] = 1; info[INFO_IN_PLACE] = hand_in_place; info[INFO_OBJ_TO_TARGET] = o2t; info[INFO_UNSCALED] = r;
    } break;
    case T_DOOR_LOCK: {   // sawyer_door_lock_v3.py:65-99,123-157 (tcp := left pad; init_left_pad aliases the live pad position)
      const real* obj = obs + 4;
      real lp[3]; mw_frame_pos(c.m, c.w, F_LPAD, lp);
      real e[3] = {(obj[0] - lp[0]) * (real)0.25, obj[1] - lp[1], (obj[2] - lp[2]) * (real)0.5};
      real tcp_to_obj = v3norm(e);
      real o2t = fabs(target[2] - obj[2]);
      real near_l = tol_long_tail(tcp_to_obj, 0, (real)0.01, tcp_to_obj);
      real pressed = tol_long_tail(o2t, 0, (real)0.005, (real)0.1);
      real r = 2 * hamacher(fmax(obs[3], (real)0), near_l) + 8 * pressed;
      *reward = r;
      info[INFO_SUCCESS] = o2t <= (real)0.02; info[INFO_NEAR_OBJECT] = tcp_to_obj <= (real)0.05; info[INFO_GRASP_SUCCESS] = obs[3] > 0;
      info[INFO_GRASP_REWARD] = near_l; info[INFO_IN_PLACE] = pressed; info[INFO_OBJ_TO_TARGET] = o2t; info[INFO_UNSCALED] = r;
    } break;
    case T_DOOR_UNLOCK: {   // sawyer_door_unlock_v3.py:63-97,126-171 (obj_init_pos aliases the live lock_link position)
      const real* grip = obs; const real* lock = obs + 4;
      real ll[3]; mw_frame_pos(c.m, c.w, F_TASK0 + 2, ll);
      real it[3] = {c.s->init_tcp[0], c.s->init_tcp[1], c.s->init_tcp[2]};
      real a[3] = {(grip[0] - lock[0]) * (real)0.25, grip[1] + (real)0.055 - lock[1], (grip[2] + (real)0.07 - lock[2]) * (real)0.5};
      real b[3] = {(it[0] - ll[0]) * (real)0.25, it[1] + (real)0.055 - ll[1], (it[2] + (real)0.07 - ll[2]) * (real)0.5};
      real s2l = v3norm(a);
      real ready = tol_long_tail(s2l, 0, (real)0.02, v3norm(b));
      real o2t = fabs(target[0] - lock[0]);
      real pushed = tol_long_tail(o2t, 0, (real)0.005, (real)0.1);
      real r = 2 * ready + 8 * pushed;
      *reward = r;
      info[INFO_SUCCESS] = o2t <= (real)0.02; info[INFO_NEAR_OBJECT] = s2l <= (real)0.05; info[INFO_GRASP_SUCCESS] = obs[3] > 0;
      info[INFO_GRASP_REWARD] = ready; info[INFO_IN_PLACE] = pushed; info[INFO_OBJ_TO_TARGET] = o2t; info[INFO_UNSCALED] = r;
    } break;
    case T_ASSEMBLY: case T_DISASSEMBLE: case T_HAMMER: {
      // sawyer_assembly_peg_v3.py:70-101,147-262 ; sawyer_disassemble_peg_v3.py:72-103,161-225 ; sawyer_hammer_v3.py:65-88,146-215
      const int id = c.tc->task_id;
      const real* hand = obs; const real* obj = obs + 4;
      real th[3] = {obj[0], obj[1], obj[2]};
      if (fabs(obj[0] - hand[0]) < (id == T_HAMMER ? (real)0.07 : (real)0.01)) th[0] = hand[0];
      real ideal[4] = {(real)0.707, 0, 0, (real)0.707};
      if (id == T_HAMMER) { ideal[0] = 1; ideal[3] = 0; }
      real qe = 0; for (int i = 0; i < 4; i++) { real d = obs[7 + i] - ideal[i]; qe += d * d; }
      real rq = fmax(1 - sqrt(qe) / (real)0.4, (real)0);
      real grab = gripper_caging_reward(c, th, (real)0.015, (real)0.02, (real)0.01, (real)0.01, 1, id == T_ASSEMBLY ? 2 : 1);
      real in_place; bool succ;
      if (id == T_HAMMER) {
        real head[3] = {obj[0] + (real)0.16, obj[1] + (real)0.06, obj[2]};
        in_place = (real)0.1 * (head[2] > (real)0.02) + (real)0.9 * tol_long_tail(dist3(target, head), 0, (real)0.02, (real)0.2);
        succ = c.w->qpos[16] > (real)0.09;                                   // joint NailSlideJoint
      } else {
        real wc[3]; mw_frame_pos(c.m, c.w, F_TASK0 + 2, wc);               // site "RoundNut"
        if (id == T_ASSEMBLY) {
          real pe[3] = {target[0] - wc[0], target[1] - wc[1], target[2] - wc[2]};
          real radius = sqrt(pe[0] * pe[0] + pe[1] * pe[1]);
          succ = radius < (real)0.02 && pe[2] > 0;
          real thr = succ ? (real)0.02 : (real)0.01;
          real theight = radius > thr ? (real)0.02 * log(radius - thr) + (real)0.2 : (real)0;
          pe[2] = (theight - wc[2]) * 3;
          bool lifted = wc[2] > (real)0.02 || radius < thr;
          in_place = (real)0.1 * lifted + (real)0.9 * tol_long_tail(v3norm(pe), 0, (real)0.02, (real)0.4);
        } else {
          real pe[3] = {target[0] - wc[0], target[1] - wc[1], target[2] + (real)0.1 - wc[2]};
          in_place = (real)0.1 * (wc[2] > (real)0.02) + (real)0.9 * tol_long_tail(v3norm(pe), 0, (real)0.02, (real)0.2);
          succ = obs[6] > target[2];
        }
      }
      real r = (2 * grab + 6 * in_place) * rq;
      if (id == T_HAMMER) { if (succ && r > 5) r = 10; } else if (succ) r = 10;
      *reward = r;
      info[INFO_SUCCESS] = succ; info[INFO_NEAR_OBJECT] = rq; info[INFO_GRASP_SUCCESS] = grab >= (real)0.5;
      info[INFO_GRASP_REWARD] = grab; info[INFO_IN_PLACE] = in_place; info[INFO_OBJ_TO_TARGET] = 0; info[INFO_UNSCALED] = r;
    } break;
    case T_BASKETBALL: {   // sawyer_basketball_v3.py:72-101,135-194
      const real* obj = obs + 4; real opened = obs[3];
      real oi[3] = {c.s->obj_init[0], c.s->obj_init[1], c.s->obj_init[2]};
      real a[3] = {obj[0] - target[0], obj[1] - target[1], 2 * (obj[2] - (real)0.3)};
      real b[3] = {oi[0] - target[0], oi[1] - target[1], 2 * (oi[2] - (real)0.3)};
      real t2o = v3norm(a);
      real in_place = tol_long_tail(t2o, 0, (real)0.08, v3norm(b));
      real tcp_to_obj = dist3(obj, tcp);
      real g = gripper_caging_reward(c, obj, (real)0.025, (real)0.06, (real)0.01, (real)0.005, 1, 1);
      bool lifted = tcp_to_obj < (real)0.035 && opened > 0 && (obj[2] - (real)0.01 > oi[2]);
      if (lifted) g = 1;
      real r = hamacher(g, in_place);
      if (lifted) r += 1 + 5 * in_place;
      if (t2o < (real)0.08) r = 10;
      *reward = r;
      info[INFO_SUCCESS] = t2o <= (real)0.08; info[INFO_NEAR_OBJECT] = tcp_to_obj <= (real)0.05;
      info[INFO_GRASP_SUCCESS] = opened > 0 && (obj[2] - (real)0.03 > oi[2]);
      info[INFO_GRASP_REWARD] = g; info[INFO_IN_PLACE] = in_place; info[INFO_OBJ_TO_TARGET] = t2o; info[INFO_UNSCALED] = r;
    } break;
    case T_BIN_PICKING: {   // sawyer_bin_picking_v3.py:94-129,161-243 (in-place margin is latched on the first reward after reset)
      const real* hand = obs; const real* obj = obs + 4;
      real oi[3] = {c.s->obj_init[0], c.s->obj_init[1], c.s->obj_init[2]};
      real t2o = dist3(obj, target);
      if (c.s->scal[0] < 0.f) c.s->scal[0] = (float)t2o;
      real in_place = tol_long_tail(t2o, 0, (real)0.05, (real)c.s->scal[0]);
      real r1 = sqrt((hand[0] - oi[0]) * (hand[0] - oi[0]) + (hand[1] - oi[1]) * (hand[1] - oi[1]));
      real r2 = sqrt((hand[0] - target[0]) * (hand[0] - target[0]) + (hand[1] - target[1]) * (hand[1] - target[1]));
      real f1 = r1 > (real)0.03 ? (real)0.02 * log(r1 - (real)0.03) + (real)0.2 : (real)0;
      real f2 = r2 > (real)0.03 ? (real)0.02 * log(r2 - (real)0.03) + (real)0.2 : (real)0;
      real floorh = fmin(f1, f2);
      real above = hand[2] >= floorh ? (real)1 : tol_long_tail(fmax(floorh - hand[2], (real)0), 0, (real)0.01, (real)0.05);
      real g = gripper_caging_reward(c, obj, (real)0.015, (real)0.05, (real)0.01, (real)0.01, (real)0.7, 1);
      real r = hamacher(g, in_place);
      bool nearo = dist3(obj, hand) < (real)0.04;
      bool gs = nearo && (obj[2] - (real)0.02 > oi[2]) && !(obs[3] < (real)0.43);
      if (gs) r += 1 + 5 * hamacher(above, in_place);
      if (t2o < (real)0.05) r = 10;
      *reward = r;
      info[INFO_SUCCESS] = t2o <= (real)0.05; info[INFO_NEAR_OBJECT] = nearo; info[INFO_GRASP_SUCCESS] = gs;
      info[INFO_GRASP_REWARD] = g; info[INFO_IN_PLACE] = in_place; info[INFO_OBJ_TO_TARGET] = t2o; info[INFO_UNSCALED] = r;
    } break;
    case T_BOX_CLOSE: {   // sawyer_box_close_v3.py:71-99,146-238
      const real* hand = obs;
      real lid[3] = {obs[4], obs[5], obs[6] + (real)0.02};
      real grab = fmin(fmax((fmin(fmax(c.action[3], (real)-1), (real)1) + 1) / 2, (real)0), (real)1);
      real ideal[4] = {(real)0.707, 0, 0, (real)0.707};
      real qe = 0; for (int i = 0; i < 4; i++) { real d = obs[7 + i] - ideal[i]; qe += d * d; }
      real rq = fmax(1 - sqrt(qe) / (real)0.2, (real)0);
      real dx = hand[0] - lid[0], dy = hand[1] - lid[1], radius = sqrt(dx * dx + dy * dy);
      real floorh = radius <= (real)0.02 ? (real)0 : (real)0.04 * log(radius - (real)0.02) + (real)0.4;
      real above = hand[2] >= floorh ? (real)1 : tol_long_tail(floorh - hand[2], 0, (real)0.01, floorh / 2);
      real ready = hamacher(above, tol_long_tail(dist3(hand, lid), 0, (real)0.02, (real)0.5));
      real pe[3] = {target[0] - lid[0], target[1] - lid[1], 3 * (target[2] - lid[2])};
      real lifted = (real)0.2 * (lid[2] > (real)0.04) + (real)0.8 * tol_long_tail(v3norm(pe), 0, (real)0.05, (real)0.25);
      real r = 2 * hamacher(grab, ready) + 8 * lifted;
      bool succ = dist3(obs + 4, target) < (real)0.08;
      if (succ) r = 10;
      r *= rq;
      *reward = r;
      info[INFO_SUCCESS] = succ; info[INFO_NEAR_OBJECT] = ready; info[INFO_GRASP_SUCCESS] = grab >= (real)0.5;
      info[INFO_GRASP_REWARD] = grab; info[INFO_IN_PLACE] = lifted; info[INFO_OBJ_TO_TARGET] = 0; info[INFO_UNSCALED] = r;
    } break;
    case T_FAUCET_OPEN: case T_FAUCET_CLOSE: {   // sawyer_faucet_open_v3.py:63-85,125-175 ; sawyer_faucet_close_v3.py:64-86,127-174
      real obj[3] = {obs[4], obs[5], obs[6]};
      if (c.tc->task_id == T_FAUCET_OPEN) { obj[0] -= (real)0.04; obj[2] += (real)0.03; }
      real oi[3] = {c.s->obj_init[0], c.s->obj_init[1], c.s->obj_init[2]}, it[3] = {c.s->init_tcp[0], c.s->init_tcp[1], c.s->init_tcp[2]};
      real t2o = dist3(obj, target);
      real in_place = tol_long_tail(t2o, 0, (real)0.07, fabs(dist3(oi, target) - (real)0.07));
      real tcp_to_obj = dist3(obj, tcp);
      real reach = tol_gaussian(tcp_to_obj, 0, (real)0.01, fabs(dist3(oi, it) - (real)0.01));
      real r = 2 * (2 * reach + 3 * in_place);
      if (t2o <= (real)0.07) r = 10;
      *reward = r;
      info[INFO_SUCCESS] = t2o <= (real)0.07; info[INFO_NEAR_OBJECT] = tcp_to_obj <= (real)0.01; info[INFO_GRASP_SUCCESS] = 1;
      info[INFO_GRASP_REWARD] = reach; info[INFO_IN_PLACE] = in_place; info[INFO_OBJ_TO_TARGET] = t2o; info[INFO_UNSCALED] = r;
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
    case T_HANDLE_PRESS: case T_HANDLE_PRESS_SIDE: {   // sawyer_handle_press_v3.py:61-86,120-160 ; sawyer_handle_press_side_v3.py
      const real* obj = obs + 4;
      real h0[3] = {c.s->scal[0], c.s->scal[1], c.s->scal[2]}, it[3] = {c.s->init_tcp[0], c.s->init_tcp[1], c.s->init_tcp[2]};
      real t2o = fabs(obj[2] - target[2]), t2oi = fabs(h0[2] - target[2]);
      real in_place = tol_long_tail(t2o, 0, (real)0.02, fabs(t2oi - (real)0.02));
      real tcp_to_obj = dist3(obj, tcp);
      real reach = tol_long_tail(tcp_to_obj, 0, (real)0.02, fabs(dist3(h0, it) - (real)0.02));
      real r = hamacher(reach, in_place);
      if (t2o <= (real)0.02) r = 1;
      r *= 10;
      *reward = r;
      info[INFO_SUCCESS] = t2o <= (real)0.02; info[INFO_NEAR_OBJECT] = tcp_to_obj <= (real)0.05; info[INFO_GRASP_SUCCESS] = 1;
      info[INFO_GRASP_REWARD] = reach; info[INFO_IN_PLACE] = in_place; info[INFO_OBJ_TO_TARGET] = t2o; info[INFO_UNSCALED] = r;
    } break;
    case T_HANDLE_PULL: case T_HANDLE_PULL_SIDE: {   // sawyer_handle_pull_v3.py:60-90,124-166 ; sawyer_handle_pull_side_v3.py:62-92,128-174
      const bool side = c.tc->task_id == T_HANDLE_PULL_SIDE;
      const real* obj = obs + 4; real opened = obs[3];
      real oi[3] = {c.s->obj_init[0], c.s->obj_init[1], c.s->obj_init[2]};
      real t2o = side ? dist3(obj, target) : fabs(target[2] - obj[2]);
      real t2oi = side ? dist3(oi, target) : fabs(target[2] - oi[2]);
      real g = side ? gripper_caging_reward(c, obj, (real)0.032, (real)0.06, (real)0.01, (real)0.01, 1, 1)
                    : gripper_caging_reward(c, obj, (real)0.022, (real)0.05, (real)0.01, (real)0.01, 1, 1);
      real in_place = tol_long_tail(t2o, 0, (real)0.05, t2oi);
      real r = hamacher(g, in_place);
      real tcp_to_obj = dist3(obj, tcp);
      if (tcp_to_obj < (real)0.035 && opened > 0 && (obj[side ? 2 : 1] - (real)0.01 > oi[2])) r += 1 + 5 * in_place;   // (sic) y vs z in the non-side task
      if (t2o < (real)0.05) r = 10;
      *reward = r;
      info[INFO_SUCCESS] = t2o <= (side ? (real)0.08 : (real)0.05); info[INFO_NEAR_OBJECT] = tcp_to_obj <= (real)0.05;
      info[INFO_GRASP_SUCCESS] = opened > 0 && (obj[2] - (real)0.03 > oi[2]);
      info[INFO_GRASP_REWARD] = g; info[INFO_IN_PLACE] = in_place; info[INFO_OBJ_TO_TARGET] = t2o; info[INFO_UNSCALED] = r;
    } break;
    case T_LEVER_PULL: {   // sawyer_lever_pull_v3.py:70-100,128-190
      const real* grip = obs; const real* lever = obs + 4;
      real l0[3] = {c.s->scal[0], c.s->scal[1], c.s->scal[2]}, it[3] = {c.s->init_tcp[0], c.s->init_tcp[1], c.s->init_tcp[2]};
      const real off[3] = {0, (real)0.055, (real)0.07}, sc[3] = {4, 1, 4};
      real a[3], b[3];
      for (int i = 0; i < 3; i++) { a[i] = (grip[i] + off[i] - lever[i]) * sc[i]; b[i] = (it[i] + off[i] - l0[i]) * sc[i]; }
      real s2l = v3norm(a);
      real ready = tol_long_tail(s2l, 0, (real)0.02, v3norm(b));
      const real PI = (real)3.141592653589793;
      real err = fabs(-c.w->qpos[9] - PI / 2);
      real engagement = tol_long_tail(err, 0, PI / 48, PI / 2 - PI / 12);
      real in_place = tol_long_tail(dist3(lever, target), 0, (real)0.04, dist3(l0, target));
      real r = 10 * hamacher(ready, in_place);
      *reward = r;
      info[INFO_SUCCESS] = err <= PI / 24; info[INFO_NEAR_OBJECT] = s2l < (real)0.03; info[INFO_GRASP_SUCCESS] = ready > (real)0.9;
      info[INFO_GRASP_REWARD] = ready; info[INFO_IN_PLACE] = engagement; info[INFO_OBJ_TO_TARGET] = s2l; info[INFO_UNSCALED] = r;
    } break;
    case T_PEG_UNPLUG_SIDE: {   // sawyer_peg_unplug_side_v3.py:61-88,118-169
      const real* obj = obs + 4; real opened = obs[3];
      real oi[3] = {c.s->obj_init[0], c.s->obj_init[1], c.s->obj_init[2]};
      real tcp_to_obj = dist3(obj, tcp), o2t = dist3(obj, target);
      real g = gripper_caging_reward(c, obj, (real)0.025, (real)0.05, (real)0.01, (real)0.005, (real)0.8, 1);
      real in_place = tol_long_tail(o2t, 0, (real)0.05, dist3(oi, target));
      bool gs = opened > (real)0.5 && (obj[0] - oi[0] > (real)0.015);
      real r = 2 * g;
      if (gs && tcp_to_obj < (real)0.035) r = 1 + 2 * g + 5 * in_place;
      if (o2t <= (real)0.05) r = 10;
      *reward = r;
      info[INFO_SUCCESS] = o2t <= (real)0.07; info[INFO_NEAR_OBJECT] = tcp_to_obj <= (real)0.03; info[INFO_GRASP_SUCCESS] = gs;
      info[INFO_GRASP_REWARD] = g; info[INFO_IN_PLACE] = in_place; info[INFO_OBJ_TO_TARGET] = o2t; info[INFO_UNSCALED] = r;
    } break;
    case T_PLATE_SLIDE: case T_PLATE_SLIDE_SIDE: case T_PLATE_SLIDE_BACK: case T_PLATE_SLIDE_BACK_SIDE: {
      // sawyer_plate_slide_v3.py:66-95,130-172 ; _side_v3.py:128-170 ; _back_v3.py:127-168 ; _back_side_v3.py:151-194
      const bool vb = c.tc->task_id != T_PLATE_SLIDE;
      const real* obj = obs + 4;
      real oi[3] = {c.s->obj_init[0], c.s->obj_init[1], c.s->obj_init[2]}, it[3] = {c.s->init_tcp[0], c.s->init_tcp[1], c.s->init_tcp[2]};
      real o2t = dist3(obj, target), tcp_to_obj = dist3(tcp, obj);
      real in_place = tol_long_tail(o2t, 0, (real)0.05, dist3(oi, target) - (vb ? (real)0.05 : (real)0));
      real grasped = tol_long_tail(tcp_to_obj, 0, (real)0.05, dist3(it, oi) - (vb ? (real)0.05 : (real)0));
      real r;
      if (vb) { r = (real)1.5 * grasped; if (tcp[2] <= (real)0.03 && tcp_to_obj < (real)0.07) r = 2 + 7 * in_place; }
      else r = 8 * hamacher(grasped, in_place);
      if (o2t < (real)0.05) r = 10;
      *reward = r;
      info[INFO_SUCCESS] = o2t <= (real)0.07; info[INFO_NEAR_OBJECT] = tcp_to_obj <= (real)0.03; info[INFO_GRASP_SUCCESS] = 0;
      info[INFO_GRASP_REWARD] = grasped; info[INFO_IN_PLACE] = in_place; info[INFO_OBJ_TO_TARGET] = o2t; info[INFO_UNSCALED] = r;
    } break;
    case T_SHELF_PLACE: {   // sawyer_shelf_place_v3.py:70-100,150-220
      const real* obj = obs + 4; real opened = obs[3];
      real oi[3] = {c.s->obj_init[0], c.s->obj_init[1], c.s->obj_init[2]};
      real o2t = dist3(obj, target), tcp_to_obj = dist3(obj, tcp);
      real in_place = tol_long_tail(o2t, 0, (real)0.05, dist3(oi, target));
      real g = gripper_caging_reward(c, obj, (real)0.02, (real)0.05, (real)0.01, (real)0.01, 1, 0);
      real r = hamacher(g, in_place);
      bool inx = target[0] - (real)0.15 < obj[0] && obj[0] < target[0] + (real)0.15, inz = (real)0 < obj[2] && obj[2] < (real)0.24;
      if (inz && inx && (target[1] - (real)0.15 < obj[1] && obj[1] < target[1])) {
        real zs = ((real)0.24 - obj[2]) / (real)0.24, ys = (obj[1] - (target[1] - (real)0.15)) / (real)0.15;
        in_place = fmin(fmax(in_place - hamacher(ys, zs), (real)0), (real)1);
      }
      if (inz && inx && obj[1] > target[1]) in_place = 0;
      if (tcp_to_obj < (real)0.025 && opened > 0 && (obj[2] - (real)0.01 > oi[2])) r += 1 + 5 * in_place;
      if (o2t < (real)0.05) r = 10;
      *reward = r;
      bool touch = touching_object(c, c.tc->main_geom, (int)c.tc->p[14], (int)c.tc->p[15]);
      info[INFO_SUCCESS] = o2t <= (real)0.07; info[INFO_NEAR_OBJECT] = tcp_to_obj <= (real)0.03;
      info[INFO_GRASP_SUCCESS] = touch && opened > 0 && (obj[2] - (real)0.02 > oi[2]);
      info[INFO_GRASP_REWARD] = g; info[INFO_IN_PLACE] = in_place; info[INFO_OBJ_TO_TARGET] = o2t; info[INFO_UNSCALED] = r;
    } break;
    case T_SOCCER: {   // sawyer_soccer_v3.py:68-97,131-262
      const real* obj = obs + 4; real opened = obs[3];
      real oi[3] = {c.s->obj_init[0], c.s->obj_init[1], c.s->obj_init[2]};
      real a[3] = {3 * (obj[0] - target[0]), obj[1] - target[1], obj[2] - target[2]};
      real b[3] = {3 * (obj[0] - oi[0]), obj[1] - oi[1], obj[2] - oi[2]};
      real t2o = v3norm(a), tcp_to_obj = dist3(obj, tcp);
      real in_place = tol_long_tail(t2o, 0, (real)0.07, v3norm(b));
      real gl = target[1] - (real)0.1;
      if (obj[1] > gl && fabs(obj[0] - target[0]) > (real)0.10) in_place = fmin(fmax(in_place - 2 * ((obj[1] - gl) / (1 - gl)), (real)0), (real)1);
      real g = grip_caging(c, obj, (real)0.013, (real)0.01, (real)0.005);
      real r = 3 * g + (real)6.5 * in_place;
      if (t2o < (real)0.07) r = 10;
      *reward = r;
      real plain = dist3(obj, target);
      bool touch = touching_object(c, c.tc->main_geom, (int)c.tc->p[14], (int)c.tc->p[15]);
      info[INFO_SUCCESS] = plain <= (real)0.07; info[INFO_NEAR_OBJECT] = tcp_to_obj <= (real)0.03;
      info[INFO_GRASP_SUCCESS] = touch && opened > 0 && (obj[2] - (real)0.02 > oi[2]);
      info[INFO_GRASP_REWARD] = g; info[INFO_IN_PLACE] = in_place; info[INFO_OBJ_TO_TARGET] = plain; info[INFO_UNSCALED] = r;
    } break;
    case T_STICK_PULL: {   // sawyer_stick_pull_v3.py:68-104,170-290 (shared caging with obj_init_pos = the container's initial position)
      const real* stick = obs + 4; const real* handle = obs + 11; real opened = obs[3];
      real eos[3]; mw_frame_pos(c.m, c.w, F_TASK0 + 3, eos);
      real oi[3] = {c.s->obj_init[0], c.s->obj_init[1], c.s->obj_init[2]}, s0[3] = {c.s->scal[0], c.s->scal[1], c.s->scal[2]};
      real cont[3] = {handle[0] + (real)0.05, handle[1], handle[2]}, cont0[3] = {oi[0] + (real)0.05, oi[1], oi[2]};
      real tcp_to_stick = dist3(stick, tcp), h2t = dist3(handle, target);
      real a[3] = {stick[0] - cont[0], stick[1] - cont[1], 2 * (stick[2] - cont[2])};
      real b[3] = {s0[0] - cont0[0], s0[1] - cont0[1], 2 * (s0[2] - cont0[2])};
      real sip = tol_long_tail(v3norm(a), 0, (real)0.05, v3norm(b));
      real sip2 = tol_long_tail(dist3(stick, target), 0, (real)0.05, dist3(s0, target));
      real cip = tol_long_tail(dist3(cont, target), 0, (real)0.05, dist3(oi, target));
      real g = gripper_caging_reward(c, stick, (real)0.014, (real)0.05, (real)0.01, (real)0.01, 1, 1);
      bool gs = tcp_to_stick < (real)0.02 && opened > 0 && (stick[2] - (real)0.01 > s0[2]);
      if (gs) g = 1;
      bool inserted = eos[0] >= handle[0] && fabs(eos[1] - handle[1]) <= (real)0.040 && fabs(eos[2] - handle[2]) <= (real)0.060;
      real ipg = hamacher(g, sip), r = ipg;
      if (gs) {
        r = 1 + ipg + 5 * sip;
        if (inserted) { r = 1 + ipg + 5 + 2 * sip2 + cip; if (h2t <= (real)0.12) r = 10; }
      }
      *reward = r;
      bool touch = touching_object(c, c.tc->main_geom, (int)c.tc->p[14], (int)c.tc->p[15]);
      info[INFO_SUCCESS] = h2t <= (real)0.12 && inserted; info[INFO_NEAR_OBJECT] = tcp_to_stick <= (real)0.03;
      info[INFO_GRASP_SUCCESS] = touch && opened > 0 && (stick[2] - (real)0.02 > oi[2]);
      info[INFO_GRASP_REWARD] = g; info[INFO_IN_PLACE] = sip; info[INFO_OBJ_TO_TARGET] = h2t; info[INFO_UNSCALED] = r;
    } break;
    case T_STICK_PUSH: {   // sawyer_stick_push_v3.py:66-98,168-345 (own caging copy: stick_init_pos in place of obj_init_pos)
      const real* cont = obs + 11; real opened = obs[3];
      real stick[3] = {obs[4] + (real)0.015, obs[5], obs[6]};
      real oi[3] = {c.s->obj_init[0], c.s->obj_init[1], c.s->obj_init[2]}, s0[3] = {c.s->scal[0], c.s->scal[1], c.s->scal[2]};
      real tcp_to_stick = dist3(stick, tcp), c2t = dist3(cont, target);
      real sip = tol_long_tail(dist3(stick, target), 0, (real)0.12, dist3(s0, target) - (real)0.12);
      real cip = tol_long_tail(c2t, 0, (real)0.12, dist3(oi, target) - (real)0.12);
      for (int i = 0; i < 3; i++) c.s->obj_init[i] = (float)s0[i];
      real g = gripper_caging_reward(c, stick, (real)0.04, (real)0.05, (real)0.01, (real)0.01, 1, 1);
      for (int i = 0; i < 3; i++) c.s->obj_init[i] = (float)oi[i];
      real r = g;
      if (tcp_to_stick < (real)0.02 && opened > 0 && (stick[2] - (real)0.01 > s0[2])) {
        g = 1; r = 2 + 5 * sip + 3 * cip;
        if (c2t <= (real)0.12) r = 10;
      }
      *reward = r;
      bool touch = touching_object(c, c.tc->main_geom, (int)c.tc->p[14], (int)c.tc->p[15]);
      bool gsx = touch && opened > 0 && (obs[6] - (real)0.01 > s0[2]);
      info[INFO_SUCCESS] = gsx && c2t <= (real)0.12; info[INFO_NEAR_OBJECT] = tcp_to_stick <= (real)0.03; info[INFO_GRASP_SUCCESS] = gsx;
      info[INFO_GRASP_REWARD] = g; info[INFO_IN_PLACE] = sip; info[INFO_OBJ_TO_TARGET] = c2t; info[INFO_UNSCALED] = r;
    } break;
    default: *reward = 0; break;
  }
  (void)obs;
}

// ---------------------------------------------------------------- reset_model (after _reset_hand); all lanes call
// rv = the task's frozen rand_vec (Task.data['rand_vec']); writes env constants into c.s (lane 0)
DEV void task_reset_model(const TaskCtx& c, const double* rv, int lane) {
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
        QSET(c.w, 9, 0); c.w->qvel[9] = 0;
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
        QSET(c.w, 9, (real)-0.15);
      }
      SYNCW();
      eng_forward(c, lane);
      if (lane == 0) { real o[3]; mw_frame_pos(c.m, c.w, F_TASK0, o); c.s->obj_init[0] = (float)o[0]; c.s->obj_init[1] = (float)(o[1] - (real)0.16); c.s->obj_init[2] = (float)(o[2] + (real)0.05); }
      SYNCW();
    } break;
    case T_BUTTON_PRESS_TOPDOWN: case T_BUTTON_PRESS_TOPDOWN_WALL: case T_BUTTON_PRESS: case T_BUTTON_PRESS_WALL: {
      // sawyer_button_press_topdown_v3.py:105-120 (mj_forward) ; sawyer_button_press_v3.py:105-124 (_set_obj_xyz(0): qpos[9]=qvel[9]=0, forward)
      const bool side = c.tc->task_id == T_BUTTON_PRESS || c.tc->task_id == T_BUTTON_PRESS_WALL;
      if (lane == 0) {
        for (int i = 0; i < 3; i++) { c.s->obj_init[i] = rv[i]; c.w->shift[i] = rv[i] - c.tc->movable_pos0[i]; }
        if (side) { QSET(c.w, 9, 0); c.w->qvel[9] = 0; }
      }
      SYNCW();
      eng_forward(c, lane);
      if (lane == 0) {
        real hole[3], bs[3]; mw_frame_pos(c.m, c.w, F_TASK0 + 1, hole); mw_frame_pos(c.m, c.w, F_TASK0 + 2, bs);
        for (int i = 0; i < 3; i++) c.s->target[i] = (float)hole[i];
        const int ax = side ? 1 : 2;
        c.s->scal[0] = (float)fabs(hole[ax] - bs[ax]);
      }
      SYNCW();
    } break;
    case T_COFFEE_BUTTON: case T_COFFEE_PULL: case T_COFFEE_PUSH: {
      // sawyer_coffee_button_v3.py:114-131 ; sawyer_coffee_pull_v3.py:117-137 ; sawyer_coffee_push_v3.py:117-138
      // the mug's free joint precedes the robot: _set_obj_xyz writes qpos[0:3] and (reference quirk) zeroes qvel[9:15]
      const int id = c.tc->task_id;
      real mug[3] = {rv[0], rv[1], rv[2]}, mach[3];
      if (id == T_COFFEE_BUTTON) { mug[1] -= (real)0.22; for (int i = 0; i < 3; i++) mach[i] = rv[i]; }
      else { for (int i = 0; i < 3; i++) mach[i] = (id == T_COFFEE_PULL ? rv[i] : rv[3 + i]); mach[1] += (real)0.22; }
      if (lane == 0 && id == T_COFFEE_BUTTON) for (int i = 0; i < 3; i++) c.w->shift[i] = mach[i] - c.tc->movable_pos0[i];   // body pos set BEFORE the forward
      if (lane == 0) { for (int i = 0; i < 3; i++) QSET(c.w, i, mug[i]); for (int i = 9; i < 15; i++) c.w->qvel[i] = 0; }
      SYNCW();
      eng_forward(c, lane);
      if (lane == 0) {
        for (int i = 0; i < 3; i++) c.w->shift[i] = mach[i] - c.tc->movable_pos0[i];
        if (id == T_COFFEE_BUTTON) { for (int i = 0; i < 3; i++) c.s->obj_init[i] = rv[i]; c.s->target[0] = rv[0]; c.s->target[1] = rv[1] - 0.22f + 0.03f; c.s->target[2] = rv[2] + 0.3f; }
        else for (int i = 0; i < 3; i++) { c.s->obj_init[i] = rv[i]; c.s->target[i] = rv[3 + i]; }
      }
      SYNCW();
    } break;
    case T_ASSEMBLY: {   // sawyer_assembly_peg_v3.py:115-145
      real p[3] = {rv[0], rv[1], rv[2]};
      if (lane == 0) for (int i = 0; i < 3; i++) { c.s->obj_init[i] = rv[i]; c.s->target[i] = rv[3 + i]; }
      SYNCW();
      set_obj_xyz(c, p, lane);
      if (lane == 0) { c.w->shift[0] = rv[3] - c.tc->movable_pos0[0]; c.w->shift[1] = rv[4] - c.tc->movable_pos0[1]; c.w->shift[2] = rv[5] - (real)0.05 - c.tc->movable_pos0[2]; }
      SYNCW();
    } break;
    case T_DISASSEMBLE: {   // sawyer_disassemble_peg_v3.py:116-133 (mj_forward, then _set_obj_xyz)
      real p[3] = {rv[0], rv[1], rv[2]};
      if (lane == 0) {
        for (int i = 0; i < 3; i++) { c.s->obj_init[i] = rv[i]; c.s->target[i] = rv[i]; c.w->shift[i] = rv[i] - c.tc->movable_pos0[i]; }
        c.s->target[2] = rv[2] + 0.15f; c.w->shift[2] += (real)0.03;
      }
      SYNCW();
      eng_forward(c, lane);
      set_obj_xyz(c, p, lane);
    } break;
    case T_BASKETBALL: {   // sawyer_basketball_v3.py:109-123.  `_target_pos` is a live view of data.site("goal").xpos and the site's LOCAL
      // offset is overwritten with that world position (model.site("goal").pos = _target_pos): scal[3:6] carries the local offset.
      real p[3] = {rv[0], rv[1], (real)0.03};
      if (lane == 0) { for (int i = 0; i < 3; i++) { c.s->obj_init[i] = (float)p[i]; c.w->shift[i] = rv[3 + i] - c.tc->movable_pos0[i]; } }
      SYNCW();
      set_obj_xyz(c, p, lane);
      if (lane == 0) {
        real b[3]; mw_frame_pos(c.m, c.w, F_TASK0 + 1, b);
        for (int i = 0; i < 3; i++) { float t = (float)(b[i] + (real)c.s->scal[3 + i]); c.s->target[i] = t; c.s->scal[3 + i] = t; }
      }
      SYNCW();
    } break;
    case T_BIN_PICKING: {   // sawyer_bin_picking_v3.py:131-159
      real ob[3]; mw_frame_pos(c.m, c.w, F_TASK0, ob);
      real p[3] = {rv[0], rv[1], ob[2]};
      if (lane == 0) for (int i = 0; i < 3; i++) c.s->obj_init[i] = (float)p[i];
      SYNCW();
      set_obj_xyz(c, p, lane);
      if (lane == 0) { real g[3]; mw_frame_pos(c.m, c.w, F_TASK0 + 1, g); for (int i = 0; i < 3; i++) c.s->target[i] = (float)g[i]; c.s->scal[0] = -1.f; }
      SYNCW();
    } break;
    case T_BOX_CLOSE: {   // sawyer_box_close_v3.py:107-144 (5 x mj_step after moving the box)
      real bb[3]; mw_frame_pos(c.m, c.w, F_TASK0 + 1, bb);
      real p[3] = {rv[0], rv[1], (real)0.02};
      if (lane == 0) {
        for (int i = 0; i < 3; i++) { c.s->obj_init[i] = (float)p[i]; c.s->target[i] = rv[3 + i]; }
        c.w->shift[0] = rv[3] - c.tc->movable_pos0[0]; c.w->shift[1] = rv[4] - c.tc->movable_pos0[1]; c.w->shift[2] = bb[2] - c.tc->movable_pos0[2];
      }
      SYNCW();
      eng_sim(c, 5, lane);
      set_obj_xyz(c, p, lane);
    } break;
    case T_FAUCET_OPEN: case T_FAUCET_CLOSE: {   // sawyer_faucet_open_v3.py:103-123 ; sawyer_faucet_close_v3.py:104-125 (mj_forward)
      const bool closing = c.tc->task_id == T_FAUCET_CLOSE;
      if (lane == 0) {
        for (int i = 0; i < 3; i++) { c.s->obj_init[i] = rv[i]; c.s->target[i] = rv[i]; c.w->shift[i] = rv[i] - c.tc->movable_pos0[i]; }
        c.s->target[0] = rv[0] + (closing ? -0.175f : 0.175f); c.s->target[2] = rv[2] + 0.125f;
      }
      SYNCW();
      if (closing) eng_forward(c, lane);
    } break;
    case T_HAMMER: {   // sawyer_hammer_v3.py:108-144
      real p[3] = {rv[0], rv[1], rv[2]};
      if (lane == 0) {
        c.w->shift[0] = (real)0.24 - c.tc->movable_pos0[0]; c.w->shift[1] = (real)0.85 - c.tc->movable_pos0[1]; c.w->shift[2] = -c.tc->movable_pos0[2];
        real g[3]; mw_frame_pos(c.m, c.w, F_TASK0 + 2, g);
        for (int i = 0; i < 3; i++) { c.s->target[i] = (float)g[i]; c.s->obj_init[i] = rv[i]; }
      }
      SYNCW();
      set_obj_xyz(c, p, lane);
    } break;
    case T_DIAL_TURN: {   // sawyer_dial_turn_v3.py:103-120 (no mj_forward)
      if (lane == 0) {
        for (int i = 0; i < 3; i++) { c.s->obj_init[i] = rv[i]; c.w->shift[i] = rv[i] - c.tc->movable_pos0[i]; }
        c.s->target[0] = rv[0]; c.s->target[1] = rv[1] + 0.03f; c.s->target[2] = rv[2] + 0.03f;
      }
      SYNCW();
      if (lane == 0) { real ob[14]; task_obs_objects(c, ob); c.s->scal[0] = (float)(ob[0] + (real)0.05); c.s->scal[1] = (float)(ob[1] + (real)0.02); c.s->scal[2] = (float)(ob[2] + (real)0.09); }
      SYNCW();
    } break;
    case T_DOOR_CLOSE: {   // sawyer_door_close_v3.py:82-103
      if (lane == 0) {
        for (int i = 0; i < 3; i++) { c.s->obj_init[i] = rv[i]; c.w->shift[i] = rv[i] - c.tc->movable_pos0[i]; }
        c.s->target[0] = rv[0] + 0.2f; c.s->target[1] = rv[1] - 0.2f; c.s->target[2] = rv[2];
        QSET(c.w, 9, (real)-1.5708); c.w->qvel[9] = 0;
      }
      SYNCW();
      eng_forward(c, lane);
    } break;
    case T_DOOR_LOCK: case T_DOOR_UNLOCK: {   // sawyer_door_lock_v3.py:108-121 (5 x mj_step) ; sawyer_door_unlock_v3.py:113-124 (qpos[9]=1.5708, forward)
      const bool lock = c.tc->task_id == T_DOOR_LOCK;
      if (lane == 0) {
        for (int i = 0; i < 3; i++) c.w->shift[i] = rv[i] - c.tc->movable_pos0[i];
        if (!lock) { QSET(c.w, 9, (real)1.5708); c.w->qvel[9] = 0; }
      }
      SYNCW();
      if (lock) eng_sim(c, 5, lane); else eng_forward(c, lane);
      if (lane == 0) {
        real ll[3]; mw_frame_pos(c.m, c.w, F_TASK0 + 2, ll);   // data.body("lock_link").xpos as left by the last kinematics
        for (int i = 0; i < 3; i++) c.s->obj_init[i] = (float)ll[i];
        c.s->target[0] = (float)(ll[0] + (lock ? (real)0 : (real)0.1)); c.s->target[1] = (float)(ll[1] - (real)0.04); c.s->target[2] = (float)(ll[2] - (lock ? (real)0.1 : (real)0));
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
        QSET(c.w, 9, closing ? (real)0.2 : (real)0);            // data.joint("window_slide").qpos = ..., no mj_forward
      }
      SYNCW();
    } break;
    case T_HANDLE_PRESS: case T_HANDLE_PRESS_SIDE: case T_HANDLE_PULL: case T_HANDLE_PULL_SIDE: {
      // sawyer_handle_press_v3.py:107-118 ; sawyer_handle_pull_v3.py:109-122 ; _side variants
      const int id = c.tc->task_id;
      const bool pull = id == T_HANDLE_PULL || id == T_HANDLE_PULL_SIDE;
      if (lane == 0) {
        for (int i = 0; i < 3; i++) { c.s->obj_init[i] = rv[i]; c.w->shift[i] = rv[i] - c.tc->movable_pos0[i]; }
        QSET(c.w, 9, pull ? (real)-0.1 : (real)-0.001); c.w->qvel[9] = 0;
      }
      SYNCW();
      eng_forward(c, lane);
      if (lane == 0) {
        real g[3], h[3]; mw_frame_pos(c.m, c.w, F_TASK0 + 1, g); mw_frame_pos(c.m, c.w, F_TASK0, h);
        for (int i = 0; i < 3; i++) { c.s->target[i] = (float)g[i]; c.s->scal[i] = (float)h[i]; }
        if (id == T_HANDLE_PULL_SIDE) for (int i = 0; i < 3; i++) c.s->obj_init[i] = (float)h[i];
      }
      SYNCW();
    } break;
    case T_LEVER_PULL: {   // sawyer_lever_pull_v3.py:114-126 (no mj_forward)
      if (lane == 0) {
        for (int i = 0; i < 3; i++) { c.s->obj_init[i] = rv[i]; c.w->shift[i] = rv[i] - c.tc->movable_pos0[i]; }
        c.s->scal[0] = rv[0] + 0.12f; c.s->scal[1] = rv[1] - 0.2f; c.s->scal[2] = rv[2] + 0.25f;
        c.s->target[0] = rv[0] + 0.12f; c.s->target[1] = rv[1]; c.s->target[2] = rv[2] + 0.25f + 0.2f;
      }
      SYNCW();
    } break;
    case T_PEG_UNPLUG_SIDE: {   // sawyer_peg_unplug_side_v3.py:96-116
      if (lane == 0) {
        for (int i = 0; i < 3; i++) c.w->shift[i] = rv[i] - c.tc->movable_pos0[i];
        QSET(c.w, 9, rv[0] + (real)0.044); QSET(c.w, 10, rv[1]); QSET(c.w, 11, rv[2] + (real)0.131);
        QSET(c.w, 12, 1); QSET(c.w, 13, 0); QSET(c.w, 14, 0); QSET(c.w, 15, 0);
        for (int i = 9; i < 12; i++) c.w->qvel[i] = 0;
      }
      SYNCW();
      eng_forward(c, lane);
      if (lane == 0) {
        real e[3]; mw_frame_pos(c.m, c.w, F_TASK0, e);
        for (int i = 0; i < 3; i++) c.s->obj_init[i] = (float)e[i];
        c.s->target[0] = (float)(rv[0] + (real)0.044 + (real)0.15); c.s->target[1] = rv[1]; c.s->target[2] = (float)(rv[2] + (real)0.131);
      }
      SYNCW();
    } break;
    case T_PLATE_SLIDE: case T_PLATE_SLIDE_SIDE: case T_PLATE_SLIDE_BACK: case T_PLATE_SLIDE_BACK_SIDE: {
      // sawyer_plate_slide_v3.py:104-123 (model.body puck_goal := goal) ; _side/_back (data xpos write: transient) ; _back_side (:= obj_init)
      const int id = c.tc->task_id;
      if (lane == 0) {
        for (int i = 0; i < 3; i++) { c.s->obj_init[i] = rv[i]; c.s->target[i] = rv[3 + i]; }
        if (id == T_PLATE_SLIDE) for (int i = 0; i < 3; i++) c.w->shift[i] = rv[3 + i] - c.tc->movable_pos0[i];
        if (id == T_PLATE_SLIDE_BACK_SIDE) for (int i = 0; i < 3; i++) c.w->shift[i] = rv[i] - c.tc->movable_pos0[i];
        QSET(c.w, 9, id == T_PLATE_SLIDE_BACK_SIDE ? (real)-0.15 : (real)0);
        QSET(c.w, 10, id == T_PLATE_SLIDE_BACK ? (real)0.15 : (real)0);
      }
      SYNCW();
      eng_forward(c, lane);
    } break;
    case T_SHELF_PLACE: {   // sawyer_shelf_place_v3.py:118-140
      real ob[3]; mw_frame_pos(c.m, c.w, F_TASK0, ob);
      real p[3] = {rv[0], rv[1], ob[2]};
      if (lane == 0) {
        real sh[3] = {rv[3], rv[4], rv[5] - (real)0.3};
        for (int i = 0; i < 3; i++) { c.s->obj_init[i] = (float)p[i]; c.w->shift[i] = sh[i] - c.tc->movable_pos0[i]; c.s->target[i] = (float)((real)c.tc->p[i] + sh[i]); }
      }
      SYNCW();
      eng_forward(c, lane);
      set_obj_xyz(c, p, lane);
    } break;
    case T_SOCCER: {   // sawyer_soccer_v3.py:106-129
      real p[3] = {rv[0], rv[1], (real)0.03};
      if (lane == 0) for (int i = 0; i < 3; i++) { c.s->obj_init[i] = (float)p[i]; c.s->target[i] = rv[3 + i]; c.w->shift[i] = rv[3 + i] - c.tc->movable_pos0[i]; }
      SYNCW();
      set_obj_xyz(c, p, lane);
    } break;
    case T_STICK_PUSH: case T_STICK_PULL: {   // sawyer_stick_push_v3.py:139-166 ; sawyer_stick_pull_v3.py:142-169
      const bool pull = c.tc->task_id == T_STICK_PULL;
      real ins[3]; mw_frame_pos(c.m, c.w, F_TASK0 + 1, ins);          // stale insertion-site height (push target z)
      real p[3] = {rv[0], rv[1], (real)0.02};
      if (lane == 0) {
        for (int i = 0; i < 3; i++) c.s->scal[i] = (float)p[i];
        c.s->target[0] = rv[3]; c.s->target[1] = rv[4]; c.s->target[2] = pull ? 0.02f : (float)ins[2];
      }
      SYNCW();
      set_obj_xyz(c, p, lane);                                         // _set_stick_xyz: qpos[9:12], qvel[9:15] = 0, forward
      if (lane == 0) { QSET(c.w, 16, 0); QSET(c.w, 17, pull ? (real)0.09 : (real)0); c.w->qvel[16] = 0; }   // qvel[16:18] on nv = 17 reaches dof 16 only
      SYNCW();
      eng_forward(c, lane);
      if (lane == 0) { real o[3]; mw_frame_pos(c.m, c.w, F_TASK0 + 2, o); for (int i = 0; i < 3; i++) c.s->obj_init[i] = (float)o[i]; }
      SYNCW();
    } break;
    default: break;
  }
}
