/*
 * ballbot_oracle.h -- C interface of the fp64 single-env CPU ORACLE.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing in the product package may import, link or
 * call this library; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do.
 *
 * PARITY UNPINNED: the reference's arithmetic lives in MuJoCo (git 99490163...
 * + tools/mujoco_fix.patch), `noise` (snoise2) and numpy-quaternion, none of
 * which exist in this container, and the reference's own tests pin no numeric
 * physics result.  This file restates MuJoCo's published
 * pipeline for the one model in ballbot_gym/models/ballbot.xml.  See
 * oracle/ORACLE_ASSUMPTIONS.md for every from-memory choice.
 */
#ifndef BALLBOT_ORACLE_H
#define BALLBOT_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define BBO_NQ 17
#define BBO_NV 15
#define BBO_NU 3
#define BBO_HF_N 293
#define BBO_MAXCON 64
#define BBO_OBS_PROPRIO 16 /* orientation3 angular_vel3 vel3 motor_state3 actions3 rel_img_ts1 */

typedef struct bbo_env bbo_env;

typedef struct bbo_config {
  int max_ep_steps;          /* ballbot_env.py:221  (4000) */
  double max_allowed_tilt;   /* ballbot_env.py:222  (20 deg) */
  double max_wheel_velocity; /* ballbot_env.py:223  (10) */
  double camera_frame_rate;  /* ballbot_env.py:224  (90 Hz) */
  double reward_scale;       /* ballbot_env.py:229  (0.01) */
  double action_reg_coef;    /* ballbot_env.py:230  (-1e-4) */
  double survival_bonus;     /* ballbot_env.py:231  (0.02) */
  double target_dir[2];      /* rewards/directional.py:33 */
  double hfield_zscale;      /* ballbot.xml:23 size[2] (2.0); ramp/gradient mutate it, ballbot_env.py:486-495 */
  int cameras;               /* 0: disable_cameras=True */
  int im_h, im_w;            /* 64 x 64 */
  int reward_type;           /* 0 directional (rewards/directional.py:33-54), 1 distance (rewards/distance.py:33-50 on info pos2d) */
  double goal[2];            /* distance reward: goal_position */
  double distance_scale;     /* distance reward: scale */
} bbo_config;

void bbo_default_config(bbo_config* cfg);
bbo_env* bbo_create(const bbo_config* cfg);
void bbo_destroy(bbo_env* e);

/* reset: hfield (293*293 float32, row-major [row=y][col=x], values in [0,1]) or NULL = flat.
 * Follows ballbot_env.py:567-671 (+ _reset_terrain :442-565).  obs: 16 floats. */
int bbo_reset(bbo_env* e, const float* hfield, float* obs16);
/* step: ballbot_env.py:854-1036.  info4 = {pos2d.x, pos2d.y, step_counter, cam_refreshed} */
int bbo_step(bbo_env* e, const float* action3, float* obs16, float* reward,
             uint8_t* terminated, uint8_t* failure, float* info4);
/* depth images of the last camera refresh, 2 x (h*w) float32 */
int bbo_get_depth(bbo_env* e, float* img0, float* img1);

int bbo_get_state(bbo_env* e, double* qpos17, double* qvel15, double* warm15, double* time);
int bbo_set_state(bbo_env* e, const double* qpos17, const double* qvel15, const double* warm15, double time);
int bbo_set_hfield(bbo_env* e, const float* hfield);

/* raw mj_step equivalent on the current state with ctrl (already in actuator units) */
int bbo_mj_step(bbo_env* e, const double* ctrl3);
/* one mj_forward at the current state; returns intermediate quantities (any pointer may be NULL) */
int bbo_forward(bbo_env* e, const double* ctrl3, double* qM225, double* qfrc_bias15,
                double* qacc_smooth15, double* qacc15, int* ncon, int* niter);
/* contacts of the most recent forward: dist[ncon], pos[3*ncon], frame[9*ncon], geom pair id */
int bbo_get_contacts(bbo_env* e, int maxcon, double* dist, double* pos, double* frame, int* pair);
/* efc of the most recent forward: J [nefc*15], aref, D, force, returns nefc */
int bbo_get_efc(bbo_env* e, int maxefc, double* J, double* aref, double* D, double* force);
/* model constants: body masses (8), invweight0 (16), meaninertia, subtree etc. */
int bbo_get_model(double* body_mass8, double* body_ipos24, double* body_inertia72,
                  double* invweight0_16, double* meaninertia);
/* obs-related kinematic quantities of the most recent forward */
int bbo_get_kin(bbo_env* e, double* xpos_base3, double* xquat_base4, double* cvel_base6);

/* noise.snoise2(x, y, octaves, persistence, lacunarity, repeatx=1024, repeaty=1024, base=seed),
 * restated (terrain/perlin.py:56-65), and the full terrain/perlin.py:8-74 generator */
float bbo_snoise2_tiled(float x, float y, int octaves, float persistence, float lacunarity,
                        float repeatx, float repeaty, int base);
/* noise.snoise2(x, y, octaves, persistence, lacunarity, base=seed) without repeats (terrain/gradient.py:74-80) */
float bbo_snoise2(float x, float y, int octaves, float persistence, float lacunarity, int base);
int bbo_perlin_terrain(int n, double scale, int octaves, double persistence, double lacunarity,
                       double amplitude, int seed, float* out);
/* spawn offset of ballbot_env.py:528-565 for a given hfield */
double bbo_spawn_offset(const float* hfield, double zscale);

/* CPU ray-cast depth for given base pose (test helper) */
int bbo_render_depth(bbo_env* e, int cam, float* img);

#ifdef __cplusplus
}
#endif
#endif
