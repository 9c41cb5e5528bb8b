/*
 * ssl_oracle.h -- CPU oracle for the batched SSL world-stepper.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (ssl-sim_b200/,
 * include/) may include, link or call this.  Only tests/, bench.py's
 * cpu_baseline / --impl reference leg and __graft_entry__.smoke() use it,
 * and there only as the checker / the timed CPU baseline.
 *
 * PARITY UNPINNED: the reference (roboime/ssl-sim) delegates the whole step to
 * btDiscreteDynamicsWorld::stepSimulation (reference src/sslsim.cpp:323), a
 * Bullet Physics dependency that is neither vendored nor version-pinned
 * (reference CMakeLists.txt:31, .travis.yml:24) and is absent from this
 * environment; the reference ships no tests, golden vectors or fixtures.
 * This file is therefore a plain-C restatement of the Bullet 2.82 step *as
 * exercised by the reference scene*, following docs/STEP_SPEC.md.  Every
 * scene constant cites the reference line it comes from.  (What the reference
 * can still build without Bullet -- its field geometry, src/field.c -- is
 * compiled in place into oracle/_ref/ and pins the field constants used here:
 * tests/test_reference_field.py.)
 *
 * One world at a time, scalar, single thread, straightforward loops.
 */
#ifndef SSL_ORACLE_H
#define SSL_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_BODIES 32 /* 31 robots + 1 ball (lane-per-body limit of the product) */

typedef struct { float x, y, z, w; } orc_vec4;

/* reference src/field.h:18-34 -- 15 floats */
typedef struct {
  float line_width, field_length, field_width, boundary_width, referee_width;
  float goal_width, goal_depth, goal_wall_width, center_circle_radius;
  float defense_radius, defense_stretch, free_kick_from_defense_dist;
  float penalty_spot_from_field_line_dist, penalty_line_from_spot_dist;
  float goal_height;
} orc_field;

/* builder-defined tunables (docs/STEP_SPEC.md section 2); same layout as
 * struct SslSimParams in include/sslsim_batch.h (64 bytes) */
typedef struct {
  uint32_t flags;            /* bit0: ball spin (rolling) model */
  int32_t solver_iterations; /* Bullet default 10 */
  float kick_reach, kick_halfwidth, kick_max_height;
  float wheel_kp, wheel_fmax, robot_yaw_inertia;
  float dribble_kd, dribble_cd, dribble_amax, dribble_reach;
  float mu_roll;
  float restitution_threshold; /* Bullet m_restitutionVelocityThreshold: 0 = Bullet 2.82 (none); 0.2 = Bullet >= 2.87 */
  float warmstart_factor;      /* Bullet m_warmstartingFactor, default 0.85; 0 = no warm starting (spec v2) */
  float reserved[1];
} orc_params;

/* per robot command, 32 bytes; schema: reference protos/grSim_Commands.proto:1-14 */
typedef struct {
  float wheel[4];  /* rad/s if ORC_CMD_WHEELS else {veltangent, velnormal, velangular, -} */
  float kickspeedx, kickspeedz;
  uint32_t flags;
  uint32_t pad;
} orc_cmd;

#define ORC_CMD_DRIVE 1u   /* robot is wheel-driven (else passive rigid body, reference behaviour) */
#define ORC_CMD_WHEELS 2u  /* grSim 'wheelsspeed' */
#define ORC_CMD_SPINNER 4u /* grSim 'spinner' */

#define ORC_F_BALL_SPIN 1u

/* aux words (8 x u32 = 32 bytes per world) */
enum {
  ORC_AUX_STATUS = 0, /* [7:0] last touching robot (0xFF none); bit8/9/10 prev in +x goal / -x goal / out;
                         bits12..15 events of the last step: goal_blue, goal_yellow, out, kicked;
                         bit16 row overflow (sticky); bit17 non-finite state (sticky);
                         bits18..19 ball activation state (0 active, 1 wants deactivation, 2 asleep);
                         bits20..31 consecutive substeps with the ball below the sleeping thresholds */
  ORC_AUX_PEAK2 = 1,  /* f32 bits: peak |v_ball|^2 since last kick */
  ORC_AUX_TOUCH = 2,  /* robots that touched the ball during the last step (bit i = robot i) */
  ORC_AUX_RESTARTS = 3, /* restarts placed by orc_referee */
  ORC_AUX_GOALS = 4,  /* goals_blue | goals_yellow << 16 */
  ORC_AUX_OUTS = 5,   /* ball_out count | kicks << 16 */
  ORC_AUX_TOUCHES = 6, /* sum over steps of popcount(touch mask) */
  ORC_AUX_CONTACTS = 7 /* total active contact rows over all substeps */
};

void orc_params_default(orc_params *p);

/* Advance one world by n_steps steps of `substeps` substeps of length h.
 * pos/vel: n_robots + 1 entries, ball last.  cmd may be NULL (reference mode:
 * every robot passive).  aux: 8 words. */
void orc_step(const orc_field *f, const orc_params *p, int n_robots,
              orc_vec4 *pos, orc_vec4 *vel, uint32_t *aux,
              const orc_cmd *cmd, int n_steps, int substeps, float h);

/* The same step with the world's warm table (STEP_SPEC 4.5a: Bullet's warm starting, every contact row that had a row
 * with the same bodies / static solid in the previous substep starts at p->warmstart_factor x the impulses it ended
 * that substep with).  `warm` is the world's table, orc_warm_words(n_robots) u32 words, layout documented in
 * ssl_oracle.c and include/sslsim_batch.h (identical to the device's); all zero = empty; it is read and rewritten by
 * every substep.  warm == NULL (orc_step) or warmstart_factor == 0: every row starts from zero impulse (spec v2) and the
 * table is neither read nor written. */
int orc_warm_words(int n_robots);
void orc_step_warm(const orc_field *f, const orc_params *p, int n_robots, orc_vec4 *pos, orc_vec4 *vel,
                   uint32_t *aux, const orc_cmd *cmd, int n_steps, int substeps, float h, uint32_t *warm);

/* deterministic initial state: mode 0 = reference-style uniform
 * placement dropped from 0.5 m (reference src/sslsim.cpp:34-42, :446);
 * mode 1 = crowded +x penalty area, resting. */
void orc_init_world(const orc_field *f, int n_robots, uint64_t seed,
                    uint64_t world_id, int mode, orc_vec4 *pos, orc_vec4 *vel,
                    uint32_t *aux);

/* synthetic command stream: kind 0 = wheel
 * speeds U[-40,40] rad/s; kind 1 = kind 0 + spinner + kicks. */
void orc_gen_commands(int n_robots, uint64_t seed, uint64_t world_id,
                      uint64_t period, int kind, orc_cmd *cmd);

/* detection records (reference src/serialize.cpp:25-65): out gets
 * 6 floats per ball {conf,x,y,z,px,py} then 7 per robot {conf,id,x,y,yaw,px,py}
 * in world robot order.  ids: robot ids. Returns number of floats. */
int orc_gather(int n_robots, const orc_vec4 *pos, const int *ids, float *out);
float orc_readout_yaw(float yaw); /* Bullet's matrix -> quaternion -> sign(axis.z) * angle readout, range (-2pi/3, 4pi/3) */

/* referee restart placement (STEP_SPEC 7, builder-defined): if the last step raised a goal event the ball goes to
 * the centre spot, if it raised ball-out the ball goes back 0.1 m inside the line it crossed; placed at rest
 * on the ground (z = ball radius, zero velocity and spin), in-goal/out edge state cleared, aux[3]++.
 * Returns 1 if the ball was re-placed. */
int orc_referee(const orc_field *f, int n_robots, orc_vec4 *pos, orc_vec4 *vel, uint32_t *aux);

/* number of substeps of length h after which Bullet's m_deactivationTime (binary32 accumulation) first exceeds
 * gDeactivationTime = 2 s; 4096 if that is beyond the 12-bit counter */
int orc_sleep_substeps(float h);

/* splitmix64-keyed uniform in [0,1) with 24 bits */
float orc_u01(uint64_t seed, uint64_t a, uint64_t b, uint64_t c);

void orc_sincos(float x, float *s, float *c);

/* multi-world convenience for timing / tests: steps W independent worlds laid
 * out world-major (pos[w*(R+1)+b]); threads>1 uses pthreads, one contiguous
 * world range per thread. */
void orc_step_many(const orc_field *f, const orc_params *p, int n_robots,
                   int n_worlds, orc_vec4 *pos, orc_vec4 *vel, uint32_t *aux,
                   const orc_cmd *cmd, int n_steps, int substeps, float h,
                   int threads, uint32_t *warm /* [W][orc_warm_words] or NULL = cold rows */);

#ifdef __cplusplus
}
#endif
#endif
