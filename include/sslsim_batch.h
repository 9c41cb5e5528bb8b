/*
 * sslsim_batch.h -- batched extension of the ssl-sim C API: thousands of
 * independent SSL worlds stepped in lockstep by hand-written sm_100a kernels.
 *
 * Additive to sslsim.h (no reference symbol changes).  Plain C ABI: pointers,
 * sizes and PODs only.  Each entry point names the reference interface it
 * stands in for (paths relative to the reference tree).
 *
 * Conventions
 *  - One CUDA stream per batch.  world_batch_step / set_commands / replace_*
 *    are asynchronous on it; read/gather/reduce calls synchronise.
 *    Calls on one batch are not thread-safe; different batches are independent.
 *  - New calls return int status: 0 ok, <0 SSLSIM_E_*; CUDA errors are sticky
 *    per batch (world_batch_last_error / world_batch_error_string).
 *  - World-major host layouts: state[w][body][4] floats with the ball as the
 *    last body, aux[w][8] words, commands[w][robot].
 *  - There is no CPU fallback: without a CUDA device new_world_batch fails.
 */
#ifndef SSLSIM_B200_BATCH_H
#define SSLSIM_B200_BATCH_H
#include <stddef.h>
#include <stdint.h>
#include "sslsim.h"
#ifdef __cplusplus
extern "C" {
#endif

struct WorldBatch;

#define SSLSIM_MAX_ROBOTS 31 /* one lane per body: 31 robots + ball per warp */

enum {
  SSLSIM_OK = 0,
  SSLSIM_E_ARG = -1,    /* bad argument */
  SSLSIM_E_CUDA = -2,   /* CUDA runtime error (sticky) */
  SSLSIM_E_NOMEM = -3,
  SSLSIM_E_SPACE = -4   /* output buffer too small */
};

/* Builder-defined tunables of the parts of the step that have no upstream
 * implementation (docs/STEP_SPEC.md section 2).  64 bytes. */
struct SslSimParams {
  uint32_t flags;            /* SSLSIM_F_* */
  int32_t solver_iterations; /* Bullet default 10 (btContactSolverInfo) */
  float kick_reach, kick_halfwidth, kick_max_height;
  float wheel_kp, wheel_fmax, robot_yaw_inertia;
  float dribble_kd, dribble_cd, dribble_amax, dribble_reach;
  float mu_roll;
  float restitution_threshold; /* Bullet's m_restitutionVelocityThreshold: 0 = Bullet 2.82 (no threshold, the default
                                  here); 0.2 = btContactSolverInfo's default from Bullet 2.87 on */
  float warmstart_factor;      /* Bullet's btContactSolverInfo::m_warmstartingFactor (SOLVER_USE_WARMSTARTING is in
                                  Bullet's default solver mode): a contact row that had a row in the previous substep
                                  starts at this factor x the impulses it ended with (docs/STEP_SPEC.md 4.5a).  Default
                                  0.85 = Bullet; 0 = every row starts from zero (the cold-start step of spec v2) */
  float reserved[1];
};
#define SSLSIM_F_BALL_SPIN 1u /* ball rolls (spin state + rolling resistance) instead of the reference's pure sliding */
void sslsim_params_default(struct SslSimParams *p);

/* Per-robot command, 32 bytes.  Schema: protos/grSim_Commands.proto:1-14
 * (kickspeedx, kickspeedz, veltangent, velnormal, velangular, spinner,
 * wheelsspeed, wheel1..4); wheel order and mounting: src/sslsim.cpp:26-28,115-122. */
struct RobotCommand {
  float wheel[4]; /* SSLSIM_CMD_WHEELS: wheel1..4 [rad/s]; else {veltangent, velnormal, velangular, unused} */
  float kickspeedx, kickspeedz;
  uint32_t flags; /* SSLSIM_CMD_* */
  uint32_t pad;
};
#define SSLSIM_CMD_DRIVE 1u   /* wheel-driven; 0 = passive rigid body (the reference's behaviour, sslsim.cpp:361) */
#define SSLSIM_CMD_WHEELS 2u  /* grSim 'wheelsspeed' */
#define SSLSIM_CMD_SPINNER 4u /* grSim 'spinner' */

/* Replacement records.  Schema: protos/grSim_Replacement.proto:1-19; effect on
 * the pose as robot_set_pos / ball_set_vec (src/sslsim.cpp:445-450, 384-388):
 * the body is re-dropped from 0.5 m.  Robot velocity is untouched (as
 * upstream); the ball takes (vx, vy, 0). */
struct RobotReplacement { int32_t world; int32_t robot; float x, y, dir; };
struct BallReplacement { int32_t world; float x, y, vx, vy; };

/* The warm table: the fourth state array of a world beside pos / vel / aux (docs/STEP_SPEC.md 4.5a) -- the accumulated
 * normal and friction impulse every contact row of the last substep ended with, keyed by what the row is between (Bullet
 * keeps them in its persistent manifolds).  u32 words per world, N = n_robots + 1, cap = 32 for N <= 16, else 64:
 *   [0, N)        sig[b]    byte k = 1 + id of the static solid in slot k of body b (0 = empty slot), k = 0..3; slots
 *                           hold the ceiling (id 0) and the six goal boxes (ids 5..10)
 *   [N, 5N)       on[k][b]  normal impulse of that row (f32 bits);   [5N, 9N)  of[k][b]  its friction impulse
 *   [9N, 11N)     gn[b], gf[b]     ground row of body b: normal, friction impulse
 *   [11N, 13N)    wxn[b], wxf[b]   its row against a field wall of the x axis (a body cannot leave one wall of an axis
 *                                  and reach the other within a substep, so the axis is the key)
 *   [13N, 15N)    wyn[b], wyf[b]   the same for the y axis
 *   [15N]         np        number of pair entries
 *   [15N + 1 ..)  pab[cap], pn[cap], pf[cap]   robot-robot (a = i < b = j) and ball-robot (a = ball, b = robot) rows:
 *                           a | b << 8, normal impulse, friction impulse
 * Slot and entry order carry no meaning; an entry with two zero impulses is no entry; all zero = empty.  The step reads
 * and rewrites it; snapshot / restore / fork carry it; every external write of a world's state (world_batch_write_state,
 * replace_*, referee placement, the legacy setters, reset) clears that world's table -- the moved bodies' contact
 * points are gone, and the other rows of that world start their next substep cold, as in spec v2.
 * world_batch_read_warm / write_warm move it explicitly. */
#define SSLSIM_WARM_WORDS(n_robots) (15 * ((n_robots) + 1) + 1 + 3 * (((n_robots) + 1) <= 16 ? 32 : 64))

/* aux words per world (8 x u32) */
enum {
  SSLSIM_AUX_STATUS = 0,  /* [7:0] last touching robot index (0xFF none); bit8/9/10 ball currently in +x goal /
                             -x goal / out; bits 12..15 events of the last step: goal for blue (+x goal),
                             goal for yellow (-x goal), ball out, kicked; bit16 contact-row overflow (sticky);
                             bit17 non-finite state (sticky); bits 18..19 ball activation state (Bullet's
                             ACTIVE_TAG = 0 / WANTS_DEACTIVATION = 1 / ISLAND_SLEEPING = 2; the reference's ball may
                             sleep, src/sslsim.cpp:51-61, its robots may not, :113); bits 20..31 consecutive
                             substeps with the ball below the sleeping thresholds (saturating) */
  SSLSIM_AUX_PEAK2 = 1,   /* f32 bits: peak |v_ball|^2 since the last kick */
  SSLSIM_AUX_TOUCH = 2,   /* robots that touched the ball during the last step (bit i = robot i) */
  SSLSIM_AUX_RESTARTS = 3, /* restarts placed by world_batch_referee */
  SSLSIM_AUX_GOALS = 4,   /* goals_blue | goals_yellow << 16 */
  SSLSIM_AUX_OUTS = 5,    /* ball-out count | kicks << 16 */
  SSLSIM_AUX_TOUCHES = 6, /* sum over steps of popcount(touch mask) */
  SSLSIM_AUX_CONTACTS = 7 /* active contact rows summed over all substeps */
};

/* Episode statistics of one batch (K4), also the NCCL all-gather payload. */
struct BatchStats {
  uint64_t goals_blue, goals_yellow, ball_out, kicks, touches, contacts;
  uint64_t bad_worlds;   /* worlds with the non-finite or overflow flag */
  uint64_t checksum;     /* order-independent sum over worlds of a hash of (world id, state bits) */
};

/* ---- lifecycle: stands in for new_world / world_add_robot x 2R / delete_world
 * (src/world.h:21-23,41; src/sslsim.cpp:238-252,358-365) for n_worlds at once.
 * Roster order is B0,Y0,B1,Y1,... as src/main.c:65-68.  Initial poses come from
 * a counter RNG keyed (seed, first_world_id + w, body) over the ranges of
 * random_robot_pos2 (src/sslsim.cpp:34-42), so shards of a larger batch are
 * reproducible on any GPU. */
struct WorldBatch *new_world_batch(const struct FieldGeometry *field, int n_worlds,
                                   int robots_per_team, int device, uint64_t seed);
/* init_mode 0: reference-style uniform placement dropped from 0.5 m;
 * 1: all bodies resting inside the +x penalty area (crowded).
 * params NULL = defaults.  n_robots is the total per world (any split; roster
 * alternates blue/yellow). */
struct WorldBatch *new_world_batch_ex(const struct FieldGeometry *field, int n_worlds,
                                      int n_robots, int device, uint64_t seed,
                                      uint64_t first_world_id, int init_mode,
                                      const struct SslSimParams *params);
void delete_world_batch(struct WorldBatch *batch);
int world_batch_reset(struct WorldBatch *batch, uint64_t seed, int init_mode); /* re-run the init kernel */

int world_batch_world_count(const struct WorldBatch *batch);
int world_batch_robot_count(const struct WorldBatch *batch); /* robots per world */
int world_batch_device(const struct WorldBatch *batch);
unsigned int world_batch_frame_number(const struct WorldBatch *batch); /* src/sslsim.cpp:325 */
double world_batch_timestamp(const struct WorldBatch *batch);          /* src/sslsim.cpp:324 (fp32 accumulator) */

/* ---- the step: world_step / world_step_delta (src/world.h:25-27,
 * src/sslsim.cpp:307-326) for every world, n_steps times in one launch with
 * the current commands held fixed.  world_batch_step = n_steps x (2 substeps
 * of 1/120 s). */
int world_batch_step(struct WorldBatch *batch, int n_steps);
int world_batch_step_delta(struct WorldBatch *batch, int n_steps, int substeps, float fixed_time_step);
/* A command sequence in one launch: n_periods blocks of [n_worlds][n_robots] records (device / host memory),
 * each held for steps_per_period world-steps.  Same result as n_periods x (set_commands + world_batch_step),
 * but every world runs through all periods without a batch-wide barrier in between, so worlds that are
 * expensive in one period do not stall the others (open-loop rollouts, pre-sampled action sequences).
 * The last block stays installed as the current commands. */
int world_batch_step_sequence(struct WorldBatch *batch, const struct RobotCommand *dev_cmds, int n_periods,
                              int steps_per_period);
int world_batch_step_sequence_host(struct WorldBatch *batch, const struct RobotCommand *host_cmds, int n_periods,
                                   int steps_per_period);
int world_batch_sync(struct WorldBatch *batch);

/* ---- pipelined closed loop.  The reference's caller is a loop of "read commands, world_step, send state"
 * (src/main.c:79-96); with thousands of worlds the batch-wide barrier of that loop (every world waits for the
 * slowest one, and for the copies) is what limits it.  world_batch_set_chunks splits the batch into n contiguous
 * ranges of worlds, each with its own CUDA stream.  world_batch_chunk_step queues on the chunk's stream: the
 * H2D copy of the chunk's commands (host_cmds = [chunk worlds][n_robots] records, pinned memory for a truly
 * asynchronous copy; NULL keeps what the batch's command buffer holds for the chunk -- passive robots until
 * the chunk was given commands), n_steps world-steps of the chunk's worlds, and the D2H copy
 * of their aux words into host_aux ([chunk worlds][8], may be NULL) -- and returns at once.
 * world_batch_chunk_wait blocks until that work is done and host_aux is valid.  A closed-loop caller cycles
 * over the chunks (wait c, decide, submit c): copies, launches and the tail of each launch overlap with the
 * other chunks' work.  Results are bit-identical to whole-batch stepping.  Every other call on the batch is
 * ordered after all queued chunk work; n_chunks 0 or 1 switches the pipeline off. */
int world_batch_set_chunks(struct WorldBatch *batch, int n_chunks);
int world_batch_chunk_count(const struct WorldBatch *batch);
int world_batch_chunk_range(const struct WorldBatch *batch, int chunk, int *first_world, int *n_worlds);
int world_batch_chunk_step(struct WorldBatch *batch, int chunk, const struct RobotCommand *host_cmds, int n_steps,
                           uint32_t *host_aux);
/* the same with the chunk's full state copied back as well (the reference's loop reads every pose after every step,
 * src/main.c:80-85): host_pos / host_vel [chunk worlds][n_robots + 1][4] floats, NULL to skip */
int world_batch_chunk_step_ex(struct WorldBatch *batch, int chunk, const struct RobotCommand *host_cmds, int n_steps,
                              uint32_t *host_aux, float *host_pos, float *host_vel);
/* a command sequence (world_batch_step_sequence's layout: n_periods blocks of the whole batch's [n_worlds][n_robots]
 * records, device memory) for one chunk, on the chunk's stream; chunks submitted back to back overlap, which fills
 * the tail of small batches (a launch ends with its slowest world) */
int world_batch_chunk_step_sequence(struct WorldBatch *batch, int chunk, const struct RobotCommand *dev_cmds,
                                    int n_periods, int steps_per_period);
int world_batch_chunk_wait(struct WorldBatch *batch, int chunk);

/* ---- command input (no upstream implementation; schema grSim_Commands.proto).
 * host/device array of n_worlds * n_robots records, world-major.  NULL (or
 * world_batch_clear_commands) = every robot passive, i.e. reference mode. */
int world_batch_set_commands(struct WorldBatch *batch, const struct RobotCommand *host_cmds);
int world_batch_set_commands_device(struct WorldBatch *batch, const struct RobotCommand *dev_cmds);
int world_batch_clear_commands(struct WorldBatch *batch);
/* synthetic command stream generated on the device: kind 0 = wheel
 * speeds U[-40,40] rad/s, kind 1 = kind 0 + spinner + kicks; keyed
 * (seed, world id, robot, period). */
int world_batch_gen_commands(struct WorldBatch *batch, uint64_t period, int kind);
/* same generator, written to a caller-owned device buffer (n_worlds * n_robots records)
 * without installing it: pre-stages command sets that set_commands_device then selects */
int world_batch_gen_commands_into(struct WorldBatch *batch, struct RobotCommand *dev_out, uint64_t period, int kind);

/* ---- replacement input: robot_set_pos / ball_set_vec (src/robot.h:23,
 * src/ball.h:19) as a sparse device scatter (K2). */
int world_batch_replace_robots(struct WorldBatch *batch, const struct RobotReplacement *recs, int n);
int world_batch_replace_balls(struct WorldBatch *batch, const struct BallReplacement *recs, int n);

/* ---- state readout / injection (bulk form of the getters/setters of
 * src/robot.h:22-26, src/ball.h:18-26).  Any pointer may be NULL. */
int world_batch_read_state(struct WorldBatch *batch, float *pos, float *vel, uint32_t *aux);
int world_batch_write_state(struct WorldBatch *batch, const float *pos, const float *vel, const uint32_t *aux);
int world_batch_read_aux(struct WorldBatch *batch, uint32_t *aux);
/* the warm tables, [n_worlds][SSLSIM_WARM_WORDS(n_robots)] words (layout above).  write_state clears them, so a caller
 * that moves a complete state between batches writes the state first and the tables second. */
int world_batch_read_warm(struct WorldBatch *batch, uint32_t *warm);
int world_batch_write_warm(struct WorldBatch *batch, const uint32_t *warm);
int world_batch_warm_words(const struct WorldBatch *batch); /* words per world */
/* selected worlds only: pos/vel [n][n_robots + 1][4] floats, aux [n][8] words (bulk form of the per-world getters) */
int world_batch_read_worlds(struct WorldBatch *batch, const int *world_ids, int n, float *pos, float *vel, uint32_t *aux);
/* zero-copy: device pointers of the live buffers (pos/vel: float4 [W][R+1];
 * aux: u32 [W][8]; cmd: RobotCommand [W][R] or NULL) */
int world_batch_device_state(struct WorldBatch *batch, void **pos, void **vel, void **aux, void **cmd);

/* ---- detection gather (K3): the per-object loops of serialize_world
 * (src/serialize.cpp:25-65) for n observed worlds.  Record per world, floats:
 * ball {confidence, x, y, z, pixel_x, pixel_y} then per robot in world order
 * {confidence, robot_id, x, y, orientation, pixel_x, pixel_y}; 6 + 7*R floats.
 * Returns the number of bytes written or <0. */
int world_batch_gather_detections(struct WorldBatch *batch, const int *world_ids, int n,
                                  void *out, size_t capacity);
size_t world_batch_detection_record_size(const struct WorldBatch *batch);

/* ---- legacy view: member i of the batch as a `struct World *` usable with
 * every sslsim.h call (host-mirrored; getters synchronise).  Owned by the batch. */
struct World *world_batch_get_world(struct WorldBatch *batch, int index);
/* the reverse: the batch a `struct World` lives in (a batch of one for new_world / clone_world results) and, if
 * index is not NULL, its index there -- for what the reference's per-world getters do not expose (a robot's height,
 * the aux words, the warm table: world_batch_read_worlds / read_warm on the result).  Borrowed, never NULL for a live world. */
struct WorldBatch *world_get_batch(struct World *world, int *index);

/* ---- referee restart placement (README goal 4, no upstream implementation).  For every world
 * whose last step raised a goal event the ball is put on the centre spot; after ball-out it is put back 0.1 m
 * inside the line it crossed; at rest on the ground.  The in-goal/out edge state is cleared and
 * aux[SSLSIM_AUX_RESTARTS] counts.  Asynchronous on the batch stream. */
int world_batch_referee(struct WorldBatch *batch);

/* ---- snapshot / fork: what reference clone_world (src/world.h:22, src/sslsim.cpp:254-273,
 * which never re-adds the bodies upstream) is for -- AI look-ahead (README goal 2) -- done batched and on the
 * device.  The state is flat POD, so a snapshot is three device-to-device copies. */
struct WorldSnapshot;
struct WorldSnapshot *world_batch_snapshot(struct WorldBatch *batch);           /* NULL on failure */
int world_batch_restore(struct WorldBatch *batch, const struct WorldSnapshot *snap); /* same shape required */
void world_snapshot_free(struct WorldSnapshot *snap);
/* copy the state (pose, velocity, aux words) of world `src` into the n worlds `dst[]` of the same batch */
int world_batch_fork(struct WorldBatch *batch, int src, const int *dst, int n);

/* ---- episode statistics (K4): per-batch reduction on the device. */
int world_batch_reduce_stats(struct WorldBatch *batch, struct BatchStats *out);

/* ---- stream / timing helpers (CUDA events on the batch's own stream) */
void *world_batch_stream(struct WorldBatch *batch);             /* cudaStream_t */
int world_batch_set_stream(struct WorldBatch *batch, void *cuda_stream); /* run on a caller-owned stream */
int world_batch_timer_start(struct WorldBatch *batch);
int world_batch_timer_stop(struct WorldBatch *batch, float *elapsed_ms); /* synchronises */
/* kernel-only time: enable, then read the sum over step_worlds launches since
 * the last reset (events bracketing each launch); reading synchronises */
int world_batch_enable_kernel_timing(struct WorldBatch *batch, int on);
int world_batch_kernel_time(struct WorldBatch *batch, float *total_ms, int *launches, int reset);
/* tuning: lanes per world (16 or 32; 0 = automatic) */
int world_batch_set_lanes_per_world(struct WorldBatch *batch, int lanes);

/* self-test: the step kernel computes its square roots and divisions with branch-free sequences that are
 * correctly rounded on the operand ranges it guarantees; this counts mismatches against sqrtf and operator/ over
 * n random operands from those ranges (expected: 0 and 0) */
int sslsim_selftest_math(int device, unsigned long long n, unsigned long long seed, unsigned long long *bad_sqrt,
                         unsigned long long *bad_div);

int world_batch_last_error(const struct WorldBatch *batch);
const char *world_batch_error_string(const struct WorldBatch *batch);
const char *sslsim_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif
