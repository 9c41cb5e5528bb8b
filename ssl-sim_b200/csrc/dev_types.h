/*
 * dev_types.h -- internal types shared by the kernels and the host C-ABI layer.
 * Not part of the public boundary (include/ is).
 */
#ifndef SSLSIM_B200_DEV_TYPES_H
#define SSLSIM_B200_DEV_TYPES_H
#include <cuda_runtime.h>
#include <stdint.h>
#include "sslsim_batch.h"

/* Field geometry derived once on the host from struct FieldGeometry
 * (reference src/field.h:18-42, solids built at src/sslsim.cpp:145-203) and
 * passed to every kernel by value, i.e. through the constant bank. */
struct DevField {
  float half_len, half_wid, limit_x, limit_y;
  float goal_half_w, goal_back, goal_height;
  float goal_reach_y; /* goal_width/2 + goal_wall_width: |y| extent of the goal solids */
  float goal_reach_x; /* field_length/2 + goal_depth + goal_wall_width: |x| extent of the goal solids */
  float lo[6][3], hi[6][3]; /* 6 goal boxes: -x goal {side -y, side +y, rear}, +x goal mirrored */
};

struct StepArgs {
  float4 *pos;   /* [W][R+1] {x,y,z,yaw}; ball (last): {x,y,z,omega_x} */
  float4 *vel;   /* [W][R+1] {vx,vy,vz,yaw rate}; ball: {vx,vy,vz,omega_y} */
  uint32_t *aux; /* [W][8] */
  uint32_t *warm; /* [W][SSLSIM_WARM_WORDS(R)]: the warm tables (STEP_SPEC 4.5a), or NULL = every row starts from zero */
  const RobotCommand *cmd; /* [W][R] or NULL */
  int n_worlds, n_robots, n_steps, substeps;
  int steps_per_period;         /* > 0: command sequence, a new [W][R] block every so many steps */
  long long cmd_period_stride;  /* records between the blocks of a sequence */
  float h;
  /* launch invariants derived from h on the host (IEEE fp32, the same values the kernel would compute): they
   * are read as constant-bank operands instead of occupying registers for the whole launch */
  float inv_h, gh, h_over_m, h_over_inertia, ez;
  int sleep_substeps;    /* slow substeps after which the ball wants to sleep (Bullet: > 2 s) */
  float warm_l2max;      /* (contact breaking threshold / h)^2: a contact point whose bodies slid faster did not persist (4.5a) */
  DevField f;            /* by value: constant bank */
  const DevField *f_dev; /* the same in global memory, for the out-of-line general path */
  SslSimParams p;
  unsigned *prof; /* SSLSIM_PROFILE builds: [W][8] per-world counters, else unused */
};

/* status-word bits that survive waking the ball (everything below the activation state / slow-substep count) */
#define SSLSIM_STATUS_KEEP_ON_WAKE 0x3ffffu

/* The orientation robot_get_pos reports (reference src/sslsim.cpp:437-443): sign(axis.z) * btQuaternion::getAngle()
 * of the quaternion btMatrix3x3::getRotation extracts.  For a rotation about z by t in (-pi, pi] that is t itself
 * while cos t > -1/2 (trace branch) and for t >= 2pi/3; for t in (-pi, -2pi/3) the largest-diagonal branch yields a
 * quaternion with negative w and the call returns 2pi - |t|.  Bullet's identity rotation has axis (1,0,0), so yaw 0
 * reads -0.0.  Range (-2pi/3, 4pi/3).  Same arithmetic on host and device (no contraction). */
#if defined(__CUDACC__)
__host__ __device__
#endif
static inline float sslsim_readout_yaw(float yaw) {
  const float k = rintf(yaw * 0.15915494f);
  float t = yaw - k * 6.2831855f;
  if (t < -2.0943952f) t = t + 6.2831855f;
  return t == 0.0f ? -0.0f : t;
}

void sslsim_derive_field(const FieldGeometry *g, DevField *d);
void sslsim_step_invariants(StepArgs *a, const SslSimParams &p); /* fills inv_h ... ez from a->h (step_kernel.cu) */

/* kernel launchers (kernels.cu / step_kernel.cu) */
cudaError_t sslsim_launch_step(const StepArgs &a, int lanes_per_world, cudaStream_t s);
cudaError_t sslsim_launch_init(float4 *pos, float4 *vel, uint32_t *aux, int n_worlds, int n_robots,
                               const FieldGeometry &g, uint64_t seed, uint64_t world0, int mode, cudaStream_t s);
cudaError_t sslsim_launch_gen_commands(RobotCommand *cmd, int n_worlds, int n_robots, uint64_t seed,
                                       uint64_t world0, uint64_t period, int kind, cudaStream_t s);
/* warm / warm_words: the batch's warm tables (include/sslsim_batch.h); kernels that write a world's state from outside
 * the step clear that world's table, fork copies it */
cudaError_t sslsim_launch_replace_robots(float4 *pos, uint32_t *warm, int warm_words, int n_worlds, int n_robots,
                                         const RobotReplacement *recs, int n, cudaStream_t s);
cudaError_t sslsim_launch_replace_balls(float4 *pos, float4 *vel, uint32_t *aux, uint32_t *warm, int warm_words,
                                        int n_worlds, int n_robots, const BallReplacement *recs, int n, cudaStream_t s);
cudaError_t sslsim_launch_gather(const float4 *pos, int n_worlds, int n_robots, const int *world_ids, int n,
                                 const int *robot_ids, float *out, cudaStream_t s);
cudaError_t sslsim_launch_referee(float4 *pos, float4 *vel, uint32_t *aux, uint32_t *warm, int warm_words, int n_worlds,
                                  int n_robots, float field_length, float field_width, cudaStream_t s);
cudaError_t sslsim_launch_fork(float4 *pos, float4 *vel, uint32_t *aux, uint32_t *warm, int warm_words, int n_worlds,
                               int n_robots, int src, const int *dst, int n, cudaStream_t s);
cudaError_t sslsim_selftest_math_launch(unsigned long long n, unsigned long long seed, unsigned long long *d_bad,
                                        cudaStream_t s);
cudaError_t sslsim_launch_reduce(const float4 *pos, const float4 *vel, const uint32_t *aux, int n_worlds,
                                 int n_robots, uint64_t world0, unsigned long long *out8, cudaStream_t s);
#endif
