/*
 * legacy_api.cu -- the reference's single-world C API (include/sslsim.h) on top
 * of the batched device path.  A `struct World` is either a private batch of one
 * world (new_world) or a view of member i of a WorldBatch
 * (world_batch_get_world).  Getters read a host mirror that is refreshed from
 * the device after a step; setters write through to the device.  The step is
 * always the CUDA kernel: there is no host implementation.
 *
 * Reference bodies being replaced: src/sslsim.cpp:298-473.
 */
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include "batch_internal.h"

namespace {

constexpr float ROBOT_R = 0.09f;  /* reference src/sslsim.cpp:19 */
constexpr float DROP_Z = 0.5f;    /* :22 */
constexpr float TWO_PI = 6.2831855f;
constexpr uint64_t LEGACY_SEED = 0x5351; /* SURVEY 8d config 1 */

uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
float u01(uint64_t seed, uint64_t a, uint64_t b, uint64_t c) {
  uint64_t h = mix64(seed);
  h = mix64(h ^ a);
  h = mix64(h ^ b);
  h = mix64(h ^ c);
  return (float)(uint32_t)(h >> 40) * 5.9604645e-08f;
}

WorldBatch *B(const World *w) { return w->batch; }

/* device -> host mirror of this world's slice.  Ordered after everything queued on the batch, chunk streams
 * included.  On failure (sticky CUDA error) the mirror is zeroed, so the reference-style getters -- which have no
 * way to report an error -- return zeros instead of stale state; world_batch_last_error() tells why. */
bool pull(World *w) {
  if (w->mirror_valid) return true;
  WorldBatch *b = B(w);
  const int N = b->R + 1;
  if (b->last_error == SSLSIM_E_CUDA || sslsim_join_chunks(b) != SSLSIM_OK) {
    memset(w->pos, 0, sizeof w->pos); memset(w->vel, 0, sizeof w->vel); memset(w->aux, 0, sizeof w->aux);
    return false;
  }
  cudaSetDevice(b->device);
  const size_t off = (size_t)w->index * N;
  bool ok = sslsim_check(b, cudaMemcpyAsync(w->pos, b->d_pos + off, N * sizeof(float4), cudaMemcpyDeviceToHost, b->stream), "pull pos") &&
            sslsim_check(b, cudaMemcpyAsync(w->vel, b->d_vel + off, N * sizeof(float4), cudaMemcpyDeviceToHost, b->stream), "pull vel") &&
            sslsim_check(b, cudaMemcpyAsync(w->aux, b->d_aux + (size_t)w->index * 8, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, b->stream), "pull aux") &&
            sslsim_check(b, cudaStreamSynchronize(b->stream), "pull sync");
  w->mirror_valid = ok;
  if (!ok) { memset(w->pos, 0, sizeof w->pos); memset(w->vel, 0, sizeof w->vel); memset(w->aux, 0, sizeof w->aux); }
  return ok;
}

/* host mirror -> device */
bool push(World *w) {
  WorldBatch *b = B(w);
  if (b->last_error == SSLSIM_E_CUDA || sslsim_join_chunks(b) != SSLSIM_OK) return false;
  cudaSetDevice(b->device);
  const int N = b->R + 1;
  const size_t off = (size_t)w->index * N;
  return sslsim_check(b, cudaMemcpyAsync(b->d_pos + off, w->pos, N * sizeof(float4), cudaMemcpyHostToDevice, b->stream), "push pos") &&
         sslsim_check(b, cudaMemcpyAsync(b->d_vel + off, w->vel, N * sizeof(float4), cudaMemcpyHostToDevice, b->stream), "push vel") &&
         sslsim_check(b, cudaMemcpyAsync(b->d_aux + (size_t)w->index * 8, w->aux, 8 * sizeof(uint32_t), cudaMemcpyHostToDevice, b->stream), "push aux") &&
         sslsim_check(b, sslsim_clear_warm(b, w->index, 1), "push warm") && /* a setter moved a body: its contact points are gone */
         sslsim_check(b, cudaStreamSynchronize(b->stream), "push sync");
}

World *make_view(WorldBatch *b, int index, bool owns) {
  World *w = new (std::nothrow) World();
  if (!w) return nullptr;
  w->batch = b; w->index = index; w->owns_batch = owns;
  w->n_robots = b->R;
  for (int i = 0; i < b->R; ++i) w->robots[i] = Robot{w, i, b->robot_ids[i], (enum Team)b->robot_teams[i]};
  w->n_balls = 1;
  w->balls[0] = Ball{w, 0};
  memset(w->pos, 0, sizeof w->pos); memset(w->vel, 0, sizeof w->vel); memset(w->aux, 0, sizeof w->aux);
  return w;
}

} // namespace

extern "C" {

/* ---- world ---------------------------------------------------------------------- */

/* reference src/sslsim.cpp:301 -> World ctor :238-252: statics + one ball dropped from 0.5 m at the origin */
struct World *new_world(const struct FieldGeometry *field) {
  WorldBatch *b = sslsim_batch_create(field, 1, 0, SSLSIM_MAX_BODIES, 0, LEGACY_SEED, 0, 0, nullptr, true);
  if (!b) return nullptr;
  World *w = make_view(b, 0, true);
  if (!w) { delete_world_batch(b); return nullptr; }
  return w;
}

/* reference src/sslsim.cpp:304 (copy-ctor :254-273 never re-adds the bodies upstream); here: a deep copy */
struct World *clone_world(const struct World *world) {
  if (!world) return nullptr;
  World *src = const_cast<World *>(world);
  if (!pull(src)) return nullptr;
  WorldBatch *sb = B(src);
  WorldBatch *b = sslsim_batch_create(sb->field_ptr, 1, src->n_robots, SSLSIM_MAX_BODIES, sb->device, sb->seed,
                                      sb->world0 + (uint64_t)src->index, 0, &sb->params, false);
  if (!b) return nullptr;
  b->robot_ids = sb->robot_ids; b->robot_teams = sb->robot_teams;
  b->frame = sb->frame; b->timestamp = sb->timestamp;
  World *w = make_view(b, 0, true);
  if (!w) { delete_world_batch(b); return nullptr; }
  for (int i = 0; i < src->n_robots; ++i) { w->robots[i].id = src->robots[i].id; w->robots[i].team = src->robots[i].team; }
  memcpy(w->pos, src->pos, sizeof w->pos); memcpy(w->vel, src->vel, sizeof w->vel); memcpy(w->aux, src->aux, sizeof w->aux);
  w->local_time = src->local_time;
  if (src->n_robots > 0)
    cudaMemcpy(b->d_robot_ids, b->robot_ids.data(), src->n_robots * sizeof(int), cudaMemcpyHostToDevice);
  if (sb->cmd_active) { /* carry the current commands of that world over */
    RobotCommand tmp[SSLSIM_MAX_BODIES];
    cudaSetDevice(sb->device);
    cudaStreamSynchronize(sb->stream);
    if (src->n_robots > 0 &&
        cudaMemcpy(tmp, sb->cmd_active + (size_t)src->index * sb->R, src->n_robots * sizeof(RobotCommand),
                   cudaMemcpyDeviceToHost) == cudaSuccess)
      world_batch_set_commands(b, tmp);
  }
  if (!push(w)) { delete_world_batch(b); delete w; return nullptr; }
  /* the copy continues exactly where the source is: its warm table comes along (push emptied the copy's) */
  if (cudaMemcpy(b->d_warm, sb->d_warm + (size_t)src->index * sslsim_warm_words(sb),
                 (size_t)sslsim_warm_words(sb) * sizeof(uint32_t), cudaMemcpyDeviceToDevice) != cudaSuccess) {
    delete_world_batch(b); delete w; return nullptr;
  }
  w->mirror_valid = true;
  return w;
}

void delete_world(struct World *world) {
  if (!world) return;
  if (world->owns_batch) { delete_world_batch(world->batch); delete world; }
  /* views are owned by their batch */
}

void world_step(struct World *world) {
  world_step_delta(world, (Float)(1.0 / 60), 2, (Float)(1.0 / 60 / 2)); /* reference src/sslsim.cpp:309 */
}

/* reference src/sslsim.cpp:312-326; substep scheduling = btDiscreteDynamicsWorld::stepSimulation
 * (accumulate, n = int(local/fixed), clamp to max_substeps; max_substeps 0 = one variable step) */
void world_step_delta(struct World *world, Float time_step, int max_substeps, Float fixed_time_step) {
  if (!world) return;
  WorldBatch *b = B(world);
  int n = 0;
  float h = fixed_time_step;
  if (max_substeps > 0) {
    world->local_time += time_step;
    if (world->local_time >= fixed_time_step) {
      n = (int)(world->local_time / fixed_time_step);
      world->local_time -= (float)n * fixed_time_step;
    }
    if (n > max_substeps) n = max_substeps;
  } else {
    world->local_time = time_step;
    h = time_step;
    n = fabsf(time_step) < 1.1920929e-07f ? 0 : 1;
  }
  if (!world->owns_batch) {
    /* a view of a batch member shares the batch's clock and launch: stepping one member alone would tear the
     * lockstep, stepping all of them behind the caller's back is worse.  Refused: no state changes, the batch
     * reports SSLSIM_E_ARG (not sticky) through world_batch_last_error(). */
    world->local_time = 0.0f;
    b->last_error = SSLSIM_E_ARG;
    snprintf(b->errstr, sizeof b->errstr, "world_step on a batch member view is not supported; use world_batch_step");
    fprintf(stderr, "sslsim_b200: %s\n", b->errstr);
    return;
  }
  sslsim_batch_step(b, 1, n, h, time_step);
  world->mirror_valid = false;
}

const struct FieldGeometry *world_get_field(const struct World *world) { return B(world)->field_ptr; }
int world_ball_count(const struct World *world) { return world->n_balls; }
const struct Ball *world_get_ball(const struct World *world, int index) { return &world->balls[index]; }
const struct Ball *world_get_game_ball(const struct World *world) { return &world->balls[0]; }
struct Ball *world_get_mut_ball(struct World *world, int index) { return &world->balls[index]; }
struct Ball *world_get_mut_game_ball(struct World *world) { return &world->balls[0]; }
int world_robot_count(const struct World *world) { return world->n_robots; }
const struct Robot *world_get_robot(const struct World *world, int index) { return &world->robots[index]; }
struct Robot *world_get_mut_robot(struct World *world, int index) { return &world->robots[index]; }

/* reference src/sslsim.cpp:358-365 + random_robot_pos2 :34-42 (uniform pose, dropped from 0.5 m).
 * The reference draws from std::random_device; here the draw is the keyed counter RNG of the
 * batch initialiser, so new_world + 2R x world_add_robot == world 0 of new_world_batch. */
void world_add_robot(struct World *world, int id, enum Team team) {
  if (!world) return;
  WorldBatch *b = B(world);
  if (!world->owns_batch) {
    fprintf(stderr, "sslsim_b200: world_add_robot: the roster of a batch member is fixed\n");
    return;
  }
  if (world->n_robots >= SSLSIM_MAX_ROBOTS) {
    fprintf(stderr, "sslsim_b200: world_add_robot: at most %d robots per world on the device path\n", SSLSIM_MAX_ROBOTS);
    return;
  }
  if (!pull(world)) return;
  const int r = world->n_robots;
  /* the ball stays the last body */
  memcpy(world->pos[r + 1], world->pos[r], sizeof world->pos[r]);
  memcpy(world->vel[r + 1], world->vel[r], sizeof world->vel[r]);
  const float wd = b->field.field_length / 2 - ROBOT_R, hg = b->field.field_width / 2 - ROBOT_R;
  const uint64_t wid = b->world0 + (uint64_t)world->index;
  const float q0 = u01(b->seed, wid, (uint64_t)r, 0), q1 = u01(b->seed, wid, (uint64_t)r, 1), q2 = u01(b->seed, wid, (uint64_t)r, 2);
  world->pos[r][0] = -wd + q0 * (2.0f * wd);
  world->pos[r][1] = -hg + q1 * (2.0f * hg);
  world->pos[r][2] = DROP_Z;
  world->pos[r][3] = q2 * TWO_PI;
  memset(world->vel[r], 0, sizeof world->vel[r]);
  world->robots[r] = Robot{world, r, id, team};
  world->n_robots = r + 1;
  b->R = r + 1;
  b->robot_ids.push_back(id);
  b->robot_teams.push_back((int)team);
  cudaSetDevice(b->device);
  sslsim_check(b, cudaMemcpyAsync(b->d_robot_ids, b->robot_ids.data(), b->R * sizeof(int), cudaMemcpyHostToDevice, b->stream), "ids");
  /* a passive robot joins: any installed command buffer no longer matches the roster */
  if (b->d_cmd) { cudaStreamSynchronize(b->stream); cudaFree(b->d_cmd); b->d_cmd = nullptr; }
  b->cmd_active = nullptr;
  push(world);
}

struct WorldBatch *world_get_batch(struct World *world, int *index) {
  if (!world) return nullptr;
  if (index) *index = world->index;
  return B(world);
}

unsigned int world_get_frame_number(const struct World *world) { return B(world)->frame; }
double world_get_timestamp(const struct World *world) { return (double)B(world)->timestamp; }
struct btDiscreteDynamicsWorld *world_bt_dynamics(struct World *world) { (void)world; return nullptr; }

/* ---- ball ------------------------------------------------------------------------- */
#define WB(ball) ((ball)->world)
#define BI(ball) ((ball)->world->n_robots) /* body index of the (single) ball */

struct Vec3 ball_get_vec(const struct Ball *ball) { /* reference src/sslsim.cpp:378-382 */
  World *w = WB(ball);
  pull(w);
  const float *p = w->pos[BI(ball)];
  return Vec3{p[0], p[1], p[2]};
}

void ball_set_vec(struct Ball *ball, const struct Vec2 vec) { /* reference :384-388: re-drop, velocity kept */
  World *w = WB(ball);
  if (!pull(w)) return;
  float *p = w->pos[BI(ball)];
  p[0] = vec.x; p[1] = vec.y; p[2] = DROP_Z;
  w->aux[SSLSIM_AUX_STATUS] &= SSLSIM_STATUS_KEEP_ON_WAKE; /* :387 body.activate(true) */
  push(w);
}

struct Pos3 ball_get_pos(const struct Ball *ball) {
  World *w = WB(ball);
  pull(w);
  const float *p = w->pos[BI(ball)];
  return Pos3{p[0], p[1], p[2], 0.0f, 0.0f, 0.0f};
}

void ball_set_pos(struct Ball *ball, const struct Pos3 pos) {
  World *w = WB(ball);
  if (!pull(w)) return;
  float *p = w->pos[BI(ball)];
  p[0] = pos.x; p[1] = pos.y; p[2] = pos.z;
  w->aux[SSLSIM_AUX_STATUS] &= SSLSIM_STATUS_KEEP_ON_WAKE;
  push(w);
}

struct Pos3 ball_get_vel(const struct Ball *ball) {
  World *w = WB(ball);
  pull(w);
  const float *p = w->pos[BI(ball)], *v = w->vel[BI(ball)];
  return Pos3{v[0], v[1], v[2], p[3], v[3], 0.0f}; /* spin (omega_x, omega_y) lives in pos.w / vel.w */
}

void ball_set_vel(struct Ball *ball, const struct Pos3 vel) {
  World *w = WB(ball);
  if (!pull(w)) return;
  float *p = w->pos[BI(ball)], *v = w->vel[BI(ball)];
  v[0] = vel.x; v[1] = vel.y; v[2] = vel.z; p[3] = vel.wx; v[3] = vel.wy;
  w->aux[SSLSIM_AUX_STATUS] &= SSLSIM_STATUS_KEEP_ON_WAKE;
  push(w);
}

int ball_is_touching_robot(const struct Ball *ball, const struct Robot *robot) {
  World *w = WB(ball);
  pull(w);
  if (!robot || robot->world != w) return 0;
  return (w->aux[SSLSIM_AUX_TOUCH] >> robot->index) & 1u;
}

struct Robot *ball_last_touching_robot(struct Ball *ball) {
  World *w = WB(ball);
  pull(w);
  const unsigned last = w->aux[SSLSIM_AUX_STATUS] & 0xFFu;
  return last < (unsigned)w->n_robots ? &w->robots[last] : nullptr;
}

Float ball_get_speed2(const struct Ball *ball) {
  World *w = WB(ball);
  pull(w);
  const float *v = w->vel[BI(ball)];
  return v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
}

Float ball_get_peak_speed2_from_last_kick(const struct Ball *ball) {
  World *w = WB(ball);
  pull(w);
  float f;
  memcpy(&f, &w->aux[SSLSIM_AUX_PEAK2], 4);
  return f;
}

struct btRigidBody *ball_bt_rigid_body(struct Ball *ball) { (void)ball; return nullptr; }

/* ---- robot ------------------------------------------------------------------------ */
int get_id(const struct Robot *robot) { return robot->id; }
enum Team get_team(const struct Robot *robot) { return robot->team; }

struct Pos2 robot_get_pos(const struct Robot *robot) { /* reference src/sslsim.cpp:437-443 */
  World *w = robot->world;
  pull(w);
  const float *p = w->pos[robot->index];
  return Pos2{p[0], p[1], sslsim_readout_yaw(p[3])};
}

void robot_set_pos(struct Robot *robot, const struct Pos2 pos) { /* reference :445-450: re-drop, velocity kept */
  World *w = robot->world;
  if (!pull(w)) return;
  float *p = w->pos[robot->index];
  p[0] = pos.x; p[1] = pos.y; p[2] = DROP_Z; p[3] = pos.w;
  push(w);
}

struct Pos2 robot_get_vel(const struct Robot *robot) {
  World *w = robot->world;
  pull(w);
  const float *v = w->vel[robot->index];
  return Pos2{v[0], v[1], v[3]};
}

void robot_set_vel(struct Robot *robot, const struct Pos2 vel) {
  World *w = robot->world;
  if (!pull(w)) return;
  float *v = w->vel[robot->index];
  v[0] = vel.x; v[1] = vel.y; v[3] = vel.w;
  push(w);
}

/* geometric test on the current poses: the two chassis cylinders within the contact margin */
int robot_is_touching_robot(const struct Robot *robot, const struct Robot *tobor) {
  if (!robot || !tobor || robot->world != tobor->world || robot == tobor) return 0;
  World *w = robot->world;
  pull(w);
  const float *a = w->pos[robot->index], *c = w->pos[tobor->index];
  const float dx = a[0] - c[0], dy = a[1] - c[1];
  const float lim = 2.0f * ROBOT_R + 0.02f;
  if (!(dx * dx + dy * dy <= lim * lim)) return 0;
  const float s_up = (a[2] - 0.02f) - (c[2] + 0.125f), s_dn = (c[2] - 0.02f) - (a[2] + 0.125f);
  return (s_up > s_dn ? s_up : s_dn) <= 0.02f;
}

Float robot_get_speed2(const struct Robot *robot) {
  World *w = robot->world;
  pull(w);
  const float *v = w->vel[robot->index];
  return v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
}

/* ---- batch member views ---------------------------------------------------------------- */
struct World *world_batch_get_world(struct WorldBatch *b, int index) {
  if (!b || index < 0 || index >= b->W) return nullptr;
  for (World *v : b->views)
    if (v->index == index) return v;
  World *w = make_view(b, index, false);
  if (w) b->views.push_back(w);
  return w;
}

} /* extern "C" */
