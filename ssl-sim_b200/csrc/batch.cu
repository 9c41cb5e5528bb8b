/*
 * batch.cu -- host side of the batched C ABI (include/sslsim_batch.h): device
 * buffers, stream, launches.  There is deliberately no CPU implementation of
 * the step here: every compute entry point launches a kernel or fails.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include "batch_internal.h"

/* ---- field constants: values of reference src/field.c:11-63 -------------------- */
extern "C" {
const struct FieldGeometry FIELD_2014_SINGLE = {0.010f, 6.050f, 4.050f, 0.250f, 0.425f, 0.700f, 0.180f, 0.020f,
                                                0.500f, 0.800f, 0.350f, 0.200f, 0.750f, 0.400f, 0.165f};
const struct FieldGeometry FIELD_2014_DOUBLE = {0.010f, 8.090f, 6.050f, 0.250f, 0.425f, 1.000f, 0.180f, 0.020f,
                                                0.500f, 1.000f, 0.500f, 0.200f, 1.000f, 0.400f, 0.165f};
const struct FieldGeometry FIELD_2015 = {0.010f, 9.000f, 6.000f, 0.250f, 0.450f, 1.000f, 0.180f, 0.020f,
                                         0.500f, 1.000f, 0.500f, 0.200f, 1.000f, 0.400f, 0.165f};
/* not in the reference: Division-A field, SSL rulebook values */
const struct FieldGeometry FIELD_DIV_A = {0.010f, 12.0f, 9.0f, 0.30f, 0.40f, 1.8f, 0.18f, 0.02f,
                                          0.5f, 1.0f, 0.5f, 0.2f, 1.0f, 0.4f, 0.16f};
}

bool sslsim_check(WorldBatch *b, cudaError_t e, const char *what) {
  if (e == cudaSuccess) return true;
  if (b) {
    b->last_error = SSLSIM_E_CUDA;
    snprintf(b->errstr, sizeof b->errstr, "%s: %s", what, cudaGetErrorString(e));
  }
  fprintf(stderr, "sslsim_b200: %s: %s\n", what, cudaGetErrorString(e));
  return false;
}

#define CK(call)                                                   \
  do {                                                             \
    if (!sslsim_check(b, (call), #call)) return SSLSIM_E_CUDA;     \
  } while (0)
/* work queued on chunk streams is ordered before anything that follows on the batch stream */
int sslsim_join_chunks(WorldBatch *b) {
  for (WorldBatch::Chunk &c : b->chunks) {
    if (!c.dirty) continue;
    if (!sslsim_check(b, cudaEventRecord(c.done, c.stream), "chunk record") ||
        !sslsim_check(b, cudaStreamWaitEvent(b->stream, c.done, 0), "chunk join")) return SSLSIM_E_CUDA;
    c.dirty = false;
  }
  return SSLSIM_OK;
}
#define ENTER_NOJOIN(b)                           \
  if (!(b)) return SSLSIM_E_ARG;                  \
  if ((b)->last_error == SSLSIM_E_CUDA) return SSLSIM_E_CUDA; \
  cudaSetDevice((b)->device)
#define ENTER(b)                                  \
  ENTER_NOJOIN(b);                                \
  if (!(b)->chunks.empty() && sslsim_join_chunks(b)) return SSLSIM_E_CUDA

static void free_chunks(WorldBatch *b) {
  for (WorldBatch::Chunk &c : b->chunks) {
    if (c.stream) { cudaStreamSynchronize(c.stream); cudaStreamDestroy(c.stream); }
    if (c.done) cudaEventDestroy(c.done);
  }
  b->chunks.clear();
}

static void batch_free(WorldBatch *b) {
  if (!b) return;
  cudaSetDevice(b->device);
  free_chunks(b);
  if (b->ev_main) cudaEventDestroy(b->ev_main);
  if (b->stream) cudaStreamSynchronize(b->stream);
  for (World *w : b->views) delete w;
  cudaFree(b->d_pos); cudaFree(b->d_vel); cudaFree(b->d_aux); cudaFree(b->d_warm); cudaFree(b->d_cmd);
  cudaFree(b->d_robot_ids); cudaFree(b->d_world_ids); cudaFree(b->d_gather); cudaFree(b->d_stats);
  cudaFree(b->d_rep); cudaFree(b->d_field); cudaFree(b->d_prof); cudaFree(b->d_seq);
  if (b->h_gather) cudaFreeHost(b->h_gather);
  for (cudaEvent_t e : b->kev) cudaEventDestroy(e);
  if (b->ev0) cudaEventDestroy(b->ev0);
  if (b->ev1) cudaEventDestroy(b->ev1);
  if (b->stream && b->own_stream) cudaStreamDestroy(b->stream);
  delete b;
}

WorldBatch *sslsim_batch_create(const FieldGeometry *field, int n_worlds, int n_robots, int cap_bodies, int device,
                                uint64_t seed, uint64_t world0, int init_mode, const SslSimParams *params,
                                bool run_init) {
  if (!field || n_worlds <= 0 || n_robots < 0 || n_robots > SSLSIM_MAX_ROBOTS) {
    fprintf(stderr, "sslsim_b200: new_world_batch: bad argument (n_worlds=%d, n_robots=%d, max robots %d)\n",
            n_worlds, n_robots, SSLSIM_MAX_ROBOTS);
    return nullptr;
  }
  if (!(field_limit_x(field) >= 0.5f) || !(field_limit_y(field) >= 0.5f)) {
    fprintf(stderr, "sslsim_b200: new_world_batch: field walls closer than 0.5 m to the centre are not supported\n");
    return nullptr;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) {
    fprintf(stderr, "sslsim_b200: no usable CUDA device (%s); there is no CPU fallback\n",
            e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return nullptr;
  }
  if (device < 0 || device >= ndev) {
    fprintf(stderr, "sslsim_b200: device %d out of range (%d devices)\n", device, ndev);
    return nullptr;
  }
  WorldBatch *b = new (std::nothrow) WorldBatch();
  if (!b) return nullptr;
  b->device = device; b->W = n_worlds; b->R = n_robots;
  b->cap_bodies = cap_bodies > n_robots + 1 ? cap_bodies : n_robots + 1;
  b->seed = seed; b->world0 = world0;
  b->field = *field; b->field_ptr = field;
  sslsim_derive_field(field, &b->dfield);
  if (params) b->params = *params; else sslsim_params_default(&b->params);
  b->robot_ids.resize(n_robots); b->robot_teams.resize(n_robots);
  for (int i = 0; i < n_robots; ++i) { b->robot_ids[i] = i / 2; b->robot_teams[i] = (i & 1) ? TEAM_YELLOW : TEAM_BLUE; }
  const size_t nb = (size_t)n_worlds * b->cap_bodies;
  bool ok = sslsim_check(b, cudaSetDevice(device), "cudaSetDevice") &&
            sslsim_check(b, cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking), "cudaStreamCreate") &&
            sslsim_check(b, cudaEventCreate(&b->ev0), "cudaEventCreate") &&
            sslsim_check(b, cudaEventCreate(&b->ev1), "cudaEventCreate") &&
            sslsim_check(b, cudaMalloc(&b->d_pos, nb * sizeof(float4)), "cudaMalloc(pos)") &&
            sslsim_check(b, cudaMalloc(&b->d_vel, nb * sizeof(float4)), "cudaMalloc(vel)") &&
            sslsim_check(b, cudaMalloc(&b->d_aux, (size_t)n_worlds * 8 * sizeof(uint32_t)), "cudaMalloc(aux)") &&
            sslsim_check(b, cudaMalloc(&b->d_warm, (size_t)n_worlds * SSLSIM_WARM_WORDS(b->cap_bodies - 1) * sizeof(uint32_t)),
                         "cudaMalloc(warm)") &&
            sslsim_check(b, cudaMemsetAsync(b->d_warm, 0, (size_t)n_worlds * SSLSIM_WARM_WORDS(b->cap_bodies - 1) * sizeof(uint32_t),
                                            b->stream), "cudaMemset(warm)") &&
            sslsim_check(b, cudaMalloc(&b->d_stats, 8 * sizeof(unsigned long long)), "cudaMalloc(stats)") &&
            sslsim_check(b, cudaMalloc(&b->d_robot_ids, SSLSIM_MAX_BODIES * sizeof(int)), "cudaMalloc(ids)") &&
            sslsim_check(b, cudaMalloc(&b->d_field, sizeof(DevField)), "cudaMalloc(field)") &&
            sslsim_check(b, cudaMemcpyAsync(b->d_field, &b->dfield, sizeof(DevField), cudaMemcpyHostToDevice, b->stream),
                         "cudaMemcpy(field)");
  if (ok && n_robots > 0)
    ok = sslsim_check(b, cudaMemcpyAsync(b->d_robot_ids, b->robot_ids.data(), n_robots * sizeof(int),
                                         cudaMemcpyHostToDevice, b->stream), "cudaMemcpy(ids)");
  if (ok && run_init)
    ok = sslsim_check(b, sslsim_launch_init(b->d_pos, b->d_vel, b->d_aux, n_worlds, n_robots, b->field, seed, world0,
                                            init_mode, b->stream), "init_worlds");
  if (ok) ok = sslsim_check(b, cudaStreamSynchronize(b->stream), "init sync");
  if (!ok) { batch_free(b); return nullptr; }
  return b;
}

cudaError_t sslsim_clear_warm(WorldBatch *b, int first, int count) {
  const size_t ww = (size_t)sslsim_warm_words(b);
  return cudaMemsetAsync(b->d_warm + (size_t)first * ww, 0, (size_t)count * ww * sizeof(uint32_t), b->stream);
}

static StepArgs step_args(const WorldBatch *b, int first, int count, const RobotCommand *cmd, int n_steps, int substeps,
                          float h, int steps_per_period, long long cmd_period_stride) {
  const size_t N = (size_t)b->R + 1;
  StepArgs a;
  a.pos = b->d_pos + first * N; a.vel = b->d_vel + first * N; a.aux = b->d_aux + (size_t)first * 8;
  a.warm = b->d_warm + (size_t)first * sslsim_warm_words(b);
  a.cmd = cmd ? cmd + (size_t)first * b->R : nullptr;
  a.n_worlds = count; a.n_robots = b->R; a.n_steps = n_steps; a.substeps = substeps; a.h = h;
  a.steps_per_period = steps_per_period; a.cmd_period_stride = cmd_period_stride;
  sslsim_step_invariants(&a, b->params);
  a.f = b->dfield; a.f_dev = b->d_field; a.p = b->params; a.prof = b->d_prof;
  return a;
}

int sslsim_batch_step(WorldBatch *b, int n_steps, int substeps, float h, float time_step_total, int steps_per_period,
                      long long cmd_period_stride) {
  ENTER(b);
  SSLSIM_RANGE("world_batch_step");
  if (n_steps < 0 || substeps < 0) return SSLSIM_E_ARG;
  if (n_steps > 0 && substeps >= 0) {
    const StepArgs a = step_args(b, 0, b->W, b->cmd_active, n_steps, substeps, h, steps_per_period, cmd_period_stride);
    if (b->ktiming) {
      if (b->kev_used + 2 > b->kev.size()) {
        for (int k = 0; k < 2; ++k) {
          cudaEvent_t ev;
          CK(cudaEventCreate(&ev));
          b->kev.push_back(ev);
        }
      }
      CK(cudaEventRecord(b->kev[b->kev_used], b->stream));
    }
    CK(sslsim_launch_step(a, b->lanes, b->stream));
    if (b->ktiming) {
      CK(cudaEventRecord(b->kev[b->kev_used + 1], b->stream));
      b->kev_used += 2;
    }
  }
  for (int s = 0; s < n_steps; ++s) {
    b->timestamp += time_step_total; /* fp32 accumulator, reference src/sslsim.cpp:324 */
    b->frame++;                      /* :325 */
  }
  for (World *w : b->views) w->mirror_valid = false;
  return SSLSIM_OK;
}

extern "C" {

void sslsim_params_default(struct SslSimParams *p) {
  memset(p, 0, sizeof *p);
  p->flags = 0;
  p->solver_iterations = 10;
  p->kick_reach = 0.015f;
  p->kick_halfwidth = 0.04f;
  p->kick_max_height = 0.03f;
  p->wheel_kp = 24.0f;
  p->wheel_fmax = 4.0f;
  p->robot_yaw_inertia = 0.0081f;
  p->dribble_kd = 60.0f;
  p->dribble_cd = 8.0f;
  p->dribble_amax = 12.0f;
  p->dribble_reach = 0.03f;
  p->mu_roll = 0.03f;
  p->restitution_threshold = 0.0f;
  p->warmstart_factor = 0.85f; /* Bullet: btContactSolverInfo::m_warmstartingFactor, SOLVER_USE_WARMSTARTING on by default */
}

struct WorldBatch *new_world_batch(const struct FieldGeometry *field, int n_worlds, int robots_per_team, int device,
                                   uint64_t seed) {
  return sslsim_batch_create(field, n_worlds, 2 * robots_per_team, 0, device, seed, 0, 0, nullptr, true);
}

struct WorldBatch *new_world_batch_ex(const struct FieldGeometry *field, int n_worlds, int n_robots, int device,
                                      uint64_t seed, uint64_t first_world_id, int init_mode,
                                      const struct SslSimParams *params) {
  return sslsim_batch_create(field, n_worlds, n_robots, 0, device, seed, first_world_id, init_mode, params, true);
}

void delete_world_batch(struct WorldBatch *batch) { batch_free(batch); }

int world_batch_reset(struct WorldBatch *b, uint64_t seed, int init_mode) {
  ENTER(b);
  SSLSIM_RANGE("world_batch_reset");
  b->seed = seed;
  CK(sslsim_launch_init(b->d_pos, b->d_vel, b->d_aux, b->W, b->R, b->field, seed, b->world0, init_mode, b->stream));
  CK(sslsim_clear_warm(b, 0, b->W));
  b->frame = 0; b->timestamp = 0.0f;
  for (World *w : b->views) w->mirror_valid = false;
  return SSLSIM_OK;
}

int world_batch_world_count(const struct WorldBatch *b) { return b ? b->W : 0; }
int world_batch_robot_count(const struct WorldBatch *b) { return b ? b->R : 0; }
int world_batch_device(const struct WorldBatch *b) { return b ? b->device : -1; }
unsigned int world_batch_frame_number(const struct WorldBatch *b) { return b ? b->frame : 0; }
double world_batch_timestamp(const struct WorldBatch *b) { return b ? (double)b->timestamp : 0.0; }

int world_batch_step(struct WorldBatch *b, int n_steps) {
  /* reference src/sslsim.cpp:309: world_step_delta(world, 1.0 / 60, 2, 1.0 / 60 / 2) with Float = float */
  return sslsim_batch_step(b, n_steps, 2, (float)(1.0 / 60 / 2), (float)(1.0 / 60));
}

int world_batch_step_delta(struct WorldBatch *b, int n_steps, int substeps, float fixed_time_step) {
  return sslsim_batch_step(b, n_steps, substeps, fixed_time_step, fixed_time_step * (float)substeps);
}

int world_batch_step_sequence(struct WorldBatch *b, const struct RobotCommand *dev_cmds, int n_periods,
                              int steps_per_period) {
  ENTER(b);
  SSLSIM_RANGE("world_batch_step_sequence");
  if (!dev_cmds || n_periods <= 0 || steps_per_period <= 0) return SSLSIM_E_ARG;
  const long long stride = (long long)b->W * b->R;
  b->cmd_active = dev_cmds;
  const int rc = sslsim_batch_step(b, n_periods * steps_per_period, 2, (float)(1.0 / 60 / 2), (float)(1.0 / 60),
                                   steps_per_period, stride);
  b->cmd_active = dev_cmds + (size_t)(n_periods - 1) * stride; /* the last block stays installed */
  return rc;
}

int world_batch_step_sequence_host(struct WorldBatch *b, const struct RobotCommand *host_cmds, int n_periods,
                                   int steps_per_period) {
  ENTER(b);
  SSLSIM_RANGE("world_batch_step_sequence_host");
  if (!host_cmds || n_periods <= 0 || steps_per_period <= 0) return SSLSIM_E_ARG;
  const size_t n = (size_t)n_periods * b->W * b->R;
  if (n > b->seq_cap) {
    CK(cudaStreamSynchronize(b->stream));
    cudaFree(b->d_seq); b->d_seq = nullptr; b->seq_cap = 0;
    CK(cudaMalloc(&b->d_seq, n * sizeof(RobotCommand)));
    b->seq_cap = n;
  }
  CK(cudaMemcpyAsync(b->d_seq, host_cmds, n * sizeof(RobotCommand), cudaMemcpyHostToDevice, b->stream));
  return world_batch_step_sequence(b, b->d_seq, n_periods, steps_per_period);
}

int world_batch_sync(struct WorldBatch *b) {
  ENTER(b); /* joins the chunk streams first */
  SSLSIM_RANGE("world_batch_sync");
  CK(cudaStreamSynchronize(b->stream));
  return SSLSIM_OK;
}

/* ---- chunked pipeline -------------------------------------------------------------------------------------- */
int world_batch_set_chunks(struct WorldBatch *b, int n_chunks) {
  ENTER(b);
  if (n_chunks < 0 || n_chunks > b->W || n_chunks > 256) return SSLSIM_E_ARG;
  CK(cudaStreamSynchronize(b->stream));
  free_chunks(b);
  if (n_chunks <= 1) return SSLSIM_OK;
  if (!b->ev_main) CK(cudaEventCreateWithFlags(&b->ev_main, cudaEventDisableTiming));
  if (!b->d_cmd && b->R > 0) {
    CK(cudaMalloc(&b->d_cmd, (size_t)b->W * b->R * sizeof(RobotCommand)));
    /* zeroed = passive robots: a chunk that is stepped before it was ever given commands must not read garbage */
    CK(cudaMemsetAsync(b->d_cmd, 0, (size_t)b->W * b->R * sizeof(RobotCommand), b->stream));
  }
  b->chunks.resize(n_chunks);
  for (int c = 0; c < n_chunks; ++c) {
    WorldBatch::Chunk &k = b->chunks[c];
    k.first = (int)((long long)b->W * c / n_chunks);
    k.count = (int)((long long)b->W * (c + 1) / n_chunks) - k.first;
    CK(cudaStreamCreateWithFlags(&k.stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&k.done, cudaEventDisableTiming));
  }
  return SSLSIM_OK;
}

int world_batch_chunk_count(const struct WorldBatch *b) { return b ? (int)b->chunks.size() : 0; }

int world_batch_chunk_range(const struct WorldBatch *b, int chunk, int *first_world, int *n_worlds) {
  if (!b || chunk < 0 || chunk >= (int)b->chunks.size()) return SSLSIM_E_ARG;
  if (first_world) *first_world = b->chunks[chunk].first;
  if (n_worlds) *n_worlds = b->chunks[chunk].count;
  return SSLSIM_OK;
}

int world_batch_chunk_step(struct WorldBatch *b, int chunk, const struct RobotCommand *host_cmds, int n_steps,
                           uint32_t *host_aux) {
  ENTER_NOJOIN(b);
  SSLSIM_RANGE("world_batch_chunk_step");
  if (chunk < 0 || chunk >= (int)b->chunks.size() || n_steps < 0) return SSLSIM_E_ARG;
  WorldBatch::Chunk &k = b->chunks[chunk];
  /* after whatever was queued on the batch stream before this call */
  CK(cudaEventRecord(b->ev_main, b->stream));
  CK(cudaStreamWaitEvent(k.stream, b->ev_main, 0));
  if (host_cmds && b->R > 0) {
    CK(cudaMemcpyAsync(b->d_cmd + (size_t)k.first * b->R, host_cmds, (size_t)k.count * b->R * sizeof(RobotCommand),
                       cudaMemcpyHostToDevice, k.stream));
    b->cmd_active = b->d_cmd;
  }
  if (n_steps > 0) {
    const StepArgs a = step_args(b, k.first, k.count, b->cmd_active, n_steps, 2, (float)(1.0 / 60 / 2), 0, 0);
    CK(sslsim_launch_step(a, b->lanes, k.stream));
  }
  if (host_aux)
    CK(cudaMemcpyAsync(host_aux, b->d_aux + (size_t)k.first * 8, (size_t)k.count * 8 * sizeof(uint32_t),
                       cudaMemcpyDeviceToHost, k.stream));
  k.dirty = true;
  if (chunk == 0) { /* the batch clock follows chunk 0; a caller keeps the chunks within one period of each other */
    for (int s = 0; s < n_steps; ++s) { b->timestamp += (float)(1.0 / 60); b->frame++; }
  }
  for (World *w : b->views) w->mirror_valid = false;
  return SSLSIM_OK;
}

/* chunk_step with the full state read back: what the reference's loop does after every step (src/main.c:80-85 reads
 * every pose for the detection packet).  host_pos / host_vel: [chunk worlds][n_robots + 1][4] floats, may be NULL. */
int world_batch_chunk_step_ex(struct WorldBatch *b, int chunk, const struct RobotCommand *host_cmds, int n_steps,
                              uint32_t *host_aux, float *host_pos, float *host_vel) {
  const int rc = world_batch_chunk_step(b, chunk, host_cmds, n_steps, host_aux);
  if (rc != SSLSIM_OK) return rc;
  WorldBatch::Chunk &k = b->chunks[chunk];
  const size_t N = (size_t)b->R + 1, off = (size_t)k.first * N, nb = (size_t)k.count * N * sizeof(float4);
  if (host_pos) CK(cudaMemcpyAsync(host_pos, b->d_pos + off, nb, cudaMemcpyDeviceToHost, k.stream));
  if (host_vel) CK(cudaMemcpyAsync(host_vel, b->d_vel + off, nb, cudaMemcpyDeviceToHost, k.stream));
  return SSLSIM_OK;
}

/* A command sequence for one chunk, on the chunk's stream: dev_cmds holds n_periods blocks of the WHOLE batch's
 * [n_worlds][n_robots] records (the layout of world_batch_step_sequence); the chunk reads its own rows of every block.
 * Chunks submitted one after the other run concurrently, so the slow worlds at the end of one chunk's launch overlap
 * the next chunk's work instead of idling the GPU (small batches: the launch ends with its slowest world). */
int world_batch_chunk_step_sequence(struct WorldBatch *b, int chunk, const struct RobotCommand *dev_cmds, int n_periods,
                                    int steps_per_period) {
  ENTER_NOJOIN(b);
  SSLSIM_RANGE("world_batch_chunk_step_sequence");
  if (chunk < 0 || chunk >= (int)b->chunks.size() || !dev_cmds || n_periods <= 0 || steps_per_period <= 0)
    return SSLSIM_E_ARG;
  WorldBatch::Chunk &k = b->chunks[chunk];
  CK(cudaEventRecord(b->ev_main, b->stream));
  CK(cudaStreamWaitEvent(k.stream, b->ev_main, 0));
  const long long stride = (long long)b->W * b->R;
  const StepArgs a = step_args(b, k.first, k.count, dev_cmds, n_periods * steps_per_period, 2, (float)(1.0 / 60 / 2),
                               steps_per_period, stride);
  CK(sslsim_launch_step(a, b->lanes, k.stream));
  k.dirty = true;
  b->cmd_active = dev_cmds + (size_t)(n_periods - 1) * stride;
  if (chunk == 0)
    for (int s = 0; s < n_periods * steps_per_period; ++s) { b->timestamp += (float)(1.0 / 60); b->frame++; }
  for (World *w : b->views) w->mirror_valid = false;
  return SSLSIM_OK;
}

int world_batch_chunk_wait(struct WorldBatch *b, int chunk) {
  ENTER_NOJOIN(b);
  SSLSIM_RANGE("world_batch_chunk_wait");
  if (chunk < 0 || chunk >= (int)b->chunks.size()) return SSLSIM_E_ARG;
  CK(cudaStreamSynchronize(b->chunks[chunk].stream));
  return SSLSIM_OK;
}

int sslsim_ensure_cmd_buffer(WorldBatch *b) {
  if (!b->d_cmd && b->R > 0) {
    CK(cudaMalloc(&b->d_cmd, (size_t)b->W * b->R * sizeof(RobotCommand)));
    CK(cudaMemsetAsync(b->d_cmd, 0, (size_t)b->W * b->R * sizeof(RobotCommand), b->stream)); /* passive robots */
  }
  return SSLSIM_OK;
}

int world_batch_set_commands(struct WorldBatch *b, const struct RobotCommand *host_cmds) {
  ENTER(b);
  SSLSIM_RANGE("world_batch_set_commands");
  if (!host_cmds) { b->cmd_active = nullptr; return SSLSIM_OK; }
  if (b->R == 0) return SSLSIM_OK;
  int rc = sslsim_ensure_cmd_buffer(b);
  if (rc) return rc;
  CK(cudaMemcpyAsync(b->d_cmd, host_cmds, (size_t)b->W * b->R * sizeof(RobotCommand), cudaMemcpyHostToDevice,
                     b->stream));
  b->cmd_active = b->d_cmd;
  return SSLSIM_OK;
}

int world_batch_set_commands_device(struct WorldBatch *b, const struct RobotCommand *dev_cmds) {
  ENTER(b);
  b->cmd_active = dev_cmds;
  return SSLSIM_OK;
}

int world_batch_clear_commands(struct WorldBatch *b) {
  ENTER(b);
  b->cmd_active = nullptr;
  return SSLSIM_OK;
}

int world_batch_gen_commands(struct WorldBatch *b, uint64_t period, int kind) {
  ENTER(b);
  SSLSIM_RANGE("world_batch_gen_commands");
  if (b->R == 0) return SSLSIM_OK;
  int rc = sslsim_ensure_cmd_buffer(b);
  if (rc) return rc;
  CK(sslsim_launch_gen_commands(b->d_cmd, b->W, b->R, b->seed, b->world0, period, kind, b->stream));
  b->cmd_active = b->d_cmd;
  return SSLSIM_OK;
}

int world_batch_gen_commands_into(struct WorldBatch *b, struct RobotCommand *dev_out, uint64_t period, int kind) {
  ENTER(b);
  SSLSIM_RANGE("world_batch_gen_commands_into");
  if (!dev_out) return SSLSIM_E_ARG;
  CK(sslsim_launch_gen_commands(dev_out, b->W, b->R, b->seed, b->world0, period, kind, b->stream));
  return SSLSIM_OK;
}

static int ensure_rep(WorldBatch *b, size_t bytes) {
  if (bytes > b->rep_cap) {
    CK(cudaStreamSynchronize(b->stream));
    cudaFree(b->d_rep); b->d_rep = nullptr; b->rep_cap = 0;
    size_t cap = bytes < 4096 ? 4096 : bytes * 2;
    CK(cudaMalloc(&b->d_rep, cap));
    b->rep_cap = cap;
  }
  return SSLSIM_OK;
}

int world_batch_replace_robots(struct WorldBatch *b, const struct RobotReplacement *recs, int n) {
  ENTER(b);
  SSLSIM_RANGE("world_batch_replace_robots");
  if (n < 0 || (n > 0 && !recs)) return SSLSIM_E_ARG;
  if (n == 0) return SSLSIM_OK;
  int rc = ensure_rep(b, (size_t)n * sizeof(RobotReplacement));
  if (rc) return rc;
  CK(cudaMemcpyAsync(b->d_rep, recs, (size_t)n * sizeof(RobotReplacement), cudaMemcpyHostToDevice, b->stream));
  CK(sslsim_launch_replace_robots(b->d_pos, b->d_warm, sslsim_warm_words(b), b->W, b->R, (const RobotReplacement *)b->d_rep, n, b->stream));
  CK(cudaStreamSynchronize(b->stream)); /* the staging buffer is reused by the next call */
  for (World *w : b->views) w->mirror_valid = false;
  return SSLSIM_OK;
}

int world_batch_replace_balls(struct WorldBatch *b, const struct BallReplacement *recs, int n) {
  ENTER(b);
  SSLSIM_RANGE("world_batch_replace_balls");
  if (n < 0 || (n > 0 && !recs)) return SSLSIM_E_ARG;
  if (n == 0) return SSLSIM_OK;
  int rc = ensure_rep(b, (size_t)n * sizeof(BallReplacement));
  if (rc) return rc;
  CK(cudaMemcpyAsync(b->d_rep, recs, (size_t)n * sizeof(BallReplacement), cudaMemcpyHostToDevice, b->stream));
  CK(sslsim_launch_replace_balls(b->d_pos, b->d_vel, b->d_aux, b->d_warm, sslsim_warm_words(b), b->W, b->R, (const BallReplacement *)b->d_rep, n, b->stream));
  CK(cudaStreamSynchronize(b->stream));
  for (World *w : b->views) w->mirror_valid = false;
  return SSLSIM_OK;
}

int world_batch_read_state(struct WorldBatch *b, float *pos, float *vel, uint32_t *aux) {
  ENTER(b);
  SSLSIM_RANGE("world_batch_read_state");
  const size_t nb = (size_t)b->W * (b->R + 1) * sizeof(float4);
  if (pos) CK(cudaMemcpyAsync(pos, b->d_pos, nb, cudaMemcpyDeviceToHost, b->stream));
  if (vel) CK(cudaMemcpyAsync(vel, b->d_vel, nb, cudaMemcpyDeviceToHost, b->stream));
  if (aux) CK(cudaMemcpyAsync(aux, b->d_aux, (size_t)b->W * 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, b->stream));
  CK(cudaStreamSynchronize(b->stream));
  return SSLSIM_OK;
}

/* bulk readout of selected worlds: pos/vel [n][n_robots + 1][4], aux [n][8]; any pointer may be NULL */
int world_batch_read_worlds(struct WorldBatch *b, const int *world_ids, int n, float *pos, float *vel, uint32_t *aux) {
  ENTER(b);
  SSLSIM_RANGE("world_batch_read_worlds");
  if (n < 0 || (n > 0 && !world_ids)) return SSLSIM_E_ARG;
  const size_t N = (size_t)b->R + 1;
  for (int i = 0; i < n; ++i) {
    const int w = world_ids[i];
    if (w < 0 || w >= b->W) return SSLSIM_E_ARG;
    if (pos) CK(cudaMemcpyAsync(pos + (size_t)i * N * 4, b->d_pos + (size_t)w * N, N * sizeof(float4), cudaMemcpyDeviceToHost, b->stream));
    if (vel) CK(cudaMemcpyAsync(vel + (size_t)i * N * 4, b->d_vel + (size_t)w * N, N * sizeof(float4), cudaMemcpyDeviceToHost, b->stream));
    if (aux) CK(cudaMemcpyAsync(aux + (size_t)i * 8, b->d_aux + (size_t)w * 8, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, b->stream));
  }
  CK(cudaStreamSynchronize(b->stream));
  return SSLSIM_OK;
}

int world_batch_read_aux(struct WorldBatch *b, uint32_t *aux) { return world_batch_read_state(b, nullptr, nullptr, aux); }

int world_batch_write_state(struct WorldBatch *b, const float *pos, const float *vel, const uint32_t *aux) {
  ENTER(b);
  SSLSIM_RANGE("world_batch_write_state");
  const size_t nb = (size_t)b->W * (b->R + 1) * sizeof(float4);
  if (pos) CK(cudaMemcpyAsync(b->d_pos, pos, nb, cudaMemcpyHostToDevice, b->stream));
  if (vel) CK(cudaMemcpyAsync(b->d_vel, vel, nb, cudaMemcpyHostToDevice, b->stream));
  if (aux) CK(cudaMemcpyAsync(b->d_aux, aux, (size_t)b->W * 8 * sizeof(uint32_t), cudaMemcpyHostToDevice, b->stream));
  CK(sslsim_clear_warm(b, 0, b->W)); /* the written bodies' contact points are gone (include/sslsim_batch.h) */
  CK(cudaStreamSynchronize(b->stream));
  for (World *w : b->views) w->mirror_valid = false;
  return SSLSIM_OK;
}

int world_batch_warm_words(const struct WorldBatch *b) { return b ? sslsim_warm_words(b) : 0; }

int world_batch_read_warm(struct WorldBatch *b, uint32_t *warm) {
  ENTER(b);
  SSLSIM_RANGE("world_batch_read_warm");
  if (!warm) return SSLSIM_E_ARG;
  CK(cudaMemcpyAsync(warm, b->d_warm, (size_t)b->W * sslsim_warm_words(b) * sizeof(uint32_t), cudaMemcpyDeviceToHost, b->stream));
  CK(cudaStreamSynchronize(b->stream));
  return SSLSIM_OK;
}

int world_batch_write_warm(struct WorldBatch *b, const uint32_t *warm) {
  ENTER(b);
  SSLSIM_RANGE("world_batch_write_warm");
  if (!warm) return SSLSIM_E_ARG;
  CK(cudaMemcpyAsync(b->d_warm, warm, (size_t)b->W * sslsim_warm_words(b) * sizeof(uint32_t), cudaMemcpyHostToDevice, b->stream));
  CK(cudaStreamSynchronize(b->stream));
  return SSLSIM_OK;
}

int world_batch_device_state(struct WorldBatch *b, void **pos, void **vel, void **aux, void **cmd) {
  if (!b) return SSLSIM_E_ARG;
  if (pos) *pos = b->d_pos;
  if (vel) *vel = b->d_vel;
  if (aux) *aux = b->d_aux;
  if (cmd) *cmd = (void *)b->cmd_active;
  return SSLSIM_OK;
}

size_t world_batch_detection_record_size(const struct WorldBatch *b) {
  return b ? (size_t)(6 + 7 * b->R) * sizeof(float) : 0;
}

int world_batch_gather_detections(struct WorldBatch *b, const int *world_ids, int n, void *out, size_t capacity) {
  ENTER(b);
  SSLSIM_RANGE("world_batch_gather_detections");
  if (n < 0 || (n > 0 && (!world_ids || !out))) return SSLSIM_E_ARG;
  const size_t rec = world_batch_detection_record_size(b);
  if ((size_t)n * rec > capacity) return SSLSIM_E_SPACE;
  if (n == 0) return 0;
  if (n > b->gather_cap) {
    CK(cudaStreamSynchronize(b->stream));
    cudaFree(b->d_world_ids); cudaFree(b->d_gather);
    if (b->h_gather) cudaFreeHost(b->h_gather);
    b->d_world_ids = nullptr; b->d_gather = nullptr; b->h_gather = nullptr; b->gather_cap = 0;
    const int cap = n < 64 ? 64 : n;
    CK(cudaMalloc(&b->d_world_ids, (size_t)cap * sizeof(int)));
    CK(cudaMalloc(&b->d_gather, (size_t)cap * rec));
    CK(cudaMallocHost(&b->h_gather, (size_t)cap * rec));
    b->gather_cap = cap;
  }
  CK(cudaMemcpyAsync(b->d_world_ids, world_ids, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, b->stream));
  CK(sslsim_launch_gather(b->d_pos, b->W, b->R, b->d_world_ids, n, b->d_robot_ids, b->d_gather, b->stream));
  CK(cudaMemcpyAsync(b->h_gather, b->d_gather, (size_t)n * rec, cudaMemcpyDeviceToHost, b->stream));
  CK(cudaStreamSynchronize(b->stream));
  memcpy(out, b->h_gather, (size_t)n * rec);
  return (int)((size_t)n * rec);
}

int world_batch_referee(struct WorldBatch *b) {
  ENTER(b);
  SSLSIM_RANGE("world_batch_referee");
  CK(sslsim_launch_referee(b->d_pos, b->d_vel, b->d_aux, b->d_warm, sslsim_warm_words(b), b->W, b->R, b->field.field_length, b->field.field_width, b->stream));
  for (World *w : b->views) w->mirror_valid = false;
  return SSLSIM_OK;
}

struct WorldSnapshot {
  int device, W, R;
  float4 *pos, *vel;
  uint32_t *aux, *warm;
  unsigned int frame;
  float timestamp;
};

struct WorldSnapshot *world_batch_snapshot(struct WorldBatch *b) {
  if (!b || b->last_error == SSLSIM_E_CUDA) return nullptr;
  cudaSetDevice(b->device);
  if (!b->chunks.empty() && sslsim_join_chunks(b)) return nullptr; /* after all queued chunk work */
  WorldSnapshot *s = new (std::nothrow) WorldSnapshot();
  if (!s) return nullptr;
  s->device = b->device; s->W = b->W; s->R = b->R; s->frame = b->frame; s->timestamp = b->timestamp;
  s->pos = s->vel = nullptr; s->aux = nullptr; s->warm = nullptr;
  const size_t nb = (size_t)b->W * (b->R + 1) * sizeof(float4), na = (size_t)b->W * 8 * sizeof(uint32_t);
  const size_t nw = (size_t)b->W * sslsim_warm_words(b) * sizeof(uint32_t);
  const bool ok = sslsim_check(b, cudaMalloc(&s->pos, nb), "snapshot alloc") && sslsim_check(b, cudaMalloc(&s->vel, nb), "snapshot alloc") &&
                  sslsim_check(b, cudaMalloc(&s->aux, na), "snapshot alloc") &&
                  sslsim_check(b, cudaMalloc(&s->warm, nw), "snapshot alloc") &&
                  sslsim_check(b, cudaMemcpyAsync(s->warm, b->d_warm, nw, cudaMemcpyDeviceToDevice, b->stream), "snapshot copy") &&
                  sslsim_check(b, cudaMemcpyAsync(s->pos, b->d_pos, nb, cudaMemcpyDeviceToDevice, b->stream), "snapshot copy") &&
                  sslsim_check(b, cudaMemcpyAsync(s->vel, b->d_vel, nb, cudaMemcpyDeviceToDevice, b->stream), "snapshot copy") &&
                  sslsim_check(b, cudaMemcpyAsync(s->aux, b->d_aux, na, cudaMemcpyDeviceToDevice, b->stream), "snapshot copy") &&
                  sslsim_check(b, cudaStreamSynchronize(b->stream), "snapshot sync");
  if (!ok) { world_snapshot_free(s); return nullptr; }
  return s;
}

int world_batch_restore(struct WorldBatch *b, const struct WorldSnapshot *s) {
  ENTER(b);
  SSLSIM_RANGE("world_batch_restore");
  if (!s || s->W != b->W || s->R != b->R || s->device != b->device) return SSLSIM_E_ARG;
  const size_t nb = (size_t)b->W * (b->R + 1) * sizeof(float4), na = (size_t)b->W * 8 * sizeof(uint32_t);
  CK(cudaMemcpyAsync(b->d_pos, s->pos, nb, cudaMemcpyDeviceToDevice, b->stream));
  CK(cudaMemcpyAsync(b->d_vel, s->vel, nb, cudaMemcpyDeviceToDevice, b->stream));
  CK(cudaMemcpyAsync(b->d_aux, s->aux, na, cudaMemcpyDeviceToDevice, b->stream));
  CK(cudaMemcpyAsync(b->d_warm, s->warm, (size_t)b->W * sslsim_warm_words(b) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, b->stream));
  b->frame = s->frame; b->timestamp = s->timestamp;
  for (World *w : b->views) w->mirror_valid = false;
  return SSLSIM_OK;
}

void world_snapshot_free(struct WorldSnapshot *s) {
  if (!s) return;
  cudaSetDevice(s->device);
  cudaFree(s->pos); cudaFree(s->vel); cudaFree(s->aux); cudaFree(s->warm);
  delete s;
}

int world_batch_fork(struct WorldBatch *b, int src, const int *dst, int n) {
  ENTER(b);
  SSLSIM_RANGE("world_batch_fork");
  if (src < 0 || src >= b->W || n < 0 || (n > 0 && !dst)) return SSLSIM_E_ARG;
  if (n == 0) return SSLSIM_OK;
  int rc = ensure_rep(b, (size_t)n * sizeof(int));
  if (rc) return rc;
  CK(cudaMemcpyAsync(b->d_rep, dst, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, b->stream));
  CK(sslsim_launch_fork(b->d_pos, b->d_vel, b->d_aux, b->d_warm, sslsim_warm_words(b), b->W, b->R, src, (const int *)b->d_rep, n, b->stream));
  CK(cudaStreamSynchronize(b->stream));
  for (World *w : b->views) w->mirror_valid = false;
  return SSLSIM_OK;
}

int world_batch_reduce_stats(struct WorldBatch *b, struct BatchStats *out) {
  ENTER(b);
  SSLSIM_RANGE("world_batch_reduce_stats");
  if (!out) return SSLSIM_E_ARG;
  CK(sslsim_launch_reduce(b->d_pos, b->d_vel, b->d_aux, b->W, b->R, b->world0, b->d_stats, b->stream));
  unsigned long long h[8];
  CK(cudaMemcpyAsync(h, b->d_stats, sizeof h, cudaMemcpyDeviceToHost, b->stream));
  CK(cudaStreamSynchronize(b->stream));
  out->goals_blue = h[0]; out->goals_yellow = h[1]; out->ball_out = h[2]; out->kicks = h[3];
  out->touches = h[4]; out->contacts = h[5]; out->bad_worlds = h[6]; out->checksum = h[7];
  return SSLSIM_OK;
}

void *world_batch_stream(struct WorldBatch *b) { return b ? (void *)b->stream : nullptr; }

int world_batch_set_stream(struct WorldBatch *b, void *cuda_stream) {
  ENTER(b);
  CK(cudaStreamSynchronize(b->stream));
  if (b->own_stream && b->stream) cudaStreamDestroy(b->stream);
  b->stream = (cudaStream_t)cuda_stream;
  b->own_stream = false;
  return SSLSIM_OK;
}

int world_batch_timer_start(struct WorldBatch *b) {
  ENTER(b);
  CK(cudaEventRecord(b->ev0, b->stream));
  return SSLSIM_OK;
}

int world_batch_timer_stop(struct WorldBatch *b, float *elapsed_ms) {
  ENTER(b);
  CK(cudaEventRecord(b->ev1, b->stream));
  CK(cudaEventSynchronize(b->ev1));
  float ms = 0.0f;
  CK(cudaEventElapsedTime(&ms, b->ev0, b->ev1));
  if (elapsed_ms) *elapsed_ms = ms;
  return SSLSIM_OK;
}

int world_batch_enable_kernel_timing(struct WorldBatch *b, int on) {
  ENTER(b);
  b->ktiming = on != 0;
  return SSLSIM_OK;
}

int world_batch_kernel_time(struct WorldBatch *b, float *total_ms, int *launches, int reset) {
  ENTER(b);
  CK(cudaStreamSynchronize(b->stream));
  float tot = 0.0f;
  for (size_t k = 0; k + 1 < b->kev_used; k += 2) {
    float ms = 0.0f;
    CK(cudaEventElapsedTime(&ms, b->kev[k], b->kev[k + 1]));
    tot += ms;
  }
  if (total_ms) *total_ms = tot;
  if (launches) *launches = (int)(b->kev_used / 2);
  if (reset) b->kev_used = 0;
  return SSLSIM_OK;
}

#ifdef SSLSIM_PROFILE
/* profile builds only (not part of the public ABI): per-world counters of the last step launch */
int world_batch_profile_counters(struct WorldBatch *b, unsigned *out) {
  ENTER(b);
  if (!b->d_prof) {
    CK(cudaMalloc(&b->d_prof, (size_t)b->W * 32 * sizeof(unsigned)));
    CK(cudaMemset(b->d_prof, 0, (size_t)b->W * 32 * sizeof(unsigned)));
    return 1;
  }
  CK(cudaStreamSynchronize(b->stream));
  CK(cudaMemcpy(out, b->d_prof, (size_t)b->W * 32 * sizeof(unsigned), cudaMemcpyDeviceToHost));
  return SSLSIM_OK;
}
#endif

/* self-test hook: mismatches of the kernel's branch-free sqrt / division against sqrtf and operator/ */
int sslsim_selftest_math(int device, unsigned long long n, unsigned long long seed, unsigned long long *bad_sqrt,
                         unsigned long long *bad_div) {
  if (cudaSetDevice(device) != cudaSuccess) return SSLSIM_E_CUDA;
  unsigned long long *d = nullptr, h[2] = {0, 0};
  if (cudaMalloc(&d, sizeof h) != cudaSuccess) return SSLSIM_E_NOMEM;
  cudaMemset(d, 0, sizeof h);
  cudaError_t e = sslsim_selftest_math_launch(n, seed, d, 0);
  if (e == cudaSuccess) e = cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return SSLSIM_E_CUDA;
  if (bad_sqrt) *bad_sqrt = h[0];
  if (bad_div) *bad_div = h[1];
  return SSLSIM_OK;
}

int world_batch_set_lanes_per_world(struct WorldBatch *b, int lanes) {
  if (!b) return SSLSIM_E_ARG;
  if (lanes != 0 && lanes != 16 && lanes != 32) return SSLSIM_E_ARG;
  if (lanes == 16 && b->R + 1 > 16) return SSLSIM_E_ARG;
  b->lanes = lanes;
  return SSLSIM_OK;
}

int world_batch_last_error(const struct WorldBatch *b) { return b ? b->last_error : SSLSIM_E_ARG; }
const char *world_batch_error_string(const struct WorldBatch *b) { return b ? b->errstr : "null batch"; }
const char *sslsim_b200_version(void) { return "sslsim_b200 0.1 (sm_100a)"; }

} /* extern "C" */
