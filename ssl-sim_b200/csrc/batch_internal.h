/*
 * batch_internal.h -- host-side objects behind the C ABI (not public).
 */
#ifndef SSLSIM_B200_BATCH_INTERNAL_H
#define SSLSIM_B200_BATCH_INTERNAL_H
#include <vector>
#include "dev_types.h"

/* NVTX ranges around every launch site of the C ABI (tracing hooks), so a timeline (nsys, ncu --nvtx) shows
 * which API call a kernel or copy belongs to.  Header-only NVTX v3: without a tool attached a range is one
 * predictable branch.  -DSSLSIM_NO_NVTX compiles them out. */
#ifndef SSLSIM_NO_NVTX
#include <nvtx3/nvToolsExt.h>
struct SslsimRange {
  explicit SslsimRange(const char *name) { nvtxRangePushA(name); }
  ~SslsimRange() { nvtxRangePop(); }
  SslsimRange(const SslsimRange &) = delete;
  SslsimRange &operator=(const SslsimRange &) = delete;
};
#define SSLSIM_RANGE(name) SslsimRange sslsim_range_(name)
#else
#define SSLSIM_RANGE(name) do { } while (0)
#endif

#define SSLSIM_LEGACY_MAX_ROBOTS 100 /* reference src/sslsim.cpp:215 stack_vector<Robot,100> */
#define SSLSIM_LEGACY_MAX_BALLS 20   /* reference src/sslsim.cpp:214 */
#define SSLSIM_MAX_BODIES 32

struct World;

/* interior-pointer handles, stable for the world's lifetime (reference src/sslsim.cpp:334-356) */
struct Robot { World *world; int index; int id; enum Team team; };
struct Ball { World *world; int index; };

struct WorldBatch {
  int device = 0, W = 0, R = 0;
  int cap_bodies = 0; /* allocated bodies per world (>= R+1; 32 for a growable legacy world) */
  uint64_t seed = 0, world0 = 0;
  FieldGeometry field{};
  const FieldGeometry *field_ptr = nullptr; /* the caller's (borrowed) pointer, returned by world_get_field */
  DevField dfield{};
  DevField *d_field = nullptr; /* device copy for the out-of-line general path */
  unsigned *d_prof = nullptr;  /* SSLSIM_PROFILE builds only */
  SslSimParams params{};
  float4 *d_pos = nullptr, *d_vel = nullptr;
  uint32_t *d_aux = nullptr;
  uint32_t *d_warm = nullptr;             /* warm tables [W][SSLSIM_WARM_WORDS(R)] (allocated for cap_bodies - 1 robots) */
  RobotCommand *d_cmd = nullptr;          /* owned command buffer (lazy) */
  const RobotCommand *cmd_active = nullptr; /* what the step kernel reads; NULL = passive robots */
  int *d_robot_ids = nullptr;
  int *d_world_ids = nullptr; float *d_gather = nullptr; int gather_cap = 0;
  float *h_gather = nullptr; /* pinned */
  unsigned long long *d_stats = nullptr;
  void *d_rep = nullptr; size_t rep_cap = 0;
  RobotCommand *d_seq = nullptr; size_t seq_cap = 0; /* staging for host command sequences */
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool ktiming = false;
  std::vector<cudaEvent_t> kev; /* pairs around step launches */
  size_t kev_used = 0;
  unsigned int frame = 0;
  float timestamp = 0.0f;
  int lanes = 0;
  int last_error = 0;
  char errstr[256] = {0};
  std::vector<int> robot_ids; /* per robot slot */
  std::vector<int> robot_teams;
  std::vector<World *> views;
  /* chunked pipeline (world_batch_set_chunks): contiguous world ranges, one stream each */
  struct Chunk { int first = 0, count = 0; cudaStream_t stream = nullptr; cudaEvent_t done = nullptr; bool dirty = false; };
  std::vector<Chunk> chunks;
  cudaEvent_t ev_main = nullptr; /* orders chunk work after earlier work on the batch stream */
};

struct World {
  WorldBatch *batch = nullptr;
  int index = 0;
  bool owns_batch = false;
  Robot robots[SSLSIM_LEGACY_MAX_ROBOTS];
  int n_robots = 0;
  Ball balls[SSLSIM_LEGACY_MAX_BALLS];
  int n_balls = 1;
  /* host mirror of this world's device state */
  float pos[SSLSIM_MAX_BODIES][4];
  float vel[SSLSIM_MAX_BODIES][4];
  uint32_t aux[8];
  bool mirror_valid = false;
  float local_time = 0.0f; /* Bullet's stepSimulation accumulator (m_localTime) */
};

bool sslsim_check(WorldBatch *b, cudaError_t e, const char *what);
static inline int sslsim_warm_words(const WorldBatch *b) { return SSLSIM_WARM_WORDS(b->R); }
/* empties the warm tables of worlds [first, first + count) on the batch stream (every external state write does) */
cudaError_t sslsim_clear_warm(WorldBatch *b, int first, int count);
int sslsim_join_chunks(WorldBatch *b); /* orders the batch stream after everything queued on the chunk streams */
extern "C" int sslsim_ensure_cmd_buffer(WorldBatch *b); /* the batch-owned command buffer, zeroed (= passive robots) on first use */
WorldBatch *sslsim_batch_create(const FieldGeometry *field, int n_worlds, int n_robots, int cap_bodies, int device,
                                uint64_t seed, uint64_t world0, int init_mode, const SslSimParams *params,
                                bool run_init);
int sslsim_batch_step(WorldBatch *b, int n_steps, int substeps, float h, float time_step_total,
                      int steps_per_period = 0, long long cmd_period_stride = 0);
#endif
