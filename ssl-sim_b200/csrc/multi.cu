/*
 * multi.cu -- include/sslsim_multi.h: contiguous world ranges over several GPUs, no per-step collective, one
 * ncclAllGather of the episode statistics.  NCCL is resolved with dlopen so that the library has
 * no link-time dependency on it.
 */
#include <dlfcn.h>
#include <nccl.h>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>
#include "batch_internal.h"
#include "sslsim_multi.h"

namespace {

struct NcclApi {
  void *handle = nullptr;
  bool tried = false;
  char version[32] = "unavailable";
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GetVersion)(int *) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

bool nccl_load() {
  NcclApi &n = g_nccl;
  if (n.tried) return n.handle != nullptr;
  n.tried = true;
  /* a process that already carries an NCCL (PyTorch) gets that one: same soname */
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) { fprintf(stderr, "sslsim_b200: NCCL not found (%s)\n", dlerror()); return false; }
#define SYM(field, name) *(void **)(&n.field) = dlsym(h, name); if (!n.field) { fprintf(stderr, "sslsim_b200: NCCL lacks %s\n", name); dlclose(h); return false; }
  SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommInitAll, "ncclCommInitAll")
  SYM(AllGather, "ncclAllGather") SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd")
  SYM(CommDestroy, "ncclCommDestroy") SYM(GetVersion, "ncclGetVersion") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  int v = 0;
  if (n.GetVersion(&v) == ncclSuccess) snprintf(n.version, sizeof n.version, "%d.%d.%d", v / 10000, (v / 100) % 100, v % 100);
  n.handle = h;
  return true;
}

} // namespace

struct WorldMulti {
  int n_ranks = 0;
  long long total = 0;
  int R = 0;
  std::vector<WorldBatch *> shards;   /* local */
  std::vector<int> ranks;             /* global shard index of each local shard */
  std::vector<ncclComm_t> comms;
  std::vector<unsigned long long *> d_all; /* per local shard: gathered stats [n_ranks][8] on its device */
  std::vector<cudaEvent_t> ev0, ev1;
  char errstr[256] = {0};
};

static int fail(WorldMulti *m, int code, const char *what, const char *detail) {
  if (m) snprintf(m->errstr, sizeof m->errstr, "%s: %s", what, detail ? detail : "");
  fprintf(stderr, "sslsim_b200: %s: %s\n", what, detail ? detail : "");
  return code;
}

static bool add_shard(WorldMulti *m, const FieldGeometry *field, int rank, int device, uint64_t seed, int init_mode,
                      const SslSimParams *params) {
  long long first = 0; int count = 0;
  sslsim_shard_range(m->total, rank, m->n_ranks, &first, &count);
  if (count <= 0) { fail(m, SSLSIM_E_ARG, "new_world_multi", "fewer worlds than shards"); return false; }
  WorldBatch *b = sslsim_batch_create(field, count, m->R, 0, device, seed, (uint64_t)first, init_mode, params, true);
  if (!b) { fail(m, SSLSIM_E_CUDA, "new_world_multi", "shard creation failed"); return false; }
  m->shards.push_back(b); m->ranks.push_back(rank);
  unsigned long long *d = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaSetDevice(device);
  if (cudaMalloc(&d, (size_t)m->n_ranks * 8 * sizeof(unsigned long long)) != cudaSuccess ||
      cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) {
    fail(m, SSLSIM_E_NOMEM, "new_world_multi", "stats buffer / events");
    return false;
  }
  m->d_all.push_back(d); m->ev0.push_back(e0); m->ev1.push_back(e1);
  return true;
}

extern "C" {

void sslsim_shard_range(long long total_worlds, int rank, int n_ranks, long long *first, int *count) {
  const long long base = total_worlds / n_ranks, rem = total_worlds % n_ranks;
  if (first) *first = rank * base + (rank < rem ? rank : rem);
  if (count) *count = (int)(base + (rank < rem ? 1 : 0));
}

void delete_world_multi(struct WorldMulti *m) {
  if (!m) return;
  for (size_t i = 0; i < m->shards.size(); ++i) {
    cudaSetDevice(m->shards[i]->device);
    cudaStreamSynchronize(m->shards[i]->stream);
    if (i < m->comms.size() && m->comms[i] && g_nccl.CommDestroy) g_nccl.CommDestroy(m->comms[i]);
    if (i < m->d_all.size()) cudaFree(m->d_all[i]);
    if (i < m->ev0.size()) { cudaEventDestroy(m->ev0[i]); cudaEventDestroy(m->ev1[i]); }
    delete_world_batch(m->shards[i]);
  }
  delete m;
}

struct WorldMulti *new_world_multi(const struct FieldGeometry *field, long long total_worlds, int n_robots,
                                   const int *devices, int n_devices, uint64_t seed, int init_mode,
                                   const struct SslSimParams *params) {
  if (!field || !devices || n_devices <= 0 || total_worlds < n_devices) {
    fail(nullptr, SSLSIM_E_ARG, "new_world_multi", "bad argument");
    return nullptr;
  }
  if (!nccl_load()) return nullptr;
  WorldMulti *m = new (std::nothrow) WorldMulti();
  if (!m) return nullptr;
  m->n_ranks = n_devices; m->total = total_worlds; m->R = n_robots;
  for (int i = 0; i < n_devices; ++i)
    if (!add_shard(m, field, i, devices[i], seed, init_mode, params)) { delete_world_multi(m); return nullptr; }
  m->comms.assign(n_devices, nullptr);
  const ncclResult_t r = g_nccl.CommInitAll(m->comms.data(), n_devices, devices);
  if (r != ncclSuccess) {
    fail(m, SSLSIM_E_NCCL, "ncclCommInitAll", g_nccl.GetErrorString(r));
    m->comms.clear();
    delete_world_multi(m);
    return nullptr;
  }
  return m;
}

int sslsim_multi_unique_id(void *id_bytes) {
  if (!id_bytes) return SSLSIM_E_ARG;
  if (!nccl_load()) return SSLSIM_E_NCCL;
  static_assert(sizeof(ncclUniqueId) == SSLSIM_NCCL_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  const ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r != ncclSuccess) return fail(nullptr, SSLSIM_E_NCCL, "ncclGetUniqueId", g_nccl.GetErrorString(r));
  memcpy(id_bytes, &id, sizeof id);
  return SSLSIM_OK;
}

struct WorldMulti *new_world_multi_rank(const struct FieldGeometry *field, long long total_worlds, int n_robots,
                                        int rank, int n_ranks, int device, const void *id_bytes, uint64_t seed,
                                        int init_mode, const struct SslSimParams *params) {
  if (!field || !id_bytes || n_ranks <= 0 || rank < 0 || rank >= n_ranks || total_worlds < n_ranks) {
    fail(nullptr, SSLSIM_E_ARG, "new_world_multi_rank", "bad argument");
    return nullptr;
  }
  if (!nccl_load()) return nullptr;
  WorldMulti *m = new (std::nothrow) WorldMulti();
  if (!m) return nullptr;
  m->n_ranks = n_ranks; m->total = total_worlds; m->R = n_robots;
  if (!add_shard(m, field, rank, device, seed, init_mode, params)) { delete_world_multi(m); return nullptr; }
  ncclUniqueId id;
  memcpy(&id, id_bytes, sizeof id);
  m->comms.assign(1, nullptr);
  cudaSetDevice(device);
  const ncclResult_t r = g_nccl.CommInitRank(&m->comms[0], n_ranks, id, rank);
  if (r != ncclSuccess) {
    fail(m, SSLSIM_E_NCCL, "ncclCommInitRank", g_nccl.GetErrorString(r));
    m->comms.clear();
    delete_world_multi(m);
    return nullptr;
  }
  return m;
}

int world_multi_rank_count(const struct WorldMulti *m) { return m ? m->n_ranks : 0; }
int world_multi_local_count(const struct WorldMulti *m) { return m ? (int)m->shards.size() : 0; }
struct WorldBatch *world_multi_shard(struct WorldMulti *m, int i) {
  return (m && i >= 0 && i < (int)m->shards.size()) ? m->shards[i] : nullptr;
}
int world_multi_shard_rank(const struct WorldMulti *m, int i) {
  return (m && i >= 0 && i < (int)m->ranks.size()) ? m->ranks[i] : -1;
}
long long world_multi_world_count(const struct WorldMulti *m) { return m ? m->total : 0; }

#define EACH(call)                                        \
  if (!m) return SSLSIM_E_ARG;                            \
  for (WorldBatch *b : m->shards) {                       \
    const int rc = (call);                                \
    if (rc < 0) { snprintf(m->errstr, sizeof m->errstr, "shard on device %d: %.200s", b->device, b->errstr); return rc; } \
  }                                                       \
  return SSLSIM_OK

int world_multi_step(struct WorldMulti *m, int n_steps) { EACH(world_batch_step(b, n_steps)); }
int world_multi_gen_commands(struct WorldMulti *m, uint64_t period, int kind) { EACH(world_batch_gen_commands(b, period, kind)); }
int world_multi_sync(struct WorldMulti *m) { EACH(world_batch_sync(b)); }
int world_multi_set_commands(struct WorldMulti *m, const struct RobotCommand *host_cmds) {
  EACH(world_batch_set_commands(b, host_cmds ? host_cmds + (size_t)b->world0 * b->R : nullptr));
}

int world_multi_read_worlds(struct WorldMulti *m, const long long *ids, int n, float *pos, float *vel, uint32_t *aux) {
  if (!m || n < 0 || (n > 0 && !ids)) return SSLSIM_E_ARG;
  const size_t N = (size_t)m->R + 1;
  int got = 0;
  for (int i = 0; i < n; ++i)
    for (WorldBatch *b : m->shards) {
      const long long lo = (long long)b->world0;
      if (ids[i] < lo || ids[i] >= lo + b->W) continue;
      const int local = (int)(ids[i] - lo);
      const int rc = world_batch_read_worlds(b, &local, 1, pos ? pos + (size_t)i * N * 4 : nullptr,
                                             vel ? vel + (size_t)i * N * 4 : nullptr, aux ? aux + (size_t)i * 8 : nullptr);
      if (rc < 0) return rc;
      ++got;
    }
  return got;
}

int world_multi_allgather_stats(struct WorldMulti *m, struct BatchStats *per_shard, struct BatchStats *total) {
  SSLSIM_RANGE("world_multi_allgather_stats");
  if (!m) return SSLSIM_E_ARG;
  if (m->comms.size() != m->shards.size()) return fail(m, SSLSIM_E_NCCL, "world_multi_allgather_stats", "no communicator");
  /* K4 on every local shard, on its own stream */
  for (WorldBatch *b : m->shards) {
    cudaSetDevice(b->device);
    if (!b->chunks.empty() && sslsim_join_chunks(b)) return SSLSIM_E_CUDA;
    if (!sslsim_check(b, sslsim_launch_reduce(b->d_pos, b->d_vel, b->d_aux, b->W, b->R, b->world0, b->d_stats, b->stream),
                      "reduce_stats")) return SSLSIM_E_CUDA;
  }
  /* the exchange: 8 x u64 per shard, all-gathered over NVLink / NVSwitch, ordered after K4 on each stream */
  ncclResult_t r = g_nccl.GroupStart();
  for (size_t i = 0; r == ncclSuccess && i < m->shards.size(); ++i) {
    cudaSetDevice(m->shards[i]->device);
    r = g_nccl.AllGather(m->shards[i]->d_stats, m->d_all[i], 8, ncclUint64, m->comms[i], m->shards[i]->stream);
  }
  const ncclResult_t r2 = g_nccl.GroupEnd();
  if (r == ncclSuccess) r = r2;
  if (r != ncclSuccess) return fail(m, SSLSIM_E_NCCL, "ncclAllGather", g_nccl.GetErrorString(r));
  std::vector<unsigned long long> h((size_t)m->n_ranks * 8);
  WorldBatch *b0 = m->shards[0];
  cudaSetDevice(b0->device);
  if (!sslsim_check(b0, cudaMemcpyAsync(h.data(), m->d_all[0], h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, b0->stream), "stats D2H") ||
      !sslsim_check(b0, cudaStreamSynchronize(b0->stream), "stats sync")) return SSLSIM_E_CUDA;
  for (size_t i = 1; i < m->shards.size(); ++i) { /* the other local streams carried the same collective */
    cudaSetDevice(m->shards[i]->device);
    if (!sslsim_check(m->shards[i], cudaStreamSynchronize(m->shards[i]->stream), "stats sync")) return SSLSIM_E_CUDA;
  }
  BatchStats tot;
  memset(&tot, 0, sizeof tot);
  for (int k = 0; k < m->n_ranks; ++k) {
    const unsigned long long *q = &h[(size_t)k * 8];
    BatchStats s;
    s.goals_blue = q[0]; s.goals_yellow = q[1]; s.ball_out = q[2]; s.kicks = q[3];
    s.touches = q[4]; s.contacts = q[5]; s.bad_worlds = q[6]; s.checksum = q[7];
    if (per_shard) per_shard[k] = s;
    tot.goals_blue += s.goals_blue; tot.goals_yellow += s.goals_yellow; tot.ball_out += s.ball_out; tot.kicks += s.kicks;
    tot.touches += s.touches; tot.contacts += s.contacts; tot.bad_worlds += s.bad_worlds; tot.checksum += s.checksum;
  }
  if (total) *total = tot;
  return SSLSIM_OK;
}

int world_multi_timer_start(struct WorldMulti *m) {
  if (!m) return SSLSIM_E_ARG;
  for (size_t i = 0; i < m->shards.size(); ++i) {
    WorldBatch *b = m->shards[i];
    cudaSetDevice(b->device);
    if (!b->chunks.empty() && sslsim_join_chunks(b)) return SSLSIM_E_CUDA;
    if (!sslsim_check(b, cudaEventRecord(m->ev0[i], b->stream), "timer start")) return SSLSIM_E_CUDA;
  }
  return SSLSIM_OK;
}

int world_multi_timer_stop(struct WorldMulti *m, float *elapsed_ms_max, float *per_shard_ms) {
  if (!m) return SSLSIM_E_ARG;
  for (size_t i = 0; i < m->shards.size(); ++i) {
    WorldBatch *b = m->shards[i];
    cudaSetDevice(b->device);
    if (!b->chunks.empty() && sslsim_join_chunks(b)) return SSLSIM_E_CUDA;
    if (!sslsim_check(b, cudaEventRecord(m->ev1[i], b->stream), "timer stop")) return SSLSIM_E_CUDA;
  }
  float mx = 0.0f;
  for (size_t i = 0; i < m->shards.size(); ++i) {
    WorldBatch *b = m->shards[i];
    cudaSetDevice(b->device);
    float ms = 0.0f;
    if (!sslsim_check(b, cudaEventSynchronize(m->ev1[i]), "timer sync") ||
        !sslsim_check(b, cudaEventElapsedTime(&ms, m->ev0[i], m->ev1[i]), "timer elapsed")) return SSLSIM_E_CUDA;
    if (per_shard_ms) per_shard_ms[i] = ms;
    mx = ms > mx ? ms : mx;
  }
  if (elapsed_ms_max) *elapsed_ms_max = mx;
  return SSLSIM_OK;
}

const char *world_multi_error_string(const struct WorldMulti *m) { return m ? m->errstr : "null multi"; }
const char *sslsim_multi_nccl_version(void) { nccl_load(); return g_nccl.version; }

} /* extern "C" */
