/*
 * sslsim_multi.h -- the batched world-stepper over several GPUs of one box, inside the C ABI.
 *
 * Worlds are independent, so a batch shards by contiguous world range: shard g of G owns worlds
 * [g*W/G, (g+1)*W/G), initial states and synthetic commands are keyed by the GLOBAL world id (a shard of a larger
 * batch is bit-identical to the same worlds of the unsharded batch), and there is no per-step collective.  The one
 * exchange step is the end-of-episode statistics: a per-shard device reduction (K4) followed by ncclAllGather of the
 * 64-byte BatchStats vector over NVLink / NVSwitch.
 *
 * Two ways to drive it, same entry points afterwards:
 *   - one process, several GPUs:  new_world_multi(devices[], n)          (ncclCommInitAll; one stream per GPU, every
 *                                                                           call is asynchronous, one host thread is enough)
 *   - one process per GPU:        sslsim_multi_unique_id on rank 0, ship the 128 bytes to the other ranks by any
 *                                 means, new_world_multi_rank(...) on every rank (ncclCommInitRank)
 * NCCL is bound at run time (dlopen "libnccl.so.2"), so the library loads on hosts without NCCL; only
 * world_multi_allgather_stats (and the constructors, which create the communicators) need it.
 *
 * The reference has no counterpart (one world, one thread, one process: SURVEY 2.1); the per-shard semantics are
 * those of include/sslsim_batch.h, whose calls work on world_multi_shard(m, i) as well.
 */
#ifndef SSLSIM_B200_MULTI_H
#define SSLSIM_B200_MULTI_H
#include "sslsim_batch.h"
#ifdef __cplusplus
extern "C" {
#endif

struct WorldMulti;

#define SSLSIM_E_NCCL (-5)          /* NCCL missing or an NCCL call failed (world_multi_error_string tells which) */
#define SSLSIM_NCCL_ID_BYTES 128    /* sizeof(ncclUniqueId) */

/* contiguous shard of `rank` among `n_ranks`: sizes differ by at most one */
void sslsim_shard_range(long long total_worlds, int rank, int n_ranks, long long *first, int *count);

/* ---- one process, n GPUs (distinct device ordinals) */
struct WorldMulti *new_world_multi(const struct FieldGeometry *field, long long total_worlds, int n_robots,
                                   const int *devices, int n_devices, uint64_t seed, int init_mode,
                                   const struct SslSimParams *params);
/* ---- one process per GPU: rank 0 creates the id, every rank passes the same 128 bytes */
int sslsim_multi_unique_id(void *id_bytes);
struct WorldMulti *new_world_multi_rank(const struct FieldGeometry *field, long long total_worlds, int n_robots,
                                        int rank, int n_ranks, int device, const void *id_bytes, uint64_t seed,
                                        int init_mode, const struct SslSimParams *params);
void delete_world_multi(struct WorldMulti *m);

int world_multi_rank_count(const struct WorldMulti *m);    /* shards in total (all processes) */
int world_multi_local_count(const struct WorldMulti *m);   /* shards owned by this process */
struct WorldBatch *world_multi_shard(struct WorldMulti *m, int local_index);   /* borrowed; every world_batch_* call works */
int world_multi_shard_rank(const struct WorldMulti *m, int local_index);       /* its global shard index */
long long world_multi_world_count(const struct WorldMulti *m);                 /* total worlds */

/* world_batch_step / gen_commands / set_commands / sync on every local shard (asynchronous per GPU stream).
 * host_cmds: the records of ALL worlds, [total_worlds][n_robots]; each shard copies its own rows. */
int world_multi_step(struct WorldMulti *m, int n_steps);
int world_multi_gen_commands(struct WorldMulti *m, uint64_t period, int kind);
int world_multi_set_commands(struct WorldMulti *m, const struct RobotCommand *host_cmds);
int world_multi_sync(struct WorldMulti *m);
/* state of selected worlds by GLOBAL id (ids owned by other processes are skipped and their rows left untouched;
 * returns the number of worlds read) */
int world_multi_read_worlds(struct WorldMulti *m, const long long *world_ids, int n, float *pos, float *vel, uint32_t *aux);

/* K4 + ncclAllGather: per_shard (may be NULL) receives world_multi_rank_count() entries in shard order, total their
 * sum (the checksum is an order-independent wrapping sum, so the total equals the unsharded batch's). */
int world_multi_allgather_stats(struct WorldMulti *m, struct BatchStats *per_shard, struct BatchStats *total);

/* device time of the work queued between start and stop: CUDA events on every local shard's stream, the maximum over
 * the local shards (stop synchronises) */
int world_multi_timer_start(struct WorldMulti *m);
int world_multi_timer_stop(struct WorldMulti *m, float *elapsed_ms_max, float *per_shard_ms /* [local count] or NULL */);

const char *world_multi_error_string(const struct WorldMulti *m);
const char *sslsim_multi_nccl_version(void); /* "2.27.3" style, or "unavailable" */

#ifdef __cplusplus
}
#endif
#endif
