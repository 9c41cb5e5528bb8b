/*
 * aux_kernels.cu -- the cold kernels around K1:
 *   K5 init_worlds       World ctor + world_add_robot + random_robot_pos2
 *                        (reference src/sslsim.cpp:238-252, 358-365, 34-42)
 *   gen_commands         synthetic grSim-style command stream
 *   K2 replace_*         robot_set_pos / ball_set_vec (src/sslsim.cpp:445-450, 384-388)
 *   K3 gather_detections per-object loops of serialize_world (src/serialize.cpp:25-65)
 *   K4 reduce_stats      episode statistics of a batch
 * Compiled with --fmad=false; arithmetic order as in docs/STEP_SPEC.md section 6.
 */
#include "dev_types.h"

namespace {

constexpr float BALL_R = 0.0215f;  /* src/sslsim.cpp:17 */
constexpr float ROBOT_R = 0.09f;   /* :19 */
constexpr float DROP_Z = 0.5f;     /* :22 */
constexpr float WHEEL_R = 0.025f;  /* :23 */
constexpr float TWO_PI = 6.2831855f;

__host__ __device__ inline uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
/* splitmix64-keyed uniform in [0,1), 24 bits (STEP_SPEC 6.1) */
__device__ inline float u01(uint64_t seed, uint64_t a, uint64_t b, uint64_t c) {
  uint64_t h = mix64(seed);
  h = mix64(h ^ a);
  h = mix64(h ^ b);
  h = mix64(h ^ c);
  return (float)(uint32_t)(h >> 40) * 5.9604645e-08f;
}

/* one thread per world (mode 1 places bodies sequentially with rejection) */
__global__ void init_worlds(float4 *pos, float4 *vel, uint32_t *aux, int W, int R, float field_length,
                            float field_width, uint64_t seed, uint64_t world0, int mode) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= W) return;
  const int N = R + 1, B = R;
  const uint64_t wid = world0 + (uint64_t)w;
  float4 *p = pos + (size_t)w * N;
  float4 *v = vel + (size_t)w * N;
  for (int b = 0; b < N; ++b) v[b] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  uint32_t *a = aux + (size_t)w * 8;
  for (int k = 0; k < 8; ++k) a[k] = 0;
  a[SSLSIM_AUX_STATUS] = 0xFFu;
  if (mode == 0) {
    const float wd = field_length / 2 - ROBOT_R, hg = field_width / 2 - ROBOT_R;
    for (int r = 0; r < R; ++r) {
      const float q0 = u01(seed, wid, (uint64_t)r, 0), q1 = u01(seed, wid, (uint64_t)r, 1),
                  q2 = u01(seed, wid, (uint64_t)r, 2);
      p[r] = make_float4(-wd + q0 * (2.0f * wd), -hg + q1 * (2.0f * hg), DROP_Z, q2 * TWO_PI);
    }
    p[B] = make_float4(0.0f, 0.0f, DROP_Z, 0.0f);
  } else {
    const float x0 = field_length / 2 - 1.8f, xw = 1.7f, yw = 1.8f;
    for (int b = 0; b < N; ++b) {
      float x = 0.0f, y = 0.0f;
      for (int t = 0; t < 64; ++t) {
        x = x0 + u01(seed, wid, (uint64_t)b, (uint64_t)(4 * t + 8)) * xw;
        y = -yw + u01(seed, wid, (uint64_t)b, (uint64_t)(4 * t + 9)) * (2.0f * yw);
        bool ok = true;
        for (int c = 0; c < b; ++c) {
          const float dx = x - p[c].x, dy = y - p[c].y;
          if (dx * dx + dy * dy < 0.19f * 0.19f) ok = false;
        }
        if (ok) break;
      }
      p[b] = make_float4(x, y, b == B ? BALL_R : WHEEL_R, b == B ? 0.0f : u01(seed, wid, (uint64_t)b, 2) * TWO_PI);
    }
  }
}

/* one thread per (world, robot) */
__global__ void gen_commands(RobotCommand *cmd, int W, int R, uint64_t seed, uint64_t world0, uint64_t period,
                             int kind) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)W * R) return;
  const int w = (int)(t / R), r = (int)(t % R);
  const uint64_t wid = world0 + (uint64_t)w;
  const uint64_t key = period * 64u + (uint64_t)r;
  RobotCommand c;
  for (int i = 0; i < 4; ++i) c.wheel[i] = -40.0f + u01(seed, wid, key, 100u + (uint64_t)i) * 80.0f;
  c.flags = SSLSIM_CMD_DRIVE | SSLSIM_CMD_WHEELS;
  c.kickspeedx = 0.0f; c.kickspeedz = 0.0f; c.pad = 0;
  if (kind == 1) {
    c.flags |= SSLSIM_CMD_SPINNER;
    c.kickspeedx = u01(seed, wid, key, 104u) * 6.5f;
    c.kickspeedz = (r & 1) ? u01(seed, wid, key, 105u) * 3.0f : 0.0f;
  }
  float4 *o = reinterpret_cast<float4 *>(cmd + t);
  o[0] = make_float4(c.wheel[0], c.wheel[1], c.wheel[2], c.wheel[3]);
  o[1] = make_float4(c.kickspeedx, c.kickspeedz, __uint_as_float(c.flags), 0.0f);
}

/* K2: sparse scatter, one thread per record; later records win (applied in order by one thread
 * per (world, body) is not needed: duplicates are resolved by taking the highest record index). */
__device__ __forceinline__ void clear_warm(uint32_t *warm, int ww, int world) {
  /* an external write of a world's state empties its warm table (include/sslsim_batch.h): the moved body's contact
   * points are gone, the others start their next substep cold.  Zero sig / ground words and a zero pair count are the
   * empty table; the impulse words behind them are dead */
  if (!warm) return;
  uint32_t *t = warm + (size_t)world * ww;
  for (int k = 0; k < ww; ++k) t[k] = 0u;
}

__global__ void replace_robots(float4 *pos, uint32_t *warm, int ww, int W, int R, const RobotReplacement *recs, int n) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const RobotReplacement rec = recs[t];
  if (rec.world < 0 || rec.world >= W || rec.robot < 0 || rec.robot >= R) return;
  for (int u = t + 1; u < n; ++u) /* a later record for the same body supersedes this one */
    if (recs[u].world == rec.world && recs[u].robot == rec.robot) return;
  pos[(size_t)rec.world * (R + 1) + rec.robot] = make_float4(rec.x, rec.y, DROP_Z, rec.dir);
  clear_warm(warm, ww, rec.world);
}

__global__ void replace_balls(float4 *pos, float4 *vel, uint32_t *aux, uint32_t *warm, int ww, int W, int R,
                              const BallReplacement *recs, int n) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const BallReplacement rec = recs[t];
  if (rec.world < 0 || rec.world >= W) return;
  for (int u = t + 1; u < n; ++u)
    if (recs[u].world == rec.world) return;
  const size_t i = (size_t)rec.world * (R + 1) + R;
  pos[i] = make_float4(rec.x, rec.y, DROP_Z, 0.0f);
  vel[i] = make_float4(rec.vx, rec.vy, 0.0f, 0.0f);
  aux[(size_t)rec.world * 8 + SSLSIM_AUX_STATUS] &= SSLSIM_STATUS_KEEP_ON_WAKE; /* reference :387 body.activate(true) */
  clear_warm(warm, ww, rec.world);
}

/* K3: one warp per observed world; record = 6 floats ball + 7 floats per robot */
__global__ void gather_detections(const float4 *pos, int W, int R, const int *world_ids, int n,
                                  const int *robot_ids, float *out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n) return;
  const int rec = 6 + 7 * R;
  float *o = out + (size_t)warp * rec;
  const int w = world_ids[warp];
  if (w < 0 || w >= W) {
    for (int k = lane; k < rec; k += 32) o[k] = 0.0f;
    return;
  }
  const float4 *p = pos + (size_t)w * (R + 1);
  for (int b = lane; b <= R; b += 32) {
    const float4 q = p[b];
    if (b == R) {
      o[0] = 1.0f; o[1] = q.x; o[2] = q.y; o[3] = q.z; o[4] = q.x; o[5] = q.y;
    } else {
      float *r = o + 6 + 7 * b;
      const float yaw = sslsim_readout_yaw(q.w); /* reference src/sslsim.cpp:440-442 */
      r[0] = 1.0f; r[1] = (float)robot_ids[b]; r[2] = q.x; r[3] = q.y; r[4] = yaw; r[5] = q.x; r[6] = q.y;
    }
  }
}

/* referee restart placement (STEP_SPEC 7): one thread per world */
__global__ void referee_restart(float4 *pos, float4 *vel, uint32_t *aux, uint32_t *warm, int ww, int W, int R,
                                float field_length, float field_width) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= W) return;
  uint32_t *a = aux + (size_t)w * 8;
  const uint32_t st = a[SSLSIM_AUX_STATUS];
  const bool goal = ((st >> 12) & 3u) != 0, out = ((st >> 14) & 1u) != 0;
  if (!goal && !out) return;
  const size_t i = (size_t)w * (R + 1) + R;
  float x = 0.0f, y = 0.0f;
  if (!goal) {
    const float lx = field_length / 2 - 0.1f, ly = field_width / 2 - 0.1f;
    const float4 p = pos[i];
    x = p.x < -lx ? -lx : (p.x > lx ? lx : p.x);
    y = p.y < -ly ? -ly : (p.y > ly ? ly : p.y);
  }
  pos[i] = make_float4(x, y, BALL_R, 0.0f);
  vel[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  a[SSLSIM_AUX_STATUS] = st & ~(7u << 8) & SSLSIM_STATUS_KEEP_ON_WAKE; /* edge state cleared; the placed ball is awake */
  a[SSLSIM_AUX_RESTARTS] += 1u;
  clear_warm(warm, ww, w);
}

/* fork: one warp per destination world copies the source world's bodies and aux words */
__global__ void fork_worlds(float4 *pos, float4 *vel, uint32_t *aux, uint32_t *warm, int ww, int W, int R, int src,
                            const int *dst, int n) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n) return;
  const int d = dst[warp];
  if (d < 0 || d >= W || d == src) return;
  const int N = R + 1;
  for (int b = lane; b < N; b += 32) {
    pos[(size_t)d * N + b] = pos[(size_t)src * N + b];
    vel[(size_t)d * N + b] = vel[(size_t)src * N + b];
  }
  if (lane < 8) aux[(size_t)d * 8 + lane] = aux[(size_t)src * 8 + lane];
  if (warm)
    for (int k = lane; k < ww; k += 32) warm[(size_t)d * ww + k] = warm[(size_t)src * ww + k];
}

/* K4: grid-stride block reduction of the aux words into 8 u64 (atomics on the final add) */
__global__ void reduce_stats(const float4 *pos, const float4 *vel, const uint32_t *aux, int W, int R,
                             uint64_t world0, unsigned long long *out8) {
  unsigned long long acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const int N = R + 1;
  for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < W; w += gridDim.x * blockDim.x) {
    const uint4 q0 = *reinterpret_cast<const uint4 *>(aux + (size_t)w * 8);
    const uint4 q1 = *reinterpret_cast<const uint4 *>(aux + (size_t)w * 8 + 4);
    acc[0] += q1.x & 0xFFFFu;
    acc[1] += q1.x >> 16;
    acc[2] += q1.y & 0xFFFFu;
    acc[3] += q1.y >> 16;
    acc[4] += q1.z;
    acc[5] += q1.w;
    acc[6] += (q0.x >> 16) & 3u ? 1u : 0u;
    uint64_t hsh = mix64(world0 + (uint64_t)w);
    for (int b = 0; b < N; ++b) {
      const float4 p = pos[(size_t)w * N + b], v = vel[(size_t)w * N + b];
      hsh = mix64(hsh ^ ((uint64_t)__float_as_uint(p.x) | ((uint64_t)__float_as_uint(p.y) << 32)));
      hsh = mix64(hsh ^ ((uint64_t)__float_as_uint(p.z) | ((uint64_t)__float_as_uint(p.w) << 32)));
      hsh = mix64(hsh ^ ((uint64_t)__float_as_uint(v.x) | ((uint64_t)__float_as_uint(v.y) << 32)));
      hsh = mix64(hsh ^ ((uint64_t)__float_as_uint(v.z) | ((uint64_t)__float_as_uint(v.w) << 32)));
    }
    acc[7] += hsh;
  }
  __shared__ unsigned long long sh[8][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = 0; k < 8; ++k) {
    unsigned long long v = acc[k];
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    if (lane == 0) sh[k][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    unsigned long long v = 0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) v += sh[threadIdx.x][k];
    atomicAdd(out8 + threadIdx.x, v);
  }
}

} // namespace

cudaError_t sslsim_launch_init(float4 *pos, float4 *vel, uint32_t *aux, int W, int R, const FieldGeometry &g,
                               uint64_t seed, uint64_t world0, int mode, cudaStream_t s) {
  if (W <= 0) return cudaSuccess;
  init_worlds<<<(W + 127) / 128, 128, 0, s>>>(pos, vel, aux, W, R, g.field_length, g.field_width, seed, world0, mode);
  return cudaGetLastError();
}

cudaError_t sslsim_launch_gen_commands(RobotCommand *cmd, int W, int R, uint64_t seed, uint64_t world0,
                                       uint64_t period, int kind, cudaStream_t s) {
  const long long n = (long long)W * R;
  if (n <= 0) return cudaSuccess;
  gen_commands<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(cmd, W, R, seed, world0, period, kind);
  return cudaGetLastError();
}

cudaError_t sslsim_launch_replace_robots(float4 *pos, uint32_t *warm, int ww, int W, int R, const RobotReplacement *recs,
                                         int n, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  replace_robots<<<(n + 127) / 128, 128, 0, s>>>(pos, warm, ww, W, R, recs, n);
  return cudaGetLastError();
}

cudaError_t sslsim_launch_replace_balls(float4 *pos, float4 *vel, uint32_t *aux, uint32_t *warm, int ww, int W, int R,
                                        const BallReplacement *recs, int n, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  replace_balls<<<(n + 127) / 128, 128, 0, s>>>(pos, vel, aux, warm, ww, W, R, recs, n);
  return cudaGetLastError();
}

cudaError_t sslsim_launch_gather(const float4 *pos, int W, int R, const int *world_ids, int n,
                                 const int *robot_ids, float *out, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  gather_detections<<<(n + 3) / 4, 128, 0, s>>>(pos, W, R, world_ids, n, robot_ids, out);
  return cudaGetLastError();
}

cudaError_t sslsim_launch_referee(float4 *pos, float4 *vel, uint32_t *aux, uint32_t *warm, int ww, int W, int R,
                                  float field_length, float field_width, cudaStream_t s) {
  if (W <= 0) return cudaSuccess;
  referee_restart<<<(W + 255) / 256, 256, 0, s>>>(pos, vel, aux, warm, ww, W, R, field_length, field_width);
  return cudaGetLastError();
}

cudaError_t sslsim_launch_fork(float4 *pos, float4 *vel, uint32_t *aux, uint32_t *warm, int ww, int W, int R, int src,
                               const int *dst, int n, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  fork_worlds<<<(n + 3) / 4, 128, 0, s>>>(pos, vel, aux, warm, ww, W, R, src, dst, n);
  return cudaGetLastError();
}

cudaError_t sslsim_launch_reduce(const float4 *pos, const float4 *vel, const uint32_t *aux, int W, int R,
                                 uint64_t world0, unsigned long long *out8, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(out8, 0, 8 * sizeof(unsigned long long), s);
  if (e != cudaSuccess || W <= 0) return e;
  int blocks = (W + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  reduce_stats<<<blocks, 256, 0, s>>>(pos, vel, aux, W, R, world0, out8);
  return cudaGetLastError();
}
