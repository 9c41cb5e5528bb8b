/*
 * tests/warp_emu/cuda_runtime.h -- TEST HARNESS ONLY.  NOT part of the product, never linked into
 * libsslsim_b200.so, never importable from ssl-sim_b200/.
 *
 * A stand-in for <cuda_runtime.h> that lets g++ compile ssl-sim_b200/csrc/step_kernel.cu *unmodified* for the
 * host so that the CPU-only test-suite (-m "not gpu", no GPU in the development container) can execute the
 * kernel's own warp-synchronous source -- the shuffle broadphase, the level schedule, the register rows, the
 * general path -- and compare it bit-for-bit with the oracle before any GPU time is spent.  The 32 lanes of a warp
 * run as 32 coroutines on one OS thread; every warp collective (__shfl_sync, __ballot_sync, __any_sync,
 * __reduce_*_sync, __syncwarp, __syncthreads) is a rendezvous of all lanes, and the harness aborts if the lanes of
 * a warp do not arrive at the same collective (a divergence bug that would hang or corrupt on the GPU).
 * It is a checker of the kernel's logic and arithmetic, not an execution path: the product has no CPU fallback.
 */
#ifndef SSLSIM_WARP_EMU_CUDA_RUNTIME_H
#define SSLSIM_WARP_EMU_CUDA_RUNTIME_H
#ifndef SSLSIM_WARP_EMU
#error "tests/warp_emu/cuda_runtime.h is only for the warp-emulation test build (-DSSLSIM_WARP_EMU)"
#endif
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __shared__ static thread_local
#define __launch_bounds__(...)
#define __maxnreg__(...)

struct float4 { float x, y, z, w; };
struct uint4 { unsigned x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { float4 r = {x, y, z, w}; return r; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 r = {x, y, z, w}; return r; }
struct emu_dim3 { unsigned x, y, z; };
typedef int cudaError_t;
enum { cudaSuccess = 0 };
typedef void *cudaStream_t;
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }

namespace emu {
struct Warp;
extern thread_local emu_dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
/* run `body` once per thread of every block of the grid (blockDim.x must be 32: one warp per block, as the
 * kernel is built); blocks are distributed over `os_threads` OS threads */
void launch(unsigned grid, unsigned block, const std::function<void()> &body, int os_threads);
void rendezvous(int tag);                    /* all live lanes of the warp meet here */
uint64_t exchange(int tag, uint64_t mine, int src_lane); /* every lane publishes `mine`, gets src_lane's */
void gather(int tag, uint64_t mine, uint64_t out[32]);   /* every lane publishes `mine`, gets all 32 */
int lane_id();
}
#define threadIdx (emu::t_threadIdx)
#define blockIdx (emu::t_blockIdx)
#define blockDim (emu::t_blockDim)
#define gridDim (emu::t_gridDim)

enum { EMU_TAG_SYNC = 1, EMU_TAG_SHFL, EMU_TAG_SHFL_UP, EMU_TAG_BALLOT, EMU_TAG_ANY, EMU_TAG_RED_OR, EMU_TAG_RED_MAX, EMU_TAG_RED_ADD };

static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __fmaf_rn(float a, float b, float c) { return __builtin_fmaf(a, b, c); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline void __trap() { fprintf(stderr, "warp_emu: __trap()\n"); abort(); }
static inline long long clock64() { return 0; }

static inline void __syncwarp(unsigned = 0xffffffffu) { emu::rendezvous(EMU_TAG_SYNC); }
static inline void __syncthreads() { emu::rendezvous(EMU_TAG_SYNC); }

template <class T> static inline uint64_t emu_pack(T v) { uint64_t u = 0; static_assert(sizeof(T) <= 8, ""); memcpy(&u, &v, sizeof(T)); return u; }
template <class T> static inline T emu_unpack(uint64_t u) { T v; memcpy(&v, &u, sizeof(T)); return v; }

template <class T> static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
  const int me = emu::lane_id();
  const int lane = (me & ~(width - 1)) | (src & (width - 1));
  return emu_unpack<T>(emu::exchange(EMU_TAG_SHFL, emu_pack(v), lane));
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned delta, int width = 32) {
  const int me = emu::lane_id();
  const int in_seg = me & (width - 1);
  const int lane = in_seg >= (int)delta ? me - (int)delta : me;
  return emu_unpack<T>(emu::exchange(EMU_TAG_SHFL_UP, emu_pack(v), lane));
}
static inline unsigned __ballot_sync(unsigned, bool p) {
  uint64_t all[32];
  emu::gather(EMU_TAG_BALLOT, p ? 1u : 0u, all);
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= (unsigned)(all[i] & 1u) << i;
  return r;
}
static inline bool __any_sync(unsigned m, bool p) { return __ballot_sync(m, p) != 0u; }
static inline unsigned __reduce_or_sync(unsigned, unsigned v) {
  uint64_t all[32];
  emu::gather(EMU_TAG_RED_OR, v, all);
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= (unsigned)all[i];
  return r;
}
static inline int __reduce_max_sync(unsigned, int v) {
  uint64_t all[32];
  emu::gather(EMU_TAG_RED_MAX, emu_pack(v), all);
  int r = emu_unpack<int>(all[0]);
  for (int i = 1; i < 32; ++i) { const int t = emu_unpack<int>(all[i]); r = t > r ? t : r; }
  return r;
}
static inline unsigned __reduce_add_sync(unsigned, unsigned v) {
  uint64_t all[32];
  emu::gather(EMU_TAG_RED_ADD, v, all);
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r += (unsigned)all[i];
  return r;
}
template <class T> static inline T atomicAdd(T *p, T v) { T o = *p; *p += v; return o; }

/* seeds of the branch-free sqrt / division sequences (MUFU.RSQ / MUFU.RCP on the GPU): any seed within the
 * hardware's error bound gives the same correctly rounded final result; the emulation uses the rounded exact value */
static inline float emu_rsqrt_approx(float x) { return (float)(1.0 / sqrt((double)x)); }
static inline float emu_rcp_approx(float x) { return (float)(1.0 / (double)x); }
#endif
