/*
 * tests/warp_emu/emu_driver.cpp -- TEST HARNESS ONLY (see cuda_runtime.h in this directory).
 * C entry point for the CPU test-suite: builds the same StepArgs the host layer (ssl-sim_b200/csrc/batch.cu
 * step_args) builds and runs step_kernel.cu's own launch function on the warp emulator.
 */
#include "dev_types.h"

extern "C" int emu_step_worlds(const FieldGeometry *field, const SslSimParams *params, int n_robots, int n_worlds,
                               float *pos, float *vel, uint32_t *aux, const RobotCommand *cmd, int n_steps,
                               int substeps, float h, int lanes, int steps_per_period, long long cmd_period_stride,
                               int os_threads, unsigned *prof, uint32_t *warm) {
  static DevField dfield; /* f_dev points here ("global memory" of the out-of-line general path) */
  StepArgs a;
  memset(&a, 0, sizeof a);
  sslsim_derive_field(field, &dfield);
  a.pos = reinterpret_cast<float4 *>(pos); a.vel = reinterpret_cast<float4 *>(vel); a.aux = aux; a.cmd = cmd; a.warm = warm;
  a.n_worlds = n_worlds; a.n_robots = n_robots; a.n_steps = n_steps; a.substeps = substeps; a.h = h;
  a.steps_per_period = steps_per_period; a.cmd_period_stride = cmd_period_stride;
  a.p = *params;
  sslsim_step_invariants(&a, a.p);
  a.f = dfield; a.f_dev = &dfield; a.prof = prof; /* per-world counters of -DSSLSIM_PROFILE builds, else unused */
  int threads = os_threads;
  return (int)sslsim_launch_step(a, lanes, (cudaStream_t)&threads);
}
