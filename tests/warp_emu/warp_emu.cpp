/*
 * tests/warp_emu/warp_emu.cpp -- TEST HARNESS ONLY (see cuda_runtime.h in this directory).
 * Coroutine scheduler for one warp: 32 lanes, each on its own stack, switched by a few lines of x86-64 assembly;
 * a collective is a rendezvous of every live lane.  One block = one warp (the kernel is built that way).
 */
#include "cuda_runtime.h"
#include <thread>
#include <vector>

#if !defined(__x86_64__)
#error "warp_emu needs x86-64 (hand-written context switch)"
#endif

extern "C" void emu_ctx_switch(void **save_sp, void *load_sp);
__asm__(
    ".text\n.globl emu_ctx_switch\n.type emu_ctx_switch,@function\nemu_ctx_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n  ret\n"
    ".size emu_ctx_switch,.-emu_ctx_switch\n");

namespace emu {
thread_local emu_dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;

constexpr size_t STACK_BYTES = 256 * 1024;
struct Lane { void *sp; char *stack; bool done; };
struct Warp {
  Lane lanes[32];
  void *main_sp;
  int cur, alive, arrived;
  unsigned gen;
  uint64_t slot[2][32];
  int tag[32];
  unsigned long long nops[32];
  const std::function<void()> *body;
};
static thread_local Warp *g_w = nullptr;

int lane_id() { return g_w->cur; }

static void yield_to_main() {
  Warp &w = *g_w;
  emu_ctx_switch(&w.lanes[w.cur].sp, w.main_sp);
}

void rendezvous(int tag) {
  Warp &w = *g_w;
  const int me = w.cur;
  w.tag[me] = tag;
  w.nops[me]++;
  const unsigned my_gen = w.gen;
  if (++w.arrived == w.alive) {
    /* last to arrive: every lane must be at the same collective, the same number of collectives in */
    for (int i = 0; i < 32; ++i)
      if (!w.lanes[i].done && (w.tag[i] != tag || w.nops[i] != w.nops[me])) {
        fprintf(stderr, "warp_emu: lanes diverged at a collective (lane %d tag %d op %llu vs lane %d tag %d op %llu)\n",
                me, tag, w.nops[me], i, w.tag[i], w.nops[i]);
        abort();
      }
    w.arrived = 0;
    w.gen++;
    return;
  }
  while (w.gen == my_gen) yield_to_main();
}

uint64_t exchange(int tag, uint64_t mine, int src_lane) {
  Warp &w = *g_w;
  const int par = (int)(w.nops[w.cur] & 1u);
  w.slot[par][w.cur] = mine;
  rendezvous(tag);
  return w.slot[par][src_lane & 31];
}

void gather(int tag, uint64_t mine, uint64_t out[32]) {
  Warp &w = *g_w;
  const int par = (int)(w.nops[w.cur] & 1u);
  w.slot[par][w.cur] = mine;
  rendezvous(tag);
  for (int i = 0; i < 32; ++i) out[i] = w.lanes[i].done ? 0 : w.slot[par][i];
}

static void lane_entry() {
  Warp &w = *g_w;
  (*w.body)();
  Lane &l = w.lanes[w.cur];
  l.done = true;
  w.alive--;
  if (w.alive > 0 && w.arrived == w.alive) { /* the others were all waiting for this lane */
    fprintf(stderr, "warp_emu: a lane left the kernel while the others wait at a collective\n");
    abort();
  }
  for (;;) yield_to_main();
}

static void run_block(Warp &w, const std::function<void()> &body) {
  g_w = &w;
  w.body = &body;
  w.alive = 32; w.arrived = 0; w.gen = 0;
  for (int i = 0; i < 32; ++i) {
    Lane &l = w.lanes[i];
    l.done = false; w.tag[i] = 0; w.nops[i] = 0;
    /* initial frame: six callee-saved registers, then the return address = lane_entry; the stack pointer is
     * 16-byte aligned + 8 at function entry as the ABI wants */
    uintptr_t top = ((uintptr_t)(l.stack + STACK_BYTES)) & ~(uintptr_t)15;
    void **sp = (void **)(top - 8);
    *--sp = (void *)lane_entry;
    for (int k = 0; k < 6; ++k) *--sp = nullptr;
    l.sp = sp;
  }
  while (w.alive > 0)
    for (int i = 0; i < 32; ++i) {
      if (w.lanes[i].done) continue;
      w.cur = i;
      t_threadIdx.x = (unsigned)i;
      emu_ctx_switch(&w.main_sp, w.lanes[i].sp);
    }
}

void launch(unsigned grid, unsigned block, const std::function<void()> &body, int os_threads) {
  if (block != 32) { fprintf(stderr, "warp_emu: blockDim.x must be 32\n"); abort(); }
  if (os_threads < 1) os_threads = 1;
  if ((unsigned)os_threads > grid) os_threads = (int)grid;
  auto worker = [&](unsigned b0, unsigned b1) {
    Warp *w = new Warp();
    for (int i = 0; i < 32; ++i) w->lanes[i].stack = (char *)malloc(STACK_BYTES);
    t_blockDim = {block, 1, 1}; t_gridDim = {grid, 1, 1};
    t_threadIdx = {0, 0, 0};
    for (unsigned b = b0; b < b1; ++b) {
      t_blockIdx = {b, 0, 0};
      run_block(*w, body);
    }
    for (int i = 0; i < 32; ++i) free(w->lanes[i].stack);
    delete w;
  };
  if (os_threads == 1) { worker(0, grid); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < os_threads; ++t)
    th.emplace_back(worker, (unsigned)((unsigned long long)grid * t / os_threads),
                    (unsigned)((unsigned long long)grid * (t + 1) / os_threads));
  for (auto &t : th) t.join();
}
} // namespace emu
