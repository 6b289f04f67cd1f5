// TEST INFRASTRUCTURE ONLY: fibre scheduler of the CPU emulation (see tests/emu/cuda_runtime.h).
#include <stdio.h>
#include <time.h>
#include <ucontext.h>

#include <vector>

#include "cuda_runtime.h"

// EMU_TSAN (tests/emu/build_emu.py, tsan=True): every emulated CUDA thread is a ThreadSanitizer fibre.  Fibre switches
// carry NO synchronisation; happens-before edges are added exactly where CUDA gives them: __syncthreads (block-wide),
// __syncwarp (the named lanes), kernel boundaries, and device atomics (cuda_runtime.h maps them to seq_cst atomics).
// TSan then reports two emulated threads touching the same shared or global address without such an edge -- the offline
// stand-in for `compute-sanitizer --tool racecheck` (which only sees shared memory).
#ifdef EMU_TSAN
#include <sanitizer/tsan_interface.h>
#define TSAN_ONLY(x) x
#else
#define TSAN_ONLY(x)
#endif

namespace emu {

Idx g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;

namespace {
constexpr size_t kStack = 256 * 1024;
enum State { kRunnable, kAtBarrier, kAtCollective, kDone };
struct Fibre {
  TSAN_ONLY(void* tsan = nullptr;)
  ucontext_t ctx;
  char* stack = nullptr;
  State st = kDone;
  unsigned barrier_gen = 0;
  // pending warp collective
  unsigned coll_mask = 0;
  unsigned long long coll_val = 0;
  unsigned long long coll_out[32];
};
struct BlockState {
  std::vector<Fibre> f;
  unsigned n = 0, alive = 0, at_barrier = 0, barrier_gen = 0;
  unsigned cur = 0;
  TSAN_ONLY(void* sched_tsan = nullptr;)
  char barrier_obj = 0;      // TSan sync object of __syncthreads
  char warp_obj[32] = {0};   // ... of __syncwarp, one per warp
  ucontext_t sched;
  const std::function<void()>* body = nullptr;
};
BlockState B;
std::vector<char*> g_stacks;
char g_launch_obj = 0;  // TSan sync objects: scheduler -> threads of the next block ...
char g_done_obj = 0;    // ... and exiting threads -> scheduler (two objects: an exiting thread must not order itself
                        // before a thread of the same block that merely starts later)

// D3H_EMU_SHUFFLE=<seed>: blocks run in a random order and every scheduler sweep starts at a random thread, in a random
// direction.  Results must not depend on either (on the GPU both are arbitrary): a shared variable read without the
// barrier that orders it after its writer, or an assumption that lower block indices ran first, shows up as a parity
// failure under some seed.  0 / unset = blocks and threads in index order.
unsigned long long g_rng = 0;
bool g_shuffle = false, g_shuffle_init = false;
unsigned rnd(unsigned n) {  // xorshift64*, uniform enough for shuffling
  g_rng ^= g_rng >> 12; g_rng ^= g_rng << 25; g_rng ^= g_rng >> 27;
  return (unsigned)(((g_rng * 2685821657736338717ull) >> 33) % (n ? n : 1));
}
void init_shuffle() {
  if (g_shuffle_init) return;
  g_shuffle_init = true;
  const char* e = getenv("D3H_EMU_SHUFFLE");
  if (e && atoll(e) != 0) { g_shuffle = true; g_rng = 0x9E3779B97F4A7C15ull ^ (unsigned long long)atoll(e); }
}

void fibre_entry() {
  TSAN_ONLY(__tsan_acquire(&g_launch_obj);)   // everything before the launch happened before this thread
  (*B.body)();
  TSAN_ONLY(__tsan_release(&g_done_obj);)     // ... and this thread happens before whatever follows the block
  B.f[B.cur].st = kDone;
  --B.alive;
  // a thread that exits never reaches the barrier the others wait at: release them if it was the last one missing
  if (B.alive > 0 && B.at_barrier == B.alive) { B.at_barrier = 0; ++B.barrier_gen; }
  TSAN_ONLY(__tsan_switch_to_fiber(B.sched_tsan, __tsan_switch_to_fiber_no_sync);)
  swapcontext(&B.f[B.cur].ctx, &B.sched);
}

void yield_to_scheduler() {
  TSAN_ONLY(__tsan_switch_to_fiber(B.sched_tsan, __tsan_switch_to_fiber_no_sync);)
  swapcontext(&B.f[B.cur].ctx, &B.sched);
}

// completes every collective of warp `w` whose participants have all arrived
void try_complete_collectives(unsigned w) {
  const unsigned base = w * 32;
  const unsigned lanes = (B.n - base) < 32u ? (B.n - base) : 32u;
  for (unsigned l = 0; l < lanes; ++l) {
    Fibre& a = B.f[base + l];
    if (a.st != kAtCollective) continue;
    const unsigned mask = a.coll_mask;
    bool ready = true;
    for (unsigned j = 0; j < 32 && ready; ++j) {
      if (!((mask >> j) & 1u)) continue;
      if (j >= lanes) continue;                       // lanes beyond the block do not exist
      const Fibre& p = B.f[base + j];
      if (p.st == kDone) continue;                    // exited lanes cannot arrive (result for them is undefined)
      if (p.st != kAtCollective || p.coll_mask != mask) ready = false;
    }
    if (!ready) continue;
    unsigned long long vals[32];
    for (unsigned j = 0; j < 32; ++j) {
      const bool in = ((mask >> j) & 1u) && j < lanes && B.f[base + j].st == kAtCollective;
      vals[j] = in ? B.f[base + j].coll_val : 0ull;
    }
    for (unsigned j = 0; j < lanes; ++j) {
      Fibre& p = B.f[base + j];
      if (((mask >> j) & 1u) && p.st == kAtCollective && p.coll_mask == mask) {
        memcpy(p.coll_out, vals, sizeof(vals));
        p.st = kRunnable;
      }
    }
  }
}
}  // namespace

void set_shuffle(unsigned long long seed) {
  g_shuffle_init = true;
  g_shuffle = seed != 0;
  g_rng = 0x9E3779B97F4A7C15ull ^ seed;
}

void block_barrier() {
  Fibre& me = B.f[B.cur];
  me.st = kAtBarrier;
  me.barrier_gen = B.barrier_gen;
  TSAN_ONLY(__tsan_release(&B.barrier_obj);)
  if (++B.at_barrier == B.alive) { B.at_barrier = 0; ++B.barrier_gen; }
  yield_to_scheduler();
  TSAN_ONLY(__tsan_acquire(&B.barrier_obj);)  // every thread of the block released before anybody got here
}

// __syncwarp: memory ordering among the lanes of `mask` (the rendezvous itself is a warp_exchange)
void warp_sync_release() { TSAN_ONLY(__tsan_release(&B.warp_obj[(B.cur >> 5) & 31u]);) }
void warp_sync_acquire() { TSAN_ONLY(__tsan_acquire(&B.warp_obj[(B.cur >> 5) & 31u]);) }

void warp_exchange(unsigned mask, unsigned long long v, unsigned long long out[32]) {
  Fibre& me = B.f[B.cur];
  const unsigned lane = B.cur & 31u;
  if (!((mask >> lane) & 1u)) {
    fprintf(stderr, "emu: lane %u calls a warp collective whose mask %08x does not name it\n", lane, mask);
    abort();
  }
  me.coll_mask = mask;
  me.coll_val = v;
  me.st = kAtCollective;
  try_complete_collectives(B.cur >> 5);
  if (me.st != kRunnable) yield_to_scheduler();
  memcpy(out, me.coll_out, sizeof(me.coll_out));
}

unsigned long long now_ns() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (unsigned long long)ts.tv_sec * 1000000000ull + (unsigned long long)ts.tv_nsec;
}

extern "C" char __start_emu_shared[] __attribute__((weak));
extern "C" char __stop_emu_shared[] __attribute__((weak));

static void run_block(const std::function<void()>& body, unsigned nthreads) {
  // shared memory of a fresh CTA holds arbitrary data on the GPU: never zeros by contract
  if (__start_emu_shared != nullptr && __stop_emu_shared > __start_emu_shared) {
    memset(__start_emu_shared, 0xCB, (size_t)(__stop_emu_shared - __start_emu_shared));
  }
  // Every block has its OWN shared memory, but the emulation re-uses one set of static variables: blocks are therefore
  // ordered one after the other for ThreadSanitizer (block k -> scheduler -> block k+1).  This hides races between
  // DIFFERENT blocks on global memory; races inside a block -- what a missing __syncthreads / __syncwarp causes -- are
  // what this build is for.
  TSAN_ONLY(__tsan_release(&g_launch_obj);)
  if (B.f.size() < nthreads) B.f.resize(nthreads);
  while (g_stacks.size() < nthreads) g_stacks.push_back((char*)malloc(kStack));
  B.n = B.alive = nthreads;
  B.at_barrier = 0;
  B.barrier_gen = 0;
  B.body = &body;
  TSAN_ONLY(B.sched_tsan = __tsan_get_current_fiber();)
  for (unsigned t = 0; t < nthreads; ++t) {
    Fibre& f = B.f[t];
    TSAN_ONLY(if (f.tsan == nullptr) f.tsan = __tsan_create_fiber(0);)
    f.stack = g_stacks[t];
    f.st = kRunnable;
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack;
    f.ctx.uc_stack.ss_size = kStack;
    f.ctx.uc_link = nullptr;
    makecontext(&f.ctx, fibre_entry, 0);
  }
  while (B.alive > 0) {
    bool progressed = false;
    const unsigned off = g_shuffle ? rnd(nthreads) : 0u;
    const bool back = g_shuffle && (rnd(2) == 1u);
    for (unsigned k = 0; k < nthreads; ++k) {
      const unsigned t = back ? (off + nthreads - k) % nthreads : (off + k) % nthreads;
      Fibre& f = B.f[t];
      if (f.st == kAtBarrier && f.barrier_gen != B.barrier_gen) f.st = kRunnable;
      if (f.st == kAtCollective) try_complete_collectives(t >> 5);
      if (f.st != kRunnable) continue;
      B.cur = t;
      g_threadIdx = Idx{t, 0, 0};
      progressed = true;
      TSAN_ONLY(__tsan_switch_to_fiber(f.tsan, __tsan_switch_to_fiber_no_sync);)
      swapcontext(&B.sched, &f.ctx);
    }
    if (!progressed && B.alive > 0) {
      fprintf(stderr, "emu: deadlock in block (%u,%u): %u threads alive, %u at the barrier\n", g_blockIdx.x, g_blockIdx.y,
              B.alive, B.at_barrier);
      abort();
    }
  }
  TSAN_ONLY(__tsan_acquire(&g_done_obj);)
}

void run_grid(dim3 grid, dim3 block, const std::function<void()>& body) {
  if (block.y != 1 || block.z != 1) { fprintf(stderr, "emu: only 1-D blocks are supported\n"); abort(); }
  g_gridDim = Idx{grid.x, grid.y, grid.z};
  g_blockDim = Idx{block.x, block.y, block.z};
  init_shuffle();
  TSAN_ONLY(__tsan_release(&g_launch_obj);)
  struct AcquireAtExit { ~AcquireAtExit() { TSAN_ONLY(__tsan_acquire(&g_done_obj);) } } acquire_at_exit;
  if (g_shuffle) {
    const unsigned long long total = (unsigned long long)grid.x * grid.y * grid.z;
    std::vector<unsigned> order(total);
    for (unsigned long long i = 0; i < total; ++i) order[i] = (unsigned)i;
    for (unsigned long long i = total; i > 1; --i) std::swap(order[i - 1], order[rnd((unsigned)i)]);
    for (unsigned long long i = 0; i < total; ++i) {
      const unsigned b = order[i];
      g_blockIdx = Idx{b % grid.x, (b / grid.x) % grid.y, b / (grid.x * grid.y)};
      run_block(body, block.x);
    }
    return;
  }
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        g_blockIdx = Idx{bx, by, bz};
        run_block(body, block.x);
      }
}

}  // namespace emu

// tests/test_emu_shuffle.py: switch the shuffled schedule on (seed != 0) or off from Python
extern "C" void d3h_emu_set_shuffle(unsigned long long seed) { emu::set_shuffle(seed); }
