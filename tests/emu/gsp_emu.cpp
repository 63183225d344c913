// TEST-ONLY fiber scheduler behind tests/emu/gsp_emu.h (see the header for the rationale).
#include "gsp_emu.h"

namespace emu {

State* g = nullptr;
uint3 g_threadIdx, g_blockIdx;
dim3 g_blockDim, g_gridDim;

static const size_t kStack = 256 * 1024;
static std::vector<char*> stack_pool;

static void fiber_entry() {
  g->body();
  Fiber& f = g->fibers[g->cur];
  f.done = true;
  g->live--;
  swapcontext(&f.ctx, &g->sched);
}

void yield_to_sched() {
  Fiber& f = g->fibers[g->cur];
  swapcontext(&f.ctx, &g->sched);
}

void syncthreads() {
  Fiber& f = g->fibers[g->cur];
  f.wait_kind = 1;
  g->bar_waiting++;
  yield_to_sched();
}

void named_bar_sync(int id, int count) {
  Fiber& f = g->fibers[g->cur];
  int gen = g->named_gen[id];
  g->named_count[id]++;
  if (g->named_count[id] >= count) { g->named_count[id] = 0; g->named_gen[id]++; return; }
  while (g->named_gen[id] == gen) { f.wait_kind = 4; yield_to_sched(); }
  f.wait_kind = 0;
}
void named_bar_arrive(int id, int count) {
  g->named_count[id]++;
  if (g->named_count[id] >= count) { g->named_count[id] = 0; g->named_gen[id]++; }
}

static int live_lanes_in_warp(int w) {
  int n = 0;
  for (int l = 0; l < 32; ++l) {
    int t = w * 32 + l;
    if (t < g->nthreads && !g->fibers[t].done) n++;
  }
  return n;
}

static void warp_rendezvous() {
  Fiber& f = g->fibers[g->cur];
  int w = g->cur / 32;
  int gen = g->warp_gen[w];
  g->warp_arrived[w]++;
  if (g->warp_arrived[w] >= live_lanes_in_warp(w)) { g->warp_arrived[w] = 0; g->warp_gen[w]++; return; }
  while (g->warp_gen[w] == gen) { f.wait_kind = 4; yield_to_sched(); }
  f.wait_kind = 0;
}

void warp_sync() { warp_rendezvous(); }

void warp_exchange(uint64_t v, uint64_t out[32]) {
  int w = g->cur / 32, l = g->cur % 32;
  g->warp_buf[w * 32 + l] = v;
  warp_rendezvous();
  for (int i = 0; i < 32; ++i) out[i] = g->warp_buf[w * 32 + i];
  warp_rendezvous();
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  State st;
  State* saved = g;
  g = &st;
  st.body = body;
  st.nthreads = (int)(block.x * block.y * block.z);
  st.fibers.resize(st.nthreads);
  st.warp_arrived.assign((st.nthreads + 31) / 32, 0);
  st.warp_gen.assign((st.nthreads + 31) / 32, 0);
  st.warp_buf.assign(((st.nthreads + 31) / 32) * 32, 0);
  std::vector<unsigned char> dyn(smem + 16);
  st.dyn_smem = (unsigned char*)(((uintptr_t)dyn.data() + 15) & ~(uintptr_t)15);
  while ((int)stack_pool.size() < st.nthreads) stack_pool.push_back((char*)std::malloc(kStack));
  g_blockDim = block;
  g_gridDim = grid;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        g_blockIdx = uint3{bx, by, bz};
        st.live = st.nthreads;
        st.bar_waiting = 0;
        for (int i = 0; i < 16; ++i) st.named_count[i] = st.named_gen[i] = 0;
        std::fill(st.warp_arrived.begin(), st.warp_arrived.end(), 0);
        for (int t = 0; t < st.nthreads; ++t) {
          Fiber& f = st.fibers[t];
          f.done = false;
          f.wait_kind = 0;
          f.tid = uint3{(unsigned)(t % block.x), (unsigned)((t / block.x) % block.y), (unsigned)(t / (block.x * block.y))};
          getcontext(&f.ctx);
          f.ctx.uc_stack.ss_sp = stack_pool[t];
          f.ctx.uc_stack.ss_size = kStack;
          f.ctx.uc_link = &st.sched;
          makecontext(&f.ctx, (void (*)())fiber_entry, 0);
        }
        long idle_rounds = 0;
        while (st.live > 0) {
          bool progressed = false;
          for (int t = 0; t < st.nthreads; ++t) {
            Fiber& f = st.fibers[t];
            if (f.done || f.wait_kind == 1) continue;
            bool was_spin = (f.wait_kind == 4);
            st.cur = t;
            g_threadIdx = f.tid;
            swapcontext(&st.sched, &f.ctx);
            if (!(was_spin && f.wait_kind == 4)) progressed = true;
          }
          // release the block barrier when every live fiber is parked on it
          if (st.bar_waiting > 0 && st.bar_waiting >= st.live) {
            for (auto& f : st.fibers)
              if (!f.done && f.wait_kind == 1) f.wait_kind = 0;
            st.bar_waiting = 0;
            progressed = true;
          }
          if (!progressed) {
            if (++idle_rounds > 20000) {
              std::fprintf(stderr, "gsp_emu: deadlock in block (%u,%u,%u): live=%d bar_waiting=%d\n", bx, by, bz, st.live, st.bar_waiting);
              std::abort();
            }
          } else {
            idle_rounds = 0;
          }
        }
      }
  g = saved;
}

}  // namespace emu
