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
  g->blocks[f.block].live--;
  swapcontext(&f.ctx, &g->sched);
}

void yield_to_sched() {
  Fiber& f = g->fibers[g->cur];
  swapcontext(&f.ctx, &g->sched);
}

void syncthreads() {
  Fiber& f = g->fibers[g->cur];
  f.wait_kind = 1;
  g->blocks[f.block].bar_waiting++;
  yield_to_sched();
}

void named_bar_sync(int id, int count) {
  Fiber& f = g->fibers[g->cur];
  BlockState& bs = g->blocks[f.block];
  int gen = bs.named_gen[id];
  bs.named_count[id]++;
  if (bs.named_count[id] >= count) { bs.named_count[id] = 0; bs.named_gen[id]++; g->event = true; return; }
  while (bs.named_gen[id] == gen) { f.wait_kind = 4; yield_to_sched(); }
  f.wait_kind = 0;
}
void named_bar_arrive(int id, int count) {
  BlockState& bs = g->blocks[g->fibers[g->cur].block];
  bs.named_count[id]++;
  if (bs.named_count[id] >= count) { bs.named_count[id] = 0; bs.named_gen[id]++; }
}

// warps are numbered globally: block * warps_per_block + warp-in-block (blocks need not be multiples of 32 threads)
static int warp_of(int fiber) {
  const int tpb = g->tpb, nwarp = (tpb + 31) / 32;
  return (fiber / tpb) * nwarp + (fiber % tpb) / 32;
}
static int live_lanes_in_warp(int fiber) {
  const int tpb = g->tpb;
  const int b0 = (fiber / tpb) * tpb, w0 = ((fiber % tpb) / 32) * 32;
  int n = 0;
  for (int l = 0; l < 32; ++l) {
    int t = w0 + l;
    if (t < tpb && !g->fibers[b0 + t].done) n++;
  }
  return n;
}

static void warp_rendezvous() {
  Fiber& f = g->fibers[g->cur];
  int w = warp_of(g->cur);
  int gen = g->warp_gen[w];
  g->warp_arrived[w]++;
  if (g->warp_arrived[w] >= live_lanes_in_warp(g->cur)) { g->warp_arrived[w] = 0; g->warp_gen[w]++; g->event = true; return; }
  while (g->warp_gen[w] == gen) { f.wait_kind = 4; yield_to_sched(); }
  f.wait_kind = 0;
}

void warp_sync() { warp_rendezvous(); }

void warp_exchange(uint64_t v, uint64_t out[32]) {
  int w = warp_of(g->cur), l = (g->cur % g->tpb) % 32;
  g->warp_buf[w * 32 + l] = v;
  warp_rendezvous();
  for (int i = 0; i < 32; ++i) out[i] = g->warp_buf[w * 32 + i];
  warp_rendezvous();
}

// Run `nblk` consecutive blocks of the grid CONCURRENTLY (all their threads are fibers of one scheduler).
// Plain launches use nblk = 1 (blocks one after the other); launch_coop runs the whole grid at once, which is
// what persistent kernels with inter-CTA dependencies (spin-waits on global flags) need.
static void run_blocks(State& st, dim3 grid, dim3 block, size_t smem, unsigned first, unsigned nblk) {
  if (smem > 227 * 1024 || block.x * block.y * block.z > 1024) {
    std::fprintf(stderr, "gsp_emu: launch exceeds the sm_100a limits (dynamic smem %zu bytes, %u threads)\n", smem, block.x * block.y * block.z);
    std::abort();
  }
  const int tpb = (int)(block.x * block.y * block.z);
  const int nwarp = (tpb + 31) / 32;
  st.nthreads = tpb * (int)nblk;
  st.fibers.assign(st.nthreads, Fiber());
  st.warp_arrived.assign(nwarp * nblk, 0);
  st.warp_gen.assign(nwarp * nblk, 0);
  st.warp_buf.assign((size_t)nwarp * nblk * 32, 0);
  st.blocks.assign(nblk, BlockState());
  std::vector<std::vector<unsigned char>> dyn(nblk);
  while ((int)stack_pool.size() < st.nthreads) stack_pool.push_back((char*)std::malloc(kStack));
  for (unsigned b = 0; b < nblk; ++b) {
    dyn[b].resize(smem + 128);
    BlockState& bs = st.blocks[b];
    bs.dyn_smem = (unsigned char*)(((uintptr_t)dyn[b].data() + 127) & ~(uintptr_t)127);
    bs.live = tpb;
    const unsigned lin = first + b;
    bs.bid = uint3{lin % grid.x, (lin / grid.x) % grid.y, lin / (grid.x * grid.y)};
    for (int t = 0; t < tpb; ++t) {
      Fiber& f = st.fibers[b * tpb + t];
      f.done = false;
      f.wait_kind = 0;
      f.block = (int)b;
      f.tid = uint3{(unsigned)(t % block.x), (unsigned)((t / block.x) % block.y), (unsigned)(t / (block.x * block.y))};
      getcontext(&f.ctx);
      f.ctx.uc_stack.ss_sp = stack_pool[b * tpb + t];
      f.ctx.uc_stack.ss_size = kStack;
      f.ctx.uc_link = &st.sched;
      makecontext(&f.ctx, (void (*)())fiber_entry, 0);
    }
  }
  st.live = st.nthreads;
  long idle_rounds = 0;
  while (st.live > 0) {
    bool progressed = st.event;
    st.event = false;
    for (int t = 0; t < st.nthreads; ++t) {
      Fiber& f = st.fibers[t];
      if (f.done || f.wait_kind == 1) continue;
      bool was_spin = (f.wait_kind == 4);
      st.cur = t;
      st.tpb = tpb;
      g_threadIdx = f.tid;
      g_blockIdx = st.blocks[f.block].bid;
      swapcontext(&st.sched, &f.ctx);
      if (!(was_spin && f.wait_kind == 4)) progressed = true;
    }
    // release a block barrier when every live fiber of that block is parked on it
    for (unsigned b = 0; b < nblk; ++b) {
      BlockState& bs = st.blocks[b];
      if (bs.bar_waiting > 0 && bs.bar_waiting >= bs.live) {
        for (int t = 0; t < tpb; ++t) {
          Fiber& f = st.fibers[b * tpb + t];
          if (!f.done && f.wait_kind == 1) f.wait_kind = 0;
        }
        bs.bar_waiting = 0;
        progressed = true;
      }
    }
    if (!progressed) {
      if (++idle_rounds > 20000) {
        std::fprintf(stderr, "gsp_emu: deadlock (blocks %u..%u): live=%d\n", first, first + nblk - 1, st.live);
        std::abort();
      }
    } else {
      idle_rounds = 0;
    }
  }
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  State st;
  State* saved = g;
  g = &st;
  st.body = body;
  g_blockDim = block;
  g_gridDim = grid;
  const unsigned total = grid.x * grid.y * grid.z;
  for (unsigned b = 0; b < total; ++b) run_blocks(st, grid, block, smem, b, 1);
  g = saved;
}

void launch_coop(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  State st;
  State* saved = g;
  g = &st;
  st.body = body;
  g_blockDim = block;
  g_gridDim = grid;
  run_blocks(st, grid, block, smem, 0, grid.x * grid.y * grid.z);
  g = saved;
}

}  // namespace emu
