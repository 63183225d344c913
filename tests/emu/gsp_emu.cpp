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


// ---------------------------------------------------------------- stream-dependency recorder (see gsp_emu.h)
#include <map>
#include <string>
namespace emu {
namespace {
struct DepAcc { const void* base; long long r0, r1, c0, c1; bool write; };
struct DepNode { std::string name; void* stream; std::vector<int> deps; std::vector<DepAcc> acc; };
struct DepState {
  bool on = false;
  std::vector<DepNode> nodes;
  std::map<void*, int> last;                   // last node of a stream
  std::map<void*, std::vector<int>> pending;   // nodes the stream waited on since its last node
  std::map<void*, int> event_node;
  std::vector<int> host_deps;                  // every later node follows these (host synchronisations)
  std::vector<DepAcc> staged;
  uintptr_t next_stream = 16;
};
DepState D;
int dep_node(void* stream, const char* name) {
  DepNode n;
  n.name = name;
  n.stream = stream;
  auto it = D.last.find(stream);
  if (it != D.last.end()) n.deps.push_back(it->second);
  for (int d : D.pending[stream]) n.deps.push_back(d);
  D.pending[stream].clear();
  for (int d : D.host_deps) n.deps.push_back(d);
  D.nodes.push_back(std::move(n));
  D.last[stream] = (int)D.nodes.size() - 1;
  return (int)D.nodes.size() - 1;
}
}  // namespace

void dep_enable(bool on) {
  D.on = on;
  D.nodes.clear(); D.last.clear(); D.pending.clear(); D.event_node.clear(); D.host_deps.clear(); D.staged.clear();
}
bool dep_enabled() { return D.on; }
void* dep_new_stream() { D.next_stream += 16; return (void*)D.next_stream; }
void dep_access(const void* base, long long r0, long long r1, long long c0, long long c1, bool write) {
  if (D.on && r1 > r0 && c1 > c0) D.staged.push_back({base, r0, r1, c0, c1, write});
}
void dep_launch(void* stream, const char* name) {
  if (!D.on) return;
  const int id = dep_node(stream, name);
  D.nodes[id].acc = std::move(D.staged);
  D.staged.clear();
}
void dep_event_record(void* ev, void* stream) {
  if (!D.on) return;
  D.event_node[ev] = dep_node(stream, "(event)");   // a marker node: carries the stream's pending waits too
}
void dep_stream_wait(void* stream, void* ev) {
  if (!D.on) return;
  auto it = D.event_node.find(ev);
  if (it != D.event_node.end()) D.pending[stream].push_back(it->second);
}
void dep_host_sync(void* stream) {
  if (!D.on) return;
  if (stream == nullptr) {
    for (auto& kv : D.last) D.host_deps.push_back(kv.second);
  } else {
    auto it = D.last.find(stream);
    if (it != D.last.end()) D.host_deps.push_back(it->second);
  }
  // keep the list short: a host sync on X supersedes X's earlier entries transitively, but duplicates are harmless
  if (D.host_deps.size() > 4096) D.host_deps.erase(D.host_deps.begin(), D.host_deps.begin() + 2048);
}
long long dep_check(int verbose) {
  const size_t n = D.nodes.size(), words = (n + 63) / 64;
  std::vector<std::vector<uint64_t>> anc(n, std::vector<uint64_t>(words, 0));
  for (size_t j = 0; j < n; ++j)
    for (int d : D.nodes[j].deps) {
      if (d < 0 || (size_t)d >= j) continue;
      anc[j][(size_t)d / 64] |= 1ull << ((size_t)d % 64);
      for (size_t w = 0; w < words; ++w) anc[j][w] |= anc[(size_t)d][w];
    }
  long long bad = 0;
  for (size_t j = 0; j < n; ++j) {
    if (D.nodes[j].acc.empty()) continue;
    for (size_t i = 0; i < j; ++i) {
      if (D.nodes[i].acc.empty() || (anc[j][i / 64] >> (i % 64) & 1)) continue;
      bool hit = false;
      for (const DepAcc& a : D.nodes[i].acc) {
        for (const DepAcc& b : D.nodes[j].acc)
          if (a.base == b.base && (a.write || b.write) && a.r0 < b.r1 && b.r0 < a.r1 && a.c0 < b.c1 && b.c0 < a.c1) {
            hit = true;
            if (verbose && bad < 20)
              fprintf(stderr, "DEPCHECK: unordered %s (#%zu, stream %p) %s rows [%lld,%lld) cols [%lld,%lld)  vs  %s (#%zu, stream %p) %s rows [%lld,%lld) cols [%lld,%lld) of %p\n",
                      D.nodes[i].name.c_str(), i, D.nodes[i].stream, a.write ? "W" : "R", a.r0, a.r1, a.c0, a.c1, D.nodes[j].name.c_str(), j,
                      D.nodes[j].stream, b.write ? "W" : "R", b.r0, b.r1, b.c0, b.c1, a.base);
            break;
          }
        if (hit) break;
      }
      if (hit) ++bad;
    }
  }
  if (verbose) fprintf(stderr, "DEPCHECK: %zu launches, %lld unordered conflicting pairs\n", n, bad);
  return bad;
}
}  // namespace emu
