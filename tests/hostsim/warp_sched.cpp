// tests/hostsim/warp_sched.cpp — TEST SCAFFOLDING ONLY (see hostsim.h).
//
// The 32-lane variant of the simulator (-DSMX_SIM_WARP32, libsmatrix_hostsim_warp32.so): every thread
// of a block is a fiber (ucontext); a fiber runs until it reaches a warp collective or a block barrier,
// the scheduler then exchanges the lanes' values exactly as the hardware primitive is specified and
// lets the lanes continue.  This exercises what the 1-lane simulator cannot: the __match_any_sync
// pre-aggregation, shuffle scans, ballots, shared-memory staging behind __syncthreads / __syncwarp.
// Stricter than the hardware in one respect: between two collectives a lane runs alone, so code that
// relies on lanes advancing together WITHOUT a __syncwarp (implicit warp-synchronous programming)
// shows up as a wrong result here; and a collective whose mask names a lane that never arrives is
// reported as a deadlock instead of hanging.
//
// __activemask() (SMX_ACTIVEMASK(key) in the kernels) returns the lanes of the warp that wait at the
// same textual call with the same key at the moment no lane of the block can advance any further —
// the widest convergence the hardware could show.  The key (e.g. the counter an aggregated increment
// goes to) keeps lanes apart that reached the same inline helper from different places: on the GPU
// those are different instructions and never converge.
#include "hostsim.h"
#ifdef SMX_SIM_WARP32
#include <ucontext.h>
#include <sys/mman.h>
#include <cstdio>
#include <vector>

namespace {
enum { ST_RUN = 0, ST_WAIT = 1, ST_DONE = 2 };
struct Lane {
  ucontext_t ctx;
  int st, kind, arg;
  unsigned mask;
  uint64_t in, out;
  const void* site;
};
struct Sched {
  std::vector<Lane> lanes;
  std::vector<char*> stacks;
  ucontext_t main;
  unsigned n = 0, cur = 0;
  void (*entry)(void*) = nullptr;
  void* arg = nullptr;
};
constexpr size_t STACK_BYTES = 256 << 10;
thread_local Sched* S = nullptr;

void lane_main() {
  Sched* s = S;
  s->entry(s->arg);
  s->lanes[s->cur].st = ST_DONE;
  swapcontext(&s->lanes[s->cur].ctx, &s->main);
}

inline bool live(const Sched* s, unsigned l) { return l < s->n && s->lanes[l].st != ST_DONE; }

// one collective of kind `k` over the lanes in `grp` (bit i = lane base + i): fill in every lane's result
void exchange(Sched* s, unsigned base, unsigned grp, int k) {
  uint64_t in[32] = {0};
  for (unsigned i = 0; i < 32; i++)
    if (grp >> i & 1u) in[i] = s->lanes[base + i].in;
  for (unsigned i = 0; i < 32; i++) {
    if (!(grp >> i & 1u)) continue;
    Lane& L = s->lanes[base + i];
    uint64_t out = 0;
    switch (k) {
      case SIM_SHFL: { const unsigned src = (unsigned)L.arg & 31u; out = (grp >> src & 1u) ? in[src] : in[i]; break; }
      case SIM_SHFL_UP: { const unsigned d = (unsigned)L.arg; out = (i >= d && (grp >> (i - d) & 1u)) ? in[i - d] : in[i]; break; }
      case SIM_SHFL_XOR: { const unsigned src = i ^ ((unsigned)L.arg & 31u); out = (grp >> src & 1u) ? in[src] : in[i]; break; }
      case SIM_BALLOT: for (unsigned j = 0; j < 32; j++) if ((grp >> j & 1u) && in[j]) out |= 1ull << j; break;
      case SIM_MATCH: for (unsigned j = 0; j < 32; j++) if ((grp >> j & 1u) && in[j] == in[i]) out |= 1ull << j; break;
      default: break; /* SIM_SYNCWARP */
    }
    L.out = out;
    L.st = ST_RUN;
  }
}

// collectives with an explicit mask: ready when every live lane the mask names waits at the same kind with the same mask
bool release_masked(Sched* s) {
  bool any = false;
  for (unsigned base = 0; base < s->n; base += 32) {
    for (unsigned i = 0; i < 32 && base + i < s->n; i++) {
      Lane& L = s->lanes[base + i];
      if (L.st != ST_WAIT || L.kind == SIM_ACTIVE || L.kind == SIM_BLOCK) continue;
      unsigned grp = 0;
      bool ready = true;
      for (unsigned j = 0; j < 32 && ready; j++) {
        if (!(L.mask >> j & 1u) || !live(s, base + j)) continue;
        const Lane& M = s->lanes[base + j];
        if (M.st == ST_WAIT && M.kind == L.kind && M.mask == L.mask) grp |= 1u << j;
        else ready = false;
      }
      if (ready && (grp >> i & 1u)) {
        exchange(s, base, grp, L.kind);
        any = true;
      }
    }
  }
  return any;
}

bool release_block_barrier(Sched* s) {
  unsigned waiting = 0, alive = 0;
  for (unsigned l = 0; l < s->n; l++) {
    if (s->lanes[l].st == ST_DONE) continue;
    alive++;
    if (s->lanes[l].st == ST_WAIT && s->lanes[l].kind == SIM_BLOCK) waiting++;
  }
  if (!alive || waiting != alive) return false;
  for (unsigned l = 0; l < s->n; l++)
    if (s->lanes[l].st == ST_WAIT) s->lanes[l].st = ST_RUN;
  return true;
}

bool release_activemask(Sched* s) {
  bool any = false;
  for (unsigned base = 0; base < s->n; base += 32) {
    for (unsigned i = 0; i < 32 && base + i < s->n; i++) {
      Lane& L = s->lanes[base + i];
      if (L.st != ST_WAIT || L.kind != SIM_ACTIVE) continue;
      unsigned grp = 0;
      for (unsigned j = 0; j < 32 && base + j < s->n; j++) {
        const Lane& M = s->lanes[base + j];
        if (M.st == ST_WAIT && M.kind == SIM_ACTIVE && M.site == L.site && M.in == L.in) grp |= 1u << j;
      }
      for (unsigned j = 0; j < 32; j++)
        if (grp >> j & 1u) {
          s->lanes[base + j].out = grp;
          s->lanes[base + j].st = ST_RUN;
        }
      any = true;
    }
  }
  return any;
}

[[noreturn]] void deadlock(Sched* s) {
  fprintf(stderr, "hostsim warp32: no lane of block (%u,%u) can advance:\n", blockIdx.x, blockIdx.y);
  static const char* names[] = {"shfl", "shfl_up", "shfl_xor", "ballot", "match_any", "syncwarp", "activemask", "syncthreads"};
  for (unsigned l = 0; l < s->n; l++)
    if (s->lanes[l].st == ST_WAIT)
      fprintf(stderr, "  thread %u waits at %s mask %08x\n", l, names[s->lanes[l].kind], s->lanes[l].mask);
  abort();
}
}  // namespace

uint64_t smx_sim_collective(int kind, unsigned mask, uint64_t in, int arg, const void* site) {
  Sched* s = S;
  Lane& L = s->lanes[s->cur];
  L.kind = kind; L.mask = mask; L.in = in; L.arg = arg; L.site = site;
  L.st = ST_WAIT;
  swapcontext(&L.ctx, &s->main);
  return L.out;
}

unsigned smx_sim_lane() { return S->cur & 31u; }

void smx_sim_run_block(unsigned nthreads, void (*entry)(void*), void* arg) {
  static thread_local Sched sched;
  Sched* s = &sched;
  S = s;
  s->n = nthreads; s->entry = entry; s->arg = arg;
  if (s->lanes.size() < nthreads) s->lanes.resize(nthreads);
  while (s->stacks.size() < nthreads) {
    void* p = mmap(nullptr, STACK_BYTES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_STACK, -1, 0);
    if (p == MAP_FAILED) { perror("hostsim warp32: mmap"); abort(); }
    s->stacks.push_back((char*)p);
  }
  for (unsigned l = 0; l < nthreads; l++) {
    Lane& L = s->lanes[l];
    getcontext(&L.ctx);
    L.ctx.uc_stack.ss_sp = s->stacks[l];
    L.ctx.uc_stack.ss_size = STACK_BYTES;
    L.ctx.uc_link = &s->main;
    makecontext(&L.ctx, lane_main, 0);
    L.st = ST_RUN;
  }
  for (;;) {
    unsigned alive = 0;
    for (unsigned l = 0; l < nthreads; l++) {
      if (s->lanes[l].st == ST_RUN) {
        s->cur = l;
        threadIdx = {l, 0, 0};
        swapcontext(&s->main, &s->lanes[l].ctx); /* back when the lane waits or is done */
      }
      if (s->lanes[l].st != ST_DONE) alive++;
    }
    if (!alive) break;
    if (release_masked(s)) continue;
    if (release_block_barrier(s)) continue;
    if (release_activemask(s)) continue;
    deadlock(s);
  }
}
#endif

#ifdef SMX_SIM_WARP32
// Self-test of the primitives' semantics (called through ctypes by tests/test_hostlogic_warp32.py): one block of
// 70 threads (two full warps and a partial one) checks every collective against its definition.
static int g_bad;
static void selftest_kernel(unsigned* out) {
  const unsigned t = threadIdx.x, lane = t & 31u, warp = t >> 5;
  const unsigned full = 0xffffffffu, nlanes = warp < 2 ? 32u : 6u;
  const unsigned live = nlanes == 32u ? full : (1u << nlanes) - 1u;
  // inclusive prefix sum by shuffles
  unsigned v = lane + 1;
  for (unsigned d = 1; d < 32; d <<= 1) {
    const unsigned u = __shfl_up_sync(full, v, d);
    if (lane >= d) v += u;
  }
  if (v != (lane + 1) * (lane + 2) / 2) g_bad++;
  if (__shfl_sync(full, t, 3) != (t & ~31u) + 3u) g_bad++;
  if (__shfl_xor_sync(full, lane, 1) != (lane ^ 1u) && (lane ^ 1u) < nlanes) g_bad++;
  if (__ballot_sync(full, lane % 3 == 0) != (0x49249249u & live)) g_bad++;
  if (__any_sync(full, lane == 40) != 0) g_bad++;
  const unsigned grp = __match_any_sync(full, (unsigned long long)(lane % 4) << 40);
  if (grp != ((0x11111111u << (lane % 4)) & live)) g_bad++;
  // divergence: odd lanes take one path, even lanes another; each side sees only its own lanes
  unsigned m;
  if (lane & 1u) m = SMX_ACTIVEMASK(0); else m = SMX_ACTIVEMASK(0);
  if (m != ((lane & 1u ? 0xaaaaaaaau : 0x55555555u) & live)) g_bad++;
  // the same call site, different keys: never grouped
  const unsigned k = SMX_ACTIVEMASK(lane / 8);
  if (k != ((0xffu << (lane / 8 * 8)) & live)) g_bad++;
  // shared memory behind a block barrier; lanes that left early do not block the others
  __shared__ unsigned s[70];
  s[t] = t * t;
  __syncthreads();
  if (s[69 - t] != (69 - t) * (69 - t)) g_bad++;
  if (t >= 64) return;
  __syncthreads();
  __syncwarp();
  out[t] = v;
}
extern "C" int smx_sim_warp_selftest(void) {
  static unsigned out[64];
  g_bad = 0;
  memset(out, 0, sizeof out);
  smx_sim_launch(dim3(2), dim3(70), [&] { selftest_kernel(out); });
  for (unsigned t = 0; t < 64; t++)
    if (out[t] != (t % 32 + 1) * (t % 32 + 2) / 2) g_bad++;
  return g_bad;
}
#endif
