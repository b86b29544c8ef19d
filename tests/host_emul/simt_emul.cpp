// simt_emul.cpp -- see simt_emul.h.  UNIT-TEST HARNESS, not a product path.
#include "simt_emul.h"

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <ucontext.h>

#include <map>
#include <vector>

#include "simt.h"

namespace {

enum Op { OP_BALLOT = 1, OP_SHFL, OP_MATCH, OP_OR, OP_MAX, OP_ADD, OP_SYNC };

struct Rendezvous {          // one per distinct mask: disjoint or nested lane subsets meet independently, as on the device
  uint32_t mask = 0, arrived = 0;
  int op = 0;
  uint64_t gen = 0;
  uint32_t val[32], aux[32], res[32];
};

constexpr size_t STACK = 256 << 10;
ucontext_t g_sched, g_lane[32];
std::vector<char> g_stacks;
bool g_done[32];
int g_cur = -1;
uint64_t g_progress = 0;
std::map<uint32_t, Rendezvous> g_rv;
const std::function<void(int)> *g_body = nullptr;

void die(const char *msg) {
  fprintf(stderr, "simt_emul: %s (lane %d)\n", msg, g_cur);
  abort();
}

void trampoline() {
  (*g_body)(g_cur);
  g_done[g_cur] = true;
  g_progress++;
  swapcontext(&g_lane[g_cur], &g_sched);
}

uint32_t collective(int op, uint32_t mask, uint32_t v, uint32_t aux) {
  const int lane = g_cur;
  if (!((mask >> lane) & 1u)) die("a lane calls a collective whose mask does not name it");
  Rendezvous &rv = g_rv[mask];
  if (rv.arrived == 0) { rv.mask = mask; rv.op = op; }
  else if (rv.mask != mask || rv.op != op) die("lanes of one mask wait at different collectives (divergence)");
  if ((rv.arrived >> lane) & 1u) die("a lane arrived twice at one collective");
  rv.val[lane] = v; rv.aux[lane] = aux;
  rv.arrived |= 1u << lane;
  const uint64_t my_gen = rv.gen;
  if (rv.arrived == rv.mask) {
    for (int l = 0; l < 32; l++) {
      if (!((mask >> l) & 1u)) continue;
      uint32_t out = 0;
      switch (op) {
        case OP_BALLOT: for (int t = 0; t < 32; t++) if (((mask >> t) & 1u) && rv.val[t]) out |= 1u << t; break;
        case OP_SHFL: {
          const uint32_t src = rv.aux[l] & 31u;
          if (!((mask >> src) & 1u)) die("shfl from a lane outside the mask");
          out = rv.val[src];
          break;
        }
        case OP_MATCH: for (int t = 0; t < 32; t++) if (((mask >> t) & 1u) && rv.val[t] == rv.val[l]) out |= 1u << t; break;
        case OP_OR: for (int t = 0; t < 32; t++) if ((mask >> t) & 1u) out |= rv.val[t]; break;
        case OP_ADD: for (int t = 0; t < 32; t++) if ((mask >> t) & 1u) out += rv.val[t]; break;
        case OP_MAX: {
          int32_t m = INT32_MIN;
          for (int t = 0; t < 32; t++) if (((mask >> t) & 1u) && (int32_t)rv.val[t] > m) m = (int32_t)rv.val[t];
          out = (uint32_t)m;
          break;
        }
        default: break;
      }
      rv.res[l] = out;
    }
    rv.arrived = 0;
    rv.gen++;
    g_progress++;
  } else {
    while (rv.gen == my_gen) swapcontext(&g_lane[lane], &g_sched);
  }
  return rv.res[lane];
}

}  // namespace

namespace simt {
uint32_t ballot(uint32_t mask, bool p) { return collective(OP_BALLOT, mask, p ? 1u : 0u, 0); }
int32_t shfl(uint32_t mask, int32_t v, int src) { return (int32_t)collective(OP_SHFL, mask, (uint32_t)v, (uint32_t)src); }
uint32_t shfl(uint32_t mask, uint32_t v, int src) { return collective(OP_SHFL, mask, v, (uint32_t)src); }
uint32_t match_any(uint32_t mask, int32_t v) { return collective(OP_MATCH, mask, (uint32_t)v, 0); }
uint32_t reduce_or(uint32_t mask, uint32_t v) { return collective(OP_OR, mask, v, 0); }
int32_t reduce_max(uint32_t mask, int32_t v) { return (int32_t)collective(OP_MAX, mask, (uint32_t)v, 0); }
int32_t reduce_add(uint32_t mask, int32_t v) { return (int32_t)collective(OP_ADD, mask, (uint32_t)v, 0); }
void sync(uint32_t mask) { collective(OP_SYNC, mask, 0, 0); }
}  // namespace simt

namespace simt_emul {

void run_warp(const std::function<void(int)> &body, unsigned seed) {
  if (g_stacks.empty()) g_stacks.resize(32 * STACK);
  g_body = &body;
  g_rv.clear();
  for (int l = 0; l < 32; l++) {
    g_done[l] = false;
    getcontext(&g_lane[l]);
    g_lane[l].uc_stack.ss_sp = g_stacks.data() + (size_t)l * STACK;
    g_lane[l].uc_stack.ss_size = STACK;
    g_lane[l].uc_link = nullptr;
    makecontext(&g_lane[l], trampoline, 0);
  }
  uint32_t rng = seed * 2654435761u + 12345u;
  int order[32];
  for (int l = 0; l < 32; l++) order[l] = l;
  for (;;) {
    for (int l = 31; l > 0; l--) {                      // a fresh lane order for every sweep
      rng = rng * 1664525u + 1013904223u;
      const int k = (int)((rng >> 8) % (uint32_t)(l + 1));
      const int tmp = order[l]; order[l] = order[k]; order[k] = tmp;
    }
    const uint64_t before = g_progress;
    bool any = false;
    for (int i = 0; i < 32; i++) {
      const int l = order[i];
      if (g_done[l]) continue;
      any = true;
      g_cur = l;
      swapcontext(&g_sched, &g_lane[l]);
    }
    if (!any) break;
    if (g_progress == before) { g_cur = -1; die("deadlock: lanes wait at a collective that some lane of the mask never reaches"); }
  }
  g_cur = -1;
}

}  // namespace simt_emul
