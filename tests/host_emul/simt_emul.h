// simt_emul.h -- UNIT-TEST HARNESS: runs the 32 lanes of one warp as coroutines (ucontext) that meet at every
// collective of pacybara_b200/csrc/simt.h.  Lane order is shuffled at every scheduling sweep (seeded), so that
// code which forgets a warp sync between a shared-memory write and a read by another lane fails here.
#pragma once
#include <functional>

namespace simt_emul {
// Runs body(lane) for lanes 0..31 to completion.  Aborts with a message on divergence bugs: lanes waiting at
// different collectives under the same mask, or a collective some named lane never reaches.
void run_warp(const std::function<void(int)> &body, unsigned seed);
}
