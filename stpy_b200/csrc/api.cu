// Library-level entry points: version and device query.
#include "common.cuh"
#include "../../include/stpyb.h"
#include <vector>

namespace stpyb {

long long g_launches = 0;
int g_prof_on = 0;

struct ProfRec { int cat; double flops; cudaEvent_t a, b; };
static std::vector<ProfRec> g_recs;
static std::vector<cudaEvent_t> g_pool;
static size_t g_pool_used = 0;
static ProfRec g_open;
static bool g_has_open = false;

static cudaEvent_t take_event() {
  if (g_pool_used == g_pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    g_pool.push_back(e);
  }
  return g_pool[g_pool_used++];
}

void prof_begin(int cat, double flops, cudaStream_t st) {
  if (!g_prof_on) return;
  g_open.cat = cat;
  g_open.flops = flops;
  g_open.a = take_event();
  g_open.b = take_event();
  cudaEventRecord(g_open.a, st);
  g_has_open = true;
}

void prof_end(cudaStream_t st) {
  if (!g_prof_on || !g_has_open) return;
  cudaEventRecord(g_open.b, st);
  g_recs.push_back(g_open);
  g_has_open = false;
}

}  // namespace stpyb

using namespace stpyb;

extern "C" int stpyb_profile(int enable) {
  g_prof_on = enable ? 1 : 0;
  g_recs.clear();
  g_pool_used = 0;
  g_has_open = false;
  g_launches = 0;
  return 0;
}

extern "C" int stpyb_profile_read(double* out, long long* launches) {
  // out[cat*3 + {0,1,2}] = { milliseconds, flops, timed launches } ; caller synchronises first
  for (int i = 0; i < PROF_NCAT * 3; ++i) out[i] = 0.0;
  for (const ProfRec& r : g_recs) {
    float ms = 0.f;
    cudaError_t e = cudaEventElapsedTime(&ms, r.a, r.b);
    if (e != cudaSuccess) return STPYB_ERR_CUDA + (int)e;
    out[r.cat * 3 + 0] += (double)ms;
    out[r.cat * 3 + 1] += r.flops;
    out[r.cat * 3 + 2] += 1.0;
  }
  if (launches) *launches = g_launches;
  return 0;
}

extern "C" int stpyb_version(void) { return 100; }
