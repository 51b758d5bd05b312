// Optional per-kernel timing with CUDA events on the launching stream (bench.py's roofline leg).
// Off by default; when off, prof_begin/prof_end are a single branch.
#include <vector>

#include "common.cuh"

namespace v2f {

static bool g_prof_on = false;
struct Span { cudaEvent_t a, b; };
static std::vector<Span> g_spans[V2F_K_COUNT];
static std::vector<Span> g_free;
static long long g_bytes[V2F_K_COUNT];

static Span take() {
  if (!g_free.empty()) {
    Span s = g_free.back();
    g_free.pop_back();
    return s;
  }
  Span s;
  cudaEventCreate(&s.a);
  cudaEventCreate(&s.b);
  return s;
}

void prof_begin(int id, cudaStream_t st) {
  if (!g_prof_on) return;
  Span s = take();
  cudaEventRecord(s.a, st);
  g_spans[id].push_back(s);
}
void prof_bytes(int id, long long bytes) {
  if (g_prof_on) g_bytes[id] += bytes;
}
void prof_end(int id, cudaStream_t st) {
  if (!g_prof_on) return;
  cudaEventRecord(g_spans[id].back().b, st);
}

}  // namespace v2f

extern "C" int v2f_prof_enable(int on) {
  v2f::g_prof_on = on != 0;
  return V2F_OK;
}

// Sum of the recorded spans of one kernel id since the last read; synchronises on the events.
extern "C" int v2f_prof_read(int id, double* total_ms, long long* launches) {
  using namespace v2f;
  V2F_REQUIRE(id >= 0 && id < V2F_K_COUNT && total_ms && launches, V2F_ERR_BAD_ARG);
  double tot = 0.0;
  for (Span& s : g_spans[id]) {
    cudaEventSynchronize(s.b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, s.a, s.b);
    tot += ms;
    g_free.push_back(s);
  }
  *total_ms = tot;
  *launches = (long long)g_spans[id].size();
  g_spans[id].clear();
  return V2F_OK;
}

extern "C" int v2f_prof_read_bytes(int id, double* total_ms, long long* launches, long long* bytes) {
  V2F_REQUIRE(id >= 0 && id < V2F_K_COUNT && bytes, V2F_ERR_BAD_ARG);
  *bytes = v2f::g_bytes[id];
  v2f::g_bytes[id] = 0;
  return v2f_prof_read(id, total_ms, launches);
}
