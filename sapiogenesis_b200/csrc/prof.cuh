// prof.cuh -- optional per-stage device timing (CUDA events on the launching stream) and a count of our kernel launches.
// Off by default; bench.py switches it on for the timed region to obtain the dominant kernel's average duration.
#pragma once
#include <cuda_runtime.h>
#include <vector>

enum ProfTag {
    PT_RULES_BEGIN = 0, PT_GAS_SCAN, PT_GRID, PT_THINK, PT_RULES_MAIN, PT_TRAIN, PT_SETTLE, PT_REPRODUCE,
    PT_BROADPHASE, PT_NARROW, PT_ORG_SOLVE, PT_COMP_PLAN, PT_COMPONENTS, PT_COUPLED, PT_STEP_FINISH, PT_FRICTION, PT_POST, PT_FRAME, PT_END, PT_COUNT
};
static const char *const PROF_NAMES[PT_COUNT] = {
    "rules_begin", "gas_scan", "grid", "think", "rules_main", "train", "settle", "reproduce",
    "broadphase", "narrow", "org_solve", "comp_plan", "components", "coupled", "step_finish", "friction", "post", "frame", "end"
};

struct ProfState {
    bool on = false;
    std::vector<cudaEvent_t> ev;
    std::vector<int> tag;
    size_t used = 0;
    double ms[PT_COUNT] = {0};
    long count[PT_COUNT] = {0};
    unsigned long long launches = 0;
};
static ProfState g_prof;

static inline void prof_mark(cudaStream_t st, int tag) {
    if (!g_prof.on) return;
    if (g_prof.used == g_prof.ev.size()) {
        if (g_prof.ev.size() >= (1u << 20)) return;
        cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) return;
        g_prof.ev.push_back(e); g_prof.tag.push_back(0);
    }
    g_prof.tag[g_prof.used] = tag;
    cudaEventRecord(g_prof.ev[g_prof.used], st);
    g_prof.used++;
}
// a mark opens the interval of its tag; the interval ends at the next mark (PT_END closes without opening one)
static inline void prof_collect() {
    if (g_prof.used == 0) return;
    cudaEventSynchronize(g_prof.ev[g_prof.used - 1]);
    for (size_t i = 0; i + 1 < g_prof.used; i++) {
        int t = g_prof.tag[i];
        if (t == PT_END) continue;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, g_prof.ev[i], g_prof.ev[i + 1]) == cudaSuccess) { g_prof.ms[t] += ms; g_prof.count[t]++; }
    }
    g_prof.used = 0;
}
#define LAUNCHES(n) (g_prof.launches += (n))
