// Host-side work split of the persistent layer kernels (pure C++: also compiled by the host
// simulator tools/sim_layout.cpp, which property-checks it).
#pragma once
#include <cstdint>

namespace whvi {

struct Plan {
    int ctas_per_sample;
    int iters_per_group;
};

// Split the tiles of every sample over CTAs: enough CTAs to fill the chip several times
// over, but each group keeps at least `min_iters` tiles so per-CTA setup/epilogue amortise.
inline Plan make_plan(int64_t S, int64_t tiles_per_sample, int groups, int64_t target_ctas, int min_iters)
{
    const int64_t max_ctas = (tiles_per_sample + groups - 1) / groups;
    int64_t ctas = (target_ctas + S - 1) / S;
    if (ctas > max_ctas) ctas = max_ctas;
    if (ctas < 1) ctas = 1;
    int64_t iters = (tiles_per_sample + ctas * groups - 1) / (ctas * groups);
    if (iters < min_iters) iters = min_iters;
    ctas = (tiles_per_sample + iters * groups - 1) / (iters * groups);
    return Plan{static_cast<int>(ctas), static_cast<int>(iters)};
}

// Persistent kernels that run ONE CTA per SM (the backward and the loss layer: shared memory is the
// limit): the grid executes in waves of `slots` CTAs, so the time is ~ waves x iterations per CTA.
// Pick the CTAs-per-sample count that minimises that product (a grid of 608 CTAs on 148 SMs costs a
// fifth wave for 16 CTAs); ties go to the smaller grid (smaller workspace).  At least `min_iters`
// tiles per group unless the sample is smaller than that.
inline Plan make_plan_waves(int64_t S, int64_t tiles_per_sample, int groups, int64_t slots, int min_iters, int max_waves)
{
    // no more CTAs per sample than leaves each group about `min_iters` tiles
    int64_t max_cps = (tiles_per_sample + int64_t(min_iters) * groups - 1) / (int64_t(min_iters) * groups);
    if (max_cps < 1) max_cps = 1;
    int64_t best_cps = 1, best_iters = (tiles_per_sample + groups - 1) / groups, best_cost = -1;
    for (int64_t cps = 1; cps <= max_cps && cps * S <= slots * max_waves; ++cps) {
        const int64_t iters = (tiles_per_sample + cps * groups - 1) / (cps * groups);
        const int64_t eff = (tiles_per_sample + iters * groups - 1) / (iters * groups);  // CTAs actually needed
        const int64_t waves = (eff * S + slots - 1) / slots;
        const int64_t cost = waves * iters;
        if (best_cost < 0 || cost < best_cost) best_cost = cost, best_cps = eff, best_iters = iters;
    }
    return Plan{static_cast<int>(best_cps), static_cast<int>(best_iters)};
}

}  // namespace whvi
