// Host-side simulator for the view/transposition engine (whvi_b200/csrc/layout.cuh).
// No GPU needed:  g++ -std=c++17 -O2 -I whvi_b200/csrc tools/sim_layout.cpp -o /tmp/sim && /tmp/sim
//
// For every (n, c) configuration the kernels instantiate it checks
//   1. each view is a permutation of the n logical bits, FIRST/LAST are float4 + lane
//      coalesced (reg bits 0,1 = logical 0,1; lanes = logical 2..6);
//   2. the reader-owned physical layout is a bijection of the tile;
//   3. scalar writes of every transposition hit 32 distinct banks per warp instruction
//      and float4 reads cover 8 distinct 16-byte bank groups per quarter-warp;
//   4. the writer's address split (thread part XOR swizzle bits + additive rest) used by
//      engine.cuh::transpose_write reproduces view_phys exactly;
//   5. running the round structure sequentially (load FIRST -> butterflies ->
//      transpose -> MID -> ... -> LAST) on random data equals a plain FWHT over bits
//      [0,k) for every k <= n, for both the 3-view sequence and its reverse.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <set>
#include <vector>
#include "layout.cuh"
#include "plan.cuh"

using namespace whvi;

static int g_fail = 0;
#define CHECK(cond, ...)                                   \
    do {                                                   \
        if (!(cond)) {                                     \
            ++g_fail;                                      \
            std::printf("FAIL %s:%d: ", __FILE__, __LINE__); \
            std::printf(__VA_ARGS__);                      \
            std::printf("\n");                             \
        }                                                  \
    } while (0)

static View get_view_rt(int n, int c, int id)
{
    return id == 0 ? view_first(n, c) : id == 1 ? view_mid(n, c) : id == 2 ? view_last(n, c) : view_mid2(n, c);
}

static void check_view(const View& v, const char* name, bool global_ok)
{
    std::set<int> seen;
    for (int p = 0; p < v.n; ++p) seen.insert(v.bit[p]);
    CHECK((int)seen.size() == v.n && *seen.begin() == 0 && *seen.rbegin() == v.n - 1, "%s n=%d c=%d not a permutation",
          name, v.n, v.c);
    if (global_ok) {
        CHECK(v.bit[0] == 0 && v.bit[1] == 1, "%s n=%d c=%d: reg bits 0,1", name, v.n, v.c);
        for (int i = 0; i < 5; ++i) CHECK(v.bit[v.c + i] == 2 + i, "%s n=%d c=%d lane %d", name, v.n, v.c, i);
    }
}

static uint32_t logical_of(const View& v, uint32_t tid, uint32_t reg) { return view_tid_logical(v, tid) | view_reg_logical(v, reg); }

static void check_transposition(int n, int c, int ida, int idb)
{
    const View a = get_view_rt(n, c, ida), b = get_view_rt(n, c, idb);
    const int E = 1 << c, T = 1 << (n - c), N = 1 << n;
    // bijection
    const uint32_t words = scratch_words(n, c);
    std::vector<char> hit(words, 0);
    for (uint32_t i = 0; i < (uint32_t)N; ++i) {
        uint32_t p = view_phys(b, i);
        CHECK(p < words && !hit[p], "phys not bijective n=%d c=%d", n, c);
        if (p < words) hit[p] = 1;
    }
    // writer: address split + bank conflicts per warp instruction
    for (int tid0 = 0; tid0 < T; tid0 += 32) {
        for (int r = 0; r < E; ++r) {
            std::set<uint32_t> banks;
            for (int lane = 0; lane < 32; ++lane) {
                uint32_t tid = tid0 + lane;
                uint32_t want = view_phys(b, logical_of(a, tid, r));
                uint32_t wbase = 0;
                for (int j = 0; j < n - c; ++j)
                    if ((tid >> j) & 1u) wbase ^= view_phys(b, 1u << a.bit[c + j]);
                uint32_t pr = view_phys(b, view_reg_logical(a, r));
#if WHVI_PADDED
                wbase = 0;
                for (int j = 0; j < n - c; ++j)
                    if ((tid >> j) & 1u) wbase += view_phys(b, 1u << a.bit[c + j]);
                uint32_t got = wbase + pr;
#else
                uint32_t got = (wbase ^ (pr & 0x1Cu)) + (pr & ~0x1Cu);
#endif
                CHECK(got == want, "writer split n=%d c=%d %d->%d tid=%u r=%d got=%u want=%u", n, c, ida, idb, tid, r, got,
                      want);
                banks.insert(want & 31u);
            }
            CHECK(banks.size() == 32, "write bank conflict n=%d c=%d %d->%d r=%d distinct=%zu", n, c, ida, idb, r,
                  banks.size());
        }
    }
    // reader: float4 slot j of thread tid at word (tid<<c) + ((j ^ sw) << 2)
    for (int tid0 = 0; tid0 < T; tid0 += 8) {
        for (int j = 0; j < E / 4; ++j) {
            std::set<uint32_t> groups;
            for (int l = 0; l < 8; ++l) {
                uint32_t tid = tid0 + l;
#if WHVI_PADDED
                uint32_t addr = tid * (E + 4) + 4 * j;
#else
                uint32_t addr = (tid << c) + ((j ^ swz_of_tid(c, tid)) << 2);
#endif
                CHECK(addr == view_phys(b, logical_of(b, tid, 4 * j)), "reader addr n=%d c=%d", n, c);
                groups.insert((addr >> 2) & 7u);
            }
            CHECK(groups.size() == 8, "read bank-group conflict n=%d c=%d view=%d j=%d distinct=%zu", n, c, idb, j,
                  groups.size());
        }
    }
}

// STUDY (not used by the kernels yet): the same transpositions with a
// PADDED reader layout instead of the XOR swizzle -- thread tid' owns E + 4 words at tid' * (E + 4),
// register r at word r.  Every address is then (thread base) + (compile-time constant): no LOP3 per
// STS/LDS.  Checks: bijection into T * (E + 4) words, scalar writes of a warp hit 32 distinct
// banks, float4 reads of a quarter-warp hit 8 distinct 16-byte bank groups, and the writer's
// address splits into a thread part plus a register part additively.
static uint32_t phys_padded(const View& v, uint32_t idx)
{
    uint32_t reg = 0, tid = 0;
    for (int p = 0; p < v.n; ++p) {
        const uint32_t b = (idx >> v.bit[p]) & 1u;
        if (p < v.c) reg |= b << p; else tid |= b << (p - v.c);
    }
    return tid * ((1u << v.c) + 4u) + reg;
}
static int g_padded_fail = 0;
static void study_padded(int n, int c, int ida, int idb)
{
    const View a = get_view_rt(n, c, ida), b = get_view_rt(n, c, idb);
    const int E = 1 << c, T = 1 << (n - c), N = 1 << n;
    std::set<uint32_t> seen;
    for (uint32_t i = 0; i < (uint32_t)N; ++i) seen.insert(phys_padded(b, i));
    bool ok = (int)seen.size() == N && *seen.rbegin() < (uint32_t)(T * (E + 4));
    for (int tid0 = 0; tid0 < T && ok; tid0 += 32)
        for (int r = 0; r < E && ok; ++r) {
            std::set<uint32_t> banks;
            for (int lane = 0; lane < 32; ++lane) {
                const uint32_t tid = tid0 + lane;
                const uint32_t want = phys_padded(b, logical_of(a, tid, r));
                // additive split: thread part (r = 0) + register part (tid = 0)
                ok = ok && want == phys_padded(b, logical_of(a, tid, 0)) + phys_padded(b, view_reg_logical(a, r));
                banks.insert(want & 31u);
            }
            ok = ok && banks.size() == 32;
        }
    for (int tid0 = 0; tid0 < T && ok; tid0 += 8)
        for (int j = 0; j < E / 4 && ok; ++j) {
            std::set<uint32_t> groups;
            for (int l = 0; l < 8; ++l) groups.insert(((uint32_t(tid0 + l) * (E + 4) + 4 * j) >> 2) & 7u);
            ok = ok && groups.size() == 8;
        }
    if (!ok) {
        ++g_padded_fail;
        std::printf("padded-layout study: n=%d c=%d %d->%d NOT conflict-free / additive\n", n, c, ida, idb);
    }
}

static void fwht_bits(std::vector<double>& x, int k)
{
    const size_t N = x.size();
    for (int b = 0; b < k; ++b)
        for (size_t i = 0; i < N; ++i)
            if (!((i >> b) & 1)) {
                double u = x[i], w = x[i | (size_t(1) << b)];
                x[i] = u + w;
                x[i | (size_t(1) << b)] = u - w;
            }
}

// Emulate the engine: views seq[0..len), butterflies on new bits < k, transpositions.
static void check_engine(int n, int c, int k, const int* seq, int len)
{
    const int E = 1 << c, T = 1 << (n - c), N = 1 << n;
    std::vector<double> x(N), ref;
    for (auto& e : x) e = (double)(std::rand() % 2001 - 1000) / 64.0;
    ref = x;
    fwht_bits(ref, k);
    std::vector<std::vector<double>> regs(T, std::vector<double>(E));
    std::vector<double> smem(scratch_words(n, c));
    std::set<int> done;
    for (int round = 0; round < len; ++round) {
        const View v = get_view_rt(n, c, seq[round]);
        if (round == 0) {
            for (int tid = 0; tid < T; ++tid)
                for (int r = 0; r < E; ++r) regs[tid][r] = x[logical_of(v, tid, r)];
        } else {
            const View a = get_view_rt(n, c, seq[round - 1]);
            for (int tid = 0; tid < T; ++tid)
                for (int r = 0; r < E; ++r) smem[view_phys(v, logical_of(a, tid, r))] = regs[tid][r];
            for (int tid = 0; tid < T; ++tid)
                for (int j = 0; j < E / 4; ++j)
                    for (int q = 0; q < 4; ++q)
#if WHVI_PADDED
                        regs[tid][4 * j + q] = smem[tid * (E + 4) + 4 * j + q];
#else
                        regs[tid][4 * j + q] = smem[(tid << c) + ((j ^ swz_of_tid(c, tid)) << 2) + q];
#endif
        }
        for (int p = 0; p < c; ++p) {
            int b = v.bit[p];
            if (b >= k || done.count(b)) continue;
            done.insert(b);
            for (int tid = 0; tid < T; ++tid)
                for (int r = 0; r < E; ++r)
                    if (!((r >> p) & 1)) {
                        double u = regs[tid][r], w = regs[tid][r | (1 << p)];
                        regs[tid][r] = u + w;
                        regs[tid][r | (1 << p)] = u - w;
                    }
        }
    }
    bool covered = true;
    for (int b = 0; b < k; ++b) covered = covered && done.count(b);
    if (!covered) return;  // this sequence does not cover k bits: not used by the kernels
    const View v = get_view_rt(n, c, seq[len - 1]);
    double err = 0;
    for (int tid = 0; tid < T; ++tid)
        for (int r = 0; r < E; ++r) err = std::fmax(err, std::fabs(regs[tid][r] - ref[logical_of(v, tid, r)]));
    CHECK(err == 0.0, "engine mismatch n=%d c=%d k=%d len=%d err=%g", n, c, k, len, err);
}

// Work-split plans: every tile covered, at least one iteration, grids bounded, and the wave-aware
// plan never costs more waves x iterations than the plain one on a 148-SM chip.
static void check_plans()
{
    uint64_t rng = 0x9E3779B97F4A7C15ull;
    auto next = [&]() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; };
    for (int t = 0; t < 20000; ++t) {
        const int64_t S = 1 + next() % (t % 3 == 0 ? 4096 : 160);
        const int64_t tiles = 1 + next() % (t % 2 == 0 ? 70000 : 300);
        const int groups = 1 << (next() % 4);
        const Plan a = make_plan(S, tiles, groups, 148 * 4, 8);
        const Plan b = make_plan_waves(S, tiles, groups, 148, 8, 8);
        auto cost = [&](const Plan& p) { return ((int64_t(p.ctas_per_sample) * S + 147) / 148) * p.iters_per_group; };
        for (const Plan& p : {a, b}) {
            CHECK(p.ctas_per_sample >= 1 && p.iters_per_group >= 1, "plan S=%ld tiles=%ld g=%d: empty", (long)S, (long)tiles, groups);
            CHECK(int64_t(p.ctas_per_sample) * p.iters_per_group * groups >= tiles, "plan S=%ld tiles=%ld g=%d: tiles not covered",
                  (long)S, (long)tiles, groups);
            // no CTA without work: the last CTA's first tile exists
            CHECK((int64_t(p.ctas_per_sample) - 1) * p.iters_per_group * groups < tiles, "plan S=%ld tiles=%ld g=%d: idle CTA",
                  (long)S, (long)tiles, groups);
        }
        // a = make_plan pads to min_iters even when the sample has fewer tiles; compare on real work only
        if (a.iters_per_group * int64_t(groups) <= tiles)
            CHECK(cost(b) <= cost(a), "plan S=%ld tiles=%ld g=%d: waves plan costs %ld > %ld", (long)S, (long)tiles, groups,
                  (long)cost(b), (long)cost(a));
    }
}

int main()
{
    check_plans();
    const int cfgs[][2] = {{10, 5}, {11, 5}, {12, 5}, {13, 5}, {12, 6}, {13, 6}, {14, 6}, {15, 6}, {16, 6}};
    for (auto& cfg : cfgs) {
        const int n = cfg[0], c = cfg[1];
        check_view(view_first(n, c), "first", true);
        check_view(view_mid(n, c), "mid", false);
        check_view(view_last(n, c), "last", true);
        check_transposition(n, c, 0, 1);
        check_transposition(n, c, 1, 2);
        check_view(view_mid2(n, c), "mid2", false);
        check_transposition(n, c, 2, 3);
        check_transposition(n, c, 3, 0);
        if (rounds_needed(n, c, n) == 2) check_transposition(n, c, 1, 0);
        study_padded(n, c, 0, 1);
        study_padded(n, c, 1, 2);
        study_padded(n, c, 2, 3);
        study_padded(n, c, 3, 0);
        if (rounds_needed(n, c, n) == 2) study_padded(n, c, 1, 0);
        const int fwd3[3] = {0, 1, 2}, rev3[3] = {2, 3, 0}, fwd2[2] = {0, 1}, rev2[2] = {1, 0};
        for (int k = 0; k <= n; ++k) {
            check_engine(n, c, k, fwd3, 3);
            check_engine(n, c, k, rev3, 3);
            check_engine(n, c, k, fwd2, 2);
            check_engine(n, c, k, rev2, 2);
            CHECK(rounds_needed(n, c, k) <= 3, "n=%d c=%d k=%d not coverable in 3 rounds", n, c, k);
        }
        std::printf("n=%2d c=%d  T=%4d E=%2d  rounds(k=n)=%d  first=[", n, c, 1 << (n - c), 1 << c, rounds_needed(n, c, n));
        const View f = view_first(n, c), m = view_mid(n, c), l = view_last(n, c);
        for (int p = 0; p < c; ++p) std::printf("%d ", f.bit[p]);
        std::printf("] mid=[");
        for (int p = 0; p < c; ++p) std::printf("%d ", m.bit[p]);
        std::printf("| lanes ");
        for (int p = c; p < c + 5; ++p) std::printf("%d ", m.bit[p]);
        std::printf("] last=[");
        for (int p = 0; p < c; ++p) std::printf("%d ", l.bit[p]);
        std::printf("]\n");
    }
    std::printf("padded-layout study: %s\n", g_padded_fail ? "some transpositions need more than padding" :
                                                            "every transposition is conflict-free and additive with E + 4 padding");
    std::printf(g_fail ? "FAILED: %d checks\n" : "all layout checks passed\n", g_fail);
    return g_fail ? 1 : 0;
}
