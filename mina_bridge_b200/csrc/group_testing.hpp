// Planning of the group testing behind the batched accumulator / IPA checks: pure host logic, no device code.
//
// The device decides, per group of a level, one of three verdicts (k_locate, verifier.cu): every member good,
// exactly one bad member (and which), or "two or more bad".  This file turns verdicts into the next level's groups.
// It is separate so that the CPU test tier can run it against a simulated device (mina_b200_host_group_testing_sim)
// over many corruption patterns: every index must end up with the right bit, in a bounded number of levels.
// What it batches: poly-commitment `batch_dlog_accumulator_check` / `SRS::verify` over many proofs.
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstdlib>
#include <stdexcept>
#include <vector>

namespace pasta {
namespace gt {

// Every group is a contiguous range [a, b) of the batch.  Level 0 is the whole batch, combined in slices of
// COMBINE_SLICE proofs whose partial vectors are KEPT: a later group that is a union of whole slices gets its
// scalar vector by adding slices (k_sum_slices) instead of combining tables again.  If level 0 fails, every
// later level computes TWO sums per group (plain and locator-weighted, see k_locate), so a group with a single
// bad proof is resolved by one test instead of log2(size) halvings; a group with more is split (wide while few
// groups are open, so the launch set fills the GPU; at slice boundaries while groups are large).  The last
// child of every split needs no MSM (A_last = A_parent - siblings).
static constexpr uint32_t COMBINE_SLICE = 64;
// open groups x children per level the splits aim for (MINA_B200_SPLIT_TARGET overrides it: tuning only)
inline uint32_t split_target() {
    static const uint32_t v = [] {
        const char *e = std::getenv("MINA_B200_SPLIT_TARGET");
        int x = e ? std::atoi(e) : 0;
        return (uint32_t)(x >= 2 && x <= 64 ? x : 4);  // 4: best of {2..32} on 1024/10, 128/2, 1024/30, 1024/100 (tools/sweep_split.sh)
    }();
    return v;
}
struct LevelGroup {
    uint32_t a = 0, b = 0;
    uint32_t parent = 0;  // index into the previous level's group list
    uint32_t size() const { return b - a; }
};
struct LevelPlan {
    // groups in result order: MSM groups built from kept slices, MSM groups combined from tables, derived groups
    std::vector<LevelGroup> sliced, combined, derived;
    std::vector<std::array<uint32_t, 3>> derived_meta;  // parent, sibling range in MSM order
    size_t n_msm() const { return sliced.size() + combined.size(); }
    size_t size() const { return n_msm() + derived.size(); }
    const LevelGroup &at(size_t g) const {
        if (g < sliced.size()) return sliced[g];
        g -= sliced.size();
        return g < combined.size() ? combined[g] : derived[g - combined.size()];
    }
};
inline bool slice_aligned(const LevelGroup &g, uint32_t m) { return g.a % COMBINE_SLICE == 0 && (g.b % COMBINE_SLICE == 0 || g.b == m); }


// children of an unresolved group [a, b): `t` parts, cut at slice boundaries while the parts are larger than a slice
inline std::vector<uint32_t> split_points(uint32_t a, uint32_t b, uint32_t t) {
    const uint32_t size = b - a;
    uint32_t part = (size + t - 1) / t;
    if (part > COMBINE_SLICE) part = (part + COMBINE_SLICE - 1) / COMBINE_SLICE * COMBINE_SLICE;
    std::vector<uint32_t> cuts;
    uint32_t first = part;
    if (part > COMBINE_SLICE && a % COMBINE_SLICE) first = part - a % COMBINE_SLICE;  // land on slice boundaries
    for (uint32_t cut = a + first; cut < b; cut += part) cuts.push_back(cut);
    if (cuts.empty()) cuts.push_back(a + (size + 1) / 2);
    return cuts;
}


// Verdicts of one locator level -> per-index bits for the resolved groups and the next level's plan.
// status[g]: 1 = all good, 2 + j = index j is the only bad one, 0 = two or more bad.  Returns false when nothing is open.
inline bool plan_next_level(const LevelPlan &plan, const std::vector<uint32_t> &status, uint32_t m, std::vector<uint8_t> &ok, LevelPlan &next) {
    std::vector<uint32_t> open;
    for (size_t g = 0; g < plan.size(); g++) {
        const LevelGroup &grp = plan.at(g);
        if (status[g] == 1) {
            for (uint32_t i = grp.a; i < grp.b; i++) ok[i] = 1;
        } else if (status[g] >= 2) {
            const uint32_t bad = status[g] - 2;
            if (bad < grp.a || bad >= grp.b) throw std::runtime_error("group testing: locator returned an index outside its group");
            for (uint32_t i = grp.a; i < grp.b; i++) ok[i] = i == bad ? 0 : 1;
        } else if (grp.size() <= 2) {
            for (uint32_t i = grp.a; i < grp.b; i++) ok[i] = 0;  // not "none" and not "exactly one": every member is bad
        } else {
            open.push_back((uint32_t)g);
        }
    }
    if (open.empty()) return false;
    const uint32_t target = split_target();
    const uint32_t t = std::max<uint32_t>(2, std::min<uint32_t>(target, (target + (uint32_t)open.size() - 1) / (uint32_t)open.size()));
    struct Pending {
        uint32_t parent, list, begin, end;  // siblings [begin, end) inside list 0 (sliced) or 1 (combined)
    };
    std::vector<Pending> pend;
    for (uint32_t g : open) {
        const LevelGroup &grp = plan.at(g);
        std::vector<uint32_t> cuts = split_points(grp.a, grp.b, std::min(t, grp.size()));
        // the MSM children of one parent all go to the same list so that they stay adjacent
        bool all_sliced = true;
        uint32_t lo = grp.a;
        for (uint32_t cut : cuts) {
            all_sliced = all_sliced && slice_aligned(LevelGroup{lo, cut, g}, m);
            lo = cut;
        }
        std::vector<LevelGroup> &list = all_sliced ? next.sliced : next.combined;
        Pending pd{g, all_sliced ? 0u : 1u, (uint32_t)list.size(), 0};
        lo = grp.a;
        for (uint32_t cut : cuts) {
            list.push_back(LevelGroup{lo, cut, g});
            lo = cut;
        }
        pd.end = (uint32_t)list.size();
        pend.push_back(pd);
        next.derived.push_back(LevelGroup{lo, grp.b, g});
    }
    for (const Pending &pd : pend) {
        const uint32_t shift = pd.list ? (uint32_t)next.sliced.size() : 0u;
        next.derived_meta.push_back({pd.parent, pd.begin + shift, pd.end + shift});
    }
    return true;
}

// The device replaced by the truth: which verdict k_locate would give a group when `bad` marks the bad indices.
inline uint32_t simulated_verdict(const LevelGroup &g, const std::vector<uint8_t> &bad) {
    uint32_t count = 0, last = 0;
    for (uint32_t i = g.a; i < g.b; i++)
        if (bad[i]) {
            count++;
            last = i;
        }
    return count == 0 ? 1u : count == 1 ? 2u + last : 0u;
}
// The whole procedure of rlc_levels against the simulated device.  Returns the number of levels; msms = MSM
// evaluations over the resident SRS it would have cost (1 at level 0, 2 per MSM group afterwards).
inline uint32_t simulate(uint32_t m, const std::vector<uint8_t> &bad, std::vector<uint8_t> &ok, uint32_t &msms) {
    ok.assign(m, 0);
    msms = 1;
    bool any = false;
    for (uint32_t i = 0; i < m; i++) any = any || bad[i];
    if (!any) {
        ok.assign(m, 1);
        return 1;
    }
    LevelPlan plan;
    plan.sliced.push_back(LevelGroup{0, m, 0});
    uint32_t levels = 1;
    for (;;) {
        levels++;
        if (levels > 64) throw std::runtime_error("group testing: too many levels");
        msms += 2 * (uint32_t)plan.n_msm();
        std::vector<uint32_t> status(plan.size());
        for (size_t g = 0; g < plan.size(); g++) status[g] = simulated_verdict(plan.at(g), bad);
        // structural checks the device path relies on: children tile their parents, groups are disjoint and non-empty
        for (size_t g = 0; g < plan.size(); g++)
            if (plan.at(g).a >= plan.at(g).b || plan.at(g).b > m) throw std::runtime_error("group testing: empty or out-of-range group");
        LevelPlan next;
        if (!plan_next_level(plan, status, m, ok, next)) break;
        if (next.derived.size() != next.derived_meta.size()) throw std::runtime_error("group testing: derived bookkeeping out of step");
        for (size_t d = 0; d < next.derived.size(); d++) {
            const auto &dm = next.derived_meta[d];
            if (dm[1] >= dm[2] || dm[2] > next.n_msm()) throw std::runtime_error("group testing: bad sibling range");
            const LevelGroup &parent = plan.at(dm[0]);
            uint32_t lo = parent.a;
            for (uint32_t sib = dm[1]; sib < dm[2]; sib++) {
                const LevelGroup &c = next.at(sib);
                if (c.a != lo || c.parent != dm[0]) throw std::runtime_error("group testing: siblings do not tile their parent");
                lo = c.b;
            }
            if (next.derived[d].a != lo || next.derived[d].b != parent.b) throw std::runtime_error("group testing: derived child does not close its parent");
        }
        plan = std::move(next);
    }
    return levels;
}

}  // namespace gt
}  // namespace pasta
