// test_weights.cu — host-only check of the work-window arithmetic of k_shared.cuh
// (weight_unrank / weight_of_child / subtree_weight): the device walks exactly
// these functions, so the windows tile the child tasks iff these are consistent.
//   * child intervals are contiguous and in lexicographic order;
//   * weight_unrank inverts weight_of_child anywhere inside an interval and
//     reports the interval's header;
//   * the total equals the sum of the level-0 subtrees;
//   * tail_group_start (the closed form a window uses to step back to the first child of a tail group) agrees
//     with weight_of_child / weight_unrank;
//   * handout_window: the hand-outs of all shards of a launch tile [w_lo, w_hi) exactly once, whole units first
//     and the last round in kFineSplit pieces.
//   * the closed form by which k_shared books the bases of a child's items [0, i) (restated here from the leaf code:
//     the kernel's copy is a lambda over its shared-memory tables) against a walk over the item tables;
//   * the sizes of the survivor stacks against the bound they must hold.
// Exit code 0 = all passed.  Needs no GPU.
#include <cstdio>
#include <algorithm>
#include <cstdlib>
#include <utility>
#include <vector>

#include "../k_shared.cuh"

using namespace enumgpu;

static uint64_t binom(int top, int k)
{
    if (top < 0 || k < 0 || k > top) return 0;
    unsigned __int128 r = 1;
    for (int i = 1; i <= k; ++i) r = r * (unsigned)(top - k + i) / (unsigned)i;
    return (uint64_t)r;
}

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "FAILED %s:%d: %s  (n=%d m=%d)\n", __FILE__, __LINE__, #c, n, m); return 1; } } while (0)

// make_weight_prefix: the table the device descent walks is the running sum of subtree_weight over the valid candidates
static int check_prefix(int n, int m)
{
    auto C = [](int top, int k) -> uint64_t { return binom(top, k); };
    const int P = m - kT;
    const std::vector<uint64_t> ps = make_weight_prefix(C, n, m);
    CHECK((int)ps.size() == P * (n + 1));
    for (int i = 0; i < P; ++i) {
        CHECK(ps[(size_t)i * (n + 1)] == 0);
        for (int v = 0; v < n; ++v) {
            const uint64_t want = v <= n - m + i ? subtree_weight(C, n, m, i, v) : 0;
            CHECK(ps[(size_t)i * (n + 1) + v + 1] - ps[(size_t)i * (n + 1) + v] == want);
        }
    }
    uint64_t total = 0;
    for (int v = 0; v <= n - m; ++v) total += subtree_weight(C, n, m, 0, v);
    CHECK(ps[n - m + 1] == total);
    return 0;
}

static int run(int n, int m)
{
    auto C = [](int top, int k) -> uint64_t { return binom(top, k); };
    const int P = m - kT, Q = P - 2;
    std::vector<int> S(P), T(kMaxM);
    for (int i = 0; i < P; ++i) S[i] = i;
    uint64_t expect = 0, n_children = 0, bases = 0;
    for (;;) {
        // header of this child: first child of its parent (+ first parent of its depth-q node)
        uint32_t hdr = 0;
        const bool first_child = (P >= 2) ? (S[P - 1] == S[P - 2] + 1) : false;
        if (first_child) {
            hdr += kWParent;
            bool first_parent = true;
            for (int j = Q; j < P - 1; ++j) first_parent &= (j == 0) ? false : (S[j] == S[j - 1] + 1);
            if (Q >= 1 && first_parent) hdr += kWNode;
        }
        const uint64_t leaves = binom(n - 1 - S[P - 1], kT);
        const uint32_t wchild = (S[P - 1] >= n - kTailR) ? kWTailChild : kWChild;   // pooled tail children are cheaper
        const uint64_t w = weight_of_child(C, n, m, S.data());
        CHECK(w == expect);
        uint64_t off; uint32_t h;
        const uint64_t probes[3] = {0, (hdr + wchild + leaves) / 2, hdr + wchild + leaves - 1};
        for (uint64_t pr : probes) {
            weight_unrank(C, n, m, w + pr, T.data(), &off, &h);
            for (int i = 0; i < P; ++i) CHECK(T[i] == S[i]);
            CHECK(off == pr);
            CHECK(h == hdr);
        }
        {   // stepping back from this child to the first child of its tail group
            const int t0 = (S[P - 2] + 1 > n - kTailR) ? S[P - 2] + 1 : n - kTailR;
            if (S[P - 1] > t0) {
                uint64_t wg = w; uint32_t hg = 12345;
                tail_group_start(C, n, m, [&](int i) { return S[i]; }, &wg, &hg);
                std::vector<int> G(S);
                G[P - 1] = t0;
                CHECK(wg == weight_of_child(C, n, m, G.data()));
                weight_unrank(C, n, m, wg, T.data(), &off, &h);
                CHECK(off == 0 && h == hg && T[P - 1] == t0);
            }
        }
        expect += hdr + wchild + leaves;
        ++n_children; bases += leaves;
        int i = P - 1;                                         // next valid prefix
        while (i >= 0 && S[i] == n - m + i) --i;
        if (i < 0) break;
        ++S[i];
        for (int j = i + 1; j < P; ++j) S[j] = S[j - 1] + 1;
    }
    uint64_t total = 0;
    for (int v = 0; v <= n - m; ++v) total += subtree_weight(C, n, m, 0, v);
    CHECK(total == expect);
    CHECK(bases == binom(n, m));
    CHECK(n_children == binom(n - kT, P));
    return 0;
}

// every hand-out of every shard, sorted: the windows must tile [w_lo, w_hi)
static int run_handouts(uint64_t w_lo, uint64_t span, uint64_t G, uint32_t shards, uint64_t warps)
{
    const int n = (int)shards, m = (int)warps;        // for CHECK's message
    std::vector<std::pair<uint64_t, uint64_t>> wins;
    const uint64_t nu_all = (span + G - 1) / G;
    uint64_t whole = 0, pieces = 0;
    for (uint32_t sh = 0; sh < shards; ++sh) {
        HandoutPlan hp{};
        hp.unit_weight = G; hp.w_lo = w_lo; hp.w_hi = w_lo + span;
        CHECK(plan_handouts(nu_all, sh, shards, warps, &hp));
        CHECK(hp.n_units >= hp.n_coarse_first && hp.n_handouts == hp.n_coarse_first + (hp.n_units - hp.n_coarse_first) * kFineSplit);
        for (uint64_t k = 0; k < hp.n_handouts; ++k) {
            uint64_t a, b;
            if (!handout_window(hp, k, &a, &b)) continue;
            CHECK(a < b && b - a <= G);
            (k < hp.n_coarse_first ? whole : pieces) += 1;
            wins.emplace_back(a, b);
        }
    }
    std::sort(wins.begin(), wins.end());
    uint64_t at = w_lo;
    for (auto& w : wins) { CHECK(w.first == at); at = w.second; }
    CHECK(at == w_lo + span);
    CHECK(nu_all < 2 * shards || pieces > 0);          // there is a fine tail whenever a shard has 2+ units
    return 0;
}

// Items are colex tuples; item i with largest element z carries R - 1 - z bases (R = candidates of the child, or
// columns of the tail group).  k_shared: bases of the items before i = R C(z,t) - t C(z+1,t+1) + (i - C(z,t)) (R-1-z),
// t = 3 for the triples of a child, 4 for the 4-tuples of a tail group (there with the tables sC3 / sC4 / C(.,5)).
static int check_seen_closed_form()
{
    const int n = 0, m = 0;
    for (int tail = 0; tail < 2; ++tail) {
        const int r_lo = tail ? 5 : 4, r_hi = tail ? kTailR : 64;
        for (int R = r_lo; R <= r_hi; ++R) {
            const std::vector<uint32_t> items = tail ? make_quads(R - 1) : make_triples(R - 1);
            CHECK(items.size() == binom(R - 1, tail ? 4 : 3));
            uint64_t acc = 0;
            for (size_t i = 0; i <= items.size(); ++i) {
                uint64_t closed;
                if (i == items.size()) closed = binom(R, tail ? 5 : 4);          // the kernel's `leaves`
                else if (!tail) {
                    const uint32_t z = (items[i] >> 16) & 255u;
                    closed = (uint64_t)R * binom(z, 3) - 3u * binom(z + 1, 4) + (i - binom(z, 3)) * (uint64_t)(R - 1 - (int)z);
                } else {
                    const uint32_t z = items[i] >> 24;
                    closed = (uint64_t)R * binom(z, 4) - 4u * binom(z + 1, 5) + (i - binom(z, 4)) * (uint64_t)(R - 1 - (int)z);
                }
                CHECK(closed == acc);
                CHECK(closed < (1ull << 32));                                     // the kernel computes it in 32 bits
                if (i < items.size()) {
                    const uint32_t z = tail ? items[i] >> 24 : (items[i] >> 16) & 255u;
                    CHECK((int)z <= R - 2);                                       // every item has at least one basis
                    acc += (uint64_t)(R - 1 - (int)z);
                }
            }
        }
    }
    return 0;
}

// The first-stage survivor stack of a warp: at most 31 entries are left over when a batch of 32 items starts, and a lane
// pushes at most one entry per column it owns — columns c+1 .. n-1 with c >= 4 (the prefix has at least two columns,
// a < b < c above them).  16-byte stores need 16-byte aligned entries at every warp's offset.
static int check_stack_sizes()
{
    const int m = 0;
    for (int n = kSharedMinM; n <= kMaxN; ++n) {
        CHECK(queue1_cap(n) >= 31 + 32 * (n - 5));
        CHECK(queue_warp_bytes(n) % 16 == 0 && kQueue2Bytes % 16 == 0 && kQEntryBytes % 16 == 0);
        CHECK(queue_warp_bytes(n) == (size_t)kQueue2Bytes + (size_t)queue1_cap(n) * kQEntryBytes);
    }
    static_assert(kQueue2Cap >= 31 + 32, "second stage: up to 31 waiting + 32 promoted at once");
    return 0;
}

int main()
{
    if (check_seen_closed_form() || check_stack_sizes()) return 1;
    const uint64_t hcases[][5] = {{0, 1000000, 1024, 1, 16}, {77, 123457, 1024, 8, 4}, {5, 4096, 1024, 3, 64}, {0, 1023, 1024, 2, 8},
                                  {0, 5611770000ull, 21404, 8, 2368}, {1000, 5611770000ull, 21404, 1, 2368}, {0, 99999, 1028, 5, 7}};
    for (auto& h : hcases)
        if (run_handouts(h[0], h[1], h[2] - h[2] % kFineSplit, (uint32_t)h[3], h[4])) return 1;
    const int cases[][2] = {{6, 6}, {9, 6}, {12, 7}, {13, 8}, {16, 9}, {15, 10}, {18, 12}, {20, 16}, {24, 8}, {22, 9}};
    for (auto& c : cases)
        if (run(c[0], c[1]) || check_prefix(c[0], c[1])) return 1;
    std::puts("test_weights: all passed");
    return 0;
}
