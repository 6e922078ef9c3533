// Compares gd::stdalgo (gd-slam_b200/csrc/stdalgo.cuh, the restatement that runs in one device thread) with the real
// libstdc++ algorithms on randomised inputs with many ties.  Built and run by tests/test_stdalgo_cpu.py (g++, no GPU).
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "stdalgo.cuh"

struct E {
    float r;
    int i;
};

static int check_retain_best(std::mt19937& rng, int n, int levels, int n_points)
{
    std::vector<E> a(n), b;
    for (int i = 0; i < n; ++i) a[i] = {(float)(rng() % levels), i};
    b = a;
    // the library, exactly as cv::KeyPointsFilter::retainBest calls it
    size_t sz_a = a.size();
    if (n_points >= 0 && (int)a.size() > n_points) {
        if (n_points == 0)
            sz_a = 0;
        else {
            std::nth_element(a.begin(), a.begin() + n_points - 1, a.end(), [](const E& x, const E& y) { return x.r > y.r; });
            const float amb = a[n_points - 1].r;
            auto e = std::partition(a.begin() + n_points, a.end(), [amb](const E& x) { return x.r >= amb; });
            sz_a = (size_t)(e - a.begin());
        }
    }
    const int sz_b = gd::stdalgo::retain_best(b.data(), n, n_points, [](const E& x) { return x.r; });
    if ((size_t)sz_b != sz_a) return 1;
    for (size_t i = 0; i < sz_a; ++i)
        if (a[i].i != b[i].i) return 1;
    return 0;
}

static int check_sort(std::mt19937& rng, int n, int levels, int pattern)
{
    std::vector<E> a(n), b;
    for (int i = 0; i < n; ++i) {
        float v = (float)(rng() % levels);
        if (pattern == 1) v = (float)(i % levels);          // saw-tooth
        if (pattern == 2) v = (float)((n - i) / 3);          // descending runs
        if (pattern == 3) v = (float)(i < n / 2 ? i : n - i); // organ pipe (median-of-three killer-ish)
        a[i] = {v, i};
    }
    b = a;
    std::sort(a.begin(), a.end(), [](const E& x, const E& y) { return x.r < y.r; });
    gd::stdalgo::sort(b.data(), b.data() + n, [](const E& x, const E& y) { return x.r < y.r; });
    for (int i = 0; i < n; ++i)
        if (a[i].i != b[i].i) return 1;
    return 0;
}

static int check_sort_prefix(std::mt19937& rng, int n, int levels, int k)
{
    std::vector<E> a(n), b;
    for (int i = 0; i < n; ++i) a[i] = {(float)(rng() % levels), i};
    b = a;
    std::sort(a.begin(), a.end(), [](const E& x, const E& y) { return x.r < y.r; });
    const long done = gd::stdalgo::sort_prefix(b.data(), b.data() + n, (long)k, [](const E& x, const E& y) { return x.r < y.r; });
    if (done < (long)std::min(k, n)) return 1;
    for (long i = 0; i < done; ++i)
        if (a[i].i != b[i].i) return 1;
    return 0;
}

// the data-parallel statement of the Hoare partition (rank_pair_swap) against the two pointer loops
static int check_rank_rule(std::mt19937& rng, int n, int levels)
{
    if (n < 8) return 0;
    std::vector<E> a(n), b;
    for (int i = 0; i < n; ++i) a[i] = {(float)(rng() % levels), i};
    // unguarded partition as introselect calls it: median of three to the front, range [first + 1, last), pivot = *first
    auto less = [](const E& x, const E& y) { return x.r > y.r; };
    gd::stdalgo::move_median_to_first(a.data(), a.data() + 1, a.data() + n / 2, a.data() + n - 1, less);
    b = a;
    const E pivot = a[0];
    E* cut_a = gd::stdalgo::unguarded_partition(a.data() + 1, a.data() + n, a.data(), less);
    const long cut_b = gd::stdalgo::rank_pair_swap(b.data(), 1L, (long)n, [pivot, less](const E& e) { return !less(e, pivot); },
                                                   [pivot, less](const E& e) { return !less(pivot, e); });
    int bad = (cut_a - a.data()) != cut_b;
    for (int i = 0; i < n; ++i) bad += a[i].i != b[i].i;
    // std::partition(first, last, pred)
    std::vector<E> c(n), d;
    for (int i = 0; i < n; ++i) c[i] = {(float)(rng() % levels), i};
    d = c;
    const float amb = (float)(rng() % levels);
    auto pa = std::partition(c.begin(), c.end(), [amb](const E& e) { return e.r >= amb; });
    long npred = 0;
    for (int i = 0; i < n; ++i) npred += d[i].r >= amb;
    gd::stdalgo::rank_pair_swap(d.data(), 0L, (long)n, [amb](const E& e) { return !(e.r >= amb); }, [amb](const E& e) { return e.r >= amb; });
    bad += (pa - c.begin()) != npred;
    for (int i = 0; i < n; ++i) bad += c[i].i != d[i].i;
    return bad ? 1 : 0;
}

// adversarial input that drives introsort / introselect into the heap fallback (depth limit 0): median-of-3 killer sequence
static std::vector<E> killer(int n)
{
    std::vector<E> v(n);
    std::vector<int> key(n);
    // Musser's median-of-three killer for the libstdc++ pivot choice (first+1, mid, last-1) is approximated by running the
    // library's own sort with an adversary comparator (McIlroy's "antiquicksort")
    std::vector<int> val(n, -1);
    int nsolid = 0, candidate = 0;
    const int gas = n;
    std::vector<int> idx(n);
    for (int i = 0; i < n; ++i) idx[i] = i;
    auto cmp = [&](int x, int y) {
        if (val[x] == -1 && val[y] == -1) {
            if (x == candidate)
                val[x] = nsolid++;
            else
                val[y] = nsolid++;
        }
        if (val[x] == -1)
            candidate = x;
        else if (val[y] == -1)
            candidate = y;
        const int vx = val[x] == -1 ? gas : val[x], vy = val[y] == -1 ? gas : val[y];
        return vx < vy;
    };
    std::sort(idx.begin(), idx.end(), cmp);
    for (int i = 0; i < n; ++i) v[i] = {(float)(val[i] == -1 ? gas : val[i]), i};
    return v;
}

extern "C" int gd_stdalgo_selfcheck(int seed, int rounds)
{
    std::mt19937 rng((unsigned)seed);
    int bad = 0;
    for (int r = 0; r < rounds; ++r) {
        const int n = 1 + (int)(rng() % 3000);
        const int levels = 1 + (int)(rng() % (r % 3 == 0 ? 5 : (r % 3 == 1 ? 236 : 100000)));
        const int n_points = (int)(rng() % (n + 50));
        bad += check_retain_best(rng, n, levels, n_points);
        bad += check_sort(rng, n, levels, r % 4);
        bad += check_sort_prefix(rng, n, levels, 100);
        bad += check_rank_rule(rng, n, levels);
        bad += check_rank_rule(rng, n, 1 + (int)(rng() % 3));
        bad += check_sort_prefix(rng, n, 1 + (int)(rng() % 257), 16 + (int)(rng() % 300));
    }
    // sizes around the thresholds (3 / 16) and the FAST-like case: 20 000 small-integer responses, keep 868
    for (int n = 0; n <= 40; ++n)
        for (int np = 0; np <= n + 1; ++np) bad += check_retain_best(rng, n, 3, np) + check_sort(rng, n, 4, 0);
    bad += check_retain_best(rng, 20000, 236, 868);
    bad += check_retain_best(rng, 20000, 40, 868);
    // heap fallback
    for (int n : {200, 1000, 4096}) {
        std::vector<E> a = killer(n), b = a, c = a, d = a;
        std::sort(a.begin(), a.end(), [](const E& x, const E& y) { return x.r < y.r; });
        gd::stdalgo::sort(b.data(), b.data() + n, [](const E& x, const E& y) { return x.r < y.r; });
        for (int i = 0; i < n; ++i) bad += a[i].i != b[i].i;
        std::nth_element(c.begin(), c.begin() + n / 3, c.end(), [](const E& x, const E& y) { return x.r < y.r; });
        gd::stdalgo::nth_element(d.data(), d.data() + n / 3, d.data() + n, [](const E& x, const E& y) { return x.r < y.r; });
        for (int i = 0; i < n; ++i) bad += c[i].i != d[i].i;
        std::vector<E> e = killer(n);
        const long done = gd::stdalgo::sort_prefix(e.data(), e.data() + n, 100L, [](const E& x, const E& y) { return x.r < y.r; });
        for (long i = 0; i < done; ++i) bad += a[i].i != e[i].i;
    }
    return bad;
}

// instrumented: did the killer input really reach the heap fallback of the library's introsort?  (depth 2 lg n exhausted
// <=> more than 2 lg n nested partitions)  Reported so that the test can assert the fallback path is exercised.
extern "C" int gd_stdalgo_killer_depth(int n)
{
    std::vector<E> a = killer(n);
    long cmps = 0;
    std::sort(a.begin(), a.end(), [&](const E& x, const E& y) {
        ++cmps;
        return x.r < y.r;
    });
    return (int)(cmps / n);  // ~ lg n for quicksort behaviour, several times that when the heap sort kicks in
}
