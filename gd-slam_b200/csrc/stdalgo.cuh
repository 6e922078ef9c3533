// stdalgo.cuh — libstdc++'s std::nth_element, std::partition and std::sort restated step by step for plain arrays, callable
// from one device thread (and from the host, where tests/test_stdalgo_cpu.py compares them with the real library).
//
// Why: the ORDER of cv::ORB's keypoints is the permutation std::nth_element + std::partition leave behind inside
// KeyPointsFilter::retainBest, and the first 100 matches GeoMaskMaker::GetRt keeps (GD-SLAM src/GeoMaskMaker.cc:95-97) are
// what the unstable std::sort leaves in front.  Both are implementation defined; the reference binary links libstdc++, so
// "identical results" means running exactly its sequence of comparisons and moves (GCC's bits/stl_algo.h, bits/stl_heap.h:
// introselect / introsort with median-of-three to first, unguarded Hoare partition, insertion sort below 4 / 16 elements,
// heap fallback at depth 2 lg n).  Written from the published algorithm; T is a trivially copyable element, Less a functor.
#pragma once

#if defined(__CUDACC__)
#define GD_HD __host__ __device__ __forceinline__
#else
#define GD_HD inline
#endif

namespace gd {
namespace stdalgo {

template <class T>
GD_HD void swap_(T& a, T& b)
{
    const T t = a;
    a = b;
    b = t;
}

GD_HD int lg_(long n)  // std::__lg: floor(log2(n)), n > 0
{
    int k = 0;
    while (n > 1) {
        n >>= 1;
        ++k;
    }
    return k;
}

template <class T, class Less>
GD_HD void move_median_to_first(T* result, T* a, T* b, T* c, Less less)
{
    if (less(*a, *b)) {
        if (less(*b, *c))
            swap_(*result, *b);
        else if (less(*a, *c))
            swap_(*result, *c);
        else
            swap_(*result, *a);
    } else if (less(*a, *c))
        swap_(*result, *a);
    else if (less(*b, *c))
        swap_(*result, *c);
    else
        swap_(*result, *b);
}

template <class T, class Less>
GD_HD T* unguarded_partition(T* first, T* last, T* pivot, Less less)
{
    while (true) {
        while (less(*first, *pivot)) ++first;
        --last;
        while (less(*pivot, *last)) --last;
        if (!(first < last)) return first;
        swap_(*first, *last);
        ++first;
    }
}

template <class T, class Less>
GD_HD T* unguarded_partition_pivot(T* first, T* last, Less less)
{
    T* mid = first + (last - first) / 2;
    move_median_to_first(first, first + 1, mid, last - 1, less);
    return unguarded_partition(first + 1, last, first, less);
}

template <class T, class Less>
GD_HD void unguarded_linear_insert(T* last, Less less)
{
    const T val = *last;
    T* next = last;
    --next;
    while (less(val, *next)) {
        *last = *next;
        last = next;
        --next;
    }
    *last = val;
}

template <class T, class Less>
GD_HD void insertion_sort(T* first, T* last, Less less)
{
    if (first == last) return;
    for (T* i = first + 1; i != last; ++i) {
        if (less(*i, *first)) {
            const T val = *i;
            for (T* p = i; p != first; --p) *p = *(p - 1);  // move_backward(first, i, i + 1)
            *first = val;
        } else
            unguarded_linear_insert(i, less);
    }
}

// ---- heap helpers (only reached when the depth limit 2 lg n runs out)
template <class T, class Less>
GD_HD void push_heap_(T* first, long hole, long top, T value, Less less)
{
    long parent = (hole - 1) / 2;
    while (hole > top && less(first[parent], value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}

template <class T, class Less>
GD_HD void adjust_heap(T* first, long hole, long len, T value, Less less)
{
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (less(first[child], first[child - 1])) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    push_heap_(first, hole, top, value, less);
}

template <class T, class Less>
GD_HD void pop_heap_(T* first, T* last, T* result, Less less)
{
    const T value = *result;
    *result = *first;
    adjust_heap(first, 0L, (long)(last - first), value, less);
}

template <class T, class Less>
GD_HD void make_heap_(T* first, T* last, Less less)
{
    if (last - first < 2) return;
    const long len = (long)(last - first);
    long parent = (len - 2) / 2;
    while (true) {
        const T value = first[parent];
        adjust_heap(first, parent, len, value, less);
        if (parent == 0) return;
        parent--;
    }
}

template <class T, class Less>
GD_HD void heap_select(T* first, T* middle, T* last, Less less)
{
    make_heap_(first, middle, less);
    for (T* i = middle; i < last; ++i)
        if (less(*i, *first)) pop_heap_(first, middle, i, less);
}

template <class T, class Less>
GD_HD void sort_heap_(T* first, T* last, Less less)
{
    while (last - first > 1) {
        --last;
        pop_heap_(first, last, last, less);
    }
}

// std::__introselect(first, nth, last, depth_limit, less): the loop of std::nth_element from a given state
template <class T, class Less>
GD_HD void introselect(T* first, T* nth, T* last, int depth_limit, Less less)
{
    while (last - first > 3) {
        if (depth_limit == 0) {
            heap_select(first, nth + 1, last, less);
            swap_(*first, *nth);
            return;
        }
        --depth_limit;
        T* cut = unguarded_partition_pivot(first, last, less);
        if (cut <= nth)
            first = cut;
        else
            last = cut;
    }
    insertion_sort(first, last, less);
}

// std::nth_element(first, nth, last, less)
template <class T, class Less>
GD_HD void nth_element(T* first, T* nth, T* last, Less less)
{
    if (first == last || nth == last) return;
    int depth_limit = lg_((long)(last - first)) * 2;
    while (last - first > 3) {
        if (depth_limit == 0) {
            heap_select(first, nth + 1, last, less);
            swap_(*first, *nth);
            return;
        }
        --depth_limit;
        T* cut = unguarded_partition_pivot(first, last, less);
        if (cut <= nth)
            first = cut;
        else
            last = cut;
    }
    insertion_sort(first, last, less);
}

// std::partition(first, last, pred) for bidirectional iterators; returns the partition point
template <class T, class Pred>
GD_HD T* partition(T* first, T* last, Pred pred)
{
    while (true) {
        while (true) {
            if (first == last) return first;
            if (pred(*first))
                ++first;
            else
                break;
        }
        --last;
        while (true) {
            if (first == last) return first;
            if (!pred(*last))
                --last;
            else
                break;
        }
        swap_(*first, *last);
        ++first;
    }
}

// std::sort(first, last, less): introsort loop with an explicit stack instead of the recursion on the right part
template <class T, class Less>
GD_HD void sort(T* first, T* last, Less less)
{
    if (first == last) return;
    constexpr int THRESHOLD = 16;
    struct Job {
        T* first;
        T* last;
        int depth;
    };
    Job stack[64];  // the recursion depth is bounded by the depth limit 2 lg n <= 2 * 31
    int sp = 0;
    stack[sp++] = {first, last, lg_((long)(last - first)) * 2};
    while (sp > 0) {
        Job j = stack[--sp];
        // __introsort_loop(j.first, j.last, j.depth): the recursive call on [cut, last) runs BEFORE the loop continues on
        // [first, cut) — the two ranges are disjoint, so deferring the left part on the stack gives the same result
        while (j.last - j.first > THRESHOLD) {
            if (j.depth == 0) {
                heap_select(j.first, j.last, j.last, less);  // __partial_sort(first, last, last)
                sort_heap_(j.first, j.last, less);
                break;
            }
            --j.depth;
            T* cut = unguarded_partition_pivot(j.first, j.last, less);
            stack[sp++] = {cut, j.last, j.depth};  // right part (processed later: disjoint from the left part)
            j.last = cut;
        }
    }
    // __final_insertion_sort
    if (last - first > THRESHOLD) {
        insertion_sort(first, first + THRESHOLD, less);
        for (T* i = first + THRESHOLD; i != last; ++i) unguarded_linear_insert(i, less);
    } else
        insertion_sort(first, last, less);
}

// The first k elements std::sort(first, last, less) would leave, computed without sorting the rest: partitions whose range
// starts at or beyond position k are skipped.  Exact because (1) the recursive calls of introsort work on disjoint ranges,
// so a skipped range cannot influence another one, and (2) the final insertion pass moves an element backwards only past
// strictly greater ones, and every element of an earlier partition block is <= every element of a later one: nothing from
// beyond the end b of the block that holds position k - 1 ever moves in front of b.  Returns b (the sorted prefix length,
// >= min(k, n)).  k >= 16.
template <class T, class Less>
GD_HD long sort_prefix(T* first, T* last, long k, Less less)
{
    if (first == last) return 0;
    constexpr int THRESHOLD = 16;
    struct Job {
        T* first;
        T* last;
        int depth;
    };
    Job stack[64];
    int sp = 0;
    T* done_end = first;  // end of the last finished block that starts before position k
    stack[sp++] = {first, last, lg_((long)(last - first)) * 2};
    while (sp > 0) {
        Job j = stack[--sp];
        while (j.last - j.first > THRESHOLD) {
            if (j.depth == 0) {
                heap_select(j.first, j.last, j.last, less);
                sort_heap_(j.first, j.last, less);
                break;
            }
            --j.depth;
            T* cut = unguarded_partition_pivot(j.first, j.last, less);
            if (cut - first < k) stack[sp++] = {cut, j.last, j.depth};  // the right part still reaches into the prefix
            j.last = cut;
        }
        if (j.last > done_end) done_end = j.last;  // j.first < first + k for every job that is run
    }
    T* stop = done_end;
    if (last - first > THRESHOLD) {
        insertion_sort(first, first + THRESHOLD, less);
        if (stop < first + THRESHOLD) stop = first + THRESHOLD;
        for (T* i = first + THRESHOLD; i != stop; ++i) unguarded_linear_insert(i, less);
    } else {
        insertion_sort(first, last, less);
        stop = last;
    }
    return (long)(stop - first);
}

// ---- the Hoare partition as a data-parallel rule (what the CTA-wide selection kernel implements) ------------------------
// std::__unguarded_partition and std::partition both run two pointers towards each other: the left one stops at elements
// for which is_left_stop holds, the right one at is_right_stop, the two stopped elements are swapped, and the scan goes on
// until the pointers meet.  The pointers only ever look at elements they have not visited, so with l_k = the k-th left
// stop from the left and r_k = the k-th right stop from the right (in the ORIGINAL array) the result is: swap (l_k, r_k)
// for every k with l_k < r_k, nothing else.  A left stop x of rank k is swapped iff more than k right stops lie to its
// right; symmetrically for a right stop.  K = number of swapped pairs.  The position the unguarded variant returns is
// l_K when that left stop lies before r_{K-1} (or K = 0), else r_{K-1}.  This sequential statement of the rule is compared
// with the pointer loops above in tests/native/stdalgo_check.cpp; the kernel evaluates the same rule with prefix sums.
template <class T, class IsL, class IsR>
GD_HD long rank_pair_swap(T* v, long lo, long hi, IsL is_left_stop, IsR is_right_stop)
{
    long nR = 0;
    for (long i = lo; i < hi; ++i) nR += is_right_stop(v[i]) ? 1 : 0;
    // first pass: ranks on the original values; find K, l_K and r_{K-1}
    long K = 0, lK = -1, rK1 = -1, l0 = -1;
    {
        long L = 0, R_right = nR;  // left stops strictly left of i / right stops strictly right of i
        for (long i = lo; i < hi; ++i) {
            const bool isr = is_right_stop(v[i]), isl = is_left_stop(v[i]);
            if (isr) R_right -= 1;
            if (isl) {
                if (l0 < 0) l0 = i;
                if (R_right > L) K += 1;                    // swapped left stop (rank L)
                else if (lK < 0) lK = i;                    // first unswapped left stop = l_K
            }
            if (isr && L > R_right && rK1 < 0) rK1 = i;     // leftmost swapped right stop = r_{K-1}
            if (isl) L += 1;
        }
    }
    // second pass: exchange the values of the swapped pairs (pair k: left rank k <-> right rank k)
    {
        long li = lo, ri = hi - 1;
        for (long k = 0; k < K; ++k) {
            while (!is_left_stop(v[li])) ++li;
            while (!is_right_stop(v[ri])) --ri;
            swap_(v[li], v[ri]);
            ++li;
            --ri;
        }
    }
    if (K == 0) return l0;
    return (lK >= 0 && lK < rK1) ? lK : rK1;
}

// cv::KeyPointsFilter::retainBest(keypoints, n_points) on an array of n elements ordered by `greater` on the response:
// returns the new size; the surviving prefix is in the library's order.  resp(e) reads the response of an element.
template <class T, class Resp>
GD_HD int retain_best(T* v, int n, int n_points, Resp resp)
{
    if (n_points < 0 || n <= n_points) return n;
    if (n_points == 0) return 0;
    nth_element(v, v + n_points - 1, v + n, [resp](const T& a, const T& b) { return resp(a) > resp(b); });
    const auto amb = resp(v[n_points - 1]);
    T* e = partition(v + n_points, v + n, [resp, amb](const T& a) { return resp(a) >= amb; });
    return (int)(e - v);
}

}  // namespace stdalgo
}  // namespace gd
