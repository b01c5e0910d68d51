// pk_stdsort.cuh -- libstdc++'s std::sort (introsort: median-of-3 quicksort to depth 2*floor(log2 n), heapsort
// fallback, final insertion sort with threshold 16; bits/stl_algo.h) on (key, index) pairs compared by key only,
// exactly as the reference sorts its reliabilities (src/KanekoKernelProcessor.cpp:148,343).  std::sort is not
// stable, so when two |alpha| are EQUAL the resulting order is an artefact of this very algorithm; replaying it
// step by step is the only way to stay bit-exact on inputs with ties (quantised or text-derived samples).
// Runs on ONE lane, only for frames where the parallel rank sort detected a tie and n > 16 (rare).
#pragma once
#include <stdint.h>

namespace pk_stdsort {

struct Arr {
    double *k;
    uint8_t *x;
    __device__ __forceinline__ bool lt(int a, int b) const { return k[a] < k[b]; }
    __device__ __forceinline__ void swap(int a, int b) const {
        const double tk = k[a]; k[a] = k[b]; k[b] = tk;
        const uint8_t tx = x[a]; x[a] = x[b]; x[b] = tx;
    }
};

__device__ inline void unguarded_linear_insert(const Arr &a, int last) {
    const double vk = a.k[last];
    const uint8_t vx = a.x[last];
    int next = last - 1;
    while (vk < a.k[next]) { a.k[last] = a.k[next]; a.x[last] = a.x[next]; last = next; --next; }
    a.k[last] = vk; a.x[last] = vx;
}
__device__ inline void insertion_sort(const Arr &a, int first, int last) {
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        if (a.k[i] < a.k[first]) {
            const double vk = a.k[i];
            const uint8_t vx = a.x[i];
            for (int j = i; j > first; --j) { a.k[j] = a.k[j - 1]; a.x[j] = a.x[j - 1]; }   // move_backward
            a.k[first] = vk; a.x[first] = vx;
        } else
            unguarded_linear_insert(a, i);
    }
}
// heap helpers of the depth-limit fallback (__partial_sort(first, last, last) = make_heap + sort_heap)
__device__ inline void push_heap(const Arr &a, int first, int hole, int top, double vk, uint8_t vx) {
    int parent = (hole - 1) / 2;
    while (hole > top && a.k[first + parent] < vk) {
        a.k[first + hole] = a.k[first + parent]; a.x[first + hole] = a.x[first + parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    a.k[first + hole] = vk; a.x[first + hole] = vx;
}
__device__ inline void adjust_heap(const Arr &a, int first, int hole, int len, double vk, uint8_t vx) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (a.k[first + child] < a.k[first + child - 1]) child--;
        a.k[first + hole] = a.k[first + child]; a.x[first + hole] = a.x[first + child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        a.k[first + hole] = a.k[first + child - 1]; a.x[first + hole] = a.x[first + child - 1];
        hole = child - 1;
    }
    push_heap(a, first, hole, top, vk, vx);
}
__device__ inline void heapsort(const Arr &a, int first, int last) {
    const int len = last - first;
    if (len < 2) return;
    for (int parent = (len - 2) / 2;; --parent) {
        adjust_heap(a, first, parent, len, a.k[first + parent], a.x[first + parent]);
        if (parent == 0) break;
    }
    while (last - first > 1) {
        --last;
        const double vk = a.k[last];
        const uint8_t vx = a.x[last];
        a.k[last] = a.k[first]; a.x[last] = a.x[first];
        adjust_heap(a, first, 0, last - first, vk, vx);
    }
}
__device__ inline void move_median_to_first(const Arr &a, int result, int p, int q, int r) {
    if (a.lt(p, q)) {
        if (a.lt(q, r)) a.swap(result, q);
        else if (a.lt(p, r)) a.swap(result, r);
        else a.swap(result, p);
    } else if (a.lt(p, r)) a.swap(result, p);
    else if (a.lt(q, r)) a.swap(result, r);
    else a.swap(result, q);
}
__device__ inline int unguarded_partition(const Arr &a, int first, int last, int pivot) {
    for (;;) {
        while (a.lt(first, pivot)) ++first;
        --last;
        while (a.lt(pivot, last)) --last;
        if (!(first < last)) return first;
        a.swap(first, last);
        ++first;
    }
}
// std::sort(k, k + n): the right-hand partitions the reference handles by recursion go on an explicit stack
// (disjoint ranges, so the processing order does not change the result).
__device__ inline void sort(double *keys, uint8_t *idx, int n) {
    if (n < 2) return;
    Arr a{keys, idx};
    int lg = 0;
    while ((1 << (lg + 1)) <= n) ++lg;
    int sf[40], sl[40], sd[40], sp = 0;
    sf[0] = 0; sl[0] = n; sd[0] = 2 * lg; sp = 1;
    while (sp) {
        --sp;
        int first = sf[sp], last = sl[sp], depth = sd[sp];
        while (last - first > 16) {
            if (depth == 0) { heapsort(a, first, last); break; }
            --depth;
            const int mid = first + (last - first) / 2;
            move_median_to_first(a, first, first + 1, mid, last - 1);
            const int cut = unguarded_partition(a, first + 1, last, first);
            sf[sp] = cut; sl[sp] = last; sd[sp] = depth; ++sp;   // __introsort_loop(cut, last, depth)
            last = cut;
        }
    }
    if (n > 16) {
        insertion_sort(a, 0, 16);
        for (int i = 16; i != n; ++i) unguarded_linear_insert(a, i);
    } else
        insertion_sort(a, 0, n);
}

}  // namespace pk_stdsort
