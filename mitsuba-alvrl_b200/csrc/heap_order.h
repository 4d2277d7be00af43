/*
 * The multi-cluster queue of Clustering (Preprocessor.h: boost::heap::priority_queue<ClusterNode>, which is a
 * std::vector driven by std::push_heap / std::pop_heap) restated as plain array operations that compile for the host
 * and the device.  The ARRAY ORDER of the queue is part of the result -- getVrlsPerCluster / sampleRepresentatives
 * (Preprocessor.cpp:354-378, 526-543) walk it front to back -- so the sift sequences below reproduce libstdc++'s
 * __push_heap / __adjust_heap step by step, ties included (tests/test_abi_host_cpu.py checks them against the
 * standard library on random keys with duplicates).
 */
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define ALVRL_HEAP_HD __host__ __device__ __forceinline__
#else
#define ALVRL_HEAP_HD inline
#endif

namespace alvrl {

struct HeapEntry { float key; uint32_t id; };      /* key = undersamplingVar + integrationVar (ClusterNode::operator<, 289-298) */

/* std::push_heap(first, first + count + 1) with the new element `v` at the back; H: HeapEntry * or any indexable view */
template <typename H>
ALVRL_HEAP_HD void heap_push(H h, uint32_t &count, HeapEntry v) {
    uint32_t hole = count++;
    while (hole > 0) {
        const uint32_t parent = (hole - 1) >> 1;
        if (!(h[parent].key < v.key)) break;
        h[hole] = h[parent]; hole = parent;
    }
    h[hole] = v;
}

/* std::pop_heap(first, first + count); back(); pop_back() */
template <typename H>
ALVRL_HEAP_HD HeapEntry heap_pop(H h, uint32_t &count) {
    const HeapEntry top = h[0];
    if (count > 1) {
        const uint32_t len = count - 1;
        const HeapEntry value = h[len];
        uint32_t hole = 0, child = 0;
        while (child < (len - 1) / 2) {
            child = 2 * (child + 1);
            if (h[child].key < h[child - 1].key) child--;
            h[hole] = h[child]; hole = child;
        }
        if ((len & 1u) == 0 && child == (len - 2) / 2) {
            child = 2 * (child + 1);
            h[hole] = h[child - 1]; hole = child - 1;
        }
        while (hole > 0) {
            const uint32_t parent = (hole - 1) >> 1;
            if (!(h[parent].key < value.key)) break;
            h[hole] = h[parent]; hole = parent;
        }
        h[hole] = value;
    }
    count--;
    return top;
}

/* a queue whose first `cap` entries live in one array (shared memory) and the rest in another (global memory) */
struct SplitHeap {
    HeapEntry *lo, *hi; uint32_t cap;
    ALVRL_HEAP_HD HeapEntry &operator[](uint32_t i) const { return i < cap ? lo[i] : hi[i - cap]; }
};

} // namespace alvrl
