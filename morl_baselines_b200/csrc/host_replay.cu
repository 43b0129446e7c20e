// host_replay.cu  -- (host code only; compiled by nvcc with the rest of the library)
// host-side (CPU) halves of the replay path, behind the same C-ABI.
//
// The prioritised replay buffer keeps the reference's semantics exactly (prioritized_buffer.py:12-82: a float64 sum tree whose
// inner nodes are updated INCREMENTALLY, node += new - old, in array order), so that the sampled indices are the reference's for
// the same numpy RNG stream.  That order-dependent float64 arithmetic is inherently sequential, so it stays on the host -- but in
// C instead of ~17 numpy calls per operation (100 + 120 us per minibatch at 65,536 leaves -> ~10 us), because it sits on the
// critical path between two GPU steps: priorities(t) -> tree -> indices(t+1).  The row gathers pack a host minibatch into the
// pinned staging buffer the single H2D copy of a step reads (reference buffer.py:84-94 builds six pageable temporaries).
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace {
// flat layout: level l (2^l nodes) starts at element 2^l - 1; level 0 is the root, level n_levels - 1 the leaves
inline double* level_ptr(double* tree, int l) { return tree + ((size_t)1 << l) - 1; }
inline const double* level_ptr(const double* tree, int l) { return tree + ((size_t)1 << l) - 1; }
}  // namespace

extern "C" int morl_host_sumtree_walk(const double* tree, int n_levels, const double* queries, int n, long long* out_index) {
    MORL_REQUIRE(tree && queries && out_index, MORL_ERR_NULL, "morl_host_sumtree_walk: NULL pointer argument");
    MORL_REQUIRE(n_levels >= 1 && n_levels <= 40 && n >= 0, MORL_ERR_SHAPE, "morl_host_sumtree_walk: bad shape n_levels=%d n=%d", n_levels, n);
    // SumTree.sample, prioritized_buffer.py:35-49, level by level over blocks of 16 independent queries (branch-free: the
    // comparisons are coin flips, and the loads of a block overlap)
    constexpr int kBlk = 16;
    for (int i0 = 0; i0 < n; i0 += kBlk) {
        const int m = n - i0 < kBlk ? n - i0 : kBlk;
        double q[kBlk];
        long long node[kBlk];
        for (int i = 0; i < m; ++i) {
            q[i] = queries[i0 + i];
            node[i] = 0;
        }
        for (int l = 1; l < n_levels; ++l) {
            const double* lv = level_ptr(tree, l);
            for (int i = 0; i < m; ++i) {
                const double left = lv[2 * node[i]];
                const bool right = q[i] > left;  // np.greater(query, left); query -= left * greater
                node[i] = 2 * node[i] + (long long)right;
                q[i] -= left * (double)right;  // arithmetic select: no branch for the compiler to mispredict
            }
        }
        for (int i = 0; i < m; ++i) out_index[i0 + i] = node[i];
    }
    return MORL_OK;
}

extern "C" int morl_host_sumtree_batch_set(double* tree, int n_levels, const long long* index, const double* priority, int n) {
    MORL_REQUIRE(tree && index && priority, MORL_ERR_NULL, "morl_host_sumtree_batch_set: NULL pointer argument");
    MORL_REQUIRE(n_levels >= 1 && n_levels <= 40 && n >= 0 && n < (1 << 24), MORL_ERR_SHAPE, "morl_host_sumtree_batch_set: bad shape n_levels=%d n=%d",
                 n_levels, n);
    const long long n_leaves = (long long)1 << (n_levels - 1);
    // np.unique(node_index, return_index=True): sorted unique leaves, each with the priority of its FIRST occurrence
    std::vector<uint64_t> key((size_t)n);
    for (int i = 0; i < n; ++i) {
        MORL_REQUIRE(index[i] >= 0 && index[i] < n_leaves, MORL_ERR_SHAPE, "morl_host_sumtree_batch_set: leaf index %lld out of range", index[i]);
        key[i] = ((uint64_t)index[i] << 24) | (uint64_t)i;  // n < 2^24
    }
    std::sort(key.begin(), key.end());
    std::vector<long long> node;
    std::vector<double> diff;
    node.reserve(n);
    diff.reserve(n);
    double* leaves = level_ptr(tree, n_levels - 1);
    for (int i = 0; i < n; ++i) {
        const long long idx = (long long)(key[i] >> 24);
        if (i > 0 && idx == (long long)(key[i - 1] >> 24)) continue;
        node.push_back(idx);
        diff.push_back(priority[key[i] & 0xFFFFFFu] - leaves[idx]);  // diff = new - old (prioritized_buffer.py:75-77)
    }
    const size_t m = node.size();
    for (int l = n_levels - 1; l >= 0; --l) {  // np.add.at(nodes, node_index, diff) level by level, in array order; node_index //= 2
        double* lv = level_ptr(tree, l);
        for (size_t i = 0; i < m; ++i) {
            lv[node[i]] += diff[i];
            node[i] >>= 1;
        }
    }
    return MORL_OK;
}

extern "C" int morl_host_gather_rows(const void* src, long long row_bytes, const long long* index, int n, void* dst) {
    MORL_REQUIRE(src && index && dst, MORL_ERR_NULL, "morl_host_gather_rows: NULL pointer argument");
    MORL_REQUIRE(row_bytes > 0 && n >= 0, MORL_ERR_SHAPE, "morl_host_gather_rows: bad shape row_bytes=%lld n=%d", row_bytes, n);
    const char* s = static_cast<const char*>(src);
    char* d = static_cast<char*>(dst);
    for (int i = 0; i < n; ++i) memcpy(d + (size_t)i * row_bytes, s + (size_t)index[i] * row_bytes, (size_t)row_bytes);
    return MORL_OK;
}

extern "C" int morl_host_gather_u8_to_i32(const unsigned char* src, long long row_elems, const long long* index, int n, int* dst) {
    MORL_REQUIRE(src && index && dst, MORL_ERR_NULL, "morl_host_gather_u8_to_i32: NULL pointer argument");
    MORL_REQUIRE(row_elems > 0 && n >= 0, MORL_ERR_SHAPE, "morl_host_gather_u8_to_i32: bad shape row_elems=%lld n=%d", row_elems, n);
    for (int i = 0; i < n; ++i)
        for (long long e = 0; e < row_elems; ++e) dst[(size_t)i * row_elems + e] = (int)src[(size_t)index[i] * row_elems + e];
    return MORL_OK;
}
