// adjacency.cuh -- k_rowptr: row pointers of a sorted (key, value) entry list; used by model.cu (adjacency of a model) and
// builders.cu (per-vertex edge lists of a graph).  One copy per translation unit.
#pragma once

namespace {

// rowptr[x] = first sorted position whose key >= x  (x in [0, rows]); rows beyond are clamped to `entries`
__global__ void k_rowptr(int64_t rows_alloc, int64_t entries, const uint32_t *sorted_keys, int32_t *rowptr, int *maxdeg) {
    const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= rows_alloc) return;
    int64_t lo = 0, hi = entries;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if ((int64_t)sorted_keys[mid] < x) lo = mid + 1; else hi = mid;
    }
    rowptr[x] = (int32_t)lo;
    // degree of row x-1 is rowptr[x]-rowptr[x-1]; computed by the thread of x via a second search
    if (x > 0) {
        int64_t lo2 = 0, hi2 = entries;
        while (lo2 < hi2) {
            const int64_t mid = (lo2 + hi2) >> 1;
            if ((int64_t)sorted_keys[mid] < x - 1) lo2 = mid + 1; else hi2 = mid;
        }
        atomicMax(maxdeg, (int)(lo - lo2));
    }
}

}  // namespace
