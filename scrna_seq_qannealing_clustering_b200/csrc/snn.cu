// snn.cu -- shared-nearest-neighbour graph construction on the device: the step BEFORE the hot path.
//
// Reference recipe (Seurat FindNeighbors as used by R/pbmc3k/Pbmc3k_general_data_preparation.Rmd:47-75 and
// R/benchmarks/Benchmark.Rmd:150-166; validated against the shipped R/benchmarks/graph_*.gexf fixtures):
//   exact kNN INCLUDING self (k.param = k)  ->  s_ij = |kNN(i) & kNN(j)|  ->  w_ij = s / (2k - s)  ->  drop the diagonal
//   ->  prune w < prune.SNN  ->  symmetric degree trim to `ord`: for i = 0..n-1, sequentially and in place, keep the `ord`
//   heaviest entries of column i (ties to the lower index) and zero the rest in column i AND row i.
// The host implementation snn.py is the specification; the kernels here reproduce its edge set and weights exactly
// (integer shared-neighbour counts; w is the same fp64 division).
//
// Batched: `num_problems` independent point sets (config 4: 512 subsets of 1000 cells, each with its own graph) are built in
// the same launches; a point only sees candidates of its own problem.
//
//   k_snn_knn      one warp per query: lanes stride over the problem's points (XT[d][j] coalesced), each lane keeps a private
//                  sorted top-k, the 32 lists are merged by k warp-argmin rounds.  Direct sum of squared differences in fp64.
//   k_snn_indeg /  reverse lists R(p) = { i : p in kNN(i) }
//   k_snn_revfill
//   k_snn_rows     one warp per row i: candidates = multiset union of R(p) over p in kNN(i), bitonic-sorted in shared memory,
//                  run lengths = s_ij; pass 0 counts the row's entries, pass 1 writes (j, w) ascending in j
//   k_snn_trim     the sequential trim as a dependency-driven persistent kernel: warps take rows in index order; row i waits
//                  (acquire) until every earlier neighbour that may still delete edges has published `done`, ranks its live
//                  entries by (-w, j), deletes the tail in both rows, then publishes its own flag (release).  Rows that never
//                  exceed `ord` are born done.  Forward progress: a row only waits on lower rows, which were handed out
//                  earlier to resident warps.
//   k_snn_count /  live entries with j > i -> edge list sorted by (u, v), local indices
//   k_snn_emit

#include "common.cuh"

#include <cub/device/device_scan.cuh>

using namespace qa;

namespace {

struct SnnDims {
    int32_t num_problems;
    int32_t dim;
    int32_t k;          // neighbours including self (clamped per problem to its size)
    int32_t max_degree; // <= 0: no trim
    double prune;
    int64_t total;      // points over all problems
};

__device__ __forceinline__ int snn_problem_of(const int64_t *off, int P, int64_t i) {
    int lo = 0, hi = P;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (off[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

constexpr int SNN_KMAX = 64;

// XT[d][total]: coordinates transposed (coalesced candidate loads).  nn[i][k]: global point indices, ascending distance, self first.
__global__ void __launch_bounds__(256) k_snn_knn(SnnDims S, const int64_t *off, const double *XT, int32_t *nn) {
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= S.total) return;
    const int p = snn_problem_of(off, S.num_problems, i);
    const int64_t lo = off[p], hi = off[p + 1];
    const int k = (int)min((int64_t)S.k, hi - lo);
    double bd[SNN_KMAX];
    int bj[SNN_KMAX];
    int cnt = 0;
    for (int64_t j = lo + lane; j < hi; j += 32) {
        double d = 0.0;
        for (int q = 0; q < S.dim; ++q) {
            const double t = __ldg(XT + (size_t)q * S.total + i) - __ldg(XT + (size_t)q * S.total + j);
            d += t * t;
        }
        if (j == i) d = -1.0;   // self first
        if (cnt == k && !(d < bd[k - 1] || (d == bd[k - 1] && (int)j < bj[k - 1]))) continue;
        int pos = cnt < k ? cnt : k - 1;
        while (pos > 0 && (d < bd[pos - 1] || (d == bd[pos - 1] && (int)j < bj[pos - 1]))) {
            bd[pos] = bd[pos - 1];
            bj[pos] = bj[pos - 1];
            --pos;
        }
        bd[pos] = d;
        bj[pos] = (int)j;
        if (cnt < k) ++cnt;
    }
    // merge: k rounds of warp-argmin over the heads of the 32 sorted lists
    int head = 0;
    for (int r = 0; r < k; ++r) {
        double d = head < cnt ? bd[head] : INFINITY;
        int j = head < cnt ? bj[head] : 0x7fffffff;
        double md = d;
        int mj = j;
        for (int o = 16; o > 0; o >>= 1) {
            const double od = __shfl_xor_sync(FULL_MASK, md, o);
            const int oj = __shfl_xor_sync(FULL_MASK, mj, o);
            if (od < md || (od == md && oj < mj)) { md = od; mj = oj; }
        }
        if (j == mj && d == md) ++head;
        if (lane == 0) nn[i * S.k + r] = mj;
    }
    if (lane == 0)
        for (int r = k; r < S.k; ++r) nn[i * S.k + r] = -1;
}

__global__ void k_snn_indeg(int64_t entries, const int32_t *nn, int32_t *indeg) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= entries) return;
    const int p = nn[e];
    if (p >= 0) atomicAdd(indeg + p, 1);
}

__global__ void k_snn_revfill(int64_t entries, int k, const int32_t *nn, const int64_t *rev_ptr, int32_t *fill, int32_t *rev) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= entries) return;
    const int p = nn[e];
    if (p >= 0) rev[rev_ptr[p] + atomicAdd(fill + p, 1)] = (int32_t)(e / k);
}

// candidate volume of a row: sum of |R(p)| over its neighbours (sizes the shared-memory sort)
__global__ void k_snn_volume(SnnDims S, const int32_t *nn, const int64_t *rev_ptr, int32_t *vol, int32_t *maxvol) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S.total) return;
    int v = 0;
    for (int r = 0; r < S.k; ++r) {
        const int p = nn[i * S.k + r];
        if (p >= 0) v += (int)(rev_ptr[p + 1] - rev_ptr[p]);
    }
    vol[i] = v;
    atomicMax(maxvol, v);
}

// one warp per row.  pass 0: deg[i] = number of j != i with w >= prune.  pass 1: adj_j / adj_w at row_ptr[i], ascending j.
__global__ void k_snn_rows(SnnDims S, const int64_t *off, const int32_t *nn, const int64_t *rev_ptr, const int32_t *rev, int cap,
                           int pass, int32_t *deg, const int64_t *row_ptr, int32_t *adj_j, double *adj_w) {
    extern __shared__ int32_t snn_sm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
    if (i >= S.total) return;
    int32_t *buf = snn_sm + (size_t)wib * cap;
    const int p = snn_problem_of(off, S.num_problems, i);
    const int kk = (int)min((int64_t)S.k, off[p + 1] - off[p]);
    // gather the candidate multiset
    int count = 0;
    for (int r = 0; r < kk; ++r) {
        const int q = nn[i * S.k + r];
        const int64_t b = rev_ptr[q], e = rev_ptr[q + 1];
        for (int64_t t = b + lane; t < e; t += 32) buf[count + (int)(t - b)] = rev[t];
        count += (int)(e - b);
    }
    int N = 32;
    while (N < count) N <<= 1;
    for (int t = count + lane; t < N; t += 32) buf[t] = 0x7fffffff;
    __syncwarp();
    // bitonic sort, ascending
    for (int size = 2; size <= N; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = lane; t < (N >> 1); t += 32) {
                const int a = 2 * t - (t & (stride - 1));
                const int b = a + stride;
                const bool up = (a & size) == 0;
                const int x = buf[a], y = buf[b];
                if ((x > y) == up) { buf[a] = y; buf[b] = x; }
            }
            __syncwarp();
        }
    }
    // run lengths -> (j, s); a lane starts a run where buf[t] != buf[t-1]
    const double twok = 2.0 * kk;
    int out = 0;
    int64_t base = pass ? row_ptr[i] : 0;
    for (int t0 = 0; t0 < count; t0 += 32) {
        const int t = t0 + lane;
        bool keep = false;
        int j = 0;
        double w = 0.0;
        if (t < count) {
            j = buf[t];
            if ((t == 0 || buf[t - 1] != j) && j != (int)i) {
                int s = 1;
                while (t + s < count && buf[t + s] == j) ++s;
                w = (double)s / (twok - (double)s);
                keep = w >= S.prune;
            }
        }
        const unsigned m = __ballot_sync(FULL_MASK, keep);
        if (pass && keep) {
            const int pos = out + __popc(m & ((1u << lane) - 1u));
            adj_j[base + pos] = j;
            adj_w[base + pos] = w;
        }
        out += __popc(m);
    }
    if (!pass && lane == 0) deg[i] = out;
}

__device__ __forceinline__ int snn_ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void snn_st_release(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void k_snn_trim_init(int64_t total, int max_degree, const int64_t *row_ptr, int *done, unsigned char *alive) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    done[i] = (row_ptr[i + 1] - row_ptr[i]) <= max_degree ? 1 : 0;
    for (int64_t e = row_ptr[i]; e < row_ptr[i + 1]; ++e) alive[e] = 1;
}

// position of neighbour `j` in row r (rows are ascending in j)
__device__ __forceinline__ int64_t snn_find(const int64_t *row_ptr, const int32_t *adj_j, int64_t r, int j) {
    int64_t lo = row_ptr[r], hi = row_ptr[r + 1];
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (adj_j[mid] < j) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(128) k_snn_trim(int64_t total, int max_degree, const int64_t *row_ptr, const int32_t *adj_j,
                                                  const double *adj_w, int *done, unsigned char *alive, unsigned long long *next) {
    const int lane = threadIdx.x & 31;
    for (;;) {
        unsigned long long it = 0;
        if (lane == 0) it = atomicAdd(next, 1ull);
        it = __shfl_sync(FULL_MASK, it, 0);
        if ((int64_t)it >= total) return;
        const int64_t i = (int64_t)it;
        const int64_t b = row_ptr[i], e = row_ptr[i + 1];
        if (e - b <= max_degree) continue;   // never exceeds the limit: born done
        // wait for the earlier neighbours that may still delete (i, j)
        for (int64_t t = b + lane; t < e; t += 32) {
            const int j = adj_j[t];
            if (j < (int)i)
                while (snn_ld_acquire(done + j) == 0) __nanosleep(64);
        }
        __syncwarp();
        __threadfence();
        // live entries, ranked by (-w, j): rank = number of live entries that come first.  The flags of row i were last written
        // by rows that have published `done` (acquired above) and stay untouched until this row publishes its own; they are
        // read from L2 (another row's flags may share the line in a stale L1 copy).  2 = live, marked for deletion.
        int live = 0;
        for (int64_t t = b + lane; t < e; t += 32) live += __ldcg(alive + t) ? 1 : 0;
        for (int o = 16; o > 0; o >>= 1) live += __shfl_xor_sync(FULL_MASK, live, o);
        if (live > max_degree) {
            for (int64_t t = b + lane; t < e; t += 32) {
                if (!__ldcg(alive + t)) continue;
                const double w = __ldg(adj_w + t);
                const int j = __ldg(adj_j + t);
                int rank = 0;
                for (int64_t u = b; u < e; ++u) {
                    if (u == t || !__ldcg(alive + u)) continue;
                    const double wu = __ldg(adj_w + u);
                    if (wu > w || (wu == w && __ldg(adj_j + u) < j)) ++rank;
                }
                if (rank >= max_degree) __stcg(alive + t, (unsigned char)2);
            }
            __syncwarp();
            for (int64_t t = b + lane; t < e; t += 32) {
                if (__ldcg(alive + t) == 2) {   // delete in column i and row i: the entry and its mirror in row j
                    __stcg(alive + t, (unsigned char)0);
                    __stcg(alive + snn_find(row_ptr, adj_j, __ldg(adj_j + t), (int)i), (unsigned char)0);
                }
            }
        }
        __syncwarp();
        __threadfence();
        if (lane == 0) snn_st_release(done + i, 1);
    }
}

__global__ void k_snn_count(int64_t total, const int64_t *row_ptr, const int32_t *adj_j, const unsigned char *alive, int32_t *cnt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int c = 0;
    for (int64_t t = row_ptr[i]; t < row_ptr[i + 1]; ++t) c += (alive[t] && adj_j[t] > (int)i) ? 1 : 0;
    cnt[i] = c;
}

__global__ void k_snn_emit(SnnDims S, const int64_t *off, const int64_t *row_ptr, const int32_t *adj_j, const double *adj_w,
                           const unsigned char *alive, const int64_t *edge_ptr, int32_t *eu, int32_t *ev, double *w) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S.total) return;
    const int64_t base = off[snn_problem_of(off, S.num_problems, i)];
    int64_t o = edge_ptr[i];
    for (int64_t t = row_ptr[i]; t < row_ptr[i + 1]; ++t) {
        if (alive[t] && adj_j[t] > (int)i) {
            eu[o] = (int32_t)(i - base);
            ev[o] = (int32_t)(adj_j[t] - base);
            w[o] = adj_w[t];
            ++o;
        }
    }
}

__global__ void k_snn_transpose(int64_t total, int dim, const double *X, double *XT) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total * dim) return;
    const int64_t i = t / dim;
    const int d = (int)(t % dim);
    XT[(size_t)d * total + i] = X[t];
}

// exclusive prefix sums into 64-bit offsets: out[0..count] (out[count] = total)
__global__ void k_snn_widen(int64_t count, const int32_t *in, int64_t *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= count) out[i] = i < count ? (int64_t)in[i] : 0;
}

int snn_scan(qa_ctx *ctx, int64_t count, const int32_t *in, int64_t *out, int64_t *total_host) {
    const int tpb = 256;
    k_snn_widen<<<(unsigned)((count + 1 + tpb - 1) / tpb), tpb, 0, ctx->stream>>>(count, in, out);
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, out, out, (int)(count + 1), ctx->stream);
    int rc = ensure(ctx->cubtmp, tmp);
    if (rc) return rc;
    cudaError_t ce = cub::DeviceScan::ExclusiveSum(ctx->cubtmp.p, tmp, out, out, (int)(count + 1), ctx->stream);
    if (ce != cudaSuccess) return fail(QA_ERR_CUDA, std::string("scan: ") + cudaGetErrorString(ce));
    ctx->launches += 2;
    if (total_host) {
        QA_CUDA(cudaMemcpyAsync(total_host, out + count, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        QA_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return QA_OK;
}

}  // namespace

extern "C" {

int qa_graph_destroy(qa_graph *g) {
    if (!g) return QA_OK;
    cudaSetDevice(g->ctx->device);
    void *ptrs[] = {g->eu, g->ev, g->w, g->node_ids};
    for (void *p : ptrs) if (p) cudaFree(p);
    delete g;
    return QA_OK;
}

int qa_snn_build(qa_ctx *ctx, int32_t num_problems, const int64_t *point_offsets, int32_t dim, const double *X, int32_t k,
                 double prune, int32_t max_degree, qa_graph **out) {
    if (!ctx || !out || !point_offsets) return fail(QA_ERR_ARG, "null argument");
    *out = nullptr;
    if (num_problems < 1 || dim < 1) return fail(QA_ERR_ARG, "need at least one point set and one dimension");
    if (k < 1 || k > SNN_KMAX) return fail(QA_ERR_LIMIT, "k must be in [1, 64]");
    const int64_t total = point_offsets[num_problems];
    if (point_offsets[0] != 0 || total < 0 || total >= (int64_t)0x7fffffff) return fail(QA_ERR_ARG, "bad point offsets");
    for (int p = 0; p < num_problems; ++p)
        if (point_offsets[p + 1] < point_offsets[p]) return fail(QA_ERR_ARG, "point offsets must be non-decreasing");
    if (total > 0 && !X) return fail(QA_ERR_ARG, "null coordinates");
    QA_CUDA(cudaSetDevice(ctx->device));
    qa_graph *G = new qa_graph();
    G->ctx = ctx;
    G->num_problems = num_problems;
    G->point_off.assign(point_offsets, point_offsets + num_problems + 1);
    G->edge_off.assign(num_problems + 1, 0);
    if (total == 0) { *out = G; return QA_OK; }

    SnnDims S;
    S.num_problems = num_problems;
    S.dim = dim;
    S.k = k;
    S.max_degree = max_degree;
    S.prune = prune;
    S.total = total;
    const int tpb = 256;
    auto blocks = [&](int64_t c) { return (unsigned)std::max<int64_t>(1, (c + tpb - 1) / tpb); };
    // scratch (freed on every path through the guard)
    SnnScratch sc;
    int rc = QA_OK;
    auto bail = [&](int code) { qa_graph_destroy(G); return code; };
#define SNN_CUDA(call)                                                                                        \
    do {                                                                                                      \
        cudaError_t e__ = (call);                                                                             \
        if (e__ != cudaSuccess) return bail(fail(QA_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__))); \
    } while (0)
    int64_t *d_off = nullptr, *rev_ptr = nullptr, *row_ptr = nullptr, *edge_ptr = nullptr;
    double *Xd = nullptr, *XT = nullptr, *adj_w = nullptr;
    int32_t *nn = nullptr, *indeg = nullptr, *fill = nullptr, *rev = nullptr, *vol = nullptr, *deg = nullptr, *adj_j = nullptr, *cnt = nullptr;
    int *done = nullptr;
    unsigned char *alive = nullptr;
    SNN_CUDA(sc.get(&d_off, (size_t)num_problems + 1));
    SNN_CUDA(cudaMemcpyAsync(d_off, point_offsets, ((size_t)num_problems + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    const double *Xsrc = X;
    if (!is_device_ptr(X)) {
        SNN_CUDA(sc.get(&Xd, (size_t)total * dim));
        SNN_CUDA(cudaMemcpyAsync(Xd, X, (size_t)total * dim * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        Xsrc = Xd;
    }
    SNN_CUDA(sc.get(&XT, (size_t)total * dim));
    k_snn_transpose<<<blocks(total * dim), tpb, 0, ctx->stream>>>(total, dim, Xsrc, XT);
    SNN_CUDA(sc.get(&nn, (size_t)total * k));
    k_snn_knn<<<blocks(total * 32), tpb, 0, ctx->stream>>>(S, d_off, XT, nn);
    ctx->launches += 2;
    // reverse lists
    SNN_CUDA(sc.get(&indeg, (size_t)total));
    SNN_CUDA(sc.get(&fill, (size_t)total));
    SNN_CUDA(cudaMemsetAsync(indeg, 0, (size_t)total * sizeof(int32_t), ctx->stream));
    SNN_CUDA(cudaMemsetAsync(fill, 0, (size_t)total * sizeof(int32_t), ctx->stream));
    k_snn_indeg<<<blocks(total * k), tpb, 0, ctx->stream>>>(total * k, nn, indeg);
    SNN_CUDA(sc.get(&rev_ptr, (size_t)total + 1));
    int64_t rev_total = 0;
    if ((rc = snn_scan(ctx, total, indeg, rev_ptr, &rev_total))) return bail(rc);
    SNN_CUDA(sc.get(&rev, (size_t)rev_total));
    k_snn_revfill<<<blocks(total * k), tpb, 0, ctx->stream>>>(total * k, k, nn, rev_ptr, fill, rev);
    // candidate volume -> shared-memory capacity of the row kernel
    SNN_CUDA(sc.get(&vol, (size_t)total));
    SNN_CUDA(cudaMemsetAsync(ctx->d_flag, 0, 2 * sizeof(int), ctx->stream));
    k_snn_volume<<<blocks(total), tpb, 0, ctx->stream>>>(S, nn, rev_ptr, vol, ctx->d_flag);
    ctx->launches += 3;
    int maxvol = 0;
    SNN_CUDA(cudaMemcpyAsync(&maxvol, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    SNN_CUDA(cudaStreamSynchronize(ctx->stream));
    int cap = 32;
    while (cap < maxvol) cap <<= 1;
    int warps = 4;
    while (warps > 1 && (size_t)warps * cap * sizeof(int32_t) > 96 * 1024) warps >>= 1;
    const size_t smem = (size_t)warps * cap * sizeof(int32_t);
    if (smem > 200 * 1024) return bail(fail(QA_ERR_LIMIT, "a point's shared-neighbour candidate list exceeds shared memory"));
    SNN_CUDA(cudaFuncSetAttribute(k_snn_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SNN_CUDA(sc.get(&deg, (size_t)total));
    SNN_CUDA(sc.get(&row_ptr, (size_t)total + 1));
    const unsigned row_blocks = (unsigned)((total + warps - 1) / warps);
    k_snn_rows<<<row_blocks, warps * 32, smem, ctx->stream>>>(S, d_off, nn, rev_ptr, rev, cap, 0, deg, nullptr, nullptr, nullptr);
    int64_t adj_total = 0;
    if ((rc = snn_scan(ctx, total, deg, row_ptr, &adj_total))) return bail(rc);
    SNN_CUDA(sc.get(&adj_j, (size_t)adj_total));
    SNN_CUDA(sc.get(&adj_w, (size_t)adj_total));
    SNN_CUDA(sc.get(&alive, (size_t)adj_total));
    k_snn_rows<<<row_blocks, warps * 32, smem, ctx->stream>>>(S, d_off, nn, rev_ptr, rev, cap, 1, deg, row_ptr, adj_j, adj_w);
    ctx->launches += 2;
    // symmetric degree trim
    SNN_CUDA(sc.get(&done, (size_t)total));
    const int md = max_degree > 0 ? max_degree : 0x7fffffff;
    k_snn_trim_init<<<blocks(total), tpb, 0, ctx->stream>>>(total, md, row_ptr, done, alive);
    ctx->launches++;
    if (max_degree > 0) {
        int bps = 0;
        SNN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_snn_trim, 128, 0));
        if (bps < 1) return bail(fail(QA_ERR_CUDA, "trim kernel does not fit on an SM"));
        SNN_CUDA(cudaMemsetAsync(ctx->d_stats, 0, sizeof(unsigned long long), ctx->stream));
        // persistent and fully resident: a waiting warp depends on rows held by other resident warps
        const int64_t grid = std::min<int64_t>((int64_t)bps * ctx->num_sms, (total + 3) / 4);
        k_snn_trim<<<(unsigned)grid, 128, 0, ctx->stream>>>(total, md, row_ptr, adj_j, adj_w, done, alive, ctx->d_stats);
        ctx->launches++;
    }
    // edge list
    SNN_CUDA(sc.get(&cnt, (size_t)total));
    SNN_CUDA(sc.get(&edge_ptr, (size_t)total + 1));
    k_snn_count<<<blocks(total), tpb, 0, ctx->stream>>>(total, row_ptr, adj_j, alive, cnt);
    int64_t m_total = 0;
    if ((rc = snn_scan(ctx, total, cnt, edge_ptr, &m_total))) return bail(rc);
    SNN_CUDA(cudaMalloc((void **)&G->eu, std::max<size_t>((size_t)m_total, 1) * sizeof(int32_t)));
    SNN_CUDA(cudaMalloc((void **)&G->ev, std::max<size_t>((size_t)m_total, 1) * sizeof(int32_t)));
    SNN_CUDA(cudaMalloc((void **)&G->w, std::max<size_t>((size_t)m_total, 1) * sizeof(double)));
    k_snn_emit<<<blocks(total), tpb, 0, ctx->stream>>>(S, d_off, row_ptr, adj_j, adj_w, alive, edge_ptr, G->eu, G->ev, G->w);
    ctx->launches += 2;
    // per-problem edge offsets: edge_ptr at the first point of every problem
    for (int p = 0; p <= num_problems; ++p)
        SNN_CUDA(cudaMemcpyAsync(&G->edge_off[p], edge_ptr + point_offsets[p], sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    SNN_CUDA(cudaStreamSynchronize(ctx->stream));
    SNN_CUDA(cudaGetLastError());
#undef SNN_CUDA
    *out = G;
    return QA_OK;
}

int64_t qa_graph_num_edges(const qa_graph *g, int32_t problem) {
    if (!g) return (int64_t)fail(QA_ERR_ARG, "null graph");
    if (problem < 0) return g->edge_off[g->num_problems];
    if (problem >= g->num_problems) return (int64_t)fail(QA_ERR_ARG, "problem index out of range");
    return g->edge_off[problem + 1] - g->edge_off[problem];
}

int qa_graph_num_nodes(const qa_graph *g, int32_t problem) {
    if (!g) return fail(QA_ERR_ARG, "null graph");
    if (problem < 0) return (int)g->point_off[g->num_problems];
    if (problem >= g->num_problems) return fail(QA_ERR_ARG, "problem index out of range");
    return (int)(g->point_off[problem + 1] - g->point_off[problem]);
}

int qa_graph_get_edges(const qa_graph *g, int32_t problem, int32_t *eu, int32_t *ev, double *w) {
    if (!g) return fail(QA_ERR_ARG, "null graph");
    if (problem >= g->num_problems) return fail(QA_ERR_ARG, "problem index out of range");
    const int64_t b = problem < 0 ? 0 : g->edge_off[problem];
    const int64_t e = problem < 0 ? g->edge_off[g->num_problems] : g->edge_off[problem + 1];
    QA_CUDA(cudaSetDevice(g->ctx->device));
    if (e > b) {
        if (eu) QA_CUDA(cudaMemcpy(eu, g->eu + b, (size_t)(e - b) * sizeof(int32_t), cudaMemcpyDefault));
        if (ev) QA_CUDA(cudaMemcpy(ev, g->ev + b, (size_t)(e - b) * sizeof(int32_t), cudaMemcpyDefault));
        if (w) QA_CUDA(cudaMemcpy(w, g->w + b, (size_t)(e - b) * sizeof(double), cudaMemcpyDefault));
    }
    return QA_OK;
}

int qa_graph_device_edges(const qa_graph *g, int32_t problem, const int32_t **eu, const int32_t **ev, const double **w) {
    if (!g || !eu || !ev || !w) return fail(QA_ERR_ARG, "null argument");
    if (problem < 0 || problem >= g->num_problems) return fail(QA_ERR_ARG, "problem index out of range");
    const int64_t b = g->edge_off[problem];
    *eu = g->eu + b;
    *ev = g->ev + b;
    *w = g->w + b;
    return QA_OK;
}

}  // extern "C"
