// qanneal.cu -- libqanneal.so: hand-written sm_100a simulated annealing for the clustering QUBOs of
// michal7kw/scRNA_seq_QAnnealing_Clustering (C ABI in include/qanneal.h).
//
// What it replaces (reference file:line -> upstream algorithm):
//   sampler.sample_qubo / .sample / .sample_dqm / .sample_cqm call sites
//     Python_Functions/BQM_clustering.py:57,75,85,245,263,273,386; QA_subsampling.py:42,56,65;
//     DQM_clustering.py:45; CQM_clustering.py:53,89
//   -> dwave-neal neal/src/cpu_sa.cpp: general_simulated_annealing / simulated_annealing_run /
//      get_flip_energy / get_state_energy / FASTRAND.
//
// Design (B200-first, not a translation of the CPU loop):
//   * one WARP per read.  The read's local fields f[v] = h_v + sum_j J_vj s_j live in HBM as fp64 and
//     are streamed 32 variables (256 B) at a time with ld.global.cg + prefetch.global.L2 run-ahead;
//     spins are bit-packed (1 bit/attempt of traffic).  neal's dE[v] is recovered exactly as
//     -2*s_v*f[v] (scaling by +-2 commutes with rounding), so a flip needs no read of s_j: every
//     neighbour update is one fire-and-forget  red.global.add.f64 f[j], -2*s_v*J  performed in L2.
//   * the 32 lanes of a chunk decide "candidate" (dE < 44.36142/beta) in parallel; candidates are then
//     resolved in variable order with a warp-uniform xorshift128+ stream, so the sequence of RNG
//     draws, accepts and fp64 roundings is exactly neal's.  In-chunk neighbours are patched in
//     registers by shuffles in neighbour order.
//   * persistent grid (multiple of the SM count), reads pulled from an atomic counter; only the
//     resident warps own fp64 scratch (n*8 B each), so 100k reads x 131k variables need ~5 GB, not 118.
//   * optional rank-1 "group" terms evaluated lazily from per-warp integer counters in shared memory.
//   * energies are evaluated by a thread-per-read kernel in neal's summation order on a
//     read-transposed packed-spin layout (coalesced), followed by a warp-shuffle argmin.
// Compile: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo (see __graft_entry__.build()).

#include "../../include/qanneal.h"

#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cmath>
#include <ctime>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define QA_VERSION 100
#define FULL_MASK 0xffffffffu

namespace {

thread_local std::string g_err;

int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}

#define QA_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(QA_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));         \
    } while (0)

// ------------------------------------------------------------------------------------------------
// device-side problem description (one per independent problem; a plain model is a batch of one)
// ------------------------------------------------------------------------------------------------
struct ProblemDesc {
    int32_t n;        // variables
    int32_t nch;      // 32-variable chunks = ceil(n/32)
    int32_t ngroups;  // rank-1 groups (0: none)
    int32_t reads;    // reads of this problem
    int32_t rpad;     // reads padded to a multiple of 32 (stride of packedT)
    int32_t pad_;
    int64_t m;        // couplers
    int64_t read_base;  // index of this problem's read 0 in the global read numbering
    const int32_t *rowptr;  // [nch*32 + 1] entries valid (padding rows are empty)
    const int32_t *col;     // CSR neighbour, local variable index
    const double *val;      // CSR coupling
    const double *h;        // [n]
    const int32_t *starts;  // COO in caller order (energy summation order)
    const int32_t *ends;
    const double *w;
    const int32_t *grp;     // [nch*32] or null
    const int32_t *coef;    // [nch*32]
    const double *lambda;   // [ngroups]
    const long long *kappa; // [ngroups]
    int8_t *states;         // [reads][n]
    uint32_t *packedT;      // [nch][rpad] final spins, bit i of word (c, r) = spin of variable 32c+i (1: +1)
    double *energies;       // [reads]
    // block word tables of the pull variant (null until built): for every block of 16 variables the distinct spin words
    // its CSR rows refer to, and per CSR entry the index of its word in that list (255 = the block's own word)
    const int32_t *bw_ptr;      // [nblk + 1], this problem's first block at index 0 (values are global offsets)
    const int32_t *bw_words;    // global array
    const unsigned short *ent_slot; // indexed like col: slot | (bit << 8)
    // coupling slabs of the replay kernel (null until built): one {RpHdr, RpEntry[]} per block of 16 variables
    const unsigned char *rp_slabs;  // global slab storage
    const uint32_t *rp_off;         // [rp_nslabs + 1] slab offsets of this problem in 16-byte units
    int32_t rp_nslabs;              // blocks (slabs) of this problem
    int32_t pad2_;
    const double *betas;            // this problem's own beta schedule [num_betas], or null: the launch's shared schedule
};

struct AnnealParams {
    const ProblemDesc *descs;
    int32_t num_problems;
    int32_t reads_per_problem;   // uniform (batch) ; == total reads for a single model
    int64_t total_reads;
    const double *betas;
    int32_t num_betas;
    int32_t sweeps_per_beta;
    const unsigned long long *seeds;
    int32_t seed_mode;
    double *f_scratch;           // [slots][f_stride]
    uint32_t *spw_scratch;       // [slots][spw_stride]
    int64_t f_stride;
    int64_t spw_stride;
    unsigned long long *counter; // next read to hand out
    unsigned long long *stats;   // QA_NSTAT counters
    int *error_flag;
    int64_t read_begin;          // wave support: global read range [read_begin, read_end)
    int64_t read_end;
    // lockstep kernels (one warp = 32 reads of one problem)
    double *fT_scratch;          // [slots][fT_stride]  read-interleaved local fields f[v][lane]
    int64_t fT_stride;           // n_pad_max * 32
    int32_t tiles_per_problem;
    int32_t max_groups;          // stride of the per-thread group counters in shared memory
    int64_t total_tiles;
    // replay kernel (one CTA = `warps` consecutive 32-read tiles of one problem)
    void *sf_scratch;            // [slots][sf_stride] {S | F << 16} half-word words, read-interleaved (uint32)
    int64_t sf_stride;           // nch_max * 2 * 32
    int64_t groups_per_problem;  // ceil(tiles_per_problem / warps per CTA)
    int64_t total_items;         // num_problems * groups_per_problem
    int32_t switch_permille;     // replay -> push hand-over: CTA-wide acceptance of a sweep below this many per mille
    int32_t rp_slab_init;        // adjacency lists ascending: the field set-up pass runs through the slab ring
    const int *interrupt_flag;   // host-mapped flag (or null): CTAs stop pulling work once it is non-zero
    // dense k-way kernel (dense.cuh)
    uint32_t *dn_spins;          // [slots][dn_stride] bit-packed spins, one slot per resident warp
    int64_t dn_stride;           // (ncp / 32) * 32 * K words
};

enum { ST_CAND = 0, ST_DRAWS, ST_ACC, ST_NBR, ST_ACTIVE, ST_CHUNKS, ST_TIES, QA_NSTAT };
constexpr int QA_NDEBUG = 8;   // development counters behind the read counter (QA_RP_PROFILE builds print them)

constexpr int QA_TPB_MAX = 256;
// measured on B200 (config 3): lockstep wins from ~3.5 tiles of 32 reads per SM, warp-per-read below that
constexpr int QA_PREFETCH_CHUNKS = 16;  // run-ahead distance of the L2 prefetch, in 256-byte chunks
constexpr double QA_TWO64 = 18446744073709551616.0;

__device__ __forceinline__ unsigned long long rng_next(unsigned long long &s0, unsigned long long &s1) {
    // neal FASTRAND (xorshift128+)
    unsigned long long x = s0;
    const unsigned long long y = s1;
    s0 = y;
    x ^= x << 23;
    s1 = x ^ y ^ (x >> 17) ^ (y >> 26);
    return s1 + y;
}

__device__ __forceinline__ void prefetch_l2(const void *p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

__device__ __forceinline__ void red_add_f64(double *addr, double v) {
    // fire-and-forget fp64 reduction performed at L2 (RED.E.ADD.F64); round-to-nearest like the CPU add
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
}

__device__ __forceinline__ void red_add_f64_if(double *addr, double v, bool pred) {
    // predicated form: no branch / reconvergence point in the neighbour loop
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p red.global.add.f64 [%0], %1;\n\t}" ::"l"(addr), "d"(v), "r"((int)pred)
                 : "memory");
}

struct WarpStats {
    unsigned long long cand, draws, acc, nbr, active, chunks, ties;
};

// ------------------------------------------------------------------------------------------------
// one read, reference order.  Restates neal simulated_annealing_run() for one `state`.
// ------------------------------------------------------------------------------------------------
template <bool GROUPS>
__device__ void anneal_read(const ProblemDesc &D, const AnnealParams &P, int64_t r_local, double *__restrict__ f,
                            uint32_t *__restrict__ spw, unsigned long long &s0, unsigned long long &s1,
                            long long *Mw, WarpStats &st, int *error_flag) {
    const int lane = threadIdx.x & 31;
    const int n = D.n;
    const int nch = D.nch;
    int8_t *state_row = D.states + r_local * (int64_t)n;

    // ---- pack the initial +-1 bytes into spin words (bit = 1 <=> s = +1); padding variables are +1
    for (int cb = 0; cb < nch; cb += 32) {
        uint32_t wg = 0xffffffffu;
        const int cend = min(32, nch - cb);
        for (int k = 0; k < cend; ++k) {
            const int v = (cb + k) * 32 + lane;
            int s = 1;
            if (v < n) {
                s = state_row[v];
                if (s != 1 && s != -1) atomicExch(error_flag, QA_ERR_STATE);
            }
            const uint32_t w = __ballot_sync(FULL_MASK, s > 0);
            if (lane == k) wg = w;
        }
        __stcg(spw + cb + lane, wg);
    }
    __syncwarp();

    // ---- local fields, neal get_flip_energy(): energy = h[v]; for nbr in adjacency order: energy += s_nbr*J
    for (int c = 0; c < nch; ++c) {
        const int v = c * 32 + lane;
        double fv = -INFINITY;  // padding: dE = -2*(+1)*(-inf) = +inf, never a candidate
        if (v < n) {
            fv = D.h[v];
            const int e0 = D.rowptr[v], e1 = D.rowptr[v + 1];
            for (int e = e0; e < e1; ++e) {
                const int j = D.col[e];
                const uint32_t wj = __ldcg(spw + (j >> 5));
                const double J = D.val[e];
                fv += ((wj >> (j & 31)) & 1u) ? J : -J;
            }
        }
        __stcg(f + v, fv);
    }
    if (GROUPS) {
        for (int g = lane; g < D.ngroups; g += 32) Mw[g] = 0;
        __syncwarp();
        for (int c = 0; c < nch; ++c) {
            const int v = c * 32 + lane;
            const int g = D.grp[v];
            if (g >= 0) {
                const uint32_t wv = __ldcg(spw + c);
                const long long a = D.coef[v];
                atomicAdd(reinterpret_cast<unsigned long long *>(Mw + g),
                          (unsigned long long)(((wv >> lane) & 1u) ? a : -a));
            }
        }
    }
    __syncwarp();

    // ---- the anneal: for beta: for sweep: for var (neal order)
    for (int b = 0; b < P.num_betas; ++b) {
        const double beta = (D.betas ? D.betas : P.betas)[b];
        const double thr = 44.36142 / beta;
        for (int sw = 0; sw < P.sweeps_per_beta; ++sw) {
            uint32_t wg = 0;
            bool gdirty = false;
            for (int c = 0; c < nch; ++c) {
                const int k = c & 31;
                if (k == 0) {
                    wg = __ldcg(spw + c + lane);
                    gdirty = false;
                }
                if ((lane & 15) == 0) {
                    int cp = c + QA_PREFETCH_CHUNKS;
                    if (cp >= nch) cp %= nch;
                    prefetch_l2(f + cp * 32 + lane);
                }
                const int v = c * 32 + lane;
                double fv = __ldcg(f + v);
                uint32_t w = __shfl_sync(FULL_MASK, wg, k);
                int g = -1;
                long long a = 0, kap = 0;
                double lam = 0.0;
                if (GROUPS) {
                    g = __ldg(D.grp + v);
                    if (g >= 0) {
                        a = __ldg(D.coef + v);
                        lam = __ldg(D.lambda + g);
                        kap = __ldg(D.kappa + g);
                    }
                }
                // neal's delta_energy[v] == -2*s_v*f[v] exactly; plus the lazily evaluated rank-1 cost
                auto flip_cost = [&](uint32_t word) -> double {
                    const bool up = (word >> lane) & 1u;
                    double d = up ? -2.0 * fv : 2.0 * fv;
                    if (GROUPS) {
                        if (g >= 0) {
                            const long long t = a * (a - (up ? 1 : -1) * (Mw[g] + kap));
                            d = d + lam * (double)t;
                        }
                    }
                    return d;
                };
                double dE = flip_cost(w);
                bool cand = !(dE >= thr);  // neal: if (delta_energy[var] >= threshold) continue;
                uint32_t pend = __ballot_sync(FULL_MASK, cand);
                if (pend) {
                    st.active++;
                    const int e0 = __ldg(D.rowptr + v), e1 = __ldg(D.rowptr + v + 1);
                    bool pvalid = false;
                    double p = 0.0;
                    while (pend) {
                        const int l = __ffs(pend) - 1;
                        pend &= pend - 1;
                        st.cand++;
                        const double dEl = __shfl_sync(FULL_MASK, dE, l);
                        bool acc = true;
                        if (dEl > 0.0) {
                            const uint32_t pv = __ballot_sync(FULL_MASK, pvalid);
                            if (!((pv >> l) & 1u)) {
                                // exp(-dE*beta)*2^64 for every lane that may need it (one SIMT pass)
                                if (!pvalid && cand && dE > 0.0) p = exp(-dE * beta) * QA_TWO64;
                                pvalid = true;
                            }
                            const unsigned long long rnd = rng_next(s0, s1);
                            st.draws++;
                            const double pl = __shfl_sync(FULL_MASK, p, l);
                            const double rd = __ull2double_rn(rnd);
                            acc = pl > rd;  // neal: exp(-dE*beta) * RANDMAX > rand
                            if (fabs(pl - rd) <= pl * 3.5527136788005009e-15) st.ties++;
                        }
                        if (acc) {
                            st.acc++;
                            const int r0 = __shfl_sync(FULL_MASK, e0, l);
                            const int r1 = __shfl_sync(FULL_MASK, e1, l);
                            const bool up_l = (w >> l) & 1u;
                            const double cf = up_l ? -2.0 : 2.0;  // f[j] += -2*s_l*J  <=> dE[j] += 4*s_l*J*s_j
                            st.nbr += (unsigned long long)(r1 - r0);
                            uint32_t touched = 0;
                            for (int base = r0; base < r1; base += 32) {
                                const int e = base + lane;
                                const bool valid = e < r1;
                                int j = -1;
                                double d = 0.0;
                                if (valid) {
                                    j = __ldg(D.col + e);
                                    d = cf * __ldg(D.val + e);
                                    red_add_f64(f + j, d);
                                }
                                uint32_t inm = __ballot_sync(FULL_MASK, valid && (j >> 5) == c);
                                while (inm) {  // patch register copies of this chunk, in neighbour order
                                    const int kk = __ffs(inm) - 1;
                                    inm &= inm - 1;
                                    const int tj = __shfl_sync(FULL_MASK, j, kk) & 31;
                                    const double dk = __shfl_sync(FULL_MASK, d, kk);
                                    if (lane == tj) fv += dk;
                                    touched |= 1u << tj;
                                }
                            }
                            w ^= 1u << l;
                            gdirty = true;
                            if (lane == k) wg = w;
                            if (GROUPS) {
                                const int gl = __shfl_sync(FULL_MASK, g, l);
                                if (gl >= 0) {
                                    const long long al = __shfl_sync(FULL_MASK, a, l);
                                    __syncwarp();
                                    if (lane == 0) Mw[gl] -= 2 * al * (up_l ? 1 : -1);
                                    __syncwarp();
                                    touched |= __ballot_sync(FULL_MASK, g == gl);
                                }
                            }
                            if ((touched >> lane) & 1u) pvalid = false;
                            dE = flip_cost(w);
                            cand = !(dE >= thr);
                            pend = __ballot_sync(FULL_MASK, cand) & ~((2u << l) - 1u);
                        }
                    }
                    __syncwarp();  // order this chunk's reductions before later loads of the same addresses
                }
                if ((k == 31 || c == nch - 1) && gdirty) __stcg(spw + (c & ~31) + lane, wg);
            }
            st.chunks += (unsigned long long)nch;
        }
    }
    __syncwarp();

    // ---- results: +-1 bytes back into the caller's row, packed transposed copy for the energy kernel
    for (int c = 0; c < nch; ++c) {
        const int v = c * 32 + lane;
        const uint32_t w = __ldcg(spw + c);
        if (v < n) state_row[v] = ((w >> lane) & 1u) ? 1 : -1;
    }
    for (int c = lane; c < nch; c += 32) D.packedT[(int64_t)c * D.rpad + r_local] = __ldcg(spw + c);
}

template <bool GROUPS>
__global__ void __launch_bounds__(QA_TPB_MAX, 4) k_anneal_ref(AnnealParams P) {
    __shared__ long long Msh[GROUPS ? (QA_TPB_MAX / 32) * QA_MAX_GROUPS : 1];
    __shared__ unsigned long long next_read[QA_TPB_MAX / 32];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int64_t slot = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
    double *f = P.f_scratch + slot * P.f_stride;
    uint32_t *spw = P.spw_scratch + slot * P.spw_stride;
    long long *Mw = GROUPS ? (Msh + wib * QA_MAX_GROUPS) : Msh;
    WarpStats st = {0, 0, 0, 0, 0, 0, 0};

    if (P.seed_mode == QA_SEED_STREAM) {
        // one xorshift128+ stream across reads (neal num_reads=R): inherently serial, validation only
        if (slot != 0) return;
        unsigned long long s0 = P.seeds[0] ? P.seeds[0] : ~0ull, s1 = 0;
        for (int64_t r = P.read_begin; r < P.read_end; ++r) {
            const int p = (int)(r / P.reads_per_problem);
            const ProblemDesc D = P.descs[p];
            anneal_read<GROUPS>(D, P, r - D.read_base, f, spw, s0, s1, Mw, st, P.error_flag);
        }
    } else {
        for (;;) {
            if (lane == 0) next_read[wib] = P.read_begin + atomicAdd(P.counter, 1ull);
            __syncwarp();
            const int64_t r = (int64_t)next_read[wib];
            __syncwarp();
            if (r >= P.read_end) break;
            const int p = (int)(r / P.reads_per_problem);
            const ProblemDesc D = P.descs[p];
            const unsigned long long sd = P.seeds[r];
            unsigned long long s0 = sd ? sd : ~0ull, s1 = 0;  // neal: rng_state = {seed ? seed : RANDMAX, 0}
            anneal_read<GROUPS>(D, P, r - D.read_base, f, spw, s0, s1, Mw, st, P.error_flag);
        }
    }
    if (lane == 0) {
        atomicAdd(P.stats + ST_CAND, st.cand);
        atomicAdd(P.stats + ST_DRAWS, st.draws);
        atomicAdd(P.stats + ST_ACC, st.acc);
        atomicAdd(P.stats + ST_NBR, st.nbr);
        atomicAdd(P.stats + ST_ACTIVE, st.active);
        atomicAdd(P.stats + ST_CHUNKS, st.chunks);
        atomicAdd(P.stats + ST_TIES, st.ties);
    }
}

// ------------------------------------------------------------------------------------------------
// lockstep kernels: one warp = 32 reads of the same problem; the lanes walk the variables together, every lane
// owning one read (its own xorshift128+ state, its own spins, its own decisions).  The CSR row of the current
// variable is shared by the warp, so a neighbour update is ONE coalesced 256-byte reduction for up to 32 reads
// instead of 32 scattered 8-byte ones, and exp()/RNG run on all lanes at once.
//   VARIANT 0 "push": local fields f[v][lane] (fp64, read-interleaved) are kept in HBM and updated by
//                     red.global.add.f64 exactly like the warp-per-read kernel -> bit-exact against the oracle.
//   VARIANT 2 "init": one pass that evaluates f[v] = h_v + sum_j J_vj s_j from the bit-packed spins through per-block
//                     tables of distinct spin words (neal's get_flip_energy order) -- the set-up of the push variant.
//   (A former VARIANT 1 re-evaluated the fields at every attempt as a non-bit-exact throughput mode for sparse models; the
//   exact replay kernel overtook it and it was retired, DESIGN.md 4.4.  QA_MODE_THROUGHPUT now means the dense tensor-core
//   kernel of dense.cuh.)
// Spins live in the read-transposed packed layout packedT[word][read] the energy kernel consumes.
// ------------------------------------------------------------------------------------------------
constexpr int QA_LS_TPB = 128;   // threads per block of the lockstep kernels
constexpr int QA_LS_WPB = QA_LS_TPB / 32;
constexpr int QA_LS_D = 16;      // variables per staged block
constexpr int QA_LS_CAP = 384;   // CSR entries staged per block (longer blocks fall back to global loads)
constexpr int QA_LS_CAPW = 32;   // distinct spin words per block held in shared memory (aliases the `cur` staging area)
static_assert(QA_LS_CAPW * sizeof(uint32_t) <= QA_LS_D * sizeof(double), "spin-word staging must fit into the field staging area");

struct LaneStats {
    unsigned int cand, draws, acc, ties;
    unsigned long long nbr;
};

// per-warp staging area in shared memory (double buffered by block parity)
struct LsStage {
    double eJ[2][QA_LS_CAP];                    // couplings of the block's CSR entries
    int ej[2][QA_LS_CAP];                       // neighbour indices
    int rowp[QA_LS_D + 1];
    int gm[QA_LS_D];
    int am[QA_LS_D];
    int pad_;
};
struct LsStagePull {                            // field (re-)evaluation from spins: init pass and pull variant
    double hb[2][QA_LS_D];                      // h of the block's variables
    unsigned int slotw[2][QA_LS_CAP / 2 + 2];   // per-entry (slot | bit << 8) as 16-bit pairs; slot 255 = own word
    int bw[2][QA_LS_CAPW];                      // distinct spin words referenced by the block
};

__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// accept test of one variable for every lane (read) of the warp; returns the per-lane accept flag.
// neal: flip iff exp(-dE*beta) * 2^64 > (double)rand.  The fp64 exp, the 64-bit conversion and the fp64 products are the
// longest dependent chain of a step, so the draw is first screened in SINGLE precision on its upper 32 bits:
//   pa = ex2.approx(float(dE) * float(-beta * log2 e)) is within 3e-5 relative of exp(-dE*beta) for -dE*beta in [-44.4, 0]
//        (two roundings of 6e-8 on an exponent of magnitude <= 64 -> 1.2e-5 absolute in the exponent -> 8e-6, + 2 ulp of
//        ex2.approx);
//   rh = float(rand >> 32): rand / 2^32 lies in [rh', rh' + 1) with rh' = floor(rand / 2^32), |rh - rh'| <= 6e-8 rh'.
// rh + 1 < 0.9999 * pa * 2^32  =>  rand < 0.99994 * exp(..) * 2^64: every fp64 evaluation accepts;
// rh     > 1.0001 * pa * 2^32  =>  rand > 1.00006 * exp(..) * 2^64: every fp64 evaluation rejects.
// Only draws inside that band (2e-4 of them, + 2^-32 / p) take the fp64 path -- the result is identical by construction,
// and near ties (|p - r| <= 2^-48 p) can only occur inside the band, where they are still counted.
__device__ __forceinline__ bool ls_accept(double dE, bool cand, double beta, unsigned long long &s0, unsigned long long &s1,
                                          LaneStats &st) {
    bool acc = cand;
    const bool need = cand && dE > 0.0;
    bool exact = false;
    unsigned long long rnd = 0ull;
    if (need) {
        rnd = rng_next(s0, s1);
        st.draws++;
        float pa;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pa) : "f"(__double2float_rn(dE) * __double2float_rn(beta * -1.4426950408889634)));
        const float rh = __uint2float_rn((unsigned)(rnd >> 32));
        acc = rh + 1.0f < pa * 4294537799.0f;               // 0.9999 * 2^32
        exact = !acc && !(rh > pa * 4295396793.0f);         // 1.0001 * 2^32
    }
    if (__any_sync(FULL_MASK, exact)) {
        if (exact) {
            const double p = exp(-dE * beta) * QA_TWO64;
            double rd;   // volatile: the conversion must stay inside this rare path (the compiler would hoist it)
            asm volatile("cvt.rn.f64.u64 %0, %1;" : "=d"(rd) : "l"(rnd));
            acc = p > rd;
            if (fabs(p - rd) <= p * 3.5527136788005009e-15) st.ties++;
        }
    }
    return acc;
}

struct LsCtx {
    const ProblemDesc &D;
    const AnnealParams &P;
    int64_t r;
    bool active;
    double *fT;
    int *Mcol;
    double *cur;            // smem, per warp [QA_LS_D][32], this lane's column
    LsStage &sg;
    LsStagePull &sp;
    uint32_t *words;        // smem [QA_LS_CAPW][32], this lane's column; aliases `cur` (never live at the same time)
    const double *lam_sh;
    const long long *kap_sh;
};

// h_v + sum over the CSR row of (+-J) in adjacency order, spin words loaded 8 at a time (loads batched, adds sequential)
__device__ __forceinline__ double ls_field_direct(const ProblemDesc &D, const uint32_t *pk, int64_t rpad, int v, int e0, int e1,
                                                  int own_word, uint32_t w_own) {
    double fv = __ldg(D.h + v);
    for (int e = e0; e < e1; e += 8) {
        int jq[8];
        uint32_t wq[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            jq[q] = __ldg(D.col + min(e + q, e1 - 1));
            const int wj = jq[q] >> 5;
            wq[q] = (wj != own_word) ? pk[(int64_t)wj * rpad] : w_own;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (e + q < e1) {
                const double J = __ldg(D.val + e + q);
                fv += ((wq[q] >> (jq[q] & 31)) & 1u) ? J : -J;
            }
        }
    }
    return fv;
}

// Runs the schedule from (bi, swi) on.  VARIANT 0 = push sweeps (bit-exact), 2 = one pass that only evaluates the local fields
// from the spins and stores them (neal get_flip_energy order): the initialisation of the push variant.
template <int VARIANT, bool GROUPS>
__device__ bool ls_sweeps(const LsCtx &c, int &bi, int &swi, bool allow_switch, unsigned long long &s0,
                          unsigned long long &s1, LaneStats &st) {
    const ProblemDesc &D = c.D;
    const AnnealParams &P = c.P;
    LsStage &sg = c.sg;
    LsStagePull &sp = c.sp;
    const int lane = threadIdx.x & 31;
    const int n = D.n;
    const int nch = D.nch;
    const int64_t rpad = D.rpad;
    uint32_t *pk = D.packedT + c.r;
    double *fT = c.fT;
    double *cur = c.cur;
    int *Mcol = c.Mcol;
    const bool active = c.active;
    const bool tables = VARIANT >= 1 && D.bw_ptr != nullptr;
    const int nblk = nch * (32 / QA_LS_D);

    // software pipeline over blocks of QA_LS_D variables:
    //   rows (and, for pull, h / slot bytes / distinct-word list) of block b+1 are copied to shared memory by cp.async while
    //   block b is processed; row pointers and group metadata run two blocks ahead in registers; (push) the local fields
    //   of block b+1 are prefetched into registers and moved to shared memory at the switch.
    auto load_meta = [&](int blk, int &rp, int &gmv, int &amv) {
        const int v = blk * QA_LS_D + lane;
        rp = 0;
        if (lane <= QA_LS_D) rp = __ldg(D.rowptr + v);
        else if (tables && lane <= QA_LS_D + 2) rp = __ldg(D.bw_ptr + blk + (lane - QA_LS_D - 1));
        gmv = -1;
        amv = 0;
        if (GROUPS) {
            if (lane < QA_LS_D) {
                gmv = __ldg(D.grp + v);
                amv = __ldg(D.coef + v);
            }
        }
    };
    auto stage_rows = [&](int buf, int blk, int rp) {
        const int eb = __shfl_sync(FULL_MASK, rp, 0);
        const int ee = __shfl_sync(FULL_MASK, rp, QA_LS_D);
        const int cnt = ee - eb;
        if (cnt <= QA_LS_CAP) {
            for (int k = lane; k < cnt; k += 32) {
                cp_async4(&sg.ej[buf][k], D.col + eb + k);
                cp_async8(&sg.eJ[buf][k], D.val + eb + k);
            }
            if (VARIANT >= 1) {
                if (tables) {
                    const int b0 = __shfl_sync(FULL_MASK, rp, QA_LS_D + 1);
                    const int b1 = __shfl_sync(FULL_MASK, rp, QA_LS_D + 2);
                    for (int k = lane; k < b1 - b0; k += 32) cp_async4(&sp.bw[buf][k], D.bw_words + b0 + k);
                    const int a0 = eb & ~1;                       // 16-bit entries staged as aligned 32-bit words
                    const int nw2 = ((ee + 1) & ~1) - a0;
                    for (int k = lane * 2; k < nw2; k += 64) cp_async4(&sp.slotw[buf][k >> 1], D.ent_slot + a0 + k);
                }
            }
        }
        if (VARIANT >= 1) {
            const int v = blk * QA_LS_D + lane;
            if (lane < QA_LS_D && v < n) cp_async8(&sp.hb[buf][lane], D.h + v);
        }
        cp_async_commit();
    };
    int rp_cur, gm_cur, am_cur, rp_nxt, gm_nxt, am_nxt;
    load_meta(0, rp_cur, gm_cur, am_cur);
    load_meta(nblk > 1 ? 1 : 0, rp_nxt, gm_nxt, am_nxt);
    __syncwarp();
    stage_rows(0, 0, rp_cur);
    double nr[QA_LS_D];
    if (VARIANT == 0) {
#pragma unroll
        for (int i = 0; i < QA_LS_D; ++i) nr[i] = __ldcg(fT + (int64_t)i * 32 + lane);
    }
    int parity = 0;
    bool finished = true;

    const int nbeta = VARIANT == 2 ? 1 : P.num_betas;
    const int nspb = VARIANT == 2 ? 1 : P.sweeps_per_beta;
    for (; bi < nbeta; ++bi, swi = 0) {
        const double beta = VARIANT == 2 ? 1.0 : (D.betas ? D.betas : P.betas)[bi];
        const double thr = 44.36142 / beta;
        for (; swi < nspb; ++swi) {
            uint32_t w = 0;
            bool dirty = false;
            for (int blk = 0; blk < nblk; ++blk) {
                const int v0 = blk * QA_LS_D;
                const int wi = v0 >> 5;
                const int sub = v0 & 31;
                if (sub == 0) {
                    w = pk[(int64_t)wi * rpad];
                    dirty = false;
                }
                int nb = blk + 1;
                if (nb == nblk) nb = 0;
                int nb2 = nb + 1;
                if (nb2 == nblk) nb2 = 0;
                // ---- pipeline: rows of the next block, metadata two blocks ahead, fields of the next block
                __syncwarp();
                stage_rows(parity ^ 1, nb, rp_nxt);
                int rp_nn, gm_nn, am_nn;
                load_meta(nb2, rp_nn, gm_nn, am_nn);
                if (lane <= QA_LS_D) sg.rowp[lane] = rp_cur;
                if (GROUPS) {
                    if (lane < QA_LS_D) {
                        sg.gm[lane] = gm_cur;
                        sg.am[lane] = am_cur;
                    }
                }
                bool stale = false;
                bool blk_dirty = false;
                if (VARIANT == 0) {
#pragma unroll
                    for (int i = 0; i < QA_LS_D; ++i) cur[i * 32] = nr[i];
#pragma unroll
                    for (int i = 0; i < QA_LS_D; ++i) nr[i] = __ldcg(fT + ((int64_t)nb * QA_LS_D + i) * 32 + lane);
                }
                cp_async_wait<1>();
                __syncwarp();
                const int eb = sg.rowp[0];
                const bool staged = (sg.rowp[QA_LS_D] - eb) <= QA_LS_CAP;
                const int *ej = sg.ej[parity];
                const double *eJ = sg.eJ[parity];
                bool tabled = false;
                const unsigned short *slots = nullptr;
                if (VARIANT >= 1) {
                    if (tables && staged) {
                        const int nbw = __shfl_sync(FULL_MASK, rp_cur, QA_LS_D + 2) - __shfl_sync(FULL_MASK, rp_cur, QA_LS_D + 1);
                        tabled = !(nbw == 1 && sp.bw[parity][0] < 0);
                        if (tabled) {
                            // this lane's copy of every distinct spin word the block refers to (its own word stays in `w`)
                            for (int s = 0; s < nbw; s += 8) {
                                uint32_t t[8];
#pragma unroll
                                for (int q = 0; q < 8; ++q) t[q] = (s + q < nbw) ? pk[(int64_t)sp.bw[parity][s + q] * rpad] : 0u;
#pragma unroll
                                for (int q = 0; q < 8; ++q)
                                    if (s + q < nbw) c.words[(s + q) * 32] = t[q];
                            }
                            slots = reinterpret_cast<const unsigned short *>(sp.slotw[parity]) + (eb & 1);
                        }
                    }
                }

                for (int i = 0; i < QA_LS_D; ++i) {
                    const int v = v0 + i;
                    if (v >= n) break;  // uniform: padding variables
                    const int e0 = sg.rowp[i], e1 = sg.rowp[i + 1];
                    const bool up = (w >> (sub + i)) & 1u;
                    double fv;
                    if (VARIANT == 0) {
                        fv = cur[i * 32];
                    } else {
                        // re-evaluate the local field from the spins: h_v + sum_j (+-J) in adjacency order
                        if (tabled) {
                            fv = sp.hb[parity][i];
#pragma unroll 4
                            for (int e = e0; e < e1; ++e) {
                                const int k = e - eb;
                                const unsigned sb = slots[k];             // low byte: slot (255 = own word), high byte: bit
                                const unsigned sl = sb & 255u;
                                const uint32_t wv = (sl == 255u) ? w : c.words[sl * 32];
                                const double J = eJ[k];
                                fv += ((wv >> (sb >> 8)) & 1u) ? J : -J;
                            }
                        } else {
                            fv = ls_field_direct(D, pk, rpad, v, e0, e1, wi, w);
                        }
                        if (VARIANT == 2) {
                            __stcg(fT + (int64_t)v * 32 + lane, fv);
                            continue;
                        }
                    }
                    double dE = up ? -2.0 * fv : 2.0 * fv;
                    int g = -1, a = 0;
                    if (GROUPS) {
                        g = sg.gm[i];
                        if (g >= 0) {
                            a = sg.am[i];
                            const long long t = (long long)a * ((long long)a - (up ? 1 : -1) * ((long long)Mcol[g * QA_LS_TPB] + c.kap_sh[g]));
                            dE = dE + c.lam_sh[g] * (double)t;
                        }
                    }
                    const bool cand = active && !(dE >= thr);
                    if (!__any_sync(FULL_MASK, cand)) continue;
                    if (cand) st.cand++;
                    const bool acc = ls_accept(dE, cand, beta, s0, s1, st);
                    const unsigned accm = __ballot_sync(FULL_MASK, acc);
                    if (accm == 0) continue;
                    if (acc) {
                        st.acc++;
                        st.nbr += (unsigned long long)(e1 - e0);
                    }
                    if (VARIANT == 0) {
                        const double cf = up ? -2.0 : 2.0;  // f[j] += -2*s_v*J  <=> neal dE[j] += 4*s_v*J*s_j
                        double *fTl = fT + lane;
                        const int nbv0 = nb * QA_LS_D;
                        unsigned hit_next = 0;
                        // neighbours inside the staged block are updated in shared memory only (the block is written back
                        // once, below); all others get one predicated fire-and-forget reduction
                        if (staged) {
#pragma unroll 4
                            for (int k = e0 - eb; k < e1 - eb; ++k) {
                                const int j = ej[k];
                                const double d = cf * eJ[k];
                                const unsigned rel = (unsigned)(j - v0);          // uniform
                                if (rel < (unsigned)QA_LS_D) {
                                    if (acc) cur[rel * 32] += d;
                                    blk_dirty = true;
                                } else {
                                    red_add_f64_if(fTl + (int64_t)j * 32, d, acc);
                                    hit_next |= (unsigned)((unsigned)(j - nbv0) < (unsigned)QA_LS_D);
                                }
                            }
                        } else {
                            for (int e = e0; e < e1; ++e) {
                                const int j = __ldg(D.col + e);
                                const double d = cf * __ldg(D.val + e);
                                const unsigned rel = (unsigned)(j - v0);
                                if (rel < (unsigned)QA_LS_D) {
                                    if (acc) cur[rel * 32] += d;
                                    blk_dirty = true;
                                } else {
                                    red_add_f64_if(fTl + (int64_t)j * 32, d, acc);
                                    hit_next |= (unsigned)((unsigned)(j - nbv0) < (unsigned)QA_LS_D);
                                }
                            }
                        }
                        stale = stale || (hit_next != 0);  // prefetched registers of the next block are stale
                    }
                    if (acc) {
                        w ^= 1u << (sub + i);
                        dirty = true;
                        if (GROUPS) {
                            if (g >= 0) Mcol[g * QA_LS_TPB] -= 2 * a * (up ? 1 : -1);
                        }
                    }
                }
                if (VARIANT == 0) {
                    if (blk_dirty) {  // uniform: write the staged fields of this block back (coalesced 256 B rows)
#pragma unroll
                        for (int i = 0; i < QA_LS_D; ++i) __stcg(fT + ((int64_t)v0 + i) * 32 + lane, cur[i * 32]);
                    }
                    if (stale) {  // uniform; rare: a flip touched a variable of the prefetched block
#pragma unroll
                        for (int i = 0; i < QA_LS_D; ++i) nr[i] = __ldcg(fT + ((int64_t)nb * QA_LS_D + i) * 32 + lane);
                    }
                }
                if ((sub + QA_LS_D == 32 || blk == nblk - 1) && dirty) pk[(int64_t)wi * rpad] = w;
                rp_cur = rp_nxt; gm_cur = gm_nxt; am_cur = am_nxt;
                rp_nxt = rp_nn; gm_nxt = gm_nn; am_nxt = am_nn;
                parity ^= 1;
            }
        }
    }
    cp_async_wait<0>();
    __syncwarp();
    return finished;
}

template <int VARIANT, bool GROUPS>
__device__ void lockstep_tile(const LsCtx &c, unsigned long long &s0, unsigned long long &s1, LaneStats &st, int *error_flag) {
    const ProblemDesc &D = c.D;
    const int n = D.n;
    const int nch = D.nch;
    const int64_t rpad = D.rpad;
    const int64_t r = c.r;
    uint32_t *pk = D.packedT + r;  // r < rpad always: padding lanes own a scratch column of packedT

    // ---- pack this read's +-1 bytes (padding lanes and padding variables are +1)
    for (int wi = 0; wi < nch; ++wi) {
        uint32_t w = 0xffffffffu;
        if (c.active) {
            const int8_t *row = D.states + r * (int64_t)n + wi * 32;
            const int lim = min(32, n - wi * 32);
            for (int i = 0; i < lim; ++i) {
                const int s = row[i];
                if (s != 1 && s != -1) atomicExch(error_flag, QA_ERR_STATE);
                if (s < 0) w &= ~(1u << i);
            }
        }
        pk[(int64_t)wi * rpad] = w;
    }
    if (GROUPS) {
        for (int g = 0; g < D.ngroups; ++g) c.Mcol[g * QA_LS_TPB] = 0;
        for (int wi = 0; wi < nch; ++wi) {
            const uint32_t w = pk[(int64_t)wi * rpad];
            for (int i = 0; i < 32; ++i) {
                const int v = wi * 32 + i;
                const int g = __ldg(D.grp + v);  // uniform
                if (g >= 0) {
                    const int a = __ldg(D.coef + v);
                    c.Mcol[g * QA_LS_TPB] += ((w >> i) & 1u) ? a : -a;
                }
            }
        }
    }
    int bi = 0, swi = 0;
    bool finished = false;
    if (!finished) {
        int ib = 0, is = 0;
        ls_sweeps<2, false>(c, ib, is, false, s0, s1, st);  // local fields from the current spins
        ls_sweeps<0, GROUPS>(c, bi, swi, false, s0, s1, st);
    }

    // ---- final spins back to the caller's +-1 rows
    if (c.active) {
        for (int wi = 0; wi < nch; ++wi) {
            const uint32_t w = pk[(int64_t)wi * rpad];
            int8_t *row = D.states + r * (int64_t)n + wi * 32;
            const int lim = min(32, n - wi * 32);
            for (int i = 0; i < lim; ++i) row[i] = ((w >> i) & 1u) ? 1 : -1;
        }
    }
}

__host__ __device__ inline size_t ls_smem_bytes(int max_groups) {
    size_t b = sizeof(double) * QA_LS_D * QA_LS_TPB;   // cur (push phases) / spin words (field evaluation phases)
    b += (sizeof(LsStage) + sizeof(LsStagePull)) * QA_LS_WPB;
    b += (sizeof(double) + sizeof(long long)) * (size_t)max_groups;
    b += sizeof(int) * (size_t)max_groups * QA_LS_TPB;
    return b;
}

template <int VARIANT, bool GROUPS>
__global__ void __launch_bounds__(QA_LS_TPB, 3) k_anneal_lockstep(AnnealParams P) {
    extern __shared__ __align__(16) unsigned char ls_smem[];
    // layout: [cur: D x TPB doubles, aliased by the spin words] [LsStage x warps] [LsStagePull x warps] [lambda] [kappa] [M]
    unsigned char *sp = ls_smem;
    double *cur_all = reinterpret_cast<double *>(sp);
    sp += sizeof(double) * QA_LS_D * QA_LS_TPB;
    LsStage *stages = reinterpret_cast<LsStage *>(sp);
    sp += sizeof(LsStage) * QA_LS_WPB;
    LsStagePull *pstages = reinterpret_cast<LsStagePull *>(sp);
    sp += sizeof(LsStagePull) * QA_LS_WPB;
    uint32_t *words_all = reinterpret_cast<uint32_t *>(cur_all);
    double *lam_sh = reinterpret_cast<double *>(sp);
    sp += sizeof(double) * P.max_groups;
    long long *kap_sh = reinterpret_cast<long long *>(sp);
    sp += sizeof(long long) * P.max_groups;
    int *M_all = reinterpret_cast<int *>(sp);
    __shared__ unsigned long long next_tile[QA_LS_WPB];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int64_t slot = (int64_t)blockIdx.x * QA_LS_WPB + wib;
    double *fT = P.fT_scratch ? P.fT_scratch + slot * P.fT_stride : nullptr;
    if (GROUPS) {  // groups exist only on single-problem models: one copy of lambda / kappa per block
        const ProblemDesc &D0 = P.descs[0];
        for (int g = threadIdx.x; g < D0.ngroups; g += blockDim.x) {
            lam_sh[g] = D0.lambda[g];
            kap_sh[g] = D0.kappa[g];
        }
    }
    __syncthreads();
    LaneStats st = {0, 0, 0, 0, 0};
    for (;;) {
        if (lane == 0) next_tile[wib] = atomicAdd(P.counter, 1ull);
        __syncwarp();
        const int64_t tile = (int64_t)next_tile[wib];
        __syncwarp();
        if (tile >= P.total_tiles) break;
        const int p = (int)(tile / P.tiles_per_problem);
        const int64_t tip = tile % P.tiles_per_problem;
        const ProblemDesc D = P.descs[p];
        const int64_t r = tip * 32 + lane;
        const bool active = r < D.reads;
        const unsigned long long sd = active ? P.seeds[D.read_base + r] : 1ull;
        unsigned long long s0 = sd ? sd : ~0ull, s1 = 0;
        const LsCtx c = {D, P, r, active, fT, M_all + threadIdx.x, cur_all + wib * (QA_LS_D * 32) + lane, stages[wib],
                         pstages[wib], words_all + wib * (QA_LS_CAPW * 32) + lane, lam_sh, kap_sh};
        lockstep_tile<VARIANT, GROUPS>(c, s0, s1, st, P.error_flag);
    }
    // warp-reduce the per-lane counters
    unsigned long long v[5] = {st.cand, st.draws, st.acc, st.ties, st.nbr};
#pragma unroll
    for (int q = 0; q < 5; ++q)
        for (int off = 16; off > 0; off >>= 1) v[q] += __shfl_xor_sync(FULL_MASK, v[q], off);
    if (lane == 0) {
        atomicAdd(P.stats + ST_CAND, v[0]);
        atomicAdd(P.stats + ST_DRAWS, v[1]);
        atomicAdd(P.stats + ST_ACC, v[2]);
        atomicAdd(P.stats + ST_TIES, v[3]);
        atomicAdd(P.stats + ST_NBR, v[4]);
    }
}

#include "replay.cuh"
#include "dense.cuh"

// ------------------------------------------------------------------------------------------------
// energies: neal get_state_energy(), one thread per read, reads on lanes (coalesced packedT loads)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_energy(const ProblemDesc *descs) {
    const ProblemDesc D = descs[blockIdx.y];
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= D.reads) return;
    const uint32_t *pk = D.packedT + r;
    const int64_t stride = D.rpad;
    double E = 0.0;
    // rank-1 group sums M_g = sum a_v s_v ride along with the linear pass (one pass over the spins for ALL groups)
    long long M[QA_MAX_GROUPS];
    for (int g = 0; g < D.ngroups; ++g) M[g] = 0;
    for (int c = 0; c < D.nch; ++c) {
        const uint32_t w = pk[c * stride];
        const int base = c * 32;
        const int lim = min(32, D.n - base);
        for (int i = 0; i < lim; ++i) {
            const double hv = __ldg(D.h + base + i);
            E += ((w >> i) & 1u) ? hv : -hv;  // state[v]*h[v]
        }
        if (D.ngroups) {
            for (int i = 0; i < lim; ++i) {
                const int g = __ldg(D.grp + base + i);   // uniform
                if (g >= 0) {
                    const long long a = __ldg(D.coef + base + i);
                    M[g] += ((w >> i) & 1u) ? a : -a;
                }
            }
        }
    }
    // couplers in the caller's order: the additions are one dependent chain, the (random-row) spin loads are not -- issue
    // them eight couplers at a time
    int64_t e = 0;
    for (; e + 8 <= D.m; e += 8) {
        uint32_t bu[8], bv[8];
        double wt[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int u = __ldg(D.starts + e + q), v = __ldg(D.ends + e + q);
            wt[q] = __ldg(D.w + e + q);
            bu[q] = pk[(int64_t)(u >> 5) * stride] >> (u & 31);
            bv[q] = pk[(int64_t)(v >> 5) * stride] >> (v & 31);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) E += ((bu[q] ^ bv[q]) & 1u) ? -wt[q] : wt[q];  // state[u]*w*state[v]
    }
    for (; e < D.m; ++e) {
        const int u = __ldg(D.starts + e), v = __ldg(D.ends + e);
        const double wt = __ldg(D.w + e);
        const uint32_t bu = (pk[(int64_t)(u >> 5) * stride] >> (u & 31)) & 1u;
        const uint32_t bv = (pk[(int64_t)(v >> 5) * stride] >> (v & 31)) & 1u;
        E += (bu ^ bv) ? -wt : wt;
    }
    for (int g = 0; g < D.ngroups; ++g) {
        const long long t = M[g] + D.kappa[g];
        E += D.lambda[g] * (double)(t * t) * 0.25;
    }
    D.energies[r] = E;
}

// pack caller-provided +-1 states into the read-transposed layout (for qa_energy_argmin)
__global__ void k_pack_states(ProblemDesc D, int *error_flag) {
    const int lane = threadIdx.x & 31;
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= D.reads) return;
    const int8_t *row = D.states + r * (int64_t)D.n;
    for (int c = 0; c < D.nch; ++c) {
        const int v = c * 32 + lane;
        int s = 1;
        if (v < D.n) {
            s = row[v];
            if (s != 1 && s != -1) atomicExch(error_flag, QA_ERR_STATE);
        }
        const uint32_t w = __ballot_sync(FULL_MASK, s > 0);
        if (lane == 0) D.packedT[(int64_t)c * D.rpad + r] = w;
    }
}

// warp-shuffle min / argmin (lowest index wins ties, like a stable sort by energy -> SampleSet.first)
__global__ void __launch_bounds__(1024) k_argmin(const double *energies, int64_t count, double *best_e, long long *best_i) {
    __shared__ double se[32];
    __shared__ long long si[32];
    double e = INFINITY;
    long long idx = 0x7fffffffffffffffll;
    for (int64_t i = threadIdx.x; i < count; i += blockDim.x) {
        const double x = energies[i];
        if (x < e || (x == e && i < idx)) { e = x; idx = i; }
    }
    for (int off = 16; off > 0; off >>= 1) {
        const double oe = __shfl_xor_sync(FULL_MASK, e, off);
        const long long oi = __shfl_xor_sync(FULL_MASK, idx, off);
        if (oe < e || (oe == e && oi < idx)) { e = oe; idx = oi; }
    }
    if ((threadIdx.x & 31) == 0) { se[threadIdx.x >> 5] = e; si[threadIdx.x >> 5] = idx; }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int nw = (blockDim.x + 31) >> 5;
        e = threadIdx.x < nw ? se[threadIdx.x] : INFINITY;
        idx = threadIdx.x < nw ? si[threadIdx.x] : 0x7fffffffffffffffll;
        for (int off = 16; off > 0; off >>= 1) {
            const double oe = __shfl_xor_sync(FULL_MASK, e, off);
            const long long oi = __shfl_xor_sync(FULL_MASK, idx, off);
            if (oe < e || (oe == e && oi < idx)) { e = oe; idx = oi; }
        }
        if (threadIdx.x == 0) { *best_e = e; *best_i = idx; }
    }
}

// ------------------------------------------------------------------------------------------------
// adjacency construction on the device, preserving neal's push_back order (stable sort by vertex)
// ------------------------------------------------------------------------------------------------
__global__ void k_make_entries(int64_t m_total, int32_t num_problems, const int64_t *var_off, const int64_t *cpl_off,
                               const int32_t *starts, const int32_t *ends, uint32_t *keys, uint32_t *vals, int *error_flag) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m_total) return;
    // problem of coupler c (binary search in coupler offsets)
    int lo = 0, hi = num_problems;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (cpl_off[mid] <= c) lo = mid; else hi = mid;
    }
    const int64_t vbase = var_off[lo];
    const int64_t np = var_off[lo + 1] - vbase;
    const int u = starts[c], v = ends[c];
    if (u < 0 || v < 0 || u >= np || v >= np || u == v) {
        atomicExch(error_flag, QA_ERR_INDEX);
        keys[2 * c] = keys[2 * c + 1] = 0;
        vals[2 * c] = (uint32_t)(2 * c);
        vals[2 * c + 1] = (uint32_t)(2 * c + 1);
        return;
    }
    keys[2 * c] = (uint32_t)(vbase + u);      // entry 2c   : row u, neighbour v
    keys[2 * c + 1] = (uint32_t)(vbase + v);  // entry 2c+1 : row v, neighbour u
    vals[2 * c] = (uint32_t)(2 * c);
    vals[2 * c + 1] = (uint32_t)(2 * c + 1);
}

__global__ void k_fill_csr(int64_t entries, const uint32_t *sorted_vals, const int32_t *starts, const int32_t *ends,
                           const double *w, int32_t *col, double *val) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= entries) return;
    const uint32_t e = sorted_vals[i];
    const int64_t c = e >> 1;
    col[i] = (e & 1u) ? starts[c] : ends[c];
    val[i] = w[c];
}

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

// ------------------------------------------------------------------------------------------------
// model builders on the device (reference: the Python Q-dict loops of BQM_clustering.py:36-47, 228-236,
// DQM_clustering.py:29-43, CQM_clustering.py:30-48, QA_subsampling.py:26-35 followed by dimod's
// from_qubo / change_vartype(SPIN) / to_numpy_vectors).  Every floating-point accumulation keeps the order of the
// Python code: per-vertex sums run sequentially over the vertex's edges in G.edges order, h accumulates Q_ij/4 in
// ascending-neighbour order, so the vectors are bit-identical to models.py (tests/test_gpu_builders.py).
// ------------------------------------------------------------------------------------------------
__global__ void k_graph_entries(int64_t m, int32_t n, const int32_t *eu, const int32_t *ev, uint32_t *keys, uint32_t *vals,
                                int *error_flag) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= m) return;
    const int u = eu[e], v = ev[e];
    if (u < 0 || v < 0 || u >= n || v >= n || u == v) {
        atomicExch(error_flag, QA_ERR_INDEX);
        keys[2 * e] = keys[2 * e + 1] = 0;
    } else {
        keys[2 * e] = (uint32_t)u;
        keys[2 * e + 1] = (uint32_t)v;
    }
    vals[2 * e] = (uint32_t)(2 * e);
    vals[2 * e + 1] = (uint32_t)(2 * e + 1);
}

// out[v] = base + sum over v's edges, in edge order, of f(w_e):  mode 0: scale*w   1: scale*(1-w)   2: 1 (degree)
// mode 3: weight of the LAST edge touching v (base if isolated)  -- the `set_linear` overwrite of DQM_clustering.py:42-43
__global__ void k_vertex_accumulate(int32_t n, const int32_t *rowptr, const uint32_t *sorted_vals, const double *w, int mode,
                                    double scale, double base, double *out) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    double acc = mode == 3 ? base : 0.0;
    for (int i = rowptr[v]; i < rowptr[v + 1]; ++i) {
        const double we = w[sorted_vals[i] >> 1];
        if (mode == 0) acc += scale * we;
        else if (mode == 1) acc += scale * (1 - we);
        else if (mode == 2) acc += 1.0;
        else acc = we;
    }
    out[v] = (mode == 3) ? acc : acc + base;
}

__global__ void k_seq_sum(int64_t count, const double *x, double scale, double *out) {
    // deterministic left-to-right sum (python / numpy cumsum order); setup only
    if (blockIdx.x || threadIdx.x) return;
    double s = 0.0;
    for (int64_t i = 0; i < count; ++i) s += x[i] * scale;
    *out = s;
}

// QUBO couplers of the k-way models: K copies of every graph edge (same case) + the one-hot pairs of every cell
__global__ void k_kway_couplers(int64_t m, int32_t n, int32_t K, const int32_t *eu, const int32_t *ev, const double *w,
                                double edge_scale, double edge_pre, double edge_shift, double onehot_q, unsigned long long *keys,
                                double *q) {
    const int64_t npairs = (int64_t)K * (K - 1) / 2;
    const int64_t total = m * K + (int64_t)n * npairs;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    int64_t a, b;
    double val;
    if (t < m * K) {
        const int64_t e = t / K;
        const int p = (int)(t % K);
        a = (int64_t)eu[e] * K + p;
        b = (int64_t)ev[e] * K + p;
        val = (edge_pre + edge_scale * w[e]) + edge_shift;  // python: (2g - 2w) - 2g, or -2w - 2g, or -2w
    } else {
        const int64_t u = t - m * K;
        const int64_t i = u / npairs;
        int64_t pr = u % npairs;
        int ca = 0;
        while (pr >= K - 1 - ca) { pr -= K - 1 - ca; ++ca; }
        const int cb = ca + 1 + (int)pr;
        a = i * K + ca;
        b = i * K + cb;
        val = onehot_q;
    }
    const unsigned long long hi = (unsigned long long)max(a, b), lo = (unsigned long long)min(a, b);
    keys[t] = (hi << 32) | lo;
    q[t] = val;
}

__global__ void k_edge_couplers(int64_t m, const int32_t *eu, const int32_t *ev, const double *w, int mode, double scale,
                                unsigned long long *keys, double *q) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= m) return;
    const unsigned long long hi = (unsigned long long)max(eu[e], ev[e]), lo = (unsigned long long)min(eu[e], ev[e]);
    keys[e] = (hi << 32) | lo;
    q[e] = mode == 0 ? scale * w[e] : scale * (1 - w[e]);
}

__global__ void k_split_keys(int64_t m, const unsigned long long *keys, const uint32_t *perm, const double *q, int32_t *r,
                             int32_t *c, double *J) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    r[i] = (int32_t)(keys[i] >> 32);
    c[i] = (int32_t)(keys[i] & 0xffffffffu);
    J[i] = q[perm[i]] / 4.0;  // dimod change_vartype(SPIN): J = Q_ij / 4
}

__global__ void k_kway_linear(int32_t n, int32_t K, int32_t nvar, const double *cell_lin, double shift, double *lin) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nvar) return;
    lin[v] = v < n * K ? cell_lin[v / K] + shift : 0.0;
}

// h_v = Q_vv/2 + sum over the row (ascending neighbour = coupler order) of Q_vj/4
__global__ void k_h_from_rows(int32_t n, const double *lin, const int32_t *rowptr, const double *val, double *h) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    double acc = lin[v] / 2.0;
    for (int e = rowptr[v]; e < rowptr[v + 1]; ++e) acc += val[e];
    h[v] = acc;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
};

}  // namespace

struct qa_ctx {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    DevBuf f, spw, fT, sf, states, energies, seeds, betas, packed, misc, cubtmp;
    int kernel = 0;  // QA_KERNEL_*
    int replay_switch_permille = 20;  // replay -> push hand-over threshold (QA_REPLAY_SWITCH_PERMILLE overrides)
    int replay_warps = 0;             // warps per CTA of the replay kernel (0: automatic; QA_REPLAY_WARPS overrides)
    int last_kernel = 0;              // QA_KERNEL_* the last sampling call ran on
    unsigned rp_smem_base = 1024;     // shared-window offset of dynamic shared memory (verified by the replay kernel)
    bool rp_base_checked = false;
    bool betas_per_problem = false;   // transient: the running call carries one beta schedule per problem ([P][num_betas])
    int *h_iflag = nullptr, *d_iflag = nullptr;  // host-mapped interrupt flag polled by the replay kernel
    unsigned long long *d_stats = nullptr;  // QA_NSTAT counters + 1 read counter
    int *d_flag = nullptr;
    double *d_best_e = nullptr;
    long long *d_best_i = nullptr;
    uint32_t launches = 0;
};

struct qa_model {
    qa_ctx *ctx = nullptr;
    int32_t num_problems = 1;
    int64_t n_total = 0;   // sum of n_p
    int64_t m_total = 0;
    int32_t n_max = 0, nch_max = 0;
    int32_t max_deg = 0;
    int32_t ngroups = 0;
    std::vector<int64_t> var_off, cpl_off;  // host copies [P+1]
    // device arrays (owned)
    double *h = nullptr;
    int32_t *starts = nullptr, *ends = nullptr;
    double *w = nullptr;
    int32_t *rowptr = nullptr, *col = nullptr;
    double *val = nullptr;
    int32_t *grp = nullptr, *coef = nullptr;
    double *lambda = nullptr;
    long long *kappa = nullptr;
    ProblemDesc *d_descs = nullptr;
    std::vector<ProblemDesc> descs;  // host mirror (pointers are device pointers)
    // block word tables of the pull variant (built on first use)
    int32_t *bw_ptr = nullptr, *bw_words = nullptr;
    unsigned short *ent_slot = nullptr;
    bool tables_built = false;
    // coupling slabs of the replay kernel (built on first use; rp_ok = the model fits the slab format)
    unsigned char *rp_slabs = nullptr;
    uint32_t *rp_off = nullptr;
    bool rp_built = false, rp_ok = false;
    bool rp_uniform = false;   // every block holds exactly RP_D variables
    int rp_slots = 32;         // half-word slots per warp the slabs were packed for (32 or 64)
    bool rp_adj_sorted = false; // adjacency lists ascending: field set-up through the slab ring
    bool groups_i32 = false;   // every group term a*(a - s*(M+kappa)) fits 32-bit integers
    // dense k-way form (dense.cuh): W and P derived from the CSR by qa_model_enable_dense
    double *dn_W = nullptr;
    bool dn_ok = false;
    DenseDesc dn = {};
};

namespace {

int ensure(DevBuf &b, size_t bytes) {
    if (bytes <= b.bytes && b.p) return QA_OK;
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.bytes = 0;
    if (bytes == 0) bytes = 256;
    QA_CUDA(cudaMalloc(&b.p, bytes));
    b.bytes = bytes;
    return QA_OK;
}

void release(DevBuf &b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.bytes = 0;
}

bool is_device_ptr(const void *p) {
    if (!p) return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

float elapsed(cudaEvent_t a, cudaEvent_t b) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

template <typename T>
int upload(qa_ctx *ctx, T **dst, const T *src, size_t count) {
    *dst = nullptr;
    QA_CUDA(cudaMalloc((void **)dst, std::max<size_t>(count, 1) * sizeof(T)));
    if (count)
        QA_CUDA(cudaMemcpyAsync(*dst, src, count * sizeof(T), cudaMemcpyDefault, ctx->stream));
    return QA_OK;
}

// build CSR (all problems at once) from the device COO already stored in the model
int build_adjacency(qa_model *M) {
    qa_ctx *ctx = M->ctx;
    const int64_t m = M->m_total;
    const int64_t entries = 2 * m;
    const int64_t rows_alloc = M->n_total + 64 + 1;  // padding rows read by the last chunk of the last problem
    if (entries >= (int64_t)0x7fffffff) return fail(QA_ERR_LIMIT, "more than 2^30 couplers: use a structured (group) model");
    if (M->n_total >= (int64_t)0x7fffffff) return fail(QA_ERR_LIMIT, "too many variables");
    QA_CUDA(cudaMalloc((void **)&M->rowptr, rows_alloc * sizeof(int32_t)));
    QA_CUDA(cudaMalloc((void **)&M->col, std::max<int64_t>(entries, 1) * sizeof(int32_t)));
    QA_CUDA(cudaMalloc((void **)&M->val, std::max<int64_t>(entries, 1) * sizeof(double)));
    QA_CUDA(cudaMemsetAsync(ctx->d_flag, 0, 2 * sizeof(int), ctx->stream));
    int64_t *d_off = nullptr;
    const int P = M->num_problems;
    QA_CUDA(cudaMalloc((void **)&d_off, 2 * (P + 1) * sizeof(int64_t)));
    QA_CUDA(cudaMemcpyAsync(d_off, M->var_off.data(), (P + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    QA_CUDA(cudaMemcpyAsync(d_off + P + 1, M->cpl_off.data(), (P + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    uint32_t *keys = nullptr, *vals = nullptr, *keys2 = nullptr, *vals2 = nullptr;
    int rc = QA_OK;
    if (entries > 0) {
        int rc2 = ensure(ctx->misc, (size_t)entries * 4 * sizeof(uint32_t));
        if (rc2) { cudaFree(d_off); return rc2; }
        keys = (uint32_t *)ctx->misc.p;
        vals = keys + entries;
        keys2 = vals + entries;
        vals2 = keys2 + entries;
        const int tpb = 256;
        k_make_entries<<<(unsigned)((m + tpb - 1) / tpb), tpb, 0, ctx->stream>>>(m, P, d_off, d_off + P + 1, M->starts, M->ends,
                                                                                 keys, vals, ctx->d_flag);
        ctx->launches++;
        int end_bit = 1;
        while (((int64_t)1 << end_bit) < M->n_total + 1 && end_bit < 32) ++end_bit;
        size_t tmp_bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, vals, vals2, (int)entries, 0, end_bit, ctx->stream);
        rc2 = ensure(ctx->cubtmp, tmp_bytes);
        if (rc2) { cudaFree(d_off); return rc2; }
        cudaError_t ce = cub::DeviceRadixSort::SortPairs(ctx->cubtmp.p, tmp_bytes, keys, keys2, vals, vals2, (int)entries, 0,
                                                         end_bit, ctx->stream);
        if (ce != cudaSuccess) { cudaFree(d_off); return fail(QA_ERR_CUDA, std::string("radix sort: ") + cudaGetErrorString(ce)); }
        ctx->launches += 4;
        k_fill_csr<<<(unsigned)((entries + tpb - 1) / tpb), tpb, 0, ctx->stream>>>(entries, vals2, M->starts, M->ends, M->w, M->col, M->val);
        ctx->launches++;
    } else {
        int rc2 = ensure(ctx->misc, 256);
        if (rc2) { cudaFree(d_off); return rc2; }
        keys2 = (uint32_t *)ctx->misc.p;
    }
    {
        const int tpb = 256;
        k_rowptr<<<(unsigned)((rows_alloc + tpb - 1) / tpb), tpb, 0, ctx->stream>>>(rows_alloc, entries, keys2, M->rowptr, ctx->d_flag + 1);
        ctx->launches++;
    }
    int flags[2] = {0, 0};
    cudaError_t ce = cudaMemcpyAsync(flags, ctx->d_flag, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_off);
    if (ce != cudaSuccess) return fail(QA_ERR_CUDA, std::string("adjacency build: ") + cudaGetErrorString(ce));
    if (flags[0] != 0) return fail(QA_ERR_INDEX, "coupler index out of range or self-loop");
    M->max_deg = flags[1];
    return rc;
}

int finalize_descs(qa_model *M) {
    const int P = M->num_problems;
    M->descs.assign(P, ProblemDesc());
    M->n_max = 0;
    for (int p = 0; p < P; ++p) {
        ProblemDesc &D = M->descs[p];
        const int64_t v0 = M->var_off[p], c0 = M->cpl_off[p];
        D.n = (int32_t)(M->var_off[p + 1] - v0);
        D.nch = (D.n + 31) / 32;
        D.ngroups = 0;
        D.m = M->cpl_off[p + 1] - c0;
        D.rowptr = M->rowptr + v0;
        D.col = M->col;   // rowptr holds global entry positions
        D.val = M->val;
        D.h = M->h + v0;
        D.starts = M->starts + c0;
        D.ends = M->ends + c0;
        D.w = M->w + c0;
        D.grp = nullptr; D.coef = nullptr; D.lambda = nullptr; D.kappa = nullptr;
        D.bw_ptr = nullptr; D.bw_words = nullptr; D.ent_slot = nullptr;
        D.rp_slabs = nullptr; D.rp_off = nullptr; D.rp_nslabs = 0;
        D.betas = nullptr;
        M->n_max = std::max(M->n_max, D.n);
    }
    M->nch_max = (M->n_max + 31) / 32;
    if (!M->d_descs) QA_CUDA(cudaMalloc((void **)&M->d_descs, P * sizeof(ProblemDesc)));
    return QA_OK;
}

int model_create(qa_ctx *ctx, int32_t P, const int64_t *var_off, const int64_t *cpl_off, const double *h,
                 const int32_t *starts, const int32_t *ends, const double *w, qa_model **out) {
    qa_model *M = new qa_model();
    M->ctx = ctx;
    M->num_problems = P;
    M->var_off.assign(var_off, var_off + P + 1);
    M->cpl_off.assign(cpl_off, cpl_off + P + 1);
    M->n_total = var_off[P];
    M->m_total = cpl_off[P];
    int rc = upload(ctx, &M->h, h, (size_t)M->n_total);
    if (!rc) rc = upload(ctx, &M->starts, starts, (size_t)M->m_total);
    if (!rc) rc = upload(ctx, &M->ends, ends, (size_t)M->m_total);
    if (!rc) rc = upload(ctx, &M->w, w, (size_t)M->m_total);
    if (!rc) rc = build_adjacency(M);
    if (!rc) rc = finalize_descs(M);
    if (rc) {
        qa_model_destroy(M);
        return rc;
    }
    *out = M;
    return QA_OK;
}

// Block word tables for the pull variant: per block of QA_LS_D variables the distinct spin words (32 variables each) its
// CSR rows refer to, and per CSR entry the slot of its word in that list.  Built once per model on the host from the
// device-built CSR (a setup step, O(entries)); blocks that exceed the shared-memory capacities get a -1 sentinel.
int build_word_tables(qa_model *M) {
    if (M->tables_built) return QA_OK;
    qa_ctx *ctx = M->ctx;
    const int64_t entries = 2 * M->m_total;
    const int64_t rows_alloc = M->n_total + 64 + 1;
    std::vector<int32_t> rowptr(rows_alloc), col(std::max<int64_t>(entries, 1));
    QA_CUDA(cudaMemcpy(rowptr.data(), M->rowptr, rows_alloc * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (entries) QA_CUDA(cudaMemcpy(col.data(), M->col, entries * sizeof(int32_t), cudaMemcpyDeviceToHost));
    std::vector<int32_t> bw_ptr, bw_words;
    std::vector<unsigned short> slot(std::max<int64_t>(entries, 1) + 8, 255);
    std::vector<int64_t> blk_base(M->num_problems + 1, 0);
    bw_ptr.push_back(0);
    std::vector<int32_t> stamp, slot_of;
    for (int p = 0; p < M->num_problems; ++p) {
        const int64_t v_off = M->var_off[p];
        const int n = (int)(M->var_off[p + 1] - v_off);
        const int nch = (n + 31) / 32;
        const int nblk = nch * (32 / QA_LS_D);
        blk_base[p] = (int64_t)bw_ptr.size() - 1;
        stamp.assign(nch, -1);
        slot_of.assign(nch, 0);
        for (int b = 0; b < nblk; ++b) {
            const int v0 = b * QA_LS_D;
            const int64_t eb = rowptr[v_off + std::min(v0, n)];
            const int64_t ee = rowptr[v_off + std::min(v0 + QA_LS_D, n)];
            const size_t first = bw_words.size();
            const int own = v0 >> 5;
            bool ok = (ee - eb) <= QA_LS_CAP;
            for (int64_t e = eb; e < ee && ok; ++e) {
                const int wj = col[e] >> 5;
                const unsigned short bit = (unsigned short)((col[e] & 31) << 8);
                if (wj == own) { slot[e] = 255 | bit; continue; }
                if (stamp[wj] != b) {
                    if (bw_words.size() - first >= (size_t)QA_LS_CAPW) { ok = false; break; }
                    stamp[wj] = b;
                    slot_of[wj] = (int32_t)(bw_words.size() - first);
                    bw_words.push_back(wj);
                }
                slot[e] = (unsigned short)slot_of[wj] | bit;
            }
            if (!ok) {
                bw_words.resize(first);
                bw_words.push_back(-1);  // sentinel: block not tabled, kernel falls back to direct loads
                for (int w = 0; w < nch; ++w) if (stamp[w] == b) stamp[w] = -1;
            }
            bw_ptr.push_back((int32_t)bw_words.size());
        }
        // one extra pointer per problem so that the two-ahead metadata loads of the last block stay in range
    }
    blk_base[M->num_problems] = (int64_t)bw_ptr.size() - 1;
    for (int k = 0; k < 4; ++k) bw_ptr.push_back((int32_t)bw_words.size());
    if (bw_words.empty()) bw_words.push_back(-1);
    QA_CUDA(cudaMalloc((void **)&M->bw_ptr, bw_ptr.size() * sizeof(int32_t)));
    QA_CUDA(cudaMalloc((void **)&M->bw_words, bw_words.size() * sizeof(int32_t)));
    QA_CUDA(cudaMalloc((void **)&M->ent_slot, slot.size() * sizeof(unsigned short)));
    QA_CUDA(cudaMemcpyAsync(M->bw_ptr, bw_ptr.data(), bw_ptr.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    QA_CUDA(cudaMemcpyAsync(M->bw_words, bw_words.data(), bw_words.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    QA_CUDA(cudaMemcpyAsync(M->ent_slot, slot.data(), slot.size() * sizeof(unsigned short), cudaMemcpyHostToDevice, ctx->stream));
    QA_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int p = 0; p < M->num_problems; ++p) {
        M->descs[p].bw_ptr = M->bw_ptr + blk_base[p];
        M->descs[p].bw_words = M->bw_words;
        M->descs[p].ent_slot = M->ent_slot;
    }
    M->tables_built = true;
    return QA_OK;
}

// Coupling slabs of the replay kernel (replay.cuh): per block one contiguous {RpHdr, RpEntry[]} record.  Consecutive variables are packed greedily into blocks of at most RP_D variables inside
// ONE 16-variable half-word, RP_MAXBW foreign half-words and RP_CAP entry slots.  A row's entries are in REPLAY order --
// neighbours u > v ascending, then u < v ascending (stable, so duplicate couplers keep their adjacency order) -- split
// into the PRE part (u > v, then u < v0: known when the block starts; padded to rounds of 4 with
// zero-coupling entries) and the SEQ part (v0 <= u < v: decided inside the block).  The slab holds the pre parts of all
// rows, then the seq parts of all rows.  Built once per model on the host from the device-built CSR (a setup step,
// O(entries)).  Models whose blocks would hold fewer than 4 variables on average (dense rows, or sparse rows scattered over
// many half-words) leave rp_ok false and run on the other kernels.
// Pure host part of the slab construction (no CUDA calls: the CPU test-suite drives it through qa_debug_pack_slabs).
// rowptr holds global entry positions per global row (problem p's rows start at var_off[p]), col local neighbour indices.
struct RpPacked {
    std::vector<uint32_t> off;          // slab offsets in 16-byte units, one terminator
    std::vector<unsigned char> slabs;
    std::vector<int64_t> blk_base;      // first slab of every problem in `off`
    std::vector<int32_t> nslabs;
    bool uniform = true;                // every block holds exactly RP_D variables
    bool adj_sorted = true;             // every adjacency list is ascending (dimod's vector order): slab order serves the set-up
};

bool pack_replay_slabs(int P, const int64_t *var_off, const int32_t *rowptr, const int32_t *col, const double *val, int ngroups,
                       const std::vector<int32_t> &hg, const std::vector<int32_t> &hc, int slots, RpPacked &out) {
    const int RP_MAXBW = slots - 1;   // foreign half-words per block
    std::vector<uint32_t> &off = out.off;
    std::vector<unsigned char> &slabs = out.slabs;
    out.blk_base.assign(P, 0);
    out.nslabs.assign(P, 0);
    off.clear();
    slabs.clear();
    out.uniform = true;
    out.adj_sorted = true;
    struct Nb { int32_t j; double J; };
    struct Row { std::vector<Nb> later, early, seq; };   // u > v | u < v0 | v0 <= u < v
    std::vector<int32_t> stamp, slot_of;
    std::vector<Row> rows(RP_D);
    std::vector<size_t> hdr_pos;
    std::vector<int32_t> row_words;
    for (int p = 0; p < P; ++p) {
        const int64_t v_off = var_off[p];
        const int n = (int)(var_off[p + 1] - v_off);
        if (n == 0 || n > RP_MAX_VARS) return false;   // the entry word holds a 20-bit neighbour index
        const int nhw = (n + 15) / 16;
        const int npad = nhw * 16;
        out.blk_base[p] = (int64_t)off.size();
        hdr_pos.clear();
        stamp.assign(nhw, -1);
        slot_of.assign(nhw, 0);
        int b = 0;   // block (slab) index inside the problem; doubles as the stamp of the half-word -> slot map
        for (int v0 = 0; v0 < npad; ++b) {
            const int own = v0 >> 4;
            const int vmax = std::min(v0 + RP_D, (own + 1) * 16);   // never across a half-word
            RpHdr H;
            memset(&H, 0, sizeof(H));
            for (int i = 0; i < RP_D; ++i) H.ga[i] = 255;
            int nbw = 0, nv = 0;
            size_t slots_pre = 0, slots_seq = 0;
            for (int v = v0; v < vmax; ++v) {
                Row &R = rows[nv];
                R.later.clear();
                R.early.clear();
                R.seq.clear();
                if (v < n) {
                    for (int64_t e = rowptr[v_off + v]; e < rowptr[v_off + v + 1]; ++e) {
                        const Nb nb{col[e], val[e]};
                        if (nb.j > v) R.later.push_back(nb);
                        else if (nb.j < v0) R.early.push_back(nb);
                        else R.seq.push_back(nb);
                        if (e > rowptr[v_off + v] && col[e] < col[e - 1]) out.adj_sorted = false;
                    }
                }
                auto by_index = [](const Nb &a, const Nb &b2) { return a.j < b2.j; };
                std::stable_sort(R.later.begin(), R.later.end(), by_index);
                std::stable_sort(R.early.begin(), R.early.end(), by_index);
                std::stable_sort(R.seq.begin(), R.seq.end(), by_index);
                const size_t npre = R.later.size() + R.early.size();
                const size_t deg = npre + R.seq.size();
                if (R.seq.size() > 255 || deg > 0xffff) return false;
                // does the row still fit into this block?  (entry slots, and the foreign half-words it would add)
                row_words.clear();
                for (int part = 0; part < 2; ++part)
                    for (const Nb &nb : (part == 0 ? R.later : R.early)) {
                        const int wj = nb.j >> 4;
                        if (wj != own && stamp[wj] != b && std::find(row_words.begin(), row_words.end(), wj) == row_words.end())
                            row_words.push_back(wj);
                    }
                const size_t pre_slots = (npre + 3) / 4 * 4;
                const bool fits = slots_pre + slots_seq + pre_slots + R.seq.size() <= (size_t)RP_CAP &&
                                  nbw + (int)row_words.size() <= RP_MAXBW && pre_slots / 4 <= 255;
                if (!fits) {
                    if (nv == 0) return false;   // a single row exceeds the format: dense model
                    break;
                }
                for (int wj : row_words) {
                    stamp[wj] = b;
                    slot_of[wj] = nbw + 1;
                    H.bw[nbw++] = wj;
                }
                const int i = nv++;
                if (v < n && ngroups > 0 && hg[v] >= 0) {
                    if (hc[v] >= (1 << 23) || hc[v] <= -(1 << 23)) return false;  // coefficient does not fit the packed form
                    H.ga[i] = (int32_t)((uint32_t)hg[v] | ((uint32_t)hc[v] << 8));
                }
                H.rowa[i] = (uint32_t)(pre_slots / 4) | ((uint32_t)R.seq.size() << 8) | ((uint32_t)deg << 16);
                H.rowb[i] = (uint32_t)R.later.size() | ((uint32_t)npre << 16);
                slots_pre += pre_slots;
                slots_seq += R.seq.size();
            }
            // the block is closed: lay out the pre region (rows in order, padded), then the seq region
            std::vector<RpEntry> E;
            E.reserve(slots_pre + slots_seq);
            auto emit = [&](const Nb &nb) {
                const int wj = nb.j >> 4;
                RpEntry en;
                en.J = nb.J;
                en.zero = 0u;
                en.B = (uint32_t)(30 - 2 * (nb.j & 15)) | ((uint32_t)(wj == own ? 0 : slot_of[wj]) << 7) | ((uint32_t)nb.j << RP_J_SHIFT);
                E.push_back(en);
            };
            for (int i = 0; i < nv; ++i) {
                for (const Nb &nb : rows[i].later) emit(nb);
                for (const Nb &nb : rows[i].early) emit(nb);
                while (E.size() % 4) {
                    RpEntry en;
                    en.J = 0.0;          // fma(0, sigma, f) == f
                    en.zero = 0u;
                    en.B = 30u | ((uint32_t)(v0 + i) << RP_J_SHIFT);   // slot 0
                    E.push_back(en);
                }
            }
            H.seq_off = (int32_t)E.size();
            for (int i = 0; i < nv; ++i)
                for (const Nb &nb : rows[i].seq) emit(nb);
            H.nent = (int32_t)E.size();
            H.nbw = nbw;
            H.v0 = v0;
            H.nv = nv;
            out.uniform = out.uniform && nv == RP_D;
            if (slabs.size() / 16 > 0xfffffff0ull) return false;
            off.push_back((uint32_t)(slabs.size() / 16));
            hdr_pos.push_back(slabs.size());
            const unsigned char *hp = reinterpret_cast<const unsigned char *>(&H);
            slabs.insert(slabs.end(), hp, hp + sizeof(H));
            const unsigned char *ep = reinterpret_cast<const unsigned char *>(E.data());
            slabs.insert(slabs.end(), ep, ep + E.size() * sizeof(RpEntry));
            v0 += nv;
        }
        out.nslabs[p] = b;
        if ((int64_t)b * 4 > (int64_t)npad) return false;   // fewer than 4 variables per block on average: replaying does not pay
        // every slab also carries the half-word list of the next block (cyclic), staged while this block decides, and the slot
        // that holds the PREVIOUS block's half-word (the one word that staging cannot have up to date)
        for (size_t k = 0; k < hdr_pos.size(); ++k) {
            RpHdr *cur = reinterpret_cast<RpHdr *>(slabs.data() + hdr_pos[k]);
            const RpHdr *nxt = reinterpret_cast<const RpHdr *>(slabs.data() + hdr_pos[(k + 1) % hdr_pos.size()]);
            cur->nbw_next = nxt->nbw;
            memcpy(cur->bw_next, nxt->bw, sizeof(cur->bw_next));
            cur->prev_slot = 0;
            if (k > 0) {
                const RpHdr *prv = reinterpret_cast<const RpHdr *>(slabs.data() + hdr_pos[k - 1]);
                const int phw = prv->v0 >> 4;
                if (phw != (cur->v0 >> 4))
                    for (int s = 0; s < cur->nbw; ++s)
                        if (cur->bw[s] == phw) cur->prev_slot = s + 1;
            }
        }
    }
    off.push_back((uint32_t)(slabs.size() / 16));
    return !slabs.empty();
}

int build_replay_tables(qa_model *M) {
    if (M->rp_built) return QA_OK;
    qa_ctx *ctx = M->ctx;
    M->rp_built = true;
    M->rp_ok = false;
    const int64_t entries = 2 * M->m_total;
    const int64_t rows_alloc = M->n_total + 64 + 1;
    QA_CUDA(cudaStreamSynchronize(ctx->stream));
    std::vector<int32_t> rowptr(rows_alloc), col(std::max<int64_t>(entries, 1));
    std::vector<double> val(std::max<int64_t>(entries, 1));
    QA_CUDA(cudaMemcpy(rowptr.data(), M->rowptr, rows_alloc * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (entries) {
        QA_CUDA(cudaMemcpy(col.data(), M->col, entries * sizeof(int32_t), cudaMemcpyDeviceToHost));
        QA_CUDA(cudaMemcpy(val.data(), M->val, entries * sizeof(double), cudaMemcpyDeviceToHost));
    }
    std::vector<int32_t> hg, hc;
    if (M->ngroups > 0) {
        const int64_t npad = (int64_t)M->descs[0].nch * 32;
        hg.resize(npad);
        hc.resize(npad);
        QA_CUDA(cudaMemcpy(hg.data(), M->grp, npad * sizeof(int32_t), cudaMemcpyDeviceToHost));
        QA_CUDA(cudaMemcpy(hc.data(), M->coef, npad * sizeof(int32_t), cudaMemcpyDeviceToHost));
    }
    RpPacked pk;
    // 32 half-word slots per warp when the blocks fit (4 KB per warp), else 64 (scattered neighbourhoods)
    M->rp_slots = 32;
    if (!pack_replay_slabs(M->num_problems, M->var_off.data(), rowptr.data(), col.data(), val.data(), M->ngroups, hg, hc, 32, pk)) {
        M->rp_slots = 64;
        if (!pack_replay_slabs(M->num_problems, M->var_off.data(), rowptr.data(), col.data(), val.data(), M->ngroups, hg, hc, 64, pk))
            return QA_OK;   // the model does not fit the slab format: rp_ok stays false
    }
    const std::vector<uint32_t> &off = pk.off;
    const std::vector<unsigned char> &slabs = pk.slabs;
    const std::vector<int64_t> &blk_base = pk.blk_base;
    const std::vector<int32_t> &nslabs = pk.nslabs;
    const bool uniform = pk.uniform;
    M->rp_adj_sorted = pk.adj_sorted;
    QA_CUDA(cudaMalloc((void **)&M->rp_slabs, slabs.size()));
    QA_CUDA(cudaMalloc((void **)&M->rp_off, off.size() * sizeof(uint32_t)));
    QA_CUDA(cudaMemcpyAsync(M->rp_slabs, slabs.data(), slabs.size(), cudaMemcpyHostToDevice, ctx->stream));
    QA_CUDA(cudaMemcpyAsync(M->rp_off, off.data(), off.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    QA_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int p = 0; p < M->num_problems; ++p) {
        M->descs[p].rp_slabs = M->rp_slabs;
        M->descs[p].rp_off = M->rp_off + blk_base[p];
        M->descs[p].rp_nslabs = nslabs[p];
    }
    M->rp_ok = true;
    M->rp_uniform = uniform;
    return QA_OK;
}

struct RunBuffers {
    int8_t *d_states = nullptr;
    double *d_energies = nullptr;
    bool states_on_host = false, energies_on_host = false;
};

int check_schedule(int32_t num_betas, const double *betas, int32_t sweeps_per_beta) {
    if (num_betas < 0 || sweeps_per_beta < 0) return fail(QA_ERR_ARG, "negative schedule size");
    if (num_betas > 0 && !betas) return fail(QA_ERR_ARG, "beta_schedule is null");
    return QA_OK;
}

// core: anneal all reads of all problems of a model; states/energies already on the device
int run_anneal(qa_ctx *ctx, qa_model *M, int32_t reads_per_problem, int8_t *d_states, double *d_energies,
               int32_t num_betas, const double *d_betas, int32_t sweeps_per_beta, const unsigned long long *d_seeds,
               int32_t seed_mode, int32_t mode, qa_interrupt_fn interrupt, void *iuser, qa_stats *st, int64_t *completed) {
    const int P = M->num_problems;
    const int64_t total_reads = (int64_t)P * reads_per_problem;
    const int32_t rpad = (reads_per_problem + 31) & ~31;
    *completed = 0;
    if (total_reads == 0) return QA_OK;

    // packed transposed spins for the energy kernel
    size_t packed_words = 0;
    for (int p = 0; p < P; ++p) packed_words += (size_t)M->descs[p].nch * rpad;
    int rc = ensure(ctx->packed, std::max<size_t>(packed_words, 1) * sizeof(uint32_t));
    if (rc) return rc;
    {
        size_t off = 0;
        int64_t st_off = 0;
        for (int p = 0; p < P; ++p) {
            ProblemDesc &D = M->descs[p];
            D.reads = reads_per_problem;
            D.rpad = rpad;
            D.read_base = (int64_t)p * reads_per_problem;
            D.states = d_states + st_off;
            D.packedT = (uint32_t *)ctx->packed.p + off;
            D.energies = d_energies + (int64_t)p * reads_per_problem;
            D.betas = ctx->betas_per_problem ? d_betas + (int64_t)p * num_betas : nullptr;
            off += (size_t)D.nch * rpad;
            st_off += (int64_t)reads_per_problem * D.n;
        }
        QA_CUDA(cudaMemcpyAsync(M->d_descs, M->descs.data(), P * sizeof(ProblemDesc), cudaMemcpyHostToDevice, ctx->stream));
    }

    const bool groups = M->ngroups > 0;
    // kernel choice: warp-per-read (any read count, stream seeding) or lockstep (32 reads per warp)
    // QA_MODE_THROUGHPUT: a model with the dense k-way form runs the tensor-core kernel (tolerance parity); every other model
    // runs the reference-order kernels below -- for sparse models the exact replay kernel is faster than any recompute or
    // graph-coloured update, and bit-exact results meet the statistical bar trivially
    int kernel = QA_KERNEL_WARP_PER_READ;
    if (mode == QA_MODE_THROUGHPUT && M->dn_ok && P == 1 && !interrupt && seed_mode == QA_SEED_PER_READ) kernel = QA_KERNEL_DENSE;
    else if (ctx->kernel == QA_KERNEL_LOCKSTEP_PUSH && seed_mode == QA_SEED_PER_READ) kernel = QA_KERNEL_LOCKSTEP_PUSH;
    else if (ctx->kernel == QA_KERNEL_AUTO && seed_mode == QA_SEED_PER_READ && (int64_t)reads_per_problem >= 32 &&
             2 * total_reads >= (int64_t)ctx->num_sms * 32 * 7)
        kernel = QA_KERNEL_LOCKSTEP_PUSH;
    // replay kernel (deferred exact updates): sparse models whose blocks fit the slab format; explicit choice, or automatic
    // from 6144 reads on (measured on B200, config 3: 1.26e10 vs 7.5e9 attempts/s for the warp-per-read kernel at 12 500
    // reads, 4.3e9 vs 5.9e9 at 4096)
    // an interrupt callback: the warp-per-read kernel runs in read waves and polls between them; the replay kernel polls a
    // host-mapped flag whenever a CTA pulls its next group of reads; the lockstep push kernel has no stopping point
    if (interrupt && kernel == QA_KERNEL_LOCKSTEP_PUSH) kernel = QA_KERNEL_WARP_PER_READ;
    // batched models with rank-1 groups (qa_model_concat): the lockstep and replay kernels keep one copy of lambda / kappa per
    // CTA, so those run on the warp-per-read kernel, which reads them per problem
    if (P > 1 && groups) kernel = QA_KERNEL_WARP_PER_READ;
    if (kernel != QA_KERNEL_DENSE && seed_mode == QA_SEED_PER_READ && (!interrupt || P == 1) && !(P > 1 && groups) &&
        (ctx->kernel == QA_KERNEL_REPLAY ||
         (ctx->kernel == QA_KERNEL_AUTO && (int64_t)reads_per_problem >= 32 && total_reads >= 6144))) {
        rc = build_replay_tables(M);
        if (rc) return rc;
        if (M->rp_ok) {
            kernel = QA_KERNEL_REPLAY;
            QA_CUDA(cudaMemcpyAsync(M->d_descs, M->descs.data(), P * sizeof(ProblemDesc), cudaMemcpyHostToDevice, ctx->stream));
        } else if (ctx->kernel == QA_KERNEL_REPLAY) {
            kernel = interrupt ? QA_KERNEL_WARP_PER_READ : QA_KERNEL_LOCKSTEP_PUSH;  // dense model: the slab format does not apply
        }
    }

    ctx->last_kernel = kernel;
    QA_CUDA(cudaMemsetAsync(ctx->d_stats, 0, (QA_NSTAT + 1 + QA_NDEBUG) * sizeof(unsigned long long), ctx->stream));
    QA_CUDA(cudaMemsetAsync(ctx->d_flag, 0, 2 * sizeof(int), ctx->stream));

    AnnealParams A;
    memset(&A, 0, sizeof(A));
    A.descs = M->d_descs;
    A.num_problems = P;
    A.reads_per_problem = reads_per_problem;
    A.total_reads = total_reads;
    A.betas = d_betas;
    A.num_betas = num_betas;
    A.sweeps_per_beta = sweeps_per_beta;
    A.seeds = d_seeds;
    A.seed_mode = seed_mode;
    A.counter = ctx->d_stats + QA_NSTAT;
    A.stats = ctx->d_stats;
    A.error_flag = ctx->d_flag;

    int64_t done = 0;
    bool interrupted = false;
    if (kernel == QA_KERNEL_WARP_PER_READ) {
        // launch geometry: persistent grid, one warp per resident read
        int wpb = QA_TPB_MAX / 32;
        if (seed_mode == QA_SEED_STREAM) wpb = 1;
        else if (total_reads < (int64_t)ctx->num_sms * wpb) wpb = (int)std::max<int64_t>(1, (total_reads + ctx->num_sms - 1) / ctx->num_sms);
        const int tpb = wpb * 32;
        int bps = 0;
        if (groups) QA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_anneal_ref<true>, tpb, 0));
        else QA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_anneal_ref<false>, tpb, 0));
        if (bps < 1) return fail(QA_ERR_CUDA, "annealing kernel does not fit on an SM");
        int64_t grid = (int64_t)bps * ctx->num_sms;  // multiple of the SM count
        const int64_t need = (total_reads + wpb - 1) / wpb;
        if (need < grid) grid = need;
        if (seed_mode == QA_SEED_STREAM) grid = 1;
        const int64_t slots = grid * wpb;
        const int64_t f_stride = (int64_t)M->nch_max * 32;
        const int64_t spw_stride = ((int64_t)M->nch_max + 31) & ~31ll;
        rc = ensure(ctx->f, (size_t)slots * f_stride * sizeof(double));
        if (!rc) rc = ensure(ctx->spw, (size_t)slots * spw_stride * sizeof(uint32_t));
        if (rc) return rc;
        A.f_scratch = (double *)ctx->f.p;
        A.spw_scratch = (uint32_t *)ctx->spw.p;
        A.f_stride = f_stride;
        A.spw_stride = spw_stride;

        // without an interrupt callback the whole job is one launch; with one, read waves of `slots` reads
        const int64_t wave = (interrupt && seed_mode != QA_SEED_STREAM) ? slots : total_reads;
        QA_CUDA(cudaEventRecord(ctx->ev[2], ctx->stream));
        while (done < total_reads) {
            A.read_begin = done;
            A.read_end = std::min(total_reads, done + wave);
            QA_CUDA(cudaMemsetAsync(A.counter, 0, sizeof(unsigned long long), ctx->stream));
            if (groups) k_anneal_ref<true><<<(unsigned)grid, tpb, 0, ctx->stream>>>(A);
            else k_anneal_ref<false><<<(unsigned)grid, tpb, 0, ctx->stream>>>(A);
            QA_CUDA(cudaGetLastError());
            ctx->launches++;
            if (st) st->anneal_launches++;
            done = A.read_end;
            if (interrupt && done < total_reads) {
                QA_CUDA(cudaStreamSynchronize(ctx->stream));
                if (interrupt(iuser)) { interrupted = true; break; }
            }
        }
    } else if (kernel == QA_KERNEL_DENSE) {
        // dense k-way: one warp = 32 reads, fields of a block of 8 cells by fp64 tensor-core MMAs over all cells (dense.cuh)
        const int K = M->dn.K;
        const int warps = 4;
        const void *fn = K == 1 ? (const void *)k_anneal_dense<1> : K == 2 ? (const void *)k_anneal_dense<2>
                       : K == 4 ? (const void *)k_anneal_dense<4> : (const void *)k_anneal_dense<8>;
        const size_t smem = K == 1 ? dn_smem_bytes<1>(warps) : K == 2 ? dn_smem_bytes<2>(warps)
                          : K == 4 ? dn_smem_bytes<4>(warps) : dn_smem_bytes<8>(warps);
        QA_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int bps = 0;
        QA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessorWithFlags(&bps, fn, warps * 32, smem, cudaOccupancyDefault));
        if (bps < 1) return fail(QA_ERR_CUDA, "dense kernel does not fit on an SM");
        const int64_t total_tiles = (reads_per_problem + 31) / 32;
        // every warp pulls 32-read tiles from a counter; no more CTAs than the tiles can fill (warps without a tile leave at once)
        const int64_t grid = std::min<int64_t>((int64_t)bps * ctx->num_sms, (total_tiles + warps - 1) / warps);
        const int64_t stride = (int64_t)M->dn.ngrp * 32 * K;
        rc = ensure(ctx->sf, (size_t)grid * warps * stride * sizeof(uint32_t));
        if (rc) return rc;
        A.dn_spins = (uint32_t *)ctx->sf.p;
        A.dn_stride = stride;
        A.total_tiles = total_tiles;
        A.read_begin = 0;
        A.read_end = total_reads;
        QA_CUDA(cudaEventRecord(ctx->ev[2], ctx->stream));
        QA_CUDA(cudaMemsetAsync(A.counter, 0, sizeof(unsigned long long), ctx->stream));
        void *args[] = {&A, &M->dn};
        QA_CUDA(cudaLaunchKernel(fn, dim3((unsigned)grid), dim3(warps * 32), args, smem, ctx->stream));
        QA_CUDA(cudaGetLastError());
        ctx->launches++;
        if (st) st->anneal_launches++;
        done = total_reads;
    } else if (kernel == QA_KERNEL_REPLAY) {
        // replay: one CTA = `nw` consecutive 32-read tiles of one problem, coupling slabs shared through a TMA ring
        const int tpp = (reads_per_problem + 31) / 32;
        const int mg = std::max(M->ngroups, 1);
        int nw = ctx->replay_warps;
        if (nw < 1 || nw > RP_MAX_WARPS) {
            nw = 1;
            while (nw < RP_MAX_WARPS && nw < tpp) nw *= 2;
            while (nw > 1 && rp_smem_bytes(nw, mg, ctx->rp_smem_base, M->rp_slots) > (size_t)(226 * 1024 / RP_MIN_CTAS - 1024)) nw /= 2;   // RP_MIN_CTAS CTAs per SM
            // few tiles: prefer narrower CTAs on every SM to full CTAs on some of them
            while (nw > 1 && (int64_t)P * ((tpp + nw - 1) / nw) < 2 * (int64_t)ctx->num_sms) nw /= 2;
        }
        const int64_t gpp = (tpp + nw - 1) / nw;
        const int64_t total_items = (int64_t)P * gpp;
        size_t smem = rp_smem_bytes(nw, mg, ctx->rp_smem_base, M->rp_slots);
        const void *fn = nullptr;
        if (M->rp_slots == 32)
            fn = !groups ? (const void *)k_anneal_replay<0, 32>
                         : (M->groups_i32 ? (const void *)k_anneal_replay<1, 32> : (const void *)k_anneal_replay<2, 32>);
        else
            fn = !groups ? (const void *)k_anneal_replay<0, 64>
                         : (M->groups_i32 ? (const void *)k_anneal_replay<1, 64> : (const void *)k_anneal_replay<2, 64>);
        QA_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int bps = 0;
        QA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessorWithFlags(&bps, fn, nw * 32, smem, cudaOccupancyDefault));
        if (bps < 1) return fail(QA_ERR_CUDA, "replay kernel does not fit on an SM");
        int64_t grid = std::min<int64_t>((int64_t)bps * ctx->num_sms, total_items);
        const int64_t fT_stride = (int64_t)M->nch_max * 32 * 32;
        const int64_t sf_stride = (int64_t)M->nch_max * 2 * 32;
        {
            size_t free_b = 0, total_b = 0;
            QA_CUDA(cudaMemGetInfo(&free_b, &total_b));
            const size_t per_slot = (size_t)fT_stride * sizeof(double) + (size_t)sf_stride * sizeof(uint32_t);
            const size_t budget = (size_t)((double)(free_b + ctx->fT.bytes + ctx->sf.bytes) * 0.9);
            const int64_t max_slots = (int64_t)(budget / per_slot);
            if (max_slots < nw) return fail(QA_ERR_CUDA, "not enough device memory for one CTA of local fields");
            if (grid * nw > max_slots) grid = max_slots / nw;
            rc = ensure(ctx->fT, (size_t)grid * nw * fT_stride * sizeof(double) + 4096);   // + the look-ahead rows of the last slot
            if (!rc) rc = ensure(ctx->sf, (size_t)grid * nw * sf_stride * sizeof(uint32_t));
            if (rc) return rc;
        }
        A.fT_scratch = (double *)ctx->fT.p;
        A.fT_stride = fT_stride;
        A.sf_scratch = ctx->sf.p;
        A.sf_stride = sf_stride;
        A.tiles_per_problem = tpp;
        A.groups_per_problem = gpp;
        A.total_items = total_items;
        A.max_groups = mg;
        A.switch_permille = ctx->replay_switch_permille;
        A.rp_slab_init = M->rp_adj_sorted ? 1 : 0;
        A.read_begin = 0;
        A.read_end = total_reads;
        if (interrupt) {
            *ctx->h_iflag = 0;
            A.interrupt_flag = ctx->d_iflag;
        }
        QA_CUDA(cudaEventRecord(ctx->ev[2], ctx->stream));
        QA_CUDA(cudaMemsetAsync(A.counter, 0, sizeof(unsigned long long), ctx->stream));
        void *args[] = {&A};
        QA_CUDA(cudaLaunchKernel(fn, dim3((unsigned)grid), dim3(nw * 32), args, smem, ctx->stream));
        QA_CUDA(cudaGetLastError());
        ctx->launches++;
        {   // the kernel leaves at once if dynamic shared memory does not start where the layout assumed: relaunch with the real base
            int fl[2] = {0, 0};
            QA_CUDA(cudaMemcpyAsync(fl, ctx->d_flag, sizeof(fl), cudaMemcpyDeviceToHost, ctx->stream));
            if (!ctx->rp_base_checked) {   // first launch of this context only (costs a synchronisation)
                QA_CUDA(cudaStreamSynchronize(ctx->stream));
                if (fl[0] == QA_ERR_SMEM_BASE) {
                    ctx->rp_smem_base = (unsigned)fl[1];
                    smem = rp_smem_bytes(nw, mg, ctx->rp_smem_base, M->rp_slots);
                    QA_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    QA_CUDA(cudaMemsetAsync(ctx->d_flag, 0, 2 * sizeof(int), ctx->stream));
                    QA_CUDA(cudaMemsetAsync(A.counter, 0, sizeof(unsigned long long), ctx->stream));
                    QA_CUDA(cudaEventRecord(ctx->ev[2], ctx->stream));
                    QA_CUDA(cudaLaunchKernel(fn, dim3((unsigned)grid), dim3(nw * 32), args, smem, ctx->stream));
                    QA_CUDA(cudaGetLastError());
                    ctx->launches++;
                }
                ctx->rp_base_checked = true;
            }
        }
        if (st) st->anneal_launches++;
        done = total_reads;
        if (interrupt) {   // poll the callback while the launch runs; CTAs stop pulling work once the flag is up
            QA_CUDA(cudaEventRecord(ctx->ev[3], ctx->stream));
            for (;;) {
                const cudaError_t q = cudaEventQuery(ctx->ev[3]);
                if (q == cudaSuccess) break;
                if (q != cudaErrorNotReady) return fail(QA_ERR_CUDA, std::string("replay kernel: ") + cudaGetErrorString(q));
                if (!interrupted && interrupt(iuser)) {
                    *reinterpret_cast<volatile int *>(ctx->h_iflag) = 1;
                    interrupted = true;
                }
                struct timespec ts = {0, 200000};
                nanosleep(&ts, nullptr);
            }
            if (interrupted) {   // groups are handed out in read order: the first `pulled` groups are complete
                unsigned long long pulled = 0;
                QA_CUDA(cudaMemcpyAsync(&pulled, A.counter, sizeof(pulled), cudaMemcpyDeviceToHost, ctx->stream));
                QA_CUDA(cudaStreamSynchronize(ctx->stream));
                done = std::min<int64_t>(total_reads, (int64_t)std::min<unsigned long long>(pulled, (unsigned long long)total_items) * nw * 32);
            }
        }
    } else {
        // lockstep: one warp = 32 reads of one problem
        const int tpp = (reads_per_problem + 31) / 32;
        const int64_t total_tiles = (int64_t)P * tpp;
        const size_t smem = ls_smem_bytes(std::max(M->ngroups, 1));
        const void *fn = nullptr;
        fn = groups ? (const void *)k_anneal_lockstep<0, true> : (const void *)k_anneal_lockstep<0, false>;
        QA_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int bps = 0;
        QA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessorWithFlags(&bps, fn, QA_LS_TPB, smem, cudaOccupancyDefault));
        if (bps < 1) return fail(QA_ERR_CUDA, "lockstep kernel does not fit on an SM");
        const int wpb = QA_LS_TPB / 32;
        int64_t grid = (int64_t)bps * ctx->num_sms;
        const int64_t need = (total_tiles + wpb - 1) / wpb;
        // spread few tiles over all SMs: prefer more blocks with idle warps to fewer full blocks
        if (need < grid) grid = std::min<int64_t>(grid, std::max<int64_t>(need, std::min<int64_t>(total_tiles, (int64_t)ctx->num_sms)));
        const int64_t fT_stride = (int64_t)M->nch_max * 32 * 32;
        {
            // read-interleaved local fields for the resident tiles (push kernel; push phase of the throughput mode)
            size_t free_b = 0, total_b = 0;
            QA_CUDA(cudaMemGetInfo(&free_b, &total_b));
            const size_t per_slot = (size_t)fT_stride * sizeof(double);
            const size_t budget = (size_t)((double)(free_b + ctx->fT.bytes) * 0.85);
            int64_t max_slots = (int64_t)(budget / per_slot);
            if (max_slots < wpb) return fail(QA_ERR_CUDA, "not enough device memory for one block of local fields");
            if (grid * wpb > max_slots) grid = max_slots / wpb;
            rc = ensure(ctx->fT, (size_t)grid * wpb * per_slot);
            if (rc) return rc;
            A.fT_scratch = (double *)ctx->fT.p;
        }
        rc = build_word_tables(M);  // field evaluation from spins (init pass of push, pull variant) runs on these
        if (rc) return rc;
        QA_CUDA(cudaMemcpyAsync(M->d_descs, M->descs.data(), P * sizeof(ProblemDesc), cudaMemcpyHostToDevice, ctx->stream));
        A.fT_stride = fT_stride;
        A.tiles_per_problem = tpp;
        A.total_tiles = total_tiles;
        A.max_groups = std::max(M->ngroups, 1);
        A.read_begin = 0;
        A.read_end = total_reads;
        QA_CUDA(cudaEventRecord(ctx->ev[2], ctx->stream));
        QA_CUDA(cudaMemsetAsync(A.counter, 0, sizeof(unsigned long long), ctx->stream));
        void *args[] = {&A};
        QA_CUDA(cudaLaunchKernel(fn, dim3((unsigned)grid), dim3(QA_LS_TPB), args, smem, ctx->stream));
        QA_CUDA(cudaGetLastError());
        ctx->launches++;
        if (st) st->anneal_launches++;
        done = total_reads;
    }
    QA_CUDA(cudaEventRecord(ctx->ev[3], ctx->stream));
    if (kernel != QA_KERNEL_DENSE) {   // the dense kernel evaluates the energies itself (one more field pass)
        dim3 g((unsigned)((reads_per_problem + 127) / 128), (unsigned)P);
        k_energy<<<g, 128, 0, ctx->stream>>>(M->d_descs);
        QA_CUDA(cudaGetLastError());
        ctx->launches++;
    }
    QA_CUDA(cudaEventRecord(ctx->ev[4], ctx->stream));
    int flag = 0;
    unsigned long long hs[QA_NSTAT + 1 + QA_NDEBUG];
    QA_CUDA(cudaMemcpyAsync(&flag, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    QA_CUDA(cudaMemcpyAsync(hs, ctx->d_stats, sizeof(hs), cudaMemcpyDeviceToHost, ctx->stream));
    QA_CUDA(cudaStreamSynchronize(ctx->stream));
#ifdef QA_RP_PROFILE
    fprintf(stderr, "[qa profile] cycles summed over warps: setup %llu slab-wait %llu stage-wait %llu prologue %llu entries %llu decide %llu pass %llu push-flip %llu\n",
            hs[QA_NSTAT + 1], hs[QA_NSTAT + 2], hs[QA_NSTAT + 3], hs[QA_NSTAT + 4], hs[QA_NSTAT + 5], hs[QA_NSTAT + 6], hs[QA_NSTAT + 7], hs[QA_NSTAT + 8]);
#endif
    if (flag == QA_ERR_SMEM_BASE) return fail(QA_ERR_CUDA, "replay kernel: dynamic shared memory does not start where the layout assumed");
    if (flag != 0) return fail(flag, "initial states must be +1/-1");
    if (st) {
        st->candidates += hs[ST_CAND];
        st->draws += hs[ST_DRAWS];
        st->accepted += hs[ST_ACC];
        st->nbr_updates += hs[ST_NBR];
        st->active_chunks += hs[ST_ACTIVE];
        st->chunks += hs[ST_CHUNKS];
        st->near_ties += hs[ST_TIES];
        st->ms_anneal += elapsed(ctx->ev[2], ctx->ev[3]);
        st->ms_energy += elapsed(ctx->ev[3], ctx->ev[4]);
        uint64_t att = 0;
        for (int p = 0; p < P; ++p) att += (uint64_t)M->descs[p].n;
        st->attempts += att * (uint64_t)num_betas * (uint64_t)sweeps_per_beta * (uint64_t)(done / P);
    }
    *completed = done;
    return QA_OK;
}

// stage caller buffers (host or device), run, and return results
int sample_common(qa_ctx *ctx, qa_model *M, int32_t reads_per_problem, int8_t *states_inout, double *energies_out,
                  int32_t num_betas, const double *beta_schedule, int32_t sweeps_per_beta, const uint64_t *seeds,
                  int32_t seed_mode, int32_t mode, qa_interrupt_fn interrupt, void *iuser, qa_stats *stats_out) {
    if (!ctx || !M) return fail(QA_ERR_ARG, "null context or model");
    if (M->ctx != ctx) return fail(QA_ERR_ARG, "model belongs to another context");
    if (reads_per_problem < 0) return fail(QA_ERR_ARG, "negative num_reads");
    if (mode != QA_MODE_REFERENCE && mode != QA_MODE_THROUGHPUT) return fail(QA_ERR_ARG, "unknown mode");
    if (seed_mode != QA_SEED_PER_READ && seed_mode != QA_SEED_STREAM) return fail(QA_ERR_ARG, "unknown seed_mode");
    if (seed_mode == QA_SEED_STREAM && M->num_problems != 1) return fail(QA_ERR_ARG, "stream seeding needs a single problem");
    int rc = check_schedule(num_betas, beta_schedule, sweeps_per_beta);
    if (rc) return rc;
    QA_CUDA(cudaSetDevice(ctx->device));
    const int64_t total_reads = (int64_t)M->num_problems * reads_per_problem;
    qa_stats st;
    memset(&st, 0, sizeof(st));
    const uint32_t launches0 = ctx->launches;
    if (total_reads == 0) {
        if (stats_out) *stats_out = st;
        return 0;
    }
    if (!states_inout || !energies_out || !seeds) return fail(QA_ERR_ARG, "null states/energies/seeds");
    const int64_t state_bytes = (int64_t)reads_per_problem * M->n_total;

    QA_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
    int8_t *d_states = states_inout;
    const bool st_host = !is_device_ptr(states_inout);
    if (st_host) {
        rc = ensure(ctx->states, (size_t)std::max<int64_t>(state_bytes, 1));
        if (rc) return rc;
        d_states = (int8_t *)ctx->states.p;
        QA_CUDA(cudaMemcpyAsync(d_states, states_inout, (size_t)state_bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    double *d_energies = energies_out;
    const bool en_host = !is_device_ptr(energies_out);
    if (en_host) {
        rc = ensure(ctx->energies, (size_t)total_reads * sizeof(double));
        if (rc) return rc;
        d_energies = (double *)ctx->energies.p;
    }
    const int64_t nseeds = seed_mode == QA_SEED_STREAM ? 1 : total_reads;
    const unsigned long long *d_seeds = (const unsigned long long *)seeds;
    if (!is_device_ptr(seeds)) {
        rc = ensure(ctx->seeds, (size_t)nseeds * sizeof(uint64_t));
        if (rc) return rc;
        QA_CUDA(cudaMemcpyAsync(ctx->seeds.p, seeds, (size_t)nseeds * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
        d_seeds = (const unsigned long long *)ctx->seeds.p;
    }
    const double *d_betas = beta_schedule;
    if (num_betas > 0 && !is_device_ptr(beta_schedule)) {
        const size_t nb = (size_t)num_betas * (ctx->betas_per_problem ? (size_t)M->num_problems : 1);
        rc = ensure(ctx->betas, nb * sizeof(double));
        if (rc) return rc;
        QA_CUDA(cudaMemcpyAsync(ctx->betas.p, beta_schedule, nb * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        d_betas = (const double *)ctx->betas.p;
    }
    QA_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));

    int64_t completed = 0;
    rc = run_anneal(ctx, M, reads_per_problem, d_states, d_energies, num_betas, d_betas, sweeps_per_beta, d_seeds, seed_mode,
                    mode, interrupt, iuser, &st, &completed);
    if (rc) return rc;

    QA_CUDA(cudaEventRecord(ctx->ev[4], ctx->stream));
    if (st_host) QA_CUDA(cudaMemcpyAsync(states_inout, d_states, (size_t)state_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (en_host) QA_CUDA(cudaMemcpyAsync(energies_out, d_energies, (size_t)total_reads * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    QA_CUDA(cudaEventRecord(ctx->ev[5], ctx->stream));
    QA_CUDA(cudaStreamSynchronize(ctx->stream));
    st.ms_h2d += elapsed(ctx->ev[0], ctx->ev[1]);
    st.ms_d2h += elapsed(ctx->ev[4], ctx->ev[5]);
    st.ms_total += elapsed(ctx->ev[0], ctx->ev[5]);
    st.total_launches = ctx->launches - launches0;
    if (stats_out) *stats_out = st;
    return (int)std::min<int64_t>(completed / M->num_problems, 0x7fffffff);
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

const char *qa_last_error(void) { return g_err.c_str(); }

int qa_version(void) { return QA_VERSION; }

int qa_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int qa_ctx_create(int device_id, qa_ctx **out) {
    if (!out) return fail(QA_ERR_ARG, "out is null");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(QA_ERR_CUDA, "no CUDA device available: libqanneal has no CPU fallback");
    }
    if (device_id < 0 || device_id >= ndev) return fail(QA_ERR_ARG, "device_id out of range");
    QA_CUDA(cudaSetDevice(device_id));
    qa_ctx *ctx = new qa_ctx();
    ctx->device = device_id;
    cudaDeviceProp prop;
    QA_CUDA(cudaGetDeviceProperties(&prop, device_id));
    ctx->num_sms = prop.multiProcessorCount;
    if (const char *e = getenv("QA_REPLAY_SWITCH_PERMILLE")) ctx->replay_switch_permille = atoi(e);  // development knobs
    if (const char *e = getenv("QA_REPLAY_WARPS")) ctx->replay_warps = atoi(e);
    QA_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    for (auto &ev : ctx->ev) QA_CUDA(cudaEventCreate(&ev));
    QA_CUDA(cudaMalloc((void **)&ctx->d_stats, (QA_NSTAT + 1 + QA_NDEBUG) * sizeof(unsigned long long)));
    QA_CUDA(cudaMalloc((void **)&ctx->d_flag, 2 * sizeof(int)));
    QA_CUDA(cudaMalloc((void **)&ctx->d_best_e, sizeof(double)));
    QA_CUDA(cudaMalloc((void **)&ctx->d_best_i, sizeof(long long)));
    QA_CUDA(cudaHostAlloc((void **)&ctx->h_iflag, sizeof(int), cudaHostAllocMapped));
    *ctx->h_iflag = 0;
    QA_CUDA(cudaHostGetDevicePointer((void **)&ctx->d_iflag, ctx->h_iflag, 0));
    *out = ctx;
    return QA_OK;
}

int qa_ctx_destroy(qa_ctx *ctx) {
    if (!ctx) return QA_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    release(ctx->f); release(ctx->spw); release(ctx->fT); release(ctx->sf); release(ctx->states); release(ctx->energies); release(ctx->seeds);
    release(ctx->betas); release(ctx->packed); release(ctx->misc); release(ctx->cubtmp);
    if (ctx->d_stats) cudaFree(ctx->d_stats);
    if (ctx->d_flag) cudaFree(ctx->d_flag);
    if (ctx->d_best_e) cudaFree(ctx->d_best_e);
    if (ctx->d_best_i) cudaFree(ctx->d_best_i);
    if (ctx->h_iflag) cudaFreeHost(ctx->h_iflag);
    for (auto &ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return QA_OK;
}

int qa_ctx_synchronize(qa_ctx *ctx) {
    if (!ctx) return fail(QA_ERR_ARG, "null context");
    QA_CUDA(cudaStreamSynchronize(ctx->stream));
    return QA_OK;
}

int qa_ctx_set_kernel(qa_ctx *ctx, int kernel) {
    if (!ctx) return fail(QA_ERR_ARG, "null context");
    if (kernel != QA_KERNEL_AUTO && kernel != QA_KERNEL_WARP_PER_READ && kernel != QA_KERNEL_LOCKSTEP_PUSH &&
        kernel != QA_KERNEL_REPLAY)
        return fail(QA_ERR_ARG, "kernel must be QA_KERNEL_AUTO, _WARP_PER_READ, _LOCKSTEP_PUSH or _REPLAY");
    ctx->kernel = kernel;
    return QA_OK;
}

int qa_ctx_last_kernel(qa_ctx *ctx) {
    if (!ctx) return fail(QA_ERR_ARG, "null context");
    return ctx->last_kernel;
}

int qa_ctx_resident_reads(qa_ctx *ctx) {
    if (!ctx) return fail(QA_ERR_ARG, "null context");
    int bps = 0;
    QA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_anneal_ref<false>, QA_TPB_MAX, 0));
    return bps * ctx->num_sms * (QA_TPB_MAX / 32);
}

int qa_model_from_ising(qa_ctx *ctx, int32_t n, const double *h, int64_t m, const int32_t *starts, const int32_t *ends,
                        const double *weights, qa_model **out) {
    if (!ctx || !out) return fail(QA_ERR_ARG, "null context or out");
    *out = nullptr;
    if (n < 0 || m < 0) return fail(QA_ERR_ARG, "negative size");
    if ((n > 0 && !h) || (m > 0 && (!starts || !ends || !weights))) return fail(QA_ERR_ARG, "null model vector");
    QA_CUDA(cudaSetDevice(ctx->device));
    const int64_t voff[2] = {0, n}, coff[2] = {0, m};
    return model_create(ctx, 1, voff, coff, h, starts, ends, weights, out);
}

int qa_model_set_groups(qa_model *M, int32_t ngroups, const int32_t *grp, const int32_t *coef, const double *lambda,
                        const int64_t *kappa) {
    if (!M) return fail(QA_ERR_ARG, "null model");
    if (M->num_problems != 1) return fail(QA_ERR_ARG, "groups need a single-problem model");
    if (ngroups < 0 || ngroups > QA_MAX_GROUPS) return fail(QA_ERR_LIMIT, "ngroups must be in [0, QA_MAX_GROUPS]");
    qa_ctx *ctx = M->ctx;
    QA_CUDA(cudaSetDevice(ctx->device));
    if (M->grp) { cudaFree(M->grp); M->grp = nullptr; }
    if (M->coef) { cudaFree(M->coef); M->coef = nullptr; }
    if (M->lambda) { cudaFree(M->lambda); M->lambda = nullptr; }
    if (M->kappa) { cudaFree(M->kappa); M->kappa = nullptr; }
    M->ngroups = ngroups;
    if (ngroups > 0) M->dn_ok = false;   // the dense form carries no group terms
    if (M->rp_slabs) { cudaFree(M->rp_slabs); M->rp_slabs = nullptr; }   // the slabs carry the group metadata
    if (M->rp_off) { cudaFree(M->rp_off); M->rp_off = nullptr; }
    M->rp_built = false;
    M->rp_ok = false;
    M->descs[0].rp_slabs = nullptr;
    M->descs[0].rp_off = nullptr;
    ProblemDesc &D = M->descs[0];
    D.ngroups = ngroups;
    D.grp = nullptr; D.coef = nullptr; D.lambda = nullptr; D.kappa = nullptr;
    if (ngroups == 0) return QA_OK;
    if (!grp || !coef || !lambda || !kappa) return fail(QA_ERR_ARG, "null group vector");
    const int32_t n = D.n;
    const int64_t npad = (int64_t)D.nch * 32;
    // host-side validation: exact integer arithmetic must stay below 2^53 (and (M+kappa)^2 below 2^62)
    std::vector<int32_t> hg(npad, -1), hc(npad, 0);
    std::vector<int32_t> tg(n), tc(n);
    QA_CUDA(cudaMemcpy(tg.data(), grp, (size_t)n * sizeof(int32_t), cudaMemcpyDefault));
    QA_CUDA(cudaMemcpy(tc.data(), coef, (size_t)n * sizeof(int32_t), cudaMemcpyDefault));
    std::vector<int64_t> hk(ngroups);
    QA_CUDA(cudaMemcpy(hk.data(), kappa, (size_t)ngroups * sizeof(int64_t), cudaMemcpyDefault));
    std::vector<double> sumabs(ngroups, 0.0), maxa(ngroups, 0.0);
    for (int v = 0; v < n; ++v) {
        if (tg[v] >= ngroups) return fail(QA_ERR_ARG, "group index out of range");
        hg[v] = tg[v] < 0 ? -1 : tg[v];
        hc[v] = tc[v];
        if (tg[v] >= 0) {
            sumabs[tg[v]] += std::fabs((double)tc[v]);
            maxa[tg[v]] = std::max(maxa[tg[v]], std::fabs((double)tc[v]));
        }
    }
    M->groups_i32 = true;
    for (int g = 0; g < ngroups; ++g) {
        const double span = sumabs[g] + std::fabs((double)hk[g]);
        if (span >= 2147483648.0 || maxa[g] * (maxa[g] + span) >= 9007199254740992.0)
            return fail(QA_ERR_LIMIT, "group coefficients too large for exact integer evaluation");
        if (maxa[g] * (maxa[g] + span) >= 2147483648.0) M->groups_i32 = false;
    }
    QA_CUDA(cudaMalloc((void **)&M->grp, npad * sizeof(int32_t)));
    QA_CUDA(cudaMalloc((void **)&M->coef, npad * sizeof(int32_t)));
    QA_CUDA(cudaMalloc((void **)&M->lambda, ngroups * sizeof(double)));
    QA_CUDA(cudaMalloc((void **)&M->kappa, ngroups * sizeof(long long)));
    QA_CUDA(cudaMemcpy(M->grp, hg.data(), npad * sizeof(int32_t), cudaMemcpyHostToDevice));
    QA_CUDA(cudaMemcpy(M->coef, hc.data(), npad * sizeof(int32_t), cudaMemcpyHostToDevice));
    QA_CUDA(cudaMemcpy(M->lambda, lambda, ngroups * sizeof(double), cudaMemcpyDefault));
    QA_CUDA(cudaMemcpy(M->kappa, hk.data(), ngroups * sizeof(long long), cudaMemcpyHostToDevice));
    D.grp = M->grp; D.coef = M->coef; D.lambda = M->lambda; D.kappa = M->kappa;
    return QA_OK;
}

// Dense k-way form for k_anneal_dense: W[i][j] = J between (i,c) and (j,c), P = J between two cases of one cell, derived
// from the device CSR and verified (every inter-cell coupler joins equal cases and does not depend on the case, every
// intra-cell coupler equals P).  Returns 1 when the model has that structure (QA_MODE_THROUGHPUT then runs the
// tensor-core kernel), 0 when it does not (nothing changes).
int qa_model_enable_dense(qa_model *M, int32_t K) {
    if (!M) return fail(QA_ERR_ARG, "null model");
    if (K != 1 && K != 2 && K != 4 && K != 8) return fail(QA_ERR_ARG, "dense form: cases per cell must be 1, 2, 4 or 8");
    if (M->num_problems != 1 || M->ngroups != 0) return fail(QA_ERR_ARG, "dense form needs a single problem without groups");
    const int64_t n = M->n_total;
    if (n == 0 || n % K != 0) return fail(QA_ERR_ARG, "number of variables is not a multiple of the cases per cell");
    qa_ctx *ctx = M->ctx;
    QA_CUDA(cudaSetDevice(ctx->device));
    if (M->dn_W) { cudaFree(M->dn_W); M->dn_W = nullptr; }
    M->dn_ok = false;
    const int32_t ncells = (int32_t)(n / K);
    const int32_t ncp = (ncells + 31) & ~31;
    // P: the first intra-cell coupler of variable 0 (its neighbours 1..K-1 come first in an ascending row; any order works)
    double Pj = 0.0;
    if (K > 1) {
        int32_t rp[2];
        QA_CUDA(cudaMemcpy(rp, M->rowptr, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost));
        const int32_t d0 = rp[1] - rp[0];
        if (d0 > 0) {
            std::vector<int32_t> c0(d0);
            std::vector<double> v0(d0);
            QA_CUDA(cudaMemcpy(c0.data(), M->col + rp[0], (size_t)d0 * sizeof(int32_t), cudaMemcpyDeviceToHost));
            QA_CUDA(cudaMemcpy(v0.data(), M->val + rp[0], (size_t)d0 * sizeof(double), cudaMemcpyDeviceToHost));
            for (int32_t e = 0; e < d0; ++e)
                if (c0[e] < K) { Pj = v0[e]; break; }
        }
    }
    const size_t cells2 = (size_t)ncp * ncp;
    QA_CUDA(cudaMalloc((void **)&M->dn_W, cells2 * sizeof(double)));
    unsigned long long *Wb = reinterpret_cast<unsigned long long *>(M->dn_W);
    unsigned long long *d_counts = nullptr;
    QA_CUDA(cudaMalloc((void **)&d_counts, 2 * sizeof(unsigned long long)));
    QA_CUDA(cudaMemsetAsync(d_counts, 0, 2 * sizeof(unsigned long long), ctx->stream));
    QA_CUDA(cudaMemsetAsync(ctx->d_flag, 0, 2 * sizeof(int), ctx->stream));
    const int tpb = 256;
    k_dense_fill<<<(unsigned)((cells2 + tpb - 1) / tpb), tpb, 0, ctx->stream>>>(cells2, Wb);
    for (int pass = 0; pass < 2; ++pass)
        k_dense_scatter<<<(unsigned)((n + tpb - 1) / tpb), tpb, 0, ctx->stream>>>((int32_t)n, K, ncp, M->rowptr, M->col, M->val, Pj, Wb,
                                                                                 ctx->d_flag, d_counts, pass);
    k_dense_finish<<<(unsigned)((cells2 + tpb - 1) / tpb), tpb, 0, ctx->stream>>>(cells2, Wb);
    ctx->launches += 4;
    int flag = 0;
    unsigned long long counts[2] = {0, 0};
    cudaError_t ce = cudaMemcpyAsync(&flag, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(counts, d_counts, sizeof(counts), cudaMemcpyDeviceToHost, ctx->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_counts);
    if (ce != cudaSuccess) return fail(QA_ERR_CUDA, std::string("dense form: ") + cudaGetErrorString(ce));
    if (flag != 0 || counts[0] * (unsigned long long)K != counts[1]) {
        cudaFree(M->dn_W);
        M->dn_W = nullptr;
        return 0;
    }
    M->dn.ncells = ncells;
    M->dn.ncp = ncp;
    M->dn.K = K;
    M->dn.ngrp = ncp / 32;
    M->dn.W = M->dn_W;
    M->dn.P = Pj;
    M->dn_ok = true;
    return 1;
}

int qa_model_num_variables(const qa_model *M) { return M ? (int)M->n_total : fail(QA_ERR_ARG, "null model"); }
int64_t qa_model_num_couplers(const qa_model *M) { return M ? M->m_total : (int64_t)fail(QA_ERR_ARG, "null model"); }
int qa_model_max_degree(const qa_model *M) { return M ? M->max_deg : fail(QA_ERR_ARG, "null model"); }

int qa_model_get_ising(const qa_model *M, double *h, int32_t *starts, int32_t *ends, double *weights) {
    if (!M) return fail(QA_ERR_ARG, "null model");
    QA_CUDA(cudaSetDevice(M->ctx->device));
    if (h) QA_CUDA(cudaMemcpy(h, M->h, (size_t)M->n_total * sizeof(double), cudaMemcpyDefault));
    if (starts) QA_CUDA(cudaMemcpy(starts, M->starts, (size_t)M->m_total * sizeof(int32_t), cudaMemcpyDefault));
    if (ends) QA_CUDA(cudaMemcpy(ends, M->ends, (size_t)M->m_total * sizeof(int32_t), cudaMemcpyDefault));
    if (weights) QA_CUDA(cudaMemcpy(weights, M->w, (size_t)M->m_total * sizeof(double), cudaMemcpyDefault));
    return QA_OK;
}

int qa_model_destroy(qa_model *M) {
    if (!M) return QA_OK;
    cudaSetDevice(M->ctx->device);
    cudaStreamSynchronize(M->ctx->stream);
    void *ptrs[] = {M->h, M->starts, M->ends, M->w, M->rowptr, M->col, M->val, M->grp, M->coef, M->lambda, M->kappa, M->d_descs,
                    M->bw_ptr, M->bw_words, M->ent_slot, M->rp_slabs, M->rp_off, M->dn_W};
    for (void *p : ptrs) if (p) cudaFree(p);
    delete M;
    return QA_OK;
}

int qa_sa_sample_model(qa_ctx *ctx, qa_model *model, int32_t num_reads, int8_t *states_inout, double *energies_out,
                       int32_t num_betas, const double *beta_schedule, int32_t sweeps_per_beta, const uint64_t *seeds,
                       int32_t seed_mode, int32_t mode, qa_interrupt_fn interrupt, void *interrupt_user, qa_stats *stats_out) {
    if (model && model->num_problems != 1) return fail(QA_ERR_ARG, "use qa_sa_sample_ising_batch for batched models");
    return sample_common(ctx, model, num_reads, states_inout, energies_out, num_betas, beta_schedule, sweeps_per_beta, seeds,
                         seed_mode, mode, interrupt, interrupt_user, stats_out);
}

int qa_sa_sample_ising(qa_ctx *ctx, int32_t n, const double *h, int64_t m, const int32_t *starts, const int32_t *ends,
                       const double *weights, int32_t num_reads, int8_t *states_inout, double *energies_out, int32_t num_betas,
                       const double *beta_schedule, int32_t sweeps_per_beta, const uint64_t *seeds, int32_t seed_mode,
                       int32_t mode, qa_stats *stats_out) {
    if (!ctx) return fail(QA_ERR_ARG, "null context");
    cudaEvent_t b0 = nullptr, b1 = nullptr;
    QA_CUDA(cudaSetDevice(ctx->device));
    QA_CUDA(cudaEventCreate(&b0));
    QA_CUDA(cudaEventCreate(&b1));
    const uint32_t l0 = ctx->launches;
    cudaEventRecord(b0, ctx->stream);
    qa_model *M = nullptr;
    int rc = qa_model_from_ising(ctx, n, h, m, starts, ends, weights, &M);
    cudaEventRecord(b1, ctx->stream);
    if (rc == QA_OK) {
        rc = sample_common(ctx, M, num_reads, states_inout, energies_out, num_betas, beta_schedule, sweeps_per_beta, seeds,
                           seed_mode, mode, nullptr, nullptr, stats_out);
        if (rc >= 0 && stats_out) {
            cudaEventSynchronize(b1);
            stats_out->ms_build = elapsed(b0, b1);
            stats_out->total_launches = ctx->launches - l0;
        }
    }
    qa_model_destroy(M);
    cudaEventDestroy(b0);
    cudaEventDestroy(b1);
    return rc;
}

int qa_sa_sample_ising_batch(qa_ctx *ctx, int32_t num_problems, const int64_t *var_offsets, const int64_t *coupler_offsets,
                             const double *h, const int32_t *starts, const int32_t *ends, const double *weights,
                             int32_t reads_per_problem, int8_t *states_inout, double *energies_out, int32_t num_betas,
                             const double *beta_schedule, int32_t sweeps_per_beta, const uint64_t *seeds, qa_stats *stats_out) {
    if (!ctx) return fail(QA_ERR_ARG, "null context");
    if (num_problems < 1 || !var_offsets || !coupler_offsets) return fail(QA_ERR_ARG, "bad batch description");
    for (int p = 0; p < num_problems; ++p)
        if (var_offsets[p + 1] < var_offsets[p] || coupler_offsets[p + 1] < coupler_offsets[p])
            return fail(QA_ERR_ARG, "offsets must be non-decreasing");
    if (var_offsets[0] != 0 || coupler_offsets[0] != 0) return fail(QA_ERR_ARG, "offsets must start at 0");
    QA_CUDA(cudaSetDevice(ctx->device));
    cudaEvent_t b0 = nullptr, b1 = nullptr;
    QA_CUDA(cudaEventCreate(&b0));
    QA_CUDA(cudaEventCreate(&b1));
    const uint32_t l0 = ctx->launches;
    cudaEventRecord(b0, ctx->stream);
    qa_model *M = nullptr;
    int rc = model_create(ctx, num_problems, var_offsets, coupler_offsets, h, starts, ends, weights, &M);
    cudaEventRecord(b1, ctx->stream);
    if (rc == QA_OK) {
        rc = sample_common(ctx, M, reads_per_problem, states_inout, energies_out, num_betas, beta_schedule, sweeps_per_beta,
                           seeds, QA_SEED_PER_READ, QA_MODE_REFERENCE, nullptr, nullptr, stats_out);
        if (rc >= 0 && stats_out) {
            cudaEventSynchronize(b1);
            stats_out->ms_build = elapsed(b0, b1);
            stats_out->total_launches = ctx->launches - l0;
        }
    }
    qa_model_destroy(M);
    cudaEventDestroy(b0);
    cudaEventDestroy(b1);
    return rc;
}

int qa_energy_argmin(qa_ctx *ctx, qa_model *M, int32_t num_reads, const int8_t *states, double *energies_out,
                     double *best_energy, int64_t *best_index, qa_stats *stats_out) {
    if (!ctx || !M) return fail(QA_ERR_ARG, "null context or model");
    if (M->num_problems != 1) return fail(QA_ERR_ARG, "single-problem model required");
    if (num_reads < 0) return fail(QA_ERR_ARG, "negative num_reads");
    qa_stats st;
    memset(&st, 0, sizeof(st));
    if (num_reads == 0) {
        if (best_energy) *best_energy = INFINITY;
        if (best_index) *best_index = -1;
        if (stats_out) *stats_out = st;
        return QA_OK;
    }
    if (!states) return fail(QA_ERR_ARG, "null states");
    QA_CUDA(cudaSetDevice(ctx->device));
    const uint32_t l0 = ctx->launches;
    ProblemDesc &D = M->descs[0];
    const int32_t rpad = (num_reads + 31) & ~31;
    const int64_t state_bytes = (int64_t)num_reads * D.n;
    QA_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
    const int8_t *d_states = states;
    if (!is_device_ptr(states)) {
        int rc = ensure(ctx->states, (size_t)std::max<int64_t>(state_bytes, 1));
        if (rc) return rc;
        QA_CUDA(cudaMemcpyAsync(ctx->states.p, states, (size_t)state_bytes, cudaMemcpyHostToDevice, ctx->stream));
        d_states = (const int8_t *)ctx->states.p;
    }
    double *d_energies = energies_out;
    const bool en_host = !energies_out || !is_device_ptr(energies_out);
    if (en_host) {
        int rc = ensure(ctx->energies, (size_t)num_reads * sizeof(double));
        if (rc) return rc;
        d_energies = (double *)ctx->energies.p;
    }
    int rc = ensure(ctx->packed, (size_t)std::max<int64_t>((int64_t)D.nch * rpad, 1) * sizeof(uint32_t));
    if (rc) return rc;
    D.reads = num_reads;
    D.rpad = rpad;
    D.read_base = 0;
    D.states = const_cast<int8_t *>(d_states);
    D.packedT = (uint32_t *)ctx->packed.p;
    D.energies = d_energies;
    QA_CUDA(cudaMemcpyAsync(M->d_descs, &D, sizeof(ProblemDesc), cudaMemcpyHostToDevice, ctx->stream));
    QA_CUDA(cudaMemsetAsync(ctx->d_flag, 0, 2 * sizeof(int), ctx->stream));
    QA_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
    {
        const int tpb = 256;
        const int64_t threads = (int64_t)num_reads * 32;
        k_pack_states<<<(unsigned)((threads + tpb - 1) / tpb), tpb, 0, ctx->stream>>>(D, ctx->d_flag);
        dim3 g((unsigned)((num_reads + 127) / 128), 1);
        k_energy<<<g, 128, 0, ctx->stream>>>(M->d_descs);
        k_argmin<<<1, 1024, 0, ctx->stream>>>(d_energies, num_reads, ctx->d_best_e, ctx->d_best_i);
        QA_CUDA(cudaGetLastError());
        ctx->launches += 3;
    }
    QA_CUDA(cudaEventRecord(ctx->ev[2], ctx->stream));
    int flag = 0;
    double be = 0;
    long long bi = 0;
    QA_CUDA(cudaMemcpyAsync(&flag, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    QA_CUDA(cudaMemcpyAsync(&be, ctx->d_best_e, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    QA_CUDA(cudaMemcpyAsync(&bi, ctx->d_best_i, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    if (energies_out && en_host)
        QA_CUDA(cudaMemcpyAsync(energies_out, d_energies, (size_t)num_reads * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    QA_CUDA(cudaEventRecord(ctx->ev[3], ctx->stream));
    QA_CUDA(cudaStreamSynchronize(ctx->stream));
    if (flag != 0) return fail(flag, "states must be +1/-1");
    if (best_energy) *best_energy = be;
    if (best_index) *best_index = bi;
    st.ms_h2d = elapsed(ctx->ev[0], ctx->ev[1]);
    st.ms_energy = elapsed(ctx->ev[1], ctx->ev[2]);
    st.ms_d2h = elapsed(ctx->ev[2], ctx->ev[3]);
    st.total_launches = ctx->launches - l0;
    if (stats_out) *stats_out = st;
    return QA_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// device model builders (host orchestration)
// ------------------------------------------------------------------------------------------------
namespace {

struct GraphDev {
    int32_t n = 0;
    int64_t m = 0;
    int32_t *eu = nullptr, *ev = nullptr;
    double *w = nullptr;
    int32_t *rowptr = nullptr;   // per-vertex edge lists, in G.edges order
    uint32_t *sorted = nullptr;  // entry ids 2e (u side) / 2e+1 (v side)
    void free_all() {
        void *ptrs[] = {eu, ev, w, rowptr, sorted};
        for (void *p : ptrs) if (p) cudaFree(p);
        eu = ev = nullptr; w = nullptr; rowptr = nullptr; sorted = nullptr;
    }
};

unsigned blocks_for(int64_t count, int tpb = 256) { return (unsigned)std::max<int64_t>(1, (count + tpb - 1) / tpb); }

int graph_to_device(qa_ctx *ctx, int32_t n, int64_t m, const int32_t *eu, const int32_t *ev, const double *w, GraphDev &g) {
    if (n < 0 || m < 0) return fail(QA_ERR_ARG, "negative graph size");
    if (m > 0 && (!eu || !ev || !w)) return fail(QA_ERR_ARG, "null edge list");
    if (2 * m >= (int64_t)0x7fffffff) return fail(QA_ERR_LIMIT, "too many edges");
    g.n = n;
    g.m = m;
    int rc = upload(ctx, &g.eu, eu, (size_t)m);
    if (!rc) rc = upload(ctx, &g.ev, ev, (size_t)m);
    if (!rc) rc = upload(ctx, &g.w, w, (size_t)m);
    if (rc) return rc;
    const int64_t entries = 2 * m;
    QA_CUDA(cudaMalloc((void **)&g.rowptr, (size_t)(n + 2) * sizeof(int32_t)));
    QA_CUDA(cudaMalloc((void **)&g.sorted, (size_t)std::max<int64_t>(entries, 1) * sizeof(uint32_t)));
    QA_CUDA(cudaMemsetAsync(ctx->d_flag, 0, 2 * sizeof(int), ctx->stream));
    rc = ensure(ctx->misc, (size_t)std::max<int64_t>(entries, 1) * 3 * sizeof(uint32_t));
    if (rc) return rc;
    uint32_t *keys = (uint32_t *)ctx->misc.p, *vals = keys + std::max<int64_t>(entries, 1), *keys2 = vals + std::max<int64_t>(entries, 1);
    if (entries > 0) {
        k_graph_entries<<<blocks_for(m), 256, 0, ctx->stream>>>(m, n, g.eu, g.ev, keys, vals, ctx->d_flag);
        int end_bit = 1;
        while (((int64_t)1 << end_bit) < (int64_t)n + 1 && end_bit < 32) ++end_bit;
        size_t tmp = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp, keys, keys2, vals, g.sorted, (int)entries, 0, end_bit, ctx->stream);
        rc = ensure(ctx->cubtmp, tmp);
        if (rc) return rc;
        cudaError_t ce = cub::DeviceRadixSort::SortPairs(ctx->cubtmp.p, tmp, keys, keys2, vals, g.sorted, (int)entries, 0, end_bit, ctx->stream);
        if (ce != cudaSuccess) return fail(QA_ERR_CUDA, std::string("radix sort: ") + cudaGetErrorString(ce));
        ctx->launches += 5;
    }
    k_rowptr<<<blocks_for(n + 1), 256, 0, ctx->stream>>>(n + 1, entries, keys2, g.rowptr, ctx->d_flag + 1);
    ctx->launches++;
    int flag = 0;
    QA_CUDA(cudaMemcpyAsync(&flag, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    QA_CUDA(cudaStreamSynchronize(ctx->stream));
    if (flag) return fail(QA_ERR_INDEX, "edge endpoint out of range or self-loop");
    return QA_OK;
}

int device_seq_sum(qa_ctx *ctx, int64_t count, const double *x, double scale, double *out_host) {
    k_seq_sum<<<1, 32, 0, ctx->stream>>>(count, x, scale, ctx->d_best_e);
    ctx->launches++;
    QA_CUDA(cudaMemcpyAsync(out_host, ctx->d_best_e, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    QA_CUDA(cudaStreamSynchronize(ctx->stream));
    return QA_OK;
}

// QUBO (lin[nvar], couplers as (hi<<32|lo) keys with values q) -> resident Ising model in dimod's vector order.
// offset_out = sum_i lin_i/2 + sum_c q_c/4 in python's left-to-right order (bqm.qubo_to_ising_vectors).
int lower_qubo(qa_ctx *ctx, int32_t nvar, const double *d_lin, int64_t mq, unsigned long long *d_keys, const double *d_q,
               qa_model **out, double *offset_out) {
    if (mq >= (int64_t)0x3fffffff) return fail(QA_ERR_LIMIT, "too many couplers");
    unsigned long long *keys2 = nullptr;
    uint32_t *perm = nullptr, *perm2 = nullptr;
    int32_t *r = nullptr, *c = nullptr;
    double *J = nullptr;
    const size_t mm = (size_t)std::max<int64_t>(mq, 1);
    int rc = QA_OK;
    auto cleanup = [&]() {
        void *ptrs[] = {keys2, perm, perm2, r, c, J};
        for (void *p : ptrs) if (p) cudaFree(p);
    };
    do {
        if (cudaMalloc((void **)&keys2, mm * 8) != cudaSuccess || cudaMalloc((void **)&perm, mm * 4) != cudaSuccess ||
            cudaMalloc((void **)&perm2, mm * 4) != cudaSuccess || cudaMalloc((void **)&r, mm * 4) != cudaSuccess ||
            cudaMalloc((void **)&c, mm * 4) != cudaSuccess || cudaMalloc((void **)&J, mm * 8) != cudaSuccess) {
            rc = fail(QA_ERR_CUDA, "out of device memory in model builder");
            break;
        }
        if (mq > 0) {
            // identity permutation, then sort by (row = larger index, col = smaller index): dimod to_numpy_vectors order
            std::vector<uint32_t> iota(mq);
            for (int64_t i = 0; i < mq; ++i) iota[i] = (uint32_t)i;
            cudaMemcpyAsync(perm, iota.data(), mq * 4, cudaMemcpyHostToDevice, ctx->stream);
            size_t tmp = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, tmp, d_keys, keys2, perm, perm2, (int)mq, 0, 64, ctx->stream);
            rc = ensure(ctx->cubtmp, tmp);
            if (rc) break;
            cudaError_t ce = cub::DeviceRadixSort::SortPairs(ctx->cubtmp.p, tmp, d_keys, keys2, perm, perm2, (int)mq, 0, 64, ctx->stream);
            if (ce != cudaSuccess) { rc = fail(QA_ERR_CUDA, std::string("radix sort: ") + cudaGetErrorString(ce)); break; }
            k_split_keys<<<blocks_for(mq), 256, 0, ctx->stream>>>(mq, keys2, perm2, d_q, r, c, J);
            cudaStreamSynchronize(ctx->stream);  // iota must outlive the copy
            ctx->launches += 9;
        }
        const int64_t voff[2] = {0, nvar}, coff[2] = {0, mq};
        qa_model *M = nullptr;
        rc = model_create(ctx, 1, voff, coff, d_lin, r, c, J, &M);
        if (rc) break;
        k_h_from_rows<<<blocks_for(nvar), 256, 0, ctx->stream>>>(nvar, d_lin, M->rowptr, M->val, M->h);
        ctx->launches++;
        double s1 = 0.0, s2 = 0.0;
        rc = device_seq_sum(ctx, nvar, d_lin, 0.5, &s1);
        if (!rc) rc = device_seq_sum(ctx, mq, J, 1.0, &s2);
        if (rc) { qa_model_destroy(M); break; }
        *offset_out = (0.0 + s1) + s2;
        *out = M;
    } while (0);
    cleanup();
    return rc;
}

std::vector<int> slack_coefficients_host(int64_t upper) {
    // binary-encoded slack spanning exactly 0..upper (dimod add_linear_inequality_constraint; models.slack_coefficients)
    std::vector<int> c;
    if (upper <= 0) return c;
    int nbits = 0;
    while (((int64_t)2 << nbits) <= upper) ++nbits;  // floor(log2(upper))
    for (int j = 0; j < nbits; ++j) c.push_back(1 << j);
    if (upper - ((int64_t)1 << nbits) >= 0) c.push_back((int)(upper - ((int64_t)1 << nbits) + 1));
    return c;
}

}  // namespace

extern "C" {

int qa_build_cut_balance(qa_ctx *ctx, int32_t n, int64_t m, const int32_t *eu, const int32_t *ev, const double *w,
                         double gamma_factor, double k, qa_model **out, double *offset_out, double *gamma_out) {
    if (!ctx || !out || !offset_out) return fail(QA_ERR_ARG, "null argument");
    *out = nullptr;
    QA_CUDA(cudaSetDevice(ctx->device));
    GraphDev g;
    double *lin = nullptr, *q = nullptr;
    unsigned long long *keys = nullptr;
    int rc = graph_to_device(ctx, n, m, eu, ev, w, g);
    do {
        if (rc) break;
        const size_t mm = (size_t)std::max<int64_t>(m, 1);
        if (cudaMalloc((void **)&lin, (size_t)std::max(n, 1) * 8) != cudaSuccess || cudaMalloc((void **)&q, mm * 8) != cudaSuccess ||
            cudaMalloc((void **)&keys, mm * 8) != cudaSuccess) { rc = fail(QA_ERR_CUDA, "out of device memory"); break; }
        // W = G.size(weight) = (sum of weighted degrees) / 2, degrees summed in adjacency (edge) order
        double sumdeg = 0.0;
        k_vertex_accumulate<<<blocks_for(n), 256, 0, ctx->stream>>>(n, g.rowptr, g.sorted, g.w, 0, 1.0, 0.0, lin);
        rc = device_seq_sum(ctx, n, lin, 1.0, &sumdeg);
        if (rc) break;
        const double W = sumdeg / 2;
        const double gamma = gamma_factor * W / n;
        k_vertex_accumulate<<<blocks_for(n), 256, 0, ctx->stream>>>(n, g.rowptr, g.sorted, g.w, 0, k, 0.0, lin);
        k_edge_couplers<<<blocks_for(m), 256, 0, ctx->stream>>>(m, g.eu, g.ev, g.w, 0, k * -2, keys, q);
        ctx->launches += 3;
        double off = 0.0;
        qa_model *M = nullptr;
        rc = lower_qubo(ctx, n, lin, m, keys, q, &M, &off);
        if (rc) break;
        std::vector<int32_t> grp(n, 0), coef(n, 1);
        const double lam = gamma;
        const int64_t kap = 0;
        rc = qa_model_set_groups(M, 1, grp.data(), coef.data(), &lam, &kap);
        if (rc) { qa_model_destroy(M); break; }
        *offset_out = off - gamma * n * n / 4.0;
        if (gamma_out) *gamma_out = gamma;
        *out = M;
    } while (0);
    g.free_all();
    if (lin) cudaFree(lin);
    if (q) cudaFree(q);
    if (keys) cudaFree(keys);
    return rc;
}

int qa_build_subsampling(qa_ctx *ctx, int32_t n, int64_t m, const int32_t *eu, const int32_t *ev, const double *w, double gamma,
                         double P, qa_model **out, double *offset_out) {
    if (!ctx || !out || !offset_out) return fail(QA_ERR_ARG, "null argument");
    *out = nullptr;
    QA_CUDA(cudaSetDevice(ctx->device));
    GraphDev g;
    double *lin = nullptr, *q = nullptr;
    unsigned long long *keys = nullptr;
    int rc = graph_to_device(ctx, n, m, eu, ev, w, g);
    do {
        if (rc) break;
        const size_t mm = (size_t)std::max<int64_t>(m, 1);
        if (cudaMalloc((void **)&lin, (size_t)std::max(n, 1) * 8) != cudaSuccess || cudaMalloc((void **)&q, mm * 8) != cudaSuccess ||
            cudaMalloc((void **)&keys, mm * 8) != cudaSuccess) { rc = fail(QA_ERR_CUDA, "out of device memory"); break; }
        k_vertex_accumulate<<<blocks_for(n), 256, 0, ctx->stream>>>(n, g.rowptr, g.sorted, g.w, 1, -P, gamma, lin);
        k_edge_couplers<<<blocks_for(m), 256, 0, ctx->stream>>>(m, g.eu, g.ev, g.w, 1, P, keys, q);
        ctx->launches += 2;
        rc = lower_qubo(ctx, n, lin, m, keys, q, out, offset_out);
    } while (0);
    g.free_all();
    if (lin) cudaFree(lin);
    if (q) cudaFree(q);
    if (keys) cudaFree(keys);
    return rc;
}

// shared by the DQM and CQM builders: K copies of the graph + one-hot pairs, then model-specific groups
static int build_kway(qa_ctx *ctx, int32_t n, int64_t m, const int32_t *eu, const int32_t *ev, const double *w, int32_t K,
                      int lin_mode, double lin_scale, double lin_base, double lin_shift, double edge_pre, double edge_shift,
                      double A, int32_t extra_vars, qa_model **out, double *off) {
    GraphDev g;
    double *cell = nullptr, *lin = nullptr, *q = nullptr;
    unsigned long long *keys = nullptr;
    int rc = graph_to_device(ctx, n, m, eu, ev, w, g);
    do {
        if (rc) break;
        const int64_t nvar64 = (int64_t)n * K + extra_vars;
        if (nvar64 >= (int64_t)0x7fffffff) { rc = fail(QA_ERR_LIMIT, "too many variables"); break; }
        const int32_t nvar = (int32_t)nvar64;
        const int64_t mq = m * K + (int64_t)n * ((int64_t)K * (K - 1) / 2);
        const size_t mm = (size_t)std::max<int64_t>(mq, 1);
        if (cudaMalloc((void **)&cell, (size_t)std::max(n, 1) * 8) != cudaSuccess ||
            cudaMalloc((void **)&lin, (size_t)std::max(nvar, 1) * 8) != cudaSuccess || cudaMalloc((void **)&q, mm * 8) != cudaSuccess ||
            cudaMalloc((void **)&keys, mm * 8) != cudaSuccess) { rc = fail(QA_ERR_CUDA, "out of device memory"); break; }
        k_vertex_accumulate<<<blocks_for(n), 256, 0, ctx->stream>>>(n, g.rowptr, g.sorted, g.w, lin_mode, lin_scale, lin_base, cell);
        k_kway_linear<<<blocks_for(nvar), 256, 0, ctx->stream>>>(n, K, nvar, cell, lin_shift, lin);
        k_kway_couplers<<<blocks_for(mq), 256, 0, ctx->stream>>>(m, n, K, g.eu, g.ev, g.w, -2.0, edge_pre, edge_shift, 2.0 * A, keys, q);
        ctx->launches += 3;
        rc = lower_qubo(ctx, nvar, lin, mq, keys, q, out, off);
    } while (0);
    g.free_all();
    void *ptrs[] = {cell, lin, q, keys};
    for (void *p : ptrs) if (p) cudaFree(p);
    return rc;
}

int qa_build_cqm_penalty(qa_ctx *ctx, int32_t n, int64_t m, const int32_t *eu, const int32_t *ev, const double *w, int32_t K,
                         int32_t min_size, double onehot_penalty, double size_penalty, qa_model **out, double *offset_out) {
    if (!ctx || !out || !offset_out) return fail(QA_ERR_ARG, "null argument");
    if (K < 1 || K > QA_MAX_GROUPS) return fail(QA_ERR_LIMIT, "number of clusters must be in [1, QA_MAX_GROUPS]");
    *out = nullptr;
    QA_CUDA(cudaSetDevice(ctx->device));
    const std::vector<int> coeffs = slack_coefficients_host((int64_t)n - min_size);
    const int nb = (int)coeffs.size();
    double off = 0.0;
    qa_model *M = nullptr;
    // linear = (number of edges at the cell) - A  (CQM_clustering.py:43: coefficient 1 per edge end, not w)
    int rc = build_kway(ctx, n, m, eu, ev, w, K, 2, 0.0, 0.0, -onehot_penalty, 0.0, 0.0, onehot_penalty, K * nb, &M, &off);
    if (rc) return rc;
    const int64_t nx = (int64_t)n * K;
    std::vector<int32_t> grp(nx + (int64_t)K * nb), coef(nx + (int64_t)K * nb);
    for (int64_t v = 0; v < nx; ++v) { grp[v] = (int32_t)(v % K); coef[v] = 1; }
    long long csum = 0;
    for (int c : coeffs) csum += c;
    for (int j = 0; j < K; ++j)
        for (int b = 0; b < nb; ++b) { grp[nx + (int64_t)j * nb + b] = j; coef[nx + (int64_t)j * nb + b] = -coeffs[b]; }
    std::vector<double> lam(K, size_penalty);
    std::vector<int64_t> kap(K, (int64_t)n - csum - 2 * (int64_t)min_size);
    rc = qa_model_set_groups(M, K, grp.data(), coef.data(), lam.data(), kap.data());
    if (rc) { qa_model_destroy(M); return rc; }
    *offset_out = off + onehot_penalty * n;
    *out = M;
    return QA_OK;
}

int qa_build_dqm_onehot(qa_ctx *ctx, int32_t n, int64_t m, const int32_t *eu, const int32_t *ev, const double *w, int32_t K,
                        double gamma, double penalty, int32_t intended, qa_model **out, double *offset_out) {
    if (!ctx || !out || !offset_out) return fail(QA_ERR_ARG, "null argument");
    if (K < 1 || K > QA_MAX_GROUPS) return fail(QA_ERR_LIMIT, "number of cases must be in [1, QA_MAX_GROUPS]");
    *out = nullptr;
    QA_CUDA(cudaSetDevice(ctx->device));
    const double base_lin = gamma * (1 - (double)n / K);
    double off = 0.0;
    qa_model *M = nullptr;
    int rc;
    if (!intended)  // as written: last set_linear wins; edges overwrite the all-pairs 2*gamma which stays in the rank-1 group
        rc = build_kway(ctx, n, m, eu, ev, w, K, 3, 0.0, base_lin, -penalty, 0.0, -(2 * gamma), penalty, 0, &M, &off);
    else
        rc = build_kway(ctx, n, m, eu, ev, w, K, 0, 1.0, base_lin, -penalty, 2 * gamma, -(2 * gamma), penalty, 0, &M, &off);
    if (rc) return rc;
    const int64_t nx = (int64_t)n * K;
    std::vector<int32_t> grp(nx), coef(nx, 1);
    for (int64_t v = 0; v < nx; ++v) grp[v] = (int32_t)(v % K);
    std::vector<double> lam(K, gamma);
    std::vector<int64_t> kap(K, (int64_t)n - 1);
    rc = qa_model_set_groups(M, K, grp.data(), coef.data(), lam.data(), kap.data());
    if (rc) { qa_model_destroy(M); return rc; }
    *offset_out = off + penalty * n - gamma * K / 4.0;
    *out = M;
    return QA_OK;
}

}  // extern "C"

#include "postprocess.cuh"
#include "snn.cuh"
#include "recursion.cuh"
