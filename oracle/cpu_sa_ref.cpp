// oracle/cpu_sa_ref.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the simulated-annealing hot loop that the reference's
// clustering functions reach through the dimod Sampler API when they run
// offline (BASELINE.json: dwave-neal's SimulatedAnnealingSampler):
//
//   reference call sites   Python_Functions/BQM_clustering.py:57,75,85,245,263,273,386
//                          Python_Functions/QA_subsampling.py:42,56,65
//                          Python_Functions/DQM_clustering.py:45, CQM_clustering.py:53,89
//   third-party algorithm  dwave-neal 0.5.x  neal/src/cpu_sa.cpp
//                          (same loop: dwave-samplers >= 1.0  dwave/samplers/sa/src/cpu_sa.cpp)
//
// The arithmetic lives in that un-vendored dependency (requirements.txt:1,
// "dwave-ocean-sdk>=3.3.0", a floor not a pin).  Neither dimod nor neal is
// installed in this image and there is no network, and the reference ships no
// tests, so:            *** PARITY UNPINNED ***
// This file restates the PUBLISHED algorithm from the upstream description in
// SURVEY.md (rows a8-a11, Appendix C); it is validated by brute force and by
// known-answer energies (tests/), not against the real library.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load the library built from this file.
//
// Functions (all restate upstream neal/src/cpu_sa.cpp by name):
//   xorshift128+ step            <- FASTRAND macro
//   flip_energy_field()          <- get_flip_energy()
//   anneal_one_read()            <- simulated_annealing_run()
//   state_energy()               <- get_state_energy()
//   oracle_sa_sample_ising()     <- general_simulated_annealing()
//
// Extension kept bit-compatible with the CUDA library (csrc/qanneal.cu): an
// optional "rank-1 group" term  sum_g lambda_g/4 * (sum_{v in g} a_v s_v + kappa_g)^2
// evaluated lazily from per-read integer counters (SURVEY.md Appendix A:
// cut+balance, DQM cluster-size, CQM size penalties).  With ngroups == 0 the
// code path is exactly the upstream loop.
//
// Build: see oracle/Makefile  (g++ -O2 -ffp-contract=off, no -ffast-math).

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <string>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

thread_local std::string g_err;

struct Rng {               // upstream: static uint64_t rng_state[2]
    uint64_t s0, s1;
};

inline void rng_seed(Rng &g, uint64_t seed) {
    // upstream general_simulated_annealing(): rng_state[0] = seed ? seed : RANDMAX; rng_state[1] = 0
    g.s0 = seed ? seed : ~uint64_t(0);
    g.s1 = 0;
}

inline uint64_t rng_next(Rng &g) {  // upstream FASTRAND
    uint64_t x = g.s0;
    const uint64_t y = g.s1;
    g.s0 = y;
    x ^= x << 23;
    g.s1 = x ^ y ^ (x >> 17) ^ (y >> 26);
    return g.s1 + y;
}

struct Adjacency {
    std::vector<int64_t> rowptr;   // n+1
    std::vector<int32_t> col;      // 2m, per-row order == coupler order (upstream push_back order)
    std::vector<double>  val;
};

struct Groups {
    int32_t ngroups = 0;
    const int32_t *grp = nullptr;    // [n]  group of variable or -1
    const int32_t *coef = nullptr;   // [n]  integer coefficient a_v
    const double  *lambda = nullptr; // [ngroups]
    const int64_t *kappa = nullptr;  // [ngroups]
};

struct Counters {
    uint64_t attempts = 0, candidates = 0, draws = 0, accepted = 0, nbr_updates = 0;
};

// lazily evaluated rank-1 flip cost: lambda*a*(a - s*(M+kappa)); integers exact, one rounding
inline double group_flip_cost(const Groups &G, int v, int s, const int64_t *M) {
    const int g = G.grp[v];
    if (g < 0) return 0.0;
    const int64_t a = G.coef[v];
    const int64_t t = a * (a - int64_t(s) * (M[g] + G.kappa[g]));
    return G.lambda[g] * double(t);
}

// upstream get_flip_energy(): energy = h[var]; for n_i: energy += state[nbr]*coupling; return -2*state[var]*energy
inline double flip_energy_field(int v, const int8_t *state, const double *h, const Adjacency &A) {
    double energy = h[v];
    for (int64_t e = A.rowptr[v]; e < A.rowptr[v + 1]; ++e)
        energy += state[A.col[e]] * A.val[e];
    return energy;
}

// upstream simulated_annealing_run()
void anneal_one_read(int8_t *state, int n, const double *h, const Adjacency &A,
                     int sweeps_per_beta, const double *betas, int num_betas,
                     Rng &rng, const Groups &G, int64_t *M, double *dE, Counters &C) {
    for (int v = 0; v < n; ++v)
        dE[v] = -2 * state[v] * flip_energy_field(v, state, h, A);
    if (G.ngroups) {
        for (int g = 0; g < G.ngroups; ++g) M[g] = 0;
        for (int v = 0; v < n; ++v)
            if (G.grp[v] >= 0) M[G.grp[v]] += int64_t(G.coef[v]) * state[v];
    }
    for (int b = 0; b < num_betas; ++b) {
        const double beta = betas[b];
        for (int sweep = 0; sweep < sweeps_per_beta; ++sweep) {
            const double threshold = 44.36142 / beta;
            for (int v = 0; v < n; ++v) {
                double d = dE[v];
                if (G.ngroups) d = d + group_flip_cost(G, v, state[v], M);
                C.attempts++;
                if (d >= threshold) continue;
                C.candidates++;
                bool flip = false;
                if (d <= 0.0) {
                    flip = true;
                } else {
                    const uint64_t r = rng_next(rng);
                    C.draws++;
                    // upstream: exp(-delta_energy[var]*beta) * RANDMAX > rand, RANDMAX=(uint64_t)-1 -> 2^64 as double
                    if (std::exp(-d * beta) * 18446744073709551616.0 > double(r)) flip = true;
                }
                if (flip) {
                    C.accepted++;
                    const int multiplier = 4 * state[v];
                    for (int64_t e = A.rowptr[v]; e < A.rowptr[v + 1]; ++e) {
                        const int j = A.col[e];
                        dE[j] += multiplier * A.val[e] * state[j];
                    }
                    C.nbr_updates += uint64_t(A.rowptr[v + 1] - A.rowptr[v]);
                    if (G.ngroups && G.grp[v] >= 0)
                        M[G.grp[v]] -= 2 * int64_t(G.coef[v]) * state[v];
                    state[v] = int8_t(-state[v]);
                    dE[v] = -dE[v];
                }
            }
        }
    }
}

// upstream get_state_energy(): linear terms in variable order, then couplers in coupler order
double state_energy(const int8_t *state, int n, const double *h, int64_t m,
                    const int32_t *starts, const int32_t *ends, const double *w, const Groups &G) {
    double energy = 0.0;
    for (int v = 0; v < n; ++v) energy += state[v] * h[v];
    for (int64_t c = 0; c < m; ++c) energy += state[starts[c]] * w[c] * state[ends[c]];
    if (G.ngroups) {
        std::vector<int64_t> M(G.ngroups, 0);
        for (int v = 0; v < n; ++v)
            if (G.grp[v] >= 0) M[G.grp[v]] += int64_t(G.coef[v]) * state[v];
        for (int g = 0; g < G.ngroups; ++g) {
            const int64_t t = M[g] + G.kappa[g];
            energy += G.lambda[g] * double(t * t) * 0.25;
        }
    }
    return energy;
}

int build_adjacency(int n, int64_t m, const int32_t *starts, const int32_t *ends, const double *w, Adjacency &A) {
    // upstream general_simulated_annealing(): for each coupler push_back on both endpoints, in coupler order
    std::vector<int64_t> deg(n, 0);
    for (int64_t c = 0; c < m; ++c) {
        const int u = starts[c], v = ends[c];
        if (u < 0 || v < 0 || u >= n || v >= n) { g_err = "coupler indices out of range"; return -2; }
        if (u == v) { g_err = "self-loop coupler"; return -2; }
        deg[u]++; deg[v]++;
    }
    A.rowptr.assign(n + 1, 0);
    for (int v = 0; v < n; ++v) A.rowptr[v + 1] = A.rowptr[v] + deg[v];
    A.col.resize(2 * m); A.val.resize(2 * m);
    std::vector<int64_t> fill(A.rowptr.begin(), A.rowptr.end() - 1);
    for (int64_t c = 0; c < m; ++c) {
        const int u = starts[c], v = ends[c];
        A.col[fill[u]] = v; A.val[fill[u]++] = w[c];
        A.col[fill[v]] = u; A.val[fill[v]++] = w[c];
    }
    return 0;
}

}  // namespace

extern "C" {

struct oracle_stats {
    uint64_t attempts, candidates, draws, accepted, nbr_updates;
};

const char *oracle_last_error(void) { return g_err.c_str(); }

// seed_mode 0: per-read (read r == upstream call with num_samples=1, seed=seeds[r], states=init[r])
// seed_mode 1: stream   (== upstream call with num_samples=R, seed=seeds[0]; one RNG stream across reads)
// Returns number of reads completed (== num_reads) or a negative error code.
int oracle_sa_sample_ising(int32_t n, const double *h, int64_t m, const int32_t *starts,
                           const int32_t *ends, const double *weights, int32_t num_reads,
                           int8_t *states_inout, double *energies_out, int32_t num_betas,
                           const double *beta_schedule, int32_t sweeps_per_beta,
                           const uint64_t *seeds, int32_t seed_mode, int32_t ngroups,
                           const int32_t *grp, const int32_t *coef, const double *lambda,
                           const int64_t *kappa, int32_t nthreads, oracle_stats *stats_out) {
    if (n < 0 || m < 0 || num_reads < 0 || num_betas < 0 || sweeps_per_beta < 0) { g_err = "negative size"; return -1; }
    Adjacency A;
    int rc = build_adjacency(n, m, starts, ends, weights, A);
    if (rc) return rc;
    Groups G;
    G.ngroups = ngroups; G.grp = grp; G.coef = coef; G.lambda = lambda; G.kappa = kappa;
    for (int64_t i = 0; i < int64_t(num_reads) * n; ++i)
        if (states_inout[i] != 1 && states_inout[i] != -1) { g_err = "states must be +-1"; return -3; }
    Counters total;
    if (seed_mode == 1) {
        Rng rng; rng_seed(rng, seeds[0]);
        std::vector<double> dE(n);
        std::vector<int64_t> M(ngroups > 0 ? ngroups : 1);
        for (int r = 0; r < num_reads; ++r) {
            int8_t *st = states_inout + int64_t(r) * n;
            anneal_one_read(st, n, h, A, sweeps_per_beta, beta_schedule, num_betas, rng, G, M.data(), dE.data(), total);
            energies_out[r] = state_energy(st, n, h, m, starts, ends, weights, G);
        }
    } else {
#ifdef _OPENMP
        if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
        #pragma omp parallel
        {
            std::vector<double> dE(n);
            std::vector<int64_t> M(ngroups > 0 ? ngroups : 1);
            Counters loc;
            #pragma omp for schedule(dynamic, 1)
            for (int r = 0; r < num_reads; ++r) {
                Rng rng; rng_seed(rng, seeds[r]);
                int8_t *st = states_inout + int64_t(r) * n;
                anneal_one_read(st, n, h, A, sweeps_per_beta, beta_schedule, num_betas, rng, G, M.data(), dE.data(), loc);
                energies_out[r] = state_energy(st, n, h, m, starts, ends, weights, G);
            }
            #pragma omp critical
            {
                total.attempts += loc.attempts; total.candidates += loc.candidates; total.draws += loc.draws;
                total.accepted += loc.accepted; total.nbr_updates += loc.nbr_updates;
            }
        }
    }
    if (stats_out) {
        stats_out->attempts = total.attempts; stats_out->candidates = total.candidates;
        stats_out->draws = total.draws; stats_out->accepted = total.accepted;
        stats_out->nbr_updates = total.nbr_updates;
    }
    return num_reads;
}

// energies only (upstream get_state_energy per row), for checking dimod-style bqm.energies
int oracle_state_energies(int32_t n, const double *h, int64_t m, const int32_t *starts, const int32_t *ends,
                          const double *weights, int32_t num_reads, const int8_t *states, double *energies_out,
                          int32_t ngroups, const int32_t *grp, const int32_t *coef, const double *lambda,
                          const int64_t *kappa) {
    Groups G;
    G.ngroups = ngroups; G.grp = grp; G.coef = coef; G.lambda = lambda; G.kappa = kappa;
    for (int64_t c = 0; c < m; ++c)
        if (starts[c] < 0 || ends[c] < 0 || starts[c] >= n || ends[c] >= n) { g_err = "coupler indices out of range"; return -2; }
    for (int r = 0; r < num_reads; ++r)
        energies_out[r] = state_energy(states + int64_t(r) * n, n, h, m, starts, ends, weights, G);
    return 0;
}

// first `count` outputs of the xorshift128+ stream (known-answer checks for the RNG)
void oracle_rng_stream(uint64_t seed, int32_t count, uint64_t *out) {
    Rng g; rng_seed(g, seed);
    for (int i = 0; i < count; ++i) out[i] = rng_next(g);
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
