// replay.cuh -- k_anneal_replay: bit-exact reference-order annealing with DEFERRED neighbour updates.
// Included by qanneal.cu (needs ProblemDesc, AnnealParams, LaneStats, ls_accept, red_add_f64_if).
//
// neal (cpu_sa.cpp simulated_annealing_run) pushes every accepted flip of u into the flip energies of
// all neighbours j at once.  On a GPU with 32 reads per warp that is a random 8-byte read-modify-write per (flip, neighbour)
// on a field matrix far larger than L2; measured on B200 (tools/ubench_update.cu) the memory system tops out at
// 1.5-2.8e10 attempts/s for every acceptance rate >= 5 % -- the lockstep push kernel already sits at that ceiling.
//
// This kernel keeps the SAME arithmetic but changes when it happens.  Per read it stores
//     f[v]  the local field h_v + sum_j J_vj s_j as of the LAST VISIT of v      (fp64, read-interleaved f[v][lane])
//     S[v]  the spin, F[v] "v flipped at its last visit"                           (bit-packed, {S,F}[word][lane])
// Between two visits of v every neighbour u is visited exactly once, so the updates neal would have pushed into f[v] are
//     + 2 s_u J_uv   for the flagged neighbours u > v (flipped later in the previous sweep), ascending u, then
//     + 2 s_u J_uv   for the flagged neighbours u < v (flipped earlier in this sweep),       ascending u
// -- the same fp64 additions in the same order (s_u is the spin after the flip; a neighbour cannot flip twice in that
// window).  Replaying them at the visit gives bit-identical fields, decisions, RNG draws and final states, but the memory
// traffic becomes sequential and fully coalesced: 8 B of field per attempt (+8 B when it changed) plus the {S,F} words of
// the neighbour cells, instead of 16 B x 4 (sector granularity) per (flip, neighbour).
// When flips become rare the cost balance inverts (a replay touches every neighbour of every variable, a push only the
// neighbours of accepted flips), so once the CTA-wide acceptance of a sweep drops below P.switch_permille the tile runs one
// catch-up pass (pending u > v updates only) and finishes the schedule pushing, as neal does.  Both forms are exact, so
// the hand-over point does not influence the result.
//
// Coupling slabs: the host packs consecutive variables greedily into blocks (at most RP_D variables, RP_MAXBW foreign spin
// words, RP_CAP entries, never across a 32-variable spin word) and builds one contiguous slab {RpHdr, RpEntry[]} per block
// (replay order, 2J premultiplied, slot/bit of the neighbour's spin word).  All warps of a CTA walk the blocks together; slabs are brought
// into a 4-stage shared-memory ring with cp.async.bulk (TMA 1-D) completing on mbarriers, and released by one mbarrier arrive
// per warp -- no __syncthreads in the sweep.  There is no fixed producer: whichever warp first gets within RP_DIST blocks of a
// slab that has not been requested yet claims it (one shared-memory CAS) and issues the copy, so the ring runs at the pace
// of the fastest warp and nobody waits for a particular (possibly de-prioritised) warp.
#pragma once

#ifndef RP_D
#define RP_D 16       // variables per block, at most
#endif
constexpr int RP_SLOTS = 2 * RP_D;            // spin-word slots per warp: slot 0 = the block's own word
constexpr int RP_MAXBW = RP_SLOTS - 1;        // foreign spin words per block
constexpr int RP_CAP = 28 * RP_D;             // entries per slab
constexpr int RP_STAGES = 4;      // power of two: stage = g & 3, phase = (g >> 2) & 1 for the running block counter g
#ifndef RP_LA
#define RP_LA 2     // local fields are loaded this many variables ahead of their visit
#endif
#ifndef RP_PF_DIST
#define RP_PF_DIST 32   // L2 run-ahead of the field rows, in variables
#endif
constexpr int RP_DIST = 2;        // slabs are requested this many blocks ahead of the first warp that will need them
constexpr int RP_WARP_BYTES = RP_SLOTS * 32 * 8;  // per warp: {~S,F}[slot][lane]; the push phase stages its fields here
constexpr uint32_t RP_SLOT_MASK = (uint32_t)(RP_SLOTS - 1) << 8;
#ifndef RP_MAX_WARPS
#define RP_MAX_WARPS 8    // warps per CTA (upper bound; the host picks a power of two)
#endif
#ifndef RP_MIN_CTAS
#define RP_MIN_CTAS 2     // resident CTAs per SM the register allocation is bounded for
#endif

struct alignas(16) RpHdr {
    int32_t nent;               // entries in this slab
    int32_t nbw;                // distinct neighbour words other than the block's own word
    int32_t v0;                 // first variable of the block
    int32_t nv;                 // variables in the block (1 .. RP_D; v0 .. v0+nv-1 lie in one 32-variable spin word)
    uint32_t row[RP_D];         // entry range of variable i: start | (end << 16)
    uint16_t nlater[RP_D];      // leading entries of row i that refer to later variables (u > v)
    uint16_t deg[RP_D];
    int32_t ga[RP_D];           // group (low byte, 255: none) | coefficient << 8
    int32_t bw[RP_MAXBW];       // neighbour word indices, slot s+1 holds word bw[s]
    int32_t nbw_next;           // the same list for the NEXT block (cyclic): L2 run-ahead of its {S,F} rows
    int32_t bw_next[RP_MAXBW];
};
static_assert(sizeof(RpHdr) % 16 == 0, "slab header layout");
constexpr uint32_t RP_H_NBW = offsetof(RpHdr, nbw), RP_H_V0 = offsetof(RpHdr, v0), RP_H_NV = offsetof(RpHdr, nv),
                   RP_H_ROW = offsetof(RpHdr, row), RP_H_NLATER = offsetof(RpHdr, nlater),
                   RP_H_DEG = offsetof(RpHdr, deg), RP_H_GA = offsetof(RpHdr, ga), RP_H_BW = offsetof(RpHdr, bw),
                   RP_H_NBWN = offsetof(RpHdr, nbw_next), RP_H_BWN = offsetof(RpHdr, bw_next);
struct __align__(16) RpEntry {
    double J2;     // 2 * J
    int32_t j;     // neighbour (local variable index)
    uint32_t B;    // bits 0-4: 31 - (j & 31); bits 8-12: slot; bit 15: neighbour inside the same block
};
constexpr int RP_STAGE_BYTES = (int)sizeof(RpHdr) + RP_CAP * (int)sizeof(RpEntry);
static_assert(RP_STAGE_BYTES % 16 == 0, "stage alignment");

// ---- mbarrier / TMA 1-D primitives -----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// shared-memory accesses by explicit address (volatile: kept in program order among themselves and after the barrier waits)
__device__ __forceinline__ int4 lds_v4(uint32_t a) {
    int4 r;
    asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
    return r;
}
__device__ __forceinline__ uint2 lds_u2(uint32_t a) {
    uint2 r;
    asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(a));
    return r;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
    uint32_t r;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(a));
    return r;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
    uint32_t r;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(r) : "r"(a));
    return r;
}
__device__ __forceinline__ void sts_u2(uint32_t a, uint32_t x, uint32_t y) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y));
}
__device__ __forceinline__ double lds_f64(uint32_t a) {
    double r;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r) : "r"(a));
    return r;
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v));
}
// f += d on the lanes whose flag word has its top bit set (one predicated DADD: the only serial dependence of a replay)
__device__ __forceinline__ void add_f64_if_neg(double &f, double d, uint32_t t) {
    asm("{\n\t.reg .pred p;\n\tsetp.lt.s32 p, %2, 0;\n\t@p add.rn.f64 %0, %0, %1;\n\t}" : "+d"(f) : "d"(d), "r"(t));
}

struct RpCtx {
    // ring (CTA-shared)
    uint32_t full_s, empty_s, stage_s;   // shared addresses: full[0], empty[0] (8 B apart), stage 0
    unsigned *issued;                    // shared: running index of the next slab to request
    const unsigned char *slabs;          // global slab storage
    const uint32_t *off;                 // [nblk + 1] slab offsets of this problem, in 16-byte units
    uint32_t gb;                         // running index of the next block this warp consumes
    // per lane
    double *fT;                          // f[v][lane], this lane's column
    uint2 *SF;                           // {S,F}[word][lane], this lane's column
    uint32_t sfbase_s;                   // shared address of this lane's column in the warp's size-aligned region
    int *Mcol;
    int mstride;
    const double *lam;
    const long long *kap;
    const double *h;                     // linear biases of the current problem (field set-up pass)
    unsigned long long s0, s1;
    LaneStats st;
    bool active;
    int n, nblk, npad;
#ifdef QA_RP_PROFILE
    unsigned long long t_full, t_empty;  // cycles spent waiting for a slab / for a free stage
    unsigned long long t_setup, t_prol, t_ent, t_dec, t_pass, t_flip;
#endif
};

// lane 0 of any warp: request every slab up to running index `tgt` that nobody has requested yet.  `blk` is the block the
// calling warp is about to consume (running index c.gb); slabs follow the cyclic block order of the passes.
__device__ __forceinline__ void rp_request(RpCtx &c, uint32_t tgt, int blk) {
    unsigned cur = *reinterpret_cast<volatile unsigned *>(c.issued);
    while ((int)(tgt - cur) >= 0) {
        const unsigned old = atomicCAS(c.issued, cur, cur + 1u);
        if (old == cur) {
            int b = blk + (int)(cur - c.gb);
            while (b >= c.nblk) b -= c.nblk;
            const uint32_t st = cur & (RP_STAGES - 1), ph = (cur / RP_STAGES) & 1u;
#ifdef QA_RP_PROFILE
            const long long t0 = clock64();
#endif
            mbar_wait(c.empty_s + 8u * st, ph ^ 1u);     // every warp has released the previous use of that stage
#ifdef QA_RP_PROFILE
            c.t_empty += (unsigned long long)(clock64() - t0);
#endif
            const uint32_t o0 = __ldg(c.off + b), o1 = __ldg(c.off + b + 1);
            const uint32_t bytes = (o1 - o0) * 16u;
            mbar_expect_tx(c.full_s + 8u * st, bytes);
            tma_load_1d(c.stage_s + st * (uint32_t)RP_STAGE_BYTES, c.slabs + (size_t)o0 * 16u, bytes, c.full_s + 8u * st);
            cur = cur + 1u;
        } else {
            cur = old;
        }
    }
}

// One pass over all blocks.  MODE 0: replay sweep (pull).  MODE 1: catch-up (pending later-neighbour updates only, no
// decisions).  MODE 2: push sweep (neal's eager form).  MODE 3: field set-up f[v] = h_v + sum_j J_vj s_j in neal's
// get_flip_energy order, for models whose adjacency lists are ascending (then adjacency order = the slab's earlier part
// followed by its later part).  Returns the number of accepted flips of the warp.
// GROUPS: 0 none, 1 rank-1 groups with 32-bit exact integer arithmetic (host-checked ranges), 2 with 64-bit.
// VAR: blocks of variable size (read from the slab header); false = every block holds exactly RP_D variables, which lets the
// compiler keep the block geometry in immediates (22 % faster on config 3: the kernel sits at the 128-register limit).
template <int MODE, int GROUPS, bool VAR>
__device__ unsigned rp_pass(RpCtx &c, double beta, bool has_next_pass) {
    const int lane = threadIdx.x & 31;
    const double thr = 44.36142 / beta;
    const int n = c.n, nblk = c.nblk;
    double *const fT = c.fT;
    uint2 *const SF = c.SF;
    const uint32_t sfb = c.sfbase_s;
    const bool active = c.active;
    unsigned sweep_acc = 0;
    uint32_t S = 0xffffffffu, F = 0u, S0 = 0xffffffffu, F0 = 0u;

#ifdef QA_RP_PROFILE
    const long long tp0 = clock64();
#endif
    const uint32_t g_last = c.gb + (uint32_t)nblk - 1u;   // running index of the last block of this pass
    for (int blk = 0; blk < nblk; ++blk) {
        const uint32_t stage = c.gb & (RP_STAGES - 1), phase = (c.gb / RP_STAGES) & 1u;
        if (lane == 0) {
            uint32_t tgt = c.gb + (uint32_t)RP_DIST;
            if (!has_next_pass && (int)(tgt - g_last) > 0) tgt = g_last;
            rp_request(c, tgt, blk);
        }
        __syncwarp();
#ifdef QA_RP_PROFILE
        const long long tw0 = clock64();
#endif
        mbar_wait(c.full_s + 8u * stage, phase);
#ifdef QA_RP_PROFILE
        if (lane == 0) c.t_full += (unsigned long long)(clock64() - tw0);
#endif
        const uint32_t hdr = c.stage_s + stage * (uint32_t)RP_STAGE_BYTES;   // shared address of the slab
        const uint32_t ent = hdr + (uint32_t)sizeof(RpHdr);
        const int v0 = VAR ? (int)lds_u32(hdr + RP_H_V0) : blk * RP_D;
        const int nv = VAR ? (int)lds_u32(hdr + RP_H_NV) : RP_D;
        const int wi = v0 >> 5;
        const int sub = v0 & 31;
        double *const fB = fT + (int64_t)v0 * 32;
        if (sub == 0) {
            const uint2 own = __ldcg(SF + (int64_t)wi * 32);
            S = own.x; F = own.y; S0 = S; F0 = F;
        }
#ifndef RP_NO_PREFETCH
        {   // L2 run-ahead of the field rows RP_PF_DIST blocks on: 16 rows x 256 B = 128 sectors, one sector per lane and instruction
            int pv = (v0 & ~(RP_D - 1)) + RP_PF_DIST;   // a 16-row group ahead (wraps into the next sweep)
            const int npad = VAR ? c.npad : nblk * RP_D;
            if (pv >= npad) pv -= npad;
            const char *pa = reinterpret_cast<const char *>(fT - lane + (int64_t)pv * 32) + lane * 32;
#pragma unroll
            for (int k = 0; k < 4; ++k) asm volatile("prefetch.global.L2 [%0];" ::"l"(pa + k * 1024));
        }
        if (MODE <= 1 || MODE == 3) {   // ... and of the {S,F} rows the next block will stage (8 sectors per word)
            const int ns = (int)lds_u32(hdr + RP_H_NBWN) * 8;
            const char *sfrow = reinterpret_cast<const char *>(SF - lane);
            for (int k = lane; k < ns; k += 32) {
                const int64_t w = (int64_t)lds_u32(hdr + RP_H_BWN + 4u * (uint32_t)(k >> 3));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(sfrow + w * 256 + (k & 7) * 32));
            }
            if (sub != 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(sfrow + (int64_t)((wi + 1) * 32 < (VAR ? c.npad : nblk * RP_D) ? wi + 1 : 0) * 256 + (lane & 7) * 32));
        }
#endif
        bool blk_dirty = false;
        if (MODE <= 1 || MODE == 3) {
            // this lane's copy of every spin/flag word the block refers to (~S so that a set top bit means "spin down")
            const int nbw = (int)lds_u32(hdr + RP_H_NBW);
            for (int s = 0; s < nbw; s += 8) {
                uint2 t[8];
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (s + q < nbw) t[q] = __ldcg(SF + (int64_t)lds_u32(hdr + RP_H_BW + 4u * (uint32_t)(s + q)) * 32);
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (s + q < nbw) sts_u2(sfb + (uint32_t)(s + q + 1) * 256u, ~t[q].x, t[q].y);
            }
            sts_u2(sfb, ~S, F);
        } else {
            double t[RP_D];
#pragma unroll
            for (int i = 0; i < RP_D; ++i)
                if (i < nv) t[i] = __ldcg(fB + i * 32);
#pragma unroll
            for (int i = 0; i < RP_D; ++i)
                if (i < nv) sts_f64(sfb + (uint32_t)i * 256u, t[i]);
        }
        double fq[RP_LA];   // fields of the next RP_LA variables, loaded ahead of their visit
#pragma unroll
        for (int k = 0; k < RP_LA; ++k) fq[k] = 0.0;
        if (MODE <= 1) {
#pragma unroll
            for (int k = 0; k < RP_LA; ++k)
                if (k < nv) fq[k] = __ldcg(fB + k * 32);
        }
        const int ilim = min(nv, n - v0);   // uniform: padding variables are not visited
#ifdef QA_RP_PROFILE
        c.t_prol += (unsigned long long)(clock64() - tw0) ;
#endif

        for (int i = 0; i < ilim; ++i) {
            const uint32_t rw = lds_u32(hdr + RP_H_ROW + 4u * (uint32_t)i);
            const uint32_t a0 = ent + (rw & 0xffffu) * 16u;
            const uint32_t bit = 1u << (sub + i);
            const bool up = (S & bit) != 0u;
            double fv;
            if (MODE == 3) {
                // h_v, then the earlier neighbours (ascending), then the later ones: 0.5 * (+-2J) is exact, one rounding per term
                fv = __ldg(c.h + v0 + i);
                const uint32_t am = a0 + (lds_u16(hdr + RP_H_NLATER + 2u * (uint32_t)i) & 0xfffu) * 16u;
                const uint32_t a1 = ent + (rw >> 16) * 16u;
                auto add1 = [&](const int4 &q, const uint2 &sf) {
                    const uint32_t sg = __funnelshift_l(0u, sf.x, (uint32_t)q.w) & 0x80000000u;   // spin down: -J
                    fv = fma(__hiloint2double(q.y ^ (int)sg, q.x), 0.5, fv);
                };
#pragma unroll
                for (int seg = 0; seg < 2; ++seg) {
                    uint32_t a = seg == 0 ? am : a0;
                    const uint32_t ae = seg == 0 ? a1 : am;
                    for (; a + 64u <= ae; a += 64u) {
                        int4 q[4];
                        uint2 sf[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) q[k] = lds_v4(a + 16u * k);
#pragma unroll
                        for (int k = 0; k < 4; ++k) sf[k] = lds_u2(((uint32_t)q[k].w & RP_SLOT_MASK) | sfb);
#pragma unroll
                        for (int k = 0; k < 4; ++k) add1(q[k], sf[k]);
                    }
                    for (; a < ae; a += 16u) {
                        const int4 q = lds_v4(a);
                        add1(q, lds_u2(((uint32_t)q.w & RP_SLOT_MASK) | sfb));
                    }
                }
                __stcg(fB + i * 32, fv);
                continue;
            } else if (MODE <= 1) {
                fv = fq[0];
#pragma unroll
                for (int k = 0; k + 1 < RP_LA; ++k) fq[k] = fq[k + 1];
                if (i + RP_LA < nv) fq[RP_LA - 1] = __ldcg(fB + (i + RP_LA) * 32);
                const double f0 = fv;
#ifdef QA_RP_PROFILE
                const long long te0 = clock64();
#endif
                // a neighbour that did not flip contributes -0.0, the exact additive identity: the only serial dependence of
                // a replay is one DADD per entry
                auto replay1 = [&](const int4 &q, const uint2 &sf) {
                    const uint32_t B = (uint32_t)q.w;
                    const uint32_t sg = __funnelshift_l(0u, sf.x, B) & 0x80000000u;   // spin down: add -2J
                    const bool flagged = (int)__funnelshift_l(0u, sf.y, B) < 0;
                    const int hi = flagged ? (q.y ^ (int)sg) : (int)0x80000000u;
                    const int lo = flagged ? q.x : 0;
                    fv += __hiloint2double(hi, lo);
                };
                const uint32_t a1 = (MODE == 0) ? ent + (rw >> 16) * 16u
                                                : a0 + lds_u16(hdr + RP_H_NLATER + 2u * (uint32_t)i) * 16u;
                uint32_t a = a0;
                // four entries per round: table entries first, then the lane's {~S,F} words, then the ordered additions
                for (; a + 64u <= a1; a += 64u) {
                    int4 q[4];
                    uint2 sf[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) q[k] = lds_v4(a + 16u * k);
#pragma unroll
                    for (int k = 0; k < 4; ++k) sf[k] = lds_u2(((uint32_t)q[k].w & RP_SLOT_MASK) | sfb);
#pragma unroll
                    for (int k = 0; k < 4; ++k) replay1(q[k], sf[k]);
                }
                for (; a < a1; a += 16u) {
                    const int4 q = lds_v4(a);
                    const uint2 sf = lds_u2(((uint32_t)q.w & RP_SLOT_MASK) | sfb);
                    replay1(q, sf);
                }
                if (fv != f0) __stcg(fB + i * 32, fv);
#ifdef QA_RP_PROFILE
                c.t_ent += (unsigned long long)(clock64() - te0);
#endif
                if (MODE == 1) continue;
            } else {
                fv = lds_f64(sfb + (uint32_t)i * 256u);
            }
#ifdef QA_RP_PROFILE
            const long long td0 = clock64();
#endif
            double dE = up ? -2.0 * fv : 2.0 * fv;
            int g = 255, a = 0;
            if (GROUPS) {
                const int ga = (int)lds_u32(hdr + RP_H_GA + 4u * (uint32_t)i);
                g = ga & 255;
                if (g != 255) {
                    a = ga >> 8;
                    if (GROUPS == 1) {
                        const int m = c.Mcol[g * c.mstride] + reinterpret_cast<const int *>(c.kap)[2 * g];
                        const int t = a * (a - (up ? m : -m));
                        dE = dE + c.lam[g] * (double)t;
                    } else {
                        const long long t = (long long)a * ((long long)a - (up ? 1 : -1) * ((long long)c.Mcol[g * c.mstride] + c.kap[g]));
                        dE = dE + c.lam[g] * (double)t;
                    }
                }
            }
            const bool cand = active && !(dE >= thr);
            bool acc = false;
            if (__any_sync(FULL_MASK, cand)) {
                if (cand) c.st.cand++;
                acc = ls_accept(dE, cand, beta, c.s0, c.s1, c.st);
            }
#ifdef QA_RP_PROFILE
            c.t_dec += (unsigned long long)(clock64() - td0);
#endif
            if (MODE == 0) {
                // F[v] := accepted (also when nothing was accepted: the flag of the previous visit must be cleared)
                F = acc ? (F | bit) : (F & ~bit);
                if (acc) S ^= bit;
                sts_u2(sfb, ~S, F);
                if (acc) {
                    c.st.acc++;
                    c.st.nbr += (unsigned long long)lds_u16(hdr + RP_H_DEG + 2u * (uint32_t)i);
                    if (GROUPS) {
                        if (g != 255) c.Mcol[g * c.mstride] -= 2 * a * (up ? 1 : -1);
                    }
                }
                sweep_acc += __popc(__ballot_sync(FULL_MASK, acc));
            } else {
                const unsigned accm = __ballot_sync(FULL_MASK, acc);
                if (accm == 0u) continue;
                sweep_acc += __popc(accm);
#ifdef QA_RP_PROFILE
                const long long tf0 = clock64();
#endif
                const int sgn = up ? (int)0x80000000u : 0;   // f[j] += -2 s_v J
                const uint32_t deg = lds_u16(hdr + RP_H_DEG + 2u * (uint32_t)i);
                const uint32_t a1 = a0 + deg * 16u;
#pragma unroll 4
                for (uint32_t ad = a0; ad < a1; ad += 16u) {
                    const int4 q = lds_v4(ad);
                    const double d = __hiloint2double(q.y ^ sgn, q.x);
                    if ((uint32_t)q.w & 0x8000u) {   // uniform: neighbour staged in this block
                        const uint32_t ca = sfb + ((uint32_t)(q.z - v0) << 8);
                        if (acc) sts_f64(ca, lds_f64(ca) + d);
                        blk_dirty = true;
                    } else {
                        red_add_f64_if(fT + (int64_t)q.z * 32, d, acc);
                    }
                }
#ifdef QA_RP_PROFILE
                c.t_flip += (unsigned long long)(clock64() - tf0);
#endif
                if (acc) {
                    S ^= bit;
                    c.st.acc++;
                    c.st.nbr += (unsigned long long)deg;
                    if (GROUPS) {
                        if (g != 255) c.Mcol[g * c.mstride] -= 2 * a * (up ? 1 : -1);
                    }
                }
            }
        }
        if (MODE == 2) {
            if (blk_dirty) {  // uniform: write the staged fields back (coalesced 256 B rows)
#pragma unroll
                for (int i = 0; i < RP_D; ++i)
                    if (i < nv) __stcg(fB + i * 32, lds_f64(sfb + (uint32_t)i * 256u));
            }
        }
        if (MODE != 1 && MODE != 3 && sub + nv == 32) {
            if (S != S0 || F != F0) __stcg(SF + (int64_t)wi * 32, make_uint2(S, F));
        }
        // release the stage: one arrive per warp
        __syncwarp();
        if (lane == 0) mbar_arrive(c.empty_s + 8u * stage);
        ++c.gb;
    }
#ifdef QA_RP_PROFILE
    c.t_pass += (unsigned long long)(clock64() - tp0);
#endif
    return sweep_acc;
}

// The set-up pass as a real call with the context passed BY VALUE: its registers do not add to the pressure of the sweep
// loops (the kernel sits at the 128-register limit and every spilled value there costs measurable time).
template <bool VAR>
__device__ __noinline__ uint32_t rp_setup_pass(RpCtx c) {
    rp_pass<3, 0, VAR>(c, 1.0, true);
    return c.gb;
}

// local field of v in neal's get_flip_energy order (adjacency order), spins from the {S,F} scratch; loads batched by 8
__device__ __forceinline__ double rp_field_direct(const ProblemDesc &D, const uint2 *SF, int v, int e0, int e1) {
    double fv = __ldg(D.h + v);
    for (int e = e0; e < e1; e += 8) {
        int jq[8];
        uint32_t wq[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            jq[q] = __ldg(D.col + min(e + q, e1 - 1));
            wq[q] = __ldcg(SF + (int64_t)(jq[q] >> 5) * 32).x;
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

// Shared-memory layout: [slab stages][mbarriers][CTA scratch][lambda][kappa][group counters] then, aligned to its size in the
// shared window (so that a slot address is formed by OR), one RP_WARP_BYTES region per warp.  `smem_base` is the shared-window
// offset of the dynamic allocation (1 KB on sm_100: the system-reserved bytes); the kernel verifies the assumption.
__host__ __device__ inline size_t rp_fixed_bytes(int nw, int max_groups) {
    size_t b = (size_t)RP_STAGES * RP_STAGE_BYTES + 8 * (2 * RP_STAGES) + 32;
    b += (sizeof(double) + sizeof(long long)) * (size_t)max_groups;
    b += sizeof(int) * (size_t)max_groups * nw * 32;
    return b;
}
__host__ __device__ inline size_t rp_smem_bytes(int nw, int max_groups, unsigned smem_base) {
    const size_t fixed = rp_fixed_bytes(nw, max_groups);
    const size_t pad = (RP_WARP_BYTES - (smem_base + fixed) % RP_WARP_BYTES) % RP_WARP_BYTES;
    return fixed + pad + (size_t)nw * RP_WARP_BYTES;
}
constexpr int QA_ERR_SMEM_BASE = -100;   // internal: the dynamic shared memory does not start where the host assumed

template <int GROUPS, bool VAR>
__global__ void __launch_bounds__(RP_MAX_WARPS * 32, RP_MIN_CTAS) k_anneal_replay(AnnealParams P) {
    extern __shared__ __align__(16) unsigned char rp_raw[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int NW = blockDim.x >> 5;
    const uint32_t raw_s = (uint32_t)__cvta_generic_to_shared(rp_raw);
    const uint32_t fixed = (uint32_t)rp_fixed_bytes(NW, P.max_groups);
    const uint32_t base_s = (raw_s + fixed + (uint32_t)RP_WARP_BYTES - 1u) & ~((uint32_t)RP_WARP_BYTES - 1u);    // warp regions
    {
        uint32_t dyn;
        asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
        if (base_s - raw_s + (uint32_t)NW * RP_WARP_BYTES > dyn) {   // uniform: report the real base and leave
            if (threadIdx.x == 0) {
                P.error_flag[1] = (int)raw_s;
                atomicExch(P.error_flag, QA_ERR_SMEM_BASE);
            }
            return;
        }
    }
    unsigned char *tail = rp_raw + (size_t)RP_STAGES * RP_STAGE_BYTES;
    const uint32_t bars_s = raw_s + (uint32_t)RP_STAGES * RP_STAGE_BYTES;
    unsigned *issued_sh = reinterpret_cast<unsigned *>(tail + 8 * (2 * RP_STAGES));
    long long *item_sh = reinterpret_cast<long long *>(tail + 8 * (2 * RP_STAGES) + 8);
    unsigned *acc_sh = reinterpret_cast<unsigned *>(tail + 8 * (2 * RP_STAGES) + 16);   // [2]
    double *lam_sh = reinterpret_cast<double *>(tail + 8 * (2 * RP_STAGES) + 32);
    long long *kap_sh = reinterpret_cast<long long *>(lam_sh + P.max_groups);
    int *M_all = reinterpret_cast<int *>(kap_sh + P.max_groups);

    if (threadIdx.x == 0) {
        for (int s = 0; s < RP_STAGES; ++s) {
            mbar_init(bars_s + 8u * s, 1);                     // full: the requester's arrive.expect_tx
            mbar_init(bars_s + 8u * (RP_STAGES + s), NW);      // empty: one arrive per warp
        }
        *issued_sh = 0u;
        acc_sh[0] = 0u;
        acc_sh[1] = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (GROUPS) {  // groups exist only on single-problem models: one copy of lambda / kappa per block
        const ProblemDesc &D0 = P.descs[0];
        for (int g = threadIdx.x; g < D0.ngroups; g += blockDim.x) {
            lam_sh[g] = D0.lambda[g];
            kap_sh[g] = D0.kappa[g];
        }
    }
    __syncthreads();

    RpCtx c;
    c.full_s = bars_s;
    c.empty_s = bars_s + 8u * RP_STAGES;
    c.stage_s = raw_s;
    c.issued = issued_sh;
    c.gb = 0u;
#ifdef QA_RP_PROFILE
    c.t_full = 0ull;
    c.t_empty = 0ull;
    c.t_setup = c.t_prol = c.t_ent = c.t_dec = c.t_pass = c.t_flip = 0ull;
#endif
    c.sfbase_s = base_s + (uint32_t)wib * RP_WARP_BYTES + (uint32_t)lane * 8u;
    c.Mcol = M_all + threadIdx.x;
    c.mstride = (int)blockDim.x;
    c.lam = lam_sh;
    c.kap = kap_sh;
    c.st = {0, 0, 0, 0, 0};
    const int64_t slot = (int64_t)blockIdx.x * NW + wib;
    c.fT = P.fT_scratch + slot * P.fT_stride + lane;
    c.SF = reinterpret_cast<uint2 *>(P.sf_scratch) + slot * P.sf_stride + lane;
    int acc_par = 0;

    for (;;) {
        if (threadIdx.x == 0) *item_sh = (long long)atomicAdd(P.counter, 1ull);
        __syncthreads();
        const int64_t item = *item_sh;
        __syncthreads();
        if (item >= P.total_items) break;
        const int p = (int)(item / P.groups_per_problem);
        const int64_t q = item % P.groups_per_problem;
        const ProblemDesc D = P.descs[p];
        const int64_t tile = q * NW + wib;
        const int64_t r = tile * 32 + lane;
        const bool active = r < D.reads;
        const long long cta_reads = min((long long)NW * 32, (long long)D.reads - (long long)q * NW * 32);
        c.active = active;
        c.n = D.n;
        c.nblk = D.rp_nslabs;
        c.npad = D.nch * 32;
        c.slabs = D.rp_slabs;
        c.h = D.h;
        c.off = D.rp_off;
        const long long total_sweeps = (long long)P.num_betas * P.sweeps_per_beta;
        if (lane == 0 && total_sweeps > 0) rp_request(c, c.gb + (uint32_t)RP_DIST, 0);   // overlaps with the set-up below

#ifdef QA_RP_PROFILE
        const long long ts0 = clock64();
#endif
        const unsigned long long sd = active ? P.seeds[D.read_base + r] : 1ull;
        c.s0 = sd ? sd : ~0ull;
        c.s1 = 0;
        const int n = D.n, nch = D.nch;
        // ---- pack this read's +-1 bytes (padding lanes and padding variables are +1), clear the flags
        for (int wi = 0; wi < nch; ++wi) {
            uint32_t w = 0xffffffffu;
            if (active) {
                const int8_t *row = D.states + r * (int64_t)n + wi * 32;
                const int lim = min(32, n - wi * 32);
                for (int i = 0; i < lim; ++i) {
                    const int s = row[i];
                    if (s != 1 && s != -1) atomicExch(P.error_flag, QA_ERR_STATE);
                    if (s < 0) w &= ~(1u << i);
                }
            }
            __stcg(c.SF + (int64_t)wi * 32, make_uint2(w, 0u));
        }
        if (GROUPS) {
            for (int g = 0; g < D.ngroups; ++g) c.Mcol[g * c.mstride] = 0;
            for (int wi = 0; wi < nch; ++wi) {
                const uint32_t w = __ldcg(c.SF + (int64_t)wi * 32).x;
                for (int i = 0; i < 32; ++i) {
                    const int v = wi * 32 + i;
                    const int g = __ldg(D.grp + v);  // uniform
                    if (g >= 0) {
                        const int a = __ldg(D.coef + v);
                        c.Mcol[g * c.mstride] += ((w >> i) & 1u) ? a : -a;
                    }
                }
            }
        }
        // ---- local fields in neal's get_flip_energy order: through the slab ring when the adjacency lists are ascending,
        // else row by row from the CSR
        if (P.rp_slab_init) {
            if (total_sweeps > 0) c.gb = rp_setup_pass<VAR>(c);
        } else {
            int e0 = __ldg(D.rowptr);
            for (int v = 0; v < n; ++v) {
                const int e1 = __ldg(D.rowptr + v + 1);
                __stcg(c.fT + (int64_t)v * 32, rp_field_direct(D, c.SF, v, e0, e1));
                e0 = e1;
            }
        }
#ifdef QA_RP_PROFILE
        c.t_setup += (unsigned long long)(clock64() - ts0);
#endif
        // ---- the schedule: replay sweeps while flips are frequent, then one catch-up pass and push sweeps
        bool push = false;
        long long done = 0;
        for (int bi = 0; bi < P.num_betas; ++bi) {
            const double beta = P.betas[bi];
            for (int swi = 0; swi < P.sweeps_per_beta; ++swi) {
                ++done;
                const bool more = done < total_sweeps;
#ifdef QA_RP_PROFILE
                if (blockIdx.x == 0 && threadIdx.x == 0)
                    printf("[qa sweep] %lld beta %.4g mode %s clock %lld\n", done, beta, push ? "push" : "replay", clock64());
#endif
                if (push) {
                    rp_pass<2, GROUPS, VAR>(c, beta, more);
                } else {
                    const unsigned wacc = rp_pass<0, GROUPS, VAR>(c, beta, more);
                    if (more) {  // CTA-uniform hand-over decision
                        if (lane == 0) atomicAdd(acc_sh + acc_par, wacc);
                        __syncthreads();
                        const unsigned long long tot = acc_sh[acc_par];
                        if (threadIdx.x == 0) acc_sh[acc_par ^ 1] = 0u;
                        acc_par ^= 1;
                        if (tot * 1000ull < (unsigned long long)P.switch_permille * (unsigned long long)n * (unsigned long long)cta_reads) {
                            rp_pass<1, 0, VAR>(c, beta, true);
                            push = true;
                        }
                    }
                }
            }
        }
        // ---- final spins: packed transposed layout for the energy kernel, +-1 bytes for the caller
        if (active) {
            for (int wi = 0; wi < nch; ++wi) {
                const uint32_t w = __ldcg(c.SF + (int64_t)wi * 32).x;
                D.packedT[(int64_t)wi * D.rpad + r] = w;
                int8_t *row = D.states + r * (int64_t)n + wi * 32;
                const int lim = min(32, n - wi * 32);
                for (int i = 0; i < lim; ++i) row[i] = ((w >> i) & 1u) ? 1 : -1;
            }
        }
    }
    // warp-reduce the per-lane counters
    unsigned long long v[5] = {c.st.cand, c.st.draws, c.st.acc, c.st.ties, c.st.nbr};
#pragma unroll
    for (int q = 0; q < 5; ++q)
        for (int off = 16; off > 0; off >>= 1) v[q] += __shfl_xor_sync(FULL_MASK, v[q], off);
#ifdef QA_RP_PROFILE
    if (lane == 0) {   // development build: cycles spent waiting on the ring, reported through the otherwise unused counters
        unsigned long long *dbg = P.stats + QA_NSTAT + 1;
        atomicAdd(dbg + 0, c.t_setup);
        atomicAdd(dbg + 1, c.t_full);
        atomicAdd(dbg + 2, c.t_empty);
        atomicAdd(dbg + 3, c.t_prol);
        atomicAdd(dbg + 4, c.t_ent);
        atomicAdd(dbg + 5, c.t_dec);
        atomicAdd(dbg + 6, c.t_pass);
        atomicAdd(dbg + 7, c.t_flip);
    }
#endif
    if (lane == 0) {
        atomicAdd(P.stats + ST_CAND, v[0]);
        atomicAdd(P.stats + ST_DRAWS, v[1]);
        atomicAdd(P.stats + ST_ACC, v[2]);
        atomicAdd(P.stats + ST_TIES, v[3]);
        atomicAdd(P.stats + ST_NBR, v[4]);
    }
}
