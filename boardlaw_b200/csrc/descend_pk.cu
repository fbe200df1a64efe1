// Packed tree descent for sm_100a (variant 7) — same arithmetic as descend.cu (boardlaw/mcts/cpp/cuda.cu:35-99,138-182), a
// different mapping of the work onto the SM.
//
// What the one-lane kernel (descend.cu) loses, measured (profiles/r01_phase_clock_final.txt): a warp's 32 envs
// share one instruction stream, so an env that needs a Newton pass waits while its neighbours are being "serviced" (child
// records fetched and normalised, lambda, alpha seed: a chain of ~1 000 dependent instructions and two memory round trips
// executed for ~11 lanes at a time) and the other way round: 58 trips per warp where an env needs 22.8 passes, 36 % of the
// lanes busy in a pass.  Here ONE CTA per SM owns 224 envs and splits the two kinds of work over different warps:
//
//   * PASS warps (one lane = one env's current evaluation): a lane claims an env whose node is ready (bit masks in shared memory,
//     one word per env residue mod 8 so that the lane-private rows stay bank-conflict free), loads lambda*pi into REGISTERS and
//     runs exactly descend.cu's pass: child terms by exact division, the two sequential fp32 sums over packed FMUL2/FFMA2
//     Markstein quotients, the Newton update — then, on convergence, the inverse-CDF search over the running sums, the step
//     to the child in O(1) (child entries are kept sorted by action: index = popcount of the child mask below the action) and
//     the asynchronous fetch of the child's row / summary / children mask (cp.async tracked by a per-env mbarrier).  The lane
//     then drops the env and claims the next ready one: pass warps only ever execute pass code, with (nearly) all lanes busy;
//   * SERVICE warps visit a node WARP-COOPERATIVELY, one env at a time: lane i fetches and normalises child i (the q
//     normalisation's two divisions run in parallel over the children instead of in a per-lane loop), N / the child mask /
//     the alpha seed are warp reductions (integer sum, OR, max: order-free, so bit-identical), the entries are written in action
//     order, and the env is published to the pass lanes.  Two ready envs are visited at a time so that the L2 round trip of
//     one's child records overlaps the other's.
//
// The exact serial fallback (rows with denormal-range values, a child term outside the safe range of the branch-free division,
// more children than entry slots) runs the reference loops verbatim on one lane of a service warp, from global memory.
// Shapes: A <= 84 (row of at most 21 four-element chunks), T <= 64 (one 64-bit children mask); others take the other kernels.
// Every wait is bounded: a stuck hand-over raises the tree's error counter and ends the launch instead of hanging the GPU.
//
// Compiled with -fmad=false -prec-div=true -ftz=false (see build.py); fused operations are explicit.
#include <cstdio>
#include <cstdlib>

#include "descend_common.cuh"

namespace {

constexpr int PK_EPC = 224;          // envs per CTA
constexpr int PK_WARPS = 12;
constexpr int PK_NT = PK_WARPS * 32;
constexpr int PK_KS = 34;            // child entries per env in shared memory (8 B each; pitch 17 x 16 B: odd)
constexpr int PK_IDLE_LIMIT = 1 << 21;

enum { E_NONE = 0, E_FETCH = 1, E_VISITING = 2, E_READY = 3, E_RUNNING = 4, E_SLOW = 5, E_DONE = 6 };

struct __align__(16) PkCtl {         // 48 bytes per env
    float alpha0, lambda, r, c_puct;
    u64 cm0, cm1;                    // bit a: action a has a child
    int16_t cur, parent, action, leaf;
    uint8_t nc, seat, pad0, pad1;
    uint32_t nzpos;
};
static_assert(sizeof(PkCtl) == 48, "control block is three 16-byte chunks (odd pitch)");

template <int NCH>
struct PkCfg {
    static constexpr int PS = 4 * NCH;
    static constexpr size_t ROW_BYTES = (size_t)PK_EPC * PS * 4;
    static constexpr size_t ENT_BYTES = (size_t)PK_EPC * PK_KS * 8;
    static constexpr size_t CTL_BYTES = (size_t)PK_EPC * sizeof(PkCtl);
    static constexpr size_t ST_BYTES = (size_t)PK_EPC * 4;
    static constexpr size_t MBAR_BYTES = (size_t)PK_EPC * 8;
    static constexpr size_t SMEM = 2 * ROW_BYTES + ENT_BYTES + CTL_BYTES + ST_BYTES + MBAR_BYTES + 64;
};

__device__ __forceinline__ void pk_cp16(uint32_t dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void pk_cp8(uint32_t dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ bool pk_mbar_test(uint32_t addr, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ uint32_t pk_ld_vol(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void pk_st_vol(uint32_t *p, uint32_t v) {
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
// k-th (0-based) set bit of a 64-bit mask
__device__ __forceinline__ int pk_kth(u64 m, int k) {
    const uint32_t l = (uint32_t)m, h = (uint32_t)(m >> 32);
    const int cl = __popc(l);
    return k < cl ? (int)__fns(l, 0, k + 1) : 32 + (int)__fns(h, 0, k - cl + 1);
}
__device__ __forceinline__ uint32_t pk_pack_meta(float q, int a, int id, int seat, int term) {
    return (uint32_t)bl_f2h(q) | ((uint32_t)a << 16) | ((uint32_t)id << 23) | ((uint32_t)seat << 29) | ((uint32_t)term << 30);
}

// everything a visit of node `n` of env `b` needs that has a known address: row summary -> pg[0..3], children mask -> pg[4..5],
// pi row -> the (dead) sums row; completion is signalled on the env's mbarrier (one arrival per phase)
template <int NCH>
__device__ __forceinline__ void pk_prefetch(const bl_tree &t, int b, int n, uint32_t ps_addr, uint32_t pg_addr, uint32_t mbar_addr) {
    const size_t slot = (size_t)b * t.T + n;
    const int nrow4 = t.AP >> 2;
    pk_cp16(pg_addr, t.aux + slot);
    pk_cp8(pg_addr + 16u, t.kids + slot);
    const float4 *row = reinterpret_cast<const float4 *>(t.pi + slot * t.AP);
#pragma unroll
    for (int c = 0; c < NCH; c++)
        if (c < nrow4) pk_cp16(ps_addr + 16u * c, row + c);
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(mbar_addr) : "memory");
}

template <int NCH>
struct PkShared {
    float *ps, *pg;
    uint2 *ent;
    PkCtl *ctl;
    uint32_t *st;
    u64 *mbar;
    uint32_t *ready;     // 8 words: bit j of word r = env slot 8 j + r is ready for a pass lane
    int *n_done, *err;
    __device__ __forceinline__ PkShared(void *raw) {
        using C = PkCfg<NCH>;
        uint8_t *p = reinterpret_cast<uint8_t *>(raw);
        ps = reinterpret_cast<float *>(p); p += C::ROW_BYTES;
        pg = reinterpret_cast<float *>(p); p += C::ROW_BYTES;
        ent = reinterpret_cast<uint2 *>(p); p += C::ENT_BYTES;
        ctl = reinterpret_cast<PkCtl *>(p); p += C::CTL_BYTES;
        st = reinterpret_cast<uint32_t *>(p); p += C::ST_BYTES;
        mbar = reinterpret_cast<u64 *>(p); p += C::MBAR_BYTES;
        ready = reinterpret_cast<uint32_t *>(p); p += 32;
        n_done = reinterpret_cast<int *>(p);
        err = n_done + 1;
    }
};

// the descent of env slot s steps from node `cur` along `action` (one thread): to an existing non-terminal child -> its data starts
// travelling and the env waits for a visit; anything else ends the descent (new leaf, existing terminal child, no legal action)
template <int NCH>
__device__ __forceinline__ void pk_advance(const bl_tree &t, const PkShared<NCH> &sh, int s, int b, int cur, int action, bool has, int id,
                                           int cseat, int cterm) {
    constexpr int PS = 4 * NCH;
    PkCtl *c = sh.ctl + s;
    const int next = (action >= 0 && has) ? id : -1;
    if (next >= 0 && !cterm) {
        c->parent = (int16_t)cur; c->cur = (int16_t)next; c->seat = (uint8_t)cseat;
        pk_prefetch<NCH>(t, b, next, smem_u32(sh.ps + (size_t)s * PS), smem_u32(sh.pg + (size_t)s * PS), smem_u32(sh.mbar + s));
        __threadfence_block();
        pk_st_vol(sh.st + s, E_FETCH);
    } else {
        c->leaf = (int16_t)next; c->parent = (int16_t)cur; c->action = (int16_t)action;
        t.leaf[b] = (int16_t)next;
        t.leaf_parent[b] = (int16_t)cur;
        t.leaf_action[b] = (int16_t)action;
        __threadfence_block();
        pk_st_vol(sh.st + s, E_DONE);
        atomicAdd(sh.n_done, 1);
    }
}

// exact serial evaluation of env slot s's current node by a service warp: the reference loops verbatim on lane 0, operands
// rebuilt from global memory (tops -> the sums row, q at the child positions -> the g row)
template <int NCH>
__device__ __noinline__ uint4 pk_slow(const bl_tree &t, const PkShared<NCH> &sh, int s, int b, const bl_qnorm &qn, uint64_t keep, bool count) {
    unsigned c_evals = 0, c_children = 0, c_iters = 0, c_desc = 0;      // returned: what lane 0 adds to the counters
    constexpr int PS = 4 * NCH;
    const int lane = threadIdx.x & 31, A = t.A, T = t.T;
    PkCtl *c = sh.ctl + s;
    const int cur = c->cur, seat = c->seat;
    const float c_puct = c->c_puct, r = c->r;
    float *ps = sh.ps + (size_t)s * PS, *pg = sh.pg + (size_t)s * PS;
    const size_t node0 = (size_t)b * T;
    const u64 mm = t.kids[node0 + cur];
    for (int a = lane; a < PS; a += 32) pg[a] = 0.f;
    __syncwarp();
    const int nc = __popcll(mm);
    int N = 0;
    for (int k0 = 0; k0 < nc; k0 += 32) {
        const int k = k0 + lane;
        if (k < nc) {
            const bl_node ch = bl_ld_node_hint(t.node + node0 + pk_kth(mm, k), keep);
            pg[ch.relation] = qn.fast(seat ? ch.w[1] : ch.w[0], ch.n);
            N += ch.n;
        }
    }
    N = __reduce_add_sync(FULL, N) + A - nc;
    const float lambda = bl_lambda(c_puct, N, A);
    const float *row = t.pi + (node0 + cur) * t.AP;
    for (int a = lane; a < PS; a += 32) ps[a] = a < A ? __fmul_rn(lambda, row[a]) : 0.f;
    __syncwarp();
    int action = -1, iters = 0;
    if (lane == 0) {
        auto topf = [&](int a) { return ps[a]; };
        auto qf = [&](int a) { return pg[a]; };
        const float al = bl_newton_f(topf, qf, A, &iters);
        action = bl_sample_f(topf, qf, A, al, r);
    }
    action = __shfl_sync(FULL, action, 0);
    int id = -1, cseat = 0, cterm = 0;
    for (int k0 = 0; k0 < nc; k0 += 32) {
        const int k = k0 + lane;
        int mid = -1, ms = 0, mt = 0;
        if (k < nc) {
            const int cid = pk_kth(mm, k);
            const bl_node ch = bl_ld_node_hint(t.node + node0 + cid, keep);
            if (ch.relation == action) { mid = cid; ms = ch.seat; mt = ch.terminal; }
        }
        const unsigned hit = __ballot_sync(FULL, mid >= 0);
        if (hit) {
            const int src = __ffs(hit) - 1;
            id = __shfl_sync(FULL, mid, src); cseat = __shfl_sync(FULL, ms, src); cterm = __shfl_sync(FULL, mt, src);
        }
    }
    if (lane == 0) {
        c_iters += iters;
        if (count) { c_evals++; c_children += nc; }
        const bool ends = !(action >= 0 && id >= 0 && !cterm);
        if (ends) c_desc++;
        pk_advance<NCH>(t, sh, s, b, cur, action, id >= 0, id, cseat, cterm);
    }
    __syncwarp();
    return make_uint4(c_evals, c_children, c_iters, c_desc);
}

struct PkVisit { uint4 rec0, rec1; u64 mm; int nc, id0, id1; };

template <int NCH>
__global__ void __launch_bounds__(PK_NT, 1) descend_pk_kernel(bl_tree t, int sim, const bl_half *__restrict__ rands, uint64_t seed, int n_pass, int fuse_expand, unsigned long long *prof) {
    constexpr int PS = 4 * NCH;
    constexpr int NW = (PS + 63) / 64;                  // 64-bit words of the child-position mask
    constexpr int SEG = PS > 64 ? 10 : (PS > 36 ? 8 : (PS > 16 ? 6 : 4));   // SEG*SEG >= PS >= A
    extern __shared__ float4 smem4[];
    const PkShared<NCH> sh(smem4);
    const int A = t.A, T = t.T;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int base = (int)blockIdx.x * PK_EPC;
    const int n_env = t.B - base < PK_EPC ? t.B - base : PK_EPC;
    const bl_qnorm qn(t.qrange + 2 * sim);
    const uint64_t move = t.counters[C_MOVE];
    const uint64_t keep = bl_policy_keep();
    unsigned c_evals = 0, c_children = 0, c_iters = 0, c_desc = 0;

    if (tid < PK_EPC) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(sh.mbar + tid)) : "memory");
        sh.st[tid] = E_NONE;
    }
    if (tid < 8) sh.ready[tid] = 0;
    if (tid == 8) { *sh.n_done = 0; *sh.err = 0; }
    __syncthreads();

    const int V = PK_WARPS - n_pass;                    // service warps: 0 .. V-1 (the pass warps take the high warp ids: the
    if (wid < V) {                                      // arbiter favours them)
        // =========================================== SERVICE ================================================================
        const int stride = V * 32, slane = wid * 32 + lane;
        // every env starts at its root: record, c_puct, first fetch
        for (int s = slane; s < n_env; s += stride) {
            const int b = base + s;
            const bl_node root = bl_ld_node_hint(t.node + (size_t)b * T, keep);
            PkCtl *c = sh.ctl + s;
            c->c_puct = bl_h2f(t.c_puct[b]);
            c->cur = 0; c->parent = 0; c->action = -1; c->leaf = -1; c->seat = root.seat;
            if (root.terminal) {                         // a terminal root ends the descent at once (leaf = 0, no action)
                c->leaf = 0;
                t.leaf[b] = 0; t.leaf_parent[b] = 0; t.leaf_action[b] = -1;
                c_desc++;
                __threadfence_block();
                pk_st_vol(sh.st + s, E_DONE);
                atomicAdd(sh.n_done, 1);
            } else {
                pk_prefetch<NCH>(t, b, 0, smem_u32(sh.ps + (size_t)s * PS), smem_u32(sh.pg + (size_t)s * PS), smem_u32(sh.mbar + s));
                __threadfence_block();
                pk_st_vol(sh.st + s, E_FETCH);
            }
        }
        uint32_t parbits = 0;                            // bit i: phase parity of this lane's i-th env
        int idle = 0;
        long long pf_t0 = prof ? clock64() : 0, pf_visit = 0;
        unsigned pf_polls = 0, pf_visits = 0, pf_slow = 0;

        auto tally = [&](const uint4 &c) { c_evals += c.x; c_children += c.y; c_iters += c.z; c_desc += c.w; };
        auto visit_issue = [&](int s) {
            PkVisit v;
            const size_t node0 = (size_t)(base + s) * T;
            v.mm = *reinterpret_cast<const volatile u64 *>(sh.pg + (size_t)s * PS + 4);
            v.nc = __popcll(v.mm);
            v.id0 = v.id1 = -1;
            v.rec0 = v.rec1 = make_uint4(0u, 0u, 0u, 0u);
            if (v.nc <= PK_KS) {
                if (lane < v.nc) { v.id0 = pk_kth(v.mm, lane); v.rec0 = bl_ld16_hint(t.node + node0 + v.id0, keep); }
                if (lane + 32 < v.nc) { v.id1 = pk_kth(v.mm, lane + 32); v.rec1 = bl_ld16_hint(t.node + node0 + v.id1, keep); }
            }
            return v;
        };
        auto visit_finish = [&](int s, const PkVisit &v) {
            const int b = base + s;
            PkCtl *c = sh.ctl + s;
            const int cur = c->cur, seat = c->seat;
            const float c_puct = c->c_puct;
            float *ps = sh.ps + (size_t)s * PS, *pg = sh.pg + (size_t)s * PS;
            float r;
            if (rands) r = bl_h2f(rands[(size_t)b * T + cur]);
            else r = bl_uniform_half_grid(bl_philox(seed ^ (move * 0x9E3779B97F4A7C15ull), (uint64_t)b, ((uint64_t)sim << 32) | (uint32_t)cur).x);
            if (lane == 0) c->r = r;
            if (v.nc > PK_KS) {                          // more children than entry slots: exact serial path
                __syncwarp();
                tally(pk_slow<NCH>(t, sh, s, b, qn, keep, true));
                return;
            }
            bl_aux ax;
            { union { float4 f; bl_aux a; } x; x.f = *reinterpret_cast<const float4 *>(pg); ax = x.a; }
            union { uint4 u; bl_node n; } x0, x1;
            x0.u = v.rec0; x1.u = v.rec1;
            const bool h0 = v.id0 >= 0, h1 = v.id1 >= 0;
            const int a0 = h0 ? x0.n.relation : 0, a1 = h1 ? x1.n.relation : 0;
            const float q0 = h0 ? qn.fast(seat ? x0.n.w[1] : x0.n.w[0], x0.n.n) : 0.f;
            const float q1 = h1 ? qn.fast(seat ? x1.n.w[1] : x1.n.w[0], x1.n.n) : 0.f;
            int N = (h0 ? (int)x0.n.n : 0) + (h1 ? (int)x1.n.n : 0);
            N = __reduce_add_sync(FULL, N) + A - v.nc;          // every child-less action counts 1 (cuda.cu:91)
            uint32_t cw[4];
#pragma unroll
            for (int w = 0; w < 4; w++) {
                uint32_t mine = 0;
                if (h0 && (a0 >> 5) == w) mine |= 1u << (a0 & 31);
                if (h1 && (a1 >> 5) == w) mine |= 1u << (a1 & 31);
                cw[w] = 2 * NW > w ? __reduce_or_sync(FULL, mine) : 0u;
            }
            const u64 cm0 = (u64)cw[0] | ((u64)cw[1] << 32), cm1 = (u64)cw[2] | ((u64)cw[3] << 32);
            const float lambda = bl_lambda(c_puct, N, A);
            const float top0 = h0 ? __fmul_rn(lambda, ps[a0]) : 0.f, top1 = h1 ? __fmul_rn(lambda, ps[a1]) : 0.f;   // the landed row holds pi
            // alpha seed (newton_search, cuda.cu:44-50): max_a (q[a] + max(lambda*pi[a], 1e-4)); max is order-free
            float am = fmaxf(h0 ? __fadd_rn(q0, fmaxf(top0, 1.e-4f)) : 0.f, h1 ? __fadd_rn(q1, fmaxf(top1, 1.e-4f)) : 0.f);
#pragma unroll
            for (int o = 16; o; o >>= 1) am = fmaxf(am, __shfl_xor_sync(FULL, am, o));
            const float alpha0 = fmaxf(fmaxf(__fmul_rn(lambda, ax.max_pi), 1.e-4f), am);
            auto rank = [&](int a) {
                const u64 below0 = a >= 64 ? ~0ull : ((1ull << a) - 1ull), below1 = a >= 64 ? ((1ull << (a - 64)) - 1ull) : 0ull;
                return __popcll(cm0 & below0) + __popcll(cm1 & below1);
            };
            uint2 *ent = sh.ent + (size_t)s * PK_KS;
            if (h0) ent[rank(a0)] = make_uint2(__float_as_uint(top0), pk_pack_meta(q0, a0, v.id0, x0.n.seat, x0.n.terminal));
            if (h1) ent[rank(a1)] = make_uint2(__float_as_uint(top1), pk_pack_meta(q1, a1, v.id1, x1.n.seat, x1.n.terminal));
            const bool tiny = __fmul_rn(lambda, bl_minnz(ax)) < BL_TINY;
            if (lane == 0) {
                c->alpha0 = alpha0; c->lambda = lambda; c->cm0 = cm0; c->cm1 = cm1; c->nc = (uint8_t)v.nc;
                c->nzpos = (uint32_t)ax.first_nz | ((uint32_t)ax.last_nz << 8);
                c_evals++; c_children += v.nc;
            }
            __syncwarp();
            if (tiny) {
                tally(pk_slow<NCH>(t, sh, s, b, qn, keep, false));
                return;
            }
            __threadfence_block();
            if (lane == 0) {
                pk_st_vol(sh.st + s, E_READY);
                atomicOr(sh.ready + (s & 7), 1u << (s >> 3));
            }
        };

        while (true) {
            if (*reinterpret_cast<volatile int *>(sh.n_done) >= n_env) break;
            bool worked = false;
            int i = 0;
            for (int s0 = wid * 32; s0 < n_env; s0 += stride, i++) {
                const int s = s0 + lane;
                const uint32_t stv = s < n_env ? pk_ld_vol(sh.st + s) : (uint32_t)E_NONE;
                bool rdy = false;
                if (stv == E_FETCH) rdy = pk_mbar_test(smem_u32(sh.mbar + s), (parbits >> i) & 1u);
                if (rdy) { parbits ^= 1u << i; pk_st_vol(sh.st + s, E_VISITING); }
                unsigned rm = __ballot_sync(FULL, rdy), sm = __ballot_sync(FULL, stv == E_SLOW);
                pf_polls++;
                while (rm) {
                    const long long tv = prof ? clock64() : 0;
                    const int l0 = __ffs(rm) - 1; rm &= rm - 1;
                    int l1 = -1;
                    if (rm) { l1 = __ffs(rm) - 1; rm &= rm - 1; }
                    const PkVisit v0 = visit_issue(s0 + l0);
                    PkVisit v1;
                    if (l1 >= 0) v1 = visit_issue(s0 + l1);
                    visit_finish(s0 + l0, v0);
                    if (l1 >= 0) visit_finish(s0 + l1, v1);
                    worked = true;
                    pf_visits += l1 >= 0 ? 2 : 1;
                    if (prof) pf_visit += clock64() - tv;
                }
                while (sm) {
                    const int l0 = __ffs(sm) - 1; sm &= sm - 1;
                    if (lane == l0) pk_st_vol(sh.st + s, E_VISITING);
                    __syncwarp();
                    tally(pk_slow<NCH>(t, sh, s0 + l0, base + s0 + l0, qn, keep, false));
                    worked = true; pf_slow++;
                }
            }
            if (worked) idle = 0;
            else {
                __nanosleep(100);
                if (++idle > PK_IDLE_LIMIT) { if (lane == 0) atomicAdd(sh.err, 1); break; }
            }
        }
        if (prof && lane == 0) {
            atomicAdd(prof + 8, 1ull); atomicAdd(prof + 9, (unsigned long long)pf_polls); atomicAdd(prof + 10, (unsigned long long)pf_visits);
            atomicAdd(prof + 11, (unsigned long long)pf_visit); atomicAdd(prof + 12, (unsigned long long)pf_slow);
            atomicAdd(prof + 13, (unsigned long long)(clock64() - pf_t0));
        }
    } else {
        // ============================================= PASS ================================================================
        u64 tp[2 * NCH];                                  // lambda*pi of the claimed env's node, element pairs
        u64 cm[NW];
        int e = -1, b = -1, cur = 0, state = ST_IDLE, nc = 0, it = 0;
        float alpha = 1.f, error = 0.f, r = 0.f;
        uint32_t nzpos = 0, ps_addr = 0, pg_addr = 0;
        float *ps = sh.ps, *pg = sh.pg;
        const uint2 *ent = sh.ent;
        int rot = (wid * 5) % 28, idle = 0;
        long long pf_t0 = prof ? clock64() : 0, pf_claim = 0, pf_pass = 0, pf_tail = 0, pf_last = pf_t0;
        unsigned pf_trips = 0, pf_idle = 0, pf_lanes = 0, pf_claims = 0;
#define PK_TICK(acc) do { if (prof) { const long long now_ = clock64(); acc += now_ - pf_last; pf_last = now_; } } while (0)
#pragma unroll
        for (int w = 0; w < NW; w++) cm[w] = 0;
#pragma unroll
        for (int c = 0; c < 2 * NCH; c++) tp[c] = 0;

        while (true) {
            if (*reinterpret_cast<volatile int *>(sh.n_done) >= n_env) break;
            // ---- claim a ready env (lanes take envs of their own residue mod 8: rows stay conflict-free) ----------------------
            const bool need = e < 0;
            const unsigned needm = __ballot_sync(FULL, need);
            bool claimed = false;
            if (needm) {
                if (need) {
                    const int r8 = lane & 7;
                    const uint32_t m = pk_ld_vol(sh.ready + r8);
                    const int k = __popc(needm & (0x01010101u << r8) & ((1u << lane) - 1u));
                    uint32_t mr = ((m >> rot) | (m << (28 - rot))) & 0x0FFFFFFFu;     // rotating start: no env starves
                    for (int i = 0; i < k; i++) mr &= mr - 1;
                    if (mr) {
                        int j = __ffs(mr) - 1 + rot;
                        if (j >= 28) j -= 28;
                        const uint32_t bit = 1u << j;
                        const uint32_t old = atomicAnd(sh.ready + r8, ~bit);
                        if (old & bit) { e = 8 * j + r8; claimed = true; }
                    }
                }
                rot = rot + 5 >= 28 ? rot + 5 - 28 : rot + 5;
            }
            if (__any_sync(FULL, claimed)) {
                if (claimed) {
                    __threadfence_block();
                    b = base + e;
                    ps = sh.ps + (size_t)e * PS; pg = sh.pg + (size_t)e * PS; ent = sh.ent + (size_t)e * PK_KS;
                    ps_addr = smem_u32(ps); pg_addr = smem_u32(pg);
                    const PkCtl *c = sh.ctl + e;
                    const float4 c0 = *reinterpret_cast<const float4 *>(c);
                    const ulonglong2 c1 = *reinterpret_cast<const ulonglong2 *>(reinterpret_cast<const uint8_t *>(c) + 16);
                    const uint4 c2 = *reinterpret_cast<const uint4 *>(reinterpret_cast<const uint8_t *>(c) + 32);
                    alpha = c0.x; r = c0.z;
                    const u64 lam2 = pk(c0.y, c0.y);
                    cm[0] = c1.x;
                    if (NW > 1) cm[NW - 1] = c1.y;
                    cur = (int)(int16_t)(c2.x & 0xFFFFu);
                    nc = (int)(c2.z & 0xFFu);
                    nzpos = c2.w;
                    const float4 *ps4 = reinterpret_cast<const float4 *>(ps);
                    const int nrow4 = t.AP >> 2;
#pragma unroll
                    for (int c = 0; c < NCH; c++) {               // top = lambda*pi, from the landed row
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (c < nrow4) v = ps4[c];
                        tp[2 * c] = mul2(pk(v.x, v.y), lam2); tp[2 * c + 1] = mul2(pk(v.z, v.w), lam2);
                    }
                    it = 0; error = BL_INF; state = ST_PASS;
                }
            }
            pf_claims += __popc(__ballot_sync(FULL, claimed));
            if (!__any_sync(FULL, e >= 0)) {
                pf_idle++;
                if (prof) pf_last = clock64();
                __nanosleep(64);
                if (++idle > PK_IDLE_LIMIT) { if (lane == 0) atomicAdd(sh.err, 1); break; }
                continue;
            }
            idle = 0;
            pf_trips++; pf_lanes += __popc(__ballot_sync(FULL, e >= 0));
            PK_TICK(pf_claim);
            // ---- child terms of this pass (full divisions), parked at their positions ----------------------------------------
            if (state == ST_PASS || state == ST_FINAL) {
                bool bad = false;
                for (int i = 0; i < nc; i += 2) {
                    float top[2], q[2], bot[2], sv[2], gv[2];
                    int a[2];
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        const uint2 v = ent[i + u < nc ? i + u : nc - 1];
                        top[u] = __uint_as_float(v.x); q[u] = bl_h2f((bl_half)(v.y & 0xFFFFu)); a[u] = (v.y >> 16) & 127;
                    }
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        bot[u] = __fsub_rn(alpha, q[u]);
                        const float bb = __fmul_rn(bot[u], bot[u]);
                        sv[u] = bl_div_fast(top[u], bot[u]);
                        gv[u] = bl_div_fast(-top[u], bb);
                    }
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        ps[a[u]] = sv[u];
                        pg[a[u]] = gv[u];
                        bad |= !(bot[u] >= 8.67e-19f && bot[u] <= 1.15e18f) || !(sv[u] >= 0.f && sv[u] <= 3.0e38f);
                    }
                }
                if (bad) {                                 // exact serial path on a service warp
                    c_iters += it;
                    __threadfence_block();
                    pk_st_vol(sh.st + e, E_SLOW);
                    e = -1; state = ST_IDLE;
                }
            }
            asm volatile("" ::: "memory");
            // ---- one Newton pass: the two sequential sums (descend.cu, block E) -----------------------------------------------
            const bool pass = state == ST_PASS || state == ST_FINAL;
            float accS = 0.f, accG = 0.f;
            {
                const float bS = alpha, bG = __fmul_rn(alpha, alpha);
                const float yS = bl_rcp_fast(bS), yG = -bl_rcp_fast(bG);
                const u64 yS2 = pk(yS, yS), yG2 = pk(yG, yG), nbS2 = pk(-bS, -bS), bG2 = pk(bG, bG);
                float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f, g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f;
#pragma unroll
                for (int c = 0; c < NCH; c++) {
                    const uint32_t kids = pass ? (uint32_t)(cm[(4 * c) >> 6] >> ((4 * c) & 63)) & 15u : 0u;
                    asm volatile(
                        "{\n.reg .pred p;\nsetp.ne.u32 p, %8, 0;\n"
                        "@p ld.shared.v4.f32 {%0,%1,%2,%3}, [%9];\n"
                        "@p ld.shared.v4.f32 {%4,%5,%6,%7}, [%10];\n}"
                        : "+f"(p0), "+f"(p1), "+f"(p2), "+f"(p3), "+f"(g0), "+f"(g1), "+f"(g2), "+f"(g3)
                        : "r"(kids), "r"(ps_addr + 16u * c), "r"(pg_addr + 16u * c));
                    const u64 t01 = tp[2 * c], t23 = tp[2 * c + 1];
                    u64 q = mul2(t01, yS2), rr = fma2(nbS2, q, t01);
                    const u64 s01 = fma2(rr, yS2, q);
                    q = mul2(t01, yG2); rr = fma2(bG2, q, t01);
                    const u64 h01 = fma2(rr, yG2, q);
                    q = mul2(t23, yS2); rr = fma2(nbS2, q, t23);
                    const u64 s23 = fma2(rr, yS2, q);
                    q = mul2(t23, yG2); rr = fma2(bG2, q, t23);
                    const u64 h23 = fma2(rr, yG2, q);
                    const bool c0 = kids & 1u, c1 = kids & 2u, c2 = kids & 4u, c3 = kids & 8u;
                    accS = __fadd_rn(accS, c0 ? p0 : lo(s01)); accG = __fadd_rn(accG, c0 ? g0 : lo(h01)); const float o0 = accS;
                    accS = __fadd_rn(accS, c1 ? p1 : hi(s01)); accG = __fadd_rn(accG, c1 ? g1 : hi(h01)); const float o1 = accS;
                    accS = __fadd_rn(accS, c2 ? p2 : lo(s23)); accG = __fadd_rn(accG, c2 ? g2 : lo(h23)); const float o2 = accS;
                    accS = __fadd_rn(accS, c3 ? p3 : hi(s23)); accG = __fadd_rn(accG, c3 ? g3 : hi(h23)); const float o3 = accS;
                    if (pass) asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(ps_addr + 16u * c), "f"(o0), "f"(o1), "f"(o2), "f"(o3) : "memory");
                }
            }
            // ---- Newton update (newton_search, cuda.cu:57-66) -----------------------------------------------------------------
            if (pass) {
                if (state == ST_PASS) {
                    it++;
                    const float ne = __fsub_rn(accS, 1.f);
                    if ((ne < 1e-3f) || (error == ne)) state = ST_SAMPLE;
                    else {
                        alpha = __fsub_rn(alpha, __fdiv_rn(ne, accG));
                        error = ne;
                        if (it == 100) state = ST_FINAL;        // loop bound hit: one more pass with the last alpha, no test
                    }
                } else state = ST_SAMPLE;
            }
            PK_TICK(pf_pass);
            // ---- converged lanes: inverse-CDF search over the running sums (cuda.cu:160-176), step to the child, drop the env ----
            if (__any_sync(FULL, state == ST_SAMPLE)) {
                if (state == ST_SAMPLE) {
                    c_iters += it;
                    int c1 = 0;
#pragma unroll
                    for (int j = 0; j < SEG; j++) {
                        const int en = (j + 1) * SEG < A ? (j + 1) * SEG : A;
                        c1 += (j * SEG < A && ps[en - 1] < r) ? 1 : 0;
                    }
                    int l = c1 * SEG;
                    if (l < A) {
                        int c2 = 0;
#pragma unroll
                        for (int j = 0; j < SEG - 1; j++) c2 += (l + j < A && ps[l + j] < r) ? 1 : 0;
                        l += c2;
                    } else l = A;
                    const int first_nz = nzpos & 255, last_nz = (nzpos >> 8) & 255;
                    const int action = first_nz == 255 ? -1 : (l < A ? (r <= 0.f ? first_nz : l) : last_nz);
                    bool has = false;
                    int id = -1, cseat = 0, cterm = 0;
                    if (action >= 0) {
                        u64 word = cm[0], below = 0;
                        int before = 0;
                        if (NW > 1 && action >= 64) { word = cm[NW - 1]; before = __popcll(cm[0]); }
                        below = word & ((1ull << (action & 63)) - 1ull);
                        has = (word >> (action & 63)) & 1ull;
                        if (has) {
                            const uint2 v = ent[before + __popcll(below)];
                            id = (v.y >> 23) & 63; cseat = (v.y >> 29) & 1; cterm = (v.y >> 30) & 3;
                        }
                    }
                    if (!(has && !cterm)) c_desc++;
                    pk_advance<NCH>(t, sh, e, b, cur, action, has, id, cseat, cterm);
                    e = -1; state = ST_IDLE;
                }
            }
            PK_TICK(pf_tail);
        }
        if (prof && lane == 0) {
            atomicAdd(prof + 0, 1ull); atomicAdd(prof + 1, (unsigned long long)pf_trips); atomicAdd(prof + 2, (unsigned long long)pf_idle);
            atomicAdd(prof + 3, (unsigned long long)pf_lanes); atomicAdd(prof + 4, (unsigned long long)(clock64() - pf_t0));
            atomicAdd(prof + 5, (unsigned long long)pf_pass); atomicAdd(prof + 6, (unsigned long long)pf_claim);
            atomicAdd(prof + 7, (unsigned long long)pf_tail); atomicAdd(prof + 14, (unsigned long long)pf_claims);
        }
#undef PK_TICK
    }
    __syncthreads();
    // ---- expand + env step of the CTA's envs (one lane per env; the rows are free now) ------------------------------------------
    if (fuse_expand && tid < n_env) {
        const PkCtl *c = sh.ctl + tid;
        bl_expand_one(t, sim, base + tid, c->leaf, c->parent, c->action, reinterpret_cast<uint32_t *>(sh.ps + (size_t)tid * PS),
                      reinterpret_cast<uint8_t *>(sh.pg + (size_t)tid * PS));
    }
    if (tid == 0 && *sh.err) atomicAdd(reinterpret_cast<unsigned long long *>(t.counters + C_ERRORS), 1ull);
    bl_count(t.counters, C_EVALS, c_evals);
    bl_count(t.counters, C_CHILDREN, c_children);
    bl_count(t.counters, C_ITERS, c_iters);
    bl_count(t.counters, C_DESCENTS, c_desc);
}

int g_pk_pass = 4;        // pass warps of the CTA's 12 (BL_PK_PASS overrides for tuning runs)

template <int NCH>
int launch_pk(const bl_tree *t, int sim, const bl_half *rands, uint64_t seed, cudaStream_t st) {
    const size_t smem = PkCfg<NCH>::SMEM;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(descend_pk_kernel<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr = true;
    }
    const int grid = (t->B + PK_EPC - 1) / PK_EPC;
    const bool fused = t->BP <= 4 * 4 * NCH;
    descend_pk_kernel<NCH><<<grid, PK_NT, smem, st>>>(*t, sim, rands, seed, g_pk_pass, fused ? 1 : 0, bl_phase_prof());
    if (cudaError_t e = cudaGetLastError()) return (int)e;
    return fused ? 0 : bl_expand_step(t, sim, st);
}

}  // namespace

int bl_descend_pk(const bl_tree *t, int sim, const bl_half *rands, uint64_t seed, cudaStream_t st) {
    static bool env_read = false;
    if (!env_read) {
        env_read = true;
        if (const char *e = getenv("BL_PK_PASS")) { const int p = atoi(e); if (p >= 1 && p <= PK_WARPS - 1) g_pk_pass = p; }
    }
    if (t->T > 64 || t->A > 84 || t->A > 127) return -2;
    const int nch = (t->A + 3) / 4;
    if (nch <= 7) return launch_pk<7>(t, sim, rands, seed, st);
    if (nch <= 13) return launch_pk<13>(t, sim, rands, seed, st);
    return launch_pk<21>(t, sim, rands, seed, st);
}
