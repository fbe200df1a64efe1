// Task-parallel tree descent for sm_100a — the hot loop of MCTS.simulate (descend_kernel + policy + newton_search,
// boardlaw/mcts/cpp/cuda.cu:35-99,138-182), restructured so that no lane waits on another env's search depth or
// Newton iteration count, and so that the regular (child-less) actions never go through a full IEEE division.
//
// Work decomposition
//   * A warp owns ES = 16 env slots; slot e is served by the lane pair (2e, 2e+1): lane 2e accumulates the S chain
//     (sum of lambda*pi/(alpha-q)), lane 2e+1 the g chain (sum of -lambda*pi/(alpha-q)^2).  The two sequential fp32
//     sums of newton_search are independent, so the pair runs them side by side, in the reference's order a = 0..A-1.
//   * Every trip round the main loop ("super-step") advances each slot by ONE Newton pass, whatever node or iteration
//     it is at.  A slot that converges samples its action (from the prefix sums of that very pass, which are exactly
//     the running `total` of the reference's sampling loop) and moves to the child; a slot whose descent is over takes
//     the next env from a global queue.  Divergence between envs therefore costs nothing but the tail.
//   * Row loads (pi row -> lambda*pi in shared memory, and the max that seeds alpha) and the inverse-CDF search are
//     warp-cooperative across a: coalesced 128-byte loads, order-independent reductions only.
//
// Arithmetic (bit-exact contract, see mcts_core.cuh)
//   For a child-less action q[a] = 0, so bot = alpha - 0 = alpha for every such a: the divisor is shared.  With
//   y = RN(1/b) (correctly rounded, __frcp_rn), q0 = RN(n*y), r = n - b*q0 (exact in one FMA), RN(q0 + r*y) is the
//   correctly rounded quotient n/b (Markstein's theorem) provided nothing under/overflows — guaranteed here because
//   rows holding a nonzero lambda*pi below 2^-100 are routed to the exact serial fallback, and alpha lies in
//   [1e-4, ~2].  Three dependent FMA-pipe instructions replace a ~10-instruction IEEE division with a subroutine
//   call; tests/test_gpu_mcts.py::test_shared_reciprocal_division checks it against __fdiv_rn on 2^32 operand pairs.
//   Actions that do have a child (bot != alpha; 1.8 per node on average) take full __fdiv_rn divisions, computed
//   before the pass and picked up at their position in the sequence.
//
// Compiled with -fmad=false -prec-div=true -ftz=false (see build.py); fused operations are explicit __fmaf_rn.
#include "engine_internal.cuh"
#include "hex_core.cuh"
#include "mcts_core.cuh"

namespace {

constexpr int ES = 16;            // env slots per warp
constexpr int DW = 4;             // warps per CTA
constexpr unsigned FULL = 0xffffffffu;
#define BL_TINY 7.888609052210118e-31f   /* 2^-100 */

enum { ST_IDLE = 0, ST_VISIT = 1, ST_PASS = 2, ST_FINAL = 3, ST_SAMPLE = 4, ST_SLOW = 5, ST_ADVANCE = 6 };

struct ChildEntry { int16_t a, id; float q; };

__global__ void __launch_bounds__(DW * 32) descend_v2_kernel(bl_tree t, int sim, const bl_half *__restrict__ rands,
                                                             uint64_t seed, ChildEntry *__restrict__ clists, int cap) {
    extern __shared__ float smf[];
    const int A = t.A, T = t.T, Sn = t.Sn;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int e = lane >> 1, chain = lane & 1;
    float *base = smf + (size_t)warp * 3 * ES * A;
    float *top = base + e * A;                        // lambda*pi of the slot's current node; negated where a child exists
    float *pre = base + (size_t)ES * A + e * A;       // S chain: child terms in, running sums (= sampling totals) out
    float *gt = base + (size_t)2 * ES * A + e * A;    // g chain: child terms in, running sums out
    float *mine = chain ? gt : pre;
    ChildEntry *cl = clists + ((size_t)(blockIdx.x * DW + warp) * ES + e) * cap;
    const bl_qnorm qn(t.qrange + 2 * sim);
    const uint64_t move = t.counters[C_MOVE];
    int *queue = reinterpret_cast<int *>(t.counters + C_QUEUE);

    int b = -1, cur = -1, parent = 0, action = -1, state = ST_VISIT, nc = 0, it = 0;
    float alpha = 0.f, error = 0.f, r = 0.f, lambda = 0.f;
    int qlo = 0, qhi = 0;
    bool exhausted = false;                           // warp-uniform
    unsigned c_evals = 0, c_children = 0, c_iters = 0, c_desc = 0;

    while (true) {
        // ---- 0a: finished descents write their result and ask for the next env ----------------------------------
        bool done = false;
        if (state == ST_VISIT) {
            bool evaluable = b >= 0 && cur >= 0 && !t.terminal[(size_t)b * T + cur];
            if (!evaluable) {
                done = true;
                if (b >= 0 && chain == 0) {
                    t.leaf[b] = (int16_t)cur;                  // existing terminal child, or -1: expand_step decides
                    t.leaf_parent[b] = (int16_t)parent;
                    t.leaf_action[b] = (int16_t)action;
                    c_desc++;
                }
            }
        }
        const unsigned req = __ballot_sync(FULL, done && chain == 0);
        if (req) {
            const int nreq = __popc(req);
            const int rank = __popc(req & ((1u << (lane & ~1)) - 1u));
            int assigned = -1, served = 0;
            while (served < nreq) {
                if (qlo == qhi) {
                    if (exhausted) break;
                    int got = 0;
                    if (lane == 0) got = atomicAdd(queue, ES);
                    got = __shfl_sync(FULL, got, 0);
                    if (got >= t.B) { exhausted = true; break; }
                    qlo = got;
                    qhi = min(got + ES, t.B);
                }
                const int take = min(nreq - served, qhi - qlo);
                if (done && rank >= served && rank < served + take) assigned = qlo + (rank - served);
                qlo += take;
                served += take;
            }
            if (done) {
                b = assigned; cur = 0; parent = 0; action = -1;
                if (b < 0) state = ST_IDLE;
            }
        }
        if (__all_sync(FULL, state == ST_IDLE)) break;

        // ---- 0b: child list, N, lambda, random number of the node to evaluate ----------------------------------------
        bool load = false;
        if (state == ST_VISIT && b >= 0 && cur >= 0 && !t.terminal[(size_t)b * T + cur]) {
            load = true;
            const size_t node0 = (size_t)b * T;
            const int seat = t.seats[node0 + cur];
            int N = 0;
            nc = 0;
            for (int c = t.first_child[node0 + cur]; c >= 0; c = t.next_sib[node0 + c]) {
                const int16_t nn = t.n[node0 + c];
                const float qv = qn(t.w[(node0 + c) * Sn + seat], nn);
                if (chain == 0 && nc < cap) cl[nc] = ChildEntry{t.relation[node0 + c], (int16_t)c, qv};
                N += nn;
                nc++;
            }
            N += A - nc;                                        // every child-less action counts 1 (cuda.cu:91)
            lambda = bl_lambda(bl_h2f(t.c_puct[b]), N, A);
            if (rands) r = bl_h2f(rands[node0 + cur]);
            else r = bl_uniform_half_grid(bl_philox(seed ^ (move * 0x9E3779B97F4A7C15ull), (uint64_t)b,
                                                    ((uint64_t)sim << 32) | (uint32_t)cur).x);
        }
        __syncwarp();

        // ---- L: cooperative row load: top = lambda*pi, alpha seed over the child-less part, tiny-value scan ----------
        float alpha0 = 0.f;
        bool tiny = false;
        for (unsigned m = __ballot_sync(FULL, load && chain == 0); m; m &= m - 1) {
            const int ol = __ffs(m) - 1, oe = ol >> 1;
            const int bb = __shfl_sync(FULL, b, ol), tt = __shfl_sync(FULL, cur, ol);
            const float lam = __shfl_sync(FULL, lambda, ol);
            const float *row = t.pi + ((size_t)bb * T + tt) * t.AP;
            float *dst = base + oe * A;
            float cand = 0.f;
            int tn = 0;
            for (int a = lane; a < A; a += 32) {
                const float tp = __fmul_rn(lam, row[a]);
                dst[a] = tp;
                cand = fmaxf(cand, fmaxf(tp, 1.e-4f));          // q = 0: 0 + gap = gap
                tn |= (tp != 0.f && tp < BL_TINY);
            }
            for (int o = 16; o; o >>= 1) cand = fmaxf(cand, __shfl_xor_sync(FULL, cand, o));
            tn = __any_sync(FULL, tn);
            if (e == oe) { alpha0 = cand; tiny = tn; }
        }
        __syncwarp();
        // children: q > 0 raises their candidate; then flag their positions by the sign of top
        if (load) {
            for (int i = 0; i < nc; i++) {
                const ChildEntry ce = cl[i];
                alpha0 = fmaxf(alpha0, __fadd_rn(ce.q, fmaxf(top[ce.a], 1.e-4f)));
            }
        }
        __syncwarp();
        if (load) {
            if (chain == 0)
                for (int i = 0; i < nc; i++) { const int a = cl[i].a; top[a] = -top[a]; }
            alpha = alpha0; it = 0; error = BL_INF;
            state = tiny ? ST_SLOW : ST_PASS;
            if (chain == 0) { c_evals++; c_children += nc; }
        }
        __syncwarp();

        // ---- exact serial fallback (rows with denormal-range lambda*pi): the reference loops verbatim -----------------
        if (state == ST_SLOW) {
            auto topf = [&](int a) { return fabsf(top[a]); };
            auto qf = [&](int a) { float q = 0.f; for (int i = 0; i < nc; i++) if (cl[i].a == a) q = cl[i].q; return q; };
            int iters;
            const float al = bl_newton_f(topf, qf, A, &iters);
            action = bl_sample_f(topf, qf, A, al, r);
            if (chain == 0) c_iters += iters;
            state = ST_ADVANCE;
        }

        // ---- C: terms of the actions that have a child (full divisions), parked at their positions -----------------------
        const bool pass = state == ST_PASS || state == ST_FINAL;
        float bdiv = 1.f, y = 1.f;
        if (pass) {
            bdiv = chain ? __fmul_rn(alpha, alpha) : alpha;
            y = __frcp_rn(bdiv);
            for (int i = 0; i < nc; i++) {
                const ChildEntry ce = cl[i];
                const float tp = -top[ce.a];
                const float bot = __fsub_rn(alpha, ce.q);
                mine[ce.a] = chain ? __fdiv_rn(-tp, __fmul_rn(bot, bot)) : __fdiv_rn(tp, bot);
            }
        }
        // ---- 2: the two sequential sums, one Newton pass ----------------------------------------------------------------------
        float acc = 0.f;
        if (pass) {
#pragma unroll 4
            for (int a = 0; a < A; a++) {
                const float tv = top[a];
                const float num = chain ? -tv : tv;
                const float q0 = __fmul_rn(num, y);
                const float r0 = __fmaf_rn(-bdiv, q0, num);
                const float q1 = __fmaf_rn(r0, y, q0);
                const float term = (__float_as_int(tv) < 0) ? mine[a] : q1;
                acc = __fadd_rn(acc, term);
                mine[a] = acc;
            }
        }
        // ---- 3: Newton update (newton_search, cuda.cu:57-66) ---------------------------------------------------------------
        const float other = __shfl_xor_sync(FULL, acc, 1);
        if (pass) {
            const float S = chain ? other : acc, g = chain ? acc : other;
            if (state == ST_PASS) {
                it++;
                if (chain == 0) c_iters++;
                const float ne = __fsub_rn(S, 1.f);
                if ((ne < 1e-3f) || (error == ne)) state = ST_SAMPLE;
                else {
                    alpha = __fsub_rn(alpha, __fdiv_rn(ne, g));
                    error = ne;
                    if (it == 100) state = ST_FINAL;            // loop bound hit: one more pass with the last alpha, no test
                }
            } else {
                state = ST_SAMPLE;
            }
        }
        __syncwarp();
        // ---- 4: inverse-CDF search over the prefix sums (descend_kernel, cuda.cu:160-176), cooperative -------------------
        for (unsigned m = __ballot_sync(FULL, state == ST_SAMPLE && chain == 0); m; m &= m - 1) {
            const int ol = __ffs(m) - 1, oe = ol >> 1;
            const float rr = __shfl_sync(FULL, r, ol);
            const float *tp_ = base + oe * A, *pr_ = base + (size_t)ES * A + oe * A;
            int first = -1, last = -1;
            for (int a0 = 0; a0 < A; a0 += 32) {
                const int a = a0 + lane;
                const bool in = a < A;
                const float tv = in ? tp_[a] : 0.f;
                const bool pos = in && (tv != 0.f);             // p > 0  <=>  lambda*pi > 0 (no underflow on this path)
                const bool hit = pos && (pr_[a] >= rr);
                const unsigned hm = __ballot_sync(FULL, hit), pm = __ballot_sync(FULL, pos);
                if (pm) last = a0 + 31 - __clz(pm);
                if (hm) { first = a0 + __ffs(hm) - 1; break; }
            }
            const int act = first >= 0 ? first : last;
            if (e == oe) { action = act; state = ST_ADVANCE; }
        }
        // ---- 5: step to the chosen child -------------------------------------------------------------------------------------
        if (state == ST_ADVANCE) {
            parent = cur;
            int next = -1;
            for (int i = 0; i < nc; i++)
                if (cl[i].a == action) next = cl[i].id;
            cur = action >= 0 ? next : -1;
            state = ST_VISIT;
        }
        __syncwarp();
    }
    bl_count(t.counters, C_EVALS, c_evals);
    bl_count(t.counters, C_CHILDREN, c_children);
    bl_count(t.counters, C_ITERS, c_iters);
    bl_count(t.counters, C_DESCENTS, c_desc);
}

// ---- expand + env step (boardlaw/mcts/__init__.py:117-129), one lane per env ---------------------------------------------------
constexpr int XNT = 64;
constexpr int XPITCH = XNT + 4;

__global__ void __launch_bounds__(XNT) expand_step_kernel(bl_tree t, int sim) {
    extern __shared__ __align__(16) uint8_t raw[];
    uint8_t *bd = raw, *stk = raw + (size_t)t.A * XPITCH;
    const int tid = threadIdx.x, lane = tid & 31, wbase = tid & ~31;
    const int A = t.A, T = t.T, Sn = t.Sn;
    const int b = blockIdx.x * XNT + tid, bw = blockIdx.x * XNT + wbase;
    const bool in_range = b < t.B;
    const size_t node0 = (size_t)(in_range ? b : 0) * T;
    int leaf = -1, parent = 0, action = -1;
    bool ok = false;
    if (in_range) {
        leaf = t.leaf[b]; parent = t.leaf_parent[b]; action = t.leaf_action[b];
        ok = action >= 0;
        if (ok) {
            if (leaf < 0) {                                     // new node in slot `sim`
                leaf = sim;
                t.parents[node0 + sim] = (int16_t)parent;
                t.relation[node0 + sim] = (int16_t)action;
                t.next_sib[node0 + sim] = t.first_child[node0 + parent];
                t.first_child[node0 + parent] = (int16_t)sim;
            }                                                   // else: stopped at an existing terminal child, reuse its slot
        } else {
            leaf = -1;
            atomicAdd(reinterpret_cast<unsigned long long *>(t.counters + C_ERRORS), 1ull);
        }
        t.leaf[b] = (int16_t)leaf;
    }
    const unsigned omask = __ballot_sync(FULL, ok);
    for (unsigned m = omask; m; m &= m - 1) {
        const int l = __ffs(m) - 1;
        const int pl = __shfl_sync(FULL, parent, l);
        const uint8_t *row = t.board + ((size_t)(bw + l) * T + pl) * t.BP;
        for (int c = lane; c < A; c += 32) bd[c * XPITCH + wbase + l] = row[c];
    }
    __syncwarp();
    if (ok) {
        const int seat = t.seats[node0 + parent];
        const int win = bl_hex_place<uint8_t>(bd + tid, stk + tid, XPITCH, t.S, seat, action);
        const float r0 = win == 1 ? 1.f : (win == 2 ? -1.f : 0.f), r1 = win == 1 ? -1.f : (win == 2 ? 1.f : 0.f);
        t.rewards[(node0 + leaf) * Sn + 0] = bl_f2h(r0);
        t.rewards[(node0 + leaf) * Sn + 1] = bl_f2h(r1);
        t.terminal[node0 + leaf] = win != 0;
        t.seats[node0 + leaf] = win ? 0 : (uint8_t)(1 - seat);
        if (win)
            for (int c = 0; c < A; c++) bd[c * XPITCH + tid] = 0;      // auto-reset (hex/__init__.py:185-188)
    }
    __syncwarp();
    for (unsigned m = omask; m; m &= m - 1) {
        const int l = __ffs(m) - 1;
        const int ll = __shfl_sync(FULL, leaf, l);
        uint8_t *row = t.board + ((size_t)(bw + l) * T + ll) * t.BP;
        for (int c = lane; c < A; c += 32) row[c] = bd[c * XPITCH + wbase + l];
    }
}

// ---- self test of the shared-reciprocal division -----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) divtest_kernel(uint64_t seed, int n_div, int n_num, unsigned long long *mismatch) {
    // divisors: alpha-like values and their squares, including edge significands; numerators: lambda*pi-like values
    unsigned long long bad = 0;
    for (int i = blockIdx.x; i < n_div; i += gridDim.x) {
        bl_philox_out o = bl_philox(seed, (uint64_t)i, 1);
        unsigned mant = (i & 7) == 0 ? 0x7FFFFFu : ((i & 7) == 1 ? 0u : (o.x & 0x7FFFFFu));
        int ex = 127 - 27 + (int)(o.y % 30);                                   // 2^-27 .. 2^2
        float bdiv = __uint_as_float(((unsigned)ex << 23) | mant);
        float y = __frcp_rn(bdiv);
        for (int j = threadIdx.x; j < n_num; j += blockDim.x) {
            bl_philox_out p = bl_philox(seed + 1, ((uint64_t)i << 32) | (unsigned)j, 2);
            unsigned nm = (j & 15) == 0 ? 0x7FFFFFu : ((j & 15) == 1 ? 0u : (p.x & 0x7FFFFFu));
            int nex = 127 - 100 + (int)(p.y % 98);                             // 2^-100 .. 2^-3
            float num = __uint_as_float(((unsigned)nex << 23) | nm | ((p.z & 1) << 31));
            float q0 = __fmul_rn(num, y), r0 = __fmaf_rn(-bdiv, q0, num), q1 = __fmaf_rn(r0, y, q0);
            bad += (__float_as_uint(q1) != __float_as_uint(__fdiv_rn(num, bdiv)));
        }
    }
    if (bad) atomicAdd(mismatch, bad);
}

}  // namespace

int bl_descend_v2(const bl_tree *t, int sim, const bl_half *rands, uint64_t seed, cudaStream_t st) {
    const int A = t->A;
    const int cap = A < t->T - 1 ? A : t->T - 1;
    const size_t smem = (size_t)DW * 3 * ES * A * sizeof(float);
    if (smem > 227 * 1024) return -2;
    static int occ_cache_A = -1, occ_cache = 0;
    if (occ_cache_A != A) {
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(descend_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
        }
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_cache, descend_v2_kernel, DW * 32, smem);
        if (e != cudaSuccess) return (int)e;
        if (occ_cache < 1) return -2;
        occ_cache_A = A;
    }
    int need = (t->B + DW * ES - 1) / (DW * ES);
    int grid = need < occ_cache * BL_NUM_SMS ? need : occ_cache * BL_NUM_SMS;
    if ((int64_t)grid * DW * ES * cap * (int64_t)sizeof(ChildEntry) > t->scratch_bytes) return -3;
    cudaError_t e = cudaMemsetAsync(t->counters + C_QUEUE, 0, sizeof(uint64_t), st);
    if (e != cudaSuccess) return (int)e;
    descend_v2_kernel<<<grid, DW * 32, smem, st>>>(*t, sim, rands, seed, reinterpret_cast<ChildEntry *>(t->scratch), cap);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    const size_t xsmem = (size_t)2 * A * XPITCH;
    if (xsmem > 48 * 1024) {
        e = cudaFuncSetAttribute(expand_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xsmem);
        if (e != cudaSuccess) return (int)e;
    }
    expand_step_kernel<<<(t->B + XNT - 1) / XNT, XNT, xsmem, st>>>(*t, sim);
    return (int)cudaGetLastError();
}

extern "C" int bl_selftest_division(uint64_t seed, int n_div, int n_num, uint64_t *mismatch, bl_stream stream) {
    divtest_kernel<<<BL_NUM_SMS * 8, 256, 0, bl_cu(stream)>>>(seed, n_div, n_num, reinterpret_cast<unsigned long long *>(mismatch));
    BL_LAUNCH_CHECK();
}
