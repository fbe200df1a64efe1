// Speculative tree descent for sm_100a (variant 6): EVERY node of every tree is evaluated, independently, then the descent is a
// pointer chase.
//
// Why this is legal: what a descent does at node x — the regularised policy at x and the action sampled from it
// (descend_kernel, boardlaw/mcts/cpp/cuda.cu:138-182) — depends on x's own row, its children's statistics, the q range of the
// simulation, c_puct and the random number of NODE x (rands is indexed by node, cuda.cu:158); not on how x was reached.  So the
// action of every node can be computed before anybody descends, and all of them at once.
// Why it is worth 5x the evaluations (sim/2 ~ 32 nodes per tree at c2 against 6.2 on the path): a descent is a chain of dependent
// evaluations, each ~750 dependent scalar instructions and 2-3 memory round trips; with one lane per env (descend_fx.cu) a launch
// lasts as long as the deepest chain's instruction stream at one issue every ~7 cycles, whatever the occupancy.  Independent
// evaluations turn that latency problem into a throughput problem: one lane per (env, node), any number of warps per SM.
//
// Each evaluation is the certified closed-form evaluation of descend_fx.cu (same error model, same exact passes on demand, so
// the same decisions as the reference bit for bit); sampling is a binary search over the node's prefix-sum row `cpi` in global
// memory.  Results go to an 8-byte record per (env, node) in scratch; chase_expand_kernel then follows them from the root, counts
// the path's evaluations for the roofline accounting, and runs the expand + env step.
//
// Compiled with -fmad=false -prec-div=true -ftz=false (see build.py); fused operations are explicit.
#include <cstdio>
#include <cstdlib>

#include "fx_common.cuh"

namespace {

constexpr int AL_WARPS = 4;                            // warps per CTA
constexpr int AL_KS = 17;                              // child entries per lane in shared memory

// result of one node: act = sampled action (255: none), nxt = the child that action leads to (-1: none), flags bit 0 = that child is
// terminal, bit 1 = evaluated; it / nc = Newton passes and children of the evaluation (for the path counters)
struct __align__(8) AlRes { int16_t nxt; uint8_t act, flags, it, nc, pad0, pad1; };

template <int KW>
__global__ void __launch_bounds__(32 * AL_WARPS, 4) eval_all_kernel(bl_tree t, int sim, const bl_half *__restrict__ rands, uint64_t seed,
                                                                   FxEntry *__restrict__ spill, int cap, AlRes *__restrict__ res, long long n_tasks) {
    extern __shared__ __align__(16) uint32_t al_smem[];
    const int A = t.A, T = t.T, AP = t.AP;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int KWT = (T + 63) >> 6;
    // per warp: slots [32][AL_KS] x 16 B, cpr [32][AL_KS] floats, xs / xg [AP] floats
    const uint32_t wbytes = 32u * AL_KS * 16u + 32u * AL_KS * 4u + 2u * 4u * (uint32_t)AP;
    const uint32_t base = smem_u32(al_smem) + (uint32_t)warp * ((wbytes + 15u) & ~15u);
    const uint32_t slot0 = base;
    const uint32_t slot_a = fx_opaque(slot0 + (uint32_t)lane * (AL_KS * 16u));
    const uint32_t cpr_a = fx_opaque(slot0 + 32u * AL_KS * 16u + (uint32_t)lane * (AL_KS * 4u));
    const uint32_t xs_a = slot0 + 32u * AL_KS * 16u + 32u * AL_KS * 4u, xg_a = xs_a + 4u * (uint32_t)AP;
    const long long gwarp = (long long)blockIdx.x * AL_WARPS + warp, nwarps = (long long)gridDim.x * AL_WARPS;
    FxEntry *my_spill = spill + ((size_t)gwarp * 32 + lane) * cap;
    const bl_qnorm qn(t.qrange + 2 * sim);
    const uint64_t move = t.counters[C_MOVE];
    const uint64_t keep = bl_policy_keep();
    unsigned c_fstop = 0, c_fsample = 0, c_fother = 0, c_xpass = 0;

    auto get = [&](int i) {
        FxEntry en;
        if (i < AL_KS) { const uint4 v = fx_lds16(slot_a + 16u * i); en.t = __uint_as_float(v.x); en.q = __uint_as_float(v.y); en.w = __uint_as_float(v.z); en.m = v.w; }
        else en = my_spill[i];
        return en;
    };
    auto put = [&](int i, const FxEntry &en) {
        if (i < AL_KS) fx_sts16(slot_a + 16u * i, make_uint4(__float_as_uint(en.t), __float_as_uint(en.q), __float_as_uint(en.w), en.m));
        else my_spill[i] = en;
    };

    for (long long wt = gwarp; wt * 32 < n_tasks; wt += nwarps) {
        const long long task = wt * 32 + lane;
        const bool has = task < n_tasks;
        const int b = has ? (int)(task / sim) : 0;
        const int node = has ? (int)(task - (long long)b * sim) : 0;
        const size_t node0 = (size_t)b * T, slot = node0 + node;
        // ---- the node: its record (seat, terminal, exists), row summary, children mask -----------------------------------------
        const bl_node me = bl_ld_node_hint(t.node + slot, keep);
        const bool evaluate = has && !me.terminal && (node == 0 || me.parent >= 0);
        union { uint4 u; bl_aux a; } ax;
        ax.u = bl_ld16_hint(t.aux + slot, keep);
        u64 mm[KW];
#pragma unroll
        for (int w = 0; w < KW; w++) mm[w] = (evaluate && w < KWT) ? t.kids[slot * KWT + w] : 0ull;
        const float pa = t.psum[slot];
        const int first_nz = ax.a.first_nz, last_nz = ax.a.last_nz, L1 = last_nz + 1;
        const float *crow = t.cpi + slot * AP;
        const float P = (evaluate && first_nz != 255) ? crow[last_nz] : 0.f;
        const float c_puct = bl_h2f(t.c_puct[b]);
        const int seat = me.seat;
        float r;
        if (rands) r = bl_h2f(rands[slot]);
        else r = bl_uniform_half_grid(bl_philox(seed ^ (move * 0x9E3779B97F4A7C15ull), (uint64_t)b, ((uint64_t)sim << 32) | (uint32_t)node).x);
        int nc = 0;
#pragma unroll
        for (int w = 0; w < KW; w++) nc += __popcll(mm[w]);
        // ---- children: records by cp.async, q, N, lambda, child-less mass, alpha seed ---------------------------------------------
        {
            int k = 0;
#pragma unroll
            for (int w = 0; w < KW; w++)
                for (u64 m = mm[w]; m; m &= m - 1) {
                    const int id = w * 64 + __ffsll((long long)m) - 1;
                    if (k < AL_KS) { fx_cp16(slot_a + 16u * k, t.node + node0 + id); fx_cp4(cpr_a + 4u * k, t.cprior + node0 + id); }
                    k++;
                }
        }
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
        float lambda = 0.f, alpha = 1.f, alpha0 = 1.f, qmax = 0.f, M = 0.f, MW = 0.f, cours = 0.f;
        int mode = FX_EVALDONE, action = -1;
        float e = 0.f, ne_prev = BL_INF, D_prev = 0.f, S = 0.f, Gs = 1.f, ESb = 0.f;
        float s_alpha = 1.f, s_nep = BL_INF, s_Dp = 0.f;
        int it = 0, s_it = 0;
        if (evaluate) {
            int N = 0, k = 0;
#pragma unroll
            for (int w = 0; w < KW; w++)
                for (u64 m = mm[w]; m; m &= m - 1) {
                    const int id = w * 64 + __ffsll((long long)m) - 1;
                    bl_node ch;
                    float pic;
                    if (k < AL_KS) { union { uint4 u; bl_node n; } x; x.u = fx_lds16(slot_a + 16u * k); ch = x.n; pic = fx_lds(cpr_a + 4u * k); }
                    else { ch = bl_ld_node_hint(t.node + node0 + id, keep); pic = t.cprior[node0 + id]; }
                    FxEntry en;
                    en.t = pic; en.q = qn.fast(seat ? ch.w[1] : ch.w[0], ch.n); en.w = (float)(L1 - ch.relation);
                    en.m = ((uint32_t)(ch.seat & 1) << 14) | ((uint32_t)(ch.terminal ? 1u : 0u) << 15) | ((uint32_t)(ch.relation & 255) << 16) | ((uint32_t)id << 24);
                    N += ch.n;
                    put(k, en);
                    k++;
                }
            N += A - nc;                                     // every child-less action counts 1 (cuda.cu:91)
            lambda = bl_lambda(c_puct, N, A);
            double Md = (double)lambda * (double)P, MWd = (double)lambda * ((double)L1 * (double)P - (double)pa);
            alpha = fmaxf(__fmul_rn(lambda, ax.a.max_pi), 1.e-4f);
            for (int i = 0; i < nc; i++) {
                FxEntry en = get(i);
                en.t = __fmul_rn(lambda, en.t);              // t_c = RN(lambda*pi_c), the reference's product
                put(i, en);
                Md -= (double)en.t;
                MWd -= (double)en.t * (double)en.w;
                alpha = fmaxf(alpha, __fadd_rn(en.q, fmaxf(en.t, 1.e-4f)));
                qmax = fmaxf(qmax, en.q);
            }
            M = fmaxf((float)Md, 0.f); MW = fmaxf((float)MWd, 0.f);
            alpha0 = alpha;
            cours = (float)(nc + 13) * FX_U;
            const bool tiny = __fmul_rn(lambda, bl_minnz(ax.a)) < BL_TINY;
            if (first_nz == 255) { action = -1; mode = FX_EVALDONE; }
            else if (tiny) { mode = FX_XALL; D_prev = -1.f; c_fother++; }
            else mode = FX_ITER;
        }
        while (__any_sync(FULL, mode != FX_EVALDONE)) {
            // ---- fast Newton, closed form (see descend_fx.cu for the error model) -------------------------------------------------
            while (__any_sync(FULL, mode == FX_ITER)) {
                if (mode == FX_ITER) {
                    if (e == 0.f) { s_alpha = alpha; s_it = it; s_nep = ne_prev; s_Dp = D_prev; }     // safe point: alpha is the reference's float
                    const float ra = bl_rcp_fast(alpha);
                    float Sn = __fmul_rn(M, ra), Gn = __fmul_rn(Sn, ra), Hn = __fmul_rn(Gn, ra);
                    float ES = __fmul_rn(MW, ra), EG = __fmul_rn(ES, ra);
                    auto term = [&](const float tc, const float q, const float w) {
                        const float rc = bl_rcp_fast(__fsub_rn(alpha, q));
                        const float s = __fmul_rn(tc, rc), g = __fmul_rn(s, rc);
                        Sn = __fadd_rn(Sn, s); Gn = __fadd_rn(Gn, g); Hn = __fadd_rn(Hn, __fmul_rn(g, rc));
                        ES = __fmaf_rn(w, s, ES); EG = __fmaf_rn(w, g, EG);
                    };
                    const int ns = nc < AL_KS ? nc : AL_KS;
                    for (int i = 0; i < ns; i++) {
                        const uint4 v = fx_lds16(slot_a + 16u * i);
                        term(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z));
                    }
                    for (int i = AL_KS; i < nc; i++) { const FxEntry en = my_spill[i]; term(en.t, en.q, en.w); }
                    const float ESn = __fmaf_rn(FX_U * 1.01f, ES, __fmul_rn(2.f * FX_U + cours, Sn));
                    const float EGn = __fmaf_rn(FX_U * 1.01f, EG, __fmul_rn(4.f * FX_U + cours, Gn));
                    const float ne = __fsub_rn(Sn, 1.f);
                    const float Dk = __fmaf_rn(__fmul_rn(Gn, e), 1.05f, __fmaf_rn(2.f * FX_U, fabsf(ne), ESn));
                    S = Sn; Gs = Gn; ESb = ESn;
                    it++;
                    const bool guard_ok = (e <= FX_GUARD * (alpha - qmax)) && (Sn < 3.0e38f) && (Gn > 0.f) && (Gn < 3.0e38f) && it <= FX_MAXIT;
                    if (!guard_ok) { mode = FX_XALL; alpha = alpha0; it = 0; ne_prev = BL_INF; D_prev = -1.f; e = 0.f; c_fother++; }
                    else if (ne < 1e-3f - Dk) mode = FX_SAMPLE;
                    else if (!((ne > 1e-3f + Dk) && (fabsf(ne - ne_prev) > Dk + D_prev))) { mode = FX_XPASS; c_fstop++; }
                    else {
                        const float rG = bl_rcp_fast(Gn);
                        const float L = 2.4f * (fmaxf(ne, 0.f) + 2.f * Gn * e) * Hn * rG * rG;      // 1.2 * F F''/F'^2, F'' = 2 Hn
                        const float R = 1.05f * rG * (Dk + fabsf(ne) * (EGn * rG + 4.f * FX_U));
                        const float eps = L * e + R;
                        const float step = __fdiv_rn(ne, Gn);
                        const float a_new = __fadd_rn(alpha, step);
                        const float err = __fsub_rn(step, __fsub_rn(a_new, alpha));          // alpha + step = a_new + err exactly (|alpha| >= |step|)
                        const uint32_t ab = __float_as_uint(a_new);
                        const float ulp = __uint_as_float(ab & 0x7f800000u) * 1.1920928955078125e-07f;
                        const bool exact = (eps < 0.5f * ulp - fabsf(err)) && (ab & 0x007fffffu) != 0u && fabsf(step) <= fabsf(alpha);
                        e = exact ? 0.f : eps + ulp;
                        alpha = a_new;
                        ne_prev = ne; D_prev = Dk;
                    }
                }
            }
            // ---- certified sampling: binary search over cum(a) = k*cpi[a] + (corrections of the children at or before a) ----------------
            if (__any_sync(FULL, mode == FX_SAMPLE)) {
                if (mode == FX_SAMPLE) {
                    const float ra = bl_rcp_fast(alpha);
                    const float k = __fmul_rn(lambda, ra);
                    for (int i = 0; i < nc; i++) {
                        const FxEntry en = get(i);
                        const float s = __fmul_rn(en.t, bl_rcp_fast(__fsub_rn(alpha, en.q)));
                        const float dlt = __fsub_rn(s, __fmul_rn(en.t, ra));
                        if (i < AL_KS) fx_sts(cpr_a + 4u * i, dlt);
                        else my_spill[i].w = dlt;
                    }
                    const float fL1 = (float)L1;
                    auto cum = [&](int a) {
                        float v = __fmul_rn(k, __ldg(crow + a));
                        const float wa = fL1 - (float)a;     // a child at or before a has w = L1 - a_c >= L1 - a
                        const int ns = nc < AL_KS ? nc : AL_KS;
                        for (int i = 0; i < ns; i++)
                            if (fx_lds(slot_a + 16u * i + 8u) >= wa) v = __fadd_rn(v, fx_lds(cpr_a + 4u * i));
                        for (int i = AL_KS; i < nc; i++) {
                            const FxEntry en = my_spill[i];
                            if (fx_a(en.m) <= a) v = __fadd_rn(v, en.w);
                        }
                        return v;
                    };
                    int lo = 0, hi = A;                      // l = #{a : cum(a) < r}
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (cum(mid) < r) lo = mid + 1; else hi = mid;
                    }
                    const float delta = __fmaf_rn(__fmul_rn(Gs, e), 1.05f, ESb);
                    bool ok = true;
                    if (lo < A) ok = ok && (cum(lo) >= r + delta);
                    if (lo > 0) ok = ok && (cum(lo - 1) < r - delta);
                    for (int i = AL_KS; i < nc; i++) my_spill[i].w = (float)(L1 - fx_a(my_spill[i].m));     // (w held the correction)
                    if (r <= 0.f) { action = first_nz; mode = FX_EVALDONE; }
                    else if (!ok) { mode = FX_XPASS; c_fsample++; }
                    else { action = lo < A ? lo : last_nz; mode = FX_EVALDONE; }
                }
            }
            // ---- exact passes, the warp working on one lane's row at a time (as descend_fx.cu) ----------------------------------------------
            for (unsigned pend = __ballot_sync(FULL, mode == FX_XPASS || mode == FX_XALL); pend; pend &= pend - 1) {
                const int ld = __ffs(pend) - 1;
                const bool leader = lane == ld;
                if (leader && mode == FX_XPASS) { alpha = s_alpha; it = s_it; ne_prev = s_nep; D_prev = s_Dp; }
                const float xalpha = __shfl_sync(FULL, alpha, ld), xlam = __shfl_sync(FULL, lambda, ld), xr = __shfl_sync(FULL, r, ld);
                const int xnc = __shfl_sync(FULL, nc, ld);
                const long long xtask = wt * 32 + ld;
                const int xb = (int)(xtask / sim);
                const size_t xslot = (size_t)xb * T + (size_t)(xtask - (long long)xb * sim);
                const float b2 = __fmul_rn(xalpha, xalpha);
                for (int a = lane; a < AP; a += 32) {
                    const float top = __fmul_rn(xlam, t.pi[xslot * AP + a]);
                    fx_sts(xs_a + 4u * a, __fdiv_rn(top, xalpha));
                    fx_sts(xg_a + 4u * a, __fdiv_rn(-top, b2));
                }
                __syncwarp();
                for (int i = lane; i < xnc; i += 32) {
                    FxEntry en;
                    if (i < AL_KS) { const uint4 v = fx_lds16(slot0 + (uint32_t)ld * (AL_KS * 16u) + 16u * i); en.t = __uint_as_float(v.x); en.q = __uint_as_float(v.y); en.m = v.w; }
                    else en = spill[((size_t)gwarp * 32 + ld) * cap + i];
                    const float bot = __fsub_rn(xalpha, en.q);
                    fx_sts(xs_a + 4u * fx_a(en.m), __fdiv_rn(en.t, bot));
                    fx_sts(xg_a + 4u * fx_a(en.m), __fdiv_rn(-en.t, __fmul_rn(bot, bot)));
                }
                __syncwarp();
                float acc = 0.f;
                if (lane < 2) {
                    const uint32_t arr = lane == 0 ? xs_a : xg_a;
                    for (int c = 0; 4 * c < AP; c++) {        // (pad terms are +-0: they leave the sums as they are)
                        const uint4 u = fx_lds16(arr + 16u * c);
                        const float v0 = __uint_as_float(u.x), v1 = __uint_as_float(u.y), v2 = __uint_as_float(u.z), v3 = __uint_as_float(u.w);
                        const float o0 = __fadd_rn(acc, v0), o1 = __fadd_rn(o0, v1), o2 = __fadd_rn(o1, v2), o3 = __fadd_rn(o2, v3);
                        acc = o3;
                        if (lane == 0)                        // |.| = running sum, sign bit set <=> p == 0
                            fx_sts16(arr + 16u * c, make_uint4(__float_as_uint(v0 > 0.f ? o0 : -o0), __float_as_uint(v1 > 0.f ? o1 : -o1),
                                                               __float_as_uint(v2 > 0.f ? o2 : -o2), __float_as_uint(v3 > 0.f ? o3 : -o3)));
                    }
                }
                const float xS = __shfl_sync(FULL, acc, 0), xg_ = __shfl_sync(FULL, acc, 1);
                __syncwarp();
                bool take = false;                            // the pass ends the Newton loop: sample from its running sums
                if (leader) {
                    c_xpass++;
                    it++;
                    const float ne = __fsub_rn(xS, 1.f);
                    const bool errknown = D_prev <= 0.f;      // the reference's previous S - 1 is known exactly
                    if (it > 100) take = true;                // the extra pass after the loop bound: sums at the final alpha, no test
                    else if (ne < 1e-3f) take = true;
                    else if (errknown && ne_prev == ne) take = true;
                    else if (!errknown && !(fabsf(ne - ne_prev) > D_prev)) {
                        mode = FX_XALL; alpha = alpha0; it = 0; ne_prev = BL_INF; D_prev = -1.f; e = 0.f;
                    } else {
                        alpha = __fsub_rn(alpha, __fdiv_rn(ne, xg_));       // the reference's update, on its own sums
                        ne_prev = ne; e = 0.f;
                        if (mode == FX_XPASS) { mode = FX_ITER; D_prev = 0.f; }
                        else D_prev = -1.f;
                    }
                }
                take = __shfl_sync(FULL, (int)take, ld) != 0;
                if (take) {
                    int hit = 1 << 20, lastv = -1;
                    for (int a = lane; a < A; a += 32) {
                        const float x = fx_lds(xs_a + 4u * a);
                        if (__float_as_int(x) >= 0) {
                            if (x >= xr && a < hit) hit = a;
                            lastv = max(lastv, a);
                        }
                    }
#pragma unroll
                    for (int o = 16; o; o >>= 1) {
                        hit = min(hit, __shfl_xor_sync(FULL, hit, o));
                        lastv = max(lastv, __shfl_xor_sync(FULL, lastv, o));
                    }
                    if (leader) { action = hit < (1 << 20) ? hit : lastv; mode = FX_EVALDONE; if (it > 100) it = 100; }
                }
                __syncwarp();
            }
        }
        // ---- the node's result -----------------------------------------------------------------------------------------------------------
        if (has) {
            AlRes rr;
            rr.nxt = -1; rr.act = 255; rr.flags = 0; rr.it = 0; rr.nc = 0; rr.pad0 = rr.pad1 = 0;
            if (evaluate) {
                rr.flags = 2;
                rr.it = (uint8_t)it; rr.nc = (uint8_t)(nc < 255 ? nc : 255);
                if (action >= 0) {
                    rr.act = (uint8_t)action;
                    for (int i = 0; i < nc; i++) {
                        const uint32_t m = get(i).m;
                        if (fx_a(m) == action) { rr.nxt = (int16_t)(m >> 24); rr.flags |= (uint8_t)((m >> 15) & 1u); }
                    }
                }
            }
            union { AlRes r; uint2 u; } x;
            x.r = rr;
            reinterpret_cast<uint2 *>(res)[slot] = x.u;
        }
        __syncwarp();
    }
    bl_count(t.counters, C_FLAG_STOP, c_fstop);
    bl_count(t.counters, C_FLAG_SAMPLE, c_fsample);
    bl_count(t.counters, C_FLAG_OTHER, c_fother);
    bl_count(t.counters, C_EXACT_PASSES, c_xpass);
}

// the descent itself: follow the nodes' sampled actions from the root; then expand + env step (boardlaw/mcts/__init__.py:117-129)
constexpr int CNT = 128;
__global__ void __launch_bounds__(CNT) chase_expand_kernel(bl_tree t, int sim, const AlRes *__restrict__ res) {
    extern __shared__ __align__(16) uint8_t raw[];
    const int tid = threadIdx.x, pw = (t.BP >> 2) | 1;            // row pitch in words (odd)
    uint32_t *bdw = reinterpret_cast<uint32_t *>(raw) + (size_t)tid * pw;
    uint8_t *stk = raw + (size_t)CNT * pw * 4 + (size_t)tid * pw * 4;
    const int b = blockIdx.x * CNT + tid;
    unsigned c_evals = 0, c_children = 0, c_iters = 0;
    if (b < t.B) {
        const size_t node0 = (size_t)b * t.T;
        int cur = 0, parent = 0, action = -1;
        const bl_node root = bl_ld_node(t.node + node0);
        if (!root.terminal) {
            while (true) {
                union { uint2 u; AlRes r; } x;
                x.u = reinterpret_cast<const uint2 *>(res)[node0 + cur];
                c_evals++; c_children += x.r.nc; c_iters += x.r.it;
                parent = cur;
                action = x.r.act == 255 ? -1 : (int)x.r.act;
                cur = action >= 0 ? (int)x.r.nxt : -1;
                if (cur < 0 || (x.r.flags & 1)) break;            // new leaf / no legal action, or an existing terminal child
            }
        }
        t.leaf[b] = (int16_t)cur;
        t.leaf_parent[b] = (int16_t)parent;
        t.leaf_action[b] = (int16_t)action;
        bl_expand_one(t, sim, b, cur, parent, action, bdw, stk);
    }
    bl_count(t.counters, C_EVALS, c_evals);
    bl_count(t.counters, C_CHILDREN, c_children);
    bl_count(t.counters, C_ITERS, c_iters);
    bl_count(t.counters, C_DESCENTS, b < t.B ? 1u : 0u);
}

int g_al_ctas = 0;         // BL_ALL_CTAS: CTAs per SM of the evaluation kernel (0 = default)

size_t al_warp_bytes(const bl_tree *t) { return ((size_t)32 * AL_KS * 16 + (size_t)32 * AL_KS * 4 + (size_t)2 * 4 * t->AP + 15) & ~(size_t)15; }
int al_grid(const bl_tree *t) {
    const size_t smem = al_warp_bytes(t) * AL_WARPS;
    int per_sm = (int)(220 * 1024 / (smem + 1024));
    if (per_sm > 4) per_sm = 4;
    if (g_al_ctas > 0 && g_al_ctas < per_sm) per_sm = g_al_ctas;
    return (per_sm < 1 ? 1 : per_sm) * BL_NUM_SMS;
}
int64_t al_spill_bytes(const bl_tree *t) {
    const int64_t cap = t->A < t->T - 1 ? t->A : (t->T > 1 ? t->T - 1 : 1);
    return (int64_t)al_grid(t) * AL_WARPS * 32 * cap * (int64_t)sizeof(FxEntry);
}

template <int KW>
int launch_all(const bl_tree *t, int sim, const bl_half *rands, uint64_t seed, cudaStream_t st) {
    const int cap = t->A < t->T - 1 ? t->A : (t->T > 1 ? t->T - 1 : 1);
    const size_t smem = al_warp_bytes(t) * AL_WARPS;
    if (smem > 227 * 1024) return -2;
    const int64_t spill_b = (al_spill_bytes(t) + 255) & ~255ll;
    if (spill_b + (int64_t)t->B * t->T * (int64_t)sizeof(AlRes) > t->scratch_bytes) return -3;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(eval_all_kernel<KW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    AlRes *res = reinterpret_cast<AlRes *>(reinterpret_cast<uint8_t *>(t->scratch) + spill_b);
    const long long n_tasks = (long long)t->B * sim;
    long long need = (n_tasks + 32 * AL_WARPS - 1) / (32 * AL_WARPS);
    const int grid = (int)(need < al_grid(t) ? need : al_grid(t));
    eval_all_kernel<KW><<<grid, 32 * AL_WARPS, smem, st>>>(*t, sim, rands, seed, reinterpret_cast<FxEntry *>(t->scratch), cap, res, n_tasks);
    if (cudaError_t e = cudaGetLastError()) return (int)e;
    const size_t xsmem = (size_t)2 * CNT * ((t->BP >> 2) | 1) * 4;
    if (xsmem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(chase_expand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xsmem);
        if (e != cudaSuccess) return (int)e;
    }
    chase_expand_kernel<<<(t->B + CNT - 1) / CNT, CNT, xsmem, st>>>(*t, sim, res);
    return (int)cudaGetLastError();
}

}  // namespace

int64_t bl_all_scratch_bytes(const bl_tree *t) {
    return ((al_spill_bytes(t) + 255) & ~255ll) + (int64_t)t->B * t->T * (int64_t)sizeof(AlRes);
}

// -2: unsupported shape, -3: scratch too small (the caller falls back to the exact kernels)
int bl_descend_all(const bl_tree *t, int sim, const bl_half *rands, uint64_t seed, cudaStream_t st) {
    static bool env_read = false;
    if (!env_read) {
        env_read = true;
        if (const char *e = getenv("BL_ALL_CTAS")) g_al_ctas = atoi(e);
    }
    if (t->A > 255 || t->T > 256 || !t->cpi) return -2;
    if (t->T <= 64) return launch_all<1>(t, sim, rands, seed, st);
    if (t->T <= 128) return launch_all<2>(t, sim, rands, seed, st);
    return launch_all<4>(t, sim, rands, seed, st);
}

bool bl_experimental_built() { return true; }
