// Certified fast tree descent for sm_100a (variant 5) — the hot loop of MCTS.simulate (descend_kernel + policy + newton_search,
// boardlaw/mcts/cpp/cuda.cu:35-99,138-182) evaluated in O(children) instead of O(actions) per Newton iteration.
//
// What the reference computes at a node: alpha by Newton on S(alpha) = sum_a t_a/(alpha - q_a) - 1 (t_a = lambda*pi_a, q_a = 0
// where there is no child), each pass an fp32 sum over a = 0..A-1, then the first action whose running sum reaches r.  Only two
// kinds of DECISIONS leave the evaluation: "stop / continue" per Newton pass and the sampled action; alpha itself is never
// output.  descend.cu reproduces the sums bit for bit (A dependent additions and 2A divisions per pass: 60 % of a move).  Here
// the sums are evaluated in closed form — every child-less action shares the divisor alpha, so
//     S(alpha) = M/alpha + sum_children t_c/(alpha - q_c),   M = lambda * (sum of pi over child-less actions)
// (c = 1.8 children per node on average at c2) — and every decision is CERTIFIED against a rigorous bound on
// |reference's value - ours|: a decision whose margin exceeds the bound is the reference's decision; one that does not is
// recomputed with the reference's arithmetic (exact path below; ~1 % of evaluations).  Results are therefore identical to
// descend.cu's and the oracle's (tests/test_gpu_mcts.py runs the stepwise oracle comparison for this variant too, and
// tests/test_gpu_fx.py compares whole c2-sized searches against variant 2 and checks the bounds against the exact values).
//
// The bound (oracle/filter_model.py is the executable specification, checked against the reference's arithmetic on the CPU;
// u = 2^-24, all terms >= 0):
//   * the reference's own rounding at a given alpha:  |S_ref - S| <= u * sum over its rounded additions of the partial sum
//     (Higham's running bound; in closed form u*(MW/alpha + sum_c wgt_c s_c) with wgt_a = number of later additions) + 2u S
//     for the rounded quotients; the same for g with 4u;
//   * ours: (children + 8) u relative (a sequential sum of children + 1 terms, <= 4 roundings per term);
//   * alpha: both iterations are perturbed Newton maps, so the distance of the UNROUNDED updates obeys e' <= L e + R with
//     L >= sup N' = F F''/F'^2 and R from the two bounds above; both updates are then rounded to fp32, and when no rounding
//     boundary lies within e' of ours (FastTwoSum gives the exact residual) the reference's alpha IS our float and e' = 0.
//     That matters: |g| e is the sensitivity of every later decision to alpha and |g| ~ 10^3 when one child dominates, so a
//     one-ulp doubt about alpha costs ~10^-4 of margin.
//
// Mapping: ONE lane per env, the 32 envs of a warp in lock step, one policy evaluation per trip.  The Newton bookkeeping is scalar
// work per env (~90 instructions per pass), so lanes are the only way to run it without redundancy (a first version with 8 lanes
// per env, redundant Newton and row slices in registers: 173 M warp instructions per launch, 360 us).  What made
// one lane per env possible: nothing in the fast path touches the A-wide row any more except the sampling step, and that is a
// binary search over the row's PREFIX SUMS (`cpi`, written next to `pi` by the network epilogue; the row travels to the lane's
// shared-memory row by cp.async while the Newton passes run); the row's mass is cpi[last], its addition-count weight comes
// from `psum`, a child's prior from `cprior`.  Per trip: visit (children records by cp.async, q, lambda, seed) -> up to FX_NIT
// Newton passes (lanes that need more continue next trip) -> certified sampling -> exact passes for the lanes that asked, the
// WARP cooperating on one lane's row at a time (terms 3-4 per lane, the two sequential sums on lanes 0 and 1) -> advance.
//
// Compiled with -fmad=false -prec-div=true -ftz=false (see build.py); fused operations are explicit.
#include <cstdio>
#include <cstdlib>

#include "fx_common.cuh"

namespace {

template <int NCH, int KW, int EPW>
__global__ void __launch_bounds__(32, EPW >= 32 ? 8 : (EPW >= 16 ? 16 : 24)) descend_fx_kernel(bl_tree t, int sim, const bl_half *__restrict__ rands, uint64_t seed,
                                                           FxEntry *__restrict__ spill, int cap, int nit, int fuse_expand) {
    constexpr int PS = 4 * NCH;                         // floats per lane row (NCH odd)
    extern __shared__ __align__(16) uint32_t fx_smem[];
    const int A = t.A, T = t.T, AP = t.AP;
    const int lane = threadIdx.x;
    // EPW envs per warp: lanes >= EPW hold no env (they only help with the exact passes).  Fewer envs per warp = more warps per SM:
    // the kernel runs at the latency of its own dependent instruction stream, and a trip costs the MAXIMUM over the warp's lanes
    const int b = lane < EPW ? blockIdx.x * EPW + lane : t.B;
    const int KWT = (T + 63) >> 6;
    // per lane: row [PS] floats (prefix sums of the node's pi row), slots [FX_KS] x 16 B (child records in flight, then entries),
    // cpr [FX_KS] floats (children's priors in flight); per warp: xs, xg [AP] floats (exact pass)
    const uint32_t base = smem_u32(fx_smem);
    const uint32_t row_a = fx_opaque(base + (uint32_t)lane * (PS * 4u));
    const uint32_t slot0 = base + (uint32_t)EPW * PS * 4u;
    const uint32_t slot_a = fx_opaque(slot0 + (uint32_t)lane * (FX_KS * 16u));
    const uint32_t cpr0 = slot0 + (uint32_t)EPW * FX_KS * 16u;
    const uint32_t cpr_a = fx_opaque(cpr0 + (uint32_t)lane * (FX_KS * 4u));
    const uint32_t xs_a = cpr0 + (uint32_t)EPW * FX_KS * 4u, xg_a = xs_a + 4u * (uint32_t)AP;
    FxEntry *my_spill = spill + (size_t)(b < t.B ? b : 0) * cap;
    const bl_qnorm qn(t.qrange + 2 * sim);
    const uint64_t move = t.counters[C_MOVE];
    const uint64_t keep = bl_policy_keep();
    const size_t node0 = (size_t)(b < t.B ? b : 0) * T;

    auto get = [&](int i) {
        FxEntry en;
        if (i < FX_KS) { const uint4 v = fx_lds16(slot_a + 16u * i); en.t = __uint_as_float(v.x); en.q = __uint_as_float(v.y); en.w = __uint_as_float(v.z); en.m = v.w; }
        else en = my_spill[i];
        return en;
    };
    auto put = [&](int i, const FxEntry &en) {
        if (i < FX_KS) fx_sts16(slot_a + 16u * i, make_uint4(__float_as_uint(en.t), __float_as_uint(en.q), __float_as_uint(en.w), en.m));
        else my_spill[i] = en;
    };
    // sampling corrections: cpr zone for the first FX_KS children, the spilled entry's w field beyond (w is recomputed from m when needed)
    auto get_dlt = [&](int i) { return i < FX_KS ? fx_lds(cpr_a + 4u * i) : my_spill[i].w; };

    int cur = 0, parent = 0, action = -1, cur_seat = 0;
    float c_puct = 0.f;
    unsigned c_evals = 0, c_children = 0, c_iters = 0, c_fstop = 0, c_fsample = 0, c_fother = 0, c_xpass = 0;
    int mode = FX_IDLE;
    // the next node's header, fetched when the descent steps to it
    uint4 nx_aux = make_uint4(0u, 0u, 0u, 0u);
    u64 nx_mm[KW];
    float nx_pa = 0.f;
#pragma unroll
    for (int w = 0; w < KW; w++) nx_mm[w] = 0ull;
    auto fetch_node = [&](int n) {
        const size_t slot = node0 + n;
        nx_aux = bl_ld16_hint(t.aux + slot, keep);
#pragma unroll
        for (int w = 0; w < KW; w++) nx_mm[w] = w < KWT ? t.kids[slot * KWT + w] : 0ull;
        nx_pa = t.psum[slot];
        const float4 *row = reinterpret_cast<const float4 *>(t.cpi + slot * AP);
#pragma unroll
        for (int c = 0; c < NCH; c++)
            if (4 * c < AP) fx_cp16(row_a + 16u * c, row + c);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (b < t.B) {
        const bl_node root = bl_ld_node_hint(t.node + node0, keep);
        c_puct = bl_h2f(t.c_puct[b]);
        cur_seat = root.seat;
        if (!root.terminal) { mode = FX_VISIT; fetch_node(0); }       // a terminal root: leaf = 0, no action (the expand step records the error)
    }
    // evaluation state
    float alpha = 1.f, alpha0 = 1.f, e = 0.f, ne_prev = BL_INF, D_prev = 0.f, S = 0.f, Gs = 1.f, ESb = 0.f;
    float s_alpha = 1.f, s_nep = BL_INF, s_Dp = 0.f, M = 0.f, MW = 0.f, lambda = 0.f, qmax = 0.f, cours = 0.f, r = 0.f;
    int it = 0, s_it = 0, nc = 0, L1 = 0, first_nz = 255, last_nz = 0;

    while (__any_sync(FULL, mode != FX_IDLE)) {
        // ---- A. visit: children records by cp.async, q, N, lambda, child-less mass, alpha seed ----------------------------------------
        if (__any_sync(FULL, mode == FX_VISIT)) {
            const bool mine = mode == FX_VISIT;
            const size_t slot = node0 + (mine ? cur : 0);
            union { uint4 u; bl_aux a; } ax;
            ax.u = nx_aux;
            if (mine) {
                nc = 0;
#pragma unroll
                for (int w = 0; w < KW; w++) nc += __popcll(nx_mm[w]);
                int k = 0;
#pragma unroll
                for (int w = 0; w < KW; w++)
                    for (u64 m = nx_mm[w]; m; m &= m - 1) {
                        const int id = w * 64 + __ffsll((long long)m) - 1;
                        if (k < FX_KS) { fx_cp16(slot_a + 16u * k, t.node + node0 + id); fx_cp4(cpr_a + 4u * k, t.cprior + node0 + id); }
                        k++;
                    }
                if (rands) r = bl_h2f(rands[slot]);
                else r = bl_uniform_half_grid(bl_philox(seed ^ (move * 0x9E3779B97F4A7C15ull), (uint64_t)b, ((uint64_t)sim << 32) | (uint32_t)cur).x);
            }
            asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
            if (mine) {
                first_nz = ax.a.first_nz; last_nz = ax.a.last_nz; L1 = last_nz + 1;
                int N = 0, k = 0;
#pragma unroll
                for (int w = 0; w < KW; w++)
                    for (u64 m = nx_mm[w]; m; m &= m - 1) {
                        const int id = w * 64 + __ffsll((long long)m) - 1;
                        bl_node ch;
                        float pic;
                        if (k < FX_KS) { union { uint4 u; bl_node n; } x; x.u = fx_lds16(slot_a + 16u * k); ch = x.n; pic = fx_lds(cpr_a + 4u * k); }
                        else { ch = bl_ld_node_hint(t.node + node0 + id, keep); pic = t.cprior[node0 + id]; }
                        const float q = qn.fast(cur_seat ? ch.w[1] : ch.w[0], ch.n);
                        N += ch.n;
                        FxEntry en;
                        en.t = pic; en.q = q; en.w = (float)(L1 - ch.relation);
                        en.m = ((uint32_t)(ch.seat & 1) << 14) | ((uint32_t)(ch.terminal ? 1u : 0u) << 15) | ((uint32_t)(ch.relation & 255) << 16) | ((uint32_t)id << 24);
                        put(k, en);
                        k++;
                    }
                N += A - nc;                                 // every child-less action counts 1 (cuda.cu:91)
                lambda = bl_lambda(c_puct, N, A);
                // mass P = cpi[last] and addition-count weight W = sum (L1 - a) pi_a = L1*P - sum a*pi_a of the row; child-less parts
                const float P = first_nz == 255 ? 0.f : fx_lds(row_a + 4u * last_nz);
                double Md = (double)lambda * (double)P, MWd = (double)lambda * ((double)L1 * (double)P - (double)nx_pa);
                alpha = fmaxf(__fmul_rn(lambda, ax.a.max_pi), 1.e-4f);
                qmax = 0.f;
                for (int i = 0; i < nc; i++) {
                    FxEntry en = get(i);
                    en.t = __fmul_rn(lambda, en.t);          // t_c = RN(lambda*pi_c), the reference's product
                    put(i, en);
                    Md -= (double)en.t;
                    MWd -= (double)en.t * (double)en.w;
                    alpha = fmaxf(alpha, __fadd_rn(en.q, fmaxf(en.t, 1.e-4f)));
                    qmax = fmaxf(qmax, en.q);
                }
                M = fmaxf((float)Md, 0.f); MW = fmaxf((float)MWd, 0.f);
                alpha0 = alpha;
                cours = (float)(nc + 13) * FX_U;             // our roundings: sequential sum of nc + 1 terms, <= 4 per term; P, W, sampling
                const bool tiny = __fmul_rn(lambda, bl_minnz(ax.a)) < BL_TINY;
                e = 0.f; ne_prev = BL_INF; it = 0;
                if (first_nz == 255) { action = -1; mode = FX_EVALDONE; }       // no action with pi != 0 (the reference would index children[-1])
                else if (tiny) { mode = FX_XALL; D_prev = -1.f; c_fother++; }
                else { mode = FX_ITER; D_prev = 0.f; }
            }
        }
        // ---- B. fast Newton, closed form; D_prev: bound on |reference's previous S - 1  -  ne_prev| (0: exact; -1: all-exact mode) -----
        for (int n = 0; n < nit && __any_sync(FULL, mode == FX_ITER); n++) {
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
                const int ns = nc < FX_KS ? nc : FX_KS;
                for (int i = 0; i < ns; i++) {
                    const uint4 v = fx_lds16(slot_a + 16u * i);
                    term(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z));
                }
                for (int i = FX_KS; i < nc; i++) { const FxEntry en = my_spill[i]; term(en.t, en.q, en.w); }
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
        // ---- C. certified sampling: binary search over cum(a) = k*cpi[a] + (corrections of the children at or before a); decided when
        //         the two prefix sums around the hit are further than delta from r -----------------------------------------------------
        if (__any_sync(FULL, mode == FX_SAMPLE)) {
            if (mode == FX_SAMPLE) {
                const float ra = bl_rcp_fast(alpha);
                const float k = __fmul_rn(lambda, ra);
                for (int i = 0; i < nc; i++) {
                    FxEntry en = get(i);
                    const float s = __fmul_rn(en.t, bl_rcp_fast(__fsub_rn(alpha, en.q)));
                    const float dlt = __fsub_rn(s, __fmul_rn(en.t, ra));
                    if (i < FX_KS) fx_sts(cpr_a + 4u * i, dlt);
                    else my_spill[i].w = dlt;
                }
                const float fL1 = (float)L1;
                auto cum = [&](int a) {
                    float v = __fmul_rn(k, fx_lds(row_a + 4u * a));
                    const float wa = fL1 - (float)a;         // a child at or before a has w = L1 - a_c >= L1 - a
                    const int ns = nc < FX_KS ? nc : FX_KS;
                    for (int i = 0; i < ns; i++)
                        if (fx_lds(slot_a + 16u * i + 8u) >= wa) v = __fadd_rn(v, fx_lds(cpr_a + 4u * i));
                    for (int i = FX_KS; i < nc; i++) {
                        const FxEntry en = my_spill[i];
                        if (fx_a(en.m) <= a) v = __fadd_rn(v, en.w);
                    }
                    return v;
                };
                int lo = 0, hi = A;                          // l = #{a : cum(a) < r}
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (cum(mid) < r) lo = mid + 1; else hi = mid;
                }
                const float delta = __fmaf_rn(__fmul_rn(Gs, e), 1.05f, ESb);
                bool ok = true;
                if (lo < A) ok = ok && (cum(lo) >= r + delta);
                if (lo > 0) ok = ok && (cum(lo - 1) < r - delta);
                for (int i = FX_KS; i < nc; i++) my_spill[i].w = (float)(L1 - fx_a(my_spill[i].m));     // (w held the correction)
                if (r <= 0.f) { action = first_nz; mode = FX_EVALDONE; }
                else if (!ok) { mode = FX_XPASS; c_fsample++; }
                else { action = lo < A ? lo : last_nz; mode = FX_EVALDONE; }
            }
        }
        // ---- D. exact passes: one pass of the reference's loops (bl_newton_f, mcts_core.cuh) for each lane that asked, the warp working on
        //         one lane's row at a time: terms of the row -> shared memory (child positions patched from the lane's list), the S sum on lane
        //         0 and the g sum on lane 1; the running sums of S are kept (sign bit = "term is zero") for the sampling loop ------------------
        for (unsigned pend = __ballot_sync(FULL, mode == FX_XPASS || mode == FX_XALL); pend; pend &= pend - 1) {
            const int ld = __ffs(pend) - 1;
            const bool leader = lane == ld;
            if (leader && mode == FX_XPASS) { alpha = s_alpha; it = s_it; ne_prev = s_nep; D_prev = s_Dp; }
            const float xalpha = __shfl_sync(FULL, alpha, ld), xlam = __shfl_sync(FULL, lambda, ld), xr = __shfl_sync(FULL, r, ld);
            const int xnc = __shfl_sync(FULL, nc, ld), xcur = __shfl_sync(FULL, cur, ld);
            const size_t xslot = (size_t)(blockIdx.x * EPW + ld) * T + xcur;
            const float b2 = __fmul_rn(xalpha, xalpha);
            for (int a = lane; a < AP; a += 32) {
                const float top = __fmul_rn(xlam, t.pi[xslot * AP + a]);
                fx_sts(xs_a + 4u * a, __fdiv_rn(top, xalpha));
                fx_sts(xg_a + 4u * a, __fdiv_rn(-top, b2));
            }
            __syncwarp();
            for (int i = lane; i < xnc; i += 32) {
                FxEntry en;
                if (i < FX_KS) { const uint4 v = fx_lds16(slot0 + (uint32_t)ld * (FX_KS * 16u) + 16u * i); en.t = __uint_as_float(v.x); en.q = __uint_as_float(v.y); en.m = v.w; }
                else en = spill[(size_t)(blockIdx.x * EPW + ld) * cap + i];
                const float bot = __fsub_rn(xalpha, en.q);
                fx_sts(xs_a + 4u * fx_a(en.m), __fdiv_rn(en.t, bot));
                fx_sts(xg_a + 4u * fx_a(en.m), __fdiv_rn(-en.t, __fmul_rn(bot, bot)));
            }
            __syncwarp();
            float acc = 0.f;
            if (lane < 2) {
                const uint32_t arr = lane == 0 ? xs_a : xg_a;
                for (int c = 0; 4 * c < AP; c++) {            // (pad terms are +-0: they leave the sums as they are)
                    const uint4 u = fx_lds16(arr + 16u * c);
                    const float v0 = __uint_as_float(u.x), v1 = __uint_as_float(u.y), v2 = __uint_as_float(u.z), v3 = __uint_as_float(u.w);
                    const float o0 = __fadd_rn(acc, v0), o1 = __fadd_rn(o0, v1), o2 = __fadd_rn(o1, v2), o3 = __fadd_rn(o2, v3);
                    acc = o3;
                    if (lane == 0)                            // |.| = running sum, sign bit set <=> p == 0
                        fx_sts16(arr + 16u * c, make_uint4(__float_as_uint(v0 > 0.f ? o0 : -o0), __float_as_uint(v1 > 0.f ? o1 : -o1),
                                                           __float_as_uint(v2 > 0.f ? o2 : -o2), __float_as_uint(v3 > 0.f ? o3 : -o3)));
                }
            }
            const float xS = __shfl_sync(FULL, acc, 0), xg_ = __shfl_sync(FULL, acc, 1);
            __syncwarp();
            bool take = false;                                // the pass ends the Newton loop: sample from its running sums
            if (leader) {
                c_xpass++;
                it++;
                const float ne = __fsub_rn(xS, 1.f);
                const bool errknown = D_prev <= 0.f;          // the reference's previous S - 1 is known exactly
                if (it > 100) take = true;                    // the extra pass after the loop bound: sums at the final alpha, no test
                else if (ne < 1e-3f) take = true;
                else if (errknown && ne_prev == ne) take = true;
                else if (!errknown && !(fabsf(ne - ne_prev) > D_prev)) {
                    // `error == new_error` cannot be decided from an approximate predecessor: the whole evaluation again, exactly
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
                // sampling loop (cuda.cu:160-176): first a with p > 0 and total >= r, else the last a with p > 0
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
        // ---- E. advance: step to the chosen child; its header and row start travelling ------------------------------------------------------
        if (mode == FX_EVALDONE) {
            c_evals++; c_children += nc; c_iters += it;
            parent = cur;
            int next = -1, nflags = 0;
            for (int i = 0; i < nc; i++) {
                const uint32_t m = get(i).m;
                if (fx_a(m) == action) { next = (int)(m >> 24); nflags = (int)((m >> 14) & 3u); }
            }
            cur = action >= 0 ? next : -1;
            if (cur >= 0 && !(nflags & 2)) { cur_seat = nflags & 1; mode = FX_VISIT; fetch_node(cur); }
            else mode = FX_IDLE;                             // new leaf, existing terminal child, or no legal action
        }
    }
    // ---- results; expand + env step of the warp's envs, fused when the lane rows are big enough for a board and its flood-fill stack ----
    if (b < t.B) {
        t.leaf[b] = (int16_t)cur;                           // existing terminal child, or -1: the expand step decides
        t.leaf_parent[b] = (int16_t)parent;
        t.leaf_action[b] = (int16_t)action;
        if (fuse_expand) bl_expand_one(t, sim, b, cur, parent, action, reinterpret_cast<uint32_t *>(fx_smem) + lane * PS,
                                       reinterpret_cast<uint8_t *>(fx_smem) + (uint32_t)EPW * PS * 4u + (uint32_t)lane * (FX_KS * 16u));
    }
    bl_count(t.counters, C_EVALS, c_evals);
    bl_count(t.counters, C_CHILDREN, c_children);
    bl_count(t.counters, C_ITERS, c_iters);
    bl_count(t.counters, C_DESCENTS, b < t.B ? 1u : 0u);
    bl_count(t.counters, C_FLAG_STOP, c_fstop);
    bl_count(t.counters, C_FLAG_SAMPLE, c_fsample);
    bl_count(t.counters, C_FLAG_OTHER, c_fother);
    bl_count(t.counters, C_EXACT_PASSES, c_xpass);
}

int g_fx_nit = 4;          // BL_FX_NIT: Newton passes per trip

template <int NCH, int KW, int EPW>
int launch_fx_e(const bl_tree *t, int sim, const bl_half *rands, uint64_t seed, cudaStream_t st) {
    const int cap = t->A < t->T - 1 ? t->A : (t->T > 1 ? t->T - 1 : 1);
    const size_t smem = (size_t)EPW * NCH * 16 + (size_t)EPW * FX_KS * 16 + (size_t)EPW * FX_KS * 4 + (size_t)2 * t->AP * 4;
    if (smem > 227 * 1024) return -2;
    if ((int64_t)((t->B + 31) / 32) * 32 * cap * (int64_t)sizeof(FxEntry) > t->scratch_bytes) return -3;
    auto kern = descend_fx_kernel<NCH, KW, EPW>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    // expand + env step in the same kernel when a lane's row holds a board and its slots the flood-fill stack
    const int fuse = (t->BP <= NCH * 16 && t->A <= FX_KS * 16) ? 1 : 0;
    kern<<<(t->B + EPW - 1) / EPW, 32, smem, st>>>(*t, sim, rands, seed, reinterpret_cast<FxEntry *>(t->scratch), cap, g_fx_nit, fuse);
    if (cudaError_t e = cudaGetLastError()) return (int)e;
    return fuse ? 0 : bl_expand_step(t, sim, st);
}

int g_fx_epw = 0;          // BL_FX_EPW: envs per warp (0 = default)
template <int NCH, int KW>
int launch_fx(const bl_tree *t, int sim, const bl_half *rands, uint64_t seed, cudaStream_t st) {
    const int epw = g_fx_epw ? g_fx_epw : 8;
    if (epw >= 32) return launch_fx_e<NCH, KW, 32>(t, sim, rands, seed, st);
    if (epw >= 16) return launch_fx_e<NCH, KW, 16>(t, sim, rands, seed, st);
    if (epw >= 8) return launch_fx_e<NCH, KW, 8>(t, sim, rands, seed, st);
    return launch_fx_e<NCH, KW, 4>(t, sim, rands, seed, st);
}

template <int KW>
int launch_fx_kw(const bl_tree *t, int sim, const bl_half *rands, uint64_t seed, cudaStream_t st) {
    const int nch = t->AP / 4;
    if (nch <= 7) return launch_fx<7, KW>(t, sim, rands, seed, st);
    if (nch <= 13) return launch_fx<13, KW>(t, sim, rands, seed, st);
    if (nch <= 21) return launch_fx<21, KW>(t, sim, rands, seed, st);
    if (nch <= 31) return launch_fx<31, KW>(t, sim, rands, seed, st);
    if (nch <= 43) return launch_fx<43, KW>(t, sim, rands, seed, st);
    if (nch <= 57) return launch_fx<57, KW>(t, sim, rands, seed, st);
    return -2;
}

}  // namespace

// -2: unsupported shape, -3: scratch too small (the caller falls back to the exact kernels)
int64_t bl_fx_scratch_bytes(const bl_tree *t) {
    const int64_t cap = t->A < t->T - 1 ? t->A : (t->T > 1 ? t->T - 1 : 1);
    return ((int64_t)t->B + 31) / 32 * 32 * cap * (int64_t)sizeof(FxEntry);
}

int bl_descend_fx(const bl_tree *t, int sim, const bl_half *rands, uint64_t seed, cudaStream_t st) {
    static bool env_read = false;
    if (!env_read) {
        env_read = true;
        if (const char *e = getenv("BL_FX_NIT")) { const int v = atoi(e); if (v > 0) g_fx_nit = v; }
        if (const char *e = getenv("BL_FX_EPW")) g_fx_epw = atoi(e);
    }
    if (t->A > 255 || t->T > 256) return -2;
    if (t->T <= 64) return launch_fx_kw<1>(t, sim, rands, seed, st);
    if (t->T <= 128) return launch_fx_kw<2>(t, sim, rands, seed, st);
    return launch_fx_kw<4>(t, sim, rands, seed, st);
}
