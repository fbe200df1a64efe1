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
// Mapping: G lanes per env (template), EPL = ceil(A/G) row elements per lane in registers (needed once per evaluation: mass
// and weight of the row, and the sampling prefix sums, in double); the children of the node live in a per-env shared-memory
// list {t, q, action, flags} filled by the lane that fetched the child's record; the Newton loop runs redundantly on every lane
// of the group over that list (broadcast reads, no shuffles).  All groups of a warp run in lock step, one evaluation per trip.
//
// Compiled with -fmad=false -prec-div=true -ftz=false (see build.py); fused operations are explicit.
#include <cstdio>
#include <cstdlib>

#include "descend_common.cuh"

namespace {

constexpr float FX_U = 5.9604644775390625e-08f;        // 2^-24
constexpr float FX_GUARD = 0.00390625f;                // e <= 2^-8 (alpha - q_max): functions of alpha vary < 2 % across the doubt
constexpr int FX_MAXIT = 24;                           // fast-path Newton passes before giving up (the exact path has the reference's 100)
constexpr int FX_THREADS = 128;

__device__ __forceinline__ float fx_shfl_xor(float v, int o, int width) { return __shfl_xor_sync(FULL, v, o, width); }
__device__ __forceinline__ double fx_shfl_xor(double v, int o, int width) { return __shfl_xor_sync(FULL, v, o, width); }

// index of the i-th (0-based) set bit of a KW-word mask
template <int KW>
__device__ __forceinline__ int fx_nth_set(const u64 (&mm)[KW], int i) {
    int base = 0;
#pragma unroll
    for (int w = 0; w < KW; w++) {
        const int pc = __popcll(mm[w]);
        if (i < pc) {
            const uint32_t lo32 = (uint32_t)mm[w], hi32 = (uint32_t)(mm[w] >> 32);
            const int pl = __popc(lo32);
            return base + (i < pl ? (int)__fns(lo32, 0, i + 1) : 32 + (int)__fns(hi32, 0, i - pl + 1));
        }
        i -= pc;
        base += 64;
    }
    return -1;
}

struct FxEntry { float t; uint32_t m; };               // m = half(q) | action << 16 | seat << 24 | terminal << 25 (terminal as a flag)

template <int G, int EPL, int KW, bool DBG>
__global__ void __launch_bounds__(FX_THREADS, 8) descend_fx_kernel(bl_tree t, int sim, const bl_half *__restrict__ rands, uint64_t seed,
                                                                int cap, int env_words, float *__restrict__ dbg) {
    extern __shared__ __align__(16) uint32_t fx_smem[];
    constexpr int EPB = FX_THREADS / G;                 // envs per block
    const int A = t.A, T = t.T, AP = t.AP;
    const int lane = threadIdx.x & 31, j = lane % G;
    const int genv = threadIdx.x / G;                   // env slot in the block
    const int b = blockIdx.x * EPB + genv;
    const int lo = j * EPL;                             // first row element of this lane
    uint32_t *envm = fx_smem + (size_t)genv * env_words;
    FxEntry *list = reinterpret_cast<FxEntry *>(envm);                  // [cap]
    float *dA = reinterpret_cast<float *>(envm + 2 * ((cap + 1) & ~1)); // [AP]: sampling corrections at child positions, 0 elsewhere
    float *xg = dA + AP;                                                // [AP]: exact path scratch
    const bl_qnorm qn(t.qrange + 2 * sim);
    const uint64_t move = t.counters[C_MOVE];
    const uint64_t keep = bl_policy_keep();

    bool alive = b < t.B;
    int cur = 0, parent = 0, action = -1, cur_seat = 0;
    float c_puct = 0.f;
    unsigned c_evals = 0, c_children = 0, c_iters = 0, c_fstop = 0, c_fsample = 0, c_fother = 0, c_xpass = 0;
    for (int a = lo; a < lo + EPL && a < AP; a++) dA[a] = 0.f;
    if (alive) {
        const bl_node root = bl_ld_node_hint(t.node + (size_t)b * T, keep);
        c_puct = bl_h2f(t.c_puct[b]);
        cur_seat = root.seat;
        if (root.terminal) alive = false;               // leaf = 0, no action: the expand step records the error (as descend.cu)
    }
    bool descending = alive;
    __syncwarp();

    while (__any_sync(FULL, descending)) {
        // ---- visit: row summary, children mask, row slice ------------------------------------------------------------------------
        const size_t node0 = (size_t)(b < t.B ? b : 0) * T;
        const size_t slot = node0 + (descending ? cur : 0);
        bl_aux ax = bl_ld_aux_hint(t.aux + slot, keep);
        u64 mm[KW];
#pragma unroll
        for (int w = 0; w < KW; w++) mm[w] = (w < ((T + 63) >> 6)) ? t.kids[slot * ((T + 63) >> 6) + w] : 0ull;
        float pr[EPL];
        {
            const float4 *row = reinterpret_cast<const float4 *>(t.pi + slot * AP);
#pragma unroll
            for (int c = 0; c < EPL / 4; c++) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (lo + 4 * c < AP) v = row[(lo >> 2) + c];
                pr[4 * c] = v.x; pr[4 * c + 1] = v.y; pr[4 * c + 2] = v.z; pr[4 * c + 3] = v.w;
            }
        }
        float r;
        if (rands) r = bl_h2f(rands[slot]);
        else r = bl_uniform_half_grid(bl_philox(seed ^ (move * 0x9E3779B97F4A7C15ull), (uint64_t)b, ((uint64_t)sim << 32) | (uint32_t)cur).x);
        int nc = 0;
#pragma unroll
        for (int w = 0; w < KW; w++) nc += __popcll(mm[w]);
        if (!descending) nc = 0;
        const int first_nz = ax.first_nz, last_nz = ax.last_nz;
        const bool empty_row = first_nz == 255;          // no action with pi != 0: the reference would index children[-1]
        // ---- adoption: lane j fetches the records of children j, j+G, ...; entry = {pi_c for now, q, action, flags} -----------------
        int N = 0;
        int ncmax = nc;
#pragma unroll
        for (int o = 16; o; o >>= 1) ncmax = max(ncmax, __shfl_xor_sync(FULL, ncmax, o));
        for (int i0 = 0; i0 < ncmax; i0 += G) {
            const int i = i0 + j;
            if (i < nc) {
                const int id = fx_nth_set<KW>(mm, i);
                const bl_node ch = bl_ld_node_hint(t.node + node0 + id, keep);
                const int a = ch.relation;
                const float pic = t.pi[slot * AP + a];
                const float q = qn.fast(cur_seat ? ch.w[1] : ch.w[0], ch.n);
                N += ch.n;
                FxEntry en;
                en.t = pic;
                en.m = (uint32_t)bl_f2h(q) | ((uint32_t)a << 16) | ((uint32_t)ch.seat << 24) | ((uint32_t)(ch.terminal ? 1u : 0u) << 25);
                list[i] = en;
            }
        }
#pragma unroll
        for (int o = G / 2; o; o >>= 1) N += __shfl_xor_sync(FULL, N, o, G);
        N += A - nc;                                     // every child-less action counts 1 (cuda.cu:91)
        const float lambda = bl_lambda(c_puct, N, A);
        for (int i0 = 0; i0 < ncmax; i0 += G) {
            const int i = i0 + j;
            if (i < nc) list[i].t = __fmul_rn(lambda, list[i].t);     // t_c = RN(lambda*pi_c), the reference's product
        }
        __syncwarp();
        // ---- once per evaluation, in double: mass P and addition-count weight W of the row; exclusive prefix of the lane ------------
        double Pl = 0., Wl = 0.;
#pragma unroll
        for (int e = 0; e < EPL; e++) {
            const double p = (double)pr[e];
            Pl += p;
            Wl = fma(p, (double)max(last_nz + 1 - (lo + e), 0), Wl);
        }
        double Pin = Pl;                                  // inclusive scan over the group
#pragma unroll
        for (int o = 1; o < G; o <<= 1) {
            const double up = __shfl_up_sync(FULL, Pin, o, G);
            if (j >= o) Pin += up;
        }
        const double Ppre = Pin - Pl;
        double Pall = __shfl_sync(FULL, Pin, G - 1, G), Wall = Wl;
#pragma unroll
        for (int o = G / 2; o; o >>= 1) Wall += fx_shfl_xor(Wall, o, G);
        // child-less mass M = lambda*P - sum t_c, weight MW likewise; alpha seed; q_max  (redundantly on every lane of the group)
        double Md = (double)lambda * Pall, MWd = (double)lambda * Wall;
        float alpha = fmaxf(__fmul_rn(lambda, ax.max_pi), 1.e-4f), qmax = 0.f;
        for (int i = 0; i < nc; i++) {
            const FxEntry en = list[i];
            const float q = bl_h2f((bl_half)(en.m & 0xFFFF));
            const int a = (en.m >> 16) & 255;
            Md -= (double)en.t;
            MWd -= (double)en.t * (double)max(last_nz + 1 - a, 0);
            alpha = fmaxf(alpha, __fadd_rn(q, fmaxf(en.t, 1.e-4f)));
            qmax = fmaxf(qmax, q);
        }
        const float M = fmaxf((float)Md, 0.f), MW = fmaxf((float)MWd, 0.f);
        const float alpha0 = alpha;
        const float cours = (float)(nc + 8) * FX_U;
        bool tiny = __fmul_rn(lambda, bl_minnz(ax)) < BL_TINY;
        // ---- Newton, closed form, certified -----------------------------------------------------------------------------------------
        bool evaluating = descending && !empty_row;
        bool iterating = evaluating && !tiny;
        int flag = (evaluating && tiny) ? 4 : 0;         // 1 stop test, 2 sample, 3 guard, 4 tiny
        float e = 0.f, ne_prev = BL_INF, D_prev = 0.f, S = 0.f, Gs = 1.f, ESb = 0.f;
        int it = 0;
        while (__any_sync(FULL, iterating)) {
            const float ra = bl_rcp_fast(alpha);
            float Sn = __fmul_rn(M, ra), Gn = __fmul_rn(Sn, ra), Hn = __fmul_rn(Gn, ra);
            float ES = __fmul_rn(MW, ra), EG = __fmul_rn(ES, ra);
            for (int i = 0; i < nc; i++) {
                const FxEntry en = list[i];
                const float q = bl_h2f((bl_half)(en.m & 0xFFFF));
                const float w = (float)max(last_nz + 1 - (int)((en.m >> 16) & 255), 0);
                const float rc = bl_rcp_fast(__fsub_rn(alpha, q));
                const float s = __fmul_rn(en.t, rc), g = __fmul_rn(s, rc);
                Sn = __fadd_rn(Sn, s); Gn = __fadd_rn(Gn, g); Hn = __fadd_rn(Hn, __fmul_rn(g, rc));
                ES = __fmaf_rn(w, s, ES); EG = __fmaf_rn(w, g, EG);
            }
            if (iterating) {
                Hn = __fmul_rn(2.f, Hn);
                const float ESn = __fmaf_rn(FX_U * 1.01f, ES, __fmul_rn(2.f * FX_U + cours, Sn));
                const float EGn = __fmaf_rn(FX_U * 1.01f, EG, __fmul_rn(4.f * FX_U + cours, Gn));
                const float ne = __fsub_rn(Sn, 1.f);
                const float Dk = __fmaf_rn(__fmul_rn(Gn, e), 1.05f, __fmaf_rn(2.f * FX_U, fabsf(ne), ESn));
                S = Sn; Gs = Gn; ESb = ESn;
                it++;
                const bool guard_ok = (e <= FX_GUARD * (alpha - qmax)) && (Sn < 3.0e38f) && (Gn > 0.f) && (Gn < 3.0e38f) && it <= FX_MAXIT;
                const bool stop_sure = ne < 1e-3f - Dk;
                const bool cont_sure = (ne > 1e-3f + Dk) && (fabsf(ne - ne_prev) > Dk + D_prev);
                if (DBG && dbg && j == 0 && it <= 8) {
                    float *d = dbg + ((size_t)b * 64) + (it - 1) * 4;
                    d[0] = alpha; d[1] = ne; d[2] = Dk; d[3] = e;
                }
                if (!guard_ok) { flag = 3; iterating = false; }
                else if (stop_sure) iterating = false;
                else if (!cont_sure) { flag = 1; iterating = false; }
                else {
                    const float L = 1.2f * (fmaxf(ne, 0.f) + 2.f * Gn * e) * Hn / (Gn * Gn);
                    const float R = 1.05f * (Dk / Gn + fabsf(ne) / Gn * (EGn / Gn + 4.f * FX_U));
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
        // ---- sample: prefix sums in double from the lane's exclusive offset; certified when no prefix sum lies within delta of r ------
        const float ra = bl_rcp_fast(alpha);
        const float k = __fmul_rn(lambda, ra);
        double off = (double)k * Ppre;
        for (int i0 = 0; i0 < ncmax; i0 += G) {           // owners park the child corrections s_c - t_c/alpha at the child positions
            const int i = i0 + j;
            if (i < nc && flag == 0) {
                const FxEntry en = list[i];
                const float q = bl_h2f((bl_half)(en.m & 0xFFFF));
                const float s = __fmul_rn(en.t, bl_rcp_fast(__fsub_rn(alpha, q)));
                dA[(en.m >> 16) & 255] = __fsub_rn(s, __fmul_rn(en.t, ra));
            }
        }
        __syncwarp();
        for (int i = 0; i < nc; i++) {                    // corrections of the children in front of this lane's slice
            const FxEntry en = list[i];
            const int a = (en.m >> 16) & 255;
            if (a < lo && flag == 0) off += (double)dA[a];
        }
        const float delta = __fmaf_rn(__fmul_rn(Gs, e), 1.05f, ESb);
        const double thr_lo = (double)r - (double)delta, thr_hi = (double)r + (double)delta;
        int cnt_lo = 0, cnt_hi = 0;
        {
            double acc = off;
#pragma unroll
            for (int c = 0; c < EPL / 4; c++) {
                float4 dv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (lo + 4 * c < AP) dv = *reinterpret_cast<const float4 *>(dA + lo + 4 * c);
                const float dd[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    acc += (double)__fmaf_rn(k, pr[4 * c + u], dd[u]);
                    const bool in = lo + 4 * c + u < A;
                    cnt_lo += (in && acc < thr_lo) ? 1 : 0;
                    cnt_hi += (in && acc <= thr_hi) ? 1 : 0;
                }
            }
        }
#pragma unroll
        for (int o = G / 2; o; o >>= 1) {
            cnt_lo += __shfl_xor_sync(FULL, cnt_lo, o, G);
            cnt_hi += __shfl_xor_sync(FULL, cnt_hi, o, G);
        }
        __syncwarp();
        for (int i0 = 0; i0 < ncmax; i0 += G) {           // the table is zero outside the sampling step
            const int i = i0 + j;
            if (i < nc) dA[(list[i].m >> 16) & 255] = 0.f;
        }
        if (evaluating && flag == 0) {
            if (r <= 0.f) action = first_nz;
            else if (cnt_lo != cnt_hi) flag = 2;
            else action = cnt_lo < A ? cnt_lo : last_nz;
        }
        if (DBG && dbg && j == 0 && evaluating) {
            float *d = dbg + ((size_t)b * 64);
            d[32] = alpha; d[33] = delta; d[34] = (float)flag; d[35] = (float)it; d[36] = e; d[37] = S; d[38] = Gs; d[39] = (float)action;
        }
        c_fstop += (flag == 1 && j == 0); c_fsample += (flag == 2 && j == 0); c_fother += (flag >= 3 && j == 0);
        __syncwarp();
        // ---- exact path: the reference's loops (bl_newton_f / bl_sample_f, mcts_core.cuh) on the group's lanes --------------------------
        // terms of the lane's slice -> shared memory (child positions patched by the child's owner lane), the two sequential sums on
        // lane 0 of the group, running sums of S kept for the sampling loop
        bool exact_run = flag != 0;
        if (__any_sync(FULL, exact_run)) {
            float xalpha = alpha0, xerror = BL_INF;
            int xit = 0;
            bool xiter = exact_run;
            float *sT = dA, *gT = xg;
            while (__any_sync(FULL, xiter)) {
                if (xiter) {
                    const float b2 = __fmul_rn(xalpha, xalpha);
#pragma unroll
                    for (int u = 0; u < EPL; u++) {
                        if (lo + u < AP) {
                            const float top = __fmul_rn(lambda, pr[u]);
                            sT[lo + u] = __fdiv_rn(top, xalpha);
                            gT[lo + u] = __fdiv_rn(-top, b2);
                        }
                    }
                }
                __syncwarp();
                for (int i0 = 0; i0 < ncmax; i0 += G) {
                    const int i = i0 + j;
                    if (i < nc && xiter) {
                        const FxEntry en = list[i];
                        const float q = bl_h2f((bl_half)(en.m & 0xFFFF));
                        const int a = (en.m >> 16) & 255;
                        const float bot = __fsub_rn(xalpha, q);
                        sT[a] = __fdiv_rn(en.t, bot);
                        gT[a] = __fdiv_rn(-en.t, __fmul_rn(bot, bot));
                    }
                }
                __syncwarp();
                float xS = 0.f, xg_ = 0.f;
                if (xiter && j == 0) {
                    for (int a = 0; a < A; a++) {
                        xS = __fadd_rn(xS, sT[a]);
                        xg_ = __fadd_rn(xg_, gT[a]);
                        gT[a] = xS;                      // running sum of S: the sampling loop's `total`
                    }
                }
                xS = __shfl_sync(FULL, xS, 0, G);
                xg_ = __shfl_sync(FULL, xg_, 0, G);
                __syncwarp();
                if (xiter) {
                    xit++;
                    c_xpass += (j == 0);
                    const float ne = __fsub_rn(xS, 1.f);
                    if (xit > 100) xiter = false;                                    // the extra pass: sums at the final alpha, no test
                    else if ((ne < 1e-3f) || (xerror == ne)) xiter = false;
                    else {
                        // (when the 100th pass does not stop, the reference's loop ends with alpha updated once more and the sampling
                        // loop recomputes the sums with it: pass 101 here, as descend.cu's ST_FINAL)
                        xalpha = __fsub_rn(xalpha, __fdiv_rn(ne, xg_));
                        xerror = ne;
                    }
                }
            }
            // sampling loop (cuda.cu:160-176): first a with p > 0 and total >= r, else the last a with p > 0
            int hit = 1 << 20, lastv = -1;
            if (exact_run) {
#pragma unroll
                for (int u = 0; u < EPL; u++) {
                    const int a = lo + u;
                    if (a < A) {
                        const float p = sT[a], tot = gT[a];
                        if (p > 0.f) {
                            if (tot >= r && hit == (1 << 20)) hit = a;
                            lastv = a;
                        }
                    }
                }
            }
#pragma unroll
            for (int o = G / 2; o; o >>= 1) {
                hit = min(hit, __shfl_xor_sync(FULL, hit, o, G));
                lastv = max(lastv, __shfl_xor_sync(FULL, lastv, o, G));
            }
            __syncwarp();
            if (exact_run) {
                action = hit < (1 << 20) ? hit : lastv;
                it = xit > 100 ? 100 : xit;
#pragma unroll
                for (int u = 0; u < EPL; u++)
                    if (lo + u < AP) sT[lo + u] = 0.f;   // dA is zero outside the sampling step
            }
            __syncwarp();
        }
        // ---- advance ------------------------------------------------------------------------------------------------------------------------
        if (descending) {
            if (j == 0) { c_evals++; c_children += nc; c_iters += it; }
            parent = cur;
            if (empty_row) action = -1;
            int next = -1, nflags = 0;
            for (int i = 0; i < nc; i++) {
                const uint32_t m = list[i].m;
                if ((int)((m >> 16) & 255) == action) { next = fx_nth_set<KW>(mm, i); nflags = (int)(m >> 24); }
            }
            cur = action >= 0 ? next : -1;
            if (cur >= 0 && !(nflags & 2)) cur_seat = nflags & 1;
            else descending = false;                    // new leaf, existing terminal child, or no legal action
        }
        __syncwarp();
    }
    if (alive || (b < t.B)) {
        if (j == 0) {
            t.leaf[b] = (int16_t)cur;                   // existing terminal child, or -1: the expand step decides
            t.leaf_parent[b] = (int16_t)parent;
            t.leaf_action[b] = (int16_t)action;
        }
    }
    bl_count(t.counters, C_EVALS, c_evals);
    bl_count(t.counters, C_CHILDREN, c_children);
    bl_count(t.counters, C_ITERS, c_iters);
    bl_count(t.counters, C_DESCENTS, (b < t.B && j == 0) ? 1u : 0u);
    bl_count(t.counters, C_FLAG_STOP, c_fstop);
    bl_count(t.counters, C_FLAG_SAMPLE, c_fsample);
    bl_count(t.counters, C_FLAG_OTHER, c_fother);
    bl_count(t.counters, C_EXACT_PASSES, c_xpass);
}

float *g_fx_dbg = nullptr;
int g_fx_lanes = 0;        // BL_FX_LANES: lanes per env (0 = by board size)

template <int G, int EPL, int KW>
int launch_fx(const bl_tree *t, int sim, const bl_half *rands, uint64_t seed, cudaStream_t st) {
    constexpr int EPB = FX_THREADS / G;
    const int cap = t->A < t->T - 1 ? t->A : (t->T > 1 ? t->T - 1 : 1);
    int env_words = 2 * ((cap + 1) & ~1) + 2 * t->AP;
    env_words = (env_words + 3) & ~3;
    if ((env_words & 31) == 0) env_words += 4;          // consecutive envs start in different banks
    const size_t smem = (size_t)EPB * env_words * 4;
    if (smem > 227 * 1024) return -2;
    auto kern = g_fx_dbg ? descend_fx_kernel<G, EPL, KW, true> : descend_fx_kernel<G, EPL, KW, false>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    kern<<<(t->B + EPB - 1) / EPB, FX_THREADS, smem, st>>>(*t, sim, rands, seed, cap, env_words, g_fx_dbg);
    return (int)cudaGetLastError();
}

template <int G, int KW>
int launch_fx_g(const bl_tree *t, int sim, const bl_half *rands, uint64_t seed, cudaStream_t st) {
    const int epl = ((t->A + G - 1) / G + 3) & ~3;
    switch (epl) {
        case 4: return launch_fx<G, 4, KW>(t, sim, rands, seed, st);
        case 8: return launch_fx<G, 8, KW>(t, sim, rands, seed, st);
        case 12: return launch_fx<G, 12, KW>(t, sim, rands, seed, st);
        case 16: return launch_fx<G, 16, KW>(t, sim, rands, seed, st);
        case 24: return launch_fx<G, 24, KW>(t, sim, rands, seed, st);
        case 32: return launch_fx<G, 32, KW>(t, sim, rands, seed, st);
        default: return -2;
    }
}

template <int KW>
int launch_fx_kw(const bl_tree *t, int sim, const bl_half *rands, uint64_t seed, cudaStream_t st) {
    int lanes = g_fx_lanes;
    if (lanes == 0) lanes = t->A <= 32 ? 4 : 8;
    int rc = -2;
    if (lanes == 4) rc = launch_fx_g<4, KW>(t, sim, rands, seed, st);
    if (lanes == 8 || rc == -2) rc = launch_fx_g<8, KW>(t, sim, rands, seed, st);
    if (rc == -2) rc = launch_fx_g<16, KW>(t, sim, rands, seed, st);
    return rc;
}

}  // namespace

// -2: unsupported shape (the caller falls back to the exact kernels)
int bl_descend_fx(const bl_tree *t, int sim, const bl_half *rands, uint64_t seed, cudaStream_t st) {
    static bool env_read = false;
    if (!env_read) {
        env_read = true;
        if (const char *e = getenv("BL_FX_LANES")) g_fx_lanes = atoi(e);
    }
    if (t->A > 255 || t->T > 256) return -2;
    int rc;
    if (t->T <= 64) rc = launch_fx_kw<1>(t, sim, rands, seed, st);
    else if (t->T <= 128) rc = launch_fx_kw<2>(t, sim, rands, seed, st);
    else rc = launch_fx_kw<4>(t, sim, rands, seed, st);
    if (rc) return rc;
    return bl_expand_step(t, sim, st);
}

extern "C" int bl_debug_set_fx_trace(float *buf) {
    g_fx_dbg = buf;
    return 0;
}
