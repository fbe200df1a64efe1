// Tree descent with FOUR lanes per env for sm_100a — the hot loop of MCTS.simulate (descend_kernel + policy + newton_search,
// boardlaw/mcts/cpp/cuda.cu:35-99,138-182), same arithmetic as descend.cu (one lane per env), different mapping.
//
// Why: with one lane per env the number of warps is B/32 (c2: 1 024 warps = 1.7 per SM sub-partition) and the kernel runs at
// the LATENCY of its own dependent chains (issue-active 33 %).  The only part of a Newton pass that is
// inherently serial is the pair of fp32 running sums (the reference adds the A terms left to right); the terms themselves —
// two correctly rounded divisions per action — are independent.  So an env gets a group of four lanes:
//
//   * term phase: each lane computes the S and g terms of every fourth 4-action chunk from its register-resident slice of
//     the lambda*pi row (packed FMUL2/FFMA2 Markstein division by the shared divisors alpha and alpha^2) and parks them in
//     the env's two shared-memory rows;
//   * child terms (actions that have a child: full divisions) are patched into the rows by the lane that owns the child;
//   * chain phase: lane 0 of the group runs the S chain and lane 1 the g chain over the rows (one 128-bit load, four
//     dependent additions and — for S — one 128-bit store of the running sums per chunk); the running sums are the sampling
//     loop's `total`;
//   * services (sample / advance / visit) are split over the four lanes as well: the inverse-CDF count is a four-way
//     partial count, a child's record is fetched and adopted by the lane (child id & 3), reductions are group shuffles.
//
// This gives 4x the warps (c2: 4 096 = 28 per SM, 7 per sub-partition) at a quarter of the registers per lane (the row slice is
// 2*ceil(NCH/4) register pairs), so the chains of different envs overlap in the issue slots the one-lane kernel leaves idle.
// Every fp32 operation, its operands and its order are those of descend.cu (and of the oracle); only who executes it differs.
//
// Compiled with -fmad=false -prec-div=true -ftz=false (see build.py); fused operations are explicit.
#include <cstdio>
#include <cstdlib>

#include "descend_common.cuh"

namespace {

template <int NCH, int L>
struct MwCfg {
    static constexpr int PS = 4 * NCH;                 // floats per row; NCH odd => eight consecutive rows start in eight different bank quartets
    static constexpr int CPL = (NCH + L - 1) / L;      // chunks per lane
    static constexpr int KS = NCH > 31 ? 15 : NCH;     // child entries per env held in shared memory (the rest go to global scratch).  13x13: 15
                                                       // entries buy a fourth CTA per SM (26.5 vs 29.5 ms/move at c5-13); at 11x11 the fifth CTA does
                                                       // not pay for the spilled entries (c5-11 20.2 vs 18.9, c3 138 vs 142)
    static constexpr int ENVS = 32;                    // envs per CTA (L warps)
    static constexpr int ROWS_BYTES = ENVS * 2 * PS * 4;
    static constexpr int SMEM = ROWS_BYTES + ENVS * 16 * KS;
    static constexpr int FIT = 233472 / (SMEM + 1024);
    static constexpr int MINB = FIT > 7 ? 7 : (FIT < 1 ? 1 : FIT);
};

__device__ __forceinline__ void cp16(uint32_t dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp8(uint32_t dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}

// shared-memory accesses by 32-bit address: one base register per env, everything else is an immediate offset (with generic
// pointers the compiler, at 72 registers, recomputed the row addresses from threadIdx at every access: 20 % of the instructions)
__device__ __forceinline__ float4 lds4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts4(uint32_t a, const float4 &v) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float lds1(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts1(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ uint32_t ldsu(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t opaque(uint32_t x) {
    uint32_t y;
    asm volatile("mov.u32 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}

template <int NCH, int L>
__global__ void __launch_bounds__(32 * L, MwCfg<NCH, L>::MINB)
descend_mw_kernel(bl_tree t, int sim, const bl_half *__restrict__ rands, uint64_t seed, ChildEntry *__restrict__ clists, int cap,
                  int gate_num, int gate_den, int fuse_expand) {
    using C = MwCfg<NCH, L>;
    constexpr int PS = C::PS, CPL = C::CPL, KS = C::KS;
    constexpr int MW = 8;                              // 32-bit words of the children-of-this-node mask (T <= 256; else list walk)
    constexpr uint32_t OWN = L == 4 ? 0x11111111u : 0x55555555u;   // node ids that are 0 mod L
    extern __shared__ float4 smem4[];
    const int A = t.A, T = t.T;
    const int lane = threadIdx.x & 31, sub = lane & (L - 1), gl = lane & ~(L - 1);
    const unsigned gmask = ((1u << L) - 1u) << gl;     // the env's lanes
    const int slot = threadIdx.x / L;
    // per env two adjacent rows: [S terms in / running S sums out | g terms]; row pitch 4*NCH words, NCH odd, so the rows the
    // chain lanes of a quarter-warp read (L = 2: S and g rows of four envs) start in different bank quartets
    // (the two bases pass through an opaque move: otherwise the compiler re-derives them from the shared window / threadIdx at
    // every access instead of keeping them in a register)
    const uint32_t ps_addr = opaque(smem_u32(smem4) + (uint32_t)slot * (2 * PS * 4)), pg_addr = ps_addr + PS * 4;
    const uint32_t pe_addr = opaque(smem_u32(smem4) + C::ROWS_BYTES + (uint32_t)slot * (16 * KS));   // child entries {q, top, action | id << 8, flags}
    const bl_qnorm qn(t.qrange + 2 * sim);
    const int nrow4 = t.AP >> 2;
    const int KW = (T + 63) >> 6;
    const bool scan_ok = T <= 32 * MW;
    const int KW32 = (T + 31) >> 5;

    int b = (int)blockIdx.x * C::ENVS + slot;
    if (b >= t.B) b = -1;
    ChildEntry *cl = clists + (size_t)(b < 0 ? 0 : b) * cap;

    u64 tp[2 * CPL];                                  // this lane's slice of lambda*pi: chunks sub, sub+4, ... as element pairs
    int cur = 0, parent = 0, action = -1, state = ST_IDLE, nc = 0, nown = 0, it = 0, cur_seat = 0;
    int res_leaf = -1, res_parent = 0, res_action = -1;
    float alpha = 1.f, error = 0.f, r = 0.f, c_puct = 0.f;
    uint32_t nzpos = 0;
    unsigned c_evals = 0, c_children = 0, c_iters = 0;
#pragma unroll
    for (int k = 0; k < 2 * CPL; k++) tp[k] = 0;

    // child entry j of this lane lives in slot L*j + sub (so a slot's owner is slot mod L)
    auto get = [&](int i) {
        ChildEntry e;
        if (i < KS) {
            const float4 v = lds4(pe_addr + 16u * i);
            e.q = v.x; e.top = v.y;
            const uint32_t u = __float_as_uint(v.z);
            e.a = u & 255; e.id = u >> 8; e.flags = __float_as_int(v.w);
        } else e = cl[i];
        return e;
    };
    auto put = [&](int i, const ChildEntry &e) {
        if (i < KS) sts4(pe_addr + 16u * i, make_float4(e.q, e.top, __uint_as_float((uint32_t)e.a | ((uint32_t)e.id << 8)), __int_as_float(e.flags)));
        else cl[i] = e;
    };
    // asynchronous fetch of what a visit of node n needs at a known address: row summary -> pg[0..3], children mask ->
    // pg[4..], pi row -> the (dead) S row; consumed at the next service
    auto prefetch_node = [&](int n) {
        const size_t s = (size_t)b * T + n;
        if (sub == 0) cp16(pg_addr, t.aux + s);
        if (sub == 1 && scan_ok)
            for (int w = 0; w < KW; w++) cp8(pg_addr + 16u + 8u * w, t.kids + s * KW + w);
        const float4 *row = reinterpret_cast<const float4 *>(t.pi + s * t.AP);
#pragma unroll
        for (int k = 0; k < CPL; k++) {
            const int c = L * k + sub;
            if (c < NCH && c < nrow4) cp16(ps_addr + 16u * c, row + c);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    if (b >= 0) {
        const bl_node root = bl_ld_node(t.node + (size_t)b * T);
        c_puct = bl_h2f(t.c_puct[b]);
        cur_seat = root.seat;
        if (root.terminal) state = ST_DONE;           // a terminal root ends the descent at the first service (leaf = 0, no action)
        else { state = ST_VISIT; prefetch_node(0); }
    }

    while (true) {
        const bool pass0 = state == ST_PASS || state == ST_FINAL;
        const unsigned livem = __ballot_sync(FULL, state != ST_IDLE), passm = __ballot_sync(FULL, pass0);
        if (livem == 0) break;
        const unsigned needm = livem & ~passm;
        if (passm == 0 || __popc(needm) * gate_den >= __popc(livem) * gate_num) {
            // ---- sample: l = #{a < A : sum[a] < r} over the running sums (descend_kernel, cuda.cu:160-176; see descend.cu) ----
            if (state == ST_SAMPLE) {
                // the sums are non-decreasing (every term is >= 0; entries past A repeat the last sum, so they only count when
                // every sum is below r and the answer is last_nz anyway): whole chunks below r are counted by their last entry,
                // four-way split over the lanes, then the one boundary chunk is counted entry by entry
                int cb = 0;
#pragma unroll
                for (int k = 0; k < CPL; k++) {
                    const int c = L * k + sub;
                    if (c < NCH) cb += lds1(ps_addr + 16u * c + 12u) < r ? 1 : 0;
                }
#pragma unroll
                for (int o = 1; o < L; o <<= 1) cb += __shfl_xor_sync(gmask, cb, o);
                int cnt = 4 * cb;
                if (cb < NCH) {
                    const float4 v = lds4(ps_addr + 16u * cb);
                    cnt += (v.x < r) + (v.y < r) + (v.z < r);              // v.w >= r: the chunk is not wholly below
                }
                const int l = cnt, first_nz = nzpos & 255, last_nz = (nzpos >> 8) & 255;
                action = first_nz == 255 ? -1 : (l < A ? (r <= 0.f ? first_nz : l) : last_nz);
                state = ST_ADVANCE;
            }
            // ---- advance: step to the chosen child; its row starts travelling ----
            if (state == ST_ADVANCE) {
                parent = cur;
                unsigned found = 0;
                for (int j = 0; j < nown; j++) {
                    const ChildEntry e = get(L * j + sub);
                    if (e.a == action) found = 0x80000000u | ((unsigned)e.id << 16) | ((unsigned)e.flags & 0xffffu);
                }
#pragma unroll
                for (int o = 1; o < L; o <<= 1) found |= __shfl_xor_sync(gmask, found, o);
                const int next = (found >> 31) ? (int)((found >> 16) & 0x7fffu) : -1, nflags = (int)(found & 0xffffu);
                cur = action >= 0 ? next : -1;
                if (cur >= 0 && !(nflags >> 8)) { cur_seat = nflags & 255; state = ST_VISIT; prefetch_node(cur); }
                else state = ST_DONE;                               // new leaf, existing terminal child, or no legal action
            }
            // ---- done: the descent's result (existing terminal child or -1: the expand step decides) ----
            if (state == ST_DONE) {
                if (sub == 0) {
                    t.leaf[b] = (int16_t)cur;
                    t.leaf_parent[b] = (int16_t)parent;
                    t.leaf_action[b] = (int16_t)action;
                }
                res_leaf = cur; res_parent = parent; res_action = action;
                state = ST_IDLE;
            }
            // ---- visit: children, N, lambda, random number, row slice into registers ----
            const bool visit = state == ST_VISIT;
            if (visit) asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();                                           // the group's copies have landed for all four lanes
            if (visit) {
                const size_t node0 = (size_t)b * T;
                const int seat = cur_seat;
                if (rands) r = bl_h2f(rands[node0 + cur]);
                else r = bl_uniform_half_grid(bl_philox(seed ^ (t.counters[C_MOVE] * 0x9E3779B97F4A7C15ull), (uint64_t)b,
                                                        ((uint64_t)sim << 32) | (uint32_t)cur).x);
                bl_aux ax;
                { union { float4 f; bl_aux a; } x; x.f = lds4(pg_addr); ax = x.a; }
                int N = 0;
                nown = 0;
                auto adopt = [&](const bl_node &ch, int id) {
                    put(L * nown + sub, ChildEntry{qn.fast(seat ? ch.w[1] : ch.w[0], ch.n), 0.f, (int)ch.relation, id,
                                                   (int)ch.seat | ((int)ch.terminal << 8)});
                    N += ch.n;
                    nown++;
                };
                if (scan_ok) {
                    // this lane's children = the mask bits whose node id is sub mod L; their records are fetched together
                    // straight into the lane's entry slots, then adopted in place
                    uint32_t mm[MW];
#pragma unroll
                    for (int w = 0; w < MW; w++) mm[w] = w < KW32 ? ldsu(pg_addr + 16u + 4u * w) & (OWN << sub) : 0u;
                    int j = 0;
#pragma unroll
                    for (int w = 0; w < MW; w++)
                        for (uint32_t m = mm[w]; m; m &= m - 1) {
                            const int id = w * 32 + __ffs((int)m) - 1, s = L * j + sub;
                            if (s < KS) cp16(pe_addr + 16u * s, t.node + node0 + id);
                            j++;
                        }
                    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
#pragma unroll
                    for (int w = 0; w < MW; w++)
                        for (uint32_t m = mm[w]; m; m &= m - 1) {
                            const int id = w * 32 + __ffs((int)m) - 1, s = L * nown + sub;
                            bl_node ch;
                            if (s < KS) { union { float4 f; bl_node n; } x; x.f = lds4(pe_addr + 16u * s); ch = x.n; }
                            else ch = bl_ld_node(t.node + node0 + id);
                            adopt(ch, id);
                        }
                } else {
                    const bl_node nd = bl_ld_node(t.node + node0 + cur);
                    for (int c = nd.first_child; c >= 0;) {
                        const bl_node ch = bl_ld_node(t.node + node0 + c);
                        if ((c & (L - 1)) == sub) adopt(ch, c);
                        c = ch.next_sib;
                    }
                }
                nc = nown;
#pragma unroll
                for (int o = 1; o < L; o <<= 1) { N += __shfl_xor_sync(gmask, N, o); nc += __shfl_xor_sync(gmask, nc, o); }
                N += A - nc;                                        // every child-less action counts 1 (cuda.cu:91)
                const float lambda = bl_lambda(c_puct, N, A);
                nzpos = (uint32_t)ax.first_nz | ((uint32_t)ax.last_nz << 8);
                // alpha seed (newton_search, cuda.cu:44-50): max_a (q[a] + max(lambda*pi[a], 1e-4)); the child-less part is
                // max(RN(lambda*max_pi), 1e-4) (rounding is monotone); max is exact, so the group reduction is order-free
                float alpha0 = fmaxf(__fmul_rn(lambda, ax.max_pi), 1.e-4f);
                for (int j = 0; j < nown; j++) {
                    ChildEntry e = get(L * j + sub);
                    e.top = __fmul_rn(lambda, lds1(ps_addr + 4u * e.a));   // the landed row still holds pi
                    alpha0 = fmaxf(alpha0, __fadd_rn(e.q, fmaxf(e.top, 1.e-4f)));
                    put(L * j + sub, e);
                }
#pragma unroll
                for (int o = 1; o < L; o <<= 1) alpha0 = fmaxf(alpha0, __shfl_xor_sync(gmask, alpha0, o));
                const u64 lam2 = pk(lambda, lambda);
#pragma unroll
                for (int k = 0; k < CPL; k++) {                   // top = lambda*pi, this lane's chunks of the landed row
                    const int c = L * k + sub;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);          // chunks past the row stay 0: their terms are +0 / -0, no-ops in the sums
                    if (c < NCH && c < nrow4) v = lds4(ps_addr + 16u * c);
                    tp[2 * k] = mul2(pk(v.x, v.y), lam2); tp[2 * k + 1] = mul2(pk(v.z, v.w), lam2);
                }
                const bool tiny = __fmul_rn(lambda, bl_minnz(ax)) < BL_TINY;
                alpha = alpha0; it = 0; error = BL_INF;
                state = tiny ? ST_SLOW : ST_PASS;
                if (sub == 0) { c_evals++; c_children += nc; }
            }
            __syncwarp();                                           // every read of the landed rows precedes the term phase's writes
        }

        // ---- terms of this pass: child-less form for every action, then the child terms patched over them ----
        bool pass = state == ST_PASS || state == ST_FINAL;
        if (__any_sync(FULL, pass)) {
            if (pass) {
                const float bS = alpha, bG = __fmul_rn(alpha, alpha);
                const float yS = bl_rcp_fast(bS), yG = -bl_rcp_fast(bG);   // g terms: divide lambda*pi by -(alpha^2)
                const u64 yS2 = pk(yS, yS), yG2 = pk(yG, yG), nbS2 = pk(-bS, -bS), bG2 = pk(bG, bG);
#pragma unroll
                for (int k = 0; k < CPL; k++) {
                    const int c = L * k + sub;
                    if (c < NCH) {
                        const u64 t01 = tp[2 * k], t23 = tp[2 * k + 1];
                        u64 q = mul2(t01, yS2), rr = fma2(nbS2, q, t01);
                        const u64 s01 = fma2(rr, yS2, q);
                        q = mul2(t01, yG2); rr = fma2(bG2, q, t01);
                        const u64 h01 = fma2(rr, yG2, q);
                        q = mul2(t23, yS2); rr = fma2(nbS2, q, t23);
                        const u64 s23 = fma2(rr, yS2, q);
                        q = mul2(t23, yG2); rr = fma2(bG2, q, t23);
                        const u64 h23 = fma2(rr, yG2, q);
                        sts4(ps_addr + 16u * c, make_float4(lo(s01), hi(s01), lo(s23), hi(s23)));
                        sts4(pg_addr + 16u * c, make_float4(lo(h01), hi(h01), lo(h23), hi(h23)));
                    }
                }
            }
            __syncwarp();
            bool bad = false;
            if (pass)
                for (int j = 0; j < nown; j++) {
                    const ChildEntry e = get(L * j + sub);
                    const float bot = __fsub_rn(alpha, e.q), bb = __fmul_rn(bot, bot);
                    const float sv = bl_div_fast(e.top, bot), gv = bl_div_fast(-e.top, bb);
                    sts1(ps_addr + 4u * e.a, sv);
                    sts1(pg_addr + 4u * e.a, gv);
                    // bot outside [2^-60, 2^60] leaves the branch-free division's safe range; a negative / non-finite term would
                    // break the monotone running sums: both go to the exact serial path
                    bad |= !(bot >= 8.67e-19f && bot <= 1.15e18f) || !(sv >= 0.f && sv <= 3.0e38f);
                }
            const unsigned badm = __ballot_sync(FULL, bad);
            if (pass && (badm & gmask)) state = ST_SLOW;
            __syncwarp();
        }
        // ---- exact serial fallback: the reference loops verbatim, run redundantly by the group's four lanes ----
        const bool slow = state == ST_SLOW;
        if (__any_sync(FULL, slow)) {
            if (slow) {
#pragma unroll
                for (int k = 0; k < CPL; k++) {
                    const int c = L * k + sub;
                    if (c < NCH) sts4(ps_addr + 16u * c, make_float4(lo(tp[2 * k]), hi(tp[2 * k]), lo(tp[2 * k + 1]), hi(tp[2 * k + 1])));
                }
            }
            __syncwarp();
            if (slow) {
                auto topf = [&](int a) { return lds1(ps_addr + 4u * a); };
                auto qf = [&](int a) {
                    float q = 0.f;
                    for (int s = 0; s < L; s++) {
                        const int n_s = __shfl_sync(gmask, nown, gl + s);
                        for (int j = 0; j < n_s; j++) { const ChildEntry e = get(L * j + s); if (e.a == a) q = e.q; }
                    }
                    return q;
                };
                int iters;
                const float al = bl_newton_f(topf, qf, A, &iters);
                action = bl_sample_f(topf, qf, A, al, r);
                if (sub == 0) c_iters += iters;
                state = ST_ADVANCE;
            }
            __syncwarp();
        }

        // ---- one Newton pass: the two sequential sums, S on the group's lane 0 and g on lane 1 ----
        pass = state == ST_PASS || state == ST_FINAL;
        if (__any_sync(FULL, pass)) {
            float acc = 0.f;
            if (pass && sub < 2) {
                // loads run AHEAD chunks in front of the additions; the S lane writes the running sums back over its terms
                const uint32_t row = sub ? pg_addr : ps_addr;
                constexpr int AHEAD = NCH < 4 ? NCH : 4;
                float4 buf[AHEAD];
#pragma unroll
                for (int i = 0; i < AHEAD; i++) buf[i] = lds4(row + 16u * i);
#pragma unroll
                for (int c = 0; c < NCH; c++) {
                    const float4 v = buf[c % AHEAD];
                    if (c + AHEAD < NCH) buf[c % AHEAD] = lds4(row + 16u * (c + AHEAD));
                    acc = __fadd_rn(acc, v.x); const float o0 = acc;
                    acc = __fadd_rn(acc, v.y); const float o1 = acc;
                    acc = __fadd_rn(acc, v.z); const float o2 = acc;
                    acc = __fadd_rn(acc, v.w); const float o3 = acc;
                    if (sub == 0) sts4(row + 16u * c, make_float4(o0, o1, o2, o3));
                }
            }
            __syncwarp();
            const float accS = __shfl_sync(FULL, acc, gl), accG = __shfl_sync(FULL, acc, gl | 1);
            // ---- Newton update (newton_search, cuda.cu:57-66) ----
            if (pass) {
                if (state == ST_PASS) {
                    it++;
                    if (sub == 0) c_iters++;
                    const float ne = __fsub_rn(accS, 1.f);
                    if ((ne < 1e-3f) || (error == ne)) state = ST_SAMPLE;
                    else {
                        alpha = __fsub_rn(alpha, __fdiv_rn(ne, accG));
                        error = ne;
                        if (it == 100) state = ST_FINAL;            // loop bound hit: one more pass with the last alpha, no test
                    }
                } else {
                    state = ST_SAMPLE;
                }
            }
        }
    }
    // ---- expand + env step of the group's env on its lane 0, in the env's (now dead) rows ----
    if (fuse_expand && b >= 0 && sub == 0)
        bl_expand_one(t, sim, b, res_leaf, res_parent, res_action, reinterpret_cast<uint32_t *>(reinterpret_cast<float *>(smem4) + slot * (2 * PS)),
                      reinterpret_cast<uint8_t *>(reinterpret_cast<float *>(smem4) + slot * (2 * PS) + PS));
    bl_count(t.counters, C_EVALS, c_evals);
    bl_count(t.counters, C_CHILDREN, c_children);
    bl_count(t.counters, C_ITERS, c_iters);
    bl_count(t.counters, C_DESCENTS, (b >= 0 && sub == 0) ? 1u : 0u);
}

int g_mw_gate_num = 1, g_mw_gate_den = 2, g_mw_fuse = 1, g_mw_lanes = 0;     // lanes per env: 0 = 2 up to 9x9, 4 above (measured)
void read_mw_env() {
    static bool done = false;
    if (done) return;
    done = true;
    if (const char *e = getenv("BL_MW_GATE")) {
        int a = 0, b = 0;
        if (sscanf(e, "%d/%d", &a, &b) == 2 && a >= 0 && b > 0) { g_mw_gate_num = a; g_mw_gate_den = b; }
    }
    if (const char *e = getenv("BL_MW_FUSE")) g_mw_fuse = atoi(e);
    if (const char *e = getenv("BL_MW_LANES")) g_mw_lanes = atoi(e) == 4 ? 4 : 2;
}

template <int NCH, int L>
int launch_mwl(const bl_tree *t, int sim, const bl_half *rands, uint64_t seed, cudaStream_t st) {
    using C = MwCfg<NCH, L>;
    static bool ready = false;
    if (!ready) {
        if (C::SMEM > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(descend_mw_kernel<NCH, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
            if (e != cudaSuccess) return (int)e;
        }
        ready = true;
    }
    const int cap = bl_mw_child_cap(t);
    const int64_t envs = ((int64_t)t->B + 31) / 32 * 32;
    if (envs * cap * (int64_t)sizeof(ChildEntry) > t->scratch_bytes) return -3;
    const bool fused = g_mw_fuse && t->BP <= 16 * NCH;               // the env's rows hold a board + flood-fill stack
    descend_mw_kernel<NCH, L><<<(unsigned)(envs / 32), 32 * L, C::SMEM, st>>>(*t, sim, rands, seed, reinterpret_cast<ChildEntry *>(t->scratch), cap,
                                                                       g_mw_gate_num, g_mw_gate_den, fused ? 1 : 0);
    if (cudaError_t e = cudaGetLastError()) return (int)e;
    return fused ? 0 : bl_expand_step(t, sim, st);
}

template <int NCH>
int launch_mw(const bl_tree *t, int sim, const bl_half *rands, uint64_t seed, cudaStream_t st) {
    const int lanes = g_mw_lanes ? g_mw_lanes : (t->A > 81 ? 4 : 2);
    return lanes == 4 ? launch_mwl<NCH, 4>(t, sim, rands, seed, st) : launch_mwl<NCH, 2>(t, sim, rands, seed, st);
}

}  // namespace

int bl_mw_child_cap(const bl_tree *t) { return (t->T + 3) & ~3; }
int64_t bl_mw_scratch_bytes(const bl_tree *t) {
    return (((int64_t)t->B + 31) / 32 * 32) * bl_mw_child_cap(t) * (int64_t)sizeof(ChildEntry);
}

int bl_descend_mw(const bl_tree *t, int sim, const bl_half *rands, uint64_t seed, cudaStream_t st) {
    read_mw_env();
    if (t->A > 255) return -2;                                      // entries pack the action into 8 bits
    const int nch = (t->A + 3) / 4;
    if (nch <= 3) return launch_mw<3>(t, sim, rands, seed, st);
    if (nch <= 7) return launch_mw<7>(t, sim, rands, seed, st);
    if (nch <= 13) return launch_mw<13>(t, sim, rands, seed, st);
    if (nch <= 21) return launch_mw<21>(t, sim, rands, seed, st);
    if (nch <= 31) return launch_mw<31>(t, sim, rands, seed, st);
    if (nch <= 43) return launch_mw<43>(t, sim, rands, seed, st);
    return -2;
}
