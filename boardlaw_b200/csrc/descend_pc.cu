// Tree descent with the Newton passes and the node services on DIFFERENT warps of a CTA (variant 4) — same arithmetic as
// descend.cu / descend_mw.cu (boardlaw/mcts/cpp/cuda.cu:35-99,138-182), different schedule.
//
// Why: in the other two kernels a warp's lanes run in lock step through "service" (sample the action, step to the child, fetch
// its row and its children, seed alpha) and "pass" (one Newton pass) phases, so an env that only needs a pass waits while
// its warp pays the service latency of its neighbours, and the other way round (profiles/r01_descend_mw_ncu_full.txt: 58 trips
// per warp where an env needs 22.8 passes + 6.2 services).  Here a CTA of 96 threads owns 32 envs:
//
//   * warps 0 and 1 are PASS warps: two lanes per env (16 envs each), exactly the term / child-term / chain / Newton-update
//     phases of descend_mw.cu at L = 2, and nothing else;
//   * warp 2 is the SERVICE warp: lane e serves env e — inverse-CDF sample over the running sums, step to the child, fetch
//     (cp.async) and visit the next node: children adopted, N, lambda, the children's tops, the alpha seed;
//   * an env is handed back and forth through one word of shared memory (`st`: who owns the env's rows), written after a
//     block-level fence.  The rows (landed pi row / terms / running sums, child entries) are only touched by the owner.
//
// A pass then costs one trip of a warp that only does passes, and a service never stalls a pass.
// EXPERIMENTAL (bl_debug_set_descend_variant(4) / BL_DESCEND_VARIANT=4): bit-exact, but measured slower than descend.cu on c2
// (19.7 vs 12.8 ms per move): one service warp per 32 envs serialises the services, which turn out to be the larger half of an
// env's critical path.  Kept as the starting point for splitting the service side further.  Every poll loop is bounded
// (a stuck hand-over raises the tree's error counter and ends the descent instead of hanging the GPU).
//
// Compiled with -fmad=false -prec-div=true -ftz=false (see build.py); fused operations are explicit.
#include <cstdio>
#include <cstdlib>

#include "descend_common.cuh"

namespace {

enum { OWN_SERVICE = 0, OWN_PASS = 1, OWN_DONE = 2, OWN_SLOW = 3 };
enum { SV_WAITPASS = 0, SV_SAMPLE = 1, SV_ADVANCE = 2, SV_VISIT = 3, SV_DONE = 4, SV_FINISHED = 5, SV_SLOW = 6 };

template <int NCH>
struct PcCfg {
    static constexpr int PS = 4 * NCH;                 // floats per row
    static constexpr int CPL = (NCH + 1) / 2;          // chunks per pass lane
    static constexpr int KS = NCH - 2 > 4 ? NCH - 2 : 4;   // child entries per env in shared memory (two fewer than the other kernels: the control words)
    static constexpr int ENVS = 32;
    static constexpr int ROWS_BYTES = ENVS * 2 * PS * 4;
    static constexpr int ENT_BYTES = ENVS * 16 * KS;
    static constexpr int CTL_WORDS = 8;                // per env: owner, alpha0, lambda, nc, (spare)
    static constexpr int SMEM = ROWS_BYTES + ENT_BYTES + ENVS * CTL_WORDS * 4;
    static constexpr int FIT = 233472 / (SMEM + 1024);
    static constexpr int MINB = FIT > 7 ? 7 : (FIT < 1 ? 1 : FIT);
};

__device__ __forceinline__ void pc_cp16(uint32_t dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void pc_cp8(uint32_t dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ float4 pc_lds4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void pc_sts4(uint32_t a, const float4 &v) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float pc_lds1(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void pc_sts1(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ uint32_t pc_ldsu(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
// hand-over word: volatile accesses, ordered against the data by __threadfence_block()
__device__ __forceinline__ uint32_t pc_ld_own(uint32_t a) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void pc_st_own(uint32_t a, uint32_t v) {
    __threadfence_block();
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t pc_opaque(uint32_t x) {
    uint32_t y;
    asm volatile("mov.u32 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}

constexpr int PC_POLL_LIMIT = 1 << 21;                 // polls of ~100 ns before a hand-over is declared stuck

template <int NCH>
__global__ void __launch_bounds__(96, PcCfg<NCH>::MINB)
descend_pc_kernel(bl_tree t, int sim, const bl_half *__restrict__ rands, uint64_t seed, ChildEntry *__restrict__ clists, int cap,
                  int fuse_expand) {
    using C = PcCfg<NCH>;
    constexpr int PS = C::PS, CPL = C::CPL, KS = C::KS;
    constexpr int MW = 8;                              // 32-bit words of the children-of-this-node mask (T <= 256; else list walk)
    extern __shared__ float4 smem4[];
    const int A = t.A, T = t.T;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem0 = smem_u32(smem4);
    const int nrow4 = t.AP >> 2;
    // every env starts on the service side
    if (threadIdx.x < C::ENVS)
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(smem0 + C::ROWS_BYTES + C::ENT_BYTES + threadIdx.x * (C::CTL_WORDS * 4)), "r"((uint32_t)OWN_SERVICE) : "memory");
    __syncthreads();

    if (warp < 2) {
        // ================================================= PASS warps =====================================================
        const int sub = lane & 1, gl = lane & ~1;
        const int slot = warp * 16 + (lane >> 1);
        const uint32_t ps_addr = pc_opaque(smem0 + (uint32_t)slot * (2 * PS * 4)), pg_addr = ps_addr + PS * 4;
        const uint32_t pe_addr = pc_opaque(smem0 + C::ROWS_BYTES + (uint32_t)slot * (16 * KS));
        const uint32_t ctl = smem0 + C::ROWS_BYTES + C::ENT_BYTES + (uint32_t)slot * (C::CTL_WORDS * 4);
        const int b = (int)blockIdx.x * C::ENVS + slot;
        ChildEntry *cl = clists + (size_t)(b < t.B ? b : 0) * cap;
        u64 tp[2 * CPL];
#pragma unroll
        for (int k = 0; k < 2 * CPL; k++) tp[k] = 0;
        bool active = false, over = b >= t.B;          // over: the env's descent has ended
        int it = 0, nc = 0, state = ST_PASS;
        float alpha = 1.f, error = 0.f;
        unsigned c_iters = 0;
        int idle = 0;
        auto get = [&](int i) {
            ChildEntry e;
            if (i < KS) {
                const float4 v = pc_lds4(pe_addr + 16u * i);
                e.q = v.x; e.top = v.y;
                const uint32_t u = __float_as_uint(v.z);
                e.a = u & 255; e.id = u >> 8; e.flags = __float_as_int(v.w);
            } else e = cl[i];
            return e;
        };
        while (true) {
            // ---- hand-over: take the env when the service side has a fresh evaluation ready ----
            uint32_t own = OWN_SERVICE;
            if (!active && !over && sub == 0) own = pc_ld_own(ctl);
            own = __shfl_sync(FULL, own, gl);
            if (!active && !over) {
                if (own == OWN_DONE) over = true;
                else if (own == OWN_PASS) {
                    __threadfence_block();
                    const float lambda = pc_lds1(ctl + 8);
                    alpha = pc_lds1(ctl + 4);
                    nc = (int)pc_ldsu(ctl + 12);
                    const u64 lam2 = pk(lambda, lambda);
#pragma unroll
                    for (int k = 0; k < CPL; k++) {               // top = lambda*pi, this lane's chunks of the landed row
                        const int c = 2 * k + sub;
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (c < NCH && c < nrow4) v = pc_lds4(ps_addr + 16u * c);
                        tp[2 * k] = mul2(pk(v.x, v.y), lam2); tp[2 * k + 1] = mul2(pk(v.z, v.w), lam2);
                    }
                    it = 0; error = BL_INF; state = ST_PASS;
                    active = true;
                }
            }
            if (__all_sync(FULL, over)) break;
            __syncwarp();                                           // every lane has read its pi chunks before terms overwrite the row
            if (!__any_sync(FULL, active)) {
                if (++idle > PC_POLL_LIMIT) {                      // stuck hand-over: report and leave
                    if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long *>(t.counters + C_ERRORS), 1ull);
                    break;
                }
                __nanosleep(100);
                continue;
            }
            idle = 0;
            // ---- terms of this pass (child-less form for every action), then the child terms patched over them ----
            if (active) {
                const float bS = alpha, bG = __fmul_rn(alpha, alpha);
                const float yS = bl_rcp_fast(bS), yG = -bl_rcp_fast(bG);
                const u64 yS2 = pk(yS, yS), yG2 = pk(yG, yG), nbS2 = pk(-bS, -bS), bG2 = pk(bG, bG);
#pragma unroll
                for (int k = 0; k < CPL; k++) {
                    const int c = 2 * k + sub;
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
                        pc_sts4(ps_addr + 16u * c, make_float4(lo(s01), hi(s01), lo(s23), hi(s23)));
                        pc_sts4(pg_addr + 16u * c, make_float4(lo(h01), hi(h01), lo(h23), hi(h23)));
                    }
                }
            }
            __syncwarp();
            bool bad = false;
            if (active)
                for (int i = sub; i < nc; i += 2) {
                    const ChildEntry e = get(i);
                    const float bot = __fsub_rn(alpha, e.q), bb = __fmul_rn(bot, bot);
                    const float sv = bl_div_fast(e.top, bot), gv = bl_div_fast(-e.top, bb);
                    pc_sts1(ps_addr + 4u * e.a, sv);
                    pc_sts1(pg_addr + 4u * e.a, gv);
                    bad |= !(bot >= 8.67e-19f && bot <= 1.15e18f) || !(sv >= 0.f && sv <= 3.0e38f);
                }
            const unsigned badm = __ballot_sync(FULL, bad);
            const bool slow = active && ((badm >> gl) & 3u);
            __syncwarp();
            if (slow) {                                             // exact serial path: the service lane runs the reference loops on the tops
#pragma unroll
                for (int k = 0; k < CPL; k++) {
                    const int c = 2 * k + sub;
                    if (c < NCH) pc_sts4(ps_addr + 16u * c, make_float4(lo(tp[2 * k]), hi(tp[2 * k]), lo(tp[2 * k + 1]), hi(tp[2 * k + 1])));
                }
            }
            __syncwarp();
            if (slow) {
                if (sub == 0) pc_st_own(ctl, OWN_SLOW);
                active = false;
            }
            // ---- one Newton pass: S chain on the group's lane 0, g chain on lane 1 ----
            float acc = 0.f;
            if (active) {
                const uint32_t row = sub ? pg_addr : ps_addr;
                constexpr int AHEAD = NCH < 4 ? NCH : 4;
                float4 buf[AHEAD];
#pragma unroll
                for (int i = 0; i < AHEAD; i++) buf[i] = pc_lds4(row + 16u * i);
#pragma unroll
                for (int c = 0; c < NCH; c++) {
                    const float4 v = buf[c % AHEAD];
                    if (c + AHEAD < NCH) buf[c % AHEAD] = pc_lds4(row + 16u * (c + AHEAD));
                    acc = __fadd_rn(acc, v.x); const float o0 = acc;
                    acc = __fadd_rn(acc, v.y); const float o1 = acc;
                    acc = __fadd_rn(acc, v.z); const float o2 = acc;
                    acc = __fadd_rn(acc, v.w); const float o3 = acc;
                    if (sub == 0) pc_sts4(row + 16u * c, make_float4(o0, o1, o2, o3));
                }
            }
            __syncwarp();
            const float accS = __shfl_sync(FULL, acc, gl), accG = __shfl_sync(FULL, acc, gl | 1);
            // ---- Newton update (newton_search, cuda.cu:57-66) ----
            if (active) {
                bool converged = false;
                if (state == ST_PASS) {
                    it++;
                    if (sub == 0) c_iters++;
                    const float ne = __fsub_rn(accS, 1.f);
                    if ((ne < 1e-3f) || (error == ne)) converged = true;
                    else {
                        alpha = __fsub_rn(alpha, __fdiv_rn(ne, accG));
                        error = ne;
                        if (it == 100) state = ST_FINAL;            // loop bound hit: one more pass with the last alpha, no test
                    }
                } else converged = true;
                if (converged) {                                   // the running sums are in the S row: over to the service lane
                    if (sub == 0) pc_st_own(ctl, OWN_SERVICE);
                    active = false;
                }
            }
        }
        bl_count(t.counters, C_ITERS, c_iters);
    } else {
        // ================================================ SERVICE warp ====================================================
        const int slot = lane;
        float *ps = reinterpret_cast<float *>(smem4) + slot * (2 * PS);
        float *pg = ps + PS;
        const uint32_t ps_addr = pc_opaque(smem0 + (uint32_t)slot * (2 * PS * 4)), pg_addr = ps_addr + PS * 4;
        const uint32_t pe_addr = pc_opaque(smem0 + C::ROWS_BYTES + (uint32_t)slot * (16 * KS));
        const uint32_t ctl = smem0 + C::ROWS_BYTES + C::ENT_BYTES + (uint32_t)slot * (C::CTL_WORDS * 4);
        const bl_qnorm qn(t.qrange + 2 * sim);
        const int KW = (T + 63) >> 6, KW32 = (T + 31) >> 5;
        const bool scan_ok = T <= 32 * MW;
        int b = (int)blockIdx.x * C::ENVS + slot;
        if (b >= t.B) b = -1;
        ChildEntry *cl = clists + (size_t)(b < 0 ? 0 : b) * cap;
        int cur = 0, parent = 0, action = -1, sv = SV_FINISHED, nc = 0, cur_seat = 0;
        int res_leaf = -1, res_parent = 0, res_action = -1;
        float r = 0.f, c_puct = 0.f;
        uint32_t nzpos = 0;
        unsigned c_evals = 0, c_children = 0, c_iters = 0;
        int idle = 0;
        auto get = [&](int i) {
            ChildEntry e;
            if (i < KS) {
                const float4 v = pc_lds4(pe_addr + 16u * i);
                e.q = v.x; e.top = v.y;
                const uint32_t u = __float_as_uint(v.z);
                e.a = u & 255; e.id = u >> 8; e.flags = __float_as_int(v.w);
            } else e = cl[i];
            return e;
        };
        auto put = [&](int i, const ChildEntry &e) {
            if (i < KS) pc_sts4(pe_addr + 16u * i, make_float4(e.q, e.top, __uint_as_float((uint32_t)e.a | ((uint32_t)e.id << 8)), __int_as_float(e.flags)));
            else cl[i] = e;
        };
        auto prefetch_node = [&](int n) {
            const size_t s = (size_t)b * T + n;
            pc_cp16(pg_addr, t.aux + s);                              // row summary -> pg[0..3]
            if (scan_ok)
                for (int w = 0; w < KW; w++) pc_cp8(pg_addr + 16u + 8u * w, t.kids + s * KW + w);   // children mask -> pg[4..]
            const float4 *row = reinterpret_cast<const float4 *>(t.pi + s * t.AP);
#pragma unroll
            for (int c = 0; c < NCH; c++)
                if (c < nrow4) pc_cp16(ps_addr + 16u * c, row + c);
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        if (b >= 0) {
            const bl_node root = bl_ld_node(t.node + (size_t)b * T);
            c_puct = bl_h2f(t.c_puct[b]);
            cur_seat = root.seat;
            if (root.terminal) sv = SV_DONE;
            else { sv = SV_VISIT; prefetch_node(0); }
        }
        if (b < 0) pc_st_own(ctl, OWN_DONE);
        while (true) {
            if (__all_sync(FULL, sv == SV_FINISHED)) break;
            // ---- the pass side has handed the env back? ----
            if (sv == SV_WAITPASS) {
                const uint32_t own = pc_ld_own(ctl);
                if (own == OWN_SERVICE) { __threadfence_block(); sv = SV_SAMPLE; }
                else if (own == OWN_SLOW) { __threadfence_block(); sv = SV_SLOW; }
            }
            if (!__any_sync(FULL, sv != SV_WAITPASS && sv != SV_FINISHED)) {
                if (++idle > PC_POLL_LIMIT) {
                    if (sv != SV_FINISHED) {
                        atomicAdd(reinterpret_cast<unsigned long long *>(t.counters + C_ERRORS), 1ull);
                        t.leaf[b] = -1; t.leaf_parent[b] = 0; t.leaf_action[b] = -1;
                        pc_st_own(ctl, OWN_DONE);
                        sv = SV_FINISHED;
                    }
                    continue;
                }
                __nanosleep(100);
                continue;
            }
            idle = 0;
            // ---- exact serial fallback on the tops the pass lanes left in the S row (the reference loops verbatim) ----
            if (sv == SV_SLOW) {
                auto topf = [&](int a) { return pc_lds1(ps_addr + 4u * a); };
                auto qf = [&](int a) {
                    float q = 0.f;
                    for (int i = 0; i < nc; i++) { const ChildEntry e = get(i); if (e.a == a) q = e.q; }
                    return q;
                };
                int iters;
                const float al = bl_newton_f(topf, qf, A, &iters);
                action = bl_sample_f(topf, qf, A, al, r);
                c_iters += iters;
                sv = SV_ADVANCE;
            }
            // ---- sample: l = #{a < A : sum[a] < r} over the running sums (cuda.cu:160-176): whole chunks by their last entry, then
            //      the boundary chunk (see descend_mw.cu) ----
            if (sv == SV_SAMPLE) {
                int cb = 0;
#pragma unroll
                for (int c = 0; c < NCH; c++) cb += pc_lds1(ps_addr + 16u * c + 12u) < r ? 1 : 0;
                int cnt = 4 * cb;
                if (cb < NCH) {
                    const float4 v = pc_lds4(ps_addr + 16u * cb);
                    cnt += (v.x < r) + (v.y < r) + (v.z < r);
                }
                const int l = cnt, first_nz = nzpos & 255, last_nz = (nzpos >> 8) & 255;
                action = first_nz == 255 ? -1 : (l < A ? (r <= 0.f ? first_nz : l) : last_nz);
                sv = SV_ADVANCE;
            }
            // ---- advance: step to the chosen child; its row starts travelling ----
            if (sv == SV_ADVANCE) {
                parent = cur;
                int next = -1, nflags = 0;
                for (int i = 0; i < nc; i++) {
                    const ChildEntry e = get(i);
                    if (e.a == action) { next = e.id; nflags = e.flags; }
                }
                cur = action >= 0 ? next : -1;
                if (cur >= 0 && !(nflags >> 8)) { cur_seat = nflags & 255; sv = SV_VISIT; prefetch_node(cur); }
                else sv = SV_DONE;
            }
            // ---- done ----
            if (sv == SV_DONE) {
                t.leaf[b] = (int16_t)cur;
                t.leaf_parent[b] = (int16_t)parent;
                t.leaf_action[b] = (int16_t)action;
                res_leaf = cur; res_parent = parent; res_action = action;
                pc_st_own(ctl, OWN_DONE);
                sv = SV_FINISHED;
            }
            // ---- visit: children, N, lambda, random number, the children's tops, alpha seed; then over to the pass lanes ----
            if (sv == SV_VISIT) {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                const size_t node0 = (size_t)b * T;
                const int seat = cur_seat;
                if (rands) r = bl_h2f(rands[node0 + cur]);
                else r = bl_uniform_half_grid(bl_philox(seed ^ (t.counters[C_MOVE] * 0x9E3779B97F4A7C15ull), (uint64_t)b,
                                                        ((uint64_t)sim << 32) | (uint32_t)cur).x);
                bl_aux ax;
                { union { float4 f; bl_aux a; } x; x.f = pc_lds4(pg_addr); ax = x.a; }
                int N = 0;
                nc = 0;
                auto adopt = [&](const bl_node &ch, int id) {
                    put(nc, ChildEntry{qn.fast(seat ? ch.w[1] : ch.w[0], ch.n), 0.f, (int)ch.relation, id, (int)ch.seat | ((int)ch.terminal << 8)});
                    N += ch.n;
                    nc++;
                };
                if (scan_ok) {
                    uint32_t mm[MW];
#pragma unroll
                    for (int w = 0; w < MW; w++) mm[w] = w < KW32 ? pc_ldsu(pg_addr + 16u + 4u * w) : 0u;
                    int k = 0;
#pragma unroll
                    for (int w = 0; w < MW; w++)
                        for (uint32_t m = mm[w]; m; m &= m - 1) {
                            const int id = w * 32 + __ffs((int)m) - 1;
                            if (k < KS) pc_cp16(pe_addr + 16u * k, t.node + node0 + id);
                            k++;
                        }
                    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
#pragma unroll
                    for (int w = 0; w < MW; w++)
                        for (uint32_t m = mm[w]; m; m &= m - 1) {
                            const int id = w * 32 + __ffs((int)m) - 1;
                            bl_node ch;
                            if (nc < KS) { union { float4 f; bl_node n; } x; x.f = pc_lds4(pe_addr + 16u * nc); ch = x.n; }
                            else ch = bl_ld_node(t.node + node0 + id);
                            adopt(ch, id);
                        }
                } else {
                    const bl_node nd = bl_ld_node(t.node + node0 + cur);
                    for (int c = nd.first_child; c >= 0;) {
                        const bl_node ch = bl_ld_node(t.node + node0 + c);
                        adopt(ch, c);
                        c = ch.next_sib;
                    }
                }
                N += A - nc;                                        // every child-less action counts 1 (cuda.cu:91)
                const float lambda = bl_lambda(c_puct, N, A);
                nzpos = (uint32_t)ax.first_nz | ((uint32_t)ax.last_nz << 8);
                float alpha0 = fmaxf(__fmul_rn(lambda, ax.max_pi), 1.e-4f);
                for (int i = 0; i < nc; i++) {
                    ChildEntry e = get(i);
                    e.top = __fmul_rn(lambda, pc_lds1(ps_addr + 4u * e.a));   // the landed row holds pi
                    alpha0 = fmaxf(alpha0, __fadd_rn(e.q, fmaxf(e.top, 1.e-4f)));
                    put(i, e);
                }
                const bool tiny = __fmul_rn(lambda, bl_minnz(ax)) < BL_TINY;
                c_evals++; c_children += nc;
                if (tiny) {
                    // rows with denormal-range values: the exact serial path, here (tops into the S row in place)
                    for (int a = 0; a < 4 * NCH; a++) pc_sts1(ps_addr + 4u * a, a < t.AP ? __fmul_rn(lambda, pc_lds1(ps_addr + 4u * a)) : 0.f);
                    sv = SV_SLOW;
                } else {
                    pc_sts1(ctl + 4, alpha0);
                    pc_sts1(ctl + 8, lambda);
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(ctl + 12), "r"((uint32_t)nc) : "memory");
                    pc_st_own(ctl, OWN_PASS);
                    sv = SV_WAITPASS;
                }
            }
        }
        // ---- expand + env step of the CTA's envs, one lane per env, in the env's (now dead) rows.  The pass warps may still be
        //      polling other envs' words; they never touch the rows of an env whose word says OWN_DONE. ----
        if (fuse_expand && b >= 0) bl_expand_one(t, sim, b, res_leaf, res_parent, res_action, reinterpret_cast<uint32_t *>(ps), reinterpret_cast<uint8_t *>(pg));
        bl_count(t.counters, C_EVALS, c_evals);
        bl_count(t.counters, C_CHILDREN, c_children);
        bl_count(t.counters, C_ITERS, c_iters);
        bl_count(t.counters, C_DESCENTS, b >= 0 ? 1u : 0u);
    }
}

template <int NCH>
int launch_pc(const bl_tree *t, int sim, const bl_half *rands, uint64_t seed, cudaStream_t st) {
    using C = PcCfg<NCH>;
    static bool ready = false;
    if (!ready) {
        if (C::SMEM > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(descend_pc_kernel<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
            if (e != cudaSuccess) return (int)e;
        }
        ready = true;
    }
    const int cap = bl_mw_child_cap(t);
    const int64_t envs = ((int64_t)t->B + 31) / 32 * 32;
    if (envs * cap * (int64_t)sizeof(ChildEntry) > t->scratch_bytes) return -3;
    const bool fused = t->BP <= 16 * NCH;
    descend_pc_kernel<NCH><<<(unsigned)(envs / 32), 96, C::SMEM, st>>>(*t, sim, rands, seed, reinterpret_cast<ChildEntry *>(t->scratch), cap, fused ? 1 : 0);
    if (cudaError_t e = cudaGetLastError()) return (int)e;
    return fused ? 0 : bl_expand_step(t, sim, st);
}

}  // namespace

int bl_descend_pc(const bl_tree *t, int sim, const bl_half *rands, uint64_t seed, cudaStream_t st) {
    if (t->A > 255) return -2;
    const int nch = (t->A + 3) / 4;
    if (nch <= 3) return launch_pc<3>(t, sim, rands, seed, st);
    if (nch <= 7) return launch_pc<7>(t, sim, rands, seed, st);
    if (nch <= 13) return launch_pc<13>(t, sim, rands, seed, st);
    if (nch <= 21) return launch_pc<21>(t, sim, rands, seed, st);
    if (nch <= 31) return launch_pc<31>(t, sim, rands, seed, st);
    if (nch <= 43) return launch_pc<43>(t, sim, rands, seed, st);
    return -2;
}
