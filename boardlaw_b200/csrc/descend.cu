// Task-parallel tree descent for sm_100a — the hot loop of MCTS.simulate (descend_kernel + policy + newton_search,
// boardlaw/mcts/cpp/cuda.cu:35-99,138-182).
//
// The reference's arithmetic is a chain of dependent fp32 additions in the order a = 0..A-1 (S = sum lambda*pi/(alpha-q),
// g = sum -lambda*pi/(alpha-q)^2, repeated per Newton iteration, then once more for the inverse-CDF sample).  Bit-exact
// parity forbids re-association, so the kernel is built around making that chain cheap rather than around bandwidth:
//
//   * one lane per env, both chains in the lane (two independent dependency chains), the node's lambda*pi row held in
//     REGISTERS for the whole evaluation (kernel templated on the row length in 4-element chunks; shared memory could not
//     feed 4 sub-partitions at the fp32 pipe's rate);
//   * every lane is its own state machine: each trip round the main loop advances the lane by ONE Newton pass, whatever
//     node or iteration it is at; a lane whose evaluation converged samples, steps to the child and loads the next row
//     in the same trip; a lane whose descent is over takes the next env from a global queue;
//   * division by a divisor shared by every child-less action (q = 0 => bot = alpha): y = RN(1/b), q0 = RN(n*y),
//     r = n - b*q0 (exact, one FMA), RN(q0 + r*y) is the correctly rounded n/b (Markstein) as long as nothing
//     under/overflows — rows holding a nonzero lambda*pi below 2^-100 go to the serial fallback that runs the reference
//     loops verbatim.  The three operations are issued as packed FMUL2/FFMA2 over element pairs;
//   * actions that do have a child (1.8 per node on average) take full IEEE divisions once per pass, parked in two
//     lane-private shared-memory rows and selected in at their positions (flag = sign bit of the register row);
//   * the S chain's running sums are the sampling loop's `total`: they are stored as the pass goes and binary-searched
//     when the pass turns out to be the last one.
//
// Compiled with -fmad=false -prec-div=true -ftz=false (see build.py); fused operations are explicit.
#include <cstdio>
#include <cstdlib>


#include "descend_common.cuh"

namespace {


constexpr int XNT = 128;
__global__ void __launch_bounds__(XNT) expand_step_kernel(bl_tree t, int sim) {
    extern __shared__ __align__(16) uint8_t raw[];
    const int tid = threadIdx.x, pw = (t.BP >> 2) | 1;            // row pitch in words (odd)
    uint32_t *bdw = reinterpret_cast<uint32_t *>(raw) + (size_t)tid * pw;
    uint8_t *stk = raw + (size_t)XNT * pw * 4 + (size_t)tid * pw * 4;
    const int b = blockIdx.x * XNT + tid;
    if (b >= t.B) return;
    bl_expand_one(t, sim, b, t.leaf[b], t.leaf_parent[b], t.leaf_action[b], bdw, stk);
}

// service gate: the visit / sample / advance phases run when at least GATE_NUM/GATE_DEN of the warp's live lanes wait for them
// (or nobody is in a pass); a lane therefore idles a trip or two now and then, and the warp does not pay the phases'
// latency on every trip
#ifndef BL_GATE_NUM
#define BL_GATE_NUM 1
#define BL_GATE_DEN 2      /* measured on c2: 1/2 17.3 ms, 1/3 18.5, 2/3 18.4, 1/4 19.6, every trip 23.4 ms per move */
#endif

template <int NCH, bool PROF>
__global__ void __launch_bounds__(32, 7) descend_v3_kernel(bl_tree t, int sim, const bl_half *__restrict__ rands, uint64_t seed,
                                                        ChildEntry *__restrict__ clists, int cap, unsigned long long *prof,
                                                        int gate_num, int gate_den, int fuse_expand) {
    constexpr int PS = 4 * NCH;                         // row pitch in floats; NCH odd => conflict-free 128-bit lane-private rows
    // the lane's third shared-memory row holds its child entries (16 B each; the rest go to global scratch)
    constexpr int NW = (PS + 63) / 64;                  // 64-bit words of the child-position mask
    constexpr int MW = 4;                               // 64-bit words of the children-of-this-node mask (T <= 256; else list walk)
    constexpr int SEG = PS > 144 ? 14 : (PS > 100 ? 12 : (PS > 64 ? 10 : (PS > 36 ? 8 : (PS > 16 ? 6 : 4))));   // SEG*SEG >= PS >= A
    extern __shared__ float4 smem4[];
    const int A = t.A, T = t.T;
    const int lane = threadIdx.x;
    float *ps = reinterpret_cast<float *>(smem4) + lane * PS;       // S child terms in, running S sums out
    float *pg = ps + 32 * PS;                                       // g child terms
    float4 *pe4 = reinterpret_cast<float4 *>(pg + 32 * PS);         // child entries {q, top, action | id << 8, -}; raw node records while in flight
    float4 *ps4 = reinterpret_cast<float4 *>(ps);
    const uint32_t ps_addr = smem_u32(ps), pg_addr = smem_u32(pg), pe_addr = smem_u32(pe4);
    ChildEntry *cl = clists + ((size_t)blockIdx.x * 32 + lane) * cap;
    const bl_qnorm qn(t.qrange + 2 * sim);
    const uint64_t move = t.counters[C_MOVE];
    const uint64_t keep = bl_policy_keep();
    int *queue = reinterpret_cast<int *>(t.counters + C_QUEUE);
    const int nrow4 = t.AP >> 2;
    const int KW = (T + 63) >> 6;                     // 64-bit words of a node's children mask (t.kids)
    const bool scan_ok = T <= 64 * MW;
    constexpr int KS = NCH;                           // child entries held in the lane's third shared-memory row; the rest spill
                                                      // to global scratch

    u64 tp[2 * NCH];                                  // lambda*pi of the current node, element pairs
    u64 cm[NW];                                       // bit a set: action a has a child
    int b = -1, cur = -1, parent = 0, action = -1, state = ST_DONE, nc = 0, it = 0, cur_seat = 0;
    int res_leaf = -1, res_parent = 0, res_action = -1;       // this lane's finished descent (one env per lane when all are resident)
    float alpha = 1.f, error = 0.f, r = 0.f, c_puct = 0.f;
    uint32_t nzpos = 0;                               // first_nz | last_nz << 8 of the current row
    bool exhausted = false;                           // warp-uniform: the queue has nothing left
    const bool all_resident = (long long)gridDim.x * 32 >= t.B;   // every env has a lane: no queue, env = global lane index
    unsigned c_evals = 0, c_children = 0, c_iters = 0, c_desc = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) cm[w] = 0;
    // optional phase clock (bl_debug_set_phase_profile): cycles per phase summed over warps, for DESIGN.md's latency budget
    long long pc[PROF ? 13 : 1], tlast = PROF ? clock64() : 0;
    unsigned n_service = 0, n_pass = 0;               // PROF: trips with the service block / with a pass (slots 13, 14)
#pragma unroll
    for (int k = 0; k < (PROF ? 13 : 1); k++) pc[k] = 0;
#define tick(k) do { if (PROF) { __syncwarp(__activemask()); const long long now_ = clock64(); pc[k] += now_ - tlast; tlast = now_; } } while (0)

    auto get = [&](int i) {
        ChildEntry e;
        if (i < KS) {
            const float4 v = pe4[i];
            e.q = v.x; e.top = v.y;
            const uint32_t u = __float_as_uint(v.z);
            e.a = u & 255; e.id = u >> 8; e.flags = __float_as_int(v.w);
        } else e = cl[i];
        return e;
    };
    auto put = [&](int i, const ChildEntry &e) {
        if (i < KS) pe4[i] = make_float4(e.q, e.top, __uint_as_float((uint32_t)e.a | ((uint32_t)e.id << 8)), __int_as_float(e.flags));
        else cl[i] = e;
    };

    // asynchronous fetch of everything a visit of node `n` needs that has a known address: its record -> entry slot 0, its
    // row summary -> slot 1, its pi row -> the (dead) sums row.  Issued when the descent steps to the node; consumed at the
    // next service, a trip or more later, so the DRAM latency is spent while the other lanes run their passes.
    auto prefetch_node = [&](int n) {
        const size_t slot = (size_t)b * T + n;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(pg_addr), "l"(t.aux + slot) : "memory");       // row summary -> pg[0..3]
        if (scan_ok)                                                                                                    // children mask -> pg[4..]
            for (int w = 0; w < KW; w++)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(pg_addr + 16u + 8u * w), "l"(t.kids + slot * KW + w) : "memory");
        const float4 *row = reinterpret_cast<const float4 *>(t.pi + slot * t.AP);
#pragma unroll
        for (int c = 0; c < NCH; c++)
            if (c < nrow4) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(ps_addr + 16u * c), "l"(row + c) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    while (true) {
        const bool pass0 = state == ST_PASS || state == ST_FINAL;
        const unsigned livem = __ballot_sync(FULL, state != ST_IDLE), passm = __ballot_sync(FULL, pass0);
        if (livem == 0) break;
        tick(0);
        const unsigned needm = livem & ~passm;
        if (passm == 0 || __popc(needm) * gate_den >= __popc(livem) * gate_num) {
            if (PROF) n_service++;
            // ---- G: inverse-CDF search over the running sums (descend_kernel, cuda.cu:160-176) ----------------------------------
            if (state == ST_SAMPLE) {
                // every term is >= 0 (checked for child terms in C), so the sums are non-decreasing: first index with sum >= r.
                // The reference additionally skips p == 0 entries: the first hit can only have p == 0 when r == 0 (then the
                // answer is the first nonzero entry), and when no sum reaches r the answer is the last nonzero entry.
                // l = #{a : sum[a] < r}, counted in two rounds of INDEPENDENT loads (segment ends, then the hit segment's entries)
                // instead of a binary search's log2(A) dependent ones
                int c1 = 0;
#pragma unroll
                for (int j = 0; j < SEG; j++) {
                    const int e = (j + 1) * SEG < A ? (j + 1) * SEG : A;        // exclusive end of segment j (empty ones repeat A)
                    c1 += (j * SEG < A && ps[e - 1] < r) ? 1 : 0;
                }
                int l = c1 * SEG;
                if (l < A) {
                    int c2 = 0;
#pragma unroll
                    for (int j = 0; j < SEG - 1; j++) c2 += (l + j < A && ps[l + j] < r) ? 1 : 0;   // the segment's last entry is >= r
                    l += c2;
                } else l = A;
                const int first_nz = nzpos & 255, last_nz = (nzpos >> 8) & 255;
                action = first_nz == 255 ? -1 : (l < A ? (r <= 0.f ? first_nz : l) : last_nz);
                state = ST_ADVANCE;
            }
            // ---- H: step to the chosen child (its seat / terminal flag were read with its statistics); its row starts travelling ----
            if (state == ST_ADVANCE) {
                parent = cur;
                int next = -1, nflags = 0;
                for (int i = 0; i < nc; i++) {
                    const ChildEntry e = get(i);
                    if (e.a == action) { next = e.id; nflags = e.flags; }
                }
                cur = action >= 0 ? next : -1;
                if (cur >= 0 && !(nflags >> 8)) { cur_seat = nflags & 255; state = ST_VISIT; prefetch_node(cur); }
                else state = ST_DONE;                               // new leaf, existing terminal child, or no legal action
            }
            tick(1);
            // ---- A: finished descents write their result and take the next env -------------------------------------------
            const bool fresh = state == ST_DONE;
            if (fresh && b >= 0) {
                t.leaf[b] = (int16_t)cur;                          // existing terminal child, or -1: the expand step decides
                t.leaf_parent[b] = (int16_t)parent;
                t.leaf_action[b] = (int16_t)action;
                res_leaf = cur; res_parent = parent; res_action = action;
                c_desc++;
            }
            const unsigned req = __ballot_sync(FULL, fresh);
            if (req) {
                int base = t.B;
                if (!exhausted && !all_resident) {
                    if (lane == 0) base = atomicAdd(queue, __popc(req));
                    base = __shfl_sync(FULL, base, 0);
                    exhausted = base + __popc(req) >= t.B;
                }
                if (fresh) {
                    int nb = base + __popc(req & ((1u << lane) - 1u));
                    if (all_resident) nb = (b < 0 && !exhausted) ? (int)blockIdx.x * 32 + lane : t.B;
                    b = nb < t.B ? nb : -1;
                    cur = 0; parent = 0; action = -1;
                    state = ST_IDLE;
                    if (b >= 0) {
                        const bl_node root = bl_ld_node_hint(t.node + (size_t)b * T, keep);
                        c_puct = bl_h2f(t.c_puct[b]);
                        cur_seat = root.seat;
                        if (root.terminal) state = ST_DONE;       // a terminal root ends the descent at the next service (leaf = 0, no action)
                        else {
                            state = ST_VISIT;
                            prefetch_node(0);
                        }
                    }
                }
                if (all_resident) exhausted = true;
            }
            tick(2);
            // ---- B: visit — child list, N, lambda, random number, row into registers ------------------------------------------
            if (state == ST_VISIT) {
                const size_t node0 = (size_t)b * T;
                const int seat = cur_seat;
                if (rands) r = bl_h2f(rands[node0 + cur]);
                else r = bl_uniform_half_grid(bl_philox(seed ^ (move * 0x9E3779B97F4A7C15ull), (uint64_t)b,
                                                        ((uint64_t)sim << 32) | (uint32_t)cur).x);
                int N = 0;
                nc = 0;
#pragma unroll
                for (int w = 0; w < NW; w++) cm[w] = 0;
                auto adopt = [&](const bl_node &ch, int id) {
                    const int a = ch.relation;
                    put(nc, ChildEntry{qn.fast(seat ? ch.w[1] : ch.w[0], ch.n), 0.f, a, id, (int)ch.seat | ((int)ch.terminal << 8)});
                    const u64 bit = 1ull << (a & 63);
#pragma unroll
                    for (int w = 0; w < NW; w++) cm[w] |= ((a >> 6) == w) ? bit : 0ull;    // (kept in registers: no indexed access)
                    N += ch.n;
                    nc++;
                };
                if (scan_ok) {
                    // children = the bits of the node's mask, landed with its row summary; their records are then fetched together
                    // by cp.async straight into the lane's entry slots
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                    u64 mm[MW];
#pragma unroll
                    for (int w = 0; w < MW; w++) mm[w] = w < KW ? *reinterpret_cast<const u64 *>(pg + 4 + 2 * w) : 0ull;
                    tick(7);
                    int k = 0;
#pragma unroll
                    for (int w = 0; w < MW; w++)
                        for (u64 m = mm[w]; m; m &= m - 1) {
                            const int id = w * 64 + __ffsll((long long)m) - 1;
                            if (k < KS)
                                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(pe_addr + 16u * k), "l"(t.node + node0 + id) : "memory");
                            k++;
                        }
                    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
                    tick(8);
                    k = 0;
#pragma unroll
                    for (int w = 0; w < MW; w++)
                        for (u64 m = mm[w]; m; m &= m - 1) {
                            const int id = w * 64 + __ffsll((long long)m) - 1;
                            bl_node ch;
                            if (k < KS) { union { float4 f; bl_node n; } x; x.f = pe4[k]; ch = x.n; }
                            else ch = bl_ld_node_hint(t.node + node0 + id, keep);
                            adopt(ch, id);
                            k++;
                        }
                } else {
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                    const bl_node nd = bl_ld_node_hint(t.node + node0 + cur, keep);
                    for (int c = nd.first_child; c >= 0;) {
                        const bl_node ch = bl_ld_node_hint(t.node + node0 + c, keep);
                        adopt(ch, c);
                        c = ch.next_sib;
                    }
                }
                tick(9);
                bl_aux ax;
                { union { float4 f; bl_aux a; } x; x.f = reinterpret_cast<const float4 *>(pg)[0]; ax = x.a; }      // landed with the row
                N += A - nc;                                        // every child-less action counts 1 (cuda.cu:91)
                const float lambda = bl_lambda(c_puct, N, A);
                nzpos = (uint32_t)ax.first_nz | ((uint32_t)ax.last_nz << 8);
                const u64 lam2 = pk(lambda, lambda);
#pragma unroll
                for (int c = 0; c < NCH; c++) {                   // top = lambda*pi, from the landed row
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (c < nrow4) v = ps4[c];
                    tp[2 * c] = mul2(pk(v.x, v.y), lam2); tp[2 * c + 1] = mul2(pk(v.z, v.w), lam2);
                }
                tick(10);
                // alpha seed (newton_search, cuda.cu:44-50): max_a (q[a] + max(lambda*pi[a], 1e-4)); rounding is monotone, so the
                // child-less part is max(RN(lambda*max_pi), 1e-4); tiny-value test on the smallest nonzero entry (conservative)
                float alpha0 = fmaxf(__fmul_rn(lambda, ax.max_pi), 1.e-4f);
                const bool tiny = __fmul_rn(lambda, bl_minnz(ax)) < BL_TINY;
                for (int i = 0; i < nc; i++) {
                    ChildEntry e = get(i);
                    e.top = __fmul_rn(lambda, ps[e.a]);             // the landed row still holds pi
                    alpha0 = fmaxf(alpha0, __fadd_rn(e.q, fmaxf(e.top, 1.e-4f)));
                    put(i, e);
                }
                alpha = alpha0; it = 0; error = BL_INF;
                state = tiny ? ST_SLOW : ST_PASS;
                c_evals++; c_children += nc;
                tick(12);
            }
            tick(11);
        }

        tick(3);
        // ---- C: child terms of this pass (full divisions), parked at their positions ----------------------------------------
        if (state == ST_PASS || state == ST_FINAL) {
            bool bad = false;
            // BL_CHILD_ILP children per step: their division chains are independent, and the loop is latency-bound (a short tail
            // recomputes the last child); measured on c2: 1 per step 15.6, 2 per step 14.4, 4 per step 14.95 ms per move
#ifndef BL_CHILD_ILP
#define BL_CHILD_ILP 2
#endif
            for (int i = 0; i < nc; i += BL_CHILD_ILP) {
                ChildEntry e[BL_CHILD_ILP];
                float bot[BL_CHILD_ILP], sv[BL_CHILD_ILP], gv[BL_CHILD_ILP];
#pragma unroll
                for (int u = 0; u < BL_CHILD_ILP; u++) e[u] = get(i + u < nc ? i + u : nc - 1);
#pragma unroll
                for (int u = 0; u < BL_CHILD_ILP; u++) {
                    bot[u] = __fsub_rn(alpha, e[u].q);
                    const float bb = __fmul_rn(bot[u], bot[u]);
                    sv[u] = bl_div_fast(e[u].top, bot[u]);
                    gv[u] = bl_div_fast(-e[u].top, bb);
                }
#pragma unroll
                for (int u = 0; u < BL_CHILD_ILP; u++) {
                    ps[e[u].a] = sv[u];
                    pg[e[u].a] = gv[u];
                    // bot outside [2^-60, 2^60] (never seen; alpha > q by construction) leaves the branch-free division's safe range;
                    // a negative / non-finite term would also break the monotone running sums: both go to the exact serial path
                    bad |= !(bot[u] >= 8.67e-19f && bot[u] <= 1.15e18f) || !(sv[u] >= 0.f && sv[u] <= 3.0e38f);
                }
            }
            if (bad) state = ST_SLOW;
        }
        // ---- D: exact serial fallback: the reference loops verbatim ---------------------------------------------------------
        if (state == ST_SLOW) {
#pragma unroll
            for (int c = 0; c < NCH; c++) ps4[c] = make_float4(lo(tp[2 * c]), hi(tp[2 * c]), lo(tp[2 * c + 1]), hi(tp[2 * c + 1]));
            auto topf = [&](int a) { return ps[a]; };
            auto qf = [&](int a) {
                float q = 0.f;
                for (int i = 0; i < nc; i++) { const ChildEntry e = get(i); if (e.a == a) q = e.q; }
                return q;
            };
            int iters;
            const float al = bl_newton_f(topf, qf, A, &iters);
            action = bl_sample_f(topf, qf, A, al, r);
            c_iters += iters;
            state = ST_ADVANCE;
        }
        asm volatile("" ::: "memory");
        tick(4);

        // ---- E: one Newton pass: the two sequential sums ----------------------------------------------------------------------
        const bool pass = state == ST_PASS || state == ST_FINAL;
        float accS = 0.f, accG = 0.f;
        if (__any_sync(FULL, pass)) {
            if (PROF) n_pass++;
            const float bS = alpha, bG = __fmul_rn(alpha, alpha);
            const float yS = bl_rcp_fast(bS), yG = -bl_rcp_fast(bG); // g terms: divide lambda*pi by -(alpha^2); alpha in [1e-4, ~c_puct+1]
            const u64 yS2 = pk(yS, yS), yG2 = pk(yG, yG), nbS2 = pk(-bS, -bS), bG2 = pk(bG, bG);
            float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f, g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f;
#pragma unroll
            for (int c = 0; c < NCH; c++) {
                // which of the chunk's 4 actions have a child (none for a lane that is not in a pass: its rows may be in flight)
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
                if (pass) ps4[c] = make_float4(o0, o1, o2, o3);    // a lane waiting for its visit must not touch the row
            }
        }
        tick(5);
        // ---- F: Newton update (newton_search, cuda.cu:57-66) ---------------------------------------------------------------
        if (pass) {
            if (state == ST_PASS) {
                it++;
                c_iters++;
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
    // ---- expand + env step of the warp's 32 envs, fused when every env has its own lane (no separate launch) ---------------
    if (fuse_expand) {
        const int bm = (int)blockIdx.x * 32 + lane;
        if (bm < t.B) bl_expand_one(t, sim, bm, res_leaf, res_parent, res_action, reinterpret_cast<uint32_t *>(ps), reinterpret_cast<uint8_t *>(pg));
    }
    if (PROF) {
        tick(6);
#pragma unroll
        for (int k = 0; k < (PROF ? 13 : 1); k++) {             // per phase: the slowest lane's total (lanes outside a phase hold 0 for it)
            long long m = pc[k];
            for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(FULL, m, o));
            if (lane == 0) atomicAdd(prof + k, (unsigned long long)m);
        }
        if (lane == 0) { atomicAdd(prof + 15, 1ull); atomicAdd(prof + 13, (unsigned long long)n_service); atomicAdd(prof + 14, (unsigned long long)n_pass); }
    }
    bl_count(t.counters, C_EVALS, c_evals);
    bl_count(t.counters, C_CHILDREN, c_children);
    bl_count(t.counters, C_ITERS, c_iters);
    bl_count(t.counters, C_DESCENTS, c_desc);
#undef tick
}

// ---- self test of the shared-reciprocal division -----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) divtest_kernel(uint64_t seed, int n_div, int n_num, unsigned long long *mismatch) {
    // divisors: alpha-like values and their squares, including edge significands; numerators: lambda*pi-like values.
    // Checks the scalar sequence and the packed (FMUL2/FFMA2) sequence the descent issues against __fdiv_rn.
    unsigned long long bad = 0;
    for (int i = blockIdx.x; i < n_div; i += gridDim.x) {
        bl_philox_out o = bl_philox(seed, (uint64_t)i, 1);
        unsigned mant = (i & 7) == 0 ? 0x7FFFFFu : ((i & 7) == 1 ? 0u : (o.x & 0x7FFFFFu));
        int ex = 127 - 27 + (int)(o.y % 44);                                   // 2^-27 .. 2^16
        float bdiv = __uint_as_float(((unsigned)ex << 23) | mant);
        float y = __frcp_rn(bdiv);
        const u64 y2 = pk(y, -y), nb2 = pk(-bdiv, bdiv);
        for (int j = threadIdx.x; j < n_num; j += blockDim.x) {
            bl_philox_out p = bl_philox(seed + 1, ((uint64_t)i << 32) | (unsigned)j, 2);
            unsigned nm = (j & 15) == 0 ? 0x7FFFFFu : ((j & 15) == 1 ? 0u : (p.x & 0x7FFFFFu));
            int nex = 127 - 100 + (int)(p.y % 98);                             // 2^-100 .. 2^-3
            float num = __uint_as_float(((unsigned)nex << 23) | nm);
            float q0 = __fmul_rn(num, y), r0 = __fmaf_rn(-bdiv, q0, num), q1 = __fmaf_rn(r0, y, q0);
            bad += (__float_as_uint(q1) != __float_as_uint(__fdiv_rn(num, bdiv)));
            // packed: (num / bdiv, num / -bdiv)
            const u64 n2 = pk(num, num);
            u64 q = mul2(n2, y2), rr = fma2(nb2, q, n2);
            const u64 d = fma2(rr, y2, q);
            bad += (__float_as_uint(lo(d)) != __float_as_uint(__fdiv_rn(num, bdiv)));
            bad += (__float_as_uint(hi(d)) != __float_as_uint(__fdiv_rn(-num, bdiv)));
            // branch-free reciprocal / quotient (mcts_core.cuh), also with q-normalisation-like operands (any sign, zero, up to 2^15)
            bad += (__float_as_uint(bl_rcp_fast(bdiv)) != __float_as_uint(y));
            bad += (__float_as_uint(bl_div_fast(-num, bdiv)) != __float_as_uint(__fdiv_rn(-num, bdiv)));
            const float big = __uint_as_float(((unsigned)(127 - 30 + (int)(p.w % 46)) << 23) | nm | ((p.z & 1u) << 31));   // 2^-30 .. 2^15
            bad += (__float_as_uint(bl_div_fast(big, bdiv)) != __float_as_uint(__fdiv_rn(big, bdiv)));
            bad += (__float_as_uint(bl_div_fast(0.f, bdiv)) != __float_as_uint(__fdiv_rn(0.f, bdiv)));
        }
    }
    if (bad) atomicAdd(mismatch, bad);
}

unsigned long long *g_phase_prof = nullptr;
bool g_last_fused = false;
// service gate (see the kernel): BL_GATE="num/den" in the environment overrides the default for tuning runs
int g_gate_num = BL_GATE_NUM, g_gate_den = BL_GATE_DEN;
int g_grid_limit = 0;      // BL_DESCEND_GRID: cap on the number of warps (fewer lanes than envs: lanes pull envs from the queue)
void read_gate_env() {
    static bool done = false;
    if (done) return;
    done = true;
    if (const char *e = getenv("BL_GATE")) {
        int a = 0, b = 0;
        if (sscanf(e, "%d/%d", &a, &b) == 2 && a >= 0 && b > 0) { g_gate_num = a; g_gate_den = b; }
    }
    if (const char *e = getenv("BL_DESCEND_GRID")) g_grid_limit = atoi(e);
}

template <int NCH, bool PROF>
int launch_v3p(const bl_tree *t, int sim, const bl_half *rands, uint64_t seed, int cap, cudaStream_t st) {
    const size_t smem = (size_t)3 * 32 * 4 * NCH * sizeof(float);
    static int occ = 0;
    if (occ == 0) {
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(descend_v3_kernel<NCH, PROF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
        }
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, descend_v3_kernel<NCH, PROF>, 32, smem);
        if (e != cudaSuccess) return (int)e;
        if (occ < 1) { occ = 0; return -2; }
    }
    const int need = (t->B + 31) / 32;
    int grid = need < occ * BL_NUM_SMS ? need : occ * BL_NUM_SMS;
    if (g_grid_limit > 0 && grid > g_grid_limit) grid = g_grid_limit;
    if ((int64_t)grid * 32 * cap * (int64_t)sizeof(ChildEntry) > t->scratch_bytes) return -3;
    if ((long long)grid * 32 < t->B) {                          // fewer lanes than envs: the lanes pull envs from the queue, which restarts at 0
        cudaError_t e = cudaMemsetAsync(t->counters + C_QUEUE, 0, sizeof(uint64_t), st);     // (every env resident: no queue, no memset node
        if (e != cudaSuccess) return (int)e;                                                  //  per simulation in the captured move)
    }
    // every env resident (one per lane) and the lane's rows big enough for a board + flood-fill stack: expand in the same kernel
    const bool fused = (long long)grid * 32 >= t->B && t->BP <= 4 * 4 * NCH;
    descend_v3_kernel<NCH, PROF><<<grid, 32, smem, st>>>(*t, sim, rands, seed, reinterpret_cast<ChildEntry *>(t->scratch), cap, g_phase_prof,
                                                         g_gate_num, g_gate_den, fused ? 1 : 0);
    g_last_fused = fused;
    return (int)cudaGetLastError();
}
template <int NCH>
int launch_v3(const bl_tree *t, int sim, const bl_half *rands, uint64_t seed, int cap, cudaStream_t st) {
    return g_phase_prof ? launch_v3p<NCH, true>(t, sim, rands, seed, cap, st) : launch_v3p<NCH, false>(t, sim, rands, seed, cap, st);
}

int child_cap(const bl_tree *t) { return t->A < t->T - 1 ? t->A : (t->T > 1 ? t->T - 1 : 1); }

}  // namespace

extern "C" int64_t bl_tree_scratch_bytes(const bl_tree *t) {
    const int64_t lanes_needed = ((int64_t)t->B + 31) / 32 * 32, lanes_max = (int64_t)BL_NUM_SMS * 32 * 32;
    const int64_t v3 = (lanes_needed < lanes_max ? lanes_needed : lanes_max) * child_cap(t) * (int64_t)sizeof(ChildEntry);
    const int64_t mw = bl_mw_scratch_bytes(t), fx = t->cpi ? bl_fx_scratch_bytes(t) : 0, al = t->cpi ? bl_all_scratch_bytes(t) : 0;
    int64_t m = v3 > mw ? v3 : mw;
    m = m > fx ? m : fx;
    return m > al ? m : al;
}

int bl_expand_step(const bl_tree *t, int sim, cudaStream_t st) {
    const size_t xsmem = (size_t)2 * XNT * ((t->BP >> 2) | 1) * 4;
    if (xsmem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(expand_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xsmem);
        if (e != cudaSuccess) return (int)e;
    }
    expand_step_kernel<<<(t->B + XNT - 1) / XNT, XNT, xsmem, st>>>(*t, sim);
    return (int)cudaGetLastError();
}

int bl_descend_v3(const bl_tree *t, int sim, const bl_half *rands, uint64_t seed, cudaStream_t st) {
    read_gate_env();
    const int cap = child_cap(t);
    const int nch = (t->A + 3) / 4;
    int rc;
    if (nch <= 3) rc = launch_v3<3>(t, sim, rands, seed, cap, st);
    else if (nch <= 7) rc = launch_v3<7>(t, sim, rands, seed, cap, st);
    else if (nch <= 13) rc = launch_v3<13>(t, sim, rands, seed, cap, st);
    else if (nch <= 21) rc = launch_v3<21>(t, sim, rands, seed, cap, st);
    else if (nch <= 31) rc = launch_v3<31>(t, sim, rands, seed, cap, st);
    else if (nch <= 43) rc = launch_v3<43>(t, sim, rands, seed, cap, st);
    else rc = -2;
    if (rc) return rc;
    return g_last_fused ? 0 : bl_expand_step(t, sim, st);
}

unsigned long long *bl_phase_prof() { return g_phase_prof; }

extern "C" int bl_debug_set_phase_profile(uint64_t *buf) {
    g_phase_prof = reinterpret_cast<unsigned long long *>(buf);
    return 0;
}

extern "C" int bl_debug_set_descend_grid(int warps) {
    read_gate_env();
    g_grid_limit = warps > 0 ? warps : 0;
    return 0;
}

extern "C" int bl_selftest_division(uint64_t seed, int n_div, int n_num, uint64_t *mismatch, bl_stream stream) {
    divtest_kernel<<<BL_NUM_SMS * 8, 256, 0, bl_cu(stream)>>>(seed, n_div, n_num, reinterpret_cast<unsigned long long *>(mismatch));
    BL_LAUNCH_CHECK();
}
