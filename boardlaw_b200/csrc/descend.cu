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
#include "engine_internal.cuh"
#include "hex_core.cuh"
#include "mcts_core.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;
#define BL_TINY 7.888609052210118e-31f   /* 2^-100 */

enum { ST_IDLE = 0, ST_VISIT = 1, ST_PASS = 2, ST_FINAL = 3, ST_SAMPLE = 4, ST_SLOW = 5, ST_ADVANCE = 6 };

struct __align__(16) ChildEntry { float q, top; int a, id; };

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { return ((u64)__float_as_uint(b) << 32) | __float_as_uint(a); }
__device__ __forceinline__ float lo(u64 v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float hi(u64 v) { return __uint_as_float((unsigned)(v >> 32)); }
// packed fp32 pipe operations: IEEE round-to-nearest per half, denormals kept (tools/ubench_fp32x2.cu checks them against the
// scalar intrinsics on 1.8e10 operand triples)
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int NCH>
__global__ void __launch_bounds__(32) descend_v3_kernel(bl_tree t, int sim, const bl_half *__restrict__ rands, uint64_t seed,
                                                        ChildEntry *__restrict__ clists, int cap) {
    constexpr int PS = 4 * NCH;                         // row pitch in floats; NCH odd => conflict-free 128-bit lane-private rows
    extern __shared__ float4 smem4[];
    const int A = t.A, T = t.T;
    const int lane = threadIdx.x;
    float *ps = reinterpret_cast<float *>(smem4) + lane * PS;       // staging row / S child terms in, running S sums out
    float *pg = ps + 32 * PS;                                       // g child terms
    float4 *ps4 = reinterpret_cast<float4 *>(ps);
    const uint32_t ps_addr = smem_u32(ps), pg_addr = smem_u32(pg);
    ChildEntry *cl = clists + ((size_t)blockIdx.x * 32 + lane) * cap;
    const bl_qnorm qn(t.qrange + 2 * sim);
    const uint64_t move = t.counters[C_MOVE];
    int *queue = reinterpret_cast<int *>(t.counters + C_QUEUE);
    const int nrow4 = t.AP >> 2;

    u64 tp[2 * NCH];                                  // lambda*pi of the current node, element pairs; sign set where a child exists
    ChildEntry e0 = {0.f, 0.f, 0, -1}, e1 = {0.f, 0.f, 0, -1};      // the first two children live in registers, the rest in `cl`
    int b = -1, cur = -1, parent = 0, action = -1, state = ST_VISIT, nc = 0, it = 0;
    float alpha = 1.f, error = 0.f, r = 0.f;
    u64 cmask = 0;                                    // bit c set: chunk c holds a child
    uint32_t nzpos = 0;                               // first_nz | last_nz << 8 of the current row
    bool exhausted = false;                           // warp-uniform
    unsigned c_evals = 0, c_children = 0, c_iters = 0, c_desc = 0;

    while (true) {
        // ---- A: finished descents write their result and take the next env -------------------------------------------
        bl_node nd;
        bool evaluable = false, fresh = false;
        if (state == ST_VISIT) {
            if (b >= 0 && cur >= 0) { nd = bl_ld_node(t.node + (size_t)b * T + cur); evaluable = !nd.terminal; }
            if (!evaluable) {
                if (b >= 0) {
                    t.leaf[b] = (int16_t)cur;                      // existing terminal child, or -1: expand_step decides
                    t.leaf_parent[b] = (int16_t)parent;
                    t.leaf_action[b] = (int16_t)action;
                    c_desc++;
                }
                fresh = true;
            }
        }
        const unsigned req = __ballot_sync(FULL, fresh);
        if (req) {
            int base = t.B;
            if (!exhausted) {
                if (lane == 0) base = atomicAdd(queue, __popc(req));
                base = __shfl_sync(FULL, base, 0);
                exhausted = base + __popc(req) >= t.B;
            }
            if (fresh) {
                const int nb = base + __popc(req & ((1u << lane) - 1u));
                b = nb < t.B ? nb : -1;
                cur = 0; parent = 0; action = -1;
                state = b < 0 ? ST_IDLE : ST_VISIT;
                if (b >= 0) { nd = bl_ld_node(t.node + (size_t)b * T); evaluable = !nd.terminal; }   // a terminal root ends the descent next trip
            }
        }
        if (__all_sync(FULL, state == ST_IDLE)) break;

        // ---- B: visit — child list, N, lambda, random number, row -> registers --------------------------------------------
        if (state == ST_VISIT && evaluable) {
            const size_t node0 = (size_t)b * T;
            const int seat = nd.seat;
            int N = 0;
            nc = 0;
            cmask = 0;
            for (int c = nd.first_child; c >= 0;) {
                const bl_node ch = bl_ld_node(t.node + node0 + c);
                const ChildEntry e = {qn(ch.w[seat], ch.n), 0.f, (int)ch.relation, c};
                if (nc == 0) e0 = e; else if (nc == 1) e1 = e; else cl[nc] = e;
                cmask |= 1ull << (ch.relation >> 2);
                N += ch.n;
                nc++;
                c = ch.next_sib;
            }
            N += A - nc;                                        // every child-less action counts 1 (cuda.cu:91)
            const float lambda = bl_lambda(bl_h2f(t.c_puct[b]), N, A);
            if (rands) r = bl_h2f(rands[node0 + cur]);
            else r = bl_uniform_half_grid(bl_philox(seed ^ (move * 0x9E3779B97F4A7C15ull), (uint64_t)b,
                                                    ((uint64_t)sim << 32) | (uint32_t)cur).x);
            const bl_aux ax = bl_ld_aux(t.aux + node0 + cur);
            nzpos = (uint32_t)ax.first_nz | ((uint32_t)ax.last_nz << 8);
            // row: top = lambda*pi, staged in the lane's shared-memory row
            const float4 *row = reinterpret_cast<const float4 *>(t.pi + (node0 + cur) * t.AP);
#pragma unroll
            for (int c = 0; c < NCH; c++) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c < nrow4) v = row[c];
                ps4[c] = make_float4(__fmul_rn(lambda, v.x), __fmul_rn(lambda, v.y), __fmul_rn(lambda, v.z), __fmul_rn(lambda, v.w));
            }
            // alpha seed (newton_search, cuda.cu:44-50): max_a (q[a] + max(lambda*pi[a], 1e-4)); rounding is monotone, so the
            // child-less part is max(RN(lambda*max_pi), 1e-4); tiny-value test on the smallest nonzero entry (conservative)
            float alpha0 = fmaxf(__fmul_rn(lambda, ax.max_pi), 1.e-4f);
            const float tmin = __fmul_rn(lambda, bl_minnz(ax));
            const bool tiny = tmin < BL_TINY;
            auto adopt = [&](ChildEntry &e) {
                const float tv = ps[e.a];
                e.top = tv;
                alpha0 = fmaxf(alpha0, __fadd_rn(e.q, fmaxf(tv, 1.e-4f)));
                ps[e.a] = -tv;                                  // sign = "has a child"; the numerator stays in the entry
            };
            if (nc > 0) adopt(e0);
            if (nc > 1) adopt(e1);
            for (int i = 2; i < nc; i++) { ChildEntry e = cl[i]; adopt(e); cl[i] = e; }
#pragma unroll
            for (int c = 0; c < NCH; c++) {
                const float4 v = ps4[c];
                tp[2 * c] = pk(v.x, v.y); tp[2 * c + 1] = pk(v.z, v.w);
            }
            alpha = alpha0; it = 0; error = BL_INF;
            state = tiny ? ST_SLOW : ST_PASS;
            c_evals++; c_children += nc;
        }

        // ---- C: child terms of this pass (full divisions), parked at their positions ----------------------------------------
        const bool pass = state == ST_PASS || state == ST_FINAL;
        if (pass) {
            bool bad = false;
            auto park = [&](const ChildEntry &e) {
                const float bot = __fsub_rn(alpha, e.q);
                const float s = __fdiv_rn(e.top, bot);
                ps[e.a] = s;
                pg[e.a] = __fdiv_rn(-e.top, __fmul_rn(bot, bot));
                bad |= !(s >= 0.f && s <= 3.0e38f);              // negative / non-finite term: the prefix would not be monotone
            };
            if (nc > 0) park(e0);
            if (nc > 1) park(e1);
            for (int i = 2; i < nc; i++) park(cl[i]);
            if (bad) state = ST_SLOW;
        }
        // ---- D: exact serial fallback: the reference loops verbatim ---------------------------------------------------------
        if (state == ST_SLOW) {
#pragma unroll
            for (int c = 0; c < NCH; c++) ps4[c] = make_float4(lo(tp[2 * c]), hi(tp[2 * c]), lo(tp[2 * c + 1]), hi(tp[2 * c + 1]));
            auto topf = [&](int a) { return fabsf(ps[a]); };
            auto qf = [&](int a) {
                float q = 0.f;
                if (nc > 0 && e0.a == a) q = e0.q;
                if (nc > 1 && e1.a == a) q = e1.q;
                for (int i = 2; i < nc; i++) if (cl[i].a == a) q = cl[i].q;
                return q;
            };
            int iters;
            const float al = bl_newton_f(topf, qf, A, &iters);
            action = bl_sample_f(topf, qf, A, al, r);
            c_iters += iters;
            state = ST_ADVANCE;
        }
        asm volatile("" ::: "memory");

        // ---- E: one Newton pass: the two sequential sums ----------------------------------------------------------------------
        const bool pass2 = state == ST_PASS || state == ST_FINAL;
        float accS = 0.f, accG = 0.f;
        if (__any_sync(FULL, pass2)) {
            const float bS = alpha, bG = __fmul_rn(alpha, alpha);
            const float yS = __frcp_rn(bS), yG = -__frcp_rn(bG);     // g terms: divide lambda*pi by -(alpha^2)
            const u64 yS2 = pk(yS, yS), yG2 = pk(yG, yG), nbS2 = pk(-bS, -bS), bG2 = pk(bG, bG);
            float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f, g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f;
#pragma unroll
            for (int c = 0; c < NCH; c++) {
                const uint32_t has = (uint32_t)(cmask >> c) & 1u;
                asm volatile(
                    "{\n.reg .pred p;\nsetp.ne.u32 p, %8, 0;\n"
                    "@p ld.shared.v4.f32 {%0,%1,%2,%3}, [%9];\n"
                    "@p ld.shared.v4.f32 {%4,%5,%6,%7}, [%10];\n}"
                    : "+f"(p0), "+f"(p1), "+f"(p2), "+f"(p3), "+f"(g0), "+f"(g1), "+f"(g2), "+f"(g3)
                    : "r"(has), "r"(ps_addr + 16u * c), "r"(pg_addr + 16u * c));
                const u64 t01 = tp[2 * c], t23 = tp[2 * c + 1];
                u64 q = mul2(t01, yS2), rr = fma2(nbS2, q, t01);
                const u64 s01 = fma2(rr, yS2, q);
                q = mul2(t01, yG2); rr = fma2(bG2, q, t01);
                const u64 h01 = fma2(rr, yG2, q);
                q = mul2(t23, yS2); rr = fma2(nbS2, q, t23);
                const u64 s23 = fma2(rr, yS2, q);
                q = mul2(t23, yG2); rr = fma2(bG2, q, t23);
                const u64 h23 = fma2(rr, yG2, q);
                const bool c0 = (int)(unsigned)t01 < 0, c1 = (int)(unsigned)(t01 >> 32) < 0;
                const bool c2 = (int)(unsigned)t23 < 0, c3 = (int)(unsigned)(t23 >> 32) < 0;
                accS = __fadd_rn(accS, c0 ? p0 : lo(s01)); accG = __fadd_rn(accG, c0 ? g0 : lo(h01)); const float o0 = accS;
                accS = __fadd_rn(accS, c1 ? p1 : hi(s01)); accG = __fadd_rn(accG, c1 ? g1 : hi(h01)); const float o1 = accS;
                accS = __fadd_rn(accS, c2 ? p2 : lo(s23)); accG = __fadd_rn(accG, c2 ? g2 : lo(h23)); const float o2 = accS;
                accS = __fadd_rn(accS, c3 ? p3 : hi(s23)); accG = __fadd_rn(accG, c3 ? g3 : hi(h23)); const float o3 = accS;
                ps4[c] = make_float4(o0, o1, o2, o3);              // harmless for lanes that are not in a pass: their row is dead
            }
        }
        // ---- F: Newton update (newton_search, cuda.cu:57-66) ---------------------------------------------------------------
        if (pass2) {
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
        // ---- G: inverse-CDF search over the running sums (descend_kernel, cuda.cu:160-176) ----------------------------------
        if (state == ST_SAMPLE) {
            // every term is >= 0 (checked for child terms above), so the sums are non-decreasing: first index with sum >= r.
            // The reference additionally skips p == 0 entries: the first hit can only have p == 0 when r == 0 (then the
            // answer is the first nonzero entry), and when no sum reaches r the answer is the last nonzero entry.
            int l = 0, h = A;
            while (l < h) {
                const int mid = (l + h) >> 1;
                if (ps[mid] >= r) h = mid; else l = mid + 1;
            }
            const int first_nz = nzpos & 255, last_nz = (nzpos >> 8) & 255;
            action = first_nz == 255 ? -1 : (l < A ? (r <= 0.f ? first_nz : l) : last_nz);
            state = ST_ADVANCE;
        }
        // ---- H: step to the chosen child -------------------------------------------------------------------------------------
        if (state == ST_ADVANCE) {
            parent = cur;
            int next = -1;
            if (nc > 0 && e0.a == action) next = e0.id;
            if (nc > 1 && e1.a == action) next = e1.id;
            for (int i = 2; i < nc; i++)
                if (cl[i].a == action) next = cl[i].id;
            cur = action >= 0 ? next : -1;
            state = ST_VISIT;
        }
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
    const int A = t.A, T = t.T;
    const int b = blockIdx.x * XNT + tid, bw = blockIdx.x * XNT + wbase;
    const bool in_range = b < t.B;
    const size_t node0 = (size_t)(in_range ? b : 0) * T;
    int leaf = -1, parent = 0, action = -1;
    bool ok = false, fresh = false;
    bl_node pn, ln;
    if (in_range) {
        leaf = t.leaf[b]; parent = t.leaf_parent[b]; action = t.leaf_action[b];
        ok = action >= 0;
        if (ok) {
            pn = bl_ld_node(t.node + node0 + parent);
            if (leaf < 0) {                                     // new node in slot `sim`
                leaf = sim;
                fresh = true;
                ln.parent = (int16_t)parent; ln.relation = (int16_t)action; ln.first_child = -1; ln.next_sib = pn.first_child;
                ln.n = 0; ln.w[0] = 0; ln.w[1] = 0;
                t.node[node0 + parent].first_child = (int16_t)sim;
            } else {                                            // stopped at an existing terminal child: reuse its slot
                ln = bl_ld_node(t.node + node0 + leaf);
            }
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
        const int seat = pn.seat;
        const int win = bl_hex_place<uint8_t>(bd + tid, stk + tid, XPITCH, t.S, seat, action);
        const float r0 = win == 1 ? 1.f : (win == 2 ? -1.f : 0.f), r1 = win == 1 ? -1.f : (win == 2 ? 1.f : 0.f);
        reinterpret_cast<uint32_t *>(t.aux + node0 + leaf)[0] = (uint32_t)bl_f2h(r0) | ((uint32_t)bl_f2h(r1) << 16);
        ln.terminal = win != 0;
        ln.seat = win ? 0 : (uint8_t)(1 - seat);
        if (fresh) bl_st_node(t.node + node0 + leaf, ln);
        else bl_st_node_stats(t.node + node0 + leaf, ln);
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
    // divisors: alpha-like values and their squares, including edge significands; numerators: lambda*pi-like values.
    // Checks the scalar sequence and the packed (FMUL2/FFMA2) sequence the descent issues against __fdiv_rn.
    unsigned long long bad = 0;
    for (int i = blockIdx.x; i < n_div; i += gridDim.x) {
        bl_philox_out o = bl_philox(seed, (uint64_t)i, 1);
        unsigned mant = (i & 7) == 0 ? 0x7FFFFFu : ((i & 7) == 1 ? 0u : (o.x & 0x7FFFFFu));
        int ex = 127 - 27 + (int)(o.y % 30);                                   // 2^-27 .. 2^2
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
        }
    }
    if (bad) atomicAdd(mismatch, bad);
}

template <int NCH>
int launch_v3(const bl_tree *t, int sim, const bl_half *rands, uint64_t seed, int cap, cudaStream_t st) {
    const size_t smem = (size_t)2 * 32 * 4 * NCH * sizeof(float);
    static int occ = 0;
    if (occ == 0) {
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(descend_v3_kernel<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
        }
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, descend_v3_kernel<NCH>, 32, smem);
        if (e != cudaSuccess) return (int)e;
        if (occ < 1) { occ = 0; return -2; }
    }
    const int need = (t->B + 31) / 32;
    const int grid = need < occ * BL_NUM_SMS ? need : occ * BL_NUM_SMS;
    if ((int64_t)grid * 32 * cap * (int64_t)sizeof(ChildEntry) > t->scratch_bytes) return -3;
    descend_v3_kernel<NCH><<<grid, 32, smem, st>>>(*t, sim, rands, seed, reinterpret_cast<ChildEntry *>(t->scratch), cap);
    return (int)cudaGetLastError();
}

int child_cap(const bl_tree *t) { return t->A < t->T - 1 ? t->A : (t->T > 1 ? t->T - 1 : 1); }

}  // namespace

extern "C" int64_t bl_tree_scratch_bytes(const bl_tree *t) {
    const int64_t lanes_needed = ((int64_t)t->B + 31) / 32 * 32, lanes_max = (int64_t)BL_NUM_SMS * 32 * 32;
    return (lanes_needed < lanes_max ? lanes_needed : lanes_max) * child_cap(t) * (int64_t)sizeof(ChildEntry);
}

int bl_expand_step(const bl_tree *t, int sim, cudaStream_t st) {
    const size_t xsmem = (size_t)2 * t->A * XPITCH;
    if (xsmem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(expand_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xsmem);
        if (e != cudaSuccess) return (int)e;
    }
    expand_step_kernel<<<(t->B + XNT - 1) / XNT, XNT, xsmem, st>>>(*t, sim);
    return (int)cudaGetLastError();
}

int bl_descend_v3(const bl_tree *t, int sim, const bl_half *rands, uint64_t seed, cudaStream_t st) {
    const int cap = child_cap(t);
    const int nch = (t->A + 3) / 4;
    cudaError_t e = cudaMemsetAsync(t->counters + C_QUEUE, 0, sizeof(uint64_t), st);
    if (e != cudaSuccess) return (int)e;
    int rc;
    if (nch <= 3) rc = launch_v3<3>(t, sim, rands, seed, cap, st);
    else if (nch <= 7) rc = launch_v3<7>(t, sim, rands, seed, cap, st);
    else if (nch <= 13) rc = launch_v3<13>(t, sim, rands, seed, cap, st);
    else if (nch <= 21) rc = launch_v3<21>(t, sim, rands, seed, cap, st);
    else if (nch <= 31) rc = launch_v3<31>(t, sim, rands, seed, cap, st);
    else if (nch <= 43) rc = launch_v3<43>(t, sim, rands, seed, cap, st);
    else rc = -2;
    if (rc) return rc;
    return bl_expand_step(t, sim, st);
}

extern "C" int bl_selftest_division(uint64_t seed, int n_div, int n_num, uint64_t *mismatch, bl_stream stream) {
    divtest_kernel<<<BL_NUM_SMS * 8, 256, 0, bl_cu(stream)>>>(seed, n_div, n_num, reinterpret_cast<unsigned long long *>(mismatch));
    BL_LAUNCH_CHECK();
}
