// Fused search engine for sm_100a: MCTS.initialize / simulate / root (boardlaw/mcts/__init__.py:72-149) on a
// persistent, privately laid out workspace (bl_tree, include/boardlaw_b200.h).
//
// Layout choices:
//   * exp(logits) is stored once per node as fp32 (`pi`, value of the host-libm table at half(logit)), so a
//     descent never evaluates exp; rows are 16-byte aligned (pitch AP).
//   * the reference's dense children (B,T,A) tensor — 38% of its tree bytes, ~98% of it -1 — is replaced by
//     first-child / next-sibling lists over the (B,T) node arrays; it is materialised only on request.
//   * transition_q's global (min,max) is produced by the preceding backup (one slot per simulation), and the
//     normalised q of a child is computed on the fly from (w, n); no q tensor exists.
// Kernel shape: one lane per env for the sequential fp32 arithmetic (bit-exact order, see mcts_core.cuh), with
// warp-cooperative coalesced row loads into lane-major shared-memory columns (odd pitch => conflict-free both ways).
//
// Compiled with -fmad=false -prec-div=true -ftz=false (see build.py).
#include <cstdlib>
#include "engine_internal.cuh"
#include "hex_core.cuh"
#include "mcts_core.cuh"

namespace {

constexpr int ENT = 64;            // lanes (= envs) per CTA
constexpr int FP = ENT + 1;        // float column pitch (odd: conflict-free for row-wise fills and lane-wise reads)
constexpr int BPITCH = ENT + 4;    // byte column pitch (17 words: consecutive cells land on distinct banks)

// counters slots
int g_descend_variant = 0;      // 0 = by board size (measured): A <= 81 one lane per env (descend.cu), larger boards four lanes (descend_mw.cu)

__global__ void __launch_bounds__(256) reset_kernel(bl_tree t, const uint8_t *__restrict__ board,
                                                    const int32_t *__restrict__ seats, bl_half c_puct) {
    const long long BT = (long long)t.B * t.T;
    const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
    for (long long i = i0; i < BT; i += stride) {
        const long long b = i / t.T;
        bl_node nd;
        nd.parent = -1; nd.relation = -1; nd.first_child = -1; nd.next_sib = -1;
        nd.n = 0; nd.w[0] = 0; nd.w[1] = 0; nd.seat = (uint8_t)seats[b]; nd.terminal = 0;
        bl_st_node(t.node + i, nd);
        reinterpret_cast<uint4 *>(t.aux)[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    for (long long i = i0; i < (long long)t.B * t.A; i += stride) {
        long long b = i / t.A; int c = (int)(i - b * t.A);
        t.board[(b * t.T) * t.BP + c] = board[i];
    }
    const int TP = (t.T + 7) & ~7;
    for (long long i = i0; i < (long long)t.B * TP; i += stride) t.parent_of[i] = -1;
    for (long long i = i0; i < BT * ((t.T + 63) >> 6); i += stride) t.kids[i] = 0ull;
    for (long long i = i0; i < t.B; i += stride) t.c_puct[i] = c_puct;
    if (i0 == 0) t.counters[C_MOVE] += 1;                       // keys this search's in-kernel random streams; no host write per move
    int *qr = reinterpret_cast<int *>(t.qrange);
    for (long long i = i0; i <= t.T; i += stride) {
        // slot 1 serves the first descent: the all-zero tree has (min,max) = (0,0)
        qr[2 * i] = (i == 1) ? bl_f2ord(0.f) : bl_f2ord(BL_INF);
        qr[2 * i + 1] = (i == 1) ? bl_f2ord(0.f) : bl_f2ord(-BL_INF);
    }
}

// Gamma(shape) draw, Marsaglia & Tsang (2000) with the shape < 1 boost, on the counter-based Philox stream (key, ctr): the
// in-kernel counterpart of torch._sample_dirichlet's gamma draws (same distribution, different stream)
__device__ __forceinline__ float bl_gamma(float shape, uint64_t key, uint64_t ctr) {
    const float a1 = shape < 1.f ? shape + 1.f : shape;
    const float d = a1 - 1.f / 3.f, c = 1.f / sqrtf(9.f * d);
    float g = d;
    for (uint32_t attempt = 0; attempt < 64; attempt++) {
        const bl_philox_out o = bl_philox(key, ctr, attempt);
        const float u1 = ((float)(o.x >> 8) + .5f) * (1.f / 16777216.f), u2 = ((float)(o.y >> 8) + .5f) * (1.f / 16777216.f);
        const float u3 = ((float)(o.z >> 8) + .5f) * (1.f / 16777216.f), u4 = ((float)(o.w >> 8) + .5f) * (1.f / 16777216.f);
        const float x = sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);                 // standard normal (Box-Muller)
        const float v = (1.f + c * x) * (1.f + c * x) * (1.f + c * x);
        if (v > 0.f && logf(u3) < .5f * x * x + d - d * v + d * logf(v)) {
            g = d * v;
            if (shape < 1.f) g *= powf(u4, 1.f / shape);
            break;
        }
    }
    return g;
}

struct PriorMix {                 // SRC == 2: what the root prior is mixed from (dirichlet_noise, boardlaw/mcts/__init__.py:13-24)
    const uint8_t *board;         // (B,A) absolute-frame root boards
    const int32_t *seats;         // (B,)
    const float *draw;            // (B,A) injected raw Dirichlet sample (before masking), or NULL: drawn here
    float eps, conc;              // noise_eps, alpha_scale / A
    uint64_t seed;
};

// logits/v of one node per env -> pi row + row summary (+ prior when node 0) + v, rounding through half like
// decisions.half() (boardlaw/mcts/__init__.py:135-136).  One warp per env.  SRC: 0 = fp32 rows, 1 = half rows, 2 = fp32 root rows
// mixed with Dirichlet noise on the way in (MCTS.initialize, boardlaw/mcts/__init__.py:72-80).
template <int SRC>
__global__ void __launch_bounds__(256) set_eval_kernel(bl_tree t, int node, const void *__restrict__ logits_,
                                                       const void *__restrict__ v_, PriorMix mix) {
    const int lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    constexpr int MAXK = 8;                                     // actions per lane (A <= 255)
    for (int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < t.B; b += nwarps) {
        const int nd = node >= 0 ? node : t.leaf[b];
        if (nd < 0) continue;
        const size_t slot = (size_t)b * t.T + nd;
        float noise[MAXK];
        if (SRC == 2) {
            // the Dirichlet sample restricted to the legal moves and renormalised (draw[~valid] = 0; draw /= draw.sum())
            const int seat = mix.seats[b], S = t.S;
            const uint64_t move = t.counters[C_MOVE];
            float tot = 0.f;
#pragma unroll
            for (int k = 0; k < MAXK; k++) {
                const int a = lane + 32 * k;
                float g = 0.f;
                if (a < t.A) {
                    const int cell = seat ? (a % S) * S + a / S : a;                  // white sees the transposed board
                    if (mix.board[(size_t)b * t.A + cell] == 0)
                        g = mix.draw ? mix.draw[(size_t)b * t.A + a]
                                     : fmaxf(bl_gamma(mix.conc, mix.seed ^ (move * 0xD1B54A32D192ED03ull), ((uint64_t)b << 8) | (uint64_t)a), 1.17549435e-38f);
                }
                noise[k] = g;
                tot += g;
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
#pragma unroll
            for (int k = 0; k < MAXK; k++) noise[k] = noise[k] / tot;
        }
        float mx = 0.f, mn = BL_INF, pa = 0.f;
        int fz = 255, lz = -1;
        double carry = 0.;                                     // running prefix sum of the row (cpi), chunk of 32 actions at a time
#pragma unroll
        for (int k = 0; k < MAXK; k++) {
            const int a0 = 32 * k;
            if (a0 >= t.AP) break;
            const int a = a0 + lane;
            float p = 0.f;
            if (a < t.A) {
                const size_t i = (size_t)b * t.A + a;
                bl_half h;
                if (SRC == 1) h = reinterpret_cast<const bl_half *>(logits_)[i];
                else if (SRC == 0) h = bl_f2h(reinterpret_cast<const float *>(logits_)[i]);
                else h = bl_f2h(logf(expf(reinterpret_cast<const float *>(logits_)[i]) * (1.f - mix.eps) + noise[k] * mix.eps));
                p = t.exp_lut[h];
                t.pi[slot * t.AP + a] = p;
                if (nd == 0) t.prior[i] = h;
                if (t.logits) t.logits[slot * t.A + a] = h;
                if (p != 0.f) { mx = fmaxf(mx, p); mn = fminf(mn, p); fz = min(fz, a); lz = max(lz, a); }
                pa = __fmaf_rn((float)a, p, pa);
            }
            if (t.cpi) {
                double run = (double)p;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const double up = __shfl_up_sync(0xffffffffu, run, o);
                    if (lane >= o) run += up;
                }
                run += carry;
                if (a < t.AP) t.cpi[slot * t.AP + a] = (float)run;
                carry = __shfl_sync(0xffffffffu, run, 31);
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) pa += __shfl_xor_sync(0xffffffffu, pa, o);
        if (t.cpi && lane == 0) t.psum[slot] = pa;
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            fz = min(fz, __shfl_xor_sync(0xffffffffu, fz, o));
            lz = max(lz, __shfl_xor_sync(0xffffffffu, lz, o));
        }
        if (lane == 0) {
            bl_half hv[2];
            for (int s = 0; s < 2; s++)
                hv[s] = SRC == 1 ? reinterpret_cast<const bl_half *>(v_)[(size_t)b * 2 + s] : bl_f2h(reinterpret_cast<const float *>(v_)[(size_t)b * 2 + s]);
            uint32_t *ax = reinterpret_cast<uint32_t *>(t.aux + slot);
            ax[1] = (uint32_t)hv[0] | ((uint32_t)hv[1] << 16);
            if (node < 0) reinterpret_cast<uint32_t *>(t.leaf_v)[b] = ax[1];
            // minnz_hi truncates the smallest nonzero pi downwards: the tiny-value test it feeds errs on the safe side
            reinterpret_cast<uint2 *>(ax)[1] = make_uint2(__float_as_uint(mx), (__float_as_uint(mn) >> 16) | ((uint32_t)(fz & 255) << 16) |
                                                                                  ((uint32_t)((lz < 0 ? 0 : lz) & 255) << 24));
        }
    }
}

// ---- descend + expand + env step, lock-step variant (on-device cross-check of descend.cu) ---------------------------
struct Smem {
    float *top, *q;        // [A][FP]
    uint8_t *bd, *stk;     // [A][BPITCH]
};
__device__ __forceinline__ Smem carve(uint8_t *raw, int A) {
    Smem s;
    s.top = reinterpret_cast<float *>(raw);
    s.q = s.top + (size_t)A * FP;
    s.bd = reinterpret_cast<uint8_t *>(s.q + (size_t)A * FP);
    s.stk = s.bd + (size_t)A * BPITCH;
    return s;
}
size_t descend_smem(int A) { return (size_t)2 * A * FP * sizeof(float) + (size_t)2 * A * BPITCH; }

__global__ void __launch_bounds__(ENT) descend_expand_kernel(bl_tree t, int sim, const bl_half *__restrict__ rands,
                                                             uint64_t seed) {
    extern __shared__ __align__(16) uint8_t raw[];
    Smem sm = carve(raw, t.A);
    const int tid = threadIdx.x, lane = tid & 31, wbase = tid & ~31;
    const int A = t.A, T = t.T;
    const int b = blockIdx.x * ENT + tid;
    const int bw = blockIdx.x * ENT + wbase;               // first env of this warp
    const bool in_range = b < t.B;
    const size_t node0 = (size_t)(in_range ? b : 0) * T;
    float *top = sm.top + tid, *q = sm.q + tid;

    for (int a = 0; a < A; a++) q[a * FP] = 0.f;

    const bl_qnorm qn(t.qrange + 2 * sim);
    const float c_puct = in_range ? bl_h2f(t.c_puct[b]) : 0.f;
    const uint64_t move = t.counters[C_MOVE];

    int cur = in_range ? 0 : -1, parent = 0, action = -1;
    unsigned c_evals = 0, c_children = 0, c_iters = 0;

    while (true) {
        bl_node nd;
        bool active = false;
        if (cur >= 0) { nd = bl_ld_node(t.node + node0 + cur); active = !nd.terminal; }
        unsigned mask = __ballot_sync(0xffffffffu, active);
        if (!mask) break;
        // cooperative, coalesced load of each active lane's pi row into its shared-memory column
        for (unsigned m = mask; m; m &= m - 1) {
            int l = __ffs(m) - 1;
            int tl = __shfl_sync(0xffffffffu, cur, l);
            const float *row = t.pi + ((size_t)(bw + l) * T + tl) * t.AP;
            for (int a = lane; a < A; a += 32) sm.top[a * FP + wbase + l] = row[a];
        }
        __syncwarp();
        if (active) {
            const int seat = nd.seat;
            int N = 0, nc = 0;
            for (int c = nd.first_child; c >= 0;) {
                const bl_node ch = bl_ld_node(t.node + node0 + c);
                q[ch.relation * FP] = qn(seat ? ch.w[1] : ch.w[0], ch.n);
                N += ch.n;
                nc++;
                c = ch.next_sib;
            }
            N += A - nc;                                        // every child-less action counts 1 (cuda.cu:91)
            const float lambda = bl_lambda(c_puct, N, A);
            for (int a = 0; a < A; a++) top[a * FP] = __fmul_rn(lambda, top[a * FP]);
            int it;
            const float alpha = bl_newton(top, q, FP, A, &it);
            float r;
            if (rands) r = bl_h2f(rands[node0 + cur]);
            else r = bl_uniform_half_grid(bl_philox(seed ^ (move * 0x9E3779B97F4A7C15ull), (uint64_t)b, ((uint64_t)sim << 32) | (uint32_t)cur).x);
            action = bl_sample(top, q, FP, A, alpha, r);
            parent = cur;
            int next = -1;
            for (int c = nd.first_child; c >= 0;) {
                const bl_node ch = bl_ld_node(t.node + node0 + c);
                q[ch.relation * FP] = 0.f;                      // leave the column zeroed for the next node
                if (ch.relation == action) next = c;
                c = ch.next_sib;
            }
            cur = action >= 0 ? next : -2;                      // -2: no positive-probability action (error)
            c_evals++; c_children += nc; c_iters += it;
        }
        __syncwarp();
    }
    if (in_range) {
        t.leaf[b] = (int16_t)(cur == -2 ? -1 : cur);            // existing terminal child, or -1
        t.leaf_parent[b] = (int16_t)parent;
        t.leaf_action[b] = (int16_t)(cur == -2 ? -1 : action);
    }
    bl_count(t.counters, C_EVALS, c_evals);
    bl_count(t.counters, C_CHILDREN, c_children);
    bl_count(t.counters, C_ITERS, c_iters);
    bl_count(t.counters, C_DESCENTS, (in_range && action >= 0 && cur != -2) ? 1u : 0u);
}

// ---- backup + q-range scan ----------------------------------------------------------------------------------------
// Warp-local: a warp owns BK_ENVS consecutive envs.  It stages their node records 0..sim (16 B each; an env's records are
// contiguous) in shared memory with coalesced loads, then lanes 0..BK_ENVS-1 walk their env's leaf->root path in the staged copies
// (shared-memory latency per step instead of a DRAM round trip), then all lanes write the touched records back and scan the
// staged records for the (min,max) of w/(n+1e-4) that the NEXT descent normalises with (slot sim+1).  No CTA barrier between the
// phases (round 1's kernel walked all paths on warp 0 while the CTA's other seven warps sat at a __syncthreads: 12 stall cycles per
// issue); the only block-level step is the final min/max combine.  Rewards are +-1 for the winner's code stored in the record's
// `terminal` byte (0 = not terminal).
// BK_ENVS: 8 for trees of up to 64 nodes; 2 for larger ones, whose staged records would otherwise leave one CTA per SM (c3, T = 256:
// 38 ms per move with 8 envs per warp against 18 for round 1's kernel).
constexpr int BK_WARPS = 4;
template <int BK_ENVS>
__global__ void __launch_bounds__(BK_WARPS * 32) backup_kernel(bl_tree t, int sim) {
    extern __shared__ uint4 bsm[];
    __shared__ int red[2 * BK_WARPS];
    const int T = t.T, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nrec = sim + 1;                                   // node slots that can be populated so far
    const int b0 = (blockIdx.x * BK_WARPS + warp) * BK_ENVS;   // first env of this warp
    uint4 *rec = bsm + (size_t)warp * BK_ENVS * (nrec + 1);    // [BK_ENVS][nrec + 1]: the odd tail word keeps the walkers on different banks
    const int stride = nrec + 1;
    // walkers: lane e < BK_ENVS takes env b0 + e; its leaf and leaf value travel while the records do
    int leaf = -1;
    float val[2] = {0.f, 0.f};
    if (lane < BK_ENVS && b0 + lane < t.B) {
        leaf = t.leaf[b0 + lane];
        const uint32_t lv = reinterpret_cast<const uint32_t *>(t.leaf_v)[b0 + lane];        // = aux[leaf].v, without the dependent load
        if (leaf >= 0) { val[0] = bl_h2f((bl_half)(lv & 0xFFFF)); val[1] = bl_h2f((bl_half)(lv >> 16)); }
    }
    // the warp's BK_ENVS x nrec records as ONE flat item list over the lanes (an env's records are contiguous, so lanes still read
    // consecutive 16-byte records): with a loop per env, nrec = 33 cost two rounds per env — a 5 us step in the launch time at sim 32
    const int nenv = t.B - b0 < BK_ENVS ? (t.B - b0 < 0 ? 0 : t.B - b0) : BK_ENVS, nitem = nenv * nrec;
    for (int i = lane; i < nitem; i += 32) {
        const int e = i / nrec, k = i - e * nrec;
        rec[e * stride + k] = reinterpret_cast<const uint4 *>(t.node + (size_t)(b0 + e) * T)[k];
    }
    __syncwarp();
    unsigned visited = 0;
    if (lane < BK_ENVS) {
        uint4 *myrec = rec + lane * stride;
        for (int cur = leaf; cur >= 0;) {
            union { uint4 u; bl_node n; } x;
            x.u = myrec[cur];
            const float r0 = x.n.terminal == 1 ? 1.f : (x.n.terminal == 2 ? -1.f : 0.f);
            const float rw[2] = {r0, x.n.terminal == 1 ? -1.f : (x.n.terminal == 2 ? 1.f : 0.f)};      // +0, never -0
#pragma unroll
            for (int s = 0; s < 2; s++) {
                if (x.n.terminal) val[s] = 0.f;
                val[s] = __fadd_rn(val[s], rw[s]);
                x.n.w[s] = bl_f2h(__fadd_rn(bl_h2f(x.n.w[s]), bl_h2f(bl_f2h(val[s]))));
            }
            x.n.n = (int16_t)(x.n.n + t.Sn);                    // quirk: +1 per seat (cuda.cu:228)
            x.n.seat |= 0x80;                                   // "touched", in the staged copy only (the seat is 0 or 1)
            myrec[cur] = x.u;
            cur = x.n.parent;
            visited++;
        }
    }
    visited = __reduce_add_sync(0xffffffffu, visited);
    __syncwarp();
    float lo = BL_INF, hi = -BL_INF;
    for (int i = lane; i < nitem; i += 32) {
        const int e = i / nrec, k = i - e * nrec;
        union { uint4 u; bl_node n; } x;
        x.u = rec[e * stride + k];
        if (x.n.seat & 0x80) {
            x.n.seat &= 0x7F;
            reinterpret_cast<uint2 *>(t.node + (size_t)(b0 + e) * T + k)[1] = make_uint2(x.u.z, x.u.w);   // (n, w, seat, terminal)
        }
        // w/(n+1e-4) through the branch-free exact division (divisors in [1e-4, 32768]: always in its safe range, as bl_qnorm::fast)
        const float den = __fadd_rn((float)x.n.n, 1.e-4f);
        const float q0 = bl_div_fast(bl_h2f(x.n.w[0]), den), q1 = bl_div_fast(bl_h2f(x.n.w[1]), den);
        lo = fminf(lo, fminf(q0, q1));
        hi = fmaxf(hi, fmaxf(q0, q1));
    }
    if (nrec < T) { lo = fminf(lo, 0.f); hi = fmaxf(hi, 0.f); }         // untouched slots: w = 0, n = 0 -> q = 0
    const int klo = __reduce_min_sync(0xffffffffu, bl_f2ord(lo)), khi = __reduce_max_sync(0xffffffffu, bl_f2ord(hi));
    if (lane == 0) {
        red[warp] = klo; red[BK_WARPS + warp] = khi;
        if (visited) atomicAdd(reinterpret_cast<unsigned long long *>(t.counters + C_BACKUP_NODES), (unsigned long long)visited);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int a = red[0], c = red[BK_WARPS];
#pragma unroll
        for (int w = 1; w < BK_WARPS; w++) { a = min(a, red[w]); c = max(c, red[BK_WARPS + w]); }
        int *qr = reinterpret_cast<int *>(t.qrange) + 2 * (sim + 1);
        atomicMin(qr, a);
        atomicMax(qr + 1, c);
    }
}

// ---- root ------------------------------------------------------------------------------------------------------------
// MCTS.root (boardlaw/mcts/__init__.py:142-149, root_kernel of cuda.cu:107-136) + the agent's action (mcts/__init__.py:220-221), one
// WARP per env: the lanes share the root's actions (3 of 81 each at 9x9) for everything that is independent per action — lambda*pi,
// the children's q (found through the env's parent_of row, one candidate node per lane), the two IEEE divisions per action and
// Newton pass, the final probabilities / half / log table / logits store (coalesced) — and the reference's sequential sums run on two
// lanes (S on lane 0, g on lane 1) over the terms parked in shared memory, in the reference's order: same operations, operands and
// order as bl_newton / bl_prob, so the logits are bit-identical.  (Round 1's kernel ran the whole evaluation serially on one lane per
// env: 230 us per move at c2.)
constexpr int RW_WARPS = 8;
__global__ void __launch_bounds__(RW_WARPS * 32) root_kernel(bl_tree t, int sim, const bl_half *__restrict__ log_lut,
                                                            bl_half *__restrict__ logits, bl_half *__restrict__ v,
                                                            int64_t *__restrict__ n_leaves, int64_t *__restrict__ actions,
                                                            const float *__restrict__ uniforms, int greedy, uint64_t seed) {
    extern __shared__ __align__(16) float rsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int A = t.A, T = t.T, Sn = t.Sn, AP = t.AP, TP = (T + 7) & ~7;
    const int b = blockIdx.x * RW_WARPS + warp;
    if (b >= t.B) return;
    float *top = rsm + (size_t)warp * 4 * AP, *q = top + AP, *sS = q + AP, *sG = sS + AP;
    const size_t node0 = (size_t)b * T;
    const float *row = t.pi + node0 * AP;
    for (int a = lane; a < AP; a += 32) { top[a] = a < A ? row[a] : 0.f; q[a] = 0.f; sS[a] = 0.f; sG[a] = 0.f; }     // (pad entries stay +0: they add nothing to the sums)
    __syncwarp();
    const bl_qnorm qn(t.qrange + 2 * sim);
    const bl_node root = bl_ld_node(t.node + node0);
    const int seat = root.seat;
    int N = 0, nc = 0;
    for (int k = 1 + lane; k < sim && k < T; k += 32) {              // the root's children: nodes whose parent is node 0
        if (t.parent_of[(size_t)b * TP + k] == 0) {
            const bl_node ch = bl_ld_node(t.node + node0 + k);
            q[ch.relation] = qn(seat ? ch.w[1] : ch.w[0], ch.n);
            N += ch.n;
            nc++;
        }
    }
    N = __reduce_add_sync(0xffffffffu, N);
    nc = __reduce_add_sync(0xffffffffu, nc);
    N += A - nc;
    const float lambda = bl_lambda(bl_h2f(t.c_puct[b]), N, A);
    __syncwarp();
    float alpha = 0.f;                                               // newton_search's seed (cuda.cu:44-50): a maximum, order-free
    for (int a = lane; a < A; a += 32) {
        const float tv = __fmul_rn(lambda, top[a]);
        top[a] = tv;
        alpha = fmaxf(alpha, __fadd_rn(q[a], fmaxf(tv, 1.e-4f)));
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) alpha = fmaxf(alpha, __shfl_xor_sync(0xffffffffu, alpha, o));
    float error = BL_INF;
    for (int it = 0; it < 100;) {
        for (int a = lane; a < A; a += 32) {
            const float tv = top[a], bot = __fsub_rn(alpha, q[a]);
            sS[a] = __fdiv_rn(tv, bot);
            sG[a] = __fdiv_rn(-tv, __fmul_rn(bot, bot));
        }
        __syncwarp();
        float acc = 0.f;
        if (lane < 2) {                                              // the reference's order a = 0, 1, 2, ...; four terms per 128-bit load
            const float4 *src = reinterpret_cast<const float4 *>(lane == 0 ? sS : sG);
#pragma unroll 3
            for (int c = 0; c < (AP >> 2); c++) {
                const float4 x = src[c];
                acc = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(acc, x.x), x.y), x.z), x.w);
            }
        }
        const float S = __shfl_sync(0xffffffffu, acc, 0), g = __shfl_sync(0xffffffffu, acc, 1);
        __syncwarp();
        it++;
        const float new_error = __fsub_rn(S, 1.f);
        if ((new_error < 1e-3f) || (error == new_error)) break;
        alpha = __fsub_rn(alpha, __fdiv_rn(new_error, g));
        error = new_error;
    }
    // probs -> half -> log -> half (MCTS.root, boardlaw/mcts/__init__.py:142-149); log through the host-libm table
    for (int a = lane; a < A; a += 32) {
        const bl_half h = log_lut[bl_f2h(bl_prob(top[a], q[a], alpha))];
        logits[(size_t)b * A + a] = h;
        if (actions) { sS[a] = bl_h2f(h); sG[a] = expf(bl_h2f(h)); }     // (the term rows are dead: log-probabilities and weights for the draw)
    }
    __syncwarp();
    if (actions) {
        // MCTSAgent.__call__ (boardlaw/mcts/__init__.py:220-221): argmax of the root policy (the first of equal maxima), or a draw from
        // Categorical(logits) with the weights summed in action order
        int act;
        if (greedy) {
            float best = -BL_INF;
            int arg = A;
            for (int a = lane; a < A; a += 32)
                if (sS[a] > best) { best = sS[a]; arg = a; }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
                if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
            }
            act = arg < A ? arg : -1;
        } else {
            act = -1;
            if (lane == 0) {
                const float4 *w4 = reinterpret_cast<const float4 *>(sG);
                float tot = 0.f;
#pragma unroll 3
                for (int c = 0; c < (AP >> 2); c++) { const float4 x = w4[c]; tot += x.x; tot += x.y; tot += x.z; tot += x.w; }
                const float u = uniforms ? uniforms[b] : ((float)(bl_philox(seed ^ (t.counters[C_MOVE] * 0x9E3779B97F4A7C15ull), (uint64_t)b, 0xAC71ull).x >> 8) + .5f) * (1.f / 16777216.f);
                const float target = u * tot;
                float cum = 0.f;
                bool done = false;
                for (int c = 0; c < (AP >> 2) && !done; c++) {
                    const float4 x = w4[c];
                    const float w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        if (!done) {
                            cum += w[j];
                            if (w[j] > 0.f) { act = 4 * c + j; done = cum >= target; }
                        }
                    }
                }
            }
        }
        if (lane == 0) actions[b] = act;
    }
    if (lane < Sn) v[(size_t)b * Sn + lane] = bl_ld_aux(t.aux + node0).v[lane];
    int leaves = 0;
    for (int k = 1 + lane; k < T; k += 32) {
        const bl_node nd = bl_ld_node(t.node + node0 + k);
        leaves += (nd.parent != -1) && (nd.first_child == -1);
    }
    leaves = __reduce_add_sync(0xffffffffu, leaves);
    if (lane == 0) n_leaves[b] = leaves;
}

__global__ void __launch_bounds__(256) children_dense_kernel(bl_tree t, int16_t *__restrict__ children) {
    const long long n = (long long)t.B * t.T * t.A;
    const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
    for (long long i = i0; i < n; i += stride) children[i] = -1;
}
__global__ void __launch_bounds__(256) children_scatter_kernel(bl_tree t, int16_t *__restrict__ children) {
    const long long n = (long long)t.B * t.T;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const bl_node nd = bl_ld_node(t.node + i);
        if (nd.parent < 0) continue;
        long long b = i / t.T; int k = (int)(i - b * t.T);
        children[((b * t.T) + nd.parent) * t.A + nd.relation] = (int16_t)k;
    }
}

// leaf boards -> compact (B,A) boards + (B,) seats for the network
__global__ void __launch_bounds__(256) gather_leaves_kernel(bl_tree t, int node, uint8_t *__restrict__ board,
                                                            int32_t *__restrict__ seats) {
    const long long n = (long long)t.B * t.A;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        long long b = i / t.A; int c = (int)(i - b * t.A);
        int nd = node >= 0 ? node : max((int)t.leaf[b], 0);
        board[i] = t.board[(b * t.T + nd) * t.BP + c];
        if (c == 0) seats[b] = t.node[b * t.T + nd].seat;
    }
}

}  // namespace
// net_tc.cu
int bl_fc_forward_tc_tree(const bl_fc_params *p, const bl_tree *t, cudaStream_t st);
bool bl_fc_tc_supported(const bl_fc_params *p);
// net_tc_wide.cu
int bl_fc_forward_wide_tree(const bl_fc_params *p, const bl_tree *t, void *scratch, cudaStream_t st);
bool bl_fc_wide_supported(const bl_fc_params *p);
namespace {

int grid1d(long long n, int block) {
    long long g = (n + block - 1) / block, cap = (long long)BL_NUM_SMS * 16;
    return (int)(g < 1 ? 1 : (g < cap ? g : cap));
}

int check_tree(const bl_tree *t) {
    if (!t || t->B < 0 || t->T < 1 || t->S < 1 || t->A != t->S * t->S || t->Sn != 2) return -1;
    if (t->A > 255 || t->T > 32767) return -1;
    if (t->AP < t->A || (t->AP & 3) || t->BP < t->A || (t->BP & 15)) return -1;
    return 0;
}

}  // namespace

extern "C" int bl_tree_reset(const bl_tree *t, const uint8_t *board, const int32_t *seats, float c_puct, bl_stream stream) {
    if (int e = check_tree(t)) return e;
    if (t->B == 0) return 0;
    __half h = __float2half_rn(c_puct);
    reset_kernel<<<grid1d((long long)t->B * t->A, 256), 256, 0, bl_cu(stream)>>>(*t, board, seats, __half_as_ushort(h));
    BL_LAUNCH_CHECK();
}

extern "C" int bl_tree_set_eval(const bl_tree *t, int node, const void *logits, const void *v, int inputs_are_half,
                                bl_stream stream) {
    if (int e = check_tree(t)) return e;
    if (t->B == 0) return 0;
    if (node >= t->T) return -1;
    int grid = grid1d((long long)t->B * 32, 256);            // one warp per env
    const PriorMix none = {nullptr, nullptr, nullptr, 0.f, 0.f, 0ull};
    if (inputs_are_half) set_eval_kernel<1><<<grid, 256, 0, bl_cu(stream)>>>(*t, node, logits, v, none);
    else set_eval_kernel<0><<<grid, 256, 0, bl_cu(stream)>>>(*t, node, logits, v, none);
    BL_LAUNCH_CHECK();
}

extern "C" int bl_tree_set_root_prior(const bl_tree *t, const float *logits, const float *v, const uint8_t *board, const int32_t *seats,
                                      const float *draw, float noise_eps, float alpha_scale, uint64_t seed, bl_stream stream) {
    if (int e = check_tree(t)) return e;
    if (t->B == 0) return 0;
    const PriorMix mix = {board, seats, draw, noise_eps, alpha_scale / (float)t->A, seed};
    set_eval_kernel<2><<<grid1d((long long)t->B * 32, 256), 256, 0, bl_cu(stream)>>>(*t, 0, logits, v, mix);
    BL_LAUNCH_CHECK();
}

extern "C" int bl_tree_descend_expand(const bl_tree *t, int sim, const bl_half *rands, uint64_t seed, bl_stream stream) {
    if (int e = check_tree(t)) return e;
    if (t->B == 0) return 0;
    if (sim < 1 || sim >= t->T) return -1;
    static bool env_read = false;
    if (!env_read) {                                          // BL_DESCEND_VARIANT=1|2|3 overrides the default for tuning runs
        env_read = true;
        if (const char *e = getenv("BL_DESCEND_VARIANT")) { const int v = atoi(e); if (v >= 0 && v <= 7 && ((v != 4 && v != 6 && v != 7) || bl_experimental_built())) g_descend_variant = v; }
    }
    if (g_descend_variant == 7) {
        const int rc = bl_descend_pk(t, sim, rands, seed, bl_cu(stream));
        if (rc != -2 && rc != -3) return rc;                 // unsupported shape: the other kernels take it
    }
    if (g_descend_variant == 6 && t->cpi) {
        const int rc = bl_descend_all(t, sim, rands, seed, bl_cu(stream));
        if (rc != -2 && rc != -3) return rc;
    }
    if (g_descend_variant == 5 && t->cpi) {
        const int rc = bl_descend_fx(t, sim, rands, seed, bl_cu(stream));
        if (rc != -2 && rc != -3) return rc;                 // unsupported shape / scratch: the exact kernels take it
    }
    if (g_descend_variant == 4) {
        const int rc = bl_descend_pc(t, sim, rands, seed, bl_cu(stream));
        if (rc != -2 && rc != -3) return rc;
    }
    if (g_descend_variant == 3 || (g_descend_variant == 0 && t->A > 81)) {
        const int rc = bl_descend_mw(t, sim, rands, seed, bl_cu(stream));
        if (rc != -2 && rc != -3) return rc;                 // unsupported shape / scratch: the one-lane kernel takes it
    }
    if (g_descend_variant != 1) return bl_descend_v3(t, sim, rands, seed, bl_cu(stream));
    size_t smem = descend_smem(t->A);
    if (smem > 227 * 1024) return -2;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(descend_expand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    descend_expand_kernel<<<(t->B + ENT - 1) / ENT, ENT, smem, bl_cu(stream)>>>(*t, sim, rands, seed);
    if (cudaError_t e = cudaGetLastError()) return (int)e;
    return bl_expand_step(t, sim, bl_cu(stream));
}

extern "C" int bl_debug_set_descend_variant(int variant) {
    if (variant < 0 || variant > 7) return -1;
    if ((variant == 4 || variant == 6 || variant == 7) && !bl_experimental_built()) return -2;     // not compiled in
    g_descend_variant = variant;
    return 0;
}

extern "C" int bl_tree_backup(const bl_tree *t, int sim, bl_stream stream) {
    if (int e = check_tree(t)) return e;
    if (t->B == 0) return 0;
    if (sim < 1 || sim >= t->T) return -1;
    const int envs = t->T <= 64 ? 8 : 2;
    const size_t smem = (size_t)BK_WARPS * envs * (sim + 2) * sizeof(uint4);
    if (smem > 226 * 1024) return -2;
    if (smem + 256 > 48 * 1024) {                                // (the kernel's static shared memory counts towards the default limit)
        const int full = (int)(BK_WARPS * envs * (t->T + 1) * sizeof(uint4));
        cudaError_t e = envs == 8 ? cudaFuncSetAttribute(backup_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, full)
                                  : cudaFuncSetAttribute(backup_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, full);
        if (e != cudaSuccess) return (int)e;
    }
    const int per_cta = BK_WARPS * envs;
    if (envs == 8) backup_kernel<8><<<(t->B + per_cta - 1) / per_cta, BK_WARPS * 32, smem, bl_cu(stream)>>>(*t, sim);
    else backup_kernel<2><<<(t->B + per_cta - 1) / per_cta, BK_WARPS * 32, smem, bl_cu(stream)>>>(*t, sim);
    BL_LAUNCH_CHECK();
}

extern "C" int64_t bl_tree_eval_scratch_bytes(const bl_tree *t, const bl_fc_params *p) {
    // compact boards + seats + fp32 logits + fp32 v + the network's own scratch
    int64_t B = t->B;
    return ((B * t->A + 255) & ~255ll) + ((B * 4 + 255) & ~255ll) + ((B * t->A * 4 + 255) & ~255ll) +
           ((B * 2 * 4 + 255) & ~255ll) + bl_fc_scratch_bytes(p, t->B);
}

namespace {
struct EvalScratch { uint8_t *board; int32_t *seats; float *logits; float *v; void *net; };
EvalScratch split_scratch(const bl_tree *t, void *scratch) {
    int64_t B = t->B;
    uint8_t *p = reinterpret_cast<uint8_t *>(scratch);
    EvalScratch s;
    s.board = p; p += (B * t->A + 255) & ~255ll;
    s.seats = reinterpret_cast<int32_t *>(p); p += (B * 4 + 255) & ~255ll;
    s.logits = reinterpret_cast<float *>(p); p += (B * t->A * 4 + 255) & ~255ll;
    s.v = reinterpret_cast<float *>(p); p += (B * 2 * 4 + 255) & ~255ll;
    s.net = p;
    return s;
}
}  // namespace

extern "C" int bl_tree_eval_leaves(const bl_tree *t, const bl_fc_params *p, int sim, void *scratch, bl_stream stream) {
    if (int e = check_tree(t)) return e;
    if (t->B == 0) return 0;
    (void)sim;
    // tensor-core path: one kernel reads the leaf boards from the tree and writes pi rows / summaries / values back into it
    if (bl_fc_tc_supported(p)) return bl_fc_forward_tc_tree(p, t, bl_cu(stream));
    EvalScratch s = split_scratch(t, scratch);
    if (bl_fc_wide_supported(p)) return bl_fc_forward_wide_tree(p, t, s.net, bl_cu(stream));
    gather_leaves_kernel<<<grid1d((long long)t->B * t->A, 256), 256, 0, bl_cu(stream)>>>(*t, -1, s.board, s.seats);
    int e = bl_fc_forward(p, s.board, s.seats, s.logits, s.v, s.net, t->B, stream);
    if (e) return e;
    return bl_tree_set_eval(t, -1, s.logits, s.v, 0, stream);
}

extern "C" int bl_tree_eval_root(const bl_tree *t, const bl_fc_params *p, float *logits, float *v, void *scratch,
                                 bl_stream stream) {
    if (int e = check_tree(t)) return e;
    if (t->B == 0) return 0;
    EvalScratch s = split_scratch(t, scratch);
    gather_leaves_kernel<<<grid1d((long long)t->B * t->A, 256), 256, 0, bl_cu(stream)>>>(*t, 0, s.board, s.seats);
    return bl_fc_forward(p, s.board, s.seats, logits, v, s.net, t->B, stream);
}

extern "C" int bl_tree_root_act(const bl_tree *t, int sim, const bl_half *log_lut, bl_half *logits, bl_half *v, int64_t *n_leaves,
                                int64_t *actions, const float *uniforms, int greedy, uint64_t seed, bl_stream stream) {
    if (int e = check_tree(t)) return e;
    if (t->B == 0) return 0;
    if (sim < 1 || sim > t->T || !actions) return -1;
    const size_t smem = (size_t)RW_WARPS * 4 * t->AP * sizeof(float);
    if (smem > 227 * 1024) return -2;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(root_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    root_kernel<<<(t->B + RW_WARPS - 1) / RW_WARPS, RW_WARPS * 32, smem, bl_cu(stream)>>>(*t, sim, log_lut, logits, v, n_leaves, actions, uniforms, greedy, seed);
    BL_LAUNCH_CHECK();
}

// One (B, R) uint8 trajectory record per env for the move just played — the row selfplay.pack_records assembles with eight torch
// casts and two concatenations: board A u8 | seat u8 | terminal u8 | action i16 | rewards 2 x f16 | v 2 x f16 | logits A x f16 |
// prior A x f16 | zero padding to R (a multiple of 16).  One warp per env, byte-granular (the fields are not aligned).
__global__ void __launch_bounds__(256) pack_records_kernel(const uint8_t *__restrict__ board, const int32_t *__restrict__ seats, const uint8_t *__restrict__ terminal,
                                                           const int64_t *__restrict__ actions, const float *__restrict__ rewards, const bl_half *__restrict__ v,
                                                           const bl_half *__restrict__ logits, const bl_half *__restrict__ prior, uint8_t *__restrict__ rec,
                                                           int B, int A, int R) {
    const int lane = threadIdx.x & 31, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < B; b += nwarps) {
        const uint16_t act = (uint16_t)(int16_t)actions[b];
        const bl_half r0 = bl_f2h(rewards[2 * b]), r1 = bl_f2h(rewards[2 * b + 1]);
        for (int o = lane; o < R; o += 32) {
            uint8_t x = 0;
            int f = o;
            if (f < A) x = board[(size_t)b * A + f];
            else if ((f -= A) == 0) x = (uint8_t)seats[b];
            else if (f == 1) x = terminal[b];
            else if (f < 4) x = (uint8_t)(act >> (8 * (f - 2)));
            else if (f < 8) x = (uint8_t)((f < 6 ? r0 : r1) >> (8 * (f & 1)));
            else if (f < 12) x = (uint8_t)(v[2 * b + ((f - 8) >> 1)] >> (8 * (f & 1)));
            else if ((f -= 12) < 2 * A) x = (uint8_t)(logits[(size_t)b * A + (f >> 1)] >> (8 * (f & 1)));
            else if ((f -= 2 * A) < 2 * A) x = (uint8_t)(prior[(size_t)b * A + (f >> 1)] >> (8 * (f & 1)));
            rec[(size_t)b * R + o] = x;
        }
    }
}

extern "C" int bl_pack_records(const uint8_t *board, const int32_t *seats, const uint8_t *terminal, const int64_t *actions, const float *rewards,
                               const bl_half *v, const bl_half *logits, const bl_half *prior, uint8_t *records, int B, int A, int R, bl_stream stream) {
    if (B == 0) return 0;
    if (R < 5 * A + 12) return -1;
    pack_records_kernel<<<grid1d((long long)B * 32, 256), 256, 0, bl_cu(stream)>>>(board, seats, terminal, actions, rewards, v, logits, prior, records, B, A, R);
    BL_LAUNCH_CHECK();
}

extern "C" int bl_tree_root(const bl_tree *t, int sim, const bl_half *log_lut, bl_half *logits, bl_half *v,
                            int64_t *n_leaves, bl_stream stream) {
    if (int e = check_tree(t)) return e;
    if (t->B == 0) return 0;
    if (sim < 1 || sim > t->T) return -1;
    const size_t smem = (size_t)RW_WARPS * 4 * t->AP * sizeof(float);
    if (smem > 227 * 1024) return -2;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(root_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    root_kernel<<<(t->B + RW_WARPS - 1) / RW_WARPS, RW_WARPS * 32, smem, bl_cu(stream)>>>(*t, sim, log_lut, logits, v, n_leaves, nullptr, nullptr, 0, 0ull);
    BL_LAUNCH_CHECK();
}

extern "C" int bl_tree_children_dense(const bl_tree *t, int16_t *children, bl_stream stream) {
    if (int e = check_tree(t)) return e;
    if (t->B == 0) return 0;
    children_dense_kernel<<<grid1d((long long)t->B * t->T * t->A, 256), 256, 0, bl_cu(stream)>>>(*t, children);
    children_scatter_kernel<<<grid1d((long long)t->B * t->T, 256), 256, 0, bl_cu(stream)>>>(*t, children);
    BL_LAUNCH_CHECK();
}
