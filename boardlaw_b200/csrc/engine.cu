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
#include "engine_internal.cuh"
#include "hex_core.cuh"
#include "mcts_core.cuh"

namespace {

constexpr int ENT = 64;            // lanes (= envs) per CTA
constexpr int FP = ENT + 1;        // float column pitch (odd: conflict-free for row-wise fills and lane-wise reads)
constexpr int BPITCH = ENT + 4;    // byte column pitch (17 words: consecutive cells land on distinct banks)

// counters slots
int g_descend_variant = 2;

__global__ void __launch_bounds__(256) reset_kernel(bl_tree t, const uint8_t *__restrict__ board,
                                                    const int32_t *__restrict__ seats, bl_half c_puct) {
    const long long BT = (long long)t.B * t.T;
    const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
    for (long long i = i0; i < BT; i += stride) {
        t.parents[i] = -1; t.relation[i] = -1; t.first_child[i] = -1; t.next_sib[i] = -1;
        t.n[i] = 0; t.terminal[i] = 0;
        long long b = i / t.T;
        t.seats[i] = (uint8_t)seats[b];
        for (int s = 0; s < t.Sn; s++) { t.w[i * t.Sn + s] = 0; t.rewards[i * t.Sn + s] = 0; }
    }
    for (long long i = i0; i < (long long)t.B * t.A; i += stride) {
        long long b = i / t.A; int c = (int)(i - b * t.A);
        t.board[(b * t.T) * t.BP + c] = board[i];
    }
    for (long long i = i0; i < t.B; i += stride) t.c_puct[i] = c_puct;
    int *qr = reinterpret_cast<int *>(t.qrange);
    for (long long i = i0; i <= t.T; i += stride) {
        // slot 1 serves the first descent: the all-zero tree has (min,max) = (0,0)
        qr[2 * i] = (i == 1) ? bl_f2ord(0.f) : bl_f2ord(BL_INF);
        qr[2 * i + 1] = (i == 1) ? bl_f2ord(0.f) : bl_f2ord(-BL_INF);
    }
}

// logits/v of one node per env -> pi row (+ prior when node 0) + v, rounding through half like decisions.half()
template <bool HALF_IN>
__global__ void __launch_bounds__(256) set_eval_kernel(bl_tree t, int node, const void *__restrict__ logits_,
                                                       const void *__restrict__ v_) {
    const long long n = (long long)t.B * t.A;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        long long b = i / t.A; int a = (int)(i - b * t.A);
        int nd = node >= 0 ? node : t.leaf[b];
        if (nd < 0) continue;
        bl_half h = HALF_IN ? reinterpret_cast<const bl_half *>(logits_)[i] : bl_f2h(reinterpret_cast<const float *>(logits_)[i]);
        t.pi[(b * t.T + nd) * t.AP + a] = t.exp_lut[h];
        if (nd == 0) t.prior[i] = h;
        if (t.logits) t.logits[(b * t.T + nd) * t.A + a] = h;
        if (a < t.Sn) {
            long long j = b * t.Sn + a;
            bl_half hv = HALF_IN ? reinterpret_cast<const bl_half *>(v_)[j] : bl_f2h(reinterpret_cast<const float *>(v_)[j]);
            t.v[(b * t.T + nd) * t.Sn + a] = hv;
        }
    }
}

// ---- descend + expand + env step -------------------------------------------------------------------------------
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
    const int A = t.A, T = t.T, Sn = t.Sn;
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
        bool active = cur >= 0 && !t.terminal[node0 + cur];
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
            const int seat = t.seats[node0 + cur];
            int N = 0, nc = 0;
            for (int c = t.first_child[node0 + cur]; c >= 0; c = t.next_sib[node0 + c]) {
                int16_t nn = t.n[node0 + c];
                q[t.relation[node0 + c] * FP] = qn(t.w[(node0 + c) * Sn + seat], nn);
                N += nn;
                nc++;
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
            for (int c = t.first_child[node0 + cur]; c >= 0; c = t.next_sib[node0 + c]) {
                int rel = t.relation[node0 + c];
                q[rel * FP] = 0.f;                              // leave the column zeroed for the next node
                if (rel == action) next = c;
            }
            cur = action >= 0 ? next : -2;                      // -2: no positive-probability action (error)
            c_evals++; c_children += nc; c_iters += it;
        }
        __syncwarp();
    }

    // expand (boardlaw/mcts/__init__.py:117-122)
    const bool ok = in_range && action >= 0 && cur != -2;
    int leaf = -1;
    if (ok) {
        if (cur >= 0) leaf = cur;                               // stopped at an existing terminal child: reuse its slot
        else {
            leaf = sim;
            t.parents[node0 + sim] = (int16_t)parent;
            t.relation[node0 + sim] = (int16_t)action;
            t.next_sib[node0 + sim] = t.first_child[node0 + parent];
            t.first_child[node0 + parent] = (int16_t)sim;
        }
    } else if (in_range) {
        atomicAdd(reinterpret_cast<unsigned long long *>(t.counters + C_ERRORS), 1ull);
    }
    if (in_range) {
        t.leaf[b] = (int16_t)leaf;
        t.leaf_parent[b] = (int16_t)parent;
        t.leaf_action[b] = (int16_t)action;
    }

    // env step of the parent's board into the leaf slot (boardlaw/mcts/__init__.py:124-129, Hex.step)
    unsigned omask = __ballot_sync(0xffffffffu, ok);
    for (unsigned m = omask; m; m &= m - 1) {
        int l = __ffs(m) - 1;
        int pl = __shfl_sync(0xffffffffu, parent, l);
        const uint8_t *row = t.board + ((size_t)(bw + l) * T + pl) * t.BP;
        for (int c = lane; c < A; c += 32) sm.bd[c * BPITCH + wbase + l] = row[c];
    }
    __syncwarp();
    if (ok) {
        const int seat = t.seats[node0 + parent];
        int win = bl_hex_place<uint8_t>(sm.bd + tid, sm.stk + tid, BPITCH, t.S, seat, action);
        float r0 = win == 1 ? 1.f : (win == 2 ? -1.f : 0.f), r1 = win == 1 ? -1.f : (win == 2 ? 1.f : 0.f);
        t.rewards[(node0 + leaf) * Sn + 0] = bl_f2h(r0);
        t.rewards[(node0 + leaf) * Sn + 1] = bl_f2h(r1);
        t.terminal[node0 + leaf] = win != 0;
        t.seats[node0 + leaf] = win ? 0 : (uint8_t)(1 - seat);
        if (win)
            for (int c = 0; c < A; c++) sm.bd[c * BPITCH + tid] = 0;   // auto-reset (hex/__init__.py:185-188)
    }
    __syncwarp();
    for (unsigned m = omask; m; m &= m - 1) {
        int l = __ffs(m) - 1;
        int ll = __shfl_sync(0xffffffffu, leaf, l);
        uint8_t *row = t.board + ((size_t)(bw + l) * T + ll) * t.BP;
        for (int c = lane; c < A; c += 32) row[c] = sm.bd[c * BPITCH + wbase + l];
    }
    bl_count(t.counters, C_EVALS, c_evals);
    bl_count(t.counters, C_CHILDREN, c_children);
    bl_count(t.counters, C_ITERS, c_iters);
    bl_count(t.counters, C_DESCENTS, ok ? 1u : 0u);
}

// ---- backup + q-range scan ----------------------------------------------------------------------------------------
constexpr int BNT = 128;
__global__ void __launch_bounds__(BNT) backup_kernel(bl_tree t, int sim) {
    const int b = blockIdx.x * BNT + threadIdx.x;
    const int T = t.T, Sn = t.Sn;
    unsigned visited = 0;
    if (b < t.B) {
        const size_t base = (size_t)b * T;
        int cur = t.leaf[b];
        float val[2] = {0.f, 0.f};
        if (cur >= 0) { val[0] = bl_h2f(t.v[(base + cur) * Sn]); val[1] = bl_h2f(t.v[(base + cur) * Sn + 1]); }
        while (cur >= 0) {
            const size_t node = base + cur;
            const bool term = t.terminal[node];
#pragma unroll
            for (int s = 0; s < 2; s++) {
                if (term) val[s] = 0.f;
                val[s] = __fadd_rn(val[s], bl_h2f(t.rewards[node * Sn + s]));
                t.w[node * Sn + s] = bl_f2h(__fadd_rn(bl_h2f(t.w[node * Sn + s]), bl_h2f(bl_f2h(val[s]))));
            }
            t.n[node] = (int16_t)(t.n[node] + Sn);              // quirk: +1 per seat (cuda.cu:228)
            cur = t.parents[node];
            visited++;
        }
    }
    __syncthreads();
    // q-range of the block's envs (contiguous (w, n) spans): feeds the NEXT descent (slot sim+1)
    const int nb = min(BNT, t.B - blockIdx.x * BNT);
    const size_t first = (size_t)blockIdx.x * BNT * T;
    float lo = BL_INF, hi = -BL_INF;
    for (int i = threadIdx.x; i < nb * T; i += BNT) {
        int16_t nn = t.n[first + i];
        const __half2 ww = reinterpret_cast<const __half2 *>(t.w)[first + i];
        float q0 = bl_qraw(__half_as_ushort(__low2half(ww)), nn), q1 = bl_qraw(__half_as_ushort(__high2half(ww)), nn);
        lo = fminf(lo, fminf(q0, q1));
        hi = fmaxf(hi, fmaxf(q0, q1));
    }
    int klo = __reduce_min_sync(0xffffffffu, bl_f2ord(lo)), khi = __reduce_max_sync(0xffffffffu, bl_f2ord(hi));
    if (bl_lane() == 0) {
        int *qr = reinterpret_cast<int *>(t.qrange) + 2 * (sim + 1);
        atomicMin(qr, klo);
        atomicMax(qr + 1, khi);
    }
    bl_count(t.counters, C_BACKUP_NODES, visited);
}

// ---- root ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ENT) root_kernel(bl_tree t, int sim, const bl_half *__restrict__ log_lut,
                                                   bl_half *__restrict__ logits, bl_half *__restrict__ v,
                                                   int64_t *__restrict__ n_leaves) {
    extern __shared__ __align__(16) uint8_t raw[];
    Smem sm = carve(raw, t.A);
    const int tid = threadIdx.x, lane = tid & 31, wbase = tid & ~31;
    const int A = t.A, T = t.T, Sn = t.Sn;
    const int b = blockIdx.x * ENT + tid, bw = blockIdx.x * ENT + wbase;
    const bool in_range = b < t.B;
    const size_t node0 = (size_t)(in_range ? b : 0) * T;
    float *top = sm.top + tid, *q = sm.q + tid;
    unsigned mask = __ballot_sync(0xffffffffu, in_range);
    for (unsigned m = mask; m; m &= m - 1) {
        int l = __ffs(m) - 1;
        const float *row = t.pi + ((size_t)(bw + l) * T) * t.AP;
        for (int a = lane; a < A; a += 32) sm.top[a * FP + wbase + l] = row[a];
    }
    __syncwarp();
    if (!in_range) return;
    for (int a = 0; a < A; a++) q[a * FP] = 0.f;
    const bl_qnorm qn(t.qrange + 2 * sim);
    const int seat = t.seats[node0];
    int N = 0, nc = 0;
    for (int c = t.first_child[node0]; c >= 0; c = t.next_sib[node0 + c]) {
        int16_t nn = t.n[node0 + c];
        q[t.relation[node0 + c] * FP] = qn(t.w[(node0 + c) * Sn + seat], nn);
        N += nn;
        nc++;
    }
    N += A - nc;
    const float lambda = bl_lambda(bl_h2f(t.c_puct[b]), N, A);
    for (int a = 0; a < A; a++) top[a * FP] = __fmul_rn(lambda, top[a * FP]);
    int it;
    const float alpha = bl_newton(top, q, FP, A, &it);
    // probs -> half -> log -> half (MCTS.root, boardlaw/mcts/__init__.py:142-149); log through the host-libm table
    for (int a = 0; a < A; a++)
        logits[(size_t)b * A + a] = log_lut[bl_f2h(bl_prob(top[a * FP], q[a * FP], alpha))];
    for (int s = 0; s < Sn; s++) v[(size_t)b * Sn + s] = t.v[node0 * Sn + s];
    int leaves = 0;
    for (int k = 1; k < T; k++) leaves += (t.parents[node0 + k] != -1) && (t.first_child[node0 + k] == -1);
    n_leaves[b] = leaves;
}

__global__ void __launch_bounds__(256) children_dense_kernel(bl_tree t, int16_t *__restrict__ children) {
    const long long n = (long long)t.B * t.T * t.A;
    const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
    for (long long i = i0; i < n; i += stride) children[i] = -1;
}
__global__ void __launch_bounds__(256) children_scatter_kernel(bl_tree t, int16_t *__restrict__ children) {
    const long long n = (long long)t.B * t.T;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int p = t.parents[i];
        if (p < 0) continue;
        long long b = i / t.T; int k = (int)(i - b * t.T);
        children[((b * t.T) + p) * t.A + t.relation[i]] = (int16_t)k;
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
        if (c == 0) seats[b] = t.seats[b * t.T + nd];
    }
}

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
    int grid = grid1d((long long)t->B * t->A, 256);
    if (inputs_are_half) set_eval_kernel<true><<<grid, 256, 0, bl_cu(stream)>>>(*t, node, logits, v);
    else set_eval_kernel<false><<<grid, 256, 0, bl_cu(stream)>>>(*t, node, logits, v);
    BL_LAUNCH_CHECK();
}

extern "C" int bl_tree_descend_expand(const bl_tree *t, int sim, const bl_half *rands, uint64_t seed, bl_stream stream) {
    if (int e = check_tree(t)) return e;
    if (t->B == 0) return 0;
    if (sim < 1 || sim >= t->T) return -1;
    if (g_descend_variant == 2) return bl_descend_v2(t, sim, rands, seed, bl_cu(stream));
    size_t smem = descend_smem(t->A);
    if (smem > 227 * 1024) return -2;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(descend_expand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    descend_expand_kernel<<<(t->B + ENT - 1) / ENT, ENT, smem, bl_cu(stream)>>>(*t, sim, rands, seed);
    BL_LAUNCH_CHECK();
}

extern "C" int bl_debug_set_descend_variant(int variant) {
    if (variant != 1 && variant != 2) return -1;
    g_descend_variant = variant;
    return 0;
}

extern "C" int bl_tree_backup(const bl_tree *t, int sim, bl_stream stream) {
    if (int e = check_tree(t)) return e;
    if (t->B == 0) return 0;
    if (sim < 1 || sim >= t->T) return -1;
    backup_kernel<<<(t->B + BNT - 1) / BNT, BNT, 0, bl_cu(stream)>>>(*t, sim);
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
    EvalScratch s = split_scratch(t, scratch);
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

extern "C" int bl_tree_root(const bl_tree *t, int sim, const bl_half *log_lut, bl_half *logits, bl_half *v,
                            int64_t *n_leaves, bl_stream stream) {
    if (int e = check_tree(t)) return e;
    if (t->B == 0) return 0;
    if (sim < 1 || sim > t->T) return -1;
    size_t smem = descend_smem(t->A);
    if (smem > 227 * 1024) return -2;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(root_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    root_kernel<<<(t->B + ENT - 1) / ENT, ENT, smem, bl_cu(stream)>>>(*t, sim, log_lut, logits, v, n_leaves);
    BL_LAUNCH_CHECK();
}

extern "C" int bl_tree_children_dense(const bl_tree *t, int16_t *children, bl_stream stream) {
    if (int e = check_tree(t)) return e;
    if (t->B == 0) return 0;
    children_dense_kernel<<<grid1d((long long)t->B * t->T * t->A, 256), 256, 0, bl_cu(stream)>>>(*t, children);
    children_scatter_kernel<<<grid1d((long long)t->B * t->T, 256), 256, 0, bl_cu(stream)>>>(*t, children);
    BL_LAUNCH_CHECK();
}
