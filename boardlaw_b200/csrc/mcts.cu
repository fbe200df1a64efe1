// MCTS ops on the reference's tensor layout, for sm_100a: transition_q, descend, root, backup.
//
// Replaces mctscuda.{descend,root,backup} (boardlaw/mcts/cpp/cuda.cu:101-248).  These are the op-level
// drop-ins behind boardlaw.mcts.cuda; the self-play hot path uses the fused engine (engine.cu), which
// shares the arithmetic in mcts_core.cuh.
//
// Compiled with -fmad=false -prec-div=true -ftz=false (see build.py): bit-exact arithmetic contract.
#include "mcts_core.cuh"

namespace {

// ---- transition_q: global (min, max) of w/(n+1e-4) over (B,T,Sn) -----------------------------------------
__global__ void qrange_init_kernel(int *qrange) {
    qrange[0] = bl_f2ord(BL_INF);
    qrange[1] = bl_f2ord(-BL_INF);
}

__global__ void __launch_bounds__(256) qrange_kernel(const bl_half *__restrict__ w, const int16_t *__restrict__ n,
                                                     int *__restrict__ qrange, long long BT, int Sn) {
    float lo = BL_INF, hi = -BL_INF;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < BT; i += (long long)gridDim.x * blockDim.x) {
        int16_t ni = n[i];
        for (int s = 0; s < Sn; s++) {
            float q = bl_qraw(w[i * Sn + s], ni);
            lo = fminf(lo, q);
            hi = fmaxf(hi, q);
        }
    }
    int klo = bl_f2ord(lo), khi = bl_f2ord(hi);
    klo = __reduce_min_sync(0xffffffffu, klo);
    khi = __reduce_max_sync(0xffffffffu, khi);
    __shared__ int slo[8], shi[8];
    int wid = threadIdx.x >> 5;
    if (bl_lane() == 0) { slo[wid] = klo; shi[wid] = khi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); k++) { klo = min(klo, slo[k]); khi = max(khi, shi[k]); }
        atomicMin(&qrange[0], klo);
        atomicMax(&qrange[1], khi);
    }
}

__global__ void __launch_bounds__(256) transition_q_kernel(const bl_half *__restrict__ w, const int16_t *__restrict__ n,
                                                           const float *__restrict__ qrange, bl_half *__restrict__ q,
                                                           long long BT, int Sn) {
    bl_qnorm qn(qrange);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < BT; i += (long long)gridDim.x * blockDim.x) {
        int16_t ni = n[i];
        for (int s = 0; s < Sn; s++) q[i * Sn + s] = bl_f2h(qn(w[i * Sn + s], ni));
    }
}

int launch_qrange(const bl_half *w, const int16_t *n, float *qrange, int B, int T, int Sn, cudaStream_t st) {
    long long BT = (long long)B * T;
    qrange_init_kernel<<<1, 1, 0, st>>>(reinterpret_cast<int *>(qrange));
    long long g = (BT + 255) / 256;
    int grid = (int)(g < BL_NUM_SMS * 8 ? g : BL_NUM_SMS * 8);
    qrange_kernel<<<grid, 256, 0, st>>>(w, n, reinterpret_cast<int *>(qrange), BT, Sn);
    return (int)cudaGetLastError();
}

// ---- policy on the reference layout ------------------------------------------------------------------------
struct RefTree {
    const bl_half *logits, *w, *c_puct;
    const int16_t *n, *seats, *children;
    const uint8_t *terminal;
    const float *exp_lut;
    int B, T, A, Sn;
};

constexpr int DNT = 64;   // lanes (= envs) per CTA in descend/root

// policy() of cuda.cu:70-99 for (env b, node t) into the lane's shared-memory columns; returns alpha.
__device__ __forceinline__ float ref_policy(const RefTree &m, const bl_qnorm &qn, int b, int t, float *top, float *q,
                                            int stride, unsigned *n_children, int *iters) {
    const int A = m.A;
    const size_t node = (size_t)b * m.T;
    const size_t row = (node + t) * A;
    const int seat = m.seats[node + t];
    int N = 0;
    unsigned nc = 0;
    for (int a = 0; a < A; a++) {
        int child = m.children[row + a];
        top[a * stride] = m.exp_lut[m.logits[row + a]];
        if (child > -1) {
            q[a * stride] = qn(m.w[(node + child) * m.Sn + seat], m.n[node + child]);
            N += m.n[node + child];
            nc++;
        } else {
            q[a * stride] = 0.f;
            N += 1;
        }
    }
    float lambda = bl_lambda(bl_h2f(m.c_puct[b]), N, A);
    for (int a = 0; a < A; a++) top[a * stride] = __fmul_rn(lambda, top[a * stride]);
    *n_children = nc;
    return bl_newton(top, q, stride, A, iters);
}

__global__ void __launch_bounds__(DNT) descend_kernel(RefTree m, const float *__restrict__ qrange,
                                                      const bl_half *__restrict__ rands, int16_t *__restrict__ parents,
                                                      int16_t *__restrict__ actions, uint64_t *counters) {
    extern __shared__ float sm[];
    float *top = sm + threadIdx.x, *q = sm + (size_t)m.A * DNT + threadIdx.x;
    const int b = blockIdx.x * DNT + threadIdx.x;
    unsigned c_evals = 0, c_children = 0, c_iters = 0, c_desc = 0;
    if (b < m.B) {
        bl_qnorm qn(qrange);
        int t = 0, parent = 0, action = -1;
        c_desc = 1;
        while (true) {
            if (t == -1) break;
            if (m.terminal[(size_t)b * m.T + t]) break;
            unsigned nc; int it;
            float alpha = ref_policy(m, qn, b, t, top, q, DNT, &nc, &it);
            c_evals++; c_children += nc; c_iters += it;
            float r = bl_h2f(rands[(size_t)b * m.T + t]);
            action = bl_sample(top, q, DNT, m.A, alpha, r);
            parent = t;
            if (action < 0) break;   // no positive-probability action: the reference indexes children[-1] here (UB)
            t = m.children[((size_t)b * m.T + t) * m.A + action];
        }
        parents[b] = (int16_t)parent;
        actions[b] = (int16_t)action;
    }
    bl_count(counters, 0, c_evals);
    bl_count(counters, 1, c_children);
    bl_count(counters, 2, c_iters);
    bl_count(counters, 3, c_desc);
}

__global__ void __launch_bounds__(DNT) root_kernel(RefTree m, const float *__restrict__ qrange, bl_half *__restrict__ probs) {
    extern __shared__ float sm[];
    float *top = sm + threadIdx.x, *q = sm + (size_t)m.A * DNT + threadIdx.x;
    const int b = blockIdx.x * DNT + threadIdx.x;
    if (b >= m.B) return;
    bl_qnorm qn(qrange);
    unsigned nc; int it;
    float alpha = ref_policy(m, qn, b, 0, top, q, DNT, &nc, &it);
    for (int a = 0; a < m.A; a++)
        probs[(size_t)b * m.A + a] = bl_f2h(bl_prob(top[a * DNT], q[a * DNT], alpha));
}

// backup_kernel of cuda.cu:205-236.  MAXSN bounds the per-lane value registers.
constexpr int MAXSN = 4;
__global__ void __launch_bounds__(128) backup_kernel(const bl_half *__restrict__ v, bl_half *__restrict__ w,
                                                     int16_t *__restrict__ n, const bl_half *__restrict__ rewards,
                                                     const int16_t *__restrict__ parents, const uint8_t *__restrict__ terminal,
                                                     const int16_t *__restrict__ leaves, int B, int T, int Sn) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float val[MAXSN];
    int cur = leaves[b];
    const size_t base = (size_t)b * T;
#pragma unroll
    for (int s = 0; s < MAXSN; s++) val[s] = (s < Sn && cur >= 0) ? bl_h2f(v[(base + cur) * Sn + s]) : 0.f;
    while (cur != -1) {
        const size_t node = base + cur;
        const bool term = terminal[node];
        int16_t nn = n[node];
#pragma unroll
        for (int s = 0; s < MAXSN; s++) {
            if (s < Sn) {
                if (term) val[s] = 0.f;
                val[s] = __fadd_rn(val[s], bl_h2f(rewards[node * Sn + s]));
                nn = (int16_t)(nn + 1);                                   // quirk: +1 per seat (cuda.cu:228)
                // c10::Half += : half(float(w) + float(half(v)))
                w[node * Sn + s] = bl_f2h(__fadd_rn(bl_h2f(w[node * Sn + s]), bl_h2f(bl_f2h(val[s]))));
            }
        }
        n[node] = nn;
        cur = parents[node];
    }
}

size_t policy_smem(int A) { return (size_t)2 * A * DNT * sizeof(float); }

template <typename K>
int ensure_smem(K kern, size_t bytes) {
    if (bytes > 227 * 1024) return -2;
    if (bytes > 48 * 1024) return (int)cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    return 0;
}

}  // namespace

extern "C" int bl_mcts_transition_q(const bl_half *w, const int16_t *n, bl_half *q, float *qrange, int B, int T, int Sn,
                                    bl_stream stream) {
    if (B <= 0 || T <= 0 || Sn <= 0) return B == 0 ? 0 : -1;
    cudaStream_t st = bl_cu(stream);
    int e = launch_qrange(w, n, qrange, B, T, Sn, st);
    if (e) return e;
    long long BT = (long long)B * T;
    long long g = (BT + 255) / 256;
    transition_q_kernel<<<(int)(g < BL_NUM_SMS * 8 ? g : BL_NUM_SMS * 8), 256, 0, st>>>(w, n, qrange, q, BT, Sn);
    BL_LAUNCH_CHECK();
}

extern "C" int bl_mcts_descend(const bl_half *logits, const bl_half *w, const int16_t *n, const bl_half *c_puct,
                               const int16_t *seats, const uint8_t *terminal, const int16_t *children,
                               const bl_half *rands, const float *exp_lut, float *qrange, int16_t *parents,
                               int16_t *actions, uint64_t *counters, int B, int T, int A, int Sn, bl_stream stream) {
    if (B < 0 || T <= 0 || A <= 0 || Sn <= 0 || Sn > MAXSN) return -1;
    if (B == 0) return 0;
    cudaStream_t st = bl_cu(stream);
    int e = launch_qrange(w, n, qrange, B, T, Sn, st);
    if (e) return e;
    e = ensure_smem(descend_kernel, policy_smem(A));
    if (e) return e;
    RefTree m{logits, w, c_puct, n, seats, children, terminal, exp_lut, B, T, A, Sn};
    descend_kernel<<<(B + DNT - 1) / DNT, DNT, policy_smem(A), st>>>(m, qrange, rands, parents, actions, counters);
    BL_LAUNCH_CHECK();
}

extern "C" int bl_mcts_root(const bl_half *logits, const bl_half *w, const int16_t *n, const bl_half *c_puct,
                            const int16_t *seats, const uint8_t *terminal, const int16_t *children,
                            const float *exp_lut, float *qrange, bl_half *probs, int B, int T, int A, int Sn,
                            bl_stream stream) {
    if (B < 0 || T <= 0 || A <= 0 || Sn <= 0 || Sn > MAXSN) return -1;
    if (B == 0) return 0;
    cudaStream_t st = bl_cu(stream);
    int e = launch_qrange(w, n, qrange, B, T, Sn, st);
    if (e) return e;
    e = ensure_smem(root_kernel, policy_smem(A));
    if (e) return e;
    RefTree m{logits, w, c_puct, n, seats, children, terminal, exp_lut, B, T, A, Sn};
    root_kernel<<<(B + DNT - 1) / DNT, DNT, policy_smem(A), st>>>(m, qrange, probs);
    BL_LAUNCH_CHECK();
}

extern "C" int bl_mcts_backup(const bl_half *v, bl_half *w, int16_t *n, const bl_half *rewards, const int16_t *parents,
                              const uint8_t *terminal, const int16_t *leaves, int B, int T, int Sn, bl_stream stream) {
    if (B < 0 || T <= 0 || Sn <= 0 || Sn > MAXSN) return -1;
    if (B == 0) return 0;
    backup_kernel<<<(B + 127) / 128, 128, 0, bl_cu(stream)>>>(v, w, n, rewards, parents, terminal, leaves, B, T, Sn);
    BL_LAUNCH_CHECK();
}
