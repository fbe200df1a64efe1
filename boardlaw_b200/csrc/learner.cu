// Learner-side kernels for sm_100a: what consumes the trajectories the self-play path produces.
//
//   bl_reward_to_go        learning.reward_to_go + present_value (boardlaw/learning.py:57-76): a reverse scan over the
//                          time axis of a chunk, one thread per (env, seat) column; HBM-bound, 13 bytes in + 4 (or 2) out
//                          per (t, env, seat)
//   bl_policy_value_loss   the loss of main.optimize (boardlaw/main.py:86-101) and its gradient with respect to the network's
//                          pre-softmax scores and pre-tanh value in ONE pass over the batch (one warp per sample): the
//                          reference launches ~20 elementwise / reduction kernels forward and as many backward through autograd
//   bl_adam_step           torch.optim.Adam's update (no weight decay, no amsgrad) over one flat parameter buffer
//
// The dense contractions of the learner's forward / backward are plain GEMMs and go to cuBLAS from the host side
// (boardlaw_b200/learner.py); they are ~3x one network launch of the self-play path per 64 moves.
#include "common.cuh"

namespace {

// result[T-1] = fallback[T-1]; result[t] = terminal[t] ? fallback[t] : reward[t] + gamma * result[t+1], with
// fallback = terminal ? reward : value (learning.py:72-76: `fallback = value; fallback[terminal] = reward[terminal]`).
// fp32, the product rounded before the sum, as torch evaluates `deltas[t] + alpha*result[t+1]`.
template <bool HALF_OUT>
__global__ void __launch_bounds__(256) reward_to_go_kernel(const float *__restrict__ reward, const float *__restrict__ value,
                                                           const uint8_t *__restrict__ terminal, void *__restrict__ out,
                                                           int T, int B, int Sn, float gamma) {
    const int col = blockIdx.x * 256 + threadIdx.x;               // (env, seat) column
    const int C = B * Sn;
    if (col >= C) return;
    const int b = col / Sn;
    float res = 0.f;
    for (int t = T - 1; t >= 0; t--) {
        const size_t i = (size_t)t * C + col;
        const bool term = terminal[(size_t)t * B + b] != 0;
        const float r = reward[i];
        if (t == T - 1) res = term ? r : value[i];
        else res = term ? r : __fadd_rn(r, __fmul_rn(gamma, res));
        if (HALF_OUT) reinterpret_cast<__half *>(out)[i] = __float2half_rn(res);
        else reinterpret_cast<float *>(out)[i] = res;
    }
}

// One warp per sample.  logp (N,A) f32: the network's masked log-softmax output (-inf on illegal moves); v (N,Sn=2) f32;
// target_logits (N,A) half: the search policy's log-probabilities; target_v (N,2) half: reward-to-go; seats (N,) i32.
//   policy_loss = -mean_n sum_a exp(l0_a) * l_a     with -inf entries of either replaced by 0 (main.py:90-94)
//   value_loss  = mean_{n,s} (target_s - v_s)^2     (main.py:96-97)
// Gradients of (policy_loss + value_loss):
//   g_a = dL/dl_a = -exp(l0_a)/N on legal moves; through the log-softmax: dscores_a = g_a - exp(l_a) * sum_b g_b
//   dL/dv_s = -2 (target_s - v_s)/(2N); v[seat] = t, v[1-seat] = -t, t = tanh(z): dz = (dL/dv[seat] - dL/dv[1-seat]) (1 - t^2)
// sums[0] += sum exp(l0)*l, sums[1] += sum (target - v)^2 (the caller divides): one atomic pair per CTA.
__global__ void __launch_bounds__(256) policy_value_loss_kernel(const float *__restrict__ logp, const float *__restrict__ v,
                                                                const __half *__restrict__ target_logits,
                                                                const __half *__restrict__ target_v, const int32_t *__restrict__ seats,
                                                                float *__restrict__ dscores, float *__restrict__ dz,
                                                                float *__restrict__ sums, int N, int A) {
    __shared__ float part[2][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = blockIdx.x * 8 + warp;
    float ps = 0.f, vs = 0.f;
    if (n < N) {
        const float invN = 1.f / (float)N;
        const float *l = logp + (size_t)n * A;
        const __half *l0 = target_logits + (size_t)n * A;
        float gsum = 0.f;
        for (int a = lane; a < A; a += 32) {
            const float la = l[a], l0a = __half2float(l0[a]);
            const float lz = la > -BL_INF_F ? la : 0.f, l0z = l0a > -BL_INF_F ? l0a : 0.f;
            const float p0 = __expf(l0z) ;
            ps += p0 * lz;
            gsum += la > -BL_INF_F ? -p0 * invN : 0.f;
        }
        for (int o = 16; o; o >>= 1) { ps += __shfl_xor_sync(0xffffffffu, ps, o); gsum += __shfl_xor_sync(0xffffffffu, gsum, o); }
        float *d = dscores + (size_t)n * A;
        for (int a = lane; a < A; a += 32) {
            const float la = l[a], l0a = __half2float(l0[a]);
            const float l0z = l0a > -BL_INF_F ? l0a : 0.f;
            d[a] = la > -BL_INF_F ? (-__expf(l0z) * invN - __expf(la) * gsum) : 0.f;
        }
        if (lane == 0) {
            const int s = seats[n] & 1;
            const float v0 = v[2 * n], v1 = v[2 * n + 1];
            const float e0 = __half2float(target_v[2 * n]) - v0, e1 = __half2float(target_v[2 * n + 1]) - v1;
            vs = e0 * e0 + e1 * e1;
            const float g0 = -e0 * invN, g1 = -e1 * invN;          // -2 e / (2N)
            const float t = s ? v1 : v0;
            dz[n] = (s ? g1 - g0 : g0 - g1) * (1.f - t * t);
        }
    }
    if (lane == 0) { part[0][warp] = ps; part[1][warp] = vs; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int w = 0; w < 8; w++) { a += part[0][w]; b += part[1][w]; }
        atomicAdd(sums, a);
        atomicAdd(sums + 1, b);
    }
}

// torch.optim.Adam (weight_decay 0, amsgrad off): m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
// p -= (lr / (1 - b1^step)) * m / (sqrt(v) / sqrt(1 - b2^step) + eps)
__global__ void __launch_bounds__(256) adam_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                                                   float *__restrict__ v, long long n, float lr, float b1, float b2, float eps,
                                                   float bias1, float sqrt_bias2) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= (lr / bias1) * mi / (sqrtf(vi) / sqrt_bias2 + eps);
}

}  // namespace

extern "C" int bl_reward_to_go(const float *reward, const float *value, const uint8_t *terminal, void *out, int out_is_half,
                               int T, int B, int Sn, float gamma, bl_stream stream) {
    if (T <= 0 || B <= 0 || Sn <= 0) return T < 0 || B < 0 || Sn < 0 ? -1 : 0;
    const int C = B * Sn;
    if (out_is_half) reward_to_go_kernel<true><<<(C + 255) / 256, 256, 0, bl_cu(stream)>>>(reward, value, terminal, out, T, B, Sn, gamma);
    else reward_to_go_kernel<false><<<(C + 255) / 256, 256, 0, bl_cu(stream)>>>(reward, value, terminal, out, T, B, Sn, gamma);
    BL_LAUNCH_CHECK();
}

extern "C" int bl_policy_value_loss(const float *logp, const float *v, const bl_half *target_logits, const bl_half *target_v,
                                    const int32_t *seats, float *dscores, float *dz, float *sums, int N, int A, bl_stream stream) {
    if (N <= 0) return N < 0 ? -1 : 0;
    if (A <= 0) return -1;
    policy_value_loss_kernel<<<(N + 7) / 8, 256, 0, bl_cu(stream)>>>(logp, v, reinterpret_cast<const __half *>(target_logits),
                                                                     reinterpret_cast<const __half *>(target_v), seats, dscores, dz, sums, N, A);
    BL_LAUNCH_CHECK();
}

extern "C" int bl_adam_step(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, int64_t n, float lr, float beta1,
                            float beta2, float eps, int step, bl_stream stream) {
    if (n <= 0) return n < 0 ? -1 : 0;
    if (step < 1) return -1;
    const float bias1 = 1.f - powf(beta1, (float)step), bias2 = 1.f - powf(beta2, (float)step);
    adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, bl_cu(stream)>>>(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, bias1,
                                                                        sqrtf(bias2));
    BL_LAUNCH_CHECK();
}
