// Policy/value network forward for sm_100a — replaces FCModel.forward (boardlaw/networks.py:37-41):
//   x0 = W_in . obs + b_in                    TensorIntake   boardlaw/heads.py:41-52
//   x_{k+1} = x_k + alpha_k (W_k relu(x_k) + b_k)   ReZeroResidual boardlaw/networks.py:10-18
//   logits = log_softmax(where(valid, W_p x + b_p, -inf))      MaskedOutput boardlaw/heads.py:93-104
//   v[seat] = tanh(w_v x + b_v), v[1-seat] = -that              ValueOutput  boardlaw/heads.py:128-142
// The observation is never materialised: the first layer's operand is generated from the board bytes.
//
// This file holds the fp32 CUDA-core path (exact fp32 products, fp32 accumulation) used for shapes the
// tensor-core path (net_tc.cu) does not cover and as its on-device cross-check.
#include "common.cuh"
#include "hex_core.cuh"

namespace {

constexpr int TM = 64, TN = 64, TK = 16, NTHREADS = 256;

// Generic row-tile GEMM with fused prologue/epilogue:
//   OBS_IN : A operand is the one-hot observation generated from (board, seats), K = 2*S*S
//   else   : A operand is X (M,K) fp32, relu applied on load when RELU_IN
//   out    : Y = (RES ? Xres + alpha * (acc + bias) : acc + bias)
template <bool OBS_IN, bool RELU_IN, bool RES>
__global__ void __launch_bounds__(NTHREADS) fc_layer_kernel(
    const float *__restrict__ X, const uint8_t *__restrict__ board, const int32_t *__restrict__ seats, int S,
    const float *__restrict__ Wt /* (N,K) */, const float *__restrict__ bias, const float *__restrict__ alpha_p,
    const float *__restrict__ Xres, float *__restrict__ Y, int M, int N, int K) {
    __shared__ __align__(16) float As[TK][TM + 4];
    __shared__ __align__(16) float Bs[TK][TN + 4];
    const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;        // 16 x 16 threads, 4 x 4 outputs each
    float acc[4][4] = {};
    const int A = S * S;

    for (int k0 = 0; k0 < K; k0 += TK) {
        // stage A tile: TM x TK, 4 elements per thread
#pragma unroll
        for (int i = 0; i < 4; i++) {
            int e = tid + i * NTHREADS;             // 0..1023
            int kk = e & (TK - 1), mm = e >> 4;
            int m = m0 + mm, k = k0 + kk;
            float val = 0.f;
            if (m < M && k < K) {
                if (OBS_IN) {
                    int cell = k >> 1, ch = k & 1;
                    int seat = seats[m];
                    int r = cell / S, c = cell - r * S;
                    uint8_t cv = board[(size_t)m * A + (seat ? c * S + r : cell)];
                    bool black = cv == BL_BLACK || cv == BL_TOP || cv == BL_BOT;
                    bool white = cv == BL_WHITE || cv == BL_LEFT || cv == BL_RIGHT;
                    bool own = seat ? white : black, opp = seat ? black : white;
                    val = (ch == 0 ? own : opp) ? 1.f : 0.f;
                } else {
                    val = X[(size_t)m * K + k];
                    if (RELU_IN) val = fmaxf(val, 0.f);
                }
            }
            As[kk][mm] = val;
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            int e = tid + i * NTHREADS;
            int kk = e & (TK - 1), nn = e >> 4;
            int n = n0 + nn, k = k0 + kk;
            Bs[kk][nn] = (n < N && k < K) ? Wt[(size_t)n * K + k] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; kk++) {
            float4 a = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
            float4 b = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
            float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    const float alpha = RES ? *alpha_p : 0.f;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float y = acc[i][j] + bias[n];
            if (RES) y = Xres[(size_t)m * N + n] + alpha * y;
            Y[(size_t)m * N + n] = y;
        }
    }
}

// Heads: one warp per env.  logits_raw (M,A) already holds W_p x + b_p.
__global__ void __launch_bounds__(256) heads_kernel(
    const float *__restrict__ X /* (M,W) neck */, const float *__restrict__ raw /* (M,A) */,
    const uint8_t *__restrict__ board, const int32_t *__restrict__ seats, int S,
    const float *__restrict__ w_val, const float *__restrict__ b_val,
    float *__restrict__ logits, float *__restrict__ v, int M, int W) {
    const int A = S * S;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = bl_lane();
    if (warp >= M) return;
    const int m = warp, seat = seats[m];
    // masked log-softmax (heads.py:101-104)
    float mx = -BL_INF_F;
    for (int a = lane; a < A; a += 32) {
        int r = a / S, c = a - r * S;
        bool valid = board[(size_t)m * A + (seat ? c * S + r : a)] == BL_EMPTY;
        if (valid) mx = fmaxf(mx, raw[(size_t)m * A + a]);
    }
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int a = lane; a < A; a += 32) {
        int r = a / S, c = a - r * S;
        bool valid = board[(size_t)m * A + (seat ? c * S + r : a)] == BL_EMPTY;
        if (valid) sum += expf(raw[(size_t)m * A + a] - mx);
    }
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float lse = logf(sum);
    for (int a = lane; a < A; a += 32) {
        int r = a / S, c = a - r * S;
        bool valid = board[(size_t)m * A + (seat ? c * S + r : a)] == BL_EMPTY;
        logits[(size_t)m * A + a] = valid ? (raw[(size_t)m * A + a] - mx) - lse : -BL_INF_F;
    }
    // value head (heads.py:128-142)
    float dot = 0.f;
    for (int k = lane; k < W; k += 32) dot = fmaf(X[(size_t)m * W + k], w_val[k], dot);
    for (int o = 16; o; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if (lane == 0) {
        float t = tanhf(dot + b_val[0]);
        v[(size_t)m * 2 + seat] = t;
        v[(size_t)m * 2 + (1 - seat)] = -t;
    }
}

}  // namespace

// net_tc.cu
int bl_fc_forward_tc(const bl_fc_params *p, const uint8_t *board, long long board_pitch, const int32_t *seats, float *logits,
                     float *v, int B, cudaStream_t st);
bool bl_fc_tc_supported(const bl_fc_params *p);
// net_tc_wide.cu
int bl_fc_forward_wide(const bl_fc_params *p, const uint8_t *board, long long board_pitch, const int32_t *seats, float *logits,
                       float *v, void *scratch, int B, cudaStream_t st);
bool bl_fc_wide_supported(const bl_fc_params *p);
int64_t bl_fc_wide_scratch_bytes(const bl_fc_params *p, int B);

extern "C" int bl_fc_uses_tensor_cores(const bl_fc_params *p) { return (bl_fc_tc_supported(p) || bl_fc_wide_supported(p)) ? 1 : 0; }

extern "C" int64_t bl_fc_scratch_bytes(const bl_fc_params *p, int B) {
    if (bl_fc_tc_supported(p)) return 0;
    if (bl_fc_wide_supported(p)) return bl_fc_wide_scratch_bytes(p, B);
    return (int64_t)sizeof(float) * B * (2 * (int64_t)p->W + (int64_t)p->S * p->S);
}

extern "C" int bl_fc_forward(const bl_fc_params *p, const uint8_t *board, const int32_t *seats, float *logits,
                             float *v, void *scratch_, int B, bl_stream stream) {
    float *scratch = reinterpret_cast<float *>(scratch_);
    if (B < 0 || p->S < 1 || p->W < 1 || p->D < 0) return -1;
    if (B == 0) return 0;
    cudaStream_t st = bl_cu(stream);
    const int S = p->S, A = S * S, W = p->W;
    if (bl_fc_tc_supported(p)) return bl_fc_forward_tc(p, board, (long long)A, seats, logits, v, B, st);
    if (bl_fc_wide_supported(p)) return bl_fc_forward_wide(p, board, (long long)A, seats, logits, v, scratch_, B, st);
    float *x0 = scratch, *x1 = scratch + (size_t)B * W, *raw = scratch + (size_t)2 * B * W;
    dim3 grid((B + TM - 1) / TM, (W + TN - 1) / TN);
    fc_layer_kernel<true, false, false><<<grid, NTHREADS, 0, st>>>(nullptr, board, seats, S, p->w_in, p->b_in, nullptr,
                                                                    nullptr, x0, B, W, 2 * A);
    for (int d = 0; d < p->D; d++) {
        fc_layer_kernel<false, true, true><<<grid, NTHREADS, 0, st>>>(x0, nullptr, nullptr, S, p->w_res + (size_t)d * W * W,
                                                                      p->b_res + (size_t)d * W, p->alpha + d, x0, x1, B, W, W);
        float *t = x0; x0 = x1; x1 = t;
    }
    dim3 gridp((B + TM - 1) / TM, (A + TN - 1) / TN);
    fc_layer_kernel<false, false, false><<<gridp, NTHREADS, 0, st>>>(x0, nullptr, nullptr, S, p->w_pol, p->b_pol, nullptr,
                                                                     nullptr, raw, B, A, W);
    heads_kernel<<<(int)(((long long)B * 32 + 255) / 256), 256, 0, st>>>(x0, raw, board, seats, S, p->w_val, p->b_val,
                                                                         logits, v, B, W);
    BL_LAUNCH_CHECK();
}
