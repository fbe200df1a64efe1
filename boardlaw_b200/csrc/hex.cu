// Hex environment kernels for sm_100a.
//
// Replaces hexcuda.step / hexcuda.observe (boardlaw/hex/cpp/cuda.cu:76-217) and the host-side glue of
// Hex.step / Hex.valid (boardlaw/hex/__init__.py:148-195).
//
// Data movement: a CTA owns NT consecutive envs.  Their boards are one contiguous NT*A-byte span of
// HBM, which the CTA streams in with fully coalesced loads and scatters into a lane-major shared
// memory tile (cell c of env e at [c*PITCH + e], PITCH = NT+4 keeps consecutive cells on distinct banks).
// Each lane then plays its env's move out of shared memory (data-dependent flood fill, no global
// traffic), and the tile is streamed back coalesced.  Algorithmic traffic: 2A + 8 + 8 bytes per env.
#include "hex_core.cuh"

namespace {

constexpr int NT = 128;          // envs (= threads) per CTA
constexpr int PITCH = NT + 4;    // shared-memory pitch in elements

// ---- cooperative tile movement -------------------------------------------------------------
// 16 bytes per thread per step (the CTA's span starts at a multiple of NT*A bytes, so it is 16-byte aligned whenever the tensor
// is), one division per 16 bytes to locate (env, cell), then incremental; unaligned tensors and the tail move byte by byte.
__device__ __forceinline__ void tile_load(uint8_t *sb, const uint8_t *__restrict__ g, int n_bytes, int A) {
    // g points at the first env of the CTA; n_bytes = (#envs in this CTA) * A
    const bool vec = (reinterpret_cast<uintptr_t>(g) & 15) == 0;
    const int n16 = vec ? n_bytes >> 4 : 0;
    for (int i = threadIdx.x; i < n16; i += NT) {
        const uint4 v = reinterpret_cast<const uint4 *>(g)[i];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        int e = (i * 16) / A, c = i * 16 - e * A;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            sb[c * PITCH + e] = (uint8_t)(w[k >> 2] >> (8 * (k & 3)));
            if (++c == A) { c = 0; e++; }
        }
    }
    for (int i = n16 * 16 + threadIdx.x; i < n_bytes; i += NT) {
        int e = i / A, c = i - e * A;
        sb[c * PITCH + e] = g[i];
    }
}
__device__ __forceinline__ void tile_store(const uint8_t *sb, uint8_t *__restrict__ g, int n_bytes, int A) {
    const bool vec = (reinterpret_cast<uintptr_t>(g) & 15) == 0;
    const int n16 = vec ? n_bytes >> 4 : 0;
    for (int i = threadIdx.x; i < n16; i += NT) {
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        int e = (i * 16) / A, c = i * 16 - e * A;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            w[k >> 2] |= (uint32_t)sb[c * PITCH + e] << (8 * (k & 3));
            if (++c == A) { c = 0; e++; }
        }
        reinterpret_cast<uint4 *>(g)[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    for (int i = n16 * 16 + threadIdx.x; i < n_bytes; i += NT) {
        int e = i / A, c = i - e * A;
        g[i] = sb[c * PITCH + e];
    }
}

// mode bits
constexpr int F_TRANSITION = 1;   // out-of-place + reset + seat flip + rule check (Hex.step)
constexpr int F_RANDOM = 2;       // the action is drawn here: the k-th legal move, k = floor(u * n_legal), u a caller-supplied uniform

template <typename StkT, int MODE>
__global__ void __launch_bounds__(NT) hex_step_kernel(
    const uint8_t *__restrict__ board_in, uint8_t *__restrict__ board_out,
    const int32_t *__restrict__ seats, const void *__restrict__ actions_, float *__restrict__ rewards,
    int32_t *__restrict__ new_seats, uint8_t *__restrict__ terminal, int32_t *__restrict__ error_word,
    int64_t *__restrict__ actions_out, int reset, int B, int S) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int A = S * S;
    uint8_t *sb = smem;                                           // A * PITCH bytes
    StkT *stk = reinterpret_cast<StkT *>(smem + ((A * PITCH + 15) & ~15));

    const int b0 = blockIdx.x * NT;
    const int nb = min(NT, B - b0);
    tile_load(sb, board_in + (size_t)b0 * A, nb * A, A);
    __syncthreads();

    const int tid = threadIdx.x;
    const int b = b0 + tid;
    if (tid < nb) {
        const int seat = seats[b];
        long long action;
        if (MODE & F_RANDOM) {
            // uniformly random legal move (learning.mix, boardlaw/learning.py:6-10 draws Categorical(probs=valid)): count the empty
            // cells, take the k-th in the mover's frame order (white sees the transpose, hex/__init__.py:154-159)
            int n_empty = 0;
            for (int c = 0; c < A; c++) n_empty += sb[c * PITCH + tid] == BL_EMPTY;
            const float u = reinterpret_cast<const float *>(actions_)[b];
            int k = (int)__fmul_rn(u, (float)n_empty);
            k = k < n_empty - 1 ? k : n_empty - 1;
            k = k > 0 ? k : 0;
            action = -1;
            int r = 0, c = 0;
            for (int a = 0; a < A && action < 0; a++) {
                if (sb[(seat ? c * S + r : a) * PITCH + tid] == BL_EMPTY && k-- == 0) action = a;
                if (++c == S) { c = 0; r++; }
            }
            actions_out[b] = action;
        } else if (MODE & F_TRANSITION) action = reinterpret_cast<const int64_t *>(actions_)[b];
        else action = reinterpret_cast<const int32_t *>(actions_)[b];

        int win = 0;
        bool ok = true;
        if (MODE & F_TRANSITION) {
            // Hex.step's asserts (hex/__init__.py:174,179) become bits in a device-side error word
            ok = action >= 0 && action < A;
            if (ok) {
                int a = (int)action;
                int cell = seat ? (a % S) * S + a / S : a;
                ok = sb[cell * PITCH + tid] == BL_EMPTY;
                if (!ok) atomicOr(error_word, 2);
            } else {
                atomicOr(error_word, 1);
            }
        }
        if (ok) win = bl_hex_place<StkT>(sb + tid, stk + tid, PITCH, S, seat, (int)action);

        float r0 = win == 1 ? 1.f : (win == 2 ? -1.f : 0.f), r1 = win == 1 ? -1.f : (win == 2 ? 1.f : 0.f);
        reinterpret_cast<float2 *>(rewards)[b] = make_float2(r0, r1);
        if (MODE & F_TRANSITION) {
            bool term = reset && win != 0;                        // hex/__init__.py:183
            terminal[b] = term ? 1 : 0;
            new_seats[b] = term ? 0 : 1 - seat;                   // hex/__init__.py:187-188
            if (term)
                for (int c = 0; c < A; c++) sb[c * PITCH + tid] = 0;   // hex/__init__.py:185
        }
    }
    __syncthreads();
    tile_store(sb, board_out + (size_t)b0 * A, nb * A, A);
}

// One thread per observation cell (b, i, j): writes the (own, opp) pair as one float2.
__global__ void __launch_bounds__(256) hex_observe_kernel(
    const uint8_t *__restrict__ board, const int32_t *__restrict__ seats, float2 *__restrict__ obs,
    long long n_cells, int S) {
    const int A = S * S;
    for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < n_cells;
         o += (long long)gridDim.x * blockDim.x) {
        long long b = o / A;
        int c = (int)(o - b * A);
        int i = c / S, j = c - i * S;
        int seat = seats[b];
        uint8_t v = board[b * A + (seat ? j * S + i : c)];       // cuda.cu:179-194: white sees the transpose
        bool black = v == BL_BLACK || v == BL_TOP || v == BL_BOT;
        bool white = v == BL_WHITE || v == BL_LEFT || v == BL_RIGHT;
        bool own = seat ? white : black, opp = seat ? black : white;
        obs[o] = make_float2(own ? 1.f : 0.f, opp ? 1.f : 0.f);
    }
}

__global__ void __launch_bounds__(256) hex_valid_kernel(
    const uint8_t *__restrict__ board, const int32_t *__restrict__ seats, uint8_t *__restrict__ valid,
    long long n_cells, int S) {
    const int A = S * S;
    for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < n_cells;
         o += (long long)gridDim.x * blockDim.x) {
        long long b = o / A;
        int c = (int)(o - b * A);
        int i = c / S, j = c - i * S;
        valid[o] = board[b * A + (seats[b] ? j * S + i : c)] == BL_EMPTY;
    }
}

template <typename StkT, int MODE>
int launch_step(const uint8_t *bin, uint8_t *bout, const int32_t *seats, const void *actions, float *rewards,
                int32_t *new_seats, uint8_t *terminal, int32_t *error_word, int64_t *actions_out, int reset, int B, int S,
                cudaStream_t st) {
    const int A = S * S;
    size_t smem = ((A * PITCH + 15) & ~15) + (size_t)A * PITCH * sizeof(StkT);
    auto kern = hex_step_kernel<StkT, MODE>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    kern<<<(B + NT - 1) / NT, NT, smem, st>>>(bin, bout, seats, actions, rewards, new_seats, terminal,
                                              error_word, actions_out, reset, B, S);
    BL_LAUNCH_CHECK();
}

int grid_for(long long n, int block) {
    long long g = (n + block - 1) / block;
    long long cap = (long long)BL_NUM_SMS * 16;
    return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace

extern "C" int bl_hex_step(uint8_t *board, const int32_t *seats, const int32_t *actions, float *rewards,
                           int B, int S, bl_stream stream) {
    if (B < 0 || S < 1 || S > 19) return -1;
    if (B == 0) return 0;
    if (S <= 15)
        return launch_step<uint8_t, 0>(board, board, seats, actions, rewards, nullptr, nullptr, nullptr, nullptr, 0, B, S, bl_cu(stream));
    return launch_step<uint16_t, 0>(board, board, seats, actions, rewards, nullptr, nullptr, nullptr, nullptr, 0, B, S, bl_cu(stream));
}

extern "C" int bl_hex_transition(const uint8_t *board, const int32_t *seats, const int64_t *actions,
                                 uint8_t *new_board, int32_t *new_seats, float *rewards, uint8_t *terminal,
                                 int32_t *error_word, int reset, int B, int S, bl_stream stream) {
    if (B < 0 || S < 1 || S > 19) return -1;
    if (B == 0) return 0;
    if (S <= 15)
        return launch_step<uint8_t, F_TRANSITION>(board, new_board, seats, actions, rewards, new_seats, terminal,
                                                  error_word, nullptr, reset, B, S, bl_cu(stream));
    return launch_step<uint16_t, F_TRANSITION>(board, new_board, seats, actions, rewards, new_seats, terminal,
                                               error_word, nullptr, reset, B, S, bl_cu(stream));
}

extern "C" int bl_hex_random_transition(const uint8_t *board, const int32_t *seats, const float *uniforms,
                                        uint8_t *new_board, int32_t *new_seats, int64_t *actions, float *rewards,
                                        uint8_t *terminal, int32_t *error_word, int reset, int B, int S, bl_stream stream) {
    if (B < 0 || S < 1 || S > 19) return -1;
    if (B == 0) return 0;
    if (S <= 15)
        return launch_step<uint8_t, F_TRANSITION | F_RANDOM>(board, new_board, seats, uniforms, rewards, new_seats, terminal,
                                                             error_word, actions, reset, B, S, bl_cu(stream));
    return launch_step<uint16_t, F_TRANSITION | F_RANDOM>(board, new_board, seats, uniforms, rewards, new_seats, terminal,
                                                          error_word, actions, reset, B, S, bl_cu(stream));
}

extern "C" int bl_hex_observe(const uint8_t *board, const int32_t *seats, float *obs, int B, int S,
                              bl_stream stream) {
    if (B < 0 || S < 1) return -1;
    if (B == 0) return 0;
    long long n = (long long)B * S * S;
    hex_observe_kernel<<<grid_for(n, 256), 256, 0, bl_cu(stream)>>>(board, seats, reinterpret_cast<float2 *>(obs), n, S);
    BL_LAUNCH_CHECK();
}

extern "C" int bl_hex_valid(const uint8_t *board, const int32_t *seats, uint8_t *valid, int B, int S,
                            bl_stream stream) {
    if (B < 0 || S < 1) return -1;
    if (B == 0) return 0;
    long long n = (long long)B * S * S;
    hex_valid_kernel<<<grid_for(n, 256), 256, 0, bl_cu(stream)>>>(board, seats, valid, n, S);
    BL_LAUNCH_CHECK();
}
