// Tensor-core policy/value network forward for sm_100a: the whole FCModel (boardlaw/networks.py:20-41) for a tile of
// 128 envs in ONE kernel — tcgen05.mma with TMEM accumulators, weights streamed by the TMA engine (cp.async.bulk),
// activations never leaving the SM.
//
//   layer 0     x  = obs . W_in^T + b_in                 (obs generated from the board bytes, exact in fp16)
//   layer 1..D  x += alpha_k (relu(x) . W_k^T + b_k)     (ReZero residual, boardlaw/networks.py:10-18)
//   heads       [policy | value] = x . [W_p ; w_v]^T      -> masked log-softmax (heads.py:101-104), tanh (heads.py:136-142)
//
// fp32 accuracy on fp16 tensor cores: every fp32 operand is split x = hi + lo with hi = fp16(x), lo = fp16(x - hi)
// (22 significand bits), and a product is accumulated as hi*hi + hi*lo + lo*hi in fp32 (the dropped lo*lo term is
// 2^-22 relative).  Three tcgen05.mma per K-step instead of one; measured error vs the fp32 reference ~1e-6
// (tests/test_gpu_net.py, tolerance 1e-5).  precision=1 issues only hi*hi (the reference's autocast precision class).
//
// Per-CTA roles (320 threads, 1 CTA/SM, persistent over tiles):
//   warps 0-7  epilogue: TMEM -> registers, bias / ReZero / relu, split to fp16, write the next layer's A operand into
//              shared memory in the UMMA canonical K-major (no-swizzle) layout; residual stream x kept in TMEM cols 256+
//   warp 8     weight loader: one elected thread, cp.async.bulk of host-prepacked operand tiles into a smem ring
//   warp 9     MMA issuer: one elected thread, tcgen05.mma.cta_group::1.kind::f16 (M=128, N<=256, K=16), tcgen05.commit
// Synchronisation is mbarrier-only between roles (full/empty ring, a_ready, acc_full).
#include "common.cuh"
#include "hex_core.cuh"

namespace {

constexpr int TILE_M = 128;
constexpr int KC = 32;                     // K elements per weight stage
constexpr int EPI_WARPS = 8;
constexpr int WARP_LOAD = 8, WARP_MMA = 9;
constexpr int TC_THREADS = 320;
constexpr int X_COL = 256;                 // TMEM column where the residual stream lives
constexpr int MAX_STAGES = 4;

struct TcParams {
    const uint8_t *board;                  // env e's board at board + e*board_pitch (absolute frame, A bytes)
    const int32_t *seats;                  // (B,)
    long long board_pitch;
    const uint8_t *blob;                   // packed operand tiles
    const float *b_in, *b_res, *alpha, *b_head;
    float *logits, *v;                     // (B,A), (B,2)
    int B, S, A, W, D, K0p, Np, precision, nstages;
};

// ---- PTX wrappers ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

#define BL_R8(a, o) "%" #a ""
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
          "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
          "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor): start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version 1 <<46
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// UMMA instruction descriptor (cute::UMMA::InstrDescriptor): F32 accumulate, F16 x F16, both K-major
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct Carve {
    uint8_t *a_hi, *a_lo, *stage0;
    uint32_t stage_bytes;
    uint64_t *full, *empty, *a_ready, *acc_full;
    uint32_t *tmem_ptr;
};
__host__ __device__ inline size_t carve_sizes(int W, int K0p, int Np, int nstages, size_t *a_hi_b, size_t *a_lo_b, size_t *stage_b) {
    int kmax = W > K0p ? W : K0p, nmax = W > Np ? W : Np;
    *a_hi_b = (size_t)TILE_M * kmax * 2;
    *a_lo_b = (size_t)TILE_M * W * 2;
    *stage_b = (size_t)nmax * KC * 2 * 2;
    return *a_hi_b + *a_lo_b + *stage_b * nstages + 256;
}

__global__ void __launch_bounds__(TC_THREADS, 1) fc_tc_kernel(TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    size_t a_hi_b, a_lo_b, stage_b;
    carve_sizes(p.W, p.K0p, p.Np, p.nstages, &a_hi_b, &a_lo_b, &stage_b);
    uint8_t *a_hi = smem, *a_lo = smem + a_hi_b, *stage0 = a_lo + a_lo_b;
    uint64_t *bars = reinterpret_cast<uint64_t *>(stage0 + stage_b * p.nstages);
    uint64_t *full = bars, *empty = bars + MAX_STAGES, *a_ready = bars + 2 * MAX_STAGES, *acc_full = a_ready + 1;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int W = p.W, D = p.D, A = p.A, S = p.S, Np = p.Np, K0p = p.K0p;
    const int ntiles = (p.B + TILE_M - 1) / TILE_M;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.nstages; s++) { mbar_init(smem_u32(full + s), 1); mbar_init(smem_u32(empty + s), 1); }
        mbar_init(smem_u32(a_ready), EPI_WARPS);
        mbar_init(smem_u32(acc_full), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == WARP_MMA) tmem_alloc(smem_u32(tmem_ptr), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;

    if (warp == WARP_LOAD) {
        if (lane == 0) {
            int stage = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const uint8_t *src = p.blob;
                for (int L = 0; L <= D + 1; L++) {
                    const int nch = (L == 0 ? K0p : W) / KC;
                    const uint32_t cbytes = (uint32_t)(L == D + 1 ? Np : W) * KC * 2 * 2;
                    for (int c = 0; c < nch; c++) {
                        mbar_wait(smem_u32(empty + stage), ph ^ 1);
                        mbar_expect_tx(smem_u32(full + stage), cbytes);
                        bulk_g2s(smem_u32(stage0 + stage_b * stage), src, cbytes, smem_u32(full + stage));
                        src += cbytes;
                        if (++stage == p.nstages) { stage = 0; ph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == WARP_MMA) {
        if (lane == 0) {
            int stage = 0;
            uint32_t ph = 0, aph = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int L = 0; L <= D + 1; L++) {
                    const int K = L == 0 ? K0p : W, N = L == D + 1 ? Np : W;
                    const uint32_t idesc = make_idesc(TILE_M, N);
                    const uint32_t sbo_a = (uint32_t)(K / 8) * 128;
                    const bool lo_a = L > 0 && p.precision == 0, lo_b = p.precision == 0;
                    mbar_wait(smem_u32(a_ready), aph);
                    aph ^= 1;
                    tc_fence_after();
                    for (int c = 0; c < K / KC; c++) {
                        mbar_wait(smem_u32(full + stage), ph);
                        tc_fence_after();
                        const uint32_t sb = smem_u32(stage0 + stage_b * stage);
#pragma unroll
                        for (int kk = 0; kk < KC / 16; kk++) {
                            const uint32_t a_off = (uint32_t)(c * (KC / 8) + kk * 2) * 128, b_off = (uint32_t)kk * 2 * 128;
                            const uint64_t d_ahi = make_desc(smem_u32(a_hi) + a_off, 128, sbo_a);
                            const uint64_t d_bhi = make_desc(sb + b_off, 128, (KC / 8) * 128);
                            umma_f16(tmem, d_ahi, d_bhi, idesc, (c | kk) != 0);
                            if (lo_b) umma_f16(tmem, d_ahi, make_desc(sb + (uint32_t)N * KC * 2 + b_off, 128, (KC / 8) * 128), idesc, 1);
                            if (lo_a) umma_f16(tmem, make_desc(smem_u32(a_lo) + a_off, 128, sbo_a), d_bhi, idesc, 1);
                        }
                        umma_commit(smem_u32(empty + stage));      // frees the weight slot when these MMAs retire
                        if (++stage == p.nstages) { stage = 0; ph ^= 1; }
                    }
                    umma_commit(smem_u32(acc_full));
                }
            }
        }
    } else {
        // ---- epilogue warps -------------------------------------------------------------------------------------------
        const int quad = warp & 3, hh = warp >> 2;
        const int row = quad * 32 + lane;
        const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
        uint32_t accph = 0;
        int c_begin, c_end;
        if (W >= 64) { c_begin = hh * (W / 2); c_end = c_begin + W / 2; }
        else { c_begin = 0; c_end = hh == 0 ? W : 0; }
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int m = tile * TILE_M + row;
            const bool live = m < p.B;
            const uint8_t *brow = p.board + (long long)(live ? m : 0) * p.board_pitch;
            const int seat = live ? p.seats[m] : 0;
            // observation operand (TensorIntake, heads.py:47-52): feature k = 2*cell + channel, mover's frame
            if (hh == 0) {
                const uint32_t sbo = (uint32_t)(K0p / 8) * 128;
                uint8_t *dst = a_hi + (row >> 3) * sbo + (row & 7) * 16;
                for (int j = 0; j < K0p / 8; j++) {                  // 8 features = 4 cells per 16-byte chunk
                    uint32_t wds[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int cell = j * 4 + u;
                        uint32_t wd = 0;
                        if (live && cell < A) {
                            const int r = cell / S, c = cell - r * S;
                            const uint8_t cv = brow[seat ? c * S + r : cell];
                            const bool black = cv == BL_BLACK || cv == BL_TOP || cv == BL_BOT;
                            const bool white = cv == BL_WHITE || cv == BL_LEFT || cv == BL_RIGHT;
                            const bool own = seat ? white : black, opp = seat ? black : white;
                            wd = (own ? 0x3C00u : 0u) | (opp ? 0x3C000000u : 0u);      // fp16 1.0
                        }
                        wds[u] = wd;
                    }
                    *reinterpret_cast<uint4 *>(dst + j * 128) = make_uint4(wds[0], wds[1], wds[2], wds[3]);
                }
            }
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(a_ready));

            for (int L = 0; L <= D; L++) {
                mbar_wait(smem_u32(acc_full), accph);
                accph ^= 1;
                tc_fence_after();
                const float *bias = L == 0 ? p.b_in : p.b_res + (size_t)(L - 1) * W;
                const float alpha = L == 0 ? 0.f : p.alpha[L - 1];
                const bool relu_out = L < D;
                const uint32_t sbo = (uint32_t)(W / 8) * 128;
                uint8_t *dhi = a_hi + (row >> 3) * sbo + (row & 7) * 16, *dlo = a_lo + (row >> 3) * sbo + (row & 7) * 16;
                for (int c0 = c_begin; c0 < c_end; c0 += 32) {
                    uint32_t acc[32], xr[32];
                    tmem_ld32(tmem + lane_base + c0, acc);
                    if (L > 0) tmem_ld32(tmem + lane_base + X_COL + c0, xr);
                    tmem_wait_ld();
                    uint32_t hi2[16], lo2[16];
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        float y = __uint_as_float(acc[j]) + bias[c0 + j];
                        float xn = L == 0 ? y : fmaf(alpha, y, __uint_as_float(xr[j]));
                        xr[j] = __float_as_uint(xn);
                        float a = relu_out ? fmaxf(xn, 0.f) : xn;
                        __half h = __float2half_rn(a);
                        __half l = __float2half_rn(a - __half2float(h));
                        if (j & 1) { hi2[j >> 1] |= (uint32_t)__half_as_ushort(h) << 16; lo2[j >> 1] |= (uint32_t)__half_as_ushort(l) << 16; }
                        else { hi2[j >> 1] = __half_as_ushort(h); lo2[j >> 1] = __half_as_ushort(l); }
                    }
                    tmem_st32(tmem + lane_base + X_COL + c0, xr);
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        *reinterpret_cast<uint4 *>(dhi + (c0 / 8 + u) * 128) = make_uint4(hi2[4 * u], hi2[4 * u + 1], hi2[4 * u + 2], hi2[4 * u + 3]);
                        *reinterpret_cast<uint4 *>(dlo + (c0 / 8 + u) * 128) = make_uint4(lo2[4 * u], lo2[4 * u + 1], lo2[4 * u + 2], lo2[4 * u + 3]);
                    }
                }
                tmem_wait_st();
                fence_proxy_async();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(a_ready));
            }
            // ---- heads --------------------------------------------------------------------------------------------------
            mbar_wait(smem_u32(acc_full), accph);
            accph ^= 1;
            tc_fence_after();
            if (hh == 0) {
                float mx = -BL_INF_F, tanh_v = 0.f;
                for (int c0 = 0; c0 < Np; c0 += 32) {
                    uint32_t acc[32];
                    tmem_ld32(tmem + lane_base + c0, acc);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        const int a = c0 + j;
                        if (a < A) {
                            const int r = a / S, c = a - r * S;
                            if (brow[seat ? c * S + r : a] == BL_EMPTY) mx = fmaxf(mx, __uint_as_float(acc[j]) + p.b_head[a]);
                        } else if (a == A) {
                            tanh_v = tanhf(__uint_as_float(acc[j]) + p.b_head[A]);
                        }
                    }
                }
                float sum = 0.f;
                for (int c0 = 0; c0 < Np; c0 += 32) {
                    uint32_t acc[32];
                    tmem_ld32(tmem + lane_base + c0, acc);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        const int a = c0 + j;
                        if (a < A) {
                            const int r = a / S, c = a - r * S;
                            if (brow[seat ? c * S + r : a] == BL_EMPTY) sum += expf((__uint_as_float(acc[j]) + p.b_head[a]) - mx);
                        }
                    }
                }
                const float lse = logf(sum);
                for (int c0 = 0; c0 < Np; c0 += 32) {
                    uint32_t acc[32];
                    tmem_ld32(tmem + lane_base + c0, acc);
                    tmem_wait_ld();
                    if (live) {
#pragma unroll
                        for (int j = 0; j < 32; j++) {
                            const int a = c0 + j;
                            if (a < A) {
                                const int r = a / S, c = a - r * S;
                                const bool valid = brow[seat ? c * S + r : a] == BL_EMPTY;
                                p.logits[(size_t)m * A + a] = valid ? ((__uint_as_float(acc[j]) + p.b_head[a]) - mx) - lse : -BL_INF_F;
                            }
                        }
                    }
                }
                if (live) {
                    p.v[(size_t)m * 2 + seat] = tanh_v;
                    p.v[(size_t)m * 2 + (1 - seat)] = -tanh_v;
                }
            }
            tc_fence_before();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == WARP_MMA) tmem_dealloc(tmem, 512);
}

}  // namespace

// Host-side launcher, called from bl_fc_forward (net.cu) when a packed operand blob is present and the shape fits.
int bl_fc_forward_tc(const bl_fc_params *p, const uint8_t *board, long long board_pitch, const int32_t *seats, float *logits,
                     float *v, int B, cudaStream_t st) {
    const int S = p->S, A = S * S, W = p->W;
    TcParams k;
    k.board = board; k.seats = seats; k.board_pitch = board_pitch;
    k.blob = reinterpret_cast<const uint8_t *>(p->packed);
    k.b_in = p->b_in; k.b_res = p->b_res; k.alpha = p->alpha; k.b_head = p->b_head;
    k.logits = logits; k.v = v;
    k.B = B; k.S = S; k.A = A; k.W = W; k.D = p->D; k.precision = p->precision;
    k.K0p = (2 * A + KC - 1) / KC * KC;
    k.Np = (A + 1 + 31) / 32 * 32;
    size_t a, b, c;
    int ns = MAX_STAGES;
    while (ns >= 2 && carve_sizes(W, k.K0p, k.Np, ns, &a, &b, &c) > 227 * 1024) ns--;
    if (ns < 2) return -2;
    k.nstages = ns;
    const size_t smem = carve_sizes(W, k.K0p, k.Np, ns, &a, &b, &c);
    cudaError_t e = cudaFuncSetAttribute(fc_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int ntiles = (B + TILE_M - 1) / TILE_M;
    const int grid = ntiles < BL_NUM_SMS ? ntiles : BL_NUM_SMS;
    fc_tc_kernel<<<grid, TC_THREADS, smem, st>>>(k);
    return (int)cudaGetLastError();
}

// shapes the tensor-core path covers: W in {32, 64, 128, 256} (tile N = W <= 256 TMEM columns next to the residual stream)
bool bl_fc_tc_supported(const bl_fc_params *p) {
    const int A = p->S * p->S, W = p->W;
    return p->packed != nullptr && p->b_head != nullptr && (W == 32 || W == 64 || W == 128 || W == 256) && A + 1 <= 256;
}
