// fp32-accurate GEMM on the tcgen05 tensor cores for the learner step (the forward / dgrad / wgrad contractions of
// FCModel under main.optimize, boardlaw/main.py:75-101, which the reference runs as autocast cuBLAS GEMMs):
//
//     C[M,N] = op(A)[M,K] . op(B)[N,K]^T (+ bias[N])        A, B, C fp32 in global memory, any (row, column) strides for A and B
//
// Same arithmetic as the self-play network kernels (net_tc.cu): every fp32 operand is split x = hi + lo (hi = fp16(x), lo = fp16(x - hi))
// and a product is accumulated in fp32 as hi*hi + hi*lo + lo*hi — three tcgen05.mma (M128 x N<=256 x K16, "SS" form, accumulator in
// tensor memory) per K step.  Gradients do not fit fp16's exponent range, so each operand is first scaled by a power of two taken
// from its max |x| (a device scalar the caller provides: no host sync) — exact — and the product of the two inverse scales is applied
// on the way out of the accumulator.
//
// The operands arrive as fp32, so the TMA engine cannot stage them: sixteen loader warps read fp32 (coalesced for either operand
// orientation: K-major rows as 32-byte row pieces, MN-major as 32 consecutive rows per K), scale / relu / split in registers and write
// 16-byte core-matrix rows of the UMMA canonical K-major layout into a 4-stage shared-memory ring (generic-proxy stores published with
// fence.proxy.async + an mbarrier arrive per warp); one thread issues the MMAs and frees stages with tcgen05.commit; the same
// warps read the accumulator back (tcgen05.ld) and store C through a shared-memory transpose (row-contiguous stores).  wgrad (K = the sample axis) is split over the grid's z dimension into partial
// products that a second kernel sums in a fixed order (deterministic).
#include <cstdlib>

#include "engine_internal.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int G_BM = 128, G_BN = 256, G_KC = 32, G_STAGES = 4;
constexpr int G_LOADERS = 16, G_THREADS = (G_LOADERS + 1) * 32;
constexpr uint32_t G_LBO = 128, G_SBO = (G_KC / 8) * 128;
constexpr uint32_t G_A_BLOCK = G_BM * G_KC * 2, G_B_BLOCK = G_BN * G_KC * 2;
constexpr uint32_t G_STAGE_BYTES = 2 * G_A_BLOCK + 2 * G_B_BLOCK;           // A hi, A lo, B hi, B lo

struct GemmParams {
    const float *A, *B, *bias, *a_amax, *b_amax;
    float *C, *partial;
    long long a_rs, a_cs, b_rs, b_cs, ldc;
    int M, N, K, a_relu, b_relu, splits, kper;
};

// power-of-two scale that brings max|x| into [2^13, 2^14) (fp16 holds up to 65504), and the exponent it used
__device__ __forceinline__ float gemm_scale(float amax, int &e) {
    const uint32_t bits = __float_as_uint(amax) & 0x7FFFFFFFu;
    const int ex = (int)(bits >> 23);
    if (ex == 0 || ex == 255) { e = 0; return 1.f; }                           // zero / denormal / non-finite: leave as is
    e = 13 - (ex - 127);
    e = e > 100 ? 100 : (e < -100 ? -100 : e);
    return __uint_as_float((uint32_t)(127 + e) << 23);
}

// one (row, 8 consecutive k) item of an operand: the global loads ...
__device__ __forceinline__ void gemm_load_item(const float *src, long long rs, long long cs, int grow, int rows, int k0, int kend, float (&v)[8]) {
    if (grow < rows && k0 < kend) {
        const float *p = src + (long long)grow * rs + (long long)k0 * cs;
        if (cs == 1 && k0 + 8 <= kend && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
            const float4 x = __ldg(reinterpret_cast<const float4 *>(p)), y = __ldg(reinterpret_cast<const float4 *>(p) + 1);
            v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = k0 + j < kend ? __ldg(p + (long long)j * cs) : 0.f;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = 0.f;
    }
}
// ... and, a chunk later, its scale / relu / split into the hi and lo core-matrix rows of the stage
__device__ __forceinline__ void gemm_store_item(const float (&v)[8], int row, int kq, float floor, float scale, uint8_t *blk_hi, uint8_t *blk_lo) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
        const float a = fmaxf(v[j] * scale, floor), b = fmaxf(v[j + 1] * scale, floor);     // floor = 0 (relu) or -inf
        const uint32_t h = pack_h2(a, b);
        const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&h));
        hi[j / 2] = h;
        lo[j / 2] = pack_h2(a - f.x, b - f.y);
    }
    const uint32_t off = (uint32_t)(row >> 3) * G_SBO + (uint32_t)kq * G_LBO + (uint32_t)(row & 7) * 16;
    *reinterpret_cast<uint4 *>(blk_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4 *>(blk_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

__global__ void __launch_bounds__(G_THREADS, 1) gemm_tc_kernel(const __grid_constant__ GemmParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)G_STAGES * G_STAGE_BYTES);
    uint64_t *full = bars, *empty = bars + G_STAGES, *acc_full = bars + 2 * G_STAGES;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(acc_full + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = (int)blockIdx.x * G_BM, n0 = (int)blockIdx.y * G_BN;
    const int nrem = p.N - n0 < G_BN ? p.N - n0 : G_BN;
    const int NI = (nrem + 31) & ~31;                              // N of the MMA instruction (multiple of 32, <= 256)
    const int kbeg = (int)blockIdx.z * p.kper, kend = kbeg + p.kper < p.K ? kbeg + p.kper : p.K;
    const int nchunks = (kend - kbeg + G_KC - 1) / G_KC;

    if (threadIdx.x == 0) {
        for (int s = 0; s < G_STAGES; s++) { mbar_init(smem_u32(full + s), G_LOADERS); mbar_init(smem_u32(empty + s), 1); }
        mbar_init(smem_u32(acc_full), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == G_LOADERS) tmem_alloc(smem_u32(tmem_ptr), 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;
    int ea, eb;
    const float sa = gemm_scale(p.a_amax ? *p.a_amax : 1.f, ea), sb = gemm_scale(p.b_amax ? *p.b_amax : 1.f, eb);

    if (warp < G_LOADERS) {
        // ---- loaders -------------------------------------------------------------------------------------------------------------
        // software-pipelined: the global loads of chunk c+1 are in flight (registers) while chunk c is converted and stored, so the
        // loaders never sit out a global-memory round trip per stage
        const bool a_kmajor = p.a_cs == 1, b_kmajor = p.b_cs == 1;
        const int tid = threadIdx.x;                               // 0..511
        const float a_floor = p.a_relu ? 0.f : -BL_INF_F, b_floor = p.b_relu ? 0.f : -BL_INF_F;
        constexpr int AI = G_BM * 4 / (G_LOADERS * 32), BI = G_BN * 4 / (G_LOADERS * 32);     // items per thread: 1 of A, up to 2 of B
        int arow[AI], akq[AI], brow[BI], bkq[BI];
#pragma unroll
        for (int j = 0; j < AI; j++) {
            const int i = tid + j * G_LOADERS * 32;
            if (a_kmajor) { arow[j] = (i & 7) + 8 * (i >> 5); akq[j] = (i >> 3) & 3; }      // a quarter-warp = 8 rows of one octet: conflict-free stores
            else { arow[j] = i & (G_BM - 1); akq[j] = i >> 7; }                          // 32 consecutive rows per k: coalesced loads
        }
#pragma unroll
        for (int j = 0; j < BI; j++) {
            const int i = tid + j * G_LOADERS * 32;
            if (i < NI * 4) {
                if (b_kmajor) { brow[j] = (i & 7) + 8 * (i >> 5); bkq[j] = (i >> 3) & 3; }
                else { brow[j] = i % NI; bkq[j] = i / NI; }
            } else { brow[j] = -1; bkq[j] = 0; }
        }
        float buf0[AI + BI][8], buf1[AI + BI][8];
        auto load = [&](int c, float (&buf)[AI + BI][8]) {
            const int k0 = kbeg + c * G_KC;
#pragma unroll
            for (int j = 0; j < AI; j++) gemm_load_item(p.A, p.a_rs, p.a_cs, m0 + arow[j], p.M, k0 + 8 * akq[j], kend, buf[j]);
#pragma unroll
            for (int j = 0; j < BI; j++)
                if (brow[j] >= 0) gemm_load_item(p.B, p.b_rs, p.b_cs, n0 + brow[j], p.N, k0 + 8 * bkq[j], kend, buf[AI + j]);
        };
        int stage = 0;
        uint32_t ph = 0;
        auto store = [&](const float (&buf)[AI + BI][8]) {
            mbar_wait(smem_u32(empty + stage), ph ^ 1);
            uint8_t *st = smem + (size_t)stage * G_STAGE_BYTES;
#pragma unroll
            for (int j = 0; j < AI; j++) gemm_store_item(buf[j], arow[j], akq[j], a_floor, sa, st, st + G_A_BLOCK);
#pragma unroll
            for (int j = 0; j < BI; j++)
                if (brow[j] >= 0) gemm_store_item(buf[AI + j], brow[j], bkq[j], b_floor, sb, st + 2 * G_A_BLOCK, st + 2 * G_A_BLOCK + G_B_BLOCK);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(full + stage));
            if (++stage == G_STAGES) { stage = 0; ph ^= 1; }
        };
        if (nchunks > 0) load(0, buf0);
        for (int c = 0; c < nchunks; c += 2) {
            if (c + 1 < nchunks) load(c + 1, buf1);
            store(buf0);
            if (c + 1 < nchunks) {
                if (c + 2 < nchunks) load(c + 2, buf0);
                store(buf1);
            }
        }
    } else if (lane == 0) {
        // ---- MMA issuer ------------------------------------------------------------------------------------------------------------
        int stage = 0;
        uint32_t ph = 0;
        const uint32_t idesc = make_idesc(G_BM, NI);
        for (int c = 0; c < nchunks; c++) {
            mbar_wait(smem_u32(full + stage), ph);
            tc_fence_after();
            const uint32_t sbase = smem_u32(smem + (size_t)stage * G_STAGE_BYTES);
#pragma unroll
            for (int kk = 0; kk < G_KC / 16; kk++) {
                const uint32_t off = (uint32_t)kk * 2 * G_LBO;
                const uint64_t a_hi = make_desc(sbase + off, G_LBO, G_SBO), a_lo = make_desc(sbase + G_A_BLOCK + off, G_LBO, G_SBO);
                const uint64_t b_hi = make_desc(sbase + 2 * G_A_BLOCK + off, G_LBO, G_SBO), b_lo = make_desc(sbase + 2 * G_A_BLOCK + G_B_BLOCK + off, G_LBO, G_SBO);
                umma_ss(tmem, a_hi, b_hi, idesc, !(c == 0 && kk == 0));
                umma_ss(tmem, a_hi, b_lo, idesc, 1);
                umma_ss(tmem, a_lo, b_hi, idesc, 1);
            }
            umma_commit(smem_u32(empty + stage));
            if (++stage == G_STAGES) { stage = 0; ph ^= 1; }
        }
        umma_commit(smem_u32(acc_full));
    }
    // ---- epilogue: the sixteen loader warps — warp w reads TMEM lanes 32 (w % 4) .. (its rows), the four warps of a lane quarter take
    //      every fourth 32-column block; a block goes through a private shared-memory tile (the ring is idle by now) so that the
    //      global stores are row-contiguous (8 lanes = 128 B of one row) instead of 32 scattered 16-byte pieces ----------------------
    if (warp < G_LOADERS) {
        if (nchunks > 0) { mbar_wait(smem_u32(acc_full), 0); tc_fence_after(); }
        const int quad = warp & 3, part = warp >> 2;               // G_LOADERS / 4 warps share a lane quarter
        constexpr int TP = 36;                                      // tile row pitch in floats (144 B: conflict-free both ways)
        float *tile = reinterpret_cast<float *>(smem) + (size_t)warp * 32 * TP;
        const float ia = __uint_as_float((uint32_t)(127 - ea) << 23), ib = __uint_as_float((uint32_t)(127 - eb) << 23);
        const int rrow = lane >> 3, c4 = (lane & 7) * 4;            // read-back role: row inside a group of 4, first of 4 columns
        for (int cb = part * 32; cb < NI; cb += 32 * (G_LOADERS / 4)) {
            uint32_t r[32];
            if (nchunks > 0) { tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + cb, r); tmem_wait_ld(); }
            else {
#pragma unroll
                for (int j = 0; j < 32; j++) r[j] = 0;
            }
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4 *>(tile + lane * TP + j) = make_float4(__uint_as_float(r[j]) * ia * ib, __uint_as_float(r[j + 1]) * ia * ib,
                                                                                  __uint_as_float(r[j + 2]) * ia * ib, __uint_as_float(r[j + 3]) * ia * ib);
            __syncwarp();
            const int n = n0 + cb + c4;                             // first of my 4 columns
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias && !p.partial) {
                if (n < p.N) bv.x = p.bias[n];
                if (n + 1 < p.N) bv.y = p.bias[n + 1];
                if (n + 2 < p.N) bv.z = p.bias[n + 2];
                if (n + 3 < p.N) bv.w = p.bias[n + 3];
            }
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int row = i * 4 + rrow, m = m0 + quad * 32 + row;
                float4 x = *reinterpret_cast<const float4 *>(tile + row * TP + c4);
                x.x += bv.x; x.y += bv.y; x.z += bv.z; x.w += bv.w;
                if (m < p.M) {
                    float *dst = p.partial ? p.partial + ((size_t)blockIdx.z * p.M + m) * p.N + n : p.C + (size_t)m * p.ldc + n;
                    if (n + 3 < p.N && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) *reinterpret_cast<float4 *>(dst) = x;
                    else {
                        if (n < p.N) dst[0] = x.x;
                        if (n + 1 < p.N) dst[1] = x.y;
                        if (n + 2 < p.N) dst[2] = x.z;
                        if (n + 3 < p.N) dst[3] = x.w;
                    }
                }
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == G_LOADERS) tmem_dealloc(tmem, 256);
}

// C = sum over the splits of the partial products, in split order (deterministic), + bias
__global__ void __launch_bounds__(256) gemm_reduce_kernel(const float *__restrict__ partial, const float *__restrict__ bias, float *__restrict__ C,
                                                          long long ldc, int M, int N, int splits) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)M * N) return;
    const int m = (int)(i / N), n = (int)(i - (long long)m * N);
    float acc = 0.f;
    for (int z = 0; z < splits; z++) acc += partial[(size_t)z * M * N + i];
    if (bias) acc += bias[n];
    C[(size_t)m * ldc + n] = acc;
}

int gemm_splits(int M, int N, int K) {
    const int tiles = ((M + G_BM - 1) / G_BM) * ((N + G_BN - 1) / G_BN);
    if (tiles >= BL_NUM_SMS / 2 || K < 16 * G_KC) return 1;
    int s = BL_NUM_SMS / tiles;
    const int maxs = K / (8 * G_KC);                               // at least 8 chunks per split
    s = s < maxs ? s : maxs;
    return s < 1 ? 1 : s;
}

}  // namespace

extern "C" int64_t bl_gemm_f32_workspace_bytes(int M, int N, int K) {
    const int s = gemm_splits(M, N, K);
    return s > 1 ? (int64_t)s * M * N * (int64_t)sizeof(float) : 0;
}

extern "C" int bl_gemm_f32(const float *A, long long a_rs, long long a_cs, int a_relu, const float *a_amax, const float *B, long long b_rs,
                           long long b_cs, int b_relu, const float *b_amax, const float *bias, float *C, long long ldc, int M, int N, int K,
                           void *workspace, int64_t workspace_bytes, bl_stream stream) {
    if (M <= 0 || N <= 0 || K < 0 || !A || !B || !C) return -1;
    GemmParams p = {};
    p.A = A; p.B = B; p.bias = bias; p.a_amax = a_amax; p.b_amax = b_amax; p.C = C;
    p.a_rs = a_rs; p.a_cs = a_cs; p.b_rs = b_rs; p.b_cs = b_cs; p.ldc = ldc;
    p.M = M; p.N = N; p.K = K; p.a_relu = a_relu; p.b_relu = b_relu;
    int splits = gemm_splits(M, N, K);
    if (splits > 1 && (!workspace || workspace_bytes < (int64_t)splits * M * N * (int64_t)sizeof(float))) splits = 1;
    p.splits = splits;
    p.kper = splits > 1 ? ((K + splits - 1) / splits + G_KC - 1) / G_KC * G_KC : (K > 0 ? K : 1);
    p.partial = splits > 1 ? reinterpret_cast<float *>(workspace) : nullptr;
    const size_t smem = (size_t)G_STAGES * G_STAGE_BYTES + 256;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr = true;
    }
    const dim3 grid((M + G_BM - 1) / G_BM, (N + G_BN - 1) / G_BN, splits);
    gemm_tc_kernel<<<grid, G_THREADS, smem, bl_cu(stream)>>>(p);
    if (cudaError_t e = cudaGetLastError()) return (int)e;
    if (splits > 1) {
        const long long n = (long long)M * N;
        gemm_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, bl_cu(stream)>>>(p.partial, bias, C, ldc, M, N, splits);
    }
    BL_LAUNCH_CHECK();
}
