// Wide (W = 512) tensor-core policy/value network forward for sm_100a: the whole FCModel (boardlaw/networks.py:20-41) for a
// tile of 128 envs in ONE persistent kernel, like net_tc.cu, for the width whose residual stream fills tensor memory.
//
//   layer 0     z  = obs . W_in^T                           (obs generated from the board bytes, exact in fp16)     c_0 = b_in
//   layer 1..D  z += relu(z + c_{k-1}) . (alpha_k W_k)^T      (ReZero residual, boardlaw/networks.py:10-18; in place on z)
//   heads       [policy | value] = (z + c_D) . [W_p ; w_v]^T -> masked log-softmax (heads.py:101-104), tanh (heads.py:136-142)
//
// At W = 512 the fp32 residual stream z of a 128-env tile alone would fill tensor memory (128 lanes x 512 columns), so neither
// z nor the activation operand (net_tc.cu's "TS" form) lives there, and the operand, split hi/lo, does not fit shared memory
// beside a weight ring either (256 KB).  Both take the one remaining short path: per-CTA scratch strips in global memory that
// never leave the L2 (operand 2 x 256 KB, z 256 KB per CTA).  The epilogue writes relu(z + c) as split-fp16 operand tiles,
// already in the UMMA canonical K-major layout, and the TMA engine streams them back through an mbarrier ring exactly like the
// weights; both operands of tcgen05.mma come through shared-memory descriptors ("SS" form).
// fp32 accuracy: x = hi + lo, products hi*hi + hi*lo + lo*hi (three MMAs per K-step).  The tensor core's fp32 accumulate
// truncates, one ulp of the ACCUMULATOR per instruction whatever the addend's size (measured: accumulating all 96 MMAs of a
// layer in place on z costs 3e-5 on the logits at D = 8), so a layer is computed one N = 256 half at a time into two FRESH
// accumulators — hi*hi in columns [0,256), the two small cross terms in [256,512), where their truncation is 2^-11 smaller —
// and z_new = z_old + (acc_hh + acc_lo) is one fp32 addition in the epilogue.
//
// Per-CTA roles (480 threads, 1 CTA/SM, persistent over tiles):
//   warps 0-7   layer group: board staging, one-hot operand, layer epilogues (tcgen05.ld z -> + c -> relu -> split -> scratch)
//   warps 8-11  heads group: softmax / tanh / tree write-back
//   warp 12     weight loader (one thread): cp.async.bulk of host-prepacked weight chunks, runs ahead of the layers freely
//   warp 13     operand loader (one thread): cp.async.bulk of the scratch strip's chunks once the layer group has published them
//   warp 14     MMA issuer (one thread): per N half, per 16-wide K chunk {hi*hi -> acc_hh; hi*lo, lo*hi -> acc_lo}, M128 x N256 x K16
#include <cstdlib>
#include "engine_internal.cuh"
#include "hex_core.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int TILE_M = 128;
constexpr int KC = 16;                     // K elements per ring stage (one K-step)
constexpr int NI = 256;                    // N per instruction
constexpr int L_WARPS = 8, IO_WARPS = 4;
constexpr int WARP_LOADW = 12, WARP_LOADA = 13, WARP_MMA = 14;
constexpr int THREADS = 480;
constexpr int MAX_STAGES = 8;
constexpr int LO_COL = 256;                // TMEM columns of the cross-term accumulator
constexpr int CL = 2;                      // CTAs per cluster: they work on different tiles in lockstep and share every weight chunk —
                                           // each loads 1/CL of it and multicasts it to all (the kernel is L2-bandwidth-bound)
constexpr uint32_t ACT_BLOCK = TILE_M * KC * 2;    // one hi (or lo) operand block of a chunk: 128 rows x 16 K halves
constexpr uint32_t ACT_CHUNK = 2 * ACT_BLOCK;      // hi block, lo block
constexpr uint32_t LBO = 128, SBO = (KC / 8) * 128;   // canonical K-major no-swizzle: K-adjacent / row-group-adjacent core matrices

struct WideParams {
    const uint8_t *board;                  // env e's board at board + e*board_pitch (absolute frame, A bytes); tree mode: tree.board
    const int32_t *seats;                  // (B,); unused in tree mode
    long long board_pitch;
    const uint8_t *blob;                   // packed weight chunks, in consumption order
    const float *cbias;                    // (D+1, W) cumulative biases c_k, then (Np) head bias
    float *logits, *v;                     // (B,A), (B,2) fp32 outputs (NULL in tree mode)
    uint8_t *scratch;                      // per CTA: two operand strips of strip_bytes() each, then all CTAs' z strips (TILE_M * W * 4)
    int B, S, A, W, D, K0p, Np, precision, nstages;
    int tree_mode;
    int l2mode;                            // bit 0: the z strips, bit 1: the operand strips are accessed with an L2 evict_last policy (BL_WIDE_L2)
    bl_tree tree;
    unsigned long long *prof;
};

// scratch-strip accesses with an L2 eviction policy: the strips are rewritten every layer and should never reach DRAM
__device__ __forceinline__ uint64_t wide_policy_keep() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void wide_st16(void *ptr, const uint4 &v, uint64_t pol, bool hint) {
    if (hint) asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(ptr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
    else *reinterpret_cast<uint4 *>(ptr) = v;
}
__device__ __forceinline__ void wide_st16_cg(void *ptr, const uint4 &v, uint64_t pol, bool hint) {
    if (hint) asm volatile("st.global.cg.L2::cache_hint.v4.b32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(ptr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
    else __stcg(reinterpret_cast<uint4 *>(ptr), v);
}
__device__ __forceinline__ uint4 wide_ld16_cg(const void *ptr, uint64_t pol, bool hint) {
    uint4 v;
    if (hint) asm volatile("ld.global.cg.L2::cache_hint.v4.b32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr), "l"(pol) : "memory");
    else v = __ldcg(reinterpret_cast<const uint4 *>(ptr));
    return v;
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}

__host__ __device__ inline int board_pitch_bytes(int A) { return 4 * (((A + 3) / 4) | 1); }
__host__ __device__ inline size_t strip_bytes(int W, int K0p) { return (size_t)TILE_M * (W > K0p ? W : K0p) * 4; }
size_t bias_bytes(int W, int D, int Np) { return ((size_t)(D + 1) * W + Np) * sizeof(float); }
size_t smem_bytes(int nstages, int A, int W, int D, int Np) {
    return ((size_t)NI * KC * 4 + ACT_CHUNK) * nstages + 2 * (size_t)TILE_M * board_pitch_bytes(A) + 1024 + 2 * TILE_M * sizeof(int32_t) +
           2 * TILE_M * 32 + bias_bytes(W, D, Np);
}

#define TCK(k) do { if (p.prof) { const long long now_ = clock64(); pc[k] += now_ - tl; tl = now_; } } while (0)

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(THREADS, 1) fc_tc_wide_kernel(const __grid_constant__ WideParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int W = p.W, D = p.D, A = p.A, S = p.S, Np = p.Np, K0p = p.K0p;
    const uint32_t wstage_bytes = (uint32_t)NI * KC * 4;             // one N half of a K chunk: hi block, lo block
    const int bpitch = board_pitch_bytes(A);
    uint8_t *wstage0 = smem;
    uint8_t *astage0 = wstage0 + (size_t)wstage_bytes * p.nstages;
    uint8_t *btile0 = astage0 + (size_t)ACT_CHUNK * p.nstages;     // [2][TILE_M][bpitch] boards of two consecutive tiles
    const size_t btile_bytes = (size_t)TILE_M * bpitch;
    uint64_t *bars = reinterpret_cast<uint64_t *>(btile0 + 2 * btile_bytes);
    uint64_t *full_w = bars, *full_a = bars + MAX_STAGES, *empty = bars + 2 * MAX_STAGES, *a_ready = bars + 3 * MAX_STAGES,
             *acc_full = a_ready + 1, *oh_ready = a_ready + 2, *heads_full = a_ready + 3, *board_free = a_ready + 4,   // board_free[2]
             *tmem_free = a_ready + 6;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(a_ready + 8);     // keeps sbias 16-byte aligned
    int32_t *tseat0 = reinterpret_cast<int32_t *>(tmem_ptr + 4);  // [2][TILE_M] seat | node << 8 of the tiles' rows
    uint8_t *vmask0 = reinterpret_cast<uint8_t *>(tseat0 + 2 * TILE_M);      // [2][TILE_M][32] legal-move bits of the tiles' rows, 8 cells per byte
    float *sbias = reinterpret_cast<float *>(vmask0 + 2 * TILE_M * 32);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = (p.B + TILE_M - 1) / TILE_M;
    const int nk0 = K0p / KC, nk = W / KC;
    // the cluster's CTAs run the same number of tiles (tile0 is cluster-uniform; a CTA whose tile is past the end works on dead rows)
    uint32_t crank;
    asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    const int tile_first = (int)blockIdx.x - (int)crank;
    const size_t strip = strip_bytes(W, K0p);
    uint8_t *scr = p.scratch + (size_t)blockIdx.x * 2 * strip;      // operand strips: layer L reads strip L & 1
    float *zscr = reinterpret_cast<float *>(p.scratch + (size_t)gridDim.x * 2 * strip + (size_t)blockIdx.x * TILE_M * W * 4);   // [W/16][4][TILE_M] x 16 bytes

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.nstages; s++) {
            mbar_init(smem_u32(full_w + s), 1); mbar_init(smem_u32(full_a + s), 1); mbar_init(smem_u32(empty + s), CL);   // every CTA's MMAs free a slot
        }
        mbar_init(smem_u32(a_ready), L_WARPS);
        mbar_init(smem_u32(acc_full), 1);
        mbar_init(smem_u32(oh_ready), L_WARPS + IO_WARPS);
        mbar_init(smem_u32(heads_full), 1);
        mbar_init(smem_u32(board_free + 0), IO_WARPS);
        mbar_init(smem_u32(board_free + 1), IO_WARPS);
        mbar_init(smem_u32(tmem_free), L_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == WARP_MMA) tmem_alloc(smem_u32(tmem_ptr), 512);
    for (int i = threadIdx.x; i < (D + 1) * W + Np; i += THREADS) sbias[i] = p.cbias[i];
    tc_fence_before();
    __syncthreads();
    cluster_sync();                                                // the peers' barriers exist before anything is multicast at them
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;
    const float *bh = sbias + (size_t)(D + 1) * W;                 // head bias
    const uint64_t keep = wide_policy_keep();
    const bool keep_z = p.l2mode & 1, keep_a = p.l2mode & 2;

    if (warp == WARP_LOADW) {
        // ---- weight loader: the blob is laid out in consumption order, one chunk per stage ----------------------------------
        if (lane == 0) {
            int stage = 0;
            uint32_t ph = 0;
            const uint32_t body_bytes = wstage_bytes, head_bytes = (uint32_t)Np * KC * 4;
            for (int tile0 = tile_first; tile0 < ntiles; tile0 += gridDim.x) {
                const uint8_t *src = p.blob;
                for (int L = 0; L <= D + 1; L++) {
                    const int nch = (L == 0 ? nk0 : nk) * (L <= D ? W / NI : 1);     // the body layers: N half 0's chunks, then half 1's
                    const uint32_t cbytes = L <= D ? body_bytes : head_bytes;
                    for (int c = 0; c < nch; c++) {
                        mbar_wait(smem_u32(empty + stage), ph ^ 1);       // freed by every CTA of the cluster
                        mbar_expect_tx(smem_u32(full_w + stage), cbytes);   // my share + the peers', all landing in my slot
                        const uint32_t share = cbytes / CL;
                        bulk_g2s_multicast(smem_u32(wstage0 + (size_t)wstage_bytes * stage) + crank * share, src + crank * share, share,
                                           smem_u32(full_w + stage), (uint16_t)((1u << CL) - 1));
                        src += cbytes;
                        if (++stage == p.nstages) { stage = 0; ph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == WARP_LOADA) {
        // ---- operand loader: a layer's chunks become loadable when the layer group has published them (a_ready / oh_ready) ------
        if (lane == 0) {
            int stage = 0;
            uint32_t ph = 0, aph = 0, oph = 0;
            for (int tile0 = tile_first; tile0 < ntiles; tile0 += gridDim.x) {
                for (int L = 0; L <= D + 1; L++) {
                    const int nch = L == 0 ? nk0 : nk;
                    // layer 0: the one-hot operand is written AND the previous tile's head accumulator has left z
                    if (L == 0) { mbar_wait(smem_u32(oh_ready), oph); oph ^= 1; }
                    const uint8_t *src = scr + (size_t)(L & 1) * strip;
                    for (int h = 0; h < (L <= D ? W / NI : 1); h++)       // every N half streams the whole operand again
                        for (int c = 0; c < nch; c++) {
                            // the producing layer publishes its operand half by half (a_ready after each half's epilogue): the first
                            // K half can be in the ring while the second is still being written
                            if (L > 0 && h == 0 && c % (nk / (W / NI)) == 0) { mbar_wait(smem_u32(a_ready), aph); aph ^= 1; }
                            mbar_wait(smem_u32(empty + stage), ph ^ 1);
                            mbar_expect_tx(smem_u32(full_a + stage), ACT_CHUNK);
                            if (keep_a) bulk_g2s_hint(smem_u32(astage0 + (size_t)ACT_CHUNK * stage), src + (size_t)c * ACT_CHUNK, ACT_CHUNK, smem_u32(full_a + stage), keep);
                            else bulk_g2s(smem_u32(astage0 + (size_t)ACT_CHUNK * stage), src + (size_t)c * ACT_CHUNK, ACT_CHUNK, smem_u32(full_a + stage));
                            if (++stage == p.nstages) { stage = 0; ph ^= 1; }
                        }
                }
            }
        }
    } else if (warp == WARP_MMA) {
        // ---- MMA issuer ---------------------------------------------------------------------------------------------------------
        if (lane == 0) {
            int stage = 0;
            uint32_t ph = 0, fph = 0;
            const bool lo_b = p.precision == 0;
            long long pc[3] = {0, 0, 0}, tl = p.prof ? clock64() : 0;       // [0] wait operand / accumulator hand-over, [1] wait weights, [2] issue
            for (int tile0 = tile_first; tile0 < ntiles; tile0 += gridDim.x) {
                for (int L = 0; L <= D + 1; L++) {
                    const bool heads = L == D + 1;
                    const int nch = L == 0 ? nk0 : nk;
                    const int ninst = heads ? Np : NI, nhalf = heads ? 1 : W / NI;
                    const bool lo_a = L > 0 && p.precision == 0;
                    const uint32_t idesc = make_idesc(TILE_M, ninst);
                    for (int h = 0; h < nhalf; h++) {
                        // the accumulators are free when the layer group has drained the previous half (the tile's first pass is
                        // gated through the operand loader's oh_ready instead)
                        if (L > 0 || h > 0) { TCK(2); mbar_wait(smem_u32(tmem_free), fph); fph ^= 1; TCK(0); }
                        for (int c = 0; c < nch; c++) {
                            TCK(2);
                            mbar_wait(smem_u32(full_w + stage), ph);
                            TCK(1);
                            mbar_wait(smem_u32(full_a + stage), ph);
                            TCK(0);
                            tc_fence_after();
                            const uint32_t wb = smem_u32(wstage0 + (size_t)wstage_bytes * stage), ab = smem_u32(astage0 + (size_t)ACT_CHUNK * stage);
                            const uint64_t a_hi = make_desc(ab, LBO, SBO), b_hi = make_desc(wb, LBO, SBO);
                            umma_ss(tmem, a_hi, b_hi, idesc, c > 0);
                            if (lo_b) umma_ss(tmem + LO_COL, a_hi, make_desc(wb + (uint32_t)ninst * KC * 2, LBO, SBO), idesc, c > 0);
                            if (lo_a) umma_ss(tmem + LO_COL, make_desc(ab + ACT_BLOCK, LBO, SBO), b_hi, idesc, lo_b || c > 0);
                            umma_commit_multicast(smem_u32(empty + stage), (uint16_t)((1u << CL) - 1));   // frees the stage in every CTA
                            if (++stage == p.nstages) { stage = 0; ph ^= 1; }
                        }
                        umma_commit(smem_u32(heads ? heads_full : acc_full));
                    }
                }
            }
            if (p.prof) { TCK(2); for (int k = 0; k < 3; k++) atomicAdd(p.prof + 16 + k, (unsigned long long)pc[k]); atomicAdd(p.prof + 31, 1ull); }
        }
    } else if (warp < L_WARPS) {
        // ---- layer group -------------------------------------------------------------------------------------------------------------
        const int quad = warp & 3, hh = warp >> 2;
        const int row = quad * 32 + lane;
        const int et = threadIdx.x;                                // 0..255
        const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
        uint32_t accph = 0, hph = 0, bfph[2] = {0, 0};
        long long pc[6] = {0, 0, 0, 0, 0, 0}, tl = p.prof ? clock64() : 0;   // [0] board staging [1] one-hot [2] wait acc [3] layer epilogue [5] wait heads of the previous tile
        uint8_t *my_scr = scr + (size_t)(row >> 3) * SBO + (size_t)(row & 7) * 16;      // my row inside every operand block
        // board staging: two threads per row, each moves half of the row's 4-byte words with cp.async straight into the tile buffer
        // (no registers held across the tile); the NEXT tile's rows are fetched after this tile's first layer
        const int wpr = (A + 3) / 4, wfirst = (et & 1) ? (wpr + 1) / 2 : 0, wcount = (et & 1) ? wpr / 2 : (wpr + 1) / 2;
        const int srow = et >> 1;
        auto fetch_tile = [&](int tile, int nbuf) {
            const int m = tile * TILE_M + srow;
            int32_t sv = -1;
            const uint8_t *src = nullptr;
            if (tile < ntiles && m < p.B) {
                if (p.tree_mode) {
                    const int nd = p.tree.leaf[m];
                    if (nd >= 0) {
                        sv = (int32_t)p.tree.node[(size_t)m * p.tree.T + nd].seat | (nd << 8);
                        src = p.tree.board + ((size_t)m * p.tree.T + nd) * p.tree.BP;
                    }
                } else {
                    sv = p.seats[m];
                    src = p.board + (size_t)m * p.board_pitch;
                }
            }
            uint8_t *dst = btile0 + nbuf * btile_bytes + (size_t)srow * bpitch;
            if (sv >= 0) {                                         // dead rows: the buffer's content is never looked at
                if ((reinterpret_cast<uintptr_t>(src) & 3) == 0) {
                    for (int k = 0; k < wcount; k++)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst + (wfirst + k) * 4)), "l"(src + (wfirst + k) * 4) : "memory");
                } else {
                    for (int k = 0; k < wcount; k++) {
                        uint32_t wd = 0;
                        for (int u = 0; u < 4; u++)
                            if ((wfirst + k) * 4 + u < A) wd |= (uint32_t)src[(wfirst + k) * 4 + u] << (8 * u);
                        *reinterpret_cast<uint32_t *>(dst + (wfirst + k) * 4) = wd;
                    }
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            if ((et & 1) == 0) tseat0[nbuf * TILE_M + srow] = sv;
        };
        fetch_tile(blockIdx.x, 0);
        int it = 0;
        for (int tile0 = tile_first; tile0 < ntiles; tile0 += gridDim.x, it++) {
            const int tile = tile0 + (int)crank;
            const int buf = it & 1;
            uint8_t *btile = btile0 + buf * btile_bytes;
            int32_t *tseat = tseat0 + buf * TILE_M;
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            epi_barrier();
            TCK(0);
            const int m = tile * TILE_M + row;
            const int32_t sv = tseat[row];
            const bool live = m < p.B && sv >= 0;
            const int seat = live ? (sv & 1) : 0;
            const uint8_t *brow = btile + (size_t)row * bpitch;
            // the scratch strip is read by the previous tile's head MMAs: wait for them to retire
            if (it >= 1) { mbar_wait(smem_u32(heads_full), hph); hph ^= 1; tc_fence_after(); }
            TCK(5);
            // ---- observation operand (TensorIntake, heads.py:47-52): feature 2*cell + channel; a 16-feature chunk is 8 cells, a
            //      16-byte core-matrix row is 4 cells (own stone in the low half of each word, opponent's in the high half) ----------
            {
                // cell codes: black = {1,3,4}, white = {2,5,6} (hex_core.cuh); channel 0 = the mover's own stones
                const uint32_t own_mask = seat ? 0x64u : 0x1Au, opp_mask = seat ? 0x1Au : 0x64u;
                for (int c = hh; c < nk0; c += 2) {
                    uint32_t vbits8 = 0;
#pragma unroll
                    for (int g = 0; g < 2; g++) {
                        uint32_t wd[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const int cell = c * 8 + g * 4 + u;
                            uint32_t x = 0;
                            if (live && cell < A) {
                                const int r = cell / S, cc = cell - r * S;
                                const uint32_t cv = brow[seat ? cc * S + r : cell];
                                x = ((own_mask >> cv) & 1u) * 0x3C00u | ((opp_mask >> cv) & 1u) * 0x3C000000u;      // fp16 1.0
                                vbits8 |= (x == 0u ? 1u : 0u) << (g * 4 + u);   // neither side's stone: a legal move (heads.py:101)
                            }
                            wd[u] = x;
                        }
                        *reinterpret_cast<uint4 *>(my_scr + (size_t)c * ACT_CHUNK + g * LBO) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
                    }
                    if (c < 32) vmask0[((size_t)buf * TILE_M + row) * 32 + c] = (uint8_t)vbits8;   // the heads group reads these
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(oh_ready));
            }
            TCK(1);
            // ---- body layers, one N half at a time: z_new = z_old + (acc_hh + acc_lo) -> scratch; relu(z_new + c) -> the next
            //      layer's operand strip (the last layer feeds the heads, without the relu) ---------------------------------------------
            for (int L = 0; L <= D; L++) {
                const float *cb = sbias + (size_t)L * W;
                const bool relu_out = L < D;
                uint8_t *dst_strip = my_scr + (size_t)((L + 1) & 1) * strip;
                for (int h = 0; h < W / NI; h++) {
                    // z_old of a unit is fetched (L2) while the previous unit is processed — the first one while the MMAs still run;
                    // the thread that reads a unit is the one that wrote it a layer ago
                    // z strip layout [unit of 16 columns][4 x 16 bytes][row]: a warp's 16-byte accesses are contiguous (512 B)
                    auto zaddr = [&](int q) {
                        const int col = h * NI + (hh * (NI / 32) + q) * 16;
                        return reinterpret_cast<uint4 *>(zscr) + (size_t)(col / 16) * 4 * TILE_M + row;
                    };
                    uint4 zn4[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) zn4[j] = L > 0 ? wide_ld16_cg(zaddr(0) + j * TILE_M, keep, keep_z) : make_uint4(0, 0, 0, 0);
                    mbar_wait(smem_u32(acc_full), accph);
                    accph ^= 1;
                    tc_fence_after();
                    TCK(2);
                    uint32_t acc_n[16], acl_n[16];
#pragma unroll
                    for (int j = 0; j < 16; j++) acl_n[j] = 0;
                    tmem_ld16(tmem + lane_base + hh * (NI / 2), acc_n);
                    if (p.precision == 0) tmem_ld16(tmem + lane_base + LO_COL + hh * (NI / 2), acl_n);
#pragma unroll 1
                    for (int q = 0; q < NI / 32; q++) {
                        const int lc = (hh * (NI / 32) + q) * 16;      // column inside the half
                        const int col = h * NI + lc;                   // feature column of the unit
                        uint4 *zrow = zaddr(q);
                        uint4 zo[4];
#pragma unroll
                        for (int j = 0; j < 4; j++) zo[j] = zn4[j];
                        if (L > 0 && q + 1 < NI / 32) {
#pragma unroll
                            for (int j = 0; j < 4; j++) zn4[j] = wide_ld16_cg(zaddr(q + 1) + j * TILE_M, keep, keep_z);
                        }
                        uint32_t acc[16], acl[16];
                        tmem_wait_ld();
#pragma unroll
                        for (int j = 0; j < 16; j++) { acc[j] = acc_n[j]; acl[j] = acl_n[j]; }
                        if (q + 1 < NI / 32) {                           // the next unit's accumulators travel during this unit's arithmetic
                            tmem_ld16(tmem + lane_base + lc + 16, acc_n);
                            if (p.precision == 0) tmem_ld16(tmem + lane_base + LO_COL + lc + 16, acl_n);
                        }
                        const float *zf = reinterpret_cast<const float *>(zo);
                        float zn[16];
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            float d = __uint_as_float(acc[j]);
                            if (p.precision == 0) d += __uint_as_float(acl[j]);
                            zn[j] = zf[j] + d;
                        }
                        if (L < D) {
#pragma unroll
                            for (int j = 0; j < 4; j++)
                                wide_st16_cg(zrow + j * TILE_M, make_uint4(__float_as_uint(zn[4 * j]), __float_as_uint(zn[4 * j + 1]),
                                                                          __float_as_uint(zn[4 * j + 2]), __float_as_uint(zn[4 * j + 3])), keep, keep_z);
                        }
                        uint32_t hi2[8], lo2[8];
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float4 c4 = *reinterpret_cast<const float4 *>(cb + col + j);
                            float a0 = zn[j] + c4.x, a1 = zn[j + 1] + c4.y, a2 = zn[j + 2] + c4.z, a3 = zn[j + 3] + c4.w;
                            if (relu_out) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); a2 = fmaxf(a2, 0.f); a3 = fmaxf(a3, 0.f); }
                            const uint32_t h01 = pack_h2(a0, a1), h23 = pack_h2(a2, a3);
                            const float2 f01 = __half22float2(*reinterpret_cast<const __half2 *>(&h01));
                            const float2 f23 = __half22float2(*reinterpret_cast<const __half2 *>(&h23));
                            hi2[j / 2] = h01; hi2[j / 2 + 1] = h23;
                            lo2[j / 2] = pack_h2(a0 - f01.x, a1 - f01.y); lo2[j / 2 + 1] = pack_h2(a2 - f23.x, a3 - f23.y);
                        }
                        // 16 features = one K chunk: two 8-feature core-matrix rows
#pragma unroll
                        for (int u = 0; u < 2; u++) {
                            uint8_t *dst = dst_strip + (size_t)(col / KC) * ACT_CHUNK + u * LBO;
                            wide_st16(dst, make_uint4(hi2[4 * u], hi2[4 * u + 1], hi2[4 * u + 2], hi2[4 * u + 3]), keep, keep_a);
                            if (p.precision == 0)
                                wide_st16(dst + ACT_BLOCK, make_uint4(lo2[4 * u], lo2[4 * u + 1], lo2[4 * u + 2], lo2[4 * u + 3]), keep, keep_a);
                        }
                    }
                    // the accumulators are drained: the next half's (or the heads') MMAs may overwrite them; this half of the next
                    // operand is published
                    tc_fence_before();
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) { mbar_arrive(smem_u32(tmem_free)); mbar_arrive(smem_u32(a_ready)); }
                    TCK(3);
                }
                if (L == 0) {
                    // the next tile's boards: its buffer was the previous tile's, and the heads group is done with that one by now
                    if (it >= 1) { mbar_wait(smem_u32(board_free + (buf ^ 1)), bfph[buf ^ 1]); bfph[buf ^ 1] ^= 1; }
                    fetch_tile(tile + gridDim.x, buf ^ 1);
                }
            }
        }
        if (p.prof && threadIdx.x == 0) for (int k = 0; k < 6; k++) atomicAdd(p.prof + 20 + k, (unsigned long long)pc[k]);
    } else if (warp < L_WARPS + IO_WARPS) {
        // ---- heads group: one warp per lane quadrant, one thread per env row; the accumulator stays in tensor memory (Np up to 256
        //      columns do not fit the registers of a 480-thread CTA), so z is handed to the next tile when the row is written -------
        const int quad = warp - L_WARPS;
        const int row = quad * 32 + lane;
        const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
        uint32_t hph = 0;
        long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tl = p.prof ? clock64() : 0;     // [0] wait heads [3] mask [4] max [5] sum [6] write
        const int nu = Np / 16;
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(oh_ready));            // the first tile needs no hand-over of z
        int it = 0;
        for (int tile0 = tile_first; tile0 < ntiles; tile0 += gridDim.x, it++) {
            const int tile = tile0 + (int)crank;
            const int buf = it & 1;
            const uint8_t *brow = btile0 + buf * btile_bytes + (size_t)row * bpitch;
            const bool last = tile0 + (int)gridDim.x >= ntiles;
            mbar_wait(smem_u32(heads_full), hph);
            hph ^= 1;
            tc_fence_after();
            TCK(0);
            const int m = tile * TILE_M + row;
            const int32_t sv = tseat0[buf * TILE_M + row];
            const bool live = m < p.B && sv >= 0;
            const int seat = live ? (sv & 1) : 0;
            // legal moves of my row, left by the layer group beside the one-hot operand (8 cells per byte); the value column and
            // the padding columns read as illegal
            const uint4 *vrow = reinterpret_cast<const uint4 *>(vmask0 + ((size_t)buf * TILE_M + row) * 32);
            const uint4 va = vrow[0], vc = vrow[1];
            auto vbits = [&](int u) {
                const uint32_t wsel = (u >> 1) == 0 ? va.x : (u >> 1) == 1 ? va.y : (u >> 1) == 2 ? va.z : (u >> 1) == 3 ? va.w
                                    : (u >> 1) == 4 ? vc.x : (u >> 1) == 5 ? vc.y : (u >> 1) == 6 ? vc.z : vc.w;
                const uint32_t bits = (u & 1) ? wsel >> 16 : wsel & 0xFFFFu;
                const int rem = A - u * 16;                        // groups the one-hot stage did not cover hold stale bits
                return rem >= 16 ? bits : (rem <= 0 ? 0u : bits & ((1u << rem) - 1u));
            };
            // raw head outputs of columns [16u, 16u+16) = acc_hh + acc_lo + bias; the NEXT unit's accumulators travel meanwhile
            uint32_t nxt[16], nxl[16];
#pragma unroll
            for (int j = 0; j < 16; j++) nxl[j] = 0;
            auto fetch = [&](int u) {
                tmem_ld16(tmem + lane_base + u * 16, nxt);
                if (p.precision == 0) tmem_ld16(tmem + lane_base + LO_COL + u * 16, nxl);
            };
            auto unit = [&](int u, float (&y)[16]) {
                tmem_wait_ld();
                uint32_t acc[16], acl[16];
#pragma unroll
                for (int j = 0; j < 16; j++) { acc[j] = nxt[j]; acl[j] = nxl[j]; }
                fetch(u + 1 < nu ? u + 1 : 0);                     // (wraps to unit 0: the next pass starts there)
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    float d = __uint_as_float(acc[j]);
                    if (p.precision == 0) d += __uint_as_float(acl[j]);
                    y[j] = d + bh[u * 16 + j];
                }
            };
            fetch(0);
            TCK(3);
            // pass 1: max and sum of exp over the legal actions in ONE sweep (running max, sum rescaled when it moves); the value
            // head's column sits right after the policy's
            float mx = -BL_INF_F, sum = 0.f, vraw = 0.f;
            for (int u = 0; u < nu; u++) {
                float y[16];
                unit(u, y);
                const uint32_t vb = vbits(u);
                float m0 = -BL_INF_F, m1 = -BL_INF_F;
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    m0 = fmaxf(m0, ((vb >> j) & 1u) ? y[j] : -BL_INF_F);
                    m1 = fmaxf(m1, ((vb >> (j + 1)) & 1u) ? y[j + 1] : -BL_INF_F);
                    if (u * 16 + j == A) vraw = y[j];
                    if (u * 16 + j + 1 == A) vraw = y[j + 1];
                }
                const float mn = fmaxf(mx, fmaxf(m0, m1));
                if (mn > -BL_INF_F) {                              // (all-illegal so far: nothing to add, and -inf - -inf is NaN)
                    float s0 = 0.f, s1 = 0.f;
#pragma unroll
                    for (int j = 0; j < 16; j += 2) {
                        if ((vb >> j) & 1u) s0 += __expf(y[j] - mn);
                        if ((vb >> (j + 1)) & 1u) s1 += __expf(y[j + 1] - mn);
                    }
                    sum = sum * __expf(mx - mn) + (s0 + s1);       // exp(-inf) = 0 on the first legal unit
                    mx = mn;
                }
            }
            const float tanh_v = tanhf(vraw);
            TCK(4);
            const float lse = logf(sum);
            TCK(5);
            // pass 3 — tree mode: logits -> half -> exp table -> pi row + row summary, straight into the search tree (what
            // bl_tree_set_eval does for injected evaluations); otherwise fp32 logits / v for the caller
            const int nd = sv >> 8;
            const size_t slot = p.tree_mode && live ? (size_t)m * p.tree.T + nd : 0;
            float pmax = 0.f, pmin = BL_INF_F, pa = 0.f;
            double prun = 0.;                                      // running sum of the pi row -> cpi (prefix sums, rounded once per entry)
            int fz = 255, lz = -1;
            for (int u = 0; u < nu; u++) {
                float y[16];
                unit(u, y);
                const uint32_t vb = vbits(u);
                if (live) {
                    float lg[16];
#pragma unroll
                    for (int j = 0; j < 16; j++) lg[j] = ((vb >> j) & 1u) ? (y[j] - mx) - lse : -BL_INF_F;
                    if (p.tree_mode) {
                        float pv[16];
#pragma unroll
                        for (int j = 0; j < 16; j++) {             // 16 independent table look-ups in flight
                            const int a = u * 16 + j;
                            const bl_half hl = bl_f2h(lg[j]);
                            pv[j] = ((vb >> j) & 1u) ? p.tree.exp_lut[hl] : 0.f;     // exp(-inf) = 0 for illegal / padding columns
                            if (a < A && p.tree.logits) p.tree.logits[slot * A + a] = hl;
                        }
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            const int a = u * 16 + j;
                            if (pv[j] != 0.f) { pmax = fmaxf(pmax, pv[j]); pmin = fminf(pmin, pv[j]); fz = min(fz, a); lz = max(lz, a); }
                        }
#pragma unroll
                        for (int j = 0; j < 16; j += 4)
                            if (u * 16 + j < p.tree.AP)
                                *reinterpret_cast<float4 *>(p.tree.pi + slot * p.tree.AP + u * 16 + j) = make_float4(pv[j], pv[j + 1], pv[j + 2], pv[j + 3]);
                        if (p.tree.cpi) {                          // prefix sums of the row for the certified fast descent (variant 5)
                            float cp[16];
#pragma unroll
                            for (int j = 0; j < 16; j++) {
                                prun += (double)pv[j];
                                cp[j] = (float)prun;
                                pa = __fmaf_rn((float)(u * 16 + j), pv[j], pa);
                            }
#pragma unroll
                            for (int j = 0; j < 16; j += 4)
                                if (u * 16 + j < p.tree.AP)
                                    *reinterpret_cast<float4 *>(p.tree.cpi + slot * p.tree.AP + u * 16 + j) = make_float4(cp[j], cp[j + 1], cp[j + 2], cp[j + 3]);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; j++)
                            if (u * 16 + j < A) p.logits[(size_t)m * A + u * 16 + j] = lg[j];
                    }
                }
            }
            tmem_wait_ld();                                        // (the wrapped prefetch of the last unit)
            // z is free: the next tile's layer 0 may overwrite it
            tc_fence_before();
            __syncwarp();
            if (lane == 0 && !last) mbar_arrive(smem_u32(oh_ready));
            if (live) {
                const float v0 = seat ? -tanh_v : tanh_v, v1 = -v0;
                if (p.tree_mode) {
                    uint32_t *ax = reinterpret_cast<uint32_t *>(p.tree.aux + slot);
                    ax[1] = (uint32_t)bl_f2h(v0) | ((uint32_t)bl_f2h(v1) << 16);
                    reinterpret_cast<uint32_t *>(p.tree.leaf_v)[m] = ax[1];
                    if (p.tree.cpi) p.tree.psum[slot] = pa;
                    reinterpret_cast<uint2 *>(ax)[1] = make_uint2(__float_as_uint(pmax), (__float_as_uint(pmin) >> 16) | ((uint32_t)(fz & 255) << 16) |
                                                                                               ((uint32_t)((lz < 0 ? 0 : lz) & 255) << 24));
                } else {
                    p.v[(size_t)m * 2] = v0;
                    p.v[(size_t)m * 2 + 1] = v1;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(board_free + buf));
            TCK(6);
        }
        if (p.prof && lane == 0 && quad == 0) {
            atomicAdd(p.prof + 25, (unsigned long long)pc[0]);
            for (int k = 3; k < 7; k++) atomicAdd(p.prof + 24 + k, (unsigned long long)pc[k]);     // slots 27..30
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();                                                // no peer multicasts into a CTA that has left
    if (warp == WARP_MMA) tmem_dealloc(tmem, 512);
}

int grid_for(int B) {
    const int ntiles = ((B + TILE_M - 1) / TILE_M + CL - 1) / CL * CL;
    return ntiles < BL_NUM_SMS / CL * CL ? ntiles : BL_NUM_SMS / CL * CL;
}

int launch(const bl_fc_params *p, WideParams &k, int B, void *scratch, cudaStream_t st) {
    const int S = p->S, A = S * S, W = p->W;
    if (scratch == nullptr) return -3;
    k.blob = reinterpret_cast<const uint8_t *>(p->packed);
    k.cbias = p->b_head;
    k.scratch = reinterpret_cast<uint8_t *>(scratch);
    k.prof = bl_phase_prof();
    static int l2mode = -1;
    if (l2mode < 0) { const char *e = getenv("BL_WIDE_L2"); l2mode = e ? atoi(e) & 3 : 0; }
    k.l2mode = l2mode;
    k.B = B; k.S = S; k.A = A; k.W = W; k.D = p->D; k.precision = p->precision;
    k.K0p = (2 * A + KC - 1) / KC * KC;
    k.Np = (A + 1 + 31) / 32 * 32;
    int ns = MAX_STAGES;
    while (ns >= 2 && smem_bytes(ns, A, W, p->D, k.Np) > 227 * 1024) ns--;
    if (ns < 2) return -2;
    k.nstages = ns;
    const size_t smem = smem_bytes(ns, A, W, p->D, k.Np);
    cudaError_t e = cudaFuncSetAttribute(fc_tc_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    fc_tc_wide_kernel<<<grid_for(B), THREADS, smem, st>>>(k);
    return (int)cudaGetLastError();
}

}  // namespace

// shapes the wide path covers: W = 512 (two N = 256 halves; z fills the 512 TMEM columns), heads up to 256 columns
bool bl_fc_wide_supported(const bl_fc_params *p) {
    const int A = p->S * p->S;
    return p->packed != nullptr && p->b_head != nullptr && p->W == 512 && A + 1 <= 256 && p->S <= 13;
}

// bytes of operand scratch the wide kernel needs for B envs (one strip per CTA), 0 for other shapes
int64_t bl_fc_wide_scratch_bytes(const bl_fc_params *p, int B) {
    if (!bl_fc_wide_supported(p)) return 0;
    const int A = p->S * p->S, K0p = (2 * A + KC - 1) / KC * KC;
    return (int64_t)grid_for(B) * (2 * (int64_t)strip_bytes(p->W, K0p) + (int64_t)TILE_M * p->W * 4);
}

int bl_fc_forward_wide(const bl_fc_params *p, const uint8_t *board, long long board_pitch, const int32_t *seats, float *logits,
                       float *v, void *scratch, int B, cudaStream_t st) {
    WideParams k = {};
    k.board = board; k.seats = seats; k.board_pitch = board_pitch;
    k.logits = logits; k.v = v;
    k.tree_mode = 0;
    return launch(p, k, B, scratch, st);
}

// Leaf evaluation straight from / into the search tree (bl_tree_eval_leaves)
int bl_fc_forward_wide_tree(const bl_fc_params *p, const bl_tree *t, void *scratch, cudaStream_t st) {
    WideParams k = {};
    k.tree_mode = 1;
    k.tree = *t;
    return launch(p, k, t->B, scratch, st);
}
