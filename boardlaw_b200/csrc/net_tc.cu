// Tensor-core policy/value network forward for sm_100a: the whole FCModel (boardlaw/networks.py:20-41) for a tile of
// 128 envs in ONE kernel — tcgen05.mma with both the accumulator AND the activation operand in tensor memory, weights
// streamed by the TMA engine (cp.async.bulk) through an mbarrier ring; activations never touch shared or global memory.
//
//   layer 0     z  = obs . W_in^T                        (obs generated from the board bytes, exact in fp16)      c_0 = b_in
//   layer 1..D  z += relu(z + c_{k-1}) . (alpha_k W_k)^T   (ReZero residual, boardlaw/networks.py:10-18; the MMA accumulates IN PLACE on z)
//                                                          c_k = c_{k-1} + alpha_k b_k   (biases never enter the accumulator)
//   heads       [policy | value] = (z + c_D) . [W_p ; w_v]^T -> masked log-softmax (heads.py:101-104), tanh (heads.py:136-142)
//
// fp32 accuracy on fp16 tensor cores: every fp32 operand is split x = hi + lo with hi = fp16(x), lo = fp16(x - hi)
// (22 significand bits), and a product is accumulated as hi*hi + hi*lo + lo*hi in fp32 (the dropped lo*lo term is
// 2^-22 relative): three tcgen05.mma per K-step.  Measured error vs the fp32 reference ~1e-6 (tests/test_gpu_net.py,
// tolerance 1e-5).  precision=1 issues only hi*hi (the reference's autocast precision class).
//
// Tensor memory (512 columns x 128 lanes, lane = env row of the tile):
//   [0,256)   z, fp32 (the heads' accumulator re-uses it once z is dead)
//   [256,384) activation operand, hi halves, two K elements per column (A of tcgen05.mma in "TS" form)
//   [384,512) activation operand, lo halves (layer 0's one-hot operand has no lo part and may use both regions)
//
// Per-CTA roles (448 threads, 1 CTA/SM, persistent over tiles):
//   warps 0-7   layer group: board staging, one-hot operand, layer epilogues (tcgen05.ld z -> + c -> relu -> split -> tcgen05.st the
//               next layer's operand)
//   warps 8-11  heads group: masked log-softmax / tanh / tree write-back of the PREVIOUS tile, off the critical path
//   warp 12     weight loader: one elected thread, cp.async.bulk of host-prepacked operand tiles into a smem ring
//   warp 13     MMA issuer: one elected thread, tcgen05.mma.cta_group::1.kind::f16 (M=128, whole-N tiles, K=16)
// A layer is issued as whole-N MMAs in K order (tc_nsplit = 1): the (N half, K half) block order that lets half-epilogues run under the
// other half's MMAs is kept in the packer (tc_nsplit = 2) but loses, because one tcgen05.mma issue costs ~120 cycles from a single
// thread and N = 128 instructions carry only 64 cycles of tensor work.
// Synchronisation is mbarrier-only between roles (weight ring full/empty, a_ready, acc_full, oh_ready, heads_full, board_free[2]).
#include "engine_internal.cuh"
#include "hex_core.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int TILE_M = 128;
constexpr int KC = 32;                     // K elements per weight stage
constexpr int L_WARPS = 8;                 // layer group: board staging, one-hot operand, layer epilogues
constexpr int IO_WARPS = 4;                // heads group: softmax / tanh / tree write-back of the PREVIOUS tile, off the critical path
constexpr int WARP_LOAD = 12, WARP_MMA = 13;
constexpr int TC_THREADS = 448;
constexpr int AH_COL = 256, AL_COL = 384;  // TMEM columns of the activation operand (hi, lo)
constexpr int MAX_STAGES = 8;
constexpr int VGROUPS = 16;                // 16-cell groups of legal-move bits per row (A <= 256)

struct TcParams {
    const uint8_t *board;                  // env e's board at board + e*board_pitch (absolute frame, A bytes); tree mode: tree.board
    const int32_t *seats;                  // (B,); unused in tree mode
    long long board_pitch;
    const uint8_t *blob;                   // packed operand tiles, in consumption order
    const float *cbias;                    // (D+1, W) cumulative biases c_k, then (Np) head bias
    float *logits, *v;                     // (B,A), (B,2) fp32 outputs (NULL in tree mode)
    int B, S, A, W, D, K0p, Np, precision, nstages, nsplit;
    int tree_mode;                         // 1: inputs are the current leaves of `tree`, outputs go straight into the tree
    bl_tree tree;
    unsigned long long *prof;              // optional phase clock (bl_debug_set_phase_profile), slots 16..31
};

struct Shape {              // derived sizes shared by the roles and the host packer (networks.py mirrors them)
    int nk0, nk;            // K chunks of layer 0 / of the residual layers and heads
    uint32_t stage_bytes;
};
__host__ __device__ inline Shape make_shape(int W, int K0p, int Np) {
    Shape s;
    s.nk0 = K0p / KC;
    s.nk = W / KC;
    const int rows = W > Np ? W : Np;
    s.stage_bytes = (uint32_t)rows * KC * 2 * 2;
    return s;
}
__host__ __device__ inline int board_pitch_bytes(int A) { return 4 * (((A + 3) / 4) | 1); }   // odd number of words: conflict-free rows
size_t bias_bytes(int W, int D, int Np) { return ((size_t)(D + 1) * W + Np) * sizeof(float); }
size_t smem_bytes(const Shape &s, int nstages, int A, int W, int D, int Np) {
    return (size_t)s.stage_bytes * nstages + 2 * (size_t)TILE_M * board_pitch_bytes(A) + 1024 + 2 * TILE_M * sizeof(int32_t) +
           2 * TILE_M * VGROUPS * sizeof(uint16_t) + bias_bytes(W, D, Np);
}

#define TCK(k) do { if (p.prof) { const long long now_ = clock64(); pc[k] += now_ - tl; tl = now_; } } while (0)

__global__ void __launch_bounds__(TC_THREADS, 1) fc_tc_kernel(const __grid_constant__ TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const Shape sh = make_shape(p.W, p.K0p, p.Np);
    const int bpitch = board_pitch_bytes(p.A);                     // board tile row pitch
    uint8_t *stage0 = smem;
    uint8_t *btile0 = stage0 + (size_t)sh.stage_bytes * p.nstages; // [2][TILE_M][bpitch] boards of two consecutive tiles
    const size_t btile_bytes = (size_t)TILE_M * bpitch;            // multiple of 512
    uint64_t *bars = reinterpret_cast<uint64_t *>(btile0 + 2 * btile_bytes);
    uint64_t *full = bars, *empty = bars + MAX_STAGES, *a_ready = bars + 2 * MAX_STAGES, *acc_full = a_ready + 1, *oh_ready = a_ready + 2,
             *heads_full = a_ready + 3, *board_free = a_ready + 4;   // board_free[2]
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(a_ready + 6);
    int32_t *tseat0 = reinterpret_cast<int32_t *>(tmem_ptr + 4);  // [2][TILE_M] seat | node << 8 of the tiles' rows
    uint16_t *vmask0 = reinterpret_cast<uint16_t *>(tseat0 + 2 * TILE_M);   // [2][TILE_M][VGROUPS] legal-move bits of the tiles' rows, 16 cells per entry
    float *sbias = reinterpret_cast<float *>(vmask0 + 2 * TILE_M * VGROUPS); // (D+1, W) cumulative biases, then the head bias (Np)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int W = p.W, D = p.D, A = p.A, S = p.S, Np = p.Np, K0p = p.K0p;
    const int ntiles = (p.B + TILE_M - 1) / TILE_M;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.nstages; s++) { mbar_init(smem_u32(full + s), 1); mbar_init(smem_u32(empty + s), 1); }
        mbar_init(smem_u32(a_ready), L_WARPS);
        mbar_init(smem_u32(acc_full), 1);
        mbar_init(smem_u32(oh_ready), L_WARPS + IO_WARPS);
        mbar_init(smem_u32(heads_full), 1);
        mbar_init(smem_u32(board_free + 0), IO_WARPS);
        mbar_init(smem_u32(board_free + 1), IO_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == WARP_MMA) tmem_alloc(smem_u32(tmem_ptr), 512);
    for (int i = threadIdx.x; i < (p.D + 1) * p.W + p.Np; i += TC_THREADS) sbias[i] = p.cbias[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;
    const float *bh = sbias + (size_t)(D + 1) * W;                 // head bias

    if (warp == WARP_LOAD) {
        // ---- weight loader: the blob is laid out in consumption order, one chunk per stage ----------------------------------
        if (lane == 0) {
            int stage = 0;
            uint32_t ph = 0;
            const uint32_t body_bytes = (uint32_t)W * KC * 2 * 2, head_bytes = (uint32_t)Np * KC * 2 * 2;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const uint8_t *src = p.blob;
                for (int L = 0; L <= D + 1; L++) {
                    const int nch = L == 0 ? sh.nk0 : sh.nk;
                    const uint32_t cbytes = L <= D ? body_bytes : head_bytes;
                    for (int c = 0; c < nch; c++) {
                        mbar_wait(smem_u32(empty + stage), ph ^ 1);
                        mbar_expect_tx(smem_u32(full + stage), cbytes);
                        bulk_g2s(smem_u32(stage0 + (size_t)sh.stage_bytes * stage), src, cbytes, smem_u32(full + stage));
                        src += cbytes;
                        if (++stage == p.nstages) { stage = 0; ph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == WARP_MMA) {
        // ---- MMA issuer ---------------------------------------------------------------------------------------------------------
        if (lane == 0) {
            int stage = 0;
            uint32_t ph = 0, aph = 0, oph = 0;
            const bool lo_b = p.precision == 0;
            long long pc[3] = {0, 0, 0}, tl = p.prof ? clock64() : 0;       // [0] wait operand, [1] wait weights, [2] issue
            // one K chunk (KC = 2 K-steps of 16): D columns [0, N), A columns from the chunk's K offset
            auto chunk = [&](int N, int kchunk, bool lo_a, bool fresh_acc) {
                TCK(2);
                mbar_wait(smem_u32(full + stage), ph);
                TCK(1);
                tc_fence_after();
                const uint32_t sb = smem_u32(stage0 + (size_t)sh.stage_bytes * stage);
                const uint32_t idesc = make_idesc(TILE_M, N);
#pragma unroll
                for (int kk = 0; kk < KC / 16; kk++) {
                    const uint32_t acol = (uint32_t)(kchunk * KC + kk * 16) / 2, b_off = (uint32_t)kk * 2 * 128;
                    const uint64_t d_bhi = make_desc(sb + b_off, 128, (KC / 8) * 128);
                    umma_ts(tmem, tmem + AH_COL + acol, d_bhi, idesc, !(fresh_acc && kk == 0));
                    if (lo_b) umma_ts(tmem, tmem + AH_COL + acol, make_desc(sb + (uint32_t)N * KC * 2 + b_off, 128, (KC / 8) * 128), idesc, 1);
                    if (lo_a) umma_ts(tmem, tmem + AL_COL + acol, d_bhi, idesc, 1);
                }
                umma_commit(smem_u32(empty + stage));              // frees the weight slot when these MMAs retire
                if (++stage == p.nstages) { stage = 0; ph ^= 1; }
            };
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int L = 0; L <= D; L++) {
                    const int nk = L == 0 ? sh.nk0 : sh.nk;
                    const bool lo_a = L > 0 && p.precision == 0;
                    TCK(2);
                    // layer 0: the one-hot operand is built AND the previous tile's head accumulator has left z
                    if (L == 0) { mbar_wait(smem_u32(oh_ready), oph); oph ^= 1; }
                    else { mbar_wait(smem_u32(a_ready), aph); aph ^= 1; }
                    TCK(0);
                    tc_fence_after();
                    // layer 0 overwrites z (first K chunk), the residual layers accumulate onto it
                    for (int c = 0; c < nk; c++) chunk(W, c, lo_a, L == 0 && c == 0);
                    umma_commit(smem_u32(acc_full));
                }
                // heads: the accumulator re-uses z's columns [0, Np) once the last layer's epilogue has consumed z
                TCK(2);
                mbar_wait(smem_u32(a_ready), aph); aph ^= 1;
                TCK(0);
                tc_fence_after();
                for (int c = 0; c < sh.nk; c++) chunk(Np, c, p.precision == 0, c == 0);
                umma_commit(smem_u32(heads_full));
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
        const int cpw = W / 32;                                    // 32-column chunks of a layer
        const int ch_begin = cpw >= 2 ? hh * (cpw / 2) : 0, ch_end = cpw >= 2 ? ch_begin + cpw / 2 : (hh == 0 ? cpw : 0);
        // board staging: two threads per row, each moves half of the row's 4-byte words; the NEXT tile's words travel in
        // registers while the current tile is computed
        constexpr int MAXW = 22;                                   // words per thread: A <= 169 -> 43 words per row
        const int wpr = (A + 3) / 4, wfirst = (et & 1) ? (wpr + 1) / 2 : 0, wcount = (et & 1) ? wpr / 2 : (wpr + 1) / 2;
        const int srow = et >> 1;
        uint32_t bw[MAXW];
        int32_t sv_next = -1;
        auto fetch_tile = [&](int tile) {                          // issue the loads of `tile`'s board words and seat for row `srow`
            const int m = tile * TILE_M + srow;
            sv_next = -1;
#pragma unroll
            for (int k = 0; k < MAXW; k++) bw[k] = 0;
            if (tile < ntiles && m < p.B) {
                const uint8_t *src;
                if (p.tree_mode) {
                    const int nd = p.tree.leaf[m];
                    if (nd >= 0) {
                        sv_next = (int32_t)p.tree.node[(size_t)m * p.tree.T + nd].seat | (nd << 8);
                        src = p.tree.board + ((size_t)m * p.tree.T + nd) * p.tree.BP;
                    }
                } else {
                    sv_next = p.seats[m];
                    src = p.board + (size_t)m * p.board_pitch;
                }
                if (sv_next >= 0) {
                    if (p.tree_mode || ((p.board_pitch & 3) == 0 && (reinterpret_cast<uintptr_t>(p.board) & 3) == 0)) {
#pragma unroll
                        for (int k = 0; k < MAXW; k++)
                            if (k < wcount) bw[k] = reinterpret_cast<const uint32_t *>(src)[wfirst + k];
                    } else {
#pragma unroll
                        for (int k = 0; k < MAXW; k++)
                            if (k < wcount)
                                for (int u = 0; u < 4; u++)
                                    if ((wfirst + k) * 4 + u < A) bw[k] |= (uint32_t)src[(wfirst + k) * 4 + u] << (8 * u);
                    }
                }
            }
        };
        fetch_tile(blockIdx.x);
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
            const int buf = it & 1;
            uint8_t *btile = btile0 + buf * btile_bytes;
            int32_t *tseat = tseat0 + buf * TILE_M;
            // ---- the tile's boards (fetched during the previous tile) -> shared memory; the heads group must be done with the
            //      buffer's previous tenant (two tiles ago) ----------------------------------------------------------------------------
            if (it >= 2) { mbar_wait(smem_u32(board_free + buf), bfph[buf]); bfph[buf] ^= 1; }
#pragma unroll
            for (int k = 0; k < MAXW; k++)
                if (k < wcount) *reinterpret_cast<uint32_t *>(btile + (size_t)srow * bpitch + (wfirst + k) * 4) = bw[k];
            if ((et & 1) == 0) tseat[srow] = sv_next;
            epi_barrier();
            fetch_tile(tile + gridDim.x);
            TCK(0);
            const int m = tile * TILE_M + row;
            const int32_t sv = tseat[row];
            const bool live = m < p.B && sv >= 0;
            const int seat = live ? (sv & 1) : 0;
            const uint8_t *brow = btile + (size_t)row * bpitch;
            // the operand columns are read by the previous tile's head MMAs: wait for them to retire
            if (it >= 1) { mbar_wait(smem_u32(heads_full), hph); hph ^= 1; tc_fence_after(); }
            TCK(5);
            // ---- observation operand (TensorIntake, heads.py:47-52): feature 2*cell + channel = one packed column per cell ---------
            {
                // columns split between the two warps of the quadrant when both parts are whole 16-column stores
                const int ncol = K0p / 2;
                const bool split = (ncol / 2) % 16 == 0;
                const int my_cols = split ? ncol / 2 : (hh == 0 ? ncol : 0);
                const int cb = split ? hh * (ncol / 2) : 0;
                // cell codes: black = {1,3,4}, white = {2,5,6} (hex_core.cuh); channel 0 = the mover's own stones
                const uint32_t own_mask = seat ? 0x64u : 0x1Au, opp_mask = seat ? 0x1Au : 0x64u;
                int r = cb / S, c = cb - r * S;                    // (row, col) of cell `cb`
                for (int j0 = 0; j0 < my_cols; j0 += 16) {
                    uint32_t wds[16], vbits16 = 0;
#pragma unroll
                    for (int u = 0; u < 16; u++) {
                        const int cell = cb + j0 + u;
                        uint32_t wd = 0;
                        if (live && cell < A) {
                            const uint32_t cv = brow[seat ? c * S + r : cell];
                            wd = ((own_mask >> cv) & 1u) * 0x3C00u | ((opp_mask >> cv) & 1u) * 0x3C000000u;      // fp16 1.0
                            vbits16 |= (wd == 0u ? 1u : 0u) << u;      // neither side's stone: a legal move (heads.py:101)
                        }
                        wds[u] = wd;
                        if (++c == S) { c = 0; r++; }
                    }
                    tmem_st16(tmem + lane_base + AH_COL + cb + j0, wds);
                    // the heads group reads the legal-move bits instead of walking the board again
                    vmask0[((size_t)buf * TILE_M + row) * VGROUPS + ((cb + j0) >> 4)] = (uint16_t)vbits16;
                }
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(oh_ready));
            }
            TCK(1);
            // ---- body layers: z -> next operand -------------------------------------------------------------------------------------
            for (int L = 0; L <= D; L++) {
                const float *cb = sbias + (size_t)L * W;
                const bool relu_out = L < D;
                mbar_wait(smem_u32(acc_full), accph);
                accph ^= 1;
                tc_fence_after();
                TCK(2);
                uint32_t nxt[32];
                if (ch_begin < ch_end) tmem_ld32(tmem + lane_base + ch_begin * 32, nxt);
                for (int ch = ch_begin; ch < ch_end; ch++) {
                    const int col = ch * 32;
                    uint32_t acc[32];
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 32; j++) acc[j] = nxt[j];
                    if (ch + 1 < ch_end) tmem_ld32(tmem + lane_base + col + 32, nxt);      // in flight during this chunk's arithmetic
                    uint32_t hi2[16], lo2[16];
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 c4 = *reinterpret_cast<const float4 *>(cb + col + j);
                        float a0 = __uint_as_float(acc[j]) + c4.x, a1 = __uint_as_float(acc[j + 1]) + c4.y;
                        float a2 = __uint_as_float(acc[j + 2]) + c4.z, a3 = __uint_as_float(acc[j + 3]) + c4.w;
                        if (relu_out) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); a2 = fmaxf(a2, 0.f); a3 = fmaxf(a3, 0.f); }
                        const uint32_t h01 = pack_h2(a0, a1), h23 = pack_h2(a2, a3);
                        const float2 f01 = __half22float2(*reinterpret_cast<const __half2 *>(&h01));
                        const float2 f23 = __half22float2(*reinterpret_cast<const __half2 *>(&h23));
                        hi2[j / 2] = h01; hi2[j / 2 + 1] = h23;
                        lo2[j / 2] = pack_h2(a0 - f01.x, a1 - f01.y); lo2[j / 2 + 1] = pack_h2(a2 - f23.x, a3 - f23.y);
                    }
                    tmem_st16(tmem + lane_base + AH_COL + col / 2, hi2);
                    if (p.precision == 0) tmem_st16(tmem + lane_base + AL_COL + col / 2, lo2);
                }
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(a_ready));
                TCK(3);
            }
        }
        if (p.prof && threadIdx.x == 0) for (int k = 0; k < 6; k++) atomicAdd(p.prof + 20 + k, (unsigned long long)pc[k]);
    } else {
        // ---- heads group: one warp per lane quadrant, one thread per env row; works on tile i while the other groups are already
        //      on tile i+1 (the accumulator is pulled into registers first when it fits, which frees z for the next tile) -----------
        const int quad = warp - L_WARPS;
        const int row = quad * 32 + lane;
        const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
        uint32_t hph = 0;
        long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tl = p.prof ? clock64() : 0;     // [0] wait heads [1] tail [2] acc->regs [3] mask [4] max [5] sum [6] write
        const int nu = Np / 16;
        // the first tile needs no hand-over of z
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(oh_ready));
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
            const int buf = it & 1;
            const uint8_t *brow = btile0 + buf * btile_bytes + (size_t)row * bpitch;
            const bool last = tile + (int)gridDim.x >= ntiles;
            mbar_wait(smem_u32(heads_full), hph);
            hph ^= 1;
            tc_fence_after();
            TCK(0);
            const int m = tile * TILE_M + row;
            const int32_t sv = tseat0[buf * TILE_M + row];
            const bool live = m < p.B && sv >= 0;
            const int seat = live ? (sv & 1) : 0;
            TCK(2);
            // legal moves of my row, 16 cells per entry, left by the layer group beside the one-hot operand (the value column and the
            // padding columns read as illegal).  The accumulator row stays in tensor memory and each pass below re-reads it unit by
            // unit: a register copy of the row would have to be indexed by the (runtime) unit number, i.e. live in local memory, and the
            // L1 left beside 225 KB of shared memory does not hold 128 such rows — measured, that version spent 58k cycles per tile here.
            const uint4 *vrow = reinterpret_cast<const uint4 *>(vmask0 + ((size_t)buf * TILE_M + row) * VGROUPS);
            const uint4 va = vrow[0], vc = vrow[1];
            auto vbits = [&](int u) {
                const uint32_t wsel = (u >> 1) == 0 ? va.x : (u >> 1) == 1 ? va.y : (u >> 1) == 2 ? va.z : (u >> 1) == 3 ? va.w
                                    : (u >> 1) == 4 ? vc.x : (u >> 1) == 5 ? vc.y : (u >> 1) == 6 ? vc.z : vc.w;
                const uint32_t bits = (u & 1) ? wsel >> 16 : wsel & 0xFFFFu;
                // cells past A (value column, padding) are never legal; groups the one-hot stage did not cover hold stale bits
                const int rem = A - u * 16;
                return rem >= 16 ? bits : (rem <= 0 ? 0u : bits & ((1u << rem) - 1u));
            };
            // raw head outputs of columns [16u, 16u+16); the NEXT unit's accumulators travel from tensor memory meanwhile
            uint32_t nxt[16];
            auto unit = [&](int u, float (&y)[16]) {
                tmem_wait_ld();
                uint32_t acc[16];
#pragma unroll
                for (int j = 0; j < 16; j++) acc[j] = nxt[j];
                tmem_ld16(tmem + lane_base + (u + 1 < nu ? u + 1 : 0) * 16, nxt);       // (wraps to unit 0: the next pass starts there)
#pragma unroll
                for (int j = 0; j < 16; j++) y[j] = __uint_as_float(acc[j]) + bh[u * 16 + j];
            };
            tmem_ld16(tmem + lane_base, nxt);
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
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (!last) mbar_arrive(smem_u32(oh_ready));                    // z was needed until now
                mbar_arrive(smem_u32(board_free + buf));
            }
            TCK(6);
        }
        if (p.prof && lane == 0 && quad == 0) {
            atomicAdd(p.prof + 25, (unsigned long long)pc[0]);
            for (int k = 2; k < 7; k++) atomicAdd(p.prof + 24 + k, (unsigned long long)pc[k]);     // slots 26..30
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == WARP_MMA) tmem_dealloc(tmem, 512);
}

int launch(const bl_fc_params *p, TcParams &k, int B, cudaStream_t st) {
    const int S = p->S, A = S * S, W = p->W;
    k.blob = reinterpret_cast<const uint8_t *>(p->packed);
    k.cbias = p->b_head;
    k.prof = bl_phase_prof();
    k.B = B; k.S = S; k.A = A; k.W = W; k.D = p->D; k.precision = p->precision;
    k.K0p = (2 * A + KC - 1) / KC * KC;
    k.Np = (A + 1 + 31) / 32 * 32;
    k.nsplit = 1;
    const Shape sh = make_shape(W, k.K0p, k.Np);
    int ns = MAX_STAGES;
    while (ns >= 2 && smem_bytes(sh, ns, A, W, p->D, k.Np) > 227 * 1024) ns--;
    if (ns < 2) return -2;
    k.nstages = ns;
    const size_t smem = smem_bytes(sh, ns, A, W, p->D, k.Np);
    cudaError_t e = cudaFuncSetAttribute(fc_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int ntiles = (B + TILE_M - 1) / TILE_M;
    const int grid = ntiles < BL_NUM_SMS ? ntiles : BL_NUM_SMS;
    fc_tc_kernel<<<grid, TC_THREADS, smem, st>>>(k);
    return (int)cudaGetLastError();
}

}  // namespace

// Host-side launcher, called from bl_fc_forward (net.cu) when a packed operand blob is present and the shape fits.
int bl_fc_forward_tc(const bl_fc_params *p, const uint8_t *board, long long board_pitch, const int32_t *seats, float *logits,
                     float *v, int B, cudaStream_t st) {
    TcParams k = {};
    k.board = board; k.seats = seats; k.board_pitch = board_pitch;
    k.logits = logits; k.v = v;
    k.tree_mode = 0;
    return launch(p, k, B, st);
}

// Leaf evaluation straight from / into the search tree (bl_tree_eval_leaves): no gather, no fp32 logits round trip, no set_eval.
int bl_fc_forward_tc_tree(const bl_fc_params *p, const bl_tree *t, cudaStream_t st) {
    TcParams k = {};
    k.tree_mode = 1;
    k.tree = *t;
    return launch(p, k, t->B, st);
}

// shapes the tensor-core path covers: W in {32, 64, 128, 256} (z fits 256 TMEM columns, the operand regions 2 x 128), S <= 13
bool bl_fc_tc_supported(const bl_fc_params *p) {
    const int A = p->S * p->S, W = p->W;
    const int K0p = (2 * A + KC - 1) / KC * KC;
    return p->packed != nullptr && p->b_head != nullptr && (W == 32 || W == 64 || W == 128 || W == 256) && A + 1 <= 256 && K0p / 2 <= 256;
}
