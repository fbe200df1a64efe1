/*
 * boardlaw_b200 — C ABI of the B200-native (sm_100a) Hex / MCTS / policy-value-net hot path.
 *
 * This is the drop-in boundary.  Every entry point is `extern "C"`, takes plain device
 * pointers, sizes and a CUDA stream (as void*), never allocates or frees device memory,
 * never synchronises, and returns the cudaError_t of its last launch (0 = success,
 * negative = argument error detected on the host).  All launches are CUDA-graph capturable.
 *
 * Each function cites the reference interface (andyljones/boardlaw) it replaces.  The
 * reference binds these through pybind11 torch extensions; the ctypes binding a maintainer
 * would add is shown in INTEGRATION.md and implemented in boardlaw_b200/_lib.py.
 *
 * Layout conventions (all tensors contiguous, row-major):
 *   B = n_envs, T = n_nodes, S = boardsize, A = S*S actions, Sn = n_seats (2 for Hex)
 *   "half" arguments are IEEE binary16 bit patterns (uint16_t here, __half on the device).
 */
#ifndef BOARDLAW_B200_H
#define BOARDLAW_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint16_t bl_half;
typedef void *bl_stream;

/* ABI version, bumped on any signature change. */
int bl_abi_version(void);

/* Selects the CUDA device for subsequent launches from the calling thread (the library links the CUDA
 * runtime statically; replaces the CUDAGuard of boardlaw/hex/cpp/cuda.cu:140, boardlaw/mcts/cpp/cuda.cu:121). */
int bl_set_device(int device);

/* Fills out[65536] (HOST memory) with expf() of every binary16 bit pattern as evaluated by
 * the host libm — the table the descend kernels index so that exp(logit) is identical to the
 * value the reference's CPU build obtains from libm (boardlaw/mcts/cpp/cpu.cpp:89).  The
 * caller uploads it once per device and passes it as `exp_lut`. */
void bl_exp_table_host(float *out);

/* ---------------------------------------------------------------------------------------------
 * Hex — replaces hexcuda.step / hexcuda.observe (boardlaw/hex/cpp/wrappers.cpp:39-40,
 * kernels boardlaw/hex/cpp/cuda.cu:76-152 and :154-217).
 * ------------------------------------------------------------------------------------------- */

/* In place on board (B,S,S) u8.  rewards (B,2) f32 is fully written (zeros unless a win). */
int bl_hex_step(uint8_t *board, const int32_t *seats, const int32_t *actions, float *rewards,
                int B, int S, bl_stream stream);

/* obs (B,S,S,2) f32, fully written.  B is the product of all leading batch dims. */
int bl_hex_observe(const uint8_t *board, const int32_t *seats, float *obs, int B, int S,
                   bl_stream stream);

/* Fused env transition = Hex.step of boardlaw/hex/__init__.py:161-195 without the host syncs:
 * copies board->new_board, applies the move, auto-resets finished games (board and seat zeroed),
 * flips the seat, and records rule violations (negative / out-of-range / occupied cell) by
 * OR-ing bits into *error_word instead of asserting on the host.
 * terminal (B,) u8 (bool), rewards (B,2) f32.  reset!=0 mirrors `reset=True`. */
int bl_hex_transition(const uint8_t *board, const int32_t *seats, const int64_t *actions,
                      uint8_t *new_board, int32_t *new_seats, float *rewards, uint8_t *terminal,
                      int32_t *error_word, int reset, int B, int S, bl_stream stream);

/* One step of a uniformly random playout, fused: replaces the loop body of learning.mix
 * (boardlaw/learning.py:6-10: `Categorical(probs=worlds.valid.float()).sample()` then `worlds.step(actions)`)
 * and the moves of validation.RandomAgent.  uniforms (B,) f32 in [0,1) are drawn by the caller (the reference draws
 * inside torch.distributions); env b plays its k-th legal move in the mover's frame order, k = min(floor(u_b * n_legal),
 * n_legal - 1), which is written to actions (B,) i64; everything else as bl_hex_transition. */
int bl_hex_random_transition(const uint8_t *board, const int32_t *seats, const float *uniforms,
                             uint8_t *new_board, int32_t *new_seats, int64_t *actions, float *rewards,
                             uint8_t *terminal, int32_t *error_word, int reset, int B, int S,
                             bl_stream stream);

/* valid (B,A) u8 (bool) in the mover's frame = (obs == 0).all(-1) of boardlaw/hex/__init__.py:154-159,
 * computed from the board without materialising obs. */
int bl_hex_valid(const uint8_t *board, const int32_t *seats, uint8_t *valid, int B, int S,
                 bl_stream stream);

/* ---------------------------------------------------------------------------------------------
 * MCTS ops on the reference's tensor layout — replace mctscuda.descend / root / backup
 * (boardlaw/mcts/cpp/wrappers.cpp:52-73; kernels boardlaw/mcts/cpp/cuda.cu:101-248).
 *   logits (B,T,A) half, w (B,T,Sn) half, n (B,T) i16, c_puct (B,) half, seats (B,T) i16,
 *   terminal (B,T) bool, children (B,T,A) i16.
 * `qrange` is 2 floats of device scratch that receive (min, max) of w/(n+1e-4) over the whole
 * batch (transition_q, cuda.cu:101-105).  `exp_lut` is the uploaded bl_exp_table_host table.
 * `counters` (optional, may be NULL) is 4 x uint64 on the device, incremented by: policy
 * evaluations, existing children seen, Newton iterations, descents.
 * ------------------------------------------------------------------------------------------- */

/* rands (B,T) half is the tensor the reference draws with at::rand_like inside descend
 * (cuda.cu:191); here the caller draws it.  parents, actions (B,) i16 out. */
int bl_mcts_descend(const bl_half *logits, const bl_half *w, const int16_t *n, const bl_half *c_puct,
                    const int16_t *seats, const uint8_t *terminal, const int16_t *children,
                    const bl_half *rands, const float *exp_lut, float *qrange,
                    int16_t *parents, int16_t *actions, uint64_t *counters,
                    int B, int T, int A, int Sn, bl_stream stream);

/* probs (B,A) half out: the regularised policy at node 0 (cuda.cu:107-136). */
int bl_mcts_root(const bl_half *logits, const bl_half *w, const int16_t *n, const bl_half *c_puct,
                 const int16_t *seats, const uint8_t *terminal, const int16_t *children,
                 const float *exp_lut, float *qrange, bl_half *probs,
                 int B, int T, int A, int Sn, bl_stream stream);

/* In place on n (B,T) i16 and w (B,T,Sn) half (cuda.cu:205-248).  leaves (B,) i16. */
int bl_mcts_backup(const bl_half *v, bl_half *w, int16_t *n, const bl_half *rewards,
                   const int16_t *parents, const uint8_t *terminal, const int16_t *leaves,
                   int B, int T, int Sn, bl_stream stream);

/* transition_q alone (cuda.cu:101-105): q (B,T,Sn) half out, qrange as above. */
int bl_mcts_transition_q(const bl_half *w, const int16_t *n, bl_half *q, float *qrange,
                         int B, int T, int Sn, bl_stream stream);

/* ---------------------------------------------------------------------------------------------
 * Policy/value network — replaces FCModel.forward (boardlaw/networks.py:37-41) with the heads of
 * boardlaw/heads.py:41-52 (TensorIntake), :93-104 (MaskedOutput), :128-142 (ValueOutput) and the
 * ReZero residual blocks of boardlaw/networks.py:10-18.  The observation is never materialised:
 * the kernels read the board and the seat.
 *
 * Weights are the fp32 tensors of the reference's state_dict (body.0.*, body.k.*, policy.core.*,
 * value.core.*), contiguous, on the device.
 * ------------------------------------------------------------------------------------------- */

typedef struct bl_fc_params {
    int S;            /* boardsize                                                             */
    int W;            /* width                                                                 */
    int D;            /* number of ReZero residual blocks                                      */
    int precision;    /* 0: fp32-accurate (exact fp32 products or split-fp16 tensor-core products,
                            fp32 accumulation); 1: fp16 operands / fp32 accumulation (mirrors the
                            reference's autocast, boardlaw/mcts/__init__.py:131-133)               */
    const float *w_in;    /* (W, 2A)  body.0.weight       */
    const float *b_in;    /* (W,)     body.0.bias         */
    const float *w_res;   /* (D, W, W) body.k.weight      */
    const float *b_res;   /* (D, W)   body.k.bias         */
    const float *alpha;   /* (D,)     body.k.α            */
    const float *w_pol;   /* (A, W)   policy.core.weight  */
    const float *b_pol;   /* (A,)     policy.core.bias    */
    const float *w_val;   /* (W,)     value.core.weight   */
    const float *b_val;   /* (1,)     value.core.bias     */
    const void *packed;   /* tensor-core operand tiles (split-fp16, UMMA canonical K-major layout, ReZero gate folded in)
                             in the kernel's consumption order, built by the host packer
                             boardlaw_b200/networks.py:pack_tensor_core_operands; NULL: CUDA-core path              */
    const float *b_head;  /* with `packed`: (D+1, W) cumulative biases c_0 = b_in, c_k = c_{k-1} + alpha_k b_k, then
                             (roundup(A+1,32),) = [policy bias (A), value bias, zeros]                              */
    int tc_nsplit;        /* how `packed` orders a layer's tiles: 1 = whole-N tiles in K order, 2 = the four (N half, K half)
                             blocks of net_tc.cu (lets the epilogue overlap the MMAs; N = W/2 per instruction)      */
} bl_fc_params;

/* 1 when bl_fc_forward / bl_tree_eval_leaves run this shape on the tcgen05 kernel (W in {32,64,128,256}, packed operands
 * present), 0 when they run the CUDA-core kernels. */
int bl_fc_uses_tensor_cores(const bl_fc_params *p);

/* Bytes of device scratch bl_fc_forward needs for a batch of B envs. */
int64_t bl_fc_scratch_bytes(const bl_fc_params *p, int B);

/* board (B,S,S) u8, seats (B,) i32 -> logits (B,A) f32 (masked log-softmax, -inf on occupied cells),
 * v (B,2) f32 with v[seat]=tanh(.), v[1-seat]=-tanh(.). */
int bl_fc_forward(const bl_fc_params *p, const uint8_t *board, const int32_t *seats,
                  float *logits, float *v, void *scratch, int B, bl_stream stream);

/* ---------------------------------------------------------------------------------------------
 * Fused search engine — the whole of MCTS.initialize / simulate / root
 * (boardlaw/mcts/__init__.py:72-149) on a persistent, privately laid out workspace.
 * The workspace arrays are allocated by the caller (torch) and described by bl_tree.
 * ------------------------------------------------------------------------------------------- */

/* One record per (env, node); 16 bytes so that a node visit is a single 128-bit load. */
typedef struct bl_node {
    int16_t parent;       /* -1 = none                                                              */
    int16_t relation;     /* action that led here                                                   */
    int16_t first_child;  /* head of the child list, -1 = none                                      */
    int16_t next_sib;     /* next sibling                                                           */
    int16_t n;            /* visit count (incremented Sn per visit, as the reference)               */
    bl_half w[2];         /* accumulated value per seat                                             */
    uint8_t seat;         /* seat to move                                                           */
    uint8_t terminal;     /* 0, or the code of the seat whose move ended the game here (1 = seat 0, 2 = seat 1)  */
} bl_node;

/* Second 16-byte record per (env, node): what only the backup and the row load need. */
typedef struct bl_aux {
    bl_half rewards[2];
    bl_half v[2];
    float max_pi;         /* max of the node's pi row (seeds alpha: max_a RN(lambda*pi_a) = RN(lambda*max_pi)) */
    uint16_t minnz_hi;    /* upper 16 bits of the smallest nonzero pi (truncated, i.e. conservative)  */
    uint8_t first_nz;     /* first / last action with pi != 0                                        */
    uint8_t last_nz;
} bl_aux;

typedef struct bl_tree {
    int B, T, S, A, Sn;
    int AP;               /* row pitch of pi in floats (A rounded up to a multiple of 4)            */
    int BP;               /* row pitch of board in bytes (A rounded up to a multiple of 16)         */
    float *pi;            /* (B,T,AP) exp(logits) as fp32, value of exp_lut[half(logit)]; pad = 0    */
    float *cpi;           /* NULL, or (with psum and cprior; needed by descent variant 5 only) (B,T,AP) inclusive prefix sums of the pi row (accumulated in double, rounded once per entry; pad = total):
                             what the certified fast descent samples from (binary search) and takes the row's mass from              */
    float *psum;          /* (B,T)  sum_a a*pi[a] per node (fp32; feeds only the descent's error bound)                              */
    float *cprior;        /* (B,T)  pi[parent][relation] of each node: its prior under its parent, copied by the expand step so that a
                             visit needs no access to the parent's pi row for its children                                           */
    bl_half *logits;      /* (B,T,A) half or NULL: reference-layout mirror (kept only in debug mode) */
    uint8_t *board;       /* (B,T,BP) u8 absolute-frame boards                                      */
    bl_node *node;        /* (B,T)                                                                  */
    bl_aux *aux;          /* (B,T)                                                                  */
    int16_t *parent_of;   /* (B,TP) i16, TP = T rounded up to a multiple of 8: parent of every node (-1 = none) as one
                             contiguous row per env (trees deeper than the kids masks cover walk the sibling lists)   */
    uint64_t *kids;       /* (B,T,KW) u64, KW = ceil(T/64): bit k of a node's mask = node k is its child; set by the expand step,
                             read by the descent's visit (one 8*KW-byte fetch instead of a scan of parent_of)           */
    bl_half *c_puct;      /* (B,)   half                                                            */
    int16_t *leaf;        /* (B,)   i16 leaf of the current simulation                              */
    int16_t *leaf_parent; /* (B,)   i16                                                             */
    int16_t *leaf_action; /* (B,)   i16                                                             */
    bl_half *leaf_v;      /* (B,Sn) half: the value just evaluated at each env's leaf (copy of aux[leaf].v, so the backup does not
                             chase leaf -> aux)                                                                       */
    bl_half *prior;       /* (B,A)  half: the (noised) root logits as stored, = decisions.logits[:,0]       */
    float *qrange;        /* (T+1,2) per-simulation (min,max) of w/(n+1e-4), ordered-int encoded    */
    uint64_t *counters;   /* (16,) policy evals, children seen, newton iters, descents, backup nodes, errors, move, queue; certified fast
                             descent: [8] evaluations whose stop test, [9] sampled action, [10] guards could not be certified (exact path),
                             [11] exact passes run */
    const float *exp_lut; /* (65536,)                                                               */
    void *scratch;        /* device scratch for the descent's per-lane child lists                          */
    int64_t scratch_bytes;/* >= bl_tree_scratch_bytes(t)                                                    */
} bl_tree;

/* Bytes of `scratch` the descent needs for a tree of this shape. */
int64_t bl_tree_scratch_bytes(const bl_tree *t);

/* Resets the workspace for a new search rooted at (board (B,S,S) u8, seats (B,) i32).  The in-kernel random
 * stream is keyed by (seed, counters[6]); the reset increments counters[6], the move index, itself (no host write per move). */
int bl_tree_reset(const bl_tree *t, const uint8_t *board, const int32_t *seats, float c_puct,
                  bl_stream stream);

/* Writes the network evaluation of node `node` for every env: logits (B,A) and v (B,2), f32 or half
 * (inputs_are_half), rounded to half exactly as `decisions.half()` does (boardlaw/mcts/__init__.py:135-136)
 * and stored as pi = exp_lut[half(logit)].  node < 0 means "the current leaf of each env" (t->leaf). */
int bl_tree_set_eval(const bl_tree *t, int node, const void *logits, const void *v,
                     int inputs_are_half, bl_stream stream);

/* MCTS.initialize's second half in one launch (boardlaw/mcts/__init__.py:72-80 with dirichlet_noise, :13-24): the root evaluation
 * (logits f32 (B,A) as bl_tree_eval_root returns them, v f32 (B,2)) is mixed in probability space with Dirichlet(alpha_scale/A) noise
 * restricted to the legal moves of (board (B,S,S) u8, seats (B,) i32) and renormalised, log(exp(l)(1-eps) + d eps), and stored as node
 * 0's evaluation exactly as bl_tree_set_eval(node=0) would.  draw: (B,A) f32 raw Dirichlet sample to use (the value
 * torch.distributions.Dirichlet.sample returns, before masking), or NULL to draw the gammas in-kernel (Marsaglia-Tsang on Philox
 * keyed by (seed, move counter, env, action)). */
int bl_tree_set_root_prior(const bl_tree *t, const float *logits, const float *v, const uint8_t *board, const int32_t *seats,
                           const float *draw, float noise_eps, float alpha_scale, uint64_t seed, bl_stream stream);

/* descend + expand + env step of simulation `sim` (boardlaw/mcts/__init__.py:108-129): writes
 * t->leaf / leaf_parent / leaf_action, links new nodes, steps the parent's board into the leaf slot, records
 * rewards / terminal.  rands: (B,T) half injected random numbers (indexed by node, as cuda.cu:158), or NULL to
 * draw them in-kernel from Philox4x32-10 keyed by (seed, move counter) and counted by (env, sim, node). */
int bl_tree_descend_expand(const bl_tree *t, int sim, const bl_half *rands, uint64_t seed,
                           bl_stream stream);

/* Selects the descent kernel: 0 (default) = by board size (2 up to 9x9, 3 above: measured on c2/c3/c5),
 * 5 = certified fast descent (descend_fx.cu; needs t->cpi / psum / cprior, A <= 255, T <= 256): closed-form sums over the children
 *     only, every decision certified against a rigorous bound on |reference - ours| and recomputed with the reference's arithmetic
 *     when the bound does not separate it (bit-identical results, fewer instructions, but not faster: DESIGN.md 5.1c),
 * 2 = task-parallel descent, one lane per env with register-resident rows (descend.cu),
 * 3 = two or four lanes per env: terms split over the lanes, the S and g chains on two of them (descend_mw.cu),
 * 6 = speculative descent (descend_all.cu; same requirements as 5): the certified evaluation of EVERY node of every tree, independently
 *     (a node's sampled action does not depend on how it was reached), then a pointer chase from the root,
 * 4 = experimental: passes and node services on different warps of a CTA (descend_pc.cu; measured slower),
 * 7 = experimental: one CTA per SM, pass lanes that claim ready envs + warp-cooperative node visits (descend_pk.cu; A <= 84, T <= 64;
 *     measured slower),
 * 1 = one lane per env in lock step, reference loops verbatim (engine.cu; kept as an on-device cross-check).
 * All produce identical results (tests/test_gpu_mcts.py runs the oracle comparison for each).  4, 6 and 7 were measured slower and are
 * compiled in only when the library is built with BL_EXPERIMENTAL=1; without it the call returns -2 for them. */
int bl_debug_set_descend_variant(int variant);

/* Caps the number of warps of the one-lane descent (variant 2): with fewer lanes than envs the lanes pull envs from a global
 * queue (the path large batches take when not every env fits on the device at once).  0 = no cap.  For tests and tuning. */
int bl_debug_set_descend_grid(int warps);


/* Phase clock of the descent and network kernels: when `buf` (32 x uint64 on the device, zeroed by the caller) is non-NULL every warp adds the
 * cycles it spent per phase — [0] loop head, [1] sample+advance, [2] finish/fetch, [3] visit, [4] child terms, [5] pass,
 * [6] Newton update/tail, [7..10] visit sub-phases — and [15] += 1; the network kernel's MMA thread adds [16] wait for operand,
 * [17] wait for weights, [18] issue, its first epilogue thread [20] board staging, [21] one-hot operand, [22] wait for accumulator,
 * [23] layer epilogue, [24] heads, [25] wait for heads — and [31] += 1 per CTA.  NULL (default) switches it off. */
int bl_debug_set_phase_profile(uint64_t *buf);

/* Self test: counts operand pairs for which the shared-reciprocal division of descend.cu differs from the IEEE
 * division (expected 0) over n_div x n_num pseudo-random pairs drawn from the descent's operand ranges. */
int bl_selftest_division(uint64_t seed, int n_div, int n_num, uint64_t *mismatch, bl_stream stream);

/* backup of the current leaves (boardlaw/mcts/cpp/cuda.cu:205-248) + the q-range scan used by the next
 * descent (transition_q's global min/max, cuda.cu:101-105), stored in qrange[sim+1]. */
int bl_tree_backup(const bl_tree *t, int sim, bl_stream stream);

/* Device scratch needed by bl_tree_eval_leaves / bl_tree_eval_root. */
int64_t bl_tree_eval_scratch_bytes(const bl_tree *t, const bl_fc_params *p);

/* Network evaluation of the current leaves straight into the tree. */
int bl_tree_eval_leaves(const bl_tree *t, const bl_fc_params *p, int sim, void *scratch, bl_stream stream);

/* Root evaluation: logits f32 (B,A) and v f32 (B,2) of node 0 are returned to the caller, who mixes the
 * Dirichlet noise (boardlaw/mcts/__init__.py:13-24) and hands the result to bl_tree_set_eval(node=0). */
int bl_tree_eval_root(const bl_tree *t, const bl_fc_params *p, float *logits, float *v, void *scratch,
                      bl_stream stream);

/* Regularised root policy -> log-probabilities as half (B,A) (MCTS.root, boardlaw/mcts/__init__.py:142-149),
 * v (B,2) half, n_leaves (B,) i64 (boardlaw/mcts/__init__.py:151-152).  `sim` = number of nodes evaluated so far.
 * log_lut[h] (65536 halves, device) = half(log(float(h))) as the caller's framework evaluates
 * `r.float().log().half()` (boardlaw/mcts/__init__.py:147). */
int bl_tree_root(const bl_tree *t, int sim, const bl_half *log_lut, bl_half *logits, bl_half *v,
                 int64_t *n_leaves, bl_stream stream);

/* bl_tree_root plus the agent's move (MCTSAgent.__call__, boardlaw/mcts/__init__.py:216-229): actions (B,) i64 = argmax of the root
 * policy when greedy != 0 (eval=True), else a draw from Categorical(logits) by inverse CDF over exp(half logits) with uniforms (B,)
 * f32 in [0,1) supplied by the caller, or drawn in-kernel from Philox keyed by (seed, move counter, env) when NULL. */
int bl_tree_root_act(const bl_tree *t, int sim, const bl_half *log_lut, bl_half *logits, bl_half *v, int64_t *n_leaves,
                     int64_t *actions, const float *uniforms, int greedy, uint64_t seed, bl_stream stream);

/* The trajectory record of one move for every env — what the actor appends to its buffer (boardlaw/main.py:179:
 * arrdict(worlds, decisions.half(), transitions)) — packed for the per-move all-gather: records (B,R) u8, R a multiple of 16 >= 5A+12:
 * board A u8 | seat u8 | terminal u8 | action i16 | rewards 2 x f16 | v 2 x f16 | logits A x f16 | prior A x f16 | zero padding.
 * board (B,A) u8 and seats (B,) i32 are the worlds the agent moved in; terminal (B,) u8, rewards (B,2) f32 the transition's. */
int bl_pack_records(const uint8_t *board, const int32_t *seats, const uint8_t *terminal, const int64_t *actions, const float *rewards,
                    const bl_half *v, const bl_half *logits, const bl_half *prior, uint8_t *records, int B, int A, int R, bl_stream stream);

/* Materialises the reference's dense children (B,T,A) i16 tensor from the child lists. */
int bl_tree_children_dense(const bl_tree *t, int16_t *children, bl_stream stream);

/* ---- learner side: the consumers of the self-play trajectories ------------------------------------------ */

/* learning.reward_to_go (boardlaw/learning.py:57-76, called from main.as_chunk, boardlaw/main.py:61-67) over one chunk:
 * reward, value (T,B,Sn) f32, terminal (T,B) u8 (bool; the reference stacks it over the seats), out (T,B,Sn) f32, or half when
 * out_is_half (as_chunk's `.half()`).  out[T-1] = terminal ? reward : value; out[t] = terminal[t] ? reward[t] :
 * reward[t] + gamma*out[t+1], fp32 with the product rounded before the sum: bit-identical to the reference's loop.
 * Unlike the reference it does not write the fallback into `value`. */
int bl_reward_to_go(const float *reward, const float *value, const uint8_t *terminal, void *out, int out_is_half,
                    int T, int B, int Sn, float gamma, bl_stream stream);

/* The loss of main.optimize (boardlaw/main.py:86-101) and its gradient in one pass: logp (N,A) f32 masked log-softmax output
 * of the network, v (N,2) f32, target_logits (N,A) half, target_v (N,2) half (reward-to-go), seats (N,) i32.
 * sums[0] += sum_n sum_a exp(l0)*l (policy_loss = -sums[0]/N), sums[1] += sum (target - v)^2 (value_loss = sums[1]/(2N));
 * dscores (N,A) f32 = d(policy_loss + value_loss)/d(pre-softmax scores), dz (N,) f32 = d/d(pre-tanh value). */
int bl_policy_value_loss(const float *logp, const float *v, const bl_half *target_logits, const bl_half *target_v,
                         const int32_t *seats, float *dscores, float *dz, float *sums, int N, int A, bl_stream stream);

/* torch.optim.Adam's update (boardlaw/main.py:154, defaults: no weight decay, no amsgrad) over one flat fp32 buffer of n
 * parameters; step counts from 1. */
int bl_adam_step(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, int64_t n, float lr, float beta1,
                 float beta2, float eps, int step, bl_stream stream);

/* The learner's dense contractions (forward, dgrad, wgrad of FCModel under main.optimize, boardlaw/main.py:75-101 — torch.nn.Linear
 * GEMMs in the reference) on the tcgen05 tensor cores with fp32 accuracy (split-fp16 products, fp32 accumulation in tensor memory):
 *     C[m*ldc + n] = sum_k a(m,k) * b(n,k) (+ bias[n]),   a(m,k) = A[m*a_rs + k*a_cs], b(n,k) = B[n*b_rs + k*b_cs]
 * for any strides (a transposed operand is a stride swap); a_relu / b_relu apply max(.,0) to the operand on the way in.
 * a_amax / b_amax: device scalars holding max|x| of each operand (NULL = values already within fp16's range): the operands are
 * scaled by exact powers of two into range before the split and the result is scaled back.  Small M*N with a long K (wgrad) is split
 * over K into partial products in `workspace` (bl_gemm_f32_workspace_bytes; without it the split is skipped) and summed in a fixed order. */
int64_t bl_gemm_f32_workspace_bytes(int M, int N, int K);
int bl_gemm_f32(const float *A, long long a_rs, long long a_cs, int a_relu, const float *a_amax, const float *B, long long b_rs,
                long long b_cs, int b_relu, const float *b_amax, const float *bias, float *C, long long ldc, int M, int N, int K,
                void *workspace, int64_t workspace_bytes, bl_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* BOARDLAW_B200_H */
