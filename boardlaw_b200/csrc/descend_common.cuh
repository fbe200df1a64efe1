// Shared pieces of the descent kernels (descend.cu: one lane per env; descend_mw.cu: four lanes per env): states, child
// entries, packed fp32 helpers and the expand + env step of one env.
#pragma once
#include "engine_internal.cuh"
#include "hex_core.cuh"
#include "mcts_core.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;
#define BL_TINY 7.888609052210118e-31f   /* 2^-100 */

enum { ST_IDLE = 0, ST_VISIT = 1, ST_PASS = 2, ST_FINAL = 3, ST_SAMPLE = 4, ST_SLOW = 5, ST_ADVANCE = 6, ST_DONE = 7 };

struct __align__(16) ChildEntry { float q, top; int a, id, flags; };   // flags: seat | terminal << 8 of the child (shared-memory form packs a|id and flags)

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { return ((u64)__float_as_uint(b) << 32) | __float_as_uint(a); }
__device__ __forceinline__ float lo(u64 v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float hi(u64 v) { return __uint_as_float((unsigned)(v >> 32)); }
// packed fp32 pipe operations: IEEE round-to-nearest per half, denormals kept (tools/ubench_fp32x2.cu checks them against the
// scalar intrinsics on 1.8e10 operand triples)
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- expand + env step (boardlaw/mcts/__init__.py:117-129), one lane per env ---------------------------------------------------
// The lane moves its parent's board (BP bytes, 16-byte loads) into a lane-private shared-memory row `bdw`, places the stone /
// relabels the group there (`stk`: flood-fill stack), and writes the row to the leaf's slot with 16-byte stores.
__device__ __forceinline__ void bl_expand_one(const bl_tree &t, int sim, int b, int leaf, int parent, int action, uint32_t *bdw, uint8_t *stk) {
    const int T = t.T, nq = t.BP >> 4;
    const size_t node0 = (size_t)b * T;
    if (action < 0) {
        t.leaf[b] = -1;
        atomicAdd(reinterpret_cast<unsigned long long *>(t.counters + C_ERRORS), 1ull);
        return;
    }
    uint8_t *bd = reinterpret_cast<uint8_t *>(bdw);
    const uint4 *src = reinterpret_cast<const uint4 *>(t.board + (node0 + parent) * t.BP);
    for (int i = 0; i < nq; i++) {
        const uint4 v = src[i];
        bdw[4 * i] = v.x; bdw[4 * i + 1] = v.y; bdw[4 * i + 2] = v.z; bdw[4 * i + 3] = v.w;
    }
    const bl_node pn = bl_ld_node(t.node + node0 + parent);
    bl_node ln;
    bool fresh = false;
    if (leaf < 0) {                                             // new node in slot `sim`
        leaf = sim;
        fresh = true;
        ln.parent = (int16_t)parent; ln.relation = (int16_t)action; ln.first_child = -1; ln.next_sib = pn.first_child;
        ln.n = 0; ln.w[0] = 0; ln.w[1] = 0;
        t.node[node0 + parent].first_child = (int16_t)sim;
        t.parent_of[(size_t)b * ((T + 7) & ~7) + sim] = (int16_t)parent;
        t.kids[(node0 + parent) * ((T + 63) >> 6) + (sim >> 6)] |= 1ull << (sim & 63);
        if (t.cprior) t.cprior[node0 + sim] = t.pi[(node0 + parent) * t.AP + action];
        t.leaf[b] = (int16_t)leaf;
    } else {                                                    // stopped at an existing terminal child: reuse its slot
        ln = bl_ld_node(t.node + node0 + leaf);
    }
    const int seat = pn.seat;
    const int win = bl_hex_place<uint8_t>(bd, stk, 1, t.S, seat, action);
    const float r0 = win == 1 ? 1.f : (win == 2 ? -1.f : 0.f), r1 = win == 1 ? -1.f : (win == 2 ? 1.f : 0.f);   // +0, never -0
    reinterpret_cast<uint32_t *>(t.aux + node0 + leaf)[0] = (uint32_t)bl_f2h(r0) | ((uint32_t)bl_f2h(r1) << 16);
    ln.terminal = (uint8_t)win;                                 // 0, or the winner's code (1 = seat 0, 2 = seat 1): the backup
    ln.seat = win ? 0 : (uint8_t)(1 - seat);                    // reads the rewards (+-1) off it
    if (fresh) bl_st_node(t.node + node0 + leaf, ln);
    else bl_st_node_stats(t.node + node0 + leaf, ln);
    uint4 *dst = reinterpret_cast<uint4 *>(t.board + (node0 + leaf) * t.BP);
    for (int i = 0; i < nq; i++)                                // a won game auto-resets to the empty board (hex/__init__.py:185-188)
        dst[i] = win ? make_uint4(0u, 0u, 0u, 0u) : make_uint4(bdw[4 * i], bdw[4 * i + 1], bdw[4 * i + 2], bdw[4 * i + 3]);
}

}  // namespace
