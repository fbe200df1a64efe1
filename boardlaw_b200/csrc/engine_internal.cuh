// Internal interfaces between the engine translation units (not part of the C ABI).
#pragma once
#include "common.cuh"

// bl_tree.counters slots
enum { C_EVALS = 0, C_CHILDREN = 1, C_ITERS = 2, C_DESCENTS = 3, C_BACKUP_NODES = 4, C_ERRORS = 5, C_MOVE = 6, C_QUEUE = 7,
       // certified fast descent (descend_fx.cu): evaluations sent to the exact path because the stop test / the sampled action / a
       // guard (doubt too large, tiny values) could not be certified, and exact passes run there
       C_FLAG_STOP = 8, C_FLAG_SAMPLE = 9, C_FLAG_OTHER = 10, C_EXACT_PASSES = 11 };

static_assert(sizeof(bl_node) == 16 && sizeof(bl_aux) == 16, "tree records are 16 bytes");

// 128-bit record loads/stores (records are 16-byte aligned: the arrays come from the caller's allocator, >= 256 B aligned)
__device__ __forceinline__ bl_node bl_ld_node(const bl_node *p) {
    union { uint4 u; bl_node n; } x;
    x.u = *reinterpret_cast<const uint4 *>(p);
    return x.n;
}
__device__ __forceinline__ void bl_st_node(bl_node *p, const bl_node &n) {
    union { uint4 u; bl_node n; } x;
    x.n = n;
    *reinterpret_cast<uint4 *>(p) = x.u;
}
__device__ __forceinline__ bl_aux bl_ld_aux(const bl_aux *p) {
    union { uint4 u; bl_aux a; } x;
    x.u = *reinterpret_cast<const uint4 *>(p);
    return x.a;
}
// the statistics half of a node record (n, w, seat, terminal) as one 8-byte store
__device__ __forceinline__ void bl_st_node_stats(bl_node *p, const bl_node &n) {
    union { uint4 u; bl_node n; } x;
    x.n = n;
    reinterpret_cast<uint2 *>(p)[1] = make_uint2(x.u.z, x.u.w);
}
// L2 residency: the 32 bytes of node/aux records per (env, node) are re-read on every simulation and fit in the 126 MB L2
// (c2: 64 MB), the pi rows and boards stream through it; record accesses carry an evict_last policy.
__device__ __forceinline__ uint64_t bl_policy_keep() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint4 bl_ld16_hint(const void *p, uint64_t pol) {
    uint4 v;
    asm("ld.global.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
    return v;
}
// read-only use (descent): the records are not written by the kernel that calls these
__device__ __forceinline__ bl_node bl_ld_node_hint(const bl_node *p, uint64_t pol) {
    union { uint4 u; bl_node n; } x;
    x.u = bl_ld16_hint(p, pol);
    return x.n;
}
__device__ __forceinline__ bl_aux bl_ld_aux_hint(const bl_aux *p, uint64_t pol) {
    union { uint4 u; bl_aux a; } x;
    x.u = bl_ld16_hint(p, pol);
    return x.a;
}
__device__ __forceinline__ float bl_minnz(const bl_aux &a) { return __uint_as_float((uint32_t)a.minnz_hi << 16); }

// descend.cu: task-parallel descent (writes t.leaf = existing terminal child or -1, t.leaf_parent, t.leaf_action)
// followed by expand + env step.  Returns a cudaError_t / negative argument error.
int bl_descend_v3(const bl_tree *t, int sim, const bl_half *rands, uint64_t seed, cudaStream_t st);
// descend_mw.cu: the same descent with four lanes per env (4x the warps); -3 = scratch too small, -2 = unsupported
int bl_descend_mw(const bl_tree *t, int sim, const bl_half *rands, uint64_t seed, cudaStream_t st);
int bl_mw_child_cap(const bl_tree *t);
// descend_pc.cu: the same descent with passes and services on different warps of a CTA (variant 4); same return codes
int bl_descend_pc(const bl_tree *t, int sim, const bl_half *rands, uint64_t seed, cudaStream_t st);
int64_t bl_mw_scratch_bytes(const bl_tree *t);
// descend_fx.cu: certified fast descent (variant 5): closed-form sums over the children, decisions certified against an error
// bound, exact path otherwise; followed by expand + env step.  -2 = unsupported shape
int bl_descend_fx(const bl_tree *t, int sim, const bl_half *rands, uint64_t seed, cudaStream_t st);
int64_t bl_fx_scratch_bytes(const bl_tree *t);
// descend_all.cu: speculative descent (variant 6): every node evaluated independently (certified closed form), then a pointer chase
// + expand + env step.  -2 = unsupported shape, -3 = scratch too small
int bl_descend_all(const bl_tree *t, int sim, const bl_half *rands, uint64_t seed, cudaStream_t st);
int64_t bl_all_scratch_bytes(const bl_tree *t);
// descend_pk.cu: packed descent (variant 7): one CTA per SM, pass warps whose lanes claim ready envs + service warps that visit a
// node warp-cooperatively; followed by expand + env step.  -2 = unsupported shape (A > 84 or T > 64)
int bl_descend_pk(const bl_tree *t, int sim, const bl_half *rands, uint64_t seed, cudaStream_t st);
bool bl_experimental_built();      // variants 4, 6 and 7 are compiled in (BL_EXPERIMENTAL=1 at build time)
// descend.cu: expand + env step of the descents recorded in t.leaf / leaf_parent / leaf_action
int bl_expand_step(const bl_tree *t, int sim, cudaStream_t st);
// descend.cu: device buffer of the optional phase clock (NULL = off); slots 0..15 descent, 16..31 network
unsigned long long *bl_phase_prof();
