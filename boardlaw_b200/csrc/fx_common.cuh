// Shared pieces of the certified fast descent kernels (descend_fx.cu: lock-step descents, one lane per env; descend_all.cu: every
// node of every tree evaluated independently, then a pointer chase): constants of the error model, shared-memory helpers, the
// child entry.  The error model itself is documented in descend_fx.cu and specified by oracle/filter_model.py.
#pragma once
#include "descend_common.cuh"

namespace {

constexpr float FX_U = 5.9604644775390625e-08f;        // 2^-24
constexpr float FX_GUARD = 0.00390625f;                // e <= 2^-8 (alpha - q_max): functions of alpha vary < 2 % across the doubt
constexpr int FX_MAXIT = 24;                           // fast-path Newton passes before the evaluation is handed to the exact path
constexpr int FX_KS = 17;                              // child entries per lane in shared memory (odd: conflict-free 16-byte lane rows)

// evaluation modes of a lane
enum { FX_IDLE = 0, FX_VISIT = 1, FX_ITER = 2, FX_SAMPLE = 3, FX_XPASS = 4, FX_XALL = 5, FX_EVALDONE = 6 };

// shared memory by 32-bit address
__device__ __forceinline__ float fx_lds(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void fx_sts(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ uint4 fx_lds16(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void fx_sts16(uint32_t a, const uint4 &v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void fx_cp16(uint32_t dst, const void *src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void fx_cp4(uint32_t dst, const void *src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ uint32_t fx_opaque(uint32_t x) { uint32_t y; asm volatile("mov.u32 %0, %1;" : "=r"(y) : "r"(x)); return y; }

// child entry (16 bytes): t = RN(lambda*pi_c) (pi_c until lambda is known), q, w, m = seat << 14 | terminal << 15 | action << 16 |
// node id << 24 (T <= 256); the sampling correction dlt = s_c - t_c/alpha lives in the lane's cpr zone (free after the visit)
struct FxEntry { float t, q, w; uint32_t m; };            // w = L1 - action (as float): the term's addition-count weight
__device__ __forceinline__ int fx_a(uint32_t m) { return (int)((m >> 16) & 255u); }


}  // namespace
