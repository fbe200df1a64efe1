// Shared device helpers for the boardlaw_b200 kernels (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/boardlaw_b200.h"

#define BL_NUM_SMS 148
#define BL_INF_F __int_as_float(0x7f800000)

#define BL_LAUNCH_CHECK() return (int)cudaGetLastError()

static inline cudaStream_t bl_cu(bl_stream s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- binary16 <-> fp32, round-to-nearest-even, as c10::Half (reference: boardlaw/cpp/common.h typedefs) ----
__device__ __forceinline__ float bl_h2f(bl_half h) { return __half2float(__ushort_as_half(h)); }
__device__ __forceinline__ bl_half bl_f2h(float f) { return __half_as_ushort(__float2half_rn(f)); }

// ---- order-preserving float <-> int encoding for atomicMin/atomicMax ----
__device__ __forceinline__ int bl_f2ord(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float bl_ord2f(int k) {
    return __int_as_float(k >= 0 ? k : k ^ 0x7FFFFFFF);
}

// ---- Philox4x32-10 (Salmon et al. 2011), counter-based; used for the in-kernel random stream ----
struct bl_philox_out { uint32_t x, y, z, w; };
__device__ __forceinline__ bl_philox_out bl_philox(uint64_t key, uint64_t ctr_lo, uint64_t ctr_hi) {
    uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
    uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return {c0, c1, c2, c3};
}
// uniform on the binary16 grid {k/2048 : k = 0..2047}, the support of torch.rand(dtype=half)
__device__ __forceinline__ float bl_uniform_half_grid(uint32_t bits) { return (float)(bits >> 21) * (1.0f / 2048.0f); }

__device__ __forceinline__ int bl_lane() { return threadIdx.x & 31; }
