// Micro-benchmarks behind DESIGN.md's descent-kernel choices (sm_100a): throughput of the packed fp32 pipe instructions
// (FMUL2/FFMA2/FADD2) against their scalar forms, their bit-exactness (round-to-nearest, denormals kept), and the cost of
// the shared-reciprocal exact division used by the descent's sequential sums.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o ubench tools/ubench_fp32x2.cu && ./ubench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { return ((u64)__float_as_uint(b) << 32) | __float_as_uint(a); }
__device__ __forceinline__ float lo(u64 v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float hi(u64 v) { return __uint_as_float((unsigned)(v >> 32)); }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// ---- throughput: ILP independent chains of FFMA (scalar) / FFMA2 (packed) -----------------------------------------------
template <int ILP> __global__ void k_ffma(float *out, int n, float a, float b) {
    float x[ILP];
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x + i;
    for (int j = 0; j < n; j++)
#pragma unroll
        for (int i = 0; i < ILP; i++) x[i] = __fmaf_rn(x[i], a, b);
    float s = 0; for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP> __global__ void k_ffma2(float *out, int n, float a, float b) {
    u64 x[ILP]; u64 a2 = pk(a, a), b2 = pk(b, b);
    for (int i = 0; i < ILP; i++) x[i] = pk(threadIdx.x + i, i);
    for (int j = 0; j < n; j++)
#pragma unroll
        for (int i = 0; i < ILP; i++) x[i] = fma2(x[i], a2, b2);
    float s = 0; for (int i = 0; i < ILP; i++) s += lo(x[i]) + hi(x[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP> __global__ void k_fadd2(float *out, int n, float a) {
    u64 x[ILP]; u64 a2 = pk(a, a);
    for (int i = 0; i < ILP; i++) x[i] = pk(threadIdx.x + i, i);
    for (int j = 0; j < n; j++)
#pragma unroll
        for (int i = 0; i < ILP; i++) x[i] = add2(x[i], a2);
    float s = 0; for (int i = 0; i < ILP; i++) s += lo(x[i]) + hi(x[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- bit-exactness of the packed ops, including denormal operands/results ---------------------------------------------------
__device__ __forceinline__ unsigned rnd(unsigned &s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }
__device__ __forceinline__ float rndf(unsigned &s, int mode) {
    unsigned m = rnd(s) & 0x7FFFFF, sg = rnd(s) & 0x80000000u;
    int e = mode == 0 ? 100 + rnd(s) % 56 : (mode == 1 ? rnd(s) % 4 : 1 + rnd(s) % 253);   // normal-ish / denormal edge / any
    return __uint_as_float(sg | ((unsigned)e << 23) | m);
}
__global__ void k_exact(u64 *bad, int n) {
    unsigned s = 0x9E3779B9u * (blockIdx.x * blockDim.x + threadIdx.x + 1);
    u64 b = 0;
    for (int j = 0; j < n; j++) {
        int mode = j % 3;
        float a0 = rndf(s, mode), a1 = rndf(s, mode), b0 = rndf(s, mode), b1 = rndf(s, 0), c0 = rndf(s, mode), c1 = rndf(s, mode);
        u64 m = mul2(pk(a0, a1), pk(b0, b1)), f = fma2(pk(a0, a1), pk(b0, b1), pk(c0, c1)), d = add2(pk(a0, a1), pk(c0, c1));
        b += __float_as_uint(lo(m)) != __float_as_uint(__fmul_rn(a0, b0));
        b += __float_as_uint(hi(m)) != __float_as_uint(__fmul_rn(a1, b1));
        b += __float_as_uint(lo(f)) != __float_as_uint(__fmaf_rn(a0, b0, c0));
        b += __float_as_uint(hi(f)) != __float_as_uint(__fmaf_rn(a1, b1, c1));
        b += __float_as_uint(lo(d)) != __float_as_uint(__fadd_rn(a0, c0));
        b += __float_as_uint(hi(d)) != __float_as_uint(__fadd_rn(a1, c1));
    }
    if (b) atomicAdd(bad, b);
}

// ---- the descent's inner loop, three formulations, per (lane = env): A elements, both chains --------------------------------
// smem rows: top[A] (pitch P), ext[2][A]; returns S and g; writes S prefix
template <int MODE> __global__ void k_pass(float *out, int A, int P, int passes, float alpha) {
    extern __shared__ float4 sm4[];
    float *sm = reinterpret_cast<float *>(sm4);
    const int nt = blockDim.x;
    float *top = sm + (size_t)threadIdx.x * P, *es = sm + (size_t)nt * P + (size_t)threadIdx.x * P, *eg = sm + (size_t)2 * nt * P + (size_t)threadIdx.x * P;
    for (int a = 0; a < P; a++) { top[a] = a < A ? 1e-3f * (1 + ((a * 7 + threadIdx.x) % 13)) : 0.f; es[a] = 0.f; eg[a] = 0.f; }
    __syncthreads();
    float accS = 0, accG = 0;
    for (int p = 0; p < passes; p++) {
        const float bS = alpha, bG = __fmul_rn(alpha, alpha);
        const float yS = __frcp_rn(bS), yG = -__frcp_rn(bG);
        accS = 0; accG = 0;
        if (MODE == 0) {            // scalar ops
            for (int a = 0; a < P; a += 4) {
                float4 t4 = *reinterpret_cast<float4 *>(top + a), e4 = *reinterpret_cast<float4 *>(es + a), g4 = *reinterpret_cast<float4 *>(eg + a);
                float t[4] = {t4.x, t4.y, t4.z, t4.w}, e[4] = {e4.x, e4.y, e4.z, e4.w}, g[4] = {g4.x, g4.y, g4.z, g4.w}, o[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    float q0 = __fmul_rn(t[u], yS), r0 = __fmaf_rn(-bS, q0, t[u]), q1 = __fmaf_rn(r0, yS, q0);
                    float h0 = __fmul_rn(t[u], yG), s0 = __fmaf_rn(bG, h0, t[u]), h1 = __fmaf_rn(s0, yG, h0);
                    accS = __fadd_rn(__fadd_rn(accS, q1), e[u]);
                    accG = __fadd_rn(__fadd_rn(accG, h1), g[u]);
                    o[u] = accS;
                }
                *reinterpret_cast<float4 *>(es + a) = make_float4(o[0], o[1], o[2], o[3]);
            }
        } else if (MODE == 1) {     // packed over element pairs of one chain
            const u64 yS2 = pk(yS, yS), yG2 = pk(yG, yG), nbS2 = pk(-bS, -bS), bG2 = pk(bG, bG);
            for (int a = 0; a < P; a += 4) {
                const ulonglong2 t2 = *reinterpret_cast<ulonglong2 *>(top + a);
                float4 e4 = *reinterpret_cast<float4 *>(es + a), g4 = *reinterpret_cast<float4 *>(eg + a);
                u64 tt[2] = {t2.x, t2.y};
                float e[4] = {e4.x, e4.y, e4.z, e4.w}, g[4] = {g4.x, g4.y, g4.z, g4.w}, o[4];
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    u64 q0 = mul2(tt[u], yS2), r0 = fma2(nbS2, q0, tt[u]), q1 = fma2(r0, yS2, q0);
                    u64 h0 = mul2(tt[u], yG2), s0 = fma2(bG2, h0, tt[u]), h1 = fma2(s0, yG2, h0);
                    accS = __fadd_rn(__fadd_rn(accS, lo(q1)), e[2 * u]); o[2 * u] = accS;
                    accG = __fadd_rn(__fadd_rn(accG, lo(h1)), g[2 * u]);
                    accS = __fadd_rn(__fadd_rn(accS, hi(q1)), e[2 * u + 1]); o[2 * u + 1] = accS;
                    accG = __fadd_rn(__fadd_rn(accG, hi(h1)), g[2 * u + 1]);
                }
                *reinterpret_cast<float4 *>(es + a) = make_float4(o[0], o[1], o[2], o[3]);
            }
        } else {                    // packed over the two chains of one element
            const u64 y2 = pk(yS, yG), b2 = pk(-bS, bG);
            u64 acc = pk(0.f, 0.f);
            for (int a = 0; a < P; a += 4) {
                float4 t4 = *reinterpret_cast<float4 *>(top + a), e4 = *reinterpret_cast<float4 *>(es + a), g4 = *reinterpret_cast<float4 *>(eg + a);
                float t[4] = {t4.x, t4.y, t4.z, t4.w}, e[4] = {e4.x, e4.y, e4.z, e4.w}, g[4] = {g4.x, g4.y, g4.z, g4.w}, o[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    u64 n2 = pk(t[u], t[u]);
                    u64 q0 = mul2(n2, y2), r0 = fma2(b2, q0, n2), q1 = fma2(r0, y2, q0);
                    acc = add2(add2(acc, q1), pk(e[u], g[u]));
                    o[u] = lo(acc);
                }
                *reinterpret_cast<float4 *>(es + a) = make_float4(o[0], o[1], o[2], o[3]);
            }
            accS = lo(acc); accG = hi(acc);
        }
        alpha = __fadd_rn(alpha, 1e-7f * accS);
        __syncwarp();
        for (int a = 0; a < P; a += 4) *reinterpret_cast<float4 *>(es + a) = make_float4(0, 0, 0, 0);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = accS + accG;
}

template <class F> float timeit(F f, int reps = 5) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); best = ms < best ? ms : best; }
    return best;
}

int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%s, %d SMs, %d kHz\n", pr.name, pr.multiProcessorCount, clk);
    const int SMS = pr.multiProcessorCount;
    float *out; cudaMalloc(&out, 1 << 26);
    const int n = 4096;
    for (int wps : {4, 8, 16, 32}) {        // warps per SM
        int blocks = SMS, threads = wps * 32;
        if (threads > 1024) { blocks = SMS * (threads / 1024); threads = 1024; }
        float m1 = timeit([&] { k_ffma<8><<<blocks, threads>>>(out, n, 1.0001f, 0.5f); });
        float m2 = timeit([&] { k_ffma2<8><<<blocks, threads>>>(out, n, 1.0001f, 0.5f); });
        float m3 = timeit([&] { k_fadd2<8><<<blocks, threads>>>(out, n, 0.5f); });
        double inst = (double)n * 8 * wps;    // warp-instructions per SM
        double cyc1 = m1 * 1e-3 * clk * 1e3, cyc2 = m2 * 1e-3 * clk * 1e3, cyc3 = m3 * 1e-3 * clk * 1e3;
        printf("warps/SM %2d: FFMA %.3f ms (%.2f warp-inst/clk/SM)  FFMA2 %.3f ms (%.2f)  FADD2 %.3f ms (%.2f)\n", wps, m1, inst / cyc1, m2, inst / cyc2, m3, inst / cyc3);
    }
    u64 *bad; cudaMalloc(&bad, 8); cudaMemset(bad, 0, 8);
    k_exact<<<SMS * 4, 256>>>(bad, 20000);
    u64 hb; cudaMemcpy(&hb, bad, 8, cudaMemcpyDeviceToHost);
    printf("packed-op mismatches vs scalar RN ops (incl. denormals): %llu of %llu\n", hb, (u64)SMS * 4 * 256 * 20000 * 6);

    const int A = 81, P = 84, passes = 200;
    for (int threads : {32, 64, 128}) {
        size_t smem = (size_t)3 * threads * P * 4;
        for (int mode = 0; mode < 3; mode++) {
            for (int bps : {1, 2}) {
                if (smem * bps > 220 * 1024) continue;
                auto launch = [&] {
                    if (mode == 0) { cudaFuncSetAttribute(k_pass<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k_pass<0><<<SMS * bps, threads, smem>>>(out, A, P, passes, 0.7f); }
                    if (mode == 1) { cudaFuncSetAttribute(k_pass<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k_pass<1><<<SMS * bps, threads, smem>>>(out, A, P, passes, 0.7f); }
                    if (mode == 2) { cudaFuncSetAttribute(k_pass<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k_pass<2><<<SMS * bps, threads, smem>>>(out, A, P, passes, 0.7f); }
                };
                float ms = timeit(launch);
                double cyc = ms * 1e-3 * clk * 1e3;
                double lane_elems_per_sm = (double)threads * bps * passes * P;
                printf("pass mode %d, %3d thr x %d CTA/SM: %.3f ms  %.2f cyc per element per warp, %.1f lane-elements/clk/SM\n", mode, threads, bps, ms,
                       cyc / (passes * P), lane_elems_per_sm / cyc);
            }
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
