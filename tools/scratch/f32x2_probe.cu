// probe: issue / pipe cost of packed fma.rn.f32x2 (FFMA2) against scalar FFMA on sm_100a
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint64_t pack(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
template <int MODE>
__global__ void k(float* out, int iters, float s) {
    float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    uint64_t p0 = pack(a0, a1), p1 = pack(a2, a3), p2 = pack(a4, a5), p3 = pack(a6, a7), ps = pack(s, s), pc = pack(0.5f, 0.25f);
    for (int i = 0; i < iters; i++) {
        if (MODE == 0) {   // 8 scalar FFMA (independent chains)
            a0 = fmaf(a0, s, 0.5f); a1 = fmaf(a1, s, 0.25f); a2 = fmaf(a2, s, 0.5f); a3 = fmaf(a3, s, 0.25f);
            a4 = fmaf(a4, s, 0.5f); a5 = fmaf(a5, s, 0.25f); a6 = fmaf(a6, s, 0.5f); a7 = fmaf(a7, s, 0.25f);
        } else {           // 4 FFMA2 = the same 8 multiply-adds
            p0 = fma2(p0, ps, pc); p1 = fma2(p1, ps, pc); p2 = fma2(p2, ps, pc); p3 = fma2(p3, ps, pc);
        }
    }
    if (MODE == 1) { unpack(p0, a0, a1); unpack(p1, a2, a3); unpack(p2, a4, a5); unpack(p3, a6, a7); }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 2; mode++) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 1024>>>(d, iters, 0.999f); else k<1><<<148 * 8, 1024>>>(d, iters, 0.999f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double fma = 148.0 * 8 * 1024 * 8.0 * iters;
            if (rep) printf("mode %d (%s): %.3f ms, %.1f Gfma/s (%.1f per SM per clk at 1.965 GHz)\n", mode, mode ? "fma.rn.f32x2" : "scalar fma", ms, fma / ms / 1e6, fma / ms / 1e6 / 148 / 1.965);
        }
    }
    return 0;
}
