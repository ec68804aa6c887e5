// synccheck probe: warp-specialised named barrier (different call sites, explicit count = block size)
#include <cstdio>
__device__ __noinline__ void cta_bar() { asm volatile("bar.sync 1, 288;" ::: "memory"); }
__global__ void k(int* out) {
    __shared__ int s[9];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 8) {
        if (lane == 0) s[8] = 100;
        cta_bar();
        if (lane == 0) out[blockIdx.x * 2] = s[0] + s[8];
        return;
    }
    if (lane == 0) s[warp] = warp + 1;
    cta_bar();
    if (threadIdx.x == 0) out[blockIdx.x * 2 + 1] = s[8];
}
int main() {
    int* d; cudaMalloc(&d, 64);
    k<<<4, 288>>>(d);
    int h[8]; cudaError_t e = cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
    printf("%s %d %d\n", cudaGetErrorString(e), h[0], h[1]);
    return 0;
}
