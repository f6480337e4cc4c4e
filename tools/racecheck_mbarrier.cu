// Does compute-sanitizer's racecheck model an inline-PTX mbarrier as synchronisation?  Warp 0 writes a shared array and
// arrives; warp 1 waits on the barrier's parity and reads.  Correct by construction; if racecheck reports a RAW hazard
// here, its reports on the staging rings of sweep_lean / sweep_quad (same protocol) are this limitation.
//   nvcc -arch=sm_100a -lineinfo -o /tmp/rcmb tools/racecheck_mbarrier.cu && compute-sanitizer --tool racecheck /tmp/rcmb
#include <cstdio>
__global__ void k(float* out) {
    __shared__ float buf[32];
    __shared__ unsigned long long bar;
    const unsigned b = (unsigned)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        buf[threadIdx.x] = (float)threadIdx.x;
        __syncwarp();
        if (threadIdx.x == 0) asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(b) : "memory");
    } else {
        asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@!p bra W;\n\t}" ::"r"(b) : "memory");
        out[threadIdx.x - 32] = buf[threadIdx.x - 32];
    }
}
int main() {
    float* d;
    cudaMalloc(&d, 32 * sizeof(float));
    k<<<1, 64>>>(d);
    float h[32];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("out[31] = %g (%s)\n", h[31], cudaGetErrorString(cudaGetLastError()));
    return 0;
}
