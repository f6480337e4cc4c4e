// Instruction-throughput probes for the sweep kernel's design choices (B200, sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/microbench tools/microbench.cu && ./microbench
// Prints warp-instructions per clock per SM for FFMA, FFMA2 (fma.rn.f32x2), FADD2, SHFL, LDS.128
// (8 distinct addresses per warp, like the geometry broadcast), FRND.FLOOR, MUFU.RCP, and a copy
// bandwidth probe for 32-byte-segment stores vs 128-byte-row stores.
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 4096;

__global__ void k_ffma(float* out, float a, float b) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma2(float* out, float a, float b) {
    float2 x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = make_float2(threadIdx.x + i, i);
    float2 aa = make_float2(a, a * 1.01f), bb = make_float2(b, b * 0.99f);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = __ffma2_rn(x[i], aa, bb);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FFMA2 where one operand pair is a broadcast scalar {s,s} produced each iteration (needs MOVs?)
__global__ void k_ffma2_bcast(float* out, const float* in) {
    float2 x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = make_float2(threadIdx.x + i, i);
    float s0 = in[threadIdx.x & 3];
    for (int it = 0; it < ITERS; ++it) {
        float s = s0 + it;                    // a fresh scalar every iteration
        float2 ss = make_float2(s, s);
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = __ffma2_rn(x[i], ss, x[(i + 1) & 15]);
    }
    float t = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) t += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

__global__ void k_fadd2(float* out, float a) {
    float2 x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = make_float2(threadIdx.x + i, i);
    float2 aa = make_float2(a, a * 1.01f);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = __fadd2_rn(x[i], aa);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_shfl(float* out) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = __shfl_sync(0xffffffffu, x[i], (threadIdx.x + 1) & 3, 4);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_lds128(float* out) {
    __shared__ float4 sm[8 * 64];
    for (int i = threadIdx.x; i < 8 * 64; i += blockDim.x) sm[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    float4 acc = make_float4(0, 0, 0, 0);
    int q = (threadIdx.x & 31) >> 2;         // 8 distinct 16-byte addresses per warp
    int base = (threadIdx.x >> 5) * 8;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float4 v = sm[((base + i * 8) & 511) + q + ((it & 1) ? 0 : 0)];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

__global__ void k_floor(float* out, float a) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 0.37f + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = floorf(x[i] * a) + 0.5f;
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_rcp(float* out, float a) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 0.37f + i + 1;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = __frcp_rn(x[i]) + a;   // IEEE rcp
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_rcp_approx(float* out, float a) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 0.37f + i + 1;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x[i])); x[i] = r + a; }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// store patterns: [C=32 rows][N] floats; seg = contiguous floats each group of lanes writes per row
__global__ void k_store(float* out, long long row_stride, long long n, int seg) {
    // warp writes 32 floats per instruction: (32/seg) rows x seg contiguous floats, 8 instructions -> 32 rows..
    int lane = threadIdx.x & 31;
    long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int rows_per_inst = 32 / seg;
    long long col0 = warp * seg;
    if (col0 + seg > n) return;
    int r0 = lane / seg, c = lane % seg;
    for (int i = 0; i < 32 / rows_per_inst; ++i) {
        int row = i * rows_per_inst + r0;
        out[row * row_stride + col0 + c] = (float)lane;
    }
}

template <typename F>
float time_ms(F f, int reps = 3) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(a);
        f();
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return best;
}

// legacy tensor-core path (mma.sync, SASS HMMA): TF32 m16n8k8 and m16n8k4, 8 independent accumulator sets.
// Asked for DESIGN.md "what comes next": would a 3xTF32 split of the bilinear interpolation (one [planes x
// corners] x [corners x channels] product per pixel and view) take the interpolation off the FP32 pipe?
__global__ void k_mma_tf32_k8(float* out) {
    float c[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    unsigned a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = threadIdx.x * 5, a3 = threadIdx.x * 7, b0 = 0x3f800000u, b1 = 0x3f000000u;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_mma_tf32_k4(float* out) {
    float c[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    unsigned a0 = threadIdx.x, a1 = threadIdx.x * 3, b0 = 0x3f800000u;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a0), "r"(a1), "r"(b0));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double clk = khz * 1e3;
    printf("%s: %d SMs, %.0f MHz nominal\n", prop.name, sms, clk / 1e6);
    float* out;
    CK(cudaMalloc(&out, (size_t)64 << 30 >> 2));   // 16 GB for the store probe
    float* in;
    CK(cudaMalloc(&in, 64));
    CK(cudaMemset(in, 0, 64));
    int blocks = sms * 8, threads = 256;
    double warps = (double)blocks * threads / 32;
    auto rep = [&](const char* name, float ms, double inst_per_thread) {
        double total_warp_inst = warps * inst_per_thread;
        double per_clk_sm = total_warp_inst / (ms * 1e-3 * clk) / sms;
        printf("%-28s %8.3f ms   %.2f warp-inst/clk/SM  (%.1f lanes/clk/SM)\n", name, ms, per_clk_sm, per_clk_sm * 32);
    };
    rep("FFMA", time_ms([&] { k_ffma<<<blocks, threads>>>(out, 1.0001f, 0.5f); }), 16.0 * ITERS);
    rep("FFMA2 (f32x2)", time_ms([&] { k_ffma2<<<blocks, threads>>>(out, 1.0001f, 0.5f); }), 16.0 * ITERS);
    rep("FFMA2 bcast scalar operand", time_ms([&] { k_ffma2_bcast<<<blocks, threads>>>(out, in); }), 16.0 * ITERS);
    rep("FADD2 (f32x2)", time_ms([&] { k_fadd2<<<blocks, threads>>>(out, 0.5f); }), 16.0 * ITERS);
    rep("SHFL.IDX width 4", time_ms([&] { k_shfl<<<blocks, threads>>>(out); }), 8.0 * ITERS);
    rep("LDS.128 8 addr/warp", time_ms([&] { k_lds128<<<blocks, threads>>>(out); }), 8.0 * ITERS);
    rep("FRND.FLOOR (+FFMA)", time_ms([&] { k_floor<<<blocks, threads>>>(out, 1.0001f); }), 8.0 * ITERS);
    rep("rcp IEEE (+FADD)", time_ms([&] { k_rcp<<<blocks, threads>>>(out, 0.5f); }), 8.0 * ITERS);
    rep("rcp.approx (+FADD)", time_ms([&] { k_rcp_approx<<<blocks, threads>>>(out, 0.5f); }), 8.0 * ITERS);
    rep("mma.sync m16n8k8 tf32", time_ms([&] { k_mma_tf32_k8<<<blocks, threads>>>(out); }), 8.0 * ITERS);
    rep("mma.sync m16n8k4 tf32", time_ms([&] { k_mma_tf32_k4<<<blocks, threads>>>(out); }), 8.0 * ITERS);
    // store probe: 32 rows x n floats = 12 GB
    long long n = 96LL << 20;
    for (int seg : {4, 8, 16, 32}) {
        long long nwarps = n / seg;
        int tb = 256;
        long long nb = (nwarps + tb / 32 - 1) / (tb / 32);
        float ms = time_ms([&] { k_store<<<(unsigned)nb, tb>>>(out, n, n, seg); });
        printf("store %3d-byte segments       %8.3f ms   %.0f GB/s\n", seg * 4, ms, 32.0 * n * 4 / (ms * 1e-3) / 1e9);
    }
    // plain copy-like store ceiling: contiguous float4 stores
    return 0;
}
