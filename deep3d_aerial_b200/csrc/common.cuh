// Shared helpers for libd3dsweep (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "d3d_sweep.h"

namespace d3d {

// error plumbing (abi.cu)
int fail(int code, const char* fmt, ...);
int check_launch(const char* what);
void count_launch(int n = 1);

constexpr int kMaxSrc = 8;  // V-1

// Kernel parameters of the fused sweep (by value: lands in the constant bank).
struct SweepParams {
    const float* __restrict__ feats;    // [V,H,W,C]
    const float* __restrict__ pose;     // [V-1,4,4]
    const float* __restrict__ hyps;     // [D] or [D,H,W]
    const float* __restrict__ weights;  // [V-1,H,W] or null
    const float* __restrict__ rays;     // [V-1,3,H*W] rot @ [x,y,1] per source view, or null (computed here)
    float* __restrict__ out;
    long long out_sc, out_sd;
    int C, H, W, HW;
    int d_begin, d_end;   // planes computed by this launch
    int d_chunk;          // planes per blockIdx.y
    int lpp_log2;         // log2(lanes per pixel) = log2(C / CPT)
    int perpix;           // hyps layout
    int groups, eps_num;
    int flags;            // bit 0: prefetch moved footprints one pass ahead
    float inv_half_w, inv_half_h;  // 1/((W-1)/2), 1/((H-1)/2)   (module.py:543-544)
    float wm1, hm1;                // W-1, H-1                   (GridSampler.h:27-32)
    // where view v starts in `feats`, as a texel index (slot * H*W): v * H*W for the dense [V,H,W,C] layout, anything else when
    // `feats` is a texel pool [S,H,W,C] and the views are named by slot (D3dCostVolumeArgs.texel_slots)
    int view_tex[kMaxSrc + 1];
    int pooled;                    // 1: view_tex is not the dense layout (kernels that address views by stride refuse)
};

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute: remember the size each kernel was opted
// into per device, so a process that drives several GPUs (or several host threads) configures every one of them.
constexpr int kMaxDevices = 64;
struct SmemOptIn {
    size_t bytes[kMaxDevices] = {};
    template <typename Kernel>
    int ensure(Kernel kern, size_t smem) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
        if (bytes[dev] >= smem) return D3D_OK;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(D3D_ERR_CUDA, "cudaFuncSetAttribute(smem=%zu): %s", smem, cudaGetErrorString(e));
        bytes[dev] = smem;
        return D3D_OK;
    }
};

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

}  // namespace d3d
