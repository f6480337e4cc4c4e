/*
 * d3d_sweep.h -- C ABI of libd3dsweep.so, the sm_100a plane-sweep cost-volume engine.
 *
 * The reference (gpcv-liujin/Deep3D_Aerial) has no FFI layer: its hot path is a set of Python
 * functions under mvs/mvs_cas/models/ that call ATen ops.  Every entry point below therefore
 * cites the reference *Python* interface it stands in for; the Python shim that re-exports those
 * names lives in deep3d_aerial_b200/ (module.py, cas_mvsnet.py, adamvs.py, msrednet.py, ucsnet.py)
 * and INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the current CUDA device; the library never allocates,
 *     frees or keeps caller memory;
 *   - work is enqueued on the caller's stream (a cudaStream_t passed as void*) and the call
 *     returns without synchronising; it is safe under CUDA-graph capture;
 *   - return value 0 = success, otherwise a D3D_ERR_* code and d3d_last_error() (thread local)
 *     describes it; nothing is launched on error;
 *   - all floating point data is fp32; image extents are rows x cols = height x width;
 *   - batch is handled by the caller (one call per batch item; the reference runs B = 1,
 *     mvs/mvs_cas/predict.py:49).
 *   - structs start with struct_size = sizeof(struct) so fields can be appended compatibly.
 */
#ifndef D3D_SWEEP_H_
#define D3D_SWEEP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define D3D_VERSION 100 /* 0.1.0 */

enum {
    D3D_OK = 0,
    D3D_ERR_BAD_ARGUMENT = 1, /* null pointer, non-positive extent, unknown enum, bad struct_size */
    D3D_ERR_UNSUPPORTED = 2,  /* shape outside what the kernels are instantiated for            */
    D3D_ERR_CUDA = 3          /* a CUDA runtime call failed; message holds cudaGetErrorString     */
};

/* How the per-view warped features are aggregated into the cost volume. */
enum {
    /* out[C,D,H,W] = the warped source view itself (num_views must be 2).
     * Replaces homo_warping_float, mvs/mvs_cas/models/module.py:516-557. */
    D3D_AGG_WARP = 0,
    /* out[C,D,H,W] = sq/V - (sum/V)^2 over the reference and the V-1 warped views.
     * Replaces cas_mvsnet.py:46-60, msrednet.py:217-230 and 401-414, ucsnet.py:119-134. */
    D3D_AGG_VARIANCE = 1,
    /* out[G,D,H,W] = mean over source views of mean_{C/G}(ref * warped).
     * Stands in for the commented-out groupwise_correlation(ref, warped, 8, 1) at
     * adamvs.py:271,295 (body not in the reference; defined by analogy with adamvs.py:473). */
    D3D_AGG_GROUP_CORR = 2,
    /* out[C,D,H,W] = sum_i (warped_i*ref)*w_i / (1e-5 + sum_i w_i)      (adamvs.py:492-509), or with
     * eps_in_numerator: (1e-5 + sum_i (warped_i*ref)*w_i) / sum_i w_i   (adamvs.py:262,287-301). */
    D3D_AGG_WEIGHTED_PRODUCT = 3,
    /* out[V-1,D,H,W]: out[i] = mean_C(ref * warped_i)                   (adamvs.py:466-475). */
    D3D_AGG_PAIR_MEAN = 4
};

typedef struct D3dCostVolumeArgs {
    uint32_t struct_size;
    int32_t mode;             /* D3D_AGG_*                                                        */
    int32_t num_views;        /* V: reference + V-1 sources, 2 <= V <= 9                          */
    int32_t channels;         /* C: multiple of 4; C/4 or C/8 must be a power of two <= 32        */
    int32_t height, width;    /* feature-map rows, cols                                           */
    int32_t num_depth;        /* D: planes in `hyps`                                              */
    int32_t d_begin, d_count; /* planes [d_begin, d_begin+d_count) are computed (slice mode for
                                 the plane-at-a-time GRU regularisers, adamvs.py:492, msrednet.py:400);
                                 d_count <= 0 means "through the last plane"                      */
    int32_t hyps_per_pixel;   /* 0: hyps[D] (fronto-parallel sweep); 1: hyps[D,H,W]               */
    int32_t groups;           /* G for D3D_AGG_GROUP_CORR (C % G == 0)                            */
    int32_t eps_in_numerator; /* D3D_AGG_WEIGHTED_PRODUCT: training-form epsilon placement        */
    int32_t variant;          /* 0 = production kernel; others select A/B kernels */
    int32_t texel_slots;      /* 0: `feats` is the dense [V,H,W,C] block of this reference view.  S > 0: `feats` is a
                                 texel POOL [S,H,W,C] of per-image maps and view v lives in slot view_slot[v] -- the
                                 images of a scene block are laid out once (d3d_nchw_to_nhwc per IMAGE) and every
                                 reference view names the five it uses (adamvs.py:570-574 re-encodes them per view) */
    const float* feats;       /* [V,H,W,C] channels-last, view 0 = reference (d3d_nchw_to_nhwc); or the pool       */
    const float* pose;        /* [V-1,4,4] row-major P_src @ inverse(P_ref), module.py:528-530    */
    const float* hyps;        /* depth hypotheses, see hyps_per_pixel                             */
    const float* weights;     /* [V-1,H,W] view weights (WEIGHTED_PRODUCT only), already at H x W  */
    float* out;               /* plane d of channel c at out[c*out_stride_c + (d-d_begin)*out_stride_d
                                 + y*W + x]                                                       */
    int64_t out_stride_c;     /* elements; 0 selects the dense [Cout, d_count, H, W] layout        */
    int64_t out_stride_d;     /* elements; 0 selects H*W                                          */
    const float* rays;        /* optional [V-1,3,H*W]: rot_i @ [x,y,1] for every reference pixel, i.e. the
                                 result of module.py:538 computed by the caller with the reference's own
                                 matmul.  NULL: the kernel forms the rays itself as
                                 fma(r2,1,fma(r1,y,r0*x)), the order cuBLAS uses for most -- not all --
                                 problem sizes (Numerics).  Passing them makes the sample
                                 coordinates bit-identical to the reference's at every size           */
    int32_t view_slot[9];     /* texel_slots > 0: pool slot of view v (v = 0: the reference), each < texel_slots  */
    int32_t reserved1;
} D3dCostVolumeArgs;

/* Fused homography warp + bilinear sample + aggregation; the V x C x D x H x W warped volume is
 * never materialised. */
int d3d_cost_volume(const D3dCostVolumeArgs* args, void* cuda_stream);

enum { D3D_SOFTMAX_STABLE = 0, /* p = softmax_D(logit)  (F.softmax, cas_mvsnet.py:69)               */
       D3D_SOFTMAX_RAW_EXP = 1,/* un-normalised exp(logit), no max subtraction, streaming
                                  accumulators (adamvs.py:514-529, msrednet.py:418-437)             */
       D3D_SOFTMAX_NONE = 2    /* input already is a probability volume: depth = sum p*d, conf =
                                  max p, no normalisation (depth_regression, module.py:605-613)     */ };
enum { D3D_CONF_MAX_PROB = 0,  /* conf = max_d p, index = argmax (msrednet.py:238, adamvs.py:310)   */
       D3D_CONF_WINDOW4 = 1    /* i = clamp(floor(sum p*k)); conf = p[i-1]+p[i]+p[i+1]+p[i+2]
                                  (cas_mvsnet.py:72-76, ucsnet.py:140-146)                          */ };
enum { D3D_HYPS_UNIFORM = 0,   /* hyps[D]                                                           */
       D3D_HYPS_PER_PIXEL = 1, /* hyps[D,H,W]                                                       */
       D3D_HYPS_RESIZED = 2    /* hyps[D,hyps_height,hyps_width] bilinearly resized to H x W with
                                  align_corners=False (module.py:608-610, adamvs.py:519-520)        */ };

#define D3D_REGRESS_MAX_PLANES 16

typedef struct D3dRegressArgs {
    uint32_t struct_size;
    int32_t num_depth;        /* D: planes of the whole sweep                                     */
    int32_t height, width;    /* rows, cols of the logit maps                                     */
    int32_t d_begin, d_count; /* `logits` holds planes [d_begin, d_begin+d_count); d_count <= 0:
                                 all D.  Slices need D3D_SOFTMAX_RAW_EXP or _NONE                 */
    int32_t softmax_mode;     /* D3D_SOFTMAX_*                                                    */
    int32_t conf_mode;        /* D3D_CONF_* (RAW_EXP implies MAX_PROB)                            */
    int32_t hyps_mode;        /* D3D_HYPS_*                                                       */
    int32_t hyps_height, hyps_width; /* D3D_HYPS_RESIZED only                                     */
    int32_t finalize;         /* RAW_EXP/NONE: 1 = also write depth and conf (RAW_EXP: acc/(sum+1e-10)) */
    int32_t next_num_depth;   /* > 0: also emit next-stage hypotheses (module.py:616-630)         */
    float lamb;               /* exp_variance scale (ucsnet.py:148-149)                           */
    int32_t reserved0;
    double next_interval;     /* depth_inteval_pixel of the next stage (a Python float upstream)  */
    const float* logits;      /* [d_count,H,W], plane stride logits_stride_d                      */
    int64_t logits_stride_d;  /* elements; 0 selects H*W                                          */
    const float* hyps;        /* all D planes (indexed with d_begin+k), layout per hyps_mode      */
    float* depth;             /* [H,W] expected depth (module.py:605-613); may be NULL if !finalize */
    float* conf;              /* [H,W] photometric confidence                                     */
    int32_t* index;           /* [H,W] argmax plane (MAX_PROB) or window base i (WINDOW4); NULL ok */
    float* state;             /* RAW_EXP: [3,H,W] = exp_sum, depth_sum, max_prob; read unless
                                 d_begin == 0, always written.  NULL allowed when the call covers
                                 all D planes                                                      */
    float* exp_variance;      /* [H,W] lamb*sqrt(sum p (d-depth)^2), STABLE only; NULL = skip     */
    float* next_hyps;         /* [next_num_depth,H,W]: depth -/+ next_num_depth/2*next_interval   */
    const float* logit_planes[D3D_REGRESS_MAX_PLANES]; /* RAW_EXP / NONE only, used when `logits` is NULL:
                                 plane d_begin+k of the slice is the [H,W] map logit_planes[k] (k < d_count
                                 <= D3D_REGRESS_MAX_PLANES).  Lets a caller whose regulariser hands back one
                                 freshly allocated plane per call (adamvs.py:512, msrednet.py:416) keep K
                                 planes alive and fold them into the accumulators with ONE launch: the
                                 3-map state is then read and written once per K planes, not once per plane */
} D3dRegressArgs;

/* Fused softmax over D + expected depth + confidence (+ argmax, + UCS-Net spread, + next-stage
 * hypotheses).  One pass over the logits. */
int d3d_depth_regress(const D3dRegressArgs* args, void* cuda_stream);

enum { D3D_SAMPLES_RANGE = 0,     /* [dmin,dmax] -> linspace planes broadcast to [D,H,W]
                                     (module.py:637-645)                                          */
       D3D_SAMPLES_AROUND = 1,    /* cur[H,W] -> cur -/+ D/2*interval, D planes (module.py:616-630) */
       D3D_SAMPLES_CASCADE = 2    /* Cas-MVSNet / RED-Net stage glue (cas_mvsnet.py:206-226,
                                     msrednet.py:495-515): cur[src_h,src_w] (or the range) bilinear
                                     -> full res, samples at full res, trilinear -> [D,H,W]        */,
       D3D_SAMPLES_SPREAD = 3     /* UCS-Net: cur[H,W] -/+ spread[H,W] (the exp_variance map of the stage before) in D
                                     steps, + 1e-12 (uncertainty_aware_samples, ucsnet.py:41-51)       */ };

typedef struct D3dSamplesArgs {
    uint32_t struct_size;
    int32_t mode;              /* D3D_SAMPLES_*                                                    */
    int32_t num_depth;         /* D of the stage being prepared                                    */
    int32_t height, width;     /* output rows, cols                                                */
    int32_t src_height, src_width; /* extent of `cur` (CASCADE); ignored otherwise                 */
    int32_t full_height, full_width; /* CASCADE: full-resolution extent                            */
    float dmin, dmax;          /* RANGE, and CASCADE when cur == NULL                              */
    int32_t reserved0;
    double interval;           /* depth_inteval_pixel = ratio * (dmax-dmin)/num_depth_total        */
    const float* cur;          /* previous depth estimate; NULL with CASCADE = first stage (range) */
    float* out;                /* [D,H,W]                                                          */
    const float* spread;       /* SPREAD: per-pixel half width of the sampled interval, [H,W]      */
} D3dSamplesArgs;

/* Per-stage depth-hypothesis resampling, get_depth_range_samples (module.py:633-650). */
int d3d_depth_samples(const D3dSamplesArgs* args, void* cuda_stream);

/* out[V-1,3,H*W] = rot_i @ [x,y,1] for every reference pixel, rounded as the sweep kernels round it when
 * D3dCostVolumeArgs.rays is NULL: fma(r2, 1, fma(r1, y, r0*x)).  The binding compares this with the reference's own
 * torch.matmul (module.py:538) once per image size to learn whether cuBLAS rounds in that order at that size; where
 * it does, the sweeps form their rays themselves and the matmul (a 0.75 TB/s skinny GEMM) is skipped. */
int d3d_pixel_rays(const float* pose, int32_t num_src, int32_t height, int32_t width, float* out, void* cuda_stream);

/* out[N,Ho,Wo] = F.interpolate(in[N,Hi,Wi], [Ho,Wo], mode="bilinear", align_corners=False): the resize of the AdaMVS pair
 * confidences between stages (mvs/mvs_cas/models/adamvs.py:291-302), ATen's upsample_bilinear2d formula. */
int d3d_resize_bilinear(const float* in, float* out, int32_t maps, int32_t in_height, int32_t in_width,
                        int32_t out_height, int32_t out_width, void* cuda_stream);

/* ---- homo_warping_double: the warp with fp64 coordinate arithmetic (mvs/mvs_cas/models/module.py:560-601) -----
 * The reference forms rot @ [x,y,1], X = rot_xyz * d + trans, X/Z and u / ((W-1)/2) - 1 in fp64 (its projection
 * matrices must be fp64 for that: torch.matmul does not promote), casts the normalised grid to fp32 and samples it with
 * the fp32 grid_sample (bilinear, zeros padding, align_corners=True).  out[C,D,H,W] = the warped source view.
 *   texels: [H,W,C] channels-last source features;  pose64: [4,4] row-major fp64 P_src @ inverse(P_ref);
 *   hyps: fp32 [D] or [D,H,W] (widened to fp64 in the kernel, as `.double()` does). */
int d3d_homo_warp_f64(const float* texels, const double* pose64, const float* hyps, int32_t hyps_per_pixel,
                      int32_t channels, int32_t num_depth, int32_t height, int32_t width, float* out, void* cuda_stream);

/* Feature relayout [C,H,W] -> [H,W,C] (the sweep kernel gathers whole texels). */
int d3d_nchw_to_nhwc(const float* in, float* out, int32_t channels, int32_t height, int32_t width,
                     void* cuda_stream);

/* ---- depth-map fusion: geometric consistency check ------------------------------------
 * Replaces ConsistencyChecker.check / check_cupy (fuse/consistency_check_n.py:29-148, ~25 CuPy launches and
 * host round trips per source view) and the per-source accumulation loop of Fuse_Depth_Map.fuse_depths
 * (fuse/fusion_3d_normal.py:436-530) with ONE launch per reference view over all its source views.
 *
 * Per reference pixel (x, y) and source view s, in the reference's own types (points fp64, normals fp32):
 *   P = Kinv_ref (x d, y d, d);  Q = (E_s Einv_ref)[P;1];  (xs, ys) = trunc(K_s Q / z + 0.5)        (:51-72)
 *   gather depth_src / normal_src at (ys, xs), indices wrapped modulo the map extent (CuPy's out-of-bounds rule)
 *   back-project with the sampled depth to the world and into the reference camera                     (:76-92)
 *   consistent = |p' - p| < position_threshold  &&  |d' - d| / d < depth_threshold  &&  prob > confidence_threshold
 *                &&  cos(n_ref, n_src) > normal_threshold_cos  &&  d > 0                                (:95-123)
 * The 4x4 / 3x3 matrix products and inverses are formed by the CALLER with the reference's own numpy calls, in
 * the matrices' own dtype (fp32 in fusion_3d_normal.py:118-126), and handed over as fp64 in `geometry`:
 * one block of 64 doubles per view, matrices row-major with every row padded to 4 entries (3x3 -> 3x4, pad 0):
 *   doubles [0..11] A 3x4, [12..23] B 3x4, [24..35] C 3x4, [36..51] D 4x4, [52..63] R 3x4
 *   block 0 (the reference view):  A = Kinv_ref, B = rows 0-2 of E_ref, C = K_ref, D = Einv_ref,
 *                                  R = inverse(E_ref[:3,:3])
 *   block 1+s (source view s):     A = rows 0-2 of E_s @ Einv_ref, B = K_s, C = Kinv_s, D = Einv_s,
 *                                  R = inverse(E_s[:3,:3])
 * Every output pointer may be NULL (= not wanted).
 */
#define D3D_FUSE_MAX_SRC 16
#define D3D_FUSE_GEOM_DOUBLES 64

typedef struct D3dFuseArgs {
    uint32_t struct_size;
    int32_t num_src;              /* S: 1..D3D_FUSE_MAX_SRC source views                              */
    int32_t height, width;        /* reference maps                                                   */
    int32_t src_height, src_width;/* source maps (the reference assumes the same extent)              */
    int32_t min_consistent;       /* final_mask = 1 + #consistent sources >= this (fusion_3d_normal.py:536) */
    int32_t accumulate;           /* 0: the accumulators start from the reference pixel itself (count 1, its world
                                     point, confidence 1: fusion_3d_normal.py:449-455, 466); 1: they continue from
                                     `accum` and `consistent_count` as an earlier call over OTHER source views of the
                                     same reference view left them (a view list that names a source twice must see
                                     the map the first visit modified, so it is split into several calls)          */
    double position_threshold;    /* pixels, compared in fp64                                         */
    float depth_threshold;        /* relative, compared in fp32                                       */
    float confidence_threshold;   /* on prob_ref, fp32                                                */
    float normal_threshold_cos;   /* cos(normal_threshold degrees), fp32                              */
    int32_t reserved1;
    const float* depth_ref;       /* [H,W]                                                            */
    const float* normal_ref;      /* [H,W,3] camera-frame normals                                     */
    const float* prob_ref;        /* [H,W] photometric confidence                                     */
    const double* geometry;       /* [(1+S) * D3D_FUSE_GEOM_DOUBLES], see above                       */
    const float* depth_src[D3D_FUSE_MAX_SRC];   /* S x [Hs,Ws]                                        */
    const float* normal_src[D3D_FUSE_MAX_SRC];  /* S x [Hs,Ws,3]                                      */
    float* depth_src_out[D3D_FUSE_MAX_SRC];     /* S x [Hs,Ws]: depth_src with the pixels consumed by this
                                     reference view set to 0 (:128-131); the library copies depth_src[s]
                                     into it first.  Must not alias depth_src[s]                      */
    uint8_t* mask;                /* [S,H,W] consistency mask per source                              */
    float* depth_reprojected;     /* [S,H,W] d' where consistent, else 0                              */
    float* xyz_world_src;         /* [S,3,H,W] world point of the source sample where consistent, else 0 */
    float* angle_conf;            /* [S,H,W] max(cos, 0) where consistent, else 0 (one plane; the reference
                                     repeats it three times, :112)                                    */
    int32_t* consistent_count;    /* [H,W] geo_mask_sum = 1 + sum_s mask_s                            */
    float* xyz_fused;             /* [3,H,W] (xyz_ref + sum_s conf_s xyz_s) / (1 + sum_s conf_s)      */
    uint8_t* final_mask;          /* [H,W]                                                            */
    float* depth_ref_filtered;    /* [H,W] depth_ref where final_mask, else 0 (fusion_3d_normal.py:541) */
    float* accum;                 /* [4,H,W] running sum_s conf_s x, y, z and sum_s conf_s (incl. the reference
                                     pixel's own point at confidence 1): written when not NULL, read when `accumulate` */
} D3dFuseArgs;

int d3d_consistency_fuse(const D3dFuseArgs* args, void* cuda_stream);

/* Thread-local description of the last error returned on this thread ("" if none). */
const char* d3d_last_error(void);
int d3d_version(void);
/* Number of kernels this library has launched in the calling process (bench accounting). */
int64_t d3d_launch_count(void);
/* sizeof() of an argument struct as compiled into the library: 0 = D3dCostVolumeArgs,
 * 1 = D3dRegressArgs, 2 = D3dSamplesArgs, 3 = D3dFuseArgs; -1 for anything else.  Lets a binding check its layout
 * without a GPU. */
int32_t d3d_abi_sizeof(int32_t which);

#ifdef __cplusplus
}
#endif
#endif /* D3D_SWEEP_H_ */
