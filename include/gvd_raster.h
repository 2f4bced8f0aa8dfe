/*
 * gvd_raster.h -- C ABI of the B200-native differentiable Gaussian rasterizer.
 *
 * Drop-in boundary for the reference's native extension
 *   submodules/diff-gaussian-rasterization-confidence  (abbrev. DGR below).
 * Each entry point names the reference interface it replaces.  All pointers
 * named "dev" are CUDA device pointers; NULL means "argument absent" (the
 * reference passes empty tensors, DGR/diff_gaussian_rasterization/__init__.py:202-212,
 * which its C++ sees as nullptr).  Nothing here allocates device memory: every
 * buffer, including scratch, is owned by the caller (DGR/rasterize_points.cu:73-80
 * uses resize callbacks on torch tensors; we keep that contract with plain C
 * callbacks).  Functions return 0 on success, non-zero on error; the message is
 * available from gvd_last_error().  Nothing throws across the ABI.
 *
 * All launches go to the cudaStream_t passed in (the reference uses the legacy
 * default stream).  The library keeps no global state besides the per-thread
 * error string and cached function attributes.
 */
#ifndef GVD_RASTER_H_
#define GVD_RASTER_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define GVD_API __attribute__((visibility("default")))
#else
#define GVD_API
#endif

typedef struct CUstream_st* gvd_stream_t; /* == cudaStream_t */

/* Resize callback: must return a device pointer to at least `bytes` bytes,
 * 128-byte aligned, valid until the matching backward has run.
 * Replaces std::function<char*(size_t)> of DGR/cuda_rasterizer/rasterizer.h:31-34. */
typedef void* (*gvd_alloc_fn)(void* user, size_t bytes);

/* Layout/version of the scratch buffers (bumped when the packed layouts change). */
#define GVD_RASTER_ABI_VERSION 9

typedef struct GvdRasterForwardArgs {
    /* sizes -- DGR/cuda_rasterizer/rasterizer_impl.cu:197-224 */
    int P;            /* number of Gaussians                       */
    int D;            /* active SH degree 0..3                      */
    int M;            /* SH coefficients stored per Gaussian (16)   */
    int width, height;
    /* inputs (dev) */
    const float* background;     /* [3]                                   */
    const float* means3D;        /* [P,3]                                 */
    const float* shs;            /* [P,M,3] or NULL                       */
    const float* colors_precomp; /* [P,3]   or NULL                       */
    const float* opacities;      /* [P]                                   */
    const float* scales;         /* [P,3]   or NULL                       */
    const float* rotations;      /* [P,4]   or NULL (used un-normalised)  */
    const float* cov3D_precomp;  /* [P,6]   or NULL                       */
    const float* viewmatrix;     /* [16], transposed (row-vector) layout  */
    const float* projmatrix;     /* [16], transposed                      */
    const float* campos;         /* [3]                                   */
    float scale_modifier;
    float tan_fovx, tan_fovy;
    int prefiltered;
    int debug;                   /* sync + check after every stage        */
    int export_keys;             /* also write the sorted 64-bit keys (tile<<32|depth bits) for parity checks */
    /* outputs (dev) */
    float* out_color;            /* [3,H,W]  */
    float* out_depth;            /* [1,H,W]  sum depth*alpha*T (un-normalised) */
    float* out_alpha;            /* [1,H,W]  sum alpha*T                       */
    int*   radii;                /* [P]      */
    /* scratch (caller-owned, via callbacks like the reference) */
    gvd_alloc_fn geom_alloc;     /* called once with gvd_raster_geom_bytes(P,W,H)    */
    gvd_alloc_fn binning_alloc;  /* called once with gvd_raster_binning_bytes(R,export_keys) */
    gvd_alloc_fn img_alloc;      /* called once with gvd_raster_img_bytes(W,H)       */
    gvd_alloc_fn temp_alloc;     /* forward-only scratch: called with gvd_raster_sort_bytes(P) at the start (unless
                                  * sort_buffer is given) and with gvd_raster_hist_bytes(V,W,H) once V is known; both
                                  * buffers may be released (stream-ordered) as soon as the call returns            */
    void* alloc_user;
    /* Optional pre-sized scratch: both sizes are pure functions of (P, W, H), so a caller that knows them can hand the
     * buffers over directly and save the two callbacks.  Used when non-NULL (must hold gvd_raster_geom_bytes(P,W,H) /
     * gvd_raster_img_bytes(W,H) bytes, 128-byte aligned); geom_alloc / img_alloc may then be NULL. */
    void* geom_buffer;
    size_t geom_bytes;
    void* img_buffer;
    size_t img_bytes;
    void* sort_buffer;           /* gvd_raster_sort_bytes(P), forward-only */
    size_t sort_bytes;
    /* R (instances) and V (visible Gaussians) are produced by the second kernel of the forward, ~40 us into the frame.
     * num_rendered_pinned: optional int[2] in pinned host memory that is mapped into the device address space (as
     * cudaHostAlloc memory is under unified addressing); the kernel stores {R, V} there directly.
     *   EXACT path (default, spec_binning_buffer == NULL): the library queues the depth sort behind that kernel, waits
     *   on the host for R and V only (event r_ready_event if given, else an internal one; without num_rendered_pinned
     *   it falls back to cudaMemcpyAsync + stream synchronisation like the reference, rasterizer_impl.cu:281-282),
     *   calls binning_alloc / temp_alloc with the exact sizes and queues the rest.  The GPU never idles for the round
     *   trip and the outputs are always valid.
     *   SPECULATIVE path (spec_binning_buffer != NULL; needs spec_hist_buffer and num_rendered_pinned): no host wait at
     *   all.  The remaining stages are queued against the caller's buffers (writes and reads are clamped to them), the
     *   library records r_ready_event behind the kernel that stores {R, V}; the caller waits on it later and, if
     *   gvd_raster_binning_bytes(R, export_keys) > spec_binning_bytes or gvd_raster_hist_bytes(V,W,H) > spec_hist_bytes,
     *   the outputs are invalid and the call must be repeated.  num_rendered / num_visible are -1 on this path. */
    void* spec_binning_buffer;
    size_t spec_binning_bytes;
    void* spec_hist_buffer;
    size_t spec_hist_bytes;
    int* num_rendered_pinned;
    void* r_ready_event;
    /* result */
    int num_rendered;            /* out: R = number of (Gaussian,tile) instances     */
    int num_visible;             /* out: V = number of Gaussians with radii > 0      */
    /* SURVEY 8 row f3 -- the activations of gaussian_renderer.render() folded into the kernels.  raw_params = 1:
     * `scales`, `rotations` and `opacities` are the UN-activated GaussianModel parameters (_scaling, _rotation, _opacity;
     * scene/gaussian_model.py:36-43,106-130) and the kernels apply exp / normalize / sigmoid themselves; `shs` is
     * _features_dc [P,1,3] and `shs_rest` _features_rest [P,M-1,3] -- the [P,M,3] torch.cat (96 MB at 500 k Gaussians) and
     * the four activation launches of every render() call disappear.  Needs scales + rotations + shs (no precomputed
     * covariance / colours). */
    int raw_params;
    const float* shs_rest;       /* [P,M-1,3] when raw_params                        */
} GvdRasterForwardArgs;

typedef struct GvdRasterBackwardArgs {
    int P, D, M, R;
    int num_visible;             /* V from the forward, or -1 when the host does not know it (costs idle CTAs) */
    int width, height;
    /* forward inputs again (dev) -- DGR/rasterize_points.cu:121-146 */
    const float* background;
    const float* means3D;
    const float* shs;
    const float* colors_precomp;
    const float* scales;
    const float* rotations;
    const float* cov3D_precomp;
    const float* viewmatrix;
    const float* projmatrix;
    const float* campos;
    float scale_modifier;
    float tan_fovx, tan_fovy;
    const int*   radii;          /* [P]     from forward                          */
    const float* alphas;         /* [1,H,W] out_alpha from forward                */
    /* forward scratch (dev) */
    const void* geom_buffer;
    const void* binning_buffer;
    const void* img_buffer;
    /* incoming cotangents (dev) */
    const float* dL_dpix;        /* [3,H,W] */
    const float* dL_ddepth_pix;  /* [1,H,W] */
    const float* dL_dalpha_pix;  /* [1,H,W] */
    /* per-Gaussian confidence, fused: DGR/diff_gaussian_rasterization/__init__.py:147-157
     * multiplies every returned gradient except dL_dmeans2D by confidence[P,1]. NULL = ones. */
    const float* confidence;     /* [P] or NULL */
    /* scratch: zero-filled by this call; gvd_raster_backward_scratch_bytes(P) bytes */
    void* scratch;
    /* Optional: one region that contains ALL the gradient outputs below (e.g. a flat allocation the eight tensors are
     * views of).  The outputs are zero except on the V visible Gaussians; given the region, its zero-fill is folded
     * into the (issue-bound) render kernel instead of costing one streaming launch per output. */
    void* zero_region;
    size_t zero_region_bytes;
    /* outputs (dev); all fully written by this call (no pre-zeroing needed).      */
    float* dL_dmeans2D;          /* [P,3]  (z = 0), NOT confidence-scaled          */
    float* dL_dmeans3D;          /* [P,3]                                          */
    float* dL_dopacity;          /* [P]                                            */
    float* dL_dcolors;           /* [P,3]   or NULL (needed iff colors_precomp)    */
    float* dL_dcov3D;            /* [P,6]   or NULL (needed iff cov3D_precomp)     */
    float* dL_dsh;               /* [P,M,3] or NULL (needed iff shs)               */
    float* dL_dscales;           /* [P,3]   or NULL (needed iff scales)            */
    float* dL_drotations;        /* [P,4]   or NULL (needed iff rotations)         */
    int debug;
    /* raw_params = 1 (see GvdRasterForwardArgs): scales / rotations are the raw parameters, `opacities` the raw opacity
     * [P]; the gradients come back with respect to the RAW parameters (chain rule of exp / normalize / sigmoid applied in
     * the kernel), dL_dsh is [P,1,3] (d/d_features_dc) and dL_dsh_rest [P,M-1,3] (d/d_features_rest). */
    int raw_params;
    const float* shs_rest;
    const float* opacities;
    float* dL_dsh_rest;
} GvdRasterBackwardArgs;

/* sizes of the caller-owned scratch buffers; replaces
 * CudaRasterizer::required<GeometryState|BinningState|ImageState>
 * (DGR/cuda_rasterizer/rasterizer_impl.h:63-69). */
GVD_API size_t gvd_raster_geom_bytes(int P, int width, int height);
GVD_API size_t gvd_raster_binning_bytes(int R, int export_keys);
GVD_API size_t gvd_raster_img_bytes(int width, int height);
GVD_API size_t gvd_raster_sort_bytes(int P);                                  /* forward-only scratch, stage 1 */
GVD_API size_t gvd_raster_hist_bytes(int num_visible, int width, int height);  /* forward-only scratch, stage 2 */
GVD_API size_t gvd_raster_backward_scratch_bytes(int P);

/* Replaces CudaRasterizer::Rasterizer::forward (DGR/cuda_rasterizer/rasterizer_impl.cu:197-339)
 * as reached from RasterizeGaussiansCUDA (DGR/rasterize_points.cu:35-119). */
GVD_API int gvd_raster_forward(GvdRasterForwardArgs* args, gvd_stream_t stream);

/* Replaces CudaRasterizer::Rasterizer::backward (rasterizer_impl.cu:343-447) as reached
 * from RasterizeGaussiansBackwardCUDA (rasterize_points.cu:121-208), plus the Python-side
 * confidence scaling (diff_gaussian_rasterization/__init__.py:147-157). */
GVD_API int gvd_raster_backward(const GvdRasterBackwardArgs* args, gvd_stream_t stream);

/* Replaces CudaRasterizer::Rasterizer::markVisible (rasterizer_impl.cu:141-153):
 * present[i] = (view-space z of means3D[i] > 0.2).  present: dev uint8[P]. */
GVD_API int gvd_raster_mark_visible(int P, const float* means3D, const float* viewmatrix,
                            const float* projmatrix, uint8_t* present, gvd_stream_t stream);

/* Introspection for parity tests: byte offsets of the bit-exact-comparable arrays inside
 * the scratch buffers of THIS library (the reference's own layout is
 * rasterizer_impl.cu:155-195). */
typedef struct GvdRasterLayout {
    size_t geom_splat;          /* float4[4P]: {x,y,conA,conB},{conC,opac,r,g},{b,depth,hx,hy},{rect_lo,rect_hi,0,0} */
    size_t geom_clamped;        /* uint8[P]: bit c set = channel c was clamped at 0                          */
    size_t geom_tiles_touched;  /* uint32[P]                                                                 */
    size_t geom_visible_ids;    /* uint32[P] ids of the V visible Gaussians, ascending (V entries valid)       */
    size_t geom_counts;         /* uint32[8] {V, R, ...}                                                     */
    size_t bin_point_list;      /* uint32[R] sorted Gaussian ids (tile-major, depth order inside a tile)     */
    size_t bin_point_list_keys; /* uint64[R] sorted keys (tile<<32 | depth bits); only with export_keys      */
    size_t img_ranges;          /* uint2[T]                                                                  */
    size_t img_n_contrib;       /* uint32[H*W]                                                               */
} GvdRasterLayout;
GVD_API int gvd_raster_layout(int P, int R, int width, int height, GvdRasterLayout* out);

/* Optional per-stage device timing (CUDA events recorded on the caller's stream around each stage).
 * Off by default; used by bench.py for the live roofline numbers. Not thread-safe; single stream. */
enum {
    GVD_STAGE_PREPROCESS = 0, GVD_STAGE_SCAN /* bin count+prefix+ranges */, GVD_STAGE_EMIT /* bin fill */,
    GVD_STAGE_SORT /* compaction + depth sort */, GVD_STAGE_PACK /* export_keys (tests only) */,
    GVD_STAGE_RENDER_FWD, GVD_STAGE_RENDER_BWD, GVD_STAGE_GAUSSIAN_BWD, GVD_STAGE_COUNT
};
typedef struct GvdRasterStageTimes {
    double ms[GVD_STAGE_COUNT];   /* accumulated device milliseconds per stage        */
    int    calls[GVD_STAGE_COUNT];/* number of timed launches accumulated per stage   */
} GvdRasterStageTimes;
GVD_API int gvd_raster_profile_enable(int on);                   /* resets the accumulators   */
GVD_API int gvd_raster_profile_read(GvdRasterStageTimes* out);   /* syncs pending events      */

GVD_API int gvd_raster_abi_version(void);
GVD_API const char* gvd_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* GVD_RASTER_H_ */
