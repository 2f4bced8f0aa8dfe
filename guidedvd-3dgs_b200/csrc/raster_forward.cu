// raster_forward.cu -- forward kernels of the B200-native Gaussian rasterizer (sm_100a).
//
//   preprocess_kernel     per Gaussian: cull, project, EWA cov2D, conic, radius, tile rect, SH->RGB
//                         (follows DGR/cuda_rasterizer/forward.cu:155-256 arithmetic exactly so that
//                         radii / tile rects / depth bits are bit-identical to the reference), plus the
//                         alpha>=1/255 extent used for sub-tile culling
//   bin_count / bin_prefix / bin_ranges / bin_fill
//                         rect-aware stable counting sort of the (Gaussian, tile) instances on the tile id,
//                         run over the depth-sorted Gaussians; produces point_list + ranges exactly as the
//                         reference's duplicateWithKeys + 64-bit radix sort + identifyTileRanges
//                         (rasterizer_impl.cu:70-138,304-319) without materialising keys
//   render_forward_kernel SPLIT CTAs per 16x16 tile (default 2: 16x8 pixels, 4 warps each): ids staged by TMA bulk
//                         copy (cp.async.bulk + mbarrier), records gathered from the L2-resident per-Gaussian
//                         array one batch ahead; each warp owns an 8x4 pixel sub-tile and visits only the
//                         instances that can touch it
//                         (forward.cu:261-381 semantics preserved: skipped instances are exactly those every
//                         pixel of the warp would `continue` on).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include "raster_common.cuh"
#include "../../include/gvd_raster.h"

namespace {

// ------------------------------------------------------------------------------------------
// forward.cu:20-71
__device__ __forceinline__ float3 sh_to_rgb(int idx, int deg, int max_coeffs, const float3 pos, const float3 campos,
                                            const float* __restrict__ shs, const float* __restrict__ shs_rest,
                                            uint8_t* clamped_bits) {
    float3 dir = f3_sub(pos, campos);
    float len = sqrtf(f3_dot(dir, dir));
    dir = {dir.x / len, dir.y / len, dir.z / len};

    float v[48];
    if (shs_rest != nullptr) load_sh_split(shs, shs_rest, idx, deg, max_coeffs, v);  // raw parameters: dc | rest
    else load_sh(shs, idx, deg, max_coeffs, v);
#define SH(k) make_float3(v[3 * (k)], v[3 * (k) + 1], v[3 * (k) + 2])
    float3 result = f3_scale(GVD_SH_C0, SH(0));
    if (deg > 0) {
        float x = dir.x, y = dir.y, z = dir.z;
        result = f3_sub(f3_add(f3_sub(result, f3_scale(GVD_SH_C1 * y, SH(1))), f3_scale(GVD_SH_C1 * z, SH(2))),
                        f3_scale(GVD_SH_C1 * x, SH(3)));
        if (deg > 1) {
            float xx = x * x, yy = y * y, zz = z * z;
            float xy = x * y, yz = y * z, xz = x * z;
            result = f3_add(
                f3_add(f3_add(f3_add(f3_add(result, f3_scale(GVD_SH_C2_0 * xy, SH(4))), f3_scale(GVD_SH_C2_1 * yz, SH(5))),
                              f3_scale(GVD_SH_C2_2 * (2.0f * zz - xx - yy), SH(6))),
                       f3_scale(GVD_SH_C2_3 * xz, SH(7))),
                f3_scale(GVD_SH_C2_4 * (xx - yy), SH(8)));
            if (deg > 2) {
                result = f3_add(
                    f3_add(f3_add(f3_add(f3_add(f3_add(f3_add(result, f3_scale(GVD_SH_C3_0 * y * (3.0f * xx - yy), SH(9))),
                                                       f3_scale(GVD_SH_C3_1 * xy * z, SH(10))),
                                                f3_scale(GVD_SH_C3_2 * y * (4.0f * zz - xx - yy), SH(11))),
                                         f3_scale(GVD_SH_C3_3 * z * (2.0f * zz - 3.0f * xx - 3.0f * yy), SH(12))),
                                  f3_scale(GVD_SH_C3_4 * x * (4.0f * zz - xx - yy), SH(13))),
                           f3_scale(GVD_SH_C3_5 * z * (xx - yy), SH(14))),
                    f3_scale(GVD_SH_C3_6 * x * (xx - 3.0f * yy), SH(15)));
            }
        }
    }
#undef SH
    result.x += 0.5f;
    result.y += 0.5f;
    result.z += 0.5f;
    *clamped_bits = (uint8_t)((result.x < 0 ? 1 : 0) | (result.y < 0 ? 2 : 0) | (result.z < 0 ? 4 : 0));
    return {fmaxf(result.x, 0.0f), fmaxf(result.y, 0.0f), fmaxf(result.z, 0.0f)};
}

// forward.cu:74-113
__device__ __forceinline__ float3 cov2d_from_cov3d(const float3& mean, float focal_x, float focal_y, float tan_fovx,
                                                   float tan_fovy, const float* cov3D,
                                                   const float* __restrict__ viewmatrix) {
    float3 t = xform_point_4x3(mean, viewmatrix);
    const float limx = 1.3f * tan_fovx;
    const float limy = 1.3f * tan_fovy;
    const float txtz = t.x / t.z;
    const float tytz = t.y / t.z;
    t.x = fminf(limx, fmaxf(-limx, txtz)) * t.z;
    t.y = fminf(limy, fmaxf(-limy, tytz)) * t.z;

    M3 J = m3_make(focal_x / t.z, 0.0f, -(focal_x * t.x) / (t.z * t.z), 0.0f, focal_y / t.z,
                   -(focal_y * t.y) / (t.z * t.z), 0, 0, 0);
    M3 W = m3_make(viewmatrix[0], viewmatrix[4], viewmatrix[8], viewmatrix[1], viewmatrix[5], viewmatrix[9],
                   viewmatrix[2], viewmatrix[6], viewmatrix[10]);
    M3 T = m3_mul(W, J);
    M3 Vrk = m3_make(cov3D[0], cov3D[1], cov3D[2], cov3D[1], cov3D[3], cov3D[4], cov3D[2], cov3D[4], cov3D[5]);
    M3 cov = m3_mul(m3_mul(m3_transpose(T), m3_transpose(Vrk)), T);
    cov.m[0][0] += 0.3f;
    cov.m[1][1] += 0.3f;
    return {cov.m[0][0], cov.m[0][1], cov.m[1][1]};
}

// auxiliary.h:46-56 (float arithmetic, truncation toward zero, clamp to [0, grid])
__device__ __forceinline__ void tile_rect(const float2 p, int max_radius, uint2& rect_min, uint2& rect_max, dim3 grid) {
    rect_min = {min(grid.x, max((int)0, (int)((p.x - max_radius) / GVD_TILE_X))),
                min(grid.y, max((int)0, (int)((p.y - max_radius) / GVD_TILE_Y)))};
    rect_max = {min(grid.x, max((int)0, (int)((p.x + max_radius + GVD_TILE_X - 1) / GVD_TILE_X))),
                min(grid.y, max((int)0, (int)((p.y + max_radius + GVD_TILE_Y - 1) / GVD_TILE_Y)))};
}

// Half extents of the axis-aligned bound of { d : opac * exp(power(d)) >= 1/255 } (power = -q(d)/2 with the
// conic as quadratic form), inflated by a slack that dwarfs any rounding in the per-pixel evaluation.
// -1e30: can never contribute (opac < 1/255: alpha <= opac).  +1e30: no usable bound.
// The extents are those of the level set of the ROUNDED conic the render kernels evaluate, sqrt(t * C / (A C - B^2)) and
// sqrt(t * A / (A C - B^2)); the determinant is formed in double (products of floats are exact there). In fp32 it
// cancels for elongated, rotated footprints (A C ~ B^2): round 1's fp32 version under-estimated the extents of such
// Gaussians and dropped a few of their faintest contributions -- invisible at 1e-4 in the images of ordinary scenes,
// 9e-4 in the gradients of a scene of large anisotropic Gaussians (tests/test_render_dropin_gpu.py).
__device__ __forceinline__ void alpha_extent(const float3 conic, float opac, float& hx, float& hy) {
    hx = hy = 1e30f;
    if (opac * 255.0f < 0.999f) {
        hx = hy = -1e30f;
        return;
    }
    const double A = conic.x, B = conic.y, C = conic.z;
    const float t = 2.0f * logf(opac * 255.0f) * 1.0005f + 1e-3f;  // q(d) <= t  <=>  power >= -t/2
    const double det = A * C - B * B;
    if (!(det > 0.0) || !(A > 0.0) || !(C > 0.0) || !(t >= 0.0f)) return;
    const float ex = (float)sqrt((double)t * C / det) * 1.0005f + 0.02f;
    const float ey = (float)sqrt((double)t * A / det) * 1.0005f + 0.02f;
    if (!(ex < 1e9f) || !(ey < 1e9f)) return;
    hx = ex;
    hy = ey;
}

// One Gaussian of preprocessCUDA (forward.cu:155-256). Returns the number of tiles its rect covers; 0 = not rendered
// (near-culled, degenerate covariance or empty rect), in which case nothing but radii = 0 has been written.
template <bool RAW>  // RAW: un-activated GaussianModel parameters (compile-time, so the standard instantiation keeps the
                      // reference's expression trees -- and with them nvcc's FMA contraction -- untouched)
__device__ __forceinline__ uint32_t preprocess_one(
    int idx, int D, int M, const float* __restrict__ orig_points, const float3* __restrict__ scales,
    const float scale_modifier, const float4* __restrict__ rotations, const float* __restrict__ opacities,
    const float* __restrict__ shs, uint8_t* __restrict__ clamped, const float* __restrict__ cov3D_precomp,
    const float* __restrict__ colors_precomp, const float* __restrict__ viewmatrix,
    const float* __restrict__ projmatrix, const float3* __restrict__ cam_pos, const int W, int H,
    const float tan_fovx, float tan_fovy, const float focal_x, float focal_y, int* __restrict__ radii,
    SplatRec* __restrict__ splat, const dim3 grid, uint32_t* __restrict__ depth_key, int prefiltered, int no_cull,
    const float* __restrict__ shs_rest) {
    radii[idx] = 0;

    // near cull (auxiliary.h:139-164): only p_view.z <= 0.2 rejects.
    const float3 p_orig = {orig_points[3 * idx], orig_points[3 * idx + 1], orig_points[3 * idx + 2]};
    const float3 p_view = xform_point_4x3(p_orig, viewmatrix);
    if (p_view.z <= 0.2f) {
        if (prefiltered) {
            printf("Point is filtered although prefiltered is set. This shouldn't happen!");
            __trap();
        }
        return 0;
    }

    const float4 p_hom = xform_point_4x4(p_orig, projmatrix);
    const float p_w = 1.0f / (p_hom.w + 0.0000001f);
    const float3 p_proj = {p_hom.x * p_w, p_hom.y * p_w, p_hom.z * p_w};

    float cov3D_local[6];
    const float* cov3D;
    if (cov3D_precomp != nullptr) {
        cov3D = cov3D_precomp + (size_t)idx * 6;
    } else {
        // RAW: the GaussianModel activations happen here (gaussian_renderer/__init__.py:60-87 does them as torch launches)
        if (RAW) cov3d_from_scale_rot(act_scale(scales[idx]), scale_modifier, act_rotation(rotations[idx]), cov3D_local);
        else cov3d_from_scale_rot(scales[idx], scale_modifier, rotations[idx], cov3D_local);
        cov3D = cov3D_local;
    }

    const float3 cov = cov2d_from_cov3d(p_orig, focal_x, focal_y, tan_fovx, tan_fovy, cov3D, viewmatrix);

    const float det = (cov.x * cov.z - cov.y * cov.y);
    if (det == 0.0f) return 0;
    const float det_inv = 1.f / det;
    const float3 conic = {cov.z * det_inv, -cov.y * det_inv, cov.x * det_inv};

    const float mid = 0.5f * (cov.x + cov.z);
    const float lambda1 = mid + sqrtf(fmaxf(0.1f, mid * mid - det));
    const float lambda2 = mid - sqrtf(fmaxf(0.1f, mid * mid - det));
    const float my_radius = ceilf(3.f * sqrtf(fmaxf(lambda1, lambda2)));
    const float2 point_image = {ndc_to_pix(p_proj.x, W), ndc_to_pix(p_proj.y, H)};
    uint2 rect_min, rect_max;
    tile_rect(point_image, (int)my_radius, rect_min, rect_max, grid);
    if ((rect_max.x - rect_min.x) * (rect_max.y - rect_min.y) == 0) return 0;

    float3 rgb;
    if (colors_precomp == nullptr) {
        uint8_t cl;
        rgb = sh_to_rgb(idx, D, M, p_orig, *cam_pos, shs, RAW ? shs_rest : nullptr, &cl);
        clamped[idx] = cl;
    } else {
        rgb = {colors_precomp[3 * idx], colors_precomp[3 * idx + 1], colors_precomp[3 * idx + 2]};
    }

    radii[idx] = (int)my_radius;
    depth_key[idx] = __float_as_uint(p_view.z);

    SplatRec rec;
    rec.a = make_float4(point_image.x, point_image.y, conic.x, conic.y);
    const float opac = RAW ? act_opacity(opacities[idx]) : opacities[idx];
    rec.b = make_float4(conic.z, opac, rgb.x, rgb.y);
    float hx, hy;
    alpha_extent(conic, opac, hx, hy);
    if (no_cull) hx = hy = 1e30f;  // GVD_NO_SUBTILE_CULL=1: every warp visits every instance of its tile (A/B and diagnosis knob)
    rec.c = make_float4(rgb.z, p_view.z, hx, hy);
    rec.d = make_float4(__uint_as_float(rect_min.x | (rect_min.y << 16)), __uint_as_float(rect_max.x | (rect_max.y << 16)),
                        0.f, 0.f);
    splat[idx] = rec;
    return (rect_max.y - rect_min.y) * (rect_max.x - rect_min.x);
}

// Per Gaussian: preprocess_one. Per CTA: how many of its Gaussians are rendered and how many (Gaussian, tile) instances
// they make -- the compaction kernel turns these into offsets, V and R (R is therefore known ~40 us into the frame,
// long before the instance list is needed). On the side the grid clears the depth sort's digit histograms.
template <bool RAW>
__global__ void __launch_bounds__(GVD_PRE_BLOCK) preprocess_kernel(
    int P, int D, int M, const float* __restrict__ orig_points, const float3* __restrict__ scales,
    const float scale_modifier, const float4* __restrict__ rotations, const float* __restrict__ opacities,
    const float* __restrict__ shs, uint8_t* __restrict__ clamped, const float* __restrict__ cov3D_precomp,
    const float* __restrict__ colors_precomp, const float* __restrict__ viewmatrix,
    const float* __restrict__ projmatrix, const float3* __restrict__ cam_pos, const int W, int H,
    const float tan_fovx, float tan_fovy, const float focal_x, float focal_y, int* __restrict__ radii,
    SplatRec* __restrict__ splat, const dim3 grid, uint32_t* __restrict__ tiles_touched,
    uint32_t* __restrict__ depth_key, uint32_t* __restrict__ blk_vis, uint32_t* __restrict__ blk_tiles,
    uint32_t* __restrict__ zeroed, size_t zeroed_words, int prefiltered, int no_cull,
    const float* __restrict__ shs_rest) {
    pdl_wait();
    pdl_trigger();
    __shared__ uint32_t warp_tiles[GVD_PRE_BLOCK / 32];
    const int idx = blockIdx.x * GVD_PRE_BLOCK + threadIdx.x;
    for (size_t i = (size_t)idx; i < zeroed_words; i += (size_t)gridDim.x * GVD_PRE_BLOCK) zeroed[i] = 0u;
    uint32_t tiles = 0;
    if (idx < P) {
        tiles = preprocess_one<RAW>(idx, D, M, orig_points, scales, scale_modifier, rotations, opacities, shs, clamped, cov3D_precomp,
                               colors_precomp, viewmatrix, projmatrix, cam_pos, W, H, tan_fovx, tan_fovy, focal_x, focal_y, radii,
                               splat, grid, depth_key, prefiltered, no_cull, shs_rest);
        tiles_touched[idx] = tiles;
    }
    const uint32_t wsum = __reduce_add_sync(0xffffffffu, tiles);
    if ((threadIdx.x & 31) == 0) warp_tiles[threadIdx.x >> 5] = wsum;
    const int nvis = __syncthreads_count(tiles > 0);
    if (threadIdx.x == 0) {
        uint32_t t = 0;
#pragma unroll
        for (int w = 0; w < GVD_PRE_BLOCK / 32; ++w) t += warp_tiles[w];
        blk_vis[blockIdx.x] = (uint32_t)nvis;
        blk_tiles[blockIdx.x] = t;
    }
}

// ------------------------------------------------------------------------------------------
// ---- compaction of the visible Gaussians + depth sort ---------------------------------------------

// Block sum for up to 1024 threads; warp_buf[32]. The result is valid in every thread of warp 0.
__device__ __forceinline__ uint32_t block_sum(uint32_t v, uint32_t* warp_buf) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = __reduce_add_sync(0xffffffffu, v);
    if (lane == 0) warp_buf[warp] = v;
    __syncthreads();
    uint32_t r = 0;
    if (warp == 0) r = __reduce_add_sync(0xffffffffu, lane < nw ? warp_buf[lane] : 0u);
    return r;
}

// CTA b owns the ids of preprocess CTA b: its visible ones go, in id order, to the compacted positions base_b + rank
// (base_b = visible Gaussians of all earlier preprocess CTAs, summed from the per-CTA counts). Writes the (depth bits,
// id) pairs the sort starts from, vis_id (kept for the backward) and the four 8-bit digit histograms of all keys
// (warp-aggregated with MATCH.ANY first: the top byte of a depth takes a handful of values, plain shared-memory atomics
// on it serialise); the last CTA publishes V and R (device words and, if given, pinned host words).
// 256-thread CTAs: a first version with 1024-thread CTAs (two resident per SM, six block barriers each) took 25 us at
// C2 -- a chain of barrier waits, not work (ncu: 13 % issue slots used).
__global__ void __launch_bounds__(GVD_COMPACT_BLOCK) compact_kernel(
    int P, int nb, const uint32_t* __restrict__ tiles_touched, const uint32_t* __restrict__ depth_key,
    const uint32_t* __restrict__ blk_vis, const uint32_t* __restrict__ blk_tiles, uint32_t* __restrict__ key0,
    uint32_t* __restrict__ val0, uint32_t* __restrict__ vis_id, uint32_t* __restrict__ counts, uint32_t* ghist, int* r_host) {
    pdl_wait();
    pdl_trigger();
    static_assert(GVD_COMPACT_BLOCK == GVD_PRE_BLOCK && GVD_COMPACT_BLOCK == 256, "one compaction CTA per preprocess CTA");
    __shared__ uint32_t hist_s[4 * 256];
    __shared__ uint32_t warp_buf[32];
    __shared__ uint32_t warp_cnt[GVD_COMPACT_BLOCK / 32];
    __shared__ uint32_t s_base, s_total;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int k = 0; k < 4; ++k) hist_s[k * 256 + tid] = 0u;
    uint32_t s = 0;
    for (int k = (int)tid; k < (int)blockIdx.x; k += GVD_COMPACT_BLOCK) s += blk_vis[k];
    const int idx = (int)(blockIdx.x * GVD_COMPACT_BLOCK + tid);
    const bool vis = idx < P && tiles_touched[idx] > 0;
    const uint32_t key = vis ? depth_key[idx] : 0u;
    const uint32_t ballot = __ballot_sync(0xffffffffu, vis);
    if (lane == 0) warp_cnt[warp] = __popc(ballot);
    const uint32_t base_w0 = block_sum(s, warp_buf);  // contains a __syncthreads: warp_cnt and hist_s = 0 are visible below
    if (warp == 0) {
        const uint32_t c = lane < GVD_COMPACT_BLOCK / 32 ? warp_cnt[lane] : 0u;
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (uint32_t)o) incl += u;
        }
        if (lane < GVD_COMPACT_BLOCK / 32) warp_cnt[lane] = incl - c;
        if (lane == 31) s_total = incl;
        if (lane == 0) s_base = base_w0;
    }
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
        const uint32_t d = vis ? ((key >> (8 * pass)) & 255u) : (256u + lane);  // invisible lanes match nobody
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        if (vis && (peers & lt_mask) == 0u) atomicAdd(&hist_s[256 * pass + d], (uint32_t)__popc(peers));
    }
    __syncthreads();
    const uint32_t base = s_base;
    if (vis) {
        const uint32_t pos = base + warp_cnt[warp] + __popc(ballot & lt_mask);
        key0[pos] = key;
        val0[pos] = (uint32_t)idx;
        vis_id[pos] = (uint32_t)idx;
    }
    // GVD_GHIST_COPIES replicas of the global histogram spread the CTAs' atomics; the sort passes add the replicas up
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t c = hist_s[k * 256 + tid];
        if (c) atomicAdd(&ghist[(blockIdx.x % GVD_GHIST_COPIES) * 1024 + k * 256 + tid], c);
    }
    if (blockIdx.x == gridDim.x - 1) {
        uint32_t r = 0;
        for (int k = (int)tid; k < nb; k += GVD_COMPACT_BLOCK) r += blk_tiles[k];
        __syncthreads();  // warp_buf is reused
        const uint32_t R = block_sum(r, warp_buf);
        if (tid == 0) {
            const uint32_t V = base + s_total;
            counts[0] = V;
            counts[1] = R;
            counts[7] = 0u;  // error word of the depth sort (a flag wait that gave up)
            // R (and V) go straight into the caller's pinned, device-mapped host words: a cudaMemcpyAsync in the compute
            // stream queues behind whatever the copy engines are busy with (measured +46 us per step in round 1).
            if (r_host != nullptr) {
                reinterpret_cast<volatile int*>(r_host)[0] = (int)R;
                reinterpret_cast<volatile int*>(r_host)[1] = (int)V;
                __threadfence_system();
            }
        }
    }
}

// One pass of the stable LSD radix sort: a CTA ranks a tile of 1024 keys on the digit (key >> shift) & 255.
//   tile id : a ticket (atomic counter), so every predecessor of a tile has started before it -- the waits below cannot
//             deadlock whatever the order in which the hardware dispatches CTAs
//   ranking : warp w owns 128 consecutive keys, 4 rounds of 32 (lane = consecutive index, so lane order = input order);
//             same-digit lanes of a round are ranked with MATCH.ANY, rounds and warps through per-warp counters wh[w][d]
//   offsets : thread = digit d. The tile publishes its count of d (flagged word), then needs the number of keys with
//             digit d in all earlier tiles: the flagged counts of the <= 31 predecessors in its group of 32 tiles,
//             fetched with independent loads (re-fetched until their flags are up: all tiles of a wave publish at about
//             the same time), plus the running total the last tile of the previous group published. Keys with smaller
//             digits come from the global digit histogram the compaction kernel made.
// No per-key atomics, no histogram kernel, no chain of dependent round trips longer than the number of groups.
// A first version let each pass accumulate the next pass's per-tile histograms with two RED.ADD per key while it
// scattered: 18-30 us per pass at C2 against 7 us for the last pass, which had none (profiles/r02_ncu_binning_*).
#define GVD_SORT_FLAG 0x80000000u
#define GVD_SORT_SPIN_LIMIT (1u << 22)
#ifdef GVD_HOST_EMU
#define gvd_nanosleep(ns) emu_yield()
#else
#define gvd_nanosleep(ns) __nanosleep(ns)
#endif
__device__ __forceinline__ uint32_t ld_flagged(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }

__global__ void __launch_bounds__(GVD_SORT_THREADS) sort_pass_kernel(
    const uint32_t* __restrict__ counts, const uint32_t* __restrict__ kin, const uint32_t* __restrict__ vin,
    uint32_t* __restrict__ kout, uint32_t* __restrict__ vout, const uint32_t* __restrict__ ghist, uint32_t* agg,
    uint32_t* incl, uint32_t* ticket, uint32_t* error_word, int shift) {
    pdl_wait();
    pdl_trigger();
    constexpr int WARPS = GVD_SORT_THREADS / 32, PER_WARP = GVD_SORT_TILE / WARPS, ROUNDS = PER_WARP / 32;
    static_assert(GVD_SORT_THREADS == 256, "one thread per digit in the offset phase");
    __shared__ uint32_t wh[WARPS][256];
    __shared__ uint32_t warp_tot[WARPS];
    __shared__ uint32_t s_tile;
    const uint32_t V = counts[0];
    if (blockIdx.x * GVD_SORT_TILE >= V) return;  // as many tickets are drawn as there are tiles with keys
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
#pragma unroll
    for (int w = 0; w < WARPS; ++w) wh[w][tid] = 0u;
    __syncthreads();
    const uint32_t tile = s_tile, start = tile * GVD_SORT_TILE;
    const uint32_t n = min((uint32_t)GVD_SORT_TILE, V - start);
    const uint32_t lt_mask = (1u << lane) - 1u;

    uint32_t key[ROUNDS], val[ROUNDS];
    bool valid[ROUNDS];
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        const uint32_t off = warp * PER_WARP + r * 32 + lane;
        valid[r] = off < n;
        key[r] = valid[r] ? kin[start + off] : 0u;
        val[r] = valid[r] ? vin[start + off] : 0u;
    }
    uint32_t g = 0;
#pragma unroll
    for (int c = 0; c < GVD_GHIST_COPIES; ++c) g += ghist[c * 1024 + tid];
    // phase A: per-warp digit counts
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        const uint32_t d = valid[r] ? ((key[r] >> shift) & 255u) : (256u + lane);  // invalid lanes match nobody
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        if (valid[r] && (peers & lt_mask) == 0u) wh[warp][d] += (uint32_t)__popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // phase B: thread = digit
    {
        uint32_t cnt = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) cnt += wh[w][tid];
        *reinterpret_cast<volatile uint32_t*>(agg + (size_t)tile * 256 + tid) = cnt | GVD_SORT_FLAG;

        uint32_t ginc = g;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, ginc, o);
            if (lane >= (uint32_t)o) ginc += u;
        }
        if (lane == 31) warp_tot[warp] = ginc;

        const uint32_t group = tile / GVD_SORT_SUPER, g0 = group * GVD_SORT_SUPER;
        uint32_t prior = 0, spins = 0;
        bool gave_up = false;
        for (uint32_t i = g0; i < tile;) {  // the predecessors inside the group, eight independent loads at a time
            uint32_t v[8];
            bool ready = true;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                v[k] = i + k < tile ? ld_flagged(agg + (size_t)(i + k) * 256 + tid) : GVD_SORT_FLAG;
                ready = ready && (v[k] & GVD_SORT_FLAG);
            }
            if (!ready) {
                if (++spins > GVD_SORT_SPIN_LIMIT) { gave_up = true; break; }
                gvd_nanosleep(40);
                continue;
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) prior += v[k] & ~GVD_SORT_FLAG;
            i += 8;
        }
        if (group > 0 && !gave_up) {  // everything before the group: the running total its predecessor group ended with
            uint32_t v;
            while (!((v = ld_flagged(incl + (size_t)(group - 1) * 256 + tid)) & GVD_SORT_FLAG)) {
                if (++spins > GVD_SORT_SPIN_LIMIT) { gave_up = true; break; }
                gvd_nanosleep(40);
            }
            prior += v & ~GVD_SORT_FLAG;
        }
        if (gave_up) *error_word = 1u;  // never hang the GPU: finish with wrong offsets and say so
        if (tile % GVD_SORT_SUPER == GVD_SORT_SUPER - 1)
            *reinterpret_cast<volatile uint32_t*>(incl + (size_t)group * 256 + tid) = (prior + cnt) | GVD_SORT_FLAG;
        __syncthreads();
        uint32_t run = ginc - g + prior;
        for (uint32_t w = 0; w < warp; ++w) run += warp_tot[w];
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            const uint32_t c = wh[w][tid];
            wh[w][tid] = run;
            run += c;
        }
    }
    __syncthreads();
    // phase C: stable scatter
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        const uint32_t d = valid[r] ? ((key[r] >> shift) & 255u) : (256u + lane);
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        uint32_t pos = 0;
        if (valid[r]) pos = wh[warp][d] + (uint32_t)__popc(peers & lt_mask);
        __syncwarp();
        if (valid[r] && (peers & lt_mask) == 0u) wh[warp][d] += (uint32_t)__popc(peers);
        __syncwarp();
        if (valid[r]) {
            kout[pos] = key[r];
            vout[pos] = val[r];
        }
    }
}

// ------------------------------------------------------------------------------------------
// ---- binning: rect-aware stable counting sort on the tile id -------------------------------------
// Instance order required (and produced by the reference's stable sort on tile<<32|depth): tile-major;
// inside a tile by depth, ties by Gaussian id. `order` lists the V visible Gaussians by (depth, id).

__device__ __forceinline__ void unpack_rect(const float4 d, uint32_t& x0, uint32_t& y0, uint32_t& x1, uint32_t& y1) {
    const uint32_t w0 = __float_as_uint(d.x), w1 = __float_as_uint(d.y);
    x0 = w0 & 0xffff; y0 = w0 >> 16; x1 = w1 & 0xffff; y1 = w1 >> 16;
}

// Pass 1: chunk c = GVD_BIN_CHUNK depth-consecutive Gaussians order[c*CHUNK ..]. hist[c][t] = how many of them cover tile t.
// A rect adds +1/-1 at its four corners of a (gy+1) x (gx+1) difference grid; a 2-D prefix sum then yields
// the per-tile counts. Cost per chunk is O(CHUNK + T) whatever the rect sizes (no per-instance atomics).
__global__ void __launch_bounds__(GVD_BIN_CHUNK) bin_count_kernel(const uint32_t* __restrict__ counts, uint32_t gx, uint32_t gy,
                                                                 const SplatRec* __restrict__ splat,
                                                                 const uint32_t* __restrict__ order,
                                                                 uint32_t* __restrict__ hist) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ int diff[];  // (gy+1) rows of stride ld
    const uint32_t V = counts[0];
    if (blockIdx.x * GVD_BIN_CHUNK >= V) return;  // only a speculatively sized grid has such CTAs
    const int ld = (int)(gx + 1) | 1;  // odd stride: column walks are bank-conflict free
    const int rows = (int)gy + 1, cols = (int)gx + 1;
    const uint32_t i = blockIdx.x * GVD_BIN_CHUNK + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    const bool live = i < V;
    if (live) unpack_rect(splat[order[i]].d, x0, y0, x1, y1);
    for (int t = threadIdx.x; t < rows * ld; t += GVD_BIN_CHUNK) diff[t] = 0;
    __syncthreads();
    if (live) {
        atomicAdd(&diff[y0 * ld + x0], 1);
        atomicAdd(&diff[y0 * ld + x1], -1);
        atomicAdd(&diff[y1 * ld + x0], -1);
        atomicAdd(&diff[y1 * ld + x1], 1);
    }
    __syncthreads();
    // prefix along x: one warp per row
    for (int y = warp; y < rows; y += GVD_BIN_CHUNK / 32) {
        int carry = 0;
        for (int xb = 0; xb < cols; xb += 32) {
            const int x = xb + lane;
            int v = (x < cols) ? diff[y * ld + x] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += u;
            }
            v += carry;
            if (x < cols) diff[y * ld + x] = v;
            carry = __shfl_sync(0xffffffffu, v, 31);
        }
    }
    __syncthreads();
    // prefix along y: one thread per column
    for (int x = threadIdx.x; x < cols; x += GVD_BIN_CHUNK) {
        int run = 0;
        for (int y = 0; y < rows; ++y) {
            run += diff[y * ld + x];
            diff[y * ld + x] = run;
        }
    }
    __syncthreads();
    uint32_t* row = hist + (size_t)blockIdx.x * (gx * gy);
    for (uint32_t t = threadIdx.x; t < gx * gy; t += GVD_BIN_CHUNK) row[t] = (uint32_t)diff[(t / gx) * ld + (t % gx)];
}

// Pass 2: per tile, exclusive prefix over the chunks, in place, and the tile's total.
// CTA = GVD_PFX_TILES tiles x GVD_PFX_SEGS chunk segments: every thread sums its segment, the segment sums are scanned
// through shared memory, then every thread rewrites its segment as running prefixes. 8 consecutive tiles = one 32-byte
// sector per hist row, and 150 CTAs at C2 (round 1's 32-tile CTAs left 110 of the 148 SMs idle: 17 us).
#define GVD_PFX_TILES 8
#define GVD_PFX_SEGS 128
__global__ void __launch_bounds__(GVD_PFX_TILES * GVD_PFX_SEGS) bin_prefix_kernel(int T, int rows_cap, const uint32_t* __restrict__ counts,
                                                                                uint32_t* hist, uint32_t* __restrict__ tile_total) {
    pdl_wait();
    pdl_trigger();
    __shared__ uint32_t seg_sum[GVD_PFX_SEGS][GVD_PFX_TILES + 1];
    const int tx = threadIdx.x % GVD_PFX_TILES, seg = threadIdx.x / GVD_PFX_TILES;
    const int nv = min(rows_cap, (int)((counts[0] + GVD_BIN_CHUNK - 1) / GVD_BIN_CHUNK));
    const int per = (nv + GVD_PFX_SEGS - 1) / GVD_PFX_SEGS;
    const int c0 = min(nv, seg * per), c1 = min(nv, c0 + per);
    const int t = blockIdx.x * GVD_PFX_TILES + tx;
    uint32_t sum = 0;
    if (t < T)
        for (int c = c0; c < c1; ++c) sum += hist[(size_t)c * T + t];
    seg_sum[seg][tx] = sum;
    __syncthreads();
    uint32_t run = 0;
    for (int k = 0; k < seg; ++k) run += seg_sum[k][tx];
    if (t < T) {
        for (int c = c0; c < c1; c += 8) {
            uint32_t v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = (c + k < c1) ? hist[(size_t)(c + k) * T + t] : 0u;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (c + k < c1) {
                    hist[(size_t)(c + k) * T + t] = run;
                    run += v[k];
                }
        }
        if (seg == GVD_PFX_SEGS - 1) {
            uint32_t tot = 0;
#pragma unroll 8
            for (int k = 0; k < GVD_PFX_SEGS; ++k) tot += seg_sum[k][tx];
            tile_total[t] = tot;
        }
    }
}

// Pass 3: exclusive scan over the tiles -> ranges (rasterizer_impl.cu:116-138 semantics: empty tiles (0,0)).
// One CTA; T <= GVD_MAX_TILES. The scan total equals counts[1] (R) unless a speculative hist buffer was too small.
__global__ void __launch_bounds__(1024) bin_ranges_kernel(int T, const uint32_t* __restrict__ tile_total,
                                                          uint2* __restrict__ ranges) {
    pdl_wait();
    pdl_trigger();
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < T; base += 1024) {
        const int t = base + tid;
        const uint32_t v = (t < T) ? tile_total[t] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += u;
            }
            warp_sums[lane] = w;  // inclusive over warps
        }
        __syncthreads();
        const uint32_t carry = carry_s;
        const uint32_t start = carry + (warp ? warp_sums[warp - 1] : 0u) + incl - v;
        if (t < T) ranges[t] = v ? make_uint2(start, start + v) : make_uint2(0u, 0u);
        __syncthreads();
        if (tid == 1023) carry_s = carry + warp_sums[31];
        __syncthreads();
    }
}

// Pass 4: chunk c writes its Gaussians' ids into the tile lists. The slot of Gaussian j (0..63, depth order inside the
// chunk) in tile t is  ranges[t].x + hist[c][t] (the chunk's starting rank in t, from pass 2) + the number of earlier
// Gaussians of the chunk that cover t.  That last term comes from two 32-bit coverage masks per tile in shared memory:
//   phase 0  every (Gaussian, tile) instance of the chunk ORs bit j into mask[j / 32][t]   (native 32-bit ATOMS.OR)
//   phase 1  every instance reads its tile's masks back: rank = popc(bits below j); one 4-byte store.
// The work unit is a (Gaussian, tile row) pair, found by a binary search over the 64 row offsets: one thread per unit,
// walking along the row. No serial walk over the chunk, no ordering between instances. The per-tile bases (ranges +
// hist row) of the chunk's bounding box are staged in shared memory once, with full memory-level parallelism, so the
// inner loops touch shared memory only.
// Chunks whose Gaussians cover their bounding box more than GVD_FILL_DENSE times over (the Gaussians nearest to the
// camera: rects up to the whole screen, 40 000 instances in one chunk at C2) go tile-major instead: one thread per tile
// of the box tests the 64 rects (one broadcast 16-byte shared load each), builds the two masks in REGISTERS and emits
// the tile's ids in order -- no atomics, no barriers, and the chunk's work is spread evenly over the CTA whatever the
// rect sizes.
// History: round 1 walked each chunk with ONE warp, 32 instances at a time (load-balanced search + MATCH.ANY ranking,
// 8.3 warp-instructions per instance, 80 us at C2); tools/ubench_scatter_store.cu showed the 3.7 M scattered 4-byte stores
// themselves cost 25 us against 18.5 us for a coalesced stream, i.e. the walk was the cost, not the stores. Mask
// versions that gave a heavy chunk's (Gaussian, row) units to 8 warps took 171-207 us: 200 dependent iterations per
// warp in the heaviest chunk were the whole kernel (ncu: SMs active 49 % of the kernel's duration).
// The planes cover a BAND of tile rows (the whole image when 3 * T words fit the shared-memory budget: always at the
// benchmark sizes; larger images are walked band by band, restricted to the rows the chunk's Gaussians reach).
struct FillChunk {  // the chunk's Gaussians in depth order
    uint4 box[GVD_BIN_CHUNK];  // {x0, x1 - x0, y0, y1 - y0} in tiles
    uint32_t id[GVD_BIN_CHUNK], roff[GVD_BIN_CHUNK + 1];
    uint32_t wn[2], wr[2], wy0[2], wy1[2], wx0[2], wx1[2];
};

#define GVD_FILL_THREADS 256
#define GVD_FILL_DENSE 8
#define GVD_FILL_SMEM_BYTES (192 * 1024)  // shared-memory budget of the three planes
__device__ __forceinline__ uint32_t upper_slot(const uint32_t* off, uint32_t m, uint32_t q) {  // largest j < m with off[j] <= q
    uint32_t lo = 0, hi = m;  // invariant: off[lo] <= q < off[hi]
#pragma unroll
    for (int it = 0; it < 7; ++it) {
        if (hi - lo <= 1) break;
        const uint32_t mid = (lo + hi) >> 1;
        if (off[mid] <= q) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(GVD_FILL_THREADS) bin_fill_kernel(const uint32_t* __restrict__ counts, int T, uint32_t tiles_x,
                                                                   uint32_t band_rows, const SplatRec* __restrict__ splat,
                                                                   const uint32_t* __restrict__ order,
                                                                   const uint32_t* __restrict__ hist,
                                                                   const uint2* __restrict__ ranges,
                                                                   uint32_t* __restrict__ point_list, uint32_t capacity) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ uint32_t planes[];  // mask bits 0..31 | mask bits 32..63 | base slot, each [band_rows * tiles_x]
    __shared__ FillChunk fc;
    static_assert(GVD_BIN_CHUNK == 64, "two warps load the chunk; two mask words per tile");
    const uint32_t V = counts[0];
    if (blockIdx.x * GVD_BIN_CHUNK >= V) return;
    const uint32_t m = min((uint32_t)GVD_BIN_CHUNK, V - blockIdx.x * GVD_BIN_CHUNK);
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t* row = hist + (size_t)blockIdx.x * T;
    const uint32_t plane = band_rows * tiles_x;
    uint32_t* mask = planes;
    uint32_t* base = planes + 2 * plane;
    if (tid < GVD_BIN_CHUNK) {
        uint32_t n = 0, nr = 0, r0 = 0, r1 = 0, id = 0, ymin = 0xffffu, ymax = 0u, xmin = 0xffffu, xmax = 0u;
        if (tid < m) {
            id = order[blockIdx.x * GVD_BIN_CHUNK + tid];
            const float4 d = splat[id].d;
            r0 = __float_as_uint(d.x);
            r1 = __float_as_uint(d.y);
            ymin = r0 >> 16;
            ymax = r1 >> 16;
            xmin = r0 & 0xffff;
            xmax = r1 & 0xffff;
            nr = ymax - ymin;
            n = nr * (xmax - xmin);
        }
        uint32_t incl = n, rincl = nr;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o), ur = __shfl_up_sync(0xffffffffu, rincl, o);
            if (lane >= (uint32_t)o) { incl += u; rincl += ur; }
        }
        ymin = __reduce_min_sync(0xffffffffu, ymin);
        ymax = __reduce_max_sync(0xffffffffu, ymax);
        xmin = __reduce_min_sync(0xffffffffu, xmin);
        xmax = __reduce_max_sync(0xffffffffu, xmax);
        if (lane == 31) {
            fc.wn[warp] = incl; fc.wr[warp] = rincl;
            fc.wy0[warp] = ymin; fc.wy1[warp] = ymax; fc.wx0[warp] = xmin; fc.wx1[warp] = xmax;
        }
        fc.id[tid] = id;
        fc.box[tid] = tid < m ? make_uint4(r0 & 0xffff, (r1 & 0xffff) - (r0 & 0xffff), r0 >> 16, (r1 >> 16) - (r0 >> 16))
                              : make_uint4(0u, 0u, 0u, 0u);
        fc.roff[tid] = rincl - nr;  // warp 1's entries get warp 0's total added below
    }
    __syncthreads();
    if (tid >= 32 && tid < GVD_BIN_CHUNK) fc.roff[tid] += fc.wr[0];
    if (tid == 0) fc.roff[GVD_BIN_CHUNK] = fc.wr[0] + fc.wr[1];
    __syncthreads();
    const uint32_t total = fc.wn[0] + fc.wn[1], units = fc.roff[GVD_BIN_CHUNK];
    const uint32_t ylo = min(fc.wy0[0], fc.wy0[1]), yhi = max(fc.wy1[0], fc.wy1[1]);  // tile rows / columns the chunk reaches
    const uint32_t xlo = min(fc.wx0[0], fc.wx0[1]), xhi = max(fc.wx1[0], fc.wx1[1]);
    const uint32_t bbw = xhi - xlo;

    if (total > GVD_FILL_DENSE * bbw * (yhi - ylo)) {
        // tile-major: thread = tile of the bounding box
        const uint32_t cells = (yhi - ylo) * bbw;
        for (uint32_t c = tid; c < cells; c += blockDim.x) {
            const uint32_t cy = c / bbw, tx = xlo + (c - cy * bbw), ty = ylo + cy;
            const uint32_t t = ty * tiles_x + tx;
            uint32_t slot = ranges[t].x + row[t];
            uint32_t lo = 0u, hi = 0u;
#pragma unroll 8
            for (uint32_t j = 0; j < 32; ++j) {
                const uint4 bx = fc.box[j];
                lo |= (uint32_t)((tx - bx.x < bx.y) && (ty - bx.z < bx.w)) << j;
            }
#pragma unroll 8
            for (uint32_t j = 0; j < 32; ++j) {
                const uint4 bx = fc.box[32 + j];
                hi |= (uint32_t)((tx - bx.x < bx.y) && (ty - bx.z < bx.w)) << j;
            }
            while (lo) {
                const uint32_t j = (uint32_t)__ffs((int)lo) - 1u;
                lo &= lo - 1u;
                if (slot < capacity) point_list[slot] = fc.id[j];  // capacity: speculative buffers may be too small
                ++slot;
            }
            while (hi) {
                const uint32_t j = (uint32_t)__ffs((int)hi) - 1u;
                hi &= hi - 1u;
                if (slot < capacity) point_list[slot] = fc.id[32 + j];
                ++slot;
            }
        }
        return;
    }

    for (uint32_t b0 = ylo - ylo % band_rows; b0 < yhi; b0 += band_rows) {
        const uint32_t b1 = b0 + band_rows;
        {   // clear the masks and stage the bases of the bounding box inside this band
            const uint32_t y0 = max(b0, ylo), y1 = min(b1, yhi);
            const uint32_t cells = (y1 - y0) * bbw;
            for (uint32_t c = tid; c < cells; c += blockDim.x) {
                const uint32_t cy = c / bbw, cx = c - cy * bbw;
                const uint32_t t = (y0 + cy) * tiles_x + xlo + cx, lt = (y0 + cy - b0) * tiles_x + xlo + cx;
                mask[lt] = 0u;
                mask[plane + lt] = 0u;
                base[lt] = ranges[t].x + row[t];
            }
        }
        __syncthreads();
        for (int phase = 0; phase < 2; ++phase) {
            for (uint32_t u = tid; u < units; u += blockDim.x) {
                const uint32_t j = upper_slot(fc.roff, GVD_BIN_CHUNK, u);
                const uint4 bx = fc.box[j];
                const uint32_t ty = bx.z + (u - fc.roff[j]), x0 = bx.x, x1 = bx.x + bx.y;
                if (ty < b0 || ty >= b1) continue;
                const uint32_t gid = fc.id[j], bit = 1u << (j & 31), below = bit - 1u;
                uint32_t* mrow = mask + (ty - b0) * tiles_x;
                const uint32_t* brow = base + (ty - b0) * tiles_x;
                for (uint32_t tx = x0; tx < x1; ++tx) {
                    if (phase == 0) {
                        atomicOr(&mrow[(j >> 5) * plane + tx], bit);
                    } else {
                        const uint32_t lo = mrow[tx];
                        const uint32_t rank = j < 32 ? __popc(lo & below) : __popc(lo) + __popc(mrow[plane + tx] & below);
                        const uint32_t slot = brow[tx] + rank;
                        if (slot < capacity) point_list[slot] = gid;
                    }
                }
            }
            __syncthreads();
        }
    }
}

// Optional (export_keys): rebuild the reference's sorted 64-bit keys for parity checks.
__global__ void __launch_bounds__(256) export_keys_kernel(uint32_t capacity, int T, const uint2* __restrict__ ranges,
                                                          const uint32_t* __restrict__ point_list,
                                                          const uint32_t* __restrict__ depth_key,
                                                          uint64_t* __restrict__ keys) {
    pdl_wait();
    pdl_trigger();
    const int tile = blockIdx.x;
    const uint2 r = ranges[tile];
    for (uint32_t k = r.x + threadIdx.x; k < r.y && k < capacity; k += blockDim.x)
        keys[k] = ((uint64_t)tile << 32) | depth_key[point_list[k]];
}

// ------------------------------------------------------------------------------------------
template <int SPLIT>
__global__ void __launch_bounds__(256 / SPLIT, 4 * SPLIT) render_forward_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, const SplatRec* __restrict__ splat,
    int W, int H, uint32_t tiles_x, const float* __restrict__ bg_color, float* __restrict__ out_color,
    float* __restrict__ out_depth, float* __restrict__ out_alpha, uint32_t* __restrict__ n_contrib, uint32_t capacity) {
    pdl_wait();
    pdl_trigger();
    // SPLIT CTAs share one 16x16 tile (8 / SPLIT warps of 8x4 pixels each): shorter CTAs, finer early exit, smaller tail
    constexpr int BLOCK = 256 / SPLIT, BATCH = BLOCK;
    __shared__ __align__(128) float4 buf[2][BATCH * 3];
    __shared__ __align__(128) IdSlot ids[3];
    __shared__ __align__(8) uint64_t bar[3];

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t tile = blockIdx.x / SPLIT;
    const uint32_t gw = (blockIdx.x % SPLIT) * (8 / SPLIT) + warp;  // this warp's 8x4 block inside the tile
    const uint32_t tile_x = tile % tiles_x, tile_y = tile / tiles_x;
    const uint32_t sub_x = tile_x * GVD_TILE_X + (gw & 1) * 8, sub_y = tile_y * GVD_TILE_Y + (gw >> 1) * 4;
    const uint32_t px = sub_x + (lane & 7), py = sub_y + (lane >> 3);
    const bool inside = px < (uint32_t)W && py < (uint32_t)H;
    const uint32_t pix_id = W * py + px;
    const float2 pixf = {(float)px, (float)py};
    const float sxf = (float)sub_x, syf = (float)sub_y;

    uint2 range = ranges[tile];
    // a speculative instance buffer may be smaller than R: never read past it (the caller discards such a frame)
    range.x = min(range.x, capacity);
    range.y = min(range.y, capacity);
    const int n = (int)(range.y - range.x);
    const int rounds = (n + BATCH - 1) / BATCH;
    const uint32_t* list = point_list + range.x;

    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_init(&bar[2], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        if (rounds > 0) issue_id_copy(&ids[0], &bar[0], list, min(BATCH, n));
        if (rounds > 1) issue_id_copy(&ids[1], &bar[1], list + BATCH, min(BATCH, n - BATCH));
    }
    // batch 0 -> buf[0]
    if (rounds > 0) {
        mbar_wait(&bar[0], 0);
        if ((int)tid < min(BATCH, n)) {
            const float4* src = reinterpret_cast<const float4*>(splat + ids[0].v[id_lead(list) + tid]);
            buf[0][tid * 3 + 0] = __ldg(src);
            buf[0][tid * 3 + 1] = __ldg(src + 1);
            buf[0][tid * 3 + 2] = __ldg(src + 2);
        }
    }

    bool done = !inside;
    float T = 1.0f;
    uint32_t last_contributor = 0;
    float C0 = 0.f, C1 = 0.f, C2 = 0.f, weight = 0.f, Dsum = 0.f;

    for (int i = 0; i < rounds; ++i) {
        // whole tile saturated? (forward.cu:310-313). Also publishes buf[i&1] and retires buf[(i+1)&1].
        const int num_done = __syncthreads_count(done);
        if (num_done == BLOCK) {
            // a CTA must not exit under a pending bulk copy: batch i+1's ids may still be in flight
            if (i + 1 < rounds) mbar_wait(&bar[(i + 1) % 3], (uint32_t)(((i + 1) / 3) & 1));
            break;
        }
        const int cur = i & 1;
        const int cnt = min(BATCH, n - i * BATCH);
        // prefetch: records of batch i+1 into registers (ids landed a round ago), ids of batch i+2 by TMA
        float4 pa, pb, pc;
        const int ncnt = min(BATCH, n - (i + 1) * BATCH);
        const bool have_next = (i + 1 < rounds) && ((int)tid < ncnt);
        if (i + 1 < rounds) {
            mbar_wait(&bar[(i + 1) % 3], (uint32_t)(((i + 1) / 3) & 1));
            if (have_next) {
                const uint32_t* first = list + (size_t)(i + 1) * BATCH;
                const float4* src = reinterpret_cast<const float4*>(splat + ids[(i + 1) % 3].v[id_lead(first) + tid]);
                pa = __ldg(src);
                pb = __ldg(src + 1);
                pc = __ldg(src + 2);
            }
            if (tid == 0 && i + 2 < rounds)
                issue_id_copy(&ids[(i + 2) % 3], &bar[(i + 2) % 3], list + (size_t)(i + 2) * BATCH,
                              min(BATCH, n - (i + 2) * BATCH));
        }

        const float4* rec = buf[cur];
        const uint32_t base = (uint32_t)i * BATCH;
        for (int chunk = 0; chunk * 32 < cnt; ++chunk) {
            if (__all_sync(0xffffffffu, done)) break;
            const int e = chunk * 32 + (int)lane;
            bool hit = false;
            if (e < cnt) {
                const float4 ea = rec[e * 3], ec = rec[e * 3 + 2];
                hit = subtile_hit(ea.x, ea.y, ec.z, ec.w, sxf, syf);
            }
            unsigned m = __ballot_sync(0xffffffffu, hit);
            while (m) {
                const int j = chunk * 32 + __ffs(m) - 1;
                m &= m - 1;
                const float4 ra = rec[j * 3], rb = rec[j * 3 + 1];
                // forward.cu:335-366, same expression trees
                const float2 d = {ra.x - pixf.x, ra.y - pixf.y};
                const float power = -0.5f * (ra.z * d.x * d.x + rb.x * d.y * d.y) - ra.w * d.x * d.y;
                if (done || power > 0.0f) continue;
                const float alpha = fminf(0.99f, rb.y * expf(power));
                if (alpha < 1.0f / 255.0f) continue;
                const float test_T = T * (1 - alpha);
                if (test_T < 0.0001f) {
                    done = true;
                    continue;
                }
                const float4 rc = rec[j * 3 + 2];
                C0 += rb.z * alpha * T;
                C1 += rb.w * alpha * T;
                C2 += rc.x * alpha * T;
                weight += alpha * T;
                Dsum += rc.y * alpha * T;
                T = test_T;
                last_contributor = base + (uint32_t)j + 1u;
            }
        }
        if (have_next) {
            buf[cur ^ 1][tid * 3 + 0] = pa;
            buf[cur ^ 1][tid * 3 + 1] = pb;
            buf[cur ^ 1][tid * 3 + 2] = pc;
        }
    }

    if (inside) {
        const size_t HW = (size_t)H * W;
        n_contrib[pix_id] = last_contributor;
        out_color[0 * HW + pix_id] = C0 + T * bg_color[0];
        out_color[1 * HW + pix_id] = C1 + T * bg_color[1];
        out_color[2 * HW + pix_id] = C2 + T * bg_color[2];
        out_alpha[pix_id] = weight;
        out_depth[pix_id] = Dsum;
    }
}

__global__ void __launch_bounds__(256) mark_visible_kernel(int P, const float* __restrict__ orig_points,
                                                           const float* __restrict__ viewmatrix,
                                                           uint8_t* __restrict__ present) {
    pdl_wait();
    pdl_trigger();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const float3 p = {orig_points[3 * idx], orig_points[3 * idx + 1], orig_points[3 * idx + 2]};
    const float3 v = xform_point_4x3(p, viewmatrix);
    present[idx] = v.z > 0.2f ? 1 : 0;
}

}  // namespace

void gvd_launch_preprocess(const GvdRasterForwardArgs& a, const RasterGeomPtrs& g, const RasterSortPtrs& so, float focal_x,
                           float focal_y, dim3 grid, cudaStream_t s) {
    static int no_cull = -1;
    if (no_cull < 0) {
        const char* e = getenv("GVD_NO_SUBTILE_CULL");
        no_cull = (e && e[0] == '1') ? 1 : 0;
    }
    if (a.raw_params)
        gvd_launch(preprocess_kernel<true>, dim3((unsigned)so.nb), dim3(GVD_PRE_BLOCK), 0, s,
            a.P, a.D, a.M, a.means3D, (const float3*)a.scales, a.scale_modifier, (const float4*)a.rotations,
            a.opacities, a.shs, g.clamped, a.cov3D_precomp, a.colors_precomp, a.viewmatrix, a.projmatrix,
            (const float3*)a.campos, a.width, a.height, a.tan_fovx, a.tan_fovy, focal_x, focal_y, a.radii, g.splat, grid,
            g.tiles_touched, so.depth_key, so.blk_vis, so.blk_tiles, so.zeroed, so.zeroed_words, a.prefiltered, no_cull,
            a.shs_rest);
    else
        gvd_launch(preprocess_kernel<false>, dim3((unsigned)so.nb), dim3(GVD_PRE_BLOCK), 0, s,
            a.P, a.D, a.M, a.means3D, (const float3*)a.scales, a.scale_modifier, (const float4*)a.rotations,
            a.opacities, a.shs, g.clamped, a.cov3D_precomp, a.colors_precomp, a.viewmatrix, a.projmatrix,
            (const float3*)a.campos, a.width, a.height, a.tan_fovx, a.tan_fovy, focal_x, focal_y, a.radii, g.splat, grid,
            g.tiles_touched, so.depth_key, so.blk_vis, so.blk_tiles, so.zeroed, so.zeroed_words, a.prefiltered, no_cull,
            a.shs_rest);
}

void gvd_launch_compact(int P, const RasterGeomPtrs& g, const RasterSortPtrs& so, int* r_host, cudaStream_t s) {
    const unsigned blocks = (unsigned)((P + GVD_COMPACT_BLOCK - 1) / GVD_COMPACT_BLOCK);
    gvd_launch(compact_kernel, dim3(blocks), dim3(GVD_COMPACT_BLOCK), 0, s, P, (int)so.nb, g.tiles_touched, so.depth_key, so.blk_vis,
               so.blk_tiles, so.key[0], so.val[0], g.vis_id, g.counts, so.ghist, r_host);
}

void gvd_launch_depth_sort(int P, const RasterGeomPtrs& g, const RasterSortPtrs& so, cudaStream_t s) {
    // grid sized for V = P (V lives in device memory); CTAs whose tile starts at or past V return at once
    const unsigned blocks = (unsigned)so.nt;
    (void)P;
    for (int pass = 0; pass < 4; ++pass) {
        const int in = pass & 1, out = in ^ 1;
        gvd_launch(sort_pass_kernel, dim3(blocks), dim3(GVD_SORT_THREADS), 0, s, g.counts, so.key[in], so.val[in], so.key[out],
                   so.val[out], so.ghist + pass * 256 /* replica c at + c * 1024 */, so.agg + (size_t)pass * so.nt * 256, so.incl + (size_t)pass * so.ns * 256,
                   so.ticket + pass, g.counts + 7, 8 * pass);
    }
}

static cudaError_t ensure_smem(const void* fn, size_t bytes) {
    if (bytes <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

cudaError_t gvd_launch_bin_count(const RasterGeomPtrs& g, const RasterSortPtrs& so, const RasterHistPtrs& h, const RasterImgPtrs& im,
                                 dim3 grid, cudaStream_t s) {
    const int T = (int)(grid.x * grid.y);
    const size_t smem = (size_t)(grid.y + 1) * ((grid.x + 1) | 1) * sizeof(int);
    cudaError_t e = ensure_smem((const void*)bin_count_kernel, smem);
    if (e != cudaSuccess) return e;
    if (h.rows > 0)
        gvd_launch(bin_count_kernel, dim3((unsigned)h.rows), dim3(GVD_BIN_CHUNK), smem, s, g.counts, grid.x, grid.y, g.splat, so.val[0], h.hist);
    gvd_launch(bin_prefix_kernel, dim3((T + GVD_PFX_TILES - 1) / GVD_PFX_TILES), dim3(GVD_PFX_TILES * GVD_PFX_SEGS), 0, s, T, (int)h.rows,
               g.counts, h.hist, h.tile_total);
    gvd_launch(bin_ranges_kernel, dim3(1), dim3(1024), 0, s, T, h.tile_total, im.ranges);
    return cudaGetLastError();
}

cudaError_t gvd_launch_bin_fill(const RasterGeomPtrs& g, const RasterSortPtrs& so, const RasterHistPtrs& h, const RasterBinPtrs& b,
                                const RasterImgPtrs& im, dim3 grid, uint32_t capacity, cudaStream_t s) {
    const int T = (int)(grid.x * grid.y);
    // rows of tiles whose three planes fit the budget (all of them up to 16 384 tiles)
    const uint32_t band_rows = (uint32_t)std::max<size_t>(1, std::min<size_t>(grid.y, GVD_FILL_SMEM_BYTES / (3 * sizeof(uint32_t) * grid.x)));
    const size_t smem = (size_t)3 * band_rows * grid.x * sizeof(uint32_t);
    cudaError_t e = ensure_smem((const void*)bin_fill_kernel, smem);
    if (e != cudaSuccess) return e;
    if (h.rows > 0)
        gvd_launch(bin_fill_kernel, dim3((unsigned)h.rows), dim3(GVD_FILL_THREADS), smem, s, g.counts, T, grid.x, band_rows, g.splat,
                   so.val[0], h.hist, im.ranges, b.point_list, capacity);
    return cudaGetLastError();
}

void gvd_launch_export_keys(uint32_t capacity, const RasterSortPtrs& so, const RasterBinPtrs& b, const RasterImgPtrs& im,
                            dim3 grid, cudaStream_t s) {
    if (capacity == 0 || !b.keys) return;
    const int T = (int)(grid.x * grid.y);
    gvd_launch(export_keys_kernel, dim3(T), dim3(256), 0, s, capacity, T, im.ranges, b.point_list, so.depth_key, b.keys);
}

int gvd_render_split() {
    // CTAs per 16x16 tile in the render kernels (1, 2 or 4); GVD_RENDER_SPLIT overrides (A/B timing knob)
    static int split = 0;
    if (!split) {
        const char* e = getenv("GVD_RENDER_SPLIT");
        split = (e && (e[0] == '1' || e[0] == '2' || e[0] == '4')) ? e[0] - '0' : 2;
    }
    return split;
}

template <int SPLIT>
static void launch_render_forward(const GvdRasterForwardArgs& a, const RasterGeomPtrs& g, const RasterBinPtrs& b,
                                  const RasterImgPtrs& im, dim3 grid, uint32_t capacity, cudaStream_t s) {
    static bool carveout_set = false;
    if (!carveout_set) {  // many small CTAs per SM need the large shared-memory carveout
        cudaFuncSetAttribute((const void*)render_forward_kernel<SPLIT>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        carveout_set = true;
    }
    gvd_launch(render_forward_kernel<SPLIT>, dim3(grid.x * grid.y * SPLIT), dim3(256 / SPLIT), 0, s, im.ranges, b.point_list,
               g.splat, a.width, a.height, grid.x, a.background, a.out_color, a.out_depth, a.out_alpha, im.n_contrib, capacity);
}

void gvd_launch_render_forward(const GvdRasterForwardArgs& a, const RasterGeomPtrs& g, const RasterBinPtrs& b,
                               const RasterImgPtrs& im, dim3 grid, uint32_t capacity, cudaStream_t s) {
    switch (gvd_render_split()) {
        case 1: launch_render_forward<1>(a, g, b, im, grid, capacity, s); break;
        case 4: launch_render_forward<4>(a, g, b, im, grid, capacity, s); break;
        default: launch_render_forward<2>(a, g, b, im, grid, capacity, s); break;
    }
}

bool gvd_pdl_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("GVD_PDL");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on != 0;
}

void gvd_launch_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present,
                             cudaStream_t s) {
    gvd_launch(mark_visible_kernel, dim3((P + 255) / 256), dim3(256), 0, s, P, means3D, viewmatrix, present);
}
