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
#include <cstdio>
#include <cstdlib>
#include "raster_common.cuh"
#include "../../include/gvd_raster.h"

namespace {

// ------------------------------------------------------------------------------------------
// forward.cu:20-71
__device__ __forceinline__ float3 sh_to_rgb(int idx, int deg, int max_coeffs, const float3 pos, const float3 campos,
                                            const float* __restrict__ shs, uint8_t* clamped_bits) {
    float3 dir = f3_sub(pos, campos);
    float len = sqrtf(f3_dot(dir, dir));
    dir = {dir.x / len, dir.y / len, dir.z / len};

    float v[48];
    load_sh(shs, idx, deg, max_coeffs, v);
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
__device__ __forceinline__ void alpha_extent(const float3 conic, float opac, float& hx, float& hy) {
    hx = hy = 1e30f;
    if (opac * 255.0f < 0.999f) {
        hx = hy = -1e30f;
        return;
    }
    const float A = conic.x, B = conic.y, C = conic.z;
    const float t = 2.0f * logf(opac * 255.0f) * 1.0005f + 1e-3f;  // q(d) <= t  <=>  power >= -t/2
    const float det = A * C - B * B;
    if (!(det > 0.0f) || !(A > 0.0f) || !(C > 0.0f) || !(t >= 0.0f)) return;
    const float ex = sqrtf(t * C / det) * 1.0005f + 0.02f;
    const float ey = sqrtf(t * A / det) * 1.0005f + 0.02f;
    if (!(ex < 1e9f) || !(ey < 1e9f)) return;
    hx = ex;
    hy = ey;
}

__global__ void __launch_bounds__(256) preprocess_kernel(
    int P, int D, int M, const float* __restrict__ orig_points, const float3* __restrict__ scales,
    const float scale_modifier, const float4* __restrict__ rotations, const float* __restrict__ opacities,
    const float* __restrict__ shs, uint8_t* __restrict__ clamped, const float* __restrict__ cov3D_precomp,
    const float* __restrict__ colors_precomp, const float* __restrict__ viewmatrix,
    const float* __restrict__ projmatrix, const float3* __restrict__ cam_pos, const int W, int H,
    const float tan_fovx, float tan_fovy, const float focal_x, float focal_y, int* __restrict__ radii,
    SplatRec* __restrict__ splat, const dim3 grid, uint32_t* __restrict__ tiles_touched,
    uint32_t* __restrict__ depth_key, uint32_t* __restrict__ gidx, int prefiltered) {
    pdl_wait();
    pdl_trigger();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;

    radii[idx] = 0;
    tiles_touched[idx] = 0;
    depth_key[idx] = 0xFFFFFFFFu;  // culled Gaussians sort behind every visible one
    gidx[idx] = idx;

    // near cull (auxiliary.h:139-164): only p_view.z <= 0.2 rejects.
    const float3 p_orig = {orig_points[3 * idx], orig_points[3 * idx + 1], orig_points[3 * idx + 2]};
    const float3 p_view = xform_point_4x3(p_orig, viewmatrix);
    if (p_view.z <= 0.2f) {
        if (prefiltered) {
            printf("Point is filtered although prefiltered is set. This shouldn't happen!");
            __trap();
        }
        return;
    }

    const float4 p_hom = xform_point_4x4(p_orig, projmatrix);
    const float p_w = 1.0f / (p_hom.w + 0.0000001f);
    const float3 p_proj = {p_hom.x * p_w, p_hom.y * p_w, p_hom.z * p_w};

    float cov3D_local[6];
    const float* cov3D;
    if (cov3D_precomp != nullptr) {
        cov3D = cov3D_precomp + (size_t)idx * 6;
    } else {
        cov3d_from_scale_rot(scales[idx], scale_modifier, rotations[idx], cov3D_local);
        cov3D = cov3D_local;
    }

    const float3 cov = cov2d_from_cov3d(p_orig, focal_x, focal_y, tan_fovx, tan_fovy, cov3D, viewmatrix);

    const float det = (cov.x * cov.z - cov.y * cov.y);
    if (det == 0.0f) return;
    const float det_inv = 1.f / det;
    const float3 conic = {cov.z * det_inv, -cov.y * det_inv, cov.x * det_inv};

    const float mid = 0.5f * (cov.x + cov.z);
    const float lambda1 = mid + sqrtf(fmaxf(0.1f, mid * mid - det));
    const float lambda2 = mid - sqrtf(fmaxf(0.1f, mid * mid - det));
    const float my_radius = ceilf(3.f * sqrtf(fmaxf(lambda1, lambda2)));
    const float2 point_image = {ndc_to_pix(p_proj.x, W), ndc_to_pix(p_proj.y, H)};
    uint2 rect_min, rect_max;
    tile_rect(point_image, (int)my_radius, rect_min, rect_max, grid);
    if ((rect_max.x - rect_min.x) * (rect_max.y - rect_min.y) == 0) return;

    float3 rgb;
    if (colors_precomp == nullptr) {
        uint8_t cl;
        rgb = sh_to_rgb(idx, D, M, p_orig, *cam_pos, shs, &cl);
        clamped[idx] = cl;
    } else {
        rgb = {colors_precomp[3 * idx], colors_precomp[3 * idx + 1], colors_precomp[3 * idx + 2]};
    }

    radii[idx] = (int)my_radius;
    tiles_touched[idx] = (rect_max.y - rect_min.y) * (rect_max.x - rect_min.x);
    depth_key[idx] = __float_as_uint(p_view.z);

    SplatRec rec;
    rec.a = make_float4(point_image.x, point_image.y, conic.x, conic.y);
    rec.b = make_float4(conic.z, opacities[idx], rgb.x, rgb.y);
    float hx, hy;
    alpha_extent(conic, opacities[idx], hx, hy);
    rec.c = make_float4(rgb.z, p_view.z, hx, hy);
    rec.d = make_float4(__uint_as_float(rect_min.x | (rect_min.y << 16)), __uint_as_float(rect_max.x | (rect_max.y << 16)),
                        0.f, 0.f);
    splat[idx] = rec;
}

// ------------------------------------------------------------------------------------------
// ---- binning: rect-aware stable counting sort on the tile id -------------------------------------
// Instance order required (and produced by the reference's stable sort on tile<<32|depth): tile-major;
// inside a tile by depth, ties by Gaussian id. `order` already lists the Gaussians by (depth, id).

__device__ __forceinline__ void unpack_rect(const float4 d, uint32_t& x0, uint32_t& y0, uint32_t& x1, uint32_t& y1) {
    const uint32_t w0 = __float_as_uint(d.x), w1 = __float_as_uint(d.y);
    x0 = w0 & 0xffff; y0 = w0 >> 16; x1 = w1 & 0xffff; y1 = w1 >> 16;
}

// Pass 1: chunk c = GVD_BIN_CHUNK depth-consecutive Gaussians order[c*CHUNK ..]. hist[c][t] = how many of them cover tile t.
// A rect adds +1/-1 at its four corners of a (gy+1) x (gx+1) difference grid; a 2-D prefix sum then yields
// the per-tile counts. Cost per chunk is O(CHUNK + T) whatever the rect sizes (no per-instance atomics).
__global__ void __launch_bounds__(GVD_BIN_CHUNK) bin_count_kernel(int P, uint32_t gx, uint32_t gy,
                                                                 const SplatRec* __restrict__ splat,
                                                                 const uint32_t* __restrict__ order,
                                                                 const uint32_t* __restrict__ tiles_touched,
                                                                 uint32_t* __restrict__ chunk_flags,
                                                                 uint32_t* __restrict__ hist) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ int diff[];  // (gy+1) rows of stride ld
    const int ld = (int)(gx + 1) | 1;  // odd stride: column walks are bank-conflict free
    const int rows = (int)gy + 1, cols = (int)gx + 1;
    const int i = blockIdx.x * GVD_BIN_CHUNK + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t n = 0, x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    if (i < P) {
        const uint32_t id = order[i];
        n = tiles_touched[id];
        if (n > 0) unpack_rect(splat[id].d, x0, y0, x1, y1);
    }
    const int any = __syncthreads_or(n > 0);
    if (threadIdx.x == 0) chunk_flags[blockIdx.x] = any ? 1u : 0u;
    if (!any) return;  // culled Gaussians sort last: this and all later chunks are empty
    for (int t = threadIdx.x; t < rows * ld; t += GVD_BIN_CHUNK) diff[t] = 0;
    __syncthreads();
    if (n > 0) {
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

// Pass 2: per tile, exclusive prefix over the (valid) chunks, in place, and the tile's total.
// CTA = 32 tiles x 32 chunk segments: every thread sums its segment, the segment sums are scanned through
// shared memory, then every thread rewrites its segment as running prefixes.
__global__ void __launch_bounds__(1024) bin_prefix_kernel(int T, int chunks, const uint32_t* __restrict__ chunk_flags,
                                                          uint32_t* hist, uint32_t* __restrict__ tile_total) {
    pdl_wait();
    pdl_trigger();
    __shared__ uint32_t seg_sum[32][33];
    __shared__ int s_valid;
    const int tx = threadIdx.x & 31, seg = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_valid = 0;
    __syncthreads();
    {   // valid chunks form a prefix of the chunk array: count them
        int c = 0;
        for (int k = threadIdx.x; k < chunks; k += 1024) c += chunk_flags[k] ? 1 : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (tx == 0 && c) atomicAdd(&s_valid, c);
    }
    __syncthreads();
    const int nv = s_valid;
    const int per = (nv + 31) / 32;
    const int c0 = seg * per, c1 = min(nv, c0 + per);
    const int t = blockIdx.x * 32 + tx;
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
        if (seg == 31) {
            uint32_t tot = 0;
#pragma unroll 8
            for (int k = 0; k < 32; ++k) tot += seg_sum[k][tx];
            tile_total[t] = tot;
        }
    }
}

// Pass 3: exclusive scan over the tiles -> ranges (rasterizer_impl.cu:116-138 semantics: empty tiles (0,0))
// and R. One CTA; T <= GVD_MAX_TILES.
__global__ void __launch_bounds__(1024) bin_ranges_kernel(int T, const uint32_t* __restrict__ tile_total,
                                                          uint2* __restrict__ ranges, uint32_t* __restrict__ num_rendered,
                                                          int* r_host) {
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
    if (tid == 0) {
        *num_rendered = carry_s;
        // speculative path: R goes straight into the caller's pinned (device-mapped) host word. A cudaMemcpyAsync in
        // the compute stream queues behind whatever the copy engines are busy with -- measured +46 us per step while
        // a 6 MB host-to-device upload of the next step's inputs was in flight.
        if (r_host != nullptr) {
            *reinterpret_cast<volatile int*>(r_host) = (int)carry_s;
            __threadfence_system();
        }
    }
}

// Pass 4: chunk c writes its Gaussians' ids into the tile lists in depth order: the slot of Gaussian g in tile t is
// ranges[t].x + (chunk's starting rank in t, from pass 2) + (number of earlier Gaussians of the chunk covering t);
// cnt[t] holds the next free slot of tile t. The only order constraint is per tile.
//   chunks up to GVD_FILL_HEAVY instances: ONE warp walks the chunk's flattened (Gaussian, tile) sequence 32 instances
//     at a time -- lane = instance, located by a load-balanced search over the chunk's instance offsets (REDUX.OR of
//     the "a Gaussian starts here" bits + popc). Instances of a batch that fall into the same tile are ranked with
//     MATCH.ANY (lane order = depth order); the first of each group advances cnt[t].
//   heavier chunks (the Gaussians nearest to the camera, rects up to the whole screen; C2: median chunk 800 instances,
//     largest 40 000): the tile ROWS are dealt out to the CTA's 8 warps (warp w owns rows ty = w mod 8); every warp
//     walks the chunk on its own, lanes across the columns of the rect, no block barrier in the walk.
// Measured: 85 us at C2, 265 us at C4 (1600x1066; a two-warp walk with a barrier per Gaussian took 445 us there).
// Four other organisations (two-warp serial walk, rows over 4/8 warps for every chunk, shared-memory bit masks,
// heavy chunks cut into slices over up to 32 CTAs) all land at 83-92 us at C2: the pass is bound by its 3.7 M
// scattered 4-byte stores -- every 32-byte sector of a tile list is written by ~8 different chunks -- not by the walk.
struct FillChunk {  // the chunk's non-empty Gaussians, compacted in order
    uint32_t id[GVD_BIN_CHUNK], r0[GVD_BIN_CHUNK], r1[GVD_BIN_CHUNK], off[GVD_BIN_CHUNK + 1], magic[GVD_BIN_CHUNK];
    uint32_t wn[2], wc[2];
};

#define GVD_FILL_THREADS 256
#define GVD_FILL_HEAVY 6144
__global__ void __launch_bounds__(GVD_FILL_THREADS) bin_fill_kernel(int P, int T, uint32_t tiles_x,
                                                                   const SplatRec* __restrict__ splat,
                                                                   const uint32_t* __restrict__ order,
                                                                   const uint32_t* __restrict__ tiles_touched,
                                                                   const uint32_t* __restrict__ chunk_flags,
                                                                   const uint32_t* __restrict__ hist,
                                                                   const uint2* __restrict__ ranges,
                                                                   uint32_t* __restrict__ point_list, uint32_t capacity) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ uint32_t cnt[];
    __shared__ FillChunk fc;
    static_assert(GVD_BIN_CHUNK == 64, "two warps load the chunk");
    if (!chunk_flags[blockIdx.x]) return;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t* row = hist + (size_t)blockIdx.x * T;
    // threads 0..63 load the chunk and compact its non-empty Gaussians (order kept)
    uint32_t id = 0, n = 0, r0 = 0, r1 = 0, incl = 0, ballot = 0;
    if (tid < GVD_BIN_CHUNK) {
        const int i = blockIdx.x * GVD_BIN_CHUNK + (int)tid;
        if (i < P) {
            id = order[i];
            n = tiles_touched[id];
            if (n > 0) {
                const float4 d = splat[id].d;
                r0 = __float_as_uint(d.x);
                r1 = __float_as_uint(d.y);
            }
        }
        ballot = __ballot_sync(0xffffffffu, n > 0);
        incl = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (uint32_t)o) incl += u;
        }
        if (lane == 31) {
            fc.wn[warp] = incl;
            fc.wc[warp] = __popc(ballot);
        }
    }
    for (uint32_t t = tid; t < (uint32_t)T; t += blockDim.x) cnt[t] = ranges[t].x + row[t];
    __syncthreads();
    const uint32_t m = fc.wc[0] + fc.wc[1], total = fc.wn[0] + fc.wn[1];
    if (tid < GVD_BIN_CHUNK) {
        const uint32_t base_n = warp ? fc.wn[0] : 0u, base_c = warp ? fc.wc[0] : 0u;
        if (n > 0) {
            const uint32_t j = base_c + __popc(ballot & lt_mask);
            const uint32_t bw = (r1 & 0xffff) - (r0 & 0xffff);
            fc.id[j] = id;
            fc.r0[j] = r0;
            fc.r1[j] = r1;
            fc.off[j] = base_n + incl - n;
            fc.magic[j] = 0xffffffffu / bw + 1u;  // floor(k / bw) == umulhi(k, magic) for k * bw < 2^32 (bw > 1)
        }
        if (tid == 0) fc.off[m] = total;
    }
    __syncthreads();

    if (total <= GVD_FILL_HEAVY) {
        if (warp != 0) return;
        const uint32_t le_mask = lt_mask | (1u << lane);
        const uint32_t o_lo = lane < m ? fc.off[lane] : 0xffffffffu;
        const uint32_t o_hi = lane + 32 < m ? fc.off[lane + 32] : 0xffffffffu;
        uint32_t before = 0;  // compacted Gaussians that start before this batch
        for (uint32_t b0 = 0; b0 < total; b0 += 32) {
            uint32_t bits = 0;
            if (o_lo - b0 < 32u) bits |= 1u << (o_lo - b0);
            if (o_hi - b0 < 32u) bits |= 1u << (o_hi - b0);
            const uint32_t starts = __reduce_or_sync(0xffffffffu, bits);
            const uint32_t q = b0 + lane;
            const bool valid = q < total;
            uint32_t t = 0xffffff00u + lane, gid = 0;  // invalid lanes: distinct keys, they match nobody
            if (valid) {
                const uint32_t j = before + __popc(starts & le_mask) - 1u;
                const uint32_t g0 = fc.r0[j], g1 = fc.r1[j];
                const uint32_t x0 = g0 & 0xffff, bw = (g1 & 0xffff) - x0;
                const uint32_t k = q - fc.off[j];
                const uint32_t ry = bw > 1 ? __umulhi(k, fc.magic[j]) : k;
                t = ((g0 >> 16) + ry) * tiles_x + x0 + (k - ry * bw);
                gid = fc.id[j];
            }
            before += __popc(starts);
            const uint32_t peers = __match_any_sync(0xffffffffu, t);
            const uint32_t rank = __popc(peers & lt_mask);
            const uint32_t base = valid ? cnt[t] : 0u;
            __syncwarp();
            if (valid) {
                if (rank == 0) cnt[t] = base + __popc(peers);
                const uint32_t slot = base + rank;
                if (slot < capacity) point_list[slot] = gid;  // capacity: speculative buffers may be too small
            }
            __syncwarp();
        }
        return;
    }

    for (uint32_t g = 0; g < m; ++g) {
        const uint32_t g0 = fc.r0[g], g1 = fc.r1[g];
        const uint32_t y0 = g0 >> 16, y1 = g1 >> 16, x0 = g0 & 0xffff, x1 = g1 & 0xffff;
        const uint32_t gid = fc.id[g];
        for (uint32_t ty = y0 + (warp + nw - y0 % nw) % nw; ty < y1; ty += nw)
            for (uint32_t tx = x0 + lane; tx < x1; tx += 32) {
                const uint32_t t = ty * tiles_x + tx;
                const uint32_t slot = cnt[t];
                cnt[t] = slot + 1;
                if (slot < capacity) point_list[slot] = gid;
            }
        __syncwarp();  // lanes change tiles from one Gaussian to the next: keep the warp's walk in step
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

void gvd_launch_preprocess(const GvdRasterForwardArgs& a, const RasterGeomPtrs& g, float focal_x, float focal_y,
                           dim3 grid, cudaStream_t s) {
    gvd_launch(preprocess_kernel, dim3((a.P + 255) / 256), dim3(256), 0, s, 
        a.P, a.D, a.M, a.means3D, (const float3*)a.scales, a.scale_modifier, (const float4*)a.rotations,
        a.opacities, a.shs, g.clamped, a.cov3D_precomp, a.colors_precomp, a.viewmatrix, a.projmatrix,
        (const float3*)a.campos, a.width, a.height, a.tan_fovx, a.tan_fovy, focal_x, focal_y, a.radii, g.splat, grid,
        g.tiles_touched, g.depth_key, g.gidx, a.prefiltered);
}

static cudaError_t ensure_smem(const void* fn, size_t bytes) {
    if (bytes <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

cudaError_t gvd_launch_bin_count(int P, const RasterGeomPtrs& g, const RasterImgPtrs& im, dim3 grid, int* r_host,
                                 cudaStream_t s) {
    const int T = (int)(grid.x * grid.y);
    const size_t smem = (size_t)(grid.y + 1) * ((grid.x + 1) | 1) * sizeof(int);
    cudaError_t e = ensure_smem((const void*)bin_count_kernel, smem);
    if (e != cudaSuccess) return e;
    gvd_launch(bin_count_kernel, dim3((unsigned)g.chunks), dim3(GVD_BIN_CHUNK), smem, s, P, grid.x, grid.y, g.splat, g.order,
                                                                    g.tiles_touched, g.chunk_flags, g.hist);
    gvd_launch(bin_prefix_kernel, dim3((T + 31) / 32), dim3(1024), 0, s, T, (int)g.chunks, g.chunk_flags, g.hist, g.tile_total);
    gvd_launch(bin_ranges_kernel, dim3(1), dim3(1024), 0, s, T, g.tile_total, im.ranges, g.num_rendered, r_host);
    return cudaGetLastError();
}

cudaError_t gvd_launch_bin_fill(int P, const RasterGeomPtrs& g, const RasterBinPtrs& b, const RasterImgPtrs& im,
                                dim3 grid, uint32_t capacity, cudaStream_t s) {
    const int T = (int)(grid.x * grid.y);
    const size_t smem = (size_t)T * sizeof(uint32_t);
    cudaError_t e = ensure_smem((const void*)bin_fill_kernel, smem);
    if (e != cudaSuccess) return e;
    gvd_launch(bin_fill_kernel, dim3((unsigned)g.chunks), dim3(GVD_FILL_THREADS), smem, s, P, T, grid.x, g.splat, g.order,
               g.tiles_touched, g.chunk_flags, g.hist, im.ranges, b.point_list, capacity);
    return cudaGetLastError();
}

void gvd_launch_export_keys(uint32_t capacity, const RasterGeomPtrs& g, const RasterBinPtrs& b, const RasterImgPtrs& im,
                            dim3 grid, cudaStream_t s) {
    if (capacity == 0 || !b.keys) return;
    const int T = (int)(grid.x * grid.y);
    gvd_launch(export_keys_kernel, dim3(T), dim3(256), 0, s, capacity, T, im.ranges, b.point_list, g.depth_key, b.keys);
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
