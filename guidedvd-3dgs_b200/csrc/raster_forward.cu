// raster_forward.cu -- forward kernels of the B200-native Gaussian rasterizer (sm_100a).
//
//   preprocess_kernel     per Gaussian: cull, project, EWA cov2D, conic, radius, tile rect, SH->RGB
//                         (follows DGR/cuda_rasterizer/forward.cu:155-256 arithmetic exactly so that
//                         radii / tile rects / depth bits are bit-identical to the reference)
//   emit_keys_kernel      per Gaussian: (tile<<32 | depth bits, id) per covered tile
//                         (rasterizer_impl.cu:70-111; large rects are spread over the warp)
//   pack_kernel           per sorted instance: tile ranges (rasterizer_impl.cu:116-138) + gather the
//                         48-B splat record into sorted order + per-instance 8x4 sub-tile mask
//   render_forward_kernel per tile: TMA bulk (cp.async.bulk + mbarrier) double-buffered staging of the
//                         contiguous packed list; each warp owns an 8x4 pixel sub-tile and visits only
//                         the instances whose mask bit is set (forward.cu:261-381 semantics preserved:
//                         skipped instances are exactly those every pixel of the warp would `continue` on).
#include <cstdio>
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

    const float3* sh = reinterpret_cast<const float3*>(shs) + (size_t)idx * max_coeffs;
    float3 result = f3_scale(GVD_SH_C0, sh[0]);
    if (deg > 0) {
        float x = dir.x, y = dir.y, z = dir.z;
        result = f3_sub(f3_add(f3_sub(result, f3_scale(GVD_SH_C1 * y, sh[1])), f3_scale(GVD_SH_C1 * z, sh[2])),
                        f3_scale(GVD_SH_C1 * x, sh[3]));
        if (deg > 1) {
            float xx = x * x, yy = y * y, zz = z * z;
            float xy = x * y, yz = y * z, xz = x * z;
            result = f3_add(
                f3_add(f3_add(f3_add(f3_add(result, f3_scale(GVD_SH_C2_0 * xy, sh[4])), f3_scale(GVD_SH_C2_1 * yz, sh[5])),
                              f3_scale(GVD_SH_C2_2 * (2.0f * zz - xx - yy), sh[6])),
                       f3_scale(GVD_SH_C2_3 * xz, sh[7])),
                f3_scale(GVD_SH_C2_4 * (xx - yy), sh[8]));
            if (deg > 2) {
                result = f3_add(
                    f3_add(f3_add(f3_add(f3_add(f3_add(f3_add(result, f3_scale(GVD_SH_C3_0 * y * (3.0f * xx - yy), sh[9])),
                                                       f3_scale(GVD_SH_C3_1 * xy * z, sh[10])),
                                                f3_scale(GVD_SH_C3_2 * y * (4.0f * zz - xx - yy), sh[11])),
                                         f3_scale(GVD_SH_C3_3 * z * (2.0f * zz - 3.0f * xx - 3.0f * yy), sh[12])),
                                  f3_scale(GVD_SH_C3_4 * x * (4.0f * zz - xx - yy), sh[13])),
                           f3_scale(GVD_SH_C3_5 * z * (xx - yy), sh[14])),
                    f3_scale(GVD_SH_C3_6 * x * (xx - 3.0f * yy), sh[15]));
            }
        }
    }
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

__global__ void __launch_bounds__(256) preprocess_kernel(
    int P, int D, int M, const float* __restrict__ orig_points, const float3* __restrict__ scales,
    const float scale_modifier, const float4* __restrict__ rotations, const float* __restrict__ opacities,
    const float* __restrict__ shs, uint8_t* __restrict__ clamped, const float* __restrict__ cov3D_precomp,
    const float* __restrict__ colors_precomp, const float* __restrict__ viewmatrix,
    const float* __restrict__ projmatrix, const float3* __restrict__ cam_pos, const int W, int H,
    const float tan_fovx, float tan_fovy, const float focal_x, float focal_y, int* __restrict__ radii,
    SplatRec* __restrict__ splat, const dim3 grid, uint32_t* __restrict__ tiles_touched, int prefiltered) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;

    radii[idx] = 0;
    tiles_touched[idx] = 0;

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

    SplatRec rec;
    rec.a = make_float4(point_image.x, point_image.y, conic.x, conic.y);
    rec.b = make_float4(conic.z, opacities[idx], rgb.x, rgb.y);
    rec.c = make_float4(rgb.z, p_view.z, __uint_as_float(rect_min.x | (rect_min.y << 16)),
                        __uint_as_float(rect_max.x | (rect_max.y << 16)));
    splat[idx] = rec;
}

// ------------------------------------------------------------------------------------------
// Key emission. One thread per Gaussian for small rects; rects with more than EMIT_SERIAL_MAX
// tiles are handed to the whole warp (the reference's serial per-thread loop leaves 31 lanes
// idle behind one screen-filling Gaussian).
#define EMIT_SERIAL_MAX 16
__global__ void __launch_bounds__(256) emit_keys_kernel(int P, const SplatRec* __restrict__ splat,
                                                        const uint32_t* __restrict__ tiles_touched,
                                                        const uint32_t* __restrict__ offsets,
                                                        uint64_t* __restrict__ keys_unsorted,
                                                        uint32_t* __restrict__ values_unsorted, dim3 grid) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    uint32_t n = 0, off = 0, w0 = 0, w1 = 0, depth_bits = 0;
    if (idx < P) {
        n = tiles_touched[idx];
        if (n > 0) {
            off = (idx == 0) ? 0 : offsets[idx - 1];
            const float4 c = splat[idx].c;
            depth_bits = __float_as_uint(c.y);
            w0 = __float_as_uint(c.z);
            w1 = __float_as_uint(c.w);
        }
    }
    const uint32_t x0 = w0 & 0xffff, y0 = w0 >> 16, x1 = w1 & 0xffff, y1 = w1 >> 16;
    if (n > 0 && n <= EMIT_SERIAL_MAX) {
        for (uint32_t y = y0; y < y1; y++)
            for (uint32_t x = x0; x < x1; x++) {
                uint64_t key = y * grid.x + x;
                key <<= 32;
                key |= depth_bits;
                keys_unsorted[off] = key;
                values_unsorted[off] = idx;
                off++;
            }
    }
    unsigned big = __ballot_sync(0xffffffffu, n > EMIT_SERIAL_MAX);
    while (big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        const uint32_t bn = __shfl_sync(0xffffffffu, n, src);
        const uint32_t boff = __shfl_sync(0xffffffffu, off, src);
        const uint32_t bx0 = __shfl_sync(0xffffffffu, x0, src);
        const uint32_t by0 = __shfl_sync(0xffffffffu, y0, src);
        const uint32_t bx1 = __shfl_sync(0xffffffffu, x1, src);
        const uint32_t bdepth = __shfl_sync(0xffffffffu, depth_bits, src);
        const uint32_t bidx = __shfl_sync(0xffffffffu, (uint32_t)idx, src);
        const uint32_t bw = bx1 - bx0;
        for (uint32_t k = lane; k < bn; k += 32) {
            const uint32_t y = by0 + k / bw, x = bx0 + k % bw;  // y-major then x, as the reference
            uint64_t key = y * grid.x + x;
            key <<= 32;
            key |= bdepth;
            keys_unsorted[boff + k] = key;
            values_unsorted[boff + k] = bidx;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Sub-tile mask: bit w set iff the instance can reach alpha >= 1/255 on some pixel of warp w's
// 8x4 block (sx = w&1, sy = w>>1). Conservative (axis-aligned bound of the level-set ellipse
// plus slack), so skipping a cleared bit never changes a pixel: on those pixels the reference
// takes `continue` at forward.cu:348.
__device__ __forceinline__ uint32_t subtile_mask(const float4 a, const float4 b, uint32_t tile_x, uint32_t tile_y) {
    const float opac = b.y;
    const float A = a.z, B = a.w, C = b.x;
    // alpha = min(.99, opac*exp(power)) with power <= 0, so opac < 1/255 can never pass.
    if (opac * 255.0f < 0.999f) return 0u;
    const float t = 2.0f * logf(opac * 255.0f) * 1.0005f + 1e-3f;  // q(d) <= t  <=>  power >= -t/2
    const float det = A * C - B * B;
    if (!(det > 0.0f) || !(A > 0.0f) || !(C > 0.0f)) return 0xffu;
    const float hx = sqrtf(t * C / det) * 1.0005f + 0.02f;
    const float hy = sqrtf(t * A / det) * 1.0005f + 0.02f;
    if (!(hx < 1e9f) || !(hy < 1e9f)) return 0xffu;
    const float ox = (float)(tile_x * GVD_TILE_X), oy = (float)(tile_y * GVD_TILE_Y);
    const float xmin = a.x - hx - ox, xmax = a.x + hx - ox;  // tile-local
    const float ymin = a.y - hy - oy, ymax = a.y + hy - oy;
    uint32_t colmask = 0, m = 0;
    if (xmin <= 7.0f && xmax >= 0.0f) colmask |= 1u;
    if (xmin <= 15.0f && xmax >= 8.0f) colmask |= 2u;
#pragma unroll
    for (int sy = 0; sy < 4; ++sy)
        if (ymin <= (float)(sy * 4 + 3) && ymax >= (float)(sy * 4)) m |= colmask << (2 * sy);
    return m;
}

__global__ void __launch_bounds__(256) pack_kernel(int R, const uint64_t* __restrict__ keys,
                                                   const uint32_t* __restrict__ point_list,
                                                   const SplatRec* __restrict__ splat, SplatRec* __restrict__ packed,
                                                   uint2* __restrict__ ranges, uint32_t tiles_x) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= R) return;
    const uint64_t key = keys[idx];
    const uint32_t currtile = key >> 32;
    if (idx == 0)
        ranges[currtile].x = 0;
    else {
        const uint32_t prevtile = keys[idx - 1] >> 32;
        if (currtile != prevtile) {
            ranges[prevtile].y = idx;
            ranges[currtile].x = idx;
        }
    }
    if (idx == R - 1) ranges[currtile].y = R;

    const uint32_t id = point_list[idx];
    const float4* src = reinterpret_cast<const float4*>(splat + id);
    float4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2);
    c.z = __uint_as_float(id);
    c.w = __uint_as_float(subtile_mask(a, b, currtile % tiles_x, currtile / tiles_x));
    float4* dst = reinterpret_cast<float4*>(packed + idx);
    dst[0] = a;
    dst[1] = b;
    dst[2] = c;
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GVD_BLOCK) render_forward_kernel(
    const uint2* __restrict__ ranges, const SplatRec* __restrict__ packed, int W, int H, uint32_t tiles_x,
    const float* __restrict__ bg_color, float* __restrict__ out_color, float* __restrict__ out_depth,
    float* __restrict__ out_alpha, uint32_t* __restrict__ n_contrib) {
    __shared__ __align__(128) float4 buf[2][GVD_BATCH * 3];
    __shared__ __align__(8) uint64_t bar[2];

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t tile = blockIdx.x;
    const uint32_t tile_x = tile % tiles_x, tile_y = tile / tiles_x;
    const uint32_t px = tile_x * GVD_TILE_X + (warp & 1) * 8 + (lane & 7);
    const uint32_t py = tile_y * GVD_TILE_Y + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < (uint32_t)W && py < (uint32_t)H;
    const uint32_t pix_id = W * py + px;
    const float2 pixf = {(float)px, (float)py};

    const uint2 range = ranges[tile];
    const int n = (int)(range.y - range.x);
    const int rounds = (n + GVD_BATCH - 1) / GVD_BATCH;
    const SplatRec* list = packed + range.x;

    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0 && rounds > 0) {
        const uint32_t bytes = (uint32_t)min(GVD_BATCH, n) * (uint32_t)sizeof(SplatRec);
        mbar_arrive_expect_tx(&bar[0], bytes);
        tma_bulk_g2s(&buf[0][0], list, bytes, &bar[0]);
    }

    bool done = !inside;
    float T = 1.0f;
    uint32_t last_contributor = 0;
    float C0 = 0.f, C1 = 0.f, C2 = 0.f, weight = 0.f, Dsum = 0.f;

    for (int i = 0; i < rounds; ++i) {
        // whole tile saturated? (forward.cu:310-313). Also frees buf[(i+1)&1] for the prefetch.
        const int num_done = __syncthreads_count(done);
        if (num_done == GVD_BLOCK) {
            // batch i is already in flight: a CTA must not exit under a pending bulk copy.
            mbar_wait(&bar[i & 1], (uint32_t)((i >> 1) & 1));
            break;
        }
        const int cur = i & 1;
        const int cnt = min(GVD_BATCH, n - i * GVD_BATCH);
        if (tid == 0 && i + 1 < rounds) {
            const uint32_t bytes = (uint32_t)min(GVD_BATCH, n - (i + 1) * GVD_BATCH) * (uint32_t)sizeof(SplatRec);
            mbar_arrive_expect_tx(&bar[cur ^ 1], bytes);
            tma_bulk_g2s(&buf[cur ^ 1][0], list + (size_t)(i + 1) * GVD_BATCH, bytes, &bar[cur ^ 1]);
        }
        mbar_wait(&bar[cur], (uint32_t)((i >> 1) & 1));

        const float4* rec = buf[cur];
        const uint32_t base = (uint32_t)i * GVD_BATCH;
        for (int chunk = 0; chunk * 32 < cnt; ++chunk) {
            if (__all_sync(0xffffffffu, done)) break;
            const int e = chunk * 32 + (int)lane;
            const uint32_t mk = (e < cnt) ? __float_as_uint(rec[e * 3 + 2].w) : 0u;
            unsigned m = __ballot_sync(0xffffffffu, (mk >> warp) & 1u);
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
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const float3 p = {orig_points[3 * idx], orig_points[3 * idx + 1], orig_points[3 * idx + 2]};
    const float3 v = xform_point_4x3(p, viewmatrix);
    present[idx] = v.z > 0.2f ? 1 : 0;
}

}  // namespace

void gvd_launch_preprocess(const GvdRasterForwardArgs& a, const RasterGeomPtrs& g, float focal_x, float focal_y,
                           dim3 grid, cudaStream_t s) {
    preprocess_kernel<<<(a.P + 255) / 256, 256, 0, s>>>(
        a.P, a.D, a.M, a.means3D, (const float3*)a.scales, a.scale_modifier, (const float4*)a.rotations,
        a.opacities, a.shs, g.clamped, a.cov3D_precomp, a.colors_precomp, a.viewmatrix, a.projmatrix,
        (const float3*)a.campos, a.width, a.height, a.tan_fovx, a.tan_fovy, focal_x, focal_y, a.radii, g.splat, grid,
        g.tiles_touched, a.prefiltered);
}

void gvd_launch_emit_keys(int P, const RasterGeomPtrs& g, const RasterBinPtrs& b, dim3 grid, cudaStream_t s) {
    emit_keys_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, g.splat, g.tiles_touched, g.point_offsets, b.keys_unsorted,
                                                     b.point_list_unsorted, grid);
}

void gvd_launch_pack(int R, const RasterGeomPtrs& g, const RasterBinPtrs& b, const RasterImgPtrs& im, dim3 grid,
                     cudaStream_t s) {
    if (R > 0)
        pack_kernel<<<(R + 255) / 256, 256, 0, s>>>(R, b.keys, b.point_list, g.splat, b.packed, im.ranges, grid.x);
}

void gvd_launch_render_forward(const GvdRasterForwardArgs& a, const RasterBinPtrs& b, const RasterImgPtrs& im,
                               dim3 grid, cudaStream_t s) {
    render_forward_kernel<<<grid.x * grid.y, GVD_BLOCK, 0, s>>>(im.ranges, b.packed, a.width, a.height, grid.x,
                                                                 a.background, a.out_color, a.out_depth, a.out_alpha,
                                                                 im.n_contrib);
}

void gvd_launch_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present,
                             cudaStream_t s) {
    mark_visible_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, means3D, viewmatrix, present);
}
