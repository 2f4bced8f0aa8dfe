// raster_backward.cu -- backward kernels of the B200-native Gaussian rasterizer (sm_100a).
//
//   render_backward_kernel   SPLIT CTAs per tile (default 2), back-to-front replay (DGR/cuda_rasterizer/backward.cu:415-601).
//                            Same staging as the forward (ids by TMA bulk copy, records gathered from the
//                            L2-resident per-Gaussian array one batch ahead, sub-tile culling). The ten
//                            per-(pixel,Gaussian) partial gradients are summed across the warp with a
//                            transposing butterfly (12 shuffles instead of 50) and leave the SM as ONE
//                            predicated RED.ADD.F32 instruction per (warp, instance) into a packed
//                            48-B-per-Gaussian accumulator (the reference issues 10 atomics per pixel hit).
//                            On the side every CTA clears its share of the dense gradient outputs (the kernel is
//                            issue-bound; the 118 MB of zero stores at C2 ride on an idle memory system).
//   gaussian_backward_kernel per VISIBLE Gaussian (the forward's compacted id list): computeCov2DCUDA +
//                            preprocessCUDA(bwd) fused (backward.cu:144-274, 346-412, 20-139, 278-341) with the
//                            Python-side confidence scaling (diff_gaussian_rasterization/__init__.py:147-157) in
//                            the epilogue; writes the rows of the visible Gaussians over the zeros.
#include <algorithm>
#include <cstdlib>
#include "raster_common.cuh"
#include "../../include/gvd_raster.h"

namespace {

// Transposing butterfly: in = 10 per-lane values; out = on even lanes with a valid slot, the sum over
// all 32 lanes of value `slot`. Returns slot (0..9) or -1.
__device__ __forceinline__ int warp_reduce10(float (&v)[10], uint32_t lane, float& out) {
    const unsigned FULL = 0xffffffffu;
    // step xor 16: 10 -> 5
    const bool b4 = lane & 16;
    float w[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const float send = b4 ? v[i] : v[i + 5];
        const float keep = b4 ? v[i + 5] : v[i];
        w[i] = keep + __shfl_xor_sync(FULL, send, 16);
    }
    // step xor 8: 5 -> 3   (w[5] == 0 implicitly)
    const bool b3 = lane & 8;
    float x[3];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = b3 ? w[i] : w[i + 3];
        const float keep = b3 ? w[i + 3] : w[i];
        x[i] = keep + __shfl_xor_sync(FULL, send, 8);
    }
    {
        const float send = b3 ? w[2] : 0.0f;
        const float keep = b3 ? 0.0f : w[2];
        x[2] = keep + __shfl_xor_sync(FULL, send, 8);
    }
    // step xor 4: 3 -> 2   (x[3] == 0 implicitly)
    const bool b2 = lane & 4;
    float y[2];
    {
        const float send = b2 ? x[0] : x[2];
        const float keep = b2 ? x[2] : x[0];
        y[0] = keep + __shfl_xor_sync(FULL, send, 4);
    }
    {
        const float send = b2 ? x[1] : 0.0f;
        const float keep = b2 ? 0.0f : x[1];
        y[1] = keep + __shfl_xor_sync(FULL, send, 4);
    }
    // step xor 2: 2 -> 1
    const bool b1 = lane & 2;
    float z;
    {
        const float send = b1 ? y[0] : y[1];
        const float keep = b1 ? y[1] : y[0];
        z = keep + __shfl_xor_sync(FULL, send, 2);
    }
    // step xor 1
    out = z + __shfl_xor_sync(FULL, z, 1);
    // which value does this lane hold?
    int local;
    if (!b3)
        local = b2 ? (b1 ? -1 : 2) : (b1 ? 1 : 0);
    else
        local = b2 ? -1 : (b1 ? 4 : 3);
    if (local < 0 || (lane & 1)) return -1;
    return local + (b4 ? 5 : 0);
}

// MOM: accumulate the raw moments  sum w, sum w dx, sum w dy, sum w dx^2, sum w dx dy, sum w dy^2  (w = G dL/dalpha-term) per
// Gaussian and let gaussian_backward apply the conic / opacity factors once per Gaussian instead of once per pixel
// (14 of ~155 instructions per visited (warp, instance); tools/moments_error.py: 2e-6 relative on the mean2D gradients).
template <int SPLIT, bool EXACT_DIV, bool MOM>
__global__ void __launch_bounds__(256 / SPLIT, 3 * SPLIT) render_backward_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, const SplatRec* __restrict__ splat,
    int W, int H, uint32_t tiles_x, const float* __restrict__ bg_color, const float* __restrict__ alphas,
    const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpixels,
    const float* __restrict__ dL_dpixel_depths, const float* __restrict__ dL_dalphas, float* __restrict__ acc,
    float4* __restrict__ zero, size_t zero_n4) {
    pdl_wait();
    pdl_trigger();
    // SPLIT CTAs share one 16x16 tile (8 / SPLIT warps of 8x4 pixels each): shorter CTAs, finer early exit, smaller tail
    constexpr int BLOCK = 256 / SPLIT, BATCH = BLOCK;
    __shared__ __align__(128) float4 buf[2][BATCH * 3];
    __shared__ __align__(128) IdSlot ids[3];
    __shared__ __align__(8) uint64_t bar[3];
    __shared__ uint32_t warp_max[BLOCK / 32];

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (zero_n4) {  // this CTA's slice of the dense gradient outputs (streaming stores; nothing here reads them)
        const size_t per = (zero_n4 + gridDim.x - 1) / gridDim.x;
        const size_t i0 = (size_t)blockIdx.x * per, i1 = i0 + per < zero_n4 ? i0 + per : zero_n4;
        for (size_t i = i0 + tid; i < i1; i += BLOCK) __stcs(zero + i, make_float4(0.f, 0.f, 0.f, 0.f));
    }
    const uint32_t tile = blockIdx.x / SPLIT;
    const uint32_t gw = (blockIdx.x % SPLIT) * (8 / SPLIT) + warp;  // this warp's 8x4 block inside the tile
    const uint32_t tile_x = tile % tiles_x, tile_y = tile / tiles_x;
    const uint32_t sub_x = tile_x * GVD_TILE_X + (gw & 1) * 8, sub_y = tile_y * GVD_TILE_Y + (gw >> 1) * 4;
    const uint32_t px = sub_x + (lane & 7), py = sub_y + (lane >> 3);
    const bool inside = px < (uint32_t)W && py < (uint32_t)H;
    const uint32_t pix_id = W * py + px;
    const float2 pixf = {(float)px, (float)py};
    const float sxf = (float)sub_x, syf = (float)sub_y;

    const uint2 range = ranges[tile];
    const int n_all = (int)(range.y - range.x);

    // The forward stored, per pixel, how far into the list it got (backward.cu:471-472).
    const uint32_t last_contributor = inside ? n_contrib[pix_id] : 0u;
    const uint32_t warp_last = __reduce_max_sync(0xffffffffu, last_contributor);
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_init(&bar[2], 1);
        mbar_fence_init();
    }
    if (lane == 0) warp_max[warp] = warp_last;
    __syncthreads();
    uint32_t block_last = 0;
#pragma unroll
    for (int w = 0; w < BLOCK / 32; ++w) block_last = max(block_last, warp_max[w]);
    // Entries at positions >= block_last are skipped by every pixel: never stage them.
    const int n = min(n_all, (int)block_last);
    const int rounds = (n + BATCH - 1) / BATCH;
    const uint32_t* list = point_list + range.x;
    // batch i (counted from the back) covers list positions [bstart(i), bstart(i) + bcnt(i))
    auto bstart = [n](int i) { return max(0, n - (i + 1) * BATCH); };
    auto bcnt = [n, &bstart](int i) { return (n - i * BATCH) - bstart(i); };

    if (tid == 0) {
        if (rounds > 0) issue_id_copy(&ids[0], &bar[0], list + bstart(0), bcnt(0));
        if (rounds > 1) issue_id_copy(&ids[1], &bar[1], list + bstart(1), bcnt(1));
    }
    if (rounds > 0) {
        mbar_wait(&bar[0], 0);
        if ((int)tid < bcnt(0)) {
            const float4* src = reinterpret_cast<const float4*>(splat + ids[0].v[id_lead(list + bstart(0)) + tid]);
            buf[0][tid * 3 + 0] = __ldg(src);
            buf[0][tid * 3 + 1] = __ldg(src + 1);
            buf[0][tid * 3 + 2] = __ldg(src + 2);
        }
    }

    const size_t HW = (size_t)H * W;
    const float T_final = inside ? (1 - alphas[pix_id]) : 0;
    float T = T_final;
    float accum_rec0 = 0, accum_rec1 = 0, accum_rec2 = 0, accum_depth_rec = 0, accum_alpha_rec = 0;
    float dL_dpixel0 = 0, dL_dpixel1 = 0, dL_dpixel2 = 0, dL_dpixel_depth = 0, dL_dalpha = 0;
    if (inside) {
        dL_dpixel0 = dL_dpixels[0 * HW + pix_id];
        dL_dpixel1 = dL_dpixels[1 * HW + pix_id];
        dL_dpixel2 = dL_dpixels[2 * HW + pix_id];
        dL_dpixel_depth = dL_dpixel_depths[pix_id];
        dL_dalpha = dL_dalphas[pix_id];
    }
    float last_alpha = 0, last_color0 = 0, last_color1 = 0, last_color2 = 0, last_depth = 0;
    const float ddelx_dx = 0.5 * W;
    const float ddely_dy = 0.5 * H;
    float bg_dot_dpixel = 0;
    bg_dot_dpixel += bg_color[0] * dL_dpixel0;
    bg_dot_dpixel += bg_color[1] * dL_dpixel1;
    bg_dot_dpixel += bg_color[2] * dL_dpixel2;

    for (int i = 0; i < rounds; ++i) {
        __syncthreads();  // publishes buf[i&1]; everyone is done with buf[(i+1)&1] and ids[(i+2)%3]
        const int cur = i & 1;
        const int start = bstart(i);
        const int cnt = bcnt(i);
        float4 pa, pb, pc;
        const bool have_next = (i + 1 < rounds) && ((int)tid < bcnt(i + 1));
        if (i + 1 < rounds) {
            mbar_wait(&bar[(i + 1) % 3], (uint32_t)(((i + 1) / 3) & 1));
            if (have_next) {
                const float4* src = reinterpret_cast<const float4*>(
                    splat + ids[(i + 1) % 3].v[id_lead(list + bstart(i + 1)) + tid]);
                pa = __ldg(src);
                pb = __ldg(src + 1);
                pc = __ldg(src + 2);
            }
            if (tid == 0 && i + 2 < rounds)
                issue_id_copy(&ids[(i + 2) % 3], &bar[(i + 2) % 3], list + bstart(i + 2), bcnt(i + 2));
        }

        if ((uint32_t)start < warp_last) {  // warp-uniform: otherwise nothing in this batch for us
            const float4* rec = buf[cur];
            const uint32_t* idv = ids[i % 3].v + id_lead(list + start);
            for (int chunk = (cnt - 1) / 32; chunk >= 0; --chunk) {
                const int e = chunk * 32 + (int)lane;
                bool hit = false;
                if (e < cnt && (uint32_t)(start + e) < warp_last) {
                    const float4 ea = rec[e * 3], ec = rec[e * 3 + 2];
                    hit = subtile_hit(ea.x, ea.y, ec.z, ec.w, sxf, syf);
                }
                unsigned m = __ballot_sync(0xffffffffu, hit);
                while (m) {
                    const int bsel = 31 - __clz(m);
                    m &= ~(1u << bsel);
                    const int j = chunk * 32 + bsel;
                    const uint32_t pos = (uint32_t)(start + j);
                    const float4 ra = rec[j * 3], rb = rec[j * 3 + 1], rc = rec[j * 3 + 2];

                    // backward.cu:509-528
                    const float2 d = {ra.x - pixf.x, ra.y - pixf.y};
                    const float power = -0.5f * (ra.z * d.x * d.x + rb.x * d.y * d.y) - ra.w * d.x * d.y;
                    const float G = expf(power);
                    const float alpha = fminf(0.99f, rb.y * G);
                    const bool contrib = (pos < last_contributor) && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
                    if (!__any_sync(0xffffffffu, contrib)) continue;

                    float g[10];
#pragma unroll
                    for (int k = 0; k < 10; ++k) g[k] = 0.0f;
                    if (contrib) {
                        // backward.cu:530,576 divide twice by (1 - alpha). EXACT_DIV keeps the two IEEE divisions, so the
                        // transmittance chain is bit-identical to the reference's; otherwise one IEEE reciprocal feeds
                        // both quotients (each within 1 ulp of the division, ~10 instructions fewer per contribution)
                        const float one_m_alpha = 1.f - alpha;
                        const float rcp_1ma = 1.0f / one_m_alpha;
                        T = EXACT_DIV ? T / one_m_alpha : T * rcp_1ma;
                        const float dchannel_dcolor = alpha * T;
                        const float dpixel_depth_ddepth = alpha * T;
                        float dL_dopa = 0.0f;
                        // colours (backward.cu:538-551)
                        accum_rec0 = last_alpha * last_color0 + (1.f - last_alpha) * accum_rec0;
                        last_color0 = rb.z;
                        dL_dopa += (rb.z - accum_rec0) * dL_dpixel0;
                        g[6] = dchannel_dcolor * dL_dpixel0;
                        accum_rec1 = last_alpha * last_color1 + (1.f - last_alpha) * accum_rec1;
                        last_color1 = rb.w;
                        dL_dopa += (rb.w - accum_rec1) * dL_dpixel1;
                        g[7] = dchannel_dcolor * dL_dpixel1;
                        accum_rec2 = last_alpha * last_color2 + (1.f - last_alpha) * accum_rec2;
                        last_color2 = rc.x;
                        dL_dopa += (rc.x - accum_rec2) * dL_dpixel2;
                        g[8] = dchannel_dcolor * dL_dpixel2;
                        // depth (backward.cu:553-563)
                        const float c_d = rc.y;
                        accum_depth_rec = last_alpha * last_depth + (1.f - last_alpha) * accum_depth_rec;
                        last_depth = c_d;
                        dL_dopa += (c_d - accum_depth_rec) * dL_dpixel_depth;
                        g[9] = dpixel_depth_ddepth * dL_dpixel_depth;
                        // alpha (backward.cu:565-567)
                        accum_alpha_rec = last_alpha + (1.f - last_alpha) * accum_alpha_rec;
                        dL_dopa += (1 - accum_alpha_rec) * dL_dalpha;

                        dL_dopa *= T;
                        last_alpha = alpha;
                        // background (backward.cu:573-578)
                        dL_dopa += (EXACT_DIV ? -T_final / one_m_alpha : -T_final * rcp_1ma) * bg_dot_dpixel;

                        if (MOM) {
                            const float w = G * dL_dopa;
                            const float wdx = w * d.x, wdy = w * d.y;
                            g[0] = wdx;
                            g[1] = wdy;
                            g[2] = wdx * d.x;
                            g[3] = wdx * d.y;
                            g[4] = wdy * d.y;
                            g[5] = w;
                        } else {
                            const float dL_dG = rb.y * dL_dopa;
                            const float gdx = G * d.x;
                            const float gdy = G * d.y;
                            const float dG_ddelx = -gdx * ra.z - gdy * ra.w;
                            const float dG_ddely = -gdy * rb.x - gdx * ra.w;
                            g[0] = dL_dG * dG_ddelx * ddelx_dx;
                            g[1] = dL_dG * dG_ddely * ddely_dy;
                            g[2] = -0.5f * gdx * d.x * dL_dG;
                            g[3] = -0.5f * gdx * d.y * dL_dG;
                            g[4] = -0.5f * gdy * d.y * dL_dG;
                            g[5] = G * dL_dopa;
                        }
                    }
                    float total;
                    const int slot = warp_reduce10(g, lane, total);
                    if (slot >= 0) atomicAdd(acc + (size_t)idv[j] * GVD_ACC_STRIDE + slot, total);
                }
            }
        }
        if (have_next) {
            buf[cur ^ 1][tid * 3 + 0] = pa;
            buf[cur ^ 1][tid * 3 + 1] = pb;
            buf[cur ^ 1][tid * 3 + 2] = pc;
        }
    }
}

// ------------------------------------------------------------------------------------------
// auxiliary.h:107-117
__device__ __forceinline__ float3 dnormvdv3(float3 v, float3 dv) {
    float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
    float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
    float3 r;
    r.x = ((+sum2 - v.x * v.x) * dv.x - v.y * v.x * dv.y - v.z * v.x * dv.z) * invsum32;
    r.y = (-v.x * v.y * dv.x + (sum2 - v.y * v.y) * dv.y - v.z * v.y * dv.z) * invsum32;
    r.z = (-v.x * v.z * dv.x - v.y * v.z * dv.y + (sum2 - v.z * v.z) * dv.z) * invsum32;
    return r;
}

#define SH(k) make_float3(v[3 * (k)], v[3 * (k) + 1], v[3 * (k) + 2])
#define OUT(k, val)                                 \
    {                                               \
        const float3 _t = f3_scale((val), dL_dRGB); \
        o_sh[3 * (k)] = _t.x * conf;                \
        o_sh[3 * (k) + 1] = _t.y * conf;            \
        o_sh[3 * (k) + 2] = _t.z * conf;            \
    }

template <int MIN_CTAS, bool MOM, bool RAW>  // RAW: gradients with respect to the un-activated GaussianModel parameters (compile-time:
                                             // the standard instantiation keeps the reference's expression trees)
__global__ void __launch_bounds__(256, MIN_CTAS) gaussian_backward_kernel(
    const SplatRec* __restrict__ splat, int W, int H, int D, int M, const float3* __restrict__ means, const uint32_t* __restrict__ vis_id, const uint32_t* __restrict__ counts,
    const float* __restrict__ shs, const uint8_t* __restrict__ clamped, const float3* __restrict__ scales,
    const float4* __restrict__ rotations, const float scale_modifier, const float* __restrict__ cov3D_precomp,
    const float* __restrict__ view, const float* __restrict__ proj, const float h_x, float h_y, const float tan_fovx,
    float tan_fovy, const float3* __restrict__ campos, const float* __restrict__ acc,
    const float* __restrict__ confidence, float* __restrict__ dL_dmeans2D, float* __restrict__ dL_dmeans3D,
    float* __restrict__ dL_dopacity, float* __restrict__ dL_dcolors, float* __restrict__ dL_dcov3D,
    float* __restrict__ dL_dsh, float* __restrict__ dL_dscales, float* __restrict__ dL_drots,
    const float* __restrict__ shs_rest, const float* __restrict__ opacities_raw, float* __restrict__ dL_dsh_rest) {
    pdl_wait();
    pdl_trigger();
    // One thread per visible Gaussian, in ascending id order (vis_id comes from the forward's compaction), so the
    // gathers of a warp stay close together. Everything else was zero-filled before this kernel.
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= counts[0]) return;
    const int idx = (int)vis_id[i];

    float2 o_m2 = {0.f, 0.f};
    float3 o_m3 = {0.f, 0.f, 0.f}, o_sc = {0.f, 0.f, 0.f}, o_col = {0.f, 0.f, 0.f};
    float4 o_rot = {0.f, 0.f, 0.f, 0.f};
    float o_op = 0.f;
    float o_cov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float o_sh[48];
#pragma unroll
    for (int k = 0; k < 48; ++k) o_sh[k] = 0.f;

    {
        const float conf = confidence ? confidence[idx] : 1.0f;
        const float4* ap = reinterpret_cast<const float4*>(acc + (size_t)idx * GVD_ACC_STRIDE);
        const float4 a0 = ap[0], a1 = ap[1], a2 = ap[2];
        float2 dL_dmean2D = {a0.x, a0.y};
        float3 dL_dconic = {a0.z, a0.w, a1.x};
        const float dL_dopac = a1.y;
        if (MOM) {  // the accumulator holds raw moments (see render_backward_kernel): apply conic / opacity here, once
            const float4* rec = reinterpret_cast<const float4*>(splat + idx);
            const float4 ra = __ldg(rec), rb = __ldg(rec + 1);  // ra.z, ra.w, rb.x = conic; rb.y = opacity (as composited)
            const float o = rb.y;
            dL_dmean2D.x = -o * (0.5f * W) * (ra.z * a0.x + ra.w * a0.y);
            dL_dmean2D.y = -o * (0.5f * H) * (rb.x * a0.y + ra.w * a0.x);
            dL_dconic.x = -0.5f * o * a0.z;
            dL_dconic.y = -0.5f * o * a0.w;
            dL_dconic.z = -0.5f * o * a1.x;
        }
        const float3 dL_dcolor = {a1.z, a1.w, a2.x};
        const float dL_ddepth = a2.y;

        // ---------------- computeCov2DCUDA (backward.cu:144-274) ----------------
        float cov3D_local[6];
        const float* cov3D;
        if (cov3D_precomp != nullptr)
            cov3D = cov3D_precomp + 6 * (size_t)idx;
        else {
            if (RAW) cov3d_from_scale_rot(act_scale(scales[idx]), scale_modifier, act_rotation(rotations[idx]), cov3D_local);
            else cov3d_from_scale_rot(scales[idx], scale_modifier, rotations[idx], cov3D_local);
            cov3D = cov3D_local;
        }
        const float3 mean = means[idx];
        float3 t = xform_point_4x3(mean, view);
        const float limx = 1.3f * tan_fovx;
        const float limy = 1.3f * tan_fovy;
        const float txtz = t.x / t.z;
        const float tytz = t.y / t.z;
        t.x = fminf(limx, fmaxf(-limx, txtz)) * t.z;
        t.y = fminf(limy, fmaxf(-limy, tytz)) * t.z;
        const float x_grad_mul = txtz < -limx || txtz > limx ? 0 : 1;
        const float y_grad_mul = tytz < -limy || tytz > limy ? 0 : 1;

        M3 J = m3_make(h_x / t.z, 0.0f, -(h_x * t.x) / (t.z * t.z), 0.0f, h_y / t.z, -(h_y * t.y) / (t.z * t.z), 0, 0, 0);
        M3 Wm = m3_make(view[0], view[4], view[8], view[1], view[5], view[9], view[2], view[6], view[10]);
        M3 Vrk = m3_make(cov3D[0], cov3D[1], cov3D[2], cov3D[1], cov3D[3], cov3D[4], cov3D[2], cov3D[4], cov3D[5]);
        M3 Tm = m3_mul(Wm, J);
        M3 cov2D = m3_mul(m3_mul(m3_transpose(Tm), m3_transpose(Vrk)), Tm);

        const float a = cov2D.m[0][0] += 0.3f;
        const float b = cov2D.m[0][1];
        const float c = cov2D.m[1][1] += 0.3f;
        const float denom = a * c - b * b;
        float dL_da = 0, dL_db = 0, dL_dc = 0;
        const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
        float dcov[6];
        if (denom2inv != 0) {
            dL_da = denom2inv * (-c * c * dL_dconic.x + 2 * b * c * dL_dconic.y + (denom - a * c) * dL_dconic.z);
            dL_dc = denom2inv * (-a * a * dL_dconic.z + 2 * a * b * dL_dconic.y + (denom - a * c) * dL_dconic.x);
            dL_db = denom2inv * 2 * (b * c * dL_dconic.x - (denom + 2 * b * b) * dL_dconic.y + a * b * dL_dconic.z);

            dcov[0] = (Tm.m[0][0] * Tm.m[0][0] * dL_da + Tm.m[0][0] * Tm.m[1][0] * dL_db + Tm.m[1][0] * Tm.m[1][0] * dL_dc);
            dcov[3] = (Tm.m[0][1] * Tm.m[0][1] * dL_da + Tm.m[0][1] * Tm.m[1][1] * dL_db + Tm.m[1][1] * Tm.m[1][1] * dL_dc);
            dcov[5] = (Tm.m[0][2] * Tm.m[0][2] * dL_da + Tm.m[0][2] * Tm.m[1][2] * dL_db + Tm.m[1][2] * Tm.m[1][2] * dL_dc);
            dcov[1] = 2 * Tm.m[0][0] * Tm.m[0][1] * dL_da + (Tm.m[0][0] * Tm.m[1][1] + Tm.m[0][1] * Tm.m[1][0]) * dL_db +
                      2 * Tm.m[1][0] * Tm.m[1][1] * dL_dc;
            dcov[2] = 2 * Tm.m[0][0] * Tm.m[0][2] * dL_da + (Tm.m[0][0] * Tm.m[1][2] + Tm.m[0][2] * Tm.m[1][0]) * dL_db +
                      2 * Tm.m[1][0] * Tm.m[1][2] * dL_dc;
            dcov[4] = 2 * Tm.m[0][2] * Tm.m[0][1] * dL_da + (Tm.m[0][1] * Tm.m[1][2] + Tm.m[0][2] * Tm.m[1][1]) * dL_db +
                      2 * Tm.m[1][1] * Tm.m[1][2] * dL_dc;
        } else {
    #pragma unroll
            for (int k = 0; k < 6; ++k) dcov[k] = 0;
        }

        const float dL_dT00 = 2 * (Tm.m[0][0] * Vrk.m[0][0] + Tm.m[0][1] * Vrk.m[0][1] + Tm.m[0][2] * Vrk.m[0][2]) * dL_da +
                              (Tm.m[1][0] * Vrk.m[0][0] + Tm.m[1][1] * Vrk.m[0][1] + Tm.m[1][2] * Vrk.m[0][2]) * dL_db;
        const float dL_dT01 = 2 * (Tm.m[0][0] * Vrk.m[1][0] + Tm.m[0][1] * Vrk.m[1][1] + Tm.m[0][2] * Vrk.m[1][2]) * dL_da +
                              (Tm.m[1][0] * Vrk.m[1][0] + Tm.m[1][1] * Vrk.m[1][1] + Tm.m[1][2] * Vrk.m[1][2]) * dL_db;
        const float dL_dT02 = 2 * (Tm.m[0][0] * Vrk.m[2][0] + Tm.m[0][1] * Vrk.m[2][1] + Tm.m[0][2] * Vrk.m[2][2]) * dL_da +
                              (Tm.m[1][0] * Vrk.m[2][0] + Tm.m[1][1] * Vrk.m[2][1] + Tm.m[1][2] * Vrk.m[2][2]) * dL_db;
        const float dL_dT10 = 2 * (Tm.m[1][0] * Vrk.m[0][0] + Tm.m[1][1] * Vrk.m[0][1] + Tm.m[1][2] * Vrk.m[0][2]) * dL_dc +
                              (Tm.m[0][0] * Vrk.m[0][0] + Tm.m[0][1] * Vrk.m[0][1] + Tm.m[0][2] * Vrk.m[0][2]) * dL_db;
        const float dL_dT11 = 2 * (Tm.m[1][0] * Vrk.m[1][0] + Tm.m[1][1] * Vrk.m[1][1] + Tm.m[1][2] * Vrk.m[1][2]) * dL_dc +
                              (Tm.m[0][0] * Vrk.m[1][0] + Tm.m[0][1] * Vrk.m[1][1] + Tm.m[0][2] * Vrk.m[1][2]) * dL_db;
        const float dL_dT12 = 2 * (Tm.m[1][0] * Vrk.m[2][0] + Tm.m[1][1] * Vrk.m[2][1] + Tm.m[1][2] * Vrk.m[2][2]) * dL_dc +
                              (Tm.m[0][0] * Vrk.m[2][0] + Tm.m[0][1] * Vrk.m[2][1] + Tm.m[0][2] * Vrk.m[2][2]) * dL_db;

        const float dL_dJ00 = Wm.m[0][0] * dL_dT00 + Wm.m[0][1] * dL_dT01 + Wm.m[0][2] * dL_dT02;
        const float dL_dJ02 = Wm.m[2][0] * dL_dT00 + Wm.m[2][1] * dL_dT01 + Wm.m[2][2] * dL_dT02;
        const float dL_dJ11 = Wm.m[1][0] * dL_dT10 + Wm.m[1][1] * dL_dT11 + Wm.m[1][2] * dL_dT12;
        const float dL_dJ12 = Wm.m[2][0] * dL_dT10 + Wm.m[2][1] * dL_dT11 + Wm.m[2][2] * dL_dT12;

        const float tz = 1.f / t.z;
        const float tz2 = tz * tz;
        const float tz3 = tz2 * tz;
        const float dL_dtx = x_grad_mul * -h_x * tz2 * dL_dJ02;
        const float dL_dty = y_grad_mul * -h_y * tz2 * dL_dJ12;
        const float dL_dtz = -h_x * tz2 * dL_dJ00 - h_y * tz2 * dL_dJ11 + (2 * h_x * t.x) * tz3 * dL_dJ02 +
                             (2 * h_y * t.y) * tz3 * dL_dJ12;
        // auxiliary.h:89-97 (transformVec4x3Transpose)
        float3 dL_dmean = {view[0] * dL_dtx + view[1] * dL_dty + view[2] * dL_dtz,
                           view[4] * dL_dtx + view[5] * dL_dty + view[6] * dL_dtz,
                           view[8] * dL_dtx + view[9] * dL_dty + view[10] * dL_dtz};

        // ---------------- preprocessCUDA backward (backward.cu:346-412) ----------------
        const float3 m = mean;
        const float4 m_hom = xform_point_4x4(m, proj);
        const float m_w = 1.0f / (m_hom.w + 0.0000001f);
        const float mul1 = (proj[0] * m.x + proj[4] * m.y + proj[8] * m.z + proj[12]) * m_w * m_w;
        const float mul2 = (proj[1] * m.x + proj[5] * m.y + proj[9] * m.z + proj[13]) * m_w * m_w;
        float3 dm;
        dm.x = (proj[0] * m_w - proj[3] * mul1) * dL_dmean2D.x + (proj[1] * m_w - proj[3] * mul2) * dL_dmean2D.y;
        dm.y = (proj[4] * m_w - proj[7] * mul1) * dL_dmean2D.x + (proj[5] * m_w - proj[7] * mul2) * dL_dmean2D.y;
        dm.z = (proj[8] * m_w - proj[11] * mul1) * dL_dmean2D.x + (proj[9] * m_w - proj[11] * mul2) * dL_dmean2D.y;
        dL_dmean = f3_add(dL_dmean, dm);

        // depth -> mean (backward.cu:391-403)
        const float mul3 = view[2] * m.x + view[6] * m.y + view[10] * m.z + view[14];
        float3 dm2;
        dm2.x = (view[2] - view[3] * mul3) * dL_ddepth;
        dm2.y = (view[6] - view[7] * mul3) * dL_ddepth;
        dm2.z = (view[10] - view[11] * mul3) * dL_ddepth;
        dL_dmean = f3_add(dL_dmean, dm2);

        // ---------------- SH backward (backward.cu:20-139) ----------------
        if (shs) {
            const float3 cp = *campos;
            const float3 dir_orig = f3_sub(m, cp);
            const float len = sqrtf(f3_dot(dir_orig, dir_orig));
            const float3 dir = {dir_orig.x / len, dir_orig.y / len, dir_orig.z / len};
            float v[48];
            if (RAW) load_sh_split(shs, shs_rest, idx, D, M, v);
            else load_sh(shs, idx, D, M, v);
            const uint8_t cl = clamped[idx];
            float3 dL_dRGB = dL_dcolor;
            dL_dRGB.x *= (cl & 1) ? 0 : 1;
            dL_dRGB.y *= (cl & 2) ? 0 : 1;
            dL_dRGB.z *= (cl & 4) ? 0 : 1;
            float3 dRGBdx = {0, 0, 0}, dRGBdy = {0, 0, 0}, dRGBdz = {0, 0, 0};
            const float x = dir.x, y = dir.y, z = dir.z;
            OUT(0, GVD_SH_C0);
            if (D > 0) {
                OUT(1, -GVD_SH_C1 * y);
                OUT(2, GVD_SH_C1 * z);
                OUT(3, -GVD_SH_C1 * x);
                dRGBdx = f3_scale(-GVD_SH_C1, SH(3));
                dRGBdy = f3_scale(-GVD_SH_C1, SH(1));
                dRGBdz = f3_scale(GVD_SH_C1, SH(2));
                if (D > 1) {
                    const float xx = x * x, yy = y * y, zz = z * z;
                    const float xy = x * y, yz = y * z, xz = x * z;
                    OUT(4, GVD_SH_C2_0 * xy);
                    OUT(5, GVD_SH_C2_1 * yz);
                    OUT(6, GVD_SH_C2_2 * (2.f * zz - xx - yy));
                    OUT(7, GVD_SH_C2_3 * xz);
                    OUT(8, GVD_SH_C2_4 * (xx - yy));

                    dRGBdx = f3_add(dRGBdx, f3_add(f3_add(f3_add(f3_scale(GVD_SH_C2_0 * y, SH(4)),
                                                                 f3_scale(GVD_SH_C2_2 * 2.f * -x, SH(6))),
                                                          f3_scale(GVD_SH_C2_3 * z, SH(7))),
                                                   f3_scale(GVD_SH_C2_4 * 2.f * x, SH(8))));
                    dRGBdy = f3_add(dRGBdy, f3_add(f3_add(f3_add(f3_scale(GVD_SH_C2_0 * x, SH(4)),
                                                                 f3_scale(GVD_SH_C2_1 * z, SH(5))),
                                                          f3_scale(GVD_SH_C2_2 * 2.f * -y, SH(6))),
                                                   f3_scale(GVD_SH_C2_4 * 2.f * -y, SH(8))));
                    dRGBdz = f3_add(dRGBdz, f3_add(f3_add(f3_scale(GVD_SH_C2_1 * y, SH(5)),
                                                          f3_scale(GVD_SH_C2_2 * 2.f * 2.f * z, SH(6))),
                                                   f3_scale(GVD_SH_C2_3 * x, SH(7))));
                    if (D > 2) {
                        OUT(9, GVD_SH_C3_0 * y * (3.f * xx - yy));
                        OUT(10, GVD_SH_C3_1 * xy * z);
                        OUT(11, GVD_SH_C3_2 * y * (4.f * zz - xx - yy));
                        OUT(12, GVD_SH_C3_3 * z * (2.f * zz - 3.f * xx - 3.f * yy));
                        OUT(13, GVD_SH_C3_4 * x * (4.f * zz - xx - yy));
                        OUT(14, GVD_SH_C3_5 * z * (xx - yy));
                        OUT(15, GVD_SH_C3_6 * x * (xx - 3.f * yy));

                        float3 ax = f3_scale(GVD_SH_C3_0 * 3.f * 2.f * xy, SH(9));
                        ax = f3_add(ax, f3_scale(GVD_SH_C3_1 * yz, SH(10)));
                        ax = f3_add(ax, f3_scale(GVD_SH_C3_2 * -2.f * xy, SH(11)));
                        ax = f3_add(ax, f3_scale(GVD_SH_C3_3 * -3.f * 2.f * xz, SH(12)));
                        ax = f3_add(ax, f3_scale(GVD_SH_C3_4 * (-3.f * xx + 4.f * zz - yy), SH(13)));
                        ax = f3_add(ax, f3_scale(GVD_SH_C3_5 * 2.f * xz, SH(14)));
                        ax = f3_add(ax, f3_scale(GVD_SH_C3_6 * 3.f * (xx - yy), SH(15)));
                        dRGBdx = f3_add(dRGBdx, ax);

                        float3 ay = f3_scale(GVD_SH_C3_0 * 3.f * (xx - yy), SH(9));
                        ay = f3_add(ay, f3_scale(GVD_SH_C3_1 * xz, SH(10)));
                        ay = f3_add(ay, f3_scale(GVD_SH_C3_2 * (-3.f * yy + 4.f * zz - xx), SH(11)));
                        ay = f3_add(ay, f3_scale(GVD_SH_C3_3 * -3.f * 2.f * yz, SH(12)));
                        ay = f3_add(ay, f3_scale(GVD_SH_C3_4 * -2.f * xy, SH(13)));
                        ay = f3_add(ay, f3_scale(GVD_SH_C3_5 * -2.f * yz, SH(14)));
                        ay = f3_add(ay, f3_scale(GVD_SH_C3_6 * -3.f * 2.f * xy, SH(15)));
                        dRGBdy = f3_add(dRGBdy, ay);

                        float3 az = f3_scale(GVD_SH_C3_1 * xy, SH(10));
                        az = f3_add(az, f3_scale(GVD_SH_C3_2 * 4.f * 2.f * yz, SH(11)));
                        az = f3_add(az, f3_scale(GVD_SH_C3_3 * 3.f * (2.f * zz - xx - yy), SH(12)));
                        az = f3_add(az, f3_scale(GVD_SH_C3_4 * 4.f * 2.f * xz, SH(13)));
                        az = f3_add(az, f3_scale(GVD_SH_C3_5 * (xx - yy), SH(14)));
                        dRGBdz = f3_add(dRGBdz, az);
                    }
                }
            }

            const float3 dL_ddir = {f3_dot(dRGBdx, dL_dRGB), f3_dot(dRGBdy, dL_dRGB), f3_dot(dRGBdz, dL_dRGB)};
            dL_dmean = f3_add(dL_dmean, dnormvdv3(dir_orig, dL_ddir));
        }

        // ---------------- cov3D backward (backward.cu:278-341) ----------------
        if (scales) {
            const float4 q_in = rotations[idx];
            const float4 q = RAW ? act_rotation(q_in) : q_in;
            const float r = q.x, x = q.y, y = q.z, z = q.w;
            M3 R = m3_make(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
                           2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
                           2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
            const float3 sc = RAW ? act_scale(scales[idx]) : scales[idx];
            const float3 s = {scale_modifier * sc.x, scale_modifier * sc.y, scale_modifier * sc.z};
            M3 Mm;
    #pragma unroll
            for (int cc = 0; cc < 3; ++cc) {
                Mm.m[cc][0] = s.x * R.m[cc][0];
                Mm.m[cc][1] = s.y * R.m[cc][1];
                Mm.m[cc][2] = s.z * R.m[cc][2];
            }
            M3 dL_dSigma = m3_make(dcov[0], 0.5f * dcov[1], 0.5f * dcov[2], 0.5f * dcov[1], dcov[3], 0.5f * dcov[4],
                                   0.5f * dcov[2], 0.5f * dcov[4], dcov[5]);
            // dL_dM = 2 * M * dL_dSigma   (scalar*matrix first, then product, as written in the reference)
            M3 M2;
    #pragma unroll
            for (int cc = 0; cc < 3; ++cc)
    #pragma unroll
                for (int rr = 0; rr < 3; ++rr) M2.m[cc][rr] = Mm.m[cc][rr] * 2.0f;
            M3 dL_dM = m3_mul(M2, dL_dSigma);
            M3 Rt = m3_transpose(R);
            M3 dL_dMt = m3_transpose(dL_dM);

            float3 dscale;
            dscale.x = Rt.m[0][0] * dL_dMt.m[0][0] + Rt.m[0][1] * dL_dMt.m[0][1] + Rt.m[0][2] * dL_dMt.m[0][2];
            dscale.y = Rt.m[1][0] * dL_dMt.m[1][0] + Rt.m[1][1] * dL_dMt.m[1][1] + Rt.m[1][2] * dL_dMt.m[1][2];
            dscale.z = Rt.m[2][0] * dL_dMt.m[2][0] + Rt.m[2][1] * dL_dMt.m[2][1] + Rt.m[2][2] * dL_dMt.m[2][2];
            if (RAW) {  // d exp(s) / ds = exp(s)
                dscale.x *= sc.x;
                dscale.y *= sc.y;
                dscale.z *= sc.z;
            }
            o_sc = {dscale.x * conf, dscale.y * conf, dscale.z * conf};

    #pragma unroll
            for (int rr = 0; rr < 3; ++rr) {
                dL_dMt.m[0][rr] *= s.x;
                dL_dMt.m[1][rr] *= s.y;
                dL_dMt.m[2][rr] *= s.z;
            }
            float4 dq;
            dq.x = 2 * z * (dL_dMt.m[0][1] - dL_dMt.m[1][0]) + 2 * y * (dL_dMt.m[2][0] - dL_dMt.m[0][2]) +
                   2 * x * (dL_dMt.m[1][2] - dL_dMt.m[2][1]);
            dq.y = 2 * y * (dL_dMt.m[1][0] + dL_dMt.m[0][1]) + 2 * z * (dL_dMt.m[2][0] + dL_dMt.m[0][2]) +
                   2 * r * (dL_dMt.m[1][2] - dL_dMt.m[2][1]) - 4 * x * (dL_dMt.m[2][2] + dL_dMt.m[1][1]);
            dq.z = 2 * x * (dL_dMt.m[1][0] + dL_dMt.m[0][1]) + 2 * r * (dL_dMt.m[2][0] - dL_dMt.m[0][2]) +
                   2 * z * (dL_dMt.m[1][2] + dL_dMt.m[2][1]) - 4 * y * (dL_dMt.m[2][2] + dL_dMt.m[0][0]);
            dq.w = 2 * r * (dL_dMt.m[0][1] - dL_dMt.m[1][0]) + 2 * x * (dL_dMt.m[2][0] + dL_dMt.m[0][2]) +
                   2 * y * (dL_dMt.m[1][2] + dL_dMt.m[2][1]) - 4 * z * (dL_dMt.m[1][1] + dL_dMt.m[0][0]);
            if (RAW) dq = dnormvdv4(q_in, dq);  // through q / ||q||
            o_rot = make_float4(dq.x * conf, dq.y * conf, dq.z * conf, dq.w * conf);
        }


        o_m2 = dL_dmean2D;  // not confidence-scaled (diff_gaussian_rasterization/__init__.py:149)
        o_m3 = {dL_dmean.x * conf, dL_dmean.y * conf, dL_dmean.z * conf};
        o_op = dL_dopac * conf;
        if (RAW) {  // d sigmoid(o) / do = sigma (1 - sigma)
            const float sg = act_opacity(opacities_raw[idx]);
            o_op *= sg * (1.0f - sg);
        }
        o_col = {dL_dcolor.x * conf, dL_dcolor.y * conf, dL_dcolor.z * conf};
#pragma unroll
        for (int k = 0; k < 6; ++k) o_cov[k] = dcov[k] * conf;
    }

    // ---------------- stores: the rows of this Gaussian ----------------
    dL_dmeans2D[3 * (size_t)idx] = o_m2.x;
    dL_dmeans2D[3 * (size_t)idx + 1] = o_m2.y;
    dL_dmeans3D[3 * (size_t)idx] = o_m3.x;
    dL_dmeans3D[3 * (size_t)idx + 1] = o_m3.y;
    dL_dmeans3D[3 * (size_t)idx + 2] = o_m3.z;
    dL_dopacity[idx] = o_op;
    if (dL_dscales) {
        dL_dscales[3 * (size_t)idx] = o_sc.x;
        dL_dscales[3 * (size_t)idx + 1] = o_sc.y;
        dL_dscales[3 * (size_t)idx + 2] = o_sc.z;
    }
    if (dL_dcolors) {
        dL_dcolors[3 * (size_t)idx] = o_col.x;
        dL_dcolors[3 * (size_t)idx + 1] = o_col.y;
        dL_dcolors[3 * (size_t)idx + 2] = o_col.z;
    }
    if (dL_drots) reinterpret_cast<float4*>(dL_drots)[idx] = o_rot;
    if (dL_dcov3D) {
#pragma unroll
        for (int k = 0; k < 6; ++k) dL_dcov3D[6 * (size_t)idx + k] = o_cov[k];
    }
    if (dL_dsh && RAW) {  // d/d_features_dc [P,1,3] and d/d_features_rest [P,M-1,3]: two tensors, no split afterwards
        float* dc = dL_dsh + 3 * (size_t)idx;
        dc[0] = o_sh[0];
        dc[1] = o_sh[1];
        dc[2] = o_sh[2];
        float* row = dL_dsh_rest + (size_t)idx * 3 * (M - 1);
#pragma unroll
        for (int k = 3; k < 48; ++k)
            if (k < 3 * M) row[k - 3] = o_sh[k];
    } else if (dL_dsh) {
        float* row = dL_dsh + (size_t)idx * 3 * M;
        if (M == 16 && ((reinterpret_cast<uintptr_t>(dL_dsh) & 15) == 0)) {
            float4* r4 = reinterpret_cast<float4*>(row);  // one 192-byte row: six full sectors
#pragma unroll
            for (int k = 0; k < 12; ++k) r4[k] = make_float4(o_sh[4 * k], o_sh[4 * k + 1], o_sh[4 * k + 2], o_sh[4 * k + 3]);
        } else {
#pragma unroll
            for (int k = 0; k < 48; ++k)
                if (k < 3 * M) row[k] = o_sh[k];
        }
    }
}
#undef SH
#undef OUT

}  // namespace

// Zero-fill of the backward accumulator as a kernel: cudaMemsetAsync may be served by a copy engine and then queues
// behind any host<->device transfer in flight on another stream.
__global__ void __launch_bounds__(256) zero_fill_kernel(float4* __restrict__ p, size_t n4) {
    pdl_wait();
    pdl_trigger();
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

void gvd_launch_zero_fill(float* p, size_t floats, cudaStream_t s) {
    const size_t n4 = floats / 4;  // GVD_ACC_STRIDE is a multiple of 4 floats and the buffer is 128-byte aligned
    const unsigned blocks = (unsigned)((n4 + 255) / 256 < 148 * 16 ? (n4 + 255) / 256 : 148 * 16);
    if (blocks) gvd_launch(zero_fill_kernel, dim3(blocks), dim3(256), 0, s, reinterpret_cast<float4*>(p), n4);
}

__global__ void __launch_bounds__(256) zero_words_kernel(uint32_t* __restrict__ p, size_t n) {
    pdl_wait();
    pdl_trigger();
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) p[i] = 0u;
}

// Generic clear (gradient tensors are float arrays: 4-byte aligned, size a multiple of 4): head words up to the first
// 16-byte boundary, float4 body, tail words.
void gvd_launch_zero_bytes(void* p, size_t bytes, cudaStream_t s) {
    if (!p || bytes < 4) return;
    char* c = reinterpret_cast<char*>(p);
    const size_t head = std::min<size_t>(bytes, (16 - (reinterpret_cast<uintptr_t>(c) & 15)) & 15);
    const size_t body = (bytes - head) & ~(size_t)15, tail = bytes - head - body;
    if (head) gvd_launch(zero_words_kernel, dim3(1), dim3(256), 0, s, reinterpret_cast<uint32_t*>(c), head / 4);
    if (body) gvd_launch_zero_fill(reinterpret_cast<float*>(c + head), body / 4, s);
    if (tail) gvd_launch(zero_words_kernel, dim3(1), dim3(256), 0, s, reinterpret_cast<uint32_t*>(c + head + body), tail / 4);
}

// GVD_BWD_MOMENTS=0: per-pixel conic factors as in the reference (A/B switch); default: raw moments.  Both backward
// kernels of a frame read the same switch.
static bool gvd_bwd_moments() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("GVD_BWD_MOMENTS");
        const char* x = getenv("GVD_BWD_EXACT_DIV");
        on = ((e && e[0] == '0') || (x && x[0] == '1')) ? 0 : 1;
    }
    return on == 1;
}

template <int SPLIT>
static void launch_render_backward(const GvdRasterBackwardArgs& a, const RasterGeomPtrs& g, const RasterBinPtrs& b,
                                   const RasterImgPtrs& im, float* acc, float4* zero, size_t zero_n4, dim3 grid, cudaStream_t s) {
    // GVD_BWD_EXACT_DIV=1: the reference's two IEEE divisions per contribution instead of one reciprocal + two
    // multiplications. Measured: +12 us (200 -> 212 us at C2), no change in any parity figure; off by default (and it
    // implies the per-pixel conic factors, i.e. no moments).
    static int exact = -1;
    if (exact < 0) {
        const char* e = getenv("GVD_BWD_EXACT_DIV");
        exact = (e && e[0] == '1') ? 1 : 0;
        cudaFuncSetAttribute((const void*)render_backward_kernel<SPLIT, true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        cudaFuncSetAttribute((const void*)render_backward_kernel<SPLIT, false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        cudaFuncSetAttribute((const void*)render_backward_kernel<SPLIT, false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    }
    if (exact)
        gvd_launch(render_backward_kernel<SPLIT, true, false>, dim3(grid.x * grid.y * SPLIT), dim3(256 / SPLIT), 0, s, im.ranges, b.point_list,
                   g.splat, a.width, a.height, grid.x, a.background, a.alphas, im.n_contrib, a.dL_dpix, a.dL_ddepth_pix,
                   a.dL_dalpha_pix, acc, zero, zero_n4);
    else if (gvd_bwd_moments())
        gvd_launch(render_backward_kernel<SPLIT, false, true>, dim3(grid.x * grid.y * SPLIT), dim3(256 / SPLIT), 0, s, im.ranges, b.point_list,
                   g.splat, a.width, a.height, grid.x, a.background, a.alphas, im.n_contrib, a.dL_dpix, a.dL_ddepth_pix,
                   a.dL_dalpha_pix, acc, zero, zero_n4);
    else
        gvd_launch(render_backward_kernel<SPLIT, false, false>, dim3(grid.x * grid.y * SPLIT), dim3(256 / SPLIT), 0, s, im.ranges, b.point_list,
                   g.splat, a.width, a.height, grid.x, a.background, a.alphas, im.n_contrib, a.dL_dpix, a.dL_ddepth_pix,
                   a.dL_dalpha_pix, acc, zero, zero_n4);
}

void gvd_launch_render_backward(const GvdRasterBackwardArgs& a, const RasterGeomPtrs& g, const RasterBinPtrs& b,
                                const RasterImgPtrs& im, float* acc, float4* zero, size_t zero_n4, dim3 grid, cudaStream_t s) {
    switch (gvd_render_split()) {
        case 1: launch_render_backward<1>(a, g, b, im, acc, zero, zero_n4, grid, s); break;
        case 4: launch_render_backward<4>(a, g, b, im, acc, zero, zero_n4, grid, s); break;
        default: launch_render_backward<2>(a, g, b, im, acc, zero, zero_n4, grid, s); break;
    }
}

template <int MIN_CTAS>
static void launch_gaussian_backward(const GvdRasterBackwardArgs& a, const RasterGeomPtrs& g, const float* acc,
                                     float focal_x, float focal_y, int num_visible, cudaStream_t s) {
    const int n = num_visible >= 0 ? num_visible : a.P;  // V unknown on the host: cover P, surplus CTAs return at once
    if (n <= 0) return;
    auto go = [&](auto kernel) {
        gvd_launch(kernel, dim3((n + 255) / 256), dim3(256), 0, s, g.splat, a.width, a.height,
            a.D, a.M, (const float3*)a.means3D, g.vis_id, g.counts, a.shs, g.clamped, (const float3*)a.scales,
            (const float4*)a.rotations, a.scale_modifier, a.cov3D_precomp, a.viewmatrix, a.projmatrix, focal_x, focal_y,
            a.tan_fovx, a.tan_fovy, (const float3*)a.campos, acc, a.confidence, a.dL_dmeans2D, a.dL_dmeans3D,
            a.dL_dopacity, a.dL_dcolors, a.dL_dcov3D, a.dL_dsh, a.dL_dscales, a.dL_drotations, a.shs_rest,
            a.opacities, a.dL_dsh_rest);
    };
    const bool mom = gvd_bwd_moments();
    if (a.raw_params) mom ? go(gaussian_backward_kernel<MIN_CTAS, true, true>) : go(gaussian_backward_kernel<MIN_CTAS, false, true>);
    else mom ? go(gaussian_backward_kernel<MIN_CTAS, true, false>) : go(gaussian_backward_kernel<MIN_CTAS, false, false>);
}

void gvd_launch_gaussian_backward(const GvdRasterBackwardArgs& a, const RasterGeomPtrs& g, const float* acc,
                                  float focal_x, float focal_y, int num_visible, cudaStream_t s) {
    // GVD_GBWD_CTAS=3: 80 registers (spills) for three resident CTAs per SM instead of two (A/B timing knob)
    static int ctas = 0;
    if (!ctas) {
        const char* e = getenv("GVD_GBWD_CTAS");
        ctas = (e && e[0] == '3') ? 3 : 2;
    }
    if (ctas == 3) launch_gaussian_backward<3>(a, g, acc, focal_x, focal_y, num_visible, s);
    else launch_gaussian_backward<2>(a, g, acc, focal_x, focal_y, num_visible, s);
}
