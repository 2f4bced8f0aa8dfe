// raster_common.cuh -- shared definitions for the B200-native Gaussian rasterizer.
//
// Data layout in HBM (all caller-owned scratch, see include/gvd_raster.h):
//   geom   : kept until the backward. SplatRec[P] (64 B per Gaussian, written by preprocess, gathered on demand by the
//            render kernels -- the visible set is a few MB and lives in the 126 MB L2), clamped[P], tiles_touched[P],
//            vis_id[P] (ids of the V visible Gaussians, ascending), counts {V, R}.
//   sort   : forward-only, sized by P. depth keys by id, per-block visible / instance counts, the two (key, id)
//            ping-pong arrays of the depth sort and its digit histograms.
//   hist   : forward-only, sized by V. Per-chunk tile histogram matrix [ceil(V/GVD_BIN_CHUNK)][T], tile totals.
//   binning: uint32 point_list[R] -- Gaussian ids, tile-major, depth-sorted inside a tile (identical to
//            the reference's sorted value array); optional uint64 keys[R] for parity checks.
//   img    : uint2 ranges[T], uint32 n_contrib[H*W]
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define GVD_TILE_X 16
#define GVD_TILE_Y 16
#define GVD_BLOCK 256      // threads per render CTA = pixels per tile
#define GVD_BATCH 256      // tile-list entries staged per round
#define GVD_ACC_STRIDE 12  // floats per Gaussian in the backward accumulator
#define GVD_BIN_CHUNK 64   // Gaussians (in depth order) per binning chunk (= threads per binning CTA)
#define GVD_MAX_TILES 49152  // per-CTA shared-memory tile counters (4 B each) must fit in 227 KB

// 64-byte per-Gaussian record:
//   a = {mean2D.x, mean2D.y, conic.x, conic.y}
//   b = {conic.z, opacity, rgb.r, rgb.g}
//   c = {rgb.b, depth, hx, hy}   hx,hy = half extents (pixels, with slack) of the axis-aligned bound of the
//                                 region where alpha can reach 1/255; -1e30 = never, +1e30 = unbounded
//   d = {rect_min.x | rect_min.y<<16, rect_max.x | rect_max.y<<16, 0, 0}  (tile units, as uint bits)
// The render kernels read a,b,c (48 B); the binning kernels read d.
struct __align__(16) SplatRec {
    float4 a, b, c, d;
};

struct RasterGeomPtrs {
    SplatRec* splat;
    uint8_t* clamped;
    uint32_t* tiles_touched;   // per Gaussian id
    uint32_t* vis_id;          // [P] ids of the visible Gaussians in ascending id order; V valid entries
    uint32_t* counts;          // [8] {V = visible Gaussians, R = instances, ...}
};

// Depth sort = least-significant-digit radix sort (4 passes of 8 bits) over the V compacted (depth bits, id) pairs,
// V read from device memory. One kernel per pass: a CTA ranks its tile of GVD_SORT_TILE keys stably, publishes its
// digit counts as flagged words and sums those of its predecessors (groups of GVD_SORT_SUPER tiles, one running total
// per group), so there is no histogram kernel per pass and no per-key atomic.
#define GVD_PRE_BLOCK 256        // threads per preprocess CTA
#define GVD_COMPACT_BLOCK 256    // threads per compaction CTA (= one preprocess CTA)
#define GVD_SORT_TILE 1024       // keys per sort CTA
#define GVD_SORT_THREADS 256
#define GVD_GHIST_COPIES 16      // replicas of the global digit histogram (spreads the compaction kernel's atomics)
#define GVD_SORT_SUPER 32        // tiles per group (bounds the number of predecessor counts a tile has to fetch)
struct RasterSortPtrs {
    uint32_t* depth_key;       // [P] per Gaussian id: depth bits (undefined when culled)
    uint32_t* blk_vis;         // [nb] visible Gaussians per preprocess CTA
    uint32_t* blk_tiles;       // [nb] instances per preprocess CTA
    uint32_t* key[2];          // [P] ping-pong
    uint32_t* val[2];          // [P] ping-pong; val[0] holds the depth-sorted ids after the 4th pass
    uint32_t* zeroed;          // ghist[GVD_GHIST_COPIES][4][256] | agg[4][nt][256] | incl[4][ns][256] | ticket[4], cleared by preprocess
    size_t zeroed_words;
    uint32_t *ghist, *agg, *incl, *ticket;
    size_t nb, nt, ns;
};

struct RasterHistPtrs {
    uint32_t* hist;            // [rows][T] per-chunk tile counts, turned into exclusive prefixes in place
    uint32_t* tile_total;      // [T]
    size_t rows;               // chunk rows available (>= ceil(V / GVD_BIN_CHUNK) unless a speculative guess was too small)
};

// Binning = depth sort of the Gaussians (P 32-bit keys) followed by a rect-aware STABLE counting sort on
// the tile id: chunks of GVD_BIN_CHUNK depth-consecutive Gaussians histogram their tile rects, a column prefix over
// chunks gives every chunk its starting rank in every tile, and each chunk then writes its Gaussians'
// ids tile by tile in depth order. The result equals the reference's single stable radix sort on
// (tile<<32 | depth bits): same tile -> by depth -> ties by Gaussian id (rasterizer_impl.cu:70-111,304-309).
struct RasterBinPtrs {
    uint32_t* point_list;
    uint64_t* keys;            // optional (export_keys): (tile<<32 | depth bits) per sorted instance
};

struct RasterImgPtrs {
    uint2* ranges;
    uint32_t* n_contrib;
};

// ---- launchers implemented in the .cu files -------------------------------------------
struct GvdRasterForwardArgs;
struct GvdRasterBackwardArgs;

void gvd_launch_preprocess(const GvdRasterForwardArgs& a, const RasterGeomPtrs& g, const RasterSortPtrs& so, float focal_x,
                           float focal_y, dim3 grid, cudaStream_t s);
// compaction of the visible Gaussians (+ V, R to `counts` and, when given, to the pinned host words r_host[0] = R, r_host[1] = V)
void gvd_launch_compact(int P, const RasterGeomPtrs& g, const RasterSortPtrs& so, int* r_host, cudaStream_t s);
void gvd_launch_depth_sort(int P, const RasterGeomPtrs& g, const RasterSortPtrs& so, cudaStream_t s);
cudaError_t gvd_launch_bin_count(const RasterGeomPtrs& g, const RasterSortPtrs& so, const RasterHistPtrs& h, const RasterImgPtrs& im,
                                 dim3 grid, cudaStream_t s);
cudaError_t gvd_launch_bin_fill(const RasterGeomPtrs& g, const RasterSortPtrs& so, const RasterHistPtrs& h, const RasterBinPtrs& b,
                                const RasterImgPtrs& im, dim3 grid, uint32_t capacity, cudaStream_t s);
void gvd_launch_export_keys(uint32_t capacity, const RasterSortPtrs& so, const RasterBinPtrs& b, const RasterImgPtrs& im,
                            dim3 grid, cudaStream_t s);
void gvd_launch_render_forward(const GvdRasterForwardArgs& a, const RasterGeomPtrs& g, const RasterBinPtrs& b,
                               const RasterImgPtrs& im, dim3 grid, uint32_t capacity, cudaStream_t s);
// zero / zero_n4: a region (float4 units) the render CTAs clear on the side (the dense gradient outputs), or null
void gvd_launch_render_backward(const GvdRasterBackwardArgs& a, const RasterGeomPtrs& g, const RasterBinPtrs& b,
                                const RasterImgPtrs& im, float* acc, float4* zero, size_t zero_n4, dim3 grid, cudaStream_t s);
void gvd_launch_zero_fill(float* p, size_t floats, cudaStream_t s);
void gvd_launch_zero_bytes(void* p, size_t bytes, cudaStream_t s);  // any alignment / size
// num_visible < 0: unknown on the host, the grid covers P and reads V from g.counts
void gvd_launch_gaussian_backward(const GvdRasterBackwardArgs& a, const RasterGeomPtrs& g, const float* acc,
                                  float focal_x, float focal_y, int num_visible, cudaStream_t s);
void gvd_launch_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present,
                             cudaStream_t s);

// ---- small device helpers ----------------------------------------------------------------
#ifdef __CUDACC__

// Column-major 3x3 with the exact operation order of the math library the reference uses
// (glm::mat3: m[col][row]; operator* and transpose written out term by term), so that the
// compiler sees the same expression trees as in DGR/cuda_rasterizer/forward.cu:74-152.
struct M3 {
    float m[3][3];
};

__device__ __forceinline__ M3 m3_make(float x0, float y0, float z0, float x1, float y1, float z1, float x2,
                                      float y2, float z2) {
    M3 r;
    r.m[0][0] = x0; r.m[0][1] = y0; r.m[0][2] = z0;
    r.m[1][0] = x1; r.m[1][1] = y1; r.m[1][2] = z1;
    r.m[2][0] = x2; r.m[2][1] = y2; r.m[2][2] = z2;
    return r;
}

__device__ __forceinline__ M3 m3_mul(const M3& p, const M3& q) {
    M3 r;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int w = 0; w < 3; ++w)
            r.m[c][w] = p.m[0][w] * q.m[c][0] + p.m[1][w] * q.m[c][1] + p.m[2][w] * q.m[c][2];
    return r;
}

__device__ __forceinline__ M3 m3_transpose(const M3& p) {
    M3 r;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int w = 0; w < 3; ++w) r.m[c][w] = p.m[w][c];
    return r;
}

// auxiliary.h:58-77 (matrices arrive transposed; index as column-major)
__device__ __forceinline__ float3 xform_point_4x3(const float3& p, const float* __restrict__ m) {
    float3 t = {
        m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12],
        m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
        m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14],
    };
    return t;
}

__device__ __forceinline__ float4 xform_point_4x4(const float3& p, const float* __restrict__ m) {
    float4 t = {
        m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12],
        m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
        m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14],
        m[3] * p.x + m[7] * p.y + m[11] * p.z + m[15],
    };
    return t;
}

// auxiliary.h:41-44 -- evaluated in DOUBLE (the literals 1.0 / 0.5 promote), then narrowed.
__device__ __forceinline__ float ndc_to_pix(float v, int S) { return ((v + 1.0) * S - 1.0) * 0.5; }

// forward.cu:118-152. Quaternion used as given (not normalised). Sigma = (S R)^T (S R).
__device__ __forceinline__ void cov3d_from_scale_rot(const float3 scale, float mod, const float4 rot,
                                                     float* cov3D) {
    // S is diagonal; the reference multiplies the full matrices (zeros included) -- the extra
    // "+ 0*x" terms are exact, so M[c][r] = fl(s_r * R[c][r]).
    const float sx = mod * scale.x, sy = mod * scale.y, sz = mod * scale.z;
    const float r = rot.x, x = rot.y, y = rot.z, z = rot.w;
    M3 R = m3_make(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
                   2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
                   2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
    M3 M;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        M.m[c][0] = sx * R.m[c][0];
        M.m[c][1] = sy * R.m[c][1];
        M.m[c][2] = sz * R.m[c][2];
    }
    M3 Sigma = m3_mul(m3_transpose(M), M);
    cov3D[0] = Sigma.m[0][0];
    cov3D[1] = Sigma.m[0][1];
    cov3D[2] = Sigma.m[0][2];
    cov3D[3] = Sigma.m[1][1];
    cov3D[4] = Sigma.m[1][2];
    cov3D[5] = Sigma.m[2][2];
}

// SH basis constants, auxiliary.h:21-39
#define GVD_SH_C0 0.28209479177387814f
#define GVD_SH_C1 0.4886025119029199f
#define GVD_SH_C2_0 1.0925484305920792f
#define GVD_SH_C2_1 -1.0925484305920792f
#define GVD_SH_C2_2 0.31539156525252005f
#define GVD_SH_C2_3 -1.0925484305920792f
#define GVD_SH_C2_4 0.5462742152960396f
#define GVD_SH_C3_0 -0.5900435899266435f
#define GVD_SH_C3_1 2.890611442640554f
#define GVD_SH_C3_2 -0.4570457994644658f
#define GVD_SH_C3_3 0.3731763325901154f
#define GVD_SH_C3_4 -0.4570457994644658f
#define GVD_SH_C3_5 1.445305721320277f
#define GVD_SH_C3_6 -0.5900435899266435f

__device__ __forceinline__ float3 f3_add(float3 a, float3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ float3 f3_sub(float3 a, float3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ float3 f3_scale(float s, float3 a) { return {s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ float f3_dot(float3 a, float3 b) {
    float tx = a.x * b.x, ty = a.y * b.y, tz = a.z * b.z;
    return tx + ty + tz;
}

// Loads one Gaussian's SH coefficients into registers: 12 x LDG.128 when the row is 16-byte aligned
// (M == 16: 192 B per Gaussian), scalar otherwise. Only the (deg+1)^2 active coefficients are read.
__device__ __forceinline__ void load_sh(const float* __restrict__ shs, int idx, int deg, int max_coeffs, float (&v)[48]) {
    const float* row = shs + (size_t)idx * max_coeffs * 3;
    const int nfl = 3 * (deg + 1) * (deg + 1);
    if (max_coeffs == 16 && ((reinterpret_cast<uintptr_t>(row) & 15) == 0)) {
        const float4* r4 = reinterpret_cast<const float4*>(row);
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            if (4 * k < nfl) {
                const float4 t = __ldg(r4 + k);
                v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < 48; ++k)
            if (k < nfl) v[k] = row[k];
    }
}

// ---- GaussianModel activations (scene/gaussian_model.py:36-43), applied in-kernel when raw_params is set ----
__device__ __forceinline__ float3 act_scale(float3 s) { return {expf(s.x), expf(s.y), expf(s.z)}; }            // torch.exp
__device__ __forceinline__ float act_opacity(float o) { return 1.0f / (1.0f + expf(-o)); }                      // torch.sigmoid
__device__ __forceinline__ float4 act_rotation(float4 q) {  // torch.nn.functional.normalize: q / max(||q||_2, 1e-12)
    const float n = fmaxf(sqrtf((q.x * q.x + q.y * q.y) + (q.z * q.z + q.w * q.w)), 1e-12f);
    return {q.x / n, q.y / n, q.z / n, q.w / n};
}
// Gradient through the normalisation (auxiliary.h:119-131, float4 form): dv of q / ||q||
__device__ __forceinline__ float4 dnormvdv4(float4 v, float4 dv) {
    const float sum2 = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
    const float4 vdv = {v.x * dv.x, v.y * dv.y, v.z * dv.z, v.w * dv.w};
    const float vdv_sum = vdv.x + vdv.y + vdv.z + vdv.w;
    float4 r;
    r.x = ((sum2 - v.x * v.x) * dv.x - v.x * (vdv_sum - vdv.x)) * invsum32;
    r.y = ((sum2 - v.y * v.y) * dv.y - v.y * (vdv_sum - vdv.y)) * invsum32;
    r.z = ((sum2 - v.z * v.z) * dv.z - v.z * (vdv_sum - vdv.z)) * invsum32;
    r.w = ((sum2 - v.w * v.w) * dv.w - v.w * (vdv_sum - vdv.w)) * invsum32;
    return r;
}
// SH coefficients from the two parameter tensors (no [P,M,3] concatenation): coefficient 0 from _features_dc [P,1,3],
// coefficients 1.. from _features_rest [P,M-1,3] (rows of 180 bytes at M = 16: scalar loads)
__device__ __forceinline__ void load_sh_split(const float* __restrict__ dc, const float* __restrict__ rest, int idx, int deg,
                                              int max_coeffs, float (&v)[48]) {
    const int nfl = 3 * (deg + 1) * (deg + 1);
    v[0] = __ldg(dc + 3 * (size_t)idx);
    v[1] = __ldg(dc + 3 * (size_t)idx + 1);
    v[2] = __ldg(dc + 3 * (size_t)idx + 2);
    const float* row = rest + (size_t)idx * (max_coeffs - 1) * 3;
#pragma unroll
    for (int k = 3; k < 48; ++k)
        if (k < nfl) v[k] = __ldg(row + k - 3);
}

// ---- programmatic dependent launch (sm_90+) --------------------------------------------------------------------
// Every kernel of the forward/backward chain is launched with programmaticStreamSerialization: its CTAs may become
// resident and run their prologue while the previous kernel of the stream drains. pdl_wait() -- the first statement
// of every such kernel, executed by every thread -- returns once the previous grid has completed and its writes are
// visible, so the chain stays transitively ordered; pdl_trigger() lets the NEXT kernel's CTAs be scheduled as soon as
// all CTAs of this grid have started. Both are no-ops for a launch without the attribute.  Measured at C2: +1.4 %.
#ifndef GVD_HOST_EMU
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#else  // host execution of this source by tests/cuda_emu: launches are serialised, nothing to wait for
__device__ __forceinline__ void pdl_wait() {}
__device__ __forceinline__ void pdl_trigger() {}
#endif

int gvd_render_split();  // CTAs per tile in the render kernels
bool gvd_pdl_enabled();  // GVD_PDL=0 turns the launch attribute off (A/B timing knob)

template <typename... Exp, typename... Act>
inline cudaError_t gvd_launch(void (*kernel)(Exp...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Act&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = gvd_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<Exp>(args)...);
}

// ---- mbarrier / TMA bulk-copy wrappers (sm_90+; SASS: UBLKCP + SYNCS) ---------------------
#ifndef GVD_HOST_EMU
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk async copy, completion signalled on an mbarrier (bytes % 16 == 0).
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
#else  // GVD_HOST_EMU (tests/cuda_emu): the bulk copy is a synchronous memcpy that flips the barrier's phase bit itself
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t) { *bar = 0; }
__device__ __forceinline__ void mbar_fence_init() {}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t*, uint32_t) {}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (((__atomic_load_n(bar, __ATOMIC_ACQUIRE)) & 1u) == parity) emu_yield();  // phase `parity` completes when the bit flips
}
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    memcpy(smem_dst, gmem_src, bytes);
    __atomic_fetch_xor(bar, 1ull, __ATOMIC_RELEASE);
}
#endif

// ---- tile-list staging shared by the forward and backward render kernels --------------------------
// Ids of one batch of a tile's list are fetched with ONE TMA bulk copy (the list is contiguous); the copy
// starts at the 16-byte boundary below the first id, `lead` is the number of leading pad elements.
struct IdSlot {
    uint32_t v[GVD_BATCH + 8];
};

__device__ __forceinline__ uint32_t issue_id_copy(IdSlot* slot, uint64_t* bar, const uint32_t* first, int cnt) {
    const uintptr_t addr = reinterpret_cast<uintptr_t>(first);
    const uint32_t lead = (uint32_t)((addr & 15) >> 2);
    const uint32_t bytes = (((lead + (uint32_t)cnt) * 4u) + 15u) & ~15u;
    mbar_arrive_expect_tx(bar, bytes);
    tma_bulk_g2s(slot->v, reinterpret_cast<const void*>(addr & ~(uintptr_t)15), bytes, bar);
    return lead;
}
__device__ __forceinline__ uint32_t id_lead(const uint32_t* first) {
    return (uint32_t)((reinterpret_cast<uintptr_t>(first) & 15) >> 2);
}

// Can this Gaussian reach alpha >= 1/255 on some pixel of the 8x4 block whose top-left pixel is (sx, sy)?
// Conservative; written with negated comparisons so that NaNs count as "yes".
__device__ __forceinline__ bool subtile_hit(float x, float y, float hx, float hy, float sx, float sy) {
    return !(x - hx > sx + 7.0f) && !(x + hx < sx) && !(y - hy > sy + 3.0f) && !(y + hy < sy);
}

#endif  // __CUDACC__
