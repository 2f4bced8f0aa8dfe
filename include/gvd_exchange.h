/* gvd_exchange.h -- C ABI of the cross-GPU gradient sum of the view-parallel rasterizer (libgvd_raster.so).
 *
 * The reference trains one view per step on one GPU (train_baseline.py:73-83, train_guidedvd.py) and has no multi-GPU
 * path; SURVEY.md section 8(e) shards the rasterizer over independent views with ONE exchange step: the per-Gaussian
 * gradients the backward produced (rasterize_points.cu:158-167 lists them: dL_dmeans3D, dL_dsh, dL_dopacity,
 * dL_dscales, dL_drotations, dL_dmeans2D ... 62 floats per Gaussian) are summed over the ranks.  This header is that
 * exchange as a hand-written kernel over NVLink peer memory (one process per GPU):
 *
 *   - every rank owns one "exchange buffer" (device memory it allocated here and exported to its peers through CUDA
 *     IPC) that the backward writes its gradients into;
 *   - gvd_exchange_allreduce_sum() launches ONE kernel per rank: cross-GPU flag barrier -> rank r sums slice r of all
 *     peers' buffers with direct peer loads (fixed rank order, so all ranks end with bit-identical sums) and writes the
 *     result into every peer's buffer with direct peer stores -> cross-GPU flag barrier.  No staging copies, no
 *     library collective; traffic per rank = (world-1)/world of the buffer in each direction.
 *   - with a multicast (NVLS) mapping of the buffers the slice is summed INSIDE the NVSwitch (multimem.ld_reduce) and
 *     written to all replicas by one multicast store (multimem.st): one request per element and direction instead of
 *     world-1 peer loads and world-1 peer stores.
 *
 * Plain C types only; the handle bytes travel between processes by whatever means the host has (the Python host
 * uses one torch.distributed all_gather at set-up time).  Return value 0 = ok; message via gvd_last_error(). */
#ifndef GVD_EXCHANGE_H
#define GVD_EXCHANGE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
#ifndef GVD_API
#define GVD_API __attribute__((visibility("default")))
#endif

#define GVD_EXCHANGE_MAX_RANKS 8
#define GVD_EXCHANGE_HANDLE_BYTES 64   /* == sizeof(cudaIpcMemHandle_t) */
#define GVD_EXCHANGE_FLAG_BYTES 256    /* per-rank signal words, placed after the payload inside the allocation */

/* Allocates payload_bytes (rounded up to 16) + GVD_EXCHANGE_FLAG_BYTES of zeroed device memory on the current device
 * and writes its IPC handle.  Free with gvd_exchange_free (after every peer closed its mapping). */
GVD_API int gvd_exchange_alloc(size_t payload_bytes, void** dev_ptr, unsigned char handle[GVD_EXCHANGE_HANDLE_BYTES]);
GVD_API int gvd_exchange_free(void* dev_ptr);
/* Maps a peer's allocation into this process (enables peer access lazily); close before the owner frees it. */
GVD_API int gvd_exchange_open(const unsigned char handle[GVD_EXCHANGE_HANDLE_BYTES], void** peer_ptr);
GVD_API int gvd_exchange_close(void* peer_ptr);

typedef struct {
    int world;                               /* 2..GVD_EXCHANGE_MAX_RANKS */
    int rank;
    void* bufs[GVD_EXCHANGE_MAX_RANKS];      /* bufs[rank] = own allocation, bufs[q] = mapping of rank q's allocation */
    size_t payload_bytes;                    /* as given to gvd_exchange_alloc on every rank (same value everywhere) */
    size_t n_floats;                         /* leading floats of the payload to sum (multiple of 4) */
    uint32_t epoch;                          /* 1, 2, 3, ... : +1 on every call, the same value on every rank */
    /* Optional NVLS path: a multicast mapping of the SAME buffers (one multicast object all `world` allocations are
     * bound to, e.g. torch.distributed._symmetric_memory's multicast_ptr).  When non-NULL the slice sums are formed inside
     * the NVSwitch (multimem.ld_reduce) and written to every replica with one multicast store (multimem.st); bufs[]
     * then only carries the flag words.  NULL: direct peer loads / stores. */
    void* multicast;
} GvdExchangeArgs;

/* In place: after the kernel, the first n_floats of EVERY rank's buffer hold the sum over ranks.  Stream-ordered: the
 * kernels that produced the local gradients must precede it on `stream`, consumers follow it on `stream`.  All ranks
 * must make the call (like any collective). */
GVD_API int gvd_exchange_allreduce_sum(const GvdExchangeArgs* args, void* stream);

/* The kernel's cross-GPU flag waits are bounded (20 s): a peer that never makes the call cannot wedge this GPU.  A wait
 * that gave up leaves the payload unspecified and records its epoch; this reads it back (synchronising copy).
 * *timed_out_epoch == 0 = every call so far completed its barriers. */
GVD_API int gvd_exchange_status(const void* own_ptr, size_t payload_bytes, uint32_t* timed_out_epoch);

#ifdef __cplusplus
}
#endif
#endif
