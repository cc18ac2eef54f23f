/*
 * ndpb200.h -- C ABI of libndpb200.so, the B200 (sm_100a) implementation of
 * ndpolator's batched `ndpolate` hot path.
 *
 * This is the drop-in boundary.  It sits beside the reference's CPython
 * extension `cndpolator` (reference src/ndp_py.c) at the seam between
 * ndpolator/ndpolator.py and native code: every entry point below replaces one
 * `cndpolator.*` call (or the C function behind it), takes plain pointers and
 * sizes, and returns an int status.  INTEGRATION.md shows the binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - All arrays are C-contiguous.  float64 for coordinates/values, int32 for
 *     indices/flags (the reference's `int`), int64_t for element counts.
 *   - Every data pointer may be a HOST pointer (pageable or pinned) or a
 *     DEVICE pointer on the table's GPU; the library detects which
 *     (cudaPointerGetAttributes) and stages host buffers through the GPU in
 *     pipelined chunks.  Outputs are caller-allocated.
 *   - Enum values equal the reference's (src/ndpolator.h:28-63).
 *   - There is NO CPU fallback: with no usable CUDA device every call returns
 *     NDPB_ERR_CUDA and ndpb_last_error() says why.
 *   - Calls on one table are serialised internally; different tables may be
 *     used from different threads.
 */
#ifndef NDPB200_H
#define NDPB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NDPB_MAX_AXES 12

/* reference src/ndpolator.h:28-32 (ndp_extrapolation_method) */
enum { NDPB_METHOD_NONE = 0, NDPB_METHOD_NEAREST = 1, NDPB_METHOD_LINEAR = 2 };
/* reference src/ndpolator.h:44-47 (ndp_search_algorithm) */
enum { NDPB_SEARCH_KDTREE = 0, NDPB_SEARCH_LINEAR = 1 };
/* reference src/ndpolator.h:59-63 (ndp_vertex_flag; ON_VERTEX|OUT_OF_BOUNDS = 3 occurs) */
enum { NDPB_ON_GRID = 0, NDPB_ON_VERTEX = 1, NDPB_OUT_OF_BOUNDS = 2 };

/* status codes (the reference has only NDP_SUCCESS / NDP_INVALID_TYPE, src/ndp_types.h:48-51) */
enum {
    NDPB_OK = 0,
    NDPB_ERR_INVALID = 1,   /* bad argument (shape, enum, NULL) */
    NDPB_ERR_CUDA = 2,      /* CUDA runtime failure or no device */
    NDPB_ERR_NOMEM = 3,     /* device allocation failed */
    NDPB_ERR_STATE = 4      /* call not valid for this table (e.g. axes-only table) */
};

typedef struct ndpb_table ndpb_table;

/* Replaces ndp_table_new_from_python -> ndp_axes_new_from_data +
 * ndp_table_new_from_data (reference src/ndp_py.c:91-124, src/ndp_types.c:51-68,
 * :159-269): uploads axes and grid to GPU `device`, computes strides, the
 * defined-vertex mask and the fully-defined-hypercube mask on the GPU, and (when
 * it pays and fits) a cell-major copy of the grid for single-burst gathers.
 * `axes[a]` has `axlen[a]` ticks; the first `nbasic` axes are basic
 * (nbasic == 0 means all).  `grid` holds prod(axlen) * vdim doubles; grid ==
 * NULL creates an axes-only table on which only ndpb_find is valid (the
 * reference's cndpolator.find builds exactly that, src/ndp_py.c:157-165). */
int ndpb_table_create(const double *const *axes, const int32_t *axlen, int naxes, int nbasic,
                      const double *grid, int vdim, int device, ndpb_table **out);

/* Replaces ndp_table_free via py_table_capsule_destructor (src/ndp_py.c:335-340). */
int ndpb_table_destroy(ndpb_table *t);

/* Replaces cndpolator.find -> ndp_query_pts_import (src/ndp_py.c:131-193,
 * src/ndpolator.c:354-414).  query_pts: n x naxes.  Outputs n x naxes each. */
int ndpb_find(ndpb_table *t, const double *query_pts, int64_t n,
              int32_t *indices, int32_t *flags, double *normed);

/* Replaces cndpolator.hypercubes -> ndp_find_hypercubes (src/ndp_py.c:254-318,
 * src/ndpolator.c:416-508).  Inputs are what ndpb_find returned.  `v` has a
 * fixed slot of 2^naxes * vdim doubles per query, of which the first
 * 2^dim[i] * vdim are written (corner-major, first free axis = most significant
 * bit, as in the reference); fdhc is the reference's ndp_hypercube.fdhc, which
 * its Python API does not expose. */
int ndpb_hypercubes(ndpb_table *t, const double *normed, const int32_t *indices, const int32_t *flags,
                    int64_t n, int32_t *dim, int32_t *fdhc, double *v);

/* Parity tap for ndp_find_nearest (src/ndpolator.c:188-352), C-only in the
 * reference.  Takes raw query points (imports them first), searches for EVERY
 * point, returns coords (n x naxes: nearest vertex, or superior corner of the
 * nearest fully defined hypercube) and the squared index-space distance. */
int ndpb_find_nearest(ndpb_table *t, const double *query_pts, int64_t n, int method, int search,
                      int32_t *coords, double *dist);

/* Replaces cndpolator.ndpolate -> ndp_query_pts_import + ndpolate
 * (src/ndp_py.c:343-409, src/ndpolator.c:510-671).  interps: n x vdim;
 * dists: n (0 for points interpolated inside a fully defined hypercube). */
int ndpb_ndpolate(ndpb_table *t, const double *query_pts, int64_t n, int method, int search,
                  double *interps, double *dists);

/* The hypercube-caching path (reference README "Hypercube caching", ndpolator/ndpolator.py:110-187): ndpolate
 * from the import products of a batch (what ndpb_find returned; host or DEVICE pointers -- kept on the device
 * they never round-trip through the host), e.g. on several tables that share their axes.  Same results as
 * ndpb_ndpolate on the original query points.  The reference has no such entry point: there the caller
 * interpolates the cached hypercubes itself.  n < 2^31. */
int ndpb_ndpolate_cached(ndpb_table *t, const double *normed, const int32_t *indices, const int32_t *flags,
                         int64_t n, int method, int search, double *interps, double *dists);

/* Replaces cndpolator.distance (src/ndp_py.c:196-251): LINEAR method, KDTREE
 * search, for every point. */
int ndpb_distance(ndpb_table *t, const double *query_pts, int64_t n, double *dists);

/* Register-time products, for parity tests of src/ndp_types.c:174-258.
 * vmask/hcmask: nverts int32 each (host or device). */
int64_t ndpb_table_nverts(const ndpb_table *t);
int ndpb_table_masks(ndpb_table *t, int32_t *vmask, int32_t *hcmask);

/* Shape of the reference's insertion-ordered kd-tree (src/ndpolator.c:114-136 +
 * external/kdtree/kdtree.c:167-209) as rebuilt on the GPU for exact tie
 * resolution: parent vertex id (-1 root, -2 not in tree) and depth (-1 not in
 * tree) per basic vertex.  method is NDPB_METHOD_NEAREST or _LINEAR. */
int ndpb_table_tree_topology(ndpb_table *t, int method, int32_t *parent, int32_t *depth);

/* Stream ordering for DEVICE-resident inputs.  The table launches its kernels on its own non-blocking
 * stream, which is not ordered against the stream that produced a device query buffer (e.g. torch's
 * current stream).  Call this with the producer's cudaStream_t (as integer; 0 = the legacy default stream)
 * before a call that takes device pointers: the table's stream then waits for everything enqueued on the
 * producer stream so far.  Every entry point returns only after its results are complete, so no ordering
 * is needed on the output side.  Host pointers need none of this.  (The reference is synchronous host
 * code, src/ndp_py.c:343-409; this is the price of accepting device pointers.) */
int ndpb_table_wait_stream(ndpb_table *t, uint64_t producer_stream);

/* Stream-ordered calls.  By default every entry point returns when its results are complete -- the reference's
 * semantics (synchronous host code, src/ndp_py.c:343-409).  With option "async" = 1, ndpb_ndpolate /
 * ndpb_find_nearest / ndpb_distance calls whose buffers are ALL device-resident (and hold at most 2^30 queries)
 * only enqueue their kernels on the table's compute stream (option "stream": the caller's own stream) and
 * return; results are complete in stream order, like any CUDA library call.  The one host decision of the
 * synchronous path (building the reference-shaped kd topology the first time an exact distance tie is met) is
 * taken ahead of time: the first async call of a kd-search extrapolating method builds that topology
 * (8 bytes per basic vertex and method).  ndpb_table_sync waits for the compute stream and folds the counters
 * of the last call into ndpb_table_stat.  Calls with host buffers are unaffected.  Results are identical. */
int ndpb_table_sync(ndpb_table *t);

/* Tuning knobs that never change results.  Keys: "gather" = 0 auto | 1 strided
 * (original layout) | 2 cell-major; "linear_search" = 0 auto | 1 brute force |
 * 2 lattice; "chunk" = queries per staged chunk for host buffers; "block_order" = 0 auto | K: blocked copy of a scalar table with 2^K corners per
 * contiguous block (K = naxes is the cell-major copy; smaller K costs 2^K times the grid instead of 2^naxes);
 * "fused" = 0 auto | 1 never use the
 * fused kernel (import + mask/nearest-site-map plan + one cell gather per query); "staged" = 0 auto: one-stage
 * cp.async gather, 4 CTAs/SM | 1 never use the staged gather kernels | 2 two-stage ring | 3 one stage,
 * 5 CTAs/SM (only used when the fused kernel is off or does not apply);
 * "nsm_subcells" = 1 | 2 | 4 | 8 | 16 (default; the largest that keeps a map under 2^21 regions is used): sub-cells per unit cell of the nearest-site maps (finer maps hold
 * fewer candidates per region and are larger; rebuilt lazily); "async" = 0 | 1, see ndpb_table_sync;
 * "search_fast_blocks_per_sm" / "search_hard_blocks_per_sm" =
 * grid-stride CTAs per SM of the two search passes (default 64 / 256); "stream" = a
 * cudaStream_t (as integer) to launch the kernels on instead of the table's own
 * stream, 0 to restore it. */
int ndpb_table_set_option(ndpb_table *t, const char *key, int64_t value);

/* Per-table counters of the last call, for bench.py / tests:
 * "launches" kernels launched, "offgrid" queries sent to the search,
 * "hard" queries that needed the lattice descent, "ties" queries resolved
 * through the kd topology, "slow" queries the fused kernel handed to the exact generic
 * kernel, "fused" batches that went through the fused kernel, "h2d_bytes", "d2h_bytes",
 * "cellmajor" (0/1), "block_order" (order of the blocked copy in use, 0 = none), "use_hc" (0/1: the hypercube-mask bit decides "fully defined"),
 * "main_ns" / "search_ns" / "finish_ns" device time (CUDA events) of the gather kernel
 * (fused or staged), of the search / slow-list kernels and of the extrapolation finish,
 * "nsm_regions_nearest" / "nsm_regions_linear" regions of the nearest-site maps built so
 * far, "variant_axes_*", "tree_levels_*", "mask_count_nearest" / "mask_count_linear" set vertices of the vertex /
 * hypercube mask, "sparse_search_*" (0/1: that mask is searched by scoring its few set vertices, k_search_sparse).
 * Unknown keys return -1. */
int64_t ndpb_table_stat(const ndpb_table *t, const char *key);

/* ---- one call, several GPUs -------------------------------------------------------------------
 * The reference's ndpolate is ONE call in one process (ndpolator/ndpolator.py:253-349).  A group
 * holds one replicated table per listed device (grid, masks, cell-major copy and
 * search structures replicated at create time, one host thread per device) and serves a HOST
 * batch by contiguous query ranges [g N / G, (g + 1) N / G), one host thread + stream set per
 * device; each device copies its results into its own slice of the caller's output arrays, which
 * is the "final result gather" -- there is no collective on the path.  Results are identical to
 * a single table's.  Device buffers are rejected (they belong to one GPU: use ndpb_group_table). */
typedef struct ndpb_group ndpb_group;
int ndpb_group_create(const double *const *axes, const int32_t *axlen, int naxes, int nbasic,
                      const double *grid, int vdim, const int32_t *devices, int ndev, ndpb_group **out);
int ndpb_group_destroy(ndpb_group *g);
int ndpb_group_size(const ndpb_group *g);
ndpb_table *ndpb_group_table(ndpb_group *g, int i);               /* borrowed; owned by the group */
int ndpb_group_ndpolate(ndpb_group *g, const double *query_pts, int64_t n, int method, int search,
                        double *interps, double *dists);
int ndpb_group_distance(ndpb_group *g, const double *query_pts, int64_t n, double *dists);
/* table options applied to every member; plus "min_queries_per_device" (default 65536: smaller batches
 * use fewer devices) */
int ndpb_group_set_option(ndpb_group *g, const char *key, int64_t value);
/* sum of a table stat over the members; "used_devices" = devices the last call ran on */
int64_t ndpb_group_stat(const ndpb_group *g, const char *key);

const char *ndpb_last_error(void);
const char *ndpb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* NDPB200_H */
