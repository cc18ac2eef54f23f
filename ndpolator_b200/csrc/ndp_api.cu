// ndp_api.cu -- the C ABI of libndpb200.so (include/ndpb200.h): device table
// object, host<->device staging pipeline and kernel launches.  No CPU fallback:
// every entry point fails with NDPB_ERR_CUDA when there is no usable device.
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>

#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "../../include/ndpb200.h"
#include "ndp_core.cuh"
#include "ndp_launch.h"
#include "ndp_table.cuh"

static_assert(NDPB_MAX_AXES == NDP_MAX_AXES, "header and core disagree on the axis limit");

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
static thread_local std::string g_last_error;

static int fail(int code, const std::string &msg)
{
    g_last_error = msg;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            int code_ = (e_ == cudaErrorMemoryAllocation) ? NDPB_ERR_NOMEM : NDPB_ERR_CUDA;        \
            return fail(code_, std::string(#call) + ": " + cudaGetErrorString(e_));                \
        }                                                                                          \
    } while (0)

extern "C" const char *ndpb_last_error(void) { return g_last_error.c_str(); }
extern "C" const char *ndpb_version(void) { return "ndpb200 0.1.0 (sm_100a; parity target ndpolator 1.3.0)"; }

// device temporaries that are released on every exit path
struct DevTemps {
    std::vector<void *> ptrs;
    ~DevTemps() { for (void *p : ptrs) cudaFree(p); }
    template <class T>
    cudaError_t alloc(T **out, size_t count)
    {
        void *p = nullptr;
        cudaError_t e = cudaMalloc(&p, sizeof(T) * (count ? count : 1));
        if (e == cudaSuccess) ptrs.push_back(p);
        *out = (T *) p;
        return e;
    }
    void release(void *p)      // hand ownership to the caller
    {
        for (auto &q : ptrs) if (q == p) q = nullptr;
    }
    void free_now(void *p)
    {
        for (auto &q : ptrs) if (q == p) { cudaFree(p); q = nullptr; }
    }
};

// ---------------------------------------------------------------------------
// table
// ---------------------------------------------------------------------------
struct Pyramid {
    NdpPyramid dev{};                       // host copy of the descriptor (device pointers inside)
    NdpPyramid *d_desc = nullptr;           // the same descriptor in device memory (kernels read it there)
    std::vector<uint32_t *> owned;
    short *prevv = nullptr, *nextv = nullptr;
    int count = 0;                          // set bits at level 0
};

struct Topology {
    int *parent = nullptr, *depth = nullptr;
    bool built = false;
    int levels = 0;
};

struct Slot {                               // one in-flight batch
    double *q = nullptr, *interps = nullptr, *dists = nullptr;     // staging (host callers only)
    int *coords = nullptr;
    int *list = nullptr, *chosen = nullptr, *tielist = nullptr, *hardlist = nullptr;
    int *rec_idx = nullptr; double *rec_nrm = nullptr;             // import products of the off-grid queries
    long long cap_q = 0, cap_work = 0, cap_coords = 0, cap_lists = 0;
    NdpCounters *ctr = nullptr;             // device
    NdpCounters *ctr_host = nullptr;        // pinned
    cudaEvent_t in_done = nullptr, searched = nullptr, out_ready = nullptr, out_done = nullptr;
    cudaEvent_t t0 = nullptr, t1 = nullptr, t2 = nullptr, t3 = nullptr;   // phase timing
    bool used = false;
    bool fused = false;                     // this batch went through k_fused (its tie list holds query ids)
};

struct ndpb_table {
    int device = 0;
    NdpGeom g{};
    int nticks = 0;
    bool has_grid = false;
    double *ticks = nullptr, *grid = nullptr;
    double *cells = nullptr;                // full cell-major copy (block order == naxes), what the staged / direct kernels read
    double *blocks = nullptr;               // blocked copy of a scalar table (ndp_fused.cuh: NdpBlocks), what k_fused reads
    int block_order = 0;
    long long grid_doubles = 0;
    Pyramid pyr[2];                         // [0] vmask, [1] hcmask: over the whole basic lattice
    Pyramid red[2];                         // the same masks over their variant axes only
    Topology topo[2];
    int *sparse_ids[2] = {nullptr, nullptr};   // ascending ids of a SPARSE mask's set vertices (lazy; see sparse_ready)
    bool sparse_tried[2] = {false, false};
    // fused path (ndp_fused_kernels.cuh): nearest-site maps per geometry (lazy), hypercube-mask slice
    NdpSiteMap nsm[3]{};
    bool nsm_tried[3] = {false, false, false};
    int *nsm_offs[3] = {nullptr, nullptr, nullptr};
    uint32_t *nsm_ents[3] = {nullptr, nullptr, nullptr};
    uint32_t *nsm_head[3] = {nullptr, nullptr, nullptr};
    NdpHcSlice hcs{};
    bool use_hc = false;                    // hypercube-mask bit <=> hypercube fully defined
    cudaStream_t s_compute = nullptr, s_in = nullptr, s_out = nullptr;
    cudaStream_t s_owned = nullptr;         // the compute stream this table created (s_compute may be a caller's)
    Slot slot[2];
    std::mutex mu;
    int sm_count = 148;
    // options
    int opt_gather = 0, opt_linear_search = 0, opt_staged = 0, opt_fused = 0, opt_block_order = 0, opt_nsm_subcells = 16;
    int opt_async = 0;                      // 1: device-resident calls only enqueue (ndpb_table_sync completes them)
    bool async_pending = false;             // an enqueued call whose counters have not been folded into the stats yet
    long long opt_chunk = 1 << 20;          // profiles/r01p_sweep_chunk.txt: 1 Mi beats 2, 4 and 8 Mi end to end
    int opt_fast_blocks = 64, opt_hard_blocks = 256;   // grid-stride CTAs per SM of the two search passes (profiles/r01p_sweep_search.txt)
    // stats of the last call
    long long st_launches = 0, st_offgrid = 0, st_hard = 0, st_ties = 0, st_h2d = 0, st_d2h = 0, st_slow = 0, st_fused = 0;
    double st_main_ms = 0, st_search_ms = 0, st_finish_ms = 0;
};

static int blocks_for(long long items, int per_block = NDP_BLOCK)
{
    long long b = (items + per_block - 1) / per_block;
    return (int) std::max<long long>(1, std::min<long long>(b, INT_MAX));
}

static bool is_device_ptr(const void *p)
{
    if (!p) return false;
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

static size_t tick_smem(const ndpb_table *t) { return (size_t) t->nticks * sizeof(double); }
static int ticks_in_smem(const ndpb_table *t) { return tick_smem(t) <= 40 * 1024; }
static size_t smem_bytes(const ndpb_table *t) { return ticks_in_smem(t) ? tick_smem(t) : 0; }

// Builds the occupancy pyramid above `level0` (nvar axes of sizes dim0[], mapped to the basic
// axes var_axis[]); takes ownership of level0.
static int build_pyramid(ndpb_table *t, Pyramid &P, uint32_t *level0, int nvar, const int *var_axis, const int *dim0, bool full)
{
    const NdpGeom &g = t->g;
    NdpPyramid &D = P.dev;
    D.nlev = 0;
    D.nvar = nvar;
    int dim[NDP_MAX_AXES];
    for (int j = 0; j < nvar; j++) { dim[j] = dim0[j]; D.var_axis[j] = var_axis[j]; }
    P.owned.push_back(level0);
    uint32_t *prev = level0;
    for (int lev = 0; lev < NDP_MAX_LEVELS; lev++) {
        long long nblocks = 1;
        for (int j = nvar - 1, s = 1; j >= 0; j--) { D.dim[lev][j] = dim[j]; D.stride[lev][j] = s; s *= dim[j]; nblocks *= dim[j]; }
        if (lev == 0) D.bits[0] = level0;
        else {
            uint32_t *bits = nullptr;
            CU(cudaMalloc(&bits, sizeof(uint32_t) * ((nblocks + 31) / 32)));
            P.owned.push_back(bits);
            PyrLevelArgs A{};
            A.nb = nvar;
            for (int j = 0; j < nvar; j++) { A.cdim[j] = D.dim[lev - 1][j]; A.cstride[j] = D.stride[lev - 1][j]; A.pdim[j] = dim[j]; }
            A.nparent = (int) nblocks; A.cbits = prev; A.pbits = bits;
            k_pyr_level<<<blocks_for(((nblocks + 31) / 32) * 32), NDP_BLOCK, 0, t->s_compute>>>(A);
            CU(cudaGetLastError());
            D.bits[lev] = bits;
            prev = bits;
        }
        D.nlev = lev + 1;
        if (nblocks == 1) break;
        for (int j = 0; j < nvar; j++) dim[j] = (dim[j] + 1) / 2;
    }
    // any_set: the single top block; first_clear: lowest clear bit of level 0 (full pyramid only)
    uint32_t top = 0;
    int first = INT_MAX;
    if (full) {
        int *d_first = nullptr, both[2] = {INT_MAX, 0};     // [0] first clear bit, [1] number of set bits
        CU(cudaMalloc(&d_first, 2 * sizeof(int)));
        CU(cudaMemcpyAsync(d_first, both, sizeof(both), cudaMemcpyHostToDevice, t->s_compute));
        k_first_clear<<<blocks_for((g.nverts + 31) / 32), NDP_BLOCK, 0, t->s_compute>>>(level0, g.nverts, d_first);
        k_count_bits<<<blocks_for((g.nverts + 31) / 32), NDP_BLOCK, 0, t->s_compute>>>(level0, g.nverts, d_first + 1);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(both, d_first, sizeof(both), cudaMemcpyDeviceToHost, t->s_compute));
        CU(cudaStreamSynchronize(t->s_compute));
        CU(cudaFree(d_first));
        first = both[0];
        P.count = both[1];
    }
    CU(cudaMemcpyAsync(&top, D.bits[D.nlev - 1], sizeof(uint32_t), cudaMemcpyDeviceToHost, t->s_compute));
    CU(cudaStreamSynchronize(t->s_compute));
    D.first_clear = (first == INT_MAX) ? -1 : first;
    D.any_set = (top & 1u) ? 1 : 0;
    D.prevv = D.nextv = nullptr;
    if (nvar >= 1 && dim0[nvar - 1] <= 32767) {       // row tables along the last pyramid axis
        long long npts = 1;
        for (int j = 0; j < nvar; j++) npts *= dim0[j];
        const long long nrows = npts / dim0[nvar - 1];
        CU(cudaMalloc(&P.prevv, sizeof(short) * npts));
        CU(cudaMalloc(&P.nextv, sizeof(short) * npts));
        k_row_tables<<<blocks_for(nrows), NDP_BLOCK, 0, t->s_compute>>>(level0, nrows, dim0[nvar - 1], P.prevv, P.nextv);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(t->s_compute));
        D.prevv = P.prevv; D.nextv = P.nextv;
    }
    CU(cudaMalloc(&P.d_desc, sizeof(NdpPyramid)));
    // (copies the kernels of the table's non-blocking stream depend on are issued on that stream and waited for:
    // a plain cudaMemcpy is ordered on the legacy default stream only and may return before the DMA has finished)
    CU(cudaMemcpyAsync(P.d_desc, &D, sizeof(NdpPyramid), cudaMemcpyHostToDevice, t->s_compute));
    CU(cudaStreamSynchronize(t->s_compute));
    return NDPB_OK;
}

// Finds the basic axes the mask actually depends on and builds the pyramid of the slice they span.
static int build_reduced(ndpb_table *t, int m)
{
    const NdpGeom &g = t->g;
    const uint32_t *bits = t->pyr[m].dev.bits[0];
    const int ref = m;                          // vertex mask: coordinate 0; hypercube mask: coordinate 1
    int *d_diff = nullptr, diff[NDP_MAX_AXES] = {0};
    {
        DevTemps hold;
        CU(hold.alloc(&d_diff, (size_t) NDP_MAX_AXES));
        CU(cudaMemsetAsync(d_diff, 0, sizeof(diff), t->s_compute));
        k_mask_variance<<<blocks_for(g.nverts), NDP_BLOCK, 0, t->s_compute>>>(g, bits, ref, d_diff);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(diff, d_diff, sizeof(diff), cudaMemcpyDeviceToHost, t->s_compute));
        CU(cudaStreamSynchronize(t->s_compute));
    }
    SliceArgs A{};
    A.g = g; A.ref = ref; A.bits = bits; A.nvar = 0;
    long long nslice = 1;
    for (int a = 0; a < g.nbasic; a++)
        if (diff[a]) { A.var_axis[A.nvar] = a; A.dim[A.nvar] = g.len[a]; nslice *= g.len[a]; A.nvar++; }
    A.nslice = (int) nslice;
    uint32_t *level0 = nullptr;
    CU(cudaMalloc(&level0, sizeof(uint32_t) * ((nslice + 31) / 32)));
    A.out = level0;
    k_mask_slice<<<blocks_for(((nslice + 31) / 32) * 32), NDP_BLOCK, 0, t->s_compute>>>(A);
    if (cudaError_t e = cudaGetLastError(); e != cudaSuccess) { cudaFree(level0); CU(e); }
    return build_pyramid(t, t->red[m], level0, A.nvar, A.var_axis, A.dim, false);      // takes ownership of level0
}

static void free_slot(Slot &s)
{
    cudaFree(s.q); cudaFree(s.interps); cudaFree(s.dists); cudaFree(s.coords);
    cudaFree(s.list); cudaFree(s.chosen); cudaFree(s.tielist); cudaFree(s.hardlist); cudaFree(s.ctr);
    cudaFree(s.rec_idx); cudaFree(s.rec_nrm);
    if (s.ctr_host) cudaFreeHost(s.ctr_host);
    cudaEvent_t *ev[] = {&s.in_done, &s.searched, &s.out_ready, &s.out_done, &s.t0, &s.t1, &s.t2, &s.t3};
    for (auto e : ev) if (*e) cudaEventDestroy(*e);
    s = Slot{};
}

extern "C" int ndpb_table_destroy(ndpb_table *t)
{
    if (!t) return NDPB_OK;
    cudaSetDevice(t->device);
    cudaDeviceSynchronize();
    cudaFree(t->ticks); cudaFree(t->grid);
    if (t->blocks) cudaFree(t->blocks); else cudaFree(t->cells);      // cells aliases blocks when the order is naxes
    for (Pyramid *set : {t->pyr, t->red})
        for (int m = 0; m < 2; m++) {
            for (auto p : set[m].owned) cudaFree(p);
            cudaFree(set[m].prevv); cudaFree(set[m].nextv);
            cudaFree(set[m].d_desc);
        }
    for (auto &T : t->topo) { cudaFree(T.parent); cudaFree(T.depth); }
    for (int m = 0; m < 2; m++) cudaFree(t->sparse_ids[m]);
    for (int geo = 0; geo < 3; geo++) { cudaFree(t->nsm_offs[geo]); cudaFree(t->nsm_ents[geo]); cudaFree(t->nsm_head[geo]); }
    for (auto &s : t->slot) free_slot(s);
    if (t->s_owned) cudaStreamDestroy(t->s_owned);
    if (t->s_in) cudaStreamDestroy(t->s_in);
    if (t->s_out) cudaStreamDestroy(t->s_out);
    delete t;
    return NDPB_OK;
}

static void drop_cells(ndpb_table *t)
{
    if (t->blocks) cudaFree(t->blocks); else cudaFree(t->cells);
    t->blocks = t->cells = nullptr;
    t->block_order = 0;
}

// The gather-friendly copy of the grid.  Scalar tables: the blocked layout of ndp_fused.cuh with the largest block
// order whose copy (2^order times the grid) fits in 60 % of the free device memory -- order == naxes is the
// cell-major copy (one contiguous (8 << n)-byte burst per query), smaller orders cost 2^(n - order) bursts of
// (8 << order) bytes per query and serve tables whose full copy would not fit.  Small-vector tables (vdim * 8 <
// 128): the full cell-major copy when it fits.  Vectors of a full line and more are gathered from the grid itself.
static int maybe_build_cells(ndpb_table *t)
{
    const NdpGeom &g = t->g;
    if (t->cells || t->blocks || t->opt_gather == 1 || g.naxes > 6) return NDPB_OK;
    if (t->opt_gather != 2 && g.vdim * sizeof(double) >= 128) return NDPB_OK;
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    if (g.vdim != 1) {
        const long long bytes = (g.ncells << g.naxes) * g.vdim * (long long) sizeof(double);
        if (bytes > (long long) (free_b / 2)) return NDPB_OK;
        CU(cudaMalloc(&t->cells, bytes));
        k_build_cells<<<t->sm_count * 16, NDP_BLOCK, 0, t->s_compute>>>(g, t->grid, t->cells);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(t->s_compute));
        return NDPB_OK;
    }
    NdpBlocks B{};
    int order = 0;
    for (int k = g.naxes; k >= 1; k--) {
        if (t->opt_block_order && k != t->opt_block_order) continue;
        ndp_blocks_setup(g, k, B);
        if (((B.nblocks << k) * (long long) sizeof(double)) <= (long long) (free_b / 10 * 6)) { order = k; break; }
    }
    if (!order) return NDPB_OK;
    CU(cudaMalloc(&t->blocks, (size_t) (B.nblocks << order) * sizeof(double)));
    BlockBuildArgs A{};
    A.g = g; A.order = order; A.nblocks = B.nblocks; A.grid = t->grid; A.blocks = t->blocks;
    for (int a = 0; a < g.naxes; a++) A.stride[a] = B.stride[a];
    k_build_blocks<<<t->sm_count * 16, NDP_BLOCK, 0, t->s_compute>>>(A);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(t->s_compute));
    t->block_order = order;
    if (order == g.naxes) t->cells = t->blocks;
    return NDPB_OK;
}

// Register-time products of the fused path that do not depend on the search mode: the hypercube-mask
// slice, and whether its bit decides "fully defined" -- always when every axis is basic; with associated
// axes only if component 0 of a basic vertex is NaN at every associated index or at none
// (ndp_types.c:179-184 looks at associated index 0 only, ndpolator.c:491 at the corner actually used).
static int setup_fused(ndpb_table *t)
{
    const NdpGeom &g = t->g;
    const NdpPyramid &R = t->red[1].dev;
    t->hcs = NdpHcSlice{};
    long long nslice = 1;
    for (int j = 0; j < R.nvar; j++) { t->hcs.astride[R.var_axis[j]] = R.stride[0][j]; nslice *= g.len[R.var_axis[j]]; }
    t->hcs.bits = R.bits[0];
    t->hcs.nwords = (int) ((nslice + 31) / 32);
    t->use_hc = true;
    if (g.nbasic < g.naxes) {
        int *d_flag = nullptr, flag = 1;
        CU(cudaMalloc(&d_flag, sizeof(int)));
        CU(cudaMemcpyAsync(d_flag, &flag, sizeof(int), cudaMemcpyHostToDevice, t->s_compute));
        cudaError_t e = ndp_launch_assoc_consistent(t->grid, g.nverts, g.vstride[g.nbasic - 1], g.vdim, d_flag, t->sm_count * 8, t->s_compute);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, t->s_compute);
        if (e == cudaSuccess) e = cudaStreamSynchronize(t->s_compute);
        cudaFree(d_flag);
        CU(e);
        t->use_hc = flag != 0;
    }
    return NDPB_OK;
}

// Lazily builds the nearest-site map of one geometry (k_nsm_count -> exclusive scan -> k_nsm_fill).
static int ensure_nsm(ndpb_table *t, int geo)
{
    if (t->nsm_tried[geo]) return NDPB_OK;
    t->nsm_tried[geo] = true;
    NdpSiteMap M;
    nsm_setup(t->g, t->red[geo == NSM_GEO_VERTEX ? 0 : 1].dev, geo, M, t->opt_nsm_subcells);
    if (!M.valid) return NDPB_OK;
    cudaStream_t st = t->s_owned;
    int *counts = nullptr, *offs = nullptr;
    uint32_t *ents = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    int total = 0;
    cudaError_t e = cudaMalloc(&counts, sizeof(int) * ((size_t) M.nregions + 1));
    if (e == cudaSuccess) e = cudaMalloc(&offs, sizeof(int) * ((size_t) M.nregions + 1));
    if (e == cudaSuccess) e = cudaMemsetAsync(counts, 0, sizeof(int) * ((size_t) M.nregions + 1), st);
    if (e == cudaSuccess) e = ndp_launch_nsm_count(M, counts, st);
    if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, counts, offs, M.nregions + 1, st);
    if (e == cudaSuccess) e = cudaMalloc(&tmp, tmp_bytes);
    if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, counts, offs, M.nregions + 1, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&total, offs + M.nregions, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    uint32_t *head = nullptr;
    const size_t head_bytes = sizeof(uint32_t) * (size_t) M.nregions * nsm_head_words(M.ew);
    if (e == cudaSuccess) e = cudaMalloc(&ents, sizeof(uint32_t) * ((size_t) std::max(total, 1) * M.ew));
    if (e == cudaSuccess) e = cudaMalloc(&head, head_bytes);
    if (e == cudaSuccess) e = cudaMemsetAsync(head, 0, head_bytes, st);
    if (e == cudaSuccess) { M.offs = offs; e = ndp_launch_nsm_fill(M, ents, head, st); }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(counts); cudaFree(tmp);
    if (e != cudaSuccess) { cudaFree(offs); cudaFree(ents); cudaFree(head); CU(e); }
    M.ents = ents; M.head = head;
    t->nsm[geo] = M; t->nsm_offs[geo] = offs; t->nsm_ents[geo] = ents; t->nsm_head[geo] = head;
    return NDPB_OK;
}

extern "C" int ndpb_table_create(const double *const *axes, const int32_t *axlen, int naxes, int nbasic,
                                 const double *grid, int vdim, int device, ndpb_table **out)
{
    if (!out) return fail(NDPB_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!axes || !axlen || naxes < 1 || naxes > NDP_MAX_AXES) return fail(NDPB_ERR_INVALID, "naxes must be in [1, 12]");
    if (nbasic == 0) nbasic = naxes;
    if (nbasic < 1 || nbasic > naxes) return fail(NDPB_ERR_INVALID, "nbasic must be in [1, naxes]");
    if (vdim < 1) return fail(NDPB_ERR_INVALID, "vdim must be >= 1");
    long long nverts = 1, nall = 1, ncells = 1;
    int nticks = 0;
    for (int a = 0; a < naxes; a++) {
        if (axlen[a] < 2 || axlen[a] > (1 << 16)) return fail(NDPB_ERR_INVALID, "every axis needs 2..65536 ticks");
        if (!axes[a]) return fail(NDPB_ERR_INVALID, "axis pointer is NULL");
        nall *= axlen[a]; ncells *= axlen[a] - 1; nticks += axlen[a];
        if (a < nbasic) nverts *= axlen[a];
        if (nall > (1LL << 40)) return fail(NDPB_ERR_INVALID, "grid too large");
    }
    if (nverts >= INT_MAX - 64) return fail(NDPB_ERR_INVALID, "basic lattice must have < 2^31 vertices (as in the reference)");

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(NDPB_ERR_CUDA, "no CUDA device available (libndpb200 has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return fail(NDPB_ERR_INVALID, "device ordinal out of range");
    CU(cudaSetDevice(device));

    ndpb_table *t = new ndpb_table();
    t->device = device;
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) t->sm_count = prop.multiProcessorCount;
    NdpGeom &g = t->g;
    g.naxes = naxes; g.nbasic = nbasic; g.vdim = vdim;
    for (int a = 0, off = 0; a < naxes; a++) { g.len[a] = axlen[a]; g.tickoff[a] = off; off += axlen[a]; }
    for (int a = naxes - 1; a >= 0; a--) {
        g.vstride[a] = (a == naxes - 1) ? 1 : g.vstride[a + 1] * g.len[a + 1];
        g.cstride[a] = (a == naxes - 1) ? 1 : g.cstride[a + 1] * (g.len[a + 1] - 1);
    }
    for (int a = nbasic - 1; a >= 0; a--) g.bstride[a] = (a == nbasic - 1) ? 1 : g.bstride[a + 1] * g.len[a + 1];
    g.nverts = (int) nverts;
    g.ncells = ncells;
    t->nticks = nticks;
    t->grid_doubles = nall * vdim;

    auto bail = [&](int rc) { ndpb_table_destroy(t); return rc; };
#define CUT(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
        fail(e_ == cudaErrorMemoryAllocation ? NDPB_ERR_NOMEM : NDPB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
        return bail(e_ == cudaErrorMemoryAllocation ? NDPB_ERR_NOMEM : NDPB_ERR_CUDA); } } while (0)

    CUT(cudaStreamCreateWithFlags(&t->s_owned, cudaStreamNonBlocking));
    t->s_compute = t->s_owned;
    CUT(cudaStreamCreateWithFlags(&t->s_in, cudaStreamNonBlocking));
    CUT(cudaStreamCreateWithFlags(&t->s_out, cudaStreamNonBlocking));

    std::vector<double> flat(nticks);
    for (int a = 0; a < naxes; a++) {
        if (is_device_ptr(axes[a])) CUT(cudaMemcpy(flat.data() + g.tickoff[a], axes[a], sizeof(double) * axlen[a], cudaMemcpyDeviceToHost));
        else memcpy(flat.data() + g.tickoff[a], axes[a], sizeof(double) * axlen[a]);
    }
    ndp_geom_axes(g, flat.data());
    g.small = (nall < (1LL << 31) - 64) ? 1 : 0;
    for (int a = 0; a < naxes; a++) { g.vstride32[a] = g.small ? (int) g.vstride[a] : 0; g.cstride32[a] = g.small ? (int) g.cstride[a] : 0; }
    for (int a = 0; a < nbasic; a++) g.inv_bstride[a] = 1.0 / (double) g.bstride[a];
    CUT(cudaMalloc(&t->ticks, sizeof(double) * nticks));
    CUT(cudaMemcpyAsync(t->ticks, flat.data(), sizeof(double) * nticks, cudaMemcpyHostToDevice, t->s_compute));
    CUT(cudaStreamSynchronize(t->s_compute));

    if (grid) {
        t->has_grid = true;
        CUT(cudaMalloc(&t->grid, sizeof(double) * t->grid_doubles));
        // A device-resident grid: cudaMemcpy of device memory returns before the copy has run and is ordered on the
        // legacy default stream only -- the table's kernels run on a NON-BLOCKING stream and would read the copy's
        // destination too early (masks computed from whatever the recycled pages held; seen as 1 flaky run in 17).
        // Registering a table is a one-off: wait for the grid's producer, copy on the table's stream, wait again.
        if (is_device_ptr(grid)) CUT(cudaDeviceSynchronize());
        CUT(cudaMemcpyAsync(t->grid, grid, sizeof(double) * t->grid_doubles, cudaMemcpyDefault, t->s_compute));
        CUT(cudaStreamSynchronize(t->s_compute));
        const long long words = (nverts + 31) / 32;
        uint32_t *vbits = nullptr, *hbits = nullptr;
        CUT(cudaMalloc(&vbits, sizeof(uint32_t) * words));
        CUT(cudaMalloc(&hbits, sizeof(uint32_t) * words));
        const int nb = blocks_for(words * 32);
        k_vmask_bits<<<nb, NDP_BLOCK, 0, t->s_compute>>>(t->grid, g.nverts, g.vstride[nbasic - 1] * vdim, vbits);
        k_hcmask_bits<<<nb, NDP_BLOCK, 0, t->s_compute>>>(g, vbits, hbits);
        CUT(cudaGetLastError());
        int all_axes[NDP_MAX_AXES];
        for (int a = 0; a < nbasic; a++) all_axes[a] = a;
        int rc = build_pyramid(t, t->pyr[0], vbits, nbasic, all_axes, g.len, true);
        if (rc == NDPB_OK) rc = build_pyramid(t, t->pyr[1], hbits, nbasic, all_axes, g.len, true);
        if (rc == NDPB_OK) rc = build_reduced(t, 0);
        if (rc == NDPB_OK) rc = build_reduced(t, 1);
        if (rc == NDPB_OK) rc = maybe_build_cells(t);
        if (rc == NDPB_OK) rc = setup_fused(t);
        if (rc != NDPB_OK) return bail(rc);
    }
#undef CUT
    *out = t;
    return NDPB_OK;
}

extern "C" int64_t ndpb_table_nverts(const ndpb_table *t) { return t ? t->g.nverts : 0; }

extern "C" int ndpb_table_set_option(ndpb_table *t, const char *key, int64_t value)
{
    if (!t || !key) return fail(NDPB_ERR_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(t->mu);
    CU(cudaSetDevice(t->device));
    std::string k(key);
    if (k == "gather") {
        if (value < 0 || value > 2) return fail(NDPB_ERR_INVALID, "gather must be 0, 1 or 2");
        t->opt_gather = (int) value;
        if (value == 1 && (t->cells || t->blocks)) { CU(cudaDeviceSynchronize()); drop_cells(t); }
        if (value != 1 && t->has_grid) return maybe_build_cells(t);
        return NDPB_OK;
    }
    if (k == "block_order") {
        if (value < 0 || value > t->g.naxes) return fail(NDPB_ERR_INVALID, "block_order must be 0 (auto: the largest that fits) or 1..naxes");
        t->opt_block_order = (int) value;
        if (t->has_grid && t->opt_gather != 1) { CU(cudaDeviceSynchronize()); drop_cells(t); return maybe_build_cells(t); }
        return NDPB_OK;
    }
    if (k == "linear_search") {
        if (value < 0 || value > 2) return fail(NDPB_ERR_INVALID, "linear_search must be 0, 1 or 2");
        t->opt_linear_search = (int) value;
        return NDPB_OK;
    }
    if (k == "stream") {
        // run the kernels on a caller-owned stream (e.g. torch's current stream) so that the
        // caller's CUDA events bracket them; 0 restores the table's own stream
        CU(cudaStreamSynchronize(t->s_compute));
        t->s_compute = value ? (cudaStream_t) (uintptr_t) value : t->s_owned;
        return NDPB_OK;
    }
    if (k == "staged") {
        if (value < 0 || value > 3) return fail(NDPB_ERR_INVALID, "staged must be 0 (auto: one stage, 4 CTAs/SM), 1 (off), 2 (two-stage ring) or 3 (one stage, 5 CTAs/SM)");
        t->opt_staged = (int) value;
        return NDPB_OK;
    }
    if (k == "fused") {
        if (value < 0 || value > 1) return fail(NDPB_ERR_INVALID, "fused must be 0 (auto) or 1 (never use the fused kernel)");
        t->opt_fused = (int) value;
        return NDPB_OK;
    }
    if (k == "nsm_subcells") {
        if (value != 1 && value != 2 && value != 4 && value != 8 && value != 16)
            return fail(NDPB_ERR_INVALID, "nsm_subcells must be 1, 2, 4, 8 or 16");
        CU(cudaDeviceSynchronize());
        t->opt_nsm_subcells = (int) value;
        for (int geo = 0; geo < 3; geo++) {          // maps are rebuilt lazily at the new resolution
            cudaFree(t->nsm_offs[geo]); cudaFree(t->nsm_ents[geo]); cudaFree(t->nsm_head[geo]);
            t->nsm_offs[geo] = nullptr; t->nsm_ents[geo] = nullptr; t->nsm_head[geo] = nullptr;
            t->nsm[geo] = NdpSiteMap{}; t->nsm_tried[geo] = false;
        }
        return NDPB_OK;
    }
    if (k == "async") {
        if (value < 0 || value > 1) return fail(NDPB_ERR_INVALID, "async must be 0 or 1");
        CU(cudaStreamSynchronize(t->s_compute));
        t->opt_async = (int) value;
        return NDPB_OK;
    }
    if (k == "chunk") {
        if (value < 1024) return fail(NDPB_ERR_INVALID, "chunk must be >= 1024");
        t->opt_chunk = value;
        return NDPB_OK;
    }
    if (k == "search_fast_blocks_per_sm" || k == "search_hard_blocks_per_sm") {
        if (value < 1 || value > 1024) return fail(NDPB_ERR_INVALID, k + " must be in 1..1024");
        (k == "search_fast_blocks_per_sm" ? t->opt_fast_blocks : t->opt_hard_blocks) = (int) value;
        return NDPB_OK;
    }
    return fail(NDPB_ERR_INVALID, "unknown option " + k);
}

extern "C" int ndpb_table_wait_stream(ndpb_table *t, uint64_t producer_stream)
{
    if (!t) return fail(NDPB_ERR_INVALID, "table is NULL");
    std::lock_guard<std::mutex> lk(t->mu);
    CU(cudaSetDevice(t->device));
    cudaStream_t src = (cudaStream_t) (uintptr_t) producer_stream;
    if (src == t->s_compute) return NDPB_OK;
    cudaEvent_t ev = nullptr;
    CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    cudaError_t e = cudaEventRecord(ev, src);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(t->s_compute, ev, 0);
    cudaEventDestroy(ev);
    CU(e);
    return NDPB_OK;
}

extern "C" int64_t ndpb_table_stat(const ndpb_table *t, const char *key)
{
    if (!t || !key) return -1;
    std::string k(key);
    if (k == "launches") return t->st_launches;
    if (k == "offgrid") return t->st_offgrid;
    if (k == "hard") return t->st_hard;
    if (k == "ties") return t->st_ties;
    if (k == "h2d_bytes") return t->st_h2d;
    if (k == "d2h_bytes") return t->st_d2h;
    if (k == "cellmajor") return t->cells ? 1 : 0;
    if (k == "block_order") return t->block_order;
    if (k == "slow") return t->st_slow;
    if (k == "fused") return t->st_fused;
    if (k == "use_hc") return t->use_hc ? 1 : 0;
    if (k == "nsm_regions_nearest") return t->nsm[NSM_GEO_VERTEX].valid ? t->nsm[NSM_GEO_VERTEX].nregions : 0;
    if (k == "nsm_regions_linear") return t->nsm[NSM_GEO_CENTRE].valid ? t->nsm[NSM_GEO_CENTRE].nregions : 0;
    if (k == "main_ns") return (int64_t) (t->st_main_ms * 1e6);
    if (k == "search_ns") return (int64_t) (t->st_search_ms * 1e6);
    if (k == "finish_ns") return (int64_t) (t->st_finish_ms * 1e6);
    if (k == "variant_axes_nearest") return t->red[0].dev.nvar;
    if (k == "variant_axes_linear") return t->red[1].dev.nvar;
    if (k == "tree_levels_nearest") return t->topo[0].levels;
    if (k == "tree_levels_linear") return t->topo[1].levels;
    if (k == "mask_count_nearest") return t->pyr[0].count;
    if (k == "mask_count_linear") return t->pyr[1].count;
    if (k == "sparse_search_nearest") return t->sparse_ids[0] ? 1 : 0;
    if (k == "sparse_search_linear") return t->sparse_ids[1] ? 1 : 0;
    return -1;
}

extern "C" int ndpb_table_masks(ndpb_table *t, int32_t *vmask, int32_t *hcmask)
{
    if (!t || !vmask || !hcmask) return fail(NDPB_ERR_INVALID, "NULL argument");
    if (!t->has_grid) return fail(NDPB_ERR_STATE, "axes-only table has no masks");
    std::lock_guard<std::mutex> lk(t->mu);
    CU(cudaSetDevice(t->device));
    int32_t *outs[2] = {vmask, hcmask};
    for (int m = 0; m < 2; m++) {
        int32_t *dst = outs[m];
        int32_t *tmp = nullptr;
        const bool dev = is_device_ptr(dst);
        if (!dev) CU(cudaMalloc(&tmp, sizeof(int32_t) * t->g.nverts));
        k_unpack_bits<<<blocks_for(t->g.nverts), NDP_BLOCK, 0, t->s_compute>>>(t->pyr[m].dev.bits[0], t->g.nverts, dev ? dst : tmp);
        CU(cudaGetLastError());
        if (!dev) {
            CU(cudaMemcpyAsync(dst, tmp, sizeof(int32_t) * t->g.nverts, cudaMemcpyDeviceToHost, t->s_compute));
            CU(cudaStreamSynchronize(t->s_compute));
            CU(cudaFree(tmp));
        }
    }
    CU(cudaStreamSynchronize(t->s_compute));
    return NDPB_OK;
}

// ---------------------------------------------------------------------------
// kd topology (lazy, per method -- like the reference's lazy trees, ndpolator.c:218-221)
// ---------------------------------------------------------------------------
struct BitSet {
    const uint32_t *bits;
    __host__ __device__ bool operator()(int v) const { return (bits[v >> 5] >> (v & 31)) & 1u; }
};

static int build_topology(ndpb_table *t, int m, int *parent, int *depthv, int &levels)
{
    const NdpGeom &g = t->g;
    cudaStream_t st = t->s_compute;
    const int nv = g.nverts;
    k_fill_int<<<t->sm_count * 8, NDP_BLOCK, 0, st>>>(parent, nv, -2);
    k_fill_int<<<t->sm_count * 8, NDP_BLOCK, 0, st>>>(depthv, nv, -1);

    // ascending ids of the masked vertices
    DevTemps tmp_bufs;
    int *ids[2] = {nullptr, nullptr}, *ss[2] = {nullptr, nullptr}, *se[2] = {nullptr, nullptr};
    int *flag = nullptr, *scan = nullptr, *d_count = nullptr, *d_active = nullptr;
    unsigned char *tmp = nullptr;
    size_t tmp_bytes = 0, need = 0;
    CU(tmp_bufs.alloc(&ids[0], (size_t) nv));
    CU(tmp_bufs.alloc(&d_count, 1));
    CU(tmp_bufs.alloc(&d_active, 1));
    thrust::counting_iterator<int> counting(0);
    BitSet pred{t->pyr[m].dev.bits[0]};
    CU(cub::DeviceSelect::If(nullptr, need, counting, ids[0], d_count, nv, pred, st));
    tmp_bytes = need;
    CU(tmp_bufs.alloc(&tmp, tmp_bytes));
    CU(cub::DeviceSelect::If(tmp, tmp_bytes, counting, ids[0], d_count, nv, pred, st));
    int N = 0;
    CU(cudaMemcpyAsync(&N, d_count, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    t->pyr[m].count = N;

    levels = 0;
    if (N > 0) {
        CU(tmp_bufs.alloc(&ids[1], (size_t) N));
        for (int b = 0; b < 2; b++) { CU(tmp_bufs.alloc(&ss[b], (size_t) N)); CU(tmp_bufs.alloc(&se[b], (size_t) N)); }
        CU(tmp_bufs.alloc(&flag, (size_t) N + 1));
        CU(tmp_bufs.alloc(&scan, (size_t) N + 1));
        CU(cub::DeviceScan::ExclusiveSum(nullptr, need, flag, scan, N + 1, st));
        if (need > tmp_bytes) { tmp_bufs.free_now(tmp); tmp_bytes = need; CU(tmp_bufs.alloc(&tmp, tmp_bytes)); }
        k_topo_init<<<blocks_for(N), NDP_BLOCK, 0, st>>>(N, ss[0], se[0], ids[0], parent, depthv);
        int cur = 0;
        for (int depth = 0;; depth++) {
            TopoArgs A{};
            A.g = g; A.N = N; A.axis = depth % g.nbasic; A.depth = depth;
            A.ids = ids[cur]; A.seg_start = ss[cur]; A.seg_end = se[cur];
            A.flag = flag; A.scan = scan;
            A.ids2 = ids[cur ^ 1]; A.seg_start2 = ss[cur ^ 1]; A.seg_end2 = se[cur ^ 1];
            A.parent = parent; A.depthv = depthv; A.active = d_active;
            CU(cudaMemsetAsync(d_active, 0, sizeof(int), st));
            k_topo_flags<<<blocks_for(N + 1), NDP_BLOCK, 0, st>>>(A);
            CU(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, flag, scan, N + 1, st));
            k_topo_scatter<<<blocks_for(N), NDP_BLOCK, 0, st>>>(A);
            CU(cudaGetLastError());
            int active = 0;
            CU(cudaMemcpyAsync(&active, d_active, sizeof(int), cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            cur ^= 1;
            levels = depth + 2;
            if (active == 0) break;
        }
    }
    CU(cudaStreamSynchronize(st));
    return NDPB_OK;
}

static int ensure_topology(ndpb_table *t, int m)
{
    Topology &T = t->topo[m];
    if (T.built) return NDPB_OK;
    int *parent = nullptr, *depthv = nullptr, levels = 0;
    {
        DevTemps hold;                   // parent / depth are released to the table only when the build succeeded
        CU(hold.alloc(&parent, (size_t) t->g.nverts));
        CU(hold.alloc(&depthv, (size_t) t->g.nverts));
        int rc = build_topology(t, m, parent, depthv, levels);
        if (rc) return rc;
        hold.release(parent); hold.release(depthv);
    }
    T.parent = parent; T.depth = depthv; T.levels = levels; T.built = true;
    return NDPB_OK;
}

extern "C" int ndpb_table_tree_topology(ndpb_table *t, int method, int32_t *parent, int32_t *depth)
{
    if (!t || !parent || !depth) return fail(NDPB_ERR_INVALID, "NULL argument");
    if (!t->has_grid) return fail(NDPB_ERR_STATE, "axes-only table has no tree");
    if (method != NDPB_METHOD_NEAREST && method != NDPB_METHOD_LINEAR) return fail(NDPB_ERR_INVALID, "method must be NEAREST or LINEAR");
    std::lock_guard<std::mutex> lk(t->mu);
    CU(cudaSetDevice(t->device));
    const int m = method == NDPB_METHOD_LINEAR ? 1 : 0;
    int rc = ensure_topology(t, m);
    if (rc != NDPB_OK) return rc;
    CU(cudaMemcpy(parent, t->topo[m].parent, sizeof(int) * t->g.nverts, cudaMemcpyDefault));
    CU(cudaMemcpy(depth, t->topo[m].depth, sizeof(int) * t->g.nverts, cudaMemcpyDefault));
    return NDPB_OK;
}

// ---------------------------------------------------------------------------
// batches
// ---------------------------------------------------------------------------
// work: 0 none, 1 the fused path's lists (slow list, tie list: 8 bytes per query), 2 also the three-pass path's
// work lists and import records (up to 84 more bytes per query with 5 axes -- only allocated when that path runs)
static int ensure_slot(ndpb_table *t, Slot &s, long long nq, bool stage_q, bool stage_interps, bool stage_dists,
                       bool stage_coords, int work)
{
    const NdpGeom &g = t->g;
    if (!s.ctr) {
        CU(cudaMalloc(&s.ctr, sizeof(NdpCounters)));
        CU(cudaMallocHost(&s.ctr_host, sizeof(NdpCounters)));
        cudaEvent_t *ev[] = {&s.in_done, &s.searched, &s.out_ready, &s.out_done};
        for (auto e : ev) CU(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
        cudaEvent_t *tv[] = {&s.t0, &s.t1, &s.t2, &s.t3};
        for (auto e : tv) CU(cudaEventCreate(e));
    }
    if ((stage_q || stage_interps || stage_dists) && nq > s.cap_q) {
        CU(cudaStreamSynchronize(t->s_compute)); CU(cudaStreamSynchronize(t->s_in)); CU(cudaStreamSynchronize(t->s_out));
        cudaFree(s.q); cudaFree(s.interps); cudaFree(s.dists);
        s.q = s.interps = s.dists = nullptr; s.cap_q = 0;
        CU(cudaMalloc(&s.q, sizeof(double) * nq * g.naxes));
        CU(cudaMalloc(&s.interps, sizeof(double) * nq * g.vdim));
        CU(cudaMalloc(&s.dists, sizeof(double) * nq));
        s.cap_q = nq;
    }
    if (stage_coords && nq > s.cap_coords) {
        CU(cudaStreamSynchronize(t->s_compute)); CU(cudaStreamSynchronize(t->s_out));
        cudaFree(s.coords); s.coords = nullptr; s.cap_coords = 0;
        CU(cudaMalloc(&s.coords, sizeof(int) * nq * g.naxes));
        s.cap_coords = nq;
    }
    if (work >= 1 && nq > s.cap_lists) {
        CU(cudaStreamSynchronize(t->s_compute));
        cudaFree(s.list); cudaFree(s.tielist);
        s.list = s.tielist = nullptr; s.cap_lists = 0;
        CU(cudaMalloc(&s.list, sizeof(int) * nq));
        CU(cudaMalloc(&s.tielist, sizeof(int) * nq));
        s.cap_lists = nq;
    }
    if (work >= 2 && nq > s.cap_work) {
        CU(cudaStreamSynchronize(t->s_compute));
        cudaFree(s.chosen); cudaFree(s.hardlist); cudaFree(s.rec_idx); cudaFree(s.rec_nrm);
        s.chosen = s.hardlist = s.rec_idx = nullptr; s.rec_nrm = nullptr; s.cap_work = 0;
        CU(cudaMalloc(&s.chosen, sizeof(int) * nq));
        CU(cudaMalloc(&s.hardlist, sizeof(int) * nq));
        CU(cudaMalloc(&s.rec_idx, sizeof(int) * nq * g.naxes));
        CU(cudaMalloc(&s.rec_nrm, sizeof(double) * nq * g.naxes));
        s.cap_work = nq;
    }
    return NDPB_OK;
}

static int launch_main(ndpb_table *t, const MainArgs &A, cudaStream_t st)
{
    const NdpGeom &g = t->g;
    int G = 1;
    while (G < g.vdim && G < 32) G <<= 1;
    const long long groups_per_block = NDP_BLOCK / G;
    const int blocks = (int) std::min<long long>((A.n + groups_per_block - 1) / groups_per_block, (long long) t->sm_count * 8);
    const size_t smem = smem_bytes(t);
    if (A.cells && g.naxes <= 6) CU(ndp_launch_main_cellmajor(A, g.naxes, G, blocks, smem, st));
    else CU(ndp_launch_main_strided(A, g.naxes, G, blocks, smem, st));
    t->st_launches++;
    return NDPB_OK;
}

static int launch_finish(ndpb_table *t, const FinishArgs &A, cudaStream_t st)
{
    int G = 1;
    while (G < t->g.vdim && G < 32) G <<= 1;
    CU(ndp_launch_finish(A, t->g.naxes, G, A.cells != nullptr, t->sm_count * 8, smem_bytes(t), st));
    t->st_launches++;
    return NDPB_OK;
}

// The bulk-copy staged gather serves scalar tables with a cell-major copy, 2..6 axes.
static bool staged_ok(const ndpb_table *t)
{
    return t->opt_staged != 1 && t->cells && t->g.vdim == 1 && t->g.naxes >= 2 && t->g.naxes <= 6 &&
           (size_t) t->nticks * sizeof(double) <= 16 * 1024;
}

// 0 (auto) -> one stage per warp at 4 CTAs per SM (measured best on B200, profiles/r01j); 3 -> one stage at 5 CTAs
static int staged_stages(const ndpb_table *t) { return t->opt_staged == 2 ? 2 : (t->opt_staged == 3 ? 1 : 3); }

static size_t staged_smem(const ndpb_table *t)
{
    const size_t ticks = ((size_t) t->nticks * 8 + 127) & ~(size_t) 127;
    const size_t slot = ((size_t) 8 << t->g.naxes) + 16;
    return ticks + (size_t) NDP_STAGED_WARPS * (staged_stages(t) == 2 ? 2 : 1) * 32 * slot;
}

template <int MODE>
static int launch_staged(ndpb_table *t, const StagedArgs &A, long long items, cudaStream_t st)
{
    CU(ndp_launch_staged(A, t->g.naxes, MODE, staged_stages(t), t->sm_count, items, staged_smem(t), st));
    t->st_launches++;
    return NDPB_OK;
}

static bool use_brute(const ndpb_table *t, int search)
{
    if (search != NDPB_SEARCH_LINEAR) return false;
    if (t->opt_linear_search == 1) return true;
    if (t->opt_linear_search == 2) return false;
    return t->g.nverts <= 4096;      // the literal scan only where it is cheap; identical results either way
}

// A mask with at most NDP_SPARSE_MAX set vertices, fewer than one in 64: the nearest search scores the set vertices
// themselves (k_search_sparse) instead of descending a pyramid that is empty almost everywhere.
#define NDP_SPARSE_MAX 16384
static bool sparse_ready(ndpb_table *t, int m)
{
    const Pyramid &P = t->pyr[m];
    if (P.count > NDP_SPARSE_MAX || (long long) P.count * 64 > t->g.nverts) return false;
    if (t->sparse_tried[m]) return t->sparse_ids[m] != nullptr;
    t->sparse_tried[m] = true;
    DevTemps hold;
    int *ids = nullptr, *d_count = nullptr;
    void *tmp = nullptr;
    size_t need = 0;
    cudaStream_t st = t->s_owned;
    thrust::counting_iterator<int> counting(0);
    BitSet pred{P.dev.bits[0]};
    if (hold.alloc(&ids, (size_t) std::max(P.count, 1)) != cudaSuccess || hold.alloc(&d_count, 1) != cudaSuccess) { cudaGetLastError(); return false; }
    if (cub::DeviceSelect::If(nullptr, need, counting, ids, d_count, t->g.nverts, pred, st) != cudaSuccess) { cudaGetLastError(); return false; }
    if (hold.alloc((unsigned char **) &tmp, need) != cudaSuccess) { cudaGetLastError(); return false; }
    if (cub::DeviceSelect::If(tmp, need, counting, ids, d_count, t->g.nverts, pred, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) { cudaGetLastError(); return false; }
    hold.release(ids);
    t->sparse_ids[m] = ids;
    return true;
}

// hard == false: pass 1 over the work list (or the literal scan); hard == true: the descent over A.sublist
static int launch_search(ndpb_table *t, SearchArgs &A, long long upper, bool hard, cudaStream_t st)
{
    const int m = A.method == NDPB_METHOD_LINEAR ? 1 : 0;
    A.g = t->g; A.F = t->pyr[m].d_desc; A.R = t->red[m].d_desc;
    A.have_topo = t->topo[m].built ? 1 : 0;
    A.topo.parent = t->topo[m].parent; A.topo.depth = t->topo[m].depth;
    A.ticks = t->ticks; A.nticks = t->nticks; A.ticks_in_smem = ticks_in_smem(t);
    const size_t smem = smem_bytes(t);
    const long long cap = (long long) t->sm_count * (hard ? t->opt_hard_blocks : (use_brute(t, A.search) ? 8 : t->opt_fast_blocks));
    if (hard) {
        const int blocks = (int) std::min<long long>((upper + NDP_BLOCK - 1) / NDP_BLOCK, cap);
        CU(ndp_launch_search_hard(A, std::max(blocks, 1), smem, st));
    } else if (use_brute(t, A.search)) {
        const int blocks = (int) std::min<long long>((upper * 32 + NDP_BLOCK - 1) / NDP_BLOCK, cap);
        CU(ndp_launch_search_brute(A, std::max(blocks, 1), smem, st));
    } else if (sparse_ready(t, m)) {
        A.sparse_ids = t->sparse_ids[m]; A.sparse_n = t->pyr[m].count;
        const int blocks = (int) std::min<long long>((upper * 32 + NDP_BLOCK - 1) / NDP_BLOCK, (long long) t->sm_count * 8);
        CU(ndp_launch_search_sparse(A, std::max(blocks, 1), smem, st));
    } else {
        const int blocks = (int) std::min<long long>((upper + NDP_BLOCK - 1) / NDP_BLOCK, cap);
        CU(ndp_launch_search_fast(A, t->g.naxes, std::max(blocks, 1), smem, st));
    }
    CU(cudaGetLastError());
    t->st_launches++;
    return NDPB_OK;
}

enum Task { TASK_NDPOLATE, TASK_NEAREST, TASK_DISTANCE };

// The fused kernel serves scalar tables with a cell-major copy, 2..6 axes, whose hypercube-mask bit decides
// "fully defined", and -- when extrapolating -- a search mode that has a nearest-site map.
static bool fused_ok(ndpb_table *t, int method, int search)
{
    if (t->opt_fused == 1 || t->g.naxes < 2 || t->g.naxes > 6 || !t->use_hc) return false;
    if (t->g.vdim == 1 && !t->blocks) return false;     // scalar tables gather from the blocked copy, vector tables from the grid
    if ((size_t) t->nticks * sizeof(double) > 16 * 1024) return false;
    if (method == NDPB_METHOD_NONE) return true;
    if (search == NDPB_SEARCH_LINEAR && t->opt_linear_search == 1) return false;    // the literal scan was asked for
    const int geo = nsm_geo_of_mode(ndp_mode_of(method, search));
    if (ensure_nsm(t, geo) != NDPB_OK) return false;
    return t->nsm[geo].valid != 0;
}

struct CachedImport { const int *idx = nullptr, *flg = nullptr; const double *nrm = nullptr; };

static int launch_fused(ndpb_table *t, Slot &s, const double *q, long long n, int method, int search,
                        double *interps, double *dists, cudaStream_t st, const CachedImport &cache = CachedImport())
{
    const NdpGeom &g = t->g;
    const int mode = ndp_mode_of(method, search);
    FusedArgs A{};
    A.g = g; A.ticks = t->ticks; A.nticks = t->nticks; A.cells = t->blocks; A.grid = t->grid;
    A.q = q; A.n = n; A.interps = interps; A.dists = dists;
    A.c_idx = cache.idx; A.c_flg = cache.flg; A.c_nrm = cache.nrm;
    if (g.vdim == 1) {
        // the plan's cell number is a block number of the blocked copy
        ndp_blocks_setup(g, t->block_order, A.B);
        for (int a = 0; a < g.naxes; a++) { A.g.cstride[a] = A.B.stride[a]; A.g.cstride32[a] = g.small ? (int) A.B.stride[a] : 0; }
    }
    // the hypercube-mask slice sits in shared memory when small (<= 3 axes have the room for 16 KB at full occupancy)
    const size_t hc_limit = (g.naxes <= 3 ? 16 : 8) * 1024;
    A.H = t->hcs; A.hc_in_smem = t->hcs.nwords * sizeof(uint32_t) <= hc_limit;
    if (method != NDPB_METHOD_NONE) A.M = t->nsm[nsm_geo_of_mode(mode)];
    A.slowlist = s.list; A.ctr = s.ctr;
    if (g.vdim > 1) {
        // k_fused_vec: the plan's cell number is the vertex below the hypercube; 2^gs lanes share a query
        for (int a = 0; a < g.naxes; a++) { A.g.cstride[a] = g.vstride[a]; A.g.cstride32[a] = g.vstride32[a]; }
        int G = 1;
        while (G < g.vdim && G < 32) { G <<= 1; A.gs++; }
    }
    const size_t fixed = (((size_t) t->nticks * 8 + 127) & ~(size_t) 127) +
                         (A.hc_in_smem ? (((size_t) t->hcs.nwords * 4 + 127) & ~(size_t) 127) : 0);
    cudaError_t e;
    switch (g.naxes) {
    case 2: e = ndp_launch_fused_2(A, method, mode, t->sm_count, fixed, st); break;
    case 3: e = ndp_launch_fused_3(A, method, mode, t->sm_count, fixed, st); break;
    case 4: e = ndp_launch_fused_4(A, method, mode, t->sm_count, fixed, st); break;
    case 5: e = ndp_launch_fused_5(A, method, mode, t->sm_count, fixed, st); break;
    default: e = ndp_launch_fused_6(A, method, mode, t->sm_count, fixed, st); break;
    }
    CU(e);
    t->st_launches++;
    return NDPB_OK;
}

// k_slow over a device list of query ids: the slow list of the fused kernel, or (tie == true) the tie list
static int launch_slow(ndpb_table *t, Slot &s, const double *q, long long n, int method, int search, double *interps, double *dists,
                       bool tie, cudaStream_t st, const CachedImport &cache = CachedImport())
{
    const int m = method == NDPB_METHOD_LINEAR ? 1 : 0;
    SlowArgs A{};
    A.g = t->g; A.F = t->pyr[m].d_desc; A.R = t->red[m].d_desc;
    A.have_topo = t->topo[m].built ? 1 : 0;
    A.topo.parent = t->topo[m].parent; A.topo.depth = t->topo[m].depth;
    A.ticks = t->ticks; A.nticks = t->nticks; A.ticks_in_smem = ticks_in_smem(t);
    A.grid = t->grid; A.q = q; A.method = method; A.search = search;
    A.list = tie ? s.tielist : s.list;
    A.count = tie ? &s.ctr->n_tie : &s.ctr->n_slow;
    A.count_host = 0; A.c_idx = cache.idx; A.c_flg = cache.flg; A.c_nrm = cache.nrm;
    A.interps = interps; A.dists = dists; A.tielist = s.tielist; A.count_searched = tie ? 0 : 1; A.ctr = s.ctr;
    const int blocks = (int) std::min<long long>((long long) t->sm_count * 2, (n + NDP_BLOCK - 1) / NDP_BLOCK);
    CU(ndp_launch_slow(A, std::max(blocks, 1), smem_bytes(t), st));
    t->st_launches++;
    return NDPB_OK;
}

// Enqueues everything that needs no host decision for one batch resident on the device.
static int batch_enqueue(ndpb_table *t, Slot &s, Task task, const double *q, long long n, int method, int search,
                         double *interps, double *dists, int *coords)
{
    cudaStream_t st = t->s_compute;
    CU(cudaMemsetAsync(s.ctr, 0, sizeof(NdpCounters), st));
    CU(cudaEventRecord(s.t0, st));
    s.fused = (task == TASK_NDPOLATE) && fused_ok(t, method, search);
    if (s.fused) {
        int rc = launch_fused(t, s, q, n, method, search, interps, dists, st);
        if (rc) return rc;
        CU(cudaEventRecord(s.t1, st));
        rc = launch_slow(t, s, q, n, method, search, interps, dists, false, st);
        if (rc) return rc;
        CU(cudaEventRecord(s.t2, st));
        CU(cudaEventRecord(s.t3, st));
        CU(cudaMemcpyAsync(s.ctr_host, s.ctr, sizeof(NdpCounters), cudaMemcpyDeviceToHost, st));
        CU(cudaEventRecord(s.searched, st));
        s.used = true;
        t->st_fused++;
        return NDPB_OK;
    }
    StagedArgs G{};
    G.g = t->g; G.ticks = t->ticks; G.nticks = t->nticks; G.cells = t->cells; G.q = q; G.n = n;
    G.interps = interps; G.dists = dists; G.method = method; G.offlist = s.list; G.chosen = s.chosen; G.ctr = s.ctr;
    G.rec_idx = s.rec_idx; G.rec_nrm = s.rec_nrm;
    if (task == TASK_NDPOLATE && staged_ok(t)) {
        int rc = launch_staged<0>(t, G, n, st);
        if (rc) return rc;
    } else if (task == TASK_NDPOLATE) {
        MainArgs M{};
        M.g = t->g; M.ticks = t->ticks; M.nticks = t->nticks; M.ticks_in_smem = ticks_in_smem(t);
        M.grid = t->grid; M.cells = t->cells; M.q = q; M.n = n;
        M.interps = interps; M.dists = dists; M.method = method; M.offlist = s.list; M.ctr = s.ctr;
        M.rec_idx = s.rec_idx; M.rec_nrm = s.rec_nrm;
        int rc = launch_main(t, M, st);
        if (rc) return rc;
    }
    CU(cudaEventRecord(s.t1, st));
    if (task != TASK_NDPOLATE || method != NDPB_METHOD_NONE) {
        SearchArgs S{};
        S.q = q; S.method = method; S.search = search;
        S.list = (task == TASK_NDPOLATE) ? s.list : nullptr;
        S.sublist = nullptr;
        S.count = (task == TASK_NDPOLATE) ? &s.ctr->n_off : nullptr;
        S.count_host = n;
        S.chosen = s.chosen; S.dists = dists; S.coords = coords; S.hardlist = s.hardlist; S.tielist = s.tielist; S.ctr = s.ctr;
        S.rec_idx = s.rec_idx; S.rec_nrm = s.rec_nrm;
        int rc = launch_search(t, S, n, false, st);
        if (rc) return rc;
        if (!use_brute(t, search) && !sparse_ready(t, method == NDPB_METHOD_LINEAR ? 1 : 0)) {   // pass 2: whatever pass 1 could not prove
            S.sublist = s.hardlist; S.count = &s.ctr->n_hard;
            rc = launch_search(t, S, n, true, st);
            if (rc) return rc;
        }
    }
    CU(cudaEventRecord(s.t2, st));
    if (task == TASK_NDPOLATE && method == NDPB_METHOD_LINEAR && staged_ok(t)) {
        int rc = launch_staged<1>(t, G, n, st);      // item count is read from the device counter; n bounds the grid
        if (rc) return rc;
    } else if (task == TASK_NDPOLATE && method != NDPB_METHOD_NONE) {
        FinishArgs F{};
        F.g = t->g; F.ticks = t->ticks; F.nticks = t->nticks; F.ticks_in_smem = ticks_in_smem(t);
        F.grid = t->grid; F.cells = t->cells; F.q = q; F.method = method;
        F.list = s.list; F.rec_idx = s.rec_idx; F.rec_nrm = s.rec_nrm; F.sublist = nullptr; F.count = &s.ctr->n_off; F.chosen = s.chosen; F.interps = interps;
        int rc = launch_finish(t, F, st);
        if (rc) return rc;
    }
    CU(cudaEventRecord(s.t3, st));
    CU(cudaMemcpyAsync(s.ctr_host, s.ctr, sizeof(NdpCounters), cudaMemcpyDeviceToHost, st));
    CU(cudaEventRecord(s.searched, st));
    s.used = true;
    return NDPB_OK;
}

// Waits for the batch, settles kd ties through the topology if any turned up.
static int batch_complete(ndpb_table *t, Slot &s, Task task, const double *q, long long n, int method, int search,
                          double *interps, double *dists, int *coords)
{
    cudaStream_t st = t->s_compute;
    CU(cudaEventSynchronize(s.searched));
    float ms = 0;
    if (cudaEventElapsedTime(&ms, s.t0, s.t1) == cudaSuccess) t->st_main_ms += ms;
    if (cudaEventElapsedTime(&ms, s.t1, s.t2) == cudaSuccess) t->st_search_ms += ms;
    if (cudaEventElapsedTime(&ms, s.t2, s.t3) == cudaSuccess) t->st_finish_ms += ms;
    const NdpCounters c = *s.ctr_host;
    t->st_offgrid += (task == TASK_NDPOLATE) ? c.n_off : (method == NDPB_METHOD_NONE ? 0 : n);
    t->st_hard += c.n_hard;
    t->st_slow += c.n_slow;
    if (s.fused && c.n_tie > 0) {
        // kd distance ties met by k_slow before the reference-shaped topology existed: build it, redo them
        const int m = method == NDPB_METHOD_LINEAR ? 1 : 0;
        int rc = ensure_topology(t, m);
        if (rc) return rc;
        rc = launch_slow(t, s, q, n, method, search, interps, dists, true, st);
        if (rc) return rc;
        t->st_ties += c.n_tie;
    } else if (c.n_tie > 0) {
        const int m = method == NDPB_METHOD_LINEAR ? 1 : 0;
        int rc = ensure_topology(t, m);
        if (rc) return rc;
        SearchArgs S{};
        S.q = q; S.method = method; S.search = search;
        S.list = (task == TASK_NDPOLATE) ? s.list : nullptr;
        S.sublist = s.tielist; S.count = &s.ctr->n_tie; S.count_host = 0;
        S.chosen = s.chosen; S.dists = dists; S.coords = coords; S.hardlist = s.hardlist; S.tielist = s.tielist; S.ctr = s.ctr;
        S.rec_idx = s.rec_idx; S.rec_nrm = s.rec_nrm;
        rc = launch_search(t, S, c.n_tie, true, st);
        if (rc) return rc;
        if (task == TASK_NDPOLATE) {
            FinishArgs F{};
            F.g = t->g; F.ticks = t->ticks; F.nticks = t->nticks; F.ticks_in_smem = ticks_in_smem(t);
            F.grid = t->grid; F.cells = t->cells; F.q = q; F.method = method;
            F.list = s.list; F.rec_idx = s.rec_idx; F.rec_nrm = s.rec_nrm; F.sublist = s.tielist; F.count = &s.ctr->n_tie; F.chosen = s.chosen; F.interps = interps;
            rc = launch_finish(t, F, st);
            if (rc) return rc;
        }
        t->st_ties += c.n_tie;
    }
    CU(cudaEventRecord(s.out_ready, st));
    return NDPB_OK;
}

static void reset_stats(ndpb_table *t)
{
    t->st_launches = t->st_offgrid = t->st_hard = t->st_ties = t->st_h2d = t->st_d2h = t->st_slow = t->st_fused = 0;
    t->st_main_ms = t->st_search_ms = t->st_finish_ms = 0;
    t->async_pending = false;
}

// Drives one API call: device-resident callers run as one batch per 2^30 queries;
// host callers are cut into chunks whose H2D copy, kernels and D2H copy overlap
// across two slots and three streams.
static int run_call_inner(ndpb_table *t, Task task, const double *q, long long n, int method, int search,
                          double *interps, double *dists, int *coords);

// On any error inside the chunk pipeline, copies of earlier chunks may still be in flight against the caller's
// host buffers and the slots' staging buffers: drain all three streams and reset the slots before reporting.
static int run_call(ndpb_table *t, Task task, const double *q, long long n, int method, int search,
                    double *interps, double *dists, int *coords)
{
    const int rc = run_call_inner(t, task, q, n, method, search, interps, dists, coords);
    if (rc != NDPB_OK) {
        const std::string keep = g_last_error;
        cudaStreamSynchronize(t->s_in); cudaStreamSynchronize(t->s_compute); cudaStreamSynchronize(t->s_out);
        cudaGetLastError();
        for (auto &sl : t->slot) sl.used = false;
        g_last_error = keep;
    }
    return rc;
}

static int run_call_inner(ndpb_table *t, Task task, const double *q, long long n, int method, int search,
                          double *interps, double *dists, int *coords)
{
    const NdpGeom &g = t->g;
    reset_stats(t);
    if (n == 0) return NDPB_OK;
    const bool q_dev = is_device_ptr(q);
    const bool i_dev = !interps || is_device_ptr(interps);
    const bool d_dev = !dists || is_device_ptr(dists);
    const bool c_dev = !coords || is_device_ptr(coords);
    const bool all_dev = q_dev && i_dev && d_dev && c_dev;
    const long long chunk = all_dev ? (1LL << 30) : t->opt_chunk;
    const long long nchunks = (n + chunk - 1) / chunk;
    const bool need_dists = true;     // the search always writes distances
    const int work_level = (task == TASK_NDPOLATE && fused_ok(t, method, search)) ? 1 : 2;

    if (t->opt_async && all_dev && nchunks == 1 && dists && (task != TASK_NEAREST || coords)) {
        // Stream-ordered call: everything is enqueued on the compute stream and the call returns.  The one host
        // decision of the synchronous path -- building the reference-shaped kd topology when an exact distance
        // tie turns up -- is taken ahead of time: with the topology in place the kernels settle ties themselves.
        Slot &s = t->slot[0];
        int rc = ensure_slot(t, s, n, false, false, false, false, work_level);
        if (rc) return rc;
        if (method != NDPB_METHOD_NONE && search == NDPB_SEARCH_KDTREE) {
            rc = ensure_topology(t, method == NDPB_METHOD_LINEAR ? 1 : 0);
            if (rc) return rc;
        }
        rc = batch_enqueue(t, s, task, q, n, method, search, interps, dists, coords);
        if (rc) return rc;
        t->async_pending = true;
        return NDPB_OK;
    }

    auto stage_in = [&](long long c) -> int {
        Slot &s = t->slot[c & 1];
        const long long lo = c * chunk, cnt = std::min(chunk, n - lo);
        int rc = ensure_slot(t, s, cnt, !q_dev || !i_dev || !d_dev, false, false, coords && !c_dev, work_level);
        if (rc) return rc;
        if (!q_dev) {
            if (s.used) CU(cudaStreamWaitEvent(t->s_in, s.out_done, 0));   // slot's previous results are out
            CU(cudaMemcpyAsync(s.q, q + lo * g.naxes, sizeof(double) * cnt * g.naxes, cudaMemcpyHostToDevice, t->s_in));
            CU(cudaEventRecord(s.in_done, t->s_in));
            t->st_h2d += sizeof(double) * cnt * g.naxes;
        }
        return NDPB_OK;
    };

    int rc = stage_in(0);
    if (rc) return rc;
    for (long long c = 0; c < nchunks; c++) {
        Slot &s = t->slot[c & 1];
        const long long lo = c * chunk, cnt = std::min(chunk, n - lo);
        const double *qd = q_dev ? q + lo * g.naxes : s.q;
        double *id = !interps ? nullptr : (i_dev ? interps + lo * g.vdim : s.interps);
        double *dd = (dists && d_dev) ? dists + lo : s.dists;
        int *cd = !coords ? nullptr : (c_dev ? coords + lo * g.naxes : s.coords);
        if (!dd) { rc = ensure_slot(t, s, cnt, true, false, false, false, work_level); if (rc) return rc; dd = s.dists; }
        if (!q_dev) CU(cudaStreamWaitEvent(t->s_compute, s.in_done, 0));
        if (s.used) CU(cudaStreamWaitEvent(t->s_compute, s.out_done, 0));
        rc = batch_enqueue(t, s, task, qd, cnt, method, search, id, dd, cd);
        if (rc) return rc;
        if (c + 1 < nchunks) { rc = stage_in(c + 1); if (rc) return rc; }
        rc = batch_complete(t, s, task, qd, cnt, method, search, id, dd, cd);
        if (rc) return rc;
        CU(cudaStreamWaitEvent(t->s_out, s.out_ready, 0));
        if (interps && !i_dev) {
            CU(cudaMemcpyAsync(interps + lo * g.vdim, s.interps, sizeof(double) * cnt * g.vdim, cudaMemcpyDeviceToHost, t->s_out));
            t->st_d2h += sizeof(double) * cnt * g.vdim;
        }
        if (dists && !d_dev && need_dists) {
            CU(cudaMemcpyAsync(dists + lo, s.dists, sizeof(double) * cnt, cudaMemcpyDeviceToHost, t->s_out));
            t->st_d2h += sizeof(double) * cnt;
        }
        if (coords && !c_dev) {
            CU(cudaMemcpyAsync(coords + lo * g.naxes, s.coords, sizeof(int) * cnt * g.naxes, cudaMemcpyDeviceToHost, t->s_out));
            t->st_d2h += sizeof(int) * cnt * g.naxes;
        }
        CU(cudaEventRecord(s.out_done, t->s_out));
    }
    CU(cudaStreamSynchronize(t->s_compute));
    CU(cudaStreamSynchronize(t->s_out));
    return NDPB_OK;
}

// Folds the counters of the last stream-ordered call into the stats (the stream has been drained).
static void fold_async(ndpb_table *t)
{
    if (!t->async_pending) return;
    t->async_pending = false;
    Slot &s = t->slot[0];
    float ms = 0;
    if (cudaEventElapsedTime(&ms, s.t0, s.t1) == cudaSuccess) t->st_main_ms += ms;
    if (cudaEventElapsedTime(&ms, s.t1, s.t2) == cudaSuccess) t->st_search_ms += ms;
    if (cudaEventElapsedTime(&ms, s.t2, s.t3) == cudaSuccess) t->st_finish_ms += ms;
    cudaGetLastError();
    const NdpCounters c = *s.ctr_host;
    t->st_offgrid += c.n_off;
    t->st_hard += c.n_hard;
    t->st_slow += c.n_slow;
    t->st_ties += c.n_tie_resolved;
}

extern "C" int ndpb_table_sync(ndpb_table *t)
{
    if (!t) return fail(NDPB_ERR_INVALID, "table is NULL");
    std::lock_guard<std::mutex> lk(t->mu);
    CU(cudaSetDevice(t->device));
    CU(cudaStreamSynchronize(t->s_compute));
    fold_async(t);
    return NDPB_OK;
}

static int check_call(ndpb_table *t, const void *q, long long n, bool needs_grid)
{
    if (!t) return fail(NDPB_ERR_INVALID, "table is NULL");
    if (n < 0) return fail(NDPB_ERR_INVALID, "negative query count");
    if (n > 0 && !q) return fail(NDPB_ERR_INVALID, "query_pts is NULL");
    if (needs_grid && !t->has_grid) return fail(NDPB_ERR_STATE, "this call needs a table created with a grid");
    return NDPB_OK;
}

extern "C" int ndpb_ndpolate(ndpb_table *t, const double *query_pts, int64_t n, int method, int search,
                             double *interps, double *dists)
{
    int rc = check_call(t, query_pts, n, true);
    if (rc) return rc;
    if (method < NDPB_METHOD_NONE || method > NDPB_METHOD_LINEAR) return fail(NDPB_ERR_INVALID, "invalid extrapolation method");
    if (search < NDPB_SEARCH_KDTREE || search > NDPB_SEARCH_LINEAR) return fail(NDPB_ERR_INVALID, "invalid search algorithm");
    if (n > 0 && (!interps || !dists)) return fail(NDPB_ERR_INVALID, "output pointer is NULL");
    std::lock_guard<std::mutex> lk(t->mu);
    CU(cudaSetDevice(t->device));
    return run_call(t, TASK_NDPOLATE, query_pts, n, method, search, interps, dists, nullptr);
}

extern "C" int ndpb_find_nearest(ndpb_table *t, const double *query_pts, int64_t n, int method, int search,
                                 int32_t *coords, double *dist)
{
    int rc = check_call(t, query_pts, n, true);
    if (rc) return rc;
    if (method != NDPB_METHOD_NEAREST && method != NDPB_METHOD_LINEAR) return fail(NDPB_ERR_INVALID, "method must be NEAREST or LINEAR");
    if (search < NDPB_SEARCH_KDTREE || search > NDPB_SEARCH_LINEAR) return fail(NDPB_ERR_INVALID, "invalid search algorithm");
    if (n > 0 && (!coords || !dist)) return fail(NDPB_ERR_INVALID, "output pointer is NULL");
    std::lock_guard<std::mutex> lk(t->mu);
    CU(cudaSetDevice(t->device));
    return run_call(t, TASK_NEAREST, query_pts, n, method, search, nullptr, dist, coords);
}

extern "C" int ndpb_distance(ndpb_table *t, const double *query_pts, int64_t n, double *dists)
{
    int rc = check_call(t, query_pts, n, true);
    if (rc) return rc;
    if (n > 0 && !dists) return fail(NDPB_ERR_INVALID, "output pointer is NULL");
    std::lock_guard<std::mutex> lk(t->mu);
    CU(cudaSetDevice(t->device));
    return run_call(t, TASK_DISTANCE, query_pts, n, NDPB_METHOD_LINEAR, NDPB_SEARCH_KDTREE, nullptr, dists, nullptr);
}

// taps: simple synchronous staging (they are parity instruments, not the hot path)
template <class T>
struct Staged {
    T *dev = nullptr; T *user = nullptr; size_t count = 0; bool own = false;
    int in(const T *p, size_t n, cudaStream_t st) {       // st: the stream of the kernels that will read the copy
        user = const_cast<T *>(p); count = n;
        if (is_device_ptr(p) || n == 0) { dev = user; return NDPB_OK; }
        own = true;
        CU(cudaMalloc(&dev, sizeof(T) * n));
        CU(cudaMemcpyAsync(dev, p, sizeof(T) * n, cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));
        return NDPB_OK;
    }
    int out(T *p, size_t n) {
        user = p; count = n;
        if (is_device_ptr(p) || n == 0) { dev = p; return NDPB_OK; }
        own = true;
        CU(cudaMalloc(&dev, sizeof(T) * n));
        return NDPB_OK;
    }
    int back() {
        if (own) CU(cudaMemcpy(user, dev, sizeof(T) * count, cudaMemcpyDeviceToHost));
        return NDPB_OK;
    }
    ~Staged() { if (own) cudaFree(dev); }
};

extern "C" int ndpb_find(ndpb_table *t, const double *query_pts, int64_t n,
                         int32_t *indices, int32_t *flags, double *normed)
{
    int rc = check_call(t, query_pts, n, false);
    if (rc) return rc;
    if (n > 0 && (!indices || !flags || !normed)) return fail(NDPB_ERR_INVALID, "output pointer is NULL");
    std::lock_guard<std::mutex> lk(t->mu);
    CU(cudaSetDevice(t->device));
    reset_stats(t);
    if (n == 0) return NDPB_OK;
    const size_t e = (size_t) n * t->g.naxes;
    Staged<double> q, nrm; Staged<int32_t> idx, flg;
    if ((rc = q.in(query_pts, e, t->s_compute)) || (rc = idx.out(indices, e)) || (rc = flg.out(flags, e)) || (rc = nrm.out(normed, e))) return rc;
    ImportArgs A{};
    A.g = t->g; A.ticks = t->ticks; A.nticks = t->nticks; A.ticks_in_smem = ticks_in_smem(t);
    A.q = q.dev; A.n = n; A.indices = idx.dev; A.flags = flg.dev; A.normed = nrm.dev;
    const int blocks = (int) std::min<long long>(((long long) e + NDP_BLOCK - 1) / NDP_BLOCK, (long long) t->sm_count * 16);
    CU(ndp_launch_import(A, blocks, smem_bytes(t), t->s_compute));
    t->st_launches++;
    CU(cudaStreamSynchronize(t->s_compute));
    if ((rc = idx.back()) || (rc = flg.back()) || (rc = nrm.back())) return rc;
    return NDPB_OK;
}

extern "C" int ndpb_hypercubes(ndpb_table *t, const double *normed, const int32_t *indices, const int32_t *flags,
                               int64_t n, int32_t *dim, int32_t *fdhc, double *v)
{
    int rc = check_call(t, normed, n, true);
    if (rc) return rc;
    if (n > 0 && (!indices || !flags || !dim || !fdhc || !v)) return fail(NDPB_ERR_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(t->mu);
    CU(cudaSetDevice(t->device));
    reset_stats(t);
    if (n == 0) return NDPB_OK;
    const size_t e = (size_t) n * t->g.naxes;
    const size_t slot = ((size_t) 1 << t->g.naxes) * t->g.vdim;
    Staged<double> nrm, vv; Staged<int32_t> idx, flg, dm, fd;
    if ((rc = nrm.in(normed, e, t->s_compute)) || (rc = idx.in(indices, e, t->s_compute)) || (rc = flg.in(flags, e, t->s_compute)) ||
        (rc = dm.out(dim, n)) || (rc = fd.out(fdhc, n)) || (rc = vv.out(v, (size_t) n * slot))) return rc;
    HypercubeArgs A{};
    A.g = t->g; A.grid = t->grid; A.normed = nrm.dev; A.indices = idx.dev; A.flags = flg.dev; A.n = n;
    A.dim = dm.dev; A.fdhc = fd.dev; A.v = vv.dev;
    CU(ndp_launch_hypercubes(A, blocks_for(n), t->s_compute));
    t->st_launches++;
    CU(cudaStreamSynchronize(t->s_compute));
    if ((rc = dm.back()) || (rc = fd.back()) || (rc = vv.back())) return rc;
    return NDPB_OK;
}

// ---------------------------------------------------------------------------
// ndpb_group: ONE call, several GPUs.  The reference's ndpolate is one call on one process
// (ndpolator/ndpolator.py:253-349); the batch is embarrassingly parallel over queries, so a group
// holds one replicated table per device and cuts a host batch into contiguous ranges, one host
// thread + stream set per device.  Every device copies its results straight into its own slice of
// the caller's output arrays -- that IS the final gather; there is no collective.
// ---------------------------------------------------------------------------
struct ndpb_group {
    std::vector<ndpb_table *> tables;
    std::mutex mu;
    long long min_per_device = 1 << 16;     // below this a device is not worth waking up
    long long st_used_devices = 0;
};

extern "C" int ndpb_group_destroy(ndpb_group *g)
{
    if (!g) return NDPB_OK;
    for (auto t : g->tables) ndpb_table_destroy(t);
    delete g;
    return NDPB_OK;
}

extern "C" int ndpb_group_create(const double *const *axes, const int32_t *axlen, int naxes, int nbasic,
                                 const double *grid, int vdim, const int32_t *devices, int ndev, ndpb_group **out)
{
    if (!out) return fail(NDPB_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!devices || ndev < 1 || ndev > 64) return fail(NDPB_ERR_INVALID, "devices must list 1..64 device ordinals");
    for (int i = 0; i < ndev; i++)
        for (int j = 0; j < i; j++)
            if (devices[i] == devices[j]) return fail(NDPB_ERR_INVALID, "a device is listed twice");
    ndpb_group *g = new ndpb_group();
    g->tables.assign(ndev, nullptr);
    std::vector<int> rcs(ndev, NDPB_OK);
    std::vector<std::string> errs(ndev);
    // replicate the table: one host thread per device uploads the grid and builds masks, pyramids and the
    // cell-major copy on its own GPU (per-device H2D)
    std::vector<std::thread> th;
    for (int i = 0; i < ndev; i++)
        th.emplace_back([&, i] {
            rcs[i] = ndpb_table_create(axes, axlen, naxes, nbasic, grid, vdim, devices[i], &g->tables[i]);
            if (rcs[i]) errs[i] = ndpb_last_error();
        });
    for (auto &x : th) x.join();
    for (int i = 0; i < ndev; i++)
        if (rcs[i]) { const int rc = rcs[i]; const std::string msg = errs[i]; ndpb_group_destroy(g); return fail(rc, msg); }
    *out = g;
    return NDPB_OK;
}

extern "C" int ndpb_group_size(const ndpb_group *g) { return g ? (int) g->tables.size() : 0; }
extern "C" ndpb_table *ndpb_group_table(ndpb_group *g, int i)
{
    return (g && i >= 0 && i < (int) g->tables.size()) ? g->tables[i] : nullptr;
}

extern "C" int ndpb_group_set_option(ndpb_group *g, const char *key, int64_t value)
{
    if (!g || !key) return fail(NDPB_ERR_INVALID, "NULL argument");
    if (std::string(key) == "min_queries_per_device") {
        if (value < 1) return fail(NDPB_ERR_INVALID, "min_queries_per_device must be >= 1");
        g->min_per_device = value;
        return NDPB_OK;
    }
    for (auto t : g->tables) { int rc = ndpb_table_set_option(t, key, value); if (rc) return rc; }
    return NDPB_OK;
}

extern "C" int64_t ndpb_group_stat(const ndpb_group *g, const char *key)
{
    if (!g || !key) return -1;
    if (std::string(key) == "used_devices") return g->st_used_devices;
    int64_t sum = 0;
    for (auto t : g->tables) { const int64_t v = ndpb_table_stat(t, key); if (v < 0) return v; sum += v; }
    return sum;
}

// task: 0 ndpolate, 1 distance
static int group_call(ndpb_group *g, int task, const double *q, int64_t n, int method, int search, double *interps, double *dists)
{
    if (!g) return fail(NDPB_ERR_INVALID, "group is NULL");
    if (n < 0) return fail(NDPB_ERR_INVALID, "negative query count");
    if (n > 0 && (!q || !dists || (task == 0 && !interps))) return fail(NDPB_ERR_INVALID, "NULL argument");
    if (is_device_ptr(q) || is_device_ptr(interps) || is_device_ptr(dists))
        return fail(NDPB_ERR_INVALID, "ndpb_group takes host buffers (a device buffer belongs to one GPU: use that GPU's table)");
    std::lock_guard<std::mutex> lk(g->mu);
    const int ndev = (int) g->tables.size();
    int use = (int) std::min<long long>(ndev, std::max<long long>(1, n / g->min_per_device));
    g->st_used_devices = use;
    const int na = g->tables[0]->g.naxes, V = g->tables[0]->g.vdim;
    std::vector<int> rcs(use, NDPB_OK);
    std::vector<std::string> errs(use);
    auto work = [&](int i) {
        const long long lo = (n * i) / use, hi = (n * (i + 1)) / use;      // contiguous range [g N / G, (g + 1) N / G)
        ndpb_table *t = g->tables[i];
        if (task == 0) rcs[i] = ndpb_ndpolate(t, q + lo * na, hi - lo, method, search, interps + lo * V, dists + lo);
        else rcs[i] = ndpb_distance(t, q + lo * na, hi - lo, dists + lo);
        if (rcs[i]) errs[i] = ndpb_last_error();
    };
    if (use == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int i = 0; i < use; i++) th.emplace_back(work, i);
        for (auto &x : th) x.join();
    }
    for (int i = 0; i < use; i++)
        if (rcs[i]) return fail(rcs[i], errs[i]);
    return NDPB_OK;
}

extern "C" int ndpb_group_ndpolate(ndpb_group *g, const double *query_pts, int64_t n, int method, int search,
                                   double *interps, double *dists)
{
    return group_call(g, 0, query_pts, n, method, search, interps, dists);
}

extern "C" int ndpb_group_distance(ndpb_group *g, const double *query_pts, int64_t n, double *dists)
{
    return group_call(g, 1, query_pts, n, NDPB_METHOD_LINEAR, NDPB_SEARCH_KDTREE, nullptr, dists);
}

// ---------------------------------------------------------------------------
// Hypercube-caching path (reference README "Hypercube caching", ndpolator/ndpolator.py:110-187): the caller
// keeps the import products of a batch (ndpb_find) -- on the device, if it passes device pointers -- and
// interpolates from them, e.g. on several tables that share their axes.  Tables the fused path serves run
// k_fused / k_fused_vec with the cached products in place of the import (then k_slow for the few the plan
// cannot settle), the others the exact generic kernel (k_slow) over all queries; identical results to
// ndpb_ndpolate on the original query points.
// ---------------------------------------------------------------------------
extern "C" int ndpb_ndpolate_cached(ndpb_table *t, const double *normed, const int32_t *indices, const int32_t *flags,
                                    int64_t n, int method, int search, double *interps, double *dists)
{
    int rc = check_call(t, normed, n, true);
    if (rc) return rc;
    if (method < NDPB_METHOD_NONE || method > NDPB_METHOD_LINEAR) return fail(NDPB_ERR_INVALID, "invalid extrapolation method");
    if (search < NDPB_SEARCH_KDTREE || search > NDPB_SEARCH_LINEAR) return fail(NDPB_ERR_INVALID, "invalid search algorithm");
    if (n > 0 && (!indices || !flags || !interps || !dists)) return fail(NDPB_ERR_INVALID, "NULL argument");
    if (n >= (1LL << 31)) return fail(NDPB_ERR_INVALID, "at most 2^31 - 1 cached queries per call");
    std::lock_guard<std::mutex> lk(t->mu);
    CU(cudaSetDevice(t->device));
    reset_stats(t);
    if (n == 0) return NDPB_OK;
    const size_t e = (size_t) n * t->g.naxes;
    Staged<double> nrm, out_i, out_d; Staged<int32_t> idx, flg;
    if ((rc = nrm.in(normed, e, t->s_compute)) || (rc = idx.in(indices, e, t->s_compute)) || (rc = flg.in(flags, e, t->s_compute)) ||
        (rc = out_i.out(interps, (size_t) n * t->g.vdim)) || (rc = out_d.out(dists, (size_t) n))) return rc;
    Slot &s = t->slot[0];
    rc = ensure_slot(t, s, n, false, false, false, false, 1);
    if (rc) return rc;
    cudaStream_t st = t->s_compute;
    const int m = method == NDPB_METHOD_LINEAR ? 1 : 0;
    CachedImport cache;
    cache.idx = idx.dev; cache.flg = flg.dev; cache.nrm = nrm.dev;
    if (fused_ok(t, method, search)) {
        // the fused kernels with the cached import products in place of the import; what their plan cannot settle
        // goes through k_slow (also on the cached products), kd ties through the topology as in ndpb_ndpolate
        CU(cudaMemsetAsync(s.ctr, 0, sizeof(NdpCounters), st));
        rc = launch_fused(t, s, nullptr, n, method, search, out_i.dev, out_d.dev, st, cache);
        if (rc) return rc;
        rc = launch_slow(t, s, nullptr, n, method, search, out_i.dev, out_d.dev, false, st, cache);
        if (rc) return rc;
        CU(cudaMemcpyAsync(s.ctr_host, s.ctr, sizeof(NdpCounters), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        t->st_fused++;
        t->st_slow += s.ctr_host->n_slow;
        t->st_offgrid += s.ctr_host->n_off;
        if (s.ctr_host->n_tie > 0) {
            t->st_ties += s.ctr_host->n_tie;
            rc = ensure_topology(t, m);
            if (rc) return rc;
            rc = launch_slow(t, s, nullptr, n, method, search, out_i.dev, out_d.dev, true, st, cache);
            if (rc) return rc;
            CU(cudaStreamSynchronize(st));
        }
    } else
    for (int pass = 0; pass < 2; pass++) {
        if (pass == 0) CU(cudaMemsetAsync(s.ctr, 0, sizeof(NdpCounters), st));
        SlowArgs A{};
        A.g = t->g; A.F = t->pyr[m].d_desc; A.R = t->red[m].d_desc;
        A.have_topo = t->topo[m].built ? 1 : 0;
        A.topo.parent = t->topo[m].parent; A.topo.depth = t->topo[m].depth;
        A.ticks = t->ticks; A.nticks = t->nticks; A.ticks_in_smem = ticks_in_smem(t);
        A.grid = t->grid; A.q = nullptr; A.method = method; A.search = search;
        A.list = pass ? s.tielist : nullptr; A.count = &s.ctr->n_tie; A.count_host = n;
        A.c_idx = idx.dev; A.c_flg = flg.dev; A.c_nrm = nrm.dev;
        A.interps = out_i.dev; A.dists = out_d.dev; A.tielist = s.tielist; A.count_searched = pass ? 0 : 1; A.ctr = s.ctr;
        const int blocks = (int) std::min<long long>((n + NDP_BLOCK - 1) / NDP_BLOCK, (long long) t->sm_count * 8);
        CU(ndp_launch_slow(A, std::max(blocks, 1), smem_bytes(t), st));
        t->st_launches++;
        CU(cudaMemcpyAsync(s.ctr_host, s.ctr, sizeof(NdpCounters), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (pass == 1 || s.ctr_host->n_tie == 0) break;
        t->st_ties += s.ctr_host->n_tie;
        rc = ensure_topology(t, m);              // kd distance ties: build the reference-shaped topology, redo them
        if (rc) return rc;
    }
    t->st_hard += s.ctr_host->n_hard;
    if ((rc = out_i.back()) || (rc = out_d.back())) return rc;
    return NDPB_OK;
}
