// ndp_kernels.cuh -- sm_100a kernels of the ndpolate hot path.
//
//   k_import        per-axis index search tap            (ndp_query_pts_import, src/ndpolator.c:354-414)
//   k_hypercubes    hypercube assembly tap               (ndp_find_hypercubes, :416-508)
//   k_main          import + gather + reduce, fused      (:354-414 + :416-508 + :78-99 + :547-554, :648-667)
//                   -> interps/dists for points in a fully defined hypercube,
//                      compacted list of the rest
//   k_search_fast   exact nearest on the lattice, pass 1 (ndp_find_nearest, :188-352; kdtree.c:345-457)
//   k_search_hard   pass 2: pyramid descent for what pass 1 could not prove; also the tie pass
//   k_search_brute  warp-cooperative full scan           (:279-329)
//   k_finish        nearest copy / linear extrapolation  (:556-635)
//
// All kernels are grid-stride over their work list and read list lengths from
// device counters, so the host never synchronises between them.
#pragma once

#include <cuda_runtime.h>
#include "ndp_core.cuh"

#define NDP_BLOCK 256
// resident CTAs per SM the two search passes are compiled for (register budget = 65536 / (256 * OCC))
#ifdef NDP_FAST_OCC
#define NDP_FAST_BOUNDS __launch_bounds__(NDP_BLOCK, NDP_FAST_OCC)
#else
#define NDP_FAST_BOUNDS __launch_bounds__(NDP_BLOCK)
#endif
#ifndef NDP_HARD_OCC
#define NDP_HARD_OCC 2
#endif

struct NdpCounters {      // one per in-flight batch, zeroed before k_main
    int n_off;            // queries not in a fully defined hypercube (search needed)
    int n_tie;            // kd-mode ties met without topology
    int n_hard;           // queries that needed the pyramid descent
    int n_tie_resolved;   // ties settled through the kd topology
    int n_slow;           // fused path: queries the plan could not settle (redone by k_slow)
};

// ---------------------------------------------------------------------------
// axis staging: all ticks of all axes in shared memory when they fit
// ---------------------------------------------------------------------------
__device__ __forceinline__ const double *ndp_stage_ticks(const double *ticks, int nticks, bool in_smem, double *smem)
{
    if (!in_smem) return ticks;
    for (int i = threadIdx.x; i < nticks; i += blockDim.x) smem[i] = ticks[i];
    __syncthreads();
    return smem;
}

// ---------------------------------------------------------------------------
// k_import: indices / flags / normed for every (query, axis)
// ---------------------------------------------------------------------------
struct ImportArgs {
    NdpGeom g;
    const double *ticks; int nticks; int ticks_in_smem;
    const double *q; long long n;
    int *indices; int *flags; double *normed;
};

#if defined(NDP_TU_TAPS)
__global__ void __launch_bounds__(NDP_BLOCK) k_import(const __grid_constant__ ImportArgs A)
{
    extern __shared__ double s_ticks[];
    const double *tk = ndp_stage_ticks(A.ticks, A.nticks, A.ticks_in_smem, s_ticks);
    const int n = A.g.naxes;
    const long long total = A.n * n;
    // one thread per (query, axis) element: fully coalesced loads and stores
    for (long long e = (long long) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long) gridDim.x * blockDim.x) {
        const int a = (int) (e % n);
        int idx, flg; double nrm;
        ndp_import_axis_guess(tk + A.g.tickoff[a], A.g.len[a], A.g.inv_step[a], A.g.first[a], A.g.last[a], A.q[e], idx, flg, nrm);
        A.indices[e] = idx; A.flags[e] = flg; A.normed[e] = nrm;
    }
}

#endif  // NDP_TU_TAPS

// ---------------------------------------------------------------------------
// k_hypercubes: the tap behind cndpolator.hypercubes
// ---------------------------------------------------------------------------
struct HypercubeArgs {
    NdpGeom g;
    const double *grid;
    const double *normed; const int *indices; const int *flags; long long n;
    int *dim; int *fdhc; double *v;
};

#if defined(NDP_TU_TAPS)
__global__ void __launch_bounds__(NDP_BLOCK) k_hypercubes(const __grid_constant__ HypercubeArgs A)
{
    const NdpGeom &g = A.g;
    const int n = g.naxes, V = g.vdim;
    const long long slot = ((long long) 1 << n) * V;
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < A.n; i += (long long) gridDim.x * blockDim.x) {
        const int *idx = A.indices + i * n;
        const int *flg = A.flags + i * n;
        const double *nrm = A.normed + i * n;
        int full = 1, d = n;
        for (int a = 0; a < n; a++) {
            if (flg[a] & NDP_F_OOB) full = 0;
            if (flg[a] & NDP_F_ON_VERTEX) d--;
        }
        double *out = A.v + i * slot;
        for (int j = 0; j < (1 << d); j++) {
            long long vert = 0;
            int l = 0;
            for (int a = 0; a < n; a++) {
                int c;
                if (flg[a] & NDP_F_ON_VERTEX) c = nrm[a] > 0.5 ? idx[a] : idx[a] - 1;
                else {
                    int bit = (j >> (d - l - 1)) & 1;
                    c = idx[a] - 1 + bit;
                    if (c < bit) c = bit;
                    l++;
                }
                vert += g.vstride[a] * c;
            }
            const double *src = A.grid + vert * V;
            if (src[0] != src[0]) full = 0;
            for (int k = 0; k < V; k++) out[(long long) j * V + k] = src[k];
        }
        A.dim[i] = d;
        A.fdhc[i] = full;
    }
}

#endif  // NDP_TU_TAPS

// ---------------------------------------------------------------------------
// k_main
// ---------------------------------------------------------------------------
struct MainArgs {
    NdpGeom g;
    const double *ticks; int nticks; int ticks_in_smem;
    const double *grid;       // original layout
    const double *cells;      // cell-major copy or nullptr
    const double *q; long long n;
    double *interps; double *dists;
    int method;
    int *offlist;             // ids of queries that need the search (method != NONE)
    int *rec_idx; double *rec_nrm;   // import products of those queries, by list position (n x naxes), or nullptr
    NdpCounters *ctr;
};

#if defined(NDP_TU_MAIN)
// N: compile-time axis count (0 = runtime).  G: lanes cooperating on one query,
// each owning components k = lane, lane+G, ... (G = 1 for scalar tables).
template <int N, int G, bool CM>
__global__ void __launch_bounds__(NDP_BLOCK, (N >= 6 ? 1 : 2)) k_main(const __grid_constant__ MainArgs A)
{
    extern __shared__ double s_ticks[];
    const double *tk = ndp_stage_ticks(A.ticks, A.nticks, A.ticks_in_smem, s_ticks);
    const NdpGeom &g = A.g;
    const int n = N ? N : g.naxes;
    const int V = (G == 1) ? 1 : g.vdim;              // G == 1 is only launched for scalar tables
    const int lig = threadIdx.x % G;                  // lane in group
    const unsigned lane = threadIdx.x & 31;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(unsigned) (G - 1)));
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);

    const long long per_pass = (long long) gridDim.x * (NDP_BLOCK / G);
    const long long rounds = (A.n + per_pass - 1) / per_pass;
    for (long long r = 0; r < rounds; r++) {
        const long long i = r * per_pass + (long long) blockIdx.x * (NDP_BLOCK / G) + threadIdx.x / G;
        const bool live = i < A.n;
        bool need_search = false;
        int idx[N ? N : NDP_MAX_AXES], flg[N ? N : NDP_MAX_AXES];
        double nrm[N ? N : NDP_MAX_AXES];

        if (live) {
            if (G > 1 && N > 0 && G >= N) {
                // the lanes of a group split the axes (one index search each) and exchange the results,
                // instead of all repeating the whole import
                int my_idx = 1, my_flg = 0;
                double my_nrm = 0.0;
                if (lig < n)
                    ndp_import_axis_guess(tk + g.tickoff[lig], g.len[lig], g.inv_step[lig], g.first[lig], g.last[lig],
                                          A.q[i * n + lig], my_idx, my_flg, my_nrm);
                const int group_base = (int) (lane & ~(unsigned) (G - 1));
#pragma unroll
                for (int a = 0; a < (N ? N : 1); a++) {
                    idx[a] = __shfl_sync(gmask, my_idx, group_base + a);
                    flg[a] = __shfl_sync(gmask, my_flg, group_base + a);
                    nrm[a] = __shfl_sync(gmask, my_nrm, group_base + a);
                }
            } else ndp_import_query<N>(g, tk, A.q + i * n, idx, flg, nrm);
            unsigned pinmask, selbits;
            const bool oob = ndp_classify<N>(g, flg, nrm, pinmask, selbits);

            bool fdhc = !oob;
            if (!oob) {
                // component 0 decides whether the hypercube is fully defined (:491)
                for (int k0 = 0; k0 < V && fdhc; k0 += G) {
                    const int k = k0 + lig;
                    bool bad = false;
                    double val = 0.0;
                    if (k < V) val = ndp_eval_cell<N, CM, G == 1>(g, A.grid, A.cells, idx, nrm, pinmask, selbits, k, bad);
                    if (k0 == 0) {
                        if (G > 1) bad = __shfl_sync(gmask, (int) bad, lane & ~(unsigned) (G - 1)) != 0;
                        if (bad) { fdhc = false; break; }
                    }
                    if (k < V) A.interps[i * V + k] = val;
                }
            }
            if (fdhc) {
                if (lig == 0) A.dists[i] = 0.0;
            } else if (A.method == NDP_M_NONE) {
                for (int k = lig; k < V; k += G) A.interps[i * V + k] = qnan;     // :551-554
                if (lig == 0) A.dists[i] = 0.0;
            } else {
                need_search = (lig == 0);
            }
        }

        // warp-aggregated append to the off-grid list
        const unsigned votes = __ballot_sync(0xffffffffu, need_search);
        if (votes) {
            int base = 0;
            const int leader = __ffs(votes) - 1;
            if ((int) lane == leader) base = atomicAdd(&A.ctr->n_off, __popc(votes));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (need_search) {
                const int pos = base + __popc(votes & ((1u << lane) - 1u));
                A.offlist[pos] = (int) i;
                if (A.rec_idx)
                    for (int a = 0; a < n; a++) { A.rec_idx[(long long) pos * n + a] = idx[a]; A.rec_nrm[(long long) pos * n + a] = nrm[a]; }
            }
        }
    }
}

#endif  // NDP_TU_MAIN

// ---------------------------------------------------------------------------
// k_search / k_search_brute
// ---------------------------------------------------------------------------
struct SearchArgs {
    NdpGeom g;
    const NdpPyramid *F;    // occupancy pyramid over the whole basic lattice (device memory)
    const NdpPyramid *R;    // pyramid over the mask-variant axes only
    NdpTopo topo; int have_topo;
    const double *ticks; int nticks; int ticks_in_smem;
    const double *q;
    int method, search;
    const int *list;        // query id per work item, or nullptr (work item t = query t)
    const int *sublist;     // k_search_hard: positions into list to do (hard list or tie list)
    const int *count;       // device count of work items, or nullptr
    long long count_host;   // used when count == nullptr
    int *chosen;            // winning basic vertex per work item
    double *dists;          // per query id
    int *coords;            // optional (n x naxes) tap, per query id
    int *hardlist;          // positions whose fast path failed
    int *tielist;           // positions whose kd tie could not be settled yet
    const int *rec_idx; const double *rec_nrm;   // import products by list position (valid when list != nullptr), or nullptr
    const int *sparse_ids; int sparse_n;         // k_search_sparse: ascending ids of the mask's set vertices
    NdpCounters *ctr;
};

#if defined(NDP_TU_SEARCH)
__device__ __forceinline__ void ndp_publish(const SearchArgs &A, const NdpQuery &Q, long long t, long long i, NdpHit hit)
{
    const NdpGeom &g = A.g;
    const int mode = ndp_mode_of(A.method, A.search);
    int coords[NDP_MAX_AXES];
    double dist;
    if (hit.status == NDP_SEARCH_EMPTY) {
        // empty kd-tree: the reference falls through to the scan (:273-275), which scores every vertex 1e10
        hit.vertex = 0; hit.dsel = 1e10; hit.status = NDP_SEARCH_OK;
        ndp_coords_of(g, Q, 0, coords);
        dist = 1e10;
    } else {
        if (hit.has_coords) {
            for (int a = 0; a < g.nbasic; a++) coords[a] = hit.coords[a];
            for (int a = g.nbasic; a < g.naxes; a++) coords[a] = Q.acoord[a];
        } else ndp_coords_of(g, Q, hit.vertex, coords);
        dist = (mode >= NDP_MODE_LS_NEAREST) ? hit.dsel : ndp_reported_dist(g, A.method, Q, coords);
    }
    A.chosen[t] = hit.vertex;
    A.dists[i] = dist;
    if (A.coords) for (int a = 0; a < g.naxes; a++) A.coords[i * g.naxes + a] = coords[a];
}

// Pass 1: the rounded lattice point, proven optimal by its axis neighbours.  What it
// cannot settle goes to the hard list so that pass 2 runs with full warps.
template <int N, int MODE>
__global__ void NDP_FAST_BOUNDS k_search_fast(const __grid_constant__ SearchArgs A)
{
    extern __shared__ double s_ticks[];
    const double *tk = ndp_stage_ticks(A.ticks, A.nticks, A.ticks_in_smem, s_ticks);
    const NdpGeom &g = A.g;
    const int n = N ? N : g.naxes;
    const int mode = ndp_mode_of(A.method, A.search);
    const unsigned lane = threadIdx.x & 31;
    const long long total = A.count ? (long long) *A.count : A.count_host;
    const long long stride = (long long) gridDim.x * blockDim.x;
    const long long rounds = (total + stride - 1) / stride;
    for (long long r = 0; r < rounds; r++) {
        const long long t = r * stride + (long long) blockIdx.x * blockDim.x + threadIdx.x;
        bool hard = false;
        if (t < total) {
            const long long i = A.list ? A.list[t] : t;
            int idx[N ? N : NDP_MAX_AXES], flg[N ? N : NDP_MAX_AXES];
            double nrm[N ? N : NDP_MAX_AXES];
            if (A.list && A.rec_idx) {
#pragma unroll
                for (int a = 0; a < (N ? N : NDP_MAX_AXES); a++)
                    if (a < n) { idx[a] = A.rec_idx[t * n + a]; nrm[a] = A.rec_nrm[t * n + a]; }
            } else ndp_import_query_seq<N>(g, tk, A.q + i * n, idx, flg, nrm);
            NdpQuery Q;
            ndp_make_query(g, idx, nrm, Q);
            NdpHit hit;
            hit.has_coords = 0;
            if (ndp_search_fast<N, MODE>(g, *A.F, mode, Q, hit)) {
                const int v0 = hit.vertex;
                ndp_scan_sentinel(*A.F, mode, hit);
                if (hit.vertex != v0) hit.has_coords = 0;
                ndp_publish(A, Q, t, i, hit);
            } else hard = true;
        }
        const unsigned votes = __ballot_sync(0xffffffffu, hard);
        if (votes) {
            int base = 0;
            const int leader = __ffs(votes) - 1;
            if ((int) lane == leader) base = atomicAdd(&A.ctr->n_hard, __popc(votes));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (hard) A.hardlist[base + __popc(votes & ((1u << lane) - 1u))] = (int) t;
        }
    }
}

// Pass 2 (and the tie pass): shell search over mask rows, then pyramid descent with an explicit
// per-thread stack for whatever the shells cannot prove within their budget.
template <int MODE>
__global__ void __launch_bounds__(NDP_BLOCK, NDP_HARD_OCC) k_search_hard(const __grid_constant__ SearchArgs A)
{
    extern __shared__ double s_ticks[];
    const double *tk = ndp_stage_ticks(A.ticks, A.nticks, A.ticks_in_smem, s_ticks);
    const NdpGeom &g = A.g;
    const int n = g.naxes;
    const int mode = ndp_mode_of(A.method, A.search);
    const long long total = *A.count;
    for (long long s = (long long) blockIdx.x * blockDim.x + threadIdx.x; s < total; s += (long long) gridDim.x * blockDim.x) {
        const long long t = A.sublist[s];
        const long long i = A.list ? A.list[t] : t;
        int idx[NDP_MAX_AXES], flg[NDP_MAX_AXES];
        double nrm[NDP_MAX_AXES];
        if (A.list && A.rec_idx) {
            for (int a = 0; a < n; a++) { idx[a] = A.rec_idx[t * n + a]; nrm[a] = A.rec_nrm[t * n + a]; }
        } else ndp_import_query_seq<0>(g, tk, A.q + i * n, idx, flg, nrm);
        NdpQuery Q;
        ndp_make_query(g, idx, nrm, Q);
        NdpHit hit;
        hit.has_coords = 0;
        ndp_search_descend<MODE>(g, *A.F, *A.R, mode, Q, A.have_topo ? &A.topo : nullptr, hit);
        if (hit.status == NDP_SEARCH_TIE) {
            // provisional (valid) vertex so that k_finish stays in bounds; redone once the topology exists
            A.chosen[t] = hit.vertex;
            A.tielist[atomicAdd(&A.ctr->n_tie, 1)] = (int) t;
            continue;
        }
        if (A.have_topo && A.sublist == A.tielist) atomicAdd(&A.ctr->n_tie_resolved, 1);
        ndp_scan_sentinel(*A.F, mode, hit);
        ndp_publish(A, Q, t, i, hit);
    }
}

// One warp per query: lanes stride over every basic vertex exactly as the
// reference's scan does (:279-324), each keeping its (distance, id) minimum, then
// a shuffle arg-min with the reference's tie rule (:101-112).  For small lattices.
__global__ void __launch_bounds__(NDP_BLOCK) k_search_brute(const __grid_constant__ SearchArgs A)
{
    extern __shared__ double s_ticks[];
    const double *tk = ndp_stage_ticks(A.ticks, A.nticks, A.ticks_in_smem, s_ticks);
    const NdpGeom &g = A.g;
    const int n = g.naxes;
    const int mode = ndp_mode_of(A.method, A.search);
    const unsigned lane = threadIdx.x & 31;
    const long long total = A.count ? (long long) *A.count : A.count_host;
    const long long warps = ((long long) gridDim.x * blockDim.x) >> 5;
    for (long long t = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < total; t += warps) {
        const long long i = A.list ? A.list[t] : t;
        int idx[NDP_MAX_AXES], flg[NDP_MAX_AXES];
        double nrm[NDP_MAX_AXES];
        ndp_import_query<0>(g, tk, A.q + i * n, idx, flg, nrm);
        NdpQuery Q;
        ndp_make_query(g, idx, nrm, Q);

        int best = -1;
        double bestd = 0.0;
        for (int v = (int) lane; v < g.nverts; v += 32) {
            double d;
            if (!ndp_bit(A.F->bits[0], v)) d = 1e10;                     // :283-286
            else {
                int c[NDP_MAX_AXES];
                for (int a = 0; a < g.nbasic; a++) c[a] = (v / g.bstride[a]) % g.len[a];
                d = ndp_metric_vertex(g, mode, Q, c);
            }
            if (best < 0 || d < bestd) { best = v; bestd = d; }         // ascending v per lane: first wins ties
        }
        for (int off = 16; off; off >>= 1) {
            const int ob = __shfl_xor_sync(0xffffffffu, best, off);
            const double od = __shfl_xor_sync(0xffffffffu, bestd, off);
            if (ob >= 0 && (best < 0 || od < bestd || (od == bestd && ob < best))) { best = ob; bestd = od; }
        }
        if (lane == 0) {
            NdpHit hit; hit.vertex = best; hit.dsel = bestd; hit.status = NDP_SEARCH_OK; hit.hard = 0; hit.has_coords = 0;
            ndp_publish(A, Q, t, i, hit);
        }
    }
}

// Sparse masks (a hypercube mask with a handful of fully defined cells among 10^8: SURVEY 8d's i.i.d.-void stress
// table): one warp per query scores the set vertices themselves -- the reference's candidates, with its distance
// expression -- instead of descending a pyramid that is empty almost everywhere.  Full-scan modes take the lowest id
// among equals (:101-112, ids ascend); kd modes settle exact ties by the reference's traversal order (topology) or
// hand them to the tie pass when the topology does not exist yet.
template <int MODE>
__global__ void __launch_bounds__(NDP_BLOCK) k_search_sparse(const __grid_constant__ SearchArgs A)
{
    extern __shared__ double s_ticks[];
    const double *tk = ndp_stage_ticks(A.ticks, A.nticks, A.ticks_in_smem, s_ticks);
    const NdpGeom &g = A.g;
    const int n = g.naxes;
    constexpr bool kd = MODE < NDP_MODE_LS_NEAREST;
    const unsigned lane = threadIdx.x & 31;
    const long long total = A.count ? (long long) *A.count : A.count_host;
    const long long warps = ((long long) gridDim.x * blockDim.x) >> 5;
    for (long long t = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < total; t += warps) {
        const long long i = A.list ? A.list[t] : t;
        int idx[NDP_MAX_AXES], flg[NDP_MAX_AXES];
        double nrm[NDP_MAX_AXES];
        ndp_import_query<0>(g, tk, A.q + i * n, idx, flg, nrm);
        NdpQuery Q;
        ndp_make_query(g, idx, nrm, Q);

        int best = -1;
        double bestd = 0.0;
        bool tie = false;
        for (int j = (int) lane; j < A.sparse_n; j += 32) {
            const int v = A.sparse_ids[j];
            int c[NDP_MAX_AXES];
            ndp_decode_vertex(g, v, c);
            const double d = ndp_metric_vertex(g, MODE, Q, c);
            if (best < 0 || d < bestd) { best = v; bestd = d; tie = false; }
            else if (kd && d == bestd) tie = true;                       // LS: ids ascend per lane, the first keeps the tie
        }
        for (int off = 16; off; off >>= 1) {
            const int ob = __shfl_xor_sync(0xffffffffu, best, off);
            const double od = __shfl_xor_sync(0xffffffffu, bestd, off);
            const bool ot = __shfl_xor_sync(0xffffffffu, (int) tie, off) != 0;
            if (ob >= 0) {
                if (best < 0 || od < bestd) { best = ob; bestd = od; tie = ot; }
                else if (od == bestd) { if (kd) tie = true; else if (ob < best) best = ob; }
            }
        }
        if (lane != 0) continue;
        NdpHit hit;
        hit.vertex = best; hit.dsel = bestd; hit.status = best < 0 ? NDP_SEARCH_EMPTY : NDP_SEARCH_OK; hit.hard = 0; hit.has_coords = 0;
        if (kd && tie && best >= 0) {
            if (!A.have_topo) {
                A.chosen[t] = best;                                      // provisional (valid) vertex; redone by the tie pass
                A.tielist[atomicAdd(&A.ctr->n_tie, 1)] = (int) t;
                continue;
            }
            int first = -1;                                              // rare: one lane walks the tied candidates
            for (int j = 0; j < A.sparse_n; j++) {
                const int v = A.sparse_ids[j];
                int c[NDP_MAX_AXES];
                ndp_decode_vertex(g, v, c);
                if (ndp_metric_vertex(g, MODE, Q, c) == bestd && (first < 0 || ndp_tree_beats(g, MODE, A.topo, Q, v, first))) first = v;
            }
            hit.vertex = first;
            atomicAdd(&A.ctr->n_tie_resolved, 1);
        }
        ndp_scan_sentinel(*A.F, MODE, hit);
        ndp_publish(A, Q, t, i, hit);
    }
}

#endif  // NDP_TU_SEARCH

// ---------------------------------------------------------------------------
// k_finish
// ---------------------------------------------------------------------------
struct FinishArgs {
    NdpGeom g;
    const double *ticks; int nticks; int ticks_in_smem;
    const double *grid; const double *cells;
    const double *q;
    int method;
    const int *list;        // query id per work item
    const int *sublist;     // positions into list to redo (after tie resolution), or nullptr
    const int *count;       // device count of work items (of sublist when given)
    const int *chosen;
    const int *rec_idx; const double *rec_nrm;   // import products by list position, or nullptr
    double *interps;
};

#if defined(NDP_TU_FINISH)
// G lanes share a work item when vdim > 1: lane k owns components k, k+G, ... so that a corner
// vector is one coalesced G*8-byte load (as in k_main).
template <int N, int G, bool CM>
__global__ void __launch_bounds__(NDP_BLOCK, (N >= 6 ? 1 : 2)) k_finish(const __grid_constant__ FinishArgs A)
{
    extern __shared__ double s_ticks[];
    const double *tk = ndp_stage_ticks(A.ticks, A.nticks, A.ticks_in_smem, s_ticks);
    const NdpGeom &g = A.g;
    const int n = N ? N : g.naxes;
    const int V = g.vdim;
    const int lig = threadIdx.x % G;
    const long long total = *A.count;
    const long long groups = ((long long) gridDim.x * blockDim.x) / G;
    for (long long s = ((long long) blockIdx.x * blockDim.x + threadIdx.x) / G; s < total; s += groups) {
        const long long t = A.sublist ? A.sublist[s] : s;
        const long long i = A.list[t];
        int idx[N ? N : NDP_MAX_AXES], flg[N ? N : NDP_MAX_AXES];
        double nrm[N ? N : NDP_MAX_AXES];
        if (A.rec_idx) {
#pragma unroll
            for (int a = 0; a < (N ? N : NDP_MAX_AXES); a++)
                if (a < n) { idx[a] = A.rec_idx[t * n + a]; nrm[a] = A.rec_nrm[t * n + a]; }
        } else ndp_import_query_seq<N>(g, tk, A.q + i * n, idx, flg, nrm);
        int coords[NDP_MAX_AXES];
        const int vertex = A.chosen[t];
        ndp_decode_vertex(g, vertex, coords);
        for (int a = g.nbasic; a < n; a++)
            coords[a] = ndp_assoc_coord((double) idx[a] + nrm[a] - 1.0, g.len[a]);
        for (int k = lig; k < V; k += G)
            A.interps[i * V + k] = ndp_finish_component<N, CM, false>(g, A.grid, A.cells, A.method, idx, nrm, coords, k);
    }
}
#endif  // NDP_TU_FINISH


// ---------------------------------------------------------------------------
// Staged gather (scalar tables, cell-major layout): k_gather_staged
//
// The cell of a query is one contiguous (8 << N)-byte block.  Instead of holding
// 2^N load destinations in registers while DRAM answers, the warp copies the
// blocks of a 32-query tile asynchronously into shared memory (cp.async, LDGSTS)
// and goes on to import the NEXT tile's queries; the reduction of a tile runs one
// iteration later out of shared memory.  Each warp owns a private two-stage ring,
// so there is no block-wide synchronisation in the loop and memory-level
// parallelism no longer depends on register-limited occupancy.
//
// The copy is cooperative: each cell is fetched by 2^(N-1) consecutive lanes with
// one 16-byte cp.async each, so the memory system sees one contiguous request per
// cell.  Slots are padded by 16 bytes so that the 128-bit shared loads of a
// quarter-warp fall into distinct bank groups.
//
// (Round-1 note: the first version issued one cp.async.bulk per lane with an
// mbarrier per warp; ncu showed the bulk copy takes uniform-register operands, so
// the compiler serialises the 32 lanes through an ELECT/R2UR loop -- 25 % of the
// kernel's instructions.  profiles/r01c_*.  Per-thread LDGSTS has no such loop.)
//
// MODE 0: the k_main work (import + classify + gather + reduce + off-grid list)
// MODE 1: the k_finish work for LINEAR extrapolation (cell below the found corner)
// ---------------------------------------------------------------------------
#define NDP_STAGED_WARPS 4

struct StagedArgs {
    NdpGeom g;
    const double *ticks; int nticks;
    const double *cells;
    const double *q; long long n;      // MODE 0: number of queries; MODE 1: unused
    double *interps; double *dists;
    int method;
    int *offlist;                      // MODE 0: out; MODE 1: in (query id per work item)
    const int *chosen;                 // MODE 1: winning vertex per work item
    int *rec_idx; double *rec_nrm;     // import products by list position: MODE 0 writes, MODE 1 reads (nullptr: re-import)
    NdpCounters *ctr;
};

#if defined(NDP_TU_STAGED)
__device__ __forceinline__ unsigned ndp_smem_u32(const void *p) { return (unsigned) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void ndp_cp_async16(unsigned dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void ndp_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void ndp_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory"); }

template <int N>
struct NdpStagedItem {                 // what a lane keeps between issuing a tile and reducing it
    long long i;                       // query id (-1: lane idle)
    int idx[N];                        // MODE 0: superior corner indices (for the off-grid record)
    double x[N];                       // interpolation weights
    unsigned pinmask, selbits;
    bool issued;                       // this lane's cell is being copied
};

// STAGES == 2: each warp overlaps its own copy with its own next import (fewer, fatter warps);
// STAGES == 1: no intra-warp overlap, twice the resident warps hide the latency instead.
template <int N, int MODE, int STAGES, int OCC = 5>
__global__ void __launch_bounds__(NDP_STAGED_WARPS * 32, (N <= 5 ? (STAGES == 1 ? OCC : 3) : 1))
k_gather_staged(const __grid_constant__ StagedArgs A)
{
    constexpr unsigned CELL_BYTES = 8u << N;
    constexpr unsigned SLOT = CELL_BYTES + 16u;
    constexpr int CHUNKS = (int) (CELL_BYTES / 16u);      // 16-byte pieces per cell (N >= 2: at least 2)
    extern __shared__ __align__(128) unsigned char ndp_smem[];
    const NdpGeom &g = A.g;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // shared layout: ticks | slots
    double *s_ticks = reinterpret_cast<double *>(ndp_smem);
    const unsigned ticks_bytes = ((unsigned) A.nticks * 8u + 127u) & ~127u;
    unsigned char *s_slots = ndp_smem + ticks_bytes;

    for (int i = threadIdx.x; i < A.nticks; i += blockDim.x) s_ticks[i] = A.ticks[i];
    __syncthreads();

    // slot of lane L in stage st of this warp
    auto slot_of = [&](int st, unsigned L) -> unsigned char * {
        return s_slots + ((size_t) (warp * STAGES + st) * 32 + L) * SLOT;
    };

    const long long total = (MODE == 0) ? A.n : (long long) A.ctr->n_off;
    const long long ntiles = (total + 31) / 32;
    const long long tile_stride = (long long) gridDim.x * NDP_STAGED_WARPS;
    const long long tile0 = (long long) blockIdx.x * NDP_STAGED_WARPS + warp;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);

    // the global-memory inputs of one work item, loaded one tile ahead so that their latency
    // overlaps the previous tile's copy and reduction
    struct Raw { long long i; double v[N]; int idx[N]; int vertex; };
    const bool have_rec = (MODE == 1) && A.rec_idx != nullptr;
    auto fetch = [&](long long tile, Raw &raw) {
        const long long w = tile * 32 + lane;
        raw.i = -1; raw.vertex = 0;
        if (w >= total) return;
        raw.i = (MODE == 0) ? w : (long long) A.offlist[w];
        if (MODE == 1) raw.vertex = A.chosen[w];
        if (have_rec) {
#pragma unroll
            for (int a = 0; a < N; a++) { raw.idx[a] = A.rec_idx[w * N + a]; raw.v[a] = A.rec_nrm[w * N + a]; }
        } else {
#pragma unroll
            for (int a = 0; a < N; a++) raw.v[a] = A.q[raw.i * N + a];
        }
    };

    // issue: import the lane's query of `tile`, start the cooperative copy into stage st
    auto produce = [&](const Raw &raw, int st, NdpStagedItem<N> &it) {
        it.i = -1; it.issued = false; it.pinmask = 0; it.selbits = 0;
        long long cell = 0;
        if (raw.i >= 0) {
            const long long i = raw.i;
            it.i = i;
            int idx[N], flg[N];
            double nrm[N];
            if (have_rec) {
#pragma unroll
                for (int a = 0; a < N; a++) { idx[a] = raw.idx[a]; nrm[a] = raw.v[a]; }
            } else ndp_import_query<N>(g, s_ticks, raw.v, idx, flg, nrm);
            if (MODE == 0) {
#pragma unroll
                for (int a = 0; a < N; a++) it.idx[a] = idx[a];
                const bool oob = ndp_classify<N>(g, flg, nrm, it.pinmask, it.selbits);
                it.issued = !oob;
                if (g.small) {
                    int c32 = 0;
#pragma unroll
                    for (int a = 0; a < N; a++) { it.x[a] = nrm[a]; c32 += (idx[a] - 1) * g.cstride32[a]; }
                    cell = c32;
                } else {
#pragma unroll
                    for (int a = 0; a < N; a++) { it.x[a] = nrm[a]; cell += (long long) (idx[a] - 1) * g.cstride[a]; }
                }
            } else {
                const int vertex = raw.vertex;
                it.issued = true;
                int vc[NDP_MAX_AXES];
                ndp_decode_vertex(g, vertex, vc);
#pragma unroll
                for (int a = 0; a < N; a++) {
                    const int c = (a < g.nbasic) ? vc[a]
                                                 : ndp_assoc_coord((double) idx[a] + nrm[a] - 1.0, g.len[a]);
                    const int sup = c > 1 ? c : 1;                      // :601
                    it.x[a] = nrm[a] + (double) (idx[a] - c);           // :621-622
                    cell += (long long) (sup - 1) * g.cstride[a];
                }
            }
        }
        // Cooperative copy: a cell is fetched by LPB consecutive lanes, 16 bytes each, i.e. as ONE
        // contiguous (8 << N)-byte request -- measured on B200 (tools/probe_gather.cu), random
        // 256-byte blocks fetched this way stream at the full HBM rate (27 G blocks/s), 5.5x the rate
        // of one thread walking its own block with sixteen 128-bit loads.
        constexpr int LPB = CHUNKS < 32 ? CHUNKS : 32;      // lanes per cell
        constexpr int BPI = 32 / LPB;                       // cells per instruction
        // owners broadcast their cell number (one 32-bit shuffle when the table is small enough)
        const unsigned sub = lane % LPB, grp = lane / LPB;
        const unsigned dst0 = ndp_smem_u32(slot_of(st, grp)) + sub * 16u;
        const unsigned char *src0 = reinterpret_cast<const unsigned char *>(A.cells) + sub * 16u;
        if (g.small) {
            const int mine = it.issued ? (int) cell : -1;
#pragma unroll
            for (int k = 0; k < 32 / BPI; k++) {
                const int oc = __shfl_sync(0xffffffffu, mine, k * BPI + (int) grp);
                if (oc >= 0) {
#pragma unroll
                    for (int c = 0; c < CHUNKS / LPB; c++)      // cells wider than 512 bytes take several sweeps
                        ndp_cp_async16(dst0 + (unsigned) (k * BPI) * SLOT + (unsigned) (c * LPB) * 16u,
                                       src0 + (long long) oc * (long long) CELL_BYTES + (long long) (c * LPB) * 16);
                }
            }
        } else {
            const long long mine = it.issued ? cell : -1;
#pragma unroll
            for (int k = 0; k < 32 / BPI; k++) {
                const long long oc = __shfl_sync(0xffffffffu, mine, k * BPI + (int) grp);
                if (oc >= 0) {
#pragma unroll
                    for (int c = 0; c < CHUNKS / LPB; c++)
                        ndp_cp_async16(dst0 + (unsigned) (k * BPI) * SLOT + (unsigned) (c * LPB) * 16u,
                                       src0 + oc * (long long) CELL_BYTES + (long long) (c * LPB) * 16);
                }
            }
        }
        ndp_cp_async_commit();
    };

    // reduce: the tile's copies have landed; interpolate out of shared memory, write results
    auto consume = [&](int st, const NdpStagedItem<N> &it) {
        bool need_search = false;
        if (it.i >= 0) {
            bool fdhc = false;
            if (it.issued) {
                const double2 *b2 = reinterpret_cast<const double2 *>(slot_of(st, lane));
                bool bad = false;
                const double val = ndp_cell_reduce_halves<N>(b2, it.x, it.pinmask, it.selbits, bad);
                fdhc = !bad;
                if (MODE == 1 || fdhc) A.interps[it.i] = val;
            }
            if (MODE == 0) {
                if (fdhc) A.dists[it.i] = 0.0;
                else if (A.method == NDP_M_NONE) { A.interps[it.i] = qnan; A.dists[it.i] = 0.0; }
                else need_search = true;
            }
        }
        if (MODE == 0) {
            const unsigned votes = __ballot_sync(0xffffffffu, need_search);
            if (votes) {
                int base = 0;
                const int leader = __ffs(votes) - 1;
                if ((int) lane == leader) base = atomicAdd(&A.ctr->n_off, __popc(votes));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (need_search) {
                    const int pos = base + __popc(votes & ((1u << lane) - 1u));
                    A.offlist[pos] = (int) it.i;
                    if (A.rec_idx) {
#pragma unroll
                        for (int a = 0; a < N; a++) { A.rec_idx[(long long) pos * N + a] = it.idx[a]; A.rec_nrm[(long long) pos * N + a] = it.x[a]; }
                    }
                }
            }
        }
    };

    if (STAGES == 1) {
        NdpStagedItem<N> item;
        Raw raw;
        if (tile0 < ntiles) fetch(tile0, raw);
        for (long long tile = tile0; tile < ntiles; tile += tile_stride) {
            produce(raw, 0, item);
            if (tile + tile_stride < ntiles) fetch(tile + tile_stride, raw);   // in flight during the wait below
            ndp_cp_async_wait<0>();
            __syncwarp();                      // the other lanes' copies into my slot are visible
            consume(0, item);
            __syncwarp();                      // the slots may be overwritten by the next tile's copies
        }
        return;
    }
    NdpStagedItem<N> item[2];
    Raw raw;
    if (tile0 < ntiles) { fetch(tile0, raw); produce(raw, 0, item[0]); }
    int st = 0;
    for (long long tile = tile0; tile < ntiles; tile += tile_stride) {
        const long long next = tile + tile_stride;
        // one commit per iteration (possibly empty) keeps the group count uniform: everything
        // but the newest group -- the tile being consumed included -- is complete after wait<1>
        if (st == 0) {
            if (next < ntiles) { fetch(next, raw); produce(raw, STAGES - 1, item[1]); } else ndp_cp_async_commit();
            ndp_cp_async_wait<1>();
            __syncwarp();
            consume(0, item[0]);
        } else {
            if (next < ntiles) { fetch(next, raw); produce(raw, 0, item[0]); } else ndp_cp_async_commit();
            ndp_cp_async_wait<1>();
            __syncwarp();
            consume(STAGES - 1, item[1]);
        }
        __syncwarp();
        st ^= 1;
    }
}
#endif  // NDP_TU_STAGED
