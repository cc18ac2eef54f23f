// ndp_fused.cuh -- the per-query PLAN of the fused ndpolate kernel (k_fused): after the
// import, everything the reference's driver decides for one query (ndpolate,
// src/ndpolator.c:510-671) -- is the enclosing hypercube fully defined (:436-505), and if not,
// which vertex / hypercube the extrapolation uses (:556-635) -- is settled here from two
// register-time products, the hypercube-mask slice and the nearest-site map (ndp_nsm.cuh),
// without touching the grid.  What is left for the kernel is ONE gather per query (the
// query's own cell, or the cell below the found hypercube) and its reduction, so that
// in-grid and extrapolated queries share one memory pipeline.  Queries the plan cannot
// settle exactly (see ndp_nsm.cuh) are marked SLOW and redone by the exact generic path.
#pragma once

#include "ndp_nsm.cuh"

#define NDP_PLAN_IDLE 0
#define NDP_PLAN_GATHER 1     // interps = reduce(cell `cell`, weights x, pin/sel); dists = dist
#define NDP_PLAN_VALUE 2      // interps = grid[vflat] (NEAREST extrapolation, :556-567); dists = dist
#define NDP_PLAN_NAN 3        // interps = NaN, dists = 0 (extrapolation NONE, :551-554)
#define NDP_PLAN_SLOW 4       // to the slow list

// The slice of the hypercube mask over the axes it depends on: bit set <=> every corner of the
// cell below that superior corner is defined (ndp_types.c:201-258).
struct NdpHcSlice {
    int astride[NDP_MAX_AXES];     // per basic axis: its stride in the slice, 0 when the mask does not depend on it
    const uint32_t *bits;
    int nwords;
};

// Blocked cell layout.  With block order K (1 <= K <= n) the corner values of a cell are stored as 2^(n-K)
// BLOCKS of 2^K doubles: a block holds the corners that differ on the LAST K axes, and the blocks of one cell
// differ on the first n-K ("outer") axes.  Blocks are indexed by vertex coordinates on the outer axes and cell
// coordinates on the inner ones, so neighbouring cells share their outer-axis blocks: the copy costs 2^K times
// the grid instead of 2^n (K = n: one block per cell = the cell-major copy).  The image of a cell in shared
// memory is the same for every K -- corner j at byte 8 j, axis 0 = most significant bit -- so the reduction
// does not change; only the source address of each 16-byte piece does:
//   address(piece s of the cell below `sup`) = blocks + cell_of(sup) * (8 << K) + ndp_block_piece_offset(s),
// cell_of() taken with the block strides.
struct NdpBlocks {
    int order;                          // K
    long long stride[NDP_MAX_AXES];     // blocks per +1 on axis a (vertex coordinate on outer, cell coordinate on inner axes)
    long long nblocks;
};

inline void ndp_blocks_setup(const NdpGeom &g, int order, NdpBlocks &B)
{
    const int n = g.naxes;
    B.order = order;
    long long st = 1;
    for (int a = n - 1; a >= 0; a--) {
        B.stride[a] = st;
        st *= (a >= n - order) ? g.len[a] - 1 : g.len[a];
    }
    B.nblocks = st;
}

// byte offset, relative to the cell's first block, of the 16-byte piece s (corners 2 s and 2 s + 1) of a cell
NDP_HD long long ndp_block_piece_offset(int n, int order, const long long *bstride, int s)
{
    const int j = 2 * s;                                   // corner number of the piece's first double
    const int outer = j >> order;                          // bits of the outer axes, axis 0 = most significant
    long long blocks = 0;
    for (int a = 0; a < n - order; a++)
        if ((outer >> (n - order - 1 - a)) & 1) blocks += bstride[a];
    return blocks * ((long long) 8 << order) + (long long) (j & ((1 << order) - 1)) * 8;
}

template <int N>
struct NdpPlan {
    int kind;
    long long cell;            // GATHER: cell number in the cell-major copy
    long long vflat;           // VALUE: vertex number in the grid (all axes)
    double x[N ? N : NDP_MAX_AXES];
    unsigned pin, sel;
    double dist;
    int searched;              // 1: the query went to the nearest search (map or slow path)
};

template <int N>
NDP_HD long long ndp_cell_of(const NdpGeom &g, const int *sup)
{
    constexpr int MM = N ? N : NDP_MAX_AXES;
    const int n = N ? N : g.naxes;
    if (g.small) {
        int c32 = 0;
#pragma unroll
        for (int a = 0; a < MM; a++) if (a < n) c32 += (sup[a] - 1) * g.cstride32[a];
        return c32;
    }
    long long c = 0;
#pragma unroll
    for (int a = 0; a < MM; a++) if (a < n) c += (long long) (sup[a] - 1) * g.cstride[a];
    return c;
}

// METHOD / MODE: compile-time extrapolation method and search mode (MODE is ignored for NONE).
// `hcbits` is H.bits or a shared-memory copy of it.
template <int N, int METHOD, int MODE, int VS = 0>
NDP_HD void ndp_fused_plan(const NdpGeom &g, const NdpHcSlice &H, const uint32_t *hcbits, const NdpSiteMap &M,
                           const int *idx, const int *flg, const double *nrm, NdpPlan<N> &P)
{
    constexpr int MM = N ? N : NDP_MAX_AXES;
    const int n = N ? N : g.naxes;
    // the head of the query's site-map region is requested first, for every query: its L2 round trip then runs
    // beside everything up to the candidate loop (and costs in-grid lanes nothing: they would idle meanwhile)
    NdpQuery Q;
    NdpNsmProbe pr;
    if (METHOD != NDP_M_NONE) {
        ndp_make_query_n<N>(g, idx, nrm, Q);
        ndp_nsm_probe<N, MODE, VS>(g, M, Q, pr);
    }
    const bool oob = ndp_classify<N>(g, flg, nrm, P.pin, P.sel);
    P.dist = 0.0;
    P.searched = 0;
    if (!oob) {
        int h = 0;
#pragma unroll
        for (int a = 0; a < MM; a++) if (a < g.nbasic) h += idx[a] * H.astride[a];
        if (ndp_bit(hcbits, h)) {
            // every corner of the query's cell is defined: the hypercube is fully defined whatever the
            // on-vertex reduction keeps (:455-505)
            P.kind = NDP_PLAN_GATHER;
            P.cell = ndp_cell_of<N>(g, idx);
#pragma unroll
            for (int a = 0; a < MM; a++) if (a < n) P.x[a] = nrm[a];
            return;
        }
        // a corner is missing.  With on-vertex axes the reduced hypercube may avoid it: the generic path looks
        if (P.pin != 0) { P.kind = NDP_PLAN_SLOW; return; }
    }
    if (METHOD == NDP_M_NONE) { P.kind = NDP_PLAN_NAN; return; }
    P.searched = 1;
    NdpHit hit;
    if (!ndp_nsm_search<N, MODE, VS>(g, M, Q, pr, hit)) { P.kind = NDP_PLAN_SLOW; return; }
    int coords[MM];
#pragma unroll
    for (int a = 0; a < MM; a++) if (a < n) coords[a] = a < g.nbasic ? hit.coords[a] : Q.acoord[a];
    P.dist = (MODE >= NDP_MODE_LS_NEAREST) ? hit.dsel : ndp_reported_dist_n<N, METHOD>(g, Q, coords);
    if (METHOD == NDP_M_NEAREST) {
        long long v = 0;
#pragma unroll
        for (int a = 0; a < MM; a++) if (a < n) v += (long long) coords[a] * g.vstride[a];
        P.kind = NDP_PLAN_VALUE;
        P.vflat = v;
        return;
    }
    // LINEAR: full-dimensional cell below the found corner, clamped at 0 (:601); query shifted into that
    // cell's unit coordinates (:621-622)
    int sup[MM];
#pragma unroll
    for (int a = 0; a < MM; a++)
        if (a < n) {
            sup[a] = coords[a] > 1 ? coords[a] : 1;
            P.x[a] = nrm[a] + (double) (idx[a] - coords[a]);
        }
    P.kind = NDP_PLAN_GATHER;
    P.cell = ndp_cell_of<N>(g, sup);
    P.pin = 0; P.sel = 0;
}
