// ndp_core.cuh -- per-query arithmetic of the ndpolate hot path, written once as
// __host__ __device__ inline functions.  The CUDA kernels in ndp_kernels.cuh are
// the product; tests/hostsim/ compiles the very same functions for the host so
// that the logic can be diffed against the oracle in the CPU-only container
// before any GPU time is spent.  (That host build is test scaffolding: it is not
// part of libndpb200.so and the product never falls back to it.)
//
// Numerics contract (BASELINE.json north_star): fp64 everywhere, compiled with
// -fmad=false so that every a+b*c keeps the two roundings of the reference's
// x86-64 build; sums run in the reference's order.  Citations are to the
// reference tree (src/ndpolator.c unless another file is named).
#pragma once

#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define NDP_HD __host__ __device__ __forceinline__
#define NDP_HD_CALL static __host__ __device__ __noinline__      // big, rarely-hot helpers: keep the kernels' code small
#else
#define NDP_HD_CALL inline
#define NDP_HD inline
struct double2 { double x, y; };
#endif

#define NDP_MAX_AXES 12
#define NDP_MAX_LEVELS 18      // occupancy pyramid: ceil(log2(65536)) + 2

// extrapolation methods / search algorithms / flags: src/ndpolator.h:28-63
#define NDP_M_NONE 0
#define NDP_M_NEAREST 1
#define NDP_M_LINEAR 2
#define NDP_S_KDTREE 0
#define NDP_S_LINEAR 1
#define NDP_F_ON_VERTEX 1
#define NDP_F_OOB 2

// search "mode" = which metric picks the winner and how ties are broken
#define NDP_MODE_KD_NEAREST 0   // kd_nearest over vmask vertices           (:223-254, kdtree.c:345-457)
#define NDP_MODE_KD_LINEAR 1    // kd_nearest over hcmask cell centres       (:126-127, :255-258)
#define NDP_MODE_LS_NEAREST 2   // full scan, all-axes vertex distance       (:279-329, :300-309)
#define NDP_MODE_LS_LINEAR 3    // full scan, ndp_distance to the cell       (:310-313)

NDP_HD int ndp_mode_of(int method, int search)
{
    return (search == NDP_S_KDTREE ? 0 : 2) + (method == NDP_M_LINEAR ? 1 : 0);
}

// Table geometry, passed to kernels by value.
struct NdpGeom {
    int naxes, nbasic, vdim;
    int len[NDP_MAX_AXES];            // ticks per axis
    int tickoff[NDP_MAX_AXES];        // start of axis a in the concatenated tick array
    long long vstride[NDP_MAX_AXES];  // vertices per +1 on axis a, all axes (ndp_types.c:59-65 `cplen`)
    long long cstride[NDP_MAX_AXES];  // cells per +1 on axis a in the (len-1)^n cell lattice
    int bstride[NDP_MAX_AXES];        // same as vstride but within the basic-only lattice
    int nverts;                       // basic lattice size (ndp_types.c:174-176); < 2^31 like the reference's int
    long long ncells;
    double inv_step[NDP_MAX_AXES];    // (len-1)/(last-first): index guess for the import; 0 = always bisect
    double first[NDP_MAX_AXES];       // tick[0] and tick[len-1] of every axis (constant-bank copies)
    double last[NDP_MAX_AXES];
    int vstride32[NDP_MAX_AXES];      // 32-bit copies of vstride / cstride, valid when small != 0
    int cstride32[NDP_MAX_AXES];
    int small;                        // every vertex and cell index fits in 31 bits
    double inv_bstride[NDP_MAX_AXES]; // 1.0 / bstride: first guess of the vertex-id decode (then fixed up exactly)
    int easy;                         // every axis is evenly spaced by an exact power of two (see ndp_geom_axes)
    double inv_pow2[NDP_MAX_AXES];    // 1 / spacing of such an axis (exact)
};

// Per-axis facts the import uses, from the concatenated tick array (host side, at table creation):
//   inv_step / first / last  index guess of ndp_import_axis_guess (0 = always bisect);
//   easy                     every axis has tick[i+1] - tick[i] == s_a for all i IN FLOATING POINT, with s_a an
//                            exact power of two.  Then the quotient (x - t[l-1]) / (t[l] - t[l-1]) of :37 / :381
//                            equals (x - t[l-1]) * (1 / s_a) bit for bit (scaling by a power of two is exact,
//                            under- and overflow included), and the index guess needs no corrective step.
inline void ndp_geom_axes(NdpGeom &g, const double *ticks)
{
    g.easy = 1;
    for (int a = 0; a < g.naxes; a++) {
        const double *tk = ticks + g.tickoff[a];
        const int len = g.len[a];
        bool ascending = true;
        for (int i = 1; i < len; i++) if (!(tk[i] > tk[i - 1])) ascending = false;
        const double span = tk[len - 1] - tk[0];
        g.inv_step[a] = (ascending && span > 0 && span < 1e300) ? (double) (len - 1) / span : 0.0;
        g.first[a] = tk[0]; g.last[a] = tk[len - 1];
        const double s = tk[1] - tk[0];
        int e = 0;
        bool pow2 = ascending && s > 0 && s < 1e300 && frexp(s, &e) == 0.5 && e > -1000 && e < 1000;
        for (int i = 1; i < len && pow2; i++) if (tk[i] - tk[i - 1] != s) pow2 = false;
        g.inv_pow2[a] = pow2 ? 1.0 / s : 0.0;
        if (!pow2) g.easy = 0;
    }
}



// Bit-packed occupancy pyramid over the basic lattice.  Level 0 = one bit per
// basic vertex (vmask or hcmask); level L = one bit per 2^L-wide block, set when
// any vertex inside is set; the last level is a single block.
struct NdpPyramid {
    int nlev;
    int nvar;                         // axes this pyramid spans (all basic axes, or only the mask-variant ones)
    int var_axis[NDP_MAX_AXES];       // pyramid axis j -> basic axis
    int dim[NDP_MAX_LEVELS][NDP_MAX_AXES];      // indexed by pyramid axis j
    int stride[NDP_MAX_LEVELS][NDP_MAX_AXES];
    const uint32_t *bits[NDP_MAX_LEVELS];
    // per level-0 lattice point, along the LAST pyramid axis: nearest set coordinate at or below /
    // at or above it (-1: none); nullptr when not built (axis longer than 32767 ticks)
    const short *prevv;
    const short *nextv;
    int first_clear;                  // lowest vertex id whose level-0 bit is clear, or -1 (full pyramid only)
    int any_set;                      // 0 when the mask is empty
};

// Shape of the reference's insertion-ordered kd-tree (ndp_kdtree_create, :114-136),
// per basic vertex id: parent vertex (-1 root, -2 absent) and depth.
struct NdpTopo {
    const int *parent;
    const int *depth;
};

NDP_HD bool ndp_bit(const uint32_t *bits, int i)
{
    return (bits[i >> 5] >> (i & 31)) & 1u;
}

// Basic vertex id -> coordinates (src/ndpolator.c:235-237), without integer division:
// the quotient is guessed in floating point and corrected by at most one, which is exact.
NDP_HD void ndp_decode_vertex(const NdpGeom &g, int vertex, int *coords)
{
    int rem = vertex;
    for (int a = 0; a < g.nbasic; a++) {
        const int d = g.bstride[a];
        int q = (int) ((double) rem * g.inv_bstride[a]);
        int r = rem - q * d;
        if (r < 0) { q--; r += d; } else if (r >= d) { q++; r -= d; }
        coords[a] = q;
        rem = r;
    }
}

// ---------------------------------------------------------------------------
// A1: per-axis index search (find_first_geq_than, :21-45; normalisation :379-381)
// ---------------------------------------------------------------------------
NDP_HD void ndp_import_axis(const double *tick, int len, double x, int &index, int &flag, double &normed)
{
    const double rtol = 1e-6;                   // :358
    int lo = 1, hi = len - 1;
    while (lo != hi) {                          // :24-33 -- note the absolute "+ rtol"
        int mid = lo + ((hi - lo) >> 1);
        if (x + rtol > tick[mid]) lo = mid + 1; else hi = mid;
    }
    int f = (x < tick[0] || x > tick[len - 1]) ? NDP_F_OOB : 0;    // :35
    double t = (x - tick[lo - 1]) / (tick[lo] - tick[lo - 1]);     // :37 and :381 compute the same quotient
    if (fabs(t) < rtol || (lo == len - 1 && fabs(t - 1) < rtol)) f |= NDP_F_ON_VERTEX;   // :37-39
    index = lo; flag = f; normed = t;
}

// Same result as ndp_import_axis for an ascending axis, without the dependent
// chain of the bisection: guess the index from the mean spacing, then verify the
// two conditions that characterise the bisection's answer on a sorted axis
// (smallest l in [1, len-1] with tick[l] >= x + rtol, else len-1) and walk at most
// two ticks; anything else falls back to the bisection itself.  inv_step == 0
// (axis not ascending, or degenerate) always bisects.
NDP_HD void ndp_import_axis_guess(const double *tick, int len, double inv_step, double first, double last,
                                  double x, int &index, int &flag, double &normed)
{
    const double rtol = 1e-6;
    if (inv_step > 0.0) {
        const double xr = x + rtol;
        // truncation instead of floor only differs below the first tick, where the clamp decides anyway;
        // a NaN or huge quotient converts to some integer, is clamped, and then fails or passes the
        // verification below like any other guess
        int l = (int) ((xr - first) * inv_step);
        l = l < len - 2 ? l : len - 2;
        l = l > 0 ? l + 1 : 1;
        double tl = tick[l], tp = tick[l - 1];
        // one corrective tick either way (the guess from the mean spacing is at most one off on near-uniform axes)
        const bool up = (xr > tl) && (l < len - 1);
        const bool dn = !up && !(xr > tp) && (l > 1);
        if (up || dn) { l += up ? 1 : -1; tl = tick[l]; tp = tick[l - 1]; }
        const bool ok = (!(xr > tl) || l == len - 1) && (l == 1 || xr > tp);
        if (ok) {
            int f = (x < first || x > last) ? NDP_F_OOB : 0;
            const double t = (x - tp) / (tl - tp);
            if (fabs(t) < rtol || (l == len - 1 && fabs(t - 1) < rtol)) f |= NDP_F_ON_VERTEX;
            index = l; flag = f; normed = t;
            return;
        }
    }
    ndp_import_axis(tick, len, x, index, flag, normed);
}

// One lerp of c_ndpolate (:91): three roundings; -fmad=false keeps them apart.
NDP_HD double ndp_lerp(double lo, double hi, double x)
{
    double diff = hi - lo;
    double step = diff * x;
    return lo + step;
}

// ---------------------------------------------------------------------------
// A2 + A3 on one cell, one component.
//
// Corner j of the cell whose superior corner is `sup` has bit (n-1-a) <-> axis a
// (axis 0 = most significant, :480 / :601).  Pinned axes (ON_VERTEX, :474-478)
// select one side instead of interpolating; `sel` says which.  Corners that the
// reference would not have gathered (the other side of a pinned axis) neither
// count for the NaN test (:491) nor influence the result.
//
// LOAD(fv) fills fv[0 .. 2^N) with the corner values.  N > 0: axes known at compile
// time, the whole cell lives in registers.  Returns the reduced value; nan_used
// reports whether any *used* corner was NaN.
// ---------------------------------------------------------------------------
template <int N, class Load>
NDP_HD double ndp_cell_reduce(Load load, const double *x, unsigned pinmask, unsigned selbits, bool &nan_used)
{
    double fv[1 << N];
    load(fv);

    if (pinmask == 0) {
        // the common case: every axis interpolates (constant trip counts so that fv[] stays in registers)
#pragma unroll
        for (int a = 0; a < N; a++) {
            const int h = 1 << (N - 1 - a);
            const double xa = x[a];
#pragma unroll
            for (int j = 0; j < (N ? (1 << (N - 1)) : 1); j++)
                if (j < h) fv[j] = ndp_lerp(fv[j], fv[j + h], xa);
        }
    } else {
#pragma unroll
        for (int a = 0; a < N; a++) {
            const int h = 1 << (N - 1 - a);
            const bool pinned = (pinmask >> (N - 1 - a)) & 1u;
            const bool upper = (selbits >> (N - 1 - a)) & 1u;
            const double xa = x[a];
#pragma unroll
            for (int j = 0; j < (N ? (1 << (N - 1)) : 1); j++) {
                if (j < h) {
                    double l = ndp_lerp(fv[j], fv[j + h], xa);
                    double s = upper ? fv[j + h] : fv[j];
                    fv[j] = pinned ? s : l;
                }
            }
        }
    }
    const double result = fv[0];

    // A NaN in any used corner reaches the result through every lerp and select, so a
    // non-NaN result proves all used corners defined; only a NaN result needs the
    // corner-by-corner test of :491 (it may also stem from a NaN query coordinate).
    bool bad = false;
    if (result != result) {
        load(fv);
        if (pinmask == 0) {
#pragma unroll
            for (int j = 0; j < (1 << N); j++) bad |= (fv[j] != fv[j]);
        } else {
#pragma unroll
            for (int j = 0; j < (1 << N); j++) {
                bool used = (((unsigned) j ^ selbits) & pinmask) == 0;
                bad |= used && (fv[j] != fv[j]);
            }
        }
    }
    nan_used = bad;
    return result;
}

// The same reduction for a scalar cell stored contiguously (cell-major layout), evaluated in two
// halves to halve the live registers: corners are split on the second-to-last axis (the 16-byte
// chunks of the cell alternate between the halves), each half is reduced over axes 0..N-3 down to
// its two last-axis values, then the halves meet on axis N-2 and finally on axis N-1.  Every lerp
// pairs exactly the operands c_ndpolate (:83-96) pairs, in an order that respects all data
// dependences, so the result is bit-identical to ndp_cell_reduce.
template <int N>
NDP_HD double ndp_cell_reduce_halves(const double2 *b2, const double *x, unsigned pinmask, unsigned selbits, bool &nan_used)
{
    static_assert(N >= 2, "needs at least two axes");
    constexpr int HC = 1 << (N - 2);            // 16-byte chunks per half
    double res[2][2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        double fv[2 * HC];                      // index bits: axes 0..N-3 (MSB first), then the last axis
#pragma unroll
        for (int mp = 0; mp < HC; mp++) { const double2 p = b2[2 * mp + h]; fv[2 * mp] = p.x; fv[2 * mp + 1] = p.y; }
#pragma unroll
        for (int a = 0; a < N - 2; a++) {
            const int hb = 1 << (N - 2 - a);
            const bool pinned = (pinmask >> (N - 1 - a)) & 1u;
            const bool upper = (selbits >> (N - 1 - a)) & 1u;
            const double xa = x[a];
#pragma unroll
            for (int j = 0; j < HC; j++)        // constant trip count keeps fv[] in registers
                if (j < hb) {
                    const double l = ndp_lerp(fv[j], fv[j + hb], xa);
                    const double sel = upper ? fv[j + hb] : fv[j];
                    fv[j] = pinned ? sel : l;
                }
        }
        res[h][0] = fv[0]; res[h][1] = fv[1];
    }
    double r[2];
    {
        const bool pinned = (pinmask >> 1) & 1u, upper = (selbits >> 1) & 1u;
#pragma unroll
        for (int b = 0; b < 2; b++) {
            const double l = ndp_lerp(res[0][b], res[1][b], x[N - 2]);
            const double sel = upper ? res[1][b] : res[0][b];
            r[b] = pinned ? sel : l;
        }
    }
    double result;
    {
        const bool pinned = pinmask & 1u, upper = selbits & 1u;
        const double l = ndp_lerp(r[0], r[1], x[N - 1]);
        const double sel = upper ? r[1] : r[0];
        result = pinned ? sel : l;
    }
    bool bad = false;
    if (result != result) {                     // see ndp_cell_reduce: only a NaN result needs the corner test
#pragma unroll
        for (int m = 0; m < 2 * HC; m++) {
            const double2 p = b2[m];
            const bool u0 = (((unsigned) (2 * m) ^ selbits) & pinmask) == 0;
            const bool u1 = (((unsigned) (2 * m + 1) ^ selbits) & pinmask) == 0;
            bad |= (u0 && (p.x != p.x)) || (u1 && (p.y != p.y));
        }
    }
    nan_used = bad;
    return result;
}

// Runtime-n version (any n <= NDP_MAX_AXES): walks only the used corners in
// "axis 0 fastest" order and folds them with a carry chain, which performs
// exactly the lerps of c_ndpolate (:83-96) in an order that respects every data
// dependence, with O(n) state.  free_ax[0..d) lists the non-pinned axes in
// ascending order; FETCHV(bits) gets the corner whose free axis i is at the
// upper side iff bit i of `bits` is set.
template <class FetchV>
NDP_HD double ndp_cell_reduce_rt(int d, FetchV fetchv, const double *xfree, bool &nan_used)
{
    double acc[NDP_MAX_AXES];
    bool bad = false;
    double val = 0.0;
    for (unsigned c = 0; c < (1u << d); c++) {
        val = fetchv(c);
        bad |= (val != val);
        int level = 0;
        while ((c >> level) & 1u) {      // partner at this level is complete: combine (lower, upper)
            val = ndp_lerp(acc[level], val, xfree[level]);
            level++;
        }
        if (level < d) acc[level] = val;
    }
    nan_used = bad;
    return val;
}

// ---------------------------------------------------------------------------
// A5/A6: selection metric of the nearest search
// ---------------------------------------------------------------------------

// max(0, min(len-1, round(qc))) of :241, evaluated in double like the C macros.
NDP_HD int ndp_assoc_coord(double qc, int len)
{
    double r = round(qc);
    double hi = (double) (len - 1);
    double m = hi < r ? hi : r;
    double c = 0.0 > m ? 0.0 : m;
    return (int) c;
}

// Per-axis contribution of a block of lattice coordinates [lo, hi] on a basic
// axis.  For lo == hi this is the exact per-axis term of the reference metric;
// for lo < hi it is the smallest term any coordinate inside can produce,
// evaluated with the same roundings, so block sums never exceed vertex sums
// (fl(a+b) is monotone in both arguments).
NDP_HD double ndp_axis_term(int mode, double q, int lo, int hi)
{
    double d;
    if (mode == NDP_MODE_KD_LINEAR) {            // cell centres c-0.5 (:126-127); kdtree.c:379-381, :745-759
        double plo = (double) lo - 0.5, phi = (double) hi - 0.5;
        if (q < plo) d = plo - q; else if (q > phi) d = phi - q; else d = 0.0;
    } else if (mode == NDP_MODE_LS_LINEAR) {     // ndp_distance, :158-176: cell [c-1, c]
        double clo = (double) lo - 1.0, chi = (double) hi;
        if (q < clo) d = clo - q; else if (q > chi) d = q - chi; else d = 0.0;
    } else {                                     // vertex at integer position (:245-254, :300-309; kdtree.c:379-381)
        double plo = (double) lo, phi = (double) hi;
        if (q < plo) d = plo - q; else if (q > phi) d = phi - q; else d = 0.0;
    }
    return d * d;
}

struct NdpQuery {                 // one imported query in index space
    double qc[NDP_MAX_AXES];      // idx + normed - 1.0 (:226, :250)
    int acoord[NDP_MAX_AXES];     // rounded/clamped tick of associated axes (:240-242)
};

// Sum of the associated-axis terms that the full-scan metrics append (:179-183, :304-308).
NDP_HD double ndp_add_assoc(const NdpGeom &g, int mode, const NdpQuery &Q, double acc)
{
    if (mode >= NDP_MODE_LS_NEAREST)
        for (int a = g.nbasic; a < g.naxes; a++) {
            double diff = Q.qc[a] - (double) Q.acoord[a];
            acc += diff * diff;
        }
    return acc;
}

NDP_HD double ndp_metric_vertex(const NdpGeom &g, int mode, const NdpQuery &Q, const int *c)
{
    double acc = 0.0;
    for (int a = 0; a < g.nbasic; a++) acc += ndp_axis_term(mode, Q.qc[a], c[a], c[a]);
    return ndp_add_assoc(g, mode, Q, acc);
}

// Reported squared distance once the winner is known (:245-258).
NDP_HD double ndp_reported_dist(const NdpGeom &g, int method, const NdpQuery &Q, const int *coords)
{
    double acc = 0.0;
    if (method == NDP_M_NEAREST) {
        for (int a = 0; a < g.naxes; a++) {
            double diff = Q.qc[a] - (double) coords[a];
            acc += diff * diff;
        }
    } else {
        for (int a = 0; a < g.nbasic; a++) acc += ndp_axis_term(NDP_MODE_LS_LINEAR, Q.qc[a], coords[a], coords[a]);
        for (int a = g.nbasic; a < g.naxes; a++) {
            double diff = Q.qc[a] - (double) coords[a];
            acc += diff * diff;
        }
    }
    return acc;
}

// ---------------------------------------------------------------------------
// Tie resolution for the kd modes.
//
// kd_nearest (kdtree.c:404-457) seeds the best guess with the root and replaces
// it only on a strictly smaller distance (:383); kd_nearest_i visits, at every
// node, the nearer child's subtree, then the node, then (if not pruned) the
// farther child's subtree.  A subtree holding a vertex at the minimum distance
// cannot be pruned before that distance has been reached, so among several
// vertices at exactly the minimum distance the reference returns the one that
// comes first in that traversal order, except that the root beats everything.
// For two candidates the order is decided at their lowest common ancestor.
// ---------------------------------------------------------------------------
NDP_HD double ndp_tree_pos(const NdpGeom &g, int mode, int v, int ax)
{
    double p = (double) ((v / g.bstride[ax]) % g.len[ax]);
    return (mode == NDP_MODE_KD_LINEAR) ? p - 0.5 : p;
}

// true when vertex A precedes vertex B (A != B) in the reference's traversal for query Q
NDP_HD_CALL bool ndp_tree_beats(const NdpGeom &g, int mode, const NdpTopo &T, const NdpQuery &Q, int A, int B)
{
    int a = A, b = B, ca = -1, cb = -1;          // ca/cb: child of the meeting node on A's / B's side
    int da = T.depth[a], db = T.depth[b];
    while (da > db) { ca = a; a = T.parent[a]; da--; }
    while (db > da) { cb = b; b = T.parent[b]; db--; }
    while (a != b) { ca = a; a = T.parent[a]; cb = b; b = T.parent[b]; da--; }
    const int C = a;                             // lowest common ancestor, depth da
    const int ax = da % g.nbasic;                // split axis cycles with depth (kdtree.c:189)
    const double cpos = ndp_tree_pos(g, mode, C, ax);
    const bool near_is_lo = (Q.qc[ax] - cpos) <= 0;          // kdtree.c:354-355
    if (C == A) {                                // A is an ancestor of B
        if (da == 0) return true;                // the root is the initial guess
        bool b_is_lo = ndp_tree_pos(g, mode, cb, ax) < cpos; // kdtree.c:190
        return !(b_is_lo == near_is_lo);         // B wins only from A's nearer subtree
    }
    if (C == B) {
        if (da == 0) return false;
        bool a_is_lo = ndp_tree_pos(g, mode, ca, ax) < cpos;
        return a_is_lo == near_is_lo;
    }
    bool a_is_lo = ndp_tree_pos(g, mode, ca, ax) < cpos;
    return a_is_lo == near_is_lo;
}

// ---------------------------------------------------------------------------
// Exact nearest search on the lattice.
//
// Every candidate of the reference's search sits on the integer basic lattice,
// so instead of walking its (unbalanced, 100s of levels deep) pointer tree we
//  (1) try the lattice point the query rounds to and prove it the unique
//      minimiser by checking its 2*nbasic axis neighbours (per-axis convexity +
//      monotone rounding make that sufficient), and otherwise
//  (2) run a branch-and-bound descent over the occupancy pyramid with an
//      explicit per-thread stack, nearest child first, pruning blocks whose
//      lower bound exceeds the best distance (ties are kept and resolved by the
//      reference's own rule: tree order for the kd modes, lowest index for the
//      full scan).
// Distances are the reference's own expressions, so the winner is bit-exact.
// ---------------------------------------------------------------------------

#ifndef NDP_RING_BUDGET
#define NDP_RING_BUDGET 300    // lattice points the shell search may touch before the pyramid descent takes over
#endif
#define NDP_SEARCH_OK 0
#define NDP_SEARCH_TIE 1      // kd mode tie met without topology: caller must redo with NdpTopo
#define NDP_SEARCH_EMPTY 2    // mask has no set bit

struct NdpHit {
    int vertex;      // winning basic vertex id
    double dsel;     // its selection metric
    int status;
    int hard;        // 1 when step (2) ran
    int tied;        // 1 when another candidate scored exactly the winning (rounded) total during the search
    int has_coords;  // coords[] below already holds the basic coordinates of `vertex`
    int coords[NDP_MAX_AXES];
};

// N > 0: number of axes known at compile time (everything stays in registers).
// MODE >= 0: search mode known at compile time (the metric's branches fold away).
template <int N, int MODE = -1>
NDP_HD bool ndp_search_fast(const NdpGeom &g, const NdpPyramid &P, int mode_rt, const NdpQuery &Q, NdpHit &hit)
{
    const int mode = (MODE >= 0) ? MODE : mode_rt;
    constexpr int M = N ? N : NDP_MAX_AXES;
    const int nb = g.nbasic;
    const bool cells = (mode == NDP_MODE_KD_LINEAR || mode == NDP_MODE_LS_LINEAR);
    const int lo = cells ? 1 : 0;
    int c[M];
    int flat = 0;
    bool ok = true;
#ifdef NDP_FAST_MARGIN
    unsigned clamped = 0;
#endif
#pragma unroll
    for (int a = 0; a < M; a++)
        if (a < nb) {
            const double hi = (double) (g.len[a] - 1);
            double p = cells ? floor(Q.qc[a]) + 1.0 : rint(Q.qc[a]);
#ifdef NDP_FAST_MARGIN
            if (p < (double) lo || p > hi) clamped |= 1u << a;         // NaN: not clamped, poisons the margin below
#endif
            p = p < (double) lo ? (double) lo : (p > hi ? hi : p);     // NaN falls through and is rejected below
            c[a] = (int) p;
            ok &= (c[a] >= lo && c[a] <= g.len[a] - 1);
            flat += c[a] * g.bstride[a];
        }
    if (!ok) return false;
    if (!ndp_bit(P.bits[0], flat)) return false;

#ifdef NDP_FAST_MARGIN
    // Margin shortcut (off by default until measured on a B200; validated on the host build).
    // Replacing the candidate's term on axis a by a neighbour's raises the exact sum by at least
    //   point metrics (vertex at c, cell centre at c - 0.5):  1 - 2|q - pos|   (>= 1 where c was clamped)
    //   box metric of ndp_distance (cell [c-1, c]):           min(f, 1 - f)^2, f = q - (c - 1)   (>= 1 if clamped)
    // fl-sums are monotone and every rounding error of the terms and of the ordered sums is below
    // ~50 u (3 d0 + 3), u = 2^-53; a margin above 2^-30 (d0 + 1) therefore proves every neighbour strictly
    // farther without evaluating the 2 nbasic neighbour sums.  Near-ties (and NaN) take the full proof.
    {
        double e0 = 0.0, G = 1.0;
#pragma unroll
        for (int a = 0; a < M; a++)
            if (a < nb) {
                e0 += ndp_axis_term(mode, Q.qc[a], c[a], c[a]);
                double ga;
                if ((clamped >> a) & 1u) ga = 1.0;
                else if (mode == NDP_MODE_LS_LINEAR) {
                    const double f = Q.qc[a] - (double) (c[a] - 1);
                    const double h = f < 1.0 - f ? f : 1.0 - f;
                    ga = h * h;
                } else {
                    const double x = Q.qc[a] - ((double) c[a] - (mode == NDP_MODE_KD_LINEAR ? 0.5 : 0.0));
                    ga = 1.0 - 2.0 * fabs(x);
                }
                G = (ga >= G) ? G : ga;                                 // NaN propagates
            }
        e0 = ndp_add_assoc(g, mode, Q, e0);
        if (!(e0 == e0)) return false;
        if (G > 0x1p-30 * (e0 + 1.0)) {
            hit.vertex = flat; hit.dsel = e0; hit.status = NDP_SEARCH_OK; hit.hard = 0;
#pragma unroll
            for (int a = 0; a < M; a++)
                if (a < nb) hit.coords[a] = c[a];
            hit.has_coords = 1;
            return true;
        }
    }
#endif
    // per-axis terms of the candidate and of its two axis neighbours, each computed once
    double t0[M], tm[M], tp[M];
    bool hm[M], hp[M];
    double d0 = 0.0;
#pragma unroll
    for (int a = 0; a < M; a++)
        if (a < nb) {
            t0[a] = ndp_axis_term(mode, Q.qc[a], c[a], c[a]);
            hm[a] = c[a] - 1 >= lo;
            hp[a] = c[a] + 1 <= g.len[a] - 1;
            tm[a] = hm[a] ? ndp_axis_term(mode, Q.qc[a], c[a] - 1, c[a] - 1) : 0.0;
            tp[a] = hp[a] ? ndp_axis_term(mode, Q.qc[a], c[a] + 1, c[a] + 1) : 0.0;
            d0 += t0[a];
        }
    d0 = ndp_add_assoc(g, mode, Q, d0);
    if (!(d0 == d0)) return false;

    // a neighbour's metric shares the candidate's ordered prefix sum up to its axis
    double prefix = 0.0;
#pragma unroll
    for (int a = 0; a < M; a++)
        if (a < nb) {
            double sm = prefix + tm[a], sp = prefix + tp[a];
#pragma unroll
            for (int b = a + 1; b < M; b++)
                if (b < nb) { sm += t0[b]; sp += t0[b]; }
            sm = ndp_add_assoc(g, mode, Q, sm);
            sp = ndp_add_assoc(g, mode, Q, sp);
            if (hm[a] && !(sm > d0)) ok = false;
            if (hp[a] && !(sp > d0)) ok = false;
            prefix += t0[a];
        }
    if (!ok) return false;
    hit.vertex = flat; hit.dsel = d0; hit.status = NDP_SEARCH_OK; hit.hard = 0;
#pragma unroll
    for (int a = 0; a < M; a++)
        if (a < nb) hit.coords[a] = c[a];
    hit.has_coords = 1;
    return true;
}

// Descent over pyramid P.  When P spans only some basic axes (the mask does not
// depend on the others), fixed[] holds the coordinate of every other axis and the
// descent runs on that slice; the metric still sums all basic axes in order.
template <int MODE>
NDP_HD_CALL void ndp_search_hard(const NdpGeom &g, const NdpPyramid &P, int mode_rt, const NdpQuery &Q,
                            const int *fixed, const NdpTopo *topo, NdpHit &hit, bool seeded = false)
{
    const int mode = (MODE >= 0) ? MODE : mode_rt;
    const int nb = g.nbasic, nv = P.nvar;
    const unsigned nchild = 1u << nv;
    const int top = P.nlev - 1;
    const bool kd = mode < NDP_MODE_LS_NEAREST;
    // children are ordered by which half the query leans to; for cell modes lean on q + 0.5
    const double lean = (mode == NDP_MODE_KD_NEAREST || mode == NDP_MODE_LS_NEAREST) ? 0.0 : 0.5;

    int bc[NDP_MAX_LEVELS][NDP_MAX_AXES];     // block coordinates along the current path (pyramid axes)
    unsigned next[NDP_MAX_LEVELS];            // next child ordinal to try at each level
    int jof[NDP_MAX_AXES];                    // basic axis -> pyramid axis, or -1
    double fterm[NDP_MAX_AXES];               // metric term of the fixed axes
    int base = 0;                             // vertex id contribution of the fixed axes

    for (int a = 0; a < nb; a++) jof[a] = -1;
    for (int j = 0; j < nv; j++) jof[P.var_axis[j]] = j;
    for (int a = 0; a < nb; a++)
        if (jof[a] < 0) { fterm[a] = ndp_axis_term(mode, Q.qc[a], fixed[a], fixed[a]); base += fixed[a] * g.bstride[a]; }

    // a seed (from the ring search) is an upper bound that lets the descent prune from the first block on
    int best = seeded ? hit.vertex : -1;
    double bestd = seeded ? hit.dsel : 0.0;
    bool tie = seeded && hit.status == NDP_SEARCH_TIE;
    bool eq = seeded && hit.tied;             // an equal total was met at the current minimum (any mode)

    hit.hard = 1;
    hit.tied = 0;
    if (!P.any_set) { hit.vertex = -1; hit.dsel = 0.0; hit.status = NDP_SEARCH_EMPTY; return; }

    int lev = top;
    for (int j = 0; j < nv; j++) bc[top][j] = 0;
    next[top] = 0;
    if (top == 0) {
        // a one-point slice (every axis is mask-invariant): that point is the only candidate
        double d = 0.0;
        for (int a = 0; a < nb; a++) d += fterm[a];
        hit.vertex = base; hit.dsel = ndp_add_assoc(g, mode, Q, d); hit.status = NDP_SEARCH_OK;
        return;
    }

    for (;;) {
        if (next[lev] == nchild) {
            if (lev == top) break;
            lev++;
            continue;
        }
        const unsigned k = next[lev]++;
        const int cl = lev - 1;               // child level
        int cb[NDP_MAX_AXES];
        int flat = 0;
        bool inside = true;
        for (int j = 0; j < nv; j++) {
            const int a = P.var_axis[j];
            const int mid = ((2 * bc[lev][j] + 1) << cl);          // first coordinate of the upper child
            const unsigned nearbit = (Q.qc[a] + lean >= (double) mid) ? 1u : 0u;
            const int b = 2 * bc[lev][j] + (int) (((k >> j) & 1u) ^ nearbit);
            if (b >= P.dim[cl][j]) { inside = false; break; }
            cb[j] = b;
            flat += b * P.stride[cl][j];
        }
        if (!inside) continue;
        if (!ndp_bit(P.bits[cl], flat)) continue;
        // lower bound of the block, summed in the reference's axis order; stop as soon as it is hopeless
        double bound = 0.0;
        bool hopeless = false;
        for (int a = 0; a < nb; a++) {
            const int j = jof[a];
            if (j < 0) bound += fterm[a];
            else {
                int lo = cb[j] << cl;
                int hi = ((cb[j] + 1) << cl) - 1;
                if (hi > g.len[a] - 1) hi = g.len[a] - 1;
                bound += ndp_axis_term(mode, Q.qc[a], lo, hi);
            }
            if (best >= 0 && bound > bestd) { hopeless = true; break; }   // terms are >= 0: the sum only grows
        }
        if (hopeless) continue;
        bound = ndp_add_assoc(g, mode, Q, bound);
        if (best >= 0 && bound > bestd) continue;     // equal bounds are explored: ties matter
        if (cl == 0) {
            // a level-0 block is one vertex and its bound is the reference's distance expression
            int vid = base;
            for (int j = 0; j < nv; j++) vid += cb[j] * g.bstride[P.var_axis[j]];
            if (best < 0 || bound < bestd) { best = vid; bestd = bound; tie = false; eq = false; }
            else if (bound == bestd && vid != best) {
                eq = true;
                if (!kd) { if (vid < best) best = vid; }                 // _compare_indexed_dists, :101-112
                else if (topo) { if (ndp_tree_beats(g, mode, *topo, Q, vid, best)) best = vid; }
                else tie = true;
            }
            continue;
        }
        lev = cl;
        for (int j = 0; j < nv; j++) bc[lev][j] = cb[j];
        next[lev] = 0;
    }
    hit.vertex = best;
    hit.dsel = bestd;
    hit.status = tie ? NDP_SEARCH_TIE : NDP_SEARCH_OK;
    hit.tied = eq ? 1 : 0;
}

// Expanding-shell search around the lattice point p0 the query rounds to, over the
// axes pyramid P spans.  Shells are taken over all pyramid axes but the last: a shell
// point is a ROW of the mask along the last axis, and along a row the metric is
// convex in the coordinate, so only the set coordinates nearest to p0 on either
// side can win -- read in O(1) from the prevv/nextv tables (plus any further ones
// that score exactly the same after rounding, for the tie rules).  After shell R every unscanned
// row differs from p0 by more than R on some shell axis, so its metric is at least
// LB_R = min over (axis, side) of the ordered sum with that axis' term evaluated
// R+1 ticks away and every other axis at its own minimum; the best point so far is
// final once it is strictly below LB_R (equality keeps going: ties matter).
// Returns true when proven; otherwise leaves the best so far in `hit` as a seed
// (hit.vertex < 0: nothing found yet) after about `budget` rows.
template <int MODE>
NDP_HD_CALL bool ndp_search_ring(const NdpGeom &g, const NdpPyramid &P, int mode_rt, const NdpQuery &Q,
                            const int *fixed, const NdpTopo *topo, int budget, NdpHit &hit)
{
    const int mode = (MODE >= 0) ? MODE : mode_rt;
    const int nb = g.nbasic, nv = P.nvar;
    const bool cells = (mode == NDP_MODE_KD_LINEAR || mode == NDP_MODE_LS_LINEAR);
    const bool kd = mode < NDP_MODE_LS_NEAREST;
    const int lo = cells ? 1 : 0;
    int jof[NDP_MAX_AXES], p0[NDP_MAX_AXES];
    double tmin[NDP_MAX_AXES];        // per basic axis: the smallest term any admissible coordinate can give
    double term[NDP_MAX_AXES];        // per basic axis: the term of the row being scanned (row axis: unused)
    int base = 0;

    hit.vertex = -1; hit.dsel = 0.0; hit.status = NDP_SEARCH_OK; hit.hard = 1; hit.tied = 0;
    if (nv < 1 || !P.prevv) return false;
    for (int a = 0; a < nb; a++) jof[a] = -1;
    for (int j = 0; j < nv; j++) jof[P.var_axis[j]] = j;
    for (int a = 0; a < nb; a++) {
        if (jof[a] < 0) { tmin[a] = ndp_axis_term(mode, Q.qc[a], fixed[a], fixed[a]); base += fixed[a] * g.bstride[a]; term[a] = tmin[a]; continue; }
        const double hi = (double) (g.len[a] - 1);
        double p = cells ? floor(Q.qc[a]) + 1.0 : rint(Q.qc[a]);
        p = p < (double) lo ? (double) lo : (p > hi ? hi : p);
        const int c = (int) p;
        if (!(c >= lo && c <= g.len[a] - 1)) return false;
        const double t0 = ndp_axis_term(mode, Q.qc[a], c, c);
        // p0 must be the per-axis minimiser for the shell bound and the row argument to hold
        if (c - 1 >= lo && ndp_axis_term(mode, Q.qc[a], c - 1, c - 1) < t0) return false;
        if (c + 1 <= g.len[a] - 1 && ndp_axis_term(mode, Q.qc[a], c + 1, c + 1) < t0) return false;
        if (!(t0 == t0)) return false;
        p0[jof[a]] = c; tmin[a] = t0; term[a] = t0;
    }
    const int ns = nv - 1;                         // shell axes; the last pyramid axis is the row axis
    const int aL = P.var_axis[ns];
    const int lenL = g.len[aL];
    const int strideL = g.bstride[aL];
    const double qL = Q.qc[aL];

    int best = -1;
    double bestd = 0.0;
    bool tie = false, eq = false;
    int rows = 0;

    // scores the set coordinate c of the current row (row_vid = vertex id without the row axis):
    // the ordered sum of the row's terms with the row axis' own term dropped in; returns the metric
    auto score = [&](int row_vid, int c) -> double {
        const double tL = ndp_axis_term(mode, qL, c, c);
        double d = 0.0;
        for (int a = 0; a < nb; a++) d += (a == aL) ? tL : term[a];
        d = ndp_add_assoc(g, mode, Q, d);
        const int vid = row_vid + c * strideL;
        if (best < 0 || d < bestd) { best = vid; bestd = d; tie = false; eq = false; }
        else if (d == bestd && vid != best) {
            eq = true;
            if (!kd) { if (vid < best) best = vid; }                     // _compare_indexed_dists, :101-112
            else if (topo) { if (ndp_tree_beats(g, mode, *topo, Q, vid, best)) best = vid; }
            else tie = true;
        }
        return d;
    };

    for (int R = 0;; R++) {
        int off[NDP_MAX_AXES];
        for (int j = 0; j < ns; j++) off[j] = -R;
        for (;;) {
            int cheb = 0, rowflat = 0, row_vid = base;
            bool inside = true;
            for (int j = 0; j < ns; j++) {
                const int o = off[j] < 0 ? -off[j] : off[j];
                if (o > cheb) cheb = o;
                const int a = P.var_axis[j];
                const int c = p0[j] + off[j];
                if (c < lo || c > g.len[a] - 1) inside = false;
                rowflat += c * P.stride[0][j];
                row_vid += c * g.bstride[a];
            }
            if (cheb == R && inside) {
                rows++;
                for (int j = 0; j < ns; j++) {
                    const int a = P.var_axis[j];
                    term[a] = ndp_axis_term(mode, Q.qc[a], p0[j] + off[j], p0[j] + off[j]);
                }
                // Set coordinates of the row, outward from p0 on either side.  The metric grows
                // (weakly, after rounding) with the distance from p0, so a side is exhausted as
                // soon as a coordinate scores strictly above the best so far; equal scores keep
                // going because ties are decided by the reference's rules, not by us.
                int c = P.prevv[rowflat + p0[ns]];
                while (c >= lo) {
                    if (score(row_vid, c) > bestd) break;
                    c = (c - 1 >= lo) ? P.prevv[rowflat + c - 1] : -1;
                }
                c = (p0[ns] + 1 <= lenL - 1) ? P.nextv[rowflat + p0[ns] + 1] : -1;
                while (c >= lo) {
                    if (score(row_vid, c) > bestd) break;
                    c = (c + 1 <= lenL - 1) ? P.nextv[rowflat + c + 1] : -1;
                }
            }
            // next shell offset; interior offsets (norm < R) are skipped by jumping the last shell axis
            int j = ns - 1;
            if (ns > 0 && R > 0) {
                bool lead_inside = true;
                for (int jj = 0; jj < ns - 1; jj++) { const int o = off[jj] < 0 ? -off[jj] : off[jj]; if (o == R) lead_inside = false; }
                if (lead_inside && off[j] == -R) { off[j] = R; continue; }
            }
            while (j >= 0) { if (++off[j] <= R) break; off[j] = -R; j--; }
            if (j < 0) break;
        }
        // lower bound for every row beyond shell R
        bool any_beyond = false;
        double lb = 0.0;
        for (int j = 0; j < ns; j++) {
            const int a = P.var_axis[j];
            for (int sgn = -1; sgn <= 1; sgn += 2) {
                const int c = p0[j] + sgn * (R + 1);
                if (c < lo || c > g.len[a] - 1) continue;
                const double ta = ndp_axis_term(mode, Q.qc[a], c, c);
                double d = 0.0;
                for (int b = 0; b < nb; b++) d += (b == a) ? ta : tmin[b];
                d = ndp_add_assoc(g, mode, Q, d);
                if (!any_beyond || d < lb) lb = d;
                any_beyond = true;
            }
        }
        hit.vertex = best; hit.dsel = bestd; hit.status = tie ? NDP_SEARCH_TIE : NDP_SEARCH_OK; hit.tied = eq ? 1 : 0;
        if (!any_beyond) {                       // every row of the slice has been scanned
            if (best < 0) hit.status = NDP_SEARCH_EMPTY;
            return true;
        }
        if (best >= 0 && bestd < lb) return true;
        if (rows >= budget) return false;
    }
}

// The same shell search with the number of basic axes fixed at compile time: every per-axis
// quantity is indexed by the basic axis with a constant index, so the whole state lives in
// registers (the generic version above keeps its arrays in local memory).  kind[a]: 0 = fixed
// (mask-invariant axis solved in closed form), 1 = shell axis, 2 = the row axis.
template <int MODE, int NB>
NDP_HD_CALL bool ndp_search_ring_nb(const NdpGeom &g, const NdpPyramid &P, const NdpQuery &Q,
                               const int *fixed, const NdpTopo *topo, int budget, NdpHit &hit)
{
    constexpr int mode = MODE;
    const int nv = P.nvar;
    constexpr bool cells = (mode == NDP_MODE_KD_LINEAR || mode == NDP_MODE_LS_LINEAR);
    constexpr bool kd = mode < NDP_MODE_LS_NEAREST;
    constexpr int lo = cells ? 1 : 0;

    hit.vertex = -1; hit.dsel = 0.0; hit.status = NDP_SEARCH_OK; hit.hard = 1; hit.tied = 0;
    if (nv < 1 || !P.prevv) return false;

    int kind[NB], pstr[NB], p0[NB], off[NB], len[NB], bstr[NB];
    double q[NB], tmin[NB], term[NB];
    bool last_shell[NB];
#pragma unroll
    for (int a = 0; a < NB; a++) { kind[a] = 0; pstr[a] = 0; off[a] = 0; last_shell[a] = false; len[a] = g.len[a]; bstr[a] = g.bstride[a]; q[a] = Q.qc[a]; }
#pragma unroll
    for (int j = 0; j < NB; j++)
        if (j < nv) {
            const int va = P.var_axis[j], st = P.stride[0][j];
#pragma unroll
            for (int a = 0; a < NB; a++)
                if (va == a) { kind[a] = (j == nv - 1) ? 2 : 1; pstr[a] = st; last_shell[a] = (j == nv - 2); }
        }
    int base = 0;
    bool ok = true;
#pragma unroll
    for (int a = 0; a < NB; a++) {
        if (kind[a] == 0) {
            p0[a] = fixed[a];
            tmin[a] = ndp_axis_term(mode, q[a], p0[a], p0[a]);
            base += p0[a] * bstr[a];
        } else {
            const double hi = (double) (len[a] - 1);
            double p = cells ? floor(q[a]) + 1.0 : rint(q[a]);
            p = p < (double) lo ? (double) lo : (p > hi ? hi : p);
            const int c = (int) p;
            ok &= (c >= lo && c <= len[a] - 1);
            const double t0 = ndp_axis_term(mode, q[a], c, c);
            // p0 must be the per-axis minimiser for the shell bound and the row argument to hold
            if (c - 1 >= lo && ndp_axis_term(mode, q[a], c - 1, c - 1) < t0) ok = false;
            if (c + 1 <= len[a] - 1 && ndp_axis_term(mode, q[a], c + 1, c + 1) < t0) ok = false;
            ok &= (t0 == t0);
            p0[a] = c; tmin[a] = t0;
        }
        term[a] = tmin[a];
    }
    if (!ok) return false;
    int p0row = 0, lenrow = 0, bstrrow = 0;
    double qrow = 0.0;
#pragma unroll
    for (int a = 0; a < NB; a++)
        if (kind[a] == 2) { p0row = p0[a]; lenrow = len[a]; bstrrow = bstr[a]; qrow = q[a]; }

    int best = -1;
    double bestd = 0.0;
    bool tie = false, eq = false;
    int rows = 0;

    auto score = [&](int row_vid, int c) -> double {
        const double tL = ndp_axis_term(mode, qrow, c, c);
        double d = 0.0;
#pragma unroll
        for (int a = 0; a < NB; a++) d += (kind[a] == 2) ? tL : term[a];
        d = ndp_add_assoc(g, mode, Q, d);
        const int vid = row_vid + c * bstrrow;
        if (best < 0 || d < bestd) { best = vid; bestd = d; tie = false; eq = false; }
        else if (d == bestd && vid != best) {
            eq = true;
            if (!kd) { if (vid < best) best = vid; }                     // _compare_indexed_dists, :101-112
            else if (topo) { if (ndp_tree_beats(g, mode, *topo, Q, vid, best)) best = vid; }
            else tie = true;
        }
        return d;
    };

    for (int R = 0;; R++) {
#pragma unroll
        for (int a = 0; a < NB; a++) off[a] = (kind[a] == 1) ? -R : 0;
        for (;;) {
            int cheb = 0, rowflat = 0, row_vid = base;
            bool inside = true, lead_inside = true;
#pragma unroll
            for (int a = 0; a < NB; a++)
                if (kind[a] == 1) {
                    const int o = off[a] < 0 ? -off[a] : off[a];
                    cheb = o > cheb ? o : cheb;
                    if (!last_shell[a] && o == R) lead_inside = false;
                    const int c = p0[a] + off[a];
                    inside &= (c >= lo && c <= len[a] - 1);
                    rowflat += c * pstr[a];
                    row_vid += c * bstr[a];
                }
            if (cheb == R && inside) {
                rows++;
#pragma unroll
                for (int a = 0; a < NB; a++)
                    if (kind[a] == 1) term[a] = ndp_axis_term(mode, q[a], p0[a] + off[a], p0[a] + off[a]);
                // set coordinates of the row, outward from p0 on either side, while the rounded total
                // does not exceed the best (equal totals are ties the reference's rules must decide)
                int c = P.prevv[rowflat + p0row];
                while (c >= lo) {
                    if (score(row_vid, c) > bestd) break;
                    c = (c - 1 >= lo) ? P.prevv[rowflat + c - 1] : -1;
                }
                c = (p0row + 1 <= lenrow - 1) ? P.nextv[rowflat + p0row + 1] : -1;
                while (c >= lo) {
                    if (score(row_vid, c) > bestd) break;
                    c = (c + 1 <= lenrow - 1) ? P.nextv[rowflat + c + 1] : -1;
                }
            }
            // next shell offset: an odometer over the shell axes, last shell axis fastest; interior
            // offsets (all norms < R) are skipped by jumping the last shell axis from -R to +R
            bool carry = true, jumped = false;
#pragma unroll
            for (int a = NB - 1; a >= 0; a--)
                if (kind[a] == 1 && carry) {
                    if (last_shell[a] && R > 0 && lead_inside && off[a] == -R) { off[a] = R; carry = false; jumped = true; }
                    else if (++off[a] <= R) carry = false;
                    else off[a] = -R;
                }
            (void) jumped;
            if (carry) break;
        }
        // lower bound for every row beyond shell R
        bool any_beyond = false;
        double lb = 0.0;
#pragma unroll
        for (int a = 0; a < NB; a++)
            if (kind[a] == 1) {
#pragma unroll
                for (int sgn = -1; sgn <= 1; sgn += 2) {
                    const int c = p0[a] + sgn * (R + 1);
                    if (c >= lo && c <= len[a] - 1) {
                        const double ta = ndp_axis_term(mode, q[a], c, c);
                        double d = 0.0;
#pragma unroll
                        for (int b = 0; b < NB; b++) d += (b == a) ? ta : tmin[b];
                        d = ndp_add_assoc(g, mode, Q, d);
                        if (!any_beyond || d < lb) lb = d;
                        any_beyond = true;
                    }
                }
            }
        hit.vertex = best; hit.dsel = bestd; hit.status = tie ? NDP_SEARCH_TIE : NDP_SEARCH_OK; hit.tied = eq ? 1 : 0;
        if (!any_beyond) {                       // every row of the slice has been scanned
            if (best < 0) hit.status = NDP_SEARCH_EMPTY;
            return true;
        }
        if (best >= 0 && bestd < lb) return true;
        if (rows >= budget) return false;
    }
}

// dispatch on the number of basic axes (MODE known at compile time), generic version otherwise
template <int MODE>
NDP_HD bool ndp_search_ring_any(const NdpGeom &g, const NdpPyramid &P, int mode, const NdpQuery &Q,
                                const int *fixed, const NdpTopo *topo, int budget, NdpHit &hit)
{
    if (MODE >= 0) {
        constexpr int M = MODE >= 0 ? MODE : 0;
        switch (g.nbasic) {
        case 1: return ndp_search_ring_nb<M, 1>(g, P, Q, fixed, topo, budget, hit);
        case 2: return ndp_search_ring_nb<M, 2>(g, P, Q, fixed, topo, budget, hit);
        case 3: return ndp_search_ring_nb<M, 3>(g, P, Q, fixed, topo, budget, hit);
        case 4: return ndp_search_ring_nb<M, 4>(g, P, Q, fixed, topo, budget, hit);
        case 5: return ndp_search_ring_nb<M, 5>(g, P, Q, fixed, topo, budget, hit);
        case 6: return ndp_search_ring_nb<M, 6>(g, P, Q, fixed, topo, budget, hit);
        default: break;
        }
    }
    return ndp_search_ring<MODE>(g, P, mode, Q, fixed, topo, budget, hit);
}

// The search proper.  Axes along which the mask does not vary (a property of the
// table, found at register time; typical of atmosphere-like tables whose voids
// depend on two of many axes) are solved in closed form: the winner takes the
// coordinate that minimises that axis' own term, provided that minimiser is strict
// both per axis and in the final rounded sum -- otherwise (exact half-way queries,
// absorbed differences) the full-lattice descent decides.  R is the pyramid over
// the variant axes only, F the full one.
template <int MODE>
NDP_HD void ndp_search_descend(const NdpGeom &g, const NdpPyramid &F, const NdpPyramid &R, int mode_rt,
                               const NdpQuery &Q, const NdpTopo *topo, NdpHit &hit)
{
    const int mode = (MODE >= 0) ? MODE : mode_rt;
    const int nb = g.nbasic;
    if (R.nvar < nb && R.any_set) {
        const bool cells = (mode == NDP_MODE_KD_LINEAR || mode == NDP_MODE_LS_LINEAR);
        const int lo = cells ? 1 : 0;
        int fixed[NDP_MAX_AXES];
        bool is_var[NDP_MAX_AXES];
        double talt_m[NDP_MAX_AXES], talt_p[NDP_MAX_AXES];   // terms of a fixed axis' two neighbours (+inf: none)
        const double inf = 1e308 * 10.0;
        bool ok = true;
        for (int a = 0; a < nb; a++) is_var[a] = false;
        for (int j = 0; j < R.nvar; j++) is_var[R.var_axis[j]] = true;
        for (int a = 0; a < nb && ok; a++) {
            if (is_var[a]) { fixed[a] = 0; continue; }
            const double hi = (double) (g.len[a] - 1);
            double p = cells ? floor(Q.qc[a]) + 1.0 : rint(Q.qc[a]);
            p = p < (double) lo ? (double) lo : (p > hi ? hi : p);
            const int c = (int) p;
            if (!(c >= lo && c <= g.len[a] - 1)) { ok = false; break; }
            const double t0 = ndp_axis_term(mode, Q.qc[a], c, c);
            talt_m[a] = (c - 1 >= lo) ? ndp_axis_term(mode, Q.qc[a], c - 1, c - 1) : inf;
            talt_p[a] = (c + 1 <= g.len[a] - 1) ? ndp_axis_term(mode, Q.qc[a], c + 1, c + 1) : inf;
            if (!(talt_m[a] > t0) || !(talt_p[a] > t0)) ok = false;
            fixed[a] = c;
        }
        if (ok) {
            if (!ndp_search_ring_any<MODE>(g, R, mode, Q, fixed, topo, NDP_RING_BUDGET, hit))
                ndp_search_hard<MODE>(g, R, mode, Q, fixed, topo, hit, hit.vertex >= 0);
            // Another candidate of the slice scored exactly the winning total: the re-summation below vouches for the
            // reported winner only, while a tied candidate's neighbour along a closed-form axis may tie too (absorbed
            // terms) and win by the reference's tie rule -- found by tools/fuzz_hostsim.py: the full lattice decides.
            if (hit.tied) ok = false;
            if (ok && hit.status != NDP_SEARCH_EMPTY && hit.vertex >= 0) {
                // the closed-form axes must also be strict minimisers of the rounded total: re-sum the
                // winner's terms in order with one fixed axis' term swapped for a neighbour's
                int c[NDP_MAX_AXES];
                double tw[NDP_MAX_AXES];
                ndp_decode_vertex(g, hit.vertex, c);
                for (int a = 0; a < nb; a++) tw[a] = ndp_axis_term(mode, Q.qc[a], c[a], c[a]);
                for (int a = 0; a < nb && ok; a++) {
                    if (is_var[a]) continue;
                    for (int s = 0; s < 2; s++) {
                        const double alt = s ? talt_p[a] : talt_m[a];
                        if (alt >= inf) continue;
                        double d = 0.0;
                        for (int b = 0; b < nb; b++) d += (b == a) ? alt : tw[b];
                        d = ndp_add_assoc(g, mode, Q, d);
                        if (!(d > hit.dsel)) ok = false;
                    }
                }
                if (ok) return;
            }
        }
    }
    if (!F.any_set) { hit.vertex = -1; hit.dsel = 0.0; hit.status = NDP_SEARCH_EMPTY; hit.hard = 1; return; }
    if (!ndp_search_ring_any<MODE>(g, F, mode, Q, nullptr, topo, NDP_RING_BUDGET, hit))
        ndp_search_hard<MODE>(g, F, mode, Q, nullptr, topo, hit, hit.vertex >= 0);
}

// The full scan scores masked-out vertices 1e10 (:283-286), so beyond 1e5 cells
// from every valid vertex the lowest masked-out id wins; an empty mask yields
// vertex 0 (all scores equal, lowest id first).
NDP_HD void ndp_scan_sentinel(const NdpPyramid &P, int mode, NdpHit &hit)
{
    if (mode < NDP_MODE_LS_NEAREST) return;
    const double sentinel = 1e10;
    if (hit.status == NDP_SEARCH_EMPTY) { hit.vertex = 0; hit.dsel = sentinel; hit.status = NDP_SEARCH_OK; return; }
    if (P.first_clear >= 0 && (sentinel < hit.dsel || (sentinel == hit.dsel && P.first_clear < hit.vertex))) {
        hit.vertex = P.first_clear; hit.dsel = sentinel;
    }
}

// Decodes a basic vertex id into coords[0..nbasic) (:235-237) and appends the
// associated ticks.
NDP_HD void ndp_coords_of(const NdpGeom &g, const NdpQuery &Q, int vertex, int *coords)
{
    ndp_decode_vertex(g, vertex, coords);
    for (int a = g.nbasic; a < g.naxes; a++) coords[a] = Q.acoord[a];
}

// Builds the index-space query from the import products.
NDP_HD void ndp_make_query(const NdpGeom &g, const int *idx, const double *nrm, NdpQuery &Q)
{
    for (int a = 0; a < g.naxes; a++) {
        Q.qc[a] = (double) idx[a] + nrm[a] - 1.0;
        Q.acoord[a] = (a >= g.nbasic) ? ndp_assoc_coord(Q.qc[a], g.len[a]) : 0;
    }
}

// ---------------------------------------------------------------------------
// Per-query glue shared by the kernels (and by the test-only host build)
// ---------------------------------------------------------------------------

// N > 0: number of axes fixed at compile time; N == 0: read it from the geometry.
//
// For N > 0 the guess-and-verify search of ndp_import_axis_guess is laid out in phases over ALL
// axes -- index guess and tick loads, one corrective tick, verification, then the quotients, then
// the flags -- as straight-line code, so that the independent per-axis chains interleave in the
// instruction stream (a warp issues in order: one axis after the other would leave every fp64
// latency exposed).  Axes whose guess does not verify are redone by the reference bisection.
template <int N>
NDP_HD void ndp_import_query(const NdpGeom &g, const double *ticks, const double *qrow,
                             int *idx, int *flg, double *nrm)
{
    if (N == 0) {
        const int n = g.naxes;
        for (int a = 0; a < n; a++)
            ndp_import_axis_guess(ticks + g.tickoff[a], g.len[a], g.inv_step[a], g.first[a], g.last[a],
                                  qrow[a], idx[a], flg[a], nrm[a]);
        return;
    }
    constexpr int M = N ? N : 1;
    const double rtol = 1e-6;
    if (g.easy) {
        // evenly spaced axes with power-of-two spacing (ndp_geom_axes): guess, verify, scale -- no corrective
        // step, no division; an unverified guess (a query within an ulp of a tick) takes the bisection
        bool all_ok = true;
#pragma unroll
        for (int a = 0; a < M; a++) {
            const double *tka = ticks + g.tickoff[a];
            const int lena = g.len[a];
            const double xa = qrow[a], xra = xa + rtol;
            int c = (int) ((xra - g.first[a]) * g.inv_pow2[a]);
            c = c < lena - 2 ? c : lena - 2;
            const int la = c > 0 ? c + 1 : 1;
            const double tla = tka[la], tpa = tka[la - 1];
            const bool oka = (!(xra > tla) || la == lena - 1) && (la == 1 || xra > tpa);
            const double ta = (xa - tpa) * g.inv_pow2[a];
            int f = (xa < g.first[a] || xa > g.last[a]) ? NDP_F_OOB : 0;
            if (fabs(ta) < rtol || (la == lena - 1 && fabs(ta - 1) < rtol)) f |= NDP_F_ON_VERTEX;
            idx[a] = la; flg[a] = f; nrm[a] = ta;
            all_ok &= oka;
            if (!oka) idx[a] = -1;
        }
        if (!all_ok) {
#pragma unroll
            for (int a = 0; a < M; a++)
                if (idx[a] < 0) ndp_import_axis(ticks + g.tickoff[a], g.len[a], qrow[a], idx[a], flg[a], nrm[a]);
        }
        return;
    }
    double x[M], xr[M], tl[M], tp[M];
    int l[M], len[M];
    const double *tk[M];
    bool ok[M];
#pragma unroll
    for (int a = 0; a < M; a++) {
        tk[a] = ticks + g.tickoff[a];
        len[a] = g.len[a];
        x[a] = qrow[a];
        xr[a] = x[a] + rtol;
        int c = (int) ((xr[a] - g.first[a]) * g.inv_step[a]);
        c = c < len[a] - 2 ? c : len[a] - 2;
        l[a] = c > 0 ? c + 1 : 1;
    }
#pragma unroll
    for (int a = 0; a < M; a++) { tl[a] = tk[a][l[a]]; tp[a] = tk[a][l[a] - 1]; }
#pragma unroll
    for (int a = 0; a < M; a++) {
        const bool up = (xr[a] > tl[a]) && (l[a] < len[a] - 1);
        const bool dn = !up && !(xr[a] > tp[a]) && (l[a] > 1);
        l[a] += (up ? 1 : 0) - (dn ? 1 : 0);
    }
#pragma unroll
    for (int a = 0; a < M; a++) { tl[a] = tk[a][l[a]]; tp[a] = tk[a][l[a] - 1]; }
    bool all_ok = true;
#pragma unroll
    for (int a = 0; a < M; a++) {
        ok[a] = g.inv_step[a] > 0.0 && (!(xr[a] > tl[a]) || l[a] == len[a] - 1) && (l[a] == 1 || xr[a] > tp[a]);
        all_ok &= ok[a];
    }
    double t[M];
#pragma unroll
    for (int a = 0; a < M; a++) t[a] = (x[a] - tp[a]) / (tl[a] - tp[a]);
#pragma unroll
    for (int a = 0; a < M; a++) {
        int f = (x[a] < g.first[a] || x[a] > g.last[a]) ? NDP_F_OOB : 0;
        if (fabs(t[a]) < rtol || (l[a] == len[a] - 1 && fabs(t[a] - 1) < rtol)) f |= NDP_F_ON_VERTEX;
        idx[a] = l[a]; flg[a] = f; nrm[a] = t[a];
    }
    if (!all_ok) {
#pragma unroll
        for (int a = 0; a < M; a++)
            if (!ok[a]) ndp_import_axis(tk[a], len[a], x[a], idx[a], flg[a], nrm[a]);
    }
}

// The plain per-axis form of the same import, for kernels that only import when no record is
// available: compact code and few registers matter more there than instruction-level parallelism.
template <int N>
NDP_HD void ndp_import_query_seq(const NdpGeom &g, const double *ticks, const double *qrow,
                                 int *idx, int *flg, double *nrm)
{
    const int n = N ? N : g.naxes;
#pragma unroll
    for (int a = 0; a < (N ? N : NDP_MAX_AXES); a++)
        if (a < n) ndp_import_axis_guess(ticks + g.tickoff[a], g.len[a], g.inv_step[a], g.first[a], g.last[a],
                                         qrow[a], idx[a], flg[a], nrm[a]);
}

// Reduced value of component k of the cell whose superior corner is sup[] (:472-498
// gather + :78-99 reduction).  CM selects the cell-major copy of the grid
// (2^n * vdim contiguous doubles per cell, `cells`), else the original strided
// layout (`grid`).  V1 asserts vdim == 1 at compile time (128-bit pair loads).
// pinmask/selbits use bit (n-1-a) for axis a.
template <int N, bool CM, bool V1>
NDP_HD double ndp_eval_cell(const NdpGeom &g, const double *__restrict__ grid, const double *__restrict__ cells,
                            const int *sup, const double *x, unsigned pinmask, unsigned selbits, int k, bool &nan_used)
{
    const int V = V1 ? 1 : g.vdim;
    if (N > 0) {
        constexpr int NN = N ? N : 1;
        if (CM) {
            long long cell = 0;
            if (g.small) {
                int c32 = 0;
#pragma unroll
                for (int a = 0; a < NN; a++) c32 += (sup[a] - 1) * g.cstride32[a];
                cell = c32;
            } else {
#pragma unroll
                for (int a = 0; a < NN; a++) cell += (long long) (sup[a] - 1) * g.cstride[a];
            }
            const double *base = cells + (cell << NN) * V + k;
            if (V1 && NN >= 2) {
                constexpr int N2 = NN >= 2 ? NN : 2;
                return ndp_cell_reduce_halves<N2>(reinterpret_cast<const double2 *>(base), x, pinmask, selbits, nan_used);
            }
            if (V1) {
                const double2 *b2 = reinterpret_cast<const double2 *>(base);
                return ndp_cell_reduce<NN>([&](double *fv) {
#pragma unroll
                    for (int j = 0; j < (1 << NN) / 2; j++) { double2 p = b2[j]; fv[2 * j] = p.x; fv[2 * j + 1] = p.y; }
                }, x, pinmask, selbits, nan_used);
            }
            return ndp_cell_reduce<NN>([&](double *fv) {
#pragma unroll
                for (int j = 0; j < (1 << NN); j++) fv[j] = base[(long long) j * V];
            }, x, pinmask, selbits, nan_used);
        }
        long long origin = 0;
#pragma unroll
        for (int a = 0; a < NN; a++) origin += (long long) (sup[a] - 1) * g.vstride[a];
        const double *base = grid + origin * V + k;
        return ndp_cell_reduce<NN>([&](double *fv) {
#pragma unroll
            for (int j = 0; j < (1 << NN); j++) {
                long long off = 0;
#pragma unroll
                for (int a = 0; a < NN; a++) if ((j >> (NN - 1 - a)) & 1) off += g.vstride[a];
                fv[j] = base[off * V];
            }
        }, x, pinmask, selbits, nan_used);
    }
    // runtime n: used corners only, carry-chain reduction, original layout
    const int n = g.naxes;
    long long fixed = 0;
    long long fstride[NDP_MAX_AXES];
    double xfree[NDP_MAX_AXES];
    int d = 0;
    for (int a = 0; a < n; a++) {
        const bool pinned = (pinmask >> (n - 1 - a)) & 1u;
        const bool upper = (selbits >> (n - 1 - a)) & 1u;
        if (pinned) fixed += (long long) (upper ? sup[a] : sup[a] - 1) * g.vstride[a];
        else { fixed += (long long) (sup[a] - 1) * g.vstride[a]; fstride[d] = g.vstride[a]; xfree[d] = x[a]; d++; }
    }
    const double *base = grid + fixed * V + k;
    return ndp_cell_reduce_rt(d, [&](unsigned c) {
        long long off = 0;
        for (int i = 0; i < d; i++) if ((c >> i) & 1u) off += fstride[i];
        return base[off * V];
    }, xfree, nan_used);
}

// Classification of an imported query (:438-460): out-of-bounds, pinned axes and
// which side each pinned axis selects (:476).
template <int N>
NDP_HD bool ndp_classify(const NdpGeom &g, const int *flg, const double *nrm, unsigned &pinmask, unsigned &selbits)
{
    const int n = N ? N : g.naxes;
    bool oob = false;
    pinmask = 0; selbits = 0;
#pragma unroll
    for (int a = 0; a < (N ? N : NDP_MAX_AXES); a++)
        if (a < n) {
            oob |= (flg[a] & NDP_F_OOB) != 0;
            if (flg[a] & NDP_F_ON_VERTEX) {
                pinmask |= 1u << (n - 1 - a);
                if (nrm[a] > 0.5) selbits |= 1u << (n - 1 - a);
            }
        }
    return oob;
}

// Off-grid finish for component k once the winner `coords` is known.
//   NEAREST (:556-567): the vertex' own value.
//   LINEAR  (:569-635): full-dimensional cell below coords (clamped at 0, :601),
//                       query shifted into that cell's unit coordinates (:621-622).
template <int N, bool CM, bool V1>
NDP_HD double ndp_finish_component(const NdpGeom &g, const double *__restrict__ grid, const double *__restrict__ cells,
                                   int method, const int *idx, const double *nrm, const int *coords, int k)
{
    const int n = N ? N : g.naxes;
    if (method == NDP_M_NEAREST) {
        long long v = 0;
        for (int a = 0; a < n; a++) v += (long long) coords[a] * g.vstride[a];
        return grid[v * (V1 ? 1 : g.vdim) + k];
    }
    int sup[NDP_MAX_AXES];
    double x[NDP_MAX_AXES];
#pragma unroll
    for (int a = 0; a < (N ? N : NDP_MAX_AXES); a++)
        if (a < n) {
            sup[a] = coords[a] > 1 ? coords[a] : 1;
            x[a] = nrm[a] + (double) (idx[a] - coords[a]);
        }
    bool unused;
    return ndp_eval_cell<N, CM, V1>(g, grid, cells, sup, x, 0u, 0u, k, unused);
}
