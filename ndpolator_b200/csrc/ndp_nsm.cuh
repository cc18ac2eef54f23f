// ndp_nsm.cuh -- nearest-site map: the register-time classification that turns the
// off-grid nearest search (ndp_find_nearest, src/ndpolator.c:188-352; kd_nearest,
// external/kdtree/kdtree.c:345-457) into a table lookup plus a handful of exact
// distance evaluations.
//
// The candidates of the reference's search are the set lattice points of a mask
// (vmask for NEAREST, hcmask for LINEAR).  On the axes the mask does not depend on
// the winner is known in closed form (ndp_search_descend); on the remaining
// "variant" axes the index space is cut into REGIONS -- per axis: S sub-cells per
// unit cell inside the box, and geometric bands [0,1] [1,2] [2,4] ... [16,32]
// outside either end -- and for every region the map stores the sites that can be
// nearest (within a slack NSM_EPS of the minimum) for SOME query inside it:
//
//   site s is dropped iff another site s' is closer by more than NSM_EPS for EVERY
//   query of the region:   inf_{q in region} [ d(q,s) - d(q,s') ] > NSM_EPS.
//
// The metric is separable and each per-axis difference is linear (point metrics)
// or a difference of single smooth pieces (box metric; sub-cells never straddle an
// integer), so the infimum over the product region is the sum of per-axis minima
// over the two interval ends -- exact in fp64 (small dyadic numbers).
//
// At query time every stored candidate is scored with the REFERENCE'S OWN
// expressions in its order, so among the candidates the decision is bit-exact; a
// dropped site's true distance exceeds the winner's by more than NSM_EPS = 2^-20,
// while all rounding errors of a total below NSM_TMAX = 2^20 stay under 2^-28, so
// no dropped site can win or tie after rounding either.  (On the closed-form axes the
// same argument needs a margin of 2^-32 only while the winning total is below 2^8.)
// Whatever the map cannot
// settle -- exact ties in the kd modes (the reference's traversal order decides),
// totals above NSM_TMAX, queries more than 32 cells outside the box on a variant
// axis, NaN, a non-strict closed-form axis -- is reported as "not settled" and goes
// through the exact lattice search (ndp_search_descend) instead.
#pragma once

#include "ndp_core.cuh"

#define NSM_BANDS 6                 // exterior bands per side: [0,1] [1,2] [2,4] [4,8] [8,16] [16,32]
#define NSM_KMAX 160                // candidates a region may hold while being built
#define NSM_EPS 0x1p-20
#define NSM_TMAX 0x1p20
#define NSM_EPS_NEAR 0x1p-32        // margin that is enough when the winning total stays below NSM_TNEAR: the rounding
#define NSM_TNEAR 0x1p8             // errors of such ordered sums are below 12 * 2^-52 * 2^8 < 2^-40
#define NSM_MAX_VAR 4               // variant axes a map may span (16-bit coordinates, 64-bit packing)

#define NSM_GEO_VERTEX 0            // vmask sites at integer c            (modes KD_NEAREST, LS_NEAREST)
#define NSM_GEO_CENTRE 1            // hcmask sites at c - 0.5             (mode KD_LINEAR)
#define NSM_GEO_BOX 2               // hcmask cells [c-1, c], box distance (mode LS_LINEAR)

NDP_HD int nsm_geo_of_mode(int mode)
{
    return mode == NDP_MODE_KD_LINEAR ? NSM_GEO_CENTRE : (mode == NDP_MODE_LS_LINEAR ? NSM_GEO_BOX : NSM_GEO_VERTEX);
}
NDP_HD int nsm_mode_of_geo(int geo)      // a mode whose per-axis term has this geometry
{
    return geo == NSM_GEO_CENTRE ? NDP_MODE_KD_LINEAR : (geo == NSM_GEO_BOX ? NDP_MODE_LS_LINEAR : NDP_MODE_KD_NEAREST);
}

struct NdpSiteMap {
    int valid;                        // 0: no map for this geometry (use the lattice search)
    int geo;
    int nvar;                         // variant axes (same set and order as the reduced pyramid's)
    int jof[NDP_MAX_AXES];            // basic axis -> variant index, or -1 (closed form)
    int var_axis[NSM_MAX_VAR];
    int vlen[NSM_MAX_VAR];            // ticks of the variant axes
    int sstride[NSM_MAX_VAR];         // strides of the slice mask (reduced pyramid level 0)
    int S;                            // sub-cells per unit cell
    int nr[NSM_MAX_VAR];              // regions per variant axis = 2 NSM_BANDS + S (len - 1)
    int rstride[NSM_MAX_VAR];
    int nregions;
    int ew;                           // 32-bit words per entry = (nvar + 1) / 2, 16 bits per coordinate
    int lo;                           // lowest site coordinate: 0 (vertices) or 1 (cells)
    const int *offs;                  // nregions + 1 entries
    const uint32_t *ents;
    const uint32_t *head;             // per region 2 * ew + 2 words (16 / 32 bytes with padding): its first TWO candidates, its
                                      // candidate count and the start of its list in ents: one load settles most regions
    const uint32_t *slice;            // the slice mask itself (level 0 of the reduced pyramid)
};

#define NSM_MAX_REGIONS (1 << 21)

// Fills the geometry of a map over the variant axes of reduced pyramid R (mask slice R.bits[0]);
// offs / ents are attached by the caller once the regions have been counted.  valid = 0 when the
// mask spans too many variant axes or regions (such tables keep the lattice search).
inline void nsm_setup(const NdpGeom &g, const NdpPyramid &R, int geo, NdpSiteMap &M, int max_subcells = 16)
{
    M = NdpSiteMap{};
    M.geo = geo;
    M.nvar = R.nvar;
    M.lo = geo == NSM_GEO_VERTEX ? 0 : 1;
    M.slice = R.bits[0];
    M.ew = R.nvar > 2 ? 2 : 1;
    for (int a = 0; a < NDP_MAX_AXES; a++) M.jof[a] = -1;
    if (R.nvar > NSM_MAX_VAR || !R.any_set) return;
    for (int j = 0; j < R.nvar; j++) {
        M.var_axis[j] = R.var_axis[j];
        M.jof[R.var_axis[j]] = j;
        M.vlen[j] = g.len[R.var_axis[j]];
        M.sstride[j] = R.stride[0][j];
    }
    for (int S = max_subcells; S >= 1; S >>= 1) {
        long long total = 1;
        for (int j = 0; j < R.nvar; j++) total *= 2 * NSM_BANDS + (long long) S * (M.vlen[j] - 1);
        if (total > NSM_MAX_REGIONS) continue;
        M.S = S;
        M.nregions = (int) total;
        for (int j = R.nvar - 1, st = 1; j >= 0; j--) { M.nr[j] = 2 * NSM_BANDS + S * (M.vlen[j] - 1); M.rstride[j] = st; st *= M.nr[j]; }
        M.valid = 1;
        return;
    }
}

NDP_HD int nsm_band_lo(int b) { return b == 0 ? 0 : (1 << (b - 1)); }
NDP_HD int nsm_band_hi(int b) { return 1 << b; }

// interval [u, v] of region r on an axis with L ticks
NDP_HD void nsm_interval(int S, int L, int r, double &u, double &v)
{
    const int ni = S * (L - 1);
    if (r < NSM_BANDS) { const int b = NSM_BANDS - 1 - r; u = -(double) nsm_band_hi(b); v = -(double) nsm_band_lo(b); }
    else if (r < NSM_BANDS + ni) { const int k = r - NSM_BANDS; u = (double) k / (double) S; v = (double) (k + 1) / (double) S; }
    else { const int b = r - NSM_BANDS - ni; u = (double) (L - 1) + (double) nsm_band_lo(b); v = (double) (L - 1) + (double) nsm_band_hi(b); }
}

// region of coordinate q on an axis with L ticks; -1: outside every band (or NaN)
NDP_HD int nsm_region_of(int S, int L, double q)
{
    if (!(q == q)) return -1;
    const double top = (double) (L - 1);
    if (q > 0.0 && q < top) return NSM_BANDS + (int) (q * (double) S);
    const bool low = !(q > 0.0);
    double t = low ? -q : q - top;
    if (!(t < 64.0)) return -1;
    const int it = (int) t;
    // 0 -> 0, 1 -> 1, 2..3 -> 2, 4..7 -> 3, ...
#if defined(__CUDA_ARCH__)
    const int band = 32 - __clz(it);
#else
    const int band = it ? 32 - __builtin_clz((unsigned) it) : 0;
#endif
    if (band >= NSM_BANDS) return -1;
    return low ? NSM_BANDS - 1 - band : NSM_BANDS + S * (L - 1) + band;
}

NDP_HD unsigned long long nsm_pack(const int *c, int nvar)
{
    unsigned long long p = 0;
    for (int j = 0; j < nvar; j++) p |= (unsigned long long) (unsigned) c[j] << (16 * j);
    return p;
}
NDP_HD int nsm_coord(unsigned long long p, int j) { return (int) ((p >> (16 * j)) & 0xffffu); }

// inf over region box of [d(q, a) - d(q, b)] (variant axes only; the other axes cancel)
NDP_HD double nsm_gap(const NdpSiteMap &M, const double *u, const double *v, unsigned long long a, unsigned long long b)
{
    const int mode = nsm_mode_of_geo(M.geo);
    double sum = 0.0;
    for (int j = 0; j < M.nvar; j++) {
        const int ca = nsm_coord(a, j), cb = nsm_coord(b, j);
        if (ca == cb) continue;
        const double gu = ndp_axis_term(mode, u[j], ca, ca) - ndp_axis_term(mode, u[j], cb, cb);
        const double gv = ndp_axis_term(mode, v[j], ca, ca) - ndp_axis_term(mode, v[j], cb, cb);
        sum += gu < gv ? gu : gv;
    }
    return sum;
}

// Candidates of one region (see the header comment).  cand[] receives packed site coordinates;
// returns their number, or -1 when the region holds more than NSM_KMAX (it is then left empty and
// its queries take the lattice search).
NDP_HD_CALL int nsm_region_candidates(const NdpSiteMap &M, int region, unsigned long long *cand)
{
    const int nv = M.nvar;
    const int mode = nsm_mode_of_geo(M.geo);
    double u[NSM_MAX_VAR], v[NSM_MAX_VAR], m[NSM_MAX_VAR];
    int c0[NSM_MAX_VAR];
    double rho2 = 0.0;
    int rem = region;
    for (int j = 0; j < nv; j++) {
        const int r = rem / M.rstride[j];
        rem -= r * M.rstride[j];
        nsm_interval(M.S, M.vlen[j], r, u[j], v[j]);
        m[j] = 0.5 * (u[j] + v[j]);
        const double h = 0.5 * (v[j] - u[j]);
        rho2 += h * h;
        double p = (M.geo == NSM_GEO_VERTEX) ? rint(m[j]) : floor(m[j]) + 1.0;
        const double hi = (double) (M.vlen[j] - 1);
        p = p < (double) M.lo ? (double) M.lo : (p > hi ? hi : p);
        c0[j] = (int) p;
    }
    if (nv == 0) { cand[0] = 0; return ndp_bit(M.slice, 0) ? 1 : 0; }

    // an upper bound U on the distance from the region's centre to its nearest site: best site of the
    // smallest doubling window around the centre that holds one
    int maxlen = 0;
    for (int j = 0; j < nv; j++) maxlen = M.vlen[j] > maxlen ? M.vlen[j] : maxlen;
    double U = -1.0;
    unsigned long long s0 = 0;
    for (int W = 1;; W *= 2) {
        int c[NSM_MAX_VAR], lo[NSM_MAX_VAR], hi[NSM_MAX_VAR];
        for (int j = 0; j < nv; j++) {
            lo[j] = c0[j] - W < M.lo ? M.lo : c0[j] - W;
            hi[j] = c0[j] + W > M.vlen[j] - 1 ? M.vlen[j] - 1 : c0[j] + W;
            c[j] = lo[j];
        }
        for (;;) {
            int flat = 0;
            for (int j = 0; j < nv; j++) flat += c[j] * M.sstride[j];
            if (ndp_bit(M.slice, flat)) {
                double d = 0.0;
                for (int j = 0; j < nv; j++) d += ndp_axis_term(mode, m[j], c[j], c[j]);
                if (U < 0.0 || d < U) { U = d; s0 = nsm_pack(c, nv); }
            }
            int j = nv - 1;
            while (j >= 0) { if (++c[j] <= hi[j]) break; c[j] = lo[j]; j--; }
            if (j < 0) break;
        }
        if (U >= 0.0 || W >= maxlen) break;
    }
    if (U < 0.0) return 0;                          // empty mask

    // every site that is not farther than s0 by more than NSM_EPS somewhere in the region lies within
    // Rw ticks of the centre on every axis (distances are 1-Lipschitz; +2 covers the cell metrics' offsets)
    const double Rw = sqrt(U) + 2.0 * sqrt(rho2) + 2.0;
    int n = 0;
    {
        int c[NSM_MAX_VAR], lo[NSM_MAX_VAR], hi[NSM_MAX_VAR];
        for (int j = 0; j < nv; j++) {
            const double a = floor(m[j] - Rw), b = ceil(m[j] + Rw);
            lo[j] = a < (double) M.lo ? M.lo : (int) a;
            hi[j] = b > (double) (M.vlen[j] - 1) ? M.vlen[j] - 1 : (int) b;
            c[j] = lo[j];
        }
        for (;;) {
            int flat = 0;
            for (int j = 0; j < nv; j++) flat += c[j] * M.sstride[j];
            if (ndp_bit(M.slice, flat)) {
                const unsigned long long s = nsm_pack(c, nv);
                if (s == s0 || !(nsm_gap(M, u, v, s, s0) > NSM_EPS)) {
                    if (n == NSM_KMAX) return -1;
                    cand[n++] = s;
                }
            }
            int j = nv - 1;
            while (j >= 0) { if (++c[j] <= hi[j]) break; c[j] = lo[j]; j--; }
            if (j < 0) break;
        }
    }
    // pairwise pruning: a site beaten everywhere in the region by another survivor of the first pass.
    // (Marks are only read after the sweep: the dominator of a dropped site may itself be dropped, by a
    // site that is closer still -- the chain ends at a kept site because every step gains NSM_EPS.)
    unsigned long long drop[(NSM_KMAX + 63) / 64];
    for (int w = 0; w < (NSM_KMAX + 63) / 64; w++) drop[w] = 0;
    for (int i = 0; i < n; i++)
        for (int k = 0; k < n; k++)
            if (k != i && nsm_gap(M, u, v, cand[i], cand[k]) > NSM_EPS) { drop[i >> 6] |= 1ull << (i & 63); break; }
    int kept = 0;
    for (int i = 0; i < n; i++)
        if (!((drop[i >> 6] >> (i & 63)) & 1ull)) cand[kept++] = cand[i];
    return kept;
}

// Writes one region's entries (ew words each) at ents + offs[region] * ew.
NDP_HD void nsm_write_entries(const NdpSiteMap &M, const unsigned long long *cand, int n, uint32_t *dst)
{
    for (int i = 0; i < n; i++)
        for (int w = 0; w < M.ew; w++) dst[i * M.ew + w] = (uint32_t) (cand[i] >> (32 * w));
}

// Words per region of the head table: ew == 1: c0, c1, count, beg (16 bytes); ew == 2: c0 (2), c1 (2), count, beg, 0, 0 (32 bytes).
NDP_HD int nsm_head_words(int ew) { return ew == 1 ? 4 : 8; }

// Writes one region's head: first two candidates, count, start of the region's list in ents.
NDP_HD void nsm_write_head(const NdpSiteMap &M, const unsigned long long *cand, int n, int beg, uint32_t *dst)
{
    const unsigned long long c0 = n > 0 ? cand[0] : 0ull, c1 = n > 1 ? cand[1] : 0ull;
    if (M.ew == 1) { dst[0] = (uint32_t) c0; dst[1] = (uint32_t) c1; dst[2] = (uint32_t) n; dst[3] = (uint32_t) beg; }
    else {
        dst[0] = (uint32_t) c0; dst[1] = (uint32_t) (c0 >> 32); dst[2] = (uint32_t) c1; dst[3] = (uint32_t) (c1 >> 32);
        dst[4] = (uint32_t) n; dst[5] = (uint32_t) beg; dst[6] = 0u; dst[7] = 0u;
    }
}

// int in [0, 2^31) -> double, exactly.  On the device: a bit pattern + one DADD instead of I2F.F64
// (which runs on the slow conversion pipe); both give the same value.
NDP_HD double nsm_i2d(int c)
{
#if defined(__CUDA_ARCH__)
    return __hiloint2double(0x43300000, c) - 4503599627370496.0;               // (2^52 + c) - 2^52
#else
    return (double) c;
#endif
}

// position of site coordinate c on a point-metric axis: c (vertex) or c - 0.5 (cell centre); exact
template <bool CENTRE>
NDP_HD double nsm_pos(int c)
{
#if defined(__CUDA_ARCH__)
    return CENTRE ? __hiloint2double(0x43200000, c << 1) - 2251799813685248.5  // (2^51 + c) - (2^51 + 0.5)
                  : __hiloint2double(0x43300000, c) - 4503599627370496.0;
#else
    return CENTRE ? (double) c - 0.5 : (double) c;
#endif
}

// ndp_axis_term(mode, q, c, c) for a non-NaN q, without its branches where the metric is a point metric:
// all three of its cases are (pos - q)^2 there (pos - q is +0 when they are equal).
template <int MODE>
NDP_HD double nsm_term(double q, int c)
{
    if (MODE == NDP_MODE_LS_LINEAR) return ndp_axis_term(MODE, q, c, c);
    const double d = nsm_pos<MODE == NDP_MODE_KD_LINEAR>(c) - q;
    return d * d;
}

// Compile-time-unrolled twins of ndp_make_query / ndp_add_assoc / ndp_reported_dist (ndp_core.cuh): every
// array index is a constant, so the per-query state stays in registers inside the fused kernel.
template <int N>
NDP_HD void ndp_make_query_n(const NdpGeom &g, const int *idx, const double *nrm, NdpQuery &Q)
{
    if (N == 0) { ndp_make_query(g, idx, nrm, Q); return; }
#pragma unroll
    for (int a = 0; a < (N ? N : 1); a++) {
        Q.qc[a] = nsm_i2d(idx[a]) + nrm[a] - 1.0;             // idx >= 1 (:226, :250)
        Q.acoord[a] = (a >= g.nbasic) ? ndp_assoc_coord(Q.qc[a], g.len[a]) : 0;
    }
}

template <int N, int MODE>
NDP_HD double ndp_add_assoc_n(const NdpGeom &g, const NdpQuery &Q, double acc)
{
    if (N == 0) return ndp_add_assoc(g, MODE, Q, acc);
    if (MODE >= NDP_MODE_LS_NEAREST) {
#pragma unroll
        for (int a = 0; a < (N ? N : 1); a++)
            if (a >= g.nbasic) {
                const double diff = Q.qc[a] - nsm_i2d(Q.acoord[a]);
                acc += diff * diff;
            }
    }
    return acc;
}

// coords[] are valid lattice coordinates (>= 0; >= 1 on the basic axes for LINEAR)
template <int N, int METHOD>
NDP_HD double ndp_reported_dist_n(const NdpGeom &g, const NdpQuery &Q, const int *coords)
{
    if (N == 0) return ndp_reported_dist(g, METHOD, Q, coords);
    double acc = 0.0;
#pragma unroll
    for (int a = 0; a < (N ? N : 1); a++) {
        const double q = Q.qc[a], hi = nsm_i2d(coords[a]);
        double diff;
        if (METHOD == NDP_M_NEAREST || a >= g.nbasic) diff = q - hi;            // :245-254, :179-183
        else {                                                                  // ndp_distance, :158-176: cell [c-1, c]
            const double lo = hi - 1.0;
            diff = q < lo ? lo - q : (q > hi ? q - hi : 0.0);
        }
        acc += diff * diff;
    }
    return acc;
}

// ---------------------------------------------------------------------------
// Query time.  N: compile-time axis count (0 = run time); MODE: search mode; VS: 0 = the variant axes are
// read from the map at run time, 1 = the caller guarantees they are exactly axes {0, 1} with one-word entries
// (the usual shape of tables whose voids depend on their two leading axes: no per-axis tests in the loop).
// Returns true when the map settles the query: hit.coords (basic axes) and hit.dsel are then exactly
// what the reference's search returns (hit.vertex is filled for the full-scan modes only, whose tie rule
// needs it).  false: take the lattice search.
// ---------------------------------------------------------------------------
// What the probe leaves for the search: the region's head, loaded as early as the plan allows so that its L2
// round trip is covered by the classification and the closed-form axes.
struct NdpNsmProbe {
    unsigned long long cand0, cand1;
    int count, beg;
    bool ok;                           // no NaN on a basic axis, every variant axis inside the map's bands
};

// Region of the query on the variant axes + the load of the region's head.  Valid for ANY imported query (the
// region index is clamped into the map), so the plan issues it before it knows whether the query needs a search.
template <int N, int MODE, int VS = 0>
NDP_HD void ndp_nsm_probe(const NdpGeom &g, const NdpSiteMap &M, const NdpQuery &Q, NdpNsmProbe &pr)
{
    constexpr int MM = N ? N : NDP_MAX_AXES;
    const int nb = g.nbasic;
    int region = 0;
    bool ok = true;
#pragma unroll
    for (int a = 0; a < MM; a++) {
        if (a < nb) {
            const bool nan = !(Q.qc[a] == Q.qc[a]);
            ok &= !nan;
            const int j = VS ? (a < 2 ? a : -1) : M.jof[a];
            if (j >= 0) {
                const int r = nsm_region_of(M.S, g.len[a], nan ? 0.0 : Q.qc[a]);
                ok &= r >= 0;
                region += (r < 0 ? 0 : r) * M.rstride[j];
            }
        }
    }
    pr.ok = ok;
    // one 16-byte load (two for two-word entries) gives the region's first two candidates, its candidate count and the
    // start of its list: regions with at most two candidates -- nearly all -- need no dependent load
    if (VS || M.ew == 1) {
#if defined(__CUDA_ARCH__)
        const ulonglong2 hv = *reinterpret_cast<const ulonglong2 *>(M.head + (long long) region * 4);     // one 128-bit load
        const unsigned long long h0 = hv.x, h1 = hv.y;
#else
        const unsigned long long *hp = reinterpret_cast<const unsigned long long *>(M.head + (long long) region * 4);
        const unsigned long long h0 = hp[0], h1 = hp[1];
#endif
        pr.cand0 = h0 & 0xffffffffull; pr.cand1 = h0 >> 32; pr.count = (int) (h1 & 0xffffffffull); pr.beg = (int) (h1 >> 32);
    } else {
#if defined(__CUDA_ARCH__)
        const ulonglong2 *hq = reinterpret_cast<const ulonglong2 *>(M.head + (long long) region * 8);     // two 128-bit loads
        const ulonglong2 ha = hq[0], hb = hq[1];
        pr.cand0 = ha.x; pr.cand1 = ha.y;
        const unsigned long long h2 = hb.x;
#else
        const unsigned long long *hp = reinterpret_cast<const unsigned long long *>(M.head + (long long) region * 8);
        pr.cand0 = hp[0]; pr.cand1 = hp[1];
        const unsigned long long h2 = hp[2];
#endif
        pr.count = (int) (h2 & 0xffffffffull); pr.beg = (int) (h2 >> 32);
    }
}

template <int N, int MODE, int VS = 0>
NDP_HD bool ndp_nsm_search(const NdpGeom &g, const NdpSiteMap &M, const NdpQuery &Q, const NdpNsmProbe &pr, NdpHit &hit)
{
    constexpr int MM = N ? N : NDP_MAX_AXES;
    constexpr int mode = MODE;
    constexpr bool cells = (mode == NDP_MODE_KD_LINEAR || mode == NDP_MODE_LS_LINEAR);
    constexpr bool kd = mode < NDP_MODE_LS_NEAREST;
    constexpr bool point = mode != NDP_MODE_LS_LINEAR;
    constexpr int lo = cells ? 1 : 0;
    const int nb = g.nbasic;
    double t0[MM];                     // closed-form axes: the term of the chosen coordinate (0 elsewhere)
    int fix[MM];
    int base = 0;
    bool ok = pr.ok;
    bool wide = true;                  // every closed-form axis has the NSM_EPS margin (else only NSM_EPS_NEAR)
    double qa[MM];
#pragma unroll
    for (int a = 0; a < MM; a++) {
        t0[a] = 0.0; fix[a] = 0;
        qa[a] = (a < nb && Q.qc[a] == Q.qc[a]) ? Q.qc[a] : 0.0;
    }
    const unsigned long long cand0 = pr.cand0, cand1 = pr.cand1;
    const int count = pr.count, beg = pr.beg;
#pragma unroll
    for (int a = 0; a < MM; a++) {
        if (a < nb) {
            const double q = qa[a];
            const int j = VS ? (a < 2 ? a : -1) : M.jof[a];
            if (j < 0) {
                // the mask does not depend on this axis: the winner takes the coordinate that minimises the
                // axis' own term, provided both neighbours lose by a margin no rounding can bridge
                const double hi = (double) (g.len[a] - 1);
                double p = cells ? floor(q) + 1.0 : rint(q);
                p = p < (double) lo ? (double) lo : (p > hi ? hi : p);
                const int c = (int) p;
                if (point) {
                    // (pos - q)^2 with pos = p or p - 0.5.  The neighbours at pos -+ 1 score 1 -+ 2 d0 more: a
                    // margin unless |d0| is within the tolerance of 1/2.  (|d0| > 1/2 only where p was clamped to the
                    // end of the axis, and then the one neighbour that exists lies on the far side.)
                    // The margin must exceed twice the rounding error of the ordered totals the reference compares:
                    // NSM_EPS whatever the total (< NSM_TMAX), NSM_EPS_NEAR when the winner's total is below NSM_TNEAR.
                    const double d0 = (cells ? p - 0.5 : p) - q;
                    t0[a] = d0 * d0;
                    const double m = fabs(fabs(d0) - 0.5);
                    ok &= m > NSM_EPS_NEAR;
                    wide &= m > NSM_EPS;
                } else {
                    const double t = ndp_axis_term(mode, q, c, c);
                    const double tm = (c - 1 >= lo) ? ndp_axis_term(mode, q, c - 1, c - 1) : 4.0 * NSM_TMAX;
                    const double tp = (c + 1 <= g.len[a] - 1) ? ndp_axis_term(mode, q, c + 1, c + 1) : 4.0 * NSM_TMAX;
                    ok &= (tm - t > NSM_EPS) && (tp - t > NSM_EPS);
                    t0[a] = t;
                }
                fix[a] = c;
                if (!kd) base += c * g.bstride[a];
            }
        }
    }
    if (!ok) return false;
    int best = 0x7fffffff;
    double bestd = 8.0 * NSM_TMAX;
    unsigned long long bestp = 0;
    bool tie = false;
    const uint32_t *ep = M.ents + (long long) (beg + 2) * (VS ? 1 : M.ew);
    for (int e = 0; e < count; e++) {
        unsigned long long p = e == 0 ? cand0 : cand1;
        if (e > 1) {
            p = ep[0];
            if (!VS && M.ew > 1) p |= (unsigned long long) ep[1] << 32;
            ep += VS ? 1 : M.ew;
        }
        // the reference's sum, in its axis order: closed-form axes contribute their fixed term (associated axes
        // and axes beyond nbasic add +0.0, which changes nothing)
        double d = 0.0;
        int vid = base;
#pragma unroll
        for (int a = 0; a < MM; a++) {
            const int j = VS ? (a < 2 ? a : -1) : M.jof[a];
            if (j < 0) d += t0[a];
            else {
                const int c = nsm_coord(p, j);
                d += nsm_term<MODE>(Q.qc[a], c);
                if (!kd) vid += c * g.bstride[a];
            }
        }
        d = ndp_add_assoc_n<N, MODE>(g, Q, d);
        if (d < bestd) { best = vid; bestd = d; bestp = p; tie = false; }
        else if (d == bestd) {
            if (kd) tie = true;                                     // the reference's traversal order decides
            else if (vid < best) { best = vid; bestp = p; }         // _compare_indexed_dists, :101-112
        }
    }
    if (tie || !(bestd < NSM_TMAX)) return false;                   // also: no candidate stored for this region
    if (!wide && !(bestd < NSM_TNEAR)) return false;
    hit.vertex = kd ? -1 : best; hit.dsel = bestd; hit.status = NDP_SEARCH_OK; hit.hard = 0; hit.has_coords = 1;
#pragma unroll
    for (int a = 0; a < MM; a++)
        if (a < nb) { const int j = VS ? (a < 2 ? a : -1) : M.jof[a]; hit.coords[a] = j < 0 ? fix[a] : nsm_coord(bestp, j); }
    return true;
}

// true when a map qualifies for the VS = 1 specialisation of ndp_nsm_search
inline bool nsm_leading_pair(const NdpSiteMap &M)
{
    return M.valid && M.nvar == 2 && M.var_axis[0] == 0 && M.var_axis[1] == 1 && M.ew == 1;
}
