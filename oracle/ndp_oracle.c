/*
 * ndp_oracle.c -- CPU restatement of the batched `ndpolate` hot path of
 * aprsa/ndpolator v1.3.0.
 *
 * THIS FILE IS TEST INFRASTRUCTURE.  It is the parity checker for the CUDA
 * path in ndpolator_b200/csrc/.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product
 * library (libndpb200.so) never links, loads or calls anything in oracle/.
 *
 * Parity status: PINNED.  tests/test_oracle_cpu.py checks this restatement
 *   (a) against the known-answer vectors of the reference's own tests
 *       (tests/test_interp.py, test_extrapolation.py, test_ndp_distance.py,
 *       test_kdtree_integration.py of the reference), restated in
 *       tests/golden/, and
 *   (b) bit-for-bit against the reference C sources compiled unmodified into
 *       oracle/_ref/libndp_ref.so (oracle/Makefile), on random + adversarial
 *       inputs, whenever that library is present.
 *
 * It is a restatement, not a copy: flat SoA arrays, 64-bit positions, an
 * index-linked kd-tree, and a running arg-min instead of a full sort.  Every
 * function cites the reference lines whose behaviour it reproduces
 * (paths relative to the reference root).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off; no FMA contraction so
 * that every double rounding matches the reference's x86-64 -O3 build).
 */

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NDPO_MAX_AXES 16

enum { M_NONE = 0, M_NEAREST = 1, M_LINEAR = 2 };   /* src/ndpolator.h:28-32 */
enum { S_KDTREE = 0, S_LINEAR = 1 };                /* src/ndpolator.h:44-47 */
enum { F_ON_VERTEX = 1, F_OUT_OF_BOUNDS = 2 };      /* src/ndpolator.h:59-63 */

/* ------------------------------------------------------------------ */
/* index-linked kd-tree with the reference's insertion-order topology  */
/* ------------------------------------------------------------------ */

typedef struct {
    int k;              /* dimensionality (= nbasic) */
    int64_t n, cap;     /* nodes used / allocated */
    double *pos;        /* n*k coordinates */
    int64_t *payload;   /* vertex id */
    int32_t *lo, *hi;   /* children: "negative" / "positive" side, -1 = none */
    double bbmin[NDPO_MAX_AXES], bbmax[NDPO_MAX_AXES];  /* bounding box of all inserted points */
} okd_tree;

static okd_tree *okd_new(int k, int64_t cap)
{
    okd_tree *t = calloc(1, sizeof *t);
    t->k = k;
    t->cap = cap > 0 ? cap : 1;
    t->pos = malloc(sizeof(double) * t->cap * k);
    t->payload = malloc(sizeof(int64_t) * t->cap);
    t->lo = malloc(sizeof(int32_t) * t->cap);
    t->hi = malloc(sizeof(int32_t) * t->cap);
    return t;
}

static void okd_free(okd_tree *t)
{
    if (!t) return;
    free(t->pos); free(t->payload); free(t->lo); free(t->hi); free(t);
}

/* external/kdtree/kdtree.c:167-209 (insert_rec + kd_insert): the split axis
 * cycles with depth starting at 0; strictly-smaller goes to the negative side,
 * everything else to the positive side; the bounding box grows with each point. */
static void okd_insert(okd_tree *t, const double *p, int64_t payload)
{
    int64_t id = t->n++;
    memcpy(t->pos + id * t->k, p, sizeof(double) * t->k);
    t->payload[id] = payload;
    t->lo[id] = t->hi[id] = -1;

    if (id == 0) {
        for (int a = 0; a < t->k; a++) t->bbmin[a] = t->bbmax[a] = p[a];
        return;
    }
    for (int a = 0; a < t->k; a++) {
        if (p[a] < t->bbmin[a]) t->bbmin[a] = p[a];
        if (p[a] > t->bbmax[a]) t->bbmax[a] = p[a];
    }

    int64_t cur = 0;
    int depth = 0;
    for (;;) {
        int ax = depth % t->k;
        int32_t *slot = (p[ax] < t->pos[cur * t->k + ax]) ? &t->lo[cur] : &t->hi[cur];
        if (*slot < 0) { *slot = (int32_t) id; return; }
        cur = *slot;
        depth++;
    }
}

/* external/kdtree/kdtree.c:745-759 (hyperrect_dist_sq) */
static double okd_box_dist2(const okd_tree *t, const double *bmin, const double *bmax, const double *q)
{
    double r = 0;
    for (int a = 0; a < t->k; a++) {
        if (q[a] < bmin[a]) r += (bmin[a] - q[a]) * (bmin[a] - q[a]);
        else if (q[a] > bmax[a]) r += (bmax[a] - q[a]) * (bmax[a] - q[a]);
    }
    return r;
}

/* external/kdtree/kdtree.c:345-402 (kd_nearest_i): nearer child first with the
 * box sliced at the node, then the node itself (strict <), then the farther
 * child iff the sliced box is strictly closer than the best so far. */
static void okd_visit(const okd_tree *t, int64_t node, int depth, const double *q,
                      int64_t *best, double *best_d2, double *bmin, double *bmax)
{
    int ax = depth % t->k;
    const double *np = t->pos + node * t->k;
    double side = q[ax] - np[ax];
    int32_t nearer, farther;
    double *near_edge, *far_edge;

    if (side <= 0) { nearer = t->lo[node]; farther = t->hi[node]; near_edge = bmax + ax; far_edge = bmin + ax; }
    else           { nearer = t->hi[node]; farther = t->lo[node]; near_edge = bmin + ax; far_edge = bmax + ax; }

    if (nearer >= 0) {
        double keep = *near_edge;
        *near_edge = np[ax];
        okd_visit(t, nearer, depth + 1, q, best, best_d2, bmin, bmax);
        *near_edge = keep;
    }

    double d2 = 0;
    for (int a = 0; a < t->k; a++) d2 += (np[a] - q[a]) * (np[a] - q[a]);
    if (d2 < *best_d2) { *best = node; *best_d2 = d2; }

    if (farther >= 0) {
        double keep = *far_edge;
        *far_edge = np[ax];
        if (okd_box_dist2(t, bmin, bmax, q) < *best_d2)
            okd_visit(t, farther, depth + 1, q, best, best_d2, bmin, bmax);
        *far_edge = keep;
    }
}

/* external/kdtree/kdtree.c:404-457 (kd_nearest): the root seeds the best guess. */
static int64_t okd_nearest(const okd_tree *t, const double *q)
{
    double bmin[NDPO_MAX_AXES], bmax[NDPO_MAX_AXES];
    if (t->n == 0) return -1;
    memcpy(bmin, t->bbmin, sizeof(double) * t->k);
    memcpy(bmax, t->bbmax, sizeof(double) * t->k);
    int64_t best = 0;
    double d2 = 0;
    for (int a = 0; a < t->k; a++) d2 += (t->pos[a] - q[a]) * (t->pos[a] - q[a]);
    okd_visit(t, 0, 0, q, &best, &d2, bmin, bmax);
    return t->payload[best];
}

/* ------------------------------------------------------------------ */
/* table                                                                */
/* ------------------------------------------------------------------ */

typedef struct {
    int naxes, nbasic, vdim;
    int len[NDPO_MAX_AXES];
    double *tick[NDPO_MAX_AXES];
    int64_t vstride[NDPO_MAX_AXES];  /* vertices (not doubles) skipped per +1 on axis a, all axes: ndp_types.c:59-65 */
    int64_t bstride[NDPO_MAX_AXES];  /* same but in the basic-only lattice */
    const double *grid;              /* borrowed */
    int64_t nverts;                  /* basic lattice size */
    unsigned char *vmask, *hcmask;
    okd_tree *tree[3];               /* per extrapolation method, lazily built */
} ndpo_table;

static inline int64_t basic_coord(const ndpo_table *t, int64_t v, int a)
{
    return (v / t->bstride[a]) % t->len[a];
}

/* src/ndp_types.c:159-269 (ndp_table_new_from_data): strides, vertex mask
 * (component 0 at associated index 0 is not NaN) and hypercube mask (vertex is
 * the superior corner of a cell all of whose 2^nbasic corners are defined). */
void *ndpo_table_new(int naxes, int nbasic, const int *axlen, const double *ticks, int vdim, const double *grid)
{
    if (naxes < 1 || naxes > NDPO_MAX_AXES || nbasic < 1 || nbasic > naxes) return NULL;
    ndpo_table *t = calloc(1, sizeof *t);
    t->naxes = naxes; t->nbasic = nbasic; t->vdim = vdim; t->grid = grid;

    const double *src = ticks;
    for (int a = 0; a < naxes; a++) {
        t->len[a] = axlen[a];
        t->tick[a] = malloc(sizeof(double) * axlen[a]);
        memcpy(t->tick[a], src, sizeof(double) * axlen[a]);
        src += axlen[a];
    }
    for (int a = naxes - 1; a >= 0; a--)
        t->vstride[a] = (a == naxes - 1) ? 1 : t->vstride[a + 1] * t->len[a + 1];
    for (int a = nbasic - 1; a >= 0; a--)
        t->bstride[a] = (a == nbasic - 1) ? 1 : t->bstride[a + 1] * t->len[a + 1];

    t->nverts = 1;
    for (int a = 0; a < nbasic; a++) t->nverts *= t->len[a];

    if (!grid) return t;   /* axes-only table: import is the only legal call */

    t->vmask = calloc(t->nverts, 1);
    t->hcmask = calloc(t->nverts, 1);
    int64_t assoc_block = t->vstride[nbasic - 1] * vdim;   /* doubles per basic vertex */
    for (int64_t v = 0; v < t->nverts; v++) {
        double first = grid[v * assoc_block];
        t->vmask[v] = (first == first);
    }
    for (int64_t v = 0; v < t->nverts; v++) {
        if (!t->vmask[v]) continue;
        int ok = 1;
        for (int a = 0; a < nbasic && ok; a++)
            if (basic_coord(t, v, a) == 0) ok = 0;
        for (int corner = 0; corner < (1 << nbasic) && ok; corner++) {
            int64_t w = v;
            for (int a = 0; a < nbasic; a++)
                if (!((corner >> (nbasic - a - 1)) & 1)) w -= t->bstride[a];
            if (!t->vmask[w]) ok = 0;
        }
        t->hcmask[v] = (unsigned char) ok;
    }
    return t;
}

void ndpo_table_free(void *h)
{
    ndpo_table *t = h;
    if (!t) return;
    for (int a = 0; a < t->naxes; a++) free(t->tick[a]);
    free(t->vmask); free(t->hcmask);
    for (int m = 0; m < 3; m++) okd_free(t->tree[m]);
    free(t);
}

int64_t ndpo_table_nverts(void *h) { return ((ndpo_table *) h)->nverts; }

void ndpo_table_masks(void *h, int32_t *vmask, int32_t *hcmask)
{
    ndpo_table *t = h;
    for (int64_t v = 0; v < t->nverts; v++) { vmask[v] = t->vmask[v]; hcmask[v] = t->hcmask[v]; }
}

/* src/ndpolator.c:114-136 (ndp_kdtree_create): masked vertices in ascending
 * flat order at their integer basic coordinates, shifted by -0.5 (cell centre)
 * for the LINEAR method; the payload is the vertex id. */
static okd_tree *build_tree(const ndpo_table *t, int method)
{
    const unsigned char *mask = (method == M_NEAREST) ? t->vmask : t->hcmask;
    int64_t count = 0;
    for (int64_t v = 0; v < t->nverts; v++) count += mask[v];
    okd_tree *tree = okd_new(t->nbasic, count);
    double p[NDPO_MAX_AXES];
    for (int64_t v = 0; v < t->nverts; v++) {
        if (!mask[v]) continue;
        for (int a = 0; a < t->nbasic; a++) {
            p[a] = (double) basic_coord(t, v, a);
            if (method == M_LINEAR) p[a] -= 0.5;
        }
        okd_insert(tree, p, v);
    }
    return tree;
}

/* Builds the lazily created trees up front (so that timed / threaded callers
 * never race on the lazy build, src/ndpolator.c:218-221). */
void ndpo_table_warm(void *h, int method)
{
    ndpo_table *t = h;
    if ((method == M_NEAREST || method == M_LINEAR) && !t->tree[method])
        t->tree[method] = build_tree(t, method);
}

/* ------------------------------------------------------------------ */
/* A1: import  (src/ndpolator.c:21-45 find_first_geq_than, :354-414)    */
/* ------------------------------------------------------------------ */

static void import_one(const double *tick, int len, double x, int32_t *index, int32_t *flag, double *normed)
{
    const double rtol = 1e-6;                       /* src/ndpolator.c:358 */
    int lo = 1, hi = len - 1;
    while (lo != hi) {                              /* :24-33, absolute +rtol in the search */
        int mid = lo + (hi - lo) / 2;
        if (x + rtol > tick[mid]) lo = mid + 1; else hi = mid;
    }
    int f = (x < tick[0] || x > tick[len - 1]) ? F_OUT_OF_BOUNDS : 0;   /* :35 */
    double t = (x - tick[lo - 1]) / (tick[lo] - tick[lo - 1]);          /* :37, :379-381 */
    if (fabs(t) < rtol || (lo == len - 1 && fabs(t - 1) < rtol)) f |= F_ON_VERTEX;  /* :37-39 */
    *index = lo; *flag = f; *normed = t;
}

void ndpo_import(void *h, int64_t n, const double *q, int32_t *indices, int32_t *flags, double *normed)
{
    ndpo_table *t = h;
    for (int64_t i = 0; i < n; i++)
        for (int a = 0; a < t->naxes; a++) {
            int64_t k = i * t->naxes + a;
            import_one(t->tick[a], t->len[a], q[k], &indices[k], &flags[k], &normed[k]);
        }
}

/* ------------------------------------------------------------------ */
/* A2: hypercube assembly (src/ndpolator.c:416-508)                     */
/* ------------------------------------------------------------------ */

/* Fills v[0 .. 2^dim * vdim) and returns dim; *fdhc as the reference computes it. */
static int hypercube_one(const ndpo_table *t, const int32_t *idx, const int32_t *flg, const double *nrm,
                         double *v, int *fdhc)
{
    int n = t->naxes, V = t->vdim;
    int full = 1, dim = n;
    for (int a = 0; a < n; a++) {
        if (flg[a] & F_OUT_OF_BOUNDS) full = 0;        /* :441-445 */
        if (flg[a] & F_ON_VERTEX) dim--;               /* :455-460 */
    }
    for (int corner = 0; corner < (1 << dim); corner++) {
        int64_t vert = 0;
        int freeax = 0;
        for (int a = 0; a < n; a++) {
            int c;
            if (flg[a] & F_ON_VERTEX) {
                c = nrm[a] > 0.5 ? idx[a] : idx[a] - 1;   /* :474-478 */
            } else {
                int bit = (corner >> (dim - freeax - 1)) & 1;   /* first free axis = MSB, :480 */
                c = idx[a] - 1 + bit;
                if (c < bit) c = bit;
                freeax++;
            }
            vert += t->vstride[a] * c;
        }
        const double *src = t->grid + vert * V;        /* ndp_idx2pos, :47-55 */
        if (src[0] != src[0]) full = 0;                /* NaN test on component 0 only, :491 */
        memcpy(v + (int64_t) corner * V, src, sizeof(double) * V);
    }
    *fdhc = full;
    return dim;
}

/* v is n * 2^naxes * vdim doubles; each query's slot is filled from its start. */
void ndpo_hypercubes(void *h, int64_t n, const int32_t *indices, const int32_t *flags, const double *normed,
                     int32_t *dim, int32_t *fdhc, double *v)
{
    ndpo_table *t = h;
    int64_t slot = ((int64_t) 1 << t->naxes) * t->vdim;
    for (int64_t i = 0; i < n; i++) {
        int f;
        dim[i] = hypercube_one(t, indices + i * t->naxes, flags + i * t->naxes, normed + i * t->naxes, v + i * slot, &f);
        fdhc[i] = f;
    }
}

/* ------------------------------------------------------------------ */
/* A3: successive-dimension reduction (src/ndpolator.c:78-99)           */
/* ------------------------------------------------------------------ */

static void reduce(int dim, int V, const double *x, double *fv)
{
    for (int a = 0; a < dim; a++) {
        int half = 1 << (dim - a - 1);
        for (int j = 0; j < half; j++)
            for (int k = 0; k < V; k++) {
                double lo = fv[j * V + k];
                double diff = fv[(half + j) * V + k] - lo;
                double step = diff * x[a];
                fv[j * V + k] = lo + step;             /* three roundings, never an FMA */
            }
    }
}

/* ------------------------------------------------------------------ */
/* A5/A6: nearest search and distances (src/ndpolator.c:138-352)        */
/* ------------------------------------------------------------------ */

static int assoc_coord(const ndpo_table *t, int a, double qc)
{
    /* max(0, min(len-1, round(qc))) evaluated in double, then truncated: :241 */
    double r = round(qc);
    double hi = (double) (t->len[a] - 1);
    double m = hi < r ? hi : r;
    double c = 0 > m ? 0 : m;
    return (int) c;
}

/* src/ndpolator.c:138-186 (ndp_distance) */
static double cube_dist2(const ndpo_table *t, const double *nrm, const int32_t *idx, const int *coords)
{
    double acc = 0.0;
    for (int a = 0; a < t->nbasic; a++) {
        double qc = idx[a] + nrm[a] - 1.0;
        double diff;
        if (qc < coords[a] - 1.0) diff = (coords[a] - 1.0) - qc;
        else if (qc > coords[a]) diff = qc - coords[a];
        else diff = 0.0;
        acc += diff * diff;
    }
    for (int a = t->nbasic; a < t->naxes; a++) {
        double qc = idx[a] + nrm[a] - 1.0;
        double diff = qc - coords[a];
        acc += diff * diff;
    }
    return acc;
}

/* src/ndpolator.c:245-254 / :300-309 */
static double vertex_dist2(const ndpo_table *t, const double *nrm, const int32_t *idx, const int *coords)
{
    double acc = 0.0;
    for (int a = 0; a < t->naxes; a++) {
        double qc = idx[a] + nrm[a] - 1.0;
        double diff = qc - coords[a];
        acc += diff * diff;
    }
    return acc;
}

/* src/ndpolator.c:188-352 (ndp_find_nearest).  coords gets naxes entries. */
static void find_nearest_one(ndpo_table *t, const double *nrm, const int32_t *idx, int method, int search,
                             int *coords, double *dist)
{
    const unsigned char *mask = (method == M_NEAREST) ? t->vmask : t->hcmask;
    int n = t->naxes, nb = t->nbasic;

    for (int a = nb; a < n; a++)
        coords[a] = assoc_coord(t, a, idx[a] + nrm[a] - 1);           /* :240-242, :295-297, :341-342 */

    if (search == S_KDTREE) {
        if (!t->tree[method]) t->tree[method] = build_tree(t, method);   /* :218-221 */
        double qc[NDPO_MAX_AXES];
        for (int a = 0; a < nb; a++) qc[a] = idx[a] + nrm[a] - 1.0;       /* :225-226 */
        int64_t v = okd_nearest(t->tree[method], qc);
        if (v >= 0) {
            for (int a = 0; a < nb; a++) coords[a] = (int) basic_coord(t, v, a);   /* :235-237 */
            *dist = (method == M_NEAREST) ? vertex_dist2(t, nrm, idx, coords)
                                          : cube_dist2(t, nrm, idx, coords);
            return;
        }
        /* an empty tree falls through to the scan below (the reference would
         * crash in kd_res_free(NULL) first; we define the outcome as the scan's) */
    }

    /* :279-329: every basic vertex; masked-out ones score 1e10; the winner is
     * the minimum of (distance, vertex id) -- _compare_indexed_dists, :101-112 */
    int64_t best = -1;
    double best_d = 0;
    int trial[NDPO_MAX_AXES];
    for (int a = nb; a < n; a++) trial[a] = coords[a];
    for (int64_t v = 0; v < t->nverts; v++) {
        double d;
        if (!mask[v]) d = 1e10;
        else {
            for (int a = 0; a < nb; a++) trial[a] = (int) basic_coord(t, v, a);
            d = (method == M_NEAREST) ? vertex_dist2(t, nrm, idx, trial) : cube_dist2(t, nrm, idx, trial);
        }
        if (best < 0 || d < best_d) { best = v; best_d = d; }
    }
    for (int a = 0; a < nb; a++) coords[a] = (int) basic_coord(t, best, a);
    *dist = best_d;
}

void ndpo_find_nearest(void *h, int64_t n, const int32_t *indices, const int32_t *flags, const double *normed,
                       int method, int search, int32_t *coords, double *dist)
{
    ndpo_table *t = h;
    (void) flags;
    for (int64_t i = 0; i < n; i++) {
        int c[NDPO_MAX_AXES];
        find_nearest_one(t, normed + i * t->naxes, indices + i * t->naxes, method, search, c, &dist[i]);
        for (int a = 0; a < t->naxes; a++) coords[i * t->naxes + a] = c[a];
    }
}

/* ------------------------------------------------------------------ */
/* A4: the ndpolate driver (src/ndpolator.c:510-671)                    */
/* ------------------------------------------------------------------ */

int ndpo_ndpolate(void *h, int64_t n, const double *q, int method, int search, double *interps, double *dists)
{
    ndpo_table *t = h;
    int N = t->naxes, V = t->vdim;
    if (method < M_NONE || method > M_LINEAR) return 1;     /* :637-640 */
    double *cube = malloc(sizeof(double) * ((size_t) 1 << N) * V);

    for (int64_t i = 0; i < n; i++) {
        int32_t idx[NDPO_MAX_AXES], flg[NDPO_MAX_AXES];
        double nrm[NDPO_MAX_AXES], x[NDPO_MAX_AXES];
        for (int a = 0; a < N; a++)
            import_one(t->tick[a], t->len[a], q[i * N + a], &idx[a], &flg[a], &nrm[a]);

        int fdhc;
        int dim = hypercube_one(t, idx, flg, nrm, cube, &fdhc);
        dists[i] = 0.0;                                       /* calloc, :545 */

        if (fdhc) {
            int k = 0;
            for (int a = 0; a < N; a++)
                if (flg[a] != F_ON_VERTEX) x[k++] = nrm[a];   /* == 1 exactly, :648-654 */
            reduce(dim, V, x, cube);
            memcpy(interps + i * V, cube, sizeof(double) * V);
            continue;
        }

        if (method == M_NONE) {                               /* :551-554 */
            for (int k = 0; k < V; k++) interps[i * V + k] = NAN;
            continue;
        }

        int coords[NDPO_MAX_AXES];
        find_nearest_one(t, nrm, idx, method, search, coords, &dists[i]);

        if (method == M_NEAREST) {                            /* :556-567 */
            int64_t vert = 0;
            for (int a = 0; a < N; a++) vert += t->vstride[a] * coords[a];
            memcpy(interps + i * V, t->grid + vert * V, sizeof(double) * V);
            continue;
        }

        /* M_LINEAR, :569-635: the full-dimensional cell below `coords`, with
         * the query re-expressed in that cell's unit coordinates */
        for (int corner = 0; corner < (1 << N); corner++) {
            int64_t vert = 0;
            for (int a = 0; a < N; a++) {
                int bit = (corner >> (N - a - 1)) & 1;
                int c = coords[a] - 1 + bit;
                if (c < bit) c = bit;                         /* :601 */
                vert += t->vstride[a] * c;
            }
            memcpy(cube + (int64_t) corner * V, t->grid + vert * V, sizeof(double) * V);
        }
        for (int a = 0; a < N; a++) x[a] = nrm[a] + (idx[a] - coords[a]);   /* :621-622 */
        reduce(N, V, x, cube);
        memcpy(interps + i * V, cube, sizeof(double) * V);
    }
    free(cube);
    return 0;
}

/* src/ndp_py.c:196-251 (py_distance): every point, LINEAR method, kd search. */
void ndpo_distance(void *h, int64_t n, const double *q, double *dists)
{
    ndpo_table *t = h;
    int N = t->naxes;
    for (int64_t i = 0; i < n; i++) {
        int32_t idx[NDPO_MAX_AXES], flg[NDPO_MAX_AXES];
        double nrm[NDPO_MAX_AXES];
        int coords[NDPO_MAX_AXES];
        for (int a = 0; a < N; a++)
            import_one(t->tick[a], t->len[a], q[i * N + a], &idx[a], &flg[a], &nrm[a]);
        find_nearest_one(t, nrm, idx, M_LINEAR, S_KDTREE, coords, &dists[i]);
    }
}

/* Tree shape probe used by tests of the device-side topology builder:
 * parent vertex id (or -1 for the root) and depth of every tree node, indexed
 * by vertex id; entries of vertices outside the mask are left untouched. */
void ndpo_tree_topology(void *h, int method, int64_t *parent, int32_t *depth)
{
    ndpo_table *t = h;
    ndpo_table_warm(t, method);
    okd_tree *tr = t->tree[method];
    if (!tr || tr->n == 0) return;
    int64_t *stack = malloc(sizeof(int64_t) * tr->n);
    int32_t *dstack = malloc(sizeof(int32_t) * tr->n);
    int64_t sp = 0;
    parent[tr->payload[0]] = -1; depth[tr->payload[0]] = 0;
    stack[sp] = 0; dstack[sp++] = 0;
    while (sp) {
        int64_t node = stack[--sp];
        int32_t d = dstack[sp];
        int32_t kids[2] = { tr->lo[node], tr->hi[node] };
        for (int s = 0; s < 2; s++) if (kids[s] >= 0) {
            parent[tr->payload[kids[s]]] = tr->payload[node];
            depth[tr->payload[kids[s]]] = d + 1;
            stack[sp] = kids[s]; dstack[sp++] = d + 1;
        }
    }
    free(stack); free(dstack);
}
