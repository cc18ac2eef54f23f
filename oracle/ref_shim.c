/*
 * ref_shim.c -- flat-array entry points onto the UNMODIFIED reference C core.
 *
 * TEST INFRASTRUCTURE.  This file contains no ndpolator algorithm: it only
 * marshals flat arrays into the reference's own structs and calls the
 * reference's own functions (ndp_table_new_from_data, ndp_query_pts_import,
 * ndp_find_hypercubes, ndp_find_nearest, ndpolate).  oracle/Makefile compiles it
 * together with the reference sources *where they lie* under /root/reference
 * (src/ndpolator.c, src/ndp_types.c, external/kdtree/kdtree.c) into
 * oracle/_ref/libndp_ref.so.  Nothing from the reference is copied into this
 * repository; oracle/_ref/ is git-ignored.
 *
 * The entry points mirror oracle/ndp_oracle.c one for one (prefix ndpref_
 * instead of ndpo_), so the same ctypes wrapper (oracle/oracle.py) drives both
 * and tests can diff them bit for bit.
 */

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "ndpolator.h"
#include "ndp_types.h"
/* The reference's kd-tree is compiled as part of THIS translation unit, from where it lies
 * (-I$(REF)/external/kdtree): struct kdnode / struct kdtree are private to kdtree.c
 * (external/kdtree/kdtree.c:56-86), and ndpref_tree_topology below has to walk them. */
#include "kdtree.c"

/* defined in src/ndpolator.c:114 but absent from its header */
struct kdtree *ndp_kdtree_create(ndp_table *table, ndp_extrapolation_method extrapolation_method);

typedef struct {
    ndp_table *table;   /* NULL for an axes-only handle */
    ndp_axes *axes;     /* owned by table when table != NULL */
} ref_handle;

void *ndpref_table_new(int naxes, int nbasic, const int *axlen, const double *ticks, int vdim, const double *grid)
{
    ref_handle *h = calloc(1, sizeof *h);
    ndp_axis **axis = malloc(naxes * sizeof *axis);
    const double *src = ticks;
    for (int a = 0; a < naxes; a++) {
        double *copy = malloc(axlen[a] * sizeof *copy);
        memcpy(copy, src, axlen[a] * sizeof *copy);
        src += axlen[a];
        axis[a] = ndp_axis_new_from_data(axlen[a], copy, /* owns_data = */ 1);
    }
    h->axes = ndp_axes_new_from_data(naxes, nbasic, axis);
    if (grid)
        h->table = ndp_table_new_from_data(h->axes, vdim, (double *) grid, /* owns_data = */ 0);
    return h;
}

void ndpref_table_free(void *hh)
{
    ref_handle *h = hh;
    if (!h) return;
    if (h->table) ndp_table_free(h->table);   /* frees the axes too */
    else ndp_axes_free(h->axes);
    free(h);
}

int64_t ndpref_table_nverts(void *hh) { return ((ref_handle *) hh)->table->nverts; }

void ndpref_table_masks(void *hh, int32_t *vmask, int32_t *hcmask)
{
    ndp_table *t = ((ref_handle *) hh)->table;
    memcpy(vmask, t->vmask, t->nverts * sizeof(int));
    memcpy(hcmask, t->hcmask, t->nverts * sizeof(int));
}

void ndpref_table_warm(void *hh, int method)
{
    ndp_table *t = ((ref_handle *) hh)->table;
    if (method == NDP_METHOD_NEAREST && !t->vtree) t->vtree = ndp_kdtree_create(t, method);
    if (method == NDP_METHOD_LINEAR && !t->hctree) t->hctree = ndp_kdtree_create(t, method);
}

/* Shape of the reference's own tree (ndp_kdtree_create, src/ndpolator.c:114-136; kd_insert,
 * external/kdtree/kdtree.c:167-209): parent vertex id (-1 root, -2 not in the tree) and depth
 * (-1 not in the tree) per basic vertex id.  Iterative walk (the tree is 100s of levels deep);
 * reads the reference's structs, computes nothing. */
void ndpref_tree_topology(void *hh, int method, int64_t *parent, int32_t *depth)
{
    ndp_table *t = ((ref_handle *) hh)->table;
    ndpref_table_warm(hh, method);
    struct kdtree *tree = method == NDP_METHOD_NEAREST ? t->vtree : t->hctree;
    for (int i = 0; i < t->nverts; i++) { parent[i] = -2; depth[i] = -1; }
    if (!tree || !tree->root) return;
    size_t cap = 1024, top = 0;
    struct kdnode **stack = malloc(cap * sizeof *stack);
    parent[(uintptr_t) tree->root->data] = -1;
    depth[(uintptr_t) tree->root->data] = 0;
    stack[top++] = tree->root;
    while (top) {
        struct kdnode *nd = stack[--top];
        int64_t id = (int64_t) (uintptr_t) nd->data;
        struct kdnode *kids[2] = {nd->left, nd->right};
        for (int k = 0; k < 2; k++)
            if (kids[k]) {
                int64_t cid = (int64_t) (uintptr_t) kids[k]->data;
                parent[cid] = id;
                depth[cid] = depth[id] + 1;
                if (top == cap) { cap *= 2; stack = realloc(stack, cap * sizeof *stack); }
                stack[top++] = kids[k];
            }
    }
    free(stack);
}

void ndpref_import(void *hh, int64_t n, const double *q, int32_t *indices, int32_t *flags, double *normed)
{
    ref_handle *h = hh;
    int na = h->axes->len;
    ndp_query_pts *qp = ndp_query_pts_import((int) n, (double *) q, h->axes);
    memcpy(indices, qp->indices, n * na * sizeof(int));
    memcpy(flags, qp->flags, n * na * sizeof(int));
    memcpy(normed, qp->normed, n * na * sizeof(double));
    ndp_query_pts_free(qp);
}

void ndpref_hypercubes(void *hh, int64_t n, const int32_t *indices, const int32_t *flags, const double *normed,
                       int32_t *dim, int32_t *fdhc, double *v)
{
    ref_handle *h = hh;
    ndp_table *t = h->table;
    int na = h->axes->len;
    int64_t slot = ((int64_t) 1 << na) * t->vdim;
    ndp_query_pts *qp = ndp_query_pts_new_from_data((int) n, na, (int *) indices, (int *) flags, NULL, (double *) normed);
    ndp_hypercube **hc = ndp_find_hypercubes(qp, t);
    for (int64_t i = 0; i < n; i++) {
        dim[i] = hc[i]->dim;
        fdhc[i] = hc[i]->fdhc;
        memcpy(v + i * slot, hc[i]->v, sizeof(double) * ((size_t) 1 << hc[i]->dim) * t->vdim);
        ndp_hypercube_free(hc[i]);
    }
    free(hc);
    free(qp);   /* container only; the arrays belong to the caller */
}

void ndpref_find_nearest(void *hh, int64_t n, const int32_t *indices, const int32_t *flags, const double *normed,
                         int method, int search, int32_t *coords, double *dist)
{
    ref_handle *h = hh;
    int na = h->axes->len;
    for (int64_t i = 0; i < n; i++) {
        int *c = ndp_find_nearest((double *) normed + i * na, (int *) indices + i * na, (int *) flags + i * na,
                                  h->table, method, search, &dist[i]);
        memcpy(coords + i * na, c, na * sizeof(int));
        free(c);
    }
}

int ndpref_ndpolate(void *hh, int64_t n, const double *q, int method, int search, double *interps, double *dists)
{
    ref_handle *h = hh;
    ndp_table *t = h->table;
    ndp_query_pts *qp = ndp_query_pts_import((int) n, (double *) q, t->axes);
    ndp_query *res = ndpolate(qp, t, method, search);
    if (!res) { ndp_query_pts_free(qp); return 1; }
    memcpy(interps, res->interps, n * t->vdim * sizeof(double));
    memcpy(dists, res->dists, n * sizeof(double));
    ndp_query_pts_free(qp);
    ndp_query_free(res);
    return 0;
}

/* the body of the reference's py_distance loop (src/ndp_py.c:229-243) */
void ndpref_distance(void *hh, int64_t n, const double *q, double *dists)
{
    ref_handle *h = hh;
    ndp_table *t = h->table;
    int na = t->axes->len;
    ndp_query_pts *qp = ndp_query_pts_import((int) n, (double *) q, t->axes);
    for (int64_t i = 0; i < n; i++) {
        int *c = ndp_find_nearest(qp->normed + i * na, qp->indices + i * na, qp->flags + i * na,
                                  t, NDP_METHOD_LINEAR, NDP_SEARCH_KDTREE, &dists[i]);
        free(c);
    }
    ndp_query_pts_free(qp);
}
