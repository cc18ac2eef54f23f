// ndp_table.cuh -- register-time kernels: vertex / hypercube masks
// (reference src/ndp_types.c:174-258), the occupancy pyramid the nearest search
// descends, the cell-major copy of the grid, and the level-synchronous rebuild of
// the reference's insertion-ordered kd-tree shape (src/ndpolator.c:114-136,
// external/kdtree/kdtree.c:167-209) used to settle exact distance ties.
#pragma once

#include <cuda_runtime.h>
#include "ndp_core.cuh"
#include "ndp_fused.cuh"

// vmask[v] = component 0 of the basic vertex (associated indices 0) is not NaN
// (ndp_types.c:179-184).  One warp packs 32 vertices into one word.
__global__ void k_vmask_bits(const double *__restrict__ grid, int nverts, long long doubles_per_vertex, uint32_t *bits)
{
    const long long v = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    bool ok = false;
    if (v < nverts) { double x = grid[v * doubles_per_vertex]; ok = (x == x); }
    const unsigned word = __ballot_sync(0xffffffffu, ok);
    if ((threadIdx.x & 31) == 0 && v < nverts) bits[v >> 5] = word;
}

// hcmask[v] = v is the superior corner of a cell whose 2^nbasic corners are all
// defined (ndp_types.c:201-258).
__global__ void k_hcmask_bits(const __grid_constant__ NdpGeom g, const uint32_t *__restrict__ vbits, uint32_t *bits)
{
    const long long v = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    bool ok = false;
    if (v < g.nverts && ndp_bit(vbits, (int) v)) {
        ok = true;
        for (int a = 0; a < g.nbasic && ok; a++)
            if (((int) v / g.bstride[a]) % g.len[a] == 0) ok = false;
        for (int corner = 0; corner < (1 << g.nbasic) && ok; corner++) {
            int w = (int) v;
            for (int a = 0; a < g.nbasic; a++)
                if (!((corner >> a) & 1)) w -= g.bstride[a];
            if (!ndp_bit(vbits, w)) ok = false;
        }
    }
    const unsigned word = __ballot_sync(0xffffffffu, ok);
    if ((threadIdx.x & 31) == 0 && v < g.nverts) bits[v >> 5] = word;
}

struct PyrLevelArgs {
    int nb;
    int cdim[NDP_MAX_AXES], cstride[NDP_MAX_AXES];   // child level
    int pdim[NDP_MAX_AXES];                           // parent level
    int nparent;
    const uint32_t *cbits;
    uint32_t *pbits;
};

__global__ void k_pyr_level(const __grid_constant__ PyrLevelArgs A)
{
    const long long p = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    bool any = false;
    if (p < A.nparent) {
        int pc[NDP_MAX_AXES];
        int rem = (int) p;
        for (int a = A.nb - 1; a >= 0; a--) { pc[a] = rem % A.pdim[a]; rem /= A.pdim[a]; }
        for (int k = 0; k < (1 << A.nb) && !any; k++) {
            int flat = 0;
            bool inside = true;
            for (int a = 0; a < A.nb; a++) {
                int c = 2 * pc[a] + ((k >> a) & 1);
                if (c >= A.cdim[a]) { inside = false; break; }
                flat += c * A.cstride[a];
            }
            if (inside && ndp_bit(A.cbits, flat)) any = true;
        }
    }
    const unsigned word = __ballot_sync(0xffffffffu, any);
    if ((threadIdx.x & 31) == 0 && p < A.nparent) A.pbits[p >> 5] = word;
}

// Does the mask depend on axis a?  differs[a] is raised when some vertex disagrees with the
// vertex that has the reference coordinate `ref` on axis a (0 for the vertex mask; 1 for the
// hypercube mask, whose coordinate-0 entries are clear by construction and never searched).
__global__ void k_mask_variance(const __grid_constant__ NdpGeom g, const uint32_t *__restrict__ bits, int ref, int *differs)
{
    const long long v = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= g.nverts) return;
    const bool mine = ndp_bit(bits, (int) v);
    for (int a = 0; a < g.nbasic; a++) {
        const int c = ((int) v / g.bstride[a]) % g.len[a];
        if (c < ref || c == ref) continue;
        const int w = (int) v - (c - ref) * g.bstride[a];
        if (ndp_bit(bits, w) != mine) differs[a] = 1;
    }
}

struct SliceArgs {
    NdpGeom g;
    int nvar; int var_axis[NDP_MAX_AXES]; int dim[NDP_MAX_AXES];
    int ref; int nslice;
    const uint32_t *bits; uint32_t *out;
};

// level 0 of the reduced pyramid: the mask restricted to the slice where every invariant axis = ref
__global__ void k_mask_slice(const __grid_constant__ SliceArgs A)
{
    const long long p = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    bool on = false;
    if (p < A.nslice) {
        int rem = (int) p, v = 0;
        bool isvar[NDP_MAX_AXES];
        for (int a = 0; a < A.g.nbasic; a++) isvar[a] = false;
        for (int j = A.nvar - 1; j >= 0; j--) {
            const int c = rem % A.dim[j]; rem /= A.dim[j];
            v += c * A.g.bstride[A.var_axis[j]];
            isvar[A.var_axis[j]] = true;
        }
        for (int a = 0; a < A.g.nbasic; a++) if (!isvar[a]) v += A.ref * A.g.bstride[a];
        on = ndp_bit(A.bits, v);
    }
    const unsigned word = __ballot_sync(0xffffffffu, on);
    if ((threadIdx.x & 31) == 0 && p < A.nslice) A.out[p >> 5] = word;
}

// Row tables of a level-0 mask whose last axis has `len` ticks: per lattice point the nearest set
// coordinate at or below (prevv) and at or above (nextv) along that axis, -1 when there is none.
__global__ void k_row_tables(const uint32_t *__restrict__ bits, long long nrows, int len, short *prevv, short *nextv)
{
    const long long r = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    const long long b0 = r * len;
    short last = -1;
    for (int c = 0; c < len; c++) {
        if (ndp_bit(bits, (int) (b0 + c))) last = (short) c;
        prevv[b0 + c] = last;
    }
    last = -1;
    for (int c = len - 1; c >= 0; c--) {
        if (ndp_bit(bits, (int) (b0 + c))) last = (short) c;
        nextv[b0 + c] = last;
    }
}

// lowest vertex id whose bit is clear (for the 1e10 sentinel of the full scan); *out preset to INT_MAX
__global__ void k_first_clear(const uint32_t *__restrict__ bits, int nverts, int *out)
{
    const long long w = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const int nwords = (nverts + 31) >> 5;
    if (w >= nwords) return;
    uint32_t word = ~bits[w];
    if (w == nwords - 1 && (nverts & 31)) word &= (1u << (nverts & 31)) - 1u;
    if (word) atomicMin(out, (int) (w * 32 + (__ffs(word) - 1)));
}

// number of set bits among the first nverts (pads of the last word are clear by construction); *out preset to 0
__global__ void k_count_bits(const uint32_t *__restrict__ bits, int nverts, int *out)
{
    const long long w = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const int nwords = (nverts + 31) >> 5;
    int c = 0;
    if (w < nwords) {
        uint32_t word = bits[w];
        if (w == nwords - 1 && (nverts & 31)) word &= (1u << (nverts & 31)) - 1u;
        c = __popc(word);
    }
    for (int off = 16; off; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

__global__ void k_unpack_bits(const uint32_t *__restrict__ bits, int n, int32_t *out)
{
    const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = ndp_bit(bits, (int) i) ? 1 : 0;
}

// Cell-major copy: for every cell (superior corner c, all axes) the 2^n corner
// vectors in the reference's corner order (axis 0 = most significant bit),
// contiguous: cells[(cell * 2^n + j) * vdim + k].  One in-grid query then reads
// one contiguous, aligned burst instead of 2^(n-1) scattered sectors.
__global__ void k_build_cells(const __grid_constant__ NdpGeom g, const double *__restrict__ grid, double *__restrict__ cells)
{
    const int n = g.naxes, V = g.vdim;
    const long long total = g.ncells << n;
    for (long long e = (long long) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long) gridDim.x * blockDim.x) {
        const int j = (int) (e & ((1LL << n) - 1));
        long long cell = e >> n;
        long long vert = 0;
        for (int a = n - 1; a >= 0; a--) {
            const int cd = g.len[a] - 1;
            const int c = (int) (cell % cd);          // inferior corner on axis a
            cell /= cd;
            vert += (long long) (c + ((j >> (n - 1 - a)) & 1)) * g.vstride[a];
        }
        for (int k = 0; k < V; k++) cells[e * V + k] = grid[vert * V + k];
    }
}

// Blocked copy (ndp_fused.cuh: NdpBlocks): blocks[(block * 2^K + i)] = corner whose last-K-axes bits are i of the
// cell whose inferior corner is the block's coordinates (vertex coordinates on the outer axes).  K = n gives
// k_build_cells' layout.  Scalar tables only.
struct BlockBuildArgs {
    NdpGeom g;
    int order;
    long long stride[NDP_MAX_AXES];
    long long nblocks;
    const double *grid; double *blocks;
};

__global__ void k_build_blocks(const __grid_constant__ BlockBuildArgs A)
{
    const int n = A.g.naxes, K = A.order;
    const long long total = A.nblocks << K;
    for (long long e = (long long) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long) gridDim.x * blockDim.x) {
        const int i = (int) (e & ((1LL << K) - 1));
        long long blk = e >> K;
        long long vert = 0;
        for (int a = n - 1; a >= 0; a--) {
            const bool inner = a >= n - K;
            const int d = inner ? A.g.len[a] - 1 : A.g.len[a];
            int c = (int) (blk % d);
            blk /= d;
            if (inner) c += (i >> (n - 1 - a)) & 1;
            vert += (long long) c * A.g.vstride[a];
        }
        A.blocks[e] = A.grid[vert];
    }
}

// ---------------------------------------------------------------------------
// kd-tree shape, level-synchronous.
//
// The reference inserts the masked vertices in ascending id (ndpolator.c:119-131)
// with the split axis cycling by depth (kdtree.c:189-193).  Hence every subtree
// holds a set of vertices whose smallest id is the subtree's root, and the rest
// split by (coordinate < root coordinate) on axis depth % nbasic.  Keeping each
// subtree as a contiguous, id-ascending segment, one round per tree level turns
// every segment head into a node and stably partitions the remainder into the
// two child segments; all segments of a level are processed at once.
// ---------------------------------------------------------------------------
struct TopoArgs {
    NdpGeom g;
    int N;               // vertices in the tree
    int axis;            // split axis of this level = depth % nbasic
    int depth;
    const int *ids; const int *seg_start; const int *seg_end;
    int *flag;           // N + 1 entries (last = 0)
    const int *scan;     // exclusive scan of flag, N + 1 entries
    int *ids2; int *seg_start2; int *seg_end2;
    int *parent; int *depthv;
    int *active;         // number of vertices still waiting to become a node
};

__global__ void k_topo_flags(const __grid_constant__ TopoArgs A)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > A.N) return;
    int f = 0;
    if (i < A.N) {
        const int s = A.seg_start[i];
        if (s >= 0 && i != s) {
            const int st = A.g.bstride[A.axis], ln = A.g.len[A.axis];
            f = ((A.ids[i] / st) % ln) < ((A.ids[s] / st) % ln);      // kdtree.c:190
        }
    }
    A.flag[i] = f;
}

__global__ void k_topo_scatter(const __grid_constant__ TopoArgs A)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.N) return;
    const int s = A.seg_start[i];
    if (s < 0 || i == s) {             // already a node, or becomes one now
        A.ids2[i] = A.ids[i]; A.seg_start2[i] = -1; A.seg_end2[i] = -1;
        return;
    }
    const int e = A.seg_end[i];
    const int nl = A.scan[e] - A.scan[s + 1];
    const int lrank = A.scan[i] - A.scan[s + 1];
    int np, ns, ne;
    if (A.flag[i]) { ns = s + 1; ne = s + 1 + nl; np = ns + lrank; }
    else { ns = s + 1 + nl; ne = e; np = ns + (i - (s + 1) - lrank); }
    A.ids2[np] = A.ids[i]; A.seg_start2[np] = ns; A.seg_end2[np] = ne;
    if (np == ns) { A.parent[A.ids[i]] = A.ids[s]; A.depthv[A.ids[i]] = A.depth + 1; }
    else atomicAdd(A.active, 1);       // still inside a segment after this round
}

__global__ void k_fill_int(int *p, long long n, int value)
{
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) p[i] = value;
}

__global__ void k_topo_init(int N, int *seg_start, int *seg_end, const int *ids, int *parent, int *depthv)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    seg_start[i] = 0; seg_end[i] = N;
    if (i == 0) { parent[ids[0]] = -1; depthv[ids[0]] = 0; }
}
