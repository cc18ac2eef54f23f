// ndp_fused_kernels.cuh -- the fused ndpolate kernel and its companions.
//
//   k_fused      import + plan + ONE cell gather + reduce for every query, in-grid and
//                extrapolated alike (ndp_query_pts_import + ndp_find_hypercubes + ndpolate,
//                src/ndpolator.c:354-414, :416-508, :510-671; the nearest search of :188-352
//                is a lookup in the nearest-site map, ndp_nsm.cuh)
//   k_slow       the exact generic path, per query, for the few the plan cannot settle
//                (kd distance ties, on-vertex queries beside a void, far-out points, NaN)
//   k_nsm_count / k_nsm_fill   register time: the nearest-site map
//   k_assoc_consistent         register time: is "vertex defined" independent of the associated index?
//
// Scalar tables (vdim 1) with a cell-major copy, 2..6 axes.  Per query the kernel moves
// naxes*8 bytes of query, one contiguous (8 << naxes)-byte cell and 16 bytes of results --
// exactly the algorithmic bytes of SURVEY 8(d) -- and nothing else touches HBM: no off-grid
// records, no second pass, no fetch of cells that turn out to be incomplete.
#pragma once

#include <cuda_runtime.h>
#include "ndp_fused.cuh"
#include "ndp_kernels.cuh"

#ifndef NDP_FUSED_WARPS
#define NDP_FUSED_WARPS 4       // warps per CTA (tuning variants: -DNDP_FUSED_WARPS=3 -DNDP_FUSED_OCC=6 = 18 warps per SM at 112 registers)
#endif
#ifndef NDP_FUSED_TMAQ
#define NDP_FUSED_TMAQ 1      // query tiles by TMA bulk copy + mbarrier (0: per-lane cp.async pieces; same results)
#endif

struct FusedArgs {
    NdpGeom g;
    const double *ticks; int nticks;
    const double *cells; const double *grid;
    const double *q; long long n;
    const int *c_idx; const int *c_flg; const double *c_nrm;   // CACHED kernels: import products (n x naxes) instead of q
    double *interps; double *dists;
    NdpHcSlice H; int hc_in_smem;
    NdpSiteMap M;
    int *slowlist;
    NdpCounters *ctr;
    int gs;                   // k_fused_vec: log2 of the lanes that share one query (each owns components k, k + G, ...)
    NdpBlocks B;              // k_fused: layout of `cells` (block order n = the cell-major copy); g.cstride holds B.stride
};

#if defined(NDP_TU_FUSED)
__device__ __forceinline__ unsigned ndpf_smem_u32(const void *p) { return (unsigned) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void ndpf_cp_async16(unsigned dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// the copy is issued only when `flag` is non-negative (a lane without a cell passes -1)
__device__ __forceinline__ void ndpf_cp_async16_if(unsigned dst, const void *src, int flag)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ge.s32 p, %2, 0;\n\t@p cp.async.cg.shared.global [%0], [%1], 16;\n\t}" ::"r"(dst), "l"(src), "r"(flag) : "memory");
}
__device__ __forceinline__ void ndpf_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void ndpf_cp_async_wait0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// c_ndpolate (:78-99) on a contiguous scalar cell without on-vertex axes: the common case, no selects.
// Same pairing and order of operations as ndp_cell_reduce_halves, hence the same bits.
template <int N>
__device__ __forceinline__ double ndpf_reduce_plain(const double2 *b2, const double *x)
{
    constexpr int HC = 1 << (N - 2);
    double res[2][2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        double fv[2 * HC];
#pragma unroll
        for (int mp = 0; mp < HC; mp++) { const double2 p = b2[2 * mp + h]; fv[2 * mp] = p.x; fv[2 * mp + 1] = p.y; }
#pragma unroll
        for (int a = 0; a < N - 2; a++) {
            const int hb = 1 << (N - 2 - a);
            const double xa = x[a];
#pragma unroll
            for (int j = 0; j < HC; j++)
                if (j < hb) fv[j] = ndp_lerp(fv[j], fv[j + hb], xa);
        }
        res[h][0] = fv[0]; res[h][1] = fv[1];
    }
    const double r0 = ndp_lerp(res[0][0], res[1][0], x[N - 2]);
    const double r1 = ndp_lerp(res[0][1], res[1][1], x[N - 2]);
    return ndp_lerp(r0, r1, x[N - 1]);
}

// src = base + cell * bytes as ONE instruction (IMAD.WIDE.U32 with a 64-bit addend)
__device__ __forceinline__ const unsigned char *ndpf_cell_ptr(const unsigned char *base, unsigned cell, unsigned bytes)
{
    unsigned long long r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(cell), "r"(bytes), "l"((unsigned long long) base));
    return reinterpret_cast<const unsigned char *>(r);
}

// 16-byte copy of which only the first `bytes` (0, 8 or 16) are read from global memory, the rest zero-filled
__device__ __forceinline__ void ndpf_cp_async16_part(unsigned dst, const void *src, unsigned bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ndpf_cp_async8(unsigned dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void ndpf_cp_async_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// TMA bulk copy (cp.async.bulk, SASS UBLKCP) of a contiguous block into shared memory, completion counted in
// bytes on an mbarrier (SYNCS): one elected lane moves a whole tile of queries with one instruction.
__device__ __forceinline__ void ndpf_mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void ndpf_mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void ndpf_mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ndpf_bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// A bulk copy (async proxy) that refills a shared-memory buffer must not be issued while a lane's reads of that buffer
// are still in flight.  Program order issues the reads first, but only a register dependency makes the warp WAIT for
// their data (a refill placed right after the bare loads was measured to overtake them under load: whole tiles of
// wrong rows, profiles/r02w).  Passing a value computed from everything that was read through this empty asm is
// that dependency; every lane passes it before the (converged) warp reaches the copy.
__device__ __forceinline__ void ndpf_settle(int v) { asm volatile("" ::"r"(v) : "memory"); }
__device__ __forceinline__ void ndpf_mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tNDPF_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra NDPF_DONE;\n\t"
                 "bra NDPF_WAIT;\n\tNDPF_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

// shared memory per warp: 32 copy slots; PIPE: + 2 query tiles + their 2 mbarriers; PIPE 1: + 2 tiles of saved plans
__host__ __device__ constexpr unsigned ndpf_warp_bytes(int n, int pipe)
{
    return 32u * ((8u << n) + 16u) + (pipe ? (pipe == 3 ? 1u : 2u) * 32u * 8u * (unsigned) n + 16u : 0u) +
           (pipe == 1 ? 2u * 32u * (8u * ((unsigned) n + 1u) + 16u) : 0u);
}

// What a lane keeps of its plan between issuing a tile's copies and reducing it.
template <int N>
struct NdpFusedItem {
    double x[N];
    double dist;
    unsigned pin, sel;
    int kind, searched;
};

// PIPE 0: per warp and tile  import -> plan -> copy -> wait -> reduce  (latency hidden by the other warps only).
// PIPE 1: software pipeline -- the queries of tile t+2 are prefetched into shared memory, and the import + plan
//         of tile t+1 run while the cell copies of tile t are in flight; the plan of the tile in flight waits in
//         shared memory, so the long plan section has the whole register file.
// PIPE 2: the same pipeline with the waiting plan kept in registers (less shared memory, one more CTA per SM).
// PIPE 3: PIPE 2 with ONE query buffer per warp (refilled as soon as the import has read it): 1.3 KB less shared
//         memory per warp, which is what a fifth CTA per SM needs for 5 axes.
// VS: see ndp_nsm_search.
// CACHED: the import products of the queries (ndpb_find: indices, flags, normed) are read instead of the query
//         points -- the hypercube-caching path (ndpb_ndpolate_cached); runs the PIPE 0 loop.
#ifdef NDP_FUSED_MAXNREG       // tuning variants: an explicit register cap instead of the one launch bounds imply
#define NDP_FUSED_BOUNDS(occ) __maxnreg__(NDP_FUSED_MAXNREG)
#else
#define NDP_FUSED_BOUNDS(occ) __launch_bounds__(NDP_FUSED_WARPS * 32, occ)
#endif
template <int N, int METHOD, int MODE, int OCC, int VS, int PIPE_, int CACHED = 0>
__global__ void NDP_FUSED_BOUNDS(OCC) k_fused(const __grid_constant__ FusedArgs A)
{
    constexpr int PIPE = CACHED ? 0 : PIPE_;
    constexpr unsigned CELL_BYTES = 8u << N;
    constexpr unsigned SLOT = CELL_BYTES + 16u;           // +16: the 128-bit shared loads of a quarter warp hit distinct bank groups
    constexpr int CHUNKS = (int) (CELL_BYTES / 16u);
    constexpr int LPB = CHUNKS < 32 ? CHUNKS : 32;        // lanes per cell
    constexpr unsigned QROW = 8u * N;                     // bytes of one query row
    constexpr unsigned QTILE = 32u * QROW;                // one tile of queries, laid out as in global memory
    constexpr unsigned QCHUNKS = QTILE / 16u;
    constexpr unsigned WARP_BYTES = ndpf_warp_bytes(N, PIPE);
    extern __shared__ __align__(128) unsigned char ndp_smem[];
    const NdpGeom &g = A.g;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // shared layout: ticks | hypercube-mask slice (when small) | per warp: copy slots [| 2 query tiles | 2 plan tiles]
    double *s_ticks = reinterpret_cast<double *>(ndp_smem);
    const unsigned ticks_bytes = ((unsigned) A.nticks * 8u + 127u) & ~127u;
    uint32_t *s_hc = reinterpret_cast<uint32_t *>(ndp_smem + ticks_bytes);
    const unsigned hc_bytes = A.hc_in_smem ? (((unsigned) A.H.nwords * 4u + 127u) & ~127u) : 0u;
    unsigned char *s_warp = ndp_smem + ticks_bytes + hc_bytes + (size_t) warp * WARP_BYTES;
    for (int i = threadIdx.x; i < A.nticks; i += blockDim.x) s_ticks[i] = A.ticks[i];
    if (A.hc_in_smem)
        for (int i = threadIdx.x; i < A.H.nwords; i += blockDim.x) s_hc[i] = A.H.bits[i];
    __syncthreads();
    const uint32_t *hcbits = A.hc_in_smem ? s_hc : A.H.bits;

    unsigned char *s_slots = s_warp;
    unsigned char *s_q = s_warp + 32u * SLOT;                       // PIPE: [2][32][N] doubles
    const unsigned s_bar = ndpf_smem_u32(s_q + 2u * QTILE);         // PIPE: one mbarrier per query buffer
    const unsigned s_bar_1 = ndpf_smem_u32(s_q + QTILE);            // PIPE 3: the mbarrier of its one query buffer
    unsigned char *s_item = s_q + 2u * QTILE + 16u;                 // PIPE 1: [2][N + 1][32] doubles + [2][32] uint4
    const unsigned char *my_slot = s_slots + (size_t) lane * SLOT;
    // cooperative copy: a cell is fetched by LPB consecutive lanes, 16 bytes each, i.e. as ONE contiguous (8 << N)-
    // byte request (measured on B200, tools/probe_gather.cu: 5.5x the rate of per-thread walks).  In sweep k the
    // LPB lanes of group `grp` fetch the cell of lane grp * LPB + k: a width-LPB shuffle with a constant lane.
    const unsigned sub = lane % LPB, grp = lane / LPB;
    const unsigned dst0 = ndpf_smem_u32(s_slots + (size_t) (grp * LPB) * SLOT) + sub * 16u;
    // blocked layout (NdpBlocks): piece `sub` of a cell sits at a fixed offset from the cell's first block
    const unsigned BLOCK_BYTES = 8u << A.B.order;
    const unsigned char *src0 = reinterpret_cast<const unsigned char *>(A.cells) + ndp_block_piece_offset(N, A.B.order, A.B.stride, (int) sub);
    const long long src_step = CHUNKS > LPB ? ndp_block_piece_offset(N, A.B.order, A.B.stride, LPB) : 0;   // 512-byte cells: second sweep

    const long long ntiles = (A.n + 31) / 32;
    const long long tile_stride = (long long) gridDim.x * NDP_FUSED_WARPS;
    const long long tile0 = (long long) blockIdx.x * NDP_FUSED_WARPS + warp;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    const bool q_aligned = (reinterpret_cast<unsigned long long>(A.q) & 15ull) == 0;
    int n_searched = 0;

    // import + plan of this lane's query of `tile`; qrow: its N coordinates (registers or shared memory)
    auto make_plan = [&](long long tile, const double *qrow, NdpPlan<N> &P) {
        P.kind = NDP_PLAN_IDLE; P.cell = 0; P.vflat = 0; P.pin = 0; P.sel = 0; P.dist = 0.0; P.searched = 0;
        const long long w = tile * 32 + lane;
        if (w < A.n) {
            int idx[N], flg[N];
            double nrm[N];
            if (CACHED) {
#pragma unroll
                for (int a = 0; a < N; a++) { idx[a] = A.c_idx[w * N + a]; flg[a] = A.c_flg[w * N + a]; nrm[a] = A.c_nrm[w * N + a]; }
            } else ndp_import_query<N>(g, s_ticks, qrow, idx, flg, nrm);
            ndp_fused_plan<N, METHOD, MODE, VS>(g, A.H, hcbits, A.M, idx, flg, nrm, P);
        }
    };
    // starts the copies of a tile's cells (no commit); a NEAREST-extrapolated lane fetches its vertex value instead
    auto issue_cells = [&](const NdpPlan<N> &P) {
        if (METHOD == NDP_M_NEAREST && P.kind == NDP_PLAN_VALUE) ndpf_cp_async8(ndpf_smem_u32(my_slot), A.grid + P.vflat);
        if (g.small) {
            const int mine = P.kind == NDP_PLAN_GATHER ? (int) P.cell : -1;
#pragma unroll
            for (int k = 0; k < LPB; k++) {
                const int oc = __shfl_sync(0xffffffffu, mine, k, LPB);
                const unsigned char *src = ndpf_cell_ptr(src0, (unsigned) oc, BLOCK_BYTES);
#pragma unroll
                for (int c = 0; c < CHUNKS / LPB; c++)
                    ndpf_cp_async16_if(dst0 + (unsigned) k * SLOT + (unsigned) (c * LPB) * 16u, src + c * src_step, oc);
            }
        } else {
            const long long mine = P.kind == NDP_PLAN_GATHER ? P.cell : -1;
#pragma unroll
            for (int k = 0; k < LPB; k++) {
                const long long oc = __shfl_sync(0xffffffffu, mine, k, LPB);
                const unsigned char *src = src0 + (unsigned long long) oc * BLOCK_BYTES;
#pragma unroll
                for (int c = 0; c < CHUNKS / LPB; c++)
                    ndpf_cp_async16_if(dst0 + (unsigned) k * SLOT + (unsigned) (c * LPB) * 16u, src + c * src_step, (int) (oc >> 32));
            }
        }
    };
    // reduces / finishes this lane's query of `tile` once its cell has landed, writes the results
    auto finish = [&](long long tile, const NdpFusedItem<N> &it) {
        const long long w = tile * 32 + lane;
        double val = qnan;
        if (METHOD == NDP_M_NEAREST && it.kind == NDP_PLAN_VALUE) val = *reinterpret_cast<const double *>(my_slot);
        if (it.kind == NDP_PLAN_GATHER) {
            const double2 *b2 = reinterpret_cast<const double2 *>(my_slot);
            if (it.pin == 0) val = ndpf_reduce_plain<N>(b2, it.x);
            else { bool unused; val = ndp_cell_reduce_halves<N>(b2, it.x, it.pin, it.sel, unused); }
        }
        if (it.kind != NDP_PLAN_IDLE && it.kind != NDP_PLAN_SLOW) { A.interps[w] = val; A.dists[w] = it.dist; }
        if (METHOD != NDP_M_NONE) n_searched += it.searched;
        const unsigned slow = __ballot_sync(0xffffffffu, it.kind == NDP_PLAN_SLOW);
        if (slow) {                            // rare: warp-aggregated append to the slow list
            int base = 0;
            const int leader = __ffs(slow) - 1;
            if ((int) lane == leader) base = atomicAdd(&A.ctr->n_slow, __popc(slow));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (it.kind == NDP_PLAN_SLOW) A.slowlist[base + __popc(slow & ((1u << lane) - 1u))] = (int) w;
        }
    };
    auto item_of = [&](const NdpPlan<N> &P, NdpFusedItem<N> &it) {
#pragma unroll
        for (int a = 0; a < N; a++) it.x[a] = P.x[a];
        it.dist = P.dist; it.pin = P.pin; it.sel = P.sel; it.kind = P.kind; it.searched = P.searched;
    };

    if (PIPE == 3) {
        // one query buffer: {q(t)} lands -> plan(t) -> the buffer is refilled with q(t + 1) at once.  cp.async groups
        // alternate {q(next + 1)} {cells(next)}: "all but the newest" (wait_group 1) is the cells before a finish and
        // the queries before a plan; with a TMA tile the query group is empty and its mbarrier is waited for instead.
        unsigned by_tma = 0, phase = 0;
        if (lane == 0) { ndpf_mbar_init(s_bar_1, 1); ndpf_mbar_fence_init(); }
        __syncwarp();
        auto prefetch_q1 = [&](long long tile) {
            by_tma = 0;
            if (tile >= ntiles) return;
            const unsigned dst = ndpf_smem_u32(s_q);
            const unsigned char *src = reinterpret_cast<const unsigned char *>(A.q) + tile * (long long) QTILE;
            const long long left = (A.n - tile * 32) * (long long) QROW;
            if (NDP_FUSED_TMAQ && q_aligned && left >= (long long) QTILE) {
                by_tma = 1;
                if (lane == 0) { ndpf_mbar_expect_tx(s_bar_1, QTILE); ndpf_bulk_g2s(dst, src, QTILE, s_bar_1); }
            } else if (q_aligned) {
#pragma unroll
                for (unsigned c = 0; c < (QCHUNKS + 31u) / 32u; c++) {
                    const unsigned piece = c * 32u + lane;
                    const long long off = (long long) piece * 16;
                    if (piece < QCHUNKS && off < left) ndpf_cp_async16_part(dst + piece * 16u, src + off, left - off >= 16 ? 16u : 8u);
                }
            } else {
#pragma unroll
                for (unsigned c = 0; c < N; c++) {
                    const long long off = (long long) (c * 32u + lane) * 8;
                    if (off < left) ndpf_cp_async8(dst + (c * 32u + lane) * 8u, src + off);
                }
            }
        };
        auto await_q1 = [&]() {
            if (by_tma) { ndpf_mbar_wait(s_bar_1, phase); phase ^= 1u; }
        };
        if (tile0 < ntiles) {
            prefetch_q1(tile0);
            ndpf_cp_async_commit();
            ndpf_cp_async_wait0();
            await_q1();
            __syncwarp();
            NdpPlan<N> P;
            make_plan(tile0, reinterpret_cast<const double *>(s_q) + lane * N, P);
            ndpf_settle(P.kind);
            __syncwarp();                          // every lane has read its row of the query buffer
            prefetch_q1(tile0 + tile_stride);
            ndpf_cp_async_commit();                // {q(t1)}
            issue_cells(P);
            ndpf_cp_async_commit();                // {cells(t0)}
            NdpFusedItem<N> it_cur;
            item_of(P, it_cur);
            for (long long tile = tile0; tile < ntiles; tile += tile_stride) {
                const long long next = tile + tile_stride;
                NdpPlan<N> Pn;
                Pn.kind = NDP_PLAN_IDLE; Pn.cell = 0; Pn.vflat = 0;
                NdpFusedItem<N> itn;
                itn.kind = NDP_PLAN_IDLE; itn.pin = 0; itn.sel = 0; itn.searched = 0; itn.dist = 0.0;
                if (next < ntiles) {
                    ndpf_cp_async_wait1();         // q(next) is in the group before cells(tile)
                    await_q1();
                    __syncwarp();
                    make_plan(next, reinterpret_cast<const double *>(s_q) + lane * N, Pn);
                    item_of(Pn, itn);
                    ndpf_settle(Pn.kind);
                    __syncwarp();                  // the query buffer has been read by every lane
                    prefetch_q1(next + tile_stride);
                    ndpf_cp_async_commit();        // {q(next + 1)}: now the newest group, cells(tile) the one before
                    ndpf_cp_async_wait1();
                } else ndpf_cp_async_wait0();
                __syncwarp();                      // the cells of `tile` have landed, for every lane
                finish(tile, it_cur);
                it_cur = itn;
                __syncwarp();                      // the slots are free again
                if (next < ntiles) {
                    issue_cells(Pn);
                    ndpf_cp_async_commit();        // {cells(next)}
                }
            }
        }
    } else if (!PIPE) {
        for (long long tile = tile0; tile < ntiles; tile += tile_stride) {
            const long long w = tile * 32 + lane;
            double qv[N];
            if (!CACHED && w < A.n) {
#pragma unroll
                for (int a = 0; a < N; a++) qv[a] = A.q[w * N + a];
            }
            NdpPlan<N> P;
            make_plan(tile, qv, P);
            issue_cells(P);
            ndpf_cp_async_commit();
            NdpFusedItem<N> it;
            item_of(P, it);
            ndpf_cp_async_wait0();
            __syncwarp();                          // the other lanes' copies into my slot are visible
            finish(tile, it);
            __syncwarp();                          // the slots may be overwritten by the next tile's copies
        }
    } else {
        // query prefetch: the tile's 32 rows are contiguous in global memory.  A whole, 16-byte-aligned tile is ONE
        // TMA bulk copy issued by lane 0, completing on the buffer's mbarrier; a partial last tile or an unaligned
        // query array goes through per-lane cp.async pieces in the current cp.async group instead.
        unsigned by_tma = 0, phase = 0;                                            // per buffer: bit `buf`
        if (lane == 0) { ndpf_mbar_init(s_bar, 1); ndpf_mbar_init(s_bar + 8u, 1); ndpf_mbar_fence_init(); }
        __syncwarp();
        auto prefetch_q = [&](long long tile, int buf) {
            if (tile >= ntiles) return;
            const unsigned dst = ndpf_smem_u32(s_q + (unsigned) buf * QTILE);
            const unsigned char *src = reinterpret_cast<const unsigned char *>(A.q) + tile * (long long) QTILE;
            const long long left = (A.n - tile * 32) * (long long) QROW;            // bytes of this tile that exist
            by_tma &= ~(1u << buf);
            if (NDP_FUSED_TMAQ && q_aligned && left >= (long long) QTILE) {
                by_tma |= 1u << buf;
                if (lane == 0) {
                    ndpf_mbar_expect_tx(s_bar + 8u * (unsigned) buf, QTILE);
                    ndpf_bulk_g2s(dst, src, QTILE, s_bar + 8u * (unsigned) buf);
                }
            } else if (q_aligned) {
#pragma unroll
                for (unsigned c = 0; c < (QCHUNKS + 31u) / 32u; c++) {
                    const unsigned piece = c * 32u + lane;
                    const long long off = (long long) piece * 16;
                    if (piece < QCHUNKS && off < left)
                        ndpf_cp_async16_part(dst + piece * 16u, src + off, left - off >= 16 ? 16u : 8u);
                }
            } else {
#pragma unroll
                for (unsigned c = 0; c < N; c++) {
                    const long long off = (long long) (c * 32u + lane) * 8;
                    if (off < left) ndpf_cp_async8(dst + (c * 32u + lane) * 8u, src + off);
                }
            }
        };
        // the queries of buffer `buf` have landed (its cp.async group, if any, was waited for by the caller)
        auto await_q = [&](int buf) {
            if (by_tma & (1u << buf)) {
                ndpf_mbar_wait(s_bar + 8u * (unsigned) buf, (phase >> buf) & 1u);
                phase ^= 1u << buf;
            }
        };
        double *itemd = reinterpret_cast<double *>(s_item);                         // [2][N + 1][32]
        uint4 *itemu = reinterpret_cast<uint4 *>(s_item + 2u * (N + 1) * 32u * 8u);  // [2][32]
        auto save_item = [&](int buf, const NdpFusedItem<N> &it) {
            double *d = itemd + (size_t) buf * (N + 1) * 32 + lane;
#pragma unroll
            for (int a = 0; a < N; a++) d[a * 32] = it.x[a];
            d[N * 32] = it.dist;
            itemu[buf * 32 + lane] = make_uint4(it.pin, it.sel, (unsigned) it.kind, (unsigned) it.searched);
        };
        auto load_item = [&](int buf, NdpFusedItem<N> &it) {
            const double *d = itemd + (size_t) buf * (N + 1) * 32 + lane;
#pragma unroll
            for (int a = 0; a < N; a++) it.x[a] = d[a * 32];
            it.dist = d[N * 32];
            const uint4 u = itemu[buf * 32 + lane];
            it.pin = u.x; it.sel = u.y; it.kind = (int) u.z; it.searched = (int) u.w;
        };
        if (tile0 < ntiles) {
            // groups: {q(t0)} {q(t1)} {cells(t0), q(t2)} | per iteration {cells(t+1), q(t+3)}
            prefetch_q(tile0, 0);
            ndpf_cp_async_commit();
            prefetch_q(tile0 + tile_stride, 1);
            ndpf_cp_async_commit();
            ndpf_cp_async_wait1();
            await_q(0);
            __syncwarp();
            NdpPlan<N> P;
            make_plan(tile0, reinterpret_cast<const double *>(s_q) + lane * N, P);
            ndpf_settle(P.kind);
            __syncwarp();                          // every lane has read its row of query buffer 0
            issue_cells(P);
            prefetch_q(tile0 + 2 * tile_stride, 0);
            ndpf_cp_async_commit();
            NdpFusedItem<N> it_cur;
            item_of(P, it_cur);
            if (PIPE == 1) save_item(0, it_cur);
            int cur = 0;
            for (long long tile = tile0; tile < ntiles; tile += tile_stride) {
                const long long next = tile + tile_stride;
                NdpPlan<N> Pn;
                Pn.kind = NDP_PLAN_IDLE; Pn.cell = 0; Pn.vflat = 0;
                NdpFusedItem<N> itn;
                itn.kind = NDP_PLAN_IDLE; itn.pin = 0; itn.sel = 0; itn.searched = 0; itn.dist = 0.0;
                if (next < ntiles) {
                    // q(next) sits in the group before the newest one; the cells of `tile` stay in flight meanwhile
                    ndpf_cp_async_wait1();
                    await_q(cur ^ 1);
                    __syncwarp();
                    make_plan(next, reinterpret_cast<const double *>(s_q + (unsigned) (cur ^ 1) * QTILE) + lane * N, Pn);
                    item_of(Pn, itn);
                    ndpf_settle(Pn.kind);
                    if (PIPE == 1) save_item(cur ^ 1, itn);
                }
                ndpf_cp_async_wait0();
                __syncwarp();                      // the cells of `tile` have landed, for every lane
                if (PIPE == 1) load_item(cur, it_cur);
                finish(tile, it_cur);
                if (PIPE == 2) it_cur = itn;
                __syncwarp();                      // slots and query buffer cur^1 are free again
                if (next < ntiles) {
                    issue_cells(Pn);
                    prefetch_q(next + 2 * tile_stride, cur ^ 1);
                    ndpf_cp_async_commit();
                }
                cur ^= 1;
            }
        }
    }
    if (METHOD != NDP_M_NONE) {
        for (int off = 16; off; off >>= 1) n_searched += __shfl_xor_sync(0xffffffffu, n_searched, off);
        if (lane == 0 && n_searched) atomicAdd(&A.ctr->n_off, n_searched);
    }
}

// ---------------------------------------------------------------------------
// k_fused_vec: the same plan for vector-valued tables (vdim > 1).  A corner vector is contiguous in the
// ORIGINAL grid (C3: 16 passbands = one 128-byte line), so no cell-major copy is involved: the plan of a
// query (one lane each) is broadcast to the G = 2^gs lanes that share it, lane k of the group gathers
// component k, k + G, ... of the 2^N corners with coalesced G*8-byte loads and reduces them in registers.
// In A.g the cell strides hold the VERTEX strides (the plan's "cell" is the vertex below the hypercube).
// ---------------------------------------------------------------------------
template <int N, int METHOD, int MODE, int VS>
__global__ void __launch_bounds__(NDP_FUSED_WARPS * 32, 4) k_fused_vec(const __grid_constant__ FusedArgs A)
{
    extern __shared__ __align__(128) unsigned char ndp_smem[];
    const NdpGeom &g = A.g;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *s_ticks = reinterpret_cast<double *>(ndp_smem);
    const unsigned ticks_bytes = ((unsigned) A.nticks * 8u + 127u) & ~127u;
    uint32_t *s_hc = reinterpret_cast<uint32_t *>(ndp_smem + ticks_bytes);
    for (int i = threadIdx.x; i < A.nticks; i += blockDim.x) s_ticks[i] = A.ticks[i];
    if (A.hc_in_smem)
        for (int i = threadIdx.x; i < A.H.nwords; i += blockDim.x) s_hc[i] = A.H.bits[i];
    __syncthreads();
    const uint32_t *hcbits = A.hc_in_smem ? s_hc : A.H.bits;

    const int V = g.vdim, gs = A.gs, G = 1 << gs;
    const int qi = (int) lane >> gs, kl = (int) lane & (G - 1);
    const int per_sweep = 32 >> gs;
    const long long ntiles = (A.n + 31) / 32;
    const long long tile_stride = (long long) gridDim.x * NDP_FUSED_WARPS;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    int n_searched = 0;

    for (long long tile = (long long) blockIdx.x * NDP_FUSED_WARPS + warp; tile < ntiles; tile += tile_stride) {
        const long long w = tile * 32 + lane;
        NdpPlan<N> P;
        P.kind = NDP_PLAN_IDLE; P.cell = 0; P.vflat = 0; P.pin = 0; P.sel = 0; P.dist = 0.0; P.searched = 0;
#pragma unroll
        for (int a = 0; a < N; a++) P.x[a] = 0.0;
        if (w < A.n) {
            int idx[N], flg[N];
            double nrm[N];
            if (A.c_idx) {            // hypercube-caching path: cached import products instead of query points
#pragma unroll
                for (int a = 0; a < N; a++) { idx[a] = A.c_idx[w * N + a]; flg[a] = A.c_flg[w * N + a]; nrm[a] = A.c_nrm[w * N + a]; }
            } else {
                double qv[N];
#pragma unroll
                for (int a = 0; a < N; a++) qv[a] = A.q[w * N + a];
                ndp_import_query<N>(g, s_ticks, qv, idx, flg, nrm);
            }
            ndp_fused_plan<N, METHOD, MODE, VS>(g, A.H, hcbits, A.M, idx, flg, nrm, P);
        }
        const long long mybase = P.kind == NDP_PLAN_VALUE ? P.vflat : P.cell;
        const int meta = (P.kind << 16) | ((int) P.pin << 8) | (int) P.sel;
        for (int s = 0; s < G; s++) {
            const int src = s * per_sweep + qi;
            const int m = __shfl_sync(0xffffffffu, meta, src);
            const long long base = __shfl_sync(0xffffffffu, mybase, src);
            double x[N];
#pragma unroll
            for (int a = 0; a < N; a++) x[a] = __shfl_sync(0xffffffffu, P.x[a], src);
            const int kind = m >> 16;
            const unsigned pin = (unsigned) (m >> 8) & 0xffu, sel = (unsigned) m & 0xffu;
            if (kind == NDP_PLAN_IDLE || kind == NDP_PLAN_SLOW) continue;
            double *out = A.interps + (tile * 32 + src) * (long long) V;
            for (int k = kl; k < V; k += G) {
                double val = qnan;
                if (kind == NDP_PLAN_GATHER) {
                    const double *b = A.grid + base * V + k;
                    bool unused;
                    val = ndp_cell_reduce<N>([&](double *fv) {
#pragma unroll
                        for (int j = 0; j < (1 << N); j++) {
                            long long off = 0;
#pragma unroll
                            for (int a = 0; a < N; a++) if ((j >> (N - 1 - a)) & 1) off += g.cstride[a];
                            fv[j] = b[off * V];
                        }
                    }, x, pin, sel, unused);
                } else if (METHOD == NDP_M_NEAREST && kind == NDP_PLAN_VALUE) val = A.grid[base * V + k];
                out[k] = val;
            }
        }
        if (P.kind != NDP_PLAN_IDLE && P.kind != NDP_PLAN_SLOW) A.dists[w] = P.dist;
        if (METHOD != NDP_M_NONE) n_searched += P.searched;
        const unsigned slow = __ballot_sync(0xffffffffu, P.kind == NDP_PLAN_SLOW);
        if (slow) {
            int base = 0;
            const int leader = __ffs(slow) - 1;
            if ((int) lane == leader) base = atomicAdd(&A.ctr->n_slow, __popc(slow));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (P.kind == NDP_PLAN_SLOW) A.slowlist[base + __popc(slow & ((1u << lane) - 1u))] = (int) w;
        }
    }
    if (METHOD != NDP_M_NONE) {
        for (int off = 16; off; off >>= 1) n_searched += __shfl_xor_sync(0xffffffffu, n_searched, off);
        if (lane == 0 && n_searched) atomicAdd(&A.ctr->n_off, n_searched);
    }
}
#endif  // NDP_TU_FUSED

// ---------------------------------------------------------------------------
// k_slow: one thread per listed query, the whole reference pipeline with the exact
// lattice search (import, hypercube with NaN test, search, publish, finish).
// ---------------------------------------------------------------------------
struct SlowArgs {
    NdpGeom g;
    const NdpPyramid *F; const NdpPyramid *R;
    NdpTopo topo; int have_topo;
    const double *ticks; int nticks; int ticks_in_smem;
    const double *grid;
    const double *q;
    int method, search;
    const int *list; const int *count;     // query ids to do, device count; list == nullptr: queries 0 .. count_host
    long long count_host;
    const int *c_idx; const int *c_flg; const double *c_nrm;   // cached import products (n x naxes) instead of q, or nullptr
    double *interps; double *dists;
    int *tielist;                          // out: query ids whose kd tie needs the topology (have_topo == 0)
    int count_searched;                    // 1: add the searched queries to ctr->n_off (slow pass), 0: tie pass
    NdpCounters *ctr;
};

#if defined(NDP_TU_SEARCH)
template <int MODE>
__global__ void __launch_bounds__(NDP_BLOCK, 2) k_slow(const __grid_constant__ SlowArgs A)
{
    extern __shared__ double s_ticks[];
    const double *tk = ndp_stage_ticks(A.ticks, A.nticks, A.ticks_in_smem, s_ticks);
    const NdpGeom &g = A.g;
    const int n = g.naxes, V = g.vdim;
    const int mode = ndp_mode_of(A.method, A.search);
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    const long long total = A.list ? (long long) *A.count : A.count_host;
    for (long long s = (long long) blockIdx.x * blockDim.x + threadIdx.x; s < total; s += (long long) gridDim.x * blockDim.x) {
        const long long i = A.list ? (long long) A.list[s] : s;
        int idx[NDP_MAX_AXES], flg[NDP_MAX_AXES];
        double nrm[NDP_MAX_AXES];
        if (A.c_idx) {
            for (int a = 0; a < n; a++) { idx[a] = A.c_idx[i * n + a]; flg[a] = A.c_flg[i * n + a]; nrm[a] = A.c_nrm[i * n + a]; }
        } else ndp_import_query_seq<0>(g, tk, A.q + i * n, idx, flg, nrm);
        unsigned pin, sel;
        const bool oob = ndp_classify<0>(g, flg, nrm, pin, sel);
        bool fdhc = !oob;
        if (!oob) {
            for (int k = 0; k < V; k++) {
                bool bad = false;
                const double val = ndp_eval_cell<0, false, false>(g, A.grid, nullptr, idx, nrm, pin, sel, k, bad);
                if (k == 0 && bad) { fdhc = false; break; }
                A.interps[i * V + k] = val;
            }
        }
        if (fdhc) { A.dists[i] = 0.0; continue; }
        if (A.method == NDP_M_NONE) {
            for (int k = 0; k < V; k++) A.interps[i * V + k] = qnan;
            A.dists[i] = 0.0;
            continue;
        }
        if (A.count_searched) atomicAdd(&A.ctr->n_hard, 1);
        NdpQuery Q;
        ndp_make_query(g, idx, nrm, Q);
        NdpHit hit;
        hit.has_coords = 0;
        if (!ndp_search_fast<0, MODE>(g, *A.F, mode, Q, hit))
            ndp_search_descend<MODE>(g, *A.F, *A.R, mode, Q, A.have_topo ? &A.topo : nullptr, hit);
        if (hit.status == NDP_SEARCH_TIE) {
            A.tielist[atomicAdd(&A.ctr->n_tie, 1)] = (int) i;
            continue;                                       // redone once the topology exists
        }
        if (A.have_topo && !A.count_searched) atomicAdd(&A.ctr->n_tie_resolved, 1);
        ndp_scan_sentinel(*A.F, mode, hit);
        int coords[NDP_MAX_AXES];
        double dist;
        if (hit.status == NDP_SEARCH_EMPTY) {               // empty kd-tree: see ndp_publish
            hit.vertex = 0;
            ndp_coords_of(g, Q, 0, coords);
            dist = 1e10;
        } else {
            ndp_coords_of(g, Q, hit.vertex, coords);
            dist = (mode >= NDP_MODE_LS_NEAREST) ? hit.dsel : ndp_reported_dist(g, A.method, Q, coords);
        }
        A.dists[i] = dist;
        for (int k = 0; k < V; k++)
            A.interps[i * V + k] = ndp_finish_component<0, false, false>(g, A.grid, nullptr, A.method, idx, nrm, coords, k);
    }
}
#endif  // NDP_TU_SEARCH

// ---------------------------------------------------------------------------
// register time
// ---------------------------------------------------------------------------
#if defined(NDP_TU_TAPS)
// pass 1: candidates per region (counts[region]; overflowing regions count 0 and keep the lattice search)
__global__ void __launch_bounds__(128) k_nsm_count(const __grid_constant__ NdpSiteMap M, int *counts)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= M.nregions) return;
    unsigned long long cand[NSM_KMAX];
    const int n = nsm_region_candidates(M, r, cand);
    counts[r] = n < 0 ? 0 : n;
}

// pass 2: M.offs holds the exclusive scan of the counts; the candidates are recomputed and written
__global__ void __launch_bounds__(128) k_nsm_fill(const __grid_constant__ NdpSiteMap M, uint32_t *ents, uint32_t *head)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= M.nregions) return;
    const int beg = M.offs[r], n = M.offs[r + 1] - beg;
    if (n == 0) return;                                  // head[] was zeroed: no candidate, the lattice search decides
    unsigned long long cand[NSM_KMAX];
    const int got = nsm_region_candidates(M, r, cand);
    if (got == n) {
        nsm_write_entries(M, cand, n, ents + (long long) beg * M.ew);
        nsm_write_head(M, cand, n, beg, head + (long long) r * nsm_head_words(M.ew));
    }
}

// *flag is cleared when component 0 of some basic vertex is NaN at one associated index but not at another
// (then the hypercube-mask bit does not decide "fully defined" and the fused path is not used)
__global__ void k_assoc_consistent(const double *__restrict__ grid, int nverts, long long per_vertex, int vdim, int *flag)
{
    const long long total = (long long) nverts * per_vertex;
    for (long long e = (long long) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long) gridDim.x * blockDim.x) {
        const long long v = e / per_vertex;
        const double first = grid[v * per_vertex * vdim], mine = grid[e * vdim];
        if ((first != first) != (mine != mine)) *flag = 0;
    }
}
#endif  // NDP_TU_TAPS
