// Shared-memory-tiled exact kNN (k = 8) over the deformation nodes.
//
// Replaces Warpfield::findNeighborsIndex -> nanoflann KD-tree (src/dynfu/warp_field.cpp:111-122,
// include/nanoflann/nanoflann.hpp:1229-1235).  Node positions are staged into shared memory in tiles
// with the TMA 1-D bulk copy (cp.async.bulk ... mbarrier::complete_tx, SASS UBLKCP), double buffered;
// every thread owns one query and keeps its running top-8 in registers.  Shared-memory reads are
// warp-wide broadcasts (all lanes read the same node), so there are no bank conflicts.
#pragma once

#include "dfu_math.cuh"

namespace dfu {

constexpr int KNN_TILE = 1024;  // nodes per shared-memory tile (16 KB); two tiles in flight

DFU_DEV uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

DFU_DEV void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
DFU_DEV void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
DFU_DEV void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
DFU_DEV void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (bytes % 16 == 0, 16 B aligned)
DFU_DEV void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
DFU_DEV void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct KnnSmem {
    float4 tile[2][KNN_TILE];
    uint64_t bar[2];
};

// Block-cooperative scan: every thread of the CTA must call this (it contains __syncthreads).
// nodes: padded float4 array (Npad % 32 == 0, padding at +inf).  On return t holds the thread's
// 8 nearest nodes, ascending by (dist2, index).  `active` = false threads still take part in the
// barriers but skip the arithmetic.
DFU_DEV void knn8_scan_block(KnnSmem& sm, const float4* __restrict__ nodes, int Npad, float qx, float qy, float qz,
                             bool active, Top8& t) {
    const int ntiles = (Npad + KNN_TILE - 1) / KNN_TILE;
    if (threadIdx.x == 0) {
        mbar_init(&sm.bar[0], 1);
        mbar_init(&sm.bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int s = 0; s < 2 && s < ntiles; ++s) {
            const int cnt = min(KNN_TILE, Npad - s * KNN_TILE);
            mbar_expect_tx(&sm.bar[s], (uint32_t) cnt * 16u);
            tma_bulk_g2s(sm.tile[s], nodes + (size_t) s * KNN_TILE, (uint32_t) cnt * 16u, &sm.bar[s]);
        }
    }
    top8_init(t);
    for (int tl = 0; tl < ntiles; ++tl) {
        const int s = tl & 1;
        const int cnt = min(KNN_TILE, Npad - tl * KNN_TILE);
        mbar_wait(&sm.bar[s], (uint32_t) ((tl >> 1) & 1));
        if (active) {
            const float4* __restrict__ tile = sm.tile[s];
            const int base = tl * KNN_TILE;
#pragma unroll 1
            for (int j = 0; j < cnt; j += 4) {  // cnt % 32 == 0
                float dd[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float4 p = tile[j + u];
                    dd[u] = dist2(qx, qy, qz, p.x, p.y, p.z);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (dd[u] < t.d[DFU_KNN - 1]) top8_insert(t, dd[u], base + j + u);
            }
        }
        __syncthreads();  // everyone is done with buffer s
        if (threadIdx.x == 0 && tl + 2 < ntiles) {
            const int c2 = min(KNN_TILE, Npad - (tl + 2) * KNN_TILE);
            fence_proxy_async();
            mbar_expect_tx(&sm.bar[s], (uint32_t) c2 * 16u);
            tma_bulk_g2s(sm.tile[s], nodes + (size_t) (tl + 2) * KNN_TILE, (uint32_t) c2 * 16u, &sm.bar[s]);
        }
    }
}


// ---- G lanes per query ------------------------------------------------------------------------------
// The serial scan above has a critical path of N dependent compare/insert steps per thread.  For point
// queries (tens of thousands, not millions) the node range is split over G consecutive lanes: lane `sub`
// scans nodes sub, sub+G, ... (the G lanes read G consecutive float4 = conflict-free, other groups of the
// warp get the same addresses by broadcast) and the G sorted lists are merged with warp shuffles.
DFU_DEV bool lex_less(float d1, int i1, float d2, int i2) { return d1 < d2 || (d1 == d2 && i1 < i2); }

DFU_DEV void lex_cswap(float& d1, int& i1, float& d2, int& i2) {
    if (lex_less(d2, i2, d1, i1)) {
        const float td = d1; d1 = d2; d2 = td;
        const int ti = i1; i1 = i2; i2 = ti;
    }
}

// merge the sorted top-8 lists of the G lanes of a group; afterwards every lane holds the group's top-8,
// ascending by (dist2, idx)
template <int G>
DFU_DEV void top8_merge_group(Top8& t) {
#pragma unroll
    for (int off = 1; off < G; off <<= 1) {
        float od[DFU_KNN];
        int oi[DFU_KNN];
#pragma unroll
        for (int k = 0; k < DFU_KNN; ++k) {  // partner's list, reversed
            od[k] = __shfl_xor_sync(0xffffffffu, t.d[DFU_KNN - 1 - k], off);
            oi[k] = __shfl_xor_sync(0xffffffffu, t.i[DFU_KNN - 1 - k], off);
        }
#pragma unroll
        for (int k = 0; k < DFU_KNN; ++k)  // the 8 smallest of the union, as a bitonic sequence
            if (lex_less(od[k], oi[k], t.d[k], t.i[k])) {
                t.d[k] = od[k];
                t.i[k] = oi[k];
            }
#pragma unroll
        for (int st = DFU_KNN / 2; st > 0; st >>= 1)  // bitonic merge network
#pragma unroll
            for (int k = 0; k < DFU_KNN; ++k)
                if ((k & st) == 0) lex_cswap(t.d[k], t.i[k], t.d[k + st], t.i[k + st]);
    }
}

template <int G>
DFU_DEV void knn8_scan_block_split(KnnSmem& sm, const float4* __restrict__ nodes, int Npad, float qx, float qy, float qz,
                                   bool active, int sub, Top8& t) {
    const int ntiles = (Npad + KNN_TILE - 1) / KNN_TILE;
    if (threadIdx.x == 0) {
        mbar_init(&sm.bar[0], 1);
        mbar_init(&sm.bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int s = 0; s < 2 && s < ntiles; ++s) {
            const int cnt = min(KNN_TILE, Npad - s * KNN_TILE);
            mbar_expect_tx(&sm.bar[s], (uint32_t) cnt * 16u);
            tma_bulk_g2s(sm.tile[s], nodes + (size_t) s * KNN_TILE, (uint32_t) cnt * 16u, &sm.bar[s]);
        }
    }
    top8_init(t);
    for (int tl = 0; tl < ntiles; ++tl) {
        const int s = tl & 1;
        const int cnt = min(KNN_TILE, Npad - tl * KNN_TILE);
        mbar_wait(&sm.bar[s], (uint32_t) ((tl >> 1) & 1));
        if (active) {
            const float4* __restrict__ tile = sm.tile[s];
            const int base = tl * KNN_TILE;
#pragma unroll 1
            for (int j = sub; j < cnt; j += 4 * G) {  // cnt % 32 == 0 and G | 8
                float dd[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float4 p = tile[j + u * G];
                    dd[u] = dist2(qx, qy, qz, p.x, p.y, p.z);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (dd[u] < t.d[DFU_KNN - 1]) top8_insert(t, dd[u], base + j + u * G);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0 && tl + 2 < ntiles) {
            const int c2 = min(KNN_TILE, Npad - (tl + 2) * KNN_TILE);
            fence_proxy_async();
            mbar_expect_tx(&sm.bar[s], (uint32_t) c2 * 16u);
            tma_bulk_g2s(sm.tile[s], nodes + (size_t) (tl + 2) * KNN_TILE, (uint32_t) c2 * 16u, &sm.bar[s]);
        }
    }
    top8_merge_group<G>(t);
}

}  // namespace dfu
