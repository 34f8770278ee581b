// Ratio test, optional symmetric cross-check and ordered compaction of the 2-NN lists.
//
// Replaces the loop at /root/reference/src/Sfm.cpp:603-607
//     if (knnMatches[i][0].distance <= NN_MATCH_RATIO * knnMatches[i][1].distance)
//         goodMatches->push_back(knnMatches[i][0]);
// for a whole launch of image pairs at once.  The comparison is the reference's: fp32 product
// (no fma, no widening), '<='; Hamming distances are converted int -> float first, exactly as
// cv::BFMatcher::knnMatchImpl does before the reference sees them.  Output records are
// cv::DMatch-layout, ascending queryIdx inside a pair, pairs back to back in launch order.
//
// Three tiny kernels (all HBM-bound, 16 B read per query row):
//   filter_count  : per filter tile (FILTER_TILE query rows) number of surviving rows
//   tile_scan     : exclusive scan of the tile counts (single CTA) + grand total
//   filter_write  : re-evaluate, block-scan, write SfmDMatch records + per-pair (offset,count)
#pragma once
#include "common.cuh"
#include "../../include/sfm_match.h"

namespace sfmm {

static constexpr int FILTER_THREADS = 256;
static constexpr int FILTER_PER_THREAD = FILTER_TILE / FILTER_THREADS;

// Merge the n_splits partial lists of one query row into its two best keys.
__device__ __forceinline__ void merged_top2(const KnnEntry* __restrict__ knn, const PairDesc& pd, uint32_t q,
                                            unsigned long long& k1, unsigned long long& k2) {
    k1 = k2 = KEY_NONE;
    for (uint32_t s = 0; s < pd.n_splits; ++s) {
        const KnnEntry e = knn[pd.knn_off + (size_t)s * pd.nq + q];
        // e.x <= e.y; insert both
        unsigned long long hi = max(k1, e.x);
        k1 = min(k1, e.x);
        k2 = min(k2, hi);
        k2 = min(k2, e.y);
    }
}

template <bool IS_FLOAT>
__device__ __forceinline__ float key_distance(unsigned long long key) {
    const uint32_t hi = static_cast<uint32_t>(key >> 32);
    if constexpr (IS_FLOAT) return __uint_as_float(hi);
    else return static_cast<float>(static_cast<int32_t>(hi));
}

// Does query row q of pair pd survive, and with which record?
template <bool IS_FLOAT, bool CROSS>
__device__ __forceinline__ bool evaluate_row(const KnnEntry* __restrict__ knn,
                                             const unsigned long long* __restrict__ colmin, const PairDesc& pd,
                                             uint32_t q, float ratio, SfmDMatch& m) {
    unsigned long long k1, k2;
    merged_top2(knn, pd, q, k1, k2);
    if (k2 == KEY_NONE) return false;  // fewer than two train rows: no ratio test, no match
    const float d1 = key_distance<IS_FLOAT>(k1);
    const float d2 = key_distance<IS_FLOAT>(k2);
    if (!(d1 <= __fmul_rn(ratio, d2))) return false;
    const uint32_t t = static_cast<uint32_t>(k1);
    if constexpr (CROSS) {
        // mutual nearest neighbour: q must be the lowest-index argmin over q' of d(q', t)
        if (static_cast<uint32_t>(colmin[pd.col_off + t]) != q) return false;
    }
    m.queryIdx = static_cast<int32_t>(q);
    m.trainIdx = static_cast<int32_t>(t);
    m.imgIdx = 0;
    m.distance = d1;
    return true;
}

template <bool IS_FLOAT, bool CROSS>
__global__ void __launch_bounds__(FILTER_THREADS)
filter_count_kernel(const FilterTile* __restrict__ ftiles, const PairDesc* __restrict__ pairs,
                    const KnnEntry* __restrict__ knn, const unsigned long long* __restrict__ colmin, float ratio,
                    uint32_t* __restrict__ tile_count) {
    const FilterTile ft = ftiles[blockIdx.x];
    const PairDesc pd = pairs[ft.pair];
    uint32_t n = 0;
#pragma unroll
    for (int k = 0; k < FILTER_PER_THREAD; ++k) {
        const uint32_t q = ft.q0 + threadIdx.x * FILTER_PER_THREAD + k;
        SfmDMatch m;
        if (q < pd.nq && evaluate_row<IS_FLOAT, CROSS>(knn, colmin, pd, q, ratio, m)) ++n;
    }
    n = __reduce_add_sync(0xFFFFFFFFu, n);
    __shared__ uint32_t warp_sum[FILTER_THREADS / 32];
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = n;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t s = 0;
#pragma unroll
        for (int w = 0; w < FILTER_THREADS / 32; ++w) s += warp_sum[w];
        tile_count[blockIdx.x] = s;
    }
}

// Exclusive scan of n tile counts into tile_off[0..n] (tile_off[n] = total).  One CTA.
static constexpr int SCAN_THREADS = 1024;
__global__ void __launch_bounds__(SCAN_THREADS)
tile_scan_kernel(const uint32_t* __restrict__ tile_count, unsigned long long* __restrict__ tile_off, uint32_t n) {
    __shared__ unsigned long long warp_tot[SCAN_THREADS / 32];
    __shared__ unsigned long long carry_sh;
    if (threadIdx.x == 0) carry_sh = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += SCAN_THREADS) {
        const uint32_t i = base + threadIdx.x;
        const unsigned long long v = i < n ? tile_count[i] : 0;
        unsigned long long incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if ((threadIdx.x & 31) >= d) incl += o;
        }
        if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
        __syncthreads();
        unsigned long long wbase = 0;
        for (uint32_t w = 0; w < (threadIdx.x >> 5); ++w) wbase += warp_tot[w];
        const unsigned long long carry = carry_sh;
        if (i < n) tile_off[i] = carry + wbase + incl - v;
        __syncthreads();
        if (threadIdx.x == SCAN_THREADS - 1) carry_sh = carry + wbase + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_off[n] = carry_sh;
}

template <bool IS_FLOAT, bool CROSS>
__global__ void __launch_bounds__(FILTER_THREADS)
filter_write_kernel(const FilterTile* __restrict__ ftiles, const PairDesc* __restrict__ pairs,
                    const KnnEntry* __restrict__ knn, const unsigned long long* __restrict__ colmin, float ratio,
                    const unsigned long long* __restrict__ tile_off, SfmDMatch* __restrict__ out,
                    unsigned long long out_capacity, int32_t* __restrict__ pair_count,
                    unsigned long long* __restrict__ pair_off, const double2* __restrict__ points,
                    double2* __restrict__ out_left, double2* __restrict__ out_right) {
    const FilterTile ft = ftiles[blockIdx.x];
    const PairDesc pd = pairs[ft.pair];
    SfmDMatch m[FILTER_PER_THREAD];
    bool keep[FILTER_PER_THREAD];
    uint32_t n = 0;
#pragma unroll
    for (int k = 0; k < FILTER_PER_THREAD; ++k) {
        const uint32_t q = ft.q0 + threadIdx.x * FILTER_PER_THREAD + k;
        keep[k] = q < pd.nq && evaluate_row<IS_FLOAT, CROSS>(knn, colmin, pd, q, ratio, m[k]);
        n += keep[k];
    }
    // exclusive scan of n over the block (warp shuffle + one smem hop)
    uint32_t incl = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if ((threadIdx.x & 31) >= d) incl += o;
    }
    __shared__ uint32_t warp_tot[FILTER_THREADS / 32];
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t wbase = 0;
    for (uint32_t w = 0; w < (threadIdx.x >> 5); ++w) wbase += warp_tot[w];
    unsigned long long pos = tile_off[blockIdx.x] + wbase + incl - n;
#pragma unroll
    for (int k = 0; k < FILTER_PER_THREAD; ++k)
        if (keep[k]) {
            if (pos < out_capacity) {
                out[pos] = m[k];
                if (points) {
                    // AlignedPointsFromMatch (/root/reference/src/Sfm.cpp:694-711) fused into the compaction:
                    // alignedL[i] = imagesPts2D[q][queryIdx], alignedR[i] = imagesPts2D[t][trainIdx]
                    out_left[pos] = points[pd.q_row0 + m[k].queryIdx];
                    out_right[pos] = points[pd.t_row0 + m[k].trainIdx];
                }
            }
            ++pos;
        }
    // the first tile of a pair publishes the pair's segment
    if (threadIdx.x == 0 && blockIdx.x == pd.first_ftile) {
        const unsigned long long b = tile_off[pd.first_ftile];
        const unsigned long long e = tile_off[pd.first_ftile + pd.n_ftiles];
        pair_off[ft.pair] = b;
        pair_count[ft.pair] = static_cast<int32_t>(e - b);
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Symmetric cross-check without a second full pass (tensor kernels).  Only the train rows that some query row of the pair
// selected as its nearest neighbour AND that passed the ratio test can end up in the match list, so only THEIR column minima
// are needed: typically an eighth of the train image instead of all of it.
//   cross_mark    : every query row that passes the ratio test flags its nearest train row        (flags: one byte per train row)
//   cross_compact : per pair, the flagged rows in ascending order -> candidate list; the pair's "reverse" work items (128
//                   candidate rows each, bit 30/31 of KnnTile::split) are appended to the launch's reverse item list
// The 2-NN kernel then runs the reverse items only (rows gathered through the list, the whole query image streamed past
// them) and the filter looks the minima up as before.  Cost: matches/Nt of a forward pass instead of a whole one.
template <bool IS_FLOAT>
__global__ void __launch_bounds__(FILTER_THREADS)
cross_mark_kernel(const FilterTile* __restrict__ ftiles, const PairDesc* __restrict__ pairs, const KnnEntry* __restrict__ knn, float ratio,
                  unsigned char* __restrict__ flags) {
    const FilterTile ft = ftiles[blockIdx.x];
    const PairDesc pd = pairs[ft.pair];
#pragma unroll
    for (int k = 0; k < FILTER_PER_THREAD; ++k) {
        const uint32_t q = ft.q0 + threadIdx.x * FILTER_PER_THREAD + k;
        if (q >= pd.nq) continue;
        unsigned long long k1, k2;
        merged_top2(knn, pd, q, k1, k2);
        if (k2 == KEY_NONE) continue;
        if (!(key_distance<IS_FLOAT>(k1) <= __fmul_rn(ratio, key_distance<IS_FLOAT>(k2)))) continue;
        flags[pd.col_off + static_cast<uint32_t>(k1)] = 1;  // (several rows may write the same byte: same value)
    }
}

// One CTA per pair.  cand[col_off ..] receives the flagged train rows in ascending order, n_cand[pair] their number.
static constexpr int COMPACT_THREADS = 256;
__global__ void __launch_bounds__(COMPACT_THREADS)
cross_compact_kernel(const PairDesc* __restrict__ pairs, const unsigned char* __restrict__ flags, uint32_t* __restrict__ cand,
                     uint32_t* __restrict__ n_cand, KnnTile* __restrict__ rtiles, uint32_t* __restrict__ n_rtiles, uint32_t rows_per_item) {
    PairDesc pd = pairs[blockIdx.x];
    if (pd.n_splits == 0) pd.nt = 0;  // a pair without work (no query rows / fewer than two train rows) owns no flags
    __shared__ uint32_t warp_tot[COMPACT_THREADS / 32];
    __shared__ uint32_t carry_sh, first_item;
    if (threadIdx.x == 0) carry_sh = 0;
    __syncthreads();
    for (uint32_t base = 0; base < pd.nt; base += COMPACT_THREADS) {
        const uint32_t t = base + threadIdx.x;
        const uint32_t f = (t < pd.nt && flags[pd.col_off + t]) ? 1u : 0u;
        const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, f);
        const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        if (lane == 0) warp_tot[w] = __popc(ballot);
        __syncthreads();
        uint32_t wbase = 0;
        for (uint32_t i = 0; i < w; ++i) wbase += warp_tot[i];
        const uint32_t carry = carry_sh;
        if (f) cand[pd.col_off + carry + wbase + __popc(ballot & ((1u << lane) - 1u))] = t;
        __syncthreads();
        if (threadIdx.x == COMPACT_THREADS - 1) carry_sh = carry + wbase + __popc(ballot);
        __syncthreads();
    }
    const uint32_t nc = carry_sh;
    const uint32_t items = (nc + rows_per_item - 1) / rows_per_item;
    if (threadIdx.x == 0) {
        n_cand[blockIdx.x] = nc;
        first_item = items ? atomicAdd(n_rtiles, items) : 0u;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < items; i += COMPACT_THREADS) {
        KnnTile kt;
        kt.pair = blockIdx.x;
        kt.q0 = i * rows_per_item;  // first candidate-list entry of the item
        kt.t0 = 0;                  // the whole query image is streamed past the candidates
        kt.t1 = pd.nq;
        kt.split = TILE_REVERSE | TILE_GATHER;
        rtiles[first_item + i] = kt;
    }
}

// Raw merged 2-NN list of ONE pair as (train index, float distance) arrays -- what
// cv::BFMatcher::knnMatch(k=2) returns (src/Sfm.cpp:599); for parity tests (sfmm_knn_pair).
template <bool IS_FLOAT>
__global__ void knn_decode_kernel(const PairDesc* __restrict__ pairs, const KnnEntry* __restrict__ knn,
                                  int32_t* __restrict__ idx, float* __restrict__ dist) {
    const PairDesc pd = pairs[0];
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= pd.nq) return;
    unsigned long long k[2];
    merged_top2(knn, pd, q, k[0], k[1]);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const bool none = k[j] == KEY_NONE;
        idx[2 * q + j] = none ? -1 : static_cast<int32_t>(static_cast<uint32_t>(k[j]));
        dist[2 * q + j] = none ? 3.402823466e+38f : key_distance<IS_FLOAT>(k[j]);
    }
}

}  // namespace sfmm
