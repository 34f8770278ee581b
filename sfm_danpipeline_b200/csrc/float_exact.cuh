// Exact fp32 L2 2-NN kernel for float descriptors (SIFT 128-d) on sm_100a -- SFMM_FLOAT_EXACT.
//
// Replaces, for CV_32F descriptor sets, matcher->knnMatch(q, t, knn, 2) with
// cv::BFMatcher(cv::NORM_L2,false) at /root/reference/src/Sfm.cpp:593,599, i.e. OpenCV's
// batchDistance(K=2) over  d = sqrtf( sum_k (a_k - b_k)^2 )  in fp32, direct-difference form
// (normL2Sqr_), which is what this kernel evaluates for EVERY (query, train) combination:
// no |a|^2+|b|^2-2ab cancellation, so distances agree with OpenCV to the last ulp or two
// (summation order) and are bit-identical on integer-valued data such as real SIFT output.
// It is the reference-accurate mode and the refinement arithmetic of the tensor-core mode
// (float_tensor.cuh ranks with tcgen05 TF32 and re-evaluates only the winners this way).
//
// Tiling: CTA = BQ(64) query rows x the whole train range in BT(64)-row stages, 256 threads as
// a 16x16 grid, 4x4 register micro-tile per thread, rows in shared memory with a 16-byte
// padded pitch (conflict-free LDS.128: 33 quad-words per row for 128-d), train stages double
// buffered with cp.async.  Top-2 is private per thread on 64-bit (float bits << 32 | index)
// keys and merged across the 16 threads sharing a query row with warp shuffles at the end.
#pragma once
#include "common.cuh"

namespace sfmm {

static constexpr int FX_BQ = 64;
static constexpr int FX_BT = 64;
static constexpr int FX_THREADS = 256;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(
                     static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))),
                 "l"(gmem_src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void top2_insert(unsigned long long& k1, unsigned long long& k2, unsigned long long key) {
    const unsigned long long hi = max(k1, key);
    k1 = min(k1, key);
    k2 = min(k2, hi);
}

// dynamic smem: float q[FX_BQ][KP]; float t[2][FX_BT][KP];   KP = kq*4 + 4
static inline size_t float_exact_smem_bytes(int kq /* quad-words per row */) {
    return (size_t)(FX_BQ + 2 * FX_BT) * (kq * 4 + 4) * sizeof(float);
}

__global__ void __launch_bounds__(FX_THREADS, 2)
float_exact_knn2_kernel(const float* __restrict__ blob, int kq /* row pitch in float4 */,
                        const KnnTile* __restrict__ tiles, const PairDesc* __restrict__ pairs,
                        KnnEntry* __restrict__ knn, unsigned long long* __restrict__ colmin, int cross,
                        const uint32_t* __restrict__ n_items_dev /* non-NULL: the tile count lives on the device */,
                        const uint32_t* __restrict__ xcand, const uint32_t* __restrict__ n_xcand /* candidate train rows per pair (filter.cuh) */) {
    extern __shared__ __align__(16) float fx_smem[];
    const int KP = kq * 4 + 4;  // padded pitch in floats
    float* sq = fx_smem;
    float* st = fx_smem + FX_BQ * KP;

    if (n_items_dev && blockIdx.x >= __ldg(n_items_dev)) return;
    KnnTile tile = tiles[blockIdx.x];
    PairDesc pd = pairs[tile.pair];
    // "reverse + gather" tile (cross-check through candidate columns, filter.cuh): the tile's rows are candidate train rows
    // addressed through the pair's list, the whole query image is streamed past them, and each row's nearest neighbour is the
    // forward problem's column minimum (lowest query index on ties) -> colmin instead of a 2-NN entry.
    const bool rgather = (tile.split & (TILE_REVERSE | TILE_GATHER)) == (TILE_REVERSE | TILE_GATHER);
    tile.split &= TILE_SPLIT_MASK;
    const uint32_t* list = nullptr;
    if (rgather) {
        const uint32_t r0 = pd.q_row0;
        pd.q_row0 = pd.t_row0; pd.t_row0 = r0;
        pd.nt = pd.nq;
        pd.nq = __ldg(n_xcand + tile.pair);
        list = xcand + pd.col_off;
    }
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;

    const float4* gq = reinterpret_cast<const float4*>(blob) + (size_t)pd.q_row0 * kq;
    const float4* gt = reinterpret_cast<const float4*>(blob) + (size_t)pd.t_row0 * kq;

    // query tile (clamped rows) -> smem
    for (int i = tid; i < FX_BQ * kq; i += FX_THREADS) {
        const int r = i / kq, c = i - r * kq;
        uint32_t row = min(tile.q0 + r, pd.nq - 1);
        if (rgather) row = __ldg(list + row);
        cp_async16(sq + r * KP + c * 4, gq + (size_t)row * kq + c);
    }
    const uint32_t n_rows = tile.t1 - tile.t0;
    const uint32_t n_stages = (n_rows + FX_BT - 1) / FX_BT;
    auto load_stage = [&](uint32_t s) {
        float* dst = st + (s & 1) * FX_BT * KP;
        const uint32_t rows = min((uint32_t)FX_BT, n_rows - s * FX_BT);
        for (int i = tid; i < (int)rows * kq; i += FX_THREADS) {
            const int r = i / kq, c = i - r * kq;
            cp_async16(dst + r * KP + c * 4, gt + (size_t)(tile.t0 + s * FX_BT + r) * kq + c);
        }
    };
    if (n_stages > 0) load_stage(0);
    cp_async_commit();

    unsigned long long k1[4], k2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) k1[i] = k2[i] = KEY_NONE;

    for (uint32_t s = 0; s < n_stages; ++s) {
        if (s + 1 < n_stages) load_stage(s + 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();

        const float* tq = sq + (ty * 4) * KP;
        const float* tt = st + (s & 1) * FX_BT * KP + tx * KP;
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 2
        for (int c = 0; c < kq; ++c) {
            float4 a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(tq + i * KP + c * 4);
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(tt + (16 * j) * KP + c * 4);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float d;
                    d = a[i].x - b[j].x; acc[i][j] = fmaf(d, d, acc[i][j]);
                    d = a[i].y - b[j].y; acc[i][j] = fmaf(d, d, acc[i][j]);
                    d = a[i].z - b[j].z; acc[i][j] = fmaf(d, d, acc[i][j]);
                    d = a[i].w - b[j].w; acc[i][j] = fmaf(d, d, acc[i][j]);
                }
        }
        const uint32_t tbase = tile.t0 + s * FX_BT;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t t = tbase + tx + 16 * j;
            const bool tvalid = t < tile.t1;
            unsigned long long cmin = KEY_NONE;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t dbits = __float_as_uint(sqrtf(acc[i][j]));
                if (tvalid) top2_insert(k1[i], k2[i], make_key(dbits, t));
                const uint32_t qr = tile.q0 + ty * 4 + i;
                if (qr < pd.nq) cmin = min(cmin, make_key(dbits, qr));
            }
            if (cross) {  // kernel-uniform branch: the shuffle is executed by the whole warp
                // lanes 0-15 / 16-31 hold different query rows for the same 16 train rows
                const unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, cmin, 16);
                cmin = min(cmin, o);
                if (tvalid && (tid & 16) == 0 && cmin != KEY_NONE) atomicMin(colmin + pd.col_off + t, cmin);
            }
        }
        __syncthreads();
    }
    cp_async_wait<0>();

    // merge the 16 partial lists of each query row (lanes sharing ty: xor 1,2,4,8)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int d = 1; d < 16; d <<= 1) {
            const unsigned long long o1 = __shfl_xor_sync(0xFFFFFFFFu, k1[i], d);
            const unsigned long long o2 = __shfl_xor_sync(0xFFFFFFFFu, k2[i], d);
            top2_insert(k1[i], k2[i], o1);
            k2[i] = min(k2[i], o2);
        }
        const uint32_t qr = tile.q0 + ty * 4 + i;
        if (tx == 0 && qr < pd.nq) {
            if (rgather) {
                if (k1[i] != KEY_NONE) atomicMin(colmin + pd.col_off + __ldg(list + qr), k1[i]);
            } else {
                KnnEntry e;
                e.x = k1[i];
                e.y = k2[i];
                knn[pd.knn_off + (size_t)tile.split * pd.nq + qr] = e;
            }
        }
    }
}

}  // namespace sfmm
