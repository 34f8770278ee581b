// Hamming 2-NN kernel for binary descriptors (AKAZE-MLDB 61 B, ORB 32 B) on sm_100a.
//
// Replaces, for CV_8U descriptor sets, the body of
//     matcher->knnMatch(query_descriptor, train_descriptor, knnMatches, 2)
// at /root/reference/src/Sfm.cpp:599 (cv::BFMatcher -> cv::batchDistance(K=2) ->
// hal::normHamming in OpenCV): every query row against every train row, keep the two smallest
// (distance, train index) lexicographically.
//
// Shape of the computation (compute bound, not HBM bound: a 20k x 20k pair reads 2.6 MB and
// does 6.4e9 32-bit popcounts):
//   * one thread owns TQ query rows (2 for 512-bit rows), held in registers for the whole tile;
//   * the CTA streams the train rows through shared memory in TT-row stages filled by the
//     TMA bulk-copy engine (cp.async.bulk + mbarrier, 2 stages) -- train rows are contiguous
//     in the blob, so a stage is one linear copy;
//   * every lane of a warp reads the SAME train word (LDS.128 broadcast, conflict free), XORs
//     it against its own query words, and reduces the XOR words to a bit count;
//   * the bit count is the bottleneck: POPC issues at 16 lanes/clk/SM, a quarter of the
//     LOP3/IADD3 rate.  CSA_LEVEL > 0 therefore compresses the W XOR words with carry-save
//     adders (two LOP3 per 3:2 compressor) before counting -- e.g. 16 words -> 2 "ones" + 7
//     "twos" words = 9 POPC instead of 16 -- trading POPC-pipe slots for ALU-pipe slots;
//   * top-2 is a branch-free min/max network on a packed 32-bit key (distance << 18 | index),
//     private to the owning thread: no shuffles, no atomics, and ties resolve to the lowest
//     train index exactly like OpenCV's strict-'<' insertion.
//   * CROSS adds the other direction of the symmetric cross-check from the same distances:
//     per train row the minimum (distance << 18 | query index) over the CTA's query rows,
//     reduced warp-wide with REDUX and merged through shared memory into a global 64-bit
//     atomicMin per (CTA, train row).
#pragma once
#include "common.cuh"

namespace sfmm {

// ------------------------------------------------------------------ PTX helpers (TMA bulk copy)
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// Same, for roles that wait for a long time off the critical path (item prefetch, query loaders): back off
// between polls so that the spin does not take issue slots from the epilogue warps of the same sub-partition
// (ncu of the fp16 kernel: 16 % of all executed instructions were try_wait spins).
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    for (;;) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(256);
    }
}
// Linear global -> shared copy by the TMA engine; completion is signalled on `bar` (complete_tx).
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ------------------------------------------------------------------ bit counting
// 3:2 compressor on 32 bit lanes: sum = a^b^c (LOP3 0x96), carry = majority(a,b,c) (LOP3 0xE8).
// Written as PTX so that neither NVVM nor ptxas re-associates the XOR of query and train word
// into the compressor (that costs 6 instead of 5 LOP3 per three words, and the ALU pipe is the
// binding one: ncu profiles/ncu_binary_r01.txt shows alu 94 %, xu 79 %, fma 3 %).
__device__ __forceinline__ void csa(uint32_t a, uint32_t b, uint32_t c, uint32_t& sum, uint32_t& carry) {
    uint32_t s, m;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(s) : "r"(a), "r"(b), "r"(c));
    asm("lop3.b32 %0, %1, %2, %3, 0xe8;" : "=r"(m) : "r"(a), "r"(b), "r"(c));
    sum = s;
    carry = m;
}

// One compression pass: N words of equal weight -> N - 2*(N/3) words of that weight (in place
// at the front of `same`) + N/3 words of twice the weight appended to `carry`.
template <int N>
__device__ __forceinline__ void csa_pass(uint32_t (&same)[N], uint32_t* carry) {
    constexpr int G = N / 3;
#pragma unroll
    for (int g = 0; g < G; ++g) csa(same[3 * g], same[3 * g + 1], same[3 * g + 2], same[g], carry[g]);
#pragma unroll
    for (int r = 0; r < N - 3 * G; ++r) same[G + r] = same[3 * G + r];
}

// acc + sum_i popc(x[i]) * weight.  `weight` comes from kernel parameters (constant bank), so
// the multiply-add stays an IMAD on the otherwise idle FMA pipe instead of being strength-
// reduced to shift/add instructions on the ALU pipe.
template <int N>
__device__ __forceinline__ uint32_t popc_mad(const uint32_t* x, uint32_t weight, uint32_t acc) {
#pragma unroll
    for (int i = 0; i < N; ++i) asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc) : "r"(__popc(x[i])), "r"(weight));
    return acc;
}

// Weights of the bit-count planes already shifted into key position: w[k] = 2^k << IDX_BITS.
struct KeyWeights {
    uint32_t w[3];
};

// Packed key  (hamming(q,b) << IDX_BITS) + t  of two W-word rows.
//   CSA_LEVEL 0: W POPC.
//   CSA_LEVEL 1: one pass on the XOR words                          (W=16: 6 ones + 5 twos = 11 POPC)
//   CSA_LEVEL 2: + a second pass on the surviving "ones"            (W=16: 2 ones + 7 twos =  9 POPC)
//   CSA_LEVEL 3: + a pass on the "twos"                             (W=16: 2 + 3 twos + 2 fours = 7 POPC)
// Pipe budget per 32 evaluations at level 2, W=16: 30 LOP3 + 3 VIMNMX on ALU (66 cycles/SMSP),
// 9 POPC on XU (72 cycles), 9 IMAD on FMA (18 cycles)  ->  16/9 of the plain-POPC rate.
template <int W, int CSA_LEVEL>
__device__ __forceinline__ uint32_t hamming_key(const uint32_t (&q)[W], const uint32_t (&b)[W], uint32_t t,
                                                const KeyWeights& kw) {
    uint32_t x[W];
#pragma unroll
    for (int j = 0; j < W; ++j) x[j] = q[j] ^ b[j];
    if constexpr (CSA_LEVEL == 0 || W < 3) {
        return popc_mad<W>(x, kw.w[0], t);
    } else {
        constexpr int G1 = W / 3;         // carries of pass 1
        constexpr int N1 = W - 2 * G1;    // ones left after pass 1
        constexpr int G2 = (CSA_LEVEL >= 2 && N1 >= 3) ? N1 / 3 : 0;
        constexpr int N2 = N1 - 2 * G2;   // ones left after pass 2
        constexpr int NT = G1 + G2;       // twos
        uint32_t twos[NT > 0 ? NT : 1];
        csa_pass<W>(x, twos);
        if constexpr (G2 > 0) {
            uint32_t ones1[N1];
#pragma unroll
            for (int i = 0; i < N1; ++i) ones1[i] = x[i];
            csa_pass<N1>(ones1, twos + G1);
#pragma unroll
            for (int i = 0; i < N2; ++i) x[i] = ones1[i];
        }
        uint32_t key = popc_mad<N2>(x, kw.w[0], t);
        if constexpr (CSA_LEVEL >= 3 && NT >= 3) {
            constexpr int G3 = NT / 3;
            constexpr int NT2 = NT - 2 * G3;
            uint32_t fours[G3];
            uint32_t t2[NT];
#pragma unroll
            for (int i = 0; i < NT; ++i) t2[i] = twos[i];
            csa_pass<NT>(t2, fours);
            key = popc_mad<NT2>(t2, kw.w[1], key);
            return popc_mad<G3>(fours, kw.w[2], key);
        } else {
            return popc_mad<NT>(twos, kw.w[1], key);
        }
    }
}

// ------------------------------------------------------------------ the kernel
template <int W, int TQ, int THREADS, int TT>
struct BinaryKnnSmem {
    alignas(128) uint32_t stage[2][TT * W];
    alignas(8) uint64_t full[2];
    uint32_t colmin[2][TT];  // CROSS only: per-stage column minima (distance << 18 | query index)
};

#ifndef BK_UNROLL
#define BK_UNROLL 2
#endif
static constexpr int BK_UNROLL_N = BK_UNROLL;  // train rows per unrolled step of the inner loop
template <int W, int TQ, int THREADS, int TT, int CSA_LEVEL, bool CROSS>
__global__ void __launch_bounds__(THREADS, (W >= 32) ? 4 : 7)  // <= 72 registers: 28 warps/SM
binary_knn2_kernel(const uint32_t* __restrict__ blob, const KnnTile* __restrict__ tiles,
                   const PairDesc* __restrict__ pairs, KnnEntry* __restrict__ knn,
                   unsigned long long* __restrict__ colmin, const KeyWeights kw) {
    static_assert(W % 4 == 0, "rows are 16-byte multiples");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    auto& sm = *reinterpret_cast<BinaryKnnSmem<W, TQ, THREADS, TT>*>(smem_raw);

    const KnnTile tile = tiles[blockIdx.x];
    const PairDesc pd = pairs[tile.pair];
    const uint32_t tid = threadIdx.x;

    // ---- query rows -> registers (row k of this thread = q0 + k*THREADS + tid: a warp reads
    //      32 consecutive rows = one contiguous 32*W*4-byte span, 128-bit loads)
    uint32_t q[TQ][W];
    uint32_t qrow[TQ];
#pragma unroll
    for (int k = 0; k < TQ; ++k) {
        qrow[k] = tile.q0 + k * THREADS + tid;
        const uint32_t r = min(qrow[k], pd.nq - 1);  // clamp: out-of-range rows compute but never store
        const uint4* src = reinterpret_cast<const uint4*>(blob + (size_t)(pd.q_row0 + r) * W);
#pragma unroll
        for (int j = 0; j < W / 4; ++j) {
            const uint4 v = __ldg(src + j);
            q[k][4 * j + 0] = v.x; q[k][4 * j + 1] = v.y; q[k][4 * j + 2] = v.z; q[k][4 * j + 3] = v.w;
        }
    }

    uint32_t m1[TQ], m2[TQ];
#pragma unroll
    for (int k = 0; k < TQ; ++k) m1[k] = m2[k] = 0xFFFFFFFFu;

    // ---- train rows: 2-stage TMA pipeline
    const uint32_t n_rows = tile.t1 - tile.t0;
    const uint32_t n_stages = (n_rows + TT - 1) / TT;
    const uint32_t* train = blob + (size_t)(pd.t_row0 + tile.t0) * W;

    if (tid == 0) {
        mbar_init(&sm.full[0], 1);
        mbar_init(&sm.full[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](uint32_t i) {
        const uint32_t rows = min((uint32_t)TT, n_rows - i * TT);
        const uint32_t bytes = rows * W * 4;
        mbar_expect_tx(&sm.full[i & 1], bytes);
        tma_load_1d(sm.stage[i & 1], train + (size_t)i * TT * W, bytes, &sm.full[i & 1]);
    };
    if (tid == 0) {
        if (n_stages > 0) issue(0);
        if (n_stages > 1) issue(1);
    }

    for (uint32_t i = 0; i < n_stages; ++i) {
        const uint32_t s = i & 1;
        if constexpr (CROSS) {
            for (uint32_t r = tid; r < TT; r += THREADS) sm.colmin[s][r] = 0xFFFFFFFFu;
            __syncthreads();
        }
        mbar_wait(&sm.full[s], (i >> 1) & 1);
        const uint32_t rows = min((uint32_t)TT, n_rows - i * TT);
        const uint32_t tbase = tile.t0 + i * TT;
        const uint4* st = reinterpret_cast<const uint4*>(sm.stage[s]);
#pragma unroll BK_UNROLL_N
        for (uint32_t r = 0; r < rows; ++r) {
            uint32_t b[W];
#pragma unroll
            for (int j = 0; j < W / 4; ++j) {
                const uint4 v = st[r * (W / 4) + j];  // same address in every lane: broadcast
                b[4 * j + 0] = v.x; b[4 * j + 1] = v.y; b[4 * j + 2] = v.z; b[4 * j + 3] = v.w;
            }
            const uint32_t t = tbase + r;
            [[maybe_unused]] uint32_t cmin = 0xFFFFFFFFu;
#pragma unroll
            for (int k = 0; k < TQ; ++k) {
                const uint32_t key = hamming_key<W, CSA_LEVEL>(q[k], b, t, kw);
                const uint32_t hi = max(m1[k], key);
                m1[k] = min(m1[k], key);
                m2[k] = min(m2[k], hi);
                if constexpr (CROSS) {
                    // same distance, query index in the low bits instead of the train index
                    const uint32_t ck = qrow[k] < pd.nq ? key - t + qrow[k] : 0xFFFFFFFFu;
                    cmin = min(cmin, ck);
                }
            }
            if constexpr (CROSS) {
                cmin = __reduce_min_sync(0xFFFFFFFFu, cmin);
                if ((tid & 31) == 0) atomicMin(&sm.colmin[s][r], cmin);
            }
        }
        __syncthreads();  // every warp is done with stage s (and its colmin row is complete)
        if constexpr (CROSS) {
            for (uint32_t r = tid; r < rows; r += THREADS) {
                const uint32_t c = sm.colmin[s][r];
                if (c != 0xFFFFFFFFu)
                    atomicMin(colmin + pd.col_off + tbase + r,
                              make_key(c >> IDX_BITS, c & ((1u << IDX_BITS) - 1)));
            }
        }
        if (tid == 0 && i + 2 < n_stages) issue(i + 2);
    }

    // ---- partial 2-NN list of this (tile, split)
#pragma unroll
    for (int k = 0; k < TQ; ++k) {
        if (qrow[k] < pd.nq) {
            KnnEntry e;
            e.x = m1[k] == 0xFFFFFFFFu ? KEY_NONE : make_key(m1[k] >> IDX_BITS, m1[k] & ((1u << IDX_BITS) - 1));
            e.y = m2[k] == 0xFFFFFFFFu ? KEY_NONE : make_key(m2[k] >> IDX_BITS, m2[k] & ((1u << IDX_BITS) - 1));
            knn[pd.knn_off + (size_t)tile.split * pd.nq + qrow[k]] = e;
        }
    }
}

}  // namespace sfmm
