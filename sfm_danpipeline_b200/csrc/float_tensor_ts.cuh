// Tensor-core 2-NN kernel, "TS" form: the QUERY tile lives in tensor memory, only train tiles go through
// shared memory.  Same maths, same epilogue pieces and the same results as tensor_knn2_kernel (float_tensor.cuh).
// Default for the binary engine (the FP4 modes TM_F4P / TM_F4X, which exist in this form only; TM_I8P, and TM_I8 up to 256 bit) and
// the fp16 float path (TM_F16X / TM_F16_EXACT), where it measured faster; TF32 modes default to the shared-memory-A kernel (SFMM_TENSOR_TS=0/1 forces either;
// profiles/tensor_variants_r01.txt has every A/B).
//
// Why: with both operands in shared memory every tcgen05.mma (M=128, N=128, 32 bytes of K) reads 4 KB of A
// and 4 KB of B in its 64-cycle slot = 128 B/clk, which is ALL of the SM's shared-memory bandwidth
// (128 B/clk/SM, B300_MICROARCH.md), and the TMA engine has to write the next train tile into the same
// memory (64 KB per 1024-cycle tile = another 64 B/clk).  192 B/clk of demand on a 128 B/clk port predicts
// 1536 cycles per tile; the SS kernel measured 1430-1575 (tensor pipe 66-76 % active, ncu
// profiles/ncu_tensor_persist_r01.txt) whatever else was tuned.  Reading A from TMEM removes 64 B/clk.
// It also frees the 64 KB query-tile buffer: the train ring grows from 2 to 3 stages (more TMA latency
// hidden), and the query tile of the NEXT item is loaded while the current one computes (two A buffers in
// TMEM), which removes the pipeline bubble at item boundaries.  The price is TMEM: two accumulator stages
// instead of four, so the epilogue must release a stage within one MMA tile time -- which it does only once
// its per-column work is small (one IMAD + 1.25 VIMNMX.U16x2 for binary; profiles/ncu_tensor_ts_r01.txt shows
// the first version with the MMA warp 30 % of its time in the acc_empty wait).
//
// TMEM map (512 columns): query tile A0 | A1 (KB*32 columns each) | 128-column accumulators, two (KB = 3, 4) or three (KB <= 2).
// (Round 2: the FP4 modes and the skipping TM_F16X kernel keep ONE query-tile buffer and always three accumulators; the FP4 modes
// add 32 columns of unit scale factors behind them -- see A_BUFS in the kernel.)
// A row r of the tile is TMEM lane r; its K bytes are packed in order into 32-bit columns (32 bytes = 8
// columns per MMA), which is exactly what a thread gets when it reads its row from global memory as
// 32-bit words -- so four loader warps (one per TMEM lane quarter) copy rows global -> registers ->
// tcgen05.st, no swizzle and no staging buffer.
//
//   warp 0       TMA producer  : train tiles (3-stage ring) + their norms (4-slot ring)
//   warp 1       MMA issuer    : tcgen05.mma [d_tmem], [a_tmem], b_desc  (A from TMEM)
//   warp 2       TMEM allocator
//   warp 3       item prefetch : next KnnTile/PairDesc + per-row constants -> 2-slot smem ring
//   warps 4..    epilogue      : GROUPS (2 or 4) groups of four warps taking tiles round robin
//   last 4 warps query loaders : global -> registers -> tcgen05.st, one item ahead
#pragma once
#include "float_tensor.cuh"

namespace sfmm {

static constexpr int FTS_B_STAGES = 3;
static constexpr int FTS_MAX_ACC_STAGES = 3;
static constexpr int FTS_NB_STAGES = 4;
static constexpr int FTS_MAX_GROUPS = 4;
// TMEM budget (512 columns): two query-tile buffers of KB*32 columns + as many 128-column accumulators as fit, at most three.
// KB = 4 (486-bit binary): 256 + 2 x 128; KB <= 2 (fp16 128-d float, 256-bit ORB): 128 + 3 x 128 -- the third stage lets the
// MMA warp run two tiles ahead of the slowest epilogue group.
__host__ __device__ constexpr int fts_acc_stages(int kb, int a_bufs = 2) { return (512 - a_bufs * kb * 32) / 128 >= 3 ? 3 : 2; }
// Epilogue groups (four warps each, one per TMEM lane quarter) take tiles round robin.  Two groups are the default everywhere.
// Four groups (GROUPS = 4: 768 threads x 80 registers, 16-column chunks) were tried for the float path, whose 32-bit keys cost
// 2.5 VIMNMX per column, on the theory that its epilogue warps were latency bound (ncu r01: alu 64 %, issue 65 %, tensor 42 %).
// Measured (cfg4s, same box): 54.5 k pairs/s with two groups, 51.8 k with four -- the epilogue is bound by issue slots and the
// ALU pipe, not by latency: per 128x128 tile and SM sub-partition it executes ~730 instructions (322 VIMNMX/VIMNMX3 at one per
// two clocks = 644 ALU cycles, 128 FFMA, 146 IMAD, 32 LDS, ~100 others; ncu source page of r01) next to ~290 mbarrier-poll
// instructions of the other warps, against 512 cycles of kind::f16 MMAs.  The ALU pipe's 2.5 VIMNMX per column is the floor of an
// exact 32-bit-key top-2 (~740 cycles per tile = 69 % tensor-pipe activity at best); the packed 16-bit keys that halve it for the
// binary engine need distances below 2^10.  The four-group variant stays selectable (SFMM_EPI_GROUPS=4) and tested.
__host__ __device__ constexpr int fts_threads(int groups) { return 32 * (4 + 4 * groups + 4); }

struct FtsSmem {  // after the 1024-byte aligned operand area
    uint64_t a_full[2], a_empty[2];
    uint64_t b_full[FTS_B_STAGES], b_empty[FTS_B_STAGES];
    uint64_t acc_full[FTS_MAX_GROUPS], acc_empty[FTS_MAX_ACC_STAGES];  // full: ring of max(GROUPS, stages) (see the kernel), empty: one per TMEM stage
    uint64_t nb_full[FTS_NB_STAGES], nb_empty[FTS_NB_STAGES];
    uint64_t item_full[2], item_empty[2];
    uint32_t tmem_base;
    uint32_t pad;
    FtItem item[2];
    alignas(16) float nb[FTS_NB_STAGES][FT_N];
    float rowval[2][FT_M];
    uint32_t arow[2][FT_M];  // blob row behind every row of the item's query tile (0xFFFFFFFF: none -> zeros); gathered for TILE_GATHER items
    uint32_t orow[2][FT_M];  // image-relative row the result of that tile row belongs to
    uint4 merge[2][(FTS_MAX_GROUPS - 1) * FT_M];
    uint32_t thr[2][FT_M];  // SKIP: the epilogue groups' common skip threshold per row of the item (a hint: stale values are only looser)
};

static inline size_t float_tensor_ts_smem_bytes(int kblocks) {
    return 1024 /*alignment slack*/ + (size_t)kblocks * FTS_B_STAGES * FT_B_KBLOCK_BYTES + sizeof(FtsSmem);
}

// D[tmem] (+)= A[tmem] * B[smem]^T; whole warp calls, one elected lane issues (see tc_mma).
template <int KIND, bool ACCUMULATE>
__device__ __forceinline__ void tc_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t tmem_sf = 0) {
    if constexpr (KIND == OK_F4)  // e2m1 x e2m1 -> f32, K = 64 per instruction, one UE8M0 scale per 32 elements (all 2^0: tmem_sf.. hold 0x7F bytes)
        asm volatile(
            "{\n\t"
            ".reg .pred pe, p;\n\t"
            "elect.sync _|pe, 0xffffffff;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "@pe tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], [%1], %2, %3, [%5], [%6], p;\n\t"
            "}" ::"r"(tmem_d),
            "r"(tmem_a), "l"(desc_b), "r"(idesc), "n"(ACCUMULATE ? 1 : 0), "r"(tmem_sf), "r"(tmem_sf + 8)
            : "memory");
    else if constexpr (KIND == OK_F16)
        asm volatile(
            "{\n\t"
            ".reg .pred pe, p;\n\t"
            "elect.sync _|pe, 0xffffffff;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
            "}" ::"r"(tmem_d),
            "r"(tmem_a), "l"(desc_b), "r"(idesc), "n"(ACCUMULATE ? 1 : 0)
            : "memory");
    else if constexpr (KIND == OK_I8)
        asm volatile(
            "{\n\t"
            ".reg .pred pe, p;\n\t"
            "elect.sync _|pe, 0xffffffff;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "@pe tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t"
            "}" ::"r"(tmem_d),
            "r"(tmem_a), "l"(desc_b), "r"(idesc), "n"(ACCUMULATE ? 1 : 0)
            : "memory");
    else
        asm volatile(
            "{\n\t"
            ".reg .pred pe, p;\n\t"
            "elect.sync _|pe, 0xffffffff;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
            "}" ::"r"(tmem_d),
            "r"(tmem_a), "l"(desc_b), "r"(idesc), "n"(ACCUMULATE ? 1 : 0)
            : "memory");
}

__device__ __forceinline__ void tc_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
        "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tc_st_32x8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                 "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int KB /* 128-byte K-blocks per row */, int MODE, int GROUPS = 2 /* epilogue groups of four warps: 2 or 4 */,
          bool SKIP = false /* TM_F16X: threshold skipping in the epilogue (chunk_top2_skipx) */>
__global__ void __launch_bounds__(fts_threads(GROUPS), 1)
tensor_knn2_ts_kernel(const __grid_constant__ CUtensorMap tmap, const uint4* __restrict__ a_src /* the matrix the tensor map describes: rows of KB*128 bytes */,
                      const uint32_t total_rows, const float* __restrict__ norms /* int32 popcounts when INT8 */,
                   const float* __restrict__ nb_src /* per train row: |t|^2, or binary_nbkey_kernel's key part when INT8 */,
                      const KnnTile* __restrict__ tiles, const uint32_t n_items_arg, const PairDesc* __restrict__ pairs,
                      KnnEntry* __restrict__ knn, unsigned long long* __restrict__ colmin, const uint32_t key_mul /* = 512 */, const uint32_t i8_bias /* TM_I8P: the descriptors' bit length; rank / collect modes: float bits of the key-table offset C */,
                      uint32_t* __restrict__ cand_count, uint32_t* __restrict__ cand_idx /* TM_TF32_COLLECT */,
                      const uint32_t* __restrict__ n_items_dev /* non-NULL: the item count lives on the device (cross-check reverse pass) */,
                      const uint32_t* __restrict__ xcand /* per pair at col_off: candidate train rows */, const uint32_t* __restrict__ n_xcand) {
    constexpr int KIND = OperandOf<MODE>::kind;
    constexpr int KB_ELEMS = OperandOf<MODE>::kb_elems;
    // SKIP with three K-blocks (128-d TM_F16X): ONE query-tile buffer, so that a third accumulator stage fits (96 + 3 x 128
    // columns).  With two stages and two groups the MMA warp cannot start tile g + 2 before group g % 2 has read the last column of
    // tile g, and that group then waits a whole MMA time for its next tile: ncu (r02, two stages) had the epilogue warps 26 % of
    // their time in that wait while ALU and tensor pipes were both half idle.  The price is the item boundary: the next query tile
    // is stored only after the last MMA of the item has read the old one (the loaders hold it in registers by then).
    // TM_F4P: one buffer as well -- 64 + 3 x 128 columns leave room for the scale-factor words of kind::mxf4 (32 columns of 0x7F
    // bytes = 2^0 in UE8M0 whatever the exact scale layout is), and its 512-cycle tiles need the third stage even more.
    constexpr int A_BUFS = ((SKIP && KB == 3) || KIND == OK_F4) ? 1 : 2;
    constexpr int ACC_STAGES = KIND == OK_F4 ? 3 : fts_acc_stages(KB, A_BUFS);
    static_assert(KIND != OK_F4 || KB <= 2, "TM_F4P: query tile + three accumulators + scale factors must fit 512 TMEM columns");
    const uint32_t n_items = n_items_dev ? __ldg(n_items_dev) : n_items_arg;
    constexpr int FULL_RING = GROUPS > ACC_STAGES ? GROUPS : ACC_STAGES;  // "accumulator ready" barriers, see their initialisation
    constexpr uint32_t A_COLS = KB * 32;            // TMEM columns of one query-tile buffer
    constexpr uint32_t ACC_COL0 = A_BUFS * A_COLS;  // first accumulator column
    [[maybe_unused]] constexpr uint32_t SF_COL0 = ACC_COL0 + ACC_STAGES * FT_N;  // TM_F4P: 32 columns of unit scale factors
    constexpr int EPI_WARPS = 4 * GROUPS;
    constexpr uint32_t LOADER_WARP0 = 4 + EPI_WARPS;  // a multiple of 4: warp % 4 is the TMEM lane quarter it may access
    static_assert(GROUPS >= 2 && GROUPS <= 4, "two to four epilogue groups");
    static_assert(!(GROUPS > 2 && (MODE == TM_I8P || tm_is_collect(MODE) || tm_is_rank(MODE))), "more than two groups: exact top-2 modes only");
    static_assert(!(GROUPS == 4 && MODE == TM_F4P), "the packed fold works on 32-column chunks");
    static_assert(MODE != TM_F16X || KB >= 2, "TM_F16X rows carry at least one data K-block and the key-term K-block");
    static_assert(!SKIP || MODE == TM_F16X || MODE == TM_F4X, "threshold skipping needs accumulators that order like the keys");
    static_assert(MODE != TM_F4X || SKIP, "TM_F4X exists with the skipping epilogue only (TM_F4P is the plain fold)");
    constexpr bool NO_NB = MODE == TM_F16X || MODE == TM_F4X;  // no per-column table: the train-side key term rides in the operand rows
    extern __shared__ unsigned char ft_smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ft_smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* sB = base;  // FTS_B_STAGES x KB x 16 KB
    FtsSmem& sm = *reinterpret_cast<FtsSmem*>(base + (size_t)KB * FTS_B_STAGES * FT_B_KBLOCK_BYTES);

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    [[maybe_unused]] constexpr uint32_t FTS_THR_NEUTRAL = MODE == TM_F4X ? 0xBF800000u /* -1.f */ : 0xFFFFFFF0u;  // "no threshold yet" in sm.thr
    if constexpr (SKIP) {
        for (uint32_t i = threadIdx.x; i < 2 * FT_M; i += blockDim.x) (&sm.thr[0][0])[i] = FTS_THR_NEUTRAL;
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&sm.a_full[s], 4);   // the four loader warps
            mbar_init(&sm.a_empty[s], 1);  // tcgen05.commit after the item's last MMA
            // every thread that writes / reads an item slot arrives itself (release / acquire pair per thread)
            mbar_init(&sm.item_full[s], 32);                             // the prefetch warp's lanes
            mbar_init(&sm.item_empty[s], 1 + 32 + 32 * EPI_WARPS + 128);  // producer thread, MMA warp, epilogue warps, loader warps
        }
        for (int s = 0; s < FTS_B_STAGES; ++s) {
            mbar_init(&sm.b_full[s], 1);
            mbar_init(&sm.b_empty[s], 1);
        }
        // "accumulator of tile g is ready" is signalled on barrier g % FULL_RING, FULL_RING = max(GROUPS, ACC_STAGES).  try_wait.parity
        // cannot tell "phase k done" from "phase k-2 done", so a barrier must never be two phases away from its waiter, either way:
        //  - the waiter (group g % GROUPS) has seen its previous tile g - GROUPS complete, and tiles complete in order, so tile
        //    g - FULL_RING (the barrier's previous phase) is complete because FULL_RING >= GROUPS;
        //  - the barrier's next phase (tile g + FULL_RING) cannot be signalled before tile g has been consumed, because the MMA warp
        //    stops at tile g + ACC_STAGES <= g + FULL_RING until tile g's TMEM stage is released.
        // (One barrier per TMEM stage is wrong with four groups on three stages: a fast group waits for a stage's NEXT use before its
        // current one has completed and reads the wrong tile.  One barrier per group is wrong with two groups on three stages: tile
        // g + 2 is signalled before tile g has been consumed.  Both were measured the hard way.)
        for (int s = 0; s < FULL_RING; ++s) mbar_init(&sm.acc_full[s], 1);
        for (int s = 0; s < ACC_STAGES; ++s) mbar_init(&sm.acc_empty[s], 4);  // the four warps of the group that took the tile
        for (int s = 0; s < FTS_NB_STAGES; ++s) {
            mbar_init(&sm.nb_full[s], 1);
            mbar_init(&sm.nb_empty[s], 4);
        }
        mbar_fence_init();
    }
    if (warp == 2) {  // whole warp: allocate all 512 TMEM columns (1 CTA per SM: smem-limited)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&sm.tmem_base))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sm.tmem_base;
    if constexpr (KIND == OK_F4) {  // unit scale factors, written once by the four loader warps (one per TMEM lane quarter)
        if (warp >= 4 + 4 * GROUPS) {
            uint32_t ones[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) ones[i] = 0x7F7F7F7Fu;
            tc_st_32x32(tmem_base + ((((warp - 4 - 4 * GROUPS) * 32) << 16)) + SF_COL0, ones);
            tc_wait_st();
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }

    if (warp == 0) {
        // ===================== TMA producer: train tiles only =====================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
            uint32_t g = 0, it = 0;  // tiles / items this CTA has gone through: ring slots and mbarrier phases
            for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const uint32_t slot = it & 1;
                mbar_wait(&sm.item_full[slot], (it >> 1) & 1);
                const uint32_t b_row0 = sm.item[slot].b_row0, n_tiles = sm.item[slot].n_tiles;
                mbar_arrive(&sm.item_empty[slot]);
#pragma unroll 1
                for (uint32_t j = 0; j < n_tiles; ++j, ++g) {
                    const uint32_t s = g % FTS_B_STAGES;
                    if constexpr (!NO_NB) {
                        const uint32_t a = g % FTS_NB_STAGES;
                        mbar_wait(&sm.nb_empty[a], ((g / FTS_NB_STAGES) & 1) ^ 1);
                        mbar_expect_tx(&sm.nb_full[a], FT_N * sizeof(float));
                        // image rows start at multiples of 4 and t0 at multiples of 128: 16-byte aligned source;
                        // the norms array is padded so that the copy may run past the image's last row
                        tma_load_1d(sm.nb[a], nb_src + b_row0 + j * FT_N, FT_N * sizeof(float), &sm.nb_full[a]);
                    }
                    mbar_wait(&sm.b_empty[s], ((g / FTS_B_STAGES) & 1) ^ 1);
                    mbar_expect_tx(&sm.b_full[s], KB * FT_B_KBLOCK_BYTES);
                    unsigned char* dst = sB + (size_t)s * KB * FT_B_KBLOCK_BYTES;
                    const int row = (int)(b_row0 + j * FT_N);
#pragma unroll
                    for (int kb = 0; kb < KB; ++kb)
#pragma unroll
                        for (int h = 0; h < FT_N / FT_BOX_ROWS; ++h)
                            tma_load_2d(dst + kb * FT_B_KBLOCK_BYTES + h * FT_BOX_BYTES, &tmap, kb * KB_ELEMS,
                                        row + h * FT_BOX_ROWS, &sm.b_full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp, uniform; one elected lane issues) =====================
        const uint32_t tb = __shfl_sync(0xFFFFFFFFu, tmem_base, 0);
        const uint32_t idesc = KIND == OK_F4 ? FT_IDESC_MXF4 : (KIND == OK_I8 ? FT_IDESC_I8 : (KIND == OK_F16 ? FT_IDESC_F16 : FT_IDESC));
        [[maybe_unused]] const uint32_t sf_tmem = tb + SF_COL0;
        uint32_t g = 0, it = 0;
        for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const uint32_t slot = it & 1;
            mbar_wait(&sm.item_full[slot], (it >> 1) & 1);
            const uint32_t n_tiles = __shfl_sync(0xFFFFFFFFu, sm.item[slot].n_tiles, 0);
            mbar_arrive(&sm.item_empty[slot]);  // every thread, after its own reads of the slot
            const uint32_t a_idx = A_BUFS == 2 ? slot : 0u, a_phase = A_BUFS == 2 ? (it >> 1) & 1 : it & 1;
            mbar_wait(&sm.a_full[a_idx], a_phase);  // the loaders have stored this item's query tile
            tc_fence_after();
            const uint32_t a_tmem = tb + a_idx * A_COLS;
#pragma unroll 1
            for (uint32_t j = 0; j < n_tiles; ++j, ++g) {
                const uint32_t s = g % FTS_B_STAGES, a = g % ACC_STAGES;
                mbar_wait(&sm.b_full[s], (g / FTS_B_STAGES) & 1);
                mbar_wait(&sm.acc_empty[a], ((g / ACC_STAGES) & 1) ^ 1);
                tc_fence_after();
                const uint64_t b_desc0 = umma_desc_sw128(smem_u32(sB + (size_t)s * KB * FT_B_KBLOCK_BYTES));
                const uint32_t d_tmem = tb + ACC_COL0 + a * FT_N;
                tc_mma_ts<KIND, false>(d_tmem, a_tmem, b_desc0, idesc, sf_tmem);
                // TM_F16X: only the first 32 bytes of the last K-block carry data (the three key-term columns), the rest is zero
                constexpr int N_MMA = MODE == TM_F16X ? 4 * (KB - 1) + 1 : 4 * KB;
#pragma unroll
                for (int i = 1; i < N_MMA; ++i) {  // i = kb*4 + k: 32 bytes of K = 8 TMEM columns of A, 32 bytes inside B's swizzle row
                    const int kb = i >> 2, k = i & 3;
                    tc_mma_ts<KIND, true>(d_tmem, a_tmem + i * 8, b_desc0 + ((kb * FT_B_KBLOCK_BYTES + k * 32) >> 4), idesc, sf_tmem);
                }
                tc_commit_elect(&sm.b_empty[s]);   // smem stage reusable once these MMAs have read it
                tc_commit_elect(&sm.acc_full[g % FULL_RING]);  // accumulator ready for the group that takes this tile
            }
            tc_commit_elect(&sm.a_empty[a_idx]);  // every MMA of this item has read the query tile: its TMEM buffer may be refilled
        }
    } else if (warp == 3) {
        // ===================== item prefetch =====================
        uint32_t it = 0;
        for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const uint32_t slot = it & 1;
            mbar_wait_relaxed(&sm.item_empty[slot], ((it >> 1) & 1) ^ 1);
            KnnTile tile = tiles[item];
            PairDesc pd = pairs[tile.pair];
            // "reverse" tile: the symmetric cross-check's column minima, roles of the two images swapped; with TILE_GATHER its
            // rows are the pair's candidate train rows (filter.cuh), addressed through the candidate list
            const bool reverse = (tile.split & TILE_REVERSE) != 0, gather = (tile.split & TILE_GATHER) != 0;
            tile.split &= TILE_SPLIT_MASK;
            if (reverse) {
                const uint32_t r0 = pd.q_row0, n = pd.nq;
                pd.q_row0 = pd.t_row0; pd.nq = pd.nt;
                pd.t_row0 = r0; pd.nt = n;
            }
            if (gather) pd.nq = __ldg(n_xcand + tile.pair);  // rows of this side = entries of the candidate list
            if (lane == 0) {
                FtItem& o = sm.item[slot];
                o.a_row = pd.q_row0 + tile.q0;
                o.b_row0 = pd.t_row0 + tile.t0;
                o.n_rows = tile.t1 - tile.t0;
                o.n_tiles = (tile.t1 - tile.t0 + FT_N - 1) / FT_N;
                o.q0 = tile.q0; o.nq = pd.nq; o.t0 = tile.t0; o.split = tile.split; o.reverse = reverse ? 1u : 0u;
                o.n_splits = pd.n_splits; o.q_off = pd.q_off; o.knn_off = pd.knn_off; o.col_off = pd.col_off;
            }
#pragma unroll
            for (int rr = 0; rr < FT_M / 32; ++rr) {
                const uint32_t row = rr * 32 + lane, qrow = tile.q0 + row;
                const bool valid = qrow < pd.nq;
                // image-relative row behind this tile row; rows past the image (but inside the blob) are another image's:
                // computed on, never read -- as with TMA's box
                const uint32_t irow = gather ? (valid ? __ldg(xcand + pd.col_off + qrow) : 0xFFFFFFFFu) : qrow;
                const uint32_t grow = (gather && !valid) ? 0xFFFFFFFFu : pd.q_row0 + irow;
                sm.arow[slot][row] = grow;
                sm.orow[slot][row] = irow;
                const float nq2 = valid ? __ldg(norms + grow) : 0.f;
                float v;
                if constexpr (tm_is_collect(MODE)) {
                    v = collect_threshold<MODE == TM_F16_COLLECT>(valid, nq2, pd, knn, qrow, __uint_as_float(i8_bias));
                } else {
                    v = nq2;  // popc(q) as integer bits (i8) / |q|^2 (exact float modes, rank)
                }
                sm.rowval[slot][row] = v;
            }
            mbar_arrive(&sm.item_full[slot]);  // every lane, after its own writes
        }
    } else if (warp >= LOADER_WARP0) {
        // ===================== query loaders: global -> registers -> TMEM, one item ahead =====================
        const uint32_t lw = warp - LOADER_WARP0;  // == warp % 4: the TMEM lane quarter this warp may access
        const uint32_t row = lw * 32 + lane;
        uint32_t it = 0;
        for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const uint32_t slot = it & 1;
            mbar_wait_relaxed(&sm.item_full[slot], (it >> 1) & 1);
            const uint32_t grow = sm.arow[slot][row];
            mbar_arrive(&sm.item_empty[slot]);  // every thread, after its own reads of the slot
            const uint32_t a_idx = A_BUFS == 2 ? slot : 0u, a_phase = A_BUFS == 2 ? (it >> 1) & 1 : it & 1;
            const bool ok = grow < total_rows;  // rows past the blob: zeros (rows past the image but inside the blob are
                                                // another image's: computed on, never read -- as with TMA's box)
            const uint4* src = a_src + (size_t)(ok ? grow : 0) * (KB * 8);
            const uint32_t taddr = tmem_base + ((lw * 32) << 16) + a_idx * A_COLS;
            if constexpr (A_BUFS == 2) {
                mbar_wait_relaxed(&sm.a_empty[a_idx], a_phase ^ 1);  // the MMAs of the item before last are done with this buffer
                tc_fence_after();
#pragma unroll
                for (int kb = 0; kb < KB; ++kb) {
                    uint32_t r[32];
#pragma unroll
                    for (int v = 0; v < 8; ++v) {
                        const uint4 x = ok ? __ldg(src + kb * 8 + v) : make_uint4(0, 0, 0, 0);
                        r[4 * v + 0] = x.x; r[4 * v + 1] = x.y; r[4 * v + 2] = x.z; r[4 * v + 3] = x.w;
                    }
                    tc_st_32x32(taddr + kb * 32, r);
                }
            } else {
                // one buffer: the row waits in registers until the previous item's last MMA has read the old tile
                // (TM_F16X: only the first 32 bytes of the last K-block carry data, the rest of it is never read)
                constexpr int FULL_KB = MODE == TM_F16X ? KB - 1 : KB;
                uint32_t r[FULL_KB][32];
                [[maybe_unused]] uint32_t rx[8];
#pragma unroll
                for (int kb = 0; kb < FULL_KB; ++kb)
#pragma unroll
                    for (int v = 0; v < 8; ++v) {
                        const uint4 x = ok ? __ldg(src + kb * 8 + v) : make_uint4(0, 0, 0, 0);
                        r[kb][4 * v + 0] = x.x; r[kb][4 * v + 1] = x.y; r[kb][4 * v + 2] = x.z; r[kb][4 * v + 3] = x.w;
                    }
                if constexpr (MODE == TM_F16X) {
#pragma unroll
                    for (int v = 0; v < 2; ++v) {
                        const uint4 x = ok ? __ldg(src + (KB - 1) * 8 + v) : make_uint4(0, 0, 0, 0);
                        rx[4 * v + 0] = x.x; rx[4 * v + 1] = x.y; rx[4 * v + 2] = x.z; rx[4 * v + 3] = x.w;
                    }
                }
                mbar_wait_relaxed(&sm.a_empty[a_idx], a_phase ^ 1);
                tc_fence_after();
#pragma unroll
                for (int kb = 0; kb < FULL_KB; ++kb) tc_st_32x32(taddr + kb * 32, r[kb]);
                if constexpr (MODE == TM_F16X) tc_st_32x8(taddr + (KB - 1) * 32, rx);
            }
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.a_full[a_idx]);
        }
    } else if (warp >= 4) {
        // ===================== epilogue: fused top-2 =====================
        // GROUPS groups of four warps (one warp per TMEM lane quarter); group h takes the tiles whose running counter is
        // h mod GROUPS.  (Splitting every tile's columns between the groups instead, so that a stage is released as soon as
        // it has been read, measured 10 % slower: the TMEM loads no longer overlap the fold inside a warp.)
        const uint32_t ew = warp - 4;
        const uint32_t quarter = ew & 3, half = ew >> 2;   // TMEM lanes 32*quarter.. ; `half` = the group index
        const uint32_t row = quarter * 32 + lane;          // row of the query tile == TMEM lane
        constexpr int CW = GROUPS == 4 ? 16 : 32;          // columns per chunk streamed from TMEM (register budget: 80 per thread with four groups)
        uint32_t g0 = 0, it = 0;
        for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const uint32_t slot = it & 1;
            mbar_wait(&sm.item_full[slot], (it >> 1) & 1);
            const FtItem& im = sm.item[slot];
            const uint32_t n_rows = im.n_rows, n_tiles = im.n_tiles, nq = im.nq, t0 = im.t0, split = im.split, n_splits = im.n_splits;
            const uint32_t qrow = im.q0 + row, q_off = im.q_off;
            const bool reverse = im.reverse != 0;
            const unsigned long long knn_off = im.knn_off, col_off = im.col_off;
            const float cq = sm.rowval[slot][row];  // TM_TF32_COLLECT: the threshold tau
            const uint32_t out_row = sm.orow[slot][row];  // == qrow unless the item's rows are gathered
            mbar_arrive(&sm.item_empty[slot]);  // every thread, after its own reads of the slot

            Top2 best;
            best.init();
            [[maybe_unused]] uint32_t* cand_count_row = nullptr;
            [[maybe_unused]] uint32_t* cand_idx_row = nullptr;
            [[maybe_unused]] uint32_t fill = 0;
            [[maybe_unused]] int r1 = 0x7FFFFFFF, r2 = 0x7FFFFFFF;  // TM_TF32_RANK: two smallest approximate d^2 (float bits), over all tiles
            if constexpr (tm_is_collect(MODE)) {
                // each epilogue group owns one counter and half of the row's list
                cand_count_row = cand_count + 2 * (size_t)(q_off + min(qrow, nq - 1)) + half;
                cand_idx_row = cand_idx + (size_t)(q_off + min(qrow, nq - 1)) * FT_CAND_CAP + half * (FT_CAND_CAP / 2);
            }
            // SKIP: only accumulators below thrv can still enter the row's top-2 (chunk_top2_skipx); rows whose result is never
            // written never ask for the slow path.  The groups of a row share their thresholds through sm.thr: every group sees
            // 1/GROUPS of the columns, together they see all of them, and the number of slow-path visits follows the columns
            // seen.  The other groups' tiles may lie AFTER this group's next tile, where an equal distance must not win, hence
            // their threshold + 1.  The word is a hint (a stale value is only looser), exchanged with ONE shared-memory atomic per
            // tile (min / max returns what the others had).  Items alternate between two slots; group 0 resets the next item's slot
            // before the named barrier that ends every item (finish_rows), which nobody passes before all groups have finished the
            // item -- so a slot never carries a previous item's value and plain stores never meet the atomics (racecheck clean).
            [[maybe_unused]] uint32_t thrv = qrow < nq ? 0xFFFFFFF0u : 0u;
            [[maybe_unused]] uint32_t* thr_shared = &sm.thr[it & 1][row];
            // TM_F4X: larger accumulators are nearer, the threshold is the accumulator VALUE of the second-nearest column so far
            // (-1: none yet; +inf: row never written); the shared word holds float bits, the other groups' threshold counts - 1
            [[maybe_unused]] float thrf = qrow < nq ? -1.f : __int_as_float(0x7f800000);
#pragma unroll 1
            for (uint32_t j = (half + GROUPS - g0 % GROUPS) % GROUPS; j < n_tiles; j += GROUPS) {
                const uint32_t g = g0 + j;               // g % GROUPS == half
                const uint32_t a = g % ACC_STAGES;
                const uint32_t nbs = g % FTS_NB_STAGES;
                mbar_wait(&sm.acc_full[g % FULL_RING], (g / FULL_RING) & 1);
                tc_fence_after();
                uint32_t acc[2][CW];  // register double buffer: chunk c+1 streams in from TMEM while chunk c is folded
                const uint32_t taddr = tmem_base + ((quarter * 32) << 16) + ACC_COL0 + a * FT_N;
                tc_ld_32x32(taddr, acc[0]);
                if constexpr (!NO_NB) mbar_wait(&sm.nb_full[nbs], (g / FTS_NB_STAGES) & 1);
                tc_wait_ld(acc[0]);
                const uint32_t col0 = j * FT_N;                // first column of the tile, relative to t0
                const bool partial = col0 + FT_N > n_rows;     // warp-uniform: only the last tile
                uint32_t m1 = 0xFFFFFFFFu, m2 = 0xFFFFFFFFu;   // two smallest keys of this tile
                constexpr int NCH = FT_N / CW;  // chunks per tile
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    if (c < NCH - 1) tc_ld_32x32(taddr + (c + 1) * CW, acc[(c + 1) & 1]);
                    const uint32_t nb_saddr = smem_u32(&sm.nb[nbs][c * CW]);
                    if constexpr (tm_is_collect(MODE)) {
                        const uint32_t c0 = col0 + c * 32;  // columns at or past n_rows belong to another image / padding
                        const uint32_t valid = c0 + 32 <= n_rows ? 0xFFFFFFFFu : (c0 < n_rows ? (1u << (n_rows - c0)) - 1u : 0u);
                        chunk_collect(acc[c & 1], nb_saddr, cq, valid, t0 + c0, n_splits != 1, fill, cand_count_row, cand_idx_row);
                    } else if constexpr (tm_is_rank(MODE)) {
                        if (!partial) chunk_rank<false>(acc[c & 1], nb_saddr, col0 + c * 32, n_rows, r1, r2);
                        else chunk_rank<true>(acc[c & 1], nb_saddr, col0 + c * 32, n_rows, r1, r2);
                    } else if constexpr (MODE == TM_I8P || MODE == TM_F4P) {
                        // (key_mul - 640 = -128 from the kernel parameter: stays an IMAD on the FMA pipe)
                        if (!partial) chunk_top2_packed<false, MODE == TM_F4P>(acc[c & 1], nb_saddr, key_mul - 640u, (key_mul - 640u) << 16, col0 + c * 32, n_rows, m1, m2);
                        else chunk_top2_packed<true, MODE == TM_F4P>(acc[c & 1], nb_saddr, key_mul - 640u, (key_mul - 640u) << 16, col0 + c * 32, n_rows, m1, m2);
                    } else if constexpr (SKIP && MODE == TM_F4X) {
                        if (!partial) chunk_top2_skipx_f4<CW>(acc[c & 1], key_mul, c * CW, thrf, m1, m2);
                        else chunk_top2<true, MODE, CW>(acc[c & 1], nb_saddr, cq, key_mul, c * CW, col0 + c * CW, n_rows, m1, m2);
                    } else if constexpr (SKIP) {
                        if (!partial) chunk_top2_skipx<CW>(acc[c & 1], key_mul, c * CW, thrv, m1, m2);
                        else chunk_top2<true, MODE, CW>(acc[c & 1], nb_saddr, cq, key_mul, c * CW, col0 + c * CW, n_rows, m1, m2);
                    } else {
                        if (!partial) chunk_top2<false, MODE, CW>(acc[c & 1], nb_saddr, cq, key_mul, c * CW, col0 + c * CW, n_rows, m1, m2);
                        else chunk_top2<true, MODE, CW>(acc[c & 1], nb_saddr, cq, key_mul, c * CW, col0 + c * CW, n_rows, m1, m2);
                    }
                    if (c < NCH - 1) tc_wait_ld(acc[(c + 1) & 1]);
                    if (c == (NCH > 1 ? NCH - 2 : 0)) {  // the last TMEM read of this tile has landed: the MMA that reuses the stage may start
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&sm.acc_empty[a]);
                    }
                }
                if constexpr (!NO_NB) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sm.nb_empty[nbs]);
                }
                // merge the tile's two best into the running pair (ascending tiles = arrival order)
                if constexpr (!tm_is_rank(MODE) && !tm_is_collect(MODE)) {
                    const int tbase = (int)(t0 + col0);
                    if constexpr (MODE == TM_I8P || MODE == TM_F4P) {  // four 16-bit lane winners -> the tile's two smallest (hamming, column)
                        uint32_t q1, q2;
                        unpack_top2_u16x2(m1, m2, __float_as_uint(cq) - i8_bias, q1, q2);
                        if (q1 != 0xFFFFFFFFu) best.offer(q1 >> 9, tbase + (int)(q1 & 511u));
                        if (q2 != 0xFFFFFFFFu) best.offer(q2 >> 9, tbase + (int)(q2 & 511u));
                    } else {
                        // i8: the key carries popc(t) - 2 q.t + I8_BIAS; popc(q) (the bits of cq) completes the Hamming distance.
                        // exact float modes: the key carries d^2 - |q|^2 + 2^20 (float_nbexact_kernel); |q|^2 is an integer <= 2^20
                        const uint32_t dadd = (MODE == TM_I8 || MODE == TM_F4X) ? __float_as_uint(cq) - I8_BIAS : static_cast<uint32_t>(cq) - 1048576u;
                        if (m1 != 0xFFFFFFFFu) best.offer((m1 >> 9) + dadd, tbase + (int)(m1 & 511u));
                        if (m2 != 0xFFFFFFFFu) best.offer((m2 >> 9) + dadd, tbase + (int)(m2 & 511u));
                        if constexpr (SKIP && MODE == TM_F4X) {  // the same in accumulator values: 1024 - key value
                            if (best.i2 >= 0) thrf = fmaxf(thrf, F4X_KMAGIC - __uint_as_float(0x4B000000u | (best.d2 - dadd)));
                            // (-1, +inf and the positive accumulator values order as signed integers like as floats)
                            const float other = __int_as_float(atomicMax(reinterpret_cast<int*>(thr_shared), __float_as_int(thrf)));
                            thrf = fmaxf(thrf, other - 1.f);
                        } else if constexpr (SKIP) {  // thresholds of the next tiles: own merged list, the other groups' + 1
                            if (best.i2 >= 0) thrv = min(thrv, 0x4B000000u + (best.d2 - dadd));
                            const uint32_t other = atomicMin(thr_shared, thrv);
                            thrv = min(thrv, other + 1u);
                        }
                    }
                }
            }
            g0 += n_tiles;
            if constexpr (tm_is_rank(MODE)) {  // values only, clamped at 0; the index field is unused
                // x' = d~^2 - |q|^2 + C (C = the set's max |x|^2, passed in i8_bias as float bits): back to d~^2, clamped at 0
                const float back = cq - __uint_as_float(i8_bias);
                if (r1 != 0x7FFFFFFF) { best.d1 = __float_as_uint(fmaxf(__int_as_float(max(r1, 0)) + back, 0.f)); best.i1 = 0; }
                if (r2 != 0x7FFFFFFF) { best.d2 = __float_as_uint(fmaxf(__int_as_float(max(r2, 0)) + back, 0.f)); best.i2 = 0; }
            }
            if constexpr (tm_is_collect(MODE)) {
                if (n_splits == 1 && qrow < nq) *cand_count_row = fill;
            } else {
                if constexpr (SKIP) {
                    if (half == 0) sm.thr[(it + 1) & 1][row] = FTS_THR_NEUTRAL;  // next item's slot, see above
                }
                finish_rows<MODE, GROUPS>(sm.merge[it & 1], best, half, row, qrow, nq, reverse, split, knn_off, col_off, knn, colmin, out_row);
            }
        }
    }

    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

}  // namespace sfmm
