// Tensor-core 2-NN kernels for sm_100a (tcgen05 + TMEM + TMA): shared helpers, the prepare kernels, the epilogue
// pieces, and the kernel that keeps BOTH operands in shared memory (tensor_knn2_kernel, "SS").  The default for
// the binary engine and the fp16 float path is the TMEM-A kernel of float_tensor_ts.cuh, which shares everything here.
//
// Replaces matcher->knnMatch(q, t, knn, 2) with cv::BFMatcher at /root/reference/src/Sfm.cpp:593,599.
//
//   float : d^2(q,t)     = |q|^2 + |t|^2 - 2 q.t      tcgen05.mma kind::f16 on an exact fp16 copy / kind::tf32
//   binary: hamming(q,t) = popc(q) + popc(t) - 2 q.t   tcgen05.mma kind::i8 on bits unpacked to {0,1} bytes (u8 x u8 -> s32, exact)
//
// The contraction runs on the tensor cores (operands staged by TMA with the 128-byte swizzle, accumulators in
// TMEM); the rest is a fused epilogue that never writes the distance matrix: per accumulator element one FFMA /
// IMAD forms a packed integer key (distance << 9 | column), a branch-free network keeps the two smallest keys of
// the tile, and once per tile they are merged into the row's running (distance, train index) pair with a strict
// '<' in arrival order -- cv::batchDistance's insertion rule.
//
// Exactness contract (float, TM_TF32_EXACT / TM_F16_EXACT).  Selected only for descriptor sets the prepare kernel
// proved exact: every value an integer with |v| <= 2047 (exactly representable in TF32 and fp16) and every row
// norm^2 <= 2^20, so every product, every partial sum of the contraction and d^2 are integers below 2^22 -- exact
// in fp32 whatever the summation order, and small enough for sqrtf to keep distinct d^2 distinct.  Real SIFT
// output is of this kind (OpenCV quantises to 0..255, norm ~512; verified on data/temple).  Then d = sqrtf(d^2)
// is bit-identical to OpenCV's sqrtf(sum (a-b)^2) and indices, distances and ties are bit-exact.  Arbitrary floats
// take the TM_TF32_RANK -> TM_TF32_COLLECT -> float_refine_kernel path below (bit-identical to float_exact.cuh).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include "binary_knn.cuh"  // mbarrier / PTX helpers
#include "common.cuh"
#include "float_exact.cuh"  // top2_insert

namespace sfmm {

// What the kernel computes (template parameter MODE):
enum TensorMode {
    TM_TF32_EXACT = 0,    // TF32-exact float data: keys carry the exact integer d^2, result is final
    TM_I8 = 1,            // binary descriptors unpacked to bytes (kind::i8), result is final
    TM_TF32_RANK = 2,     // arbitrary floats, pass 1: approximate d^2 (TF32 truncation) -> approximate top-2 per row
    TM_I8P = 4,           // TM_I8 with two 16-bit keys per register (descriptors < 512 bit): half the min/max work
    TM_F16_RANK = 6,      // TM_TF32_RANK / TM_TF32_COLLECT on an fp16 (round-to-nearest) copy: same 10-bit significand as TF32,
    TM_F16_COLLECT = 7,   // half the MMA time and far less power (the TF32 passes run into the power cap on dense mantissas)
    TM_F16_EXACT = 5,     // TM_TF32_EXACT on an fp16 copy of the descriptors (integers |v| <= 2048 are exact in fp16):
                          // kind::f16 contracts 16 elements per MMA, half the tensor time and half the operand bytes
    TM_F16X = 8,          // TM_F16_EXACT with the key's train-side term contracted by the tensor core as well: the operand rows carry one
                          // extra 128-byte K-block -- query side (-2q, 1, 2048, 2048, 0...), train side (t, v0, v1, 2048 v2, 0...) with
                          // v0 + 2048 v1 + 2048^2 v2 = |t|^2 + 2^23 + 2^20 -- so that the accumulator IS the key argument
                          // x' = |t|^2 - 2 q.t + 2^23 + 2^20 (exact: every term an integer, every partial sum below 2^24): no FFMA, no
                          // norm table read per column in the epilogue (float_to_half_kx_kernel; one MMA more per tile)
    TM_F4P = 9,           // TM_I8P on the FP4 pipe: a bit is the E2M1 value 1.0 or 0.0 (one nibble), every block scale is 2^0, and
                          // tcgen05.mma kind::mxf4.block_scale contracts 64 bits per instruction in the 64 cycles kind::i8 needs for 32 --
                          // twice the rate on HALF the operand bytes (256 B per 512-bit row); products are 0 or 1 and the fp32 accumulator
                          // holds integers <= 512, so q.t is exact (tools/mxf4_bench.cu checks it against popc(q & t)); TMEM-A kernel only
    TM_F4X = 10,          // TM_F4P with the key's train-side term contracted by the tensor core as well (what TM_F16X is to TM_F16_EXACT), for
                          // descriptors that leave 17 spare elements in their last K-block (AKAZE: 488 of 512).  Query bits are 1.0, train
                          // bits 2.0, and the spare elements carry 6,...,6,1 on the query side and a digit expansion of 512 - popc(t) on the
                          // train side (binary_unpack4x_kernel), so the accumulator is  x = 2 q.t - popc(t) + 512  -- an exact positive
                          // integer that orders the columns of a row like the Hamming distance, descending.  The epilogue compares raw
                          // accumulators (no table, no IMAD per column) and forms keys only where a column can still enter the row's top-2
    TM_TF32_COLLECT = 3   // arbitrary floats, pass 2: every column whose approximate d^2 can still be in the exact
                          // top-2 (<= m2 + 2*eps, a rigorous bound) is appended to the row's candidate list, which
                          // float_refine_kernel then evaluates exactly (fp32 direct difference, float_exact.cuh's arithmetic)
};
static constexpr int FT_CAND_CAP = 32;  // candidate slots per query row (two halves of 16); a half overflowing => exact rescan of the row

static constexpr int FT_M = 128;         // query rows per CTA (UMMA M)
static constexpr int FT_N = 128;         // train rows per MMA tile (UMMA N); 64 was measured slower (per-tile costs double)
static constexpr int FT_KB_ELEMS = 32;   // floats per 128-byte swizzle row
static constexpr int FT_B_STAGES = 2;    // train tiles in shared memory (2 x 64 KB next to the 64 KB query tile)
static constexpr int FT_ACC_STAGES = 4;  // 4 x 128 fp32 columns = 512 TMEM columns
static constexpr int FT_THREADS = 384;   // 4 control warps + 8 epilogue warps
static constexpr int FT_EPI_WARPS = 8;
static constexpr uint32_t FT_KBLOCK_BYTES = FT_M * 128;  // one K-block of the 128-row query tile: 16 KB
static constexpr uint32_t FT_B_KBLOCK_BYTES = FT_N * 128;  // one K-block of a train tile
static constexpr int FT_BOX_ROWS = 64;                   // TMA box: 128 bytes x 64 rows (half a K-block)
static constexpr uint32_t FT_BOX_BYTES = FT_BOX_ROWS * 128;

// ------------------------------------------------------------------ tcgen05 / TMA PTX
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__host__ __device__ constexpr bool tm_is_rank(int mode) { return mode == TM_TF32_RANK || mode == TM_F16_RANK; }
__host__ __device__ constexpr bool tm_is_collect(int mode) { return mode == TM_TF32_COLLECT || mode == TM_F16_COLLECT; }
// Operand type of a mode: what the tensor map describes and which tcgen05.mma kind contracts it.
enum OperandKind { OK_TF32 = 0, OK_I8 = 1, OK_F16 = 2, OK_F4 = 3 };
template <int MODE>
struct OperandOf {
    static constexpr int kind = (MODE == TM_F4P || MODE == TM_F4X) ? OK_F4 : (MODE == TM_I8 || MODE == TM_I8P) ? OK_I8 : ((MODE == TM_F16_EXACT || MODE == TM_F16X || MODE == TM_F16_RANK || MODE == TM_F16_COLLECT) ? OK_F16 : OK_TF32);
    static constexpr int kb_elems = (kind == OK_I8 || kind == OK_F4) ? 128 : (kind == OK_F16 ? 64 : 32);  // tensor-map elements per 128-byte swizzle row (nibble pairs travel as bytes)
};
// D[tmem] (+)= A[smem] * B[smem]^T; M=128, N=128, 32 bytes of K per instruction (8 tf32 / 16 f16 / 32 u8), fp32 or s32
// accumulate.  Called by the WHOLE warp with warp-uniform operands; one elected lane issues.  (Issuing from inside
// an `if (lane == 0)` made ptxas wrap every MMA in a R2UR.BROADCAST / BRA.U.ANY uniformisation loop
// and recompute the descriptors: ~90 cycles per instruction for 64 cycles of tensor work.)
template <int KIND, bool ACCUMULATE>
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc) {
    if constexpr (KIND == OK_F16)  // f16 x f16 -> f32, K = 16 per instruction
        asm volatile(
            "{\n\t"
            ".reg .pred pe, p;\n\t"
            "elect.sync _|pe, 0xffffffff;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
            "}" ::"r"(tmem_d),
            "l"(desc_a), "l"(desc_b), "r"(idesc), "n"(ACCUMULATE ? 1 : 0)
            : "memory");
    else if constexpr (KIND == OK_I8)  // u8 x u8 -> s32, K = 32 per instruction
        asm volatile(
            "{\n\t"
            ".reg .pred pe, p;\n\t"
            "elect.sync _|pe, 0xffffffff;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "@pe tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
            "}" ::"r"(tmem_d),
            "l"(desc_a), "l"(desc_b), "r"(idesc), "n"(ACCUMULATE ? 1 : 0)
            : "memory");
    else  // tf32 x tf32 -> f32, K = 8 per instruction
        asm volatile(
            "{\n\t"
            ".reg .pred pe, p;\n\t"
            "elect.sync _|pe, 0xffffffff;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
            "}" ::"r"(tmem_d),
            "l"(desc_a), "l"(desc_b), "r"(idesc), "n"(ACCUMULATE ? 1 : 0)
            : "memory");
}
__device__ __forceinline__ void tc_commit_elect(uint64_t* bar) {
    asm volatile(
        "{\n\t"
        ".reg .pred pe;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
        "}" ::"r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tc_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_ld_32x32(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                   "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}
// Waits for this thread's outstanding tcgen05.ld; the registers are listed as in/out operands so that
// the compiler cannot schedule a use of them above the wait (the loads complete asynchronously).
__device__ __forceinline__ void tc_wait_ld(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                   "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]),
                   "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]),
                   "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}

// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle (cute::UMMA::SmemDescriptor):
// start address >> 4 in bits [0,14), LBO bits [16,30) (unused for swizzled K-major), SBO >> 4 in
// bits [32,46) = 1024 B (8 rows x 128 B), version 1 in bits [46,48), layout SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 (bits 4-5 = 1), A=B=TF32 (bits 7-9,
// 10-12 = 2), both K-major, N>>3 in bits [17,23), M>>4 in bits [24,29).
static constexpr uint32_t FT_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((FT_N >> 3) << 17) | ((FT_M >> 4) << 24);
// kind::i8: D=S32 (bits 4-5 = 2), A=B=UINT8 (format 0), both K-major.
static constexpr uint32_t FT_IDESC_I8 = (2u << 4) | (0u << 7) | (0u << 10) | ((FT_N >> 3) << 17) | ((FT_M >> 4) << 24);
// kind::f16: D=F32 (1), A=B=F16 (format 0), both K-major.
// kind::mxf4.block_scale (CUTLASS mma_sm100_desc.hpp, InstrDescriptorBlockScaled): a/b format E2M1 = 1 at bits 7 / 10, both K-major,
// N >> 3 at 17, scale format UE8M0 = 1 at 23, M >> 4 at 24, scale-factor ids 0, K = 64 per instruction
static constexpr uint32_t FT_IDESC_MXF4 = (1u << 7) | (1u << 10) | ((FT_N >> 3) << 17) | (1u << 23) | ((FT_M >> 4) << 24);
static constexpr uint32_t FT_IDESC_F16 = (1u << 4) | (0u << 7) | (0u << 10) | ((FT_N >> 3) << 17) | ((FT_M >> 4) << 24);

__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// ------------------------------------------------------------------ prepare: norms + TF32-exactness proof
// One warp per row.  flags[0] |= 1 when a value is not an integer in [-2047, 2047];
// flags[1] = max over rows of the float bits of |x|^2 (non-negative floats order as unsigned).
__global__ void float_prepare_kernel(const float* __restrict__ blob, int kq, uint32_t total_rows, int cols,
                                     float* __restrict__ norms, unsigned int* __restrict__ flags) {
    const uint32_t row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= total_rows) return;
    const int lane = threadIdx.x & 31;
    const float* p = blob + (size_t)row * kq * 4;
    float acc = 0.f;
    bool bad = false;
    for (int c = lane; c < cols; c += 32) {
        const float v = p[c];
        bad |= !(fabsf(v) <= 2047.f) || (v != truncf(v));
        acc = fmaf(v, v, acc);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, d);
    bad = __any_sync(0xFFFFFFFFu, bad);
    if (lane == 0) {
        norms[row] = acc;
        if (bad) atomicOr(&flags[0], 1u);
        atomicMax(&flags[1], __float_as_uint(acc));
    }
}

// TF32-exact sets: the query-independent part of the epilogue's key argument, |t|^2 + 2^23 + 2^20 (an exact
// integer below 2^24).  x' = fmaf(q.t, -2, nb') = d^2 - |q|^2 + 2^23 + 2^20 lies in [2^23, 2^24): its low mantissa
// bits are that integer, it orders the columns of a row exactly like d^2, and |q|^2 is added back once per tile
// winner -- one FADD per accumulator element less than forming d^2 + 2^23 in place.
// The ranking pass of arbitrary floats uses the same table with offset C = max |x|^2 over the set: x' = |t|^2 - 2 q.t + C
// is >= 0 (up to the approximation error), so its float bits still order as signed integers, and d~^2 = x' - C + |q|^2
// is formed once per row at the end of the item.
static constexpr float FT_NB_OFFSET = 8388608.f + 1048576.f;
__global__ void float_nbexact_kernel(const float* __restrict__ norms, uint32_t n, float offset, float* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = norms[i] + offset;
}

// fp16 copy of a TF32-exact set (integers |v| <= 2047: exact in fp16's 11-bit significand), rows of `cols` halves.
__global__ void float_to_half_kernel(const float* __restrict__ blob, int kq, uint32_t total_rows, int cols, __half* __restrict__ out) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;  // one thread per pair of elements
    const size_t per_row = static_cast<size_t>(cols) / 2;
    if (i >= static_cast<size_t>(total_rows) * per_row) return;
    const size_t row = i / per_row, c = (i % per_row) * 2;
    const float2 v = *reinterpret_cast<const float2*>(blob + row * kq * 4 + c);
    *reinterpret_cast<__half2*>(out + row * cols + c) = __floats2half2_rn(v.x, v.y);
}

// TM_F16X operands of a TF32-exact set, rows of cols + 64 halves (one extra 128-byte K-block): see the mode's description.
//   train side  out_b : t_0 .. t_{c-1}, v0, v1, 2048 v2, 0 ...     v0 + 2048 v1 + 2048^2 v2 = |t|^2 + 2^23 + 2^20 (v0, v1 < 2048, v2 = 2)
//   query side  out_a : -2 q_0 .. -2 q_{c-1}, 1, 2048, 2048, 0 ...
// Every value is an integer of at most 12 significant bits times a power of two: exact in fp16.
__global__ void float_to_half_kx_kernel(const float* __restrict__ blob, int kq, uint32_t total_rows, int cols, const float* __restrict__ norms,
                                        __half* __restrict__ out_b, __half* __restrict__ out_a) {
    const int kx = cols + 64;
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;  // one thread per pair of output elements
    const size_t per_row = static_cast<size_t>(kx) / 2;
    if (i >= static_cast<size_t>(total_rows) * per_row) return;
    const size_t row = i / per_row;
    const int c = static_cast<int>(i % per_row) * 2;
    float b0 = 0.f, b1 = 0.f, a0 = 0.f, a1 = 0.f;
    if (c < cols) {
        const float2 v = *reinterpret_cast<const float2*>(blob + row * kq * 4 + c);
        b0 = v.x; b1 = v.y;
        a0 = -2.f * v.x; a1 = -2.f * v.y;
    } else if (c == cols || c == cols + 2) {
        const uint32_t V = static_cast<uint32_t>(norms[row] + FT_NB_OFFSET);  // exact integer below 2^24
        if (c == cols) {
            b0 = static_cast<float>(V & 2047u); b1 = static_cast<float>((V >> 11) & 2047u);
            a0 = 1.f; a1 = 2048.f;
        } else {
            b0 = static_cast<float>((V >> 22) * 2048u);
            a0 = 2048.f;
        }
    }
    *reinterpret_cast<__half2*>(out_b + row * kx + c) = __floats2half2_rn(b0, b1);
    *reinterpret_cast<__half2*>(out_a + row * kx + c) = __floats2half2_rn(a0, a1);
}

// ------------------------------------------------------------------ prepare (binary tensor engine)
// Bits -> {0,1} bytes, kbytes per row (a multiple of 128: whole swizzle rows), plus popcount per row.
// One warp per row; lane l expands the 16 bits [16l, 16l+16) of every 512-bit group.
__global__ void binary_unpack_kernel(const uint32_t* __restrict__ blob, int words, uint32_t total_rows, int kbytes,
                                     uint8_t* __restrict__ out, int32_t* __restrict__ norms) {
    const uint32_t row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= total_rows) return;
    const int lane = threadIdx.x & 31;
    const uint32_t* src = blob + (size_t)row * words;
    int pc = 0;
    for (int base = 0; base < kbytes; base += 512) {
        const int w = base / 32 + lane / 2;  // word holding this lane's 16 bits
        const uint32_t word = w < words ? src[w] : 0u;
        const uint32_t half = (word >> (16 * (lane & 1))) & 0xFFFFu;
        pc += __popc(half);
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t nib = (half >> (4 * j)) & 0xFu;
            o[j] = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
        }
        if (base + lane * 16 < kbytes)
            *reinterpret_cast<uint4*>(out + (size_t)row * kbytes + base + lane * 16) = make_uint4(o[0], o[1], o[2], o[3]);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) pc += __shfl_xor_sync(0xFFFFFFFFu, pc, d);
    if (lane == 0) norms[row] = pc;
}

// TM_F4P operands: bits -> E2M1 nibbles (1.0 = 0x2, 0.0 = 0x0), element 2j in the low nibble of byte j; kbytes per row (a multiple of
// 128), plus the popcount per row.  One warp per row; lane l expands the 8 bits [8l, 8l+8) of every 256-bit group into one 32-bit word.
__global__ void binary_unpack4_kernel(const uint32_t* __restrict__ blob, int words, uint32_t total_rows, int kbytes,
                                      uint8_t* __restrict__ out, int32_t* __restrict__ norms) {
    const uint32_t row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= total_rows) return;
    const int lane = threadIdx.x & 31;
    const uint32_t* src = blob + (size_t)row * words;
    int pc = 0;
    for (int base = 0; base < kbytes; base += 128) {  // 128 output bytes = 256 bits = 8 words per step
        const int w = base / 16 + lane / 4;
        const uint32_t word = w < words ? src[w] : 0u;
        const uint32_t byte = (word >> (8 * (lane & 3))) & 0xFFu;
        pc += __popc(byte);
        uint32_t o = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) o |= ((byte >> j) & 1u) << (4 * j + 1);
        *reinterpret_cast<uint32_t*>(out + (size_t)row * kbytes + base + lane * 4) = o;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) pc += __shfl_xor_sync(0xFFFFFFFFu, pc, d);
    if (lane == 0) norms[row] = pc;
}

// TM_F4X operands (see the mode): out_a = query side, bits as 1.0 (nibble 0x2), spare elements 6 x16, 1; out_b = train side, bits as 2.0
// (nibble 0x4), spare elements = digits of E = 512 - popc(t): floor(E / 36) elements of 6.0 (x 6 = 36), then the rest r = 3 b + c as
// two elements u/2 + v/2 (x 6) with u + v = b and one element c (x 1).  Every value is an E2M1 number, every product and every partial
// sum a small integer: exact.  One warp per row; the caller guarantees bits + 17 <= 2 * kbytes.
__global__ void binary_unpack4x_kernel(const uint32_t* __restrict__ blob, int words, uint32_t total_rows, int kbytes, int bits,
                                       uint8_t* __restrict__ out_b, uint8_t* __restrict__ out_a, int32_t* __restrict__ norms) {
    const uint32_t row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= total_rows) return;
    const int lane = threadIdx.x & 31;
    const uint32_t* src = blob + (size_t)row * words;
    int pc = lane < words ? __popc(src[lane]) : 0;  // words <= 32
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) pc += __shfl_xor_sync(0xFFFFFFFFu, pc, d);
    const int E = 512 - pc, a = E / 36, r = E - 36 * a, b = r / 3, c = r - 3 * b;
    // b = u + v in units of 3 (element values u/2, v/2 from {0, .5, 1, 1.5, 2, 3, 4}); nibble codes of the E2M1 values
    const int u = b <= 4 ? b : (b <= 7 ? (b == 5 ? 4 : 6) : 8), v = b - u;
    auto code = [](int units) -> uint32_t { return units <= 4 ? (uint32_t)units : (units == 6 ? 5u : 6u); };  // value units/2 -> nibble
    for (int base = 0; base < kbytes; base += 128) {
        const int w = base / 16 + lane / 4;
        const uint32_t word = w < words ? src[w] : 0u;
        const uint32_t byte = (word >> (8 * (lane & 3))) & 0xFFu;
        uint32_t oa = 0, ob = 0;
        const int e0 = (base + lane * 4) * 2;  // first of this lane's 8 elements
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int s = e0 + j - bits;  // spare element index
            uint32_t na = ((byte >> j) & 1u) << 1, nb = ((byte >> j) & 1u) << 2;
            if (s >= 0 && s < 17) {
                na = s < 16 ? 7u : 2u;
                nb = s < 14 ? (s < a ? 7u : 0u) : (s == 14 ? code(u) : (s == 15 ? code(v) : (uint32_t)(2 * c)));
            }
            oa |= na << (4 * j);
            ob |= nb << (4 * j);
        }
        *reinterpret_cast<uint32_t*>(out_a + (size_t)row * kbytes + base + lane * 4) = oa;
        *reinterpret_cast<uint32_t*>(out_b + (size_t)row * kbytes + base + lane * 4) = ob;
    }
    if (lane == 0) norms[row] = pc;
}

// Per train row of the binary tensor engine: the part of the top-2 key that does not depend on the query,
//     nbkey = (popc(t) + I8_BIAS) * 512 + (row inside its image mod 128),
// so that the epilogue forms its key with ONE IMAD per accumulator element:
//     key = acc * (-1024) + nbkey = (popc(t) - 2 q.t + I8_BIAS) << 9 | column-in-tile.
// popc(q) is the same for every column of a row: it does not change the order and is added when the
// tile's two winners are merged (hamming = (key >> 9) - I8_BIAS + popc(q)).  Train tiles start at multiples
// of 128 rows of the image, so the column inside the tile is a property of the row.  IMAD runs at 64
// lanes/clk/SM (profiles/pipe_bench_r01.txt): three of them per element were 768 cycles per 128x128 tile,
// more than the 512 cycles the MMAs of a 256-bit descriptor need.
static constexpr uint32_t I8_BIAS = 512;  // popc(t) - 2 q.t >= -popc(q) >= -512
//
// packed (TM_I8P, descriptors of fewer than 512 bits, bias = the bit length): the distance part needs 10 bits,
// so a key fits 16 bits with a 6-bit column code, and ONE register carries the keys of columns c (low half)
// and c + 16 (high half) of a 32-column chunk: col6 = (c & 15) | (chunk << 4).  The entry of a row with
// (column & 16) == 0 holds the query-independent parts of BOTH keys,
//     nbkey = ((popc(t_c) + bias) << 6 | col6) | ((popc(t_c+16) + bias) << 6 | col6) << 16,
// and the epilogue forms the pair with two IMADs:  hi = acc[c+16] * (-128 << 16) + nbkey;  key2 = acc[c] * -128 + hi.
// VIMNMX.U16x2 then keeps the two smallest keys of each half: 1.25 ALU ops per column instead of 2.5 --
// the ALU pipe (64 lanes/clk/SM) was the epilogue's floor at 640 cycles per 128x128 tile.
__global__ void binary_nbkey_kernel(const int32_t* __restrict__ popc, const uint32_t* __restrict__ row0 /* n_images, ascending */,
                                    int n_images, uint32_t total_rows, uint32_t* __restrict__ nbkey, uint32_t packed_bias /* 0 = 32-bit keys */,
                                    uint32_t f4_offset = 0 /* TM_F4P: 2^31, cancels the magic number's contribution (chunk_top2_packed) */) {
    const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= total_rows) return;
    int lo = 0, hi = n_images;  // last image whose first row is <= row
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (row0[mid] <= row) lo = mid; else hi = mid;
    }
    const uint32_t lc = (row - row0[lo]) & 127u;  // column inside its 128-row train tile
    if (packed_bias == 0) {
        nbkey[row] = (static_cast<uint32_t>(popc[row]) + I8_BIAS) * 512u + lc;
    } else {
        const uint32_t col6 = (lc & 15u) | ((lc >> 5) << 4);
        const uint32_t own = (static_cast<uint32_t>(popc[row]) + packed_bias) << 6 | col6;
        // the partner row may lie past the image (another image's row, or padding): its key is masked in the
        // kernel (partial tile), it only has to stay inside its 16 bits
        const uint32_t pc16 = row + 16 < total_rows ? static_cast<uint32_t>(popc[row + 16]) : 0u;
        const uint32_t partner = (pc16 + packed_bias) << 6 | col6;
        nbkey[row] = ((lc & 16u) ? own : (own | partner << 16)) + f4_offset;  // (entries with bit 4 set are not read)
    }
}

// ------------------------------------------------------------------ the kernel
// One work item (= one KnnTile), decoded by the prefetch warp for the other roles.
struct FtItem {
    uint32_t a_row;     // blob row of the tile's first query row
    uint32_t b_row0;    // blob row of the first train row of the range
    uint32_t n_rows;    // train rows in the range
    uint32_t n_tiles;   // 128-row train tiles
    uint32_t q0, nq;    // first query row of the tile (image-relative), rows of the query image
    uint32_t t0;        // first train row of the range (image-relative)
    uint32_t split;     // which partial list the tile writes
    uint32_t reverse;   // roles of the two images swapped (cross-check)
    uint32_t n_splits;
    uint32_t q_off;     // first row of the pair in the per-launch candidate scratch
    uint32_t pad;
    uint64_t knn_off, col_off;
};

struct FtSmem {  // after the 1024-byte aligned operand area
    uint64_t a_full, a_empty;
    uint64_t b_full[FT_B_STAGES];
    uint64_t b_empty[FT_B_STAGES];
    uint64_t acc_full[FT_ACC_STAGES];
    uint64_t acc_empty[FT_ACC_STAGES];
    uint64_t nb_full[FT_ACC_STAGES];
    uint64_t nb_empty[FT_ACC_STAGES];
    uint64_t item_full[2];
    uint64_t item_empty[2];
    uint32_t tmem_base;
    uint32_t pad;
    FtItem item[2];
    alignas(16) float nb[FT_ACC_STAGES][FT_N];  // |t|^2 of the train tile, bulk-copied next to the operands
    float rowval[2][FT_M];                      // per query row of the item: |q|^2 (+2^23) / popc(q) / collect threshold
    uint4 merge[2][FT_M];                       // (d2_1, i1, d2_2, i2) of the epilogue group with the odd tiles
};

static inline size_t float_tensor_smem_bytes(int kblocks) {
    return 1024 /*alignment slack*/ + (size_t)kblocks * (FT_KBLOCK_BYTES + FT_B_STAGES * FT_B_KBLOCK_BYTES) + sizeof(FtSmem);
}

// Top-2 bookkeeping of the epilogue.
//
// For TF32-exact data with row norm^2 <= 2^20 every d^2 = |q|^2 + |t|^2 - 2 q.t is an integer below
// 2^22.  Adding 2^23 in fp32 (exact) leaves that integer in the low mantissa bits, so
//     key = float_bits(d^2 + 2^23) * 512 + column      (one IMAD; the exponent bits shift out)
// is the 31-bit integer  d^2 << 9 | column-in-tile : unsigned min/max on it orders by
// (distance, lowest column) -- no branches, no divergence, data-independent cost.  Inside a tile
// the two smallest keys are kept with 5 VIMNMX per two columns; once per tile they are merged
// into the running (d^2, train index) pair with a strict '<' in arrival order, which is
// cv::batchDistance's insertion rule.  sqrtf is injective on integers below 2^22 (consecutive
// roots are more than an ulp apart), so ties and order on d^2 are ties and order on d.
struct Top2 {
    uint32_t d1, d2;  // squared distances (integers), 0xFFFFFFFF = none
    int i1, i2;
    __device__ __forceinline__ void init() {
        d1 = d2 = 0xFFFFFFFFu;
        i1 = i2 = -1;
    }
    __device__ __forceinline__ void offer(uint32_t d, int t) {  // strict '<', arrival order
        const bool lt1 = d < d1, lt2 = d < d2;
        d2 = lt1 ? d1 : (lt2 ? d : d2);
        i2 = lt1 ? i1 : (lt2 ? t : i2);
        d1 = lt1 ? d : d1;
        i1 = lt1 ? t : i1;
    }
};

// TM_F4X: accumulator x = 2 q.t - popc(t) + 512 (fp32, an integer in [24, 1000]) -> key (popc(t) - 2 q.t + 512) << 9 | column, the
// TM_I8 key.  (2^23 + 1024) - x is exact and leaves 1024 - x in the low mantissa bits; the exponent bits vanish in the product by 512.
static constexpr float F4X_KMAGIC = 8388608.f + 1024.f;
__device__ __forceinline__ uint32_t f4x_key(uint32_t acc_bits, uint32_t key_mul, uint32_t lc) {
    const uint32_t b = __float_as_uint(F4X_KMAGIC - __uint_as_float(acc_bits));
    uint32_t k;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(k) : "r"(b), "r"(key_mul), "r"(lc));
    return k;
}

// (m1 <= m2) <- two smallest of {m1, m2, a, b}
__device__ __forceinline__ void top2_pair(uint32_t& m1, uint32_t& m2, uint32_t a, uint32_t b) {
    const uint32_t lo = min(a, b), hi = max(a, b);
    const uint32_t loser = max(m1, lo);
    m1 = min(m1, lo);
    m2 = __vimin3_u32(m2, loser, hi);
}

// Keys of this thread's 64 columns of one accumulator tile and their two smallest.
//   d^2 + 2^23 = (|t|^2 + |q|^2 + 2^23) - 2 q.t   FADD + FFMA, exact
//   key        = bits * 512 + column               one IMAD (key_mul = 512 comes from the kernel
//                                                  parameters so that it is not strength-reduced
//                                                  into a shift + add on the ALU pipe)
// PARTIAL: columns past the train set (another image's rows / padding) get the "none" key.
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}

template <bool PARTIAL, int MODE, int W = 32>
__device__ __forceinline__ void chunk_top2(const uint32_t (&acc)[W], uint32_t nb_saddr, float cq, uint32_t key_mul,
                                           uint32_t lc0 /* first column of the chunk inside the tile */, uint32_t col0 /* same, relative to t0 */,
                                           uint32_t n_rows, uint32_t& m1, uint32_t& m2) {
#pragma unroll
    for (int e = 0; e < W; e += 4) {
        float4 nb = make_float4(0.f, 0.f, 0.f, 0.f);
        if constexpr (MODE != TM_F16X) nb = lds128(nb_saddr + e * 4);  // same address in every lane: broadcast
        [[maybe_unused]] const float nbv[4] = {nb.x, nb.y, nb.z, nb.w};
        uint32_t k[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            [[maybe_unused]] uint32_t bits = 0;
            if constexpr (MODE == TM_F16X) {
                // the accumulator is the key argument itself (see TM_F16X): key = bits * 512 + column, one IMAD
                const uint32_t lc = lc0 + e + i;
                asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(k[i]) : "r"(acc[e + i]), "r"(key_mul), "r"(lc));
            } else if constexpr (MODE == TM_F4X) {
                k[i] = f4x_key(acc[e + i], key_mul, lc0 + e + i);
            } else if constexpr (MODE == TM_I8) {
                // key = acc * (-1024) + nbkey: see binary_nbkey_kernel (key_mul - 1536 = -1024 comes from a kernel
                // parameter so that the multiply stays one IMAD on the FMA pipe instead of a shift + subtract on the ALU)
                asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(k[i]) : "r"(acc[e + i]), "r"(key_mul - 1536u), "r"(__float_as_uint(nbv[i])));
            } else {
                bits = __float_as_uint(fmaf(__uint_as_float(acc[e + i]), -2.f, nbv[i]));  // see float_nbexact_kernel
                const uint32_t lc = lc0 + e + i;  // column inside the 128-wide tile
                asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(k[i]) : "r"(bits), "r"(key_mul), "r"(lc));
            }
            if (PARTIAL) k[i] = col0 + e + i < n_rows ? k[i] : 0xFFFFFFFFu;
        }
        top2_pair(m1, m2, k[0], k[1]);
        top2_pair(m1, m2, k[2], k[3]);
    }
}

// TM_F16X with threshold skipping (round 2).  The exact fold above costs 2.5 VIMNMX per column at 64 lanes/clk/SM = 640 ALU
// cycles per 128x128 tile against 576 cycles of MMAs: the ALU pipe, not the tensor pipe, set the pace (ncu r02: alu 71 %,
// tensor 53 %).  But once a row has seen a few hundred train rows almost no column can still enter its top-2 -- a column matters
// only if its distance is below the row's current second-smallest, which happens ~2 ln(Nt) times per row -- and in TM_F16X the
// accumulator itself orders like the key (float bits of an integer in [2^23, 2^24): monotone as unsigned).  So the chunk is
// reduced to one minimum per 8 columns with VIMNMX3 on the RAW accumulators (0.56 ALU ops per column, no IMAD), one warp vote
// decides whether any of the warp's 32 rows has a column below its threshold `thrv`, and only then the 8-column blocks that hold
// such a column get their keys formed and folded exactly.  The four block votes are issued together (no dependent vote -> branch
// chain); `thrv` only ever decreases: from the tile's own second key after a chunk, from the merged list at the end of a tile and
// from the other epilogue group's threshold (+1, see the kernel).  A column whose value EQUALS the threshold is skipped: columns
// arrive in ascending index order, so an equal distance further on never replaces (cv::batchDistance's strict '<').
// Cost is data dependent (train rows arriving by descending distance take the slow path every time), the result is not.
template <int W = 32>
__device__ __forceinline__ void chunk_top2_skipx(const uint32_t (&acc)[W], uint32_t key_mul, uint32_t lc0 /* first column of the chunk inside the tile */,
                                                 uint32_t& thrv /* accumulator bits: only columns below can matter */, uint32_t& m1, uint32_t& m2) {
    static_assert(W % 8 == 0, "8-column blocks");
    constexpr int NB = W / 8;
    uint32_t s[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const uint32_t* a = &acc[8 * b];
        s[b] = min(__vimin3_u32(__vimin3_u32(a[0], a[1], a[2]), __vimin3_u32(a[3], a[4], a[5]), a[6]), a[7]);
    }
    uint32_t mn = NB == 4 ? min(__vimin3_u32(s[0], s[1], s[2]), s[NB - 1]) : min(s[0], s[NB - 1]);
    if (__any_sync(0xFFFFFFFFu, mn < thrv)) {
        bool hit[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) hit[b] = __any_sync(0xFFFFFFFFu, s[b] < thrv);
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            if (hit[b]) {
                uint32_t k[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint32_t lc = lc0 + 8 * b + i;
                    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(k[i]) : "r"(acc[8 * b + i]), "r"(key_mul), "r"(lc));
                }
                top2_pair(m1, m2, k[0], k[1]);
                top2_pair(m1, m2, k[2], k[3]);
                top2_pair(m1, m2, k[4], k[5]);
                top2_pair(m1, m2, k[6], k[7]);
            }
        }
        // two keys of this tile are <= m2, so the row's second-smallest is; back to accumulator bits (m2 = none: 0x4B7FFFFF, above every accumulator)
        thrv = min(thrv, 0x4B000000u | (m2 >> 9));
    }
}

// The same for TM_F4X, where LARGER accumulators are nearer: block maxima of the raw (positive) floats against the float threshold
// `thrf` = the accumulator value of the row's current second-nearest column (-1: none yet, +inf: row never written).
template <int W = 32>
__device__ __forceinline__ void chunk_top2_skipx_f4(const uint32_t (&acc)[W], uint32_t key_mul, uint32_t lc0, float& thrf, uint32_t& m1, uint32_t& m2) {
    static_assert(W % 8 == 0, "8-column blocks");
    constexpr int NB = W / 8;
    uint32_t s[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const uint32_t* a = &acc[8 * b];
        s[b] = max(__vimax3_u32(__vimax3_u32(a[0], a[1], a[2]), __vimax3_u32(a[3], a[4], a[5]), a[6]), a[7]);
    }
    const uint32_t mx = NB == 4 ? max(__vimax3_u32(s[0], s[1], s[2]), s[NB - 1]) : max(s[0], s[NB - 1]);
    if (__any_sync(0xFFFFFFFFu, __uint_as_float(mx) > thrf)) {
        bool hit[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) hit[b] = __any_sync(0xFFFFFFFFu, __uint_as_float(s[b]) > thrf);
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            if (hit[b]) {
                uint32_t k[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) k[i] = f4x_key(acc[8 * b + i], key_mul, lc0 + 8 * b + i);
                top2_pair(m1, m2, k[0], k[1]);
                top2_pair(m1, m2, k[2], k[3]);
                top2_pair(m1, m2, k[4], k[5]);
                top2_pair(m1, m2, k[6], k[7]);
            }
        }
        // accumulator value of the tile's second key: (2^23 + 1024) - (2^23 + v) = 1024 - v, exact (m2 = none: far below zero)
        thrf = fmaxf(thrf, F4X_KMAGIC - __uint_as_float(0x4B000000u | (m2 >> 9)));
    }
}

// TM_I8P: see binary_nbkey_kernel.  16 packed key pairs per 32-column chunk.
__device__ __forceinline__ void top2_pair_u16x2(uint32_t& m1, uint32_t& m2, uint32_t a, uint32_t b) {
    const uint32_t lo = __vminu2(a, b), hi = __vmaxu2(a, b);
    const uint32_t loser = __vmaxu2(m1, lo);
    m1 = __vminu2(m1, lo);
    m2 = __vimin3_u16x2(m2, loser, hi);
}

// F4 (TM_F4P): the accumulators are fp32 integers; adding 2^23 leaves the integer in the low mantissa bits, and the exponent bits
// 0x4B000000 drop out of the products -- times (-128 << 16) entirely, times -128 up to the constant 2^31 that binary_nbkey_kernel
// folds into the table -- so the same two IMADs form the same key pair: one FADD per column more, on the FMA pipe.
template <bool PARTIAL, bool F4 = false>
__device__ __forceinline__ void chunk_top2_packed(const uint32_t (&acc)[32], uint32_t nb_saddr, uint32_t mul_lo /* -128 */, uint32_t mul_hi /* -128 << 16 */,
                                                  uint32_t col0 /* first column of the chunk, relative to t0 */, uint32_t n_rows,
                                                  uint32_t& m1, uint32_t& m2) {
#pragma unroll
    for (int e = 0; e < 16; e += 4) {
        const float4 nb = lds128(nb_saddr + e * 4);  // entries 0..15 of the chunk: the pair constants
        const uint32_t nbv[4] = {__float_as_uint(nb.x), __float_as_uint(nb.y), __float_as_uint(nb.z), __float_as_uint(nb.w)};
        uint32_t k[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t hi, a_hi = acc[e + i + 16], a_lo = acc[e + i];
            if constexpr (F4) {
                a_hi = __float_as_uint(__uint_as_float(a_hi) + 8388608.f);
                a_lo = __float_as_uint(__uint_as_float(a_lo) + 8388608.f);
            }
            asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(hi) : "r"(a_hi), "r"(mul_hi), "r"(nbv[i]));
            asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(k[i]) : "r"(a_lo), "r"(mul_lo), "r"(hi));
            if (PARTIAL) k[i] |= (col0 + e + i < n_rows ? 0u : 0xFFFFu) | (col0 + e + i + 16 < n_rows ? 0u : 0xFFFF0000u);
        }
        top2_pair_u16x2(m1, m2, k[0], k[1]);
        top2_pair_u16x2(m1, m2, k[2], k[3]);
    }
}

// Tile end of TM_I8P: the two smallest 32-bit keys (hamming << 9 | column-in-tile) among the four 16-bit lane winners.
__device__ __forceinline__ void unpack_top2_u16x2(uint32_t m1, uint32_t m2, uint32_t dadd /* popc(q) - bias */, uint32_t& q1, uint32_t& q2) {
    uint32_t c[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t k16 = ((i & 2) ? m2 : m1) >> (16 * (i & 1)) & 0xFFFFu;
        const uint32_t col6 = k16 & 63u;
        const uint32_t col = (col6 & 15u) + 16u * (i & 1) + 32u * (col6 >> 4);
        c[i] = k16 >= 0xFFC0u ? 0xFFFFFFFFu : (((k16 >> 6) + dadd) << 9 | col);
    }
    const uint32_t lo01 = min(c[0], c[1]), hi01 = max(c[0], c[1]), lo23 = min(c[2], c[3]), hi23 = max(c[2], c[3]);
    q1 = min(lo01, lo23);
    q2 = min(max(lo01, lo23), min(hi01, hi23));
}

// TM_TF32_RANK: only the two smallest VALUES of the row matter (pass 2 re-derives the columns), so the key is
// the raw float, ordered as a signed integer: the float order for values >= 0.  Slightly negative values (only
// possible within the error bound of 0) sort first in scrambled order; the caller clamps the result at 0, which
// keeps the threshold an upper bound (see collect_threshold).  FFMA + 2.5 VIMNMX per column.
__device__ __forceinline__ void top2_pair_s32(int& m1, int& m2, int a, int b) {
    const int lo = min(a, b), hi = max(a, b);
    const int loser = max(m1, lo);
    m1 = min(m1, lo);
    m2 = __vimin3_s32(m2, loser, hi);
}

template <bool PARTIAL>
__device__ __forceinline__ void chunk_rank(const uint32_t (&acc)[32], uint32_t nb_saddr, uint32_t col0,
                                           uint32_t n_rows, int& m1, int& m2) {
#pragma unroll
    for (int e = 0; e < 32; e += 4) {
        const float4 nb = lds128(nb_saddr + e * 4);
        const float nbv[4] = {nb.x, nb.y, nb.z, nb.w};
        int k[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            k[i] = __float_as_int(fmaf(__uint_as_float(acc[e + i]), -2.f, nbv[i]));  // nbv = |t|^2 + C (float_nbexact_kernel)
            if (PARTIAL) k[i] = col0 + e + i < n_rows ? k[i] : 0x7FFFFFFF;
        }
        top2_pair_s32(m1, m2, k[0], k[1]);
        top2_pair_s32(m1, m2, k[2], k[3]);
    }
}

// TM_TF32_COLLECT: append every column of the chunk whose approximate d^2 is <= tau to this thread's
// half of the row's candidate list (FT_CAND_CAP/2 slots per epilogue group).  One FFMA + ~half an FMNMX
// per column and ONE branch per 32 columns: hits are rare (a handful of columns per row over the whole
// train set), and a data-dependent branch cannot be scheduled across -- a branch every 4 columns left the
// epilogue latency-bound at 19 % issue utilisation (profiles/ncu_float_refine_r01.txt).
// The slow path is kept SMALL: it only builds a per-thread bit mask of the hits (FSETP + predicated OR per
// column); the appends run in a rolled loop over the set bits.  The first version appended inline, column by
// column -- 6 instructions x 32 columns x 4 chunks x 2 variants = an 88 KB kernel whose warps stalled on
// instruction fetch (stall_no_inst on every opcode, same profile).
// With one split per pair the two owners of a row are the only writers, so the fill count lives in a
// register (`shared_fill` false); with several splits it is a global counter.
__device__ __forceinline__ void chunk_collect(const uint32_t (&acc)[32], uint32_t nb_saddr, float tau, uint32_t valid /* bit i: column i is a train row */,
                                              uint32_t t_first, bool shared_fill, uint32_t& fill, uint32_t* fill_global,
                                              uint32_t* __restrict__ list) {
    float x[32];
#pragma unroll
    for (int e = 0; e < 32; e += 4) {
        const float4 nb = lds128(nb_saddr + e * 4);
        // tau already has |q|^2 subtracted: compare |t|^2 - 2 q.t, one FFMA per column
        x[e + 0] = fmaf(__uint_as_float(acc[e + 0]), -2.f, nb.x);
        x[e + 1] = fmaf(__uint_as_float(acc[e + 1]), -2.f, nb.y);
        x[e + 2] = fmaf(__uint_as_float(acc[e + 2]), -2.f, nb.z);
        x[e + 3] = fmaf(__uint_as_float(acc[e + 3]), -2.f, nb.w);
    }
    float mn = fmin3(x[0], x[1], x[2]);
#pragma unroll
    for (int e = 3; e < 31; e += 2) mn = fmin3(mn, x[e], x[e + 1]);
    mn = fminf(mn, x[31]);
    if (mn <= tau) {
        uint32_t h[4] = {0, 0, 0, 0};  // four independent chains
#pragma unroll
        for (int i = 0; i < 32; ++i) h[i & 3] |= x[i] <= tau ? (1u << i) : 0u;
        uint32_t hits = ((h[0] | h[1]) | (h[2] | h[3])) & valid;
#pragma unroll 1
        while (hits) {
            const uint32_t i = __ffs(hits) - 1;
            hits &= hits - 1;
            const uint32_t pos = shared_fill ? atomicAdd(fill_global, 1u) : fill++;
            if (pos < FT_CAND_CAP / 2) list[pos] = t_first + i;
        }
    }
}

// Pass-2 threshold of one query row (TM_TF32_COLLECT), with |q|^2 already moved to the threshold's side.
// m2 = approximate second-smallest d^2 of this row from pass 1 (merged over the splits, clamped at 0: if two or
// more approximations were negative, pass 1's signed-integer order may have kept the wrong two of them, but then
// every one of them -- and the true second-smallest -- is <= 0 + eps).  |approx - exact| <= eps with
// eps = 2^-8 |q| max|t| (both operands truncated to TF32: relative error < 2^-10 each, Cauchy-Schwarz)
// + accumulation / norm rounding slack; every column of the exact top-2 has approx <= m2 + 2 eps.
template <bool F16>
__device__ __forceinline__ float collect_threshold(bool valid, float nq2, const PairDesc& pd, const KnnEntry* __restrict__ knn, uint32_t qrow,
                                                   float rank_offset /* C of the ranking pass's key table */) {
    if (!valid) return __int_as_float(0xff800000);  // -inf: rows past the image collect nothing (their accumulators are
                                                    // dot products with some other image's rows and can be anything)
    unsigned long long k1 = KEY_NONE, k2 = KEY_NONE;
    for (uint32_t sidx = 0; sidx < pd.n_splits; ++sidx) {
        const KnnEntry e = knn[pd.knn_off + (size_t)sidx * pd.nq + qrow];
        const unsigned long long hi = max(k1, e.x);
        k1 = min(k1, e.x);
        k2 = min(min(k2, hi), e.y);
    }
    const uint32_t m2bits = static_cast<uint32_t>(k2 >> 32);
    const float m2 = m2bits >= 0x7f800000u ? __int_as_float(0x7f7fffff) : __uint_as_float(m2bits);  // none / inf / NaN: collect all
    // TF32 operands are truncated (relative error < 2^-10 each): |2 q~.t~ - 2 q.t| <= 2^-8 |q||t|.
    // fp16 operands are rounded to nearest (relative error <= 2^-11 each, plus <= 2^-25 absolute in the subnormal range):
    // <= 2^-9 |q||t| + 2^-24 sqrt(d) (|q| + |t|).  The 2 % on top covers the fp32 accumulation of the tensor core.
    const float rel = F16 ? 0.001953125f : 0.00390625f;
    // The ranking pass formed x' = |t|^2 - 2 q.t + C and d~^2 = x' - C + |q|^2 in fp32 at magnitudes up to 4C: two roundings
    // of at most 2^-24 * 4C each, covered by 1e-6 * C (negligible unless one row of the set dwarfs this pair's norms).
    const float eps = rel * 1.02f * sqrtf(nq2) * sqrtf(pd.t_maxnorm2) + (F16 ? 2e-6f : 1e-6f) * (nq2 + pd.t_maxnorm2) + (F16 ? 2e-6f : 0.f) +
                      1e-6f * rank_offset;
    return m2 + 2.f * eps - nq2 + 1e-6f * (m2 + nq2);  // (last term: rounding of moving |q|^2 across)
}

// End of an item for the eight epilogue warps: merge the two groups' lists of each row -- lexicographic
// (d^2, index) -- and write the row's entry; d = sqrtf(d^2) of an exact integer is bit-identical to OpenCV's
// sqrtf(sum (a-b)^2).  `merge` alternates between items: group 1 may already be an item ahead when group 0 reads.
template <int MODE, int GROUPS = 2>
__device__ __forceinline__ void finish_rows(uint4* merge /* [GROUPS-1][FT_M] */, const Top2& best, uint32_t grp, uint32_t row, uint32_t qrow, uint32_t nq,
                                            bool reverse, uint32_t split, unsigned long long knn_off, unsigned long long col_off,
                                            KnnEntry* __restrict__ knn, unsigned long long* __restrict__ colmin, uint32_t out_row) {
    if (grp != 0) merge[(grp - 1) * FT_M + row] = make_uint4(best.d1, (uint32_t)best.i1, best.d2, (uint32_t)best.i2);
    asm volatile("bar.sync 1, %0;" ::"n"(128 * GROUPS) : "memory");  // the epilogue warps only
    if (grp == 0 && qrow < nq) {
        unsigned long long k1 = best.i1 < 0 ? KEY_NONE : make_key(best.d1, (uint32_t)best.i1);
        unsigned long long k2 = best.i2 < 0 ? KEY_NONE : make_key(best.d2, (uint32_t)best.i2);
#pragma unroll
        for (int g = 0; g < GROUPS - 1; ++g) {
            const uint4 o = merge[g * FT_M + row];
            const unsigned long long o1 = (int)o.y < 0 ? KEY_NONE : make_key(o.x, o.y);
            const unsigned long long o2 = (int)o.w < 0 ? KEY_NONE : make_key(o.z, o.w);
            const unsigned long long hi = max(k1, o1);
            k1 = min(k1, o1);
            k2 = min(min(k2, hi), o2);
        }
        if (reverse) {  // column minimum of the forward problem: (d^2, lowest query index); splits merge by atomicMin
            if (k1 != KEY_NONE) atomicMin(colmin + col_off + out_row, k1);  // (out_row == qrow unless the rows were gathered)
        } else {
            KnnEntry e;
            if constexpr (MODE == TM_I8 || MODE == TM_I8P || MODE == TM_F4P || MODE == TM_F4X || tm_is_rank(MODE)) {
                // i8: the Hamming distance stays an integer in the key (binary_knn.cuh's convention);
                // rank pass: float bits of the approximate d^2 (only pass 2 reads it)
                e.x = k1;
                e.y = k2;
            } else {  // integer d^2 -> float bits of d
                e.x = k1 == KEY_NONE ? KEY_NONE : make_key(__float_as_uint(sqrtf((float)(uint32_t)(k1 >> 32))), (uint32_t)k1);
                e.y = k2 == KEY_NONE ? KEY_NONE : make_key(__float_as_uint(sqrtf((float)(uint32_t)(k2 >> 32))), (uint32_t)k2);
            }
            knn[knn_off + (size_t)split * nq + qrow] = e;
        }
    }
}

// PERSISTENT kernel: one CTA per SM walks the launch's KnnTile list with stride gridDim.x.  The TMA / MMA /
// epilogue pipeline keeps running across item boundaries (ring indices and mbarrier phases follow a tile counter
// that never resets), so the epilogue of item k -- draining the 4 TMEM stages, merging the two groups' lists,
// writing the row results -- overlaps the loads and MMAs of item k+1, and barrier init / TMEM allocation /
// tensor-map prefetch are paid once per SM instead of once per 128 query rows.  With one CTA per (pair, query
// tile) the fixed cost was ~4 us against 30 us of tensor work at cfg-2 (40 train tiles per item) and against
// 6 us for ORB-size images.
//
//   warp 0      TMA producer  : query tile per item (as soon as the previous item's MMAs released it), then
//                               128-row train tiles through a 2-stage smem ring + their norms (4-slot ring)
//   warp 1      MMA issuer    : whole warp, uniform control flow, elect.sync picks the issuing lane
//   warp 2      TMEM allocator
//   warp 3      item prefetch : decodes the next KnnTile/PairDesc and computes the per-row constants
//                               (|q|^2, popc(q), or the pass-2 threshold) into a 2-slot smem ring, off the critical path
//   warps 4-11  epilogue      : two groups of four warps on alternate tiles
template <int KB /* 128-byte K-blocks per row: 4 for 128-d float / 512-bit binary */, int MODE>
__global__ void __launch_bounds__(FT_THREADS, 1)
tensor_knn2_kernel(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ norms /* int32 popcounts when INT8 */,
                   const float* __restrict__ nb_src /* per train row: |t|^2, or binary_nbkey_kernel's key part when INT8 */,
                   const KnnTile* __restrict__ tiles, const uint32_t n_items, const PairDesc* __restrict__ pairs,
                   KnnEntry* __restrict__ knn, unsigned long long* __restrict__ colmin, const uint32_t key_mul /* = 512 */, const uint32_t i8_bias /* TM_I8P: the descriptors' bit length; rank / collect modes: float bits of the key-table offset C */,
                   uint32_t* __restrict__ cand_count, uint32_t* __restrict__ cand_idx /* TM_TF32_COLLECT */) {
    constexpr int KIND = OperandOf<MODE>::kind;
    constexpr int KB_ELEMS = OperandOf<MODE>::kb_elems;
    extern __shared__ unsigned char ft_smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ft_smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* sA = base;                                   // KB x 16 KB
    unsigned char* sB = base + (size_t)KB * FT_KBLOCK_BYTES;    // FT_B_STAGES x KB x 16 KB
    FtSmem& sm = *reinterpret_cast<FtSmem*>(base + (size_t)KB * (FT_KBLOCK_BYTES + FT_B_STAGES * FT_B_KBLOCK_BYTES));

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbar_init(&sm.a_full, 1);
        mbar_init(&sm.a_empty, 1);
        for (int s = 0; s < FT_B_STAGES; ++s) {
            mbar_init(&sm.b_full[s], 1);
            mbar_init(&sm.b_empty[s], 1);
        }
        for (int s = 0; s < FT_ACC_STAGES; ++s) {
            mbar_init(&sm.acc_full[s], 1);
            mbar_init(&sm.acc_empty[s], FT_EPI_WARPS / 2);  // the four warps of the group that owns the tile
            mbar_init(&sm.nb_full[s], 1);
            mbar_init(&sm.nb_empty[s], FT_EPI_WARPS / 2);
        }
        for (int s = 0; s < 2; ++s) {
            // every thread that writes / reads an item slot arrives itself (release / acquire pair per thread)
            mbar_init(&sm.item_full[s], 32);                       // the prefetch warp's lanes
            mbar_init(&sm.item_empty[s], 1 + 32 + 32 * FT_EPI_WARPS);  // producer thread, MMA warp, epilogue warps
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

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
            uint32_t g = 0, it = 0;  // tiles / items this CTA has gone through: ring slots and mbarrier phases
            for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const uint32_t slot = it & 1;
                mbar_wait(&sm.item_full[slot], (it >> 1) & 1);
                const uint32_t a_row = sm.item[slot].a_row, b_row0 = sm.item[slot].b_row0, n_tiles = sm.item[slot].n_tiles;
                mbar_arrive(&sm.item_empty[slot]);
                mbar_wait(&sm.a_empty, (it & 1) ^ 1);  // the previous item's MMAs have read the query tile
                mbar_expect_tx(&sm.a_full, KB * FT_KBLOCK_BYTES);
#pragma unroll
                for (int kb = 0; kb < KB; ++kb)
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        tma_load_2d(sA + kb * FT_KBLOCK_BYTES + h * FT_BOX_BYTES, &tmap, kb * KB_ELEMS,
                                    (int)a_row + h * FT_BOX_ROWS, &sm.a_full);
#pragma unroll 1
                for (uint32_t j = 0; j < n_tiles; ++j, ++g) {
                    const uint32_t s = g % FT_B_STAGES;
                    const uint32_t a = g % FT_ACC_STAGES;
                    mbar_wait(&sm.nb_empty[a], ((g / FT_ACC_STAGES) & 1) ^ 1);
                    mbar_expect_tx(&sm.nb_full[a], FT_N * sizeof(float));
                    // image rows start at multiples of 4 and t0 at multiples of 128: 16-byte aligned source;
                    // the norms array is padded so that the copy may run past the image's last row
                    tma_load_1d(sm.nb[a], nb_src + b_row0 + j * FT_N, FT_N * sizeof(float), &sm.nb_full[a]);
                    mbar_wait(&sm.b_empty[s], ((g / FT_B_STAGES) & 1) ^ 1);
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
        const uint64_t a_desc0 = umma_desc_sw128(smem_u32(sA));
        const uint32_t idesc = KIND == OK_I8 ? FT_IDESC_I8 : (KIND == OK_F16 ? FT_IDESC_F16 : FT_IDESC);
        uint32_t g = 0, it = 0;
        for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const uint32_t slot = it & 1;
            mbar_wait(&sm.item_full[slot], (it >> 1) & 1);
            const uint32_t n_tiles = __shfl_sync(0xFFFFFFFFu, sm.item[slot].n_tiles, 0);
            mbar_arrive(&sm.item_empty[slot]);  // every thread, after its own reads of the slot
            mbar_wait(&sm.a_full, it & 1);
#pragma unroll 1
            for (uint32_t j = 0; j < n_tiles; ++j, ++g) {
                const uint32_t s = g % FT_B_STAGES, a = g % FT_ACC_STAGES;
                mbar_wait(&sm.b_full[s], (g / FT_B_STAGES) & 1);
                mbar_wait(&sm.acc_empty[a], ((g / FT_ACC_STAGES) & 1) ^ 1);
                tc_fence_after();
                const uint64_t b_desc0 = umma_desc_sw128(smem_u32(sB + (size_t)s * KB * FT_B_KBLOCK_BYTES));
                const uint32_t d_tmem = tb + a * FT_N;
                // the start-address field counts 16-byte units: stepping inside the tile is an integer add
                tc_mma<KIND, false>(d_tmem, a_desc0, b_desc0, idesc);
#pragma unroll
                for (int i = 1; i < 4 * KB; ++i) {  // i = kb*4 + k: 4 x (32 bytes of K) inside each 128-byte swizzle row
                    const int kb = i >> 2, k = i & 3;
                    tc_mma<KIND, true>(d_tmem, a_desc0 + ((kb * FT_KBLOCK_BYTES + k * 32) >> 4), b_desc0 + ((kb * FT_B_KBLOCK_BYTES + k * 32) >> 4), idesc);
                }
                tc_commit_elect(&sm.b_empty[s]);   // smem stage reusable once these MMAs have read it
                tc_commit_elect(&sm.acc_full[a]);  // accumulator ready for the epilogue
            }
            tc_commit_elect(&sm.a_empty);  // every MMA of this item has read the query tile: the producer may replace it
        }
    } else if (warp == 3) {
        // ===================== item prefetch =====================
        uint32_t it = 0;
        for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const uint32_t slot = it & 1;
            mbar_wait_relaxed(&sm.item_empty[slot], ((it >> 1) & 1) ^ 1);
            KnnTile tile = tiles[item];
            PairDesc pd = pairs[tile.pair];
            // Symmetric cross-check = the same problem with the roles swapped: a "reverse" tile (bit 31 of
            // split) takes its rows from the train image and streams the query image; its row-wise 1-NN is
            // the column minimum the filter needs (lowest query index on ties, by the same insertion rule).
            const bool reverse = (tile.split & TILE_REVERSE) != 0;  // (gathered reverse tiles are a TMEM-A kernel feature)
            tile.split &= TILE_SPLIT_MASK;
            if (reverse) {
                const uint32_t r0 = pd.q_row0, n = pd.nq;
                pd.q_row0 = pd.t_row0; pd.nq = pd.nt;
                pd.t_row0 = r0; pd.nt = n;
            }
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
                const float nq2 = valid ? __ldg(norms + pd.q_row0 + qrow) : 0.f;
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
    } else if (warp >= 4) {
        // ===================== epilogue: fused top-2 =====================
        // Two groups of four warps (one warp per TMEM lane quarter) take alternate tiles (by the running tile
        // counter, so that odd tile counts do not favour one group), so that on every SM sub-partition one
        // warp computes while the other waits for its barrier / TMEM load.
        const uint32_t ew = warp - 4;
        const uint32_t quarter = ew & 3, half = ew >> 2;   // TMEM lanes 32*quarter..
        const uint32_t row = quarter * 32 + lane;          // row of the query tile == TMEM lane
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
#pragma unroll 1
            for (uint32_t j = (half ^ g0) & 1; j < n_tiles; j += 2) {
                const uint32_t g = g0 + j;
                const uint32_t a = g % FT_ACC_STAGES;
                mbar_wait(&sm.acc_full[a], (g / FT_ACC_STAGES) & 1);
                tc_fence_after();
                uint32_t acc[2][32];  // register double buffer: chunk c+1 streams in from TMEM while chunk c is folded
                const uint32_t taddr = tmem_base + ((quarter * 32) << 16) + a * FT_N;
                tc_ld_32x32(taddr, acc[0]);
                mbar_wait(&sm.nb_full[a], (g / FT_ACC_STAGES) & 1);
                tc_wait_ld(acc[0]);
                const uint32_t col0 = j * FT_N;                // first column of the tile, relative to t0
                const bool partial = col0 + FT_N > n_rows;     // warp-uniform: only the last tile
                uint32_t m1 = 0xFFFFFFFFu, m2 = 0xFFFFFFFFu;   // two smallest keys of this tile
                constexpr int NCH = FT_N / 32;  // 32-column chunks per tile
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    if (c < NCH - 1) tc_ld_32x32(taddr + (c + 1) * 32, acc[(c + 1) & 1]);
                    const uint32_t nb_saddr = smem_u32(&sm.nb[a][c * 32]);
                    if constexpr (tm_is_collect(MODE)) {
                        const uint32_t c0 = col0 + c * 32;  // columns at or past n_rows belong to another image / padding
                        const uint32_t valid = c0 + 32 <= n_rows ? 0xFFFFFFFFu : (c0 < n_rows ? (1u << (n_rows - c0)) - 1u : 0u);
                        chunk_collect(acc[c & 1], nb_saddr, cq, valid, t0 + c0, n_splits != 1, fill, cand_count_row, cand_idx_row);
                    } else if constexpr (tm_is_rank(MODE)) {
                        if (!partial) chunk_rank<false>(acc[c & 1], nb_saddr, col0 + c * 32, n_rows, r1, r2);
                        else chunk_rank<true>(acc[c & 1], nb_saddr, col0 + c * 32, n_rows, r1, r2);
                    } else if constexpr (MODE == TM_I8P) {
                        // (key_mul - 640 = -128 from the kernel parameter: stays an IMAD on the FMA pipe)
                        if (!partial) chunk_top2_packed<false>(acc[c & 1], nb_saddr, key_mul - 640u, (key_mul - 640u) << 16, col0 + c * 32, n_rows, m1, m2);
                        else chunk_top2_packed<true>(acc[c & 1], nb_saddr, key_mul - 640u, (key_mul - 640u) << 16, col0 + c * 32, n_rows, m1, m2);
                    } else {
                        if (!partial) chunk_top2<false, MODE>(acc[c & 1], nb_saddr, cq, key_mul, c * 32, col0 + c * 32, n_rows, m1, m2);
                        else chunk_top2<true, MODE>(acc[c & 1], nb_saddr, cq, key_mul, c * 32, col0 + c * 32, n_rows, m1, m2);
                    }
                    if (c < NCH - 1) tc_wait_ld(acc[(c + 1) & 1]);
                    if (c == (NCH > 1 ? NCH - 2 : 0)) {  // the last TMEM read of this tile has landed: the MMA that reuses the stage may start
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&sm.acc_empty[a]);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.nb_empty[a]);
                // merge the tile's two best into the running pair (ascending tiles = arrival order)
                if constexpr (!tm_is_rank(MODE) && !tm_is_collect(MODE)) {
                    const int tbase = (int)(t0 + col0);
                    if constexpr (MODE == TM_I8P) {  // four 16-bit lane winners -> the tile's two smallest (hamming, column)
                        uint32_t q1, q2;
                        unpack_top2_u16x2(m1, m2, __float_as_uint(cq) - i8_bias, q1, q2);
                        if (q1 != 0xFFFFFFFFu) best.offer(q1 >> 9, tbase + (int)(q1 & 511u));
                        if (q2 != 0xFFFFFFFFu) best.offer(q2 >> 9, tbase + (int)(q2 & 511u));
                    } else {
                        // i8: the key carries popc(t) - 2 q.t + I8_BIAS; popc(q) (the bits of cq) completes the Hamming distance.
                        // exact float modes: the key carries d^2 - |q|^2 + 2^20 (float_nbexact_kernel); |q|^2 is an integer <= 2^20
                        const uint32_t dadd = MODE == TM_I8 ? __float_as_uint(cq) - I8_BIAS : static_cast<uint32_t>(cq) - 1048576u;
                        if (m1 != 0xFFFFFFFFu) best.offer((m1 >> 9) + dadd, tbase + (int)(m1 & 511u));
                        if (m2 != 0xFFFFFFFFu) best.offer((m2 >> 9) + dadd, tbase + (int)(m2 & 511u));
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
                finish_rows<MODE, 2>(sm.merge[it & 1], best, half, row, qrow, nq, reverse, split, knn_off, col_off, knn, colmin, qrow);
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

// ------------------------------------------------------------------ pass 3: exact refinement of the candidates
// Four lanes per query row of the launch.  The candidates (or, if the list overflowed, the whole train
// image) are evaluated with exactly float_exact_knn2_kernel's arithmetic -- acc = fmaf(a-b, a-b, acc)
// over ascending k, then sqrtf -- so the result is bit-identical to SFMM_FLOAT_EXACT; top-2 by the
// same (float bits << 32 | index) key.  The entry goes to split 0; the other splits are neutralised.
__global__ void float_refine_kernel(const float* __restrict__ blob, int kq, const PairDesc* __restrict__ pairs, uint32_t n_pairs,
                                    const uint32_t* __restrict__ pair_of_row, const uint32_t* __restrict__ cand_count,
                                    const uint32_t* __restrict__ cand_idx, KnnEntry* __restrict__ knn, uint32_t total_rows) {
    // four lanes per query row (a list holds ~3 candidates): 8 rows per warp, one candidate per lane and step
    const uint32_t lane = threadIdx.x & 31, sub = lane & 3;
    const uint32_t grow_raw = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 8 + (lane >> 2);  // row index in the launch
    const bool live = grow_raw < total_rows;
    const uint32_t grow = live ? grow_raw : total_rows - 1;
    const PairDesc pd = pairs[pair_of_row[grow]];
    const uint32_t q = grow - pd.q_off;
    const uint32_t c0 = cand_count[2 * (size_t)grow], c1 = cand_count[2 * (size_t)grow + 1];  // the two epilogue groups' halves
    const bool overflow = c0 > FT_CAND_CAP / 2 || c1 > FT_CAND_CAP / 2;
    const uint32_t n_items = live ? (overflow ? pd.nt : c0 + c1) : 0;
    const float4* qa = reinterpret_cast<const float4*>(blob) + (size_t)(pd.q_row0 + q) * kq;
    unsigned long long k1 = KEY_NONE, k2 = KEY_NONE;
    for (uint32_t i = sub; i < n_items; i += 4) {
        const uint32_t t = overflow ? i : cand_idx[(size_t)grow * FT_CAND_CAP + (i < c0 ? i : FT_CAND_CAP / 2 + (i - c0))];
        const float4* tb = reinterpret_cast<const float4*>(blob) + (size_t)(pd.t_row0 + t) * kq;
        float acc = 0.f;
#pragma unroll 8
        for (int c = 0; c < kq; ++c) {
            const float4 a = __ldg(qa + c), b = __ldg(tb + c);
            float d;
            d = a.x - b.x; acc = fmaf(d, d, acc);
            d = a.y - b.y; acc = fmaf(d, d, acc);
            d = a.z - b.z; acc = fmaf(d, d, acc);
            d = a.w - b.w; acc = fmaf(d, d, acc);
        }
        top2_insert(k1, k2, make_key(__float_as_uint(sqrtf(acc)), t));
    }
#pragma unroll
    for (int d = 1; d < 4; d <<= 1) {
        const unsigned long long o1 = __shfl_xor_sync(0xFFFFFFFFu, k1, d);
        const unsigned long long o2 = __shfl_xor_sync(0xFFFFFFFFu, k2, d);
        top2_insert(k1, k2, o1);
        k2 = min(k2, o2);
    }
    if (live)
        for (uint32_t sidx = sub; sidx < pd.n_splits; sidx += 4) {
            KnnEntry e;
            e.x = sidx == 0 ? k1 : KEY_NONE;
            e.y = sidx == 0 ? k2 : KEY_NONE;
            knn[pd.knn_off + (size_t)sidx * pd.nq + q] = e;
        }
}

}  // namespace sfmm
