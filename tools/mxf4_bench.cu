// tcgen05.mma kind::mxf4 (block-scaled FP4, E2M1 operands, UE8M0 scales) as a BINARY dot-product engine: correctness + issue rate.
//
// A descriptor bit b is stored as the E2M1 value 1.0 (nibble 0x2) or 0.0 (nibble 0x0); with every scale factor 2^0 (UE8M0 byte 0x7F)
// the instruction computes  acc[r][c] = sum_k a[r][k] * b[c][k] = popc(q_r & t_c)  exactly (products are 0 or 1, the fp32 accumulator
// holds integers <= 512), which is the contraction of the Hamming distance popc(q) + popc(t) - 2 q.t -- at twice the rate of kind::i8
// and with HALF the operand bytes (256 B instead of 512 B per 512-bit row).  This tool checks that
//   (1) a K-major 128-byte-swizzled smem tile (B) and a plain row-per-lane TMEM tile (A: 32 bytes of K = 8 columns per MMA) of packed
//       nibbles are read consistently (same element order on both sides), and that scale-factor columns filled with 0x7F7F7F7F work
//       whatever their exact layout is, by comparing all 128 x 128 results of a K = 512 product with the CPU;
//   (2) how many cycles one M=128 N=128 K=64 instruction takes back to back (the denominator of the roofline).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mxf4_bench tools/mxf4_bench.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

static constexpr int M = 128, N = 128, KBITS = 512, ROWB = KBITS / 2;  // 256 bytes of nibbles per row
static constexpr int ITERS = 2048, TRAIN = 16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(
            smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {  // K-major, 128-byte swizzle, 8-row groups 1024 bytes apart
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
// instruction descriptor of the block-scaled kinds (CUTLASS cute/arch/mma_sm100_desc.hpp, InstrDescriptorBlockScaled):
// a/b format E2M1 = 1 at bits 7 / 10, K-major both, N >> 3 at 17, scale format UE8M0 = 1 at 23, M >> 4 at 24, scale-factor ids 0, K = 64
static constexpr uint32_t IDESC_MXF4 = (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (1u << 23) | ((uint32_t)(M >> 4) << 24);

__device__ __forceinline__ void mma_mxf4_ts(uint32_t d, uint32_t a_tmem, uint64_t b_desc, uint32_t sfa, uint32_t sfb, int acc) {
    asm volatile(
        "{\n\t.reg .pred pe, p;\n\telect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], [%1], %2, %3, [%5], [%6], p;\n\t}" ::"r"(d),
        "r"(a_tmem), "l"(b_desc), "r"(IDESC_MXF4), "r"(acc), "r"(sfa), "r"(sfb)
        : "memory");
}
__device__ __forceinline__ void st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct Ctl {
    uint64_t bar;
    uint32_t tmem_base;
};

// TMEM map: [0,64) query tile A (256 bytes per row), [128,256) and [256,384) accumulators, [384,400) scale-factor words
// a_rows / b_rows: M x 256 and N x 256 bytes of nibbles (row-major).  out: M x N floats (when check != 0).
__global__ void __launch_bounds__(128, 1) mxf4_kernel(const uint8_t* a_rows, const uint8_t* b_rows, float* out, long long* cycles, int check) {
    extern __shared__ unsigned char raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    unsigned char* sB = base;  // 2 K-blocks x (128 rows x 128 B)
    Ctl& ctl = *reinterpret_cast<Ctl*>(sB + 2 * N * 128);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, row = threadIdx.x;
    // B: row r, byte b -> K-block b / 128, 16-byte chunk (b % 128) / 16 XOR (r % 8) inside the row's 128-byte line
    for (uint32_t i = threadIdx.x; i < N * ROWB / 16; i += blockDim.x) {
        const uint32_t r = i / (ROWB / 16), ch = i % (ROWB / 16), kb = ch / 8, c = ch % 8;
        const uint4 v = *reinterpret_cast<const uint4*>(b_rows + (size_t)r * ROWB + ch * 16);
        *reinterpret_cast<uint4*>(sB + kb * (N * 128) + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)) = v;
    }
    if (threadIdx.x == 0) {
        mbar_init(&ctl.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&ctl.tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = ctl.tmem_base;
    {
        const uint32_t taddr = tb + ((warp * 32) << 16);
        uint32_t r[32];
        for (int h = 0; h < 2; ++h) {  // the row's 256 bytes, in order, into columns [0,64)
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = reinterpret_cast<const uint32_t*>(a_rows + (size_t)row * ROWB)[h * 32 + i];
            st32(taddr + h * 32, r);
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = 0x7F7F7F7Fu;  // UE8M0 1.0 in every byte of every lane: any scale-factor layout reads 2^0
        st32(taddr + 384, r);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const uint32_t sfa = tb + 384, sfb = tb + 392;
    const uint64_t b_desc = desc_sw128(smem_u32(sB));
    if (check) {
        if (warp == 0) {
#pragma unroll
            for (int i = 0; i < KBITS / 64; ++i)  // 64 elements = 32 bytes of K per instruction: 8 TMEM columns of A, 32 bytes inside B's swizzle line
                mma_mxf4_ts(tb + 128, tb + i * 8, b_desc + (((i >> 2) * (N * 128) + (i & 3) * 32) >> 4), sfa, sfb, i > 0);
            asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(
                             smem_u32(&ctl.bar))
                         : "memory");
        }
        mbar_wait(&ctl.bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int c = 0; c < 4; ++c) {
            uint32_t r[32];
            ld32(tb + ((warp * 32) << 16) + 128 + c * 32, r);
            for (int i = 0; i < 32; ++i) out[(size_t)row * N + c * 32 + i] = __uint_as_float(r[i]);
        }
    } else if (warp == 0) {
        const long long t0 = clock64();
#pragma unroll 1
        for (int it = 0; it < ITERS; ++it) {
            const uint32_t d = tb + 128 + (it & 1) * 128;
#pragma unroll
            for (int i = 0; i < TRAIN; ++i) {
                const int k = i & 7;
                mma_mxf4_ts(d, tb + k * 8, b_desc + (((k >> 2) * (N * 128) + (k & 3) * 32) >> 4), sfa, sfb, i > 0);
            }
            asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(
                             smem_u32(&ctl.bar))
                         : "memory");
            if (it >= 1) mbar_wait(&ctl.bar, (it - 1) & 1);
        }
        mbar_wait(&ctl.bar, (ITERS - 1) & 1);
        const long long t1 = clock64();
        if (lane == 0) cycles[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
    }
}

int main() {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) {
        fprintf(stderr, "no CUDA device\n");
        return 1;
    }
    const int sms = p.multiProcessorCount;
    // random bits -> nibbles (element 2j in the low nibble of byte j)
    std::vector<uint8_t> abits(M * KBITS), bbits(N * KBITS), an(M * ROWB), bn(N * ROWB);
    uint32_t s = 12345u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (s >> 16) & 1u; };
    for (auto& v : abits) v = rnd();
    for (auto& v : bbits) v = rnd();
    for (int r = 0; r < M; ++r)
        for (int j = 0; j < ROWB; ++j) an[r * ROWB + j] = (abits[r * KBITS + 2 * j] ? 0x2 : 0) | (abits[r * KBITS + 2 * j + 1] ? 0x20 : 0);
    for (int r = 0; r < N; ++r)
        for (int j = 0; j < ROWB; ++j) bn[r * ROWB + j] = (bbits[r * KBITS + 2 * j] ? 0x2 : 0) | (bbits[r * KBITS + 2 * j + 1] ? 0x20 : 0);
    uint8_t *da, *db;
    float* dout;
    long long* dcyc;
    cudaMalloc(&da, an.size());
    cudaMalloc(&db, bn.size());
    cudaMalloc(&dout, sizeof(float) * M * N);
    cudaMalloc(&dcyc, sizeof(long long) * sms);
    cudaMemcpy(da, an.data(), an.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(db, bn.data(), bn.size(), cudaMemcpyHostToDevice);
    const size_t smem = 1024 + 2 * N * 128 + 64;
    cudaFuncSetAttribute(mxf4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    mxf4_kernel<<<1, 128, smem>>>(da, db, dout, dcyc, 1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) {
        printf("check launch failed: %s\n", cudaGetErrorString(err));
        return 2;
    }
    std::vector<float> out(M * N);
    cudaMemcpy(out.data(), dout, sizeof(float) * M * N, cudaMemcpyDeviceToHost);
    long long bad = 0;
    double worst = 0;
    for (int r = 0; r < M; ++r)
        for (int c = 0; c < N; ++c) {
            int ref = 0;
            for (int k = 0; k < KBITS; ++k) ref += abits[r * KBITS + k] & bbits[c * KBITS + k];
            const double d = out[r * N + c] - ref;
            if (d != 0) {
                if (bad < 5) printf("  mismatch r=%d c=%d got=%g ref=%d\n", r, c, out[r * N + c], ref);
                ++bad;
                worst = d > worst ? d : (-d > worst ? -d : worst);
            }
        }
    printf("check: %lld of %d results differ from popc(q & t) (worst %g)  -> %s\n", bad, M * N, worst, bad ? "FAIL" : "exact");

    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    mxf4_kernel<<<sms, 128, smem>>>(da, db, dout, dcyc, 0);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        mxf4_kernel<<<sms, 128, smem>>>(da, db, dout, dcyc, 0);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    err = cudaDeviceSynchronize();
    std::vector<long long> h(sms);
    cudaMemcpy(h.data(), dcyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (auto c : h) mx = c > mx ? c : mx;
    const double n_mma = (double)ITERS * TRAIN;
    const double tops = 2.0 * M * N * 64 * n_mma * sms / (best * 1e-3) / 1e12;
    printf("kind::mxf4 A=TMEM (TS) N=128 K=64: cycles/MMA=%.2f ms=%.3f eff_clock_GHz=%.3f chip_TOP/s=%.1f %s\n", (double)mx / n_mma, best, mx / (best * 1e6), tops,
           err == cudaSuccess ? "" : cudaGetErrorString(err));
    printf("{\"mxf4_tops\": %.1f, \"mxf4_ts_n128_cycles_per_mma\": %.2f, \"check\": \"%s\"}\n", tops, (double)mx / n_mma, bad ? "FAIL" : "exact");
    return bad ? 3 : 0;
}
