// tcgen05.mma issue-rate microbenchmark: the tensor-pipe denominators of bench.py's roofline, MEASURED on the box
// instead of inferred from cuBLAS bf16.  One CTA per SM; one warp issues a long train of
// back-to-back tcgen05.mma (cta_group::1, M=128, N=128 or 256, 32 bytes of K per instruction) on operands that are
// already in shared / tensor memory -- no loads, no epilogue: the rate the pipe can sustain for this instruction shape.
//   kinds : i8 (u8 x u8 -> s32, K=32), f16 (f16 x f16 -> f32, K=16), tf32 (K=8)
//   A from: shared memory (SS, the float_tensor.cuh kernel) or tensor memory (TS, float_tensor_ts.cuh)
// Operands hold a fixed pseudo-random bit pattern (not zeros: data toggling affects power and therefore clocks).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tc_bench tools/tc_bench.cu
//   gpurun -- './tools/tc_bench > gpurun_out/tc_bench.txt'        # last line is JSON for profiles/tcgen05_peaks_r02.json
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

static constexpr int M = 128;
static constexpr int ITERS = 2048;   // trains of MMAs per launch
static constexpr int TRAIN = 16;     // MMAs between commits

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
// K-major operand, 128-byte swizzle (same encoding as csrc/float_tensor.cuh: umma_desc_sw128)
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

enum Kind { K_I8 = 0, K_F16 = 1, K_TF32 = 2 };

template <int KIND, bool TS>
__device__ __forceinline__ void mma(uint32_t d, uint32_t a_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, int acc) {
    if constexpr (TS) {
        if constexpr (KIND == K_I8)
            asm volatile("{\n\t.reg .pred pe, p;\n\telect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t@pe tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
        else if constexpr (KIND == K_F16)
            asm volatile("{\n\t.reg .pred pe, p;\n\telect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
        else
            asm volatile("{\n\t.reg .pred pe, p;\n\telect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
    } else {
        if constexpr (KIND == K_I8)
            asm volatile("{\n\t.reg .pred pe, p;\n\telect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t@pe tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
        else if constexpr (KIND == K_F16)
            asm volatile("{\n\t.reg .pred pe, p;\n\telect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
        else
            asm volatile("{\n\t.reg .pred pe, p;\n\telect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
    }
}

struct Ctl {
    uint64_t bar;
    uint32_t tmem_base;
};

// smem: [A tile 128 x 128 B = 16 KB][B tile N x 128 B][Ctl]; 4 K-steps of 32 bytes inside the 128-byte swizzle row
template <int KIND, bool TS, int N>
__global__ void __launch_bounds__(128, 1) tc_kernel(long long* cycles, uint32_t seed) {
    extern __shared__ unsigned char raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    unsigned char* sA = base;
    unsigned char* sB = base + M * 128;
    Ctl& ctl = *reinterpret_cast<Ctl*>(sB + N * 128);
    const uint32_t warp = threadIdx.x >> 5;
    // fill the operands with a bit pattern whose values are small finite numbers in every interpretation
    for (uint32_t i = threadIdx.x; i < (M + N) * 128 / 4; i += blockDim.x) {
        uint32_t x = (i + 1) * 2654435761u ^ seed;
        uint32_t v;
        if (KIND == K_I8) v = x & 0x01010101u;                                   // {0,1} bytes, like the unpacked descriptors
        else if (KIND == K_F16) v = 0x3C003C00u | (x & 0x03FF03FFu);             // halves in [1,2)
        else v = 0x3F800000u | (x & 0x007FE000u);                                // tf32 in [1,2)
        reinterpret_cast<uint32_t*>(base)[i] = v;
    }
    if (threadIdx.x == 0) {
        mbar_init(&ctl.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core's async proxy
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&ctl.tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = ctl.tmem_base;
    if (TS) {  // query tile into TMEM columns [0,32): row r = lane r, 128 bytes = 32 columns
        uint32_t r[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = reinterpret_cast<uint32_t*>(sA)[threadIdx.x * 32 + i];
        const uint32_t taddr = tb + ((warp * 32) << 16);
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
            "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
            "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
            "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
            "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
            "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
            "r"(r[31])
            : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    long long t0 = 0, t1 = 0;
    if (warp == 0) {
        const uint64_t a_desc = desc_sw128(smem_u32(sA)), b_desc = desc_sw128(smem_u32(sB));
        const uint32_t dfmt = KIND == K_I8 ? 2u : 1u, abfmt = KIND == K_TF32 ? 2u : 0u;
        const uint32_t idesc = (dfmt << 4) | (abfmt << 7) | (abfmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t d0 = tb + 128, d1 = tb + 128 + (N == 256 ? 0 : 128);  // two accumulators when they fit next to the A tile
        t0 = clock64();
#pragma unroll 1
        for (int it = 0; it < ITERS; ++it) {
            const uint32_t d = (it & 1) ? d1 : d0;
#pragma unroll
            for (int i = 0; i < TRAIN; ++i) {
                const int k = i & 3;  // 32 bytes of K inside the swizzle row: +2 in the descriptor's 16-byte units, 8 TMEM columns
                mma<KIND, TS>(d, tb + k * 8, a_desc + k * 2, b_desc + k * 2, idesc, i > 0);
            }
            asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(
                             smem_u32(&ctl.bar))
                         : "memory");
            if (it >= 1) mbar_wait(&ctl.bar, (it - 1) & 1);  // one train (1024+ cycles of tensor work) always queued behind the one we wait for
        }
        mbar_wait(&ctl.bar, (ITERS - 1) & 1);
        t1 = clock64();
        if ((threadIdx.x & 31) == 0) cycles[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
    }
}

struct Result {
    double tops, cyc_per_mma, ghz;
};

template <int KIND, bool TS, int N>
Result run(const char* name, int sms) {
    long long* cyc;
    cudaMalloc(&cyc, sizeof(long long) * sms);
    const size_t smem = 1024 + (M + N) * 128 + 64;
    auto kern = tc_kernel<KIND, TS, N>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    kern<<<sms, 128, smem>>>(cyc, 1u);  // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        kern<<<sms, 128, smem>>>(cyc, 7u + rep);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    const cudaError_t err = cudaDeviceSynchronize();
    std::vector<long long> h(sms);
    cudaMemcpy(h.data(), cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (auto c : h) mx = c > mx ? c : mx;
    const int k_per = KIND == K_I8 ? 32 : (KIND == K_F16 ? 16 : 8);
    const double n_mma = (double)ITERS * TRAIN;
    const double ops = 2.0 * M * N * k_per * n_mma * sms;
    Result r{ops / (best * 1e-3) / 1e12, (double)mx / n_mma, mx / (best * 1e6)};
    printf("%-28s N=%3d  cycles/MMA=%7.2f  ms=%.3f  eff_clock_GHz=%.3f  chip_T%s/s=%8.1f  %s\n", name, N, r.cyc_per_mma, best, r.ghz,
           KIND == K_I8 ? "OP" : "FLOP", r.tops, err == cudaSuccess ? "" : cudaGetErrorString(err));
    cudaFree(cyc);
    return r;
}

int main() {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) {
        fprintf(stderr, "no CUDA device\n");
        return 1;
    }
    const int sms = p.multiProcessorCount;
    printf("device=%s sms=%d clock_khz=%d  (kernel timed with CUDA events incl. ~10 us of set-up; cycles = clock64 around the MMA train)\n", p.name, sms, p.clockRate);
    const Result i8_ts = run<K_I8, true, 128>("kind::i8   A=TMEM  (TS)", sms);
    const Result i8_ss = run<K_I8, false, 128>("kind::i8   A=smem  (SS)", sms);
    const Result i8_ss256 = run<K_I8, false, 256>("kind::i8   A=smem  (SS)", sms);
    const Result f16_ts = run<K_F16, true, 128>("kind::f16  A=TMEM  (TS)", sms);
    const Result f16_ss = run<K_F16, false, 128>("kind::f16  A=smem  (SS)", sms);
    const Result f16_ss256 = run<K_F16, false, 256>("kind::f16  A=smem  (SS)", sms);
    const Result tf_ts = run<K_TF32, true, 128>("kind::tf32 A=TMEM  (TS)", sms);
    const Result tf_ss = run<K_TF32, false, 128>("kind::tf32 A=smem  (SS)", sms);
    const Result tf_ss256 = run<K_TF32, false, 256>("kind::tf32 A=smem  (SS)", sms);
    auto mx = [](double a, double b, double c) { return a > b ? (a > c ? a : c) : (b > c ? b : c); };
    printf("{\"i8_tops\": %.1f, \"f16_tflops\": %.1f, \"tf32_tflops\": %.1f, \"i8_ts_n128_cycles_per_mma\": %.2f, \"f16_ts_n128_cycles_per_mma\": %.2f, "
           "\"tf32_ss_n128_cycles_per_mma\": %.2f, \"how\": \"tools/tc_bench.cu: back-to-back tcgen05.mma cta_group::1 M=128, best of N=128/256 and "
           "A in smem/TMEM, operands resident, no loads, no epilogue; CUDA-event time of the whole launch on all SMs\"}\n",
           mx(i8_ts.tops, i8_ss.tops, i8_ss256.tops), mx(f16_ts.tops, f16_ss.tops, f16_ss256.tops), mx(tf_ts.tops, tf_ss.tops, tf_ss256.tops),
           i8_ts.cyc_per_mma, f16_ts.cyc_per_mma, tf_ss.cyc_per_mma);
    return 0;
}
