// tcgen05.ld throughput with and without concurrent tcgen05.mma traffic: is the tensor-memory read path what the 2-NN kernels share?
//
// One CTA per SM.  Warps 4..4+LW-1 (LW = 4 or 8: one or two per TMEM lane quarter) read a 128-column fp32 accumulator region with
// tcgen05.ld.32x32b.x32 (4 loads of 32 columns = the epilogue's chunking), back to back; warp 0 optionally issues a continuous train
// of kind::mxf4 MMAs (M=128, N=128, A from TMEM) into ANOTHER 128-column accumulator -- no data dependence between the two, only the
// shared hardware.  Reported: bytes per clock and SM the loads sustain, cycles per 128 x 128 x 4 B "tile read", cycles per MMA.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_bench tools/tmem_bench.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

static constexpr int N = 128, M = 128;
static constexpr int LD_ITERS = 4096;    // tile reads per loader warp
static constexpr int MMA_TRAINS = 4096;  // trains of 16 MMAs

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
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
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
}

struct Ctl {
    uint64_t bar;
    uint32_t tmem_base;
};

// TMEM map: [0,64) A tile, [128,256) accumulator the MMAs write, [256,384) region the loaders read, [384,416) scale factors
template <int LW>
__global__ void __launch_bounds__(128 + 32 * LW, 1) tmem_kernel(long long* ld_cycles, long long* mma_cycles, uint32_t* sink, int with_ld, int with_mma) {
    extern __shared__ unsigned char raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    unsigned char* sB = base;  // 2 K-blocks x (128 rows x 128 B)
    Ctl& ctl = *reinterpret_cast<Ctl*>(sB + 2 * N * 128);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t i = threadIdx.x; i < 2 * N * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sB)[i] = ((i * 2654435761u) >> 7) & 0x22222222u;
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
    if (warp < 4) {  // A tile, scale factors and the read region: some bit pattern
        const uint32_t taddr = tb + ((warp * 32) << 16);
        uint32_t r[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = ((threadIdx.x * 32 + i) * 2246822519u) & 0x22222222u;
        st32(taddr, r);
        st32(taddr + 32, r);
        for (int c = 256; c < 384; c += 32) st32(taddr + c, r);
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = 0x7F7F7F7Fu;
        st32(taddr + 384, r);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 0 && with_mma) {
        const uint64_t b_desc = desc_sw128(smem_u32(sB));
        const uint32_t sfa = tb + 384, sfb = tb + 392;
        const long long t0 = clock64();
#pragma unroll 1
        for (int it = 0; it < MMA_TRAINS; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int k = i & 7;
                mma_mxf4_ts(tb + 128, tb + k * 8, b_desc + (((k >> 2) * (N * 128) + (k & 3) * 32) >> 4), sfa, sfb, i > 0);
            }
            asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(
                             smem_u32(&ctl.bar))
                         : "memory");
            if (it >= 1) mbar_wait(&ctl.bar, (it - 1) & 1);
        }
        mbar_wait(&ctl.bar, (MMA_TRAINS - 1) & 1);
        const long long t1 = clock64();
        if (lane == 0) mma_cycles[blockIdx.x] = t1 - t0;
    } else if (warp >= 4 && with_ld) {
        const uint32_t q = (warp - 4) & 3;  // warp % 4 == the lane quarter it may access (warps 4..7, 8..11)
        const uint32_t taddr = tb + ((q * 32) << 16) + 256;
        uint32_t acc[2][32], x = 0;
        const long long t0 = clock64();
        ld32(taddr, acc[0]);
#pragma unroll 1
        for (int it = 0; it < LD_ITERS; ++it) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {  // the epilogue's register double buffer: issue the next load, then consume this one
                ld32(taddr + ((c + 1) & 3) * 32, acc[(c + 1) & 1]);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int i = 0; i < 32; i += 8) x ^= acc[c & 1][i];
            }
        }
        const long long t1 = clock64();
        if (lane == 0) ld_cycles[blockIdx.x * LW + (warp - 4)] = t1 - t0;
        sink[blockIdx.x * blockDim.x + threadIdx.x] = x;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
    }
}

template <int LW>
void run(int sms, int with_ld, int with_mma) {
    long long *d_ld, *d_mma;
    uint32_t* d_sink;
    const int threads = 128 + 32 * LW;
    cudaMalloc(&d_ld, sizeof(long long) * sms * LW);
    cudaMalloc(&d_mma, sizeof(long long) * sms);
    cudaMalloc(&d_sink, sizeof(uint32_t) * sms * threads);
    cudaMemset(d_ld, 0, sizeof(long long) * sms * LW);
    cudaMemset(d_mma, 0, sizeof(long long) * sms);
    const size_t smem = 1024 + 2 * N * 128 + 64;
    cudaFuncSetAttribute(tmem_kernel<LW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    tmem_kernel<LW><<<sms, threads, smem>>>(d_ld, d_mma, d_sink, with_ld, with_mma);
    tmem_kernel<LW><<<sms, threads, smem>>>(d_ld, d_mma, d_sink, with_ld, with_mma);
    const cudaError_t err = cudaDeviceSynchronize();
    std::vector<long long> ld(sms * LW), mm(sms);
    cudaMemcpy(ld.data(), d_ld, sizeof(long long) * sms * LW, cudaMemcpyDeviceToHost);
    cudaMemcpy(mm.data(), d_mma, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    long long ld_max = 0, mm_max = 0;
    for (auto c : ld) ld_max = c > ld_max ? c : ld_max;
    for (auto c : mm) mm_max = c > mm_max ? c : mm_max;
    printf("loader warps %d  ld %d  mma %d : ", LW, with_ld, with_mma);
    if (with_ld) {
        const double per_tile = (double)ld_max / LD_ITERS;  // one loader warp reads its 32 lanes x 128 columns per iteration
        // all LW warps read concurrently: LW x 16 KB per iteration
        printf("cycles per 32x128 fp32 read %.1f  -> %.1f B/clk/SM (%.0f cycles per 128 x 128 tile at this rate)  ", per_tile, LW * 16384.0 / per_tile,
               65536.0 / (LW * 16384.0 / per_tile));
    }
    if (with_mma) printf("cycles per MMA %.2f", (double)mm_max / (MMA_TRAINS * 16.0));
    printf("  %s\n", err == cudaSuccess ? "" : cudaGetErrorString(err));
    cudaFree(d_ld);
    cudaFree(d_mma);
    cudaFree(d_sink);
}

int main() {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) {
        fprintf(stderr, "no CUDA device\n");
        return 1;
    }
    const int sms = p.multiProcessorCount;
    printf("device=%s sms=%d\n", p.name, sms);
    run<4>(sms, 1, 0);
    run<8>(sms, 1, 0);
    run<4>(sms, 0, 1);
    run<4>(sms, 1, 1);
    run<8>(sms, 1, 1);
    return 0;
}
