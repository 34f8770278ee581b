// Register-only issue-rate microbenchmark for the pipes the Hamming kernel lives on.
// SURVEY.md section 8(d) asks for the POPC roofline denominator to be MEASURED on the box
// (nominal: 16 POPC/clk/SM).  Prints ops/clk/SM for POPC, LOP3, IADD3, VIMNMX, IMAD and for
// POPC+LOP3 issued together (are the pipes independent?).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_bench tools/pipe_bench.cu
//   gpurun -- ./tools/pipe_bench > gpurun_out/pipe_bench.txt
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int CHAINS = 8;
constexpr int ITERS = 4096;

enum Op { POPC, LOP3, IADD3, VIMNMX, IMAD, POPC_LOP3, POPC_LOP3_4 };

template <Op OP>
__global__ void __launch_bounds__(1024) pipe_kernel(unsigned* out, long long* cycles, unsigned seed) {
    unsigned x[CHAINS], y[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
        x[i] = seed * (threadIdx.x + 1) + i * 0x9E3779B9u;
        y[i] = x[i] ^ 0xA5A5A5A5u;
    }
    const unsigned a = seed | 1u, b = seed ^ 0x55555555u;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (OP == POPC) asm volatile("popc.b32 %0, %0;" : "+r"(x[i]));
            if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(a), "r"(b));
            if (OP == IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(a));
            if (OP == VIMNMX) asm volatile("min.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i]));
            if (OP == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
            if (OP == POPC_LOP3) {
                asm volatile("popc.b32 %0, %0;" : "+r"(x[i]));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(a), "r"(b));
            }
            if (OP == POPC_LOP3_4) {  // 1 POPC : 4 LOP3, the kernel's mix
                asm volatile("popc.b32 %0, %0;" : "+r"(x[i]));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(a), "r"(b));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0xe8;" : "+r"(y[i]) : "r"(a), "r"(b));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[(i + 1) % CHAINS]) : "r"(b), "r"(a));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0xe8;" : "+r"(y[(i + 3) % CHAINS]) : "r"(b), "r"(a));
            }
        }
    }
    const long long t1 = clock64();
    unsigned acc = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) acc ^= x[i] ^ y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <Op OP>
void run(const char* name, int sms, int threads, double ops_per_iter_per_chain) {
    unsigned* out;
    long long* cyc;
    cudaMalloc(&out, sizeof(unsigned) * sms * threads);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    pipe_kernel<OP><<<sms, threads>>>(out, cyc, 12345u);  // warm-up
    cudaEventRecord(e0);
    pipe_kernel<OP><<<sms, threads>>>(out, cyc, 12345u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(sms);
    cudaMemcpy(h.data(), cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (auto c : h) mx = c > mx ? c : mx;
    const double ops_per_sm = (double)threads * CHAINS * ITERS * ops_per_iter_per_chain;
    printf("%-12s threads/SM=%4d  ops/clk/SM=%7.2f  cycles=%lld  ms=%.3f  eff_clock_GHz=%.3f  chip_Gops/s=%.1f\n", name,
           threads, ops_per_sm / (double)mx, mx, ms, mx / (ms * 1e6), ops_per_sm * sms / (ms * 1e6));
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) {
        fprintf(stderr, "no CUDA device\n");
        return 1;
    }
    printf("device=%s sms=%d clock_khz=%d\n", p.name, p.multiProcessorCount, p.clockRate);
    const int sms = p.multiProcessorCount;
    for (int threads : {256, 512, 1024}) {
        run<POPC>("POPC", sms, threads, 1);
        run<LOP3>("LOP3", sms, threads, 1);
        run<IADD3>("IADD", sms, threads, 1);
        run<VIMNMX>("VIMNMX", sms, threads, 1);
        run<IMAD>("IMAD", sms, threads, 1);
        run<POPC_LOP3>("POPC+LOP3", sms, threads, 2);
        run<POPC_LOP3_4>("POPC+4LOP3", sms, threads, 5);
    }
    return 0;
}
