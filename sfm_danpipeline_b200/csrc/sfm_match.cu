// libsfmmatch.so -- C ABI (include/sfm_match.h) + host-side pair scheduler for the sm_100a
// descriptor-matching kernels.  There is no CPU path in this file: if a CUDA device is
// missing every entry point fails with SFMM_ENODEVICE.
//
// Reference path being replaced: StructFromMotion::getMatching (/root/reference/src/Sfm.cpp:590-608)
// driven by findBestPair's q<t loop (:511-515).  The scheduler turns a list of image pairs into
//   1. "knn tiles"   (query-row tile x train-row range)  -> binary_knn2_kernel / float kernels
//   2. "filter tiles" (1024 query rows)                  -> ratio test + cross-check + compaction
// and moves the compacted cv::DMatch-layout records to host memory.
#include <algorithm>
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/sfm_match.h"
#include "binary_knn.cuh"
#include "common.cuh"
#include "filter.cuh"
#include "float_exact.cuh"
#include "float_tensor.cuh"

using namespace sfmm;

static_assert(sizeof(SfmDMatch) == 16, "SfmDMatch must be layout-identical to cv::DMatch");

namespace {

thread_local std::string g_create_error;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            want = bytes;
            e = cudaMalloc(&p, want);
        }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T* as() const { return static_cast<T*>(p); }
};

struct HostChunk {  // one block of match records on the host (stable address)
    std::unique_ptr<SfmDMatch[]> data;
    int64_t n = 0;
};

struct PairSlot {  // where a computed pair lives
    const SfmDMatch* ptr;
    int32_t count;
};

// Binary kernel geometry (see binary_knn.cuh)
constexpr int BK_THREADS = 128;
constexpr int BK_TT = 128;
template <int W> struct BkTq { static constexpr int v = (W >= 32) ? 2 : 4; };

struct ChunkPlan {
    std::vector<PairDesc> pairs;
    std::vector<KnnTile> tiles;
    std::vector<FilterTile> ftiles;
    uint64_t knn_entries = 0;
    uint64_t col_entries = 0;
    uint64_t max_matches = 0;
    double work = 0;  // algorithmic POPC32 ops / FLOPs of the knn launch
};

}  // namespace

struct SfmmCtx {
    SfmmConfig cfg{};
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    mutable std::string err;
    int csa_level = 2;
    size_t fx_attr_smem = 0;
    size_t ft_attr_smem = 0;

    // float path state (norms + TF32-exactness proof, see float_tensor.cuh)
    bool float_prepared = false;
    bool tensor_eligible = false;
    bool use_tensor = false;
    CUtensorMap tmap{};

    // descriptors (imagesDescriptors, include/Sfm.h:29)
    int32_t n_images = 0;
    std::vector<int32_t> rows;
    std::vector<uint32_t> row0;
    int32_t cols = 0;
    int32_t elem_type = -1;
    size_t pitch = 0;
    uint64_t total_rows = 0;
    DevBuf blob;
    size_t blob_bytes = 0;

    // per-launch scratch
    DevBuf d_pairs, d_tiles, d_ftiles, d_knn, d_colmin, d_tile_count, d_tile_off, d_pair_count, d_pair_off,
        d_matches, d_idx, d_dist, d_norms, d_flags;
    void* pinned = nullptr;
    size_t pinned_cap = 0;

    // result table
    std::vector<int32_t> res_qt;
    std::vector<int32_t> res_counts;
    std::vector<int64_t> res_offsets;  // offsets into the consolidated view
    std::vector<PairSlot> res_slots;
    std::vector<HostChunk> chunks;
    std::unordered_map<uint64_t, int64_t> index;
    mutable std::vector<SfmDMatch> flat;  // consolidated copy (built lazily when >1 chunk)
    mutable bool flat_valid = false;
    int64_t n_matches = 0;

    SfmmStats stats{};
};

namespace {

int fail(const SfmmCtx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    else g_create_error = msg;
    return code;
}

#define CU_TRY(ctx, expr)                                                                         \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            (void)cudaGetLastError();                                                             \
            return fail(ctx, _e == cudaErrorMemoryAllocation ? SFMM_ENOMEM : SFMM_ECUDA,          \
                        std::string(#expr) + ": " + cudaGetErrorString(_e));                      \
        }                                                                                         \
    } while (0)

int binary_words(int cols) {  // 32-bit words per packed row
    if (cols <= 16) return 4;
    if (cols <= 32) return 8;
    if (cols <= 64) return 16;
    if (cols <= 128) return 32;
    return 0;
}

int ensure_pinned(SfmmCtx* ctx, size_t bytes) {
    if (bytes <= ctx->pinned_cap) return SFMM_OK;
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    ctx->pinned = nullptr;
    ctx->pinned_cap = 0;
    const size_t want = bytes + bytes / 4 + 4096;
    CU_TRY(ctx, cudaMallocHost(&ctx->pinned, want));
    ctx->pinned_cap = want;
    return SFMM_OK;
}

void clear_results(SfmmCtx* ctx) {
    ctx->res_qt.clear();
    ctx->res_counts.clear();
    ctx->res_offsets.clear();
    ctx->res_slots.clear();
    ctx->chunks.clear();
    ctx->index.clear();
    ctx->flat.clear();
    ctx->flat_valid = false;
    ctx->n_matches = 0;
}

inline uint64_t pair_key(int32_t q, int32_t t) { return (static_cast<uint64_t>(static_cast<uint32_t>(q)) << 32) | static_cast<uint32_t>(t); }

// ---------------------------------------------------------------------------- planning
// Turn pairs [begin,end) of qt into device work descriptors.
int plan_chunk(SfmmCtx* ctx, const int32_t* qt, int64_t n, ChunkPlan& plan) {
    const bool is_float = ctx->elem_type == SFMM_F32;
    const int W = is_float ? 0 : binary_words(ctx->cols);
    const int q_tile = is_float ? (ctx->use_tensor ? FT_M : FX_BQ) : BK_THREADS * (W >= 32 ? 2 : 4);
    const int t_gran = is_float ? (ctx->use_tensor ? FT_N : FX_BT) : BK_TT;
    plan.pairs.resize(n);
    uint64_t base_tiles = 0;
    for (int64_t i = 0; i < n; ++i) {
        const int32_t q = qt[2 * i], t = qt[2 * i + 1];
        if (q < 0 || q >= ctx->n_images || t < 0 || t >= ctx->n_images)
            return fail(ctx, SFMM_ERANGE, "image index out of range in pair list");
        if (ctx->rows[q] > 0 && ctx->rows[t] >= 2) base_tiles += (ctx->rows[q] + q_tile - 1) / q_tile;
    }
    // Small jobs (a single getMatching call, the temple set) do not fill 148 SMs with one tile
    // per query-row tile: split the train range so that at least ~2 waves of CTAs exist.
    const uint64_t want_tiles = static_cast<uint64_t>(ctx->sm_count) * 8;
    uint32_t splits_wanted = 1;
    if (base_tiles > 0 && base_tiles < want_tiles)
        splits_wanted = static_cast<uint32_t>(std::min<uint64_t>(32, (want_tiles + base_tiles - 1) / base_tiles));

    const double work_per_eval = is_float ? 2.0 * ctx->cols : static_cast<double>((ctx->cols * 8 + 31) / 32);
    for (int64_t i = 0; i < n; ++i) {
        const int32_t q = qt[2 * i], t = qt[2 * i + 1];
        PairDesc& pd = plan.pairs[i];
        pd.q_row0 = ctx->row0[q];
        pd.nq = static_cast<uint32_t>(ctx->rows[q]);
        pd.t_row0 = ctx->row0[t];
        pd.nt = static_cast<uint32_t>(ctx->rows[t]);
        pd.knn_off = plan.knn_entries;
        pd.col_off = plan.col_entries;
        pd.n_splits = 0;
        pd.first_ftile = static_cast<uint32_t>(plan.ftiles.size());
        pd.n_ftiles = 0;
        pd.pad = 0;
        if (pd.nq == 0 || pd.nt < 2) continue;  // defined: no matches (see sfm_match.h)
        uint32_t splits = std::min<uint32_t>(splits_wanted, std::max<uint32_t>(1, pd.nt / (2 * t_gran)));
        const uint32_t per = ((pd.nt + splits - 1) / splits + t_gran - 1) / t_gran * t_gran;
        splits = (pd.nt + per - 1) / per;
        pd.n_splits = splits;
        for (uint32_t q0 = 0; q0 < pd.nq; q0 += q_tile)
            for (uint32_t s = 0; s < splits; ++s) {
                KnnTile kt;
                kt.pair = static_cast<uint32_t>(i);
                kt.q0 = q0;
                kt.t0 = s * per;
                kt.t1 = std::min(pd.nt, (s + 1) * per);
                kt.split = s;
                plan.tiles.push_back(kt);
            }
        pd.n_ftiles = (pd.nq + FILTER_TILE - 1) / FILTER_TILE;
        for (uint32_t f = 0; f < pd.n_ftiles; ++f) plan.ftiles.push_back(FilterTile{static_cast<uint32_t>(i), f * FILTER_TILE});
        plan.knn_entries += static_cast<uint64_t>(splits) * pd.nq;
        plan.col_entries += pd.nt;
        plan.max_matches += pd.nq;
        plan.work += static_cast<double>(pd.nq) * pd.nt * work_per_eval;
    }
    return SFMM_OK;
}

// ---------------------------------------------------------------------------- launches
template <int W, int CSA, bool CROSS>
cudaError_t launch_binary_t(SfmmCtx* ctx, uint32_t n_tiles) {
    constexpr int TQ = BkTq<W>::v;
    using Smem = BinaryKnnSmem<W, TQ, BK_THREADS, BK_TT>;
    auto kern = binary_knn2_kernel<W, TQ, BK_THREADS, BK_TT, CSA, CROSS>;
    static_assert(sizeof(Smem) <= 48 * 1024, "fits the default dynamic shared memory limit");
    kern<<<n_tiles, BK_THREADS, sizeof(Smem), ctx->stream>>>(ctx->blob.as<uint32_t>(), ctx->d_tiles.as<KnnTile>(),
                                                             ctx->d_pairs.as<PairDesc>(), ctx->d_knn.as<KnnEntry>(),
                                                             ctx->d_colmin.as<unsigned long long>(),
                                                             KeyWeights{{1u << IDX_BITS, 2u << IDX_BITS, 4u << IDX_BITS}});
    return cudaGetLastError();
}

template <int W, bool CROSS>
cudaError_t launch_binary_w(SfmmCtx* ctx, uint32_t n_tiles) {
    switch (ctx->csa_level) {
        case 0: return launch_binary_t<W, 0, CROSS>(ctx, n_tiles);
        case 1: return launch_binary_t<W, 1, CROSS>(ctx, n_tiles);
        case 3: return launch_binary_t<W, 3, CROSS>(ctx, n_tiles);
        default: return launch_binary_t<W, 2, CROSS>(ctx, n_tiles);
    }
}

template <bool CROSS>
cudaError_t launch_binary(SfmmCtx* ctx, uint32_t n_tiles) {
    switch (binary_words(ctx->cols)) {
        case 4: return launch_binary_w<4, CROSS>(ctx, n_tiles);
        case 8: return launch_binary_w<8, CROSS>(ctx, n_tiles);
        case 16: return launch_binary_w<16, CROSS>(ctx, n_tiles);
        case 32: return launch_binary_w<32, CROSS>(ctx, n_tiles);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_float_exact(SfmmCtx* ctx, uint32_t n_tiles) {
    const int kq = static_cast<int>(ctx->pitch / 16);
    const size_t smem = float_exact_smem_bytes(kq);
    if (smem > ctx->fx_attr_smem) {  // per context == per device
        cudaError_t e = cudaFuncSetAttribute(float_exact_knn2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        ctx->fx_attr_smem = smem;
    }
    float_exact_knn2_kernel<<<n_tiles, FX_THREADS, smem, ctx->stream>>>(
        ctx->blob.as<float>(), kq, ctx->d_tiles.as<KnnTile>(), ctx->d_pairs.as<PairDesc>(), ctx->d_knn.as<KnnEntry>(),
        ctx->d_colmin.as<unsigned long long>(), ctx->cfg.cross_check ? 1 : 0);
    return cudaGetLastError();
}

template <int KB>
cudaError_t launch_float_tensor_t(SfmmCtx* ctx, uint32_t n_tiles) {
    const size_t smem = float_tensor_smem_bytes(KB);
    if (smem > ctx->ft_attr_smem) {
        cudaError_t e = cudaFuncSetAttribute(float_tensor_knn2_kernel<KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        ctx->ft_attr_smem = smem;
    }
    float_tensor_knn2_kernel<KB><<<n_tiles, FT_THREADS, smem, ctx->stream>>>(
        ctx->tmap, ctx->d_norms.as<float>(), ctx->d_tiles.as<KnnTile>(), ctx->d_pairs.as<PairDesc>(), ctx->d_knn.as<KnnEntry>(), 512u);
    return cudaGetLastError();
}

cudaError_t launch_float_tensor(SfmmCtx* ctx, uint32_t n_tiles) {
    switch (ctx->cols / FT_KB_ELEMS) {
        case 1: return launch_float_tensor_t<1>(ctx, n_tiles);
        case 2: return launch_float_tensor_t<2>(ctx, n_tiles);
        case 3: return launch_float_tensor_t<3>(ctx, n_tiles);
        case 4: return launch_float_tensor_t<4>(ctx, n_tiles);
    }
    return cudaErrorInvalidValue;
}

// Float descriptors only, once per descriptor set (lazily, so that a blob filled by an NCCL
// broadcast is seen): row norms, the TF32-exactness proof and the TMA tensor map; picks the path.
int prepare_float(SfmmCtx* ctx) {
    if (ctx->elem_type != SFMM_F32 || ctx->float_prepared) return SFMM_OK;
    ctx->tensor_eligible = false;
    ctx->use_tensor = false;
    const bool shape_ok = ctx->cols % FT_KB_ELEMS == 0 && ctx->cols <= 4 * FT_KB_ELEMS && ctx->total_rows > 0;
    if (ctx->cfg.float_mode != SFMM_FLOAT_EXACT && shape_ok && !ctx->cfg.cross_check) {
        CU_TRY(ctx, ctx->d_norms.ensure((static_cast<size_t>(ctx->total_rows) + 2 * FT_N) * sizeof(float)));  // + tail for the bulk copies
        CU_TRY(ctx, ctx->d_flags.ensure(2 * sizeof(unsigned int)));
        CU_TRY(ctx, cudaMemsetAsync(ctx->d_flags.p, 0, 2 * sizeof(unsigned int), ctx->stream));
        const uint32_t rows = static_cast<uint32_t>(ctx->total_rows);
        float_prepare_kernel<<<(rows + 7) / 8, 256, 0, ctx->stream>>>(ctx->blob.as<float>(), static_cast<int>(ctx->pitch / 16), rows,
                                                                     ctx->cols, ctx->d_norms.as<float>(), ctx->d_flags.as<unsigned int>());
        CU_TRY(ctx, cudaGetLastError());
        ctx->stats.kernel_launches += 1;
        unsigned int flags[2] = {1, 0};
        CU_TRY(ctx, cudaMemcpyAsync(flags, ctx->d_flags.p, sizeof(flags), cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        float max_norm2;
        std::memcpy(&max_norm2, &flags[1], sizeof(float));
        ctx->tensor_eligible = flags[0] == 0 && max_norm2 <= 1048576.f;  // integers, |v|<=2047, |x|^2 <= 2^20
        if (ctx->tensor_eligible) {
            typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                         const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
            void* fn = nullptr;
            cudaDriverEntryPointQueryResult qres;
            CU_TRY(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
            if (!fn || qres != cudaDriverEntryPointSuccess) return fail(ctx, SFMM_ECUDA, "cuTensorMapEncodeTiled is not available in this driver");
            const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(ctx->pitch / 4), static_cast<cuuint64_t>(ctx->total_rows)};
            const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ctx->pitch)};
            const cuuint32_t box[2] = {FT_KB_ELEMS, FT_M};
            const cuuint32_t estr[2] = {1, 1};
            const CUresult r = reinterpret_cast<EncodeFn>(fn)(&ctx->tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ctx->blob.p, gdim, gstride, box, estr,
                                                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return fail(ctx, SFMM_ECUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
            ctx->use_tensor = true;
        }
    }
    if (ctx->cfg.float_mode == SFMM_FLOAT_TENSOR && !ctx->use_tensor)
        return fail(ctx, SFMM_EINVAL,
                    "SFMM_FLOAT_TENSOR needs TF32-exact descriptors (integer values |v|<=2047, row norm^2 <= 2^20), a width that is a "
                    "multiple of 32 up to 128 and cross_check off; use SFMM_FLOAT_AUTO or SFMM_FLOAT_EXACT");
    ctx->float_prepared = true;
    ctx->stats.float_path = ctx->use_tensor ? SFMM_FLOAT_TENSOR : SFMM_FLOAT_EXACT;
    return SFMM_OK;
}

// Uploads the plan and runs the 2-NN kernel; leaves merged-able partial lists in d_knn (+ d_colmin).
int run_knn(SfmmCtx* ctx, const ChunkPlan& plan, bool timed) {
    const bool cross = ctx->cfg.cross_check != 0;
    CU_TRY(ctx, ctx->d_pairs.ensure(std::max<size_t>(1, plan.pairs.size()) * sizeof(PairDesc)));
    CU_TRY(ctx, ctx->d_tiles.ensure(std::max<size_t>(1, plan.tiles.size()) * sizeof(KnnTile)));
    CU_TRY(ctx, ctx->d_ftiles.ensure(std::max<size_t>(1, plan.ftiles.size()) * sizeof(FilterTile)));
    CU_TRY(ctx, ctx->d_knn.ensure(std::max<uint64_t>(1, plan.knn_entries) * sizeof(KnnEntry)));
    CU_TRY(ctx, ctx->d_colmin.ensure(std::max<uint64_t>(1, cross ? plan.col_entries : 1) * sizeof(unsigned long long)));
    if (!plan.pairs.empty())
        CU_TRY(ctx, cudaMemcpyAsync(ctx->d_pairs.p, plan.pairs.data(), plan.pairs.size() * sizeof(PairDesc),
                                    cudaMemcpyHostToDevice, ctx->stream));
    if (!plan.tiles.empty())
        CU_TRY(ctx, cudaMemcpyAsync(ctx->d_tiles.p, plan.tiles.data(), plan.tiles.size() * sizeof(KnnTile),
                                    cudaMemcpyHostToDevice, ctx->stream));
    if (!plan.ftiles.empty())
        CU_TRY(ctx, cudaMemcpyAsync(ctx->d_ftiles.p, plan.ftiles.data(), plan.ftiles.size() * sizeof(FilterTile),
                                    cudaMemcpyHostToDevice, ctx->stream));
    ctx->stats.h2d_bytes += plan.pairs.size() * sizeof(PairDesc) + plan.tiles.size() * sizeof(KnnTile) +
                            plan.ftiles.size() * sizeof(FilterTile);
    if (cross && plan.col_entries)
        CU_TRY(ctx, cudaMemsetAsync(ctx->d_colmin.p, 0xFF, plan.col_entries * sizeof(unsigned long long), ctx->stream));
    if (plan.tiles.empty()) return SFMM_OK;
    if (timed) CU_TRY(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
    cudaError_t e;
    if (ctx->elem_type == SFMM_F32) e = ctx->use_tensor ? launch_float_tensor(ctx, static_cast<uint32_t>(plan.tiles.size()))
                                                    : launch_float_exact(ctx, static_cast<uint32_t>(plan.tiles.size()));
    else e = cross ? launch_binary<true>(ctx, static_cast<uint32_t>(plan.tiles.size()))
                   : launch_binary<false>(ctx, static_cast<uint32_t>(plan.tiles.size()));
    CU_TRY(ctx, e);
    if (timed) CU_TRY(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
    ctx->stats.kernel_launches += 1;
    return SFMM_OK;
}

template <bool IS_FLOAT, bool CROSS>
cudaError_t launch_filter(SfmmCtx* ctx, const ChunkPlan& plan, int32_t* d_counts, SfmDMatch* d_matches, uint64_t capacity) {
    const uint32_t nft = static_cast<uint32_t>(plan.ftiles.size());
    const float ratio = ctx->cfg.ratio;
    filter_count_kernel<IS_FLOAT, CROSS><<<nft, FILTER_THREADS, 0, ctx->stream>>>(
        ctx->d_ftiles.as<FilterTile>(), ctx->d_pairs.as<PairDesc>(), ctx->d_knn.as<KnnEntry>(),
        ctx->d_colmin.as<unsigned long long>(), ratio, ctx->d_tile_count.as<uint32_t>());
    tile_scan_kernel<<<1, SCAN_THREADS, 0, ctx->stream>>>(ctx->d_tile_count.as<uint32_t>(),
                                                          ctx->d_tile_off.as<unsigned long long>(), nft);
    filter_write_kernel<IS_FLOAT, CROSS><<<nft, FILTER_THREADS, 0, ctx->stream>>>(
        ctx->d_ftiles.as<FilterTile>(), ctx->d_pairs.as<PairDesc>(), ctx->d_knn.as<KnnEntry>(),
        ctx->d_colmin.as<unsigned long long>(), ratio, ctx->d_tile_off.as<unsigned long long>(), d_matches, capacity,
        d_counts, ctx->d_pair_off.as<unsigned long long>());
    return cudaGetLastError();
}

// Ratio test + cross-check + compaction of the lists left by run_knn into caller-chosen DEVICE
// buffers.  On return (stream synchronised) *total = records produced (may exceed capacity:
// nothing past capacity is written).
int run_filter(SfmmCtx* ctx, const ChunkPlan& plan, int32_t* d_counts, SfmDMatch* d_matches, uint64_t capacity,
               uint64_t* total) {
    const size_t np = plan.pairs.size();
    const size_t nft = plan.ftiles.size();
    *total = 0;
    if (np) CU_TRY(ctx, cudaMemsetAsync(d_counts, 0, np * sizeof(int32_t), ctx->stream));
    CU_TRY(ctx, ctx->d_pair_off.ensure(std::max<size_t>(1, np) * sizeof(unsigned long long)));
    if (np) CU_TRY(ctx, cudaMemsetAsync(ctx->d_pair_off.p, 0, np * sizeof(unsigned long long), ctx->stream));
    if (nft == 0) {
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        return SFMM_OK;
    }
    CU_TRY(ctx, ctx->d_tile_count.ensure(nft * sizeof(uint32_t)));
    CU_TRY(ctx, ctx->d_tile_off.ensure((nft + 1) * sizeof(unsigned long long)));
    const bool is_float = ctx->elem_type == SFMM_F32, cross = ctx->cfg.cross_check != 0;
    cudaError_t e;
    if (is_float) e = cross ? launch_filter<true, true>(ctx, plan, d_counts, d_matches, capacity)
                            : launch_filter<true, false>(ctx, plan, d_counts, d_matches, capacity);
    else e = cross ? launch_filter<false, true>(ctx, plan, d_counts, d_matches, capacity)
                   : launch_filter<false, false>(ctx, plan, d_counts, d_matches, capacity);
    CU_TRY(ctx, e);
    ctx->stats.kernel_launches += 3;
    CU_TRY(ctx, ensure_pinned(ctx, 64) == SFMM_OK ? cudaSuccess : cudaErrorMemoryAllocation);
    CU_TRY(ctx, cudaMemcpyAsync(ctx->pinned, ctx->d_tile_off.as<unsigned long long>() + nft, sizeof(unsigned long long),
                                cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *total = *static_cast<unsigned long long*>(ctx->pinned);
    ctx->stats.d2h_bytes += sizeof(unsigned long long);
    return SFMM_OK;
}

void account_knn_time(SfmmCtx* ctx, const ChunkPlan& plan) {
    if (plan.tiles.empty()) return;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]) == cudaSuccess) {
        ctx->stats.last_knn_ms += ms;
        ctx->stats.last_knn_work += plan.work;
        ctx->stats.last_knn_launches += 1;
    } else {
        (void)cudaGetLastError();
    }
}

// How many pairs of qt[from..n) fit the per-launch budget.
int64_t chunk_extent(const SfmmCtx* ctx, const int32_t* qt, int64_t from, int64_t n) {
    const uint64_t row_budget = 8u << 20;  // query rows per launch: 128 MB of 2-NN scratch, 128 MB of records
    const int64_t pair_cap = ctx->cfg.pair_batch > 0 ? ctx->cfg.pair_batch : (1 << 20);
    uint64_t rows = 0;
    int64_t i = from;
    for (; i < n && i - from < pair_cap; ++i) {
        const int32_t q = qt[2 * i];
        const uint64_t r = (q >= 0 && q < ctx->n_images) ? static_cast<uint64_t>(ctx->rows[q]) : 0;
        if (i > from && rows + r > row_budget) break;
        rows += r;
    }
    return i - from;
}

int require_descriptors(const SfmmCtx* ctx) {
    if (!ctx) return SFMM_EINVAL;
    if (ctx->elem_type < 0) return fail(ctx, SFMM_ESTATE, "sfmm_set_descriptors has not been called");
    return SFMM_OK;
}

int bind_device(const SfmmCtx* ctx) {
    CU_TRY(ctx, cudaSetDevice(ctx->cfg.device));
    return SFMM_OK;
}

}  // namespace

// =============================================================================== C ABI
extern "C" {

SFMM_API const char* sfmm_version(void) { return "0.1.0 (sm_100a)"; }

SFMM_API void sfmm_default_config(SfmmConfig* cfg) {
    if (!cfg) return;
    std::memset(cfg, 0, sizeof(*cfg));
    cfg->struct_size = static_cast<int32_t>(sizeof(SfmmConfig));
    cfg->device = 0;
    cfg->norm = SFMM_NORM_L2;  // src/Sfm.cpp:593
    cfg->ratio = 0.8f;         // include/Sfm.h:60
    cfg->cross_check = 0;      // src/Sfm.cpp:593 (crossCheck=false)
    cfg->float_mode = SFMM_FLOAT_AUTO;
    cfg->pair_batch = 0;
}

SFMM_API const char* sfmm_last_error(const SfmmCtx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

SFMM_API int sfmm_create(const SfmmConfig* cfg, SfmmCtx** out) {
    if (!cfg || !out) return fail(nullptr, SFMM_EINVAL, "sfmm_create: NULL argument");
    *out = nullptr;
    if (cfg->struct_size != static_cast<int32_t>(sizeof(SfmmConfig)))
        return fail(nullptr, SFMM_EINVAL, "sfmm_create: SfmmConfig.struct_size mismatch (use sfmm_default_config)");
    if (cfg->norm != SFMM_NORM_HAMMING && cfg->norm != SFMM_NORM_L2)
        return fail(nullptr, SFMM_EINVAL, "sfmm_create: unknown norm");
    if (!(cfg->ratio >= 0.f)) return fail(nullptr, SFMM_EINVAL, "sfmm_create: ratio must be >= 0");
    if (cfg->float_mode < SFMM_FLOAT_AUTO || cfg->float_mode > SFMM_FLOAT_TENSOR)
        return fail(nullptr, SFMM_EINVAL, "sfmm_create: unknown float_mode");
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        (void)cudaGetLastError();
        return fail(nullptr, SFMM_ENODEVICE,
                    std::string("sfmm_create: no CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback");
    }
    if (cfg->device < 0 || cfg->device >= n_dev) return fail(nullptr, SFMM_ERANGE, "sfmm_create: device ordinal out of range");
    cudaDeviceProp prop{};
    if ((e = cudaGetDeviceProperties(&prop, cfg->device)) != cudaSuccess)
        return fail(nullptr, SFMM_ECUDA, std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, SFMM_ENODEVICE, "sfmm_create: kernels are built for sm_100a (Blackwell B200) only, found " +
                                                 std::string(prop.name));
    SfmmCtx* ctx = new (std::nothrow) SfmmCtx();
    if (!ctx) return fail(nullptr, SFMM_ENOMEM, "sfmm_create: out of host memory");
    ctx->cfg = *cfg;
    ctx->sm_count = prop.multiProcessorCount;
    if (const char* s = std::getenv("SFMM_CSA_LEVEL")) ctx->csa_level = std::max(0, std::min(3, std::atoi(s)));
    if ((e = cudaSetDevice(cfg->device)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        delete ctx;
        return fail(nullptr, SFMM_ECUDA, std::string("stream setup: ") + cudaGetErrorString(e));
    }
    for (auto& ev : ctx->ev)
        if ((e = cudaEventCreate(&ev)) != cudaSuccess) {
            sfmm_destroy(ctx);
            return fail(nullptr, SFMM_ECUDA, std::string("cudaEventCreate: ") + cudaGetErrorString(e));
        }
    *out = ctx;
    return SFMM_OK;
}

SFMM_API void sfmm_destroy(SfmmCtx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->cfg.device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (DevBuf* b : {&ctx->blob, &ctx->d_pairs, &ctx->d_tiles, &ctx->d_ftiles, &ctx->d_knn, &ctx->d_colmin, &ctx->d_tile_count,
                      &ctx->d_tile_off, &ctx->d_pair_count, &ctx->d_pair_off, &ctx->d_matches, &ctx->d_idx, &ctx->d_dist, &ctx->d_norms, &ctx->d_flags})
        b->release();
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    for (auto& ev : ctx->ev)
        if (ev) cudaEventDestroy(ev);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

SFMM_API size_t sfmm_row_pitch(int32_t cols, int32_t elem_type) {
    if (cols <= 0) return 0;
    if (elem_type == SFMM_U8) return static_cast<size_t>(binary_words(cols)) * 4;
    if (elem_type == SFMM_F32) return (static_cast<size_t>(cols) * 4 + 15) / 16 * 16;
    return 0;
}

SFMM_API int sfmm_set_descriptors(SfmmCtx* ctx, int32_t n_images, const void* const* data, const int32_t* rows,
                                  int32_t cols, const size_t* step_bytes, int32_t elem_type) {
    if (!ctx) return SFMM_EINVAL;
    if (n_images < 0 || (n_images > 0 && !rows) || cols <= 0) return fail(ctx, SFMM_EINVAL, "set_descriptors: bad sizes");
    if (elem_type != SFMM_U8 && elem_type != SFMM_F32) return fail(ctx, SFMM_EINVAL, "set_descriptors: unknown elem_type");
    // cv::BFMatcher asserts the same pairing: NORM_HAMMING needs CV_8U, the L2 path here takes CV_32F
    if ((ctx->cfg.norm == SFMM_NORM_HAMMING) != (elem_type == SFMM_U8))
        return fail(ctx, SFMM_EINVAL, "set_descriptors: NORM_HAMMING needs SFMM_U8 rows and NORM_L2 needs SFMM_F32 rows");
    const size_t pitch = sfmm_row_pitch(cols, elem_type);
    if (pitch == 0) return fail(ctx, SFMM_EINVAL, "set_descriptors: unsupported descriptor width (binary <= 128 bytes)");
    if (elem_type == SFMM_F32 && cols > 256) return fail(ctx, SFMM_EINVAL, "set_descriptors: float descriptors wider than 256 are not supported");
    const size_t elem = elem_type == SFMM_U8 ? 1 : 4;
    // every image starts at a blob row that is a multiple of 4 (16-byte aligned slices of the
    // per-row norm array for the TMA bulk copies of the tensor path); pad rows are zero
    uint64_t total = 0;
    for (int32_t i = 0; i < n_images; ++i) {
        if (rows[i] < 0) return fail(ctx, SFMM_EINVAL, "set_descriptors: negative row count");
        if (rows[i] >= (1 << IDX_BITS)) return fail(ctx, SFMM_ERANGE, "set_descriptors: an image has >= 2^18 rows (OpenCV's BFMatcher limit)");
        if (data && rows[i] > 0 && !data[i]) return fail(ctx, SFMM_EINVAL, "set_descriptors: NULL image data");
        if (step_bytes && rows[i] > 1 && step_bytes[i] < static_cast<size_t>(cols) * elem)
            return fail(ctx, SFMM_EINVAL, "set_descriptors: row step smaller than a row");
        total += (static_cast<uint64_t>(rows[i]) + 3) & ~3ull;
    }
    if (total >= (1ull << 32)) return fail(ctx, SFMM_ERANGE, "set_descriptors: more than 2^32 rows in total");
    int rc = bind_device(ctx);
    if (rc) return rc;
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    clear_results(ctx);
    ctx->elem_type = -1;
    ctx->float_prepared = false;
    ctx->use_tensor = false;
    ctx->stats.float_path = 0;
    const size_t bytes = static_cast<size_t>(total) * pitch;
    CU_TRY(ctx, ctx->blob.ensure(std::max<size_t>(bytes, 16)));
    ctx->rows.assign(rows, rows + n_images);
    ctx->row0.resize(n_images);
    uint64_t r0 = 0;
    for (int32_t i = 0; i < n_images; ++i) {
        ctx->row0[i] = static_cast<uint32_t>(r0);
        r0 += (static_cast<uint64_t>(rows[i]) + 3) & ~3ull;
    }
    if (data && bytes) {
        // re-pitch on the host into pinned memory (zeroed padding: zeros are Hamming/L2 neutral), one H2D per slab
        const size_t slab = std::min<size_t>(bytes, 256u << 20);
        rc = ensure_pinned(ctx, slab);
        if (rc) return rc;
        const size_t row_bytes = static_cast<size_t>(cols) * elem;
        unsigned char* stage = static_cast<unsigned char*>(ctx->pinned);
        size_t fill = 0, dev_off = 0;
        auto flush = [&]() -> int {
            if (!fill) return SFMM_OK;
            CU_TRY(ctx, cudaMemcpyAsync(static_cast<unsigned char*>(ctx->blob.p) + dev_off, stage, fill, cudaMemcpyHostToDevice, ctx->stream));
            CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
            ctx->stats.h2d_bytes += static_cast<int64_t>(fill);
            dev_off += fill;
            fill = 0;
            return SFMM_OK;
        };
        for (int32_t i = 0; i < n_images; ++i) {
            const unsigned char* src = static_cast<const unsigned char*>(data[i]);
            const size_t step = step_bytes ? step_bytes[i] : row_bytes;
            for (int32_t r = 0; r < rows[i]; ++r) {
                if (fill + pitch > slab && (rc = flush())) return rc;
                std::memcpy(stage + fill, src + static_cast<size_t>(r) * step, row_bytes);
                if (pitch > row_bytes) std::memset(stage + fill + row_bytes, 0, pitch - row_bytes);
                fill += pitch;
            }
            for (int32_t r = rows[i]; r & 3; ++r) {  // zero rows up to the next multiple of 4
                if (fill + pitch > slab && (rc = flush())) return rc;
                std::memset(stage + fill, 0, pitch);
                fill += pitch;
            }
        }
        if ((rc = flush())) return rc;
    }
    ctx->n_images = n_images;
    ctx->cols = cols;
    ctx->pitch = pitch;
    ctx->total_rows = total;
    ctx->blob_bytes = bytes;
    ctx->elem_type = elem_type;
    return SFMM_OK;
}

SFMM_API int sfmm_descriptor_blob(SfmmCtx* ctx, void** device_ptr, size_t* bytes) {
    int rc = require_descriptors(ctx);
    if (rc) return rc;
    if (!device_ptr || !bytes) return fail(ctx, SFMM_EINVAL, "descriptor_blob: NULL argument");
    *device_ptr = ctx->blob.p;
    *bytes = ctx->blob_bytes;
    return SFMM_OK;
}

SFMM_API int sfmm_clear_results(SfmmCtx* ctx) {
    if (!ctx) return SFMM_EINVAL;
    clear_results(ctx);
    return SFMM_OK;
}

SFMM_API int sfmm_match_pairs_device(SfmmCtx* ctx, const int32_t* qt, int64_t n_pairs, int32_t* d_counts,
                                     SfmDMatch* d_matches, int64_t match_capacity, int64_t* n_matches) {
    int rc = require_descriptors(ctx);
    if (rc) return rc;
    if (n_pairs < 0 || (n_pairs > 0 && (!qt || !d_counts)) || match_capacity < 0 || !n_matches)
        return fail(ctx, SFMM_EINVAL, "match_pairs_device: bad argument");
    if ((rc = bind_device(ctx))) return rc;
    if ((rc = prepare_float(ctx))) return rc;
    *n_matches = 0;
    ctx->stats.last_knn_ms = ctx->stats.last_knn_work = 0;
    ctx->stats.last_knn_launches = 0;
    CU_TRY(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    int64_t done = 0;
    uint64_t written = 0;
    while (done < n_pairs) {
        const int64_t n = chunk_extent(ctx, qt, done, n_pairs);
        ChunkPlan plan;
        if ((rc = plan_chunk(ctx, qt + 2 * done, n, plan))) return rc;
        if ((rc = run_knn(ctx, plan, true))) return rc;
        uint64_t total = 0;
        const uint64_t room = static_cast<uint64_t>(match_capacity) - std::min<uint64_t>(written, match_capacity);
        if ((rc = run_filter(ctx, plan, d_counts + done, d_matches ? d_matches + written : nullptr, d_matches ? room : 0, &total))) return rc;
        account_knn_time(ctx, plan);
        if (total > room) return fail(ctx, SFMM_ERANGE, "match_pairs_device: match_capacity too small");
        written += total;
        done += n;
    }
    CU_TRY(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    CU_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
    ctx->stats.last_match_ms = ms;
    ctx->stats.pairs_matched += n_pairs;
    *n_matches = static_cast<int64_t>(written);
    return SFMM_OK;
}

SFMM_API int sfmm_match_pairs(SfmmCtx* ctx, const int32_t* qt, int64_t n_pairs) {
    int rc = require_descriptors(ctx);
    if (rc) return rc;
    if (n_pairs < 0 || (n_pairs > 0 && !qt)) return fail(ctx, SFMM_EINVAL, "match_pairs: bad argument");
    if ((rc = bind_device(ctx))) return rc;
    if ((rc = prepare_float(ctx))) return rc;
    ctx->stats.last_knn_ms = ctx->stats.last_knn_work = 0;
    ctx->stats.last_knn_launches = 0;
    CU_TRY(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    int64_t done = 0;
    while (done < n_pairs) {
        const int64_t n = chunk_extent(ctx, qt, done, n_pairs);
        ChunkPlan plan;
        if ((rc = plan_chunk(ctx, qt + 2 * done, n, plan))) return rc;
        if ((rc = run_knn(ctx, plan, true))) return rc;
        CU_TRY(ctx, ctx->d_pair_count.ensure(static_cast<size_t>(n) * sizeof(int32_t)));
        CU_TRY(ctx, ctx->d_matches.ensure(std::max<uint64_t>(1, plan.max_matches) * sizeof(SfmDMatch)));
        uint64_t total = 0;
        if ((rc = run_filter(ctx, plan, ctx->d_pair_count.as<int32_t>(), ctx->d_matches.as<SfmDMatch>(), plan.max_matches, &total))) return rc;
        account_knn_time(ctx, plan);
        // device -> pinned -> the chunk's host block
        const size_t meta = (static_cast<size_t>(n) * (sizeof(int32_t) + sizeof(unsigned long long)) + 15) / 16 * 16;
        if ((rc = ensure_pinned(ctx, meta + static_cast<size_t>(total) * sizeof(SfmDMatch)))) return rc;
        unsigned char* pin = static_cast<unsigned char*>(ctx->pinned);
        unsigned long long* h_off = reinterpret_cast<unsigned long long*>(pin);
        int32_t* h_cnt = reinterpret_cast<int32_t*>(pin + static_cast<size_t>(n) * sizeof(unsigned long long));
        SfmDMatch* h_m = reinterpret_cast<SfmDMatch*>(pin + meta);
        CU_TRY(ctx, cudaMemcpyAsync(h_off, ctx->d_pair_off.p, static_cast<size_t>(n) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(ctx, cudaMemcpyAsync(h_cnt, ctx->d_pair_count.p, static_cast<size_t>(n) * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        if (total)
            CU_TRY(ctx, cudaMemcpyAsync(h_m, ctx->d_matches.p, static_cast<size_t>(total) * sizeof(SfmDMatch), cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->stats.d2h_bytes += static_cast<int64_t>(meta + total * sizeof(SfmDMatch));
        HostChunk hc;
        hc.n = static_cast<int64_t>(total);
        hc.data.reset(new (std::nothrow) SfmDMatch[std::max<uint64_t>(1, total)]);
        if (!hc.data) return fail(ctx, SFMM_ENOMEM, "match_pairs: out of host memory for the match table");
        if (total) std::memcpy(hc.data.get(), h_m, static_cast<size_t>(total) * sizeof(SfmDMatch));
        const SfmDMatch* base = hc.data.get();
        int64_t running = 0;
        ctx->chunks.push_back(std::move(hc));
        for (int64_t i = 0; i < n; ++i) {
            const int32_t q = qt[2 * (done + i)], t = qt[2 * (done + i) + 1];
            ctx->index[pair_key(q, t)] = static_cast<int64_t>(ctx->res_slots.size());
            ctx->res_qt.push_back(q);
            ctx->res_qt.push_back(t);
            ctx->res_counts.push_back(h_cnt[i]);
            // records are laid out in pair order, so the offset is the running total (h_off[i] for
            // non-empty pairs; empty pairs have no filter tile and get the position they would occupy)
            const int64_t local = h_cnt[i] ? static_cast<int64_t>(h_off[i]) : running;
            running = local + h_cnt[i];
            ctx->res_offsets.push_back(ctx->n_matches + local);
            ctx->res_slots.push_back(PairSlot{base + local, h_cnt[i]});
        }
        ctx->n_matches += static_cast<int64_t>(total);
        ctx->flat_valid = false;
        done += n;
    }
    CU_TRY(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    CU_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
    ctx->stats.last_match_ms = ms;
    ctx->stats.pairs_matched += n_pairs;
    return SFMM_OK;
}

SFMM_API int sfmm_match_all_pairs(SfmmCtx* ctx) {
    int rc = require_descriptors(ctx);
    if (rc) return rc;
    clear_results(ctx);
    // findBestPair's enumeration, src/Sfm.cpp:511-512
    std::vector<int32_t> qt;
    const int64_t n = ctx->n_images;
    qt.reserve(static_cast<size_t>(n > 1 ? n * (n - 1) : 0));
    for (int32_t q = 0; q + 1 < n; ++q)
        for (int32_t t = q + 1; t < n; ++t) {
            qt.push_back(q);
            qt.push_back(t);
        }
    return sfmm_match_pairs(ctx, qt.data(), static_cast<int64_t>(qt.size() / 2));
}

SFMM_API int sfmm_get_pair(const SfmmCtx* ctx, int32_t q, int32_t t, const SfmDMatch** matches, int32_t* count) {
    int rc = require_descriptors(ctx);
    if (rc) return rc;
    if (!matches || !count) return fail(ctx, SFMM_EINVAL, "get_pair: NULL argument");
    if (q < 0 || q >= ctx->n_images || t < 0 || t >= ctx->n_images) return fail(ctx, SFMM_ERANGE, "get_pair: image index out of range");
    auto it = ctx->index.find(pair_key(q, t));
    if (it == ctx->index.end()) return fail(ctx, SFMM_ESTATE, "get_pair: pair has not been matched (call sfmm_match_all_pairs / sfmm_match_pairs)");
    const PairSlot& s = ctx->res_slots[static_cast<size_t>(it->second)];
    *matches = s.ptr;
    *count = s.count;
    return SFMM_OK;
}

SFMM_API int sfmm_result_table(const SfmmCtx* ctx, int64_t* n_pairs, const int32_t** qt, const int32_t** counts,
                               const int64_t** offsets, const SfmDMatch** matches, int64_t* n_matches) {
    int rc = require_descriptors(ctx);
    if (rc) return rc;
    if (!n_pairs || !qt || !counts || !offsets || !matches || !n_matches) return fail(ctx, SFMM_EINVAL, "result_table: NULL argument");
    *n_pairs = static_cast<int64_t>(ctx->res_counts.size());
    *qt = ctx->res_qt.data();
    *counts = ctx->res_counts.data();
    *offsets = ctx->res_offsets.data();
    *n_matches = ctx->n_matches;
    if (ctx->chunks.size() == 1) {
        *matches = ctx->chunks[0].data.get();
    } else {
        if (!ctx->flat_valid) {
            ctx->flat.clear();
            ctx->flat.reserve(static_cast<size_t>(ctx->n_matches));
            for (const HostChunk& c : ctx->chunks) ctx->flat.insert(ctx->flat.end(), c.data.get(), c.data.get() + c.n);
            ctx->flat_valid = true;
        }
        *matches = ctx->flat.data();
    }
    return SFMM_OK;
}

SFMM_API int sfmm_match_pair(SfmmCtx* ctx, int32_t q, int32_t t, SfmDMatch* out, int32_t cap, int32_t* count) {
    int rc = require_descriptors(ctx);
    if (rc) return rc;
    if (!count || cap < 0 || (cap > 0 && !out)) return fail(ctx, SFMM_EINVAL, "match_pair: bad argument");
    if (q < 0 || q >= ctx->n_images || t < 0 || t >= ctx->n_images) return fail(ctx, SFMM_ERANGE, "match_pair: image index out of range");
    if ((rc = bind_device(ctx))) return rc;
    if ((rc = prepare_float(ctx))) return rc;
    *count = 0;
    const int32_t qt[2] = {q, t};
    ChunkPlan plan;
    if ((rc = plan_chunk(ctx, qt, 1, plan))) return rc;
    if ((rc = run_knn(ctx, plan, false))) return rc;
    CU_TRY(ctx, ctx->d_pair_count.ensure(sizeof(int32_t)));
    CU_TRY(ctx, ctx->d_matches.ensure(std::max<uint64_t>(1, plan.max_matches) * sizeof(SfmDMatch)));
    uint64_t total = 0;
    if ((rc = run_filter(ctx, plan, ctx->d_pair_count.as<int32_t>(), ctx->d_matches.as<SfmDMatch>(), plan.max_matches, &total))) return rc;
    ctx->stats.pairs_matched += 1;
    if (total > static_cast<uint64_t>(cap)) {
        *count = static_cast<int32_t>(total);
        return fail(ctx, SFMM_ERANGE, "match_pair: output capacity too small (count holds the size needed)");
    }
    if (total) {
        CU_TRY(ctx, cudaMemcpyAsync(out, ctx->d_matches.p, static_cast<size_t>(total) * sizeof(SfmDMatch), cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->stats.d2h_bytes += static_cast<int64_t>(total * sizeof(SfmDMatch));
    }
    *count = static_cast<int32_t>(total);
    return SFMM_OK;
}

SFMM_API int sfmm_knn_pair(SfmmCtx* ctx, int32_t q, int32_t t, int32_t* train_idx, float* distance) {
    int rc = require_descriptors(ctx);
    if (rc) return rc;
    if (q < 0 || q >= ctx->n_images || t < 0 || t >= ctx->n_images) return fail(ctx, SFMM_ERANGE, "knn_pair: image index out of range");
    const int32_t nq = ctx->rows[q], nt = ctx->rows[t];
    if (nq == 0) return SFMM_OK;
    if (!train_idx || !distance) return fail(ctx, SFMM_EINVAL, "knn_pair: NULL output");
    if (nt == 0) {
        for (int32_t i = 0; i < 2 * nq; ++i) {
            train_idx[i] = -1;
            distance[i] = FLT_MAX;
        }
        return SFMM_OK;
    }
    if ((rc = bind_device(ctx))) return rc;
    if ((rc = prepare_float(ctx))) return rc;
    // plan as a normal pair but force the knn launch even when nt == 1
    const int32_t qt[2] = {q, t};
    ChunkPlan plan;
    if ((rc = plan_chunk(ctx, qt, 1, plan))) return rc;
    if (plan.tiles.empty()) {  // nt == 1: planned as "no matches"; build the single split by hand
        PairDesc& pd = plan.pairs[0];
        pd.n_splits = 1;
        const bool is_float = ctx->elem_type == SFMM_F32;
        const int W = is_float ? 0 : binary_words(ctx->cols);
        const uint32_t q_tile = is_float ? (ctx->use_tensor ? FT_M : FX_BQ) : BK_THREADS * (W >= 32 ? 2 : 4);
        for (uint32_t q0 = 0; q0 < pd.nq; q0 += q_tile) plan.tiles.push_back(KnnTile{0, q0, 0, pd.nt, 0});
        plan.knn_entries = pd.nq;
        plan.col_entries = pd.nt;
    }
    if ((rc = run_knn(ctx, plan, false))) return rc;
    CU_TRY(ctx, ctx->d_idx.ensure(static_cast<size_t>(nq) * 2 * sizeof(int32_t)));
    CU_TRY(ctx, ctx->d_dist.ensure(static_cast<size_t>(nq) * 2 * sizeof(float)));
    const int threads = 256, blocks = (nq + threads - 1) / threads;
    if (ctx->elem_type == SFMM_F32)
        knn_decode_kernel<true><<<blocks, threads, 0, ctx->stream>>>(ctx->d_pairs.as<PairDesc>(), ctx->d_knn.as<KnnEntry>(), ctx->d_idx.as<int32_t>(), ctx->d_dist.as<float>());
    else
        knn_decode_kernel<false><<<blocks, threads, 0, ctx->stream>>>(ctx->d_pairs.as<PairDesc>(), ctx->d_knn.as<KnnEntry>(), ctx->d_idx.as<int32_t>(), ctx->d_dist.as<float>());
    CU_TRY(ctx, cudaGetLastError());
    ctx->stats.kernel_launches += 1;
    CU_TRY(ctx, cudaMemcpyAsync(train_idx, ctx->d_idx.p, static_cast<size_t>(nq) * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaMemcpyAsync(distance, ctx->d_dist.p, static_cast<size_t>(nq) * 2 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stats.d2h_bytes += static_cast<int64_t>(nq) * 16;
    return SFMM_OK;
}

SFMM_API int sfmm_get_stats(const SfmmCtx* ctx, SfmmStats* out) {
    if (!ctx || !out) return SFMM_EINVAL;
    *out = ctx->stats;
    return SFMM_OK;
}

}  // extern "C"
