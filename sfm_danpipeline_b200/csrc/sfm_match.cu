// libsfmmatch.so -- C ABI (include/sfm_match.h) + host-side pair scheduler for the sm_100a
// descriptor-matching kernels.  There is no CPU path in this file: if a CUDA device is
// missing every entry point fails with SFMM_ENODEVICE.
//
// Reference path being replaced: StructFromMotion::getMatching (/root/reference/src/Sfm.cpp:590-608)
// driven by findBestPair's q<t loop (:511-515).  The scheduler turns a list of image pairs into
// "chunks" (one kernel launch each):
//   1. "knn tiles"    (query-row tile x train-row range) -> binary_knn2_kernel / float kernels
//   2. "filter tiles" (1024 query rows)                  -> ratio test + cross-check + compaction
// Two chunk slots (own stream, own scratch, own pinned staging) are kept in flight so that the
// device->host copy and host-side bookkeeping of chunk k overlap the kernels of chunk k+1.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <nccl.h>  // types and prototypes only: the library is dlopen'ed by sfmm_group_create, never linked

#include "../../include/sfm_match.h"
#include "binary_knn.cuh"
#include "common.cuh"
#include "filter.cuh"
#include "float_exact.cuh"
#include "float_tensor.cuh"
#include "float_tensor_ts.cuh"

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

struct PinBuf {  // page-locked host staging
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        const size_t want = bytes + bytes / 4 + 4096;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

// The match table: every SfmDMatch record of every pair matched so far, in page-locked host memory so that a chunk's
// records are copied device -> host straight into their final place (no staging buffer, no second copy on the host), on
// the copy stream, while the next chunk's kernels run.  Falls back to pageable memory when the pinned allocation fails
// (very large tables): the copies then go through the driver's own staging and are synchronous, nothing else changes.
struct HostTable {
    SfmDMatch* p = nullptr;
    size_t n = 0, cap = 0;
    bool pinned = false;
    // Shared mode (sfmm_share_table): the block is a POSIX shared-memory segment "<prefix>.<generation>", page-locked with
    // cudaHostRegister, so that another process of the same host (rank 0 of a multi-process job) can map the records this GPU
    // wrote over its own PCIe link -- the multi-process form of "every device copies into one host table".
    std::string shm_prefix, shm_name;
    int shm_generation = 0;
    size_t map_bytes = 0;
    SfmDMatch* data() const { return p; }
    size_t size() const { return n; }
    size_t capacity() const { return cap; }
    void clear() { n = 0; }
    void resize_down(size_t m) { if (m < n) n = m; }
    void release() {
        if (p) {
            if (!shm_name.empty()) {
                if (pinned) cudaHostUnregister(p);
                munmap(p, map_bytes);
                shm_unlink(shm_name.c_str());
                shm_name.clear();
            } else if (pinned) {
                cudaFreeHost(p);
            } else {
                std::free(p);
            }
        }
        (void)cudaGetLastError();
        p = nullptr;
        n = cap = 0;
        map_bytes = 0;
    }
    // Capacity for `want` records.  `drain` is synchronised before the old block is given up: copies into it may be in flight.
    int reserve(size_t want, cudaStream_t drain) {
        if (want <= cap) return SFMM_OK;
        const size_t ncap = std::max<size_t>(want, 1 << 12);
        void* q = nullptr;
        bool pin = true;
        std::string name;
        size_t bytes = ncap * sizeof(SfmDMatch);
        if (!shm_prefix.empty()) {
            name = shm_prefix + "." + std::to_string(shm_generation++);
            bytes = (bytes + 4095) & ~size_t(4095);
            const int fd = shm_open(name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
            if (fd < 0) return SFMM_ENOMEM;
            // posix_fallocate (not ftruncate): the blocks are reserved now, so a /dev/shm that is too small is an error code here and
            // not a SIGBUS at the first copy into the mapping
            if (posix_fallocate(fd, 0, static_cast<off_t>(bytes)) != 0) {
                close(fd);
                shm_unlink(name.c_str());
                return SFMM_ENOMEM;
            }
            q = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_POPULATE, fd, 0);
            close(fd);
            if (q == MAP_FAILED) {
                shm_unlink(name.c_str());
                return SFMM_ENOMEM;
            }
            if (cudaHostRegister(q, bytes, cudaHostRegisterPortable) != cudaSuccess) {
                (void)cudaGetLastError();
                pin = false;  // still correct: copies into it are staged by the driver
            }
        } else if (cudaMallocHost(&q, bytes) != cudaSuccess) {
            (void)cudaGetLastError();
            pin = false;
            q = std::malloc(bytes);
            if (!q) return SFMM_ENOMEM;
        }
        if (drain) cudaStreamSynchronize(drain);
        if (n) std::memcpy(q, p, n * sizeof(SfmDMatch));
        const size_t keep = n;
        const std::string prefix = shm_prefix;
        const int gen = shm_generation;
        release();
        shm_prefix = prefix;
        shm_generation = gen;
        p = static_cast<SfmDMatch*>(q);
        n = keep;
        cap = ncap;
        pinned = pin;
        shm_name = name;
        map_bytes = bytes;
        return SFMM_OK;
    }
};

struct PairSlot {  // where a computed pair lives in the host table
    int64_t offset;
    int32_t count;
};

// Binary kernel geometry (see binary_knn.cuh)
constexpr int BK_THREADS = 128;
constexpr int BK_TT = 128;
// query rows per thread: 32 registers of query words whatever the width (measured best: more
// resident warps beat more reuse of the broadcast train words, profiles/binary_variants_r01.txt)
template <int W> struct BkTq { static constexpr int v = (W >= 16) ? 2 : 4; };

struct ChunkPlan {
    std::vector<PairDesc> pairs;
    std::vector<KnnTile> tiles;
    std::vector<FilterTile> ftiles;
    std::vector<uint32_t> pair_of_row;  // TF32 rank + refine path: owning pair of every query row of the launch
    uint64_t knn_entries = 0;
    uint64_t col_entries = 0;
    uint64_t max_matches = 0;
    uint64_t max_rtiles = 0;  // upper bound of the cross-check's gathered reverse items (cross_compact_kernel)
    double work = 0;  // algorithmic POPC32 ops / FLOPs of the knn launch
    void clear() {
        pairs.clear(); tiles.clear(); ftiles.clear(); pair_of_row.clear();
        knn_entries = col_entries = max_matches = max_rtiles = 0;
        work = 0;
    }
};

// One chunk in flight: its stream, device scratch, pinned staging and events.
struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_knn0 = nullptr, ev_knn1 = nullptr, ev_done = nullptr, ev_copied = nullptr;
    bool copy_pending = false;  // records of this slot's last chunk are still being copied to the host table (ev_copied)
    DevBuf d_pairs, d_tiles, d_ftiles, d_knn, d_colmin, d_tile_count, d_tile_off, d_pair_count, d_pair_off, d_matches, d_left, d_right;
    DevBuf d_cand_count, d_cand_idx, d_pair_of_row;  // TF32 rank + refine path
    DevBuf d_xflags, d_xcand, d_n_xcand, d_rtiles, d_n_rtiles;  // cross-check through candidate columns (filter.cuh)
    PinBuf meta;     // [total u64][pair_off u64 x n][pair_count i32 x n]
    PinBuf points;   // aligned-point staging (left then right)
    ChunkPlan plan;
    int64_t first = 0, n = 0;  // pair range [first, first+n) of the caller's list
    bool busy = false;
    void release() {
        for (DevBuf* b : {&d_pairs, &d_tiles, &d_ftiles, &d_knn, &d_colmin, &d_tile_count, &d_tile_off, &d_pair_count, &d_pair_off, &d_matches, &d_left, &d_right,
                          &d_cand_count, &d_cand_idx, &d_pair_of_row, &d_xflags, &d_xcand, &d_n_xcand, &d_rtiles, &d_n_rtiles})
            b->release();
        meta.release();
        points.release();
        if (ev_knn0) cudaEventDestroy(ev_knn0);
        if (ev_knn1) cudaEventDestroy(ev_knn1);
        if (ev_done) cudaEventDestroy(ev_done);
        if (ev_copied) cudaEventDestroy(ev_copied);
        if (stream) cudaStreamDestroy(stream);
        stream = nullptr;
        ev_knn0 = ev_knn1 = ev_done = ev_copied = nullptr;
    }
};

}  // namespace

struct SfmmCtx {
    SfmmConfig cfg{};
    int sm_count = 0;
    Slot slot[2];
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr, ev_pack[2] = {nullptr, nullptr};
    mutable std::string err;
    int csa_level = 2;
    bool cross_full_reverse = false;  // SFMM_CROSS_FULL=1: the round-1 cross-check (a full reverse pass), for A/B measurements
    bool epi_groups_set = false;
    int epi_groups = 2;  // epilogue groups of the TMEM-A float kernels (SFMM_EPI_GROUPS=4: measured slower, see float_tensor_ts.cuh)
    size_t fx_attr_smem = 0;

    // float path state (norms + TF32-exactness proof, see float_tensor.cuh)
    bool float_prepared = false;
    bool tensor_eligible = false;
    bool use_tensor = false;   // a tcgen05 kernel is in use (float TF32, or binary through kind::i8)
    int tensor_kblocks = 0;    // 128-byte K-blocks per operand row
    bool tensor_refine = false; // arbitrary floats: TF32 ranking pass + candidate collection + exact refinement
    std::vector<float> img_maxnorm2;  // per image max |x|^2 (error bound of the ranking pass)
    int tensor_ts = -1;        // query tile in tensor memory (float_tensor_ts.cuh): -1 = where it measured faster (binary engine), 0/1 = SFMM_TENSOR_TS
    CUtensorMap tmap{};
    DevBuf d_norms, d_flags;
    float rank_offset = 0.f;  // C of the ranking pass's key table (float_nbexact_kernel)
    bool tensor_f16 = false;  // float tensor path runs on an fp16 copy (d_half) with kind::f16
    DevBuf d_half;
    DevBuf d_f4_a;            // TM_F4X: the query-side operand rows (bits as 1.0 | 6 x16, 1), d_unpacked holds the train side
    DevBuf d_half_a;          // TM_F16X: the query-side operand rows (-2q | 1, 2048, 2048), d_half holds the train side
    bool tensor_kx = false;   // TM_F16X in use (key term contracted by the tensor core)
    bool no_kx = false;       // SFMM_NO_KX=1: keep TM_F16_EXACT (A/B measurements)
    bool tensor_f4 = false;   // binary tensor engine on the FP4 pipe (TM_F4P: descriptors below 512 bit; SFMM_NO_F4=1 keeps kind::i8)
    bool tensor_f4x = false;  // TM_F4X: key term in the MMA + threshold-skipping epilogue (needs 17 spare elements per row; SFMM_NO_F4X=1 / SFMM_NO_SKIP=1: TM_F4P)
    bool no_f4 = false;
    bool no_skip = false;     // SFMM_NO_SKIP=1: TM_F16X folds every column (no threshold skipping; A/B measurements)
    uint32_t i8_bias = 0;    // binary tensor engine: descriptor bit length when the packed 16-bit keys apply (< 512 bit), else 0
    DevBuf d_nbkey, d_row0;  // binary tensor engine: per-row key part (binary_nbkey_kernel) and the images' first rows
    DevBuf d_unpacked;  // SFMM_BINARY_TENSOR: one byte per descriptor bit

    // descriptors (imagesDescriptors, include/Sfm.h:29)
    int32_t n_images = 0;
    std::vector<int32_t> rows;
    std::vector<uint32_t> row0;
    int32_t cols = 0;       // elements per blob row as the kernels see them
    int32_t elem_type = -1; // element type of the blob (SFMM_F32 for widened CV_8U rows under NORM_L2)
    int32_t src_cols = 0;   // what the caller passed to sfmm_set_descriptors
    int32_t src_elem_type = -1;
    int32_t elem_type_pending = -1;  // blob element type while sfmm_set_descriptors is still filling it
    size_t pitch = 0;
    uint64_t total_rows = 0;  // blob rows, including the zero rows that align every image to 4 rows
    DevBuf blob;
    size_t blob_bytes = 0;
    PinBuf pack[2];  // double-buffered pinned staging for set_descriptors
    std::vector<uint64_t> raw_row0;  // first row of every image counted without alignment rows (n_images + 1)
    std::vector<uint32_t> raw_row0_32;
    DevBuf d_raw, d_raw_row0;        // the caller's rows as uploaded (tightly packed), before ingest_rows_kernel
    cudaEvent_t ev_blob = nullptr;   // the blob is complete on the device (recorded on slot[0]'s stream)
    bool blob_async = false;

    DevBuf d_idx, d_dist;  // sfmm_knn_pair
    DevBuf d_points;       // imagesPts2D (double2 per blob row), optional
    bool have_points = false;

    // result table
    std::vector<int32_t> res_qt;
    std::vector<int32_t> res_counts;
    std::vector<int64_t> res_offsets;  // offsets into the consolidated view
    std::vector<PairSlot> res_slots;
    HostTable table;  // every record, in the order the pairs were given
    std::vector<double2> pts_left, pts_right;  // aligned points of every record (when points are set)
    std::unordered_map<uint64_t, int64_t> index;
    int64_t n_matches = 0;
    int64_t call_pairs = 0;  // pairs of the sfmm_match_pairs call in progress
    size_t call_base = 0;    // table size when it started

    SfmmStats stats{};
};

namespace {

int fail(const SfmmCtx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    else g_create_error = msg;
    return code;
}

// Nothing may throw across the C ABI: entry points that allocate on the host run their body through this.
template <class F>
int guarded(const SfmmCtx* ctx, F&& body) {
    try {
        return body();
    } catch (const std::bad_alloc&) {
        return fail(ctx, SFMM_ENOMEM, "out of host memory");
    } catch (const std::exception& e) {
        return fail(ctx, SFMM_EINVAL, std::string("unexpected C++ exception: ") + e.what());
    } catch (...) {
        return fail(ctx, SFMM_EINVAL, "unexpected C++ exception");
    }
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

void clear_results(SfmmCtx* ctx) {
    ctx->res_qt.clear();
    ctx->res_counts.clear();
    ctx->res_offsets.clear();
    ctx->res_slots.clear();
    ctx->table.clear();
    ctx->pts_left.clear();
    ctx->pts_right.clear();
    ctx->index.clear();
    ctx->n_matches = 0;
}

inline uint64_t pair_key(int32_t q, int32_t t) { return (static_cast<uint64_t>(static_cast<uint32_t>(q)) << 32) | static_cast<uint32_t>(t); }

uint32_t query_tile_rows(const SfmmCtx* ctx) {
    if (ctx->use_tensor) return FT_M;
    if (ctx->elem_type == SFMM_F32) return FX_BQ;
    return BK_THREADS * (binary_words(ctx->cols) >= 16 ? 2 : 4);
}

// Does the tensor path of the current descriptor set run the TMEM-A kernel (float_tensor_ts.cuh)?  Mirrors launch_tensor_t.
bool tensor_uses_ts(const SfmmCtx* ctx) {
    if (!ctx->use_tensor) return false;
    if (ctx->tensor_kx) return true;
    if (ctx->tensor_ts >= 0) return ctx->tensor_ts != 0;
    if (ctx->elem_type == SFMM_F32) return ctx->tensor_f16;
    return ctx->i8_bias != 0 || ctx->tensor_kblocks <= 2;
}
// Cross-check through candidate columns + gathered reverse items: the TMEM-A kernel's query loaders can gather rows.
bool cross_by_candidates(const SfmmCtx* ctx) { return ctx->cfg.cross_check && tensor_uses_ts(ctx) && !ctx->tensor_refine && !ctx->cross_full_reverse; }
// Arbitrary floats (tensor ranking + exact refinement): the candidate columns' minima come from the exact fp32 kernel, which
// gathers its rows the same way -- candidates/Nt of an exact pass instead of running the whole pair on the exact kernel.
bool cross_by_candidates_exact(const SfmmCtx* ctx) { return ctx->cfg.cross_check && ctx->use_tensor && ctx->tensor_refine; }

// ---------------------------------------------------------------------------- planning
// Turn n pairs of qt into device work descriptors.
int plan_chunk(SfmmCtx* ctx, const int32_t* qt, int64_t n, ChunkPlan& plan) {
    plan.clear();
    const bool is_float = ctx->elem_type == SFMM_F32;
    const uint32_t q_tile = query_tile_rows(ctx);
    const uint32_t t_gran = ctx->use_tensor ? FT_N : (is_float ? FX_BT : BK_TT);
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
    const uint64_t want_tiles = static_cast<uint64_t>(ctx->sm_count) * (ctx->use_tensor ? 2 : 8);
    uint32_t splits_wanted = 1;
    if (base_tiles > 0 && base_tiles < want_tiles)
        splits_wanted = static_cast<uint32_t>(std::min<uint64_t>(32, (want_tiles + base_tiles - 1) / base_tiles));

    const double work_per_eval = is_float ? 2.0 * ctx->src_cols : static_cast<double>((ctx->cols * 8 + 31) / 32);
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
        pd.q_off = static_cast<uint32_t>(plan.pair_of_row.size());
        pd.t_maxnorm2 = ctx->tensor_refine ? ctx->img_maxnorm2[t] : 0.f;
        if (pd.nq == 0 || pd.nt < 2) continue;  // defined: no matches (see sfm_match.h)
        if (ctx->tensor_refine) plan.pair_of_row.insert(plan.pair_of_row.end(), pd.nq, static_cast<uint32_t>(i));
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
        if (cross_by_candidates_exact(ctx)) {
            plan.max_rtiles += (pd.nt + FX_BQ - 1) / FX_BQ;
        } else if (cross_by_candidates(ctx)) {
            plan.max_rtiles += (pd.nt + q_tile - 1) / q_tile;  // reverse items are made on the device, from the candidate columns
        } else if (ctx->use_tensor && ctx->cfg.cross_check) {
            // tensor kernels: the cross-check's column minima come from "reverse" tiles (roles swapped,
            // bit 31 of split), see float_tensor.cuh; rows = train rows, streamed = query rows
            uint32_t rsplits = std::min<uint32_t>(splits_wanted, std::max<uint32_t>(1, pd.nq / (2 * t_gran)));
            const uint32_t rper = ((pd.nq + rsplits - 1) / rsplits + t_gran - 1) / t_gran * t_gran;
            rsplits = (pd.nq + rper - 1) / rper;
            for (uint32_t r0 = 0; r0 < pd.nt; r0 += q_tile)
                for (uint32_t s2 = 0; s2 < rsplits; ++s2)
                    plan.tiles.push_back(KnnTile{static_cast<uint32_t>(i), r0, s2 * rper, std::min(pd.nq, (s2 + 1) * rper), s2 | TILE_REVERSE});
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
cudaError_t launch_binary_t(SfmmCtx* ctx, Slot& sl, uint32_t n_tiles) {
    constexpr int TQ = BkTq<W>::v;
    using Smem = BinaryKnnSmem<W, TQ, BK_THREADS, BK_TT>;
    static_assert(sizeof(Smem) <= 48 * 1024, "fits the default dynamic shared memory limit");
    binary_knn2_kernel<W, TQ, BK_THREADS, BK_TT, CSA, CROSS><<<n_tiles, BK_THREADS, sizeof(Smem), sl.stream>>>(
        ctx->blob.as<uint32_t>(), sl.d_tiles.as<KnnTile>(), sl.d_pairs.as<PairDesc>(), sl.d_knn.as<KnnEntry>(),
        sl.d_colmin.as<unsigned long long>(), KeyWeights{{1u << IDX_BITS, 2u << IDX_BITS, 4u << IDX_BITS}});
    return cudaGetLastError();
}

template <int W, bool CROSS>
cudaError_t launch_binary_w(SfmmCtx* ctx, Slot& sl, uint32_t n_tiles) {
    switch (ctx->csa_level) {
        case 0: return launch_binary_t<W, 0, CROSS>(ctx, sl, n_tiles);
        case 1: return launch_binary_t<W, 1, CROSS>(ctx, sl, n_tiles);
        case 3: return launch_binary_t<W, 3, CROSS>(ctx, sl, n_tiles);
        default: return launch_binary_t<W, 2, CROSS>(ctx, sl, n_tiles);
    }
}

template <bool CROSS>
cudaError_t launch_binary(SfmmCtx* ctx, Slot& sl, uint32_t n_tiles) {
    switch (binary_words(ctx->cols)) {
        case 4: return launch_binary_w<4, CROSS>(ctx, sl, n_tiles);
        case 8: return launch_binary_w<8, CROSS>(ctx, sl, n_tiles);
        case 16: return launch_binary_w<16, CROSS>(ctx, sl, n_tiles);
        case 32: return launch_binary_w<32, CROSS>(ctx, sl, n_tiles);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_float_exact(SfmmCtx* ctx, Slot& sl, uint32_t n_tiles, const KnnTile* tiles_dev = nullptr, const uint32_t* n_items_dev = nullptr) {
    const int kq = static_cast<int>(ctx->pitch / 16);
    const size_t smem = float_exact_smem_bytes(kq);
    if (smem > ctx->fx_attr_smem) {  // per context == per device
        cudaError_t e = cudaFuncSetAttribute(float_exact_knn2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        ctx->fx_attr_smem = smem;
    }
    float_exact_knn2_kernel<<<n_tiles, FX_THREADS, smem, sl.stream>>>(
        ctx->blob.as<float>(), kq, tiles_dev ? tiles_dev : (const KnnTile*)sl.d_tiles.as<KnnTile>(), sl.d_pairs.as<PairDesc>(), sl.d_knn.as<KnnEntry>(),
        sl.d_colmin.as<unsigned long long>(), (ctx->cfg.cross_check && !tiles_dev) ? 1 : 0, n_items_dev, (const uint32_t*)sl.d_xcand.as<uint32_t>(),
        (const uint32_t*)sl.d_n_xcand.as<uint32_t>());
    return cudaGetLastError();
}

// `tiles` / `n_items_dev`: NULL = the slot's planned tile list (n_tiles of them); otherwise a device-made list whose length
// lives on the device (at most n_tiles) -- the cross-check's gathered reverse items.
template <int KB, int MODE>
cudaError_t launch_tensor_t(SfmmCtx* ctx, Slot& sl, uint32_t n_tiles, const KnnTile* tiles_dev = nullptr, const uint32_t* n_items_dev = nullptr) {
    // persistent: one CTA per SM (shared memory allows no more) walks the tile list with stride gridDim.x
    const uint32_t grid = std::min<uint32_t>(n_tiles, static_cast<uint32_t>(ctx->sm_count));
    const KnnTile* tiles = tiles_dev ? tiles_dev : (const KnnTile*)sl.d_tiles.as<KnnTile>();
    const float* nb_src = (MODE == TM_I8 || MODE == TM_I8P || MODE == TM_F4P || MODE == TM_TF32_EXACT || MODE == TM_F16_EXACT || MODE == TM_F16X || tm_is_rank(MODE)) ? (const float*)ctx->d_nbkey.as<float>()
                                                                                                              : (const float*)ctx->d_norms.as<float>();
    // measured (profiles/tensor_variants_r01.txt): TMEM-A wins for the binary engine and the fp16 float path, shared-memory-A for TF32
    uint32_t aux = ctx->i8_bias;  // TM_I8P: descriptor bit length; rank modes: float bits of the key-table offset
    if (tm_is_rank(MODE) || tm_is_collect(MODE)) std::memcpy(&aux, &ctx->rank_offset, sizeof(aux));
    const bool ts = MODE == TM_F16X || MODE == TM_F4P || MODE == TM_F4X || (ctx->tensor_ts < 0 ? (MODE == TM_I8P || OperandOf<MODE>::kind == OK_F16 || (MODE == TM_I8 && KB <= 2)) : ctx->tensor_ts != 0);
    if constexpr (MODE != TM_F4P && MODE != TM_F4X && MODE != TM_F16X) if (!ts) {  // query tile in shared memory (float_tensor.cuh)
        const size_t smem = float_tensor_smem_bytes(KB);
        auto kern = tensor_knn2_kernel<KB, MODE>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (tiles_dev) return cudaErrorInvalidValue;  // gathered reverse items need the TMEM-A kernel (cross_by_candidates)
        kern<<<grid, FT_THREADS, smem, sl.stream>>>(ctx->tmap, (const float*)ctx->d_norms.as<float>(), nb_src, tiles, n_tiles,
                                                    (const PairDesc*)sl.d_pairs.as<PairDesc>(), sl.d_knn.as<KnnEntry>(),
                                                    sl.d_colmin.as<unsigned long long>(), 512u, aux, sl.d_cand_count.as<uint32_t>(),
                                                    sl.d_cand_idx.as<uint32_t>());
        return cudaGetLastError();
    }
    // query tile in tensor memory (float_tensor_ts.cuh)
    const uint4* a_src = MODE == TM_F4X ? ctx->d_f4_a.as<uint4>() : (MODE == TM_I8 || MODE == TM_I8P || MODE == TM_F4P) ? ctx->d_unpacked.as<uint4>() : (MODE == TM_F16X ? ctx->d_half_a.as<uint4>() : (OperandOf<MODE>::kind == OK_F16 ? ctx->d_half.as<uint4>() : ctx->blob.as<uint4>()));
    const size_t smem = float_tensor_ts_smem_bytes(KB);
    auto go = [&](auto kern, int threads) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<grid, threads, smem, sl.stream>>>(ctx->tmap, a_src, static_cast<uint32_t>(ctx->total_rows), (const float*)ctx->d_norms.as<float>(), nb_src,
                                                 tiles, n_tiles, (const PairDesc*)sl.d_pairs.as<PairDesc>(),
                                                 sl.d_knn.as<KnnEntry>(), sl.d_colmin.as<unsigned long long>(), 512u, aux,
                                                 sl.d_cand_count.as<uint32_t>(), sl.d_cand_idx.as<uint32_t>(), n_items_dev,
                                                 (const uint32_t*)sl.d_xcand.as<uint32_t>(), (const uint32_t*)sl.d_n_xcand.as<uint32_t>());
        return cudaGetLastError();
    };
    // SFMM_EPI_GROUPS=4: four epilogue groups for the 32-bit-key float modes (an experiment that measured 5 % slower than two,
    // profiles/tensor_variants_r02.txt; kept selectable and covered by the parity tests)
    if constexpr (MODE == TM_F16X) {  // threshold skipping in the epilogue (float_tensor.cuh, chunk_top2_skipx) unless SFMM_NO_SKIP=1
        if (!ctx->no_skip) {
            if (ctx->epi_groups == 4) return go(tensor_knn2_ts_kernel<KB, MODE, 4, true>, fts_threads(4));
            if (ctx->epi_groups == 3) return go(tensor_knn2_ts_kernel<KB, MODE, 3, true>, fts_threads(3));
            return go(tensor_knn2_ts_kernel<KB, MODE, 2, true>, fts_threads(2));
        }
    }
    if constexpr (MODE == TM_F4P) {
        if (ctx->epi_groups == 3) return go(tensor_knn2_ts_kernel<KB, MODE, 3>, fts_threads(3));
    }
    if constexpr (MODE == TM_F4X) {  // three epilogue groups by default (+2-3 %, same variants file); SFMM_EPI_GROUPS=2 for two
        if (ctx->epi_groups_set && ctx->epi_groups == 2) return go(tensor_knn2_ts_kernel<KB, MODE, 2, true>, fts_threads(2));
        return go(tensor_knn2_ts_kernel<KB, MODE, 3, true>, fts_threads(3));
    } else {
        if constexpr (MODE == TM_F16_EXACT || MODE == TM_TF32_EXACT || MODE == TM_F16X) {
            if (ctx->epi_groups == 4) return go(tensor_knn2_ts_kernel<KB, MODE, 4>, fts_threads(4));
        }
        return go(tensor_knn2_ts_kernel<KB, MODE, 2>, fts_threads(2));
    }
}

template <int MODE>
cudaError_t launch_tensor(SfmmCtx* ctx, Slot& sl, uint32_t n_tiles, int kblocks, const KnnTile* tiles_dev = nullptr, const uint32_t* n_items_dev = nullptr) {
    switch (kblocks) {
        case 1: return launch_tensor_t<1, MODE>(ctx, sl, n_tiles, tiles_dev, n_items_dev);
        case 2: return launch_tensor_t<2, MODE>(ctx, sl, n_tiles, tiles_dev, n_items_dev);
        case 3: return launch_tensor_t<3, MODE>(ctx, sl, n_tiles, tiles_dev, n_items_dev);
        case 4: return launch_tensor_t<4, MODE>(ctx, sl, n_tiles, tiles_dev, n_items_dev);
    }
    return cudaErrorInvalidValue;
}

// The exact tensor modes (results final after one pass) dispatched on the descriptor type; used for the forward pass and for the
// cross-check's gathered reverse pass.
cudaError_t launch_tensor_exact(SfmmCtx* ctx, Slot& sl, uint32_t n_tiles, const KnnTile* tiles_dev = nullptr, const uint32_t* n_items_dev = nullptr) {
    if (ctx->elem_type == SFMM_F32) {
        if (ctx->tensor_kx) return ctx->tensor_kblocks == 2 ? launch_tensor_t<2, TM_F16X>(ctx, sl, n_tiles, tiles_dev, n_items_dev)
                                                           : launch_tensor_t<3, TM_F16X>(ctx, sl, n_tiles, tiles_dev, n_items_dev);
        return ctx->tensor_f16 ? launch_tensor<TM_F16_EXACT>(ctx, sl, n_tiles, ctx->tensor_kblocks, tiles_dev, n_items_dev)
                               : launch_tensor<TM_TF32_EXACT>(ctx, sl, n_tiles, ctx->tensor_kblocks, tiles_dev, n_items_dev);
    }
    if (ctx->tensor_f4x) return ctx->tensor_kblocks == 1 ? launch_tensor_t<1, TM_F4X>(ctx, sl, n_tiles, tiles_dev, n_items_dev)
                                                         : launch_tensor_t<2, TM_F4X>(ctx, sl, n_tiles, tiles_dev, n_items_dev);
    if (ctx->tensor_f4) return ctx->tensor_kblocks == 1 ? launch_tensor_t<1, TM_F4P>(ctx, sl, n_tiles, tiles_dev, n_items_dev)
                                                        : launch_tensor_t<2, TM_F4P>(ctx, sl, n_tiles, tiles_dev, n_items_dev);
    return ctx->i8_bias ? launch_tensor<TM_I8P>(ctx, sl, n_tiles, ctx->tensor_kblocks, tiles_dev, n_items_dev)
                        : launch_tensor<TM_I8>(ctx, sl, n_tiles, ctx->tensor_kblocks, tiles_dev, n_items_dev);
}

// Arbitrary float data: TF32 ranking pass -> candidate collection (same tiles) -> exact refinement.
cudaError_t launch_tensor_refine(SfmmCtx* ctx, Slot& sl, uint32_t n_tiles, int kblocks) {
    const uint32_t rows = static_cast<uint32_t>(sl.plan.pair_of_row.size());
    cudaError_t e = cudaMemsetAsync(sl.d_cand_count.p, 0, std::max<size_t>(1, rows) * 2 * sizeof(uint32_t), sl.stream);
    if (e != cudaSuccess) return e;
    if (ctx->tensor_f16) {
        if ((e = launch_tensor<TM_F16_RANK>(ctx, sl, n_tiles, kblocks)) != cudaSuccess) return e;
        if ((e = launch_tensor<TM_F16_COLLECT>(ctx, sl, n_tiles, kblocks)) != cudaSuccess) return e;
    } else {
        if ((e = launch_tensor<TM_TF32_RANK>(ctx, sl, n_tiles, kblocks)) != cudaSuccess) return e;
        if ((e = launch_tensor<TM_TF32_COLLECT>(ctx, sl, n_tiles, kblocks)) != cudaSuccess) return e;
    }
    if (rows)
        float_refine_kernel<<<(rows + 63) / 64, 256, 0, sl.stream>>>(ctx->blob.as<float>(), static_cast<int>(ctx->pitch / 16), sl.d_pairs.as<PairDesc>(),
                                                                  static_cast<uint32_t>(sl.plan.pairs.size()), sl.d_pair_of_row.as<uint32_t>(),
                                                                  sl.d_cand_count.as<uint32_t>(), sl.d_cand_idx.as<uint32_t>(),
                                                                  sl.d_knn.as<KnnEntry>(), rows);
    ctx->stats.kernel_launches += 2;
    if (std::getenv("SFMM_DEBUG_CAND") && rows) {  // development aid: candidate-list statistics of this launch
        std::vector<uint32_t> h(2 * static_cast<size_t>(rows));
        cudaMemcpyAsync(h.data(), sl.d_cand_count.p, h.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, sl.stream);
        cudaStreamSynchronize(sl.stream);
        double sum = 0;
        uint32_t mx = 0, over = 0;
        for (uint32_t r = 0; r < rows; ++r) {
            const uint32_t a = h[2 * r], b = h[2 * r + 1];
            sum += a + b; mx = std::max(mx, a + b); over += a > FT_CAND_CAP / 2 || b > FT_CAND_CAP / 2;
        }
        std::fprintf(stderr, "[sfmm] candidates per row: mean %.2f max %u, rows over capacity %u of %u\n", sum / rows, mx, over, rows);
    }
    return cudaGetLastError();
}

typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 2-D tensor map over a row-major matrix of `row_bytes`-byte rows: boxes of 128 bytes x 128 rows, 128-byte swizzle.
int make_tensor_map(SfmmCtx* ctx, void* base, CUtensorMapDataType dtype, size_t elem_bytes, size_t row_bytes, uint64_t rows) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CU_TRY(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail(ctx, SFMM_ECUDA, "cuTensorMapEncodeTiled is not available in this driver");
    const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(row_bytes / elem_bytes), static_cast<cuuint64_t>(rows)};
    const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(row_bytes)};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / elem_bytes), FT_BOX_ROWS};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = reinterpret_cast<TensorMapEncodeFn>(fn)(&ctx->tmap, dtype, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, SFMM_ECUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return SFMM_OK;
}

// Tensor engine for Hamming (SFMM_BINARY_TENSOR / AUTO): unpack the bit rows to bytes once per descriptor set
// (lazily, like prepare_float).
int prepare_binary_tensor(SfmmCtx* ctx) {
    if (ctx->elem_type != SFMM_U8 || ctx->float_prepared) return SFMM_OK;
    ctx->use_tensor = false;
    ctx->tensor_f4 = false;
    ctx->tensor_f4x = false;
    const int words = binary_words(ctx->cols);
    const int kbytes8 = (words * 32 + 127) / 128 * 128;  // unpacked row: one byte per bit, whole 128-byte K-blocks
    // descriptors below 512 bit (packed 16-bit keys apply): one NIBBLE per bit and the FP4 pipe (TM_F4P) -- half the bytes, twice the rate
    const bool f4 = ctx->cols * 8 < 512 && !ctx->no_f4 && !std::getenv("SFMM_I8_KEYS32");
    const int kbytes = f4 ? (words * 16 + 127) / 128 * 128 : kbytes8;
    // TM_F4X pays once a row's threshold has settled, i.e. on long train images: measured (one B200, 486 bit) -4 % at 5 000 rows per
    // image, +13 % at 10 000 (profiles/tensor_variants_r02.txt); SFMM_F4X=1 / SFMM_NO_F4X=1 force either.  It keeps TWO operand arrays.
    const int64_t max_rows = ctx->rows.empty() ? 0 : *std::max_element(ctx->rows.begin(), ctx->rows.end());
    const char* f4x_env = std::getenv("SFMM_F4X");
    bool f4x = f4 && ctx->cols * 8 + 17 <= kbytes * 2 && !ctx->no_skip && !std::getenv("SFMM_NO_F4X") &&
               ((f4x_env && std::atoi(f4x_env) != 0) || max_rows >= 7000);
    bool want = ctx->cfg.binary_engine == SFMM_BINARY_TENSOR;
    if (want && kbytes8 > 512) return fail(ctx, SFMM_EINVAL, "SFMM_BINARY_TENSOR supports descriptors of at most 512 bits");
    if (ctx->cfg.binary_engine == SFMM_BINARY_AUTO && kbytes8 <= 512 && ctx->total_rows > 0) {
        size_t free_b = 0, total_b = 0;
        CU_TRY(ctx, cudaMemGetInfo(&free_b, &total_b));
        const size_t one = static_cast<size_t>(ctx->total_rows) * kbytes, have = free_b / 2 + ctx->d_unpacked.cap + ctx->d_f4_a.cap;  // (buffers we already own count as free)
        if (f4x && 2 * one > have) f4x = false;  // the second array does not fit: TM_F4P needs one
        want = one <= have;
    }
    if (want) {
        if (ctx->total_rows > 0) {
            cudaStream_t st = ctx->slot[0].stream;
            CU_TRY(ctx, ctx->d_unpacked.ensure(static_cast<size_t>(ctx->total_rows) * kbytes));
            CU_TRY(ctx, ctx->d_norms.ensure((static_cast<size_t>(ctx->total_rows) + 2 * FT_N) * sizeof(int32_t)));
            const uint32_t rows = static_cast<uint32_t>(ctx->total_rows);
            if (f4x) {
                CU_TRY(ctx, ctx->d_f4_a.ensure(static_cast<size_t>(ctx->total_rows) * kbytes));
                binary_unpack4x_kernel<<<(rows + 7) / 8, 256, 0, st>>>(ctx->blob.as<uint32_t>(), words, rows, kbytes, ctx->cols * 8, ctx->d_unpacked.as<uint8_t>(),
                                                                       ctx->d_f4_a.as<uint8_t>(), ctx->d_norms.as<int32_t>());
            } else if (f4)
                binary_unpack4_kernel<<<(rows + 7) / 8, 256, 0, st>>>(ctx->blob.as<uint32_t>(), words, rows, kbytes, ctx->d_unpacked.as<uint8_t>(),
                                                                      ctx->d_norms.as<int32_t>());
            else
                binary_unpack_kernel<<<(rows + 7) / 8, 256, 0, st>>>(ctx->blob.as<uint32_t>(), words, rows, kbytes, ctx->d_unpacked.as<uint8_t>(),
                                                                     ctx->d_norms.as<int32_t>());
            CU_TRY(ctx, cudaGetLastError());
            // query-independent part of the top-2 key per train row (popcount, bias, column inside its 128-row tile)
            CU_TRY(ctx, ctx->d_nbkey.ensure((static_cast<size_t>(ctx->total_rows) + 2 * FT_N) * sizeof(uint32_t)));
            CU_TRY(ctx, ctx->d_row0.ensure((static_cast<size_t>(ctx->n_images) + 1) * sizeof(uint32_t)));
            CU_TRY(ctx, cudaMemcpyAsync(ctx->d_row0.p, ctx->row0.data(), static_cast<size_t>(ctx->n_images) * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
            ctx->i8_bias = ctx->cols * 8 < 512 && !std::getenv("SFMM_I8_KEYS32") ? static_cast<uint32_t>(ctx->cols) * 8u : 0u;
            binary_nbkey_kernel<<<(rows + 255) / 256, 256, 0, st>>>(ctx->d_norms.as<int32_t>(), ctx->d_row0.as<uint32_t>(), ctx->n_images, rows,
                                                                     ctx->d_nbkey.as<uint32_t>(), ctx->i8_bias, f4 ? 0x80000000u : 0u);
            ctx->tensor_f4 = f4;
            ctx->tensor_f4x = f4x;
            CU_TRY(ctx, cudaGetLastError());
            CU_TRY(ctx, cudaStreamSynchronize(st));
            ctx->stats.kernel_launches += 2;
            int rc = make_tensor_map(ctx, ctx->d_unpacked.p, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, static_cast<size_t>(kbytes), ctx->total_rows);
            if (rc) return rc;
            ctx->tensor_kblocks = kbytes / 128;
            ctx->use_tensor = true;
        }
    }
    ctx->float_prepared = true;
    ctx->stats.float_path = ctx->use_tensor ? SFMM_FLOAT_TENSOR : 0;
    ctx->stats.tensor_kind = ctx->use_tensor ? (ctx->tensor_f4 ? 2 : 1) : 0;
    return SFMM_OK;
}

// Float descriptors only, once per descriptor set (lazily, so that a blob filled by an NCCL
// broadcast is seen): row norms, the TF32-exactness proof and the TMA tensor map; picks the path.
int prepare_float(SfmmCtx* ctx) {
    if (ctx->elem_type == SFMM_U8) return prepare_binary_tensor(ctx);
    if (ctx->elem_type != SFMM_F32 || ctx->float_prepared) return SFMM_OK;
    cudaStream_t st = ctx->slot[0].stream;
    ctx->tensor_eligible = false;
    ctx->use_tensor = false;
    const bool shape_ok = ctx->cols % FT_KB_ELEMS == 0 && ctx->cols <= 4 * FT_KB_ELEMS && ctx->total_rows > 0;
    if (ctx->cfg.float_mode != SFMM_FLOAT_EXACT && shape_ok) {
        CU_TRY(ctx, ctx->d_norms.ensure((static_cast<size_t>(ctx->total_rows) + 2 * FT_N) * sizeof(float)));  // + tail for the bulk copies
        CU_TRY(ctx, ctx->d_flags.ensure(2 * sizeof(unsigned int)));
        CU_TRY(ctx, cudaMemsetAsync(ctx->d_flags.p, 0, 2 * sizeof(unsigned int), st));
        const uint32_t rows = static_cast<uint32_t>(ctx->total_rows);
        float_prepare_kernel<<<(rows + 7) / 8, 256, 0, st>>>(ctx->blob.as<float>(), static_cast<int>(ctx->pitch / 16), rows, ctx->cols,
                                                             ctx->d_norms.as<float>(), ctx->d_flags.as<unsigned int>());
        CU_TRY(ctx, cudaGetLastError());
        ctx->stats.kernel_launches += 1;
        unsigned int flags[2] = {1, 0};
        CU_TRY(ctx, cudaMemcpyAsync(flags, ctx->d_flags.p, sizeof(flags), cudaMemcpyDeviceToHost, st));
        CU_TRY(ctx, cudaStreamSynchronize(st));
        float max_norm2;
        std::memcpy(&max_norm2, &flags[1], sizeof(float));
        ctx->tensor_eligible = flags[0] == 0 && max_norm2 <= 1048576.f;  // integers, |v|<=2047, |x|^2 <= 2^20
        ctx->tensor_refine = false;
        const bool finite = std::isfinite(max_norm2);  // NaN / inf rows: leave those sets to the exact kernel
        ctx->tensor_f16 = false;
        if (ctx->tensor_eligible) {  // per train row: |t|^2 + 2^23 + 2^20, the epilogue's key argument (float_nbexact_kernel)
            const uint32_t n = rows + 2 * FT_N;
            CU_TRY(ctx, ctx->d_nbkey.ensure(static_cast<size_t>(n) * sizeof(float)));
            float_nbexact_kernel<<<(n + 255) / 256, 256, 0, st>>>(ctx->d_norms.as<float>(), n, FT_NB_OFFSET, ctx->d_nbkey.as<float>());
            CU_TRY(ctx, cudaGetLastError());
            ctx->stats.kernel_launches += 1;
        }
        ctx->tensor_kx = false;
        if (ctx->tensor_eligible && ctx->cols % 64 == 0 && !std::getenv("SFMM_NO_F16")) {
            // TF32-exact data is fp16-exact too: contract an fp16 copy with kind::f16 (16 elements per MMA instead of 8)
            const bool kx = !ctx->no_kx && ctx->cols <= 128;  // TM_F16X: one more K-block carries the key's train-side term (<= 3 K-blocks)
            const int kx_cols = ctx->cols + (kx ? 64 : 0);
            CU_TRY(ctx, ctx->d_half.ensure(static_cast<size_t>(ctx->total_rows) * kx_cols * sizeof(__half)));
            const size_t n2 = static_cast<size_t>(ctx->total_rows) * kx_cols / 2;
            if (kx) {
                CU_TRY(ctx, ctx->d_half_a.ensure(static_cast<size_t>(ctx->total_rows) * kx_cols * sizeof(__half)));
                float_to_half_kx_kernel<<<static_cast<unsigned>((n2 + 255) / 256), 256, 0, st>>>(ctx->blob.as<float>(), static_cast<int>(ctx->pitch / 16), rows,
                                                                                              ctx->cols, ctx->d_norms.as<float>(), ctx->d_half.as<__half>(),
                                                                                              ctx->d_half_a.as<__half>());
            } else {
                float_to_half_kernel<<<static_cast<unsigned>((n2 + 255) / 256), 256, 0, st>>>(ctx->blob.as<float>(), static_cast<int>(ctx->pitch / 16), rows, ctx->cols,
                                                                                           ctx->d_half.as<__half>());
            }
            CU_TRY(ctx, cudaGetLastError());
            CU_TRY(ctx, cudaStreamSynchronize(st));
            ctx->stats.kernel_launches += 1;
            int rc = make_tensor_map(ctx, ctx->d_half.p, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, static_cast<size_t>(kx_cols) * 2, ctx->total_rows);
            if (rc) return rc;
            ctx->tensor_kblocks = kx_cols * 2 / 128;
            ctx->use_tensor = true;
            ctx->tensor_f16 = true;
            ctx->tensor_kx = kx;
        } else if (ctx->tensor_eligible || finite) {
            // arbitrary floats whose magnitudes fit fp16 (|v| <= sqrt(max |x|^2) < 65504): the ranking and collection passes
            // contract an fp16 round-to-nearest copy (same 10-bit significand as TF32, half the MMA time, no power cap);
            // the refinement reads the fp32 blob either way
            const bool f16_rank = !ctx->tensor_eligible && ctx->cols % 64 == 0 && max_norm2 < 4.0e9f && !std::getenv("SFMM_NO_F16");
            int rc;
            if (f16_rank) {
                CU_TRY(ctx, ctx->d_half.ensure(static_cast<size_t>(ctx->total_rows) * ctx->cols * sizeof(__half)));
                const size_t n2 = static_cast<size_t>(ctx->total_rows) * ctx->cols / 2;
                float_to_half_kernel<<<static_cast<unsigned>((n2 + 255) / 256), 256, 0, st>>>(ctx->blob.as<float>(), static_cast<int>(ctx->pitch / 16), rows,
                                                                                           ctx->cols, ctx->d_half.as<__half>());
                CU_TRY(ctx, cudaGetLastError());
                ctx->stats.kernel_launches += 1;
                rc = make_tensor_map(ctx, ctx->d_half.p, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, static_cast<size_t>(ctx->cols) * 2, ctx->total_rows);
                ctx->tensor_kblocks = ctx->cols * 2 / 128;
                ctx->tensor_f16 = true;
            } else {
                rc = make_tensor_map(ctx, ctx->blob.p, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, ctx->pitch, ctx->total_rows);
                ctx->tensor_kblocks = ctx->cols / FT_KB_ELEMS;
            }
            if (rc) return rc;
            ctx->use_tensor = true;
            if (!ctx->tensor_eligible) {
                // arbitrary floats: the tensor passes only rank; per-image max |x|^2 feeds the error bound, and the
                // ranking pass reads |t|^2 + C (C = the set's max |x|^2) from the key table
                ctx->tensor_refine = true;
                ctx->rank_offset = max_norm2;
                {
                    const uint32_t n = rows + 2 * FT_N;
                    CU_TRY(ctx, ctx->d_nbkey.ensure(static_cast<size_t>(n) * sizeof(float)));
                    float_nbexact_kernel<<<(n + 255) / 256, 256, 0, st>>>(ctx->d_norms.as<float>(), n, max_norm2, ctx->d_nbkey.as<float>());
                    CU_TRY(ctx, cudaGetLastError());
                    ctx->stats.kernel_launches += 1;
                }
                std::vector<float> h(static_cast<size_t>(ctx->total_rows));
                CU_TRY(ctx, cudaMemcpyAsync(h.data(), ctx->d_norms.p, h.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
                CU_TRY(ctx, cudaStreamSynchronize(st));
                ctx->stats.d2h_bytes += static_cast<int64_t>(h.size() * sizeof(float));
                ctx->img_maxnorm2.assign(static_cast<size_t>(ctx->n_images), 0.f);
                for (int32_t i = 0; i < ctx->n_images; ++i)
                    for (int32_t r = 0; r < ctx->rows[i]; ++r) ctx->img_maxnorm2[i] = std::max(ctx->img_maxnorm2[i], h[ctx->row0[i] + r]);
            }
        }
    }
    if (ctx->cfg.float_mode == SFMM_FLOAT_TENSOR && !ctx->use_tensor)
        return fail(ctx, SFMM_EINVAL,
                    "SFMM_FLOAT_TENSOR needs a descriptor width that is a multiple of 32 up to 128 and finite values; use SFMM_FLOAT_AUTO or SFMM_FLOAT_EXACT");
    ctx->float_prepared = true;
    ctx->stats.float_path = ctx->use_tensor ? (ctx->tensor_refine ? 3 : SFMM_FLOAT_TENSOR) : SFMM_FLOAT_EXACT;
    ctx->stats.tensor_kind = ctx->use_tensor ? (ctx->tensor_f16 ? 3 : 4) : 0;
    return SFMM_OK;
}

template <bool IS_FLOAT, bool CROSS>
cudaError_t launch_filter(SfmmCtx* ctx, Slot& sl, int32_t* d_counts, SfmDMatch* d_matches, uint64_t capacity, bool with_points) {
    const uint32_t nft = static_cast<uint32_t>(sl.plan.ftiles.size());
    const float ratio = ctx->cfg.ratio;
    filter_count_kernel<IS_FLOAT, CROSS><<<nft, FILTER_THREADS, 0, sl.stream>>>(
        sl.d_ftiles.as<FilterTile>(), sl.d_pairs.as<PairDesc>(), sl.d_knn.as<KnnEntry>(), sl.d_colmin.as<unsigned long long>(), ratio,
        sl.d_tile_count.as<uint32_t>());
    tile_scan_kernel<<<1, SCAN_THREADS, 0, sl.stream>>>(sl.d_tile_count.as<uint32_t>(), sl.d_tile_off.as<unsigned long long>(), nft);
    filter_write_kernel<IS_FLOAT, CROSS><<<nft, FILTER_THREADS, 0, sl.stream>>>(
        sl.d_ftiles.as<FilterTile>(), sl.d_pairs.as<PairDesc>(), sl.d_knn.as<KnnEntry>(), sl.d_colmin.as<unsigned long long>(), ratio,
        sl.d_tile_off.as<unsigned long long>(), d_matches, capacity, d_counts, sl.d_pair_off.as<unsigned long long>(),
        with_points ? ctx->d_points.as<double2>() : nullptr, sl.d_left.as<double2>(), sl.d_right.as<double2>());
    return cudaGetLastError();
}

// Plans `n` pairs and enqueues everything for them on the slot's stream, without waiting:
// plan upload, 2-NN kernel, ratio/cross-check/compaction into (d_counts, d_matches) -- the
// slot's own buffers when NULL --, and the device->host copy of the chunk's metadata.
//   knn_only: stop after the 2-NN kernel (sfmm_knn_pair);  force_single_split: plan nt==1 pairs too.
int launch_chunk(SfmmCtx* ctx, Slot& sl, const int32_t* qt, int64_t first, int64_t n, int32_t* d_counts, SfmDMatch* d_matches,
                 uint64_t capacity, bool knn_only = false, bool force_single_split = false) {
    int rc = plan_chunk(ctx, qt + 2 * first, n, sl.plan);
    if (rc) return rc;
    ChunkPlan& plan = sl.plan;
    if (force_single_split && plan.tiles.empty() && n == 1 && plan.pairs[0].nq > 0 && plan.pairs[0].nt > 0) {
        PairDesc& pd = plan.pairs[0];  // nt == 1: planned as "no matches"; the raw list still has one neighbour
        pd.n_splits = 1;
        const uint32_t q_tile = query_tile_rows(ctx);
        for (uint32_t q0 = 0; q0 < pd.nq; q0 += q_tile) plan.tiles.push_back(KnnTile{0, q0, 0, pd.nt, 0});
        plan.knn_entries = pd.nq;
        plan.col_entries = pd.nt;
        if (ctx->tensor_refine) {
            pd.q_off = 0;
            plan.pair_of_row.assign(pd.nq, 0u);
        }
    }
    sl.first = first;
    sl.n = n;
    sl.busy = true;
    if (ctx->blob_async) CU_TRY(ctx, cudaStreamWaitEvent(sl.stream, ctx->ev_blob, 0));  // sfmm_set_descriptors does not wait for its upload
    if (sl.copy_pending) {  // the previous chunk's records are still leaving this slot's buffers on the copy stream
        CU_TRY(ctx, cudaStreamWaitEvent(sl.stream, sl.ev_copied, 0));
        sl.copy_pending = false;
    }
    const bool cross = ctx->cfg.cross_check != 0;
    const size_t np = plan.pairs.size(), nft = plan.ftiles.size();
    CU_TRY(ctx, sl.d_pairs.ensure(std::max<size_t>(1, np) * sizeof(PairDesc)));
    CU_TRY(ctx, sl.d_tiles.ensure(std::max<size_t>(1, plan.tiles.size()) * sizeof(KnnTile)));
    CU_TRY(ctx, sl.d_ftiles.ensure(std::max<size_t>(1, nft) * sizeof(FilterTile)));
    CU_TRY(ctx, sl.d_knn.ensure(std::max<uint64_t>(1, plan.knn_entries) * sizeof(KnnEntry)));
    CU_TRY(ctx, sl.d_colmin.ensure(std::max<uint64_t>(1, cross ? plan.col_entries : 1) * sizeof(unsigned long long)));
    if (np) CU_TRY(ctx, cudaMemcpyAsync(sl.d_pairs.p, plan.pairs.data(), np * sizeof(PairDesc), cudaMemcpyHostToDevice, sl.stream));
    if (!plan.tiles.empty())
        CU_TRY(ctx, cudaMemcpyAsync(sl.d_tiles.p, plan.tiles.data(), plan.tiles.size() * sizeof(KnnTile), cudaMemcpyHostToDevice, sl.stream));
    if (nft) CU_TRY(ctx, cudaMemcpyAsync(sl.d_ftiles.p, plan.ftiles.data(), nft * sizeof(FilterTile), cudaMemcpyHostToDevice, sl.stream));
    ctx->stats.h2d_bytes += np * sizeof(PairDesc) + plan.tiles.size() * sizeof(KnnTile) + nft * sizeof(FilterTile);
    if (cross && plan.col_entries)
        CU_TRY(ctx, cudaMemsetAsync(sl.d_colmin.p, 0xFF, plan.col_entries * sizeof(unsigned long long), sl.stream));
    if (ctx->tensor_refine) {
        const size_t rows = std::max<size_t>(1, plan.pair_of_row.size());
        CU_TRY(ctx, sl.d_cand_count.ensure(rows * 2 * sizeof(uint32_t)));
        CU_TRY(ctx, sl.d_cand_idx.ensure(rows * FT_CAND_CAP * sizeof(uint32_t)));
        CU_TRY(ctx, sl.d_pair_of_row.ensure(rows * sizeof(uint32_t)));
        if (!plan.pair_of_row.empty())
            CU_TRY(ctx, cudaMemcpyAsync(sl.d_pair_of_row.p, plan.pair_of_row.data(), plan.pair_of_row.size() * sizeof(uint32_t),
                                        cudaMemcpyHostToDevice, sl.stream));
        ctx->stats.h2d_bytes += static_cast<int64_t>(plan.pair_of_row.size() * sizeof(uint32_t));
    }
    CU_TRY(ctx, cudaEventRecord(sl.ev_knn0, sl.stream));
    if (!plan.tiles.empty()) {
        const uint32_t nt = static_cast<uint32_t>(plan.tiles.size());
        cudaError_t e;
        if (ctx->elem_type == SFMM_F32 && ctx->use_tensor && ctx->tensor_refine) e = launch_tensor_refine(ctx, sl, nt, ctx->tensor_kblocks);
        else if (ctx->use_tensor) e = launch_tensor_exact(ctx, sl, nt);
        else if (ctx->elem_type == SFMM_F32) e = launch_float_exact(ctx, sl, nt);
        else e = cross ? launch_binary<true>(ctx, sl, nt) : launch_binary<false>(ctx, sl, nt);
        CU_TRY(ctx, e);
        ctx->stats.kernel_launches += 1;
        const bool by_exact = cross_by_candidates_exact(ctx);
        if ((cross_by_candidates(ctx) || by_exact) && !knn_only && nft && plan.max_rtiles) {
            // Symmetric cross-check through candidate columns (filter.cuh): flag the train rows that ratio-passing query rows chose,
            // compact them per pair, and run the 2-NN kernel again on "reverse" items whose rows are gathered through those lists.
            CU_TRY(ctx, sl.d_xflags.ensure(plan.col_entries));
            CU_TRY(ctx, sl.d_xcand.ensure(plan.col_entries * sizeof(uint32_t)));
            CU_TRY(ctx, sl.d_n_xcand.ensure(np * sizeof(uint32_t)));
            CU_TRY(ctx, sl.d_rtiles.ensure(plan.max_rtiles * sizeof(KnnTile)));
            CU_TRY(ctx, sl.d_n_rtiles.ensure(sizeof(uint32_t)));
            CU_TRY(ctx, cudaMemsetAsync(sl.d_xflags.p, 0, plan.col_entries, sl.stream));
            CU_TRY(ctx, cudaMemsetAsync(sl.d_n_rtiles.p, 0, sizeof(uint32_t), sl.stream));
            const float ratio = ctx->cfg.ratio;
            if (ctx->elem_type == SFMM_F32)
                cross_mark_kernel<true><<<static_cast<unsigned>(nft), FILTER_THREADS, 0, sl.stream>>>(sl.d_ftiles.as<FilterTile>(), sl.d_pairs.as<PairDesc>(),
                                                                                                     sl.d_knn.as<KnnEntry>(), ratio, sl.d_xflags.as<unsigned char>());
            else
                cross_mark_kernel<false><<<static_cast<unsigned>(nft), FILTER_THREADS, 0, sl.stream>>>(sl.d_ftiles.as<FilterTile>(), sl.d_pairs.as<PairDesc>(),
                                                                                                      sl.d_knn.as<KnnEntry>(), ratio, sl.d_xflags.as<unsigned char>());
            cross_compact_kernel<<<static_cast<unsigned>(np), COMPACT_THREADS, 0, sl.stream>>>(sl.d_pairs.as<PairDesc>(), sl.d_xflags.as<unsigned char>(),
                                                                                               sl.d_xcand.as<uint32_t>(), sl.d_n_xcand.as<uint32_t>(),
                                                                                               sl.d_rtiles.as<KnnTile>(), sl.d_n_rtiles.as<uint32_t>(), by_exact ? FX_BQ : FT_M);
            CU_TRY(ctx, cudaGetLastError());
            const uint32_t max_items = static_cast<uint32_t>(std::min<uint64_t>(plan.max_rtiles, 0xFFFFFFFFull));
            CU_TRY(ctx, by_exact ? launch_float_exact(ctx, sl, max_items, sl.d_rtiles.as<KnnTile>(), sl.d_n_rtiles.as<uint32_t>())
                                 : launch_tensor_exact(ctx, sl, max_items, sl.d_rtiles.as<KnnTile>(), sl.d_n_rtiles.as<uint32_t>()));
            ctx->stats.kernel_launches += 3;
        }
    }
    CU_TRY(ctx, cudaEventRecord(sl.ev_knn1, sl.stream));
    if (knn_only) {
        CU_TRY(ctx, cudaEventRecord(sl.ev_done, sl.stream));
        return SFMM_OK;
    }
    // ---- ratio test + cross-check + compaction
    const bool own = d_counts == nullptr;
    if (own) {
        CU_TRY(ctx, sl.d_pair_count.ensure(std::max<size_t>(1, np) * sizeof(int32_t)));
        CU_TRY(ctx, sl.d_matches.ensure(std::max<uint64_t>(1, plan.max_matches) * sizeof(SfmDMatch)));
        d_counts = sl.d_pair_count.as<int32_t>();
        d_matches = sl.d_matches.as<SfmDMatch>();
        capacity = plan.max_matches;
        if (ctx->have_points) {
            CU_TRY(ctx, sl.d_left.ensure(std::max<uint64_t>(1, plan.max_matches) * sizeof(double2)));
            CU_TRY(ctx, sl.d_right.ensure(std::max<uint64_t>(1, plan.max_matches) * sizeof(double2)));
        }
    }
    const bool with_points = own && ctx->have_points;
    CU_TRY(ctx, sl.d_pair_off.ensure(std::max<size_t>(1, np) * sizeof(unsigned long long)));
    CU_TRY(ctx, sl.meta.ensure(sizeof(unsigned long long) + np * (sizeof(unsigned long long) + sizeof(int32_t)) + 64));
    if (np) {
        CU_TRY(ctx, cudaMemsetAsync(d_counts, 0, np * sizeof(int32_t), sl.stream));
        CU_TRY(ctx, cudaMemsetAsync(sl.d_pair_off.p, 0, np * sizeof(unsigned long long), sl.stream));
    }
    unsigned long long* h_total = static_cast<unsigned long long*>(sl.meta.p);
    *h_total = 0;
    if (nft) {
        CU_TRY(ctx, sl.d_tile_count.ensure(nft * sizeof(uint32_t)));
        CU_TRY(ctx, sl.d_tile_off.ensure((nft + 1) * sizeof(unsigned long long)));
        const bool is_float = ctx->elem_type == SFMM_F32;
        cudaError_t e;
        if (is_float) e = cross ? launch_filter<true, true>(ctx, sl, d_counts, d_matches, capacity, with_points)
                                : launch_filter<true, false>(ctx, sl, d_counts, d_matches, capacity, with_points);
        else e = cross ? launch_filter<false, true>(ctx, sl, d_counts, d_matches, capacity, with_points)
                       : launch_filter<false, false>(ctx, sl, d_counts, d_matches, capacity, with_points);
        CU_TRY(ctx, e);
        ctx->stats.kernel_launches += 3;
        CU_TRY(ctx, cudaMemcpyAsync(h_total, sl.d_tile_off.as<unsigned long long>() + nft, sizeof(unsigned long long),
                                    cudaMemcpyDeviceToHost, sl.stream));
        ctx->stats.d2h_bytes += sizeof(unsigned long long);
    }
    if (own && np) {  // host table wanted: per-pair offsets and counts ride along
        unsigned long long* h_off = h_total + 1;
        int32_t* h_cnt = reinterpret_cast<int32_t*>(h_off + np);
        CU_TRY(ctx, cudaMemcpyAsync(h_off, sl.d_pair_off.p, np * sizeof(unsigned long long), cudaMemcpyDeviceToHost, sl.stream));
        CU_TRY(ctx, cudaMemcpyAsync(h_cnt, d_counts, np * sizeof(int32_t), cudaMemcpyDeviceToHost, sl.stream));
        ctx->stats.d2h_bytes += static_cast<int64_t>(np * (sizeof(unsigned long long) + sizeof(int32_t)));
    }
    CU_TRY(ctx, cudaEventRecord(sl.ev_done, sl.stream));
    return SFMM_OK;
}

// Waits for the chunk's kernels; *total = records produced (may exceed the capacity given: nothing
// past it was written).  Accounts the 2-NN kernel time.
int wait_chunk(SfmmCtx* ctx, Slot& sl, uint64_t* total) {
    CU_TRY(ctx, cudaEventSynchronize(sl.ev_done));
    if (total) *total = sl.meta.p ? *static_cast<unsigned long long*>(sl.meta.p) : 0;
    if (!sl.plan.tiles.empty()) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, sl.ev_knn0, sl.ev_knn1) == cudaSuccess) {
            ctx->stats.last_knn_ms += ms;
            ctx->stats.last_knn_work += sl.plan.work;
            ctx->stats.last_knn_launches += 1;
        } else {
            (void)cudaGetLastError();
        }
    }
    sl.busy = false;
    return SFMM_OK;
}

// Host-table path: bring the finished chunk's records to the host and index them.
int collect_chunk(SfmmCtx* ctx, Slot& sl, const int32_t* qt) {
    uint64_t total = 0;
    int rc = wait_chunk(ctx, sl, &total);
    if (rc) return rc;
    const size_t np = static_cast<size_t>(sl.n);
    const unsigned long long* h_off = static_cast<unsigned long long*>(sl.meta.p) + 1;
    const int32_t* h_cnt = reinterpret_cast<const int32_t*>(h_off + np);
    const size_t old_size = ctx->table.size();
    if (total) {
        // grow the table once per call where possible: extrapolate from the chunks seen so far
        if (ctx->table.capacity() < old_size + total) {
            const double seen = static_cast<double>(sl.first + sl.n), all = static_cast<double>(std::max<int64_t>(ctx->call_pairs, sl.first + sl.n));
            const size_t guess = static_cast<size_t>((static_cast<double>(old_size - ctx->call_base + total) * all / seen) * 1.05) + 1024;
            if (ctx->table.reserve(std::max(old_size + static_cast<size_t>(total), ctx->call_base + guess), ctx->copy_stream))
                return fail(ctx, SFMM_ENOMEM, ctx->table.shm_prefix.empty() ? "match_pairs: out of host memory for the match table"
                                                                            : "match_pairs: the shared-memory match table does not fit (/dev/shm too small?)");
        }
        // device -> host straight into the table, asynchronously: the next chunk's kernels (other slot) run meanwhile; this slot's
        // next launch waits for ev_copied, the call ends with a synchronisation of the copy stream
        CU_TRY(ctx, cudaMemcpyAsync(ctx->table.p + old_size, sl.d_matches.p, static_cast<size_t>(total) * sizeof(SfmDMatch), cudaMemcpyDeviceToHost,
                                    ctx->copy_stream));
        ctx->table.n = old_size + static_cast<size_t>(total);
        ctx->stats.d2h_bytes += static_cast<int64_t>(total * sizeof(SfmDMatch));
        if (ctx->have_points) {
            const size_t pb = static_cast<size_t>(total) * sizeof(double2);
            CU_TRY(ctx, sl.points.ensure(2 * pb));
            double2* stage = static_cast<double2*>(sl.points.p);
            CU_TRY(ctx, cudaMemcpyAsync(stage, sl.d_left.p, pb, cudaMemcpyDeviceToHost, ctx->copy_stream));
            CU_TRY(ctx, cudaMemcpyAsync(stage + total, sl.d_right.p, pb, cudaMemcpyDeviceToHost, ctx->copy_stream));
            CU_TRY(ctx, cudaStreamSynchronize(ctx->copy_stream));
            try {
                ctx->pts_left.insert(ctx->pts_left.end(), stage, stage + total);
                ctx->pts_right.insert(ctx->pts_right.end(), stage + total, stage + 2 * total);
            } catch (const std::bad_alloc&) {
                return fail(ctx, SFMM_ENOMEM, "match_pairs: out of host memory for the aligned points");
            }
            ctx->stats.d2h_bytes += static_cast<int64_t>(2 * pb);
        }
        CU_TRY(ctx, cudaEventRecord(sl.ev_copied, ctx->copy_stream));
        sl.copy_pending = true;
    }
    int64_t running = 0;
    for (size_t i = 0; i < np; ++i) {
        const int32_t q = qt[2 * (sl.first + i)], t = qt[2 * (sl.first + i) + 1];
        ctx->index[pair_key(q, t)] = static_cast<int64_t>(ctx->res_slots.size());
        ctx->res_qt.push_back(q);
        ctx->res_qt.push_back(t);
        ctx->res_counts.push_back(h_cnt[i]);
        // records are laid out in pair order, so the offset is the running total (h_off[i] for
        // non-empty pairs; empty pairs have no filter tile and get the position they would occupy)
        const int64_t local = h_cnt[i] ? static_cast<int64_t>(h_off[i]) : running;
        running = local + h_cnt[i];
        ctx->res_offsets.push_back(ctx->n_matches + local);
        ctx->res_slots.push_back(PairSlot{ctx->n_matches + local, h_cnt[i]});
    }
    ctx->n_matches += static_cast<int64_t>(total);
    return SFMM_OK;
}

// How many pairs of qt[from..n) fit a launch of at most `row_budget` query rows.
int64_t chunk_extent(const SfmmCtx* ctx, const int32_t* qt, int64_t from, int64_t n, uint64_t row_budget) {
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
constexpr uint64_t MAX_CHUNK_ROWS = 8u << 20;  // query rows per launch: 128 MB of 2-NN scratch, 128 MB of records

int require_descriptors(const SfmmCtx* ctx) {
    if (!ctx) return SFMM_EINVAL;
    if (ctx->elem_type < 0) return fail(ctx, SFMM_ESTATE, "sfmm_set_descriptors has not been called");
    return SFMM_OK;
}

int bind_device(const SfmmCtx* ctx) {
    CU_TRY(ctx, cudaSetDevice(ctx->cfg.device));
    return SFMM_OK;
}

int sync_all(SfmmCtx* ctx) {
    for (Slot& sl : ctx->slot) CU_TRY(ctx, cudaStreamSynchronize(sl.stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->copy_stream));
    return SFMM_OK;
}

void begin_stats(SfmmCtx* ctx) {
    ctx->stats.last_knn_ms = ctx->stats.last_knn_work = 0;
    ctx->stats.last_knn_launches = 0;
}

// Copies the caller's rows [r0, r1) -- counted over all images back to back, WITHOUT the blob's alignment rows -- into `dst`,
// tightly packed (row_bytes per row).  A cv::Mat is normally continuous (step == row_bytes): its share is one memcpy.  The
// re-pitch to 16-byte-multiple rows, the zero padding and the u8 -> fp32 widening happen on the GPU (ingest_rows_kernel).
void pack_raw(const SfmmCtx* ctx, const void* const* data, const size_t* step_bytes, size_t row_bytes, uint64_t r0, uint64_t r1,
              unsigned char* dst) {
    int32_t img = static_cast<int32_t>(std::upper_bound(ctx->raw_row0.begin(), ctx->raw_row0.end(), r0) - ctx->raw_row0.begin()) - 1;
    uint64_t r = r0;
    while (r < r1) {
        while (img + 1 < ctx->n_images && ctx->raw_row0[img + 1] <= r) ++img;
        const uint64_t local = r - ctx->raw_row0[img];
        const uint64_t n = std::min<uint64_t>(r1 - r, static_cast<uint64_t>(ctx->rows[img]) - local);
        const size_t step = step_bytes ? step_bytes[img] : row_bytes;
        const unsigned char* src = static_cast<const unsigned char*>(data[img]) + local * step;
        if (step == row_bytes || n == 1) {
            std::memcpy(dst, src, n * row_bytes);
        } else {
            for (uint64_t k = 0; k < n; ++k) std::memcpy(dst + k * row_bytes, src + k * step, row_bytes);
        }
        dst += n * row_bytes;
        r += n;
    }
}

// raw (tightly packed caller rows) -> blob rows: 16-byte-multiple pitch, zeroed padding, zero rows that align every image to a
// multiple of four rows; CV_8U rows under NORM_L2 are widened to fp32 here (exact).  One thread per 16 output bytes.
template <bool WIDEN>
__global__ void ingest_rows_kernel(const unsigned char* __restrict__ raw, const uint32_t* __restrict__ raw_row0 /* n_images + 1 */,
                                   const uint32_t* __restrict__ row0 /* n_images */, int n_images, uint32_t total_rows, uint32_t chunks_per_row,
                                   uint32_t row_bytes, uint4* __restrict__ blob) {
    const uint64_t idx = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const uint32_t r = static_cast<uint32_t>(idx / chunks_per_row), c = static_cast<uint32_t>(idx % chunks_per_row);
    if (r >= total_rows) return;
    int lo = 0, hi = n_images;  // last image whose first blob row is <= r (empty images share their successor's first row)
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (row0[mid] <= r) lo = mid; else hi = mid;
    }
    const uint32_t local = r - row0[lo];
    const bool valid = local < raw_row0[lo + 1] - raw_row0[lo];
    const unsigned char* src = raw + static_cast<size_t>(raw_row0[lo] + (valid ? local : 0)) * row_bytes;
    uint4 out = make_uint4(0, 0, 0, 0);
    if (valid) {
        if constexpr (WIDEN) {
            float f[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) f[k] = 4 * c + k < row_bytes ? static_cast<float>(src[4 * c + k]) : 0.f;
            out = make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
        } else {
            uint32_t w[4] = {0, 0, 0, 0};
            if ((row_bytes & 15u) == 0) {  // float rows, 32- and 64-byte binary rows: aligned 16-byte loads
                if (16 * c < row_bytes) blob[idx] = *reinterpret_cast<const uint4*>(src + 16 * c);
                else blob[idx] = out;
                return;
            }
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const uint32_t b = 16 * c + k;
                if (b < row_bytes) w[k >> 2] |= static_cast<uint32_t>(src[b]) << (8 * (k & 3));
            }
            out = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
    blob[idx] = out;
}

}  // namespace

// =============================================================================== C ABI
extern "C" {

SFMM_API const char* sfmm_version(void) { return "0.2.0 (sm_100a)"; }

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
    cfg->binary_engine = SFMM_BINARY_AUTO;
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
    if (cfg->binary_engine != SFMM_BINARY_AUTO && cfg->binary_engine != SFMM_BINARY_POPC && cfg->binary_engine != SFMM_BINARY_TENSOR)
        return fail(nullptr, SFMM_EINVAL, "sfmm_create: unknown binary_engine");
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
    if (const char* s = std::getenv("SFMM_TENSOR_TS")) ctx->tensor_ts = std::atoi(s) != 0 ? 1 : 0;
    if (const char* s = std::getenv("SFMM_CSA_LEVEL")) ctx->csa_level = std::max(0, std::min(3, std::atoi(s)));
    if (const char* s = std::getenv("SFMM_NO_KX")) ctx->no_kx = std::atoi(s) != 0;
    if (const char* s = std::getenv("SFMM_NO_SKIP")) ctx->no_skip = std::atoi(s) != 0;
    if (const char* s = std::getenv("SFMM_NO_F4")) ctx->no_f4 = std::atoi(s) != 0;
    if (const char* s = std::getenv("SFMM_CROSS_FULL")) ctx->cross_full_reverse = std::atoi(s) != 0;
    if (const char* s = std::getenv("SFMM_EPI_GROUPS")) { ctx->epi_groups = std::max(2, std::min(4, std::atoi(s))); ctx->epi_groups_set = true; }
    bool ok = (e = cudaSetDevice(cfg->device)) == cudaSuccess;
    for (Slot& sl : ctx->slot) {
        ok = ok && (e = cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking)) == cudaSuccess;
        ok = ok && (e = cudaEventCreate(&sl.ev_knn0)) == cudaSuccess && (e = cudaEventCreate(&sl.ev_knn1)) == cudaSuccess &&
             (e = cudaEventCreate(&sl.ev_done)) == cudaSuccess && (e = cudaEventCreateWithFlags(&sl.ev_copied, cudaEventDisableTiming)) == cudaSuccess;
    }
    ok = ok && (e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking)) == cudaSuccess;
    ok = ok && (e = cudaEventCreate(&ctx->ev_begin)) == cudaSuccess && (e = cudaEventCreate(&ctx->ev_end)) == cudaSuccess;
    for (auto& ev : ctx->ev_pack) ok = ok && (e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) == cudaSuccess;
    ok = ok && (e = cudaEventCreateWithFlags(&ctx->ev_blob, cudaEventDisableTiming)) == cudaSuccess;
    if (!ok) {
        sfmm_destroy(ctx);
        return fail(nullptr, SFMM_ECUDA, std::string("stream/event setup: ") + cudaGetErrorString(e));
    }
    *out = ctx;
    return SFMM_OK;
}

SFMM_API void sfmm_destroy(SfmmCtx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->cfg.device);
    for (Slot& sl : ctx->slot) {
        if (sl.stream) cudaStreamSynchronize(sl.stream);
        sl.release();
    }
    if (ctx->copy_stream) {
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamDestroy(ctx->copy_stream);
    }
    for (DevBuf* b : {&ctx->blob, &ctx->d_idx, &ctx->d_dist, &ctx->d_norms, &ctx->d_flags, &ctx->d_points, &ctx->d_unpacked, &ctx->d_nbkey, &ctx->d_row0, &ctx->d_half, &ctx->d_half_a, &ctx->d_f4_a, &ctx->d_raw, &ctx->d_raw_row0}) b->release();
    for (PinBuf& p : ctx->pack) p.release();
    ctx->table.release();
    for (cudaEvent_t ev : {ctx->ev_begin, ctx->ev_end, ctx->ev_pack[0], ctx->ev_pack[1], ctx->ev_blob})
        if (ev) cudaEventDestroy(ev);
    delete ctx;
}

SFMM_API size_t sfmm_row_pitch(int32_t cols, int32_t elem_type) {
    if (cols <= 0) return 0;
    if (elem_type == SFMM_U8) return static_cast<size_t>(binary_words(cols)) * 4;
    if (elem_type == SFMM_F32) return (static_cast<size_t>(cols) * 4 + 15) / 16 * 16;
    return 0;
}

static int impl_set_descriptors(SfmmCtx* ctx, int32_t n_images, const void* const* data, const int32_t* rows,
                                  int32_t cols, const size_t* step_bytes, int32_t elem_type) {
    if (!ctx) return SFMM_EINVAL;
    if (n_images < 0 || (n_images > 0 && !rows) || cols <= 0) return fail(ctx, SFMM_EINVAL, "set_descriptors: bad sizes");
    if (elem_type != SFMM_U8 && elem_type != SFMM_F32) return fail(ctx, SFMM_EINVAL, "set_descriptors: unknown elem_type");
    // cv::BFMatcher asserts NORM_HAMMING <-> CV_8U.  NORM_L2 takes both depths: on CV_8U rows it is what the reference
    // literally does for its AKAZE / ORB detectors (cv::BFMatcher(cv::NORM_L2) at src/Sfm.cpp:593 whatever the
    // detector; OpenCV's batchDistL2_8u32f = sqrtf of the exact integer sum).  Those rows are widened to fp32 on
    // upload (exact, zero-padded to a multiple of 32 columns) and take the float kernels.
    if (ctx->cfg.norm == SFMM_NORM_HAMMING && elem_type != SFMM_U8)
        return fail(ctx, SFMM_EINVAL, "set_descriptors: NORM_HAMMING needs SFMM_U8 rows");
    const bool widen = ctx->cfg.norm == SFMM_NORM_L2 && elem_type == SFMM_U8;
    const int32_t blob_type = widen ? SFMM_F32 : elem_type;
    const int32_t blob_cols = widen ? (cols + 31) / 32 * 32 : cols;
    const size_t pitch = sfmm_row_pitch(blob_cols, blob_type);
    if (pitch == 0) return fail(ctx, SFMM_EINVAL, "set_descriptors: unsupported descriptor width (binary <= 128 bytes)");
    if (blob_type == SFMM_F32 && blob_cols > 256) return fail(ctx, SFMM_EINVAL, "set_descriptors: float descriptors wider than 256 are not supported");
    const size_t elem = elem_type == SFMM_U8 ? 1 : 4;
    // every image starts at a blob row that is a multiple of 4 (16-byte aligned slices of the
    // per-row norm array for the TMA bulk copies of the tensor path); the rows in between are zero
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
    if ((rc = sync_all(ctx))) return rc;
    clear_results(ctx);
    ctx->elem_type = -1;
    ctx->src_elem_type = elem_type;
    ctx->elem_type_pending = blob_type;
    ctx->float_prepared = false;
    ctx->use_tensor = false;
    ctx->tensor_refine = false;
    ctx->have_points = false;
    ctx->stats.float_path = 0;
    ctx->stats.tensor_kind = 0;
    const size_t bytes = static_cast<size_t>(total) * pitch;
    CU_TRY(ctx, ctx->blob.ensure(std::max<size_t>(bytes, 16)));
    ctx->rows.assign(rows, rows + n_images);
    ctx->row0.resize(n_images);
    uint64_t r0 = 0;
    for (int32_t i = 0; i < n_images; ++i) {
        ctx->row0[i] = static_cast<uint32_t>(r0);
        r0 += (static_cast<uint64_t>(rows[i]) + 3) & ~3ull;
    }
    ctx->n_images = n_images;
    ctx->cols = blob_cols;
    ctx->src_cols = cols;
    ctx->pitch = pitch;
    ctx->total_rows = total;
    ctx->blob_bytes = bytes;
    ctx->raw_row0.assign(static_cast<size_t>(n_images) + 1, 0);
    for (int32_t i = 0; i < n_images; ++i) ctx->raw_row0[i + 1] = ctx->raw_row0[i] + static_cast<uint64_t>(rows[i]);
    ctx->blob_async = false;
    if (data && bytes) {
        // The caller's rows travel tightly packed: a few host threads copy them into two pinned slabs (one memcpy per image when the
        // Mat is continuous), each slab followed by its own async H2D so that filling slab k+1 overlaps the copy of slab k; one
        // kernel then re-pitches / pads / widens them into the blob.  Nothing here waits for the device: the blob's completion is an
        // event the matching streams wait for (ev_blob), so the tail of the upload overlaps the caller's next steps.
        cudaStream_t st = ctx->slot[0].stream;
        const size_t row_bytes = static_cast<size_t>(cols) * elem;
        const uint64_t raw_rows = ctx->raw_row0[n_images];
        CU_TRY(ctx, ctx->d_raw.ensure(std::max<size_t>(16, static_cast<size_t>(raw_rows) * row_bytes + 16)));
        CU_TRY(ctx, ctx->d_raw_row0.ensure((static_cast<size_t>(n_images) + 1) * sizeof(uint32_t)));
        CU_TRY(ctx, ctx->d_row0.ensure((static_cast<size_t>(n_images) + 1) * sizeof(uint32_t)));
        ctx->raw_row0_32.assign(ctx->raw_row0.begin(), ctx->raw_row0.end());
        CU_TRY(ctx, cudaMemcpyAsync(ctx->d_raw_row0.p, ctx->raw_row0_32.data(), ctx->raw_row0_32.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        CU_TRY(ctx, cudaMemcpyAsync(ctx->d_row0.p, ctx->row0.data(), static_cast<size_t>(n_images) * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        // Slabs of 8 MB: the first slab's fill and the last slab's copy are the only parts of the pipeline that do not overlap.  The
        // worker threads live for the whole call (one spawn per call, not per slab): slab k is announced through `gen`, every worker
        // copies its share and counts `left` down; the calling thread takes the first share, waits for the others and issues the copy.
        const uint64_t slab_rows = std::max<uint64_t>(1, (8u << 20) / row_bytes);
        const size_t slab_bytes = static_cast<size_t>(std::min<uint64_t>(slab_rows, raw_rows)) * row_bytes;
        for (PinBuf& p : ctx->pack) CU_TRY(ctx, p.ensure(slab_bytes));
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        unsigned workers = static_cast<unsigned>(std::min<uint64_t>(std::min(16u, hw), std::max<uint64_t>(1, (raw_rows * row_bytes) >> 21)));
        struct Slab { uint64_t b = 0, e = 0; unsigned char* dst = nullptr; };
        Slab cur;
        std::atomic<int64_t> gen{0};   // slab number + 1 that `cur` describes; -1: done
        std::atomic<int> left{0};      // workers that have not finished the current slab yet
        auto share = [&](const Slab& sl, unsigned w) {
            const uint64_t n_rows = sl.e - sl.b, per = (n_rows + workers - 1) / workers;
            const uint64_t wb = sl.b + w * per, we = std::min(sl.e, wb + per);
            if (wb < we) pack_raw(ctx, data, step_bytes, row_bytes, wb, we, sl.dst + (wb - sl.b) * row_bytes);
        };
        std::vector<std::thread> pool;
        try {
            pool.reserve(workers);
            for (unsigned w = 1; w < workers; ++w)
                pool.emplace_back([&, w]() {
                    int64_t seen = 0;
                    for (;;) {
                        int64_t g;
                        while ((g = gen.load(std::memory_order_acquire)) == seen) std::this_thread::yield();
                        if (g < 0) return;
                        seen = g;
                        share(cur, w);
                        left.fetch_sub(1, std::memory_order_release);
                    }
                });
        } catch (...) {  // out of threads: go on with the ones that started (nothing has been announced yet)
        }
        workers = static_cast<unsigned>(pool.size()) + 1;
        auto stop_pool = [&]() {
            gen.store(-1, std::memory_order_release);
            for (auto& t : pool) t.join();
        };
        int k = 0;
        for (uint64_t b = 0; b < raw_rows; b += slab_rows, ++k) {
            const uint64_t e = std::min(raw_rows, b + slab_rows);
            PinBuf& pin = ctx->pack[k & 1];
            if (k >= 2) {  // its previous copy has left the slab
                const cudaError_t ce = cudaEventSynchronize(ctx->ev_pack[k & 1]);
                if (ce != cudaSuccess) {
                    stop_pool();
                    CU_TRY(ctx, ce);
                }
            }
            cur = Slab{b, e, static_cast<unsigned char*>(pin.p)};
            left.store(static_cast<int>(pool.size()), std::memory_order_relaxed);
            gen.store(k + 1, std::memory_order_release);
            share(cur, 0);  // this thread takes the first share
            while (left.load(std::memory_order_acquire) != 0) std::this_thread::yield();
            const uint64_t n_rows = e - b;
            cudaError_t ce = cudaMemcpyAsync(static_cast<unsigned char*>(ctx->d_raw.p) + b * row_bytes, cur.dst, n_rows * row_bytes, cudaMemcpyHostToDevice, st);
            if (ce == cudaSuccess) ce = cudaEventRecord(ctx->ev_pack[k & 1], st);
            if (ce != cudaSuccess) {
                stop_pool();
                CU_TRY(ctx, ce);
            }
            ctx->stats.h2d_bytes += static_cast<int64_t>(n_rows * row_bytes);
        }
        stop_pool();
        const uint32_t chunks = static_cast<uint32_t>(pitch / 16);
        const uint64_t n_thr = total * chunks;
        const unsigned grid = static_cast<unsigned>((n_thr + 255) / 256);
        if (widen)
            ingest_rows_kernel<true><<<grid, 256, 0, st>>>(ctx->d_raw.as<unsigned char>(), ctx->d_raw_row0.as<uint32_t>(), ctx->d_row0.as<uint32_t>(), n_images,
                                                           static_cast<uint32_t>(total), chunks, static_cast<uint32_t>(row_bytes), ctx->blob.as<uint4>());
        else
            ingest_rows_kernel<false><<<grid, 256, 0, st>>>(ctx->d_raw.as<unsigned char>(), ctx->d_raw_row0.as<uint32_t>(), ctx->d_row0.as<uint32_t>(), n_images,
                                                            static_cast<uint32_t>(total), chunks, static_cast<uint32_t>(row_bytes), ctx->blob.as<uint4>());
        CU_TRY(ctx, cudaGetLastError());
        ctx->stats.kernel_launches += 1;
        CU_TRY(ctx, cudaEventRecord(ctx->ev_blob, st));
        ctx->blob_async = true;
    }
    ctx->elem_type = blob_type;
    return SFMM_OK;
}

SFMM_API int sfmm_image_rows(const SfmmCtx* ctx, int32_t image, int32_t* rows) {
    if (!ctx || !rows) return SFMM_EINVAL;
    if (ctx->elem_type < 0) return SFMM_ESTATE;
    if (image < 0 || image >= ctx->n_images) return SFMM_ERANGE;
    *rows = ctx->rows[static_cast<size_t>(image)];
    return SFMM_OK;
}

SFMM_API int sfmm_descriptor_blob(SfmmCtx* ctx, void** device_ptr, size_t* bytes) {
    int rc = require_descriptors(ctx);
    if (rc) return rc;
    if (!device_ptr || !bytes) return fail(ctx, SFMM_EINVAL, "descriptor_blob: NULL argument");
    if (ctx->blob_async) {  // the caller will touch the blob from its own streams
        CU_TRY(ctx, cudaEventSynchronize(ctx->ev_blob));
        ctx->blob_async = false;
    }
    *device_ptr = ctx->blob.p;
    *bytes = ctx->blob_bytes;
    ctx->float_prepared = false;  // the caller may overwrite the blob (broadcast): re-derive norms/eligibility
    return SFMM_OK;
}

static int impl_set_points(SfmmCtx* ctx, int32_t n_images, const double* const* xy) {
    int rc = require_descriptors(ctx);
    if (rc) return rc;
    if (n_images != ctx->n_images || (n_images > 0 && !xy)) return fail(ctx, SFMM_EINVAL, "set_points: image count differs from the descriptor set");
    for (int32_t i = 0; i < n_images; ++i)
        if (ctx->rows[i] > 0 && !xy[i]) return fail(ctx, SFMM_EINVAL, "set_points: NULL point list");
    if ((rc = bind_device(ctx))) return rc;
    if ((rc = sync_all(ctx))) return rc;
    clear_results(ctx);
    CU_TRY(ctx, ctx->d_points.ensure(std::max<uint64_t>(1, ctx->total_rows) * sizeof(double2)));
    cudaStream_t st = ctx->slot[0].stream;
    for (int32_t i = 0; i < n_images; ++i)
        if (ctx->rows[i] > 0) {
            CU_TRY(ctx, cudaMemcpyAsync(ctx->d_points.as<double2>() + ctx->row0[i], xy[i], static_cast<size_t>(ctx->rows[i]) * sizeof(double2),
                                        cudaMemcpyHostToDevice, st));
            ctx->stats.h2d_bytes += static_cast<int64_t>(ctx->rows[i]) * static_cast<int64_t>(sizeof(double2));
        }
    CU_TRY(ctx, cudaStreamSynchronize(st));
    ctx->have_points = true;
    return SFMM_OK;
}

SFMM_API int sfmm_get_pair_points(const SfmmCtx* ctx, int32_t q, int32_t t, const double** left_xy, const double** right_xy, int32_t* count) {
    if (!ctx) return SFMM_EINVAL;
    if (ctx->elem_type < 0) return SFMM_ESTATE;
    if (!left_xy || !right_xy || !count) return SFMM_EINVAL;
    if (q < 0 || q >= ctx->n_images || t < 0 || t >= ctx->n_images) return SFMM_ERANGE;
    auto it = ctx->index.find(pair_key(q, t));
    if (it == ctx->index.end()) return SFMM_ESTATE;
    if (ctx->pts_left.size() != ctx->table.size()) return SFMM_ESTATE;
    const PairSlot& s = ctx->res_slots[static_cast<size_t>(it->second)];
    *left_xy = reinterpret_cast<const double*>(ctx->pts_left.data() + s.offset);
    *right_xy = reinterpret_cast<const double*>(ctx->pts_right.data() + s.offset);
    *count = s.count;
    return SFMM_OK;
}

namespace {
struct TableHeader {  // 64 bytes, little endian
    char magic[8];    // "SFMMTBL1"
    int32_t n_images, norm, cross_check, elem_type;
    float ratio;
    int32_t cols;
    int64_t n_pairs, n_matches;
    int32_t reserved[4];
};
static_assert(sizeof(TableHeader) == 64, "table header is 64 bytes");
}  // namespace

SFMM_API int sfmm_save_table(const SfmmCtx* ctx, const char* path) {
    int rc = require_descriptors(ctx);
    if (rc) return rc;
    if (!path) return fail(ctx, SFMM_EINVAL, "save_table: NULL path");
    FILE* f = std::fopen(path, "wb");
    if (!f) return fail(ctx, SFMM_EINVAL, std::string("save_table: cannot open ") + path);
    TableHeader h{};
    std::memcpy(h.magic, "SFMMTBL1", 8);
    h.n_images = ctx->n_images; h.norm = ctx->cfg.norm; h.cross_check = ctx->cfg.cross_check; h.elem_type = ctx->src_elem_type;
    h.ratio = ctx->cfg.ratio; h.cols = ctx->src_cols;
    h.n_pairs = static_cast<int64_t>(ctx->res_counts.size()); h.n_matches = ctx->n_matches;
    bool ok = std::fwrite(&h, sizeof h, 1, f) == 1;
    auto put = [&](const void* p, size_t bytes) { ok = ok && (bytes == 0 || std::fwrite(p, 1, bytes, f) == bytes); };
    put(ctx->rows.data(), ctx->rows.size() * sizeof(int32_t));
    put(ctx->res_qt.data(), ctx->res_qt.size() * sizeof(int32_t));
    put(ctx->res_counts.data(), ctx->res_counts.size() * sizeof(int32_t));
    put(ctx->res_offsets.data(), ctx->res_offsets.size() * sizeof(int64_t));
    put(ctx->table.data(), ctx->table.size() * sizeof(SfmDMatch));
    ok = (std::fclose(f) == 0) && ok;
    return ok ? SFMM_OK : fail(ctx, SFMM_EINVAL, std::string("save_table: short write to ") + path);
}

SFMM_API int sfmm_load_table(SfmmCtx* ctx, const char* path) {
    int rc = require_descriptors(ctx);
    if (rc) return rc;
    if (!path) return fail(ctx, SFMM_EINVAL, "load_table: NULL path");
    return guarded(ctx, [&]() -> int {
        FILE* f = std::fopen(path, "rb");
        if (!f) return fail(ctx, SFMM_EINVAL, std::string("load_table: cannot open ") + path);
        struct Closer { FILE* f; ~Closer() { std::fclose(f); } } closer{f};
        TableHeader h{};
        if (std::fread(&h, sizeof h, 1, f) != 1 || std::memcmp(h.magic, "SFMMTBL1", 8) != 0)
            return fail(ctx, SFMM_EINVAL, std::string("load_table: not a match table: ") + path);
        if (h.n_images != ctx->n_images || h.norm != ctx->cfg.norm || h.elem_type != ctx->src_elem_type || h.cols != ctx->src_cols)
            return fail(ctx, SFMM_EINVAL, "load_table: the file belongs to a different descriptor set / norm");
        // a table is only valid for the filter settings it was computed with
        if (std::memcmp(&h.ratio, &ctx->cfg.ratio, sizeof(float)) != 0 || (h.cross_check != 0) != (ctx->cfg.cross_check != 0))
            return fail(ctx, SFMM_EINVAL, "load_table: the file was computed with a different ratio / cross_check setting");
        // the directory sizes come from the file: bound them by the file itself before allocating anything
        if (std::fseek(f, 0, SEEK_END) != 0) return fail(ctx, SFMM_EINVAL, "load_table: cannot seek");
        const long long file_bytes = std::ftell(f);
        if (std::fseek(f, static_cast<long>(sizeof h), SEEK_SET) != 0) return fail(ctx, SFMM_EINVAL, "load_table: cannot seek");
        const long long per_pair = 2 * sizeof(int32_t) + sizeof(int32_t) + sizeof(int64_t);
        const long long fixed = static_cast<long long>(sizeof h) + static_cast<long long>(h.n_images) * sizeof(int32_t);
        if (h.n_pairs < 0 || h.n_matches < 0 || file_bytes < fixed || h.n_pairs > (file_bytes - fixed) / per_pair ||
            h.n_matches > (file_bytes - fixed - h.n_pairs * per_pair) / static_cast<long long>(sizeof(SfmDMatch)) ||
            fixed + h.n_pairs * per_pair + h.n_matches * static_cast<long long>(sizeof(SfmDMatch)) != file_bytes)
            return fail(ctx, SFMM_EINVAL, std::string("load_table: header does not match the file size (truncated or corrupt): ") + path);
        std::vector<int32_t> rows(static_cast<size_t>(h.n_images)), qt(2 * static_cast<size_t>(h.n_pairs)), counts(static_cast<size_t>(h.n_pairs));
        std::vector<int64_t> offsets(static_cast<size_t>(h.n_pairs));
        std::vector<SfmDMatch> table(static_cast<size_t>(h.n_matches));
        bool ok = true;
        auto get = [&](void* p, size_t bytes) { ok = ok && (bytes == 0 || std::fread(p, 1, bytes, f) == bytes); };
        get(rows.data(), rows.size() * sizeof(int32_t));
        get(qt.data(), qt.size() * sizeof(int32_t));
        get(counts.data(), counts.size() * sizeof(int32_t));
        get(offsets.data(), offsets.size() * sizeof(int64_t));
        get(table.data(), table.size() * sizeof(SfmDMatch));
        if (!ok) return fail(ctx, SFMM_EINVAL, std::string("load_table: short read: ") + path);
        if (rows != ctx->rows) return fail(ctx, SFMM_EINVAL, "load_table: per-image row counts differ from the current descriptor set");
        for (int64_t i = 0; i < h.n_pairs; ++i) {
            const int32_t q = qt[2 * i], t = qt[2 * i + 1];
            if (q < 0 || q >= ctx->n_images || t < 0 || t >= ctx->n_images || counts[i] < 0 || offsets[i] < 0 ||
                offsets[i] > h.n_matches - counts[i])
                return fail(ctx, SFMM_EINVAL, "load_table: corrupt pair directory");
        }
        clear_results(ctx);
        ctx->res_qt = std::move(qt);
        ctx->res_counts = std::move(counts);
        ctx->res_offsets = std::move(offsets);
        if (ctx->table.reserve(table.size(), ctx->copy_stream)) return fail(ctx, SFMM_ENOMEM, "load_table: out of host memory");
        if (!table.empty()) std::memcpy(ctx->table.p, table.data(), table.size() * sizeof(SfmDMatch));
        ctx->table.n = table.size();
        ctx->n_matches = h.n_matches;
        ctx->res_slots.reserve(static_cast<size_t>(h.n_pairs));
        for (int64_t i = 0; i < h.n_pairs; ++i) {
            ctx->index[pair_key(ctx->res_qt[2 * i], ctx->res_qt[2 * i + 1])] = i;
            ctx->res_slots.push_back(PairSlot{ctx->res_offsets[i], ctx->res_counts[i]});
        }
        return SFMM_OK;
    });
}

SFMM_API int sfmm_share_table(SfmmCtx* ctx, const char* shm_prefix) {
    if (!ctx) return SFMM_EINVAL;
    if (shm_prefix && shm_prefix[0] && (shm_prefix[0] != '/' || std::strchr(shm_prefix + 1, '/') || std::strlen(shm_prefix) > 200))
        return fail(ctx, SFMM_EINVAL, "share_table: the prefix must be a POSIX shared-memory name: '/' + a name without further slashes");
    int rc = bind_device(ctx);
    if (rc) return rc;
    if ((rc = sync_all(ctx))) return rc;
    return guarded(ctx, [&]() -> int {
        clear_results(ctx);
        ctx->table.release();  // the next match call allocates the table in the new mode
        ctx->table.shm_prefix = shm_prefix ? shm_prefix : "";
        return SFMM_OK;
    });
}

SFMM_API int sfmm_shared_table_info(const SfmmCtx* ctx, char* name, size_t name_capacity, int64_t* n_records) {
    if (!ctx || !name || !n_records || name_capacity == 0) return SFMM_EINVAL;
    if (ctx->table.shm_prefix.empty()) return SFMM_ESTATE;
    if (ctx->table.shm_name.size() + 1 > name_capacity) return SFMM_ERANGE;
    std::memcpy(name, ctx->table.shm_name.c_str(), ctx->table.shm_name.size() + 1);  // empty while nothing has been matched yet
    *n_records = static_cast<int64_t>(ctx->table.size());
    return SFMM_OK;
}

SFMM_API int sfmm_clear_results(SfmmCtx* ctx) {
    if (!ctx) return SFMM_EINVAL;
    clear_results(ctx);
    return SFMM_OK;
}

static int impl_match_pairs_device(SfmmCtx* ctx, const int32_t* qt, int64_t n_pairs, int32_t* d_counts,
                                     SfmDMatch* d_matches, int64_t match_capacity, int64_t* n_matches) {
    int rc = require_descriptors(ctx);
    if (rc) return rc;
    if (n_pairs < 0 || (n_pairs > 0 && (!qt || !d_counts)) || match_capacity < 0 || !n_matches)
        return fail(ctx, SFMM_EINVAL, "match_pairs_device: bad argument");
    if ((rc = bind_device(ctx))) return rc;
    if ((rc = prepare_float(ctx))) return rc;
    *n_matches = 0;
    begin_stats(ctx);
    Slot& sl = ctx->slot[0];  // results chain through `written`: chunks run back to back on one stream
    CU_TRY(ctx, cudaEventRecord(ctx->ev_begin, sl.stream));
    int64_t done = 0;
    uint64_t written = 0;
    while (done < n_pairs) {
        const int64_t n = chunk_extent(ctx, qt, done, n_pairs, ctx->tensor_refine ? MAX_CHUNK_ROWS / 4 : MAX_CHUNK_ROWS);
        const uint64_t room = static_cast<uint64_t>(match_capacity) - std::min<uint64_t>(written, match_capacity);
        if ((rc = launch_chunk(ctx, sl, qt, done, n, d_counts + done, d_matches ? d_matches + written : nullptr, d_matches ? room : 0))) return rc;
        uint64_t total = 0;
        if ((rc = wait_chunk(ctx, sl, &total))) return rc;
        if (total > room) return fail(ctx, SFMM_ERANGE, "match_pairs_device: match_capacity too small");
        written += total;
        done += n;
    }
    CU_TRY(ctx, cudaEventRecord(ctx->ev_end, sl.stream));
    CU_TRY(ctx, cudaStreamSynchronize(sl.stream));
    float ms = 0.f;
    CU_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev_begin, ctx->ev_end));
    ctx->stats.last_match_ms = ms;
    ctx->stats.pairs_matched += n_pairs;
    *n_matches = static_cast<int64_t>(written);
    return SFMM_OK;
}

static int impl_match_pairs(SfmmCtx* ctx, const int32_t* qt, int64_t n_pairs) {
    int rc = require_descriptors(ctx);
    if (rc) return rc;
    if (n_pairs < 0 || (n_pairs > 0 && !qt)) return fail(ctx, SFMM_EINVAL, "match_pairs: bad argument");
    if ((rc = bind_device(ctx))) return rc;
    const bool trace = std::getenv("SFMM_TRACE_HOST") != nullptr;  // development aid: where the host time of a call goes
    const auto t_start = std::chrono::steady_clock::now();
    if ((rc = prepare_float(ctx))) return rc;
    const auto t_prepared = std::chrono::steady_clock::now();
    begin_stats(ctx);
    ctx->call_pairs = n_pairs;
    ctx->call_base = ctx->table.size();
    CU_TRY(ctx, cudaEventRecord(ctx->ev_begin, ctx->slot[0].stream));
    // Chunk size: about an eighth of the job (so that copies and bookkeeping overlap kernels; 4 chunks measured the
    // same end-to-end time at cfg-2, 2 chunks 10 % more),
    // never more than the scratch budget, never so small that a launch cannot fill the GPU.
    uint64_t q_rows = 0;
    for (int64_t i = 0; i < n_pairs; ++i) {
        const int32_t q = qt[2 * i];
        if (q >= 0 && q < ctx->n_images) q_rows += static_cast<uint64_t>(ctx->rows[q]);
    }
    const char* env_c = std::getenv("SFMM_CHUNKS");  // experiments: chunks per job
    const uint64_t env_chunks = env_c ? std::max(1, std::atoi(env_c)) : 0;
    const uint64_t min_rows = static_cast<uint64_t>(query_tile_rows(ctx)) * ctx->sm_count * 8;
    const uint64_t budget = std::min<uint64_t>(ctx->tensor_refine ? MAX_CHUNK_ROWS / 4 : MAX_CHUNK_ROWS, std::max<uint64_t>(min_rows, q_rows / (env_chunks ? env_chunks : 8) + 1));
    int64_t next = 0;
    int k = 0;
    int pending[2] = {-1, -1};  // slot indices in launch order
    // A failure in the middle of the call must not leave a chunk in flight or a half-indexed table behind:
    // drain both slots and roll the table back to where this call started.
    const size_t base_pairs = ctx->res_counts.size(), base_pts = ctx->pts_left.size();
    const int64_t base_matches = ctx->n_matches;
    auto abort_call = [&](int code) -> int {
        for (Slot& sl : ctx->slot) {
            if (sl.stream) cudaStreamSynchronize(sl.stream);
            sl.busy = false;
        }
        cudaStreamSynchronize(ctx->copy_stream);
        (void)cudaGetLastError();
        for (size_t i = base_pairs; i < ctx->res_counts.size(); ++i) ctx->index.erase(pair_key(ctx->res_qt[2 * i], ctx->res_qt[2 * i + 1]));
        ctx->res_qt.resize(2 * base_pairs);
        ctx->res_counts.resize(base_pairs);
        ctx->res_offsets.resize(base_pairs);
        ctx->res_slots.resize(base_pairs);
        ctx->table.resize_down(ctx->call_base);
        ctx->pts_left.resize(base_pts);
        ctx->pts_right.resize(base_pts);
        ctx->n_matches = base_matches;
        return code;
    };
    auto collect_oldest = [&]() -> int {
        const int s = pending[0];
        pending[0] = pending[1];
        pending[1] = -1;
        return collect_chunk(ctx, ctx->slot[s], qt);
    };
    try {
        while (next < n_pairs) {
            if (pending[1] >= 0 && (rc = collect_oldest())) return abort_call(rc);  // both slots busy: free the older one
            const int s = k & 1;
            const int64_t n = chunk_extent(ctx, qt, next, n_pairs, budget);
            if ((rc = launch_chunk(ctx, ctx->slot[s], qt, next, n, nullptr, nullptr, 0))) return abort_call(rc);
            (pending[0] < 0 ? pending[0] : pending[1]) = s;
            next += n;
            ++k;
        }
        while (pending[0] >= 0)
            if ((rc = collect_oldest())) return abort_call(rc);
    } catch (...) {
        abort_call(0);
        throw;  // mapped to an error code by the entry point's guard
    }
    // device time of the whole call: from the first launch to the later of the two streams
    const auto t_issued = std::chrono::steady_clock::now();
    CU_TRY(ctx, cudaEventRecord(ctx->ev_end, ctx->slot[0].stream));
    if ((rc = sync_all(ctx))) return rc;
    float ms = 0.f;
    CU_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev_begin, ctx->ev_end));
    if (trace) {
        const auto t_done = std::chrono::steady_clock::now();
        auto us = [](auto a, auto b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
        std::fprintf(stderr, "[sfmm] match_pairs(%lld): prepare %.0f us, launch+collect loop %.0f us, final sync %.0f us; device %.0f us\n",
                     static_cast<long long>(n_pairs), us(t_start, t_prepared), us(t_prepared, t_issued), us(t_issued, t_done), ms * 1e3);
    }
    ctx->stats.last_match_ms = ms;
    ctx->stats.pairs_matched += n_pairs;
    return SFMM_OK;
}

static int impl_match_all_pairs(SfmmCtx* ctx) {
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
    return impl_match_pairs(ctx, qt.data(), static_cast<int64_t>(qt.size() / 2));
}

SFMM_API int sfmm_get_pair(const SfmmCtx* ctx, int32_t q, int32_t t, const SfmDMatch** matches, int32_t* count) {
    if (!ctx) return SFMM_EINVAL;
    if (ctx->elem_type < 0) return SFMM_ESTATE;
    if (!matches || !count) return SFMM_EINVAL;
    if (q < 0 || q >= ctx->n_images || t < 0 || t >= ctx->n_images) return SFMM_ERANGE;
    auto it = ctx->index.find(pair_key(q, t));
    if (it == ctx->index.end()) return SFMM_ESTATE;
    const PairSlot& s = ctx->res_slots[static_cast<size_t>(it->second)];
    *matches = ctx->table.data() + s.offset;
    *count = s.count;
    return SFMM_OK;
}

SFMM_API int sfmm_result_table(const SfmmCtx* ctx, int64_t* n_pairs, const int32_t** qt, const int32_t** counts,
                               const int64_t** offsets, const SfmDMatch** matches, int64_t* n_matches) {
    if (!ctx) return SFMM_EINVAL;
    if (ctx->elem_type < 0) return SFMM_ESTATE;
    if (!n_pairs || !qt || !counts || !offsets || !matches || !n_matches) return SFMM_EINVAL;
    *n_pairs = static_cast<int64_t>(ctx->res_counts.size());
    *qt = ctx->res_qt.data();
    *counts = ctx->res_counts.data();
    *offsets = ctx->res_offsets.data();
    *n_matches = ctx->n_matches;
    *matches = ctx->table.data();
    return SFMM_OK;
}

static int impl_match_pair(SfmmCtx* ctx, int32_t q, int32_t t, SfmDMatch* out, int32_t cap, int32_t* count) {
    int rc = require_descriptors(ctx);
    if (rc) return rc;
    if (!count || cap < 0 || (cap > 0 && !out)) return fail(ctx, SFMM_EINVAL, "match_pair: bad argument");
    if (q < 0 || q >= ctx->n_images || t < 0 || t >= ctx->n_images) return fail(ctx, SFMM_ERANGE, "match_pair: image index out of range");
    if ((rc = bind_device(ctx))) return rc;
    if ((rc = prepare_float(ctx))) return rc;
    *count = 0;
    const int32_t qt[2] = {q, t};
    Slot& sl = ctx->slot[0];
    if ((rc = launch_chunk(ctx, sl, qt, 0, 1, nullptr, nullptr, 0))) return rc;
    uint64_t total = 0;
    if ((rc = wait_chunk(ctx, sl, &total))) return rc;
    ctx->stats.pairs_matched += 1;
    if (total > static_cast<uint64_t>(cap)) {
        *count = static_cast<int32_t>(total);
        return fail(ctx, SFMM_ERANGE, "match_pair: output capacity too small (count holds the size needed)");
    }
    if (total) {
        CU_TRY(ctx, cudaMemcpyAsync(out, sl.d_matches.p, static_cast<size_t>(total) * sizeof(SfmDMatch), cudaMemcpyDeviceToHost, sl.stream));
        CU_TRY(ctx, cudaStreamSynchronize(sl.stream));
        ctx->stats.d2h_bytes += static_cast<int64_t>(total * sizeof(SfmDMatch));
    }
    *count = static_cast<int32_t>(total);
    return SFMM_OK;
}

static int impl_knn_pair(SfmmCtx* ctx, int32_t q, int32_t t, int32_t* train_idx, float* distance) {
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
    const int32_t qt[2] = {q, t};
    Slot& sl = ctx->slot[0];
    if ((rc = launch_chunk(ctx, sl, qt, 0, 1, nullptr, nullptr, 0, /*knn_only=*/true, /*force_single_split=*/true))) return rc;
    CU_TRY(ctx, ctx->d_idx.ensure(static_cast<size_t>(nq) * 2 * sizeof(int32_t)));
    CU_TRY(ctx, ctx->d_dist.ensure(static_cast<size_t>(nq) * 2 * sizeof(float)));
    const int threads = 256, blocks = (nq + threads - 1) / threads;
    if (ctx->elem_type == SFMM_F32)
        knn_decode_kernel<true><<<blocks, threads, 0, sl.stream>>>(sl.d_pairs.as<PairDesc>(), sl.d_knn.as<KnnEntry>(), ctx->d_idx.as<int32_t>(), ctx->d_dist.as<float>());
    else
        knn_decode_kernel<false><<<blocks, threads, 0, sl.stream>>>(sl.d_pairs.as<PairDesc>(), sl.d_knn.as<KnnEntry>(), ctx->d_idx.as<int32_t>(), ctx->d_dist.as<float>());
    CU_TRY(ctx, cudaGetLastError());
    ctx->stats.kernel_launches += 1;
    CU_TRY(ctx, cudaMemcpyAsync(train_idx, ctx->d_idx.p, static_cast<size_t>(nq) * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, sl.stream));
    CU_TRY(ctx, cudaMemcpyAsync(distance, ctx->d_dist.p, static_cast<size_t>(nq) * 2 * sizeof(float), cudaMemcpyDeviceToHost, sl.stream));
    CU_TRY(ctx, cudaStreamSynchronize(sl.stream));
    sl.busy = false;
    ctx->stats.d2h_bytes += static_cast<int64_t>(nq) * 16;
    return SFMM_OK;
}

SFMM_API int sfmm_get_stats(const SfmmCtx* ctx, SfmmStats* out) {
    if (!ctx || !out) return SFMM_EINVAL;
    *out = ctx->stats;
    return SFMM_OK;
}

SFMM_API int sfmm_set_descriptors(SfmmCtx* ctx, int32_t n_images, const void* const* data, const int32_t* rows,
                                  int32_t cols, const size_t* step_bytes, int32_t elem_type) {
    return guarded(ctx, [&]() -> int { return impl_set_descriptors(ctx, n_images, data, rows, cols, step_bytes, elem_type); });
}

SFMM_API int sfmm_set_points(SfmmCtx* ctx, int32_t n_images, const double* const* xy) {
    return guarded(ctx, [&]() -> int { return impl_set_points(ctx, n_images, xy); });
}

SFMM_API int sfmm_match_pairs_device(SfmmCtx* ctx, const int32_t* qt, int64_t n_pairs, int32_t* d_counts,
                                     SfmDMatch* d_matches, int64_t match_capacity, int64_t* n_matches) {
    return guarded(ctx, [&]() -> int { return impl_match_pairs_device(ctx, qt, n_pairs, d_counts, d_matches, match_capacity, n_matches); });
}

SFMM_API int sfmm_match_pairs(SfmmCtx* ctx, const int32_t* qt, int64_t n_pairs) {
    return guarded(ctx, [&]() -> int { return impl_match_pairs(ctx, qt, n_pairs); });
}

SFMM_API int sfmm_match_pair(SfmmCtx* ctx, int32_t q, int32_t t, SfmDMatch* out, int32_t cap, int32_t* count) {
    return guarded(ctx, [&]() -> int { return impl_match_pair(ctx, q, t, out, cap, count); });
}

SFMM_API int sfmm_knn_pair(SfmmCtx* ctx, int32_t q, int32_t t, int32_t* train_idx, float* distance) {
    return guarded(ctx, [&]() -> int { return impl_knn_pair(ctx, q, t, train_idx, distance); });
}

SFMM_API int sfmm_match_all_pairs(SfmmCtx* ctx) {
    return guarded(ctx, [&]() -> int { return impl_match_all_pairs(ctx); });
}

}  // extern "C"

// =============================================================================== device groups (single process, NCCL)
namespace {

// The handful of NCCL entry points the group needs, resolved at run time so that libsfmmatch.so has no link-time
// dependency on NCCL (a process that already loaded an NCCL -- e.g. torch's -- gets that one: same SONAME).
struct NcclApi {
    void* handle = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    std::string error;
    bool load() {
        if (handle) return true;
        const char* env = std::getenv("SFMM_NCCL_LIB");
        const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n || !*n) continue;
            handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
            error = dlerror();
        }
        if (!handle) return false;
        bool ok = true;
        auto sym = [&](const char* name) -> void* {
            void* p = dlsym(handle, name);
            if (!p) { ok = false; error = std::string("missing NCCL symbol ") + name; }
            return p;
        };
        CommInitAll = reinterpret_cast<decltype(CommInitAll)>(sym("ncclCommInitAll"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(sym("ncclCommDestroy"));
        Broadcast = reinterpret_cast<decltype(Broadcast)>(sym("ncclBroadcast"));
        GroupStart = reinterpret_cast<decltype(GroupStart)>(sym("ncclGroupStart"));
        GroupEnd = reinterpret_cast<decltype(GroupEnd)>(sym("ncclGroupEnd"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(sym("ncclGetErrorString"));
        if (!ok) { dlclose(handle); handle = nullptr; }
        return ok;
    }
};
NcclApi g_nccl;

}  // namespace

struct SfmmGroup {
    std::vector<SfmmCtx*> ctx;
    std::vector<int> devices;
    std::vector<ncclComm_t> comms;
    std::vector<int8_t> owner;  // n_images x n_images: member that matched (q,t), -1 = nobody
    int32_t n_images = 0;
    int64_t h2d_bytes = 0, nccl_bytes = 0;
    mutable std::string err;
};

namespace {
thread_local std::string g_group_create_error;
int gfail(const SfmmGroup* g, int code, const std::string& msg) {
    if (g) g->err = msg;
    else g_group_create_error = msg;
    return code;
}
}  // namespace

extern "C" {

SFMM_API const char* sfmm_group_last_error(const SfmmGroup* g) { return g ? g->err.c_str() : g_group_create_error.c_str(); }
SFMM_API int32_t sfmm_group_size(const SfmmGroup* g) { return g ? static_cast<int32_t>(g->ctx.size()) : 0; }
SFMM_API SfmmCtx* sfmm_group_context(SfmmGroup* g, int32_t i) {
    return (g && i >= 0 && i < static_cast<int32_t>(g->ctx.size())) ? g->ctx[static_cast<size_t>(i)] : nullptr;
}

SFMM_API void sfmm_group_destroy(SfmmGroup* g) {
    if (!g) return;
    for (size_t i = 0; i < g->comms.size(); ++i)
        if (g->comms[i]) {
            cudaSetDevice(g->devices[i]);
            g_nccl.CommDestroy(g->comms[i]);
        }
    for (SfmmCtx* c : g->ctx) sfmm_destroy(c);
    delete g;
}

SFMM_API int sfmm_group_create(const SfmmConfig* cfg, int32_t n_devices, const int32_t* devices, SfmmGroup** out) {
    if (!cfg || !out || n_devices < 1 || n_devices > 64) return gfail(nullptr, SFMM_EINVAL, "group_create: bad argument");
    *out = nullptr;
    try {
        std::unique_ptr<SfmmGroup, void (*)(SfmmGroup*)> g(new SfmmGroup(), sfmm_group_destroy);
        for (int32_t i = 0; i < n_devices; ++i) {
            const int dev = devices ? devices[i] : i;
            for (int d : g->devices)
                if (d == dev) return gfail(nullptr, SFMM_EINVAL, "group_create: a device is listed twice");
            SfmmConfig c = *cfg;
            c.device = dev;
            SfmmCtx* ctx = nullptr;
            const int rc = sfmm_create(&c, &ctx);
            if (rc) return gfail(nullptr, rc, std::string("group_create: device ") + std::to_string(dev) + ": " + sfmm_last_error(nullptr));
            g->ctx.push_back(ctx);
            g->devices.push_back(dev);
        }
        if (n_devices > 1) {  // one communicator per device, single process
            if (!g_nccl.load()) return gfail(nullptr, SFMM_ENODEVICE, "group_create: NCCL could not be loaded (" + g_nccl.error + ")");
            g->comms.assign(static_cast<size_t>(n_devices), nullptr);
            const ncclResult_t r = g_nccl.CommInitAll(g->comms.data(), n_devices, g->devices.data());
            if (r != ncclSuccess) {
                g->comms.clear();
                return gfail(nullptr, SFMM_ECUDA, std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(r));
            }
        }
        *out = g.release();
        return SFMM_OK;
    } catch (const std::exception& e) {
        return gfail(nullptr, SFMM_ENOMEM, std::string("group_create: ") + e.what());
    }
}

SFMM_API int sfmm_group_set_descriptors(SfmmGroup* g, int32_t n_images, const void* const* data, const int32_t* rows, int32_t cols,
                                        const size_t* step_bytes, int32_t elem_type) {
    if (!g) return SFMM_EINVAL;
    g->owner.clear();
    g->n_images = 0;
    g->h2d_bytes = g->nccl_bytes = 0;
    // member 0 packs and uploads (the only host->device copy of the descriptors); the others reserve the same layout
    const int64_t h2d0 = g->ctx[0]->stats.h2d_bytes;
    int rc = sfmm_set_descriptors(g->ctx[0], n_images, data, rows, cols, step_bytes, elem_type);
    if (rc) return gfail(g, rc, std::string("member 0: ") + sfmm_last_error(g->ctx[0]));
    g->h2d_bytes = g->ctx[0]->stats.h2d_bytes - h2d0;
    for (size_t i = 1; i < g->ctx.size(); ++i)
        if ((rc = sfmm_set_descriptors(g->ctx[i], n_images, nullptr, rows, cols, nullptr, elem_type)))
            return gfail(g, rc, "member " + std::to_string(i) + ": " + sfmm_last_error(g->ctx[i]));
    const size_t bytes = g->ctx[0]->blob_bytes;
    if (g->ctx.size() > 1 && bytes) {
        ncclResult_t r = g_nccl.GroupStart();
        for (size_t i = 0; i < g->ctx.size() && r == ncclSuccess; ++i) {
            cudaSetDevice(g->devices[i]);
            r = g_nccl.Broadcast(g->ctx[i]->blob.p /* in place: only the root's is read */, g->ctx[i]->blob.p, bytes, ncclUint8, 0, g->comms[i],
                                 g->ctx[i]->slot[0].stream);
        }
        const ncclResult_t r2 = g_nccl.GroupEnd();
        if (r != ncclSuccess || r2 != ncclSuccess)
            return gfail(g, SFMM_ECUDA, std::string("ncclBroadcast: ") + g_nccl.GetErrorString(r != ncclSuccess ? r : r2));
        for (size_t i = 0; i < g->ctx.size(); ++i) {
            cudaSetDevice(g->devices[i]);
            const cudaError_t e = cudaStreamSynchronize(g->ctx[i]->slot[0].stream);
            if (e != cudaSuccess) return gfail(g, SFMM_ECUDA, std::string("broadcast sync: ") + cudaGetErrorString(e));
            g->ctx[i]->float_prepared = false;  // the blob was filled behind the context's back: re-derive norms / operand copies
        }
        g->nccl_bytes = static_cast<int64_t>(bytes) * static_cast<int64_t>(g->ctx.size() - 1);
    }
    g->n_images = n_images;
    return SFMM_OK;
}

SFMM_API int sfmm_group_match_pairs(SfmmGroup* g, const int32_t* qt, int64_t n_pairs) {
    if (!g || n_pairs < 0 || (n_pairs > 0 && !qt)) return SFMM_EINVAL;
    if (g->ctx[0]->elem_type < 0) return gfail(g, SFMM_ESTATE, "group_match_pairs: sfmm_group_set_descriptors has not been called");
    try {
        const size_t nd = g->ctx.size();
        const int64_t n = g->n_images;
        const std::vector<int32_t>& rows = g->ctx[0]->rows;
        for (int64_t i = 0; i < n_pairs; ++i)
            if (qt[2 * i] < 0 || qt[2 * i] >= n || qt[2 * i + 1] < 0 || qt[2 * i + 1] >= n)
                return gfail(g, SFMM_ERANGE, "group_match_pairs: image index out of range in pair list");
        // cost-sorted snake deal (the rule of sfm_danpipeline_b200/distributed.py: shard_pairs), then ascending pair order per member
        std::vector<int64_t> order(static_cast<size_t>(n_pairs));
        for (int64_t i = 0; i < n_pairs; ++i) order[static_cast<size_t>(i)] = i;
        std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
            return static_cast<int64_t>(rows[qt[2 * a]]) * rows[qt[2 * a + 1]] > static_cast<int64_t>(rows[qt[2 * b]]) * rows[qt[2 * b + 1]];
        });
        std::vector<std::vector<int64_t>> members(nd);
        for (size_t pos = 0; pos < order.size(); ++pos) {
            const size_t lap = pos / nd, off = pos % nd;
            members[(lap % 2 == 0) ? off : nd - 1 - off].push_back(order[pos]);
        }
        if (g->owner.size() != static_cast<size_t>(n * n)) g->owner.assign(static_cast<size_t>(n * n), -1);
        std::vector<std::vector<int32_t>> shard(nd);
        for (size_t d = 0; d < nd; ++d) {
            std::sort(members[d].begin(), members[d].end());
            shard[d].reserve(2 * members[d].size());
            for (int64_t k : members[d]) {
                shard[d].push_back(qt[2 * k]);
                shard[d].push_back(qt[2 * k + 1]);
                g->owner[static_cast<size_t>(qt[2 * k]) * n + qt[2 * k + 1]] = static_cast<int8_t>(d);
            }
        }
        std::vector<int> codes(nd, SFMM_OK);
        std::vector<std::thread> pool;
        for (size_t d = 1; d < nd; ++d)
            pool.emplace_back([&, d]() { codes[d] = sfmm_match_pairs(g->ctx[d], shard[d].data(), static_cast<int64_t>(shard[d].size() / 2)); });
        codes[0] = sfmm_match_pairs(g->ctx[0], shard[0].data(), static_cast<int64_t>(shard[0].size() / 2));
        for (auto& t : pool) t.join();
        for (size_t d = 0; d < nd; ++d)
            if (codes[d]) return gfail(g, codes[d], "member " + std::to_string(d) + ": " + sfmm_last_error(g->ctx[d]));
        return SFMM_OK;
    } catch (const std::exception& e) {
        return gfail(g, SFMM_ENOMEM, std::string("group_match_pairs: ") + e.what());
    }
}

SFMM_API int sfmm_group_match_all_pairs(SfmmGroup* g) {
    if (!g) return SFMM_EINVAL;
    try {
        for (SfmmCtx* c : g->ctx) clear_results(c);
        std::fill(g->owner.begin(), g->owner.end(), static_cast<int8_t>(-1));
        std::vector<int32_t> qt;  // findBestPair's enumeration, src/Sfm.cpp:511-512
        const int64_t n = g->n_images;
        qt.reserve(static_cast<size_t>(n > 1 ? n * (n - 1) : 0));
        for (int32_t q = 0; q + 1 < n; ++q)
            for (int32_t t = q + 1; t < n; ++t) {
                qt.push_back(q);
                qt.push_back(t);
            }
        return sfmm_group_match_pairs(g, qt.data(), static_cast<int64_t>(qt.size() / 2));
    } catch (const std::exception& e) {
        return gfail(g, SFMM_ENOMEM, std::string("group_match_all_pairs: ") + e.what());
    }
}

SFMM_API int sfmm_group_get_pair(const SfmmGroup* g, int32_t q, int32_t t, const SfmDMatch** matches, int32_t* count) {
    if (!g || !matches || !count) return SFMM_EINVAL;
    if (q < 0 || q >= g->n_images || t < 0 || t >= g->n_images) return SFMM_ERANGE;
    if (g->owner.empty()) return SFMM_ESTATE;
    const int8_t o = g->owner[static_cast<size_t>(q) * g->n_images + t];
    if (o < 0) return SFMM_ESTATE;
    return sfmm_get_pair(g->ctx[static_cast<size_t>(o)], q, t, matches, count);
}

SFMM_API int sfmm_group_transfer_stats(const SfmmGroup* g, int64_t* h2d_bytes, int64_t* nccl_bytes) {
    if (!g || !h2d_bytes || !nccl_bytes) return SFMM_EINVAL;
    *h2d_bytes = g->h2d_bytes;
    *nccl_bytes = g->nccl_bytes;
    return SFMM_OK;
}

}  // extern "C"
