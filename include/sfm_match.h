/*
 * sfm_match.h -- C ABI of libsfmmatch.so, the B200 (sm_100a) replacement for the descriptor
 * matching hot path of codebydant/sfM_danPipeline (iTree3DMap).
 *
 * Every entry point cites the reference interface it replaces; file:line are relative to
 * the reference tree.  The reference path is
 *
 *     StructFromMotion::getMatching(const int& idx_query, const int& idx_train,
 *                                   Matching* goodMatches)          src/Sfm.cpp:590-608
 *         cv::BFMatcher(cv::NORM_L2,false).knnMatch(q, t, knn, 2)   src/Sfm.cpp:593,599
 *         ratio loop  knn[i][0].distance <= NN_MATCH_RATIO * knn[i][1].distance   :603-607
 *     driven over every pair q<t by findBestPair                    src/Sfm.cpp:511-515
 *     and re-invoked by baseReconstruction / addMoreViews / find2D3DMatches
 *                                                                   src/Sfm.cpp:426,977,1031
 *
 * Conventions
 *   - plain C types only: pointers, sizes, POD structs.  Nothing throws across the ABI (every entry
 *     point that allocates maps C++ exceptions to SFMM_ENOMEM / SFMM_EINVAL).
 *   - every call returns SFMM_OK (0) or a negative SFMM_E* code; sfmm_last_error() gives text.
 *   - a context is bound to ONE CUDA device and is not thread-safe (the reference calls
 *     getMatching from a single thread); the look-ups sfmm_get_pair / sfmm_get_pair_points /
 *     sfmm_result_table / sfmm_image_rows are read-only, do not touch the error text (their codes are
 *     self-explanatory: SFMM_ESTATE = not computed) and may be called concurrently once
 *     sfmm_match_all_pairs / sfmm_match_pairs has returned.  A failing sfmm_match_pairs leaves no chunk
 *     in flight and rolls the table back to where the call started.
 *   - there is NO CPU fallback: without a usable CUDA device sfmm_create fails with
 *     SFMM_ENODEVICE.
 *   - match lists are in ascending queryIdx, at most one entry per query row, imgIdx == 0,
 *     distance == the 1-NN distance as float -- the order and fields the reference's
 *     push_back loop produces and its consumers rely on (AlignedPointsFromMatch
 *     src/Sfm.cpp:700-711, RANSAC mask indexing :549-553).
 *   - defined edge cases (the reference has undefined behaviour for the last one,
 *     knnMatches[i][1] at src/Sfm.cpp:604): a query set with 0 rows, or a train set with
 *     fewer than 2 rows, yields 0 matches.
 */
#ifndef SFM_MATCH_H_
#define SFM_MATCH_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SFMM_API __attribute__((visibility("default")))
#else
#define SFMM_API
#endif

/* Layout-identical to cv::DMatch {int queryIdx; int trainIdx; int imgIdx; float distance;}
 * (the element type of `Matching`, include/Utilities.h:27), so a match list can be copied
 * into a std::vector<cv::DMatch> with memcpy. */
typedef struct SfmDMatch {
    int32_t queryIdx;
    int32_t trainIdx;
    int32_t imgIdx;
    float distance;
} SfmDMatch;

enum {
    SFMM_OK = 0,
    SFMM_EINVAL = -1,    /* bad argument: NULL, negative size, type/width mismatch, unknown enum */
    SFMM_ENOMEM = -2,    /* host or device allocation failed */
    SFMM_ECUDA = -3,     /* a CUDA runtime call or kernel failed (text in sfmm_last_error) */
    SFMM_ESTATE = -4,    /* call out of order (no descriptors yet, pair not computed, ...) */
    SFMM_ERANGE = -5,    /* image index / capacity out of range; rows >= 2^18 (OpenCV's own limit) */
    SFMM_ENODEVICE = -6  /* no CUDA device / wrong architecture: there is no CPU fallback */
};

/* cv::NORM_HAMMING / cv::NORM_L2 of the BFMatcher constructor (src/Sfm.cpp:593).  The reference
 * hard-wires NORM_L2 for all three detectors; HAMMING is the right norm for its AKAZE/ORB detectors
 * (src/Sfm.cpp:331-384).  Both pairings the reference can produce are supported: NORM_L2 over CV_32F rows
 * (SIFT) and NORM_L2 over CV_8U rows (AKAZE / ORB through the literal src/Sfm.cpp:593 call: OpenCV's
 * batchDistL2_8u32f, sqrtf of the exact integer sum -- the bytes are widened to fp32 on upload and the float
 * kernels take over, bit-exact).  NORM_HAMMING needs CV_8U rows, as cv::BFMatcher asserts. */
enum { SFMM_NORM_HAMMING = 0, SFMM_NORM_L2 = 1 };
/* cv::Mat depth of imagesDescriptors (include/Sfm.h:29): CV_8U (AKAZE, ORB) or CV_32F (SIFT). */
enum { SFMM_U8 = 0, SFMM_F32 = 1 };
/* How the L2 path ranks candidates (distances REPORTED are always fp32 direct-difference). */
enum {
    SFMM_FLOAT_AUTO = 0,   /* tensor cores whenever the shape allows (exact keys for integer-valued data such as SIFT,
                              contracted from an fp16 copy; TF32 ranking + exact fp32 refinement of the candidates for
                              arbitrary values), else the exact kernel */
    SFMM_FLOAT_EXACT = 1,  /* fp32 CUDA-core direct difference for every candidate */
    SFMM_FLOAT_TENSOR = 2  /* tcgen05 TF32 |a|^2+|b|^2-2ab ranking + fp32 refinement of the winners */
};

/* Which pipe evaluates Hamming distances.  Results are bit-identical. */
enum {
    SFMM_BINARY_AUTO = 0,   /* default: the tensor engine when the descriptors are <= 512 bit and their unpacked copy
                               (4x the packed bytes, 8x at exactly 512 bit) fits in half of the free device memory, else POPC */
    SFMM_BINARY_POPC = 1,   /* the north-star design: XOR + carry-save + POPC on the integer pipes, packed descriptors */
    SFMM_BINARY_TENSOR = 2  /* popc(a)+popc(b)-2a.b with a.b on the tensor cores: below 512 bit every bit an E2M1 nibble on the
                               FP4 pipe (tcgen05 kind::mxf4, unit block scales), at 512 bit a {0,1} byte on kind::i8 (~8-10x the
                               POPC rate); SFMM_EINVAL for descriptors > 512 bit */
};

typedef struct SfmmConfig {
    int32_t struct_size;  /* = sizeof(SfmmConfig); set by sfmm_default_config */
    int32_t device;       /* CUDA device ordinal this context lives on */
    int32_t norm;         /* SFMM_NORM_* -- first argument of cv::BFMatcher(...), src/Sfm.cpp:593 */
    float ratio;          /* NN_MATCH_RATIO, include/Sfm.h:27,60 (0.8f) */
    int32_t cross_check;  /* 0 = reference behaviour (crossCheck=false, src/Sfm.cpp:593);
                             1 = additionally require q == lowest-index argmin_q' d(q',t) */
    int32_t float_mode;   /* SFMM_FLOAT_* */
    int32_t pair_batch;   /* max image pairs per kernel launch; 0 = automatic */
    int32_t binary_engine; /* SFMM_BINARY_* (Hamming only) */
} SfmmConfig;

typedef struct SfmmStats {
    int64_t kernel_launches;   /* kernels this context has launched so far */
    int64_t pairs_matched;     /* image pairs matched so far */
    int64_t h2d_bytes;         /* host->device bytes copied so far */
    int64_t d2h_bytes;         /* device->host bytes copied so far */
    double last_match_ms;      /* device time (CUDA events) of the last sfmm_match_* call */
    double last_knn_ms;        /* of which: the 2-NN distance kernel(s) */
    double last_knn_work;      /* algorithmic work of those launches: POPC32 ops (Hamming) or FLOPs (L2) */
    int64_t last_knn_launches; /* number of 2-NN kernel launches behind last_knn_ms */
    int64_t float_path;        /* L2: 0 = not decided yet, SFMM_FLOAT_EXACT or SFMM_FLOAT_TENSOR = the kernel in use,
                                  3 = tensor-core fp16/TF32 ranking + exact refinement (arbitrary floats);
                                  Hamming: SFMM_FLOAT_TENSOR when the tensor engine is in use, else 0 */
    int64_t tensor_kind;       /* tcgen05.mma kind of the 2-NN kernel in use: 0 = none (CUDA-core kernels), 1 = kind::i8 (bits as bytes),
                                  2 = kind::mxf4 (bits as E2M1 nibbles, unit block scales), 3 = kind::f16, 4 = kind::tf32 */
} SfmmStats;

typedef struct SfmmCtx SfmmCtx;

/* Fills *cfg with the reference's constants: NORM_L2 (src/Sfm.cpp:593), ratio 0.8f
 * (include/Sfm.h:60), cross_check 0, device 0. */
SFMM_API void sfmm_default_config(SfmmConfig* cfg);

/* Replaces `new cv::BFMatcher(cv::NORM_L2,false)` (src/Sfm.cpp:593; the reference leaks one per
 * call).  One context per StructFromMotion object. */
SFMM_API int sfmm_create(const SfmmConfig* cfg, SfmmCtx** out);
SFMM_API void sfmm_destroy(SfmmCtx* ctx);
/* Text of the last error on this context (or of the last failed sfmm_create when ctx==NULL). */
SFMM_API const char* sfmm_last_error(const SfmmCtx* ctx);

/*
 * Hands the context the std::vector<cv::Mat> imagesDescriptors (include/Sfm.h:29, filled at
 * src/Sfm.cpp:326,353,381): image i is rows[i] x cols elements of elem_type starting at data[i]
 * with a row step of step_bytes[i] (cv::Mat::step; NULL = tightly packed).  The data is COPIED
 * (re-pitched to 16-byte-aligned rows with zeroed padding) into device memory; the caller keeps
 * ownership of its cv::Mats.  data == NULL only reserves the device layout (used by non-root ranks
 * before a broadcast into sfmm_descriptor_blob).  Invalidates all results.
 */
SFMM_API int sfmm_set_descriptors(SfmmCtx* ctx, int32_t n_images, const void* const* data,
                                  const int32_t* rows, int32_t cols, const size_t* step_bytes,
                                  int32_t elem_type);

/* rows[image] of the current descriptor set (cv::Mat::rows of imagesDescriptors[image]). */
SFMM_API int sfmm_image_rows(const SfmmCtx* ctx, int32_t image, int32_t* rows);

/* Device address and size of the packed descriptor blob (all images back to back, row pitch
 * sfmm_row_pitch(cols, elem_type)).  A multi-GPU host broadcasts rank 0's blob into the other
 * ranks' blobs (NCCL) instead of re-uploading from the host. */
SFMM_API int sfmm_descriptor_blob(SfmmCtx* ctx, void** device_ptr, size_t* bytes);
SFMM_API size_t sfmm_row_pitch(int32_t cols, int32_t elem_type);

/* The all-pairs loop of findBestPair (src/Sfm.cpp:511-515): computes getMatching for every
 * q<t once; afterwards getMatching(q,t) is a table look-up (sfmm_get_pair). */
SFMM_API int sfmm_match_all_pairs(SfmmCtx* ctx);

/* Same for an explicit list of ordered pairs (qt[2*i], qt[2*i+1]) -- the unit a multi-GPU host
 * shards across ranks, and what addMoreViews / find2D3DMatches ask for one view at a time
 * (src/Sfm.cpp:964-977,1020-1042).  Adds to the table built so far. */
SFMM_API int sfmm_match_pairs(SfmmCtx* ctx, const int32_t* qt, int64_t n_pairs);

/* Body of the patched getMatching (src/Sfm.cpp:590-608): borrow the match list of pair (q,t).
 * *matches stays valid until the next sfmm_match_pairs / sfmm_match_all_pairs (the table may grow),
 * sfmm_set_descriptors, sfmm_clear_results or sfmm_destroy.
 * SFMM_ESTATE if the pair has not been computed. */
SFMM_API int sfmm_get_pair(const SfmmCtx* ctx, int32_t q, int32_t t, const SfmDMatch** matches,
                           int32_t* count);

/* getMatching computed on demand for one pair into caller memory (capacity `cap` records,
 * rows[q] always suffices).  Does not touch the table. */
SFMM_API int sfmm_match_pair(SfmmCtx* ctx, int32_t q, int32_t t, SfmDMatch* out, int32_t cap,
                             int32_t* count);

/* The raw result of matcher->knnMatch(q, t, knnMatches, 2) (src/Sfm.cpp:599) as arrays:
 * train_idx[2*i+k], distance[2*i+k] for query row i, neighbour k; missing neighbours are
 * (-1, FLT_MAX) like cv::batchDistance's initialisation.  For parity tests. */
SFMM_API int sfmm_knn_pair(SfmmCtx* ctx, int32_t q, int32_t t, int32_t* train_idx, float* distance);

/* Flat view of everything matched since the last clear, in the order the pairs were given:
 * pair i = (qt[2i], qt[2i+1]) owns matches[offsets[i] .. offsets[i]+counts[i]).  Host pointers,
 * same lifetime as sfmm_get_pair's. */
SFMM_API int sfmm_result_table(const SfmmCtx* ctx, int64_t* n_pairs, const int32_t** qt,
                               const int32_t** counts, const int64_t** offsets,
                               const SfmDMatch** matches, int64_t* n_matches);

/* Device-resident variant for hosts that move results between GPUs themselves (NCCL gather to
 * rank 0): matches the given pairs and leaves counts (int32[n_pairs]) and the packed SfmDMatch
 * records in caller-provided DEVICE buffers; *n_matches is the total written.  SFMM_ERANGE when
 * match_capacity (records) is too small -- sum of rows[q] over the pairs always suffices. */
SFMM_API int sfmm_match_pairs_device(SfmmCtx* ctx, const int32_t* qt, int64_t n_pairs,
                                     int32_t* d_counts, SfmDMatch* d_matches,
                                     int64_t match_capacity, int64_t* n_matches);

/* ---- next rows of the path --------------------------------------------------- */

/* Optional: the std::vector<std::vector<cv::Point2d>> imagesPts2D (include/Sfm.h:30, filled by
 * keypointstoPoints at src/Sfm.cpp:323,350,378): image i has rows[i] points, xy[i] points at
 * rows[i] x {double x, double y}.  Call after sfmm_set_descriptors and before matching; from then
 * on sfmm_match_all_pairs / sfmm_match_pairs also gather, on the GPU, the aligned point lists that
 * AlignedPointsFromMatch (src/Sfm.cpp:694-711) builds on the CPU for every pair. */
SFMM_API int sfmm_set_points(SfmmCtx* ctx, int32_t n_images, const double* const* xy);

/* alignedL / alignedR of pair (q,t): count x {x,y} doubles each, in match order (left[i] belongs to
 * matches[i].queryIdx, right[i] to matches[i].trainIdx).  Same lifetime as sfmm_get_pair. */
SFMM_API int sfmm_get_pair_points(const SfmmCtx* ctx, int32_t q, int32_t t, const double** left_xy,
                                  const double** right_xy, int32_t* count);

/* Persisted all-pairs match table (the reference has no checkpointing): a flat
 * little-endian file -- 64-byte header, rows[n_images], qt[2*n_pairs], counts[n_pairs],
 * offsets[n_pairs], SfmDMatch[n_matches] -- so that later runs / downstream stages skip matching.
 * sfmm_load_table needs the same image count, row counts, norm, ratio and cross_check setting as the
 * current context (a table is only valid for the filter it was computed with; SFMM_EINVAL otherwise) and a
 * file whose size agrees with its header; it replaces the table, afterwards sfmm_get_pair serves the stored
 * lists.  Aligned points are not part of the file: sfmm_get_pair_points answers SFMM_ESTATE after a load. */
SFMM_API int sfmm_save_table(const SfmmCtx* ctx, const char* path);
SFMM_API int sfmm_load_table(SfmmCtx* ctx, const char* path);

/* ---- all GPUs of one box from ONE host process ----------------------------
 * The reference is a single C++ program (main.cpp:18): a group gives it every B200 of the box without
 * becoming multi-process.  One context per device plus one NCCL communicator per device created with
 * ncclCommInitAll (NCCL is dlopen'ed at the first sfmm_group_create: "libnccl.so.2"; SFMM_ENODEVICE
 * if it cannot be loaded -- single-device contexts never need it).
 *   sfmm_group_set_descriptors : imagesDescriptors re-pitched and copied host->device ONCE (device 0's
 *       PCIe link), then ncclBroadcast of the packed blob to the other devices over NVLink;
 *   sfmm_group_match_all_pairs / _match_pairs : findBestPair's q<t pairs (src/Sfm.cpp:511-515) dealt to
 *       the devices by descending cost rows_q*rows_t in snake order (deterministic, balanced), one host
 *       thread per device drives its shard through the pipelined chunk scheduler; every device copies
 *       its records to host memory over its own PCIe link as chunks finish (no rank-0 funnel);
 *   sfmm_group_get_pair : getMatching(q,t) (src/Sfm.cpp:590-608) as a look-up in the owning device's
 *       host table.  Same ownership / lifetime rules as sfmm_get_pair. */
typedef struct SfmmGroup SfmmGroup;
SFMM_API int sfmm_group_create(const SfmmConfig* cfg /* .device ignored */, int32_t n_devices,
                               const int32_t* devices /* NULL = 0..n_devices-1 */, SfmmGroup** out);
SFMM_API void sfmm_group_destroy(SfmmGroup* g);
SFMM_API const char* sfmm_group_last_error(const SfmmGroup* g);
SFMM_API int32_t sfmm_group_size(const SfmmGroup* g);
/* The context of member i (statistics, on-demand sfmm_match_pair, ...); owned by the group. */
SFMM_API SfmmCtx* sfmm_group_context(SfmmGroup* g, int32_t i);
SFMM_API int sfmm_group_set_descriptors(SfmmGroup* g, int32_t n_images, const void* const* data,
                                        const int32_t* rows, int32_t cols, const size_t* step_bytes,
                                        int32_t elem_type);
SFMM_API int sfmm_group_match_all_pairs(SfmmGroup* g);
SFMM_API int sfmm_group_match_pairs(SfmmGroup* g, const int32_t* qt, int64_t n_pairs);
SFMM_API int sfmm_group_get_pair(const SfmmGroup* g, int32_t q, int32_t t, const SfmDMatch** matches,
                                 int32_t* count);
/* Bytes moved by the last sfmm_group_set_descriptors: host->device (once) and device->device (NCCL). */
SFMM_API int sfmm_group_transfer_stats(const SfmmGroup* g, int64_t* h2d_bytes, int64_t* nccl_bytes);

/* Multi-PROCESS hosts (one rank per GPU): put this context's match table into a POSIX shared-memory segment
 * ("<shm_prefix>.<generation>", page-locked with cudaHostRegister) instead of private pinned memory, so that the
 * rank that assembles the all-pairs result maps the records this GPU copied over its OWN PCIe link -- no inter-GPU
 * transfer and no second host copy ("gathered to rank 0" without a rank-0 funnel; the NCCL gather of
 * sfm_danpipeline_b200/distributed.py remains the alternative).  shm_prefix = "/name" (no further slashes); NULL or ""
 * returns to private memory.  Drops all results.  sfmm_shared_table_info reports the segment currently holding the
 * table (it changes when the table grows) and the number of records in it; the segment is unlinked by
 * sfmm_destroy / the next growth, a reader's existing mapping stays valid until it unmaps. */
SFMM_API int sfmm_share_table(SfmmCtx* ctx, const char* shm_prefix);
SFMM_API int sfmm_shared_table_info(const SfmmCtx* ctx, char* name, size_t name_capacity, int64_t* n_records);

SFMM_API int sfmm_clear_results(SfmmCtx* ctx);
SFMM_API int sfmm_get_stats(const SfmmCtx* ctx, SfmmStats* out);
/* "major.minor.patch (sm_100a)" */
SFMM_API const char* sfmm_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SFM_MATCH_H_ */
