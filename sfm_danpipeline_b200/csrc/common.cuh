// Shared device-side types of libsfmmatch (sm_100a only).
//
// Vocabulary follows the reference's domain: an *image* owns a descriptor set (a cv::Mat of
// imagesDescriptors, /root/reference/include/Sfm.h:29); a *pair* is one ordered (query image,
// train image) call of StructFromMotion::getMatching (/root/reference/src/Sfm.cpp:590-608);
// a *tile* is the slice of a pair one thread block works on.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sfmm {

// One getMatching call, as the kernels see it.  Rows are addressed in the packed blob:
// image rows are contiguous, `pitch` bytes apart (sfmm_row_pitch).
struct PairDesc {
    uint32_t q_row0;      // first blob row of the query image
    uint32_t nq;          // rows of the query image
    uint32_t t_row0;      // first blob row of the train image
    uint32_t nt;          // rows of the train image
    uint64_t knn_off;     // first entry of this pair in the 2-NN scratch (n_splits * nq entries)
    uint64_t col_off;     // first entry of this pair in the column-minimum scratch (nt entries)
    uint32_t n_splits;    // how many train-range splits produced partial 2-NN lists
    uint32_t first_ftile; // index of this pair's first filter tile
    uint32_t n_ftiles;    // number of filter tiles (ceil(nq / FILTER_TILE))
    uint32_t q_off;       // first query row of this pair in per-launch row-indexed scratch (candidate lists)
    float t_maxnorm2;     // max |t|^2 over the train image (error bound of the TF32 ranking pass)
    uint32_t pad;
};

// One thread block of the 2-NN kernels: a tile of query rows against a train-row range.
struct KnnTile {
    uint32_t pair;   // index into the PairDesc array of the launch
    uint32_t q0;     // first query row (image-relative) of the tile
    uint32_t t0;     // train range [t0, t1), image-relative
    uint32_t t1;
    uint32_t split;  // which partial list this tile writes; bit 31 = "reverse" tile (roles of the two images swapped: its rows'
                     // 1-NN are the forward problem's column minima), bit 30 = the reverse tile's rows are GATHERED through the pair's
                     // candidate list (q0 counts list entries), see filter.cuh: cross_mark / cross_compact
};
static constexpr uint32_t TILE_REVERSE = 0x80000000u, TILE_GATHER = 0x40000000u, TILE_SPLIT_MASK = 0x3FFFFFFFu;

// One thread block of the ratio/cross-check/compaction kernels.
struct FilterTile {
    uint32_t pair;
    uint32_t q0;
};

// 2-NN scratch entry: two 64-bit keys (distance << 32 | train index), smallest first.
// Hamming: distance is the integer bit count; L2: the IEEE bits of a non-negative float
// (monotone as unsigned).  A missing neighbour is KEY_NONE, which sorts last.
// Lexicographic (distance, index) order on the key == cv::batchDistance's strict-'<'
// insertion order: lowest train index wins ties, in both slots.
typedef ulonglong2 KnnEntry;
static constexpr unsigned long long KEY_NONE = 0xFFFFFFFFFFFFFFFFull;

static constexpr int FILTER_TILE = 1024;  // query rows per filter tile (256 threads x 4)
static constexpr int IDX_BITS = 18;       // OpenCV packs (imgIdx, trainIdx) with an 18-bit shift:
                                          // train sets are limited to < 2^18 rows there too.

__device__ __forceinline__ unsigned long long make_key(uint32_t dist_bits, uint32_t idx) {
    return (static_cast<unsigned long long>(dist_bits) << 32) | idx;
}

}  // namespace sfmm
