// sfm_match_opencv.hpp -- header-only C++11 adapter between the reference's OpenCV types and the
// C ABI of libsfmmatch.so (sfm_match.h).  Compile it inside iTree3DMap (it only needs
// <opencv2/core.hpp>); nothing here touches CUDA.
//
// Reference types kept at the boundary (file:line in the reference tree):
//     std::vector<cv::Mat>  imagesDescriptors                 include/Sfm.h:29
//     using Matching = std::vector<cv::DMatch>                include/Utilities.h:27
//     void getMatching(const int&, const int&, Matching*)     include/Sfm.h:89, src/Sfm.cpp:590-608
#ifndef SFM_MATCH_OPENCV_HPP_
#define SFM_MATCH_OPENCV_HPP_

#include <algorithm>
#include <cstring>
#include <memory>
#include <numeric>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include <opencv2/core.hpp>

#include "sfm_features.h"
#include "sfm_match.h"

namespace sfmm {

static_assert(sizeof(cv::DMatch) == sizeof(SfmDMatch), "cv::DMatch and SfmDMatch must have the same layout");

// The reference has no error channel on this path (void functions; cv::Exception propagates):
// ABI error codes become exceptions of the same spirit.
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

// std::vector<cv::Mat> imagesDescriptors (include/Sfm.h:29) -> the pointer / rows / step arrays of sfmm_set_descriptors.
// One BFMatcher, one descriptor type: every non-empty Mat must have the same width and depth.
inline void describe(const std::vector<cv::Mat>& imagesDescriptors, std::vector<const void*>& data, std::vector<int32_t>& rows,
                     std::vector<size_t>& steps, int& cols, int& type) {
    const int n = static_cast<int>(imagesDescriptors.size());
    data.assign(n, nullptr);
    rows.assign(n, 0);
    steps.assign(n, 0);
    cols = 0;
    type = SFMM_F32;
    for (int i = 0; i < n; ++i) {
        const cv::Mat& m = imagesDescriptors[i];
        if (m.empty()) {  // an image without keypoints: cv::Mat() -- zero rows
            data[i] = nullptr; rows[i] = 0; steps[i] = 0;
            continue;
        }
        if (m.dims != 2 || m.channels() != 1 || (m.depth() != CV_8U && m.depth() != CV_32F))
            throw Error(SFMM_EINVAL, "descriptors must be single-channel 2-D CV_8U or CV_32F");
        const int t = (m.depth() == CV_8U) ? SFMM_U8 : SFMM_F32;
        if (cols == 0) {
            cols = m.cols;
            type = t;
        } else if (m.cols != cols || t != type) {  // one BFMatcher, one descriptor type: every image must agree
            throw Error(SFMM_EINVAL, "all descriptor sets must have the same width and depth");
        }
        if (static_cast<size_t>(m.step) < static_cast<size_t>(m.cols) * m.elemSize())
            throw Error(SFMM_EINVAL, "descriptor row step smaller than a row");
        data[i] = m.data; rows[i] = m.rows; steps[i] = m.step;
    }
    if (cols == 0) cols = 1;  // no image has descriptors: an empty table
}

class AllPairsMatcher {
  public:
    // normType: cv::NORM_L2 (what src/Sfm.cpp:593 passes) or cv::NORM_HAMMING / NORM_HAMMING2-free
    // binary matching for the AKAZE / ORB detectors (src/Sfm.cpp:331-384).
    explicit AllPairsMatcher(int normType = cv::NORM_L2, float ratio = 0.8f, bool crossCheck = false, int device = 0)
        : ctx_(nullptr) {
        SfmmConfig cfg;
        sfmm_default_config(&cfg);
        cfg.device = device;
        cfg.norm = (normType == cv::NORM_HAMMING) ? SFMM_NORM_HAMMING : SFMM_NORM_L2;
        cfg.ratio = ratio;
        cfg.cross_check = crossCheck ? 1 : 0;
        const int rc = sfmm_create(&cfg, &ctx_);
        if (rc != SFMM_OK) throw Error(rc, sfmm_last_error(nullptr));
    }
    ~AllPairsMatcher() { sfmm_destroy(ctx_); }
    AllPairsMatcher(const AllPairsMatcher&) = delete;
    AllPairsMatcher& operator=(const AllPairsMatcher&) = delete;

    // Call once after extractFeature() (src/Sfm.cpp:21): uploads every image's descriptors and
    // runs findBestPair's whole q<t loop (src/Sfm.cpp:511-515) on the GPU.
    void compute(const std::vector<cv::Mat>& imagesDescriptors) {
        upload(imagesDescriptors);
        check(sfmm_match_all_pairs(ctx_));
    }

    // Building blocks (also used by MultiGpuMatcher): descriptors to this device; match an explicit
    // list of ordered pairs (q0,t0,q1,t1,...) -- the unit that is sharded across devices.
    void matchPairs(const std::vector<int32_t>& qt) { check(sfmm_match_pairs(ctx_, qt.data(), static_cast<int64_t>(qt.size() / 2))); }
    void upload(const std::vector<cv::Mat>& imagesDescriptors) {
        std::vector<const void*> data;
        std::vector<int32_t> rows;
        std::vector<size_t> steps;
        int cols = 0, type = SFMM_F32;
        describe(imagesDescriptors, data, rows, steps, cols, type);
        check(sfmm_set_descriptors(ctx_, static_cast<int32_t>(rows.size()), data.data(), rows.data(), cols, steps.data(), type));
    }

  public:

    // Drop-in body of StructFromMotion::getMatching: APPENDS to *goodMatches like the
    // reference's push_back loop (src/Sfm.cpp:603-607); no clear().
    void getMatching(const int& idx_query, const int& idx_train, std::vector<cv::DMatch>* goodMatches) {
        const SfmDMatch* m = nullptr;
        int32_t n = 0;
        int rc = sfmm_get_pair(ctx_, idx_query, idx_train, &m, &n);
        if (rc == SFMM_ESTATE) {  // a pair outside the q<t table (never requested by the reference): on demand, one call
            int32_t nq = 0;
            if (sfmm_image_rows(ctx_, idx_query, &nq) != SFMM_OK) throw Error(SFMM_ERANGE, "getMatching: image index out of range");
            std::vector<SfmDMatch> tmp(static_cast<size_t>(nq > 0 ? nq : 1));  // at most one match per query row
            check(sfmm_match_pair(ctx_, idx_query, idx_train, tmp.data(), static_cast<int32_t>(tmp.size()), &n));
            append(tmp.data(), n, goodMatches);
            return;
        }
        if (rc != SFMM_OK) throw Error(rc, rc == SFMM_ERANGE ? "getMatching: image index out of range" : "getMatching: no descriptors uploaded");
        append(m, n, goodMatches);
    }

    // ---- next rows --------------------------------------------------------------
    // imagesPts2D (include/Sfm.h:30): after this, compute() also gathers on the GPU what
    // AlignedPointsFromMatch (src/Sfm.cpp:694-711) builds per pair on the CPU.  Call between the
    // descriptor upload and the matching: compute(desc, &imagesPts2D) does both in order.
    void compute(const std::vector<cv::Mat>& imagesDescriptors, const std::vector<std::vector<cv::Point2d> >& imagesPts2D) {
        upload(imagesDescriptors);
        static_assert(sizeof(cv::Point2d) == 2 * sizeof(double), "cv::Point2d is two doubles");
        std::vector<const double*> xy(imagesPts2D.size());
        for (size_t i = 0; i < imagesPts2D.size(); ++i)
            xy[i] = imagesPts2D[i].empty() ? nullptr : reinterpret_cast<const double*>(imagesPts2D[i].data());
        check(sfmm_set_points(ctx_, static_cast<int32_t>(xy.size()), xy.data()));
        check(sfmm_match_all_pairs(ctx_));
    }

    // Drop-in for AlignedPointsFromMatch(imagesPts2D[q], imagesPts2D[t], matches, alignedL, alignedR):
    // appends, like the reference's push_back loop.
    void getAlignedPoints(int idx_query, int idx_train, std::vector<cv::Point2d>& alignedL, std::vector<cv::Point2d>& alignedR) {
        const double *l = nullptr, *r = nullptr;
        int32_t n = 0;
        const int rc = sfmm_get_pair_points(ctx_, idx_query, idx_train, &l, &r, &n);  // (look-ups do not set the error text)
        if (rc != SFMM_OK) throw Error(rc, "getAlignedPoints: pair not matched with points (compute(desc, pts) first; not restored by loadTable)");
        const size_t ol = alignedL.size(), orr = alignedR.size();
        alignedL.resize(ol + n);
        alignedR.resize(orr + n);
        if (n > 0) {
            std::memcpy(static_cast<void*>(alignedL.data() + ol), l, static_cast<size_t>(n) * 2 * sizeof(double));
            std::memcpy(static_cast<void*>(alignedR.data() + orr), r, static_cast<size_t>(n) * 2 * sizeof(double));
        }
    }

    // Persisted match table: later runs skip matching (the reference has no checkpointing).
    void saveTable(const std::string& path) { check(sfmm_save_table(ctx_, path.c_str())); }
    void loadTable(const std::vector<cv::Mat>& imagesDescriptors, const std::string& path) {
        upload(imagesDescriptors);
        check(sfmm_load_table(ctx_, path.c_str()));
    }

    SfmmCtx* handle() { return ctx_; }

  private:
    static void append(const SfmDMatch* m, int32_t n, std::vector<cv::DMatch>* out) {
        if (n <= 0) return;
        const size_t old = out->size();
        out->resize(old + static_cast<size_t>(n));
        std::memcpy(static_cast<void*>(out->data() + old), m, static_cast<size_t>(n) * sizeof(SfmDMatch));
    }
    void check(int rc) {
        if (rc != SFMM_OK) throw Error(rc, sfmm_last_error(ctx_));
    }
    SfmmCtx* ctx_;
};

// All GPUs of the box from ONE host process -- the shape of the reference (a single C++ program, main.cpp:18).
// Thin wrapper over the library's device group (sfmm_group_*, sfm_match.h): the descriptors are re-pitched and
// copied host->device once, broadcast to the other devices with NCCL over NVLink (single-process
// ncclCommInitAll inside the library), the q<t pairs are dealt by descending cost rows_q*rows_t in snake order,
// every device matches its shard and copies its records to host memory over its own PCIe link, and
// getMatching() is a look-up in the owning device's table.  (The multi-PROCESS variant -- one rank per GPU,
// torch.distributed -- lives in sfm_danpipeline_b200/distributed.py.)
class MultiGpuMatcher {
  public:
    explicit MultiGpuMatcher(int nDevices, int normType = cv::NORM_L2, float ratio = 0.8f, bool crossCheck = false) : g_(nullptr) {
        if (nDevices < 1) throw Error(SFMM_EINVAL, "MultiGpuMatcher: need at least one device");
        SfmmConfig cfg;
        sfmm_default_config(&cfg);
        cfg.norm = (normType == cv::NORM_HAMMING) ? SFMM_NORM_HAMMING : SFMM_NORM_L2;
        cfg.ratio = ratio;
        cfg.cross_check = crossCheck ? 1 : 0;
        const int rc = sfmm_group_create(&cfg, nDevices, nullptr, &g_);
        if (rc != SFMM_OK) throw Error(rc, sfmm_group_last_error(nullptr));
    }
    ~MultiGpuMatcher() { sfmm_group_destroy(g_); }
    MultiGpuMatcher(const MultiGpuMatcher&) = delete;
    MultiGpuMatcher& operator=(const MultiGpuMatcher&) = delete;

    // Call once after extractFeature() (src/Sfm.cpp:21), like AllPairsMatcher::compute.
    void compute(const std::vector<cv::Mat>& imagesDescriptors) {
        std::vector<const void*> data;
        std::vector<int32_t> rows;
        std::vector<size_t> steps;
        int cols = 0, type = SFMM_F32;
        describe(imagesDescriptors, data, rows, steps, cols, type);
        check(sfmm_group_set_descriptors(g_, static_cast<int32_t>(rows.size()), data.data(), rows.data(), cols, steps.data(), type));
        check(sfmm_group_match_all_pairs(g_));
    }

    // Drop-in body of StructFromMotion::getMatching (appends).  Pairs outside the q<t table are computed on demand by member 0.
    void getMatching(const int& idx_query, const int& idx_train, std::vector<cv::DMatch>* goodMatches) {
        const SfmDMatch* m = nullptr;
        int32_t n = 0;
        const int rc = sfmm_group_get_pair(g_, idx_query, idx_train, &m, &n);
        if (rc == SFMM_ESTATE) {
            SfmmCtx* c0 = sfmm_group_context(g_, 0);
            int32_t nq = 0;
            if (sfmm_image_rows(c0, idx_query, &nq) != SFMM_OK) throw Error(SFMM_ERANGE, "getMatching: image index out of range");
            std::vector<SfmDMatch> tmp(static_cast<size_t>(nq > 0 ? nq : 1));
            const int rc2 = sfmm_match_pair(c0, idx_query, idx_train, tmp.data(), static_cast<int32_t>(tmp.size()), &n);
            if (rc2 != SFMM_OK) throw Error(rc2, sfmm_last_error(c0));
            m = tmp.data();
            appendTo(m, n, goodMatches);
            return;
        }
        if (rc != SFMM_OK) throw Error(rc, "getMatching: image index out of range");
        appendTo(m, n, goodMatches);
    }

    int devices() const { return static_cast<int>(sfmm_group_size(g_)); }
    SfmmGroup* handle() { return g_; }

  private:
    static void appendTo(const SfmDMatch* m, int32_t n, std::vector<cv::DMatch>* out) {
        if (n <= 0) return;
        const size_t old = out->size();
        out->resize(old + static_cast<size_t>(n));
        std::memcpy(static_cast<void*>(out->data() + old), m, static_cast<size_t>(n) * sizeof(SfmDMatch));
    }
    void check(int rc) {
        if (rc != SFMM_OK) throw Error(rc, sfmm_group_last_error(g_));
    }
    SfmmGroup* g_;
};

// Drop-in for the ORB branch of StructFromMotion::getFeature (src/Sfm.cpp:358-384): replaces
//     cv::Ptr<cv::ORB> detector = cv::ORB::create(500, 1.2f, 8, 31, 0, 2, cv::ORB::HARRIS_SCORE, 31, 20);
//     detector->detectAndCompute(image, cv::noArray(), kps, descriptors, false);
// with the GPU extractor of sfm_features.h (the reference's parameters are built in).  Same keypoints on every pyramid level,
// bit-identical angle / response / size / descriptor; the order inside a level is row-major instead of std::nth_element's.
static_assert(sizeof(cv::KeyPoint) == sizeof(SfmKeyPoint), "cv::KeyPoint and SfmKeyPoint must have the same layout");
class OrbExtractor {
  public:
    explicit OrbExtractor(int device = 0) : orb_(nullptr) {
        const int rc = sfmm_orb_create(device, &orb_);
        if (rc != SFMM_OK) throw Error(rc, sfmm_orb_last_error(nullptr));
    }
    ~OrbExtractor() { sfmm_orb_destroy(orb_); }
    OrbExtractor(const OrbExtractor&) = delete;
    OrbExtractor& operator=(const OrbExtractor&) = delete;

    void detectAndCompute(const cv::Mat& image, std::vector<cv::KeyPoint>& kps, cv::Mat& descriptors) {
        if (image.empty() || image.depth() != CV_8U || (image.channels() != 1 && image.channels() != 3))
            throw Error(SFMM_EINVAL, "OrbExtractor: 8-bit gray or BGR image expected");
        std::vector<SfmKeyPoint> k(1024);
        std::vector<uint8_t> d(1024 * 32);
        int32_t n = 0;
        int rc = sfmm_orb_detect_and_compute(orb_, image.data, image.rows, image.cols, image.step, image.channels(), k.data(), d.data(), 1024, &n);
        if (rc == SFMM_ERANGE && n > 1024) {  // more than 1024 keypoints: only through ties
            k.resize(n);
            d.resize(static_cast<size_t>(n) * 32);
            rc = sfmm_orb_detect_and_compute(orb_, image.data, image.rows, image.cols, image.step, image.channels(), k.data(), d.data(), n, &n);
        }
        if (rc != SFMM_OK) throw Error(rc, sfmm_orb_last_error(orb_));
        kps.resize(static_cast<size_t>(n));
        if (n > 0) std::memcpy(static_cast<void*>(kps.data()), k.data(), static_cast<size_t>(n) * sizeof(SfmKeyPoint));
        if (n > 0) {
            descriptors.create(n, 32, CV_8U);
            for (int i = 0; i < n; ++i) std::memcpy(descriptors.data + static_cast<size_t>(i) * descriptors.step, d.data() + static_cast<size_t>(i) * 32, 32);
        } else {
            descriptors.release();
        }
    }

  private:
    SfmmOrb* orb_;
};

}  // namespace sfmm
#endif  // SFM_MATCH_OPENCV_HPP_
