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

#include "sfm_match.h"

namespace sfmm {

static_assert(sizeof(cv::DMatch) == sizeof(SfmDMatch), "cv::DMatch and SfmDMatch must have the same layout");

// The reference has no error channel on this path (void functions; cv::Exception propagates):
// ABI error codes become exceptions of the same spirit.
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

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
        const int n = static_cast<int>(imagesDescriptors.size());
        std::vector<const void*> data(n);
        std::vector<int32_t> rows(n);
        std::vector<size_t> steps(n);
        int cols = 1, type = SFMM_F32;
        for (int i = 0; i < n; ++i) {
            const cv::Mat& m = imagesDescriptors[i];
            if (m.empty()) {  // an image without keypoints: cv::Mat() -- zero rows
                data[i] = nullptr; rows[i] = 0; steps[i] = 0;
                continue;
            }
            if (m.depth() != CV_8U && m.depth() != CV_32F) throw Error(SFMM_EINVAL, "descriptors must be CV_8U or CV_32F");
            cols = m.cols;
            type = (m.depth() == CV_8U) ? SFMM_U8 : SFMM_F32;
            data[i] = m.data; rows[i] = m.rows; steps[i] = m.step;
        }
        check(sfmm_set_descriptors(ctx_, n, data.data(), rows.data(), cols, steps.data(), type));
    }

  public:

    // Drop-in body of StructFromMotion::getMatching: APPENDS to *goodMatches like the
    // reference's push_back loop (src/Sfm.cpp:603-607); no clear().
    void getMatching(const int& idx_query, const int& idx_train, std::vector<cv::DMatch>* goodMatches) {
        const SfmDMatch* m = nullptr;
        int32_t n = 0;
        int rc = sfmm_get_pair(ctx_, idx_query, idx_train, &m, &n);
        if (rc == SFMM_ESTATE) {  // a pair outside the q<t table (never requested by the reference): on demand
            std::vector<SfmDMatch> tmp(1);
            rc = sfmm_match_pair(ctx_, idx_query, idx_train, tmp.data(), 0, &n);
            if (rc != SFMM_ERANGE && rc != SFMM_OK) check(rc);
            tmp.resize(n > 0 ? n : 1);
            check(sfmm_match_pair(ctx_, idx_query, idx_train, tmp.data(), static_cast<int32_t>(tmp.size()), &n));
            append(tmp.data(), n, goodMatches);
            return;
        }
        check(rc);
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
        check(sfmm_get_pair_points(ctx_, idx_query, idx_train, &l, &r, &n));
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

// All GPUs of the box from ONE host process -- the shape of the reference (a single C++ program).
// Every device gets the descriptors (its own H2D over its own PCIe link, in parallel host threads),
// the q<t pairs are dealt by descending cost rows_q*rows_t in a snake order (the same rule as
// sfm_danpipeline_b200/distributed.py, so shards are balanced and deterministic), each device
// matches its shard into its own host table, and getMatching() looks the pair up in the owning
// context.  No inter-GPU traffic is needed: results are wanted in host memory anyway.  (The
// multi-PROCESS variant -- NCCL broadcast + gather to rank 0 -- lives in distributed.py.)
class MultiGpuMatcher {
  public:
    explicit MultiGpuMatcher(int nDevices, int normType = cv::NORM_L2, float ratio = 0.8f, bool crossCheck = false) {
        if (nDevices < 1) throw Error(SFMM_EINVAL, "MultiGpuMatcher: need at least one device");
        for (int d = 0; d < nDevices; ++d) dev_.emplace_back(new AllPairsMatcher(normType, ratio, crossCheck, d));
    }

    void compute(const std::vector<cv::Mat>& imagesDescriptors) {
        const int n = static_cast<int>(imagesDescriptors.size()), nd = static_cast<int>(dev_.size());
        // findBestPair's enumeration (src/Sfm.cpp:511-512), then the cost-sorted snake deal
        std::vector<std::pair<int, int> > pairs;
        for (int q = 0; q + 1 < n; ++q)
            for (int t = q + 1; t < n; ++t) pairs.push_back(std::make_pair(q, t));
        std::vector<size_t> order(pairs.size());
        std::iota(order.begin(), order.end(), size_t(0));
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) {
            const long long ca = 1LL * imagesDescriptors[pairs[a].first].rows * imagesDescriptors[pairs[a].second].rows;
            const long long cb = 1LL * imagesDescriptors[pairs[b].first].rows * imagesDescriptors[pairs[b].second].rows;
            return ca > cb;
        });
        owner_.assign(static_cast<size_t>(n) * n, -1);
        std::vector<std::vector<int32_t> > shard(nd);
        std::vector<std::vector<size_t> > members(nd);
        for (size_t pos = 0; pos < order.size(); ++pos) {
            const size_t lap = pos / nd, off = pos % nd;
            members[(lap % 2 == 0) ? off : nd - 1 - off].push_back(order[pos]);
        }
        for (int d = 0; d < nd; ++d) {
            std::sort(members[d].begin(), members[d].end());  // ascending pair order inside a device
            for (size_t k : members[d]) {
                shard[d].push_back(pairs[k].first);
                shard[d].push_back(pairs[k].second);
                owner_[static_cast<size_t>(pairs[k].first) * n + pairs[k].second] = d;
            }
        }
        n_images_ = n;
        std::vector<std::string> errors(nd);
        std::vector<int> codes(nd, SFMM_OK);
        std::vector<std::thread> pool;
        for (int d = 0; d < nd; ++d)
            pool.emplace_back([&, d]() {
                try {
                    dev_[d]->upload(imagesDescriptors);
                    dev_[d]->matchPairs(shard[d]);
                } catch (const Error& e) {
                    codes[d] = e.code;
                    errors[d] = e.what();
                }
            });
        for (auto& t : pool) t.join();
        for (int d = 0; d < nd; ++d)
            if (codes[d] != SFMM_OK) throw Error(codes[d], "device " + std::to_string(d) + ": " + errors[d]);
    }

    // Drop-in body of StructFromMotion::getMatching (appends).
    void getMatching(const int& idx_query, const int& idx_train, std::vector<cv::DMatch>* goodMatches) {
        int d = 0;
        if (idx_query >= 0 && idx_train >= 0 && idx_query < n_images_ && idx_train < n_images_) {
            const int o = owner_[static_cast<size_t>(idx_query) * n_images_ + idx_train];
            if (o >= 0) d = o;  // pairs outside the table are computed on demand by device 0
        }
        dev_[d]->getMatching(idx_query, idx_train, goodMatches);
    }

    int devices() const { return static_cast<int>(dev_.size()); }

  private:
    std::vector<std::unique_ptr<AllPairsMatcher> > dev_;
    std::vector<int> owner_;
    int n_images_ = 0;
};

}  // namespace sfmm
#endif  // SFM_MATCH_OPENCV_HPP_
