// Drives include/sfm_match_opencv.hpp the way a patched StructFromMotion would: imagesDescriptors in,
// getMatching(q,t,&vector<cv::DMatch>) out (append semantics).  Reads descriptor sets and the
// expected lists from a flat binary file written by tests/test_cpp_adapter.py (oracle output).
// exit 0 = identical, 1 = mismatch, 2 = usage/io, 77 = no GPU.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "sfm_match_opencv.hpp"

static bool rd(FILE* f, void* p, size_t n) { return fread(p, 1, n, f) == n; }

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    int32_t hdr[5];  // n_images, cols, depth(0=u8,1=f32), cross, norm(0=HAMMING,1=L2; L2 over CV_8U rows = the reference's literal call)
    if (!rd(f, hdr, sizeof hdr)) return 2;
    const int n = hdr[0], cols = hdr[1], depth = hdr[2] ? CV_32F : CV_8U, esz = hdr[2] ? 4 : 1;
    std::vector<int32_t> rows(n);
    if (!rd(f, rows.data(), sizeof(int32_t) * n)) return 2;
    std::vector<std::vector<unsigned char>> store(n);
    std::vector<cv::Mat> imagesDescriptors(n);
    for (int i = 0; i < n; ++i) {
        store[i].resize(static_cast<size_t>(rows[i]) * cols * esz);
        if (rows[i] && !rd(f, store[i].data(), store[i].size())) return 2;
        if (rows[i]) imagesDescriptors[i] = cv::Mat(rows[i], cols, depth, store[i].data());
    }
    try {
        sfmm::AllPairsMatcher matcher(hdr[4] ? cv::NORM_L2 : cv::NORM_HAMMING, 0.8f, hdr[3] != 0, 0);
        matcher.compute(imagesDescriptors);
        for (int q = 0; q < n - 1; ++q)
            for (int t = q + 1; t < n; ++t) {
                int32_t cnt;
                if (!rd(f, &cnt, 4)) return 2;
                std::vector<cv::DMatch> expect(cnt);
                if (cnt && !rd(f, expect.data(), sizeof(cv::DMatch) * cnt)) return 2;
                std::vector<cv::DMatch> good(1);  // pre-existing element: getMatching must append
                matcher.getMatching(q, t, &good);
                if (static_cast<int>(good.size()) != cnt + 1 || good[0].queryIdx != -1) { printf("size mismatch %d,%d\n", q, t); return 1; }
                if (cnt && memcmp(static_cast<const void*>(good.data() + 1), expect.data(), sizeof(cv::DMatch) * cnt)) { printf("content mismatch %d,%d\n", q, t); return 1; }
            }
        {   // next rows: aligned points gathered on the GPU + persisted table, through the adapter
            std::vector<std::vector<cv::Point2d> > pts(n);
            for (int i = 0; i < n; ++i)
                for (int r = 0; r < rows[i]; ++r) pts[i].push_back(cv::Point2d(i * 1000.0 + r, -0.5 * r));
            sfmm::AllPairsMatcher m2(hdr[4] ? cv::NORM_L2 : cv::NORM_HAMMING, 0.8f, hdr[3] != 0, 0);
            m2.compute(imagesDescriptors, pts);
            for (int q = 0; q < n - 1; ++q)
                for (int t = q + 1; t < n; ++t) {
                    std::vector<cv::DMatch> mm;
                    std::vector<cv::Point2d> L, R;
                    m2.getMatching(q, t, &mm);
                    m2.getAlignedPoints(q, t, L, R);
                    if (L.size() != mm.size() || R.size() != mm.size()) { printf("aligned size mismatch %d,%d\n", q, t); return 1; }
                    for (size_t i = 0; i < mm.size(); ++i)  // AlignedPoints, src/Sfm.cpp:700-711
                        if (L[i].x != pts[q][mm[i].queryIdx].x || L[i].y != pts[q][mm[i].queryIdx].y ||
                            R[i].x != pts[t][mm[i].trainIdx].x || R[i].y != pts[t][mm[i].trainIdx].y) { printf("aligned mismatch %d,%d\n", q, t); return 1; }
                }
            const std::string path = std::string(argv[1]) + ".tbl";
            m2.saveTable(path);
            sfmm::AllPairsMatcher m3(hdr[4] ? cv::NORM_L2 : cv::NORM_HAMMING, 0.8f, hdr[3] != 0, 0);
            m3.loadTable(imagesDescriptors, path);
            for (int q = 0; q < n - 1; ++q)
                for (int t = q + 1; t < n; ++t) {
                    std::vector<cv::DMatch> a, b;
                    m2.getMatching(q, t, &a);
                    m3.getMatching(q, t, &b);
                    if (a.size() != b.size() || (a.size() && memcmp(static_cast<const void*>(a.data()), b.data(), a.size() * sizeof(cv::DMatch)))) { printf("table mismatch %d,%d\n", q, t); return 1; }
                }
        }
        {   // single process, several devices (as many as the box has, at most 2 here): same lists
            int n_dev = 1;
            if (argc > 2) n_dev = atoi(argv[2]);
            sfmm::MultiGpuMatcher multi(n_dev, hdr[4] ? cv::NORM_L2 : cv::NORM_HAMMING, 0.8f, hdr[3] != 0);
            multi.compute(imagesDescriptors);
            sfmm::AllPairsMatcher single(hdr[4] ? cv::NORM_L2 : cv::NORM_HAMMING, 0.8f, hdr[3] != 0, 0);
            single.compute(imagesDescriptors);
            for (int q = 0; q < n - 1; ++q)
                for (int t = q + 1; t < n; ++t) {
                    std::vector<cv::DMatch> a, b;
                    multi.getMatching(q, t, &a);
                    single.getMatching(q, t, &b);
                    if (a.size() != b.size() || (a.size() && memcmp(static_cast<const void*>(a.data()), b.data(), a.size() * sizeof(cv::DMatch)))) { printf("multi-gpu mismatch %d,%d\n", q, t); return 1; }
                }
            printf("multi-gpu ok on %d device(s)\n", multi.devices());
        }
        {   // a mixed set (different widths) must be refused, not read out of bounds
            std::vector<cv::Mat> bad = imagesDescriptors;
            std::vector<unsigned char> narrow(64, 0);
            bad.push_back(cv::Mat(2, cols > 8 ? cols - 8 : cols + 8, depth, narrow.data()));
            bool refused = false;
            try { matcher.upload(bad); } catch (const sfmm::Error& e) { refused = e.code == SFMM_EINVAL; }
            if (!refused) { printf("mixed descriptor widths were accepted\n"); return 1; }
            matcher.compute(imagesDescriptors);
        }
        if (argc > 3) {  // ORB extraction through the adapter: argv[3] = raw 8-bit gray image file "rows cols" header + pixels, expected count
            FILE* g = fopen(argv[3], "rb");
            int32_t dims[3];
            if (!g || !rd(g, dims, sizeof dims)) return 2;
            std::vector<unsigned char> pix(static_cast<size_t>(dims[0]) * dims[1]);
            if (!rd(g, pix.data(), pix.size())) return 2;
            fclose(g);
            cv::Mat image(dims[0], dims[1], CV_8U, pix.data());
            sfmm::OrbExtractor orb(0);
            std::vector<cv::KeyPoint> kps;
            cv::Mat desc;
            orb.detectAndCompute(image, kps, desc);
            if (static_cast<int>(kps.size()) != dims[2] || desc.rows != dims[2] || desc.cols != 32) { printf("orb count mismatch %zu\n", kps.size()); return 1; }
            for (size_t i = 0; i < kps.size(); ++i)
                if (kps[i].class_id != -1 || kps[i].octave < 0 || kps[i].octave > 7 || kps[i].angle < 0.f || kps[i].angle >= 360.f) { printf("orb keypoint fields\n"); return 1; }
            printf("orb ok: %zu keypoints\n", kps.size());
        }
        std::vector<cv::DMatch> rev;  // q>t is never asked by the reference; the adapter computes it on demand
        matcher.getMatching(1, 0, &rev);
        printf("adapter ok: %d images, reverse pair gave %zu matches\n", n, rev.size());
    } catch (const sfmm::Error& e) {
        printf("sfmm::Error %d: %s\n", e.code, e.what());
        return e.code == SFMM_ENODEVICE ? 77 : 1;
    }
    return 0;
}
