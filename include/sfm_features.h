/*
 * sfm_features.h -- C ABI of the descriptor-extraction step in front of the matching path,
 * part of libsfmmatch.so (sm_100a, no CPU fallback).
 *
 * Reference interface being replaced (file:line relative to the reference tree):
 *
 *     void StructFromMotion::getFeature(const cv::Mat& image, const int& numImage)      src/Sfm.cpp:303-392
 *         detector == 3:  cv::ORB::create(500, 1.2f, 8, 31, 0, 2, cv::ORB::HARRIS_SCORE, 31, 20)   src/Sfm.cpp:360-371
 *                         detector->detectAndCompute(image, cv::noArray(), kps, descriptors, false)  src/Sfm.cpp:373
 *         imagesKeypoints[numImage] = kps;  imagesDescriptors[numImage] = descriptors;           src/Sfm.cpp:380-381
 *
 * Only the ORB branch is built (the cheapest of the three detectors and the one whose CV_8U x 32 rows the binary matching
 * kernels consume directly); SIFT (:307-329) and AKAZE (:331-356) stay on the CPU.  Parameters are the reference's
 * hard-wired ones.  Results equal cv::ORB's on the same image: the same keypoints on every pyramid level and, at each of
 * them, bit-identical angle, Harris response, size and 256-bit descriptor (tests/test_orb_gpu.py, cv2 golden on data/temple).
 * The ORDER of keypoints inside a pyramid level is not reproduced: OpenCV leaves it to std::nth_element
 * (KeyPointsFilter::retainBest); here it is level by level, row-major.
 */
#ifndef SFM_FEATURES_H_
#define SFM_FEATURES_H_

#include <stddef.h>
#include <stdint.h>

#include "sfm_match.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Layout-identical to cv::KeyPoint {Point2f pt; float size; float angle; float response; int octave; int class_id;}
 * (the element type of `Keypoints`, include/Utilities.h:26), so a list can be copied into std::vector<cv::KeyPoint>. */
typedef struct SfmKeyPoint {
    float x, y;       /* pt, in coordinates of the input image */
    float size;       /* 31 * 1.2^octave */
    float angle;      /* degrees, [0, 360) */
    float response;   /* Harris response */
    int32_t octave;   /* pyramid level 0..7 */
    int32_t class_id; /* -1 */
} SfmKeyPoint;

typedef struct SfmmOrb SfmmOrb;

/* Replaces cv::ORB::create(...) with the reference's arguments (src/Sfm.cpp:360-371).  One extractor per device. */
SFMM_API int sfmm_orb_create(int32_t device, SfmmOrb** out);
SFMM_API void sfmm_orb_destroy(SfmmOrb* orb);
SFMM_API const char* sfmm_orb_last_error(const SfmmOrb* orb);

/* Replaces detector->detectAndCompute(image, noArray(), kps, descriptors, false) (src/Sfm.cpp:373) for one image:
 * rows x cols pixels of `channels` (1 = gray, 3 = BGR as cv::imread returns it; converted like cv::cvtColor(BGR2GRAY))
 * 8-bit samples, row step `step_bytes`.  Writes at most `capacity` keypoints and `capacity` x 32 descriptor bytes
 * (descriptor i belongs to keypoint i); *count is the number found -- SFMM_ERANGE when it exceeds `capacity`
 * (nothing is written then; nfeatures = 500 plus ties: 1024 always suffices for the reference's parameters). */
SFMM_API int sfmm_orb_detect_and_compute(SfmmOrb* orb, const uint8_t* image, int32_t rows, int32_t cols, size_t step_bytes,
                                         int32_t channels, SfmKeyPoint* keypoints, uint8_t* descriptors, int32_t capacity,
                                         int32_t* count);

/* Kernel launches so far and the device time (CUDA events) of the last sfmm_orb_detect_and_compute. */
SFMM_API int sfmm_orb_stats(const SfmmOrb* orb, int64_t* kernel_launches, double* last_ms);

#ifdef __cplusplus
}
#endif
#endif /* SFM_FEATURES_H_ */
