/*
 * oracle/bf_oracle.c -- CPU restatement of the reference's matching hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (sfm_danpipeline_b200/csrc) never links, imports or falls back to anything here.
 *
 * What it restates (all file:line relative to /root/reference):
 *   - StructFromMotion::getMatching            src/Sfm.cpp:590-608
 *       cv::BFMatcher(norm,false).knnMatch(q,t,knn,2)      src/Sfm.cpp:593,599
 *       ratio test  d1 <= NN_MATCH_RATIO * d2 (fp32, <=)   src/Sfm.cpp:603-607
 *       NN_MATCH_RATIO = 0.8f                              include/Sfm.h:60
 *   - findBestPair's all-pairs enumeration q<t             src/Sfm.cpp:511-515
 *
 * The arithmetic itself lives in a third-party dependency that is NOT vendored under
 * /root/reference: OpenCV 3.4.1 (README.md:29, CMakeLists.txt:20,46), modules
 * features2d (BFMatcher::knnMatchImpl) and core (batchDistance, hal::normHamming,
 * normL2Sqr_).  Its published algorithm, restated here:
 *   batchDistance(K=2): for every query row, distance to every train row in ascending
 *   train order; a sorted K-list initialised to (INT_MAX|FLT_MAX, -1) is updated by
 *   insertion with a STRICT '<' -- so on equal distances the LOWEST train index wins,
 *   for the first and for the second neighbour alike.
 *   NORM_HAMMING: sum popcount(a^b) over `cols` bytes, int32 -> float in the DMatch.
 *   NORM_L2:      sqrtf( sum (a-b)^2 ) accumulated in fp32 in direct-difference form.
 *
 * Parity pin: the reference has no tests or golden vectors of its own, so
 * the pin is the OpenCV code itself, reached through the cv2 4.13.0 wheel of this image:
 * tests/golden/make_golden.py ran cv2.BFMatcher.knnMatch on descriptors extracted
 * from /root/reference/data/temple and on seeded synthetic sets and committed the
 * results; tests/test_oracle.py checks this file against those fixtures (and against
 * cv2 live, when importable).
 *
 * The symmetric cross-check is NOT in the reference (crossCheck=false, src/Sfm.cpp:593);
 * it is the north-star extra stage, defined as membership in
 * cv2.BFMatcher(norm, crossCheck=True).match(Q,T): keep (q,t) iff q is the lowest-index
 * argmin over q' of d(q',t).
 */
#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int32_t queryIdx;
    int32_t trainIdx;
    int32_t imgIdx;
    float distance;
} OracleDMatch; /* layout of cv::DMatch (include/Utilities.h:27 -> std::vector<cv::DMatch>) */

#define ORACLE_NORM_HAMMING 0
#define ORACLE_NORM_L2 1

#if defined(__x86_64__) && defined(__GNUC__) && !defined(ORACLE_NO_CLONES)
#define ORACLE_CLONES __attribute__((target_clones("default", "popcnt", "avx2", "arch=x86-64-v4")))
#else
#define ORACLE_CLONES
#endif

/* hal::normHamming: popcount of the XOR over `cols` bytes. */
static inline int hamming_row(const uint8_t* a, const uint8_t* b, int cols) {
    int d = 0, j = 0;
    for (; j + 8 <= cols; j += 8) {
        uint64_t x, y;
        memcpy(&x, a + j, 8);
        memcpy(&y, b + j, 8);
        d += __builtin_popcountll(x ^ y);
    }
    for (; j < cols; ++j) d += __builtin_popcount((unsigned)(a[j] ^ b[j]));
    return d;
}

/* normL2Sqr_ (fp32, direct difference) followed by sqrtf.  Eight partial sums mirror
 * the SIMD-lane accumulation of the OpenCV build; the last ulp is build dependent and
 * is covered by the 1e-4 relative tolerance of the float path. */
static inline float l2_row(const float* a, const float* b, int cols) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int j = 0;
    for (; j + 8 <= cols; j += 8)
        for (int l = 0; l < 8; ++l) {
            float t = a[j + l] - b[j + l];
            acc[l] += t * t;
        }
    float s = ((acc[0] + acc[4]) + (acc[2] + acc[6])) + ((acc[1] + acc[5]) + (acc[3] + acc[7]));
    for (; j < cols; ++j) {
        float t = a[j] - b[j];
        s += t * t;
    }
    return sqrtf(s);
}

/*
 * batchDistance(Q, T, K=2, NORM_HAMMING): dist/idx are [nq][2]; unfilled slots keep
 * (INT_MAX, -1) exactly like OpenCV's initialisation.
 */
ORACLE_CLONES
void oracle_knn2_hamming(const uint8_t* Q, int nq, size_t qstep, const uint8_t* T, int nt,
                         size_t tstep, int cols, int32_t* dist, int32_t* idx) {
    for (int i = 0; i < nq; ++i) {
        const uint8_t* q = Q + (size_t)i * qstep;
        int32_t d0 = INT_MAX, d1 = INT_MAX, i0 = -1, i1 = -1;
        for (int j = 0; j < nt; ++j) {
            int32_t d = hamming_row(q, T + (size_t)j * tstep, cols);
            if (d < d0) {
                d1 = d0; i1 = i0; d0 = d; i0 = j;
            } else if (d < d1) {
                d1 = d; i1 = j;
            }
        }
        dist[2 * i] = d0; dist[2 * i + 1] = d1;
        idx[2 * i] = i0;  idx[2 * i + 1] = i1;
    }
}

/* batchDistance(Q, T, K=2, NORM_L2) on CV_32F rows; steps are in BYTES. */
ORACLE_CLONES
void oracle_knn2_l2(const float* Q, int nq, size_t qstep, const float* T, int nt, size_t tstep,
                    int cols, float* dist, int32_t* idx) {
    for (int i = 0; i < nq; ++i) {
        const float* q = (const float*)((const char*)Q + (size_t)i * qstep);
        float d0 = FLT_MAX, d1 = FLT_MAX;
        int32_t i0 = -1, i1 = -1;
        for (int j = 0; j < nt; ++j) {
            float d = l2_row(q, (const float*)((const char*)T + (size_t)j * tstep), cols);
            if (d < d0) {
                d1 = d0; i1 = i0; d0 = d; i0 = j;
            } else if (d < d1) {
                d1 = d; i1 = j;
            }
        }
        dist[2 * i] = d0; dist[2 * i + 1] = d1;
        idx[2 * i] = i0;  idx[2 * i + 1] = i1;
    }
}

/* For every train row: the lowest-index query row at minimum distance (K=1 of the
 * transposed problem) -- the other half of BFMatcher(crossCheck=true). */
ORACLE_CLONES
void oracle_colmin_hamming(const uint8_t* Q, int nq, size_t qstep, const uint8_t* T, int nt,
                           size_t tstep, int cols, int32_t* best_q) {
    for (int j = 0; j < nt; ++j) {
        const uint8_t* t = T + (size_t)j * tstep;
        int32_t bd = INT_MAX, bi = -1;
        for (int i = 0; i < nq; ++i) {
            int32_t d = hamming_row(t, Q + (size_t)i * qstep, cols);
            if (d < bd) { bd = d; bi = i; }
        }
        best_q[j] = bi;
    }
}

ORACLE_CLONES
void oracle_colmin_l2(const float* Q, int nq, size_t qstep, const float* T, int nt, size_t tstep,
                      int cols, int32_t* best_q) {
    for (int j = 0; j < nt; ++j) {
        const float* t = (const float*)((const char*)T + (size_t)j * tstep);
        float bd = FLT_MAX;
        int32_t bi = -1;
        for (int i = 0; i < nq; ++i) {
            /* BFMatcher(crossCheck=true) runs batchDistance(T,Q): the train row is the
             * first operand of the difference. (a-b)^2 == (b-a)^2 exactly in fp32. */
            float d = l2_row(t, (const float*)((const char*)Q + (size_t)i * qstep), cols);
            if (d < bd) { bd = d; bi = i; }
        }
        best_q[j] = bi;
    }
}

/*
 * getMatching (src/Sfm.cpp:590-608) for one ordered pair.
 *   norm        ORACLE_NORM_HAMMING (uint8 rows) or ORACLE_NORM_L2 (float rows)
 *   ratio       NN_MATCH_RATIO; the test is  d1 <= ratio*d2  with an fp32 product
 *   cross_check 0 = reference behaviour; 1 = add the mutual-nearest-neighbour filter
 * Writes at most nq matches to `out` in ascending queryIdx (the order the reference's
 * push_back loop produces) and returns how many.  Train sets with fewer than two rows
 * give zero matches: the reference indexes knnMatches[i][1] there (undefined behaviour,
 * src/Sfm.cpp:604); "no second neighbour => no ratio test => no match" is the defined
 * behaviour of the new path and the oracle states it the same way.
 */
int oracle_match_pair(const void* Q, int nq, size_t qstep, const void* T, int nt, size_t tstep,
                      int cols, int norm, float ratio, int cross_check, OracleDMatch* out) {
    if (nq <= 0 || nt < 2) return 0;
    int32_t* idx = (int32_t*)malloc(sizeof(int32_t) * 2 * (size_t)nq);
    float* fd = (float*)malloc(sizeof(float) * 2 * (size_t)nq);
    int32_t* best_q = NULL;
    if (norm == ORACLE_NORM_HAMMING) {
        int32_t* id = (int32_t*)malloc(sizeof(int32_t) * 2 * (size_t)nq);
        oracle_knn2_hamming((const uint8_t*)Q, nq, qstep, (const uint8_t*)T, nt, tstep, cols, id, idx);
        /* BFMatcher::knnMatchImpl converts the CV_32S distances to float */
        for (size_t k = 0; k < 2 * (size_t)nq; ++k) fd[k] = (float)id[k];
        free(id);
    } else {
        oracle_knn2_l2((const float*)Q, nq, qstep, (const float*)T, nt, tstep, cols, fd, idx);
    }
    if (cross_check) {
        best_q = (int32_t*)malloc(sizeof(int32_t) * (size_t)nt);
        if (norm == ORACLE_NORM_HAMMING)
            oracle_colmin_hamming((const uint8_t*)Q, nq, qstep, (const uint8_t*)T, nt, tstep, cols, best_q);
        else
            oracle_colmin_l2((const float*)Q, nq, qstep, (const float*)T, nt, tstep, cols, best_q);
    }
    int n = 0;
    for (int i = 0; i < nq; ++i) {
        volatile float rhs = ratio * fd[2 * i + 1]; /* fp32 product, no widening */
        if (fd[2 * i] <= rhs) {
            if (cross_check && best_q[idx[2 * i]] != i) continue;
            out[n].queryIdx = i;
            out[n].trainIdx = idx[2 * i];
            out[n].imgIdx = 0;
            out[n].distance = fd[2 * i];
            ++n;
        }
    }
    free(idx); free(fd); free(best_q);
    return n;
}

/*
 * findBestPair's loop (src/Sfm.cpp:511-515): every q<t, row-major.  Images are given as
 * arrays of row pointers/rows/steps.  `pair_counts` gets N(N-1)/2 entries; matches are
 * appended to `out` (capacity `cap` records); returns the total written or -1 on overflow.
 */
long oracle_all_pairs(const void* const* data, const int* rows, const size_t* steps, int n_images,
                      int cols, int norm, float ratio, int cross_check, OracleDMatch* out, long cap,
                      int32_t* pair_counts) {
    long total = 0;
    int p = 0;
    for (int q = 0; q < n_images - 1; ++q)
        for (int t = q + 1; t < n_images; ++t, ++p) {
            if (total + rows[q] > cap) return -1;
            int n = oracle_match_pair(data[q], rows[q], steps[q], data[t], rows[t], steps[t], cols,
                                      norm, ratio, cross_check, out + total);
            pair_counts[p] = n;
            total += n;
        }
    return total;
}
