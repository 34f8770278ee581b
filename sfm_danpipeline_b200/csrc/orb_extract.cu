// ORB feature extraction on sm_100a -- the step in front of the matching path.
//
// Replaces, for the reference's ORB branch, cv::ORB::create(500, 1.2f, 8, 31, 0, 2, HARRIS_SCORE, 31, 20) +
// detector->detectAndCompute(image, noArray(), kps, descriptors, false) at /root/reference/src/Sfm.cpp:360-373.
// The algorithm is OpenCV's (features2d orb.cpp / fast.cpp, imgproc resize / filter), restated in oracle/orb_oracle.py with
// every integer and fp32 operation spelled out; the kernels here follow that restatement operation by operation (explicit
// round-to-nearest intrinsics wherever the compiler could otherwise contract a multiply-add), so keypoints, angles, responses
// and descriptors are bit-identical to cv::ORB's.  It is HBM- / latency-bound byte and integer work on small images
// (640 x 480 in the reference's fixture): plain coalesced kernels, one launch per pyramid level and stage, no tensor cores.
//
//   level l (0..7), size round(W / 1.2^l) x round(H / 1.2^l), kept inside a 32-pixel reflect-101 frame:
//     orb_level0_kernel / orb_resize_kernel   gray conversion / INTER_LINEAR_EXACT (8.8 fixed point) from level l-1, frame included
//     orb_fast_score_kernel                   FAST-9/16 corner score of every pixel (threshold 20)
//     orb_nms_kernel                          3x3 non-maximum suppression, 31-pixel border filter -> candidate list + score histogram
//     orb_pick_fast_kernel                    the 2N strongest by FAST score (ties kept: threshold from the histogram) + Harris response
//     orb_blur_rows_kernel / _cols_kernel     7x7 sigma-2 Gaussian, separable fp32 with fused multiply-adds in OpenCV's order
//   all levels:
//     orb_pick_harris_kernel                  the N strongest by Harris response (ties kept), ordered level by level, row-major
//     orb_angle_kernel                        intensity-centroid orientation (integer moments, fastAtan2 polynomial)
//     orb_describe_kernel                     256 rotated comparisons on the blurred level -> 32 bytes
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/sfm_features.h"
#include "orb_pattern.h"

namespace {

constexpr int N_FEATURES = 500, N_LEVELS = 8, EDGE = 31, PATCH = 31, HALF_PATCH = 15, FAST_T = 20, BORDER = 32, HARRIS_BLOCK = 7;
constexpr int MAX_OUT = 4096;  // keypoints kept per image (500 + ties)

__constant__ signed char c_pattern[256 * 4];
__constant__ int c_umax[HALF_PATCH + 2];
__constant__ float c_gauss[7];

struct Level {
    int w, h, pitch;       // interior size; pitch of the framed image = w + 2 * BORDER
    float scale;           // 1.2^l as OpenCV computes it
    int quota;             // features wanted on this level
    unsigned char* ext;    // framed image, (h + 64) x pitch
    unsigned char* blur;   // same layout: blurred interior, unblurred frame
    unsigned char* score;  // w x h FAST scores
    float* rows;           // (h + 6) x w row-filtered intermediate of the blur
    int *ofsx, *ofsy;      // INTER_LINEAR_EXACT taps from level l-1: offset, weight of the second tap (8.8)
    short *cx1, *cy1;
    uint32_t* cand;        // FAST candidates after NMS: x | y << 16
    unsigned char* cand_score;
    uint32_t* pick;        // after the 2N-by-FAST selection
    float* pick_resp;      // their Harris responses
    int cand_cap;
};
struct LevelDev {  // what the kernels of the "all levels" stage need
    int w, h, pitch, quota;
    float scale;
    const unsigned char *ext, *blur;
    const uint32_t* pick;
    const float* pick_resp;
    unsigned char* flags;  // scratch, one byte per picked candidate
};
struct KeyInfo {  // one final keypoint before it becomes an SfmKeyPoint
    int x, y, level;
    float response;
};

__device__ __forceinline__ int reflect101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return i;
}

// ---- level 0: the input image (gray or BGR -> gray like cv::cvtColor: 15-bit fixed point) inside its reflect-101 frame
__global__ void orb_level0_kernel(const unsigned char* __restrict__ src, size_t step, int channels, int w, int h, unsigned char* __restrict__ ext,
                                  int pitch) {
    const int X = blockIdx.x * blockDim.x + threadIdx.x, Y = blockIdx.y * blockDim.y + threadIdx.y;
    if (X >= w + 2 * BORDER || Y >= h + 2 * BORDER) return;
    const int x = reflect101(X - BORDER, w), y = reflect101(Y - BORDER, h);
    const unsigned char* p = src + (size_t)y * step + (size_t)x * channels;
    unsigned v;
    if (channels == 1) v = p[0];
    else v = (p[0] * 3735u + p[1] * 19235u + p[2] * 9798u + (1u << 14)) >> 15;
    ext[(size_t)Y * pitch + X] = (unsigned char)v;
}

// ---- level l from level l-1: cv::resize(INTER_LINEAR_EXACT) on 8-bit data = 8.8 fixed-point taps per axis, 16.16 result rounded
__global__ void orb_resize_kernel(const unsigned char* __restrict__ prev /* interior origin */, int prev_pitch, int pw, int ph, const int* __restrict__ ofsx,
                                  const short* __restrict__ cx1, const int* __restrict__ ofsy, const short* __restrict__ cy1, int w, int h,
                                  unsigned char* __restrict__ ext, int pitch) {
    const int X = blockIdx.x * blockDim.x + threadIdx.x, Y = blockIdx.y * blockDim.y + threadIdx.y;
    if (X >= w + 2 * BORDER || Y >= h + 2 * BORDER) return;
    const int x = reflect101(X - BORDER, w), y = reflect101(Y - BORDER, h);
    const int ox = ofsx[x], oy = ofsy[y];
    const int ox1 = min(ox + 1, pw - 1), oy1 = min(oy + 1, ph - 1);
    const unsigned wx1 = cx1[x], wx0 = 256 - wx1, wy1 = cy1[y], wy0 = 256 - wy1;
    const unsigned top = prev[(size_t)oy * prev_pitch + ox] * wx0 + prev[(size_t)oy * prev_pitch + ox1] * wx1;
    const unsigned bot = prev[(size_t)oy1 * prev_pitch + ox] * wx0 + prev[(size_t)oy1 * prev_pitch + ox1] * wx1;
    ext[(size_t)Y * pitch + X] = (unsigned char)((top * wy0 + bot * wy1 + (1u << 15)) >> 16);
}

// ---- FAST-9/16: score = the largest t for which the pixel still is a corner = max over the 16 arcs of 9 ring pixels of
//      min(centre - ring) or of min(ring - centre), minus 1; kept when the pixel is a corner for t = 20
__global__ void orb_fast_score_kernel(const unsigned char* __restrict__ img /* interior origin */, int pitch, int w, int h, unsigned char* __restrict__ score) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    int out = 0;
    if (x >= 3 && y >= 3 && x < w - 3 && y < h - 3) {
        const int dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
        const int dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
        const unsigned char* c = img + (size_t)y * pitch + x;
        const int v = c[0];
        int d[25];
#pragma unroll
        for (int k = 0; k < 16; ++k) d[k] = v - c[dy[k] * pitch + dx[k]];
#pragma unroll
        for (int k = 0; k < 9; ++k) d[16 + k] = d[k];
        int dark = -512, bright = -512;
#pragma unroll
        for (int s = 0; s < 16; ++s) {
            int mn = d[s], mx = d[s];
#pragma unroll
            for (int j = 1; j < 9; ++j) {
                mn = min(mn, d[s + j]);
                mx = max(mx, d[s + j]);
            }
            dark = max(dark, mn);
            bright = max(bright, -mx);
        }
        const int m = max(dark, bright);
        if (m > FAST_T) out = m - 1;
    }
    score[(size_t)y * w + x] = (unsigned char)out;
}

// ---- 3x3 non-maximum suppression (strictly greater than all 8 neighbours), then cv's runByImageBorder(edgeThreshold)
__global__ void orb_nms_kernel(const unsigned char* __restrict__ score, int w, int h, uint32_t* __restrict__ cand, unsigned char* __restrict__ cand_score,
                               int cap, int* __restrict__ n_cand, int* __restrict__ hist /* 256 */) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x < EDGE || y < EDGE || x >= w - EDGE || y >= h - EDGE) return;
    const int s = score[(size_t)y * w + x];
    if (s == 0) return;
#pragma unroll
    for (int j = -1; j <= 1; ++j)
#pragma unroll
        for (int i = -1; i <= 1; ++i)
            if ((i || j) && score[(size_t)(y + j) * w + x + i] >= s) return;
    const int pos = atomicAdd(n_cand, 1);
    if (pos < cap) {
        cand[pos] = (uint32_t)x | ((uint32_t)y << 16);
        cand_score[pos] = (unsigned char)s;
    }
    atomicAdd(hist + s, 1);
}

// ---- KeyPointsFilter::retainBest(2 * quota) on the FAST score as a set (everything at least as strong as the 2N-th strongest),
//      and HarrisResponses(blockSize 7, k 0.04) for the survivors.  One block per level.
__global__ void orb_pick_fast_kernel(const unsigned char* __restrict__ img /* interior origin */, int pitch, const uint32_t* __restrict__ cand,
                                     const unsigned char* __restrict__ cand_score, const int* __restrict__ n_cand_p, int cap, const int* __restrict__ hist,
                                     int keep, uint32_t* __restrict__ pick, float* __restrict__ pick_resp, int* __restrict__ n_pick, float harris_k,
                                     float scale_sq_sq) {
    __shared__ int thr_sh;
    const int n = min(*n_cand_p, cap);
    if (threadIdx.x == 0) {
        int thr = 0;
        if (n > keep) {
            int acc = 0;
            thr = 256;  // keep == 0: nothing
            for (int s = 255; s > 0 && keep > 0; --s) {
                acc += hist[s];
                if (acc >= keep) {
                    thr = s;
                    break;
                }
            }
        }
        thr_sh = thr;
    }
    __syncthreads();
    const int thr = thr_sh;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if (cand_score[i] < thr) continue;
        const uint32_t xy = cand[i];
        const int x0 = xy & 0xFFFF, y0 = xy >> 16;
        int a = 0, b = 0, c = 0;
        const int r = HARRIS_BLOCK / 2;
        for (int dy = -r; dy <= r; ++dy)
            for (int dx = -r; dx <= r; ++dx) {
                const unsigned char* p = img + (size_t)(y0 + dy) * pitch + x0 + dx;
                const int ix = (p[1] - p[-1]) * 2 + (p[-pitch + 1] - p[-pitch - 1]) + (p[pitch + 1] - p[pitch - 1]);
                const int iy = (p[pitch] - p[-pitch]) * 2 + (p[pitch - 1] - p[-pitch - 1]) + (p[pitch + 1] - p[-pitch + 1]);
                a += ix * ix;
                b += iy * iy;
                c += ix * iy;
            }
        // ((float)a * b - (float)c * c - harris_k * ((float)a + b) * ((float)a + b)) * scale_sq_sq, one rounding per operation
        const float A = (float)a, B = (float)b, C = (float)c, S = __fadd_rn(A, B);
        const float resp = __fmul_rn(__fsub_rn(__fsub_rn(__fmul_rn(A, B), __fmul_rn(C, C)), __fmul_rn(__fmul_rn(harris_k, S), S)), scale_sq_sq);
        const int pos = atomicAdd(n_pick, 1);
        pick[pos] = xy;
        pick_resp[pos] = resp;
    }
}

// ---- retainBest(quota) on the Harris response as a set, per level; output level by level, row-major inside a level.  One block.
__global__ void orb_pick_harris_kernel(const LevelDev* __restrict__ lv, const int* __restrict__ n_pick /* per level */, KeyInfo* __restrict__ out,
                                       int* __restrict__ n_out, int max_out) {
    __shared__ int base_sh, kept_sh;
    if (threadIdx.x == 0) base_sh = 0;
    __syncthreads();
    for (int l = 0; l < N_LEVELS; ++l) {
        const int n = n_pick[l], quota = lv[l].quota;
        const uint32_t* pick = lv[l].pick;
        const float* resp = lv[l].pick_resp;
        unsigned char* keep = lv[l].flags;
        if (threadIdx.x == 0) kept_sh = 0;
        // kept <=> fewer than `quota` candidates are strictly stronger (ties with the last kept one stay, like retainBest)
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const float ri = resp[i];
            int stronger = 0;
            for (int j = 0; j < n; ++j) stronger += resp[j] > ri;
            keep[i] = stronger < quota;
        }
        __syncthreads();
        const int base = base_sh;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            if (!keep[i]) continue;
            const uint32_t ki = pick[i];  // y << 16 | x: row-major order
            int before = 0;
            for (int j = 0; j < n; ++j) before += keep[j] && pick[j] < ki;
            atomicAdd(&kept_sh, 1);
            if (base + before < max_out) {
                KeyInfo o;
                o.x = ki & 0xFFFF;
                o.y = ki >> 16;
                o.level = l;
                o.response = resp[i];
                out[base + before] = o;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) base_sh = base + kept_sh;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_out = base_sh;
}

// ---- orientation: intensity centroid over the radius-15 disc (integer moments), angle = cv::fastAtan2(m01, m10).  One warp per keypoint.
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float k = 57.29577951308232f;  // (float)(180 / CV_PI)
    const float p1 = __fmul_rn(0.9997878412794807f, k), p3 = __fmul_rn(-0.3258083974640975f, k), p5 = __fmul_rn(0.1555786518463281f, k),
                p7 = __fmul_rn(-0.04432655554792128f, k);
    const float eps = 2.220446049250313e-16f;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0.f) a = __fsub_rn(180.f, a);
    if (y < 0.f) a = __fsub_rn(360.f, a);
    return a;
}

__global__ void orb_angle_kernel(const LevelDev* __restrict__ lv, const KeyInfo* __restrict__ keys, const int* __restrict__ n_keys, int max_out,
                                 SfmKeyPoint* __restrict__ out) {
    const int kp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (kp >= min(*n_keys, max_out)) return;
    const KeyInfo ki = keys[kp];
    const LevelDev L = lv[ki.level];
    const unsigned char* c = L.ext + (size_t)(ki.y + BORDER) * L.pitch + ki.x + BORDER;
    int m01 = 0, m10 = 0;
    // 31 rows of the disc; lane = column offset u + 15 (31 of 32 lanes)
    const int u = lane - HALF_PATCH;
    if (lane < 2 * HALF_PATCH + 1) {
        m10 += u * c[u];
        for (int v = 1; v <= HALF_PATCH; ++v) {
            if (abs(u) <= c_umax[v]) {
                const int plus = c[v * L.pitch + u], minus = c[-v * L.pitch + u];
                m10 += u * (plus + minus);
                m01 += v * (plus - minus);
            }
        }
    }
    m01 = __reduce_add_sync(0xFFFFFFFFu, m01);
    m10 = __reduce_add_sync(0xFFFFFFFFu, m10);
    if (lane == 0) {
        SfmKeyPoint o;
        o.x = __fmul_rn((float)ki.x, L.scale);
        o.y = __fmul_rn((float)ki.y, L.scale);
        o.size = __fmul_rn((float)PATCH, L.scale);
        o.angle = fast_atan2_deg((float)m01, (float)m10);
        o.response = ki.response;
        o.octave = ki.level;
        o.class_id = -1;
        out[kp] = o;
    }
}

// ---- 7x7 sigma-2 Gaussian, separable fp32, in the order OpenCV's generic row filter / symmetric column filter evaluate it
__global__ void orb_blur_rows_kernel(const unsigned char* __restrict__ ext, int pitch, int w, int h, float* __restrict__ rows /* (h + 6) x w */) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y * blockDim.y + threadIdx.y;  // r = image row + 3
    if (x >= w || r >= h + 6) return;
    const unsigned char* p = ext + (size_t)(r - 3 + BORDER) * pitch + BORDER + x - 3;
    float acc = __fmul_rn(c_gauss[0], (float)p[0]);
#pragma unroll
    for (int i = 1; i < 7; ++i) acc = __fmaf_rn((float)p[i], c_gauss[i], acc);
    rows[(size_t)r * w + x] = acc;
}
__global__ void orb_blur_cols_kernel(const float* __restrict__ rows, int w, int h, unsigned char* __restrict__ blur /* framed */, int pitch) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const float* p = rows + (size_t)(y + 3) * w + x;
    float acc = __fmul_rn(c_gauss[3], p[0]);
#pragma unroll
    for (int j = 1; j <= 3; ++j) acc = __fmaf_rn(__fadd_rn(p[-j * w], p[j * w]), c_gauss[3 + j], acc);
    const int v = __float2int_rn(acc);
    blur[(size_t)(y + BORDER) * pitch + BORDER + x] = (unsigned char)min(max(v, 0), 255);
}

// ---- rBRIEF: 256 comparisons at the pattern points rotated by the keypoint's angle.  One warp per keypoint, one byte per lane.
__global__ void orb_describe_kernel(const LevelDev* __restrict__ lv, const KeyInfo* __restrict__ keys, const SfmKeyPoint* __restrict__ kps,
                                    const int* __restrict__ n_keys, int max_out, unsigned char* __restrict__ desc) {
    const int kp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (kp >= min(*n_keys, max_out)) return;
    const KeyInfo ki = keys[kp];
    const LevelDev L = lv[ki.level];
    const float rad = __fmul_rn(kps[kp].angle, 0.017453292519943295f);  // angle *= (float)(CV_PI / 180.f)
    const float a = (float)cos((double)rad), b = (float)sin((double)rad);
    const unsigned char* c = L.blur + (size_t)(ki.y + BORDER) * L.pitch + ki.x + BORDER;
    unsigned byte = 0;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const signed char* p = c_pattern + (lane * 8 + t) * 4;
        const float x0 = (float)p[0], y0 = (float)p[1], x1 = (float)p[2], y1 = (float)p[3];
        const int ix0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b))), iy0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
        const int ix1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b))), iy1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
        const int t0 = c[iy0 * L.pitch + ix0], t1 = c[iy1 * L.pitch + ix1];
        byte |= (t0 < t1 ? 1u : 0u) << t;
    }
    desc[(size_t)kp * 32 + lane] = (unsigned char)byte;
}

struct Buf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        const cudaError_t e = cudaMalloc(&p, bytes + 256);
        if (e == cudaSuccess) cap = bytes + 256;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

thread_local std::string g_orb_create_error;

}  // namespace

struct SfmmOrb {
    int device = 0;
    cudaStream_t st = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int rows = 0, cols = 0;  // geometry the level buffers are laid out for
    Level lv[N_LEVELS];
    Buf pool;        // every per-level array lives in one allocation
    Buf d_src;       // the input image
    Buf d_counts;    // n_cand[8], n_pick[8], n_out, hist[8][256]
    Buf d_lv, d_keys, d_kps, d_desc;
    void* h_pin = nullptr;  // pinned staging for the image
    size_t h_pin_cap = 0;
    void* h_out = nullptr;  // pinned: count + keypoints + descriptors
    // The per-image work is a fixed sequence of ~51 small launches (640 x 480: launch latency, not the GPU, set the 0.51 ms): it is
    // captured ONCE per geometry into a CUDA graph.  Only the resize chain is serial; each level's FAST -> NMS -> pick chain and its
    // blur chain fork off onto side streams as soon as the level's image exists and join before the final selection, so the graph
    // has up to 17 concurrent branches.  SFMM_ORB_NO_GRAPH=1 issues the same fork/join sequence eagerly.
    cudaStream_t side[2 * N_LEVELS] = {};
    cudaEvent_t ev_ext[N_LEVELS] = {}, ev_side[2 * N_LEVELS] = {};
    cudaGraphExec_t graph_exec = nullptr;
    const void* graph_key[6] = {};  // what the captured graph has baked in: geometry, channels, buffers
    bool use_graph = true;
    int64_t launches = 0;
    double last_ms = 0;
    mutable std::string err;
};

namespace {

int ofail(const SfmmOrb* o, int code, const std::string& msg) {
    if (o) o->err = msg;
    else g_orb_create_error = msg;
    return code;
}
#define ORB_TRY(o, expr)                                                                                                    \
    do {                                                                                                                    \
        cudaError_t _e = (expr);                                                                                            \
        if (_e != cudaSuccess) {                                                                                            \
            (void)cudaGetLastError();                                                                                       \
            return ofail(o, _e == cudaErrorMemoryAllocation ? SFMM_ENOMEM : SFMM_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
        }                                                                                                                   \
    } while (0)

// INTER_LINEAR_EXACT taps of one axis (resize.cpp: interpolationLinear): source offset and weight of the second tap in 8.8
void linear_taps(int src, int dst, std::vector<int>& ofs, std::vector<short>& c1) {
    const double scale = 1.0 / (static_cast<double>(dst) / static_cast<double>(src));
    ofs.assign(dst, 0);
    c1.assign(dst, 0);
    for (int d = 0; d < dst; ++d) {
        const double f = scale * (static_cast<double>(d) + 0.5) - 0.5;
        const int i = static_cast<int>(std::floor(f));
        if (i >= 0 && src > 1) {
            if (i < src - 1) {
                ofs[d] = i;
                c1[d] = static_cast<short>(std::nearbyint((f - i) * 256.0));
            } else {
                ofs[d] = src - 1;
            }
        }
    }
}

// Lays the level buffers out for a rows x cols image (once per geometry) and uploads the resize taps.
int layout(SfmmOrb* o, int rows, int cols) {
    if (o->rows == rows && o->cols == cols) return SFMM_OK;
    // level scales, sizes and quotas exactly as ORB_Impl computes them (fp32 / double mix included)
    const float scale_factor = 1.2f;
    const float factor = 1.f / scale_factor;
    float ndesired = N_FEATURES * (1.f - factor) / (1.f - static_cast<float>(std::pow(static_cast<double>(factor), static_cast<double>(N_LEVELS))));
    int sum = 0;
    size_t total = 0;
    auto take = [&](size_t bytes) {
        const size_t at = total;
        total += (bytes + 255) & ~size_t(255);
        return at;
    };
    size_t off[N_LEVELS][12];
    for (int l = 0; l < N_LEVELS; ++l) {
        Level& L = o->lv[l];
        L.scale = static_cast<float>(std::pow(static_cast<double>(scale_factor), static_cast<double>(l)));
        L.w = static_cast<int>(std::lrintf(cols / L.scale));
        L.h = static_cast<int>(std::lrintf(rows / L.scale));
        L.pitch = L.w + 2 * BORDER;
        if (l < N_LEVELS - 1) {
            L.quota = static_cast<int>(std::lrintf(ndesired));
            sum += L.quota;
            ndesired *= factor;
        } else {
            L.quota = std::max(N_FEATURES - sum, 0);
        }
        L.cand_cap = std::max(1, L.w * L.h / 4 + 16);
        const size_t ext = static_cast<size_t>(L.h + 2 * BORDER) * L.pitch;
        off[l][0] = take(ext);
        off[l][1] = take(ext);
        off[l][2] = take(static_cast<size_t>(L.w) * L.h);
        off[l][3] = take(static_cast<size_t>(L.h + 6) * L.w * sizeof(float));
        off[l][4] = take(static_cast<size_t>(L.w) * sizeof(int));
        off[l][5] = take(static_cast<size_t>(L.h) * sizeof(int));
        off[l][6] = take(static_cast<size_t>(L.w) * sizeof(short));
        off[l][7] = take(static_cast<size_t>(L.h) * sizeof(short));
        off[l][8] = take(static_cast<size_t>(L.cand_cap) * sizeof(uint32_t));
        off[l][9] = take(static_cast<size_t>(L.cand_cap));
        off[l][10] = take(static_cast<size_t>(L.cand_cap) * sizeof(uint32_t));
        off[l][11] = take(static_cast<size_t>(L.cand_cap) * sizeof(float));
    }
    ORB_TRY(o, o->pool.ensure(total));
    unsigned char* base = static_cast<unsigned char*>(o->pool.p);
    std::vector<LevelDev> dev(N_LEVELS);
    for (int l = 0; l < N_LEVELS; ++l) {
        Level& L = o->lv[l];
        L.ext = base + off[l][0];
        L.blur = base + off[l][1];
        L.score = base + off[l][2];
        L.rows = reinterpret_cast<float*>(base + off[l][3]);
        L.ofsx = reinterpret_cast<int*>(base + off[l][4]);
        L.ofsy = reinterpret_cast<int*>(base + off[l][5]);
        L.cx1 = reinterpret_cast<short*>(base + off[l][6]);
        L.cy1 = reinterpret_cast<short*>(base + off[l][7]);
        L.cand = reinterpret_cast<uint32_t*>(base + off[l][8]);
        L.cand_score = base + off[l][9];
        L.pick = reinterpret_cast<uint32_t*>(base + off[l][10]);
        L.pick_resp = reinterpret_cast<float*>(base + off[l][11]);
        if (l > 0) {
            std::vector<int> ofs;
            std::vector<short> c1;
            linear_taps(o->lv[l - 1].w, L.w, ofs, c1);
            ORB_TRY(o, cudaMemcpyAsync(L.ofsx, ofs.data(), ofs.size() * sizeof(int), cudaMemcpyHostToDevice, o->st));
            ORB_TRY(o, cudaMemcpyAsync(L.cx1, c1.data(), c1.size() * sizeof(short), cudaMemcpyHostToDevice, o->st));
            ORB_TRY(o, cudaStreamSynchronize(o->st));
            linear_taps(o->lv[l - 1].h, L.h, ofs, c1);
            ORB_TRY(o, cudaMemcpyAsync(L.ofsy, ofs.data(), ofs.size() * sizeof(int), cudaMemcpyHostToDevice, o->st));
            ORB_TRY(o, cudaMemcpyAsync(L.cy1, c1.data(), c1.size() * sizeof(short), cudaMemcpyHostToDevice, o->st));
            ORB_TRY(o, cudaStreamSynchronize(o->st));
        }
        dev[l] = LevelDev{L.w, L.h, L.pitch, L.quota, L.scale, L.ext, L.blur, L.pick, L.pick_resp, L.cand_score};
    }
    ORB_TRY(o, o->d_lv.ensure(sizeof(LevelDev) * N_LEVELS));
    ORB_TRY(o, cudaMemcpyAsync(o->d_lv.p, dev.data(), sizeof(LevelDev) * N_LEVELS, cudaMemcpyHostToDevice, o->st));
    ORB_TRY(o, cudaStreamSynchronize(o->st));
    o->rows = rows;
    o->cols = cols;
    return SFMM_OK;
}

}  // namespace

extern "C" {

SFMM_API const char* sfmm_orb_last_error(const SfmmOrb* o) { return o ? o->err.c_str() : g_orb_create_error.c_str(); }

SFMM_API void sfmm_orb_destroy(SfmmOrb* o) {
    if (!o) return;
    cudaSetDevice(o->device);
    if (o->st) cudaStreamSynchronize(o->st);
    if (o->graph_exec) cudaGraphExecDestroy(o->graph_exec);
    for (cudaStream_t s : o->side) if (s) cudaStreamDestroy(s);
    for (cudaEvent_t e : o->ev_ext) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : o->ev_side) if (e) cudaEventDestroy(e);
    for (Buf* b : {&o->pool, &o->d_src, &o->d_counts, &o->d_lv, &o->d_keys, &o->d_kps, &o->d_desc}) b->release();
    if (o->h_pin) cudaFreeHost(o->h_pin);
    if (o->h_out) cudaFreeHost(o->h_out);
    if (o->ev0) cudaEventDestroy(o->ev0);
    if (o->ev1) cudaEventDestroy(o->ev1);
    if (o->st) cudaStreamDestroy(o->st);
    delete o;
}

SFMM_API int sfmm_orb_create(int32_t device, SfmmOrb** out) {
    if (!out) return ofail(nullptr, SFMM_EINVAL, "orb_create: NULL argument");
    *out = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        (void)cudaGetLastError();
        return ofail(nullptr, SFMM_ENODEVICE, std::string("orb_create: no CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback");
    }
    if (device < 0 || device >= n_dev) return ofail(nullptr, SFMM_ERANGE, "orb_create: device ordinal out of range");
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10)
        return ofail(nullptr, SFMM_ENODEVICE, "orb_create: kernels are built for sm_100a (Blackwell B200) only");
    SfmmOrb* o = new (std::nothrow) SfmmOrb();
    if (!o) return ofail(nullptr, SFMM_ENOMEM, "orb_create: out of host memory");
    o->device = device;
    bool ok = cudaSetDevice(device) == cudaSuccess && cudaStreamCreateWithFlags(&o->st, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreate(&o->ev0) == cudaSuccess && cudaEventCreate(&o->ev1) == cudaSuccess;
    for (int i = 0; ok && i < 2 * N_LEVELS; ++i)
        ok = cudaStreamCreateWithFlags(&o->side[i], cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&o->ev_side[i], cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; ok && i < N_LEVELS; ++i) ok = cudaEventCreateWithFlags(&o->ev_ext[i], cudaEventDisableTiming) == cudaSuccess;
    if (const char* e2 = std::getenv("SFMM_ORB_NO_GRAPH")) o->use_graph = std::atoi(e2) == 0;
    // constants: the learned pattern, the disc's half-widths (orb.cpp: umax), cv::getGaussianKernel(7, 2, CV_32F)
    int umax[HALF_PATCH + 2] = {0};
    {
        const int vmax = static_cast<int>(std::floor(HALF_PATCH * std::sqrt(2.f) / 2 + 1)), vmin = static_cast<int>(std::ceil(HALF_PATCH * std::sqrt(2.f) / 2));
        for (int v = 0; v <= vmax; ++v) umax[v] = static_cast<int>(std::lrint(std::sqrt(static_cast<double>(HALF_PATCH) * HALF_PATCH - v * v)));
        for (int v = HALF_PATCH, v0 = 0; v >= vmin; --v) {
            while (umax[v0] == umax[v0 + 1]) ++v0;
            umax[v] = v0;
            ++v0;
        }
    }
    float gauss[7];
    {
        double k[7], s = 0;
        for (int i = 0; i < 7; ++i) {
            const double x = i - 3.0;
            k[i] = std::exp(-(x * x) / (2.0 * 2.0 * 2.0));
            s += k[i];
        }
        for (int i = 0; i < 7; ++i) gauss[i] = static_cast<float>(k[i] / s);
    }
    ok = ok && cudaMemcpyToSymbol(c_pattern, ORB_PATTERN_31, sizeof(ORB_PATTERN_31)) == cudaSuccess &&
         cudaMemcpyToSymbol(c_umax, umax, sizeof(umax)) == cudaSuccess && cudaMemcpyToSymbol(c_gauss, gauss, sizeof(gauss)) == cudaSuccess;
    ok = ok && o->d_counts.ensure((2 * N_LEVELS + 1 + N_LEVELS * 256) * sizeof(int)) == cudaSuccess &&
         o->d_keys.ensure(sizeof(KeyInfo) * MAX_OUT) == cudaSuccess && o->d_kps.ensure(sizeof(SfmKeyPoint) * MAX_OUT) == cudaSuccess &&
         o->d_desc.ensure(32 * MAX_OUT) == cudaSuccess &&
         cudaMallocHost(&o->h_out, 64 + (sizeof(SfmKeyPoint) + 32) * MAX_OUT) == cudaSuccess;
    if (!ok) {
        const std::string msg = std::string("orb_create: ") + cudaGetErrorString(cudaGetLastError());
        sfmm_orb_destroy(o);
        return ofail(nullptr, SFMM_ECUDA, msg);
    }
    *out = o;
    return SFMM_OK;
}

SFMM_API int sfmm_orb_stats(const SfmmOrb* o, int64_t* kernel_launches, double* last_ms) {
    if (!o || !kernel_launches || !last_ms) return SFMM_EINVAL;
    *kernel_launches = o->launches;
    *last_ms = o->last_ms;
    return SFMM_OK;
}

SFMM_API int sfmm_orb_detect_and_compute(SfmmOrb* o, const uint8_t* image, int32_t rows, int32_t cols, size_t step_bytes, int32_t channels,
                                         SfmKeyPoint* keypoints, uint8_t* descriptors, int32_t capacity, int32_t* count) {
    if (!o) return SFMM_EINVAL;
    if (!image || !count || rows <= 0 || cols <= 0 || (channels != 1 && channels != 3) || capacity < 0 || (capacity > 0 && (!keypoints || !descriptors)))
        return ofail(o, SFMM_EINVAL, "orb_detect_and_compute: bad argument");
    if (step_bytes < static_cast<size_t>(cols) * channels) return ofail(o, SFMM_EINVAL, "orb_detect_and_compute: row step smaller than a row");
    if (rows > 16384 || cols > 16384) return ofail(o, SFMM_ERANGE, "orb_detect_and_compute: images up to 16384 x 16384");
    *count = 0;
    ORB_TRY(o, cudaSetDevice(o->device));
    try {
        int rc = layout(o, rows, cols);
        if (rc) return rc;
    } catch (const std::exception& e) {
        return ofail(o, SFMM_ENOMEM, std::string("orb_detect_and_compute: ") + e.what());
    }
    if (o->lv[N_LEVELS - 1].w <= 2 * EDGE || o->lv[N_LEVELS - 1].h <= 2 * EDGE) {
        // (cv::ORB simply finds nothing on levels smaller than the border; supported here down to the last level being larger)
    }
    cudaStream_t st = o->st;
    // image -> pinned staging -> device
    const size_t row_bytes = static_cast<size_t>(cols) * channels, bytes = row_bytes * rows;
    if (bytes > o->h_pin_cap) {
        if (o->h_pin) cudaFreeHost(o->h_pin);
        o->h_pin = nullptr;
        o->h_pin_cap = 0;
        ORB_TRY(o, cudaMallocHost(&o->h_pin, bytes));
        o->h_pin_cap = bytes;
    }
    for (int y = 0; y < rows; ++y) std::memcpy(static_cast<unsigned char*>(o->h_pin) + y * row_bytes, image + static_cast<size_t>(y) * step_bytes, row_bytes);
    ORB_TRY(o, o->d_src.ensure(bytes));
    int* h_count = static_cast<int*>(o->h_out);
    unsigned char* h = static_cast<unsigned char*>(o->h_out) + 64;
    // the whole image: upload, pyramid, detection, description, results (count + MAX_OUT records: 245 KB, the first `count` are valid)
    auto enqueue = [&]() -> cudaError_t {
        cudaError_t e;
#define ORB_Q(expr) do { if ((e = (expr)) != cudaSuccess) return e; } while (0)
        ORB_Q(cudaMemcpyAsync(o->d_src.p, o->h_pin, bytes, cudaMemcpyHostToDevice, st));
        int* d_n_cand = static_cast<int*>(o->d_counts.p);
        int* d_n_pick = d_n_cand + N_LEVELS;
        int* d_n_out = d_n_pick + N_LEVELS;
        int* d_hist = d_n_out + 1;
        ORB_Q(cudaMemsetAsync(o->d_counts.p, 0, (2 * N_LEVELS + 1 + N_LEVELS * 256) * sizeof(int), st));
        const float harris_k = 0.04f;
        const float hs = 1.f / ((1 << 2) * HARRIS_BLOCK * 255.f);
        const float scale_sq_sq = hs * hs * hs * hs;
        const dim3 blk(32, 8);
        for (int l = 0; l < N_LEVELS; ++l) {
            Level& L = o->lv[l];
            const dim3 ge((L.pitch + 31) / 32, (L.h + 2 * BORDER + 7) / 8), gi((L.w + 31) / 32, (L.h + 7) / 8);
            if (l == 0) {
                orb_level0_kernel<<<ge, blk, 0, st>>>(static_cast<const unsigned char*>(o->d_src.p), row_bytes, channels, L.w, L.h, L.ext, L.pitch);
            } else {
                Level& P = o->lv[l - 1];
                orb_resize_kernel<<<ge, blk, 0, st>>>(P.ext + static_cast<size_t>(BORDER) * P.pitch + BORDER, P.pitch, P.w, P.h, L.ofsx, L.cx1, L.ofsy, L.cy1, L.w, L.h,
                                                      L.ext, L.pitch);
            }
            // fork: the level's image exists
            cudaStream_t sa = o->side[2 * l], sb = o->side[2 * l + 1];
            ORB_Q(cudaEventRecord(o->ev_ext[l], st));
            ORB_Q(cudaStreamWaitEvent(sa, o->ev_ext[l], 0));
            ORB_Q(cudaStreamWaitEvent(sb, o->ev_ext[l], 0));
            const unsigned char* interior = L.ext + static_cast<size_t>(BORDER) * L.pitch + BORDER;
            orb_fast_score_kernel<<<gi, blk, 0, sa>>>(interior, L.pitch, L.w, L.h, L.score);
            orb_nms_kernel<<<gi, blk, 0, sa>>>(L.score, L.w, L.h, L.cand, L.cand_score, L.cand_cap, d_n_cand + l, d_hist + l * 256);
            orb_pick_fast_kernel<<<1, 256, 0, sa>>>(interior, L.pitch, L.cand, L.cand_score, d_n_cand + l, L.cand_cap, d_hist + l * 256, 2 * L.quota, L.pick,
                                                    L.pick_resp, d_n_pick + l, harris_k, scale_sq_sq);
            // the blurred copy: frame = the unblurred frame (ORB blurs the level in place inside its framed buffer), interior = blur
            ORB_Q(cudaMemcpyAsync(L.blur, L.ext, static_cast<size_t>(L.h + 2 * BORDER) * L.pitch, cudaMemcpyDeviceToDevice, sb));
            orb_blur_rows_kernel<<<dim3((L.w + 31) / 32, (L.h + 6 + 7) / 8), blk, 0, sb>>>(L.ext, L.pitch, L.w, L.h, L.rows);
            orb_blur_cols_kernel<<<gi, blk, 0, sb>>>(L.rows, L.w, L.h, L.blur, L.pitch);
            ORB_Q(cudaEventRecord(o->ev_side[2 * l], sa));
            ORB_Q(cudaEventRecord(o->ev_side[2 * l + 1], sb));
        }
        for (int i = 0; i < 2 * N_LEVELS; ++i) ORB_Q(cudaStreamWaitEvent(st, o->ev_side[i], 0));  // join
        orb_pick_harris_kernel<<<1, 1024, 0, st>>>(static_cast<const LevelDev*>(o->d_lv.p), d_n_pick, static_cast<KeyInfo*>(o->d_keys.p), d_n_out, MAX_OUT);
        orb_angle_kernel<<<(MAX_OUT + 7) / 8, 256, 0, st>>>(static_cast<const LevelDev*>(o->d_lv.p), static_cast<const KeyInfo*>(o->d_keys.p), d_n_out, MAX_OUT,
                                                           static_cast<SfmKeyPoint*>(o->d_kps.p));
        orb_describe_kernel<<<(MAX_OUT + 7) / 8, 256, 0, st>>>(static_cast<const LevelDev*>(o->d_lv.p), static_cast<const KeyInfo*>(o->d_keys.p),
                                                              static_cast<const SfmKeyPoint*>(o->d_kps.p), d_n_out, MAX_OUT,
                                                              static_cast<unsigned char*>(o->d_desc.p));
        ORB_Q(cudaGetLastError());
        ORB_Q(cudaMemcpyAsync(h_count, d_n_out, sizeof(int), cudaMemcpyDeviceToHost, st));
        ORB_Q(cudaMemcpyAsync(h, o->d_kps.p, sizeof(SfmKeyPoint) * MAX_OUT, cudaMemcpyDeviceToHost, st));
        ORB_Q(cudaMemcpyAsync(h + sizeof(SfmKeyPoint) * MAX_OUT, o->d_desc.p, static_cast<size_t>(32) * MAX_OUT, cudaMemcpyDeviceToHost, st));
#undef ORB_Q
        return cudaSuccess;
    };
    if (o->use_graph) {
        const void* key[6] = {reinterpret_cast<const void*>(static_cast<uintptr_t>(rows)), reinterpret_cast<const void*>(static_cast<uintptr_t>(cols)),
                              reinterpret_cast<const void*>(static_cast<uintptr_t>(channels)), o->pool.p, o->d_src.p, o->h_pin};
        if (!o->graph_exec || std::memcmp(key, o->graph_key, sizeof(key)) != 0) {
            if (o->graph_exec) {
                cudaGraphExecDestroy(o->graph_exec);
                o->graph_exec = nullptr;
            }
            cudaGraph_t graph = nullptr;
            ORB_TRY(o, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            const cudaError_t qe = enqueue();
            const cudaError_t ce = cudaStreamEndCapture(st, &graph);  // always: never leave the streams capturing
            if (qe != cudaSuccess || ce != cudaSuccess) {
                if (graph) cudaGraphDestroy(graph);
                (void)cudaGetLastError();
                return ofail(o, SFMM_ECUDA, std::string("orb_detect_and_compute: graph capture: ") + cudaGetErrorString(qe != cudaSuccess ? qe : ce));
            }
            const cudaError_t ie = cudaGraphInstantiate(&o->graph_exec, graph, 0);
            cudaGraphDestroy(graph);
            ORB_TRY(o, ie);
            std::memcpy(o->graph_key, key, sizeof(key));
        }
        ORB_TRY(o, cudaEventRecord(o->ev0, st));
        ORB_TRY(o, cudaGraphLaunch(o->graph_exec, st));
    } else {
        ORB_TRY(o, cudaEventRecord(o->ev0, st));
        ORB_TRY(o, enqueue());
    }
    ORB_TRY(o, cudaEventRecord(o->ev1, st));
    ORB_TRY(o, cudaStreamSynchronize(st));
    o->launches += 6 * N_LEVELS + 3;
    const int n = *h_count;
    if (n > MAX_OUT) return ofail(o, SFMM_ERANGE, "orb_detect_and_compute: more than 4096 keypoints (ties)");
    *count = n;
    if (n > capacity) return ofail(o, SFMM_ERANGE, "orb_detect_and_compute: output capacity too small (count holds the size needed)");
    if (n) {
        std::memcpy(keypoints, h, sizeof(SfmKeyPoint) * n);
        std::memcpy(descriptors, h + sizeof(SfmKeyPoint) * MAX_OUT, static_cast<size_t>(32) * n);
    }
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, o->ev0, o->ev1) == cudaSuccess) o->last_ms = ms;
    return SFMM_OK;
}

}  // extern "C"
