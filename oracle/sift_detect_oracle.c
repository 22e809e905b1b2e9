/*
 * sift_detect_oracle.c -- CPU restatement of SIFT key-point detection and description:
 * SiftGPU::RunSIFT + GetKeyPointsAndDescriptorsCUDA as BundleFusion configures them (FL/Bundler.cpp:55-100).
 *   parameters        SiftParam::ParseSiftParam / GetInitialSmoothSigma / GetLevelSigma      FL/SiftGPU/SiftGPU.cpp:105-174, 256-259
 *                     SiftGPU::SetParams (4 octaves from octave 0, no sub-pixel step)          FL/SiftGPU/SiftGPU.cpp:224-254
 *   filter kernels    ProgramCU::CreateFilterKernel                                            FL/SiftGPU/ProgramCU.cu:431-462
 *   pyramid           SiftPyramid::BuildPyramid; FilterH / FilterV; DownsampleKernel           FL/SiftGPU/SiftPyramid.cpp:82-145, ProgramCU.cu:159-263, 330-370
 *   DoG + gradient    SiftPyramid::DetectKeypoints; ComputeDOG_Kernel                          SiftPyramid.cpp:351-394, ProgramCU.cu:550-598
 *   extrema           ComputeKEY_Kernel (depth gate, 26-neighbour test, edge test)             ProgramCU.cu:601-757
 *   count limit       SiftPyramid::LimitFeatureCount                                           SiftPyramid.cpp:227-255
 *   orientation       SiftPyramid::GetFeatureOrientations; ComputeOrientation_Kernel           SiftPyramid.cpp:426-450, ProgramCU.cu:905-1143
 *   list reshape      SiftPyramid::ReshapeFeatureList; ReshapeFeatureList_Kernel               SiftPyramid.cpp:297-314, ProgramCU.cu:1994-2047
 *   descriptor        SiftPyramid::GetFeatureDescriptors; ComputeDescriptor_Kernel, NormalizeDescriptor_Kernel   SiftPyramid.cpp:275-295, ProgramCU.cu:1178-1258, 1339-1371
 *   outputs           CreateGlobalKeyPointList(_Kernel), ConvertDescriptorToUChar_Kernel       SiftPyramid.cpp:730-781, ProgramCU.cu:2049-2121
 * (FL/ = FriedLiver/Source/.)
 *
 * TEST INFRASTRUCTURE ONLY (see tsdf_oracle.c header).  PARITY STATUS: PINNED against the reference's own kernels and host classes
 * executed on the CPU.  ProgramCU.cu is written against texture references (removed from CUDA 12: nvcc cannot rebuild it), so
 * oracle/build_ref.py compiles it -- with SiftPyramid.cpp, SiftGPU.cpp, CuTexImage.cpp, GlobalUtil.cpp -- by g++ against a CPU emulation of
 * CUDA (oracle/ref_emu, tests/cuda_emu) into oracle/_ref/libref_sift_emulated.so; its outputs on seeded images are committed as
 * tests/golden/sift_reference_emulated.npz and this file reproduces them (tests/test_sift_reference_emulated.py): the same key points
 * bit for bit (position, scale, depth; four images incl. a half-resolution depth map with holes, the minimum-scale rule, the
 * feature-count limit, and the application's 640x480 frame with its parameters), descriptors identical for > 90 % of the features and within 2 counts for the rest (the reference's histogram sums
 * run in thread order), a second orientation decided differently for about one feature in 300 (a peak sitting at the 0.8 threshold).  Also checked by
 * known answers and invariances (tests/test_sift_detect_oracle.py).  Not yet compared with a run of the reference on a GPU.
 *
 * Contract where the reference is race-dependent or uses approximate hardware instructions (what a CUDA implementation is compared with):
 *   - key points of a level are listed in raster order (row-major), an orientation pair in (first, second) order; the reference appends
 *     through atomicAdd (order, and WHICH features survive a full list, depend on scheduling);
 *   - orientation votes and descriptor bins are accumulated in raster order over the sampling window (reference: shared-memory atomics);
 *   - multiply-adds the reference's nvcc build fuses are written as fmaf (filter taps, squared distances); -ffp-contract=off elsewhere;
 *   - __fdividef, __sincosf and rsqrt are taken as exact division, sinf / cosf and 1 / sqrtf;
 *   - a descriptor sample whose orientation difference rounds to exactly 8.0 bins goes to bin 0 / 1 (the reference indexes des[8]).
 * Images whose four octave widths are not multiples of 4 are rejected (the reference pads rows to a multiple of 4 and filters the padding).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

#define DOG_LEVELS 3                 /* _dog_level_num */
#define LEVELS 6                     /* _level_num = _level_max - _level_min + 1 = 4 - (-1) + 1 */
#define OCTAVES 4                    /* GlobalUtil::_octave_num_default, SiftGPU.cpp:245 */
#define KERNEL_MAX_WIDTH 33
#define KERNEL_MIN_WIDTH 5
#define MAX_LEVEL_FEATURES 4096      /* GlobalUtil::_MaxLevelFeatureNum */

typedef struct {
    uint32_t width, height;              /* SIFT (intensity) image */
    uint32_t depthWidth, depthHeight;    /* c_siftCameraParams.m_depthWidth / Height */
    float depthMin, depthMax;            /* GlobalUtil::_SiftDepthMin / Max */
    float minKeyScale;                   /* c_siftCameraParams.m_minKeyScale */
    int32_t featureCountThreshold;       /* GlobalUtil::_FeatureCountThreshold (150 in Bundler.cpp:61) */
    uint32_t maxKeyPoints;               /* capacity of the outputs (s_maxNumKeysPerImage) */
} SiftDetectParams;

typedef struct { int w, h; float* gus[LEVELS]; float* dog[LEVELS]; float* grd[LEVELS]; float* rot[LEVELS]; } Octave;
typedef struct { float x, y, s, o; } Feature;

/* ---- parameters (SiftGPU.cpp:127-174) ---- */
typedef struct { float sigma0, sigmas[LEVELS]; float kernel[LEVELS][KERNEL_MAX_WIDTH]; int fw[LEVELS]; float dogThreshold, edgeThreshold; } SiftParam;

static void create_filter_kernel(float sigma, float* kernel, int* width) {        /* ProgramCU.cu:431-462 */
    int sz = (int)ceil(4.0f * sigma - 0.5);
    *width = 2 * sz + 1;
    if (*width > KERNEL_MAX_WIDTH) { sz = KERNEL_MAX_WIDTH >> 1; *width = KERNEL_MAX_WIDTH; }
    else if (*width < KERNEL_MIN_WIDTH) { sz = KERNEL_MIN_WIDTH >> 1; *width = KERNEL_MIN_WIDTH; }
    float rv = 1.0f / (sigma * sigma), ksum = 0.0f;
    for (int i = -sz; i <= sz; ++i) { const float v = expf(-0.5f * i * i * rv); kernel[i + sz] = v; ksum += v; }
    rv = 1.0f / ksum;
    for (int i = 0; i < *width; ++i) kernel[i] *= rv;
}

static void parse_param(SiftParam* p) {
    const int levelMin = -1;
    p->sigma0 = 1.6f * powf(2.0f, 1.0f / DOG_LEVELS);
    const float sigman = 0.5f;
    const float sigmak = powf(2.0f, 1.0f / DOG_LEVELS);
    const float dsigma0 = p->sigma0 * sqrtf(1.0f - 1.0f / (sigmak * sigmak));
    /* GetInitialSmoothSigma(octave_min = 0) */
    const float sa = p->sigma0 * powf(2.0f, (float)levelMin / (float)DOG_LEVELS), sb = sigman / powf(2.0f, 0.0f);
    p->sigmas[0] = sa > sb + 0.001 ? sqrtf(sa * sa - sb * sb) : 0.0f;
    for (int i = 0; i <= 4; ++i) p->sigmas[i + 1] = dsigma0 * powf(sigmak, (float)i);        /* _sigma[i], i = _level_min + 1 .. _level_max */
    for (int i = 0; i < LEVELS; ++i) { memset(p->kernel[i], 0, sizeof p->kernel[i]); create_filter_kernel(p->sigmas[i], p->kernel[i], &p->fw[i]); }
    p->dogThreshold = 0.02f / DOG_LEVELS;
    p->edgeThreshold = 10.0f;
}

/* ---- FilterH + FilterV (ProgramCU.cu:159-263): clamp to edge, taps accumulated left to right / top to bottom with fused multiply-add ---- */
static void filter_image(float* dst, const float* src, float* buf, int w, int h, const float* k, int fw) {
    const int half = fw >> 1;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            float v = 0.0f;
            for (int i = 0; i < fw; ++i) { int xx = x - half + i; xx = xx < 0 ? 0 : (xx > w - 1 ? w - 1 : xx); v = fmaf(src[y * w + xx], k[i], v); }
            buf[y * w + x] = v;
        }
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            float v = 0.0f;
            for (int i = 0; i < fw; ++i) { int yy = y - half + i; yy = yy < 0 ? 0 : (yy > h - 1 ? h - 1 : yy); v = fmaf(buf[yy * w + x], k[i], v); }
            dst[y * w + x] = v;
        }
}

/* ---- ComputeKEY_Kernel (ProgramCU.cu:601-757), no sub-pixel step; returns +1 / -1 for a maximum / minimum, 0 otherwise ---- */
static int cmp_rows(const float* img, int idx, float v, float* nmax, float* nmin) {      /* READ_CMP_DOG_DATA: 1 = rejected */
    const float d0 = img[idx - 1], d1 = img[idx], d2 = img[idx + 1];
    if (v > *nmax) { *nmax = fmaxf(*nmax, d0); *nmax = fmaxf(*nmax, d1); *nmax = fmaxf(*nmax, d2); if (v < *nmax) return 1; }
    else { *nmin = fminf(*nmin, d0); *nmin = fminf(*nmin, d1); *nmin = fminf(*nmin, d2); if (v > *nmin) return 1; }
    return 0;
}
static int key_test(const float* dogP, const float* dogC, const float* dogN, int w, int row, int col, float dogThreshold, float edgeT) {
    const int index = row * w + col, up = index - w, dn = index + w;
    const float v = dogC[index];
    if (fabsf(v) <= dogThreshold) return 0;
    float nmax = fmaxf(dogC[index - 1], dogC[index + 1]), nmin = fminf(dogC[index - 1], dogC[index + 1]);
    if (v <= nmax && v >= nmin) return 0;
    if (cmp_rows(dogC, up, v, &nmax, &nmin)) return 0;
    if (cmp_rows(dogC, dn, v, &nmax, &nmin)) return 0;
    const float vx2 = v * 2.0f;
    const float fxx = dogC[index - 1] + dogC[index + 1] - vx2;
    const float fyy = dogC[up] + dogC[dn] - vx2;
    const float fxy = 0.25f * (dogC[dn + 1] + dogC[up - 1] - dogC[dn - 1] - dogC[up + 1]);
    const float t1 = fxx * fyy - fxy * fxy, t2 = (fxx + fyy) * (fxx + fyy);
    if (t1 <= 0.0f || t2 > edgeT * t1) return 0;
    if (cmp_rows(dogP, up, v, &nmax, &nmin) || cmp_rows(dogP, index, v, &nmax, &nmin) || cmp_rows(dogP, dn, v, &nmax, &nmin)) return 0;
    if (cmp_rows(dogN, up, v, &nmax, &nmin) || cmp_rows(dogN, index, v, &nmax, &nmin) || cmp_rows(dogN, dn, v, &nmax, &nmin)) return 0;
    return v > nmax ? 1 : -1;
}

/* (unsigned int) of the GPU (cvt.rzi.u32.f32): negative and NaN -> 0, saturating.  A sampling window that lies outside the image has a
 * negative extent; the GPU turns it into an empty loop, plain C would make it undefined. */
static unsigned f2u_gpu(float x) { if (!(x > 0.0f)) return 0u; if (x >= 4294967296.0f) return 0xFFFFFFFFu; return (unsigned)x; }

/* ---- ComputeOrientation_Kernel (ProgramCU.cu:905-1143): two packed 16-bit orientations ---- */
static uint32_t orientation(const float* grd, const float* rot, int w, int h, float kx, float ky, float sigma) {
    const float tenDegPerRad = 5.7295779513082320876798154814105f;
    const float gsigma = sigma * 1.5f, win = fabsf(sigma) * 1.5f * 2.0f;
    const float distThreshold = (float)(win * win + 0.5);
    const float factor = -0.5f / (gsigma * gsigma);
    const float xmin = fmaxf(1.5f, floorf(kx - win) + 0.5f), ymin = fmaxf(1.5f, floorf(ky - win) + 0.5f);
    const float xmax = fminf(w - 1.5f, floorf(kx + win) + 0.5f), ymax = fminf(h - 1.5f, floorf(ky + win) + 0.5f);
    float vote[36], tmp[36];
    for (int i = 0; i < 36; ++i) vote[i] = 0.0f;
    const unsigned xlen = f2u_gpu(roundf(xmax - xmin + 1)), ylen = f2u_gpu(roundf(ymax - ymin + 1)), num = xlen * ylen;
    for (unsigned i = 0; i < num; ++i) {
        const float x = (float)(i % xlen) + xmin, y = (float)(i / xlen) + ymin;
        const float dx = x - kx, dy = y - ky;
        const float sq = fmaf(dx, dx, dy * dy);
        if (sq < distThreshold) {
            const int pix = (int)floorf(y) * w + (int)floorf(x);
            const float weight = grd[pix] * expf(sq * factor);
            int oidx = (int)floorf(rot[pix] * tenDegPerRad);
            if (oidx < 0) oidx += 36;
            vote[oidx] += weight;
        }
    }
    const float oneThird = (float)(1.0 / 3.0);
    float* src = vote; float* dst = tmp;
    for (int it = 0; it < 6; ++it) {
        for (int t = 0; t < 36; ++t) dst[t] = (src[(t + 35) % 36] + src[t] + src[(t + 1) % 36]) * oneThird;
        float* s = src; src = dst; dst = s;
    }                                                   /* six passes: the result is back in vote[] */
    float maxVote = 0.0f;
    for (int t = 0; t < 36; ++t) maxVote = fmaxf(maxVote, vote[t]);
    const float thr = maxVote * 0.8f;
    float maxRot[2] = { 0.0f, 0.0f }; int ocount = 0, maxIndex = -1;
    for (int pass = 0; pass < 2; ++pass) {
        float best = -1.0f; int arg = -1;
        for (int c = 0; c < 36; ++c) {
            if (pass == 1 && c == maxIndex) continue;
            const int m = (c + 35) % 36, p = (c + 1) % 36;
            if (vote[c] > thr && vote[c] > vote[m] && vote[c] > vote[p] && vote[c] > best) { best = vote[c]; arg = c; }    /* ties: lowest bin */
        }
        if (arg < 0) { if (pass == 0) break; else continue; }
        const int m = (arg + 35) % 36, p = (arg + 1) % 36;
        const float di = 0.5f * ((vote[p] - vote[m]) / (2.0f * vote[arg] - vote[p] - vote[m]));
        maxRot[pass] = (float)arg + di + 0.5f;
        ++ocount;
        if (pass == 0) maxIndex = arg;
    }
    float fr1 = maxRot[0] / 36.0f; if (fr1 < 0) fr1 += 1.0f;
    const uint32_t us1 = ocount == 0 ? 65535u : (uint32_t)(uint16_t)floorf(fr1 * 65535.0f);
    uint32_t us2 = 65535u;
    if (ocount > 1) { float fr2 = maxRot[1] / 36.0f; if (fr2 < 0) fr2 += 1.0f; us2 = (uint32_t)(uint16_t)floorf(fr2 * 65535.0f); }
    return (us2 << 16) | us1;
}

/* ---- ComputeDescriptor_Kernel + NormalizeDescriptor_Kernel (ProgramCU.cu:1178-1258, 1339-1371) ---- */
static void descriptor(const float* grd, const float* rot, int w, int h, Feature key, float* out /*[128]*/) {
    const float rpi = (float)(4.0 / 3.14159265358979323846);
    const float spt = fabsf(key.s * 3.0f);
    const float s = sinf(key.o), c = cosf(key.o);
    const float anglef = (double)key.o > 3.14159265358979323846 ? (float)(key.o - (2.0 * 3.14159265358979323846)) : key.o;
    const float cspt = c * spt, sspt = s * spt, crspt = c / spt, srspt = s / spt;
    for (int b = 0; b < 16; ++b) {
        const int ix = b & 3, iy = b >> 2;
        const float ox = ix - 1.5f, oy = iy - 1.5f;
        const float ptx = cspt * ox - sspt * oy + key.x, pty = cspt * oy + sspt * ox + key.y;
        const float bsz = fabsf(cspt) + fabsf(sspt);
        const float xmin = fmaxf(1.5f, floorf(ptx - bsz) + 0.5f), ymin = fmaxf(1.5f, floorf(pty - bsz) + 0.5f);
        const float xmax = fminf(w - 1.5f, floorf(ptx + bsz) + 0.5f), ymax = fminf(h - 1.5f, floorf(pty + bsz) + 0.5f);
        float des[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
        const unsigned xlen = f2u_gpu(roundf(xmax - xmin + 1)), ylen = f2u_gpu(roundf(ymax - ymin + 1)), size = xlen * ylen;
        for (unsigned i = 0; i < size; ++i) {
            const float x = (float)(i % xlen) + xmin, y = (float)(i / xlen) + ymin;
            const float dx = x - ptx, dy = y - pty;
            const float nx = crspt * dx + srspt * dy, ny = crspt * dy - srspt * dx;
            const float nxn = fabsf(nx), nyn = fabsf(ny);
            if (nxn < 1.0f && nyn < 1.0f) {
                const int pix = (int)floorf(y) * w + (int)floorf(x);
                const float dnx = nx + ox, dny = ny + oy;
                const float ww = expf(-0.125f * (dnx * dnx + dny * dny));
                const float wx = (float)(1.0 - nxn), wy = (float)(1.0 - nyn);
                const float weight = ww * wx * wy * grd[pix];
                float theta = (anglef - rot[pix]) * rpi;
                if (theta < 0) theta += 8.0f;
                const float fo = floorf(theta);
                const int fidx = (int)fo & 7;
                const float w1 = fo + 1.0f - theta, w2 = theta - fo;
                des[fidx] += w1 * weight;
                des[(fidx + 1) % 8] += w2 * weight;
            }
        }
        memcpy(out + 8 * b, des, sizeof des);
    }
    /* normalise, clamp at 0.2, normalise again; sums as the 32-lane butterfly takes them (4 values per lane, xor tree) */
    for (int pass = 0; pass < 2; ++pass) {
        float lane[32];
        for (int l = 0; l < 32; ++l) { const float* t = out + 4 * l; lane[l] = t[0] * t[0] + t[1] * t[1] + t[2] * t[2] + t[3] * t[3]; }
        for (int off = 16; off > 0; off /= 2) { float nxt[32]; for (int l = 0; l < 32; ++l) nxt[l] = lane[l] + lane[l ^ off]; memcpy(lane, nxt, sizeof lane); }
        const float inv = 1.0f / sqrtf(lane[0]);
        for (int k = 0; k < 128; ++k) out[k] = pass == 0 ? fminf(0.2f, out[k] * inv) : out[k] * inv;
    }
}

static void limit_feature_count(int* levelNum, int* featureNum, int threshold) {       /* SiftPyramid.cpp:227-255, _TruncateMethod 0 */
    if (threshold <= 0) return;
    int i = 0;
    while (i < OCTAVES * DOG_LEVELS && *featureNum - levelNum[i] > threshold) { *featureNum -= levelNum[i]; levelNum[i++] = 0; }
}

/* SiftGPU::RunSIFT + GetKeyPointsAndDescriptorsCUDA.  intensity: [height][width] float; depth: [depthHeight][depthWidth] float (-inf invalid).
 * keyPoints: [maxKeyPoints][4] = (x, y, scale, depth); descriptors: [maxKeyPoints][128] bytes; levelCounts (optional): [12] final count per
 * (octave, DoG level).  Returns the number of key points (<= maxKeyPoints), or -1 for an unsupported size. */
ORC_API int orc_sift_detect(const float* intensity, const float* depth, const SiftDetectParams* P, float* keyPoints, uint8_t* descriptors, int32_t* levelCounts) {
    if ((P->width & 31u) || P->width < 64 || P->height < 64 || (P->height & 7u)) return -1;
    SiftParam sp; parse_param(&sp);
    Octave oc[OCTAVES];
    int w = (int)P->width, h = (int)P->height;
    float* buf = (float*)malloc(sizeof(float) * (size_t)w * h);
    for (int o = 0; o < OCTAVES; ++o, w >>= 1, h >>= 1) {
        oc[o].w = w; oc[o].h = h;
        for (int l = 0; l < LEVELS; ++l) {
            oc[o].gus[l] = (float*)malloc(sizeof(float) * (size_t)w * h);
            oc[o].dog[l] = (float*)calloc((size_t)w * h, sizeof(float));
            oc[o].grd[l] = (float*)calloc((size_t)w * h, sizeof(float));
            oc[o].rot[l] = (float*)calloc((size_t)w * h, sizeof(float));
        }
        /* BuildPyramid, SiftPyramid.cpp:82-126 */
        if (o == 0) filter_image(oc[0].gus[0], intensity, buf, w, h, sp.kernel[0], sp.fw[0]);
        else {
            const float* src = oc[o - 1].gus[3];           /* GetBaseLevel(i - 1) + _level_ds - _level_min = level index 3 */
            const int sw = oc[o - 1].w;
            for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x) { int sx = x << 1; sx = sx < sw - 1 ? sx : sw - 1; oc[o].gus[0][y * w + x] = src[(y << 1) * sw + sx]; }
        }
        for (int l = 1; l < LEVELS; ++l) filter_image(oc[o].gus[l], oc[o].gus[l - 1], buf, w, h, sp.kernel[l], sp.fw[l]);
        /* ComputeDOG, ProgramCU.cu:550-598: DoG for levels 1..5, gradient where the reference allocates it (levels 1..3) */
        for (int l = 1; l < LEVELS; ++l) {
            const float* g = oc[o].gus[l]; const float* gp = oc[o].gus[l - 1];
            for (int i = 0; i < w * h; ++i) oc[o].dog[l][i] = g[i] - gp[i];
            if (l >= 1 && l < 1 + DOG_LEVELS)
                for (int y = 1; y < h - 1; ++y) for (int x = 1; x < w - 1; ++x) {      /* border gradients are never sampled (windows start at 1.5) */
                    const int i = y * w + x;
                    const float dx = g[i + 1] - g[i - 1], dy = g[i + w] - g[i - w];
                    const float gr = 0.5f * sqrtf(fmaf(dx, dx, dy * dy));
                    oc[o].grd[l][i] = gr; oc[o].rot[l][i] = gr == 0.0f ? 0.0f : atan2f(dy, dx);
                }
        }
    }
    free(buf);
    /* DetectKeypoints, SiftPyramid.cpp:351-394 */
    const float keyLocOffset = 0.5f;                          /* _LoweOrigin = 0 */
    const float edgeT = (sp.edgeThreshold + 1) * (sp.edgeThreshold + 1) / sp.edgeThreshold;
    int levelNum[OCTAVES * DOG_LEVELS], featureNum = 0, fmaxv[OCTAVES];
    int* rawX[OCTAVES * DOG_LEVELS]; int* rawY[OCTAVES * DOG_LEVELS];
    for (int o = 0; o < OCTAVES; ++o) {
        int fm = (int)(oc[o].w * oc[o].h * 0.005f);
        fm = fm > MAX_LEVEL_FEATURES ? MAX_LEVEL_FEATURES : (fm < 32 ? 32 : fm);
        fmaxv[o] = fm;
        const float keyLocScale = (float)(1 << o);
        for (int j = 0; j < DOG_LEVELS; ++j) {
            const int L = o * DOG_LEVELS + j, l = j + 2;              /* DoG level index 2..4 */
            rawX[L] = (int*)malloc(sizeof(int) * fm); rawY[L] = (int*)malloc(sizeof(int) * fm);
            int count = 0;
            for (int row = 1; row < oc[o].h - 2; ++row)              /* row > 0 && row < rowmax - 1, rowmax = height - 1 */
                for (int col = 1; col < oc[o].w - 2; ++col) {
                    const int dxp = (int)roundf((keyLocScale * (float)col + keyLocOffset) * (float)(P->depthWidth - 1) / (float)(P->width - 1));
                    const int dyp = (int)roundf((keyLocScale * (float)row + keyLocOffset) * (float)(P->depthHeight - 1) / (float)(P->height - 1));
                    if (dxp < 0 || dxp >= (int)P->depthWidth || dyp < 0 || dyp >= (int)P->depthHeight) continue;
                    const float d = depth[(size_t)dyp * P->depthWidth + dxp];
                    if (d == -INFINITY || d < P->depthMin || d > P->depthMax) continue;
                    if (!key_test(oc[o].dog[l - 1], oc[o].dog[l], oc[o].dog[l + 1], oc[o].w, row, col, sp.dogThreshold, edgeT)) continue;
                    if (count < fm) { rawX[L][count] = col; rawY[L][count] = row; }
                    ++count;
                }
            levelNum[L] = count < fm ? count : fm;
            featureNum += levelNum[L];
        }
    }
    limit_feature_count(levelNum, &featureNum, P->featureCountThreshold);
    /* GetFeatureOrientations + ReshapeFeatureList, SiftPyramid.cpp:426-450, 297-314 */
    const float factor = (float)(2.0 * 3.14159265358979323846 / 65535.0);
    Feature* fin[OCTAVES * DOG_LEVELS];
    featureNum = 0;
    for (int L = 0; L < OCTAVES * DOG_LEVELS; ++L) {
        const int o = L / DOG_LEVELS, j = L % DOG_LEVELS, l = j + 1;              /* gradient level index 1..3 */
        fin[L] = (Feature*)malloc(sizeof(Feature) * (size_t)fmaxv[o]);
        if (levelNum[L] == 0) continue;
        const float sigma = sp.sigma0 * powf(2.0f, (float)j / (float)DOG_LEVELS);   /* GetLevelSigma(j + _level_min + 1) */
        const float keyLocScale = (float)(1 << o);
        int n = 0;
        for (int k = 0; k < levelNum[L]; ++k) {
            const float kx = rawX[L][k] + 0.5f, ky = rawY[L][k] + 0.5f;
            const uint32_t pack = orientation(oc[o].grd[l], oc[o].rot[l], oc[o].w, oc[o].h, kx, ky, sigma);
            const uint32_t o0 = pack & 0xFFFFu, o1 = pack >> 16;
            if (!(sigma * keyLocScale >= P->minKeyScale)) continue;
            if (o0 == 65535u) continue;
            if (n < fmaxv[o]) { Feature f = { kx, ky, sigma, factor * (float)o0 }; fin[L][n] = f; }
            ++n;
            if (n - 1 < fmaxv[o] && o1 != 65535u && o1 != o0) {
                if (n < fmaxv[o]) { Feature f = { kx, ky, sigma, factor * (float)o1 }; fin[L][n] = f; }
                ++n;
            }
        }
        levelNum[L] = n < fmaxv[o] ? n : fmaxv[o];
        featureNum += levelNum[L];
    }
    limit_feature_count(levelNum, &featureNum, P->featureCountThreshold);
    /* GetFeatureDescriptors + outputs */
    int out = 0;
    for (int L = 0; L < OCTAVES * DOG_LEVELS; ++L) {
        const int o = L / DOG_LEVELS, l = L % DOG_LEVELS + 1;
        const float keyLocScale = (float)(1 << o);
        if (levelCounts) levelCounts[L] = levelNum[L];
        for (int k = 0; k < levelNum[L] && out < (int)P->maxKeyPoints; ++k, ++out) {
            float des[128];
            descriptor(oc[o].grd[l], oc[o].rot[l], oc[o].w, oc[o].h, fin[L][k], des);
            for (int e = 0; e < 128; ++e) descriptors[(size_t)out * 128 + e] = (uint8_t)(int)(512 * des[e] + 0.5);
            /* CreateGlobalKeyPointList_Kernel, ProgramCU.cu:2049-2080 */
            const float posX = keyLocScale * (fin[L][k].x - 0.5f) + keyLocOffset, posY = keyLocScale * (fin[L][k].y - 0.5f) + keyLocOffset;
            const int ix = (int)roundf(posX * (float)(P->depthWidth - 1) / (float)(P->width - 1)), iy = (int)roundf(posY * (float)(P->depthHeight - 1) / (float)(P->height - 1));
            float* kp = keyPoints + 4 * (size_t)out;
            kp[0] = posX; kp[1] = posY; kp[2] = keyLocScale * fin[L][k].s; kp[3] = depth[(size_t)iy * P->depthWidth + ix];
        }
    }
    for (int L = 0; L < OCTAVES * DOG_LEVELS; ++L) { free(rawX[L]); free(rawY[L]); free(fin[L]); }
    for (int o = 0; o < OCTAVES; ++o) for (int l = 0; l < LEVELS; ++l) { free(oc[o].gus[l]); free(oc[o].dog[l]); free(oc[o].grd[l]); free(oc[o].rot[l]); }
    return out;
}

/* the Gaussian filter bank, for tests: sigmas [6], widths [6], taps [6][33] */
ORC_API void orc_sift_filter_bank(float* sigmas, int32_t* widths, float* taps) {
    SiftParam sp; parse_param(&sp);
    for (int i = 0; i < LEVELS; ++i) { sigmas[i] = sp.sigmas[i]; widths[i] = sp.fw[i]; memcpy(taps + i * KERNEL_MAX_WIDTH, sp.kernel[i], sizeof sp.kernel[i]); }
}
