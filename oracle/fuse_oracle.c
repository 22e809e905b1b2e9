/*
 * fuse_oracle.c -- CPU restatement of the chunk -> keyframe fusion of the sparse features:
 *   SIFTImageManager::computeTracks   /root/reference/FriedLiver/Source/SiftGPU/SIFTImageManager.cpp:381-411 (findTrack :366-378)
 *   SIFTImageManager::fuseToGlobal    .../SIFTImageManager.cpp:413-476
 * (host code in the reference: it copies keys, descriptors, correspondences and poses to the CPU, recurses there and uploads the result).
 *
 * TEST INFRASTRUCTURE ONLY: the oracle the CUDA path (bundlefusion_b200/csrc/sift_fuse.cu) is checked against; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.  PARITY STATUS: pinned against the reference's OWN code -- fuseToGlobal / computeTracks /
 * findTrack are host code of its manager class; oracle/build_ref.py (build_fuse_emulated) compiles SIFTImageManager.cpp + .cu against the CUDA emulation, and on the solved
 * chunks of tests/test_fuse_reference_emulated.py (golden: tests/golden/fuse_reference_emulated.npz; live where oracle/_ref is built) the fused keyframe's keys and
 * descriptors agree bit for bit: which key represents a track, which correspondences contribute to its position, the order of the fused keys.
 *
 * Arithmetic: IEEE binary32, individually rounded operations (the reference code here is MSVC host code), float4x4 * float3 evaluated as
 * ((m0 x + m1 y) + m2 z) + m3 (cuda_SimpleMatrixUtil.h:937-944).  Compile with -ffp-contract=off.
 *
 * Layout: key k of image i has the global index i * keyStride + k (the reference packs keys by a prefix sum; the order of global
 * indices -- image-major, key-minor -- is what the algorithm depends on, and it is the same).  Differences from the reference, both
 * outside what its configuration can reach (<= 11 images x ~150 features): when more than maxKeys tracks survive the reference sorts
 * the KEYS by depth with an unstable sort and leaves the descriptors unsorted; here the first maxKeys tracks are kept in track order.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))
#define MAX_TRACK_CORR_ERROR 0.03f          /* SIFTImageManager.cpp:380 */

typedef struct { uint32_t i, j; float pi[3], pj[3]; } EntryJ;            /* FL/SiftGPU/SIFTImageManager.h:45-60 */
typedef struct { float x, y, scale, depth; } KeyPoint;
typedef struct { uint32_t img, key; float pos[3]; } TrackItem;          /* std::pair<uint2, float3> */

static void xform(const float* M, const float* p, float* o) {
    for (int r = 0; r < 3; ++r) o[r] = ((M[4 * r] * p[0] + M[4 * r + 1] * p[1]) + M[4 * r + 2] * p[2]) + M[4 * r + 3];
}

typedef struct { TrackItem* items; uint32_t* start; uint32_t n; } Adj;     /* corrPerKey */

static void find_track(const Adj* adj, uint8_t* marker, TrackItem* track, uint32_t* trackLen, uint32_t curKey) {
    for (uint32_t e = adj->start[curKey]; e < adj->start[curKey + 1]; ++e) {
        const TrackItem c = adj->items[e];
        if (!marker[c.key]) {
            track[(*trackLen)++] = c;
            marker[c.key] = 1;
            find_track(adj, marker, track, trackLen, c.key);
        }
    }
}

/* returns the number of keys written.  keyIdx: [numCorr][2] global key indices (uint2), transforms: [numImages][16] row-major,
 * K: colour intrinsics 4x4.  trackOfKey (optional, [numImages * keyStride]): for tests, the output index each key's track went to or -1. */
ORC_API int orc_sift_fuse_to_global(const EntryJ* corr, const uint32_t* keyIdx, uint32_t numCorr, const float* transforms, uint32_t numImages,
                                    const KeyPoint* keys, const uint8_t* descs, const int32_t* numKeysPerImage, uint32_t keyStride, const float* K,
                                    KeyPoint* outKeys, uint8_t* outDescs, uint32_t maxKeys) {
    const uint32_t M = numImages * keyStride;
    uint32_t* cnt = (uint32_t*)calloc(M + 1, sizeof(uint32_t));
    uint32_t* fill = (uint32_t*)calloc(M + 1, sizeof(uint32_t));
    /* corrPerKey in push order (ascending correspondence index) */
    for (uint32_t c = 0; c < numCorr; ++c) if (corr[c].i != 0xFFFFFFFFu) { cnt[keyIdx[2 * c]]++; cnt[keyIdx[2 * c + 1]]++; }
    Adj adj; adj.start = (uint32_t*)malloc(sizeof(uint32_t) * (M + 1)); adj.n = 0;
    for (uint32_t k = 0; k < M; ++k) { adj.start[k] = adj.n; adj.n += cnt[k]; }
    adj.start[M] = adj.n;
    adj.items = (TrackItem*)malloc(sizeof(TrackItem) * (adj.n + 1));
    const float ninf = -INFINITY;
    for (uint32_t c = 0; c < numCorr; ++c) {
        const EntryJ* e = &corr[c];
        if (e->i == 0xFFFFFFFFu) continue;
        const uint32_t kx = keyIdx[2 * c], ky = keyIdx[2 * c + 1];
        float a[3], b[3];
        xform(transforms + 16 * (size_t)e->i, e->pi, a); xform(transforms + 16 * (size_t)e->j, e->pj, b);
        const float d0 = a[0] - b[0], d1 = a[1] - b[1], d2 = a[2] - b[2];
        const float err = sqrtf((d0 * d0 + d1 * d1) + d2 * d2);
        const int ok = err < MAX_TRACK_CORR_ERROR;
        TrackItem tx = { e->j, ky, { ok ? e->pj[0] : ninf, ok ? e->pj[1] : ninf, ok ? e->pj[2] : ninf } };
        TrackItem ty = { e->i, kx, { ok ? e->pi[0] : ninf, ok ? e->pi[1] : ninf, ok ? e->pi[2] : ninf } };
        adj.items[adj.start[kx] + fill[kx]++] = tx;
        adj.items[adj.start[ky] + fill[ky]++] = ty;
    }
    uint8_t* marker = (uint8_t*)calloc(M + 1, 1);
    TrackItem* track = (TrackItem*)malloc(sizeof(TrackItem) * (M + 1));
    uint32_t numOut = 0;
    for (uint32_t i = 0; i < numImages; ++i)
        for (uint32_t k = 0; k < (uint32_t)numKeysPerImage[i]; ++k) {
            uint32_t len = 0;
            find_track(&adj, marker, track, &len, i * keyStride + k);
            if (len == 0) continue;
            /* fuseToGlobal: average of the world positions of the track's usable members, projected into the chunk's first frame */
            float pos[3] = { 0.0f, 0.0f, 0.0f }; unsigned num = 0;
            for (uint32_t t = 0; t < len; ++t)
                if (track[t].pos[0] != ninf) { float w[3]; xform(transforms + 16 * (size_t)track[t].img, track[t].pos, w); pos[0] += w[0]; pos[1] += w[1]; pos[2] += w[2]; ++num; }
            if (num == 0) continue;
            pos[0] /= (float)num; pos[1] /= (float)num; pos[2] /= (float)num;
            float p[3]; xform(K, pos, p);
            if (numOut < maxKeys) {
                const uint32_t rep = track[0].key;                          /* "arbitrarily pick a key for the descriptor": the track's front */
                outKeys[numOut].x = p[0] / p[2]; outKeys[numOut].y = p[1] / p[2]; outKeys[numOut].scale = keys[rep].scale; outKeys[numOut].depth = p[2];
                memcpy(outDescs + 128 * (size_t)numOut, descs + 128 * (size_t)rep, 128);
                ++numOut;
            }
        }
    free(cnt); free(fill); free(adj.start); free(adj.items); free(marker); free(track);
    return (int)numOut;
}
