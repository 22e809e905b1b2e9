/*
 * sift_oracle.c -- CPU restatement of the reference's SIFT descriptor matcher:
 * all-pairs u8 x u8 dot products, per-row and per-column best / second-best, distance and ratio tests,
 * mutual-best check.
 *
 * TEST INFRASTRUCTURE ONLY (see tsdf_oracle.c header for the rules).  PARITY STATUS: PINNED against the reference's own
 * kernels (MultiplyDescriptor / RowMatch / ColMatch through SiftMatchGPU::GetSiftMatch) executed on the CPU: ProgramCU.cu uses
 * texture references, which nvcc 12 no longer compiles, so oracle/build_ref.py builds it with g++ against a CPU emulation of
 * CUDA (oracle/_ref/libref_sift_emulated.so); its outputs on seeded descriptor sets, exact ties and key-point offsets included,
 * are committed as tests/golden/sift_reference_emulated.npz and this file reproduces them bit for bit -- counters, index pairs,
 * distances (tests/test_sift_reference_emulated.py).  Also pinned by the known-answer tests in tests/test_sift_oracle.py (a
 * brute-force numpy restatement of the published SiftGPU matching rule, planted matches, tie cases).  Not yet compared with a run
 * of the reference on a GPU.
 *
 * Restates (FL/ = FriedLiver/Source/):
 *   MultiplyDescriptor_Kernel   FL/SiftGPU/ProgramCU.cu:1634-1731  dot[i][j] = sum_k d1[i][k] * d2[j][k] (int32, exact),
 *                                                                   + per 4-row block (max, row, second) per column
 *   RowMatch_Kernel             FL/SiftGPU/ProgramCU.cu:1772-1831  32 strided lanes + tree: (max, col, second) per row
 *   ColMatch_Kernel             FL/SiftGPU/ProgramCU.cu:1852-1918  32 strided lanes + tree over the 4-row blocks, mutual check
 * The lane / tree structure is kept literally because it fixes which index wins a TIE: with the strict `>` everywhere and a tree
 * that folds slot t + step into slot t, the winner among equal maxima is the one with the smallest (bit-reversed lane, position)
 * -- lane = col % 32 for rows, (row / 4) % 32 for columns -- and "second" is the second largest of the MULTISET (a duplicated
 * maximum is its own second).
 * dist = acosf(min(dot * 2^-18, 1)); accepted iff dist < distmax && dist < distn * ratiomax.
 * Output order: the reference appends with atomicAdd (race-dependent); the oracle emits matches in ascending column
 * (feature-of-image-2) order.  Compare as sets, or after the distance sort of SortKeyPointMatchesCU.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))
#define LANES 32
#define ROWS_PER_BLOCK 4           /* MULT_BLOCK_DIMY, ProgramCU.cu:1627 */
#define MAX_RAW 128                /* MAX_MATCHES_PER_IMAGE_PAIR_RAW, FL/GlobalDefines.h:8 */

/* dot[n1][n2] */
ORC_API void orc_sift_multiply(const uint8_t* d1, int n1, const uint8_t* d2, int n2, int32_t* dot) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n1; ++i)
        for (int j = 0; j < n2; ++j) {
            int32_t s = 0;
            for (int k = 0; k < 128; ++k) s += (int32_t)d1[(size_t)i * 128 + k] * (int32_t)d2[(size_t)j * 128 + k];
            dot[(size_t)i * n2 + j] = s;
        }
}

static inline float dist_of(int32_t dot) { return acosf(fminf((float)dot * 0.000003814697265625f, 1.0f)); }

/* RowMatch_Kernel: rowResult[i] = best column or -1, rowDist[i] = dist of the best */
ORC_API void orc_sift_row_match(const int32_t* dot, int n1, int n2, float distmax, float ratiomax, int32_t* rowResult, float* rowDist) {
    for (int row = 0; row < n1; ++row) {
        int32_t mx[LANES], nx[LANES], ix[LANES];
        for (int t = 0; t < LANES; ++t) {
            int32_t m = 0, n = 0, id = -1;
            for (int i = 0; i < n2; i += LANES) {
                if (t + i < n2) {
                    const int32_t v = dot[(size_t)row * n2 + t + i];
                    const int test = v > m;
                    n = test ? m : (n > v ? n : v);
                    id = test ? (t + i) : id;
                    m = test ? v : m;
                }
            }
            mx[t] = m; nx[t] = n; ix[t] = id;
        }
        for (int step = LANES / 2; step > 0; step /= 2)
            for (int t = 0; t < step; ++t) {
                const int32_t v1 = mx[t], v2 = mx[t + step];
                const int test = v2 > v1;
                nx[t] = test ? (v1 > nx[t + step] ? v1 : nx[t + step]) : (nx[t] > v2 ? nx[t] : v2);
                ix[t] = test ? ix[t + step] : ix[t];
                mx[t] = test ? v2 : v1;
            }
        const float dist = dist_of(mx[0]), distn = dist_of(nx[0]);
        rowResult[row] = (dist < distmax && dist < distn * ratiomax) ? ix[0] : -1;
        rowDist[row] = dist;
    }
}

/* MultiplyDescriptor's per-block column partials + ColMatch_Kernel.  Returns the number of mutual matches found (the
 * reference's counter, which keeps counting past MAX_RAW); writes at most MAX_RAW of them, ascending in column. */
ORC_API int orc_sift_col_match(const int32_t* dot, int n1, int n2, float distmax, float ratiomax, const int32_t* rowResult, const float* rowDist,
                               uint32_t offX, uint32_t offY, uint32_t* outIdx /* [MAX_RAW][2] */, float* outDist /* [MAX_RAW] */) {
    const int height = (n1 + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK;
    int count = 0;
    for (int col = 0; col < n2; ++col) {
        int32_t rx[LANES], ry[LANES], rz[LANES];
        for (int t = 0; t < LANES; ++t) {
            int32_t lx = 0, ly = -1, lz = 0;
            for (int b = t; b < height; b += LANES) {
                /* the block's (max, row, second), ProgramCU.cu:1707-1721 */
                int32_t cx = 0, cy = -1, cz = 0;
                for (int i = 0; i < ROWS_PER_BLOCK; ++i) {
                    const int row = b * ROWS_PER_BLOCK + i;
                    if (row >= n1) break;
                    const int32_t r = dot[(size_t)row * n2 + col];
                    if (r > cx) { cz = cx; cx = r; cy = row; } else { cz = cz > r ? cz : r; }
                }
                if (lx < cx) { lz = lx > cz ? lx : cz; lx = cx; ly = cy; } else { lz = lz > cx ? lz : cx; }
            }
            rx[t] = lx; ry[t] = ly; rz[t] = lz;
        }
        for (int step = LANES / 2; step > 0; step /= 2)
            for (int t = 0; t < step; ++t) {
                if (rx[t] < rx[t + step]) { rz[t] = rx[t] > rz[t + step] ? rx[t] : rz[t + step]; rx[t] = rx[t + step]; ry[t] = ry[t + step]; }
                else rz[t] = rz[t] > rx[t + step] ? rz[t] : rx[t + step];
            }
        const float dist = dist_of(rx[0]), distn = dist_of(rz[0]);
        const int f1 = (dist < distmax && dist < distn * ratiomax) ? ry[0] : -1;
        if (f1 >= 0 && rowResult[f1] == col) {
            if (count < MAX_RAW) { outIdx[2 * count] = (uint32_t)f1 + offX; outIdx[2 * count + 1] = (uint32_t)col + offY; outDist[count] = rowDist[f1]; }
            ++count;
        }
    }
    return count;
}

/* SiftMatchGPU::GetSiftMatch (FL/SiftGPU/SiftMatch.cpp:160-196) for one pair */
ORC_API int orc_sift_match(const uint8_t* d1, int n1, const uint8_t* d2, int n2, float distmax, float ratiomax, uint32_t offX, uint32_t offY,
                           uint32_t* outIdx, float* outDist) {
    if (n1 <= 0 || n2 <= 0) return 0;
    int32_t* dot = (int32_t*)malloc(sizeof(int32_t) * (size_t)n1 * n2);
    int32_t* rr = (int32_t*)malloc(sizeof(int32_t) * n1);
    float* rd = (float*)malloc(sizeof(float) * n1);
    orc_sift_multiply(d1, n1, d2, n2, dot);
    orc_sift_row_match(dot, n1, n2, distmax, ratiomax, rr, rd);
    const int c = orc_sift_col_match(dot, n1, n2, distmax, ratiomax, rr, rd, offX, offY, outIdx, outDist);
    free(dot); free(rr); free(rd);
    return c;
}

/* SortKeyPointMatchesCU (FL/SiftGPU/SIFTImageManager.cu:59-177): ascending distance; equal distances by (image-2 feature, image-1 feature)
 * -- the reference's odd-even transposition sort never swaps equal distances, i.e. it is stable with respect to an append order that is
 * itself race-dependent; the oracle's append order is ascending image-2 feature, so both readings coincide here. */
ORC_API void orc_sift_sort_matches(unsigned curFrame, unsigned startFrame, unsigned numFrames, const int32_t* numMatches, float* dists, uint32_t* idxs) {
    for (unsigned p = startFrame; p < numFrames; ++p) {
        if (p == curFrame) continue;
        int n = numMatches[p] < MAX_RAW ? numMatches[p] : MAX_RAW;
        float* d = dists + (size_t)p * MAX_RAW; uint32_t* ix = idxs + 2 * (size_t)p * MAX_RAW;
        for (int i = 1; i < n; ++i) {                          /* insertion sort on the total order */
            const float dv = d[i]; const uint32_t a = ix[2 * i], b = ix[2 * i + 1];
            int j = i - 1;
            while (j >= 0 && (d[j] > dv || (d[j] == dv && (ix[2 * j + 1] > b || (ix[2 * j + 1] == b && ix[2 * j] > a))))) {
                d[j + 1] = d[j]; ix[2 * j + 2] = ix[2 * j]; ix[2 * j + 3] = ix[2 * j + 1]; --j;
            }
            d[j + 1] = dv; ix[2 * j + 2] = a; ix[2 * j + 3] = b;
        }
    }
}
