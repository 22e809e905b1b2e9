/*
 * marchingcubes_oracle.c -- CPU restatement of the reference's iso-surface extraction from the hashed TSDF (second half).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under bundlefusion_b200/ may include, link or call this file; only tests/ use it (as the checker).
 *
 * What it restates (FL/ = FriedLiver/Source/):
 *   extractIsoSurfaceKernel                                            FL/DepthSensing/CUDAMarchingCubesSDF.cu:15-29   (one thread per voxel of every occupied hash entry)
 *   MarchingCubesData::extractIsoSurfaceAtPosition, vertexInterp, isInBoxAA, appendTriangle   FL/DepthSensing/MarchingCubesSDFUtil.h:121-271
 *   edgeTable / triTable                                               FL/DepthSensing/Tables.h:21, 58   (Bourke's public-domain tables; the edge table is derived from the corner signs)
 *   the trilinear sample and the voxel fetch are raycast_oracle.c's (RayCastSDFUtil.h:100-121, VoxelUtilHashSDF.h:407-418)
 *
 * PARITY STATUS.  Pinned against the reference's OWN kernel executed on the CPU (oracle/build_ref.py: build_marchingcubes_emulated compiles CUDAMarchingCubesSDF.cu
 * with MarchingCubesSDFUtil.h and Tables.h against the CUDA emulation of oracle/ref_emu; tests/golden/marchingcubes_reference_emulated.npz,
 * tests/test_marchingcubes_reference_emulated.py): the multiset of triangles, every vertex position and colour bit for bit, and both tables entry by entry.
 * The order of the triangle soup is not part of the statement: the reference appends with one atomicAdd per triangle.  Here: hash-entry order, then voxel
 * index (x fastest), then the table's triangle order.
 *
 * Arithmetic contract: as raycast_oracle.c (IEEE binary32, every operation individually rounded, the reference's expression order).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../include/bf_marchingcubes.h"

#define ORC_API __attribute__((visibility("default")))

extern int orc_rc_trilinear(const BFHashDataStruct* hd, const BFHashParams* hp, float x, float y, float z, float* dist, uint8_t color[3]);     /* raycast_oracle.c */
extern BFVoxel orc_rc_voxel(const BFHashDataStruct* hd, const BFHashParams* hp, float x, float y, float z);

/* triangle table: one string per case, one hexadecimal digit per triangle corner = the cube edge it lies on */
static const char* const kTri[256] = {
    "", "083", "019", "183981", "12a", "08312a", "92a029", "2832a8a98",
    "3b2", "0b28b0", "19023b", "1b219b98b", "3a1ba3", "0a108a8ba", "3903b9ba9", "98aa8b",
    "478", "430734", "019847", "419471731", "12a847", "34730412a", "92a902847", "2a9297273794",
    "8473b2", "b47b24204", "90184723b", "47b94b9b2921", "3a13ba784", "1ba14b1047b4", "47890b9bab03", "47b4b99ba",
    "954", "954083", "054150", "854835315", "12a954", "30812a495", "52a542402", "2a5325354348",
    "95423b", "0b208b495", "05401523b", "21525828b485", "a3ba13954", "4950818a18ba", "54050b5bab03", "54858aa8b",
    "978579", "930953573", "078017157", "153357", "978957a12", "a12950530573", "802825857a52", "2a5253357",
    "7957893b2", "95797292027b", "23b018178157", "b21b17715", "958857a13a3b", "5705097b010aba0", "ba0b03a50807570", "ba57b5",
    "a65", "0835a6", "9015a6", "1831985a6", "165261", "165126308", "965906026", "598582526328",
    "23ba65", "b08b20a65", "01923b5a6", "5a61929b298b", "63b653513", "08b0b50515b6", "3b6036065059", "65969bb98",
    "5a6478", "43047365a", "1905a6847", "a65197173794", "612651478", "125526304347", "847905065026", "739794329596269",
    "3b2784a65", "5a647242027b", "01947823b5a6", "9219b294b7b45a6", "8473b53515b6", "51b5b610b7b404b", "059065036b63847", "65969b4797b9",
    "a4964a", "4a649a083", "a01a60640", "83181686461a", "149124264", "308129249264", "024426", "832824426",
    "a49a64b23", "08228b49a4a6", "3b201606461a", "64161a48121b8b1", "964936913b63", "8b1810b61914641", "3b6360064", "648b68",
    "7a678a89a", "0730a709a67a", "a671a7178180", "a67a71173", "126168189867", "269291679093739", "780706602", "732672",
    "23ba68a89867", "20727b09767a9a7", "1801781a767a23b", "b21b17a61671", "896867916b63136", "091b67", "7807063b0b60", "7b6",
    "76b", "308b76", "019b76", "819831b76", "a126b7", "12a3086b7", "2902a96b7", "6b72a3a83a98",
    "723627", "708760620", "276237019", "162186198876", "a76a17137", "a7617a187108", "03707a0a96a7", "76a7a88a9",
    "684b86", "36b306046", "86b846901", "946963931b36", "6846b82a1", "12a30b06b046", "4b846b0292a9", "a93a32943b36463",
    "823842462", "042462", "190234246438", "194142246", "8138618466a1", "a10a06604", "4634386a3039a93", "a946a4",
    "49576b", "083495b76", "50154076b", "b76834354315", "954a1276b", "6b712a083495", "76b54a42a402", "348354325a52b76",
    "723762549", "954086062687", "362376150540", "628687218485158", "954a16176137", "16a176107870954", "40a4a503a6a737a", "76a7a854a48a",
    "6956b9b89", "36b063056095", "0b805b01556b", "6b3635531", "12a95b9b8b56", "0b306b09656912a", "b85b56805a52025", "6b36352a3a53",
    "589528562382", "956960062", "158180568382628", "156216", "13616a386569896", "a10a06950560", "03856a", "a56",
    "b5a75b", "b5ab75830", "5b75ab190", "a75ab7981831", "b12b71751", "08312717572b", "9759279022b7", "75272b592328982",
    "25a235375", "820852875a25", "9015a35373a2", "982921872a25752", "135375", "087071175", "903935537", "987597",
    "5845a8ab8", "5045b05abb30", "01984a8aba45", "ab4a45b34941314", "2512852b8458", "04b0b345b2b151b", "0250592b5458b85", "9452b3",
    "25a352345384", "5a2524420", "3a235a385458019", "5a2524192942", "845853351", "045105", "845853905035", "945",
    "4b749b9ab", "0834979b79ab", "1ab1b414074b", "3143481a474bab4", "4b79b492b912", "9749b791b2b1083", "b74b42240", "b74b42834324",
    "29a279237749", "9a7974a27870207", "37a3a274a1a040a", "1a2874", "491417713", "491417081871", "403743", "487",
    "9a8ab8", "30939bb9a", "01a0a88ab", "31ab3a", "12b1b99b8", "30939b1292b9", "02b80b", "32b",
    "23828aa89", "9a2092", "23828a0181a8", "1a2", "138918", "091", "038", "",
};
/* cube corners in the table's numbering as offsets of half a voxel: v0 = 010, v1 = 110, v2 = 100, v3 = 000, v4 = 011, v5 = 111, v6 = 101, v7 = 001 */
static const int kCorner[8][3] = { {0,1,0}, {1,1,0}, {1,0,0}, {0,0,0}, {0,1,1}, {1,1,1}, {1,0,1}, {0,0,1} };
/* cube edge -> its two corners */
static const int kEdge[12][2] = { {0,1}, {1,2}, {2,3}, {3,0}, {4,5}, {5,6}, {6,7}, {7,4}, {0,4}, {1,5}, {2,6}, {3,7} };

ORC_API void orc_marchingcubes_tables(int edgeTable[256], int triTable[256][16]) {
    for (int c = 0; c < 256; ++c) {
        int m = 0;
        for (int e = 0; e < 12; ++e) if (((c >> kEdge[e][0]) & 1) != ((c >> kEdge[e][1]) & 1)) m |= 1 << e;
        edgeTable[c] = m;
        const int n = (int)strlen(kTri[c]);
        for (int i = 0; i < 16; ++i) {
            if (i >= n) { triTable[c][i] = -1; continue; }
            const char ch = kTri[c][i];
            triTable[c][i] = ch <= '9' ? ch - '0' : ch - 'a' + 10;
        }
    }
}

typedef struct { float p[3]; float c[3]; } vert;

/* vertexInterp, MarchingCubesSDFUtil.h:211-233 */
static vert vertex_interp(float isolevel, const float p1[3], const float p2[3], float d1, float d2, const uint8_t c1[4], const uint8_t c2[4]) {
    vert r1, r2, res;
    for (int k = 0; k < 3; ++k) { r1.p[k] = p1[k]; r1.c[k] = (float)c1[k] / 255.f; r2.p[k] = p2[k]; r2.c[k] = (float)c2[k] / 255.f; }
    if (fabsf(isolevel - d1) < 0.00001f) return r1;
    if (fabsf(isolevel - d2) < 0.00001f) return r2;
    if (fabsf(d1 - d2) < 0.00001f) return r1;
    const float mu = (isolevel - d1) / (d2 - d1);
    for (int k = 0; k < 3; ++k) {
        res.p[k] = p1[k] + mu * (p2[k] - p1[k]);
        res.c[k] = (float)((float)c1[k] + mu * (float)((int)c2[k] - (int)c1[k])) / 255.f;
    }
    return res;
}

/* extractIsoSurfaceAtPosition, MarchingCubesSDFUtil.h:121-209: appends to tri[*n] while *n < cap; returns the number of triangles of this cell */
static int extract_at(const BFHashDataStruct* hd, const BFHashParams* hp, const BFMarchingCubesParams* p, const float wp[3], BFMarchingCubesTriangle* tri, unsigned* n) {
    if (p->m_boxEnabled == 1) for (int k = 0; k < 3; ++k) if (wp[k] < p->m_minCorner[k] || wp[k] > p->m_maxCorner[k]) return 0;
    const float isolevel = 0.0f;
    const float P = hp->m_virtualVoxelSize / 2.0f, M = -P;
    float pos[8][3], d[8];
    int valid = 1;
    for (int v = 0; v < 8; ++v) {
        uint8_t col[3];
        for (int k = 0; k < 3; ++k) pos[v][k] = wp[k] + (kCorner[v][k] ? P : M);
        d[v] = 0.0f;
        if (!orc_rc_trilinear(hd, hp, pos[v][0], pos[v][1], pos[v][2], &d[v], col)) valid = 0;
    }
    if (!valid) return 0;
    unsigned cube = 0;
    for (int v = 0; v < 8; ++v) if (d[v] < isolevel) cube += 1u << v;
    /* the reference's distArray = {000, 100, 010, 001, 110, 011, 101, 111} = corners 3 2 0 7 1 4 6 5, all ordered pairs */
    static const int order[8] = { 3, 2, 0, 7, 1, 4, 6, 5 };
    const float thres = p->m_threshMarchingCubes;
    for (int k = 0; k < 8; ++k)
        for (int l = 0; l < 8; ++l) {
            const float a = d[order[k]], b = d[order[l]];
            if (a * b < 0.0f) { if (fabsf(a) + fabsf(b) > thres) return 0; }
            else if (fabsf(a - b) > thres) return 0;
        }
    for (int v = 0; v < 8; ++v) if (fabsf(d[v]) > p->m_threshMarchingCubes2) return 0;
    int mask = 0;
    for (int e = 0; e < 12; ++e) if (((cube >> kEdge[e][0]) & 1) != ((cube >> kEdge[e][1]) & 1)) mask |= 1 << e;
    if (mask == 0 || mask == 255) return 0;
    const BFVoxel vox = orc_rc_voxel(hd, hp, wp[0], wp[1], wp[2]);
    vert list[12];
    for (int e = 0; e < 12; ++e)
        if (mask & (1 << e)) list[e] = vertex_interp(isolevel, pos[kEdge[e][0]], pos[kEdge[e][1]], d[kEdge[e][0]], d[kEdge[e][1]], vox.color, vox.color);
    const char* row = kTri[cube];
    int made = 0;
    for (int i = 0; row[i]; i += 3, ++made) {
        if (*n >= p->m_maxNumTriangles) continue;                     /* appendTriangle: a full buffer drops the triangle */
        vert t[3];
        for (int c = 0; c < 3; ++c) { const char ch = row[i + c]; t[c] = list[ch <= '9' ? ch - '0' : ch - 'a' + 10]; }
        BFMarchingCubesTriangle* o = &tri[(*n)++];
        memcpy(&o->v0, &t[0], sizeof(vert)); memcpy(&o->v1, &t[1], sizeof(vert)); memcpy(&o->v2, &t[2], sizeof(vert));
    }
    return made;
}

/* extractIsoSurfaceKernel over the whole hash table; returns the number of triangles written (<= m_maxNumTriangles); *found = triangles the cells produced */
ORC_API unsigned orc_marchingcubes_extract(const BFHashDataStruct* hd, const BFHashParams* hp, const BFMarchingCubesParams* p, BFMarchingCubesTriangle* tri, unsigned long long* found) {
    unsigned n = 0;
    unsigned long long all = 0;
    const unsigned total = hp->m_hashNumBuckets * BF_HASH_BUCKET_SIZE;
    for (unsigned e = 0; e < total; ++e) {
        const BFHashEntry* he = &hd->d_hash[e];
        if (he->ptr == BF_FREE_ENTRY) continue;
        for (int t = 0; t < BF_SDF_BLOCK_SIZE * BF_SDF_BLOCK_SIZE * BF_SDF_BLOCK_SIZE; ++t) {
            const int pi[3] = { he->pos[0] * BF_SDF_BLOCK_SIZE + (t & 7), he->pos[1] * BF_SDF_BLOCK_SIZE + ((t >> 3) & 7), he->pos[2] * BF_SDF_BLOCK_SIZE + (t >> 6) };
            const float wp[3] = { (float)pi[0] * hp->m_virtualVoxelSize, (float)pi[1] * hp->m_virtualVoxelSize, (float)pi[2] * hp->m_virtualVoxelSize };
            all += (unsigned long long)extract_at(hd, hp, p, wp, tri, &n);
        }
    }
    if (found) *found = all;
    return n;
}
