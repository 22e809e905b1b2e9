/*
 * raycast_oracle.c -- CPU restatement of the reference's ray cast of the hashed TSDF.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under bundlefusion_b200/ may include, link or call this file; only tests/ use it (as the checker).
 *
 * What it restates (FL/ = FriedLiver/Source/):
 *   rayIntervalSplatKernel                    FL/DepthSensing/CUDARayCastSDF.cu:90-172      one screen-space quad per in-frustum block
 *   the D3D11 pass that rasterises the quads  FL/DepthSensing/DX11RayIntervalSplatting.cpp:137-229, FriedLiver/Shaders/RayIntervalSplatting.hlsl
 *   renderKernel                              FL/DepthSensing/CUDARayCastSDF.cu:17-58
 *   RayCastData::traverseCoarseGridSimpleSampleAll, trilinearInterpolationSimpleFastFast, findIntersectionBisection, gradientForPoint,
 *   depthToCamera, cameraToDepthProj, depthProjToCameraZ       FL/DepthSensing/RayCastSDFUtil.h:87-280
 *   HashDataStruct::getVoxel(float3), worldToVirtualVoxelPos, virtualVoxelPosToLocalSDFBlockIndex   FL/DepthSensing/VoxelUtilHashSDF.h:276-358, 407-418
 *   computeNormals_Kernel                     FL/CUDAImageUtil.cu:404-431                   (normals from the rendered positions when m_useGradients is off)
 *
 * PARITY STATUS.  renderKernel and its helpers: pinned against the reference's OWN kernel executed on the CPU (oracle/build_ref.py builds
 * CUDARayCastSDF.cu against the CUDA emulation of oracle/ref_emu; tests/golden/raycast_reference_emulated.npz, tests/test_raycast_reference_emulated.py):
 * depth, positions, colours and gradient normals bit for bit on the committed scene.  The interval splat has TWO halves: the quad of a block (the
 * reference's CUDA kernel: pinned the same way) and its rasterisation, which the reference leaves to Direct3D 11 (depth-tested triangle
 * draws into two render targets) -- there is no Direct3D here and its fixed-function rasteriser (1/256-pixel vertex snapping, top-left fill rule) is not
 * the reference's code.  That half is restated from the D3D11 rules on exact float coordinates: PARITY UNPINNED for the set of pixels a quad covers when
 * an edge falls within 1/256 pixel of a pixel centre.  A pixel's interval only fixes where its ray starts marching; tests bound the effect.
 *
 * Arithmetic contract (shared with bundlefusion_b200/csrc/raycast.cu, compared bit for bit): IEEE binary32, every + - * / individually rounded
 * (-ffp-contract=off here, -fmad=false there), expressions in the reference's order; normalize() multiplies by 1 / sqrtf(v.v) (the reference's
 * rsqrtf is an approximate GPU instruction; its CPU emulation uses the same exact form); float -> int as cvt.rzi (truncate, saturate, NaN -> 0).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../include/bf_raycast.h"
#include "../include/bf_tsdf.h"

#define ORC_API __attribute__((visibility("default")))

typedef struct { float x, y, z; } f3;
typedef struct { int x, y, z; } i3;

extern int orc_tsdf_find(const BFHashDataStruct* hd, const BFHashParams* hp, int bx, int by, int bz);      /* tsdf_oracle.c */

static inline int f2i(float v) {
    if (v != v) return 0;
    if (v >= 2147483648.0f) return INT32_MAX;
    if (v <= -2147483648.0f) return INT32_MIN;
    return (int)v;
}
static inline int isign(float v) { return (0.0f < v) - (v < 0.0f); }
static inline f3 add3(f3 a, f3 b) { f3 r = { a.x + b.x, a.y + b.y, a.z + b.z }; return r; }
static inline f3 sub3(f3 a, f3 b) { f3 r = { a.x - b.x, a.y - b.y, a.z - b.z }; return r; }
static inline f3 scale3(float s, f3 a) { f3 r = { s * a.x, s * a.y, s * a.z }; return r; }
static inline float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline f3 normalize3(f3 v) { const float inv = 1.0f / sqrtf(dot3(v, v)); f3 r = { v.x * inv, v.y * inv, v.z * inv }; return r; }    /* cutil_math.h:1207-1211 */
/* float4x4 * float3 (w = 1) and the xyz of float4x4 * float4(v, 0), cuda_SimpleMatrixUtil.h:925-944 */
static inline f3 mul_point(const float* M, f3 v) {
    f3 r = { M[0] * v.x + M[1] * v.y + M[2] * v.z + M[3] * 1.0f, M[4] * v.x + M[5] * v.y + M[6] * v.z + M[7] * 1.0f, M[8] * v.x + M[9] * v.y + M[10] * v.z + M[11] * 1.0f };
    return r;
}
static inline f3 mul_dir(const float* M, f3 v) {
    f3 r = { M[0] * v.x + M[1] * v.y + M[2] * v.z + M[3] * 0.0f, M[4] * v.x + M[5] * v.y + M[6] * v.z + M[7] * 0.0f, M[8] * v.x + M[9] * v.y + M[10] * v.z + M[11] * 0.0f };
    return r;
}

/* VoxelUtilHashSDF.h:283-287, 290-299, 347-358 */
static inline i3 world_to_voxel(const BFHashParams* hp, f3 pos) {
    f3 p = { pos.x / hp->m_virtualVoxelSize, pos.y / hp->m_virtualVoxelSize, pos.z / hp->m_virtualVoxelSize };
    i3 r = { f2i(p.x + (float)isign(p.x) * 0.5f), f2i(p.y + (float)isign(p.y) * 0.5f), f2i(p.z + (float)isign(p.z) * 0.5f) };
    return r;
}
static inline i3 voxel_to_block(i3 v) {
    if (v.x < 0) v.x -= BF_SDF_BLOCK_SIZE - 1;
    if (v.y < 0) v.y -= BF_SDF_BLOCK_SIZE - 1;
    if (v.z < 0) v.z -= BF_SDF_BLOCK_SIZE - 1;
    i3 r = { v.x / BF_SDF_BLOCK_SIZE, v.y / BF_SDF_BLOCK_SIZE, v.z / BF_SDF_BLOCK_SIZE };
    return r;
}
static inline int local_index(i3 v) {
    i3 l = { v.x % BF_SDF_BLOCK_SIZE, v.y % BF_SDF_BLOCK_SIZE, v.z % BF_SDF_BLOCK_SIZE };
    if (l.x < 0) l.x += BF_SDF_BLOCK_SIZE;
    if (l.y < 0) l.y += BF_SDF_BLOCK_SIZE;
    if (l.z < 0) l.z += BF_SDF_BLOCK_SIZE;
    return l.z * BF_SDF_BLOCK_SIZE * BF_SDF_BLOCK_SIZE + l.y * BF_SDF_BLOCK_SIZE + l.x;
}
/* HashDataStruct::getVoxel(const float3&), VoxelUtilHashSDF.h:407-418: a missing block reads as the empty voxel */
static inline BFVoxel get_voxel(const BFHashDataStruct* hd, const BFHashParams* hp, f3 worldPos) {
    const i3 vv = world_to_voxel(hp, worldPos), b = voxel_to_block(vv);
    const int e = orc_tsdf_find(hd, hp, b.x, b.y, b.z);
    BFVoxel v; memset(&v, 0, sizeof(v));
    if (e >= 0) v = hd->d_SDFBlocks[hd->d_hash[e].ptr + local_index(vv)];
    return v;
}
static inline float fracf_(float v) { return v - floorf(v); }

/* trilinearInterpolationSimpleFastFast, RayCastSDFUtil.h:100-121: returns 0 at the first empty voxel, LEAVING the partial sum in *dist (gradientForPoint reads it) */
static int trilinear(const BFHashDataStruct* hd, const BFHashParams* hp, f3 pos, float* dist, uint8_t color[3]) {
    const float oSet = hp->m_virtualVoxelSize;
    const f3 half = { oSet / 2.0f, oSet / 2.0f, oSet / 2.0f };
    const f3 posDual = sub3(pos, half);
    const f3 w = { fracf_(pos.x / oSet), fracf_(pos.y / oSet), fracf_(pos.z / oSet) };
    *dist = 0.0f;
    f3 col = { 0.0f, 0.0f, 0.0f };
    /* the eight corners in the reference's order, each weight product evaluated left to right */
    static const int corner[8][3] = { {0,0,0}, {1,0,0}, {0,1,0}, {0,0,1}, {1,1,0}, {0,1,1}, {1,0,1}, {1,1,1} };
    for (int k = 0; k < 8; ++k) {
        const f3 off = { corner[k][0] ? oSet : 0.0f, corner[k][1] ? oSet : 0.0f, corner[k][2] ? oSet : 0.0f };
        const BFVoxel v = get_voxel(hd, hp, add3(posDual, off));
        if (v.weight == 0) return 0;
        const float wx = corner[k][0] ? w.x : (1.0f - w.x), wy = corner[k][1] ? w.y : (1.0f - w.y), wz = corner[k][2] ? w.z : (1.0f - w.z);
        const float ww = wx * wy * wz;
        *dist += ww * v.sdf;
        col.x += ww * (float)v.color[0]; col.y += ww * (float)v.color[1]; col.z += ww * (float)v.color[2];
    }
    color[0] = (uint8_t)f2i(col.x); color[1] = (uint8_t)f2i(col.y); color[2] = (uint8_t)f2i(col.z);          /* make_uchar3(float, float, float) */
    return 1;
}

/* the two samplers by themselves, for marchingcubes_oracle.c (same library, not exported) */
int orc_rc_trilinear(const BFHashDataStruct* hd, const BFHashParams* hp, float x, float y, float z, float* dist, uint8_t color[3]) {
    const f3 p = { x, y, z };
    return trilinear(hd, hp, p, dist, color);
}
BFVoxel orc_rc_voxel(const BFHashDataStruct* hd, const BFHashParams* hp, float x, float y, float z) {
    const f3 p = { x, y, z };
    return get_voxel(hd, hp, p);
}

/* findIntersectionBisection, RayCastSDFUtil.h:148-170 (three steps of regula falsi) */
static int bisection(const BFHashDataStruct* hd, const BFHashParams* hp, f3 camPos, f3 dir, float d0, float r0, float d1, float r1, float* alpha, uint8_t color[3]) {
    float a = r0, aDist = d0, b = r1, bDist = d1, c = 0.0f;
    for (unsigned i = 0; i < 3; ++i) {
        c = a + (aDist / (aDist - bDist)) * (b - a);
        float cDist;
        if (!trilinear(hd, hp, add3(camPos, scale3(c, dir)), &cDist, color)) return 0;
        if (aDist * cDist > 0.0f) { a = c; aDist = cDist; } else { b = c; bDist = cDist; }
    }
    *alpha = c;
    return 1;
}

/* gradientForPoint, RayCastSDFUtil.h:173-199: central differences of six interpolated samples, whatever each returned */
static f3 gradient(const BFHashDataStruct* hd, const BFHashParams* hp, f3 pos) {
    const float vs = hp->m_virtualVoxelSize;
    uint8_t c[3];
    float dm[3], dp[3];
    for (int k = 0; k < 3; ++k) { f3 o = { k == 0 ? 0.5f * vs : 0.0f, k == 1 ? 0.5f * vs : 0.0f, k == 2 ? 0.5f * vs : 0.0f }; trilinear(hd, hp, sub3(pos, o), &dm[k], c); }
    for (int k = 0; k < 3; ++k) { f3 o = { k == 0 ? 0.5f * vs : 0.0f, k == 1 ? 0.5f * vs : 0.0f, k == 2 ? 0.5f * vs : 0.0f }; trilinear(hd, hp, add3(pos, o), &dp[k], c); }
    const f3 g = { (dm[0] - dp[0]) / vs, (dm[1] - dp[1]) / vs, (dm[2] - dp[2]) / vs };
    const float l = sqrtf(dot3(g, g));
    f3 r = { 0.0f, 0.0f, 0.0f };
    if (l == 0.0f) return r;
    r.x = -g.x / l; r.y = -g.y / l; r.z = -g.z / l;
    return r;
}

static inline f3 depth_to_camera(const BFRayCastParams* p, unsigned ux, unsigned uy, float depth) {     /* RayCastSDFUtil.h:207-212 */
    const float x = ((float)ux - p->mx) / p->fx, y = ((float)uy - p->my) / p->fy;
    f3 r = { depth * x, depth * y, depth };
    return r;
}

/* renderKernel + traverseCoarseGridSimpleSampleAll for every pixel; rayMin / rayMax: the splatted interval images ([height][width], 0 or -inf = no interval) */
ORC_API void orc_raycast_render(const BFHashDataStruct* hd, const BFHashParams* hp, const BFRayCastParams* p, const float* rayMin, const float* rayMax,
                                float* depth, float* depth4, float* normals, float* colors) {
    const unsigned W = p->m_width, H = p->m_height;
    for (unsigned y = 0; y < H; ++y)
        for (unsigned x = 0; x < W; ++x) {
            const size_t i = (size_t)y * W + x;
            depth[i] = -INFINITY;
            for (int k = 0; k < 4; ++k) { depth4[4 * i + k] = -INFINITY; normals[4 * i + k] = -INFINITY; colors[4 * i + k] = -INFINITY; }
            const f3 camDir = normalize3(depth_to_camera(p, x, y, 1.0f));
            const f3 zero = { 0.0f, 0.0f, 0.0f };
            const f3 camPos = mul_point(p->m_viewMatrixInverse.m, zero);
            const f3 dir = normalize3(mul_dir(p->m_viewMatrixInverse.m, camDir));
            float mn = rayMin[i], mx = rayMax[i];
            if (mn == 0 || mn == -INFINITY) continue;
            if (mx == 0 || mx == -INFINITY) continue;
            mn = fmaxf(mn, p->m_minDepth); mx = fminf(mx, p->m_maxDepth);
            /* traverseCoarseGridSimpleSampleAll, RayCastSDFUtil.h:231-294 */
            float lastSdf = 0.0f, lastAlpha = 0.0f; unsigned lastWeight = 0;
            const float depthToRayLength = 1.0f / camDir.z;
            float rayCurrent = depthToRayLength * fmaxf(p->m_minDepth, mn);
            const float rayEnd = depthToRayLength * fminf(p->m_maxDepth, mx);
            while (rayCurrent < rayEnd) {
                const f3 cur = add3(camPos, scale3(rayCurrent, dir));
                float dist; uint8_t col[3];
                if (trilinear(hd, hp, cur, &dist, col)) {
                    if (lastWeight > 0 && lastSdf > 0.0f && dist < 0.0f) {
                        float alpha = 0.0f; uint8_t col2[3] = { 0, 0, 0 };
                        const int b = bisection(hd, hp, camPos, dir, lastSdf, lastAlpha, dist, rayCurrent, &alpha, col2);
                        const f3 iso = add3(camPos, scale3(alpha, dir));
                        if (b && fabsf(lastSdf - dist) < p->m_thresSampleDist && fabsf(dist) < p->m_thresDist) {
                            const float d = alpha / depthToRayLength;
                            depth[i] = d;
                            const f3 c3 = depth_to_camera(p, x, y, d);
                            depth4[4 * i] = c3.x; depth4[4 * i + 1] = c3.y; depth4[4 * i + 2] = c3.z; depth4[4 * i + 3] = 1.0f;
                            colors[4 * i] = (float)col2[0] / 255.f; colors[4 * i + 1] = (float)col2[1] / 255.f; colors[4 * i + 2] = (float)col2[2] / 255.f; colors[4 * i + 3] = 1.0f;
                            if (p->m_useGradients) {
                                const f3 g = gradient(hd, hp, iso);
                                const f3 nrm = { -g.x, -g.y, -g.z };
                                const f3 n = mul_dir(p->m_viewMatrix.m, nrm);
                                normals[4 * i] = n.x; normals[4 * i + 1] = n.y; normals[4 * i + 2] = n.z; normals[4 * i + 3] = 1.0f;
                            }
                            break;
                        }
                    }
                    lastSdf = dist; lastAlpha = rayCurrent; lastWeight = 1;
                    rayCurrent += p->m_rayIncrement;
                } else {
                    lastWeight = 0;
                    rayCurrent += p->m_rayIncrement;
                }
            }
        }
}

/* computeNormals_Kernel, FL/CUDAImageUtil.cu:404-431 over the rendered camera-space positions (CUDARayCastSDF::render when m_useGradients is off) */
ORC_API void orc_raycast_normals(const float* depth4, unsigned W, unsigned H, float* normals) {
    for (unsigned y = 0; y < H; ++y)
        for (unsigned x = 0; x < W; ++x) {
            float* o = normals + 4 * ((size_t)y * W + x);
            o[0] = o[1] = o[2] = o[3] = -INFINITY;
            if (x > 0 && x < W - 1 && y > 0 && y < H - 1) {
                const float* CC = depth4 + 4 * ((size_t)y * W + x);
                const float* PC = depth4 + 4 * ((size_t)(y + 1) * W + x); const float* CP = depth4 + 4 * ((size_t)y * W + x + 1);
                const float* MC = depth4 + 4 * ((size_t)(y - 1) * W + x); const float* CM = depth4 + 4 * ((size_t)y * W + x - 1);
                if (CC[0] != -INFINITY && PC[0] != -INFINITY && CP[0] != -INFINITY && MC[0] != -INFINITY && CM[0] != -INFINITY) {
                    const float ax = PC[0] - MC[0], ay = PC[1] - MC[1], az = PC[2] - MC[2], bx = CP[0] - CM[0], by = CP[1] - CM[1], bz = CP[2] - CM[2];
                    const float nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
                    const float l = sqrtf((nx * nx + ny * ny) + nz * nz);
                    if (l > 0.0f) { o[0] = nx / -l; o[1] = ny / -l; o[2] = nz / -l; o[3] = 0.0f; }
                }
            }
        }
}

/* ---- ray interval splat ------------------------------------------------------------------------------------------------------------ */
static inline f3 camera_to_depth_proj(const BFRayCastParams* p, f3 pos) {                  /* RayCastSDFUtil.h:213-228 */
    const float px = pos.x * p->fx / pos.z + p->mx, py = pos.y * p->fy / pos.z + p->my;
    f3 r;
    r.x = (2.0f * px - ((float)p->m_width - 1.0f)) / ((float)p->m_width - 1.0f);
    r.y = (((float)p->m_height - 1.0f) - 2.0f * py) / ((float)p->m_height - 1.0f);
    r.z = (pos.z - p->m_minDepth) / (p->m_maxDepth - p->m_minDepth);
    return r;
}
/* isSDFBlockInCameraFrustumApprox with the HASH parameters' pose and the depth camera (VoxelUtilHashSDF.h:322-326, DepthCameraUtil.h:138-144) */
static int block_in_frustum(const BFHashParams* hp, const BFDepthCameraParams* cp, i3 b) {
    const float vs = hp->m_virtualVoxelSize;
    const float off = vs * 0.5f * ((float)BF_SDF_BLOCK_SIZE - 1.0f);
    const f3 w = { (float)(b.x * BF_SDF_BLOCK_SIZE) * vs + off, (float)(b.y * BF_SDF_BLOCK_SIZE) * vs + off, (float)(b.z * BF_SDF_BLOCK_SIZE) * vs + off };
    const f3 pc = mul_point(hp->m_rigidTransformInverse.m, w);
    const float px = pc.x * cp->fx / pc.z + cp->mx, py = pc.y * cp->fy / pc.z + cp->my;
    const float w1 = (float)cp->m_imageWidth - 1.0f, h1 = (float)cp->m_imageHeight - 1.0f;
    float ix = (2.0f * px - w1) / w1, iy = (h1 - 2.0f * py) / h1;
    float iz = (pc.z - cp->m_sensorDepthWorldMin) / (cp->m_sensorDepthWorldMax - cp->m_sensorDepthWorldMin);
    ix *= 0.95f; iy *= 0.95f; iz *= 0.95f;
    return !(ix < -1.0f || ix > 1.0f || iy < -1.0f || iy > 1.0f || iz < 0.0f || iz > 1.0f);
}
/* rayIntervalSplatKernel for one compactified entry: quad[0..3] = (min x, min y, max x, max y) in normalised device coordinates, quad[4] = the quad's
 * depth in [0, 1] (nearest corner when splatMinimum, else farthest), quad[5] = that depth in metres.  Returns 0 when the block draws nothing. */
ORC_API int orc_raycast_block_quad(const BFHashParams* hp, const BFDepthCameraParams* cp, const BFRayCastParams* p, const BFHashEntry* e, float quad[6]) {
    if (e->ptr == BF_FREE_ENTRY) return 0;
    const i3 b = { e->pos[0], e->pos[1], e->pos[2] };
    if (!block_in_frustum(hp, cp, b)) return 0;
    const float vs = hp->m_virtualVoxelSize;
    const f3 wv = { (float)(b.x * BF_SDF_BLOCK_SIZE) * vs, (float)(b.y * BF_SDF_BLOCK_SIZE) * vs, (float)(b.z * BF_SDF_BLOCK_SIZE) * vs };
    const f3 mn = { wv.x - vs / 2.0f, wv.y - vs / 2.0f, wv.z - vs / 2.0f };
    const f3 mx = { mn.x + (float)BF_SDF_BLOCK_SIZE * vs, mn.y + (float)BF_SDF_BLOCK_SIZE * vs, mn.z + (float)BF_SDF_BLOCK_SIZE * vs };
    /* corner order of the reference's two reduction trees: 000 100 | 010 001 | 110 011 | 101 111 */
    const f3 c[8] = { { mn.x, mn.y, mn.z }, { mx.x, mn.y, mn.z }, { mn.x, mx.y, mn.z }, { mn.x, mn.y, mx.z }, { mx.x, mx.y, mn.z }, { mn.x, mx.y, mx.z }, { mx.x, mn.y, mx.z }, { mx.x, mx.y, mx.z } };
    f3 lo = { INFINITY, INFINITY, INFINITY }, hi = { -INFINITY, -INFINITY, -INFINITY };
    for (int k = 0; k < 8; ++k) {                    /* fminf / fmaxf are exact and order-free but for NaN, which a corner behind the camera plane (z = 0) can produce */
        const f3 q = camera_to_depth_proj(p, mul_point(p->m_viewMatrix.m, c[k]));
        lo.x = fminf(lo.x, q.x); lo.y = fminf(lo.y, q.y); lo.z = fminf(lo.z, q.z);
        hi.x = fmaxf(hi.x, q.x); hi.y = fmaxf(hi.y, q.y); hi.z = fmaxf(hi.z, q.z);
    }
    const float d = p->m_splatMinimum == 1 ? lo.z : hi.z;
    quad[0] = lo.x; quad[1] = lo.y; quad[2] = hi.x; quad[3] = hi.y; quad[4] = d;
    quad[5] = d * (p->m_maxDepth - p->m_minDepth) + p->m_minDepth;          /* depthProjToCameraZ */
    return 1;
}
/* The Direct3D 11 pass (DX11RayIntervalSplatting.cpp:137-229): two triangles per quad, constant depth, depth test LESS into the cleared-to-far "min"
 * target / GREATER into the cleared-to-near "max" target, pixel value = the quad's metric depth; default depth clip (0 <= z <= 1).  A pixel belongs to
 * the axis-aligned quad when its centre (i + 0.5, j + 0.5) lies in [x0, x1) x [y0, y1) of the viewport rectangle (top-left rule on exact coordinates).
 * out: [height][width], -inf where nothing was drawn. */
ORC_API void orc_raycast_splat(const BFHashParams* hp, const BFDepthCameraParams* cp, const BFRayCastParams* p, const BFHashEntry* compactified, unsigned numOccupied, float* out) {
    const unsigned W = p->m_width, H = p->m_height;
    for (size_t i = 0; i < (size_t)W * H; ++i) out[i] = -INFINITY;
    for (unsigned k = 0; k < numOccupied; ++k) {
        float q[6];
        if (!orc_raycast_block_quad(hp, cp, p, &compactified[k], q)) continue;
        if (!(q[4] >= 0.0f && q[4] <= 1.0f)) continue;                      /* depth clip (also drops NaN) */
        const float X0 = (q[0] + 1.0f) * 0.5f * (float)W, X1 = (q[2] + 1.0f) * 0.5f * (float)W;
        const float Y0 = (1.0f - q[3]) * 0.5f * (float)H, Y1 = (1.0f - q[1]) * 0.5f * (float)H;
        if (!(X0 < X1) || !(Y0 < Y1)) continue;
        /* first / last pixel whose centre is inside: i + 0.5 >= X0  <=>  i >= ceil(X0 - 0.5) */
        float fi0 = ceilf(X0 - 0.5f), fi1 = ceilf(X1 - 0.5f), fj0 = ceilf(Y0 - 0.5f), fj1 = ceilf(Y1 - 0.5f);
        if (fi0 < 0.0f) fi0 = 0.0f;
        if (fj0 < 0.0f) fj0 = 0.0f;
        if (fi1 > (float)W) fi1 = (float)W;
        if (fj1 > (float)H) fj1 = (float)H;
        for (int j = (int)fj0; j < (int)fj1; ++j)
            for (int i = (int)fi0; i < (int)fi1; ++i) {
                float* o = &out[(size_t)j * W + i];
                if (*o == -INFINITY) *o = q[5];
                else if (p->m_splatMinimum == 1) { if (q[5] < *o) *o = q[5]; }
                else if (q[5] > *o) *o = q[5];
            }
    }
}
