/*
 * bf_raycast.h -- C-ABI of the ray cast of the hashed TSDF: depth / camera-space position / normal / colour images of
 * the fused model seen from a pose.
 *
 * Reference interface this replaces (FL/ = /root/reference/FriedLiver/Source/):
 *   extern "C" renderCS, resetRayIntervalSplatCUDA, rayIntervalSplatCUDA        FL/DepthSensing/CUDARayCastSDF.cpp:10-16 (kernels: CUDARayCastSDF.cu)
 *   extern "C" updateConstantRayCastParams                                      FL/DepthSensing/CUDAConstant.cu:36
 *   struct RayCastParams                                                        FL/DepthSensing/CUDARayCastParams.h:8-27
 *   struct RayCastData                                                          FL/DepthSensing/RayCastSDFUtil.h:35-302
 *   class CUDARayCastSDF (render = interval splat + renderCS + computeNormals)  FL/DepthSensing/CUDARayCastSDF.{h,cpp}
 *   class DX11RayIntervalSplatting (the Direct3D 11 draw of the block quads)    FL/DepthSensing/DX11RayIntervalSplatting.cpp:137-229
 * The reference hands the per-block quads to Direct3D 11, which rasterises them into two depth-tested render targets that come back to CUDA as
 * cudaArray textures.  Here the interval images are plain device float images and bfRayCastSplat fills them with a CUDA kernel (same quads, same
 * depth tests, atomic min / max per pixel), so the path has no graphics API in it.  The reference-named stubs keep their names and argument order
 * (struct references are pointers at the ABI level); d_rayIntervalSplatMin/MaxArray are float* instead of cudaArray*.
 */
#ifndef BF_RAYCAST_H
#define BF_RAYCAST_H

#include <stddef.h>
#include <stdint.h>

#include "bf_tsdf.h"

#ifdef __cplusplus
extern "C" {
#endif

/* FL/DepthSensing/CUDARayCastParams.h:8-27 (192 bytes, 16-byte aligned) */
typedef struct BFRayCastParams {
    BFFloat4x4 m_viewMatrix;                /*   0  world -> camera (inverse of the rigid transform)  */
    BFFloat4x4 m_viewMatrixInverse;         /*  64  camera -> world                                   */
    float    mx, my, fx, fy;                /* 128  ray-cast intrinsics                               */
    uint32_t m_width, m_height;             /* 144 */
    uint32_t m_numOccupiedSDFBlocks;        /* 152  entries of d_hashCompactified to splat            */
    uint32_t m_maxNumVertices;
    int32_t  m_splatMinimum;                /* 160  1: quads carry their nearest depth, 0: farthest   */
    float    m_minDepth, m_maxDepth;        /* 164 */
    float    m_rayIncrement;                /* 172  s_SDFRayIncrementFactor * s_SDFTruncation         */
    float    m_thresSampleDist;             /* 176 */
    float    m_thresDist;
    uint8_t  m_useGradients;                /* 184  bool: normals from the TSDF gradient instead of the rendered positions */
    uint8_t  m_pad[3];
    uint32_t dummy0;
} BFRayCastParams;

/* FL/DepthSensing/RayCastSDFUtil.h:296-302; all device pointers, caller-owned (RayCastData::allocate, :56-61) */
typedef struct BFRayCastData {
    float* d_depth;                         /* [height][width]      ray-cast depth, -inf where no surface                    */
    float* d_depth4;                        /* [height][width][4]   camera-space position (x, y, z, 1)                        */
    float* d_normals;                       /* [height][width][4]   camera-space normal                                      */
    float* d_colors;                        /* [height][width][4]   colour / 255, alpha 1                                    */
    float* d_vertexBuffer;                  /* [6 * numOccupied][4] the quads of rayIntervalSplatCUDA (may be NULL for bfRayCastSplat) */
    float* d_rayIntervalSplatMin;           /* [height][width]      nearest block depth along the pixel's ray, -inf = none    */
    float* d_rayIntervalSplatMax;           /* [height][width]      farthest                                                  */
} BFRayCastData;

/* ---- the reference's stubs, same names.  Errors print and exit(-1) like cutilSafeCall. ---- */
void updateConstantRayCastParams(const BFRayCastParams* params);
/* six vertices (two triangles) per compactified entry, (ndc x, ndc y, depth in [0, 1], depth in metres); uses the hash / depth-camera parameters last
 * given to updateConstantHashParams / updateConstantDepthCameraParams for the frustum test, as the reference's kernel uses its __constant__ copies */
void rayIntervalSplatCUDA(const BFHashDataStruct* hashData, const BFRayCastData* rayCastData, const BFRayCastParams* rayCastParams);
void resetRayIntervalSplatCUDA(BFRayCastData* data, const BFRayCastParams* params);
/* the ray march: reads d_rayIntervalSplatMin / Max, writes d_depth, d_depth4, d_colors and (m_useGradients) d_normals */
void renderCS(const BFHashDataStruct* hashData, const BFRayCastData* rayCastData, const BFRayCastParams* rayCastParams);

/* ---- bf* extension: explicit parameters, cudaError_t return codes, no host synchronisation ---- */
/* Interval images for both directions in one call: what DX11RayIntervalSplatting::rayIntervalSplatting produces with two draws.  The number of
 * entries is read on the device from hashData->d_hashCompactifiedCounter (the last compactify), rayCastParams->m_numOccupiedSDFBlocks is ignored. */
int bfRayCastSplat(const BFHashDataStruct* hashData, const BFHashParams* hashParams, const BFDepthCameraParams* cameraParams, const BFRayCastData* rayCastData,
                   const BFRayCastParams* rayCastParams);
int bfRayCastRender(const BFHashDataStruct* hashData, const BFHashParams* hashParams, const BFRayCastData* rayCastData, const BFRayCastParams* rayCastParams);
/* computeNormals (FL/CUDAImageUtil.cu:404-445) over d_depth4 into d_normals: what CUDARayCastSDF::render runs when m_useGradients is off */
int bfRayCastComputeNormals(const BFRayCastData* rayCastData, unsigned int width, unsigned int height);
/* CUDARayCastSDF::render (FL/DepthSensing/CUDARayCastSDF.cpp:42-73) for the pose `rigidTransform` (camera -> world, 4x4 row-major, host): fills the view
 * matrices of *rayCastParams, then splat, render, normals -- four launches, nothing comes back to the host */
int bfRayCastRenderPose(const BFHashDataStruct* hashData, const BFHashParams* hashParams, const BFDepthCameraParams* cameraParams, const BFRayCastData* rayCastData,
                        BFRayCastParams* rayCastParams, const float* rigidTransform);

#ifdef __cplusplus
}
#endif
#endif /* BF_RAYCAST_H */
