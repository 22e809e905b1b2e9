/*
 * bf_marchingcubes.h -- C-ABI of the iso-surface extraction of the hashed TSDF (second half): the triangle mesh the
 * reference saves at the end of a scan (Ctrl+9 / s_generateMeshDir), and the mesh clean-up and PLY writer behind CUDAMarchingCubesHashSDF::saveMesh.
 *
 * Reference interface this replaces (FL/ = /root/reference/FriedLiver/Source/, mLib = /root/reference/external/mLib/include):
 *   extern "C" resetMarchingCubesCUDA, extractIsoSurfaceCUDA                  FL/DepthSensing/CUDAMarchingCubesHashSDF.cpp:7-11 (kernels: CUDAMarchingCubesSDF.cu:10-52)
 *   struct MarchingCubesParams, MarchingCubesData (Vertex, Triangle)          FL/DepthSensing/MarchingCubesSDFUtil.h:9-23, 27-287
 *   MarchingCubesData::extractIsoSurfaceAtPosition, vertexInterp              FL/DepthSensing/MarchingCubesSDFUtil.h:121-242
 *   class CUDAMarchingCubesHashSDF (create, extractIsoSurface, copyTrianglesToCPU, saveMesh)   FL/DepthSensing/CUDAMarchingCubesHashSDF.{h,cpp}
 *   MeshData::mergeCloseVertices(thresh, approx = true), removeDuplicateFaces, removeDegeneratedFaces   mLib core-mesh/meshData.cpp:40-100, 200-300
 *   MeshIO::saveToPLY                                                         mLib core-mesh/meshIO.cpp:556-640
 * Out of this header's scope: the chunk-grid overload of extractIsoSurface (FL/DepthSensing/CUDAMarchingCubesHashSDF.cpp:117-160) -- it walks the
 * host-streamed chunk grid, which BundleFusion runs with streaming off (out of scope).
 *
 * The reference appends triangles with one atomicAdd each, so the ORDER of its triangle soup differs from run to run; the parity statement is on the
 * multiset of triangles (each with its three vertices in the reference's order).  Here a block's triangles are contiguous and ordered by voxel.
 */
#ifndef BF_MARCHINGCUBES_H
#define BF_MARCHINGCUBES_H

#include <stddef.h>
#include <stdint.h>

#include "bf_raycast.h"
#include "bf_tsdf.h"

#ifdef __cplusplus
extern "C" {
#endif

/* FL/DepthSensing/MarchingCubesSDFUtil.h:9-23 (64 bytes) */
typedef struct BFMarchingCubesParams {
    uint8_t  m_boxEnabled;                  /*  0  bool: only voxels whose centre lies in [m_minCorner, m_maxCorner] */
    uint8_t  m_pad[3];
    float    m_minCorner[3];                /*  4 */
    uint32_t m_maxNumTriangles;             /* 16  capacity of d_triangles */
    float    m_maxCorner[3];                /* 20 */
    uint32_t m_sdfBlockSize;                /* 32  8 */
    uint32_t m_hashNumBuckets;              /* 36 */
    uint32_t m_hashBucketSize;              /* 40 */
    float    m_threshMarchingCubes;         /* 44  largest jump between two corners of a cell (s_SDFMarchingCubeThreshFactor * voxel size) */
    float    m_threshMarchingCubes2;        /* 48  largest |sdf| of a corner */
    float    dummy[3];
} BFMarchingCubesParams;

typedef struct BFMarchingCubesVertex { float p[3]; float c[3]; } BFMarchingCubesVertex;                       /* position (world), colour / 255 */
typedef struct BFMarchingCubesTriangle { BFMarchingCubesVertex v0, v1, v2; } BFMarchingCubesTriangle;         /* 72 bytes */

/* FL/DepthSensing/MarchingCubesSDFUtil.h:281-286; device pointers, caller-owned (MarchingCubesData::allocate, :56-70) */
typedef struct BFMarchingCubesData {
    BFMarchingCubesParams*   d_params;      /* device copy of the parameters (MarchingCubesData::updateParams) */
    uint32_t*                d_numTriangles;
    BFMarchingCubesTriangle* d_triangles;   /* [m_maxNumTriangles] */
    uint8_t                  m_bIsOnGPU;
} BFMarchingCubesData;

/* ---- the reference's stubs, same names.  Errors print and exit(-1) like cutilSafeCall. ---- */
void resetMarchingCubesCUDA(BFMarchingCubesData* data);
/* reads the cell parameters from data->d_params (device) and the hash parameters last given to updateConstantHashParams, as the reference's kernel reads
 * *d_params and c_hashParams; rayCastData is not read (the reference passes it only for its member functions) and may be NULL */
void extractIsoSurfaceCUDA(const BFHashDataStruct* hashData, const BFRayCastData* rayCastData, const BFMarchingCubesParams* params, BFMarchingCubesData* data);

/* ---- bf* extension: explicit parameters, cudaError_t return codes, no host synchronisation ---- */
/* reset + extraction in one call: *d_numTriangles = min(triangles found, params->m_maxNumTriangles) when the stream reaches the end of the call */
int bfMarchingCubesExtract(const BFHashDataStruct* hashData, const BFHashParams* hashParams, const BFMarchingCubesParams* params,
                           BFMarchingCubesTriangle* d_triangles, uint32_t* d_numTriangles);

/* ---- host side: class CUDAMarchingCubesHashSDF ---- */
typedef struct BFMarchingCubes BFMarchingCubes;
int  bfMarchingCubesCreate(const BFMarchingCubesParams* params, BFMarchingCubes** out);        /* create(): device triangle buffer of m_maxNumTriangles */
void bfMarchingCubesDestroy(BFMarchingCubes* mc);
/* extractIsoSurface(hashData, hashParams, rayCastData, minCorner, maxCorner, boxEnabled) + copyTrianglesToCPU: the triangles are APPENDED to the host mesh buffer */
int  bfMarchingCubesExtractIsoSurface(BFMarchingCubes* mc, const BFHashDataStruct* hashData, const BFHashParams* hashParams, const float* minCorner, const float* maxCorner,
                                      int boxEnabled);
void bfMarchingCubesClearMeshBuffer(BFMarchingCubes* mc);
/* the host mesh buffer: triangle soup, three vertices per triangle; pointers stay valid until the next call on `mc` */
size_t bfMarchingCubesGetSoup(const BFMarchingCubes* mc, const float** positions /* [n][3] */, const float** colors /* [n][4], alpha 1 */);
/* saveMesh(filename, transform, overwriteExistingFile): index buffer, mergeCloseVertices(0.00001f, approx), removeDuplicateFaces, optional 4x4 row-major transform,
 * binary little-endian PLY; clears the mesh buffer.  With overwrite == 0 an existing file is kept and the name gets a numeric suffix, as in the reference;
 * actualPath (may be NULL) receives the name written. */
int  bfMarchingCubesSaveMesh(BFMarchingCubes* mc, const char* filename, const float* transform, int overwrite, char* actualPath, size_t actualPathCapacity);

/* ---- the mesh clean-up by itself (host memory; what saveMesh runs) ---- */
/* in: soup of numVertices vertices (positions [n][3], colours [n][4]) and numFaces index triples; out: merged vertices (in place, the first *numVerticesOut entries) and
 * faces (in place, the first *numFacesOut triples) -- mergeCloseVertices(thresh, true) then removeDuplicateFaces, in the reference's order of operations */
int  bfMeshMergeCloseVertices(float* positions, float* colors, size_t numVertices, uint32_t* faces, size_t numFaces, float thresh, size_t* numVerticesOut, size_t* numFacesOut);
int  bfMeshRemoveDuplicateFaces(uint32_t* faces, size_t numFaces, size_t* numFacesOut);
int  bfMeshSavePly(const char* filename, const float* positions, const float* colors /* [n][4] or NULL */, size_t numVertices, const uint32_t* faces, size_t numFaces);

#ifdef __cplusplus
}
#endif
#endif /* BF_MARCHINGCUBES_H */
