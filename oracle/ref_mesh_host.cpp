// ref_mesh_host.cpp -- TEST INFRASTRUCTURE ONLY (oracle/build_ref.py: build_mesh_host).  The host half of the reference's mesh export, run by the reference's own
// classes: the body of CUDAMarchingCubesHashSDF::copyTrianglesToCPU and ::saveMesh (FL/DepthSensing/CUDAMarchingCubesHashSDF.cpp:26-46, 70-100) over mLib's
// MeshDataf (mergeCloseVertices, removeDuplicateFaces, applyTransform) and MeshIOf::saveToFile, compiled from /root/reference/external/mLib/include where it lies
// (a scratch copy with five one-line patches g++ needs; see build_mesh_host).  The statements below are those of the two member functions, applied to a
// triangle soup handed in instead of one copied from the GPU.
#include <cmath>
#include <cstring>
#include <string>
using std::isnan;                                           // meshIO.cpp calls it unqualified (MSVC's <math.h> declares it globally)

#include "mLibCore.h"

using namespace ml;

extern "C" int ref_mesh_save(const float* triangles, unsigned nTriangles, const float* transform, const char* filename,
                             float* outPositions, float* outColors, unsigned* outFaces, unsigned* counts) {
    MeshDataf m_meshData;
    {   // copyTrianglesToCPU
        unsigned int baseIdx = (unsigned int)m_meshData.m_Vertices.size();
        m_meshData.m_Vertices.resize(baseIdx + 3 * nTriangles);
        m_meshData.m_Colors.resize(baseIdx + 3 * nTriangles);
        const vec3f* vc = (const vec3f*)triangles;
        for (unsigned int i = 0; i < 3 * nTriangles; i++) {
            m_meshData.m_Vertices[baseIdx + i] = vc[2 * i + 0];
            m_meshData.m_Colors[baseIdx + i] = vec4f(vc[2 * i + 1]);
        }
    }
    // saveMesh
    m_meshData.m_FaceIndicesVertices.resize(m_meshData.m_Vertices.size());
    for (unsigned int i = 0; i < (unsigned int)m_meshData.m_Vertices.size() / 3; i++) {
        m_meshData.m_FaceIndicesVertices[i][0] = 3 * i + 0;
        m_meshData.m_FaceIndicesVertices[i][1] = 3 * i + 1;
        m_meshData.m_FaceIndicesVertices[i][2] = 3 * i + 2;
    }
    m_meshData.mergeCloseVertices(0.00001f, true);
    m_meshData.removeDuplicateFaces();
    if (transform) { mat4f t(transform); m_meshData.applyTransform(t); }
    MeshIOf::saveToFile(filename, m_meshData);
    counts[0] = (unsigned)m_meshData.m_Vertices.size(); counts[1] = (unsigned)m_meshData.m_FaceIndicesVertices.size();
    for (size_t v = 0; v < m_meshData.m_Vertices.size(); ++v) {
        memcpy(outPositions + 3 * v, &m_meshData.m_Vertices[v], 12);
        memcpy(outColors + 4 * v, &m_meshData.m_Colors[v], 16);
    }
    for (size_t f = 0; f < m_meshData.m_FaceIndicesVertices.size(); ++f) {
        if (m_meshData.m_FaceIndicesVertices[f].size() != 3) return 2;
        for (int k = 0; k < 3; ++k) outFaces[3 * f + k] = m_meshData.m_FaceIndicesVertices[f][k];
    }
    return 0;
}

// mat4f::getInverse, mLib core-math/matrix4x4.h:587-710 (the ray cast's view matrix, FL/DepthSensing/CUDARayCastSDF.cpp:92)
extern "C" void ref_mlib_mat4_inverse(const float* m16, float* out16) {
    const mat4f r = mat4f(m16).getInverse();
    memcpy(out16, r.getData(), 64);
}
