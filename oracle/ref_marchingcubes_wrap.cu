// ref_marchingcubes_wrap.cu -- TEST INFRASTRUCTURE ONLY (oracle/build_ref.py: build_marchingcubes_emulated).  Entry points around the reference's OWN iso-surface
// kernel (FL/DepthSensing/CUDAMarchingCubesSDF.cu: extractIsoSurfaceKernel through its extern "C" stubs, with MarchingCubesSDFUtil.h and Tables.h), compiled from
// the sources where they lie under /root/reference against the CPU emulation of CUDA, so that oracle/marchingcubes_oracle.c can be pinned without a GPU.
// Nothing of the reference is copied: the .cu files are #included from the scratch tree build_ref.py prepares (launch syntax and the HashEntry alignment patched there).
#include "CUDAConstant.cu"
#include "CUDAMarchingCubesSDF.cu"

extern "C" int ref_marchingcubes_extract(void* hash, void* sdfBlocks, const HashParams* hp, MarchingCubesParams* p, void* triangles, unsigned* numTriangles) {
    HashDataStruct hd;
    hd.d_hash = (HashEntry*)hash; hd.d_SDFBlocks = (Voxel*)sdfBlocks;
    updateConstantHashParams(*hp);
    MarchingCubesData data;
    data.d_params = p;                                   // "device" memory is host memory under the emulation (MarchingCubesData::updateParams)
    data.d_triangles = (MarchingCubesData::Triangle*)triangles; data.d_numTriangles = numTriangles;
    RayCastData rc;
    resetMarchingCubesCUDA(data);
    extractIsoSurfaceCUDA(hd, rc, *p, data);
    return 0;
}
extern "C" int ref_marchingcubes_tables(int* edge, int* tri) {
    for (int c = 0; c < 256; ++c) { edge[c] = edgeTable[c]; for (int i = 0; i < 16; ++i) tri[16 * c + i] = triTable[c][i]; }
    return 0;
}
extern "C" int ref_marchingcubes_sizes(int* out) {
    out[0] = (int)sizeof(HashEntry); out[1] = (int)sizeof(HashParams); out[2] = (int)sizeof(MarchingCubesParams); out[3] = (int)sizeof(Voxel); out[4] = (int)sizeof(MarchingCubesData::Triangle);
    return 0;
}
