// ref_raycast_wrap.cu -- TEST INFRASTRUCTURE ONLY (oracle/build_ref.py: build_raycast_emulated).  Entry points around the reference's OWN ray-cast
// kernels (FL/DepthSensing/CUDARayCastSDF.cu: renderKernel, rayIntervalSplatKernel, through its extern "C" stubs renderCS / rayIntervalSplatCUDA),
// compiled from their sources where they lie under /root/reference against the CPU emulation of CUDA, so that the restatement in
// oracle/raycast_oracle.c can be pinned without a GPU and without Direct3D.  Nothing of the reference is copied: the two .cu files are #included from the
// scratch tree build_ref.py prepares (launch syntax, texture binding over linear memory and the HashEntry alignment patched there).
#include "CUDAConstant.cu"
#include "CUDARayCastSDF.cu"

extern "C" int ref_raycast_render(void* hash, void* sdfBlocks, const HashParams* hp, const RayCastParams* p, float* rayMin, float* rayMax,
                                  float* depth, float4* depth4, float4* normals, float4* colors) {
    HashDataStruct hd;
    hd.d_hash = (HashEntry*)hash; hd.d_SDFBlocks = (Voxel*)sdfBlocks;
    updateConstantHashParams(*hp);
    updateConstantRayCastParams(*p);
    RayCastData d;
    d.d_depth = depth; d.d_depth4 = depth4; d.d_normals = normals; d.d_colors = colors;
    d.d_rayIntervalSplatMinArray = (cudaArray*)rayMin; d.d_rayIntervalSplatMaxArray = (cudaArray*)rayMax;      // bound as linear 2-D textures (patched renderCS)
    renderCS(hd, d, *p);
    return 0;
}

extern "C" int ref_raycast_quads(void* hashCompactified, const HashParams* hp, const DepthCameraParams* cp, const RayCastParams* p, float4* vertexBuffer) {
    HashDataStruct hd;
    hd.d_hashCompactified = (HashEntry*)hashCompactified;
    updateConstantHashParams(*hp);
    updateConstantDepthCameraParams(*cp);
    updateConstantRayCastParams(*p);
    RayCastData d;
    d.d_vertexBuffer = vertexBuffer;
    rayIntervalSplatCUDA(hd, d, *p);
    return 0;
}
extern "C" int ref_raycast_sizes(int* out) { out[0] = (int)sizeof(HashEntry); out[1] = (int)sizeof(HashParams); out[2] = (int)sizeof(RayCastParams); out[3] = (int)sizeof(Voxel); out[4] = (int)sizeof(DepthCameraParams); return 0; }
