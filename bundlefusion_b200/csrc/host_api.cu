// host_api.cu -- host-side C-ABI helpers (include/bf_host.h): pose inverse and the TSDF op replay loop.
#include "../../include/bf_host.h"
#include "bf_common.cuh"

// 4x4 inverse by cofactor expansion in fp32, as the reference's float4x4::getInverse does on the host
// (FL/SiftGPU/cuda_SimpleMatrixUtil.h:980-1100).  Written from the textbook adjugate formula:
// inv = adj(M) / det(M), cofactors expanded as 2x2 sub-determinant products.
BF_API void bfMat4Inverse(const float* m, float* out) {
    // 2x2 sub-determinants of the lower two rows (s*) and upper two rows (c*)
    const float a00 = m[0], a01 = m[1], a02 = m[2], a03 = m[3];
    const float a10 = m[4], a11 = m[5], a12 = m[6], a13 = m[7];
    const float a20 = m[8], a21 = m[9], a22 = m[10], a23 = m[11];
    const float a30 = m[12], a31 = m[13], a32 = m[14], a33 = m[15];
    const float s0 = a00 * a11 - a10 * a01, s1 = a00 * a12 - a10 * a02, s2 = a00 * a13 - a10 * a03;
    const float s3 = a01 * a12 - a11 * a02, s4 = a01 * a13 - a11 * a03, s5 = a02 * a13 - a12 * a03;
    const float c5 = a22 * a33 - a32 * a23, c4 = a21 * a33 - a31 * a23, c3 = a21 * a32 - a31 * a22;
    const float c2 = a20 * a33 - a30 * a23, c1 = a20 * a32 - a30 * a22, c0 = a20 * a31 - a30 * a21;
    const float det = s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0;
    const float r = 1.0f / det;
    out[0] = (a11 * c5 - a12 * c4 + a13 * c3) * r;
    out[1] = (-a01 * c5 + a02 * c4 - a03 * c3) * r;
    out[2] = (a31 * s5 - a32 * s4 + a33 * s3) * r;
    out[3] = (-a21 * s5 + a22 * s4 - a23 * s3) * r;
    out[4] = (-a10 * c5 + a12 * c2 - a13 * c1) * r;
    out[5] = (a00 * c5 - a02 * c2 + a03 * c1) * r;
    out[6] = (-a30 * s5 + a32 * s2 - a33 * s1) * r;
    out[7] = (a20 * s5 - a22 * s2 + a23 * s1) * r;
    out[8] = (a10 * c4 - a11 * c2 + a13 * c0) * r;
    out[9] = (-a00 * c4 + a01 * c2 - a03 * c0) * r;
    out[10] = (a30 * s4 - a31 * s2 + a33 * s0) * r;
    out[11] = (-a20 * s4 + a21 * s2 - a23 * s0) * r;
    out[12] = (-a10 * c3 + a11 * c1 - a12 * c0) * r;
    out[13] = (a00 * c3 - a01 * c1 + a02 * c0) * r;
    out[14] = (-a30 * s3 + a31 * s1 - a32 * s0) * r;
    out[15] = (a20 * s3 - a21 * s1 + a22 * s0) * r;
}

namespace bf {
int tsdf_lanes_begin(const BFHashDataStruct* hd, const BFHashParams* hp);     // tsdf.cu: two-lane replay bracket
int tsdf_lanes_end(const BFHashDataStruct* hd);
}

static int run_ops(BFHashDataStruct* hd, BFHashParams* hp, const BFDepthCameraParams* cam, const BFTsdfOp* ops, int numOps,
                   const float* const* d_depthFrames, const uint8_t* const* d_colorFrames);

// One call replays a whole batch of TSDF operations.  Inside, the stencil of op k runs on the library's back lane while alloc +
// compactify of op k+1 run on the caller's stream (tsdf.cu, "two lanes"); on return the caller's stream is ordered after all of it.
BF_API int bfTsdfRunOps(BFHashDataStruct* hd, BFHashParams* hp, const BFDepthCameraParams* cam, const BFTsdfOp* ops, int numOps,
                        const float* const* d_depthFrames, const uint8_t* const* d_colorFrames) {
    int rc = bf::tsdf_lanes_begin(hd, hp);
    if (rc) return rc;
    rc = run_ops(hd, hp, cam, ops, numOps, d_depthFrames, d_colorFrames);
    const int rc2 = bf::tsdf_lanes_end(hd);
    return rc ? rc : rc2;
}

static int run_ops(BFHashDataStruct* hd, BFHashParams* hp, const BFDepthCameraParams* cam, const BFTsdfOp* ops, int numOps,
                   const float* const* d_depthFrames, const uint8_t* const* d_colorFrames) {
    static int fuse = -1;        // BF_TSDF_FUSE_REINT=0 replays every op separately (A/B measurements)
    if (fuse < 0) { const char* e = getenv("BF_TSDF_FUSE_REINT"); fuse = (e && e[0] == '0') ? 0 : 1; }
    for (int i = 0; i < numOps; ++i) {
        const BFTsdfOp& op = ops[i];
        if (fuse && op.kind == BF_TSDF_OP_DEINTEGRATE && i + 1 < numOps && ops[i + 1].kind == BF_TSDF_OP_INTEGRATE && ops[i + 1].frame == op.frame) {
            // the reference's re-integration pair: one fused pass (bfTsdfReintegrateFrame)
            BFHashParams hpOld = *hp;
            for (int k = 0; k < 16; ++k) hpOld.m_rigidTransform.m[k] = op.pose[k];
            bfMat4Inverse(op.pose, hpOld.m_rigidTransformInverse.m);
            const BFTsdfOp& nx = ops[i + 1];
            for (int k = 0; k < 16; ++k) hp->m_rigidTransform.m[k] = nx.pose[k];
            bfMat4Inverse(nx.pose, hp->m_rigidTransformInverse.m);
            BFDepthCameraData dd;
            dd.d_depthData = d_depthFrames[op.frame];
            dd.d_colorData = d_colorFrames ? d_colorFrames[op.frame] : nullptr;
            int rc = bfTsdfReintegrateFrame(hd, &hpOld, hp, &dd, cam);
            if (rc) return rc;
            ++i;
            continue;
        }
        if (op.kind == BF_TSDF_OP_GARBAGE_COLLECT) {
            int rc = bfTsdfGarbageCollect(hd, hp);
            if (rc) return rc;
            continue;
        }
        for (int k = 0; k < 16; ++k) hp->m_rigidTransform.m[k] = op.pose[k];
        bfMat4Inverse(op.pose, hp->m_rigidTransformInverse.m);
        BFDepthCameraData dd;
        dd.d_depthData = d_depthFrames[op.frame];
        dd.d_colorData = d_colorFrames ? d_colorFrames[op.frame] : nullptr;
        int rc = bfTsdfIntegrateFrame(hd, hp, &dd, cam, op.kind == BF_TSDF_OP_DEINTEGRATE ? 1 : 0);
        if (rc) return rc;
    }
    return 0;
}
