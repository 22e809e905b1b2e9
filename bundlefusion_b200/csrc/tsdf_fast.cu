// tsdf_fast.cu -- the integrate / de-integrate / fused re-integration stencil in TOLERANCE arithmetic (the library default).
//
// Behavioural source: integrateDepthMapKernel<deIntegrate>, FL/DepthSensing/CUDASceneRepHashSDF.cu:420-521 (per-voxel rule),
// DepthCameraUtil.h:71-82 (projection), the re-integration loop of FL/DepthSensing/DepthSensing.cpp:854-902.
//
// Why a second stencil.  tsdf.cu reproduces the reference's IEEE build bit for bit (individually rounded operations, two IEEE
// divides per voxel probe, the full 4x4 product per voxel): ~130 thread-instructions per probe, issue-bound at 18 % of the HBM
// roof on the re-integration stream (profiles/r1_ncu_full_reintegrate_alloc.txt).  The reference itself ships --use_fast_math
// and BASELINE's north star asks for "weights/sdf within a stated fp32 tolerance", so the product path is this one:
//   * the voxel -> camera transform is affine in the integer voxel coordinate: one FMA chain per (thread, block) and three adds
//     per further x-consecutive voxel instead of a 4x4 product per voxel;
//   * one approximate reciprocal (MUFU.RCP) of z per probe instead of two IEEE divides;
//   * the truncation clamp is dropped (|sdf| < trunc makes it the identity), the colour blend 0.2 c + 0.8 o is evaluated in
//     exact integer arithmetic (its fraction is a multiple of 0.2, never a rounding tie: identical to the float expression),
//     byte <-> float moves use PRMT / magic-number adds, no conversion instructions;
//   * a fused re-integration shares everything but the two poses' projections; each voxel is read and written once.
// Contract (tests/test_tsdf_fast_gpu.py): allocated block set identical (alloc / compactify stay in tsdf.cu), weights identical,
// |d sdf| <= 1e-5 m, colour +-1, except for a <= 1e-4 fraction of voxels whose projected pixel or truncation test sits within
// rounding distance of its decision boundary -- the same kind and size of difference the reference's own --use_fast_math build
// shows against its IEEE build (profiles/r1_tsdf_parity_vs_reference_cuda.txt).  bfTsdfSetArithmetic(BF_TSDF_ARITH_EXACT)
// selects tsdf.cu's kernels.
#include "bf_common.cuh"
#include "tsdf_shared.cuh"

namespace bf {

struct FastPose {
    float cx[4], cy[4], cz[4];       // camera-space x, y, z of a voxel = c[0] vx + c[1] vy + c[2] vz + c[3]   (rows of Tinv * diag(voxelSize))
    float maxDist, trunc0, truncScale, wMax;
};
struct FastArgs {
    BFVoxel* blocks; const BFHashEntry* list; int* listCounterOut; const int4* work;
    const float* depth; const uchar4* color; unsigned* ctrs; int* live;
    unsigned W, H; int set; int useListCount; unsigned countOverride;
    float fx, fy, mx5, my5;          // mx + 0.5, my + 0.5: the reference's int(proj + 0.5)
    FastPose A, B;                   // MODE 0 / 1: A.  MODE 2: A = old pose (de-integrated), B = new pose (integrated)
};

#ifndef BF_FAST_MINBLOCKS
#define BF_FAST_MINBLOCKS 10
#endif

#ifdef BF_EMU_SEQUENTIAL      // tests/test_tsdf_fast_emulated.py: this source executed on the CPU, one CUDA thread after the other
__device__ __forceinline__ float rcp_approx(float x) { return 1.0f / x; }
#else
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#endif
// byte k of w as a float, exactly: bits 0x4B0000bb = 2^23 + b
template <int K> __device__ __forceinline__ float byte_to_float(unsigned w) {
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440u | K)) - 8388608.0f;
}
template <int K> __device__ __forceinline__ unsigned put_byte(unsigned acc, unsigned v) {
    return __byte_perm(acc, v, K == 0 ? 0x3214u : (K == 1 ? 0x3240u : 0x3410u));
}

struct Probe { float sdf; unsigned col; };

// truncation test of one voxel at camera-space (X, Y, Z); .cu:433-463 without the (identity) clamp
__device__ __forceinline__ bool probe_fast(const FastArgs& a, const FastPose& p, float X, float Y, float Z, Probe& out) {
    const float rz = rcp_approx(Z);
    const float sx = fmaf(X * rz, a.fx, a.mx5), sy = fmaf(Y * rz, a.fy, a.my5);
    const unsigned ix = (unsigned)__float2int_rz(sx), iy = (unsigned)__float2int_rz(sy);       // cvt.rzi: (-1, 0) -> 0, NaN -> 0, as the reference's (int)
    if (ix >= a.W || iy >= a.H) return false;
    const unsigned idx = iy * a.W + ix;
    const float d = __ldg(&a.depth[idx]);
    if (!(d != -INFINITY && d < p.maxDist)) return false;
    const float sdf = d - Z;
    if (!(fabsf(sdf) < fmaf(p.truncScale, d, p.trunc0))) return false;
    out.sdf = sdf;
    out.col = __ldg(reinterpret_cast<const unsigned*>(a.color) + idx);
    return true;
}

// integrate one sample into (sdf, weight, colour) words; .cu:486-500
__device__ __forceinline__ void integrate_fast(const FastPose& p, const Probe& s, unsigned& wSdf, unsigned& wWeight, unsigned& wColor) {
    const float oldSdf = __uint_as_float(wSdf), oldW = __uint_as_float(wWeight);
    const float den = oldW + 1.0f;
    const float nSdf = fmaf(oldSdf, oldW, s.sdf) * rcp_approx(den);
    unsigned nc = 0xFF000000u;
    if (oldW == 0.0f) {
        // colour = the sample, through clamp [0, 254.5] and the byte conversion: 255 -> 254
        const unsigned c0 = s.col & 0xffu, c1 = (s.col >> 8) & 0xffu, c2 = (s.col >> 16) & 0xffu;
        nc = put_byte<0>(nc, min(c0, 254u)); nc = put_byte<1>(nc, min(c1, 254u)); nc = put_byte<2>(nc, min(c2, 254u));
    } else {
        // round(0.2 c + 0.8 o) = floor((2 c + 8 o + 5) / 10): the fraction of (c + 4 o) / 5 is a multiple of 0.2, never a tie
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const unsigned c = (s.col >> (8 * k)) & 0xffu, o = (wColor >> (8 * k)) & 0xffu;
            const unsigned q = __umulhi(8u * o + 2u * c + 5u, 429496730u);
            const unsigned v = min(q, 254u);
            nc = (k == 0) ? put_byte<0>(nc, v) : (k == 1 ? put_byte<1>(nc, v) : put_byte<2>(nc, v));
        }
    }
    wSdf = __float_as_uint(nSdf); wWeight = __float_as_uint(fminf(p.wMax, den)); wColor = nc;
}

// de-integrate one sample; .cu:501-514
__device__ __forceinline__ void deintegrate_fast(const Probe& s, unsigned& wSdf, unsigned& wWeight, unsigned& wColor) {
    const float oldSdf = __uint_as_float(wSdf), oldW = __uint_as_float(wWeight);
    const float den = oldW - 1.0f;
    if (!(den > 0.001f)) { wSdf = 0u; wWeight = 0u; wColor = 0u; return; }           // weight max(0, w - 1) <= 0.001: the voxel is cleared
    const float r = rcp_approx(den);
    const float nSdf = fmaf(oldSdf, oldW, -s.sdf) * r;
    // round((o w - c) / (w - 1)), half away from zero, clamped to [0, 254.5] -> byte.  The quotient is biased up by 2^-20 so that the
    // exact ties (w - 1 = 2, 4, ...) fall on the reference's side under an approximate reciprocal; then round-to-nearest via 2^23.
    const float rb = r * 1.00000095367431640625f;
    unsigned nc = 0xFF000000u;
    {
        const float q = fminf(fmaxf(fmaf(byte_to_float<0>(wColor), oldW, -byte_to_float<0>(s.col)) * rb, 0.0f), 254.4f);
        nc = put_byte<0>(nc, __float_as_uint(q + 8388608.0f));
    }
    {
        const float q = fminf(fmaxf(fmaf(byte_to_float<1>(wColor), oldW, -byte_to_float<1>(s.col)) * rb, 0.0f), 254.4f);
        nc = put_byte<1>(nc, __float_as_uint(q + 8388608.0f));
    }
    {
        const float q = fminf(fmaxf(fmaf(byte_to_float<2>(wColor), oldW, -byte_to_float<2>(s.col)) * rb, 0.0f), 254.4f);
        nc = put_byte<2>(nc, __float_as_uint(q + 8388608.0f));
    }
    wSdf = __float_as_uint(nSdf); wWeight = __float_as_uint(den); wColor = nc;
}

template <int MODE>
__device__ __forceinline__ void update_fast(const FastArgs& a, bool passA, const Probe& sA, bool passB, const Probe& sB,
                                            unsigned& wSdf, unsigned& wWeight, unsigned& wColor, int& liveDelta) {
    const bool wasLive = __uint_as_float(wWeight) > 0.0f;
    if (MODE == 0) { integrate_fast(a.A, sA, wSdf, wWeight, wColor); }
    else if (MODE == 1) { deintegrate_fast(sA, wSdf, wWeight, wColor); }
    else {
        if (passA) deintegrate_fast(sA, wSdf, wWeight, wColor);
        if (passB) integrate_fast(a.B, sB, wSdf, wWeight, wColor);
    }
    liveDelta += (int)(__uint_as_float(wWeight) > 0.0f) - (int)wasLive;
}

// 128 threads per SDF block, 4 x-consecutive voxels (48 B = three 16-byte vectors) per thread; persistent grid over the work list.
template <int MODE>
__global__ void __launch_bounds__(128, BF_FAST_MINBLOCKS)
stencil_fast_kernel(const __grid_constant__ FastArgs a) {
    const unsigned listCount = a.useListCount ? a.ctrs[a.set + SET_COUNT] : a.countOverride;
    const unsigned count = a.work ? a.ctrs[a.set + SET_WORK] : listCount;
    const unsigned t = threadIdx.x;
    if (blockIdx.x == 0 && t == 0) {
        if (a.useListCount) { a.listCounterOut[0] = (int)listCount; a.ctrs[CTR_E] = listCount; }
        if (!a.work) atomicAdd(reinterpret_cast<unsigned long long*>(&a.ctrs[CTR_E_TOT_LO]), (unsigned long long)listCount);
    }
    // this thread's first voxel inside a block: i = 4t -> x = (4t) % 8, y = (4t % 64) / 8, z = 4t / 64
    const float flx = (float)((4 * t) & 7), fly = (float)(((4 * t) & 63) >> 3), flz = (float)((4 * t) >> 6);
    const float oAx = fmaf(flx, a.A.cx[0], fmaf(fly, a.A.cx[1], fmaf(flz, a.A.cx[2], a.A.cx[3])));
    const float oAy = fmaf(flx, a.A.cy[0], fmaf(fly, a.A.cy[1], fmaf(flz, a.A.cy[2], a.A.cy[3])));
    const float oAz = fmaf(flx, a.A.cz[0], fmaf(fly, a.A.cz[1], fmaf(flz, a.A.cz[2], a.A.cz[3])));
    float oBx = 0.0f, oBy = 0.0f, oBz = 0.0f;
    if (MODE == 2) {
        oBx = fmaf(flx, a.B.cx[0], fmaf(fly, a.B.cx[1], fmaf(flz, a.B.cx[2], a.B.cx[3])));
        oBy = fmaf(flx, a.B.cy[0], fmaf(fly, a.B.cy[1], fmaf(flz, a.B.cy[2], a.B.cy[3])));
        oBz = fmaf(flx, a.B.cz[0], fmaf(fly, a.B.cz[1], fmaf(flz, a.B.cz[2], a.B.cz[3])));
    }
    unsigned passed = 0;
    int4 wNext = make_int4(0, 0, 0, 0);
    if (a.work && blockIdx.x < count) wNext = __ldg(&a.work[blockIdx.x]);

    for (unsigned b = blockIdx.x; b < count; b += gridDim.x) {
        int bx, by, bz; unsigned ptr, fl = 3u;
        if (a.work) {
            const int4 w = wNext;
            if (b + gridDim.x < count) wNext = __ldg(&a.work[b + gridDim.x]);
            bx = w.x; by = w.y; bz = w.z; ptr = ((unsigned)w.w & 0x0FFFFFFFu) * BF_SDF_BLOCK_VOXELS;
            if (MODE == 2) fl = (unsigned)w.w >> 28;          // bit0: the old pose can touch this block, bit1: the new pose can
        } else {
            const BFHashEntry* ep = &a.list[b];
            bx = __ldg(&ep->pos[0]); by = __ldg(&ep->pos[1]); bz = __ldg(&ep->pos[2]); ptr = (unsigned)__ldg(&ep->ptr);
        }
        const float fbx = (float)(bx * BF_SDF_BLOCK_SIZE), fby = (float)(by * BF_SDF_BLOCK_SIZE), fbz = (float)(bz * BF_SDF_BLOCK_SIZE);
        Probe sA[4], sB[4];
        unsigned maskA = 0, maskB = 0;
        if (a.color != nullptr) {                             // without colour nothing passes (.cu:441-448)
            if (fl & 1u) {
                const float X = fmaf(fbx, a.A.cx[0], fmaf(fby, a.A.cx[1], fmaf(fbz, a.A.cx[2], oAx)));
                const float Y = fmaf(fbx, a.A.cy[0], fmaf(fby, a.A.cy[1], fmaf(fbz, a.A.cy[2], oAy)));
                const float Z = fmaf(fbx, a.A.cz[0], fmaf(fby, a.A.cz[1], fmaf(fbz, a.A.cz[2], oAz)));
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (probe_fast(a, a.A, fmaf((float)k, a.A.cx[0], X), fmaf((float)k, a.A.cy[0], Y), fmaf((float)k, a.A.cz[0], Z), sA[k])) maskA |= 1u << k;
            }
            if (MODE == 2 && (fl & 2u)) {
                const float X = fmaf(fbx, a.B.cx[0], fmaf(fby, a.B.cx[1], fmaf(fbz, a.B.cx[2], oBx)));
                const float Y = fmaf(fbx, a.B.cy[0], fmaf(fby, a.B.cy[1], fmaf(fbz, a.B.cy[2], oBy)));
                const float Z = fmaf(fbx, a.B.cz[0], fmaf(fby, a.B.cz[1], fmaf(fbz, a.B.cz[2], oBz)));
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (probe_fast(a, a.B, fmaf((float)k, a.B.cx[0], X), fmaf((float)k, a.B.cy[0], Y), fmaf((float)k, a.B.cz[0], Z), sB[k])) maskB |= 1u << k;
            }
        }
        const unsigned mask = maskA | maskB;
        int liveDelta = 0;
        if (mask) {
            uint4* const vp = reinterpret_cast<uint4*>(a.blocks + (size_t)ptr) + 3 * t;      // 48 B per thread, 16-B aligned
            uint4 qa = vp[0], qb = vp[1], qc = vp[2];
            // 4 voxels = 12 words: v0{a.x,a.y,a.z} v1{a.w,b.x,b.y} v2{b.z,b.w,c.x} v3{c.y,c.z,c.w}
            if (mask & 1u) update_fast<MODE>(a, maskA & 1u, sA[0], maskB & 1u, sB[0], qa.x, qa.y, qa.z, liveDelta);
            if (mask & 2u) update_fast<MODE>(a, maskA & 2u, sA[1], maskB & 2u, sB[1], qa.w, qb.x, qb.y, liveDelta);
            if (mask & 4u) update_fast<MODE>(a, maskA & 4u, sA[2], maskB & 4u, sB[2], qb.z, qb.w, qc.x, liveDelta);
            if (mask & 8u) update_fast<MODE>(a, maskA & 8u, sA[3], maskB & 8u, sB[3], qc.y, qc.z, qc.w, liveDelta);
            if (mask & 0x3u) vp[0] = qa;                      // only the 16-byte pieces that hold an updated voxel
            if (mask & 0x6u) vp[1] = qb;
            if (mask & 0xCu) vp[2] = qc;
            passed += __popc(maskA) + __popc(maskB);
        }
        // live-voxel bookkeeping for the O(E) garbage collection: one RED per warp, only when a weight crossed zero
#ifdef BF_EMU_SEQUENTIAL
        if (liveDelta != 0) atomicAdd(&a.live[ptr / BF_SDF_BLOCK_VOXELS], liveDelta);
#else
        if (__any_sync(0xffffffffu, liveDelta != 0)) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) liveDelta += __shfl_xor_sync(0xffffffffu, liveDelta, o);
            if ((t & 31) == 0 && liveDelta != 0) atomicAdd(&a.live[ptr / BF_SDF_BLOCK_VOXELS], liveDelta);
        }
#endif
    }
    // U statistics (voxel updates, the roofline's byte count): one 64-bit atomic per CTA
#ifdef BF_EMU_SEQUENTIAL
    if (passed) { atomicAdd(reinterpret_cast<unsigned long long*>(&a.ctrs[a.set + SET_U_LO]), (unsigned long long)passed); atomicAdd(reinterpret_cast<unsigned long long*>(&a.ctrs[CTR_U_TOT_LO]), (unsigned long long)passed); }
    return;
#endif
    passed = warp_sum_u(passed);
    __shared__ unsigned sPassed[4];
    if ((t & 31) == 0) sPassed[t >> 5] = passed;
    __syncthreads();
    if (t == 0) {
        const unsigned long long tot = (unsigned long long)sPassed[0] + sPassed[1] + sPassed[2] + sPassed[3];
        if (tot) { atomicAdd(reinterpret_cast<unsigned long long*>(&a.ctrs[a.set + SET_U_LO]), tot); atomicAdd(reinterpret_cast<unsigned long long*>(&a.ctrs[CTR_U_TOT_LO]), tot); }
    }
}

// ---- host side ----------------------------------------------------------------------------------------------------------
static void make_pose(const BFHashParams* hp, FastPose* p) {
    const float* M = hp->m_rigidTransformInverse.m;
    const double vs = (double)hp->m_virtualVoxelSize;
    for (int c = 0; c < 3; ++c) {
        p->cx[c] = (float)((double)M[0 + c] * vs); p->cy[c] = (float)((double)M[4 + c] * vs); p->cz[c] = (float)((double)M[8 + c] * vs);
    }
    p->cx[3] = M[3]; p->cy[3] = M[7]; p->cz[3] = M[11];
    p->maxDist = hp->m_maxIntegrationDistance; p->trunc0 = hp->m_truncation; p->truncScale = hp->m_truncScale; p->wMax = (float)hp->m_integrationWeightMax;
}
static void make_args(FastArgs* a, const BFHashDataStruct* hd, const BFDepthCameraParams* cp, const float* depth, const void* color,
                      bool useListCount, unsigned countOverride, unsigned* ctrs, int* live, const int4* work, int set) {
    a->blocks = hd->d_SDFBlocks; a->list = hd->d_hashCompactified; a->listCounterOut = hd->d_hashCompactifiedCounter; a->work = work;
    a->depth = depth; a->color = reinterpret_cast<const uchar4*>(color); a->ctrs = ctrs; a->live = live;
    a->W = cp->m_imageWidth; a->H = cp->m_imageHeight; a->set = set; a->useListCount = useListCount ? 1 : 0; a->countOverride = countOverride;
    a->fx = cp->fx; a->fy = cp->fy; a->mx5 = cp->mx + 0.5f; a->my5 = cp->my + 0.5f;
}

int fast_stencil_ctas_per_sm() { return BF_FAST_MINBLOCKS; }

int launch_integrate_fast(const BFHashDataStruct* hd, const BFHashParams* hp, const BFDepthCameraParams* cp, const float* depth, const void* color,
                          bool deIntegrate, bool useListCount, unsigned countOverride, unsigned* ctrs, int* live, const int4* work, int set,
                          int grid, cudaStream_t s) {
    FastArgs a;
    make_args(&a, hd, cp, depth, color, useListCount, countOverride, ctrs, live, work, set);
    make_pose(hp, &a.A); a.B = a.A;
    if (deIntegrate) stencil_fast_kernel<1><<<grid, 128, 0, s>>>(a);
    else             stencil_fast_kernel<0><<<grid, 128, 0, s>>>(a);
    BF_CHECK(cudaGetLastError());
    return 0;
}

int launch_reintegrate_fast(const BFHashDataStruct* hd, const BFHashParams* hpOld, const BFHashParams* hpNew, const BFDepthCameraParams* cp,
                            const float* depth, const void* color, const int4* work, int set, unsigned* ctrs, int* live, int grid, cudaStream_t s) {
    FastArgs a;
    make_args(&a, hd, cp, depth, color, true, 0, ctrs, live, work, set);
    make_pose(hpOld, &a.A); make_pose(hpNew, &a.B);
    stencil_fast_kernel<2><<<grid, 128, 0, s>>>(a);
    BF_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace bf
