// ref_sensordata_host.cpp -- TEST INFRASTRUCTURE ONLY (oracle/build_ref.py: build_sensordata_host).  The reference's own `.sens` container class, ml::SensorData
// (external/mLib/include/ext-depthcamera/sensorData.h: initDefault / addFrame / saveToFile :832-866, 1040-1048; loadFromFile :1187-1227; decompressColorAlloc /
// decompressDepthAlloc :543-668), compiled by g++ from where it lies (scratch copy of mLib's core headers with the patches of build_mesh_host) so that
// include/bf_sens.h's reader and writer can be pinned against files the reference writes and against what the reference reads.  Colour can only be WRITTEN raw here:
// the reference compresses JPEG / PNG through the Windows-only uplink codec (sensorData.h:508-523); it READS both through stb_image.
#include <cmath>
#include <cstring>
#include <string>
using std::isnan;
namespace std { using ::ceilf; using ::log10f; }           // sensorData.h:1257 spells them std::ceilf / std::log10f (MSVC's <cmath> has them)

#include "mLibCore.h"
#include "mLibDepthCamera.h"

using namespace ml;

extern "C" int ref_sensordata_write(const char* path, unsigned w, unsigned h, const float* intrinsic, float depthShift, int depthType, const char* name, unsigned n,
                                    const unsigned char* rgb, const unsigned short* depth, const float* poses, const unsigned long long* timeStamps) {
    try {
        SensorData sd;
        const SensorData::CalibrationData calib{ mat4f(intrinsic) };
        sd.initDefault(w, h, w, h, calib, calib, SensorData::TYPE_RAW, (SensorData::COMPRESSION_TYPE_DEPTH)depthType, depthShift, name);
        for (unsigned i = 0; i < n; ++i)
            sd.addFrame((const vec3uc*)(rgb + (size_t)3 * w * h * i), depth + (size_t)w * h * i, mat4f(poses + 16 * i), timeStamps[2 * i], timeStamps[2 * i + 1]);
        sd.saveToFile(path);
    } catch (const std::exception& e) { fprintf(stderr, "ref_sensordata_write: %s\n", e.what()); return 1; }
    return 0;
}

// dims: colorWidth, colorHeight, depthWidth, depthHeight, colorCompression, depthCompression, numFrames, numIMUFrames; calib: colour intrinsic, colour extrinsic, depth intrinsic, depth extrinsic
extern "C" int ref_sensordata_read(const char* path, unsigned* dims, float* calib, float* depthShift, char* name, unsigned nameCap, unsigned capFrames,
                                   unsigned char* rgb, unsigned short* depth, float* poses, unsigned long long* timeStamps) {
    try {
        SensorData sd;
        sd.loadFromFile(path);
        dims[0] = sd.m_colorWidth; dims[1] = sd.m_colorHeight; dims[2] = sd.m_depthWidth; dims[3] = sd.m_depthHeight;
        dims[4] = (unsigned)sd.m_colorCompressionType; dims[5] = (unsigned)sd.m_depthCompressionType; dims[6] = (unsigned)sd.m_frames.size(); dims[7] = (unsigned)sd.m_IMUFrames.size();
        memcpy(calib, sd.m_calibrationColor.m_intrinsic.getData(), 64); memcpy(calib + 16, sd.m_calibrationColor.m_extrinsic.getData(), 64);
        memcpy(calib + 32, sd.m_calibrationDepth.m_intrinsic.getData(), 64); memcpy(calib + 48, sd.m_calibrationDepth.m_extrinsic.getData(), 64);
        *depthShift = sd.m_depthShift;
        strncpy(name, sd.m_sensorName.c_str(), nameCap - 1); name[nameCap - 1] = 0;
        for (unsigned i = 0; i < sd.m_frames.size() && i < capFrames; ++i) {
            vec3uc* c = sd.decompressColorAlloc(i);
            unsigned short* d = sd.decompressDepthAlloc(i);
            memcpy(rgb + (size_t)3 * dims[0] * dims[1] * i, c, (size_t)3 * dims[0] * dims[1]);
            memcpy(depth + (size_t)dims[2] * dims[3] * i, d, (size_t)2 * dims[2] * dims[3]);
            std::free(c); std::free(d);
            memcpy(poses + 16 * i, sd.m_frames[i].getCameraToWorld().getData(), 64);
            timeStamps[2 * i] = sd.m_frames[i].getTimeStampColor(); timeStamps[2 * i + 1] = sd.m_frames[i].getTimeStampDepth();
        }
    } catch (const std::exception& e) { fprintf(stderr, "ref_sensordata_read: %s\n", e.what()); return 1; }
    return 0;
}
