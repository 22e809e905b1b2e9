/*
 * bf_sens.h -- C-ABI of the `.sens` reader / writer (first half): the recorded-sequence container the reference's
 * SensorDataReader plays back, so that real sequences (BundleFusion / ScanNet recordings) can drive the frame loop.
 *
 * Reference interface this replaces (external/mLib = /root/reference/external/mLib/include):
 *   ml::SensorData::loadFromFile / saveToFile, RGBDFrame::loadFromFile / saveToFile      ext-depthcamera/sensorData.h:676-700, 1040-1048, 1187-1227   (on-disk layout, version 4)
 *   RGBDFrame::decompressDepthAlloc / decompressColorAlloc                                ext-depthcamera/sensorData.h:540-600, 640-668
 *   SensorDataReader::processDepth (ushort -> metres, 0 -> -inf; RGB -> RGBX)             FL/SensorDataReader.cpp:100-117
 * mLib decodes through a vendored stb_image v2.08 (JPEG / PNG / zlib).  Here: zlib from the system library, PNG (what the reference's decoder reads: 8 bits per channel for every colour type, 1 / 2 / 4 bits for grey and palette images,
 * Adam7 interlacing; not 16 bits) and
 * baseline and progressive JPEG (Huffman; up to 2x2 chroma subsampling, restart markers) decoded by this library's own code; arithmetic-coded / lossless JPEG
 * (which the reference's decoder refuses too) and OCCI depth are reported as unsupported.  JPEG decoding is not normative in its last bit (IDCT, chroma up-sampling, colour conversion) and SIFT sees that bit:
 * the decoder restates the reference decoder's fixed-point pipeline and is bit-identical to it (tests/test_sens_reference_stb.py: golden outputs of the
 * reference's stb_image compiled from /root/reference; libjpeg's output is compared under a tolerance of 3 levels besides).  The container: files written by the
 * reference's ml::SensorData are read field for field, the writer's raw-depth file equals the reference's byte for byte (tests/test_sens_reference_sensordata.py).
 *
 * Host-only code: no CUDA call is made by this header's functions; buffers are host memory (pin them to hand them to bfFrameLoopStep with onHost = 1).
 */
#ifndef BF_SENS_H
#define BF_SENS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { BF_SENS_COLOR_RAW = 0, BF_SENS_COLOR_PNG = 1, BF_SENS_COLOR_JPEG = 2 };                  /* COMPRESSION_TYPE_COLOR, sensorData.h:289-294 */
enum { BF_SENS_DEPTH_RAW_USHORT = 0, BF_SENS_DEPTH_ZLIB_USHORT = 1, BF_SENS_DEPTH_OCCI_USHORT = 2 };   /* COMPRESSION_TYPE_DEPTH, :295-300 */
enum { BF_SENS_OK = 0, BF_SENS_ERR_IO = 1, BF_SENS_ERR_FORMAT = 2, BF_SENS_ERR_UNSUPPORTED = 3, BF_SENS_ERR_RANGE = 4, BF_SENS_ERR_ARGUMENT = 5 };

typedef struct BFSensHeader {
    uint32_t version;                       /* 4 (M_SENSOR_DATA_VERSION) */
    char     sensorName[256];               /* zero-terminated (longer names are cut) */
    float    colorIntrinsic[16], colorExtrinsic[16];        /* CalibrationData, 4x4 row-major */
    float    depthIntrinsic[16], depthExtrinsic[16];
    int32_t  colorCompression, depthCompression;
    uint32_t colorWidth, colorHeight, depthWidth, depthHeight;
    float    depthShift;                    /* depth in metres = ushort / depthShift (1000 for millimetres) */
    uint64_t numFrames, numIMUFrames;
} BFSensHeader;

typedef struct BFSensReader BFSensReader;
/* opens the file, reads the header and indexes the frames (one pass over the frame sizes; pixel data is read on demand) */
int bfSensOpen(const char* path, BFSensReader** out, BFSensHeader* header);
/* Frame `index` as the frame loop takes it: depth float [depthHeight][depthWidth] in metres with -inf where the sensor had none (ushort 0), colour uchar4
 * [colorHeight][colorWidth] = (r, g, b, 1); cameraToWorld: the recorded 4x4 pose (all -inf when the recording has none); timeStamps[2] = colour, depth.
 * Any output pointer may be NULL. */
int bfSensReadFrame(BFSensReader* r, uint64_t index, float* depthMetres, uint8_t* colorRGBX, float* cameraToWorld, uint64_t* timeStamps);
/* the same frame undecorated: depth ushort, colour RGB (3 bytes per pixel) */
int bfSensReadFrameRaw(BFSensReader* r, uint64_t index, uint16_t* depth, uint8_t* colorRGB);
void bfSensClose(BFSensReader* r);

typedef struct BFSensWriter BFSensWriter;
/* creates a version-4 file; header->colorCompression must be RAW, header->depthCompression RAW_USHORT or ZLIB_USHORT; numFrames is filled in by bfSensFinish */
int bfSensCreate(const char* path, const BFSensHeader* header, BFSensWriter** out);
int bfSensAppendFrame(BFSensWriter* w, const uint16_t* depth, const uint8_t* colorRGB, const float* cameraToWorld, uint64_t timeStampColor, uint64_t timeStampDepth);
int bfSensFinish(BFSensWriter* w);          /* writes the frame count and an empty IMU list, closes and frees the writer */

/* stand-alone decoders (what bfSensReadFrame uses): *width / *height are outputs; rgb must hold 3 * width * height bytes -- call with rgb == NULL to get the size first */
int bfSensDecodeJpeg(const uint8_t* data, size_t bytes, uint8_t* rgb, uint32_t* width, uint32_t* height);
int bfSensDecodePng(const uint8_t* data, size_t bytes, uint8_t* rgb, uint32_t* width, uint32_t* height);
const char* bfSensErrorString(int code);

#ifdef __cplusplus
}
#endif
#endif /* BF_SENS_H */
