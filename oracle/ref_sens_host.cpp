// ref_sens_host.cpp -- TEST INFRASTRUCTURE ONLY (oracle/build_ref.py: build_sens_host).  The codecs the reference's `.sens` payloads go through: mLib's
// ml::SensorData (external/mLib/include/ext-depthcamera/sensorData.h:540-668) decodes colour with stbi_load_from_memory and depth with stbi_zlib_decode_malloc,
// and compresses depth with stbi_zlib_compress -- the stb_image v2.08 / stb_image_write it vendors under ext-depthcamera/sensorData/.  Those two headers are
// compiled here by g++ from where they lie under /root/reference (sensorData.h itself needs all of mLib, which g++ cannot compile; its on-disk layout is pinned by
// an independent reader in tests/test_sens_io.py).  STBI_NO_SIMD: stb documents its SSE2 paths as bit-identical to the portable ones.
#include <cstring>

#define STBI_NO_SIMD
#define STB_IMAGE_IMPLEMENTATION
#define STB_IMAGE_WRITE_IMPLEMENTATION
#include "sensorData/stb_image.h"
#include "sensorData/stb_image_write.h"

// RGBDFrame::decompressColorAlloc_stb: 3 channels requested
extern "C" int ref_stb_decode(const unsigned char* data, int bytes, unsigned char* rgb, int* w, int* h) {
    int comp = 0;
    unsigned char* raw = stbi_load_from_memory(data, bytes, w, h, &comp, 3);
    if (!raw) return 1;
    if (rgb) memcpy(rgb, raw, (size_t)3 * (*w) * (*h));
    stbi_image_free(raw);
    return 0;
}
// RGBDFrame::compressDepth (TYPE_ZLIB_USHORT): returns the compressed size, or -1 when `cap` is too small
extern "C" int ref_stb_zlib_compress(const unsigned char* data, int bytes, unsigned char* out, int cap, int quality) {
    int n = 0;
    unsigned char* z = stbi_zlib_compress(const_cast<unsigned char*>(data), bytes, &n, quality);
    if (!z) return -1;
    if (n <= cap) memcpy(out, z, n);
    free(z);
    return n <= cap ? n : -1;
}
// RGBDFrame::decompressDepthAlloc_stb
extern "C" int ref_stb_zlib_decode(const unsigned char* data, int bytes, unsigned char* out, int cap) {
    int n = 0;
    char* r = stbi_zlib_decode_malloc(reinterpret_cast<const char*>(data), bytes, &n);
    if (!r) return -1;
    if (n <= cap) memcpy(out, r, n);
    free(r);
    return n <= cap ? n : -1;
}
