/*
 * hostio.h -- the file formats either side of the path (host only, no CUDA):
 *
 *   read_vol_file   the grid volume file a `gridvolume` plugin maps (src/volume/gridvolume.cpp:217-287: "VOL", version 3,
 *                   little endian, type / xres / yres / zres / channels as int32, the data AABB as six float32, then the
 *                   voxels, x fastest).  The path takes the density of a heterogeneous medium from it: one channel, float32
 *                   or uint8.  uint8 voxels go through the reference's density map i / 255.0f (gridvolume.cpp:212-215,
 *                   374-389), which gives the same eight fp32 corner values its trilinear lookup reads, so the float
 *                   lookup on the converted grid returns the same density.
 *   write_npy_f32   the film's NumPy output (src/films/mfilm.cpp:337-348 through cnpy::npy_save, src/films/cnpy.h:207-236):
 *                   format 1.0, '<f4', C order, shape (height, width, channels) -- (height, width) for one channel --, the
 *                   dictionary padded with spaces to a multiple of 16 bytes and closed by a newline.
 *
 *   read_vrl_file   the ASCII VRL file of the `vrlFile` property (see below).
 *
 * All throw HostIoError {code, message}; capi.cu maps it to the ABI's status + alvrl_last_error(), host_test_api.cpp
 * exposes them to the CPU tests.
 */
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <locale>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace alvrl {

struct HostIoError : std::runtime_error {
    int code;                         /* ALVRL_ERR_IO (-4), ALVRL_ERR_ARG (-1) or ALVRL_ERR_UNSUPPORTED (-5) */
    HostIoError(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

enum VolType { VOL_FLOAT32 = 1, VOL_FLOAT16 = 2, VOL_UINT8 = 3, VOL_QUANTIZED_DIRECTIONS = 4 };   /* gridvolume.cpp:101-106 */

struct VolFile {
    int32_t type = 0, res[3] = {0, 0, 0}, channels = 0;
    float bmin[3] = {0, 0, 0}, bmax[3] = {0, 0, 0};      /* the AABB stored in the file */
    std::vector<float> density;                          /* res[0] * res[1] * res[2] values, x fastest */
};

namespace hostio_detail {
struct File {
    FILE *f;
    explicit File(FILE *f_) : f(f_) {}
    ~File() { if (f) fclose(f); }
    File(const File &) = delete;
    File &operator=(const File &) = delete;
};
inline uint32_t le32(const unsigned char *p) { return (uint32_t) p[0] | ((uint32_t) p[1] << 8) | ((uint32_t) p[2] << 16) | ((uint32_t) p[3] << 24); }
inline float lef32(const unsigned char *p) { const uint32_t u = le32(p); float v; memcpy(&v, &u, 4); return v; }
}   // namespace hostio_detail

/* headerOnly: type, resolution, channels and AABB without reading the voxels */
inline void read_vol_file(const char *path, VolFile &out, bool headerOnly = false) {
    using namespace hostio_detail;
    if (!path) throw HostIoError(-1, "null volume file name");
    File fh(fopen(path, "rb"));
    if (!fh.f) throw HostIoError(-4, std::string("cannot open volume data file ") + path);
    unsigned char h[48];                                /* 3 + 1 + 5 * 4 + 6 * 4: the voxels start at float 12 (gridvolume.cpp:285) */
    if (fread(h, 1, sizeof(h), fh.f) != sizeof(h)) throw HostIoError(-4, "Encountered an invalid volume data file (truncated header)");
    if (h[0] != 'V' || h[1] != 'O' || h[2] != 'L') throw HostIoError(-1, "Encountered an invalid volume data file (incorrect header identifier)");
    if (h[3] != 3) throw HostIoError(-1, "Encountered an invalid volume data file (incorrect file version)");
    out.type = (int32_t) le32(h + 4);
    for (int i = 0; i < 3; i++) out.res[i] = (int32_t) le32(h + 8 + 4 * i);
    out.channels = (int32_t) le32(h + 20);
    for (int i = 0; i < 3; i++) { out.bmin[i] = lef32(h + 24 + 4 * i); out.bmax[i] = lef32(h + 36 + 4 * i); }
    const std::string what = " (type=" + std::to_string(out.type) + ", channels=" + std::to_string(out.channels) + ")";
    switch (out.type) {
        case VOL_FLOAT32:
        case VOL_UINT8:
            if (out.channels != 1 && out.channels != 3)
                throw HostIoError(-1, "Encountered an unsupported volume data file, only 1 and 3 channels are supported" + what);
            if (out.channels != 1)      /* a density is looked up with lookupFloat: one channel (heterogeneous.cpp:227-260) */
                throw HostIoError(-5, "the density of the medium needs a one-channel volume" + what);
            break;
        case VOL_FLOAT16: throw HostIoError(-5, "Error: float16 volumes are not yet supported!");
        case VOL_QUANTIZED_DIRECTIONS: throw HostIoError(-5, "quantized direction volumes hold no density" + what);
        default: throw HostIoError(-1, "Encountered a volume data file of unknown type" + what + "!");
    }
    for (int i = 0; i < 3; i++)
        if (out.res[i] < 1 || out.res[i] > (1 << 14)) throw HostIoError(-1, "Encountered an invalid volume data file (resolution out of range)");
    if (headerOnly) { out.density.clear(); return; }
    const size_t n = (size_t) out.res[0] * (size_t) out.res[1] * (size_t) out.res[2];
    /* the voxels the header announces must be in the file: checked against its length before anything is allocated (a corrupt
     * header can announce terabytes; the reference maps the file and would read past the mapping) */
    if (fseek(fh.f, 0, SEEK_END) != 0) throw HostIoError(-4, "cannot seek in the volume data file");
    const long fileSize = ftell(fh.f);
    if (fileSize < 0 || fseek(fh.f, (long) sizeof(h), SEEK_SET) != 0) throw HostIoError(-4, "cannot seek in the volume data file");
    if ((size_t) fileSize - sizeof(h) < n * (out.type == VOL_FLOAT32 ? 4u : 1u))
        throw HostIoError(-4, "Encountered an invalid volume data file (fewer voxels than the header announces)");
    out.density.resize(n);
    if (out.type == VOL_FLOAT32) {
        if (fread(out.density.data(), 4, n, fh.f) != n) throw HostIoError(-4, "Encountered an invalid volume data file (fewer voxels than the header announces)");
        const uint16_t one = 1;
        if (*(const unsigned char *) &one != 1)         /* big-endian host: the file is little endian (gridvolume.cpp:222) */
            for (size_t i = 0; i < n; i++) out.density[i] = lef32((const unsigned char *) &out.density[i]);
    } else {
        float map[256];
        for (int i = 0; i < 255; i++) map[i] = i / 255.0f;               /* m_densityMap, gridvolume.cpp:212-215 */
        map[255] = 1.0f;
        std::vector<unsigned char> buf(1 << 20);
        size_t done = 0;
        while (done < n) {
            const size_t want = std::min(buf.size(), n - done);
            if (fread(buf.data(), 1, want, fh.f) != want) throw HostIoError(-4, "Encountered an invalid volume data file (fewer voxels than the header announces)");
            for (size_t i = 0; i < want; i++) out.density[done + i] = map[buf[i]];
            done += want;
        }
    }
}

/* The ASCII VRL file of the `vrlFile` property (src/integrators/vrl/VRL.h:43-54, 120-128): one VRL per line, nine numbers
 * "sx sy sz ex ey ez r g b".  The reference reads lines until the stream throws at the end of the file; a VRL whose power is
 * not valid (NaN, infinite or negative) makes its constructor Log(EError), which throws into the same catch block (124-126) --
 * the reading ends there and the VRLs read so far are kept.  A line that does not hold nine numbers leaves the reference with
 * uninitialised members (undefined behaviour); here it ends the reading like the end of the file.  The put() filter (zero
 * power, zero length, 148-158) is applied by the caller (alvrl_set_vrls). */
inline void read_vrl_file(const char *path, std::vector<float> &start, std::vector<float> &end, std::vector<float> &power) {
    using namespace hostio_detail;
    if (!path) throw HostIoError(-1, "null VRL file name");
    File fh(fopen(path, "rb"));
    if (!fh.f) throw HostIoError(-4, std::string("cannot open VRL file ") + path);
    std::string line;
    bool eof = false;
    while (!eof) {
        line.clear();
        int ch;
        while ((ch = fgetc(fh.f)) != EOF && ch != 10) if (ch != 13) line.push_back((char) ch);      /* Stream::readLine, stream.cpp:392-414 */
        eof = ch == EOF;
        if (eof && line.empty()) break;
        float v[9];
        int k = 0;
        std::istringstream ss(line);                        /* as VRL.h:47-50: operator>> on a stringstream of the line ... */
        ss.imbue(std::locale::classic());                   /* ... in the "C" locale whatever the host process has set */
        while (k < 9 && (ss >> v[k])) k++;
        if (k < 9) break;
        bool valid = true;
        for (int i = 6; i < 9; i++) valid = valid && std::isfinite(v[i]) && v[i] >= 0;               /* Spectrum::isValid, VRL.h:51-53 */
        if (!valid) break;
        start.insert(start.end(), v, v + 3); end.insert(end.end(), v + 3, v + 6); power.insert(power.end(), v + 6, v + 9);
    }
}

/* data: height x width x channels, row-major */
inline void write_npy_f32(const char *path, const float *data, uint32_t height, uint32_t width, uint32_t channels) {
    using namespace hostio_detail;
    if (!path || !data) throw HostIoError(-1, "write_npy: null argument");
    if (!height || !width || !channels) throw HostIoError(-1, "write_npy: empty image");
    const uint16_t probe = 1;
    const bool little = *(const unsigned char *) &probe == 1;
    std::string dict = std::string("{'descr': '") + (little ? '<' : '>') + "f4', 'fortran_order': False, 'shape': (" + std::to_string(height) + ", " + std::to_string(width);
    if (channels != 1) dict += ", " + std::to_string(channels);                     /* mfilm.cpp:343-344: N = 2 for one channel */
    dict += "), }";
    dict.append(16 - (10 + dict.size()) % 16, ' ');                                /* preamble (10 bytes) + dictionary: a multiple of 16 */
    dict.back() = '\n';
    File fh(fopen(path, "wb"));
    if (!fh.f) throw HostIoError(-4, std::string("Output file cannot be created: ") + path);
    const unsigned char pre[10] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0, (unsigned char) (dict.size() & 0xff), (unsigned char) (dict.size() >> 8)};
    const size_t n = (size_t) height * width * channels;
    if (fwrite(pre, 1, 10, fh.f) != 10 || fwrite(dict.data(), 1, dict.size(), fh.f) != dict.size() || fwrite(data, 4, n, fh.f) != n)
        throw HostIoError(-4, std::string("short write to ") + path);
}

}   // namespace alvrl
