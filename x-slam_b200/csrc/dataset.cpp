// dataset.cpp — depth-stream readers in front of the frame loop: the reference's Dataset API
// (XKinectFusion/include/Dataset.h:18-81, src/Dataset.cpp:3-124) and loadTxtMatrix (src/IOHelper.cpp:4-19) without
// OpenCV / Eigen: a PNG decoder for the 16-bit greyscale depth images both benchmarks ship (zlib inflate + the five PNG
// scan-line filters), the ICL-NUIM layout (`depth/<i>.png`, raw / 5 -> millimetres, `livingRoom1n.gt.sim` poses) and the
// 7-Scenes layout (`seq-XX/frame-%06d.depth.png` + `.pose.txt`).  Host code only; the frames it returns are what
// xs_kinfu_process_frame uploads.
#include "../../include/xslam_b200.h"

#include <zlib.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace xs {
void set_error(const std::string &msg);
}
using xs::set_error;

namespace {

uint32_t be32(const unsigned char *p) { return ((uint32_t) p[0] << 24) | ((uint32_t) p[1] << 16) | ((uint32_t) p[2] << 8) | p[3]; }

// Decodes a non-interlaced greyscale PNG (bit depth 8 or 16) into 16-bit samples, i.e. what
// cv::imread(path, cv::IMREAD_UNCHANGED) yields for the benchmark depth images (Dataset.cpp:7); 8-bit samples are widened.
bool decode_png_gray_impl(const std::string &path, int &rows, int &cols, std::vector<uint16_t> &out, std::string &err) {
    std::ifstream in(path, std::ios::binary);
    if (!in) {
        err = "cannot open " + path;
        return false;
    }
    std::vector<unsigned char> file((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (file.size() < 8 + 25 || std::memcmp(file.data(), sig, 8) != 0) {
        err = path + ": not a PNG file";
        return false;
    }
    size_t pos = 8;
    int depth = 0, color = -1, interlace = 0;
    rows = cols = 0;
    std::vector<unsigned char> idat;
    bool end = false;
    while (!end && pos + 12 <= file.size()) {
        const uint32_t len = be32(&file[pos]);
        const unsigned char *type = &file[pos + 4];
        if (pos + 12 + (size_t) len > file.size()) {
            err = path + ": truncated chunk";
            return false;
        }
        const unsigned char *data = &file[pos + 8];
        if (be32(data + len) != (uint32_t) crc32(crc32(0L, Z_NULL, 0), type, 4 + len)) {
            err = path + ": chunk CRC mismatch";
            return false;
        }
        if (!std::memcmp(type, "IHDR", 4) && len >= 13) {
            cols = (int) be32(data);
            rows = (int) be32(data + 4);
            depth = data[8];
            color = data[9];
            interlace = data[12];
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), data, data + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            end = true;
        }
        pos += 12 + (size_t) len;
    }
    if (rows <= 0 || cols <= 0 || color != 0 || (depth != 16 && depth != 8) || interlace != 0) {
        err = path + ": only non-interlaced 8/16-bit greyscale PNG depth images are supported";
        return false;
    }
    // the header is untrusted: bound the allocation (a depth frame is far below 64 Mpixel) before trusting it
    if ((uint64_t) rows * (uint64_t) cols > (1ull << 26)) {
        err = path + ": IHDR announces an implausible image size";
        return false;
    }
    const size_t bpp = depth / 8, stride = (size_t) cols * bpp;
    std::vector<unsigned char> raw((stride + 1) * (size_t) rows);
    uLongf raw_len = (uLongf) raw.size();
    if (uncompress(raw.data(), &raw_len, idat.data(), (uLong) idat.size()) != Z_OK || raw_len != raw.size()) {
        err = path + ": zlib inflate failed";
        return false;
    }
    // scan-line filters (PNG spec 9.2): a = left, b = up, c = up-left, at a distance of one pixel (bpp bytes)
    std::vector<unsigned char> prev(stride, 0), cur(stride);
    out.resize((size_t) rows * cols);
    for (int y = 0; y < rows; ++y) {
        const unsigned char *line = &raw[(stride + 1) * (size_t) y];
        const int ft = line[0];
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
            int pred = 0;
            switch (ft) {
            case 0: pred = 0; break;
            case 1: pred = a; break;
            case 2: pred = b; break;
            case 3: pred = (a + b) >> 1; break;
            case 4: {
                const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
                pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                break;
            }
            default: err = path + ": bad scan-line filter"; return false;
            }
            cur[i] = (unsigned char) (line[1 + i] + pred);
        }
        for (int x = 0; x < cols; ++x)
            out[(size_t) y * cols + x] = bpp == 2 ? (uint16_t) ((cur[2 * x] << 8) | cur[2 * x + 1]) : (uint16_t) cur[x];
        prev.swap(cur);
    }
    return true;
}

// no exception leaves the library through an extern "C" function: allocation failures on corrupt input become errors
bool decode_png_gray(const std::string &path, int &rows, int &cols, std::vector<uint16_t> &out, std::string &err) {
    try {
        return decode_png_gray_impl(path, rows, cols, out, err);
    } catch (const std::exception &e) {
        err = path + ": " + e.what();
        return false;
    }
}

// IOHelper.cpp:4-19 loadTxtMatrix: whitespace-separated floats, row by row
bool load_txt_matrix(const std::string &path, int rows, int cols, float *out) {
    std::ifstream in(path);
    if (!in) return false;
    for (int i = 0; i < rows * cols; ++i) {
        float x = 0.f;
        in >> x;
        out[i] = x;
    }
    return true;
}

}  // namespace

struct xs_dataset {
    int factor = 1;
    bool flip = false;
    std::vector<std::string> depth_files, stamps;
    std::vector<float> poses;  // [n][16] row-major
};

extern "C" {

int xs_read_png16(const char *path, uint16_t *out_host, long capacity, int *rows, int *cols) {
    std::vector<uint16_t> img;
    std::string err;
    int r = 0, c = 0;
    if (!path || !rows || !cols || !decode_png_gray(path, r, c, img, err)) {
        set_error(err.empty() ? "xs_read_png16: null argument" : err);
        return XS_ERR_ARG;
    }
    *rows = r;
    *cols = c;
    if (out_host) {
        if ((long) img.size() > capacity) {
            set_error("xs_read_png16: output buffer too small");
            return XS_ERR_ARG;
        }
        std::memcpy(out_host, img.data(), img.size() * sizeof(uint16_t));
    }
    return XS_OK;
}

int xs_load_txt_matrix(const char *path, int rows, int cols, float *out) {
    if (!path || !out || rows <= 0 || cols <= 0 || !load_txt_matrix(path, rows, cols, out)) {
        set_error(std::string("loadTxtMatrix: cannot open ") + (path ? path : "(null)"));
        return XS_ERR_ARG;
    }
    return XS_OK;
}

// ICL_Dataset::readPoseFile, Dataset.cpp:90-124: lines [start, end) of the .gt.sim file are the 3x4 top of the pose (one
// blank line follows every pose, so pose i starts at line 4 i); the last row is set to 0 0 0 1.
int xs_icl_read_pose_file(const char *poses_path, int start, int end, float *pose16) {
    std::ifstream poses_file(poses_path ? poses_path : "");
    if (!poses_file) {
        std::cout << "Error opening poses file." << std::endl;
        set_error(std::string("cannot open poses file ") + (poses_path ? poses_path : "(null)"));
        return XS_ERR_ARG;
    }
    for (int i = 0; i < 16; ++i) pose16[i] = (i % 5 == 0) ? 1.f : 0.f;
    int i = 0;
    std::string temp;
    while (std::getline(poses_file, temp)) {
        if (i < start) {
            i++;
        } else if (i >= start && i < end) {
            int j = 0;
            std::stringstream linestream(temp);
            std::string sub;
            while (linestream >> sub) {
                if (i - start < 4 && j < 4) pose16[(i - start) * 4 + j] = (float) std::strtod(sub.c_str(), nullptr);
                j++;
            }
            i++;
        } else {
            break;
        }
    }
    pose16[12] = pose16[13] = pose16[14] = 0.f;
    pose16[15] = 1.f;
    return XS_OK;
}

// ICL_Dataset::ICL_Dataset, Dataset.cpp:69-88
static xs_dataset *open_icl_impl(const char *dataset_dir, int start_frame, int end_frame, int is_flip);
xs_dataset *xs_dataset_open_icl(const char *dataset_dir, int start_frame, int end_frame, int is_flip) {
    try {
        return open_icl_impl(dataset_dir, start_frame, end_frame, is_flip);
    } catch (const std::exception &e) {
        set_error(std::string("xs_dataset_open_icl: ") + e.what());
        return nullptr;
    }
}
static xs_dataset *open_icl_impl(const char *dataset_dir, int start_frame, int end_frame, int is_flip) {
    if (!dataset_dir) return nullptr;
    xs_dataset *d = new xs_dataset;
    d->flip = is_flip != 0;
    d->factor = 5;
    const std::string dir(dataset_dir), poses_path = dir + "livingRoom1n.gt.sim";
    std::cout << "pose path:" << poses_path << std::endl;
    // one pass over the pose file instead of the reference's re-read per frame
    std::vector<std::string> lines;
    {
        std::ifstream f(poses_path);
        if (!f) {
            std::cout << "Error opening poses file." << std::endl;
            set_error("cannot open poses file " + poses_path);
            delete d;
            return nullptr;
        }
        std::string l;
        while (std::getline(f, l)) lines.push_back(l);
    }
    for (int i = start_frame; i <= end_frame; ++i) {
        const std::string format = std::to_string(i);
        d->stamps.push_back(format);
        d->depth_files.push_back(dir + "depth/" + format + ".png");
        float pose[16];
        for (int e = 0; e < 16; ++e) pose[e] = (e % 5 == 0) ? 1.f : 0.f;
        for (int r = 0; r < 3; ++r) {
            const size_t ln = (size_t) 4 * i + r;
            if (ln >= lines.size()) break;
            std::stringstream ss(lines[ln]);
            std::string sub;
            for (int j = 0; ss >> sub; ++j)
                if (j < 4) pose[r * 4 + j] = (float) std::strtod(sub.c_str(), nullptr);
        }
        pose[12] = pose[13] = pose[14] = 0.f;
        pose[15] = 1.f;
        d->poses.insert(d->poses.end(), pose, pose + 16);
    }
    return d;
}

// seven_scenes_Dataset::seven_scenes_Dataset, Dataset.cpp:13-39.  seq_names as readInfo returns them ("seq-01/").
static xs_dataset *open_seven_scenes_impl(const char *dataset_dir, const int *start_frames, const int *end_frames,
                                          const char *const *seq_names, int nseq, int is_flip);
xs_dataset *xs_dataset_open_seven_scenes(const char *dataset_dir, const int *start_frames, const int *end_frames,
                                         const char *const *seq_names, int nseq, int is_flip) {
    try {
        return open_seven_scenes_impl(dataset_dir, start_frames, end_frames, seq_names, nseq, is_flip);
    } catch (const std::exception &e) {
        set_error(std::string("xs_dataset_open_seven_scenes: ") + e.what());
        return nullptr;
    }
}
static xs_dataset *open_seven_scenes_impl(const char *dataset_dir, const int *start_frames, const int *end_frames,
                                          const char *const *seq_names, int nseq, int is_flip) {
    if (!dataset_dir || !start_frames || !end_frames || !seq_names || nseq < 0) return nullptr;
    xs_dataset *d = new xs_dataset;
    d->flip = is_flip != 0;
    d->factor = 1;
    const std::string dir(dataset_dir);
    for (int seq = 0; seq < nseq; ++seq)
        for (int frame = start_frames[seq]; frame <= end_frames[seq]; ++frame) {
            char num[32];
            std::snprintf(num, sizeof(num), "%06d", frame);
            const std::string base = std::string(seq_names[seq]) + "frame-" + num;
            d->stamps.push_back(base);
            d->depth_files.push_back(dir + base + ".depth.png");
            float pose[16];
            if (!load_txt_matrix(dir + base + ".pose.txt", 4, 4, pose)) {
                set_error("cannot open " + dir + base + ".pose.txt");
                delete d;
                return nullptr;
            }
            d->poses.insert(d->poses.end(), pose, pose + 16);
        }
    return d;
}

// seven_scenes_Dataset::readInfo, Dataset.cpp:41-67: line 0 = start frames, line 1 = end frames, line 2 = sequence ids.
// Writes up to max_seq entries; seq_names_out: max_seq x 16 chars ("seq-XX/").  Returns the number of sequences.
int xs_seven_scenes_read_info(const char *filename, int *start_frames, int *end_frames, char *seq_names_out, int max_seq) {
    std::ifstream in(filename ? filename : "");
    if (!in) {
        set_error(std::string("cannot open ") + (filename ? filename : "(null)"));
        return -1;
    }
    std::string line;
    int count = 0, n[3] = {0, 0, 0};
    while (std::getline(in, line)) {
        std::stringstream ss(line);
        std::string x;
        while (ss >> x) {
            if (count > 2 || n[count] >= max_seq) continue;
            if (count == 0) start_frames[n[0]] = std::atoi(x.c_str());
            if (count == 1) end_frames[n[1]] = std::atoi(x.c_str());
            if (count == 2) std::snprintf(seq_names_out + (size_t) n[2] * 16, 16, "seq-%s/", x.c_str());
            ++n[count];
        }
        count++;
    }
    return n[2];
}

int xs_dataset_size(const xs_dataset *d) { return d ? (int) d->depth_files.size() : 0; }

// Dataset::getDepthData, Dataset.cpp:3-11: imread(IMREAD_UNCHANGED); depth /= factor_ (cv::Mat arithmetic: the quotient is
// rounded to nearest, ties to even); optional horizontal flip.
int xs_dataset_get_depth(const xs_dataset *d, int index, uint16_t *out_host, int rows, int cols) {
    if (!d || !out_host || index < 0 || index >= (int) d->depth_files.size()) {
        set_error("xs_dataset_get_depth: bad index");
        return XS_ERR_ARG;
    }
    std::vector<uint16_t> img;
    std::string err;
    int r = 0, c = 0;
    if (!decode_png_gray(d->depth_files[index], r, c, img, err)) {
        set_error(err);
        return XS_ERR_ARG;
    }
    if (r != rows || c != cols) {
        set_error(d->depth_files[index] + ": image size differs from depth_height x depth_width");
        return XS_ERR_ARG;
    }
    const double scale = 1.0 / d->factor;
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            uint16_t v = img[(size_t) y * cols + x];
            if (d->factor != 1) v = (uint16_t) std::nearbyint(v * scale);
            out_host[(size_t) y * cols + (d->flip ? cols - 1 - x : x)] = v;
        }
    return XS_OK;
}

int xs_dataset_get_pose(const xs_dataset *d, int index, float *pose16) {
    if (!d || !pose16 || index < 0 || index >= (int) d->depth_files.size()) return XS_ERR_ARG;
    std::memcpy(pose16, &d->poses[(size_t) index * 16], 16 * sizeof(float));
    return XS_OK;
}

int xs_dataset_set_pose(xs_dataset *d, int index, const float *pose16) {
    if (!d || !pose16 || index < 0 || index >= (int) d->depth_files.size()) return XS_ERR_ARG;
    std::memcpy(&d->poses[(size_t) index * 16], pose16, 16 * sizeof(float));
    return XS_OK;
}

const char *xs_dataset_timestamp(const xs_dataset *d, int index) {
    return (d && index >= 0 && index < (int) d->stamps.size()) ? d->stamps[index].c_str() : "";
}

const char *xs_dataset_depth_filename(const xs_dataset *d, int index) {
    return (d && index >= 0 && index < (int) d->depth_files.size()) ? d->depth_files[index].c_str() : "";
}

void xs_dataset_close(xs_dataset *d) { delete d; }

}  // extern "C"
