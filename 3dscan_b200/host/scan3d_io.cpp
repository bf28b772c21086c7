// scan3d_io.cpp -- the file formats on either side of the hot path: BMP captures as
// cvLoadImage(..., CV_LOAD_IMAGE_GRAYSCALE) returns them (3/wrapped_phase.cpp:44,
// 4/phase_unwrap.cpp:78-90), the OpenCV-XML calibration matrices load_matrices() reads
// (6/system_calibration.cpp:1526-1554), and an "x y z red green blue" PLY like
// pcl::io::savePLYFile writes for PointXYZRGB (8/save_point_cloud.cpp:217).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/scan3d_host.h"

static thread_local std::string g_err;
static int fail(int code, const std::string& m)
{
    g_err = m;
    return code;
}

static uint32_t rd32(const uint8_t* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint16_t rd16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }

static bool slurp(const char* path, std::vector<uint8_t>& out)
{
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? (size_t)n : 0);
    const size_t got = n > 0 ? fread(out.data(), 1, (size_t)n, f) : 0;
    fclose(f);
    return got == out.size();
}

// OpenCV's fixed-point BGR->grey (cvtColor CV_BGR2GRAY, 14-bit coefficients)
static inline uint8_t grey_of(uint8_t b, uint8_t g, uint8_t r)
{
    return (uint8_t)((b * 1868 + g * 9617 + r * 4899 + 8192) >> 14);
}

extern "C" {

const char* scan3d_host_last_error(void) { return g_err.c_str(); }

// shared by the grey and the colour reader: header checks of an uncompressed 8- or 24-bit BMP
static int bmp_open(const char* path, std::vector<uint8_t>& b, int* w, int* h, int* hs, int* bpp, uint32_t* off, size_t* stride,
                    uint32_t* ncol, const uint8_t** pal)
{
    if (!slurp(path, b) || b.size() < 54 || b[0] != 'B' || b[1] != 'M')
        return fail(SCAN3D_ERR_IO, std::string("cannot read BMP ") + path);
    *off = rd32(&b[10]);
    const uint32_t hdr = rd32(&b[14]);
    *w = (int)rd32(&b[18]);
    *hs = (int)rd32(&b[22]);
    *bpp = rd16(&b[28]);
    const uint32_t comp = rd32(&b[30]);
    *h = *hs < 0 ? -*hs : *hs;
    if (comp != 0 || (*bpp != 8 && *bpp != 24) || *w <= 0 || *h <= 0 || *w > 65536 || *h > 65536)
        return fail(SCAN3D_ERR_IO, std::string("unsupported BMP flavour: ") + path);
    *stride = (((size_t)*w * *bpp + 31) / 32) * 4;
    if (hdr < 40 || hdr > 1024 || *off < 14 + hdr || *off > b.size() || b.size() - *off < *stride * (size_t)*h)
        return fail(SCAN3D_ERR_IO, std::string("truncated or malformed BMP ") + path);
    *ncol = 0;
    *pal = nullptr;
    if (*bpp == 8) {
        uint32_t n = rd32(&b[46]);
        if (n == 0 || n > 256) n = 256;
        // the palette lies between the info header and the pixel data: never read past either
        const size_t room = (*off - (14 + (size_t)hdr)) / 4;
        *ncol = n < room ? n : (uint32_t)room;
        *pal = &b[14 + hdr];
    }
    return SCAN3D_OK;
}

int scan3d_read_bmp8(const char* path, int* W, int* H, uint8_t* buf, int64_t buf_bytes)
{
    if (!path || !W || !H) return fail(SCAN3D_ERR_ARG, "null argument");
    std::vector<uint8_t> b;
    int w, h, hs, bpp;
    uint32_t off, ncol;
    size_t stride;
    const uint8_t* pal;
    const int rc = bmp_open(path, b, &w, &h, &hs, &bpp, &off, &stride, &ncol, &pal);
    if (rc) return rc;
    *W = w;
    *H = h;
    if (!buf) return SCAN3D_OK;
    if (buf_bytes < (int64_t)w * h) return fail(SCAN3D_ERR_ARG, "buffer too small");
    uint8_t lut[256];
    if (bpp == 8)
        for (uint32_t i = 0; i < 256; i++) lut[i] = i < ncol ? grey_of(pal[4 * i], pal[4 * i + 1], pal[4 * i + 2]) : 0;
    for (int y = 0; y < h; y++) {
        const uint8_t* src = &b[off + stride * (size_t)(hs > 0 ? h - 1 - y : y)];
        uint8_t* dst = buf + (size_t)y * w;
        if (bpp == 8)
            for (int x = 0; x < w; x++) dst[x] = lut[src[x]];
        else
            for (int x = 0; x < w; x++) dst[x] = grey_of(src[3 * x], src[3 * x + 1], src[3 * x + 2]);
    }
    return SCAN3D_OK;
}

// cvLoadImage(path) with its default flag (colour): [H][W][3] B,G,R bytes, top row first
// (8/save_point_cloud.cpp:59-66 loads Point_cloud/texture.bmp this way and cvSplits it into blue, green, red)
int scan3d_read_bmp_bgr(const char* path, int* W, int* H, uint8_t* buf, int64_t buf_bytes)
{
    if (!path || !W || !H) return fail(SCAN3D_ERR_ARG, "null argument");
    std::vector<uint8_t> b;
    int w, h, hs, bpp;
    uint32_t off, ncol;
    size_t stride;
    const uint8_t* pal;
    const int rc = bmp_open(path, b, &w, &h, &hs, &bpp, &off, &stride, &ncol, &pal);
    if (rc) return rc;
    *W = w;
    *H = h;
    if (!buf) return SCAN3D_OK;
    if (buf_bytes < (int64_t)3 * w * h) return fail(SCAN3D_ERR_ARG, "buffer too small");
    for (int y = 0; y < h; y++) {
        const uint8_t* src = &b[off + stride * (size_t)(hs > 0 ? h - 1 - y : y)];
        uint8_t* dst = buf + (size_t)3 * y * w;
        if (bpp == 24) {
            memcpy(dst, src, (size_t)3 * w);
        } else {
            for (int x = 0; x < w; x++)
                for (int k = 0; k < 3; k++) dst[3 * x + k] = src[x] < ncol ? pal[4 * src[x] + k] : 0;
        }
    }
    return SCAN3D_OK;
}

int scan3d_write_bmp8(const char* path, int W, int H, const uint8_t* buf)
{
    if (!path || !buf || W <= 0 || H <= 0) return fail(SCAN3D_ERR_ARG, "bad argument");
    const size_t stride = ((size_t)W + 3) & ~(size_t)3;
    std::vector<uint8_t> out(54 + 1024 + stride * H, 0);
    out[0] = 'B'; out[1] = 'M';
    auto w32 = [&](size_t o, uint32_t v) { out[o] = v; out[o + 1] = v >> 8; out[o + 2] = v >> 16; out[o + 3] = v >> 24; };
    w32(2, (uint32_t)out.size()); w32(10, 54 + 1024); w32(14, 40); w32(18, (uint32_t)W); w32(22, (uint32_t)H);
    out[26] = 1; out[28] = 8; w32(34, (uint32_t)(stride * H)); w32(46, 256);
    for (int i = 0; i < 256; i++) out[54 + 4 * i] = out[54 + 4 * i + 1] = out[54 + 4 * i + 2] = (uint8_t)i;
    for (int y = 0; y < H; y++) memcpy(&out[54 + 1024 + stride * (size_t)(H - 1 - y)], buf + (size_t)y * W, W);
    FILE* f = fopen(path, "wb");
    if (!f) return fail(SCAN3D_ERR_IO, std::string("cannot write ") + path);
    const bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
    return (fclose(f) == 0 && ok) ? SCAN3D_OK : fail(SCAN3D_ERR_IO, "short write");
}

int scan3d_read_cv_matrix(const char* path, const char* name, int rows, int cols, double* out)
{
    if (!path || !name || !out) return fail(SCAN3D_ERR_ARG, "null argument");
    std::vector<uint8_t> b;
    if (!slurp(path, b)) return fail(SCAN3D_ERR_IO, std::string("cannot read ") + path);
    const std::string t(b.begin(), b.end());
    const size_t tag = t.find(std::string("<") + name);
    if (tag == std::string::npos) return fail(SCAN3D_ERR_IO, std::string("no <") + name + "> in " + path);
    auto int_of = [&](const char* key) {
        const size_t p = t.find(key, tag);
        return p == std::string::npos ? -1 : atoi(t.c_str() + p + strlen(key));
    };
    if (int_of("<rows>") * int_of("<cols>") != rows * cols)
        return fail(SCAN3D_ERR_IO, std::string("unexpected matrix size in ") + path);
    const size_t d0 = t.find("<data>", tag), d1 = t.find("</data>", tag);
    if (d0 == std::string::npos || d1 == std::string::npos) return fail(SCAN3D_ERR_IO, "no <data>");
    const std::string data = t.substr(d0 + 6, d1 - d0 - 6);
    const char* p = data.c_str();
    for (int i = 0; i < rows * cols; i++) {
        char* end = nullptr;
        out[i] = strtod(p, &end);
        if (end == p) return fail(SCAN3D_ERR_IO, std::string("short <data> in ") + path);
        p = end;
    }
    return SCAN3D_OK;
}

int scan3d_load_calibration(const char* root, scan3d_calib* cal)
{
    if (!root || !cal) return fail(SCAN3D_ERR_ARG, "null argument");
    const std::string r = std::string(root) + "/";
    struct Item { const char* file; const char* name; int rows, cols; double* dst; };
    const Item items[] = {
        {"Camera_calibration/Matrices/cam_intrinsic_mat.xml", "cam_intrinsic_mat", 3, 3, cal->Kc},
        {"Camera_calibration/Matrices/cam_distortion_vect.xml", "cam_distortion_vect", 5, 1, cal->dc},
        {"Projector_calibration/Matrices/proj_intrinsic_mat.xml", "proj_intrinsic_mat", 3, 3, cal->Kp},
        {"Projector_calibration/Matrices/proj_distortion_vect.xml", "proj_distortion_vect", 5, 1, cal->dp},
        {"Triangulation/Camera_extrinsic_parametrs/world_to_cam_rot_vect.xml", "world_to_cam_rot_vect", 3, 1, cal->rc},
        {"Triangulation/Camera_extrinsic_parametrs/world_to_cam_trans_vect.xml", "world_to_cam_trans_vect", 3, 1, cal->tc},
        {"Triangulation/Projector_extrinsic_parametrs/world_to_proj_rot_vect.xml", "world_to_proj_rot_vect", 3, 1, cal->rp},
        {"Triangulation/Projector_extrinsic_parametrs/world_to_proj_trans_vect.xml", "world_to_proj_trans_vect", 3, 1, cal->tp},
    };
    for (const Item& it : items) {
        const int rc = scan3d_read_cv_matrix((r + it.file).c_str(), it.name, it.rows, it.cols, it.dst);
        if (rc) return rc;
    }
    return SCAN3D_OK;
}

static int load_one(const std::string& dir, const char* prefix, int i, int W, int H, uint8_t* dst)
{
    // the colour originals ("Captured_image_i.bmp") are read like cvLoadImage(GRAYSCALE); the
    // grey re-saves the reference writes next to them ("Gray_captured_image_i.bmp") are identical
    const std::string a = dir + prefix + "Captured_image_" + std::to_string(i) + ".bmp";
    const std::string b = dir + prefix + "Gray_captured_image_" + std::to_string(i) + ".bmp";
    int w = 0, h = 0;
    int rc = scan3d_read_bmp8(a.c_str(), &w, &h, nullptr, 0);
    const std::string& use = rc == SCAN3D_OK ? a : b;
    if (rc != SCAN3D_OK) rc = scan3d_read_bmp8(b.c_str(), &w, &h, nullptr, 0);
    if (rc != SCAN3D_OK) return rc;
    if (w != W || h != H) return fail(SCAN3D_ERR_CONFIG, "captured image size differs from the config: " + use);
    return scan3d_read_bmp8(use.c_str(), &w, &h, dst, (int64_t)W * H);
}

int scan3d_load_captured_set(const char* root, const scan3d_config* cfg, uint8_t* stack)
{
    if (!root || !cfg || !stack) return fail(SCAN3D_ERR_ARG, "null argument");
    const size_t plane = (size_t)cfg->W * cfg->H;
    uint8_t* p = stack;
    const char* dirs[2] = {"Vertical", "Horizontal"};
    for (int d = 0; d < cfg->dirs; d++) {
        const int M = d == 0 ? cfg->M_v : cfg->M_h;
        const std::string fr = std::string(root) + "/Captured_patterns/Fringe_patterns/" + dirs[d] + "/Undistorted/";
        const std::string gc = std::string(root) + "/Captured_patterns/Coded_patterns/Gray_coded/" + dirs[d] + "/Undistorted/";
        for (int i = 0; i < cfg->N; i++, p += plane) {
            const int rc = load_one(fr, "", i, cfg->W, cfg->H, p);
            if (rc) return rc;
        }
        for (int i = 0; i < M; i++, p += plane) {
            const int rc = load_one(gc, "", i, cfg->W, cfg->H, p);
            if (rc) return rc;
        }
        for (int i = 0; i < M; i++, p += plane) {
            const int rc = load_one(gc, "inverse_", i, cfg->W, cfg->H, p);
            if (rc) return rc;
        }
    }
    return SCAN3D_OK;
}

int scan3d_write_ply_points(const char* path, const float* xyz, const uint8_t* rgb, int64_t n, int binary)
{
    if (!path || (!xyz && n > 0) || n < 0) return fail(SCAN3D_ERR_ARG, "bad argument");
    FILE* f = fopen(path, "wb");
    if (!f) return fail(SCAN3D_ERR_IO, std::string("cannot write ") + path);
    fprintf(f, "ply\nformat %s 1.0\ncomment scan3d-b200\nelement vertex %lld\n"
               "property float x\nproperty float y\nproperty float z\n"
               "property uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n",
            binary ? "binary_little_endian" : "ascii", (long long)n);
    for (int64_t i = 0; i < n; i++) {
        const uint8_t c[3] = {rgb ? rgb[3 * i] : (uint8_t)0, rgb ? rgb[3 * i + 1] : (uint8_t)0, rgb ? rgb[3 * i + 2] : (uint8_t)0};
        if (binary) {
            fwrite(&xyz[3 * i], 4, 3, f);
            fwrite(c, 1, 3, f);
        } else {
            fprintf(f, "%.9g %.9g %.9g %u %u %u\n", xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], c[0], c[1], c[2]);
        }
    }
    return fclose(f) == 0 ? SCAN3D_OK : fail(SCAN3D_ERR_IO, "short write");
}

// pcl::io::savePCDFileASCII of a PointXYZRGB cloud (8/save_point_cloud.cpp:212; PCL 1.6 writer):
// one "x y z rgb" line per point, rgb = (r<<16 | g<<8 | b) reinterpreted as a float, all four
// printed with the stream precision PCL sets (8 significant digits).  No PCL-written file
// survives in the reference tree, so the header text follows the PCD v0.7 specification.
int scan3d_write_pcd_points(const char* path, const float* xyz, const uint8_t* rgb, int64_t n)
{
    if (!path || (!xyz && n > 0) || n < 0) return fail(SCAN3D_ERR_ARG, "bad argument");
    FILE* f = fopen(path, "wb");
    if (!f) return fail(SCAN3D_ERR_IO, std::string("cannot write ") + path);
    fprintf(f, "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z rgb\nSIZE 4 4 4 4\n"
               "TYPE F F F F\nCOUNT 1 1 1 1\nWIDTH %lld\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %lld\nDATA ascii\n",
            (long long)n, (long long)n);
    for (int64_t i = 0; i < n; i++) {
        const uint32_t packed = rgb ? ((uint32_t)rgb[3 * i] << 16) | ((uint32_t)rgb[3 * i + 1] << 8) | rgb[3 * i + 2] : 0u;
        float c;
        memcpy(&c, &packed, 4);
        fprintf(f, "%.8g %.8g %.8g %.8g\n", xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], c);
    }
    return fclose(f) == 0 ? SCAN3D_OK : fail(SCAN3D_ERR_IO, "short write");
}

// pcl::io::loadPLYFile's job in register_point_clouds (9/register_point_clouds.cpp:66,87): vertex
// positions and colours of a PLY written by scan3d_write_ply_points / PCL ("x y z" floats or doubles,
// optional uchar "red green blue"), ascii or binary_little_endian.  Other vertex properties are skipped.
int scan3d_read_ply_points(const char* path, float* xyz, uint8_t* rgb, int64_t capacity, int64_t* n_out)
{
    if (!path || !n_out) return fail(SCAN3D_ERR_ARG, "bad argument");
    std::vector<uint8_t> b;
    if (!slurp(path, b)) return fail(SCAN3D_ERR_IO, std::string("cannot read ") + path);
    b.push_back(0);   // strtod below needs a terminator
    struct Prop { std::string type, name; int size; };
    std::vector<Prop> props;
    size_t pos = 0;
    bool in_vertex = false, binary = false, seen_end = false;
    long long count = -1;
    auto type_size = [](const std::string& t) {
        if (t == "char" || t == "uchar" || t == "int8" || t == "uint8") return 1;
        if (t == "short" || t == "ushort" || t == "int16" || t == "uint16") return 2;
        if (t == "int" || t == "uint" || t == "float" || t == "int32" || t == "uint32" || t == "float32") return 4;
        if (t == "double" || t == "float64") return 8;
        return 0;
    };
    while (pos < b.size()) {
        size_t e = pos;
        while (e < b.size() && b[e] != '\n') e++;
        std::string line((const char*)&b[pos], e - pos);
        if (!line.empty() && line.back() == '\r') line.pop_back();
        pos = e + 1;
        char a[64] = {0}, c[64] = {0}, d[64] = {0};
        const int nf = sscanf(line.c_str(), "%63s %63s %63s", a, c, d);
        if (nf >= 1 && !strcmp(a, "end_header")) { seen_end = true; break; }
        if (nf >= 2 && !strcmp(a, "format")) {
            if (!strcmp(c, "binary_little_endian")) binary = true;
            else if (strcmp(c, "ascii")) return fail(SCAN3D_ERR_IO, std::string("unsupported PLY format in ") + path);
        } else if (nf >= 3 && !strcmp(a, "element")) {
            in_vertex = !strcmp(c, "vertex");
            if (in_vertex) count = atoll(d);
            else if (count < 0) return fail(SCAN3D_ERR_IO, "PLY: an element precedes the vertices");
        } else if (nf >= 3 && !strcmp(a, "property") && in_vertex) {
            if (!strcmp(c, "list")) return fail(SCAN3D_ERR_IO, "PLY: list property on vertices");
            const int sz = type_size(c);
            if (!sz) return fail(SCAN3D_ERR_IO, std::string("PLY: unknown property type ") + c);
            props.push_back({c, d, sz});
        }
    }
    if (!seen_end || count < 0) return fail(SCAN3D_ERR_IO, std::string("not a PLY vertex file: ") + path);
    int ix = -1, iy = -1, iz = -1, ir = -1, ig = -1, ib = -1;
    for (int k = 0; k < (int)props.size(); k++) {
        const std::string& n = props[k].name;
        if (n == "x") ix = k; else if (n == "y") iy = k; else if (n == "z") iz = k;
        else if (n == "red" || n == "r") ir = k; else if (n == "green" || n == "g") ig = k;
        else if (n == "blue" || n == "b") ib = k;
    }
    if (ix < 0 || iy < 0 || iz < 0) return fail(SCAN3D_ERR_IO, "PLY: no x/y/z properties");
    *n_out = count;
    if (!xyz) return SCAN3D_OK;   // size query
    if (count > capacity) return fail(SCAN3D_ERR_ARG, "PLY: more vertices than the buffer holds");
    std::vector<double> v(props.size());
    for (long long i = 0; i < count; i++) {
        for (size_t k = 0; k < props.size(); k++) {
            if (binary) {
                if (pos + props[k].size > b.size()) return fail(SCAN3D_ERR_IO, "PLY: truncated");
                const uint8_t* q = &b[pos];
                const std::string& t = props[k].type;
                if (t == "float" || t == "float32") { float f; memcpy(&f, q, 4); v[k] = f; }
                else if (t == "double" || t == "float64") { double f; memcpy(&f, q, 8); v[k] = f; }
                else if (props[k].size == 1) v[k] = (t[0] == 'u') ? (double)q[0] : (double)(int8_t)q[0];
                else if (props[k].size == 2) v[k] = (t[0] == 'u') ? (double)rd16(q) : (double)(int16_t)rd16(q);
                else v[k] = (t[0] == 'u') ? (double)rd32(q) : (double)(int32_t)rd32(q);
                pos += props[k].size;
            } else {
                while (pos < b.size() && (b[pos] == ' ' || b[pos] == '\n' || b[pos] == '\r' || b[pos] == '\t')) pos++;
                if (pos >= b.size()) return fail(SCAN3D_ERR_IO, "PLY: truncated");
                char* endp = nullptr;
                v[k] = strtod((const char*)&b[pos], &endp);
                if (endp == (const char*)&b[pos]) return fail(SCAN3D_ERR_IO, "PLY: bad number");
                pos = (size_t)(endp - (const char*)b.data());
            }
        }
        xyz[3 * i] = (float)v[ix]; xyz[3 * i + 1] = (float)v[iy]; xyz[3 * i + 2] = (float)v[iz];
        if (rgb) {
            rgb[3 * i] = ir >= 0 ? (uint8_t)v[ir] : 0;
            rgb[3 * i + 1] = ig >= 0 ? (uint8_t)v[ig] : 0;
            rgb[3 * i + 2] = ib >= 0 ? (uint8_t)v[ib] : 0;
        }
    }
    return SCAN3D_OK;
}

}  // extern "C"
