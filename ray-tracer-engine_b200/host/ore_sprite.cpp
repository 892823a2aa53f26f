// ore_sprite.cpp - sprite/buffer for the headless build (format of /root/reference/Sprite.cpp:13-65, no OpenCV)
#include "ore_sprite.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

buffer::buffer(float* pixels, int length) {
    size = length * (int)sizeof(float);
    checkCudaErrors(cudaMallocHost((void**)&data, size > 0 ? size : 4));
    memcpy(data, pixels, size);
}
buffer::~buffer() {
    if (data) cudaFreeHost(data);
}

namespace {
unsigned lcg(unsigned& s) {  // MSVC rand()
    s = s * 214013u + 2531011u;
    return (s >> 16) & 0x7fff;
}
bool load_ppm(const std::string& file, int& w, int& h, std::vector<unsigned char>& rgb) {
    FILE* f = fopen(file.c_str(), "rb");
    if (!f) return false;
    char magic[3] = {0, 0, 0};
    int maxv = 0;
    bool ok = fscanf(f, "%2s", magic) == 1 && strcmp(magic, "P6") == 0;
    // skip comments
    auto skip = [&]() {
        int c;
        while ((c = fgetc(f)) != EOF) {
            if (c == '#') {
                while ((c = fgetc(f)) != EOF && c != '\n') {}
            } else if (c > ' ') {
                ungetc(c, f);
                break;
            }
        }
    };
    if (ok) {
        skip();
        ok = fscanf(f, "%d", &w) == 1;
        skip();
        ok = ok && fscanf(f, "%d", &h) == 1;
        skip();
        ok = ok && fscanf(f, "%d", &maxv) == 1 && maxv == 255 && w > 0 && h > 0;
        fgetc(f);
    }
    if (ok) {
        rgb.resize((size_t)w * h * 3);
        ok = fread(rgb.data(), 1, rgb.size(), f) == rgb.size();
    }
    fclose(f);
    return ok;
}
// uncompressed 24-bit BMP (bottom-up or top-down), the simplest "real image file" a Windows tool chain writes
bool load_bmp(const std::string& file, int& w, int& h, std::vector<unsigned char>& rgb) {
    FILE* f = fopen(file.c_str(), "rb");
    if (!f) return false;
    unsigned char hdr[54];
    bool ok = fread(hdr, 1, 54, f) == 54 && hdr[0] == 'B' && hdr[1] == 'M';
    auto u32 = [&](int o) { return (unsigned)hdr[o] | ((unsigned)hdr[o + 1] << 8) | ((unsigned)hdr[o + 2] << 16) | ((unsigned)hdr[o + 3] << 24); };
    int bw = 0, bh = 0;
    unsigned off = 0;
    if (ok) {
        off = u32(10);
        bw = (int)u32(18);
        bh = (int)u32(22);
        const unsigned bpp = (unsigned)hdr[28] | ((unsigned)hdr[29] << 8), comp = u32(30);
        ok = bpp == 24 && comp == 0 && bw > 0 && bh != 0;
    }
    if (ok) {
        const bool bottom_up = bh > 0;
        w = bw;
        h = bottom_up ? bh : -bh;
        const size_t stride = ((size_t)w * 3 + 3) & ~(size_t)3;
        std::vector<unsigned char> row(stride);
        rgb.resize((size_t)w * h * 3);
        ok = fseek(f, (long)off, SEEK_SET) == 0;
        for (int y = 0; ok && y < h; y++) {
            ok = fread(row.data(), 1, stride, f) == stride;
            const int dst = bottom_up ? h - 1 - y : y;  // plane row 0 = top of the image (cv::imread order)
            for (int x = 0; ok && x < w; x++) {
                unsigned char* p = &rgb[((size_t)dst * w + x) * 3];
                p[0] = row[3 * x + 2];  // BMP stores B,G,R like cv::Vec3b (Sprite.cpp:44-46 reads r = pixel[2])
                p[1] = row[3 * x + 1];
                p[2] = row[3 * x + 0];
            }
        }
    }
    fclose(f);
    return ok;
}
void procedural(const std::string& spec, int& w, int& h, std::vector<unsigned char>& rgb) {
    char kind[32] = {0};
    int a = 0, b = 0, c = 0;
    if (sscanf(spec.c_str(), "proc:%31[^:]:%d:%d:%d", kind, &a, &b, &c) != 4 || a <= 0 || b <= 0) {
        w = h = 1;
        rgb.assign(3, 128);
        return;
    }
    w = a;
    h = b;
    rgb.resize((size_t)w * h * 3);
    if (strcmp(kind, "checker") == 0) {
        const int cells = c > 0 ? c : 16;
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                const int on = ((x * cells / w) + (y * cells / h)) & 1;
                unsigned char* p = &rgb[((size_t)y * w + x) * 3];
                p[0] = (unsigned char)(40 + 200 * on);
                p[1] = (unsigned char)(220 - 180 * on);
                p[2] = (unsigned char)(60 + 120 * on);
            }
    } else {  // smooth: same construction as scene.smooth_texture in the Python harness
        unsigned s = (unsigned)c;
        double ph[6];
        for (double& v : ph) v = lcg(s) / 32768.0 * 2 * M_PI;
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                const double u = (double)x / w, v = (double)y / h;
                unsigned char* p = &rgb[((size_t)y * w + x) * 3];
                for (int ch = 0; ch < 3; ch++) {
                    double f = 0.55 + 0.20 * sin(2 * M_PI * ((ch + 1) * u + v) + ph[ch]) +
                               0.15 * cos(2 * M_PI * (u - (ch + 2) * v) + ph[3 + ch]);
                    double q = nearbyint(f * 255);
                    p[ch] = (unsigned char)(q < 0 ? 0 : (q > 255 ? 255 : q));
                }
            }
    }
}
}  // namespace

sprite::sprite(std::string file) {
    std::vector<unsigned char> rgb;
    int w = 0, h = 0;
    if (file.rfind("proc:", 0) == 0) {
        procedural(file, w, h, rgb);
    } else if (!(load_ppm(file, w, h, rgb) || load_bmp(file, w, h, rgb))) {
        // The reference decodes any OpenCV-readable image (Sprite.cpp:30); this shim reads binary PPM (P6) and
        // 24-bit BMP only.  An unreadable file is an error in the shim's own convention (memManager.cpp:3-11:
        // message on stderr, exit(99)) - never a silent stand-in texture.
        fprintf(stderr, "sprite: cannot read '%s' (accepted: binary PPM (P6), 24-bit uncompressed BMP, or a proc: spec)\n",
                file.c_str());
        exit(99);
    }
    width = w;
    height = h;
    std::vector<float> r((size_t)w * h), g((size_t)w * h), b((size_t)w * h);
    for (size_t i = 0; i < (size_t)w * h; i++) {
        r[i] = (float)rgb[3 * i + 0] / 255;  // Sprite.cpp:44-46
        g[i] = (float)rgb[3 * i + 1] / 255;
        b[i] = (float)rgb[3 * i + 2] / 255;
    }
    rBuff = new buffer(r.data(), w * h);
    gBuff = new buffer(g.data(), w * h);
    bBuff = new buffer(b.data(), w * h);
}
sprite::~sprite() {
    delete rBuff;
    delete gBuff;
    delete bBuff;
}
// C hook for tests: loads `file` through the sprite constructor and copies the planes out
extern "C" int ore_host_sprite_load(const char* file, int* w, int* h, float* r, float* g, float* b, int cap) {
    sprite* s = new sprite(std::string(file));
    *w = s->width;
    *h = s->height;
    const int n = s->width * s->height;
    int rc = n <= cap ? 0 : 1;
    if (!rc) {
        memcpy(r, s->rBuff->data, sizeof(float) * n);
        memcpy(g, s->gBuff->data, sizeof(float) * n);
        memcpy(b, s->bBuff->data, sizeof(float) * n);
    }
    delete s;
    return rc;
}
int sprite::getBytes() { return (int)sizeof(float) * width * height * 3; }
int sprite::getWidth() { return this->width - 1; }
int sprite::getHeight() { return this->height - 1; }
