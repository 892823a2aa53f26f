// sprite_raw.cpp - stand-in for the reference's Sprite.cpp (which decodes image files
// with OpenCV, absent here).  TEST INFRASTRUCTURE ONLY (oracle/_ref host build).
//
// Keeps the reference's sprite format exactly (sprite.h:11-47, Sprite.cpp:13-52):
// three planar float planes r,g,b, row-major y*width+x, each a `buffer{data,size}`.
// Instead of cv::imread the "file name" is a key into a registry of raw planes the
// driver registers before constructing the sprite.
//
// The reference indexes one row (+1 texel) past the end of a plane at the poles
// ((int)(ty*maxY) can equal maxY, kernel.cu:1402-1403,1653; sky :1157-1160) - undefined
// behaviour there.  The harness DEFINES it: every plane gets width+1 trailing floats
// equal to its last texel, i.e. the out-of-range read behaves like an index clamp.
// The CUDA path clamps the index, so both read identical values.
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include "sprite.h"
#include "sprite_raw.h"
#define checkCudaErrors(val) check_cuda((val), #val, __FILE__, __LINE__)

static std::map<std::string, raw_image>& registry() {
    static std::map<std::string, raw_image> r;
    return r;
}
void sprite_raw_register(const std::string& key, const raw_image& img) { registry()[key] = img; }
void sprite_raw_unregister(const std::string& key) { registry().erase(key); }

// same body as the reference's buffer ctor contract: size in bytes, managed copy
buffer::buffer(float* pixels, int length) {
    size = length * (int)sizeof(float);
    checkCudaErrors(cudaMallocManaged((void**)&data, size));
    memcpy(data, pixels, size);
}

static buffer* padded_plane(const float* src, int w, int h) {
    const size_t n = (size_t)w * h, pad = (size_t)w + 1;
    std::vector<float> tmp(n + pad);
    for (size_t i = 0; i < n; i++) tmp[i] = src[i];
    for (size_t i = 0; i < pad; i++) tmp[n + i] = src[n - 1];
    return new buffer(tmp.data(), (int)(n + pad));
}

sprite::sprite(std::string file) {
    auto it = registry().find(file);
    static const float grey[1] = {0.5f};
    raw_image img{1, 1, grey, grey, grey};   // unknown key (the reference's C:\ paths)
    if (it != registry().end()) img = it->second;
    width = img.w;
    height = img.h;
    rBuff = padded_plane(img.r, img.w, img.h);
    gBuff = padded_plane(img.g, img.w, img.h);
    bBuff = padded_plane(img.b, img.w, img.h);
}
int sprite::getBytes() { return (int)sizeof(float) * width * height * 3; }
int sprite::getWidth() { return this->width - 1; }
int sprite::getHeight() { return this->height - 1; }

void sprite_raw_free(sprite* s) {
    if (!s) return;
    buffer* b[3] = {s->rBuff, s->gBuff, s->bBuff};
    for (buffer* p : b) {
        cudaFree(p->data);
        delete p;
    }
    delete s;
}
