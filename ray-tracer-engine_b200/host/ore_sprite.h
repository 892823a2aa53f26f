// ore_sprite.h - the reference's texture format, unchanged in layout (/root/reference/sprite.h:11-47):
// three planar float planes r,g,b (value = byte/255, row-major y*width+x, Sprite.cpp:41-46), each a
// `buffer{float* data; int size /*bytes*/}`; `sprite{rBuff,gBuff,bBuff,width,height}`.
// Planes live in pinned host memory (memManager) and are uploaded once by ore_set_texture/ore_set_sky.
#pragma once
#include <string>

#include "ore_memmanager.h"

class buffer : public memManager {
public:
    float* data;
    int size;
    buffer(float* pixels, int length);
    ~buffer();
};

class sprite : public memManager {
public:
    // `file` is a binary PPM (P6) or an uncompressed 24-bit BMP path, or a procedural source:
    //   "proc:smooth:<w>:<h>:<seed>"  low-frequency sinusoid, "proc:checker:<w>:<h>:<cells>"
    // (the reference decodes image files with OpenCV, Sprite.cpp:30; OpenCV is out of scope here)
    sprite(std::string file);
    ~sprite();
    int getBytes();
    int getWidth();
    int getHeight();

    buffer* rBuff;
    buffer* gBuff;
    buffer* bBuff;
    int width;
    int height;
};
