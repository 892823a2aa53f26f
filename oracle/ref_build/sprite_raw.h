// registry of raw planar images standing in for image files (see sprite_raw.cpp)
#pragma once
#include <string>
struct raw_image { int w, h; const float *r, *g, *b; };
class sprite;
void sprite_raw_register(const std::string& key, const raw_image& img);
void sprite_raw_unregister(const std::string& key);
void sprite_raw_free(sprite* s);
