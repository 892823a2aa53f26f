// headless_window.cpp - replaces the Win32 window (window.cpp) on the benchmark path: provides the
// window.h callbacks and a main() that drives onStart()/update() like wWinMain (window.cpp:57-84) does,
// with a scripted camera orbit instead of the message pump, and writes the last frame as a PPM.
//
//   ore_headless [width height frames spheres out.ppm mesh.obj|- pipelined(0|1) present(copy|pointer)]
//
// present = copy (default): setPixelBuff memcpy's the frame like window.cpp:130-132 (one host thread: ~10 GB/s, i.e. 13 ms
// for an 8K frame - the reference's window, not this path, then bounds the frame rate).  present = pointer: the "window"
// keeps the pointer it is handed (valid until the frame after next in pipelined mode) and reads the pixels when it needs
// them - what a consumer that encodes or uploads the frame would do.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ore_host.h"
#include "ore_window_callbacks.h"

namespace {
struct RenderState {  // window.cpp:8-17
    int width = 0, height = 0;
    std::vector<unsigned int> buffmemory;
    const unsigned int* front = nullptr;   // present = pointer: the last frame handed over
    bool copy = true;
} render;
}

int getScreenHeight() { return render.height; }
int getScreenWidth() { return render.width; }
void setPixelBuff(unsigned int* pixels) {  // window.cpp:130-132
    if (render.copy)
        memcpy(render.buffmemory.data(), pixels, sizeof(unsigned int) * render.width * render.height);
    else
        render.front = pixels;
}

int main(int argc, char** argv) {
    render.width = argc > 1 ? atoi(argv[1]) : 640;
    render.height = argc > 2 ? atoi(argv[2]) : 480;
    const int frames = argc > 3 ? atoi(argv[3]) : 8;
    const int spheres = argc > 4 ? atoi(argv[4]) : 64;
    const char* out = argc > 5 ? argv[5] : nullptr;
    if (argc > 6 && strcmp(argv[6], "-") != 0) oreSetMeshFile(argv[6]);
    const int pipelined = argc > 7 ? atoi(argv[7]) : 0;
    render.copy = !(argc > 8 && strcmp(argv[8], "pointer") == 0);
    oreConfigurePresentation(pipelined);
    render.buffmemory.assign((size_t)render.width * render.height, 0u);

    oreConfigureScene(spheres, 1u, nullptr, nullptr);
    onStart();
    update();  // warm-up (context buffers, module load)
    oreFlush();
    const auto t0 = std::chrono::steady_clock::now();
    for (int f = 0; f < frames; f++) {
        // orbit about the cube centre (5,5,5), radius 12, facing the centre (see scene.orbit_camera)
        const double yaw = 180.0 + 360.0 * f / (frames > 1 ? frames : 1), pitch = 15.0;
        const double yr = yaw * M_PI / 180, pr = pitch * M_PI / 180;
        oreSetCamera((float)(5 - 12 * cos(pr) * sin(yr)), (float)(5 + 12 * sin(pr)), (float)(5 - 12 * cos(pr) * cos(yr)),
                     (float)yaw, (float)pitch);
        update();
    }
    oreFlush();  // pipelined mode: the last frame is still in flight
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (!render.copy && render.front)   // (outside the timed region: checksum / PPM of the last frame)
        memcpy(render.buffmemory.data(), render.front, sizeof(unsigned int) * render.width * render.height);
    unsigned long long sum = 0;
    for (unsigned v : render.buffmemory) sum += v;
    printf("{\"width\": %d, \"height\": %d, \"frames\": %d, \"spheres\": %d, \"ms_per_frame\": %.3f, \"mrays_per_s\": %.1f, "
           "\"checksum\": %llu, \"pipelined\": %d, \"present\": \"%s\"}\n",
           render.width, render.height, frames, spheres, sec / frames * 1e3,
           (double)render.width * render.height * frames / sec / 1e6, sum, pipelined, render.copy ? "copy" : "pointer");
    if (out) {
        FILE* fp = fopen(out, "wb");
        if (fp) {
            fprintf(fp, "P6\n%d %d\n255\n", render.width, render.height);
            for (int y = render.height - 1; y >= 0; y--)  // row 0 is the bottom scanline (window.cpp:43)
                for (int x = 0; x < render.width; x++) {
                    const unsigned p = render.buffmemory[(size_t)y * render.width + x];
                    const unsigned char rgb[3] = {(unsigned char)(p >> 16), (unsigned char)(p >> 8), (unsigned char)p};
                    fwrite(rgb, 1, 3, fp);
                }
            fclose(fp);
        }
    }
    oreShutdown();
    return 0;
}
