// ore_host.h - entry points of the host shim.
// onStart()/update() are the reference's kernel-launch entry points with unchanged signatures
// (/root/reference/kernel.cuh:3-4): window.cpp calls them from wWinMain (window.cpp:71,80).
#pragma once

void onStart();
void update();

// ---- additions for headless use (not in the reference) ----
// scripted camera instead of GetKeyState polling (kernel.cu:1716-1759): position + yaw/pitch in degrees
void oreSetCamera(float x, float y, float z, float yaw_deg, float pitch_deg);
// scene size / seed for the reference sphere generator (kernel.cu:1189-1191); call before onStart()
void oreConfigureScene(int sphere_count, unsigned seed, const char* texture, const char* sky);
// triangle mesh for onStart(): an OBJ file in the reference's dialect (kernel.cu:1706); call before onStart()
void oreSetMeshFile(const char* obj_path);
// Presentation mode of update().  0 (default): the reference's synchronous semantics - update() returns after the
// frame has been handed to setPixelBuff (kernel.cu:1786-1788).  1: pipelined - update() enqueues frame f (kernels +
// asynchronous copy into one of two pinned host frames) and hands setPixelBuff frame f-1, whose copy overlapped
// frame f's kernels; oreFlush() presents what is still in flight.  Call before the first update().
void oreConfigurePresentation(int pipelined);
void oreFlush();
// release the render context (the reference never frees its globals)
void oreShutdown();
