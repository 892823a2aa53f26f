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
// release the render context (the reference never frees its globals)
void oreShutdown();
