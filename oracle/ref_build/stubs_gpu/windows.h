// stand-in for <windows.h> in the nvcc build of the reference kernel (oracle/_ref/libref_sm100.so):
// checkKey() polls GetKeyState (kernel.cu:1723-1757); headless = no key is ever down.
#pragma once
#define VK_SHIFT 0x10
#define VK_SPACE 0x20
static inline short GetKeyState(int) { return 0; }
