// The three callbacks the kernel translation unit uses from the window layer
// (declared in /root/reference/window.h:7-11, implemented by window.cpp:86-91,130-132).
// In a drop-in build the maintainer's own window.h provides them; the headless harness
// (headless_window.cpp) implements them without Win32.
#pragma once

int getScreenWidth();
int getScreenHeight();
void setPixelBuff(unsigned int* pixels);  // copies width*height packed 0x00RRGGBB pixels from a host-readable pointer
