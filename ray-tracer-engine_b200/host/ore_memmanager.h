// ore_memmanager.h - drop-in for the reference's memManager.h (same two public names, same signatures:
// /root/reference/memManager.h:11-18) on top of the B200 render library.
//
// The reference backs `operator new` with cudaMallocManaged + cudaDeviceSynchronize and lets the kernel
// chase pointers into that managed memory (memManager.cpp:12-22).  Here scene DATA lives in structure-of-
// arrays device buffers owned by the render context (include/ore_render.h), fed through pinned-host
// staging; objects deriving from memManager are host-side descriptors only, so `operator new` hands out
// PINNED host memory (cudaMallocHost): uploads from them need no extra staging copy.
//
// Error convention kept from the reference (memManager.cpp:3-11): print
// "CUDA error = <n> at <file>:<line> '<expr>'", cudaDeviceReset(), exit(99).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <iostream>

void check_cuda(cudaError_t result, char const* const func, const char* const file, int const line);
#define checkCudaErrors(val) check_cuda((val), #val, __FILE__, __LINE__)

// status codes of the C ABI follow the same print-and-exit(99) convention in the shim layer
void check_ore(int status, const char* what, const char* detail, const char* file, int line);

class memManager {
public:
    void* operator new(size_t len);
    void operator delete(void* ptr);
};
