// Stand-in for the CUDA runtime header so the reference's kernel.cu / memManager.cpp
// (read in place from /root/reference, never copied into the repo) compile as plain
// host C++ with g++.  TEST INFRASTRUCTURE ONLY (oracle/_ref build) - never on the
// product path.  Qualifiers are erased; managed/device allocations become calloc.
#pragma once
#include <cstdlib>
#include <cstring>
#include <cstddef>

#define __device__
#define __host__
#define __global__

typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1,
                      cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };

static inline cudaError_t cudaMallocManaged(void** p, size_t n) {
    *p = calloc(n ? n : 1, 1);
    return *p ? 0 : 2;
}
static inline cudaError_t cudaMalloc(void** p, size_t n) { return cudaMallocManaged(p, n); }
static inline cudaError_t cudaFree(void* p) { free(p); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) {
    memcpy(d, s, n);
    return 0;
}
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaDeviceReset() { return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }

struct uint3 { unsigned int x, y, z; };
struct dim3 {
    unsigned int x, y, z;
    dim3(unsigned int a = 1, unsigned int b = 1, unsigned int c = 1) : x(a), y(b), z(c) {}
};
// one "CUDA thread" per host call; set by the driver before each rayTrace() call
extern thread_local uint3 threadIdx;
extern thread_local uint3 blockIdx;
extern thread_local dim3 blockDim;
