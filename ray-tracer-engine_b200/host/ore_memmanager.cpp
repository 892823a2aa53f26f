// ore_memmanager.cpp - see ore_memmanager.h (reference: /root/reference/memManager.cpp:3-22)
#include "ore_memmanager.h"

#include <cstdlib>

void check_cuda(cudaError_t result, char const* const func, const char* const file, int const line) {
    if (result) {
        std::cerr << "CUDA error = " << static_cast<unsigned int>(result) << " at " << file << ":" << line << " '"
                  << func << "' \n";
        cudaDeviceReset();
        exit(99);
    }
}

void check_ore(int status, const char* what, const char* detail, const char* file, int line) {
    if (status) {
        std::cerr << "CUDA error = " << static_cast<unsigned int>(status) << " at " << file << ":" << line << " '"
                  << what << "' " << (detail ? detail : "") << "\n";
        cudaDeviceReset();
        exit(99);
    }
}

void* memManager::operator new(size_t len) {
    void* ptr = nullptr;
    checkCudaErrors(cudaMallocHost(&ptr, len ? len : 1));  // pinned: direct source for async uploads
    return ptr;
}

void memManager::operator delete(void* ptr) {
    if (ptr) checkCudaErrors(cudaFreeHost(ptr));
}
