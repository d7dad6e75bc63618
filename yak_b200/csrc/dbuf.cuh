// dbuf.cuh - CUDA error handling and the grow-only device buffer used for all per-chunk scratch
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <stdexcept>
#include <cuda_runtime.h>

namespace yakb {

struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };
void cuda_fail(cudaError_t e, const char *what, const char *file, int line);
#define YAKB_CUDA(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) ::yakb::cuda_fail(e__, #x, __FILE__, __LINE__); } while (0)

// grow-only device buffer
struct DBuf {
	void *p = nullptr; size_t cap = 0;
	void *need(size_t bytes);
	template<class T> T *as(size_t n) { return (T*)need(n * sizeof(T)); }
	void release();
};

} // namespace yakb
