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

// Device memory is cached by the library: a freed block (a 16 GiB filter, a table that was regrown,
// per-chunk scratch) is kept and handed to the next request of about that size, because cudaFree of
// gigabytes stalls the host for tens to hundreds of milliseconds (seconds, at times) and yak_count
// builds and drops a table per pass.  Idle blocks are bounded (YAKB_CACHE_GB, default 48) and all go
// back to the driver when an allocation fails.  Both calls keep the synchronous meaning of
// cudaMalloc / cudaFree.  YAKB_NO_POOL=1 turns the cache off.
void *dev_alloc(size_t bytes);
void dev_free(void *p);
size_t dev_pool_idle(); // bytes held idle (cudaMemGetInfo counts them as taken)
void dev_trim();        // give every idle block back

// grow-only device buffer
struct DBuf {
	void *p = nullptr; size_t cap = 0;
	void *need(size_t bytes);
	template<class T> T *as(size_t n) { return (T*)need(n * sizeof(T)); }
	void release();
};

} // namespace yakb
