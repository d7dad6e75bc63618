// zone_sweep.cu - does blocking a huge hash table into L2-sized zones pay on B200?
// A table of T GiB is swept zone by zone (zone = Z MiB); for every zone one kernel (a) streams the NEXT zone
// through L2 with coalesced loads (prefetch) and (b) does this zone's share of E random 32-byte-bucket
// read + counter-update operations (what k1_fused does per k-mer event).  Compared with the same E operations
// spread over the whole table at once.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/zone_sweep.cu -o tools/zone_sweep
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix(uint64_t z){ z+=0x9E3779B97F4A7C15ull; z=(z^(z>>30))*0xBF58476D1CE4E5B9ull; z=(z^(z>>27))*0x94D049BB133111EBull; return z^(z>>31);}

// ops on [base, base+n) slots (u64), n_ops total over the grid; MODE 0 load+CAS, 1 load+atomicAdd
template<int MODE>
__global__ void __launch_bounds__(256) zone_kernel(unsigned long long *tab, uint64_t base, uint64_t n, uint64_t n_ops, uint64_t seed,
                                                   const ulonglong2 *pf, uint64_t pf_n16, unsigned long long *sink)
{
	const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x, nth = gridDim.x * (uint64_t)blockDim.x;
	unsigned long long acc = 0;
	// (a) prefetch the next zone: coalesced 16-byte loads, L2 only
	for (uint64_t i = tid; i < pf_n16; i += nth) { ulonglong2 v = __ldcg(pf + i); acc += v.x ^ v.y; }
	// (b) this zone's operations, 4 in flight per thread
	for (uint64_t o = tid * 4; o < n_ops; o += nth * 4) {
		uint64_t idx[4]; ulonglong2 a[4], b[4];
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			idx[j] = (mix(seed + o + j) % (n / 4)) * 4; // a 32-byte bucket
			const ulonglong2 *p = (const ulonglong2*)(tab + base + idx[j]);
			a[j] = __ldcg(p); b[j] = __ldcg(p + 1);
		}
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			if (o + j >= n_ops) break;
			const int s = (int)((a[j].x ^ b[j].y ^ idx[j]) & 3);
			unsigned long long *q = tab + base + idx[j] + s;
			const unsigned long long cur = s == 0 ? a[j].x : s == 1 ? a[j].y : s == 2 ? b[j].x : b[j].y;
			if (MODE == 0) acc += atomicCAS(q, cur, cur + 1);
			else { atomicAdd(q, 1ull); acc += cur; }
		}
	}
	if (acc == 0x123456789ull) *sink = acc;
}

int main(int argc, char **argv)
{
	const double T = argc > 1 ? atof(argv[1]) : 48.0;      // table GiB
	const double E = argc > 2 ? atof(argv[2]) : 1.7e9;     // operations per sweep
	const uint64_t n = (uint64_t)(T * (1ull << 30) / 8);
	unsigned long long *tab, *sink;
	if (cudaMalloc(&tab, n * 8) != cudaSuccess) { printf("alloc failed\n"); return 1; }
	cudaMalloc(&sink, 8);
	cudaMemset(tab, 0, n * 8);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	float ms;
	for (int mode = 0; mode < 2; ++mode) {
		// unblocked: the whole table is one zone, no prefetch
		for (int rep = 0; rep < 2; ++rep) {
			cudaEventRecord(e0);
			if (mode == 0) zone_kernel<0><<<148 * 8, 256>>>(tab, 0, n, (uint64_t)E, 17 + rep, nullptr, 0, sink);
			else zone_kernel<1><<<148 * 8, 256>>>(tab, 0, n, (uint64_t)E, 17 + rep, nullptr, 0, sink);
			cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
		}
		printf("%-14s unblocked            : %8.2f ms  %6.2f G ops/s\n", mode ? "load+atomicAdd" : "load+CAS", ms, E / ms / 1e6);
		for (int zi = 3; zi < (argc > 3 ? argc : 4); ++zi) {
			const double zmb = argc > 3 ? atof(argv[zi]) : 32.0;
			for (int pf = 0; pf < (getenv("ZS_PREFETCH") ? 2 : 1); ++pf) {
				const uint64_t zn = (uint64_t)(zmb * (1 << 20) / 8), nz = (n + zn - 1) / zn;
				const uint64_t per = (uint64_t)(E / nz);
				for (int rep = 0; rep < 2; ++rep) {
					cudaEventRecord(e0);
					for (uint64_t z = 0; z < nz; ++z) {
						const uint64_t base = z * zn, len = (base + zn <= n ? zn : n - base) & ~3ull;
						const uint64_t nb = (z + 1) * zn, nlen = z + 1 < nz ? ((nb + zn <= n ? zn : n - nb) & ~3ull) : 0;
						const ulonglong2 *pfp = pf && nlen ? (const ulonglong2*)(tab + nb) : nullptr;
						if (mode == 0) zone_kernel<0><<<148 * 4, 256>>>(tab, base, len, per, z * 1000003ull + rep, pfp, pfp ? nlen / 2 : 0, sink);
						else zone_kernel<1><<<148 * 4, 256>>>(tab, base, len, per, z * 1000003ull + rep, pfp, pfp ? nlen / 2 : 0, sink);
					}
					cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
				}
				printf("%-14s zones of %3.0f MiB %s: %8.2f ms  %6.2f G ops/s  (%llu zones, %llu ops each)\n", mode ? "load+atomicAdd" : "load+CAS",
				       zmb, pf ? "+prefetch" : "         ", ms, (double)per * nz / ms / 1e6, (unsigned long long)nz, (unsigned long long)per);
			}
		}
	}
	cudaError_t err = cudaDeviceSynchronize();
	if (err != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(err));
	return 0;
}
