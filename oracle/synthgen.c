/* synthgen.c - CPU generator of the seeded synthetic read stream (the specification is
 * yak_b200/synth.py, the device twin is csrc/extras.cu synth_reads_kernel).  TEST / BENCH INFRASTRUCTURE ONLY:
 * bench.py's reference arm writes its input with this program so that the reference's process never loads
 * the product library.  Genome bases are evaluated on the fly (counter-based), so a 3 Gbp genome costs nothing.
 *
 *   synthgen <seed_g> <G> <seed_r> <first> <n_reads> <L> <err> <n_pct> <fmt> <k> <out> [threads]
 *     fmt 1: ">r\nSEQ\n"   fmt 2: "@r\nSEQ\n+\nQUAL\n" (quality 'I')   - fixed record size
 *   prints the number of k-mer events (windows of k bases without N) on stdout.
 */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include <pthread.h>
#include <fcntl.h>
#include <unistd.h>
#include <sys/mman.h>

#define GMUL 0xD1342543DE82EF95ull

static inline uint64_t smix(uint64_t z)
{
	z += 0x9E3779B97F4A7C15ull;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

typedef struct {
	uint64_t seed_g, G, seed_r, first, r0, r1, thr, rec;
	int L, n_pct, fmt, k;
	uint8_t *out;
	uint64_t n_events;
} job_t;

static void *worker(void *p)
{
	job_t *j = (job_t*)p;
	const int L = j->L;
	uint64_t ev = 0;
	for (uint64_t i = j->r0; i < j->r1; ++i) {
		uint8_t *o = j->out + i * j->rec;
		const uint64_t a = smix(j->seed_r * GMUL + (j->first + i));
		const uint64_t start = smix(a + 1) % (j->G - L + 1);
		const uint64_t f = smix(a + 2);
		const int rev = f & 1;
		const int has_n = ((f >> 8) % 100) < (uint64_t)j->n_pct;
		const int npos = (int)((f >> 32) % (uint64_t)L);
		int run = 0;
		*o++ = j->fmt == 1 ? '>' : '@'; *o++ = 'r'; *o++ = '\n';
		for (int q = 0; q < L; ++q) {
			const uint64_t gi = rev ? start + (uint64_t)(L - 1 - q) : start + (uint64_t)q;
			uint32_t b = (uint32_t)(smix(j->seed_g * GMUL + gi) >> 62);
			if (rev) b = 3 - b;
			const uint64_t e = smix(a + 16 + (uint64_t)q);
			if ((e & 0xFFFFFF) < j->thr) b = (uint32_t)(e >> 24) & 3;
			uint8_t ch = "ACGT"[b];
			if (has_n && npos == q) ch = 'N';
			*o++ = ch;
			run = ch == 'N' ? 0 : run + 1;
			if (run >= j->k) ++ev;
		}
		*o++ = '\n';
		if (j->fmt == 2) {
			*o++ = '+'; *o++ = '\n';
			memset(o, 'I', L); o += L;
			*o++ = '\n';
		}
	}
	j->n_events = ev;
	return 0;
}

int main(int argc, char **argv)
{
	if (argc < 12) {
		fprintf(stderr, "usage: synthgen seed_g G seed_r first n_reads L err n_pct fmt k out [threads]\n");
		return 1;
	}
	job_t base;
	memset(&base, 0, sizeof(base));
	base.seed_g = strtoull(argv[1], 0, 10); base.G = strtoull(argv[2], 0, 10); base.seed_r = strtoull(argv[3], 0, 10);
	base.first = strtoull(argv[4], 0, 10);
	const uint64_t n_reads = strtoull(argv[5], 0, 10);
	base.L = atoi(argv[6]);
	base.thr = (uint64_t)(atof(argv[7]) * 16777216.0);
	base.n_pct = atoi(argv[8]); base.fmt = atoi(argv[9]); base.k = atoi(argv[10]);
	int nt = argc > 12 ? atoi(argv[12]) : 8;
	if (nt < 1) nt = 1;
	if (nt > 256) nt = 256;
	if (base.fmt != 1 && base.fmt != 2) { fprintf(stderr, "synthgen: fmt must be 1 or 2\n"); return 1; }
	base.rec = base.fmt == 1 ? (uint64_t)base.L + 4 : 2 * (uint64_t)base.L + 7;
	const uint64_t bytes = n_reads * base.rec;
	int fd = open(argv[11], O_RDWR | O_CREAT | O_TRUNC, 0644);
	if (fd < 0) { perror("synthgen: open"); return 1; }
	if (bytes == 0) { close(fd); printf("0\n"); return 0; }
	if (ftruncate(fd, (off_t)bytes) != 0) { perror("synthgen: ftruncate"); return 1; }
	uint8_t *out = (uint8_t*)mmap(0, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
	if (out == MAP_FAILED) { perror("synthgen: mmap"); return 1; }
	pthread_t *th = (pthread_t*)calloc(nt, sizeof(pthread_t));
	job_t *jobs = (job_t*)calloc(nt, sizeof(job_t));
	for (int t = 0; t < nt; ++t) {
		jobs[t] = base;
		jobs[t].out = out;
		jobs[t].r0 = n_reads * (uint64_t)t / nt; jobs[t].r1 = n_reads * (uint64_t)(t + 1) / nt;
		pthread_create(&th[t], 0, worker, &jobs[t]);
	}
	uint64_t ev = 0;
	for (int t = 0; t < nt; ++t) { pthread_join(th[t], 0); ev += jobs[t].n_events; }
	munmap(out, bytes);
	close(fd);
	printf("%llu\n", (unsigned long long)ev);
	free(th); free(jobs);
	return 0;
}
