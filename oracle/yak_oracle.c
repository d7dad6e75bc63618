/* yak_oracle.c - CPU restatement of lh3/yak's k-mer count / lookup path (see yak_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY: the checker, never the product.  Sequential on purpose: the
 * reference's result is independent of -t and -K, so one thread walking the
 * events of every sub-table in file order is the specification.
 */
#include "yak_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include <zlib.h>

/* ------------------------------------------------------------------ hashes */

/* yak-priv.h:11-21 - the 7-step invertible mix, every add masked to 2k bits */
uint64_t yo_hash64(uint64_t key, uint64_t mask)
{
	key = (~key + (key << 21)) & mask;
	key ^= key >> 24;
	key = (key * 265) & mask;          /* key + key<<3 + key<<8 */
	key ^= key >> 14;
	key = (key * 21) & mask;           /* key + key<<2 + key<<4 */
	key ^= key >> 28;
	key = (key + (key << 31)) & mask;
	return key;
}

/* yak-priv.h:23-33 - same mix without masking */
uint64_t yo_hash64_64(uint64_t key)
{
	key = ~key + (key << 21);
	key ^= key >> 24;
	key *= 265;
	key ^= key >> 14;
	key *= 21;
	key ^= key >> 28;
	key += key << 31;
	return key;
}

/* yak-priv.h:35-39 - strand picked by the high-bit planes only; ties take the reverse (quirk Q5) */
uint64_t yo_hash_long(const uint64_t x[4])
{
	int rev = !(x[1] < x[3]);
	return yo_hash64_64(x[rev * 2]) + yo_hash64_64(x[rev * 2 + 1]);
}

/* yak-priv.h:41-68 - inverse of yo_hash64 */
uint64_t yo_hash64_inv(uint64_t key, uint64_t mask)
{
	uint64_t t;
	t = key - (key << 31);
	key = (key - (t << 31)) & mask;
	t = key ^ key >> 28;
	key ^= t >> 28;
	key = (key * 14933078535860113213ull) & mask;   /* 21^-1 mod 2^64 */
	t = key ^ key >> 14;
	t = key ^ t >> 14;
	t = key ^ t >> 14;
	key ^= t >> 14;
	key = (key * 15244667743933553977ull) & mask;   /* 265^-1 mod 2^64 */
	t = key ^ key >> 24;
	key ^= t >> 24;
	t = ~key;
	t = ~(key - (t << 21));
	t = ~(key - (t << 21));
	key = ~(key - (t << 21)) & mask;
	return key;
}

/* khashl.h:98 with htab.c:9-10 and the khint_t (uint32) return type at khashl.h:262:
 * only the low 32 bits of (key >> 10) reach the Fibonacci multiply (quirk Q2). */
uint32_t yo_slot_home(uint64_t stored_key, uint32_t bits)
{
	uint32_t h = (uint32_t)(stored_key >> YO_COUNTER_BITS);
	return (uint32_t)(h * 2654435769u) >> (32 - bits);
}

/* misc.c:4-21: A/a C/c G/g T/t U/u -> 0..3, raw bytes 0..3 -> 0..3, everything else 4 */
const unsigned char yo_nt4[256] = {
	0,1,2,3, 4,4,4,4, 4,4,4,4, 4,4,4,4,  4,4,4,4, 4,4,4,4, 4,4,4,4, 4,4,4,4,
	4,4,4,4, 4,4,4,4, 4,4,4,4, 4,4,4,4,  4,4,4,4, 4,4,4,4, 4,4,4,4, 4,4,4,4,
	4,0,4,1, 4,4,4,2, 4,4,4,4, 4,4,4,4,  4,4,4,4, 3,3,4,4, 4,4,4,4, 4,4,4,4,
	4,0,4,1, 4,4,4,2, 4,4,4,4, 4,4,4,4,  4,4,4,4, 3,3,4,4, 4,4,4,4, 4,4,4,4,
	4,4,4,4, 4,4,4,4, 4,4,4,4, 4,4,4,4,  4,4,4,4, 4,4,4,4, 4,4,4,4, 4,4,4,4,
	4,4,4,4, 4,4,4,4, 4,4,4,4, 4,4,4,4,  4,4,4,4, 4,4,4,4, 4,4,4,4, 4,4,4,4,
	4,4,4,4, 4,4,4,4, 4,4,4,4, 4,4,4,4,  4,4,4,4, 4,4,4,4, 4,4,4,4, 4,4,4,4,
	4,4,4,4, 4,4,4,4, 4,4,4,4, 4,4,4,4,  4,4,4,4, 4,4,4,4, 4,4,4,4, 4,4,4,4
};

/* ------------------------------------------------------------------ the set */

#define SAME_KMER(a, b) (((a) >> YO_COUNTER_BITS) == ((b) >> YO_COUNTER_BITS)) /* htab.c:9 */

static inline int bit_get(const uint32_t *f, uint32_t i) { return f[i >> 5] >> (i & 31) & 1; }
static inline void bit_set(uint32_t *f, uint32_t i) { f[i >> 5] |= 1u << (i & 31); }
static inline void bit_clr(uint32_t *f, uint32_t i) { f[i >> 5] &= ~(1u << (i & 31)); }
static inline uint32_t flag_words(uint32_t n) { return n < 32 ? 1 : n >> 5; }

uint32_t yo_set_capacity(const yo_set_t *s) { return s->keys ? 1u << s->bits : 0; } /* khashl.h:309 */
int yo_set_used(const yo_set_t *s, uint32_t i) { return bit_get(s->used, i); }

/* khashl.h:152-195: round the request up to a power of two (min 4), refuse if it cannot hold the
 * current keys at 3/4 load, then re-place every key IN PLACE: walk old slots upward; a lifted key
 * goes to the first slot free in the NEW occupancy map starting at its new home; if that slot
 * still holds a not-yet-moved old key the two swap and the evicted key continues. */
int yo_set_resize(yo_set_t *s, uint32_t request)
{
	uint32_t lg = 0, t = request, new_bits, new_n, old_n, new_mask, j;
	uint32_t *occ;
	while ((t >>= 1) != 0) ++lg;
	if (request & (request - 1)) ++lg;
	new_bits = lg > 2 ? lg : 2;
	new_n = 1u << new_bits;
	if (s->count > (new_n >> 1) + (new_n >> 2)) return 0;
	occ = (uint32_t*)calloc(flag_words(new_n), sizeof(uint32_t));
	old_n = yo_set_capacity(s);
	if (old_n < new_n) s->keys = (uint64_t*)realloc(s->keys, (size_t)new_n * sizeof(uint64_t));
	new_mask = new_n - 1;
	for (j = 0; j != old_n; ++j) {
		uint64_t key;
		if (!bit_get(s->used, j)) continue;
		key = s->keys[j];
		bit_clr(s->used, j);
		for (;;) {
			uint32_t i = yo_slot_home(key, new_bits);
			while (bit_get(occ, i)) i = (i + 1) & new_mask;
			bit_set(occ, i);
			if (i < old_n && bit_get(s->used, i)) {
				uint64_t evicted = s->keys[i];
				s->keys[i] = key;
				key = evicted;
				bit_clr(s->used, i);
			} else {
				s->keys[i] = key;
				break;
			}
		}
	}
	if (old_n > new_n) s->keys = (uint64_t*)realloc(s->keys, (size_t)new_n * sizeof(uint64_t));
	free(s->used);
	s->used = occ;
	s->bits = new_bits;
	return 0;
}

/* khashl.h:137-150 */
uint32_t yo_set_get(const yo_set_t *s, uint64_t key)
{
	uint32_t n, mask, i, start;
	if (s->keys == 0) return 0;
	n = 1u << s->bits, mask = n - 1;
	i = start = yo_slot_home(key, s->bits);
	while (bit_get(s->used, i) && !SAME_KMER(s->keys[i], key)) {
		i = (i + 1) & mask;
		if (i == start) return n;
	}
	return bit_get(s->used, i) ? i : n;
}

/* khashl.h:197-221: the load test comes BEFORE the lookup (quirk Q3) */
uint32_t yo_set_put(yo_set_t *s, uint64_t key, int *absent)
{
	uint32_t n = yo_set_capacity(s), mask, i, start;
	if (s->count >= (n >> 1) + (n >> 2)) {
		yo_set_resize(s, n + 1);
		n = 1u << s->bits;
	}
	mask = n - 1;
	i = start = yo_slot_home(key, s->bits);
	while (bit_get(s->used, i) && !SAME_KMER(s->keys[i], key)) {
		i = (i + 1) & mask;
		if (i == start) break;
	}
	if (!bit_get(s->used, i)) {
		s->keys[i] = key;
		bit_set(s->used, i);
		++s->count;
		*absent = 1;
	} else *absent = 0;
	return i;
}

static void set_free(yo_set_t *s) { free(s->keys); free(s->used); s->keys = 0; s->used = 0; s->bits = s->count = 0; }

/* ------------------------------------------------------------------ bloom */

/* bbf.c:5-18 */
yo_bloom_t *yo_bloom_init(int n_shift, int n_hashes)
{
	yo_bloom_t *b;
	if (n_shift + 9 > 64 || n_shift < 9) return 0;
	b = (yo_bloom_t*)calloc(1, sizeof(*b));
	b->n_shift = n_shift, b->n_hashes = n_hashes;
	b->b = (uint8_t*)calloc((size_t)1 << (n_shift - 3), 1);
	return b;
}

void yo_bloom_destroy(yo_bloom_t *b) { if (b) { free(b->b); free(b); } }

/* bbf.c:25-42: one 512-bit block chosen by the low bits; n_hashes bits by double hashing */
int yo_bloom_insert(yo_bloom_t *b, uint64_t hash)
{
	int sh = b->n_shift - 9, i, z, cnt = 0;
	uint8_t *blk = b->b + ((hash & ((1ULL << sh) - 1)) << 6);
	int h1 = hash >> sh & 511, h2 = hash >> b->n_shift & 511;
	if ((h2 & 31) == 0) h2 = (h2 + 1) & 511;
	for (i = 0, z = h1; i < b->n_hashes; ++i, z = (z + h2) & 511) {
		uint8_t m = 1u << (z & 7);
		cnt += (blk[z >> 3] & m) != 0;
		blk[z >> 3] |= m;
	}
	return cnt;
}

/* ------------------------------------------------------------------ count table */

/* htab.c:13-29 */
yo_ch_t *yo_ch_init(int k, int pre, int n_hash, int n_shift)
{
	yo_ch_t *h;
	int i, n;
	if (pre < YO_COUNTER_BITS) return 0;
	h = (yo_ch_t*)calloc(1, sizeof(*h));
	h->k = k, h->pre = pre;
	n = 1 << pre;
	h->h = (yo_set_t*)calloc(n, sizeof(yo_set_t));
	h->b = (yo_bloom_t**)calloc(n, sizeof(yo_bloom_t*));
	if (n_hash > 0 && n_shift > pre) {
		h->n_hash = n_hash, h->n_shift = n_shift;
		for (i = 0; i < n; ++i) h->b[i] = yo_bloom_init(n_shift - pre, n_hash);
	}
	return h;
}

/* htab.c:31-39 */
void yo_ch_destroy_bf(yo_ch_t *h)
{
	int i;
	for (i = 0; i < 1 << h->pre; ++i) { yo_bloom_destroy(h->b[i]); h->b[i] = 0; }
}

/* htab.c:41-49 */
void yo_ch_destroy(yo_ch_t *h)
{
	int i;
	if (h == 0) return;
	yo_ch_destroy_bf(h);
	for (i = 0; i < 1 << h->pre; ++i) set_free(&h->h[i]);
	free(h->h); free(h->b); free(h);
}

/* htab.c:51-78 */
int yo_ch_insert_list(yo_ch_t *h, int create_new, int n, const uint64_t *a)
{
	uint64_t lowmask = (1ULL << h->pre) - 1;
	int j, n_new = 0, w;
	yo_set_t *g;
	if (n == 0) return 0;
	w = (int)(a[0] & lowmask);
	g = &h->h[w];
	for (j = 0; j < n; ++j) {
		uint64_t x = a[j] >> h->pre;
		if ((a[j] & lowmask) != (uint64_t)w) continue;     /* htab.c:61: foreign elements skipped */
		if (create_new) {
			int take = 1, absent;
			uint32_t i;
			if (h->b[w]) take = yo_bloom_insert(h->b[w], x) == h->n_hash;
			if (!take) continue;
			i = yo_set_put(g, x << YO_COUNTER_BITS, &absent);
			n_new += absent;
			if ((g->keys[i] & YO_MAX_COUNT) < YO_MAX_COUNT) ++g->keys[i];
		} else {
			uint32_t i = yo_set_get(g, x << YO_COUNTER_BITS);
			if (i != yo_set_capacity(g) && (g->keys[i] & YO_MAX_COUNT) < YO_MAX_COUNT) ++g->keys[i];
		}
	}
	return n_new;
}

/* a stream of events of mixed sub-tables in file order, one by one (what count.c:129-143 does
 * per chunk through the per-sub-table lists); h->tot maintained like count.c:138 */
void yo_ch_insert_events(yo_ch_t *h, int create_new, int64_t n, const uint64_t *a)
{
	int64_t i;
	for (i = 0; i < n; ++i) h->tot += yo_ch_insert_list(h, create_new, 1, &a[i]);
}

/* htab.c:93-100 */
int yo_ch_get(const yo_ch_t *h, uint64_t x)
{
	const yo_set_t *g = &h->h[x & ((1ULL << h->pre) - 1)];
	uint32_t i = yo_set_get(g, x >> h->pre << YO_COUNTER_BITS);
	return i == yo_set_capacity(g) ? -1 : (int)(g->keys[i] & YO_MAX_COUNT);
}

/* htab.c:80-91 */
int yo_ch_inc(yo_ch_t *h, uint64_t x)
{
	yo_set_t *g = &h->h[x & ((1ULL << h->pre) - 1)];
	uint32_t i = yo_set_get(g, x >> h->pre << YO_COUNTER_BITS);
	if (i == yo_set_capacity(g)) return -1;
	if ((g->keys[i] & YO_MAX_COUNT) < YO_MAX_COUNT) ++g->keys[i];
	return (int)(g->keys[i] & YO_MAX_COUNT);
}

/* htab.c:116-130 */
void yo_ch_clear(yo_ch_t *h)
{
	int w;
	for (w = 0; w < 1 << h->pre; ++w) {
		yo_set_t *g = &h->h[w];
		uint32_t i, n = yo_set_capacity(g);
		for (i = 0; i < n; ++i)
			if (bit_get(g->used, i)) g->keys[i] &= ~(uint64_t)YO_MAX_COUNT;
	}
}

/* htab.c:136-169 */
void yo_ch_hist(const yo_ch_t *h, int64_t cnt[YO_N_COUNTS])
{
	int w;
	memset(cnt, 0, YO_N_COUNTS * sizeof(int64_t));
	for (w = 0; w < 1 << h->pre; ++w) {
		const yo_set_t *g = &h->h[w];
		uint32_t i, n = yo_set_capacity(g);
		for (i = 0; i < n; ++i)
			if (bit_get(g->used, i)) ++cnt[g->keys[i] & YO_MAX_COUNT];
	}
}

/* htab.c:175-208: fresh set pre-sized to the OLD size, old slots visited upward (quirk Q9) */
void yo_ch_shrink(yo_ch_t *h, int min, int max)
{
	int w;
	if (!(max >= min && max <= YO_MAX_COUNT)) max = YO_MAX_COUNT;
	h->tot = 0;
	for (w = 0; w < 1 << h->pre; ++w) {
		yo_set_t *g = &h->h[w], f;
		uint32_t i, n = yo_set_capacity(g);
		memset(&f, 0, sizeof(f));
		yo_set_resize(&f, g->count);
		for (i = 0; i < n; ++i) {
			int c, absent;
			if (!bit_get(g->used, i)) continue;
			c = (int)(g->keys[i] & YO_MAX_COUNT);
			if (c >= min && c <= max) yo_set_put(&f, g->keys[i], &absent);
		}
		set_free(g);
		*g = f;
		h->tot += g->count;
	}
}

/* htab.c:102-110 */
void yo_ch_tighten(yo_ch_t *h)
{
	int w;
	for (w = 0; w < 1 << h->pre; ++w) {
		yo_set_t *g = &h->h[w];
		if (g->count * 3 < yo_set_capacity(g)) yo_set_resize(g, g->count * 3);
	}
}

/* htab.c:214-235 */
void yo_ch_setcnt(yo_ch_t *h, int cnt)
{
	int w;
	for (w = 0; w < 1 << h->pre; ++w) {
		yo_set_t *g = &h->h[w];
		uint32_t i, n = yo_set_capacity(g);
		for (i = 0; i < n; ++i)
			if (bit_get(g->used, i)) g->keys[i] = (g->keys[i] & ~(uint64_t)YO_MAX_COUNT) | (uint64_t)cnt;
	}
}

/* htab.c:241-285: h1 is consumed */
void yo_ch_merge(yo_ch_t *h0, yo_ch_t *h1, int min, int max, int pre_resize)
{
	int w;
	if (!(max >= min && max <= YO_MAX_COUNT)) max = YO_MAX_COUNT;
	h0->tot = 0;
	for (w = 0; w < 1 << h0->pre; ++w) {
		yo_set_t *g0 = &h0->h[w], *g1 = &h1->h[w];
		uint32_t i, n1 = yo_set_capacity(g1);
		if (pre_resize) {
			uint32_t want = (g0->count + g1->count) * 4 / 3 + 1;
			if (want > yo_set_capacity(g0)) yo_set_resize(g0, want);
		}
		for (i = 0; i < n1; ++i) {
			int c, absent;
			uint32_t l;
			if (!bit_get(g1->used, i)) continue;
			c = (int)(g1->keys[i] & YO_MAX_COUNT);
			if (c < min || c > max) continue;
			l = yo_set_put(g0, g1->keys[i] & ~(uint64_t)YO_MAX_COUNT, &absent);
			if ((g0->keys[l] & YO_MAX_COUNT) < YO_MAX_COUNT) ++g0->keys[l];
		}
		h0->tot += g0->count;
	}
	yo_ch_destroy(h1);
}

/* htab.c:287-347: keep the keys of h0 that are absent from (keep_present=0) / present in h1 */
static void filter_members(yo_ch_t *h0, const yo_ch_t *h1, int keep_present)
{
	int w;
	h0->tot = 0;
	for (w = 0; w < 1 << h0->pre; ++w) {
		yo_set_t *g0 = &h0->h[w], f;
		const yo_set_t *g1 = &h1->h[w];
		uint32_t i, n0 = yo_set_capacity(g0);
		memset(&f, 0, sizeof(f));
		yo_set_resize(&f, g0->count);
		for (i = 0; i < n0; ++i) {
			int absent, present;
			if (!bit_get(g0->used, i)) continue;
			present = yo_set_get(g1, g0->keys[i]) != yo_set_capacity(g1);
			if (present == keep_present) yo_set_put(&f, g0->keys[i], &absent);
		}
		set_free(g0);
		*g0 = f;
		h0->tot += g0->count;
	}
}
void yo_ch_subtract(yo_ch_t *h0, const yo_ch_t *h1) { filter_members(h0, h1, 0); }
void yo_ch_isec(yo_ch_t *h0, const yo_ch_t *h1) { filter_members(h0, h1, 1); }

/* htab.c:373-394 */
static int64_t dump_to(const yo_ch_t *h, FILE *fp, uint8_t *mem)
{
	int64_t off = 0;
	uint32_t t[3];
	int w;
#define EMIT(ptr, len) do { if (fp) fwrite((ptr), 1, (len), fp); if (mem) memcpy(mem + off, (ptr), (len)); off += (len); } while (0)
	EMIT("YAK\2", 4);
	t[0] = h->k, t[1] = h->pre, t[2] = YO_COUNTER_BITS;
	EMIT(t, 12);
	for (w = 0; w < 1 << h->pre; ++w) {
		const yo_set_t *g = &h->h[w];
		uint32_t i, n = yo_set_capacity(g);
		t[0] = n, t[1] = g->count;
		EMIT(t, 8);
		for (i = 0; i < n; ++i)
			if (bit_get(g->used, i)) EMIT(&g->keys[i], 8);
	}
#undef EMIT
	return off;
}

int yo_ch_dump(const yo_ch_t *h, const char *fn)
{
	FILE *fp = strcmp(fn, "-") ? fopen(fn, "wb") : stdout;
	if (fp == 0) return -1;
	dump_to(h, fp, 0);
	if (fp != stdout) fclose(fp); else fflush(fp);
	return 0;
}

int64_t yo_ch_dump_mem(const yo_ch_t *h, uint8_t **out)
{
	int64_t len = 16 + 8LL * (1 << h->pre);
	int w;
	for (w = 0; w < 1 << h->pre; ++w) len += 8LL * h->h[w].count;
	*out = (uint8_t*)malloc(len);
	return dump_to(h, 0, *out);
}

/* htab.c:396-481, YAK_LOAD_ALL only */
yo_ch_t *yo_ch_restore(const char *fn)
{
	FILE *fp;
	char magic[4];
	uint32_t t[3], j;
	int w, absent;
	yo_ch_t *h;
	if ((fp = fopen(fn, "rb")) == 0) return 0;
	if (fread(magic, 1, 4, fp) != 4) { fclose(fp); return 0; }
	if (memcmp(magic, "YAK\2", 4) != 0) { fprintf(stderr, "ERROR: wrong file magic.\n"); fclose(fp); return 0; }
	if (fread(t, 4, 3, fp) != 3 || t[2] != YO_COUNTER_BITS) { fclose(fp); return 0; }
	h = yo_ch_init(t[0], t[1], 0, 0);
	for (w = 0; w < 1 << h->pre; ++w) {
		if (fread(t, 4, 2, fp) != 2) break;
		yo_set_resize(&h->h[w], t[0]);
		for (j = 0; j < t[1]; ++j) {
			uint64_t key;
			if (fread(&key, 8, 1, fp) != 1) break;
			yo_set_put(&h->h[w], key, &absent);
		}
	}
	fclose(fp);
	return h;
}

/* htab.c:396-476 with the variadic (min_cnt, mid_cnt) made explicit.  mode: 1 ALL, 2/3 TRIOBIN1/2
 * (count -> 2-bit class, second file in bits 2-3), 4/5/6 SEXCHR1/2/3 (membership bits 0/1/2).
 * n_io (may be NULL) receives {#put calls, #new keys} as printed at htab.c:474. */
yo_ch_t *yo_ch_restore_core(yo_ch_t *ch0, const char *fn, int mode, int min_cnt, int mid_cnt, int64_t n_io[2])
{
	FILE *fp;
	char magic[4];
	uint32_t t[3], j;
	int w, absent;
	const uint64_t mask = YO_MAX_COUNT;
	int64_t n_ins = 0, n_new = 0;
	yo_ch_t *h;
	if (mode < 1 || mode > 6) return 0;                                        /* htab.c:419 */
	if (ch0 == 0 && (mode == 3 || mode == 5 || mode == 6)) return 0;          /* htab.c:412-418 */
	if ((fp = fopen(fn, "rb")) == 0) return 0;
	if (fread(magic, 1, 4, fp) != 4) { fclose(fp); return 0; }
	if (memcmp(magic, "YAK\2", 4) != 0) { fprintf(stderr, "ERROR: wrong file magic.\n"); fclose(fp); return 0; }
	if (fread(t, 4, 3, fp) != 3 || t[2] != YO_COUNTER_BITS) { fclose(fp); return 0; }
	h = ch0 ? ch0 : yo_ch_init(t[0], t[1], 0, 0);
	for (w = 0; w < 1 << h->pre; ++w) {
		if (fread(t, 4, 2, fp) != 2) break;
		yo_set_resize(&h->h[w], t[0]);                                         /* htab.c:438 */
		for (j = 0; j < t[1]; ++j) {
			uint64_t key;
			uint32_t it;
			if (fread(&key, 8, 1, fp) != 1) break;
			if (mode == 1) {
				++n_ins;
				yo_set_put(&h->h[w], key, &absent);
				if (absent) ++n_new;
			} else {
				int x;
				if (mode == 2 || mode == 3) {                                  /* htab.c:448-452 */
					const int cnt = key & mask, shift = mode == 2 ? 0 : 2;
					x = cnt >= mid_cnt ? 2 << shift : cnt >= min_cnt ? 1 << shift : -1;
				} else x = 1 << (mode - 4);                                    /* htab.c:462 */
				if (x < 0) continue;
				key = (key & ~mask) | x;
				++n_ins;
				it = yo_set_put(&h->h[w], key, &absent);
				if (absent) ++n_new;
				else h->h[w].keys[it] |= x;                                    /* htab.c:459, 468 */
			}
		}
	}
	fclose(fp);
	if (n_io) n_io[0] = n_ins, n_io[1] = n_new;
	return h;
}

/* the loop every scanner shares (triobin.c:62-86, trioeval.c:61-89, chkerr.c:35-56, sexchr.c:42-66):
 * out[i] = yak_ch_get of the k-mer ending at base i (-1 absent), or -2 where no k-mer ends */
void yo_scan_seq(const yo_ch_t *ch, int64_t len, const char *seq, int16_t *out)
{
	const int k = ch->k;
	int64_t i;
	int l = 0;
	uint64_t x[4] = {0, 0, 0, 0}, mask = k < 32 ? (1ULL << 2 * k) - 1 : (1ULL << k) - 1;
	const int shift = k < 32 ? 2 * (k - 1) : k - 1;
	for (i = 0; i < len; ++i) {
		int c = yo_nt4[(uint8_t)seq[i]];
		out[i] = -2;
		if (c >= 4) { l = 0, x[0] = x[1] = x[2] = x[3] = 0; continue; }
		if (k < 32) {
			x[0] = (x[0] << 2 | c) & mask;
			x[1] = x[1] >> 2 | (uint64_t)(3 - c) << shift;
		} else {
			x[0] = (x[0] << 1 | (c & 1)) & mask;
			x[1] = (x[1] << 1 | (c >> 1)) & mask;
			x[2] = x[2] >> 1 | (uint64_t)(1 - (c & 1)) << shift;
			x[3] = x[3] >> 1 | (uint64_t)(1 - (c >> 1)) << shift;
		}
		if (++l >= k) out[i] = (int16_t)yo_ch_get(ch, k < 32 ? yo_hash64(x[0] < x[1] ? x[0] : x[1], mask) : yo_hash_long(x));
	}
}

/* ------------------------------------------------------------------ event stream */

/* count.c:28-43 (k < 32) and count.c:45-60 (32 <= k < 64) */
int64_t yo_extract(int k, int64_t len, const char *seq, uint64_t *out)
{
	int64_t i, n = 0;
	int l = 0;
	if (k < 32) {
		uint64_t fwd = 0, rev = 0, mask = (1ULL << 2 * k) - 1;
		int shift = 2 * (k - 1);
		for (i = 0; i < len; ++i) {
			int c = yo_nt4[(uint8_t)seq[i]];
			if (c >= 4) { l = 0, fwd = rev = 0; continue; }
			fwd = (fwd << 2 | c) & mask;
			rev = rev >> 2 | (uint64_t)(3 - c) << shift;
			if (++l >= k) out[n++] = yo_hash64(fwd < rev ? fwd : rev, mask);
		}
	} else {
		uint64_t x[4] = {0, 0, 0, 0}, mask = (1ULL << k) - 1;
		int shift = k - 1;
		for (i = 0; i < len; ++i) {
			int c = yo_nt4[(uint8_t)seq[i]];
			if (c >= 4) { l = 0, x[0] = x[1] = x[2] = x[3] = 0; continue; }
			x[0] = (x[0] << 1 | (c & 1)) & mask;
			x[1] = (x[1] << 1 | (c >> 1)) & mask;
			x[2] = x[2] >> 1 | (uint64_t)(1 - (c & 1)) << shift;
			x[3] = x[3] >> 1 | (uint64_t)(1 - (c >> 1)) << shift;
			if (++l >= k) out[n++] = yo_hash_long(x);
		}
	}
	return n;
}

/* count.c:85-145 done one sequence at a time: partition this read's events by the low `pre`
 * bits (count.c:17-26) and feed each sub-table its events in file order (8.A.1). */
typedef struct { uint64_t *ev; int64_t m; uint64_t **bucket; int *bn, *bm; } feeder_t;

static void feed_seq(yo_ch_t *h, int create_new, feeder_t *f, int64_t len, const char *seq, int64_t *n_events)
{
	int64_t n, i;
	if (len < h->k) return;                                   /* count.c:95 */
	if (len > f->m) { f->m = len + (len >> 1) + 64; f->ev = (uint64_t*)realloc(f->ev, f->m * 8); }
	n = yo_extract(h->k, len, seq, f->ev);
	if (n_events) *n_events += n;
	/* sequential restatement: inserting event by event in file order is identical to the
	 * reference's per-chunk per-sub-table lists because sub-tables are independent */
	for (i = 0; i < n; ++i) h->tot += yo_ch_insert_list(h, create_new, 1, &f->ev[i]);
}

yo_ch_t *yo_count_seqs(int64_t n_seq, const int64_t *lens, const char *cat, int k, int pre,
                       int bf_shift, int bf_n_hash, yo_ch_t *h0, int64_t *n_events)
{
	feeder_t f;
	yo_ch_t *h = h0 ? h0 : yo_ch_init(k, pre, bf_n_hash, bf_shift);
	int64_t i, off = 0;
	if (h == 0) return 0;
	memset(&f, 0, sizeof(f));
	if (n_events) *n_events = 0;
	for (i = 0; i < n_seq; ++i) { feed_seq(h, h0 == 0, &f, lens[i], cat + off, n_events); off += lens[i]; }
	free(f.ev);
	return h;
}

/* count.c:88-110: step 0 of the pipeline collects records until they hold chunk_size bases (records shorter than k do not
 * count) or kseq_read fails; a failure that is a truncated FASTQ record (kseq's -2) only ends that chunk, the next call
 * resumes at the next header character (kseq.h:192-199).  A call that collects nothing returns NULL, which retires the
 * pipeline worker that made it (kthread.c:119); count.c:162 starts 3 workers, so reading goes on until the third empty
 * call - the end of the file gives three in a row, a truncated record that is the first thing a call meets costs one. */
yo_ch_t *yo_count_file_chunked(const char *fn, int k, int pre, int bf_shift, int bf_n_hash, yo_ch_t *h0, int64_t *n_events,
                               int64_t chunk_size)
{
	feeder_t f;
	yo_reader_t *r = yo_reader_open(fn);
	yo_ch_t *h;
	const char *seq;
	int64_t len, sum_len = 0;
	int workers = 3;                                           /* count.c:162 */
	if (r == 0) return 0;                                      /* count.c:152 */
	h = h0 ? h0 : yo_ch_init(k, pre, bf_n_hash, bf_shift);
	memset(&f, 0, sizeof(f));
	if (n_events) *n_events = 0;
	for (;;) {
		len = yo_reader_next(r, &seq, 0);
		if (len < 0) {                                         /* the call ends here */
			if (len == -1) break;                              /* end of file: every later call is empty as well */
			if (sum_len == 0 && --workers == 0) break;         /* count.c:109 + kthread.c:119 */
			sum_len = 0;                                       /* the next call starts a new chunk */
			continue;
		}
		feed_seq(h, h0 == 0, &f, len, seq, n_events);
		if (len < k) continue;                                 /* count.c:95 */
		sum_len += len;
		if (sum_len >= chunk_size) sum_len = 0;                /* count.c:106 */
	}
	free(f.ev);
	yo_reader_close(r);
	return h;
}

yo_ch_t *yo_count_file(const char *fn, int k, int pre, int bf_shift, int bf_n_hash, yo_ch_t *h0, int64_t *n_events)
{
	return yo_count_file_chunked(fn, k, pre, bf_shift, bf_n_hash, h0, n_events, 10000000); /* misc.c:31 */
}

/* ------------------------------------------------------------------ qv scan */

/* qv.c:34-86 without the printing; qv.c:129-133 reduction */
void yo_qv_seqs(const yo_ch_t *ch, int64_t n_seq, const int64_t *lens, const char *cat,
                int min_len, double min_frac, int64_t cnt[YO_N_COUNTS], int32_t *out_tot, int32_t *out_non0)
{
	int64_t s, off = 0, m = 0;
	uint64_t *ev = 0;
	memset(cnt, 0, YO_N_COUNTS * sizeof(int64_t));
	for (s = 0; s < n_seq; off += lens[s], ++s) {
		int64_t n, i;
		int tot = 0, non0 = 0;
		if (out_tot) out_tot[s] = 0;
		if (out_non0) out_non0[s] = 0;
		if (lens[s] < min_len) continue;                       /* qv.c:44 */
		if (lens[s] > m) { m = lens[s] + 64; ev = (uint64_t*)realloc(ev, m * 8); }
		n = yo_extract(ch->k, lens[s], cat + off, ev);
		for (i = 0; i < n; ++i) {
			int t = yo_ch_get(ch, ev[i]);
			if (t < 0) t = 0;
			if (t > 0) ++non0;
			ev[i] = t;
			++tot;
		}
		if (out_tot) out_tot[s] = tot;
		if (out_non0) out_non0[s] = non0;
		if (non0 < tot * min_frac) continue;                   /* qv.c:83 */
		for (i = 0; i < n; ++i) ++cnt[ev[i]];
	}
	free(ev);
}

/* ------------------------------------------------------------------ reader */

struct yo_reader_s {
	gzFile fp;
	unsigned char buf[16384];
	int beg, end, eof;
	int last;              /* header char already consumed, or 0 */
	char *name, *seq, *qual;
	int64_t name_l, name_m, seq_l, seq_m, qual_l, qual_m;
};

static int rd_getc(yo_reader_t *r)
{
	if (r->beg >= r->end) {
		if (r->eof) return -1;
		r->beg = 0;
		r->end = gzread(r->fp, r->buf, sizeof(r->buf));
		if (r->end < (int)sizeof(r->buf)) r->eof = 1;
		if (r->end <= 0) return -1;
	}
	return r->buf[r->beg++];
}

static void push(char **s, int64_t *l, int64_t *m, int c)
{
	if (*l + 2 > *m) { *m = *m ? *m * 2 : 256; *s = (char*)realloc(*s, *m); }
	(*s)[(*l)++] = (char)c;
	(*s)[*l] = 0;
}

/* read to end of line into (s,l,m) appending; strips one trailing '\r' like kseq.h:146; returns
 * -1 if nothing could be read because the stream is exhausted */
static int rd_line(yo_reader_t *r, char **s, int64_t *l, int64_t *m)
{
	int c, got = 0;
	if (r->beg >= r->end && r->eof) return -1;
	while ((c = rd_getc(r)) != -1) { got = 1; if (c == '\n') break; push(s, l, m, c); }
	if (!got && c == -1) return -1;
	if (*l > 1 && (*s)[*l - 1] == '\r') (*s)[--*l] = 0;
	return 0;
}

yo_reader_t *yo_reader_open(const char *fn)
{
	gzFile fp = (fn == 0 || strcmp(fn, "-") == 0) ? gzdopen(0, "r") : gzopen(fn, "r");
	yo_reader_t *r;
	if (fp == 0) return 0;
	r = (yo_reader_t*)calloc(1, sizeof(*r));
	r->fp = fp;
	return r;
}

void yo_reader_close(yo_reader_t *r)
{
	if (!r) return;
	gzclose(r->fp);
	free(r->name); free(r->seq); free(r->qual); free(r);
}

/* kseq.h:192-232 */
int64_t yo_reader_next(yo_reader_t *r, const char **seq, const char **name)
{
	int c;
	if (r->last == 0) {
		while ((c = rd_getc(r)) != -1 && c != '>' && c != '@') {}
		if (c == -1) return -1;
		r->last = c;
	}
	r->name_l = r->seq_l = r->qual_l = 0;
	if (r->name) r->name[0] = 0;
	/* name: up to the first whitespace; rest of the line is the comment */
	{
		int got = 0;
		while ((c = rd_getc(r)) != -1) { got = 1; if (isspace(c)) break; push(&r->name, &r->name_l, &r->name_m, c); }
		if (!got) return -1;
		if (c != -1 && c != '\n') while ((c = rd_getc(r)) != -1 && c != '\n') {}
	}
	if (r->seq == 0) { r->seq_m = 256; r->seq = (char*)malloc(256); }
	r->seq[0] = 0;
	while ((c = rd_getc(r)) != -1 && c != '>' && c != '+' && c != '@') {
		if (c == '\n') continue;
		push(&r->seq, &r->seq_l, &r->seq_m, c);
		rd_line(r, &r->seq, &r->seq_l, &r->seq_m);
	}
	if (c == '>' || c == '@') r->last = c;
	if (seq) *seq = r->seq;
	if (name) *name = r->name ? r->name : "";
	if (c != '+') return r->seq_l;
	while ((c = rd_getc(r)) != -1 && c != '\n') {}
	if (c == -1) return -2;
	while (rd_line(r, &r->qual, &r->qual_l, &r->qual_m) >= 0 && r->qual_l < r->seq_l) {}
	r->last = 0;
	if (r->qual_l != r->seq_l) return -2;
	return r->seq_l;
}
