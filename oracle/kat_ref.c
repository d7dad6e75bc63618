/* kat_ref.c - thin exports of the reference's static-inline primitives, compiled AGAINST the
 * reference headers where they lie (-I/root/reference) into oracle/_ref/libyakref.so.
 * Test infrastructure only; contains no reference code itself. */
#include <stdint.h>
#include "yak-priv.h"
#include "khashl.h"

uint64_t ref_hash64(uint64_t key, uint64_t mask) { return yak_hash64(key, mask); }
uint64_t ref_hash64_64(uint64_t key) { return yak_hash64_64(key); }
uint64_t ref_hash_long(uint64_t *x) { return yak_hash_long(x); }
uint64_t ref_hash64_inv(uint64_t key, uint64_t mask) { return yak_hash64_inv(key, mask); }
uint32_t ref_h2b(uint32_t hash, uint32_t bits) { return __kh_h2b(hash, bits); }
