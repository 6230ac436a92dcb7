/* TEST INFRASTRUCTURE — rand()/srand() for the reference oracle binary.
 *
 * The reference draws every random number from libc rand() (src/rand_gen.c:9-12, seeded by
 * srand(seed) at src/main.c:403,1590-1593).  On this image that is glibc's TYPE_3 additive
 * feedback generator; it is restated here (and again, independently, in odexp_b200/glibc_rand.py
 * and oracle/odx_oracle.c) so that
 *   (1) the oracle binary does not depend on which libc it runs against, and
 *   (2) every draw can be logged (ODX_RAND_LOG=<file>: raw int32 per draw) for replay tests.
 * tests/test_oracle_ref.py checks this restatement against the libc of the machine it runs on.
 */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>

static int32_t ring_[31];
static int fptr_ = 3, rptr_ = 0;
static int seeded_ = 0;
static FILE *log_ = NULL;
static int log_checked_ = 0;

void srand(unsigned int seed) {
  int i; int32_t word;
  if (seed == 0) seed = 1;
  ring_[0] = (int32_t)seed; word = (int32_t)seed;
  for (i = 1; i < 31; i++) {           /* Park-Miller minimal standard via Schrage, as glibc srandom_r */
    long hi = word / 127773, lo = word % 127773;
    word = (int32_t)(16807 * lo - 2836 * hi);
    if (word < 0) word += 2147483647;
    ring_[i] = word;
  }
  fptr_ = 3; rptr_ = 0;
  for (i = 0; i < 310; i++) {          /* discard 10 * 31 outputs */
    ring_[fptr_] = (int32_t)((uint32_t)ring_[fptr_] + (uint32_t)ring_[rptr_]);
    if (++fptr_ >= 31) fptr_ = 0;
    if (++rptr_ >= 31) rptr_ = 0;
  }
  seeded_ = 1;
}

int rand(void) {
  uint32_t v; int out;
  if (!seeded_) srand(1);
  v = (uint32_t)ring_[fptr_] + (uint32_t)ring_[rptr_];
  ring_[fptr_] = (int32_t)v;
  out = (int)(v >> 1);
  if (++fptr_ >= 31) fptr_ = 0;
  if (++rptr_ >= 31) rptr_ = 0;
  if (!log_checked_) { const char *p = getenv("ODX_RAND_LOG"); log_checked_ = 1; if (p && *p) log_ = fopen(p, "wb"); }
  if (log_) { int32_t o = out; fwrite(&o, sizeof o, 1, log_); fflush(log_); }
  return out;
}
