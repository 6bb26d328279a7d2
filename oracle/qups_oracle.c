/* qups_oracle.c — CPU oracle (plain C + OpenMP); see qups_oracle.h for the
 * TEST-INFRASTRUCTURE-ONLY notice and the parity-pinning status.
 *
 * Build (oracle/Makefile):
 *   gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp -fPIC -shared
 * -ffp-contract=off is REQUIRED: the fp32 build must round every operation
 * individually (canonical fp32 sequence).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>
#include "qups_oracle.h"

int oracle_num_threads(void) { return omp_get_max_threads(); }
void oracle_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }

/* ---- fp32 instantiation ---- */
#define REAL float
#define SFX _f
#define ROUND roundf
#define FLOOR floorf
#define SQRT sqrtf
#define COS cosf
#define SIN sinf
#include "oracle_body.inc"
#undef REAL
#undef SFX
#undef ROUND
#undef FLOOR
#undef SQRT
#undef COS
#undef SIN

/* ---- fp64 instantiation (arbiter) ---- */
#define REAL double
#define SFX _d
#define ROUND round
#define FLOOR floor
#define SQRT sqrt
#define COS cos
#define SIN sin
#include "oracle_body.inc"
#undef REAL
#undef SFX
#undef ROUND
#undef FLOOR
#undef SQRT
#undef COS
#undef SIN
