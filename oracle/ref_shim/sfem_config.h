/* Build shim for compiling the reference's hot-path operator sources out of tree.
 * TEST INFRASTRUCTURE ONLY (see oracle/README.md). Mirrors the typedef set that the
 * reference generates from cmake/SFEMNumbers.cmake:5-14 through
 * configuration/sfem_config.h.in:46-107 (real/scalar/accumulator = fp64 by default,
 * geom/jacobian = fp32, idx/count/element_idx = int32, local_idx = int16).
 * -DORACLE_REAL_FLOAT32 gives the SFEM_REAL_TYPE=float32 configuration. */
#ifndef SFEM_CONFIG_HPP
#define SFEM_CONFIG_HPP

#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define SFEM_ENABLE_OPENMP

typedef float geom_t;
#ifdef ORACLE_REAL_FLOAT32
typedef float real_t;
typedef float scalar_t;
typedef float accumulator_t;
#else
typedef double real_t;
typedef double scalar_t;
typedef double accumulator_t;
#endif
typedef float jacobian_t;
typedef float cu_jacobian_t;
typedef int32_t idx_t;
typedef int32_t count_t;
typedef int32_t element_idx_t;
typedef int16_t local_idx_t;
typedef char mask_t;

#define SFEM_VEC_SIZE 1
typedef scalar_t vec_t;

#define SFEM_IDX_INVALID (-1)
#define SFEM_COUNT_INVALID (-1)
#define SFEM_ELEMENT_IDX_INVALID (-1)
#define SFEM_PTRDIFF_INVALID (-1)
#define SFEM_LOG_LEVEL 0
#define SFEM_PRINT_THROUGHPUT 0

#endif
