/* Trimmed stand-in for base/sfem_defs.hpp: the real one re-exports a long list of
 * smesh:: helpers that are unavailable; the element kernels need none of them. */
#ifndef SFEM_DEFS_H
#define SFEM_DEFS_H
#include "sfem_base.hpp"
#include "smesh_elem_type.hpp"
#include <assert.h>
namespace sfem { static void *SFEM_DEFAULT_STREAM = 0; }
#endif
