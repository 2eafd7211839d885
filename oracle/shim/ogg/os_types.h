/* TEST INFRASTRUCTURE (oracle build only) -- minimal stand-in for libogg's
 * <ogg/os_types.h>.  libogg is an external dependency of the reference
 * (configure.ac:247, "ogg >= 1.0", not vendored, not installed here); the
 * reference only needs its integer typedefs and allocator macros from this
 * header.  Written from the published libogg API, no libogg source used. */
#ifndef VB200_SHIM_OS_TYPES_H
#define VB200_SHIM_OS_TYPES_H
#include <stdint.h>
#include <stdlib.h>
#define _ogg_malloc  malloc
#define _ogg_calloc  calloc
#define _ogg_realloc realloc
#define _ogg_free    free
typedef int16_t  ogg_int16_t;
typedef uint16_t ogg_uint16_t;
typedef int32_t  ogg_int32_t;
typedef uint32_t ogg_uint32_t;
typedef int64_t  ogg_int64_t;
typedef uint64_t ogg_uint64_t;
#endif
