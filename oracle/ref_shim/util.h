/* oracle/ref_shim/util.h -- minimal stand-in for the reference's util.h so that its
 * libcsdr.c / fastddc.c compile standalone (the real util.h pulls in libacars, glib types).
 * Only the allocation / assertion macros those two files use are provided. Test infrastructure only. */
#pragma once
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <stddef.h>
#define nop() do {} while (0)
#define ASSERT(expr) do { if(!(expr)) { fprintf(stderr, "ASSERT %s failed %s:%d\n", #expr, __FILE__, __LINE__); abort(); } } while(0)
#define XCALLOC(nmemb, size) calloc((nmemb), (size))
#define XFREE(ptr) do { free(ptr); ptr = NULL; } while(0)
#define NEW(type, x) type *(x) = XCALLOC(1, sizeof(type))
#define UNUSED(x) (void)(x)
#define D_DSP 0
#define debug_print(cls, ...) nop()
