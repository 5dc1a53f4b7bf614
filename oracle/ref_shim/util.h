/* oracle/ref_shim/util.h -- minimal stand-in for the reference's util.h so that its hot-path sources
 * (libcsdr.c, fastddc.c, block.c, fft.c, hfdl.c, input-helpers.c) compile where they lie without libacars / glib.
 * Only what those files use: allocation / assertion / debug macros, container_of, REVERSE_BYTE (the bit hack of
 * util.h:109 is the interface: hfdl.c:1052 relies on its value), struct octet_string, thread helpers.
 * The functions are defined in ref_shim/ref_host.c.  Test infrastructure only. */
#pragma once
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <stdbool.h>
#include <stddef.h>
#include <pthread.h>
#define D_SDR (1 << 0)
#define D_DSP (1 << 1)
#define D_DSP_DETAIL (1 << 2)
#define D_FRAME (1 << 3)
#define D_FRAME_DETAIL (1 << 4)
#define D_MISC (1 << 31)
#define nop() do {} while (0)
#define LIKELY(x)   (__builtin_expect(!!(x),1))
#define UNLIKELY(x) (__builtin_expect(!!(x),0))
#define ASSERT(expr) do { if(!(expr)) { fprintf(stderr, "ASSERT %s failed %s:%d\n", #expr, __FILE__, __LINE__); abort(); } } while(0)
#define XCALLOC(nmemb, size) calloc((nmemb), (size))
#define XFREE(ptr) do { free(ptr); ptr = NULL; } while(0)
#define NEW(type, x) type *(x) = XCALLOC(1, sizeof(type))
#define UNUSED(x) (void)(x)
#define container_of(ptr, type, member) ((type *)((char *)(ptr) - offsetof(type, member)))
#define max(a, b) ((a) > (b) ? (a) : (b))
#define debug_print(cls, ...) nop()
#define debug_print_buf_hex(cls, buf, len, ...) nop()
#define REVERSE_BYTE(x) (uint8_t)((((x) * 0x80200802ULL) & 0x0884422110ULL) * 0x0101010101ULL >> 32)

struct dumphfdl_config { int32_t nf_stats_interval; bool datadumps; };
extern struct dumphfdl_config Config;

int32_t start_thread(pthread_t *pth, void *(*start_routine)(void *), void *thread_ctx);
int32_t pthread_barrier_create(pthread_barrier_t *barrier, unsigned count);
int32_t pthread_cond_initialize(pthread_cond_t *cond);
int32_t pthread_mutex_initialize(pthread_mutex_t *mutex);

struct octet_string { uint8_t *buf; size_t len; };
struct octet_string *octet_string_new(void *buf, size_t len);
void octet_string_destroy(struct octet_string *ostring);
