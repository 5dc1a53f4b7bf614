/* oracle/ref_shim/front/libacars/hash.h -- DECLARATION-ONLY stand-in for libacars 2.x's <libacars/hash.h> (not installed here), just enough for the
 * reference's pdu.c / mpdu.c / spdu.c / lpdu.c / util.c to compile where they lie; the few functions the parse path
 * really calls are defined in ref_front_host.c, every other one aborts (ref_front_stubs.c).  Test infrastructure. */
#pragma once
#include <stdint.h>
#include <stdbool.h>
typedef struct la_hash_s la_hash;
typedef uint32_t (la_hash_func)(void const *key);
typedef bool (la_hash_compare_func)(void const *key1, void const *key2);
typedef void (la_hash_key_destroy_func)(void *key);
typedef void (la_hash_value_destroy_func)(void *value);
typedef bool (la_hash_if_func)(void const *key, void const *value, void *ctx);
