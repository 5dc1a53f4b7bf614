/* oracle/ref_shim/front/libacars/list.h -- DECLARATION-ONLY stand-in for libacars 2.x's <libacars/list.h> (not installed here), just enough for the
 * reference's pdu.c / mpdu.c / spdu.c / lpdu.c / util.c to compile where they lie; the few functions the parse path
 * really calls are defined in ref_front_host.c, every other one aborts (ref_front_stubs.c).  Test infrastructure. */
#pragma once
#include <stddef.h>
typedef struct la_list la_list;
struct la_list { void *data; la_list *next; };
la_list *la_list_next(la_list const *l);
la_list *la_list_append(la_list *l, void *data);
size_t la_list_length(la_list const *l);
void la_list_foreach(la_list *l, void (*cb)(void *, void *), void *ctx);
void la_list_free(la_list *l);
void la_list_free_full(la_list *l, void (*node_free)(void *));
void la_list_free_full_with_ctx(la_list *l, void (*node_free)(void *, void *), void *ctx);
