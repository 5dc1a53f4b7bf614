/* oracle/ref_shim/front/glib.h -- DECLARATION-ONLY stand-in for GLib's <glib.h> (GAsyncQueue only) (not installed here), just enough for the
 * reference's pdu.c / mpdu.c / spdu.c / lpdu.c / util.c to compile where they lie; the few functions the parse path
 * really calls are defined in ref_front_host.c, every other one aborts (ref_front_stubs.c).  Test infrastructure. */
#pragma once
typedef struct _GAsyncQueue GAsyncQueue;
typedef void *gpointer;
GAsyncQueue *g_async_queue_new(void);
void g_async_queue_push(GAsyncQueue *q, gpointer data);
gpointer g_async_queue_pop(GAsyncQueue *q);
int g_async_queue_length(GAsyncQueue *q);
