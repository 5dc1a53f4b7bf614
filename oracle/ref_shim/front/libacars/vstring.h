/* oracle/ref_shim/front/libacars/vstring.h -- DECLARATION-ONLY stand-in for libacars 2.x's <libacars/vstring.h> (not installed here), just enough for the
 * reference's pdu.c / mpdu.c / spdu.c / lpdu.c / util.c to compile where they lie; the few functions the parse path
 * really calls are defined in ref_front_host.c, every other one aborts (ref_front_stubs.c).  Test infrastructure. */
#pragma once
#include <stddef.h>
typedef struct { char *str; size_t len; size_t allocated_size; } la_vstring;
la_vstring *la_vstring_new(void);
void la_vstring_destroy(la_vstring *vstr, _Bool destroy_buffer);
void la_vstring_append_sprintf(la_vstring *vstr, char const *fmt, ...);
void la_vstring_append_buffer(la_vstring *vstr, void const *buffer, size_t size);
void la_isprintf_multiline_text(la_vstring *vstr, int indent, char const *text);
#define LA_ISPRINTF(vstr, i, f, ...) la_vstring_append_sprintf(vstr, "%*s" f, i, "", ##__VA_ARGS__)
