/* oracle/ref_shim/front/libacars/dict.h -- DECLARATION-ONLY stand-in for libacars 2.x's <libacars/dict.h> (not installed here), just enough for the
 * reference's pdu.c / mpdu.c / spdu.c / lpdu.c / util.c to compile where they lie; the few functions the parse path
 * really calls are defined in ref_front_host.c, every other one aborts (ref_front_stubs.c).  Test infrastructure. */
#pragma once
typedef struct { int id; void *val; } la_dict;
void *la_dict_search(la_dict const *list, int id);
