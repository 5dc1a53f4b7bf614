/*
 * oracle/ref_shim/front/ref_front_stubs.c -- TEST INFRASTRUCTURE (see ref_front_host.c).
 * Link-time stand-ins for what lies DOWNSTREAM of the LPDU frame check and outside the hot path's scope (SURVEY 2,
 * "OUT OF SCOPE"): HFNPDU / ACARS parsing, aircraft cache, aircraft database, system table, outputs, and libacars'
 * text / JSON formatting.  Deliberately compiled without any reference header (K&R-style definitions), so that no
 * signature has to be restated.  "return 0" stubs are on the parse path (lpdu.c:160-196 calls them and goes on);
 * formatting is never requested by the recording formatter, so those abort if they are ever reached.
 */
#include <stdio.h>
#include <stdlib.h>
#define RETURNS_0(name) void *name() { return 0; }
#define MUST_NOT_RUN(name) void *name() { fprintf(stderr, "ref_front_stubs: %s called\n", #name); abort(); return 0; }

RETURNS_0(hfnpdu_parse)                      /* lpdu.c:196: the LPDU node simply has no child */
RETURNS_0(ac_cache_entry_create)             /* lpdu.c:171 */
RETURNS_0(ac_cache_entry_delete)             /* lpdu.c:161 */
RETURNS_0(ac_cache_entry_lookup)
RETURNS_0(ac_data_entry_lookup)
RETURNS_0(systable_get_station_name)
RETURNS_0(systable_get_station_frequency)
RETURNS_0(shutdown_outputs)                  /* pdu.c:110, on hfdl_pdu_decoder_stop */
RETURNS_0(output_queue_push)                 /* pdu.c:149: never reached, the recording formatter returns NULL */

MUST_NOT_RUN(hfnpdu_position_info_extract)
MUST_NOT_RUN(position_info_destroy)
MUST_NOT_RUN(la_dict_search)
MUST_NOT_RUN(la_vstring_append_sprintf)
MUST_NOT_RUN(la_isprintf_multiline_text)
MUST_NOT_RUN(la_json_object_start)
MUST_NOT_RUN(la_json_object_end)
MUST_NOT_RUN(la_json_array_start)
MUST_NOT_RUN(la_json_array_end)
MUST_NOT_RUN(la_json_append_bool)
MUST_NOT_RUN(la_json_append_double)
MUST_NOT_RUN(la_json_append_int64)
MUST_NOT_RUN(la_json_append_string)
MUST_NOT_RUN(la_json_append_octet_string)
