/* oracle/ref_shim/front/libacars/json.h -- DECLARATION-ONLY stand-in for libacars 2.x's <libacars/json.h> (not installed here), just enough for the
 * reference's pdu.c / mpdu.c / spdu.c / lpdu.c / util.c to compile where they lie; the few functions the parse path
 * really calls are defined in ref_front_host.c, every other one aborts (ref_front_stubs.c).  Test infrastructure. */
#pragma once
#include <stdint.h>
#include <stdbool.h>
#include "vstring.h"
void la_json_start(la_vstring *vstr);
void la_json_end(la_vstring *vstr);
void la_json_object_start(la_vstring *vstr, char const *key);
void la_json_object_end(la_vstring *vstr);
void la_json_array_start(la_vstring *vstr, char const *key);
void la_json_array_end(la_vstring *vstr);
void la_json_append_bool(la_vstring *vstr, char const *key, bool val);
void la_json_append_double(la_vstring *vstr, char const *key, double val);
void la_json_append_int64(la_vstring *vstr, char const *key, int64_t val);
void la_json_append_long(la_vstring *vstr, char const *key, long val);
void la_json_append_char(la_vstring *vstr, char const *key, char val);
void la_json_append_string(la_vstring *vstr, char const *key, char const *val);
void la_json_append_octet_string(la_vstring *vstr, char const *key, uint8_t const *buf, size_t len);
