/* oracle/ref_shim/front/libacars/libacars.h -- DECLARATION-ONLY stand-in for libacars 2.x's <libacars/libacars.h> (not installed here), just enough for the
 * reference's pdu.c / mpdu.c / spdu.c / lpdu.c / util.c to compile where they lie; the few functions the parse path
 * really calls are defined in ref_front_host.c, every other one aborts (ref_front_stubs.c).  Test infrastructure. */
#pragma once
#include <stddef.h>
#include <stdbool.h>
#include <stdint.h>
#include "vstring.h"
typedef enum { LA_MSG_DIR_UNKNOWN = 0, LA_MSG_DIR_GND2AIR, LA_MSG_DIR_AIR2GND } la_msg_dir;
typedef void (la_format_text_func)(la_vstring *vstr, void const *data, int indent);
typedef void (la_format_json_func)(la_vstring *vstr, void const *data);
typedef void (la_destroy_type_f)(void *data);
typedef struct { la_format_text_func *format_text; la_destroy_type_f *destroy; la_format_json_func *format_json; char const *json_key; } la_type_descriptor;
typedef struct la_proto_node la_proto_node;
struct la_proto_node { la_type_descriptor const *td; void *data; la_proto_node *next; };
la_proto_node *la_proto_node_new(void);
void la_proto_tree_destroy(la_proto_node *root);
la_proto_node *la_proto_tree_find_protocol(la_proto_node *root, la_type_descriptor const *td);
la_vstring *la_proto_tree_format_text(la_vstring *vstr, la_proto_node const *root);
la_vstring *la_proto_tree_format_json(la_vstring *vstr, la_proto_node const *root);
