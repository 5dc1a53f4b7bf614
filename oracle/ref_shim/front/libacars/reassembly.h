/* oracle/ref_shim/front/libacars/reassembly.h -- DECLARATION-ONLY stand-in for libacars 2.x's <libacars/reassembly.h> (not installed here), just enough for the
 * reference's pdu.c / mpdu.c / spdu.c / lpdu.c / util.c to compile where they lie; the few functions the parse path
 * really calls are defined in ref_front_host.c, every other one aborts (ref_front_stubs.c).  Test infrastructure. */
#pragma once
typedef struct la_reasm_ctx_s la_reasm_ctx;
la_reasm_ctx *la_reasm_ctx_new(void);
void la_reasm_ctx_destroy(void *ctx);
