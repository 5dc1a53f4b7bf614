/*
 * oracle/ref_shim/front/ref_front_host.c -- TEST INFRASTRUCTURE (see oracle/orc.h).
 *
 * Host-program stand-in around the reference's OWN pdu.c, mpdu.c, spdu.c, lpdu.c, util.c and crc.c, compiled where
 * they lie into oracle/_ref/libref_front.so (recipe: oracle/Makefile).  The reference's real pdu_decoder_thread
 * (pdu.c:91-178) is started with hfdl_pdu_decoder_init / hfdl_pdu_decoder_start, PDUs are handed to it with
 * pdu_decoder_queue_push exactly as hfdl.c:1058-1080 does, and what it does with them is recorded:
 *   - every statsd_increment_per_channel() hook it fires (frames.processed / frames.good / frame.errors.* /
 *     frame.dir.* / lpdus.processed / lpdus.good / lpdu.errors.*), per PDU;
 *   - every protocol node it hands to a formatter (which MPDU / SPDU / LPDU nodes survive the FCS filters).
 * That is the ground truth for the front parser: tests compare it with
 * orc_pdu_front_parse and with the device's pdu_front_parse.
 *
 * Defined here: what main.c owns (Config and the other globals), a GAsyncQueue, the handful of libacars list / node
 * functions the parse path calls, the statsd capture, and a formatter instance that records instead of formatting.
 * Everything downstream of the LPDU frame check (hfnpdu_parse, aircraft cache, system table) is out of the hot path's
 * scope and stubbed in ref_front_stubs.c.
 */
#include <string.h>
#include <glib.h>
#include <libacars/libacars.h>
#include <libacars/list.h>
#include <libacars/reassembly.h>
#include "util.h"
#include "globals.h"
#include "pdu.h"
#include "statsd.h"
#include "output-common.h"

/* ---- what main.c defines ---- */
struct dumphfdl_config Config;
int32_t do_exit, exitcode;
systable *Systable;
pthread_mutex_t Systable_lock = PTHREAD_MUTEX_INITIALIZER;
ac_cache *AC_cache;
pthread_mutex_t AC_cache_lock = PTHREAD_MUTEX_INITIALIZER;
ac_data *AC_data;

/* ---- GAsyncQueue: FIFO with a blocking pop; "idle" = the consumer waits in pop on an empty queue ---- */
struct qnode { void *data; struct qnode *next; };
struct _GAsyncQueue { pthread_mutex_t mu; pthread_cond_t cv, idle_cv; struct qnode *head, *tail; int waiting; };
static GAsyncQueue *the_queue;          /* pdu.c keeps its queue private (pdu.c:23); it creates exactly one (pdu.c:46) */
GAsyncQueue *g_async_queue_new(void) {
	GAsyncQueue *q = calloc(1, sizeof(*q));
	pthread_mutex_init(&q->mu, NULL); pthread_cond_init(&q->cv, NULL); pthread_cond_init(&q->idle_cv, NULL);
	the_queue = q;
	return q;
}
void g_async_queue_push(GAsyncQueue *q, gpointer data) {
	struct qnode *n = calloc(1, sizeof(*n));
	n->data = data;
	pthread_mutex_lock(&q->mu);
	if(q->tail) q->tail->next = n; else q->head = n;
	q->tail = n;
	pthread_cond_signal(&q->cv);
	pthread_mutex_unlock(&q->mu);
}
gpointer g_async_queue_pop(GAsyncQueue *q) {
	pthread_mutex_lock(&q->mu);
	while(!q->head) {
		q->waiting = 1;
		pthread_cond_broadcast(&q->idle_cv);
		pthread_cond_wait(&q->cv, &q->mu);
	}
	q->waiting = 0;
	struct qnode *n = q->head;
	q->head = n->next;
	if(!q->head) q->tail = NULL;
	pthread_mutex_unlock(&q->mu);
	void *d = n->data;
	free(n);
	return d;
}
static void wait_idle(void) {
	GAsyncQueue *q = the_queue;
	pthread_mutex_lock(&q->mu);
	while(q->head || !q->waiting) pthread_cond_wait(&q->idle_cv, &q->mu);
	pthread_mutex_unlock(&q->mu);
}

/* ---- libacars: la_list, la_proto_node (published semantics of libacars 2.x list.c / libacars.c) ---- */
la_list *la_list_next(la_list const *l) { return l ? l->next : NULL; }
la_list *la_list_append(la_list *l, void *data) {
	la_list *n = calloc(1, sizeof(*n));
	n->data = data;
	if(!l) return n;
	la_list *p = l;
	while(p->next) p = p->next;
	p->next = n;
	return l;
}
size_t la_list_length(la_list const *l) { size_t n = 0; for(; l; l = l->next) n++; return n; }
void la_list_foreach(la_list *l, void (*cb)(void *, void *), void *ctx) { for(; l; l = l->next) cb(l->data, ctx); }
void la_list_free_full(la_list *l, void (*node_free)(void *)) {
	while(l) { la_list *nx = l->next; if(node_free) node_free(l->data); else free(l->data); free(l); l = nx; }
}
void la_list_free(la_list *l) { la_list_free_full(l, NULL); }
la_proto_node *la_proto_node_new(void) { return calloc(1, sizeof(la_proto_node)); }
void la_proto_tree_destroy(la_proto_node *root) {
	if(!root) return;
	if(root->next) la_proto_tree_destroy(root->next);
	if(root->td && root->td->destroy) root->td->destroy(root->data); else free(root->data);
	free(root);
}
la_proto_node *la_proto_tree_find_protocol(la_proto_node *root, la_type_descriptor const *td) {
	for(; root; root = root->next) if(root->td == td) return root;
	return NULL;
}
struct la_reasm_ctx_s { int unused; };
la_reasm_ctx *la_reasm_ctx_new(void) { return calloc(1, sizeof(struct la_reasm_ctx_s)); }
void la_reasm_ctx_destroy(void *ctx) { free(ctx); }

/* ---- capture ---- */
enum { K_FRAMES_PROCESSED, K_FRAMES_GOOD, K_FRAME_BAD_FCS, K_FRAME_TOO_SHORT, K_DIR_AIR2GND, K_DIR_GND2AIR,
	K_LPDUS_PROCESSED, K_LPDUS_GOOD, K_LPDU_BAD_FCS, K_LPDU_TOO_SHORT, K_OTHER,
	K_NODES_MPDU, K_NODES_SPDU, K_NODES_LPDU, K_NODES_OTHER, K_COUNT };
static const char *const k_names[K_OTHER] = { "frames.processed", "frames.good", "frame.errors.bad_fcs", "frame.errors.too_short",
	"frame.dir.air2gnd", "frame.dir.gnd2air", "lpdus.processed", "lpdus.good", "lpdu.errors.bad_fcs", "lpdu.errors.too_short" };
static int64_t cur[K_COUNT];
static int32_t cur_freq;
static int64_t wrong_freq;

void statsd_counter_per_channel_increment(int32_t freq, char *counter) {          /* statsd.c, capture instead of UDP */
	int k = K_OTHER;
	for(int i = 0; i < K_OTHER; i++) if(strcmp(counter, k_names[i]) == 0) { k = i; break; }
	cur[k]++;
	if(freq != cur_freq) wrong_freq++;
}

extern la_type_descriptor const proto_DEF_hfdl_mpdu, proto_DEF_hfdl_spdu, proto_DEF_hfdl_lpdu;
static struct octet_string *record_decoded(struct metadata *m, la_proto_node *root) {      /* fmt_decoded_fun_t */
	(void)m;
	if(root->td == &proto_DEF_hfdl_mpdu) cur[K_NODES_MPDU]++;
	else if(root->td == &proto_DEF_hfdl_spdu) cur[K_NODES_SPDU]++;
	else if(root->td == &proto_DEF_hfdl_lpdu) cur[K_NODES_LPDU]++;
	else cur[K_NODES_OTHER]++;
	return NULL;                                  /* "this formatter does not handle the message": nothing goes to an output */
}
static bool any_type(fmtr_input_type_t t) { (void)t; return true; }
static fmtr_descriptor_t rec_td = { "record", "records what the decoder thread delivers", record_decoded, NULL, any_type, OFMT_TEXT };
static fmtr_instance_t rec_fmtr = { &rec_td, FMTR_INTYPE_DECODED_FRAME, NULL };
static la_list *fmtr_list;
static int started;

static int ref_front_start(int output_mpdus, int output_corrupted) {
	if(started) return 0;
	memset(&Config, 0, sizeof(Config));
	Config.output_mpdus = output_mpdus != 0;
	Config.output_corrupted_pdus = output_corrupted != 0;
	hfdl_pdu_decoder_init();
	fmtr_list = la_list_append(NULL, &rec_fmtr);
	if(hfdl_pdu_decoder_start(fmtr_list) != 0) return -1;
	started = 1;
	return 0;
}

int32_t ref_front_counter_count(void) { return K_COUNT; }
const char *ref_front_counter_name(int32_t k) {
	static const char *const extra[] = { "other", "nodes.mpdu", "nodes.spdu", "nodes.lpdu", "nodes.other" };
	return k < 0 || k >= K_COUNT ? NULL : k < K_OTHER ? k_names[k] : extra[k - K_OTHER];
}

/* Runs n PDUs (concatenated in bufs, lengths in lens) through the reference's decoder thread, one at a time, and writes
 * K_COUNT counters per PDU.  Metadata as dispatch_pdu builds it (hfdl.c:1061-1069). */
int32_t ref_front_run(const uint8_t *bufs, const int32_t *lens, int32_t n, int32_t freq, int32_t output_mpdus, int32_t output_corrupted, int64_t *out) {
	if(ref_front_start(output_mpdus, output_corrupted)) return -1;
	Config.output_mpdus = output_mpdus != 0;
	Config.output_corrupted_pdus = output_corrupted != 0;
	wait_idle();
	size_t at = 0;
	for(int32_t i = 0; i < n; i++) {
		memset(cur, 0, sizeof(cur));
		cur_freq = freq;
		struct metadata *m = hfdl_pdu_metadata_create();
		struct hfdl_pdu_metadata *hm = container_of(m, struct hfdl_pdu_metadata, metadata);
		hm->version = 1;
		hm->freq = freq;
		uint8_t *copy = malloc((size_t)lens[i] + 2);          /* hfdl.c:1056 hands over a malloc'd copy */
		memcpy(copy, bufs + at, (size_t)lens[i]);
		at += (size_t)lens[i];
		pdu_decoder_queue_push(m, octet_string_new(copy, (size_t)lens[i]), 0);
		wait_idle();
		memcpy(out + (size_t)i * K_COUNT, cur, sizeof(cur));
	}
	return wrong_freq ? -2 : 0;
}

/* the reference's own hfdl_pdu_fcs_check (pdu.c:68-79) and parse_icao_hex (util.c:236-242), for direct comparison */
int32_t ref_front_fcs_check(uint8_t *buf, uint32_t hdr_len) { return hfdl_pdu_fcs_check(buf, hdr_len) ? 1 : 0; }
