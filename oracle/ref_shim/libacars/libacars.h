/* oracle/ref_shim/libacars/libacars.h -- stand-in: the reference's statsd.h only needs the la_msg_dir type. Test infrastructure. */
#pragma once
typedef enum { LA_MSG_DIR_UNKNOWN = 0, LA_MSG_DIR_GND2AIR, LA_MSG_DIR_AIR2GND } la_msg_dir;
