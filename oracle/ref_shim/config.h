/* oracle/ref_shim/config.h -- stands in for the cmake-generated src/config.h (config.h.in); no optional features. */
#pragma once
#define HAVE_PTHREAD_BARRIERS
