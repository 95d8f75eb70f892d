// stand-in for <gnuradio/attributes.h> (include/lora/api.h:25): symbol visibility macros only
#pragma once
#define __GR_ATTR_EXPORT __attribute__((visibility("default")))
#define __GR_ATTR_IMPORT __attribute__((visibility("default")))
