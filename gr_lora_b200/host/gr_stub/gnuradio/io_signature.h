#pragma once
#include <gnuradio/sync_block.h>
