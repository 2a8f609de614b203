/* TEST INFRASTRUCTURE ONLY -- see block.h */
#include <gnuradio/block.h>
